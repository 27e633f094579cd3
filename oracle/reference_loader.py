"""Import the unmodified reference (RchalYang/torchrl): from /root/reference where it is mounted (the build
container), else from oracle/_ref -- the byte-for-byte copy oracle/build_ref.py places beside the oracle so
that it travels to the GPU box (git-ignored: reference sources never enter this repository's history).

TEST / BASELINE INFRASTRUCTURE ONLY.  ``available()`` is False when neither location exists; nothing in the
product may depend on it.

The reference imports third-party modules that are absent from this image
(gym, toolz, tensorboardX, wandb -- SURVEY.md Appendix C); oracle/shims holds
stand-ins for those and is put *ahead* of the reference on sys.path.  They are
real packages on sys.path (not sys.modules injection) because SubProcVecEnv
workers are spawned and re-import everything
(/root/reference/torchrl/env/subproc_vecenv.py:7).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("TORCHRL_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "torchrl")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()
SHIMS = os.path.join(_HERE, "shims")
REPO_ROOT = os.path.dirname(_HERE)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "torchrl"))


def add_shims():
    """Put the third-party stand-ins and the repo root on sys.path (idempotent)."""
    for p in (REPO_ROOT, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    # make sure `gym.make` knows the synthetic env ids
    import oracle.synth_env  # noqa: F401


def load():
    """Return the reference's top-level ``torchrl`` module (imports lazily)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    add_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    import torchrl  # the reference package
    assert os.path.abspath(torchrl.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), torchrl.__file__
    import torchrl.replay_buffers  # noqa: F401
    import torchrl.env  # noqa: F401
    import torchrl.collector  # noqa: F401
    import torchrl.algo  # noqa: F401
    import torchrl.policies  # noqa: F401
    import torchrl.networks  # noqa: F401
    return torchrl
