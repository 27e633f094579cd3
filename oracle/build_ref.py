"""Recipe for oracle/_ref: a byte-for-byte copy of the reference's own Python sources, made where they lie.

TEST / BASELINE INFRASTRUCTURE ONLY (the product never imports anything under oracle/).

The reference (RchalYang/torchrl) is pure Python: there is nothing to compile, "building" it means placing its
package (`torchrl/`), its example scripts (`examples/`) and its configs (`config/`) under oracle/_ref/ so that
they travel to the GPU box with the repo snapshot (`/root/reference` does not exist there).  oracle/_ref/ is
git-ignored -- reference sources never enter this repository's history -- and is NOT gpurun-ignored.

    python oracle/build_ref.py            # no-op when /root/reference is absent (e.g. on the GPU box)

Consumers: oracle/reference_loader.py (falls back to oracle/_ref when /root/reference is absent), the CPU arm of
bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`) and tests/test_reference_examples.py.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("TORCHRL_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
PARTS = ("torchrl", "examples", "config")


def build(verbose=False):
    """Copy SRC/{torchrl,examples,config} -> oracle/_ref/.  Returns the destination, or None if there is no source."""
    if not os.path.isdir(os.path.join(SRC, "torchrl")):
        return DST if os.path.isdir(os.path.join(DST, "torchrl")) else None
    os.makedirs(DST, exist_ok=True)
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        if not os.path.isdir(s):
            continue
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "ORIGIN.txt"), "w") as f:
        f.write("copied by oracle/build_ref.py from %s (unmodified reference sources; not part of this repository)\n" % SRC)
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print("oracle/_ref: %d files from %s" % (n, SRC))
    return DST


if __name__ == "__main__":
    out = build(verbose=True)
    sys.exit(0 if out else 1)
