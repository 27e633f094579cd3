"""Stand-in for the third-party `gym` module the reference's example scripts import at top level
(`import gym`; the examples never call it -- environments come from `torchrl.env.get_vec_env`).  Only what a
script or a user-side isinstance check can touch is provided: the space classes of torchrl_b200.spaces."""
from . import spaces  # noqa: F401


class Env:
    pass


def make(env_id, **kwargs):
    raise RuntimeError("gym.make(%r): this stand-in only exists so that `import gym` succeeds; device environments are "
                       "created by torchrl.env.get_vec_env (SynthHalfCheetah-v0, SynthAnt-v0, SynthAtari-v0), real host "
                       "environments by torchrl_b200.hostenv" % (env_id,))
