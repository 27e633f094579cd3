from .get_env import get_env, get_vec_env, get_subprocvec_env  # noqa: F401
from .synth import SynthVecEnv, DeviceNormalizer  # noqa: F401
from .synth import SynthVecEnv as VecEnv  # noqa: F401  (the reference exports VecEnv / SubProcVecEnv)
from .synth import SynthVecEnv as SubProcVecEnv  # noqa: F401
from .synth_atari import SynthAtariVecEnv  # noqa: F401
from .bridge import HostEnvBridge  # noqa: F401
from ..hostenv import VecEnv as HostVecEnv, SubProcVecEnv as HostSubProcVecEnv  # noqa: F401
