"""Env factories with the reference's names (/root/reference/torchrl/env/get_env.py:32-87).

Synthetic ids ("SynthHalfCheetah-v0", "SynthAnt-v0") build the device-resident
SynthVecEnv; any other id needs a real gym + the host-env bridge, which is
outside this round's hot path and raises.
"""
import torch

from . import synth_spec
from .synth import SynthVecEnv
from .synth_atari import SynthAtariVecEnv, ENV_ID as ATARI_ID


def _device(device):
    if device is not None:
        return device
    if not torch.cuda.is_available():
        raise RuntimeError("torchrl_b200 envs live on the GPU: no CUDA device available (there is no CPU path)")
    return "cuda"


def get_vec_env(env_id, env_param, vec_env_nums, device=None, **kwargs):
    if synth_spec.is_synth(env_id):
        return SynthVecEnv(env_id, vec_env_nums, env_param, device=_device(device), **kwargs)
    if env_id == ATARI_ID:
        return SynthAtariVecEnv(vec_env_nums, env_param, device=_device(device), **kwargs)
    raise NotImplementedError("only the synthetic device envs are built in this round: %r" % (env_id,))


def get_subprocvec_env(env_id, env_param, vec_env_nums, proc_nums, device=None, **kwargs):
    """`proc_nums` is accepted for API compatibility: the device env needs no worker processes."""
    return get_vec_env(env_id, env_param, vec_env_nums, device=device, **kwargs)


def get_env(env_id, env_param, device=None):
    return get_vec_env(env_id, env_param, 1, device=device)
