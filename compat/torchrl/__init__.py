"""`torchrl` -- the reference's package name -- served by torchrl_b200.

Put this directory ahead of the reference on sys.path and the reference's own example scripts
(/root/reference/examples/*.py: `from torchrl.utils import get_args`, `import torchrl.policies as policies`,
`from torchrl.collector.on_policy import VecOnPolicyCollector`, ...) run UNMODIFIED on the device path:

    PYTHONPATH=compat:. python <reference>/examples/ppo_continuous_vec.py \
        --config config/ppo_synth_halfcheetah.json --vec_env_nums 4096 --seed 0

Every `torchrl.<sub>` module is the `torchrl_b200.<sub>` module object itself (no copies): the API mirror is what
the examples exercise (tests/test_reference_examples.py).
"""
import importlib
import sys

import torchrl_b200 as _impl

_SUBMODULES = ("utils", "env", "env.get_env", "collector", "collector.base", "collector.on_policy", "replay_buffers",
               "replay_buffers.base", "replay_buffers.on_policy", "policies", "policies.continuous_policy",
               "policies.discrete_policies", "policies.distribution", "networks", "networks.base", "networks.nets",
               "networks.init", "algo", "algo.utils")

for _name in _SUBMODULES:
    try:
        _mod = importlib.import_module("torchrl_b200." + _name)
    except ImportError:
        continue
    sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod

__version__ = getattr(_impl, "__version__", "0")
