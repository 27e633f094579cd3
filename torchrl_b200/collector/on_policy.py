"""On-policy collectors (API of /root/reference/torchrl/collector/on_policy.py:8-155): adds the
value net, stores V(obs) per step and bootstraps rewards of envs cut by `max_episode_frames`."""
from .base import VecCollector


class VecOnPolicyCollector(VecCollector):
    on_policy = True

    def __init__(self, vf, discount=0.99, **kwargs):
        self.vf = vf
        self.discount = discount
        super().__init__(**kwargs)

    @property
    def funcs(self):
        return {"pf": self.pf, "vf": self.vf}


OnPolicyCollectorBase = VecOnPolicyCollector
