"""Twin-Q soft actor-critic on a device-resident vectorised env -- the torchrl_b200 counterpart of the reference's
examples/twin_sac_q_continuous_vec.py (same flags, same JSON schema).

    python examples/twin_sac_q_continuous_vec.py --config config/twin_sac_q_synth_ant.json --vec_env_nums 1024
"""
import torch

from _common import Run, main  # noqa: F401  (also puts the repository root on sys.path)
import torchrl_b200.networks as networks
import torchrl_b200.policies as policies
from torchrl_b200.algo import TwinSACQ
from torchrl_b200.collector import VecCollector
from torchrl_b200.replay_buffers import BaseReplayBuffer


def experiment(run):
    cfg = run.params
    trunk = dict(cfg["net"], base_type=networks.MLPBase, activation_func=torch.nn.ReLU)
    o, a = run.obs_dim, run.act_dim
    pf = policies.GuassianContPolicy(input_shape=o, output_shape=2 * a, **trunk, **cfg["policy"])
    critics = [networks.QNet(input_shape=o + a, output_shape=1, **trunk) for _ in range(2)]
    ring = BaseReplayBuffer(**run.buffer_kwargs())
    collector = VecCollector(**run.collector_kwargs(pf, ring))
    TwinSACQ(pf=pf, qf1=critics[0], qf2=critics[1], **cfg["twin_sac_q"], **run.agent_kwargs(ring, collector)).train()


if __name__ == "__main__":
    main(experiment)
