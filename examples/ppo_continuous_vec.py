"""PPO on a device-resident vectorised env -- the torchrl_b200 counterpart of the reference's
examples/ppo_continuous_vec.py (same flags, same JSON schema).

    python examples/ppo_continuous_vec.py --config config/ppo_synth_halfcheetah.json --vec_env_nums 4096 --seed 0
"""
import torch

from _common import Run, main  # noqa: F401  (also puts the repository root on sys.path)
import torchrl_b200.networks as networks
import torchrl_b200.policies as policies
from torchrl_b200.algo import PPO
from torchrl_b200.collector.on_policy import VecOnPolicyCollector
from torchrl_b200.replay_buffers.on_policy import OnPolicyReplayBuffer


def experiment(run):
    cfg = run.params
    trunk = dict(cfg["net"], base_type=networks.MLPBase, activation_func=torch.nn.Tanh)
    pf = policies.GuassianContPolicyBasicBias(input_shape=run.obs_dim, output_shape=run.act_dim, **trunk,
                                              **cfg["policy"])
    vf = networks.Net(input_shape=run.env.observation_space.shape, output_shape=1, **trunk)
    rollout = OnPolicyReplayBuffer(**run.buffer_kwargs())
    collector = VecOnPolicyCollector(vf, **run.collector_kwargs(pf, rollout))
    PPO(pf=pf, vf=vf, **cfg["ppo"], **run.agent_kwargs(rollout, collector)).train()


if __name__ == "__main__":
    main(experiment)
