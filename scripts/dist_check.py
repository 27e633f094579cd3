"""2-rank NCCL equivalence check (run under torchrun --nproc-per-node 2 on a 2-GPU box):
PPO with envs sharded over 2 ranks == single-process PPO over the same 2N envs, given the same
exploration noise (drawn from the CPU generator for ALL envs and sliced per rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_check.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchrl_b200.networks as networks  # noqa: E402
import torchrl_b200.policies as policies  # noqa: E402
from torchrl_b200.algo import PPO  # noqa: E402
from torchrl_b200.collector import VecOnPolicyCollector  # noqa: E402
from torchrl_b200.distributed import DataParallelContext  # noqa: E402
from torchrl_b200.env import get_vec_env  # noqa: E402
from torchrl_b200.policies import distribution as D  # noqa: E402
from torchrl_b200.replay_buffers import OnPolicyReplayBuffer  # noqa: E402
from torchrl_b200.utils import NullLogger  # noqa: E402

N_TOTAL, T, HID, ROWS, OE = 128, 16, (32, 32), 4, 2


def build(n_local, first, total, dev, ctx, use_graph):
    params = {"reward_scale": 1, "obs_norm": True}
    env = get_vec_env("SynthHalfCheetah-v0", params, n_local, device=dev, first_env=first, total_envs=total)
    eval_env = get_vec_env("SynthHalfCheetah-v0", params, n_local, device=dev, first_env=first, total_envs=total)
    env.dist = ctx
    env.seed(0); torch.manual_seed(0); np.random.seed(0)
    buf = OnPolicyReplayBuffer(env_nums=n_local, max_replay_buffer_size=T * n_local, time_limit_filter=True)
    net = dict(hidden_shapes=list(HID), append_hidden_shapes=[], base_type=networks.MLPBase,
               activation_func=torch.nn.Tanh)
    pf = policies.GuassianContPolicyBasicBias(input_shape=17, output_shape=6, tanh_action=True, **net)
    vf = networks.Net(input_shape=(17,), output_shape=1, **net)
    col = VecOnPolicyCollector(vf, env=env, eval_env=eval_env, pf=pf, replay_buffer=buf, device=dev,
                               epoch_frames=T * n_local, max_episode_frames=7, use_cuda_graph=use_graph)
    agent = PPO(pf=pf, vf=vf, plr=3e-4, vlr=3e-4, opt_epochs=OE, tau=0.95, shuffle=True, entropy_coeff=0.005,
                env=env, replay_buffer=buf, collector=col, logger=NullLogger(), discount=0.99, num_epochs=10,
                batch_size=ROWS * n_local, gae=True, device=dev, save_dir=None, use_cuda_graph=use_graph, dist=ctx)
    return agent, col, buf, env


def run(agent, col, epochs=2):
    for e in range(epochs):
        agent.current_epoch = e
        col.train_one_epoch()
        agent.update_per_epoch()


def main():
    ctx = DataParallelContext()
    assert ctx.world_size == 2, "run under torchrun with 2 ranks"
    dev = ctx.device
    first, n_local = ctx.shard(N_TOTAL)
    D.set_noise_mode("reference_cpu")
    full_draw = D.draw_reference_noise

    def sliced(shape, device):       # the single process draws (N_TOTAL, a); a rank uses its rows
        full = torch.normal(torch.zeros((N_TOTAL,) + tuple(shape[1:])), torch.ones((N_TOTAL,) + tuple(shape[1:])))
        return full[first:first + shape[0]].to(device)

    for use_graph in (False, True):
        D.draw_reference_noise = sliced
        agent, col, buf, env = build(n_local, first, N_TOTAL, dev, ctx, use_graph)
        run(agent, col)
        flat = agent.opt.data.clone()
        other = [torch.zeros_like(flat) for _ in range(2)]
        torch.distributed.all_gather(other, flat)
        assert torch.equal(other[0], other[1]), "ranks diverged"
        nrm = env._obs_normalizer
        if ctx.rank == 0:
            D.draw_reference_noise = full_draw
            a1, c1, b1, e1 = build(N_TOTAL, 0, N_TOTAL, dev, None, use_graph)
            run(a1, c1)
            torch.testing.assert_close(flat, a1.opt.data, rtol=1e-3, atol=2e-5)
            torch.testing.assert_close(nrm._mean, e1._obs_normalizer._mean, rtol=1e-6, atol=1e-8)
            torch.testing.assert_close(nrm._var, e1._obs_normalizer._var, rtol=1e-6, atol=1e-8)
            assert abs(nrm._count.item() - e1._obs_normalizer._count.item()) < 1e-9
            torch.testing.assert_close(buf._advs, b1._advs[:, first:first + n_local], rtol=1e-3, atol=2e-4)
            torch.testing.assert_close(buf._obs, b1._obs[:, first:first + n_local], rtol=1e-4, atol=1e-4)
            print("dist_check ok (graph=%s): 2-rank == single-process, max |dparam| = %.3g"
                  % (use_graph, (flat - a1.opt.data).abs().max().item()), flush=True)
        ctx.barrier()
    ctx.destroy()


if __name__ == "__main__":
    main()
