"""2-rank data-parallel equivalence check (run under torchrun --nproc-per-node 2 on a 2-GPU box; driven by
tests/test_multi_gpu.py):

  ppo : PPO with envs sharded over 2 ranks == single-process PPO over the same 2N envs (rollout, advantages, the
        observation-normaliser state and the parameters after two epochs), eager and CUDA-graph paths;
  sac : TwinSAC-Q with the replay ring sharded by env over 2 ranks == the single-process agent over all envs
        (parameters, target networks, log-alpha after three epochs).

Both cases draw the exploration / reparameterisation noise from the CPU generator for ALL envs and slice it per rank,
so the two runs see the same numbers.  TORCHRL_B200_COMM=peer (default: csrc/comm.cu over NVLink peer memory) or nccl.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_check.py [ppo|sac]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchrl_b200.networks as networks  # noqa: E402
import torchrl_b200.policies as policies  # noqa: E402
from torchrl_b200.algo import PPO, TwinSACQ  # noqa: E402
from torchrl_b200.collector import VecCollector, VecOnPolicyCollector  # noqa: E402
from torchrl_b200.distributed import DataParallelContext  # noqa: E402
from torchrl_b200.env import get_vec_env  # noqa: E402
from torchrl_b200.policies import distribution as D  # noqa: E402
from torchrl_b200.replay_buffers import BaseReplayBuffer, OnPolicyReplayBuffer  # noqa: E402
from torchrl_b200.utils import NullLogger  # noqa: E402

N_TOTAL, T, HID, ROWS, OE = 128, 16, (32, 32), 4, 2


def build_ppo(n_local, first, total, dev, ctx, use_graph):
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


def build_sac(n_local, first, total, dev, ctx, use_graph):
    env = get_vec_env("SynthAnt-v0", {"reward_scale": 1, "obs_norm": False}, n_local, device=dev, first_env=first,
                      total_envs=total)
    eval_env = get_vec_env("SynthAnt-v0", {"reward_scale": 1, "obs_norm": False}, n_local, device=dev, first_env=first,
                           total_envs=total)
    env.dist = ctx
    env.seed(0); torch.manual_seed(0); np.random.seed(0)
    o, a = env.observation_space.shape[0], env.action_space.shape[0]
    buf = BaseReplayBuffer(env_nums=n_local, max_replay_buffer_size=64 * n_local, time_limit_filter=False)
    net = dict(hidden_shapes=[32, 32], append_hidden_shapes=[], base_type=networks.MLPBase, activation_func=torch.nn.ReLU)
    pf = policies.GuassianContPolicy(input_shape=o, output_shape=2 * a, tanh_action=True, **net)
    qf1 = networks.QNet(input_shape=o + a, output_shape=1, **net)
    qf2 = networks.QNet(input_shape=o + a, output_shape=1, **net)
    col = VecCollector(env=env, eval_env=eval_env, pf=pf, replay_buffer=buf, device=dev, epoch_frames=8 * n_local,
                       max_episode_frames=20, use_cuda_graph=use_graph)
    agent = TwinSACQ(pf=pf, qf1=qf1, qf2=qf2, plr=3e-4, qlr=3e-4, policy_std_reg_weight=1e-3, policy_mean_reg_weight=1e-3,
                     env=env, replay_buffer=buf, collector=col, logger=NullLogger(), discount=0.99,
                     batch_size=ROWS * n_local, device=dev, save_dir=None, tau=0.005, use_soft_update=True, opt_times=6,
                     pretrain_epochs=1, num_epochs=3, use_cuda_graph=use_graph, dist=ctx)
    return agent, col, buf, env


def run(agent, col, epochs, pretrain=False):
    if pretrain:
        agent.pretrain()
    for e in range(epochs):
        agent.current_epoch = e
        col.train_one_epoch()
        agent.update_per_epoch()


def main():
    algo = sys.argv[1] if len(sys.argv) > 1 else "ppo"
    ctx = DataParallelContext()
    assert ctx.world_size == 2, "run under torchrun with 2 ranks"
    dev = ctx.device
    first, n_local = ctx.shard(N_TOTAL)
    D.set_noise_mode("reference_cpu")
    full_draw = D.draw_reference_noise

    def sliced(shape, device):
        """What a single process over all N_TOTAL envs would draw for a (k rows x envs, ...) batch, cut down to this
        rank's envs: batches are laid out (rows, envs, ...), a rank owns envs [first, first + n_local)."""
        k = shape[0] // n_local
        tail = tuple(shape[1:])
        full = torch.normal(torch.zeros((k * N_TOTAL,) + tail), torch.ones((k * N_TOTAL,) + tail))
        mine = full.reshape((k, N_TOTAL) + tail)[:, first:first + n_local]
        return mine.reshape((k * n_local,) + tail).to(device)

    build = build_ppo if algo == "ppo" else build_sac
    epochs = 2 if algo == "ppo" else 3
    for use_graph in (False, True):
        # SAC draws its reparameterisation noise inside the update: with the CPU reference generator that cannot be
        # captured, so its CUDA-graph pass runs on the device Philox streams (one per rank) and checks what does not
        # depend on the noise source: the ranks stay bit-identical
        philox = (algo == "sac" and use_graph)
        D.set_noise_mode("philox" if philox else "reference_cpu")
        D.draw_reference_noise = sliced
        agent, col, buf, env = build(n_local, first, N_TOTAL, dev, ctx, use_graph)
        run(agent, col, epochs, pretrain=(algo == "sac"))
        flat = agent.opt.data.clone()
        other = [torch.zeros_like(flat) for _ in range(2)]
        torch.distributed.all_gather(other, flat)
        assert torch.equal(other[0], other[1]), "ranks diverged"
        assert bool(torch.isfinite(flat).all())
        if philox:
            if ctx.rank == 0:
                print("dist_check ok (%s, graph=%s, comm=%s): ranks bit-identical after %d epochs (device noise)"
                      % (algo, use_graph, "peer" if ctx.peer is not None else "nccl", epochs), flush=True)
        elif ctx.rank == 0:
            D.draw_reference_noise = full_draw
            a1, c1, b1, e1 = build(N_TOTAL, 0, N_TOTAL, dev, None, use_graph)
            run(a1, c1, epochs, pretrain=(algo == "sac"))
            torch.testing.assert_close(flat, a1.opt.data, rtol=1e-3, atol=2e-5)
            if algo == "ppo":
                nrm = env._obs_normalizer
                torch.testing.assert_close(nrm._mean, e1._obs_normalizer._mean, rtol=1e-6, atol=1e-8)
                torch.testing.assert_close(nrm._var, e1._obs_normalizer._var, rtol=1e-6, atol=1e-8)
                assert abs(nrm._count.item() - e1._obs_normalizer._count.item()) < 1e-9
                torch.testing.assert_close(buf._advs, b1._advs[:, first:first + n_local], rtol=1e-3, atol=2e-4)
                torch.testing.assert_close(buf._obs, b1._obs[:, first:first + n_local], rtol=1e-4, atol=1e-4)
                for i0, i1 in zip(agent._last_infos, a1._last_infos):
                    for key in ("grad_norm/pf", "grad_norm/vf", "advs/mean", "advs/std"):
                        assert abs(i0[key] - i1[key]) <= 2e-3 * abs(i1[key]) + 1e-5, (key, i0[key], i1[key])
            else:
                torch.testing.assert_close(agent._target_flat.data, a1._target_flat.data, rtol=1e-3, atol=2e-5)
                torch.testing.assert_close(agent.log_alpha, a1.log_alpha, rtol=1e-3, atol=1e-5)
                torch.testing.assert_close(buf._obs[:buf._size], b1._obs[:b1._size, first:first + n_local], rtol=1e-4,
                                           atol=1e-4)
            print("dist_check ok (%s, graph=%s, comm=%s): 2-rank == single-process, max |dparam| = %.3g"
                  % (algo, use_graph, "peer" if ctx.peer is not None else "nccl",
                     (flat - a1.opt.data).abs().max().item()), flush=True)
        ctx.barrier()
    # hard exit: graphs that captured collectives make an orderly teardown hang (see bench.py)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
