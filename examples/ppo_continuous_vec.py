"""PPO on a device-resident vectorised env -- the torchrl_b200 counterpart of the reference's
examples/ppo_continuous_vec.py (same flags, same JSON schema, same object wiring).

    python examples/ppo_continuous_vec.py --config config/ppo_synth_halfcheetah.json --vec_env_nums 4096 --seed 0
"""
import os
import os.path as osp
import random
import sys

import numpy as np
import torch

sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torchrl_b200.networks as networks  # noqa: E402
import torchrl_b200.policies as policies  # noqa: E402
from torchrl_b200.algo import PPO  # noqa: E402
from torchrl_b200.collector.on_policy import VecOnPolicyCollector  # noqa: E402
from torchrl_b200.env import get_vec_env  # noqa: E402
from torchrl_b200.replay_buffers.on_policy import OnPolicyReplayBuffer  # noqa: E402
from torchrl_b200.utils import Logger, get_args, get_params  # noqa: E402


def experiment(args, params):
    if not args.cuda:
        raise SystemExit("torchrl_b200 needs a CUDA device (there is no CPU path)")
    device = torch.device("cuda:{}".format(args.device))
    env = get_vec_env(params["env_name"], params["env"], args.vec_env_nums, device=device)
    eval_env = get_vec_env(params["env_name"], params["env"], args.vec_env_nums, device=device)
    env.seed(args.seed)
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    random.seed(args.seed)
    torch.cuda.manual_seed_all(args.seed)

    experiment_name = os.path.split(os.path.splitext(args.config)[0])[-1] if args.id is None else args.id
    logger = Logger(experiment_name, params['env_name'], args.seed, params, args.log_dir, args.overwrite)
    general = dict(params['general_setting'])
    buffer_param = params['replay_buffer']
    replay_buffer = OnPolicyReplayBuffer(env_nums=args.vec_env_nums, max_replay_buffer_size=int(buffer_param['size']),
                                         time_limit_filter=buffer_param['time_limit_filter'])
    net = dict(params['net'], base_type=networks.MLPBase, activation_func=torch.nn.Tanh)
    pf = policies.GuassianContPolicyBasicBias(input_shape=env.observation_space.shape[0],
                                              output_shape=env.action_space.shape[0], **net, **params['policy'])
    vf = networks.Net(input_shape=env.observation_space.shape, output_shape=1, **net)
    collector = VecOnPolicyCollector(vf, env=env, eval_env=eval_env, pf=pf, replay_buffer=replay_buffer,
                                     device=device, train_render=False, **params["collector"])
    general.update(env=env, replay_buffer=replay_buffer, logger=logger, device=device, collector=collector,
                   save_dir=osp.join(logger.work_dir, "model"))
    agent = PPO(pf=pf, vf=vf, **params["ppo"], **general)
    agent.train()


if __name__ == "__main__":
    _args = get_args()
    experiment(_args, get_params(_args.config))
