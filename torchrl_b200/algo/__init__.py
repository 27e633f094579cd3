from .on_policy import A2C, PPO, TRPO, VMPO  # noqa: F401
from .off_policy import TwinSACQ, TD3, DQN, QRDQN, DDPG  # noqa: F401
from .rl_algo import RLAlgo  # noqa: F401

__all__ = ['TwinSACQ', 'TD3', 'DQN', 'QRDQN', 'DDPG', 'A2C', 'PPO', 'TRPO', 'VMPO', 'RLAlgo']
