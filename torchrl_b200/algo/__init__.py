from .on_policy import A2C, PPO  # noqa: F401
from .rl_algo import RLAlgo  # noqa: F401

__all__ = ['A2C', 'PPO', 'RLAlgo']
