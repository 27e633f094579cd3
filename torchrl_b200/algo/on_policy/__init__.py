from .a2c import A2C  # noqa: F401
from .ppo import PPO  # noqa: F401
