from .base import BaseReplayBuffer  # noqa: F401
from .on_policy import OnPolicyReplayBuffer, OnPolicyReplayBufferBase  # noqa: F401
from .prioritized import PrioritizedReplayBuffer  # noqa: F401
from .memory_efficient import MemoryEfficientReplayBuffer  # noqa: F401
