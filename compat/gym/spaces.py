from torchrl_b200.spaces import Box, Discrete  # noqa: F401
