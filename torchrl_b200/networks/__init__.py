from .nets import *  # noqa: F401,F403
from .nets import Net, FlattenNet, QNet, ZeroNet, BootstrappedNet, FlattenBootstrappedNet  # noqa: F401
from .base import MLPBase, CNNBase, calc_next_shape  # noqa: F401
from .init import basic_init, uniform_init, orthogonal_init, layer_init  # noqa: F401
