from .base import BaseCollector, VecCollector  # noqa: F401
from .on_policy import OnPolicyCollectorBase, VecOnPolicyCollector  # noqa: F401
from .pixel import PixelVecCollector  # noqa: F401
