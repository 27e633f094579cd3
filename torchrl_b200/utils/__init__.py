from .args import get_args, get_params  # noqa: F401
from .logger import Logger, NullLogger  # noqa: F401
from .checkpoint import save_checkpoint, load_checkpoint  # noqa: F401
