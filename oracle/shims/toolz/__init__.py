"""Stand-in for the third-party `toolz` package: only dicttoolz.merge_with is used."""
from . import dicttoolz  # noqa: F401
