"""Offline no-op stand-in for wandb."""


def init(*args, **kwargs):
    return None


def log(*args, **kwargs):
    return None


def finish(*args, **kwargs):
    return None
