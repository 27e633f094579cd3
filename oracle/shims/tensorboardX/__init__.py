"""No-op stand-in for tensorboardX (absent here): SummaryWriter that records nothing."""


class SummaryWriter:
    def __init__(self, logdir=None, **kwargs):
        import os
        if logdir:
            os.makedirs(logdir, exist_ok=True)
        self.logdir = logdir

    def add_scalar(self, *args, **kwargs):
        pass

    def close(self):
        pass
