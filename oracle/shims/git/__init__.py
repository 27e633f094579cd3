"""Stand-in for GitPython (absent from the image): the reference's Logger only reads
``git.Repo(search_parent_directories=True).head.object.hexsha`` (torchrl/utils/logger.py:50-53).
Not reference code; used only by oracle/reference_loader.py."""


class _Obj:
    hexsha = "0" * 40


class _Head:
    object = _Obj()


class Repo:
    def __init__(self, *args, **kwargs):
        self.head = _Head()
