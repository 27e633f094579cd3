"""Command-line flags and JSON config loader (API of /root/reference/torchrl/utils/args.py:6-53): the flag
names, types and defaults the reference's example scripts read from the returned namespace."""
import argparse
import json

import torch

# (flag, type or None for a store_true switch, default, help)
_FLAGS = (
    ("seed", int, 0, "random seed"),
    ("vec_env_nums", int, 4, "number of vectorised envs"),
    ("proc_nums", int, 4, "env worker processes (host envs; device envs need none)"),
    ("eval_worker_nums", int, 2, "evaluation workers"),
    ("config", str, None, "JSON config file"),
    ("save_dir", str, "./snapshots", "directory for snapshots"),
    ("log_dir", str, "./log", "directory for logs"),
    ("no_cuda", None, False, "disable CUDA (the agents of this package then refuse to start)"),
    ("overwrite", None, False, "overwrite a previous experiment with the same id"),
    ("device", int, 0, "GPU index"),
    ("id", str, None, "experiment id"),
)


def build_parser():
    parser = argparse.ArgumentParser(description="RL")
    for name, kind, default, text in _FLAGS:
        if kind is None:
            parser.add_argument("--" + name, action="store_true", default=default, help=text)
        else:
            parser.add_argument("--" + name, type=kind, default=default, help=text)
    return parser


def get_args(argv=None):
    args = build_parser().parse_args(argv)
    args.cuda = torch.cuda.is_available() and not args.no_cuda
    return args


def get_params(file_name):
    with open(file_name) as handle:
        return json.load(handle)
