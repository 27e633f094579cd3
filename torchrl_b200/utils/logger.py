"""Experiment logger (API of /root/reference/torchrl/utils/logger.py:17-158).

Same constructor, work_dir layout (log_dir/experiment_id/env_name/seed), log.csv / params.json /
git_hash.txt files, key names and Mean/Std/Max/Min aggregation of the per-update infos.
tensorboardX, wandb and GitPython are used when importable and skipped otherwise (the reference
hard-requires all three plus params['project']).
"""
import csv
import json
import logging
import os
import shutil
import sys

import numpy as np

try:
    from tabulate import tabulate
except ImportError:  # pragma: no cover
    tabulate = None
try:
    import tensorboardX
except ImportError:
    tensorboardX = None
try:
    import wandb
except ImportError:
    wandb = None


_AGGREGATES = (("Mean", np.mean), ("Std", np.std), ("Max", np.max), ("Min", np.min))


def _console_logger(name, quiet):
    log = logging.getLogger(name)
    log.handlers = []
    log.propagate = False
    handler = logging.StreamHandler(sys.stdout)
    handler.setFormatter(logging.Formatter("%(asctime)s %(threadName)s %(levelname)s: %(message)s"))
    handler.setLevel(logging.INFO)
    log.addHandler(handler)
    log.setLevel(logging.WARNING if quiet else logging.INFO)
    return log


def _git_revision():
    try:
        import git
        return git.Repo(search_parent_directories=True).head.object.hexsha
    except Exception:
        return "unknown"


class Logger:
    """Run directory `log_dir/experiment_id/env_name/seed` holding params.json, git_hash.txt and log.csv; per-update
    scalars are buffered by `add_update_info` and folded into Mean / Std / Max / Min columns by the next
    `add_epoch_info`, which also appends the epoch's row (values printed with 5 decimals)."""

    def __init__(self, experiment_id, env_name, seed, params, log_dir="./log", overwrite=False, quiet=False):
        run = "{}_{}_{}".format(experiment_id, env_name, str(seed))
        self.logger = _console_logger(run, quiet)
        self.quiet = quiet
        self.work_dir = os.path.join(log_dir, experiment_id, env_name, str(seed))
        if os.path.exists(self.work_dir):
            assert overwrite, "Experiment Exists and Did not set overwrite"
            shutil.rmtree(self.work_dir)
        os.makedirs(self.work_dir, exist_ok=True)
        self.tf_writer = tensorboardX.SummaryWriter(self.work_dir) if tensorboardX is not None else None
        self.csv_file_path = os.path.join(self.work_dir, 'log.csv')
        self.git_file_path = os.path.join(self.work_dir, 'git_hash.txt')
        with open(self.git_file_path, 'w') as f:
            f.write(_git_revision())
        self.update_count = 0
        self.stored_infos = {}
        config = {k: v for k, v in params.items() if _jsonable(v)}
        with open(os.path.join(self.work_dir, 'params.json'), 'w') as f:
            json.dump(config, f, indent=2)
        self.logger.info("Experiment Name:{}".format(experiment_id))
        self.logger.info(json.dumps(config, indent=2))
        params["name_combine"] = "{}_{}".format(experiment_id, env_name)
        self.use_wb = wandb is not None and 'project' in params
        if self.use_wb:
            wandb.init(project=params['project'], name=run, group=params["name_combine"], config=config)

    def finish(self):
        if self.use_wb:
            wandb.finish()

    def log(self, info):
        self.logger.info(info)

    def add_update_info(self, infos):
        for key, value in infos.items():
            self.stored_infos.setdefault(key, []).append(value)
        self.update_count += 1

    def _columns(self, infos):
        """(name, value) pairs of one epoch row after the three fixed columns: the epoch scalars in the order
        given, then Mean / Std / Max / Min of every buffered per-update scalar."""
        for key, value in infos.items():
            yield key, value
        for key, values in self.stored_infos.items():
            for suffix, fold in _AGGREGATES:
                yield "{}_{}".format(key, suffix), fold(values)

    def add_epoch_info(self, epoch_num, total_frames, total_time, infos, csv_write=True):
        fixed = [("EPOCH", epoch_num), ("Time Consumed", total_time), ("Total Frames", total_frames)]
        for label, value in fixed:
            self.logger.info("{}:{}{}".format(label, value, "" if label == "EPOCH" else "s"))
        columns = list(self._columns(infos))
        if self.tf_writer is not None:
            for name, value in columns:
                self.tf_writer.add_scalar(name, value, total_frames)
        if self.use_wb:
            wandb.log(dict(fixed + columns))
        if csv_write:
            with open(self.csv_file_path, 'a') as f:
                writer = csv.writer(f)
                if epoch_num == 0:
                    writer.writerow([label for label, _ in fixed] + [name for name, _ in columns])
                writer.writerow([value for _, value in fixed] + ["{:.5f}".format(value) for _, value in columns])
        if not self.quiet:
            self._print_table(infos, columns[len(infos):])
        self.stored_infos = {}

    def _print_table(self, infos, aggregated):
        table = [["Name", "Value"]] + [[k, "{:.5f}".format(v)] for k, v in infos.items()] + [[]]
        table.append(["Name"] + [suffix for suffix, _ in _AGGREGATES])
        for i in range(0, len(aggregated), len(_AGGREGATES)):
            group = aggregated[i:i + len(_AGGREGATES)]
            key = group[0][0][:-len("_" + _AGGREGATES[0][0])]
            table.append([key] + ["{:.5f}".format(v) for _, v in group])
        if tabulate is not None:
            print(tabulate(table))
        else:
            for row in table:
                print(*row)


class NullLogger:
    """Collects update infos without any I/O (benchmarks, tests)."""

    def __init__(self):
        self.stored_infos = {}
        self.update_count = 0
        self.epochs = []
        self.work_dir = None

    def add_update_info(self, infos):
        for k, v in infos.items():
            self.stored_infos.setdefault(k, []).append(v)
        self.update_count += 1

    def add_epoch_info(self, epoch_num, total_frames, total_time, infos, csv_write=True):
        self.epochs.append((epoch_num, total_frames, dict(infos)))
        self.stored_infos = {}

    def log(self, info):
        pass

    def finish(self):
        pass


def _jsonable(v):
    try:
        json.dumps(v)
        return True
    except (TypeError, ValueError):
        return False
