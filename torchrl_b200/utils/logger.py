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


class Logger:
    def __init__(self, experiment_id, env_name, seed, params, log_dir="./log", overwrite=False, quiet=False):
        self.logger = logging.getLogger("{}_{}_{}".format(experiment_id, env_name, str(seed)))
        self.logger.handlers = []
        self.logger.propagate = False
        sh = logging.StreamHandler(sys.stdout)
        sh.setFormatter(logging.Formatter("%(asctime)s %(threadName)s %(levelname)s: %(message)s"))
        sh.setLevel(logging.INFO)
        self.logger.addHandler(sh)
        self.logger.setLevel(logging.WARNING if quiet else logging.INFO)
        self.quiet = quiet

        work_dir = os.path.join(log_dir, experiment_id, env_name, str(seed))
        self.work_dir = work_dir
        if os.path.exists(work_dir):
            assert overwrite, "Experiment Exists and Did not set overwrite"
            shutil.rmtree(work_dir)
        os.makedirs(work_dir, exist_ok=True)
        self.tf_writer = tensorboardX.SummaryWriter(work_dir) if tensorboardX is not None else None
        self.csv_file_path = os.path.join(work_dir, 'log.csv')
        self.git_file_path = os.path.join(work_dir, 'git_hash.txt')
        try:
            import git
            sha = git.Repo(search_parent_directories=True).head.object.hexsha
        except Exception:
            sha = "unknown"
        with open(self.git_file_path, 'w') as f:
            f.write(sha)
        self.update_count = 0
        self.stored_infos = {}
        serialisable = {k: v for k, v in params.items() if _jsonable(v)}
        with open(os.path.join(work_dir, 'params.json'), 'w') as f:
            json.dump(serialisable, f, indent=2)
        self.logger.info("Experiment Name:{}".format(experiment_id))
        self.logger.info(json.dumps(serialisable, indent=2))
        params["name_combine"] = "{}_{}".format(experiment_id, env_name)
        self.use_wb = wandb is not None and 'project' in params
        if self.use_wb:
            wandb.init(project=params['project'], name="{}_{}_{}".format(experiment_id, env_name, str(seed)),
                       group="{}_{}".format(experiment_id, env_name), config=serialisable)

    def finish(self):
        if self.use_wb:
            wandb.finish()

    def log(self, info):
        self.logger.info(info)

    def add_update_info(self, infos):
        for k, v in infos.items():
            self.stored_infos.setdefault(k, []).append(v)
        self.update_count += 1

    def add_epoch_info(self, epoch_num, total_frames, total_time, infos, csv_write=True):
        if csv_write:
            csv_titles = ["EPOCH", "Time Consumed", "Total Frames"] if epoch_num == 0 else None
            csv_values = [epoch_num, total_time, total_frames]
        self.logger.info("EPOCH:{}".format(epoch_num))
        self.logger.info("Time Consumed:{}s".format(total_time))
        self.logger.info("Total Frames:{}s".format(total_frames))
        table = [["Name", "Value"]]
        wb = {"EPOCH": epoch_num, "Time Consumed": total_time, "Total Frames": total_frames}
        for k, v in infos.items():
            if self.tf_writer is not None:
                self.tf_writer.add_scalar(k, v, total_frames)
            table.append([k, "{:.5f}".format(v)])
            if csv_write:
                if csv_titles is not None:
                    csv_titles.append(k)
                csv_values.append("{:.5f}".format(v))
            wb[k] = v
        table.append([])
        names, methods = ["Mean", "Std", "Max", "Min"], [np.mean, np.std, np.max, np.min]
        table.append(["Name"] + names)
        for k, vals in self.stored_infos.items():
            row = [k]
            for name, method in zip(names, methods):
                val = method(vals)
                if self.tf_writer is not None:
                    self.tf_writer.add_scalar("{}_{}".format(k, name), val, total_frames)
                row.append("{:.5f}".format(val))
                if csv_write:
                    if csv_titles is not None:
                        csv_titles.append("{}_{}".format(k, name))
                    csv_values.append("{:.5f}".format(val))
                wb["{}_{}".format(k, name)] = val
            table.append(row)
        if self.use_wb:
            wandb.log(wb)
        self.stored_infos = {}
        if csv_write:
            with open(self.csv_file_path, 'a') as f:
                w = csv.writer(f)
                if csv_titles is not None:
                    w.writerow(csv_titles)
                w.writerow(csv_values)
        if not self.quiet:
            if tabulate is not None:
                print(tabulate(table))
            else:
                for r in table:
                    print(*r)


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
