"""The example scripts (same flags / JSON schema / wiring as the reference's examples) run end to end:
agent.train() with collection, updates, evaluation, logging (log.csv) and snapshots (model_*.pth,
_obs_normalizer_*.pkl -- the reference's snapshot file names, algo/rl_algo.py:83-94)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, cfg_src, patch, nenv, tmp_path):
    cfg = json.load(open(os.path.join(ROOT, "config", cfg_src)))
    patch(cfg)
    cfg_path = tmp_path / cfg_src
    json.dump(cfg, open(cfg_path, "w"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), "--config", str(cfg_path),
                        "--vec_env_nums", str(nenv), "--seed", "1", "--log_dir", str(tmp_path / "log"), "--overwrite"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return tmp_path / "log" / os.path.splitext(cfg_src)[0] / cfg["env_name"] / "1"


@pytest.mark.gpu
def test_ppo_example_trains_and_snapshots(tmp_path):
    def patch(c):
        c["replay_buffer"]["size"] = 64 * 16
        c["collector"].update(epoch_frames=64 * 16, max_episode_frames=40)
        c["general_setting"].update(num_epochs=3, batch_size=64 * 4, eval_interval=2, save_interval=2)
        c["net"]["hidden_shapes"] = [32, 32]
        c["ppo"]["opt_epochs"] = 2
    work = _run("ppo_continuous_vec.py", "ppo_synth_halfcheetah.json", patch, 64, tmp_path)
    assert (work / "log.csv").exists() and (work / "params.json").exists()
    files = set(os.listdir(work / "model"))
    for f in ("model_pf_best.pth", "model_vf_0.pth", "model_pf_finish.pth", "_obs_normalizer_finish.pkl"):
        assert f in files, files
    header = open(work / "log.csv").readline()
    for key in ("Train_Epoch_Reward", "Training/policy_loss_Mean", "grad_norm/pf_Max", "eval_traj_length"):
        assert key in header, header


@pytest.mark.gpu
def test_sac_example_trains(tmp_path):
    def patch(c):
        c["replay_buffer"]["size"] = 32 * 64
        c["collector"].update(epoch_frames=32 * 8, max_episode_frames=30)
        c["general_setting"].update(num_epochs=2, batch_size=32 * 4, opt_times=5, eval_interval=1, save_interval=1)
        c["net"]["hidden_shapes"] = [32, 32]
    work = _run("twin_sac_q_continuous_vec.py", "twin_sac_q_synth_ant.json", patch, 32, tmp_path)
    files = set(os.listdir(work / "model"))
    assert "model_qf1_finish.pth" in files and "model_pf_best.pth" in files
    assert "Alpha_Mean" in open(work / "log.csv").readline()
