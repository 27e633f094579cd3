"""bench.py's reference arm (`--impl reference`: the reference's own CPU pipeline -- the unmodified classes when a copy
of the reference is present, else the oracle port -- timed for whole epochs on the host cores) prints ONE JSON line
with the contract's keys.  CPU-only; a reduced env count keeps it at a few seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--envs-per-gpu", "64", "--cpu-procs", "4"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "64 envs/GPU" in d["config"]["workload"] and "horizon 128" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    from oracle import reference_loader
    assert cb["kind"] == ("reference" if reference_loader.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # every step is a whole epoch timed for real: the per-epoch times add up to steps * ms_per_step
    assert len(cb["detail"]["epoch_s"]) == 2
    assert abs(sum(cb["detail"]["epoch_s"]) * 1e3 - d["steps"] * d["ms_per_step"]) < 1e-6 * d["ms_per_step"] + 1e-3
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
