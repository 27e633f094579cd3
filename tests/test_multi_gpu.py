"""K12 on real GPUs (skipped below 2 visible GPUs): 2 ranks, one process per GPU, launched exactly like bench.py's N > 1
case; scripts/dist_check.py asserts that the data-parallel run equals the single-process run over the same envs.
The reference has no multi-GPU path: the pin is "sharding changes nothing"."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(algo, comm, port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, TORCHRL_B200_COMM=comm)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "scripts", "dist_check.py"), algo],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    try:                                        # keep the ranks' full output for diagnosis
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "dist_check_%s_%s.txt" % (algo, comm)), "w") as f:
            f.write(r.stdout + "\n==== stderr ====\n" + r.stderr)
    except OSError:
        pass
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    oks = [l for l in r.stdout.splitlines() if l.startswith("dist_check ok")]
    assert len(oks) == 2, r.stdout[-3000:]                       # eager and CUDA-graph passes
    assert all(("comm=%s" % comm) in l for l in oks), oks
    return oks


@pytest.mark.gpu
@pytest.mark.parametrize("comm", ["peer", "nccl"])
def test_ppo_two_ranks_equal_single_process(comm):
    _run("ppo", comm, 29531 if comm == "peer" else 29532)


@pytest.mark.gpu
@pytest.mark.parametrize("comm", ["peer", "nccl"])
def test_sac_two_ranks_equal_single_process(comm):
    _run("sac", comm, 29533 if comm == "peer" else 29534)
