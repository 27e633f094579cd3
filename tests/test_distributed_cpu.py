"""K12 host-side logic on CPU: world_size-2 `gloo` processes exercise the same DataParallelContext the
NCCL path uses (sharding, gradient all-reduce + 1/G scale, global advantage statistics, identical
minibatch row order on every rank)."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def test_shard_range_and_moments():
    from torchrl_b200.distributed import combine_moments, shard_range
    assert shard_range(32768, 8, 3) == (3 * 4096, 4096)
    with pytest.raises(ValueError):
        shard_range(10, 4, 0)
    x = torch.randn(1000, dtype=torch.float64)
    mean, std = combine_moments(x.sum(), (x * x).sum(), torch.tensor(1000.0, dtype=torch.float64))
    assert abs(mean - x.mean()) < 1e-12 and abs(std - x.std()) < 1e-12


def _worker(rank, world, port, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from torchrl_b200.distributed import DataParallelContext
    ctx = DataParallelContext(backend="gloo", device="cpu")
    assert ctx.active and ctx.shard(64) == (rank * 32, 32)
    # gradient exchange: SUM then scale 1/G == mean of the per-rank gradients
    g = torch.full((1000,), float(rank + 1))
    scale = ctx.all_reduce_grads(g)
    assert scale == 0.5 and torch.allclose(g * scale, torch.full((1000,), 1.5))
    # advantage statistics over the union of the ranks' minibatches (ppo.py:147 semantics)
    torch.manual_seed(0)
    full = torch.randn(2, 512) * 3 + 1
    out = torch.zeros(4)
    ctx.global_vec_stats(full[rank], out)
    ref = full.reshape(-1)
    exp = torch.stack([ref.mean(), ref.std(), ref.max(), ref.min()])
    assert torch.allclose(out, exp, rtol=1e-5, atol=1e-6), (out, exp)
    # every rank draws the same row order from the same host seed
    np.random.seed(7)
    perm = torch.from_numpy(np.random.permutation(128))
    gathered = [torch.zeros_like(perm) for _ in range(world)]
    torch.distributed.all_gather(gathered, perm)
    assert all(torch.equal(gathered[0], p) for p in gathered)
    assert abs(ctx.max_over_ranks(float(rank)) - (world - 1)) < 1e-12
    ctx.barrier()
    ctx.destroy()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("1")


def test_two_rank_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
