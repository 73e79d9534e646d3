"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: shard bounds, global-sample-index keying, final all_gather."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusionmodelscustom_b200 import sharding


def test_shard_ranges_cover_batch_exactly():
    for n in (1, 2, 7, 64, 256, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeDiffusion:
    """Stands in for DiffusionUtils on CPU: 'samples' are a pure function of the GLOBAL sample index and the seed,
    exactly the property the Philox keying gives the real kernels."""

    def sample(self, x, model, y=None, cond_img=None, lsm_cond=None, topo_cond=None, seed=0, sample_offset=0):
        idx = torch.arange(sample_offset, sample_offset + x.shape[0], dtype=torch.float32)
        return x * 2 + idx[:, None, None, None] * 1000 + seed + (0 if lsm_cond is None else lsm_cond)


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n_total, 1, 8, 8, generator=g)
        lsm = torch.randn(n_total, 1, 8, 8, generator=g)
        full = sharding.sample_sharded(_FakeDiffusion(), None, x, lsm_cond=lsm, seed=5)
        ref = _FakeDiffusion().sample(x, None, lsm_cond=lsm, seed=5, sample_offset=0)
        q.put((rank, bool(torch.equal(full, ref)), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_sampling_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_total, world, port = 7, 2, 29731          # ragged: ranks get 4 and 3 samples
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and all(shape == (7, 1, 8, 8) for _, _, shape in res)
