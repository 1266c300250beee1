"""World-size-2 gloo test (CPU) of the multi-rank host logic: shard bounds, index-meta broadcast, per-rank
search of the rank's shard (the C oracle stands in for the GPU here) and independence from the shard count."""
import importlib
import os
import socket
import sys

import numpy as np

import helpers


def test_shard_bounds_match_c_driver():
    sharding = importlib.import_module(helpers.PKG_NAME + ".sharding")
    for nq in (0, 1, 31, 32, 33, 1000, 2001, 10_000_000, 100_000_000):
        for g in (1, 2, 3, 4, 8):
            b = sharding.shard_bounds(nq, g)
            assert b[0] == 0 and b[-1] == nq and all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % 32 == 0 for x in b[:-1] if x < nq)
            assert sum(y - x for x, y in zip(b, b[1:])) == nq
            assert max(y - x for x, y in zip(b, b[1:])) <= ((nq + g - 1) // g + 31)


def _worker(rank, world, port, golden, out_dir):
    sys.path.insert(0, helpers.ROOT); sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    import torch
    import torch.distributed as dist
    pkg = helpers.pkg()
    sharding = importlib.import_module(helpers.PKG_NAME + ".sharding")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(golden)
    length, reads = int(g["length"]), g["reads"]
    nq = reads.size // length
    # rank 0 "owns" the index: broadcast its shape like bench.py does for the device table
    meta = pkg.fmgpu_index_meta_t(2, int(g["n"]) + 1, 16, 216, 100, 0xFFFFFFFF, 0, 0, 16 * 16 * 216) if rank == 0 else None
    meta = sharding.broadcast_meta(meta, 0, dist, pkg.fmgpu_index_meta_t)
    assert (meta.steps, meta.nsymbols, meta.nbytes) == (2, 16, 16 * 16 * 216)
    lo, hi = sharding.shard_range(nq, world, rank)
    o = helpers.Oracle()
    h = o.wrap(g["image_100"])
    local = o.search(h, reads[lo * length: hi * length], length)
    full = sharding.gather_intervals(torch.from_numpy(local.astype(np.int64)), nq, dist, torch).numpy().astype(np.uint32)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                    # the bench's max-over-ranks timing reduction
    assert t.item() == world
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_search_gloo(built, tmp_path):
    import torch.multiprocessing as mp
    golden = os.path.join(helpers.ROOT, "tests", "golden", "small_k2_d64.npz")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, golden, str(tmp_path)), nprocs=2, join=True)
    want = np.load(golden)["expected_std"]
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"rank{r}.npy"), want)
