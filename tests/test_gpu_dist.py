"""Region sharding with the real GPU backend over NCCL (world_size 2; skipped on a one-GPU box).

Same logic as tests/test_dist_cpu.py, but every rank counts its region with the CUDA path and the
halo columns travel GPU to GPU with NCCL send/recv; the merged matrix must equal the oracle's
single pass and the all-reduced summary the oracle's summary."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_dist_cpu import _free_port, _reads

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ref_len, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from basecount_b200 import dist as bdist
        from basecount_b200.engine import Engine
        batch = _reads(7, ref_len)
        with Engine(rank) as eng:
            be = bdist.GpuBackend(eng, torch.device("cuda", rank))
            bounds = bdist.count_region_sharded(be, dist, rank, world, batch, ref_len, min_base_quality=20)
            pc, depth, ent = bdist.summary_region_sharded(be, dist, world, ref_len)
            np.savez(os.path.join(out_dir, f"r{rank}.npz"), counts=eng.counts(0), lo=bounds[rank], hi=bounds[rank + 1],
                     summary=np.array([pc, depth, ent]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_region_sharding_over_nccl_matches_single_pass(tmp_path):
    from oracle import bcount as obc
    from oracle import stats as ost
    world, ref_len = 2, 6001
    mp.spawn(_worker, args=(world, _free_port(), ref_len, str(tmp_path)), nprocs=world, join=True)
    want = obc.bcount_flat(ref_len, 20, _reads(7, ref_len)).astype(np.int64)
    got = np.zeros_like(want)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert z["counts"].shape[0] == z["hi"] - z["lo"]
        got[int(z["lo"]):int(z["hi"])] = z["counts"]
    assert np.array_equal(got, want)
    cov, ent, _ = ost.per_position_vectors(want.tolist())
    pc, depth, avg_ent = ost.summary(cov, ent, ref_len)
    for r in range(world):
        s = np.load(tmp_path / f"r{r}.npz")["summary"]
        assert s[0] == pc and s[1] == depth
        assert s[2] == pytest.approx(float(avg_ent), rel=1e-12)
