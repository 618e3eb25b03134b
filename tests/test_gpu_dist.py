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


def _worker(rank, world, port, ref_len, out_dir, native):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    if native:                                   # the process group only carries the communicator id
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from basecount_b200 import dist as bdist
        from basecount_b200.engine import Engine
        batch = _reads(7, ref_len)
        with Engine(rank) as eng:
            if native:
                bdist.engine_comm(eng, dist, rank, world)
            be = bdist.GpuBackend(eng, torch.device("cuda", rank), native_comm=native)
            bounds = bdist.count_region_sharded(be, dist, rank, world, batch, ref_len, min_base_quality=20)
            pc, depth, ent = bdist.summary_region_sharded(be, dist, world, ref_len)
            np.savez(os.path.join(out_dir, f"r{rank}.npz"), counts=eng.counts(0), lo=bounds[rank], hi=bounds[rank + 1],
                     summary=np.array([pc, depth, ent]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("native", [True, False])      # the library's own communicator / the host moving the halo
def test_region_sharding_over_nccl_matches_single_pass(tmp_path, native):
    from oracle import bcount as obc
    from oracle import stats as ost
    world, ref_len = min(torch.cuda.device_count(), 4), 6001
    mp.spawn(_worker, args=(world, _free_port(), ref_len, str(tmp_path), native), nprocs=world, join=True)
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


def test_library_communicator_on_one_gpu():
    """World size 1 on whatever GPU there is: the communicator, the (empty) halo merge with its asynchronous
    truncation, the all-reduced summary and bc_set_length restoring the slot -- the same calls a rank of a
    region-sharded run makes, checked against the oracle."""
    from basecount_b200 import _lib, synth
    from basecount_b200.engine import Engine
    from basecount_b200.pack import pack_batches
    from basecount_b200.records import select_reads
    from oracle import bcount as obc
    from oracle import stats as ost
    ref_len, halo = 3000, 300
    b = select_reads(synth.amplicon_sample(seed=12, n_reads=4000, ref_len=ref_len, ref_name="x"), 0, 0)
    want = obc.bcount_flat(ref_len, 0, b).astype(np.int64)
    own = ref_len - halo
    with Engine(0) as eng:
        eng.comm_init(1, 0, Engine.comm_unique_id())
        assert eng.allgather_u32(halo) == [halo]
        out = (_lib.pinned_empty(1, np.int64), _lib.pinned_empty(1, np.int64), _lib.pinned_empty(1, np.float64))
        eng.begin([ref_len])
        for step in range(3):                       # a stream of samples: restore the halo columns, count, merge, summarise
            eng.set_length(0, ref_len)
            eng.reset()
            eng.push(pack_batches(b, 0))
            eng.halo_merge(0, [0, own], [halo])
            eng.summary_allreduce_async(out, False)
        eng.sync()
        got = eng.counts(0)
        assert got.shape[0] == own and np.array_equal(got, want[:own])
        cov, ent, _ = ost.per_position_vectors(want[:own].tolist())
        assert int(out[0][0]) == sum(1 for x in cov if x) and int(out[1][0]) == sum(cov)
        assert float(out[2][0]) == pytest.approx(float(np.sum(ent)), rel=1e-12)
        eng.comm_destroy()
