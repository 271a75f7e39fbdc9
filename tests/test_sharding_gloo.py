"""N>1 host logic on CPU: world_size-2 gloo processes shard the sequences, generate only their own
streams, and the gathered read-outs equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ingvio_b200.sharding import gather_readouts, gather_readouts_device, max_over_ranks, shard_range
from ingvio_b200.synth import WORKLOADS, SyntheticStream

TOTAL = 5


def _readout(lo, hi):
    st = SyntheticStream(WORKLOADS["tiny"], hi - lo, seq0=lo)
    ini = st.initial_state()
    fr = st.next_frame()
    return np.concatenate([ini["p"], ini["v"], fr.gyro[:, 0, :], fr.pf_w[:, 0, :]], 1)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(TOTAL, world, rank)
    local = _readout(lo, hi)
    full = gather_readouts(local, TOTAL, dist)
    full_dev = gather_readouts_device(torch.from_numpy(local), TOTAL, dist).numpy()   # the tensor path bench.py's c4 leg uses
    assert np.array_equal(full, full_dev)
    ms = max_over_ranks(10.0 + rank, dist)
    q.put((rank, lo, hi, full, ms))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    for total in (1, 5, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_matches_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _readout(0, TOTAL)
    for rank, lo, hi, full, ms in res:
        assert (lo, hi) == shard_range(TOTAL, 2, rank)
        assert np.array_equal(full, ref)      # sequence b is the same stream wherever it is generated
        assert ms == 11.0
