"""world_size-2 gloo test of the multi-GPU host logic: contiguous shards of
independent problems, each rank materialises its shard from (seed, index), no data-path collective,
one all_gather of results at the end.  The per-shard solve is done by the CPU oracle here (a test
may use it); on GPUs bench.py runs the same partition with the CUDA path and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tinyopt_b200.shard import gather_results, gather_rows, shard_range


def test_shard_range_partitions():
    for B in (0, 1, 5, 64, 100, 100_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-B // world) or B == 0


def _worker(rank, world, port, B, m, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    lo, hi = shard_range(B, rank, world)
    A, y, xs, x0 = O.synth_generate(hi - lo, m, n, np.float64, p0=lo)
    x, res, _ = O.synth_lm_run(A, y, x0, O.default_options(), nthreads=1)
    X = gather_rows(torch.from_numpy(x), B, rank, world)
    R = gather_results(res, B, rank, world, "cpu")
    if rank == 0:
        q.put((X.numpy(), R["num_iters"].copy(), R["stop_reason"].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_process():
    from oracle import oracle as O
    B, m, n = 101, 12, 3  # ragged: 51 + 50
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, m, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    X, iters, stops = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    A, y, xs, x0 = O.synth_generate(B, m, n, np.float64)
    x, res, _ = O.synth_lm_run(A, y, x0, O.default_options(), nthreads=1)
    assert np.array_equal(X, x)
    assert np.array_equal(iters, res["num_iters"]) and np.array_equal(stops, res["stop_reason"])
