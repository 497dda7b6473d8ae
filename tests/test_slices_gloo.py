"""(e) multi-GPU host logic at world_size 2 over gloo on CPU: every rank seeds its contiguous time
slice analytically (doppler_b200.slicing, no data-path collective), mixes it, and the gathered
bytes equal the single-process stream.  The CUDA path cannot run here, so the per-slice mixing is
done by the oracle as a stand-in; what is under test is the slicing, the seed and the chaining."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from doppler_b200 import slicing
        from tests.oracle_lib import BPS, Oracle
        oracle = Oracle()
        intype, outtype, fs, total = case["intype"], case["outtype"], case["fs"], case["total"]
        rng = np.random.default_rng(77)  # same stream on every rank
        if intype == 0:
            stream = rng.integers(-32768, 32768, 2 * total, dtype=np.int32).astype("<i2").view(np.uint8)
        else:
            stream = rng.uniform(-1, 1, 2 * total).astype("<f4").view(np.uint8)
        begin, end = slicing.slice_bounds(total, world, rank, intype)
        mine = stream[begin * BPS[intype]:end * BPS[intype]]
        if case["mode"] == "const":
            seed = slicing.seed_const(case["shift"], fs, begin)
            out, sn = oracle.mix(mine, intype, outtype, case["shift"], fs, samplenum=seed)
        else:
            shifts = np.array(case["shifts"], dtype=np.float32)
            seed = slicing.seed_blocks(shifts, intype, fs, begin)
            b0 = begin // slicing.block_samples(intype)
            out, sn = oracle.mix_blocks(mine, intype, outtype, shifts[b0:], fs, samplenum=seed)
        # gather sizes, then payloads (gloo); also the max-over-ranks reduction bench.py uses
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([out.size], dtype=torch.int64))
        cap = int(max(s.item() for s in sizes))
        pad = torch.zeros(cap, dtype=torch.uint8)
        pad[:out.size] = torch.from_numpy(out.copy())
        parts = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, pad)
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        last_sn = torch.tensor([sn if rank == world - 1 else 0], dtype=torch.int64)
        dist.all_reduce(last_sn, op=dist.ReduceOp.MAX)
        if rank == 0:
            got = np.concatenate([p.numpy()[:int(s.item())] for p, s in zip(parts, sizes)])
            if case["mode"] == "const":
                want, sn_want = oracle.mix(stream, intype, outtype, case["shift"], fs)
            else:
                want, sn_want = oracle.mix_blocks(stream, intype, outtype, np.array(case["shifts"], dtype=np.float32), fs)
            ok = bool(np.array_equal(got, want)) and int(last_sn.item()) == sn_want and t.item() == float(world)
            q.put((ok, begin, end))
    finally:
        dist.destroy_process_group()


CASES = [
    {"mode": "const", "intype": 0, "outtype": 0, "fs": 256000, "shift": -15000.0, "total": 7 * 2048 + 5},
    {"mode": "const", "intype": 1, "outtype": 0, "fs": 1_024_000, "shift": 7321.7, "total": 9 * 1024 + 333},
    {"mode": "blocks", "intype": 0, "outtype": 1, "fs": 1_024_000, "total": 6 * 2048 - 100,
     "shifts": [-9876.54, -9876.54, -9871.02, 5000.0, 0.0, 7321.7]},
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['mode']}-{c['intype']}{c['outtype']}")
def test_two_rank_time_slices_reproduce_the_stream(case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, begin, end = q.get(timeout=10)
    assert ok


def test_slice_bounds_cover_the_stream_on_block_boundaries():
    from doppler_b200 import slicing
    for total in (0, 1, 2047, 2048, 10 * 2048 + 17, 8 * 2048):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                b, e = slicing.slice_bounds(total, world, r, 0)
                assert b == prev and e >= b
                if r < world - 1:
                    assert e % 2048 == 0
                prev = e
            assert prev == total
