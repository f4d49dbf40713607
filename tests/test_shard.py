"""Multi-GPU host logic on CPU: GOP-segment splitting and the rank assignment (no collective on the data path).  The
world_size-2 gloo test checks that two ranks derive disjoint, complete shards independently and that the max-over-ranks timing
reduction bench.py uses works."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "streams")


def test_split_segments_on_golden_inter_stream(built):
    import av1recon
    from av1recon import shard
    from tools.obuio import read_ivf
    tus = read_ivf(os.path.join(GOLD, "inter_8b_sb128_tiles_640x360.ivf"))   # kf_max_dist 6, 10 frames
    segs = shard.split_segments(tus, av1recon.scan_headers(tus))
    assert segs[0][0] == 0 and segs[-1][1] == len(tus) and len(segs) >= 2
    for (a, b), (c, d) in zip(segs, segs[1:]):
        assert b == c and a < b
    # every segment is decodable on its own: the oracle decodes each one from a fresh state to the same frames
    from oracle import oracle_lib
    import numpy as np
    whole, _ = oracle_lib.decode_stream(tus)
    k = 0
    for a, b in segs:
        part, _ = oracle_lib.decode_stream(tus[a:b])
        for fr in part:
            for p in range(3):
                assert np.array_equal(fr[p], whole[k][p])
            k += 1
    assert k == len(whole)


def test_assign_is_balanced_and_complete():
    from av1recon import shard
    items = [((f, s), 1000 + 37 * ((f * 7 + s * 3) % 11)) for f in range(32) for s in range(4)]
    for world in (1, 2, 4, 8):
        parts = shard.assign(items, world)
        flat = sorted(k for p in parts for k in p)
        assert flat == sorted(k for k, _ in items)
        w = dict(items)
        loads = [sum(w[k] for k in p) for p in parts]
        assert max(loads) - min(loads) <= max(w.values())


def _rank_main(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "av1-go_b200"))
    from av1recon import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = [((f, s), 100 + f + s) for f in range(8) for s in range(4)]
    mine = shard.assign(items, world)[rank]
    # ranks share nothing on the data path; only the timing is reduced (max over ranks), as in bench.py
    t = torch.tensor([float(10 + rank)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([len(mine)])
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    q.put((rank, mine, float(t.item()), int(n.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_disjoint_shards():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    a, b = set(res[0][1]), set(res[1][1])
    assert not (a & b) and len(a | b) == 32
    assert res[0][2] == res[1][2] == 11.0 and res[0][3] == 32
