"""CPU test of the N>1 path: two gloo ranks shard the frame pairs, track their block and gather on rank 0.
The compute stand-in is the oracle (no GPU here); on the GPU box bench.py --gpus N exercises the same sharding."""
import importlib
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from _common import ROOT, load_gray, random_points


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle
    sh = importlib.import_module("3dr_b200.sharding")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = np.load(os.path.join(tmp, "in.npz"))

    def track(prev, nxt, pts, offs):
        outs = [oracle.calc_optical_flow_pyr_lk(prev[b], nxt[b], pts[offs[b]:offs[b + 1]], nthreads=1) for b in range(len(offs) - 1)]
        return (np.concatenate([o[0] for o in outs]), np.concatenate([o[1] for o in outs]), np.concatenate([o[2] for o in outs]))

    res = sh.track_sharded(track, d["prev"], d["nxt"], d["pts"], d["offs"], rank, world, dist)
    if rank == 0:
        np.savez(os.path.join(tmp, "out.npz"), p=res[0], s=res[1], e=res[2])
    else:
        assert res is None
    dist.destroy_process_group()


def test_shard_ranges_cover_without_overlap():
    sh = importlib.import_module("3dr_b200.sharding")
    for n in (1, 5, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [sh.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1


def test_two_rank_gloo_sharded_tracking_matches_single_process(tmp_path):
    import oracle
    frames = [load_gray("kitti%d.png" % i)[100:260, 200:520].copy() for i in range(6)]
    prev, nxt = np.stack(frames[:5]), np.stack(frames[1:6])
    rng = np.random.default_rng(8)
    pts_list = [random_points(rng, 320, 160, n, margin=10) for n in (120, 0, 77, 200, 31)]  # ragged, one empty pair
    offs = np.concatenate([[0], np.cumsum([len(p) for p in pts_list])]).astype(np.int32)
    pts = np.concatenate(pts_list).astype(np.float32)
    np.savez(str(tmp_path / "in.npz"), prev=prev, nxt=nxt, pts=pts, offs=offs)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out = np.load(str(tmp_path / "out.npz"))
    exp = [oracle.calc_optical_flow_pyr_lk(prev[b], nxt[b], pts_list[b]) for b in range(5)]
    assert np.array_equal(out["p"].view(np.uint32), np.concatenate([e[0] for e in exp]).view(np.uint32))
    assert np.array_equal(out["s"], np.concatenate([e[1] for e in exp]))
    assert np.array_equal(out["e"], np.concatenate([e[2] for e in exp]))
