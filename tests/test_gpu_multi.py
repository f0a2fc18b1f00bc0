"""dr3lk_multi: the host-buffer batch sharded by frame pair over several contexts / GPUs from ONE process.
Results must be bit-identical to the single-device call whatever the number of ranks -- also with ragged point counts, empty
pairs and more ranks than pairs.  Listing device 0 twice exercises the sharding on a single-GPU box; with >= 2 devices the
same batch is also split over two GPUs."""
import numpy as np
import pytest

import oracle
from _common import load_gray, random_points

pytestmark = pytest.mark.gpu


def _batch(rng, n_pairs, ragged=True):
    frames = [load_gray("kitti%d.png" % i) for i in range(6)]
    prev = np.stack([frames[i % 5] for i in range(n_pairs)])
    nxt = np.stack([frames[i % 5 + 1] for i in range(n_pairs)])
    counts = rng.integers(0, 400, n_pairs) if ragged else np.full(n_pairs, 256)
    if ragged and n_pairs > 2:
        counts[1] = 0  # a pair without points
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    pts = random_points(rng, 1240, 376, int(offs[-1]))
    return prev, nxt, pts, offs


def _equal(a, b):
    return all((x is None and y is None) or np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))


def test_shard_range_matches_python_rule(dr3):
    from importlib import import_module
    sharding = import_module("3dr_b200.sharding")
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            blocks = [dr3.shard_range(n, r, world) for r in range(world)]
            assert blocks == [sharding.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0], [0] * 8])
def test_multi_on_one_device_matches_single_context(ctx, dr3, devices):
    rng = np.random.default_rng(len(devices))
    for n_pairs, ragged in ((13, True), (8, False), (3, True)):
        prev, nxt, pts, offs = _batch(rng, n_pairs, ragged)
        exp = ctx.track_batch_host(prev, nxt, pts, offs, want_stats=True)
        with dr3.MultiContext(devices) as m:
            got = m.track_batch_host(prev, nxt, pts, offs, want_stats=True)
            assert m.launch_count > 0
        assert _equal(got, exp), (devices, n_pairs)
    # ... and the sharded result is the oracle's, pair by pair
    for b in range(3):
        sl = slice(offs[b], offs[b + 1])
        po, so, eo = oracle.calc_optical_flow_pyr_lk(prev[b], nxt[b], pts[sl])
        assert np.array_equal(got[1][sl], so) and np.array_equal(got[0][sl].view(np.uint32), po.view(np.uint32))


def test_multi_initial_flow_and_errors(ctx, dr3):
    rng = np.random.default_rng(5)
    prev, nxt, pts, offs = _batch(rng, 6)
    init = (pts + rng.normal(0, 1.0, pts.shape)).astype(np.float32)
    exp = ctx.track_batch_host(prev, nxt, pts, offs, init, flags=dr3.USE_INITIAL_FLOW)
    with dr3.MultiContext([0, 0]) as m:
        got = m.track_batch_host(prev, nxt, pts, offs, init, flags=dr3.USE_INITIAL_FLOW)
        assert _equal(got, exp)
        with pytest.raises(dr3.Dr3lkError) as e:
            m.track_batch_host(prev, nxt, pts, offs, win=(2, 2))
        assert e.value.code == dr3.E_ARG and "rank" in str(e.value)
        bad = offs.copy(); bad[2] = bad[1] - 1
        with pytest.raises(dr3.Dr3lkError):
            m.track_batch_host(prev, nxt, pts, bad)
    with pytest.raises(dr3.Dr3lkError):
        dr3.MultiContext([9999])


def test_multi_two_gpus_matches_one(ctx, dr3):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    rng = np.random.default_rng(9)
    prev, nxt, pts, offs = _batch(rng, 24)
    exp = ctx.track_batch_host(prev, nxt, pts, offs, want_stats=True)
    n_dev = min(torch.cuda.device_count(), 8)
    with dr3.MultiContext(list(range(n_dev))) as m:
        got = m.track_batch_host(prev, nxt, pts, offs, want_stats=True)
    assert _equal(got, exp)
