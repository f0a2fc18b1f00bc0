"""GPU parity: box pyramid, Gaussian pyramid and Scharr derivatives, bit-exact against the oracle (and golden hashes)."""
import numpy as np
import pytest

import oracle
from _common import IMAGES, golden_json, load_gray, sha

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", IMAGES)
def test_lk_pyramid_bit_exact_fixture_images(ctx, name):
    im = load_gray(name)
    lv, dv = ctx.build_lk_pyramid(im, (21, 21), 3, True)
    g = golden_json("pyramid_hashes.json")["images"][name]
    assert [sha(a) for a in lv] == g["gauss"]
    assert [sha(a) for a in dv] == g["scharr"]
    lo, do = oracle.build_lk_pyramid(im, (21, 21), 3, True)
    assert all(np.array_equal(a, b) for a, b in zip(lv, lo)) and all(np.array_equal(a, b) for a, b in zip(dv, do))


@pytest.mark.parametrize("shape", [(47, 155), (64, 64), (33, 70), (135, 240), (1, 40), (40, 1), (2, 2), (129, 257), (376, 1241),
                                   (5, 300), (540, 960)])
def test_lk_pyramid_bit_exact_random_shapes(ctx, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    im = rng.integers(0, 256, shape, dtype=np.uint8)
    for win, ml in [((3, 3), 6), ((21, 21), 3)]:
        lv, dv = ctx.build_lk_pyramid(im, win, ml, True)
        lo, do = oracle.build_lk_pyramid(im, win, ml, True)
        assert len(lv) == len(lo)
        for l, (a, b) in enumerate(zip(lv, lo)):
            assert np.array_equal(a, b), (shape, win, "gauss", l)
        for l, (a, b) in enumerate(zip(dv, do)):
            assert np.array_equal(a, b), (shape, win, "scharr", l)


def test_lk_pyramid_strided_input(ctx):
    im = load_gray("kitti0.png")
    roi = im[7:300, 13:900]  # non-contiguous view: step 1240, width 887
    lv = ctx.build_lk_pyramid(roi, (21, 21), 3)
    lo = oracle.build_lk_pyramid(np.ascontiguousarray(roi), (21, 21), 3)
    assert all(np.array_equal(a, b) for a, b in zip(lv, lo))


@pytest.mark.parametrize("name", IMAGES[:3])
@pytest.mark.parametrize("mode", [0, 1])
def test_box_pyramid_bit_exact(ctx, dr3, name, mode):
    im = load_gray(name)
    got = ctx.box_pyramid(im, 3, mode)
    exp = oracle.box_pyramid(im, 3, mode)
    assert got[0] is not None and np.array_equal(got[0], im)
    for a, b in zip(got[1:], exp[1:]):
        assert np.array_equal(a, b)


def test_box_pyramid_known_answers_and_modes(ctx, dr3):
    box = ctx.box_pyramid(load_gray("kitti0.png"), 3)
    assert [sha(b)[:16] for b in box[1:]] == ["8befb7026fb4f64b", "6a236274936f5ccc"]
    odd = ctx.box_pyramid(load_gray("kitti_000000.png"), 3)
    assert [sha(b)[:16] for b in odd[1:]] == ["31c9d41ecba6551a", "0d90d380417b6584"]
    rng = np.random.default_rng(3)
    for shape in [(64, 96), (2160 // 4, 3840 // 4), (52, 36), (376, 1241), (10, 7)]:
        im = rng.integers(0, 256, shape, dtype=np.uint8)
        for mode in ([0, 1, 2] if shape[1] % 16 == 0 else [0, 1]):
            n_levels = 3 if min(shape) >= 8 else 2
            got, exp = ctx.box_pyramid(im, n_levels, mode), oracle.box_pyramid(im, n_levels, mode)
            for a, b in zip(got[1:], exp[1:]):
                assert np.array_equal(a, b), (shape, mode)
    # shapes on which the reference overruns its own buffers are rejected, not "fixed"
    with pytest.raises(dr3.Dr3lkError) as e:
        ctx.box_pyramid(np.zeros((375, 501), np.uint8), 2)
    assert e.value.code == dr3.E_UNSUPPORTED
    with pytest.raises(dr3.Dr3lkError):
        ctx.box_pyramid(np.zeros((64, 90), np.uint8), 2, dr3.BOX_SSE2)


def test_box_pyramid_device_batch(ctx, dr3):
    import torch
    rng = np.random.default_rng(5)
    B, h, w = 5, 96, 311
    ims = rng.integers(0, 256, (B, h, w), dtype=np.uint8)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        d = torch.from_numpy(ims).cuda()
        l1 = torch.zeros((B, h // 2, w // 2), dtype=torch.uint8, device="cuda")
        l2 = torch.zeros((B, h // 4, w // 4), dtype=torch.uint8, device="cuda")
        ctx.set_stream(stream.cuda_stream)
        ctx.box_pyramid_device(d.data_ptr(), w, h, w, h * w, B, [l1.data_ptr(), l2.data_ptr()], dr3.BOX_AUTO_X86)
        ctx.synchronize()
        ctx.set_stream(None)
    for b in range(B):
        exp = oracle.box_pyramid(ims[b], 3)
        assert np.array_equal(l1[b].cpu().numpy(), exp[1]) and np.array_equal(l2[b].cpu().numpy(), exp[2])
