"""GPU parity: ccv.grayscale + ccv.detect_objects through the C ABI vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest

import oracle
from headtrackr_b200 import synth

pytestmark = pytest.mark.gpu


def rect_tuple(d):
    return (d["x"], d["y"], d["width"], d["height"], d["confidence"], d.get("neighbors", d.get("neighbor")))


CASES = [("faces", 320, 240, 0), ("faces", 320, 240, 1), ("faces", 640, 480, 0), ("faces", 640, 480, 3),
         ("noise", 320, 240, 5), ("constant", 320, 240, 0), ("gradient", 320, 240, 0), ("faces", 333, 251, 2),
         ("faces", 1280, 720, 4)]


@pytest.mark.parametrize("kind,W,H,idx", CASES)
def test_detect_matches_oracle(ctx, blob, kind, W, H, idx):
    f = synth.frame(idx, W, H, kind=kind)
    want, want_raw = oracle.detect(f, blob, want_raw=True)
    got = ctx.detect(f, 5, 1)[0]
    got_raw, n_raw = ctx.debug_raw(0)
    assert n_raw == len(want_raw)
    assert got_raw == want_raw                       # raw list, reference (scale,q,y,x) order, bit-exact
    assert [rect_tuple(d) for d in got] == want      # grouped list, bit-exact, same order
    if kind == "faces":
        assert len(want) >= 1                        # parity must not be vacuous


def test_planes_match_oracle(ctx, blob):
    W, H = 320, 240
    f = synth.frame(7, W, H)
    ctx.detect(f, 5, 1)
    gray = oracle.grayscale(f)
    assert np.array_equal(ctx.debug_plane(0, 0, 0), gray)
    pyr = oracle.Pyramid(gray, 5)
    g = pyr.geom
    info = ctx.plan_info(W, H, 5)
    assert info["n_slots"] == g.n_slots and info["scale_upto"] == g.scale_upto
    assert info["w"] == list(g.w[: g.n_slots]) and info["h"] == list(g.h[: g.n_slots])
    for s in range(g.n_slots):
        for q in range(4 if s >= 2 * g.next else 1):
            assert np.array_equal(ctx.debug_plane(0, s, q), pyr.plane(s, q)), (s, q)


@pytest.mark.parametrize("interval,min_neighbors", [(5, 0), (3, 1), (5, 2), (2, 1)])
def test_detect_parameters(ctx, blob, interval, min_neighbors):
    f = synth.frame(11, 640, 480)
    want = oracle.detect(f, blob, interval=interval, min_neighbors=min_neighbors)
    got = ctx.detect(f, interval, min_neighbors)[0]
    assert [rect_tuple(d) for d in got] == want
    assert len(want) >= 1


def test_batch_equals_single(ctx, blob):
    frames = synth.batch(6, 320, 240, start=20)
    got = ctx.detect(frames, 5, 1)
    for i in range(6):
        want = oracle.detect(frames[i], blob)
        assert [rect_tuple(d) for d in got[i]] == want


def test_device_resident_input(ctx, blob):
    import torch
    frames = synth.batch(3, 640, 480, start=40)
    t = torch.from_numpy(frames).cuda()
    got = ctx.detect(t, 5, 1)
    for i in range(3):
        assert [rect_tuple(d) for d in got[i]] == oracle.detect(frames[i], blob)


def test_too_small_frame_is_rejected(ctx):
    from headtrackr_b200._lib import HtError, HT_ERR_SIZE
    f = synth.frame(0, 64, 48, kind="noise")
    with pytest.raises(HtError) as e:
        ctx.detect(f, 5, 1)
    assert e.value.code == HT_ERR_SIZE


def test_raw_list_overflow_is_reported(blob):
    """A raw list longer than max_raw_per_frame is truncated and reported as HT_WARN_OVERFLOW, never silently."""
    from headtrackr_b200 import Context
    from headtrackr_b200._lib import HT_WARN_OVERFLOW
    f = synth.frame(0, 320, 240)
    assert len(oracle.detect(f, blob, min_neighbors=0)) > 4
    c = Context(max_width=320, max_height=240, max_frames=2, max_raw_per_frame=4, max_rects_per_frame=2)
    try:
        rects, counts = c.detect_raw(f, 5, 0)
        assert c.last_warning and "overflow" in c.last_warning
        assert counts[0] == 2
    finally:
        c.close()
