"""GPU: frame quads, L2 waves and the exactness fallbacks of the round-2 detector.

The pyramid arena is frame-quad-interleaved (4 frames per word) and batches are processed in waves; these tests pin
the cases that layout adds: batches that are not a multiple of 4, several waves, the piped variant, odd frame sizes
in multi-frame batches, and - with ht_debug_set_exactness - the fallback branches that real data never takes.
"""
import math

import numpy as np
import pytest

import oracle
from headtrackr_b200 import Context, synth

pytestmark = pytest.mark.gpu


def tup(d):
    return (d["x"], d["y"], d["width"], d["height"], d["confidence"], d.get("neighbors", d.get("neighbor")))


def test_partial_quads_and_planes(ctx, blob):
    frames = synth.batch(7, 320, 240, start=60)            # 1 full quad + 3 frames
    got = ctx.detect(frames, 5, 1)
    for i in range(7):
        assert [tup(d) for d in got[i]] == oracle.detect(frames[i], blob), i
    for i in (2, 5, 6):                                      # byte lanes 2, 1, 2 of two different quads
        pyr = oracle.Pyramid(oracle.grayscale(frames[i]), 5)
        g = pyr.geom
        for s in (0, 3, 6, 13, g.n_slots - 1):
            for q in range(4 if s >= 2 * g.next else 1):
                assert np.array_equal(ctx.debug_plane(i, s, q), pyr.plane(s, q)), (i, s, q)


@pytest.mark.parametrize("env", [{"HT_WAVE": "8"}, {"HT_WAVE": "8", "HT_DETECT_PIPE": "1"}, {"HT_WAVE": "4"}])
def test_waves_do_not_change_results(blob, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    frames = synth.batch(21, 320, 240, start=100)
    c = Context(max_width=320, max_height=240, max_frames=32)
    try:
        for _ in range(2):                                   # twice: the arenas are re-used across calls
            got = c.detect(frames, 5, 1)
            for i in range(21):
                assert [tup(d) for d in got[i]] == oracle.detect(frames[i], blob), i
        dets, found, objs, wins = c.detect_track(frames, 5, 1, calc_angles=False, n_calls=2)
        for i in (0, 9, 20):
            _, fnd, obj = oracle.detect_track(frames[i], blob, 5, 1, False, 2)
            assert [tup(d) for d in dets[i]] == oracle.detect(frames[i], blob)
            assert found[i] == fnd
            if fnd:
                assert (objs[i]["x"], objs[i]["y"], objs[i]["width"], objs[i]["height"]) == (obj["x"], obj["y"], obj["width"], obj["height"])
    finally:
        c.close()


def test_table_driven_cascade_matches(blob, monkeypatch):
    monkeypatch.setenv("HT_NO_FAST", "1")                   # the path any OTHER cascade blob takes
    c = Context(max_width=640, max_height=480, max_frames=8)
    try:
        frames = synth.batch(5, 320, 240, start=7)
        got = c.detect(frames, 5, 1)
        for i in range(5):
            assert [tup(d) for d in got[i]] == oracle.detect(frames[i], blob)
        f = synth.frame(3, 640, 480)
        assert [tup(d) for d in c.detect(f, 5, 0)[0]] == oracle.detect(f, blob, min_neighbors=0)
    finally:
        c.close()


@pytest.mark.parametrize("flags", [1, 2, 3])
def test_forced_cascade_ties_do_not_change_results(blob, flags):
    """bit 0 / 1: every generated / late stage decision is re-decided by the reference's ordered fp64 adds."""
    c = Context(max_width=640, max_height=480, max_frames=8, max_raw_per_frame=4096)
    try:
        c.debug_set_exactness(flags)
        for W, H, idx in ((320, 240, 0), (640, 480, 3), (333, 251, 2)):
            f = synth.frame(idx, W, H)
            want, want_raw = oracle.detect(f, blob, want_raw=True)
            assert [tup(d) for d in c.detect(f, 5, 1)[0]] == want
            got_raw, n_raw = c.debug_raw(0)
            assert got_raw == want_raw and n_raw >= 1
    finally:
        c.close()


def test_forced_serial_moments_do_not_change_results(blob):
    """bit 2: every mean-shift pass takes moments_serial (the reference's x-outer / y-inner order)."""
    c = Context(max_width=640, max_height=480, max_frames=8)
    try:
        frames = synth.batch(4, 320, 240, start=30)
        c.debug_set_exactness(4)
        c.debug_track_stats(reset=True)
        dets, found, objs, wins = c.detect_track(frames, 5, 1, calc_angles=True, n_calls=5)
        st = c.debug_track_stats(reset=True)
        assert st["serial_passes"] > 0 and st["serial_passes"] >= st["passes"] - st["memo_hits"] - 4 * 5
        for i in range(4):
            cand = None
            for r in oracle.detect(frames[i], blob):
                if cand is None or r[4] > cand[4]:
                    cand = r
            assert found[i] == (1 if cand else 0)
            if not cand:
                continue
            ot = oracle.CamshiftTracker(calc_angles=True)
            ot.init_tracker(frames[i], *[int(math.floor(v)) for v in cand[:4]])
            for _ in range(5):
                ot.track(frames[i])
            w = ot.track_obj()
            assert (objs[i]["x"], objs[i]["y"], objs[i]["width"], objs[i]["height"]) == (w["x"], w["y"], w["width"], w["height"])
            assert abs(objs[i]["angle"] - w["angle"]) <= 1e-4          # north_star tolerance
            assert wins[i] == ot.search_window()
    finally:
        c.close()


def test_odd_sized_multi_frame_track(ctx, blob):
    """with an odd w*h every odd frame of a batch starts at 4 mod 8 bytes (bin plane at 2 mod 4)."""
    W, H = 333, 251
    frames = synth.batch(3, W, H, start=2)
    rects = []
    for i in range(3):
        want = oracle.detect(frames[i], blob)
        assert want
        rects.append([int(math.floor(v)) for v in want[0][:4]])
    ctx.track_init(frames, rects, calc_angles=False)
    objs, wins = ctx.track(frames, n_calls=3)
    dets, found, objs2, wins2 = ctx.detect_track(frames, 5, 1, calc_angles=False, n_calls=3)
    for i in range(3):
        ot = oracle.CamshiftTracker(calc_angles=False)
        ot.init_tracker(frames[i], *rects[i])
        for _ in range(3):
            ot.track(frames[i])
        w = ot.track_obj()
        assert (objs[i]["x"], objs[i]["y"], objs[i]["width"], objs[i]["height"]) == (w["x"], w["y"], w["width"], w["height"])
        assert wins[i] == ot.search_window()
        assert [tup(d) for d in dets[i]] == oracle.detect(frames[i], blob)


def test_track_on_a_4_byte_aligned_device_pointer(ctx, blob):
    """The public API only asks for 4-byte alignment of `rgba` (a torch slice at 4 mod 8 is legal)."""
    import torch
    W, H = 320, 240
    f = synth.frame(9, W, H)
    want = oracle.detect(f, blob)
    rect = [int(math.floor(v)) for v in want[0][:4]]
    buf = torch.zeros(W * H * 4 + 16, dtype=torch.uint8, device="cuda")
    view = buf[4:4 + W * H * 4]
    view.copy_(torch.from_numpy(f.reshape(-1)).cuda())
    t = view.view(1, H, W, 4)
    assert t.data_ptr() % 8 == 4
    ctx.track_init(t, [rect], calc_angles=False)
    objs, wins = ctx.track(t, n_calls=2)
    ot = oracle.CamshiftTracker(calc_angles=False)
    ot.init_tracker(f, *rect)
    for _ in range(2):
        ot.track(f)
    w = ot.track_obj()
    assert (objs[0]["x"], objs[0]["y"], objs[0]["width"], objs[0]["height"]) == (w["x"], w["y"], w["width"], w["height"])
    assert [tup(d) for d in ctx.detect(t, 5, 1)[0]] == want


@pytest.mark.parametrize("calc_angles", [False, True])
def test_crafted_frame_takes_the_serial_fallback_by_itself(ctx, calc_angles):
    """A frame on which mean-shift steps are EXACT integers in real arithmetic: a 40x60 block of one colour on a
    background of another, tracker seeded inside the block (so only the block's bin has weight, v = 1000/2400, not
    dyadic).  Once the window contains the block, (xc - w/2) is an integer k in exact arithmetic and k +- 1e-14 in
    fp64 depending on the summation order - `>>0` (src/camshift.js:295-296) would truncate to k or k-1.  The kernel
    must notice (trunc_ambiguous) and re-derive the moments in the reference's order: serial_passes > 0 WITHOUT the
    debug knob, and every call equal to the oracle."""
    W, H = 160, 120
    f = np.zeros((H, W, 4), np.uint8)
    f[..., 3] = 255
    f[..., :3] = (200, 50, 50)
    f[30:90, 60:100, :3] = (50, 200, 50)
    rect = [70, 40, 25, 40]
    ctx.set_track_memo(False)
    try:
        ctx.debug_set_exactness(0)
        ctx.track_init(f, [rect], calc_angles=calc_angles)
        ot = oracle.CamshiftTracker(calc_angles=calc_angles)
        ot.init_tracker(f, *rect)
        ctx.debug_track_stats(reset=True)
        for call in range(8):
            objs, wins = ctx.track(f, n_calls=1)
            ot.track(f)
            w = ot.track_obj()
            assert (objs[0]["x"], objs[0]["y"], objs[0]["width"], objs[0]["height"]) == (w["x"], w["y"], w["width"], w["height"]), call
            assert abs(objs[0]["angle"] - w["angle"]) <= 1e-4
            assert wins[0] == ot.search_window(), call
        st = ctx.debug_track_stats(reset=True)
        assert st["serial_passes"] > 0, st
    finally:
        ctx.set_track_memo(True)
