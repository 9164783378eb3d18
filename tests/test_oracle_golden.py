"""CPU: the C oracle against the golden vectors produced by executing the reference's own JavaScript
(tools/make_goldens.py: /root/reference/src/*.js run by oracle/jsmini.py over the canvas shim)."""
import hashlib
import json
import math
from pathlib import Path

import numpy as np
import pytest

import oracle
from headtrackr_b200 import synth

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_js.json").read_text())


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_cascade_blob_is_the_one_the_goldens_used(blob):
    assert hashlib.sha256(blob).hexdigest() == GOLD["cascade_blob_sha256"]


@pytest.mark.parametrize("case", GOLD["detect"], ids=lambda c: c["name"])
def test_detect_golden(case, blob):
    f = synth.frame(case["index"], case["W"], case["H"], n_faces=case["n_faces"], kind=case["kind"])
    assert sha(f) == case["frame_sha256"], "synthetic frame generator drifted"
    assert sha(oracle.grayscale(f)) == case["gray_sha256"]
    got = oracle.detect(f, blob, interval=case["interval"], min_neighbors=case["min_neighbors"])
    assert got == [tuple(r) for r in case["rects"]]            # bit-exact doubles, same order


def track_frames(case):
    f = synth.frame(case["index"], case["W"], case["H"], n_faces=case["n_faces"])
    if case["track_frame"] == "same":
        t = f
    elif case["track_frame"] == "constant":
        t = synth.frame(0, case["W"], case["H"], kind="constant")
    else:
        t = np.roll(f, 3, axis=1)
    return f, t


@pytest.mark.parametrize("case", GOLD["track"], ids=lambda c: c["name"])
def test_track_golden(case):
    f, t = track_frames(case)
    assert sha(f) == case["frame_sha256"]
    ot = oracle.CamshiftTracker(calc_angles=case["calc_angles"])
    ot.init_tracker(f, *case["rect"])
    for call in case["calls"]:
        ot.track(t)
        o = ot.track_obj()
        assert [o["x"], o["y"], o["width"], o["height"]] == call["obj"][:4]
        assert abs(o["angle"] - call["obj"][4]) <= 1e-12
        assert list(ot.search_window()) == call["window"]


def test_lost_face_is_zero_sized():
    case = [c for c in GOLD["track"] if c["name"] == "track_lost"][0]
    assert case["calls"][-1]["obj"][2:4] == [0, 0]               # src/main.js:230 lost-face condition


def test_whitebalance_golden():
    for c in GOLD["whitebalance"]:
        f = synth.frame(c["index"], c["W"], c["H"], kind=c["kind"])
        assert sha(f) == c["frame_sha256"]
        assert oracle.whitebalance(f) == c["value"]


def test_shim_two_implementations_agree():
    """The defined resampler exists twice (C in the oracle, numpy in synth/jsmini): they must agree."""
    rng = np.random.default_rng(7)
    src = rng.integers(0, 256, (61, 83), dtype=np.uint8)
    for (dw, dh, sx, sy, sw, sh) in [(41, 30, 0, 0, 83, 61), (39, 30, 1, 0, 82, 61), (41, 28, 0, 1, 83, 60),
                                     (39, 28, 1, 1, 82, 60), (83, 61, 0, 0, 83, 61), (74, 54, 0, 0, 83, 61),
                                     (1, 1, 0, 0, 83, 61), (120, 90, 0, 0, 83, 61)]:
        a = synth.shim_resize(src, dw, dh, sx, sy, sw, sh)
        b = oracle.draw_image(src, sx, sy, sw, sh, dw, dh, dw, dh)
        assert np.array_equal(a, b), (dw, dh, sx, sy)
    even = rng.integers(0, 256, (40, 64), dtype=np.uint8)
    half = oracle.draw_image(even, 0, 0, 64, 40, 32, 20, 32, 20)
    s = even.astype(np.int32)
    mean = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) // 4
    assert np.array_equal(half, mean.astype(np.uint8))            # 2:1 of even sizes == rounded 2x2 mean


def test_geometry_matches_survey():
    g = oracle.geometry(640, 480, 5)
    assert (g.scale_upto, g.n_slots) == (27, 39)
    assert [g.w[i] for i in range(6)] == [640, 570, 507, 452, 403, 359]
    st = oracle.detect(synth.frame(0, 640, 480, kind="constant"), synth.load_cascade_blob(), want_stats=True)[1]
    assert st.windows == 312640 and st.feature_evals == 4 * 312640    # 4 evals/window on flat input
    g = oracle.geometry(1280, 720, 3)
    assert g.scale_upto == 18
    with pytest.raises(ValueError):
        oracle.geometry(64, 48, 5)                                 # a pyramid level would be 0-sized


def test_grayscale_rounding_is_half_even():
    px = np.zeros((1, 4, 4), np.uint8)
    px[0, :, 3] = 255
    px[0, 0, :3] = (255, 255, 255)
    px[0, 1, :3] = (5, 0, 0)       # 1.5 -> 2
    px[0, 2, :3] = (15, 0, 0)      # 4.5 -> 4 (ties to even)
    px[0, 3, :3] = (0, 0, 50)      # 5.5 (fp: 5.5 exactly) -> 6
    g = oracle.grayscale(px)[0]
    assert g[0] == 255 and g[1] == 2 and g[2] == 4 and g[3] == int(np.rint(50 * 0.11))


def test_group_is_order_and_cap_safe(blob):
    f = synth.frame(0, 320, 240)
    res, raw = oracle.detect(f, blob, want_raw=True)
    assert oracle.group(raw, 1) == res
    assert oracle.group(raw, 0) == raw
    assert oracle.group([], 1) == []
