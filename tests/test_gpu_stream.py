"""GPU: ht_stream_step - facetrackr's state machine for n streams on the device - against the
reference's own JS event streams (tests/golden/reference_js_post.json: src/facetrackr.js; reference_js_main.json:
src/main.js with a face that disappears and comes back), replayed frame by frame."""
import sys
from pathlib import Path

import numpy as np
import pytest

from headtrackr_b200 import Context, synth
from test_host_post import GOLD as POST, _frames

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def check_step(ev, want):
    assert ev["detection"] == want["detection"]
    for k in ("x", "y", "width", "height", "confidence"):
        assert float(ev[k]) == float(want[k]), (k, ev, want)
    for we in want.get("events", []):          # facetrackingEvent is sent exactly for CS results
        assert ev["detection"] == "CS" and abs(ev["angle"] - we["angle"]) <= 1e-12
    assert (ev["detection"] == "CS") == (len(want.get("events", [])) == 1)


@pytest.mark.parametrize("case", POST["facetrackr"], ids=lambda c: c["name"])
def test_stream_step_replays_facetrackr_js(case):
    frames = _frames(case)
    steps = case["steps"]
    i0 = next(i for i, s in enumerate(steps) if s["detection"] == "VJ")     # the WB gate stays on the host
    frames, steps = frames[i0:], steps[i0:]
    H, W = frames[0].shape[:2]
    blank = np.full_like(frames[0], 90)
    blank[..., 3] = 255
    c = Context(max_width=W, max_height=H, max_frames=8)
    try:
        # stream 0: the golden stream; stream 1: never a face; stream 2: the golden stream one frame late;
        # streams 3..5: more copies, so that two frame quads are in play and modes are mixed inside a quad
        for t in range(len(frames) + 1):
            f_now = frames[t] if t < len(frames) else frames[-1]
            f_late = blank if t == 0 else frames[t - 1]
            batch = np.stack([f_now, blank, f_late, f_now, f_late, f_now])
            ev = c.stream_step(batch, 5, 1, calc_angles=False)
            if t < len(frames):
                check_step(ev[0], steps[t])
                check_step(ev[3], steps[t])
                check_step(ev[5], steps[t])
            assert ev[1]["detection"] == "VJ" and ev[1]["confidence"] == -10000 and not ev[1]["found"]
            if t >= 1:
                check_step(ev[2], steps[t - 1])
                check_step(ev[4], steps[t - 1])
        assert any(s["detection"] == "CS" for s in steps)
        # a reset stream starts over with detection
        c.stream_reset(0, 1)
        ev = c.stream_step(np.stack([frames[0]] * 6), 5, 1)
        check_step(ev[0], steps[0])
        assert ev[0]["found"] and ev[3]["detection"] == "CS"
    finally:
        c.close()


def test_stream_step_replays_main_js_lost_and_found():
    """src/main.js over 42 frames: detecting -> found -> tracking -> (face gone) redetecting -> found again."""
    sys.path.insert(0, str(ROOT / "tools"))
    import make_goldens_main as mg
    from test_host_main import GOLD_M
    case = next(c for c in GOLD_M["cases"] if c["name"] == "default")
    spec = [tuple(s["frame"]) for s in case["steps"]]
    i0 = next(i for i, s in enumerate(case["steps"]) if s["status"] == "detecting")
    W, H = GOLD_M["width"], GOLD_M["height"]
    c = Context(max_width=W, max_height=H, max_frames=4)
    try:
        n_cs = n_lost = 0
        for i in range(i0, len(spec)):
            f = mg.make_frame(*spec[i])
            ev = c.stream_step(np.stack([f, f]), 5, 1, calc_angles=bool(case["params"].get("calcAngles", False)))[1]
            want = [e for e in case["steps"][i]["events"] if e["type"] == "facetrackingEvent"]
            assert (ev["detection"] == "CS") == (len(want) == 1), (i, ev, want)
            if want:
                n_cs += 1
                w = want[0]
                assert (ev["x"], ev["y"], ev["width"], ev["height"], ev["confidence"]) == (w["x"], w["y"], w["width"], w["height"], w["confidence"])
                assert abs(ev["angle"] - w["angle"]) <= 1e-12
            redetect = any(e.get("status") == "redetecting" for e in case["steps"][i]["events"])
            assert ev["lost"] == redetect, (i, ev)
            n_lost += int(redetect)
        assert n_cs >= 10 and n_lost >= 1
    finally:
        c.close()


@pytest.mark.parametrize("case_name", ["default", "no_smoothing_fov"])
def test_stream_step_head_replays_main_js_head_events(case_name):
    """f3: every headtrackingEvent {x, y, z} and "found" status of the reference's src/main.js run, produced by the
    head-position epilogue of ht_stream_step on the device (smoother + stable diagonal + headposition.Tracker)."""
    sys.path.insert(0, str(ROOT / "tools"))
    import make_goldens_main as mg
    from test_host_main import GOLD_M
    case = next(c for c in GOLD_M["cases"] if c["name"] == case_name)
    p = case["params"] or {}
    spec = [tuple(s["frame"]) for s in case["steps"]]
    i0 = next(i for i, s in enumerate(case["steps"]) if s["status"] == "detecting")
    W, H = GOLD_M["width"], GOLD_M["height"]
    c = Context(max_width=W, max_height=H, max_frames=4)
    try:
        c.stream_head_config(smoothing=p.get("smoothing", True), fov=p.get("fov"), camera_offset=p.get("cameraOffset", 11.5),
                             head_position=p.get("headPosition", True))
        c.stream_reset(0, 3)
        n_head = 0
        for i in range(i0, len(spec)):
            f = mg.make_frame(*spec[i])
            ev, heads = c.stream_step_head(np.stack([f, f, f]), 5, 1, calc_angles=bool(p.get("calcAngles", False)))
            want = [e for e in case["steps"][i]["events"] if e["type"] == "headtrackingEvent"]
            for k in (0, 2):
                assert heads[k]["valid"] == (len(want) == 1), (i, k)
                if want:
                    for key in ("x", "y", "z"):
                        a, b = heads[k][key], want[0][key]
                        assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (i, key, a, b)
                found = any(e["type"] == "headtrackrStatus" and e["status"] == "found" for e in case["steps"][i]["events"])
                assert heads[k]["found"] == found, (i, k)
            n_head += len(want)
        assert n_head >= 5
    finally:
        c.close()
