"""CPU: host-side post-processing mirrors (Smoother, headposition.Tracker) against the reference's own JS
(tests/golden/reference_js_post.json, produced by executing src/smoother.js and src/headposition.js)."""
import json
import math
from pathlib import Path

from headtrackr_b200 import headposition, smoother

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_js_post.json").read_text())
FACES = GOLD["faces"]


def close(a, b):
    return all(abs(x - y) <= 1e-12 * max(1.0, abs(y)) for x, y in zip(a, b))


def test_smoother_matches_reference_js():
    for case in GOLD["smoother"]:
        sm = smoother.Smoother(case["alpha"], case["interval"])
        assert sm.smooth(dict(FACES[1])) is False                 # not initialised yet (src/smoother.js:57)
        sm.init(dict(FACES[0]))
        for f, want in zip(FACES[1:], case["out"]):
            r = sm.smooth(dict(f))
            assert close([r["x"], r["y"], r["width"], r["height"]], want)
            assert math.isnan(r["z"])                              # z is never initialised by the caller (src/main.js:259)


def test_headposition_matches_reference_js():
    for case in GOLD["headposition"]:
        events = []
        hp = headposition.Tracker(dict(FACES[0]), 320, 240, case["params"])
        hp.addEventListener(events.append)
        assert abs(hp.getFOV() - case["fov"]) <= 1e-12 * case["fov"]
        for f, want in zip(FACES[1:], case["out"]):
            r = hp.track(dict(f))
            assert close([r.x, r.y, r.z], want)
        assert len(events) == len(FACES) - 1 and set(events[0]) == {"type", "x", "y", "z"}
        assert events[-1]["type"] == "headtrackingEvent"


def _frames(case):
    import numpy as np
    from headtrackr_b200 import synth
    base = synth.frame(3, 160, 120, n_faces=1)
    if case["whitebalancing"]:
        return [base] * 16 + [np.roll(base, (t, 2 * t), axis=(0, 1)) for t in range(2)]
    return [np.roll(base, (t, 2 * t), axis=(0, 1)) for t in range(case["n_frames"])]


def run_facetrackr(case, backend):
    from headtrackr_b200 import Canvas, facetrackr
    frames = _frames(case)
    events = []
    canvas = Canvas(frames[0])
    ft = facetrackr.Tracker({"whitebalancing": case["whitebalancing"]}, backend=backend)
    ft.addEventListener(lambda e: events.append({k: v for k, v in e.items() if k != "time"}))
    ft.init(canvas)
    steps = []
    for f in frames:
        canvas.pixels = f
        n0 = len(events)
        ft.track()
        o = ft.getTrackingObject()
        steps.append(dict(detection=o.detection, x=o.x, y=o.y, width=o.width, height=o.height, confidence=o.confidence,
                          events=events[n0:]))
    return steps


def check_steps(got, want):
    assert [s["detection"] for s in got] == [s["detection"] for s in want]
    for g, w in zip(got, want):
        for k in ("x", "y", "width", "height", "confidence"):
            assert float(g[k]) == float(w[k]), (k, g, w)
        assert len(g["events"]) == len(w["events"])
        for ge, we in zip(g["events"], w["events"]):
            assert set(ge) == set(we)                              # facetrackingEvent shape (src/facetrackr.js:112-125)
            for k in we:
                assert ge[k] == we[k] or (k == "angle" and abs(ge[k] - we[k]) <= 1e-12), (k, ge, we)


def test_facetrackr_state_machine_matches_reference_js(blob):
    """The Python mirror of src/facetrackr.js driven by the CPU oracle == the reference's own facetrackr.js
    (WB gate, VJ -> CS hand-off without an event on the hand-off frame, CS events)."""
    from test_host_logic import OracleBackend
    for case in GOLD["facetrackr"]:
        check_steps(run_facetrackr(case, OracleBackend(blob)), case["steps"])
