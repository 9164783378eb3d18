"""CPU: the Python mirror of headtrackr.Tracker (src/main.js) against the reference's own main.js executed by
oracle/jsmini.py (tests/golden/reference_js_main.json, tools/make_goldens_main.py): every status
(whitebalance -> detecting -> found -> tracking -> redetecting -> found), every facetrackingEvent and every
headtrackingEvent of a 42-frame stream in which the face disappears and comes back."""
import json
import math
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
GOLD_M = json.loads((Path(__file__).resolve().parent / "golden" / "reference_js_main.json").read_text())


def run_main(case, backend):
    import make_goldens_main as mg
    from headtrackr_b200 import Canvas, main
    import numpy as np
    spec = [tuple(s["frame"]) for s in case["steps"]]
    assert spec == [tuple(x) for x in mg.stream_frames()]
    W, H = GOLD_M["width"], GOLD_M["height"]
    video = Canvas(mg.make_frame(*spec[0]))
    canvas = Canvas(np.zeros((H, W, 4), np.uint8))
    clock = [1.0e12]
    ht = main.Tracker(dict(case["params"], ui=False), backend=backend, clock=lambda: clock[0])
    log = []
    for t in ("headtrackrStatus", "facetrackingEvent", "headtrackingEvent"):
        ht.addEventListener(t, lambda e: log.append({k: v for k, v in e.items() if k != "time"}))
    ht.init(video, canvas, False)
    steps = []
    for n, (kind, t) in enumerate(spec):
        video.pixels = mg.make_frame(kind, t)
        clock[0] += 35.0
        n0 = len(log)
        if n == 0:
            assert ht.start() is True
        else:
            assert ht.step() is True
        steps.append(dict(status=ht.status, events=log[n0:]))
    n0 = len(log)
    ht.stop()
    return steps, log[n0:], ht


def same(a, b):
    if isinstance(a, float) or isinstance(b, float):
        a, b = float(a), float(b)
        return (a != a and b != b) or a == b or abs(a - b) <= 1e-9 * max(1.0, abs(a), abs(b))
    return a == b


def check_events(got, want):
    assert [e["type"] for e in got] == [e["type"] for e in want]
    for g, w in zip(got, want):
        assert set(g) == set(w), (g, w)
        for k in w:
            assert same(g[k], w[k]), (k, g, w)


@pytest.mark.parametrize("case", GOLD_M["cases"], ids=lambda c: c["name"])
def test_main_tracker_matches_reference_js(case, blob):
    from test_host_logic import OracleBackend
    steps, stop_events, ht = run_main(case, OracleBackend(blob))
    assert [s["status"] for s in steps] == [s["status"] for s in case["steps"]]
    for n, (g, w) in enumerate(zip(steps, case["steps"])):
        check_events(g["events"], w["events"])
    check_events(stop_events, case["stop_events"])
    assert same(ht.getFOV(), case["fov"])
    assert ht.status == "stopped" and ht.step() is False           # stop() cancels the pending timer
    seen = {s["status"] for s in case["steps"]}
    assert {"whitebalance", "detecting", "found", "tracking", "redetecting"} <= seen
    assert any(e["type"] == "headtrackingEvent" for s in case["steps"] for e in s["events"])
