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
