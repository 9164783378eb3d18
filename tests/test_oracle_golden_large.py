"""CPU: the C oracle against the reference-JS golden vectors at the benchmark resolutions (640x480 bench frames with
30 track() calls, 1280x720 interval 3) - tests/golden/reference_js_large.json, tools/make_goldens_large.py."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from headtrackr_b200 import synth

PATH = Path(__file__).resolve().parent / "golden" / "reference_js_large.json"
GOLD_L = json.loads(PATH.read_text()) if PATH.exists() else {"detect": [], "track": []}


def large_frame(case):
    f = synth.frame(case["index"], case["W"], case["H"])
    if case.get("roll"):
        f = np.roll(f, case["roll"], axis=1)
    assert hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest() == case["frame_sha256"], "synthetic frame changed"
    return f


def test_large_goldens_are_present():
    assert len(GOLD_L["detect"]) >= 2 and len(GOLD_L["track"]) >= 1


@pytest.mark.parametrize("case", GOLD_L["detect"], ids=lambda c: c["name"])
def test_detect_large_golden_oracle(case, blob):
    f = large_frame(case)
    got = [list(r) for r in oracle.detect(f, blob, case["interval"], case["min_neighbors"])]
    assert got == case["rects"]


@pytest.mark.parametrize("case", GOLD_L["track"], ids=lambda c: c["name"])
def test_track_large_golden_oracle(case):
    f = large_frame(case)
    ot = oracle.CamshiftTracker(calc_angles=case["calc_angles"])
    ot.init_tracker(f, *case["rect"])
    for call in case["calls"]:
        ot.track(f)
        o = ot.track_obj()
        assert [o["x"], o["y"], o["width"], o["height"]] == call["obj"][:4]
        assert abs(o["angle"] - call["obj"][4]) <= 1e-12
        assert list(ot.search_window()) == call["window"]


@pytest.mark.parametrize("case", GOLD_L.get("backprojection", []), ids=lambda c: c["name"])
def test_backprojection_golden_oracle(case):
    """getBackProjectionImg (src/camshift.js:177-196): the image the reference JS produced, by hash."""
    f = synth.frame(case["index"], case["W"], case["H"], n_faces=case["n_faces"])
    assert hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest() == case["frame_sha256"]
    ot = oracle.CamshiftTracker(calc_angles=False)
    ot.init_tracker(f, *case["rect"])
    ot.track(f)
    img = ot.backprojection_img(f)
    assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == case["image_sha256"]
    assert int((img[..., 0] > 0).sum()) == case["nonzero"]
