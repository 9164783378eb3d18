"""GPU: the CUDA path (through the C ABI) against the golden vectors produced by executing the
reference's own JavaScript (tests/golden/reference_js.json) — no oracle in between."""
import json
from pathlib import Path

import numpy as np
import pytest

from headtrackr_b200 import Canvas, camshift, ccv, facetrackr, synth
from test_oracle_golden import GOLD, track_frames
from test_oracle_golden_large import GOLD_L, large_frame

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", GOLD["detect"], ids=lambda c: c["name"])
def test_detect_golden_cuda(ctx, case):
    f = synth.frame(case["index"], case["W"], case["H"], n_faces=case["n_faces"], kind=case["kind"])
    got = ctx.detect(f, case["interval"], case["min_neighbors"])[0]
    got = [[d["x"], d["y"], d["width"], d["height"], d["confidence"], d.get("neighbors", d.get("neighbor"))] for d in got]
    assert got == case["rects"]                                   # bit-exact doubles, reference order


@pytest.mark.parametrize("case", GOLD["track"], ids=lambda c: c["name"])
def test_track_golden_cuda(ctx, case):
    f, t = track_frames(case)
    ctx.track_init(f, [case["rect"]], calc_angles=case["calc_angles"])
    for call in case["calls"]:
        objs, wins = ctx.track(t)
        o = objs[0]
        assert [o["x"], o["y"], o["width"], o["height"]] == call["obj"][:4]
        assert abs(o["angle"] - call["obj"][4]) <= 1e-4          # north_star tolerance
        assert list(wins[0]) == call["window"]


@pytest.mark.parametrize("case", GOLD_L["detect"], ids=lambda c: c["name"])
def test_detect_large_golden_cuda(ctx, case):
    """Benchmark-resolution frames (BASELINE configs 2 and 4) against the reference JS itself."""
    got = ctx.detect(large_frame(case), case["interval"], case["min_neighbors"])[0]
    got = [[d["x"], d["y"], d["width"], d["height"], d["confidence"], d["neighbors"]] for d in got]
    assert got == case["rects"]


@pytest.mark.parametrize("case", GOLD_L["track"], ids=lambda c: c["name"])
@pytest.mark.parametrize("memo", [False, True], ids=["strict", "memo"])
def test_track_large_golden_cuda(ctx, case, memo):
    """BASELINE config 3: 30 track() calls on a 640x480 VJ frame, call by call and as one launch, against the JS."""
    f = large_frame(case)
    try:
        ctx.set_track_memo(memo)
        ctx.track_init(f, [case["rect"]], calc_angles=case["calc_angles"])
        for call in case["calls"]:
            objs, wins = ctx.track(f)
            o = objs[0]
            assert [o["x"], o["y"], o["width"], o["height"]] == call["obj"][:4]
            assert abs(o["angle"] - call["obj"][4]) <= 1e-4
            assert list(wins[0]) == call["window"]
        ctx.track_init(f, [case["rect"]], calc_angles=case["calc_angles"])
        objs, wins = ctx.track(f, n_calls=len(case["calls"]))
        last = case["calls"][-1]
        assert [objs[0]["x"], objs[0]["y"], objs[0]["width"], objs[0]["height"]] == last["obj"][:4]
        assert list(wins[0]) == last["window"]
    finally:
        ctx.set_track_memo(True)


@pytest.mark.parametrize("case", GOLD_L.get("backprojection", []), ids=lambda c: c["name"])
def test_backprojection_golden_cuda(ctx, case):
    import hashlib
    f = synth.frame(case["index"], case["W"], case["H"], n_faces=case["n_faces"])
    ctx.track_init(f, [case["rect"]], calc_angles=False)
    img = ctx.backprojection(f, 0)
    assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == case["image_sha256"]


def test_whitebalance_golden_cuda(ctx):
    for c in GOLD["whitebalance"]:
        f = synth.frame(c["index"], c["W"], c["H"], kind=c["kind"])
        assert ctx.whitebalance(f)[0] == c["value"]


def test_reference_api_mirror(ctx):
    """headtrackr.ccv / camshift / facetrackr names and shapes on top of the CUDA library."""
    case = GOLD["detect"][0]
    f = synth.frame(case["index"], case["W"], case["H"], n_faces=case["n_faces"], kind=case["kind"])
    canvas = Canvas(f)
    comp = ccv.detect_objects(ccv.grayscale(canvas), None, 5, 1, context=ctx)
    assert [[d["x"], d["y"], d["width"], d["height"], d["confidence"], d["neighbors"]] for d in comp] == case["rects"]
    events = []
    ft = facetrackr.Tracker({"whitebalancing": False}, backend=facetrackr.CudaBackend(ctx))
    ft.addEventListener(events.append)
    ft.init(canvas)
    ft.track()
    assert ft.getTrackingObject().detection == "VJ"
    ft.track()
    o = ft.getTrackingObject()
    assert o.detection == "CS" and o.width > 0 and len(events) == 1 and events[0]["confidence"] == 1
    trk = camshift.Tracker({"calcAngles": True}, context=ctx, slot=3)
    best = max(comp, key=lambda d: d["confidence"])
    trk.initTracker(canvas, camshift.Rectangle(int(best["x"]), int(best["y"]), int(best["width"]), int(best["height"])))
    trk.track(canvas)
    assert trk.getTrackObj().width > 0 and trk.getSearchWindow().width > 0
    assert trk.getBackProjectionImg().shape == f.shape
