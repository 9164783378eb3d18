"""GPU: the top-level headtrackr.Tracker mirror (src/main.js) running on the CUDA library, against the reference's
own main.js event stream (tests/golden/reference_js_main.json): whitebalance gate, detection, tracking, smoothing,
head positions, lost face -> re-detection -> found again."""
import pytest

from headtrackr_b200 import facetrackr
from test_host_main import GOLD_M, check_events, run_main, same

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", GOLD_M["cases"], ids=lambda c: c["name"])
def test_main_tracker_on_cuda_matches_reference_js(ctx, case):
    steps, stop_events, ht = run_main(case, facetrackr.CudaBackend(ctx))
    assert [s["status"] for s in steps] == [s["status"] for s in case["steps"]]
    for g, w in zip(steps, case["steps"]):
        check_events(g["events"], w["events"])
    check_events(stop_events, case["stop_events"])
    assert same(ht.getFOV(), case["fov"])
