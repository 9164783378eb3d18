"""CPU: the head-position epilogue of ht_stream_step (head_step in ht_track.cuh: Smoother + stable-diagonal wait +
headposition.Tracker, src/main.js:246-300, src/smoother.js, src/headposition.js) run on the host over the CS results
of the reference's own src/main.js run (tests/golden/reference_js_main.json): every `headtrackingEvent {x, y, z}` and
every "found" status of the 42-frame stream - face lost and found again, with and without smoothing / a given fov."""
import ctypes as C
import math

import pytest

from test_cascade_host import st  # noqa: F401  (fixture: the host-only build of ht_api.cu)
from test_host_main import GOLD_M


class HeadParams(C.Structure):
    _fields_ = [("smoothing", C.c_int32), ("head_position", C.c_int32), ("edgecorrection", C.c_int32), ("pad_", C.c_int32),
                ("alpha", C.c_double), ("fov_deg", C.c_double), ("camera_offset", C.c_double), ("distance_to_screen", C.c_double)]


class HeadEvent(C.Structure):
    _fields_ = [("valid", C.c_int32), ("status", C.c_int32), ("x", C.c_double), ("y", C.c_double), ("z", C.c_double),
                ("fx", C.c_double), ("fy", C.c_double), ("fwidth", C.c_double), ("fheight", C.c_double)]


def params_of(case):
    p = case["params"] or {}
    return HeadParams(int(p.get("smoothing", True)), int(p.get("headPosition", True)), 1, 0, 0.35,
                      float(p["fov"]) if p.get("fov") is not None else 0.0, float(p.get("cameraOffset", 11.5)), 60.0)


def close(a, b):
    return (a != a and b != b) or abs(a - b) <= 1e-9 * max(1.0, abs(a), abs(b))


@pytest.mark.parametrize("case", GOLD_M["cases"], ids=lambda c: c["name"])
def test_head_epilogue_replays_main_js(st, case):
    steps = case["steps"]
    cs = (C.c_double * (5 * len(steps)))()
    for i, s in enumerate(steps):
        ft = [e for e in s["events"] if e["type"] == "facetrackingEvent"]
        if ft:
            cs[5 * i: 5 * i + 5] = [1.0, ft[0]["x"], ft[0]["y"], ft[0]["width"], ft[0]["height"]]
    out = (HeadEvent * len(steps))()
    p = params_of(case)
    st.ht_selftest_head.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    assert st.ht_selftest_head(C.byref(p), len(steps), cs, GOLD_M["width"], GOLD_M["height"], out) == 0
    n_head = n_found = 0
    for i, s in enumerate(steps):
        want = [e for e in s["events"] if e["type"] == "headtrackingEvent"]
        assert out[i].valid == len(want), i
        if want:
            n_head += 1
            assert close(out[i].x, want[0]["x"]) and close(out[i].y, want[0]["y"]) and close(out[i].z, want[0]["z"]), (i, want)
        found = any(e["type"] == "headtrackrStatus" and e["status"] == "found" for e in s["events"])
        assert bool(out[i].status & 1) == found, i
        n_found += int(found)
    assert n_head >= 5 and n_found >= 2          # found, lost, found again
    assert math.isfinite(out[len(steps) - 1].fx)
