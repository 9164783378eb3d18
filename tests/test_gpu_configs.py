"""GPU: the BASELINE.json configurations that are parity cases rather than bench lines.

config 3: detect once, then 30 CAMShift track() calls per frame (state carried)       -> test_gpu_track.py
config 4: 1280x720, "4-scale pyramid" (interval=3: 4 scales per octave), per-frame DP -> here (+ test_distributed_gloo.py)
config 5: independent 640x480 streams, detect -> track steady state, redetect on loss -> here
"""
import math

import numpy as np
import pytest

import oracle
from headtrackr_b200 import Canvas, facetrackr, synth
from test_host_logic import OracleBackend

pytestmark = pytest.mark.gpu


def tup(d):
    return (d["x"], d["y"], d["width"], d["height"], d["confidence"], d["neighbors"])


@pytest.mark.parametrize("interval", [3, 5])
def test_config4_1280x720(ctx, blob, interval):
    frames = synth.batch(2, 1280, 720, start=300)
    got = ctx.detect(frames, interval, 1)
    info = ctx.plan_info(1280, 720, interval)
    assert info["scale_upto"] == (18 if interval == 3 else 27)
    for i in range(2):
        want = oracle.detect(frames[i], blob, interval=interval)
        assert [tup(d) for d in got[i]] == want and len(want) >= 1


def moving_stream(n, W=640, H=480, seed=500):
    """Frames t = 0..n-1 of one synthetic stream: the background of frame `seed`, the faces translated 3 px/frame."""
    base = synth.frame(seed, W, H, n_faces=1)
    return [np.roll(base, (2 * t, 3 * t), axis=(0, 1)) for t in range(n)]


def run_stream(backend, frames):
    """facetrackr state machine (src/facetrackr.js) over a stream; returns the emitted events and the modes."""
    events, modes = [], []
    canvas = Canvas(frames[0])
    ft = facetrackr.Tracker({"whitebalancing": False}, backend=backend)
    ft.addEventListener(lambda e: events.append({k: v for k, v in e.items() if k != "time"}))
    ft.init(canvas)
    for f in frames:
        canvas.pixels = f
        ft.track()
        modes.append(ft.getTrackingObject().detection)
    return events, modes


def test_config5_stream_events_match_oracle(ctx, blob):
    frames = moving_stream(8)
    ev_gpu, modes_gpu = run_stream(facetrackr.CudaBackend(ctx), frames)
    ev_cpu, modes_cpu = run_stream(OracleBackend(blob), frames)
    assert modes_gpu == modes_cpu == ["VJ"] + ["CS"] * 7
    assert len(ev_gpu) == 7
    for a, b in zip(ev_gpu, ev_cpu):
        assert (a["x"], a["y"], a["width"], a["height"], a["confidence"], a["detection"]) == \
               (b["x"], b["y"], b["width"], b["height"], b["confidence"], b["detection"])
        assert abs(a["angle"] - b["angle"]) <= 1e-4


def test_config5_many_streams_in_one_batch(ctx, blob):
    """8 independent streams advanced together: one ht_detect_track for the VJ frame, then one ht_track per time step."""
    n_streams, steps = 8, 4
    streams = [moving_stream(steps + 1, seed=600 + s) for s in range(n_streams)]
    first = np.stack([s[0] for s in streams])
    dets, found, objs, wins = ctx.detect_track(first, 5, 1, calc_angles=False, n_calls=0)
    assert all(found)
    trackers = []
    for s in range(n_streams):
        res = oracle.detect(first[s], blob)
        cand = res[0]
        for r in res[1:]:
            if r[4] > cand[4]:
                cand = r
        ot = oracle.CamshiftTracker(calc_angles=False)
        ot.init_tracker(first[s], *[int(math.floor(v)) for v in cand[:4]])
        trackers.append(ot)
    for t in range(1, steps + 1):
        batch = np.stack([s[t] for s in streams])
        objs, wins = ctx.track(batch)
        for s in range(n_streams):
            trackers[s].track(batch[s])
            w = trackers[s].track_obj()
            assert (objs[s]["x"], objs[s]["y"], objs[s]["width"], objs[s]["height"]) == (w["x"], w["y"], w["width"], w["height"])
            assert wins[s] == trackers[s].search_window()


def test_facetrackr_on_cuda_matches_reference_js(ctx):
    """The reference's own facetrackr.js event stream (tests/golden/reference_js_post.json) reproduced with the
    CUDA library underneath the host state machine."""
    from test_host_post import GOLD as POST, check_steps, run_facetrackr
    for case in POST["facetrackr"]:
        check_steps(run_facetrackr(case, facetrackr.CudaBackend(ctx)), case["steps"])
