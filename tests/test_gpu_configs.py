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


def test_configs_2_and_3_full_batch_properties(blob):
    """BASELINE configs 2/3 at their full size (1024 x 640x480 per launch), through size-independent properties:
    the batch is 16 distinct frames in a scrambled order, so (a) every copy of a frame must give the same record
    wherever it sits in the batch (idempotence / no cross-frame leakage), (b) each distinct frame must equal the
    oracle, (c) the chunk-pipelined host path must equal the device-resident path, (d) the window memo must not
    change a single output of the 30 track() calls."""
    import torch
    from headtrackr_b200 import Context
    N, U, CALLS = 1024, 16, 30
    uniq = synth.batch(U, 640, 480, start=40)
    rng = np.random.default_rng(7)
    which = rng.integers(0, U, size=N)
    which[:U] = np.arange(U)
    frames = uniq[which]                                            # 1.26 GB
    want = [oracle.detect_track(uniq[u], blob, n_calls=CALLS) for u in range(U)]
    want_rects = [oracle.detect(uniq[u], blob) for u in range(U)]
    c = Context(max_width=640, max_height=480, max_frames=N)
    try:
        dev = torch.from_numpy(frames).cuda()
        runs = {}
        for name, src, memo in (("device_strict", dev, False), ("device_memo", dev, True), ("host_memo", frames, True)):
            c.set_track_memo(memo)
            runs[name] = c.detect_track(src, 5, 1, calc_angles=False, n_calls=CALLS)
        del dev
        ref = runs["device_strict"]
        assert runs["device_memo"] == ref and runs["host_memo"] == ref            # (c), (d)
        dets, found, objs, wins = ref
        first = {}
        for i in range(N):
            u = int(which[i])
            rec = ([tup(d) for d in dets[i]], found[i], objs[i], wins[i])
            if u not in first:
                first[u] = rec
                n_det, fnd, obj = want[u]                                         # (b)
                assert rec[0] == want_rects[u] and len(rec[0]) == n_det and rec[1] == fnd
                if fnd:
                    o = rec[2]
                    assert (o["x"], o["y"], o["width"], o["height"]) == (obj["x"], obj["y"], obj["width"], obj["height"])
                    assert abs(o["angle"] - obj["angle"]) <= 1e-4
            else:
                assert rec == first[u], f"frame {i} (copy of distinct frame {u}) differs from its first copy"   # (a)
        # detect-only entry point on the same batch (config 2), host path
        d2 = c.detect(frames, 5, 1)
        assert all([tup(d) for d in d2[i]] == first[int(which[i])][0] for i in range(N))
    finally:
        c.close()
