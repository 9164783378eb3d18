#!/usr/bin/env python
"""Golden vectors AT THE BENCHMARK RESOLUTIONS, produced like tools/make_goldens.py by executing the reference's own
JavaScript (oracle/jsmini.py over the canvas shim) and asserting equality with the C oracle on the way:

  * BASELINE configs 2/3 frames: 640x480 bench frames (one plain, one of the rolled copies the bench uses),
    detect_objects(…, 5, 1), and 30 consecutive track() calls on the VJ frame (config 3), one converging stream and
    one that never converges (10 mean-shift iterations in every call);
  * BASELINE config 4: a 1280x720 frame, interval 3.

-> tests/golden/reference_js_large.json; replayed by tests/test_oracle_golden_large.py (CPU) and
tests/test_gpu_golden.py (CUDA).  Takes ~1 h (tree-walking interpreter); only runs where /root/reference exists.
"""
import json
import math
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import make_goldens as mg  # noqa: E402
import oracle  # noqa: E402
from headtrackr_b200 import synth  # noqa: E402
from oracle import jsmini  # noqa: E402

OUT = ROOT / "tests" / "golden" / "reference_js_large.json"

DETECT = [
    # name, W, H, index, roll, interval, min_neighbors
    ("config2_640x480_bench_frame0", 640, 480, 0, 0, 5, 1),
    ("config2_640x480_bench_frame568", 640, 480, 56, 128, 5, 1),     # bench.py: roll(base[568 % 64], (568 // 64) * 16)
    ("config4_1280x720_interval3", 1280, 720, 2, 0, 3, 1),
]
TRACK = [
    # name, W, H, index, roll, n_calls
    ("config3_640x480_30calls_converging", 640, 480, 20, 0, 30),
    ("config3_640x480_30calls_oscillating", 640, 480, 5, 0, 30),
]


def make_frame(W, H, index, roll):
    f = synth.frame(index, W, H)
    return np.roll(f, roll, axis=1) if roll else f


def main():
    only = set(sys.argv[1:])
    t_start = time.time()
    it = mg.load_reference()
    blob = synth.load_cascade_blob()
    gold = json.loads(OUT.read_text()) if OUT.exists() else {
        "generator": "tools/make_goldens_large.py (reference JS executed by oracle/jsmini.py over the canvas shim)",
        "detect": [], "track": []}
    have = {c["name"] for c in gold["detect"] + gold["track"]}
    for name, W, H, idx, roll, interval, mn in DETECT:
        if name in have or (only and name not in only):
            continue
        t0 = time.time()
        frame = make_frame(W, H, idx, roll)
        rects, gray_sha = mg.js_detect(it, frame, interval, mn)
        want = [list(r) for r in oracle.detect(frame, blob, interval, mn)]
        assert rects == want, f"{name}: reference JS != C oracle\n{rects}\n{want}"
        assert gray_sha == mg.sha(oracle.grayscale(frame)), f"{name}: grayscale differs"
        gold["detect"].append(dict(name=name, W=W, H=H, index=idx, roll=roll, interval=interval, min_neighbors=mn,
                                   frame_sha256=mg.sha(frame), gray_sha256=gray_sha, rects=rects))
        OUT.write_text(json.dumps(gold, indent=1))
        print(f"{name}: {len(rects)} rects, JS == oracle  ({time.time() - t0:.0f}s)", flush=True)
    Tracker = it.get(["headtrackr", "camshift", "Tracker"])
    Rectangle = it.get(["headtrackr", "camshift", "Rectangle"])
    for name, W, H, idx, roll, n_calls in TRACK:
        if name in have or (only and name not in only):
            continue
        t0 = time.time()
        frame = make_frame(W, H, idx, roll)
        det = oracle.detect(frame, blob)
        best = det[0]
        for r in det[1:]:
            if r[4] > best[4]:
                best = r
        rect = [int(math.floor(v)) for v in best[:4]]              # src/facetrackr.js:101-106
        params = jsmini.JSObject()
        params.props["calcAngles"] = False
        trk = Tracker.construct([params])
        canvas = jsmini.CanvasShim(frame.copy())
        it.call(trk.get("initTracker"), trk, canvas, Rectangle.construct([float(v) for v in rect]))
        ot = oracle.CamshiftTracker(calc_angles=False)
        ot.init_tracker(frame, *rect)
        calls = []
        for k in range(n_calls):
            it.call(trk.get("track"), trk, canvas)
            o = jsmini.to_py(it.call(trk.get("getTrackObj"), trk))
            w = jsmini.to_py(it.call(trk.get("getSearchWindow"), trk))
            tr = ot.track(frame)
            oo = ot.track_obj()
            js_obj = [int(o["x"]), int(o["y"]), int(o["width"]), int(o["height"]), o["angle"]]
            js_win = [int(w["x"]), int(w["y"]), int(w["width"]), int(w["height"])]
            assert js_obj[:4] == [oo["x"], oo["y"], oo["width"], oo["height"]], f"{name} call {k}: {js_obj} vs {oo}"
            assert abs(js_obj[4] - oo["angle"]) < 1e-12, f"{name} call {k}: angle"
            assert tuple(js_win) == ot.search_window(), f"{name} call {k}: window {js_win} vs {ot.search_window()}"
            calls.append(dict(obj=js_obj, window=js_win, oracle_iterations=int(tr.n_iter)))
            print(f"  {name} call {k}: {js_obj[:4]} iterations {tr.n_iter} ({time.time() - t0:.0f}s)", flush=True)
        gold["track"].append(dict(name=name, W=W, H=H, index=idx, roll=roll, calc_angles=False, rect=rect,
                                  frame_sha256=mg.sha(frame), calls=calls))
        OUT.write_text(json.dumps(gold, indent=1))
        print(f"{name}: {n_calls} track() calls, JS == oracle  ({time.time() - t0:.0f}s)", flush=True)
    if "backprojection" not in gold and not only:
        # getBackProjectionImg (src/camshift.js:177-196) after initTracker + one track() on a 160x120 frame
        W, H, idx = 160, 120, 3
        frame = synth.frame(idx, W, H, n_faces=1)
        best = max(oracle.detect(frame, blob), key=lambda r: r[4])
        rect = [int(math.floor(v)) for v in best[:4]]
        params = jsmini.JSObject()
        params.props["calcAngles"] = False
        trk = Tracker.construct([params])
        canvas = jsmini.CanvasShim(frame.copy())
        it.call(trk.get("initTracker"), trk, canvas, Rectangle.construct([float(v) for v in rect]))
        it.call(trk.get("track"), trk, canvas)
        img = it.call(trk.get("getBackProjectionImg"), trk)
        arr = np.array(img.get("data").buf).reshape(H, W, 4).astype(np.uint8)
        ot = oracle.CamshiftTracker(calc_angles=False)
        ot.init_tracker(frame, *rect)
        ot.track(frame)
        assert np.array_equal(arr, ot.backprojection_img(frame)), "getBackProjectionImg: reference JS != C oracle"
        gold["backprojection"] = [dict(name="backprojection_160x120", W=W, H=H, index=idx, n_faces=1, rect=rect,
                                       frame_sha256=mg.sha(frame), image_sha256=mg.sha(arr),
                                       nonzero=int((arr[..., 0] > 0).sum()))]
        OUT.write_text(json.dumps(gold, indent=1))
        print("backprojection_160x120: JS == oracle", flush=True)
    print(f"wrote {OUT} in {time.time() - t_start:.0f}s")


if __name__ == "__main__":
    main()
