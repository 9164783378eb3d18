#!/usr/bin/env python
"""Generate tests/golden/reference_js.json by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/src/{ccv,cascade,camshift,whitebalance}.js) with oracle/jsmini.py over the canvas shim.

This is the pin of the parity chain:
    reference JS (executed here)  ==  C oracle (oracle/ht_oracle.c)  ==  CUDA path (on the GPU box)
The script asserts the first equality while writing the vectors; tests/test_oracle_golden.py re-checks
the oracle against the committed file everywhere, tests/test_gpu_golden.py checks the CUDA path.

Only runs where /root/reference exists (this container).  Takes a few minutes (tree-walking interpreter).
"""
import hashlib
import json
import math
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from headtrackr_b200 import synth  # noqa: E402
from oracle import jsmini  # noqa: E402

REF = Path("/root/reference/src")
OUT = ROOT / "tests" / "golden" / "reference_js.json"

DETECT_CASES = [
    # name, kind, W, H, index, n_faces, interval, min_neighbors
    ("config1_320x240", "faces", 320, 240, 0, None, 5, 1),          # BASELINE.json configs[0]
    ("config1_320x240_raw", "faces", 320, 240, 0, None, 5, 0),      # raw list, reference (scale,q,y,x) order
    ("faces_200x152_i3", "faces", 200, 152, 5, 2, 3, 1),
    ("faces_160x120_mn2", "faces", 160, 120, 3, 1, 5, 2),
    ("odd_171x133", "faces", 171, 133, 9, 1, 5, 1),
    ("noise_160x120", "noise", 160, 120, 1, None, 5, 1),
    ("constant_160x120", "constant", 160, 120, 0, None, 5, 1),
    ("gradient_160x120", "gradient", 160, 120, 0, None, 5, 1),
]

TRACK_CASES = [
    # name, W, H, index, n_faces, calc_angles, n_calls, rect (None = floor of the best detection), track-frame kind
    ("track_160x120_angles", 160, 120, 3, 1, True, 6, None, "same"),
    ("track_160x120_noangles", 160, 120, 3, 1, False, 6, None, "same"),
    ("track_200x152_shifted", 200, 152, 5, 2, False, 5, None, "shifted"),
    ("track_rect_outside", 160, 120, 3, 1, True, 3, [140, 100, 40, 36], "same"),
    ("track_lost", 160, 120, 3, 1, False, 2, None, "constant"),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_reference():
    it = jsmini.Interpreter()
    for f in ("ccv.js", "cascade.js", "camshift.js", "whitebalance.js"):
        it.run((REF / f).read_text())
    return it


def js_detect(it, frame, interval, min_neighbors):
    canvas = jsmini.CanvasShim(frame.copy())
    gray = it.call(it.get(["headtrackr", "ccv", "grayscale"]), jsmini.undefined, canvas)
    res = it.call(it.get(["headtrackr", "ccv", "detect_objects"]), jsmini.undefined, gray,
                  it.get(["headtrackr", "cascade"]), float(interval), float(min_neighbors))
    out = []
    for r in jsmini.to_py(res):
        nb = r.get("neighbors", r.get("neighbor"))
        out.append([r["x"], r["y"], r["width"], r["height"], r["confidence"], int(nb)])
    return out, sha(canvas.pix[..., 0])


def main():
    t_start = time.time()
    it = load_reference()
    blob = synth.load_cascade_blob()
    gold = {"generator": "tools/make_goldens.py (reference JS executed by oracle/jsmini.py over the canvas shim)",
            "reference": "auduno/headtrackr src/ccv.js src/cascade.js src/camshift.js src/whitebalance.js",
            "cascade_blob_sha256": hashlib.sha256(blob).hexdigest(), "detect": [], "track": [], "whitebalance": []}
    for name, kind, W, H, idx, nf, interval, mn in DETECT_CASES:
        t0 = time.time()
        frame = synth.frame(idx, W, H, n_faces=nf, kind=kind)
        rects, gray_sha = js_detect(it, frame, interval, mn)
        want = oracle.detect(frame, blob, interval=interval, min_neighbors=mn)
        assert [tuple(r) for r in rects] == want, f"{name}: reference JS != C oracle\n{rects}\n{want}"
        assert gray_sha == sha(oracle.grayscale(frame)), f"{name}: grayscale differs"
        gold["detect"].append(dict(name=name, kind=kind, W=W, H=H, index=idx, n_faces=nf, interval=interval,
                                   min_neighbors=mn, frame_sha256=sha(frame), gray_sha256=gray_sha, rects=rects))
        print(f"{name}: {len(rects)} rects, JS == oracle  ({time.time() - t0:.1f}s)", flush=True)

    Tracker = it.get(["headtrackr", "camshift", "Tracker"])
    Rectangle = it.get(["headtrackr", "camshift", "Rectangle"])
    for name, W, H, idx, nf, calc, n_calls, rect, tkind in TRACK_CASES:
        t0 = time.time()
        frame = synth.frame(idx, W, H, n_faces=nf)
        if rect is None:
            det = oracle.detect(frame, blob)
            best = det[0]
            for r in det[1:]:
                if r[4] > best[4]:
                    best = r
            rect = [int(math.floor(v)) for v in best[:4]]          # src/facetrackr.js:101-106
        if tkind == "same":
            tframe = frame
        elif tkind == "constant":
            tframe = synth.frame(0, W, H, kind="constant")
        else:
            tframe = np.roll(frame, 3, axis=1)                     # the face moved 3 px to the right
        params = jsmini.JSObject()
        params.props["calcAngles"] = bool(calc)
        trk = Tracker.construct([params])
        c0, c1 = jsmini.CanvasShim(frame.copy()), jsmini.CanvasShim(tframe.copy())
        it.call(trk.get("initTracker"), trk, c0, Rectangle.construct([float(v) for v in rect]))
        ot = oracle.CamshiftTracker(calc_angles=calc)
        ot.init_tracker(frame, *rect)
        calls = []
        for _ in range(n_calls):
            it.call(trk.get("track"), trk, c1)
            o = jsmini.to_py(it.call(trk.get("getTrackObj"), trk))
            w = jsmini.to_py(it.call(trk.get("getSearchWindow"), trk))
            ot.track(tframe)
            oo = ot.track_obj()
            js_obj = [int(o["x"]), int(o["y"]), int(o["width"]), int(o["height"]), o["angle"]]
            js_win = [int(w["x"]), int(w["y"]), int(w["width"]), int(w["height"])]
            assert js_obj[:4] == [oo["x"], oo["y"], oo["width"], oo["height"]], f"{name}: {js_obj} vs {oo}"
            assert js_obj[4] == oo["angle"] or abs(js_obj[4] - oo["angle"]) < 1e-12, f"{name}: angle"
            assert tuple(js_win) == ot.search_window(), f"{name}: window {js_win} vs {ot.search_window()}"
            calls.append(dict(obj=js_obj, window=js_win))
        gold["track"].append(dict(name=name, W=W, H=H, index=idx, n_faces=nf, calc_angles=calc, rect=rect,
                                  track_frame=tkind, frame_sha256=sha(frame), calls=calls))
        print(f"{name}: {n_calls} track() calls, JS == oracle  ({time.time() - t0:.1f}s)", flush=True)

    wb = it.get(["headtrackr", "getWhitebalance"])
    for (W, H, idx, kind) in [(160, 120, 3, "faces"), (64, 48, 1, "noise"), (32, 32, 0, "constant")]:
        frame = synth.frame(idx, W, H, kind=kind)
        v = it.call(wb, jsmini.undefined, jsmini.CanvasShim(frame.copy()))
        assert v == oracle.whitebalance(frame)
        gold["whitebalance"].append(dict(W=W, H=H, index=idx, kind=kind, value=v, frame_sha256=sha(frame)))
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_text(json.dumps(gold, indent=1))
    print(f"wrote {OUT} in {time.time() - t_start:.0f}s")


if __name__ == "__main__":
    main()
