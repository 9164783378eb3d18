#!/usr/bin/env python
"""Goldens for the host-side post-processing (src/smoother.js, src/headposition.js) from the reference's own
source executed by oracle/jsmini.py -> tests/golden/reference_js_post.json.  jsmini's Math uses the C libm
(V8 uses fdlibm): sin/cos/tan/atan may differ from a browser in the last ulp, so the tests compare to 1e-12."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import jsmini  # noqa: E402

REF = Path("/root/reference/src")
OUT = ROOT / "tests" / "golden" / "reference_js_post.json"

FACES = [  # centre x, y, width, height in a 320x240 camera image: centre, edges, corners, moving
    dict(x=160, y=120, width=80, height=95), dict(x=150, y=118, width=84, height=99), dict(x=42, y=120, width=76, height=90),
    dict(x=285, y=130, width=70, height=84), dict(x=160, y=44, width=78, height=92), dict(x=170, y=200, width=72, height=86),
    dict(x=36, y=40, width=70, height=80), dict(x=290, y=205, width=64, height=76), dict(x=100, y=100, width=60, height=72),
]


def obj(d):
    o = jsmini.JSObject()
    o.props.update({k: float(v) for k, v in d.items()})
    return o


def main():
    it = jsmini.Interpreter()
    it.run("headtrackr.headposition = {};")
    for f in ("smoother.js", "headposition.js"):
        it.run((REF / f).read_text())
    gold = {"generator": "tools/make_goldens_post.py (reference JS executed by oracle/jsmini.py)", "faces": FACES,
            "smoother": [], "headposition": []}
    for alpha, interval in [(0.35, 35.0), (0.8, 20.0)]:
        sm = it.get(["headtrackr", "Smoother"]).construct([alpha, interval])
        it.call(sm.get("init"), sm, obj(FACES[0]))
        outs = []
        for f in FACES[1:]:
            r = jsmini.to_py(it.call(sm.get("smooth"), sm, obj(f)))
            outs.append([r["x"], r["y"], r["width"], r["height"]])     # z is NaN by construction (never initialised)
        gold["smoother"].append(dict(alpha=alpha, interval=interval, out=outs))
    for params in [None, {"edgecorrection": False}, {"fov": 60.0}, {"distance_to_screen": 50.0, "distance_from_camera_to_screen": 9.0}]:
        args = [obj(FACES[0]), 320.0, 240.0]
        if params is not None:
            p = jsmini.JSObject()
            p.props.update(params)
            args.append(p)
        hp = it.get(["headtrackr", "headposition", "Tracker"]).construct(args)
        outs = []
        for f in FACES[1:]:
            r = jsmini.to_py(it.call(hp.get("track"), hp, obj(f)))
            outs.append([r["x"], r["y"], r["z"]])
        gold["headposition"].append(dict(params=params, fov=it.call(hp.get("getFOV"), hp), out=outs))
    # ---- facetrackr state machine (src/facetrackr.js) over a short synthetic stream ----
    import numpy as np
    from headtrackr_b200 import synth
    for f in ("ccv.js", "cascade.js", "camshift.js", "whitebalance.js", "facetrackr.js"):
        it.run((REF / f).read_text())
    gold["facetrackr"] = []
    for name, wb, n_frames in [("vj_cs", False, 5), ("wb_vj_cs", True, 18)]:
        base = synth.frame(3, 160, 120, n_faces=1)
        frames = [np.roll(base, (t, 2 * t), axis=(0, 1)) for t in range(n_frames)]
        if wb:   # whitebalance gate: identical frames until the 15-sample window is stable (src/facetrackr.js:79-95)
            frames = [base] * 16 + frames[:2]
        params = jsmini.JSObject()
        params.props["whitebalancing"] = wb
        canvas = jsmini.CanvasShim(frames[0].copy())
        ft = it.get(["headtrackr", "facetrackr", "Tracker"]).construct([params])
        it.call(ft.get("init"), ft, canvas)
        it.events.clear()
        steps = []
        for fr in frames:
            canvas.pix = fr.copy()
            n_before = len(it.events)
            it.call(ft.get("track"), ft)
            o = jsmini.to_py(it.call(ft.get("getTrackingObject"), ft))
            ev = [jsmini.to_py(e) for e in it.events[n_before:]]
            for e in ev:
                e.pop("time", None)
            steps.append(dict(detection=o["detection"], x=o["x"], y=o["y"], width=o["width"], height=o["height"],
                              confidence=o["confidence"], events=ev))
        gold["facetrackr"].append(dict(name=name, whitebalancing=wb, n_frames=len(frames), steps=steps))
        print(name, [s_["detection"] for s_ in steps])
    OUT.write_text(json.dumps(gold, indent=1))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
