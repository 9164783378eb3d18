#!/usr/bin/env python
"""Golden event stream of the reference's TOP-LEVEL object, headtrackr.Tracker (src/main.js), executed by
oracle/jsmini.py on top of the unmodified ccv / cascade / camshift / whitebalance / facetrackr / smoother /
headposition sources -> tests/golden/reference_js_main.json.

main.js is executed with two cuts, both of code that never runs when `init(video, canvas, false)` is used and that
jsmini cannot parse (regex literals, `throw`, `instanceof`): the getUserMedia block of `init` (src/main.js:100-157) and
everything from the `Function.prototype.bind` polyfill down (src/main.js:381-430; jsmini provides `bind`).  The
`<video>` element is a canvas shim holding the current frame; `window.setTimeout` only records its callback, and the
harness fires the newest one once per frame - the browser's 20 ms timer.

The stream (160x120): a face is detected (VJ), tracked (CS) until the head diagonal is stable and head positions
are emitted, then disappears (lost -> "redetecting" -> a fresh facetrackr), stays away for a few frames, and comes
back ("found" again, head position re-created with the remembered field of view).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from headtrackr_b200 import synth  # noqa: E402
from oracle import jsmini  # noqa: E402

REF = Path("/root/reference/src")
OUT = ROOT / "tests" / "golden" / "reference_js_main.json"
W, H = 160, 120


def cut_main():
    s = (REF / "main.js").read_text()
    i = s.index("if (setupVideo === undefined || setupVideo == true) {")
    k, depth = s.index("{", i), 0
    while True:
        if s[k] == "{":
            depth += 1
        elif s[k] == "}":
            depth -= 1
            if depth == 0:
                break
        k += 1
    s = s[:i] + s[k + 1:]
    return s[: s.index("// bind shim")]


def stream_frames():
    """(kind, index/shift) list shared with tests/test_host_main.py"""
    # 15 frames fill the whitebalance window (src/facetrackr.js:79-95; rolling a frame keeps its mean), then VJ, CS ...
    spec = [("face", t) for t in range(28)] + [("empty", 0)] * 3 + [("face", 40 + t) for t in range(11)]
    return spec


def make_frame(kind, t):
    if kind == "empty":
        return synth.frame(0, W, H, kind="constant")
    base = synth.frame(3, W, H, n_faces=1)
    return np.roll(base, (t % 3, (2 * t) % 5), axis=(0, 1))          # a little jitter, like a hand-held head


def event_record(e):
    d = jsmini.to_py(e)
    d.pop("time", None)
    d.pop("initEvent", None)
    return d


def main():
    it = jsmini.Interpreter()
    it.run(cut_main())                     # main.js comes first in the bundle: it declares `var headtrackr = {}`
    it.run("headtrackr.headposition = {};")
    for f in ("ccv.js", "cascade.js", "camshift.js", "whitebalance.js", "facetrackr.js", "smoother.js", "headposition.js"):
        it.run((REF / f).read_text())
    cases = []
    for name, params in (("default", {}), ("no_smoothing_fov", {"smoothing": False, "fov": 55.0})):
        p = jsmini.JSObject()
        p.props["ui"] = False
        for k, v in params.items():
            p.props[k] = v
        spec = stream_frames()
        video = jsmini.CanvasShim(make_frame(*spec[0]).copy())
        video.props.update(currentTime=1.0, paused=False, ended=False)
        canvas = jsmini.CanvasShim(np.zeros((H, W, 4), np.uint8))
        ht = it.get(["headtrackr", "Tracker"]).construct([p])
        it.events.clear()
        it.timers.clear()
        it.call(ht.get("init"), ht, video, canvas, False)
        steps = []
        for n, (kind, t) in enumerate(spec):
            video.pix = make_frame(kind, t).copy()
            it.now_ms += 35.0
            n0 = len(it.events)
            if n == 0:
                assert it.call(ht.get("start"), ht) is True
            else:
                live = [tm for tm in it.timers if not tm[3]]
                assert live, "no pending timer"
                tm = live[-1]
                tm[3] = True
                it.call(tm[1])
            ev = [event_record(e) for e in it.events[n0:]]
            steps.append(dict(frame=[kind, t], status=ht.get("status"), events=ev))
            print(name, n, kind, ht.get("status"), [(e.get("type"), e.get("status", e.get("detection", ""))) for e in ev], flush=True)
        cases.append(dict(name=name, params=params, steps=steps, fov=it.call(ht.get("getFOV"), ht)))
        # stop(): status event + no further timer
        n0 = len(it.events)
        it.call(ht.get("stop"), ht)
        cases[-1]["stop_events"] = [event_record(e) for e in it.events[n0:]]
    OUT.write_text(json.dumps(dict(generator="tools/make_goldens_main.py (src/main.js executed by oracle/jsmini.py)",
                                   width=W, height=H, cases=cases), indent=1))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
