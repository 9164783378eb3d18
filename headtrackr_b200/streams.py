"""Many independent video streams through facetrackr's state machine, batched on the GPU (ht_stream_step).

What `headtrackr.facetrackr.Tracker.track()` does for ONE stream per call in the reference
(/root/reference/src/facetrackr.js:67-126, with the lost-face re-detection of src/main.js:230-244) this class does
for n streams per call: stream k owns tracker slot k of the context; the mode switches ("VJ" -> "CS" on a face,
"CS" -> "VJ" on a lost one), the max-confidence pick and the tracker seeding run in kernels, and the host only
drains one event record per stream and frame.  `facetrackingEvent`s are dispatched exactly as the reference does:
for records with detection == "CS" (src/facetrackr.js:112-125).
"""
import time


class StreamSet:
    def __init__(self, context, n_streams, interval=5, min_neighbors=1, calc_angles=False):
        if n_streams > context.max_frames:
            raise ValueError("more streams than tracker slots in the context")
        self.ctx, self.n = context, n_streams
        self.interval, self.min_neighbors, self.calc_angles = interval, min_neighbors, calc_angles
        self._listeners = []
        self.current = [None] * n_streams            # getTrackingObject() per stream
        context.stream_reset(0, n_streams)

    def addEventListener(self, fn):
        """fn(stream_index, evt): evt is the facetrackingEvent dict (src/facetrackr.js:112-125)."""
        self._listeners.append(fn)

    def reset(self, stream):
        """A new facetrackr.Tracker({whitebalancing: false}) for one stream (src/main.js:236)."""
        self.ctx.stream_reset(stream, 1)

    def enable_head_position(self, **params):
        """headtrackr.Tracker's smoothing + head position per stream (src/main.js:246-300) as a GPU epilogue: the
        listeners then also receive `headtrackingEvent {x, y, z}` and `headtrackrStatus {status: "found"}` dicts."""
        self.ctx.stream_head_config(**params)
        self._head = True

    def track(self, frames):
        """frames: (n, H, W, 4) u8 (numpy or torch CUDA) - the current frame of every stream."""
        t0 = time.time()
        heads = None
        if getattr(self, "_head", False):
            events, heads = self.ctx.stream_step_head(frames, self.interval, self.min_neighbors, self.calc_angles)
        else:
            events = self.ctx.stream_step(frames, self.interval, self.min_neighbors, self.calc_angles)
        dt = int((time.time() - t0) * 1000)
        for k, e in enumerate(events):
            e["time"] = dt
            self.current[k] = e
            if e["detection"] == "CS":
                evt = dict(type="facetrackingEvent", height=e["height"], width=e["width"], angle=e["angle"], x=e["x"],
                           y=e["y"], confidence=e["confidence"], detection="CS", time=dt)
                for fn in self._listeners:
                    fn(k, evt)
            if heads is not None:
                if heads[k]["found"]:
                    for fn in self._listeners:
                        fn(k, dict(type="headtrackrStatus", status="found"))
                if heads[k]["valid"]:
                    for fn in self._listeners:
                        fn(k, dict(type="headtrackingEvent", x=heads[k]["x"], y=heads[k]["y"], z=heads[k]["z"]))
        return events

    def getTrackingObject(self, stream):
        return dict(self.current[stream]) if self.current[stream] is not None else None
