"""headtrackr.camshift mirror — /root/reference/src/camshift.js.

    Tracker(params).initTracker(canvas, Rectangle) / track(canvas) / getTrackObj() / getSearchWindow()
    / getBackProjectionImg()                                     src/camshift.js:148-220
    Rectangle(x, y, w, h)                                        src/camshift.js:127-141
    TrackObj                                                     src/camshift.js:362-377
Each Tracker owns one tracker slot of a Context; all arithmetic runs in the CUDA library.
"""
from .canvas import as_pixels
from .runtime import default_context


class Rectangle:
    def __init__(self, x=0, y=0, w=0, h=0):
        self.x, self.y, self.width, self.height = x, y, w, h

    def clone(self):
        return Rectangle(self.x, self.y, self.width, self.height)


class TrackObj:
    def __init__(self):
        self.height = 0
        self.width = 0
        self.angle = 0
        self.x = 0
        self.y = 0

    def clone(self):
        c = TrackObj()
        c.height, c.width, c.angle, c.x, c.y = self.height, self.width, self.angle, self.x, self.y
        return c


def _take_slot(ctx):
    """Tracker slots are a per-context resource: a free list on the Context, no silent sharing when it runs dry."""
    free = ctx.__dict__.setdefault("_free_slots", list(range(ctx.max_frames - 1, -1, -1)))
    if not free:
        raise RuntimeError(f"all {ctx.max_frames} tracker slots of this context are in use "
                           "(create the Context with a larger max_frames, or drop Trackers you no longer need)")
    ctx.__dict__["_live_trackers"] = ctx.__dict__.get("_live_trackers", 0) + 1
    return free.pop()


def _give_slot(ctx, slot):
    ctx.__dict__.setdefault("_free_slots", []).append(slot)
    ctx.__dict__["_live_trackers"] = max(0, ctx.__dict__.get("_live_trackers", 1) - 1)


class Tracker:
    def __init__(self, params=None, context=None, slot=None):
        params = dict(params or {})
        self.calcAngles = params.get("calcAngles", True)          # src/camshift.js:151
        self._ctx = context
        self._slot = slot
        self._searchWindow = None
        self._trackObj = None
        self._last_frame = None

    def _context(self, px):
        if self._ctx is None:
            self._ctx = default_context(px.shape[1], px.shape[0])
        if self._slot is None:
            self._slot = _take_slot(self._ctx)
            self._own_slot = True
        return self._ctx

    def __del__(self):
        try:
            if getattr(self, "_own_slot", False) and self._ctx is not None:
                _give_slot(self._ctx, self._slot)
        except Exception:
            pass

    def initTracker(self, canvas, trackedArea):
        px = as_pixels(canvas)
        ctx = self._context(px)
        rect = [int(trackedArea.x), int(trackedArea.y), int(trackedArea.width), int(trackedArea.height)]
        ctx.track_init(px, [rect], slots=[self._slot], calc_angles=self.calcAngles)
        self._searchWindow = trackedArea.clone()
        self._trackObj = TrackObj()

    def track(self, canvas):
        px = as_pixels(canvas)
        if px.shape[0] == 0 or px.shape[1] == 0:                   # src/camshift.js:219
            return
        ctx = self._context(px)
        objs, wins = ctx.track(px, slots=[self._slot])
        o, w = objs[0], wins[0]
        t = TrackObj()
        t.x, t.y, t.width, t.height, t.angle = o["x"], o["y"], o["width"], o["height"], o["angle"]
        self._trackObj = t
        self._searchWindow = Rectangle(*w)
        self._last_frame = px

    def getTrackObj(self):
        return self._trackObj.clone()

    def getSearchWindow(self):
        return self._searchWindow.clone()

    def getBackProjectionImg(self):
        return self._ctx.backprojection(self._last_frame, self._slot)
