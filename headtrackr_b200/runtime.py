"""Process-wide default Context (one GPU context per (device, cascade), sized on demand)."""
from .context import Context

_ctx = {}


def default_context(width, height, cascade=None, max_frames=16, device=0):
    key = (device, None if cascade is None else hash(cascade))
    c = _ctx.get(key)
    if c is None or c.max_w < width or c.max_h < height:
        # a context that still has live camshift.Trackers is NOT closed (they hold its slots and state): the larger
        # one simply replaces it as the default, the old one lives as long as its trackers do
        if c is not None and not c.__dict__.get("_live_trackers", 0):
            c.close()
        c = Context(max_width=max(width, 640), max_height=max(height, 480), max_frames=max_frames, device=device,
                    cascade=cascade)
        c.max_w, c.max_h = max(width, 640), max(height, 480)
        _ctx[key] = c
    return c
