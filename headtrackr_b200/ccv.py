"""headtrackr.ccv mirror — /root/reference/src/ccv.js.

    detect_objects(canvas, cascade, interval, min_neighbors)   src/ccv.js:109
    grayscale(canvas)                                           src/ccv.js:22

The reference calls detect_objects(grayscale(copy), cascade, 5, 1) (src/facetrackr.js:147-149); the
CUDA path fuses the grayscale pass into ht_detect, so grayscale() here only tags the canvas.
"""
from .canvas import Canvas, as_pixels
from .runtime import default_context


class _GrayTagged(Canvas):
    """Result of grayscale(): the colour pixels plus the promise that detect_objects grays them."""


def grayscale(canvas):
    return _GrayTagged(as_pixels(canvas))


def detect_objects(canvas, cascade=None, interval=5, min_neighbors=1, context=None):
    """-> list of {x, y, width, height, neighbors, confidence} exactly as src/ccv.js:293-330
    (or the raw {.., neighbor: 1, ..} list when min_neighbors <= 0, src/ccv.js:249-250)."""
    if not isinstance(canvas, _GrayTagged):
        # The reference's detect_objects does no graying: it reads channel 0 of whatever it is given
        # (src/ccv.js:171-192).  ht_detect always applies ccv.grayscale first, so only the call the reference
        # actually makes - detect_objects(grayscale(canvas), ...) - is served; anything else would silently differ.
        raise TypeError("detect_objects() expects the result of ccv.grayscale(canvas) (src/facetrackr.js:147-149)")
    px = as_pixels(canvas)
    ctx = context or default_context(px.shape[1], px.shape[0], cascade)
    return ctx.detect(px, interval, min_neighbors)[0]
