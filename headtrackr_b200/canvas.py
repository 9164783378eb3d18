"""Minimal stand-in for the HTML canvas the reference passes around (width, height, RGBA pixels).

The reference's L1 functions take a canvas and read it with getImageData (src/ccv.js:25,
src/camshift.js:206,218, src/whitebalance.js:12).  Here a canvas is an (H, W, 4) uint8 array with
alpha 255 — host numpy memory, or a torch CUDA tensor for zero-copy device frames.
"""
import numpy as np


class Canvas:
    def __init__(self, pixels):
        if hasattr(pixels, "is_cuda"):
            assert pixels.dim() == 3 and pixels.shape[2] == 4
        else:
            pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
            assert pixels.ndim == 3 and pixels.shape[2] == 4
        self.pixels = pixels

    @property
    def width(self):
        return int(self.pixels.shape[1])

    @property
    def height(self):
        return int(self.pixels.shape[0])


def as_pixels(canvas):
    return canvas.pixels if isinstance(canvas, Canvas) else canvas
