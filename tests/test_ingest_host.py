"""CPU: frame ingest (ht_ingest, src/main.js:170 drawImage(video -> canvas)) - k_ingest's per-pixel code run on the
host against the oracle's canvas-shim drawImage, channel by channel, for down- and up-scaling and odd sizes."""
import ctypes as C

import numpy as np
import pytest

import oracle
from headtrackr_b200 import synth
from test_cascade_host import st  # noqa: F401  (fixture: the host-only build of ht_api.cu)


def oracle_resize(frame, dw, dh):
    sh, sw = frame.shape[:2]
    out = np.zeros((dh, dw, 4), np.uint8)
    for c in range(4):
        out[..., c] = oracle.draw_image(np.ascontiguousarray(frame[..., c]), 0, 0, sw, sh, dw, dh, dw, dh)
    return out


@pytest.mark.parametrize("sw,sh,dw,dh", [(640, 480, 320, 240), (333, 251, 320, 240), (160, 120, 320, 240), (320, 240, 171, 97)])
def test_ingest_pixels_equal_the_oracle(st, sw, sh, dw, dh):
    frames = np.stack([synth.frame(70 + i, sw, sh) for i in range(2)])
    frames[1, ..., 3] = np.arange(sw, dtype=np.uint8)[None, :]          # a non-constant alpha channel too
    out = np.zeros((2, dh, dw, 4), np.uint8)
    st.ht_selftest_ingest.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    assert st.ht_selftest_ingest(frames.ctypes.data, 2, sw, sh, out.ctypes.data, dw, dh) == 0
    for i in range(2):
        assert np.array_equal(out[i], oracle_resize(frames[i], dw, dh)), i
