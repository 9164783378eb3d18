"""CPU: k_gray's and k_resample's per-thread code (gray_item, resample_thread in ht_detect.cuh) executed thread by
thread on the host, then k_cascade's tile evaluation on the result (test_cascade_host.py): the whole detector from
RGBA frames to raw detection lists, with the kernels' own code and the planner's own tables, against the oracle.

Every plane of the frame-quad-interleaved arena must equal the oracle's pyramid (src/ccv.js:110-147 over the defined
canvas shim) in each of its four byte lanes, pad columns and unpainted rows/columns must be 0, and the raw lists must
equal src/ccv.js:178-243 bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from headtrackr_b200 import synth
from test_cascade_host import st, want_raw  # noqa: F401  (fixture + helper)


def host_arena(st, frames, W, H, interval):
    info = np.zeros(2 + 6 * 256, np.int32)
    assert st.ht_selftest_planes(W, H, interval, info.ctypes.data, info.size) == 0
    stride = int(info[1])
    arena = np.full(stride, 0xDEADBEEF, np.uint32)          # the kernels must write every word they own
    rgba = np.ascontiguousarray(np.stack(frames))
    st.ht_selftest_pyramid.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    assert st.ht_selftest_pyramid(W, H, interval, rgba.ctypes.data, len(frames), arena.ctypes.data, arena.size) == 0
    return arena, info


@pytest.mark.parametrize("W,H,interval,n", [(160, 120, 5, 4), (171, 133, 3, 3), (320, 240, 5, 2)])
def test_host_run_of_gray_and_resample_equals_the_oracle(st, blob, W, H, interval, n):
    frames = [synth.frame(20 + i, W, H) for i in range(n)]
    arena, info = host_arena(st, frames, W, H, interval)
    pyrs = [oracle.Pyramid(oracle.grayscale(f), interval) for f in frames]
    for i in range(int(info[0])):
        off, pitch, w, h, slot, q = (int(v) for v in info[2 + 6 * i: 8 + 6 * i])
        words = arena[off: off + pitch * h].reshape(h, pitch)
        assert not (words[:, w:] != 0).any(), (slot, q, "pad columns")
        for f in range(4):
            lane = ((words[:, :w] >> (8 * f)) & 0xFF).astype(np.uint8)
            if f < n:
                assert np.array_equal(lane, pyrs[f].plane(slot, q)), (slot, q, f)
            else:
                assert not lane.any(), (slot, q, f, "missing frames are 0")
    # ... and the cascade on exactly this arena
    out = np.zeros((4, 8192, 4), np.float64)
    counts = np.zeros(4, np.int32)
    assert st.ht_selftest_cascade(blob, len(blob), W, H, interval, arena.ctypes.data, n, 0, 2, out.ctypes.data,
                                  counts.ctypes.data, 8192) == 0
    for f in range(n):
        assert [tuple(out[f, i]) for i in range(counts[f])] == want_raw(frames[f], blob, interval)
