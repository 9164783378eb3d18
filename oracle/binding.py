"""ctypes binding of oracle/libht_oracle.so (the plain-C restatement of src/ccv.js + src/camshift.js)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libht_oracle.so"


def build(force=False):
    src = [_HERE / "ht_oracle.c", _HERE / "ht_oracle.h"]
    if force or not _SO.exists() or any(s.stat().st_mtime > _SO.stat().st_mtime for s in src):
        subprocess.check_call(["make", "-s", "-C", str(_HERE)] + (["-B"] if force else []))
    return _SO


class Rect(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("width", C.c_double), ("height", C.c_double),
                ("confidence", C.c_double), ("neighbors", C.c_int32), ("pad_", C.c_int32)]

    def astuple(self):
        return (self.x, self.y, self.width, self.height, self.confidence, self.neighbors)


class Geom(C.Structure):
    _fields_ = [("interval", C.c_int), ("next", C.c_int), ("scale_upto", C.c_int), ("n_slots", C.c_int),
                ("w", C.c_int * 128), ("h", C.c_int * 128)]


class DetectStats(C.Structure):
    _fields_ = [("windows", C.c_int64), ("feature_evals", C.c_int64), ("stage_entries", C.c_int64 * 64),
                ("n_raw", C.c_int)]


class Tracker(C.Structure):
    _fields_ = [("model_hist", C.c_uint32 * 4096),
                ("sx", C.c_int32), ("sy", C.c_int32), ("sw", C.c_int32), ("sh", C.c_int32),
                ("tx", C.c_int32), ("ty", C.c_int32), ("tw", C.c_int32), ("th", C.c_int32),
                ("angle", C.c_double), ("calc_angles", C.c_int32), ("initialised", C.c_int32)]


class TrackTrace(C.Structure):
    _fields_ = [("n_iter", C.c_int32), ("converged", C.c_int32),
                ("wx", C.c_int32 * 11), ("wy", C.c_int32 * 11),
                ("m00", C.c_double), ("m10", C.c_double), ("m01", C.c_double),
                ("m11", C.c_double), ("m20", C.c_double), ("m02", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        u8p = C.POINTER(C.c_uint8)
        L.hto_geometry.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(Geom)]
        L.hto_grayscale.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.hto_grayscale.restype = None
        L.hto_draw_image.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
        L.hto_draw_image.restype = None
        L.hto_pyramid_build.argtypes = [u8p, C.c_int, C.c_int, C.c_int]
        L.hto_pyramid_build.restype = C.c_void_p
        L.hto_pyramid_free.argtypes = [C.c_void_p]
        L.hto_pyramid_free.restype = None
        L.hto_pyramid_geom.argtypes = [C.c_void_p]
        L.hto_pyramid_geom.restype = C.POINTER(Geom)
        L.hto_pyramid_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hto_pyramid_plane.restype = u8p
        L.hto_detect.argtypes = [u8p, C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.c_int, C.c_int,
                                 C.POINTER(Rect), C.c_int, C.POINTER(Rect), C.c_int, C.POINTER(DetectStats)]
        L.hto_cascade_raw.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(Rect), C.c_int,
                                      C.POINTER(DetectStats)]
        L.hto_group.argtypes = [C.POINTER(Rect), C.c_int, C.c_int, C.POINTER(Rect), C.c_int]
        L.hto_histogram.argtypes = [u8p, C.c_size_t, C.POINTER(C.c_uint32)]
        L.hto_histogram.restype = None
        L.hto_weights.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
        L.hto_weights.restype = None
        L.hto_tracker_init.argtypes = [C.POINTER(Tracker), u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int]
        L.hto_tracker_track.argtypes = [C.POINTER(Tracker), u8p, C.c_int, C.c_int, C.POINTER(TrackTrace)]
        L.hto_backprojection_img.argtypes = [C.POINTER(Tracker), u8p, C.c_int, C.c_int, u8p]
        L.hto_backprojection_img.restype = None
        L.hto_detect_track.argtypes = [u8p, C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.POINTER(Tracker), C.POINTER(C.c_int)]
        L.hto_whitebalance.argtypes = [u8p, C.c_int, C.c_int]
        L.hto_whitebalance.restype = C.c_double
        _lib = L
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def geometry(W, H, interval=5):
    g = Geom()
    rc = lib().hto_geometry(W, H, interval, C.byref(g))
    if rc != 0:
        raise ValueError(f"hto_geometry({W},{H},{interval}) -> {rc}")
    return g


def grayscale(rgba):
    rgba, p = _u8(rgba)
    H, W = rgba.shape[:2]
    out = np.empty((H, W), np.uint8)
    lib().hto_grayscale(p, W, H, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def draw_image(src, sx, sy, sw, sh, dst_w, dst_h, dw, dh):
    src, p = _u8(src)
    dst = np.zeros((dst_h, dst_w), np.uint8)
    lib().hto_draw_image(p, src.shape[1], sx, sy, sw, sh, dst.ctypes.data_as(C.POINTER(C.c_uint8)), dst_w, dw, dh)
    return dst


class Pyramid:
    def __init__(self, gray, interval=5):
        gray, p = _u8(gray)
        H, W = gray.shape
        self._h = lib().hto_pyramid_build(p, W, H, interval)
        if not self._h:
            raise ValueError("pyramid build failed (frame too small?)")
        self.geom = lib().hto_pyramid_geom(self._h).contents

    def plane(self, slot, q=0):
        w, h = C.c_int(), C.c_int()
        p = lib().hto_pyramid_plane(self._h, slot, q, C.byref(w), C.byref(h))
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(h.value, w.value)).copy()

    def cascade_raw(self, blob, cap=65536):
        raw = (Rect * cap)()
        st = DetectStats()
        n = lib().hto_cascade_raw(self._h, blob, len(blob), raw, cap, C.byref(st))
        if n < 0:
            raise RuntimeError(f"hto_cascade_raw -> {n}")
        return [raw[i].astuple() for i in range(min(n, cap))], st

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hto_pyramid_free(self._h)
            self._h = None


def rects_from_tuples(t):
    arr = (Rect * max(len(t), 1))()
    for i, r in enumerate(t):
        arr[i].x, arr[i].y, arr[i].width, arr[i].height, arr[i].confidence, arr[i].neighbors = r
    return arr


def group(raw, min_neighbors=1, cap=4096):
    arr = rects_from_tuples(raw)
    out = (Rect * cap)()
    n = lib().hto_group(arr, len(raw), min_neighbors, out, cap)
    return [out[i].astuple() for i in range(min(n, cap))]


def detect(rgba, blob, interval=5, min_neighbors=1, cap=4096, raw_cap=65536, want_raw=False, want_stats=False):
    """ccv.detect_objects(ccv.grayscale(canvas), cascade, interval, min_neighbors) on an (H,W,4) u8 frame."""
    rgba, p = _u8(rgba)
    H, W = rgba.shape[:2]
    out = (Rect * cap)()
    raw = (Rect * raw_cap)()
    st = DetectStats()
    n = lib().hto_detect(p, W, H, blob, len(blob), interval, min_neighbors, out, cap, raw, raw_cap, C.byref(st))
    if n < 0:
        raise RuntimeError(f"hto_detect -> {n}")
    res = [out[i].astuple() for i in range(min(n, cap))]
    ret = [res]
    if want_raw:
        ret.append([raw[i].astuple() for i in range(min(st.n_raw, raw_cap))])
    if want_stats:
        ret.append(st)
    return ret[0] if len(ret) == 1 else tuple(ret)


def histogram(rgba):
    rgba, p = _u8(rgba)
    bins = np.zeros(4096, np.uint32)
    lib().hto_histogram(p, rgba.size // 4, bins.ctypes.data_as(C.POINTER(C.c_uint32)))
    return bins


def weights(model, cur):
    model = np.ascontiguousarray(model, np.uint32)
    cur = np.ascontiguousarray(cur, np.uint32)
    w = np.zeros(4096, np.float64)
    lib().hto_weights(model.ctypes.data_as(C.POINTER(C.c_uint32)), cur.ctypes.data_as(C.POINTER(C.c_uint32)),
                      w.ctypes.data_as(C.POINTER(C.c_double)))
    return w


class CamshiftTracker:
    """camshift.Tracker restated (src/camshift.js:148-354)."""

    def __init__(self, calc_angles=True):
        self.t = Tracker()
        self.calc_angles = bool(calc_angles)

    def init_tracker(self, rgba, x, y, w, h):
        rgba, p = _u8(rgba)
        H, W = rgba.shape[:2]
        rc = lib().hto_tracker_init(C.byref(self.t), p, W, H, x, y, w, h, int(self.calc_angles))
        if rc != 0:
            raise ValueError("initTracker: empty rectangle")

    def track(self, rgba):
        rgba, p = _u8(rgba)
        H, W = rgba.shape[:2]
        tr = TrackTrace()
        rc = lib().hto_tracker_track(C.byref(self.t), p, W, H, C.byref(tr))
        if rc != 0:
            raise RuntimeError("track before initTracker")
        return tr

    def track_obj(self):
        t = self.t
        return dict(x=t.tx, y=t.ty, width=t.tw, height=t.th, angle=t.angle)

    def search_window(self):
        t = self.t
        return (t.sx, t.sy, t.sw, t.sh)

    def backprojection_img(self, rgba):
        rgba, p = _u8(rgba)
        H, W = rgba.shape[:2]
        out = np.zeros((H, W, 4), np.uint8)
        lib().hto_backprojection_img(C.byref(self.t), p, W, H, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out


def detect_track(rgba, blob, interval=5, min_neighbors=1, calc_angles=False, n_calls=30):
    """One C call: detect, VJ->CS hand-off, n_calls x track().  -> (n_detections, found, track_obj dict)."""
    rgba, p = _u8(rgba)
    H, W = rgba.shape[:2]
    t = Tracker()
    found = C.c_int()
    n = lib().hto_detect_track(p, W, H, blob, len(blob), interval, min_neighbors, int(calc_angles), n_calls,
                               C.byref(t), C.byref(found))
    return n, found.value, dict(x=t.tx, y=t.ty, width=t.tw, height=t.th, angle=t.angle)


def whitebalance(rgba):
    rgba, p = _u8(rgba)
    H, W = rgba.shape[:2]
    return lib().hto_whitebalance(p, W, H)
