"""ctypes binding of libheadtrackr_b200.so (C ABI: include/headtrackr_b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc (sm_100a); if that
is impossible the import fails loudly.  Nothing here imports the oracle.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# HT_LIB selects another build of the same library (A/B variants of compile-time knobs, tools/build_variants.sh)
SO_PATH = Path(os.environ["HT_LIB"]).resolve() if os.environ.get("HT_LIB") else _PKG / "libheadtrackr_b200.so"
CSRC = _PKG / "csrc"

HT_OK, HT_WARN_OVERFLOW = 0, 1
HT_ERR_ARG, HT_ERR_CUDA, HT_ERR_SIZE, HT_ERR_CASCADE, HT_ERR_STATE = -1, -2, -3, -4, -5


class HtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"headtrackr_b200 error {code}: {msg}")
        self.code = code


class Rect(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("width", C.c_double), ("height", C.c_double),
                ("confidence", C.c_double), ("neighbors", C.c_int32), ("pad_", C.c_int32)]


class TrackObj(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("angle", C.c_double)]


class Window(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("width", C.c_int32), ("height", C.c_int32)]


class StreamEvent(C.Structure):
    _fields_ = [("detection", C.c_int32), ("status", C.c_int32), ("x", C.c_double), ("y", C.c_double),
                ("width", C.c_double), ("height", C.c_double), ("angle", C.c_double), ("confidence", C.c_double)]


class HeadParams(C.Structure):
    _fields_ = [("smoothing", C.c_int32), ("head_position", C.c_int32), ("edgecorrection", C.c_int32), ("pad_", C.c_int32),
                ("alpha", C.c_double), ("fov_deg", C.c_double), ("camera_offset", C.c_double),
                ("distance_to_screen", C.c_double)]


class HeadEvent(C.Structure):
    _fields_ = [("valid", C.c_int32), ("status", C.c_int32), ("x", C.c_double), ("y", C.c_double), ("z", C.c_double),
                ("fx", C.c_double), ("fy", C.c_double), ("fwidth", C.c_double), ("fheight", C.c_double)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_width", C.c_int32), ("max_height", C.c_int32),
                ("max_frames", C.c_int32), ("max_raw_per_frame", C.c_int32), ("max_rects_per_frame", C.c_int32),
                ("cuda_stream", C.c_void_p)]


def sources_newer_than_so():
    if not SO_PATH.exists():
        return True
    t = SO_PATH.stat().st_mtime
    srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inc")) + \
        [_PKG.parent / "include" / "headtrackr_b200.h"]
    return any(s.stat().st_mtime > t for s in srcs)


def build(force=False):
    """Compile the sm_100a shared library in-tree (nvcc cross-compiles without a GPU)."""
    if force or sources_newer_than_so():
        subprocess.check_call(["make", "-s", "-C", str(CSRC)] + (["-B"] if force else []))
    if not SO_PATH.exists():
        raise ImportError(f"{SO_PATH} was not produced by the build")
    return SO_PATH


_lib = None

EXPORTS = ["ht_version", "ht_create", "ht_destroy", "ht_last_error", "ht_sync", "ht_max_rects", "ht_detect",
           "ht_track_init", "ht_track_init_from_detect", "ht_track", "ht_detect_track", "ht_stream_reset", "ht_stream_step", "ht_stream_head_config", "ht_stream_step_head", "ht_ingest", "ht_backprojection", "ht_whitebalance",
           "ht_plan_info", "ht_debug_plane", "ht_debug_raw", "ht_debug_model_hist", "ht_debug_track_stats", "ht_set_track_memo", "ht_set_pipeline", "ht_join", "ht_debug_set_exactness", "ht_debug_track_trace", "ht_debug_track_phases", "ht_launch_count",
           "ht_profile", "ht_profile_read"]

PROF_CLASSES = ["gray", "pyramid", "cascade", "group", "hist", "track_init", "track"]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not SO_PATH.exists():
        try:
            build()
        except Exception as e:  # no silent fallback
            raise ImportError(f"libheadtrackr_b200.so is missing and could not be built: {e}") from e
    L = C.CDLL(str(SO_PATH))
    vp = C.c_void_p  # raw addresses: host or device pointers
    L.ht_version.restype = C.c_uint32
    L.ht_create.argtypes = [C.POINTER(vp), C.POINTER(Config), C.c_char_p, C.c_size_t]
    L.ht_destroy.argtypes = [vp]
    L.ht_destroy.restype = None
    L.ht_last_error.argtypes = [vp]
    L.ht_last_error.restype = C.c_char_p
    L.ht_sync.argtypes = [vp]
    L.ht_max_rects.argtypes = [vp]
    L.ht_detect.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.ht_track_init.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int]
    L.ht_track_init_from_detect.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]
    L.ht_track.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.ht_detect_track.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  vp, vp, vp, vp, vp]
    L.ht_stream_reset.argtypes = [vp, C.c_int, C.c_int]
    L.ht_stream_step.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    L.ht_stream_head_config.argtypes = [vp, vp]
    L.ht_stream_step_head.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.ht_ingest.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int]
    L.ht_backprojection.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp]
    L.ht_whitebalance.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    L.ht_plan_info.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int]
    L.ht_debug_plane.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp]
    L.ht_debug_raw.argtypes = [vp, C.c_int, vp, C.c_int, vp]
    L.ht_debug_model_hist.argtypes = [vp, C.c_int, vp]
    L.ht_debug_track_stats.argtypes = [vp, vp, C.c_int]
    L.ht_debug_track_trace.argtypes = [vp, vp, C.c_int]
    L.ht_debug_track_phases.argtypes = [vp, vp, C.c_int]
    L.ht_set_track_memo.argtypes = [vp, C.c_int]
    L.ht_set_pipeline.argtypes = [vp, C.c_int]
    L.ht_join.argtypes = [vp]
    L.ht_debug_set_exactness.argtypes = [vp, C.c_int]
    L.ht_launch_count.argtypes = [vp]
    L.ht_launch_count.restype = C.c_uint64
    L.ht_profile.argtypes = [vp, C.c_int]
    L.ht_profile_read.argtypes = [vp, vp, vp, C.c_int]
    _lib = L
    return L
