"""CPU: the C-ABI shared library loads, exports every symbol include/headtrackr_b200.h declares, and
fails loudly without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import pytest

from headtrackr_b200 import _lib

HEADER = Path(__file__).resolve().parent.parent / "include" / "headtrackr_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ht_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(names) == sorted(_lib.EXPORTS)                  # the Python binding covers the whole ABI


def test_struct_layouts_match_the_header():
    assert C.sizeof(_lib.Rect) == 48 and C.sizeof(_lib.TrackObj) == 24 and C.sizeof(_lib.Window) == 16
    assert _lib.Rect.confidence.offset == 32 and _lib.Rect.neighbors.offset == 40
    assert _lib.TrackObj.angle.offset == 16


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu suite")
    from headtrackr_b200 import Context
    from headtrackr_b200._lib import HtError, HT_ERR_CUDA
    with pytest.raises(HtError) as e:
        Context()
    assert e.value.code == HT_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_bad_cascade_is_rejected_before_touching_the_gpu():
    L = _lib.lib()
    h = C.c_void_p()
    cfg = _lib.Config(0, 640, 480, 4, 0, 0, None)
    rc = L.ht_create(C.byref(h), C.byref(cfg), b"nope", 4)
    assert rc in (_lib.HT_ERR_CASCADE, _lib.HT_ERR_CUDA) and not h.value
    assert L.ht_version() >> 16 == 1


def test_product_does_not_import_the_oracle():
    pkg = Path(_lib.__file__).parent
    for p in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if p.suffix in (".py", ".cu", ".cuh", ".inc"):
            text = p.read_text()
            assert "import oracle" not in text and "from oracle" not in text, p      # no Python import
            assert "libht_oracle" not in text and '#include "ht_oracle' not in text, p  # no link / include
