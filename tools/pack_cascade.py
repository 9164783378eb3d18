#!/usr/bin/env python
"""Pack the reference's BBF cascade model (`/root/reference/src/cascade.js:19`) into a binary blob.

The cascade is MODEL DATA (learned weights), the one reference artefact that is consumed
verbatim.  This tool parses the JSON literal after
``headtrackr.cascade = `` and writes ``headtrackr_b200/data/cascade_face.bin``.

Blob layout (little endian), "HTC1":
    0   char[4]  magic "HTC1"
    4   u32      n_stages
    8   u32      n_features (total)
    12  u32      window width  (24)
    16  u32      window height (24)
    20  u32      reserved (0)
    24  stages   n_stages x { u32 count; u32 first_feature; f64 threshold }       (16 B each)
    ..  features n_features x { u8 size; u8 pad; i8 pz[5]; u8 px[5]; u8 py[5];
                                i8 nz[5]; u8 nx[5]; u8 ny[5] }                      (32 B each)
    ..  alpha    n_features x { f64 alpha_fail (alpha[2k]); f64 alpha_pass (alpha[2k+1]) }
Unused point slots (index >= size, or z == -1) are stored as z = -1, x = y = 0.
Decimal literals are converted with Python's float() == JS Number() (correctly rounded).
"""
import json
import struct
import sys
from pathlib import Path

REF = Path("/root/reference/src/cascade.js")
OUT = Path(__file__).resolve().parent.parent / "headtrackr_b200" / "data" / "cascade_face.bin"

EXPECTED_COUNTS = [4, 4, 7, 13, 20, 22, 32, 45, 61, 80, 115, 153, 203, 301, 391, 564]


def parse_cascade_js(path=REF):
    text = Path(path).read_text()
    key = "headtrackr.cascade = "
    body = text[text.index(key) + len(key):].strip()
    if body.endswith(";"):
        body = body[:-1]
    return json.loads(body)


def pack(c):
    stages = c["stage_classifier"]
    assert c["count"] == len(stages)
    n_feat = sum(st["count"] for st in stages)
    out = bytearray()
    out += b"HTC1" + struct.pack("<5I", len(stages), n_feat, c["width"], c["height"], 0)
    first = 0
    for st in stages:
        assert len(st["feature"]) == st["count"] and len(st["alpha"]) == 2 * st["count"]
        out += struct.pack("<IId", st["count"], first, float(st["threshold"]))
        first += st["count"]
    for st in stages:
        for f in st["feature"]:
            size = f["size"]
            assert 1 <= size <= 5
            rec = bytearray(32)
            rec[0] = size
            for side, base in (("p", 2), ("n", 17)):
                zs, xs, ys = f[side + "z"], f[side + "x"], f[side + "y"]
                assert len(zs) == len(xs) == len(ys) == size
                assert zs[0] >= 0, "slot 0 is read unconditionally (src/ccv.js:191-192)"
                for q in range(5):
                    z = zs[q] if q < size else -1
                    x = xs[q] if (q < size and z >= 0) else 0
                    y = ys[q] if (q < size and z >= 0) else 0
                    assert -1 <= z <= 2
                    if z >= 0:
                        lim = (c["width"] >> z) - 1
                        assert 0 <= x <= lim and 0 <= y <= lim, (x, y, z)
                    rec[base + q] = z & 0xFF
                    rec[base + 5 + q] = x
                    rec[base + 10 + q] = y
            out += rec
    for st in stages:
        a = st["alpha"]
        for k in range(st["count"]):
            out += struct.pack("<dd", float(a[2 * k]), float(a[2 * k + 1]))
    return bytes(out)


def main():
    c = parse_cascade_js(sys.argv[1] if len(sys.argv) > 1 else REF)
    counts = [st["count"] for st in c["stage_classifier"]]
    assert counts == EXPECTED_COUNTS, counts
    for st in c["stage_classifier"]:
        a = st["alpha"]
        for k in range(st["count"]):
            assert a[2 * k] == -a[2 * k + 1] and a[2 * k] < 0
    blob = pack(c)
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_bytes(blob)
    print(f"wrote {OUT} ({len(blob)} bytes, {len(counts)} stages, {sum(counts)} features)")


if __name__ == "__main__":
    main()
