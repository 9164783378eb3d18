#!/usr/bin/env python
"""Aggregate an ncu `--page source --print-source cuda,sass --csv` dump of k_cascade by kernel region (source-line ranges of
ht_detect.cuh / the generated include).  Usage: python tools/ncu_regions.py dump.csv [--lines]"""
import csv
import sys


def num(x):
    try:
        return float(x)
    except Exception:
        return 0.0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    cur = None
    hdr = None
    out = []
    for r in rows:
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            cur = r[1].split('/')[-1]
            continue
        if r and r[0] == "Line No" and len(r) > 5:
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            d = {}
            for k, v in zip(hdr, r):
                if k not in d:
                    d[k] = v
            out.append((cur or "?", int(r[0]), r[1], d))
    tot_s = sum(num(d['# Samples']) for _, _, _, d in out)
    tot_i = sum(num(d['Instructions Executed']) for _, _, _, d in out)
    tot_w = sum(num(d['L1 Wavefronts Shared']) for _, _, _, d in out)
    print("total samples", tot_s, "inst", tot_i, "shared wavefronts", tot_w)
    src = open(__file__.rsplit('/tools/', 1)[0] + '/headtrackr_b200/csrc/ht_detect.cuh').read().split('\n')
    # region markers: first line containing the marker starts the region
    marks = [("prologue", "k_cascade(DevPlan plan"), ("tma issue", "const bool use_tma = tmaps"), ("stage L0", "level 0, columns split by parity"),
             ("stage L1 cp.async", "if (!use_tma) {  // level 1"), ("stage L2", "level 2: the four phase copies"),
             ("tma wait + sync", "if (use_tma) {   // every thread observes"), ("lambdas", "const uint8_t *tile_b ="),
             ("dense group", "---- dense group"), ("survivor groups", "---- survivor masks"), ("no-late tail", "if (!has_late) {"),
             ("late stages", "---- late stages"), ("after", "K4  sort raw detections")]
    bounds = []
    for name, m in marks:
        for i, l in enumerate(src):
            if m in l:
                bounds.append((i + 1, name))
                break
    bounds.sort()
    agg = {}
    for f, l, s, d in out:
        key = f
        if f == "ht_detect.cuh":
            key = "ht_detect.cuh: helpers (feat_fires, ordered sums, nth_set_bit, ...)"
            for b, name in bounds:
                if l >= b:
                    key = name
            if key == "after":
                key = "ht_detect.cuh: after"
            if l < bounds[0][0]:
                key = "ht_detect.cuh: helpers (feat_fires, ordered sums, nth_set_bit, ...)"
        a = agg.setdefault(key, [0, 0, 0, 0, 0])
        a[0] += num(d['# Samples'])
        a[1] += num(d['Instructions Executed'])
        a[2] += num(d['L1 Wavefronts Shared'])
        a[3] += num(d.get('stall_barrier', 0))
        a[4] += num(d.get('stall_long_sb', 0))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{k:72s} samples {100*a[0]/tot_s:5.1f}%  inst {100*a[1]/tot_i:5.1f}%  wavefr {100*a[2]/max(tot_w,1):5.1f}%  "
              f"barrier {100*a[3]/tot_s:5.1f}%  long_sb {100*a[4]/tot_s:5.1f}%")
    if "--lines" in sys.argv:
        print("\ntop lines:")
        for f, l, s, d in sorted(out, key=lambda x: -num(x[3]['# Samples']))[:40]:
            print(f"{f}:{l:4d} {100*num(d['# Samples'])/tot_s:5.2f}% inst {100*num(d['Instructions Executed'])/tot_i:5.2f}%  {s.strip()[:110]}")


if __name__ == "__main__":
    main()
