#!/usr/bin/env python
"""Would early exits inside a stage pay?  (CPU model, like tools/cascade_wavefront_model.py)

k_cascade's byte stages 2-7 decide `U > T` on an exact integer U that only grows (a fired feature adds 2*alpha_int), so a
window is DECIDED as soon as U > T (pass) or U + (everything the remaining features could add) < T (fail).  A warp
iteration could stop loading at the first checkpoint at which all of its live lanes are decided.  This script replays
the kernel's distribution of survivors over warp iterations (quad of 4 frames, 32x8 tiles, bank classes, rank slices per
warp, groups {2} {3} {4,5} {6,7}) and counts the byte loads with and without such exits.

    python tools/early_exit_model.py [first_frame] [checkpoint_every]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import oracle  # noqa: E402  (analysis tool, not the product)
from headtrackr_b200 import synth  # noqa: E402
from cascade_wavefront_model import stage_tables  # noqa: E402

TW, TH, NWARP = 32, 8, 8
GROUPS = [[2], [3], [4, 5], [6, 7]]


def evaluate(frame, blob, interval=5):
    """per scale: (qw, qh, kstar[stage 0..7][4, qh, qw]) kstar = -1: window not alive at that stage; else the index of the
    feature after which the window is decided; passed[stage] bool"""
    c = synth.parse_blob(blob)
    st = stage_tables(c)
    pyr = oracle.Pyramid(oracle.grayscale(frame), interval)
    g = pyr.geom
    nxt = g.next
    out = []
    for i in range(g.scale_upto):
        qw, qh = g.w[i + 2 * nxt] - c["width"] // 4, g.h[i + 2 * nxt] - c["height"] // 4
        if qw <= 0 or qh <= 0:
            continue
        pad = 32
        padded = lambda a: np.pad(a, ((0, pad), (0, pad)))   # noqa: E731
        p0, p1 = padded(pyr.plane(i)), padded(pyr.plane(i + nxt))
        p2 = [padded(pyr.plane(i + 2 * nxt, q)) for q in range(4)]
        kstar = np.full((8, 4, qh, qw), -1, np.int32)
        for q in range(4):
            dx, dy = q & 1, q >> 1
            gy, gx = np.mgrid[0:qh, 0:qw]
            gy, gx = gy.ravel(), gx.ravel()
            alive = np.arange(gx.size)
            for j in range(8):
                s = st[j]
                if alive.size == 0:
                    break
                ax, ay = gx[alive], gy[alive]

                def px(pt):
                    z, x, y = pt
                    if z == 0:
                        return p0[4 * ay + 2 * dy + y, 4 * ax + 2 * dx + x]
                    if z == 1:
                        return p1[2 * ay + dy + y, 2 * ax + dx + x]
                    return p2[q][ay + y, ax + x]
                # exact integers: U = sum over fired features of (a_pass - a_fail) * 1e8, threshold T' = thr - sum(a_fail)
                gains = [int(round((a_pass - a_fail) * 1e8)) for (_, _, a_fail, a_pass) in s["feats"]]
                T = int(round(s["thr"] * 1e8)) - sum(int(round(a_fail * 1e8)) for (_, _, a_fail, _) in s["feats"])
                rem = np.cumsum(gains[::-1])[::-1]                       # rem[k] = gains[k] + ... (what features k.. can add)
                U = np.zeros(alive.size, np.int64)
                dec = np.full(alive.size, len(gains) - 1, np.int32)
                undecided = np.ones(alive.size, bool)
                for k, (pp, nn, a_fail, a_pass) in enumerate(s["feats"]):
                    pm = np.minimum.reduce([px(p) for p in pp])
                    nm = np.maximum.reduce([px(p) for p in nn])
                    U = U + np.where(pm > nm, gains[k], 0)
                    rest = rem[k + 1] if k + 1 < len(gains) else 0
                    now = undecided & ((U > T) | (U + rest < T))          # !(sum < thr)  <=>  U >= T; ties stay undecided
                    dec[now] = k
                    undecided &= ~now
                kstar[j, q].ravel()[alive] = dec
                alive = alive[U >= T]
        out.append((qw, qh, kstar))
    return st, out


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    every = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    blob = synth.load_cascade_blob()
    frames = [synth.frame(first + f, 640, 480) for f in range(4)]
    evs = []
    for f in frames:
        st, sc = evaluate(f, blob)
        evs.append(sc)
    loads = [[len(pp) + len(nn) for (pp, nn, _, _) in s["feats"]] for s in st]
    full = dict((j, 0) for j in range(2, 8))
    early = dict((j, 0) for j in range(2, 8))
    iters = dict((j, 0) for j in range(2, 8))
    lanes_sum = dict((j, 0) for j in range(2, 8))
    for si in range(len(evs[0])):
        qw, qh, _ = evs[0][si]
        for ty in range((qh + TH - 1) // TH):
            for tx in range((qw + TW - 1) // TW):
                ys = np.arange(ty * TH, min(qh, ty * TH + TH))
                xs = np.arange(tx * TW, min(qw, tx * TW + TW))
                ks = np.stack([evs[f][si][2][:, :, ys[:, None], xs[None, :]] for f in range(4)])   # [f, stage, q, ny, nx]
                ff, qq, ly, lx = np.meshgrid(np.arange(4), np.arange(4), ys - ty * TH, xs - tx * TW, indexing="ij")
                ks = np.moveaxis(ks, 1, 0).reshape(8, -1)                                          # [stage, windows]
                ff, qq, ly, lx = ff.ravel(), qq.ravel(), ly.ravel(), lx.ravel()
                u, v = 2 * lx + (qq & 1), 2 * ly + (qq >> 1)
                cls = (u + 12 * v) & 31
                bit = 8 * v + 4 * (u >> 5) + ff
                for grp in GROUPS:
                    j0 = grp[0]
                    idx = np.nonzero(ks[j0] >= 0)[0]                    # windows alive at the group's first stage
                    if idx.size == 0:
                        break
                    # per class: entries in bit order; warp w takes ranks [w n / 8, (w+1) n / 8)
                    per_warp_lane = [[[] for _ in range(32)] for _ in range(NWARP)]
                    for c in range(32):
                        e = idx[cls[idx] == c]
                        e = e[np.argsort(bit[e])]
                        n = e.size
                        for w in range(NWARP):
                            per_warp_lane[w][c] = list(e[(w * n) // NWARP:((w + 1) * n) // NWARP])
                    for w in range(NWARP):
                        n_it = max(len(l) for l in per_warp_lane[w])
                        for it in range(n_it):
                            lanes = np.array([l[it] for l in per_warp_lane[w] if len(l) > it])
                            live = lanes
                            for j in grp:
                                live = live[ks[j, live] >= 0]
                                if live.size == 0:
                                    break
                                iters[j] += 1
                                lanes_sum[j] += live.size
                                full[j] += sum(loads[j])
                                kmax = int(ks[j, live].max())
                                stop = min(len(loads[j]) - 1, ((kmax // every) + 1) * every - 1)   # next checkpoint
                                early[j] += sum(loads[j][: stop + 1])
    print(f"frames {first}..{first + 3}, checkpoint every {every} features; byte loads (= wavefronts) per quad")
    tf = te = 0
    for j in range(2, 8):
        print(f"  stage {j}: {len(loads[j]):3d} features, warp iterations {iters[j]:7d}, live lanes/iteration {lanes_sum[j] / max(iters[j], 1):5.1f}, "
              f"loads full {full[j]:9d}  with exits {early[j]:9d}  ({100 * early[j] / max(full[j], 1):5.1f} %)")
        tf += full[j]
        te += early[j]
    print(f"  total stages 2-7: {tf} -> {te} ({100 * te / tf:.1f} %), per frame {tf / 4 / 1e3:.0f} k -> {te / 4 / 1e3:.0f} k wavefronts")


if __name__ == "__main__":
    main()
