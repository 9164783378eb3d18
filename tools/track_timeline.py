#!/usr/bin/env python
"""Per-stream timeline of one k_track launch on the bench mix (HT_TRACK_TRACE=1): when each stream's cluster
started and ended, on which SM, and how many passes it ran.  Prints the schedule's summary; used to decide the
launch order / cluster sizes.  Usage (GPU box): python tools/track_timeline.py [n] [env=val ...]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from headtrackr_b200 import synth  # noqa: E402
from headtrackr_b200.context import Context  # noqa: E402

W, H = 640, 480


def main():
    n = 1024
    for a in sys.argv[1:]:
        if "=" in a:
            k, v = a.split("=", 1)
            os.environ[k] = v
        else:
            n = int(a)
    os.environ["HT_TRACK_TRACE"] = "1"
    frames = np.stack([synth.frame(i, W, H) for i in range(64)])
    # bench.py's mix: repetition r of the 64 base frames is rolled by 16 r pixels
    batch = np.stack([np.roll(frames[j % 64], (j // 64) * 16, axis=1) for j in range(n)])
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = Context(max_width=W, max_height=H, max_frames=n, device=0, stream=stream.cuda_stream)
    ctx.set_track_memo(os.environ.get("TIMELINE_MEMO", "0") == "1")   # strict (bench headline) unless asked
    d = torch.from_numpy(batch).cuda()
    K = ctx.K
    outs = (torch.zeros(n * K * 12, dtype=torch.int32, device="cuda"), torch.zeros(n, dtype=torch.int32, device="cuda"),
            torch.zeros(n, dtype=torch.int32, device="cuda"), torch.zeros(n * 6, dtype=torch.int32, device="cuda"),
            torch.zeros(n * 4, dtype=torch.int32, device="cuda"))
    for _ in range(2):
        ctx.detect_track(d, n_calls=30, outputs=outs)
    torch.cuda.synchronize()
    ctx.profile(True)
    ctx.detect_track(d, n_calls=30, outputs=outs)
    torch.cuda.synchronize()
    ms = ctx.profile_read()["track"][0]
    print("stats of the 3 launches:", ctx.debug_track_stats())
    tr = ctx.debug_track_trace(n).astype(np.int64)
    t0 = tr[:, 0].min()
    start = (tr[:, 0] - t0) / 1e3
    end = (tr[:, 1] - t0) / 1e3
    ok = tr[:, 1] > 0
    end = np.where(ok, end, start)
    dur = end - start
    passes = tr[:, 3]
    print(f"n={n} track kernel(s) {ms:.3f} ms; span of the timeline {end.max():.0f} us")
    print(f"stream duration us: mean {dur.mean():.0f} p50 {np.median(dur):.0f} p90 {np.percentile(dur, 90):.0f} max {dur.max():.0f}"
          f"   sum {dur.sum() / 1e3:.0f} ms-slots")
    print(f"us per pass: mean {(dur / np.maximum(passes, 1)).mean():.1f}  (heaviest 16 streams: "
          f"{(dur / np.maximum(passes, 1))[np.argsort(-dur)[:16]].mean():.1f})")
    # resident streams over time
    edges = np.linspace(0, end.max(), 25)
    for a, b in zip(edges[:-1], edges[1:]):
        mid = (a + b) / 2
        print(f"  t={mid:7.0f} us  resident streams {int(((start <= mid) & (end > mid)).sum()):5d}   started so far "
              f"{int((start <= mid).sum()):5d}")
    area = None
    try:
        wins = outs[4].cpu().numpy().reshape(n, 4)
        area = wins[:, 2].astype(np.int64) * wins[:, 3]
    except Exception:
        pass
    order = np.argsort(-dur)
    print("  heaviest streams: " + ", ".join(f"{i}:{dur[i]:.0f}us/{passes[i]}p" + (f"/{area[i]}px" if area is not None else "")
                                           for i in order[:12]))
    print(f"  passes: mean {passes.mean():.1f} p50 {np.median(passes):.0f} p90 {np.percentile(passes, 90):.0f} max {passes.max()}")
    late = np.argsort(-end)[:8]
    for i in late:
        print(f"  last finishers: stream {i} (frame {i % 64}) start {start[i]:.0f} end {end[i]:.0f} passes {passes[i]} sm {tr[i, 2]}")

    phases(ctx, n, tr)




def phases(ctx, n, tr):
    """Per-phase cycles of a pass (profiling builds, HT_LIB=...libht_ptrace.so): where does a pass spend its time?"""
    ph = ctx.debug_track_phases(n).astype(np.float64)
    if ph.sum() == 0:
        return
    passes = np.maximum(tr[:, 3].astype(np.float64), 1)
    names = ["pixel loop", "warp sums + CTA barrier", "exchange + cluster barrier", "mean-shift step", "publish barrier"]
    order = np.argsort(-passes * 1e6 - (tr[:, 1] - tr[:, 0]))
    groups = {"16 longest chains": order[:16], "next 64": order[16:80], "median 128": order[len(order) // 2 - 64: len(order) // 2 + 64],
              "all": order}
    print("cycles per pass (leader thread, SM clock) by phase:")
    for gname, idx in groups.items():
        per = ph[idx, :5].sum(axis=0) / passes[idx].sum()
        tot = per.sum()
        print(f"  {gname:18s} total {tot:8.0f}  " + "  ".join(f"{nm}: {v:7.0f} ({100 * v / tot:4.1f}%)" for nm, v in zip(names, per)))


if __name__ == "__main__":
    main()
