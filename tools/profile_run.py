#!/usr/bin/env python
"""Small fixed workload for ncu captures: N frames of 640x480 through ht_detect_track, a few iterations.

    ncu --set full --clock-control none --import-source on -k regex:k_cascade -s 1 -c 1 -o gpurun_out/prof \
        python tools/profile_run.py --frames 32 --iters 2
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--track-calls", type=int, default=30)
    args = ap.parse_args()
    import torch
    from headtrackr_b200 import Context, synth
    frames = np.stack([synth.frame(i, args.width, args.height) for i in range(args.frames)])
    dev = torch.from_numpy(frames).cuda()
    ctx = Context(max_width=args.width, max_height=args.height, max_frames=args.frames)
    K = ctx.K
    outs = (torch.zeros((args.frames, K, 6), dtype=torch.float64, device="cuda"),
            torch.zeros((args.frames,), dtype=torch.int32, device="cuda"),
            torch.zeros((args.frames,), dtype=torch.int32, device="cuda"),
            torch.zeros((args.frames, 6), dtype=torch.int32, device="cuda"),
            torch.zeros((args.frames, 4), dtype=torch.int32, device="cuda"))
    for _ in range(args.iters):
        ctx.detect_track(dev, 5, 1, calc_angles=False, n_calls=args.track_calls, outputs=outs)
    ctx.sync()
    print("counts", outs[1][:8].tolist(), "stats", ctx.debug_track_stats())
    ctx.close()


if __name__ == "__main__":
    main()
