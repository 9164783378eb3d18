"""GPU: pipelined ht_detect_track calls (ht_set_pipeline) give the results of unpipelined ones.

The tracking of call s is left on the library's second stream and runs under the detection of call s+1
(src/facetrackr.js:97-108,190 for a batch); these tests pin what that must not change: every output
of every call, the ordering of accesses to output arrays that the next call re-uses, and that any other entry point
of the context joins first."""
import numpy as np
import pytest

from headtrackr_b200 import synth
from headtrackr_b200.context import Context

pytestmark = pytest.mark.gpu

W, H, N = 320, 240, 12


def batches():
    return [np.stack([synth.frame(7 * b + i, W, H) for i in range(N)]) for b in range(4)]


def new_outputs(torch, K, n):
    return (torch.zeros((n, K, 6), dtype=torch.float64, device="cuda"), torch.zeros((n,), dtype=torch.int32, device="cuda"),
            torch.zeros((n,), dtype=torch.int32, device="cuda"), torch.zeros((n, 6), dtype=torch.int32, device="cuda"),
            torch.zeros((n, 4), dtype=torch.int32, device="cuda"))


def host_copy(outs):
    return [o.cpu().numpy().copy() for o in outs]


def same_results(got, exp):
    """(rects, counts, found, objs, windows): rect entries beyond a frame's count are never written by the library and
    hold whatever an earlier call left there."""
    if not all(np.array_equal(a, b) for a, b in zip(got[1:], exp[1:])):
        return False
    return all(np.array_equal(got[0][f, :c], exp[0][f, :c]) for f, c in enumerate(exp[1]))


@pytest.mark.parametrize("n_calls", [1, 30])
def test_pipelined_calls_match_unpipelined(n_calls):
    import torch
    ctx = Context(max_width=W, max_height=H, max_frames=N, max_raw_per_frame=4096)
    try:
        ctx.set_track_memo(False)
        bs = [torch.from_numpy(b).cuda() for b in batches()]
        # reference run: pipeline off, sync after every call
        want = []
        for b in bs:
            outs = new_outputs(torch, ctx.K, N)      # fresh arrays: entries beyond a frame's count are never written
            ctx.detect_track(b, 5, 1, calc_angles=True, n_calls=n_calls, outputs=outs)
            ctx.sync()
            want.append(host_copy(outs))
        assert any(w[2].any() for w in want), "the batches must contain faces"
        # pipelined run: one output set per call, nothing is synchronised between the calls
        ctx.set_pipeline(True)
        sets = [new_outputs(torch, ctx.K, N) for _ in bs]
        for s_, b in enumerate(bs):
            ctx.detect_track(b, 5, 1, calc_angles=True, n_calls=n_calls, outputs=sets[s_])
        ctx.sync()
        for s_ in range(len(bs)):
            for got, exp in zip(host_copy(sets[s_]), want[s_]):
                assert np.array_equal(got, exp), s_
    finally:
        ctx.close()


def test_pipelined_same_output_set_every_call():
    """The caller may pass the SAME rectangle / count arrays to consecutive pipelined calls: k_group of call s+1 waits
    for the tracking of call s, which still reads them (hand-off, src/facetrackr.js:97-108)."""
    import torch
    ctx = Context(max_width=W, max_height=H, max_frames=N, max_raw_per_frame=4096)
    try:
        ctx.set_track_memo(False)
        bs = [torch.from_numpy(b).cuda() for b in batches()]
        outs = new_outputs(torch, ctx.K, N)
        want = []
        for b in bs:
            ctx.detect_track(b, 5, 1, calc_angles=False, n_calls=30, outputs=outs)
            ctx.sync()
            want.append(host_copy(outs))
        ctx.set_pipeline(True)
        for rep in range(3):
            for s, b in enumerate(bs):
                ctx.detect_track(b, 5, 1, calc_angles=False, n_calls=30, outputs=outs)
                ctx.sync()                       # joins: the results of THIS call are complete
                assert same_results(host_copy(outs), want[s]), (rep, s)
        # back-to-back without a sync: the last call's results after one final sync
        for s, b in enumerate(bs):
            ctx.detect_track(b, 5, 1, calc_angles=False, n_calls=30, outputs=outs)
        ctx.sync()
        assert same_results(host_copy(outs), want[-1])
    finally:
        ctx.close()


def test_other_entry_points_join_first(blob):
    """ht_track right after a pipelined ht_detect_track continues the trackers that call initialised."""
    import torch
    ctx = Context(max_width=W, max_height=H, max_frames=N, max_raw_per_frame=4096)
    try:
        ctx.set_track_memo(False)
        b = torch.from_numpy(batches()[0]).cuda()
        outs = new_outputs(torch, ctx.K, N)
        ctx.detect_track(b, 5, 1, calc_angles=False, n_calls=3, outputs=outs)
        ctx.sync()
        objs_a, wins_a = ctx.track(b, n_calls=2)
        ctx.set_pipeline(True)
        ctx.detect_track(b, 5, 1, calc_angles=False, n_calls=3, outputs=outs)
        objs_b, wins_b = ctx.track(b, n_calls=2)           # no sync in between: ht_track joins
        assert objs_a == objs_b and wins_a == wins_b
        # host outputs are never deferred
        dets, found, objs, wins = ctx.detect_track(batches()[0], 5, 1, calc_angles=False, n_calls=3)
        assert [o["x"] for o in objs] == outs[3].cpu().numpy()[:, 0].tolist()
    finally:
        ctx.close()


def test_pipelined_large_batch_with_tiers():
    """n >= 128 streams: k_track runs as three concurrent tiers on prioritised side streams (launch_track); pipelined
    and unpipelined calls, and a second pass whose launch order comes from the first one's history, agree."""
    import torch
    w, h, n = 160, 120, 192
    ctx = Context(max_width=w, max_height=h, max_frames=n, max_raw_per_frame=2048)
    try:
        ctx.set_track_memo(False)
        fr = np.stack([synth.frame(i % 48, w, h) for i in range(n)])
        for j in range(n):
            fr[j] = np.roll(fr[j], (j // 48) * 8, axis=1)
        b = torch.from_numpy(fr).cuda()
        outs = new_outputs(torch, ctx.K, n)
        ctx.detect_track(b, 5, 1, calc_angles=False, n_calls=30, outputs=outs)
        ctx.sync()
        want = host_copy(outs)
        assert want[2].sum() > n // 8
        ctx.set_pipeline(True)
        sets = [new_outputs(torch, ctx.K, n) for _ in range(3)]
        for s_ in range(3):
            ctx.detect_track(b, 5, 1, calc_angles=False, n_calls=30, outputs=sets[s_])
        ctx.sync()
        for s_ in range(3):
            for got, exp in zip(host_copy(sets[s_]), want):
                assert np.array_equal(got, exp), s_
    finally:
        ctx.close()
