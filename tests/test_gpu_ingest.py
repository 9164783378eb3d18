"""GPU: ht_ingest - the video -> canvas drawImage of src/main.js:170 - against the oracle's canvas-shim resampler,
and the ingested (device-resident) canvases fed straight into the detector."""
import numpy as np
import pytest

import oracle
from headtrackr_b200 import synth
from test_ingest_host import oracle_resize

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sw,sh,dw,dh", [(640, 480, 320, 240), (333, 251, 320, 240), (160, 120, 320, 240), (1280, 720, 640, 360)])
def test_ingest_matches_oracle(ctx, sw, sh, dw, dh):
    frames = synth.batch(3, sw, sh, start=80)
    frames[1, ..., 3] = np.arange(sw, dtype=np.uint8)[None, :]
    out = ctx.ingest(frames, dw, dh)
    assert out.shape == (3, dh, dw, 4)
    for i in range(3):
        assert np.array_equal(out[i], oracle_resize(frames[i], dw, dh)), i


def test_one_to_one_ingest_is_a_copy(ctx):
    f = synth.batch(2, 320, 240, start=5)
    assert np.array_equal(ctx.ingest(f, 320, 240), f)


def test_ingested_canvas_feeds_the_detector_on_the_device(ctx, blob):
    import torch
    video = synth.batch(2, 640, 480, start=90)
    canvas = torch.zeros((2, 240, 320, 4), dtype=torch.uint8, device="cuda")
    ctx.ingest(torch.from_numpy(video).cuda(), 320, 240, out=canvas)
    got = ctx.detect(canvas, 5, 1)
    for i in range(2):
        want = oracle.detect(oracle_resize(video[i], 320, 240), blob)
        assert [(d["x"], d["y"], d["width"], d["height"], d["confidence"], d["neighbors"]) for d in got[i]] == want
