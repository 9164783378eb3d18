"""CPU: host-side logic — the facetrackr state machine mirror (src/facetrackr.js) driven by an
oracle-backed backend, the cascade packer invariants, the synthetic frame generator."""
import math

import numpy as np

import oracle
from headtrackr_b200 import camshift, facetrackr, synth
from headtrackr_b200.canvas import Canvas


class OracleBackend:
    """facetrackr's three pixel calls served by the CPU oracle (test double of CudaBackend)."""

    def __init__(self, blob):
        self.blob = blob

    def detect_objects(self, canvas, interval, min_neighbors):
        return [dict(x=r[0], y=r[1], width=r[2], height=r[3], confidence=r[4], neighbors=r[5])
                for r in oracle.detect(canvas.pixels, self.blob, interval, min_neighbors)]

    def new_tracker(self, calc_angles):
        ot = oracle.CamshiftTracker(calc_angles=calc_angles)

        class T:
            def initTracker(self, canvas, rect):
                ot.init_tracker(canvas.pixels, int(rect.x), int(rect.y), int(rect.width), int(rect.height))

            def track(self, canvas):
                ot.track(canvas.pixels)

            def getTrackObj(self):
                o = camshift.TrackObj()
                d = ot.track_obj()
                o.x, o.y, o.width, o.height, o.angle = d["x"], d["y"], d["width"], d["height"], d["angle"]
                return o
        return T()

    def whitebalance(self, canvas):
        return oracle.whitebalance(canvas.pixels)


def test_state_machine_wb_vj_cs(blob):
    canvas = Canvas(synth.frame(3, 160, 120, n_faces=1))
    events = []
    ft = facetrackr.Tracker({"whitebalancing": True}, backend=OracleBackend(blob))
    ft.addEventListener(events.append)
    ft.init(canvas)
    for _ in range(15):                                           # 15 stable whitebalance samples (src/facetrackr.js:59,82-94)
        ft.track()
        assert ft.getTrackingObject().detection == "WB"
    assert not events
    ft.track()
    vj = ft.getTrackingObject()
    assert vj.detection == "VJ" and vj.confidence > -10           # no event on the hand-off frame (:112)
    assert not events
    ft.track()
    cs = ft.getTrackingObject()
    assert cs.detection == "CS" and cs.confidence == 1 and cs.width > 0
    assert len(events) == 1 and events[0]["type"] == "facetrackingEvent" and events[0]["detection"] == "CS"
    assert set(events[0]) == {"type", "height", "width", "angle", "x", "y", "confidence", "detection", "time"}


def test_state_machine_stays_in_vj_without_a_face(blob):
    canvas = Canvas(synth.frame(0, 160, 120, kind="constant"))
    ft = facetrackr.Tracker({"whitebalancing": False}, backend=OracleBackend(blob))
    ft.init(canvas)
    ft.track()
    o = ft.getTrackingObject()
    assert o.detection == "VJ" and o.confidence == -10000 and o.width == 0
    ft.track()
    assert ft.getTrackingObject().detection == "VJ"


def test_cascade_blob_invariants(blob):
    c = synth.parse_blob(blob)
    assert c["n_stages"] == 16 and c["n_features"] == 2015 and (c["width"], c["height"]) == (24, 24)
    assert [s[0] for s in c["stages"]] == [4, 4, 7, 13, 20, 22, 32, 45, 61, 80, 115, 153, 203, 301, 391, 564]
    for f in c["features"]:
        assert f["a_fail"] == -f["a_pass"] < 0 and 2 <= f["size"] <= 5
        assert f["p"][0][0] >= 0 and f["n"][0][0] >= 0
        assert abs(round(f["a_pass"] * 1e8) / 1e8 - f["a_pass"]) == 0   # 8-digit decimals: exact integer stage sums


def test_template_passes_every_stage(blob):
    """A face synthesised from the cascade itself survives all 16 stages."""
    t = synth.face_template(blob)
    frame = np.zeros((240, 320, 4), np.uint8)
    frame[..., 3] = 255
    frame[..., :3] = 128
    big = synth.shim_resize(t, 96, 96)
    frame[60:156, 100:196, :3] = big[..., None]
    res = oracle.detect(frame, blob)
    assert len(res) >= 1 and max(r[4] for r in res) > 0


def test_synthetic_frames_are_deterministic():
    a, b = synth.frame(5, 160, 120), synth.frame(5, 160, 120)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.shape == (120, 160, 4) and (a[..., 3] == 255).all()
    assert not np.array_equal(a, synth.frame(6, 160, 120))


def test_tracker_slots_are_a_per_context_free_list():
    """the 17th Tracker of a 16-slot context used to share slot 0 silently."""
    import pytest

    class FakeCtx:
        max_frames = 3
    c = FakeCtx()
    got = [camshift._take_slot(c) for _ in range(3)]
    assert sorted(got) == [0, 1, 2] and c._live_trackers == 3
    with pytest.raises(RuntimeError):
        camshift._take_slot(c)
    camshift._give_slot(c, got[1])
    assert camshift._take_slot(c) == got[1]
    other = FakeCtx()
    assert camshift._take_slot(other) == 0                        # the pool is per context, not global


def test_detect_objects_only_accepts_a_grayscaled_canvas():
    """the reference's detect_objects does no graying (src/ccv.js:171-192); ht_detect always does, so the
    mirror refuses the one call shape it could not serve identically."""
    import pytest
    from headtrackr_b200 import ccv
    with pytest.raises(TypeError):
        ccv.detect_objects(Canvas(synth.frame(0, 160, 120)), None, 5, 1, context=object())
