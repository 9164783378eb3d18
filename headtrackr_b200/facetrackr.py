"""headtrackr.facetrackr mirror — /root/reference/src/facetrackr.js (host-side state machine).

The reference keeps this layer in JavaScript; only its three inner calls touch pixels
(ccv.detect_objects, camshift.Tracker, getWhitebalance) and those go to the CUDA library.  The
`backend` argument lets the CPU tests drive the same state machine with the oracle.
DOM events (document.dispatchEvent, src/facetrackr.js:112-125) become callbacks.
"""
import math
import time

from . import camshift as _camshift


class TrackObj:                                                    # src/facetrackr.js:233-255
    def __init__(self):
        self.height = 0
        self.width = 0
        self.angle = 0
        self.x = 0
        self.y = 0
        self.confidence = -10000
        self.detection = ""
        self.time = 0

    def clone(self):
        c = TrackObj()
        c.__dict__.update(self.__dict__)
        return c


class CudaBackend:
    """The three pixel entry points of facetrackr, served by libheadtrackr_b200."""

    def __init__(self, context=None):
        self.context = context

    def detect_objects(self, canvas, interval, min_neighbors):
        from . import ccv
        return ccv.detect_objects(ccv.grayscale(canvas), None, interval, min_neighbors, context=self.context)

    def new_tracker(self, calc_angles):
        return _camshift.Tracker({"calcAngles": calc_angles}, context=self.context)

    def whitebalance(self, canvas):
        from . import getWhitebalance
        return getWhitebalance(canvas, context=self.context)


class Tracker:
    def __init__(self, params=None, backend=None):
        params = dict(params or {})
        self.sendEvents = params.get("sendEvents", True)           # src/facetrackr.js:41-53
        self.whitebalancing = params.get("whitebalancing", True)
        self.calcAngles = params.get("calcAngles", False)
        self._currentDetection = "WB" if self.whitebalancing else "VJ"
        self._confidenceThreshold = -10                            # :57
        self.previousWhitebalances = []                            # :58
        self.pwbLength = 15                                        # :59
        self._backend = backend or CudaBackend()
        self._listeners = []
        self._inputcanvas = None
        self._curtracked = None
        self._cstracker = None

    def addEventListener(self, fn):
        """fn(evt) receives the facetrackingEvent dict (src/facetrackr.js:112-125)."""
        self._listeners.append(fn)

    def init(self, inputcanvas):                                   # :61-65
        self._inputcanvas = inputcanvas
        self._cstracker = self._backend.new_tracker(self.calcAngles)

    def track(self):                                               # :67-126
        if self._currentDetection == "WB":
            result = self._checkWhitebalance()
        elif self._currentDetection == "VJ":
            result = self._doVJDetection()
        else:
            result = self._doCSDetection()
        if result.detection == "WB":                               # :79-95
            if len(self.previousWhitebalances) >= self.pwbLength:
                self.previousWhitebalances.pop()
            self.previousWhitebalances.insert(0, result.wb)
            if len(self.previousWhitebalances) == self.pwbLength:
                if (max(self.previousWhitebalances) - min(self.previousWhitebalances)) < 2:
                    self._currentDetection = "VJ"
        if result.detection == "VJ" and result.confidence > self._confidenceThreshold:   # :97-108
            self._currentDetection = "CS"
            rect = _camshift.Rectangle(math.floor(result.x), math.floor(result.y), math.floor(result.width),
                                       math.floor(result.height))
            self._cstracker.initTracker(self._inputcanvas, rect)
        self._curtracked = result
        if result.detection == "CS" and self.sendEvents:           # :112-125
            evt = dict(type="facetrackingEvent", height=result.height, width=result.width, angle=result.angle,
                       x=result.x, y=result.y, confidence=result.confidence, detection=result.detection,
                       time=result.time)
            for fn in self._listeners:
                fn(evt)

    def getTrackingObject(self):                                   # :128-130
        return self._curtracked.clone()

    def _doVJDetection(self):                                      # :133-182
        start = time.time()
        comp = self._backend.detect_objects(self._inputcanvas, 5, 1)
        diff = int((time.time() - start) * 1000)
        candidate = comp[0] if len(comp) > 0 else None
        for i in range(1, len(comp)):
            if comp[i]["confidence"] > candidate["confidence"]:
                candidate = comp[i]
        result = TrackObj()
        if candidate is not None:
            result.width = candidate["width"]
            result.height = candidate["height"]
            result.x = candidate["x"]
            result.y = candidate["y"]
            result.confidence = candidate["confidence"]
        result.time = diff
        result.detection = "VJ"
        return result

    def _doCSDetection(self):                                      # :185-217
        start = time.time()
        self._cstracker.track(self._inputcanvas)
        cs = self._cstracker.getTrackObj()
        diff = int((time.time() - start) * 1000)
        result = TrackObj()
        result.width, result.height, result.x, result.y, result.angle = cs.width, cs.height, cs.x, cs.y, cs.angle
        result.confidence = 1
        result.time = diff
        result.detection = "CS"
        return result

    def _checkWhitebalance(self):                                  # :220-227
        result = TrackObj()
        result.wb = self._backend.whitebalance(self._inputcanvas)
        result.detection = "WB"
        return result
