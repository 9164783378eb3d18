"""headtrackr.headposition mirror — /root/reference/src/headposition.js:35-218 (host-side, ~30 flops/frame).

Face box (centre x,y, width, height in camera pixels) + field-of-view model -> head position (x, y, z) in cm
relative to the screen centre; emits the `headtrackingEvent` {x, y, z} through callbacks (the reference uses
document.dispatchEvent, src/headposition.js:183-188).
"""
import math


class TrackObj:                                                     # src/headposition.js:206-218
    def __init__(self, x=None, y=None, z=None):
        self.x, self.y, self.z = x, y, z

    def clone(self):
        return TrackObj(self.x, self.y, self.z)


class Tracker:
    def __init__(self, facetrackrObj, camwidth, camheight, params=None):
        params = dict(params or {})
        self._params = params
        self.edgecorrection = params.get("edgecorrection", True)   # :44-48
        self.camheight_cam = camheight
        self.camwidth_cam = camwidth
        head_width_cm, head_height_cm = 16, 19                      # :53-54
        self._hsa = math.atan(head_width_cm / head_height_cm)       # :57
        self._head_diag_cm = math.sqrt((head_width_cm * head_width_cm) + (head_height_cm * head_height_cm))
        self._sin, self._cos, self._tan = math.sin(self._hsa), math.cos(self._hsa), math.tan(self._hsa)
        iw, ih = facetrackrObj["width"], facetrackrObj["height"]   # :66-68
        self._head_diag_cam = math.sqrt((iw * iw) + (ih * ih))
        if params.get("fov") is None:                               # :69-84
            head_width_cam = self._sin * self._head_diag_cam
            camwidth_at_default_face_cm = (self.camwidth_cam / head_width_cam) * head_width_cm
            distance_to_screen = params.get("distance_to_screen", 60)
            self._fov_width = math.atan((camwidth_at_default_face_cm / 2) / distance_to_screen) * 2
        else:
            self._fov_width = params["fov"] * math.pi / 180
        self._tan_fov_width = 2 * math.tan(self._fov_width / 2)     # :87
        self._x = self._y = self._z = None
        self._listeners = []

    def addEventListener(self, fn):
        self._listeners.append(fn)

    def track(self, facetrackrObj):                                 # :91-191
        w, h = facetrackrObj["width"], facetrackrObj["height"]
        fx, fy = facetrackrObj["x"], facetrackrObj["y"]
        sin_hsa, cos_hsa, tan_hsa = self._sin, self._cos, self._tan
        hdc = self._head_diag_cam
        if self.edgecorrection:
            margin = 11
            leftDistance = fx - (w / 2)
            rightDistance = self.camwidth_cam - (fx + (w / 2))
            topDistance = fy - (h / 2)
            bottomDistance = self.camheight_cam - (fy + (h / 2))
            onVerticalEdge = leftDistance < margin or rightDistance < margin
            onHorizontalEdge = topDistance < margin or bottomDistance < margin
            if onHorizontalEdge:
                if onVerticalEdge:                                  # corner: keep the previous diagonal
                    if leftDistance < margin:
                        fx = w - (hdc * sin_hsa / 2)
                    else:
                        fx = fx - (w / 2) + (hdc * sin_hsa / 2)
                    if topDistance < margin:
                        fy = h - (hdc * cos_hsa / 2)
                    else:
                        fy = fy - (h / 2) + (hdc * cos_hsa / 2)
                else:                                               # top / bottom edge
                    if topDistance < margin:
                        ow, ew = topDistance / margin, (margin - topDistance) / margin
                        fy = h - (ow * (h / 2) + ew * ((w / tan_hsa) / 2))
                        hdc = ew * (w / sin_hsa) + ow * (math.sqrt((w * w) + (h * h)))
                    else:
                        ow, ew = bottomDistance / margin, (margin - bottomDistance) / margin
                        fy = fy - (h / 2) + (ow * (h / 2) + ew * ((w / tan_hsa) / 2))
                        hdc = ew * (w / sin_hsa) + ow * (math.sqrt((w * w) + (h * h)))
            elif onVerticalEdge:                                    # left / right edge
                if leftDistance < margin:
                    ow, ew = leftDistance / margin, (margin - leftDistance) / margin
                    hdc = ew * (h / cos_hsa) + ow * (math.sqrt((w * w) + (h * h)))
                    fx = w - (ow * (w / 2) + ew * (h * tan_hsa / 2))
                else:
                    ow, ew = rightDistance / margin, (margin - rightDistance) / margin
                    hdc = ew * (h / cos_hsa) + ow * (math.sqrt((w * w) + (h * h)))
                    fx = fx - (w / 2) + (ow * (w / 2) + ew * (h * tan_hsa / 2))
            else:
                hdc = math.sqrt((w * w) + (h * h))
        else:
            hdc = math.sqrt((w * w) + (h * h))
        self._head_diag_cam = hdc
        z = (self._head_diag_cm * self.camwidth_cam) / (self._tan_fov_width * hdc)          # :165
        x = -((fx / self.camwidth_cam) - 0.5) * z * self._tan_fov_width                      # :170
        y = -((fy / self.camheight_cam) - 0.5) * z * self._tan_fov_width * (self.camheight_cam / self.camwidth_cam)
        y = y + self._params.get("distance_from_camera_to_screen", 11.5)                      # :175-180
        self._x, self._y, self._z = x, y, z
        evt = dict(type="headtrackingEvent", x=x, y=y, z=z)
        for fn in self._listeners:
            fn(evt)
        return TrackObj(x, y, z)

    def getTrackerObj(self):
        return TrackObj(self._x, self._y, self._z)

    def getFOV(self):
        return self._fov_width * 180 / math.pi
