"""headtrackr.Smoother mirror — /root/reference/src/smoother.js:13-88 (host-side scalar post-processing).

Kept host-side as in the reference (5 scalars per frame).  Quirks preserved on purpose:
  * `sp2 = sp` aliases the two state arrays (src/smoother.js:28), so the "double" exponential smoother is a
    single one: sp2[i] = alpha*sp[i] + (1-alpha)*sp2[i] is evaluated on the value just written;
  * init() copies initPos.z, which the caller never sets (src/main.js:259) -> z stays NaN;
  * predict() is called unbound, so `this.interpolate` is undefined -> the non-interpolating branch always runs;
  * updateTime is refreshed just before it is subtracted (src/smoother.js:44-46): msDiff is 0, step is 0.
"""
import math


def _num(v):
    return math.nan if v is None else float(v)


class Smoother:
    def __init__(self, alpha, interval):
        self.alpha = alpha
        self.interval = interval
        self.initialized = False
        self.interpolate = False
        self._sp = None
        self._sp2 = None

    def init(self, initPos):                                        # src/smoother.js:25-30
        self.initialized = True
        self._sp = [_num(initPos.get("x")), _num(initPos.get("y")), _num(initPos.get("z")),
                    _num(initPos.get("width")), _num(initPos.get("height"))]
        self._sp2 = self._sp                                        # aliasing, as in the reference

    def smooth(self, pos):                                          # src/smoother.js:32-59
        if not self.initialized:
            return False
        a = self.alpha
        positions = [_num(pos.get("x")), _num(pos.get("y")), _num(pos.get("z")), _num(pos.get("width")),
                     _num(pos.get("height"))]
        sp, sp2 = self._sp, self._sp2
        for i in range(5):
            sp[i] = a * positions[i] + (1 - a) * sp[i]
            sp2[i] = a * sp[i] + (1 - a) * sp2[i]
        new = self._predict(0.0)
        pos["x"], pos["y"], pos["z"], pos["width"], pos["height"] = new
        return pos

    def _predict(self, time):                                       # src/smoother.js:61-87, non-interpolating branch
        step = float(int(time / self.interval))
        ratio = (self.alpha * step) / (1 - self.alpha)
        a, b = 2 + ratio, 1 + ratio
        return [a * self._sp[i] - b * self._sp2[i] for i in range(5)]
