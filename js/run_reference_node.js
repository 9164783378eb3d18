// js/run_reference_node.js — where a real Node exists: run the UNMODIFIED reference source over the same
// canvas shim the oracle defines and print the rect list as JSON, to compare with tests/golden/reference_js.json.
// (In this container the same job is done by oracle/jsmini.py; see tools/make_goldens.py.)
//
//   node js/run_reference_node.js /path/to/headtrackr/src frame.rgba 320 240 > out.json
'use strict';
var fs = require('fs'), path = require('path'), vm = require('vm');
var src = process.argv[2], W = +process.argv[4], H = +process.argv[5];
var rgba = new Uint8ClampedArray(fs.readFileSync(process.argv[3]));

function Canvas(w, h, data) {
  this.width = w || 0; this.height = h || 0;
  var self = this, pix = data || new Uint8ClampedArray(this.width * this.height * 4);
  function ensure() { if (pix.length !== self.width * self.height * 4) pix = new Uint8ClampedArray(self.width * self.height * 4); }
  this.getContext = function () {
    return {
      getImageData: function (x, y, w, h) {
        ensure();
        var out = new Uint8ClampedArray(w * h * 4);
        for (var j = 0; j < h; j++) for (var i = 0; i < w; i++) {
          var cx = x + i, cy = y + j;
          if (cx >= 0 && cx < self.width && cy >= 0 && cy < self.height)
            for (var c = 0; c < 4; c++) out[(j * w + i) * 4 + c] = pix[(cy * self.width + cx) * 4 + c];
        }
        return {width: w, height: h, data: out};
      },
      putImageData: function (img) { ensure(); pix.set(img.data); },
      createImageData: function (w, h) { return {width: w, height: h, data: new Uint8ClampedArray(w * h * 4)}; },
      // the resampler DEFINED in oracle/ht_oracle.h: exact integer bilinear, pixel centres, round half up
      drawImage: function (s, sx, sy, sw, sh, dx, dy, dw, dh) {
        ensure();
        if (arguments.length === 5) { dw = dx; dh = dy; dx = sx; dy = sy; sx = 0; sy = 0; sw = s.width; sh = s.height; }
        if (dw <= 0 || dh <= 0) return;
        var sp = s.getContext().getImageData(0, 0, s.width, s.height).data;
        for (var Y = 0; Y < dh; Y++) {
          var vn = (2 * Y + 1) * sh - dh, y0 = Math.floor(vn / (2 * dh)), fy = vn - y0 * 2 * dh;
          var ya = Math.min(Math.max(y0, 0), sh - 1) + sy, yb = Math.min(Math.max(y0 + 1, 0), sh - 1) + sy;
          for (var X = 0; X < dw; X++) {
            var un = (2 * X + 1) * sw - dw, x0 = Math.floor(un / (2 * dw)), fx = un - x0 * 2 * dw;
            var xa = Math.min(Math.max(x0, 0), sw - 1) + sx, xb = Math.min(Math.max(x0 + 1, 0), sw - 1) + sx;
            for (var c = 0; c < 4; c++) {
              var num = (2 * dw - fx) * (2 * dh - fy) * sp[(ya * s.width + xa) * 4 + c] + fx * (2 * dh - fy) * sp[(ya * s.width + xb) * 4 + c] +
                        (2 * dw - fx) * fy * sp[(yb * s.width + xa) * 4 + c] + fx * fy * sp[(yb * s.width + xb) * 4 + c];
              pix[(Y * self.width + X) * 4 + c] = Math.floor((num + 2 * dw * dh) / (4 * dw * dh));
            }
          }
        }
      }
    };
  };
}
var sandbox = {headtrackr: {}, document: {createElement: function () { return new Canvas(); }}, Math: Math};
vm.createContext(sandbox);
['ccv.js', 'cascade.js', 'camshift.js', 'whitebalance.js'].forEach(function (f) {
  vm.runInContext(fs.readFileSync(path.join(src, f), 'utf8'), sandbox);
});
var canvas = new Canvas(W, H, rgba);
var res = sandbox.headtrackr.ccv.detect_objects(sandbox.headtrackr.ccv.grayscale(canvas), sandbox.headtrackr.cascade, 5, 1);
console.log(JSON.stringify(res));
