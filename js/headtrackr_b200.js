// js/headtrackr_b200.js — drop the CUDA path under an unmodified headtrackr bundle.
//
// UNTESTED SOURCE (no Node in this image; see js/addon.cc).  Usage in a Node/Electron host that already
// loads headtrackr.js and has a canvas implementation:
//
//   var headtrackr = require('./headtrackr.js');
//   require('./headtrackr_b200.js').install(headtrackr, fs.readFileSync('cascade_face.bin'));
//
// After install(), facetrackr.js keeps calling headtrackr.ccv.detect_objects(headtrackr.ccv.grayscale(c),
// headtrackr.cascade, 5, 1), new headtrackr.camshift.Tracker(...) and headtrackr.getWhitebalance(c)
// (src/facetrackr.js:64,107,147-149,190-191,223) — the same names, argument meaning and result shapes —
// but the pixel work runs in libheadtrackr_b200.so.  main.js, smoother.js, headposition.js, ui.js stay as is.
'use strict';
var addon = require('./build/Release/headtrackr_b200_addon.node');

function pixels(canvas) {
  var d = canvas.getContext('2d').getImageData(0, 0, canvas.width, canvas.height);
  return d.data;  // Uint8ClampedArray RGBA, what every reference function reads
}

exports.install = function (headtrackr, cascadeBlob, opts) {
  var h = addon.create(cascadeBlob, opts || {});
  var nextSlot = 0;

  // src/ccv.js:22 — the CUDA detect fuses the grayscale pass, so grayscale() only tags the canvas.
  headtrackr.ccv.grayscale = function (canvas) { canvas.__ht_gray = true; return canvas; };

  // src/ccv.js:109
  headtrackr.ccv.detect_objects = function (canvas, cascade, interval, min_neighbors) {
    return addon.detect(h, pixels(canvas), 1, canvas.width, canvas.height, interval, min_neighbors)[0];
  };

  // src/camshift.js:148-354
  headtrackr.camshift.Tracker = function (params) {
    if (params === undefined) params = {};
    if (params.calcAngles === undefined) params.calcAngles = true;
    var slot = nextSlot++, trackObj = new headtrackr.camshift.TrackObj(), win = null, last = null;
    this.initTracker = function (canvas, area) {
      addon.trackInit(h, slot, pixels(canvas), canvas.width, canvas.height, area.x, area.y, area.width,
                      area.height, params.calcAngles ? 1 : 0);
      win = area.clone();
      trackObj = new headtrackr.camshift.TrackObj();
    };
    this.track = function (canvas) {
      if (canvas.width == 0 || canvas.height == 0) return;               // src/camshift.js:219
      last = canvas;
      var r = addon.track(h, slot, pixels(canvas), canvas.width, canvas.height, 1);
      trackObj = new headtrackr.camshift.TrackObj();
      trackObj.x = r.x; trackObj.y = r.y; trackObj.width = r.width; trackObj.height = r.height; trackObj.angle = r.angle;
      win = new headtrackr.camshift.Rectangle(r.window.x, r.window.y, r.window.width, r.window.height);
    };
    this.getTrackObj = function () { return trackObj.clone(); };
    this.getSearchWindow = function () { return win.clone(); };
    this.getBackProjectionImg = function () {
      var img = last.getContext('2d').createImageData(last.width, last.height);
      img.data.set(addon.backprojection(h, slot, pixels(last), last.width, last.height));
      return img;
    };
  };

  // Many streams at once: facetrackr's per-frame state machine for n canvases in one call (ht_stream_step).
  // Dispatches the same `facetrackingEvent` the reference sends (src/facetrackr.js:112-125), with `stream` added.
  headtrackr.b200StreamSet = function (n, params) {
    params = params || {};
    addon.streamReset(h, 0, n);
    this.track = function (canvases) {
      var w = canvases[0].width, hgt = canvases[0].height, all = new Uint8Array(n * w * hgt * 4);
      for (var k = 0; k < n; k++) all.set(pixels(canvases[k]), k * w * hgt * 4);
      var ev = addon.streamStep(h, all, n, w, hgt, 5, 1, params.calcAngles ? 1 : 0);
      for (k = 0; k < n; k++) {
        if (ev[k].detection !== 'CS') continue;
        var evt = document.createEvent('Event');
        evt.initEvent('facetrackingEvent', true, true);
        evt.stream = k; evt.height = ev[k].height; evt.width = ev[k].width; evt.angle = ev[k].angle;
        evt.x = ev[k].x; evt.y = ev[k].y; evt.confidence = ev[k].confidence; evt.detection = 'CS'; evt.time = 0;
        document.dispatchEvent(evt);
      }
      return ev;
    };
  };

  // src/whitebalance.js:5
  headtrackr.getWhitebalance = function (canvas) {
    return addon.whitebalance(h, pixels(canvas), canvas.width, canvas.height);
  };
  return h;
};
