// js/addon.cc — Node N-API binding of libheadtrackr_b200.so (include/headtrackr_b200.h).
//
// UNTESTED SOURCE: this image (and the GPU box) has no node, no node_api.h and no node-gyp, so this file
// cannot be compiled here.  It is the thin binding a maintainer adds; all logic lives
// below the C ABI, which IS tested (tests/test_gpu_*.py through ctypes).  Build: `node-gyp rebuild` with
// js/binding.gyp on a machine with Node >= 12 and the CUDA runtime.
//
// Exposed to JavaScript (1:1 with the C ABI):
//   create(cascadeBlob: Buffer, {device,maxWidth,maxHeight,maxFrames}) -> handle (External)
//   detect(handle, rgba: Uint8ClampedArray|Buffer, n, w, h, interval, minNeighbors) -> Array<Array<rect>>
//   trackInit(handle, slot, rgba, w, h, x, y, rw, rh, calcAngles)
//   track(handle, slot, rgba, w, h, nCalls) -> {x,y,width,height,angle, window:{x,y,width,height}}
//   whitebalance(handle, rgba, w, h) -> Number
//   backprojection(handle, slot, rgba, w, h) -> Uint8ClampedArray
//   streamStep(handle, rgba /* n frames, one per stream */, n, w, h, interval, minNeighbors, calcAngles)
//        -> Array<{detection:"VJ"|"CS", x,y,width,height,angle,confidence, found, lost}>   (ht_stream_step)
//   streamReset(handle, first, n)
//   streamHeadConfig(handle, {smoothing, fov, cameraOffset, headPosition} | null)     (ht_stream_head_config)
//   streamStepHead(handle, rgba, n, w, h, interval, minNeighbors, calcAngles)
//        -> {events: [...as streamStep...], heads: Array<{valid, found, x, y, z}>}    (ht_stream_step_head)
//   ingest(handle, rgba /* n video frames */, n, sw, sh, dw, dh) -> Uint8ClampedArray (n canvases)   (ht_ingest)
//   destroy(handle)
#include <node_api.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/headtrackr_b200.h"

#define NAPI_OK(call)                                             \
  do {                                                            \
    if ((call) != napi_ok) {                                      \
      napi_throw_error(env, nullptr, "N-API call failed: " #call); \
      return nullptr;                                             \
    }                                                             \
  } while (0)

static napi_value Throw(napi_env env, ht_ctx *ctx, int rc) {
  const char *msg = ht_last_error(ctx);
  napi_throw_error(env, nullptr, (msg && *msg) ? msg : (rc == HT_ERR_CUDA ? "CUDA error" : "headtrackr_b200 error"));
  return nullptr;
}

static bool GetBytes(napi_env env, napi_value v, uint8_t **data, size_t *len) {
  bool is_buf = false, is_ta = false;
  napi_is_buffer(env, v, &is_buf);
  if (is_buf) return napi_get_buffer_info(env, v, reinterpret_cast<void **>(data), len) == napi_ok;
  napi_is_typedarray(env, v, &is_ta);
  if (!is_ta) return false;
  napi_typedarray_type t;
  napi_value ab;
  size_t off;
  return napi_get_typedarray_info(env, v, &t, len, reinterpret_cast<void **>(data), &ab, &off) == napi_ok;
}

static void Finalize(napi_env, void *p, void *) { ht_destroy(static_cast<ht_ctx *>(p)); }

static napi_value Create(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  uint8_t *blob;
  size_t blob_len;
  if (!GetBytes(env, argv[0], &blob, &blob_len)) return Throw(env, nullptr, HT_ERR_ARG);
  ht_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.max_width = 1280; cfg.max_height = 720; cfg.max_frames = 16;
  auto geti = [&](const char *k, int32_t *dst) {
    napi_value v; bool has = false;
    if (argc > 1 && napi_has_named_property(env, argv[1], k, &has) == napi_ok && has &&
        napi_get_named_property(env, argv[1], k, &v) == napi_ok) napi_get_value_int32(env, v, dst);
  };
  geti("device", &cfg.device); geti("maxWidth", &cfg.max_width); geti("maxHeight", &cfg.max_height);
  geti("maxFrames", &cfg.max_frames);
  ht_ctx *ctx = nullptr;
  int rc = ht_create(&ctx, &cfg, blob, blob_len);
  if (rc != HT_OK) return Throw(env, nullptr, rc);
  napi_value ext;
  NAPI_OK(napi_create_external(env, ctx, Finalize, nullptr, &ext));
  return ext;
}

static ht_ctx *Ctx(napi_env env, napi_value v) {
  void *p = nullptr;
  napi_get_value_external(env, v, &p);
  return static_cast<ht_ctx *>(p);
}

static napi_value SetNum(napi_env env, napi_value obj, const char *k, double v) {
  napi_value n;
  napi_create_double(env, v, &n);
  napi_set_named_property(env, obj, k, n);
  return obj;
}

// detect(handle, rgba, n, w, h, interval, minNeighbors) -> [[{x,y,width,height,neighbors,confidence}]]
static napi_value Detect(napi_env env, napi_callback_info info) {
  size_t argc = 7;
  napi_value argv[7];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  uint8_t *rgba; size_t len;
  if (!GetBytes(env, argv[1], &rgba, &len)) return Throw(env, ctx, HT_ERR_ARG);
  int32_t n, w, h, interval, mn;
  napi_get_value_int32(env, argv[2], &n); napi_get_value_int32(env, argv[3], &w);
  napi_get_value_int32(env, argv[4], &h); napi_get_value_int32(env, argv[5], &interval);
  napi_get_value_int32(env, argv[6], &mn);
  if (len < (size_t)n * w * h * 4) return Throw(env, ctx, HT_ERR_ARG);
  const int K = ht_max_rects(ctx);
  std::vector<ht_rect> rects((size_t)n * K);
  std::vector<int32_t> counts(n);
  int rc = ht_detect(ctx, rgba, n, w, h, interval, mn, rects.data(), counts.data());
  if (rc < 0) return Throw(env, ctx, rc);
  napi_value out;
  NAPI_OK(napi_create_array_with_length(env, n, &out));
  for (int f = 0; f < n; ++f) {
    napi_value list;
    napi_create_array_with_length(env, counts[f], &list);
    for (int i = 0; i < counts[f]; ++i) {
      const ht_rect &r = rects[(size_t)f * K + i];
      napi_value o;
      napi_create_object(env, &o);
      SetNum(env, o, "x", r.x); SetNum(env, o, "y", r.y); SetNum(env, o, "width", r.width);
      SetNum(env, o, "height", r.height);
      SetNum(env, o, mn > 0 ? "neighbors" : "neighbor", r.neighbors);   // src/ccv.js:232 vs :301
      SetNum(env, o, "confidence", r.confidence);
      napi_set_element(env, list, i, o);
    }
    napi_set_element(env, out, f, list);
  }
  return out;
}

static napi_value TrackInit(napi_env env, napi_callback_info info) {
  size_t argc = 10;
  napi_value argv[10];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  int32_t slot, w, h, r[4], calc;
  uint8_t *rgba; size_t len;
  napi_get_value_int32(env, argv[1], &slot);
  if (!GetBytes(env, argv[2], &rgba, &len)) return Throw(env, ctx, HT_ERR_ARG);
  napi_get_value_int32(env, argv[3], &w); napi_get_value_int32(env, argv[4], &h);
  for (int i = 0; i < 4; ++i) napi_get_value_int32(env, argv[5 + i], &r[i]);
  napi_get_value_int32(env, argv[9], &calc);
  int rc = ht_track_init(ctx, &slot, 1, rgba, w, h, r, calc);
  if (rc < 0) return Throw(env, ctx, rc);
  return nullptr;
}

static napi_value Track(napi_env env, napi_callback_info info) {
  size_t argc = 6;
  napi_value argv[6];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  int32_t slot, w, h, n_calls;
  uint8_t *rgba; size_t len;
  napi_get_value_int32(env, argv[1], &slot);
  if (!GetBytes(env, argv[2], &rgba, &len)) return Throw(env, ctx, HT_ERR_ARG);
  napi_get_value_int32(env, argv[3], &w); napi_get_value_int32(env, argv[4], &h);
  napi_get_value_int32(env, argv[5], &n_calls);
  ht_trackobj o; ht_window win;
  int rc = ht_track(ctx, &slot, 1, rgba, w, h, n_calls, &o, &win);
  if (rc < 0) return Throw(env, ctx, rc);
  napi_value out, wo;
  napi_create_object(env, &out); napi_create_object(env, &wo);
  SetNum(env, out, "x", o.x); SetNum(env, out, "y", o.y); SetNum(env, out, "width", o.width);
  SetNum(env, out, "height", o.height); SetNum(env, out, "angle", o.angle);
  SetNum(env, wo, "x", win.x); SetNum(env, wo, "y", win.y); SetNum(env, wo, "width", win.width);
  SetNum(env, wo, "height", win.height);
  napi_set_named_property(env, out, "window", wo);
  return out;
}

// facetrackr's state machine for n streams in one call (src/facetrackr.js:67-126 + src/main.js:230-244 on the device)
static napi_value StreamStep(napi_env env, napi_callback_info info) {
  size_t argc = 8;
  napi_value argv[8];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  uint8_t *rgba; size_t len;
  int32_t n, w, h, interval, min_neighbors, calc;
  if (!GetBytes(env, argv[1], &rgba, &len)) return Throw(env, ctx, HT_ERR_ARG);
  napi_get_value_int32(env, argv[2], &n); napi_get_value_int32(env, argv[3], &w); napi_get_value_int32(env, argv[4], &h);
  napi_get_value_int32(env, argv[5], &interval); napi_get_value_int32(env, argv[6], &min_neighbors);
  napi_get_value_int32(env, argv[7], &calc);
  if (n <= 0 || len < (size_t)n * w * h * 4) return Throw(env, ctx, HT_ERR_ARG);
  std::vector<ht_stream_event> ev((size_t)n);
  int rc = ht_stream_step(ctx, rgba, n, w, h, interval, min_neighbors, calc, ev.data());
  if (rc < 0) return Throw(env, ctx, rc);
  napi_value out;
  napi_create_array_with_length(env, (size_t)n, &out);
  for (int k = 0; k < n; ++k) {
    napi_value o, s, b;
    napi_create_object(env, &o);
    napi_create_string_utf8(env, ev[k].detection == 2 ? "CS" : "VJ", 2, &s);
    napi_set_named_property(env, o, "detection", s);
    SetNum(env, o, "x", ev[k].x); SetNum(env, o, "y", ev[k].y); SetNum(env, o, "width", ev[k].width);
    SetNum(env, o, "height", ev[k].height); SetNum(env, o, "angle", ev[k].angle); SetNum(env, o, "confidence", ev[k].confidence);
    napi_get_boolean(env, (ev[k].status & 1) != 0, &b); napi_set_named_property(env, o, "found", b);
    napi_get_boolean(env, (ev[k].status & 2) != 0, &b); napi_set_named_property(env, o, "lost", b);
    napi_set_element(env, out, (uint32_t)k, o);
  }
  return out;
}

static bool GetBoolProp(napi_env env, napi_value obj, const char *k, bool dflt) {
  bool has = false; napi_value v; bool out = dflt;
  if (napi_has_named_property(env, obj, k, &has) == napi_ok && has && napi_get_named_property(env, obj, k, &v) == napi_ok) napi_get_value_bool(env, v, &out);
  return out;
}
static double GetNumProp(napi_env env, napi_value obj, const char *k, double dflt) {
  bool has = false; napi_value v; double out = dflt;
  if (napi_has_named_property(env, obj, k, &has) == napi_ok && has && napi_get_named_property(env, obj, k, &v) == napi_ok) napi_get_value_double(env, v, &out);
  return out;
}

// headtrackr.Tracker's {smoothing, fov, cameraOffset, headPosition} (src/main.js:35-56) for the GPU head epilogue
static napi_value StreamHeadConfig(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  napi_valuetype t;
  napi_typeof(env, argv[1], &t);
  int rc;
  if (t != napi_object) rc = ht_stream_head_config(ctx, nullptr);
  else {
    ht_head_params p;
    std::memset(&p, 0, sizeof(p));
    p.smoothing = GetBoolProp(env, argv[1], "smoothing", true);
    p.head_position = GetBoolProp(env, argv[1], "headPosition", true);
    p.edgecorrection = 1;
    p.alpha = 0.35;                                              // src/main.js:163
    p.fov_deg = GetNumProp(env, argv[1], "fov", 0.0);            // <= 0: estimate (src/main.js:283-288)
    p.camera_offset = GetNumProp(env, argv[1], "cameraOffset", 11.5);
    p.distance_to_screen = 60.0;
    rc = ht_stream_head_config(ctx, &p);
  }
  if (rc < 0) return Throw(env, ctx, rc);
  return nullptr;
}

static napi_value StreamStepHead(napi_env env, napi_callback_info info) {
  size_t argc = 8;
  napi_value argv[8];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  uint8_t *rgba; size_t len;
  int32_t n, w, h, interval, min_neighbors, calc;
  if (!GetBytes(env, argv[1], &rgba, &len)) return Throw(env, ctx, HT_ERR_ARG);
  napi_get_value_int32(env, argv[2], &n); napi_get_value_int32(env, argv[3], &w); napi_get_value_int32(env, argv[4], &h);
  napi_get_value_int32(env, argv[5], &interval); napi_get_value_int32(env, argv[6], &min_neighbors);
  napi_get_value_int32(env, argv[7], &calc);
  if (n <= 0 || len < (size_t)n * w * h * 4) return Throw(env, ctx, HT_ERR_ARG);
  std::vector<ht_stream_event> ev((size_t)n);
  std::vector<ht_head_event> he((size_t)n);
  int rc = ht_stream_step_head(ctx, rgba, n, w, h, interval, min_neighbors, calc, ev.data(), he.data());
  if (rc < 0) return Throw(env, ctx, rc);
  napi_value out, events, heads;
  napi_create_object(env, &out);
  napi_create_array_with_length(env, (size_t)n, &events);
  napi_create_array_with_length(env, (size_t)n, &heads);
  for (int k = 0; k < n; ++k) {
    napi_value o, s, b;
    napi_create_object(env, &o);
    napi_create_string_utf8(env, ev[k].detection == 2 ? "CS" : "VJ", 2, &s);
    napi_set_named_property(env, o, "detection", s);
    SetNum(env, o, "x", ev[k].x); SetNum(env, o, "y", ev[k].y); SetNum(env, o, "width", ev[k].width);
    SetNum(env, o, "height", ev[k].height); SetNum(env, o, "angle", ev[k].angle); SetNum(env, o, "confidence", ev[k].confidence);
    napi_get_boolean(env, (ev[k].status & 2) != 0, &b); napi_set_named_property(env, o, "lost", b);
    napi_set_element(env, events, (uint32_t)k, o);
    napi_value hobj;
    napi_create_object(env, &hobj);
    napi_get_boolean(env, he[k].valid != 0, &b); napi_set_named_property(env, hobj, "valid", b);
    napi_get_boolean(env, (he[k].status & 1) != 0, &b); napi_set_named_property(env, hobj, "found", b);
    SetNum(env, hobj, "x", he[k].x); SetNum(env, hobj, "y", he[k].y); SetNum(env, hobj, "z", he[k].z);
    napi_set_element(env, heads, (uint32_t)k, hobj);
  }
  napi_set_named_property(env, out, "events", events);
  napi_set_named_property(env, out, "heads", heads);
  return out;
}

// canvasContext.drawImage(video, 0, 0, dw, dh) for n frames (src/main.js:170)
static napi_value Ingest(napi_env env, napi_callback_info info) {
  size_t argc = 7;
  napi_value argv[7];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  uint8_t *rgba; size_t len;
  int32_t n, sw, sh, dw, dh;
  if (!GetBytes(env, argv[1], &rgba, &len)) return Throw(env, ctx, HT_ERR_ARG);
  napi_get_value_int32(env, argv[2], &n); napi_get_value_int32(env, argv[3], &sw); napi_get_value_int32(env, argv[4], &sh);
  napi_get_value_int32(env, argv[5], &dw); napi_get_value_int32(env, argv[6], &dh);
  if (n <= 0 || len < (size_t)n * sw * sh * 4) return Throw(env, ctx, HT_ERR_ARG);
  void *data; napi_value ab, ta;
  NAPI_OK(napi_create_arraybuffer(env, (size_t)n * dw * dh * 4, &data, &ab));
  int rc = ht_ingest(ctx, rgba, n, sw, sh, static_cast<uint8_t *>(data), dw, dh);
  if (rc < 0) return Throw(env, ctx, rc);
  NAPI_OK(napi_create_typedarray(env, napi_uint8_clamped_array, (size_t)n * dw * dh * 4, ab, 0, &ta));
  return ta;
}

static napi_value StreamReset(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value argv[3];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  int32_t first, n;
  napi_get_value_int32(env, argv[1], &first); napi_get_value_int32(env, argv[2], &n);
  int rc = ht_stream_reset(ctx, first, n);
  if (rc < 0) return Throw(env, ctx, rc);
  return nullptr;
}

static napi_value Whitebalance(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value argv[4];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  uint8_t *rgba; size_t len; int32_t w, h;
  if (!GetBytes(env, argv[1], &rgba, &len)) return Throw(env, ctx, HT_ERR_ARG);
  napi_get_value_int32(env, argv[2], &w); napi_get_value_int32(env, argv[3], &h);
  double v = 0;
  int rc = ht_whitebalance(ctx, rgba, 1, w, h, &v);
  if (rc < 0) return Throw(env, ctx, rc);
  napi_value out;
  napi_create_double(env, v, &out);
  return out;
}

static napi_value Backprojection(napi_env env, napi_callback_info info) {
  size_t argc = 5;
  napi_value argv[5];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  ht_ctx *ctx = Ctx(env, argv[0]);
  int32_t slot, w, h; uint8_t *rgba; size_t len;
  napi_get_value_int32(env, argv[1], &slot);
  if (!GetBytes(env, argv[2], &rgba, &len)) return Throw(env, ctx, HT_ERR_ARG);
  napi_get_value_int32(env, argv[3], &w); napi_get_value_int32(env, argv[4], &h);
  void *data; napi_value ab, ta;
  NAPI_OK(napi_create_arraybuffer(env, (size_t)w * h * 4, &data, &ab));
  int rc = ht_backprojection(ctx, slot, rgba, w, h, static_cast<uint8_t *>(data));
  if (rc < 0) return Throw(env, ctx, rc);
  NAPI_OK(napi_create_typedarray(env, napi_uint8_clamped_array, (size_t)w * h * 4, ab, 0, &ta));
  return ta;
}

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor d[] = {
      {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"detect", nullptr, Detect, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"trackInit", nullptr, TrackInit, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"track", nullptr, Track, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"streamStep", nullptr, StreamStep, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"streamReset", nullptr, StreamReset, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"streamHeadConfig", nullptr, StreamHeadConfig, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"streamStepHead", nullptr, StreamStepHead, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"ingest", nullptr, Ingest, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"whitebalance", nullptr, Whitebalance, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"backprojection", nullptr, Backprojection, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof(d) / sizeof(d[0]), d);
  return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
