/*
 * headtrackr_b200.h — C ABI of libheadtrackr_b200.so: the B200-native (sm_100a) replacement for
 * headtrackr's per-frame detect-then-track pixel kernels.
 *
 * This is the drop-in boundary.  Every entry point replaces a call that
 * headtrackr's facetrackr.js makes into ccv.js / camshift.js / whitebalance.js; a Node N-API addon
 * (js/addon.cc) or any FFI (ctypes: headtrackr_b200/_lib.py) binds these 1:1.  Plain pointers and
 * sizes only; no C++/torch types.  All citations are into /root/reference/.
 *
 *   reference call (file:line)                                       ->  C ABI
 *   ---------------------------------------------------------------      ---------------------------
 *   headtrackr.ccv.detect_objects(headtrackr.ccv.grayscale(canvas),
 *       headtrackr.cascade, 5, 1)          src/facetrackr.js:147-149 ->  ht_detect
 *       (= src/ccv.js:22-32 grayscale + src/ccv.js:109-333 detect_objects)
 *   headtrackr.cascade                     src/cascade.js:19         ->  cascade blob given to ht_create
 *   new headtrackr.camshift.Tracker({calcAngles})
 *                                          src/facetrackr.js:64      ->  tracker slots inside ht_ctx
 *   cstracker.initTracker(canvas, Rectangle)
 *                                          src/facetrackr.js:101-107 ->  ht_track_init / ht_track_init_from_detect
 *       (= src/camshift.js:198-211)
 *   cstracker.track(canvas); getTrackObj() src/facetrackr.js:190-191 ->  ht_track
 *       (= src/camshift.js:213-312, 167-170)
 *   cstracker.getSearchWindow()            src/camshift.js:162-165   ->  ht_track (out_windows)
 *   cstracker.getBackProjectionImg()       src/facetrackr.js:195     ->  ht_backprojection
 *   headtrackr.getWhitebalance(canvas)     src/facetrackr.js:223     ->  ht_whitebalance
 *       (= src/whitebalance.js:5-29)
 *
 * Conventions
 *   - A "canvas" is a tightly packed RGBA8 frame: w*h*4 bytes, row-major (what getImageData returns).
 *     Frame batches are n contiguous frames.  `rgba` and all out pointers may be HOST or DEVICE
 *     pointers (resolved with cudaPointerGetAttributes).  With device outputs the call only enqueues
 *     work on the context's stream (use ht_sync); with host outputs it returns after the results
 *     have landed.
 *   - Return value: 0 = ok; >0 = completed with a warning (HT_WARN_*); <0 = error (HT_ERR_*).
 *     ht_last_error(ctx) describes the last non-zero return.  Nothing throws across the ABI.
 *   - One context per host thread / GPU; calls on one context must be serialised by the caller
 *     (the reference is single-threaded, src/main.js:303).
 *   - There is NO CPU fallback: every entry point fails with HT_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef HEADTRACKR_B200_H
#define HEADTRACKR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HT_OK 0
#define HT_WARN_OVERFLOW 1      /* a per-frame raw or result list hit its capacity; lists were truncated */
#define HT_ERR_ARG (-1)
#define HT_ERR_CUDA (-2)
#define HT_ERR_SIZE (-3)        /* frame too small/large for the pyramid (a browser would throw on a 0-sized level) */
#define HT_ERR_CASCADE (-4)
#define HT_ERR_STATE (-5)       /* e.g. ht_track on a slot that was never initialised */

typedef struct ht_ctx ht_ctx;

/* One element of the array detect_objects returns (src/ccv.js:228-233 raw, :297-302 grouped):
 * x,y = top-left, all Numbers (fp64).  For raw lists (min_neighbors <= 0) neighbors == 1. */
typedef struct {
  double x, y, width, height, confidence;
  int32_t neighbors;
  int32_t pad_;
} ht_rect;

/* camshift TrackObj (src/camshift.js:362-377): x,y = centre. */
typedef struct {
  int32_t x, y, width, height;
  double angle;
} ht_trackobj;

/* camshift _searchWindow (src/camshift.js:156) */
typedef struct {
  int32_t x, y, width, height;
} ht_window;

typedef struct {
  int32_t device;             /* CUDA device ordinal */
  int32_t max_width;          /* largest frame the context must handle */
  int32_t max_height;
  int32_t max_frames;         /* largest batch per call == number of tracker slots */
  int32_t max_raw_per_frame;  /* capacity of the pre-grouping list per frame (0 -> 1024) */
  int32_t max_rects_per_frame;/* K: capacity of the result list per frame (0 -> 64) */
  void *cuda_stream;          /* cudaStream_t to run on; NULL -> the context creates its own */
} ht_config;

/* Library/ABI version (major<<16 | minor). */
uint32_t ht_version(void);

/* cascade_blob: "HTC1" blob (tools/pack_cascade.py) of headtrackr.cascade (src/cascade.js:19). */
int ht_create(ht_ctx **out, const ht_config *cfg, const void *cascade_blob, size_t blob_len);
void ht_destroy(ht_ctx *ctx);
const char *ht_last_error(const ht_ctx *ctx);   /* ctx may be NULL: error of the last failed ht_create */
int ht_sync(ht_ctx *ctx);
int ht_max_rects(const ht_ctx *ctx);            /* K */

/* ccv.detect_objects(ccv.grayscale(frame), cascade, interval, min_neighbors) for n frames.
 *   out_rects : [n][K] ht_rect, reference order (src/ccv.js:293-330; raw order (scale,q,y,x) if min_neighbors<=0)
 *   out_counts: [n]    number of rects written for each frame
 * The input frames are not modified (the reference works on a copy, src/facetrackr.js:140-145). */
int ht_detect(ht_ctx *ctx, const uint8_t *rgba, int n, int w, int h, int interval, int min_neighbors,
              ht_rect *out_rects, int32_t *out_counts);

/* camshift.Tracker.initTracker(frame, Rectangle(x,y,w,h)) for n tracker slots (src/camshift.js:198-211).
 *   slots : [n] DISTINCT slot ids in [0,max_frames) (NULL -> 0..n-1); rgba: n frames; rects: [n][4] = x,y,w,h
 *           (host arrays are checked: out-of-range or repeated ids -> HT_ERR_ARG; a device-resident slot array is
 *           used as is - ids out of range or repeated are undefined behaviour, like any bad device pointer)
 * Pixels of the rectangle outside the frame count as (0,0,0) like canvas getImageData. */
int ht_track_init(ht_ctx *ctx, const int32_t *slots, int n, const uint8_t *rgba, int w, int h,
                  const int32_t *rects, int calc_angles);

/* The VJ->CS hand-off of facetrackr (src/facetrackr.js:157-165 first-max-confidence pick, :97-108
 * confidence > -10 gate and Math.floor of x,y,w,h) done on the device from ht_detect's outputs:
 *   det_rects [n][K], det_counts [n] (host or device).  out_found [n] (optional): 1 if the slot was seeded. */
int ht_track_init_from_detect(ht_ctx *ctx, const int32_t *slots, int n, const uint8_t *rgba, int w, int h,
                              const ht_rect *det_rects, const int32_t *det_counts, int calc_angles,
                              int32_t *out_found);

/* n_calls successive camshift.Tracker.track(frame) calls on each slot's frame, state carried
 * (src/camshift.js:213-312).  out_objs [n] = getTrackObj() after the last call; out_windows [n]
 * (optional) = getSearchWindow().  Slots that were never initialised yield HT_ERR_STATE. */
int ht_track(ht_ctx *ctx, const int32_t *slots, int n, const uint8_t *rgba, int w, int h, int n_calls,
             ht_trackobj *out_objs, ht_window *out_windows);

/* One facetrackr VJ frame followed by its CS frames, for a batch (src/facetrackr.js:67-126):
 *   ht_detect -> first-max-confidence pick, confidence > -10 gate, floor -> initTracker on slot k for frame k
 *   -> n_calls x track().  Frames with no usable face leave out_found[k] = 0 and a zero TrackObj.
 * With HOST frames the upload is pipelined in chunks against the kernels of the previous chunk.
 * out_found and out_windows are optional. */
int ht_detect_track(ht_ctx *ctx, const uint8_t *rgba, int n, int w, int h, int interval, int min_neighbors,
                    int calc_angles, int n_calls, ht_rect *out_rects, int32_t *out_counts, int32_t *out_found,
                    ht_trackobj *out_objs, ht_window *out_windows);

/* facetrackr's per-frame state machine for n independent video streams, on the device.
 * Stream k (tracker slot k) is in one of the reference's detection modes (src/facetrackr.js:57,75-81; whitebalancing
 * off, as after src/main.js:236):
 *   "VJ": ccv.detect_objects on the frame, first-max-confidence pick (src/facetrackr.js:157-165); if its confidence
 *         exceeds -10 the tracker is seeded with the floored rectangle on this same frame and the stream switches
 *         to "CS" (src/facetrackr.js:97-108)
 *   "CS": one camshift track() on the frame (src/facetrackr.js:178-209); a result with width or height 0 means the
 *         face is lost and the stream starts over in "VJ" on the next frame (src/main.js:230-244, retryDetection)
 * ht_stream_step consumes ONE frame per stream (rgba = n frames, stream-major) and returns the TrackObj facetrackr
 * would hold after track() for each stream; the host emits `facetrackingEvent` for records with detection == 2
 * (src/facetrackr.js:112-125).  Mode switches, the pick and the tracker seeding happen in kernels: the only host
 * traffic per frame is the frame upload (if `rgba` is host memory) and the n event records. */
typedef struct {
  int32_t detection;   /* 1 = "VJ", 2 = "CS" */
  int32_t status;      /* bit 0: face found on this frame (VJ -> CS); bit 1: face lost on this frame (CS -> VJ) */
  double x, y, width, height, angle, confidence;   /* VJ: top-left, fp64 as ccv returns; CS: centre, integers */
} ht_stream_event;
/* Head position per stream and frame as an epilogue of the state machine: what src/main.js does
 * with a "CS" result - headtrackrStatus "found" (src/main.js:246-249), headtrackr.Smoother (src/smoother.js:25-87,
 * including its quirks: sp2 aliases sp, z is NaN, predict() runs with step 0), the wait for six head diagonals within
 * 5 px (src/main.js:262-281), headposition.Tracker with its field-of-view estimate and edge correction
 * (src/headposition.js:35-191) - evaluated by the kernel that already writes the stream events.  One ht_head_event
 * per stream and frame; `valid` marks the frames on which the reference dispatches `headtrackingEvent {x, y, z}`. */
typedef struct {
  int32_t smoothing;           /* params.smoothing      (src/main.js:39, default 1) */
  int32_t head_position;       /* params.headPosition   (src/main.js:55, default 1) */
  int32_t edgecorrection;      /* headposition params   (src/headposition.js:44-48, default 1) */
  int32_t pad_;
  double alpha;                /* Smoother alpha        (src/main.js:163: 0.35) */
  double fov_deg;              /* params.fov in degrees; <= 0: estimate it from the first stable face (src/main.js:283-288) */
  double camera_offset;        /* params.cameraOffset   (src/main.js:53: 11.5) */
  double distance_to_screen;   /* 60 cm                 (src/headposition.js:75-79) */
} ht_head_params;
typedef struct {
  int32_t valid;               /* 1: headtrackingEvent dispatched on this frame */
  int32_t status;              /* bit 0: headtrackrStatus "found" on this frame */
  double x, y, z;              /* head position in cm relative to the screen centre (src/headposition.js:165-188) */
  double fx, fy, fwidth, fheight;   /* the (smoothed) face object it was computed from */
} ht_head_event;
/* params == NULL switches the epilogue off.  Changing parameters does not reset stream state; ht_stream_reset does. */
int ht_stream_head_config(ht_ctx *ctx, const ht_head_params *params);
/* ht_stream_step plus one ht_head_event per stream (out_head may be NULL) */
int ht_stream_step_head(ht_ctx *ctx, const uint8_t *rgba, int n, int w, int h, int interval, int min_neighbors,
                        int calc_angles, ht_stream_event *out_events, ht_head_event *out_head);

/* put streams [first, first+n) back into "VJ" (new facetrackr.Tracker) */
int ht_stream_reset(ht_ctx *ctx, int first, int n);
int ht_stream_step(ht_ctx *ctx, const uint8_t *rgba, int n, int w, int h, int interval, int min_neighbors,
                   int calc_angles, ht_stream_event *out_events);

/* Frame ingest: canvasContext.drawImage(videoElement, 0, 0, canvas.width, canvas.height)
 * (src/main.js:170) for n frames - the video frame (sw x sh) scaled onto the working canvas (dw x dh), all four
 * channels, with the canvas resampler this build defines.  src and dst may be host or device
 * memory; a device dst can be passed straight to ht_detect / ht_track / ht_stream_step.  (The 1:1 copy facetrackr
 * makes before detection, src/facetrackr.js:140-145, needs no call: no entry point modifies its input frames.) */
int ht_ingest(ht_ctx *ctx, const uint8_t *src_rgba, int n, int sw, int sh, uint8_t *dst_rgba, int dw, int dh);

/* getBackProjectionImg() of the last track() state for one slot: RGBA w*h*4, floor(255*weight) gray
 * (src/camshift.js:177-196).  Debug path of the reference (src/facetrackr.js:194-196). */
int ht_backprojection(ht_ctx *ctx, int slot, const uint8_t *rgba, int w, int h, uint8_t *out_rgba);

/* headtrackr.getWhitebalance(frame) for n frames: out[n] doubles (src/whitebalance.js:5-29). */
int ht_whitebalance(ht_ctx *ctx, const uint8_t *rgba, int n, int w, int h, double *out);

/* ---- introspection used by the parity tests (not needed by a caller of the reference API) ---- */

/* pyramid geometry for (w,h,interval): n_slots, scale_upto and the per-slot sizes (src/ccv.js:110-127) */
int ht_plan_info(ht_ctx *ctx, int w, int h, int interval, int32_t *n_slots, int32_t *scale_upto,
                 int32_t *slot_w, int32_t *slot_h, int cap);
/* copy one pyramid plane (slot, q) of frame `frame` of the LAST ht_detect call to host memory (w*h bytes) */
int ht_debug_plane(ht_ctx *ctx, int frame, int slot, int q, uint8_t *out, int cap_bytes, int32_t *w, int32_t *h);
/* raw (pre-grouping) list of frame `frame` of the LAST ht_detect call, reference order */
int ht_debug_raw(ht_ctx *ctx, int frame, ht_rect *out, int cap, int32_t *count);
/* tracker slot state: model histogram (4096 u32, optional) */
int ht_debug_model_hist(ht_ctx *ctx, int slot, uint32_t *out4096);
/* mean-shift counters since the last reset: {moment passes summed on the device, passes redone in strict reference
 * order, window pixels visited, track() calls, passes answered from the per-launch window memo} */
int ht_debug_track_stats(ht_ctx *ctx, uint64_t *out5, int reset);
/* Exactness fallbacks (tests only).  The kernels decide the integer outputs of the reference exactly with cheap
 * arithmetic plus a fallback that reproduces the reference's own operation order when the cheap path is not
 * conclusive.  flags force the fallbacks so that they are exercised (results must not change):
 *   bit 0: every generated cascade-stage decision of the survivor lists is treated as an exact tie and re-decided
 *          with ordered fp64 adds (src/ccv.js:186-222)
 *   bit 1: the same for the late (warp-per-window, exact-integer) stages
 *   bit 2: every mean-shift pass of ht_track re-derives its moments in the reference's strict summation order
 *          (src/camshift.js:90-107) as if a truncation had been ambiguous */
int ht_debug_set_exactness(ht_ctx *ctx, int flags);
/* Window memo of ht_track / ht_detect_track (default on).  The moments of a search window depend only on the frame,
 * the histogram weights and the window, and all three are fixed for the n_calls track() calls of one launch; the
 * kernel therefore keeps the moments of the last 8 windows of a stream and re-uses them when mean-shift comes back
 * to one of them (a converged stream; a stream oscillating between two windows - src/camshift.js:283-306 would
 * re-sum them).  Results are identical either way; enable = 0 re-sums every pass like the reference. */
int ht_set_track_memo(ht_ctx *ctx, int enable);
/* Pipelined batches (default off; HT_PIPELINE=1 in the environment turns it on at ht_create).  With enable != 0 an
 * ht_detect_track call whose frames AND outputs are all device pointers returns as soon as its detection is enqueued
 * and leaves its tracking (hand-off + n_calls x track(), src/facetrackr.js:97-108,190) on a second, higher-priority
 * stream, where it runs under the detection kernels of the NEXT ht_detect_track call: CAMShift is a latency chain per
 * stream that leaves most of the GPU idle, the detector is throughput-bound.  Results are identical to the
 * unpipelined call.  The caller's side of the contract: out_found / out_objs / out_windows of call s are complete
 * after ht_sync or ht_join (or any other entry point of the context, which all join first) - NOT merely after the
 * next ht_detect_track; out_rects / out_counts may be reused by the next call (the library orders the accesses);
 * the frames of call s must stay unchanged until then as well. */
int ht_set_pipeline(ht_ctx *ctx, int enable);
/* Stream-level join: later work on the context's stream waits for a pipelined call's tracking.  No host wait. */
int ht_join(ht_ctx *ctx);
/* per-stream timeline of the last ht_track / ht_detect_track launch, 4 x u64 per stream: {globaltimer ns at start,
 * at end, SM id of the leading CTA, moment passes}.  Only for contexts created with HT_TRACK_TRACE=1 in the
 * environment (tools/track_timeline.py); HT_ERR_ARG otherwise. */
int ht_debug_track_trace(ht_ctx *ctx, uint64_t *out, int n_streams);
/* profiling builds (-DHT_TRACK_PASSTRACE=1) only, zeros otherwise: 8 x u64 per stream, the SM clock cycles the leading
 * thread spent in each phase of its passes {pixel loop, warp sums + CTA barrier, exchange + cluster barrier, scalar
 * mean-shift step, publishing barrier, 0, 0, 0}, summed over the launch. */
int ht_debug_track_phases(ht_ctx *ctx, uint64_t *out, int n_streams);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t ht_launch_count(const ht_ctx *ctx);

/* Per-kernel-class device time, measured with CUDA events recorded on the context's stream around
 * every launch while enabled (bench.py's roofline uses the cascade class).  ht_profile_read syncs the
 * stream and returns the accumulated milliseconds and launch counts per class. */
enum { HT_PROF_GRAY = 0, HT_PROF_PYRAMID, HT_PROF_CASCADE, HT_PROF_GROUP, HT_PROF_HIST, HT_PROF_TRACK_INIT,
       HT_PROF_TRACK, HT_PROF_N };
int ht_profile(ht_ctx *ctx, int enable);
int ht_profile_read(ht_ctx *ctx, double *ms_out, uint64_t *launches_out, int reset);

#ifdef __cplusplus
}
#endif
#endif
