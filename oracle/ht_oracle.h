/*
 * ht_oracle.h — CPU oracle for the headtrackr detect+track hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under headtrackr_b200/ (the product) may include, link
 * or call this.  Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
 * --impl reference legs.
 *
 * What it is: a plain-C restatement of the reference's JavaScript arithmetic
 *   /root/reference/src/ccv.js        (grayscale, pyramid geometry, BBF cascade, grouping)
 *   /root/reference/src/camshift.js   (histogram, weights, back-projection, moments, mean-shift, camShift)
 *   /root/reference/src/whitebalance.js
 * with the same loop nesting and the same IEEE-754 binary64 operation order (JS Number == C double;
 * compiled with -ffp-contract=off).  Each function cites the reference lines it follows.
 *
 * PARITY STATUS: the reference ships no tests, no golden vectors and no fixtures
 * (SURVEY.md §4, §8c).  The oracle is pinned instead against the reference's OWN source executed
 * in this container by oracle/jsmini (a small ES5 interpreter) over the canvas shim below; the
 * resulting vectors are committed under tests/golden/ (see tools/make_goldens.py).  One boundary
 * stays unpinned by construction: the browser's canvas 2D resampler used to build the pyramid
 * (src/ccv.js:121,128,135,140,145) is not part of the reference source.  It is DEFINED here
 * ("canvas shim") and used identically by jsmini's canvas, by this oracle and by the CUDA path.
 *
 * Canvas shim (the single non-reference-defined decision):
 *   - new canvas = all zero (transparent black); alpha of inputs assumed 255;
 *   - getImageData outside the canvas reads 0,0,0,0;
 *   - drawImage(src, sx,sy,sw,sh, 0,0,dw,dh): dw<=0 or dh<=0 paints nothing; otherwise, for each
 *     destination pixel (X,Y), exact bilinear interpolation at the pixel-centre mapping
 *         u = sx + (X+1/2)*sw/dw - 1/2,   v = sy + (Y+1/2)*sh/dh - 1/2
 *     with the 2x2 taps clamped to the source rectangle, evaluated in exact integer arithmetic
 *     and rounded half up:
 *         un = (2X+1)*sw - dw;  x0 = floor(un / (2dw));  fx = un - x0*2dw   (0 <= fx < 2dw)
 *         vn = (2Y+1)*sh - dh;  y0 = floor(vn / (2dh));  fy = vn - y0*2dh
 *         num = (2dw-fx)(2dh-fy) S[y0][x0] + fx(2dh-fy) S[y0][x0+1] + (2dw-fx)fy S[y0+1][x0] + fx fy S[y0+1][x0+1]
 *         D[Y][X] = floor((num + 2 dw dh) / (4 dw dh))
 *     (1:1 draws are exact copies; 2:1 draws of even sizes are the rounded 2x2 mean).
 */
#ifndef HT_ORACLE_H
#define HT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double x, y, width, height, confidence;
  int32_t neighbors;
  int32_t pad_;
} hto_rect;

/* Pyramid geometry, src/ccv.js:110-147.  slot index s in [0, n_slots), copy q in [0,4). */
typedef struct {
  int interval, next, scale_upto, n_slots;
  int w[128], h[128];          /* base (q=0) slot sizes */
} hto_geom;

typedef struct hto_pyramid hto_pyramid;

/* returns 0 ok, <0 error (too many slots / zero-sized level) */
int hto_geometry(int W, int H, int interval, hto_geom *g);

/* src/ccv.js:22-32 : gray[y*w+x] = ToUint8Clamp(r*0.3 + g*0.59 + b*0.11) */
void hto_grayscale(const uint8_t *rgba, int w, int h, uint8_t *gray);

/* canvas-shim drawImage on single-channel planes; dst must be zero-initialised by the caller */
void hto_draw_image(const uint8_t *src, int src_pitch, int sx, int sy, int sw, int sh,
                    uint8_t *dst, int dst_pitch, int dw, int dh);

/* pyramid of gray planes, src/ccv.js:113-147 */
hto_pyramid *hto_pyramid_build(const uint8_t *gray, int W, int H, int interval);
void hto_pyramid_free(hto_pyramid *p);
const hto_geom *hto_pyramid_geom(const hto_pyramid *p);
/* plane pointer (pitch == width), NULL if the slot/copy does not exist */
const uint8_t *hto_pyramid_plane(const hto_pyramid *p, int slot, int q, int *w, int *h);

typedef struct {
  int64_t windows;           /* windows visited */
  int64_t feature_evals;     /* features evaluated */
  int64_t stage_entries[64]; /* windows entering stage j */
  int n_raw;                 /* raw detections before grouping */
} hto_detect_stats;

/* full ccv.detect_objects(ccv.grayscale(canvas), cascade, interval, min_neighbors), src/ccv.js:109-333.
 * blob = HTC1 cascade blob.  Results written to out[0..max_out); returns number of rects (may exceed
 * max_out: then only max_out were written), <0 on error.  raw_out/raw_cap (optional) receive the
 * pre-grouping list in reference order; stats optional. */
int hto_detect(const uint8_t *rgba, int W, int H, const void *blob, size_t blob_len,
               int interval, int min_neighbors, hto_rect *out, int max_out,
               hto_rect *raw_out, int raw_cap, hto_detect_stats *stats);

/* detection on a prebuilt pyramid: returns raw count; raw list in (scale, q, y, x) order */
int hto_cascade_raw(const hto_pyramid *p, const void *blob, size_t blob_len,
                    hto_rect *raw_out, int raw_cap, hto_detect_stats *stats);

/* grouping of a raw list, src/ccv.js:34-107, 249-332 */
int hto_group(const hto_rect *raw, int n_raw, int min_neighbors, hto_rect *out, int max_out);

/* ---- camshift, src/camshift.js ---- */
typedef struct {
  uint32_t model_hist[4096];
  int32_t sx, sy, sw, sh;            /* _searchWindow */
  int32_t tx, ty, tw, th;            /* _trackObj x,y,width,height */
  double angle;                      /* _trackObj.angle */
  int32_t calc_angles;
  int32_t initialised;
} hto_tracker;

typedef struct {
  int32_t n_iter;                    /* number of Moments() evaluated inside the loop (1..10) */
  int32_t converged;                 /* 1 if the break at src/camshift.js:299-301 was taken */
  int32_t wx[11], wy[11];            /* _searchWindow.x/y after each iteration's shift (unclamped) */
  double m00, m10, m01, m11, m20, m02; /* final Moments */
} hto_track_trace;

/* src/camshift.js:49-72 over an RGBA byte array of n pixels */
void hto_histogram(const uint8_t *rgba, size_t n_px, uint32_t *bins4096);
/* src/camshift.js:314-330 */
void hto_weights(const uint32_t *model, const uint32_t *cur, double *w4096);
/* src/camshift.js:198-211; rect may stick out of the canvas (reads 0,0,0,0 there). returns <0 if w<=0||h<=0 */
int hto_tracker_init(hto_tracker *t, const uint8_t *rgba, int W, int H,
                     int x, int y, int w, int h, int calc_angles);
/* src/camshift.js:213-312 ; one track() call */
int hto_tracker_track(hto_tracker *t, const uint8_t *rgba, int W, int H, hto_track_trace *trace);
/* src/camshift.js:177-196 : floor(255*w[bin]) gray image (debug getBackProjectionImg) as RGBA */
void hto_backprojection_img(const hto_tracker *t, const uint8_t *rgba, int W, int H, uint8_t *out_rgba);

/* detect -> VJ->CS hand-off (src/facetrackr.js:97-108,157-165) -> n_calls x track() in one call (CPU baseline) */
int hto_detect_track(const uint8_t *rgba, int W, int H, const void *blob, size_t blob_len, int interval,
                     int min_neighbors, int calc_angles, int n_calls, hto_tracker *t_out, int *found);

/* src/whitebalance.js:5-29 */
double hto_whitebalance(const uint8_t *rgba, int W, int H);

#ifdef __cplusplus
}
#endif
#endif
