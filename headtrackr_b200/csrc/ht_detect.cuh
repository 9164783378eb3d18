// ht_detect.cuh — sm_100a kernels for ccv.grayscale + ccv.detect_objects
// (/root/reference/src/ccv.js:22-32, 109-333).  No tensor cores: byte compares + ordered fp64 adds.
// The whole library is compiled with -fmad=false so that every a*b+c below is two IEEE roundings,
// as in JavaScript.
#pragma once
#include "ht_common.cuh"

namespace ht {

// ------------------------------------------------------------------------------------------------
// K1  grayscale — src/ccv.js:28-29:  gray = ToUint8Clamp(r*0.3 + g*0.59 + b*0.11)  (fp64, RN-even)
// One thread per 4 horizontally adjacent pixels: one 16 B load, one 4 B store into plane 0.
// HBM-bound: 4 B read + 1 B written per pixel.
__device__ __forceinline__ uint32_t gray_of(uint32_t px) {
  const double r = (double)(px & 0xffu), g = (double)((px >> 8) & 0xffu), b = (double)((px >> 16) & 0xffu);
  const double v = __dadd_rn(__dadd_rn(__dmul_rn(r, 0.3), __dmul_rn(g, 0.59)), __dmul_rn(b, 0.11));
  int iv = __double2int_rn(v);  // round half to even == Uint8ClampedArray store
  return (uint32_t)min(iv, 255);
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_gray(const uint8_t *__restrict__ rgba, size_t frame_bytes,
                                              uint8_t *__restrict__ arena, size_t arena_stride,
                                              int w, int h, int pitch0, int quads_per_row) {
  const int frame = blockIdx.y;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t row = t / (uint32_t)quads_per_row;
  if (row >= (uint32_t)h) return;
  const int col = (int)(t - row * (uint32_t)quads_per_row) * 4;
  const uint8_t *src = rgba + (size_t)frame * frame_bytes + ((size_t)row * w + col) * 4;
  uint32_t px[4] = {0, 0, 0, 0};
  if (VEC) {  // w % 4 == 0 and 16 B aligned base
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src));
    px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (col + i < w) px[i] = (uint32_t)src[4 * i] | ((uint32_t)src[4 * i + 1] << 8) | ((uint32_t)src[4 * i + 2] << 16);
  }
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t gv = (col + i < w) ? gray_of(px[i]) : 0u;  // pad columns are written as 0
    out |= gv << (8 * i);
  }
  *reinterpret_cast<uint32_t *>(arena + (size_t)frame * arena_stride + (size_t)row * pitch0 + col) = out;
}

// ------------------------------------------------------------------------------------------------
// K2  pyramid level = canvas-shim drawImage (exact integer bilinear, see oracle/ht_oracle.h and
// src/ccv.js:121,128,135,140,145).  One launch per pyramid "generation" (levels whose sources are
// complete).  Block = 32 x 32 pixels of one destination plane; thread = one column x 4 rows.
// Tap positions / weights come from per-job tables read with 16 B loads (computing them on the fly with magic
// divisions was measured slower: 4.04 vs 3.89 ms per 1024 frames).
__global__ void __launch_bounds__(256) k_resample(DevPlan plan, int tile0, uint8_t *__restrict__ arena,
                                                  size_t arena_stride) {
  // per-block metadata: one 8 B tile record and one 64 B job record, fetched with vector loads
  const uint2 tl = __ldg(reinterpret_cast<const uint2 *>(plan.pyr_tiles + tile0 + blockIdx.x));
  const int job_id = (int)(tl.x & 0xffffu), tx = (int)(tl.x >> 16), ty = (int)(tl.y & 0xffffu);
  const uint4 *jp = reinterpret_cast<const uint4 *>(plan.jobs + job_id);
  const uint4 j0 = __ldg(jp), j1 = __ldg(jp + 1), j2 = __ldg(jp + 2), j3 = __ldg(jp + 3);
  // DevJob: {src_off, dst_off, src_pitch, dst_pitch} {dst_h, dw, dh, col_off} {row_off, magic, shift, half} {..}
  const uint32_t src_off = j0.x, dst_off = j0.y;
  const int src_pitch = (int)j0.z, dst_pitch = (int)j0.w;
  const int dst_h = (int)j1.x, dw = (int)j1.y, dh = (int)j1.z;
  const uint32_t col_off = j1.w, row_off = j2.x, magic = j2.y, shift = j2.z, half = j2.w;
  (void)j3;
  // Block = 32 columns x 32 rows: lane = column (adjacent lanes read adjacent-ish source bytes, so one
  // warp-wide byte load touches one or two 128 B lines), each thread produces 4 consecutive rows and reuses
  // its column taps for all of them.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int X = tx * 32 + lane;
  const int Y0 = ty * 32 + warp * 4;
  if (Y0 >= dst_h || X >= dst_pitch) return;
  uint8_t *frame = arena + (size_t)blockIdx.y * arena_stride;
  uint32_t xa = 0, xb = 0, wx0 = 0, wx1 = 0;
  const bool col_ok = X < dw;
  if (col_ok) {
    const uint2 cx = __ldg(reinterpret_cast<const uint2 *>(plan.taps + col_off + X));   // {a | b<<16, f}
    xa = cx.x & 0xffffu; xb = cx.x >> 16;
    wx1 = cx.y & 0xffffu; wx0 = 2u * (uint32_t)dw - wx1;
  }
  const uint32_t Dy = 2u * (uint32_t)dh;
  const uint8_t *src = frame + src_off;
  // All 16 source bytes of the thread's 4 rows are requested before any arithmetic (one exposed memory latency per
  // thread instead of four); rows below the painted area read row taps {0, 0} and are zeroed afterwards, so the
  // loads need no branches.  Offsets inside a plane fit 32 bits.
  uint32_t oa[4], ob[4], wy1[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int Y = Y0 + r;
    uint2 ry = make_uint2(0u, 0u);
    if (Y < dh) ry = __ldg(reinterpret_cast<const uint2 *>(plan.taps + row_off + Y));   // warp-uniform
    oa[r] = (ry.x & 0xffffu) * (uint32_t)src_pitch;
    ob[r] = (ry.x >> 16) * (uint32_t)src_pitch;
    wy1[r] = ry.y & 0xffffu;
  }
  uint32_t p[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    p[r][0] = src[oa[r] + xa]; p[r][1] = src[oa[r] + xb];
    p[r][2] = src[ob[r] + xa]; p[r][3] = src[ob[r] + xb];
  }
  uint8_t *dst = frame + dst_off + (uint32_t)Y0 * (uint32_t)dst_pitch + (uint32_t)X;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int Y = Y0 + r;
    if (Y >= dst_h) break;
    const uint32_t top = wx0 * p[r][0] + wx1 * p[r][1];   // <= 255 * 2dw
    const uint32_t bot = wx0 * p[r][2] + wx1 * p[r][3];
    const uint32_t num = top * (Dy - wy1[r]) + bot * wy1[r] + half;  // <= 255.5 * 4 dw dh  < 2^32 (checked by the planner)
    uint32_t q = (uint32_t)(((uint64_t)num * magic) >> shift);
    if (!col_ok || Y >= dh) q = 0;                        // unpainted columns / rows and the pitch padding are 0
    dst[(uint32_t)r * (uint32_t)dst_pitch] = (uint8_t)q;
  }
}

// ------------------------------------------------------------------------------------------------
// K3  BBF cascade over all windows of one (frame, scale, tile) — src/ccv.js:178-243.
//
// Feature test: min over the p-points > max over the n-points.  The reference's early-outs
// (src/ccv.js:193-218) break exactly when a running min(p) <= running max(n); since min is
// non-increasing and max non-decreasing this is equivalent to the final comparison.
// Stage sum: sequential fp64 adds of alpha in feature order (bit-exact with the JS).
//
// Windows are evaluated in stage groups; survivors of a group are compacted (ballot + prefix) into
// a shared-memory queue so that later, longer stages run on dense warps.

__constant__ ConstCascade c_casc;

// ---- stages specialised at build time (tools/gen_cascade_code.py) ----
__host__ __device__ constexpr unsigned gen_off(int z, int x, int y) {
  return z == 0 ? (unsigned)(y * TP + x) : z == 1 ? (unsigned)(REGION + TP + 2 * x + 2 * y * TP) : (unsigned)(REGION + 4 * x + 4 * y * TP);
}
// sum += (pmin > nmax) ? alpha[2k+1] : alpha[2k]  with alpha[2k] == -alpha[2k+1] (src/ccv.js:194,219):
// add `a` with its sign bit flipped unless pm > nm.  d = nm - pm is negative exactly when the feature
// fires, so the flip mask is ~d & 0x80000000 — one IADD + one LOP3 instead of a compare and two selects.
__device__ __forceinline__ double bbf_accumulate(double s, unsigned pm, unsigned nm, double a) {
  const int d = (int)nm - (int)pm;
  const int hi = __double2hiint(a) ^ (~d & (int)0x80000000);
  return s + __hiloint2double(hi, __double2loint(a));
}
#define HT_W(z, x, y) ((unsigned)win[gen_off(z, x, y)])
#define HT_MIN2(a, b) __vimin3_u32(a, b, b)      /* VIMNMX3.U32 (plain min() is turned into U16x2 + masks) */
#define HT_MIN3(a, b, c) __vimin3_u32(a, b, c)
#define HT_MAX2(a, b) __vimax3_u32(a, b, b)
#define HT_MAX3(a, b, c) __vimax3_u32(a, b, c)
#define HT_ACC(s, pm, nm, k) bbf_accumulate(s, pm, nm, c_casc.alpha[k])
#define HT_THRESHOLD(j) (c_casc.stage[j].threshold)
#include "cascade_face_gen.inc"
#undef HT_W
#undef HT_MIN2
#undef HT_MIN3
#undef HT_MAX2
#undef HT_MAX3
#undef HT_ACC
#undef HT_THRESHOLD

__device__ __forceinline__ unsigned ldpx(const uint8_t *__restrict__ win, unsigned off) { return win[off]; }

// predicated ld.shared.u8: lanes with p == false issue no shared-memory access and return dflt
__device__ __forceinline__ unsigned lds_u8_if(unsigned saddr, bool p, unsigned dflt) {
  unsigned v = dflt;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.shared.u8 %0, [%1];\n\t}" : "+r"(v) : "r"(saddr), "r"((unsigned)p));
  return v;
}

// one stage for one window per lane; all control flow is warp-uniform (table reads are uniform)
__device__ __forceinline__ bool stage_pass(const uint8_t *__restrict__ win, int j, bool alive, double &sum_out) {
  const int first = c_casc.stage[j].first, last = first + c_casc.stage[j].count;
  double sum = 0.0;
  for (int k = first; k < last; ++k) {
    const unsigned kind = c_casc.np_nn[k];
    const unsigned np = kind & 15u, nn = kind >> 4;
    unsigned pmin = ldpx(win, c_casc.off[k][0]);
    unsigned nmax = ldpx(win, c_casc.off[k][5]);
    if (np > 1) {
      pmin = min(pmin, ldpx(win, c_casc.off[k][1]));
      if (np > 2) {
        pmin = min(pmin, ldpx(win, c_casc.off[k][2]));
        if (np > 3) {
          pmin = min(pmin, ldpx(win, c_casc.off[k][3]));
          if (np > 4) pmin = min(pmin, ldpx(win, c_casc.off[k][4]));
        }
      }
    }
    if (nn > 1) {
      nmax = max(nmax, ldpx(win, c_casc.off[k][6]));
      if (nn > 2) {
        nmax = max(nmax, ldpx(win, c_casc.off[k][7]));
        if (nn > 3) {
          nmax = max(nmax, ldpx(win, c_casc.off[k][8]));
          if (nn > 4) nmax = max(nmax, ldpx(win, c_casc.off[k][9]));
        }
      }
    }
    sum = bbf_accumulate(sum, pmin, nmax, c_casc.alpha[k]);   // src/ccv.js:194,219
  }
  sum_out = sum;
  return alive && !(sum < c_casc.stage[j].threshold);  // src/ccv.js:222
}

template <bool FAST, int MINB>
__global__ void __launch_bounds__(CASCADE_THREADS, MINB) k_cascade(DevPlan plan, const LateFeat *__restrict__ late,
                                                              const void *__restrict__ tmaps, int tma_frame0,
                                                              const uint8_t *__restrict__ arena, size_t arena_stride,
                                                              uint32_t *__restrict__ raw_keys,
                                                              double *__restrict__ raw_conf,
                                                              uint32_t *__restrict__ raw_count, int raw_cap) {
  __shared__ __align__(128) uint8_t tile[2 * REGION];
  __shared__ __align__(8) unsigned long long tma_bar;
  __shared__ uint16_t raw[NWIN];   // [slot][class] survivor cells of the group that just ran
  __shared__ uint16_t cl[NWIN];    // [entry][class] compacted per-bank-class lists
  __shared__ int cnt[8][32];
  __shared__ int len_s[32], pre_s[32];
  __shared__ int maxlen_s, total_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int frame = blockIdx.y;
  const DevCascTile tl = plan.casc_tiles[blockIdx.x];
  const DevScale sc = plan.scales[tl.scale];
  const uint8_t *fr = arena + (size_t)frame * arena_stride;
  const int x0 = tl.tx * TW, y0 = tl.ty * TH;  // quarter-res origin of the tile

  // ---- stage the three levels in shared memory (layout in ht_common.cuh) ----
  // Level 0 (13.8 KB, a plain 2-D box of the plane) is staged by the TMA engine when tensor maps are
  // available: one elected thread issues cp.async.bulk.tensor (3-D map: column, row, frame; out-of-bounds
  // elements are zero-filled) and the CTA waits on an mbarrier after it has scattered levels 1 and 2 itself.
  const bool use_tma = (tmaps != nullptr) && (TP == TILE_FILL_COLS);
  if (use_tma) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&tma_bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();   // nobody may poll the barrier before it is initialised
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)REGION) : "memory");
      const unsigned dst = (unsigned)__cvta_generic_to_shared(tile);
      const unsigned long long map = (unsigned long long)(reinterpret_cast<const uint8_t *>(tmaps) + 128 * (size_t)tl.scale);
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"(dst), "l"(map), "r"(4 * x0), "r"(4 * y0), "r"(tma_frame0 + frame), "r"(bar)
          : "memory");
    }
  } else {
    const DevPlane pl = plan.planes[sc.p0];
    const uint8_t *src = fr + pl.off;
    const int X0 = 4 * x0, Y0 = 4 * y0;
    for (int i = tid; i < TILE_ROWS * (TILE_FILL_COLS / 16); i += CASCADE_THREADS) {
      const int r = i / (TILE_FILL_COLS / 16), c = (i % (TILE_FILL_COLS / 16)) * 16;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (Y0 + r < pl.h && X0 + c < pl.pitch)
        v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)(Y0 + r) * pl.pitch + X0 + c));
      if (TP % 16 == 0) {
        *reinterpret_cast<uint4 *>(tile + r * TP + c) = v;
      } else {   // rows are only 4 B aligned
        uint32_t *d = reinterpret_cast<uint32_t *>(tile + r * TP + c);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    }
  }
  {
    const DevPlane pl = plan.planes[sc.p1];
    const uint8_t *src = fr + pl.off;
    const int X0 = 2 * x0, Y0 = 2 * y0;
    constexpr int QC = (L1_COLS + 3) / 4;
    for (int i = tid; i < L1_ROWS * QC; i += CASCADE_THREADS) {
      const int r = i / QC, c = (i % QC) * 4;
      uint32_t v = 0;
      if (Y0 + r < pl.h && X0 + c < pl.pitch)
        v = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)(Y0 + r) * pl.pitch + X0 + c));
      uint8_t *d = tile + REGION + (2 * r + 1) * TP + 2 * c;
      d[0] = (uint8_t)v; d[2] = (uint8_t)(v >> 8); d[4] = (uint8_t)(v >> 16); d[6] = (uint8_t)(v >> 24);
    }
  }
  {
    constexpr int QC = (L2_COLS + 3) / 4;
    for (int i = tid; i < 4 * L2_ROWS * QC; i += CASCADE_THREADS) {
      const int q = i / (L2_ROWS * QC), rem = i % (L2_ROWS * QC);
      const int r = rem / QC, c = (rem % QC) * 4;
      const DevPlane pl = plan.planes[plan.scales[tl.scale].p2[q]];  // (indexing the register copy `sc` would spill it)
      uint32_t v = 0;
      if (y0 + r < pl.h && x0 + c < pl.pitch)
        v = __ldg(reinterpret_cast<const uint32_t *>(fr + pl.off + (size_t)(y0 + r) * pl.pitch + x0 + c));
      uint8_t *d = tile + REGION + (4 * r + 2 * (q >> 1)) * TP + 4 * c + 2 * (q & 1);
      d[0] = (uint8_t)v; d[4] = (uint8_t)(v >> 8); d[8] = (uint8_t)(v >> 16); d[12] = (uint8_t)(v >> 24);
    }
  }
  if (use_tma) {   // every thread observes the completion of the bulk copy (phase 0 of the barrier)
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&tma_bar);
    unsigned done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar) : "memory");
    }
  }
  __syncthreads();

  // ---- stage groups ----
  // Survivors of a group are kept in 32 per-BANK-CLASS lists: class(window) = (lx + 16*dy) & 31 is the shared-
  // memory bank of the window's base a0 (TP = 160: four tile rows are a multiple of 128 B).  Lane L of every warp
  // only ever evaluates windows of class L, so a warp's 32 pixel loads hit 32 different banks in the compacted
  // groups too (ballot compaction of arbitrary survivors measured ~2.7 wavefronts per load).
  //   raw[slot][class] : one writer per cell (window id or 0xFFFF), slot-major so a warp's stores are contiguous
  //   cl[entry][class] : compacted per-class lists; len_s[class], maxlen_s, pre_s[class] (exclusive scan), total_s
  const int warp = tid >> 5;
  auto decode = [&](int wid, int &lx, int &ly, int &q) {
    lx = wid & (TW - 1); ly = (wid / TW) & (TH - 1); q = wid / (TW * TH);
    return tile + (4 * lx + 2 * (q & 1)) + (4 * ly + 2 * (q >> 1)) * TP;
  };
  auto emit = [&](int lx, int ly, int q, double sum) {  // src/ccv.js:227-234: (window id in reference order, last stage sum)
    const uint32_t key = sc.win_base + (uint32_t)((q * sc.qh + (y0 + ly)) * sc.qw + (x0 + lx));
    const uint32_t pos = atomicAdd(&raw_count[frame], 1u);
    if (pos < (uint32_t)raw_cap) {
      raw_keys[(size_t)frame * raw_cap + pos] = key;
      raw_conf[(size_t)frame * raw_cap + pos] = sum;
    }
  };
  // compaction of raw[0..n_slots) into the class lists; every thread calls it
  auto compact = [&](int n_slots) {
    __syncthreads();
    uint16_t v[8];
    int c = 0;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int slot = warp * 8 + s;
      v[s] = (slot < n_slots) ? raw[slot * 32 + lane] : (uint16_t)0xFFFFu;
      c += (v[s] != 0xFFFFu) ? 1 : 0;
    }
    cnt[warp][lane] = c;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) {
      const int x = cnt[w2][lane];
      off += (w2 < warp) ? x : 0;
      tot += x;
    }
#pragma unroll
    for (int s = 0; s < 8; ++s)
      if (v[s] != 0xFFFFu) cl[(off++) * 32 + lane] = v[s];
    if (warp == 0) {
      len_s[lane] = tot;
      int mx = tot, incl = tot;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      pre_s[lane] = incl - tot;
      if (lane == 31) total_s = incl;
      if (lane == 0) maxlen_s = mx;
    }
    __syncthreads();
  };
  // dense first group: thread (warp, lane) owns lx = lane; iteration `it` -> q = it >> 1, ly = (it & 1) * 8 + warp
  auto run_dense = [&](bool emit_here, auto eval) {
    for (int it = 0; it < NWIN / CASCADE_THREADS; ++it) {
      const int wid = it * CASCADE_THREADS + tid;
      int lx, ly, q;
      const uint8_t *win = decode(wid, lx, ly, q);
      bool alive = (x0 + lx < sc.qw) && (y0 + ly < sc.qh);
      double sum = 0.0;
      alive = eval(win, alive, sum);
      if (emit_here) { if (alive) emit(lx, ly, q, sum); }
      else raw[(warp * 8 + it) * 32 + bank_class(lx, ly, q >> 1)] = alive ? (uint16_t)wid : (uint16_t)0xFFFFu;
    }
    return NWIN / 32;
  };
  // later groups: lane L takes entries warp, warp+8, ... of class L
  auto run_lists = [&](bool emit_here, auto eval) {
    const int ml = maxlen_s, mylen = len_s[lane];
    for (int e = warp; e < ml; e += 8) {
      bool alive = e < mylen;
      const int wid = alive ? cl[e * 32 + lane] : 0;
      int lx, ly, q;
      const uint8_t *win = decode(wid, lx, ly, q);
      double sum = 0.0;
      alive = eval(win, alive, sum);
      if (emit_here) { if (alive) emit(lx, ly, q, sum); }
      else raw[e * 32 + lane] = alive ? (uint16_t)wid : (uint16_t)0xFFFFu;
    }
    return ml;
  };
  auto table_stages = [&](const int jb, const int je) {
    return [=](const uint8_t *win, bool alive, double &sum) {
      for (int j = jb; j < je; ++j) {
        if (!__any_sync(0xffffffffu, alive)) break;
        alive = stage_pass(win, j, alive, sum);
      }
      return alive;
    };
  };
  const int late_first = c_casc.group_first[c_casc.n_groups];
  const bool has_late = late_first < c_casc.n_stages;
  int g = 0;
  if (FAST) {
    // specialised groups {0,1} {2,3} {4,5} {6,7} {8,9}: straight-line code generated from the cascade
    // (cascade_face_gen.inc).  Lane-per-window stays cheaper than warp-per-window while a warp iteration still
    // carries >~2 live windows, i.e. up to stage 9 for this cascade (stage-10 entrants: ~2 per tile).
    static_assert(HT_GEN_STAGES == 6 || HT_GEN_STAGES == 8 || HT_GEN_STAGES == 10, "generated stages come in pairs");
#define HT_GEN_PAIR(A, B)                                                   \
  [&](const uint8_t *win, bool alive, double &sum) {                        \
    alive = alive && gen_stage##A(win, sum);                                \
    if (__any_sync(0xffffffffu, alive)) alive = gen_stage##B(win, sum) && alive; \
    return alive;                                                           \
  }
#define HT_GEN_ONE(A)                                                       \
  [&](const uint8_t *win, bool alive, double &sum) { return alive && gen_stage##A(win, sum); }
#ifndef HT_SPLIT_GROUPS
#define HT_SPLIT_GROUPS 1
#endif
#if HT_SPLIT_GROUPS >= 3   // stage 1 on compacted lists (43 % of the windows survive stage 0): measured SLOWER, 12.0 vs
                           // 11.4 ms per 1024 frames - the extra compaction of ~880 survivors per tile costs more
                           // than the 16 % of shared-memory wavefronts it saves
    compact(run_dense(false, HT_GEN_ONE(0)));
    if (total_s == 0) return;
    compact(run_lists(false, HT_GEN_ONE(1)));
    if (total_s == 0) return;
#else
    compact(run_dense(false, HT_GEN_PAIR(0, 1)));
    if (total_s == 0) return;
#endif
#if HT_SPLIT_GROUPS >= 1   // stages 2 and 3 as separate groups: stage 3 (91 loads) runs on re-compacted lists
    compact(run_lists(false, HT_GEN_ONE(2)));
    if (total_s == 0) return;
    compact(run_lists(false, HT_GEN_ONE(3)));
    if (total_s == 0) return;
#else
    compact(run_lists(false, HT_GEN_PAIR(2, 3)));
    if (total_s == 0) return;
#endif
#if HT_SPLIT_GROUPS >= 2
    compact(run_lists(false, HT_GEN_ONE(4)));
    if (total_s == 0) return;
    compact(run_lists(false, HT_GEN_ONE(5)));
    if (total_s == 0) return;
#else
    compact(run_lists(false, HT_GEN_PAIR(4, 5)));
    if (total_s == 0) return;
#endif
#if HT_GEN_STAGES >= 8
    compact(run_lists(false, HT_GEN_PAIR(6, 7)));
    if (total_s == 0) return;
#endif
#if HT_GEN_STAGES >= 10
    compact(run_lists(false, HT_GEN_PAIR(8, 9)));
    if (total_s == 0) return;
#endif
#undef HT_GEN_PAIR
#undef HT_GEN_ONE
    g = HT_GEN_STAGES / 2;
  }
  for (; g < c_casc.n_groups; ++g) {
    const bool emit_here = (g == c_casc.n_groups - 1) && !has_late;
    auto ev = table_stages(c_casc.group_first[g], c_casc.group_first[g + 1]);
    const int n_slots = (g == 0) ? run_dense(emit_here, ev) : run_lists(emit_here, ev);
    if (emit_here) return;
    compact(n_slots);
    if (total_s == 0) return;
  }
  if (!has_late) return;

  // ---- late stages: one warp per surviving window, one feature per lane, exact integer sums ----
  {
    // flatten the class lists (raw[] is free again after the last compaction)
    {
      const int mylen = len_s[lane], base = pre_s[lane];
      for (int e = warp; e < mylen; e += 8) raw[base + e] = cl[e * 32 + lane];
    }
    __syncthreads();
    const int n_in = total_s;
    for (int w = warp; w < n_in; w += CASCADE_THREADS / 32) {
      const int wid = raw[w];
      int lx, ly, q;
      const uint8_t *win = decode(wid, lx, ly, q);
      bool pass = true;
      for (int j = late_first; j < c_casc.n_stages && pass; ++j) {
        const int first = c_casc.stage[j].first, count = c_casc.stage[j].count;
        long long acc = 0;
        for (int base = 0; base < count; base += 32) {
          const int kk = base + lane;
          if (kk < count) {
            const uint4 *f = reinterpret_cast<const uint4 *>(late + first + kk);
            const uint4 a = __ldg(f), b = __ldg(f + 1);
            const unsigned np = b.z & 0xffu, nn = (b.z >> 8) & 0xffu;
            // branch-free: predicated shared loads (inactive slots keep the neutral element and cost no bank traffic)
            const unsigned wbase = (unsigned)__cvta_generic_to_shared(win);
            unsigned pm = lds_u8_if(wbase + (a.x & 0xffffu), true, 255u);
            unsigned nm = lds_u8_if(wbase + (a.z >> 16), true, 0u);
            pm = __vimin3_u32(pm, lds_u8_if(wbase + (a.x >> 16), np > 1, 255u), lds_u8_if(wbase + (a.y & 0xffffu), np > 2, 255u));
            pm = __vimin3_u32(pm, lds_u8_if(wbase + (a.y >> 16), np > 3, 255u), lds_u8_if(wbase + (a.z & 0xffffu), np > 4, 255u));
            nm = __vimax3_u32(nm, lds_u8_if(wbase + (a.w & 0xffffu), nn > 1, 0u), lds_u8_if(wbase + (a.w >> 16), nn > 2, 0u));
            nm = __vimax3_u32(nm, lds_u8_if(wbase + (b.x & 0xffffu), nn > 3, 0u), lds_u8_if(wbase + (b.x >> 16), nn > 4, 0u));
            const int ai = (int)b.y;
            acc += (pm > nm) ? (long long)ai : -(long long)ai;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const long long thr = c_casc.thr_int[j];
        if (acc == thr) {  // exact tie of the decimal sums: decide with the reference's ordered fp64 adds
          double s;
          pass = stage_pass(win, j, true, s);
        } else {
          pass = acc > thr;
        }
      }
      if (pass) {  // confidence = ordered fp64 sum of the last stage
        double s;
        stage_pass(win, c_casc.n_stages - 1, true, s);
        if (lane == 0) emit(lx, ly, q, s);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4  sort raw detections into the reference's (i,q,y,x) order and group them —
// src/ccv.js:34-107 (array_group), 249-332.  One warp per frame.
//
// array_group's union-find yields the connected components of the symmetric closure of the
// predicate, numbered by smallest member index (src/ccv.js:90-105); any components algorithm
// gives the same partition, so min-label propagation is used.  Sums run in list order per class
// (fp64, order-sensitive) exactly as src/ccv.js:274-289.

__device__ __forceinline__ bool group_pred(const Rect &r1, const Rect &r2) {  // src/ccv.js:252-261
  const double distance = floor(r1.width * 0.25 + 0.5);
  return r2.x <= r1.x + distance && r2.x >= r1.x - distance && r2.y <= r1.y + distance &&
         r2.y >= r1.y - distance && r2.width <= floor(r1.width * 1.5 + 0.5) &&
         floor(r2.width * 1.5 + 0.5) >= r1.width;
}

__global__ void __launch_bounds__(128) k_group(DevPlan plan, int n_frames, const uint32_t *__restrict__ raw_keys,
                                               const double *__restrict__ raw_conf,
                                               const uint32_t *__restrict__ raw_count, int raw_cap,
                                               Rect *__restrict__ sorted, int *__restrict__ labels,
                                               Rect *__restrict__ seq2, int min_neighbors,
                                               Rect *__restrict__ out_rects, int32_t *__restrict__ out_counts, int K,
                                               int32_t *__restrict__ overflow_flag) {
  const int lane = threadIdx.x & 31;
  const int frame = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (frame >= n_frames) return;
  const unsigned FULL = 0xffffffffu;
  const uint32_t n_true = raw_count[frame];
  const int n = (int)min(n_true, (uint32_t)raw_cap);
  if (n_true > (uint32_t)raw_cap && lane == 0) atomicOr(overflow_flag, 1);
  const uint32_t *keys = raw_keys + (size_t)frame * raw_cap;
  const double *conf = raw_conf + (size_t)frame * raw_cap;
  Rect *S = sorted + (size_t)frame * raw_cap;
  int *L = labels + (size_t)frame * raw_cap;
  Rect *S2 = seq2 + (size_t)frame * raw_cap;
  Rect *O = out_rects + (size_t)frame * K;

  // 1. rank sort by window id (ids are unique) and decode to rectangles, src/ccv.js:228-233
  for (int i = lane; i < n; i += 32) {
    const uint32_t key = keys[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (keys[j] < key) ? 1 : 0;
    int s = 0;
    for (int t = 1; t < plan.n_scales; ++t)
      if (plan.scales[t].win_base <= key && plan.scales[t].qw > 0 && plan.scales[t].qh > 0) s = t;
    const DevScale sc = plan.scales[s];
    uint32_t rem = key - sc.win_base;
    const uint32_t per_q = (uint32_t)(sc.qw * sc.qh);
    const uint32_t q = rem / per_q;
    rem -= q * per_q;
    const uint32_t y = rem / (uint32_t)sc.qw, x = rem - y * (uint32_t)sc.qw;
    Rect r;
    r.x = (double)(x * 4 + (q & 1) * 2) * sc.scale_x;
    r.y = (double)(y * 4 + (q >> 1) * 2) * sc.scale_x;  // scale_y == scale_x, src/ccv.js:244-245
    r.width = 24.0 * sc.scale_x;
    r.height = 24.0 * sc.scale_x;
    r.confidence = conf[i];
    r.neighbors = 1;
    r.pad_ = 0;
    S[rank] = r;
  }
  __syncwarp();

  if (!(min_neighbors > 0)) {  // src/ccv.js:249-250: raw list
    for (int i = lane; i < n && i < K; i += 32) O[i] = S[i];
    if (lane == 0) {
      out_counts[frame] = min(n, K);
      if (n > K) atomicOr(overflow_flag, 1);
    }
    return;
  }

  // 2. connected components by min-label propagation + pointer jumping
  for (int i = lane; i < n; i += 32) L[i] = i;
  __syncwarp();
  for (;;) {
    bool changed = false;
    for (int i = lane; i < n; i += 32) {
      const Rect ri = S[i];
      int li = L[i];
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        const Rect rj = S[j];
        if (group_pred(ri, rj) || group_pred(rj, ri)) li = min(li, L[j]);
      }
      li = min(li, L[li]);
      if (li < L[i]) { L[i] = li; changed = true; }
    }
    __syncwarp();
    if (!__any_sync(FULL, changed)) break;
  }
  // flatten: every label points at the component's smallest index
  for (int i = lane; i < n; i += 32) {
    int li = L[i];
    while (L[li] != li) li = L[li];
    L[i] = li;
  }
  __syncwarp();

  // 3. per class (in order of smallest member): ordered sums, src/ccv.js:274-303
  int n2 = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const bool is_root = (i < n) && (L[i] == i);
    Rect c;
    c.x = c.y = c.width = c.height = c.confidence = 0.0;
    c.neighbors = 0; c.pad_ = 0;
    if (is_root) {
      for (int j = i; j < n; ++j) {
        if (L[j] != i) continue;
        const Rect r1 = S[j];
        if (c.neighbors == 0) c.confidence = r1.confidence;
        ++c.neighbors;
        c.x += r1.x; c.y += r1.y; c.width += r1.width; c.height += r1.height;
        c.confidence = fmax(c.confidence, r1.confidence);
      }
    }
    const bool keep = is_root && c.neighbors >= min_neighbors;
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) {
      const double nn = (double)c.neighbors;
      Rect r;
      r.x = (c.x * 2 + nn) / (2 * nn);
      r.y = (c.y * 2 + nn) / (2 * nn);
      r.width = (c.width * 2 + nn) / (2 * nn);
      r.height = (c.height * 2 + nn) / (2 * nn);
      r.neighbors = c.neighbors;
      r.confidence = c.confidence;
      r.pad_ = 0;
      S2[n2 + __popc(m & ((1u << lane) - 1u))] = r;
    }
    n2 += __popc(m);
  }
  __syncwarp();

  // 4. drop rectangles contained in a better one, src/ccv.js:307-330
  int n_out = 0;
  for (int base = 0; base < n2; base += 32) {
    const int i = base + lane;
    bool flag = i < n2;
    Rect r1;
    if (flag) {
      r1 = S2[i];
      for (int j = 0; j < n2; ++j) {
        const Rect r2 = S2[j];
        const double distance = floor(r2.width * 0.25 + 0.5);
        if (i != j && r1.x >= r2.x - distance && r1.y >= r2.y - distance &&
            r1.x + r1.width <= r2.x + r2.width + distance && r1.y + r1.height <= r2.y + r2.height + distance &&
            (r2.neighbors > max(3, r1.neighbors) || r1.neighbors < 3)) {
          flag = false;
          break;
        }
      }
    }
    const unsigned m = __ballot_sync(FULL, flag);
    if (flag) {
      const int pos = n_out + __popc(m & ((1u << lane) - 1u));
      if (pos < K) O[pos] = r1;
    }
    n_out += __popc(m);
  }
  if (lane == 0) {
    out_counts[frame] = min(n_out, K);
    if (n_out > K) atomicOr(overflow_flag, 1);
  }
}

}  // namespace ht
