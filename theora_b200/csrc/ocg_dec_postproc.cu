/* Decoder post-processing on the device: the reference's out-of-loop de-blocking and de-ringing filters
 * (lib/decode.c:1609-1957; TH_DECCTL_SET_PPLEVEL 2..7), applied to the frame the reconstruction kernels
 * left in HBM, before it is copied to the host.  Non-normative, but th_decode_ycbcr_out hands out its
 * result, so it is restated bit for bit.
 *
 * The reference runs the filters fragment row by fragment row inside the decode loop; their results do
 * not depend on that striping, only on the order inside a plane, which is restated here as:
 *   ocg_pp_hedge_kernel   oc_filter_hedge for every horizontal block edge (decode.c:1610-1660): reads the
 *                         reconstructed frame, writes the post-processing frame; rows 0-3 and the last 4
 *                         rows are copied (decode.c:1739-1743, 1770-1774).  Order-free.
 *   ocg_pp_vedge_kernel   oc_filter_vedge (decode.c:1663-1699), in place on the post-processing frame.
 *                         Along a pixel row the edges form a chain: an edge's first sample (x-5) is the last
 *                         pixel the previous edge may have rewritten.  Whether it did is one bit, and both
 *                         candidates for that sample are known from untouched pixels, so every lane
 *                         evaluates its filter condition for both and the bits of a 32-edge chunk are
 *                         resolved by a 32-step walk over two ballots.
 *                         Both kernels accumulate the per-block "variances" (sums of min(255, sum0/sum1)).
 *   ocg_pp_dering_kernel  oc_dec_dering_frag_rows + oc_dering_block (decode.c:1788-1957): in place, blocks
 *                         in raster order, a block reading the already filtered pixels of its left and
 *                         upper neighbours and -- inside the block -- of the pixel to the left and above.
 *                         One lane per block ROW, lane l one block behind lane l-1 (a wave-front in lock
 *                         step inside the warp); warps of a plane share one CTA and hand over through a
 *                         progress counter in shared memory.
 */
#include <algorithm>
#include "ocg_internal.h"

namespace {

struct PpPlane {
  int32_t W, H, nh, nv, froffset;
  int32_t src_off;    /* reconstructed plane: bottom-left pixel relative to the buffer's luma base */
  int32_t src_stride; /* negative */
  int32_t dst_off;    /* post-processing plane: row 0 (the bottom row, last in memory); row stride is -W */
};

struct PpArgs {
  PpPlane p[3];
  int32_t dc_scale[64];
  int32_t sharp_mod[64];
  int32_t level;
};

__device__ __forceinline__ int pp_abs(int v) { return v < 0 ? -v : v; }

/* the eight outputs of the low-pass filter both edge filters share (decode.c:1638-1650, 1687-1696) */
__device__ __forceinline__ void pp_lowpass(const int (&r)[10], int (&o)[8]) {
  o[0] = (r[0] * 3 + r[1] * 2 + r[2] + r[3] + r[4] + 4) >> 3;
  o[1] = (r[0] * 2 + r[1] + r[2] * 2 + r[3] + r[4] + r[5] + 4) >> 3;
#pragma unroll
  for (int k = 0; k < 4; k++) o[2 + k] = (r[k] + r[k + 1] + r[k + 2] + r[k + 3] * 2 + r[k + 4] + r[k + 5] + r[k + 6] + 4) >> 3;
  o[6] = (r[4] + r[5] + r[6] + r[7] * 2 + r[8] + r[9] * 2 + 4) >> 3;
  o[7] = (r[5] + r[6] + r[7] + r[8] * 2 + r[9] * 3 + 4) >> 3;
}

/* ---- horizontal block edges + the copied rows -------------------------------------------------------
   One thread per (4 pixel columns, edge e): e = 0 copies rows 0..3, e = nv copies the last four rows,
   0 < e < nv filters the edge between block rows e-1 and e. */
__global__ void __launch_bounds__(256)
ocg_pp_hedge_kernel(const PpArgs A, const uint8_t *__restrict__ src_base, uint8_t *__restrict__ pp, const uint8_t *__restrict__ dc_qis,
                    int32_t *__restrict__ variances, int pli0, int pli1) {
  int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  for (int pli = pli0; pli <= pli1; pli++) {
    const PpPlane &P = A.p[pli];
    const int gw = P.W >> 2, n = gw * (P.nv + 1);
    if (t >= n) { t -= n; continue; }
    const int e = t / gw, x = (t - e * gw) * 4;
    const uint8_t *src = src_base + P.src_off + x;
    uint8_t *dst = pp + P.dst_off + x;
    if (e == 0 || e == P.nv) {
      const int y0 = e == 0 ? 0 : P.H - 4;
#pragma unroll
      for (int k = 0; k < 4; k++)
        *(uint32_t *)(dst - (ptrdiff_t)(y0 + k) * P.W) = *(const uint32_t *)(src + (ptrdiff_t)(y0 + k) * P.src_stride);
      return;
    }
    uint32_t w[10], ow[8];
#pragma unroll
    for (int k = 0; k < 10; k++) w[k] = *(const uint32_t *)(src + (ptrdiff_t)(8 * e - 5 + k) * P.src_stride);
    const int bx = x >> 3;
    const int qstep = A.dc_scale[dc_qis[P.froffset + (e - 1) * P.nh + bx]];
    const int flimit = (qstep * 3) >> 2;
    int v0 = 0, v1 = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) ow[k] = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      int r[10], o[8];
#pragma unroll
      for (int k = 0; k < 10; k++) r[k] = (int)((w[k] >> (8 * c)) & 0xFFu);
      int sum0 = 0, sum1 = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) { sum0 += pp_abs(r[k + 1] - r[k]); sum1 += pp_abs(r[k + 5] - r[k + 6]); }
      v0 += min(255, sum0);
      v1 += min(255, sum1);
      if (sum0 < flimit && sum1 < flimit && r[5] - r[4] < qstep && r[4] - r[5] < qstep) pp_lowpass(r, o);
      else {
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = r[k + 1];
      }
#pragma unroll
      for (int k = 0; k < 8; k++) ow[k] |= (uint32_t)(o[k] & 0xFF) << (8 * c);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) *(uint32_t *)(dst - (ptrdiff_t)(8 * e - 4 + k) * P.W) = ow[k];
    atomicAdd(variances + P.froffset + (e - 1) * P.nh + bx, v0);
    atomicAdd(variances + P.froffset + e * P.nh + bx, v1);
    return;
  }
}

/* ---- vertical block edges ---------------------------------------------------------------------------
   One warp per pixel row; a lane per edge, 32 edges per round. */
struct PpEdge {
  uint32_t w0, w1, w2, w3; /* the aligned words at x-8, x-4, x, x+4 */
  int qstep;
};

__device__ __forceinline__ void pp_vedge_load(PpEdge &E, const uint8_t *row, int x, bool valid, const PpArgs &A, const uint8_t *dcq) {
  E.w0 = E.w1 = E.w2 = E.w3 = 0;
  E.qstep = 0;
  if (valid) {
    const uint32_t *q = (const uint32_t *)(row + x - 8);
    E.w0 = __ldcg(q);
    E.w1 = __ldcg(q + 1);
    E.w2 = __ldcg(q + 2);
    E.w3 = __ldcg(q + 3);
    E.qstep = A.dc_scale[dcq[x >> 3]];
  }
}

__global__ void __launch_bounds__(256)
ocg_pp_vedge_kernel(const PpArgs A, uint8_t *__restrict__ pp, const uint8_t *__restrict__ dc_qis, int32_t *__restrict__ variances,
                    int pli0, int pli1) {
  int wrow = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)threadIdx.x & 31;
  for (int pli = pli0; pli <= pli1; pli++) {
    const PpPlane &P = A.p[pli];
    if (wrow >= P.H) { wrow -= P.H; continue; }
    const int y = wrow, by = y >> 3;
    uint8_t *row = pp + P.dst_off - (ptrdiff_t)y * P.W;
    const uint8_t *dcq = dc_qis + P.froffset + by * P.nh;
    int32_t *var = variances + P.froffset + by * P.nh;
    const int nedges = P.nh - 1; /* edges at x = 8, 16, ... */
    unsigned carry = 0;          /* did the previous edge rewrite its pixels? */
    PpEdge cur, nxt;
    pp_vedge_load(cur, row, 8 * (1 + lane), lane < nedges, A, dcq);
    for (int e0 = 0; e0 < nedges; e0 += 32) {
      /* the next round's pixels are read before this round's are written */
      const int en = e0 + 32 + lane;
      pp_vedge_load(nxt, row, 8 * (1 + en), en < nedges, A, dcq);
      const bool valid = e0 + lane < nedges;
      const int x = 8 * (1 + e0 + lane);
      int r[10];
      r[0] = (int)(cur.w0 >> 24);
#pragma unroll
      for (int k = 0; k < 4; k++) { r[1 + k] = (int)((cur.w1 >> (8 * k)) & 0xFFu); r[5 + k] = (int)((cur.w2 >> (8 * k)) & 0xFFu); }
      r[9] = (int)(cur.w3 & 0xFFu);
      /* what x-5 holds if the previous edge was filtered: its last output, from pixels no edge has touched */
      int alt;
      {
        const int p5 = (int)(cur.w0 & 0xFFu), p6 = (int)((cur.w0 >> 8) & 0xFFu), p7 = (int)((cur.w0 >> 16) & 0xFFu),
                  p8 = (int)(cur.w0 >> 24), p9 = r[1];
        alt = (p5 + p6 + p7 + p8 * 2 + p9 * 3 + 4) >> 3;
      }
      const int qstep = cur.qstep, flimit = (qstep * 3) >> 2;
      const int rest0 = pp_abs(r[2] - r[1]) + pp_abs(r[3] - r[2]) + pp_abs(r[4] - r[3]);
      int sum1 = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) sum1 += pp_abs(r[k + 5] - r[k + 6]);
      const bool other = sum1 < flimit && r[5] - r[4] < qstep && r[4] - r[5] < qstep;
      const bool g0 = valid && other && rest0 + pp_abs(r[1] - r[0]) < flimit;
      const bool g1 = valid && other && rest0 + pp_abs(r[1] - alt) < flimit;
      const unsigned G0 = __ballot_sync(0xFFFFFFFFu, g0), G1 = __ballot_sync(0xFFFFFFFFu, g1);
      unsigned D = 0, d = carry;
#pragma unroll 8
      for (int k = 0; k < 32; k++) {
        d = ((d ? G1 : G0) >> k) & 1u;
        D |= d << k;
      }
      const unsigned dprev = lane == 0 ? carry : (D >> (lane - 1)) & 1u;
      carry = (D >> 31) & 1u;
      if (valid) {
        if (dprev) r[0] = alt;
        const int sum0 = rest0 + pp_abs(r[1] - r[0]);
        atomicAdd(var + (x >> 3) - 1, min(255, sum0));
        atomicAdd(var + (x >> 3), min(255, sum1));
        if ((D >> lane) & 1u) {
          int o[8];
          pp_lowpass(r, o);
          uint32_t a = 0, b = 0;
#pragma unroll
          for (int k = 0; k < 4; k++) { a |= (uint32_t)(o[k] & 0xFF) << (8 * k); b |= (uint32_t)(o[4 + k] & 0xFF) << (8 * k); }
          __stcg((uint32_t *)(row + x - 4), a);
          __stcg((uint32_t *)(row + x), b);
        }
      }
      __syncwarp();
      cur = nxt;
    }
    return;
  }
}

/* ---- de-ringing ---------------------------------------------------------------------------------------
   oc_dering_block on one block, by one thread; pixels of the block and its rim in registers.
   px[r] = row r of the block (8 bytes), up/dn = the rows above/below (or the block's own edge row at the
   frame border), lf/rt = the columns left/right of it, one byte per row (or the block's own edge column). */
struct PpBlock {
  uint32_t px[8][2];
  uint32_t up[2], dn[2];
  uint32_t lf[2], rt[2];
};

__device__ __forceinline__ int pp_byte(const uint32_t (&w)[2], int i) { return (int)((w[i >> 2] >> (8 * (i & 3))) & 0xFFu); }

__device__ __forceinline__ int pp_mod(int diff, int dc_scale, int sharp_mod, int mod_hi, int shift) {
  const int mod = 32 + dc_scale - (pp_abs(diff) << shift);
  return mod < -64 ? sharp_mod : max(0, min(mod, mod_hi));
}

__device__ void pp_dering_block(PpBlock &B, int dc_scale, int sharp_mod, int strong) {
  const int mod_hi = min(3 * dc_scale, strong ? 32 : 24);
  const int shift = strong ? 0 : 1;
  /* vmod[by][bx], by = 0..8: between row by-1 and row by (decode.c:1806-1813); hmod[bx][by], bx = 0..8:
     between column bx-1 and column bx (1816-1827).  The reference tabulates all of them from the block as
     it is before the pass; here a row's are made just before the row is rewritten, from rows that are still
     untouched: vcur = vmod[by], vnext = vmod[by+1], h = hmod[.][by]. */
  int vcur[8], vnext[8], h[9];
#pragma unroll
  for (int bx = 0; bx < 8; bx++) vcur[bx] = pp_mod(pp_byte(B.px[0], bx) - pp_byte(B.up, bx), dc_scale, sharp_mod, mod_hi, shift);
  /* decode.c:1828-1886: raster order, in place: left and upper neighbours are already filtered */
  uint32_t prev[2] = {B.up[0], B.up[1]};
#pragma unroll
  for (int by = 0; by < 8; by++) {
#pragma unroll
    for (int bx = 0; bx < 8; bx++) {
      const int below = by < 7 ? pp_byte(B.px[by + 1], bx) : pp_byte(B.dn, bx);
      vnext[bx] = pp_mod(below - pp_byte(B.px[by], bx), dc_scale, sharp_mod, mod_hi, shift);
    }
    h[0] = pp_mod(pp_byte(B.px[by], 0) - pp_byte(B.lf, by), dc_scale, sharp_mod, mod_hi, shift);
#pragma unroll
    for (int bx = 1; bx < 8; bx++)
      h[bx] = pp_mod(pp_byte(B.px[by], bx) - pp_byte(B.px[by], bx - 1), dc_scale, sharp_mod, mod_hi, shift);
    h[8] = pp_mod(pp_byte(B.rt, by) - pp_byte(B.px[by], 7), dc_scale, sharp_mod, mod_hi, shift);
    uint32_t out[2] = {0u, 0u};
    int left = pp_byte(B.lf, by);
#pragma unroll
    for (int bx = 0; bx < 8; bx++) {
      const int cur = pp_byte(B.px[by], bx);
      const int upv = pp_byte(prev, bx);
      const int dnv = by < 7 ? pp_byte(B.px[by + 1], bx) : pp_byte(B.dn, bx);
      const int rtv = bx < 7 ? pp_byte(B.px[by], bx + 1) : pp_byte(B.rt, by);
      const int a = 128 - h[bx] - vcur[bx] - vnext[bx] - h[bx + 1];
      const int b = 64 + h[bx] * left + vcur[bx] * upv + vnext[bx] * dnv + h[bx + 1] * rtv;
      const int v = max(0, min(255, (a * cur + b) >> 7));
      out[bx >> 2] |= (uint32_t)v << (8 * (bx & 3));
      left = v;
    }
    B.px[by][0] = out[0];
    B.px[by][1] = out[1];
    prev[0] = out[0];
    prev[1] = out[1];
#pragma unroll
    for (int bx = 0; bx < 8; bx++) vcur[bx] = vnext[bx];
  }
}

/* At the frame border the rim aliases the block's own edge (decode.c:1802-1805, 1814-1826, 1829-1885):
   refreshed before every pass because the edge itself changes. */
__device__ __forceinline__ void pp_alias_rim(PpBlock &B, int b) {
  if (b & 4) { B.up[0] = B.px[0][0]; B.up[1] = B.px[0][1]; }
  if (b & 8) { B.dn[0] = B.px[7][0]; B.dn[1] = B.px[7][1]; }
  if (b & 1) {
    B.lf[0] = B.lf[1] = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) B.lf[r >> 2] |= (B.px[r][0] & 0xFFu) << (8 * (r & 3));
  }
  if (b & 2) {
    B.rt[0] = B.rt[1] = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) B.rt[r >> 2] |= (B.px[r][1] >> 24) << (8 * (r & 3));
  }
}

#define OCG_PP_T1 384
#define OCG_PP_T2 (4 * OCG_PP_T1)
#define OCG_PP_T3 (5 * OCG_PP_T1)
#define OCG_PP_T4 (10 * OCG_PP_T1)

/* grid.x = planes to process, one CTA per plane, 32 * ceil(nv/32) threads; shared: int progress[nv] */
__global__ void __launch_bounds__(1024)
ocg_pp_dering_kernel(const PpArgs A, uint8_t *__restrict__ pp, const uint8_t *__restrict__ qis, const int32_t *__restrict__ variances,
                     int pli0) {
  extern __shared__ int pp_progress[]; /* blocks finished per block row */
  const int pli = pli0 + (int)blockIdx.x;
  const PpPlane &P = A.p[pli];
  const int nh = P.nh, nv = P.nv;
  const int by = (int)threadIdx.x, lane = (int)threadIdx.x & 31;
  for (int i = (int)threadIdx.x; i < nv; i += (int)blockDim.x) pp_progress[i] = 0;
  __syncthreads();
  const bool live = by < nv;
  const int strong = A.level >= (pli ? 7 : 4);
  const int sthresh = pli ? OCG_PP_T4 : OCG_PP_T3;
  const int32_t *var = variances + P.froffset + (live ? by : 0) * nh;
  const uint8_t *qrow = qis + P.froffset + (live ? by : 0) * nh;
  volatile int *prog = pp_progress;
  /* step t: lane l of a warp works on block t - l of its row; the row above is lane l-1 (one block ahead,
     finished in the previous step) or, for lane 0, the last row of the previous warp */
  const int nsteps = nh + 31;
  for (int t = 0; t < nsteps; t++) {
    const int bx = t - lane;
    if (live && bx >= 0 && bx < nh) {
      const int v = var[bx];
      const int b = (bx == 0 ? 1 : 0) | (bx == nh - 1 ? 2 : 0) | (by == 0 ? 4 : 0) | (by == nv - 1 ? 8 : 0);
      int passes = 0, st = 0;
      if (strong && v > sthresh) {
        passes = 1; st = 1;
        if (pli || (!(b & 1) && var[bx - 1] > OCG_PP_T4) || (!(b & 2) && var[bx + 1] > OCG_PP_T4) ||
            (!(b & 4) && var[bx - nh] > OCG_PP_T4) || (!(b & 8) && var[bx + nh] > OCG_PP_T4))
          passes = 3;
      } else if (v > OCG_PP_T2) { passes = 1; st = 1; }
      else if (v > OCG_PP_T1) { passes = 1; st = 0; }
      if (passes) {
        if (lane == 0 && by > 0) {
          while (prog[by - 1] <= bx) __nanosleep(100);
          __threadfence_block();
        }
        uint8_t *o = pp + P.dst_off - (ptrdiff_t)(8 * by) * P.W + 8 * bx; /* row r of the block at o - r*W */
        PpBlock B;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const uint2 q = __ldcg((const uint2 *)(o - (ptrdiff_t)r * P.W));
          B.px[r][0] = q.x;
          B.px[r][1] = q.y;
        }
        B.up[0] = B.up[1] = B.dn[0] = B.dn[1] = B.lf[0] = B.lf[1] = B.rt[0] = B.rt[1] = 0;
        if (!(b & 4)) { const uint2 q = __ldcg((const uint2 *)(o + P.W)); B.up[0] = q.x; B.up[1] = q.y; }
        if (!(b & 8)) { const uint2 q = __ldcg((const uint2 *)(o - (ptrdiff_t)8 * P.W)); B.dn[0] = q.x; B.dn[1] = q.y; }
        if (!(b & 1)) {
#pragma unroll
          for (int r = 0; r < 8; r++) B.lf[r >> 2] |= (uint32_t)__ldcg(o - (ptrdiff_t)r * P.W - 1) << (8 * (r & 3));
        }
        if (!(b & 2)) {
#pragma unroll
          for (int r = 0; r < 8; r++) B.rt[r >> 2] |= (uint32_t)__ldcg(o - (ptrdiff_t)r * P.W + 8) << (8 * (r & 3));
        }
        const int qi = qrow[bx];
        const int dc_scale = A.dc_scale[qi], sharp_mod = A.sharp_mod[qi];
#pragma unroll 1
        for (int k = 0; k < passes; k++) {
          pp_alias_rim(B, b);
          pp_dering_block(B, dc_scale, sharp_mod, st);
        }
#pragma unroll
        for (int r = 0; r < 8; r++) __stcg((uint2 *)(o - (ptrdiff_t)r * P.W), make_uint2(B.px[r][0], B.px[r][1]));
        __threadfence_block();
      }
      prog[by] = bx + 1;
    }
    __syncwarp();
  }
}

} /* namespace */

/* ------------------------------------------------------------------------ */
struct ocg_pp {
  ocg_ctx *ctx = nullptr;
  int device = 0;
  PpArgs args;
  size_t pp_bytes = 0, plane_bytes[3] = {0, 0, 0}, plane_start[3] = {0, 0, 0};
  uint8_t *d_pp = nullptr;
  int32_t *d_var = nullptr;
  uint8_t *d_q = nullptr;  /* [2][nfrags]: dc_qis, qis */
  uint8_t *h_q = nullptr;  /* pinned staging of the same, two slots used in turn */
  cudaEvent_t h_q_free[2] = {nullptr, nullptr}; /* recorded behind the upload that reads a slot */
  int h_q_slot = 0;
  int nfrags = 0;
  int last_level = 0;
};

extern "C" {

OCG_API void ocg_pp_destroy(ocg_pp *pp) {
  if (pp == nullptr) return;
  ocg_set_device(pp->device);
  cudaStreamSynchronize((cudaStream_t)ocg_ctx_stream(pp->ctx));
  cudaFree(pp->d_pp);
  cudaFree(pp->d_var);
  cudaFree(pp->d_q);
  if (pp->h_q) cudaFreeHost(pp->h_q);
  for (int i = 0; i < 2; i++) if (pp->h_q_free[i]) cudaEventDestroy(pp->h_q_free[i]);
  delete pp;
}

OCG_API int ocg_pp_create(ocg_pp **out, ocg_ctx *ctx) {
  if (out == nullptr || ctx == nullptr) return OCG_EFAULT;
  const ocg_geometry *g = ocg_ctx_geometry(ctx);
  if (ocg_set_device(ocg_ctx_device(ctx)) != cudaSuccess) return OCG_ECUDA;
  ocg_pp *pp = new ocg_pp;
  pp->ctx = ctx;
  pp->device = ocg_ctx_device(ctx);
  pp->nfrags = g->nfrags;
  memset(&pp->args, 0, sizeof(pp->args));
  size_t at = 0;
  for (int pli = 0; pli < 3; pli++) {
    const ocg_plane_geom &p = g->planes[pli];
    PpPlane &q = pp->args.p[pli];
    q.W = p.width; q.H = p.height; q.nh = p.nhfrags; q.nv = p.nvfrags; q.froffset = p.froffset;
    q.src_off = (int32_t)p.plane_off;
    q.src_stride = p.ystride;
    pp->plane_start[pli] = at;
    pp->plane_bytes[pli] = (size_t)p.width * p.height;
    q.dst_off = (int32_t)(at + (size_t)(p.height - 1) * p.width);
    at += pp->plane_bytes[pli];
  }
  pp->pp_bytes = at;
#define PP_CU(x) do { if ((x) != cudaSuccess) { cudaGetLastError(); ocg_pp_destroy(pp); return OCG_ECUDA; } } while (0)
  PP_CU(cudaMalloc(&pp->d_pp, at + 64));
  PP_CU(cudaMalloc(&pp->d_var, (size_t)g->nfrags * sizeof(int32_t)));
  PP_CU(cudaMalloc(&pp->d_q, (size_t)g->nfrags * 2));
  PP_CU(cudaHostAlloc(&pp->h_q, (size_t)g->nfrags * 4, cudaHostAllocDefault));
  for (int i = 0; i < 2; i++) PP_CU(cudaEventCreateWithFlags(&pp->h_q_free[i], cudaEventDisableTiming));
#undef PP_CU
  *out = pp;
  return OCG_OK;
}

/* Queues the filters for the frame in buffer `self_buf` on the context's stream (behind that frame's
   reconstruction).  level: the reference's OC_PP_LEVEL_* (2 de-block luma, 3 + de-ring luma, 4 strong,
   5..7 the same for chroma on top).  dc_qis / qis: per fragment, the DC quantiser index the reference
   tracks (decode.c:1204-1243) and state.qis[frag.qii]. */
OCG_API int ocg_pp_run(ocg_pp *pp, int self_buf, int level, const int32_t *dc_scale, const int32_t *sharp_mod, const uint8_t *dc_qis,
                       const uint8_t *qis) {
  if (pp == nullptr || dc_scale == nullptr || sharp_mod == nullptr || dc_qis == nullptr || qis == nullptr) return OCG_EFAULT;
  if (level < 2 || level > 7) return OCG_EINVAL;
  const ocg_geometry *g = ocg_ctx_geometry(pp->ctx);
  if (self_buf < 0 || self_buf >= g->nrefs) return OCG_EINVAL;
  if (ocg_set_device(pp->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(pp->ctx);
  /* nothing is waited for here except the upload that last read this staging slot (two frames ago) */
  const int slot = pp->h_q_slot;
  pp->h_q_slot ^= 1;
  uint8_t *hq = pp->h_q + (size_t)slot * 2 * pp->nfrags;
  if (cudaEventSynchronize(pp->h_q_free[slot]) != cudaSuccess) return OCG_ECUDA;
  memcpy(pp->args.dc_scale, dc_scale, sizeof(pp->args.dc_scale));
  memcpy(pp->args.sharp_mod, sharp_mod, sizeof(pp->args.sharp_mod));
  pp->args.level = level;
  memcpy(hq, dc_qis, (size_t)pp->nfrags);
  memcpy(hq + pp->nfrags, qis, (size_t)pp->nfrags);
  if (cudaMemcpyAsync(pp->d_q, hq, (size_t)pp->nfrags * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaEventRecord(pp->h_q_free[slot], st) != cudaSuccess ||
      cudaMemsetAsync(pp->d_var, 0, (size_t)pp->nfrags * sizeof(int32_t), st) != cudaSuccess)
    return OCG_ECUDA;
  const uint8_t *src = (const uint8_t *)ocg_ctx_frame_devptr(pp->ctx, self_buf) + g->base_off;
  const int pli1 = level >= 5 ? 2 : 0;
  long nh = 0, nvw = 0;
  for (int pli = 0; pli <= pli1; pli++) {
    nh += (long)(pp->args.p[pli].W >> 2) * (pp->args.p[pli].nv + 1);
    nvw += pp->args.p[pli].H;
  }
  ocg_pp_hedge_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(pp->args, src, pp->d_pp, pp->d_q, pp->d_var, 0, pli1);
  ocg_pp_vedge_kernel<<<(unsigned)((nvw * 32 + 255) / 256), 256, 0, st>>>(pp->args, pp->d_pp, pp->d_q, pp->d_var, 0, pli1);
  ocg_count_launch(2);
  /* de-ringing: luma from level 3, chroma from level 6 */
  const int dr1 = level >= 6 ? 2 : (level >= 3 ? 0 : -1);
  for (int pli = 0; pli <= dr1; pli++) {
    const int nv = pp->args.p[pli].nv;
    const int threads = 32 * ((nv + 31) / 32);
    if (threads > 1024) return OCG_EINVAL; /* planes taller than 8192 pixels */
    ocg_pp_dering_kernel<<<1, threads, (size_t)nv * sizeof(int), st>>>(pp->args, pp->d_pp, pp->d_q + pp->nfrags, pp->d_var, pli);
    ocg_count_launch(1);
  }
  pp->last_level = level;
  if (cudaGetLastError() != cudaSuccess) return OCG_ECUDA;
  return OCG_OK;
}

/* Waits for the filters and copies the post-processed planes (tightly packed: luma, Cb, Cr, each W x H,
   top row first -- the layout of the reference's pp_frame_data, decode.c:1283-1315) to the host.  Only the
   planes the level processed are copied: luma, and chroma from level 5. */
OCG_API int ocg_pp_download(ocg_pp *pp, uint8_t *host_dst) {
  if (pp == nullptr || host_dst == nullptr) return OCG_EFAULT;
  if (pp->last_level < 2) return OCG_EINVAL;
  if (ocg_set_device(pp->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(pp->ctx);
  const size_t n = pp->last_level >= 5 ? pp->pp_bytes : pp->plane_bytes[0];
  if (cudaMemcpyAsync(host_dst, pp->d_pp, n, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    cudaGetLastError();
    return OCG_ECUDA;
  }
  return OCG_OK;
}

/* Test hook: the variances of the last run. */
OCG_API int ocg_pp_download_variances(ocg_pp *pp, int32_t *host_dst) {
  if (pp == nullptr || host_dst == nullptr) return OCG_EFAULT;
  if (ocg_set_device(pp->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(pp->ctx);
  if (cudaMemcpyAsync(host_dst, pp->d_var, (size_t)pp->nfrags * sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return OCG_ECUDA;
  return OCG_OK;
}

} /* extern "C" */
