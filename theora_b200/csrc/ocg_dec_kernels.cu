/* Decode-side kernels for sm_100a.
 *
 *  ocg_recon_simple_kernel / ocg_recon_xform_kernel
 *                     fused DC dequant + 8x8 iDCT + intra/inter/inter2
 *                     reconstruction, plus the uncoded-fragment copy, for a
 *                     batch of (stream, frame) jobs: pass A moves everything
 *                     that needs no transform, pass B transforms the rest
 *                     from a compact list.  Replaces
 *                     oc_state_frag_recon (reference lib/state.c:959),
 *                     oc_idct8x8 (idct.c:301), oc_frag_recon_* (fragment.c:49-80)
 *                     and oc_frag_copy_list (fragment.c:37).
 *  ocg_lf_kernel      the normative loop filter (state.c:1002-1105) in its
 *                     order-free "corner cell" form.
 *  ocg_border_kernel  apron replication (state.c:770-835).
 *
 * Work mapping: a fragment is handled by 4 adjacent lanes, lane l owning rows 2l
 * and 2l+1 (a warp covers 8 raster-consecutive fragments, so the 64-bit row
 * stores of adjacent lanes coalesce).  Pass B partitions each 64-fragment
 * chunk of its list by footprint class in shared memory (stable) first.
 * The two 1-D passes run in registers; the transposes between them
 * are two __shfl_xor stages on 16-bit pairs; the final add/clamp uses the
 * packed-halfword DPX instructions (VIADDMNMX.S16x2.RELU / VIMNMX.S16x2).
 * All loads/stores are 64- or 128-bit.  There is no dense contraction here, so
 * tensor cores are not used.
 */
#include <algorithm>
#include "ocg_internal.h"

namespace {

/* Programmatic dependent launch: a kernel launched with ocg_launch_pdl may become resident while the kernel
   before it in the stream is still running; it must call pdl_wait() before it touches anything that kernel
   wrote (the wait returns when the predecessor has completed and its writes are visible).  pdl_release()
   lets the NEXT kernel's CTAs in; it is placed behind pdl_wait(), so a kernel's successor never runs ahead
   of the kernel's own predecessor.  Hides the launch latency and the prologue (table staging, flag loads)
   of the four kernels of a frame behind the tail of the one before. */
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int K1 = 64277, K2 = 60547, K3 = 54491, K4 = 46341, K5 = 36410, K6 = 25080, K7 = 12785;

__device__ __forceinline__ int sext16(int v) { return (int)(short)v; }
__device__ __forceinline__ int mulhi16(int k, int v) { return (k * v) >> 16; }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }
__device__ __forceinline__ int lo16(uint32_t p) { return (int)(short)(p & 0xFFFFu); }
__device__ __forceinline__ int hi16(uint32_t p) { return (int)p >> 16; }

/* idct.c:30-203.  NR = number of leading inputs that may be non-zero (the
   reduced reference variants idct8_2/3/4 are exactly this with the trailing
   inputs dropped).  Inputs are sign-extended int16; outputs are 32-bit sums
   whose low 16 bits are the reference's (ogg_int16_t) results. */
template <int NR>
__device__ __forceinline__ void idct8(const int (&x)[8], int (&y)[8]) {
  const int x0 = x[0];
  const int x1 = NR > 1 ? x[1] : 0;
  const int x2 = NR > 2 ? x[2] : 0;
  const int x3 = NR > 3 ? x[3] : 0;
  const int x4 = NR > 4 ? x[4] : 0;
  const int x5 = NR > 5 ? x[5] : 0;
  const int x6 = NR > 6 ? x[6] : 0;
  const int x7 = NR > 7 ? x[7] : 0;
  int e0, e1;
  if (NR > 4) {
    e0 = mulhi16(K4, sext16(x0 + x4));
    e1 = mulhi16(K4, sext16(x0 - x4));
  } else {
    e0 = e1 = mulhi16(K4, x0);
  }
  const int e2 = mulhi16(K6, x2) - mulhi16(K2, x6);
  const int e3 = mulhi16(K2, x2) + mulhi16(K6, x6);
  const int o4 = mulhi16(K7, x1) - mulhi16(K1, x7);
  const int o5 = mulhi16(K3, x5) - mulhi16(K5, x3);
  const int o6 = mulhi16(K5, x5) + mulhi16(K3, x3);
  const int o7 = mulhi16(K1, x1) + mulhi16(K7, x7);
  const int s4 = o4 + o5;
  const int s5 = mulhi16(K4, NR > 3 ? sext16(o4 - o5) : o4);
  const int s7 = o7 + o6;
  const int s6 = mulhi16(K4, NR > 3 ? sext16(o7 - o6) : o7);
  const int a0 = e0 + e3, a3 = e0 - e3;
  const int a1 = e1 + e2, a2 = e1 - e2;
  const int b6 = s6 + s5, b5 = s6 - s5;
  y[0] = a0 + s7;
  y[1] = a1 + b6;
  y[2] = a2 + b5;
  y[3] = a3 + s4;
  y[4] = a3 - s4;
  y[5] = a2 - b5;
  y[6] = a1 - b6;
  y[7] = a0 - s7;
}

/* Register layout conventions for the 4-lane fragment group (lane l = 0..3):
     L_row: lane owns rows 2l,2l+1; q[2m+r0] = (v[2l+r0][2m], v[2l+r0][2m+1])
     L_col: lane owns cols 2l,2l+1; q[r]     = (v[r][2l],     v[r][2l+1])
   The same two-stage exchange converts either layout into the other.  All 32
   lanes of the warp must call it together (xor 1 and 2 stay inside a group):
   a full-warp mask keeps the shuffles plain SHFL.BFLY instead of the
   MATCH/WARPSYNC sequences partial masks compile to. */
__device__ __forceinline__ void xpose(uint32_t (&q)[8], int l) {
  const bool hi1 = (l & 2) != 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t s = hi1 ? q[i] : q[i + 4];
    uint32_t r = __shfl_xor_sync(0xFFFFFFFFu, s, 2);
    if (hi1) q[i] = r; else q[i + 4] = r;
  }
  const bool hi0 = (l & 1) != 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int i = (j & 1) | ((j & 2) << 1); /* 0,1,4,5 */
    uint32_t s = hi0 ? q[i] : q[i | 2];
    uint32_t r = __shfl_xor_sync(0xFFFFFFFFu, s, 1);
    if (hi0) q[i] = r; else q[i | 2] = r;
  }
}

/* 8 bytes at an arbitrary byte address out of the enclosing 16-byte window
   (two aligned 64-bit loads, then byte permutes). */
__device__ __forceinline__ uint2 load8_unaligned(const uint8_t *p) {
  const uintptr_t a = (uintptr_t)p;
  const uint2 *w = (const uint2 *)(a & ~(uintptr_t)7);
  const unsigned sh = (unsigned)(a & 7);
  const uint2 w0 = __ldg(w);
  if (sh == 0) return w0;
  const uint2 w1 = __ldg(w + 1);
  const unsigned sel = 0x3210u + 0x1111u * (sh & 3);
  uint2 r;
  if (sh < 4) {
    r.x = __byte_perm(w0.x, w0.y, sel);
    r.y = __byte_perm(w0.y, w1.x, sel);
  } else {
    r.x = __byte_perm(w0.y, w1.x, sel);
    r.y = __byte_perm(w1.x, w1.y, sel);
  }
  return r;
}

/* state.c:846-957 in closed form: first tap truncates towards zero, second tap
   (present iff a component has a fractional part) one step away from zero. */
__device__ __forceinline__ void mv_taps(int mv, int qx, int qy, int ystride, int &off0, int &fx, int &fy) {
  const int dx = (int)(signed char)(mv & 0xFF);
  const int dy = mv >> 8;
  const int ax = abs(dx), ay = abs(dy);
  const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
  const int mx = sx * (ax >> (1 + qx)), my = sy * (ay >> (1 + qy));
  fx = (ax & (qx ? 3 : 1)) ? sx : 0; /* second tap = first + fy*ystride + fx, present iff fx|fy */
  fy = (ay & (qy ? 3 : 1)) ? sy : 0;
  off0 = my * ystride + mx;
}

/* clamp255(res + pred) for one row held as four (even,odd) halfword pairs.
   fragment.c:49-80.  The residue is first limited to <=255 so the packed add
   cannot wrap; values above 255 saturate the result either way. */
__device__ __forceinline__ uint2 recon_row(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint2 pred) {
  const uint32_t c255 = 0x00FF00FFu;
  const uint32_t p0 = __byte_perm(pred.x, 0, 0x4140), p1 = __byte_perm(pred.x, 0, 0x4342);
  const uint32_t p2 = __byte_perm(pred.y, 0, 0x4140), p3 = __byte_perm(pred.y, 0, 0x4342);
  const uint32_t o0 = __viaddmin_s16x2_relu(__vmins2(r0, c255), p0, c255);
  const uint32_t o1 = __viaddmin_s16x2_relu(__vmins2(r1, c255), p1, c255);
  const uint32_t o2 = __viaddmin_s16x2_relu(__vmins2(r2, c255), p2, c255);
  const uint32_t o3 = __viaddmin_s16x2_relu(__vmins2(r3, c255), p3, c255);
  return make_uint2(__byte_perm(o0, o1, 0x6420), __byte_perm(o2, o3, 0x6420));
}

/* Work classes inside a CTA, in processing order. */
enum { WC_NONE = 0, WC_COPY = 1, WC_DC = 2, WC_3 = 3, WC_10 = 4, WC_FULL = 5, WC_COUNT = 6 };

__device__ __forceinline__ int work_class(const int4 rw) {
  const int refi = (rw.w >> 16) & 0xFF, lz = (rw.w >> 8) & 0xFF;
  if (refi == OCG_FRAG_UNCODED) return WC_COPY;
  /* state.c:967 and idct.c:327-329 */
  return lz < 2 ? WC_DC : (lz <= 3 ? WC_3 : (lz <= 10 ? WC_10 : WC_FULL));
}

/* Both 1-D passes, the transposes and the (v+8)>>4 rounding for the two rows
   this lane owns (idct.c:301-330).  Executed by the whole warp with NR chosen
   from the largest footprint present in the warp; a group whose own footprint
   is smaller has had the coefficients the reference ignores zeroed, and the
   reduced reference transforms are exact specialisations of the larger ones. */
template <int NR>
__device__ __forceinline__ void idct_rows2(int (&xa)[8], int (&xb)[8], uint32_t (&q)[8], int l) {
  int ya[8], yb[8];
  idct8<NR>(xa, ya);
  idct8<NR>(xb, yb);
#pragma unroll
  for (int m = 0; m < 4; m++) {
    q[2 * m] = pack16(ya[2 * m], ya[2 * m + 1]);
    q[2 * m + 1] = pack16(yb[2 * m], yb[2 * m + 1]);
  }
  xpose(q, l);
#pragma unroll
  for (int r = 0; r < 8; r++) { xa[r] = lo16(q[r]); xb[r] = hi16(q[r]); }
  idct8<NR>(xa, ya);
  idct8<NR>(xb, yb);
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int va = (sext16(ya[r]) + 8) >> 4;
    const int vb = (sext16(yb[r]) + 8) >> 4;
    q[r] = pack16(va, vb);
  }
  xpose(q, l);
}

/* Zeroes the halfwords of a coefficient row the reference's reduced transform
   would not read: keeps the first `lim` (0..8) coefficients. */
__device__ __forceinline__ uint4 keep_first(uint4 w, int lim) {
  const uint32_t m0 = lim >= 2 ? 0xFFFFFFFFu : (lim == 1 ? 0x0000FFFFu : 0u);
  const uint32_t m1 = lim >= 4 ? 0xFFFFFFFFu : (lim == 3 ? 0x0000FFFFu : 0u);
  const uint32_t m2 = lim >= 6 ? 0xFFFFFFFFu : (lim == 5 ? 0x0000FFFFu : 0u);
  const uint32_t m3 = lim >= 8 ? 0xFFFFFFFFu : (lim == 7 ? 0x0000FFFFu : 0u);
  return make_uint4(w.x & m0, w.y & m1, w.z & m2, w.w & m3);
}

/* Pass B for one 4-lane group (class bits == WC_NONE for idle groups, which
   still take part in the warp-wide shuffles). */
__device__ __forceinline__ void xform_group(const OcgGeomDev &g, const OcgJobDev &job, const int4 it, unsigned tap,
                                            int l) {
  const int my = (it.w >> 24) & 7;
  const int wmax = __reduce_max_sync(0xFFFFFFFFu, my); /* warp-uniform */
  if (wmax == WC_NONE) return;
  const int pli = (it.w >> 29) & 3;
  const int refi = (it.w >> 27) & 3;
  const int ystride = g.p[pli].ystride;
  uint32_t q[8];
  {
    /* this lane's two coefficient rows (zero unless stored and inside the
       footprint of the group's own class) */
    const int nfoot = my == WC_FULL ? 8 : (my == WC_10 ? 4 : (my == WC_3 ? 2 : 0));
    const int ra = 2 * l, rb = 2 * l + 1;
    const unsigned rowmask = ((unsigned)it.w >> 16) & 0xFFu;
    uint4 wa = make_uint4(0, 0, 0, 0), wb = make_uint4(0, 0, 0, 0);
    if (job.dense_rows) {
      /* device-side expansion: row r of the block at coeff_row + r; whatever was stored is cleared again
         (the blocks are all-zero between frames, idct.c:245,276,295 does the same to its input) */
      uint4 *rows = (uint4 *)job.rows + (unsigned)it.z;
      const uint4 zero = make_uint4(0, 0, 0, 0);
      if (my != WC_NONE) {
        if (rowmask >> ra & 1) { if (ra < nfoot) wa = rows[ra]; rows[ra] = zero; }
        if (rowmask >> rb & 1) { if (rb < nfoot) wb = rows[rb]; rows[rb] = zero; }
      }
      if (my != WC_FULL) {
        wa = keep_first(wa, max(0, nfoot - ra));
        wb = keep_first(wb, max(0, nfoot - rb));
      }
    } else if (ra < nfoot) {
      const uint4 *rows = (const uint4 *)job.rows + (unsigned)it.z;
      if (rowmask >> ra & 1) wa = __ldg(rows + __popc(rowmask & ((1u << ra) - 1u)));
      if (rowmask >> rb & 1) wb = __ldg(rows + __popc(rowmask & ((1u << rb) - 1u)));
      if (my != WC_FULL) {
        /* triangular footprint of the reduced transforms (idct.c:213-275) */
        wa = keep_first(wa, nfoot - ra);
        wb = keep_first(wb, nfoot - rb);
      }
    }
    int xa[8], xb[8];
    xa[0] = lo16(wa.x); xa[1] = hi16(wa.x); xa[2] = lo16(wa.y); xa[3] = hi16(wa.y);
    xa[4] = lo16(wa.z); xa[5] = hi16(wa.z); xa[6] = lo16(wa.w); xa[7] = hi16(wa.w);
    xb[0] = lo16(wb.x); xb[1] = hi16(wb.x); xb[2] = lo16(wb.y); xb[3] = hi16(wb.y);
    xb[4] = lo16(wb.z); xb[5] = hi16(wb.z); xb[6] = lo16(wb.w); xb[7] = hi16(wb.w);
    if (l == 0 && my != WC_NONE) xa[0] = sext16(it.w); /* dequantised DC, state.c:978 */
    if (wmax == WC_3) idct_rows2<2>(xa, xb, q, l);
    else if (wmax == WC_10) idct_rows2<4>(xa, xb, q, l);
    else idct_rows2<8>(xa, xb, q, l);
  }
  if (my == WC_NONE) return;
  uint8_t *dst = job.base[OCG_FRAME_SELF] + it.x + (2 * l) * ystride;
  /* prediction, rows 2l and 2l+1 (fragment.c:49-80) */
  uint2 pa, pb;
  if (refi == OCG_FRAME_SELF) {
    pa = pb = make_uint2(0x80808080u, 0x80808080u);
  } else {
    const uint8_t *ref = job.base[refi] + it.y + (2 * l) * ystride;
    pa = load8_unaligned(ref);
    pb = load8_unaligned(ref + ystride);
    if (tap) {
      const int fx = (int)(tap << 30) >> 30, fy = (int)(tap << 28) >> 30;
      const uint8_t *ref2 = ref + fy * ystride + fx;
      const uint2 ta = load8_unaligned(ref2);
      const uint2 tb = load8_unaligned(ref2 + ystride);
      pa.x = __vhaddu4(pa.x, ta.x); pa.y = __vhaddu4(pa.y, ta.y);
      pb.x = __vhaddu4(pb.x, tb.x); pb.y = __vhaddu4(pb.y, tb.y);
    }
  }
  const uint2 oa = recon_row(q[0], q[2], q[4], q[6], pa);
  const uint2 ob = recon_row(q[1], q[3], q[5], q[7], pb);
  *(uint2 *)dst = oa;
  *(uint2 *)(dst + ystride) = ob;
}

/* ---- recon pass A: everything that needs no transform -------------------
   Uncoded copies and DC-only fragments (state.c:967-975) are pure data
   movement (+ one add/clamp); in inter frames they are ~97 % of the work.
   ONE LANE PER FRAGMENT: a lane loads its fragment's 16-byte record, decodes
   the motion vector once, and moves all 8 rows (up to 16 independent 64-bit
   loads in flight per lane).  The 32 lanes of a warp own 32 raster-consecutive
   fragments, so every row access of the warp is one contiguous 256-byte run.
   The byte shift of an unaligned predictor is the same for all 8 rows (row
   strides are multiples of 16), so the permute selectors are computed once.
   Fragments that need an iDCT are appended to the job's compact work list (one
   warp-aggregated atomic) for pass B.  Pass A also writes the coded map the
   loop filter reads. */
#define OCG_SIMPLE_THREADS 128

template <int OCG_A_ROWS> struct RowsN { uint2 r[OCG_A_ROWS]; };

/* rows p, p+ystride, ... : 8 bytes each at an arbitrary byte address */
template <int OCG_A_ROWS>
__device__ __forceinline__ void load_rows(const uint8_t *p, int ystride, RowsN<OCG_A_ROWS> &o) {
  const uintptr_t a = (uintptr_t)p;
  const unsigned sh = (unsigned)(a & 7);
  const uint8_t *base = (const uint8_t *)(a & ~(uintptr_t)7);
  if (sh == 0) {
#pragma unroll
    for (int i = 0; i < OCG_A_ROWS; i++) o.r[i] = __ldg((const uint2 *)(base + i * ystride));
    return;
  }
  uint2 w0[OCG_A_ROWS], w1[OCG_A_ROWS];
#pragma unroll
  for (int i = 0; i < OCG_A_ROWS; i++) {
    w0[i] = __ldg((const uint2 *)(base + i * ystride));
    w1[i] = __ldg((const uint2 *)(base + i * ystride) + 1);
  }
  const unsigned sel = 0x3210u + 0x1111u * (sh & 3);
  if (sh < 4) {
#pragma unroll
    for (int i = 0; i < OCG_A_ROWS; i++) {
      o.r[i].x = __byte_perm(w0[i].x, w0[i].y, sel);
      o.r[i].y = __byte_perm(w0[i].y, w1[i].x, sel);
    }
  } else {
#pragma unroll
    for (int i = 0; i < OCG_A_ROWS; i++) {
      o.r[i].x = __byte_perm(w0[i].y, w1[i].x, sel);
      o.r[i].y = __byte_perm(w1[i].x, w1[i].y, sel);
    }
  }
}

/* OCG_A_ROWS: rows moved per step (2 steps of 4 keep the kernel at 40 registers) */
template <int OCG_A_ROWS, int MINB>
__global__ void __launch_bounds__(OCG_SIMPLE_THREADS, MINB)
ocg_recon_simple_kernel(const OcgGeomDev g, const OcgJobDev *__restrict__ jobs) {
  pdl_wait(); /* the records (built by the kernel before, in the token path) and the previous frame's pixels */
  pdl_release();
  const OcgJobDev &job = jobs[blockIdx.y];
  const int lane = (int)threadIdx.x & 31;
  const int fragi = (int)(blockIdx.x * OCG_SIMPLE_THREADS + threadIdx.x);
  int4 rw = make_int4(0, 0, 0, 0);
  int cls = WC_NONE;
  if (fragi < g.nfrags) {
    rw = __ldg((const int4 *)(job.recs + fragi));
    cls = work_class(rw);
    job.coded[fragi] = (unsigned char)(cls != WC_COPY);
  }
  /* hand the transform fragments to pass B */
  const unsigned want = __ballot_sync(0xFFFFFFFFu, cls >= WC_3);
  if (want) {
    int base = 0;
    const int leader = __ffs(want) - 1;
    if (lane == leader) base = atomicAdd(job.xcount, __popc(want));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (cls >= WC_3) job.xlist[base + __popc(want & ((1u << lane) - 1u))] = fragi;
  }
  if (cls != WC_COPY && cls != WC_DC) return;
  const int pli = (rw.w >> 24) & 3;
  const int ystride = g.p[pli].ystride;
  uint8_t *dst = job.base[OCG_FRAME_SELF] + rw.x;
  /* source of the block: PREV co-located (copy), a motion-displaced reference
     (one or two taps), or the constant 128 of intra prediction */
  const int refi = cls == WC_COPY ? OCG_FRAME_PREV : (rw.w >> 16) & 3;
  const uint8_t *ref = nullptr;
  int tap2 = 0;
  bool two = false;
  if (refi != OCG_FRAME_SELF) {
    int off0 = 0, fx = 0, fy = 0;
    if (cls == WC_DC) mv_taps(rw.y << 16 >> 16, pli ? g.qx : 0, pli ? g.qy : 0, ystride, off0, fx, fy);
    ref = job.base[refi] + rw.x + off0;
    tap2 = fy * ystride + fx;
    two = (fx | fy) != 0;
  }
  /* state.c:967-975: p=(dc*dc_quant+15)>>5 replicated over the block; 0 for copies */
  const int p = cls == WC_DC ? sext16(((rw.y >> 16) * (int)job.dcq[pli][(rw.w >> 26) & 1] + 15) >> 5) : 0;
  const uint32_t pp = pack16(p, p);
#pragma unroll 1
  for (int r0 = 0; r0 < 8; r0 += OCG_A_ROWS) {
    RowsN<OCG_A_ROWS> px;
    if (ref == nullptr) {
#pragma unroll
      for (int i = 0; i < OCG_A_ROWS; i++) px.r[i] = make_uint2(0x80808080u, 0x80808080u);
    } else {
      load_rows(ref + r0 * ystride, ystride, px);
      if (two) {
        RowsN<OCG_A_ROWS> t2;
        load_rows(ref + r0 * ystride + tap2, ystride, t2);
#pragma unroll
        for (int i = 0; i < OCG_A_ROWS; i++) {
          px.r[i].x = __vhaddu4(px.r[i].x, t2.r[i].x);
          px.r[i].y = __vhaddu4(px.r[i].y, t2.r[i].y);
        }
      }
    }
    if (p != 0) {
#pragma unroll
      for (int i = 0; i < OCG_A_ROWS; i++) px.r[i] = recon_row(pp, pp, pp, pp, px.r[i]);
    } /* else zero residual / copy: clamp255(0 + pred) == pred */
#pragma unroll
    for (int i = 0; i < OCG_A_ROWS; i++) *(uint2 *)(dst + (r0 + i) * ystride) = px.r[i];
  }
}

/* ---- recon pass B: fragments with an 8x8 iDCT ------------------------------
   A warp takes 32 entries of the compact list pass A produced, turns their
   records into work items (MV -> tap offsets, state.c:846-957; DC dequant,
   state.c:978), partitions them by footprint class (stable, so neighbours
   stay neighbours) and runs 4 lanes per fragment.
     item.x  buf_off                 item.y  buf_off + first-tap offset
     item.z  coeff_row
     item.w  [15:0] dequantised DC  [23:16] rowmask  [26:24] class
             [28:27] refi  [30:29] plane
     stap    second-tap step: bits [1:0] x (0, 1, 3=-1), bits [3:2] y
   The list counter is cleared by the border kernel (or ocg_xlist_reset_kernel),
   which always follows in stream order. */
__global__ void __launch_bounds__(OCG_RECON_THREADS, 6)
ocg_recon_xform_kernel(const OcgGeomDev g, const OcgJobDev *__restrict__ jobs, int njobs) {
  /* Every warp works on its own: 32 list entries per round, partitioned by footprint class inside the
     warp (ballots, stable), then four sub-rounds of 8 fragments x 4 lanes.  No CTA-wide barrier (the
     CTA-wide 64-entry partition this replaces spent most of its stall cycles in __syncthreads on
     dense-coefficient frames). */
  __shared__ int4 sitem[OCG_RECON_THREADS / 32][32];
  __shared__ unsigned char stap[OCG_RECON_THREADS / 32][32];
  const int t = (int)threadIdx.x, lane = t & 31, w = t >> 5;
  /* the grid is one resident wave; CTA b serves job b % njobs as that job's rank b / njobs */
  const int jobi = (int)blockIdx.x % njobs, rank = (int)blockIdx.x / njobs;
  const int nranks = ((int)gridDim.x - 1 - jobi) / njobs + 1;
  const OcgJobDev &job = jobs[jobi];
  pdl_wait(); /* pass A's list, counter and pixels */
  pdl_release();
  const int nx = *(volatile const int *)job.xcount;
  constexpr int NW = OCG_RECON_THREADS / 32;
  /* entries per warp and round: 32, or -- when the list is short for the warps that serve it -- just enough
     (a multiple of 8) that one round covers it: a short list is latency, not throughput */
  const int E = min(32, max(8, (((nx + nranks * NW - 1) / (nranks * NW)) + 7) & ~7));
  /* rounds are claimed from a per-job counter, not dealt statically: with programmatic dependent launch the
     CTAs of this grid become resident as the kernel before drains, unevenly over the SMs, and a static
     deal made the slowest SM's share the kernel's time (dense frames: 292 -> 342 us) */
  (void)rank;
  for (;;) {
    int e0 = 0;
    if (lane == 0) e0 = atomicAdd(job.xcount + 1, E);
    e0 = __shfl_sync(0xFFFFFFFFu, e0, 0);
    if (e0 >= nx) break;
    const int nvalid = min(E, nx - e0);
    int cls = WC_NONE;
    int4 it = make_int4(0, 0, 0, 0);
    unsigned tap = 0;
    if (lane < nvalid) {
      const int fragi = job.xlist[e0 + lane];
      const int4 rw = __ldg((const int4 *)(job.recs + fragi));
      cls = work_class(rw);
      const int refi = (rw.w >> 16) & 3, pli = (rw.w >> 24) & 3, qti = (rw.w >> 26) & 1;
      const int dcv = (rw.y >> 16) * (int)job.dcq[pli][qti];
      int off0 = 0, fx = 0, fy = 0;
      if (refi != OCG_FRAME_SELF) mv_taps(rw.y << 16 >> 16, pli ? g.qx : 0, pli ? g.qy : 0, g.p[pli].ystride, off0, fx, fy);
      it.x = rw.x;
      it.y = rw.x + off0;
      it.z = rw.z;
      it.w = (dcv & 0xFFFF) | ((rw.w & 0xFF) << 16) | (cls << 24) | (refi << 27) | (pli << 29);
      tap = (unsigned)(fx & 3) | ((unsigned)(fy & 3) << 2);
    }
    int base = 0, pos = 0;
#pragma unroll
    for (int c = WC_3; c < WC_COUNT; c++) {
      const unsigned m = __ballot_sync(0xFFFFFFFFu, cls == c);
      if (cls == c) pos = base + __popc(m & ((1u << lane) - 1u));
      base += __popc(m);
    }
    if (cls != WC_NONE) {
      sitem[w][pos] = it;
      stap[w][pos] = (unsigned char)tap;
    }
    __syncwarp();
#pragma unroll 1
    for (int r = 0; r < 32; r += 8) {
      if (r >= nvalid) break; /* warp-uniform */
      const int gi = r + (lane >> 2);
      int4 gt = make_int4(0, 0, 0, 0);
      unsigned gtap = 0;
      if (gi < nvalid) {
        gt = sitem[w][gi];
        gtap = stap[w][gi];
      }
      xform_group(g, job, gt, gtap, lane & 3);
    }
    __syncwarp(); /* the staging rows are rewritten by the next round */
  }
}

/* Clears the transform work-list counters of every job (stream-ordered after
   pass B); used when the border kernel, which normally does it, is not run. */
__global__ void ocg_xlist_reset_kernel(const OcgJobDev *__restrict__ jobs, int njobs) {
  const int j = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (j < njobs) { jobs[j].xcount[0] = 0; jobs[j].xcount[1] = 0; }
}

/* Only the coded map (for running the loop filter stage on its own). */
__global__ void __launch_bounds__(256)
ocg_codedmap_kernel(const OcgGeomDev g, const OcgJobDev *__restrict__ jobs) {
  const OcgJobDev &job = jobs[blockIdx.y];
  const int f = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (f < g.nfrags) job.coded[f] = (unsigned char)(job.recs[f].refi != OCG_FRAG_UNCODED);
}

/* ------------------------------------------------------------------------ */
/* Loop filter.  One thread per 8x8 "cell" centred on a fragment corner
   (pixel columns [8cx-4,8cx+4), rows [8cy-4,8cy+4) of a plane).  Every filter
   line of state.c:1002-1031 lies inside exactly one cell; the lines through
   the central 4x4 patch are applied in the order the raster scan of
   state.c:1083-1104 reaches them, everything else is independent. */

__device__ __forceinline__ int lflim(int r, int lim) {
  /* closed form of the table built by oc_loop_filter_init_c, state.c:1036 */
  const int a = abs(r);
  const int v = a < lim ? a : max(2 * lim - a, 0);
  return r < 0 ? -v : v;
}

/* Two filter lines at once, one per halfword (state.c:1002-1031).
   a,b,c,d: the four samples straddling the edge, packed as (line0, line1)
   16-bit pairs, 0..255 each.  Values stay below 2^16 per halfword, so plain
   32-bit integer arithmetic acts on both lines:
     F = a + 3c + 1028 - (d + 3b)  ==  (a-d+3(c-b)+4) + 1024   (>= 8 per half)
     F>>3                           ==  ((f+4)>>3) + 128        (1..256)
   The bounding function comes from a 257-entry table in shared memory. */
__device__ __forceinline__ void lf_pair(uint32_t a, uint32_t &b, uint32_t &c, uint32_t d, const signed char *bv) {
  const uint32_t x = c * 3u + a + 0x04040404u;
  const uint32_t y = b * 3u + d;
  const uint32_t u = ((x - y) >> 3) & 0x1FFF1FFFu;
  const int f0 = bv[u & 0xFFFFu], f1 = bv[u >> 16];
  const uint32_t ff = __byte_perm((uint32_t)f0, (uint32_t)f1, 0x5410);
  const uint32_t nf = __vneg2(ff);
  b = __viaddmin_s16x2_relu(b, ff, 0x00FF00FFu);
  c = __viaddmin_s16x2_relu(c, nf, 0x00FF00FFu);
}

struct Cell {
  uint32_t w[8][2]; /* w[row][0] = cols 0..3, w[row][1] = cols 4..7 (cell-local) */
};

/* Vertical edge through the cell centre, rows R and R+1: samples are cell
   columns 2,3 | 4,5 (loop_filter_h, state.c:1002). */
template <int R>
__device__ __forceinline__ void cell_vpair(Cell &c, const signed char *bv) {
  const uint32_t m = 0x00FF00FFu;
  const uint32_t a = __byte_perm(c.w[R][0], c.w[R + 1][0], 0x0602) & m; /* col 2 of both rows */
  uint32_t b = __byte_perm(c.w[R][0], c.w[R + 1][0], 0x0703) & m;       /* col 3 */
  uint32_t cc = __byte_perm(c.w[R][1], c.w[R + 1][1], 0x0400) & m;      /* col 4 */
  const uint32_t d = __byte_perm(c.w[R][1], c.w[R + 1][1], 0x0501) & m; /* col 5 */
  lf_pair(a, b, cc, d, bv);
  c.w[R][0] = __byte_perm(c.w[R][0], b, 0x4210);
  c.w[R + 1][0] = __byte_perm(c.w[R + 1][0], b, 0x6210);
  c.w[R][1] = __byte_perm(c.w[R][1], cc, 0x3214);
  c.w[R + 1][1] = __byte_perm(c.w[R + 1][1], cc, 0x3216);
}

/* Horizontal edge through the cell centre, cell columns COL and COL+1 (same
   register): samples are cell rows 2,3 | 4,5 (loop_filter_v, state.c:1018;
   bottom-up rows). */
template <int COL>
__device__ __forceinline__ void cell_hpair(Cell &c, const signed char *bv) {
  constexpr int h = COL >> 2, i = COL & 3; /* i is 0 or 2 */
  constexpr unsigned ext = i == 0 ? 0x4140u : 0x4342u; /* (byte i, 0, byte i+1, 0) with a zero second operand */
  const uint32_t a = __byte_perm(c.w[2][h], 0, ext);
  uint32_t b = __byte_perm(c.w[3][h], 0, ext);
  uint32_t cc = __byte_perm(c.w[4][h], 0, ext);
  const uint32_t d = __byte_perm(c.w[5][h], 0, ext);
  lf_pair(a, b, cc, d, bv);
  constexpr unsigned ins = i == 0 ? 0x3264u : 0x6410u; /* put bytes 0,2 of the result at i,i+1 */
  c.w[3][h] = __byte_perm(c.w[3][h], b, ins);
  c.w[4][h] = __byte_perm(c.w[4][h], cc, ins);
}

/* Bounding tables for every loop-filter limit, built once per device:
   g_lf_table[lim][u] = lflim(u-128, lim), u = 0..259 (260 B = 65 words/row). */
__device__ signed char g_lf_table[128][260];

__global__ void ocg_lf_table_kernel() {
  const int lim = (int)blockIdx.x;
  for (int u = (int)threadIdx.x; u < 260; u += (int)blockDim.x) g_lf_table[lim][u] = (signed char)lflim(u - 128, lim);
}

template <bool INTERIOR>
__device__ __forceinline__ void lf_cell(uint8_t *o, int ystride, const signed char *bv, bool inl, bool inr, bool ind,
                                        bool inu, bool vd, bool vu, bool hl, bool hr, bool B, bool Cc, bool D) {
  const int rlo = ind ? 0 : 4, rhi = inu ? 8 : 4;
  Cell c;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    if (INTERIOR) {
      const uint32_t *row = (const uint32_t *)(o + r * ystride);
      c.w[r][0] = row[0];
      c.w[r][1] = row[1];
    } else {
      c.w[r][0] = c.w[r][1] = 0;
      if (r >= rlo && r < rhi) {
        const uint32_t *row = (const uint32_t *)(o + r * ystride);
        if (inl) c.w[r][0] = row[0];
        if (inr) c.w[r][1] = row[1];
      }
    }
  }
  /* independent lines */
  if (vd) cell_vpair<0>(c, bv);
  if (vu) cell_vpair<6>(c, bv);
  if (hl) cell_hpair<0>(c, bv);
  if (hr) cell_hpair<6>(c, bv);
  /* ordered lines through the central patch ("loop filter") */
  if (vd && !B) cell_vpair<2>(c, bv);
  if (hl && !Cc) cell_hpair<2>(c, bv);
  if (vd && B) cell_vpair<2>(c, bv);
  if (hr && !D) cell_hpair<4>(c, bv);
  if (hl && Cc) cell_hpair<2>(c, bv);
  if (vu && !D) cell_vpair<4>(c, bv);
  if (vu && D) cell_vpair<4>(c, bv);
  if (hr && D) cell_hpair<4>(c, bv);
#pragma unroll
  for (int r = 0; r < 8; r++) {
    if (INTERIOR) {
      uint32_t *row = (uint32_t *)(o + r * ystride);
      row[0] = c.w[r][0];
      row[1] = c.w[r][1];
    } else if (r >= rlo && r < rhi) {
      uint32_t *row = (uint32_t *)(o + r * ystride);
      if (inl) row[0] = c.w[r][0];
      if (inr) row[1] = c.w[r][1];
    }
  }
}

#define OCG_LF_ROWS 1 /* cell rows per CTA (64 cells wide); 4 rows/CTA measured 10 % slower (tail + occupancy) */

__global__ void __launch_bounds__(64 * OCG_LF_ROWS)
ocg_lf_kernel(const OcgGeomDev g, const OcgJobDev *__restrict__ jobs) {
  __shared__ __align__(4) signed char bv[260];
  const OcgJobDev &job = jobs[blockIdx.z];
  const int lim = job.lf_limit;
  if (lim == 0) return;
  if (threadIdx.y == 0) {
    const uint32_t *src = (const uint32_t *)g_lf_table[lim];
    uint32_t *dst = (uint32_t *)bv;
    dst[threadIdx.x] = src[threadIdx.x];
    if (threadIdx.x == 0) dst[64] = src[64];
  }
  __syncthreads();
  const int crow = (int)blockIdx.y * OCG_LF_ROWS + (int)threadIdx.y;
  if (crow >= g.cell_rows) return;
  const int pli = crow >= g.p[2].cell_row0 ? 2 : (crow >= g.p[1].cell_row0 ? 1 : 0);
  const OcgPlaneDev &P = g.p[pli];
  const int cy = crow - P.cell_row0;
  const int cx = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int nh = P.nhfrags, nv = P.nvfrags;
  if (cx > nh) return;
  /* coded flags of the four fragments around the corner */
  const uint8_t *cm = job.coded + P.froffset;
  const bool inl = cx > 0, inr = cx < nh, ind = cy > 0, inu = cy < nv;
  const bool A = inl && ind && cm[(cy - 1) * nh + cx - 1];
  const bool B = inr && ind && cm[(cy - 1) * nh + cx];
  const bool Cc = inl && inu && cm[cy * nh + cx - 1];
  const bool D = inr && inu && cm[cy * nh + cx];
  const bool vd = inl && inr && ind && (A || B);   /* vertical edge below the corner   */
  const bool vu = inl && inr && inu && (Cc || D);  /* vertical edge above the corner   */
  const bool hl = ind && inu && inl && (A || Cc);  /* horizontal edge left of corner   */
  const bool hr = ind && inu && inr && (B || D);   /* horizontal edge right of corner  */
  if (!(vd || vu || hl || hr)) return;
  const int ystride = P.ystride;
  uint8_t *o = job.base[OCG_FRAME_SELF] + P.plane_off + (cy * 8 - 4) * ystride + (cx * 8 - 4);
  if (inl && inr && ind && inu) lf_cell<true>(o, ystride, bv, true, true, true, true, vd, vu, hl, hr, B, Cc, D);
  else lf_cell<false>(o, ystride, bv, inl, inr, ind, inu, vd, vu, hl, hr, B, Cc, D);
}

/* ---- loop filter, strip form (the default) ---------------------------------------------------------
   The one-cell-per-thread kernel above is bound by the ALU pipe (ncu: 67 % busy at 45 % of the DRAM rate).
   This form first trims instructions:
   * lines ACROSS a vertical edge (loop_filter_h, state.c:1002) have their four samples in four adjacent
     bytes of one pixel row: one PRMT funnels them into a word and one IDP.4A (dp4a, weights 4,-12,12,-4,
     bias 4*1020) yields 4*(a-d+3(c-b)+1020) -- directly the byte offset into a 2048-entry table in shared
     memory whose 32-bit entries hold the bounded correction as a halfword pair (+f for b, -f for c), so the
     whole update is PRMT, VIADDMNMX.S16x2.RELU and two PRMT inserts;
   * lines ALONG columns (loop_filter_v, state.c:1018) stay two-per-operation on halfword pairs, but index
     the same table (two LDS, two PRMT give the +f and -f pairs);
   * a thread walks R vertically adjacent cells, so the plane look-up and the address set-up are paid once
     per strip, and the coded flags of all R+1 fragment rows are requested up front (no flag -> row
     dependency per cell);
   * the eight lines through the central 4x4 patch are issued as six blocks instead of eight: only the
     relative order of a vertical and a horizontal group matters, and of those only two pairs are not fixed
     (left edge before the lower edge iff B and not C; right edge before the upper edge iff not D);
   * every active cell moves whole rows (the cells of the plane's rim reach into the apron, which is
     allocated, read and written back unchanged), so there is no second code path for the rim;
   * row addresses are one IMAD.WIDE each (FMA pipe) instead of eight hoisted 64-bit offsets: 40 registers.
   ALU pipe 67 % -> 37 %; 94 -> 78 us per 64-stream launch.  What bounds it now is the movement itself: a
   probe that only loads and stores the cells takes 86 us with these 32-bit accesses (70 us with 64-bit ones,
   which a cell -- 4 bytes off the 8-byte grid -- cannot use directly).  Tried and measured slower: exchanging
   halves between lanes by shuffle to move aligned 64-bit words (97-108 us; stores only: 87 us), one cell of prefetch (85 us at
   62 registers), fetching each cell row's 8 pixel rows as one linear cp.async.bulk block into a double-
   buffered shared-memory tile (88 us); CTA shapes from 32x4 to 256x1 cells and R = 1..8 all land within
   78-87 us. */
#define OCG_LF2_TAB 2048

/* g_lf_tab2[lim][i], i = a-d+3(c-b)+1020 in 0..2040: f = lflim((i-1016)>>3, lim) as (f & 0xFFFF) | (-f << 16) */
__device__ uint32_t g_lf_tab2[128][OCG_LF2_TAB];

__global__ void ocg_lf_tab2_kernel() {
  const int lim = (int)blockIdx.x;
  for (int i = (int)threadIdx.x; i < OCG_LF2_TAB; i += (int)blockDim.x) {
    const int f = lflim((i - 1016) >> 3, lim);
    g_lf_tab2[lim][i] = ((uint32_t)f & 0xFFFFu) | ((uint32_t)(-f) << 16);
  }
}

__device__ __forceinline__ uint32_t lf2_entry(const uint32_t *tab, uint32_t byte_off) {
  return *(const uint32_t *)((const unsigned char *)tab + byte_off);
}

/* one line across the vertical edge in the middle of a cell row: w0 = cell columns 0..3, w1 = 4..7 */
__device__ __forceinline__ void lf2_vline(uint32_t &w0, uint32_t &w1, const uint32_t *tab) {
  const uint32_t s = __byte_perm(w0, w1, 0x5432); /* columns 2,3,4,5 = a,b,c,d */
  int off;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(off) : "r"(s), "r"(0xFC0CF404u), "r"(4 * 1020));
  const uint32_t e = lf2_entry(tab, (uint32_t)off);
  const uint32_t bc = __byte_perm(s, 0, 0x4241); /* (b, c) as halfwords */
  const uint32_t r = __viaddmin_s16x2_relu(bc, e, 0x00FF00FFu);
  w0 = __byte_perm(w0, r, 0x4210);
  w1 = __byte_perm(w1, r, 0x3216);
}

/* two lines along cell columns COL, COL+1 across the horizontal edge in the middle of the cell (rows 2..5) */
template <int COL>
__device__ __forceinline__ void lf2_hpair(uint32_t (&w)[8][2], const uint32_t *tab) {
  constexpr int h = COL >> 2, i = COL & 3; /* i is 0 or 2 */
  constexpr unsigned ext = i == 0 ? 0x4140u : 0x4342u;
  const uint32_t a = __byte_perm(w[2][h], 0, ext);
  const uint32_t b = __byte_perm(w[3][h], 0, ext);
  const uint32_t c = __byte_perm(w[4][h], 0, ext);
  const uint32_t d = __byte_perm(w[5][h], 0, ext);
  /* per halfword 4*(a-d+3(c-b)+1020) in 0..8160: no borrow between the halves of the final value */
  const uint32_t v = (c * 12u + a * 4u + 0x0FF00FF0u) - (b * 12u + d * 4u);
  const uint32_t e0 = lf2_entry(tab, v & 0xFFFFu), e1 = lf2_entry(tab, v >> 16);
  const uint32_t pf = __byte_perm(e0, e1, 0x5410), nf = __byte_perm(e0, e1, 0x7632);
  const uint32_t b2 = __viaddmin_s16x2_relu(b, pf, 0x00FF00FFu);
  const uint32_t c2 = __viaddmin_s16x2_relu(c, nf, 0x00FF00FFu);
  constexpr unsigned ins = i == 0 ? 0x3264u : 0x6410u;
  w[3][h] = __byte_perm(w[3][h], b2, ins);
  w[4][h] = __byte_perm(w[4][h], c2, ins);
}

/* o + r*ystride as one IMAD.WIDE (FMA pipe) instead of a hoisted 64-bit offset per row (16 registers) */
__device__ __forceinline__ uint32_t *lf2_row(uint8_t *o, int ystride, int r) {
  unsigned long long a;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(a) : "r"(ystride), "r"(r), "l"((unsigned long long)o));
  return (uint32_t *)a;
}

/* which lines of a cell run, from the coded flags of the fragment rows below (AB) and above (CD) the corner:
   bit 0 vd, 1 vu, 2 hl, 3 hr (vertical edge below/above the corner, horizontal edge left/right of it) */
__device__ __forceinline__ unsigned lf2_lines(unsigned AB, unsigned CD, bool both, bool mid) {
  unsigned m = 0;
  if (both && AB != 0) m |= 1u;
  if (both && CD != 0) m |= 2u;
  if (mid && ((AB | CD) & 1u)) m |= 4u;
  if (mid && ((AB | CD) & 2u)) m |= 8u;
  return m;
}

__device__ __forceinline__ void lf2_load(uint32_t (&w)[8][2], uint8_t *o, int ystride, unsigned lines) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const bool need = lines != 0 && (r < 2 ? (lines & 1u) != 0 : (r >= 6 ? (lines & 2u) != 0 : true));
    w[r][0] = w[r][1] = 0;
    if (need) {
      const uint32_t *rw = lf2_row(o, ystride, r);
      w[r][0] = rw[0];
      w[r][1] = rw[1];
    }
  }
}

__device__ __forceinline__ void lf2_filter(uint32_t (&w)[8][2], unsigned lines, unsigned AB, unsigned CD, const uint32_t *tab) {
  const bool vd = lines & 1u, vu = lines & 2u, hl = lines & 4u, hr = lines & 8u;
  const bool B = (AB & 2u) != 0, Cc = (CD & 1u) != 0, D = (CD & 2u) != 0;
  /* lines that touch nothing another line touches */
  if (vd) { lf2_vline(w[0][0], w[0][1], tab); lf2_vline(w[1][0], w[1][1], tab); }
  if (vu) { lf2_vline(w[6][0], w[6][1], tab); lf2_vline(w[7][0], w[7][1], tab); }
  if (hl) lf2_hpair<0>(w, tab);
  if (hr) lf2_hpair<6>(w, tab);
  /* the central patch, in an order equivalent to the raster scan's (state.c:1083-1104) */
  const bool hl_first = B && !Cc;
  if (hl && hl_first) lf2_hpair<2>(w, tab);
  if (vd) { lf2_vline(w[2][0], w[2][1], tab); lf2_vline(w[3][0], w[3][1], tab); }
  if (hl && !hl_first) lf2_hpair<2>(w, tab);
  if (hr && !D) lf2_hpair<4>(w, tab);
  if (vu) { lf2_vline(w[4][0], w[4][1], tab); lf2_vline(w[5][0], w[5][1], tab); }
  if (hr && D) lf2_hpair<4>(w, tab);
}

__device__ __forceinline__ void lf2_filter_store(uint32_t (&w)[8][2], uint8_t *o, int ystride, unsigned lines, unsigned AB,
                                                 unsigned CD, const uint32_t *tab) {
  lf2_filter(w, lines, AB, CD, tab);
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const bool need = r < 2 ? (lines & 1u) != 0 : (r >= 6 ? (lines & 2u) != 0 : true);
    if (need) {
      uint32_t *rw = lf2_row(o, ystride, r);
      rw[0] = w[r][0];
      rw[1] = w[r][1];
    }
  }
}

template <int TX, int TY, int R, int MINB>
__global__ void __launch_bounds__(TX * TY, MINB)
ocg_lf2_kernel(const OcgGeomDev g, const OcgJobDev *__restrict__ jobs) {
  __shared__ __align__(16) uint32_t tab[OCG_LF2_TAB];
  const OcgJobDev &job = jobs[blockIdx.z];
  const int lim = job.lf_limit;
  if (lim == 0) return;
  /* strip group -> plane: groups of TY*R cell rows, aligned per plane */
  constexpr int GR = TY * R;
  int grp = (int)blockIdx.y, pli = 0;
  {
    const int n0 = (g.p[0].nvfrags + GR) / GR, n1 = (g.p[1].nvfrags + GR) / GR;
    if (grp >= n0) { grp -= n0; pli = 1; if (grp >= n1) { grp -= n1; pli = 2; } }
  }
  const OcgPlaneDev &P = g.p[pli];
  const int nh = P.nhfrags, nv = P.nvfrags;
  const int cx = (int)(blockIdx.x * TX + threadIdx.x);
  const int cy0 = grp * GR + (int)threadIdx.y * R;
  const bool live = cx <= nh && cy0 <= nv;
  const bool inl = cx > 0, inr = cx < nh;
  {
    /* the table does not depend on the frame: staged while the reconstruction kernels drain */
    const int tid = (int)(threadIdx.y * TX + threadIdx.x);
    const uint4 *src = (const uint4 *)g_lf_tab2[lim];
#pragma unroll
    for (int k = 0; k < OCG_LF2_TAB / 4 / (TX * TY); k++) ((uint4 *)tab)[tid + k * TX * TY] = src[tid + k * TX * TY];
  }
  pdl_wait(); /* the coded map and the pixels of the reconstruction kernels */
  pdl_release();
  /* coded flags of the R+1 fragment rows around the strip's corners, all requested up front: 2 bits per row
     (bit 0 left of the corner, bit 1 right of it) */
  unsigned coded = 0;
  if (live) {
    const uint8_t *cm = job.coded + P.froffset + cx;
    unsigned char fl[R + 1][2];
#pragma unroll
    for (int k = 0; k <= R; k++) {
      const int fy = cy0 - 1 + k;
      const bool in = fy >= 0 && fy < nv;
      fl[k][0] = in && inl ? cm[fy * nh - 1] : (unsigned char)0;
      fl[k][1] = in && inr ? cm[fy * nh] : (unsigned char)0;
    }
#pragma unroll
    for (int k = 0; k <= R; k++) coded |= ((fl[k][0] ? 1u : 0u) | (fl[k][1] ? 2u : 0u)) << (2 * k);
  }
  __syncthreads();
  if (!live) return;
  const int ystride = P.ystride;
  const bool both = inl && inr;
  uint8_t *o = job.base[OCG_FRAME_SELF] + P.plane_off + (ptrdiff_t)(cy0 * 8 - 4) * ystride + (cx * 8 - 4);
  const int n = min(R, nv + 1 - cy0);
#pragma unroll 1
  for (int i = 0; i < n; i++, o += 8 * (ptrdiff_t)ystride) {
    const int cy = cy0 + i;
    const unsigned AB = (coded >> (2 * i)) & 3u, CD = (coded >> (2 * i + 2)) & 3u;
    const unsigned lines = lf2_lines(AB, CD, both, cy > 0 && cy < nv);
    if (lines == 0) continue;
    uint32_t w[8][2];
    lf2_load(w, o, ystride, lines);
    lf2_filter_store(w, o, ystride, lines, AB, CD, tab);
  }
}

/* ---- TMA variant ----------------------------------------------------------
   The cells of one cell row are self-contained 8x8 pixel squares, so the input
   of a CTA's 64 cells is one tile of the padded plane: thread 0 fetches it with
   a single cp.async.bulk.tensor.2d (TMA) into shared memory, completion is
   signalled on an mbarrier, and every thread reads its cell out of shared
   memory -- no per-thread address arithmetic or predicated global loads, and
   plane edges need no special casing because the tile lies inside the apron.
   TMA requires the box to start on a 16-byte boundary in global memory
   (measured: tools/micro/tma_test.cu faults otherwise) while cells start at
   8cx-4, so the box is 528 bytes wide (12 bytes of lead-in + 64 cells + 4) and
   boxes of neighbouring CTAs overlap by 16 bytes; for the same reason the
   results are written back with per-thread 32-bit stores rather than a TMA
   store (which would also write the overlap).  Tiles whose 64 cells touch no
   coded fragment are skipped before the fetch. */
#define OCG_LF_BOX_WORDS 132

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(64)
ocg_lf_tma_kernel(const OcgGeomDev g, const OcgJobDev *__restrict__ jobs) {
  __shared__ __align__(128) uint32_t tile[8][OCG_LF_BOX_WORDS]; /* [memory row][word] */
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ __align__(4) signed char bv[260];
  const OcgJobDev &job = jobs[blockIdx.z];
  const int lim = job.lf_limit;
  if (lim == 0) return;
  const int tid = (int)threadIdx.x;
  const int crow = (int)blockIdx.y;
  const int pli = crow >= g.p[2].cell_row0 ? 2 : (crow >= g.p[1].cell_row0 ? 1 : 0);
  const OcgPlaneDev &P = g.p[pli];
  const int cy = crow - P.cell_row0;
  const int cx0 = (int)blockIdx.x * 64;
  const int cx = cx0 + tid;
  const int nh = P.nhfrags, nv = P.nvfrags;
  if (cx0 > nh) return; /* whole CTA beyond this plane's last cell */
  /* coded flags of the four fragments around the corner */
  const uint8_t *cm = job.coded + P.froffset;
  const bool inl = cx > 0 && cx <= nh, inr = cx < nh, ind = cy > 0, inu = cy < nv;
  const bool A = inl && ind && cm[(cy - 1) * nh + cx - 1];
  const bool B = inr && ind && cm[(cy - 1) * nh + cx];
  const bool Cc = inl && inu && cm[cy * nh + cx - 1];
  const bool D = inr && inu && cm[cy * nh + cx];
  const bool vd = inl && inr && ind && (A || B);
  const bool vu = inl && inr && inu && (Cc || D);
  const bool hl = ind && inu && inl && (A || Cc);
  const bool hr = ind && inu && inr && (B || D);
  const bool active = vd || vu || hl || hr;
  {
    const uint32_t *src = (const uint32_t *)g_lf_table[lim];
    uint32_t *dst = (uint32_t *)bv;
    dst[tid] = src[tid];
    if (tid == 0) {
      dst[64] = src[64];
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  if (!__syncthreads_or(active)) return;
  /* box origin in tensor coordinates (32-bit words / top-down memory rows): the
     tensor starts 16 bytes left of the picture and vpad rows above it, so the
     box starts at picture x = 8*cx0 - 16, a multiple of 16 bytes */
  const int tx = 2 * cx0;
  const int ty = P.vpad + P.height - 8 * cy - 4;
  if (tid == 0) {
    const CUtensorMap *tm = job.lf_tmaps + pli;
    /* the tensor map lives in global memory (written by the host): acquire it
       for the tensormap proxy before the TMA unit reads it */
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)),
                 "r"((unsigned)(OCG_LF_BOX_WORDS * 4 * 8))
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(&tile[0][0])),
        "l"(tm), "r"(tx), "r"(ty), "r"(smem_u32(&mbar))
        : "memory");
  }
  if (!active) return; /* idle cells need not wait for the tile */
  {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(0u)
          : "memory");
    }
  }
  Cell c;
  /* cell row r (bottom-up) is memory row 7-r of the tile; the cell starts 12 bytes into the box */
#pragma unroll
  for (int r = 0; r < 8; r++) {
    c.w[r][0] = tile[7 - r][3 + 2 * tid];
    c.w[r][1] = tile[7 - r][4 + 2 * tid];
  }
  if (vd) cell_vpair<0>(c, bv);
  if (vu) cell_vpair<6>(c, bv);
  if (hl) cell_hpair<0>(c, bv);
  if (hr) cell_hpair<6>(c, bv);
  if (vd && !B) cell_vpair<2>(c, bv);
  if (hl && !Cc) cell_hpair<2>(c, bv);
  if (vd && B) cell_vpair<2>(c, bv);
  if (hr && !D) cell_hpair<4>(c, bv);
  if (hl && Cc) cell_hpair<2>(c, bv);
  if (vu && !D) cell_vpair<4>(c, bv);
  if (vu && D) cell_vpair<4>(c, bv);
  if (hr && D) cell_hpair<4>(c, bv);
  /* write back the parts of the cell that lie inside the plane */
  const int ystride = P.ystride;
  uint8_t *o = job.base[OCG_FRAME_SELF] + P.plane_off + (cy * 8 - 4) * ystride + (cx * 8 - 4);
  const int rlo = ind ? 0 : 4, rhi = inu ? 8 : 4;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    if (r >= rlo && r < rhi) {
      uint32_t *row = (uint32_t *)(o + r * ystride);
      if (inl) row[0] = c.w[r][0];
      if (inr) row[1] = c.w[r][1];
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Apron replication, state.c:770-835: every apron byte takes the nearest
   picture pixel (rows first, then full-width caps == clamp in both axes).
   Items are enumerated linearly per job:
     side items  every picture row x {left,right}: one thread reads the edge pixel and writes the whole
                 apron row (hpad bytes)
     cap items   every 8-byte column chunk of the padded width x {below,above}: one thread reads its 8 source
                 bytes of the edge row once and writes all vpad apron rows */
__global__ void __launch_bounds__(256)
ocg_border_kernel(const OcgGeomDev g, const OcgJobDev *__restrict__ jobs) {
  const OcgJobDev &job = jobs[blockIdx.y];
  int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  pdl_wait(); /* the filtered picture; pass B is done with the work list */
  pdl_release();
  if (t == 0) { job.xcount[0] = 0; job.xcount[1] = 0; } /* recon pass B is done with this frame's work list and its claim counter */
#pragma unroll
  for (int pli = 0; pli < 3; pli++) {
    const OcgPlaneDev &P = g.p[pli];
    const int nside = P.height * 2;
    const int capw = (P.width + 2 * P.hpad) >> 3;
    const int ncap = 2 * capw;
    uint8_t *base = job.base[OCG_FRAME_SELF] + P.plane_off;
    if (t < nside) {
      const int y = t >> 1;
      const bool right = (t & 1) != 0;
      uint8_t *srow = base + (ptrdiff_t)y * P.ystride;
      const uint32_t v = 0x01010101u * (right ? srow[P.width - 1] : srow[0]);
      uint2 *d = (uint2 *)(srow + (right ? P.width : -P.hpad));
      for (int k = 0; k < (P.hpad >> 3); k++) d[k] = make_uint2(v, v);
      return;
    }
    t -= nside;
    if (t < ncap) {
      const bool above = t >= capw;
      const int c = above ? t - capw : t;
      const uint8_t *srow = base + (ptrdiff_t)(above ? P.height - 1 : 0) * P.ystride;
      const int x = -P.hpad + 8 * c;
      uint2 v;
      if (x < 0) { const uint32_t e = 0x01010101u * srow[0]; v = make_uint2(e, e); }
      else if (x >= P.width) { const uint32_t e = 0x01010101u * srow[P.width - 1]; v = make_uint2(e, e); }
      else v = *(const uint2 *)(srow + x);
      /* rows -1..-vpad below the picture, rows height..height+vpad-1 above it */
      uint8_t *d = base + (ptrdiff_t)(above ? P.height : -1) * P.ystride + x;
      const ptrdiff_t step = above ? P.ystride : -(ptrdiff_t)P.ystride;
      for (int r = 0; r < P.vpad; r++) *(uint2 *)(d + r * step) = v;
      return;
    }
    t -= ncap;
  }
}

/* ------------------------------------------------------------------------ */
/* DC un-prediction (oc_dec_dc_unpredict_mcu_plane_c, decode.c:1392-1500) for a
   whole plane.  The recurrence dc = residual + pred(left, up-left, up,
   up-right | last value of the same reference type) is serial along a row and
   skewed by two columns between rows, so one CTA per plane runs a wave-front
   with ONE THREAD PER FRAGMENT ROW: in each step a row skips its uncoded
   fragments and finishes at most one coded fragment, as soon as the row above
   has passed column x+1; rows publish their progress through a double-buffered
   shared array, one __syncthreads_or per step.  The "no neighbour of my
   reference type" case (pred_last, decode.c:1452) points at the last coded
   fragment of that type in raster order, which is known before any value is:
   per-row last positions are tabulated first and the row simply waits until
   that particular fragment is final.  Reference types and (where the plane
   fits) the DC values live in shared memory; results go straight into the
   records the reconstruction kernels read. */
#define OCG_DC_MAX_ROWS 1024

/* Inputs/outputs of one DC job: either the fragment records (dc rewritten in place), or -- when the kernel
   runs ahead of the frame's lists -- the decoder's packed fragment words (oc_fragment, state.h:297-322:
   bit 0 coded, bits 6-7 refi, bits 16-31 dc) and a separate array of final values. */
struct OcgDcArgs {
  ocg_frag_rec *recs;
  const uint32_t *words;
  int16_t *dc_final;
  int16_t *dc_tmp;
};

template <bool DC_IN_SMEM>
__global__ void __launch_bounds__(OCG_DC_MAX_ROWS)
ocg_dc_unpredict_kernel(const OcgGeomDev g, const OcgJobDev *__restrict__ jobs, const OcgDcArgs single) {
  OcgDcArgs A = single;
  if (jobs != nullptr) {
    const OcgJobDev &job = jobs[blockIdx.y];
    if (job.dc_residual != 1) return;
    A.recs = const_cast<ocg_frag_rec *>(job.recs);
    A.words = nullptr;
    A.dc_final = nullptr;
    A.dc_tmp = job.dc_tmp;
  }
  const OcgPlaneDev &P = g.p[blockIdx.x];
  const int nh = P.nhfrags, nv = P.nvfrags, nfr = nh * nv;
  ocg_frag_rec *recs = A.recs != nullptr ? A.recs + P.froffset : nullptr;
  const uint32_t *words = A.words != nullptr ? A.words + P.froffset : nullptr;
  int16_t *dc_final = A.dc_final != nullptr ? A.dc_final + P.froffset : nullptr;
  extern __shared__ __align__(16) unsigned char smem[];
  int *progress = (int *)smem;                 /* [2][nv] */
  int *rowlast = progress + 2 * nv;            /* [3][nv] last coded x of each reference type in a row, or -1 */
  int *prevrow = rowlast + 3 * nv;             /* [3][nv] nearest earlier row that has one, or -1 */
  uint8_t *refs = (uint8_t *)(prevrow + 3 * nv); /* [nfr] 0..2 = reference type, 3 = not coded */
  volatile int16_t *dcs = DC_IN_SMEM ? (volatile int16_t *)(refs + ((nfr + 15) & ~15))
                                     : (volatile int16_t *)(A.dc_tmp + P.froffset);
  const int y = (int)threadIdx.x;
  for (int i = (int)threadIdx.x; i < nfr; i += (int)blockDim.x) {
    if (words != nullptr) {
      const uint32_t w = words[i];
      refs[i] = (uint8_t)((w & 1u) ? ((w >> 6) & 3u) : 3u);
      dcs[i] = (int16_t)(w >> 16);
    } else {
      const uint2 lo = *(const uint2 *)(recs + i); /* buf_off, mv|dc<<16 */
      const uint32_t hi = ((const uint32_t *)(recs + i))[3]; /* rowmask, last_zzi, refi, pli_qti */
      refs[i] = (uint8_t)((hi >> 16) & 0xFFu);
      dcs[i] = (int16_t)(lo.y >> 16);
    }
  }
  if (y < nv) { progress[y] = 0; progress[nv + y] = 0; }
  __syncthreads();
  if (y < nv) {
    int l0 = -1, l1 = -1, l2 = -1;
    const uint8_t *rr = refs + y * nh;
    for (int x = 0; x < nh; x++) {
      const int r = rr[x];
      l0 = r == 0 ? x : l0;
      l1 = r == 1 ? x : l1;
      l2 = r == 2 ? x : l2;
    }
    rowlast[y] = l0;
    rowlast[nv + y] = l1;
    rowlast[2 * nv + y] = l2;
  }
  __syncthreads();
  if (y < 3) {
    int last = -1;
    for (int r = 0; r < nv; r++) {
      prevrow[y * nv + r] = last;
      if (rowlast[y * nv + r] >= 0) last = r;
    }
  }
  __syncthreads();
  int x = 0, l_ref = -1, l_dc = 0, hasmask = 0, last0 = 0, last1 = 0, last2 = 0;
  const bool active = y < nv;
  const uint8_t *myref = refs + y * nh;
  const uint8_t *upref = refs + (y - 1) * nh;
  for (int it = 0;; it++) {
    const int *pcur = progress + (it & 1) * nv;
    int *pnext = progress + ((it & 1) ^ 1) * nv;
    bool more = false;
    if (active) {
      while (x < nh && myref[x] == 3) { x++; l_ref = -1; }
      if (x < nh) {
        const int r = myref[x];
        const int i = y * nh + x;
        const int mine = r == 0 ? last0 : (r == 1 ? last1 : last2);
        bool ready = true;
        int pred = 0;
        if (y == 0) pred = mine; /* decode.c:1415-1425 (0 until the first one) */
        else {
          ready = pcur[y - 1] >= min(x + 2, nh);
          if (ready) {
            const int ul = x > 0 ? upref[x - 1] : 255, u = upref[x], ur = x + 1 < nh ? upref[x + 1] : 255;
            const int pat = (l_ref == r) | (ul == r) << 1 | (u == r) << 2 | (ur == r) << 3;
            const int iu = i - nh;
            switch (pat) {
              case 0:
                if (hasmask >> r & 1) pred = mine;
                else {
                  const int yp = prevrow[r * nv + y];
                  if (yp >= 0) {
                    const int xs = rowlast[r * nv + yp];
                    if (pcur[yp] > xs) pred = dcs[yp * nh + xs];
                    else ready = false;
                  }
                }
                break;
              case 1: case 3: pred = l_dc; break;
              case 2: pred = dcs[iu - 1]; break;
              case 4: case 6: case 12: pred = dcs[iu]; break;
              case 5: pred = (l_dc + dcs[iu]) / 2; break;
              case 8: pred = dcs[iu + 1]; break;
              case 9: case 11: case 13: pred = (75 * l_dc + 53 * dcs[iu + 1]) / 128; break;
              case 10: pred = (dcs[iu - 1] + dcs[iu + 1]) / 2; break;
              case 14: pred = (3 * (dcs[iu - 1] + dcs[iu + 1]) + 10 * dcs[iu]) / 16; break;
              default: { /* 7, 15 */
                const int p0 = l_dc, p1 = dcs[iu - 1], p2 = dcs[iu];
                pred = (29 * (p0 + p2) - 26 * p1) / 32;
                if (abs(pred - p2) > 128) pred = p2;
                else if (abs(pred - p0) > 128) pred = p0;
                else if (abs(pred - p1) > 128) pred = p1;
              } break;
            }
          }
        }
        if (ready) {
          const int v = (int)(int16_t)(dcs[i] + pred); /* frags[].dc is a 16-bit field (state.h:320) */
          dcs[i] = (int16_t)v;
          if (dc_final != nullptr) dc_final[i] = (int16_t)v;
          else recs[i].dc = (int16_t)v;
          if (r == 0) last0 = v; else if (r == 1) last1 = v; else last2 = v;
          hasmask |= 1 << r;
          l_ref = r;
          l_dc = v;
          x++;
          while (x < nh && myref[x] == 3) { x++; l_ref = -1; }
        }
      }
      pnext[y] = x;
      more = x < nh;
    }
    if (!__syncthreads_or(more)) break;
  }
}

/* final DC values computed ahead of the lists (ocg_dec_dc_begin) -> the records of coded fragments */
__global__ void __launch_bounds__(256)
ocg_dc_patch_kernel(ocg_frag_rec *__restrict__ recs, const int16_t *__restrict__ dc_final, int nfrags) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i < nfrags && recs[i].refi != OCG_FRAG_UNCODED) recs[i].dc = dc_final[i];
}

} /* namespace */

/* launch with the programmatic-stream-serialisation attribute (see pdl_wait above) */
template <typename... KArgs, typename... Args>
static void ocg_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  static const int no_pdl = getenv("OCG_NO_PDL") != nullptr; /* A/B switch, read once */
  cfg.numAttrs = no_pdl ? 0 : 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

void ocg_launch_recon(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return;
  dim3 ga((unsigned)((g.nfrags + OCG_SIMPLE_THREADS - 1) / OCG_SIMPLE_THREADS), (unsigned)njobs);
  /* 4 rows per step at 12 CTAs per SM: 8 rows per step (64 registers, or 40-48 with spills) and 2 rows per step
     were measured 1-20 % slower */
  ocg_launch_pdl(ocg_recon_simple_kernel<4, 12>, ga, dim3(OCG_SIMPLE_THREADS), 0, st, g, jobs);
  /* pass B: one resident wave (6 CTAs of 256 threads per SM), CTA b striding over the list of job b % njobs;
     never more CTAs than there can be work for */
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  const int tiles = (g.nfrags + OCG_RECON_THREADS - 1) / OCG_RECON_THREADS; /* a CTA takes 256 entries per round */
  long nb = (long)sms * 6;
  if (nb < njobs) nb = njobs;
  if (nb > (long)tiles * njobs) nb = (long)tiles * njobs;
  ocg_launch_pdl(ocg_recon_xform_kernel, dim3((unsigned)nb), dim3(OCG_RECON_THREADS), 0, st, g, jobs, njobs);
  ocg_count_launch(2);
}

/* shared memory the DC kernel needs for a plane; dc_in_smem says whether the values fit too */
static size_t dc_smem_bytes(const OcgPlaneDev &P, bool dc_in_smem) {
  const size_t nv = (size_t)P.nvfrags, nfr = (size_t)P.nhfrags * P.nvfrags;
  return 8 * nv * sizeof(int) + ((nfr + 15) & ~(size_t)15) + (dc_in_smem ? nfr * 2 : 0);
}

static int dc_launch(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, const OcgDcArgs &single, cudaStream_t st) {
  static const size_t kMaxSmem = 227 * 1024;
  size_t need_all = 0, need_refs = 0;
  int rows = 0;
  for (int pli = 0; pli < 3; pli++) {
    need_all = std::max(need_all, dc_smem_bytes(g.p[pli], true));
    need_refs = std::max(need_refs, dc_smem_bytes(g.p[pli], false));
    rows = std::max(rows, (int)g.p[pli].nvfrags);
  }
  if (rows > OCG_DC_MAX_ROWS || need_refs > kMaxSmem) return -1; /* frame too tall / plane too large for this kernel */
  const unsigned threads = (unsigned)((rows + 31) & ~31);
  dim3 grid(3, (unsigned)njobs);
  if (need_all <= kMaxSmem) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(ocg_dc_unpredict_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem); attr = true; }
    ocg_dc_unpredict_kernel<true><<<grid, threads, need_all, st>>>(g, jobs, single);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(ocg_dc_unpredict_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem); attr = true; }
    ocg_dc_unpredict_kernel<false><<<grid, threads, need_refs, st>>>(g, jobs, single);
  }
  ocg_count_launch(1);
  return 0;
}

int ocg_launch_dc_unpredict(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return 0;
  return dc_launch(g, jobs, njobs, OcgDcArgs{nullptr, nullptr, nullptr, nullptr}, st);
}

int ocg_launch_dc_unpredict_words(const OcgGeomDev &g, const uint32_t *words, int16_t *dc_final, int16_t *dc_tmp,
                                  cudaStream_t st) {
  return dc_launch(g, nullptr, 1, OcgDcArgs{nullptr, words, dc_final, dc_tmp}, st);
}

void ocg_launch_dc_patch(ocg_frag_rec *recs, const int16_t *dc_final, int nfrags, cudaStream_t st) {
  ocg_dc_patch_kernel<<<(unsigned)((nfrags + 255) / 256), 256, 0, st>>>(recs, dc_final, nfrags);
  ocg_count_launch(1);
}

void ocg_launch_xlist_reset(const OcgJobDev *jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return;
  ocg_xlist_reset_kernel<<<(unsigned)((njobs + 127) / 128), 128, 0, st>>>(jobs, njobs);
  ocg_count_launch(1);
}

void ocg_launch_codedmap(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return;
  dim3 grid((unsigned)((g.nfrags + 255) / 256), (unsigned)njobs);
  ocg_codedmap_kernel<<<grid, 256, 0, st>>>(g, jobs);
  ocg_count_launch(1);
}

/* ---- flush graph: stage-in and copy-out without the copy engines -----------
   A graph made of kernels only is one cheap submission; every memcpy node in it is a separate one, and
   with a stream thread per decoder the driver's submission rate (not the GPU, not PCIe) was what capped
   the end-to-end frame rate.  So the lists come in and the picture goes out through mapped host memory. */
__device__ __forceinline__ uint4 ld_host16(const void *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

/* records (nfrags x 16 B), the frame's coefficient rows and the job header: mapped host memory -> device */
__global__ void __launch_bounds__(256)
ocg_stage_in_kernel(const OcgJobDev *__restrict__ h_job, OcgJobDev *__restrict__ d_job, const uint4 *__restrict__ h_recs,
                    uint4 *__restrict__ d_recs, int nfrags, const uint4 *__restrict__ h_rows, uint4 *__restrict__ d_rows) {
  const int nrows = h_job->ncoeff_rows;
  const int total = nfrags + nrows;
  const int stride = (int)(gridDim.x * blockDim.x);
  int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  /* four independent 16-byte reads in flight per thread: PCIe latency is ~1-2 us */
  for (; i + 3 * stride < total; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int j = i + k * stride;
      v[k] = j < nfrags ? ld_host16(h_recs + j) : ld_host16(h_rows + (j - nfrags));
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int j = i + k * stride;
      if (j < nfrags) d_recs[j] = v[k];
      else d_rows[j - nfrags] = v[k];
    }
  }
  for (; i < total; i += stride) {
    if (i < nfrags) d_recs[i] = ld_host16(h_recs + i);
    else d_rows[i - nfrags] = ld_host16(h_rows + (i - nfrags));
  }
  if (blockIdx.x == 0 && threadIdx.x < sizeof(OcgJobDev) / 8) {
    static_assert(sizeof(OcgJobDev) % 8 == 0, "job header is copied in 8-byte words");
    ((uint64_t *)d_job)[threadIdx.x] = ((const uint64_t *)h_job)[threadIdx.x];
  }
}

void ocg_launch_stage_in(const OcgJobDev *h_job, OcgJobDev *d_job, const ocg_frag_rec *h_recs, ocg_frag_rec *d_recs,
                         int nfrags, const int16_t *h_rows, int16_t *d_rows, cudaStream_t st) {
  static_assert(sizeof(OcgJobDev) / 8 <= 256, "job header larger than one CTA copies");
  ocg_stage_in_kernel<<<64, 256, 0, st>>>(h_job, d_job, (const uint4 *)h_recs, (uint4 *)d_recs, nfrags,
                                          (const uint4 *)h_rows, (uint4 *)d_rows);
  ocg_count_launch(1);
}

/* The finished frame -> mapped host memory (same layout on both sides), then the frame's sequence number
   into the host flag: every CTA fences its stores system-wide, the last one to finish publishes. */
struct OcgOutRect { int64_t off; int32_t pitch, width, height; }; /* width in bytes, multiple of 8 */
struct OcgOutPlan { OcgOutRect r[3]; int32_t nrect; int64_t base_off; };

template <typename V>
__device__ __forceinline__ void copy_rect(const uint8_t *src, uint8_t *dst, const OcgOutRect &r, int tid, int nthreads) {
  const int per_row = r.width / (int)sizeof(V);
  const int total = per_row * r.height;
  int i = tid;
  for (; i + 3 * nthreads < total; i += 4 * nthreads) {
    V v[4];
    int64_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int j = i + k * nthreads;
      const int y = j / per_row, x = j - y * per_row;
      o[k] = r.off + (int64_t)y * r.pitch + (int64_t)x * (int)sizeof(V);
      v[k] = *(const V *)(src + o[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) *(V *)(dst + o[k]) = v[k];
  }
  for (; i < total; i += nthreads) {
    const int y = i / per_row, x = i - y * per_row;
    const int64_t o = r.off + (int64_t)y * r.pitch + (int64_t)x * (int)sizeof(V);
    *(V *)(dst + o) = *(const V *)(src + o);
  }
}

__global__ void __launch_bounds__(256)
ocg_copy_out_kernel(const OcgOutPlan plan, const OcgJobDev *__restrict__ job, uint32_t *counter, uint32_t *host_flag) {
  pdl_wait(); /* the finished frame (and, in the token path, the job header staged at the head of the chain) */
  /* source and destination travel in the job header, so one graph serves every SELF buffer */
  const uint8_t *__restrict__ src = job->base[OCG_FRAME_SELF] - plan.base_off;
  uint8_t *__restrict__ host_dst = job->host_out;
  const int tid = (int)(blockIdx.x * blockDim.x + threadIdx.x), nthreads = (int)(gridDim.x * blockDim.x);
  for (int k = 0; k < plan.nrect; k++) {
    if ((plan.r[k].width & 15) == 0) copy_rect<uint4>(src, host_dst, plan.r[k], tid, nthreads);
    else copy_rect<uint2>(src, host_dst, plan.r[k], tid, nthreads);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(counter, 1u) + 1u == gridDim.x) {
      *counter = 0;
      __threadfence_system();
      *(volatile uint32_t *)host_flag = job->seq;
      __threadfence_system();
    }
  }
}

/* out_mode: OCG_OUT_PICTURE / OCG_OUT_PADDED / OCG_OUT_NONE (flag only) */
void ocg_launch_copy_out(const ocg_geometry &g, int out_mode, const OcgJobDev *job, uint32_t *counter, uint32_t *host_flag,
                         cudaStream_t st) {
  if (out_mode == 3) return; /* OCG_OUT_DEFERRED (internal): the caller moves the picture itself and flags afterwards */
  OcgOutPlan plan;
  memset(&plan, 0, sizeof(plan));
  plan.base_off = g.base_off;
  if (out_mode == OCG_OUT_PADDED) {
    plan.nrect = 1;
    plan.r[0].off = 0;
    plan.r[0].pitch = 0;
    plan.r[0].width = (int32_t)g.ref_frame_sz; /* multiple of 16 (state.c:582) */
    plan.r[0].height = 1;
  } else if (out_mode == OCG_OUT_PICTURE) {
    plan.nrect = 3;
    for (int pli = 0; pli < 3; pli++) {
      const ocg_plane_geom &p = g.planes[pli];
      plan.r[pli].off = g.base_off + p.plane_off + (int64_t)(p.height - 1) * p.ystride; /* top-left pixel */
      plan.r[pli].pitch = -p.ystride;
      plan.r[pli].width = p.width;
      plan.r[pli].height = p.height;
    }
  }
  const unsigned grid = out_mode == OCG_OUT_NONE ? 1u : 96u;
  ocg_launch_pdl(ocg_copy_out_kernel, dim3(grid), dim3(256), 0, st, plan, job, counter, host_flag);
  ocg_count_launch(1);
}

void ocg_init_device_tables(cudaStream_t st) {
  ocg_lf_table_kernel<<<128, 128, 0, st>>>();
  ocg_lf_tab2_kernel<<<128, 256, 0, st>>>();
}

int g_ocg_lf_legacy = 0; /* A/B switch (ocg_set_lf_tma(2)): the one-cell-per-thread kernel */

void ocg_launch_loop_filter(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, bool use_tma, cudaStream_t st) {
  if (njobs <= 0) return;
  const int variant = use_tma ? 1 : (g_ocg_lf_legacy ? 2 : 0);
  if (variant == 0) {
    constexpr int TX = 128, R = 4; /* 128 cells x 4 cell rows per CTA: the best of the shapes tried */
    int groups = 0;
    for (int pli = 0; pli < 3; pli++) groups += (g.p[pli].nvfrags + R) / R;
    dim3 grid((unsigned)((g.max_cells_x + TX - 1) / TX), (unsigned)groups, (unsigned)njobs);
    ocg_launch_pdl(ocg_lf2_kernel<TX, 1, R, 12>, grid, dim3(TX, 1), 0, st, g, jobs);
    ocg_count_launch(1);
    return;
  }
  if (variant == 1) {
    dim3 grid((unsigned)((g.max_cells_x + 63) / 64), (unsigned)g.cell_rows, (unsigned)njobs);
    ocg_lf_tma_kernel<<<grid, 64, 0, st>>>(g, jobs);
    ocg_count_launch(1);
    return;
  }
  dim3 grid((unsigned)((g.max_cells_x + 63) / 64), (unsigned)((g.cell_rows + OCG_LF_ROWS - 1) / OCG_LF_ROWS),
            (unsigned)njobs);
  ocg_lf_kernel<<<grid, dim3(64, OCG_LF_ROWS), 0, st>>>(g, jobs);
  ocg_count_launch(1);
}

void ocg_launch_borders(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return;
  int items = 0;
  for (int pli = 0; pli < 3; pli++) items += g.p[pli].height * 2 + 2 * ((g.p[pli].width + 2 * g.p[pli].hpad) >> 3);
  dim3 grid((unsigned)((items + 255) / 256), (unsigned)njobs);
  ocg_launch_pdl(ocg_border_kernel, grid, dim3(256), 0, st, g, jobs);
  ocg_count_launch(1);
}
