/* Encode-side batch kernels for sm_100a (reference lib/encint.h:292-326).
 *
 *  ocg_enc_metrics_kernel     SAD / SAD2, SATD / SATD2, intra SATD, SSD and
 *                             intra SAD of 8x8 blocks (encfrag.c:42-366), one
 *                             lane per block: VABSDIFF4 for the SADs, IDP.4A
 *                             for the SSD, a register-resident 8x8 Hadamard
 *                             for the SATDs.
 *  ocg_enc_fdct_quant_kernel  frag_sub / sub_128 / copy2+sub (encfrag.c:21-40,
 *                             368) -> oc_enc_fdct8x8 (fdct.c:128) ->
 *                             oc_enc_quantize (enquant.c:220), zig-zag order
 *                             output staged through shared memory so global
 *                             stores are 128-bit.
 * The serial mode decision / tokeniser that consumes these results stays on
 * the host (analyze.c, tokenize.c).
 */
#include <climits>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include "ocg_internal.h"

namespace {

constexpr int K1 = 64277, K2 = 60547, K3 = 54491, K5 = 36410, K6 = 25080, K7 = 12785;

__device__ __forceinline__ uint2 ld8u(const uint8_t *p) {
  const uintptr_t a = (uintptr_t)p;
  const uint2 *w = (const uint2 *)(a & ~(uintptr_t)7);
  const unsigned sh = (unsigned)(a & 7);
  const uint2 w0 = __ldg(w);
  if (sh == 0) return w0;
  const uint2 w1 = __ldg(w + 1);
  const unsigned sel = 0x3210u + 0x1111u * (sh & 3);
  uint2 r;
  if (sh < 4) { r.x = __byte_perm(w0.x, w0.y, sel); r.y = __byte_perm(w0.y, w1.x, sel); }
  else { r.x = __byte_perm(w0.y, w1.x, sel); r.y = __byte_perm(w1.x, w1.y, sel); }
  return r;
}

__device__ __forceinline__ void unpack8(uint2 v, int (&o)[8]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    o[i] = (int)((v.x >> (8 * i)) & 0xFF);
    o[4 + i] = (int)((v.y >> (8 * i)) & 0xFF);
  }
}

/* predictor row: none (0), one tap, or (a+b)>>1 of two taps */
__device__ __forceinline__ bool load_pred_row(const uint8_t *ref_base, const ocg_enc_frag &f, int row, int ystride,
                                              uint2 &pred) {
  if (f.ref_off0 == INT_MIN) { pred = make_uint2(0, 0); return false; }
  pred = ld8u(ref_base + f.ref_off0 + row * ystride);
  if (f.ref_off1 != INT_MIN) {
    const uint2 t = ld8u(ref_base + f.ref_off1 + row * ystride);
    pred.x = __vhaddu4(pred.x, t.x);
    pred.y = __vhaddu4(pred.y, t.y);
  }
  return true;
}

__device__ __forceinline__ void hadamard8(int (&t)[8]) {
  const int a0 = t[0] + t[4], a4 = t[0] - t[4], a1 = t[1] + t[5], a5 = t[1] - t[5];
  const int a2 = t[2] + t[6], a6 = t[2] - t[6], a3 = t[3] + t[7], a7 = t[3] - t[7];
  const int b0 = a0 + a2, b2 = a0 - a2, b1 = a1 + a3, b3 = a1 - a3;
  const int b4 = a4 + a6, b6 = a4 - a6, b5 = a5 + a7, b7 = a5 - a7;
  t[0] = b0 + b1; t[1] = b0 - b1; t[2] = b2 + b3; t[3] = b2 - b3;
  t[4] = b4 + b5; t[5] = b4 - b5; t[6] = b6 + b7; t[7] = b6 - b7;
}

/* All shuffles use the full-warp mask (xor 4/2/1 stay inside an 8-lane slice):
   every lane of the warp takes part, tail slices work on a clamped index and
   only skip the final store. */
__device__ __forceinline__ int group_sum8(int v) {
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
  return v;
}

/* ---- block metrics: ONE LANE PER 8x8 BLOCK --------------------------------
   A lane loads its block's descriptor, then all 8 source rows and all 8
   (motion-displaced, possibly two-tap) predictor rows -- up to 40 independent
   64-bit loads in flight -- and reduces in registers: no shuffles, no shared
   memory.  Lists in raster order make the row accesses of a warp contiguous.
   (The first version used 8 lanes per block with shuffle butterflies: 2-3x
   more instructions per block.) */
struct Rows8 { uint2 r[8]; };

__device__ __forceinline__ void load_rows8(const uint8_t *p, int ystride, Rows8 &o) {
  const uintptr_t a = (uintptr_t)p;
  const unsigned sh = (unsigned)(a & 7);
  const uint8_t *base = (const uint8_t *)(a & ~(uintptr_t)7);
  if (sh == 0) {
#pragma unroll
    for (int i = 0; i < 8; i++) o.r[i] = __ldg((const uint2 *)(base + i * ystride));
    return;
  }
  const unsigned sel = 0x3210u + 0x1111u * (sh & 3);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint2 w0 = __ldg((const uint2 *)(base + i * ystride));
    const uint2 w1 = __ldg((const uint2 *)(base + i * ystride) + 1);
    if (sh < 4) {
      o.r[i].x = __byte_perm(w0.x, w0.y, sel);
      o.r[i].y = __byte_perm(w0.y, w1.x, sel);
    } else {
      o.r[i].x = __byte_perm(w0.y, w1.x, sel);
      o.r[i].y = __byte_perm(w1.x, w1.y, sel);
    }
  }
}

/* predictor rows: zeros (intra), one tap, or (a+b)>>1 of two taps (encfrag.c:71-84) */
__device__ __forceinline__ void load_pred8(const uint8_t *ref_base, const ocg_enc_frag &f, int ystride, Rows8 &p) {
  if (f.ref_off0 == INT_MIN) {
#pragma unroll
    for (int i = 0; i < 8; i++) p.r[i] = make_uint2(0, 0);
    return;
  }
  load_rows8(ref_base + f.ref_off0, ystride, p);
  if (f.ref_off1 != INT_MIN) {
    Rows8 t;
    load_rows8(ref_base + f.ref_off1, ystride, t);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      p.r[i].x = __vhaddu4(p.r[i].x, t.r[i].x);
      p.r[i].y = __vhaddu4(p.r[i].y, t.r[i].y);
    }
  }
}

/* 2-D 8x8 Hadamard of (s - p), sum of magnitudes without the DC term, DC returned separately
   (encfrag.c:109-336), all in registers, TWO VALUES PER REGISTER: the residual is held as halfword pairs
   (columns 2j, 2j+1) with a bias that keeps every halfword non-negative, so that a plain 32-bit IADD3
   adds or subtracts both halves at once without borrows between them (a butterfly output is
   x + y or x - y + 2B; the bias B doubles per stage: 256 for the residual, 8192 after the three vertical
   and two of the horizontal stages, where |value| <= 32*255).  The last stage pairs the two halves of one
   register: IDP.2A with weights (1,1) and (1,-1) yields both outputs as 32-bit integers, and VABSDIFF
   accumulates their magnitudes.  ~390 instructions per block instead of ~700 for one value per register. */
template <bool HAVE_PRED>
__device__ __forceinline__ uint32_t satd8x8(const Rows8 &s, const Rows8 &p, int &dc) {
  uint32_t r[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t sw[2] = {s.r[i].x, s.r[i].y};
    const uint32_t pw[2] = {p.r[i].x, p.r[i].y};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned sel = (j & 1) ? 0x4342u : 0x4140u;
      const uint32_t a = __byte_perm(sw[j >> 1], 0, sel);
      if (HAVE_PRED) r[i][j] = a - __byte_perm(pw[j >> 1], 0, sel) + 0x01000100u;
      else r[i][j] = a + 0x01000100u;
    }
  }
  /* vertical stages: rows 4, 2, 1 apart; B = 256 -> 2048 */
#pragma unroll
  for (int st = 0; st < 3; st++) {
    const int d = 4 >> st;
    const uint32_t k2b = 0x02000200u << st;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i & d) continue;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t x = r[i][j], y = r[i + d][j];
        r[i][j] = x + y;
        r[i + d][j] = x - y + k2b;
      }
    }
  }
  /* the DC term is the plain sum of the residual: row 0 now holds the column sums (B = 2048) */
  {
    int t = -8 * 2048;
#pragma unroll
    for (int j = 0; j < 4; j++) t = __dp2a_lo((int)r[0][j], 0x0101, t);
    dc = t;
  }
  /* horizontal stages between registers: columns 4 apart (j, j+2), then 2 apart (j, j+1); B -> 8192 */
  int acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t a0 = r[i][0] + r[i][2], a2 = r[i][0] - r[i][2] + 0x10001000u;
    const uint32_t a1 = r[i][1] + r[i][3], a3 = r[i][1] - r[i][3] + 0x10001000u;
    const uint32_t b[4] = {a0 + a1, a0 - a1 + 0x20002000u, a2 + a3, a2 - a3 + 0x20002000u};
    /* last stage inside each register: lo + hi and lo - hi as integers, magnitudes accumulated */
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int u = __dp2a_lo((int)b[j], 0x0101, -2 * 8192);
      const int v = __dp2a_lo((int)b[j], 0xFF01, 0);
      acc = (int)__sad(u, 0, (unsigned)acc);
      acc = (int)__sad(v, 0, (unsigned)acc);
    }
  }
  return (uint32_t)(acc - abs(dc));
}

/* four mask bits -> four byte masks */
__device__ __forceinline__ uint32_t nibble_to_bytes(uint32_t n) { return ((n * 0x00204081u) & 0x01010101u) * 0xFFu; }

template <int METRIC>
__global__ void __launch_bounds__(128)
ocg_enc_metrics_kernel(const uint8_t *__restrict__ src_base, const uint8_t *__restrict__ ref_base, int ystride,
                       const ocg_enc_frag *__restrict__ frags, int n, uint32_t *__restrict__ out_val,
                       int32_t *__restrict__ out_dc) {
  const int fi = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (fi >= n) return;
  const int4 fw = __ldg((const int4 *)(frags + fi));
  ocg_enc_frag f;
  f.src_off = fw.x; f.ref_off0 = fw.y; f.ref_off1 = fw.z; f.aux = fw.w;
  Rows8 s, p;
  load_rows8(src_base + f.src_off, ystride, s);
  uint32_t val = 0;
  int dc = 0;
  if (METRIC == OCG_MET_SAD) {
    load_pred8(ref_base, f, ystride, p);
#pragma unroll
    for (int i = 0; i < 8; i++) val += __vsadu4(s.r[i].x, p.r[i].x) + __vsadu4(s.r[i].y, p.r[i].y);
  } else if (METRIC == OCG_MET_SAD_THRESH) {
    /* encfrag.c:55-84: rows are added in order and the sum is returned as soon as it exceeds the threshold */
    const uint32_t thresh = (uint32_t)f.aux;
    load_pred8(ref_base, f, ystride, p);
    bool out = false;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t next = val + __vsadu4(s.r[i].x, p.r[i].x) + __vsadu4(s.r[i].y, p.r[i].y);
      if (!out) val = next;
      out = out || val > thresh;
    }
  } else if (METRIC == OCG_MET_SSD) {
    /* sum (a-b)^2 = sum a^2 + sum b^2 - 2 sum ab, four pixels per IDP.4A;
       oc_enc_frag_ssd has no two-tap form (encfrag.c:338) */
    f.ref_off1 = INT_MIN;
    load_pred8(ref_base, f, ystride, p);
    unsigned aa = 0, bb = 0, ab = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      aa = __dp4a(s.r[i].x, s.r[i].x, aa); aa = __dp4a(s.r[i].y, s.r[i].y, aa);
      bb = __dp4a(p.r[i].x, p.r[i].x, bb); bb = __dp4a(p.r[i].y, p.r[i].y, bb);
      ab = __dp4a(s.r[i].x, p.r[i].x, ab); ab = __dp4a(s.r[i].y, p.r[i].y, ab);
    }
    val = aa + bb - 2u * ab;
  } else if (METRIC == OCG_MET_BORDER_SSD) {
    /* encfrag.c:352-366: only the pixels whose mask bit is set count; masked-out
       bytes are zeroed in both operands, so they add 0 to every dot product */
    const uint32_t mlo = (uint32_t)f.ref_off1, mhi = (uint32_t)f.aux;
    f.ref_off1 = INT_MIN;
    load_pred8(ref_base, f, ystride, p);
    unsigned aa = 0, bb = 0, ab = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t m8 = ((i < 4 ? mlo : mhi) >> (8 * (i & 3))) & 0xFFu;
      const uint32_t m0 = nibble_to_bytes(m8 & 15u), m1 = nibble_to_bytes(m8 >> 4);
      const uint32_t sx = s.r[i].x & m0, sy = s.r[i].y & m1, px = p.r[i].x & m0, py = p.r[i].y & m1;
      aa = __dp4a(sx, sx, aa); aa = __dp4a(sy, sy, aa);
      bb = __dp4a(px, px, bb); bb = __dp4a(py, py, bb);
      ab = __dp4a(sx, px, ab); ab = __dp4a(sy, py, ab);
    }
    val = aa + bb - 2u * ab;
  } else if (METRIC == OCG_MET_INTRA_SAD) {
    /* encfrag.c:88-107: dc=(sum+32)>>6, then sum |src-dc| */
    unsigned tot = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += __vsadu4(s.r[i].x, 0) + __vsadu4(s.r[i].y, 0);
    const uint32_t m = 0x01010101u * ((tot + 32) >> 6);
#pragma unroll
    for (int i = 0; i < 8; i++) val += __vsadu4(s.r[i].x, m) + __vsadu4(s.r[i].y, m);
  } else {
    /* SATD family, encfrag.c:109-336: 2-D Hadamard of the residual, sum of
       magnitudes without the DC term, DC returned separately. */
    if (METRIC == OCG_MET_INTRA_SATD) {
#pragma unroll
      for (int i = 0; i < 8; i++) p.r[i] = make_uint2(0, 0);
      val = satd8x8<false>(s, p, dc);
    } else {
      load_pred8(ref_base, f, ystride, p);
      val = satd8x8<true>(s, p, dc);
    }
  }
  out_val[fi] = val;
  if (out_dc != nullptr) out_dc[fi] = dc;
}


/* oc_mv (state.h:225-240): x in the low byte, y above it, half-pel units */
__device__ __forceinline__ int mv_x(int mv) { return (int)(signed char)(mv & 0xFF); }
__device__ __forceinline__ int mv_y(int mv) { return (int)(short)mv >> 8; }
__device__ __forceinline__ int mv_make(int x, int y) { return (int)(short)((x & 0xFF) | (y * 256)); }

/* ---- inter-frame analysis tables (BASELINE configs[3]) ---------------------
   oc_cost_inter* (analyze.c:2062-2286) score a macro block's 4 luma + 2..8 chroma blocks against a
   predictor that is a function of (reference frame, vector).  Once the motion analysis of the frame is
   known, the vectors of every mode but LAST/LAST2 are known too, so the predictors can be listed per
   fragment and scored in one batch.  Candidate k of a macro block:
     0 PREV (0,0)   OC_MODE_INTER_NOMV        4 GOLD unrefined      OC_MODE_GOLDEN_MV, first cost
     1 GOLD (0,0)   OC_MODE_GOLDEN_NOMV       5 GOLD refined        ... after oc_mcenc_refine1mv (speculative)
     2 PREV unrefined  OC_MODE_INTER_MV, first cost (analyze.c:2437)
     3 PREV refined    ... after oc_mcenc_refine1mv (2490)
     6 4MV unrefined block vectors, chroma vectors from all four (oc_set_chroma_mvs*, state.c:33-97)
     7 4MV refined block vectors
   The entries are keyed by what the hook will be called with: predictor tap offsets relative to the frame
   pool, so a look-up is exact whoever asks (LAST/LAST2 hit whenever their vector coincides with one above). */
__device__ __forceinline__ void enc_mv_taps(int mv, int qx, int qy, int ystride, int &off0, int &fx, int &fy) {
  /* state.c:846-957 in closed form: first tap truncates towards zero, the second one (present iff a
     component has a fractional part) lies one step further from zero */
  const int dx = mv_x(mv), dy = mv_y(mv);
  const int ax = abs(dx), ay = abs(dy);
  const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
  fx = (ax & (qx ? 3 : 1)) ? sx : 0;
  fy = (ay & (qy ? 3 : 1)) ? sy : 0;
  off0 = sy * (ay >> (1 + qy)) * ystride + sx * (ax >> (1 + qx));
}
__device__ __forceinline__ int div_round_pow2(int v, int shift, int rval) { return (v + (v < 0 ? -1 : 0) + rval) >> shift; }

struct OcgCandJob {
  const ocg_me_mb *mb;
  const int32_t *mbfrags;   /* [nmbs][12]: state.mb_maps[mbi][pli][bi], -1 = absent */
  const int32_t *frag_off;  /* [nfrags] frag_buf_offs */
  ocg_enc_frag *out;        /* luma block [K][nluma], then chroma block [K][nchroma] */
  int32_t nmbs, nfrags, nluma;
  int32_t ystride_y, ystride_c, qx, qy, fmt;
  int32_t io_off, prev_off, gold_off; /* buffer index * ref_frame_sz */
};

__global__ void __launch_bounds__(128)
ocg_enc_cand_kernel(const OcgCandJob J) {
  const int mbi = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (mbi >= J.nmbs) return;
  const int32_t *mf = J.mbfrags + (size_t)mbi * 12;
  if (mf[0] < 0) return; /* macro block outside the coded frame */
  const ocg_me_mb m = J.mb[mbi];
  const int nchroma = J.nfrags - J.nluma;
  for (int k = 0; k < OCG_ENC_NCAND; k++) {
    const bool gold = k == 1 || k == 4 || k == 5;
    const int frame_off = gold ? J.gold_off : J.prev_off;
    int lb[4], cb[4], one = 0;
    bool four = false;
    switch (k) {
      case 2: one = m.unref_mv[1]; break;
      case 3: one = m.analysis_mv[0][1]; break;
      case 4: one = m.unref_mv[0]; break;
      case 5: one = m.gold_ref_mv; break;
      case 6: four = true; for (int b = 0; b < 4; b++) lb[b] = m.block_mv[b]; break;
      case 7: four = true; for (int b = 0; b < 4; b++) lb[b] = m.ref_mv[b]; break;
      default: break;
    }
    if (four) {
      /* state.c:33-87 by pixel format: 0 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 (1 is reserved: chroma decimated in y only) */
      for (int b = 0; b < 4; b++) cb[b] = lb[b];
      if (J.fmt == 0) {
        const int dx = mv_x(lb[0]) + mv_x(lb[1]) + mv_x(lb[2]) + mv_x(lb[3]);
        const int dy = mv_y(lb[0]) + mv_y(lb[1]) + mv_y(lb[2]) + mv_y(lb[3]);
        cb[0] = mv_make(div_round_pow2(dx, 2, 2), div_round_pow2(dy, 2, 2));
      } else if (J.fmt == 1) {
        cb[0] = mv_make(div_round_pow2(mv_x(lb[0]) + mv_x(lb[2]), 1, 1), div_round_pow2(mv_y(lb[0]) + mv_y(lb[2]), 1, 1));
        cb[1] = mv_make(div_round_pow2(mv_x(lb[1]) + mv_x(lb[3]), 1, 1), div_round_pow2(mv_y(lb[1]) + mv_y(lb[3]), 1, 1));
      } else if (J.fmt == 2) {
        cb[0] = mv_make(div_round_pow2(mv_x(lb[0]) + mv_x(lb[1]), 1, 1), div_round_pow2(mv_y(lb[0]) + mv_y(lb[1]), 1, 1));
        cb[2] = mv_make(div_round_pow2(mv_x(lb[2]) + mv_x(lb[3]), 1, 1), div_round_pow2(mv_y(lb[2]) + mv_y(lb[3]), 1, 1));
      }
    }
    for (int idx = 0; idx < 12; idx++) {
      const int fragi = mf[idx];
      if (fragi < 0) continue;
      const int pli = idx >> 2, bi = idx & 3;
      const int mv = four ? (pli ? cb[bi] : lb[bi]) : one;
      const int ystride = pli ? J.ystride_c : J.ystride_y;
      int off0, fx, fy;
      enc_mv_taps(mv, pli ? J.qx : 0, pli ? J.qy : 0, ystride, off0, fx, fy);
      const int foff = J.frag_off[fragi];
      ocg_enc_frag e;
      e.src_off = J.io_off + foff;
      e.ref_off0 = frame_off + foff + off0;
      e.ref_off1 = (fx | fy) ? e.ref_off0 + fy * ystride + fx : INT_MIN;
      e.aux = 0;
      const size_t at = fragi < J.nluma ? (size_t)k * J.nluma + fragi
                                        : (size_t)OCG_ENC_NCAND * J.nluma + (size_t)k * nchroma + (fragi - J.nluma);
      J.out[at] = e;
    }
  }
}

/* ---- speculative frag_sub + fDCT + quantiser for the likeliest predictors ------------------------------
   oc_enc_block_transform_quantize (analyze.c:667-882) transforms a block against the predictor of the mode
   the serial decision picked; for the two predictors most macro blocks end up with -- the co-located block
   of PREV (OC_MODE_INTER_NOMV) and PREV displaced by the refined vector (OC_MODE_INTER_MV, and LAST/LAST2
   whenever they repeat it) -- the whole chain is computed ahead for every fragment and every quantiser of
   the frame, and handed over in compact form: only the leading `count` zig-zag coefficients, beyond which
   every quantiser's output is zero. */
#define OCG_FQ_NSEL 3
__constant__ int c_fq_sel[OCG_FQ_NSEL] = {0, 3, 2}; /* measured on the benchmark encode: 80 % of the blocks the first two
                                                       missed were coded against PREV displaced by the UNREFINED vector */

__global__ void __launch_bounds__(256)
ocg_enc_fq_list_kernel(const ocg_enc_frag *__restrict__ cand, ocg_enc_frag *__restrict__ fq, int nfrags, int nluma, int nqis) {
  const int f = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (f >= nfrags) return;
  const int nchroma = nfrags - nluma;
  const bool luma = f < nluma;
  /* plane of a chroma fragment: first half Cb, second half Cr (same quantiser class layout [pli]) */
  const int pli = luma ? 0 : (f - nluma < nchroma / 2 ? 1 : 2);
  for (int sel = 0; sel < OCG_FQ_NSEL; sel++) {
    const int k = c_fq_sel[sel];
    const size_t cat = luma ? (size_t)k * nluma + f : (size_t)OCG_ENC_NCAND * nluma + (size_t)k * nchroma + (f - nluma);
    ocg_enc_frag e = cand[cat];
    for (int q = 0; q < nqis; q++) {
      e.aux = pli | 1 << 2 | q << 3; /* inter tables */
      const int slot = q * OCG_FQ_NSEL + sel;
      const size_t at = luma ? (size_t)slot * nluma + f : (size_t)3 * OCG_FQ_NSEL * nluma + (size_t)slot * nchroma + (f - nluma);
      fq[at] = e;
    }
  }
}

struct ocg_fq_desc_dev { uint32_t off; uint8_t count, nz[3]; };

/* the candidates of a fragment side by side for the host: [k][fragment] arrays -> [fragment][k] records */
__global__ void __launch_bounds__(256)
ocg_enc_cand_pack_kernel(const ocg_enc_frag *__restrict__ cand, const uint32_t *__restrict__ satd, const int32_t *__restrict__ dc,
                         int nfrags, int nluma, ocg_enc_cand_rec *__restrict__ out) {
  const int f = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (f >= nfrags) return;
  const int nchroma = nfrags - nluma;
  const bool luma = f < nluma;
#pragma unroll
  for (int k = 0; k < OCG_ENC_NCAND; k++) {
    const size_t at = luma ? (size_t)k * nluma + f : (size_t)OCG_ENC_NCAND * nluma + (size_t)k * nchroma + (f - nluma);
    const ocg_enc_frag e = cand[at];
    ocg_enc_cand_rec r;
    r.ref_off0 = e.ref_off0; r.ref_off1 = e.ref_off1; r.satd = satd[at]; r.dc = dc[at];
    out[(size_t)f * OCG_ENC_NCAND + k] = r;
  }
}

__global__ void __launch_bounds__(256)
ocg_enc_fq_compact_kernel(const ocg_enc_frag *__restrict__ cand, const int16_t *__restrict__ dct, const int16_t *__restrict__ qdct,
                          const int32_t *__restrict__ nonzero, int nfrags, int nluma, int nqis, ocg_fq_desc_dev *__restrict__ desc,
                          uint4 *__restrict__ pool, uint32_t pool_units, uint32_t *__restrict__ counter) {
  const int f = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (f >= nfrags) return;
  const int nchroma = nfrags - nluma;
  const bool luma = f < nluma;
  ocg_enc_frag tap[OCG_FQ_NSEL];
  ocg_fq_desc_dev done[OCG_FQ_NSEL];
  for (int sel = 0; sel < OCG_FQ_NSEL; sel++) {
    const int k = c_fq_sel[sel];
    tap[sel] = cand[luma ? (size_t)k * nluma + f : (size_t)OCG_ENC_NCAND * nluma + (size_t)k * nchroma + (f - nluma)];
    /* a predictor that repeats an earlier one (a refinement that moved nothing) shares its entry */
    int same = -1;
    for (int e = 0; e < sel; e++)
      if (tap[e].ref_off0 == tap[sel].ref_off0 && tap[e].ref_off1 == tap[sel].ref_off1) same = e;
    if (same >= 0) {
      done[sel] = done[same];
      desc[(size_t)sel * nfrags + f] = done[sel];
      continue;
    }
    size_t at[3];
    int nz[3] = {0, 0, 0}, cnt = 0;
    for (int q = 0; q < nqis; q++) {
      const int slot = q * OCG_FQ_NSEL + sel;
      at[q] = luma ? (size_t)slot * nluma + f : (size_t)3 * OCG_FQ_NSEL * nluma + (size_t)slot * nchroma + (f - nluma);
      nz[q] = nonzero[at[q]];
      cnt = max(cnt, nz[q] + 1);
    }
    cnt = min(cnt, 64);
    const uint32_t rows = (uint32_t)(cnt + 7) >> 3;          /* 16-byte units per array */
    const uint32_t need = rows * (uint32_t)(1 + nqis);
    const uint32_t off = atomicAdd(counter, need);
    ocg_fq_desc_dev d;
    d.off = off + need <= pool_units ? off : 0xFFFFFFFFu;     /* pool exhausted: the host computes this one itself */
    d.count = (uint8_t)cnt;
    d.nz[0] = (uint8_t)nz[0]; d.nz[1] = (uint8_t)nz[1]; d.nz[2] = (uint8_t)nz[2];
    done[sel] = d;
    desc[(size_t)sel * nfrags + f] = d;
    if (d.off == 0xFFFFFFFFu) continue;
    const uint4 *sd = (const uint4 *)(dct + at[0] * 64);
    for (uint32_t r = 0; r < rows; r++) pool[off + r] = sd[r];
    for (int q = 0; q < nqis; q++) {
      const uint4 *sq = (const uint4 *)(qdct + at[q] * 64);
      for (uint32_t r = 0; r < rows; r++) pool[off + (uint32_t)(1 + q) * rows + r] = sq[r];
    }
  }
}

/* ------------------------------------------------------------------------ */
/* oc_mb_activity (analyze.c:1152-1237) per luma block, one lane per block:
   pixel sum and sum of squares -> variance-like activity, and for non-flat
   blocks the four directional edge energies over the 10x10 neighbourhood; an
   "edge" block gets act_th*(act/act_th)^0.7 through the reference's Q10
   log/exp polynomials (mathops.c:294-313). */
__device__ __forceinline__ uint32_t bexp32_q10(int z) {
  const int ipart = z >> 10;
  unsigned n = (unsigned)(z & 1023) << 4;
  n = (n * ((n * ((n * ((n * 3548u >> 15) + 6817u) >> 15) + 15823u) >> 15) + 22708u) >> 15) + 16384u;
  return 14 - ipart > 0 ? (n + (1u << (13 - ipart))) >> (14 - ipart) : n << (ipart - 14);
}
__device__ __forceinline__ int blog32_q10(uint32_t w) {
  if (w == 0) return -1;
  const int ipart = 32 - __clz(w);
  const int n = (int)(ipart - 16 > 0 ? w >> (ipart - 16) : w << (16 - ipart)) - 32768 - 16384;
  const int fpart = (n * ((n * ((n * ((n * -1402 >> 15) + 2546) >> 15) - 5216) >> 15) + 15745) >> 15) - 6793;
  return (ipart << 10) + (fpart >> 4);
}

/* ten pixels of a row starting one byte left of p */
__device__ __forceinline__ void load_row10(const uint8_t *p, int (&o)[10]) {
  const uint2 a = ld8u(p - 1), b = ld8u(p + 7);
  int t[8];
  unpack8(a, t);
#pragma unroll
  for (int i = 0; i < 8; i++) o[i] = t[i];
  o[8] = (int)(b.x & 0xFFu);
  o[9] = (int)((b.x >> 8) & 0xFFu);
}

__global__ void __launch_bounds__(128)
ocg_enc_activity_kernel(const uint8_t *__restrict__ src_base, int ystride, const ocg_enc_frag *__restrict__ frags, int n,
                        uint32_t *__restrict__ out_act, int32_t *__restrict__ out_sum) {
  const int fi = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (fi >= n) return;
  const uint8_t *s = src_base + frags[fi].src_off;
  int up[10], cur[10], dn[10];
  load_row10(s - ystride, up);
  load_row10(s, cur);
  unsigned x = 0, x2 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    load_row10(s + (i + 1) * ystride, dn);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const unsigned c = (unsigned)cur[j + 1];
      x += c;
      x2 += c * c;
      e1 += (unsigned)abs(((cur[j + 2] - cur[j]) << 1) + up[j + 2] - up[j] + dn[j + 2] - dn[j]);
      e2 += (unsigned)abs(((dn[j + 1] - up[j + 1]) << 1) + dn[j] - up[j] + dn[j + 2] - up[j + 2]);
      e3 += (unsigned)abs(((dn[j + 2] - up[j]) << 1) + dn[j + 1] - cur[j] + cur[j + 2] - up[j + 1]);
      e4 += (unsigned)abs(((dn[j] - up[j + 2]) << 1) + dn[j + 1] - cur[j + 2] + cur[j] - up[j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 10; j++) { up[j] = cur[j]; cur[j] = dn[j]; }
  }
  unsigned act = (x2 << 6) - x * x;
  if (act < (8u << 12)) act = min(act, 5u << 12);
  else if (5u * max(max(e1, e2), max(e3, e4)) > 2u * (e1 + e2 + e3 + e4))
    act = bexp32_q10(0x394A + (7 * (blog32_q10(act) - 0x394A + 5) / 10));
  out_act[fi] = act;
  if (out_sum != nullptr) out_sum[fi] = (int32_t)x;
}

/* fdct.c:28-120 */
__device__ __forceinline__ int fd_exp(int t, int bias) { return ((27146 * t + bias) >> 16) + t + (t != 0); }

__device__ __forceinline__ void fdct8(const int (&x)[8], int (&y)[8]) {
  const int a0 = x[0] + x[7], a7 = x[0] - x[7], a1 = x[1] + x[6], a6 = x[1] - x[6];
  const int a2 = x[2] + x[5], a5 = x[2] - x[5], a3 = x[3] + x[4], a4 = x[3] - x[4];
  const int b0 = a0 + a3, b3 = a0 - a3, b1 = a1 + a2, b2 = a1 - a2;
  const int b6 = a6 + a5, b5 = a6 - a5;
  int s = fd_exp(b5, 0xB500) >> 1;
  const int c4 = a4 + s, c5 = a4 - s;
  s = fd_exp(b6, 0xB500) >> 1;
  const int c7 = a7 + s, c6 = a7 - s;
  const int r = fd_exp(b0, 0x4000);
  s = fd_exp(b1, 0xB500);
  int u = (r + s) >> 1;
  y[0] = u;
  y[4] = r - u;
  u = ((K6 * b2 + K2 * b3 + 0x6CB7) >> 16) + (b3 != 0);
  s = ((K6 * u) >> 16) - b2;
  y[2] = u;
  y[6] = ((s * 21600 + 0x2800) >> 18) + s + (s != 0);
  u = ((K5 * c6 + K3 * c5 + 0x0E3D) >> 16) + (c5 != 0);
  s = c6 - ((K5 * u) >> 16);
  y[5] = u;
  y[3] = ((s * 26568 + 0x3400) >> 17) + s + (s != 0);
  u = ((K7 * c4 + K1 * c7 + 0x7B1B) >> 16) + (c7 != 0);
  s = ((K7 * u) >> 16) - c4;
  y[1] = u;
  y[7] = ((s * 20539 + 0x3000) >> 20) + s + (s != 0);
}

/* 8x8 transpose of 16-bit values over an 8-lane slice; p[k] = (v[2k], v[2k+1]). */
__device__ __forceinline__ void xpose8(uint32_t (&p)[4], int g) {
  {
    const bool up = (g & 4) != 0;
    const uint32_t s0 = up ? p[0] : p[2], s1 = up ? p[1] : p[3];
    const uint32_t r0 = __shfl_xor_sync(0xFFFFFFFFu, s0, 4), r1 = __shfl_xor_sync(0xFFFFFFFFu, s1, 4);
    if (up) { p[0] = r0; p[1] = r1; } else { p[2] = r0; p[3] = r1; }
  }
  {
    const bool up = (g & 2) != 0;
    const uint32_t s0 = up ? p[0] : p[1], s1 = up ? p[2] : p[3];
    const uint32_t r0 = __shfl_xor_sync(0xFFFFFFFFu, s0, 2), r1 = __shfl_xor_sync(0xFFFFFFFFu, s1, 2);
    if (up) { p[0] = r0; p[2] = r1; } else { p[1] = r0; p[3] = r1; }
  }
  {
    const bool up = (g & 1) != 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t r = __shfl_xor_sync(0xFFFFFFFFu, p[k], 1);
      p[k] = up ? __byte_perm(p[k], r, 0x3276) : __byte_perm(p[k], r, 0x5410);
    }
  }
}

__device__ __forceinline__ void pack8(const int (&v)[8], uint32_t (&p)[4]) {
#pragma unroll
  for (int k = 0; k < 4; k++) p[k] = __byte_perm((uint32_t)v[2 * k], (uint32_t)v[2 * k + 1], 0x5410);
}
__device__ __forceinline__ void unpack16(const uint32_t (&p)[4], int (&v)[8]) {
#pragma unroll
  for (int k = 0; k < 4; k++) { v[2 * k] = (int)(short)(p[k] & 0xFFFF); v[2 * k + 1] = (int)p[k] >> 16; }
}

/* natural index -> zig-zag position (internal.c:46-55) */
__constant__ uint8_t c_izig[64] = {
    0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
    41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
    46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

/* (A one-lane-per-block variant -- whole 8x8 block in registers, static zig-zag,
   tables in shared memory, coalesced stores through a padded tile -- was
   measured on B200 at 4.9 G blocks/s vs 5.8 G for this 8-lanes-per-block form:
   ~4300 instructions per block at 128 registers per lane and 25 % occupancy.
   The transform + quantiser are instruction-bound either way: 16 one-dimensional
   transforms and 64 quantiser evaluations per block.) */
__global__ void __launch_bounds__(256)
ocg_enc_fdct_quant_kernel(const uint8_t *__restrict__ src_base, const uint8_t *__restrict__ ref_base, int ystride,
                          const ocg_enc_frag *__restrict__ frags, int n, const uint16_t *__restrict__ dequant,
                          const int16_t *__restrict__ enquant, int16_t *__restrict__ dct,
                          int16_t *__restrict__ qdct, int32_t *__restrict__ nonzero) {
  __shared__ __align__(16) int16_t zz[32][64]; /* per fragment slot, zig-zag order */
  const int slot = threadIdx.x >> 3;
  const int fi_raw = (int)(blockIdx.x * 32 + slot);
  const bool live = fi_raw < n;
  const int fi = live ? fi_raw : n - 1;
  const int lane = threadIdx.x & 31;
  const int row = lane & 7;
  const int4 fw = __ldg((const int4 *)(frags + fi));
  ocg_enc_frag f;
  f.src_off = fw.x; f.ref_off0 = fw.y; f.ref_off1 = fw.z; f.aux = fw.w;
  const uint2 s = ld8u(src_base + f.src_off + row * ystride);
  uint2 p;
  int v[8], b[8], y[8];
  unpack8(s, v);
  if (load_pred_row(ref_base, f, row, ystride, p)) {
    unpack8(p, b);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] -= b[i];
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] -= 128;
  }
  /* fdct.c:135-142: two extra bits of precision and the round-trip biases */
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = (int)(short)(v[i] << 2);
  if (row == 0) { v[0] = (int)(short)(v[0] + (v[0] != 0) + 1); v[1] = (int)(short)(v[1] + 1); }
  if (row == 1) v[0] = (int)(short)(v[0] - 1);
  uint32_t pk[4];
  pack8(v, pk);
  xpose8(pk, row); /* lane c now holds column c */
  unpack16(pk, v);
  fdct8(v, y);            /* vertical frequencies of column c */
  pack8(y, pk);
  xpose8(pk, row); /* lane r holds, for vertical frequency r, the 8 columns */
  unpack16(pk, v);
  fdct8(v, y);            /* y[k] = coefficient (row r, column k), natural order */
#pragma unroll
  for (int k = 0; k < 8; k++) zz[slot][c_izig[row * 8 + k]] = (int16_t)(((int)(short)y[k] + 2) >> 2);
  __syncwarp();
  /* lane handles zig-zag positions 8*row .. 8*row+7 */
  const uint4 dv = *(const uint4 *)&zz[slot][row * 8];
  if (live) *(uint4 *)(dct + (size_t)fi * 64 + row * 8) = dv;
  const int pli = f.aux & 3, qti = (f.aux >> 2) & 1, qii = (f.aux >> 3) & 3;
  const int tab = (pli * 2 + qti) * 3 + qii;
  const uint4 dq = __ldg((const uint4 *)(dequant + (size_t)tab * 64 + row * 8));
  const uint4 e0 = __ldg((const uint4 *)(enquant + (size_t)tab * 128 + row * 16));
  const uint4 e1 = __ldg((const uint4 *)(enquant + (size_t)tab * 128 + row * 16 + 8));
  const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
  const uint32_t qw[4] = {dq.x, dq.y, dq.z, dq.w};
  const uint32_t ew[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
  int q[8], last = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    /* enquant.c:232-246 */
    const int c = (k & 1) ? (int)dw[k >> 1] >> 16 : (int)(short)(dw[k >> 1] & 0xFFFF);
    const int d = (int)((qw[k >> 1] >> (16 * (k & 1))) & 0xFFFF);
    const int m = (int)(short)(ew[k] & 0xFFFF), l = (int)ew[k] >> 16;
    int val = c << 1;
    if (abs(val) >= d) {
      const int sg = val < 0 ? -1 : 0;
      val += (d + sg) ^ sg;
      val = ((((m * val) >> 16) + val) >> l) - sg;
      q[k] = (int)(short)val;
      last = row * 8 + k;
    } else q[k] = 0;
  }
  uint32_t qp[4];
  pack8(q, qp);
  if (live) *(uint4 *)(qdct + (size_t)fi * 64 + row * 8) = make_uint4(qp[0], qp[1], qp[2], qp[3]);
  last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, 4));
  last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, 2));
  last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, 1));
  if (row == 0 && live) nonzero[fi] = last;
}

/* ------------------------------------------------------------------------ */
/* oc_mcenc_search_frame's full-pel search (mcenc.c:268-515), one warp per
   macro block.  Lane k handles row (k&7) of luma block (k>>3): a candidate
   vector costs one unaligned 8-byte load, two VABSDIFF4 and a 5-step shuffle
   reduction that yields the four block SADs and their sum.  All control state
   (best vector/error, per-block bests) is warp-uniform; the 31x31 "already
   visited" bitmap (mcenc.c:292) lives one row per lane. */
struct McWarp {
  const uint8_t *ref; /* this lane's row in the searched frame at vector (0,0) */
  uint2 src;          /* this lane's 8 source pixels */
  int ystride;
  uint32_t hit;       /* row (lane) of the visited bitmap */
  int lane;
};

/* returns the 16x16 SAD; berr = SAD of this lane's block (group of 8 lanes) */
__device__ __forceinline__ unsigned mc_sad16(const McWarp &w, int dx, int dy, unsigned &berr) {
  const uint2 r = ld8u(w.ref + dx + dy * w.ystride);
  int v = (int)(__vsadu4(w.src.x, r.x) + __vsadu4(w.src.y, r.y));
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
  berr = (unsigned)v;
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 8);
  v += __shfl_xor_sync(0xFFFFFFFFu, v, 16);
  return (unsigned)v;
}

/* test-and-set in the visited bitmap; warp-uniform result */
__device__ __forceinline__ bool mc_visited(McWarp &w, int dx, int dy) {
  const uint32_t bit = 1u << (dx + 15);
  const uint32_t row = __shfl_sync(0xFFFFFFFFu, w.hit, dy + 15);
  if (w.lane == dy + 15) w.hit |= bit;
  return (row & bit) != 0;
}

/* sum of |8x8 Hadamard| of (src - ref row) over this lane's block: what
   oc_mcenc_ysatd_check_*_fullpel accumulate per block, satd + |dc| (mcenc.c:222-266) */
__device__ __forceinline__ unsigned mc_satd_block(uint2 src, uint2 ref, int row) {
  int a[8], b[8];
  unpack8(src, a);
  unpack8(ref, b);
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] -= b[i];
  hadamard8(a);
#pragma unroll
  for (int d = 4; d >= 1; d >>= 1) {
    const bool up = (row & d) != 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int o = __shfl_xor_sync(0xFFFFFFFFu, a[i], d);
      a[i] = up ? o - a[i] : a[i] + o;
    }
  }
  int acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc += abs(a[i]);
  acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
  acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
  acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
  return (unsigned)acc;
}

__global__ void __launch_bounds__(128)
ocg_mcenc_search_kernel(const uint8_t *__restrict__ src_base, const uint8_t *__restrict__ ref_full,
                        const uint8_t *__restrict__ ref_satd, int ystride, const ocg_mb_search_in *__restrict__ in,
                        ocg_mb_search_out *__restrict__ out, int n) {
  const int mbi = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (mbi >= n) return; /* whole warp */
  const int lane = threadIdx.x & 31;
  const int bi = lane >> 3, row = lane & 7;
  const ocg_mb_search_in *m = in + mbi;
  const int foff = m->frag_off[bi] + row * ystride;
  McWarp w;
  w.ref = ref_full + foff;
  w.src = ld8u(src_base + foff);
  w.ystride = ystride;
  w.hit = 0;
  w.lane = lane;
  /* per-block state lives in the block's 8 lanes (warp-uniform inside a group) */
  unsigned berr, err;
  int cx = (int)m->cand[0][0] / 2, cy = (int)m->cand[0][1] / 2; /* OC_DIV2: towards zero */
  mc_visited(w, cx, cy);
  unsigned best_err = mc_sad16(w, cx, cy, berr);
  int bx = cx, by = cy;
  unsigned blk_err = berr; /* this lane's block */
  int blk_x = cx, blk_y = cy;
  const bool track = m->is_prev != 0;
  if (best_err > 256u) {
    unsigned t2 = m->t2_base;
    t2 += (t2 >> 4) + 64u;
    const int setb0 = m->setb0, ncand = m->ncand;
    int ci = 1;
    for (; ci < setb0; ci++) {
      cx = (int)m->cand[ci][0] / 2;
      cy = (int)m->cand[ci][1] / 2;
      if (mc_visited(w, cx, cy)) continue;
      err = mc_sad16(w, cx, cy, berr);
      if (err < best_err) { best_err = err; bx = cx; by = cy; }
      if (berr < blk_err) { blk_err = berr; blk_x = cx; blk_y = cy; }
    }
    if (best_err > t2) {
      for (; ci < ncand; ci++) {
        cx = (int)m->cand[ci][0] / 2;
        cy = (int)m->cand[ci][1] / 2;
        if (mc_visited(w, cx, cy)) continue;
        err = mc_sad16(w, cx, cy, berr);
        if (err < best_err) { best_err = err; bx = cx; by = cy; }
        if (berr < blk_err) { blk_err = berr; blk_x = cx; blk_y = cy; }
      }
      if (best_err > t2) {
        /* square-pattern descent (mcenc.c:399-431) */
        for (;;) {
          int sx = 0, sy = 0;
          bool moved = false;
          for (int dy = -1; dy <= 1; dy++) {
            for (int dx = -1; dx <= 1; dx++) {
              if ((dx | dy) == 0) continue;
              if ((bx <= -15 && dx < 0) || (bx >= 15 && dx > 0) || (by <= -15 && dy < 0) || (by >= 15 && dy > 0)) continue;
              cx = bx + dx;
              cy = by + dy;
              if (mc_visited(w, cx, cy)) continue;
              err = mc_sad16(w, cx, cy, berr);
              if (err < best_err) { best_err = err; sx = dx; sy = dy; moved = true; }
              if (berr < blk_err) { blk_err = berr; blk_x = cx; blk_y = cy; }
            }
          }
          if (!moved) break;
          bx += sx;
          by += sy;
        }
        /* per-block descents sharing the visited map (mcenc.c:437-499); the
           block being refined is broadcast from its first lane */
        if (track) {
          const unsigned t4 = t2 >> 2;
          for (int b = 0; b < 4; b++) {
            if (__shfl_sync(0xFFFFFFFFu, blk_err, b * 8) <= t4) continue;
            for (;;) {
              const int ox = __shfl_sync(0xFFFFFFFFu, blk_x, b * 8), oy = __shfl_sync(0xFFFFFFFFu, blk_y, b * 8);
              for (int dy = -1; dy <= 1; dy++) {
                for (int dx = -1; dx <= 1; dx++) {
                  if ((dx | dy) == 0) continue;
                  if ((ox <= -15 && dx < 0) || (ox >= 15 && dx > 0) || (oy <= -15 && dy < 0) || (oy >= 15 && dy > 0)) continue;
                  cx = ox + dx;
                  cy = oy + dy;
                  if (mc_visited(w, cx, cy)) continue;
                  err = mc_sad16(w, cx, cy, berr);
                  if (err < best_err) { best_err = err; bx = cx; by = cy; }
                  if (berr < blk_err) { blk_err = berr; blk_x = cx; blk_y = cy; }
                }
              }
              if (__shfl_sync(0xFFFFFFFFu, blk_x, b * 8) == ox && __shfl_sync(0xFFFFFFFFu, blk_y, b * 8) == oy) break;
            }
          }
        }
      }
    }
  }
  /* final SATDs on the reconstructed reference (mcenc.c:500-513) */
  const uint8_t *rs = ref_satd + foff;
  unsigned s_mb = mc_satd_block(w.src, ld8u(rs + bx + by * ystride), row);
  unsigned s_blk = mc_satd_block(w.src, ld8u(rs + blk_x + blk_y * ystride), row);
  unsigned tot = s_mb + __shfl_xor_sync(0xFFFFFFFFu, s_mb, 8);
  tot += __shfl_xor_sync(0xFFFFFFFFu, tot, 16);
  ocg_mb_search_out *o = out + mbi;
  if (lane == 0) {
    o->best_vec[0] = (int8_t)bx;
    o->best_vec[1] = (int8_t)by;
    o->error = (uint16_t)best_err;
    o->satd = tot;
  }
  if (row == 0) {
    o->block_vec[bi][0] = track ? (int8_t)blk_x : (int8_t)0;
    o->block_vec[bi][1] = track ? (int8_t)blk_y : (int8_t)0;
    o->block_satd[bi] = track ? s_blk : 0u;
  }
}

/* ------------------------------------------------------------------------ */
/* Half-pel refinement (mcenc.c:606-791), one warp per macro block, ONE LANE
   PER (block, site): the 4 blocks x 8 half-pel sites of a macro block are 32
   independent two-tap 8x8 scores, each computed like a metrics-kernel block
   (all rows in registers).  1MV: lane = site*4 + block, the four block scores of
   a site are summed with two shuffles; 4MV: lane = block*8 + site.  The winner
   is the first site in OC_SQUARE_SITES[0] order that is strictly better than
   everything before it, i.e. the lexicographic minimum of (score, site order)
   if that beats the entry score. */
__constant__ int c_sq_dx[8] = {-1, 0, 1, -1, 1, -1, 0, 1}; /* OC_SQUARE_SITES[0] = {0,1,2,3,5,6,7,8} */
__constant__ int c_sq_dy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};

__device__ __forceinline__ uint32_t refine_score(const uint8_t *src, const uint8_t *ref, int ystride, int vx, int vy,
                                                 int dx, int dy, bool use_sad) {
  /* mcenc.c:636-646: the two taps of half-pel vector (2v+d) */
  const int xmask = (((vx << 1) + dx) ^ dx) < 0 ? -1 : 0;
  const int ymask = (((vy << 1) + dy) ^ dy) < 0 ? -1 : 0;
  const int oy = dy * ystride;
  const int base = vx + vy * ystride;
  const int o0 = base + (dx & xmask) + (oy & ymask);
  const int o1 = base + (dx & ~xmask) + (oy & ~ymask);
  Rows8 s, p, t;
  load_rows8(src, ystride, s);
  load_rows8(ref + o0, ystride, p);
  load_rows8(ref + o1, ystride, t);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    p.r[i].x = __vhaddu4(p.r[i].x, t.r[i].x);
    p.r[i].y = __vhaddu4(p.r[i].y, t.r[i].y);
  }
  if (use_sad) {
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) v += __vsadu4(s.r[i].x, p.r[i].x) + __vsadu4(s.r[i].y, p.r[i].y);
    return v;
  }
  int dc;
  const uint32_t v = satd8x8<true>(s, p, dc);
  return v + (uint32_t)abs(dc);
}

__global__ void __launch_bounds__(128)
ocg_mcenc_refine_kernel(const uint8_t *__restrict__ src_base, const uint8_t *__restrict__ ref_base, int ystride,
                        const ocg_mb_refine_in *__restrict__ in, ocg_mb_refine_out *__restrict__ out, int n,
                        int flags) {
  const int lane = threadIdx.x & 31;
  const int mbi_raw = (int)(blockIdx.x * 4 + (threadIdx.x >> 5));
  const bool live = mbi_raw < n; /* warp-uniform */
  if (!live) return;
  const ocg_mb_refine_in *mb = in + mbi_raw;
  ocg_mb_refine_out *o = out + mbi_raw;
  if (flags & OCG_REFINE_1MV) {
    const int sitei = lane >> 2, bi = lane & 3;
    const int off = mb->frag_off[bi];
    const int vx = mb->vec[0], vy = mb->vec[1];
    uint32_t err = refine_score(src_base + off, ref_base + off, ystride, vx, vy, c_sq_dx[sitei], c_sq_dy[sitei],
                                (flags & OCG_REFINE_SAD) != 0);
    err += __shfl_xor_sync(0xFFFFFFFFu, err, 1);
    err += __shfl_xor_sync(0xFFFFFFFFu, err, 2);
    /* lexicographic min over the sites of (err, site order) */
    uint32_t key_e = err;
    int key_s = sitei;
#pragma unroll
    for (int d = 4; d < 32; d <<= 1) {
      const uint32_t oe = __shfl_xor_sync(0xFFFFFFFFu, key_e, d);
      const int os = __shfl_xor_sync(0xFFFFFFFFu, key_s, d);
      if (oe < key_e || (oe == key_e && os < key_s)) { key_e = oe; key_s = os; }
    }
    if (lane == 0) {
      const uint32_t entry = mb->satd;
      int dx = 0, dy = 0;
      uint32_t best = entry;
      if (key_e < entry) { best = key_e; dx = c_sq_dx[key_s]; dy = c_sq_dy[key_s]; }
      o->mv[0] = (int8_t)((vx << 1) + dx);
      o->mv[1] = (int8_t)((vy << 1) + dy);
      o->satd = best;
    }
  }
  if (flags & OCG_REFINE_4MV) {
    const int bi = lane >> 3, sitei = lane & 7;
    const int off = mb->frag_off[bi];
    const int vx = mb->block_vec[bi][0], vy = mb->block_vec[bi][1];
    const uint32_t err = refine_score(src_base + off, ref_base + off, ystride, vx, vy, c_sq_dx[sitei], c_sq_dy[sitei],
                                      false);
    uint32_t key_e = err;
    int key_s = sitei;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      const uint32_t oe = __shfl_xor_sync(0xFFFFFFFFu, key_e, d);
      const int os = __shfl_xor_sync(0xFFFFFFFFu, key_s, d);
      if (oe < key_e || (oe == key_e && os < key_s)) { key_e = oe; key_s = os; }
    }
    if (sitei == 0) {
      const uint32_t entry = mb->block_satd[bi];
      int dx = 0, dy = 0;
      uint32_t best = entry;
      if (key_e < entry) { best = key_e; dx = c_sq_dx[key_s]; dy = c_sq_dy[key_s]; }
      o->ref_mv[bi][0] = (int8_t)((vx << 1) + dx);
      o->ref_mv[bi][1] = (int8_t)((vy << 1) + dy);
      o->block_satd[bi] = best;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Whole-frame motion analysis: oc_mcenc_search (mcenc.c:517-548) for every
   macro block in coding order, candidates from already-searched neighbours,
   as a wave-front of warps (one per super-block row and reference frame). */
struct OcgMeJob {
  const uint8_t *src;          /* OC_FRAME_IO: buffer + base_off */
  const uint8_t *ref_full[2];  /* [frame]: the *_ORIG frame searched with SAD */
  const uint8_t *ref_satd[2];  /* [frame]: the reconstructed reference */
  const ocg_me_topo *topo;
  ocg_me_mb *mb;
  uint32_t *done;              /* [nmbs][2]: == seq once the MB's analysis against that frame is final */
  const uint8_t *gold_refine;  /* nmbs flags or NULL */
  uint32_t *ticket;            /* [frame][2]: super-block rows claimed / finished in this launch */
  int32_t ystride, nhsbs, nvsbs, nmbs, flags;
  uint32_t seq;
};

__device__ __forceinline__ int mv_add(int a, int b) { return mv_make(mv_x(a) + mv_x(b), mv_y(a) + mv_y(b)); }
__device__ __forceinline__ int mv_sub(int a, int b) { return mv_make(mv_x(a) - mv_x(b), mv_y(a) - mv_y(b)); }
__device__ __forceinline__ int clamp31(int v) { return max(-31, min(31, v)); }

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

/* oc_mcenc_ysatd_halfpel_mbrefine (mcenc.c:606-664) by the whole warp:
   lane = site*4 + block.  Returns the refined vector (half-pel) and score. */
__device__ __forceinline__ void me_refine1(const uint8_t *src, const uint8_t *ref, int ystride, int off,
                                           int lane, int vx, int vy, uint32_t entry, bool use_sad, int &ox, int &oy,
                                           uint32_t &oscore) {
  const int sitei = lane >> 2;
  uint32_t err = refine_score(src + off, ref + off, ystride, vx, vy, c_sq_dx[sitei], c_sq_dy[sitei], use_sad);
  err += __shfl_xor_sync(0xFFFFFFFFu, err, 1);
  err += __shfl_xor_sync(0xFFFFFFFFu, err, 2);
  uint32_t key_e = err;
  int key_s = sitei;
#pragma unroll
  for (int d = 4; d < 32; d <<= 1) {
    const uint32_t oe = __shfl_xor_sync(0xFFFFFFFFu, key_e, d);
    const int os = __shfl_xor_sync(0xFFFFFFFFu, key_s, d);
    if (oe < key_e || (oe == key_e && os < key_s)) { key_e = oe; key_s = os; }
  }
  int dx = 0, dy = 0;
  oscore = entry;
  if (key_e < entry) { oscore = key_e; dx = c_sq_dx[key_s]; dy = c_sq_dy[key_s]; }
  ox = (vx << 1) + dx;
  oy = (vy << 1) + dy;
}

__global__ void __launch_bounds__(32)
ocg_me_wavefront_kernel(const OcgMeJob *__restrict__ jobs) {
  const OcgMeJob J = jobs[blockIdx.z];
  const int f = (int)blockIdx.y; /* 0 = OC_FRAME_GOLD, 1 = OC_FRAME_PREV */
  const int lane = threadIdx.x;
  const int bi = lane >> 3, row = lane & 7;
  const int ystride = J.ystride;
  const bool is_prev = f == 1;
  const bool nosatd = (J.flags & OCG_ME_NOSATD) != 0;
  const bool want_blocks = is_prev && (J.flags & OCG_ME_FAST) == 0;
  /* Rows are CLAIMED in order through a ticket, not taken from blockIdx: a CTA only ever waits on rows
     with smaller tickets, and those were claimed by CTAs that are already running, so the wave-front makes
     progress whatever order the hardware dispatches the grid in and however many CTAs are resident. */
  int row_claim = 0;
  if (lane == 0) row_claim = (int)atomicAdd(J.ticket + f * 2, 1u);
  row_claim = __shfl_sync(0xFFFFFFFFu, row_claim, 0);
  const int mbi0 = row_claim * J.nhsbs * 4;
  for (int k = 0; k < J.nhsbs * 4; k++) {
    const int mbi = mbi0 + k;
    const ocg_me_topo *T = J.topo + mbi;
    if (!T->valid) continue; /* warp-uniform */
    ocg_me_mb *m = J.mb + mbi;
    const int off_search = T->frag_off[bi];      /* search: lane = block*8 + row  */
    const int off_refine = T->frag_off[lane & 3]; /* refine: lane = site*4 + block */
    const int ncn = T->ncn;
    /* history rotation, mcenc.c:523-531 / 540-541 (this MB's own values from earlier frames) */
    const int mv0 = m->analysis_mv[0][f];
    int mv1 = m->analysis_mv[1][f], mv2 = m->analysis_mv[2][f];
    int accum;
    if (is_prev) {
      accum = (J.flags & OCG_ME_DROPPED) ? mv0 : 0;
      const int old2 = mv2;
      mv2 = mv1;
      mv1 = mv_sub(mv0, old2);
    } else {
      accum = mv2;
      mv2 = mv1;
      mv1 = mv0;
      mv1 = mv_sub(mv1, mv2);
      mv2 = mv_sub(mv2, accum);
    }
    unsigned t2 = m->error[f];
    /* wait for the neighbours' final vectors, then read them around L1 */
    int cmv = 0;
    unsigned cerr = 0;
    if (lane < ncn) {
      const int n = T->cn[lane];
      const uint32_t *flag = J.done + (size_t)n * 2 + f;
      while (ld_acquire_u32(flag) != J.seq) __nanosleep(64);
      cmv = (int)__ldcg(&J.mb[n].analysis_mv[0][f]);
      cerr = (unsigned)__ldcg(&J.mb[n].error[f]);
    }
    __syncwarp();
    /* candidate list, one per lane (mcenc.c:90-164): [0] median, [1..ncn] neighbours,
       accum, clamp(mv1+accum), (0,0) | set B: clamp(2*mv1-mv2+accum) */
    const int ax = mv_x(accum), ay = mv_y(accum);
    int candx = 0, candy = 0;
    if (lane >= 1 && lane <= ncn) {
      const int v = __shfl_sync(0xFFFFFFFFu, cmv, (lane - 1) & 31);
      candx = mv_x(v);
      candy = mv_y(v);
    } else {
      (void)__shfl_sync(0xFFFFFFFFu, cmv, 0);
    }
    if (lane == ncn + 1) { candx = ax; candy = ay; }
    if (lane == ncn + 2) { candx = clamp31(mv_x(mv1) + ax); candy = clamp31(mv_y(mv1) + ay); }
    /* lane ncn+3: (0,0) */
    const int setb0 = ncn + 4;
    if (lane == setb0) {
      candx = clamp31(2 * mv_x(mv1) - mv_x(mv2) + ax);
      candy = clamp31(2 * mv_y(mv1) - mv_y(mv2) + ay);
    }
    const int ncand = setb0 + 1;
    {
      /* median of the first three of set A (mcenc.c:130-138) */
      int a0x = __shfl_sync(0xFFFFFFFFu, candx, 1), a1x = __shfl_sync(0xFFFFFFFFu, candx, 2), a2x = __shfl_sync(0xFFFFFFFFu, candx, 3);
      int a0y = __shfl_sync(0xFFFFFFFFu, candy, 1), a1y = __shfl_sync(0xFFFFFFFFu, candy, 2), a2y = __shfl_sync(0xFFFFFFFFu, candy, 3);
      const int medx = max(min(a0x, a1x), min(max(a0x, a1x), a2x));
      const int medy = max(min(a0y, a1y), min(max(a0y, a1y), a2y));
      if (lane == 0) { candx = medx; candy = medy; }
    }
    /* early-termination threshold base: own previous error and the first <=3 neighbours' (mcenc.c:337-341) */
    {
      const int ncs = min(3, ncn);
      for (int i = 0; i < ncs; i++) t2 = max(t2, __shfl_sync(0xFFFFFFFFu, cerr, i));
    }
    /* ---- full-pel search, mcenc.c:305-499 ---- */
    McWarp w;
    const int po = off_search + row * ystride;
    w.ref = J.ref_full[f] + po;
    w.src = ld8u(J.src + po);
    w.ystride = ystride;
    w.hit = 0;
    w.lane = lane;
    unsigned berr, err;
    int cx = __shfl_sync(0xFFFFFFFFu, candx, 0) / 2, cy = __shfl_sync(0xFFFFFFFFu, candy, 0) / 2;
    mc_visited(w, cx, cy);
    unsigned best_err = mc_sad16(w, cx, cy, berr);
    int bx = cx, by = cy;
    unsigned blk_err = berr;
    int blk_x = cx, blk_y = cy;
    if (best_err > 256u) {
      t2 += (t2 >> 4) + 64u;
      int ci = 1;
      for (; ci < setb0; ci++) {
        cx = __shfl_sync(0xFFFFFFFFu, candx, ci) / 2;
        cy = __shfl_sync(0xFFFFFFFFu, candy, ci) / 2;
        if (mc_visited(w, cx, cy)) continue;
        err = mc_sad16(w, cx, cy, berr);
        if (err < best_err) { best_err = err; bx = cx; by = cy; }
        if (berr < blk_err) { blk_err = berr; blk_x = cx; blk_y = cy; }
      }
      if (best_err > t2) {
        for (; ci < ncand; ci++) {
          cx = __shfl_sync(0xFFFFFFFFu, candx, ci) / 2;
          cy = __shfl_sync(0xFFFFFFFFu, candy, ci) / 2;
          if (mc_visited(w, cx, cy)) continue;
          err = mc_sad16(w, cx, cy, berr);
          if (err < best_err) { best_err = err; bx = cx; by = cy; }
          if (berr < blk_err) { blk_err = berr; blk_x = cx; blk_y = cy; }
        }
        if (best_err > t2) {
          for (;;) {
            int sx = 0, sy = 0;
            bool moved = false;
            for (int dy = -1; dy <= 1; dy++) {
              for (int dx = -1; dx <= 1; dx++) {
                if ((dx | dy) == 0) continue;
                if ((bx <= -15 && dx < 0) || (bx >= 15 && dx > 0) || (by <= -15 && dy < 0) || (by >= 15 && dy > 0)) continue;
                cx = bx + dx;
                cy = by + dy;
                if (mc_visited(w, cx, cy)) continue;
                err = mc_sad16(w, cx, cy, berr);
                if (err < best_err) { best_err = err; sx = dx; sy = dy; moved = true; }
                if (berr < blk_err) { blk_err = berr; blk_x = cx; blk_y = cy; }
              }
            }
            if (!moved) break;
            bx += sx;
            by += sy;
          }
          if (is_prev) {
            const unsigned t4 = t2 >> 2;
            for (int b = 0; b < 4; b++) {
              if (__shfl_sync(0xFFFFFFFFu, blk_err, b * 8) <= t4) continue;
              for (;;) {
                const int ox = __shfl_sync(0xFFFFFFFFu, blk_x, b * 8), oy = __shfl_sync(0xFFFFFFFFu, blk_y, b * 8);
                for (int dy = -1; dy <= 1; dy++) {
                  for (int dx = -1; dx <= 1; dx++) {
                    if ((dx | dy) == 0) continue;
                    if ((ox <= -15 && dx < 0) || (ox >= 15 && dx > 0) || (oy <= -15 && dy < 0) || (oy >= 15 && dy > 0)) continue;
                    cx = ox + dx;
                    cy = oy + dy;
                    if (mc_visited(w, cx, cy)) continue;
                    err = mc_sad16(w, cx, cy, berr);
                    if (err < best_err) { best_err = err; bx = cx; by = cy; }
                    if (berr < blk_err) { blk_err = berr; blk_x = cx; blk_y = cy; }
                  }
                }
                if (__shfl_sync(0xFFFFFFFFu, blk_x, b * 8) == ox && __shfl_sync(0xFFFFFFFFu, blk_y, b * 8) == oy) break;
              }
            }
          }
        }
      }
    }
    /* ---- final score on the reconstructed reference, mcenc.c:500-514 ---- */
    const uint8_t *rs = J.ref_satd[f] + po;
    unsigned s_mb;
    if (nosatd) {
      const uint2 r = ld8u(rs + bx + by * ystride);
      s_mb = __vsadu4(w.src.x, r.x) + __vsadu4(w.src.y, r.y);
      s_mb += __shfl_xor_sync(0xFFFFFFFFu, s_mb, 4);
      s_mb += __shfl_xor_sync(0xFFFFFFFFu, s_mb, 2);
      s_mb += __shfl_xor_sync(0xFFFFFFFFu, s_mb, 1);
    } else {
      s_mb = mc_satd_block(w.src, ld8u(rs + bx + by * ystride), row);
    }
    unsigned tot = s_mb + __shfl_xor_sync(0xFFFFFFFFu, s_mb, 8);
    tot += __shfl_xor_sync(0xFFFFFFFFu, tot, 16);
    unsigned s_blk = 0;
    if (want_blocks) s_blk = mc_satd_block(w.src, ld8u(rs + blk_x + blk_y * ystride), row); /* warp-uniform branch */
    int fmv = mv_make(bx * 2, by * 2);
    uint32_t fsatd = tot;
    /* history as oc_mcenc_search leaves it (mcenc.c:534, 546-547) */
    int h1, h2;
    if (is_prev) { h2 = accum; h1 = mv1; }
    else { h2 = mv_add(mv2, accum); h1 = mv_add(mv1, h2); }
    if (lane == 0) {
      m->analysis_mv[1][f] = (int16_t)h1;
      m->analysis_mv[2][f] = (int16_t)h2;
      m->error[f] = (uint16_t)best_err;
      m->unref_mv[f] = (int16_t)fmv;
      m->unref_satd[f] = tot;
    }
    if (want_blocks && row == 0) {
      m->block_mv[bi] = (int16_t)mv_make(blk_x * 2, blk_y * 2);
      m->block_satd[bi] = s_blk;
    }
    /* ---- oc_mcenc_refine1mv where the analysis loop runs it (analyze.c:2476-2489) ---- */
    bool refine = is_prev ? (J.flags & OCG_ME_REFINE_PREV) != 0
                          : (J.gold_refine != nullptr && J.gold_refine[mbi] != 0);
    if (refine) {
      int rx, ry;
      uint32_t rscore;
      me_refine1(J.src, J.ref_satd[f], ystride, off_refine, lane, bx, by, tot, nosatd, rx, ry, rscore);
      fmv = mv_make(rx, ry);
      fsatd = rscore;
    }
    if (lane == 0) {
      m->analysis_mv[0][f] = (int16_t)fmv;
      m->satd[f] = fsatd;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_u32(J.done + (size_t)mbi * 2 + f, J.seq);
    /* Speculative oc_mcenc_refine1mv(OC_FRAME_GOLD): what the refinement WOULD give, kept beside the
       state (the host decides per macro block whether it happens, analyze.c:2476-2485); off the chain */
    if (!is_prev && !refine && (J.flags & OCG_ME_SPEC_GOLD)) {
      int rx, ry;
      uint32_t rscore;
      me_refine1(J.src, J.ref_satd[f], ystride, off_refine, lane, bx, by, tot, nosatd, rx, ry, rscore);
      if (lane == 0) {
        m->gold_ref_mv = (int16_t)mv_make(rx, ry);
        m->gold_ref_satd = rscore;
      }
    }
  }
  /* the last row to finish re-arms the tickets for the next launch */
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(J.ticket + f * 2 + 1, 1u) + 1u == (unsigned)J.nvsbs) {
      J.ticket[f * 2] = 0;
      J.ticket[f * 2 + 1] = 0;
    }
  }
}

/* oc_mcenc_refine4mv (mcenc.c:763-791) for every macro block of the frame:
   one warp per MB, lane = block*8 + site. */
__global__ void __launch_bounds__(128)
ocg_me_refine4_kernel(const OcgMeJob *__restrict__ jobs) {
  const OcgMeJob J = jobs[blockIdx.y];
  const int mbi = (int)(blockIdx.x * 4 + (threadIdx.x >> 5));
  if (mbi >= J.nmbs) return;
  const ocg_me_topo *T = J.topo + mbi;
  if (!T->valid) return;
  const int lane = threadIdx.x & 31;
  const int bi = lane >> 3, sitei = lane & 7;
  ocg_me_mb *m = J.mb + mbi;
  const int off = T->frag_off[bi];
  const int bmv = m->block_mv[bi];
  const int vx = mv_x(bmv) / 2, vy = mv_y(bmv) / 2;
  const uint32_t err = refine_score(J.src + off, J.ref_satd[1] + off, J.ystride, vx, vy, c_sq_dx[sitei], c_sq_dy[sitei], false);
  uint32_t key_e = err;
  int key_s = sitei;
#pragma unroll
  for (int d = 1; d < 8; d <<= 1) {
    const uint32_t oe = __shfl_xor_sync(0xFFFFFFFFu, key_e, d);
    const int os = __shfl_xor_sync(0xFFFFFFFFu, key_s, d);
    if (oe < key_e || (oe == key_e && os < key_s)) { key_e = oe; key_s = os; }
  }
  if (sitei == 0) {
    const uint32_t entry = m->block_satd[bi];
    int dx = 0, dy = 0;
    uint32_t best = entry;
    if (key_e < entry) { best = key_e; dx = c_sq_dx[key_s]; dy = c_sq_dy[key_s]; }
    m->ref_mv[bi] = (int16_t)mv_make((vx << 1) + dx, (vy << 1) + dy);
    m->ref_block_satd[bi] = best;
  }
}

/* ocg_me_repair: the refinement's input is the search's output */
__global__ void ocg_me_repair_chain_kernel(const ocg_mb_search_in *in, const ocg_mb_search_out *so, ocg_mb_refine_in *ri) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; i++) { ri->frag_off[i] = in->frag_off[i]; ri->block_vec[i][0] = ri->block_vec[i][1] = 0; ri->block_satd[i] = 0; }
    ri->vec[0] = so->best_vec[0];
    ri->vec[1] = so->best_vec[1];
    ri->satd = so->satd;
  }
}

} /* namespace */

extern "C" {

/* ---- whole-frame motion analysis ------------------------------------------ */
struct ocg_me {
  ocg_ctx *ctx = nullptr;
  int nmbs = 0, nhsbs = 0, nvsbs = 0;
  ocg_me_topo *d_topo = nullptr;
  ocg_me_mb *d_mb = nullptr;
  uint32_t *d_done = nullptr;
  uint8_t *d_gold = nullptr;
  uint8_t *h_gold = nullptr;   /* pinned */
  OcgMeJob *d_job = nullptr;
  OcgMeJob *h_job = nullptr;   /* pinned */
  cudaEvent_t job_used = nullptr;
  bool job_busy = false;
  uint32_t seq = 0;
  uint32_t *d_ticket = nullptr; /* 4 words, zero between launches */
  int device = 0;
  int last_bufs[5] = {-1, -1, -1, -1, -1};
  int last_flags = 0;
  /* ocg_me_repair scratch: one macro block */
  uint8_t *h_rep = nullptr;     /* pinned: search_in | refine_in | search_out | refine_out */
  uint8_t *d_rep = nullptr;
};

/* batch scratch (one launch set for many streams), one per device; calls are serialised by g_me_batch_lock */
static std::mutex g_me_batch_lock;
struct MeBatch {
  OcgMeJob *d = nullptr, *h = nullptr;
  int cap = 0;
  cudaEvent_t used = nullptr;
  bool busy = false;
};
static MeBatch g_me_batches[16];

OCG_API int ocg_me_nmbs(const ocg_geometry *g) {
  if (g == nullptr) return OCG_EFAULT;
  return ((g->planes[0].nhfrags + 3) >> 2) * ((g->planes[0].nvfrags + 3) >> 2) * 4;
}

OCG_API int ocg_me_topology(const ocg_geometry *g, ocg_me_topo *topo) {
  if (g == nullptr || topo == nullptr) return OCG_EFAULT;
  const ocg_plane_geom &p = g->planes[0];
  const int nhsbs = (p.nhfrags + 3) >> 2, nvsbs = (p.nvfrags + 3) >> 2;
  const int nhmbs = nhsbs << 1, nvmbs = nvsbs << 1; /* state.c:501-502: counted in whole super blocks */
  static const unsigned char MBMAP[2][2] = {{0, 3}, {1, 2}};           /* internal.c:63 */
  static const unsigned char NCN[4] = {4, 3, 2, 4};                    /* encode.c:994 */
  static const signed char CDX[4][4] = {{-1, 0, 1, -1}, {-1, 0, -1, 0}, {-1, -1, 0, 0}, {-1, 0, 0, 1}};
  static const signed char CDY[4][4] = {{0, -1, -1, -1}, {0, -1, -1, 0}, {0, -1, 0, 0}, {0, -1, 1, -1}};
  const int nmbs = nhsbs * nvsbs * 4;
  memset(topo, 0, sizeof(*topo) * (size_t)nmbs);
  for (int sby = 0; sby < nvsbs; sby++)
    for (int sbx = 0; sbx < nhsbs; sbx++)
      for (int q = 0; q < 4; q++) {
        const int mbi = (sby * nhsbs + sbx) * 4 + q;
        const int mbx = 2 * sbx + (q >> 1), mby = 2 * sby + (((q + 1) >> 1) & 1);
        if (2 * mbx >= p.nhfrags || 2 * mby >= p.nvfrags) continue; /* state.c:318: outside the coded region */
        ocg_me_topo &t = topo[mbi];
        t.valid = 1;
        for (int i = 0; i < 2; i++)
          for (int j = 0; j < 2; j++)
            t.frag_off[i << 1 | j] = (int32_t)(p.plane_off + (int64_t)(2 * mby + i) * 8 * p.ystride + (2 * mbx + j) * 8);
      }
  for (int sby = 0; sby < nvsbs; sby++)
    for (int sbx = 0; sbx < nhsbs; sbx++)
      for (int q = 0; q < 4; q++) {
        const int mbi = (sby * nhsbs + sbx) * 4 + q;
        if (!topo[mbi].valid) continue;
        const int mbx = 2 * sbx + (q >> 1), mby = 2 * sby + (((q + 1) >> 1) & 1);
        for (int ni = 0; ni < NCN[q]; ni++) {
          const int nx = mbx + CDX[q][ni], ny = mby + CDY[q][ni];
          if (nx < 0 || nx >= nhmbs || ny < 0 || ny >= nvmbs) continue;
          /* encode.c:1031 */
          const int nmbi = (ny & ~1) * nhmbs + ((nx & ~1) << 1) + MBMAP[ny & 1][nx & 1];
          if (nmbi < 0 || nmbi >= nmbs || !topo[nmbi].valid) continue;
          topo[mbi].cn[topo[mbi].ncn++] = nmbi;
        }
      }
  return OCG_OK;
}

OCG_API void ocg_me_destroy(ocg_me *me) {
  if (me == nullptr) return;
  ocg_set_device(me->device);
  if (me->ctx != nullptr) ocg_ctx_sync(me->ctx);
  cudaFree(me->d_ticket);
  cudaFree(me->d_rep);
  if (me->h_rep) cudaFreeHost(me->h_rep);
  cudaFree(me->d_topo);
  cudaFree(me->d_mb);
  cudaFree(me->d_done);
  cudaFree(me->d_gold);
  cudaFree(me->d_job);
  if (me->h_gold) cudaFreeHost(me->h_gold);
  if (me->h_job) cudaFreeHost(me->h_job);
  if (me->job_used) cudaEventDestroy(me->job_used);
  delete me;
}

#define ME_CU(call)                                   \
  do {                                                \
    if ((call) != cudaSuccess) {                      \
      cudaGetLastError();                             \
      ocg_me_destroy(me);                             \
      return OCG_ECUDA;                               \
    }                                                 \
  } while (0)

OCG_API int ocg_me_create(ocg_me **out, ocg_ctx *ctx, const ocg_me_topo *topo) {
  if (out == nullptr || ctx == nullptr) return OCG_EFAULT;
  *out = nullptr;
  const ocg_geometry *g = ocg_ctx_geometry(ctx);
  if (g->nrefs < 5) return OCG_EINVAL; /* IO + two originals + two reconstructions */
  if (ocg_set_device(ocg_ctx_device(ctx)) != cudaSuccess) return OCG_ECUDA;
  ocg_me *me = new (std::nothrow) ocg_me();
  if (me == nullptr) return OCG_ENOMEM;
  me->ctx = ctx;
  me->device = ocg_ctx_device(ctx);
  me->nhsbs = (g->planes[0].nhfrags + 3) >> 2;
  me->nvsbs = (g->planes[0].nvfrags + 3) >> 2;
  me->nmbs = me->nhsbs * me->nvsbs * 4;
  const size_t n = (size_t)me->nmbs;
  std::vector<ocg_me_topo> own;
  if (topo == nullptr) {
    own.resize(n);
    ocg_me_topology(g, own.data());
    topo = own.data();
  }
  /* the wave-front only ever waits on macro blocks that precede it in coding order */
  for (size_t i = 0; i < n; i++)
    for (int k = 0; k < topo[i].ncn; k++)
      if (topo[i].ncn > 4 || topo[i].cn[k] < 0 || (size_t)topo[i].cn[k] >= i || !topo[topo[i].cn[k]].valid) {
        delete me;
        return OCG_EINVAL;
      }
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(ctx);
  ME_CU(cudaMalloc(&me->d_topo, n * sizeof(ocg_me_topo)));
  ME_CU(cudaMalloc(&me->d_mb, n * sizeof(ocg_me_mb)));
  ME_CU(cudaMalloc(&me->d_done, n * 2 * sizeof(uint32_t)));
  ME_CU(cudaMalloc(&me->d_gold, n));
  ME_CU(cudaMalloc(&me->d_job, sizeof(OcgMeJob)));
  ME_CU(cudaHostAlloc(&me->h_gold, n, cudaHostAllocDefault));
  ME_CU(cudaHostAlloc(&me->h_job, sizeof(OcgMeJob), cudaHostAllocDefault));
  ME_CU(cudaEventCreateWithFlags(&me->job_used, cudaEventDisableTiming));
  ME_CU(cudaMalloc(&me->d_ticket, 4 * sizeof(uint32_t)));
  ME_CU(cudaMemsetAsync(me->d_ticket, 0, 4 * sizeof(uint32_t), st));
  ME_CU(cudaMalloc(&me->d_rep, 256));
  ME_CU(cudaHostAlloc(&me->h_rep, 256, cudaHostAllocDefault));
  ME_CU(cudaMemcpyAsync(me->d_topo, topo, n * sizeof(ocg_me_topo), cudaMemcpyHostToDevice, st));
  ME_CU(cudaMemsetAsync(me->d_mb, 0, n * sizeof(ocg_me_mb), st));
  ME_CU(cudaMemsetAsync(me->d_done, 0, n * 2 * sizeof(uint32_t), st));
  ME_CU(cudaStreamSynchronize(st)); /* `own` goes out of scope */
  *out = me;
  return OCG_OK;
}
#undef ME_CU

static int me_fill_job(ocg_me *me, const int bufs[5], int flags, bool gold, OcgMeJob *j) {
  const ocg_geometry *g = ocg_ctx_geometry(me->ctx);
  const uint8_t *b[5];
  for (int i = 0; i < 5; i++) {
    const uint8_t *p = (const uint8_t *)ocg_ctx_frame_devptr(me->ctx, bufs[i]);
    if (p == nullptr) return OCG_EINVAL;
    b[i] = p + g->base_off;
  }
  j->src = b[0];
  j->ref_full[1] = b[1]; /* OC_FRAME_PREV <- PREV_ORIG */
  j->ref_full[0] = b[2]; /* OC_FRAME_GOLD <- GOLD_ORIG */
  j->ref_satd[1] = b[3];
  j->ref_satd[0] = b[4];
  j->topo = me->d_topo;
  j->mb = me->d_mb;
  j->done = me->d_done;
  j->gold_refine = gold ? me->d_gold : nullptr;
  j->ticket = me->d_ticket;
  j->ystride = g->planes[0].ystride;
  j->nhsbs = me->nhsbs;
  j->nvsbs = me->nvsbs;
  j->nmbs = me->nmbs;
  j->flags = flags;
  j->seq = ++me->seq;
  for (int i = 0; i < 5; i++) me->last_bufs[i] = bufs[i];
  me->last_flags = flags;
  return OCG_OK;
}

static int me_launch(const OcgMeJob *d_jobs, int n, int nvsbs, int nmbs, int flags, cudaStream_t st) {
  ocg_me_wavefront_kernel<<<dim3((unsigned)nvsbs, 2, (unsigned)n), 32, 0, st>>>(d_jobs);
  ocg_count_launch(1);
  if ((flags & OCG_ME_REFINE_4MV) && !(flags & OCG_ME_FAST)) {
    ocg_me_refine4_kernel<<<dim3((unsigned)((nmbs + 3) / 4), (unsigned)n), 128, 0, st>>>(d_jobs);
    ocg_count_launch(1);
  }
  return cudaGetLastError() == cudaSuccess ? OCG_OK : OCG_ECUDA;
}

OCG_API int ocg_me_frame(ocg_me *me, const int bufs[5], int flags, const uint8_t *gold_refine) {
  if (me == nullptr || bufs == nullptr) return OCG_EFAULT;
  if (flags & ~63) return OCG_EINVAL;
  if (ocg_set_device(me->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(me->ctx);
  if (me->job_busy) {
    if (cudaEventSynchronize(me->job_used) != cudaSuccess) return OCG_ECUDA;
    me->job_busy = false;
  }
  int r = me_fill_job(me, bufs, flags, gold_refine != nullptr, me->h_job);
  if (r < 0) return r;
  if (gold_refine != nullptr) {
    memcpy(me->h_gold, gold_refine, (size_t)me->nmbs);
    if (cudaMemcpyAsync(me->d_gold, me->h_gold, (size_t)me->nmbs, cudaMemcpyHostToDevice, st) != cudaSuccess) return OCG_ECUDA;
  }
  if (cudaMemcpyAsync(me->d_job, me->h_job, sizeof(OcgMeJob), cudaMemcpyHostToDevice, st) != cudaSuccess) return OCG_ECUDA;
  cudaEventRecord(me->job_used, st);
  me->job_busy = true;
  return me_launch(me->d_job, 1, me->nvsbs, me->nmbs, flags, st);
}

OCG_API int ocg_me_frame_batch(ocg_me *const *mes, const int *bufs, int n, int flags, void *stream) {
  if (mes == nullptr || bufs == nullptr || mes[0] == nullptr) return OCG_EFAULT;
  if (n <= 0 || (flags & ~63)) return OCG_EINVAL;
  const int device = mes[0]->device;
  if (device < 0 || device >= 16) return OCG_EINVAL;
  for (int i = 0; i < n; i++)
    if (mes[i] == nullptr || mes[i]->device != device) return OCG_EINVAL; /* one launch set = one device */
  if (ocg_set_device(device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = stream != nullptr ? (cudaStream_t)stream : (cudaStream_t)ocg_ctx_stream(mes[0]->ctx);
  std::lock_guard<std::mutex> lk(g_me_batch_lock);
  MeBatch &g_me_batch = g_me_batches[device];
  if (g_me_batch.busy) {
    if (cudaEventSynchronize(g_me_batch.used) != cudaSuccess) return OCG_ECUDA;
    g_me_batch.busy = false;
  }
  if (g_me_batch.cap < n) {
    if (g_me_batch.d) cudaFree(g_me_batch.d);
    if (g_me_batch.h) cudaFreeHost(g_me_batch.h);
    g_me_batch.d = g_me_batch.h = nullptr;
    g_me_batch.cap = 0;
    if (cudaMalloc(&g_me_batch.d, sizeof(OcgMeJob) * (size_t)n) != cudaSuccess) return OCG_ECUDA;
    if (cudaHostAlloc(&g_me_batch.h, sizeof(OcgMeJob) * (size_t)n, cudaHostAllocDefault) != cudaSuccess) return OCG_ECUDA;
    if (g_me_batch.used == nullptr && cudaEventCreateWithFlags(&g_me_batch.used, cudaEventDisableTiming) != cudaSuccess)
      return OCG_ECUDA;
    g_me_batch.cap = n;
  }
  for (int i = 0; i < n; i++) {
    if (mes[i] == nullptr || mes[i]->nmbs != mes[0]->nmbs || mes[i]->nvsbs != mes[0]->nvsbs) return OCG_EINVAL;
    int r = me_fill_job(mes[i], bufs + 5 * i, flags, false, g_me_batch.h + i);
    if (r < 0) return r;
  }
  if (cudaMemcpyAsync(g_me_batch.d, g_me_batch.h, sizeof(OcgMeJob) * (size_t)n, cudaMemcpyHostToDevice, st) != cudaSuccess)
    return OCG_ECUDA;
  cudaEventRecord(g_me_batch.used, st);
  g_me_batch.busy = true;
  return me_launch(g_me_batch.d, n, mes[0]->nvsbs, mes[0]->nmbs, flags, st);
}

OCG_API int ocg_me_read(ocg_me *me, ocg_me_mb *out) {
  if (me == nullptr || out == nullptr) return OCG_EFAULT;
  if (ocg_set_device(me->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(me->ctx);
  if (cudaMemcpyAsync(out, me->d_mb, (size_t)me->nmbs * sizeof(ocg_me_mb), cudaMemcpyDeviceToHost, st) != cudaSuccess) return OCG_ECUDA;
  return cudaStreamSynchronize(st) == cudaSuccess ? OCG_OK : OCG_ECUDA;
}

OCG_API int ocg_me_write(ocg_me *me, const ocg_me_mb *in) {
  if (me == nullptr || in == nullptr) return OCG_EFAULT;
  if (ocg_set_device(me->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(me->ctx);
  if (cudaMemcpyAsync(me->d_mb, in, (size_t)me->nmbs * sizeof(ocg_me_mb), cudaMemcpyHostToDevice, st) != cudaSuccess) return OCG_ECUDA;
  return cudaStreamSynchronize(st) == cudaSuccess ? OCG_OK : OCG_ECUDA;
}

/* Asynchronous forms for a caller that brackets them with its own synchronisation (pinned `in`/`out`). */
OCG_API int ocg_me_write_async(ocg_me *me, const ocg_me_mb *in) {
  if (me == nullptr || in == nullptr) return OCG_EFAULT;
  if (ocg_set_device(me->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(me->ctx);
  return cudaMemcpyAsync(me->d_mb, in, (size_t)me->nmbs * sizeof(ocg_me_mb), cudaMemcpyHostToDevice, st) == cudaSuccess ? OCG_OK : OCG_ECUDA;
}
OCG_API int ocg_me_read_async(ocg_me *me, ocg_me_mb *out) {
  if (me == nullptr || out == nullptr) return OCG_EFAULT;
  if (ocg_set_device(me->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(me->ctx);
  return cudaMemcpyAsync(out, me->d_mb, (size_t)me->nmbs * sizeof(ocg_me_mb), cudaMemcpyDeviceToHost, st) == cudaSuccess ? OCG_OK : OCG_ECUDA;
}

/* One macro block's search against one reference frame of the last ocg_me_frame call, with the candidate
   set given by the caller, then (refine != 0) the half-pel refinement of the result.  Synchronous.  This
   is the repair step of a caller that ran the GOLD chain speculatively (no refinements) and later learns
   that a neighbour's vector was refined after all (analyze.c:2476-2485 -> mcenc.c:104-110). */
OCG_API int ocg_me_repair(ocg_me *me, int frame, const ocg_mb_search_in *in, int refine, ocg_mb_search_out *out,
                          ocg_mb_refine_out *rout) {
  if (me == nullptr || in == nullptr || out == nullptr || (refine && rout == nullptr)) return OCG_EFAULT;
  if (frame < 0 || frame > 1 || me->last_bufs[0] < 0) return OCG_EINVAL;
  if (ocg_set_device(me->device) != cudaSuccess) return OCG_ECUDA;
  const ocg_geometry *g = ocg_ctx_geometry(me->ctx);
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(me->ctx);
  const uint8_t *b[5];
  for (int i = 0; i < 5; i++) b[i] = (const uint8_t *)ocg_ctx_frame_devptr(me->ctx, me->last_bufs[i]) + g->base_off;
  const uint8_t *ref_full = frame == 1 ? b[1] : b[2], *ref_satd = frame == 1 ? b[3] : b[4];
  ocg_mb_search_in *h_in = (ocg_mb_search_in *)me->h_rep, *d_in = (ocg_mb_search_in *)me->d_rep;
  ocg_mb_refine_in *d_rin = (ocg_mb_refine_in *)(me->d_rep + 48);
  ocg_mb_search_out *h_out = (ocg_mb_search_out *)(me->h_rep + 96), *d_out = (ocg_mb_search_out *)(me->d_rep + 96);
  ocg_mb_refine_out *h_rout = (ocg_mb_refine_out *)(me->h_rep + 128), *d_rout = (ocg_mb_refine_out *)(me->d_rep + 128);
  *h_in = *in;
  if (cudaMemcpyAsync(d_in, h_in, sizeof(*h_in), cudaMemcpyHostToDevice, st) != cudaSuccess) return OCG_ECUDA;
  const int ystride = g->planes[0].ystride;
  const bool nosatd = (me->last_flags & OCG_ME_NOSATD) != 0;
  ocg_mcenc_search_kernel<<<1, 128, 0, st>>>(b[0], ref_full, ref_satd, ystride, d_in, d_out, 1);
  ocg_count_launch(1);
  if (refine) {
    ocg_me_repair_chain_kernel<<<1, 32, 0, st>>>(d_in, d_out, d_rin);
    ocg_mcenc_refine_kernel<<<1, 128, 0, st>>>(b[0], ref_satd, ystride, d_rin, d_rout, 1,
                                               OCG_REFINE_1MV | (nosatd ? OCG_REFINE_SAD : 0));
    ocg_count_launch(2);
  }
  if (cudaMemcpyAsync(h_out, d_out, 64, cudaMemcpyDeviceToHost, st) != cudaSuccess) return OCG_ECUDA;
  if (cudaStreamSynchronize(st) != cudaSuccess) return OCG_ECUDA;
  *out = *h_out;
  if (refine) *rout = *h_rout;
  return OCG_OK;
}


/* ---- inter-frame analysis tables ------------------------------------------- */
struct ocg_enc_inter {
  ocg_ctx *ctx = nullptr;
  ocg_me *me = nullptr;
  int device = 0;
  int nfrags = 0, nluma = 0, nmbs = 0, nborder_y = 0, nborder_c = 0;
  int32_t *d_mbfrags = nullptr, *d_frag_off = nullptr;
  ocg_enc_frag *d_all = nullptr;     /* [nfrags] {foff, foff} */
  ocg_enc_frag *d_border = nullptr;  /* luma border fragments, then chroma ones: {foff, foff, mask lo, mask hi} */
  ocg_enc_frag *d_cand = nullptr;    /* [K][nfrags] in the two-block layout of ocg_enc_cand_kernel */
  uint8_t *d_out = nullptr, *h_out = nullptr;
  size_t off_isatd = 0, off_idc = 0, off_skip = 0, off_border = 0, off_key = 0, off_csatd = 0, off_cdc = 0, out_sz = 0;
  int32_t *h_border_slot = nullptr; /* [nfrags] index into border_ssd or -1 */
  /* speculative sub + fDCT + quantiser (ocg_enc_fq_*) */
  ocg_enc_cand_rec *d_rec = nullptr, *h_rec = nullptr; /* [nfrags][K] candidate records for the host (h_rec pinned) */
  ocg_enc_frag *d_fq = nullptr;          /* [3 qii][OCG_FQ_NSEL][nfrags], luma block then chroma block */
  int16_t *d_fq_dct = nullptr, *d_fq_qdct = nullptr;
  int32_t *d_fq_nz = nullptr;
  uint16_t *d_dequant = nullptr;
  int16_t *d_enquant = nullptr;
  uint8_t *h_qtab = nullptr;             /* pinned staging of the two quantiser tables */
  ocg_fq_desc_dev *d_fq_desc = nullptr, *h_fq_desc = nullptr;
  uint4 *d_fq_pool = nullptr, *h_fq_pool = nullptr;
  uint32_t *d_fq_counter = nullptr, *h_fq_counter = nullptr;
  uint32_t fq_pool_units = 0;
  int fq_nqis = 0;                       /* > 0: the last prepass produced the speculative tables */
};

OCG_API void ocg_enc_inter_destroy(ocg_enc_inter *ei) {
  if (ei == nullptr) return;
  ocg_set_device(ei->device);
  if (ei->ctx != nullptr) ocg_ctx_sync(ei->ctx);
  cudaFree(ei->d_mbfrags);
  cudaFree(ei->d_frag_off);
  cudaFree(ei->d_all);
  cudaFree(ei->d_border);
  cudaFree(ei->d_cand);
  cudaFree(ei->d_out);
  if (ei->h_out) cudaFreeHost(ei->h_out);
  cudaFree(ei->d_rec);
  if (ei->h_rec) cudaFreeHost(ei->h_rec);
  cudaFree(ei->d_fq); cudaFree(ei->d_fq_dct); cudaFree(ei->d_fq_qdct); cudaFree(ei->d_fq_nz);
  cudaFree(ei->d_dequant); cudaFree(ei->d_enquant); cudaFree(ei->d_fq_desc); cudaFree(ei->d_fq_pool); cudaFree(ei->d_fq_counter);
  if (ei->h_qtab) cudaFreeHost(ei->h_qtab);
  if (ei->h_fq_desc) cudaFreeHost(ei->h_fq_desc);
  if (ei->h_fq_pool) cudaFreeHost(ei->h_fq_pool);
  if (ei->h_fq_counter) cudaFreeHost(ei->h_fq_counter);
  free(ei->h_border_slot);
  delete ei;
}

OCG_API int ocg_enc_inter_create(ocg_enc_inter **out, ocg_ctx *ctx, ocg_me *me, const int32_t *mbfrags,
                                 const int32_t *border_fragi, const int64_t *border_mask, int nborder) {
  if (out == nullptr || ctx == nullptr || me == nullptr || mbfrags == nullptr || (nborder > 0 && (border_fragi == nullptr || border_mask == nullptr)))
    return OCG_EFAULT;
  *out = nullptr;
  const ocg_geometry *g = ocg_ctx_geometry(ctx);
  if (nborder < 0 || nborder > g->nfrags) return OCG_EINVAL;
  if (ocg_set_device(ocg_ctx_device(ctx)) != cudaSuccess) return OCG_ECUDA;
  ocg_enc_inter *ei = new (std::nothrow) ocg_enc_inter();
  if (ei == nullptr) return OCG_ENOMEM;
  ei->ctx = ctx;
  ei->me = me;
  ei->device = ocg_ctx_device(ctx);
  ei->nfrags = g->nfrags;
  ei->nluma = g->planes[0].nfrags;
  ei->nmbs = me->nmbs;
  const size_t nf = (size_t)g->nfrags, K = OCG_ENC_NCAND;
  std::vector<int32_t> foff(nf);
  ocg_geometry_frag_buf_offs(g, foff.data());
  std::vector<ocg_enc_frag> all(nf), border((size_t)nborder);
  for (size_t i = 0; i < nf; i++) all[i] = ocg_enc_frag{foff[i], foff[i], INT_MIN, 0};
  ei->h_border_slot = (int32_t *)malloc(nf * sizeof(int32_t));
  if (ei->h_border_slot == nullptr) { delete ei; return OCG_ENOMEM; }
  for (size_t i = 0; i < nf; i++) ei->h_border_slot[i] = -1;
  int nb = 0;
  for (int pass = 0; pass < 2; pass++) { /* luma fragments first: the two classes differ in stride */
    for (int i = 0; i < nborder; i++) {
      const int fi = border_fragi[i];
      if (fi < 0 || fi >= g->nfrags) { ocg_enc_inter_destroy(ei); return OCG_EINVAL; }
      if ((fi < ei->nluma) != (pass == 0)) continue;
      border[(size_t)nb] = ocg_enc_frag{foff[(size_t)fi], foff[(size_t)fi], (int32_t)(uint32_t)(uint64_t)border_mask[i],
                                        (int32_t)(uint32_t)((uint64_t)border_mask[i] >> 32)};
      ei->h_border_slot[fi] = nb++;
    }
    if (pass == 0) ei->nborder_y = nb;
  }
  ei->nborder_c = nb - ei->nborder_y;
  size_t o = 0;
  ei->off_isatd = o; o += nf * 4;
  ei->off_idc = o; o += nf * 4;
  ei->off_skip = o; o += nf * 4;
  ei->off_border = o; o += ((size_t)nborder + 1) * 4;
  o = (o + 255) & ~(size_t)255;
  ei->off_csatd = o; o += K * nf * 4;
  ei->off_cdc = o; o += K * nf * 4;
  ei->out_sz = o;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(ctx);
#define EI_CU(call) do { if ((call) != cudaSuccess) { cudaGetLastError(); ocg_enc_inter_destroy(ei); return OCG_ECUDA; } } while (0)
  EI_CU(cudaMalloc(&ei->d_mbfrags, (size_t)ei->nmbs * 12 * sizeof(int32_t)));
  EI_CU(cudaMalloc(&ei->d_frag_off, nf * sizeof(int32_t)));
  EI_CU(cudaMalloc(&ei->d_all, nf * sizeof(ocg_enc_frag)));
  EI_CU(cudaMalloc(&ei->d_border, ((size_t)nborder + 1) * sizeof(ocg_enc_frag)));
  EI_CU(cudaMalloc(&ei->d_cand, K * nf * sizeof(ocg_enc_frag)));
  EI_CU(cudaMalloc(&ei->d_out, ei->out_sz));
  EI_CU(cudaHostAlloc(&ei->h_out, ei->out_sz, cudaHostAllocDefault));
  EI_CU(cudaMalloc(&ei->d_rec, K * nf * sizeof(ocg_enc_cand_rec)));
  EI_CU(cudaHostAlloc(&ei->h_rec, K * nf * sizeof(ocg_enc_cand_rec), cudaHostAllocDefault));
  EI_CU(cudaMemcpyAsync(ei->d_mbfrags, mbfrags, (size_t)ei->nmbs * 12 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  EI_CU(cudaMemcpyAsync(ei->d_frag_off, foff.data(), nf * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  EI_CU(cudaMemcpyAsync(ei->d_all, all.data(), nf * sizeof(ocg_enc_frag), cudaMemcpyHostToDevice, st));
  if (nborder > 0) EI_CU(cudaMemcpyAsync(ei->d_border, border.data(), (size_t)nborder * sizeof(ocg_enc_frag), cudaMemcpyHostToDevice, st));
  EI_CU(cudaMemsetAsync(ei->d_cand, 0, K * nf * sizeof(ocg_enc_frag), st));
  {
    const size_t nfq = 3 * OCG_FQ_NSEL * nf;
    ei->fq_pool_units = (uint32_t)((16u << 20) / 16);
    EI_CU(cudaMalloc(&ei->d_fq, nfq * sizeof(ocg_enc_frag)));
    EI_CU(cudaMalloc(&ei->d_fq_dct, nfq * 128));
    EI_CU(cudaMalloc(&ei->d_fq_qdct, nfq * 128));
    EI_CU(cudaMalloc(&ei->d_fq_nz, nfq * 4));
    EI_CU(cudaMalloc(&ei->d_dequant, 3 * 2 * 3 * 64 * 2));
    EI_CU(cudaMalloc(&ei->d_enquant, 3 * 2 * 3 * 64 * 2 * 2));
    EI_CU(cudaMalloc(&ei->d_fq_desc, OCG_FQ_NSEL * nf * sizeof(ocg_fq_desc_dev)));
    EI_CU(cudaMalloc(&ei->d_fq_pool, (size_t)ei->fq_pool_units * 16));
    EI_CU(cudaMalloc(&ei->d_fq_counter, 4));
    EI_CU(cudaHostAlloc(&ei->h_qtab, 3 * 2 * 3 * 64 * 2 * 3, cudaHostAllocDefault));
    EI_CU(cudaHostAlloc(&ei->h_fq_desc, OCG_FQ_NSEL * nf * sizeof(ocg_fq_desc_dev), cudaHostAllocDefault));
    EI_CU(cudaHostAlloc(&ei->h_fq_pool, (size_t)ei->fq_pool_units * 16, cudaHostAllocDefault));
    EI_CU(cudaHostAlloc(&ei->h_fq_counter, 64, cudaHostAllocDefault));
  }
  EI_CU(cudaStreamSynchronize(st));
#undef EI_CU
  *out = ei;
  return OCG_OK;
}

OCG_API int ocg_enc_inter_border_slot(const ocg_enc_inter *ei, int fragi) {
  return ei != nullptr && fragi >= 0 && fragi < ei->nfrags ? ei->h_border_slot[fragi] : -1;
}

OCG_API int ocg_enc_inter_quant_tables(ocg_enc_inter *ei, const uint16_t *dequant, const int16_t *enquant, int nqis) {
  if (ei == nullptr) return OCG_EFAULT;
  ei->fq_nqis = 0;
  if (dequant == nullptr || enquant == nullptr || nqis <= 0) return OCG_OK; /* no speculative transform this frame */
  if (nqis > 3) return OCG_EINVAL;
  if (ocg_set_device(ei->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(ei->ctx);
  const size_t nd = 3 * 2 * 3 * 64 * 2, ne = nd * 2;
  memcpy(ei->h_qtab, dequant, nd);
  memcpy(ei->h_qtab + nd, enquant, ne);
  if (cudaMemcpyAsync(ei->d_dequant, ei->h_qtab, nd, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(ei->d_enquant, ei->h_qtab + nd, ne, cudaMemcpyHostToDevice, st) != cudaSuccess)
    return OCG_ECUDA;
  ei->fq_nqis = nqis;
  return OCG_OK;
}

OCG_API int ocg_enc_inter_prepass(ocg_enc_inter *ei, int io_buf, int prev_buf, int gold_buf, int with_cands,
                                  ocg_enc_inter_tables *out) {
  if (ei == nullptr || out == nullptr) return OCG_EFAULT;
  const ocg_geometry *g = ocg_ctx_geometry(ei->ctx);
  if (io_buf < 0 || io_buf >= g->nrefs || prev_buf < 0 || prev_buf >= g->nrefs || gold_buf < 0 || gold_buf >= g->nrefs) return OCG_EINVAL;
  if (ocg_set_device(ei->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(ei->ctx);
  const uint8_t *pool = (const uint8_t *)ocg_ctx_frame_devptr(ei->ctx, 0) + g->base_off;
  const uint8_t *io = pool + (size_t)io_buf * g->ref_frame_sz, *prev = pool + (size_t)prev_buf * g->ref_frame_sz;
  const int nl = ei->nluma, nc = ei->nfrags - nl, K = OCG_ENC_NCAND;
  const int ys[2] = {g->planes[0].ystride, g->planes[1].ystride};
  uint32_t *d_isatd = (uint32_t *)(ei->d_out + ei->off_isatd), *d_skip = (uint32_t *)(ei->d_out + ei->off_skip);
  uint32_t *d_border = (uint32_t *)(ei->d_out + ei->off_border), *d_csatd = (uint32_t *)(ei->d_out + ei->off_csatd);
  int32_t *d_idc = (int32_t *)(ei->d_out + ei->off_idc), *d_cdc = (int32_t *)(ei->d_out + ei->off_cdc);
  const int first[2] = {0, nl}, count[2] = {nl, nc};
  const int bfirst[2] = {0, ei->nborder_y}, bcount[2] = {ei->nborder_y, ei->nborder_c};
  int r = OCG_OK;
  for (int k = 0; k < 2 && r == OCG_OK; k++) {
    /* oc_mb_intra_satd (analyze.c:1360-1400), oc_skip_cost (analyze.c:1968-2040) */
    r = ocg_enc_metrics_batch(OCG_MET_INTRA_SATD, io, nullptr, ys[k], ei->d_all + first[k], count[k], d_isatd + first[k], d_idc + first[k], st);
    if (r == OCG_OK) r = ocg_enc_metrics_batch(OCG_MET_SSD, io, prev, ys[k], ei->d_all + first[k], count[k], d_skip + first[k], nullptr, st);
    if (r == OCG_OK && bcount[k] > 0)
      r = ocg_enc_metrics_batch(OCG_MET_BORDER_SSD, io, prev, ys[k], ei->d_border + bfirst[k], bcount[k], d_border + bfirst[k], nullptr, st);
  }
  if (r != OCG_OK) return r;
  if (with_cands) {
    OcgCandJob J;
    J.mb = ei->me->d_mb;
    J.mbfrags = ei->d_mbfrags;
    J.frag_off = ei->d_frag_off;
    J.out = ei->d_cand;
    J.nmbs = ei->nmbs;
    J.nfrags = ei->nfrags;
    J.nluma = nl;
    J.ystride_y = ys[0];
    J.ystride_c = ys[1];
    J.qx = !(g->pixel_fmt & 1);
    J.qy = !(g->pixel_fmt & 2);
    J.fmt = g->pixel_fmt;
    J.io_off = (int32_t)((int64_t)io_buf * g->ref_frame_sz);
    J.prev_off = (int32_t)((int64_t)prev_buf * g->ref_frame_sz);
    J.gold_off = (int32_t)((int64_t)gold_buf * g->ref_frame_sz);
    ocg_enc_cand_kernel<<<(unsigned)((ei->nmbs + 127) / 128), 128, 0, st>>>(J);
    ocg_count_launch(1);
    r = ocg_enc_metrics_batch(OCG_MET_SATD, pool, pool, ys[0], ei->d_cand, K * nl, d_csatd, d_cdc, st);
    if (r == OCG_OK && nc > 0)
      r = ocg_enc_metrics_batch(OCG_MET_SATD, pool, pool, ys[1], ei->d_cand + (size_t)K * nl, K * nc, d_csatd + (size_t)K * nl, d_cdc + (size_t)K * nl, st);
    if (r != OCG_OK) return r;
    if (ei->fq_nqis > 0) {
      /* sub + fDCT + quantiser against the predictors of candidates 0 and 3, every quantiser of the frame */
      const int nq = ei->fq_nqis;
      ocg_enc_fq_list_kernel<<<(unsigned)((ei->nfrags + 255) / 256), 256, 0, st>>>(ei->d_cand, ei->d_fq, ei->nfrags, nl, nq);
      ocg_count_launch(1);
      const size_t cbase = (size_t)3 * OCG_FQ_NSEL * nl;
      r = ocg_enc_fdct_quant_batch(pool, pool, ys[0], ei->d_fq, nq * OCG_FQ_NSEL * nl, ei->d_dequant, ei->d_enquant, ei->d_fq_dct,
                                   ei->d_fq_qdct, ei->d_fq_nz, st);
      if (r == OCG_OK && nc > 0)
        r = ocg_enc_fdct_quant_batch(pool, pool, ys[1], ei->d_fq + cbase, nq * OCG_FQ_NSEL * nc, ei->d_dequant, ei->d_enquant,
                                     ei->d_fq_dct + cbase * 64, ei->d_fq_qdct + cbase * 64, ei->d_fq_nz + cbase, st);
      if (r != OCG_OK) return r;
      if (cudaMemsetAsync(ei->d_fq_counter, 0, 4, st) != cudaSuccess) return OCG_ECUDA;
      ocg_enc_fq_compact_kernel<<<(unsigned)((ei->nfrags + 255) / 256), 256, 0, st>>>(
          ei->d_cand, ei->d_fq_dct, ei->d_fq_qdct, ei->d_fq_nz, ei->nfrags, nl, nq, ei->d_fq_desc, ei->d_fq_pool, ei->fq_pool_units, ei->d_fq_counter);
      ocg_count_launch(1);
      if (cudaMemcpyAsync(ei->h_fq_counter, ei->d_fq_counter, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaMemcpyAsync(ei->h_fq_desc, ei->d_fq_desc, (size_t)OCG_FQ_NSEL * ei->nfrags * sizeof(ocg_fq_desc_dev), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return OCG_ECUDA;
    }
  }
  const size_t small = ei->off_border + ((size_t)(ei->nborder_y + ei->nborder_c) + 1) * 4;
  if (cudaMemcpyAsync(ei->h_out, ei->d_out, small, cudaMemcpyDeviceToHost, st) != cudaSuccess) return OCG_ECUDA;
  if (with_cands) {
    ocg_enc_cand_pack_kernel<<<(unsigned)((ei->nfrags + 255) / 256), 256, 0, st>>>(ei->d_cand, d_csatd, d_cdc, ei->nfrags, nl, ei->d_rec);
    ocg_count_launch(1);
    if (cudaMemcpyAsync(ei->h_rec, ei->d_rec, (size_t)K * ei->nfrags * sizeof(ocg_enc_cand_rec), cudaMemcpyDeviceToHost, st) != cudaSuccess)
      return OCG_ECUDA;
  }
  out->intra_satd = (const uint32_t *)(ei->h_out + ei->off_isatd);
  out->intra_dc = (const int32_t *)(ei->h_out + ei->off_idc);
  out->skip_ssd = (const uint32_t *)(ei->h_out + ei->off_skip);
  out->border_ssd = (const uint32_t *)(ei->h_out + ei->off_border);
  out->ncand = with_cands ? K : 0;
  out->cand = ei->h_rec;
  out->nluma = nl;
  out->nfrags = ei->nfrags;
  out->d2h_bytes = (long)(small + (with_cands ? (size_t)K * ei->nfrags * sizeof(ocg_enc_cand_rec) : 0));
  out->fq_nqis = with_cands ? ei->fq_nqis : 0;
  out->fq_desc = (const ocg_enc_fq_desc *)ei->h_fq_desc;
  out->fq_pool = (const int16_t *)ei->h_fq_pool;
  return OCG_OK;
}

/* Second half of the pre-pass, after the caller has waited for the first: fetches exactly the part of the
   coefficient pool that was filled.  Synchronous. */
OCG_API int ocg_enc_inter_finish(ocg_enc_inter *ei, ocg_enc_inter_tables *out) {
  if (ei == nullptr || out == nullptr) return OCG_EFAULT;
  if (out->fq_nqis <= 0) return OCG_OK;
  if (ocg_set_device(ei->device) != cudaSuccess) return OCG_ECUDA;
  cudaStream_t st = (cudaStream_t)ocg_ctx_stream(ei->ctx);
  uint32_t used = *ei->h_fq_counter;
  if (used > ei->fq_pool_units) used = ei->fq_pool_units;
  if (used > 0 && cudaMemcpyAsync(ei->h_fq_pool, ei->d_fq_pool, (size_t)used * 16, cudaMemcpyDeviceToHost, st) != cudaSuccess) return OCG_ECUDA;
  if (cudaStreamSynchronize(st) != cudaSuccess) return OCG_ECUDA;
  out->d2h_bytes += (long)used * 16 + (long)OCG_FQ_NSEL * ei->nfrags * (long)sizeof(ocg_fq_desc_dev);
  return OCG_OK;
}

OCG_API int ocg_mcenc_refine_batch(const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                                   const ocg_mb_refine_in *in, ocg_mb_refine_out *out, int n, int flags,
                                   void *stream) {
  if (src_base == nullptr || ref_base == nullptr || in == nullptr || out == nullptr) return OCG_EFAULT;
  if (n < 0 || (flags & ~7) != 0 || (flags & (OCG_REFINE_1MV | OCG_REFINE_4MV)) == 0) return OCG_EINVAL;
  if (n == 0) return OCG_OK;
  ocg_mcenc_refine_kernel<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(src_base, ref_base, ystride, in,
                                                                                     out, n, flags);
  ocg_count_launch(1);
  return cudaGetLastError() == cudaSuccess ? OCG_OK : OCG_ECUDA;
}

OCG_API int ocg_enc_metrics_batch(int metric, const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                                  const ocg_enc_frag *frags, int n, uint32_t *out_val, int32_t *out_dc,
                                  void *stream) {
  if (src_base == nullptr || frags == nullptr || out_val == nullptr) return OCG_EFAULT;
  if (metric < OCG_MET_SAD || metric > OCG_MET_SAD_THRESH || n < 0) return OCG_EINVAL;
  if (n == 0) return OCG_OK;
  const unsigned grid = (unsigned)((n + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  switch (metric) {
    case OCG_MET_ACTIVITY: ocg_enc_activity_kernel<<<grid, 128, 0, st>>>(src_base, ystride, frags, n, out_val, out_dc); break;
    case OCG_MET_SAD: ocg_enc_metrics_kernel<OCG_MET_SAD><<<grid, 128, 0, st>>>(src_base, ref_base, ystride, frags, n, out_val, out_dc); break;
    case OCG_MET_SATD: ocg_enc_metrics_kernel<OCG_MET_SATD><<<grid, 128, 0, st>>>(src_base, ref_base, ystride, frags, n, out_val, out_dc); break;
    case OCG_MET_INTRA_SATD: ocg_enc_metrics_kernel<OCG_MET_INTRA_SATD><<<grid, 128, 0, st>>>(src_base, ref_base, ystride, frags, n, out_val, out_dc); break;
    case OCG_MET_SSD: ocg_enc_metrics_kernel<OCG_MET_SSD><<<grid, 128, 0, st>>>(src_base, ref_base, ystride, frags, n, out_val, out_dc); break;
    case OCG_MET_BORDER_SSD: ocg_enc_metrics_kernel<OCG_MET_BORDER_SSD><<<grid, 128, 0, st>>>(src_base, ref_base, ystride, frags, n, out_val, out_dc); break;
    case OCG_MET_SAD_THRESH: ocg_enc_metrics_kernel<OCG_MET_SAD_THRESH><<<grid, 128, 0, st>>>(src_base, ref_base, ystride, frags, n, out_val, out_dc); break;
    default: ocg_enc_metrics_kernel<OCG_MET_INTRA_SAD><<<grid, 128, 0, st>>>(src_base, ref_base, ystride, frags, n, out_val, out_dc); break;
  }
  ocg_count_launch(1);
  return cudaGetLastError() == cudaSuccess ? OCG_OK : OCG_ECUDA;
}

OCG_API int ocg_enc_fdct_quant_batch(const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                                     const ocg_enc_frag *frags, int n, const uint16_t *dequant,
                                     const int16_t *enquant, int16_t *dct, int16_t *qdct, int32_t *nonzero,
                                     void *stream) {
  if (src_base == nullptr || frags == nullptr || dequant == nullptr || enquant == nullptr || dct == nullptr ||
      qdct == nullptr || nonzero == nullptr)
    return OCG_EFAULT;
  if (n < 0) return OCG_EINVAL;
  if (n == 0) return OCG_OK;
  ocg_enc_fdct_quant_kernel<<<(unsigned)((n + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
      src_base, ref_base, ystride, frags, n, dequant, enquant, dct, qdct, nonzero);
  ocg_count_launch(1);
  return cudaGetLastError() == cudaSuccess ? OCG_OK : OCG_ECUDA;
}

OCG_API int ocg_mcenc_search_batch(const uint8_t *src_base, const uint8_t *ref_full_base,
                                   const uint8_t *ref_satd_base, int ystride, const ocg_mb_search_in *in,
                                   ocg_mb_search_out *out, int n, void *stream) {
  if (src_base == nullptr || ref_full_base == nullptr || ref_satd_base == nullptr || in == nullptr || out == nullptr)
    return OCG_EFAULT;
  if (n < 0) return OCG_EINVAL;
  if (n == 0) return OCG_OK;
  ocg_mcenc_search_kernel<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
      src_base, ref_full_base, ref_satd_base, ystride, in, out, n);
  ocg_count_launch(1);
  return cudaGetLastError() == cudaSuccess ? OCG_OK : OCG_ECUDA;
}

} /* extern "C" */
