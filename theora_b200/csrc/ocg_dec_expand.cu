/* Token -> coefficient expansion on the device.
 *
 * After the reference's entropy decoder has unpacked a packet (decode.c:1000-1190) a frame is held as
 *   frags[]      one packed word per fragment: coded flag, qii, reference type, mode, DC (state.h:297-322)
 *   frag_mvs[]   one vector per fragment
 *   dct_tokens[] 3 x 64 token lists, one per (plane, zig-zag index), in the decoder's INTERNAL token
 *                alphabet: one byte per token + one byte of extra bits for tokens 0..14 + a second one for
 *                token 0 (decode.c:95-300), with the start of each list in ti0[][] and the EOB run that
 *                reaches into each list in eob_runs[][]
 * and the host then walks every coded fragment through the lists serially (decode.c:1511-1586) to build its
 * 64 coefficients.  Here the host uploads those three arrays as they are and the walk runs on the device:
 *
 *   ocg_tok_parse_kernel   one CTA per list: token boundaries (a token's length is a function of its first
 *                          byte; a chunk of bytes is therefore a map {entry offset 0..2} -> {exit offset,
 *                          token count} and the maps compose associatively: block scan), decoded token
 *                          words and the exclusive prefix of "fragments served" (EOB runs serve many).
 *   ocg_tok_expand_kernel  one CTA per plane, 64 steps (one per zig-zag index, the only serial dimension):
 *                          the fragments that need a token at this index are ranked in coded order by a
 *                          block scan, each finds its token by rank in the list's prefix array, stores the
 *                          dequantised coefficient (decode.c:1573) in the fragment's dense 8x8 block and
 *                          moves on to index + run + 1, or ends (EOB) and records last_zzi.
 *   ocg_rec_build_kernel   the 16-byte records the reconstruction kernels read.
 *
 * The dense coefficient blocks are all-zero between frames: the transform pass clears what it reads, as the
 * reference's own iDCT does (idct.c:245,276,295).
 */
#include <cstring>
#include <vector>
#include "ocg_internal.h"

namespace {

/* ---- the decoder's internal token alphabet (decode.c:95-300, huffdec.c) -------------------------------
   92 tokens; what each stands for, by group, in the order the alphabet lists them.  Generated, not
   transcribed: tests/test_gpu_expand.py checks every stream against the reference's own expansion. */
struct TokInfo { uint8_t rlen, eob, neg, ebkind; uint16_t mag; }; /* ebkind: 0 none, 1 adds to eob, 2 to mag, 3 to rlen */

void build_token_table(uint32_t packed[92]) {
  std::vector<TokInfo> t;
  auto add = [&](int eob, int rlen, int mag, int neg, int ebkind) {
    TokInfo i;
    i.rlen = (uint8_t)rlen; i.eob = (uint8_t)eob; i.neg = (uint8_t)neg; i.ebkind = (uint8_t)ebkind; i.mag = (uint16_t)mag;
    t.push_back(i);
  };
  add(0, 0, 0, 0, 1);                                            /* EOB run, 12 extra bits (0 = to the end of the frame) */
  add(16, 0, 0, 0, 1);                                           /* EOB run 16..31 */
  for (int base : {13, 21, 37}) { add(0, 0, base, 0, 2); add(0, 0, base, 1, 2); } /* large values, extra bits add to |v| */
  add(0, 0, 69, 0, 2); add(0, 0, 325, 0, 2); add(0, 0, 69, 1, 2); add(0, 0, 325, 1, 2);
  add(0, 10, 1, 0, 3); add(0, 10, 1, 1, 3);                      /* 10..17 zeros, then +-1 */
  add(0, 0, 0, 0, 3);                                            /* zero run, 6 extra bits */
  for (int e = 1; e <= 3; e++) add(e, 0, 0, 0, 0);               /* EOB runs 1..3 */
  for (int r = 1; r <= 5; r++) { add(0, r, 1, 0, 0); add(0, r, 1, 1, 0); } /* r zeros, then +-1 */
  add(0, 1, 2, 0, 0); add(0, 1, 3, 0, 0); add(0, 1, 2, 1, 0); add(0, 1, 3, 1, 0); /* one zero, then +-2/3 */
  for (int neg = 0; neg < 2; neg++) for (int r = 6; r <= 9; r++) add(0, r, 1, neg, 0);
  for (int neg = 0; neg < 2; neg++) for (int m = 2; m <= 3; m++) for (int r = 2; r <= 3; r++) add(0, r, m, neg, 0);
  for (int r = 0; r < 8; r++) add(0, r, 0, 0, 0);                /* short zero runs */
  add(0, 0, 1, 0, 0); add(0, 0, 1, 1, 0); add(0, 0, 2, 0, 0); add(0, 0, 2, 1, 0);
  for (int m = 3; m <= 6; m++) { add(0, 0, m, 0, 0); add(0, 0, m, 1, 0); }
  add(0, 0, 7, 0, 0); add(0, 0, 8, 0, 0); add(0, 0, 7, 1, 0); add(0, 0, 8, 1, 0);
  for (int neg = 0; neg < 2; neg++) for (int m = 9; m <= 12; m++) add(0, 0, m, neg, 0);
  for (int e = 8; e <= 15; e++) add(e, 0, 0, 0, 0);
  for (int e = 4; e <= 7; e++) add(e, 0, 0, 0, 0);
  for (size_t i = 0; i < 92; i++) {
    const TokInfo &k = i < t.size() ? t[i] : t[0];
    packed[i] = (uint32_t)k.rlen | (uint32_t)k.eob << 8 | (uint32_t)k.mag << 13 | (uint32_t)k.neg << 23 | (uint32_t)k.ebkind << 24;
  }
}

__constant__ uint32_t c_tokinfo[92];
__constant__ uint8_t c_zigzag[64]; /* zig-zag index -> natural (row-major) position */

/* decoded token word: bit 31 = EOB token (bits 0..30: fragments served, saturated); otherwise bits 0..15 the
   coefficient (two's complement), bits 16..22 the zero run before it */
#define TOK_EOB 0x80000000u
#define COV_SAT 0x40000000u

__device__ __forceinline__ int tok_len(int b0) { return 1 + (b0 < 15) + (b0 == 0); }

__device__ __forceinline__ uint32_t tok_decode(const uint8_t *p) {
  const int b0 = p[0] < 92 ? p[0] : 15; /* never produced by the decoder; treated as an EOB */
  const uint32_t k = c_tokinfo[b0];
  int eb = 0;
  if (b0 < 15) eb = p[1];
  if (b0 == 0) eb |= (int)p[2] << 8;
  const int kind = (int)(k >> 24) & 3;
  int rlen = (int)(k & 0xFFu), eob = (int)(k >> 8) & 31, mag = (int)(k >> 13) & 1023;
  if (kind == 1) eob += eb;
  else if (kind == 2) mag += eb;
  else if (kind == 3) rlen += eb;
  if (b0 == 0 && eb == 0) return TOK_EOB | COV_SAT; /* "no more coded coefficients in the frame" */
  if (eob > 0) return TOK_EOB | (uint32_t)eob;
  const int v = (k >> 23) & 1 ? -mag : mag;
  return ((uint32_t)v & 0xFFFFu) | (uint32_t)(rlen & 127) << 16;
}

__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b) { const uint32_t s = a + b; return s > COV_SAT ? COV_SAT : s; }

/* ---- K1: one CTA per (zig-zag index, plane) list ------------------------------------------------- */
#define TP_THREADS 256
#define TP_CHUNK 32

/* A chunk of token bytes as a map of the offset at which its first token starts (0..2: a token is at most
   3 bytes long): exit offset into the next chunk (2 bits each, bits 0..5) and tokens started (16 bits each,
   from bit 8).  Maps compose associatively, so token boundaries come out of a scan. */
typedef unsigned long long cmap_t;
__device__ __forceinline__ int cm_exit(cmap_t m, int s) { return (int)(m >> (2 * s)) & 3; }
__device__ __forceinline__ unsigned cm_cnt(cmap_t m, int s) { return (unsigned)(m >> (8 + 16 * s)) & 0xFFFFu; }
__device__ __forceinline__ cmap_t cm_make(const int ex[3], const unsigned cn[3]) {
  return (cmap_t)(ex[0] | ex[1] << 2 | ex[2] << 4) | (cmap_t)cn[0] << 8 | (cmap_t)cn[1] << 24 | (cmap_t)cn[2] << 40;
}
__device__ __forceinline__ cmap_t cm_compose(cmap_t a, cmap_t b) { /* a first, then b */
  int ex[3];
  unsigned cn[3];
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const int mid = cm_exit(a, s);
    ex[s] = cm_exit(b, mid);
    cn[s] = cm_cnt(a, s) + cm_cnt(b, mid);
  }
  return cm_make(ex, cn);
}
#define CM_IDENTITY ((cmap_t)(0 | 1 << 2 | 2 << 4))

__global__ void __launch_bounds__(TP_THREADS)
ocg_tok_parse_kernel(const OcgExpandDev *__restrict__ X, const uint8_t *__restrict__ bytes, uint32_t *__restrict__ tok,
                     uint32_t *__restrict__ cov, int32_t *__restrict__ ntok_out) {
  const int list = (int)blockIdx.x;           /* storage order of the lists: index-major, plane-minor */
  const int z = list / 3, p = list - 3 * z;
  const int b0 = X->ti0[p][z];
  const int b1 = list == 191 ? X->ntoken_bytes : X->ti0[(p + 1) % 3][z + (p == 2)];
  const int t = (int)threadIdx.x, lane = t & 31, warp = t >> 5;
  __shared__ cmap_t s_wmap[TP_THREADS / 32];
  __shared__ uint32_t s_wcov[TP_THREADS / 32];
  __shared__ int s_carry[3];                /* entry offset, tokens so far, coverage so far */
  if (t == 0) { s_carry[0] = 0; s_carry[1] = 0; s_carry[2] = 0; }
  __syncthreads();
  for (int tile = b0; tile < b1; tile += TP_THREADS * TP_CHUNK) {
    const int c0 = tile + t * TP_CHUNK;
    const int clen = min(TP_CHUNK, max(0, b1 - c0));
    uint8_t loc[TP_CHUNK + 2];
#pragma unroll
    for (int i = 0; i < TP_CHUNK + 2; i++) loc[i] = (uint8_t)15;
    if (clen > 0) {
      for (int i = 0; i < TP_CHUNK + 2 && c0 + i < b1; i++) loc[i] = bytes[c0 + i];
    }
    /* this chunk as a map of the entry offset */
    cmap_t mine = CM_IDENTITY;
    if (clen > 0) {
      int ex[3];
      unsigned cn[3];
#pragma unroll
      for (int s = 0; s < 3; s++) {
        int pos = s;
        unsigned n = 0;
        while (pos < clen) { pos += tok_len(loc[pos]); n++; }
        ex[s] = clen == TP_CHUNK ? pos - TP_CHUNK : 0; /* the exit of a list's last chunk is never used */
        cn[s] = n;
      }
      mine = cm_make(ex, cn);
    }
    /* inclusive scan of the composition: inside the warp by shuffles, across the 8 warps serially */
    cmap_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const cmap_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if (lane >= d) inc = cm_compose(o, inc);
    }
    if (lane == 31) s_wmap[warp] = inc;
    __syncthreads();
    const int entry0 = s_carry[0], tok0 = s_carry[1];
    const uint32_t cov0 = (uint32_t)s_carry[2];
    cmap_t before = CM_IDENTITY; /* everything in this tile before my chunk */
    for (int w = 0; w < warp; w++) before = cm_compose(before, s_wmap[w]);
    {
      const cmap_t prev = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
      if (lane > 0) before = cm_compose(before, prev);
    }
    const int my_entry = cm_exit(before, entry0);
    const int my_tok = tok0 + (int)cm_cnt(before, entry0);
    /* fragments served by my chunk's tokens */
    uint32_t mycov = 0;
    {
      int pos = my_entry;
      while (pos < clen) {
        const uint32_t w = tok_decode(loc + pos);
        mycov = sat_add(mycov, (w & TOK_EOB) ? (w & 0x7FFFFFFFu) : 1u);
        pos += tok_len(loc[pos]);
      }
    }
    uint32_t cinc = mycov;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, cinc, d);
      if (lane >= d) cinc = sat_add(o, cinc);
    }
    if (lane == 31) s_wcov[warp] = cinc;
    __syncthreads();
    uint32_t run = cov0;
    for (int w = 0; w < warp; w++) run = sat_add(run, s_wcov[w]);
    {
      const uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, cinc, 1); /* lanes before me in this warp */
      if (lane > 0) run = sat_add(run, excl);
    }
    {
      int pos = my_entry, j = my_tok;
      while (pos < clen) {
        const uint32_t w = tok_decode(loc + pos);
        tok[b0 + j] = w;
        cov[b0 + j] = run;
        run = sat_add(run, (w & TOK_EOB) ? (w & 0x7FFFFFFFu) : 1u);
        pos += tok_len(loc[pos]);
        j++;
      }
    }
    __syncthreads();
    if (t == TP_THREADS - 1) {
      cmap_t all = before;
      all = cm_compose(all, mine);
      uint32_t tot = cov0;
      for (int w = 0; w < TP_THREADS / 32; w++) tot = sat_add(tot, s_wcov[w]);
      s_carry[0] = cm_exit(all, entry0);
      s_carry[1] = tok0 + (int)cm_cnt(all, entry0);
      s_carry[2] = (int)tot;
    }
    __syncthreads();
  }
  if (t == 0) ntok_out[list] = s_carry[1];
}

/* ---- K2: one CTA per plane, 64 steps -------------------------------------------------------------- */
#define TX_THREADS 1024

__global__ void __launch_bounds__(TX_THREADS)
ocg_tok_expand_kernel(const OcgGeomDev g, const OcgExpandDev *__restrict__ X, const int32_t *__restrict__ order,
                      const uint32_t *__restrict__ words, const uint32_t *__restrict__ tok, const uint32_t *__restrict__ cov,
                      const int32_t *__restrict__ ntok, const uint16_t *__restrict__ dequant, int16_t *__restrict__ coef,
                      uint8_t *nextz_global, size_t nextz_stride, uint8_t *__restrict__ lastz, uint8_t *__restrict__ rmask,
                      int nextz_in_smem) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  const int p = (int)blockIdx.x;
  const OcgPlaneDev &P = g.p[p];
  const int n = P.nhfrags * P.nvfrags;
  const int t = (int)threadIdx.x, lane = t & 31, warp = t >> 5;
  /* a thread owns `per` consecutive positions of the plane's coded order (a multiple of 4: the step state
     of four positions is one word) */
  const int per = (((n + TX_THREADS - 1) / TX_THREADS) + 3) & ~3;
  const int l0 = min(n, t * per), l1 = min(n, (t + 1) * per);
  uint8_t *nextz = nextz_in_smem ? s_dyn : nextz_global + (size_t)p * nextz_stride; /* padded to whole words per thread */
  const int32_t *ord = order + P.froffset;
  __shared__ int s_warp[32];
  __shared__ int s_total;
  __shared__ int s_cntz[64 + 1]; /* positions waiting at each index */
  if (t < 65) s_cntz[t] = 0;
  __syncthreads();
  {
    int mine = 0;
    for (int i = t * per; i < (t + 1) * per; i++) {
      uint8_t v = 255;
      if (i < n) {
        const int f = ord[i];
        if (words[f] & 1u) { v = 0; mine++; }
        lastz[f] = 0;
        rmask[f] = 0;
      }
      nextz[i] = v;
    }
    if (mine) atomicAdd(&s_cntz[0], mine);
  }
  __syncthreads();
  for (int z = 0; z < 64; z++) {
    if (s_cntz[z] == 0) continue; /* block-uniform: written before the previous step's barrier */
    int cnt = 0;
    {
      const uint32_t zz = 0x01010101u * (uint32_t)z;
      const uint32_t *wz = (const uint32_t *)(nextz + t * per);
      for (int k = 0; k < per / 4; k++) cnt += __popc(__vcmpeq4(wz[k], zz)) >> 3;
    }
    /* exclusive block scan of cnt */
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int v = s_warp[lane], w2 = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, w2, d);
        if (lane >= d) w2 += o;
      }
      s_warp[lane] = w2 - v;
      if (lane == 31) s_total = w2;
    }
    __syncthreads();
    int r = s_warp[warp] + inc - cnt;
    if (cnt != 0) {
      const int list = z * 3 + p;
      const int b0 = X->ti0[p][z];
      const int nt = ntok[list];
      const uint32_t eob0 = (uint32_t)min(X->eob_runs[p][z], (int)COV_SAT);
      const uint32_t *ltok = tok + b0, *lcov = cov + b0;
      int j = -1;
      for (int i = l0; i < l1; i++) {
        if (nextz[i] != z) continue;
        const int f = ord[i];
        int nz = 255;
        if ((uint32_t)r >= eob0) {
          const uint32_t rr = (uint32_t)r - eob0;
          if (j < 0) { /* largest j with cov[j] <= rr */
            int lo = 0, hi = nt;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (lcov[mid] <= rr) lo = mid + 1; else hi = mid; }
            j = lo - 1;
          } else {
            while (j + 1 < nt && lcov[j + 1] <= rr) j++;
          }
          if (j >= 0 && j < nt) {
            const uint32_t w = ltok[j];
            if (!(w & TOK_EOB)) {
              const int pos2 = z + (int)((w >> 16) & 127u);
              if (pos2 > 0 && pos2 < 64) {
                const uint32_t fw = words[f];
                const int qii = (int)(fw >> 2) & 15, qti = ((fw >> 8) & 7u) != 1u; /* mb_mode != OC_MODE_INTRA */
                const int qi = X->qis[qii < X->nqis ? qii : 0];
                const int q = dequant[(((size_t)qi * 3 + p) * 2 + qti) * 64 + pos2];
                const int v = (int)(int16_t)(w & 0xFFFFu) * q; /* decode.c:1573 */
                if ((int16_t)v != 0) {
                  const int nat = c_zigzag[pos2];
                  coef[(size_t)f * 64 + nat] = (int16_t)v;
                  rmask[f] |= (uint8_t)(1u << (nat >> 3));
                }
              }
              if (pos2 + 1 < 64) nz = pos2 + 1;
            }
          }
        }
        nextz[i] = (uint8_t)nz;
        if (nz == 255) lastz[f] = (uint8_t)z;
        else atomicAdd(&s_cntz[nz], 1);
        r++;
      }
    }
    __syncthreads(); /* s_cntz of later steps is complete; s_warp / s_total may be rewritten */
  }
}

/* ---- K3: records ---------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(256)
ocg_rec_build_kernel(const OcgGeomDev g, const uint32_t *__restrict__ words, const int16_t *__restrict__ mvs,
                     const int32_t *__restrict__ buf_off, const uint8_t *__restrict__ lastz, const uint8_t *__restrict__ rmask,
                     int16_t *__restrict__ coef, const int16_t *__restrict__ dc_final, const OcgJobDev *__restrict__ job,
                     ocg_frag_rec *__restrict__ recs) {
  const int intra_frame = job->intra_frame;
  const int f = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (f >= g.nfrags) return;
  const uint32_t w = words[f];
  const int pli = f >= g.p[2].froffset ? 2 : (f >= g.p[1].froffset ? 1 : 0);
  ocg_frag_rec r;
  r.buf_off = buf_off[f];
  r.coeff_row = (uint32_t)f * 8u;
  if (w & 1u) {
    const int qti = ((w >> 8) & 7u) != 1u;
    r.mv = intra_frame ? (int16_t)0 : mvs[f];
    r.dc = dc_final != nullptr ? dc_final[f] : (int16_t)(w >> 16);
    r.rowmask = rmask[f];
    r.last_zzi = lastz[f];
    if (r.last_zzi < 2 && r.rowmask != 0) {
      /* no transform will read (and clear) this block: a run that overshoots the block's end left
         coefficients behind (malformed stream); keep the all-zero invariant */
      for (int row = 0; row < 8; row++)
        if (r.rowmask >> row & 1) ((uint4 *)(coef + (size_t)f * 64))[row] = make_uint4(0, 0, 0, 0);
      r.rowmask = 0;
    }
    r.refi = (uint8_t)((w >> 6) & 3u);
    r.pli_qti = (uint8_t)(pli | qti << 2);
  } else {
    r.mv = 0; r.dc = 0; r.rowmask = 0; r.last_zzi = 0;
    r.refi = OCG_FRAG_UNCODED;
    r.pli_qti = (uint8_t)pli;
  }
  recs[f] = r;
}

/* stage-in for the token path: words, vectors, token bytes, the two headers */
__device__ __forceinline__ uint4 ldh16(const void *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(256)
ocg_stage_tokens_kernel(const OcgJobDev *__restrict__ h_job, OcgJobDev *__restrict__ d_job, const OcgExpandDev *__restrict__ h_x,
                        OcgExpandDev *__restrict__ d_x, const uint8_t *__restrict__ h_words, uint8_t *__restrict__ d_words,
                        int words_bytes, const uint8_t *__restrict__ h_mvs, uint8_t *__restrict__ d_mvs, int mvs_bytes,
                        const uint8_t *__restrict__ h_tok, uint8_t *__restrict__ d_tok) {
  /* the host arrays are only guaranteed 4-/2-/1-byte aligned: copy whole 16-byte lines around them (the
     device buffers mirror the host's alignment within a line) */
  const int tok_bytes = h_x->ntoken_bytes;
  const int stride = (int)(gridDim.x * blockDim.x);
  const int tid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const uint8_t *src[3] = {h_words, h_mvs, h_tok};
  uint8_t *dst[3] = {d_words, d_mvs, d_tok};
  const int len[3] = {words_bytes, mvs_bytes, tok_bytes};
  for (int a = 0; a < 3; a++) {
    if (len[a] <= 0) continue;
    const uintptr_t s0 = (uintptr_t)src[a] & ~(uintptr_t)15, s1 = ((uintptr_t)src[a] + (size_t)len[a] + 15) & ~(uintptr_t)15;
    const int nlines = (int)((s1 - s0) >> 4);
    uint8_t *d0 = dst[a] - ((uintptr_t)src[a] - s0); /* dst was chosen with the same offset inside a line */
    for (int i = tid; i < nlines; i += stride) ((uint4 *)d0)[i] = ldh16((const uint4 *)s0 + i);
  }
  if (blockIdx.x == 0) {
    for (int i = (int)threadIdx.x; i < (int)(sizeof(OcgJobDev) / 8); i += (int)blockDim.x) ((uint64_t *)d_job)[i] = ((const uint64_t *)h_job)[i];
    for (int i = (int)threadIdx.x; i < (int)(sizeof(OcgExpandDev) / 8); i += (int)blockDim.x) ((uint64_t *)d_x)[i] = ((const uint64_t *)h_x)[i];
  }
}

} /* namespace */

/* ---- launch helpers (ocg_internal.h) -------------------------------------------------------------- */
void ocg_expand_init_tables(cudaStream_t st) {
  static bool done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 16 && done[dev]) return;
  uint32_t packed[92];
  build_token_table(packed);
  /* zig-zag scan (spec 2.3 / Figure 2.8): anti-diagonals, alternating direction */
  uint8_t zz[64];
  int x = 0, y = 0;
  for (int i = 0; i < 64; i++) {
    zz[i] = (uint8_t)(y * 8 + x);
    if (((x + y) & 1) == 0) { /* moving up-right */
      if (x == 7) y++; else if (y == 0) x++; else { x++; y--; }
    } else {                  /* moving down-left */
      if (y == 7) x++; else if (x == 0) y++; else { x--; y++; }
    }
  }
  cudaMemcpyToSymbolAsync(c_tokinfo, packed, sizeof(packed), 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_zigzag, zz, sizeof(zz), 0, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  if (dev >= 0 && dev < 16) done[dev] = true;
}

void ocg_launch_stage_tokens(const OcgJobDev *h_job, OcgJobDev *d_job, const OcgExpandDev *h_x, OcgExpandDev *d_x,
                             const void *h_words, void *d_words, int words_bytes, const void *h_mvs, void *d_mvs, int mvs_bytes,
                             const void *h_tok, void *d_tok, cudaStream_t st) {
  ocg_stage_tokens_kernel<<<48, 256, 0, st>>>(h_job, d_job, h_x, d_x, (const uint8_t *)h_words, (uint8_t *)d_words, words_bytes,
                                              (const uint8_t *)h_mvs, (uint8_t *)d_mvs, mvs_bytes, (const uint8_t *)h_tok,
                                              (uint8_t *)d_tok);
  ocg_count_launch(1);
}

void ocg_launch_expand(const OcgGeomDev &g, const OcgExpandDev *d_x, const OcgExpandBufs &B, const int16_t *dc_final,
                       const OcgJobDev *d_job, ocg_frag_rec *d_recs, cudaStream_t st) {
  ocg_tok_parse_kernel<<<192, TP_THREADS, 0, st>>>(d_x, B.tokens, B.tok, B.cov, B.ntok);
  {
    /* the per-position step state lives in shared memory when the largest plane fits (4K luma: 130 KB) */
    int nmax = 0;
    for (int p = 0; p < 3; p++) nmax = nmax > g.p[p].nhfrags * g.p[p].nvfrags ? nmax : g.p[p].nhfrags * g.p[p].nvfrags;
    const size_t need = (size_t)((((nmax + TX_THREADS - 1) / TX_THREADS) + 3) & ~3) * TX_THREADS;
    const int in_smem = need <= 200 * 1024;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(ocg_tok_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
    ocg_tok_expand_kernel<<<3, TX_THREADS, in_smem ? need : 0, st>>>(g, d_x, B.order, B.words, B.tok, B.cov, B.ntok, B.dequant,
                                                                     B.coef, B.nextz, need, B.lastz, B.rmask, in_smem);
  }
  ocg_rec_build_kernel<<<(unsigned)((g.nfrags + 255) / 256), 256, 0, st>>>(g, B.words, B.mvs, B.buf_off, B.lastz, B.rmask, B.coef,
                                                                         dc_final, d_job, d_recs);
  ocg_count_launch(3);
}
