/* TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's per-fragment
 * 8x8 block pipeline, used as the parity checker by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg.  The product
 * library never links or calls anything in oracle/.
 *
 * Pinning: every function here is checked bit-for-bit against the compiled
 * reference (oracle/_ref/libth_c.so, built from /root/reference by
 * oracle/Makefile) on random and adversarial inputs by tests/test_oracle_*.py,
 * and against golden vectors generated from that reference and committed under
 * tests/golden/ (generator: tests/golden/make_golden.py).  The reference's own
 * test-suite holds no vectors for this path.
 */
#ifndef THEORA_ORACLE_H
#define THEORA_ORACLE_H
#include <stdint.h>
#include "../include/theora_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- decode-side block kernels ---- */
void oco_idct8x8(int16_t y[64], int16_t x[64], int last_zzi);           /* idct.c:301-330 */
void oco_frag_recon_intra(uint8_t *dst, int ystride, const int16_t res[64]);          /* fragment.c:49 */
void oco_frag_recon_inter(uint8_t *dst, const uint8_t *src, int ystride, const int16_t res[64]); /* :59 */
void oco_frag_recon_inter2(uint8_t *dst, const uint8_t *s1, const uint8_t *s2, int ystride,
                           const int16_t res[64]);                                     /* :70 */
void oco_frag_copy(uint8_t *dst, const uint8_t *src, int ystride);                     /* :20 */
int  oco_mv_offsets(int offs[2], int ystride, int pli, int pixel_fmt, int16_t mv);     /* state.c:846 */
/* state.c:959-1000 with explicit frame pointers instead of oc_theora_state. */
void oco_state_frag_recon(uint8_t *dst_frame, const uint8_t *ref_frame, int32_t buf_off, int ystride,
                          int pli, int pixel_fmt, int intra, int16_t mv, int16_t coeffs[128],
                          int last_zzi, uint16_t dc_quant);
int  oco_lflim(int r, int limit);                                       /* state.c:1036-1045 closed form */
void oco_loop_filter_init(signed char bv[256], int limit);              /* state.c:1036 */
/* Normative sequential order, whole plane (state.c:1055-1105). `pix` addresses
   the bottom-left pixel; ystride is negative. */
void oco_loop_filter_plane_seq(uint8_t *pix, int ystride, int nhfrags, int nvfrags,
                               const uint8_t *coded, int limit);
/* The order-free decomposition the GPU kernel uses (shifted 8x8 cells around
   fragment corners); must equal the sequential form. */
void oco_loop_filter_plane_cells(uint8_t *pix, int ystride, int nhfrags, int nvfrags,
                                 const uint8_t *coded, int limit);
void oco_borders_fill_plane(uint8_t *pix, int ystride, int width, int height, int hpad, int vpad); /* state.c:770-835 */

/* ---- geometry + whole-frame executor over the C-ABI frame description ---- */
int  oco_geometry_init(ocg_geometry *g, int fw, int fh, int pixel_fmt, int nrefs);    /* state.c:424-671 */
void oco_geometry_frag_buf_offs(const ocg_geometry *g, int32_t *offs);
/* `frames` = nrefs*ref_frame_sz bytes laid out like ref_frame_handle. Runs
   recon -> copy -> loop filter -> border fill like decode.c:2858-2945. */
void oco_dec_frame(const ocg_geometry *g, uint8_t *frames, const ocg_dec_frame *f, int stage_mask);

/* ---- encode-side block kernels ---- */
void     oco_fdct8x8(int16_t y[64], const int16_t x[64]);                              /* fdct.c:128 */
void     oco_enquant_init(int16_t enq[128], const uint16_t dequant[64]);               /* enquant.c:184-208 */
int      oco_quantize(int16_t q[64], const int16_t dct[64], const uint16_t dequant[64], const int16_t enq[128]); /* :220 */
void     oco_frag_sub(int16_t d[64], const uint8_t *src, const uint8_t *ref, int ystride);   /* encfrag.c:21 */
void     oco_frag_sub_128(int16_t d[64], const uint8_t *src, int ystride);                    /* :32 */
unsigned oco_frag_sad(const uint8_t *src, const uint8_t *ref, int ystride);                   /* :42 */
unsigned oco_frag_sad_thresh(const uint8_t *src, const uint8_t *ref, int ystride, unsigned thresh); /* :56 */
unsigned oco_frag_sad2_thresh(const uint8_t *src, const uint8_t *r1, const uint8_t *r2, int ystride, unsigned thresh); /* :71 */
unsigned oco_frag_intra_sad(const uint8_t *src, int ystride);                                 /* :88 */
unsigned oco_frag_satd(int *dc, const uint8_t *src, const uint8_t *ref, int ystride);         /* :306 */
unsigned oco_frag_satd2(int *dc, const uint8_t *src, const uint8_t *r1, const uint8_t *r2, int ystride); /* :313 */
unsigned oco_frag_intra_satd(int *dc, const uint8_t *src, int ystride);                       /* :322 */
unsigned oco_frag_ssd(const uint8_t *src, const uint8_t *ref, int ystride);                   /* :338 */
unsigned oco_frag_border_ssd(const uint8_t *src, const uint8_t *ref, int ystride, int64_t mask); /* :352 */
void     oco_frag_copy2(uint8_t *dst, const uint8_t *s1, const uint8_t *s2, int ystride);     /* :368 */

void     oco_dc_unpredict_plane(int16_t *dc, const uint8_t *refs, int nhfrags, int nvfrags);   /* decode.c:1392-1500 */
unsigned oco_block_activity(const uint8_t *src, int ystride, int *sum);                     /* analyze.c:1167-1234 */

/* Batch forms matching the C-ABI encoder entry points. */
void oco_enc_metrics_batch(int metric, const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                           const ocg_enc_frag *frags, int n, uint32_t *out_val, int32_t *out_dc);
void oco_enc_fdct_quant_batch(const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                              const ocg_enc_frag *frags, int n, const uint16_t *dequant,
                              const int16_t *enquant, int16_t *dct, int16_t *qdct, int32_t *nonzero);

void oco_mcenc_search_batch(const uint8_t *src_base, const uint8_t *ref_full_base, const uint8_t *ref_satd_base,
                            int ystride, const ocg_mb_search_in *in, ocg_mb_search_out *out, int n); /* mcenc.c:268-515 */
void oco_mcenc_refine_batch(const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                            const ocg_mb_refine_in *in, ocg_mb_refine_out *out, int n, int flags); /* mcenc.c:606-791 */

/* whole-frame motion analysis, state in/out in mb[] (mcenc.c:90-164, 517-548; analyze.c:2469-2489) */
void oco_me_frame(const uint8_t *src, const uint8_t *ref_full_gold, const uint8_t *ref_full_prev,
                  const uint8_t *ref_satd_gold, const uint8_t *ref_satd_prev, int ystride,
                  const ocg_me_topo *topo, ocg_me_mb *mb, int nmbs, int flags, const uint8_t *gold_refine);

/* ---- out-of-loop post-processing (decode.c:1609-1957), whole plane, internal orientation (row 0 = the row
   the reference processes first) ---- */
void oco_pp_deblock_plane(uint8_t *dst, int dstride, const uint8_t *src, int sstride, int W, int H,
                          const uint8_t *dc_qis, const int *dc_scale, int32_t *variances);   /* decode.c:1700 */
void oco_pp_dering_plane(uint8_t *img, int stride, int W, int H, int pli, int strong_level, const uint8_t *qis,
                         const int *dc_scale, const int *sharp_mod, const int32_t *variances); /* decode.c:1892 */

#ifdef __cplusplus
}
#endif
#endif
