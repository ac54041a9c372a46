/* TEST INFRASTRUCTURE ONLY (oracle/): thin wrappers that let Python call the
 * reference's *internal* state-taking routines on caller-provided pixels, so
 * unit-level golden vectors can be produced by the real reference code:
 *   oc_state_get_mv_offsets            lib/state.c:846
 *   oc_state_frag_recon_c              lib/state.c:959
 *   oc_state_loop_filter_frag_rows_c   lib/state.c:1055
 *   oc_state_borders_fill_rows/caps    lib/state.c:770,803
 * Each wrapper builds the minimal oc_theora_state those routines read.  This
 * file includes the reference's private headers, so it only builds where
 * /root/reference is present; its objects live in oracle/_ref/.
 */
#include <stdlib.h>
#include <string.h>
#include "state.h"

#define REFH_API __attribute__((visibility("default")))

static oc_theora_state *refh_fake_state(int pixel_fmt) {
  oc_theora_state *st = (oc_theora_state *)calloc(1, sizeof(*st));
  st->info.pixel_fmt = (th_pixel_fmt)pixel_fmt;
  oc_state_accel_init_c(st);
  return st;
}

REFH_API int refh_mv_offsets(int pixel_fmt, int ystride, int pli, int mv, int offs[2]) {
  oc_theora_state *st = refh_fake_state(pixel_fmt);
  int n;
  st->ref_ystride[pli] = ystride;
  offs[0] = offs[1] = 0;
  n = oc_state_get_mv_offsets(st, offs, pli, (oc_mv)mv);
  free(st);
  return n;
}

/* dst/ref point at the frames' bottom-left pixel (offset 0); buf_off locates
   the fragment. */
REFH_API void refh_state_frag_recon(unsigned char *dst_frame, const unsigned char *ref_frame, long buf_off,
                                    int ystride, int pli, int pixel_fmt, int intra, int mv,
                                    ogg_int16_t coeffs[128], int last_zzi, int dc_quant) {
  oc_theora_state *st = refh_fake_state(pixel_fmt);
  oc_fragment frag;
  ptrdiff_t off = buf_off;
  oc_mv fmv = (oc_mv)mv;
  memset(&frag, 0, sizeof(frag));
  frag.coded = 1;
  frag.refi = intra ? OC_FRAME_SELF : OC_FRAME_PREV;
  st->frags = &frag;
  st->frag_buf_offs = &off;
  st->frag_mvs = &fmv;
  st->ref_ystride[pli] = ystride;
  st->ref_frame_data[OC_FRAME_SELF] = dst_frame;
  st->ref_frame_data[OC_FRAME_PREV] = (unsigned char *)ref_frame;
  oc_state_frag_recon_c(st, 0, pli, coeffs, last_zzi, (ogg_uint16_t)dc_quant);
  free(st);
}

/* pix = bottom-left pixel of a plane of nhfrags x nvfrags fragments. */
REFH_API void refh_loop_filter_plane(unsigned char *pix, int ystride, int nhfrags, int nvfrags,
                                     const unsigned char *coded, int limit) {
  oc_theora_state *st = refh_fake_state(TH_PF_444);
  ptrdiff_t n = (ptrdiff_t)nhfrags * nvfrags, i;
  signed char bv[256];
  st->fplanes[0].nhfrags = nhfrags;
  st->fplanes[0].nvfrags = nvfrags;
  st->fplanes[0].froffset = 0;
  st->fplanes[0].nfrags = n;
  st->frags = (oc_fragment *)calloc((size_t)n, sizeof(oc_fragment));
  st->frag_buf_offs = (ptrdiff_t *)calloc((size_t)n, sizeof(ptrdiff_t));
  for (i = 0; i < n; i++) {
    st->frags[i].coded = coded[i] != 0;
    st->frag_buf_offs[i] = (i / nhfrags) * 8 * (ptrdiff_t)ystride + (i % nhfrags) * 8;
  }
  st->ref_ystride[0] = ystride;
  st->ref_frame_data[OC_FRAME_SELF] = pix;
  if (limit) {
    oc_loop_filter_init_c(bv, limit);
    oc_state_loop_filter_frag_rows_c(st, bv, OC_FRAME_SELF, 0, 0, nvfrags);
  }
  free(st->frags);
  free(st->frag_buf_offs);
  free(st);
}

REFH_API void refh_loop_filter_table(signed char bv[256], int limit) { oc_loop_filter_init_c(bv, limit); }

/* pix = TOP-left pixel, positive stride (borders_fill works on the un-flipped
   th_img_plane as well: state.c:787 "allows the stride to be negative"). */
REFH_API void refh_borders_fill(unsigned char *pix_bottom_left, int ystride, int width, int height,
                                int pli, int pixel_fmt) {
  oc_theora_state *st = refh_fake_state(pixel_fmt);
  st->ref_frame_bufs[0][pli].width = width;
  st->ref_frame_bufs[0][pli].height = height;
  st->ref_frame_bufs[0][pli].stride = ystride;
  st->ref_frame_bufs[0][pli].data = pix_bottom_left;
  oc_state_borders_fill_rows(st, 0, pli, 0, height);
  oc_state_borders_fill_caps(st, 0, pli);
  free(st);
}

/* ---------------------------------------------------------------------- */
/* oc_mcenc_search_frame (lib/mcenc.c:268) on caller-provided frames, with the
   minimal oc_enc_ctx it reads: one macro block (index 0) plus `ncn` already
   searched neighbours (indices 1..ncn) that feed the candidate sets. */
#include "encint.h"

REFH_API void refh_mcenc_search_frame(const unsigned char *src, const unsigned char *ref_full,
                                      const unsigned char *ref_satd, int ystride, const long frag_off[4],
                                      int frame, int accum, int ncn, const int *nb_mv, const int *nb_err,
                                      int mv1, int mv2, int own_err, int sp_level, int out[12]) {
  oc_enc_ctx *enc = (oc_enc_ctx *)calloc(1, sizeof(*enc));
  oc_mb_enc_info *embs = (oc_mb_enc_info *)calloc((size_t)ncn + 1, sizeof(*embs));
  oc_mb_map map;
  ptrdiff_t offs[4];
  int i;
  memset(map, 0xFF, sizeof(map));
  for (i = 0; i < 4; i++) { map[0][i] = i; offs[i] = frag_off[i]; }
  enc->mb_info = embs;
  enc->sp_level = sp_level;
  enc->state.mb_maps = &map;
  enc->state.frag_buf_offs = offs;
  enc->state.ref_ystride[0] = ystride;
  enc->state.ref_frame_data[OC_FRAME_IO] = (unsigned char *)src;
  enc->state.ref_frame_data[frame] = (unsigned char *)ref_satd;
  enc->state.ref_frame_data[frame == OC_FRAME_PREV ? OC_FRAME_PREV_ORIG : OC_FRAME_GOLD_ORIG] = (unsigned char *)ref_full;
  embs[0].ncneighbors = (unsigned char)ncn;
  for (i = 0; i < ncn; i++) {
    embs[0].cneighbors[i] = (unsigned)(i + 1);
    embs[i + 1].analysis_mv[0][frame] = (oc_mv)nb_mv[i];
    embs[i + 1].error[frame] = (ogg_uint16_t)nb_err[i];
  }
  embs[0].analysis_mv[1][frame] = (oc_mv)mv1;
  embs[0].analysis_mv[2][frame] = (oc_mv)mv2;
  embs[0].error[frame] = (ogg_uint16_t)own_err;
  oc_mcenc_search_frame(enc, (oc_mv)accum, 0, frame,
                        frame == OC_FRAME_PREV ? OC_FRAME_PREV_ORIG : OC_FRAME_GOLD_ORIG);
  out[0] = embs[0].analysis_mv[0][frame];
  out[1] = embs[0].error[frame];
  out[2] = (int)embs[0].satd[frame];
  for (i = 0; i < 4; i++) {
    out[3 + i] = embs[0].block_mv[i];
    out[7 + i] = (int)embs[0].block_satd[i];
  }
  free(embs);
  free(enc);
}

/* oc_mcenc_refine1mv (lib/mcenc.c:661) and oc_mcenc_refine4mv (lib/mcenc.c:762)
   on caller-provided frames.  mv / block_mv are oc_mv values in half-pel units
   as the full-pel search leaves them (even components).
   out: [0] analysis_mv[0][frame], [1] satd[frame], [2..5] ref_mv[bi],
        [6..9] block_satd[bi]. */
REFH_API void refh_mcenc_refine(const unsigned char *src, const unsigned char *ref, int ystride,
                                const long frag_off[4], int frame, int mv, unsigned satd, const int block_mv[4],
                                const unsigned block_satd[4], int sp_level, int do4mv, int out[10]) {
  oc_enc_ctx *enc = (oc_enc_ctx *)calloc(1, sizeof(*enc));
  oc_mb_enc_info *embs = (oc_mb_enc_info *)calloc(1, sizeof(*embs));
  oc_mb_map map;
  ptrdiff_t offs[4];
  int i;
  memset(map, 0xFF, sizeof(map));
  for (i = 0; i < 4; i++) { map[0][i] = i; offs[i] = frag_off[i]; }
  enc->mb_info = embs;
  enc->sp_level = sp_level;
  enc->state.mb_maps = &map;
  enc->state.frag_buf_offs = offs;
  enc->state.ref_ystride[0] = ystride;
  enc->state.ref_frame_data[OC_FRAME_IO] = (unsigned char *)src;
  enc->state.ref_frame_data[frame] = (unsigned char *)ref;
  embs[0].analysis_mv[0][frame] = (oc_mv)mv;
  embs[0].satd[frame] = satd;
  for (i = 0; i < 4; i++) {
    embs[0].block_mv[i] = (oc_mv)block_mv[i];
    embs[0].block_satd[i] = block_satd[i];
  }
  oc_mcenc_refine1mv(enc, 0, frame);
  if (do4mv) oc_mcenc_refine4mv(enc, 0);
  out[0] = embs[0].analysis_mv[0][frame];
  out[1] = (int)embs[0].satd[frame];
  for (i = 0; i < 4; i++) {
    out[2 + i] = embs[0].ref_mv[i];
    out[6 + i] = (int)embs[0].block_satd[i];
  }
  free(embs);
  free(enc);
}
