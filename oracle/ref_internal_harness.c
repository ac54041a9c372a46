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

/* ---------------------------------------------------------------------- */
/* Whole-frame motion analysis by the REAL reference: a th_enc_ctx made by
   th_encode_alloc (its own mb_info/cneighbors, mb_maps, frag_buf_offs), whose
   six reference buffers are loaded with caller-provided frames, then
   oc_mcenc_search (lib/mcenc.c:517) for every macro block in coding order and
   the refinements in the order oc_enc_analyze_inter runs them per macro block
   (lib/analyze.c:2402, 2469-2489): [refine4mv] [refine1mv GOLD if flagged]
   [refine1mv PREV].  State (analysis_mv history, error) persists in the
   context across calls, as in an encoder. */
typedef struct refh_me_mb {      /* same layout as ocg_me_mb (include/theora_b200.h) */
  ogg_int16_t  analysis_mv[3][2];
  ogg_uint16_t error[2];
  ogg_uint32_t satd[2];
  ogg_int16_t  unref_mv[2];
  ogg_uint32_t unref_satd[2];
  ogg_int16_t  block_mv[4];
  ogg_int16_t  ref_mv[4];
  ogg_uint32_t block_satd[4];
  ogg_uint32_t ref_block_satd[4];
  unsigned char pad[12];
} refh_me_mb;

typedef struct refh_me_topo {    /* same layout as ocg_me_topo */
  ogg_int32_t frag_off[4];
  ogg_int32_t cn[4];
  unsigned char ncn, valid, pad[6];
} refh_me_topo;

REFH_API void *refh_me_open(int fw, int fh, int pixel_fmt) {
  th_info ti;
  th_enc_ctx *enc;
  th_info_init(&ti);
  ti.frame_width = ti.pic_width = (ogg_uint32_t)fw;
  ti.frame_height = ti.pic_height = (ogg_uint32_t)fh;
  ti.fps_numerator = 30; ti.fps_denominator = 1;
  ti.pixel_fmt = (th_pixel_fmt)pixel_fmt;
  ti.quality = 32;
  ti.keyframe_granule_shift = 6;
  enc = th_encode_alloc(&ti);
  return enc;
}
REFH_API void refh_me_close(void *h) { th_encode_free((th_enc_ctx *)h); }
REFH_API int refh_me_nmbs(void *h) { return (int)((oc_enc_ctx *)h)->state.nmbs; }
REFH_API long refh_me_frame_size(void *h) {
  oc_enc_ctx *enc = (oc_enc_ctx *)h;
  return (long)(enc->state.ref_frame_bufs[1][0].data - enc->state.ref_frame_bufs[0][0].data);
}

REFH_API void refh_me_topology(void *h, refh_me_topo *topo) {
  oc_enc_ctx *enc = (oc_enc_ctx *)h;
  unsigned mbi;
  int i;
  memset(topo, 0, sizeof(*topo) * enc->state.nmbs);
  for (mbi = 0; mbi < enc->state.nmbs; mbi++) {
    if (enc->state.mb_modes[mbi] == OC_MODE_INVALID) continue;
    topo[mbi].valid = 1;
    for (i = 0; i < 4; i++) topo[mbi].frag_off[i] = (ogg_int32_t)enc->state.frag_buf_offs[enc->state.mb_maps[mbi][0][i]];
    topo[mbi].ncn = enc->mb_info[mbi].ncneighbors;
    for (i = 0; i < enc->mb_info[mbi].ncneighbors; i++) topo[mbi].cn[i] = (ogg_int32_t)enc->mb_info[mbi].cneighbors[i];
  }
}

/* frames: 5 buffers of refh_me_frame_size bytes in the library's own layout,
   {IO, PREV_ORIG, GOLD_ORIG, PREV, GOLD}. flags as OCG_ME_*. */
REFH_API void refh_me_frame(void *h, const unsigned char *const frames[5], int flags,
                            const unsigned char *gold_refine, refh_me_mb *out) {
  static const int ROLE[5] = {OC_FRAME_IO, OC_FRAME_PREV_ORIG, OC_FRAME_GOLD_ORIG, OC_FRAME_PREV, OC_FRAME_GOLD};
  oc_enc_ctx *enc = (oc_enc_ctx *)h;
  oc_mb_enc_info *embs = enc->mb_info;
  size_t fsz = (size_t)refh_me_frame_size(h);
  unsigned mbi;
  int i;
  for (i = 0; i < 5; i++) {
    memcpy(enc->state.ref_frame_handle + (size_t)i * fsz, frames[i], fsz);
    enc->state.ref_frame_idx[ROLE[i]] = i;
    enc->state.ref_frame_data[ROLE[i]] = enc->state.ref_frame_bufs[i][0].data;
  }
  enc->sp_level = (flags & 4) ? OC_SP_LEVEL_NOSATD : ((flags & 8) ? OC_SP_LEVEL_FAST_ANALYSIS : OC_SP_LEVEL_EARLY_SKIP);
  enc->prevframe_dropped = (flags & 16) != 0;
  for (mbi = 0; mbi < enc->state.nmbs; mbi++) {
    refh_me_mb *o = out + mbi;
    memset(o, 0, sizeof(*o));
    if (enc->state.mb_modes[mbi] == OC_MODE_INVALID) continue;
    oc_mcenc_search(enc, (int)mbi);
    for (i = 0; i < 2; i++) {
      o->unref_mv[i] = embs[mbi].analysis_mv[0][i];
      o->unref_satd[i] = embs[mbi].satd[i];
    }
    for (i = 0; i < 4; i++) {
      o->block_mv[i] = embs[mbi].block_mv[i];
      o->block_satd[i] = embs[mbi].block_satd[i];
    }
    if ((flags & 2) && !(flags & 8)) {
      oc_mcenc_refine4mv(enc, (int)mbi);
      for (i = 0; i < 4; i++) {
        o->ref_mv[i] = embs[mbi].ref_mv[i];
        o->ref_block_satd[i] = embs[mbi].block_satd[i];
      }
    }
    if (gold_refine != NULL && gold_refine[mbi]) oc_mcenc_refine1mv(enc, (int)mbi, OC_FRAME_GOLD);
    if (flags & 1) oc_mcenc_refine1mv(enc, (int)mbi, OC_FRAME_PREV);
  }
  /* the state the next frame starts from */
  for (mbi = 0; mbi < enc->state.nmbs; mbi++) {
    refh_me_mb *o = out + mbi;
    if (enc->state.mb_modes[mbi] == OC_MODE_INVALID) continue;
    memcpy(o->analysis_mv, embs[mbi].analysis_mv, sizeof(o->analysis_mv));
    for (i = 0; i < 2; i++) {
      o->error[i] = embs[mbi].error[i];
      o->satd[i] = embs[mbi].satd[i];
    }
  }
}

/* The post-processing inputs the reference decoder holds after a th_decode_packetin (decode.c:1204-1243,
   397-409): per fragment the tracked DC quantiser index and state.qis[frag.qii], the two tables, the
   variances its own filters accumulated for the frame, and the level the frame was processed at.  Returns
   the fragment count, or -1 while the decoder is not tracking (pp level 0 / no key frame seen yet). */
#include "decint.h"
REFH_API long refh_dec_pp_state(th_dec_ctx *dec_, unsigned char *dc_qis, unsigned char *qis, int *dc_scale, int *sharp_mod,
                                int *variances, int *level) {
  oc_dec_ctx *dec = (oc_dec_ctx *)dec_;
  ptrdiff_t i, n;
  if (dec == NULL || dec->dc_qis == NULL) return -1;
  n = dec->state.nfrags;
  for (i = 0; i < n; i++) {
    dc_qis[i] = dec->dc_qis[i];
    qis[i] = (unsigned char)dec->state.qis[dec->state.frags[i].qii];
    if (variances != NULL) variances[i] = dec->variances != NULL ? dec->variances[i] : 0;
  }
  for (i = 0; i < 64; i++) { dc_scale[i] = dec->pp_dc_scale[i]; sharp_mod[i] = dec->pp_sharp_mod[i]; }
  if (level != NULL) *level = dec->pipe.pp_level;
  return (long)n;
}
