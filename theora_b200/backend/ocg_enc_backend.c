/* ocg_enc_backend.c -- encoder half of the vtable back-end: INTRA frames of an
 * unmodified th_encode_* run their block pipeline on the B200.
 *
 * Every encoder hook (encint.h:292-325) returns its result synchronously to
 * serial host code (mode decision, R-D tokeniser), so none of them can launch
 * a kernel per call.  For an intra frame, however,
 *   frag_intra_satd(src)                      analyze.c:1385-1534
 *   frag_sub_128(src) -> fdct8x8 -> quantize  analyze.c:725-782
 * are functions of the input frame and the frame's quantiser tables only.
 * The first hook of a frame, enquant_table_fixup (analyze.c:564, after the
 * input has been copied into OC_FRAME_IO and the tables have been condensed),
 * therefore runs ONE batched device pre-pass over all fragments
 * (ocg_enc_intra_prepass) and the per-block hooks become table look-ups keyed
 * by the block's source pointer.  What happens after the tokeniser --
 * idct8x8 + frag_recon_intra (analyze.c:803-822), the loop filter and the
 * border fill -- is never read back by intra analysis (the SSD check at
 * analyze.c:825 is inter-only, _fr!=NULL), so it is RECORDED exactly like the
 * decoder's state_frag_recon and flushed through ocg_dec_submit at the
 * restore_fpu that opens oc_enc_frame_pack (encode.c:911); the reconstructed
 * frame is copied back into the host SELF buffer for whoever predicts from it.
 *
 * Inter frames need their reconstruction inside the analysis loop (skip
 * decision, analyze.c:825-862) and their candidates from already-analysed
 * neighbours (mcenc.c:90-164): that needs a restructured caller and is not
 * served by this back-end.  An encoder that can emit inter frames
 * (keyframe_granule_shift > 0) keeps the reference's C kernels on the host; an
 * intra-only encoder (keyframe_granule_shift == 0) takes the device path and
 * fails to allocate without a device -- there is no silent CPU fallback for it.
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include "encint.h"
#include "ocg_backend.h"

/* the reference's own entry points, renamed on encode.c's command line */
th_enc_ctx *oc_refimpl_encode_alloc(const th_info *_info);
void oc_refimpl_encode_free(th_enc_ctx *_enc);

int ocg_backend_device_(void); /* ocg_backend.c */

typedef struct ocg_enc_backend {
  oc_enc_ctx          *enc;
  ocg_ctx             *ctx;
  ocg_geometry         geom;
  ocg_staging          st;
  ocg_enc_intra_tables tab;
  int                  nqis;
  int                  frame_open;
  int                  ncoded;
  int                  nrows;
  int                  pinned;
  int                  self_on_device; /* buffer index whose reconstruction has not been copied to the host, or -1 */
  /* source/destination pointer -> fragment index */
  ogg_int32_t         *off2frag;
  ptrdiff_t            off_min;
  size_t               noff;
  /* block in flight: sub_128 -> fdct8x8 -> quantize -> [idct8x8] -> recon_intra */
  ptrdiff_t            cur_fragi;
  int                  idct_pending;
  int                  pend_last_zzi;
  ogg_uint32_t         pend_row0;
  int                  pend_mask;
  ogg_int16_t          pend_dc;
  /* quantiser tables in the layout of ocg_enc_fdct_quant_batch */
  ogg_uint16_t         dequant[3][2][3][64];
  ogg_int16_t          enquant[3][2][3][64][2];
  struct ocg_enc_backend *next;
} ocg_enc_backend;

static pthread_mutex_t g_elock = PTHREAD_MUTEX_INITIALIZER;
static ocg_enc_backend *g_elist;
static int g_enc_mode = OCG_ENC_AUTO;
static ocg_enc_spy_fn g_enc_spy;
static void *g_enc_spy_user;
static __thread ocg_enc_backend *t_enc;
static __thread int t_enc_init_failed;
static __thread ocg_enc_backend *t_enc_created; /* made by the th_encode_alloc in progress */
static ocg_enc_backend_stats g_estats;
static pthread_mutex_t g_estats_lock = PTHREAD_MUTEX_INITIALIZER;

OCG_API void ocg_backend_set_enc_mode(int mode) { g_enc_mode = mode; }
OCG_API void ocg_backend_set_enc_spy(ocg_enc_spy_fn fn, void *user) { g_enc_spy = fn; g_enc_spy_user = user; }
OCG_API void ocg_backend_get_enc_stats(ocg_enc_backend_stats *out, int reset) {
  pthread_mutex_lock(&g_estats_lock);
  if (out) *out = g_estats;
  if (reset) memset(&g_estats, 0, sizeof(g_estats));
  pthread_mutex_unlock(&g_estats_lock);
}

static double enc_now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void enc_fatal(const char *what) {
  fprintf(stderr, "theora_b200 encoder back-end: %s (%s)\n", what, ocg_last_error());
  abort();
}

static ocg_enc_backend *enc_backend_of(const oc_enc_ctx *enc) {
  ocg_enc_backend *b = t_enc;
  if (b != NULL && b->enc == enc) return b;
  pthread_mutex_lock(&g_elock);
  for (b = g_elist; b != NULL && b->enc != enc; b = b->next) {}
  pthread_mutex_unlock(&g_elock);
  return b;
}

static inline ptrdiff_t enc_fragi_of(const ocg_enc_backend *b, const unsigned char *p, int frame) {
  const unsigned char *base = b->enc->state.ref_frame_data[frame];
  size_t k = (size_t)((p - base) - b->off_min);
  ogg_int32_t fragi;
  if ((k & 7) != 0 || (k >> 3) >= b->noff || (fragi = b->off2frag[k >> 3]) < 0)
    enc_fatal("block pointer does not address a fragment of the expected frame");
  return fragi;
}

/* ---- frame life cycle ---------------------------------------------------- */
static void enc_begin_frame(ocg_enc_backend *b, int nqis) {
  oc_enc_ctx *enc = b->enc;
  oc_theora_state *st = &enc->state;
  const unsigned char *host_io;
  double t0 = enc_now_s();
  int pli, qii, zzi;
  if (st->frame_type != OC_INTRA_FRAME)
    enc_fatal("inter frame reached the intra-only device encoder (keyframe_granule_shift==0 expected)");
  if (nqis < 1 || nqis > 3) enc_fatal("unexpected quantiser count");
  /* a frame that was analysed but never packed (re-analysis) is simply dropped */
  if (ocg_dec_staging(b->ctx, &b->st) < 0) enc_fatal("ocg_dec_staging failed");
  b->ncoded = b->nrows = 0;
  b->cur_fragi = -1;
  b->idct_pending = 0;
  b->nqis = nqis;
  /* analyze.c:544-564 has just condensed the tables for this frame */
  for (pli = 0; pli < 3; pli++)
    for (qii = 0; qii < nqis; qii++) {
      const oc_iquant *iq = (const oc_iquant *)enc->enquant[pli][qii][0];
      memcpy(b->dequant[pli][0][qii], enc->dequant[pli][qii][0], 64 * sizeof(ogg_uint16_t));
      for (zzi = 0; zzi < 64; zzi++) {
        b->enquant[pli][0][qii][zzi][0] = iq[zzi].m;
        b->enquant[pli][0][qii][zzi][1] = iq[zzi].l;
      }
    }
  host_io = st->ref_frame_handle + (size_t)st->ref_frame_idx[OC_FRAME_IO] * (size_t)b->geom.ref_frame_sz;
  if (ocg_enc_intra_prepass(b->ctx, st->ref_frame_idx[OC_FRAME_IO], host_io, &b->dequant[0][0][0][0],
                            &b->enquant[0][0][0][0][0], nqis, &b->tab) < 0)
    enc_fatal("ocg_enc_intra_prepass failed");
  b->frame_open = 1;
  pthread_mutex_lock(&g_estats_lock);
  g_estats.prepass_frames++;
  g_estats.prepass_seconds += enc_now_s() - t0;
  g_estats.h2d_bytes += (long)b->geom.ref_frame_sz;
  g_estats.d2h_bytes += (long)b->geom.nfrags * (8 + 4 * nqis + 128 + 128 * nqis);
  pthread_mutex_unlock(&g_estats_lock);
}

static void enc_flush(ocg_enc_backend *b) {
  oc_theora_state *st = &b->enc->state;
  ocg_dec_frame f;
  double t0 = enc_now_s();
  int pli;
  b->frame_open = 0;
  if (b->ncoded != b->geom.nfrags) enc_fatal("intra frame did not reconstruct every fragment");
  memset(&f, 0, sizeof(f));
  f.ref_idx[OCG_FRAME_GOLD] = f.ref_idx[OCG_FRAME_PREV] = -1;
  f.ref_idx[OCG_FRAME_SELF] = st->ref_frame_idx[OC_FRAME_SELF];
  f.lf_limit = st->loop_filter_limits[st->qis[0]];
  /* the records carry already-scaled DC terms: slot 0 = dequantised DC of a
     transformed block (analyze.c:803), slot 1 = the flat residual p of a
     DC-only block (analyze.c:790-794) as (32p+15)>>5 == p */
  for (pli = 0; pli < 3; pli++) { f.dc_quant[pli][0] = 1; f.dc_quant[pli][1] = 32; }
  f.ncoded = b->ncoded;
  f.intra_frame = 1;
  f.ncoeff_rows = b->nrows;
  /* An intra-only encoder never predicts from SELF, so the reconstruction stays
     on the device (ocg_backend_enc_copy_recon fetches it on demand) and the
     flush is asynchronous: the staging slots are double-buffered and the next
     frame's pre-pass queues behind these kernels on the same stream. */
  if (ocg_dec_submit(b->ctx, &f, NULL) < 0) enc_fatal("ocg_dec_submit failed");
  b->self_on_device = f.ref_idx[OCG_FRAME_SELF];
  pthread_mutex_lock(&g_estats_lock);
  g_estats.frames++;
  g_estats.coeff_rows += b->nrows;
  g_estats.h2d_bytes += (long)b->geom.nfrags * 16 + (long)b->nrows * 16;
  g_estats.flush_seconds += enc_now_s() - t0;
  pthread_mutex_unlock(&g_estats_lock);
}

/* ---- hooks --------------------------------------------------------------- */
static void ocge_enquant_table_fixup(void *_enquant[3][3][2], int _nqis) {
  /* _enquant is _enc->enquant (analyze.c:564) */
  oc_enc_ctx *enc = (oc_enc_ctx *)((char *)_enquant - offsetof(oc_enc_ctx, enquant));
  ocg_enc_backend *b = enc_backend_of(enc);
  oc_enc_enquant_table_fixup_c(_enquant, _nqis);
  if (b == NULL) enc_fatal("enquant_table_fixup from an unknown encoder");
  t_enc = b;
  enc_begin_frame(b, _nqis);
}

static ocg_enc_backend *enc_cur(void) {
  ocg_enc_backend *b = t_enc;
  if (b == NULL || !b->frame_open) enc_fatal("block hook outside an intra frame");
  return b;
}

static unsigned ocge_frag_intra_satd(int *_dc, const unsigned char *_src, int _ystride) {
  ocg_enc_backend *b = enc_cur();
  ptrdiff_t fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
  (void)_ystride;
  *_dc = b->tab.satd_dc[fragi];
  return b->tab.satd[fragi];
}

static void ocge_frag_sub_128(ogg_int16_t _diff[64], const unsigned char *_src, int _ystride) {
  /* the residual itself stays on the device; its only consumer is fdct8x8 */
  ocg_enc_backend *b = enc_cur();
  (void)_diff; (void)_ystride;
  b->cur_fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
  b->idct_pending = 0;
}

static void ocge_fdct8x8(ogg_int16_t _y[64], const ogg_int16_t _x[64]) {
  ocg_enc_backend *b = enc_cur();
  (void)_x;
  if (b->cur_fragi < 0) enc_fatal("fdct8x8 without a preceding frag_sub_128");
  memcpy(_y, b->tab.dct + (size_t)b->cur_fragi * 64, 64 * sizeof(ogg_int16_t));
}

static int ocge_quantize(ogg_int16_t _qdct[64], const ogg_int16_t _dct[64], const ogg_uint16_t _dequant[64],
                         const void *_enquant) {
  ocg_enc_backend *b = enc_cur();
  oc_enc_ctx *enc = b->enc;
  int pli, qii;
  size_t at;
  (void)_dct; (void)_dequant;
  if (b->cur_fragi < 0) enc_fatal("quantize without a preceding frag_sub_128");
  pli = b->st.recs[b->cur_fragi].pli_qti & 3;
  for (qii = 0; qii < b->nqis && enc->enquant[pli][qii][0] != _enquant; qii++) {}
  if (qii >= b->nqis) enc_fatal("quantize with a table that is not one of the frame's intra tables");
  at = (size_t)qii * (size_t)b->geom.nfrags + (size_t)b->cur_fragi;
  memcpy(_qdct, b->tab.qdct + at * 64, 64 * sizeof(ogg_int16_t));
  return b->tab.nonzero[at];
}

static inline int enc_row_nonzero(const ogg_int16_t *row) {
  ogg_uint64_t a, c;
  memcpy(&a, row, 8);
  memcpy(&c, row + 4, 8);
  return (a | c) != 0;
}

/* oc_idct8x8 (state.h:98, called at analyze.c:806 with the dequantised
   coefficients the tokeniser left in _x): take the rows the transform of this
   footprint reads (idct.c:327-329) and leave _x zeroed (idct.c:245,276,295). */
static void ocge_idct8x8(ogg_int16_t _y[64], ogg_int16_t _x[64], int _last_zzi) {
  ocg_enc_backend *b = enc_cur();
  int nr = _last_zzi <= 3 ? 2 : (_last_zzi <= 10 ? 4 : 8);
  int r, mask = 0;
  (void)_y;
  b->pend_dc = _x[0];
  _x[0] = 0; /* DC travels in the record */
  b->pend_row0 = (ogg_uint32_t)b->nrows;
  for (r = 0; r < nr; r++) {
    ogg_int16_t *row = _x + r * 8;
    if (enc_row_nonzero(row)) {
      memcpy(b->st.coeff_rows + (size_t)b->nrows * 8, row, 16);
      memset(row, 0, 16);
      b->nrows++;
      mask |= 1 << r;
    }
  }
  b->pend_mask = mask;
  /* never the DC-only shortcut of state.c:967: analyze.c:806 always transforms */
  b->pend_last_zzi = _last_zzi < 2 ? 2 : _last_zzi;
  b->idct_pending = 1;
}

static void ocge_frag_recon_intra(unsigned char *_dst, int _ystride, const ogg_int16_t _residue[64]) {
  ocg_enc_backend *b = enc_cur();
  ptrdiff_t fragi = enc_fragi_of(b, _dst, OC_FRAME_SELF);
  ocg_frag_rec *rec = b->st.recs + fragi;
  int pli = rec->pli_qti & 3;
  (void)_ystride;
  rec->mv = 0;
  rec->refi = OC_FRAME_SELF;
  if (b->idct_pending) {
    rec->coeff_row = b->pend_row0;
    rec->dc = b->pend_dc;
    rec->rowmask = (unsigned char)b->pend_mask;
    rec->last_zzi = (unsigned char)b->pend_last_zzi;
    rec->pli_qti = (unsigned char)pli;            /* dc scale slot 0 (x1) */
  } else {
    /* analyze.c:790-794: the block is flat, _residue[] == p everywhere */
    rec->coeff_row = (ogg_uint32_t)b->nrows;
    rec->dc = _residue[0];
    rec->rowmask = 0;
    rec->last_zzi = 0;
    rec->pli_qti = (unsigned char)(pli | 1 << 2); /* dc scale slot 1 (x32, >>5) */
  }
  b->idct_pending = 0;
  b->cur_fragi = -1;
  b->ncoded++;
}

static void ocge_state_loop_filter_frag_rows(const oc_theora_state *_state, signed char _bv[256], int _refi, int _pli,
                                             int _fragy0, int _fragy_end) {
  /* filtered on the device over the whole frame at flush */
  (void)_state; (void)_bv; (void)_refi; (void)_pli; (void)_fragy0; (void)_fragy_end;
}

static void ocge_restore_fpu(void) {
  ocg_enc_backend *b = t_enc;
  if (b != NULL && b->frame_open) enc_flush(b);
}

/* hooks that only inter frames use: reaching one is a configuration error */
static void ocge_no_sub(ogg_int16_t d[64], const unsigned char *s, const unsigned char *r, int y) {
  (void)d; (void)s; (void)r; (void)y; enc_fatal("frag_sub: inter-frame hook in the intra-only device encoder");
}
static unsigned ocge_no_sad(const unsigned char *s, const unsigned char *r, int y) {
  (void)s; (void)r; (void)y; enc_fatal("frag_sad/ssd: inter-frame hook in the intra-only device encoder"); return 0;
}
static unsigned ocge_no_sad_thresh(const unsigned char *s, const unsigned char *r, int y, unsigned t) {
  (void)s; (void)r; (void)y; (void)t; enc_fatal("frag_sad_thresh: inter-frame hook in the intra-only device encoder"); return 0;
}
static unsigned ocge_no_sad2_thresh(const unsigned char *s, const unsigned char *r1, const unsigned char *r2, int y, unsigned t) {
  (void)s; (void)r1; (void)r2; (void)y; (void)t; enc_fatal("frag_sad2_thresh: inter-frame hook in the intra-only device encoder"); return 0;
}
static unsigned ocge_no_satd(int *dc, const unsigned char *s, const unsigned char *r, int y) {
  (void)dc; (void)s; (void)r; (void)y; enc_fatal("frag_satd: inter-frame hook in the intra-only device encoder"); return 0;
}
static unsigned ocge_no_satd2(int *dc, const unsigned char *s, const unsigned char *r1, const unsigned char *r2, int y) {
  (void)dc; (void)s; (void)r1; (void)r2; (void)y; enc_fatal("frag_satd2: inter-frame hook in the intra-only device encoder"); return 0;
}
static unsigned ocge_no_border_ssd(const unsigned char *s, const unsigned char *r, int y, ogg_int64_t m) {
  (void)s; (void)r; (void)y; (void)m; enc_fatal("frag_border_ssd: inter-frame hook in the intra-only device encoder"); return 0;
}
static void ocge_no_copy2(unsigned char *d, const unsigned char *a, const unsigned char *c, int y) {
  (void)d; (void)a; (void)c; (void)y; enc_fatal("frag_copy2: inter-frame hook in the intra-only device encoder");
}
static void ocge_no_recon_inter(unsigned char *d, const unsigned char *s, int y, const ogg_int16_t r[64]) {
  (void)d; (void)s; (void)y; (void)r; enc_fatal("frag_recon_inter: inter-frame hook in the intra-only device encoder");
}
static void ocge_no_copy_list(unsigned char *d, const unsigned char *s, int y, const ptrdiff_t *f, ptrdiff_t n,
                              const ptrdiff_t *o) {
  (void)d; (void)s; (void)y; (void)f; (void)o;
  if (n > 0) enc_fatal("frag_copy_list: inter-frame hook in the intra-only device encoder");
}


/* ---- test instrumentation: analysis-pass snapshots of a host encoder ------- */
static void ocge_spy_fixup(void *_enquant[3][3][2], int _nqis) {
  static const int ROLE[5] = {OC_FRAME_IO, OC_FRAME_PREV_ORIG, OC_FRAME_GOLD_ORIG, OC_FRAME_PREV, OC_FRAME_GOLD};
  oc_enc_ctx *enc = (oc_enc_ctx *)((char *)_enquant - offsetof(oc_enc_ctx, enquant));
  oc_theora_state *st = &enc->state;
  ocg_enc_spy_frame f;
  ocg_me_mb *mb;
  unsigned char *refined;
  size_t fsz, mbi;
  int i, k;
  oc_enc_enquant_table_fixup_c(_enquant, _nqis);
  if (g_enc_spy == NULL) return;
  memset(&f, 0, sizeof(f));
  fsz = (size_t)(st->ref_frame_bufs[1][0].data - st->ref_frame_bufs[0][0].data);
  mb = (ocg_me_mb *)calloc(st->nmbs, sizeof(*mb));
  refined = (unsigned char *)calloc(st->nmbs, 1);
  if (mb == NULL || refined == NULL) { free(mb); free(refined); return; }
  for (mbi = 0; mbi < st->nmbs; mbi++) {
    const oc_mb_enc_info *e = enc->mb_info + mbi;
    for (i = 0; i < 3; i++) for (k = 0; k < 2; k++) mb[mbi].analysis_mv[i][k] = e->analysis_mv[i][k];
    for (k = 0; k < 2; k++) {
      mb[mbi].error[k] = e->error[k];
      mb[mbi].satd[k] = e->satd[k];
      mb[mbi].unref_mv[k] = e->unref_mv[k];
    }
    for (k = 0; k < 4; k++) {
      mb[mbi].block_mv[k] = e->block_mv[k];
      mb[mbi].ref_mv[k] = e->ref_mv[k];
      mb[mbi].block_satd[k] = mb[mbi].ref_block_satd[k] = e->block_satd[k];
    }
    refined[mbi] = e->refined;
  }
  f.frame_type = st->frame_type;
  f.prevframe_dropped = enc->prevframe_dropped;
  f.sp_level = enc->sp_level;
  f.keyframe_frequency_force = (int)enc->keyframe_frequency_force;
  f.nmbs = (int)st->nmbs;
  f.curframe_num = st->curframe_num;
  f.ref_frame_sz = (ogg_int64_t)fsz;
  for (i = 0; i < 5; i++)
    f.frames[i] = st->ref_frame_idx[ROLE[i]] >= 0 ? st->ref_frame_handle + (size_t)st->ref_frame_idx[ROLE[i]] * fsz : NULL;
  f.state = mb;
  f.refined = refined;
  (*g_enc_spy)(g_enc_spy_user, &f);
  free(mb);
  free(refined);
}

/* ---- set-up / tear-down --------------------------------------------------- */
static void enc_backend_destroy(ocg_enc_backend *b) {
  ocg_enc_backend **pp;
  if (b == NULL) return;
  pthread_mutex_lock(&g_elock);
  for (pp = &g_elist; *pp != NULL && *pp != b; pp = &(*pp)->next) {}
  if (*pp == b) *pp = b->next;
  pthread_mutex_unlock(&g_elock);
  if (t_enc == b) t_enc = NULL;
  if (b->ctx != NULL) {
    ocg_ctx_sync(b->ctx);
    if (b->pinned) ocg_host_unregister(b->enc->state.ref_frame_handle);
    ocg_ctx_destroy(b->ctx);
  }
  free(b->off2frag);
  free(b);
}

#if defined(OC_X86_ASM)
void oc_refimpl_state_accel_init_x86(oc_theora_state *_state); /* lib/x86/x86state.c, renamed */
void oc_refimpl_enc_accel_init_x86(oc_enc_ctx *_enc);          /* lib/x86/x86enc.c, renamed */
void oc_enc_accel_init_ocg(oc_enc_ctx *_enc);
/* x86enc.h names oc_enc_accel_init_x86 as the encoder's init function (see ocg_backend.c) */
void oc_enc_accel_init_x86(oc_enc_ctx *_enc) { oc_enc_accel_init_ocg(_enc); }
#endif

void oc_enc_accel_init_ocg(oc_enc_ctx *_enc) {
  oc_theora_state *st = &_enc->state;
  ocg_enc_backend *b;
  ptrdiff_t fragi, omin, omax;
  oc_enc_accel_init_c(_enc);
  t_enc_init_failed = 0;
  /* only an encoder that cannot emit inter frames takes the device path */
  if (g_enc_mode == OCG_ENC_HOST || st->info.keyframe_granule_shift != 0) {
#if defined(OC_X86_ASM)
    /* everything stays on the host: with the reference's SIMD kernels, exactly as its own x86 build
       would set this context up (x86state.c:66-95, x86enc.c:21-62; oc_enc_init sizes the quantiser
       tables after this call, encode.c:1160-1190) */
    if (g_enc_mode != OCG_ENC_HOST) {
      oc_refimpl_state_accel_init_x86(st);
      oc_refimpl_enc_accel_init_x86(_enc);
    }
#endif
    /* the spy wraps the C fix-up: host (C kernel) mode only */
    if (g_enc_spy != NULL && g_enc_mode == OCG_ENC_HOST) _enc->opt_vtable.enquant_table_fixup = ocge_spy_fixup;
    return;
  }
  t_enc_init_failed = 1;
  b = (ocg_enc_backend *)calloc(1, sizeof(*b));
  if (b == NULL) return;
  b->enc = _enc;
  b->self_on_device = -1;
  if (ocg_geometry_init(&b->geom, (int)st->info.frame_width, (int)st->info.frame_height, (int)st->info.pixel_fmt, 6) < 0) {
    fprintf(stderr, "theora_b200 encoder back-end: %s\n", ocg_last_error());
    free(b);
    return;
  }
  /* the device mirror must be byte-compatible with state.c:545-671 */
  if (b->geom.nfrags != st->nfrags || b->geom.planes[0].ystride != st->ref_ystride[0] ||
      b->geom.planes[1].ystride != st->ref_ystride[1] ||
      st->ref_frame_bufs[0][0].data - st->ref_frame_handle != b->geom.base_off ||
      st->ref_frame_bufs[1][0].data - st->ref_frame_bufs[0][0].data != b->geom.ref_frame_sz) {
    fprintf(stderr, "theora_b200 encoder back-end: frame layout differs from the reference's\n");
    free(b);
    return;
  }
  omin = omax = st->frag_buf_offs[0];
  for (fragi = 1; fragi < st->nfrags; fragi++) {
    if (st->frag_buf_offs[fragi] < omin) omin = st->frag_buf_offs[fragi];
    if (st->frag_buf_offs[fragi] > omax) omax = st->frag_buf_offs[fragi];
  }
  b->off_min = omin;
  b->noff = (size_t)((omax - omin) >> 3) + 1;
  b->off2frag = (ogg_int32_t *)malloc(b->noff * sizeof(ogg_int32_t));
  if (b->off2frag == NULL) { free(b); return; }
  memset(b->off2frag, 0xFF, b->noff * sizeof(ogg_int32_t));
  for (fragi = 0; fragi < st->nfrags; fragi++) {
    ptrdiff_t k = st->frag_buf_offs[fragi] - omin;
    if ((k & 7) != 0 || b->off2frag[k >> 3] >= 0) {
      fprintf(stderr, "theora_b200 encoder back-end: fragment offsets are not 8-byte distinct\n");
      free(b->off2frag);
      free(b);
      return;
    }
    b->off2frag[k >> 3] = (ogg_int32_t)fragi;
  }
  if (ocg_ctx_create(&b->ctx, &b->geom, ocg_backend_device_()) < 0 ||
      (ocg_enc_intra_reserve(b->ctx) < 0 && (ocg_ctx_destroy(b->ctx), 1))) {
    fprintf(stderr, "theora_b200 encoder back-end: %s\n", ocg_last_error());
    free(b->off2frag);
    free(b);
    return; /* th_encode_alloc (below) reports the failure; no CPU fallback for an intra-only encoder */
  }
  b->pinned = ocg_host_register(st->ref_frame_handle, (size_t)b->geom.ref_frame_sz * 6) == 0;
  /* pre-pass look-ups */
  _enc->opt_vtable.enquant_table_fixup = ocge_enquant_table_fixup;
  _enc->opt_vtable.frag_intra_satd = ocge_frag_intra_satd;
  _enc->opt_vtable.frag_sub_128 = ocge_frag_sub_128;
  _enc->opt_vtable.fdct8x8 = ocge_fdct8x8;
  _enc->opt_vtable.quantize = ocge_quantize;
  /* recorded reconstruction */
  _enc->opt_vtable.frag_recon_intra = ocge_frag_recon_intra;
  st->opt_vtable.idct8x8 = ocge_idct8x8;
  st->opt_vtable.state_loop_filter_frag_rows = ocge_state_loop_filter_frag_rows;
  st->opt_vtable.restore_fpu = ocge_restore_fpu;
  /* inter-only hooks */
  _enc->opt_vtable.frag_sub = ocge_no_sub;
  _enc->opt_vtable.frag_sad = ocge_no_sad;
  _enc->opt_vtable.frag_sad_thresh = ocge_no_sad_thresh;
  _enc->opt_vtable.frag_sad2_thresh = ocge_no_sad2_thresh;
  _enc->opt_vtable.frag_satd = ocge_no_satd;
  _enc->opt_vtable.frag_satd2 = ocge_no_satd2;
  _enc->opt_vtable.frag_ssd = ocge_no_sad;
  _enc->opt_vtable.frag_border_ssd = ocge_no_border_ssd;
  _enc->opt_vtable.frag_copy2 = ocge_no_copy2;
  _enc->opt_vtable.frag_recon_inter = ocge_no_recon_inter;
  st->opt_vtable.frag_copy_list = ocge_no_copy_list;
  pthread_mutex_lock(&g_elock);
  b->next = g_elist;
  g_elist = b;
  pthread_mutex_unlock(&g_elock);
  t_enc = b;
  t_enc_created = b;
  t_enc_init_failed = 0;
}

/* ---- public API wrappers -------------------------------------------------- */
th_enc_ctx *th_encode_alloc(const th_info *_info) {
  th_enc_ctx *enc;
  t_enc_init_failed = 0;
  t_enc_created = NULL;
  enc = oc_refimpl_encode_alloc(_info);
  if (enc == NULL && t_enc_created != NULL) enc_backend_destroy(t_enc_created); /* oc_enc_init failed later */
  t_enc_created = NULL;
  if (enc != NULL && t_enc_init_failed) {
    /* an intra-only encoder without a usable device: fail the allocation */
    oc_refimpl_encode_free(enc);
    return NULL;
  }
  return enc;
}

void th_encode_free(th_enc_ctx *_enc) {
  if (_enc != NULL) enc_backend_destroy(enc_backend_of(_enc));
  oc_refimpl_encode_free(_enc);
}

/* Test accessor: the encoder's current reconstruction (OC_FRAME_SELF of the
   frame just coded) as top-down planes packed back to back, frame_width x
   frame_height, the way refh_dec_copy_frame lays out a decoded frame. */
OCG_API long ocg_backend_enc_copy_recon(th_enc_ctx *_enc, unsigned char *_dst) {
  const oc_theora_state *st;
  long n = 0;
  int pli, y, idx;
  if (_enc == NULL || _dst == NULL) return TH_EFAULT;
  st = &_enc->state;
  idx = st->ref_frame_idx[OC_FRAME_SELF];
  if (idx < 0) return TH_EINVAL;
  {
    ocg_enc_backend *b = enc_backend_of(_enc);
    if (b != NULL && b->self_on_device == idx) {
      if (ocg_ctx_download_frame(b->ctx, idx, (unsigned char *)st->ref_frame_handle + (size_t)idx * (size_t)b->geom.ref_frame_sz) < 0 ||
          ocg_ctx_sync(b->ctx) < 0)
        return TH_EFAULT;
      b->self_on_device = -1;
    }
  }
  for (pli = 0; pli < 3; pli++) {
    const th_img_plane *p = &st->ref_frame_bufs[idx][pli];
    /* data points at the displayed-bottom row, stride is negative (state.c:622-629) */
    for (y = 0; y < p->height; y++) {
      memcpy(_dst + n, p->data + (ptrdiff_t)(p->height - 1 - y) * p->stride, (size_t)p->width);
      n += p->width;
    }
  }
  return n;
}
