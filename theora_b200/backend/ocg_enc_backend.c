/* ocg_enc_backend.c -- encoder half of the vtable back-end.
 *
 * Every encoder hook (encint.h:292-325) returns its result synchronously to
 * serial host code (mode decision, R-D tokeniser), so none of them can launch
 * a kernel per call.  What the device does instead:
 *
 * INTRA frames.  frag_intra_satd(src) (analyze.c:1385-1534) and
 * frag_sub_128(src) -> fdct8x8 -> quantize (analyze.c:725-782) are functions
 * of the input frame and the frame's quantiser tables only.  The first hook of
 * an analysis pass, enquant_table_fixup (analyze.c:564), runs ONE batched
 * device pre-pass over all fragments (ocg_enc_intra_prepass) and the per-block
 * hooks become table look-ups keyed by the block's source pointer.
 *
 * EVERY frame.  What happens after the tokeniser -- idct8x8 + frag_recon_*
 * (analyze.c:803-822), the uncoded-fragment copy, the loop filter and the
 * border fill -- is RECORDED like the decoder's state_frag_recon and flushed as
 * one CUDA graph (ocg_dec_flush) at the restore_fpu that opens
 * oc_enc_frame_pack (encode.c:911): the reconstructed reference frames live on
 * the device; for an encoder that can emit inter frames the finished frame is
 * also copied back into the host's buffer.
 *
 * INTER frames (encoders with keyframe_granule_shift > 0).  The analysis loop
 * compares the SSD of the block it has just reconstructed (analyze.c:825-868)
 * and its mode costs depend on the running entropy state, so the loop stays on
 * the host; see DESIGN.md section 1 for what the device serves to it and what
 * the loop still computes itself (with the reference's plain C kernels: no
 * lib/x86 code is linked).
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include "encint.h"
#include "ocg_backend.h"

/* the reference's own entry points, renamed on encode.c's / mcenc.c's command line */
th_enc_ctx *oc_refimpl_encode_alloc(const th_info *_info);
void oc_refimpl_encode_free(th_enc_ctx *_enc);
int oc_refimpl_encode_ycbcr_in(th_enc_ctx *_enc, th_ycbcr_buffer _img);
void oc_refimpl_mcenc_search(oc_enc_ctx *_enc, int _mbi);
void oc_refimpl_mcenc_refine1mv(oc_enc_ctx *_enc, int _mbi, int _frame);
void oc_refimpl_mcenc_refine4mv(oc_enc_ctx *_enc, int _mbi);

int ocg_backend_device_(void); /* ocg_backend.c */

typedef struct ocg_enc_backend {
  oc_enc_ctx          *enc;
  ocg_ctx             *ctx;
  ocg_geometry         geom;
  ocg_staging          st;
  ocg_enc_intra_tables tab;
  int                  nqis;
  int                  frame_open;     /* an analysis pass is being recorded */
  int                  tables;         /* the intra pre-pass tables serve this pass */
  int                  inter_capable;  /* keyframe_granule_shift > 0: inter frames possible, host frames kept current */
  int                  inter_frame;    /* the pass in progress analyses an inter frame */
  int                  pending;        /* a flushed frame may still be running / copying back */
  int                  failed;         /* latched: hooks fall through to the C kernels, the API returns TH_EFAULT */
  int                  ncoded;
  int                  nrows;
  int                  pinned;
  int                  self_on_device; /* buffer index whose reconstruction has not been copied to the host, or -1 */
  /* source/destination pointer -> fragment index */
  ogg_int32_t         *off2frag;
  ptrdiff_t            off_min;
  size_t               noff;
  /* block in flight: sub_128 -> fdct8x8 -> quantize -> [idct8x8] -> recon */
  ptrdiff_t            cur_fragi;
  int                  idct_pending;
  int                  pend_last_zzi;
  ogg_uint32_t         pend_row0;
  int                  pend_mask;
  ogg_int16_t          pend_dc;
  /* whole-frame motion analysis on the device (oc_mcenc_search / refine*, served as look-ups) */
  ocg_me              *me;
  ocg_me_mb           *me_tab;         /* page-locked: state upload, then the device's results for this frame */
  int                  me_valid;       /* me_tab holds the results of the frame being analysed */
  int                  me_flags;
  int                  me_seen_frame;  /* a pass for me_frame_num has run */
  ogg_int64_t          me_frame_num;
  unsigned char       *gold_dirty;     /* [nmbs] the macro block's final GOLD vector/error differ from the speculation */
  unsigned char       *gold_fixed;     /* [nmbs] its GOLD search was redone (gold_fix holds the refinement) */
  struct { ogg_int16_t mv; ogg_uint32_t satd; } *gold_fix;
  /* inter-frame analysis tables (intra SATD, skip SSD, SATD of the candidate predictors) */
  ocg_enc_inter       *ei;
  ocg_enc_inter_tables itab;
  int                  itab_valid;     /* itab holds the tables of the frame being analysed */
  ogg_int32_t         *border_slot;    /* [nfrags] index into itab.border_ssd, or -1 */
  ogg_int64_t         *border_mask;    /* [nfrags] the mask that slot was computed with */
  const unsigned char *pool0;          /* ref_frame_handle + base_off: what the candidates' tap offsets are relative to */
  /* speculative sub + fDCT + quantiser tables (itab.fq_*) */
  int                  fq_ok;          /* they serve the pass in progress (same quantisers as when they were made) */
  int                  fq_qis[3], fq_nq;
  long                 fq_cur;         /* entry serving the block in flight (frag_sub -> fdct8x8 -> quantize), or -1 */
  int                  fq_pli;
  const unsigned char *c2_dst, *c2_r1, *c2_r2; /* the last frag_copy2 */
  long                 n_fq_hit, n_fq_miss;
  long                 n_me_repairs, n_gold_refines;
  long                 n_satd_hit, n_satd_miss, n_ssd_hit, n_ssd_host, n_isatd_hit; /* per pass, folded into the stats at the flush */
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
static double g_dbg_t[8];
static long g_dbg_n;
/* diagnostic switches, read once: OCG_ENC_TIMING (pass timing + miss histogram at exit), OCG_ENC_NO_FQ (no
   speculative transform tables) */
static int enc_dbg_flag(int which) {
  static int flags = -1;
  if (flags < 0) flags = (getenv("OCG_ENC_TIMING") != NULL ? 1 : 0) | (getenv("OCG_ENC_NO_FQ") != NULL ? 2 : 0);
  return (flags >> which) & 1;
}
static long g_dbg_miss[10]; /* fq misses by the candidate that would have served them; [8] none; [9] sub_128 in inter frames */

OCG_API void ocg_backend_get_enc_stats(ocg_enc_backend_stats *out, int reset) {
  pthread_mutex_lock(&g_estats_lock);
  if (out) *out = g_estats;
  if (reset) { memset(&g_estats, 0, sizeof(g_estats)); memset(g_dbg_t, 0, sizeof(g_dbg_t)); g_dbg_n = 0; }
  pthread_mutex_unlock(&g_estats_lock);
}

static double enc_now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* The hooks cannot report errors (encint.h:292-325: they return data).  A failed device call or a violated
   assumption is latched: from then on every hook falls through to the reference's C kernel, so the host
   code above keeps running on defined data, nothing touches the device any more, and th_encode_ycbcr_in
   returns TH_EFAULT (the encoder's reference frames are no longer trustworthy). */
static void enc_fail(ocg_enc_backend *b, const char *what) {
  if (b == NULL || !b->failed) fprintf(stderr, "theora_b200 encoder back-end: %s (%s)\n", what, ocg_last_error());
#if defined(OCG_BACKEND_ABORT_ON_ERROR)
  abort();
#endif
  if (b != NULL) {
    b->failed = 1;
    b->frame_open = 0;
    b->tables = 0;
    b->pending = 0;
  }
}

static ocg_enc_backend *enc_backend_of(const oc_enc_ctx *enc) {
  ocg_enc_backend *b = t_enc;
  if (b != NULL && b->enc == enc) return b;
  pthread_mutex_lock(&g_elock);
  for (b = g_elist; b != NULL && b->enc != enc; b = b->next) {}
  pthread_mutex_unlock(&g_elock);
  return b;
}

/* Fragment whose row 0 is at p inside the buffer playing role `frame`, or -1. */
static inline ptrdiff_t enc_fragi_of(const ocg_enc_backend *b, const unsigned char *p, int frame) {
  const unsigned char *base = b->enc->state.ref_frame_data[frame];
  size_t k = (size_t)((p - base) - b->off_min);
  if ((k & 7) != 0 || (k >> 3) >= b->noff) return -1;
  return b->off2frag[k >> 3];
}

static void enc_wait(ocg_enc_backend *b) {
  if (b->pending) {
    b->pending = 0;
    if (ocg_dec_wait(b->ctx) < 0) enc_fail(b, "ocg_dec_wait failed");
  }
}

/* ---- motion analysis pre-pass ---------------------------------------------
   oc_mcenc_search for macro block m reads the vectors and errors of m's already-analysed neighbours
   (mcenc.c:90-164, 331-337) AFTER their half-pel refinement.  Against OC_FRAME_PREV that refinement always
   runs (analyze.c:2486-2489), so the whole PREV chain is a function of the frames alone and the device's
   wave-front reproduces it exactly.  Against OC_FRAME_GOLD it runs only where the serial mode decision
   asks for it (analyze.c:2476-2485), which is not known here: the device runs the GOLD chain WITHOUT
   refinements and also reports what each refinement would give; the look-ups below check, per macro
   block, whether the neighbours' final GOLD vectors/errors still produce the candidate set the device
   used, and redo the one search on the device where they do not (ocg_me_repair). */
/* Static inputs of the inter-frame tables: which fragments make up each macro block (state.mb_maps) and
   which fragments straddle the picture border, with their pixel masks (state.c:473-543). */
static int enc_inter_setup(ocg_enc_backend *b) {
  oc_theora_state *st = &b->enc->state;
  ogg_int32_t *mbfrags = (ogg_int32_t *)malloc(st->nmbs * 12 * sizeof(*mbfrags));
  ogg_int32_t *bfragi = (ogg_int32_t *)malloc((size_t)st->nfrags * sizeof(*bfragi));
  ogg_int64_t *bmask = (ogg_int64_t *)malloc((size_t)st->nfrags * sizeof(*bmask));
  ptrdiff_t fragi;
  size_t mbi;
  int pli, bi, nb = 0, r = -1;
  b->border_slot = (ogg_int32_t *)malloc((size_t)st->nfrags * sizeof(*b->border_slot));
  b->border_mask = (ogg_int64_t *)calloc((size_t)st->nfrags, sizeof(*b->border_mask));
  if (mbfrags != NULL && bfragi != NULL && bmask != NULL && b->border_slot != NULL && b->border_mask != NULL) {
    for (mbi = 0; mbi < st->nmbs; mbi++)
      for (pli = 0; pli < 3; pli++)
        for (bi = 0; bi < 4; bi++)
          mbfrags[mbi * 12 + pli * 4 + bi] = st->mb_modes[mbi] == OC_MODE_INVALID ? -1 : (ogg_int32_t)st->mb_maps[mbi][pli][bi];
    for (fragi = 0; fragi < st->nfrags; fragi++) {
      if (st->frags[fragi].borderi >= 0) {
        bfragi[nb] = (ogg_int32_t)fragi;
        bmask[nb] = st->borders[st->frags[fragi].borderi].mask;
        b->border_mask[fragi] = bmask[nb];
        nb++;
      }
    }
    if (ocg_enc_inter_create(&b->ei, b->ctx, b->me, mbfrags, bfragi, bmask, nb) == 0) {
      for (fragi = 0; fragi < st->nfrags; fragi++) b->border_slot[fragi] = ocg_enc_inter_border_slot(b->ei, (int)fragi);
      b->pool0 = st->ref_frame_handle + b->geom.base_off;
      r = 0;
    }
  }
  free(mbfrags);
  free(bfragi);
  free(bmask);
  if (r < 0) enc_fail(b, "inter-frame table set-up failed");
  return r;
}

static int enc_me_prepass(ocg_enc_backend *b) {
  oc_enc_ctx *enc = b->enc;
  oc_theora_state *st = &enc->state;
  static const int ROLE[5] = {OC_FRAME_IO, OC_FRAME_PREV_ORIG, OC_FRAME_GOLD_ORIG, OC_FRAME_PREV, OC_FRAME_GOLD};
  const int first_pass = !(b->me_seen_frame && b->me_frame_num == st->curframe_num);
  const double t_enter = enc_now_s();
  int bufs[5], i, k, flags, wanted;
  size_t mbi;
  if (!first_pass) return 0; /* itab too: functions of the frames and the first pass's vectors */
  b->me_seen_frame = 1;
  b->me_frame_num = st->curframe_num;
  b->me_valid = 0;
  b->itab_valid = 0;
  /* who will call oc_mcenc_search in this pass: analyze.c:1723-1726 (key frames), 2402 (inter frames) */
  wanted = enc->sp_level < OC_SP_LEVEL_NOSATD &&
           (b->inter_frame ? 1 : st->curframe_num > 0 && enc->keyframe_frequency_force > 1);
  for (i = 0; i < 5; i++) {
    bufs[i] = st->ref_frame_idx[ROLE[i]];
    if (bufs[i] < 0) wanted = 0;
  }
  /* the input frame joins the device's frame pool.  A key frame's own pre-pass uploads it too, but only
     after this hook: a key frame that is searched (every key frame but the first) needs it here already */
  if ((b->inter_frame || wanted) &&
      ocg_ctx_upload_frame(b->ctx, st->ref_frame_idx[OC_FRAME_IO],
                           st->ref_frame_handle + (size_t)st->ref_frame_idx[OC_FRAME_IO] * (size_t)b->geom.ref_frame_sz) < 0) {
    enc_fail(b, "input frame upload failed");
    return -1;
  }
  /* device objects are made on the encoder's first pass (frame 0), whoever wants them first */
  if (b->me == NULL) {
    /* the reference's own tables: mb_maps (state.c:300-330), cneighbors (encode.c:967-1048) */
    ocg_me_topo *topo = (ocg_me_topo *)calloc(st->nmbs, sizeof(*topo));
    if (topo == NULL) { enc_fail(b, "out of memory"); return -1; }
    for (mbi = 0; mbi < st->nmbs; mbi++) {
      const oc_mb_enc_info *e = enc->mb_info + mbi;
      if (st->mb_modes[mbi] == OC_MODE_INVALID) continue;
      topo[mbi].valid = 1;
      topo[mbi].ncn = e->ncneighbors;
      for (k = 0; k < e->ncneighbors; k++) topo[mbi].cn[k] = (ogg_int32_t)e->cneighbors[k];
      for (k = 0; k < 4; k++) topo[mbi].frag_off[k] = (ogg_int32_t)st->frag_buf_offs[st->mb_maps[mbi][0][k]];
    }
    i = ocg_me_create(&b->me, b->ctx, topo);
    free(topo);
    b->me_tab = (ocg_me_mb *)calloc(st->nmbs, sizeof(*b->me_tab));
    b->gold_dirty = (unsigned char *)calloc(st->nmbs, 1);
    b->gold_fixed = (unsigned char *)calloc(st->nmbs, 1);
    b->gold_fix = calloc(st->nmbs, sizeof(*b->gold_fix));
    if (i < 0 || b->me_tab == NULL || b->gold_dirty == NULL || b->gold_fixed == NULL || b->gold_fix == NULL ||
        ocg_host_register(b->me_tab, st->nmbs * sizeof(*b->me_tab)) < 0) {
      enc_fail(b, "motion analysis set-up failed");
      return -1;
    }
    if (enc_inter_setup(b) < 0) return -1;
  }
  if (!wanted) return 0;
  for (mbi = 0; mbi < st->nmbs; mbi++) {
    const oc_mb_enc_info *e = enc->mb_info + mbi;
    ocg_me_mb *m = b->me_tab + mbi;
    for (i = 0; i < 3; i++) for (k = 0; k < 2; k++) m->analysis_mv[i][k] = e->analysis_mv[i][k];
    for (k = 0; k < 2; k++) { m->error[k] = e->error[k]; m->satd[k] = e->satd[k]; }
    for (k = 0; k < 4; k++) { m->block_mv[k] = e->block_mv[k]; m->block_satd[k] = e->block_satd[k]; }
  }
  flags = OCG_ME_SPEC_GOLD;
  if (b->inter_frame) flags |= OCG_ME_REFINE_PREV;
  if (enc->sp_level >= OC_SP_LEVEL_FAST_ANALYSIS) flags |= OCG_ME_FAST;
  else if (b->inter_frame) flags |= OCG_ME_REFINE_4MV;
  if (enc->prevframe_dropped) flags |= OCG_ME_DROPPED;
  b->itab_valid = 0;
  b->fq_nq = 0;
  {
    const double tq0 = enc_now_s();
    double tq1;
    int pli, qii, zzi, r = 0;
    pthread_mutex_lock(&g_estats_lock);
    g_dbg_t[4] += tq0 - t_enter;
    pthread_mutex_unlock(&g_estats_lock);
    if (b->inter_frame) {
      /* the frame's inter quantisers, condensed by analyze.c:544-564 just before this hook */
      for (pli = 0; pli < 3; pli++)
        for (qii = 0; qii < b->nqis; qii++) {
          const oc_iquant *iq = (const oc_iquant *)enc->enquant[pli][qii][1];
          memcpy(b->dequant[pli][1][qii], enc->dequant[pli][qii][1], 64 * sizeof(ogg_uint16_t));
          for (zzi = 0; zzi < 64; zzi++) {
            b->enquant[pli][1][qii][zzi][0] = iq[zzi].m;
            b->enquant[pli][1][qii][zzi][1] = iq[zzi].l;
          }
        }
      r = ocg_enc_inter_quant_tables(b->ei, &b->dequant[0][0][0][0], &b->enquant[0][0][0][0][0], b->nqis);
      for (qii = 0; qii < 3; qii++) b->fq_qis[qii] = st->qis[qii < st->nqis ? qii : 0];
      b->fq_nq = b->nqis;
    }
    if (r < 0 || ocg_me_write_async(b->me, b->me_tab) < 0 || ocg_me_frame(b->me, bufs, flags, NULL) < 0 ||
        ocg_me_read_async(b->me, b->me_tab) < 0 ||
        (b->inter_frame && ocg_enc_inter_prepass(b->ei, bufs[0], bufs[3], bufs[4], 1, &b->itab) < 0)) {
      enc_fail(b, "motion analysis on the device failed");
      return -1;
    }
    tq1 = enc_now_s();
    if (ocg_ctx_sync(b->ctx) < 0 || (b->inter_frame && ocg_enc_inter_finish(b->ei, &b->itab) < 0)) {
      enc_fail(b, "motion analysis on the device failed");
      return -1;
    }
    pthread_mutex_lock(&g_estats_lock);
    g_estats.me_queue_seconds += tq1 - tq0;
    g_estats.me_sync_seconds += enc_now_s() - tq1;
    pthread_mutex_unlock(&g_estats_lock);
  }
  b->itab_valid = b->inter_frame;
  memset(b->gold_dirty, 0, st->nmbs);
  memset(b->gold_fixed, 0, st->nmbs);
  b->me_flags = flags;
  b->me_valid = 1;
  pthread_mutex_lock(&g_estats_lock);
  g_estats.me_frames++;
  g_estats.h2d_bytes += (long)(b->inter_frame ? b->geom.ref_frame_sz : 0) + (long)(st->nmbs * sizeof(ocg_me_mb));
  g_estats.d2h_bytes += (long)(st->nmbs * sizeof(ocg_me_mb)) + (b->inter_frame ? b->itab.d2h_bytes : 0);
  pthread_mutex_unlock(&g_estats_lock);
  return 0;
}

__attribute__((destructor)) static void enc_dbg_print(void) {
  if (enc_dbg_flag(0) && g_dbg_n > 0)
    fprintf(stderr, "[enc timing, ms per pass over %ld passes] wait %.3f staging %.3f me_total %.3f (prep %.3f) rest %.3f\n", g_dbg_n,
            1e3 * g_dbg_t[0] / g_dbg_n, 1e3 * g_dbg_t[1] / g_dbg_n, 1e3 * g_dbg_t[2] / g_dbg_n, 1e3 * g_dbg_t[4] / g_dbg_n,
            1e3 * g_dbg_t[3] / g_dbg_n);
  if (enc_dbg_flag(0))
    fprintf(stderr, "[enc fq misses by candidate] %ld %ld %ld %ld %ld %ld %ld %ld none %ld; sub_128 in inter frames %ld\n", g_dbg_miss[0],
            g_dbg_miss[1], g_dbg_miss[2], g_dbg_miss[3], g_dbg_miss[4], g_dbg_miss[5], g_dbg_miss[6], g_dbg_miss[7], g_dbg_miss[8],
            g_dbg_miss[9]);
}

/* ---- frame life cycle ---------------------------------------------------- */
/* enquant_table_fixup, analyze.c:564: the first hook of every analysis pass (oc_enc_pipeline_init); the
   input frame is in OC_FRAME_IO, the references are rotated, state.frame_type says which analysis runs. */
static void enc_begin_pass(ocg_enc_backend *b, int nqis) {
  oc_enc_ctx *enc = b->enc;
  oc_theora_state *st = &enc->state;
  const unsigned char *host_io;
  double t0 = enc_now_s();
  int pli, qii, zzi;
  b->frame_open = 0;
  b->tables = 0;
  if (b->failed) return;
  /* the previous frame's reconstruction must have reached the host buffers the C kernels of an inter
     analysis read (and its staging slots are free again) */
  enc_wait(b);
  if (b->failed) return;
  {
    const double tw = enc_now_s() - t0;
    pthread_mutex_lock(&g_estats_lock);
    g_estats.prev_wait_seconds += tw;
    pthread_mutex_unlock(&g_estats_lock);
  }
  if (nqis < 1 || nqis > 3) { enc_fail(b, "unexpected quantiser count"); return; }
  b->inter_frame = st->frame_type != OC_INTRA_FRAME;
  if (b->inter_frame && !b->inter_capable) { enc_fail(b, "inter frame in an encoder that was set up as intra-only"); return; }
  /* a pass that was analysed but never packed (dry run, re-analysis as a key frame) is simply dropped */
  const double tA = enc_now_s();
  if (ocg_dec_staging(b->ctx, &b->st) < 0) { enc_fail(b, "ocg_dec_staging failed"); return; }
  const double tB = enc_now_s();
  b->ncoded = b->nrows = 0;
  b->cur_fragi = -1;
  b->idct_pending = 0;
  b->nqis = nqis;
  b->n_satd_hit = b->n_satd_miss = b->n_ssd_hit = b->n_ssd_host = b->n_isatd_hit = 0;
  b->n_fq_hit = b->n_fq_miss = 0;
  if (b->inter_capable && enc_me_prepass(b) < 0) return;
  const double tC = enc_now_s();
  pthread_mutex_lock(&g_estats_lock);
  g_dbg_t[0] += tA - t0; g_dbg_t[1] += tB - tA; g_dbg_t[2] += tC - tB; g_dbg_n++;
  pthread_mutex_unlock(&g_estats_lock);
  /* the speculative transform tables were made for one set of quantisers: a re-analysis with another uses
     the C kernels */
  b->fq_cur = -1;
  b->c2_dst = NULL;
  b->fq_ok = b->inter_frame && b->itab_valid && b->itab.fq_nqis > 0 && b->fq_nq == nqis && !enc_dbg_flag(1);
  if (b->fq_ok) {
    int qii;
    for (qii = 0; qii < nqis; qii++) if (b->fq_qis[qii] != st->qis[qii]) b->fq_ok = 0;
  }
  if (!b->inter_frame) {
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
                              &b->enquant[0][0][0][0][0], nqis, &b->tab) < 0) {
      enc_fail(b, "ocg_enc_intra_prepass failed");
      return;
    }
    b->tables = 1;
    pthread_mutex_lock(&g_estats_lock);
    g_estats.h2d_bytes += (long)b->geom.ref_frame_sz;
    g_estats.d2h_bytes += (long)b->geom.nfrags * (8 + 4 * nqis + 128 + 128 * nqis);
    pthread_mutex_unlock(&g_estats_lock);
  }
  b->frame_open = 1;
  pthread_mutex_lock(&g_estats_lock);
  g_estats.prepass_frames++;
  g_estats.prepass_seconds += enc_now_s() - t0;
  g_dbg_t[3] += enc_now_s() - tC;
  pthread_mutex_unlock(&g_estats_lock);
}

/* restore_fpu at encode.c:911 (oc_enc_frame_pack): the analysis pass that will be packed is complete. */
static void enc_flush(ocg_enc_backend *b) {
  oc_theora_state *st = &b->enc->state;
  ocg_dec_frame f;
  double t0 = enc_now_s();
  int pli, self = st->ref_frame_idx[OC_FRAME_SELF];
  b->frame_open = 0;
  b->tables = 0;
  if (!b->inter_frame && b->ncoded != b->geom.nfrags) { enc_fail(b, "intra frame did not reconstruct every fragment"); return; }
  memset(&f, 0, sizeof(f));
  f.ref_idx[OCG_FRAME_GOLD] = b->inter_frame ? st->ref_frame_idx[OC_FRAME_GOLD] : -1;
  f.ref_idx[OCG_FRAME_PREV] = b->inter_frame ? st->ref_frame_idx[OC_FRAME_PREV] : -1;
  f.ref_idx[OCG_FRAME_SELF] = self;
  f.lf_limit = st->loop_filter_limits[st->qis[0]];
  /* the records carry already-scaled DC terms: slot 0 = dequantised DC of a
     transformed block (analyze.c:803), slot 1 = the flat residual p of a
     DC-only block (analyze.c:790-794) as (32p+15)>>5 == p */
  for (pli = 0; pli < 3; pli++) { f.dc_quant[pli][0] = 1; f.dc_quant[pli][1] = 32; }
  f.ncoded = b->ncoded;
  f.intra_frame = !b->inter_frame;
  f.ncoeff_rows = b->nrows;
  /* One graph launch; nothing is waited for here.  An intra-only encoder never predicts from SELF, so
     its reconstruction stays on the device (ocg_backend_enc_copy_recon fetches it on demand).  An
     encoder that can emit inter frames gets the finished, padded frame back into its own buffer: the
     next frame's analysis reads it there wherever the loop still runs a C kernel. */
  if (ocg_dec_flush(b->ctx, &f, b->inter_capable ? st->ref_frame_handle + (size_t)self * (size_t)b->geom.ref_frame_sz : NULL,
                    b->inter_capable ? OCG_OUT_PADDED : OCG_OUT_NONE) < 0) {
    enc_fail(b, "ocg_dec_flush failed");
    return;
  }
  b->pending = 1;
  b->self_on_device = b->inter_capable ? -1 : self;
  pthread_mutex_lock(&g_estats_lock);
  g_estats.frames++;
  g_estats.me_repairs += b->n_me_repairs;
  g_estats.me_gold_refines += b->n_gold_refines;
  b->n_me_repairs = b->n_gold_refines = 0;
  g_estats.satd_lookups += b->n_satd_hit;
  g_estats.satd_host += b->n_satd_miss;
  g_estats.ssd_lookups += b->n_ssd_hit;
  g_estats.ssd_host += b->n_ssd_host;
  g_estats.intra_satd_lookups += b->n_isatd_hit;
  g_estats.fdct_quant_lookups += b->n_fq_hit;
  g_estats.fdct_quant_host += b->n_fq_miss;
  g_estats.coeff_rows += b->nrows;
  g_estats.h2d_bytes += (long)b->geom.nfrags * 16 + (long)b->nrows * 16;
  if (b->inter_capable) g_estats.d2h_bytes += (long)b->geom.ref_frame_sz;
  g_estats.flush_seconds += enc_now_s() - t0;
  pthread_mutex_unlock(&g_estats_lock);
}

/* ---- hooks --------------------------------------------------------------- */
static void ocge_enquant_table_fixup(void *_enquant[3][3][2], int _nqis) {
  /* _enquant is _enc->enquant (analyze.c:564) */
  oc_enc_ctx *enc = (oc_enc_ctx *)((char *)_enquant - offsetof(oc_enc_ctx, enquant));
  ocg_enc_backend *b = enc_backend_of(enc);
  oc_enc_enquant_table_fixup_c(_enquant, _nqis);
  t_enc = b;
  if (b != NULL) enc_begin_pass(b, _nqis);
}

/* the back-end of the pass in progress, if its tables / recorder are live */
static inline ocg_enc_backend *enc_live(void) {
  ocg_enc_backend *b = t_enc;
  return b != NULL && b->frame_open && !b->failed ? b : NULL;
}

static unsigned ocge_frag_intra_satd(int *_dc, const unsigned char *_src, int _ystride) {
  ocg_enc_backend *b = enc_live();
  if (b != NULL && b->tables) {
    ptrdiff_t fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
    if (fragi >= 0) {
      *_dc = b->tab.satd_dc[fragi];
      return b->tab.satd[fragi];
    }
    enc_fail(b, "frag_intra_satd: not a fragment of the input frame");
  } else if (b != NULL && b->inter_frame && b->itab_valid) {
    ptrdiff_t fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
    if (fragi >= 0) {
      b->n_isatd_hit++;
      *_dc = b->itab.intra_dc[fragi];
      return b->itab.intra_satd[fragi];
    }
  }
  return oc_enc_frag_intra_satd_c(_dc, _src, _ystride);
}

/* ---- inter-frame look-ups (tables of ocg_enc_inter_prepass) ----------------
   A call is served from the tables iff its source block is a fragment of the input frame and its
   predictor address(es) are the ones a table entry was computed with; anything else -- LAST/LAST2 vectors
   that coincide with no listed candidate, chroma vectors of a 4MV macro block with skipped luma blocks,
   the SSD of the block just reconstructed (analyze.c:829-835) -- runs the reference's C kernel on the host
   frames, which are kept current for exactly that. */
static inline ocg_enc_backend *enc_tables(void) {
  ocg_enc_backend *b = t_enc;
  return b != NULL && b->frame_open && !b->failed && b->inter_frame && b->itab_valid ? b : NULL;
}

static unsigned ocge_frag_satd(int *_dc, const unsigned char *_src, const unsigned char *_ref, int _ystride) {
  ocg_enc_backend *b = enc_tables();
  if (b != NULL) {
    ptrdiff_t fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
    const ptrdiff_t roff = _ref - b->pool0;
    if (fragi >= 0) {
      const ocg_enc_cand_rec *c = b->itab.cand + (size_t)fragi * (size_t)b->itab.ncand;
      int k;
      for (k = 0; k < b->itab.ncand; k++) {
        if (c[k].ref_off0 == roff && c[k].ref_off1 == INT32_MIN) {
          b->n_satd_hit++;
          *_dc = c[k].dc;
          return c[k].satd;
        }
      }
    }
    b->n_satd_miss++;
  }
  return oc_enc_frag_satd_c(_dc, _src, _ref, _ystride);
}

static unsigned ocge_frag_satd2(int *_dc, const unsigned char *_src, const unsigned char *_ref1, const unsigned char *_ref2,
                                int _ystride) {
  ocg_enc_backend *b = enc_tables();
  if (b != NULL) {
    ptrdiff_t fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
    const ptrdiff_t r1 = _ref1 - b->pool0, r2 = _ref2 - b->pool0;
    if (fragi >= 0) {
      const ocg_enc_cand_rec *c = b->itab.cand + (size_t)fragi * (size_t)b->itab.ncand;
      int k;
      for (k = 0; k < b->itab.ncand; k++, c++) {
        /* the two-tap average is symmetric in its taps */
        if ((c->ref_off0 == r1 && c->ref_off1 == r2) || (c->ref_off0 == r2 && c->ref_off1 == r1)) {
          b->n_satd_hit++;
          *_dc = c->dc;
          return c->satd;
        }
      }
    }
    b->n_satd_miss++;
  }
  return oc_enc_frag_satd2_c(_dc, _src, _ref1, _ref2, _ystride);
}

static unsigned ocge_frag_ssd(const unsigned char *_src, const unsigned char *_ref, int _ystride) {
  ocg_enc_backend *b = enc_tables();
  if (b != NULL) {
    /* oc_skip_cost (analyze.c:1996-2000, 2020-2024): the co-located block of the previous reconstruction */
    const unsigned char *io = b->enc->state.ref_frame_data[OC_FRAME_IO];
    if (_ref - b->enc->state.ref_frame_data[OC_FRAME_PREV] == _src - io) {
      ptrdiff_t fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
      if (fragi >= 0) { b->n_ssd_hit++; return b->itab.skip_ssd[fragi]; }
    }
    b->n_ssd_host++;
  }
  return oc_enc_frag_ssd_c(_src, _ref, _ystride);
}

static unsigned ocge_frag_border_ssd(const unsigned char *_src, const unsigned char *_ref, int _ystride, ogg_int64_t _mask) {
  ocg_enc_backend *b = enc_tables();
  if (b != NULL) {
    const unsigned char *io = b->enc->state.ref_frame_data[OC_FRAME_IO];
    if (_ref - b->enc->state.ref_frame_data[OC_FRAME_PREV] == _src - io) {
      ptrdiff_t fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
      if (fragi >= 0 && b->border_slot[fragi] >= 0 && b->border_mask[fragi] == _mask) {
        b->n_ssd_hit++;
        return b->itab.border_ssd[b->border_slot[fragi]];
      }
    }
    b->n_ssd_host++;
  }
  return oc_enc_frag_border_ssd_c(_src, _ref, _ystride, _mask);
}

static void ocge_frag_sub_128(ogg_int16_t _diff[64], const unsigned char *_src, int _ystride) {
  ocg_enc_backend *b = enc_live();
  if (b != NULL) {
    b->idct_pending = 0;
    b->cur_fragi = -1;
    b->fq_cur = -1;
    if (b->tables) {
      /* the residual itself stays on the device; its only consumer is fdct8x8 */
      b->cur_fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
      if (b->cur_fragi >= 0) return;
      enc_fail(b, "frag_sub_128: not a fragment of the input frame");
    }
  }
  if (b != NULL && b->inter_frame) __sync_fetch_and_add(&g_dbg_miss[9], 1);
  oc_enc_frag_sub_128_c(_diff, _src, _ystride);
}

/* oc_enc_frag_copy2 (analyze.c:741): the two-tap predictor is built in the block's place in SELF, then used
   as frag_sub's reference; remembered so that the subtraction can be recognised by its taps. */
static void ocge_frag_copy2(unsigned char *_dst, const unsigned char *_src1, const unsigned char *_src2, int _ystride) {
  ocg_enc_backend *b = enc_live();
  if (b != NULL) { b->c2_dst = _dst; b->c2_r1 = _src1; b->c2_r2 = _src2; }
  oc_enc_frag_copy2_c(_dst, _src1, _src2, _ystride);
}

static void ocge_frag_sub(ogg_int16_t _diff[64], const unsigned char *_src, const unsigned char *_ref, int _ystride) {
  ocg_enc_backend *b = enc_live();
  if (b != NULL) {
    b->idct_pending = 0;
    b->cur_fragi = -1;
    b->fq_cur = -1;
    if (b->fq_ok) {
      /* served from the speculative tables iff the predictor is one they were computed with */
      ptrdiff_t fragi = enc_fragi_of(b, _src, OC_FRAME_IO);
      if (fragi >= 0) {
        static const int SEL_CAND[OCG_ENC_FQ_NSEL] = {OCG_ENC_FQ_CAND0, OCG_ENC_FQ_CAND1, OCG_ENC_FQ_CAND2};
        ptrdiff_t r1, r2 = INT32_MIN;
        int sel;
        if (_ref == b->c2_dst) { r1 = b->c2_r1 - b->pool0; r2 = b->c2_r2 - b->pool0; }
        else r1 = _ref - b->pool0;
        for (sel = 0; sel < OCG_ENC_FQ_NSEL; sel++) {
          const ocg_enc_cand_rec *c = b->itab.cand + (size_t)fragi * (size_t)b->itab.ncand + SEL_CAND[sel];
          if ((c->ref_off0 == r1 && c->ref_off1 == r2) || (r2 != INT32_MIN && c->ref_off0 == r2 && c->ref_off1 == r1)) {
            const long at = (long)sel * b->geom.nfrags + (long)fragi;
            if (b->itab.fq_desc[at].off != 0xFFFFFFFFu) {
              b->fq_cur = at;
              b->fq_pli = b->st.recs[fragi].pli_qti & 3;
              b->n_fq_hit++;
              b->c2_dst = NULL;
              return; /* the residual's only consumer is fdct8x8 */
            }
          }
        }
      }
      b->n_fq_miss++;
      if (fragi >= 0 && enc_dbg_flag(0)) {
        ptrdiff_t r1, r2 = INT32_MIN;
        int k, hit = 8;
        if (_ref == b->c2_dst) { r1 = b->c2_r1 - b->pool0; r2 = b->c2_r2 - b->pool0; }
        else r1 = _ref - b->pool0;
        for (k = 0; k < 8 && hit == 8; k++) {
          const ocg_enc_cand_rec *c = b->itab.cand + (size_t)fragi * (size_t)b->itab.ncand + k;
          if ((c->ref_off0 == r1 && c->ref_off1 == r2) || (r2 != INT32_MIN && c->ref_off0 == r2 && c->ref_off1 == r1)) hit = k;
        }
        __sync_fetch_and_add(&g_dbg_miss[hit], 1);
      }
    }
    b->c2_dst = NULL;
  }
  oc_enc_frag_sub_c(_diff, _src, _ref, _ystride);
}

static void ocge_fdct8x8(ogg_int16_t _y[64], const ogg_int16_t _x[64]) {
  ocg_enc_backend *b = enc_live();
  if (b != NULL && b->fq_cur >= 0) {
    const ocg_enc_fq_desc *d = b->itab.fq_desc + b->fq_cur;
    const int n8 = (d->count + 7) & ~7;
    memcpy(_y, b->itab.fq_pool + (size_t)d->off * 8, (size_t)n8 * sizeof(ogg_int16_t));
    memset(_y + n8, 0, (size_t)(64 - n8) * sizeof(ogg_int16_t));
    return;
  }
  if (b != NULL && b->tables && b->cur_fragi >= 0) {
    memcpy(_y, b->tab.dct + (size_t)b->cur_fragi * 64, 64 * sizeof(ogg_int16_t));
    return;
  }
  oc_enc_fdct8x8_c(_y, _x);
}

static int ocge_quantize(ogg_int16_t _qdct[64], const ogg_int16_t _dct[64], const ogg_uint16_t _dequant[64],
                         const void *_enquant) {
  ocg_enc_backend *b = enc_live();
  if (b != NULL && b->fq_cur >= 0) {
    const ocg_enc_fq_desc *d = b->itab.fq_desc + b->fq_cur;
    const int n8 = (d->count + 7) & ~7;
    int qii;
    for (qii = 0; qii < b->fq_nq && b->enc->enquant[b->fq_pli][qii][1] != _enquant; qii++) {}
    if (qii < b->fq_nq) {
      memcpy(_qdct, b->itab.fq_pool + (size_t)d->off * 8 + (size_t)(1 + qii) * n8, (size_t)n8 * sizeof(ogg_int16_t));
      memset(_qdct + n8, 0, (size_t)(64 - n8) * sizeof(ogg_int16_t));
      return d->nz[qii];
    }
    /* _dct came from the table, so the C quantiser below is still exact */
  }
  if (b != NULL && b->tables && b->cur_fragi >= 0) {
    oc_enc_ctx *enc = b->enc;
    int pli = b->st.recs[b->cur_fragi].pli_qti & 3, qii;
    for (qii = 0; qii < b->nqis && enc->enquant[pli][qii][0] != _enquant; qii++) {}
    if (qii < b->nqis) {
      size_t at = (size_t)qii * (size_t)b->geom.nfrags + (size_t)b->cur_fragi;
      memcpy(_qdct, b->tab.qdct + at * 64, 64 * sizeof(ogg_int16_t));
      return b->tab.nonzero[at];
    }
    /* _dct came from the table, so the C quantiser below is still exact */
  }
  return oc_enc_quantize_c(_qdct, _dct, _dequant, _enquant);
}

static inline int enc_row_nonzero(const ogg_int16_t *row) {
  ogg_uint64_t a, c;
  memcpy(&a, row, 8);
  memcpy(&c, row + 4, 8);
  return (a | c) != 0;
}

/* oc_idct8x8 (state.h:98, called at analyze.c:806 with the dequantised coefficients the tokeniser left
   in _x): take the rows the transform of this footprint reads (idct.c:327-329) for the device, and
   leave _x zeroed (idct.c:245,276,295).  In an inter frame the analysis loop goes on to measure the
   reconstructed block (analyze.c:825-868), so the residual is also produced here, by the reference's C
   transform. */
static void ocge_idct8x8(ogg_int16_t _y[64], ogg_int16_t _x[64], int _last_zzi) {
  ocg_enc_backend *b = enc_live();
  int nr = _last_zzi <= 3 ? 2 : (_last_zzi <= 10 ? 4 : 8);
  int r, mask = 0;
  if (b == NULL) { oc_idct8x8_c(_y, _x, _last_zzi); return; }
  b->pend_dc = _x[0];
  b->pend_row0 = (ogg_uint32_t)b->nrows;
  for (r = 0; r < nr; r++) {
    const ogg_int16_t *row = _x + r * 8;
    if (enc_row_nonzero(row)) {
      ogg_int16_t *out = b->st.coeff_rows + (size_t)b->nrows * 8;
      memcpy(out, row, 16);
      if (r == 0) out[0] = 0; /* DC travels in the record */
      b->nrows++;
      mask |= 1 << r;
    }
  }
  b->pend_mask = mask;
  /* never the DC-only shortcut of state.c:967: analyze.c:806 always transforms */
  b->pend_last_zzi = _last_zzi < 2 ? 2 : _last_zzi;
  b->idct_pending = 1;
  if (b->inter_frame) oc_idct8x8_c(_y, _x, _last_zzi);
  else memset(_x, 0, (size_t)nr * 16);
}

/* Shared by frag_recon_intra / frag_recon_inter: the fragment's record for the device. */
static void enc_record(ocg_enc_backend *b, ptrdiff_t fragi, int refi, int mv, const ogg_int16_t _residue[64]) {
  ocg_frag_rec *rec = b->st.recs + fragi;
  int pli = rec->pli_qti & 3;
  rec->mv = (ogg_int16_t)mv;
  rec->refi = (unsigned char)refi;
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

static void ocge_frag_recon_intra(unsigned char *_dst, int _ystride, const ogg_int16_t _residue[64]) {
  ocg_enc_backend *b = enc_live();
  if (b != NULL) {
    ptrdiff_t fragi = enc_fragi_of(b, _dst, OC_FRAME_SELF);
    if (fragi >= 0) enc_record(b, fragi, OC_FRAME_SELF, 0, _residue);
    else enc_fail(b, "frag_recon_intra: not a fragment of the frame being reconstructed");
    if (b->frame_open && !b->inter_frame) return; /* nobody reads an intra frame's pixels on the host */
  }
  oc_frag_recon_intra_c(_dst, _ystride, _residue);
}

static void ocge_frag_recon_inter(unsigned char *_dst, const unsigned char *_src, int _ystride,
                                  const ogg_int16_t _residue[64]) {
  ocg_enc_backend *b = enc_live();
  if (b != NULL) {
    ptrdiff_t fragi = enc_fragi_of(b, _dst, OC_FRAME_SELF);
    if (fragi >= 0) {
      /* the predictor is a function of the reference frame and the vector (oc_state_get_mv_offsets,
         state.c:846-957, restated on the device): analyze.c:710-746 */
      const oc_fragment *frag = b->enc->state.frags + fragi;
      int mode = frag->mb_mode;
      int mv = mode == OC_MODE_INTER_NOMV || mode == OC_MODE_GOLDEN_NOMV ? 0 : b->enc->state.frag_mvs[fragi];
      enc_record(b, fragi, frag->refi, mv, _residue);
    } else enc_fail(b, "frag_recon_inter: not a fragment of the frame being reconstructed");
  }
  /* analyze.c:829-835 measures this block next */
  oc_frag_recon_inter_c(_dst, _src, _ystride, _residue);
}

static void ocge_frag_copy_list(unsigned char *_dst_frame, const unsigned char *_src_frame, int _ystride,
                                const ptrdiff_t *_fragis, ptrdiff_t _nfragis, const ptrdiff_t *_frag_buf_offs) {
  /* analyze.c:610-622: the MCU's uncoded fragments; the device copies them PREV -> SELF at the flush */
  ocg_enc_backend *b = enc_live();
  ptrdiff_t i;
  if (b == NULL) { oc_frag_copy_list_c(_dst_frame, _src_frame, _ystride, _fragis, _nfragis, _frag_buf_offs); return; }
  for (i = 0; i < _nfragis; i++) b->st.recs[_fragis[i]].refi = OCG_FRAG_UNCODED;
}

static void ocge_state_loop_filter_frag_rows(const oc_theora_state *_state, signed char _bv[256], int _refi, int _pli,
                                             int _fragy0, int _fragy_end) {
  /* filtered on the device over the whole frame at flush */
  ocg_enc_backend *b = t_enc;
  if (b != NULL && !b->failed) return;
  oc_state_loop_filter_frag_rows_c(_state, _bv, _refi, _pli, _fragy0, _fragy_end);
}

static void ocge_restore_fpu(void) {
  ocg_enc_backend *b = t_enc;
  if (b != NULL && b->frame_open && !b->failed) enc_flush(b);
}

/* ---- motion analysis (mcenc.c:517-548, 666-672, 763-791) -------------------
   Not vtable entries in the reference: analyze.c calls them by name, so the integrated build renames
   the reference's definitions (theora_b200/backend/Makefile) and these take their place.  They hand out
   the device's results of enc_me_prepass; an encoder without them (tooling mode, speed levels the
   device analysis does not cover) runs the reference's own functions. */
static inline int ocge_div2(int v) { return v / 2; } /* OC_DIV2: towards zero */
static inline int ocge_clamp31(int v) { return v < -31 ? -31 : (v > 31 ? 31 : v); }
static inline int ocge_mv_x(int mv) { return (signed char)mv; }
static inline int ocge_mv_y(int mv) { return (ogg_int16_t)mv >> 8; }
static inline int ocge_mv(int x, int y) { return (ogg_int16_t)((x & 0xFF) | y * 256); }
static inline int ocge_mv_sub(int a, int c) { return ocge_mv(ocge_mv_x(a) - ocge_mv_x(c), ocge_mv_y(a) - ocge_mv_y(c)); }
static inline int ocge_med3(int a, int c, int d) {
  int lo = a < c ? a : c, hi = a < c ? c : a;
  return d < lo ? lo : (d > hi ? hi : d);
}

/* The candidate set of a GOLD search (mcenc.c:90-164) and its threshold base (331-337), given the
   neighbours' vectors/errors: half-pel entries as ocg_mb_search_in wants them, [0] = median of [1..3]. */
static void ocge_gold_cands(ocg_mb_search_in *in, int ncn, const int nb_mv[4], const unsigned nb_err[4], int accum,
                            int m1, int m2, unsigned own_err) {
  int n = 1, i;
  unsigned t2 = own_err;
  for (i = 0; i < ncn; i++, n++) {
    in->cand[n][0] = (signed char)ocge_mv_x(nb_mv[i]);
    in->cand[n][1] = (signed char)ocge_mv_y(nb_mv[i]);
  }
  in->cand[n][0] = (signed char)ocge_mv_x(accum);
  in->cand[n][1] = (signed char)ocge_mv_y(accum);
  n++;
  in->cand[n][0] = (signed char)ocge_clamp31(ocge_mv_x(m1) + ocge_mv_x(accum));
  in->cand[n][1] = (signed char)ocge_clamp31(ocge_mv_y(m1) + ocge_mv_y(accum));
  n++;
  in->cand[n][0] = in->cand[n][1] = 0;
  n++;
  in->cand[0][0] = (signed char)ocge_med3(in->cand[1][0], in->cand[2][0], in->cand[3][0]);
  in->cand[0][1] = (signed char)ocge_med3(in->cand[1][1], in->cand[2][1], in->cand[3][1]);
  in->setb0 = (unsigned char)n;
  in->cand[n][0] = (signed char)ocge_clamp31(2 * ocge_mv_x(m1) - ocge_mv_x(m2) + ocge_mv_x(accum));
  in->cand[n][1] = (signed char)ocge_clamp31(2 * ocge_mv_y(m1) - ocge_mv_y(m2) + ocge_mv_y(accum));
  n++;
  in->ncand = (unsigned char)n;
  for (i = 0; i < (ncn < 3 ? ncn : 3); i++) if (nb_err[i] > t2) t2 = nb_err[i];
  in->t2_base = (ogg_uint16_t)t2;
  in->is_prev = 0;
}

static ocg_enc_backend *enc_me_live(oc_enc_ctx *_enc) {
  ocg_enc_backend *b = t_enc;
  if (b == NULL || b->enc != _enc) b = enc_backend_of(_enc);
  return b != NULL && b->me_valid && !b->failed ? b : NULL;
}

void oc_mcenc_search(oc_enc_ctx *_enc, int _mbi) {
  ocg_enc_backend *b = enc_me_live(_enc);
  oc_mb_enc_info *e;
  const ocg_me_mb *m;
  int k, ncn, stale = 0;
  int gold_mv, gold_satd;
  unsigned gold_err;
  if (b == NULL) { oc_refimpl_mcenc_search(_enc, _mbi); return; }
  e = _enc->mb_info + _mbi;
  m = b->me_tab + _mbi;
  ncn = e->ncneighbors;
  gold_mv = m->unref_mv[OC_FRAME_GOLD];
  gold_err = m->error[OC_FRAME_GOLD];
  gold_satd = (int)m->unref_satd[OC_FRAME_GOLD];
  for (k = 0; k < ncn; k++) stale |= b->gold_dirty[e->cneighbors[k]];
  if (stale) {
    /* did the speculation feed this search what the reference would? compare what the search consumes:
       the full-pel candidate list and the threshold base */
    ocg_mb_search_in spec, real;
    int nb_spec[4], nb_real[4], same;
    unsigned err_spec[4], err_real[4];
    /* this macro block's own history as oc_mcenc_search rotates it before the GOLD search (mcenc.c:526-541) */
    const int old0 = e->analysis_mv[0][OC_FRAME_GOLD], old1 = e->analysis_mv[1][OC_FRAME_GOLD];
    const int accum = e->analysis_mv[2][OC_FRAME_GOLD];
    const int m2 = ocge_mv_sub(old1, accum), m1 = ocge_mv_sub(old0, old1);
    for (k = 0; k < ncn; k++) {
      const unsigned n = e->cneighbors[k];
      nb_spec[k] = b->me_tab[n].analysis_mv[0][OC_FRAME_GOLD];
      err_spec[k] = b->me_tab[n].error[OC_FRAME_GOLD];
      nb_real[k] = _enc->mb_info[n].analysis_mv[0][OC_FRAME_GOLD];
      err_real[k] = _enc->mb_info[n].error[OC_FRAME_GOLD];
    }
    memset(&spec, 0, sizeof(spec));
    memset(&real, 0, sizeof(real));
    ocge_gold_cands(&spec, ncn, nb_spec, err_spec, accum, m1, m2, e->error[OC_FRAME_GOLD]);
    ocge_gold_cands(&real, ncn, nb_real, err_real, accum, m1, m2, e->error[OC_FRAME_GOLD]);
    same = spec.t2_base == real.t2_base;
    for (k = 0; k < spec.ncand && same; k++)
      same = ocge_div2(spec.cand[k][0]) == ocge_div2(real.cand[k][0]) && ocge_div2(spec.cand[k][1]) == ocge_div2(real.cand[k][1]);
    if (!same) {
      ocg_mb_search_out so;
      ocg_mb_refine_out ro;
      for (k = 0; k < 4; k++) real.frag_off[k] = (ogg_int32_t)_enc->state.frag_buf_offs[_enc->state.mb_maps[_mbi][0][k]];
      if (ocg_me_repair(b->me, OC_FRAME_GOLD, &real, 1, &so, &ro) < 0) {
        enc_fail(b, "ocg_me_repair failed");
        oc_refimpl_mcenc_search(_enc, _mbi);
        return;
      }
      gold_mv = ocge_mv(so.best_vec[0] * 2, so.best_vec[1] * 2);
      gold_err = so.error;
      gold_satd = (int)so.satd;
      b->gold_fixed[_mbi] = 1;
      b->gold_fix[_mbi].mv = (ogg_int16_t)ocge_mv(ro.mv[0], ro.mv[1]);
      b->gold_fix[_mbi].satd = ro.satd;
      /* later macro blocks were fed the speculative values of this one */
      if (gold_mv != m->analysis_mv[0][OC_FRAME_GOLD] || gold_err != m->error[OC_FRAME_GOLD]) b->gold_dirty[_mbi] = 1;
      b->n_me_repairs++;
    }
  }
  /* history as the reference leaves it (mcenc.c:534, 546-547): a function of this macro block alone */
  for (k = 0; k < 2; k++) {
    e->analysis_mv[1][k] = m->analysis_mv[1][k];
    e->analysis_mv[2][k] = m->analysis_mv[2][k];
  }
  e->analysis_mv[0][OC_FRAME_PREV] = m->unref_mv[OC_FRAME_PREV];
  e->error[OC_FRAME_PREV] = m->error[OC_FRAME_PREV];
  e->satd[OC_FRAME_PREV] = m->unref_satd[OC_FRAME_PREV];
  if (!(b->me_flags & OCG_ME_FAST)) {
    for (k = 0; k < 4; k++) { e->block_mv[k] = m->block_mv[k]; e->block_satd[k] = m->block_satd[k]; }
  }
  e->analysis_mv[0][OC_FRAME_GOLD] = (oc_mv)gold_mv;
  e->error[OC_FRAME_GOLD] = (ogg_uint16_t)gold_err;
  e->satd[OC_FRAME_GOLD] = (unsigned)gold_satd;
}

void oc_mcenc_refine1mv(oc_enc_ctx *_enc, int _mbi, int _frame) {
  ocg_enc_backend *b = enc_me_live(_enc);
  oc_mb_enc_info *e;
  const ocg_me_mb *m;
  if (b == NULL || !(b->me_flags & OCG_ME_REFINE_PREV)) { oc_refimpl_mcenc_refine1mv(_enc, _mbi, _frame); return; }
  e = _enc->mb_info + _mbi;
  m = b->me_tab + _mbi;
  if (_frame == OC_FRAME_PREV) {
    e->analysis_mv[0][OC_FRAME_PREV] = m->analysis_mv[0][OC_FRAME_PREV];
    e->satd[OC_FRAME_PREV] = m->satd[OC_FRAME_PREV];
  } else {
    const int before = e->analysis_mv[0][OC_FRAME_GOLD];
    if (b->gold_fixed[_mbi]) {
      e->analysis_mv[0][OC_FRAME_GOLD] = b->gold_fix[_mbi].mv;
      e->satd[OC_FRAME_GOLD] = b->gold_fix[_mbi].satd;
    } else {
      e->analysis_mv[0][OC_FRAME_GOLD] = m->gold_ref_mv;
      e->satd[OC_FRAME_GOLD] = m->gold_ref_satd;
    }
    /* the device's chain handed the unrefined vector to this macro block's later neighbours */
    if (e->analysis_mv[0][OC_FRAME_GOLD] != before) b->gold_dirty[_mbi] = 1;
    b->n_gold_refines++;
  }
}

void oc_mcenc_refine4mv(oc_enc_ctx *_enc, int _mbi) {
  ocg_enc_backend *b = enc_me_live(_enc);
  oc_mb_enc_info *e;
  const ocg_me_mb *m;
  int k;
  if (b == NULL || !(b->me_flags & OCG_ME_REFINE_4MV)) { oc_refimpl_mcenc_refine4mv(_enc, _mbi); return; }
  e = _enc->mb_info + _mbi;
  m = b->me_tab + _mbi;
  for (k = 0; k < 4; k++) {
    e->ref_mv[k] = m->ref_mv[k];
    e->block_satd[k] = m->ref_block_satd[k];
  }
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
    if (b->ei != NULL) ocg_enc_inter_destroy(b->ei);
    if (b->me != NULL) ocg_me_destroy(b->me);
    if (b->me_tab != NULL) ocg_host_unregister(b->me_tab);
    if (b->pinned) ocg_host_unregister(b->enc->state.ref_frame_handle);
    ocg_ctx_destroy(b->ctx);
  }
  free(b->me_tab);
  free(b->gold_dirty);
  free(b->gold_fixed);
  free(b->gold_fix);
  free(b->border_slot);
  free(b->border_mask);
  free(b->off2frag);
  free(b);
}

void oc_enc_accel_init_ocg(oc_enc_ctx *_enc) {
  oc_theora_state *st = &_enc->state;
  ocg_enc_backend *b;
  ptrdiff_t fragi, omin, omax;
  oc_enc_accel_init_c(_enc);
  t_enc_init_failed = 0;
  if (g_enc_mode == OCG_ENC_HOST) {
    /* tooling mode: the reference's C kernels for everything; the spy wraps the C fix-up */
    if (g_enc_spy != NULL) _enc->opt_vtable.enquant_table_fixup = ocge_spy_fixup;
    return;
  }
  t_enc_init_failed = 1;
  b = (ocg_enc_backend *)calloc(1, sizeof(*b));
  if (b == NULL) return;
  b->enc = _enc;
  b->self_on_device = -1;
  b->inter_capable = st->info.keyframe_granule_shift != 0;
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
    return; /* th_encode_alloc (below) reports the failure: there is no CPU fallback encoder */
  }
  /* an encoder flushes a few times per second: kernel-by-kernel launches, no graph to instantiate */
  ocg_ctx_set_flush_graph(b->ctx, -1);
  b->pinned = ocg_host_register(st->ref_frame_handle, (size_t)b->geom.ref_frame_sz * 6) == 0;
  if (b->inter_capable && !b->pinned) {
    /* the flush graph copies the finished frame straight into the encoder's own buffer */
    fprintf(stderr, "theora_b200 encoder back-end: cannot page-lock the frame buffers (%s)\n", ocg_last_error());
    ocg_ctx_destroy(b->ctx);
    free(b->off2frag);
    free(b);
    return;
  }
  /* pre-pass look-ups (intra frames) */
  _enc->opt_vtable.enquant_table_fixup = ocge_enquant_table_fixup;
  _enc->opt_vtable.frag_intra_satd = ocge_frag_intra_satd;
  _enc->opt_vtable.frag_sub_128 = ocge_frag_sub_128;
  _enc->opt_vtable.frag_sub = ocge_frag_sub;
  if (b->inter_capable) {
    /* inter-frame look-ups */
    _enc->opt_vtable.frag_satd = ocge_frag_satd;
    _enc->opt_vtable.frag_satd2 = ocge_frag_satd2;
    _enc->opt_vtable.frag_ssd = ocge_frag_ssd;
    _enc->opt_vtable.frag_border_ssd = ocge_frag_border_ssd;
    _enc->opt_vtable.frag_copy2 = ocge_frag_copy2;
  }
  _enc->opt_vtable.fdct8x8 = ocge_fdct8x8;
  _enc->opt_vtable.quantize = ocge_quantize;
  /* recorded reconstruction (every frame) */
  _enc->opt_vtable.frag_recon_intra = ocge_frag_recon_intra;
  _enc->opt_vtable.frag_recon_inter = ocge_frag_recon_inter;
  st->opt_vtable.idct8x8 = ocge_idct8x8;
  st->opt_vtable.frag_copy_list = ocge_frag_copy_list;
  st->opt_vtable.state_loop_filter_frag_rows = ocge_state_loop_filter_frag_rows;
  st->opt_vtable.restore_fpu = ocge_restore_fpu;
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
    /* no usable device: fail the allocation rather than encode on the CPU */
    oc_refimpl_encode_free(enc);
    return NULL;
  }
  return enc;
}

void th_encode_free(th_enc_ctx *_enc) {
  if (_enc != NULL) enc_backend_destroy(enc_backend_of(_enc));
  oc_refimpl_encode_free(_enc);
}

int th_encode_ycbcr_in(th_enc_ctx *_enc, th_ycbcr_buffer _img) {
  ocg_enc_backend *b = _enc != NULL ? enc_backend_of(_enc) : NULL;
  int ret;
  if (b != NULL && b->failed) return TH_EFAULT;
  ret = oc_refimpl_encode_ycbcr_in(_enc, _img);
  if (b != NULL && b->failed) return TH_EFAULT;
  return ret;
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
    if (b != NULL) enc_wait(b);
    if (b != NULL && b->failed) return TH_EFAULT;
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
