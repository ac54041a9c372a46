/* ocg_backend.c -- vtable back-end that puts libtheora's decoder block pipeline
 * on a B200 through the C ABI of include/theora_b200.h.
 *
 * Compiled with `-include ocg_hooks.h` against the reference's private headers,
 * exactly like lib/x86/x86state.c or lib/arm/armstate.c.  The reference host
 * code (decode.c, state.c, huffdec.c, bitpack.c ...) is consumed unmodified.
 *
 * The per-block hooks cannot launch kernels (call granularity),
 * so they RECORD:
 *   dc_unpredict_mcu_plane  (decint.h:72)  host DC un-prediction, frame begin
 *   state_frag_recon        (state.h:361)  -> ocg_frag_rec + coefficient rows
 *   frag_copy_list          (state.h:356)  -> uncoded offsets
 *   state_loop_filter_frag_rows            -> nothing (whole-frame on device)
 *   restore_fpu             (state.h:369)  end of frame (decode.c:2965): FLUSH =
 *                           one H2D of the lists, recon/copy + loop filter +
 *                           borders kernels, D2H of the frame into the host
 *                           reference buffer th_decode_ycbcr_out hands out.
 * th_decode_alloc / th_decode_free / th_decode_ctl are thin wrappers around
 * the reference's own functions (renamed at compile time) so device state
 * follows the decoder's life time and the stripe callback is delivered once per
 * frame after the flush.  Post-processing (TH_DECCTL_SET_PPLEVEL > 0), which the
 * reference runs inside the MCU loop on host pixels that are not there yet
 * (decode.c:2899-2914), is re-run over the whole frame after the flush by the
 * reference's own filters (ocg_dec_host.c).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include "decint.h"
#include "encint.h"
#include "ocg_backend.h"

/* the reference's own entry points, renamed on decode.c's command line */
th_dec_ctx *oc_refimpl_decode_alloc(const th_info *_info, const th_setup_info *_setup);
void oc_refimpl_decode_free(th_dec_ctx *_dec);
int oc_refimpl_decode_ctl(th_dec_ctx *_dec, int _req, void *_buf, size_t _buf_sz);
int oc_refimpl_decode_packetin(th_dec_ctx *_dec, const ogg_packet *_op, ogg_int64_t *_granpos);
int oc_refimpl_decode_ycbcr_out(th_dec_ctx *_dec, th_ycbcr_buffer _ycbcr);
void oc_state_accel_init_ocg(oc_theora_state *_state);
void ocg_host_dc_unpredict_mcu_plane(oc_dec_ctx *_dec, oc_dec_pipeline_state *_pipe, int _pli); /* ocg_dec_host.c */

typedef struct ocg_backend {
  th_dec_ctx        *dec;
  ocg_ctx           *ctx;      /* NULL in record mode */
  ocg_geometry       geom;
  ocg_staging        st;
  void              *heap_staging; /* record mode only */
  ocg_frag_rec      *heap_tmpl;    /* record mode only */
  int                mode;
  int                frame_open;
  int                ncoded;
  int                nrows;
  int                ref_idx[3];
  ogg_uint16_t       dcq[3][2];
  unsigned char      dev_valid[6];
  int                pinned;
  int                dc_device;   /* DC prediction is undone on the device (records carry residuals) */
  int                expand;      /* the device expands the tokens (ocg_dec_flush_tokens): the hooks record nothing */
  int                regs;        /* bit i: host array i (frags, frag_mvs, dct_tokens) is page-locked by us */
  int                dc_ahead;    /* ocg_dec_dc_begin was called for the frame being assembled */
  int                dc_ahead_used;
  int                pending;     /* a flushed frame's kernels / copy-back may still be running */
  int                failed;      /* a device call failed: the decoder is unusable, the API returns TH_EFAULT */
  int                out_mode;
  ocg_pp            *pp;          /* post-processing on the device (made when a frame first asks for it) */
  int                pp_pending;  /* the last flushed frame's post-processed planes are still on the device */
  int                pp_level;    /* the frame's post-processing level (taken from the pipeline at the first hook) */
  unsigned char     *pp_qis;      /* [nfrags] state.qis[frag.qii] of the frame being flushed */
  th_stripe_callback user_cb;
  ocg_backend_stats  stats;       /* this decoder's share; summed by ocg_backend_get_stats (no shared lock per frame) */
  struct ocg_backend *next;
} ocg_backend;

static void backend_unregister(ocg_backend *b);
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static ocg_backend *g_list;
static int g_mode = OCG_BACKEND_GPU;
static int g_dc_mode = OCG_DC_HOST;
static int g_expand_mode = OCG_EXPAND_DEVICE;
static int g_out_mode = OCG_OUT_PICTURE;
static ocg_capture_fn g_capture;
static void *g_capture_user;
static int g_device; /* one process per GPU: process-wide */
static __thread ocg_backend *t_cur;
static ocg_backend_stats g_stats_retired; /* of decoders that no longer exist; under g_lock */

OCG_API void ocg_backend_set_mode(int mode) { g_mode = mode; }
OCG_API void ocg_backend_set_device(int device) { g_device = device; }
OCG_API void ocg_backend_set_dc_mode(int mode) { g_dc_mode = mode; }
OCG_API void ocg_backend_set_expand_mode(int mode) { g_expand_mode = mode; }
OCG_API void ocg_backend_set_output_mode(int mode) { g_out_mode = mode; }
OCG_API void ocg_backend_set_capture(ocg_capture_fn fn, void *user) { g_capture = fn; g_capture_user = user; }
static void stats_sum(ocg_backend_stats *a, const ocg_backend_stats *b) {
  a->frames += b->frames;
  a->coded_frags += b->coded_frags;
  a->uncoded_frags += b->uncoded_frags;
  a->coeff_rows += b->coeff_rows;
  a->h2d_bytes += b->h2d_bytes;
  a->d2h_bytes += b->d2h_bytes;
  a->flush_seconds += b->flush_seconds;
  a->wait_seconds += b->wait_seconds;
}

/* Every decoder counts for itself (its own thread, no lock on the per-frame path); the totals are put
   together here.  Call it while no decoder is inside th_decode_packetin for exact figures. */
OCG_API void ocg_backend_get_stats(ocg_backend_stats *out, int reset) {
  ocg_backend *b;
  ocg_backend_stats t;
  pthread_mutex_lock(&g_lock);
  t = g_stats_retired;
  for (b = g_list; b != NULL; b = b->next) stats_sum(&t, &b->stats);
  if (reset) {
    memset(&g_stats_retired, 0, sizeof(g_stats_retired));
    for (b = g_list; b != NULL; b = b->next) memset(&b->stats, 0, sizeof(b->stats));
  }
  pthread_mutex_unlock(&g_lock);
  if (out) *out = t;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static ocg_backend *backend_of(const void *dec) {
  ocg_backend *b = t_cur;
  if (b != NULL && (const void *)b->dec == dec) return b;
  pthread_mutex_lock(&g_lock);
  for (b = g_list; b != NULL && (const void *)b->dec != dec; b = b->next) {}
  pthread_mutex_unlock(&g_lock);
  t_cur = b;
  return b;
}

/* The hooks cannot report errors (state.h:352-370: all void).  A failed device call is latched: the rest
   of the frame's hooks become no-ops and the th_decode_* wrappers below return TH_EFAULT from then on
   (the decoder's reference frames are no longer valid).  OCG_BACKEND_ABORT_ON_ERROR keeps the old
   behaviour for debugging. */
static void backend_fail(ocg_backend *b, const char *what) {
  if (b == NULL || !b->failed) fprintf(stderr, "theora_b200 back-end: %s (%s)\n", what, ocg_last_error());
#if defined(OCG_BACKEND_ABORT_ON_ERROR)
  abort();
#endif
  if (b != NULL) {
    b->failed = 1;
    b->frame_open = 0;
    b->pending = 0;
  }
}

/* Waits for the last flushed frame (kernels + copy-back). */
static void backend_wait(ocg_backend *b) {
  if (b->pending && b->ctx != NULL) {
    double t0 = now_s();
    b->pending = 0;
    if ((b->dc_ahead_used ? ocg_ctx_sync(b->ctx) : ocg_dec_wait(b->ctx)) < 0) { backend_fail(b, "waiting for the frame failed"); return; }
    if (b->pp_pending) {
      /* the planes the filters produced go where the reference's own filters would have left them
         (pp_frame_data: luma first, the chroma planes behind it; decode.c:1283-1315) */
      b->pp_pending = 0;
      if (b->dec->pp_frame_data == NULL || ocg_pp_download(b->pp, b->dec->pp_frame_data) < 0) {
        backend_fail(b, "fetching the post-processed frame failed");
        return;
      }
    }
    b->stats.wait_seconds += now_s() - t0;
  }
}

static void stats_add(ocg_backend *b, long h2d, long d2h, double secs) {
  b->stats.frames++;
  b->stats.coded_frags += b->ncoded;
  b->stats.uncoded_frags += b->geom.nfrags - b->ncoded;
  b->stats.coeff_rows += b->nrows;
  b->stats.h2d_bytes += h2d;
  b->stats.d2h_bytes += d2h;
  b->stats.flush_seconds += secs;
}

/* ---- frame life cycle ---------------------------------------------------- */
static void backend_begin_frame(ocg_backend *b) {
  const oc_theora_state *st = &b->dec->state;
  int i;
  /* staging records keep buf_off/plane from context creation; every fragment
     is visited once per frame by exactly one of the recon and copy-list hooks
     (decode.c:1584,1601), which refresh the rest */
  if (b->ctx != NULL && !b->expand && ocg_dec_staging(b->ctx, &b->st) < 0) { backend_fail(b, "ocg_dec_staging failed"); return; }
  b->ncoded = b->nrows = 0;
  /* decode.c:2790-2794 has already picked SELF; GOLD/PREV are still the
     references this frame predicts from (they rotate at 2947-2962). */
  for (i = 0; i < 3; i++) b->ref_idx[i] = st->ref_frame_idx[i];
  memset(b->dcq, 0, sizeof(b->dcq));
  b->frame_open = 1;
  b->dc_ahead = 0;
  /* The reference would run its de-blocking / de-ringing filters inside the MCU loop (decode.c:2899-2914) on
     host pixels that are not there: the level is taken over here -- this is the frame's first hook, ahead of
     the loop's first look at it -- and the filters run on the device at the flush instead.  pp_frame_buf
     (decode.c:1283-1322) already points where th_decode_ycbcr_out will look. */
  b->pp_level = 0;
  if (b->ctx != NULL) {
    b->pp_level = b->dec->pipe.pp_level;
    b->dec->pipe.pp_level = 0; /* OC_PP_LEVEL_DISABLED */
  }
  if (b->dc_device == 2 && b->ctx != NULL) {
    /* Every input of the DC recurrence (coded flags, reference types, residuals) is in frags[] once the
       tokens are unpacked (decode.c:2822), i.e. now: the device starts on it while the host expands
       the coefficients of the whole frame. */
    if (ocg_dec_dc_begin(b->ctx, (const ogg_uint32_t *)st->frags) < 0) { backend_fail(b, "ocg_dec_dc_begin failed"); return; }
    b->dc_ahead = 1;
  }
}

static void backend_flush(ocg_backend *b) {
  oc_theora_state *st = &b->dec->state;
  ocg_dec_frame f;
  double t0 = now_s();
  int i, k;
  memset(&f, 0, sizeof(f));
  for (i = 0; i < 3; i++) f.ref_idx[i] = b->ref_idx[i];
  f.lf_limit = st->loop_filter_limits[st->qis[0]];
  for (i = 0; i < 3; i++) for (k = 0; k < 2; k++) f.dc_quant[i][k] = b->dcq[i][k];
  f.ncoded = b->ncoded;
  f.intra_frame = st->frame_type == OC_INTRA_FRAME;
  f.ncoeff_rows = b->nrows;
  f.dc_residual = b->dc_device ? (b->dc_ahead ? 2 : 1) : 0;
  b->dc_ahead_used = b->dc_ahead;
  b->dc_ahead = 0;
  b->frame_open = 0;
  if (g_capture != NULL && !b->expand) (*g_capture)(g_capture_user, &f, &b->st);
  if (b->ctx == NULL) { stats_add(b, 0, 0, 0.0); return; } /* record mode */
  {
    const int self = f.ref_idx[OCG_FRAME_SELF];
    unsigned char *host_self = st->ref_frame_handle + (size_t)self * (size_t)b->geom.ref_frame_sz;
    long extra_h2d = 0, d2h;
    int r;
    /* with luma AND chroma post-processing on, th_decode_ycbcr_out hands out the post-processed planes only
       (decode.c:1299-1315): the reconstruction itself need not leave the device */
    const int out_mode = b->pp_level >= 5 && b->out_mode == OCG_OUT_PICTURE ? OCG_OUT_NONE : b->out_mode;
    /* A reference the device has never produced (stream starting on an inter
       frame: oc_dec_init_dummy_frame, decode.c:2053) is taken from the host. */
    if (st->frame_type != OC_INTRA_FRAME) {
      for (i = 0; i < 2; i++) {
        int ri = f.ref_idx[i];
        if (ri >= 0 && !b->dev_valid[ri]) {
          if (ocg_ctx_upload_frame(b->ctx, ri, st->ref_frame_handle + (size_t)ri * (size_t)b->geom.ref_frame_sz) < 0) {
            backend_fail(b, "reference upload failed");
            return;
          }
          b->dev_valid[ri] = 1;
          extra_h2d += (long)b->geom.ref_frame_sz;
        }
      }
    }
    /* Everything below is queued on the context's stream and NOT waited for here: the frame is only
       needed in host memory when the application asks for it (th_decode_ycbcr_out), a stripe callback
       is due, or post-processing reads it; meanwhile this thread is free (the next packet's entropy
       decode touches none of it) and other stream threads get the core. */
    if (b->dc_ahead_used) {
      /* the opt-in "DC ahead of the lists" variant keeps the call-by-call path */
      if (ocg_dec_submit(b->ctx, &f, NULL) < 0) { backend_fail(b, "ocg_dec_submit failed"); return; }
      r = b->out_mode == OCG_OUT_PADDED ? ocg_ctx_download_frame(b->ctx, self, host_self)
                                        : ocg_ctx_download_picture(b->ctx, self, host_self);
      if (r < 0) { backend_fail(b, "frame copy-back failed"); return; }
    } else if (b->expand) {
      /* the frame as the entropy decoder left it (decode.c:2822): the device does the rest */
      const oc_dec_ctx *dec = b->dec;
      ocg_dec_tokens t;
      int pli, zzi;
      memset(&t, 0, sizeof(t));
      for (i = 0; i < 3; i++) t.ref_idx[i] = f.ref_idx[i];
      t.lf_limit = f.lf_limit;
      t.intra_frame = f.intra_frame;
      t.dc_residual = b->dc_device ? 1 : 0;
      t.nqis = st->nqis;
      for (i = 0; i < 3; i++) t.qis[i] = st->qis[i < st->nqis ? i : 0];
      for (pli = 0; pli < 3; pli++) {
        for (k = 0; k < 2; k++) t.dc_quant[pli][k] = st->dequant_tables[st->qis[0]][pli][k][0]; /* decode.c:1366, 1534 */
        for (zzi = 0; zzi < 64; zzi++) {
          const ptrdiff_t e = dec->eob_runs[pli][zzi];
          t.ti0[pli][zzi] = (ogg_int32_t)dec->ti0[pli][zzi];
          t.eob_runs[pli][zzi] = e < 0 || e > 0x40000000 ? 0x40000000 : (ogg_int32_t)e;
        }
      }
      t.ntoken_bytes = dec->dct_tokens_count;
      if (ocg_dec_flush_tokens(b->ctx, &t, host_self, out_mode) < 0) { backend_fail(b, "ocg_dec_flush_tokens failed"); return; }
      extra_h2d += (long)b->geom.nfrags * 6 + t.ntoken_bytes - (long)b->geom.nfrags * 16;
    } else if (ocg_dec_flush(b->ctx, &f, host_self, out_mode) < 0) { backend_fail(b, "ocg_dec_flush failed"); return; }
    d2h = b->out_mode == OCG_OUT_PADDED ? (long)b->geom.ref_frame_sz : ocg_picture_bytes(&b->geom); /* reconstruction or post-processed planes */
    b->pending = 1;
    b->dev_valid[self] = 1;
    stats_add(b, extra_h2d + (long)b->geom.nfrags * 16 + (long)b->nrows * 16, d2h, now_s() - t0);
  }
  /* Out-of-loop post-processing (non-normative, decode.c:2899-2914) ran inside the MCU loop on a host
     frame that was not there yet; now that it is, run it again over the whole frame. */
  if (b->ctx != NULL && b->pp_level >= 2 /* OC_PP_LEVEL_DEBLOCKY */) {
    /* on the device, behind the frame's reconstruction on the same stream (ocg_dec_postproc.cu); fetched
       with the frame in backend_wait */
    const oc_fragment *frags = st->frags;
    const ptrdiff_t nfrags = st->nfrags;
    ptrdiff_t fragi;
    ogg_int32_t dc_scale[64], sharp_mod[64];
    if (b->pp == NULL) {
      b->pp_qis = (unsigned char *)malloc((size_t)nfrags);
      if (b->pp_qis == NULL || ocg_pp_create(&b->pp, b->ctx) < 0) { backend_fail(b, "post-processing set-up failed"); return; }
    }
    for (fragi = 0; fragi < nfrags; fragi++) b->pp_qis[fragi] = (unsigned char)st->qis[frags[fragi].qii]; /* decode.c:1926 */
    for (i = 0; i < 64; i++) { dc_scale[i] = b->dec->pp_dc_scale[i]; sharp_mod[i] = b->dec->pp_sharp_mod[i]; }
    if (ocg_pp_run(b->pp, f.ref_idx[OCG_FRAME_SELF], b->pp_level, dc_scale, sharp_mod, b->dec->dc_qis, b->pp_qis) < 0) {
      backend_fail(b, "post-processing on the device failed");
      return;
    }
    b->pp_pending = 1;
  }
  /* the stripe callback, once, with the whole (now final) frame:
     decode.c:2936-2940 flips the row range, the telemetry path at 2975 already
     calls it with the full range. */
  if (b->user_cb.stripe_decoded != NULL) {
    th_ycbcr_buffer stripe;
    backend_wait(b);
    if (b->failed) return;
    oc_ycbcr_buffer_flip(stripe, b->dec->pp_frame_buf);
    (*b->user_cb.stripe_decoded)(b->user_cb.ctx, stripe, 0, st->fplanes[0].nvfrags);
  }
}

/* ---- recorders ----------------------------------------------------------- */
static void ocg_dc_unpredict_mcu_plane(oc_dec_ctx *_dec, oc_dec_pipeline_state *_pipe, int _pli) {
  ocg_backend *b = backend_of(_dec);
  if (b != NULL && b->failed) b = NULL; /* latched error: behave like the plain C table until the API reports it */
  if (b != NULL && b->dc_device) {
    /* the device undoes the prediction (ocg_dc_unpredict_kernel); what is left of
       this hook is its side effect, decode.c:1496-1499: the MCU's fragment counts */
    const oc_fragment_plane *fplane = _dec->state.fplanes + _pli;
    const oc_fragment *frags = _dec->state.frags;
    const int fragy0 = _pipe->fragy0[_pli], fragy_end = _pipe->fragy_end[_pli];
    ptrdiff_t fragi = fplane->froffset + fragy0 * (ptrdiff_t)fplane->nhfrags;
    const ptrdiff_t end = fplane->froffset + fragy_end * (ptrdiff_t)fplane->nhfrags;
    ptrdiff_t ncoded = 0;
    for (; fragi < end; fragi++) ncoded += frags[fragi].coded;
    _pipe->ncoded_fragis[_pli] = ncoded;
    _pipe->nuncoded_fragis[_pli] = (fragy_end - fragy0) * (ptrdiff_t)fplane->nhfrags - ncoded;
  } else if (b != NULL) {
    ocg_host_dc_unpredict_mcu_plane(_dec, _pipe, _pli); /* same contract, restated for speed */
  } else {
    oc_dec_dc_unpredict_mcu_plane_c(_dec, _pipe, _pli);
  }
  if (b != NULL && !b->frame_open) backend_begin_frame(b);
  if (b != NULL && b->failed) return;
  if (b != NULL && b->expand) {
    /* The device walks the token lists itself (ocg_dec_flush_tokens): this hook owns the trip count of the
       reference's expansion loop (it computes pipe->ncoded_fragis / nuncoded_fragis, decode.c:1496-1499),
       so it leaves oc_dec_frags_recon_mcu_plane (decode.c:1511) an empty range. */
    b->ncoded += (int)_pipe->ncoded_fragis[_pli];
    _pipe->ncoded_fragis[_pli] = 0;
    _pipe->nuncoded_fragis[_pli] = 0;
  }
}

static inline int row_nonzero(const ogg_int16_t *row) {
  ogg_uint64_t a, c;
  memcpy(&a, row, 8);
  memcpy(&c, row + 4, 8);
  return (a | c) != 0;
}

static void ocg_state_frag_recon(const oc_theora_state *_state, ptrdiff_t _fragi, int _pli,
                                 ogg_int16_t _dct_coeffs[128], int _last_zzi, ogg_uint16_t _dc_quant) {
  ocg_backend *b = t_cur;
  const oc_fragment *frag = _state->frags + _fragi;
  ocg_frag_rec *rec;
  ogg_int16_t *out;
  int nr, r, qti, mask = 0;
  ogg_int16_t dc = _dct_coeffs[0];
  if (__builtin_expect(b == NULL || (const void *)b->dec != (const void *)_state, 0)) b = backend_of(_state);
  if (b == NULL || b->failed) return;
  if (!b->frame_open) { backend_fail(b, "state_frag_recon outside a frame"); return; }
  /* footprint of the transform the reference would run (state.c:967,
     idct.c:327-329): 0, 2, 4 or 8 leading rows */
  nr = _last_zzi < 2 ? 0 : (_last_zzi <= 3 ? 2 : (_last_zzi <= 10 ? 4 : 8));
  rec = b->st.recs + _fragi;
  rec->coeff_row = (ogg_uint32_t)b->nrows;
  _dct_coeffs[0] = 0; /* DC travels in the record */
  /* every footprint row is written at the list's end and kept only if it is
     non-zero (no branch per row; a zero row is overwritten by the next one) */
  out = b->st.coeff_rows + (size_t)b->nrows * 8;
  for (r = 0; r < nr; r++) {
    ogg_uint64_t a, c;
    int nz;
    memcpy(&a, _dct_coeffs + r * 8, 8);
    memcpy(&c, _dct_coeffs + r * 8 + 4, 8);
    memcpy(out, &a, 8);
    memcpy(out + 4, &c, 8);
    nz = (a | c) != 0;
    out += nz << 3;
    mask |= nz << r;
  }
  b->nrows = (int)((out - b->st.coeff_rows) >> 3);
  /* the iDCT contract: leave the coefficients zeroed for the next block
     (idct.c:245,276,295; decode.c:1385); the footprint rows are contiguous */
  memset(_dct_coeffs, 0, (size_t)nr * 16);
  qti = frag->mb_mode != OC_MODE_INTRA;
  /* buf_off and the plane are already in the record (template) */
  rec->mv = _state->frag_mvs[_fragi];
  rec->dc = dc;
  rec->rowmask = (unsigned char)mask;
  rec->last_zzi = (unsigned char)_last_zzi;
  rec->refi = (unsigned char)frag->refi;
  rec->pli_qti = (unsigned char)(_pli | qti << 2);
  b->dcq[_pli][qti] = _dc_quant;
  b->ncoded++;
}

static void ocg_frag_copy_list(unsigned char *_dst_frame, const unsigned char *_src_frame, int _ystride,
                               const ptrdiff_t *_fragis, ptrdiff_t _nfragis, const ptrdiff_t *_frag_buf_offs) {
  /* mark the listed fragments uncoded; the device copies them PREV -> SELF */
  ocg_backend *b = t_cur;
  ocg_frag_rec *recs;
  ptrdiff_t i;
  (void)_dst_frame; (void)_src_frame; (void)_ystride; (void)_frag_buf_offs;
  if (b == NULL || b->failed) return;
  if (!b->frame_open) { backend_fail(b, "frag_copy_list outside a frame"); return; }
  recs = b->st.recs;
  for (i = 0; i < _nfragis; i++) recs[_fragis[i]].refi = OCG_FRAG_UNCODED;
}

static void ocg_state_loop_filter_frag_rows(const oc_theora_state *_state, signed char _bv[256], int _refi,
                                            int _pli, int _fragy0, int _fragy_end) {
  /* filtered on the device over the whole frame at flush */
  (void)_state; (void)_bv; (void)_refi; (void)_pli; (void)_fragy0; (void)_fragy_end;
}

static void ocg_restore_fpu(void) {
  ocg_backend *b = t_cur;
  if (b != NULL && b->frame_open && !b->failed) backend_flush(b);
}

/* ---- init functions named by ocg_hooks.h --------------------------------- */
#if defined(OC_X86_ASM)
/* x86int.h names oc_state_accel_init_x86 as the init function of every unit built with OC_X86_ASM: that is
   this function here (theora_b200/backend/Makefile); the reference's own body is oc_refimpl_state_accel_init_x86
   and is used for encoders that keep their block kernels on the host (ocg_enc_backend.c). */
void oc_state_accel_init_x86(oc_theora_state *_state) { oc_state_accel_init_ocg(_state); }
#endif

void oc_state_accel_init_ocg(oc_theora_state *_state) {
  /* shared encoder/decoder table: plain C entries; the decoder and encoder inits
     override the ones they offload (ocg_enc_backend.c for intra-only encoders). */
  oc_state_accel_init_c(_state);
}

/* oc_enc_accel_init_ocg lives in ocg_enc_backend.c */
int ocg_backend_device_(void) { return g_device; }

static void backend_destroy(ocg_backend *b) {
  ocg_backend **pp;
  if (b == NULL) return;
  pthread_mutex_lock(&g_lock);
  for (pp = &g_list; *pp != NULL && *pp != b; pp = &(*pp)->next) {}
  if (*pp == b) *pp = b->next;
  stats_sum(&g_stats_retired, &b->stats);
  pthread_mutex_unlock(&g_lock);
  if (t_cur == b) t_cur = NULL;
  if (b->ctx != NULL) {
    ocg_ctx_sync(b->ctx);
    if (b->pp != NULL) ocg_pp_destroy(b->pp);
    backend_unregister(b);
    if (b->pinned) ocg_host_unregister(b->dec->state.ref_frame_handle);
    ocg_ctx_destroy(b->ctx);
  }
  free(b->heap_staging);
  free(b->pp_qis);
  free(b);
}

/* ---- device-side token expansion: one-time set-up ------------------------------------------------- */
static void backend_unregister(ocg_backend *b) {
  oc_theora_state *st = &b->dec->state;
  if (b->regs & 1) ocg_host_unregister(st->frags);
  if (b->regs & 2) ocg_host_unregister(st->frag_mvs);
  if (b->regs & 4) ocg_host_unregister(b->dec->dct_tokens);
  b->regs = 0;
}

static int backend_expand_setup(ocg_backend *b) {
  oc_dec_ctx *dec = b->dec;
  oc_theora_state *st = &dec->state;
  const size_t nf = (size_t)st->nfrags, token_cap = (64 + 64 + 1) * nf; /* decode.c:386 */
  ogg_int32_t *order = (ogg_int32_t *)malloc(nf * sizeof(*order));
  ogg_uint16_t *deq = (ogg_uint16_t *)malloc(64 * 3 * 2 * 64 * sizeof(*deq));
  size_t n = 0;
  unsigned sbi;
  int qi, pli, qti, quadi, bi, r = -1;
  oc_fragment probe;
  ogg_uint32_t w = 0;
  /* the device reads oc_fragment as a 32-bit word (state.h:297-322 as this compiler lays the bit-fields out) */
  memset(&probe, 0, sizeof(probe));
  probe.coded = 1; probe.qii = 5; probe.refi = 2; probe.mb_mode = 6; probe.dc = -3;
  if (sizeof(probe) == 4) memcpy(&w, &probe, 4);
  if (order == NULL || deq == NULL || sizeof(probe) != 4 ||
      w != (1u | 5u << 2 | 2u << 6 | 6u << 8 | 0xFFFDu << 16) || sizeof(st->frag_mvs[0]) != 2) {
    free(order);
    free(deq);
    return -1;
  }
  /* coded order: super blocks of a plane in raster order, Hilbert order inside (decode.c:548-600, 626-700) */
  for (sbi = 0; sbi < st->nsbs; sbi++)
    for (quadi = 0; quadi < 4; quadi++)
      for (bi = 0; bi < 4; bi++) {
        const ptrdiff_t fragi = st->sb_maps[sbi][quadi][bi];
        if (fragi >= 0 && n < nf) order[n++] = (ogg_int32_t)fragi;
      }
  for (qi = 0; qi < 64; qi++)
    for (pli = 0; pli < 3; pli++)
      for (qti = 0; qti < 2; qti++)
        memcpy(deq + (((size_t)qi * 3 + pli) * 2 + qti) * 64, st->dequant_tables[qi][pli][qti], 64 * sizeof(*deq));
  if (n == nf) {
    if (ocg_host_register(st->frags, nf * sizeof(st->frags[0])) == 0) b->regs |= 1;
    if (ocg_host_register(st->frag_mvs, nf * sizeof(st->frag_mvs[0])) == 0) b->regs |= 2;
    if (ocg_host_register(dec->dct_tokens, token_cap) == 0) b->regs |= 4;
    if (b->regs == 7)
      r = ocg_dec_expand_setup(b->ctx, order, deq, (const ogg_uint32_t *)st->frags, (const ogg_int16_t *)st->frag_mvs,
                               dec->dct_tokens, token_cap);
    if (r < 0) backend_unregister(b);
  }
  free(order);
  free(deq);
  return r;
}

void oc_dec_accel_init_ocg(th_dec_ctx *_dec) {
  oc_theora_state *st = &_dec->state;
  ocg_backend *b;
  ptrdiff_t last;
  oc_dec_accel_init_c(_dec);
  b = (ocg_backend *)calloc(1, sizeof(*b));
  if (b == NULL) return;
  b->dec = _dec;
  b->mode = g_mode;
  if (ocg_geometry_init(&b->geom, (int)st->info.frame_width, (int)st->info.frame_height, (int)st->info.pixel_fmt, 3) < 0) {
    fprintf(stderr, "theora_b200 back-end: %s\n", ocg_last_error());
    free(b);
    return;
  }
  /* the device mirror must be byte-compatible with state.c:545-671 */
  last = st->nfrags - 1;
  if (b->geom.nfrags != st->nfrags || b->geom.planes[0].ystride != st->ref_ystride[0] ||
      b->geom.planes[1].ystride != st->ref_ystride[1] ||
      st->ref_frame_bufs[0][0].data - st->ref_frame_handle != b->geom.base_off ||
      st->ref_frame_bufs[1][0].data - st->ref_frame_bufs[0][0].data != b->geom.ref_frame_sz ||
      st->frag_buf_offs[last] != b->geom.planes[2].plane_off +
       (ptrdiff_t)(b->geom.planes[2].nvfrags - 1) * 8 * b->geom.planes[2].ystride + (b->geom.planes[2].nhfrags - 1) * 8) {
    fprintf(stderr, "theora_b200 back-end: frame layout differs from the reference's\n");
    free(b);
    return;
  }
  /* the token walk moves to the device unless someone wants to see the lists (capture hook, record mode) */
  b->expand = g_expand_mode == OCG_EXPAND_DEVICE && b->mode == OCG_BACKEND_GPU && g_capture == NULL;
  b->out_mode = g_out_mode;
  b->dc_device = (g_dc_mode == OCG_DC_DEVICE || g_dc_mode == OCG_DC_DEVICE_AHEAD) &&
                 (b->mode != OCG_BACKEND_GPU || ocg_dc_unpredict_supported(&b->geom));
  if (b->dc_device && g_dc_mode == OCG_DC_DEVICE_AHEAD && !b->expand) b->dc_device = 2;
  if (b->dc_device && b->mode == OCG_BACKEND_GPU) {
    /* the device reads oc_fragment as a 32-bit word: bit 0 coded, bits 6-7 refi, bits 16-31 dc (state.h:297-322
       as this compiler lays the bit-fields out); verify instead of assuming */
    oc_fragment t;
    ogg_uint32_t w;
    memset(&t, 0, sizeof(t));
    t.coded = 1; t.refi = 2; t.dc = -3;
    memcpy(&w, &t, sizeof(w));
    if (sizeof(t) != 4 || w != (1u | 2u << 6 | 0xFFFDu << 16)) b->dc_device = 0;
  }
  if (b->mode == OCG_BACKEND_GPU) {
    if (ocg_ctx_create(&b->ctx, &b->geom, g_device) < 0) {
      fprintf(stderr, "theora_b200 back-end: %s\n", ocg_last_error());
      free(b);
      return; /* th_decode_alloc (below) reports the failure; there is no CPU fallback */
    }
    b->pinned = ocg_host_register(st->ref_frame_handle, (size_t)b->geom.ref_frame_sz * 3) == 0;
    if (!b->pinned) {
      /* the flush writes the finished frame straight into the decoder's own buffer */
      fprintf(stderr, "theora_b200 back-end: cannot page-lock the frame buffers (%s)\n", ocg_last_error());
      ocg_ctx_destroy(b->ctx);
      free(b);
      return;
    }
    if (b->expand && backend_expand_setup(b) < 0) {
      fprintf(stderr, "theora_b200 back-end: %s\n", ocg_last_error());
      backend_unregister(b);
      ocg_host_unregister(st->ref_frame_handle);
      ocg_ctx_destroy(b->ctx);
      free(b);
      return;
    }
  } else {
    size_t nf = (size_t)b->geom.nfrags, i;
    unsigned char *p = (unsigned char *)malloc(nf * (16 + 16 + 128) + 64);
    int pli;
    if (p == NULL) { free(b); return; }
    b->heap_staging = p;
    b->st.recs = (ocg_frag_rec *)p;
    b->heap_tmpl = (ocg_frag_rec *)(p + nf * 16);
    b->st.coeff_rows = (int16_t *)(p + nf * 32);
    memset(b->heap_tmpl, 0, nf * 16);
    for (pli = 0; pli < 3; pli++) {
      for (i = 0; i < (size_t)b->geom.planes[pli].nfrags; i++) {
        ocg_frag_rec *rc = b->heap_tmpl + b->geom.planes[pli].froffset + i;
        rc->buf_off = (ogg_int32_t)st->frag_buf_offs[b->geom.planes[pli].froffset + i];
        rc->refi = OCG_FRAG_UNCODED;
        rc->pli_qti = (unsigned char)pli;
      }
    }
    memcpy(b->st.recs, b->heap_tmpl, nf * 16);
  }
  st->opt_vtable.state_frag_recon = ocg_state_frag_recon;
  st->opt_vtable.frag_copy_list = ocg_frag_copy_list;
  st->opt_vtable.state_loop_filter_frag_rows = ocg_state_loop_filter_frag_rows;
  st->opt_vtable.restore_fpu = ocg_restore_fpu;
  _dec->opt_vtable.dc_unpredict_mcu_plane = ocg_dc_unpredict_mcu_plane;
  pthread_mutex_lock(&g_lock);
  b->next = g_list;
  g_list = b;
  pthread_mutex_unlock(&g_lock);
  t_cur = b;
}

/* ---- public API wrappers -------------------------------------------------- */
th_dec_ctx *th_decode_alloc(const th_info *_info, const th_setup_info *_setup) {
  th_dec_ctx *dec = oc_refimpl_decode_alloc(_info, _setup);
  if (dec != NULL && backend_of(dec) == NULL) {
    /* no device / layout mismatch: fail the allocation rather than decode on the CPU */
    oc_refimpl_decode_free(dec);
    return NULL;
  }
  return dec;
}

/* The flush at the end of th_decode_packetin (decode.c:2965) is asynchronous; the frame must be in host
   memory when the application looks at it. */
int th_decode_packetin(th_dec_ctx *_dec, const ogg_packet *_op, ogg_int64_t *_granpos) {
  ocg_backend *b = _dec != NULL ? backend_of(_dec) : NULL;
  int ret;
  if (b != NULL && b->failed) return TH_EFAULT;
  /* the device reads frags[] / frag_mvs[] / dct_tokens[] in place while a flush is in flight */
  if (b != NULL && b->expand) backend_wait(b);
  if (b != NULL && b->failed) return TH_EFAULT;
  ret = oc_refimpl_decode_packetin(_dec, _op, _granpos);
  if (b != NULL && b->failed) return TH_EFAULT;
  return ret;
}

int th_decode_ycbcr_out(th_dec_ctx *_dec, th_ycbcr_buffer _ycbcr) {
  ocg_backend *b = _dec != NULL ? backend_of(_dec) : NULL;
  if (b != NULL) {
    backend_wait(b);
    if (b->failed) return TH_EFAULT;
  }
  return oc_refimpl_decode_ycbcr_out(_dec, _ycbcr);
}

void th_decode_free(th_dec_ctx *_dec) {
  if (_dec != NULL) backend_destroy(backend_of(_dec));
  oc_refimpl_decode_free(_dec);
}

int th_decode_ctl(th_dec_ctx *_dec, int _req, void *_buf, size_t _buf_sz) {
  ocg_backend *b = _dec != NULL ? backend_of(_dec) : NULL;
  if (b != NULL) {
    if (_req == TH_DECCTL_SET_STRIPE_CB) {
      if (_buf == NULL) return TH_EFAULT;
      if (_buf_sz != sizeof(th_stripe_callback)) return TH_EINVAL;
      b->user_cb = *(th_stripe_callback *)_buf;
      return 0;
    }
  }
  return oc_refimpl_decode_ctl(_dec, _req, _buf, _buf_sz);
}
