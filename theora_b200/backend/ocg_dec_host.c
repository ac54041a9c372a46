/* ocg_dec_host.c -- the host-side piece of the record-and-flush decoder back-end that is a restatement
 * of a reference routine: the hook's DC un-prediction, and its self-test against the reference routine
 * (oc_dec_dc_unpredict_mcu_plane_c, reached as the ordinary external symbol of decode.c).
 * (Post-processing used to be re-run here by the reference's own file-static filters; it now runs on the
 * device, ocg_dec_postproc.cu, and lib/decode.c is no longer compiled into this unit.) */
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "decint.h"
#include "ocg_backend.h"

/* DC un-prediction in the hook, restated for speed.
 *
 * Same contract as oc_dec_dc_unpredict_mcu_plane_c (decode.c:1392-1500): undoes
 * the DC prediction of fragment rows [fragy0,fragy_end) of one plane in place
 * (frags[].dc), carries pipe->pred_last across calls, and counts the coded /
 * uncoded fragments of the MCU.  The reference dispatches every coded fragment
 * through a 16-way switch on which of its four causal neighbours share its
 * reference type; on frames that mix reference types that switch mispredicts.
 * Here the neighbour test yields a 4-bit pattern that indexes a small table of
 * weights and shifts (the same integer formulas: x/2^k with truncation towards
 * zero == (x + ((x>>31) & (2^k-1))) >> k), so only the two special cases stay
 * as branches: "no neighbour" (pred_last) and the three-neighbour gradient
 * predictor with its outlier clamps (cases 7 and 15). */
typedef struct ocg_dc_rule {
  short wl, wul, wu, wur; /* weights of left, up-left, up, up-right */
  short shift;            /* divisor 2^shift */
  short special;          /* 0 table, 1 pred_last, 2 gradient predictor with clamps */
} ocg_dc_rule;

/* index: (l==ref) | (ul==ref)<<1 | (u==ref)<<2 | (ur==ref)<<3, decode.c:1450-1484 */
static const ocg_dc_rule OCG_DC_RULES[16] = {
  {0, 0, 0, 0, 0, 1},     /*  0: pred_last                      */
  {1, 0, 0, 0, 0, 0},     /*  1: l                              */
  {0, 1, 0, 0, 0, 0},     /*  2: ul                             */
  {1, 0, 0, 0, 0, 0},     /*  3: l                              */
  {0, 0, 1, 0, 0, 0},     /*  4: u                              */
  {1, 0, 1, 0, 1, 0},     /*  5: (l+u)/2                        */
  {0, 0, 1, 0, 0, 0},     /*  6: u                              */
  {29, -26, 29, 0, 5, 2}, /*  7: (29*(l+u)-26*ul)/32 + clamps   */
  {0, 0, 0, 1, 0, 0},     /*  8: ur                             */
  {75, 0, 0, 53, 7, 0},   /*  9: (75*l+53*ur)/128               */
  {0, 1, 0, 1, 1, 0},     /* 10: (ul+ur)/2                      */
  {75, 0, 0, 53, 7, 0},   /* 11                                 */
  {0, 0, 1, 0, 0, 0},     /* 12: u                              */
  {75, 0, 0, 53, 7, 0},   /* 13                                 */
  {0, 3, 10, 3, 4, 0},    /* 14: (3*(ul+ur)+10*u)/16            */
  {29, -26, 29, 0, 5, 2}  /* 15                                 */
};

/* oc_fragment as one 32-bit word (state.h:297-322 as GCC lays the bit-fields out on little-endian targets;
   ocg_dc_words_ok() verifies it once): bit 0 coded, bits 6-7 refi, bits 16-31 dc. */
#define OCG_W_CODED(w) ((w) & 1u)
#define OCG_W_REFI(w)  ((int)((w) >> 6 & 3u))
#define OCG_W_DC(w)    ((int)(ogg_int16_t)((w) >> 16))

static int ocg_dc_words_ok(void) {
  static int ok = -1;
  if (ok < 0) {
    oc_fragment t;
    ogg_uint32_t w = 0;
    memset(&t, 0, sizeof(t));
    t.coded = 1; t.refi = 2; t.dc = -3;
    if (sizeof(t) == 4) memcpy(&w, &t, 4);
    ok = sizeof(t) == 4 && w == (1u | 2u << 6 | 0xFFFDu << 16);
  }
  return ok;
}

void ocg_host_dc_unpredict_mcu_plane(oc_dec_ctx *_dec, oc_dec_pipeline_state *_pipe, int _pli) {
  const oc_fragment_plane *fplane = _dec->state.fplanes + _pli;
  ogg_uint32_t *words = (ogg_uint32_t *)_dec->state.frags;
  int *pred_last = _pipe->pred_last[_pli];
  const int fragy0 = _pipe->fragy0[_pli], fragy_end = _pipe->fragy_end[_pli], nhfrags = fplane->nhfrags;
  ptrdiff_t ncoded = 0, fragi = fplane->froffset + fragy0 * (ptrdiff_t)nhfrags;
  int fragx, fragy;
  if (!ocg_dc_words_ok()) { /* unexpected bit-field layout: the reference routine knows its own */
    oc_dec_dc_unpredict_mcu_plane_c(_dec, _pipe, _pli);
    return;
  }
  for (fragy = fragy0; fragy < fragy_end; fragy++) {
    if (fragy == 0) {
      for (fragx = 0; fragx < nhfrags; fragx++, fragi++) {
        const ogg_uint32_t w = words[fragi];
        if (OCG_W_CODED(w)) {
          const int refi = OCG_W_REFI(w);
          const int dc = (ogg_int16_t)(OCG_W_DC(w) + pred_last[refi]);
          words[fragi] = (w & 0xFFFFu) | (ogg_uint32_t)dc << 16;
          pred_last[refi] = dc;
          ncoded++;
        }
      }
    } else {
      const ogg_uint32_t *u_words = words - nhfrags;
      /* like the reference, the rows above are judged by refi alone: an uncoded fragment carries
         OC_FRAME_NONE there (decode.c:658) */
      ogg_uint32_t uw = u_words[fragi];
      int l_ref = -1, ul_ref = -1, u_ref = OCG_W_REFI(uw);
      int l_dc = 0, ul_dc = 0, u_dc = OCG_W_DC(uw);
      for (fragx = 0; fragx < nhfrags; fragx++, fragi++) {
        const ogg_uint32_t w = words[fragi];
        int ur_ref = -1, ur_dc = 0;
        if (fragx + 1 < nhfrags) {
          const ogg_uint32_t urw = u_words[fragi + 1];
          ur_ref = OCG_W_REFI(urw);
          ur_dc = OCG_W_DC(urw);
        }
        if (OCG_W_CODED(w)) {
          const int refi = OCG_W_REFI(w);
          const ocg_dc_rule *r = OCG_DC_RULES + ((l_ref == refi) | (ul_ref == refi) << 1 | (u_ref == refi) << 2 |
                                                 (ur_ref == refi) << 3);
          int pred, dc;
          if (r->special == 2) {
            /* the pattern of dense regions (every frame's interior): constants in the code keep the
               loop-carried chain through l_dc short */
            pred = (29 * (l_dc + u_dc) - 26 * ul_dc) / 32;
            if (abs(pred - u_dc) > 128) pred = u_dc;
            else if (abs(pred - l_dc) > 128) pred = l_dc;
            else if (abs(pred - ul_dc) > 128) pred = ul_dc;
          } else if (r->special == 1) pred = pred_last[refi];
          else {
            int sum = r->wl * l_dc + r->wul * ul_dc + r->wu * u_dc + r->wur * ur_dc;
            pred = (sum + ((sum >> 31) & ((1 << r->shift) - 1))) >> r->shift;
          }
          dc = (ogg_int16_t)(OCG_W_DC(w) + pred); /* frags[].dc is a 16-bit field */
          words[fragi] = (w & 0xFFFFu) | (ogg_uint32_t)dc << 16;
          pred_last[refi] = dc;
          ncoded++;
          l_ref = refi;
          l_dc = dc;
        } else l_ref = -1;
        ul_ref = u_ref;
        ul_dc = u_dc;
        u_ref = ur_ref;
        u_dc = ur_dc;
      }
    }
  }
  _pipe->ncoded_fragis[_pli] = ncoded;
  _pipe->nuncoded_fragis[_pli] = (fragy_end - fragy0) * (ptrdiff_t)nhfrags - ncoded;
}

/* Test hook: both implementations on the same random plane (fake decoder context); returns the number of
   fragments whose DC differs, or -1 if the counts differ. */
OCG_API long ocg_host_dc_selftest(int nhfrags, int nvfrags, int mcu_rows, unsigned seed, int coded_pct, int mixed) {
  oc_dec_ctx *da = (oc_dec_ctx *)calloc(1, sizeof(*da)), *db = (oc_dec_ctx *)calloc(1, sizeof(*db));
  oc_dec_pipeline_state *pa = &da->pipe, *pb = &db->pipe;
  size_t n = (size_t)nhfrags * nvfrags, i;
  oc_fragment *fa = (oc_fragment *)calloc(n, sizeof(*fa)), *fb = (oc_fragment *)calloc(n, sizeof(*fb));
  long bad = 0;
  int y0;
  unsigned s = seed * 2654435761u + 12345u;
  for (i = 0; i < n; i++) {
    s = s * 1664525u + 1013904223u;
    fa[i].coded = (s >> 8) % 100 < (unsigned)coded_pct;
    fa[i].refi = mixed ? (s >> 16) % 3 : 1;
    s = s * 1664525u + 1013904223u;
    fa[i].dc = (int)((s >> 12) % 1201) - 600;
    if (((s >> 28) & 15) == 0) fa[i].dc = (int)(s >> 16) - 32768; /* the odd huge value: 16-bit wrap */
    fb[i] = fa[i];
  }
  da->state.frags = fa; db->state.frags = fb;
  da->state.fplanes[0].nhfrags = db->state.fplanes[0].nhfrags = nhfrags;
  da->state.fplanes[0].nvfrags = db->state.fplanes[0].nvfrags = nvfrags;
  for (y0 = 0; y0 < nvfrags; y0 += mcu_rows) {
    pa->fragy0[0] = pb->fragy0[0] = y0;
    pa->fragy_end[0] = pb->fragy_end[0] = y0 + mcu_rows < nvfrags ? y0 + mcu_rows : nvfrags;
    oc_dec_dc_unpredict_mcu_plane_c(da, pa, 0);
    ocg_host_dc_unpredict_mcu_plane(db, pb, 0);
    if (pa->ncoded_fragis[0] != pb->ncoded_fragis[0] || pa->nuncoded_fragis[0] != pb->nuncoded_fragis[0]) bad = -1;
  }
  if (bad == 0)
    for (i = 0; i < n; i++) bad += fa[i].dc != fb[i].dc;
  free(fa); free(fb); free(da); free(db);
  return bad;
}

/* Test/diagnostic hook: nanoseconds per fragment of either DC implementation (which: 0 reference routine,
   1 restated) on a random plane, best of `reps` whole-plane passes (residuals restored before each). */
OCG_API double ocg_host_dc_bench(int nhfrags, int nvfrags, int mcu_rows, unsigned seed, int coded_pct, int mixed,
                                 int which, int reps) {
  oc_dec_ctx *d = (oc_dec_ctx *)calloc(1, sizeof(*d));
  oc_dec_pipeline_state *p = &d->pipe;
  size_t n = (size_t)nhfrags * nvfrags, i;
  oc_fragment *f0 = (oc_fragment *)calloc(n, sizeof(*f0)), *f = (oc_fragment *)calloc(n, sizeof(*f));
  double best = 1e30;
  unsigned s = seed * 2654435761u + 12345u;
  int rep, y0;
  for (i = 0; i < n; i++) {
    s = s * 1664525u + 1013904223u;
    f0[i].coded = (s >> 8) % 100 < (unsigned)coded_pct;
    f0[i].refi = f0[i].coded ? (mixed ? (s >> 16) % 3 : 1) : 3;
    s = s * 1664525u + 1013904223u;
    f0[i].dc = (int)((s >> 12) % 1201) - 600;
  }
  d->state.frags = f;
  d->state.fplanes[0].nhfrags = nhfrags;
  d->state.fplanes[0].nvfrags = nvfrags;
  for (rep = 0; rep < reps; rep++) {
    struct timespec t0, t1;
    double dt;
    memcpy(f, f0, n * sizeof(*f));
    memset(p->pred_last, 0, sizeof(p->pred_last));
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (y0 = 0; y0 < nvfrags; y0 += mcu_rows) {
      p->fragy0[0] = y0;
      p->fragy_end[0] = y0 + mcu_rows < nvfrags ? y0 + mcu_rows : nvfrags;
      if (which) ocg_host_dc_unpredict_mcu_plane(d, p, 0);
      else oc_dec_dc_unpredict_mcu_plane_c(d, p, 0);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    dt = (double)(t1.tv_sec - t0.tv_sec) * 1e9 + (double)(t1.tv_nsec - t0.tv_nsec);
    if (dt < best) best = dt;
  }
  free(f0); free(f); free(d);
  return best / (double)n;
}
