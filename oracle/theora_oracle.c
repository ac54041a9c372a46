/* TEST INFRASTRUCTURE ONLY -- see theora_oracle.h.
 *
 * CPU restatement of libtheora's per-fragment 8x8 block pipeline.  Every
 * function cites the reference file:line whose behaviour it restates; the
 * arithmetic (16-bit wrap-arounds, arithmetic right shifts, rounding biases)
 * is normative and therefore identical, the code structure is our own.
 * Parity with the compiled reference is enforced by tests/test_oracle_*.py.
 */
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "theora_oracle.h"

#define OCO_EXPORT __attribute__((visibility("default")))

/* cos(k*pi/16)*65536, reference lib/dct.h:21-28 */
enum { K1 = 64277, K2 = 60547, K3 = 54491, K4 = 46341, K5 = 36410, K6 = 25080, K7 = 12785 };

static inline int32_t mulhi16(int32_t k, int32_t v) { return (k * v) >> 16; }
static inline int16_t wrap16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }
static inline uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* One 8-point pass of idct.c:30-81 (idct8).  The reduced variants idct8_1..4
   (idct.c:91-203) are this routine with the trailing inputs equal to zero, so
   callers zero what the reference ignores instead of duplicating them. */
static void idct8_pass(int16_t *out, int ostep, const int16_t in[8]) {
  int32_t e0 = mulhi16(K4, wrap16(in[0] + in[4]));
  int32_t e1 = mulhi16(K4, wrap16(in[0] - in[4]));
  int32_t e2 = mulhi16(K6, in[2]) - mulhi16(K2, in[6]);
  int32_t e3 = mulhi16(K2, in[2]) + mulhi16(K6, in[6]);
  int32_t o4 = mulhi16(K7, in[1]) - mulhi16(K1, in[7]);
  int32_t o5 = mulhi16(K3, in[5]) - mulhi16(K5, in[3]);
  int32_t o6 = mulhi16(K5, in[5]) + mulhi16(K3, in[3]);
  int32_t o7 = mulhi16(K1, in[1]) + mulhi16(K7, in[7]);
  int32_t s4 = o4 + o5;
  int32_t s5 = mulhi16(K4, wrap16(o4 - o5));
  int32_t s7 = o7 + o6;
  int32_t s6 = mulhi16(K4, wrap16(o7 - o6));
  int32_t a0 = e0 + e3, a3 = e0 - e3;
  int32_t a1 = e1 + e2, a2 = e1 - e2;
  int32_t b6 = s6 + s5, b5 = s6 - s5;
  out[0 * ostep] = wrap16(a0 + s7);
  out[1 * ostep] = wrap16(a1 + b6);
  out[2 * ostep] = wrap16(a2 + b5);
  out[3 * ostep] = wrap16(a3 + s4);
  out[4 * ostep] = wrap16(a3 - s4);
  out[5 * ostep] = wrap16(a2 - b5);
  out[6 * ostep] = wrap16(a1 - b6);
  out[7 * ostep] = wrap16(a0 - s7);
}

/* idct.c:301-330.  Class selection by last_zzi, row pass into columns of w,
   column pass, (v+8)>>4, and the input-clearing side effect the decoder relies
   on (decode.c:1385). */
OCO_EXPORT void oco_idct8x8(int16_t y[64], int16_t x[64], int last_zzi) {
  /* number of leading coefficients the reference reads in each row */
  static const uint8_t used3[8] = {2, 1, 0, 0, 0, 0, 0, 0};
  static const uint8_t used10[8] = {4, 3, 2, 1, 0, 0, 0, 0};
  static const uint8_t usedall[8] = {8, 8, 8, 8, 8, 8, 8, 8};
  const uint8_t *used = last_zzi <= 3 ? used3 : (last_zzi <= 10 ? used10 : usedall);
  int ncols = last_zzi <= 3 ? 2 : (last_zzi <= 10 ? 4 : 8);
  int16_t w[64];
  int16_t row[8];
  int i, j;
  memset(w, 0, sizeof(w));
  for (i = 0; i < 8; i++) {
    if (used[i] == 0) continue;
    for (j = 0; j < 8; j++) row[j] = j < used[i] ? x[i * 8 + j] : 0;
    idct8_pass(w + i, 8, row);
  }
  for (i = 0; i < 8; i++) {
    for (j = 0; j < 8; j++) row[j] = j < ncols ? w[i * 8 + j] : 0;
    idct8_pass(y + i, 8, row);
  }
  for (i = 0; i < 64; i++) y[i] = wrap16((y[i] + 8) >> 4);
  for (i = 0; i < 8; i++)
    for (j = 0; j < used[i]; j++) x[i * 8 + j] = 0;
}

/* fragment.c:49-57 */
OCO_EXPORT void oco_frag_recon_intra(uint8_t *dst, int ystride, const int16_t res[64]) {
  int r, c;
  for (r = 0; r < 8; r++, dst += ystride)
    for (c = 0; c < 8; c++) dst[c] = clamp255(res[r * 8 + c] + 128);
}

/* fragment.c:59-68 */
OCO_EXPORT void oco_frag_recon_inter(uint8_t *dst, const uint8_t *src, int ystride, const int16_t res[64]) {
  int r, c;
  for (r = 0; r < 8; r++, dst += ystride, src += ystride)
    for (c = 0; c < 8; c++) dst[c] = clamp255(res[r * 8 + c] + src[c]);
}

/* fragment.c:70-80 */
OCO_EXPORT void oco_frag_recon_inter2(uint8_t *dst, const uint8_t *s1, const uint8_t *s2, int ystride,
                                      const int16_t res[64]) {
  int r, c;
  for (r = 0; r < 8; r++, dst += ystride, s1 += ystride, s2 += ystride)
    for (c = 0; c < 8; c++) dst[c] = clamp255(res[r * 8 + c] + ((s1[c] + s2[c]) >> 1));
}

/* fragment.c:20-27 */
OCO_EXPORT void oco_frag_copy(uint8_t *dst, const uint8_t *src, int ystride) {
  int r;
  for (r = 0; r < 8; r++, dst += ystride, src += ystride) memcpy(dst, src, 8);
}

/* state.c:846-957.  Closed form of the OC_MVMAP/OC_MVMAP2 tables: the first
   tap truncates the vector towards zero at half-pel (quarter-pel in a
   decimated chroma direction), the second tap -- present when either component
   has a fractional part -- is one step further away from zero. */
static void mv_component(int d, int qpel, int *ipart, int *step) {
  int mag = d < 0 ? -d : d;
  int sgn = d < 0 ? -1 : 1;
  int frac = qpel ? (mag & 3) : (mag & 1);
  *ipart = sgn * (mag >> (qpel ? 2 : 1));
  *step = frac ? sgn : 0;
}

OCO_EXPORT int oco_mv_offsets(int offs[2], int ystride, int pli, int pixel_fmt, int16_t mv) {
  int dx = (signed char)(mv & 0xFF);
  int dy = mv >> 8;
  int qx = pli != 0 && !(pixel_fmt & 1);
  int qy = pli != 0 && !(pixel_fmt & 2);
  int mx, my, mx2, my2;
  mv_component(dx, qx, &mx, &mx2);
  mv_component(dy, qy, &my, &my2);
  offs[0] = my * ystride + mx;
  if (mx2 || my2) {
    offs[1] = offs[0] + my2 * ystride + mx2;
    return 2;
  }
  return 1;
}

/* state.c:959-1000 */
OCO_EXPORT void oco_state_frag_recon(uint8_t *dst_frame, const uint8_t *ref_frame, int32_t buf_off,
                                     int ystride, int pli, int pixel_fmt, int intra, int16_t mv,
                                     int16_t coeffs[128], int last_zzi, uint16_t dc_quant) {
  uint8_t *dst = dst_frame + buf_off;
  int i;
  if (last_zzi < 2) {
    int16_t p = wrap16((coeffs[0] * (int32_t)dc_quant + 15) >> 5);
    for (i = 0; i < 64; i++) coeffs[64 + i] = p;
  } else {
    coeffs[0] = wrap16(coeffs[0] * (int32_t)dc_quant);
    oco_idct8x8(coeffs + 64, coeffs, last_zzi);
  }
  if (intra) oco_frag_recon_intra(dst, ystride, coeffs + 64);
  else {
    const uint8_t *ref = ref_frame + buf_off;
    int offs[2];
    if (oco_mv_offsets(offs, ystride, pli, pixel_fmt, mv) > 1)
      oco_frag_recon_inter2(dst, ref + offs[0], ref + offs[1], ystride, coeffs + 64);
    else oco_frag_recon_inter(dst, ref + offs[0], ystride, coeffs + 64);
  }
}

/* Closed form of the bounding-value table built by state.c:1036-1045. */
OCO_EXPORT int oco_lflim(int r, int limit) {
  int a = r < 0 ? -r : r;
  int v;
  if (a < limit) v = a;
  else if (a < 2 * limit) v = 2 * limit - a;
  else v = 0;
  return r < 0 ? -v : v;
}

/* state.c:1036-1045, table form (kept to pin the closed form above). */
OCO_EXPORT void oco_loop_filter_init(signed char bv[256], int limit) {
  int r;
  for (r = -127; r <= 128; r++) bv[127 + r] = (signed char)oco_lflim(r, limit);
}

/* One line of loop_filter_h / loop_filter_v (state.c:1002-1031): p[0..3*step]
   straddle the edge; the middle two samples are corrected. */
static inline void lf_line(uint8_t *p, ptrdiff_t step, int limit) {
  int f = p[0] - p[3 * step] + 3 * (p[2 * step] - p[step]);
  f = oco_lflim((f + 4) >> 3, limit);
  p[step] = clamp255(p[step] + f);
  p[2 * step] = clamp255(p[2 * step] - f);
}

/* Filter across the vertical edge at `pix` (left column of a fragment), rows
   0..7: loop_filter_h. */
static void lf_vedge(uint8_t *pix, int ystride, int limit) {
  int r;
  for (r = 0; r < 8; r++) lf_line(pix - 2 + (ptrdiff_t)r * ystride, 1, limit);
}

/* Filter across the horizontal edge at row `pix`: loop_filter_v. */
static void lf_hedge(uint8_t *pix, int ystride, int limit) {
  int c;
  for (c = 0; c < 8; c++) lf_line(pix + c - 2 * (ptrdiff_t)ystride, ystride, limit);
}

/* state.c:1055-1105: raster over fragments from row 0 (bottom); a coded
   fragment filters its left and lower edges, and its right/upper edges only
   when that neighbour is not coded. */
OCO_EXPORT void oco_loop_filter_plane_seq(uint8_t *pix, int ystride, int nhfrags, int nvfrags,
                                          const uint8_t *coded, int limit) {
  int fx, fy;
  if (limit == 0) return;
  for (fy = 0; fy < nvfrags; fy++) {
    for (fx = 0; fx < nhfrags; fx++) {
      uint8_t *p;
      if (!coded[fy * nhfrags + fx]) continue;
      p = pix + (ptrdiff_t)fy * 8 * ystride + fx * 8;
      if (fx > 0) lf_vedge(p, ystride, limit);
      if (fy > 0) lf_hedge(p, ystride, limit);
      if (fx + 1 < nhfrags && !coded[fy * nhfrags + fx + 1]) lf_vedge(p + 8, ystride, limit);
      if (fy + 1 < nvfrags && !coded[(fy + 1) * nhfrags + fx]) lf_hedge(p + 8 * (ptrdiff_t)ystride, ystride, limit);
    }
  }
}

/* Order-free form.  Every filter line lies inside exactly one 8x8 "cell"
   centred on a fragment corner (pixel columns [8cx-4,8cx+4), rows
   [8cy-4,8cy+4)).  Lines that do not cross the central 4x4 patch touch pixels
   no other line touches; the (up to) eight lines inside the patch are applied
   in the order the raster scan of state.c:1083-1104 would reach them:
     A=(cx-1,cy-1) B=(cx,cy-1) C=(cx-1,cy) D=(cx,cy)
     Vd: vertical edge A|B   Hl: horizontal edge A/C
     Vu: vertical edge C|D   Hr: horizontal edge B/D
     1 Vd if !B&&A   2 Hl if !C&&A   3 Vd if B   4 Hr if !D&&B
     5 Hl if C       6 Vu if !D&&C   7 Vu if D   8 Hr if D            */
static int cell_coded(const uint8_t *coded, int nh, int nv, int fx, int fy) {
  if (fx < 0 || fy < 0 || fx >= nh || fy >= nv) return 0;
  return coded[fy * nh + fx] != 0;
}

OCO_EXPORT void oco_loop_filter_plane_cells(uint8_t *pix, int ystride, int nhfrags, int nvfrags,
                                            const uint8_t *coded, int limit) {
  int cx, cy, k;
  if (limit == 0) return;
  for (cy = 0; cy <= nvfrags; cy++) {
    for (cx = 0; cx <= nhfrags; cx++) {
      int a = cell_coded(coded, nhfrags, nvfrags, cx - 1, cy - 1);
      int b = cell_coded(coded, nhfrags, nvfrags, cx, cy - 1);
      int c = cell_coded(coded, nhfrags, nvfrags, cx - 1, cy);
      int d = cell_coded(coded, nhfrags, nvfrags, cx, cy);
      /* which of the four edges meeting here exist and are filtered */
      int vd = cx > 0 && cx < nhfrags && cy > 0 && (a || b);
      int vu = cx > 0 && cx < nhfrags && cy < nvfrags && (c || d);
      int hl = cy > 0 && cy < nvfrags && cx > 0 && (a || c);
      int hr = cy > 0 && cy < nvfrags && cx < nhfrags && (b || d);
      uint8_t *o = pix + (ptrdiff_t)cy * 8 * ystride + cx * 8; /* corner pixel (8cx,8cy) */
      int slot;
      /* independent lines: rows -4,-3 of Vd, rows 2,3 of Vu, cols -4,-3 of Hl, cols 2,3 of Hr */
      if (vd) for (k = -4; k < -2; k++) lf_line(o - 2 + (ptrdiff_t)k * ystride, 1, limit);
      if (vu) for (k = 2; k < 4; k++) lf_line(o - 2 + (ptrdiff_t)k * ystride, 1, limit);
      if (hl) for (k = -4; k < -2; k++) lf_line(o + k - 2 * (ptrdiff_t)ystride, ystride, limit);
      if (hr) for (k = 2; k < 4; k++) lf_line(o + k - 2 * (ptrdiff_t)ystride, ystride, limit);
      /* ordered lines inside the central patch */
      for (slot = 1; slot <= 8; slot++) {
        int run_vd = (slot == 1 && vd && !b && a) || (slot == 3 && vd && b);
        int run_hl = (slot == 2 && hl && !c && a) || (slot == 5 && hl && c);
        int run_hr = (slot == 4 && hr && !d && b) || (slot == 8 && hr && d);
        int run_vu = (slot == 6 && vu && !d && c) || (slot == 7 && vu && d);
        if (run_vd) for (k = -2; k < 0; k++) lf_line(o - 2 + (ptrdiff_t)k * ystride, 1, limit);
        if (run_vu) for (k = 0; k < 2; k++) lf_line(o - 2 + (ptrdiff_t)k * ystride, 1, limit);
        if (run_hl) for (k = -2; k < 0; k++) lf_line(o + k - 2 * (ptrdiff_t)ystride, ystride, limit);
        if (run_hr) for (k = 0; k < 2; k++) lf_line(o + k - 2 * (ptrdiff_t)ystride, ystride, limit);
      }
    }
  }
}

/* state.c:770-835: left/right replication per row, then top/bottom rows
   (full padded width).  `pix` is the bottom-left pixel, ystride negative. */
OCO_EXPORT void oco_borders_fill_plane(uint8_t *pix, int ystride, int width, int height, int hpad, int vpad) {
  int y;
  for (y = 0; y < height; y++) {
    uint8_t *row = pix + (ptrdiff_t)y * ystride;
    memset(row - hpad, row[0], (size_t)hpad);
    memset(row + width, row[width - 1], (size_t)hpad);
  }
  for (y = 1; y <= vpad; y++) {
    memcpy(pix - hpad - (ptrdiff_t)y * ystride, pix - hpad, (size_t)(width + 2 * hpad));
    memcpy(pix - hpad + (ptrdiff_t)(height - 1 + y) * ystride,
           pix - hpad + (ptrdiff_t)(height - 1) * ystride, (size_t)(width + 2 * hpad));
  }
}

/* ---------------------------------------------------------------------- */
/* state.c:424-470 (fragment planes) and 545-671 (padded buffers, flip,
   frag_buf_offs). */
OCO_EXPORT int oco_geometry_init(ocg_geometry *g, int fw, int fh, int pixel_fmt, int nrefs) {
  int hdec = !(pixel_fmt & 1), vdec = !(pixel_fmt & 2);
  int64_t ystr = fw + 32, yrows = fh + 32;
  int64_t cstr = ((ystr >> hdec) + 15) & ~(int64_t)15, crows = yrows >> vdec;
  int64_t ysz = ystr * yrows, csz = cstr * crows;
  int64_t yoff = 16 + 16 * ystr;
  int64_t coff = (16 >> hdec) + (16 >> vdec) * cstr;
  int64_t align = (-coff) & 15;
  int64_t top_left[3];
  int pli, fro = 0;
  if ((fw & 15) || (fh & 15) || fw <= 0 || fh <= 0 || pixel_fmt == 1 || pixel_fmt < 0 || pixel_fmt > 3 ||
      nrefs < 3 || nrefs > 6)
    return OCG_EINVAL;
  memset(g, 0, sizeof(*g));
  g->frame_width = fw;
  g->frame_height = fh;
  g->pixel_fmt = pixel_fmt;
  g->nrefs = nrefs;
  g->ref_frame_sz = ysz + 2 * csz + 16;
  top_left[0] = yoff;
  top_left[1] = ysz + align + coff;
  top_left[2] = ysz + align + csz + coff;
  g->base_off = yoff + (int64_t)(fh - 1) * ystr;
  for (pli = 0; pli < 3; pli++) {
    ocg_plane_geom *p = &g->planes[pli];
    int w = pli ? fw >> hdec : fw, h = pli ? fh >> vdec : fh;
    int64_t str = pli ? cstr : ystr;
    p->width = w;
    p->height = h;
    p->nhfrags = w >> 3;
    p->nvfrags = h >> 3;
    p->froffset = fro;
    p->nfrags = p->nhfrags * p->nvfrags;
    p->ystride = (int32_t)-str;
    p->hpad = pli ? 16 >> hdec : 16;
    p->vpad = pli ? 16 >> vdec : 16;
    p->plane_off = top_left[pli] + (int64_t)(h - 1) * str - g->base_off;
    fro += p->nfrags;
  }
  g->nfrags = fro;
  return 0;
}

OCO_EXPORT void oco_geometry_frag_buf_offs(const ocg_geometry *g, int32_t *offs) {
  int pli, fx, fy;
  for (pli = 0; pli < 3; pli++) {
    const ocg_plane_geom *p = &g->planes[pli];
    for (fy = 0; fy < p->nvfrags; fy++)
      for (fx = 0; fx < p->nhfrags; fx++)
        offs[p->froffset + fy * p->nhfrags + fx] =
            (int32_t)(p->plane_off + (int64_t)fy * 8 * p->ystride + fx * 8);
  }
}

/* decode.c:1392-1500 (oc_dec_dc_unpredict_mcu_plane_c) over a whole plane:
   refs[i] = reference type of fragment i (0..2) or OCG_FRAG_UNCODED, dc[i] in:
   residual, out: value.  pred_last starts at 0 for every type (decode.c:1367). */
OCO_EXPORT void oco_dc_unpredict_plane(int16_t *dc, const uint8_t *refs, int nhfrags, int nvfrags) {
  int pred_last[3] = {0, 0, 0};
  int fragx, fragy, fragi = 0;
  for (fragy = 0; fragy < nvfrags; fragy++) {
    if (fragy == 0) {
      for (fragx = 0; fragx < nhfrags; fragx++, fragi++) {
        if (refs[fragi] != OCG_FRAG_UNCODED) {
          int refi = refs[fragi];
          dc[fragi] = (int16_t)(dc[fragi] + pred_last[refi]);
          pred_last[refi] = dc[fragi];
        }
      }
    } else {
      const int16_t *u_dc = dc - nhfrags;
      const uint8_t *u_refs = refs - nhfrags;
      int l_ref = -1, ul_ref = -1, u_ref = u_refs[fragi] == OCG_FRAG_UNCODED ? -1 : u_refs[fragi];
      for (fragx = 0; fragx < nhfrags; fragx++, fragi++) {
        int ur_ref;
        if (fragx + 1 >= nhfrags) ur_ref = -1;
        else ur_ref = u_refs[fragi + 1] == OCG_FRAG_UNCODED ? -1 : u_refs[fragi + 1];
        if (refs[fragi] != OCG_FRAG_UNCODED) {
          int pred, refi = refs[fragi];
          switch ((l_ref == refi) | (ul_ref == refi) << 1 | (u_ref == refi) << 2 | (ur_ref == refi) << 3) {
            default: pred = pred_last[refi]; break;
            case 1: case 3: pred = dc[fragi - 1]; break;
            case 2: pred = u_dc[fragi - 1]; break;
            case 4: case 6: case 12: pred = u_dc[fragi]; break;
            case 5: pred = (dc[fragi - 1] + u_dc[fragi]) / 2; break;
            case 8: pred = u_dc[fragi + 1]; break;
            case 9: case 11: case 13: pred = (75 * dc[fragi - 1] + 53 * u_dc[fragi + 1]) / 128; break;
            case 10: pred = (u_dc[fragi - 1] + u_dc[fragi + 1]) / 2; break;
            case 14: pred = (3 * (u_dc[fragi - 1] + u_dc[fragi + 1]) + 10 * u_dc[fragi]) / 16; break;
            case 7: case 15: {
              int p0 = dc[fragi - 1], p1 = u_dc[fragi - 1], p2 = u_dc[fragi];
              pred = (29 * (p0 + p2) - 26 * p1) / 32;
              if (abs(pred - p2) > 128) pred = p2;
              else if (abs(pred - p0) > 128) pred = p0;
              else if (abs(pred - p1) > 128) pred = p1;
            } break;
          }
          dc[fragi] = (int16_t)(dc[fragi] + pred);
          pred_last[refi] = dc[fragi];
          l_ref = refi;
        } else l_ref = -1;
        ul_ref = u_ref;
        u_ref = ur_ref;
      }
    }
  }
}

/* The whole-frame sequence of decode.c:2858-2945 driven from the C-ABI frame
   description: recon of coded fragments (decode.c:1584 -> state.c:959), copy
   of uncoded ones (decode.c:1599), loop filter over all rows (2882), borders
   (2890, 2945).  Fragments are visited in fragment-index order; coded ones
   write only their own 8x8 block of SELF and read other buffers, so the
   reference's coded order gives the same pixels. */
OCO_EXPORT void oco_dec_frame(const ocg_geometry *g, uint8_t *frames, const ocg_dec_frame *f, int stage_mask) {
  uint8_t *base[3];
  uint8_t *coded = (uint8_t *)malloc((size_t)g->nfrags);
  int16_t *dcv = NULL;
  int i, r, c, pli;
  for (i = 0; i < 3; i++)
    base[i] = f->ref_idx[i] >= 0 ? frames + (int64_t)f->ref_idx[i] * g->ref_frame_sz + g->base_off : NULL;
  for (i = 0; i < g->nfrags; i++) coded[i] = f->recs[i].refi != OCG_FRAG_UNCODED;
  if (f->dc_residual && (stage_mask & 1)) {
    /* the records carry DC residuals: undo the prediction first (decode.c:2876) */
    uint8_t *refs = (uint8_t *)malloc((size_t)g->nfrags);
    dcv = (int16_t *)malloc((size_t)g->nfrags * sizeof(int16_t));
    for (i = 0; i < g->nfrags; i++) { refs[i] = f->recs[i].refi; dcv[i] = f->recs[i].dc; }
    for (pli = 0; pli < 3; pli++)
      oco_dc_unpredict_plane(dcv + g->planes[pli].froffset, refs + g->planes[pli].froffset, g->planes[pli].nhfrags,
                             g->planes[pli].nvfrags);
    free(refs);
  }
  if (stage_mask & 1) {
    for (i = 0; i < g->nfrags; i++) {
      const ocg_frag_rec *rec = &f->recs[i];
      int pl = rec->pli_qti & 3, qti = rec->pli_qti >> 2 & 1;
      if (!coded[i]) {
        oco_frag_copy(base[OCG_FRAME_SELF] + rec->buf_off, base[OCG_FRAME_PREV] + rec->buf_off, g->planes[pl].ystride);
      } else {
        int16_t blk[128];
        const int16_t *rows = f->coeff_rows + (size_t)rec->coeff_row * 8;
        memset(blk, 0, sizeof(blk));
        for (r = 0; r < 8; r++) {
          if (!(rec->rowmask >> r & 1)) continue;
          for (c = 0; c < 8; c++) blk[r * 8 + c] = rows[c];
          rows += 8;
        }
        blk[0] = dcv ? dcv[i] : rec->dc;
        oco_state_frag_recon(base[OCG_FRAME_SELF], rec->refi == OCG_FRAME_SELF ? NULL : base[rec->refi],
                             rec->buf_off, g->planes[pl].ystride, pl, g->pixel_fmt,
                             rec->refi == OCG_FRAME_SELF, rec->mv, blk, rec->last_zzi, f->dc_quant[pl][qti]);
      }
    }
  }
  for (pli = 0; pli < 3; pli++) {
    const ocg_plane_geom *p = &g->planes[pli];
    uint8_t *pix = base[OCG_FRAME_SELF] + p->plane_off;
    if ((stage_mask & 2) && f->lf_limit)
      oco_loop_filter_plane_seq(pix, p->ystride, p->nhfrags, p->nvfrags, coded + p->froffset, f->lf_limit);
    if (stage_mask & 4) oco_borders_fill_plane(pix, p->ystride, p->width, p->height, p->hpad, p->vpad);
  }
  free(coded);
  free(dcv);
}

/* ---------------------------------------------------------------------- */
/* encoder block kernels */

/* fdct.c:28-120 (oc_fdct8): input every `istep`-th sample, output 8 in a row. */
static inline int fd_exp(int t, int bias) { return ((27146 * t + bias) >> 16) + t + (t != 0); }

static void fdct8_pass(int16_t out[8], const int16_t *in, int istep) {
  int x0 = in[0], x1 = in[istep], x2 = in[2 * istep], x3 = in[3 * istep];
  int x4 = in[4 * istep], x5 = in[5 * istep], x6 = in[6 * istep], x7 = in[7 * istep];
  /* stage 1 */
  int a0 = x0 + x7, a7 = x0 - x7, a1 = x1 + x6, a6 = x1 - x6;
  int a2 = x2 + x5, a5 = x2 - x5, a3 = x3 + x4, a4 = x3 - x4;
  /* stage 2 */
  int b0 = a0 + a3, b3 = a0 - a3, b1 = a1 + a2, b2 = a1 - a2;
  int b6 = a6 + a5, b5 = a6 - a5;
  /* stage 3 */
  int s = fd_exp(b5, 0xB500) >> 1;
  int c4 = a4 + s, c5 = a4 - s;
  int c7, c6, r, u, v;
  s = fd_exp(b6, 0xB500) >> 1;
  c7 = a7 + s;
  c6 = a7 - s;
  /* stage 4 */
  r = fd_exp(b0, 0x4000);
  s = fd_exp(b1, 0xB500);
  u = (r + s) >> 1;
  out[0] = (int16_t)u;
  out[4] = (int16_t)(r - u);
  u = ((K6 * b2 + K2 * b3 + 0x6CB7) >> 16) + (b3 != 0);
  s = (K6 * u >> 16) - b2;
  v = ((s * 21600 + 0x2800) >> 18) + s + (s != 0);
  out[2] = (int16_t)u;
  out[6] = (int16_t)v;
  u = ((K5 * c6 + K3 * c5 + 0x0E3D) >> 16) + (c5 != 0);
  s = c6 - (K5 * u >> 16);
  v = ((s * 26568 + 0x3400) >> 17) + s + (s != 0);
  out[5] = (int16_t)u;
  out[3] = (int16_t)v;
  u = ((K7 * c4 + K1 * c7 + 0x7B1B) >> 16) + (c7 != 0);
  s = (K7 * u >> 16) - c4;
  v = ((s * 20539 + 0x3000) >> 20) + s + (s != 0);
  out[1] = (int16_t)u;
  out[7] = (int16_t)v;
}

/* zig-zag position -> natural index (internal.c:27-44, first 64 entries);
   generated, not tabulated: walk the anti-diagonals. */
static void zigzag_table(uint8_t fz[64]) {
  int d, k = 0;
  for (d = 0; d < 15; d++) {
    int lo = d < 8 ? 0 : d - 7, hi = d < 8 ? d : 7, i;
    for (i = lo; i <= hi; i++) {
      /* odd diagonals run top-right -> bottom-left, even ones the other way */
      int r = (d & 1) ? i : d - i;
      fz[k++] = (uint8_t)(r * 8 + (d - r));
    }
  }
}

/* fdct.c:128-150 */
OCO_EXPORT void oco_fdct8x8(int16_t y[64], const int16_t x[64]) {
  static uint8_t fz[64];
  static int fz_ready = 0;
  int16_t w[64], t[64];
  int i;
  if (!fz_ready) { zigzag_table(fz); fz_ready = 1; }
  for (i = 0; i < 64; i++) w[i] = (int16_t)(x[i] << 2);
  w[0] = (int16_t)(w[0] + (w[0] != 0) + 1);
  w[1]++;
  w[8]--;
  for (i = 0; i < 8; i++) fdct8_pass(t + i * 8, w + i, 8);
  for (i = 0; i < 8; i++) fdct8_pass(w + i * 8, t + i, 8);
  for (i = 0; i < 64; i++) y[i] = (int16_t)((w[fz[i]] + 2) >> 2);
}

/* enquant.c:184-208: per coefficient {m,l} with x/d == ((x*m>>16)+x>>l)+(x<0). */
OCO_EXPORT void oco_enquant_init(int16_t enq[128], const uint16_t dequant[64]) {
  int zzi;
  for (zzi = 0; zzi < 64; zzi++) {
    uint32_t d = (uint32_t)dequant[zzi] << 1;
    int l = 31 - __builtin_clz(d);
    uint32_t t = 1 + ((uint32_t)1 << (16 + l)) / d;
    enq[2 * zzi] = (int16_t)(t - 0x10000);
    enq[2 * zzi + 1] = (int16_t)l;
  }
}

/* enquant.c:220-249 */
OCO_EXPORT int oco_quantize(int16_t q[64], const int16_t dct[64], const uint16_t dequant[64],
                            const int16_t enq[128]) {
  int zzi, last = 0;
  for (zzi = 0; zzi < 64; zzi++) {
    int v = dct[zzi] << 1, d = dequant[zzi];
    if (abs(v) >= d) {
      int s = v < 0 ? -1 : 0;
      v += (d + s) ^ s;
      v = (((enq[2 * zzi] * (int32_t)v >> 16) + v) >> enq[2 * zzi + 1]) - s;
      q[zzi] = (int16_t)v;
      last = zzi;
    } else q[zzi] = 0;
  }
  return last;
}

OCO_EXPORT void oco_frag_sub(int16_t d[64], const uint8_t *src, const uint8_t *ref, int ystride) {
  int r, c;
  for (r = 0; r < 8; r++, src += ystride, ref += ystride)
    for (c = 0; c < 8; c++) d[r * 8 + c] = (int16_t)(src[c] - ref[c]);
}

OCO_EXPORT void oco_frag_sub_128(int16_t d[64], const uint8_t *src, int ystride) {
  int r, c;
  for (r = 0; r < 8; r++, src += ystride)
    for (c = 0; c < 8; c++) d[r * 8 + c] = (int16_t)(src[c] - 128);
}

OCO_EXPORT unsigned oco_frag_sad(const uint8_t *src, const uint8_t *ref, int ystride) {
  unsigned s = 0;
  int r, c;
  for (r = 0; r < 8; r++, src += ystride, ref += ystride)
    for (c = 0; c < 8; c++) s += (unsigned)abs(src[c] - ref[c]);
  return s;
}

/* encfrag.c:56-69: row-granular early out once the running sum exceeds thresh. */
OCO_EXPORT unsigned oco_frag_sad_thresh(const uint8_t *src, const uint8_t *ref, int ystride, unsigned thresh) {
  unsigned s = 0;
  int r, c;
  for (r = 0; r < 8; r++, src += ystride, ref += ystride) {
    for (c = 0; c < 8; c++) s += (unsigned)abs(src[c] - ref[c]);
    if (s > thresh) break;
  }
  return s;
}

OCO_EXPORT unsigned oco_frag_sad2_thresh(const uint8_t *src, const uint8_t *r1, const uint8_t *r2, int ystride,
                                         unsigned thresh) {
  unsigned s = 0;
  int r, c;
  for (r = 0; r < 8; r++, src += ystride, r1 += ystride, r2 += ystride) {
    for (c = 0; c < 8; c++) s += (unsigned)abs(src[c] - ((r1[c] + r2[c]) >> 1));
    if (s > thresh) break;
  }
  return s;
}

OCO_EXPORT unsigned oco_frag_intra_sad(const uint8_t *src, int ystride) {
  const uint8_t *p = src;
  unsigned s = 0;
  int r, c, dc = 0;
  for (r = 0; r < 8; r++, p += ystride)
    for (c = 0; c < 8; c++) dc += p[c];
  dc = (dc + 32) >> 6;
  for (r = 0, p = src; r < 8; r++, p += ystride)
    for (c = 0; c < 8; c++) s += (unsigned)abs(p[c] - dc);
  return s;
}

/* 8-point Hadamard butterfly network shared by encfrag.c:109-304. */
static void hadamard8(int t[8]) {
  int a0 = t[0] + t[4], a4 = t[0] - t[4], a1 = t[1] + t[5], a5 = t[1] - t[5];
  int a2 = t[2] + t[6], a6 = t[2] - t[6], a3 = t[3] + t[7], a7 = t[3] - t[7];
  int b0 = a0 + a2, b2 = a0 - a2, b1 = a1 + a3, b3 = a1 - a3;
  int b4 = a4 + a6, b6 = a4 - a6, b5 = a5 + a7, b7 = a5 - a7;
  t[0] = b0 + b1; t[1] = b0 - b1; t[2] = b2 + b3; t[3] = b2 - b3;
  t[4] = b4 + b5; t[5] = b4 - b5; t[6] = b6 + b7; t[7] = b6 - b7;
}

/* encfrag.c:262-304 (oc_hadamard_sad) applied to the row-transformed, 16-bit
   truncated, transposed buffer produced by encfrag.c:109-260. */
static unsigned hadamard_sad(int *dc, int16_t buf[64]) {
  unsigned sad = 0;
  int i, k, t[8];
  for (i = 0; i < 8; i++) {
    for (k = 0; k < 8; k++) t[k] = buf[i * 8 + k];
    hadamard8(t);
    for (k = 0; k < 8; k++) if (i > 0 || k > 0) sad += (unsigned)abs(t[k]);
  }
  *dc = buf[0] + buf[1] + buf[2] + buf[3] + buf[4] + buf[5] + buf[6] + buf[7];
  return sad;
}

static unsigned satd_core(int *dc, const uint8_t *src, const uint8_t *r1, const uint8_t *r2, int ystride) {
  int16_t buf[64];
  int r, c, t[8];
  for (r = 0; r < 8; r++) {
    for (c = 0; c < 8; c++) {
      int pred = r1 == NULL ? 0 : (r2 == NULL ? r1[c] : (r1[c] + r2[c]) >> 1);
      t[c] = src[c] - pred;
    }
    hadamard8(t);
    for (c = 0; c < 8; c++) buf[c * 8 + r] = (int16_t)t[c];
    src += ystride;
    if (r1) r1 += ystride;
    if (r2) r2 += ystride;
  }
  return hadamard_sad(dc, buf);
}

OCO_EXPORT unsigned oco_frag_satd(int *dc, const uint8_t *src, const uint8_t *ref, int ystride) {
  return satd_core(dc, src, ref, NULL, ystride);
}
OCO_EXPORT unsigned oco_frag_satd2(int *dc, const uint8_t *src, const uint8_t *r1, const uint8_t *r2, int ystride) {
  return satd_core(dc, src, r1, r2, ystride);
}
OCO_EXPORT unsigned oco_frag_intra_satd(int *dc, const uint8_t *src, int ystride) {
  return satd_core(dc, src, NULL, NULL, ystride);
}

OCO_EXPORT unsigned oco_frag_ssd(const uint8_t *src, const uint8_t *ref, int ystride) {
  unsigned s = 0;
  int r, c;
  for (r = 0; r < 8; r++, src += ystride, ref += ystride)
    for (c = 0; c < 8; c++) s += (unsigned)((src[c] - ref[c]) * (src[c] - ref[c]));
  return s;
}

OCO_EXPORT unsigned oco_frag_border_ssd(const uint8_t *src, const uint8_t *ref, int ystride, int64_t mask) {
  unsigned s = 0;
  int r, c;
  for (r = 0; r < 8; r++, src += ystride, ref += ystride)
    for (c = 0; c < 8; c++)
      if ((uint64_t)mask >> (r * 8 + c) & 1) s += (unsigned)((src[c] - ref[c]) * (src[c] - ref[c]));
  return s;
}

OCO_EXPORT void oco_frag_copy2(uint8_t *dst, const uint8_t *s1, const uint8_t *s2, int ystride) {
  int r, c;
  for (r = 0; r < 8; r++, dst += ystride, s1 += ystride, s2 += ystride)
    for (c = 0; c < 8; c++) dst[c] = (uint8_t)((s1[c] + s2[c]) >> 1);
}

/* ---- batch forms mirroring the C-ABI ---- */

/* mathops.c:294-313 */
static uint32_t oco_bexp32_q10(int z) {
  int ipart = z >> 10;
  unsigned n = (unsigned)(z & 1023) << 4;
  n = (n * ((n * ((n * ((n * 3548u >> 15) + 6817u) >> 15) + 15823u) >> 15) + 22708u) >> 15) + 16384u;
  return 14 - ipart > 0 ? (n + (1u << (13 - ipart))) >> (14 - ipart) : n << (ipart - 14);
}
static int oco_blog32_q10(uint32_t w) {
  int ipart = 0, n, fpart;
  uint32_t v = w;
  if (w == 0) return -1;
  while (v) { ipart++; v >>= 1; }
  n = (int)(ipart - 16 > 0 ? w >> (ipart - 16) : w << (16 - ipart)) - 32768 - 16384;
  fpart = (n * ((n * ((n * ((n * -1402 >> 15) + 2546) >> 15) - 5216) >> 15) + 15745) >> 15) - 6793;
  return (ipart << 10) + (fpart >> 4);
}

/* analyze.c:1167-1234, one luma block: returns the activity, *sum = pixel sum */
OCO_EXPORT unsigned oco_block_activity(const uint8_t *src, int ystride, int *sum) {
  const uint8_t *s = src;
  unsigned x = 0, x2 = 0, act;
  int i, j;
  for (i = 0; i < 8; i++) {
    for (j = 0; j < 8; j++) { unsigned c = s[j]; x += c; x2 += c * c; }
    s += ystride;
  }
  if (sum) *sum = (int)x;
  act = (x2 << 6) - x * x;
  if (act < 8u << 12) return act < (5u << 12) ? act : (5u << 12);
  {
    unsigned e1 = 0, e2 = 0, e3 = 0, e4 = 0, emax;
    s = src - 1;
    for (i = 0; i < 8; i++) {
      const uint8_t *u = s - ystride, *d = s + ystride;
      for (j = 0; j < 8; j++) {
        e1 += (unsigned)abs(((s[j + 2] - s[j]) << 1) + u[j + 2] - u[j] + d[j + 2] - d[j]);
        e2 += (unsigned)abs(((d[j + 1] - u[j + 1]) << 1) + d[j] - u[j] + d[j + 2] - u[j + 2]);
        e3 += (unsigned)abs(((d[j + 2] - u[j]) << 1) + d[j + 1] - s[j] + s[j + 2] - u[j + 1]);
        e4 += (unsigned)abs(((d[j] - u[j + 2]) << 1) + d[j + 1] - s[j + 2] + s[j] - u[j + 1]);
      }
      s += ystride;
    }
    emax = e1 > e2 ? e1 : e2;
    if (e3 > emax) emax = e3;
    if (e4 > emax) emax = e4;
    if (5 * emax > 2 * (e1 + e2 + e3 + e4)) act = oco_bexp32_q10(0x394A + (7 * (oco_blog32_q10(act) - 0x394A + 5) / 10));
  }
  return act;
}

OCO_EXPORT void oco_enc_metrics_batch(int metric, const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                                      const ocg_enc_frag *frags, int n, uint32_t *out_val, int32_t *out_dc) {
  int i;
  for (i = 0; i < n; i++) {
    const uint8_t *src = src_base + frags[i].src_off;
    const uint8_t *r0 = frags[i].ref_off0 == INT32_MIN ? NULL : ref_base + frags[i].ref_off0;
    const uint8_t *r1 = frags[i].ref_off1 == INT32_MIN ? NULL : ref_base + frags[i].ref_off1;
    int dc = 0;
    unsigned v = 0;
    switch (metric) {
      case OCG_MET_SAD: v = r1 ? oco_frag_sad2_thresh(src, r0, r1, ystride, UINT_MAX) : oco_frag_sad(src, r0, ystride); break;
      case OCG_MET_SATD: v = r1 ? oco_frag_satd2(&dc, src, r0, r1, ystride) : oco_frag_satd(&dc, src, r0, ystride); break;
      case OCG_MET_INTRA_SATD: v = oco_frag_intra_satd(&dc, src, ystride); break;
      case OCG_MET_SSD: v = oco_frag_ssd(src, r0, ystride); break;
      case OCG_MET_INTRA_SAD: v = oco_frag_intra_sad(src, ystride); break;
      case OCG_MET_BORDER_SSD:
        v = oco_frag_border_ssd(src, r0, ystride,
                                (int64_t)(((uint64_t)(uint32_t)frags[i].aux << 32) | (uint32_t)frags[i].ref_off1));
        break;
      case OCG_MET_ACTIVITY: v = oco_block_activity(src, ystride, &dc); break;
      case OCG_MET_SAD_THRESH:
        v = r1 ? oco_frag_sad2_thresh(src, r0, r1, ystride, (unsigned)frags[i].aux)
               : oco_frag_sad_thresh(src, r0, ystride, (unsigned)frags[i].aux);
        break;
      default: break;
    }
    out_val[i] = v;
    if (out_dc) out_dc[i] = dc;
  }
}

/* analyze.c:725-778: sub / sub_128 / copy2+sub, fDCT, quantise. */
OCO_EXPORT void oco_enc_fdct_quant_batch(const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                                         const ocg_enc_frag *frags, int n, const uint16_t *dequant,
                                         const int16_t *enquant, int16_t *dct, int16_t *qdct, int32_t *nonzero) {
  int i;
  for (i = 0; i < n; i++) {
    const uint8_t *src = src_base + frags[i].src_off;
    int pli = frags[i].aux & 3, qti = frags[i].aux >> 2 & 1, qii = frags[i].aux >> 3 & 3;
    int tab = ((pli * 2 + qti) * 3 + qii);
    int16_t diff[64];
    if (frags[i].ref_off0 == INT32_MIN) oco_frag_sub_128(diff, src, ystride);
    else if (frags[i].ref_off1 == INT32_MIN) oco_frag_sub(diff, src, ref_base + frags[i].ref_off0, ystride);
    else {
      const uint8_t *r0 = ref_base + frags[i].ref_off0, *r1 = ref_base + frags[i].ref_off1;
      int r, c;
      for (r = 0; r < 8; r++)
        for (c = 0; c < 8; c++)
          diff[r * 8 + c] = (int16_t)(src[(ptrdiff_t)r * ystride + c] -
                                      ((r0[(ptrdiff_t)r * ystride + c] + r1[(ptrdiff_t)r * ystride + c]) >> 1));
    }
    oco_fdct8x8(dct + (size_t)i * 64, diff);
    nonzero[i] = oco_quantize(qdct + (size_t)i * 64, dct + (size_t)i * 64, dequant + (size_t)tab * 64,
                              enquant + (size_t)tab * 128);
  }
}

/* ---------------------------------------------------------------------- */
/* mcenc.c:268-515 with the candidate lists supplied by the caller. */
typedef struct {
  const uint8_t *src, *ref;
  int ystride;
  const int32_t *frag_off;
  uint32_t hit[31]; /* visited full-pel vectors, mcenc.c:292 */
  unsigned best_err;
  int best[2];
  unsigned blk_err[4];
  int blk_vec[4][2];
  int track_blocks;
} mcs_state;

/* mcenc.c:200-220: 16x16 SAD as four block SADs */
static unsigned mcs_sad16(const mcs_state *s, int dx, int dy, unsigned berr[4]) {
  unsigned err = 0;
  int bi;
  for (bi = 0; bi < 4; bi++) {
    berr[bi] = oco_frag_sad(s->src + s->frag_off[bi], s->ref + s->frag_off[bi] + dx + dy * s->ystride, s->ystride);
    err += berr[bi];
  }
  return err;
}

static int mcs_visited(mcs_state *s, int dx, int dy) {
  uint32_t bit = (uint32_t)1 << (dx + 15);
  if (s->hit[dy + 15] & bit) return 1;
  s->hit[dy + 15] |= bit;
  return 0;
}

static void mcs_track_blocks(mcs_state *s, int dx, int dy, const unsigned berr[4]) {
  int bi;
  if (!s->track_blocks) return;
  for (bi = 0; bi < 4; bi++)
    if (berr[bi] < s->blk_err[bi]) {
      s->blk_err[bi] = berr[bi];
      s->blk_vec[bi][0] = dx;
      s->blk_vec[bi][1] = dy;
    }
}

/* sites of the 3x3 square pattern allowed at the +-15 window boundary
   (mcenc.c:50-87 tabulates the same thing) */
static int mcs_sites(int x, int y, int sites[8][2]) {
  int n = 0, dx, dy;
  for (dy = -1; dy <= 1; dy++)
    for (dx = -1; dx <= 1; dx++) {
      if (dx == 0 && dy == 0) continue;
      if ((x <= -15 && dx < 0) || (x >= 15 && dx > 0) || (y <= -15 && dy < 0) || (y >= 15 && dy > 0)) continue;
      sites[n][0] = dx;
      sites[n][1] = dy;
      n++;
    }
  return n;
}

static int div2_trunc(int v) { return v / 2; }

OCO_EXPORT void oco_mcenc_search_batch(const uint8_t *src_base, const uint8_t *ref_full_base,
                                       const uint8_t *ref_satd_base, int ystride, const ocg_mb_search_in *in,
                                       ocg_mb_search_out *out, int n) {
  int i;
  for (i = 0; i < n; i++) {
    const ocg_mb_search_in *m = &in[i];
    mcs_state s;
    unsigned berr[4], err, t2;
    int bi, ci, cx, cy;
    memset(&s, 0, sizeof(s));
    s.src = src_base;
    s.ref = ref_full_base;
    s.ystride = ystride;
    s.frag_off = m->frag_off;
    s.track_blocks = m->is_prev != 0;
    /* median predictor first (mcenc.c:295-316) */
    cx = div2_trunc(m->cand[0][0]);
    cy = div2_trunc(m->cand[0][1]);
    mcs_visited(&s, cx, cy);
    s.best_err = mcs_sad16(&s, cx, cy, berr);
    s.best[0] = cx;
    s.best[1] = cy;
    for (bi = 0; bi < 4; bi++) {
      s.blk_err[bi] = berr[bi];
      s.blk_vec[bi][0] = cx;
      s.blk_vec[bi][1] = cy;
    }
    if (s.best_err > 256) { /* OC_YSAD_THRESH1 */
      t2 = m->t2_base;
      t2 += (t2 >> 4) + 64; /* OC_YSAD_THRESH2_SCALE_BITS / _OFFSET */
      for (ci = 1; ci < m->setb0; ci++) { /* set A */
        cx = div2_trunc(m->cand[ci][0]);
        cy = div2_trunc(m->cand[ci][1]);
        if (mcs_visited(&s, cx, cy)) continue;
        err = mcs_sad16(&s, cx, cy, berr);
        if (err < s.best_err) { s.best_err = err; s.best[0] = cx; s.best[1] = cy; }
        mcs_track_blocks(&s, cx, cy, berr);
      }
      if (s.best_err > t2) {
        for (; ci < m->ncand; ci++) { /* set B */
          cx = div2_trunc(m->cand[ci][0]);
          cy = div2_trunc(m->cand[ci][1]);
          if (mcs_visited(&s, cx, cy)) continue;
          err = mcs_sad16(&s, cx, cy, berr);
          if (err < s.best_err) { s.best_err = err; s.best[0] = cx; s.best[1] = cy; }
          mcs_track_blocks(&s, cx, cy, berr);
        }
        if (s.best_err > t2) {
          int sites[8][2], ns, si, moved;
          /* square-pattern descent around the macro-block vector (mcenc.c:399-431):
             the centre moves to the LAST site that improved on the running best */
          do {
            int step[2] = {0, 0};
            moved = 0;
            ns = mcs_sites(s.best[0], s.best[1], sites);
            for (si = 0; si < ns; si++) {
              cx = s.best[0] + sites[si][0];
              cy = s.best[1] + sites[si][1];
              if (mcs_visited(&s, cx, cy)) continue;
              err = mcs_sad16(&s, cx, cy, berr);
              if (err < s.best_err) { s.best_err = err; step[0] = sites[si][0]; step[1] = sites[si][1]; moved = 1; }
              mcs_track_blocks(&s, cx, cy, berr);
            }
            s.best[0] += step[0];
            s.best[1] += step[1];
          } while (moved);
          /* per-block descents sharing the hit cache (mcenc.c:437-499) */
          if (s.track_blocks) {
            unsigned t4 = t2 >> 2;
            for (bi = 0; bi < 4; bi++) {
              if (s.blk_err[bi] <= t4) continue;
              for (;;) {
                int bx = s.blk_vec[bi][0], by = s.blk_vec[bi][1], bj;
                ns = mcs_sites(bx, by, sites);
                for (si = 0; si < ns; si++) {
                  cx = bx + sites[si][0];
                  cy = by + sites[si][1];
                  if (mcs_visited(&s, cx, cy)) continue;
                  err = mcs_sad16(&s, cx, cy, berr);
                  if (err < s.best_err) { s.best_err = err; s.best[0] = cx; s.best[1] = cy; }
                  for (bj = 0; bj < 4; bj++)
                    if (berr[bj] < s.blk_err[bj]) {
                      s.blk_err[bj] = berr[bj];
                      s.blk_vec[bj][0] = cx;
                      s.blk_vec[bj][1] = cy;
                    }
                }
                if (s.blk_vec[bi][0] == bx && s.blk_vec[bi][1] == by) break;
              }
            }
          }
        }
      }
    }
    /* results, incl. the SATD of the winner on the reconstructed reference (mcenc.c:500-513) */
    {
      ocg_mb_search_out *o = &out[i];
      unsigned satd = 0;
      int dc;
      memset(o, 0, sizeof(*o));
      o->best_vec[0] = (int8_t)s.best[0];
      o->best_vec[1] = (int8_t)s.best[1];
      o->error = (uint16_t)s.best_err;
      for (bi = 0; bi < 4; bi++) {
        satd += oco_frag_satd(&dc, src_base + m->frag_off[bi],
                              ref_satd_base + m->frag_off[bi] + s.best[0] + s.best[1] * ystride, ystride);
        satd += (unsigned)abs(dc);
      }
      o->satd = satd;
      if (m->is_prev) {
        for (bi = 0; bi < 4; bi++) {
          unsigned bs = oco_frag_satd(&dc, src_base + m->frag_off[bi],
                                      ref_satd_base + m->frag_off[bi] + s.blk_vec[bi][0] + s.blk_vec[bi][1] * ystride,
                                      ystride);
          o->block_vec[bi][0] = (int8_t)s.blk_vec[bi][0];
          o->block_vec[bi][1] = (int8_t)s.blk_vec[bi][1];
          o->block_satd[bi] = bs + (unsigned)abs(dc);
        }
      }
    }
  }
}

/* ---- half-pel refinement (mcenc.c:606-791) --------------------------------
   The two taps of half-pel vector 2v+d are found as the reference does at
   mcenc.c:636-646 (a restatement of oc_state_get_mv_offsets for luma): the
   component whose half-pel value and step have opposite signs keeps the
   full-pel tap on the first offset. */
static void ref_taps(int vx, int vy, int dx, int dy, int ystride, int *o0, int *o1) {
  int hx = 2 * vx + dx, hy = 2 * vy + dy;
  int first_x = ((hx ^ dx) < 0) ? dx : 0, second_x = ((hx ^ dx) < 0) ? 0 : dx;
  int first_y = ((hy ^ dy) < 0) ? dy * ystride : 0, second_y = ((hy ^ dy) < 0) ? 0 : dy * ystride;
  int base = vx + vy * ystride;
  *o0 = base + first_x + first_y;
  *o1 = base + second_x + second_y;
}

/* visiting order of the 8 sites: OC_SQUARE_SITES[0] = {0,1,2,3,5,6,7,8} over
   OC_SQUARE_DX/DY (mcenc.c:50-66) */
static const int k_ref_dx[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
static const int k_ref_dy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};

OCO_EXPORT void oco_mcenc_refine_batch(const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                                       const ocg_mb_refine_in *in, ocg_mb_refine_out *out, int n, int flags) {
  int i, bi, si;
  for (i = 0; i < n; i++) {
    const ocg_mb_refine_in *m = &in[i];
    ocg_mb_refine_out *o = &out[i];
    if (flags & OCG_REFINE_1MV) {
      /* oc_mcenc_ysatd_halfpel_mbrefine, mcenc.c:606-659 */
      unsigned best = m->satd;
      int bdx = 0, bdy = 0;
      for (si = 0; si < 8; si++) {
        int o0, o1, dc;
        unsigned err = 0;
        ref_taps(m->vec[0], m->vec[1], k_ref_dx[si], k_ref_dy[si], ystride, &o0, &o1);
        for (bi = 0; bi < 4; bi++) {
          const uint8_t *s = src_base + m->frag_off[bi], *r = ref_base + m->frag_off[bi];
          if (flags & OCG_REFINE_SAD) {
            /* oc_sad16_halfpel, mcenc.c:166-180: the early-out only ever reports
               "worse than the current best", the exact sum decides the same */
            err += oco_frag_sad2_thresh(s, r + o0, r + o1, ystride, UINT_MAX);
          } else {
            err += oco_frag_satd2(&dc, s, r + o0, r + o1, ystride);
            err += (unsigned)abs(dc);
          }
        }
        if (err < best) { best = err; bdx = k_ref_dx[si]; bdy = k_ref_dy[si]; }
      }
      o->mv[0] = (int8_t)(2 * m->vec[0] + bdx);
      o->mv[1] = (int8_t)(2 * m->vec[1] + bdy);
      o->satd = best;
    }
    if (flags & OCG_REFINE_4MV) {
      /* oc_mcenc_refine4mv + oc_mcenc_ysatd_halfpel_brefine, mcenc.c:713-791 */
      for (bi = 0; bi < 4; bi++) {
        const uint8_t *s = src_base + m->frag_off[bi], *r = ref_base + m->frag_off[bi];
        unsigned best = m->block_satd[bi];
        int bdx = 0, bdy = 0;
        for (si = 0; si < 8; si++) {
          int o0, o1, dc;
          unsigned err;
          ref_taps(m->block_vec[bi][0], m->block_vec[bi][1], k_ref_dx[si], k_ref_dy[si], ystride, &o0, &o1);
          err = oco_frag_satd2(&dc, s, r + o0, r + o1, ystride) + (unsigned)abs(dc);
          if (err < best) { best = err; bdx = k_ref_dx[si]; bdy = k_ref_dy[si]; }
        }
        o->ref_mv[bi][0] = (int8_t)(2 * m->block_vec[bi][0] + bdx);
        o->ref_mv[bi][1] = (int8_t)(2 * m->block_vec[bi][1] + bdy);
        o->block_satd[bi] = best;
      }
    }
  }
}

/* ---------------------------------------------------------------------- */
/* Whole-frame motion analysis: oc_mcenc_search (mcenc.c:517-548) for every
   macro block in coding order, with the candidate sets of
   oc_mcenc_find_candidates_a/_b (mcenc.c:90-164) built from the neighbours'
   current vectors, and the refinements where oc_enc_analyze_inter runs them
   (analyze.c:2469-2489).  State layout = ocg_me_mb (oc_mb_enc_info fields). */
static int oco_mv_x(int mv) { return (int)(signed char)(mv & 0xFF); }
static int oco_mv_y(int mv) { return (int)(int16_t)mv >> 8; }
static int oco_mv(int x, int y) { return (int)(int16_t)((x & 0xFF) | (y * 256)); }
static int oco_clamp31(int v) { return v < -31 ? -31 : (v > 31 ? 31 : v); }
static void oco_sort2(int *a, int *b) { if (*a > *b) { int t = *a; *a = *b; *b = t; } }

static void oco_me_search_one(const uint8_t *src, const uint8_t *ref_full, const uint8_t *ref_satd, int ystride,
                              const ocg_me_topo *topo, ocg_me_mb *mb, int mbi, int f, int flags,
                              const uint8_t *gold_refine) {
  const ocg_me_topo *t = &topo[mbi];
  ocg_me_mb *m = &mb[mbi];
  ocg_mb_search_in in;
  ocg_mb_search_out out;
  int mv0 = m->analysis_mv[0][f], mv1 = m->analysis_mv[1][f], mv2 = m->analysis_mv[2][f];
  int accum, ax, ay, nc = 1, i, a[3][2], h1, h2, refine;
  unsigned t2;
  /* mcenc.c:523-531 / 540-541 */
  if (f == 1) {
    int old2 = mv2;
    accum = (flags & OCG_ME_DROPPED) ? mv0 : 0;
    mv2 = mv1;
    mv1 = oco_mv(oco_mv_x(mv0) - oco_mv_x(old2), oco_mv_y(mv0) - oco_mv_y(old2));
  } else {
    accum = mv2;
    mv2 = mv1;
    mv1 = mv0;
    mv1 = oco_mv(oco_mv_x(mv1) - oco_mv_x(mv2), oco_mv_y(mv1) - oco_mv_y(mv2));
    mv2 = oco_mv(oco_mv_x(mv2) - oco_mv_x(accum), oco_mv_y(mv2) - oco_mv_y(accum));
  }
  ax = oco_mv_x(accum);
  ay = oco_mv_y(accum);
  memset(&in, 0, sizeof(in));
  for (i = 0; i < 4; i++) in.frag_off[i] = t->frag_off[i];
  /* set A, mcenc.c:101-127 */
  for (i = 0; i < t->ncn; i++) {
    in.cand[nc][0] = (int8_t)oco_mv_x(mb[t->cn[i]].analysis_mv[0][f]);
    in.cand[nc][1] = (int8_t)oco_mv_y(mb[t->cn[i]].analysis_mv[0][f]);
    nc++;
  }
  in.cand[nc][0] = (int8_t)ax; in.cand[nc][1] = (int8_t)ay; nc++;
  in.cand[nc][0] = (int8_t)oco_clamp31(oco_mv_x(mv1) + ax);
  in.cand[nc][1] = (int8_t)oco_clamp31(oco_mv_y(mv1) + ay);
  nc++;
  in.cand[nc][0] = in.cand[nc][1] = 0; nc++;
  /* median of the first three, mcenc.c:130-138 */
  for (i = 0; i < 3; i++) { a[i][0] = in.cand[1 + i][0]; a[i][1] = in.cand[1 + i][1]; }
  oco_sort2(&a[0][0], &a[1][0]); oco_sort2(&a[0][1], &a[1][1]);
  oco_sort2(&a[1][0], &a[2][0]); oco_sort2(&a[1][1], &a[2][1]);
  oco_sort2(&a[0][0], &a[1][0]); oco_sort2(&a[0][1], &a[1][1]);
  in.cand[0][0] = (int8_t)a[1][0];
  in.cand[0][1] = (int8_t)a[1][1];
  in.setb0 = (uint8_t)nc;
  /* set B, mcenc.c:156-163 */
  in.cand[nc][0] = (int8_t)oco_clamp31(2 * oco_mv_x(mv1) - oco_mv_x(mv2) + ax);
  in.cand[nc][1] = (int8_t)oco_clamp31(2 * oco_mv_y(mv1) - oco_mv_y(mv2) + ay);
  nc++;
  in.ncand = (uint8_t)nc;
  /* mcenc.c:337-341 */
  t2 = m->error[f];
  for (i = 0; i < (t->ncn < 3 ? t->ncn : 3); i++)
    if (mb[t->cn[i]].error[f] > t2) t2 = mb[t->cn[i]].error[f];
  in.t2_base = (uint16_t)t2;
  in.is_prev = (uint8_t)(f == 1);
  oco_mcenc_search_batch(src, ref_full, ref_satd, ystride, &in, &out, 1);
  if (flags & OCG_ME_NOSATD) { /* mcenc.c:238-241: SAD instead of SATD for the final score */
    unsigned s = 0;
    for (i = 0; i < 4; i++)
      s += oco_frag_sad(src + t->frag_off[i], ref_satd + t->frag_off[i] + out.best_vec[0] + out.best_vec[1] * ystride, ystride);
    out.satd = s;
  }
  /* mcenc.c:534, 546-547 */
  if (f == 1) { h2 = accum; h1 = mv1; }
  else {
    h2 = oco_mv(oco_mv_x(mv2) + ax, oco_mv_y(mv2) + ay);
    h1 = oco_mv(oco_mv_x(mv1) + oco_mv_x(h2), oco_mv_y(mv1) + oco_mv_y(h2));
  }
  m->analysis_mv[1][f] = (int16_t)h1;
  m->analysis_mv[2][f] = (int16_t)h2;
  m->error[f] = out.error;
  m->analysis_mv[0][f] = m->unref_mv[f] = (int16_t)oco_mv(out.best_vec[0] * 2, out.best_vec[1] * 2);
  m->satd[f] = m->unref_satd[f] = out.satd;
  if (f == 1 && !(flags & OCG_ME_FAST)) {
    for (i = 0; i < 4; i++) {
      m->block_mv[i] = (int16_t)oco_mv(out.block_vec[i][0] * 2, out.block_vec[i][1] * 2);
      m->block_satd[i] = out.block_satd[i];
    }
  }
  refine = f == 1 ? (flags & OCG_ME_REFINE_PREV) != 0 : (gold_refine != NULL && gold_refine[mbi] != 0);
  if (refine) { /* analyze.c:2476-2489 -> mcenc.c:666-675 */
    ocg_mb_refine_in rin;
    ocg_mb_refine_out rout;
    memset(&rin, 0, sizeof(rin));
    for (i = 0; i < 4; i++) rin.frag_off[i] = t->frag_off[i];
    rin.vec[0] = out.best_vec[0];
    rin.vec[1] = out.best_vec[1];
    rin.satd = out.satd;
    oco_mcenc_refine_batch(src, ref_satd, ystride, &rin, &rout, 1,
                           OCG_REFINE_1MV | ((flags & OCG_ME_NOSATD) ? OCG_REFINE_SAD : 0));
    m->analysis_mv[0][f] = (int16_t)oco_mv(rout.mv[0], rout.mv[1]);
    m->satd[f] = rout.satd;
  }
}

OCO_EXPORT void oco_me_frame(const uint8_t *src, const uint8_t *ref_full_gold, const uint8_t *ref_full_prev,
                             const uint8_t *ref_satd_gold, const uint8_t *ref_satd_prev, int ystride,
                             const ocg_me_topo *topo, ocg_me_mb *mb, int nmbs, int flags, const uint8_t *gold_refine) {
  int mbi, i;
  for (mbi = 0; mbi < nmbs; mbi++) {
    if (!topo[mbi].valid) continue;
    oco_me_search_one(src, ref_full_prev, ref_satd_prev, ystride, topo, mb, mbi, 1, flags, gold_refine);
    oco_me_search_one(src, ref_full_gold, ref_satd_gold, ystride, topo, mb, mbi, 0, flags, gold_refine);
  }
  if ((flags & OCG_ME_REFINE_4MV) && !(flags & OCG_ME_FAST)) {
    for (mbi = 0; mbi < nmbs; mbi++) { /* mcenc.c:763-791 */
      ocg_mb_refine_in rin;
      ocg_mb_refine_out rout;
      if (!topo[mbi].valid) continue;
      memset(&rin, 0, sizeof(rin));
      for (i = 0; i < 4; i++) {
        rin.frag_off[i] = topo[mbi].frag_off[i];
        rin.block_vec[i][0] = (int8_t)(oco_mv_x(mb[mbi].block_mv[i]) / 2);
        rin.block_vec[i][1] = (int8_t)(oco_mv_y(mb[mbi].block_mv[i]) / 2);
        rin.block_satd[i] = mb[mbi].block_satd[i];
      }
      oco_mcenc_refine_batch(src, ref_satd_prev, ystride, &rin, &rout, 1, OCG_REFINE_4MV);
      for (i = 0; i < 4; i++) {
        mb[mbi].ref_mv[i] = (int16_t)oco_mv(rout.ref_mv[i][0], rout.ref_mv[i][1]);
        mb[mbi].ref_block_satd[i] = rout.block_satd[i];
      }
    }
  }
}

/* ==========================================================================
 * Out-of-loop post-processing (decode.c:1609-1957), restated for a whole plane.
 * Planes are given in the decoder's internal orientation: row 0 is the row the
 * reference processes first (the displayed image's bottom row), `stride` bytes
 * from one row to the next.  W and H are multiples of 8.  dc_qis / qis are per
 * block of the plane (nh x nv, row-major in the same orientation).
 * ========================================================================== */
static int oco_pp_edge_sums(const int r[10], int *sum0, int *sum1) {
  int k, a = 0, b = 0;
  for (k = 0; k < 4; k++) { a += abs(r[k + 1] - r[k]); b += abs(r[k + 5] - r[k + 6]); }
  *sum0 = a;
  *sum1 = b;
  return 0;
}

static void oco_pp_edge_filter(const int r[10], int out[8]) {
  int k;
  out[0] = (r[0] * 3 + r[1] * 2 + r[2] + r[3] + r[4] + 4) >> 3;
  out[1] = (r[0] * 2 + r[1] + r[2] * 2 + r[3] + r[4] + r[5] + 4) >> 3;
  for (k = 0; k < 4; k++) out[2 + k] = (r[k] + r[k + 1] + r[k + 2] + r[k + 3] * 2 + r[k + 4] + r[k + 5] + r[k + 6] + 4) >> 3;
  out[6] = (r[4] + r[5] + r[6] + r[7] * 2 + r[8] + r[9] * 2 + 4) >> 3;
  out[7] = (r[5] + r[6] + r[7] + r[8] * 2 + r[9] * 3 + 4) >> 3;
}

/* oc_dec_deblock_frag_rows over the whole plane: dst <- filtered src; variances[nh*nv] accumulated from 0 */
OCO_EXPORT void oco_pp_deblock_plane(uint8_t *dst, int dstride, const uint8_t *src, int sstride, int W, int H,
                                     const uint8_t *dc_qis, const int *dc_scale, int32_t *variances) {
  const int nh = W >> 3, nv = H >> 3;
  int x, y, e, bx, by, k;
  memset(variances, 0, (size_t)nh * nv * sizeof(*variances));
  /* rows no horizontal edge reaches */
  for (y = 0; y < 4; y++) memcpy(dst + (size_t)y * dstride, src + (size_t)y * sstride, (size_t)W);
  for (y = H - 4; y < H; y++) memcpy(dst + (size_t)y * dstride, src + (size_t)y * sstride, (size_t)W);
  /* horizontal edges between block rows e-1 and e: 10 source rows in, 8 rows out (decode.c:1610-1660) */
  for (e = 1; e < nv; e++) {
    for (x = 0; x < W; x++) {
      const int qstep = dc_scale[dc_qis[(e - 1) * nh + (x >> 3)]], flimit = (qstep * 3) >> 2;
      int r[10], o[8], s0, s1;
      for (k = 0; k < 10; k++) r[k] = src[(size_t)(8 * e - 5 + k) * sstride + x];
      oco_pp_edge_sums(r, &s0, &s1);
      variances[(e - 1) * nh + (x >> 3)] += s0 < 255 ? s0 : 255;
      variances[e * nh + (x >> 3)] += s1 < 255 ? s1 : 255;
      if (s0 < flimit && s1 < flimit && r[5] - r[4] < qstep && r[4] - r[5] < qstep) oco_pp_edge_filter(r, o);
      else for (k = 0; k < 8; k++) o[k] = r[k + 1];
      for (k = 0; k < 8; k++) dst[(size_t)(8 * e - 4 + k) * dstride + x] = (uint8_t)o[k];
    }
  }
  /* vertical edges, in place, left to right along every row (decode.c:1663-1699, 1750-1790) */
  for (by = 0; by < nv; by++) {
    for (bx = 1; bx < nh; bx++) {
      const int qstep = dc_scale[dc_qis[by * nh + bx]], flimit = (qstep * 3) >> 2;
      for (y = 8 * by; y < 8 * by + 8; y++) {
        uint8_t *p = dst + (size_t)y * dstride + 8 * bx - 5;
        int r[10], o[8], s0, s1;
        for (k = 0; k < 10; k++) r[k] = p[k];
        oco_pp_edge_sums(r, &s0, &s1);
        variances[by * nh + bx - 1] += s0 < 255 ? s0 : 255;
        variances[by * nh + bx] += s1 < 255 ? s1 : 255;
        if (s0 < flimit && s1 < flimit && r[5] - r[4] < qstep && r[4] - r[5] < qstep) {
          oco_pp_edge_filter(r, o);
          for (k = 0; k < 8; k++) p[1 + k] = (uint8_t)o[k];
        }
      }
    }
  }
}

/* oc_dering_block (decode.c:1788-1886) on the block at (x0,y0) of a W x H plane, with neighbour pixels
   clamped into the plane */
static void oco_pp_dering_block(uint8_t *img, int stride, int W, int H, int x0, int y0, int dc_scale, int sharp_mod,
                                int strong) {
  const int mod_hi = 3 * dc_scale < (strong ? 32 : 24) ? 3 * dc_scale : (strong ? 32 : 24);
  const int shift = strong ? 0 : 1;
  int vmod[9][8], hmod[9][8];
  int bx, by;
#define OCO_PIX(xx, yy) ((int)img[(size_t)((yy) < 0 ? 0 : ((yy) >= H ? H - 1 : (yy))) * stride + ((xx) < 0 ? 0 : ((xx) >= W ? W - 1 : (xx)))])
#define OCO_MOD(d) (32 + dc_scale - (abs(d) << shift))
  for (by = 0; by < 9; by++)
    for (bx = 0; bx < 8; bx++) {
      const int m = OCO_MOD(OCO_PIX(x0 + bx, y0 + by) - OCO_PIX(x0 + bx, y0 + by - 1));
      vmod[by][bx] = m < -64 ? sharp_mod : (m < 0 ? 0 : (m > mod_hi ? mod_hi : m));
    }
  for (bx = 0; bx < 9; bx++)
    for (by = 0; by < 8; by++) {
      const int m = OCO_MOD(OCO_PIX(x0 + bx, y0 + by) - OCO_PIX(x0 + bx - 1, y0 + by));
      hmod[bx][by] = m < -64 ? sharp_mod : (m < 0 ? 0 : (m > mod_hi ? mod_hi : m));
    }
  /* at the frame border the reference compares a row/column with itself (the pointer does not advance),
     which the clamped coordinates above reproduce except for row/column 8 against 7: both clamp to the
     same pixel there too */
  for (by = 0; by < 8; by++)
    for (bx = 0; bx < 8; bx++) {
      int a = 128, b = 64, w, v;
      w = hmod[bx][by]; a -= w; b += w * OCO_PIX(x0 + bx - 1, y0 + by);
      w = vmod[by][bx]; a -= w; b += w * OCO_PIX(x0 + bx, y0 + by - 1);
      w = vmod[by + 1][bx]; a -= w; b += w * OCO_PIX(x0 + bx, y0 + by + 1);
      w = hmod[bx + 1][by]; a -= w; b += w * OCO_PIX(x0 + bx + 1, y0 + by);
      v = (a * OCO_PIX(x0 + bx, y0 + by) + b) >> 7;
      img[(size_t)(y0 + by) * stride + x0 + bx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
#undef OCO_PIX
#undef OCO_MOD
}

/* oc_dec_dering_frag_rows over the whole plane (decode.c:1892-1957) */
OCO_EXPORT void oco_pp_dering_plane(uint8_t *img, int stride, int W, int H, int pli, int strong_level, const uint8_t *qis,
                                    const int *dc_scale, const int *sharp_mod, const int32_t *variances) {
  const int nh = W >> 3, nv = H >> 3;
  const int t1 = 384, t2 = 4 * 384, t3 = 5 * 384, t4 = 10 * 384;
  const int sthresh = pli ? t4 : t3;
  int bx, by;
  for (by = 0; by < nv; by++)
    for (bx = 0; bx < nh; bx++) {
      const int i = by * nh + bx, var = variances[i], qi = qis[i];
      if (strong_level && var > sthresh) {
        int passes = 1, k;
        if (pli || (bx > 0 && variances[i - 1] > t4) || (bx + 1 < nh && variances[i + 1] > t4) ||
            (by > 0 && variances[i - nh] > t4) || (by + 1 < nv && variances[i + nh] > t4))
          passes = 3;
        for (k = 0; k < passes; k++) oco_pp_dering_block(img, stride, W, H, 8 * bx, 8 * by, dc_scale[qi], sharp_mod[qi], 1);
      } else if (var > t2) oco_pp_dering_block(img, stride, W, H, 8 * bx, 8 * by, dc_scale[qi], sharp_mod[qi], 1);
      else if (var > t1) oco_pp_dering_block(img, stride, W, H, 8 * bx, 8 * by, dc_scale[qi], sharp_mod[qi], 0);
    }
}
