/* TEST INFRASTRUCTURE ONLY (oracle/): lets Python call the reference's
 * file-static analysis helpers by compiling lib/analyze.c as part of this
 * translation unit (read from where it lies, never copied):
 *   oc_mb_activity     lib/analyze.c:1152
 *   oc_mb_masking      lib/analyze.c:1279
 * Built into oracle/_ref/libth_c_analyze.so (own copy of analyze.c's
 * functions, bound locally with -Bsymbolic; everything else resolves from
 * libth_c.so). */
#include "analyze.c"

#define REFH_API __attribute__((visibility("default")))

/* io = flipped luma base of the input frame (ref_frame_data[OC_FRAME_IO]). */
REFH_API unsigned refh_mb_activity(const unsigned char *io, int ystride, const long frag_off[4], unsigned act[4]) {
  oc_enc_ctx *enc = (oc_enc_ctx *)calloc(1, sizeof(*enc));
  oc_sb_map sb_map;
  ptrdiff_t offs[4];
  unsigned luma;
  int i;
  memset(sb_map, 0xFF, sizeof(sb_map));
  for (i = 0; i < 4; i++) { sb_map[0][i] = i; offs[i] = frag_off[i]; }
  enc->state.sb_maps = &sb_map;
  enc->state.frag_buf_offs = offs;
  enc->state.ref_frame_data[OC_FRAME_IO] = (unsigned char *)io;
  enc->state.ref_ystride[0] = ystride;
  luma = oc_mb_activity(enc, 0, act);
  free(enc);
  return luma;
}

REFH_API unsigned refh_mb_masking(unsigned rd_scale[5], unsigned rd_iscale[5], const unsigned short chroma_rd_scale[2],
                                  const unsigned activity[4], unsigned activity_avg, unsigned luma, unsigned luma_avg) {
  return oc_mb_masking(rd_scale, rd_iscale, chroma_rd_scale, activity, activity_avg, luma, luma_avg);
}
