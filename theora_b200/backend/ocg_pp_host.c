/* ocg_pp_host.c -- out-of-loop post-processing (TH_DECCTL_SET_PPLEVEL > 0) for
 * the record-and-flush decoder back-end.
 *
 * The de-blocking / de-ringing filters (decode.c:1609-1957) are non-normative,
 * file-static, and run inside th_decode_packetin's MCU loop on the host's copy
 * of the frame (decode.c:2899-2922) -- which, with this back-end, only becomes
 * valid at the end-of-frame flush.  They are not part of the block pipeline
 * (SURVEY 8(f)3 leaves them on the host), so the back-end re-runs them over the
 * WHOLE frame right after the flush has copied the reconstructed frame back:
 * same functions, same state (dc_qis, variances, pp_frame_buf are prepared by
 * the unmodified oc_dec_postprocess_init, decode.c:1204), row range
 * [0, nvfrags) instead of the striped ranges with their start/end delays.
 * To reach the static functions, lib/decode.c is compiled into this
 * translation unit a second time (read from where it lies, never copied) with
 * its external definitions renamed out of the way. */
#define th_decode_alloc                 ocgpp_unused_decode_alloc
#define th_decode_free                  ocgpp_unused_decode_free
#define th_decode_ctl                   ocgpp_unused_decode_ctl
#define th_decode_packetin              ocgpp_unused_decode_packetin
#define th_decode_ycbcr_out             ocgpp_unused_decode_ycbcr_out
#define oc_dec_accel_init_c             ocgpp_unused_dec_accel_init_c
#define oc_dec_dc_unpredict_mcu_plane_c ocgpp_unused_dc_unpredict_mcu_plane_c
#include "decode.c"

/* decode.c:2899-2914 for every plane, all rows at once.  refi = the buffer the
   frame was reconstructed into (SELF at decode time). */
void ocg_pp_host_whole_frame(oc_dec_ctx *_dec, int _refi) {
  int pli;
  for (pli = 0; pli < 3; pli++) {
    int pp_offset = 3 * (pli != 0);
    int nvfrags = _dec->state.fplanes[pli].nvfrags;
    if (_dec->pipe.pp_level >= OC_PP_LEVEL_DEBLOCKY + pp_offset) {
      oc_dec_deblock_frag_rows(_dec, _dec->pp_frame_buf, _dec->state.ref_frame_bufs[_refi], pli, 0, nvfrags);
      if (_dec->pipe.pp_level >= OC_PP_LEVEL_DERINGY + pp_offset) {
        oc_dec_dering_frag_rows(_dec, _dec->pp_frame_buf, pli, 0, nvfrags);
      }
    }
  }
}
