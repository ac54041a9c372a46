/* Controls of the vtable back-end that are not part of the th_* API. */
#ifndef OCG_BACKEND_H
#define OCG_BACKEND_H
#include "../../include/theora_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* OCG_BACKEND_GPU: record one frame's lists, run them on the device, copy the
   frame back (the product path).
   OCG_BACKEND_RECORD: instrumentation for tests -- lists are recorded and handed
   to the capture callback only; no device is touched and NO pixels are
   produced (th_decode_ycbcr_out is meaningless in this mode). */
#define OCG_BACKEND_GPU    0
#define OCG_BACKEND_RECORD 1
OCG_API void ocg_backend_set_mode(int mode);      /* applies to decoders allocated afterwards */
OCG_API void ocg_backend_set_device(int device);  /* CUDA device for decoders allocated afterwards (process-wide: one process per GPU) */

/* Called at every frame flush with the frame description (list pointers NULL)
   and the staged lists, before they are submitted. */
typedef void (*ocg_capture_fn)(void *user, const ocg_dec_frame *frame, const ocg_staging *lists);
OCG_API void ocg_backend_set_capture(ocg_capture_fn fn, void *user);

/* Process-wide statistics since the last reset. */
typedef struct ocg_backend_stats {
  long   frames;
  long   coded_frags;
  long   uncoded_frags;
  long   coeff_rows;
  long   h2d_bytes;
  long   d2h_bytes;
  double flush_seconds;   /* host wall time inside flush (submit + sync) */
} ocg_backend_stats;
OCG_API void ocg_backend_get_stats(ocg_backend_stats *out, int reset);

#ifdef __cplusplus
}
#endif
#endif
