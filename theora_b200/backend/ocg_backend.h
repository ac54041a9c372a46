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

/* Where the DC prediction is undone (oc_dec_dc_unpredict_mcu_plane, decode.c:1392):
   OCG_DC_HOST: the reference's C routine runs in the hook and the records carry final DC values (also
   what resident packs need).
   OCG_DC_DEVICE: the hook only counts the coded fragments of the MCU (its other duty,
   decode.c:1496-1499).  On the device path the recurrence is started at the first hook of the frame
   (ocg_dec_dc_begin: its inputs are all in frags[] by then) and runs WHILE the host expands the frame's
   coefficients; the flush patches the final values into the records (dc_residual=2).  In record mode the
   records simply carry the residuals (dc_residual=1).  Geometries the kernel does not cover
   (ocg_dc_unpredict_supported) stay on the host.  Bit-exact either way.
   The kernel alone is slow on inter frames (the "last value of the same reference type" predictor,
   decode.c:1452, links rows end-to-start: chains of thousands of fragments at ~0.25 us each vs ~7 ns on a
   CPU core; +0.5..1.1 ms per 1080p inter frame).  Started ahead of the lists it overlaps the host's
   ~0.55 ms of coefficient expansion, which is not enough to hide it: measured through
   th_decode_packetin 8.5 k vs 11.5 k frames/s for OCG_DC_HOST (6.8 k when the kernel sits in the flush),
   so the host routine stays the default. */
#define OCG_DC_DEVICE 0        /* inside the flush graph (dc_residual=1) */
#define OCG_DC_HOST   1
#define OCG_DC_DEVICE_AHEAD 2  /* started at the first hook of the frame (ocg_dec_dc_begin, dc_residual=2) */
OCG_API void ocg_backend_set_dc_mode(int mode);   /* applies to decoders allocated afterwards */

/* Who expands a coded fragment's tokens into coefficients (decode.c:1531-1586):
   OCG_EXPAND_DEVICE (default): the device (ocg_dec_flush_tokens): the decoder's fragment words, vectors and
   token lists are read in place and nothing is recorded per fragment on the host.
   OCG_EXPAND_REFERENCE: the reference's loop, one state_frag_recon hook call per fragment, which records a
   16-byte record and the non-zero coefficient rows (what the capture hook and record mode hand out; also
   used whenever a capture hook is installed). */
#define OCG_EXPAND_DEVICE    0
#define OCG_EXPAND_REFERENCE 1
OCG_API void ocg_backend_set_expand_mode(int mode);   /* applies to decoders allocated afterwards */

/* What the flush copies back into the decoder's host reference buffer (the memory th_decode_ycbcr_out
   hands out): OCG_OUT_PICTURE (default) = the coded-frame area of the three planes, which is all the API
   exposes; OCG_OUT_PADDED = the whole padded buffer including the aprons (diagnostics). */
OCG_API void ocg_backend_set_output_mode(int mode);   /* applies to decoders allocated afterwards */

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
  double flush_seconds;   /* host wall time inside the flush (queueing copies and kernels) */
  double wait_seconds;    /* host wall time waiting for a flushed frame (th_decode_ycbcr_out, stripe callback) */
} ocg_backend_stats;
OCG_API void ocg_backend_get_stats(ocg_backend_stats *out, int reset);

/* ---- encoder ---------------------------------------------------------------
   OCG_ENC_AUTO: an intra-only encoder (th_info.keyframe_granule_shift == 0, so
   every frame is a key frame) runs its block pipeline on the device: one batched
   pre-pass per frame (intra SATD, sub_128 + fDCT, quantiser) served to the
   per-block hooks as look-ups, and the reconstruction (iDCT + recon + loop
   filter + borders) recorded and flushed like a decoded frame.  th_encode_alloc
   returns NULL when no device is usable.  Encoders that can emit inter frames
   keep the reference's C kernels (their hooks need a restructured caller).
   OCG_ENC_HOST: every encoder keeps the reference's C kernels -- tooling mode
   for producing test streams on machines without a GPU. */
#define OCG_ENC_AUTO 0
#define OCG_ENC_HOST 1
OCG_API void ocg_backend_set_enc_mode(int mode);  /* applies to encoders allocated afterwards */

typedef struct ocg_enc_backend_stats {
  long   frames;            /* frames reconstructed on the device            */
  long   prepass_frames;    /* analysis passes (>= frames: dry runs, recodes) */
  long   coeff_rows;
  long   h2d_bytes;
  long   d2h_bytes;
  double prepass_seconds;   /* host wall time inside the pre-pass call        */
  double flush_seconds;     /* host wall time inside the reconstruction flush */
  long   me_frames;         /* analysis passes whose motion analysis ran on the device          */
  long   me_gold_refines;   /* oc_mcenc_refine1mv(OC_FRAME_GOLD) decisions of the host loop      */
  long   me_repairs;        /* GOLD searches redone because a neighbour's refinement changed their candidates */
  /* inter frames that were packed: how the analysis loop's block-metric calls were served */
  long   satd_lookups;      /* frag_satd / frag_satd2 answered from the device's candidate tables          */
  long   satd_host;         /* ... computed by the reference's C kernel (no candidate with that predictor) */
  long   ssd_lookups;       /* frag_ssd / frag_border_ssd of oc_skip_cost answered from the device table   */
  long   ssd_host;          /* ... of the block just reconstructed (analyze.c:829-835): host               */
  long   intra_satd_lookups;
  long   fdct_quant_lookups; /* inter blocks whose frag_sub + fdct8x8 + quantize came from the device tables */
  long   fdct_quant_host;    /* ... computed by the reference's C kernels (another predictor, pool exhausted) */
  double me_queue_seconds;  /* host time queueing the inter-frame pre-pass (copies + kernels)       */
  double me_sync_seconds;   /* ... and waiting for its results                                      */
  double prev_wait_seconds; /* host time at the start of a pass waiting for the previous frame's flush */
  double me_prep_seconds;   /* ... preparing the motion analysis state (copies of oc_mb_enc_info)       */
} ocg_enc_backend_stats;
OCG_API void ocg_backend_get_enc_stats(ocg_enc_backend_stats *out, int reset);
/* Test instrumentation: a snapshot at the start of every analysis pass of an encoder that runs on the
   host (enquant_table_fixup, analyze.c:564: input frame in place, references rotated, the pass's own
   motion search not yet started): the frame buffers by role and the per-macro-block analysis state
   (oc_mb_enc_info, encint.h:346-382) converted to the layout of ocg_me_mb, i.e. the state the previous
   pass left behind.  Lets a test replay the reference encoder's own motion analysis on the device. */
typedef struct ocg_enc_spy_frame {
  int32_t frame_type;              /* OC_INTRA_FRAME 0 / OC_INTER_FRAME 1 of the pass that is starting */
  int32_t prevframe_dropped;
  int32_t sp_level;
  int32_t keyframe_frequency_force;
  int32_t nmbs;
  int32_t reserved;
  int64_t curframe_num;
  int64_t ref_frame_sz;
  const unsigned char *frames[5];  /* whole buffers: IO, PREV_ORIG, GOLD_ORIG, PREV, GOLD (NULL: none yet) */
  const ocg_me_mb *state;          /* nmbs entries; block_satd and ref_block_satd both hold embs.block_satd */
  const unsigned char *refined;    /* oc_mb_enc_info.refined per macro block */
} ocg_enc_spy_frame;
typedef void (*ocg_enc_spy_fn)(void *user, const ocg_enc_spy_frame *frame);
OCG_API void ocg_backend_set_enc_spy(ocg_enc_spy_fn fn, void *user); /* applies to encoders allocated afterwards */

/* Test accessor: copies the encoder's current reconstruction (three top-down
   planes, frame_width x frame_height, packed); returns the byte count. */
struct th_enc_ctx;
OCG_API long ocg_backend_enc_copy_recon(struct th_enc_ctx *enc, unsigned char *dst);

#ifdef __cplusplus
}
#endif
#endif
