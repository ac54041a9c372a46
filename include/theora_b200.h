/* theora_b200.h -- C ABI of the B200 (sm_100a) back-end for libtheora's
 * per-fragment 8x8 block pipeline.
 *
 * This is the drop-in boundary.  Everything above it (bit reader, Huffman,
 * token unpack, DC un-prediction, mode decision, rate control) stays host C in
 * the reference; everything below it is hand-written CUDA.  The reference-side
 * binding is theora_b200/backend/ocg_hooks.h + ocg_backend.c (see
 * INTEGRATION.md): it fills oc_base_opt_vtable (reference lib/state.h:352-370)
 * and oc_enc_opt_vtable (lib/encint.h:292-326) with recorders and flushes one
 * frame's lists through the entry points declared here.
 *
 * Conventions shared with the reference:
 *   - frame buffers use the reference's exact padded layout (state.c:545-671):
 *     nrefs consecutive buffers of ref_frame_sz bytes, each luma | Cb | Cr with
 *     16/8-pixel aprons; pixel addressing is bottom-up (negative row stride)
 *     relative to `base_off`, the byte offset of the bottom-left luma pixel,
 *     so state->frag_buf_offs[] and state->ref_ystride[] apply verbatim.
 *   - coefficients are int16 in natural (row-major) order, AC already
 *     dequantised (decode.c:1573-1574), DC raw (state.c:972,978 applies
 *     dc_quant on the device).
 *   - plain pointers and sizes only; every call returns 0 or a negative
 *     OCG_E* code and never throws; a missing/failed CUDA device is an error,
 *     there is no CPU fallback.
 */
#ifndef THEORA_B200_H
#define THEORA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
# define OCG_API __attribute__((visibility("default")))
#else
# define OCG_API
#endif

#define OCG_OK        0
#define OCG_EFAULT   (-1)   /* NULL argument (TH_EFAULT, codec.h:77) */
#define OCG_EINVAL   (-10)  /* bad argument  (TH_EINVAL, codec.h:79) */
#define OCG_EIMPL    (-23)  /* unsupported   (TH_EIMPL,  codec.h:87) */
#define OCG_ECUDA    (-100) /* CUDA runtime failure; ocg_last_error() has the text */
#define OCG_ENOMEM   (-101)

/* Reference-frame roles, same numbering as OC_FRAME_* (state.h:270-282). */
#define OCG_FRAME_GOLD 0
#define OCG_FRAME_PREV 1
#define OCG_FRAME_SELF 2

/* Sparsity classes of a coded fragment, selected from last_zzi exactly as
   oc_state_frag_recon_c (state.c:967) and oc_idct8x8_c (idct.c:327-329) do. */
#define OCG_CLS_DC    0   /* last_zzi<2 : (dc*dc_quant+15)>>5, no iDCT        */
#define OCG_CLS_3     1   /* last_zzi<=3 : coefficients in the top-left 2x2    */
#define OCG_CLS_10    2   /* last_zzi<=10: coefficients in the top-left 4x4    */
#define OCG_CLS_FULL  3   /* everything else: full 8x8                         */
#define OCG_NCLS      4

typedef struct ocg_plane_geom {
  int32_t  nhfrags;    /* fragments per row        (state.h oc_fragment_plane) */
  int32_t  nvfrags;    /* fragment rows                                          */
  int32_t  froffset;   /* index of the plane's first fragment                    */
  int32_t  nfrags;
  int32_t  ystride;    /* NEGATIVE byte stride, == state->ref_ystride[pli]       */
  int32_t  width;      /* plane width/height in pixels (coded frame)             */
  int32_t  height;
  int32_t  hpad;       /* apron width / height in pixels                         */
  int32_t  vpad;
  int64_t  plane_off;  /* bottom-left pixel of the plane relative to base_off
                          (== frag_buf_offs[froffset])                           */
} ocg_plane_geom;

typedef struct ocg_geometry {
  int32_t        frame_width;   /* coded size, multiples of 16 */
  int32_t        frame_height;
  int32_t        pixel_fmt;     /* TH_PF_420=0, TH_PF_422=2, TH_PF_444=3 (codec.h:94-107) */
  int32_t        nrefs;         /* 3 for a decoder, 6 for an encoder (state.c:545) */
  int32_t        nfrags;
  int32_t        reserved;
  int64_t        ref_frame_sz;  /* bytes per buffer   (state.c:582)                */
  int64_t        base_off;      /* ref_frame_bufs[0][0].data - ref_frame_handle
                                   after the flip (state.c:622-629)               */
  ocg_plane_geom planes[3];
} ocg_geometry;

/* refi value of a fragment that is NOT coded in this frame: its pixels are
   copied from the previous reference (oc_frag_copy_list, fragment.c:37). */
#define OCG_FRAG_UNCODED 3

/* One fragment (16 bytes).  A frame carries exactly nfrags of these, indexed
   by the reference's fragment index (raster order inside each plane, planes
   back to back: state.h oc_fragment_plane.froffset).  For coded fragments it
   replaces the argument list of oc_state_frag_recon (state.h:361-362,
   state.c:959); uncoded ones only need buf_off and the plane. */
typedef struct ocg_frag_rec {
  int32_t  buf_off;    /* state->frag_buf_offs[fragi]                             */
  int16_t  mv;         /* state->frag_mvs[fragi]: dx=(int8)mv, dy=mv>>8           */
  int16_t  dc;         /* frags[fragi].dc after DC un-prediction (decode.c:1392)  */
  uint32_t coeff_row;  /* index of this fragment's first stored row in coeff_rows */
  uint8_t  rowmask;    /* bit j: natural-order row j is stored (16 B per row)     */
  uint8_t  last_zzi;   /* as handed to oc_state_frag_recon, 0..64                 */
  uint8_t  refi;       /* OCG_FRAME_* (SELF = intra) or OCG_FRAG_UNCODED          */
  uint8_t  pli_qti;    /* pli | qti<<2                                            */
} ocg_frag_rec;

/* One frame of decoder-side block work.  Replaces the per-MCU sequence
   oc_dec_frags_recon_mcu_plane -> oc_state_frag_recon / oc_frag_copy_list ->
   oc_state_loop_filter_frag_rows -> oc_state_borders_fill_* of
   decode.c:2858-2945.  Pointers are host pointers for ocg_dec_submit and
   device pointers inside a resident pack. */
typedef struct ocg_dec_frame {
  int32_t             ref_idx[3];      /* buffer playing GOLD, PREV, SELF          */
  int32_t             lf_limit;        /* loop_filter_limits[qis[0]]; 0 = no filter */
  uint16_t            dc_quant[3][2];  /* dequant[pli][0][qti][0] (decode.c:1534)  */
  int32_t             ncoded;          /* coded fragments in recs (informational)  */
  int32_t             intra_frame;     /* 1: key frame, no record references PREV/GOLD */
  int32_t             ncoeff_rows;
  int32_t             dc_residual;     /* 0: recs[].dc are final (the host undid the DC prediction).
                                          1: recs[].dc hold the DC-prediction RESIDUALS as decoded
                                          (decode.c:1277-1316); the device undoes the prediction
                                          (oc_dec_dc_unpredict_mcu_plane, decode.c:1392-1500) before
                                          reconstructing.  2: the final values were computed ahead of
                                          the lists by ocg_dec_dc_begin and replace recs[].dc       */
  const ocg_frag_rec *recs;            /* nfrags records, fragment-index order     */
  const int16_t      *coeff_rows;      /* ncoeff_rows x 8 int16                    */
} ocg_dec_frame;

typedef struct ocg_ctx  ocg_ctx;    /* per th_dec_ctx / th_enc_ctx device state   */
typedef struct ocg_pack ocg_pack;   /* device-resident copy of a run of frames    */

/* ---- library ----------------------------------------------------------- */
OCG_API const char *ocg_version(void);
OCG_API const char *ocg_last_error(void);
OCG_API int         ocg_device_count(void);

/* ---- geometry (restates oc_state_frarray_init/ref_bufs_init, state.c:424-671) */
OCG_API int  ocg_geometry_init(ocg_geometry *g, int frame_width, int frame_height,
                               int pixel_fmt, int nrefs);
OCG_API void ocg_geometry_frag_buf_offs(const ocg_geometry *g, int32_t *offs /* nfrags */);

/* ---- context ------------------------------------------------------------ */
OCG_API int  ocg_ctx_create(ocg_ctx **out, const ocg_geometry *g, int device);
OCG_API void ocg_ctx_destroy(ocg_ctx *ctx);
OCG_API const ocg_geometry *ocg_ctx_geometry(const ocg_ctx *ctx);
OCG_API int  ocg_ctx_sync(ocg_ctx *ctx);
/* 1 if frames of this geometry may be submitted with dc_residual=1 (the DC
   wave-front kernel runs one thread per fragment row of a plane, at most 1024
   rows, reference types of a plane in shared memory), else 0. */
OCG_API int  ocg_dc_unpredict_supported(const ocg_geometry *g);
/* Starts the DC un-prediction of the frame that is being assembled, ahead of its lists: everything the
   recurrence needs (coded flag, reference type, DC residual of every fragment) is known as soon as the
   packet's tokens are unpacked, i.e. before the host expands a single coefficient, so the wave-front
   kernel can run on the device WHILE the host builds the lists.  frag_words: the decoder's oc_fragment
   array (state.h:297-322) viewed as 32-bit words (bit 0 coded, bits 6-7 refi, bits 16-31 dc residual);
   copied before the call returns.  Submit the frame with dc_residual = 2. */
OCG_API int  ocg_dec_dc_begin(ocg_ctx *ctx, const uint32_t *frag_words);
OCG_API void *ocg_ctx_stream(ocg_ctx *ctx);                 /* cudaStream_t */
OCG_API void *ocg_ctx_frame_devptr(ocg_ctx *ctx, int buf);  /* device address of buffer `buf` */
/* Whole padded buffer, host <-> device (ref_frame_sz bytes). */
OCG_API int  ocg_ctx_upload_frame(ocg_ctx *ctx, int buf, const uint8_t *host_buf);
OCG_API int  ocg_ctx_download_frame(ocg_ctx *ctx, int buf, uint8_t *host_buf);
/* Only the coded-frame area of the three planes (what th_decode_ycbcr_out exposes, decode.c:2988-2992),
   device -> the same positions of a host buffer laid out like the reference's; asynchronous on the
   context's stream.  ocg_picture_bytes: the bytes that moves. */
OCG_API int  ocg_ctx_download_picture(ocg_ctx *ctx, int buf, uint8_t *host_buf);
OCG_API long ocg_picture_bytes(const ocg_geometry *g);
OCG_API int  ocg_ctx_fill_frame(ocg_ctx *ctx, int buf, int value);  /* oc_dec_init_dummy_frame, decode.c:2053 */
/* Page-locks caller-owned host memory (the reference's ref_frame_handle) so the
   per-frame D2H runs at full PCIe rate and asynchronously. */
OCG_API int  ocg_host_register(void *p, size_t bytes);
OCG_API int  ocg_host_unregister(void *p);

/* ---- decode: one frame, host lists (the call the vtable back-end makes) -- */
/* Pinned staging owned by the ctx; the recorder writes straight into it (no
   extra host copy): nfrags records and room for nfrags*8 coefficient rows.
   buf_off and the plane bits of pli_qti are filled in once at context creation
   and must be left alone; for every frame the caller sets `refi` of EVERY
   fragment (OCG_FRAG_UNCODED or a reference) and the remaining fields of the
   coded ones.  Valid until the next ocg_dec_submit on this ctx. */
typedef struct ocg_staging {
  ocg_frag_rec *recs;
  int16_t      *coeff_rows;
} ocg_staging;
OCG_API int  ocg_dec_staging(ocg_ctx *ctx, ocg_staging *out);
/* H2D of the lists + recon/copy + loop filter + border fill on the ctx stream.
   Asynchronous.  With f->recs==NULL the lists are taken from the staging
   regions handed out by the preceding ocg_dec_staging call (counts from `f`);
   otherwise f's host arrays are copied into staging first.  If host_out!=NULL
   the finished SELF buffer is also copied back (ref_frame_sz bytes) on the same
   stream; call ocg_ctx_sync before reading it. */
OCG_API int  ocg_dec_submit(ocg_ctx *ctx, const ocg_dec_frame *f, uint8_t *host_out);

/* The same frame flush as ONE driver call: lists H2D, [DC un-prediction], reconstruction, loop filter,
   borders, copy-back and a completion flag in host memory are a CUDA graph that is instantiated once per
   staging slot and replayed (see ocg_ctx_set_flush_graph); ocg_dec_wait polls the flag without entering
   the driver.  host_out must be page-locked (ocg_host_register) and laid out like the reference's buffer;
   out_mode selects what is copied back.  dc_residual 0 or 1. */
#define OCG_OUT_PICTURE 0   /* the coded-frame area of the three planes (what th_decode_ycbcr_out exposes) */
#define OCG_OUT_PADDED  1   /* the whole padded buffer, aprons included */
#define OCG_OUT_NONE    2   /* nothing: the frame stays on the device */
OCG_API int  ocg_dec_flush(ocg_ctx *ctx, const ocg_dec_frame *f, uint8_t *host_out, int out_mode);
OCG_API int  ocg_dec_wait(ocg_ctx *ctx);   /* until the last ocg_dec_flush has completed */
/* From which flush on the kernel sequence is replayed as a graph (default 16; earlier flushes are launched
   kernel by kernel: instantiating a graph only pays off on a stream that lives long enough); < 0: never. */
OCG_API void ocg_ctx_set_flush_graph(ocg_ctx *ctx, int after);
/* Host time spent inside ocg_dec_flush since the last reset, process-wide: before the graph launch
   (seconds), inside cudaGraphLaunch (seconds), number of flushes. */
OCG_API void ocg_flush_profile(double *prepare_s, double *launch_s, long *n, int reset);
OCG_API void ocg_flush_profile_builds(double *build_s, long *nbuilds); /* graph instantiations so far (never reset) */

/* ---- decode: one frame as the reference's decoder holds it after entropy decoding ------------------------
    The token -> coefficient walk of oc_dec_frags_recon_mcu_plane (decode.c:1511-1586) runs on
   the device: the caller hands over the decoder's own arrays -- no per-fragment host work at all.
   ocg_dec_expand_setup, once per context:
     coded_order     [nfrags] the fragments of each plane in coded (super-block Hilbert) order, planes back to
                     back: state.sb_maps walked plane by plane (state.c:200-298)
     dequant_tables  [64 qi][3 pli][2 qti][64] state.dequant_tables, gathered (AC entries are used)
     frag_words      state.frags viewed as 32-bit words (state.h:297-322: bit 0 coded, bits 2-5 qii, bits 6-7
                     refi, bits 8-10 mb_mode, bits 16-31 dc); frag_mvs: state.frag_mvs; dct_tokens:
                     dec->dct_tokens (decode.c:1000-1190), token_capacity its allocated size.  These three stay
                     where they are for the life of the context, must be page-locked (ocg_host_register) and
                     are read in place by the device while a flush is in flight: do not modify them between
                     ocg_dec_flush_tokens and ocg_dec_wait. */
typedef struct ocg_dec_tokens {
  int32_t  ref_idx[3];
  int32_t  lf_limit;
  int32_t  intra_frame;
  int32_t  dc_residual;       /* 0: the words' DC fields are final; 1: residuals, the device undoes the prediction */
  uint16_t dc_quant[3][2];    /* dequant[pli][0][qti][0] (decode.c:1534) */
  int32_t  nqis;
  int32_t  qis[3];            /* state.qis */
  int32_t  ntoken_bytes;      /* dec->dct_tokens_count */
  int32_t  ti0[3][64];        /* dec->ti0 */
  int32_t  eob_runs[3][64];   /* dec->eob_runs, saturated to int32 */
} ocg_dec_tokens;
OCG_API int  ocg_dec_expand_setup(ocg_ctx *ctx, const int32_t *coded_order, const uint16_t *dequant_tables,
                                  const uint32_t *frag_words, const int16_t *frag_mvs, const uint8_t *dct_tokens,
                                  size_t token_capacity);
/* As ocg_dec_flush (one graph launch, asynchronous, ocg_dec_wait for completion). */
OCG_API int  ocg_dec_flush_tokens(ocg_ctx *ctx, const ocg_dec_tokens *t, uint8_t *host_out, int out_mode);

/* ---- decode: device-resident frames, batched over independent streams ---- */
OCG_API int  ocg_pack_create(ocg_pack **out, const ocg_dec_frame *frames, int nframes, int nfrags,
                             int device);
OCG_API void ocg_pack_destroy(ocg_pack *p);
OCG_API int  ocg_pack_nframes(const ocg_pack *p);
/* One launch set for n independent (ctx, pack, frame) jobs of equal geometry on
   `stream` (cudaStream_t, NULL = ctx[0]'s stream).  Asynchronous. */
OCG_API int  ocg_dec_run_batch(ocg_ctx *const *ctxs, ocg_pack *const *packs,
                               const int32_t *frame_idx, int n, void *stream);
/* Stage selection for profiling/tests: bit0 recon+copy, bit1 loop filter,
   bit2 border fill.  Default 7. */
OCG_API void ocg_set_stage_mask(int mask);
OCG_API long ocg_launch_count(void);   /* kernels launched by this library so far */
/* Loop-filter variant: 0 (default) = per-thread global loads/stores, 1 = the
   input tile of each CTA is fetched by one TMA (cp.async.bulk.tensor.2d) into
   shared memory.  Both are bit-exact; on B200 the TMA variant measured slower
   (67 vs 48 us per 32-frame launch: 4 KB boxes are too small to amortise the
   TMA issue cost), so it is opt-in.  Call before creating contexts. */
OCG_API void ocg_set_lf_tma(int on);
/* How ocg_dec_flush / ocg_dec_flush_tokens hand the picture (OCG_OUT_PICTURE) to the host: 0 = a kernel
   writes it through mapped host memory (the whole flush stays one graph of kernels), 1 = three 2-D copies
   by the copy engines behind the graph, then the completion flag; n > 1 = one context in n does (the others
   keep the kernel).  Measured: no variant beats the kernel. */
OCG_API void ocg_set_out_dma(int on);
/* Wait policy of ocg_dec_wait (and of ocg_ctx_sync: 0 = cudaStreamSynchronize, non-zero = a blocking event):
   0 (default) spin on the completion flag; 1 sched_yield between looks; 2 SLEEP: one poller thread per
   process watches the flags of every sleeping waiter and wakes it (semaphore).  With more stream threads
   than cores -- the flush is asynchronous, so a thread mostly waits while the device works on its frame --
   2 keeps the cores on entropy decoding. */
OCG_API void ocg_set_blocking_sync(int policy);
/* Test hook for policy 2 (no device involved): sleeps until *flag has reached seq; 1 = it has, 0 = timed out. */
OCG_API int  ocg_test_sleep_until(volatile uint32_t *flag, uint32_t seq, int timeout_s);
/* Per-stage device timing with CUDA events on the launching stream.  While
   enabled every stage launch (0 recon+copy, 1 loop filter, 2 borders) is
   bracketed by an event pair; collect() waits for them and returns the summed
   milliseconds and launch counts since the previous collect. */
OCG_API void ocg_profile_enable(int on);
OCG_API int  ocg_profile_collect(double ms[3], long launches[3]);

/* ---- encode-side batched block kernels (encint.h:292-326) ---------------- */
/* Fragment descriptor for the encoder kernels: where the source block is, and
   (optionally) one or two predictor blocks (frag_sub / frag_satd2 style). */
typedef struct ocg_enc_frag {
  int32_t src_off;     /* byte offset of the block's row 0 in the source buffer  */
  int32_t ref_off0;    /* predictor 1 offset, or INT32_MIN for intra (sub_128)   */
  int32_t ref_off1;    /* predictor 2 offset, or INT32_MIN for single-tap        */
  int32_t aux;         /* quantiser selector: pli | qti<<2 | qii<<3              */
} ocg_enc_frag;

/* metric selector for ocg_enc_metrics_batch */
#define OCG_MET_SAD        0  /* oc_enc_frag_sad_c / sad2 when ref_off1 valid (encfrag.c:42-90, no early out) */
#define OCG_MET_SATD       1  /* oc_enc_frag_satd_c / satd2_c (encfrag.c:306-320): out = satd, dc         */
#define OCG_MET_INTRA_SATD 2  /* oc_enc_frag_intra_satd_c (encfrag.c:322)                                  */
#define OCG_MET_SSD        3  /* oc_enc_frag_ssd_c (encfrag.c:338)                                         */
#define OCG_MET_INTRA_SAD  4  /* oc_enc_frag_intra_sad_c (encfrag.c:86)                                    */
#define OCG_MET_BORDER_SSD 5  /* oc_enc_frag_border_ssd_c (encfrag.c:352): the 64-bit pixel mask (bit i = row
                                 i>>3, column i&7 in traversal order) travels in ref_off1 (low) / aux (high) */
#define OCG_MET_ACTIVITY   6  /* oc_mb_activity per luma block (analyze.c:1167-1234): out = activity (flat clamp,
                                 edge test, act_th*(act/act_th)^0.7 via mathops.c:294-313), dc = pixel sum (the
                                 block's share of the function's `luma` return value); reads a 10x10 window */
#define OCG_MET_SAD_THRESH 7  /* oc_enc_frag_sad_thresh_c / sad2_thresh_c (encfrag.c:55-84) with the early out: the
                                 running sum after the first row that takes it above the threshold (aux, unsigned),
                                 the whole sum if none does                                                      */

/* All encoder entry points take DEVICE pointers for frames and lists (the
   caller owns residency) and run on `stream`. */
OCG_API int ocg_enc_metrics_batch(int metric, const uint8_t *src_base, const uint8_t *ref_base,
                                  int ystride, const ocg_enc_frag *frags, int n,
                                  uint32_t *out_val, int32_t *out_dc, void *stream);
/* sub/sub_128 -> fDCT (fdct.c:128) -> quantise (enquant.c:220): writes dct[n][64]
   and qdct[n][64] in zig-zag order plus the last nonzero zzi per block.
   dequant/enquant tables: [3 pli][2 qti][3 qii][64] u16 and {m,l} int16 pairs. */
OCG_API int ocg_enc_fdct_quant_batch(const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                                     const ocg_enc_frag *frags, int n,
                                     const uint16_t *dequant, const int16_t *enquant,
                                     int16_t *dct, int16_t *qdct, int32_t *nonzero, void *stream);
/* oc_mcenc_search_frame's full-pel search (mcenc.c:268-515) for a batch of
   macro blocks whose candidate sets are already known (the candidates depend
   on already-searched neighbours, mcenc.c:90-164, so the host -- or a
   wave-front schedule -- supplies them): median predictor, set A, set B, the
   square-pattern descent ("diamond step", mcenc.c:399-431), the per-block 4MV
   descent (PREV only, 437-499) and the final SATD on the reconstructed
   reference (502-513).  SAD on the ORIGINAL frames as in mcenc.c:314-316. */
typedef struct ocg_mb_search_in {
  int32_t  frag_off[4];  /* frag_buf_offs[mb_maps[mbi][0][0..3]]                    */
  int8_t   cand[13][2];  /* oc_mcenc_ctx.candidates, half-pel units, [0] = median   */
  uint8_t  setb0;        /* end of set A                                            */
  uint8_t  ncand;        /* end of set B                                            */
  uint16_t t2_base;      /* max of error[frame] over the MB and its first <=3 cneighbors (mcenc.c:333-337) */
  uint8_t  is_prev;      /* frame==OC_FRAME_PREV: block vectors + 4MV search        */
  uint8_t  pad;
} ocg_mb_search_in;      /* 48 bytes */

typedef struct ocg_mb_search_out {
  int8_t   best_vec[2];     /* full-pel; analysis_mv[0][frame] = OC_MV(2x,2y)       */
  uint16_t error;           /* embs[mbi].error[frame]                               */
  uint32_t satd;            /* embs[mbi].satd[frame]                                */
  int8_t   block_vec[4][2]; /* block_mv[bi] = OC_MV(2x,2y) (is_prev only)           */
  uint32_t block_satd[4];   /* block_satd[bi]                (is_prev only)         */
} ocg_mb_search_out;     /* 32 bytes */

/* Device pointers; src = OC_FRAME_IO, ref_full = the *_ORIG frame searched with
   SAD, ref_satd = the reconstructed reference used for the final SATD. */
OCG_API int ocg_mcenc_search_batch(const uint8_t *src_base, const uint8_t *ref_full_base,
                                   const uint8_t *ref_satd_base, int ystride,
                                   const ocg_mb_search_in *in, ocg_mb_search_out *out, int n,
                                   void *stream);

/* Half-pel refinement of the vectors the full-pel search found:
   oc_mcenc_refine1mv (mcenc.c:661-670, via oc_mcenc_ysatd_halfpel_mbrefine,
   606-659) and oc_mcenc_refine4mv (mcenc.c:762-791, via ..._brefine, 713-760).
   Unlike the search, the refinement of a macro block depends on nothing but its
   own full-pel result, so a whole frame is one batch.  For each of the 8
   half-pel sites around the full-pel vector the two-tap predictor is scored
   with SATD2 + |dc| (or SAD2 when OCG_REFINE_SAD is set: speed level >=
   OC_SP_LEVEL_NOSATD, 1MV only); a site replaces the current best only if it is
   strictly better, sites visited in the order of OC_SQUARE_SITES[0]. */
typedef struct ocg_mb_refine_in {
  int32_t  frag_off[4];     /* frag_buf_offs[mb_maps[mbi][0][0..3]]                      */
  int8_t   vec[2];          /* OC_DIV2 of analysis_mv[0][frame] (full-pel)               */
  int8_t   block_vec[4][2]; /* OC_DIV2 of block_mv[bi]                                   */
  uint8_t  pad[2];
  uint32_t satd;            /* embs[mbi].satd[frame] on entry                            */
  uint32_t block_satd[4];   /* embs[mbi].block_satd[bi] on entry                         */
} ocg_mb_refine_in;         /* 48 bytes */

typedef struct ocg_mb_refine_out {
  int8_t   mv[2];           /* analysis_mv[0][frame] = OC_MV(mv[0],mv[1]) (half-pel)     */
  int8_t   ref_mv[4][2];    /* embs[mbi].ref_mv[bi]                        (half-pel)     */
  uint8_t  pad[2];
  uint32_t satd;            /* embs[mbi].satd[frame]                                     */
  uint32_t block_satd[4];   /* embs[mbi].block_satd[bi]                                  */
} ocg_mb_refine_out;        /* 32 bytes */

#define OCG_REFINE_1MV 1
#define OCG_REFINE_4MV 2
#define OCG_REFINE_SAD 4
/* Device pointers; src = OC_FRAME_IO, ref = the reconstructed reference frame. */
OCG_API int ocg_mcenc_refine_batch(const uint8_t *src_base, const uint8_t *ref_base, int ystride,
                                   const ocg_mb_refine_in *in, ocg_mb_refine_out *out, int n, int flags,
                                   void *stream);

/* ---- whole-frame motion analysis (BASELINE configs[3]) -------------------- */
/* oc_mcenc_search (mcenc.c:517-548) for EVERY macro block of a frame, in the
   reference's coding order, with the candidate sets derived on the device from
   the already-searched neighbours (mcenc.c:90-164) -- the part the per-batch
   call above leaves to its caller.  The data dependence (a macro block needs
   analysis_mv[0][frame] and error[frame] of its cneighbors, encode.c:1004-1034,
   AFTER their half-pel refinement, analyze.c:2486-2489) is honoured by a
   wave-front: one warp per super-block row and reference frame walks its row
   in coding order and waits on per-macro-block completion flags of the rows
   below it; the two reference frames are independent chains.  Per macro block:
   history rotation (mcenc.c:523-531,534,540-547), sets A/B + median predictor,
   thresholds (mcenc.c:331-342), square-pattern descent, 4MV descent, final
   SATD (or SAD at OC_SP_LEVEL_NOSATD), then oc_mcenc_refine1mv where the
   reference's analysis loop would run it: always against OC_FRAME_PREV in an
   inter frame (analyze.c:2486), against OC_FRAME_GOLD only for the macro blocks
   flagged by the caller (that refinement is decided by the host's mode costs,
   analyze.c:2476-2481).  oc_mcenc_refine4mv (analyze.c:2469) feeds nothing
   else and runs as one batch for all macro blocks when requested. */
typedef struct ocg_me_topo {
  int32_t frag_off[4];   /* frag_buf_offs[mb_maps[mbi][0][0..3]]                        */
  int32_t cn[4];         /* oc_mb_enc_info.cneighbors (macro-block indices)              */
  uint8_t ncn;           /* oc_mb_enc_info.ncneighbors                                   */
  uint8_t valid;         /* mb_modes[mbi]!=OC_MODE_INVALID                               */
  uint8_t pad[6];
} ocg_me_topo;           /* 40 bytes */

/* Per-macro-block analysis state, persistent across frames like oc_mb_enc_info
   (encint.h:352-382).  Vectors use the reference's oc_mv encoding
   ((x&0xFF)|y*256, half-pel units); index [0] = OC_FRAME_GOLD, [1] = OC_FRAME_PREV. */
typedef struct ocg_me_mb {
  int16_t  analysis_mv[3][2]; /* [age][frame]; [0] is refined where refinement ran          */
  uint16_t error[2];          /* error[frame]: 16x16 SAD of the full-pel winner             */
  uint32_t satd[2];           /* satd[frame], after refinement where refinement ran         */
  int16_t  unref_mv[2];       /* analysis_mv[0][frame] as the full-pel search left it       */
  uint32_t unref_satd[2];     /* satd[frame] as the full-pel search left it                 */
  int16_t  block_mv[4];       /* full-pel 4MV vectors (OC_FRAME_PREV only)                  */
  int16_t  ref_mv[4];         /* oc_mcenc_refine4mv result (valid after OCG_ME_REFINE_4MV)  */
  uint32_t block_satd[4];     /* block_satd[bi] as the full-pel search left it              */
  uint32_t ref_block_satd[4]; /* block_satd[bi] after oc_mcenc_refine4mv                    */
  int16_t  gold_ref_mv;       /* OCG_ME_SPEC_GOLD: what oc_mcenc_refine1mv(OC_FRAME_GOLD) would */
  uint16_t pad0;              /*   make of analysis_mv[0][GOLD] / satd[GOLD] (macro blocks whose */
  uint32_t gold_ref_satd;     /*   GOLD vector was not refined inside the chain)                */
  uint8_t  pad[4];
} ocg_me_mb;                  /* 96 bytes */

#define OCG_ME_REFINE_PREV 1   /* inter frame: oc_mcenc_refine1mv(OC_FRAME_PREV) for every MB   */
#define OCG_ME_REFINE_4MV  2   /* oc_mcenc_refine4mv for every MB                                */
#define OCG_ME_NOSATD      4   /* sp_level>=OC_SP_LEVEL_NOSATD: SAD instead of SATD (mcenc.c:233,648) */
#define OCG_ME_FAST        8   /* sp_level>=OC_SP_LEVEL_FAST_ANALYSIS: no block_mv/block_satd (mcenc.c:506) */
#define OCG_ME_DROPPED    16   /* _enc->prevframe_dropped (mcenc.c:523)                          */
#define OCG_ME_SPEC_GOLD  32   /* also compute gold_ref_mv / gold_ref_satd for every macro block */

typedef struct ocg_me ocg_me;
/* Macro blocks of the frame in the reference's numbering (4 per luma super
   block, invalid ones included): the array sizes used below. */
OCG_API int  ocg_me_nmbs(const ocg_geometry *g);
/* Native restatement of the reference's tables (state.c:300-330 mb_maps,
   encode.c:967-1048 cneighbors; macro blocks outside the coded region are
   invalid and never listed as neighbours). */
OCG_API int  ocg_me_topology(const ocg_geometry *g, ocg_me_topo *topo);
/* topo may be NULL (derive it) or the caller's copy of the reference's tables. */
OCG_API int  ocg_me_create(ocg_me **out, ocg_ctx *ctx, const ocg_me_topo *topo);
OCG_API void ocg_me_destroy(ocg_me *me);
/* bufs: pool buffer indices of {OC_FRAME_IO, OC_FRAME_PREV_ORIG, OC_FRAME_GOLD_ORIG,
   OC_FRAME_PREV, OC_FRAME_GOLD}.  gold_refine: HOST array of ocg_me_nmbs bytes
   (non-zero = refine that macro block's GOLD vector) or NULL.  Asynchronous on
   the context's stream. */
OCG_API int  ocg_me_frame(ocg_me *me, const int bufs[5], int flags, const uint8_t *gold_refine);
/* The same for n independent streams in ONE launch set on `stream`
   (bufs = n x 5 indices; no GOLD refinement flags). */
OCG_API int  ocg_me_frame_batch(ocg_me *const *mes, const int *bufs, int n, int flags, void *stream);
/* Copy the state out (after the queued work has finished) / seed it. */
OCG_API int  ocg_me_read(ocg_me *me, ocg_me_mb *out);
OCG_API int  ocg_me_write(ocg_me *me, const ocg_me_mb *in);
/* The same without the stream synchronisation (page-locked arrays; the caller orders them, e.g. with
   ocg_ctx_sync). */
OCG_API int  ocg_me_read_async(ocg_me *me, ocg_me_mb *out);
OCG_API int  ocg_me_write_async(ocg_me *me, const ocg_me_mb *in);
/* One macro block again, against reference frame `frame` (0 GOLD, 1 PREV) of the last ocg_me_frame call,
   with the caller's candidate set; refine != 0 adds oc_mcenc_refine1mv of the result.  Synchronous.  For
   callers that run the GOLD chain without refinements (their decision, analyze.c:2476-2485) and must
   redo the macro blocks whose candidates a later refinement changed (mcenc.c:104-110). */
OCG_API int  ocg_me_repair(ocg_me *me, int frame, const ocg_mb_search_in *in, int refine, ocg_mb_search_out *out,
                           ocg_mb_refine_out *rout);
OCG_API int  ocg_ctx_device(const ocg_ctx *ctx);

/* Inter-frame analysis tables (BASELINE configs[3]): everything the reference's analysis loop asks of the
   block kernels that depends only on the input frame, the reference frames and the frame's motion analysis,
   computed for the whole frame behind ocg_me_frame on the same stream:
     oc_enc_frag_intra_satd of every fragment                  (oc_mb_intra_satd, analyze.c:1360-1400)
     oc_enc_frag_ssd / _border_ssd vs the co-located PREV block (oc_skip_cost,     analyze.c:1968-2040)
     oc_enc_frag_satd / _satd2 for OCG_ENC_NCAND candidate predictors per fragment (oc_cost_inter*,
       analyze.c:2062-2286): NOMV, GOLDEN_NOMV, the unrefined and refined PREV and GOLD vectors, the
       unrefined and refined 4MV block vectors.
   A candidate is identified by what the hook is called with -- the predictor's tap offsets relative to the
   frame pool (buffer 0's bottom-left luma pixel, i.e. ocg_ctx_frame_devptr(ctx,0)+base_off on the device,
   ref_frame_handle+base_off in the reference) -- so a hit is exact whichever mode asks.  Asynchronous;
   the tables are complete after ocg_ctx_sync. */
#define OCG_ENC_NCAND 8
typedef struct ocg_enc_inter ocg_enc_inter;
typedef struct ocg_enc_inter_tables {
  const uint32_t *intra_satd;   /* [nfrags]                                                     */
  const int32_t  *intra_dc;     /* [nfrags]                                                     */
  const uint32_t *skip_ssd;     /* [nfrags] plain SSD vs PREV                                   */
  const uint32_t *border_ssd;   /* [nborder] masked SSD vs PREV, indexed by ocg_enc_inter_border_slot */
  int32_t         ncand;        /* OCG_ENC_NCAND, or 0 when the candidates were not requested   */
  int32_t         nluma, nfrags;
  const struct ocg_enc_cand_rec *cand; /* candidate k of fragment f at cand[f * ncand + k]: a fragment's candidates
                                   sit together (two cache lines), because the analysis loop asks for several
                                   predictors of the same fragment in a row                              */
  long            d2h_bytes;
  /* speculative frag_sub + fDCT + quantiser (oc_enc_block_transform_quantize, analyze.c:704-782) against the
     predictors of candidates OCG_ENC_FQ_CAND0/1/2, for every fragment and each of the frame's fq_nqis inter
     quantisers (0: not produced).  fq_desc[sel * nfrags + fragi]; the arrays of an entry sit back to back in
     fq_pool at 8 * off int16: (count+7)/8*8 coefficients of the transform output, then as many per
     quantiser; everything beyond `count` is zero for every quantiser.  off == UINT32_MAX: not available. */
  int32_t         fq_nqis;
  const struct ocg_enc_fq_desc *fq_desc;
  const int16_t  *fq_pool;
} ocg_enc_inter_tables;
typedef struct ocg_enc_cand_rec {
  int32_t  ref_off0, ref_off1;  /* predictor tap offsets relative to the frame pool (INT32_MIN: one tap) */
  uint32_t satd;                /* oc_enc_frag_satd / satd2 against that predictor                      */
  int32_t  dc;
} ocg_enc_cand_rec;
typedef struct ocg_enc_fq_desc {
  uint32_t off;
  uint8_t  count;      /* leading zig-zag coefficients stored (1..64) */
  uint8_t  nz[3];      /* oc_enc_quantize's return value per quantiser */
} ocg_enc_fq_desc;
#define OCG_ENC_FQ_NSEL  3
#define OCG_ENC_FQ_CAND0 0   /* PREV (0,0)              */
#define OCG_ENC_FQ_CAND1 3   /* PREV, refined vector    */
#define OCG_ENC_FQ_CAND2 2   /* PREV, unrefined vector  */
/* mbfrags: [ocg_me_nmbs][12] = state.mb_maps[mbi][pli][bi] (-1: absent); border_fragi/border_mask: the
   fragments with a border mask (state.borders[frags[i].borderi].mask). */
OCG_API int  ocg_enc_inter_create(ocg_enc_inter **out, ocg_ctx *ctx, ocg_me *me, const int32_t *mbfrags,
                                  const int32_t *border_fragi, const int64_t *border_mask, int nborder);
OCG_API void ocg_enc_inter_destroy(ocg_enc_inter *ei);
OCG_API int  ocg_enc_inter_border_slot(const ocg_enc_inter *ei, int fragi);
/* The frame's inter quantiser tables for the speculative transform (layout of ocg_enc_fdct_quant_batch; qti 1
   is read); NULL / nqis 0 switches it off for the next prepass.  Call before ocg_enc_inter_prepass. */
OCG_API int  ocg_enc_inter_quant_tables(ocg_enc_inter *ei, const uint16_t *dequant, const int16_t *enquant, int nqis);
OCG_API int  ocg_enc_inter_prepass(ocg_enc_inter *ei, int io_buf, int prev_buf, int gold_buf, int with_cands,
                                   ocg_enc_inter_tables *out);
/* After the stream has drained (ocg_ctx_sync): fetches the filled part of the coefficient pool.  Synchronous. */
OCG_API int  ocg_enc_inter_finish(ocg_enc_inter *ei, ocg_enc_inter_tables *out);

/* Intra-frame analysis pre-pass (BASELINE config "intra-only encode").  The
   per-block encoder hooks return their result synchronously to serial host
   code (encint.h:292-325), so they cannot launch kernels; but in an INTRA frame
   every value they return depends on the input frame and the frame's quantiser
   tables only.  This call, made once per frame from the enquant_table_fixup
   hook (analyze.c:564, the first hook after the input frame is in place),
   uploads the padded input frame (OC_FRAME_IO) into pool buffer io_buf and
   computes, for EVERY fragment of the frame,
     oc_enc_frag_intra_satd                       (encfrag.c:322)
     oc_enc_frag_sub_128 + oc_enc_fdct8x8         (encfrag.c:33, fdct.c:128)
     oc_enc_quantize for each of the nqis quantisers of the frame (enquant.c:220)
   into pinned host tables owned by the context; the hooks then only look their
   answers up.  dequant/enquant use the layout of ocg_enc_fdct_quant_batch
   ([3 pli][2 qti][3 qii][64]); only qti 0 is read.  Synchronous: returns when
   the tables are complete. */
typedef struct ocg_enc_intra_tables {
  const uint32_t *satd;     /* [nfrags]           oc_enc_frag_intra_satd return value   */
  const int32_t  *satd_dc;  /* [nfrags]           its *_dc output                       */
  const int16_t  *dct;      /* [nfrags][64]       oc_enc_fdct8x8 output (zig-zag order) */
  const int16_t  *qdct;     /* [nqis][nfrags][64] oc_enc_quantize output                */
  const int32_t  *nonzero;  /* [nqis][nfrags]     its return value                      */
} ocg_enc_intra_tables;
/* Allocates the pre-pass buffers now instead of on the first frame. */
OCG_API int ocg_enc_intra_reserve(ocg_ctx *ctx);
OCG_API int ocg_enc_intra_prepass(ocg_ctx *ctx, int io_buf, const uint8_t *host_frame,
                                  const uint16_t *dequant, const int16_t *enquant, int nqis,
                                  ocg_enc_intra_tables *out);

/* ---- decoder post-processing --------------------------------------------------------------
   The reference's out-of-loop filters, lib/decode.c:1609-1957 -- oc_filter_hedge / oc_filter_vedge under
   oc_dec_deblock_frag_rows (:1700) and oc_dering_block under oc_dec_dering_frag_rows (:1892) -- applied on
   the device to the frame a flush left in buffer `self_buf`, into a frame of the layout of the reference's
   pp_frame_data (planes back to back, W x H each, top row first; decode.c:1283-1315).
   level is the reference's pp_level (decint.h: 2 de-block luma, 3 + de-ring luma, 4 strong de-ringing, 5..7
   the same for chroma on top); dc_scale / sharp_mod are its pp_dc_scale[64] / pp_sharp_mod[64]; dc_qis[nfrags]
   is the per-fragment DC quantiser index it tracks (decode.c:1204-1243) and qis[nfrags] state.qis[frag.qii].
   ocg_pp_run queues the kernels on the context's stream; ocg_pp_download waits and copies the planes the level
   processed (luma; chroma from level 5) to host_dst. */
typedef struct ocg_pp ocg_pp;
OCG_API int  ocg_pp_create(ocg_pp **out, ocg_ctx *ctx);
OCG_API void ocg_pp_destroy(ocg_pp *pp);
OCG_API int  ocg_pp_run(ocg_pp *pp, int self_buf, int level, const int32_t *dc_scale, const int32_t *sharp_mod,
                        const uint8_t *dc_qis, const uint8_t *qis);
OCG_API int  ocg_pp_download(ocg_pp *pp, uint8_t *host_dst);
/* test hook: the per-fragment variances of the last run (the reference's dec->variances) */
OCG_API int  ocg_pp_download_variances(ocg_pp *pp, int32_t *host_dst);

#ifdef __cplusplus
}
#endif
#endif /* THEORA_B200_H */
