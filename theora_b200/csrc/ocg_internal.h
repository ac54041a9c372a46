/* Internal declarations shared by the C-ABI implementation (ocg_api.cu) and the
 * kernel translation units.  Not installed. */
#ifndef OCG_INTERNAL_H
#define OCG_INTERNAL_H
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/theora_b200.h"

/* Geometry as the kernels see it (passed by value as a kernel parameter). */
struct OcgPlaneDev {
  int32_t nhfrags, nvfrags, froffset;
  int32_t ystride;    /* negative */
  int32_t width, height, hpad, vpad;
  int32_t plane_off;  /* bottom-left pixel relative to the buffer's luma base */
  int32_t lo_off;     /* lowest offset belonging to this plane (its top-left pixel) */
  int32_t cell_row0;  /* first loop-filter cell row of this plane in the fused row index */
};

struct OcgGeomDev {
  OcgPlaneDev p[3];
  int32_t qx, qy;       /* chroma decimated horizontally / vertically */
  int32_t cell_rows;    /* sum over planes of (nvfrags+1) */
  int32_t max_cells_x;  /* max over planes of (nhfrags+1) */
  int32_t nfrags;
};

/* One (context, frame) job of a batch; lives in device memory. */
struct OcgJobDev {
  uint8_t            *base[3];   /* GOLD, PREV, SELF: buffer + base_off */
  const ocg_frag_rec *recs;      /* nfrags, fragment-index order */
  const int16_t      *rows;
  uint8_t            *coded;     /* nfrags bytes: written by recon pass A, read by the loop filter */
  int32_t            *xlist;     /* nfrags: fragments needing a transform (pass A -> pass B) */
  int32_t            *xcount;    /* [0] list length; 0 between frames (cleared by the border kernel) */
  int32_t             lf_limit;
  uint16_t            dcq[3][2];
  int32_t             dc_residual; /* 1: recs[].dc are DC-prediction residuals: run ocg_dc_unpredict_kernel first */
  int16_t            *dc_tmp;     /* nfrags scratch for planes whose DC values do not fit shared memory */
  const CUtensorMap  *lf_tmaps;  /* 3 tensor maps (one per plane) of the SELF buffer, or NULL: no TMA path */
  uint32_t            seq;       /* ocg_dec_flush: published in host memory when the frame is complete */
  int32_t             ncoeff_rows; /* ocg_dec_flush: rows to stage in */
  uint8_t            *host_out;  /* ocg_dec_flush: the device's address of the page-locked destination buffer */
  int32_t             dense_rows; /* rows live at coeff_row + r (device-side expansion) and are cleared once read */
  int32_t             intra_frame;
};

/* One frame's token lists as the reference's decoder holds them (ocg_dec_flush_tokens); device copy. */
struct OcgExpandDev {
  int32_t ti0[3][64];       /* byte offset of each (plane, zig-zag index) list in the token array */
  int32_t eob_runs[3][64];  /* EOB run reaching into each list (saturated) */
  int32_t ntoken_bytes;
  int32_t qis[3];
  int32_t nqis;
  int32_t pad[3];
};

struct OcgExpandBufs {
  const int32_t  *order;    /* [nfrags] coded order (per plane, planes back to back) -> fragment index */
  const int32_t  *buf_off;  /* [nfrags] */
  const uint16_t *dequant;  /* [64 qi][3][2][64] */
  const uint32_t *words;    /* device copies of the frame's arrays */
  const int16_t  *mvs;
  const uint8_t  *tokens;
  uint32_t *tok, *cov;      /* decoded tokens / fragments served before each, indexed from the list's byte offset */
  int32_t  *ntok;           /* [192] */
  int16_t  *coef;           /* [nfrags][64] dense blocks, all-zero between frames */
  uint8_t  *nextz, *lastz, *rmask;
};

void ocg_expand_init_tables(cudaStream_t st);
void ocg_launch_stage_tokens(const OcgJobDev *h_job, OcgJobDev *d_job, const OcgExpandDev *h_x, OcgExpandDev *d_x,
                             const void *h_words, void *d_words, int words_bytes, const void *h_mvs, void *d_mvs, int mvs_bytes,
                             const void *h_tok, void *d_tok, cudaStream_t st);
void ocg_launch_expand(const OcgGeomDev &g, const OcgExpandDev *d_x, const OcgExpandBufs &B, const int16_t *dc_final,
                       const OcgJobDev *d_job, ocg_frag_rec *d_recs, cudaStream_t st);

#define OCG_FRAGS_PER_BLOCK 64
#define OCG_RECON_THREADS   256

void ocg_launch_recon(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, cudaStream_t st);
int  ocg_launch_dc_unpredict(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, cudaStream_t st); /* <0: unsupported size */
/* the same recurrence ahead of the frame's lists: packed oc_fragment words in, final DC array out */
int  ocg_launch_dc_unpredict_words(const OcgGeomDev &g, const uint32_t *words, int16_t *dc_final, int16_t *dc_tmp,
                                   cudaStream_t st);
void ocg_launch_dc_patch(ocg_frag_rec *recs, const int16_t *dc_final, int nfrags, cudaStream_t st);
void ocg_launch_xlist_reset(const OcgJobDev *jobs, int njobs, cudaStream_t st);
void ocg_launch_codedmap(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, cudaStream_t st);
void ocg_launch_loop_filter(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, bool use_tma, cudaStream_t st);
void ocg_launch_borders(const OcgGeomDev &g, const OcgJobDev *jobs, int njobs, cudaStream_t st);

/* flush graph (ocg_dec_flush): lists in / picture out through mapped host memory, kernels only */
void ocg_launch_stage_in(const OcgJobDev *h_job, OcgJobDev *d_job, const ocg_frag_rec *h_recs, ocg_frag_rec *d_recs,
                         int nfrags, const int16_t *h_rows, int16_t *d_rows, cudaStream_t st);
void ocg_launch_copy_out(const ocg_geometry &g, int out_mode, const OcgJobDev *job, uint32_t *counter, uint32_t *host_flag,
                         cudaStream_t st);
void ocg_init_device_tables(cudaStream_t st); /* idempotent; call once per context */

/* ocg_pool.cu: caching allocators.  The translation units of the library reach them through the CUDA
   runtime's own names (below), so every block an instance owns comes from the cache. */
cudaError_t ocg_pool_host_alloc(void **pp, size_t size, unsigned flags);
cudaError_t ocg_pool_host_free(void *p);
cudaError_t ocg_pool_dev_alloc(void **pp, size_t size);
cudaError_t ocg_pool_dev_free(void *p);
#ifndef OCG_POOL_IMPLEMENTATION
#define cudaHostAlloc(pp, size, flags) ocg_pool_host_alloc((void **)(pp), (size), (flags))
#define cudaFreeHost(p) ocg_pool_host_free((void *)(p))
#define cudaMalloc(pp, size) ocg_pool_dev_alloc((void **)(pp), (size))
#define cudaFree(p) ocg_pool_dev_free((void *)(p))
#endif

void ocg_count_launch(int n);
cudaError_t ocg_set_device(int device); /* cudaSetDevice unless it is already the thread's device */

#endif
