/* C-ABI implementation: contexts, pinned staging, device-resident packs and the
 * launch sequences.  See include/theora_b200.h for the contract. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <new>
#include <vector>
#include <stdint.h>
#include <sched.h>
#include <semaphore.h>
#include <pthread.h>
#include <condition_variable>
#include <time.h>
#include "ocg_internal.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<long> g_launches{0};
std::atomic<int> g_stage_mask{7};
std::atomic<int> g_blocking_sync{0};
std::atomic<int> g_out_dma{0}; /* ocg_set_out_dma: the picture leaves through the copy engines instead of the copy-out kernel */
std::atomic<int> g_use_tma{0}; /* measured slower than per-thread loads on B200 (67 vs 48 us): opt-in */

int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
  if (e != cudaSuccess) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  else snprintf(g_err, sizeof(g_err), "%s", what);
  return code;
}

#define CU(call)                                              \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return fail(OCG_ECUDA, #call, e_); \
  } while (0)

constexpr int kSlots = 2;

} /* namespace */

/* cudaSetDevice is not free even when nothing changes; cudaGetDevice only reads the runtime's thread state */
cudaError_t ocg_set_device(int device) {
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == device) return cudaSuccess;
  return cudaSetDevice(device);
}

namespace {

struct Slot {
  ocg_frag_rec *recs = nullptr; /* pinned, nfrags */
  int16_t *rows = nullptr;      /* pinned, nfrags*8 rows */
  OcgJobDev *job = nullptr;     /* pinned */
  cudaEvent_t consumed = nullptr;
  bool busy = false;
  uint32_t flush_seq = 0;       /* ocg_dec_flush: the slot is free once the context's done flag reaches this */
  bool flush_busy = false;
  /* the device's addresses of the three pinned regions above (mapped) */
  ocg_frag_rec *m_recs = nullptr;
  int16_t *m_rows = nullptr;
  OcgJobDev *m_job = nullptr;
  OcgExpandDev *xjob = nullptr, *m_xjob = nullptr; /* token path header: pinned / its device address */
};

/* One instantiated CUDA graph of a whole frame flush (ocg_dec_flush). */
struct FlushGraph {
  int slot, out_mode, dc, lf, tokens; /* tokens: the ocg_dec_flush_tokens variant */
  cudaGraphExec_t exec;
  int kernels;
};
struct MappedOut { uint8_t *host, *dev; }; /* page-locked destination buffers seen so far */

} /* namespace */

/* buffers of the intra-frame encoder pre-pass (allocated on first use) */
struct EncPre {
  ocg_enc_frag *d_frags = nullptr; /* [3 qii][nfrags] */
  uint16_t *d_dequant = nullptr;   /* [3][2][3][64] */
  int16_t *d_enquant = nullptr;    /* [3][2][3][64][2] */
  uint8_t *d_out = nullptr;        /* one device block mirrored by h_out */
  uint8_t *h_out = nullptr;        /* pinned */
  uint8_t *h_tabs = nullptr;       /* pinned staging for the two quantiser tables */
  size_t off_satd = 0, off_dc = 0, off_dct = 0, off_qdct = 0, off_nz = 0, out_sz = 0;
};

struct ocg_ctx {
  int serial = 0; /* creation order in the process */
  EncPre *enc = nullptr;
  ocg_geometry geom;
  OcgGeomDev gdev;
  int device = 0;
  cudaStream_t stream = nullptr;
  uint8_t *frames = nullptr; /* nrefs * ref_frame_sz (+slack) */
  /* device-side lists for single-frame submits */
  ocg_frag_rec *d_recs = nullptr;
  int16_t *d_rows = nullptr;
  uint8_t *d_map = nullptr;  /* coded map, produced by the recon kernel */
  int16_t *d_dc_tmp = nullptr; /* DC wave-front scratch (planes too large for shared memory) */
  uint32_t *d_dc_words = nullptr; /* ocg_dec_dc_begin: the decoder's packed fragment words */
  int16_t *d_dc_final = nullptr;  /* ... and the final DC values computed from them */
  bool dc_ahead = false;          /* ocg_dec_dc_begin ran for the frame being assembled */
  int32_t *d_xlist = nullptr; /* transform work list + 2 counters behind it */
  CUtensorMap *d_tmaps = nullptr; /* [nrefs][3] tiled views of the padded planes for the TMA loop filter */
  OcgJobDev *d_job = nullptr;
  ocg_frag_rec *tmpl = nullptr; /* host: every fragment uncoded, buf_off/plane filled in */
  Slot slots[kSlots];
  cudaEvent_t done = nullptr; /* cudaEventBlockingSync: ocg_ctx_sync can sleep instead of spinning */
  int cur_slot = 0;      /* slot handed out by the last ocg_dec_staging */
  bool staged = false;
  /* ocg_dec_flush / ocg_dec_wait */
  std::vector<FlushGraph> graphs;
  std::vector<MappedOut> outs;
  volatile uint32_t *h_done = nullptr; /* pinned + mapped: sequence number of the last finished flush */
  uint32_t *d_done = nullptr;          /* the device's address of the same word */
  uint32_t *d_out_counter = nullptr;   /* copy-out kernel: CTAs finished */
  uint32_t flush_seq = 0;
  /* device-side token expansion (ocg_dec_expand_setup / ocg_dec_flush_tokens) */
  struct Expand {
    int32_t *d_order = nullptr, *d_buf_off = nullptr, *d_ntok = nullptr;
    uint16_t *d_dequant = nullptr;
    uint8_t *d_words_raw = nullptr, *d_mvs_raw = nullptr, *d_tokens_raw = nullptr; /* + up to 15 bytes of host misalignment */
    uint32_t *d_tok = nullptr, *d_cov = nullptr;
    int16_t *d_coef = nullptr;
    uint8_t *d_nextz = nullptr, *d_lastz = nullptr, *d_rmask = nullptr;
    OcgExpandDev *d_xjob = nullptr;
    size_t token_cap = 0;
    /* the caller's page-locked arrays and their device addresses (fixed for the context's life) */
    const uint32_t *h_words = nullptr; const int16_t *h_mvs = nullptr; const uint8_t *h_tokens = nullptr;
    const uint8_t *m_words = nullptr, *m_mvs = nullptr, *m_tokens = nullptr;
    bool ready = false;
  } x;
  long nflush = 0;            /* ocg_dec_flush calls so far */
  int graph_after = 16;       /* replay a CUDA graph from this flush on; < 0: never */
};

struct ocg_pack {
  int device = 0;
  int nframes = 0;
  int nfrags = 0;
  std::vector<ocg_dec_frame> frames; /* pointers are device pointers */
  uint8_t *blob = nullptr;
  size_t blob_sz = 0;
};

void ocg_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

/* ------------------------------------------------------------------------ */
static void geom_to_dev(const ocg_geometry &g, OcgGeomDev &d) {
  memset(&d, 0, sizeof(d));
  int cell_rows = 0, maxcx = 0;
  for (int pli = 0; pli < 3; pli++) {
    const ocg_plane_geom &p = g.planes[pli];
    OcgPlaneDev &q = d.p[pli];
    q.nhfrags = p.nhfrags;
    q.nvfrags = p.nvfrags;
    q.froffset = p.froffset;
    q.ystride = p.ystride;
    q.width = p.width;
    q.height = p.height;
    q.hpad = p.hpad;
    q.vpad = p.vpad;
    q.plane_off = (int32_t)p.plane_off;
    q.lo_off = (int32_t)(p.plane_off + (int64_t)(p.height - 1) * p.ystride);
    q.cell_row0 = cell_rows;
    cell_rows += p.nvfrags + 1;
    if (p.nhfrags + 1 > maxcx) maxcx = p.nhfrags + 1;
  }
  d.qx = !(g.pixel_fmt & 1);
  d.qy = !(g.pixel_fmt & 2);
  d.cell_rows = cell_rows;
  d.max_cells_x = maxcx;
  d.nfrags = g.nfrags;
}

static void fill_job(OcgJobDev &j, const ocg_ctx *c, const ocg_dec_frame &f, const ocg_frag_rec *recs,
                     const int16_t *rows) {
  memset(&j, 0, sizeof(j));
  for (int i = 0; i < 3; i++)
    j.base[i] = f.ref_idx[i] >= 0 ? c->frames + (int64_t)f.ref_idx[i] * c->geom.ref_frame_sz + c->geom.base_off
                                  : nullptr;
  j.recs = recs;
  j.rows = rows;
  j.coded = c->d_map;
  j.xlist = c->d_xlist;
  j.xcount = c->d_xlist + c->geom.nfrags;
  j.lf_limit = f.lf_limit;
  j.intra_frame = f.intra_frame;
  j.dc_residual = f.dc_residual == 1;
  j.dc_tmp = c->d_dc_tmp;
  j.lf_tmaps = c->d_tmaps ? c->d_tmaps + (size_t)f.ref_idx[OCG_FRAME_SELF] * 3 : nullptr;
  for (int p = 0; p < 3; p++)
    for (int q = 0; q < 2; q++) j.dcq[p][q] = f.dc_quant[p][q];
}

extern "C" OCG_API int ocg_dc_unpredict_supported(const ocg_geometry *g) {
  if (g == nullptr) return 0;
  for (int pli = 0; pli < 3; pli++) {
    const size_t nv = (size_t)g->planes[pli].nvfrags, nfr = (size_t)g->planes[pli].nfrags;
    if (nv > 1024 || 8 * nv * sizeof(int) + ((nfr + 15) & ~(size_t)15) > 227 * 1024) return 0;
  }
  return 1;
}

static int check_frame(const ocg_geometry &g, const ocg_dec_frame &f) {
  if (f.ncoded < 0 || f.ncoded > g.nfrags) return fail(OCG_EINVAL, "coded fragment count out of range");
  if (f.ncoeff_rows < 0 || f.ncoeff_rows > (long)g.nfrags * 8) return fail(OCG_EINVAL, "coefficient row count out of range");
  if (f.ref_idx[OCG_FRAME_SELF] < 0 || f.ref_idx[OCG_FRAME_SELF] >= g.nrefs) return fail(OCG_EINVAL, "bad SELF buffer index");
  for (int i = 0; i < 2; i++) {
    if (f.ref_idx[i] >= g.nrefs) return fail(OCG_EINVAL, "bad reference buffer index");
    /* the kernels form base pointers from these: a frame that may copy or predict needs both */
    if (!f.intra_frame && f.ref_idx[i] < 0) return fail(OCG_EINVAL, "inter frame without a GOLD/PREV reference buffer");
  }
  if (f.lf_limit < 0 || f.lf_limit > 127) return fail(OCG_EINVAL, "loop filter limit out of range");
  if (f.dc_residual && !ocg_dc_unpredict_supported(&g)) return fail(OCG_EIMPL, "DC un-prediction on the device does not support this frame size");
  return OCG_OK;
}

/* Tensor maps for the TMA loop filter: each padded plane of each buffer as a 2-D
   tensor of 32-bit words (x granularity 4 bytes = the cell origin's alignment),
   box = 64 cells x 8 rows.  The x origin sits 16 bytes left of the picture for
   every plane so the base address is 16-byte aligned (chroma aprons are 8 wide). */
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int build_tensor_maps(ocg_ctx *c) {
  static EncodeTiledFn encode = nullptr;
  if (encode == nullptr) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
      cudaGetLastError();
      return fail(OCG_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    }
    encode = (EncodeTiledFn)fn;
  }
  const ocg_geometry &g = c->geom;
  /* a 528-byte box row must fit inside one pitch: narrow planes keep the per-thread kernel */
  for (int pli = 0; pli < 3; pli++)
    if (-(int64_t)g.planes[pli].ystride < 528) return OCG_OK;
  std::vector<CUtensorMap> maps((size_t)g.nrefs * 3);
  for (int b = 0; b < g.nrefs; b++) {
    for (int pli = 0; pli < 3; pli++) {
      const ocg_plane_geom &p = g.planes[pli];
      const int64_t stride = -(int64_t)p.ystride;
      /* top-left picture pixel, then up vpad rows and left 16 bytes */
      const int64_t org = (int64_t)b * g.ref_frame_sz + g.base_off + p.plane_off + (int64_t)(p.height - 1) * p.ystride -
                          (int64_t)p.vpad * stride - 16;
      if (org < 0 || (org & 15) || (stride & 15)) return fail(OCG_EIMPL, "plane layout not TMA-addressable");
      /* one tensor row = one full pitch (the row extent may not exceed the pitch) */
      cuuint64_t dims[2] = {(cuuint64_t)(stride / 4), (cuuint64_t)(p.height + 2 * p.vpad)};
      cuuint64_t strides[1] = {(cuuint64_t)stride};
      cuuint32_t box[2] = {132, 8}; /* 12 bytes lead-in + 64 cells + 4: see ocg_lf_tma_kernel */
      cuuint32_t estr[2] = {1, 1};
      CUresult r = encode(&maps[(size_t)b * 3 + pli], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, c->frames + org, dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(OCG_ECUDA, "cuTensorMapEncodeTiled failed");
    }
  }
  CU(cudaMalloc(&c->d_tmaps, maps.size() * sizeof(CUtensorMap)));
  CU(cudaMemcpy(c->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  return OCG_OK;
}

/* Optional per-stage timing with CUDA events on the launching stream (used by
   bench.py for the roofline numbers; off by default). */
struct StageSpan { cudaEvent_t a, b; int stage; };
static std::mutex g_prof_lock;
static std::vector<StageSpan> g_spans;
static std::vector<cudaEvent_t> g_event_pool;
static std::atomic<int> g_profile{0};

static cudaEvent_t prof_event() {
  cudaEvent_t e = nullptr;
  if (!g_event_pool.empty()) { e = g_event_pool.back(); g_event_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}

template <typename F>
static void timed_stage(int stage, cudaStream_t st, F &&launch) {
  if (!g_profile.load(std::memory_order_relaxed)) { launch(); return; }
  std::lock_guard<std::mutex> lk(g_prof_lock);
  StageSpan sp{prof_event(), prof_event(), stage};
  cudaEventRecord(sp.a, st);
  launch();
  cudaEventRecord(sp.b, st);
  g_spans.push_back(sp);
}

static void launch_stages(const OcgGeomDev &gd, const OcgJobDev *jobs, int njobs, bool any_lf, bool use_tma,
                          cudaStream_t st, bool any_dc = false) {
  const int mask = g_stage_mask.load(std::memory_order_relaxed);
  if (any_dc && (mask & 1)) ocg_launch_dc_unpredict(gd, jobs, njobs, st);
  if (mask & 1) timed_stage(0, st, [&] { ocg_launch_recon(gd, jobs, njobs, st); });
  else if ((mask & 2) && any_lf) ocg_launch_codedmap(gd, jobs, njobs, st);
  if ((mask & 2) && any_lf) timed_stage(1, st, [&] { ocg_launch_loop_filter(gd, jobs, njobs, use_tma, st); });
  if (mask & 4) timed_stage(2, st, [&] { ocg_launch_borders(gd, jobs, njobs, st); });
  else if (mask & 1) ocg_launch_xlist_reset(jobs, njobs, st);
}

/* One stream per decoder/encoder instance and one host thread per instance is the intended use; with the
   driver's default of 8 hardware work queues, more than 8 concurrently active streams share queues and a
   flush can sit behind another instance's multi-megabyte copy (measured with 16 encoder threads: 366 vs 637
   frames/s).  Ask for the maximum before the driver initialises, unless the user has chosen a value. */
__attribute__((constructor)) static void ocg_request_work_queues(void) { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

/* ------------------------------------------------------------------------ */
extern "C" {

OCG_API const char *ocg_version(void) { return "theora_b200 0.1 (sm_100a)"; }
OCG_API const char *ocg_last_error(void) { return g_err; }

OCG_API int ocg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

OCG_API void ocg_set_stage_mask(int mask) { g_stage_mask.store(mask & 7); }

OCG_API void ocg_set_out_dma(int on) { g_out_dma.store(on < 0 ? 0 : on); }
extern int g_ocg_lf_legacy;
OCG_API void ocg_set_lf_tma(int on) { g_use_tma.store(on == 1 ? 1 : 0); g_ocg_lf_legacy = on == 2; }

OCG_API void ocg_set_blocking_sync(int policy) { g_blocking_sync.store(policy < 0 ? 0 : (policy > 2 ? 2 : policy)); }

OCG_API void ocg_profile_enable(int on) { g_profile.store(on ? 1 : 0); }

OCG_API int ocg_profile_collect(double ms[3], long launches[3]) {
  if (ms == nullptr || launches == nullptr) return fail(OCG_EFAULT, "NULL argument");
  std::lock_guard<std::mutex> lk(g_prof_lock);
  for (int i = 0; i < 3; i++) { ms[i] = 0.0; launches[i] = 0; }
  for (StageSpan &sp : g_spans) {
    float t = 0.f;
    CU(cudaEventSynchronize(sp.b));
    CU(cudaEventElapsedTime(&t, sp.a, sp.b));
    ms[sp.stage] += (double)t;
    launches[sp.stage]++;
    g_event_pool.push_back(sp.a);
    g_event_pool.push_back(sp.b);
  }
  g_spans.clear();
  return OCG_OK;
}
OCG_API long ocg_launch_count(void) { return g_launches.load(); }

/* state.c:424-470 (fragment planes), 545-671 (padded buffers + flip). */
OCG_API int ocg_geometry_init(ocg_geometry *g, int fw, int fh, int pixel_fmt, int nrefs) {
  if (g == nullptr) return fail(OCG_EFAULT, "NULL geometry");
  if (fw <= 0 || fh <= 0 || (fw & 15) || (fh & 15) || fw >= 0x100000 || fh >= 0x100000)
    return fail(OCG_EINVAL, "frame size must be a positive multiple of 16");
  if (pixel_fmt != 0 && pixel_fmt != 2 && pixel_fmt != 3) return fail(OCG_EINVAL, "unknown pixel format");
  if (nrefs < 3 || nrefs > 6) return fail(OCG_EINVAL, "nrefs must be 3..6");
  const int hdec = !(pixel_fmt & 1), vdec = !(pixel_fmt & 2);
  const int64_t ys = (int64_t)fw + 32, yr = (int64_t)fh + 32;
  const int64_t cs = ((ys >> hdec) + 15) & ~(int64_t)15, cr = yr >> vdec;
  const int64_t ysz = ys * yr, csz = cs * cr;
  const int64_t yorg = 16 + 16 * ys;                          /* luma picture origin          */
  const int64_t corg = (16 >> hdec) + (int64_t)(16 >> vdec) * cs; /* chroma origin inside its plane */
  const int64_t adj = (-corg) & 15;                           /* keeps chroma data 16-aligned */
  memset(g, 0, sizeof(*g));
  g->frame_width = fw;
  g->frame_height = fh;
  g->pixel_fmt = pixel_fmt;
  g->nrefs = nrefs;
  g->ref_frame_sz = ysz + 2 * csz + 16;
  if (g->ref_frame_sz * nrefs > ((int64_t)1 << 31) - 65536)
    return fail(OCG_EIMPL, "frame pool exceeds 32-bit fragment offsets");
  g->base_off = yorg + (int64_t)(fh - 1) * ys;
  const int64_t origin[3] = {yorg, ysz + adj + corg, ysz + adj + csz + corg};
  int fro = 0;
  for (int pli = 0; pli < 3; pli++) {
    ocg_plane_geom &p = g->planes[pli];
    p.width = pli ? fw >> hdec : fw;
    p.height = pli ? fh >> vdec : fh;
    p.nhfrags = p.width >> 3;
    p.nvfrags = p.height >> 3;
    p.froffset = fro;
    p.nfrags = p.nhfrags * p.nvfrags;
    p.ystride = (int32_t)-(pli ? cs : ys);
    p.hpad = pli ? 16 >> hdec : 16;
    p.vpad = pli ? 16 >> vdec : 16;
    /* bottom-left pixel (the reference addresses frames bottom-up) */
    p.plane_off = origin[pli] + (int64_t)(p.height - 1) * (pli ? cs : ys) - g->base_off;
    fro += p.nfrags;
  }
  g->nfrags = fro;
  return OCG_OK;
}

OCG_API void ocg_geometry_frag_buf_offs(const ocg_geometry *g, int32_t *offs) {
  for (int pli = 0; pli < 3; pli++) {
    const ocg_plane_geom &p = g->planes[pli];
    int32_t *o = offs + p.froffset;
    for (int fy = 0; fy < p.nvfrags; fy++)
      for (int fx = 0; fx < p.nhfrags; fx++)
        *o++ = (int32_t)(p.plane_off + (int64_t)fy * 8 * p.ystride + fx * 8);
  }
}

/* ---- context ------------------------------------------------------------ */
OCG_API void ocg_ctx_destroy(ocg_ctx *c) {
  if (c == nullptr) return;
  ocg_set_device(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (Slot &s : c->slots) {
    if (s.recs) cudaFreeHost(s.recs);
    if (s.rows) cudaFreeHost(s.rows);
    if (s.job) cudaFreeHost(s.job);
    if (s.xjob) cudaFreeHost(s.xjob);
    if (s.consumed) cudaEventDestroy(s.consumed);
  }
  if (c->enc != nullptr) {
    cudaFree(c->enc->d_frags);
    cudaFree(c->enc->d_dequant);
    cudaFree(c->enc->d_enquant);
    cudaFree(c->enc->d_out);
    if (c->enc->h_out) cudaFreeHost(c->enc->h_out);
    if (c->enc->h_tabs) cudaFreeHost(c->enc->h_tabs);
    delete c->enc;
  }
  if (c->done) cudaEventDestroy(c->done);
  for (FlushGraph &fg : c->graphs) cudaGraphExecDestroy(fg.exec);
  cudaFree(c->x.d_order); cudaFree(c->x.d_buf_off); cudaFree(c->x.d_ntok); cudaFree(c->x.d_dequant);
  cudaFree(c->x.d_words_raw); cudaFree(c->x.d_mvs_raw); cudaFree(c->x.d_tokens_raw);
  cudaFree(c->x.d_tok); cudaFree(c->x.d_cov); cudaFree(c->x.d_coef);
  cudaFree(c->x.d_nextz); cudaFree(c->x.d_lastz); cudaFree(c->x.d_rmask); cudaFree(c->x.d_xjob);
  if (c->h_done) cudaFreeHost((void *)c->h_done);
  cudaFree(c->d_out_counter);
  cudaFree(c->frames);
  cudaFree(c->d_recs);
  cudaFree(c->d_rows);
  cudaFree(c->d_map);
  cudaFree(c->d_dc_tmp);
  cudaFree(c->d_dc_words);
  cudaFree(c->d_dc_final);
  cudaFree(c->d_xlist);
  cudaFree(c->d_tmaps);
  cudaFree(c->d_job);
  free(c->tmpl);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

OCG_API int ocg_ctx_create(ocg_ctx **out, const ocg_geometry *g, int device) {
  if (out == nullptr || g == nullptr) return fail(OCG_EFAULT, "NULL argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(OCG_ECUDA, "no CUDA device (this library has no CPU fallback)", e);
  if (device < 0 || device >= ndev) return fail(OCG_EINVAL, "device index out of range");
  ocg_geometry chk;
  int r = ocg_geometry_init(&chk, g->frame_width, g->frame_height, g->pixel_fmt, g->nrefs);
  if (r < 0) return r;
  if (memcmp(&chk, g, sizeof(chk)) != 0) return fail(OCG_EINVAL, "geometry was not produced by ocg_geometry_init");
  CU(ocg_set_device(device));
  ocg_ctx *c = new (std::nothrow) ocg_ctx();
  if (c == nullptr) return fail(OCG_ENOMEM, "out of memory");
  c->geom = *g;
  c->device = device;
  {
    static std::atomic<int> next_serial{0};
    c->serial = next_serial.fetch_add(1);
  }
  geom_to_dev(*g, c->gdev);
  const size_t nf = (size_t)g->nfrags;
  const size_t pool = (size_t)g->ref_frame_sz * g->nrefs + 256;
#define CUX(call)                                  \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) {                       \
      ocg_ctx_destroy(c);                          \
      return fail(OCG_ECUDA, #call, e_);           \
    }                                              \
  } while (0)
  CUX(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUX(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming | cudaEventBlockingSync));
  CUX(cudaMalloc(&c->frames, pool));
  CUX(cudaMemsetAsync(c->frames, 0x80, pool, c->stream));
  ocg_init_device_tables(c->stream);
  CUX(cudaMalloc(&c->d_recs, nf * sizeof(ocg_frag_rec)));
  CUX(cudaMalloc(&c->d_rows, nf * 8 * 16));
  CUX(cudaMalloc(&c->d_map, nf));
  CUX(cudaMalloc(&c->d_dc_tmp, nf * sizeof(int16_t)));
  CUX(cudaMalloc(&c->d_dc_words, nf * sizeof(uint32_t)));
  CUX(cudaMalloc(&c->d_dc_final, nf * sizeof(int16_t)));
  CUX(cudaMalloc(&c->d_xlist, (nf + 2) * sizeof(int32_t)));
  CUX(cudaMemsetAsync(c->d_xlist, 0, (nf + 2) * sizeof(int32_t), c->stream));
  CUX(cudaMalloc(&c->d_job, sizeof(OcgJobDev)));
  CUX(cudaHostAlloc((void **)&c->h_done, 64, cudaHostAllocMapped));
  *c->h_done = 0;
  CUX(cudaHostGetDevicePointer((void **)&c->d_done, (void *)c->h_done, 0));
  CUX(cudaMalloc(&c->d_out_counter, sizeof(uint32_t)));
  CUX(cudaMemsetAsync(c->d_out_counter, 0, sizeof(uint32_t), c->stream));
  for (Slot &s : c->slots) {
    /* mapped: ocg_dec_flush's stage-in kernel reads them in place */
    CUX(cudaHostAlloc(&s.recs, nf * sizeof(ocg_frag_rec), cudaHostAllocMapped));
    CUX(cudaHostAlloc(&s.rows, nf * 8 * 16, cudaHostAllocMapped));
    CUX(cudaHostAlloc(&s.job, sizeof(OcgJobDev), cudaHostAllocMapped));
    CUX(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    CUX(cudaHostGetDevicePointer((void **)&s.m_recs, s.recs, 0));
    CUX(cudaHostGetDevicePointer((void **)&s.m_rows, s.rows, 0));
    CUX(cudaHostGetDevicePointer((void **)&s.m_job, s.job, 0));
    CUX(cudaHostAlloc(&s.xjob, sizeof(OcgExpandDev), cudaHostAllocMapped));
    CUX(cudaHostGetDevicePointer((void **)&s.m_xjob, s.xjob, 0));
  }
  /* record template: every fragment uncoded, offsets and planes filled in */
  c->tmpl = (ocg_frag_rec *)calloc(nf, sizeof(ocg_frag_rec));
  if (c->tmpl == nullptr) { ocg_ctx_destroy(c); return fail(OCG_ENOMEM, "out of memory"); }
  {
    std::vector<int32_t> offs(nf);
    ocg_geometry_frag_buf_offs(g, offs.data());
    for (int pli = 0; pli < 3; pli++) {
      const ocg_plane_geom &p = g->planes[pli];
      for (int i = 0; i < p.nfrags; i++) {
        ocg_frag_rec &rc = c->tmpl[p.froffset + i];
        rc.buf_off = offs[(size_t)(p.froffset + i)];
        rc.refi = OCG_FRAG_UNCODED;
        rc.pli_qti = (uint8_t)pli;
      }
    }
  }
  for (Slot &s : c->slots) memcpy(s.recs, c->tmpl, nf * sizeof(ocg_frag_rec));
  if (g_use_tma.load() && build_tensor_maps(c) < 0) { ocg_ctx_destroy(c); return OCG_ECUDA; }
  CUX(cudaStreamSynchronize(c->stream));
#undef CUX
  *out = c;
  return OCG_OK;
}

OCG_API const ocg_geometry *ocg_ctx_geometry(const ocg_ctx *c) { return c ? &c->geom : nullptr; }
OCG_API void *ocg_ctx_stream(ocg_ctx *c) { return c ? (void *)c->stream : nullptr; }
OCG_API int ocg_ctx_device(const ocg_ctx *c) { return c ? c->device : -1; }

OCG_API void *ocg_ctx_frame_devptr(ocg_ctx *c, int buf) {
  if (c == nullptr || buf < 0 || buf >= c->geom.nrefs) return nullptr;
  return c->frames + (size_t)buf * c->geom.ref_frame_sz;
}

OCG_API int ocg_ctx_sync(ocg_ctx *c) {
  if (c == nullptr) return fail(OCG_EFAULT, "NULL context");
  CU(ocg_set_device(c->device));
  if (g_blocking_sync.load()) {
    /* the calling thread sleeps until the stream drains: with more stream threads than
       cores the core runs another stream's entropy decode meanwhile */
    CU(cudaEventRecord(c->done, c->stream));
    CU(cudaEventSynchronize(c->done));
  } else {
    CU(cudaStreamSynchronize(c->stream));
  }
  for (Slot &s : c->slots) s.busy = false;
  return OCG_OK;
}

OCG_API int ocg_ctx_upload_frame(ocg_ctx *c, int buf, const uint8_t *host) {
  if (c == nullptr || host == nullptr) return fail(OCG_EFAULT, "NULL argument");
  if (buf < 0 || buf >= c->geom.nrefs) return fail(OCG_EINVAL, "bad buffer index");
  CU(ocg_set_device(c->device));
  CU(cudaMemcpyAsync(c->frames + (size_t)buf * c->geom.ref_frame_sz, host, (size_t)c->geom.ref_frame_sz,
                     cudaMemcpyHostToDevice, c->stream));
  return OCG_OK;
}

OCG_API int ocg_ctx_download_frame(ocg_ctx *c, int buf, uint8_t *host) {
  if (c == nullptr || host == nullptr) return fail(OCG_EFAULT, "NULL argument");
  if (buf < 0 || buf >= c->geom.nrefs) return fail(OCG_EINVAL, "bad buffer index");
  CU(ocg_set_device(c->device));
  CU(cudaMemcpyAsync(host, c->frames + (size_t)buf * c->geom.ref_frame_sz, (size_t)c->geom.ref_frame_sz,
                     cudaMemcpyDeviceToHost, c->stream));
  return OCG_OK;
}

/* The coded-frame area only (frame_width x frame_height of every plane; what th_decode_ycbcr_out exposes):
   three strided copies into the same positions of a host buffer that has the reference's layout.  The
   aprons are device-only state (motion compensation reads them there), so they need not cross PCIe:
   3 133 440 instead of 3 279 360 bytes per 1080p frame. */
OCG_API int ocg_ctx_download_picture(ocg_ctx *c, int buf, uint8_t *host) {
  if (c == nullptr || host == nullptr) return fail(OCG_EFAULT, "NULL argument");
  if (buf < 0 || buf >= c->geom.nrefs) return fail(OCG_EINVAL, "bad buffer index");
  CU(ocg_set_device(c->device));
  const uint8_t *dev = c->frames + (size_t)buf * c->geom.ref_frame_sz;
  for (int pli = 0; pli < 3; pli++) {
    const ocg_plane_geom &p = c->geom.planes[pli];
    const size_t pitch = (size_t)(-(int64_t)p.ystride);
    /* top-left pixel of the plane = lowest address (rows are addressed bottom-up) */
    const int64_t top = c->geom.base_off + p.plane_off + (int64_t)(p.height - 1) * p.ystride;
    CU(cudaMemcpy2DAsync(host + top, pitch, dev + top, pitch, (size_t)p.width, (size_t)p.height, cudaMemcpyDeviceToHost,
                         c->stream));
  }
  return OCG_OK;
}

OCG_API long ocg_picture_bytes(const ocg_geometry *g) {
  long n = 0;
  if (g == nullptr) return 0;
  for (int pli = 0; pli < 3; pli++) n += (long)g->planes[pli].width * g->planes[pli].height;
  return n;
}

OCG_API int ocg_ctx_fill_frame(ocg_ctx *c, int buf, int value) {
  if (c == nullptr) return fail(OCG_EFAULT, "NULL context");
  if (buf < 0 || buf >= c->geom.nrefs) return fail(OCG_EINVAL, "bad buffer index");
  CU(ocg_set_device(c->device));
  CU(cudaMemsetAsync(c->frames + (size_t)buf * c->geom.ref_frame_sz, value, (size_t)c->geom.ref_frame_sz, c->stream));
  return OCG_OK;
}

OCG_API int ocg_host_register(void *p, size_t bytes) {
  if (p == nullptr) return fail(OCG_EFAULT, "NULL argument");
  CU(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  return OCG_OK;
}

OCG_API int ocg_host_unregister(void *p) {
  if (p == nullptr) return fail(OCG_EFAULT, "NULL argument");
  CU(cudaHostUnregister(p));
  return OCG_OK;
}

/* ---- single-frame decode ------------------------------------------------ */
/* Waits until the context's done flag (written by the last node of a flush graph into mapped host
   memory) has reached `seq`: no driver call on the way, so stream threads do not meet in the driver's
   locks.  Wait policy as for ocg_ctx_sync: spin, or -- ocg_set_blocking_sync(1) -- give the core away
   between looks so that a host running more stream threads than cores keeps them busy. */
/* ---- sleeping waits -----------------------------------------------------------------------------------
   ocg_set_blocking_sync(2): a waiting stream thread SLEEPS (semaphore) and one poller thread per process
   watches the completion flags of everybody who sleeps, so a host running several stream threads per core
   spends its cycles on entropy decoding instead of on spinning or yielding waiters.  The poller itself
   sleeps when nobody waits. */
namespace {
struct WaitSlot {
  std::atomic<int> state{0}; /* 0 free, 1 being filled, 2 armed */
  volatile uint32_t *flag = nullptr;
  uint32_t seq = 0;
  sem_t sem;
};
constexpr int kWaitSlots = 512;
WaitSlot g_wait[kWaitSlots];
std::atomic<int> g_wait_armed{0};
/* never destroyed: the poller may be asleep on the condition variable when the process exits, and
   destroying a condition variable that has a waiter blocks (glibc) */
std::mutex &g_wait_mu = *new std::mutex;
std::condition_variable &g_wait_cv = *new std::condition_variable;

void *poller_main(void *) {
  for (;;) {
    if (g_wait_armed.load(std::memory_order_acquire) == 0) {
      std::unique_lock<std::mutex> lk(g_wait_mu);
      g_wait_cv.wait(lk, [] { return g_wait_armed.load(std::memory_order_acquire) > 0; });
    }
    for (int i = 0; i < kWaitSlots; i++) {
      WaitSlot &w = g_wait[i];
      if (w.state.load(std::memory_order_acquire) != 2) continue;
      if ((int32_t)(*w.flag - w.seq) >= 0) {
        /* claim (a waiter that times out at this moment loses the race and takes the wake-up), wake, then
           release the slot: it belongs to one thread, which re-arms it only once it is free again */
        int expect = 2;
        if (!w.state.compare_exchange_strong(expect, 3, std::memory_order_acq_rel)) continue;
        sem_post(&w.sem);
        g_wait_armed.fetch_sub(1, std::memory_order_acq_rel);
        w.state.store(0, std::memory_order_release);
      }
    }
    __builtin_ia32_pause();
  }
  return nullptr;
}

void start_poller() {
  static std::once_flag once;
  std::call_once(once, [] {
    for (int i = 0; i < kWaitSlots; i++) sem_init(&g_wait[i].sem, 0, 0);
    pthread_t th;
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
    pthread_create(&th, &at, poller_main, nullptr);
    pthread_attr_destroy(&at);
  });
}

/* returns false if there is no slot for this thread (the caller polls instead) or the wait timed out */
bool sleep_until(volatile uint32_t *flag, uint32_t seq, int timeout_s) {
  static std::atomic<int> next_slot{0};
  static thread_local int my = -1; /* a thread keeps its slot: nobody else ever waits on its semaphore */
  start_poller();
  if (my < 0) {
    const int i = next_slot.fetch_add(1);
    if (i >= kWaitSlots) { next_slot.store(kWaitSlots); return false; }
    my = i;
  }
  WaitSlot &w = g_wait[my];
  /* the poller releases the slot right after waking us (unless it was descheduled in between) */
  while (w.state.load(std::memory_order_acquire) != 0) sched_yield();
  while (sem_trywait(&w.sem) == 0) {} /* a stale post from a wait that timed out */
  w.flag = flag;
  w.seq = seq;
  w.state.store(2, std::memory_order_release);
  if (g_wait_armed.fetch_add(1, std::memory_order_acq_rel) == 0) {
    std::lock_guard<std::mutex> lk(g_wait_mu);
    g_wait_cv.notify_one();
  }
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  ts.tv_sec += timeout_s;
  if (sem_timedwait(&w.sem, &ts) == 0) return true;
  /* timed out: disarm -- unless the poller has just claimed the slot: then its wake-up is on the way */
  int expect = 2;
  if (w.state.compare_exchange_strong(expect, 0, std::memory_order_acq_rel)) {
    g_wait_armed.fetch_sub(1, std::memory_order_acq_rel);
    return false;
  }
  while (sem_wait(&w.sem) != 0) {}
  return true;
}
} /* namespace */

/* Waits until the context's done flag (written by the last kernel of a flush into mapped host memory) has
   reached `seq`: no driver call on the way, so stream threads do not meet in the driver's locks.  Wait
   policy (ocg_set_blocking_sync): 0 spin, 1 yield the core between looks, 2 sleep until the poller thread
   sees the flag. */
static int wait_done(ocg_ctx *c, uint32_t seq) {
  const int policy = g_blocking_sync.load();
  struct timespec t0;
  long spins = 0;
  bool timed = false;
  while ((int32_t)(*c->h_done - seq) < 0) {
    if (policy == 2 && spins >= 64) {
      if (!sleep_until(c->h_done, seq, 1) && (int32_t)(*c->h_done - seq) < 0) {
        /* a faulted kernel never writes the flag: look at the stream */
        cudaError_t e = cudaStreamQuery(c->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) return fail(OCG_ECUDA, "flush failed on the device", e);
        if (e == cudaSuccess && (int32_t)(*c->h_done - seq) < 0) return fail(OCG_ECUDA, "flush finished without its completion flag");
      }
      continue;
    }
    if (policy == 1) sched_yield();
    else __builtin_ia32_pause();
    if ((++spins & 0xFFF) == 0) {
      /* a faulted kernel never writes the flag: look at the stream now and then */
      struct timespec t1;
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if (!timed) { t0 = t1; timed = true; }
      else if (t1.tv_sec - t0.tv_sec >= 1) {
        cudaError_t e = cudaStreamQuery(c->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) return fail(OCG_ECUDA, "flush failed on the device", e);
        if (e == cudaSuccess && (int32_t)(*c->h_done - seq) < 0) return fail(OCG_ECUDA, "flush finished without its completion flag");
        t0 = t1;
      }
    }
  }
  return OCG_OK;
}

static int acquire_slot(ocg_ctx *c) {
  const int si = (c->cur_slot + 1) % kSlots;
  Slot &s = c->slots[si];
  if (s.busy) {
    cudaError_t e = cudaEventSynchronize(s.consumed);
    if (e != cudaSuccess) return fail(OCG_ECUDA, "cudaEventSynchronize", e);
    s.busy = false;
  }
  if (s.flush_busy) {
    int r = wait_done(c, s.flush_seq);
    if (r < 0) return r;
    s.flush_busy = false;
  }
  c->cur_slot = si;
  return OCG_OK;
}

OCG_API int ocg_dec_staging(ocg_ctx *c, ocg_staging *out) {
  if (c == nullptr || out == nullptr) return fail(OCG_EFAULT, "NULL argument");
  CU(ocg_set_device(c->device));
  int r = acquire_slot(c);
  if (r < 0) return r;
  Slot &s = c->slots[c->cur_slot];
  out->recs = s.recs;
  out->coeff_rows = s.rows;
  c->staged = true;
  return OCG_OK;
}

OCG_API int ocg_dec_submit(ocg_ctx *c, const ocg_dec_frame *f, uint8_t *host_out) {
  if (c == nullptr || f == nullptr) return fail(OCG_EFAULT, "NULL argument");
  int r = check_frame(c->geom, *f);
  if (r < 0) return r;
  CU(ocg_set_device(c->device));
  const bool from_staging = c->staged && f->recs == nullptr && f->coeff_rows == nullptr;
  if (!from_staging) {
    if (!f->recs || (f->ncoeff_rows && !f->coeff_rows)) return fail(OCG_EFAULT, "NULL list pointer (and no staged lists)");
    r = acquire_slot(c);
    if (r < 0) return r;
  }
  Slot &s = c->slots[c->cur_slot];
  c->staged = false;
  cudaStream_t st = c->stream;
  const size_t nf = (size_t)c->geom.nfrags;
  if (!from_staging) memcpy(s.recs, f->recs, nf * sizeof(ocg_frag_rec));
  CU(cudaMemcpyAsync(c->d_recs, s.recs, nf * sizeof(ocg_frag_rec), cudaMemcpyHostToDevice, st));
  if (f->ncoeff_rows) {
    if (!from_staging) memcpy(s.rows, f->coeff_rows, (size_t)f->ncoeff_rows * 16);
    CU(cudaMemcpyAsync(c->d_rows, s.rows, (size_t)f->ncoeff_rows * 16, cudaMemcpyHostToDevice, st));
  }
  fill_job(*s.job, c, *f, c->d_recs, c->d_rows);
  CU(cudaMemcpyAsync(c->d_job, s.job, sizeof(OcgJobDev), cudaMemcpyHostToDevice, st));
  CU(cudaEventRecord(s.consumed, st));
  s.busy = true;
  if (f->dc_residual == 2) {
    if (!c->dc_ahead) return fail(OCG_EINVAL, "dc_residual=2 without a preceding ocg_dec_dc_begin");
    ocg_launch_dc_patch(c->d_recs, c->d_dc_final, c->geom.nfrags, st);
  }
  c->dc_ahead = false;
  launch_stages(c->gdev, c->d_job, 1, f->lf_limit != 0, c->d_tmaps != nullptr && g_use_tma.load(), st, f->dc_residual == 1);
  CU(cudaGetLastError());
  if (host_out != nullptr) {
    CU(cudaMemcpyAsync(host_out, c->frames + (size_t)f->ref_idx[OCG_FRAME_SELF] * c->geom.ref_frame_sz,
                       (size_t)c->geom.ref_frame_sz, cudaMemcpyDeviceToHost, st));
  }
  return OCG_OK;
}

static OcgExpandBufs expand_bufs(const ocg_ctx *c) {
  OcgExpandBufs B;
  B.order = c->x.d_order;
  B.buf_off = c->x.d_buf_off;
  B.dequant = c->x.d_dequant;
  B.words = (const uint32_t *)(c->x.d_words_raw + ((uintptr_t)c->x.h_words & 15));
  B.mvs = (const int16_t *)(c->x.d_mvs_raw + ((uintptr_t)c->x.h_mvs & 15));
  B.tokens = c->x.d_tokens_raw + ((uintptr_t)c->x.h_tokens & 15);
  B.tok = c->x.d_tok;
  B.cov = c->x.d_cov;
  B.ntok = c->x.d_ntok;
  B.coef = c->x.d_coef;
  B.nextz = c->x.d_nextz;
  B.lastz = c->x.d_lastz;
  B.rmask = c->x.d_rmask;
  return B;
}

/* The kernel sequence of one frame flush on the context's stream (directly, or under capture).
   tokens: the frame comes as the decoder's token lists (ocg_dec_flush_tokens) instead of records. */
static void launch_flush_sequence(ocg_ctx *c, int si, int out_mode, int dc, int lf, int tokens, bool tma) {
  Slot &s = c->slots[si];
  cudaStream_t st = c->stream;
  if (tokens) {
    const OcgExpandBufs B = expand_bufs(c);
    ocg_launch_stage_tokens(s.m_job, c->d_job, s.m_xjob, c->x.d_xjob, c->x.m_words, (void *)B.words, c->geom.nfrags * 4,
                            c->x.m_mvs, (void *)B.mvs, c->geom.nfrags * 2, c->x.m_tokens, (void *)B.tokens, st);
    if (dc) ocg_launch_dc_unpredict_words(c->gdev, B.words, c->d_dc_final, c->d_dc_tmp, st);
    ocg_launch_expand(c->gdev, c->x.d_xjob, B, dc ? c->d_dc_final : nullptr, c->d_job, c->d_recs, st);
  } else {
    ocg_launch_stage_in(s.m_job, c->d_job, s.m_recs, c->d_recs, c->geom.nfrags, s.m_rows, c->d_rows, st);
    if (dc) ocg_launch_dc_unpredict(c->gdev, c->d_job, 1, st);
  }
  ocg_launch_recon(c->gdev, c->d_job, 1, st);
  if (lf) ocg_launch_loop_filter(c->gdev, c->d_job, 1, tma, st);
  ocg_launch_borders(c->gdev, c->d_job, 1, st);
  ocg_launch_copy_out(c->geom, out_mode, c->d_job, c->d_out_counter, c->d_done, st);
}

/* Captures and instantiates the flush of one staging slot: kernels only (see ocg_dec_flush). */
static int build_flush_graph(ocg_ctx *c, int si, int out_mode, int dc, int lf, int tokens, bool tma) {
  cudaStream_t st = c->stream;
  cudaGraph_t graph = nullptr;
  const long k0 = g_launches.load();
  /* thread-local capture: other threads' contexts keep working meanwhile */
  CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  launch_flush_sequence(c, si, out_mode, dc, lf, tokens, tma);
  cudaError_t e = cudaStreamEndCapture(st, &graph);
  if (e != cudaSuccess || graph == nullptr) { cudaGetLastError(); return fail(OCG_ECUDA, "flush graph capture failed", e); }
  (void)k0;
  size_t nnodes = 0;
  cudaGraphGetNodes(graph, nullptr, &nnodes); /* all of them kernels */
  const int kernels = (int)nnodes;
  g_launches.fetch_sub(kernels); /* the capture counted them once; they are counted per replay */
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(OCG_ECUDA, "cudaGraphInstantiate", e);
  c->graphs.push_back(FlushGraph{si, out_mode, dc, lf, tokens, exec, kernels});
  return OCG_OK;
}

static std::atomic<long> g_flush_prep_ns{0}, g_flush_launch_ns{0}, g_flush_n{0}, g_flush_build_ns{0}, g_flush_builds{0};
static inline long now_ns() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (long)ts.tv_sec * 1000000000L + ts.tv_nsec;
}

/* One frame = one driver call.  The whole flush -- records and job header H2D, [DC un-prediction],
   recon pass A/B, loop filter, borders, copy-back of the picture (or the padded buffer) into host_out,
   completion flag -- is a CUDA graph, instantiated once per (staging slot, SELF buffer, destination,
   variant) and replayed; the coefficient rows are read in place from the mapped staging memory.  With
   one stream thread per decoder the eleven driver calls of ocg_dec_submit + copy-back + sync serialise
   on the driver's locks (measured: 0.25 ms of host time per frame at 16 threads); a graph launch plus a
   flag in host memory needs one. */
/* Shared tail of ocg_dec_flush / ocg_dec_flush_tokens: the slot's job header is filled in except for the
   destination and the sequence number. */
static const int kOutDeferred = 3; /* internal out_mode: no copy-out kernel in the sequence */
static int flush_core(ocg_ctx *c, int si, uint8_t *host_out, int out_mode, int dc, int lf, int tokens, long t_in) {
  Slot &s = c->slots[si];
  cudaStream_t st = c->stream;
  int r;
  {
    /* 1: every context; n > 1: one context in n (by creation order) -- the rest keep the copy-out kernel */
    const int dma = g_out_dma.load(std::memory_order_relaxed);
    if (out_mode == OCG_OUT_PICTURE && dma > 0 && (dma == 1 || c->serial % dma == 0)) out_mode = kOutDeferred;
  }
  const bool tma = c->d_tmaps != nullptr && g_use_tma.load();
  /* the destination's device address (one driver call per distinct buffer, then remembered) */
  uint8_t *d_out_mapped = nullptr;
  if (out_mode != OCG_OUT_NONE) {
    for (const MappedOut &m : c->outs) if (m.host == host_out) d_out_mapped = m.dev;
    if (d_out_mapped == nullptr) {
      if (((uintptr_t)host_out & 15) != 0) return fail(OCG_EINVAL, "output buffer must be 16-byte aligned");
      if (cudaHostGetDevicePointer((void **)&d_out_mapped, host_out, 0) != cudaSuccess) {
        cudaGetLastError();
        return fail(OCG_EINVAL, "output buffer is not page-locked (ocg_host_register)");
      }
      if (c->outs.size() >= 16) c->outs.clear();
      c->outs.push_back(MappedOut{host_out, d_out_mapped});
    }
  }
  s.job->host_out = d_out_mapped;
  /* A stream's first flushes are launched kernel by kernel; from flush `graph_after` on the sequence is
     replayed as a CUDA graph (one driver call per frame instead of six or more).  Instantiating a graph
     costs ~0.3 ms, and tens of ms when many threads do it at once, so short-lived contexts never pay for it
     and the builds of a process are serialised here rather than inside the driver. */
  FlushGraph *fg = nullptr;
  long t_build = 0;
  if (c->graph_after >= 0 && c->nflush >= c->graph_after) {
    for (int pass = 0; pass < 2 && fg == nullptr; pass++) {
      for (FlushGraph &g : c->graphs)
        if (g.slot == si && g.out_mode == out_mode && g.dc == dc && g.lf == lf && g.tokens == tokens) fg = &g;
      if (fg == nullptr) {
        const long tb = now_ns();
        {
          static std::mutex build_lock;
          std::lock_guard<std::mutex> lk(build_lock);
          for (int k = 0; k < kSlots; k++) {
            r = build_flush_graph(c, k, out_mode, dc, lf, tokens, tma);
            if (r < 0) return r;
          }
        }
        t_build = now_ns() - tb;
        g_flush_build_ns.fetch_add(t_build, std::memory_order_relaxed);
        g_flush_builds.fetch_add(kSlots, std::memory_order_relaxed);
      }
    }
    if (fg == nullptr) return fail(OCG_ECUDA, "flush graph missing");
  }
  /* the stage-in kernel reads the host memory when it runs: everything it reads is final now */
  c->flush_seq++;
  s.job->seq = c->flush_seq;
  s.flush_seq = c->flush_seq;
  s.flush_busy = true;
  c->nflush++;
  const long t_launch = now_ns();
  int nk = 0;
  if (fg != nullptr) {
    CU(cudaGraphLaunch(fg->exec, st));
    nk = fg->kernels;
  } else {
    launch_flush_sequence(c, si, out_mode, dc, lf, tokens, tma); /* the launch helpers count their kernels */
    CU(cudaGetLastError());
  }
  if (out_mode == kOutDeferred) {
    /* the picture by the copy engines (three 2-D copies, one per plane), then the completion flag */
    const ocg_geometry &g = c->geom;
    const uint8_t *dev_frame = s.job->base[OCG_FRAME_SELF] - g.base_off;
    for (int pli = 0; pli < 3; pli++) {
      const ocg_plane_geom &p = g.planes[pli];
      const int64_t off = g.base_off + p.plane_off + (int64_t)(p.height - 1) * p.ystride; /* top-left pixel */
      CU(cudaMemcpy2DAsync(host_out + off, (size_t)-p.ystride, dev_frame + off, (size_t)-p.ystride, (size_t)p.width, (size_t)p.height,
                           cudaMemcpyDeviceToHost, st));
    }
    ocg_launch_copy_out(c->geom, OCG_OUT_NONE, c->d_job, c->d_out_counter, c->d_done, st);
    nk++;
  }
  const long t_out = now_ns();
  g_flush_prep_ns.fetch_add(t_launch - t_in - t_build, std::memory_order_relaxed);
  g_flush_launch_ns.fetch_add(t_out - t_launch, std::memory_order_relaxed);
  g_flush_n.fetch_add(1, std::memory_order_relaxed);
  g_launches.fetch_add(nk, std::memory_order_relaxed);
  return OCG_OK;
}

OCG_API int ocg_dec_flush(ocg_ctx *c, const ocg_dec_frame *f, uint8_t *host_out, int out_mode) {
  if (c == nullptr || f == nullptr) return fail(OCG_EFAULT, "NULL argument");
  int r = check_frame(c->geom, *f);
  if (r < 0) return r;
  if (f->dc_residual == 2) return fail(OCG_EINVAL, "ocg_dec_flush takes dc_residual 0 or 1");
  if (out_mode != OCG_OUT_NONE && host_out == nullptr) return fail(OCG_EFAULT, "NULL output buffer");
  const long t_in = now_ns();
  CU(ocg_set_device(c->device));
  const bool from_staging = c->staged && f->recs == nullptr && f->coeff_rows == nullptr;
  if (!from_staging) {
    if (!f->recs || (f->ncoeff_rows && !f->coeff_rows)) return fail(OCG_EFAULT, "NULL list pointer (and no staged lists)");
    r = acquire_slot(c);
    if (r < 0) return r;
  }
  const int si = c->cur_slot;
  Slot &s = c->slots[si];
  c->staged = false;
  const size_t nf = (size_t)c->geom.nfrags;
  if (!from_staging) {
    memcpy(s.recs, f->recs, nf * sizeof(ocg_frag_rec));
    if (f->ncoeff_rows) memcpy(s.rows, f->coeff_rows, (size_t)f->ncoeff_rows * 16);
  }
  fill_job(*s.job, c, *f, c->d_recs, c->d_rows);
  s.job->ncoeff_rows = f->ncoeff_rows;
  return flush_core(c, si, host_out, out_mode, f->dc_residual == 1, f->lf_limit != 0, 0, t_in);
}

/* ---- device-side token expansion ------------------------------------------------------------------ */
OCG_API int ocg_dec_expand_setup(ocg_ctx *c, const int32_t *coded_order, const uint16_t *dequant_tables,
                                 const uint32_t *frag_words, const int16_t *frag_mvs, const uint8_t *dct_tokens,
                                 size_t token_capacity) {
  if (c == nullptr || coded_order == nullptr || dequant_tables == nullptr || frag_words == nullptr || frag_mvs == nullptr ||
      dct_tokens == nullptr)
    return fail(OCG_EFAULT, "NULL argument");
  if (c->x.ready) return fail(OCG_EINVAL, "token expansion is already set up for this context");
  if (token_capacity == 0 || token_capacity > ((size_t)1 << 30)) return fail(OCG_EINVAL, "bad token capacity");
  CU(ocg_set_device(c->device));
  const size_t nf = (size_t)c->geom.nfrags;
  std::vector<uint8_t> seen(nf, 0);
  for (int pli = 0; pli < 3; pli++) {
    const ocg_plane_geom &p = c->geom.planes[pli];
    for (int i = 0; i < p.nfrags; i++) {
      const int32_t f = coded_order[p.froffset + i];
      if (f < p.froffset || f >= p.froffset + p.nfrags || seen[(size_t)f]) return fail(OCG_EINVAL, "coded order is not a permutation of each plane's fragments");
      seen[(size_t)f] = 1;
    }
  }
  ocg_ctx::Expand &x = c->x;
  cudaStream_t st = c->stream;
  ocg_expand_init_tables(st);
  std::vector<int32_t> offs(nf);
  ocg_geometry_frag_buf_offs(&c->geom, offs.data());
  const size_t tokb = (token_capacity + 63) & ~(size_t)15;
  CU(cudaMalloc(&x.d_order, nf * 4));
  CU(cudaMalloc(&x.d_buf_off, nf * 4));
  CU(cudaMalloc(&x.d_ntok, 192 * 4));
  CU(cudaMalloc(&x.d_dequant, 64 * 3 * 2 * 64 * 2));
  CU(cudaMalloc(&x.d_words_raw, nf * 4 + 64));
  CU(cudaMalloc(&x.d_mvs_raw, nf * 2 + 64));
  CU(cudaMalloc(&x.d_tokens_raw, tokb + 64));
  CU(cudaMalloc(&x.d_tok, tokb * 4));
  CU(cudaMalloc(&x.d_cov, tokb * 4));
  CU(cudaMalloc(&x.d_coef, nf * 128));
  {
    /* ocg_tok_expand_kernel's global fall-back: per plane, whole words per thread (ocg_dec_expand.cu) */
    size_t nmax = 0;
    for (int pli = 0; pli < 3; pli++) nmax = nmax > (size_t)c->geom.planes[pli].nfrags ? nmax : (size_t)c->geom.planes[pli].nfrags;
    const size_t need = ((((nmax + 1023) / 1024) + 3) & ~(size_t)3) * 1024;
    CU(cudaMalloc(&x.d_nextz, 3 * need));
  }
  CU(cudaMalloc(&x.d_lastz, nf));
  CU(cudaMalloc(&x.d_rmask, nf));
  CU(cudaMalloc(&x.d_xjob, sizeof(OcgExpandDev)));
  CU(cudaMemcpyAsync(x.d_order, coded_order, nf * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(x.d_buf_off, offs.data(), nf * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(x.d_dequant, dequant_tables, 64 * 3 * 2 * 64 * 2, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(x.d_coef, 0, nf * 128, st));
  CU(cudaStreamSynchronize(st));
  x.token_cap = token_capacity;
  x.h_words = frag_words;
  x.h_mvs = frag_mvs;
  x.h_tokens = dct_tokens;
  if (cudaHostGetDevicePointer((void **)&x.m_words, (void *)frag_words, 0) != cudaSuccess ||
      cudaHostGetDevicePointer((void **)&x.m_mvs, (void *)frag_mvs, 0) != cudaSuccess ||
      cudaHostGetDevicePointer((void **)&x.m_tokens, (void *)dct_tokens, 0) != cudaSuccess) {
    cudaGetLastError();
    return fail(OCG_EINVAL, "the fragment / vector / token arrays must be page-locked (ocg_host_register)");
  }
  x.ready = true;
  return OCG_OK;
}

OCG_API int ocg_dec_flush_tokens(ocg_ctx *c, const ocg_dec_tokens *t, uint8_t *host_out, int out_mode) {
  if (c == nullptr || t == nullptr) return fail(OCG_EFAULT, "NULL argument");
  if (!c->x.ready) return fail(OCG_EINVAL, "ocg_dec_expand_setup has not been called");
  if (out_mode != OCG_OUT_NONE && host_out == nullptr) return fail(OCG_EFAULT, "NULL output buffer");
  if (t->ntoken_bytes < 0 || (size_t)t->ntoken_bytes > c->x.token_cap) return fail(OCG_EINVAL, "token count out of range");
  if (t->nqis < 1 || t->nqis > 3) return fail(OCG_EINVAL, "nqis must be 1..3");
  for (int i = 0; i < t->nqis; i++) if (t->qis[i] < 0 || t->qis[i] > 63) return fail(OCG_EINVAL, "qi out of range");
  if (t->dc_residual != 0 && t->dc_residual != 1) return fail(OCG_EINVAL, "dc_residual must be 0 or 1");
  ocg_dec_frame f;
  memset(&f, 0, sizeof(f));
  for (int i = 0; i < 3; i++) f.ref_idx[i] = t->ref_idx[i];
  f.lf_limit = t->lf_limit;
  f.intra_frame = t->intra_frame;
  f.dc_residual = t->dc_residual;
  memcpy(f.dc_quant, t->dc_quant, sizeof(f.dc_quant));
  int r = check_frame(c->geom, f);
  if (r < 0) return r;
  if (!t->intra_frame && (t->ref_idx[OCG_FRAME_PREV] < 0 || t->ref_idx[OCG_FRAME_GOLD] < 0)) return fail(OCG_EINVAL, "inter frame without reference buffers");
  for (int p = 0; p < 3; p++) {
    int prev = -1;
    for (int z = 0; z < 64; z++) {
      const int v = t->ti0[p][z];
      if (v < 0 || v > t->ntoken_bytes || (z > 0 && v < prev)) return fail(OCG_EINVAL, "token list offsets out of order");
      prev = v;
    }
  }
  const long t_in = now_ns();
  CU(ocg_set_device(c->device));
  r = acquire_slot(c);
  if (r < 0) return r;
  const int si = c->cur_slot;
  Slot &s = c->slots[si];
  c->staged = false;
  f.dc_residual = 0; /* the records the reconstruction reads are final either way */
  fill_job(*s.job, c, f, c->d_recs, c->x.d_coef);
  s.job->dense_rows = 1;
  OcgExpandDev &X = *s.xjob;
  for (int p = 0; p < 3; p++)
    for (int z = 0; z < 64; z++) {
      X.ti0[p][z] = t->ti0[p][z];
      X.eob_runs[p][z] = t->eob_runs[p][z] < 0 ? 0 : t->eob_runs[p][z];
    }
  X.ntoken_bytes = t->ntoken_bytes;
  X.nqis = t->nqis;
  for (int i = 0; i < 3; i++) X.qis[i] = i < t->nqis ? t->qis[i] : t->qis[0];
  return flush_core(c, si, host_out, out_mode, t->dc_residual == 1, t->lf_limit != 0, 1, t_in);
}

OCG_API void ocg_ctx_set_flush_graph(ocg_ctx *c, int after) {
  if (c != nullptr) c->graph_after = after;
}

OCG_API void ocg_flush_profile_builds(double *build_s, long *nbuilds) {
  if (build_s) *build_s = 1e-9 * (double)g_flush_build_ns.load();
  if (nbuilds) *nbuilds = g_flush_builds.load();
}

OCG_API void ocg_flush_profile(double *prepare_s, double *launch_s, long *n, int reset) {
  if (prepare_s) *prepare_s = 1e-9 * (double)g_flush_prep_ns.load();
  if (launch_s) *launch_s = 1e-9 * (double)g_flush_launch_ns.load();
  if (n) *n = g_flush_n.load();
  if (reset) { g_flush_prep_ns.store(0); g_flush_launch_ns.store(0); g_flush_n.store(0); }
}

/* Test hook (no device involved): the sleeping wait on a caller-owned flag word; 1 = woken with the flag
   at or past seq, 0 = timed out / no slot. */
OCG_API int ocg_test_sleep_until(volatile uint32_t *flag, uint32_t seq, int timeout_s) {
  if (flag == nullptr) return 0;
  while ((int32_t)(*flag - seq) < 0) {
    if (!sleep_until(flag, seq, timeout_s)) return (int32_t)(*flag - seq) >= 0;
  }
  return 1;
}

OCG_API int ocg_dec_wait(ocg_ctx *c) {
  if (c == nullptr) return fail(OCG_EFAULT, "NULL context");
  int r = wait_done(c, c->flush_seq);
  if (r < 0) return r;
  for (Slot &s : c->slots) s.flush_busy = false;
  return OCG_OK;
}

OCG_API int ocg_dec_dc_begin(ocg_ctx *c, const uint32_t *frag_words) {
  if (c == nullptr || frag_words == nullptr) return fail(OCG_EFAULT, "NULL argument");
  if (!ocg_dc_unpredict_supported(&c->geom)) return fail(OCG_EIMPL, "DC un-prediction on the device does not support this frame size");
  CU(ocg_set_device(c->device));
  CU(cudaMemcpyAsync(c->d_dc_words, frag_words, (size_t)c->geom.nfrags * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  if (ocg_launch_dc_unpredict_words(c->gdev, c->d_dc_words, c->d_dc_final, c->d_dc_tmp, c->stream) < 0)
    return fail(OCG_EIMPL, "DC un-prediction launch failed");
  CU(cudaGetLastError());
  c->dc_ahead = true;
  return OCG_OK;
}

/* ---- intra-frame encoder pre-pass ---------------------------------------- */
static constexpr size_t kQTabU16 = 3 * 2 * 3 * 64;       /* dequant entries */
static constexpr size_t kQTabI16 = 3 * 2 * 3 * 64 * 2;   /* enquant {m,l} entries */

static int enc_pre_init(ocg_ctx *c) {
  EncPre *e = new (std::nothrow) EncPre();
  if (e == nullptr) return fail(OCG_ENOMEM, "out of memory");
  c->enc = e; /* owned by the context from here on (freed in ocg_ctx_destroy) */
  const size_t nf = (size_t)c->geom.nfrags;
  size_t o = 0;
  e->off_satd = o; o += nf * 4;
  e->off_dc = o; o += nf * 4;
  e->off_nz = o; o += 3 * nf * 4;
  o = (o + 255) & ~(size_t)255;
  e->off_dct = o; o += nf * 128;
  e->off_qdct = o; o += 3 * nf * 128;
  e->out_sz = o;
  CU(cudaMalloc(&e->d_frags, 3 * nf * sizeof(ocg_enc_frag)));
  CU(cudaMalloc(&e->d_dequant, kQTabU16 * 2));
  CU(cudaMalloc(&e->d_enquant, kQTabI16 * 2));
  CU(cudaMalloc(&e->d_out, e->out_sz));
  CU(cudaHostAlloc(&e->h_out, e->out_sz, cudaHostAllocDefault));
  CU(cudaHostAlloc(&e->h_tabs, kQTabU16 * 2 + kQTabI16 * 2, cudaHostAllocDefault));
  /* the block lists never change: every fragment, intra (no predictor), one copy per qii */
  std::vector<ocg_enc_frag> fl(3 * nf);
  for (int qii = 0; qii < 3; qii++)
    for (size_t i = 0; i < nf; i++) {
      ocg_enc_frag &f = fl[(size_t)qii * nf + i];
      f.src_off = c->tmpl[i].buf_off;
      f.ref_off0 = INT32_MIN;
      f.ref_off1 = INT32_MIN;
      f.aux = (c->tmpl[i].pli_qti & 3) | (qii << 3);
    }
  CU(cudaMemcpyAsync(e->d_frags, fl.data(), fl.size() * sizeof(ocg_enc_frag), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return OCG_OK;
}

OCG_API int ocg_enc_intra_reserve(ocg_ctx *c) {
  if (c == nullptr) return fail(OCG_EFAULT, "NULL context");
  CU(ocg_set_device(c->device));
  return c->enc != nullptr ? OCG_OK : enc_pre_init(c);
}

OCG_API int ocg_enc_intra_prepass(ocg_ctx *c, int io_buf, const uint8_t *host_frame, const uint16_t *dequant,
                                  const int16_t *enquant, int nqis, ocg_enc_intra_tables *out) {
  if (c == nullptr || host_frame == nullptr || dequant == nullptr || enquant == nullptr || out == nullptr)
    return fail(OCG_EFAULT, "NULL argument");
  if (io_buf < 0 || io_buf >= c->geom.nrefs) return fail(OCG_EINVAL, "bad buffer index");
  if (nqis < 1 || nqis > 3) return fail(OCG_EINVAL, "nqis must be 1..3");
  CU(ocg_set_device(c->device));
  if (c->enc == nullptr) {
    int r = enc_pre_init(c);
    if (r < 0) return r;
  }
  EncPre *e = c->enc;
  cudaStream_t st = c->stream;
  const size_t nf = (size_t)c->geom.nfrags;
  const int nluma = c->geom.planes[0].nfrags;
  const int nchroma = c->geom.nfrags - nluma;
  uint8_t *frame = c->frames + (size_t)io_buf * c->geom.ref_frame_sz;
  const uint8_t *base = frame + c->geom.base_off;
  CU(cudaMemcpyAsync(frame, host_frame, (size_t)c->geom.ref_frame_sz, cudaMemcpyHostToDevice, st));
  memcpy(e->h_tabs, dequant, kQTabU16 * 2);
  memcpy(e->h_tabs + kQTabU16 * 2, enquant, kQTabI16 * 2);
  CU(cudaMemcpyAsync(e->d_dequant, e->h_tabs, kQTabU16 * 2, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(e->d_enquant, e->h_tabs + kQTabU16 * 2, kQTabI16 * 2, cudaMemcpyHostToDevice, st));
  uint32_t *d_satd = (uint32_t *)(e->d_out + e->off_satd);
  int32_t *d_dc = (int32_t *)(e->d_out + e->off_dc);
  int32_t *d_nz = (int32_t *)(e->d_out + e->off_nz);
  int16_t *d_dct = (int16_t *)(e->d_out + e->off_dct);
  int16_t *d_qdct = (int16_t *)(e->d_out + e->off_qdct);
  /* luma and chroma differ in stride: one launch each */
  const int ys[2] = {c->geom.planes[0].ystride, c->geom.planes[1].ystride};
  const int first[2] = {0, nluma}, count[2] = {nluma, nchroma};
  int r;
  for (int k = 0; k < 2; k++) {
    r = ocg_enc_metrics_batch(OCG_MET_INTRA_SATD, base, nullptr, ys[k], e->d_frags + first[k], count[k],
                              d_satd + first[k], d_dc + first[k], st);
    if (r < 0) return fail(r, "intra SATD launch failed");
    for (int qii = 0; qii < nqis; qii++) {
      const size_t at = (size_t)qii * nf + (size_t)first[k];
      /* the transform does not depend on qii: qii > 0 rewrites the same dct rows */
      r = ocg_enc_fdct_quant_batch(base, nullptr, ys[k], e->d_frags + at, count[k], e->d_dequant, e->d_enquant,
                                   d_dct + (size_t)first[k] * 64, d_qdct + at * 64, d_nz + at, st);
      if (r < 0) return fail(r, "fDCT/quantiser launch failed");
    }
  }
  /* one D2H for the small tables + dct, one for the nqis quantised planes */
  CU(cudaMemcpyAsync(e->h_out, e->d_out, e->off_dct + nf * 128, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(e->h_out + e->off_qdct, e->d_out + e->off_qdct, (size_t)nqis * nf * 128, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  out->satd = (const uint32_t *)(e->h_out + e->off_satd);
  out->satd_dc = (const int32_t *)(e->h_out + e->off_dc);
  out->dct = (const int16_t *)(e->h_out + e->off_dct);
  out->qdct = (const int16_t *)(e->h_out + e->off_qdct);
  out->nonzero = (const int32_t *)(e->h_out + e->off_nz);
  return OCG_OK;
}

/* ---- device-resident packs ---------------------------------------------- */
OCG_API void ocg_pack_destroy(ocg_pack *p) {
  if (p == nullptr) return;
  ocg_set_device(p->device);
  cudaDeviceSynchronize(); /* a pack is read by launches on any stream; its block goes back to the cache */
  cudaFree(p->blob);
  delete p;
}

OCG_API int ocg_pack_nframes(const ocg_pack *p) { return p ? p->nframes : 0; }

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

OCG_API int ocg_pack_create(ocg_pack **out, const ocg_dec_frame *frames, int nframes, int nfrags, int device) {
  if (out == nullptr || frames == nullptr) return fail(OCG_EFAULT, "NULL argument");
  *out = nullptr;
  if (nframes <= 0 || nfrags <= 0) return fail(OCG_EINVAL, "empty pack");
  CU(ocg_set_device(device));
  size_t total = 0;
  for (int i = 0; i < nframes; i++)
    total += align256((size_t)nfrags * sizeof(ocg_frag_rec)) + align256((size_t)frames[i].ncoeff_rows * 16);
  ocg_pack *p = new (std::nothrow) ocg_pack();
  if (p == nullptr) return fail(OCG_ENOMEM, "out of memory");
  p->device = device;
  p->nframes = nframes;
  p->nfrags = nfrags;
  p->blob_sz = total;
  cudaError_t e = cudaMalloc(&p->blob, total ? total : 256);
  if (e != cudaSuccess) { delete p; return fail(OCG_ECUDA, "cudaMalloc(pack)", e); }
  size_t at = 0;
  p->frames.resize((size_t)nframes);
  for (int i = 0; i < nframes; i++) {
    const ocg_dec_frame &f = frames[i];
    ocg_dec_frame d = f;
    struct Part { const void *src; size_t n; const void **dst; } parts[2] = {
        {f.recs, (size_t)nfrags * sizeof(ocg_frag_rec), (const void **)&d.recs},
        {f.coeff_rows, (size_t)f.ncoeff_rows * 16, (const void **)&d.coeff_rows}};
    for (Part &pt : parts) {
      *pt.dst = p->blob + at;
      if (pt.n) {
        if (pt.src == nullptr) { ocg_pack_destroy(p); return fail(OCG_EFAULT, "NULL list pointer in pack frame"); }
        e = cudaMemcpy(p->blob + at, pt.src, pt.n, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { ocg_pack_destroy(p); return fail(OCG_ECUDA, "cudaMemcpy(pack)", e); }
      }
      at += align256(pt.n);
    }
    p->frames[(size_t)i] = d;
  }
  *out = p;
  return OCG_OK;
}

struct BatchScratch {
  OcgJobDev *h = nullptr; /* pinned */
  OcgJobDev *d = nullptr;
  int cap = 0;
  int device = -1;
  cudaEvent_t done = nullptr;
  bool pending = false;
};
static thread_local BatchScratch g_bs[8]; /* job tables in flight: the host may run this many batches ahead */
static thread_local int g_bs_i = 0;

OCG_API int ocg_dec_run_batch(ocg_ctx *const *ctxs, ocg_pack *const *packs, const int32_t *frame_idx, int n,
                              void *stream) {
  if (ctxs == nullptr || packs == nullptr || frame_idx == nullptr) return fail(OCG_EFAULT, "NULL argument");
  if (n <= 0) return fail(OCG_EINVAL, "empty batch");
  ocg_ctx *c0 = ctxs[0];
  if (c0 == nullptr) return fail(OCG_EFAULT, "NULL context");
  CU(ocg_set_device(c0->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c0->stream;
  g_bs_i = (g_bs_i + 1) & 7;
  BatchScratch &bs = g_bs[g_bs_i];
  if (bs.pending) { CU(cudaEventSynchronize(bs.done)); bs.pending = false; }
  if (bs.cap < n || bs.device != c0->device) {
    if (bs.h) cudaFreeHost(bs.h);
    if (bs.d) cudaFree(bs.d);
    bs.h = nullptr; bs.d = nullptr; bs.cap = 0;
    CU(cudaHostAlloc(&bs.h, sizeof(OcgJobDev) * (size_t)n, cudaHostAllocDefault));
    CU(cudaMalloc(&bs.d, sizeof(OcgJobDev) * (size_t)n));
    if (bs.done == nullptr) CU(cudaEventCreateWithFlags(&bs.done, cudaEventDisableTiming));
    bs.cap = n;
    bs.device = c0->device;
  }
  bool any_lf = false, all_tma = g_use_tma.load() != 0;
  for (int i = 0; i < n; i++) {
    ocg_ctx *c = ctxs[i];
    ocg_pack *p = packs[i];
    if (c == nullptr || p == nullptr) return fail(OCG_EFAULT, "NULL job");
    if (c->device != c0->device || p->device != c0->device) return fail(OCG_EINVAL, "batch spans devices");
    if (memcmp(&c->geom, &c0->geom, sizeof(ocg_geometry)) != 0) return fail(OCG_EINVAL, "batch mixes geometries");
    if (p->nfrags != c->geom.nfrags) return fail(OCG_EINVAL, "pack was built for another geometry");
    if (frame_idx[i] < 0 || frame_idx[i] >= p->nframes) return fail(OCG_EINVAL, "frame index out of range");
    const ocg_dec_frame &f = p->frames[(size_t)frame_idx[i]];
    int r = check_frame(c->geom, f);
    if (r < 0) return r;
    /* a resident frame is replayed: its records must not be rewritten in place */
    if (f.dc_residual != 0) return fail(OCG_EINVAL, "resident packs hold final DC values (dc_residual frames go through ocg_dec_submit)");
    fill_job(bs.h[i], c, f, f.recs, f.coeff_rows);
    any_lf |= f.lf_limit != 0;
    all_tma = all_tma && c->d_tmaps != nullptr;
  }
  CU(cudaMemcpyAsync(bs.d, bs.h, sizeof(OcgJobDev) * (size_t)n, cudaMemcpyHostToDevice, st));
  launch_stages(c0->gdev, bs.d, n, any_lf, all_tma, st);
  CU(cudaGetLastError());
  /* the job table (host and device copy) is free again once the kernels ran */
  CU(cudaEventRecord(bs.done, st));
  bs.pending = true;
  return OCG_OK;
}

} /* extern "C" */
