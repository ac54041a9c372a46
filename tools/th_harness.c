/* Test/bench tooling: ctypes-friendly driver around the
 * reference's *public* API (include/theora/theoraenc.h:456-537,
 * theoradec.h:234-322).  It is compiled into oracle/_ref/libth_{c,asm}.so next
 * to the unmodified reference objects, and into the integrated build
 * (reference host code + the B200 back-end) so the same calls exercise both.
 *
 * Nothing here restates codec arithmetic: it generates the synthetic video of
 * SURVEY.md section 8(d), feeds the encoder / decoder, keeps packets in RAM,
 * hashes decoded planes (FNV-1a 64) and runs timing loops.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <pthread.h>
#include "theora/theoraenc.h"
#include "theora/theoradec.h"

#define REFH_API __attribute__((visibility("default")))

typedef struct refh_stream {
  int             npackets;
  int             cap;
  long           *sizes;
  unsigned char **data;
  int             width;   /* picture size */
  int             height;
} refh_stream;

static double refh_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int refh_ilog(unsigned v) {
  int r = 0;
  while (v) { r++; v >>= 1; }
  return r;
}

static void refh_stream_push(refh_stream *s, const unsigned char *p, long n) {
  if (s->npackets == s->cap) {
    s->cap = s->cap ? 2 * s->cap : 64;
    s->sizes = (long *)realloc(s->sizes, sizeof(long) * (size_t)s->cap);
    s->data = (unsigned char **)realloc(s->data, sizeof(unsigned char *) * (size_t)s->cap);
  }
  s->sizes[s->npackets] = n;
  s->data[s->npackets] = (unsigned char *)malloc((size_t)(n > 0 ? n : 1));
  if (n > 0) memcpy(s->data[s->npackets], p, (size_t)n);
  s->npackets++;
}

REFH_API void refh_stream_free(refh_stream *s) {
  int i;
  if (s == NULL) return;
  for (i = 0; i < s->npackets; i++) free(s->data[i]);
  free(s->data);
  free(s->sizes);
  free(s);
}

REFH_API int refh_stream_npackets(const refh_stream *s) { return s->npackets; }
REFH_API long refh_stream_packet_size(const refh_stream *s, int i) { return s->sizes[i]; }
REFH_API const unsigned char *refh_stream_packet_data(const refh_stream *s, int i) { return s->data[i]; }

/* Serialised form: u32 magic, u32 npackets, u32 w, u32 h, u32 sizes[n], bytes. */
REFH_API long refh_stream_blob_size(const refh_stream *s) {
  long n = 16 + 4L * s->npackets;
  int i;
  for (i = 0; i < s->npackets; i++) n += s->sizes[i];
  return n;
}

REFH_API long refh_stream_to_blob(const refh_stream *s, unsigned char *out, long cap) {
  long need = refh_stream_blob_size(s);
  uint32_t *hdr = (uint32_t *)out;
  unsigned char *p;
  int i;
  if (cap < need) return -1;
  hdr[0] = 0x4F434753u;
  hdr[1] = (uint32_t)s->npackets;
  hdr[2] = (uint32_t)s->width;
  hdr[3] = (uint32_t)s->height;
  for (i = 0; i < s->npackets; i++) hdr[4 + i] = (uint32_t)s->sizes[i];
  p = out + 16 + 4L * s->npackets;
  for (i = 0; i < s->npackets; i++) {
    memcpy(p, s->data[i], (size_t)s->sizes[i]);
    p += s->sizes[i];
  }
  return need;
}

REFH_API refh_stream *refh_stream_from_blob(const unsigned char *blob, long n) {
  const uint32_t *hdr = (const uint32_t *)blob;
  refh_stream *s;
  const unsigned char *p;
  uint32_t i, np;
  if (n < 16 || hdr[0] != 0x4F434753u) return NULL;
  np = hdr[1];
  if (16 + 4L * np > n) return NULL;
  s = (refh_stream *)calloc(1, sizeof(*s));
  s->width = (int)hdr[2];
  s->height = (int)hdr[3];
  p = blob + 16 + 4L * np;
  for (i = 0; i < np; i++) {
    if (p + hdr[4 + i] > blob + n) { refh_stream_free(s); return NULL; }
    refh_stream_push(s, p, (long)hdr[4 + i]);
    p += hdr[4 + i];
  }
  return s;
}

/* Appends the data packets of `src` (skipping its 3 headers) to `dst`. */
REFH_API void refh_stream_append_data(refh_stream *dst, const refh_stream *src) {
  int i;
  for (i = 3; i < src->npackets; i++) refh_stream_push(dst, src->data[i], src->sizes[i]);
}

/* ---------------------------------------------------------------------- */
/* Synthetic content: moving gradient + checker + LCG noise,
   global motion (3,1) px/frame.  The LCG is re-seeded per frame from
   (seed, frame index) so GOP segments can be produced independently. */
static void refh_synth_frame_fmt(int w, int h, int cw, int ch, int f, int noise_shift, unsigned seed,
                                 unsigned char *y, unsigned char *cb, unsigned char *cr);
REFH_API void refh_synth_frame(int w, int h, int f, int noise_shift, unsigned seed,
                               unsigned char *y, unsigned char *cb, unsigned char *cr) {
  refh_synth_frame_fmt(w, h, w >> 1, h >> 1, f, noise_shift, seed, y, cb, cr);
}

/* chroma planes of cw x ch samples (4:2:0, 4:2:2 or 4:4:4): the chroma pattern
   is defined on the 4:2:0 grid and sampled at the plane's own resolution */
static void refh_synth_frame_fmt(int w, int h, int cw, int ch, int f, int noise_shift, unsigned seed,
                                 unsigned char *y, unsigned char *cb, unsigned char *cr) {
  uint32_t s = seed ^ ((uint32_t)f * 2654435761u);
  int sx = cw == w ? 1 : 0, sy = ch == h ? 1 : 0;
  int x, yy;
  for (yy = 0; yy < h; yy++) {
    for (x = 0; x < w; x++) {
      int v = ((2 * (x + 3 * f) + (yy + f)) & 255) + 60 * ((((x + 3 * f) >> 5) ^ ((yy + f) >> 5)) & 1);
      s = s * 1664525u + 1013904223u;
      v += (int)(s >> noise_shift);
      y[(size_t)yy * w + x] = (unsigned char)(v > 255 ? 255 : v);
    }
  }
  for (yy = 0; yy < ch; yy++) {
    for (x = 0; x < cw; x++) {
      cb[(size_t)yy * cw + x] = (unsigned char)(128 + ((((x >> sx) + f) >> 3) & 15));
      cr[(size_t)yy * cw + x] = (unsigned char)(128 - ((((yy >> sy) + 2 * f) >> 3) & 15));
    }
  }
}

/* Encodes frames [f0, f0+nframes) of the synthetic sequence at picture size
   w x h (4:2:0).  The coded frame is padded to a multiple of 16 and the picture
   centred the way encoder_example.c:1558-1563 does.  Returns 3 header packets
   followed by nframes data packets. */
/* only the integrated build (theora_b200/backend) defines this accessor */
extern long ocg_backend_enc_copy_recon(th_enc_ctx *enc, unsigned char *dst) __attribute__((weak));

REFH_API refh_stream *refh_encode_synth_recon(int w, int h, int f0, int nframes, int quality, int kf_interval,
                                              int speed, int noise_shift, unsigned seed, unsigned char *recon_out);
REFH_API refh_stream *refh_encode_synth_fmt(int w, int h, int f0, int nframes, int quality, int kf_interval,
                                            int speed, int noise_shift, unsigned seed, int pixel_fmt,
                                            unsigned char *recon_out);

REFH_API refh_stream *refh_encode_synth(int w, int h, int f0, int nframes, int quality,
                                        int kf_interval, int speed, int noise_shift, unsigned seed) {
  return refh_encode_synth_recon(w, h, f0, nframes, quality, kf_interval, speed, noise_shift, seed, NULL);
}

/* As refh_encode_synth; recon_out (optional, integrated build only) receives
   the encoder's reconstruction of the LAST frame. */
REFH_API refh_stream *refh_encode_synth_recon(int w, int h, int f0, int nframes, int quality, int kf_interval,
                                              int speed, int noise_shift, unsigned seed, unsigned char *recon_out) {
  return refh_encode_synth_fmt(w, h, f0, nframes, quality, kf_interval, speed, noise_shift, seed, TH_PF_420,
                               recon_out);
}

/* As above for any th_pixel_fmt (TH_PF_420 = 0, TH_PF_422 = 2, TH_PF_444 = 3). */
REFH_API refh_stream *refh_encode_synth_fmt(int w, int h, int f0, int nframes, int quality, int kf_interval,
                                            int speed, int noise_shift, unsigned seed, int pixel_fmt,
                                            unsigned char *recon_out) {
  th_info ti;
  th_enc_ctx *te;
  th_comment tc;
  th_ycbcr_buffer yuv;
  ogg_packet op;
  refh_stream *s;
  unsigned char *buf;
  int fw = (w + 15) & ~15, fh = (h + 15) & ~15;
  int cw = w >> !(pixel_fmt & 1), ch = h >> !(pixel_fmt & 2);
  int f, ret;
  ogg_uint32_t kf = (ogg_uint32_t)kf_interval;
  th_info_init(&ti);
  ti.frame_width = (ogg_uint32_t)fw;
  ti.frame_height = (ogg_uint32_t)fh;
  ti.pic_width = (ogg_uint32_t)w;
  ti.pic_height = (ogg_uint32_t)h;
  ti.pic_x = (ogg_uint32_t)(((fw - w) >> 1) & ~1);
  ti.pic_y = (ogg_uint32_t)(((fh - h) >> 1) & ~1);
  ti.fps_numerator = 30;
  ti.fps_denominator = 1;
  ti.aspect_numerator = 1;
  ti.aspect_denominator = 1;
  ti.colorspace = TH_CS_UNSPECIFIED;
  ti.pixel_fmt = (th_pixel_fmt)pixel_fmt;
  ti.target_bitrate = 0;
  ti.quality = quality;
  ti.keyframe_granule_shift = refh_ilog(kf_interval > 1 ? (unsigned)(kf_interval - 1) : 0);
  te = th_encode_alloc(&ti);
  if (te == NULL) { th_info_clear(&ti); return NULL; }
  th_encode_ctl(te, TH_ENCCTL_SET_KEYFRAME_FREQUENCY_FORCE, &kf, sizeof(kf));
  if (speed >= 0) th_encode_ctl(te, TH_ENCCTL_SET_SPLEVEL, &speed, sizeof(speed));
  s = (refh_stream *)calloc(1, sizeof(*s));
  s->width = w;
  s->height = h;
  th_comment_init(&tc);
  while ((ret = th_encode_flushheader(te, &tc, &op)) > 0) refh_stream_push(s, op.packet, op.bytes);
  th_comment_clear(&tc);
  buf = (unsigned char *)malloc((size_t)w * h + 2 * (size_t)cw * ch);
  yuv[0].width = w; yuv[0].height = h; yuv[0].stride = w; yuv[0].data = buf;
  yuv[1].width = cw; yuv[1].height = ch; yuv[1].stride = cw; yuv[1].data = buf + (size_t)w * h;
  yuv[2].width = cw; yuv[2].height = ch; yuv[2].stride = cw; yuv[2].data = yuv[1].data + (size_t)cw * ch;
  for (f = 0; f < nframes; f++) {
    refh_synth_frame_fmt(w, h, cw, ch, f0 + f, noise_shift, seed, yuv[0].data, yuv[1].data, yuv[2].data);
    ret = th_encode_ycbcr_in(te, yuv);
    if (ret < 0) break;
    while (th_encode_packetout(te, f + 1 >= nframes, &op) > 0) refh_stream_push(s, op.packet, op.bytes);
  }
  if (recon_out != NULL && ocg_backend_enc_copy_recon != NULL) ocg_backend_enc_copy_recon(te, recon_out);
  free(buf);
  th_encode_free(te);
  th_info_clear(&ti);
  return s;
}

/* Encoder timing loop for the CPU baseline: frames pre-generated in RAM,
   returns seconds spent in th_encode_ycbcr_in + th_encode_packetout. */
REFH_API double refh_encode_time(int w, int h, int nframes, int quality, int kf_interval, int speed,
                                 int noise_shift, unsigned seed, long *bytes_out) {
  th_info ti;
  th_enc_ctx *te;
  th_comment tc;
  th_ycbcr_buffer yuv;
  ogg_packet op;
  unsigned char *buf;
  size_t fsz;
  int fw = (w + 15) & ~15, fh = (h + 15) & ~15;
  int cw = w >> 1, ch = h >> 1;
  int f;
  long bytes = 0;
  double t0, t1;
  ogg_uint32_t kf = (ogg_uint32_t)kf_interval;
  th_info_init(&ti);
  ti.frame_width = (ogg_uint32_t)fw; ti.frame_height = (ogg_uint32_t)fh;
  ti.pic_width = (ogg_uint32_t)w; ti.pic_height = (ogg_uint32_t)h;
  ti.pic_x = (ogg_uint32_t)(((fw - w) >> 1) & ~1); ti.pic_y = (ogg_uint32_t)(((fh - h) >> 1) & ~1);
  ti.fps_numerator = 30; ti.fps_denominator = 1; ti.aspect_numerator = 1; ti.aspect_denominator = 1;
  ti.colorspace = TH_CS_UNSPECIFIED; ti.pixel_fmt = TH_PF_420; ti.target_bitrate = 0; ti.quality = quality;
  ti.keyframe_granule_shift = refh_ilog(kf_interval > 1 ? (unsigned)(kf_interval - 1) : 0);
  te = th_encode_alloc(&ti);
  if (te == NULL) return -1.0;
  th_encode_ctl(te, TH_ENCCTL_SET_KEYFRAME_FREQUENCY_FORCE, &kf, sizeof(kf));
  if (speed >= 0) th_encode_ctl(te, TH_ENCCTL_SET_SPLEVEL, &speed, sizeof(speed));
  th_comment_init(&tc);
  while (th_encode_flushheader(te, &tc, &op) > 0) {}
  th_comment_clear(&tc);
  fsz = (size_t)w * h + 2 * (size_t)cw * ch;
  buf = (unsigned char *)malloc(fsz * (size_t)nframes);
  for (f = 0; f < nframes; f++) {
    unsigned char *b = buf + fsz * (size_t)f;
    refh_synth_frame(w, h, f, noise_shift, seed, b, b + (size_t)w * h, b + (size_t)w * h + (size_t)cw * ch);
  }
  yuv[0].width = w; yuv[0].height = h; yuv[0].stride = w;
  yuv[1].width = cw; yuv[1].height = ch; yuv[1].stride = cw;
  yuv[2].width = cw; yuv[2].height = ch; yuv[2].stride = cw;
  t0 = refh_now();
  for (f = 0; f < nframes; f++) {
    unsigned char *b = buf + fsz * (size_t)f;
    yuv[0].data = b; yuv[1].data = b + (size_t)w * h; yuv[2].data = yuv[1].data + (size_t)cw * ch;
    if (th_encode_ycbcr_in(te, yuv) < 0) break;
    while (th_encode_packetout(te, f + 1 >= nframes, &op) > 0) bytes += op.bytes;
  }
  t1 = refh_now();
  free(buf);
  th_encode_free(te);
  th_info_clear(&ti);
  if (bytes_out) *bytes_out = bytes;
  return t1 - t0;
}

/* ---------------------------------------------------------------------- */
typedef struct refh_dec {
  th_info        ti;
  th_comment     tc;
  th_dec_ctx    *td;
  const refh_stream *s;
  int            next;
  th_ycbcr_buffer out;
} refh_dec;

REFH_API void refh_dec_close(refh_dec *d) {
  if (d == NULL) return;
  if (d->td) th_decode_free(d->td);
  th_comment_clear(&d->tc);
  th_info_clear(&d->ti);
  free(d);
}

REFH_API refh_dec *refh_dec_open(const refh_stream *s) {
  refh_dec *d = (refh_dec *)calloc(1, sizeof(*d));
  th_setup_info *ts = NULL;
  ogg_packet op;
  int i, ret = 0;
  th_info_init(&d->ti);
  th_comment_init(&d->tc);
  d->s = s;
  for (i = 0; i < 3 && i < s->npackets; i++) {
    memset(&op, 0, sizeof(op));
    op.packet = s->data[i];
    op.bytes = s->sizes[i];
    op.b_o_s = i == 0;
    op.packetno = i;
    ret = th_decode_headerin(&d->ti, &d->tc, &ts, &op);
    if (ret < 0) break;
  }
  if (ret >= 0) d->td = th_decode_alloc(&d->ti, ts);
  th_setup_free(ts);
  if (d->td == NULL) { refh_dec_close(d); return NULL; }
  d->next = 3;
  return d;
}

REFH_API void refh_dec_info(const refh_dec *d, int out[8]) {
  out[0] = (int)d->ti.frame_width; out[1] = (int)d->ti.frame_height;
  out[2] = (int)d->ti.pic_width; out[3] = (int)d->ti.pic_height;
  out[4] = (int)d->ti.pic_x; out[5] = (int)d->ti.pic_y;
  out[6] = (int)d->ti.pixel_fmt; out[7] = d->s->npackets - 3;
}

REFH_API void refh_dec_rewind(refh_dec *d) {
  ogg_int64_t gp = 0;
  d->next = 3;
  th_decode_ctl(d->td, TH_DECCTL_SET_GRANPOS, &gp, sizeof(gp));
}

/* Decodes the next data packet; returns th_decode_packetin's code, or 1000 at
   end of stream. */
REFH_API int refh_dec_next(refh_dec *d) {
  ogg_packet op;
  ogg_int64_t gp;
  int ret;
  if (d->next >= d->s->npackets) return 1000;
  memset(&op, 0, sizeof(op));
  op.packet = d->s->data[d->next];
  op.bytes = d->s->sizes[d->next];
  op.packetno = d->next;
  d->next++;
  ret = th_decode_packetin(d->td, &op, &gp);
  if (ret >= 0) th_decode_ycbcr_out(d->td, d->out);
  return ret;
}

static uint64_t refh_fnv_plane(const th_img_plane *p, uint64_t h) {
  int y, x;
  for (y = 0; y < p->height; y++) {
    const unsigned char *row = p->data + (ptrdiff_t)y * p->stride;
    for (x = 0; x < p->width; x++) { h ^= row[x]; h *= 1099511628211ULL; }
  }
  return h;
}

/* FNV-1a 64 of the three planes of the last decoded frame (full coded frame
   area, top-down, as th_decode_ycbcr_out presents it). */
REFH_API void refh_dec_hash(const refh_dec *d, uint64_t out[3]) {
  int pli;
  for (pli = 0; pli < 3; pli++) out[pli] = refh_fnv_plane(&d->out[pli], 14695981039346656037ULL);
}

/* Copies the last decoded frame, planes packed back to back, no padding. */
REFH_API long refh_dec_copy_frame(const refh_dec *d, unsigned char *dst) {
  long n = 0;
  int pli, y;
  for (pli = 0; pli < 3; pli++) {
    const th_img_plane *p = &d->out[pli];
    for (y = 0; y < p->height; y++) {
      memcpy(dst + n, p->data + (ptrdiff_t)y * p->stride, (size_t)p->width);
      n += p->width;
    }
  }
  return n;
}

REFH_API th_dec_ctx *refh_dec_ctx(refh_dec *d) { return d->td; }

/* TH_DECCTL_SET_PPLEVEL (theoradec.h); returns th_decode_ctl's code. */
REFH_API int refh_dec_set_pplevel(refh_dec *d, int level) {
  return th_decode_ctl(d->td, TH_DECCTL_SET_PPLEVEL, &level, sizeof(level));
}

/* ---------------------------------------------------------------------- */
/* Decode timing: `nthreads` independent decoders each decode the whole stream
   `passes` times (the library is single-threaded; streams share nothing).
   Returns wall seconds of the slowest thread; frames decoded = nthreads *
   passes * nframes. */
typedef struct refh_job {
  const refh_stream *s;
  pthread_barrier_t *bar;
  int passes;
  double secs;
  uint64_t hash;
  int fail;
} refh_job;

/* post-processing level the timed decoders run with (TH_DECCTL_SET_PPLEVEL; 0 = none) */
static int refh_timed_pplevel;
REFH_API void refh_set_timed_pplevel(int level) { refh_timed_pplevel = level; }

static void *refh_decode_worker(void *arg) {
  refh_job *j = (refh_job *)arg;
  refh_dec *d = refh_dec_open(j->s);
  if (d != NULL && refh_timed_pplevel > 0) refh_dec_set_pplevel(d, refh_timed_pplevel);
  double t0;
  int p;
  /* every worker finishes its set-up (decoder, device context, warm-up packet)
     before any starts the clock, and tears down only after all have stopped
     it, so allocation/free never overlaps a timed region */
  if (d != NULL) { refh_dec_next(d); refh_dec_rewind(d); }
  pthread_barrier_wait(j->bar);
  if (d == NULL) { j->fail = 1; pthread_barrier_wait(j->bar); return NULL; }
  t0 = refh_now();
  for (p = 0; p < j->passes; p++) {
    int ret;
    refh_dec_rewind(d);
    while ((ret = refh_dec_next(d)) != 1000) {
      if (ret < 0) { j->fail = 1; break; }
    }
  }
  j->secs = refh_now() - t0;
  pthread_barrier_wait(j->bar);
  {
    uint64_t h[3];
    refh_dec_hash(d, h);
    j->hash = h[0] ^ h[1] ^ h[2];
  }
  refh_dec_close(d);
  return NULL;
}

REFH_API double refh_decode_time(const refh_stream *s, int nthreads, int passes, uint64_t *hash_out) {
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  refh_job *jobs = (refh_job *)calloc((size_t)nthreads, sizeof(refh_job));
  pthread_barrier_t bar;
  double worst = 0.0;
  int i, fail = 0;
  pthread_barrier_init(&bar, NULL, (unsigned)nthreads);
  for (i = 0; i < nthreads; i++) {
    jobs[i].s = s;
    jobs[i].bar = &bar;
    jobs[i].passes = passes;
    pthread_create(&th[i], NULL, refh_decode_worker, &jobs[i]);
  }
  for (i = 0; i < nthreads; i++) {
    pthread_join(th[i], NULL);
    if (jobs[i].secs > worst) worst = jobs[i].secs;
    fail |= jobs[i].fail;
  }
  if (hash_out) *hash_out = jobs[0].hash;
  pthread_barrier_destroy(&bar);
  free(th);
  free(jobs);
  return fail ? -1.0 : worst;
}

/* ---------------------------------------------------------------------- */
/* Encode timing: `nthreads` independent encoders each encode the same
   `nframes` pre-generated frames (frames in RAM, packets discarded after
   hashing).  The first frame of every encoder is encoded before the clock
   starts (the library codes frame 0 twice to prime its statistics, and device
   contexts are created lazily), so the timed region is nframes-1 frames per
   thread.  Returns wall seconds of the slowest thread; the FNV-1a hash of all
   packet bytes of thread 0 goes to *hash_out. */
typedef struct refh_enc_job {
  pthread_barrier_t *bar;
  const unsigned char *frames;
  int w, h, nframes, quality, kf_interval, speed;
  double secs;
  uint64_t hash;
  long bytes;
  int fail;
} refh_enc_job;

static void *refh_encode_worker(void *arg) {
  refh_enc_job *j = (refh_enc_job *)arg;
  th_info ti;
  th_enc_ctx *te;
  th_comment tc;
  th_ycbcr_buffer yuv;
  ogg_packet op;
  int w = j->w, h = j->h;
  int fw = (w + 15) & ~15, fh = (h + 15) & ~15;
  int cw = w >> 1, ch = h >> 1;
  size_t fsz = (size_t)w * h + 2 * (size_t)cw * ch;
  uint64_t hash = 1469598103934665603ULL;
  ogg_uint32_t kf = (ogg_uint32_t)j->kf_interval;
  double t0 = 0.0;
  int f, started = 0;
  th_info_init(&ti);
  ti.frame_width = (ogg_uint32_t)fw; ti.frame_height = (ogg_uint32_t)fh;
  ti.pic_width = (ogg_uint32_t)w; ti.pic_height = (ogg_uint32_t)h;
  ti.pic_x = (ogg_uint32_t)(((fw - w) >> 1) & ~1); ti.pic_y = (ogg_uint32_t)(((fh - h) >> 1) & ~1);
  ti.fps_numerator = 30; ti.fps_denominator = 1; ti.aspect_numerator = 1; ti.aspect_denominator = 1;
  ti.colorspace = TH_CS_UNSPECIFIED; ti.pixel_fmt = TH_PF_420; ti.target_bitrate = 0; ti.quality = j->quality;
  ti.keyframe_granule_shift = refh_ilog(j->kf_interval > 1 ? (unsigned)(j->kf_interval - 1) : 0);
  te = th_encode_alloc(&ti);
  if (te != NULL) {
    th_encode_ctl(te, TH_ENCCTL_SET_KEYFRAME_FREQUENCY_FORCE, &kf, sizeof(kf));
    if (j->speed >= 0) th_encode_ctl(te, TH_ENCCTL_SET_SPLEVEL, &j->speed, sizeof(j->speed));
    th_comment_init(&tc);
    while (th_encode_flushheader(te, &tc, &op) > 0) {}
    th_comment_clear(&tc);
  }
  yuv[0].width = w; yuv[0].height = h; yuv[0].stride = w;
  yuv[1].width = cw; yuv[1].height = ch; yuv[1].stride = cw;
  yuv[2].width = cw; yuv[2].height = ch; yuv[2].stride = cw;
  for (f = 0; f < j->nframes && te != NULL; f++) {
    unsigned char *b = (unsigned char *)j->frames + fsz * (size_t)f;
    long i;
    if (f == 1) { pthread_barrier_wait(j->bar); started = 1; t0 = refh_now(); }
    yuv[0].data = b; yuv[1].data = b + (size_t)w * h; yuv[2].data = yuv[1].data + (size_t)cw * ch;
    if (th_encode_ycbcr_in(te, yuv) < 0) { j->fail = 1; break; }
    while (th_encode_packetout(te, f + 1 >= j->nframes, &op) > 0) {
      j->bytes += op.bytes;
      for (i = 0; i < op.bytes; i++) hash = (hash ^ op.packet[i]) * 1099511628211ULL;
    }
  }
  if (!started) { /* no encoder, fewer than two frames, or frame 0 failed */
    j->fail = 1;
    pthread_barrier_wait(j->bar);
    t0 = refh_now();
  }
  j->secs = refh_now() - t0;
  j->hash = hash;
  pthread_barrier_wait(j->bar);
  if (te != NULL) th_encode_free(te);
  th_info_clear(&ti);
  return NULL;
}

REFH_API double refh_encode_time_mt(int w, int h, int nframes, int quality, int kf_interval, int speed,
                                    int noise_shift, unsigned seed, int nthreads, uint64_t *hash_out,
                                    long *bytes_out) {
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  refh_enc_job *jobs = (refh_enc_job *)calloc((size_t)nthreads, sizeof(refh_enc_job));
  pthread_barrier_t bar;
  int cw = w >> 1, ch = h >> 1;
  size_t fsz = (size_t)w * h + 2 * (size_t)cw * ch;
  unsigned char *buf = (unsigned char *)malloc(fsz * (size_t)nframes);
  double worst = 0.0;
  int i, f, fail = 0;
  for (f = 0; f < nframes; f++) {
    unsigned char *b = buf + fsz * (size_t)f;
    refh_synth_frame(w, h, f, noise_shift, seed, b, b + (size_t)w * h, b + (size_t)w * h + (size_t)cw * ch);
  }
  pthread_barrier_init(&bar, NULL, (unsigned)nthreads);
  for (i = 0; i < nthreads; i++) {
    jobs[i].bar = &bar;
    jobs[i].frames = buf;
    jobs[i].w = w; jobs[i].h = h; jobs[i].nframes = nframes;
    jobs[i].quality = quality; jobs[i].kf_interval = kf_interval; jobs[i].speed = speed;
    pthread_create(&th[i], NULL, refh_encode_worker, &jobs[i]);
  }
  for (i = 0; i < nthreads; i++) {
    pthread_join(th[i], NULL);
    if (jobs[i].secs > worst) worst = jobs[i].secs;
    fail |= jobs[i].fail;
  }
  if (hash_out) *hash_out = jobs[0].hash;
  if (bytes_out) *bytes_out = jobs[0].bytes;
  pthread_barrier_destroy(&bar);
  free(th);
  free(jobs);
  free(buf);
  return fail ? -1.0 : worst;
}
