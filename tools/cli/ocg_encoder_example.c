/* ocg_encoder_example -- the video half of the reference's
 * examples/encoder_example.c (1059-1240, 1550-1830): YUV4MPEG2 in, Ogg Theora
 * out, through th_encode_alloc / th_encode_flushheader / th_encode_ycbcr_in /
 * th_encode_packetout of whichever libtheora build it is linked against.
 * Picture placement follows encoder_example.c:1556-1563 (frame padded to a
 * multiple of 16, picture centred on even offsets).  4:2:0 (C420jpeg and the
 * other 4:2:0 sitings, taken as they are), 4:2:2 and 4:4:4 input. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "theora/theoraenc.h"
#include "ogg_lite.h"

static int ilog(unsigned v) { int r = 0; while (v) { r++; v >>= 1; } return r; }

static void usage(void) {
  fprintf(stderr, "usage: ocg_encoder_example [-o out.ogv] [-v quality 0..10] [-k keyframe-freq] [-z speed] [-n max-frames] in.y4m\n");
  exit(1);
}

int main(int argc, char **argv) {
  const char *in = NULL, *out = "out.ogv";
  double vq = 6.0;
  ogg_uint32_t kf = 64;
  int speed = -1, max_frames = -1, i;
  FILE *fin, *fout;
  char line[256], chroma[32] = "420jpeg";
  int pic_w = 0, pic_h = 0, fps_n = 30, fps_d = 1, par_n = 0, par_d = 0;
  int frame_w, frame_h, pic_x, pic_y, hdec, vdec, cw, ch, ret;
  th_info ti;
  th_enc_ctx *te;
  th_comment tc;
  ogg_packet op;
  oggl_writer ow;
  th_ycbcr_buffer yuv;
  unsigned char *planes, *frame;
  size_t in_sz, nframes = 0;
  int have_next;
  for (i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
    else if (!strcmp(argv[i], "-v") && i + 1 < argc) vq = atof(argv[++i]);
    else if (!strcmp(argv[i], "-k") && i + 1 < argc) kf = (ogg_uint32_t)atoi(argv[++i]);
    else if (!strcmp(argv[i], "-z") && i + 1 < argc) speed = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-n") && i + 1 < argc) max_frames = atoi(argv[++i]);
    else if (argv[i][0] == '-' && argv[i][1]) usage();
    else in = argv[i];
  }
  if (in == NULL || vq < 0 || vq > 10 || kf < 1) usage();
  fin = strcmp(in, "-") ? fopen(in, "rb") : stdin;
  if (fin == NULL) { perror(in); return 1; }
  if (fgets(line, sizeof(line), fin) == NULL || strncmp(line, "YUV4MPEG2 ", 10)) { fprintf(stderr, "not a YUV4MPEG2 file\n"); return 1; }
  {
    char *tok = strtok(line + 10, " \n");
    while (tok != NULL) {
      switch (tok[0]) {
        case 'W': pic_w = atoi(tok + 1); break;
        case 'H': pic_h = atoi(tok + 1); break;
        case 'F': sscanf(tok + 1, "%d:%d", &fps_n, &fps_d); break;
        case 'A': sscanf(tok + 1, "%d:%d", &par_n, &par_d); break;
        case 'C': strncpy(chroma, tok + 1, sizeof(chroma) - 1); break;
        case 'I': if (tok[1] != 'p' && tok[1] != '?') { fprintf(stderr, "interlaced input is not supported\n"); return 1; } break;
        default: break;
      }
      tok = strtok(NULL, " \n");
    }
  }
  if (pic_w <= 0 || pic_h <= 0) { fprintf(stderr, "bad YUV4MPEG2 header\n"); return 1; }
  if (!strncmp(chroma, "420", 3)) { hdec = vdec = 1; }
  else if (!strncmp(chroma, "422", 3)) { hdec = 1; vdec = 0; }
  else if (!strncmp(chroma, "444", 3) && strcmp(chroma, "444alpha")) { hdec = vdec = 0; }
  else { fprintf(stderr, "unsupported chroma format C%s\n", chroma); return 1; }
  frame_w = (pic_w + 15) & ~0xF;
  frame_h = (pic_h + 15) & ~0xF;
  pic_x = ((frame_w - pic_w) >> 1) & ~1;
  pic_y = ((frame_h - pic_h) >> 1) & ~1;
  th_info_init(&ti);
  ti.frame_width = (ogg_uint32_t)frame_w;
  ti.frame_height = (ogg_uint32_t)frame_h;
  ti.pic_width = (ogg_uint32_t)pic_w;
  ti.pic_height = (ogg_uint32_t)pic_h;
  ti.pic_x = (ogg_uint32_t)pic_x;
  ti.pic_y = (ogg_uint32_t)pic_y;
  ti.fps_numerator = (ogg_uint32_t)fps_n;
  ti.fps_denominator = (ogg_uint32_t)fps_d;
  ti.aspect_numerator = (ogg_uint32_t)par_n;
  ti.aspect_denominator = (ogg_uint32_t)par_d;
  ti.colorspace = TH_CS_UNSPECIFIED;
  ti.target_bitrate = 0;
  ti.quality = (int)rint(6.3 * vq);
  ti.keyframe_granule_shift = ilog(kf - 1);
  ti.pixel_fmt = hdec ? (vdec ? TH_PF_420 : TH_PF_422) : TH_PF_444;
  te = th_encode_alloc(&ti);
  if (te == NULL) { fprintf(stderr, "th_encode_alloc failed\n"); return 1; }
  th_encode_ctl(te, TH_ENCCTL_SET_KEYFRAME_FREQUENCY_FORCE, &kf, sizeof(kf));
  if (speed >= 0) {
    int smax = 0;
    th_encode_ctl(te, TH_ENCCTL_GET_SPLEVEL_MAX, &smax, sizeof(smax));
    if (speed > smax) speed = smax;
    th_encode_ctl(te, TH_ENCCTL_SET_SPLEVEL, &speed, sizeof(speed));
  }
  fout = strcmp(out, "-") ? fopen(out, "wb") : stdout;
  if (fout == NULL) { perror(out); return 1; }
  oggl_writer_init(&ow, fout, 0x4f434742u);
  th_comment_init(&tc);
  /* header packets: the first on a page of its own, the rest flushed before any data (spec.tex A.2.3,
     encoder_example.c:1716-1790) */
  i = 0;
  while ((ret = th_encode_flushheader(te, &tc, &op)) > 0) {
    if (oggl_write_packet(&ow, op.packet, (size_t)op.bytes, op.granulepos, 0) < 0) return 1;
    if (i++ == 0 && oggl_writer_flush(&ow, 0) < 0) return 1;
  }
  if (ret < 0 || oggl_writer_flush(&ow, 0) < 0) { fprintf(stderr, "header output failed\n"); return 1; }
  cw = (pic_w + hdec) >> hdec;
  ch = (pic_h + vdec) >> vdec;
  in_sz = (size_t)pic_w * pic_h + 2 * (size_t)cw * ch;
  planes = (unsigned char *)malloc(2 * in_sz);
  /* th_encode_ycbcr_in wants frame-sized planes; only the picture region is read (encode.c:1640-1700) */
  frame = (unsigned char *)calloc((size_t)frame_w * frame_h + 2 * ((size_t)frame_w >> hdec) * ((size_t)frame_h >> vdec), 1);
  if (planes == NULL || frame == NULL) return 1;
  yuv[0].width = frame_w; yuv[0].height = frame_h; yuv[0].stride = frame_w; yuv[0].data = frame;
  yuv[1].width = frame_w >> hdec; yuv[1].height = frame_h >> vdec; yuv[1].stride = frame_w >> hdec;
  yuv[1].data = frame + (size_t)frame_w * frame_h;
  yuv[2] = yuv[1];
  yuv[2].data = yuv[1].data + (size_t)yuv[1].stride * yuv[1].height;
  /* one frame of look-ahead so the last packet can carry the end-of-stream flag (encoder_example.c:1100-1127) */
  have_next = 0;
  for (;;) {
    unsigned char *cur = planes + (nframes & 1) * in_sz;
    int last, y;
    if (!have_next) {
      if (fgets(line, sizeof(line), fin) == NULL || strncmp(line, "FRAME", 5) || fread(cur, 1, in_sz, fin) != in_sz) break;
    }
    have_next = 0;
    if (max_frames < 0 || (long)nframes + 1 < max_frames) {
      unsigned char *nxt = planes + ((nframes + 1) & 1) * in_sz;
      if (fgets(line, sizeof(line), fin) != NULL && !strncmp(line, "FRAME", 5) && fread(nxt, 1, in_sz, fin) == in_sz) have_next = 1;
    }
    last = !have_next;
    for (y = 0; y < pic_h; y++) memcpy(yuv[0].data + (size_t)(y + pic_y) * yuv[0].stride + pic_x, cur + (size_t)y * pic_w, (size_t)pic_w);
    for (i = 1; i < 3; i++) {
      const unsigned char *src = cur + (size_t)pic_w * pic_h + (size_t)(i - 1) * cw * ch;
      for (y = 0; y < ch; y++)
        memcpy(yuv[i].data + (size_t)(y + (pic_y >> vdec)) * yuv[i].stride + (pic_x >> hdec), src + (size_t)y * cw, (size_t)cw);
    }
    if (th_encode_ycbcr_in(te, yuv) < 0) { fprintf(stderr, "th_encode_ycbcr_in failed\n"); return 1; }
    while ((ret = th_encode_packetout(te, last, &op)) > 0) {
      if (oggl_write_packet(&ow, op.packet, (size_t)op.bytes, op.granulepos, (int)op.e_o_s) < 0) return 1;
      /* one page per packet keeps seeking granularity at a frame, like ogg_stream_pageout on video packets */
      if (!op.e_o_s && oggl_writer_flush(&ow, 0) < 0) return 1;
    }
    if (ret < 0) { fprintf(stderr, "th_encode_packetout failed (%d)\n", ret); return 1; }
    nframes++;
    if (last) break;
  }
  oggl_writer_flush(&ow, 0);
  fprintf(stderr, "%lu frames, %ld pages, %ld bytes\n", (unsigned long)nframes, ow.pages_written, ow.bytes_written);
  th_encode_free(te);
  th_comment_clear(&tc);
  oggl_writer_clear(&ow);
  if (fout != stdout) fclose(fout);
  if (fin != stdin) fclose(fin);
  free(planes);
  free(frame);
  return 0;
}
