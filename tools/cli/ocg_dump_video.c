/* ocg_dump_video -- the job of the reference's examples/dump_video.c (289-585)
 * on top of whichever libtheora build it is linked against (the B200
 * back-end's libth_ocg.so, or the unmodified reference for the CPU plumbing
 * case, BASELINE configs[0]): Ogg file in, YUV4MPEG2 (or raw planes) out,
 * decoded through th_decode_headerin / th_decode_alloc / th_decode_packetin
 * with the striped-decode callback, exactly as the example does.  The Ogg
 * layer is tools/cli/ogg_lite.c (libogg is not in this image). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "theora/theoradec.h"
#include "ogg_lite.h"

static const char *CHROMA_TYPES[4] = {"420jpeg", NULL, "422jpeg", "444"};
static th_info ti;
static th_ycbcr_buffer ycbcr;

/* dump_video.c:157-175 */
static void stripe_decoded(void *ctx, th_ycbcr_buffer src, int fragy0, int fragy_end) {
  th_img_plane *dst = (th_img_plane *)ctx;
  int pli;
  for (pli = 0; pli < 3; pli++) {
    int yshift = pli != 0 && !(ti.pixel_fmt & 2);
    int y_end = fragy_end << (3 - yshift), y;
    for (y = fragy0 << (3 - yshift); y < y_end; y++)
      memcpy(dst[pli].data + y * dst[pli].stride, src[pli].data + y * src[pli].stride, (size_t)src[pli].width);
  }
}

static void usage(void) {
  fprintf(stderr, "usage: ocg_dump_video [-o out.y4m] [-c] [-r] [-f] [-z] [-p pplevel] in.ogv\n"
                  "  -c crop to the picture region   -r raw planes, no YUV4MPEG2 framing   -f only report fps\n"
                  "  -z zero-copy: no stripe callback; rows are written straight out of the buffer th_decode_ycbcr_out\n"
                  "     hands out (with the B200 back-end: the page-locked buffer the device wrote the frame into)\n");
  exit(1);
}

int main(int argc, char **argv) {
  const char *in = NULL, *out = NULL;
  int crop = 0, raw = 0, fps_only = 0, pplevel = 0, zero_copy = 0, i;
  FILE *fin, *fout = NULL;
  oggl_reader rd;
  oggl_packet pk;
  th_comment tc;
  th_setup_info *ts = NULL;
  th_dec_ctx *td = NULL;
  ogg_packet op;
  long frames = 0, packetno = 0;
  int headers_done = 0, ret;
  struct timespec t0, t1;
  for (i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
    else if (!strcmp(argv[i], "-c")) crop = 1;
    else if (!strcmp(argv[i], "-r")) raw = 1;
    else if (!strcmp(argv[i], "-f")) fps_only = 1;
    else if (!strcmp(argv[i], "-z")) zero_copy = 1;
    else if (!strcmp(argv[i], "-p") && i + 1 < argc) pplevel = atoi(argv[++i]);
    else if (argv[i][0] == '-') usage();
    else in = argv[i];
  }
  if (in == NULL) usage();
  fin = fopen(in, "rb");
  if (fin == NULL) { perror(in); return 1; }
  if (out != NULL && !fps_only) {
    fout = strcmp(out, "-") ? fopen(out, "wb") : stdout;
    if (fout == NULL) { perror(out); return 1; }
  }
  oggl_reader_init(&rd, fin);
  th_info_init(&ti);
  th_comment_init(&tc);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  while ((ret = oggl_read_packet(&rd, &pk)) > 0) {
    memset(&op, 0, sizeof(op));
    op.packet = pk.data;
    op.bytes = (long)pk.len;
    op.b_o_s = pk.bos;
    op.e_o_s = pk.eos;
    op.granulepos = pk.granulepos;
    op.packetno = packetno++;
    if (!headers_done) {
      int hr = th_decode_headerin(&ti, &tc, &ts, &op);
      if (hr < 0) { fprintf(stderr, "not a Theora stream (th_decode_headerin: %d)\n", hr); return 1; }
      if (hr > 0) continue; /* a header packet was consumed */
      /* hr == 0: first data packet: set the decoder up (dump_video.c:452-511) */
      td = th_decode_alloc(&ti, ts);
      th_setup_free(ts);
      if (td == NULL) { fprintf(stderr, "th_decode_alloc failed (no usable device?)\n"); return 1; }
      if (pplevel > 0 && th_decode_ctl(td, TH_DECCTL_SET_PPLEVEL, &pplevel, sizeof(pplevel)) < 0) {
        fprintf(stderr, "post-processing level %d refused\n", pplevel);
        return 1;
      }
      if (!zero_copy) {
        th_stripe_callback cb;
        int pli;
        for (pli = 0; pli < 3; pli++) {
          int xs = pli != 0 && !(ti.pixel_fmt & 1), ys = pli != 0 && !(ti.pixel_fmt & 2);
          ycbcr[pli].width = (int)ti.frame_width >> xs;
          ycbcr[pli].height = (int)ti.frame_height >> ys;
          ycbcr[pli].stride = ycbcr[pli].width;
          ycbcr[pli].data = (unsigned char *)malloc((size_t)ycbcr[pli].width * ycbcr[pli].height);
        }
        cb.ctx = ycbcr;
        cb.stripe_decoded = (th_stripe_decoded_func)stripe_decoded;
        th_decode_ctl(td, TH_DECCTL_SET_STRIPE_CB, &cb, sizeof(cb));
      }
      if (fout != NULL && !raw) {
        int hdec = !(ti.pixel_fmt & 1), vdec = !(ti.pixel_fmt & 2);
        int w = (int)ti.frame_width, h = (int)ti.frame_height;
        if (crop) {
          if ((ti.pic_x & hdec) || (ti.pic_width & hdec) || (ti.pic_y & vdec) || (ti.pic_height & vdec)) {
            fprintf(stderr, "cropped images with odd offsets/sizes and chroma subsampling cannot be output to YUV4MPEG2\n");
            return 1;
          }
          w = (int)ti.pic_width;
          h = (int)ti.pic_height;
        }
        fprintf(fout, "YUV4MPEG2 C%s W%d H%d F%d:%d I%c A%d:%d\n", CHROMA_TYPES[ti.pixel_fmt], w, h,
                (int)ti.fps_numerator, (int)ti.fps_denominator, 'p', (int)ti.aspect_numerator, (int)ti.aspect_denominator);
      }
      headers_done = 1;
    }
    {
      ogg_int64_t gp;
      int dr = th_decode_packetin(td, &op, &gp);
      if (dr < 0 && dr != TH_DUPFRAME) { fprintf(stderr, "th_decode_packetin: %d\n", dr); return 1; }
      frames++;
      /* dump_video.c:245-250 ("normal, non-striped decoding"): the decoder's own buffer, no copy in between */
      if (zero_copy && (fout != NULL || fps_only) && th_decode_ycbcr_out(td, ycbcr) < 0) { fprintf(stderr, "th_decode_ycbcr_out failed\n"); return 1; }
      if (fout != NULL) { /* dump_video.c:203-241 */
        int x0 = 0, y0 = 0, xend = (int)ti.frame_width, yend = (int)ti.frame_height, hdec = 0, vdec = 0, pli, y;
        if (crop) { x0 = (int)ti.pic_x; y0 = (int)ti.pic_y; xend = x0 + (int)ti.pic_width; yend = y0 + (int)ti.pic_height; }
        if (!raw) fprintf(fout, "FRAME\n");
        for (pli = 0; pli < 3; pli++) {
          for (y = y0 >> vdec; y < ((yend + vdec) >> vdec); y++)
            fwrite(ycbcr[pli].data + ycbcr[pli].stride * y + (x0 >> hdec), 1, (size_t)(((xend + hdec) >> hdec) - (x0 >> hdec)), fout);
          hdec = !(ti.pixel_fmt & 1);
          vdec = !(ti.pixel_fmt & 2);
        }
      }
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (ret < 0) { fprintf(stderr, "malformed Ogg stream (%d)\n", ret); return 1; }
  if (!headers_done) { fprintf(stderr, "no Theora data packets found\n"); return 1; }
  {
    double secs = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    fprintf(stderr, "%ld frames, %ld pages (%ld bad CRC, %ld lost%s), %.2f fps\n", frames, rd.pages_read, rd.crc_errors,
            rd.lost_pages, rd.truncated ? ", file truncated inside a page" : "", secs > 0 ? frames / secs : 0.0);
  }
  th_decode_free(td);
  th_comment_clear(&tc);
  th_info_clear(&ti);
  oggl_reader_clear(&rd);
  if (fout != NULL && fout != stdout) fclose(fout);
  fclose(fin);
  if (!zero_copy) for (i = 0; i < 3; i++) free(ycbcr[i].data);
  return 0;
}
