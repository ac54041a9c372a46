/* ogg_lite -- the part of the Ogg container (RFC 3533; Theora mapping:
 * reference doc/spec/spec.tex appendix "Ogg Bitstream Encapsulation") that the
 * reference's example CLIs take from libogg (examples/dump_video.c:157-241,
 * examples/encoder_example.c:1059-1240): one logical stream, packets in,
 * pages out, and back.  libogg is not available in this image; this is an
 * independent implementation written from the RFC, not a copy of libogg. */
#ifndef OGG_LITE_H
#define OGG_LITE_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

/* ---- writer ---------------------------------------------------------------- */
typedef struct oggl_writer {
  FILE *f;
  uint32_t serial, pageno;
  int bos_pending;
  /* page under construction */
  unsigned char lacing[255];
  int nsegs;
  unsigned char *body;
  size_t body_len, body_cap;
  int64_t granulepos;      /* of the last packet that ENDS on this page, else -1 */
  int continued;           /* first segment continues a packet from the previous page */
  long pages_written, bytes_written;
} oggl_writer;

int  oggl_writer_init(oggl_writer *w, FILE *f, uint32_t serial);
/* Appends one packet; emits full pages as they fill.  eos marks the stream's last packet. */
int  oggl_write_packet(oggl_writer *w, const unsigned char *data, size_t len, int64_t granulepos, int eos);
/* Forces the page under construction out (after the header packets, at end of stream). */
int  oggl_writer_flush(oggl_writer *w, int eos);
void oggl_writer_clear(oggl_writer *w);

/* ---- reader ---------------------------------------------------------------- */
typedef struct oggl_packet {
  unsigned char *data;
  size_t len;
  int64_t granulepos;      /* page granulepos if the packet is the last to end on its page, else -1 */
  int bos, eos;
} oggl_packet;

typedef struct oggl_reader {
  FILE *f;
  int have_serial;
  uint32_t serial;         /* logical stream being followed: the first Theora stream that begins */
  uint32_t next_pageno;
  unsigned char *pkt;      /* packet being assembled */
  size_t pkt_len, pkt_cap;
  int pkt_open;
  /* current page */
  unsigned char hdr[27 + 255];
  unsigned char *body;
  size_t body_cap;
  int nsegs, seg, flags;
  size_t body_pos;
  int64_t page_granule;
  int last_packet_seg;     /* index of the segment that ends the last complete packet of the page */
  int page_loaded, first_packet_done;
  long pages_read, crc_errors, lost_pages;
  long other_streams;      /* beginning-of-stream pages of other (non-Theora) logical streams seen before ours */
  int truncated;           /* the file ended inside a page */
} oggl_reader;

int  oggl_reader_init(oggl_reader *r, FILE *f);
/* 1 = packet returned (valid until the next call), 0 = end of file, <0 = malformed stream. */
int  oggl_read_packet(oggl_reader *r, oggl_packet *out);
void oggl_reader_clear(oggl_reader *r);

uint32_t oggl_crc(const unsigned char *p, size_t n, uint32_t crc);
#endif
