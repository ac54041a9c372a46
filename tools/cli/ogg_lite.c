/* ogg_lite.c -- see ogg_lite.h. */
#include <stdlib.h>
#include <string.h>
#include "ogg_lite.h"

/* RFC 3533 section 6: CRC-32, generator 0x04c11db7, initial value and final XOR 0, no bit reflection. */
static uint32_t g_crc_tab[256];
static int g_crc_ready;
static void crc_init(void) {
  uint32_t i, j, r;
  for (i = 0; i < 256; i++) {
    r = i << 24;
    for (j = 0; j < 8; j++) r = (r & 0x80000000u) ? (r << 1) ^ 0x04c11db7u : r << 1;
    g_crc_tab[i] = r;
  }
  g_crc_ready = 1;
}
uint32_t oggl_crc(const unsigned char *p, size_t n, uint32_t crc) {
  size_t i;
  if (!g_crc_ready) crc_init();
  for (i = 0; i < n; i++) crc = (crc << 8) ^ g_crc_tab[((crc >> 24) & 0xFF) ^ p[i]];
  return crc;
}

static void put32(unsigned char *p, uint32_t v) { p[0] = v & 0xFF; p[1] = (v >> 8) & 0xFF; p[2] = (v >> 16) & 0xFF; p[3] = (v >> 24) & 0xFF; }
static uint32_t get32(const unsigned char *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

/* ---- writer ---------------------------------------------------------------- */
int oggl_writer_init(oggl_writer *w, FILE *f, uint32_t serial) {
  memset(w, 0, sizeof(*w));
  w->f = f;
  w->serial = serial;
  w->bos_pending = 1;
  w->granulepos = -1;
  return 0;
}

static int body_reserve(oggl_writer *w, size_t extra) {
  if (w->body_len + extra > w->body_cap) {
    size_t cap = w->body_cap ? w->body_cap : 65536;
    unsigned char *nb;
    while (cap < w->body_len + extra) cap *= 2;
    nb = (unsigned char *)realloc(w->body, cap);
    if (nb == NULL) return -1;
    w->body = nb;
    w->body_cap = cap;
  }
  return 0;
}

static int emit_page(oggl_writer *w, int eos) {
  unsigned char hdr[27 + 255];
  uint64_t gp = (uint64_t)w->granulepos;
  uint32_t crc;
  int i, hlen = 27 + w->nsegs;
  memcpy(hdr, "OggS", 4);
  hdr[4] = 0;
  hdr[5] = (unsigned char)((w->continued ? 1 : 0) | (w->bos_pending ? 2 : 0) | (eos ? 4 : 0));
  for (i = 0; i < 8; i++) hdr[6 + i] = (unsigned char)(gp >> (8 * i));
  put32(hdr + 14, w->serial);
  put32(hdr + 18, w->pageno);
  put32(hdr + 22, 0);
  hdr[26] = (unsigned char)w->nsegs;
  memcpy(hdr + 27, w->lacing, (size_t)w->nsegs);
  crc = oggl_crc(hdr, (size_t)hlen, 0);
  crc = oggl_crc(w->body, w->body_len, crc);
  put32(hdr + 22, crc);
  if (fwrite(hdr, 1, (size_t)hlen, w->f) != (size_t)hlen) return -1;
  if (w->body_len && fwrite(w->body, 1, w->body_len, w->f) != w->body_len) return -1;
  w->pages_written++;
  w->bytes_written += hlen + (long)w->body_len;
  w->pageno++;
  w->bos_pending = 0;
  w->nsegs = 0;
  w->body_len = 0;
  w->granulepos = -1;
  w->continued = 0;
  return 0;
}

int oggl_write_packet(oggl_writer *w, const unsigned char *data, size_t len, int64_t granulepos, int eos) {
  /* a packet is len/255 segments of 255 plus one of len%255 (possibly 0), RFC 3533 section 5 */
  size_t left = len;
  const unsigned char *p = data;
  for (;;) {
    size_t seg = left >= 255 ? 255 : left;
    if (w->nsegs == 255) {
      /* page full in the middle of this packet: the next page continues it */
      if (emit_page(w, 0) < 0) return -1;
      w->continued = 1;
    }
    if (body_reserve(w, seg) < 0) return -1;
    memcpy(w->body + w->body_len, p, seg);
    w->body_len += seg;
    w->lacing[w->nsegs++] = (unsigned char)seg;
    p += seg;
    left -= seg;
    if (seg < 255) break;
  }
  w->granulepos = granulepos; /* this packet ends on the page under construction */
  if (eos || w->nsegs == 255) return emit_page(w, eos);
  return 0;
}

int oggl_writer_flush(oggl_writer *w, int eos) {
  if (w->nsegs == 0 && !eos) return 0;
  return emit_page(w, eos);
}

void oggl_writer_clear(oggl_writer *w) {
  free(w->body);
  memset(w, 0, sizeof(*w));
}

/* ---- reader ---------------------------------------------------------------- */
int oggl_reader_init(oggl_reader *r, FILE *f) {
  memset(r, 0, sizeof(*r));
  r->f = f;
  return 0;
}

/* Loads the next page of the followed stream into r->hdr / r->body.  1 ok, 0 EOF, <0 error. */
static int load_page(oggl_reader *r) {
  for (;;) {
    size_t blen = 0;
    uint32_t crc, want, serial, pageno;
    long cap_pos;
    int i, c, matched = 0;
    /* resynchronise on the capture pattern (RFC 3533 section 6.1) */
    while (matched < 4) {
      c = fgetc(r->f);
      if (c == EOF) return 0;
      if (c == "OggS"[matched]) matched++;
      else matched = c == 'O' ? 1 : 0;
    }
    cap_pos = ftell(r->f); /* just behind the capture pattern; -1 on a pipe */
    memcpy(r->hdr, "OggS", 4);
    if (fread(r->hdr + 4, 1, 23, r->f) != 23) { r->truncated = 1; return 0; }
    if (r->hdr[4] != 0) return -2; /* stream structure version */
    r->nsegs = r->hdr[26];
    if (fread(r->hdr + 27, 1, (size_t)r->nsegs, r->f) != (size_t)r->nsegs) { r->truncated = 1; return 0; }
    for (i = 0; i < r->nsegs; i++) blen += r->hdr[27 + i];
    if (blen > r->body_cap) {
      unsigned char *nb = (unsigned char *)realloc(r->body, blen);
      if (nb == NULL) return -3;
      r->body = nb;
      r->body_cap = blen;
    }
    if (fread(r->body, 1, blen, r->f) != blen) {
      /* a damaged segment table can claim more than the file holds: look for a page behind the pattern
         first, and call the file truncated only if there is none */
      r->truncated = 1;
      if (cap_pos >= 4 && fseek(r->f, cap_pos - 3, SEEK_SET) == 0) { r->crc_errors++; continue; }
      return 0;
    }
    r->truncated = 0;
    want = get32(r->hdr + 22);
    put32(r->hdr + 22, 0);
    crc = oggl_crc(r->hdr, (size_t)(27 + r->nsegs), 0);
    crc = oggl_crc(r->body, blen, crc);
    if (crc != want) {
      /* damaged page.  Its header (segment count, lacing values) may be what is damaged, so the size just
         skipped cannot be trusted: search again from the byte after this capture pattern's first byte, the
         way libogg's ogg_sync_pageseek does.  (On a pipe: carry on behind the page.) */
      r->crc_errors++;
      if (cap_pos >= 4) fseek(r->f, cap_pos - 3, SEEK_SET);
      continue;
    }
    serial = get32(r->hdr + 14);
    pageno = get32(r->hdr + 18);
    r->flags = r->hdr[5];
    if (!r->have_serial) {
      if (!(r->flags & 2)) continue; /* wait for a beginning-of-stream page */
      /* a multiplexed file starts with one BOS page per logical stream: follow the Theora one (its first
         packet is the identification header, 0x80 "theora", spec section 6.2) */
      if (blen < 7 || r->body[0] != 0x80 || memcmp(r->body + 1, "theora", 6) != 0) { r->other_streams++; continue; }
      r->have_serial = 1;
      r->serial = serial;
      r->next_pageno = pageno;
    }
    if (serial != r->serial) continue; /* another logical stream (e.g. audio) */
    if (pageno != r->next_pageno) {
      /* pages were lost: whatever was being assembled cannot be completed */
      r->lost_pages++;
      r->pkt_open = 0;
      r->pkt_len = 0;
    }
    r->next_pageno = pageno + 1;
    r->page_granule = 0;
    for (i = 7; i >= 0; i--) r->page_granule = (int64_t)(((uint64_t)r->page_granule << 8) | r->hdr[6 + i]);
    r->last_packet_seg = -1;
    for (i = 0; i < r->nsegs; i++) if (r->hdr[27 + i] < 255) r->last_packet_seg = i;
    r->seg = 0;
    r->body_pos = 0;
    r->first_packet_done = 0;
    r->pages_read++;
    r->page_loaded = 1;
    /* a page that does not continue a packet while one is open means a lost tail */
    if (!(r->flags & 1) && r->pkt_open) { r->pkt_open = 0; r->pkt_len = 0; }
    return 1;
  }
}

static int pkt_append(oggl_reader *r, const unsigned char *p, size_t n) {
  if (r->pkt_len + n > r->pkt_cap) {
    size_t cap = r->pkt_cap ? r->pkt_cap : 65536;
    unsigned char *nb;
    while (cap < r->pkt_len + n) cap *= 2;
    nb = (unsigned char *)realloc(r->pkt, cap);
    if (nb == NULL) return -1;
    r->pkt = nb;
    r->pkt_cap = cap;
  }
  memcpy(r->pkt + r->pkt_len, p, n);
  r->pkt_len += n;
  return 0;
}

int oggl_read_packet(oggl_reader *r, oggl_packet *out) {
  for (;;) {
    if (!r->page_loaded || r->seg >= r->nsegs) {
      int ret = load_page(r);
      if (ret <= 0) return ret;
      /* a continued page with nothing open: skip the orphaned tail */
      if ((r->flags & 1) && !r->pkt_open) {
        while (r->seg < r->nsegs) {
          int l = r->hdr[27 + r->seg];
          r->body_pos += (size_t)l;
          r->seg++;
          if (l < 255) break;
        }
      }
      continue;
    }
    if (!r->pkt_open) { r->pkt_open = 1; r->pkt_len = 0; }
    while (r->seg < r->nsegs) {
      int l = r->hdr[27 + r->seg];
      if (pkt_append(r, r->body + r->body_pos, (size_t)l) < 0) return -3;
      r->body_pos += (size_t)l;
      r->seg++;
      if (l < 255) {
        const int ended_at = r->seg - 1;
        out->data = r->pkt;
        out->len = r->pkt_len;
        out->granulepos = ended_at == r->last_packet_seg ? r->page_granule : -1;
        out->bos = (r->flags & 2) != 0 && !r->first_packet_done;
        out->eos = (r->flags & 4) != 0 && ended_at == r->last_packet_seg;
        r->first_packet_done = 1;
        r->pkt_open = 0;
        return 1;
      }
    }
    /* page exhausted inside a packet: it continues on the next page */
  }
}

void oggl_reader_clear(oggl_reader *r) {
  free(r->pkt);
  free(r->body);
  memset(r, 0, sizeof(*r));
}
