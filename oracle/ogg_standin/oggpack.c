/* TEST INFRASTRUCTURE ONLY (oracle/): bit writer behind the <ogg/ogg.h>
 * stand-in.  Written from libogg's documented behaviour: values are masked to
 * `bits`, appended MSb-first (oggpackB_*) or LSb-first (oggpack_*), the buffer
 * grows on demand, and *_bytes() reports endbyte+(endbit+7)/8.
 */
#include <string.h>
#include "ogg/ogg.h"

#define OCG_PACK_CHUNK 4096

static int ocg_pack_reserve(oggpack_buffer *b, long extra) {
  if (b->ptr == NULL) return -1;
  if (b->endbyte + extra + 8 >= b->storage) {
    long nstorage = b->storage + extra + OCG_PACK_CHUNK;
    unsigned char *nb = (unsigned char *)realloc(b->buffer, (size_t)nstorage);
    if (nb == NULL) {
      free(b->buffer);
      memset(b, 0, sizeof(*b));
      return -1;
    }
    memset(nb + b->storage, 0, (size_t)(nstorage - b->storage));
    b->buffer = nb;
    b->storage = nstorage;
    b->ptr = nb + b->endbyte;
  }
  return 0;
}

void oggpackB_writeinit(oggpack_buffer *b) {
  memset(b, 0, sizeof(*b));
  b->buffer = (unsigned char *)calloc(OCG_PACK_CHUNK, 1);
  b->ptr = b->buffer;
  b->storage = b->buffer ? OCG_PACK_CHUNK : 0;
}

void oggpackB_reset(oggpack_buffer *b) {
  long used;
  if (b->buffer == NULL) return;
  used = b->endbyte + 16;
  if (used > b->storage) used = b->storage;
  memset(b->buffer, 0, (size_t)used);
  b->ptr = b->buffer;
  b->endbyte = 0;
  b->endbit = 0;
}

long oggpackB_bytes(oggpack_buffer *b) { return b->endbyte + (b->endbit + 7) / 8; }

unsigned char *oggpackB_get_buffer(oggpack_buffer *b) { return b->buffer; }

void oggpackB_writeclear(oggpack_buffer *b) {
  free(b->buffer);
  memset(b, 0, sizeof(*b));
}

void oggpack_writeclear(oggpack_buffer *b) { oggpackB_writeclear(b); }

/* MSb-first: the next bit written lands in the highest free bit of *ptr.
   Bytes past the write point are kept zero, so OR-ing is enough. */
void oggpackB_write(oggpack_buffer *b, unsigned long value, int bits) {
  if (bits < 0 || bits > 32) return;
  if (ocg_pack_reserve(b, 8) < 0) return;
  if (bits < 32) value &= (1UL << bits) - 1UL;
  else value &= 0xFFFFFFFFUL;
  while (bits > 0) {
    int room = 8 - b->endbit;
    int take = bits < room ? bits : room;
    unsigned long chunk = (value >> (bits - take)) & ((1UL << take) - 1UL);
    *b->ptr |= (unsigned char)(chunk << (room - take));
    bits -= take;
    b->endbit += take;
    if (b->endbit == 8) {
      b->endbit = 0;
      b->endbyte++;
      b->ptr++;
    }
  }
}

/* LSb-first variant (used by the reference only for byte-aligned 32-bit
   comment lengths, encinfo.c:72-80). */
void oggpack_write(oggpack_buffer *b, unsigned long value, int bits) {
  if (bits < 0 || bits > 32) return;
  if (ocg_pack_reserve(b, 8) < 0) return;
  if (bits < 32) value &= (1UL << bits) - 1UL;
  else value &= 0xFFFFFFFFUL;
  while (bits > 0) {
    int room = 8 - b->endbit;
    int take = bits < room ? bits : room;
    unsigned long chunk = value & ((1UL << take) - 1UL);
    *b->ptr |= (unsigned char)(chunk << b->endbit);
    value >>= take;
    bits -= take;
    b->endbit += take;
    if (b->endbit == 8) {
      b->endbit = 0;
      b->endbyte++;
      b->ptr++;
    }
  }
}
