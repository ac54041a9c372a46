/* TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for libogg's <ogg/ogg.h>.
 *
 * libogg (>= 1.3.4, reference configure.ac:418) is not installed in this image
 * and is not vendored by the reference.  libtheora uses it only as a set of
 * integer typedefs, allocator macros, the ogg_packet struct and an MSb-first
 * bit *writer* (encode.c, huffenc.c, enquant.c, encinfo.c); none of the 8x8
 * block arithmetic lives there.  This header restates that public ABI from
 * libogg's documentation so the unmodified reference sources compile into
 * oracle/_ref/.  It is never linked into the product library.
 */
#ifndef OCG_OGG_STANDIN_H
#define OCG_OGG_STANDIN_H
#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int16_t  ogg_int16_t;
typedef uint16_t ogg_uint16_t;
typedef int32_t  ogg_int32_t;
typedef uint32_t ogg_uint32_t;
typedef int64_t  ogg_int64_t;
typedef uint64_t ogg_uint64_t;

#define _ogg_malloc  malloc
#define _ogg_calloc  calloc
#define _ogg_realloc realloc
#define _ogg_free    free

typedef struct {
  long           endbyte;
  int            endbit;
  unsigned char *buffer;
  unsigned char *ptr;
  long           storage;
} oggpack_buffer;

typedef struct {
  unsigned char *packet;
  long           bytes;
  long           b_o_s;
  long           e_o_s;
  ogg_int64_t    granulepos;
  ogg_int64_t    packetno;
} ogg_packet;

void           oggpackB_writeinit(oggpack_buffer *b);
void           oggpackB_write(oggpack_buffer *b, unsigned long value, int bits);
void           oggpackB_reset(oggpack_buffer *b);
long           oggpackB_bytes(oggpack_buffer *b);
unsigned char *oggpackB_get_buffer(oggpack_buffer *b);
void           oggpackB_writeclear(oggpack_buffer *b);
void           oggpack_write(oggpack_buffer *b, unsigned long value, int bits);
void           oggpack_writeclear(oggpack_buffer *b);

#ifdef __cplusplus
}
#endif
#endif
