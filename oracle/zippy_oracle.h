/*
 * zippy_oracle.h -- CPU oracle for the zippy-b200 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of guzba/zippy's codec core (deflate / inflate /
 * crc32 / adler32 and the gzip/zlib framing around them).  It exists so the
 * CUDA path can be checked bit-for-bit on a CPU; nothing in zippy_b200/ may
 * link, import or call it.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: INFLATE + CHECKSUMS are PINNED by the reference's own
 * fixtures (tests/golden/, see tests/test_oracle.py).  DEFLATE bytes are
 * "parity unpinned": the reference holds no golden compressed bytes; the
 * restatement is pinned by round trip through this inflate and system zlib.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to the reference checkout, src/zippy/...).
 */
#ifndef ZIPPY_ORACLE_H
#define ZIPPY_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes: one per distinct ZippyError message in the reference
 * (SURVEY.md section 8b "Error convention").  0 = success. */
enum {
  ZO_OK = 0,
  ZO_ERR_INVALID_LEVEL = 1,   /* deflate.nim:209 */
  ZO_ERR_INVALID_FORMAT = 2,  /* zippy.nim:84 */
  ZO_ERR_UNCOMPRESS = 3,      /* internal.nim:191-192 failUncompress */
  ZO_ERR_COMPRESS = 4,        /* internal.nim:194-195 failCompress */
  ZO_ERR_END_OF_BUFFER = 5,   /* bitstreams.nim:16-17 */
  ZO_ERR_BYTE_BOUNDARY = 6,   /* bitstreams.nim:66,113 */
  ZO_ERR_BLOCK_HEADER = 7,    /* inflate.nim:289 */
  ZO_ERR_INVALID_SYMBOL = 8,  /* inflate.nim:165 */
  ZO_ERR_DETECT = 9,          /* zippy.nim:125 */
  ZO_ERR_METHOD = 10,         /* zippy.nim:141, gzip.nim:26 */
  ZO_ERR_CINFO = 11,          /* zippy.nim:144 */
  ZO_ERR_HEADER = 12,         /* zippy.nim:147 */
  ZO_ERR_FDICT = 13,          /* zippy.nim:150 */
  ZO_ERR_CHECKSUM = 14,       /* zippy.nim:162, gzip.nim:81 */
  ZO_ERR_GZIP_ID = 15,        /* gzip.nim:23 */
  ZO_ERR_GZIP_RESERVED = 16,  /* gzip.nim:29 */
  ZO_ERR_GZIP_FLAGS = 17,     /* gzip.nim:41 */
  ZO_ERR_SIZE = 18,           /* gzip.nim:85-88 */
  ZO_ERR_NOMEM = 30
};

/* CompressedDataFormat, common.nim:4-5 (same ordinal values). */
enum { ZO_DF_DETECT = 0, ZO_DF_ZLIB = 1, ZO_DF_GZIP = 2, ZO_DF_DEFLATE = 3 };

/* Growable byte buffer standing in for the Nim `string` the reference appends to. */
typedef struct {
  uint8_t *data;
  size_t len;
  size_t cap;
} zo_buf;

void zo_buf_free(zo_buf *b);

uint32_t zo_crc32(const uint8_t *src, size_t len);   /* crc.nim:53-72 */
uint32_t zo_adler32(const uint8_t *src, size_t len); /* adler32.nim:6-63 */
/* the scalar bodies alone (crc.nim:29-51, adler32.nim:17-63): what zo_crc32 / zo_adler32 fall back
 * to without SSE4.1+PCLMUL / SSSE3; the SIMD forms are checked against them in tests/test_oracle.py */
uint32_t zo_crc32_scalar(const uint8_t *src, size_t len);
uint32_t zo_adler32_scalar(const uint8_t *src, size_t len);

/* deflate.nim:207-467: appends raw RFC1951 bytes to dst. */
int zo_deflate(zo_buf *dst, const uint8_t *src, size_t len, int level);
/* inflate.nim:268-291: decodes src[pos..len) into dst (dst->len reset to 0). */
int zo_inflate(zo_buf *dst, const uint8_t *src, size_t len, size_t pos);

/* zippy.nim:11-84.  fname_len: 0..25 fixes the gzip FNAME length the
 * reference draws from urandom (zippy.nim:28-42); <0 draws it from rand(). */
int zo_compress(zo_buf *dst, const uint8_t *src, size_t len, int level,
                int data_format, int fname_len);
/* zippy.nim:100-165 + gzip.nim:3-88. */
int zo_uncompress(zo_buf *dst, const uint8_t *src, size_t len, int data_format);

/* Convenience for ctypes / bench: batch over independent inputs with
 * `threads` host threads (the reference itself is single-threaded; this is
 * "all host cores over independent inputs").
 * offsets has n+1 entries.  Outputs are malloc'ed per input; out_lens[i]
 * receives each size, statuses[i] each error code.  If keep_outputs==0 the
 * buffers are freed immediately (timing only).  Returns total output bytes. */
uint64_t zo_compress_batch(const uint8_t *base, const uint64_t *offsets, size_t n,
                           int level, int data_format, int threads,
                           uint64_t *out_lens, int *statuses);
uint64_t zo_uncompress_batch(const uint8_t *base, const uint64_t *offsets, size_t n,
                             int data_format, int threads,
                             uint64_t *out_lens, int *statuses);

const char *zo_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif
