/*
 * zippy_b200.h -- C ABI of libzippy_b200.so, the B200 (sm_100a) replacement for the codec
 * core of guzba/zippy.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * The four single-input entry points are exactly the seam the reference's framing layer
 * calls:
 *     zb200_deflate  <->  proc deflate*(dst: var string, src, len, level)   src/zippy/deflate.nim:207
 *     zb200_inflate  <->  proc inflate*(dst: var string, src, len, pos)     src/zippy/inflate.nim:268
 *     zb200_crc32    <->  proc crc32*(src: pointer, len: int): uint32       src/zippy/crc.nim:53
 *     zb200_adler32  <->  proc adler32*(src: pointer, len: int): uint32     src/zippy/adler32.nim:6
 * The batch entry points have no reference counterpart (the reference is one input per
 * call); they are what makes a GPU worthwhile and what zippy.compress/uncompress map onto
 * when a caller has many inputs (e.g. ziparchives.nim:505-540 createZipArchive's loop).
 *
 * Conventions
 *  - every function returns a ZB200_* status (0 = ok) unless documented otherwise;
 *    statuses mirror the reference's ZippyError messages one to one (zb200_strerror).
 *  - the library never keeps or frees caller memory; `dst` buffers are caller-allocated
 *    (zb200_*_bound gives a sufficient size).
 *  - a zb200_ctx is bound to one CUDA device and one stream; calls on the same ctx
 *    serialise, different ctxs may be used from different host threads.
 *  - there is no CPU fallback: without a CUDA device zb200_init fails with ZB200_ERR_CUDA.
 */
#ifndef ZIPPY_B200_H
#define ZIPPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes; 1..18 are the reference's ZippyError messages (SURVEY.md 8b "Error convention") */
enum {
  ZB200_OK = 0,
  ZB200_ERR_INVALID_LEVEL = 1,   /* "Invalid compression level"             deflate.nim:209 */
  ZB200_ERR_INVALID_FORMAT = 2,  /* "Invalid data format"                   zippy.nim:84 */
  ZB200_ERR_UNCOMPRESS = 3,      /* "Invalid buffer, unable to uncompress"  internal.nim:191 */
  ZB200_ERR_COMPRESS = 4,        /* "Unexpected error while compressing"    internal.nim:194 */
  ZB200_ERR_END_OF_BUFFER = 5,   /* "Cannot read further, at end of buffer" bitstreams.nim:16 */
  ZB200_ERR_BYTE_BOUNDARY = 6,   /* "Must be at a byte boundary"            bitstreams.nim:66 */
  ZB200_ERR_BLOCK_HEADER = 7,    /* "Invalid block header"                  inflate.nim:289 */
  ZB200_ERR_INVALID_SYMBOL = 8,  /* "Invalid symbol"                        inflate.nim:165 */
  ZB200_ERR_DETECT = 9,          /* "Unable to detect compressed data format" zippy.nim:125 */
  ZB200_ERR_METHOD = 10,         /* "Unsupported compression method"        zippy.nim:141 */
  ZB200_ERR_CINFO = 11,          /* "Invalid compression info"              zippy.nim:144 */
  ZB200_ERR_HEADER = 12,         /* "Invalid header"                        zippy.nim:147 */
  ZB200_ERR_FDICT = 13,          /* "Preset dictionary is not yet supported" zippy.nim:150 */
  ZB200_ERR_CHECKSUM = 14,       /* "Checksum verification failed"          zippy.nim:162, gzip.nim:81 */
  ZB200_ERR_GZIP_ID = 15,        /* "Failed gzip identification values check" gzip.nim:23 */
  ZB200_ERR_GZIP_RESERVED = 16,  /* "Reserved flag bits set"                gzip.nim:29 */
  ZB200_ERR_GZIP_FLAGS = 17,     /* "Currently unsupported flags are set"   gzip.nim:41 */
  ZB200_ERR_SIZE = 18,           /* "Size verification failed"              gzip.nim:85 */
  /* new classes with no reference counterpart */
  ZB200_ERR_DST_TOO_SMALL = 19,
  ZB200_ERR_CUDA = 20,
  ZB200_ERR_NOMEM = 21,
  ZB200_ERR_ARG = 22
};

/* CompressedDataFormat (src/zippy/common.nim:4-5), same ordinals */
enum { ZB200_DF_DETECT = 0, ZB200_DF_ZLIB = 1, ZB200_DF_GZIP = 2, ZB200_DF_DEFLATE = 3 };
/* levels (src/zippy/common.nim:7-12) */
enum { ZB200_NO_COMPRESSION = 0, ZB200_BEST_SPEED = 1, ZB200_BEST_COMPRESSION = 9,
       ZB200_DEFAULT_COMPRESSION = -1, ZB200_HUFFMAN_ONLY = -2 };

typedef struct zb200_ctx zb200_ctx;

/* ---- lifecycle ---- */
int zb200_init(int device, zb200_ctx **out);  /* device < 0: current device */
void zb200_shutdown(zb200_ctx *ctx);
const char *zb200_strerror(int status);       /* the reference's message for 1..18 */
const char *zb200_last_cuda_error(zb200_ctx *ctx);
/* run this ctx's work on a caller-owned CUDA stream (a cudaStream_t passed as void*;
 * NULL restores the ctx's own stream) so callers can order and time it with their events */
int zb200_set_stream(zb200_ctx *ctx, void *cuda_stream);
int zb200_device_count(void);

/* ---- sizing ---- */
/* bound on raw-deflate bytes for `len` input bytes (stored path + per-block overhead;
 * the reference itself grows its output, bitstreams.nim:96-98) */
size_t zb200_deflate_bound(size_t len);
/* bound including the gzip (<= 36 + 8 bytes) / zlib (2 + 4) framing of zippy.nim:21-78 */
size_t zb200_compress_bound(size_t len, int data_format);

/* ---- single input, host buffers (the reference seam) ---- */
/* deflate.nim:207: raw RFC1951 stream, BFINAL on the last block, byte aligned at the end */
int zb200_deflate(zb200_ctx *ctx, const uint8_t *src, size_t len, int level,
                  uint8_t *dst, size_t dst_cap, size_t *dst_len);
/* inflate.nim:268: decode the raw stream that starts at byte `pos` of src[0..len) */
int zb200_inflate(zb200_ctx *ctx, const uint8_t *src, size_t len, size_t pos,
                  uint8_t *dst, size_t dst_cap, size_t *dst_len);
/* size the output of zb200_inflate without producing it */
int zb200_inflate_size(zb200_ctx *ctx, const uint8_t *src, size_t len, size_t pos, size_t *out_len);
/* One input whose output size is unknown, decoded ONCE: begin inflates (and verifies the trailer) into
 * library-owned device memory and reports the size, finish copies the bytes out.  This is what a caller
 * with a growing destination (the reference's `dst: var string`, inflate.nim:268-291, gzip.nim:3-88) binds
 * instead of inflate_size + inflate (two decodes).  data_format as for uncompress; `pos` as for inflate
 * (raw streams only).  A gzip member whose ISIZE understates its content gets the reference's verdict
 * (data, CRC check, then "Size verification failed"), not "destination too small". */
int zb200_decode_begin(zb200_ctx *ctx, const uint8_t *src, size_t len, int data_format, size_t pos, size_t *out_len);
int zb200_decode_finish(zb200_ctx *ctx, uint8_t *dst, size_t dst_cap, size_t *dst_len);
/* crc.nim:53 / adler32.nim:6 */
int zb200_crc32(zb200_ctx *ctx, const void *src, size_t len, uint32_t *out);
int zb200_adler32(zb200_ctx *ctx, const void *src, size_t len, uint32_t *out);

/* ---- batches of independent inputs, host buffers ----
 * input i is src_base[src_offsets[i] .. src_offsets[i+1]); offsets arrays have n+1 entries.
 * compress: zippy.compress(src, level, dataFormat) per input (zippy.nim:11-84); output i is
 * written at dst_base[dst_offsets[i] .. dst_offsets[i+1]) with dst_offsets filled in.
 * fname_lens (gzip only, may be NULL = all 0): number of 'a'.. letters the reference draws
 * at random for the FNAME field (zippy.nim:28-42); each in 0..25. */
int zb200_compress_batch(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                         int level, int data_format, const uint8_t *fname_lens,
                         uint8_t *dst_base, size_t dst_cap, uint64_t *dst_offsets, int *statuses);
/* uncompressed size of every input (gzip: ISIZE trailer, gzip.nim:66; zlib/raw: a counting
 * pass over the stream).  sizes[i] is only meaningful where statuses[i] == 0. */
int zb200_uncompress_sizes(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                           int data_format, uint64_t *sizes, int *statuses);
/* zippy.uncompress(src, dataFormat) per input (zippy.nim:100-165).  dst_offsets (n+1, in)
 * gives each output's slot; slot capacity is dst_offsets[i+1]-dst_offsets[i]; dst_lens[i]
 * receives the produced size.  A failing input sets statuses[i] and produces no output.
 * Members are independent and decoded in parallel; ONE member is a serial stream (one 8-lane group, ~10 MB/s),
 * so members of 512 KiB or more are first cut into parallel segments: at every byte-aligning empty stored
 * block 00 00 ff ff when the stream has them (this library's own multi-chunk output, zlib full-flush streams),
 * otherwise at dynamic-block starts found by testing every bit offset, each segment decoded with marker symbols
 * for its unknown 32 KiB window and resolved afterwards (any gzip / zlib output).  The serial decode is the
 * fallback for anything irregular -- results, including the error reported, are identical either way.
 * Limits: on the serial path one member's output is at most 4 GiB - 33 KiB (32-bit positions inside a member);
 * members decoded as segments have no such limit.
 * With page-locked host buffers the copies in and out overlap the kernels (member groups); pageable buffers
 * are staged through an internal pinned ring. */
int zb200_uncompress_batch(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                           int data_format, uint8_t *dst_base, const uint64_t *dst_offsets,
                           uint64_t *dst_lens, int *statuses);
/* crc32 (kind 0) or adler32 (kind 1) of every input */
int zb200_checksum_batch(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                         int kind, uint32_t *out);

/* ---- device-resident variants (pointers prefixed d_ are device memory on ctx's device;
 * offsets / statuses / sizes stay host arrays).  Used when the data already lives in HBM
 * (bench.py's `value`) and by the multi-GPU sharded path.  The call returns after the
 * work has completed on the ctx stream. ---- */
int zb200_compress_batch_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                                int level, int data_format, const uint8_t *fname_lens,
                                uint8_t *d_dst, size_t dst_cap, uint64_t *dst_offsets, int *statuses);
/* host inputs -> members left in DEVICE memory (H2D pipelined with the kernels).  The sharded
 * multi-GPU path uses it: compress, exchange the sizes, then copy each shard to its place. */
int zb200_compress_batch_h2d(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                             int level, int data_format, const uint8_t *fname_lens,
                             uint8_t *d_dst, size_t dst_cap, uint64_t *dst_offsets, int *statuses);
/* device -> host copy on the ctx stream (returns when it has landed): the second half of the sharded
 * path, once the size exchange has said where a shard goes in the concatenated host stream */
int zb200_download(zb200_ctx *ctx, const uint8_t *d_src, uint8_t *h_dst, size_t bytes);
int zb200_uncompress_batch_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                                  int data_format, uint8_t *d_dst, const uint64_t *dst_offsets,
                                  uint64_t *dst_lens, int *statuses);
int zb200_uncompress_sizes_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                                  int data_format, uint64_t *sizes, int *statuses);
int zb200_checksum_batch_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                                int kind, uint32_t *out);

/* ---- host memory ----
 * The host-buffer calls overlap their copies with the kernels only for page-locked memory.  A caller
 * that reuses a buffer (a Nim string it keeps, an mmap) can page-lock it once with these; buffers that
 * are not page-locked are staged through an internal pinned ring instead (slower than PCIe). */
int zb200_host_register(void *ptr, size_t bytes);
int zb200_host_unregister(void *ptr);

/* ---- several GPUs behind one call (SURVEY 8e at the boundary) ----
 * Members shard by contiguous index range, balanced by bytes; one host thread drives each device through its
 * own ctx; there is no data-path collective.  compress: every shard is compressed on its device, the per-shard
 * sizes are gathered, and each shard's bytes are copied to their place in ONE concatenated host stream
 * (dst_offsets are global).  The multi-process form (one rank per GPU, NCCL all_gather of the sizes) is
 * zippy_b200/sharding.py; this is the same path for a caller that is a single process, e.g. a Nim program.
 * devices == NULL or n_devices <= 0: every visible device.  A device may be listed more than once. */
typedef struct zb200_mgpu zb200_mgpu;
int zb200_mgpu_init(const int *devices, int n_devices, zb200_mgpu **out);
void zb200_mgpu_shutdown(zb200_mgpu *m);
int zb200_mgpu_device_count(zb200_mgpu *m);
int zb200_mgpu_compress_batch(zb200_mgpu *m, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                              int level, int data_format, const uint8_t *fname_lens,
                              uint8_t *dst_base, size_t dst_cap, uint64_t *dst_offsets, int *statuses);
int zb200_mgpu_uncompress_batch(zb200_mgpu *m, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                                int data_format, uint8_t *dst_base, const uint64_t *dst_offsets,
                                uint64_t *dst_lens, int *statuses);
int zb200_mgpu_checksum_batch(zb200_mgpu *m, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                              int kind, uint32_t *out);

/* ---- instrumentation (bench.py): device time in ms of the kernels of the last batch call,
 * measured with CUDA events on the ctx stream, and how many kernels it launched. ---- */
typedef struct {
  float lz_ms, huff_ms, scan_ms, pack_ms;      /* compress */
  float inflate_ms, verify_ms;                 /* uncompress */
  float checksum_ms;
  float h2d_ms, d2h_ms;                        /* host-buffer variants only */
  uint64_t h2d_bytes, d2h_bytes;
  uint32_t kernel_launches;
  uint32_t n_chunks;
} zb200_timing;
int zb200_last_timing(zb200_ctx *ctx, zb200_timing *out);

#ifdef __cplusplus
}
#endif
#endif
