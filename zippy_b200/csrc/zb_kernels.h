// zb_kernels.h -- launch interface between the C-ABI layer (zb_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "zb_common.h"
#include "zb_crc.h"
#include "zb_huff.h"

// One unit of compress work: <= 64 KiB of one member, one DEFLATE block.
struct ZbChunkDesc {
  uint64_t src_off;   // byte offset of the chunk in the source buffer
  uint32_t len;       // 0..65536
  uint32_t member;    // index of the input this chunk belongs to
  uint32_t flags;     // bit0: first chunk of its member, bit1: last chunk
  uint32_t pad;       // k_lz2: bytes of the member's preceding data staged as history (<= 32768)
};
#define ZB_CHUNK_FIRST 1u
#define ZB_CHUNK_LAST 2u

struct ZbChunkCheck {
  uint32_t crc_raw;   // init-0 CRC of the chunk bytes
  uint32_t adler;     // Adler-32 of the chunk bytes as a standalone message
};

// All device scratch for one compress batch (arrays sized by n_chunks / n_members).
struct ZbCompressWork {
  const uint8_t *src;          // device
  uint8_t *dst;                // device, 4-byte aligned; every byte of the output is written explicitly by k_pack
  uint64_t dst_cap;            // bytes (multiple of 4): a group that would end beyond this is not written at all
  const ZbChunkDesc *desc;     // [n_chunks]
  const uint32_t *member_first;// [n_members + 1] first chunk index of each member
  const uint8_t *fname_len;    // [n_members] gzip FNAME letters (0..25) or nullptr
  uint2 *masks;                // [n_chunks][2048] (token-start mask, is-match mask) per 32-byte window
  uint32_t *recs;              // [n_chunks][8 sub-chunks][2048] match records, one dense stream per sub-chunk
  uint16_t *hist;              // [n_chunks][8][316]
  ZbChunkCheck *chk;           // [n_chunks]
  ZbCodebook *cb;              // [n_chunks]
  uint64_t *chunk_off;         // [n_chunks] byte offset of each chunk's deflate bytes in dst
  uint64_t *member_off;        // [n_members + 1] output offsets (member_off[n] = total)
  uint32_t *member_check;      // [n_members] crc32 (gzip) or adler32 (zlib) of the whole member
  uint32_t *member_isize;      // [n_members] input size mod 2^32 (gzip ISIZE)
  uint64_t out_base;           // byte offset in dst where this group's first member starts ...
  const uint64_t *out_base_ptr;// ... or, when non-null, a device word holding it (the previous group's end),
                               // so consecutive groups can be enqueued without a host round trip
  const ZbCrcTables *tabs;     // device
  uint2 *lz2_tables;           // k_lz2 dictionaries: [grid][8 own tables of 2048 x 4 ways + 12 segment tables of 8192 entries], u16 positions (LZ levels only)
  uint32_t n_chunks, n_members;
  int level, data_format;
};

// per-device kernel attributes (dynamic shared memory limits); call with the device current
cudaError_t zb_setup_deflate_attrs();
cudaError_t zb_setup_inflate_attrs();
struct ZbLz2Params {
  uint32_t own_ways;   // ways of the own bucket looked at (1..4), after the closest same-hash position of the window
  uint32_t hist_segs;  // preceding segments whose entry is looked at (0..4), nearest first
  uint32_t maxcand;    // candidates per position that may pass the 4-byte check and be extended
  uint32_t good;       // a match this long leaves room for one more candidate only
  uint32_t lazy;       // matches shorter than this yield to a longer match at the next position (0: greedy)
};
ZbLz2Params zb_lz2_params(int level);
size_t zb_lz2_table_bytes(int *grid_out);
cudaError_t zb_launch_lz(const ZbCompressWork &w, cudaStream_t s);
cudaError_t zb_launch_huff(const ZbCompressWork &w, cudaStream_t s);
cudaError_t zb_launch_scan(const ZbCompressWork &w, cudaStream_t s);
cudaError_t zb_launch_pack(const ZbCompressWork &w, cudaStream_t s);
// ---- inflate ----
struct ZbInflateWork {
  const uint8_t *src;          // device
  const uint64_t *src_off;     // device [n+1]
  uint8_t *dst;                // device (unused when count_only)
  const uint64_t *dst_off;     // device [n+1]; capacity of member i = dst_off[i+1]-dst_off[i]
  uint64_t *out_len;           // device [n]
  int *status;                 // device [n]
  uint32_t *expect;            // device [n] trailer checksum
  uint32_t *kind;              // device [n] resolved format (ZB_DF_*) per member
  uint32_t *counter;           // device work-queue counter (zeroed before launch)
  const ZbCrcTables *tabs;
  uint32_t n;
  int data_format;             // requested (may be ZB_DF_DETECT)
  uint64_t pos;                // payload start for raw ZB_DF_DEFLATE members (zb200_inflate's `pos`)
  int count_only;
  const uint32_t *order;       // device [n] or null: the work queue hands out members in this order (longest first)
  const uint8_t *skip;         // device [n] or null: members with skip[i] != 0 are left alone
  const uint64_t *seg_bits;    // device [2n] or null (seg_mode only): bit-exact (start, end) of segment i inside src --
                               //   speculative segments found by zb_launch_find_blocks; segment i > 0 may reference
                               //   32768 bytes before its own start (its unknown window)
  uint64_t seg_limit;          // with seg_bits: byte offset in src where the stream's payload ends
  int mark;                    // with seg_bits: dst holds uint16 elements (dst_off in elements), 32768 marker
                               //   symbols sit in front of every segment's output
  int seg_mode;                // members are independently decodable SEGMENTS of one raw deflate stream:
                               // a segment also ends, successfully, when its input is used up at a block
                               // boundary; kind[i] reports whether a final block was seen
  // Gated queue (the host pipeline: ONE launch for a whole batch whose input is still arriving).  Queue
  // positions [gate_first[g], gate_first[g + 1]) belong to copy-in group g and are not started before
  // *gate_ready > g (written in stream order behind the group's copy); gate_done[g] counts the group's
  // finished members (a stream wait on it releases the group's copy-out).  Null: no gates.
  const uint32_t *gate_first;  // device [n_gates + 1]
  const uint32_t *gate_ready;  // device word
  uint32_t *gate_done;         // device [n_gates], zeroed before the launch
  uint32_t n_gates;
};
cudaError_t zb_launch_inflate(const ZbInflateWork &w, cudaStream_t s);
// positions just past every byte sequence 00 00 ff ff (the empty stored block that byte-aligns a
// stream: zlib's sync / full flush, and the joint between this library's 64 KiB chunks) inside
// src[lo, hi): unordered, *count may exceed cap (then the list is incomplete)
cudaError_t zb_launch_find_sync(const uint8_t *src, uint64_t lo, uint64_t hi, uint64_t *out, uint32_t cap,
                                uint32_t *count, cudaStream_t s);

// candidate starts (bit positions) of dynamic deflate blocks inside src bits [lo_bit, hi_bit): unordered,
// *count may exceed cap; bytes at and beyond limit_byte read as zero
cudaError_t zb_launch_find_blocks(const uint8_t *src, uint64_t lo_bit, uint64_t hi_bit, uint64_t limit_byte, uint64_t *out,
                                  uint32_t cap, uint32_t *count, cudaStream_t s);
// speculative segments decoded as uint16 symbols (ZbInflateWork::mark): marker prefill and resolution
struct ZbMarkSegHost {
  uint64_t scr;   // element offset of the segment's first output symbol in the scratch
  uint64_t dst;   // byte offset of the segment's first output byte in dst
  uint32_t n;     // output bytes
  uint32_t pad;
};
cudaError_t zb_launch_mark_prefill(uint16_t *scr, const void *segs, uint32_t nseg, cudaStream_t s);
cudaError_t zb_launch_resolve(const uint16_t *scr, const void *segs, uint32_t nseg, uint32_t max_n, uint8_t *dst, int *bad,
                              cudaStream_t s);

// ---- checksums over a batch of buffers (standalone crc32/adler32, and the trailer
// verification after inflate) ----
#define ZB_CK_PIECE_BYTES 32768   // the checksum kernels cut every buffer into pieces of this size
struct ZbPiece {
  uint64_t rel;   // start of the piece relative to its buffer
  uint32_t buf;   // buffer index
  uint32_t pad;
};
struct ZbChecksumWork {
  const uint8_t *src;          // device: base of the buffers
  const uint64_t *off;         // device [n+1]: buffer i starts at src + off[i]
  const uint64_t *lens;        // device [n] actual lengths, or null (= off[i+1]-off[i])
  const ZbPiece *pieces;       // device [n_pieces]: ZB_CK_PIECE_BYTES pieces covering every buffer's capacity
  const uint32_t *first;       // device [n+1]: first piece of each buffer
  ZbChunkCheck *piece_out;     // device [n_pieces] scratch
  uint32_t *partials;          // device [n_pieces * 512] scratch: the CRC path's per-warp, per-lane words of full pieces
  uint32_t *out;               // device [n] checksums, or null
  int *status;                 // device [n] or null: buffers with a non-zero status are skipped;
                               //   verify mode writes ZB_ERR_CHECKSUM / ZB_ERR_SIZE here
  const uint32_t *expect;      // device [n] expected value (verify mode) or null
  const uint32_t *kinds;       // device [n] resolved ZB_DF_* per buffer (gzip -> crc32, zlib -> adler32) or null
  const uint8_t *isize_src;    // verify mode: compressed buffers (gzip ISIZE lives in their last 4 bytes)
  const uint64_t *isize_off;   //   and their offsets [n+1]
  const ZbCrcTables *tabs;
  uint32_t n, n_pieces;
  int kind;                    // 0 crc32, 1 adler32 when kinds == null
  uint32_t big_pieces;         // 0, or: buffers of more pieces than this are folded by a whole CTA (k_buffer_combine_big)
                               //   instead of one warp (the host sets it when some buffer's capacity is that large)
};
#define ZB_CK_BIG_PIECES 2048u  // 64 MiB
#define ZB_CK_PARTIAL_BYTES 2048  // per piece in ZbChecksumWork::partials
cudaError_t zb_launch_checksum(const ZbChecksumWork &w, cudaStream_t s);
