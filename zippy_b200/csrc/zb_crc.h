// zb_crc.h -- GF(2) helpers for the blocked combine-reduce CRC-32 and the closed-form
// Adler-32 combine (host + device).
//
// Replaces the reference's serial CRC (src/zippy/crc.nim:29-72 slice-by-8 +
// crc32_simd.nim:39-144 PCLMUL folding) and Adler (adler32.nim:6-63) with forms that
// split over lanes / warps / chunks and recombine exactly:
//   raw(A||B)   = raw(A) * x^(8|B|) + raw(B)            (raw = init 0, no final xor)
//   crc32(M)    = ~( 0xffffffff * x^(8|M|) + raw(M) )
//   adler(A||B) : s1 = s1A + s1B - 1, s2 = s2A + s2B + |B| * (s1A - 1)   (mod 65521)
// Register convention is the usual reflected one (poly 0xEDB88320; bit 31 holds x^0),
// the same convention zlib's crc32_combine uses.
#pragma once
#include "zb_common.h"

#define ZB_CRC_POLY 0xedb88320u
#define ZB_ADLER_MOD 65521u

// a(x) * b(x) mod P(x)
ZB_HD uint32_t zb_gf2_mul(uint32_t a, uint32_t b) {
  uint32_t p = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
  for (int i = 0; i < 32; i++) {
    p ^= (0u - ((a >> 31) & 1u)) & b;
    a <<= 1;
    b = (b >> 1) ^ ((0u - (b & 1u)) & ZB_CRC_POLY);
  }
  return p;
}

// x^(8*nbytes) mod P by square-and-multiply.
ZB_HD uint32_t zb_xpow8(uint64_t nbytes) {
  uint32_t result = 0x80000000u;  // x^0
  uint32_t base = 0x00800000u;    // x^8
  while (nbytes) {
    if (nbytes & 1) result = zb_gf2_mul(result, base);
    nbytes >>= 1;
    if (nbytes) base = zb_gf2_mul(base, base);
  }
  return result;
}

// The same with the squarings read from a table (ZbCrcTables::pow2): one multiplication per set bit of nbytes
// beyond the first (zb_xpow8 spends 2 log2 n of them; a warp per 64 KiB buffer did little else).
ZB_HD uint32_t zb_xpow8_t(const uint32_t *pow2, uint64_t nbytes) {
  uint32_t result = 0x80000000u;  // x^0
  bool first = true;
  for (int j = 0; nbytes; j++, nbytes >>= 1)
    if (nbytes & 1) {
      result = first ? pow2[j] : zb_gf2_mul(result, pow2[j]);
      first = false;
    }
  return result;
}

// raw CRC (init 0) of the single byte b followed by nothing: b(x) * x^8 mod P,
// computed bitwise (no table) -- used only for <4-byte tails.
ZB_HD uint32_t zb_crc_raw_byte(uint32_t state, uint32_t b) {
  state ^= b;
  for (int i = 0; i < 8; i++) state = (state >> 1) ^ ((0u - (state & 1u)) & ZB_CRC_POLY);
  return state;
}

// Combine finalized CRCs: crc(A||B) from crc(A), crc(B), |B|.
ZB_HD uint32_t zb_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) {
  return zb_gf2_mul(zb_xpow8(len_b), crc_a) ^ crc_b;
}

// Finalize a raw (init-0) CRC of an n-byte message.
ZB_HD uint32_t zb_crc32_finalize(uint32_t raw, uint64_t n) {
  return ~(zb_gf2_mul(zb_xpow8(n), 0xffffffffu) ^ raw);
}

// Adler-32 of a piece, given A = sum(b_i) and B = sum((n - i) * b_i), i 0-based,
// both already reduced mod 65521 (or small enough not to overflow).
ZB_HD uint32_t zb_adler_from_sums(uint64_t a_sum, uint64_t b_sum, uint64_t n) {
  uint32_t s1 = (uint32_t)((1 + a_sum) % ZB_ADLER_MOD);
  uint32_t s2 = (uint32_t)((n % ZB_ADLER_MOD + b_sum) % ZB_ADLER_MOD);
  return (s2 << 16) | s1;
}

ZB_HD uint32_t zb_adler32_combine(uint32_t ad_a, uint32_t ad_b, uint64_t len_b) {
  uint32_t s1a = ad_a & 0xffffu, s2a = ad_a >> 16, s1b = ad_b & 0xffffu, s2b = ad_b >> 16;
  uint32_t rem = (uint32_t)(len_b % ZB_ADLER_MOD);
  uint32_t s1 = (s1a + s1b + ZB_ADLER_MOD - 1) % ZB_ADLER_MOD;
  uint64_t s2 = (uint64_t)s2a + s2b + (uint64_t)rem * ((s1a + ZB_ADLER_MOD - 1) % ZB_ADLER_MOD);
  return ((uint32_t)(s2 % ZB_ADLER_MOD) << 16) | s1;
}

// Tables for the lane-strided CRC kernel body (built once on the host, copied to the
// device).  mul1024[k][b] = ((b << 8k) as a register value) * x^1024 mod P, i.e. the
// state after the byte-k contribution is followed by 128 more message bytes.
// lane_mul[j] = x^(32*j) mod P for j = 0..32.
struct ZbCrcTables {
  uint32_t mul1024[4][256];
  uint32_t lane_mul[33];
  uint32_t sub_mul[8];  // [k] = x^(8 * 8192 * k) mod P, k = 1..7: shifts a sub-chunk's CRC to the chunk end;
                        // [0] = x^(8 * 65536) mod P: shifts by one whole chunk
  uint32_t quart_mul[4];  // [k] = x^(8 * 2048 * k) mod P: shifts a quarter of a sub-chunk (16 rows of 128 B)
  uint32_t ck_sub[16];    // [k] = x^(8 * 2048 * k) mod P: shifts a warp's 2 KiB of a (ragged) checksum piece to the piece end
  uint32_t pow2[48];      // [j] = x^(8 * 2^j) mod P: x^(8 n) as a product over the set bits of n (zb_xpow8_t)
  uint32_t piece_mul[16]; // [k] = x^(8 * 4096 * k) mod P: shifts the CRC of one of k_lz's 4 KiB pieces to the chunk end
  uint32_t pq_mul[4];     // [k] = x^(8 * 1024 * k) mod P: joins the four 1 KiB chains of such a piece
  // the checksum kernel's CRC path (zb_inflate.cu): a 32 KiB piece is 256 rows of 128 B; row q + 64 k belongs to chain q
  uint32_t mul64r[4][256];       // [j][b] = (b << 8j) * x^(8 * 128 * 64): the Horner step of a chain (64 rows)
  uint32_t ck_join[3][4][256];   // [k-1][j][b] = (b << 8j) * x^(8 * 128 * 16 * k), k = 1..3: joins a warp's four chains
};

inline void zb_crc_build_tables(ZbCrcTables *t) {
  uint32_t x1024 = zb_xpow8(128);
  for (int k = 0; k < 4; k++)
    for (uint32_t b = 0; b < 256; b++) t->mul1024[k][b] = zb_gf2_mul(b << (8 * k), x1024);
  for (int j = 0; j <= 32; j++) t->lane_mul[j] = zb_xpow8(4ull * (uint64_t)j);
  for (int k = 1; k < 8; k++) t->sub_mul[k] = zb_xpow8((uint64_t)ZB_SUB_BYTES * (uint64_t)k);
  t->sub_mul[0] = zb_xpow8((uint64_t)ZB_CHUNK_BYTES);
  for (int k = 0; k < 4; k++) t->quart_mul[k] = zb_xpow8((uint64_t)(ZB_SUB_BYTES / 4) * (uint64_t)k);
  for (int k = 0; k < 16; k++) t->ck_sub[k] = zb_xpow8(2048ull * (uint64_t)k);
  {
    const uint32_t c = zb_xpow8(128ull * 64ull);
    for (int j = 0; j < 4; j++)
      for (uint32_t b = 0; b < 256; b++) t->mul64r[j][b] = zb_gf2_mul(b << (8 * j), c);
  }
  t->pow2[0] = 0x00800000u;
  for (int j = 1; j < 48; j++) t->pow2[j] = zb_gf2_mul(t->pow2[j - 1], t->pow2[j - 1]);
  for (int k = 0; k < 16; k++) t->piece_mul[k] = zb_xpow8(4096ull * (uint64_t)k);
  for (int k = 0; k < 4; k++) t->pq_mul[k] = zb_xpow8(1024ull * (uint64_t)k);
  for (int k = 1; k <= 3; k++) {
    const uint32_t c = zb_xpow8(128ull * 16ull * (uint64_t)k);
    for (int j = 0; j < 4; j++)
      for (uint32_t b = 0; b < 256; b++) t->ck_join[k - 1][j][b] = zb_gf2_mul(b << (8 * j), c);
  }
}
