// CPU build of the host/device-shared maths (zb_huff.h, zb_crc.h) for unit tests.
// The lane-strided CRC below is a lane-by-lane CPU emulation of warp_crc_raw() in
// zb_device.cuh so the algebra is checked without a GPU.
#include <string.h>
#include "../../zippy_b200/csrc/zb_crc.h"
#include "../../zippy_b200/csrc/zb_huff.h"

extern "C" {
void t_huff_lengths(const uint32_t *freq, int n, int limit, uint8_t *lens) { zb_huff_lengths(freq, n, limit, lens); }
void t_canonical(const uint8_t *lens, int n, uint32_t *out) { zb_canonical_codes(lens, n, out); }
int t_codebook_size() { return (int)sizeof(ZbCodebook); }
void t_build_codebook(const uint16_t *hist, uint32_t chunk_len, int is_final, int force, ZbCodebook *cb) {
  zb_build_codebook(hist, chunk_len, is_final, force, cb);
}
uint32_t t_gf2_mul(uint32_t a, uint32_t b) { return zb_gf2_mul(a, b); }
uint32_t t_xpow8(uint64_t n) { return zb_xpow8(n); }
uint32_t t_xpow8_t(uint64_t n) {
  static ZbCrcTables T;
  static int init = 0;
  if (!init) { zb_crc_build_tables(&T); init = 1; }
  return zb_xpow8_t(T.pow2, n);
}
uint32_t t_crc_combine(uint32_t a, uint32_t b, uint64_t lb) { return zb_crc32_combine(a, b, lb); }
uint32_t t_adler_combine(uint32_t a, uint32_t b, uint64_t lb) { return zb_adler32_combine(a, b, lb); }
int t_dist_code(uint32_t d) { return zb_dist_code(d); }
int t_len_code(uint32_t l) { return zb_len_code(l); }
uint32_t t_dist_base(int c) { return zb_dist_base(c); }
uint32_t t_len_base(int c) { return zb_len_base(c); }
int t_dist_extra(int c) { return zb_dist_extra_bits(c); }
int t_len_extra(int c) { return zb_len_extra_bits(c); }

static uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

uint32_t t_crc32_lane_model(const uint8_t *data, uint64_t n) {
  static ZbCrcTables T;
  static int init = 0;
  if (!init) { zb_crc_build_tables(&T); init = 1; }
  if (n == ZB_SUB_BYTES) {  // the device's fast path: four chains joined by quarter shifts
    const int QR = ZB_SUB_BYTES / 512;
    uint32_t total = 0;
    for (int lane = 0; lane < 32; lane++) {
      uint32_t r[4] = {0, 0, 0, 0};
      for (int k = 0; k < QR; k++)
        for (int c = 0; c < 4; c++) {
          uint32_t w = ld32(data + 128 * (QR * c + k) + 4 * lane);
          if (k) r[c] = T.mul1024[0][r[c] & 255] ^ T.mul1024[1][(r[c] >> 8) & 255] ^ T.mul1024[2][(r[c] >> 16) & 255] ^ T.mul1024[3][r[c] >> 24];
          r[c] ^= w;
        }
      uint32_t rr = zb_gf2_mul(r[0], T.quart_mul[3]) ^ zb_gf2_mul(r[1], T.quart_mul[2]) ^ zb_gf2_mul(r[2], T.quart_mul[1]) ^ r[3];
      total ^= zb_gf2_mul(rr, T.lane_mul[32 - lane]);
    }
    return zb_crc32_finalize(total, n);
  }
  uint64_t K = n / 128, t = n - 128 * K;
  uint32_t r_main = 0;
  for (int lane = 0; lane < 32; lane++) {
    uint32_t r = 0;
    for (uint64_t k = 0; k < K; k++) {
      uint32_t w = ld32(data + 128 * k + 4 * lane);
      if (k) r = T.mul1024[0][r & 255] ^ T.mul1024[1][(r >> 8) & 255] ^ T.mul1024[2][(r >> 16) & 255] ^ T.mul1024[3][r >> 24];
      r ^= w;
    }
    if (K) r_main ^= zb_gf2_mul(r, T.lane_mul[32 - lane]);
  }
  uint32_t tw = (uint32_t)(t / 4), rem = (uint32_t)(t % 4), r_tail = 0, r_rem = 0;
  for (uint32_t lane = 0; lane < tw; lane++) r_tail ^= zb_gf2_mul(ld32(data + 128 * K + 4 * lane), T.lane_mul[tw - lane]);
  for (uint32_t i = 0; i < rem; i++) r_rem = zb_crc_raw_byte(r_rem, data[128 * K + 4 * tw + i]);
  uint32_t raw = zb_gf2_mul(r_main, zb_xpow8(t)) ^ zb_gf2_mul(r_tail, zb_xpow8(rem)) ^ r_rem;
  return zb_crc32_finalize(raw, n);
}

// the checksum kernel's CRC path for a full 32 KiB piece (zb_inflate.cu: ck_warp_crc_rows + k_piece_fold): 64 chains
// of four rows with the 64-row step, four chains of a warp joined by table multiplications, then the 16 x 32
// partial words folded like a 2 KiB message
static uint32_t mul_tab(const uint32_t t[4][256], uint32_t r) {
  return t[0][r & 255] ^ t[1][(r >> 8) & 255] ^ t[2][(r >> 16) & 255] ^ t[3][r >> 24];
}
uint32_t t_crc32_piece_model(const uint8_t *data) {
  static ZbCrcTables T;
  static int init = 0;
  if (!init) { zb_crc_build_tables(&T); init = 1; }
  uint32_t total = 0;
  for (int lane = 0; lane < 32; lane++) {
    uint32_t f = 0;
    for (int warp = 0; warp < 16; warp++) {
      uint32_t r[4] = {0, 0, 0, 0};
      for (int k = 0; k < 4; k++)
        for (int c = 0; c < 4; c++) {
          if (k) r[c] = mul_tab(T.mul64r, r[c]);
          r[c] ^= ld32(data + 128 * (64 * k + warp + 16 * c) + 4 * lane);
        }
      const uint32_t m = mul_tab(T.ck_join[2], r[0]) ^ mul_tab(T.ck_join[1], r[1]) ^ mul_tab(T.ck_join[0], r[2]) ^ r[3];
      if (warp) f = mul_tab(T.mul1024, f);
      f ^= m;
    }
    total ^= zb_gf2_mul(f, T.lane_mul[32 - lane]);
  }
  return zb_crc32_finalize(total, 32768);
}

uint32_t t_adler32_model(const uint8_t *data, uint64_t n) {
  // per-4-byte-word sums exactly as the device does: A += sum(b), B += (n-o)*sum(b) - (b1+2*b2+3*b3)
  uint64_t A = 0, B = 0;
  uint64_t o = 0;
  for (; o + 4 <= n; o += 4) {
    uint32_t s = data[o] + data[o + 1] + data[o + 2] + data[o + 3];
    uint32_t wsum = data[o + 1] + 2 * data[o + 2] + 3 * data[o + 3];
    A += s;
    B += (n - o) * s - wsum;
  }
  for (; o < n; o++) { A += data[o]; B += (n - o) * data[o]; }
  return zb_adler_from_sums(A % ZB_ADLER_MOD, B % ZB_ADLER_MOD, n);
}
}
