// TEST SCAFFOLDING: a zlib-backed stand-in for the few C-ABI entry points include/zippy_b200_zip.hpp
// calls, so that the C++ container logic can be exercised on a machine without a GPU
// (tests/test_ziparchives.py).  It is never linked into the product.
#include <zlib.h>

#include <cstring>
#include <vector>

#include "../../include/zippy_b200.h"

struct zb200_ctx {
  int unused;
};

extern "C" {
int zb200_init(int, zb200_ctx **out) {
  static zb200_ctx c;
  *out = &c;
  return ZB200_OK;
}
const char *zb200_strerror(int status) { return status == ZB200_OK ? "ok" : "mock error"; }
size_t zb200_compress_bound(size_t len, int) { return compressBound((uLong)len) + 64; }
int zb200_checksum_batch(zb200_ctx *, const uint8_t *base, const uint64_t *off, size_t n, int kind, uint32_t *out) {
  for (size_t i = 0; i < n; i++) {
    const uInt l = (uInt)(off[i + 1] - off[i]);
    out[i] = kind == 0 ? (uint32_t)crc32(0L, base + off[i], l) : (uint32_t)adler32(1L, base + off[i], l);
  }
  return ZB200_OK;
}
int zb200_crc32(zb200_ctx *, const void *src, size_t len, uint32_t *out) {
  *out = (uint32_t)crc32(0L, static_cast<const Bytef *>(src), (uInt)len);
  return ZB200_OK;
}
int zb200_adler32(zb200_ctx *, const void *src, size_t len, uint32_t *out) {
  *out = (uint32_t)adler32(1L, static_cast<const Bytef *>(src), (uInt)len);
  return ZB200_OK;
}
static int raw_inflate(const uint8_t *src, size_t len, size_t pos, std::vector<uint8_t> &out) {
  z_stream zs;
  std::memset(&zs, 0, sizeof(zs));
  if (pos > len || inflateInit2(&zs, -15) != Z_OK) return ZB200_ERR_UNCOMPRESS;
  zs.next_in = const_cast<Bytef *>(src + pos);
  zs.avail_in = (uInt)(len - pos);
  out.resize(1 << 16);
  int rc;
  for (;;) {
    zs.next_out = out.data() + zs.total_out;
    zs.avail_out = (uInt)(out.size() - zs.total_out);
    rc = inflate(&zs, Z_NO_FLUSH);
    if (rc != Z_OK || zs.avail_out != 0) break;
    out.resize(out.size() * 2);
  }
  out.resize(zs.total_out);
  inflateEnd(&zs);
  return rc == Z_STREAM_END ? ZB200_OK : ZB200_ERR_UNCOMPRESS;
}
int zb200_inflate_size(zb200_ctx *, const uint8_t *src, size_t len, size_t pos, size_t *n) {
  std::vector<uint8_t> out;
  int rc = raw_inflate(src, len, pos, out);
  *n = out.size();
  return rc;
}
int zb200_inflate(zb200_ctx *, const uint8_t *src, size_t len, size_t pos, uint8_t *dst, size_t cap, size_t *n) {
  std::vector<uint8_t> out;
  int rc = raw_inflate(src, len, pos, out);
  if (rc) return rc;
  if (out.size() > cap) return ZB200_ERR_DST_TOO_SMALL;
  if (!out.empty()) std::memcpy(dst, out.data(), out.size());
  *n = out.size();
  return ZB200_OK;
}
static std::vector<uint8_t> g_pending;
int zb200_decode_begin(zb200_ctx *, const uint8_t *src, size_t len, int fmt, size_t pos, size_t *n) {
  if (fmt != ZB200_DF_DEFLATE) return ZB200_ERR_INVALID_FORMAT;  // the mock only serves the raw seam
  int rc = raw_inflate(src, len, pos, g_pending);
  *n = g_pending.size();
  return rc;
}
int zb200_decode_finish(zb200_ctx *, uint8_t *dst, size_t cap, size_t *n) {
  if (g_pending.size() > cap) return ZB200_ERR_DST_TOO_SMALL;
  if (!g_pending.empty()) std::memcpy(dst, g_pending.data(), g_pending.size());
  *n = g_pending.size();
  return ZB200_OK;
}
int zb200_compress_batch(zb200_ctx *, const uint8_t *base, const uint64_t *off, size_t n, int, int fmt, const uint8_t *,
                         uint8_t *dst, size_t cap, uint64_t *dst_off, int *st) {
  if (fmt != ZB200_DF_DEFLATE) return ZB200_ERR_INVALID_FORMAT;
  size_t pos = 0;
  dst_off[0] = 0;
  for (size_t i = 0; i < n; i++) {
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return ZB200_ERR_COMPRESS;
    zs.next_in = const_cast<Bytef *>(base + off[i]);
    zs.avail_in = (uInt)(off[i + 1] - off[i]);
    zs.next_out = dst + pos;
    zs.avail_out = (uInt)(cap - pos);
    const int rc = deflate(&zs, Z_FINISH);
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return ZB200_ERR_DST_TOO_SMALL;
    pos += zs.total_out;
    dst_off[i + 1] = pos;
    if (st) st[i] = ZB200_OK;
  }
  return ZB200_OK;
}
int zb200_uncompress_batch(zb200_ctx *, const uint8_t *base, const uint64_t *off, size_t n, int fmt, uint8_t *dst,
                           const uint64_t *dst_off, uint64_t *lens, int *st) {
  if (fmt != ZB200_DF_DEFLATE) return ZB200_ERR_INVALID_FORMAT;
  for (size_t i = 0; i < n; i++) {
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return ZB200_ERR_UNCOMPRESS;
    zs.next_in = const_cast<Bytef *>(base + off[i]);
    zs.avail_in = (uInt)(off[i + 1] - off[i]);
    zs.next_out = dst + dst_off[i];
    zs.avail_out = (uInt)(dst_off[i + 1] - dst_off[i]);
    const int rc = inflate(&zs, Z_FINISH);
    lens[i] = rc == Z_STREAM_END ? zs.total_out : 0;
    if (st) st[i] = rc == Z_STREAM_END ? ZB200_OK : ZB200_ERR_UNCOMPRESS;
    inflateEnd(&zs);
  }
  return ZB200_OK;
}
}
