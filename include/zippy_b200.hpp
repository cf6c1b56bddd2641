// zippy_b200.hpp -- header-only C++ mirror of guzba/zippy's public API over the C ABI
// (include/zippy_b200.h).  The reference is compiled code (Nim); with no Nim toolchain in
// this environment this is the compiled-language host side: same names, defaults, argument
// meaning and error behaviour as src/zippy.nim:11-177, src/zippy/common.nim:1-12,
// src/zippy/crc.nim:53-75, src/zippy/adler32.nim:6-66.  Framing for the single-input calls is
// done HERE, on the host, exactly where zippy.nim does it; the codec core (deflate, inflate,
// crc32, adler32) goes through the four seam entry points.
#pragma once
#include <cstdint>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "zippy_b200.h"

namespace zippy {

struct ZippyError : std::runtime_error {  // common.nim:2
  int code;
  ZippyError(int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

enum CompressedDataFormat { dfDetect = 0, dfZlib = 1, dfGzip = 2, dfDeflate = 3 };  // common.nim:4-5
constexpr int NoCompression = 0, BestSpeed = 1, BestCompression = 9, DefaultCompression = -1, HuffmanOnly = -2;

namespace detail {
inline void check(int rc) {
  if (rc != ZB200_OK) throw ZippyError(rc, zb200_strerror(rc));
}
inline zb200_ctx *ctx() {
  thread_local zb200_ctx *c = nullptr;
  if (!c) check(zb200_init(-1, &c));  // throws without a CUDA device: there is no CPU fallback
  return c;
}
inline const uint8_t *u8(const std::string &s) { return reinterpret_cast<const uint8_t *>(s.data()); }
// deflate.nim:207 -- appends the raw stream to dst
inline void deflate(std::string &dst, const uint8_t *src, size_t len, int level) {
  size_t start = dst.size(), n = 0;
  dst.resize(start + zb200_deflate_bound(len));
  check(zb200_deflate(ctx(), src, len, level, reinterpret_cast<uint8_t *>(&dst[start]), dst.size() - start, &n));
  dst.resize(start + n);
}
// inflate.nim:268
inline void inflate(std::string &dst, const uint8_t *src, size_t len, size_t pos) {
  // one decode: the library inflates into its own device memory, reports the size, then copies out
  size_t n = 0;
  check(zb200_decode_begin(ctx(), src, len, ZB200_DF_DEFLATE, pos, &n));
  dst.resize(n);
  uint8_t dummy = 0;
  check(zb200_decode_finish(ctx(), n ? reinterpret_cast<uint8_t *>(&dst[0]) : &dummy, n, &n));
  dst.resize(n);
}
inline uint32_t read32le(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
}  // namespace detail

inline uint32_t crc32(const void *src, size_t len) {  // crc.nim:53
  uint32_t v = 0;
  detail::check(zb200_crc32(detail::ctx(), src, len, &v));
  return v;
}
inline uint32_t crc32(const std::string &s) { return crc32(s.data(), s.size()); }
inline uint32_t adler32(const void *src, size_t len) {  // adler32.nim:6
  uint32_t v = 0;
  detail::check(zb200_adler32(detail::ctx(), src, len, &v));
  return v;
}
inline uint32_t adler32(const std::string &s) { return adler32(s.data(), s.size()); }

// zippy.nim:11-84
inline std::string compress(const void *srcp, size_t len, int level = DefaultCompression,
                            CompressedDataFormat dataFormat = dfGzip) {
  const uint8_t *src = static_cast<const uint8_t *>(srcp);
  std::string result;
  switch (dataFormat) {
    case dfGzip: {
      result.assign({31, (char)139, 8, 1 << 3, 0, 0, 0, 0, 0, 0});
      std::random_device rd;  // zippy.nim:28-42: 0..25 letters against BREACH-style length probing
      int k = (int)(rd() % 26);
      for (int i = 0; i < k; i++) result.push_back((char)(97 + i));
      result.push_back('\0');
      detail::deflate(result, src, len, level);
      uint32_t c = crc32(src, len), isz = (uint32_t)len;
      for (int s = 0; s < 32; s += 8) result.push_back((char)((c >> s) & 255));
      for (int s = 0; s < 32; s += 8) result.push_back((char)((isz >> s) & 255));
      return result;
    }
    case dfZlib: {
      result.assign({0x78, 0x01});
      detail::deflate(result, src, len, level);
      uint32_t a = adler32(src, len);
      for (int s = 24; s >= 0; s -= 8) result.push_back((char)((a >> s) & 255));
      return result;
    }
    case dfDeflate:
      detail::deflate(result, src, len, level);
      return result;
    default:
      throw ZippyError(ZB200_ERR_INVALID_FORMAT, "Invalid data format dfDetect");
  }
}
inline std::string compress(const std::string &src, int level = DefaultCompression,
                            CompressedDataFormat dataFormat = dfGzip) {
  return compress(src.data(), src.size(), level, dataFormat);
}

// gzip.nim:3-88
inline void uncompressGzip(std::string &dst, const uint8_t *src, size_t len) {
  auto fail = [] { throw ZippyError(ZB200_ERR_UNCOMPRESS, "Invalid buffer, unable to uncompress"); };
  if (len < 18) fail();
  if (src[0] != 31 || src[1] != 139) throw ZippyError(ZB200_ERR_GZIP_ID, "Failed gzip identification values check");
  if (src[2] != 8) throw ZippyError(ZB200_ERR_METHOD, "Unsupported compression method");
  uint8_t flg = src[3];
  if (flg & 0xe0) throw ZippyError(ZB200_ERR_GZIP_RESERVED, "Reserved flag bits set");
  if (flg & 4) throw ZippyError(ZB200_ERR_GZIP_FLAGS, "Currently unsupported flags are set");
  size_t pos = 10;
  for (int pass = 0; pass < 2; pass++)
    if ((pass == 0 && (flg & 8)) || (pass == 1 && (flg & 16))) {
      while (pos < len && src[pos] != 0) pos++;
      if (pos >= len) fail();
      pos++;
    }
  if (flg & 2) {
    if (pos + 2 >= len) fail();
    pos += 2;
  }
  if (pos + 8 >= len) fail();
  uint32_t checksum = detail::read32le(src + len - 8), isize = detail::read32le(src + len - 4);
  detail::inflate(dst, src, len, pos);
  if (checksum != crc32(dst)) throw ZippyError(ZB200_ERR_CHECKSUM, "Checksum verification failed");
  if (isize != (uint32_t)dst.size()) throw ZippyError(ZB200_ERR_SIZE, "Size verification failed");
}

// zippy.nim:100-165
inline std::string uncompress(const void *srcp, size_t len, CompressedDataFormat dataFormat = dfDetect) {
  const uint8_t *src = static_cast<const uint8_t *>(srcp);
  std::string result;
  switch (dataFormat) {
    case dfDetect:
      if (len > 18 && src[0] == 31 && src[1] == 139 && src[2] == 8 && (src[3] & 0xe0) == 0)
        return uncompress(src, len, dfGzip);
      if (len > 6 && (src[0] & 0x0f) == 8 && (src[0] >> 4) <= 7 && (((uint32_t)src[0] * 256u) + src[1]) % 31u == 0)
        return uncompress(src, len, dfZlib);
      throw ZippyError(ZB200_ERR_DETECT, "Unable to detect compressed data format");
    case dfGzip:
      uncompressGzip(result, src, len);
      return result;
    case dfZlib: {
      if (len < 6) throw ZippyError(ZB200_ERR_UNCOMPRESS, "Invalid buffer, unable to uncompress");
      uint8_t cmf = src[0], flg = src[1];
      if ((cmf & 0x0f) != 8) throw ZippyError(ZB200_ERR_METHOD, "Unsupported compression method");
      if ((cmf >> 4) > 7) throw ZippyError(ZB200_ERR_CINFO, "Invalid compression info");
      if ((((uint32_t)cmf * 256u) + flg) % 31u != 0) throw ZippyError(ZB200_ERR_HEADER, "Invalid header");
      if (flg & 0x20) throw ZippyError(ZB200_ERR_FDICT, "Preset dictionary is not yet supported");
      detail::inflate(result, src, len, 2);
      uint32_t checksum = ((uint32_t)src[len - 4] << 24) | ((uint32_t)src[len - 3] << 16) |
                          ((uint32_t)src[len - 2] << 8) | src[len - 1];
      if (checksum != adler32(result)) throw ZippyError(ZB200_ERR_CHECKSUM, "Checksum verification failed");
      return result;
    }
    case dfDeflate:
      detail::inflate(result, src, len, 0);
      return result;
  }
  throw ZippyError(ZB200_ERR_INVALID_FORMAT, "Invalid data format");
}
inline std::string uncompress(const std::string &src, CompressedDataFormat dataFormat = dfDetect) {
  return uncompress(src.data(), src.size(), dataFormat);
}

// One GPU launch sequence for many inputs (no reference counterpart; cf. the loop over entries
// in ziparchives.nim:505-540).
inline std::vector<std::string> compressBatch(const std::vector<std::string> &items, int level = DefaultCompression,
                                              CompressedDataFormat dataFormat = dfGzip) {
  std::string base;
  std::vector<uint64_t> offs(items.size() + 1, 0), out_offs(items.size() + 1, 0);
  size_t bound = 64;
  for (size_t i = 0; i < items.size(); i++) {
    base += items[i];
    offs[i + 1] = base.size();
    bound += zb200_compress_bound(items[i].size(), dataFormat) + 64;
  }
  std::string out(bound, '\0');
  std::vector<int> st(items.size() + 1, 0);
  uint8_t dummy = 0;
  detail::check(zb200_compress_batch(detail::ctx(), items.empty() ? &dummy : detail::u8(base), offs.data(), items.size(),
                                     level, dataFormat, nullptr, reinterpret_cast<uint8_t *>(&out[0]), out.size(),
                                     out_offs.data(), st.data()));
  std::vector<std::string> res;
  for (size_t i = 0; i < items.size(); i++) res.emplace_back(out.substr(out_offs[i], out_offs[i + 1] - out_offs[i]));
  return res;
}

}  // namespace zippy
