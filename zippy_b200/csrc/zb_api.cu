// zb_api.cu -- the C ABI (include/zippy_b200.h): context, device scratch, host<->device
// staging and the launch sequences.  No codec logic lives here and nothing here falls
// back to the CPU: every data byte is produced by the kernels in zb_deflate.cu /
// zb_inflate.cu.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/zippy_b200.h"
#include "zb_kernels.h"

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

constexpr size_t kMaxChunksPerGroup = 131072;  // 8 GiB of input per launch group

}  // namespace

struct zb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  ZbCrcTables *d_tabs = nullptr;
  DevBuf desc, member_first, fname, masks, recs, hist, chk, cb, chunk_off, member_off, member_check, member_isize;
  DevBuf src_off, dst_off, out_len, status, expect, kind, counter, ck_out;
  DevBuf in_stage, out_stage;
  cudaEvent_t ev[10];
  zb200_timing timing;
  std::string last_err;
  std::mutex mu;
};

namespace {

bool cuda_ok(zb200_ctx *c, cudaError_t e, const char *what) {
  if (e == cudaSuccess) return true;
  c->last_err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}
#define CK(call)                                      \
  do {                                                \
    if (!cuda_ok(ctx, (call), #call)) return ZB200_ERR_CUDA; \
  } while (0)

int ensure(zb200_ctx *ctx, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return ZB200_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    e = cudaMalloc(&b.p, bytes);
    want = bytes;
  }
  if (e != cudaSuccess) {
    ctx->last_err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    cudaGetLastError();
    return ZB200_ERR_NOMEM;
  }
  b.cap = want;
  return ZB200_OK;
}
#define ENSURE(buf, bytes)                         \
  do {                                             \
    int _rc = ensure(ctx, (buf), (bytes));         \
    if (_rc != ZB200_OK) return _rc;               \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};


float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

// ---- compress, device-resident src/dst ----
int compress_device_locked(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n, int level,
                           int data_format, const uint8_t *fname_lens, uint8_t *d_dst, size_t dst_cap,
                           uint64_t *dst_offsets, int *statuses) {
  if (level < -2 || level > 9) return ZB200_ERR_INVALID_LEVEL;
  if (data_format != ZB200_DF_GZIP && data_format != ZB200_DF_ZLIB && data_format != ZB200_DF_DEFLATE)
    return ZB200_ERR_INVALID_FORMAT;
  if (fname_lens)
    for (size_t i = 0; i < n; i++)
      if (fname_lens[i] > 25) return ZB200_ERR_ARG;
  if (((uintptr_t)d_dst & 3u) != 0) return ZB200_ERR_ARG;
  for (size_t i = 0; i < n; i++)
    if (statuses) statuses[i] = ZB200_OK;
  ctx->timing.lz_ms = ctx->timing.huff_ms = ctx->timing.scan_ms = ctx->timing.pack_ms = 0.f;
  ctx->timing.n_chunks = 0;
  dst_offsets[0] = 0;
  if (n == 0) return ZB200_OK;

  uint64_t out_base = 0;
  size_t m0 = 0;
  std::vector<ZbChunkDesc> desc;
  std::vector<uint32_t> first;
  while (m0 < n) {
    // ---- carve a group of members with a bounded number of chunks ----
    desc.clear();
    first.clear();
    size_t m1 = m0;
    uint64_t bound_total = 0;
    while (m1 < n) {
      uint64_t len = src_offsets[m1 + 1] - src_offsets[m1];
      size_t nc = len == 0 ? 1 : (size_t)((len + ZB_CHUNK_BYTES - 1) / ZB_CHUNK_BYTES);
      if (!desc.empty() && desc.size() + nc > kMaxChunksPerGroup) break;
      first.push_back((uint32_t)desc.size());
      for (size_t k = 0; k < nc; k++) {
        ZbChunkDesc d;
        d.src_off = src_offsets[m1] + (uint64_t)k * ZB_CHUNK_BYTES;
        d.len = (uint32_t)std::min<uint64_t>(ZB_CHUNK_BYTES, len - (uint64_t)k * ZB_CHUNK_BYTES);
        d.member = (uint32_t)(m1 - m0);
        d.flags = (k == 0 ? ZB_CHUNK_FIRST : 0u) | (k == nc - 1 ? ZB_CHUNK_LAST : 0u);
        d.pad = 0;
        desc.push_back(d);
      }
      bound_total += zb200_compress_bound((size_t)len, data_format) + 64;
      m1++;
    }
    first.push_back((uint32_t)desc.size());
    const size_t nm = m1 - m0, nc = desc.size();

    ENSURE(ctx->desc, nc * sizeof(ZbChunkDesc));
    ENSURE(ctx->member_first, (nm + 1) * sizeof(uint32_t));
    ENSURE(ctx->fname, nm + 16);
    ENSURE(ctx->masks, nc * ZB_WINDOWS_PER_CHUNK * sizeof(uint2));
    ENSURE(ctx->recs, nc * (size_t)ZB_WINDOWS_PER_CHUNK * ZB_MATCH_SLOTS * sizeof(uint32_t));
    ENSURE(ctx->hist, nc * (size_t)ZB_WARPS_PER_CHUNK * ZB_HIST_SYMS * sizeof(uint16_t));
    ENSURE(ctx->chk, nc * sizeof(ZbChunkCheck));
    ENSURE(ctx->cb, nc * sizeof(ZbCodebook));
    ENSURE(ctx->chunk_off, nc * sizeof(uint64_t));
    ENSURE(ctx->member_off, (nm + 1) * sizeof(uint64_t));
    ENSURE(ctx->member_check, nm * sizeof(uint32_t));
    ENSURE(ctx->member_isize, nm * sizeof(uint32_t));

    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->desc.p, desc.data(), nc * sizeof(ZbChunkDesc), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->member_first.p, first.data(), (nm + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    if (fname_lens && data_format == ZB200_DF_GZIP)
      CK(cudaMemcpyAsync(ctx->fname.p, fname_lens + m0, nm, cudaMemcpyHostToDevice, s));

    ZbCompressWork w;
    w.src = d_src;
    w.dst = d_dst;
    w.desc = (const ZbChunkDesc *)ctx->desc.p;
    w.member_first = (const uint32_t *)ctx->member_first.p;
    w.fname_len = (fname_lens && data_format == ZB200_DF_GZIP) ? (const uint8_t *)ctx->fname.p : nullptr;
    w.masks = (uint2 *)ctx->masks.p;
    w.recs = (uint32_t *)ctx->recs.p;
    w.hist = (uint16_t *)ctx->hist.p;
    w.chk = (ZbChunkCheck *)ctx->chk.p;
    w.cb = (ZbCodebook *)ctx->cb.p;
    w.chunk_off = (uint64_t *)ctx->chunk_off.p;
    w.member_off = (uint64_t *)ctx->member_off.p;
    w.member_check = (uint32_t *)ctx->member_check.p;
    w.member_isize = (uint32_t *)ctx->member_isize.p;
    w.tabs = ctx->d_tabs;
    w.n_chunks = (uint32_t)nc;
    w.n_members = (uint32_t)nm;
    w.level = level;
    w.data_format = data_format;
    w.out_base = out_base;

    CK(cudaEventRecord(ctx->ev[0], s));
    CK(zb_launch_lz(w, s));
    CK(cudaEventRecord(ctx->ev[1], s));
    CK(zb_launch_huff(w, s));
    CK(cudaEventRecord(ctx->ev[2], s));
    CK(zb_launch_scan(w, s));
    CK(cudaEventRecord(ctx->ev[3], s));
    // the packer ORs bits into a zero-filled stream
    uint64_t total_end = 0;
    bool known_fit = out_base + bound_total <= dst_cap;
    if (known_fit) {
      CK(cudaMemsetAsync(d_dst + out_base, 0, bound_total, s));
    } else {
      CK(cudaMemcpyAsync(&total_end, (uint64_t *)ctx->member_off.p + nm, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      if (total_end > dst_cap) return ZB200_ERR_DST_TOO_SMALL;
      size_t z0 = out_base & ~(size_t)3;
      size_t z1 = std::min<size_t>((total_end + 3) & ~(size_t)3, dst_cap);
      CK(cudaMemsetAsync(d_dst + z0, 0, z1 - z0, s));
    }
    CK(cudaEventRecord(ctx->ev[4], s));
    CK(zb_launch_pack(w, s));
    CK(cudaEventRecord(ctx->ev[5], s));
    CK(cudaMemcpyAsync(dst_offsets + m0, ctx->member_off.p, (nm + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    ctx->timing.lz_ms += ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->timing.huff_ms += ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->timing.scan_ms += ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->timing.pack_ms += ev_ms(ctx->ev[3], ctx->ev[5]);
    ctx->timing.kernel_launches += 4;
    ctx->timing.n_chunks += (uint32_t)nc;
    out_base = dst_offsets[m1];
    if (out_base > dst_cap) return ZB200_ERR_DST_TOO_SMALL;
    m0 = m1;
  }
  return ZB200_OK;
}

// ---- uncompress, device-resident ----
int uncompress_device_locked(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                             int data_format, uint64_t raw_pos, uint8_t *d_dst, const uint64_t *dst_offsets,
                             uint64_t *dst_lens, int *statuses, bool count_only) {
  if (data_format < ZB200_DF_DETECT || data_format > ZB200_DF_DEFLATE) return ZB200_ERR_INVALID_FORMAT;
  ctx->timing.inflate_ms = ctx->timing.verify_ms = 0.f;
  if (n == 0) return ZB200_OK;
  ENSURE(ctx->src_off, (n + 1) * sizeof(uint64_t));
  ENSURE(ctx->dst_off, (n + 1) * sizeof(uint64_t));
  ENSURE(ctx->out_len, n * sizeof(uint64_t));
  ENSURE(ctx->status, n * sizeof(int));
  ENSURE(ctx->expect, n * sizeof(uint32_t));
  ENSURE(ctx->kind, n * sizeof(uint32_t));
  ENSURE(ctx->counter, 64);
  cudaStream_t s = ctx->stream;
  CK(cudaMemcpyAsync(ctx->src_off.p, src_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  if (!count_only)
    CK(cudaMemcpyAsync(ctx->dst_off.p, dst_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  ZbInflateWork w;
  w.src = d_src;
  w.src_off = (const uint64_t *)ctx->src_off.p;
  w.dst = d_dst;
  w.dst_off = (const uint64_t *)ctx->dst_off.p;
  w.out_len = (uint64_t *)ctx->out_len.p;
  w.status = (int *)ctx->status.p;
  w.expect = (uint32_t *)ctx->expect.p;
  w.kind = (uint32_t *)ctx->kind.p;
  w.counter = (uint32_t *)ctx->counter.p;
  w.tabs = ctx->d_tabs;
  w.n = (uint32_t)n;
  w.data_format = data_format;
  w.pos = raw_pos;
  w.count_only = count_only ? 1 : 0;
  CK(cudaEventRecord(ctx->ev[0], s));
  CK(zb_launch_inflate(w, s));
  CK(cudaEventRecord(ctx->ev[1], s));
  CK(zb_launch_verify(w, s));
  CK(cudaEventRecord(ctx->ev[2], s));
  CK(cudaMemcpyAsync(dst_lens, ctx->out_len.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  std::vector<int> st_tmp;
  int *st = statuses;
  if (!st) {
    st_tmp.resize(n);
    st = st_tmp.data();
  }
  CK(cudaMemcpyAsync(st, ctx->status.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ctx->timing.inflate_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
  ctx->timing.verify_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
  ctx->timing.kernel_launches += count_only ? 1 : 2;
  for (size_t i = 0; i < n; i++)
    if (st[i] != ZB200_OK) dst_lens[i] = 0;
  return ZB200_OK;
}

int checksum_device_locked(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n, int kind,
                           uint32_t *out) {
  if (kind != 0 && kind != 1) return ZB200_ERR_ARG;
  if (n == 0) return ZB200_OK;
  ENSURE(ctx->src_off, (n + 1) * sizeof(uint64_t));
  ENSURE(ctx->ck_out, n * sizeof(uint32_t));
  cudaStream_t s = ctx->stream;
  CK(cudaMemcpyAsync(ctx->src_off.p, src_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  ZbChecksumWork w;
  w.src = d_src;
  w.off = (const uint64_t *)ctx->src_off.p;
  w.out = (uint32_t *)ctx->ck_out.p;
  w.tabs = ctx->d_tabs;
  w.n = (uint32_t)n;
  w.kind = kind;
  CK(cudaEventRecord(ctx->ev[0], s));
  CK(zb_launch_checksum(w, s));
  CK(cudaEventRecord(ctx->ev[1], s));
  CK(cudaMemcpyAsync(out, ctx->ck_out.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ctx->timing.checksum_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
  ctx->timing.kernel_launches += 1;
  return ZB200_OK;
}

// copy host inputs [src_offsets[0], src_offsets[n]) into the ctx staging buffer, rebased to 0
int stage_in(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
             std::vector<uint64_t> &rebased) {
  uint64_t lo = src_offsets[0], hi = src_offsets[n];
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] < src_offsets[i]) return ZB200_ERR_ARG;
  rebased.resize(n + 1);
  for (size_t i = 0; i <= n; i++) rebased[i] = src_offsets[i] - lo;
  ENSURE(ctx->in_stage, (size_t)(hi - lo) + 64);
  CK(cudaEventRecord(ctx->ev[6], ctx->stream));
  if (hi > lo)
    CK(cudaMemcpyAsync(ctx->in_stage.p, src_base + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->ev[7], ctx->stream));
  ctx->timing.h2d_bytes = hi - lo;
  return ZB200_OK;
}

}  // namespace

// =====================================================================================
extern "C" {

int zb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int zb200_init(int device, zb200_ctx **out) {
  if (!out) return ZB200_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return ZB200_ERR_CUDA;  // no CPU fallback
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return ZB200_ERR_CUDA;
  }
  if (device >= ndev) return ZB200_ERR_ARG;
  zb200_ctx *ctx = new zb200_ctx();
  ctx->device = device;
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  DeviceGuard g(device);
  bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  ctx->stream = ctx->own_stream;
  for (int i = 0; ok && i < 10; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  if (ok) ok = cudaMalloc((void **)&ctx->d_tabs, sizeof(ZbCrcTables)) == cudaSuccess;
  if (ok) {
    ZbCrcTables t;
    zb_crc_build_tables(&t);
    ok = cudaMemcpy(ctx->d_tabs, &t, sizeof(t), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (!ok) {
    cudaGetLastError();
    delete ctx;
    return ZB200_ERR_CUDA;
  }
  *out = ctx;
  return ZB200_OK;
}

void zb200_shutdown(zb200_ctx *ctx) {
  if (!ctx) return;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  DevBuf *bufs[] = {&ctx->desc, &ctx->member_first, &ctx->fname, &ctx->masks, &ctx->recs, &ctx->hist, &ctx->chk,
                    &ctx->cb, &ctx->chunk_off, &ctx->member_off, &ctx->member_check, &ctx->member_isize,
                    &ctx->src_off, &ctx->dst_off, &ctx->out_len, &ctx->status, &ctx->expect, &ctx->kind,
                    &ctx->counter, &ctx->ck_out, &ctx->in_stage, &ctx->out_stage};
  for (DevBuf *b : bufs)
    if (b->p) cudaFree(b->p);
  if (ctx->d_tabs) cudaFree(ctx->d_tabs);
  for (int i = 0; i < 10; i++) cudaEventDestroy(ctx->ev[i]);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int zb200_set_stream(zb200_ctx *ctx, void *cuda_stream) {
  if (!ctx) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return ZB200_OK;
}

const char *zb200_last_cuda_error(zb200_ctx *ctx) { return ctx ? ctx->last_err.c_str() : ""; }

const char *zb200_strerror(int s) {
  switch (s) {
    case ZB200_OK: return "ok";
    case ZB200_ERR_INVALID_LEVEL: return "Invalid compression level";
    case ZB200_ERR_INVALID_FORMAT: return "Invalid data format";
    case ZB200_ERR_UNCOMPRESS: return "Invalid buffer, unable to uncompress";
    case ZB200_ERR_COMPRESS: return "Unexpected error while compressing";
    case ZB200_ERR_END_OF_BUFFER: return "Cannot read further, at end of buffer";
    case ZB200_ERR_BYTE_BOUNDARY: return "Must be at a byte boundary";
    case ZB200_ERR_BLOCK_HEADER: return "Invalid block header";
    case ZB200_ERR_INVALID_SYMBOL: return "Invalid symbol";
    case ZB200_ERR_DETECT: return "Unable to detect compressed data format";
    case ZB200_ERR_METHOD: return "Unsupported compression method";
    case ZB200_ERR_CINFO: return "Invalid compression info";
    case ZB200_ERR_HEADER: return "Invalid header";
    case ZB200_ERR_FDICT: return "Preset dictionary is not yet supported";
    case ZB200_ERR_CHECKSUM: return "Checksum verification failed";
    case ZB200_ERR_GZIP_ID: return "Failed gzip identification values check";
    case ZB200_ERR_GZIP_RESERVED: return "Reserved flag bits set";
    case ZB200_ERR_GZIP_FLAGS: return "Currently unsupported flags are set";
    case ZB200_ERR_SIZE: return "Size verification failed";
    case ZB200_ERR_DST_TOO_SMALL: return "Destination buffer too small";
    case ZB200_ERR_CUDA: return "CUDA error (no CPU fallback)";
    case ZB200_ERR_NOMEM: return "Out of device memory";
    case ZB200_ERR_ARG: return "Invalid argument";
    default: return "unknown status";
  }
}

size_t zb200_deflate_bound(size_t len) {
  size_t chunks = len == 0 ? 1 : (len + ZB_CHUNK_BYTES - 1) / ZB_CHUNK_BYTES;
  // per chunk: worst case is the stored path (two stored pieces for a full 64 KiB chunk);
  // a coded chunk is only chosen when it is smaller than that.
  return len + chunks * 10 + 8;
}
size_t zb200_compress_bound(size_t len, int data_format) {
  size_t frame = data_format == ZB200_DF_GZIP ? 36 + 8 : data_format == ZB200_DF_ZLIB ? 6 : 0;
  return zb200_deflate_bound(len) + frame;
}

int zb200_compress_batch_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                                int level, int data_format, const uint8_t *fname_lens, uint8_t *d_dst,
                                size_t dst_cap, uint64_t *dst_offsets, int *statuses) {
  if (!ctx || !src_offsets || !dst_offsets || (n && (!d_src || !d_dst))) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  ctx->timing.kernel_launches = 0;
  return compress_device_locked(ctx, d_src, src_offsets, n, level, data_format, fname_lens, d_dst, dst_cap,
                                dst_offsets, statuses);
}

int zb200_compress_batch(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n, int level,
                         int data_format, const uint8_t *fname_lens, uint8_t *dst_base, size_t dst_cap,
                         uint64_t *dst_offsets, int *statuses) {
  if (!ctx || !src_offsets || !dst_offsets || (n && (!src_base || !dst_base))) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  std::vector<uint64_t> reb;
  int rc = stage_in(ctx, src_base, src_offsets, n, reb);
  if (rc) return rc;
  uint64_t bound = 0;
  for (size_t i = 0; i < n; i++) bound += zb200_compress_bound((size_t)(reb[i + 1] - reb[i]), data_format) + 64;
  ENSURE(ctx->out_stage, (size_t)bound + 64);
  rc = compress_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data(), n, level, data_format, fname_lens,
                              (uint8_t *)ctx->out_stage.p, ctx->out_stage.cap & ~(size_t)3, dst_offsets, statuses);
  if (rc) return rc;
  uint64_t total = dst_offsets[n];
  if (total > dst_cap) return ZB200_ERR_DST_TOO_SMALL;
  CK(cudaEventRecord(ctx->ev[8], ctx->stream));
  if (total) CK(cudaMemcpyAsync(dst_base, ctx->out_stage.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(ctx->ev[9], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.h2d_ms = ev_ms(ctx->ev[6], ctx->ev[7]);
  ctx->timing.d2h_ms = ev_ms(ctx->ev[8], ctx->ev[9]);
  ctx->timing.d2h_bytes = total;
  return ZB200_OK;
}

int zb200_uncompress_batch_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                                  int data_format, uint8_t *d_dst, const uint64_t *dst_offsets, uint64_t *dst_lens,
                                  int *statuses) {
  if (!ctx || !src_offsets || !dst_offsets || !dst_lens || (n && !d_src)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  ctx->timing.kernel_launches = 0;
  return uncompress_device_locked(ctx, d_src, src_offsets, n, data_format, 0, d_dst, dst_offsets, dst_lens, statuses,
                                  false);
}

int zb200_uncompress_sizes_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                                  int data_format, uint64_t *sizes, int *statuses) {
  if (!ctx || !src_offsets || !sizes || (n && !d_src)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  ctx->timing.kernel_launches = 0;
  return uncompress_device_locked(ctx, d_src, src_offsets, n, data_format, 0, nullptr, nullptr, sizes, statuses,
                                  true);
}

int zb200_uncompress_sizes(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                           int data_format, uint64_t *sizes, int *statuses) {
  if (!ctx || !src_offsets || !sizes || (n && !src_base)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  std::vector<uint64_t> reb;
  int rc = stage_in(ctx, src_base, src_offsets, n, reb);
  if (rc) return rc;
  return uncompress_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data(), n, data_format, 0, nullptr,
                                  nullptr, sizes, statuses, true);
}

int zb200_uncompress_batch(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                           int data_format, uint8_t *dst_base, const uint64_t *dst_offsets, uint64_t *dst_lens,
                           int *statuses) {
  if (!ctx || !src_offsets || !dst_offsets || !dst_lens || (n && !src_base)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  std::vector<uint64_t> reb;
  int rc = stage_in(ctx, src_base, src_offsets, n, reb);
  if (rc) return rc;
  for (size_t i = 0; i < n; i++)
    if (dst_offsets[i + 1] < dst_offsets[i]) return ZB200_ERR_ARG;
  uint64_t lo = n ? dst_offsets[0] : 0, hi = n ? dst_offsets[n] : 0;
  std::vector<uint64_t> dreb(n + 1);
  for (size_t i = 0; i <= n; i++) dreb[i] = dst_offsets[i] - lo;
  ENSURE(ctx->out_stage, (size_t)(hi - lo) + 64);
  rc = uncompress_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data(), n, data_format, 0,
                                (uint8_t *)ctx->out_stage.p, dreb.data(), dst_lens, statuses, false);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev[8], ctx->stream));
  if (hi > lo && dst_base)
    CK(cudaMemcpyAsync(dst_base + lo, ctx->out_stage.p, (size_t)(hi - lo), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(ctx->ev[9], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->timing.h2d_ms = ev_ms(ctx->ev[6], ctx->ev[7]);
  ctx->timing.d2h_ms = ev_ms(ctx->ev[8], ctx->ev[9]);
  ctx->timing.d2h_bytes = hi - lo;
  return ZB200_OK;
}

int zb200_checksum_batch_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n, int kind,
                                uint32_t *out) {
  if (!ctx || !src_offsets || !out || (n && !d_src)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  ctx->timing.kernel_launches = 0;
  return checksum_device_locked(ctx, d_src, src_offsets, n, kind, out);
}

int zb200_checksum_batch(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n, int kind,
                         uint32_t *out) {
  if (!ctx || !src_offsets || !out || (n && !src_base)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  std::vector<uint64_t> reb;
  int rc = stage_in(ctx, src_base, src_offsets, n, reb);
  if (rc) return rc;
  return checksum_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data(), n, kind, out);
}

// ---- the single-input seam ----
int zb200_deflate(zb200_ctx *ctx, const uint8_t *src, size_t len, int level, uint8_t *dst, size_t dst_cap,
                  size_t *dst_len) {
  if (!dst_len) return ZB200_ERR_ARG;
  uint64_t so[2] = {0, len}, dof[2] = {0, 0};
  int st = 0;
  uint8_t dummy = 0;
  int rc = zb200_compress_batch(ctx, src ? src : &dummy, so, 1, level, ZB200_DF_DEFLATE, nullptr, dst, dst_cap, dof,
                                &st);
  if (rc) return rc;
  *dst_len = (size_t)dof[1];
  return st;
}

static int inflate_one(zb200_ctx *ctx, const uint8_t *src, size_t len, size_t pos, uint8_t *dst, size_t dst_cap,
                       size_t *dst_len, bool count_only) {
  if (!ctx || !dst_len || (len && !src)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  uint64_t so[2] = {0, len};
  std::vector<uint64_t> reb;
  uint8_t dummy = 0;
  int rc = stage_in(ctx, src ? src : &dummy, so, 1, reb);
  if (rc) return rc;
  uint64_t dof[2] = {0, dst_cap}, dl = 0;
  int st = 0;
  if (!count_only) ENSURE(ctx->out_stage, dst_cap + 64);
  rc = uncompress_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data(), 1, ZB200_DF_DEFLATE, pos,
                                count_only ? nullptr : (uint8_t *)ctx->out_stage.p, dof, &dl, &st, count_only);
  if (rc) return rc;
  if (st) return st;
  if (!count_only && dl) {
    CK(cudaMemcpyAsync(dst, ctx->out_stage.p, (size_t)dl, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  *dst_len = (size_t)dl;
  return ZB200_OK;
}

int zb200_inflate(zb200_ctx *ctx, const uint8_t *src, size_t len, size_t pos, uint8_t *dst, size_t dst_cap,
                  size_t *dst_len) {
  if (dst_cap && !dst) return ZB200_ERR_ARG;
  return inflate_one(ctx, src, len, pos, dst, dst_cap, dst_len, false);
}
int zb200_inflate_size(zb200_ctx *ctx, const uint8_t *src, size_t len, size_t pos, size_t *out_len) {
  return inflate_one(ctx, src, len, pos, nullptr, 0, out_len, true);
}

int zb200_crc32(zb200_ctx *ctx, const void *src, size_t len, uint32_t *out) {
  uint64_t so[2] = {0, len};
  uint8_t dummy = 0;
  return zb200_checksum_batch(ctx, src ? (const uint8_t *)src : &dummy, so, 1, 0, out);
}
int zb200_adler32(zb200_ctx *ctx, const void *src, size_t len, uint32_t *out) {
  uint64_t so[2] = {0, len};
  uint8_t dummy = 0;
  return zb200_checksum_batch(ctx, src ? (const uint8_t *)src : &dummy, so, 1, 1, out);
}

int zb200_last_timing(zb200_ctx *ctx, zb200_timing *out) {
  if (!ctx || !out) return ZB200_ERR_ARG;
  *out = ctx->timing;
  return ZB200_OK;
}

}  // extern "C"
