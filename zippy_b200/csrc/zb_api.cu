// zb_api.cu -- the C ABI (include/zippy_b200.h): context, device scratch, host<->device
// staging and the launch sequences.  No codec logic lives here and nothing here falls
// back to the CPU: every data byte is produced by the kernels in zb_deflate.cu /
// zb_inflate.cu.
#include <cuda.h>
#include <cuda_runtime.h>
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/zippy_b200.h"
#include "zb_kernels.h"
#include "zb_wrapper.h"

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

// Stream memory operations (driver API, fetched at run time: the library links the runtime only).  A wait on a
// device word lets the D2H stream follow the progress of ONE persistent inflate launch, group by group.
struct StreamMemOps {
  CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  CUresult (*write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  bool ok = false;
};
StreamMemOps load_stream_memops() {
  StreamMemOps m;
  void *f1 = nullptr, *f2 = nullptr;
  cudaDriverEntryPointQueryResult q1, q2;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f1, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess &&
      cudaGetDriverEntryPoint("cuStreamWriteValue32", &f2, cudaEnableDefault, &q2) == cudaSuccess && q2 == cudaDriverEntryPointSuccess &&
      f1 && f2) {
    m.wait32 = reinterpret_cast<decltype(m.wait32)>(f1);
    m.write32 = reinterpret_cast<decltype(m.write32)>(f2);
    m.ok = true;
  } else {
    (void)cudaGetLastError();
  }
  return m;
}

constexpr size_t kMaxChunksPerGroup = 32768;  // device-resident batches: 2 GiB of input per launch group
constexpr size_t kHostGroupChunks = 4096;     // host batches: 256 MiB groups so transfers overlap the kernels

}  // namespace

// ---- pageable host memory ----
// cudaMemcpyAsync only overlaps with kernels for page-locked memory; a caller's ordinary buffer (a Nim
// string, malloc, numpy) would make every copy block.  Such buffers are staged through a ring of pinned
// slots: a few host threads memcpy a slice into a slot while the DMA engine drains the previous one.
class CopyPool {
 public:
  ~CopyPool() { stop(); }
  void start(int n) {
    if (!th_.empty()) return;
    for (int i = 0; i < n; i++) th_.emplace_back([this] { work(); });
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
    }
    cv_.notify_all();
    for (std::thread &t : th_) t.join();
    th_.clear();
  }
  // copy n bytes with all workers (plus the caller); returns when done
  void copy(uint8_t *dst, const uint8_t *src, size_t n) {
    const size_t piece = 1 << 20;
    if (th_.empty() || n < 2 * piece) {
      memcpy(dst, src, n);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      dst_ = dst;
      src_ = src;
      n_ = n;
      next_ = 0;
      left_ = (n + piece - 1) / piece;
    }
    cv_.notify_all();
    help(piece);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return left_ == 0; });
  }

 private:
  bool take(size_t piece, size_t &off, size_t &len) {
    std::lock_guard<std::mutex> lk(m_);
    if (next_ >= n_) return false;
    off = next_;
    len = std::min(piece, n_ - next_);
    next_ += len;
    return true;
  }
  void finish_one() {
    std::lock_guard<std::mutex> lk(m_);
    if (--left_ == 0) done_.notify_all();
  }
  void help(size_t piece) {
    size_t off, len;
    while (take(piece, off, len)) {
      memcpy(dst_ + off, src_ + off, len);
      finish_one();
    }
  }
  void work() {
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return quit_ || next_ < n_; });
        if (quit_) return;
      }
      help(1 << 20);
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  uint8_t *dst_ = nullptr;
  const uint8_t *src_ = nullptr;
  size_t n_ = 0, next_ = 0, left_ = 0;
  bool quit_ = false;
};

constexpr size_t kStageSlotBytes = 32u << 20;
constexpr int kStageSlots = 4;
struct StageRing {
  uint8_t *slot[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  bool busy[kStageSlots] = {false, false, false, false};
  uint8_t *out_dst[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};  // D2H ring: where the slot's bytes go
  size_t out_n[kStageSlots] = {0, 0, 0, 0};
  size_t next = 0;
};

struct zb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  ZbCrcTables *d_tabs = nullptr;
  DevBuf desc, member_first, fname, masks, recs, hist, chk, cb, chunk_off, member_off, member_check, member_isize;
  DevBuf src_off, dst_off, out_len, status, expect, kind, counter, ck_out, ck_pieces, ck_first, ck_piece_out, ck_partials;
  DevBuf in_stage, out_stage, lz2_tables;
  DevBuf seg_src, seg_dst, seg_len, seg_status, seg_kind, seg_expect, seg_cand, skip_mask;  // large-member segments
  DevBuf mark_scratch, mark_segs, seg_bits;  // speculative segments of a large member (uint16 symbols, descriptors)
  DevBuf order;             // work-queue order of an inflate launch (longest members first)
  DevBuf gate;              // gated inflate launch: [0] groups copied in, [1 .. ng] members done per group, then the group starts
  StreamMemOps memops;      // stream wait / write on device words (null: group-by-group launches instead)
  bool gated_unc = true;    // env ZB200_UNC_GATED=0 forces the group-by-group launches
  StageRing ring_in, ring_out;   // pinned slots for pageable callers (allocated on first use)
  CopyPool pool;
  uint64_t pending_len = 0;     // zb200_decode_begin's result, waiting in out_stage for zb200_decode_finish
  bool pending = false;
  uint64_t big_member_bytes = 512ull << 10;  // members at least this long are tried as parallel segments
  uint64_t single_member_bytes = 512ull << 10; // threshold when the call holds ONE input.  (24 KiB was measured: alice29.txt.gz has three blocks,
                                               // two decode passes over three segments cost what one serial pass costs -- 10.5 vs 9.4 ms -- so no gain)
  bool big_env = false;
  cudaEvent_t ev[10] = {};
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> gev;   // per-group events (H2D done, compute done, offsets ready)
  void *pin = nullptr;            // pinned host scratch for descriptors / offsets
  size_t pin_cap = 0;
  DevBuf group_end;               // device u64 per group: end offset of the group's output
  size_t dev_group_chunks = kMaxChunksPerGroup, host_group_chunks = kHostGroupChunks;
  uint64_t unc_group_out_bytes = 0;  // host uncompress: output bytes per pipelined member group (0: a quarter of the batch)
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

int ensure_pinned(zb200_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->pin_cap) return ZB200_OK;
  if (ctx->pin) cudaFreeHost(ctx->pin);
  ctx->pin = nullptr;
  ctx->pin_cap = 0;
  size_t want = bytes + bytes / 4 + 4096;
  if (cudaMallocHost(&ctx->pin, want) != cudaSuccess) {
    cudaGetLastError();
    ctx->last_err = "cudaMallocHost failed";
    return ZB200_ERR_NOMEM;
  }
  ctx->pin_cap = want;
  return ZB200_OK;
}

int ensure_group_events(zb200_ctx *ctx, size_t n) {
  while (ctx->gev.size() < n) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return ZB200_ERR_CUDA;
    ctx->gev.push_back(e);
  }
  return ZB200_OK;
}

bool is_pageable(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

int ring_ready(zb200_ctx *ctx, StageRing &r) {
  if (r.slot[0]) return ZB200_OK;
  for (int i = 0; i < kStageSlots; i++) {
    if (cudaMallocHost((void **)&r.slot[i], kStageSlotBytes) != cudaSuccess ||
        cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      ctx->last_err = "pinned staging ring: allocation failed";
      return ZB200_ERR_NOMEM;
    }
  }
  unsigned hc = std::thread::hardware_concurrency();
  ctx->pool.start((int)std::min<unsigned>(7, hc > 2 ? hc / 2 : 1));
  return ZB200_OK;
}

// host -> device on `st`: asynchronous for page-locked memory, staged through the pinned ring otherwise
int h2d_copy(zb200_ctx *ctx, uint8_t *d_dst, const uint8_t *h_src, size_t n, cudaStream_t st, bool pageable) {
  if (!n) return ZB200_OK;
  if (!pageable) {
    CK(cudaMemcpyAsync(d_dst, h_src, n, cudaMemcpyHostToDevice, st));
    return ZB200_OK;
  }
  StageRing &r = ctx->ring_in;
  int rc = ring_ready(ctx, r);
  if (rc) return rc;
  for (size_t off = 0; off < n; off += kStageSlotBytes) {
    const size_t len = std::min(kStageSlotBytes, n - off);
    const size_t k = r.next++ % kStageSlots;
    if (r.busy[k]) CK(cudaEventSynchronize(r.ev[k]));   // the DMA that last read this slot is done
    ctx->pool.copy(r.slot[k], h_src + off, len);
    CK(cudaMemcpyAsync(d_dst + off, r.slot[k], len, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(r.ev[k], st));
    r.busy[k] = true;
  }
  return ZB200_OK;
}

// finish the oldest / all pending device -> pageable copies (the slot's bytes go to their place)
int d2h_complete(zb200_ctx *ctx, size_t k) {
  StageRing &r = ctx->ring_out;
  if (!r.busy[k]) return ZB200_OK;
  CK(cudaEventSynchronize(r.ev[k]));
  ctx->pool.copy(r.out_dst[k], r.slot[k], r.out_n[k]);
  r.busy[k] = false;
  return ZB200_OK;
}
int d2h_flush(zb200_ctx *ctx) {
  if (!ctx->ring_out.slot[0]) return ZB200_OK;
  for (size_t i = 0; i < (size_t)kStageSlots; i++) {
    int rc = d2h_complete(ctx, (ctx->ring_out.next + i) % kStageSlots);   // oldest first
    if (rc) return rc;
  }
  return ZB200_OK;
}
// device -> host on `st`; for pageable memory the bytes land when d2h_flush (or a later d2h_copy) says so
int d2h_copy(zb200_ctx *ctx, uint8_t *h_dst, const uint8_t *d_src, size_t n, cudaStream_t st, bool pageable) {
  if (!n) return ZB200_OK;
  if (!pageable) {
    CK(cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDeviceToHost, st));
    return ZB200_OK;
  }
  StageRing &r = ctx->ring_out;
  int rc = ring_ready(ctx, r);
  if (rc) return rc;
  for (size_t off = 0; off < n; off += kStageSlotBytes) {
    const size_t len = std::min(kStageSlotBytes, n - off);
    const size_t k = r.next++ % kStageSlots;
    rc = d2h_complete(ctx, k);
    if (rc) return rc;
    CK(cudaMemcpyAsync(r.slot[k], d_src + off, len, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(r.ev[k], st));
    r.busy[k] = true;
    r.out_dst[k] = h_dst + off;
    r.out_n[k] = len;
  }
  return ZB200_OK;
}

struct Group {
  size_t m0, m1;        // members [m0, m1)
  size_t c0, nc;        // chunks [c0, c0 + nc) in the batch-wide descriptor array
  size_t first0;        // start of this group's (nm + 1) entries in the member_first / member_off arrays
  uint64_t in_lo, in_hi;  // source byte range
  uint64_t bound;       // output bound of the group
};

// ---- compress: device-resident (h_src == h_dst == nullptr) or pipelined host buffers ----
// With host buffers the batch is cut into groups and H2D(g+1) || kernels(g) || D2H(g-1) run
// on three streams; each group's output offset is chained on the device (out_base_ptr), so
// the only host waits are for the small per-group offset arrays that size the D2H copies.
int compress_locked(zb200_ctx *ctx, const uint8_t *d_src, const uint8_t *h_src, const uint64_t *src_offsets,
                    size_t n, int level, int data_format, const uint8_t *fname_lens, uint8_t *d_dst,
                    size_t dst_cap, uint8_t *h_dst, size_t h_dst_cap, uint64_t *dst_offsets, int *statuses,
                    size_t max_group_chunks) {
  if (level < -2 || level > 9) return ZB200_ERR_INVALID_LEVEL;
  if (data_format != ZB200_DF_GZIP && data_format != ZB200_DF_ZLIB && data_format != ZB200_DF_DEFLATE)
    return ZB200_ERR_INVALID_FORMAT;
  if (fname_lens)
    for (size_t i = 0; i < n; i++)
      if (fname_lens[i] > 25) return ZB200_ERR_ARG;
  if (((uintptr_t)d_dst & 3u) != 0) return ZB200_ERR_ARG;
  dst_cap &= ~(size_t)3;  // the packer writes whole 32-bit words: never touch a word that straddles the end
  for (size_t i = 0; i < n; i++) {
    if (src_offsets[i + 1] < src_offsets[i]) return ZB200_ERR_ARG;
    if (statuses) statuses[i] = ZB200_OK;
  }
  ctx->timing.lz_ms = ctx->timing.huff_ms = ctx->timing.scan_ms = ctx->timing.pack_ms = 0.f;
  ctx->timing.n_chunks = 0;
  dst_offsets[0] = 0;
  if (n == 0) return ZB200_OK;
  const uint64_t src_lo = src_offsets[0];  // d_src holds [src_lo, src_hi) rebased to 0 when staging from the host
  const bool src_pageable = h_src && is_pageable(h_src + src_lo), dst_pageable = h_dst && is_pageable(h_dst);

  // ---- plan: groups, descriptors ----
  std::vector<Group> groups;
  std::vector<ZbChunkDesc> desc;
  std::vector<uint32_t> first;
  size_t max_nc = 0, max_nm = 0;
  size_t chunks_left = 0;
  for (size_t i = 0; i < n; i++) {
    uint64_t len = src_offsets[i + 1] - src_offsets[i];
    chunks_left += len == 0 ? 1 : (size_t)((len + ZB_CHUNK_BYTES - 1) / ZB_CHUNK_BYTES);
  }
  const size_t group_cap = max_group_chunks;
  for (size_t m0 = 0; m0 < n;) {
    // host pipeline: the work after the last H2D (kernels + D2H of the last group) is not overlapped
    // with anything, so the groups shrink geometrically towards the end of the batch
    if (h_src) max_group_chunks = std::min(group_cap, std::max<size_t>(512, chunks_left / 2));
    Group g;
    g.m0 = m0;
    g.c0 = desc.size();
    g.first0 = first.size();
    g.bound = 0;
    size_t m1 = m0;
    while (m1 < n) {
      uint64_t len = src_offsets[m1 + 1] - src_offsets[m1];
      size_t nc = len == 0 ? 1 : (size_t)((len + ZB_CHUNK_BYTES - 1) / ZB_CHUNK_BYTES);
      if (desc.size() > g.c0 && desc.size() - g.c0 + nc > max_group_chunks) break;
      first.push_back((uint32_t)(desc.size() - g.c0));
      for (size_t k = 0; k < nc; k++) {
        ZbChunkDesc d;
        d.src_off = src_offsets[m1] - (h_src ? src_lo : 0) + (uint64_t)k * ZB_CHUNK_BYTES;
        d.len = (uint32_t)std::min<uint64_t>(ZB_CHUNK_BYTES, len - (uint64_t)k * ZB_CHUNK_BYTES);
        d.member = (uint32_t)(m1 - m0);
        d.flags = (k == 0 ? ZB_CHUNK_FIRST : 0u) | (k == nc - 1 ? ZB_CHUNK_LAST : 0u);
        d.pad = (level == -1 || level >= 2) ? (uint32_t)std::min<uint64_t>(32768, (uint64_t)k * ZB_CHUNK_BYTES) : 0u;
        desc.push_back(d);
      }
      g.bound += zb200_compress_bound((size_t)len, data_format) + 64;
      m1++;
    }
    g.m1 = m1;
    g.nc = desc.size() - g.c0;
    chunks_left -= std::min(chunks_left, g.nc);
    first.push_back((uint32_t)g.nc);
    g.in_lo = src_offsets[m0] - src_lo;
    g.in_hi = src_offsets[m1] - src_lo;
    max_nc = std::max(max_nc, g.nc);
    max_nm = std::max(max_nm, g.m1 - g.m0);
    groups.push_back(g);
    m0 = m1;
  }
  const size_t ng = groups.size(), nc_all = desc.size(), nfirst = first.size();

  ENSURE(ctx->desc, nc_all * sizeof(ZbChunkDesc));
  ENSURE(ctx->member_first, nfirst * sizeof(uint32_t));
  ENSURE(ctx->member_off, nfirst * sizeof(uint64_t));
  ENSURE(ctx->fname, n + 16);
  ENSURE(ctx->group_end, (ng + 1) * sizeof(uint64_t));
  ENSURE(ctx->masks, max_nc * ZB_WINDOWS_PER_CHUNK * sizeof(uint2));
  ENSURE(ctx->recs, max_nc * (size_t)ZB_RECS_PER_CHUNK * sizeof(uint32_t));
  ENSURE(ctx->hist, max_nc * (size_t)ZB_WARPS_PER_CHUNK * ZB_HIST_SYMS * sizeof(uint16_t));
  ENSURE(ctx->chk, max_nc * sizeof(ZbChunkCheck));
  ENSURE(ctx->cb, max_nc * sizeof(ZbCodebook));
  ENSURE(ctx->chunk_off, max_nc * sizeof(uint64_t));
  ENSURE(ctx->member_check, max_nm * sizeof(uint32_t));
  ENSURE(ctx->member_isize, max_nm * sizeof(uint32_t));
  if (level == -1 || level >= 2) ENSURE(ctx->lz2_tables, zb_lz2_table_bytes(nullptr));
  {
    int rc = ensure_pinned(ctx, nfirst * sizeof(uint64_t) + 64);
    if (rc) return rc;
    rc = ensure_group_events(ctx, 3 * ng + 1);
    if (rc) return rc;
  }
  uint64_t *pin_off = (uint64_t *)ctx->pin;

  cudaStream_t s = ctx->stream;
  cudaStream_t sh = h_src ? ctx->h2d_stream : s, sd = h_dst ? ctx->d2h_stream : s;
  CK(cudaMemcpyAsync(ctx->desc.p, desc.data(), nc_all * sizeof(ZbChunkDesc), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->member_first.p, first.data(), nfirst * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  if (fname_lens && data_format == ZB200_DF_GZIP)
    CK(cudaMemcpyAsync(ctx->fname.p, fname_lens, n, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(ctx->group_end.p, 0, sizeof(uint64_t), s));
  if (h_src || h_dst) {
    // the transfer streams must not run ahead of the setup above
    CK(cudaEventRecord(ctx->gev[3 * ng], s));
    if (h_src) CK(cudaStreamWaitEvent(sh, ctx->gev[3 * ng], 0));
    if (h_dst) CK(cudaStreamWaitEvent(sd, ctx->gev[3 * ng], 0));
  }
  CK(cudaEventRecord(ctx->ev[6], sh));

  auto make_work = [&](const Group &g, size_t gi) {
    ZbCompressWork w;
    w.src = d_src;
    w.dst = d_dst;
    w.dst_cap = dst_cap;
    w.desc = (const ZbChunkDesc *)ctx->desc.p + g.c0;
    w.member_first = (const uint32_t *)ctx->member_first.p + g.first0;
    w.fname_len = (fname_lens && data_format == ZB200_DF_GZIP) ? (const uint8_t *)ctx->fname.p + g.m0 : nullptr;
    w.masks = (uint2 *)ctx->masks.p;
    w.recs = (uint32_t *)ctx->recs.p;
    w.hist = (uint16_t *)ctx->hist.p;
    w.chk = (ZbChunkCheck *)ctx->chk.p;
    w.cb = (ZbCodebook *)ctx->cb.p;
    w.chunk_off = (uint64_t *)ctx->chunk_off.p;
    w.member_off = (uint64_t *)ctx->member_off.p + g.first0;
    w.member_check = (uint32_t *)ctx->member_check.p;
    w.member_isize = (uint32_t *)ctx->member_isize.p;
    w.tabs = ctx->d_tabs;
    w.lz2_tables = (uint2 *)ctx->lz2_tables.p;
    w.n_chunks = (uint32_t)g.nc;
    w.n_members = (uint32_t)(g.m1 - g.m0);
    w.level = level;
    w.data_format = data_format;
    w.out_base = 0;
    w.out_base_ptr = (const uint64_t *)ctx->group_end.p + gi;
    return w;
  };

  // ---- enqueue every group ----
  size_t d2h_done = 0;  // groups whose output has been handed to the D2H stream
  uint64_t total_out = 0;
  auto drain_d2h = [&](size_t upto) -> int {  // enqueue D2H for groups [d2h_done, upto)
    for (; d2h_done < upto; d2h_done++) {
      const Group &g = groups[d2h_done];
      CK(cudaEventSynchronize(ctx->gev[3 * d2h_done + 2]));  // offsets of this group are in pin_off
      const size_t nm = g.m1 - g.m0;
      const uint64_t lo = pin_off[g.first0], hi = pin_off[g.first0 + nm];
      for (size_t j = 0; j <= nm; j++) dst_offsets[g.m0 + j] = pin_off[g.first0 + j];
      total_out = hi;
      if (h_dst) {
        if (hi > h_dst_cap) return ZB200_ERR_DST_TOO_SMALL;
        CK(cudaStreamWaitEvent(sd, ctx->gev[3 * d2h_done + 1], 0));
        if (d2h_done == 0) CK(cudaEventRecord(ctx->ev[8], sd));
        if (hi > lo) {
          int rc2 = d2h_copy(ctx, h_dst + lo, d_dst + lo, (size_t)(hi - lo), sd, dst_pageable);
          if (rc2) return rc2;
        }
      }
    }
    return ZB200_OK;
  };

  for (size_t gi = 0; gi < ng; gi++) {
    const Group &g = groups[gi];
    if (h_src) {
      if (g.in_hi > g.in_lo) {
        int rc2 = h2d_copy(ctx, (uint8_t *)d_src + g.in_lo, h_src + src_lo + g.in_lo, (size_t)(g.in_hi - g.in_lo), sh, src_pageable);
        if (rc2) return rc2;
      }
      CK(cudaEventRecord(ctx->gev[3 * gi + 0], sh));
      CK(cudaStreamWaitEvent(s, ctx->gev[3 * gi + 0], 0));
    }
    ZbCompressWork w = make_work(g, gi);
    const bool timed = (gi == 0);  // per-kernel events on the first group; totals are scaled by chunk count
    if (timed) CK(cudaEventRecord(ctx->ev[0], s));
    CK(zb_launch_lz(w, s));
    if (timed) CK(cudaEventRecord(ctx->ev[1], s));
    CK(zb_launch_huff(w, s));
    if (timed) CK(cudaEventRecord(ctx->ev[2], s));
    CK(zb_launch_scan(w, s));
    if (timed) CK(cudaEventRecord(ctx->ev[3], s));
    // chain: the next group starts where this one ended
    CK(cudaMemcpyAsync((uint64_t *)ctx->group_end.p + gi + 1, w.member_off + w.n_members, sizeof(uint64_t),
                       cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(pin_off + g.first0, w.member_off, (w.n_members + 1) * sizeof(uint64_t),
                       cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->gev[3 * gi + 2], s));
    if (timed) CK(cudaEventRecord(ctx->ev[4], s));
    CK(zb_launch_pack(w, s));
    if (timed) CK(cudaEventRecord(ctx->ev[5], s));
    CK(cudaEventRecord(ctx->gev[3 * gi + 1], s));
    ctx->timing.kernel_launches += 5;
    ctx->timing.n_chunks += (uint32_t)g.nc;
    // keep at most two groups of output waiting on the device before draining to the host
    if (h_dst && gi >= 2) {
      int rc = drain_d2h(gi - 1);
      if (rc) return rc;
    }
  }
  if (h_src) CK(cudaEventRecord(ctx->ev[7], sh));
  {
    int rc = drain_d2h(ng);
    if (rc) return rc;
  }
  if (h_dst) CK(cudaEventRecord(ctx->ev[9], sd));
  if (dst_pageable) {
    int rc = d2h_flush(ctx);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(s));
  if (h_src) CK(cudaStreamSynchronize(sh));
  if (h_dst) CK(cudaStreamSynchronize(sd));
  if (total_out > dst_cap) return ZB200_ERR_DST_TOO_SMALL;
  // per-kernel times: measured on the first group, scaled to the batch by chunk count
  const float scale = groups[0].nc ? (float)nc_all / (float)groups[0].nc : 1.f;
  ctx->timing.lz_ms = ev_ms(ctx->ev[0], ctx->ev[1]) * scale;
  ctx->timing.huff_ms = ev_ms(ctx->ev[1], ctx->ev[2]) * scale;
  ctx->timing.scan_ms = ev_ms(ctx->ev[2], ctx->ev[3]) * scale;
  ctx->timing.pack_ms = ev_ms(ctx->ev[4], ctx->ev[5]) * scale;
  return ZB200_OK;
}

// ZB_CK_PIECE_BYTES pieces covering the capacity [offs[i], offs[i+1]) of every buffer (at least one per buffer)
int upload_pieces(zb200_ctx *ctx, const uint64_t *offs, size_t n, ZbChecksumWork &w) {
  std::vector<ZbPiece> pieces;
  std::vector<uint32_t> first(n + 1);
  for (size_t i = 0; i < n; i++) {
    first[i] = (uint32_t)pieces.size();
    uint64_t cap = offs[i + 1] - offs[i];
    uint64_t rel = 0;
    do {
      ZbPiece p;
      p.rel = rel;
      p.buf = (uint32_t)i;
      p.pad = 0;
      pieces.push_back(p);
      rel += ZB_CK_PIECE_BYTES;
    } while (rel < cap);
  }
  first[n] = (uint32_t)pieces.size();
  w.big_pieces = 0;
  for (size_t i = 0; i < n; i++)
    if (first[i + 1] - first[i] > ZB_CK_BIG_PIECES) w.big_pieces = ZB_CK_BIG_PIECES;
  ENSURE(ctx->ck_pieces, pieces.size() * sizeof(ZbPiece));
  ENSURE(ctx->ck_first, (n + 1) * sizeof(uint32_t));
  ENSURE(ctx->ck_piece_out, pieces.size() * sizeof(ZbChunkCheck));
  ENSURE(ctx->ck_partials, pieces.size() * (size_t)ZB_CK_PARTIAL_BYTES);
  w.partials = (uint32_t *)ctx->ck_partials.p;
  CK(cudaMemcpyAsync(ctx->ck_pieces.p, pieces.data(), pieces.size() * sizeof(ZbPiece), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->ck_first.p, first.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // the host vectors go out of scope
  w.pieces = (const ZbPiece *)ctx->ck_pieces.p;
  w.first = (const uint32_t *)ctx->ck_first.p;
  w.piece_out = (ZbChunkCheck *)ctx->ck_piece_out.p;
  w.n_pieces = (uint32_t)pieces.size();
  w.tabs = ctx->d_tabs;
  w.n = (uint32_t)n;
  return ZB200_OK;
}


// ---- large members as parallel segments ----
// A DEFLATE stream is serial, and one member is decoded by one 8-lane group: a single multi-MiB
// member would crawl.  But this library's own multi-chunk members (and zlib's Z_FULL_FLUSH /
// pigz -i streams) are chains of INDEPENDENT, byte-aligned pieces, each ending with the empty
// stored block 00 00 ff ff.  So a large member is handled speculatively: find every 00 00 ff ff
// in its payload, decode the pieces between them as separate raw-deflate segments (a count pass
// for the sizes, then the real pass at the prefix-summed positions), and accept the result only
// if every segment decodes cleanly, only the last one holds the final block, and the sizes add up
// inside the member's capacity.  Anything else (a false 00 00 ff ff inside data, a back-reference
// across a boundary, a corrupt stream) leaves the member to the ordinary serial decode, which
// also produces the reference's error for it.  The trailer check runs on the output either way.
struct HostWrapper {
  int fmt;
  uint64_t pos, end;  // payload [pos, end) inside the member
  uint32_t expect, isize;
};

// zippy.nim:100-165 + gzip.nim:3-66 on the first hn bytes / last 8 bytes of a member; false when
// the member is not acceptable or the header does not fit in hn (then the serial path decides).
bool host_parse_wrapper(const uint8_t *h, size_t hn, const uint8_t *t8, uint64_t len, int fmt, uint64_t raw_pos,
                        HostWrapper &w) {
  auto le32 = [](const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
  if (fmt == ZB200_DF_DETECT) {
    if (len > 18 && hn >= 4 && h[0] == 31 && h[1] == 139 && h[2] == 8 && (h[3] & 0xe0) == 0) fmt = ZB200_DF_GZIP;
    else if (len > 6 && hn >= 2 && (h[0] & 0x0f) == 8 && (h[0] >> 4) <= 7 && (((uint32_t)h[0] * 256u) + h[1]) % 31u == 0)
      fmt = ZB200_DF_ZLIB;
    else return false;
  }
  w.fmt = fmt;
  w.expect = w.isize = 0;
  if (fmt == ZB200_DF_GZIP) {
    if (len < 18 || hn < 10) return false;
    uint32_t flg = h[3];
    if (h[0] != 31 || h[1] != 139 || h[2] != 8 || (flg & 0xe0) || (flg & 4)) return false;
    uint64_t p = 10;
    for (int pass = 0; pass < 2; pass++)
      if ((pass == 0 && (flg & 8)) || (pass == 1 && (flg & 16))) {
        while (p < hn && h[p] != 0) p++;
        if (p >= hn) return false;
        p++;
      }
    if (flg & 2) p += 2;
    if (p + 8 >= len) return false;
    w.pos = p;
    w.end = len - 8;
    w.expect = le32(t8);
    w.isize = le32(t8 + 4);
    return true;
  }
  if (fmt == ZB200_DF_ZLIB) {
    if (len < 6 || hn < 2) return false;
    uint32_t cmf = h[0], flg = h[1];
    if ((cmf & 0x0f) != 8 || (cmf >> 4) > 7 || (cmf * 256u + flg) % 31u != 0 || (flg & 0x20)) return false;
    w.pos = 2;
    w.end = len - 4;
    w.expect = ((uint32_t)t8[4] << 24) | ((uint32_t)t8[5] << 16) | ((uint32_t)t8[6] << 8) | t8[7];
    return true;
  }
  if (fmt == ZB200_DF_DEFLATE) {
    if (raw_pos > len) return false;
    w.pos = raw_pos;
    w.end = len;
    return true;
  }
  return false;
}

struct BigResult {
  size_t member;
  uint64_t out_len;
  uint32_t kind, expect;
  int status = ZB200_OK;
};

// ---- a large member WITHOUT sync markers (any foreign gzip / zlib / raw stream) ----
// 1. k_find_blocks lists every plausible dynamic-block start in the payload; the list is thinned to one
//    boundary per >= 16 KiB of input.  2. A counting pass decodes every segment [boundary i, boundary i+1)
//    in parallel, each from its own block start with an UNKNOWN 32 KiB window (back-references before the
//    segment's start are allowed, nothing is written): it must end exactly on the next boundary, and only the
//    last segment may hold the final block -- this is what exposes a false boundary.  3. The sizes give every
//    segment its place; the segments are decoded again into uint16 symbols, with marker symbols standing in
//    for the bytes of the unknown window.  4. The markers are resolved: the last 32 KiB of every segment in
//    order (one CTA), then everything else in parallel.  Anything irregular -- a decode error, a boundary
//    that is not hit, a marker that points before the start of the stream -- leaves the member to the serial
//    decode, which also produces the reference's verdict for it.  The trailer check runs on the output
//    either way.
int inflate_member_speculative(zb200_ctx *ctx, const uint8_t *d_src, uint64_t m0, const HostWrapper &hw, uint8_t *d_dst,
                               uint64_t dst0, uint64_t mcap, bool count_only, bool latency, bool &ok, uint64_t &out_len,
                               bool &too_small) {
  ok = false;
  too_small = false;
  cudaStream_t s = ctx->stream;
  const uint64_t lo_bit = (m0 + hw.pos) * 8ull, hi_bit = (m0 + hw.end) * 8ull, limit_byte = m0 + hw.end;
  const uint32_t cap = (uint32_t)std::min<uint64_t>((hw.end - hw.pos) / 64 + 1024, 1u << 24);
  ENSURE(ctx->seg_cand, (size_t)cap * 8 + 16);
  ENSURE(ctx->counter, 64);
  uint32_t *d_cnt = (uint32_t *)ctx->counter.p + 8;
  CK(zb_launch_find_blocks(d_src, lo_bit, hi_bit, limit_byte, (uint64_t *)ctx->seg_cand.p, cap, d_cnt, s));
  uint32_t cnt = 0;
  CK(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ctx->timing.kernel_launches += 1;
  if (cnt == 0 || cnt > cap) return ZB200_OK;
  std::vector<uint64_t> cand(cnt);
  CK(cudaMemcpyAsync(cand.data(), ctx->seg_cand.p, (size_t)cnt * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  std::sort(cand.begin(), cand.end());
  // boundaries: the payload start, then candidates at least min_gap apart (a segment costs a 64 KiB marker prefill)
  // (at most 60000 segments: the resolve kernels index them with a grid dimension)
  // a single input is all the GPU has: cut it as finely as its blocks allow
  const uint64_t min_gap = std::max<uint64_t>((latency ? 2048ull : 16384ull) * 8ull, (hi_bit - lo_bit) / 60000ull);
  std::vector<uint64_t> bits(1, lo_bit);
  for (uint64_t c : cand)
    if (c >= bits.back() + min_gap && c + min_gap / 4 < hi_bit) bits.push_back(c);
  const size_t S = bits.size();
  if (S < 2) return ZB200_OK;
  std::vector<uint64_t> sb(2 * S);
  for (size_t i = 0; i < S; i++) {
    sb[2 * i] = bits[i];
    sb[2 * i + 1] = i + 1 < S ? bits[i + 1] : hi_bit;
  }
  ENSURE(ctx->seg_bits, 2 * S * 8);
  ENSURE(ctx->seg_dst, (S + 1) * 8);
  ENSURE(ctx->seg_len, S * 8);
  ENSURE(ctx->seg_status, S * 4);
  ENSURE(ctx->seg_kind, S * 4);
  ENSURE(ctx->seg_expect, S * 4);
  CK(cudaMemcpyAsync(ctx->seg_bits.p, sb.data(), 2 * S * 8, cudaMemcpyHostToDevice, s));
  ZbInflateWork w;
  memset(&w, 0, sizeof(w));
  w.src = d_src;
  w.seg_bits = (const uint64_t *)ctx->seg_bits.p;
  w.seg_limit = limit_byte;
  w.dst_off = (const uint64_t *)ctx->seg_dst.p;
  w.out_len = (uint64_t *)ctx->seg_len.p;
  w.status = (int *)ctx->seg_status.p;
  w.expect = (uint32_t *)ctx->seg_expect.p;
  w.kind = (uint32_t *)ctx->seg_kind.p;
  w.counter = (uint32_t *)ctx->counter.p + 4;
  w.tabs = ctx->d_tabs;
  w.n = (uint32_t)S;
  w.data_format = ZB200_DF_DEFLATE;
  w.seg_mode = 1;
  std::vector<uint64_t> sl(S);
  std::vector<int> sst(S);
  std::vector<uint32_t> sk(S);
  auto fetch = [&]() -> int {
    CK(cudaMemcpyAsync(sl.data(), ctx->seg_len.p, S * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(sst.data(), ctx->seg_status.p, S * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(sk.data(), ctx->seg_kind.p, S * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ZB200_OK;
  };
  // 2. the counting pass
  w.count_only = 1;
  CK(zb_launch_inflate(w, s));
  int rc = fetch();
  if (rc) return rc;
  ctx->timing.kernel_launches += 1;
  uint64_t total = 0, scr_elems = 0;
  uint32_t max_n = 0;
  for (size_t i = 0; i < S; i++) {
    if (sst[i] != ZB200_OK || (sk[i] != 0) != (i + 1 == S) || sl[i] > 0xf0000000ull) return ZB200_OK;
    total += sl[i];
    scr_elems += 32768ull + sl[i];
    max_n = std::max<uint32_t>(max_n, (uint32_t)sl[i]);
  }
  if (count_only) {
    ok = true;
    out_len = total;
    return ZB200_OK;
  }
  if (total > mcap) {   // the whole stream decodes, so the serial decode could only run out of room: say so now
    too_small = true;
    return ZB200_OK;
  }
  // 3. uint16 symbols, markers in front of every segment
  ENSURE(ctx->mark_scratch, (size_t)scr_elems * 2 + 64);
  ENSURE(ctx->mark_segs, S * sizeof(ZbMarkSegHost) + 16);
  std::vector<ZbMarkSegHost> segs(S);
  std::vector<uint64_t> dof(S + 1);
  uint64_t se = 0, de = dst0;
  for (size_t i = 0; i < S; i++) {
    se += 32768ull;
    segs[i].scr = se;
    segs[i].dst = de;
    segs[i].n = (uint32_t)sl[i];
    segs[i].pad = 0;
    dof[i] = se;
    se += sl[i];
    de += sl[i];
  }
  dof[S] = se;
  const std::vector<uint64_t> want = sl;
  CK(cudaMemcpyAsync(ctx->mark_segs.p, segs.data(), S * sizeof(ZbMarkSegHost), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->seg_dst.p, dof.data(), (S + 1) * 8, cudaMemcpyHostToDevice, s));
  int *d_bad = (int *)((uint32_t *)ctx->counter.p + 12);
  CK(cudaMemsetAsync(d_bad, 0, 4, s));
  CK(zb_launch_mark_prefill((uint16_t *)ctx->mark_scratch.p, ctx->mark_segs.p, (uint32_t)S, s));
  w.count_only = 0;
  w.mark = 1;
  w.dst = (uint8_t *)ctx->mark_scratch.p;
  CK(zb_launch_inflate(w, s));
  rc = fetch();
  if (rc) return rc;
  ctx->timing.kernel_launches += 2;
  for (size_t i = 0; i < S; i++)
    if (sst[i] != ZB200_OK || sl[i] != want[i]) return ZB200_OK;
  // 4. markers -> bytes
  CK(zb_launch_resolve((const uint16_t *)ctx->mark_scratch.p, ctx->mark_segs.p, (uint32_t)S, max_n, d_dst, d_bad, s));
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ctx->timing.kernel_launches += 2;
  if (bad) return ZB200_OK;
  ok = true;
  out_len = total;
  return ZB200_OK;
}

// Tries every large member; `done` gets the members that were fully decoded here (their output
// is in place; status / length / kind / expect still have to be written to the device arrays).
int inflate_big_members(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n, int data_format,
                        uint64_t raw_pos, uint8_t *d_dst, const uint64_t *dst_offsets, bool count_only,
                        std::vector<BigResult> &done) {
  cudaStream_t s = ctx->stream;
  // Every large member costs a few host round trips here, while a batch of thousands of members
  // already fills the GPU with one group per member: with many large members only the huge ones
  // (minutes of serial decode) take this path.
  uint64_t min_len = (n == 1 && !ctx->big_env) ? ctx->single_member_bytes : ctx->big_member_bytes;
  {
    size_t count = 0;
    for (size_t m = 0; m < n; m++) count += (src_offsets[m + 1] - src_offsets[m]) >= min_len;
    if (count == 0) return ZB200_OK;
    if (count > 256) min_len = std::max<uint64_t>(min_len, 64ull << 20);
  }
  for (size_t m = 0; m < n; m++) {
    const uint64_t m0 = src_offsets[m], len = src_offsets[m + 1] - m0;
    if (len < min_len) continue;
    uint8_t head[1024], tail[8];
    const size_t hn = (size_t)std::min<uint64_t>(len, sizeof(head));
    CK(cudaMemcpyAsync(head, d_src + m0, hn, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(tail, d_src + m0 + len - 8, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    HostWrapper hw;
    if (!host_parse_wrapper(head, hn, tail, len, data_format, raw_pos, hw)) continue;
    if (count_only && hw.fmt == ZB200_DF_GZIP) continue;  // ISIZE answers that (gzip.nim:66)
    if (hw.end <= hw.pos + 4) continue;
    const uint64_t dst0_m = count_only ? 0 : dst_offsets[m], mcap_m = count_only ? ~0ull : dst_offsets[m + 1] - dst_offsets[m];
    // streams without sync markers (or whose pieces are not independent): speculative segments
    auto speculative = [&]() -> int {
      bool sok = false, small = false;
      uint64_t slen = 0;
      int src_ = inflate_member_speculative(ctx, d_src, m0, hw, d_dst, dst0_m, mcap_m, count_only, n == 1, sok, slen, small);
      if (src_) return src_;
      if (sok || small) {
        BigResult r;
        r.member = m;
        r.out_len = sok ? slen : 0;
        r.kind = (uint32_t)hw.fmt;
        r.expect = hw.expect;
        r.status = small ? ZB200_ERR_DST_TOO_SMALL : ZB200_OK;
        done.push_back(r);
      }
      return ZB200_OK;
    };
#define ZB_TRY_SPECULATIVE()      \
  {                               \
    int _rc = speculative();      \
    if (_rc) return _rc;          \
    continue;                     \
  }
    // 1. candidate boundaries
    const uint32_t cap = (uint32_t)std::min<uint64_t>((hw.end - hw.pos) / 32 + 64, 1u << 24);
    ENSURE(ctx->seg_cand, (size_t)cap * 8 + 16);
    ENSURE(ctx->counter, 64);
    uint32_t *d_cnt = (uint32_t *)ctx->counter.p + 8;
    CK(zb_launch_find_sync(d_src + m0, hw.pos, hw.end, (uint64_t *)ctx->seg_cand.p, cap, d_cnt, s));
    uint32_t cnt = 0;
    CK(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (cnt == 0 || cnt > cap) ZB_TRY_SPECULATIVE();
    std::vector<uint64_t> bounds(cnt + 2);
    CK(cudaMemcpyAsync(bounds.data() + 1, ctx->seg_cand.p, (size_t)cnt * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::sort(bounds.begin() + 1, bounds.begin() + 1 + cnt);
    bounds[0] = hw.pos;
    size_t S = cnt + 1;
    if (bounds[cnt] >= hw.end) S = cnt;  // the payload ends with a marker: no trailing segment
    else bounds[cnt + 1] = hw.end;
    if (S < 2) ZB_TRY_SPECULATIVE();
    for (size_t j = 0; j <= S; j++) bounds[j] += m0;  // absolute in d_src
    ENSURE(ctx->seg_src, (S + 1) * 8);
    ENSURE(ctx->seg_dst, (S + 1) * 8);
    ENSURE(ctx->seg_len, S * 8);
    ENSURE(ctx->seg_status, S * 4);
    ENSURE(ctx->seg_kind, S * 4);
    ENSURE(ctx->seg_expect, S * 4);
    CK(cudaMemcpyAsync(ctx->seg_src.p, bounds.data(), (S + 1) * 8, cudaMemcpyHostToDevice, s));
    ZbInflateWork w;
    memset(&w, 0, sizeof(w));
    w.src = d_src;
    w.src_off = (const uint64_t *)ctx->seg_src.p;
    w.dst = d_dst;
    w.dst_off = (const uint64_t *)ctx->seg_dst.p;
    w.out_len = (uint64_t *)ctx->seg_len.p;
    w.status = (int *)ctx->seg_status.p;
    w.expect = (uint32_t *)ctx->seg_expect.p;
    w.kind = (uint32_t *)ctx->seg_kind.p;
    w.counter = (uint32_t *)ctx->counter.p + 4;
    w.tabs = ctx->d_tabs;
    w.n = (uint32_t)S;
    w.data_format = ZB200_DF_DEFLATE;
    w.pos = 0;
    w.seg_mode = 1;
    std::vector<uint64_t> sl(S), dof(S + 1);
    std::vector<int> sst(S);
    std::vector<uint32_t> sk(S);
    auto fetch = [&]() -> int {
      CK(cudaMemcpyAsync(sl.data(), ctx->seg_len.p, S * 8, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(sst.data(), ctx->seg_status.p, S * 4, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(sk.data(), ctx->seg_kind.p, S * 4, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      return ZB200_OK;
    };
    const uint64_t dst0 = count_only ? 0 : dst_offsets[m], mcap = count_only ? ~0ull : dst_offsets[m + 1] - dst_offsets[m];
    bool ok = false;
    int rc;
    // 2. the optimistic pass: this library's own members have 64 KiB of output per segment (the
    // last one takes what is left); if every segment agrees, one pass was enough
    if (!count_only && (S - 1) * (uint64_t)ZB_CHUNK_BYTES < mcap) {
      for (size_t j = 0; j < S; j++) dof[j] = dst0 + j * (uint64_t)ZB_CHUNK_BYTES;
      dof[S] = dst0 + mcap;
      CK(cudaMemcpyAsync(ctx->seg_dst.p, dof.data(), (S + 1) * 8, cudaMemcpyHostToDevice, s));
      w.count_only = 0;
      CK(zb_launch_inflate(w, s));
      rc = fetch();
      if (rc) return rc;
      ctx->timing.kernel_launches += 2;
      ok = true;
      for (size_t j = 0; j < S && ok; j++)
        ok = sst[j] == ZB200_OK && (sk[j] != 0) == (j + 1 == S) && (j + 1 == S || sl[j] == (uint64_t)ZB_CHUNK_BYTES);
      if (ok) dof[S] = dof[S - 1] + sl[S - 1];
    }
    if (!ok) {
      // 3. sizes from a count pass, then the real pass with every segment at its place
      w.count_only = 1;
      CK(zb_launch_inflate(w, s));
      rc = fetch();
      if (rc) return rc;
      ctx->timing.kernel_launches += 2;
      ok = true;
      dof[0] = dst0;
      for (size_t j = 0; j < S && ok; j++) {
        ok = sst[j] == ZB200_OK && (sk[j] != 0) == (j + 1 == S);
        dof[j + 1] = dof[j] + sl[j];
      }
      if (!ok || dof[S] - dst0 > mcap) ZB_TRY_SPECULATIVE();
      if (!count_only) {
        std::vector<uint64_t> want = sl;
        CK(cudaMemcpyAsync(ctx->seg_dst.p, dof.data(), (S + 1) * 8, cudaMemcpyHostToDevice, s));
        w.count_only = 0;
        CK(zb_launch_inflate(w, s));
        rc = fetch();
        if (rc) return rc;
        ctx->timing.kernel_launches += 1;
        for (size_t j = 0; j < S && ok; j++) ok = sst[j] == ZB200_OK && sl[j] == want[j];
        if (!ok) ZB_TRY_SPECULATIVE();
      }
    }
    BigResult r;
    r.member = m;
    r.out_len = dof[S] - dst0;
    r.kind = (uint32_t)hw.fmt;
    r.expect = hw.expect;
    done.push_back(r);
  }
  return ZB200_OK;
}

// Work-queue order of one inflate launch: a member is decoded by one 8-lane group from start to end,
// so a long member that is fetched late finishes long after everything else (the tail of the launch).
// Members much longer than the average go first, longest first; the rest keep their order.
// Returns false when no member stands out (then the launch uses index order and nothing is uploaded).
bool longest_first_order(const uint64_t *src_offsets, size_t n, std::vector<uint32_t> &order) {
  if (n < 64) return false;
  const uint64_t total = src_offsets[n] - src_offsets[0];
  const uint64_t thr = std::max<uint64_t>(4 * (total / n), 32768);
  std::vector<uint32_t> big;
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] - src_offsets[i] >= thr) big.push_back((uint32_t)i);
  if (big.empty() || big.size() > n / 2) return false;
  std::sort(big.begin(), big.end(), [&](uint32_t a, uint32_t b) {
    const uint64_t la = src_offsets[a + 1] - src_offsets[a], lb = src_offsets[b + 1] - src_offsets[b];
    return la != lb ? la > lb : a < b;
  });
  order.resize(n);
  size_t k = 0;
  for (uint32_t i : big) order[k++] = i;
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] - src_offsets[i] < thr) order[k++] = (uint32_t)i;
  return true;
}

// ---- uncompress, device-resident ----
int uncompress_device_locked(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                             int data_format, uint64_t raw_pos, uint8_t *d_dst, const uint64_t *dst_offsets,
                             uint64_t *dst_lens, int *statuses, bool count_only,
                             const std::function<int()> *after_launch = nullptr) {
  if (data_format < ZB200_DF_DETECT || data_format > ZB200_DF_DEFLATE) return ZB200_ERR_INVALID_FORMAT;
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
  memset(&w, 0, sizeof(w));
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
  w.skip = nullptr;
  w.seg_mode = 0;
  w.order = nullptr;
  {
    std::vector<uint32_t> order;
    if (longest_first_order(src_offsets, n, order)) {
      ENSURE(ctx->order, n * sizeof(uint32_t));
      CK(cudaMemcpyAsync(ctx->order.p, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
      CK(cudaStreamSynchronize(s));  // `order` goes out of scope
      w.order = (const uint32_t *)ctx->order.p;
    }
  }
  ZbChecksumWork cw;
  memset(&cw, 0, sizeof(cw));
  if (!count_only) {  // piece table for the verification pass, uploaded before anything is launched
    int rc = upload_pieces(ctx, dst_offsets, n, cw);
    if (rc) return rc;
  }
  CK(cudaEventRecord(ctx->ev[0], s));
  // large members first, as parallel segments where their streams allow it; the rest (and every
  // large member that did not work out) goes through the ordinary launch below
  std::vector<BigResult> big;
  std::vector<uint8_t> skip_host;
  {
    int rc = inflate_big_members(ctx, d_src, src_offsets, n, data_format, raw_pos, d_dst, dst_offsets, count_only, big);
    if (rc) return rc;
    if (!big.empty()) {
      skip_host.assign(n, 0);
      ENSURE(ctx->skip_mask, n);
      for (const BigResult &r : big) {
        skip_host[r.member] = 1;
        CK(cudaMemcpyAsync((int *)ctx->status.p + r.member, &r.status, 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync((uint64_t *)ctx->out_len.p + r.member, &r.out_len, 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync((uint32_t *)ctx->kind.p + r.member, &r.kind, 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync((uint32_t *)ctx->expect.p + r.member, &r.expect, 4, cudaMemcpyHostToDevice, s));
      }
      CK(cudaMemcpyAsync(ctx->skip_mask.p, skip_host.data(), n, cudaMemcpyHostToDevice, s));
      CK(cudaStreamSynchronize(s));
      w.skip = (const uint8_t *)ctx->skip_mask.p;
    }
  }
  CK(zb_launch_inflate(w, s));
  CK(cudaEventRecord(ctx->ev[1], s));
  if (!count_only) {
    // gzip.nim:80-88 / zippy.nim:154-162: checksum, then size, of every member that inflated
    cw.src = d_dst;
    cw.off = (const uint64_t *)ctx->dst_off.p;
    cw.lens = (const uint64_t *)ctx->out_len.p;
    cw.status = (int *)ctx->status.p;
    cw.expect = (const uint32_t *)ctx->expect.p;
    cw.kinds = (const uint32_t *)ctx->kind.p;
    cw.isize_src = d_src;
    cw.isize_off = (const uint64_t *)ctx->src_off.p;
    CK(zb_launch_checksum(cw, s));
  }
  CK(cudaEventRecord(ctx->ev[2], s));
  if (after_launch) {  // work to queue behind the kernels (other streams) before this thread waits
    int rc = (*after_launch)();
    if (rc) return rc;
  }
  CK(cudaMemcpyAsync(dst_lens, ctx->out_len.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  std::vector<int> st_tmp;
  int *st = statuses;
  if (!st) {
    st_tmp.resize(n);
    st = st_tmp.data();
  }
  CK(cudaMemcpyAsync(st, ctx->status.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ctx->timing.inflate_ms += ev_ms(ctx->ev[0], ctx->ev[1]);
  ctx->timing.verify_ms += ev_ms(ctx->ev[1], ctx->ev[2]);
  ctx->timing.kernel_launches += count_only ? 1 : 3;
  for (size_t i = 0; i < n; i++)
    if (st[i] != ZB200_OK) dst_lens[i] = 0;
  return ZB200_OK;
}

// ---- uncompress, host buffers, fully asynchronous ----
// Everything the device needs for the WHOLE batch (offsets, verification pieces, work-queue orders) is
// built and uploaded once; then every member group is enqueued without a host wait in between:
//   H2D stream : copy-in of group 0, 1, 2, ... back to back
//   main stream: wait copy-in(g) -> inflate(g) -> verify(g)
//   D2H stream : wait verify(g) -> copy-out(g)
// and the host synchronises once at the end.  (The first version waited for every group's kernels on the
// host before it built and launched the next group, and each launch carried its own tail of long members.)
int uncompress_host_pipelined(zb200_ctx *ctx, const uint8_t *h_src, const std::vector<uint64_t> &reb, size_t n,
                              int data_format, uint8_t *h_dst, const std::vector<uint64_t> &dreb, uint64_t *dst_lens,
                              int *statuses, const std::vector<size_t> &gb) {
  const size_t ng = gb.size() - 1;
  // gated: ONE inflate launch walks the whole batch behind the copy-in (no per-group launch tails), the D2H
  // stream waits on the per-group done counts; otherwise one launch per group, chained with events
  const bool gated = ctx->memops.ok && ctx->gated_unc && n < 0xffffffffull;
  cudaStream_t s = ctx->stream, sh = ctx->h2d_stream, sd = ctx->d2h_stream;
  const uint8_t *d_src = (const uint8_t *)ctx->in_stage.p;
  uint8_t *d_dst = (uint8_t *)ctx->out_stage.p;
  // ---- plan (host) ----
  std::vector<ZbPiece> pieces;
  std::vector<uint32_t> first;          // per group: (members + 1) entries relative to the group's first piece
  std::vector<size_t> piece0(ng + 1), first0(ng + 1);
  std::vector<uint32_t> order(n);       // per group: entries relative to the group
  std::vector<uint8_t> has_order(ng, 0);
  for (size_t gi = 0; gi < ng; gi++) {
    const size_t m0 = gb[gi], m1 = gb[gi + 1];
    piece0[gi] = pieces.size();
    first0[gi] = first.size();
    for (size_t i = m0; i < m1; i++) {
      first.push_back((uint32_t)(pieces.size() - (gated ? 0 : piece0[gi])));
      const uint64_t cap = dreb[i + 1] - dreb[i];
      uint64_t rel = 0;
      do {
        ZbPiece pc;
        pc.rel = rel;
        pc.buf = (uint32_t)(i - (gated ? 0 : m0));
        pc.pad = 0;
        pieces.push_back(pc);
        rel += ZB_CK_PIECE_BYTES;
      } while (rel < cap);
    }
    if (!gated || gi + 1 == ng) first.push_back((uint32_t)(pieces.size() - (gated ? 0 : piece0[gi])));
    std::vector<uint32_t> o;
    if (longest_first_order(reb.data() + m0, m1 - m0, o)) {
      has_order[gi] = 1;
      std::copy(o.begin(), o.end(), order.begin() + m0);
      if (gated)
        for (size_t i = m0; i < m1; i++) order[i] += (uint32_t)m0;
    } else if (gated) {
      for (size_t i = m0; i < m1; i++) order[i] = (uint32_t)i;
    }
  }
  piece0[ng] = pieces.size();
  first0[ng] = first.size();
  // ---- device arrays for the whole batch ----
  ENSURE(ctx->src_off, (n + 1) * sizeof(uint64_t));
  ENSURE(ctx->dst_off, (n + 1) * sizeof(uint64_t));
  ENSURE(ctx->out_len, n * sizeof(uint64_t));
  ENSURE(ctx->status, n * sizeof(int));
  ENSURE(ctx->expect, n * sizeof(uint32_t));
  ENSURE(ctx->kind, n * sizeof(uint32_t));
  ENSURE(ctx->order, n * sizeof(uint32_t));
  ENSURE(ctx->counter, (ng + 16) * sizeof(uint32_t));
  ENSURE(ctx->ck_pieces, pieces.size() * sizeof(ZbPiece));
  ENSURE(ctx->ck_first, first.size() * sizeof(uint32_t));
  ENSURE(ctx->ck_piece_out, pieces.size() * sizeof(ZbChunkCheck));
  ENSURE(ctx->ck_partials, pieces.size() * (size_t)ZB_CK_PARTIAL_BYTES);
  {
    int rc = ensure_group_events(ctx, 2 * ng + 2);
    if (rc) return rc;
    rc = ensure_pinned(ctx, n * (sizeof(uint64_t) + sizeof(int)) + 64);
    if (rc) return rc;
  }
  CK(cudaMemcpyAsync(ctx->src_off.p, reb.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->dst_off.p, dreb.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->ck_pieces.p, pieces.data(), pieces.size() * sizeof(ZbPiece), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->ck_first.p, first.data(), first.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->order.p, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  // the side streams start after whatever the caller's stream already holds (and after the tables above)
  CK(cudaEventRecord(ctx->gev[2 * ng], s));
  CK(cudaStreamWaitEvent(sh, ctx->gev[2 * ng], 0));
  CK(cudaStreamWaitEvent(sd, ctx->gev[2 * ng], 0));
  CK(cudaEventRecord(ctx->ev[6], sh));
  CK(cudaEventRecord(ctx->ev[8], sd));
  const bool src_pageable = is_pageable(h_src), dst_pageable = h_dst && is_pageable(h_dst);
  CK(cudaEventRecord(ctx->ev[0], s));
  if (gated) {
    // nothing that may synchronise the device (allocations, registrations) can happen while the gated kernel waits
    // for input: the staging rings of pageable callers are set up first
    if (src_pageable) {
      int rc = ring_ready(ctx, ctx->ring_in);
      if (rc) return rc;
    }
    if (dst_pageable) {
      int rc = ring_ready(ctx, ctx->ring_out);
      if (rc) return rc;
    }
    // gate words: [0] = groups copied in, [1 + g] = members of group g done, [1 + ng + g] = first queue position of group g
    std::vector<uint32_t> gate(2 * ng + 2, 0u);
    for (size_t gi = 0; gi <= ng; gi++) gate[1 + ng + gi] = (uint32_t)gb[gi];
    ENSURE(ctx->gate, gate.size() * sizeof(uint32_t));
    uint32_t *d_gate = (uint32_t *)ctx->gate.p;
    CK(cudaMemcpyAsync(d_gate, gate.data(), gate.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));  // `gate` is pageable: the copy has read it; and the zeroed words are in place before
                                   // the H2D stream's first write of d_gate[0]
    ZbInflateWork w;
    memset(&w, 0, sizeof(w));
    w.src = d_src;
    w.src_off = (const uint64_t *)ctx->src_off.p;
    w.dst = d_dst;
    w.dst_off = (const uint64_t *)ctx->dst_off.p;
    w.out_len = (uint64_t *)ctx->out_len.p;
    w.status = (int *)ctx->status.p;
    w.expect = (uint32_t *)ctx->expect.p;
    w.kind = (uint32_t *)ctx->kind.p;
    w.counter = (uint32_t *)ctx->counter.p + 16;
    w.tabs = ctx->d_tabs;
    w.n = (uint32_t)n;
    w.data_format = data_format;
    w.order = (const uint32_t *)ctx->order.p;
    w.gate_ready = d_gate;
    w.gate_done = d_gate + 1;
    w.gate_first = d_gate + 1 + ng;
    w.n_gates = (uint32_t)ng;
    CK(zb_launch_inflate(w, s));
    // from here on the kernel may be waiting for copies: if this function leaves early (a failed copy, a failed
    // enqueue), open every gate so that the kernel drains instead of waiting out its timeout
    struct GateRelease {
      zb200_ctx *ctx;
      uint32_t *word;
      bool armed;
      ~GateRelease() {
        if (armed) ctx->memops.write32((CUstream)ctx->h2d_stream, (CUdeviceptr)(uintptr_t)word, 0xffffffffu, 0);
      }
    } gate_release{ctx, d_gate, true};
    // gzip.nim:80-88 / zippy.nim:154-162: checksum, then size, of every member that inflated (after the whole
    // launch: 2.4 ms per 4 GiB, under the last groups' copy-out)
    ZbChecksumWork cw;
    memset(&cw, 0, sizeof(cw));
    cw.src = d_dst;
    cw.off = w.dst_off;
    cw.lens = w.out_len;
    cw.pieces = (const ZbPiece *)ctx->ck_pieces.p;
    cw.first = (const uint32_t *)ctx->ck_first.p;
    cw.piece_out = (ZbChunkCheck *)ctx->ck_piece_out.p;
    cw.partials = (uint32_t *)ctx->ck_partials.p;
    cw.status = w.status;
    cw.expect = w.expect;
    cw.kinds = w.kind;
    cw.isize_src = d_src;
    cw.isize_off = w.src_off;
    cw.tabs = ctx->d_tabs;
    cw.n = (uint32_t)n;
    cw.n_pieces = (uint32_t)pieces.size();
    CK(zb_launch_checksum(cw, s));
    ctx->timing.kernel_launches += 3;
    auto copy_in = [&](size_t gi) -> int {
      const uint64_t b0 = reb[gb[gi]], b1 = reb[gb[gi + 1]];
      int rc = h2d_copy(ctx, (uint8_t *)ctx->in_stage.p + b0, h_src + b0, (size_t)(b1 - b0), sh, src_pageable);
      if (rc) return rc;
      if (ctx->memops.write32((CUstream)sh, (CUdeviceptr)(uintptr_t)d_gate, (cuuint32_t)(gi + 1), 0) != CUDA_SUCCESS) return ZB200_ERR_CUDA;
      if (gi + 1 == ng) CK(cudaEventRecord(ctx->ev[7], sh));
      return ZB200_OK;
    };
    {
      int rc = copy_in(0);
      if (rc) return rc;
    }
    // ZB200_DEBUG_TIMELINE=1: per group, when its done count released the copy-out and when the copy ended (stderr)
    static const bool timeline = getenv("ZB200_DEBUG_TIMELINE") != nullptr;
    std::vector<cudaEvent_t> tl;
    if (timeline) {
      tl.resize(2 * ng + 1);
      for (auto &ev : tl) cudaEventCreate(&ev);
      cudaEventRecord(tl[2 * ng], sd);
    }
    for (size_t gi = 0; gi < ng; gi++) {
      if (gi + 1 < ng) {  // the next group's copy-in is queued before this thread may block on a pageable copy-out
        int rc = copy_in(gi + 1);
        if (rc) return rc;
      }
      const size_t m0 = gb[gi], m1 = gb[gi + 1];
      const uint64_t o0 = dreb[m0], o1 = dreb[m1];
      if (o1 > o0 && h_dst) {
        if (ctx->memops.wait32((CUstream)sd, (CUdeviceptr)(uintptr_t)(d_gate + 1 + gi), (cuuint32_t)(m1 - m0), CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
          return ZB200_ERR_CUDA;
        if (timeline) cudaEventRecord(tl[2 * gi], sd);
        int rc = d2h_copy(ctx, h_dst + o0, d_dst + o0, (size_t)(o1 - o0), sd, dst_pageable);
        if (rc) return rc;
        if (timeline) cudaEventRecord(tl[2 * gi + 1], sd);
      }
    }
    if (timeline) {
      cudaStreamSynchronize(sd);
      for (size_t gi = 0; gi < ng; gi++) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, tl[2 * ng], tl[2 * gi]);
        cudaEventElapsedTime(&b, tl[2 * ng], tl[2 * gi + 1]);
        fprintf(stderr, "group %zu: members %zu, out %.1f MiB, released %.2f ms, copied %.2f ms\n", gi, gb[gi + 1] - gb[gi],
                (double)(dreb[gb[gi + 1]] - dreb[gb[gi]]) / 1048576.0, a, b);
      }
      for (auto &ev : tl) cudaEventDestroy(ev);
    }
    gate_release.armed = false;   // every group's copy-in (and its gate word) is queued
  }
  for (size_t gi = 0; gi < ng && !gated; gi++) {
    const size_t m0 = gb[gi], m1 = gb[gi + 1], nm = m1 - m0;
    {
      const uint64_t b0 = reb[m0], b1 = reb[m1];
      int rc = h2d_copy(ctx, (uint8_t *)ctx->in_stage.p + b0, h_src + b0, (size_t)(b1 - b0), sh, src_pageable);
      if (rc) return rc;
      CK(cudaEventRecord(ctx->gev[2 * gi], sh));
      if (gi + 1 == ng) CK(cudaEventRecord(ctx->ev[7], sh));
    }
    CK(cudaStreamWaitEvent(s, ctx->gev[2 * gi], 0));
    ZbInflateWork w;
    memset(&w, 0, sizeof(w));
    w.src = d_src;
    w.src_off = (const uint64_t *)ctx->src_off.p + m0;
    w.dst = d_dst;
    w.dst_off = (const uint64_t *)ctx->dst_off.p + m0;
    w.out_len = (uint64_t *)ctx->out_len.p + m0;
    w.status = (int *)ctx->status.p + m0;
    w.expect = (uint32_t *)ctx->expect.p + m0;
    w.kind = (uint32_t *)ctx->kind.p + m0;
    w.counter = (uint32_t *)ctx->counter.p + 16 + gi;
    w.tabs = ctx->d_tabs;
    w.n = (uint32_t)nm;
    w.data_format = data_format;
    w.order = has_order[gi] ? (const uint32_t *)ctx->order.p + m0 : nullptr;
    CK(zb_launch_inflate(w, s));
    // gzip.nim:80-88 / zippy.nim:154-162: checksum, then size, of every member that inflated
    ZbChecksumWork cw;
    memset(&cw, 0, sizeof(cw));
    cw.src = d_dst;
    cw.off = w.dst_off;
    cw.lens = w.out_len;
    cw.pieces = (const ZbPiece *)ctx->ck_pieces.p + piece0[gi];
    cw.first = (const uint32_t *)ctx->ck_first.p + first0[gi];
    cw.piece_out = (ZbChunkCheck *)ctx->ck_piece_out.p + piece0[gi];
    cw.partials = (uint32_t *)ctx->ck_partials.p + piece0[gi] * (size_t)(ZB_CK_PARTIAL_BYTES / 4);
    cw.status = w.status;
    cw.expect = w.expect;
    cw.kinds = w.kind;
    cw.isize_src = d_src;
    cw.isize_off = w.src_off;
    cw.tabs = ctx->d_tabs;
    cw.n = (uint32_t)nm;
    cw.n_pieces = (uint32_t)(piece0[gi + 1] - piece0[gi]);
    CK(zb_launch_checksum(cw, s));
    CK(cudaEventRecord(ctx->gev[2 * gi + 1], s));
    CK(cudaStreamWaitEvent(sd, ctx->gev[2 * gi + 1], 0));
    const uint64_t o0 = dreb[m0], o1 = dreb[m1];
    if (o1 > o0 && h_dst) {
      int rc = d2h_copy(ctx, h_dst + o0, d_dst + o0, (size_t)(o1 - o0), sd, dst_pageable);
      if (rc) return rc;
    }
    ctx->timing.kernel_launches += 3;
  }
  CK(cudaEventRecord(ctx->ev[1], s));
  CK(cudaEventRecord(ctx->ev[9], sd));
  if (dst_pageable) {
    int rc = d2h_flush(ctx);
    if (rc) return rc;
  }
  uint64_t *pl = (uint64_t *)ctx->pin;
  int *ps = (int *)(pl + n);
  CK(cudaMemcpyAsync(pl, ctx->out_len.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(ps, ctx->status.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
  // the caller's stream ends after the last copy out
  CK(cudaEventRecord(ctx->gev[2 * ng + 1], sd));
  CK(cudaStreamWaitEvent(s, ctx->gev[2 * ng + 1], 0));
  CK(cudaStreamSynchronize(s));  // also: the planning vectors above stay alive until their uploads are done
  CK(cudaStreamSynchronize(sh));
  for (size_t i = 0; i < n; i++) {
    dst_lens[i] = ps[i] == ZB200_OK ? pl[i] : 0;
    if (statuses) statuses[i] = ps[i];
  }
  ctx->timing.inflate_ms = ev_ms(ctx->ev[0], ctx->ev[1]);  // inflate + verify of all groups (includes waits for copy-in)
  ctx->timing.verify_ms = 0.f;
  return ZB200_OK;
}

int checksum_device_locked(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n, int kind,
                           uint32_t *out) {
  if (kind != 0 && kind != 1) return ZB200_ERR_ARG;
  if (n == 0) return ZB200_OK;
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] < src_offsets[i]) return ZB200_ERR_ARG;
  ENSURE(ctx->src_off, (n + 1) * sizeof(uint64_t));
  ENSURE(ctx->ck_out, n * sizeof(uint32_t));
  cudaStream_t s = ctx->stream;
  CK(cudaMemcpyAsync(ctx->src_off.p, src_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
  ZbChecksumWork w;
  memset(&w, 0, sizeof(w));
  int rc = upload_pieces(ctx, src_offsets, n, w);
  if (rc) return rc;
  w.src = d_src;
  w.off = (const uint64_t *)ctx->src_off.p;
  w.out = (uint32_t *)ctx->ck_out.p;
  w.kind = kind;
  CK(cudaEventRecord(ctx->ev[0], s));
  CK(zb_launch_checksum(w, s));
  CK(cudaEventRecord(ctx->ev[1], s));
  CK(cudaMemcpyAsync(out, ctx->ck_out.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ctx->timing.checksum_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
  ctx->timing.kernel_launches += 2;
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

// An error return must not leave copies that use the caller's buffers in flight.
void quiesce(zb200_ctx *ctx) {
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->h2d_stream) cudaStreamSynchronize(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
  cudaGetLastError();
  // copies that were still waiting in the pinned ring belong to the failed call: their destinations are
  // the caller's buffers, which it may free now -- forget them
  for (int i = 0; i < kStageSlots; i++) ctx->ring_out.busy[i] = false;
}

// No C++ exception crosses the C ABI (std::vector growth on attacker-sized inputs, ...).
template <class F>
int guarded(zb200_ctx *ctx, F &&f) {
  int rc;
  try {
    rc = f();
  } catch (const std::bad_alloc &) {
    if (ctx) ctx->last_err = "host allocation failed";
    rc = ZB200_ERR_NOMEM;
  } catch (const std::exception &e) {
    if (ctx) ctx->last_err = e.what();
    rc = ZB200_ERR_ARG;
  }
  if (rc != ZB200_OK && ctx) quiesce(ctx);
  return rc;
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

static void zb_segv_handler(int sig) {
  void *frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "zippy_b200: fatal signal, backtrace (resolve with addr2line -e libzippy_b200.so):\n";
  if (write(2, msg, sizeof(msg) - 1) < 0) _exit(128 + sig);
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}

int zb200_init(int device, zb200_ctx **out) {
  if (!out) return ZB200_ERR_ARG;
  if (getenv("ZB200_DEBUG_SEGV")) signal(SIGSEGV, zb_segv_handler);  // debugging aid: where did a host fault happen
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
  if (const char *e = getenv("ZB200_GROUP_CHUNKS")) {  // test hook: force small launch groups
    long v = atol(e);
    if (v > 0) ctx->dev_group_chunks = ctx->host_group_chunks = (size_t)v;
  }
  if (const char *e = getenv("ZB200_BIG_MEMBER_BYTES")) {  // test hook: segment path for small members too
    long long v = atoll(e);
    if (v > 0) {
      ctx->big_member_bytes = (uint64_t)v;
      ctx->big_env = true;
    }
  }
  ctx->memops = load_stream_memops();
  if (const char *e = getenv("ZB200_UNC_GATED")) ctx->gated_unc = atoi(e) != 0;  // test hook: 0 = one launch per group
  if (const char *e = getenv("ZB200_UNC_GROUP_BYTES")) {  // test hook: small pipelined groups in the host uncompress
    long long v = atoll(e);
    if (v > 0) ctx->unc_group_out_bytes = (uint64_t)v;
  }
  if (const char *e = getenv("ZB200_DEV_GROUP_CHUNKS")) {  // device-resident batches only (bench.py)
    long v = atol(e);
    if (v > 0) ctx->dev_group_chunks = (size_t)v;
  }
  DeviceGuard g(device);
  bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  ctx->stream = ctx->own_stream;
  if (ok) ok = cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking) == cudaSuccess;
  if (ok) ok = cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; ok && i < 10; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  if (ok) ok = cudaMalloc((void **)&ctx->d_tabs, sizeof(ZbCrcTables)) == cudaSuccess;
  // kernel attributes are per device (and cheap to set again): every ctx sets them for its own
  if (ok) ok = zb_setup_deflate_attrs() == cudaSuccess && zb_setup_inflate_attrs() == cudaSuccess;
  if (ok) {
    ZbCrcTables t;
    zb_crc_build_tables(&t);
    ok = cudaMemcpy(ctx->d_tabs, &t, sizeof(t), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  if (!ok) {
    cudaGetLastError();
    zb200_shutdown(ctx);  // releases whatever was created
    return ZB200_ERR_CUDA;
  }
  *out = ctx;
  return ZB200_OK;
}

void zb200_shutdown(zb200_ctx *ctx) {
  if (!ctx) return;
  DeviceGuard g(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  DevBuf *bufs[] = {&ctx->desc, &ctx->member_first, &ctx->fname, &ctx->masks, &ctx->recs, &ctx->hist, &ctx->chk,
                    &ctx->cb, &ctx->chunk_off, &ctx->member_off, &ctx->member_check, &ctx->member_isize,
                    &ctx->src_off, &ctx->dst_off, &ctx->out_len, &ctx->status, &ctx->expect, &ctx->kind,
                    &ctx->counter, &ctx->ck_out, &ctx->ck_pieces, &ctx->ck_first, &ctx->ck_piece_out, &ctx->ck_partials, &ctx->in_stage, &ctx->out_stage, &ctx->lz2_tables,
                    &ctx->seg_src, &ctx->seg_dst, &ctx->seg_len, &ctx->seg_status, &ctx->seg_kind, &ctx->seg_expect, &ctx->seg_cand, &ctx->skip_mask, &ctx->order, &ctx->mark_scratch, &ctx->mark_segs, &ctx->seg_bits, &ctx->gate};
  for (DevBuf *b : bufs)
    if (b->p) cudaFree(b->p);
  if (ctx->d_tabs) cudaFree(ctx->d_tabs);
  for (int i = 0; i < 10; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (cudaEvent_t e : ctx->gev) cudaEventDestroy(e);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  ctx->pool.stop();
  for (StageRing *r : {&ctx->ring_in, &ctx->ring_out})
    for (int i = 0; i < kStageSlots; i++) {
      if (r->slot[i]) cudaFreeHost(r->slot[i]);
      if (r->ev[i]) cudaEventDestroy(r->ev[i]);
    }
  if (ctx->group_end.p) cudaFree(ctx->group_end.p);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  cudaGetLastError();
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
  return guarded(ctx, [&]() -> int {
  if (!ctx || !src_offsets || !dst_offsets || (n && (!d_src || !d_dst))) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  ctx->timing.kernel_launches = 0;
  return compress_locked(ctx, d_src, nullptr, src_offsets, n, level, data_format, fname_lens, d_dst, dst_cap, nullptr,
                         0, dst_offsets, statuses, ctx->dev_group_chunks);
  });
}

int zb200_compress_batch(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n, int level,
                         int data_format, const uint8_t *fname_lens, uint8_t *dst_base, size_t dst_cap,
                         uint64_t *dst_offsets, int *statuses) {
  return guarded(ctx, [&]() -> int {
  if (!ctx || !src_offsets || !dst_offsets || (n && (!src_base || !dst_base))) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  if (n == 0) {
    dst_offsets[0] = 0;
    return ZB200_OK;
  }
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] < src_offsets[i]) return ZB200_ERR_ARG;
  const uint64_t in_bytes = src_offsets[n] - src_offsets[0];
  uint64_t bound = 0;
  for (size_t i = 0; i < n; i++)
    bound += zb200_compress_bound((size_t)(src_offsets[i + 1] - src_offsets[i]), data_format) + 64;
  ENSURE(ctx->in_stage, (size_t)in_bytes + 64);
  ENSURE(ctx->out_stage, (size_t)bound + 64);
  int rc = compress_locked(ctx, (const uint8_t *)ctx->in_stage.p, src_base, src_offsets, n, level, data_format,
                           fname_lens, (uint8_t *)ctx->out_stage.p, ctx->out_stage.cap & ~(size_t)3, dst_base,
                           dst_cap, dst_offsets, statuses, ctx->host_group_chunks);
  if (rc) return rc;
  ctx->timing.h2d_ms = ev_ms(ctx->ev[6], ctx->ev[7]);
  ctx->timing.d2h_ms = ev_ms(ctx->ev[8], ctx->ev[9]);
  ctx->timing.h2d_bytes = in_bytes;
  ctx->timing.d2h_bytes = dst_offsets[n];
  return ZB200_OK;
  });
}

// host inputs (pipelined H2D over launch groups) -> members left in device memory: the sharded
// multi-GPU path compresses this way, exchanges the sizes, and only then knows where each shard's
// bytes go in the concatenated host stream
int zb200_compress_batch_h2d(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n, int level,
                             int data_format, const uint8_t *fname_lens, uint8_t *d_dst, size_t dst_cap,
                             uint64_t *dst_offsets, int *statuses) {
  return guarded(ctx, [&]() -> int {
    if (!ctx || !src_offsets || !dst_offsets || (n && (!src_base || !d_dst))) return ZB200_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    if (n == 0) {
      dst_offsets[0] = 0;
      return ZB200_OK;
    }
    for (size_t i = 0; i < n; i++)
      if (src_offsets[i + 1] < src_offsets[i]) return ZB200_ERR_ARG;
    const uint64_t in_bytes = src_offsets[n] - src_offsets[0];
    ENSURE(ctx->in_stage, (size_t)in_bytes + 64);
    int rc = compress_locked(ctx, (const uint8_t *)ctx->in_stage.p, src_base, src_offsets, n, level, data_format,
                             fname_lens, d_dst, dst_cap, nullptr, 0, dst_offsets, statuses, ctx->host_group_chunks);
    if (rc) return rc;
    ctx->timing.h2d_ms = ev_ms(ctx->ev[6], ctx->ev[7]);
    ctx->timing.h2d_bytes = in_bytes;
    return ZB200_OK;
  });
}

// second half of the sharded path: once the size exchange has told a rank where its shard lands in
// the concatenated stream, its device-resident members go straight to that place in host memory
int zb200_download(zb200_ctx *ctx, const uint8_t *d_src, uint8_t *h_dst, size_t bytes) {
  return guarded(ctx, [&]() -> int {
    if (!ctx || (bytes && (!d_src || !h_dst))) return ZB200_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (bytes) CK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZB200_OK;
  });
}

// page-lock a caller-owned host range so that the host-buffer calls can overlap their copies with
// the kernels (cudaHostRegister; a Nim string / malloc'd buffer is pageable otherwise)
int zb200_host_register(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return ZB200_ERR_ARG;
  if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) != cudaSuccess) {
    cudaGetLastError();
    return ZB200_ERR_CUDA;
  }
  return ZB200_OK;
}
int zb200_host_unregister(void *ptr) {
  if (!ptr) return ZB200_ERR_ARG;
  if (cudaHostUnregister(ptr) != cudaSuccess) {
    cudaGetLastError();
    return ZB200_ERR_CUDA;
  }
  return ZB200_OK;
}

int zb200_uncompress_batch_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                                  int data_format, uint8_t *d_dst, const uint64_t *dst_offsets, uint64_t *dst_lens,
                                  int *statuses) {
  return guarded(ctx, [&]() -> int {
  if (!ctx || !src_offsets || !dst_offsets || !dst_lens || (n && !d_src)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  ctx->timing.kernel_launches = 0;
  ctx->timing.inflate_ms = ctx->timing.verify_ms = 0.f;
  return uncompress_device_locked(ctx, d_src, src_offsets, n, data_format, 0, d_dst, dst_offsets, dst_lens, statuses,
                                  false);
  });
}

int zb200_uncompress_sizes_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n,
                                  int data_format, uint64_t *sizes, int *statuses) {
  return guarded(ctx, [&]() -> int {
  if (!ctx || !src_offsets || !sizes || (n && !d_src)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  ctx->timing.kernel_launches = 0;
  ctx->timing.inflate_ms = ctx->timing.verify_ms = 0.f;
  return uncompress_device_locked(ctx, d_src, src_offsets, n, data_format, 0, nullptr, nullptr, sizes, statuses,
                                  true);
  });
}

int zb200_uncompress_sizes(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                           int data_format, uint64_t *sizes, int *statuses) {
  return guarded(ctx, [&]() -> int {
  if (!ctx || !src_offsets || !sizes || (n && !src_base)) return ZB200_ERR_ARG;
  if (data_format < ZB200_DF_DETECT || data_format > ZB200_DF_DEFLATE) return ZB200_ERR_INVALID_FORMAT;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] < src_offsets[i]) return ZB200_ERR_ARG;
  // gzip members answer from their trailer (gzip.nim:66) after the same wrapper checks the device decoder
  // makes: nothing is copied to the device for them.  Only a batch that holds zlib / raw members needs the
  // counting pass.
  bool need_device = false;
  for (size_t i = 0; i < n && !need_device; i++) {
    uint64_t payload = 0;
    uint32_t kind = 0, expect = 0, isize = 0;
    const int st = zb_parse_wrapper(src_base + src_offsets[i], src_offsets[i + 1] - src_offsets[i], data_format, 0, payload,
                                    kind, expect, isize);
    if (st == ZB200_OK && kind != ZB200_DF_GZIP) need_device = true;
    sizes[i] = st == ZB200_OK ? isize : 0;
    if (statuses) statuses[i] = st;
  }
  if (!need_device) return ZB200_OK;
  std::vector<uint64_t> reb;
  int rc = stage_in(ctx, src_base, src_offsets, n, reb);
  if (rc) return rc;
  return uncompress_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data(), n, data_format, 0, nullptr,
                                  nullptr, sizes, statuses, true);
  });
}

int zb200_uncompress_batch(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n,
                           int data_format, uint8_t *dst_base, const uint64_t *dst_offsets, uint64_t *dst_lens,
                           int *statuses) {
  return guarded(ctx, [&]() -> int {
  if (!ctx || !src_offsets || !dst_offsets || !dst_lens || (n && !src_base)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  if (n == 0) return ZB200_OK;
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] < src_offsets[i] || dst_offsets[i + 1] < dst_offsets[i]) return ZB200_ERR_ARG;
  const uint64_t slo = src_offsets[0], shi = src_offsets[n], lo = dst_offsets[0], hi = dst_offsets[n];
  std::vector<uint64_t> reb(n + 1), dreb(n + 1);
  for (size_t i = 0; i <= n; i++) {
    reb[i] = src_offsets[i] - slo;
    dreb[i] = dst_offsets[i] - lo;
  }
  ENSURE(ctx->in_stage, (size_t)(shi - slo) + 64);
  ENSURE(ctx->out_stage, (size_t)(hi - lo) + 64);
  // Member groups: group g inflates while group g + 1 is copied in on the H2D stream and group
  // g - 1 is copied out on the D2H stream.  A launch wants ~7000 members to fill the GPU, so the
  // groups are big: an eighth of the batch, between 64 MiB and 1 GiB of output.  (The copies only
  // run asynchronously for page-locked host buffers; with pageable memory the same code is
  // correct but the copies block this thread.)
  std::vector<size_t> gb(1, 0);
  {
    uint64_t out_cap = ctx->unc_group_out_bytes;
    if (!out_cap) {
      // a member is decoded by ONE 8-lane group at ~9 MB/s, so every group's launch ends with the tail of its
      // longest member: a group must be worth several such tails (8192 x the largest output slot matches the
      // measured optimum of 512 MiB groups for 64 KiB members); otherwise an eighth of the batch
      uint64_t max_out = 0;
      for (size_t i = 0; i < n; i++) max_out = std::max<uint64_t>(max_out, dreb[i + 1] - dreb[i]);
      out_cap = std::max<uint64_t>(std::max<uint64_t>((hi - lo) / 8, 8192ull * max_out), 64ull << 20);
      // one gated launch for the whole batch has no per-group tail: the groups only set the granularity of the
      // overlap (what is exposed is the first group's copy-in and the last group's copy-out)
      if (ctx->memops.ok && ctx->gated_unc)
        out_cap = std::min<uint64_t>(std::max<uint64_t>((hi - lo) / 32, 32ull << 20), 256ull << 20);
    }
    const uint64_t in_cap = out_cap;
    size_t a = 0;
    for (size_t i = 1; i <= n; i++)
      if (i == n || dreb[i + 1] - dreb[a] > out_cap || reb[i + 1] - reb[a] > in_cap) {
        gb.push_back(i);
        a = i;
      }
  }
  const size_t ng = gb.size() - 1;
  {
    bool any_big = false;
    uint64_t big_thr = (n == 1 && !ctx->big_env) ? ctx->single_member_bytes : ctx->big_member_bytes;
    {
      // the same rule as inflate_big_members: with hundreds of large members the batch fills the GPU by itself
      size_t count = 0;
      for (size_t i = 0; i < n; i++) count += reb[i + 1] - reb[i] >= big_thr;
      if (count > 256) big_thr = std::max<uint64_t>(big_thr, 64ull << 20);
    }
    for (size_t i = 0; i < n && !any_big; i++) any_big = reb[i + 1] - reb[i] >= big_thr;
    if (!any_big) {
      int rc = uncompress_host_pipelined(ctx, src_base + slo, reb, n, data_format, dst_base ? dst_base + lo : nullptr, dreb,
                                         dst_lens, statuses, gb);
      if (rc) return rc;
      ctx->timing.h2d_ms = ev_ms(ctx->ev[6], ctx->ev[7]);
      ctx->timing.d2h_ms = ev_ms(ctx->ev[8], ctx->ev[9]);
      ctx->timing.h2d_bytes = shi - slo;
      ctx->timing.d2h_bytes = hi - lo;
      return ZB200_OK;
    }
  }
  // a batch with large members takes the group-by-group path: those members are planned on the host
  int rc = ensure_group_events(ctx, 2 * ng + 2);
  if (rc) return rc;
  cudaStream_t s = ctx->stream, sh = ctx->h2d_stream, sd = ctx->d2h_stream;
  // the side streams start after whatever the caller's stream already holds
  CK(cudaEventRecord(ctx->gev[2 * ng], s));
  CK(cudaStreamWaitEvent(sh, ctx->gev[2 * ng], 0));
  CK(cudaStreamWaitEvent(sd, ctx->gev[2 * ng], 0));
  CK(cudaEventRecord(ctx->ev[6], sh));
  CK(cudaEventRecord(ctx->ev[8], sd));
  auto copy_in = [&](size_t gi) -> int {
    const uint64_t b0 = reb[gb[gi]], b1 = reb[gb[gi + 1]];
    if (b1 > b0)
      CK(cudaMemcpyAsync((uint8_t *)ctx->in_stage.p + b0, src_base + slo + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, sh));
    CK(cudaEventRecord(ctx->gev[2 * gi], sh));
    return ZB200_OK;
  };
  rc = copy_in(0);
  if (rc) return rc;
  for (size_t gi = 0; gi < ng; gi++) {
    if (gi + 1 < ng) {
      rc = copy_in(gi + 1);
      if (rc) return rc;
    }
    CK(cudaStreamWaitEvent(s, ctx->gev[2 * gi], 0));
    const size_t m0 = gb[gi], m1 = gb[gi + 1];
    const std::function<int()> copy_out = [&]() -> int {
      CK(cudaEventRecord(ctx->gev[2 * gi + 1], s));
      CK(cudaStreamWaitEvent(sd, ctx->gev[2 * gi + 1], 0));
      const uint64_t b0 = dreb[m0], b1 = dreb[m1];
      if (b1 > b0 && dst_base)
        CK(cudaMemcpyAsync(dst_base + lo + b0, (uint8_t *)ctx->out_stage.p + b0, (size_t)(b1 - b0), cudaMemcpyDeviceToHost, sd));
      return ZB200_OK;
    };
    rc = uncompress_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data() + m0, m1 - m0, data_format, 0,
                                  (uint8_t *)ctx->out_stage.p, dreb.data() + m0, dst_lens + m0,
                                  statuses ? statuses + m0 : nullptr, false, &copy_out);
    if (rc) return rc;
  }
  CK(cudaEventRecord(ctx->ev[7], sh));
  CK(cudaEventRecord(ctx->ev[9], sd));
  // the caller's stream ends after the last copy out
  CK(cudaEventRecord(ctx->gev[2 * ng + 1], sd));
  CK(cudaStreamWaitEvent(s, ctx->gev[2 * ng + 1], 0));
  CK(cudaStreamSynchronize(sd));
  CK(cudaStreamSynchronize(sh));
  ctx->timing.h2d_ms = ev_ms(ctx->ev[6], ctx->ev[7]);
  ctx->timing.d2h_ms = ev_ms(ctx->ev[8], ctx->ev[9]);
  ctx->timing.h2d_bytes = shi - slo;
  ctx->timing.d2h_bytes = hi - lo;
  return ZB200_OK;
  });
}

int zb200_checksum_batch_device(zb200_ctx *ctx, const uint8_t *d_src, const uint64_t *src_offsets, size_t n, int kind,
                                uint32_t *out) {
  return guarded(ctx, [&]() -> int {
  if (!ctx || !src_offsets || !out || (n && !d_src)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  ctx->timing.kernel_launches = 0;
  return checksum_device_locked(ctx, d_src, src_offsets, n, kind, out);
  });
}

int zb200_checksum_batch(zb200_ctx *ctx, const uint8_t *src_base, const uint64_t *src_offsets, size_t n, int kind,
                         uint32_t *out) {
  return guarded(ctx, [&]() -> int {
  if (!ctx || !src_offsets || !out || (n && !src_base)) return ZB200_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  memset(&ctx->timing, 0, sizeof(ctx->timing));
  std::vector<uint64_t> reb;
  int rc = stage_in(ctx, src_base, src_offsets, n, reb);
  if (rc) return rc;
  return checksum_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data(), n, kind, out);
  });
}

// ---- one input of unknown size, decoded ONCE ----
// The reference's inflate appends to a string that grows as it goes (inflate.nim:268-291); a fixed-capacity
// ABI would otherwise need a counting pass before the real one.  decode_begin inflates into library-owned
// device memory (capacity from the gzip trailer, else a guess that a counting pass corrects only when it
// was too small) and reports the size; decode_finish copies the bytes to the caller.  Also what gives the
// reference's answer for a gzip member whose ISIZE understates its content: the data is produced, the CRC
// is checked, then the size check fails (gzip.nim:80-88), instead of "destination too small".
int zb200_decode_begin(zb200_ctx *ctx, const uint8_t *src, size_t len, int data_format, size_t pos, size_t *out_len) {
  return guarded(ctx, [&]() -> int {
    if (!ctx || !out_len || (len && !src)) return ZB200_ERR_ARG;
    if (data_format < ZB200_DF_DETECT || data_format > ZB200_DF_DEFLATE) return ZB200_ERR_INVALID_FORMAT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    ctx->pending = false;
    uint8_t dummy = 0;
    const uint8_t *sp = src ? src : &dummy;
    uint64_t payload = 0;
    uint32_t kind = 0, expect = 0, isize = 0;
    int st = zb_parse_wrapper(sp, len, data_format, pos, payload, kind, expect, isize);
    if (st != ZB200_OK) return st;
    uint64_t cap = kind == ZB200_DF_GZIP ? std::min<uint64_t>(isize, (uint64_t)len * 1032ull + 1024ull)
                                         : std::min<uint64_t>(std::max<uint64_t>((uint64_t)len * 8ull, 256ull << 10), 1ull << 30);
    uint64_t so[2] = {0, len};
    std::vector<uint64_t> reb;
    int rc = stage_in(ctx, sp, so, 1, reb);
    if (rc) return rc;
    for (int attempt = 0; attempt < 2; attempt++) {
      ENSURE(ctx->out_stage, (size_t)cap + 64);
      uint64_t dof[2] = {0, cap}, dl = 0;
      rc = uncompress_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data(), 1, data_format, pos,
                                    (uint8_t *)ctx->out_stage.p, dof, &dl, &st, false);
      if (rc) return rc;
      if (st != ZB200_ERR_DST_TOO_SMALL || attempt == 1) {
        if (st != ZB200_OK) return st;
        ctx->pending = true;
        ctx->pending_len = dl;
        *out_len = (size_t)dl;
        return ZB200_OK;
      }
      // too small: count the raw stream from the payload start (a gzip ISIZE is a claim, not a fact)
      int cst = ZB200_OK;
      uint64_t real = 0;
      rc = uncompress_device_locked(ctx, (const uint8_t *)ctx->in_stage.p, reb.data(), 1, ZB200_DF_DEFLATE, payload, nullptr,
                                    nullptr, &real, &cst, true);
      if (rc) return rc;
      if (cst != ZB200_OK) return cst;
      cap = real;
    }
    return ZB200_ERR_UNCOMPRESS;
  });
}

int zb200_decode_finish(zb200_ctx *ctx, uint8_t *dst, size_t dst_cap, size_t *dst_len) {
  return guarded(ctx, [&]() -> int {
    if (!ctx || !dst_len) return ZB200_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    DeviceGuard g(ctx->device);
    if (!ctx->pending) return ZB200_ERR_ARG;
    if (ctx->pending_len > dst_cap) return ZB200_ERR_DST_TOO_SMALL;
    if (ctx->pending_len && !dst) return ZB200_ERR_ARG;
    if (ctx->pending_len) {
      CK(cudaMemcpyAsync(dst, ctx->out_stage.p, (size_t)ctx->pending_len, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
    *dst_len = (size_t)ctx->pending_len;
    ctx->pending = false;
    return ZB200_OK;
  });
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
  return guarded(ctx, [&]() -> int {
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
  });
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
