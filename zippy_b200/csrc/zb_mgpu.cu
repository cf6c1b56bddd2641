// zb_mgpu.cu -- several GPUs behind ONE call of the C ABI (SURVEY 8e at the drop-in boundary).
//
// Independent members shard by contiguous index range, balanced by input bytes; there is no
// data-path collective.  One host thread drives each device through its own zb200_ctx:
//   compress  : shard g: host inputs -> device g (H2D pipelined with the kernels), members stay there;
//               the per-shard sizes are gathered (in this single-process form the "all-gather" is a
//               read of host memory -- the multi-process form, one rank per GPU, exchanges them with one
//               NCCL all_gather: zippy_b200/sharding.py, bench.py config 5); every shard's bytes are then
//               copied straight to their place in ONE concatenated host stream.
//   uncompress: members are independent and their output slots are the caller's, so each shard is simply
//               the host-buffer call on its own device.
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zippy_b200.h"

struct zb200_mgpu {
  std::vector<zb200_ctx *> ctx;
  std::vector<int> dev;
  std::vector<void *> d_out;     // per device: compressed members of its shard
  std::vector<size_t> d_out_cap;
  std::string last_err;
};

namespace {

// contiguous ranges balanced by bytes: cut[g]..cut[g+1]
std::vector<size_t> shard_cuts(const uint64_t *offs, size_t n, size_t parts) {
  std::vector<size_t> cut(parts + 1, n);
  cut[0] = 0;
  const uint64_t total = offs[n] - offs[0];
  size_t i = 0;
  for (size_t g = 1; g < parts; g++) {
    const uint64_t want = offs[0] + total * g / parts;
    while (i < n && offs[i] < want) i++;
    cut[g] = std::max(cut[g - 1], i);
  }
  return cut;
}

}  // namespace

extern "C" {

int zb200_mgpu_init(const int *devices, int n_devices, zb200_mgpu **out) {
  if (!out) return ZB200_ERR_ARG;
  *out = nullptr;
  int have = zb200_device_count();
  if (have <= 0) return ZB200_ERR_CUDA;
  zb200_mgpu *m = new zb200_mgpu();
  if (!devices || n_devices <= 0) {
    for (int d = 0; d < have; d++) m->dev.push_back(d);
  } else {
    for (int i = 0; i < n_devices; i++) m->dev.push_back(devices[i]);
  }
  for (int d : m->dev) {
    zb200_ctx *c = nullptr;
    int rc = zb200_init(d, &c);
    if (rc) {
      zb200_mgpu_shutdown(m);
      return rc;
    }
    m->ctx.push_back(c);
    m->d_out.push_back(nullptr);
    m->d_out_cap.push_back(0);
  }
  *out = m;
  return ZB200_OK;
}

void zb200_mgpu_shutdown(zb200_mgpu *m) {
  if (!m) return;
  for (size_t g = 0; g < m->ctx.size(); g++) {
    if (m->d_out[g]) {
      cudaSetDevice(m->dev[g]);
      cudaFree(m->d_out[g]);
    }
    zb200_shutdown(m->ctx[g]);
  }
  delete m;
}

int zb200_mgpu_device_count(zb200_mgpu *m) { return m ? (int)m->ctx.size() : 0; }

int zb200_mgpu_compress_batch(zb200_mgpu *m, const uint8_t *src_base, const uint64_t *src_offsets, size_t n, int level,
                              int data_format, const uint8_t *fname_lens, uint8_t *dst_base, size_t dst_cap,
                              uint64_t *dst_offsets, int *statuses) {
  if (!m || m->ctx.empty() || !src_offsets || !dst_offsets || (n && (!src_base || !dst_base))) return ZB200_ERR_ARG;
  dst_offsets[0] = 0;
  if (n == 0) return ZB200_OK;
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] < src_offsets[i]) return ZB200_ERR_ARG;
  const size_t G = m->ctx.size();
  const std::vector<size_t> cut = shard_cuts(src_offsets, n, G);
  std::vector<int> rc(G, ZB200_OK);
  std::vector<std::vector<uint64_t>> local(G);   // member offsets inside each shard's device buffer
  // ---- phase 1: every shard compresses on its own device ----
  {
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++)
      th.emplace_back([&, g] {
        try {
          const size_t lo = cut[g], cnt = cut[g + 1] - cut[g];
          local[g].assign(cnt + 1, 0);
          if (!cnt) return;
          size_t bound = 4096;
          for (size_t i = lo; i < lo + cnt; i++) bound += zb200_compress_bound((size_t)(src_offsets[i + 1] - src_offsets[i]), data_format) + 64;
          bound = (bound + 3) & ~(size_t)3;
          if (cudaSetDevice(m->dev[g]) != cudaSuccess) {
            rc[g] = ZB200_ERR_CUDA;
            return;
          }
          if (bound > m->d_out_cap[g]) {
            if (m->d_out[g]) cudaFree(m->d_out[g]);
            m->d_out[g] = nullptr;
            m->d_out_cap[g] = 0;
            if (cudaMalloc(&m->d_out[g], bound) != cudaSuccess) {
              cudaGetLastError();
              rc[g] = ZB200_ERR_NOMEM;
              return;
            }
            m->d_out_cap[g] = bound;
          }
          rc[g] = zb200_compress_batch_h2d(m->ctx[g], src_base, src_offsets + lo, cnt, level, data_format,
                                           fname_lens ? fname_lens + lo : nullptr, (uint8_t *)m->d_out[g], m->d_out_cap[g],
                                           local[g].data(), statuses ? statuses + lo : nullptr);
        } catch (...) {
          rc[g] = ZB200_ERR_NOMEM;
        }
      });
    for (std::thread &t : th) t.join();
  }
  for (size_t g = 0; g < G; g++)
    if (rc[g]) return rc[g];
  // ---- the size exchange: where does every shard land in the concatenated stream? ----
  std::vector<uint64_t> shard_off(G + 1, 0);
  for (size_t g = 0; g < G; g++) shard_off[g + 1] = shard_off[g] + local[g].back();
  if (shard_off[G] > dst_cap) return ZB200_ERR_DST_TOO_SMALL;
  for (size_t g = 0; g < G; g++)
    for (size_t i = cut[g]; i < cut[g + 1]; i++) dst_offsets[i + 1] = shard_off[g] + local[g][i - cut[g] + 1];
  // ---- phase 2: every shard goes to its place ----
  {
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++)
      th.emplace_back([&, g] {
        const uint64_t bytes = local[g].back();
        if (bytes) rc[g] = zb200_download(m->ctx[g], (const uint8_t *)m->d_out[g], dst_base + shard_off[g], (size_t)bytes);
      });
    for (std::thread &t : th) t.join();
  }
  for (size_t g = 0; g < G; g++)
    if (rc[g]) return rc[g];
  return ZB200_OK;
}

int zb200_mgpu_uncompress_batch(zb200_mgpu *m, const uint8_t *src_base, const uint64_t *src_offsets, size_t n, int data_format,
                                uint8_t *dst_base, const uint64_t *dst_offsets, uint64_t *dst_lens, int *statuses) {
  if (!m || m->ctx.empty() || !src_offsets || !dst_offsets || !dst_lens || (n && !src_base)) return ZB200_ERR_ARG;
  if (n == 0) return ZB200_OK;
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] < src_offsets[i] || dst_offsets[i + 1] < dst_offsets[i]) return ZB200_ERR_ARG;
  const size_t G = m->ctx.size();
  const std::vector<size_t> cut = shard_cuts(dst_offsets, n, G);   // balance by output bytes
  std::vector<int> rc(G, ZB200_OK);
  std::vector<std::thread> th;
  for (size_t g = 0; g < G; g++)
    th.emplace_back([&, g] {
      const size_t lo = cut[g], cnt = cut[g + 1] - cut[g];
      if (cnt)
        rc[g] = zb200_uncompress_batch(m->ctx[g], src_base, src_offsets + lo, cnt, data_format, dst_base, dst_offsets + lo,
                                       dst_lens + lo, statuses ? statuses + lo : nullptr);
    });
  for (std::thread &t : th) t.join();
  for (size_t g = 0; g < G; g++)
    if (rc[g]) return rc[g];
  return ZB200_OK;
}

int zb200_mgpu_checksum_batch(zb200_mgpu *m, const uint8_t *src_base, const uint64_t *src_offsets, size_t n, int kind,
                              uint32_t *out) {
  if (!m || m->ctx.empty() || !src_offsets || !out || (n && !src_base)) return ZB200_ERR_ARG;
  if (n == 0) return ZB200_OK;
  for (size_t i = 0; i < n; i++)
    if (src_offsets[i + 1] < src_offsets[i]) return ZB200_ERR_ARG;
  const size_t G = m->ctx.size();
  const std::vector<size_t> cut = shard_cuts(src_offsets, n, G);
  std::vector<int> rc(G, ZB200_OK);
  std::vector<std::thread> th;
  for (size_t g = 0; g < G; g++)
    th.emplace_back([&, g] {
      const size_t lo = cut[g], cnt = cut[g + 1] - cut[g];
      if (cnt) rc[g] = zb200_checksum_batch(m->ctx[g], src_base, src_offsets + lo, cnt, kind, out + lo);
    });
  for (std::thread &t : th) t.join();
  for (size_t g = 0; g < G; g++)
    if (rc[g]) return rc[g];
  return ZB200_OK;
}

}  // extern "C"
