// zb_deflate.cu -- the compress pipeline for sm_100a:
//   k_lz    : per-chunk LZ77 parse (one CTA per 64 KiB chunk staged in shared memory by
//             TMA; each warp parses an 8 KiB sub-chunk with a private hash table),
//             fused with the per-sub-chunk symbol histograms and the CRC-32/Adler-32
//             of the chunk.  Replaces encodeSnappy/encodeFragment (snappy.nim:12-163),
//             the histogram side of BlockMetadata (internal.nim:128-131) and the
//             separate crc32/adler32 passes of zippy.nim:47,73.
//   k_huff  : one thread per chunk: zb_build_codebook (zb_huff.h) -- replaces
//             huffmanCodes + the dynamic header writer (deflate.nim:13-151, 295-394)
//             and the stored/fixed/dynamic choice (deflate.nim:274-290).
//   k_scan  : exclusive scan of chunk sizes -> output offsets; per-member checksum combine.
//   k_pack  : token -> bit emission with exact, precomputed bit offsets (replaces the
//             BitStreamWriter loop, deflate.nim:396-464, bitstreams.nim:84-123) plus the
//             gzip/zlib framing bytes of zippy.nim:21-78 for batched members.
#include "zb_device.cuh"
#include "zb_kernels.h"

#define LZ_THREADS (ZB_WARPS_PER_CHUNK * 32)
#define LZ_HASH_BITS 11
#define LZ_TABLE_ENTRIES (1 << LZ_HASH_BITS)
#define LZ_PRESEED 4096
#define LZ_PAR_CAP 32  // bytes a lane extends on its own; longer matches finish cooperatively

// shared-memory layout of k_lz (bytes)
#define LZ_SM_DATA 0
#define LZ_SM_DATA_BYTES (ZB_CHUNK_BYTES + 64)
#define LZ_SM_TABLE (LZ_SM_DATA + LZ_SM_DATA_BYTES)
#define LZ_SM_TABLE_BYTES (ZB_WARPS_PER_CHUNK * LZ_TABLE_ENTRIES * 2)
#define LZ_SM_HIST (LZ_SM_TABLE + LZ_SM_TABLE_BYTES)
#define LZ_SM_HIST_BYTES (ZB_WARPS_PER_CHUNK * ZB_HIST_WORDS * 4)
#define LZ_SM_CRC (LZ_SM_HIST + LZ_SM_HIST_BYTES)
#define LZ_SM_CRC_BYTES (4 * 256 * 4 + 36 * 4)
#define LZ_SM_PART (LZ_SM_CRC + LZ_SM_CRC_BYTES)
#define LZ_SM_PART_BYTES (ZB_WARPS_PER_CHUNK * 24)
#define LZ_SM_BAR (LZ_SM_PART + LZ_SM_PART_BYTES)
#define LZ_SM_TOTAL (LZ_SM_BAR + 16)

__device__ __forceinline__ uint32_t lz_hash(uint32_t v) { return (v * 0x9E3779B1u) >> (32 - LZ_HASH_BITS); }

template <int MODE>  // 1: hash-table matcher (level 1 and, for now, the LZ levels); 0: literals only
__global__ void __launch_bounds__(LZ_THREADS, 2)
    k_lz(const uint8_t *__restrict__ src, const ZbChunkDesc *__restrict__ desc, uint2 *__restrict__ masks,
         uint32_t *__restrict__ recs, uint16_t *__restrict__ hist, ZbChunkCheck *__restrict__ chk,
         const ZbCrcTables *__restrict__ tabs) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *data = smem + LZ_SM_DATA;
  uint16_t *table_all = reinterpret_cast<uint16_t *>(smem + LZ_SM_TABLE);
  uint32_t *hist_all = reinterpret_cast<uint32_t *>(smem + LZ_SM_HIST);
  uint32_t *crc_tab = reinterpret_cast<uint32_t *>(smem + LZ_SM_CRC);
  uint32_t *lane_mul = crc_tab + 1024;
  uint64_t *part = reinterpret_cast<uint64_t *>(smem + LZ_SM_PART);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + LZ_SM_BAR);

  const uint32_t chunk = blockIdx.x;
  const ZbChunkDesc d = desc[chunk];
  const uint32_t len = d.len;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    zb_mbar_init(bar, 1);
    zb_fence_mbar_init();
  }
  __syncthreads();
  uint32_t mis = (uint32_t)((uintptr_t)(src + d.src_off) & 15u);
  if (tid == 0 && len) zb_stage_chunk(data, src + d.src_off, len, bar);

  // while the bulk copy is in flight: clear histograms, empty the hash tables, load CRC tables
  for (int i = tid; i < ZB_WARPS_PER_CHUNK * ZB_HIST_WORDS; i += LZ_THREADS) hist_all[i] = 0;
  {
    uint4 ff = make_uint4(~0u, ~0u, ~0u, ~0u);
    uint4 *t4 = reinterpret_cast<uint4 *>(table_all);
    for (int i = tid; i < LZ_SM_TABLE_BYTES / 16; i += LZ_THREADS) t4[i] = ff;
  }
  for (int i = tid; i < 1024; i += LZ_THREADS) crc_tab[i] = (&tabs->mul1024[0][0])[i];
  if (tid < 33) lane_mul[tid] = tabs->lane_mul[tid];
  __syncthreads();
  if (len) zb_mbar_wait(bar, 0);

  const uint32_t b0 = (uint32_t)warp * ZB_SUB_BYTES;
  const uint32_t b1 = min(b0 + ZB_SUB_BYTES, len);
  uint16_t *table = table_all + warp * LZ_TABLE_ENTRIES;
  uint32_t *whist = hist_all + warp * ZB_HIST_WORDS;

  // ---- checksums of this warp's piece (CRC raw + Adler sums) ----
  {
    ZbCheck c;
    c.crc_raw = 0;
    c.a_sum = c.b_sum = 0;
    uint32_t n = b0 < len ? b1 - b0 : 0;
    if (n) c = zb_warp_checksums(data, mis + b0, n, crc_tab, lane_mul);
    if (lane == 0) {
      part[warp * 3 + 0] = c.crc_raw;
      part[warp * 3 + 1] = c.a_sum;
      part[warp * 3 + 2] = c.b_sum;
    }
  }

  if (b0 < len) {
    if (MODE == 1 && warp > 0) {
      // pre-seed the private table with the positions just before this sub-chunk
      for (uint32_t s = b0 - LZ_PRESEED; s < b0; s += 32) {
        uint32_t p = s + (uint32_t)lane;
        if (p + 4 <= len) table[lz_hash(zb_ld32_unaligned(data, mis + p))] = (uint16_t)p;
      }
      __syncwarp();
    }
    uint32_t entry = b0;
    const size_t win_base = (size_t)chunk * ZB_WINDOWS_PER_CHUNK;
    for (uint32_t wb = b0; wb < b1; wb += 32) {
      const uint32_t win = wb >> 5;
      if (entry >= wb + 32) {  // whole window already covered by a long match
        if (lane == 0) masks[win_base + win] = make_uint2(0u, 0u);
        continue;
      }
      const uint32_t p = wb + (uint32_t)lane;
      const uint32_t nvalid = min(32u, b1 - wb);
      const uint32_t v = zb_ld32_unaligned(data, mis + p);
      uint32_t m = 0, c = 0;
      if (MODE == 1) {
        const bool can = (p + 4 <= len);
        const uint32_t h = lz_hash(v);
        c = table[h];
        __syncwarp();
        if (can) table[h] = (uint16_t)p;
        // a match may not cross the sub-chunk end (the next warp starts its own parse there)
        const uint32_t limit = p < b1 ? min((uint32_t)ZB_MAX_MATCH, b1 - p) : 0u;
        if (can && c < p && p - c <= ZB_MAX_DIST && p >= entry && limit >= ZB_MIN_MATCH) {
          if (zb_ld32_unaligned(data, mis + c) == v) {
            m = 4;
            while (m < LZ_PAR_CAP) {
              uint32_t x = zb_ld32_unaligned(data, mis + p + m) ^ zb_ld32_unaligned(data, mis + c + m);
              if (x) {
                m += (uint32_t)(__ffs((int)x) - 1) >> 3;
                break;
              }
              m += 4;
            }
            if (m < LZ_PAR_CAP) m = min(m, limit);  // a capped match is clamped after extension
          }
        }
      }
      // ---- greedy selection inside the window (uniform control flow) ----
      uint32_t mm = __ballot_sync(ZB_FULL, m != 0);
      uint32_t sel = 0, ism = 0, cur = entry - wb;
      uint32_t my_len = 0;
      while (cur < nvalid) {
        uint32_t rest = mm >> cur;
        uint32_t upto = rest ? cur + (uint32_t)(__ffs((int)rest) - 1) : nvalid;
        // literals [cur, upto)
        uint32_t lit_bits = (upto >= 32 ? ~0u : ((1u << upto) - 1u)) & ~((1u << cur) - 1u);
        sel |= lit_bits;
        if (!rest || upto >= nvalid) {
          cur = nvalid;
          break;
        }
        const uint32_t nx = upto;
        uint32_t mlen = __shfl_sync(ZB_FULL, m, (int)nx);
        if (mlen >= LZ_PAR_CAP) {
          // cooperative extension: lane j checks bytes [32 + 8j, 40 + 8j) of the match
          const uint32_t mc = __shfl_sync(ZB_FULL, c, (int)nx);
          const uint32_t pos = wb + nx;
          const uint32_t off = LZ_PAR_CAP + 8u * (uint32_t)lane;
          uint32_t x0 = zb_ld32_unaligned(data, mis + pos + off) ^ zb_ld32_unaligned(data, mis + mc + off);
          uint32_t x1 = zb_ld32_unaligned(data, mis + pos + off + 4) ^ zb_ld32_unaligned(data, mis + mc + off + 4);
          uint32_t nm = x0 ? ((uint32_t)(__ffs((int)x0) - 1) >> 3) : 4u + (x1 ? ((uint32_t)(__ffs((int)x1) - 1) >> 3) : 4u);
          uint32_t stop = __ballot_sync(ZB_FULL, nm < 8u);
          if (stop) {
            int first = __ffs((int)stop) - 1;
            mlen = LZ_PAR_CAP + 8u * (uint32_t)first + __shfl_sync(ZB_FULL, nm, first);
          } else {
            mlen = LZ_PAR_CAP + 256u;
          }
          mlen = min(mlen, min((uint32_t)ZB_MAX_MATCH, b1 - pos));
        }
        sel |= 1u << nx;
        ism |= 1u << nx;
        if ((uint32_t)lane == nx) my_len = mlen;
        cur = nx + mlen;
      }
      entry = wb + cur;

      // ---- per-warp histogram + token records ----
      const bool is_sel = (sel >> lane) & 1u, is_m = (ism >> lane) & 1u;
      if (is_m) {
        uint32_t dist = p - c;
        uint32_t s1 = 257u + (uint32_t)zb_len_code(my_len);
        uint32_t s2 = (uint32_t)ZB_NUM_LITLEN + (uint32_t)zb_dist_code(dist);
        atomicAdd(&whist[s1 >> 1], 1u << ((s1 & 1u) * 16u));
        atomicAdd(&whist[s2 >> 1], 1u << ((s2 & 1u) * 16u));
        uint32_t rank = (uint32_t)__popc(ism & ((1u << lane) - 1u));
        recs[(win_base + win) * ZB_MATCH_SLOTS + rank] = (my_len - 3u) | ((dist - 1u) << 9);
      } else if (is_sel) {
        uint32_t s = v & 255u;
        atomicAdd(&whist[s >> 1], 1u << ((s & 1u) * 16u));
      }
      if (lane == 0) masks[win_base + win] = make_uint2(sel, ism);
    }
  }
  __syncthreads();
  // ---- publish histograms (packed u16 pairs == the global u16 layout) and chunk checksums ----
  {
    uint32_t *gh = reinterpret_cast<uint32_t *>(hist + (size_t)chunk * ZB_WARPS_PER_CHUNK * ZB_HIST_SYMS);
    for (int i = tid; i < ZB_WARPS_PER_CHUNK * ZB_HIST_WORDS; i += LZ_THREADS) gh[i] = hist_all[i];
  }
  if (tid == 0) {
    uint32_t raw = 0;
    uint64_t a = 0, b = 0;
    for (int w = 0; w < ZB_WARPS_PER_CHUNK; w++) {
      uint32_t wb0 = (uint32_t)w * ZB_SUB_BYTES;
      if (wb0 >= len) break;
      uint32_t wb1 = min(wb0 + ZB_SUB_BYTES, len);
      uint32_t after = len - wb1;
      uint32_t r = (uint32_t)part[w * 3 + 0];
      raw ^= after ? zb_gf2_mul(r, zb_xpow8(after)) : r;
      a += part[w * 3 + 1];
      b += part[w * 3 + 2] + (uint64_t)after * part[w * 3 + 1];
    }
    ZbChunkCheck cc;
    cc.crc_raw = raw;
    cc.adler = zb_adler_from_sums(a % ZB_ADLER_MOD, b % ZB_ADLER_MOD, len);
    chk[chunk] = cc;
  }
}

// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
    k_huff(const ZbChunkDesc *__restrict__ desc, const uint16_t *__restrict__ hist, ZbCodebook *__restrict__ cb,
           uint32_t n_chunks, int level) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  ZbChunkDesc d = desc[c];
  zb_build_codebook(hist + (size_t)c * ZB_WARPS_PER_CHUNK * ZB_HIST_SYMS, d.len, (d.flags & ZB_CHUNK_LAST) ? 1 : 0,
                    level == 0 ? 0 : -1, &cb[c]);
}

// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t frame_head_bytes(int fmt, const uint8_t *fname_len, uint32_t m) {
  if (fmt == ZB_DF_GZIP) return 10u + (fname_len ? (uint32_t)fname_len[m] : 0u) + 1u;
  if (fmt == ZB_DF_ZLIB) return 2u;
  return 0u;
}
__device__ __forceinline__ uint32_t frame_tail_bytes(int fmt) {
  return fmt == ZB_DF_GZIP ? 8u : fmt == ZB_DF_ZLIB ? 4u : 0u;
}

#define SCAN_THREADS 1024
__global__ void __launch_bounds__(SCAN_THREADS)
    k_scan(ZbCompressWork w) {
  __shared__ uint64_t warp_tot[32];
  __shared__ uint64_t carry_s;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = w.out_base;
  __syncthreads();
  for (uint32_t base = 0; base < w.n_chunks; base += SCAN_THREADS) {
    uint32_t c = base + (uint32_t)tid;
    uint64_t sz = 0;
    uint32_t head = 0, flags = 0, member = 0;
    if (c < w.n_chunks) {
      ZbChunkDesc d = w.desc[c];
      flags = d.flags;
      member = d.member;
      sz = w.cb[c].total_bytes;
      if (flags & ZB_CHUNK_FIRST) {
        head = frame_head_bytes(w.data_format, w.fname_len, member);
        sz += head;
      }
      if (flags & ZB_CHUNK_LAST) sz += frame_tail_bytes(w.data_format);
    }
    uint64_t incl = sz;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t t = __shfl_up_sync(ZB_FULL, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint64_t t = warp_tot[lane], it = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint64_t u = __shfl_up_sync(ZB_FULL, it, o);
        if (lane >= o) it += u;
      }
      warp_tot[lane] = it - t;  // exclusive
    }
    __syncthreads();
    uint64_t carry = carry_s;
    uint64_t excl = carry + warp_tot[warp] + incl - sz;
    if (c < w.n_chunks) {
      w.chunk_off[c] = excl + head;
      if (flags & ZB_CHUNK_FIRST) w.member_off[member] = excl;
    }
    __syncthreads();
    if (tid == SCAN_THREADS - 1) carry_s = excl + sz;
    __syncthreads();
  }
  if (tid == 0) w.member_off[w.n_members] = carry_s;
  // whole-member checksums: sequential combine over each member's chunks
  for (uint32_t m = (uint32_t)tid; m < w.n_members; m += SCAN_THREADS) {
    uint32_t c0 = w.member_first[m], c1 = w.member_first[m + 1];
    uint32_t raw = 0, ad = 1;
    uint64_t total = 0;
    for (uint32_t c = c0; c < c1; c++) {
      uint32_t l = w.desc[c].len;
      ZbChunkCheck cc = w.chk[c];
      if (w.data_format == ZB_DF_ZLIB) {
        ad = zb_adler32_combine(ad, cc.adler, l);
      } else {
        raw = (c == c0) ? cc.crc_raw : (zb_gf2_mul(raw, zb_xpow8(l)) ^ cc.crc_raw);
      }
      total += l;
    }
    w.member_check[m] = (w.data_format == ZB_DF_ZLIB) ? ad : zb_crc32_finalize(raw, total);
    w.member_isize[m] = (uint32_t)total;
  }
}

// ------------------------------------------------------------------------------------
#define PK_SM_DATA 0
#define PK_SM_DATA_BYTES (ZB_CHUNK_BYTES + 64)
#define PK_SM_CODES (PK_SM_DATA + PK_SM_DATA_BYTES)
#define PK_SM_CODES_BYTES ((288 + 32) * 4)
#define PK_SM_STAGE (PK_SM_CODES + PK_SM_CODES_BYTES)
#define PK_STAGE_WORDS 64
#define PK_SM_STAGE_BYTES (ZB_WARPS_PER_CHUNK * PK_STAGE_WORDS * 4)
#define PK_SM_BAR (PK_SM_STAGE + PK_SM_STAGE_BYTES)
#define PK_SM_TOTAL (PK_SM_BAR + 16)

// OR `nbits` (<= 32) bits of v into the global bitstream at absolute bit position gb.
__device__ __forceinline__ void or_bits_global(uint32_t *dstw, uint64_t gb, uint32_t v, uint32_t nbits) {
  if (nbits == 0) return;
  if (nbits < 32) v &= (1u << nbits) - 1u;
  uint64_t word = gb >> 5;
  uint32_t sh = (uint32_t)(gb & 31u);
  atomicOr(&dstw[word], v << sh);
  if (sh && sh + nbits > 32) atomicOr(&dstw[word + 1], v >> (32u - sh));
}

__global__ void __launch_bounds__(LZ_THREADS, 3)
    k_pack(ZbCompressWork w) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *data = smem + PK_SM_DATA;
  uint32_t *codes = reinterpret_cast<uint32_t *>(smem + PK_SM_CODES);  // [0,288) litlen, [288,320) dist
  uint32_t *stage_all = reinterpret_cast<uint32_t *>(smem + PK_SM_STAGE);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + PK_SM_BAR);

  const uint32_t chunk = blockIdx.x;
  const ZbChunkDesc d = w.desc[chunk];
  const ZbCodebook *cb = &w.cb[chunk];
  const uint32_t len = d.len;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t btype = cb->block_type;
  const uint64_t out0 = w.chunk_off[chunk];  // byte offset of this chunk's deflate bytes
  uint32_t *dstw = reinterpret_cast<uint32_t *>(w.dst);

  if (tid == 0) {
    zb_mbar_init(bar, 1);
    zb_fence_mbar_init();
  }
  __syncthreads();
  const uint32_t mis = (uint32_t)((uintptr_t)(w.src + d.src_off) & 15u);
  if (tid == 0 && len) zb_stage_chunk(data, w.src + d.src_off, len, bar);
  for (int i = tid; i < 320; i += LZ_THREADS) codes[i] = i < 288 ? cb->ll[i] : cb->dd[i - 288];
  for (int i = tid; i < ZB_WARPS_PER_CHUNK * PK_STAGE_WORDS; i += LZ_THREADS) stage_all[i] = 0;

  // ---- framing bytes (zippy.nim:21-42, 50-58, 60-78) ----
  if (tid == 32 && (d.flags & ZB_CHUNK_FIRST)) {
    uint8_t *h = w.dst + w.member_off[d.member];
    if (w.data_format == ZB_DF_GZIP) {
      h[0] = 31; h[1] = 139; h[2] = 8; h[3] = 8;  // FNAME flag set, as the reference does
      uint32_t k = w.fname_len ? w.fname_len[d.member] : 0u;
      for (uint32_t i = 0; i < k; i++) h[10 + i] = (uint8_t)(97 + i);
    } else if (w.data_format == ZB_DF_ZLIB) {
      h[0] = 0x78; h[1] = 0x01;
    }
  }
  if (tid == 64 && (d.flags & ZB_CHUNK_LAST)) {
    uint8_t *t = w.dst + out0 + cb->total_bytes;
    uint32_t ck = w.member_check[d.member];
    if (w.data_format == ZB_DF_GZIP) {
      uint32_t isz = w.member_isize[d.member];
      t[0] = (uint8_t)ck; t[1] = (uint8_t)(ck >> 8); t[2] = (uint8_t)(ck >> 16); t[3] = (uint8_t)(ck >> 24);
      t[4] = (uint8_t)isz; t[5] = (uint8_t)(isz >> 8); t[6] = (uint8_t)(isz >> 16); t[7] = (uint8_t)(isz >> 24);
    } else if (w.data_format == ZB_DF_ZLIB) {
      t[0] = (uint8_t)(ck >> 24); t[1] = (uint8_t)(ck >> 16); t[2] = (uint8_t)(ck >> 8); t[3] = (uint8_t)ck;
    }
  }
  __syncthreads();
  if (len) zb_mbar_wait(bar, 0);

  if (btype == 0) {
    // stored blocks (deflate.nim:179-205): 1 header byte, LEN, NLEN, bytes
    uint8_t *o = w.dst + out0;
    uint32_t npieces = len == 0 ? 1u : (len + 65534u) / 65535u;
    for (uint32_t pc = 0; pc < npieces; pc++) {
      uint32_t s0 = pc * 65535u, n = min(65535u, len - s0);
      uint8_t *ob = o + (size_t)pc * 5u + s0;
      if (tid == 0) {
        ob[0] = (uint8_t)((cb->is_final && pc == npieces - 1) ? 1 : 0);
        ob[1] = (uint8_t)n; ob[2] = (uint8_t)(n >> 8);
        ob[3] = (uint8_t)~n; ob[4] = (uint8_t)((~n) >> 8);
      }
      for (uint32_t i = (uint32_t)tid; i < n; i += LZ_THREADS) ob[5 + i] = data[mis + s0 + i];
    }
    return;
  }

  const uint64_t gbit0 = out0 * 8ull;
  // ---- block header + dynamic tables ----
  if (warp == 0) {
    uint32_t hb = cb->hdr_bits;
    for (uint32_t k = (uint32_t)lane; k * 32u < hb; k += 32) {
      uint32_t piece = 0;
      for (int j = 0; j < 4; j++) piece |= (uint32_t)cb->hdr[k * 4 + j] << (8 * j);
      uint32_t nb = min(32u, hb - k * 32u);
      or_bits_global(dstw, gbit0 + k * 32ull, piece, nb);
    }
  }
  // ---- end of block (+ byte-aligning empty stored block when more chunks follow) ----
  if (tid == 96) {
    uint32_t e = codes[256];
    uint64_t eb = gbit0 + cb->eob_bit_start;
    or_bits_global(dstw, eb, e & 0xffffu, e >> 16);
    if (!cb->is_final) {
      uint64_t after = cb->eob_bit_start + (e >> 16) + 3u;
      uint64_t byte_al = (after + 7u) >> 3;
      uint8_t *o = w.dst + out0 + byte_al;
      o[2] = 0xff;
      o[3] = 0xff;
    }
  }
  // ---- tokens of this warp's sub-chunk ----
  const uint32_t b0 = (uint32_t)warp * ZB_SUB_BYTES;
  if (b0 >= len) return;
  const uint32_t b1 = min(b0 + ZB_SUB_BYTES, len);
  uint32_t *stage = stage_all + warp * PK_STAGE_WORDS;
  uint64_t bitpos = gbit0 + cb->warp_bit_start[warp];
  const size_t win_base = (size_t)chunk * ZB_WINDOWS_PER_CHUNK;
  for (uint32_t wb = b0; wb < b1; wb += 32) {
    const uint32_t win = wb >> 5;
    const uint2 mk = w.masks[win_base + win];
    if (mk.x == 0) continue;
    const bool is_sel = (mk.x >> lane) & 1u, is_m = (mk.y >> lane) & 1u;
    uint64_t bits = 0;
    uint32_t nb = 0;
    if (is_m) {
      uint32_t rank = (uint32_t)__popc(mk.y & ((1u << lane) - 1u));
      uint32_t rec = w.recs[(win_base + win) * ZB_MATCH_SLOTS + rank];
      uint32_t mlen = (rec & 511u) + 3u, dist = (rec >> 9) + 1u;
      int lc = zb_len_code(mlen), dc = zb_dist_code(dist);
      uint32_t e1 = codes[257 + lc], e2 = codes[288 + dc];
      bits = e1 & 0xffffu;
      nb = e1 >> 16;
      bits |= (uint64_t)(mlen - zb_len_base(lc)) << nb;
      nb += (uint32_t)zb_len_extra_bits(lc);
      bits |= (uint64_t)(e2 & 0xffffu) << nb;
      nb += e2 >> 16;
      bits |= (uint64_t)(dist - zb_dist_base(dc)) << nb;
      nb += (uint32_t)zb_dist_extra_bits(dc);
    } else if (is_sel) {
      uint32_t e = codes[data[mis + wb + (uint32_t)lane]];
      bits = e & 0xffffu;
      nb = e >> 16;
    }
    uint32_t incl = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(ZB_FULL, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(ZB_FULL, incl, 31);
    const uint32_t lead = (uint32_t)(bitpos & 31u);
    if (nb) {
      uint32_t lb = lead + incl - nb;  // local bit offset in the staging words
      uint32_t wd = lb >> 5, sh = lb & 31u;
      uint64_t lo = bits << sh;
      atomicOr(&stage[wd], (uint32_t)lo);
      if (sh + nb > 32) atomicOr(&stage[wd + 1], (uint32_t)(lo >> 32));
      if (sh + nb > 64) atomicOr(&stage[wd + 2], (uint32_t)(bits >> (64u - sh)));
    }
    __syncwarp();
    const uint32_t nwords = (lead + total + 31u) >> 5;
    const uint64_t word0 = bitpos >> 5;
    for (uint32_t j = (uint32_t)lane; j < nwords; j += 32) {
      uint32_t sv = stage[j];
      if (sv) atomicOr(&dstw[word0 + j], sv);
      stage[j] = 0;
    }
    __syncwarp();
    bitpos += total;
  }
}

// ------------------------------------------------------------------------------------
cudaError_t zb_launch_lz(const ZbCompressWork &w, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_lz<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, LZ_SM_TOTAL);
    cudaFuncSetAttribute(k_lz<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LZ_SM_TOTAL);
    cudaFuncSetAttribute(k_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SM_TOTAL);
    attr_set = true;
  }
  if (w.n_chunks == 0) return cudaSuccess;
  if (w.level == -2 || w.level == 0)
    k_lz<0><<<w.n_chunks, LZ_THREADS, LZ_SM_TOTAL, s>>>(w.src, w.desc, w.masks, w.recs, w.hist, w.chk, w.tabs);
  else
    k_lz<1><<<w.n_chunks, LZ_THREADS, LZ_SM_TOTAL, s>>>(w.src, w.desc, w.masks, w.recs, w.hist, w.chk, w.tabs);
  return cudaGetLastError();
}
cudaError_t zb_launch_huff(const ZbCompressWork &w, cudaStream_t s) {
  if (w.n_chunks == 0) return cudaSuccess;
  k_huff<<<(w.n_chunks + 63) / 64, 64, 0, s>>>(w.desc, w.hist, w.cb, w.n_chunks, w.level);
  return cudaGetLastError();
}
cudaError_t zb_launch_scan(const ZbCompressWork &w, cudaStream_t s) {
  k_scan<<<1, SCAN_THREADS, 0, s>>>(w);
  return cudaGetLastError();
}
cudaError_t zb_launch_pack(const ZbCompressWork &w, cudaStream_t s) {
  if (w.n_chunks == 0) return cudaSuccess;
  k_pack<<<w.n_chunks, LZ_THREADS, PK_SM_TOTAL, s>>>(w);
  return cudaGetLastError();
}
