// zb_inflate.cu -- batched inflate for sm_100a: one compressed member per warp,
// members pulled from a device work queue.
//
// Follows the behaviour (not the code) of the reference decoder:
//   format detect + wrappers : src/zippy.nim:100-165, src/zippy/gzip.nim:3-88
//   block loop               : src/zippy/inflate.nim:268-291
//   stored / fixed / dynamic : inflate.nim:252-266, 104-171
//   symbol loop + back-copy  : inflate.nim:173-250
//   canonical decode         : inflate.nim:24-102 (same accept/reject set: over-subscribed
//                              length sets are rejected, incomplete ones accepted, an
//                              undecodable code is an error)
// B200 formulation: the bit window lives in registers across the warp (two 128-byte
// lines, one word per lane, refilled by coalesced loads and read with shuffles); the
// canonical decode is lane-parallel (lane L tests the L-bit prefix; a ballot picks the
// code length), so the only per-warp shared memory is the sorted-symbol table; LZ
// back-copies are done by all 32 lanes reading the already-written output.
#include "zb_device.cuh"
#include "zb_kernels.h"

#define INF_G 32                      // lanes per member (16 = two members per warp: measured slower, the halves diverge)
#define INF_WARPS 8
#define INF_THREADS (INF_WARPS * 32)
#define INF_GROUPS (INF_THREADS / INF_G)

// A "group" is INF_G consecutive lanes that decode one member together; all collectives are
// restricted to the group's own lanes.  INF_G = 32 is one member per warp.  INF_G = 16 (two
// members per warp, hoping to halve the instructions per token) was measured 25-30 % SLOWER on
// B200: the two halves sit on different code paths almost all the time, so nothing is shared.
__device__ __forceinline__ int g_lane() { return (int)(threadIdx.x & (INF_G - 1)); }
__device__ __forceinline__ uint32_t g_shift() { return threadIdx.x & 31u & ~(uint32_t)(INF_G - 1); }
__device__ __forceinline__ uint32_t g_mask() { return (INF_G == 32 ? 0xffffffffu : ((1u << INF_G) - 1u)) << g_shift(); }
__device__ __forceinline__ uint32_t g_shfl(uint32_t v, int idx) { return __shfl_sync(g_mask(), v, idx, INF_G); }
__device__ __forceinline__ uint32_t g_shfl_up(uint32_t v, int d) { return __shfl_up_sync(g_mask(), v, d, INF_G); }
__device__ __forceinline__ uint32_t g_ballot(bool p) {
  return (__ballot_sync(g_mask(), p) >> g_shift()) & (INF_G == 32 ? 0xffffffffu : ((1u << INF_G) - 1u));
}
__device__ __forceinline__ uint32_t g_match_any(uint32_t v) { return __match_any_sync(g_mask(), v) >> g_shift(); }
__device__ __forceinline__ void g_sync() { __syncwarp(g_mask()); }
#define LL_BITS 10   // literal/length codes up to this long decode with one table lookup
#define D_BITS 9     // same for distance codes

struct WarpSmem {
  uint16_t lut_ll[1 << LL_BITS];  // symbol | code length << 9; 0 = longer code or no code: slow path
  uint16_t lut_d[1 << D_BITS];
  uint16_t syms_ll[288];          // symbols sorted by (length, symbol) for the canonical slow path
  uint16_t syms_d[32];
  uint8_t lens[320];              // code lengths of the current block (lit/len then distance)
  uint16_t cnt[16];               // per-length counts / running ranks while building
  uint16_t offs[16];
  uint16_t first[16];
};

struct Tree {  // lane L holds the entries for code length L
  uint32_t first, count, offs;
};

struct BitReader {
  const uint32_t *gbase;  // member start rounded down to 4 bytes
  uint32_t nwords;        // words that contain member bytes
  uint32_t cur, nxt;      // lane-held words of lines `line` and `line + 1` (a line = INF_G words)
  uint32_t widx;          // next word to feed into the bit buffer
  uint64_t buf;
  int cnt;                // valid bits in buf
  uint64_t end_bit;       // absolute bit (from gbase) one past the member's last byte
  bool overrun;           // set when the reader is found to have consumed bits past end_bit
};

__device__ __forceinline__ uint32_t br_load_line(const BitReader &b, uint32_t line) {
  uint32_t idx = line * (uint32_t)INF_G + (uint32_t)g_lane();
  return idx < b.nwords ? __ldg(b.gbase + idx) : 0u;
}
__device__ __forceinline__ uint64_t br_consumed_abs(const BitReader &b) {
  return (uint64_t)b.widx * 32ull - (uint64_t)b.cnt;
}
__device__ __forceinline__ bool br_past_end(const BitReader &b) { return br_consumed_abs(b) > b.end_bit; }
__device__ __forceinline__ uint32_t br_next_word(BitReader &b) {
  uint32_t w = g_shfl(b.cur, (int)(b.widx & (uint32_t)(INF_G - 1)));
  b.widx++;
  if ((b.widx & (uint32_t)(INF_G - 1)) == 0) {
    b.cur = b.nxt;
    b.nxt = br_load_line(b, b.widx / (uint32_t)INF_G + 1u);
    // bits past the end read as zero; a reader that is a whole line past the end can only be
    // decoding garbage: flag it here so that no decode loop runs away (checked by the callers)
    if ((uint64_t)b.widx * 32ull > b.end_bit + 64ull) b.overrun = true;
  }
  return w;
}
// position the reader at byte `byte_off` of the member (shift0 = member start & 3)
__device__ __forceinline__ void br_seek(BitReader &b, uint32_t shift0, uint64_t byte_off) {
  uint64_t abit = (shift0 + byte_off) * 8ull;
  b.widx = (uint32_t)(abit >> 5);
  uint32_t skip = (uint32_t)(abit & 31u);
  uint32_t line = b.widx / (uint32_t)INF_G;
  b.cur = br_load_line(b, line);
  b.nxt = br_load_line(b, line + 1);
  uint32_t w = br_next_word(b);
  b.buf = (uint64_t)(w >> skip);
  b.cnt = 32 - (int)skip;
}
__device__ __forceinline__ void br_refill(BitReader &b) {  // afterwards cnt >= 32
  if (b.cnt < 32) {
    b.buf |= (uint64_t)br_next_word(b) << b.cnt;
    b.cnt += 32;
  }
}
__device__ __forceinline__ uint32_t br_take(BitReader &b, int n) {  // n <= 32, n <= cnt
  uint32_t v = (uint32_t)b.buf & (n >= 32 ? 0xffffffffu : ((1u << n) - 1u));
  b.buf >>= n;
  b.cnt -= n;
  return v;
}

// Build the decode state from n code lengths (inflate.nim:24-65 initHuffman): the per-length
// canonical ranges (for the lane-parallel slow path), the sorted symbols, and a direct lookup
// table for codes of at most lut_bits.  Returns false for an over-subscribed set
// (inflate.nim:32-34, 45-46).
__device__ __forceinline__ bool build_tree(const uint8_t *lens, int n, uint16_t *syms, uint16_t *lut, int lut_bits,
                                           WarpSmem *ws, Tree &t) {
  const int lane = g_lane();
  if (lane < 16) ws->cnt[lane] = 0;
  {
    uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4 *l4 = reinterpret_cast<uint4 *>(lut);
    for (int i = lane; i < ((1 << lut_bits) * 2) / 16; i += INF_G) l4[i] = z;
  }
  g_sync();
  for (int base = 0; base < n; base += INF_G) {
    int s = base + lane;
    uint32_t l = s < n ? lens[s] : 0u;
    uint32_t grp = g_match_any(l);
    if (l && lane == __ffs((int)grp) - 1) ws->cnt[l] = (uint16_t)(ws->cnt[l] + __popc(grp));
    g_sync();
  }
  // per-length first code / first slot (all lanes compute the same recurrence)
  bool ok = true;
  uint32_t code = 0, k = 0, my_first = 0, my_count = 0, my_offs = 0;
  for (int i = 1; i < 16; i++) {
    uint32_t c = ws->cnt[i];
    if (i == lane) {
      my_first = code;
      my_count = c;
      my_offs = k;
    }
    code += c;
    if (c > 0 && code - 1 >= (1u << i)) ok = false;
    code <<= 1;
    k += c;
  }
  g_sync();
  if (lane < 16) {
    ws->offs[lane] = (uint16_t)my_offs;
    ws->first[lane] = (uint16_t)my_first;
    ws->cnt[lane] = 0;  // becomes the running rank per length
  }
  g_sync();
  if (!ok) return false;
  for (int base = 0; base < n; base += INF_G) {
    int s = base + lane;
    uint32_t l = s < n ? lens[s] : 0u;
    uint32_t grp = g_match_any(l);
    if (l) {
      uint32_t rank = ws->cnt[l] + (uint32_t)__popc(grp & ((1u << lane) - 1u));
      syms[(uint32_t)ws->offs[l] + rank] = (uint16_t)s;
      if ((int)l <= lut_bits) {
        uint32_t c = (uint32_t)ws->first[l] + rank;          // canonical code, MSB first
        uint32_t rev = __brev(c) >> (32 - l);                // as it appears in the LSB-first stream
        uint16_t e = (uint16_t)((uint32_t)s | (l << 9));
        for (uint32_t idx = rev; idx < (1u << lut_bits); idx += (1u << l)) lut[idx] = e;
      }
    }
    g_sync();
    if (l && lane == __ffs((int)grp) - 1) ws->cnt[l] = (uint16_t)(ws->cnt[l] + __popc(grp));
    g_sync();
  }
  t.first = my_first;
  t.count = (lane >= 1 && lane <= 15) ? my_count : 0u;
  t.offs = my_offs;
  return true;
}

// Lane-parallel canonical decode (codes longer than the lookup table, and the
// "no code matches" case, which returns 0xffff: inflate.nim:77-82).
__device__ __forceinline__ uint32_t decode_slow(BitReader &b, const Tree &t, const uint16_t *syms) {
  const int lane = g_lane();
  uint32_t rev = __brev((uint32_t)b.buf);
  uint32_t code = lane ? (rev >> (32 - lane)) : 0u;
  uint32_t rel = code - t.first;
  uint32_t hit = g_ballot(rel < t.count);
  if (!hit) return 0xffffu;
  int L = __ffs((int)hit) - 1;
  uint32_t idx = g_shfl(t.offs + rel, L);
  b.buf >>= L;
  b.cnt -= L;
  return syms[idx];
}
// One symbol (needs >= 15 buffered bits; bits past the end of the member read as zero).
template <int BITS>
__device__ __forceinline__ uint32_t decode_sym(BitReader &b, const uint16_t *lut, const Tree &t,
                                               const uint16_t *syms) {
  uint32_t e = lut[(uint32_t)b.buf & ((1u << BITS) - 1u)];
  uint32_t l = e >> 9;
  if (l) {
    b.buf >>= l;
    b.cnt -= (int)l;
    return e & 511u;
  }
  return decode_slow(b, t, syms);
}

__device__ __forceinline__ uint32_t ld_le32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// zippy.nim:100-165 + gzip.nim:3-66: resolve the format, validate the wrapper, find the
// payload start and the trailer checksum.  All lanes run this redundantly.
__device__ __forceinline__ int parse_wrapper(const uint8_t *src, uint64_t len, int fmt, uint64_t raw_pos,
                                             uint64_t &pos, uint32_t &kind, uint32_t &expect, uint32_t &isize) {
  expect = 0;
  isize = 0;
  if (fmt == ZB_DF_DETECT) {
    if (len > 18 && src[0] == 31 && src[1] == 139 && src[2] == 8 && (src[3] & 0xe0) == 0) fmt = ZB_DF_GZIP;
    else if (len > 6 && (src[0] & 0x0f) == 8 && (src[0] >> 4) <= 7 && (((uint32_t)src[0] * 256u) + src[1]) % 31u == 0)
      fmt = ZB_DF_ZLIB;
    else return ZB_ERR_DETECT;
  }
  kind = (uint32_t)fmt;
  if (fmt == ZB_DF_GZIP) {
    if (len < 18) return ZB_ERR_UNCOMPRESS;
    uint32_t flg = src[3];
    if (src[0] != 31 || src[1] != 139) return ZB_ERR_GZIP_ID;
    if (src[2] != 8) return ZB_ERR_METHOD;
    if (flg & 0xe0) return ZB_ERR_GZIP_RESERVED;
    if (flg & 4) return ZB_ERR_GZIP_FLAGS;
    uint64_t p = 10;
    for (int pass = 0; pass < 2; pass++) {
      if ((pass == 0 && (flg & 8)) || (pass == 1 && (flg & 16))) {
        while (p < len && src[p] != 0) p++;
        if (p >= len) return ZB_ERR_UNCOMPRESS;
        p++;
      }
    }
    if (flg & 2) {
      if (p + 2 >= len) return ZB_ERR_UNCOMPRESS;
      p += 2;
    }
    if (p + 8 >= len) return ZB_ERR_UNCOMPRESS;
    expect = ld_le32(src + len - 8);
    isize = ld_le32(src + len - 4);
    pos = p;
    return ZB_OK;
  }
  if (fmt == ZB_DF_ZLIB) {
    if (len < 6) return ZB_ERR_UNCOMPRESS;
    uint32_t cmf = src[0], flg = src[1];
    if ((cmf & 0x0f) != 8) return ZB_ERR_METHOD;
    if ((cmf >> 4) > 7) return ZB_ERR_CINFO;
    if ((cmf * 256u + flg) % 31u != 0) return ZB_ERR_HEADER;
    if (flg & 0x20) return ZB_ERR_FDICT;
    expect = ((uint32_t)src[len - 4] << 24) | ((uint32_t)src[len - 3] << 16) | ((uint32_t)src[len - 2] << 8) | src[len - 1];
    pos = 2;
    return ZB_OK;
  }
  if (fmt == ZB_DF_DEFLATE) {
    if (raw_pos > len) return ZB_ERR_END_OF_BUFFER;
    pos = raw_pos;
    return ZB_OK;
  }
  return ZB_ERR_INVALID_FORMAT;
}

// Materialise a batch of up to INF_G decoded tokens (lane i of the group holds token i):
//   literal: the byte;  match: 1 << 31 | (dist - 1) << 9 | len.
// One lane per token: a warp prefix sum of the lengths places every token; literals and
// matches whose source lies wholly before the batch are copied by their own lane, in
// parallel; the few matches that read bytes produced inside the batch follow in stream
// order, each copied by the whole warp (reads only touch finished output: i % dist).
__device__ __forceinline__ void flush_tokens(uint8_t *out, uint32_t batch_op, uint32_t tok, uint32_t ntok) {
  const int lane = g_lane();
  g_sync();  // stores of earlier batches are visible to every lane from here on
  const bool act = (uint32_t)lane < ntok;
  const bool is_m = act && (tok >> 31);
  const uint32_t len = act ? (is_m ? (tok & 511u) : 1u) : 0u;
  const uint32_t dist = ((tok >> 9) & 0x7fffu) + 1u;
  uint32_t incl = len;
#pragma unroll
  for (int o = 1; o < INF_G; o <<= 1) {
    uint32_t t = g_shfl_up(incl, o);
    if (lane >= o) incl += t;
  }
  const uint32_t rel = incl - len;           // output offset inside the batch
  uint8_t *to = out + batch_op + rel;
  const bool dep = is_m && dist < rel + len;  // source overlaps this batch's own output
  if (act && !is_m) {
    *to = (uint8_t)tok;
  } else if (is_m && !dep) {
    const uint8_t *from = to - dist;
    for (uint32_t k = 0; k < len; k++) to[k] = from[k];
  }
  uint32_t depmask = g_ballot(dep);
  while (depmask) {
    const int j = __ffs((int)depmask) - 1;
    depmask &= depmask - 1;
    const uint32_t rj = g_shfl(rel, j), lj = g_shfl(len, j);
    const uint32_t dj = g_shfl(dist, j);
    g_sync();
    uint8_t *tj = out + batch_op + rj;
    const uint8_t *fj = tj - dj;
    if (dj >= lj) {
      for (uint32_t i = (uint32_t)lane; i < lj; i += INF_G) tj[i] = fj[i];
    } else {
      for (uint32_t i = (uint32_t)lane; i < lj; i += INF_G) tj[i] = fj[i % dj];
    }
  }
}

template <bool COUNT_ONLY>
__device__ __forceinline__ int inflate_member(const uint8_t *src, uint64_t len, uint64_t pos, uint8_t *out,
                                              uint64_t cap64, WarpSmem *ws, const uint32_t *len_tab,
                                              const uint32_t *dist_tab, uint64_t &out_len) {
  const int lane = g_lane();
  const uint8_t clcl_order[19] = ZB_CLCL_ORDER;
  BitReader b;
  const uint32_t shift0 = (uint32_t)((uintptr_t)src & 3u);
  b.gbase = reinterpret_cast<const uint32_t *>(src - shift0);
  b.nwords = (uint32_t)((shift0 + len + 3u) >> 2);
  b.end_bit = (shift0 + len) * 8ull;
  b.overrun = false;
  br_seek(b, shift0, pos);
  // positions are 32-bit inside a member (a single member's output is limited to 4 GiB - 1)
  const uint32_t cap = COUNT_ONLY ? 0xffffffffu : (uint32_t)min(cap64, (uint64_t)0xffffffffu);
  uint32_t op = 0;
  Tree tl, td;
  bool final_block = false;
  while (!final_block) {
    br_refill(b);
    uint32_t bfinal = br_take(b, 1);
    uint32_t btype = br_take(b, 2);
    if (br_past_end(b)) return ZB_ERR_END_OF_BUFFER;
    if (bfinal) final_block = true;
    if (btype == 0) {
      // ---- stored (inflate.nim:252-266) ----
      br_take(b, b.cnt & 7);
      br_refill(b);
      uint32_t l = br_take(b, 16);
      br_refill(b);
      uint32_t nl = br_take(b, 16);
      if (br_past_end(b)) return ZB_ERR_END_OF_BUFFER;
      if (l + nl != 65535u) return ZB_ERR_UNCOMPRESS;
      if (l > 0) {
        uint64_t byte_pos = (br_consumed_abs(b) >> 3) - shift0;
        if (byte_pos + l > len) return ZB_ERR_END_OF_BUFFER;
        if (!COUNT_ONLY) {
          if (l > cap - op) return ZB_ERR_DST_TOO_SMALL;
          for (uint32_t i = (uint32_t)lane; i < l; i += INF_G) out[op + i] = src[byte_pos + i];
        } else if (l > cap - op) {
          return ZB_ERR_DST_TOO_SMALL;
        }
        op += l;
        br_seek(b, shift0, byte_pos + l);
      }
      continue;
    }
    if (btype == 3) return ZB_ERR_BLOCK_HEADER;
    int hlit, hdist;
    if (btype == 1) {
      // ---- fixed codes (inflate.nim:111-113) ----
      for (int i = lane; i < 320; i += INF_G) ws->lens[i] = (uint8_t)(i < 288 ? zb_fixed_ll_len(i) : 5);
      hlit = 288;
      hdist = 30;
      g_sync();
    } else {
      // ---- dynamic header (inflate.nim:115-171) ----
      br_refill(b);
      hlit = (int)br_take(b, 5) + 257;
      hdist = (int)br_take(b, 5) + 1;
      int hclen = (int)br_take(b, 4) + 4;
      if (hlit > ZB_NUM_LITLEN) return ZB_ERR_UNCOMPRESS;
      if (hdist > ZB_NUM_DIST) return ZB_ERR_UNCOMPRESS;
      for (int i = lane; i < 19; i += INF_G) ws->lens[i] = 0;
      g_sync();
      for (int i = 0; i < hclen; i++) {
        br_refill(b);
        uint32_t v = br_take(b, 3);
        if (lane == 0) ws->lens[clcl_order[i]] = (uint8_t)v;
      }
      g_sync();
      Tree tc;
      if (!build_tree(ws->lens, 19, ws->syms_d, ws->lut_d, 7, ws, tc)) return ZB_ERR_UNCOMPRESS;
      g_sync();
      // the code-length code now lives in syms_d / lut_d; lens[] is rewritten with the
      // unpacked literal/length + distance code lengths.
      int i = 0;
      const int total = hlit + hdist;
      uint32_t prev = 0;
      while (i != total) {
        br_refill(b);
        uint32_t sym = decode_sym<7>(b, ws->lut_d, tc, ws->syms_d);
        if (b.overrun) return ZB_ERR_END_OF_BUFFER;
        if (sym <= 15) {
          if (lane == 0) ws->lens[i] = (uint8_t)sym;
          prev = sym;
          i++;
        } else if (sym == 16) {
          if (i == 0) return ZB_ERR_UNCOMPRESS;
          int rep = (int)br_take(b, 2) + 3;
          if (i + rep > 320) return ZB_ERR_UNCOMPRESS;
          if (lane < rep) ws->lens[i + lane] = (uint8_t)prev;
          i += rep;
        } else if (sym == 17) {
          int rep = (int)br_take(b, 3) + 3;
          if (i + rep <= 320 && lane < rep) ws->lens[i + lane] = 0;
          i += rep;
          prev = 0;
        } else if (sym == 18) {
          int rep = (int)br_take(b, 7) + 11;
          for (int j = lane; j < rep && i + j < 320; j += INF_G) ws->lens[i + j] = 0;
          i += rep;
          prev = 0;
        } else {
          if (br_past_end(b)) return ZB_ERR_END_OF_BUFFER;
          return ZB_ERR_INVALID_SYMBOL;  // also the undecodable-code case (0xffff)
        }
        if (i > total) return ZB_ERR_UNCOMPRESS;
      }
      if (br_past_end(b)) return ZB_ERR_END_OF_BUFFER;
      g_sync();
    }
    if (!build_tree(ws->lens, hlit, ws->syms_ll, ws->lut_ll, LL_BITS, ws, tl)) return ZB_ERR_UNCOMPRESS;
    if (!build_tree(ws->lens + hlit, hdist, ws->syms_d, ws->lut_d, D_BITS, ws, td)) return ZB_ERR_UNCOMPRESS;
    g_sync();

    // ---- symbol loop (inflate.nim:173-250): decode into a 32-token batch, then flush ----
    uint32_t tok = 0, ntok = 0, batch_op = op;
    for (;;) {
      br_refill(b);
      if (b.overrun) return ZB_ERR_END_OF_BUFFER;  // a whole line past the end: stop decoding zeros
      uint32_t sym = decode_sym<LL_BITS>(b, ws->lut_ll, tl, ws->syms_ll);
      uint32_t t, tlen;
      if (sym < 256) {
        t = sym;
        tlen = 1;
      } else {
        if (sym == 256) break;
        uint32_t lidx = sym - 257u;
        if (lidx >= 29u) {  // includes the undecodable-code case
          if (br_past_end(b)) return ZB_ERR_END_OF_BUFFER;
          return ZB_ERR_UNCOMPRESS;
        }
        const uint32_t lt = len_tab[lidx];
        tlen = (lt & 0xffffu) + br_take(b, (int)(lt >> 16));
        br_refill(b);
        uint32_t didx = decode_sym<D_BITS>(b, ws->lut_d, td, ws->syms_d);
        if (didx >= 30u) {
          if (br_past_end(b)) return ZB_ERR_END_OF_BUFFER;
          return ZB_ERR_UNCOMPRESS;
        }
        const uint32_t dt = dist_tab[didx];
        const uint32_t dist = (dt & 0xffffu) + br_take(b, (int)(dt >> 16));
        if (dist > op) {
          if (br_past_end(b)) return ZB_ERR_END_OF_BUFFER;
          return ZB_ERR_UNCOMPRESS;
        }
        t = (1u << 31) | ((dist - 1u) << 9) | tlen;
      }
      if (tlen > cap - op) return (b.overrun || br_past_end(b)) ? ZB_ERR_END_OF_BUFFER : ZB_ERR_DST_TOO_SMALL;
      if ((uint32_t)lane == ntok) tok = t;
      ntok++;
      op += tlen;
      if (ntok == INF_G) {
        if (!COUNT_ONLY) flush_tokens(out, batch_op, tok, INF_G);
        ntok = 0;
        batch_op = op;
      }
    }
    if (!COUNT_ONLY && ntok) flush_tokens(out, batch_op, tok, ntok);
    if (br_past_end(b)) return ZB_ERR_END_OF_BUFFER;
  }
  out_len = op;
  return ZB_OK;
}

template <bool COUNT_ONLY>
__global__ void __launch_bounds__(INF_THREADS)
    k_inflate(ZbInflateWork w) {
  __shared__ __align__(16) WarpSmem wsm[INF_GROUPS];
  __shared__ uint32_t len_tab[32], dist_tab[32];  // base | extra bits << 16 (RFC 1951 3.2.5)
  const int lane = g_lane();
  WarpSmem *ws = &wsm[threadIdx.x / INF_G];
  if (threadIdx.x < 29) len_tab[threadIdx.x] = zb_len_base((int)threadIdx.x) | ((uint32_t)zb_len_extra_bits((int)threadIdx.x) << 16);
  if (threadIdx.x >= 32 && threadIdx.x < 62) {
    int c = (int)threadIdx.x - 32;
    dist_tab[c] = zb_dist_base(c) | ((uint32_t)zb_dist_extra_bits(c) << 16);
  }
  __syncthreads();
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(w.counter, 1u);
    i = g_shfl(i, 0);
    if (i >= w.n) break;
    const uint64_t s0 = w.src_off[i], s1 = w.src_off[i + 1];
    const uint8_t *src = w.src + s0;
    const uint64_t len = s1 - s0;
    uint64_t pos = 0, out_len = 0;
    uint32_t kind = 0, expect = 0, isize = 0;
    int st = parse_wrapper(src, len, w.data_format, w.pos, pos, kind, expect, isize);
    if (st == ZB_OK) {
      uint8_t *out = COUNT_ONLY ? nullptr : w.dst + w.dst_off[i];
      uint64_t cap = COUNT_ONLY ? 0 : w.dst_off[i + 1] - w.dst_off[i];
      if (COUNT_ONLY && kind == ZB_DF_GZIP) out_len = isize;  // gzip.nim:66 (trustSize's source)
      else st = inflate_member<COUNT_ONLY>(src, len, pos, out, cap, ws, len_tab, dist_tab, out_len);
    }
    g_sync();
    if (lane == 0) {
      w.status[i] = st;
      w.out_len[i] = st == ZB_OK ? out_len : 0;
      w.kind[i] = kind;
      w.expect[i] = expect;
    }
  }
}

// ------------------------------------------------------------------------------------
// Checksum kernels: one warp per buffer piece, staged through shared memory by TMA.
// k_verify checks inflate outputs against their trailers (gzip.nim:80-88, zippy.nim:154-162).
#define CK_PIECE 32768u
#define CK_WARPS 4
#define CK_THREADS (CK_WARPS * 32)
#define CK_SM_PIECE_BYTES (CK_PIECE + 64)
#define CK_SM_TOTAL (CK_WARPS * CK_SM_PIECE_BYTES + 4096 + 160 + CK_WARPS * 8 + 64)

// checksum of buf[0, n) by one warp; piece by piece through this warp's smem slot
__device__ __forceinline__ uint32_t warp_buffer_checksum(const uint8_t *buf, uint64_t n, int kind, uint8_t *slot,
                                                         uint64_t *bar, uint32_t &phase, const uint32_t *crc_tab,
                                                         const uint32_t *lane_mul) {
  const int lane = zb_lane();
  uint32_t raw = 0, ad = 1;
  for (uint64_t o = 0; o < n; o += CK_PIECE) {
    uint32_t pl = (uint32_t)min((uint64_t)CK_PIECE, n - o);
    uint32_t mis = (uint32_t)((uintptr_t)(buf + o) & 15u);
    if (lane == 0) zb_stage_chunk(slot, buf + o, pl, bar);
    zb_mbar_wait(bar, phase & 1u);
    phase++;
    ZbCheck c = zb_warp_checksums(slot, mis, pl, crc_tab, lane_mul);
    __syncwarp();
    if (kind == 0) raw = o ? (zb_gf2_mul(raw, zb_xpow8(pl)) ^ c.crc_raw) : c.crc_raw;
    else ad = zb_adler32_combine(ad, zb_adler_from_sums(c.a_sum % ZB_ADLER_MOD, c.b_sum % ZB_ADLER_MOD, pl), pl);
  }
  return kind == 0 ? zb_crc32_finalize(raw, n) : ad;
}

__global__ void __launch_bounds__(CK_THREADS)
    k_checksum(const uint8_t *base, const uint64_t *off, const uint64_t *lens_or_null, uint32_t *out, int *status,
               const uint32_t *expect, const uint32_t *kinds, const ZbCrcTables *tabs, uint32_t n, int fixed_kind,
               const uint8_t *src_for_isize, const uint64_t *src_off) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t *crc_tab = reinterpret_cast<uint32_t *>(smem + CK_WARPS * CK_SM_PIECE_BYTES);
  uint32_t *lane_mul = crc_tab + 1024;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + CK_WARPS * CK_SM_PIECE_BYTES + 4096 + 160);
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 1024; i += CK_THREADS) crc_tab[i] = (&tabs->mul1024[0][0])[i];
  if (tid < 33) lane_mul[tid] = tabs->lane_mul[tid];
  if (lane == 0) {
    zb_mbar_init(&bars[warp], 1);
    zb_fence_mbar_init();
  }
  __syncthreads();
  uint8_t *slot = smem + warp * CK_SM_PIECE_BYTES;
  uint32_t phase = 0;
  for (uint32_t i = blockIdx.x * CK_WARPS + (uint32_t)warp; i < n; i += gridDim.x * CK_WARPS) {
    if (status && status[i] != ZB_OK) continue;
    int kind = fixed_kind;
    if (kinds) {
      uint32_t k = kinds[i];
      if (k == ZB_DF_GZIP) kind = 0;
      else if (k == ZB_DF_ZLIB) kind = 1;
      else continue;  // raw deflate: nothing to verify
    }
    uint64_t len = lens_or_null ? lens_or_null[i] : off[i + 1] - off[i];
    uint32_t v = warp_buffer_checksum(base + off[i], len, kind, slot, &bars[warp], phase, crc_tab, lane_mul);
    if (lane == 0) {
      if (out) out[i] = v;
      if (expect) {
        if (v != expect[i]) status[i] = ZB_ERR_CHECKSUM;
        else if (kind == 0 && src_for_isize) {
          const uint8_t *t = src_for_isize + src_off[i + 1] - 4;
          if (ld_le32(t) != (uint32_t)len) status[i] = ZB_ERR_SIZE;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
cudaError_t zb_launch_inflate(const ZbInflateWork &w, cudaStream_t s) {
  if (w.n == 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint32_t blocks = (uint32_t)sms * 6u;
  uint32_t need = (w.n + INF_GROUPS - 1) / INF_GROUPS;
  if (blocks > need) blocks = need;
  cudaError_t e = cudaMemsetAsync(w.counter, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  if (w.count_only) k_inflate<true><<<blocks, INF_THREADS, 0, s>>>(w);
  else k_inflate<false><<<blocks, INF_THREADS, 0, s>>>(w);
  return cudaGetLastError();
}

static cudaError_t ck_attr() {
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(k_checksum, cudaFuncAttributeMaxDynamicSharedMemorySize, CK_SM_TOTAL);
    if (e != cudaSuccess) return e;
    done = true;
  }
  return cudaSuccess;
}

cudaError_t zb_launch_verify(const ZbInflateWork &w, cudaStream_t s) {
  if (w.n == 0 || w.count_only) return cudaSuccess;
  cudaError_t e = ck_attr();
  if (e != cudaSuccess) return e;
  uint32_t blocks = (w.n + CK_WARPS - 1) / CK_WARPS;
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  k_checksum<<<blocks, CK_THREADS, CK_SM_TOTAL, s>>>(w.dst, w.dst_off, w.out_len, nullptr, w.status, w.expect, w.kind,
                                                     w.tabs, w.n, 0, w.src, w.src_off);
  return cudaGetLastError();
}

cudaError_t zb_launch_checksum(const ZbChecksumWork &w, cudaStream_t s) {
  if (w.n == 0) return cudaSuccess;
  cudaError_t e = ck_attr();
  if (e != cudaSuccess) return e;
  uint32_t blocks = (w.n + CK_WARPS - 1) / CK_WARPS;
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  k_checksum<<<blocks, CK_THREADS, CK_SM_TOTAL, s>>>(w.src, w.off, nullptr, w.out, nullptr, nullptr, nullptr, w.tabs,
                                                     w.n, w.kind, nullptr, nullptr);
  return cudaGetLastError();
}
