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
  // op + tlen is computed in 32 bits: keep 512 bytes of headroom below 2^32
  const uint32_t cap = COUNT_ONLY ? 0xfffffdffu : (uint32_t)min(cap64, (uint64_t)0xfffffdffu);
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
    // One exit: errors are recorded in `err` and resolved after the loop, and a reader that ran a
    // whole line past the end (b.overrun, set on the rare line-advance path) simply makes the
    // capacity check fail -- so the hot path carries no per-token error plumbing.
    uint32_t tok = 0, ntok = 0, batch_op = op;
    int err = 0;  // 0: end of block, 1: invalid stream, 2: out of room (or overrun)
    for (;;) {
      br_refill(b);
      const uint32_t e = ws->lut_ll[(uint32_t)b.buf & ((1u << LL_BITS) - 1u)];
      uint32_t sym = e & 511u;
      const uint32_t l = e >> 9;
      if (l == 0) sym = decode_slow(b, tl, ws->syms_ll);  // long code (consumes its own bits) or none
      b.buf >>= l;
      b.cnt -= (int)l;
      uint32_t t = sym, tlen = 1;
      if (sym >= 256) {
        if (sym == 256) break;
        const uint32_t lidx = sym - 257u;
        if (lidx >= 29u) {  // includes the undecodable-code case
          err = 1;
          break;
        }
        const uint32_t lt = len_tab[lidx];
        tlen = (lt & 0xffffu) + br_take(b, (int)(lt >> 16));
        br_refill(b);
        const uint32_t didx = decode_sym<D_BITS>(b, ws->lut_d, td, ws->syms_d);
        if (didx >= 30u) {
          err = 1;
          break;
        }
        const uint32_t dt = dist_tab[didx];
        const uint32_t dist = (dt & 0xffffu) + br_take(b, (int)(dt >> 16));
        if (dist > op) {
          err = 1;
          break;
        }
        t = (1u << 31) | ((dist - 1u) << 9) | tlen;
      }
      if (op + tlen > (b.overrun ? 0u : cap)) {
        err = 2;
        break;
      }
      if ((uint32_t)lane == ntok) tok = t;
      ntok++;
      op += tlen;
      if (ntok == INF_G) {
        if (!COUNT_ONLY) flush_tokens(out, batch_op, tok, INF_G);
        ntok = 0;
        batch_op = op;
      }
    }
    if (err) {
      if (b.overrun || br_past_end(b)) return ZB_ERR_END_OF_BUFFER;  // decoding ran off the input
      return err == 1 ? ZB_ERR_UNCOMPRESS : ZB_ERR_DST_TOO_SMALL;
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
// Blocked combine-reduce checksums (replaces crc32 crc.nim:53 / crc32_simd.nim:39-144 and
// adler32 adler32.nim:6 / adler32_simd.nim:45-120, and the trailer checks gzip.nim:80-88,
// zippy.nim:154-162).  A buffer is cut into 64 KiB pieces; k_piece_checksum gives every piece
// to one CTA (TMA-staged, 8 warps x 8 KiB, lane-strided CRC stepping + dp4a Adler sums, see
// zb_device.cuh), k_buffer_combine folds a buffer's pieces with x^(8*len) multiplications /
// the closed-form Adler merge, and in verify mode compares with the trailer.
#define CK_PIECE ZB_CHUNK_BYTES
#define CK_THREADS (ZB_WARPS_PER_CHUNK * 32)
#define CK_STAGES 3
#define CK_STAGE_BYTES (CK_PIECE + 128)
#define CK_SM_CRC (CK_STAGES * CK_STAGE_BYTES)
#define CK_SM_LMUL (CK_SM_CRC + 4096)
#define CK_SM_PART (CK_SM_LMUL + 48 * 4)
#define CK_SM_BAR (CK_SM_PART + ZB_WARPS_PER_CHUNK * 24)
#define CK_SM_TOTAL (CK_SM_BAR + 8 * CK_STAGES + 8)

__device__ __forceinline__ uint32_t piece_len(uint64_t buflen, uint64_t rel) {
  return rel < buflen ? (uint32_t)min((uint64_t)CK_PIECE, buflen - rel) : 0u;
}
__device__ __forceinline__ uint32_t ck_piece_info(const ZbChecksumWork &w, uint32_t pid, const uint8_t *&src) {
  const ZbPiece pc = w.pieces[pid];
  uint64_t buflen = w.lens ? w.lens[pc.buf] : w.off[pc.buf + 1] - w.off[pc.buf];
  if (w.status && w.status[pc.buf] != ZB_OK) buflen = 0;
  src = w.src + w.off[pc.buf] + pc.rel;
  return piece_len(buflen, pc.rel);
}

// Persistent CTAs (one per SM), a 3-stage ring of 64 KiB shared-memory buffers filled by TMA
// bulk copies: while the 8 warps checksum stage k, the copies for k+1 and k+2 are in flight.
// Piece descriptors (source pointer, length: two dependent global loads) are fetched 16..32
// pieces ahead into a small shared ring so they never sit on the critical path.
// This is the one kernel on the path that is genuinely memory-bound (0.12 instructions per byte).
#define CK_INFO 32
__global__ void __launch_bounds__(CK_THREADS, 1)
    k_piece_checksum(ZbChecksumWork w) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t *crc_tab = reinterpret_cast<uint32_t *>(smem + CK_SM_CRC);
  uint32_t *lane_mul = reinterpret_cast<uint32_t *>(smem + CK_SM_LMUL);
  uint64_t *part = reinterpret_cast<uint64_t *>(smem + CK_SM_PART);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + CK_SM_BAR);
  __shared__ const uint8_t *info_src[CK_INFO];
  __shared__ uint32_t info_len[CK_INFO];
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t stride = gridDim.x;
  if (tid == 0) {
    for (int i = 0; i < CK_STAGES; i++) zb_mbar_init(&bars[i], 1);
    zb_fence_mbar_init();
  }
  for (int i = tid; i < 1024; i += CK_THREADS) crc_tab[i] = (&w.tabs->mul1024[0][0])[i];
  if (tid < 33) lane_mul[tid] = w.tabs->lane_mul[tid];
  if (tid >= 64 && tid < 72) lane_mul[33 + tid - 64] = w.tabs->sub_mul[tid - 64];
  if (tid >= 96 && tid < 100) lane_mul[41 + tid - 96] = w.tabs->quart_mul[tid - 96];
  if (tid >= 128 && tid < 128 + CK_INFO) {  // descriptors of this CTA's first 32 pieces
    const uint32_t j = (uint32_t)tid - 128u, pid = blockIdx.x + j * stride;
    const uint8_t *src = nullptr;
    info_len[j] = pid < w.n_pieces ? ck_piece_info(w, pid, src) : 0u;
    info_src[j] = src;
  }
  __syncthreads();
  if (tid == 0) {  // prologue: fill the ring
    for (uint32_t k = 0; k < CK_STAGES; k++) {
      if (blockIdx.x + k * stride >= w.n_pieces) break;
      if (info_len[k]) zb_stage_chunk(smem + k * CK_STAGE_BYTES, info_src[k], info_len[k], &bars[k]);
    }
  }
  uint32_t k = 0, phases = 0;  // bit s of `phases` = parity of the next completion of stage s
  for (uint32_t pid = blockIdx.x; pid < w.n_pieces; pid += stride, k++) {
    const uint32_t stage = k % CK_STAGES;
    const uint32_t len = info_len[k % CK_INFO];
    const uint8_t *src = info_src[k % CK_INFO];
    const uint8_t *data = smem + stage * CK_STAGE_BYTES;
    const uint32_t mis = (uint32_t)((uintptr_t)src & 15u);
    if (len) {  // empty pieces are never copied, so their stage's phase does not advance
      zb_mbar_wait(&bars[stage], (phases >> stage) & 1u);
      phases ^= 1u << stage;
    }
    const uint32_t b0 = (uint32_t)warp * ZB_SUB_BYTES, b1 = min(b0 + ZB_SUB_BYTES, len);
    ZbCheck c;
    c.crc_raw = 0;
    c.a_sum = c.b_sum = 0;
    if (b0 < len) {
      c = zb_warp_checksums(data, mis + b0, b1 - b0, crc_tab, lane_mul);
      const uint32_t after = len - b1;
      if (after) {
        const uint32_t shift = ((after & (ZB_SUB_BYTES - 1)) == 0) ? lane_mul[33 + after / ZB_SUB_BYTES] : zb_xpow8(after);
        c.crc_raw = zb_gf2_mul(c.crc_raw, shift);
        c.b_sum += (uint64_t)after * c.a_sum;
      }
    }
    if (lane == 0) {
      part[warp * 3 + 0] = c.crc_raw;
      part[warp * 3 + 1] = c.a_sum;
      part[warp * 3 + 2] = c.b_sum;
    }
    __syncthreads();  // every warp is done with this stage (and with info slot k)
    if (tid == 0) {
      // refill this stage with the piece three iterations ahead
      const uint32_t nk = k + CK_STAGES;
      if (pid + CK_STAGES * stride < w.n_pieces && info_len[nk % CK_INFO])
        zb_stage_chunk(smem + stage * CK_STAGE_BYTES, info_src[nk % CK_INFO], info_len[nk % CK_INFO], &bars[stage]);
      uint32_t raw = 0;
      uint64_t a = 0, b = 0;
      for (int j = 0; j < ZB_WARPS_PER_CHUNK; j++) {
        raw ^= (uint32_t)part[j * 3 + 0];
        a += part[j * 3 + 1];
        b += part[j * 3 + 2];
      }
      ZbChunkCheck cc;
      cc.crc_raw = raw;
      cc.adler = ((uint32_t)(b % ZB_ADLER_MOD) << 16) | (uint32_t)(a % ZB_ADLER_MOD);  // (B mod p, A mod p), not yet an Adler value
      w.piece_out[pid] = cc;
    } else if (warp == 1 && (k % 16u) == 15u && lane < 16) {
      // descriptors for pieces k+17 .. k+32 go into the half of the ring that has just been used up
      const uint32_t j = k + 17u + (uint32_t)lane, npid = blockIdx.x + j * stride;
      const uint8_t *nsrc = nullptr;
      info_len[j % CK_INFO] = npid < w.n_pieces ? ck_piece_info(w, npid, nsrc) : 0u;
      info_src[j % CK_INFO] = nsrc;
    }
    __syncthreads();  // part[] and the descriptor ring are consistent for the next piece
  }
}

// One warp per buffer: lane j folds pieces j, j+32, ... (Horner in x^(8 * 32 * 64 KiB)), the lane
// results are shifted to the end of the buffer and XOR-reduced; Adler sums add up directly.
__global__ void __launch_bounds__(128)
    k_buffer_combine(ZbChecksumWork w) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = (int)(threadIdx.x & 31u);
  if (i >= w.n) return;
  if (w.status && w.status[i] != ZB_OK) return;
  int kind = w.kind;
  if (w.kinds) {
    const uint32_t kd = w.kinds[i];
    if (kd == ZB_DF_GZIP) kind = 0;
    else if (kd == ZB_DF_ZLIB) kind = 1;
    else return;  // raw deflate: nothing to verify
  }
  const uint64_t buflen = w.lens ? w.lens[i] : w.off[i + 1] - w.off[i];
  const uint32_t p0 = w.first[i];
  const uint32_t np = (uint32_t)((buflen + CK_PIECE - 1) / CK_PIECE);  // pieces that hold data
  uint32_t r = 0;
  uint64_t a = 0, b = 0, end = 0;  // end = bytes from the buffer start to the end of this lane's last piece
  if (kind == 0) {
    uint32_t step = 0;  // x^(8 * 32 * 64 KiB), computed only if a lane has more than one piece
    for (uint32_t k = (uint32_t)lane; k < np; k += 32) {
      const uint32_t l = piece_len(buflen, (uint64_t)k * CK_PIECE);
      const uint32_t raw = w.piece_out[p0 + k].crc_raw;
      if (k >= 32) {
        if (l == CK_PIECE) {
          if (!step) step = zb_xpow8(32ull * CK_PIECE);
          r = zb_gf2_mul(r, step);
        } else {
          r = zb_gf2_mul(r, zb_xpow8(31ull * CK_PIECE + l));
        }
      }
      r ^= raw;
      end = (uint64_t)k * CK_PIECE + l;
    }
    if ((uint32_t)lane < np && end < buflen) r = zb_gf2_mul(r, zb_xpow8(buflen - end));
    r = zb_warp_xor(r);
  } else {
    for (uint32_t k = (uint32_t)lane; k < np; k += 32) {
      const uint32_t l = piece_len(buflen, (uint64_t)k * CK_PIECE);
      const uint32_t ab = w.piece_out[p0 + k].adler;
      const uint64_t ak = ab & 0xffffu, bk = ab >> 16;
      const uint64_t after = (buflen - ((uint64_t)k * CK_PIECE + l)) % ZB_ADLER_MOD;
      a += ak;
      b = (b + bk + after * ak) % ZB_ADLER_MOD;
    }
    a = zb_warp_sum64(a % ZB_ADLER_MOD);
    b = zb_warp_sum64(b);
  }
  if (lane != 0) return;
  const uint32_t v = kind == 0 ? ~(zb_gf2_mul(buflen == CK_PIECE ? w.tabs->sub_mul[0] : zb_xpow8(buflen), 0xffffffffu) ^ r)
                               : zb_adler_from_sums(a % ZB_ADLER_MOD, b % ZB_ADLER_MOD, buflen);
  if (w.out) w.out[i] = v;
  if (w.expect) {
    if (v != w.expect[i]) w.status[i] = ZB_ERR_CHECKSUM;
    else if (kind == 0 && w.isize_src) {
      const uint8_t *t = w.isize_src + w.isize_off[i + 1] - 4;
      if (ld_le32(t) != (uint32_t)buflen) w.status[i] = ZB_ERR_SIZE;
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

cudaError_t zb_launch_checksum(const ZbChecksumWork &w, cudaStream_t s) {
  if (w.n == 0) return cudaSuccess;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(k_piece_checksum, cudaFuncAttributeMaxDynamicSharedMemorySize, CK_SM_TOTAL);
    if (e != cudaSuccess) return e;
    done = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (w.n_pieces) {
    uint32_t grid = (uint32_t)sms < w.n_pieces ? (uint32_t)sms : w.n_pieces;
    k_piece_checksum<<<grid, CK_THREADS, CK_SM_TOTAL, s>>>(w);
  }
  k_buffer_combine<<<(w.n + 3) / 4, 128, 0, s>>>(w);
  return cudaGetLastError();
}
