// lzsim.c -- CPU model of the GPU matchers' PARSING DECISIONS (not their speed), used to choose
// dictionary geometry before spending GPU time: it replays what k_lz / k_lz2 would select on a
// file (64 KiB chunks, warp = sub-chunk, 32 positions per step, per-lane candidate evaluation,
// one-step lazy drop, greedy chain selection) and prices the tokens with per-chunk dynamic
// Huffman codes.  Development tool only; nothing in the product or the tests depends on it.
//
//   lzsim FILE [key=value ...]
//     sub=8192      sub-chunk bytes (warp's share of a 64 KiB chunk)
//     seg=8192      static segment bytes (0: no static segment tables)
//     sbits=10      log2 entries of each static table
//     slong=0       bytes hashed by the second static table per segment (0: none; e.g. 6..8)
//     sways=1       ways per static bucket (most recent first)
//     obits=10      log2 buckets of the warp's own incremental table
//     oways=2       ways per own bucket
//     olong=0       second own table hashed on this many bytes (0: none)
//     preseed=0     bytes before the sub-chunk inserted into the own table first
//     hist=32768    bytes of the member staged before the chunk (0 for level 1)
//     lazy=16       drop a match shorter than this when the next position has a longer one (0: off)
//     min4far=0     reject 4-byte matches farther than this (0: accept all)
//     near=1        also try the nearest same-hash position inside the window
//     good=8, budget candidates after a match >= good: stop early
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHUNK 65536
#define CAP 32
#define MAXM 258

static int P_sub = 8192, P_seg = 8192, P_sbits = 10, P_slong = 0, P_sways = 1, P_obits = 10, P_oways = 2, P_olong = 0;
static int P_slotmode = 0, P_slotcap = 0, P_order = 0;
static int P_ownstatic = 0, P_visit = 1000, P_back = 4;
static int P_preseed = 0, P_hist = 32768, P_lazy = 16, P_min4far = 0, P_near = 1, P_good = 8, P_maxcand = 64, P_minlen = 4;

static const uint8_t *D;   // whole file
static size_t N;

static inline uint32_t ld32(size_t p) { uint32_t v = 0; memcpy(&v, D + p, p + 4 <= N ? 4 : N - p); return v; }
static inline uint64_t ld64(size_t p) { uint64_t v = 0; memcpy(&v, D + p, p + 8 <= N ? 8 : N - p); return v; }
static inline uint32_t h4(uint32_t v, int bits) { return (v * 0x9E3779B1u) >> (32 - bits); }
static inline uint32_t hl(uint64_t v, int nbytes, int bits) {
  v <<= (8 - nbytes) * 8;
  return (uint32_t)((v * 0x9E3779B97F4A7C15ull) >> (64 - bits));
}

// ---- huffman cost ----
static double huff_bits(const uint32_t *f, int n, int limit) {
  // plain Huffman (heap), then clamp lengths > limit crudely (rare)
  int idx[320], m = 0;
  uint64_t w[640];
  int parent[640];
  for (int i = 0; i < n; i++)
    if (f[i]) { idx[m] = i; w[m] = f[i]; m++; }
  if (m == 0) return 0;
  if (m == 1) return f[idx[0]];
  int alive[640], na = m, tot = m;
  for (int i = 0; i < m; i++) alive[i] = i;
  while (na > 1) {
    int a = 0, b = 1;
    if (w[alive[b]] < w[alive[a]]) { int t = a; a = b; b = t; }
    for (int i = 2; i < na; i++) {
      if (w[alive[i]] < w[alive[a]]) { b = a; a = i; }
      else if (w[alive[i]] < w[alive[b]]) b = i;
    }
    w[tot] = w[alive[a]] + w[alive[b]];
    parent[alive[a]] = tot; parent[alive[b]] = tot;
    int hi = a > b ? a : b, lo = a > b ? b : a;
    alive[lo] = tot; alive[hi] = alive[na - 1]; na--;
    tot++;
  }
  double bits = 0;
  for (int i = 0; i < m; i++) {
    int d = 0, x = i;
    while (x != tot - 1) { x = parent[x]; d++; }
    if (d > limit) d = limit;
    bits += (double)d * f[idx[i]];
  }
  return bits;
}

static const int len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static int len_code(int l) {
  if (l == 258) return 28;
  int v = l - 3;
  if (v < 8) return v;
  int hb = 31 - __builtin_clz(v);
  return 4 * hb - 4 + ((v >> (hb - 2)) & 3);
}
static int dist_code(int d) {
  int v = d - 1;
  if (v < 4) return v;
  int hb = 31 - __builtin_clz(v);
  return 2 * hb + ((v >> (hb - 1)) & 1);
}

static uint32_t fl[286], fd[30];
static uint64_t extra_bits, n_match, n_lit, match_bytes;
static void tok_lit(uint8_t b) { fl[b]++; n_lit++; }
static void tok_match(int len, int dist) {
  int lc = len_code(len), dc = dist_code(dist);
  fl[257 + lc]++; fd[dc]++;
  extra_bits += len_extra[lc] + (dc < 4 ? 0 : (dc >> 1) - 1);
  n_match++; match_bytes += len;
}

// ---- dictionaries ----
typedef struct { uint32_t *e; int bits, ways, nbytes, slot; } Tab;   // entries: position + 1 (0 = empty), ways most recent first
static void tab_init(Tab *t, int bits, int ways, int nbytes) {
  t->bits = bits; t->ways = ways; t->nbytes = nbytes; t->slot = 0;
  t->e = calloc((size_t)ways << bits, 4);
}
static void tab_clear(Tab *t) { memset(t->e, 0, ((size_t)t->ways << t->bits) * 4); }
static inline uint32_t tab_hash(const Tab *t, size_t p) {
  return t->nbytes <= 4 ? h4(ld32(p), t->bits) : hl(ld64(p), t->nbytes, t->bits);
}
static inline void tab_push(Tab *t, size_t p) {
  uint32_t *b = t->e + (size_t)tab_hash(t, p) * t->ways;
  if (t->slot) {   // slot = window index mod ways: plain store, no read-modify-write
    b[(p >> 5) & (size_t)(t->ways - 1)] = (uint32_t)p + 1;
    return;
  }
  for (int k = t->ways - 1; k > 0; k--) b[k] = b[k - 1];
  b[0] = (uint32_t)p + 1;
}

// entries of bucket b in recency order into out[] (slotmode: by window residue, most recent first)
static int bucket_cands(const Tab *t, const uint32_t *b, size_t p, size_t *out) {
  int n = 0;
  if (!t->slot) {
    for (int k = 0; k < t->ways; k++) if (b[k]) out[n++] = b[k] - 1;
    return n;
  }
  size_t win = p >> 5;
  for (int i = 1; i <= t->ways; i++) {
    uint32_t e = b[(win - (size_t)i) & (size_t)(t->ways - 1)];
    if (e) out[n++] = e - 1;
  }
  return n;
}

static int match_len(size_t p, size_t c, int limit) {
  int m = 0;
  while (m < limit && p + m < N && D[p + m] == D[c + m]) m++;
  return m;
}

int main(int argc, char **argv) {
  if (argc < 2) return 1;
  for (int i = 2; i < argc; i++) {
    char *eq = strchr(argv[i], '=');
    if (!eq) continue;
    *eq = 0;
    int v = atoi(eq + 1);
    const char *k = argv[i];
#define OPT(n, var) if (!strcmp(k, n)) var = v
    OPT("sub", P_sub); OPT("seg", P_seg); OPT("sbits", P_sbits); OPT("slong", P_slong); OPT("sways", P_sways);
    OPT("obits", P_obits); OPT("oways", P_oways); OPT("olong", P_olong); OPT("preseed", P_preseed); OPT("hist", P_hist);
    OPT("lazy", P_lazy); OPT("min4far", P_min4far); OPT("near", P_near); OPT("good", P_good); OPT("maxcand", P_maxcand);
    OPT("minlen", P_minlen); OPT("ownstatic", P_ownstatic); OPT("visit", P_visit); OPT("back", P_back); OPT("slotmode", P_slotmode); OPT("slotcap", P_slotcap); OPT("order", P_order);
  }
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  fseek(f, 0, SEEK_END);
  N = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t *buf = malloc(N + 64);
  if (fread(buf, 1, N, f) != N) return 3;
  memset(buf + N, 0, 64);
  D = buf;
  fclose(f);

  int nseg_max = P_seg ? (CHUNK + 32768) / P_seg + 2 : 0;
  Tab *st4 = calloc(nseg_max, sizeof(Tab)), *stl = calloc(nseg_max, sizeof(Tab));
  for (int i = 0; i < nseg_max; i++) {
    tab_init(&st4[i], P_sbits, P_sways, 4);
    if (P_slong) tab_init(&stl[i], P_sbits, P_sways, P_slong);
  }
  for (int i = 0; i < nseg_max; i++) st4[i].slot = P_slotmode != 0;
  Tab own4, ownl;
  tab_init(&own4, P_obits, P_oways, 4);
  own4.slot = P_slotmode == 1;
  if (P_olong) tab_init(&ownl, P_obits, P_oways, P_olong);

  double total_bits = 0;
  uint64_t cand_evals = 0, positions = 0;
  for (size_t c0 = 0; c0 < N || c0 == 0; c0 += CHUNK) {
    size_t clen = N - c0 < CHUNK ? N - c0 : CHUNK;
    size_t hb = c0 < (size_t)P_hist ? c0 : (size_t)P_hist;
    size_t r0 = c0 - hb, r1 = c0 + clen;   // staged region
    memset(fl, 0, sizeof fl); memset(fd, 0, sizeof fd);
    extra_bits = 0;
    // phase 1: static tables of every segment of the region (segments are aligned to the region start)
    int nseg = 0;
    if (P_seg) {
      nseg = (int)((r1 - r0 + P_seg - 1) / P_seg);
      for (int s = 0; s < nseg; s++) {
        tab_clear(&st4[s]);
        if (P_slong) tab_clear(&stl[s]);
        size_t a = r0 + (size_t)s * P_seg, b = a + P_seg < r1 ? a + P_seg : r1;
        for (size_t p = a; p < b; p++) {
          if (p + 4 > r1) break;
          tab_push(&st4[s], p);
          if (P_slong) tab_push(&stl[s], p);
        }
      }
    }
    // phase 2: sub-chunks
    for (size_t b0 = c0; b0 < r1; b0 += P_sub) {
      size_t b1 = b0 + P_sub < r1 ? b0 + P_sub : r1;
      tab_clear(&own4);
      if (P_olong) tab_clear(&ownl);
      if (P_preseed) {
        size_t s = b0 - r0 < (size_t)P_preseed ? r0 : b0 - P_preseed;
        for (size_t p = s; p < b0; p++)
          if (p + 4 <= r1) { tab_push(&own4, p); if (P_olong) tab_push(&ownl, p); }
      }
      size_t entry = b0;
      for (size_t wb = b0; wb < b1; wb += 32) {
        int m[32], dist[32];
        size_t nvalid = b1 - wb < 32 ? b1 - wb : 32;
        uint32_t hh[32];
        for (int l = 0; l < 32; l++) { m[l] = 0; dist[l] = 0; hh[l] = 0xffffffffu; }
        for (int l = 0; l < 32; l++) {
          size_t p = wb + l;
          if (p + 4 > r1 || p >= b1 + 0) { if (p + 4 <= r1) hh[l] = tab_hash(&own4, p); continue; }
          hh[l] = tab_hash(&own4, p);
        }
        for (int l = 0; l < 32; l++) {
          size_t p = wb + l;
          if (p >= b1 || p + 4 > r1 || p < entry) continue;
          positions++;
          int limit = (int)(b1 - p < MAXM ? b1 - p : MAXM);
          if (limit < P_minlen) continue;
          int best = 0, bd = 0, budget = P_maxcand;
          size_t cands[64];
          int nc = 0;
          if (P_near) {
            for (int j = l - 1; j >= 0; j--)
              if (hh[j] == hh[l]) { cands[nc++] = wb + j; break; }
          }
          if (!P_ownstatic) {
            uint32_t *b = own4.e + (size_t)hh[l] * own4.ways;
            nc += bucket_cands(&own4, b, p, cands + nc);
            if (P_olong) {
              uint32_t *bl = ownl.e + (size_t)tab_hash(&ownl, p) * ownl.ways;
              for (int k = 0; k < ownl.ways; k++) if (bl[k]) cands[nc++] = bl[k] - 1;
            }
          }
          if (P_seg) {
            int sp = (int)((p - r0) / P_seg);
            int back = P_back;
            // own segment's static table only helps when sub < seg (positions of the segment before this sub-chunk)
            for (int s = ((P_sub < P_seg || P_ownstatic) ? sp : sp - 1); s >= 0 && s >= sp - back && nc < 60; s--) {
              uint32_t *b = st4[s].e + (size_t)tab_hash(&st4[s], p) * st4[s].ways;
              nc += bucket_cands(&st4[s], b, (size_t)0, cands + nc);
              if (P_slong) {
                uint32_t *bl = stl[s].e + (size_t)tab_hash(&stl[s], p) * stl[s].ways;
                for (int k = 0; k < stl[s].ways; k++) if (bl[k]) cands[nc++] = bl[k] - 1;
              }
            }
          }
          uint32_t v = ld32(p);
          int visits = P_visit;
          if (P_slotcap) {
            // the kernel's fixed slot list (empties count): slot 0 near, 1..4 own ways, then 4 per history segment
            size_t sl[21];
            int ns = 0;
            const size_t NONE = (size_t)-1;
            size_t nearc = NONE;
            for (int j = l - 1; j >= 0; j--) if (hh[j] == hh[l]) { nearc = wb + j; break; }
            size_t ownc[4] = {NONE, NONE, NONE, NONE};
            { uint32_t *b = own4.e + (size_t)hh[l] * own4.ways; size_t tmp[8]; int k = bucket_cands(&own4, b, p, tmp); for (int i = 0; i < k && i < 4; i++) ownc[i] = tmp[i]; }
            size_t hc[4][4];
            int sp = (int)((p - r0) / P_seg);
            for (int j = 0; j < 4; j++) {
              for (int i = 0; i < 4; i++) hc[j][i] = NONE;
              int sg = sp - 1 - j;
              if (sg < 0) continue;
              uint32_t *b = st4[sg].e + (size_t)tab_hash(&st4[sg], p) * st4[sg].ways;
              // kernel order: residue 3,2,1,0 regardless of emptiness
              for (int i = 0; i < 4; i++) { uint32_t e = st4[sg].slot ? b[3 - i] : b[i]; if (e) hc[j][i] = e - 1; }
            }
            if (P_order == 0) {
              sl[ns++] = nearc;
              for (int i = 0; i < 4; i++) sl[ns++] = ownc[i];
              for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) sl[ns++] = hc[j][i];
            } else {
              // interleaved: near, own0, own1, then way i of every history segment, own2, own3 after the first round
              sl[ns++] = nearc; sl[ns++] = ownc[0]; sl[ns++] = ownc[1];
              for (int j = 0; j < 4; j++) sl[ns++] = hc[j][0];
              sl[ns++] = ownc[2];
              for (int j = 0; j < 4; j++) sl[ns++] = hc[j][1];
              sl[ns++] = ownc[3];
              for (int j = 0; j < 4; j++) sl[ns++] = hc[j][2];
              for (int j = 0; j < 4; j++) sl[ns++] = hc[j][3];
            }
            nc = 0;
            for (int k = 0; k < ns && k < P_slotcap; k++) if (sl[k] != NONE) cands[nc++] = sl[k];
          }
          for (int k = 0; k < nc && budget > 0 && visits > 0; k++) {
            size_t c = cands[k];
            visits--;
            if (c >= p || p - c > 32768 || c < r0) continue;
            if (ld32(c) != v) continue;
            cand_evals++;
            budget--;
            int ml = match_len(p, c, limit < CAP ? limit : CAP);
            if (ml == CAP && limit > CAP) ml = CAP;   // lane cap
            if (ml > best) { best = ml; bd = (int)(p - c); }
            if (best >= P_good) budget = budget > 1 ? 1 : budget;
            if (best >= CAP || best >= limit) break;
          }
          if (best >= P_minlen && !(best == 4 && P_min4far && bd > P_min4far)) { m[l] = best; dist[l] = bd; }
        }
        // insert this window into the own tables (in order)
        for (int l = 0; l < 32; l++) {
          size_t p = wb + l;
          if (p + 4 <= r1 && p < b1 + 0) { tab_push(&own4, p); if (P_olong) tab_push(&ownl, p); }
        }
        // lazy
        if (P_lazy)
          for (int l = 0; l < 31; l++)
            if (m[l] && m[l] < P_lazy && m[l + 1] > m[l]) m[l] = 0;
        // greedy chain
        size_t cur = entry > wb ? entry - wb : 0;
        size_t endw = 0;
        size_t pos = cur;
        while (pos < nvalid) {
          if (m[pos]) {
            int ml = m[pos];
            size_t p = wb + pos;
            if (ml >= CAP) {
              int limit = (int)(b1 - p < MAXM ? b1 - p : MAXM);
              ml = match_len(p, p - dist[pos], limit);
            }
            tok_match(ml, dist[pos]);
            pos += ml;
            endw = pos;
          } else {
            tok_lit(D[wb + pos]);
            pos++;
          }
        }
        if (entry < wb + nvalid) entry = wb + (endw > nvalid ? endw : nvalid);
      }
    }
    fl[256]++;
    double bits = huff_bits(fl, 286, 15) + huff_bits(fd, 30, 15) + extra_bits + 90 * 8;
    double stored = (double)(clen + 10) * 8;
    total_bits += bits < stored ? bits : stored;
    if (N == 0) break;
  }
  printf("%s size=%zu est=%.0f ratio=%.4f matches=%llu lits=%llu avgmatch=%.2f cand_per_pos=%.2f\n", argv[1], N, total_bits / 8,
         total_bits / 8 / (double)(N ? N : 1), (unsigned long long)n_match, (unsigned long long)n_lit,
         n_match ? (double)match_bytes / n_match : 0.0, positions ? (double)cand_evals / positions : 0.0);
  return 0;
}
