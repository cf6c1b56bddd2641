/*
 * zippy_oracle.c -- CPU oracle for the zippy-b200 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of guzba/zippy v0.10.18's codec core.  See
 * zippy_oracle.h for the rules around its use.  Citations are to the
 * reference checkout (src/zippy.nim, src/zippy/<file>.nim).
 *
 * Documented deviations from the reference (all in undefined / latent-bug
 * territory, none reachable from a well-formed stream):
 *  - bit reader tail: the reference reloads the last 8 bytes of the buffer
 *    and shifts (bitstreams.nim:41-43), which indexes out of bounds for
 *    len < 8 and shifts by 64 when no bytes remain.  Here bytes past the end
 *    read as zero and running past the end is reported at the same check
 *    points the reference uses (bitsBuffered < 0) plus the block-header and
 *    stored-length reads, which the reference leaves unchecked.
 *  - huffmanCodes' canonical-code histogram is uint8 in the reference
 *    (deflate.nim:136) and wraps at 256 equal lengths with checks off; ints here.
 *  - std/heapqueue tie-breaking is restated from the CPython heapq algorithm
 *    Nim's module is a port of; nothing in the reference pins it.
 */
#include "zippy_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* internal.nim:9-24 constants                                         */
/* ------------------------------------------------------------------ */
#define MAX_CODE_LENGTH 15
#define MAX_LITLEN_CODES 286
#define MAX_DIST_CODES 30
#define MAX_FIXED_LITLEN_CODES 288
#define MAX_WINDOW_SIZE 32768
#define MAX_UNCOMPRESSED_BLOCK 65535
#define MAX_BLOCK_SIZE 4194304
#define FIRST_LENGTH_CODE 257
#define BASE_MATCH_LEN 3
#define MIN_MATCH_LEN 4
#define MAX_MATCH_LEN 258
#define MAX_LITERAL_LENGTH 32767 /* internal.nim:24 */

/* internal.nim:26-44 */
static const uint16_t base_lengths[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,
                                          15, 17, 19, 23, 27, 31, 35, 43, 51,  59,
                                          67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t base_lengths_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                               2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
/* internal.nim:75-107 */
static const uint16_t base_distances[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,
                                            33,  49,  65,  97,  129, 193,  257,  385,  513,  769,
                                            1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t base_distance_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2,  3,  3,  4,  4,  5,  5,  6,
                                                6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
/* internal.nim:109-111 */
static const uint8_t clcl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

/* internal.nim:46-73 baseLengthIndices[len-3]: derived from base_lengths
 * instead of the literal 256-entry table. */
static uint8_t base_length_indices[256];
/* internal.nim:224-249 distanceCodeIndex: derived from base_distances. */
static uint8_t dist_code_of[32768];
/* internal.nim:151-175 fixed codes */
static uint8_t fixed_litlen_lengths[MAX_FIXED_LITLEN_CODES];
static uint16_t fixed_litlen_codes[MAX_FIXED_LITLEN_CODES];
static uint8_t fixed_dist_lengths[MAX_DIST_CODES];
static uint16_t fixed_dist_codes[MAX_DIST_CODES];
/* crc.nim:6-23 */
static uint32_t crc_tables[8][256];

static pthread_once_t tables_once = PTHREAD_ONCE_INIT;

static uint16_t reverse_bits16(uint16_t v) {
  v = (uint16_t)(((v & 0x5555u) << 1) | ((v >> 1) & 0x5555u));
  v = (uint16_t)(((v & 0x3333u) << 2) | ((v >> 2) & 0x3333u));
  v = (uint16_t)(((v & 0x0f0fu) << 4) | ((v >> 4) & 0x0f0fu));
  return (uint16_t)((v << 8) | (v >> 8));
}

/* internal.nim:133-149 makeCodes */
static void make_codes(const uint8_t *lengths, int n, uint16_t *codes) {
  int counts[16] = {0};
  uint16_t next_code[16] = {0};
  for (int i = 0; i < n; i++) counts[lengths[i]]++;
  counts[0] = 0;
  for (int i = 1; i <= MAX_CODE_LENGTH; i++)
    next_code[i] = (uint16_t)((next_code[i - 1] + counts[i - 1]) << 1);
  for (int i = 0; i < n; i++) {
    codes[i] = 0;
    if (lengths[i] != 0) {
      codes[i] = (uint16_t)(reverse_bits16(next_code[lengths[i]]) >> (16 - lengths[i]));
      next_code[lengths[i]]++;
    }
  }
}

static void init_tables(void) {
  for (int idx = 0; idx < 29; idx++) {
    int lo = base_lengths[idx];
    int hi = (idx == 28) ? 258 : (idx == 27 ? 257 : base_lengths[idx + 1] - 1);
    for (int l = lo; l <= hi; l++) base_length_indices[l - 3] = (uint8_t)idx;
  }
  base_length_indices[258 - 3] = 28;
  for (int idx = 0; idx < 30; idx++) {
    int lo = base_distances[idx];
    int hi = (idx == 29) ? 32768 : base_distances[idx + 1] - 1;
    for (int d = lo; d <= hi; d++) dist_code_of[d - 1] = (uint8_t)idx;
  }
  for (int i = 0; i < MAX_FIXED_LITLEN_CODES; i++)
    fixed_litlen_lengths[i] = (uint8_t)(i <= 143 ? 8 : i <= 255 ? 9 : i <= 279 ? 7 : 8);
  make_codes(fixed_litlen_lengths, MAX_FIXED_LITLEN_CODES, fixed_litlen_codes);
  for (int i = 0; i < MAX_DIST_CODES; i++) fixed_dist_lengths[i] = 5;
  make_codes(fixed_dist_lengths, MAX_DIST_CODES, fixed_dist_codes);
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int j = 0; j < 8; j++) c = (c >> 1) ^ ((c & 1) * 0xedb88320u);
    crc_tables[0][i] = c;
  }
  for (int i = 0; i < 256; i++)
    for (int t = 1; t < 8; t++)
      crc_tables[t][i] = (crc_tables[t - 1][i] >> 8) ^ crc_tables[0][crc_tables[t - 1][i] & 255];
}

static void ensure_tables(void) { pthread_once(&tables_once, init_tables); }

/* ------------------------------------------------------------------ */
/* buffers                                                             */
/* ------------------------------------------------------------------ */
void zo_buf_free(zo_buf *b) {
  free(b->data);
  b->data = NULL;
  b->len = b->cap = 0;
}

/* Grow to hold at least `need` bytes; new bytes are zero (Nim setLen zero-fills,
 * which BitStreamWriter.addBits relies on: bitstreams.nim:96-103). */
static int buf_reserve(zo_buf *b, size_t need) {
  if (need <= b->cap) return 0;
  size_t ncap = b->cap ? b->cap : 64;
  while (ncap < need) ncap *= 2;
  uint8_t *p = (uint8_t *)realloc(b->data, ncap);
  if (!p) return -1;
  memset(p + b->cap, 0, ncap - b->cap);
  b->data = p;
  b->cap = ncap;
  return 0;
}

static inline uint32_t rd32(const uint8_t *p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
static inline uint64_t rd64(const uint8_t *p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
}

/* ------------------------------------------------------------------ */
/* checksums                                                           */
/* ------------------------------------------------------------------ */
/* crc.nim:29-51 slice-by-8 with a running (pre-inverted) state. */
static uint32_t crc32_scalar_state(const uint8_t *src, size_t len, uint32_t crc) {
  size_t i = 0;
  for (size_t n = len / 8; n > 0; n--) {
    uint32_t one = rd32(src + i) ^ crc;
    uint32_t two = rd32(src + i + 4);
    crc = crc_tables[7][one & 255] ^ crc_tables[6][(one >> 8) & 255] ^
          crc_tables[5][(one >> 16) & 255] ^ crc_tables[4][one >> 24] ^
          crc_tables[3][two & 255] ^ crc_tables[2][(two >> 8) & 255] ^
          crc_tables[1][(two >> 16) & 255] ^ crc_tables[0][two >> 24];
    i += 8;
  }
  for (; i < len; i++) crc = crc_tables[0][(crc ^ src[i]) & 255] ^ (crc >> 8);
  return crc;
}
uint32_t zo_crc32_scalar(const uint8_t *src, size_t len) {
  ensure_tables();
  return ~crc32_scalar_state(src, len, ~0u);
}

#if defined(__x86_64__)
#include <immintrin.h>
/* crc32_simd.nim:39-144 (crc32_sse41_pcmul): carry-less-multiply folding of four 128-bit
 * lanes (constants :43-46 are x^(512+-32), x^(128+-32), x^64 mod P and the Barrett pair), then
 * 128 -> 64 -> 32 bit reduction.  len >= 64 and a multiple of 16; crc is the running state. */
__attribute__((target("sse4.1,pclmul")))
static uint32_t crc32_pclmul_state(const uint8_t *p, size_t len, uint32_t crc) {
  const __m128i k12 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll);
  const __m128i k34 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll);
  const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124ll);
  const __m128i mu = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);
  const __m128i lo32 = _mm_setr_epi32(~0, 0, ~0, 0);
  __m128i a = _mm_loadu_si128((const __m128i *)(p + 0)), b = _mm_loadu_si128((const __m128i *)(p + 16));
  __m128i c = _mm_loadu_si128((const __m128i *)(p + 32)), d = _mm_loadu_si128((const __m128i *)(p + 48));
  a = _mm_xor_si128(a, _mm_cvtsi32_si128((int)crc));
  p += 64;
  len -= 64;
#define ZO_FOLD(x, k, in) \
  _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x, k, 0x00), _mm_clmulepi64_si128(x, k, 0x11)), in)
  while (len >= 64) { /* :67-92 */
    a = ZO_FOLD(a, k12, _mm_loadu_si128((const __m128i *)(p + 0)));
    b = ZO_FOLD(b, k12, _mm_loadu_si128((const __m128i *)(p + 16)));
    c = ZO_FOLD(c, k12, _mm_loadu_si128((const __m128i *)(p + 32)));
    d = ZO_FOLD(d, k12, _mm_loadu_si128((const __m128i *)(p + 48)));
    p += 64;
    len -= 64;
  }
  a = ZO_FOLD(a, k34, b); /* :94-110 */
  a = ZO_FOLD(a, k34, c);
  a = ZO_FOLD(a, k34, d);
  while (len >= 16) { /* :112-121 */
    a = ZO_FOLD(a, k34, _mm_loadu_si128((const __m128i *)p));
    p += 16;
    len -= 16;
  }
#undef ZO_FOLD
  /* :123-142 */
  __m128i t = _mm_clmulepi64_si128(a, k34, 0x10);
  a = _mm_xor_si128(_mm_srli_si128(a, 8), t);
  t = _mm_srli_si128(a, 4);
  a = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(a, lo32), k5, 0x00), t);
  t = _mm_and_si128(a, lo32);
  t = _mm_clmulepi64_si128(t, mu, 0x10);
  t = _mm_and_si128(t, lo32);
  t = _mm_clmulepi64_si128(t, mu, 0x00);
  a = _mm_xor_si128(a, t);
  return (uint32_t)_mm_extract_epi32(a, 1);
}

/* adler32_simd.nim:45-120 (adler32_ssse3): 32-byte blocks, sad for s1, maddubs taps 32..1 for
 * s2, modulo every nmax/32 blocks. */
__attribute__((target("ssse3")))
static uint32_t adler32_ssse3(const uint8_t *src, size_t len) {
  const uint32_t block = 32, nmax = 5552;
  uint32_t s1 = 1, s2 = 0;
  size_t pos = 0;
  size_t blocks = len / block, remaining = len - blocks * block;
  const __m128i tap1 = _mm_setr_epi8(32, 31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17);
  const __m128i tap2 = _mm_setr_epi8(16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1);
  const __m128i zero = _mm_setzero_si128(), ones = _mm_set1_epi16(1);
  while (blocks > 0) {
    size_t n = nmax / block;
    if (n > blocks) n = blocks;
    blocks -= n;
    __m128i ps = _mm_set_epi32(0, 0, 0, (int)(s1 * (uint32_t)n));
    __m128i v2 = _mm_set_epi32(0, 0, 0, (int)s2), v1 = zero;
    for (; n > 0; n--, pos += 32) {
      const __m128i b1 = _mm_loadu_si128((const __m128i *)(src + pos));
      const __m128i b2 = _mm_loadu_si128((const __m128i *)(src + pos + 16));
      ps = _mm_add_epi32(ps, v1);
      v1 = _mm_add_epi32(v1, _mm_sad_epu8(b1, zero));
      v2 = _mm_add_epi32(v2, _mm_madd_epi16(_mm_maddubs_epi16(b1, tap1), ones));
      v1 = _mm_add_epi32(v1, _mm_sad_epu8(b2, zero));
      v2 = _mm_add_epi32(v2, _mm_madd_epi16(_mm_maddubs_epi16(b2, tap2), ones));
    }
    v2 = _mm_add_epi32(v2, _mm_slli_epi32(ps, 5));
    v1 = _mm_add_epi32(v1, _mm_shuffle_epi32(v1, _MM_SHUFFLE(2, 3, 0, 1)));
    v1 = _mm_add_epi32(v1, _mm_shuffle_epi32(v1, _MM_SHUFFLE(1, 0, 3, 2)));
    s1 += (uint32_t)_mm_cvtsi128_si32(v1);
    v2 = _mm_add_epi32(v2, _mm_shuffle_epi32(v2, _MM_SHUFFLE(2, 3, 0, 1)));
    v2 = _mm_add_epi32(v2, _mm_shuffle_epi32(v2, _MM_SHUFFLE(1, 0, 3, 2)));
    s2 = (uint32_t)_mm_cvtsi128_si32(v2);
    s1 %= 65521;
    s2 %= 65521;
  }
  for (size_t i = 0; i < remaining; i++) {
    s1 += src[pos + i];
    s2 += s1;
  }
  s1 %= 65521;
  s2 %= 65521;
  return (s2 << 16) | s1;
}
static int cpu_has_pclmul(void) { return __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("pclmul"); }
static int cpu_has_ssse3(void) { return __builtin_cpu_supports("ssse3"); }
#else
static int cpu_has_pclmul(void) { return 0; }
static int cpu_has_ssse3(void) { return 0; }
#endif

/* crc.nim:53-72: the SIMD prefix (16-byte multiple, len >= 64, amd64 with SSE4.1 + PCLMUL),
 * then the scalar tail -- what the reference really runs on the bench hosts. */
uint32_t zo_crc32(const uint8_t *src, size_t len) {
  ensure_tables();
  uint32_t crc = ~0u;
  size_t pos = 0;
#if defined(__x86_64__)
  if (len >= 64 && cpu_has_pclmul()) {
    const size_t simd_len = (len / 16) * 16;
    crc = crc32_pclmul_state(src, simd_len, crc);
    pos = simd_len;
  }
#endif
  if (pos < len) crc = crc32_scalar_state(src + pos, len - pos, crc);
  return ~crc;
}

/* adler32.nim:17-63 scalar form (NMAX = 5552 deferred modulo). */
uint32_t zo_adler32_scalar(const uint8_t *src, size_t len) {
  const size_t nmax = 5552;
  uint32_t s1 = 1, s2 = 0;
  size_t pos = 0, l = len;
  while (l >= nmax) {
    l -= nmax;
    for (size_t i = 0; i < nmax; i++) {
      s1 += src[pos + i];
      s2 += s1;
    }
    pos += nmax;
    s1 %= 65521;
    s2 %= 65521;
  }
  for (size_t i = 0; i < l; i++) {
    s1 += src[pos + i];
    s2 += s1;
  }
  s1 %= 65521;
  s2 %= 65521;
  return (s2 << 16) | s1;
}
/* adler32.nim:6-15: the SSSE3 path is always taken on amd64 when the CPU has SSSE3. */
uint32_t zo_adler32(const uint8_t *src, size_t len) {
#if defined(__x86_64__)
  if (cpu_has_ssse3() && len <= 0xffffffffu) return len ? adler32_ssse3(src, len) : 1u;
#endif
  return zo_adler32_scalar(src, len);
}

/* ------------------------------------------------------------------ */
/* bit writer: bitstreams.nim:13-14, 84-123                            */
/* ------------------------------------------------------------------ */
typedef struct {
  size_t pos;
  int bit_pos;
} bit_writer;

/* bitstreams.nim:88-104 addBits */
static int add_bits(bit_writer *b, zo_buf *dst, uint32_t value, int bit_len) {
  if (buf_reserve(dst, b->pos + 8) != 0) return ZO_ERR_NOMEM;
  uint64_t v = (uint64_t)value & (((uint64_t)1 << bit_len) - 1);
  uint64_t cur = rd32(dst->data + b->pos);
  cur |= v << b->bit_pos;
  memcpy(dst->data + b->pos, &cur, 8);
  b->pos += (size_t)((bit_len + b->bit_pos) >> 3);
  b->bit_pos = (bit_len + b->bit_pos) & 7;
  return 0;
}

/* bitstreams.nim:121-123 */
static void writer_skip_to_byte(bit_writer *b) {
  if (b->bit_pos > 0) {
    b->pos += 1;
    b->bit_pos = 0;
  }
}

/* bitstreams.nim:106-119 addBytes */
static int add_bytes(bit_writer *b, zo_buf *dst, const uint8_t *src, size_t len) {
  if (b->bit_pos != 0) return ZO_ERR_BYTE_BOUNDARY;
  if (buf_reserve(dst, b->pos + len + 8) != 0) return ZO_ERR_NOMEM;
  memcpy(dst->data + b->pos, src, len);
  b->pos += len;
  return 0;
}

/* ------------------------------------------------------------------ */
/* token stream + block metadata: internal.nim:128-131, lz77.nim:19-50 */
/* ------------------------------------------------------------------ */
typedef struct {
  uint32_t litlen_freq[MAX_LITLEN_CODES];
  uint32_t dist_freq[MAX_DIST_CODES];
  size_t num_literals;
} block_meta;

typedef struct {
  uint16_t *v;
  size_t len, cap;
} enc_vec;

static int enc_push(enc_vec *e, uint16_t x) {
  if (e->len == e->cap) {
    size_t nc = e->cap ? e->cap * 2 : 1024;
    uint16_t *p = (uint16_t *)realloc(e->v, nc * sizeof(uint16_t));
    if (!p) return -1;
    e->v = p;
    e->cap = nc;
  }
  e->v[e->len++] = x;
  return 0;
}

/* snappy.nim:33-47 / lz77.nim:19-33 addLiteral */
static int add_literal(enc_vec *e, block_meta *m, const uint8_t *src, size_t start, size_t length) {
  for (size_t i = 0; i < length; i++) m->litlen_freq[src[start + i]]++;
  m->num_literals += length;
  size_t remaining = length;
  while (remaining > 0) {
    size_t added = remaining < MAX_LITERAL_LENGTH ? remaining : MAX_LITERAL_LENGTH;
    if (enc_push(e, (uint16_t)added)) return -1;
    remaining -= added;
  }
  return 0;
}

/* snappy.nim:49-64 / lz77.nim:35-50 addCopy */
static int add_copy(enc_vec *e, block_meta *m, size_t offset, size_t length) {
  uint16_t length_index = base_length_indices[length - BASE_MATCH_LEN];
  uint16_t dist_index = dist_code_of[offset - 1];
  m->litlen_freq[length_index + FIRST_LENGTH_CODE]++;
  m->dist_freq[dist_index]++;
  if (enc_push(e, (uint16_t)(((length_index << 8) | dist_index) | (1u << 15)))) return -1;
  if (enc_push(e, (uint16_t)offset)) return -1;
  if (enc_push(e, (uint16_t)length)) return -1;
  return 0;
}

/* internal.nim:251-270 determineMatchLength */
static inline size_t determine_match_length(const uint8_t *src, size_t s1, size_t s2, size_t limit) {
  size_t result = 0;
  while (s2 + 8 <= limit) {
    uint64_t x = rd64(src + s2) ^ rd64(src + s1 + result);
    if (x != 0) return result + ((size_t)__builtin_ctzll(x) >> 3);
    s2 += 8;
    result += 8;
  }
  while (s2 < limit) {
    if (src[s2] != src[s1 + result]) return result;
    s2++;
    result++;
  }
  return result;
}

/* ------------------------------------------------------------------ */
/* level 1 matcher: snappy.nim:12-136 encodeFragment, :138-163 encodeSnappy */
/* ------------------------------------------------------------------ */
#define MAX_COMPRESS_TABLE 16384 /* snappy.nim:7 */

static int encode_fragment(enc_vec *e, block_meta *m, const uint8_t *src, size_t start,
                           size_t bytes_to_read, uint16_t *table) {
  const size_t ip_end = start + bytes_to_read;
  size_t ip = start, next_emit = start;
  size_t table_size = 256;
  int shift = 24;
  while (table_size < MAX_COMPRESS_TABLE && table_size < bytes_to_read) { /* :27-29 */
    table_size <<= 1;
    shift--;
  }
  memset(table, 0, table_size * sizeof(uint16_t)); /* :31 */
#define SN_HASH(v) ((uint32_t)((uint32_t)(v) * 0x1e35a7bdu) >> shift) /* :70-71 */
  if (bytes_to_read >= 15) { /* :76 */
    const size_t ip_limit = start + bytes_to_read - 15;
    ip++;
    uint32_t next_hash = SN_HASH(rd32(src + ip));
    for (;;) {
      size_t skip_bytes = 32, next_ip = ip, candidate;
      for (;;) { /* :86-101 probe / skip loop */
        ip = next_ip;
        uint32_t h = next_hash;
        size_t between = skip_bytes >> 5;
        skip_bytes++;
        next_ip = ip + between;
        if (next_ip > ip_limit) goto emit_remainder;
        next_hash = SN_HASH(rd32(src + next_ip));
        candidate = start + table[h];
        table[h] = (uint16_t)(ip - start);
        if (rd32(src + ip) == rd32(src + candidate)) break;
      }
      if (add_literal(e, m, src, next_emit, ip - next_emit)) return -1; /* :103 */
      uint64_t input_bytes;
      for (;;) { /* :108-131 match-extend loop */
        size_t limit = ip_end < ip + MAX_MATCH_LEN ? ip_end : ip + MAX_MATCH_LEN;
        size_t matched = 4 + determine_match_length(src, candidate + 4, ip + 4, limit);
        size_t offset = ip - candidate;
        ip += matched;
        if (add_copy(e, m, offset, matched)) return -1;
        size_t insert_tail = ip - 1;
        next_emit = ip;
        if (ip >= ip_limit) goto emit_remainder;
        input_bytes = rd64(src + insert_tail);
        uint32_t prev_hash = SN_HASH((uint32_t)input_bytes);
        uint32_t cur_hash = SN_HASH((uint32_t)(input_bytes >> 8));
        table[prev_hash] = (uint16_t)(ip - start - 1);
        candidate = start + table[cur_hash];
        uint32_t candidate_bytes = rd32(src + candidate);
        table[cur_hash] = (uint16_t)(ip - start);
        if ((uint32_t)(input_bytes >> 8) != candidate_bytes) break;
      }
      next_hash = SN_HASH((uint32_t)(input_bytes >> 16)); /* :133 */
      ip++;
    }
  }
emit_remainder:
  if (next_emit < ip_end) /* :66-68 */
    if (add_literal(e, m, src, next_emit, ip_end - next_emit)) return -1;
  return 0;
#undef SN_HASH
}

static int encode_snappy(enc_vec *e, block_meta *m, const uint8_t *src, size_t block_start,
                         size_t block_len) {
  m->litlen_freq[256] = 1; /* :145 */
  uint16_t *table = (uint16_t *)malloc(MAX_COMPRESS_TABLE * sizeof(uint16_t));
  if (!table) return -1;
  size_t pos = block_start;
  int rc = 0;
  while (pos < block_start + block_len) { /* :150-163 */
    size_t fragment = block_start + block_len - pos;
    size_t n = fragment < MAX_WINDOW_SIZE ? fragment : MAX_WINDOW_SIZE;
    rc = encode_fragment(e, m, src, pos, n, table);
    if (rc) break;
    pos += n;
  }
  free(table);
  return rc;
}

/* ------------------------------------------------------------------ */
/* levels -1, 2..9: lz77.nim:10-130 encodeLz77                         */
/* ------------------------------------------------------------------ */
typedef struct {
  int good, lazy, nice, chain;
} comp_config;

/* internal.nim:177-189 configurationTable */
static const comp_config config_table[10] = {
    {0, 0, 0, 0},      {4, 4, 8, 4},      {4, 5, 16, 8},      {4, 6, 32, 32},      {4, 4, 16, 16},
    {8, 16, 32, 32},   {8, 16, 128, 128}, {8, 32, 256, 256},  {32, 128, 258, 1024}, {32, 258, 258, 4096}};

#define LZ_HASH_BITS 17 /* lz77.nim:4 */

static int encode_lz77(enc_vec *e, comp_config cfg, block_meta *m, const uint8_t *src,
                       size_t block_start, size_t block_len) {
  m->litlen_freq[256] = 1; /* :52 */
  if (MIN_MATCH_LEN >= block_len) /* :54-56 */
    return add_literal(e, m, src, block_start, block_len);
  const size_t block_end = block_start + block_len;
  uint16_t *head = (uint16_t *)calloc((size_t)1 << LZ_HASH_BITS, sizeof(uint16_t)); /* :63 */
  uint16_t *chain = (uint16_t *)calloc(MAX_WINDOW_SIZE, sizeof(uint16_t));           /* :64 */
  if (!head || !chain) {
    free(head);
    free(chain);
    return -1;
  }
  int rc = 0;
  size_t pos = block_start, literal_len = 0;
#define LZ_HASH4(p) ((uint32_t)(rd32(src + (p)) * 0x1e35a7bdu) >> (32 - LZ_HASH_BITS)) /* :66-67 */
  while (pos < block_end) {
    if (pos + MIN_MATCH_LEN >= block_end) { /* :74-76 */
      rc = add_literal(e, m, src, pos - literal_len, block_end - pos + literal_len);
      break;
    }
    uint16_t window_pos = (uint16_t)((pos - block_start) & (MAX_WINDOW_SIZE - 1)); /* :78 */
    uint32_t hash = LZ_HASH4(pos);
    chain[window_pos] = head[hash]; /* :69-71 updateChain */
    head[hash] = window_pos;

    uint16_t hash_pos = chain[window_pos];
    size_t limit = block_end < pos + MAX_MATCH_LEN ? block_end : pos + MAX_MATCH_LEN;
    int tries = cfg.chain;
    long prev_offset = 0, longest_offset = 0, longest_len = 0;
    while (tries > 0 && hash_pos != 0) { /* :88-112 chain walk */
      tries--;
      long offset;
      if (hash_pos <= window_pos)
        offset = (long)window_pos - (long)hash_pos;
      else
        offset = (long)window_pos - (long)hash_pos + MAX_WINDOW_SIZE;
      if (offset <= 0 || offset < prev_offset) break; /* :97-98 */
      prev_offset = offset;
      long match_len = (long)determine_match_length(src, pos - (size_t)offset, pos, limit);
      if (match_len > longest_len) {
        if (match_len >= cfg.good) tries >>= 2; /* :104-105 */
        longest_len = match_len;
        longest_offset = offset;
      }
      if (longest_len >= cfg.nice || hash_pos == chain[hash_pos]) break; /* :109 */
      hash_pos = chain[hash_pos];
    }
    if (longest_len > MIN_MATCH_LEN) { /* :114 */
      if (literal_len > 0) {
        rc = add_literal(e, m, src, pos - literal_len, literal_len);
        if (rc) break;
        literal_len = 0;
      }
      rc = add_copy(e, m, (size_t)longest_offset, (size_t)longest_len);
      if (rc) break;
      for (long i = 1; i < longest_len; i++) { /* :121-126 */
        pos++;
        window_pos = (uint16_t)(pos & (MAX_WINDOW_SIZE - 1));
        if (pos + MIN_MATCH_LEN < block_end) {
          hash = LZ_HASH4(pos);
          chain[window_pos] = head[hash];
          head[hash] = window_pos;
        }
      }
    } else {
      literal_len++;
    }
    pos++;
  }
#undef LZ_HASH4
  free(head);
  free(chain);
  return rc;
}

/* deflate.nim:153-177 encodeAllLiterals (level -2) */
static int encode_all_literals(enc_vec *e, block_meta *m, const uint8_t *src, size_t start,
                               size_t len) {
  for (size_t i = 0; i < len; i++) m->litlen_freq[src[start + i]]++;
  size_t a = len / MAX_LITERAL_LENGTH, b = len % MAX_LITERAL_LENGTH;
  for (size_t i = 0; i < a; i++)
    if (enc_push(e, (uint16_t)MAX_LITERAL_LENGTH)) return -1;
  if (b > 0)
    if (enc_push(e, (uint16_t)b)) return -1;
  m->litlen_freq[256] = 1;
  m->num_literals = len;
  return 0;
}

/* ------------------------------------------------------------------ */
/* huffmanCodes: deflate.nim:13-151                                    */
/* ------------------------------------------------------------------ */
typedef struct {
  int symbol; /* -1 = internal */
  long freq;  /* re-used for depth after the tree walk (deflate.nim:71) */
  int left, right;
} hnode;

/* std/heapqueue is a port of CPython's heapq; compare = freq only (deflate.nim:10-11). */
static void heap_siftdown(int *heap, const hnode *nodes, int startpos, int pos) {
  int newitem = heap[pos];
  while (pos > startpos) {
    int parentpos = (pos - 1) >> 1;
    int parent = heap[parentpos];
    if (nodes[newitem].freq < nodes[parent].freq) {
      heap[pos] = parent;
      pos = parentpos;
      continue;
    }
    break;
  }
  heap[pos] = newitem;
}
static void heap_siftup(int *heap, const hnode *nodes, int n, int pos) {
  int startpos = pos, newitem = heap[pos];
  int childpos = 2 * pos + 1;
  while (childpos < n) {
    int rightpos = childpos + 1;
    if (rightpos < n && !(nodes[heap[childpos]].freq < nodes[heap[rightpos]].freq)) childpos = rightpos;
    heap[pos] = heap[childpos];
    pos = childpos;
    childpos = 2 * pos + 1;
  }
  heap[pos] = newitem;
  heap_siftdown(heap, nodes, startpos, pos);
}
static void heap_push(int *heap, int *n, const hnode *nodes, int item) {
  heap[*n] = item;
  (*n)++;
  heap_siftdown(heap, nodes, 0, *n - 1);
}
static int heap_pop(int *heap, int *n, const hnode *nodes) {
  int last = heap[--(*n)];
  if (*n > 0) {
    int ret = heap[0];
    heap[0] = last;
    heap_siftup(heap, nodes, *n, 0);
    return ret;
  }
  return last;
}

/* deflate.nim:103-121 quickSort on the leaf depths */
static void quick_sort_nodes(hnode *nodes, int *order, int inl, int inr) {
  int r = inr, l = inl;
  int n = r - l + 1;
  if (n < 2) return;
  long p = nodes[order[l + 3 * n / 4]].freq;
  while (l <= r) {
    if (nodes[order[l]].freq < p) {
      l++;
    } else if (nodes[order[r]].freq > p) {
      r--;
    } else {
      int t = order[l];
      order[l] = order[r];
      order[r] = t;
      l++;
      r--;
    }
  }
  quick_sort_nodes(nodes, order, inl, r);
  quick_sort_nodes(nodes, order, l, inr);
}

/* Returns numCodes; codes/lens must hold max(nfreq, min_codes+1) entries. */
static int huffman_codes(const uint32_t *freqs, int nfreq, int min_codes, int limit,
                         uint16_t *codes, uint8_t *lens) {
  int highest = 0, used = 0;
  for (int s = 0; s < nfreq; s++)
    if (freqs[s] > 0) {
      highest = s;
      used++;
    }
  int num_codes = (highest > min_codes ? highest : min_codes) + 1; /* :30 */
  memset(codes, 0, sizeof(uint16_t) * (size_t)num_codes);
  memset(lens, 0, (size_t)num_codes);
  if (used == 0) { /* :34-36 */
    lens[0] = 1;
    lens[1] = 1;
  } else if (used == 1) { /* :37-45 */
    for (int i = 0; i < nfreq; i++)
      if (freqs[i] != 0) {
        lens[i] = 1;
        if (i == 0)
          lens[1] = 1;
        else
          lens[0] = 1;
        break;
      }
  } else {
    hnode nodes[2 * MAX_FIXED_LITLEN_CODES];
    int heap[MAX_FIXED_LITLEN_CODES];
    int order[MAX_FIXED_LITLEN_CODES];
    int nleaf = 0, nn, hn = 0;
    for (int s = 0; s < nfreq; s++)
      if (freqs[s] > 0) {
        nodes[nleaf].symbol = s;
        nodes[nleaf].freq = (long)freqs[s];
        nodes[nleaf].left = nodes[nleaf].right = -1;
        nleaf++;
      }
    nn = nleaf;
    for (int i = 0; i < nleaf; i++) heap_push(heap, &hn, nodes, i); /* :54-55 */
    while (hn >= 2) {                                              /* :57-64 */
      int l = heap_pop(heap, &hn, nodes);
      int r = heap_pop(heap, &hn, nodes);
      nodes[nn].symbol = -1;
      nodes[nn].left = l;
      nodes[nn].right = r;
      nodes[nn].freq = nodes[l].freq + nodes[r].freq;
      heap_push(heap, &hn, nodes, nn);
      nn++;
    }
    /* :66-75 depth walk (iterative) */
    int needs_limit = 0;
    {
      int stack_n[2 * MAX_FIXED_LITLEN_CODES], stack_l[2 * MAX_FIXED_LITLEN_CODES], sp = 0;
      stack_n[sp] = heap[0];
      stack_l[sp++] = 0;
      while (sp > 0) {
        sp--;
        int nd = stack_n[sp], lv = stack_l[sp];
        if (nodes[nd].symbol == -1) {
          stack_n[sp] = nodes[nd].right;
          stack_l[sp++] = lv + 1;
          stack_n[sp] = nodes[nd].left;
          stack_l[sp++] = lv + 1;
        } else {
          nodes[nd].freq = lv;
          if (lv > limit) needs_limit = 1;
        }
      }
    }
    if (needs_limit) { /* :78-131 */
      long longest = 0;
      for (int i = 0; i < nleaf; i++)
        if (nodes[i].freq > longest) longest = nodes[i].freq;
      long hist[2 * MAX_FIXED_LITLEN_CODES];
      memset(hist, 0, sizeof(hist));
      for (int i = 0; i < nleaf; i++) hist[nodes[i].freq]++;
      long i = longest;
      while (i > limit) { /* :87-101 */
        if (hist[i] == 0) {
          i--;
          continue;
        }
        long j = i - 2;
        while (j > 0 && hist[j] == 0) j--;
        hist[i] -= 2;
        hist[i - 1]++;
        hist[j + 1] += 2;
        hist[j]--;
      }
      for (int k = 0; k < nleaf; k++) order[k] = k;
      quick_sort_nodes(nodes, order, 0, nleaf - 1); /* :123 */
      int code_len = 1;
      for (int k = 0; k < nleaf; k++) { /* :125-131 */
        while (hist[code_len] == 0) code_len++;
        nodes[order[k]].freq = code_len;
        hist[code_len]--;
      }
    }
    for (int i = 0; i < nleaf; i++) lens[nodes[i].symbol] = (uint8_t)nodes[i].freq; /* :133-134 */
  }
  make_codes(lens, num_codes, codes); /* :136-149 (same canonical + bit-reverse form) */
  return num_codes;
}

/* ------------------------------------------------------------------ */
/* deflate driver: deflate.nim:179-205, 207-467                        */
/* ------------------------------------------------------------------ */
static int add_no_compression_block(bit_writer *b, zo_buf *dst, const uint8_t *src,
                                    size_t block_start, size_t block_len, int final_block) {
  size_t count = (block_len + MAX_UNCOMPRESSED_BLOCK - 1) / MAX_UNCOMPRESSED_BLOCK;
  if (count < 1) count = 1; /* :186-189 */
  for (size_t k = 0; k < count; k++) {
    int last = (k == count - 1);
    size_t ustart = block_start + k * MAX_UNCOMPRESSED_BLOCK;
    size_t ulen = block_start + block_len - ustart;
    if (ulen > MAX_UNCOMPRESSED_BLOCK) ulen = MAX_UNCOMPRESSED_BLOCK;
    int rc;
    if ((rc = add_bits(b, dst, (final_block && last) ? 1 : 0, 1))) return rc;
    if ((rc = add_bits(b, dst, 0, 2))) return rc;
    writer_skip_to_byte(b);
    if ((rc = add_bits(b, dst, (uint16_t)ulen, 16))) return rc;
    if ((rc = add_bits(b, dst, (uint16_t)(MAX_UNCOMPRESSED_BLOCK - ulen), 16))) return rc;
    if (ulen > 0)
      if ((rc = add_bytes(b, dst, src + ustart, ulen))) return rc;
  }
  return 0;
}

int zo_deflate(zo_buf *dst, const uint8_t *src, size_t len, int level) {
  ensure_tables();
  if (level < -2 || level > 9) return ZO_ERR_INVALID_LEVEL; /* :208-209 */
  bit_writer b = {dst->len, 0};                              /* :211-212 */
  int rc = 0;
  if (level == 0) { /* :214-226 */
    size_t count = (len + MAX_UNCOMPRESSED_BLOCK - 1) / MAX_UNCOMPRESSED_BLOCK;
    if (count < 1) count = 1;
    for (size_t k = 0; k < count; k++) {
      size_t bs = k * MAX_UNCOMPRESSED_BLOCK;
      size_t bl = len - bs < MAX_UNCOMPRESSED_BLOCK ? len - bs : MAX_UNCOMPRESSED_BLOCK;
      if ((rc = add_no_compression_block(&b, dst, src, bs, bl, k == count - 1))) return rc;
    }
    if (buf_reserve(dst, b.pos)) return ZO_ERR_NOMEM;
    dst->len = b.pos;
    return 0;
  }
  size_t block_count = (len + MAX_BLOCK_SIZE - 1) / MAX_BLOCK_SIZE; /* :228 */
  if (block_count < 1) block_count = 1;
  enc_vec enc = {NULL, 0, 0};
  for (size_t bn = 0; bn < block_count && rc == 0; bn++) {
    size_t block_start = bn * MAX_BLOCK_SIZE;
    size_t block_len = len - block_start < MAX_BLOCK_SIZE ? len - block_start : MAX_BLOCK_SIZE;
    int final_block = (bn == block_count - 1);
    enc.len = 0;
    block_meta meta;
    memset(&meta, 0, sizeof(meta));
    int erc;
    if (level == -2) /* :243-272 */
      erc = encode_all_literals(&enc, &meta, src, block_start, block_len);
    else if (level == 1)
      erc = encode_snappy(&enc, &meta, src, block_start, block_len);
    else
      erc = encode_lz77(&enc, config_table[level == -1 ? 6 : level], &meta, src, block_start, block_len);
    if (erc) {
      rc = ZO_ERR_NOMEM;
      break;
    }
    /* :274-277 stored fallback: float32 multiply, truncation */
    if (level != -2 && (long)meta.num_literals >= (long)((float)block_len * 0.98f)) {
      rc = add_no_compression_block(&b, dst, src, block_start, block_len, final_block);
      continue;
    }
    int use_fixed = (level <= 6 && block_len <= 2048); /* :280 */
    uint16_t ll_codes_buf[MAX_FIXED_LITLEN_CODES], d_codes_buf[MAX_DIST_CODES + 2];
    uint8_t ll_lens_buf[MAX_FIXED_LITLEN_CODES], d_lens_buf[MAX_DIST_CODES + 2];
    const uint16_t *ll_codes, *d_codes;
    const uint8_t *ll_lens, *d_lens;
    int n_ll, n_d;
    if (use_fixed) {
      ll_codes = fixed_litlen_codes;
      ll_lens = fixed_litlen_lengths;
      d_codes = fixed_dist_codes;
      d_lens = fixed_dist_lengths;
      n_ll = MAX_FIXED_LITLEN_CODES;
      n_d = MAX_DIST_CODES;
      if ((rc = add_bits(&b, dst, final_block ? 1 : 0, 1))) break; /* :292-294 */
      if ((rc = add_bits(&b, dst, 1, 2))) break;
    } else {
      n_ll = huffman_codes(meta.litlen_freq, MAX_LITLEN_CODES, 257, MAX_CODE_LENGTH, ll_codes_buf, ll_lens_buf);
      n_d = huffman_codes(meta.dist_freq, MAX_DIST_CODES, 2, MAX_CODE_LENGTH, d_codes_buf, d_lens_buf);
      ll_codes = ll_codes_buf;
      ll_lens = ll_lens_buf;
      d_codes = d_codes_buf;
      d_lens = d_lens_buf;
      /* :296-306 concatenated code lengths */
      uint8_t code_lengths[MAX_LITLEN_CODES + MAX_DIST_CODES];
      int num_codes = n_ll + n_d, cli = 0;
      for (int i = 0; i < n_ll; i++) code_lengths[cli++] = ll_lens[i];
      for (int i = 0; i < n_d; i++) code_lengths[cli++] = d_lens[i];
      /* :308-343 RLE */
      uint8_t rle[2 * (MAX_LITLEN_CODES + MAX_DIST_CODES) + 8];
      int nrle = 0;
      for (int i = 0; i < num_codes; i++) {
        int repeat = 0;
        while (i + repeat + 1 < num_codes && code_lengths[i + repeat + 1] == code_lengths[i]) repeat++;
        if (code_lengths[i] == 0 && repeat >= 2) {
          repeat++; /* initial zero */
          if (repeat <= 10) {
            rle[nrle++] = 17;
            rle[nrle++] = (uint8_t)(repeat - 3);
          } else {
            if (repeat > 138) repeat = 138;
            rle[nrle++] = 18;
            rle[nrle++] = (uint8_t)(repeat - 11);
          }
          i += repeat - 1;
        } else if (repeat >= 3) {
          int a = repeat / 6, bb = repeat % 6;
          rle[nrle++] = code_lengths[i];
          for (int j = 0; j < a; j++) {
            rle[nrle++] = 16;
            rle[nrle++] = 3;
          }
          if (bb >= 3) {
            rle[nrle++] = 16;
            rle[nrle++] = (uint8_t)(bb - 3);
          } else {
            repeat -= bb;
          }
          i += repeat;
        } else {
          rle[nrle++] = code_lengths[i];
        }
      }
      /* :345-353 */
      uint32_t cl_freq[19] = {0};
      for (int i = 0; i < nrle; i++) {
        cl_freq[rle[i]]++;
        if (rle[i] >= 16) i++;
      }
      uint16_t cl_codes[20];
      uint8_t cl_lens[20];
      huffman_codes(cl_freq, 19, 19, 7, cl_codes, cl_lens); /* :355 */
      uint16_t clcl_ordered[19];
      for (int i = 0; i < 19; i++) clcl_ordered[i] = cl_lens[clcl_order[i]];
      int hclen = 19; /* :361-364 */
      while (clcl_ordered[hclen - 1] == 0 && hclen > 4) hclen--;
      hclen -= 4;
      int hlit = n_ll - FIRST_LENGTH_CODE, hdist = n_d - 1;
      if ((rc = add_bits(&b, dst, final_block ? 1 : 0, 1))) break; /* :370-375 */
      if ((rc = add_bits(&b, dst, 2, 2))) break;
      if ((rc = add_bits(&b, dst, (uint32_t)hlit, 5))) break;
      if ((rc = add_bits(&b, dst, (uint32_t)hdist, 5))) break;
      if ((rc = add_bits(&b, dst, (uint32_t)hclen, 4))) break;
      for (int i = 0; i < hclen + 4 && rc == 0; i++) rc = add_bits(&b, dst, clcl_ordered[i], 3);
      for (int i = 0; i < nrle && rc == 0;) { /* :380-394 */
        uint8_t sym = rle[i];
        rc = add_bits(&b, dst, cl_codes[sym], cl_lens[sym]);
        i++;
        if (rc) break;
        if (sym == 16)
          rc = add_bits(&b, dst, rle[i++], 2);
        else if (sym == 17)
          rc = add_bits(&b, dst, rle[i++], 3);
        else if (sym == 18)
          rc = add_bits(&b, dst, rle[i++], 7);
      }
      if (rc) break;
    }
    (void)n_ll;
    (void)n_d;
    /* :396-459 token emission */
    size_t src_pos = block_start, enc_pos = 0;
    while (enc_pos < enc.len && rc == 0) {
      if (enc.v[enc_pos] & (1u << 15)) {
        uint16_t value = enc.v[enc_pos], offset = enc.v[enc_pos + 1], length = enc.v[enc_pos + 2];
        int length_index = (value >> 8) & 0x7f, dist_index = value & 0xff;
        int lebits = base_lengths_extra[length_index];
        uint64_t lextra = (uint64_t)(length - base_lengths[length_index]);
        int debits = base_distance_extra[dist_index];
        uint64_t dextra = (uint64_t)(offset - base_distances[dist_index]);
        enc_pos += 3;
        src_pos += length;
        uint64_t buf = ll_codes[length_index + 257];
        int bit_len = ll_lens[length_index + 257];
        buf |= lextra << bit_len;
        bit_len += lebits;
        buf |= (uint64_t)d_codes[dist_index] << bit_len;
        bit_len += d_lens[dist_index];
        buf |= dextra << bit_len;
        bit_len += debits;
        int first = bit_len < 32 ? bit_len : 32;
        rc = add_bits(&b, dst, (uint32_t)buf, first);
        buf >>= first;
        bit_len -= first;
        if (rc == 0 && bit_len > 0) rc = add_bits(&b, dst, (uint32_t)buf, bit_len);
      } else {
        size_t literals = enc.v[enc_pos++];
        uint32_t buf = 0;
        int bit_len = 0;
        for (size_t k = 0; k < literals && rc == 0; k++) {
          int cl = ll_lens[src[src_pos]];
          if (bit_len + cl > 32) {
            rc = add_bits(&b, dst, buf, bit_len);
            buf = 0;
            bit_len = 0;
          }
          buf |= (uint32_t)ll_codes[src[src_pos]] << bit_len;
          bit_len += cl;
          src_pos++;
        }
        if (rc == 0 && bit_len > 0) rc = add_bits(&b, dst, buf, bit_len);
      }
    }
    if (rc) break;
    if (enc_pos != enc.len) { /* :456-457 */
      rc = ZO_ERR_UNCOMPRESS;
      break;
    }
    if (ll_lens[256] == 0) { /* :461-462 */
      rc = ZO_ERR_COMPRESS;
      break;
    }
    rc = add_bits(&b, dst, ll_codes[256], ll_lens[256]); /* :464 */
  }
  free(enc.v);
  if (rc) return rc;
  writer_skip_to_byte(&b); /* :466-467 */
  if (buf_reserve(dst, b.pos)) return ZO_ERR_NOMEM;
  dst->len = b.pos;
  return 0;
}

/* ------------------------------------------------------------------ */
/* bit reader: bitstreams.nim:4-11, 22-82                              */
/* ------------------------------------------------------------------ */
typedef struct {
  const uint8_t *src;
  size_t len, pos;
  uint64_t bit_buffer;
  long bits_buffered;
} bit_reader;

/* bitstreams.nim:22-48 fillBitBuffer (64-bit variant); tail bytes past the
 * end read as zero (see header note). */
static inline void fill_bit_buffer(bit_reader *b) {
  if (b->bits_buffered < 0) return; /* already past the end; nothing can be added */
  size_t needed = (size_t)((64 - b->bits_buffered) / 8);
  size_t avail = b->len - b->pos;
  size_t added = needed < avail ? needed : avail;
  uint64_t s = 0;
  if (avail >= 8) {
    s = rd64(b->src + b->pos);
  } else {
    for (size_t i = 0; i < avail; i++) s |= (uint64_t)b->src[b->pos + i] << (8 * i);
  }
  b->pos += added;
  if (b->bits_buffered < 64) b->bit_buffer |= s << b->bits_buffered;
  b->bits_buffered += 8 * (long)added;
}

/* bitstreams.nim:50-62 readBits */
static inline uint16_t read_bits(bit_reader *b, int bits, int fill) {
  if (fill) fill_bit_buffer(b);
  uint16_t r = (uint16_t)(b->bit_buffer & (((uint32_t)1 << bits) - 1));
  b->bit_buffer >>= bits;
  b->bits_buffered -= bits;
  return r;
}

/* ------------------------------------------------------------------ */
/* inflate: inflate.nim                                                */
/* ------------------------------------------------------------------ */
#define FAST_BITS 9
#define FAST_MASK ((1 << FAST_BITS) - 1)

typedef struct { /* inflate.nim:14-19 */
  uint16_t first_code[16], first_symbol[16];
  uint32_t max_codes[17];
  uint16_t values[288];
  uint16_t fast[1 << FAST_BITS];
} huffman;

/* inflate.nim:24-65 initHuffman */
static int init_huffman(huffman *h, const uint8_t *code_lengths, int n) {
  memset(h, 0, sizeof(*h));
  uint16_t histogram[17] = {0};
  for (int i = 0; i < n; i++) histogram[code_lengths[i]]++;
  histogram[0] = 0;
  for (int i = 1; i < 16; i++)
    if (histogram[i] > (1u << i)) return ZO_ERR_UNCOMPRESS; /* :32-34 */
  uint32_t code = 0, next_code[16] = {0};
  uint16_t k = 0;
  for (int i = 1; i < 16; i++) { /* :40-49 */
    next_code[i] = code;
    h->first_code[i] = (uint16_t)code;
    h->first_symbol[i] = k;
    code += histogram[i];
    if (histogram[i] > 0 && code - 1 >= (1u << i)) return ZO_ERR_UNCOMPRESS;
    h->max_codes[i] = code << (16 - i);
    code <<= 1;
    k = (uint16_t)(k + histogram[i]);
  }
  h->max_codes[16] = 1u << 16;
  for (int i = 0; i < n; i++) { /* :53-65 */
    uint8_t len = code_lengths[i];
    if (len > 0) {
      uint32_t symbol_id = next_code[len] - h->first_code[len] + h->first_symbol[len];
      h->values[symbol_id] = (uint16_t)i;
      if (len <= FAST_BITS) {
        uint16_t fast = (uint16_t)((len << FAST_BITS) | i);
        uint32_t kk = (uint32_t)(reverse_bits16((uint16_t)next_code[len]) >> (16 - len));
        while (kk < (1u << FAST_BITS)) {
          h->fast[kk] = fast;
          kk += 1u << len;
        }
      }
      next_code[len]++;
    }
  }
  return 0;
}

/* inflate.nim:93-102 decodeSymbol + :67-91 decodeSymbolSlow */
static inline uint16_t decode_symbol(bit_reader *b, const huffman *h) {
  uint16_t fast = h->fast[b->bit_buffer & FAST_MASK];
  if (fast > 0) {
    int cl = fast >> FAST_BITS;
    b->bit_buffer >>= cl;
    b->bits_buffered -= cl;
    return fast & FAST_MASK;
  }
  uint16_t k = reverse_bits16((uint16_t)b->bit_buffer);
  int cl = FAST_BITS + 1;
  while (cl < 17) {
    if ((uint32_t)k < h->max_codes[cl]) break;
    cl++;
  }
  if (cl >= 16) return 0xffff; /* :77-82 */
  uint32_t symbol_id = (uint32_t)(k >> (16 - cl)) - h->first_code[cl] + h->first_symbol[cl];
  b->bit_buffer >>= cl;
  b->bits_buffered -= cl;
  if (symbol_id >= 288) return 0xffff; /* unreachable for canonical codes; guards the array */
  return h->values[symbol_id];
}

/* inflate.nim:104-250 inflateBlock */
static int inflate_block(zo_buf *dst, bit_reader *b, size_t *op_io, int fixed_codes) {
  huffman lit, dist;
  int rc;
  size_t op = *op_io;
  if (fixed_codes) { /* :111-113 */
    if ((rc = init_huffman(&lit, fixed_litlen_lengths, MAX_FIXED_LITLEN_CODES))) return rc;
    if ((rc = init_huffman(&dist, fixed_dist_lengths, MAX_DIST_CODES))) return rc;
  } else {
    int hlit = read_bits(b, 5, 1) + 257; /* :116-118 */
    int hdist = read_bits(b, 5, 1) + 1;
    int hclen = read_bits(b, 4, 1) + 4;
    if (hlit > MAX_LITLEN_CODES) return ZO_ERR_UNCOMPRESS; /* :120-124 */
    if (hdist > MAX_DIST_CODES) return ZO_ERR_UNCOMPRESS;
    uint8_t clcls[19] = {0};
    for (int i = 0; i < hclen; i++) clcls[clcl_order[i]] = (uint8_t)read_bits(b, 3, 1);
    huffman clh;
    if ((rc = init_huffman(&clh, clcls, 19))) return rc;
    uint8_t unpacked[320]; /* :135-168 */
    memset(unpacked, 0, sizeof(unpacked));
    int i = 0;
    while (i != hlit + hdist) {
      if (b->bits_buffered < 15) fill_bit_buffer(b);
      uint16_t symbol = decode_symbol(b, &clh);
      if (b->bits_buffered < 0) return ZO_ERR_END_OF_BUFFER;
      if (symbol <= 15) {
        unpacked[i++] = (uint8_t)symbol;
      } else if (symbol == 16) {
        if (i == 0) return ZO_ERR_UNCOMPRESS;
        uint8_t prev = unpacked[i - 1];
        int repeat = read_bits(b, 2, 1) + 3;
        if (i + repeat > 320) return ZO_ERR_UNCOMPRESS;
        for (int r = 0; r < repeat; r++) unpacked[i++] = prev;
      } else if (symbol == 17) {
        i += read_bits(b, 3, 1) + 3;
      } else if (symbol == 18) {
        i += read_bits(b, 7, 1) + 11;
      } else {
        return ZO_ERR_INVALID_SYMBOL; /* :165 */
      }
      if (i > hlit + hdist) return ZO_ERR_UNCOMPRESS; /* :167-168 */
    }
    if ((rc = init_huffman(&lit, unpacked, hlit))) return rc; /* :170-171 */
    if ((rc = init_huffman(&dist, unpacked + hlit, hdist))) return rc;
  }
  for (;;) { /* :173-250 */
    if (op + 320 > dst->cap && buf_reserve(dst, op + 320)) return ZO_ERR_NOMEM; /* :192-196, :227-229 */
    if (b->bits_buffered < 15) fill_bit_buffer(b);
    uint16_t symbol = decode_symbol(b, &lit);
    if (b->bits_buffered < 0) return ZO_ERR_END_OF_BUFFER; /* :190-191 */
    if (symbol <= 255) {
      dst->data[op++] = (uint8_t)symbol;
    } else if (symbol == 256) {
      break;
    } else {
      fill_bit_buffer(b);
      int length_idx = symbol - 257;
      if (length_idx >= 29) return ZO_ERR_UNCOMPRESS; /* :203-204 */
      size_t copy_len = (size_t)base_lengths[length_idx] + read_bits(b, base_lengths_extra[length_idx], 0);
      uint16_t distance_idx = decode_symbol(b, &dist);
      if (distance_idx >= 30) return ZO_ERR_UNCOMPRESS; /* :212-213 */
      size_t distance = (size_t)base_distances[distance_idx] + read_bits(b, base_distance_extra[distance_idx], 0);
      if (b->bits_buffered < 0) return ZO_ERR_END_OF_BUFFER; /* reference: caught at the next :190 */
      if (distance > op) return ZO_ERR_UNCOMPRESS;             /* :224-225 */
      uint8_t *d = dst->data;
      if (distance >= copy_len) {
        memcpy(d + op, d + op - distance, copy_len);
      } else {
        for (size_t k = 0; k < copy_len; k++) d[op + k] = d[op + k - distance]; /* :233-250, bytewise form */
      }
      op += copy_len;
    }
  }
  *op_io = op;
  return 0;
}

/* inflate.nim:252-266 inflateNoCompression */
static int inflate_no_compression(zo_buf *dst, bit_reader *b, size_t *op_io) {
  if (b->bits_buffered < 0) return ZO_ERR_END_OF_BUFFER;
  long mod8 = b->bits_buffered % 8; /* bitstreams.nim:78-82 */
  if (mod8 != 0) {
    b->bits_buffered -= mod8;
    b->bit_buffer >>= mod8;
  }
  uint32_t len = read_bits(b, 16, 1);
  uint32_t nlen = read_bits(b, 16, 1);
  if (b->bits_buffered < 0) return ZO_ERR_END_OF_BUFFER;
  if (len + nlen != 65535) return ZO_ERR_UNCOMPRESS; /* :261-262 */
  if (len > 0) {
    /* bitstreams.nim:64-76 readBytes */
    if (b->bits_buffered % 8 != 0) return ZO_ERR_BYTE_BOUNDARY;
    size_t offset = (size_t)(b->bits_buffered / 8);
    if (b->pos - offset + len > b->len) return ZO_ERR_END_OF_BUFFER;
    if (buf_reserve(dst, *op_io + len)) return ZO_ERR_NOMEM;
    memcpy(dst->data + *op_io, b->src + b->pos - offset, len);
    b->pos = b->pos - offset + len;
    b->bits_buffered = 0;
    b->bit_buffer = 0;
  }
  *op_io += len;
  return 0;
}

/* inflate.nim:268-291 */
int zo_inflate(zo_buf *dst, const uint8_t *src, size_t len, size_t pos) {
  ensure_tables();
  bit_reader b = {src, len, pos, 0, 0};
  size_t op = 0;
  int final_block = 0, rc = 0;
  dst->len = 0;
  if (pos > len) return ZO_ERR_END_OF_BUFFER;
  while (!final_block) {
    uint16_t bfinal = read_bits(&b, 1, 1);
    uint16_t btype = read_bits(&b, 2, 1);
    if (b.bits_buffered < 0) return ZO_ERR_END_OF_BUFFER;
    if (bfinal) final_block = 1;
    switch (btype) {
      case 0:
        rc = inflate_no_compression(dst, &b, &op);
        break;
      case 1:
        rc = inflate_block(dst, &b, &op, 1);
        break;
      case 2:
        rc = inflate_block(dst, &b, &op, 0);
        break;
      default:
        rc = ZO_ERR_BLOCK_HEADER; /* :289 */
    }
    if (rc) return rc;
  }
  dst->len = op; /* :291 */
  return 0;
}

/* ------------------------------------------------------------------ */
/* framing: zippy.nim:11-84 compress                                   */
/* ------------------------------------------------------------------ */
static int buf_append(zo_buf *b, const uint8_t *p, size_t n) {
  if (buf_reserve(b, b->len + n)) return -1;
  memcpy(b->data + b->len, p, n);
  b->len += n;
  return 0;
}

int zo_compress(zo_buf *dst, const uint8_t *src, size_t len, int level, int data_format,
                int fname_len) {
  dst->len = 0;
  if (dst->cap) memset(dst->data, 0, dst->cap);
  int rc;
  if (data_format == ZO_DF_GZIP) { /* :21-58 */
    uint8_t hdr[10] = {31, 139, 8, 1u << 3, 0, 0, 0, 0, 0, 0};
    if (buf_append(dst, hdr, 10)) return ZO_ERR_NOMEM;
    int k = fname_len >= 0 ? fname_len % 26 : rand() % 26; /* :28-42 */
    for (int i = 0; i < k; i++) {
      uint8_t c = (uint8_t)(97 + i);
      if (buf_append(dst, &c, 1)) return ZO_ERR_NOMEM;
    }
    uint8_t z = 0;
    if (buf_append(dst, &z, 1)) return ZO_ERR_NOMEM;
    if ((rc = zo_deflate(dst, src, len, level))) return rc;
    uint32_t crc = zo_crc32(src, len);
    uint32_t isize = (uint32_t)len;
    uint8_t tr[8] = {(uint8_t)crc,   (uint8_t)(crc >> 8),   (uint8_t)(crc >> 16),   (uint8_t)(crc >> 24),
                     (uint8_t)isize, (uint8_t)(isize >> 8), (uint8_t)(isize >> 16), (uint8_t)(isize >> 24)};
    if (buf_append(dst, tr, 8)) return ZO_ERR_NOMEM;
    return 0;
  }
  if (data_format == ZO_DF_ZLIB) { /* :60-78 */
    const uint8_t cmf = (7u << 4) | 8u;
    const uint8_t fcheck = (uint8_t)(31u - ((uint32_t)cmf * 256u) % 31u);
    uint8_t hdr[2] = {cmf, fcheck};
    if (buf_append(dst, hdr, 2)) return ZO_ERR_NOMEM;
    if ((rc = zo_deflate(dst, src, len, level))) return rc;
    uint32_t a = zo_adler32(src, len);
    uint8_t tr[4] = {(uint8_t)(a >> 24), (uint8_t)(a >> 16), (uint8_t)(a >> 8), (uint8_t)a};
    if (buf_append(dst, tr, 4)) return ZO_ERR_NOMEM;
    return 0;
  }
  if (data_format == ZO_DF_DEFLATE) return zo_deflate(dst, src, len, level); /* :80-81 */
  return ZO_ERR_INVALID_FORMAT;                                                /* :83-84 */
}

/* gzip.nim:3-88 uncompressGzip (trustSize only pre-sizes; no effect on results) */
static int uncompress_gzip(zo_buf *dst, const uint8_t *src, size_t len) {
  if (len < 18) return ZO_ERR_UNCOMPRESS;
  uint8_t flg = src[3];
  if (src[0] != 31 || src[1] != 139) return ZO_ERR_GZIP_ID;
  if (src[2] != 8) return ZO_ERR_METHOD;
  if (flg & 0xe0) return ZO_ERR_GZIP_RESERVED;
  int fhcrc = flg & 2, fextra = flg & 4, fname = flg & 8, fcomment = flg & 16;
  size_t pos = 10;
  if (fextra) return ZO_ERR_GZIP_FLAGS;
  for (int pass = 0; pass < 2; pass++) { /* :45-53 nextZeroByte for FNAME then FCOMMENT */
    if ((pass == 0 && fname) || (pass == 1 && fcomment)) {
      size_t i = pos;
      while (i < len && src[i] != 0) i++;
      if (i >= len) return ZO_ERR_UNCOMPRESS;
      pos = i + 1;
    }
  }
  if (fhcrc) { /* :55-59 */
    if (pos + 2 >= len) return ZO_ERR_UNCOMPRESS;
    pos += 2;
  }
  if (pos + 8 >= len) return ZO_ERR_UNCOMPRESS; /* :61-62 */
  uint32_t checksum = rd32(src + len - 8), isize = rd32(src + len - 4);
  int rc = zo_inflate(dst, src, len, pos);
  if (rc) return rc;
  if (checksum != zo_crc32(dst->data, dst->len)) return ZO_ERR_CHECKSUM;
  if (isize != (uint32_t)(dst->len & 0xffffffffu)) return ZO_ERR_SIZE;
  return 0;
}

/* zippy.nim:100-165 */
int zo_uncompress(zo_buf *dst, const uint8_t *src, size_t len, int data_format) {
  dst->len = 0;
  switch (data_format) {
    case ZO_DF_DETECT:
      if (len > 18 && src[0] == 31 && src[1] == 139 && src[2] == 8 && (src[3] & 0xe0) == 0)
        return zo_uncompress(dst, src, len, ZO_DF_GZIP);
      if (len > 6 && (src[0] & 0x0f) == 8 && (src[0] >> 4) <= 7 &&
          (((uint32_t)src[0] * 256u) + src[1]) % 31u == 0)
        return zo_uncompress(dst, src, len, ZO_DF_ZLIB);
      return ZO_ERR_DETECT;
    case ZO_DF_GZIP:
      return uncompress_gzip(dst, src, len);
    case ZO_DF_ZLIB: {
      if (len < 6) return ZO_ERR_UNCOMPRESS;
      uint8_t cmf = src[0], flg = src[1];
      if ((cmf & 0x0f) != 8) return ZO_ERR_METHOD;
      if ((cmf >> 4) > 7) return ZO_ERR_CINFO;
      if ((((uint32_t)cmf * 256u) + flg) % 31u != 0) return ZO_ERR_HEADER;
      if (flg & 0x20) return ZO_ERR_FDICT;
      int rc = zo_inflate(dst, src, len, 2);
      if (rc) return rc;
      uint32_t checksum = ((uint32_t)src[len - 4] << 24) | ((uint32_t)src[len - 3] << 16) |
                          ((uint32_t)src[len - 2] << 8) | src[len - 1];
      if (checksum != zo_adler32(dst->data, dst->len)) return ZO_ERR_CHECKSUM;
      return 0;
    }
    case ZO_DF_DEFLATE:
      return zo_inflate(dst, src, len, 0);
    default:
      return ZO_ERR_INVALID_FORMAT;
  }
}

/* ------------------------------------------------------------------ */
/* batch helpers for the CPU baseline (not in the reference)           */
/* ------------------------------------------------------------------ */
typedef struct {
  const uint8_t *base;
  const uint64_t *offsets;
  size_t n;
  int level, data_format, compress;
  uint64_t *out_lens;
  int *statuses;
  size_t *next;
  pthread_mutex_t *mu;
  uint64_t total;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  zo_buf out = {NULL, 0, 0};
  for (;;) {
    pthread_mutex_lock(j->mu);
    size_t i = *j->next;
    size_t hi = i + 16 < j->n ? i + 16 : j->n;
    *j->next = hi;
    pthread_mutex_unlock(j->mu);
    if (i >= j->n) break;
    for (; i < hi; i++) {
      const uint8_t *p = j->base + j->offsets[i];
      size_t l = (size_t)(j->offsets[i + 1] - j->offsets[i]);
      int rc = j->compress ? zo_compress(&out, p, l, j->level, j->data_format, 0)
                           : zo_uncompress(&out, p, l, j->data_format);
      if (j->statuses) j->statuses[i] = rc;
      if (j->out_lens) j->out_lens[i] = rc ? 0 : out.len;
      if (!rc) j->total += out.len;
    }
  }
  zo_buf_free(&out);
  return NULL;
}

static uint64_t run_batch(const uint8_t *base, const uint64_t *offsets, size_t n, int level,
                          int data_format, int threads, int compress, uint64_t *out_lens,
                          int *statuses) {
  ensure_tables();
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  size_t next = 0;
  batch_job jobs[256];
  pthread_t tids[256];
  for (int t = 0; t < threads; t++) {
    jobs[t] = (batch_job){base, offsets, n, level, data_format, compress, out_lens, statuses, &next, &mu, 0};
    if (threads > 1) pthread_create(&tids[t], NULL, batch_worker, &jobs[t]);
  }
  if (threads == 1) batch_worker(&jobs[0]);
  uint64_t total = 0;
  for (int t = 0; t < threads; t++) {
    if (threads > 1) pthread_join(tids[t], NULL);
    total += jobs[t].total;
  }
  return total;
}

uint64_t zo_compress_batch(const uint8_t *base, const uint64_t *offsets, size_t n, int level,
                           int data_format, int threads, uint64_t *out_lens, int *statuses) {
  return run_batch(base, offsets, n, level, data_format, threads, 1, out_lens, statuses);
}
uint64_t zo_uncompress_batch(const uint8_t *base, const uint64_t *offsets, size_t n,
                             int data_format, int threads, uint64_t *out_lens, int *statuses) {
  return run_batch(base, offsets, n, 0, data_format, threads, 0, out_lens, statuses);
}

const char *zo_strerror(int code) {
  switch (code) {
    case ZO_OK: return "ok";
    case ZO_ERR_INVALID_LEVEL: return "Invalid compression level";
    case ZO_ERR_INVALID_FORMAT: return "Invalid data format";
    case ZO_ERR_UNCOMPRESS: return "Invalid buffer, unable to uncompress";
    case ZO_ERR_COMPRESS: return "Unexpected error while compressing";
    case ZO_ERR_END_OF_BUFFER: return "Cannot read further, at end of buffer";
    case ZO_ERR_BYTE_BOUNDARY: return "Must be at a byte boundary";
    case ZO_ERR_BLOCK_HEADER: return "Invalid block header";
    case ZO_ERR_INVALID_SYMBOL: return "Invalid symbol";
    case ZO_ERR_DETECT: return "Unable to detect compressed data format";
    case ZO_ERR_METHOD: return "Unsupported compression method";
    case ZO_ERR_CINFO: return "Invalid compression info";
    case ZO_ERR_HEADER: return "Invalid header";
    case ZO_ERR_FDICT: return "Preset dictionary is not yet supported";
    case ZO_ERR_CHECKSUM: return "Checksum verification failed";
    case ZO_ERR_GZIP_ID: return "Failed gzip identification values check";
    case ZO_ERR_GZIP_RESERVED: return "Reserved flag bits set";
    case ZO_ERR_GZIP_FLAGS: return "Currently unsupported flags are set";
    case ZO_ERR_SIZE: return "Size verification failed";
    case ZO_ERR_NOMEM: return "out of memory";
    default: return "unknown";
  }
}
