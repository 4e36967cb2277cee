/*
 * ORACLE (test infrastructure only -- never linked into the product path).
 *
 * Plain-C restatement of the range coder the reference calls through
 * `compressai.ans` (CompressAI 1.2.4, un-vendored dependency pinned in
 * /root/reference/pyproject.toml:16; call sites:
 * src/models/subnet/context_model/minnen20_charm_context_model.py:186-187,
 * 201-202,222-224 and compressai EntropyModel.compress/decompress).
 * Algorithm: ryg_rans `rans64.h` (64-bit state, L = 2^31, 32-bit renorm
 * words) + CompressAI's `rans_interface.cpp` bypass coding (precision 16,
 * bypass_precision 4) and `pmf_to_quantized_cdf`.
 * PARITY UNPINNED: CompressAI is absent from this container, so this follows
 * the published algorithm from memory of that release; see DESIGN.md.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION 16
#define BYPASS_BITS 4
#define BYPASS_MAX ((1 << BYPASS_BITS) - 1)
#define RANS_L (1ull << 31)

typedef struct { uint16_t start, range; uint8_t bypass; } sym_t;

typedef struct { sym_t *v; size_t n, cap; } symvec_t;

static void push(symvec_t *s, uint16_t start, uint16_t range, int bypass) {
  if (s->n == s->cap) {
    s->cap = s->cap ? s->cap * 2 : 1024;
    s->v = (sym_t *)realloc(s->v, s->cap * sizeof(sym_t));
  }
  s->v[s->n].start = start; s->v[s->n].range = range; s->v[s->n].bypass = (uint8_t)bypass;
  s->n++;
}

/* cdfs: [n_cdf][cdf_stride] int32.  Returns number of bytes written to out
 * (out must hold 4*(n_queued+2) bytes; pass out=NULL to query the bound). */
long oracle_rans_encode(const int32_t *symbols, const int32_t *indexes, long n,
                        const int32_t *cdfs, long cdf_stride, const int32_t *cdf_sizes,
                        const int32_t *offsets, uint8_t *out, long out_cap) {
  symvec_t q = {0, 0, 0};
  for (long i = 0; i < n; ++i) {
    int32_t ci = indexes[i];
    const int32_t *cdf = cdfs + (long)ci * cdf_stride;
    int32_t max_value = cdf_sizes[ci] - 2;
    int32_t value = symbols[i] - offsets[ci];
    uint32_t raw = 0;
    if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
    else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
    push(&q, (uint16_t)cdf[value], (uint16_t)(cdf[value + 1] - cdf[value]), 0);
    if (value == max_value) {
      int32_t nb = 0;
      while ((raw >> (nb * BYPASS_BITS)) != 0) ++nb;
      int32_t val = nb;
      while (val >= BYPASS_MAX) { push(&q, BYPASS_MAX, BYPASS_MAX + 1, 1); val -= BYPASS_MAX; }
      push(&q, (uint16_t)val, (uint16_t)(val + 1), 1);
      for (int32_t j = 0; j < nb; ++j) {
        int32_t v = (raw >> (j * BYPASS_BITS)) & BYPASS_MAX;
        push(&q, (uint16_t)v, (uint16_t)(v + 1), 1);
      }
    }
  }
  size_t words = q.n + 2;
  uint32_t *buf = (uint32_t *)malloc(words * sizeof(uint32_t));
  uint32_t *ptr = buf + words;
  uint64_t x = RANS_L;
  for (size_t k = q.n; k-- > 0;) {
    sym_t s = q.v[k];
    if (!s.bypass) {
      uint64_t x_max = ((RANS_L >> PRECISION) << 32) * (uint64_t)s.range;
      if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
      x = ((x / s.range) << PRECISION) + (x % s.range) + s.start;
    } else {
      uint32_t freq = 1u << (16 - BYPASS_BITS);
      uint64_t x_max = ((RANS_L >> 16) << 32) * (uint64_t)freq;
      if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
      x = (x << BYPASS_BITS) | s.start;
    }
  }
  ptr -= 2;
  ptr[0] = (uint32_t)x; ptr[1] = (uint32_t)(x >> 32);
  long nbytes = (long)((buf + words) - ptr) * 4;
  long ret = nbytes;
  if (out) { if (nbytes <= out_cap) memcpy(out, ptr, (size_t)nbytes); else ret = -nbytes; }
  free(buf); free(q.v);
  return ret;
}

typedef struct { uint64_t x; const uint32_t *ptr; uint32_t *own; } oracle_dec_t;

void *oracle_rans_dec_new(const uint8_t *stream, long nbytes) {
  oracle_dec_t *d = (oracle_dec_t *)malloc(sizeof(oracle_dec_t));
  d->own = (uint32_t *)malloc((size_t)nbytes + 16);
  memset(d->own, 0, (size_t)nbytes + 16);
  memcpy(d->own, stream, (size_t)nbytes);
  d->ptr = d->own;
  d->x = (uint64_t)d->ptr[0] | ((uint64_t)d->ptr[1] << 32);
  d->ptr += 2;
  return d;
}
void oracle_rans_dec_free(void *h) { oracle_dec_t *d = (oracle_dec_t *)h; free(d->own); free(d); }

static inline uint32_t get_bits(oracle_dec_t *d, uint32_t nb) {
  uint64_t x = d->x;
  uint32_t val = (uint32_t)(x & ((1u << nb) - 1));
  x >>= nb;
  if (x < RANS_L) { x = (x << 32) | *d->ptr++; }
  d->x = x;
  return val;
}

void oracle_rans_dec_stream(void *h, const int32_t *indexes, long n, const int32_t *cdfs,
                            long cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                            int32_t *out) {
  oracle_dec_t *d = (oracle_dec_t *)h;
  for (long i = 0; i < n; ++i) {
    int32_t ci = indexes[i];
    const int32_t *cdf = cdfs + (long)ci * cdf_stride;
    int32_t max_value = cdf_sizes[ci] - 2;
    uint32_t cum = (uint32_t)(d->x & ((1u << PRECISION) - 1));
    int32_t s = 0;
    int32_t len = cdf_sizes[ci];
    while (s < len && !((uint32_t)cdf[s] > cum)) ++s;
    s -= 1;
    uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
    uint64_t x = d->x;
    x = freq * (x >> PRECISION) + (x & ((1ull << PRECISION) - 1)) - start;
    if (x < RANS_L) { x = (x << 32) | *d->ptr++; }
    d->x = x;
    int32_t value = s;
    if (value == max_value) {
      int32_t val = (int32_t)get_bits(d, BYPASS_BITS);
      int32_t nb = val;
      while (val == BYPASS_MAX) { val = (int32_t)get_bits(d, BYPASS_BITS); nb += val; }
      int32_t raw = 0;
      for (int32_t j = 0; j < nb; ++j) { val = (int32_t)get_bits(d, BYPASS_BITS); raw |= val << (j * BYPASS_BITS); }
      value = raw >> 1;
      if (raw & 1) value = -value - 1; else value += max_value;
    }
    out[i] = value + offsets[ci];
  }
}

/* pmf (float32, n entries) -> cdf (uint32, n+1 entries). returns 0 ok. */
int oracle_pmf_to_quantized_cdf(const float *pmf, long n, int precision, uint32_t *cdf) {
  for (long i = 0; i < n; ++i) if (pmf[i] < 0 || !isfinite(pmf[i])) return 1;
  cdf[0] = 0;
  for (long i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
  uint32_t total = 0;
  for (long i = 0; i <= n; ++i) total += cdf[i];
  if (total == 0) return 2;
  for (long i = 0; i <= n; ++i) cdf[i] = (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
  for (long i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
  cdf[n] = 1u << precision;
  for (long i = 0; i < n; ++i) {
    if (cdf[i] == cdf[i + 1]) {
      uint32_t best_freq = ~0u; long best = -1;
      for (long j = 0; j < n; ++j) {
        uint32_t f = cdf[j + 1] - cdf[j];
        if (f > 1 && f < best_freq) { best_freq = f; best = j; }
      }
      if (best < 0) return 3;
      if (best < i) { for (long j = best + 1; j <= i; ++j) cdf[j]--; }
      else { for (long j = i + 1; j <= best; ++j) cdf[j]++; }
    }
  }
  return 0;
}
