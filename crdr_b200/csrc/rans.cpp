// Host range coder for the CRDR bitstream (see include/crdr_rans.h).
//
// Streaming formulation: the encoder walks the symbols back to front and feeds each symbol's
// sub-symbols (escape nibbles, escape-length run, table symbol) straight into the rANS state, so no
// per-stream symbol queue is materialised; emitted 32-bit words are collected and reversed once.
#include "../../include/crdr_rans.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>
#ifdef __linux__
#include <pthread.h>
#include <sched.h>
#endif

namespace {

constexpr uint32_t kPrecision = 16;
constexpr uint32_t kBypassBits = 4;
constexpr int32_t kBypassMax = (1 << kBypassBits) - 1;
constexpr uint64_t kRansL = 1ull << 31;

// Per-table acceleration structures, built once per (table pointer, geometry) and shared by all coder threads:
//  * encoder: per (row, value) reciprocal of the frequency, so the state update is a multiply-high instead of a 64-bit
//    division (Alverson; exact for states below 2^63, which the renormalisation bound guarantees) - same words out;
//  * decoder: per row a 1024-bucket index over the 16-bit cumulative range.  A bucket that lies inside ONE slot's
//    interval (most of the probability mass of a peaked row) carries that slot's (start, freq) in the entry itself:
//    one load replaces bucket -> 8-wide compare -> two row loads on the decoder's serial dependency chain.  Other
//    buckets give the first candidate slot and the search walks forward from there (Gaussian / logistic tables put a
//    handful of slots in a bucket).
struct EncSym {
  uint64_t rcp_freq;
  uint32_t bias;
  uint32_t cmpl_freq;   // 2^16 - freq
  uint32_t rcp_shift;
  uint32_t freq;
};
constexpr int kBucketBits = 10;                       // decoder: 1024 buckets of 64 cumulative values per table row
constexpr int kBuckets = 1 << kBucketBits;
constexpr int kBucketShift = 16 - kBucketBits;
struct TableAccel {
  std::vector<EncSym> enc;        // [n_cdf][cdf_stride] (entries past a row's size are unused)
  std::vector<uint64_t> bucket;   // [n_cdf][kBuckets]: start | freq << 16 | first slot << 32 | pure << 48 (see above)
  std::vector<int32_t> cdf_pad;   // [n_cdf][pad_stride]: the rows followed by 8 sentinels (> any cumulative value), so the
  int32_t pad_stride = 0;         //   decoder's 8-wide compare may read past a row's end
};
// Built by crdr_rans_tables_prepare() and owned by the caller's handle (crdr_cdf_tables::prepared).  There is no
// cache keyed on table addresses: a rebuilt or reloaded table that lands on a recycled address must never pick up
// another table's frequencies (the encoder would then disagree with the decoder without any error).
std::shared_ptr<const TableAccel> build_accel(const crdr_cdf_tables* t) {
  auto a = std::make_shared<TableAccel>();
  a->enc.assign((size_t)t->n_cdf * (size_t)t->cdf_stride, EncSym{0, 0, 0, 0, 0});
  a->bucket.assign((size_t)t->n_cdf * kBuckets, 0);
  a->pad_stride = t->cdf_stride + 8;
  a->cdf_pad.assign((size_t)t->n_cdf * (size_t)a->pad_stride, 0x7fffffff);
  for (int32_t r = 0; r < t->n_cdf; ++r) {
    const int32_t* cdf = t->cdfs + (int64_t)r * t->cdf_stride;
    const int32_t size = std::min(t->cdf_sizes[r], t->cdf_stride);
    for (int32_t v = 0; v + 1 < size; ++v) {
      const uint32_t start = (uint32_t)cdf[v], freq = (uint32_t)(cdf[v + 1] - cdf[v]);
      EncSym& e = a->enc[(size_t)r * t->cdf_stride + v];
      e.freq = freq;
      e.cmpl_freq = (1u << kPrecision) - freq;
      if (freq < 2) {
        e.rcp_freq = ~0ull; e.rcp_shift = 0; e.bias = start + (1u << kPrecision) - 1;
      } else {
        uint32_t shift = 0;
        while (freq > (1u << shift)) ++shift;
        e.rcp_freq = (uint64_t)(((((unsigned __int128)1) << (shift + 63)) + freq - 1) / freq);
        e.rcp_shift = shift - 1;
        e.bias = start;
      }
    }
    uint64_t* bk = a->bucket.data() + (size_t)r * kBuckets;
    int32_t sidx = 0;
    for (int32_t b = 0; b < kBuckets; ++b) {
      const int32_t cum = b << kBucketShift;
      while (sidx + 2 < size && cdf[sidx + 1] <= cum) ++sidx;
      uint64_t e = (uint64_t)(uint32_t)sidx << 32;
      if (sidx + 1 < size && cdf[sidx] <= cum && cdf[sidx + 1] > cum + (1 << kBucketShift) - 1) {
        const uint32_t start = (uint32_t)cdf[sidx], freq = (uint32_t)(cdf[sidx + 1] - cdf[sidx]);
        if (start < 65536u && freq < 65536u) e |= (uint64_t)start | ((uint64_t)freq << 16) | (1ull << 48);
      }
      bk[b] = e;
    }
    // padded copy: entries [0, size-1] are the row (the last one is 2^16 > any cumulative value), sentinels behind it
    int32_t* cp = a->cdf_pad.data() + (size_t)r * (size_t)a->pad_stride;
    for (int32_t v = 0; v < size; ++v) cp[v] = cdf[v];
  }
  return a;
}

// A prepared table set: private copies of the arrays + the acceleration structures.
struct Prepared {
  std::vector<int32_t> cdfs, sizes, offsets;
  crdr_cdf_tables view;
  std::shared_ptr<const TableAccel> accel;
};

// The tables a coder call works on: the prepared copy when the caller made one, otherwise the caller's arrays with
// acceleration structures built for this call only (correct, just slower).
struct TableRef {
  const crdr_cdf_tables* t;
  std::shared_ptr<const TableAccel> accel;
  explicit TableRef(const crdr_cdf_tables* in) {
    if (in->prepared) {
      const Prepared* p = static_cast<const Prepared*>(in->prepared);
      t = &p->view;
      accel = p->accel;
    } else {
      t = in;
      accel = build_accel(in);
    }
  }
};

struct Encoder {
  uint64_t x = kRansL;
  uint32_t* wp = nullptr;  // next free word, emission order (the stream stores the words reversed)

  // Renormalisation.  JUMP (one stream per thread): a jump -- at the codec's rates (about a bit per symbol) a word is due
  // every few dozen symbols, the jump predicts well and compare + shift stay off the serial chain x -> x' (multiply-high,
  // shift, multiply, add).  !JUMP (interleaved streams): branch free -- the low word is always stored, the pointer only
  // advances when it was due, the state shifts by 0 or 32 (a ?: compiles to a jump).
  template <bool JUMP>
  __attribute__((always_inline)) inline void renorm(uint64_t x_max) {
    if (JUMP) {
      if (__builtin_expect(x >= x_max, 0)) { *wp++ = (uint32_t)x; x >>= 32; }
    } else {
      const uint64_t emit = (uint64_t)(x >= x_max);
      *wp = (uint32_t)x;
      wp += emit;
      x >>= (emit << 5);
    }
  }
  template <bool JUMP>
  __attribute__((always_inline)) inline void put(const EncSym& e) {
    renorm<JUMP>(((kRansL >> kPrecision) << 32) * (uint64_t)e.freq);
    const uint64_t q = (uint64_t)(((unsigned __int128)x * e.rcp_freq) >> 64) >> e.rcp_shift;   // == x / freq
    x = x + e.bias + q * e.cmpl_freq;
  }
  template <bool JUMP>
  __attribute__((always_inline)) inline void put_bits(uint32_t val) {
    renorm<JUMP>(((kRansL >> 16) << 32) * (uint64_t)(1u << (16 - kBypassBits)));
    x = (x << kBypassBits) | val;
  }
};

struct EncTables {
  const EncSym* esym;
  const int32_t* offsets;
  const int32_t* sizes;
  int64_t stride;
  int32_t n_cdf;
  explicit EncTables(const TableRef& tr)
      : esym(tr.accel->enc.data()), offsets(tr.t->offsets), sizes(tr.t->cdf_sizes), stride(tr.t->cdf_stride),
        n_cdf(tr.t->n_cdf) {}
};

// One symbol into one stream's state; false on an index out of range.
template <bool JUMP, class SymT, class IdxT>
__attribute__((always_inline)) inline bool encode_symbol(Encoder& enc, SymT symbol, IdxT index, const EncTables& T) {
  const int32_t ci = (int32_t)index;
  if (__builtin_expect((uint32_t)ci >= (uint32_t)T.n_cdf, 0)) return false;
  const int32_t max_value = T.sizes[ci] - 2;
  int32_t value = (int32_t)symbol - T.offsets[ci];
  if (__builtin_expect(value < 0 || value >= max_value, 0)) {
    const uint32_t raw = value < 0 ? (uint32_t)(-2 * (int64_t)value - 1) : (uint32_t)(2 * ((int64_t)value - max_value));
    value = max_value;
    int32_t nb = 0;
    while (nb < 8 && (raw >> (nb * kBypassBits)) != 0) ++nb;
    // forward order is: [run of 15s][remainder][nibble 0 .. nibble nb-1]; feed it reversed
    for (int32_t j = nb - 1; j >= 0; --j) enc.put_bits<JUMP>((raw >> (j * kBypassBits)) & kBypassMax);
    int32_t full = nb / kBypassMax, rem = nb % kBypassMax;
    enc.put_bits<JUMP>((uint32_t)rem);
    for (int32_t r = 0; r < full; ++r) enc.put_bits<JUMP>((uint32_t)kBypassMax);
  }
  enc.put<JUMP>(T.esym[(int64_t)ci * T.stride + value]);
  return true;
}

constexpr int kMaxInterleave = 4;

// K streams of n symbols each, side by side in one thread (their state-update chains overlap in the out-of-order
// core); lengths[k] = bytes written, -(bytes needed) when out_cap[k] is too small, INT64_MIN on a bad index.
// (BMI2: the variable shifts on the chain become one-cycle shrx / shlx; every x86-64 server core since 2013 has it)
template <bool BMI2, int K, class SymT, class IdxT>
#if defined(__x86_64__)
__attribute__((target(BMI2 ? "bmi2" : "sse2")))
#endif
void encode_many_impl(const SymT* const* symbols, const IdxT* const* indexes, int64_t n, const TableRef& tr, uint8_t* const* out,
                      const int64_t* out_cap, int64_t* lengths) {
  const EncTables T(tr);
  // a table symbol emits at most one word; an escape adds at most 2 + 8 + 1 nibbles (raw < 2^32) = two more words.
  // Per-thread grow-only scratch: a fresh multi-megabyte allocation per stream is an mmap / munmap pair, and a dozen
  // coder threads doing that at once serialise on the process's address-space lock.
  static thread_local std::vector<uint32_t> scratch[kMaxInterleave];
  Encoder enc[K];
  const SymT* sy[K];
  const IdxT* ix[K];
#pragma GCC unroll 8
  for (int k = 0; k < K; ++k) {
    if (scratch[k].size() < (size_t)(3 * n + 8)) scratch[k].resize((size_t)(3 * n + 8));
    enc[k].wp = scratch[k].data();
    sy[k] = symbols[k];
    ix[k] = indexes[k];
  }
  bool ok = true;
  for (int64_t i = n - 1; i >= 0 && ok; --i) {
#pragma GCC unroll 8
    for (int k = 0; k < K; ++k) ok &= encode_symbol<K == 1>(enc[k], sy[k][i], ix[k][i], T);
  }
  for (int k = 0; k < K; ++k) {
    if (!ok) { lengths[k] = std::numeric_limits<int64_t>::min(); continue; }
    const uint32_t* words = scratch[k].data();
    const int64_t nwords = (int64_t)(enc[k].wp - words);
    const int64_t nbytes = 4 * (nwords + 2);
    if (nbytes > out_cap[k]) { lengths[k] = -nbytes; continue; }
    uint8_t* o = out[k];
    uint32_t w0 = (uint32_t)enc[k].x, w1 = (uint32_t)(enc[k].x >> 32);
    std::memcpy(o, &w0, 4);
    std::memcpy(o + 4, &w1, 4);
    size_t pos = 8;
    for (int64_t j = nwords; j-- > 0; pos += 4) std::memcpy(o + pos, &words[(size_t)j], 4);
    lengths[k] = nbytes;
  }
}

template <int K, class SymT, class IdxT>
void encode_many(const SymT* const* symbols, const IdxT* const* indexes, int64_t n, const TableRef& tr, uint8_t* const* out,
                 const int64_t* out_cap, int64_t* lengths) {
#if defined(__x86_64__)
  static const bool have_bmi2 = __builtin_cpu_supports("bmi2");
  if (have_bmi2) return encode_many_impl<true, K>(symbols, indexes, n, tr, out, out_cap, lengths);
#endif
  return encode_many_impl<false, K>(symbols, indexes, n, tr, out, out_cap, lengths);
}

template <class SymT, class IdxT>
int64_t encode_one(const SymT* symbols, const IdxT* indexes, int64_t n, const TableRef& tr, uint8_t* out,
                   int64_t out_cap) {
  int64_t len = 0;
  encode_many<1>(&symbols, &indexes, n, tr, &out, &out_cap, &len);
  return len;
}

struct Decoder {
  std::vector<uint32_t> buf;          // private copy of the stream (crdr_rans_decoder_set_stream) ...
  const uint32_t* words = nullptr;    // ... or the caller's bytes (crdr_rans_decoder_set_stream_view); either way the words
  size_t nwords = 0;
  size_t pos = 0;
  uint64_t x = 0;
};

// Slot search: the bucket index gives the first candidate slot s0; the answer is s0 + #{j >= 1 : cdf[s0 + j] <= cum}.
// With 64 cumulative values per bucket a Gaussian / logistic row has a handful of slots per bucket, so one 8-wide
// compare (AVX2) or a short scalar walk resolves it; both continue in steps while a whole group compared <= cum.
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2,bmi2"))) static inline int32_t find_slot_avx2(const int32_t* row, int32_t s, int32_t cum) {
  const __m256i c = _mm256_set1_epi32(cum);
  for (;;) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(row + s + 1));
    const int gt = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpgt_epi32(v, c)));   // bit j: row[s+1+j] > cum
    if (gt) return s + __builtin_ctz((unsigned)gt);
    s += 8;
  }
}
#endif
static inline int32_t find_slot_scalar(const int32_t* row, int32_t s, int32_t cum) {
  while (row[s + 1] <= cum) ++s;
  return s;
}

// Decoder state of one stream held in locals (registers) while a loop runs.
struct DecState {
  uint64_t x;
  size_t pos;
  const uint32_t* buf;
  size_t nwords;
  inline uint32_t get_bits() {
    const uint32_t val = (uint32_t)(x & ((1u << kBypassBits) - 1));
    x >>= kBypassBits;
    if (x < kRansL) { x = (x << 32) | (pos < nwords ? buf[pos] : 0u); ++pos; }
    return val;
  }
  // the escape code behind the table's last slot: [run of 15s][remainder] = nibble count, then the nibbles
  inline int32_t escape(int32_t max_value) {
    int32_t val = (int32_t)get_bits();
    int32_t nb = val;
    while (val == kBypassMax) { val = (int32_t)get_bits(); nb += val; }
    int32_t raw = 0;
    for (int32_t j = 0; j < nb; ++j) raw |= (int32_t)get_bits() << (j * kBypassBits);
    const int32_t value = raw >> 1;
    return (raw & 1) ? -value - 1 : value + max_value;
  }
};
inline DecState load_state(const Decoder* d) { return DecState{d->x, d->pos, d->words, d->nwords}; }
inline void store_state(Decoder* d, const DecState& st) { d->x = st.x; d->pos = std::min(st.pos, d->nwords); }

struct DecTables {
  const uint64_t* bucket;
  const int32_t* cdf_pad;
  int64_t pstride;
  const int32_t* sizes;
  const int32_t* offsets;
  int32_t n_cdf;
  explicit DecTables(const TableRef& tr)
      : bucket(tr.accel->bucket.data()), cdf_pad(tr.accel->cdf_pad.data()), pstride(tr.accel->pad_stride),
        sizes(tr.t->cdf_sizes), offsets(tr.t->offsets), n_cdf(tr.t->n_cdf) {}
};

// One symbol of one stream.  The pure-bucket shortcut is taken with a jump: it is right for most of the probability
// mass, and it shortens the serial chain x -> slot -> x from ~40 to ~20 cycles.
// returns false on an index out of range
template <bool AVX2, bool RENORM_JUMP, class IdxT, class OutT>
#if defined(__x86_64__)
__attribute__((target("avx2,bmi2"), always_inline))
#else
__attribute__((always_inline))
#endif
inline bool decode_symbol(DecState& st, const DecTables& T, IdxT index, OutT* out) {
  const int32_t ci = (int32_t)index;
  if (__builtin_expect((uint32_t)ci >= (uint32_t)T.n_cdf, 0)) return false;
  const int32_t cum = (int32_t)(st.x & ((1u << kPrecision) - 1));
  const uint64_t e = T.bucket[(int64_t)ci * kBuckets + (cum >> kBucketShift)];
  int32_t s;
  uint32_t start, freq;
  if (__builtin_expect((e >> 48) != 0, 1)) {
    s = (int32_t)((e >> 32) & 0xffffu);
    start = (uint32_t)(e & 0xffffu);
    freq = (uint32_t)((e >> 16) & 0xffffu);
  } else {
    const int32_t* row = T.cdf_pad + (int64_t)ci * T.pstride;
    s = (int32_t)((e >> 32) & 0xffffu);
#if defined(__x86_64__)
    s = AVX2 ? find_slot_avx2(row, s, cum) : find_slot_scalar(row, s, cum);
#else
    s = find_slot_scalar(row, s, cum);
#endif
    start = (uint32_t)row[s];
    freq = (uint32_t)(row[s + 1] - row[s]);
  }
  uint64_t x = (uint64_t)freq * (st.x >> kPrecision) + (uint64_t)cum - start;
  // branch-free renormalisation (the refill is due every few symbols, unpredictably)
  if (RENORM_JUMP) {
    // one stream: a jump.  At the codec's rates (about a bit per symbol) a refill is due every few dozen symbols, so
    // the jump predicts well and keeps compare + select off the serial chain
    if (__builtin_expect(x < kRansL, 0)) {
      x = (x << 32) | (st.pos < st.nwords ? st.buf[st.pos] : 0u);
      ++st.pos;
    }
    st.x = x;
  } else {
    // interleaved streams: branch free (shifts and masks; a ?: compiles to a jump), a misprediction would flush the
    // work in flight for all of them
    const uint64_t need = (uint64_t)(x < kRansL);
    const uint64_t w = st.pos < st.nwords ? st.buf[st.pos] : 0u;
    st.x = (x << (need << 5)) | (w & (0 - need));
    st.pos += need;
  }
  int32_t value = s;
  const int32_t max_value = T.sizes[ci] - 2;
  if (__builtin_expect(value == max_value, 0)) value = st.escape(max_value);
  *out = (OutT)(value + T.offsets[ci]);
  return true;
}

// K streams of equal length side by side in one thread: their serial chains overlap in the out-of-order core
template <bool AVX2, int K, class IdxT, class OutT>
#if defined(__x86_64__)
__attribute__((target("avx2,bmi2")))
#endif
int decode_loop(Decoder* const* d, const IdxT* const* indexes, int64_t n, const TableRef& tr, OutT* const* out) {
  const DecTables T(tr);
  DecState st[K];
  const IdxT* ix[K];
  OutT* o[K];
#pragma GCC unroll 8
  for (int k = 0; k < K; ++k) { st[k] = load_state(d[k]); ix[k] = indexes[k]; o[k] = out[k]; }
  int bad = 0;
  for (int64_t i = 0; i < n; ++i) {
#pragma GCC unroll 8
    for (int k = 0; k < K; ++k)
      if (!decode_symbol<AVX2, K == 1>(st[k], T, ix[k][i], o[k] + i)) { bad = 1; goto done; }
  }
done:
#pragma GCC unroll 8
  for (int k = 0; k < K; ++k) store_state(d[k], st[k]);
  return bad;
}

template <int K, class IdxT, class OutT>
int decode_some(Decoder* const* d, const IdxT* const* indexes, int64_t n, const TableRef& tr, OutT* const* out) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
  if (have_avx2) return decode_loop<true, K>(d, indexes, n, tr, out);
#endif
  return decode_loop<false, K>(d, indexes, n, tr, out);
}

// How many equal-length streams one thread should code side by side: K interleaved streams cost about
// 1 + step (K - 1) single-stream times (measured: decoder 0.4 -- its chain is latency bound; encoder 1.0 -- with the
// renormalisation jump its single-stream loop is throughput bound already, so it never interleaves unless
// CRDR_CODER_INTERLEAVE forces it), and the job ends with its last round of bundles.
int pick_interleave(int32_t count, int nt, double step) {
  static const int forced = [] { const char* e = std::getenv("CRDR_CODER_INTERLEAVE"); return e && *e ? std::atoi(e) : 0; }();
  if (forced >= 1) return std::min(forced, kMaxInterleave);
  int best = 1;
  double best_cost = 1e30;
  for (int k = 1; k <= kMaxInterleave; ++k) {
    const int bundles = (count + k - 1) / k;
    const int rounds = (bundles + nt - 1) / nt;
    const double cost = rounds * (1.0 + step * (k - 1));
    if (cost < best_cost - 1e-9) { best_cost = cost; best = k; }
  }
  return best;
}

// ---------------------------------------------------------------------------------------------------------------
// Coder threads.  One persistent pool per process, sized and pinned for a one-process-per-GPU job: the CPUs this process
// may run on are cut into LOCAL_WORLD_SIZE equal slices and the pool lives on the slice of LOCAL_RANK, one worker per
// CPU (at most kMaxWorkers) -- eight ranks no longer put 8 x 12 freshly spawned threads on 32 cores.  The calling thread
// works too.  Workers spin for a few microseconds after a job (the codec issues coder calls back to back: z then y,
// chunk after chunk) and then block on a condition variable; pinned to distinct CPUs, their wake-ups do not pile up on
// the waker's core.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxWorkers = 32;

class Pool {
 public:
  static Pool& get() {
    static Pool p;
    return p;
  }
  int size() const { return (int)workers_.size() + 1; }
  int first_cpu() const { return first_cpu_; }

  template <class F>
  void run(int32_t count, int32_t threads, F&& fn) {
    if (count <= 0) return;
    int nt = threads > 0 ? std::min(threads, size()) : size();
    nt = std::max(1, std::min<int>(nt, count));
    if (nt == 1) { for (int32_t i = 0; i < count; ++i) fn(i); return; }
    std::lock_guard<std::mutex> serial(run_mutex_);   // one job at a time
    std::function<void(int32_t)> job = [&](int32_t i) { fn(i); };
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &job;
      count_ = count;
      helpers_ = nt - 1;
      next_.store(0, std::memory_order_relaxed);
      pending_.store(nt - 1, std::memory_order_relaxed);
      generation_.fetch_add(1, std::memory_order_release);
    }
    cv_work_.notify_all();
    for (int32_t i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < count;) fn(i);
    // wait for the helpers that picked this generation up
    for (int spin = 0; pending_.load(std::memory_order_acquire) != 0; ++spin) {
      if (spin > 2000) {
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0; });
        break;
      }
    }
    std::lock_guard<std::mutex> lk(m_);
    job_ = nullptr;
  }

 private:
  Pool() {
    std::vector<int> cpus;
#ifdef __linux__
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0)
      for (int c = 0; c < CPU_SETSIZE; ++c)
        if (CPU_ISSET(c, &set)) cpus.push_back(c);
#endif
    if (cpus.empty()) {
      const int hc = std::max(1u, std::thread::hardware_concurrency());
      for (int c = 0; c < hc; ++c) cpus.push_back(c);
    }
    auto env_int = [](const char* name, int dflt) { const char* e = std::getenv(name); return e && *e ? std::atoi(e) : dflt; };
    int lw = env_int("LOCAL_WORLD_SIZE", env_int("WORLD_SIZE", 1));
    int lr = env_int("LOCAL_RANK", env_int("RANK", 0));
    if (lw < 1) lw = 1;
    lr = ((lr % lw) + lw) % lw;
    int per = std::max<int>(1, (int)cpus.size() / lw);
    int want = env_int("CRDR_CODER_THREADS", per);
    want = std::max(1, std::min({want, kMaxWorkers, std::max(per, 1)}));
    const int base = std::min<int>(lr * per, (int)cpus.size() - 1);
    first_cpu_ = cpus[base];
    const bool pin = env_int("CRDR_CODER_PIN", 1) != 0 && (int)cpus.size() >= 2;
    for (int w = 1; w < want; ++w) {   // slot 0 of the slice is left to the calling thread
      const int cpu = cpus[std::min<int>(base + w, (int)cpus.size() - 1)];
      workers_.emplace_back([this, cpu, pin] { worker(cpu, pin); });
    }
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      generation_.fetch_add(1, std::memory_order_release);
    }
    cv_work_.notify_all();
    for (auto& t : workers_) t.join();
  }

  void worker(int cpu, bool pin) {
#ifdef __linux__
    if (pin) {
      cpu_set_t set;
      CPU_ZERO(&set);
      CPU_SET(cpu, &set);
      pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
#endif
    uint64_t seen = 0;
    for (;;) {
      // short spin (back-to-back coder calls), then sleep
      bool fresh = false;
      for (int spin = 0; spin < 4000; ++spin) {
        if (generation_.load(std::memory_order_acquire) != seen) { fresh = true; break; }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      std::function<void(int32_t)>* job = nullptr;
      int32_t count = 0;
      {
        std::unique_lock<std::mutex> lk(m_);
        if (!fresh) cv_work_.wait(lk, [&] { return generation_.load(std::memory_order_acquire) != seen; });
        seen = generation_.load(std::memory_order_acquire);
        if (stop_) return;
        if (helpers_ <= 0 || job_ == nullptr) continue;   // this job has all the helpers it asked for
        --helpers_;
        job = job_;
        count = count_;
      }
      for (int32_t i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < count;) (*job)(i);
      if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> lk(m_);
        cv_done_.notify_all();
      }
    }
  }

  std::vector<std::thread> workers_;
  std::mutex m_, run_mutex_;
  std::condition_variable cv_work_, cv_done_;
  std::atomic<uint64_t> generation_{0};
  std::atomic<int32_t> next_{0};
  std::atomic<int32_t> pending_{0};
  std::function<void(int32_t)>* job_ = nullptr;
  int32_t count_ = 0;
  int helpers_ = 0;
  bool stop_ = false;
  int first_cpu_ = 0;
};

template <class F>
void parallel_for(int32_t count, int32_t threads, F&& fn) {
  Pool::get().run(count, threads, std::forward<F>(fn));
}

int effective_threads(int32_t count, int32_t threads) {
  const int pool = Pool::get().size();
  const int nt = threads > 0 ? std::min<int>(threads, pool) : pool;
  return std::max(1, std::min<int>(nt, count));
}

bool equal_lengths(int32_t count, const int64_t* n) {
  for (int32_t i = 1; i < count; ++i)
    if (n[i] != n[0]) return false;
  return true;
}

// More streams than coder threads (one process per GPU on a slice of the host's cores): every thread takes bundles of
// up to four equal-length streams and codes them side by side.  Same bytes / symbols as one stream at a time.
template <class SymT, class IdxT>
int encode_batch_impl(int32_t count, const SymT* const* symbols, const IdxT* const* indexes, const int64_t* n,
                      const crdr_cdf_tables* t, uint8_t* const* out, const int64_t* out_cap, int64_t* lengths, int32_t threads) {
  if (count <= 0) return 0;
  if (!t) return 1;
  const TableRef tr(t);
  const int K = equal_lengths(count, n) ? pick_interleave(count, effective_threads(count, threads), 1.0) : 1;
  const int32_t bundles = (count + K - 1) / K;
  parallel_for(bundles, threads, [&](int32_t b) {
    const int32_t lo = b * K, m = std::min<int32_t>(K, count - lo);
    switch (m) {
      case 1: encode_many<1>(symbols + lo, indexes + lo, n[lo], tr, out + lo, out_cap + lo, lengths + lo); break;
      case 2: encode_many<2>(symbols + lo, indexes + lo, n[lo], tr, out + lo, out_cap + lo, lengths + lo); break;
      case 3: encode_many<3>(symbols + lo, indexes + lo, n[lo], tr, out + lo, out_cap + lo, lengths + lo); break;
      default: encode_many<4>(symbols + lo, indexes + lo, n[lo], tr, out + lo, out_cap + lo, lengths + lo); break;
    }
  });
  for (int32_t i = 0; i < count; ++i)
    if (lengths[i] < 0) return 1;
  return 0;
}

template <class IdxT>
int decode_batch_impl(int32_t count, void* const* decoders, const IdxT* const* indexes, const int64_t* n,
                      const crdr_cdf_tables* t, int32_t* const* out, int32_t threads) {
  if (count <= 0) return 0;
  if (!t) return 1;
  const TableRef tr(t);
  Decoder* const* d = reinterpret_cast<Decoder* const*>(decoders);
  const int K = equal_lengths(count, n) ? pick_interleave(count, effective_threads(count, threads), 0.4) : 1;
  const int32_t bundles = (count + K - 1) / K;
  std::atomic<int> bad{0};
  parallel_for(bundles, threads, [&](int32_t b) {
    const int32_t lo = b * K, m = std::min<int32_t>(K, count - lo);
    int rc;
    switch (m) {
      case 1: rc = decode_some<1>(d + lo, indexes + lo, n[lo], tr, out + lo); break;
      case 2: rc = decode_some<2>(d + lo, indexes + lo, n[lo], tr, out + lo); break;
      case 3: rc = decode_some<3>(d + lo, indexes + lo, n[lo], tr, out + lo); break;
      default: rc = decode_some<4>(d + lo, indexes + lo, n[lo], tr, out + lo); break;
    }
    if (rc) bad.store(1);
  });
  return bad.load();
}

}  // namespace

extern "C" {

int crdr_pmf_to_quantized_cdf(const float* pmf, int64_t n, int32_t precision, uint32_t* cdf) {
  if (n <= 0 || precision < 1 || precision > 16) return 1;
  for (int64_t i = 0; i < n; ++i)
    if (!(pmf[i] >= 0.f) || !std::isfinite(pmf[i])) return 1;
  const uint32_t one = 1u << precision;
  cdf[0] = 0;
  uint32_t total = 0;
  for (int64_t i = 0; i < n; ++i) { cdf[i + 1] = (uint32_t)std::round(pmf[i] * (float)one); total += cdf[i + 1]; }
  if (total == 0) return 2;
  uint32_t run = 0;
  for (int64_t i = 1; i <= n; ++i) { run += (uint32_t)(((uint64_t)one * cdf[i]) / total); cdf[i] = run; }
  cdf[n] = one;
  for (int64_t i = 0; i < n; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    // empty bin: take one count from the narrowest bin that can spare it
    uint32_t best_freq = ~0u;
    int64_t donor = -1;
    for (int64_t j = 0; j < n; ++j) {
      const uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best_freq) { best_freq = f; donor = j; }
    }
    if (donor < 0) return 3;
    if (donor < i) { for (int64_t j = donor + 1; j <= i; ++j) --cdf[j]; }
    else { for (int64_t j = i + 1; j <= donor; ++j) ++cdf[j]; }
  }
  return 0;
}

int64_t crdr_rans_encode_with_indexes(const int32_t* symbols, const int32_t* indexes, int64_t n,
                                      const crdr_cdf_tables* t, uint8_t* out, int64_t out_cap) {
  if (!t) return std::numeric_limits<int64_t>::min();
  const TableRef tr(t);
  return encode_one(symbols, indexes, n, tr, out, out_cap);
}

void* crdr_rans_tables_prepare(const crdr_cdf_tables* t) {
  if (!t || !t->cdfs || !t->cdf_sizes || !t->offsets || t->n_cdf <= 0 || t->cdf_stride <= 0) return nullptr;
  Prepared* p = new Prepared();
  p->cdfs.assign(t->cdfs, t->cdfs + (size_t)t->n_cdf * (size_t)t->cdf_stride);
  p->sizes.assign(t->cdf_sizes, t->cdf_sizes + t->n_cdf);
  p->offsets.assign(t->offsets, t->offsets + t->n_cdf);
  p->view = crdr_cdf_tables{p->cdfs.data(), t->cdf_stride, p->sizes.data(), p->offsets.data(), t->n_cdf, nullptr};
  p->accel = build_accel(&p->view);
  return p;
}

void crdr_rans_tables_free(void* prepared) { delete static_cast<Prepared*>(prepared); }

int crdr_rans_encode_batch(int32_t count, const int32_t* const* symbols, const int32_t* const* indexes,
                           const int64_t* n, const crdr_cdf_tables* t, uint8_t* const* out, const int64_t* out_cap,
                           int64_t* lengths, int32_t threads) {
  return encode_batch_impl(count, symbols, indexes, n, t, out, out_cap, lengths, threads);
}

int crdr_rans_encode_batch_i16u8(int32_t count, const int16_t* const* symbols, const uint8_t* const* indexes,
                                 const int64_t* n, const crdr_cdf_tables* t, uint8_t* const* out, const int64_t* out_cap,
                                 int64_t* lengths, int32_t threads) {
  return encode_batch_impl(count, symbols, indexes, n, t, out, out_cap, lengths, threads);
}

int crdr_rans_pool_info(int32_t* threads, int32_t* first_cpu) {
  Pool& p = Pool::get();
  if (threads) *threads = p.size();
  if (first_cpu) *first_cpu = p.first_cpu();
  return 0;
}

void* crdr_rans_decoder_new(void) { return new Decoder(); }
void crdr_rans_decoder_free(void* dec) { delete static_cast<Decoder*>(dec); }

int crdr_rans_decoder_set_stream(void* dec, const uint8_t* stream, int64_t nbytes) {
  Decoder* d = static_cast<Decoder*>(dec);
  if (!d || nbytes < 8) return 1;
  d->buf.assign((size_t)((nbytes + 3) / 4), 0u);
  std::memcpy(d->buf.data(), stream, (size_t)nbytes);
  d->words = d->buf.data();
  d->nwords = d->buf.size();
  d->x = (uint64_t)d->buf[0] | ((uint64_t)d->buf[1] << 32);
  d->pos = 2;
  return 0;
}

int crdr_rans_decoder_set_stream_view(void* dec, const uint8_t* stream, int64_t nbytes) {
  Decoder* d = static_cast<Decoder*>(dec);
  if (!d || !stream || nbytes < 8) return 1;
  if (nbytes % 4 != 0 || ((uintptr_t)stream & 3) != 0) return crdr_rans_decoder_set_stream(dec, stream, nbytes);
  d->buf.clear();
  d->words = reinterpret_cast<const uint32_t*>(stream);
  d->nwords = (size_t)(nbytes / 4);
  d->x = (uint64_t)d->words[0] | ((uint64_t)d->words[1] << 32);
  d->pos = 2;
  return 0;
}

int crdr_rans_decoder_decode_stream(void* dec, const int32_t* indexes, int64_t n, const crdr_cdf_tables* t,
                                    int32_t* out) {
  if (!dec || !t) return 1;
  const TableRef tr(t);
  Decoder* d = static_cast<Decoder*>(dec);
  return decode_some<1>(&d, &indexes, n, tr, &out);
}

int crdr_rans_decode_batch(int32_t count, void* const* decoders, const int32_t* const* indexes, const int64_t* n,
                           const crdr_cdf_tables* t, int32_t* const* out, int32_t threads) {
  return decode_batch_impl(count, decoders, indexes, n, t, out, threads);
}

int crdr_rans_decode_batch_u8(int32_t count, void* const* decoders, const uint8_t* const* indexes, const int64_t* n,
                              const crdr_cdf_tables* t, int32_t* const* out, int32_t threads) {
  return decode_batch_impl(count, decoders, indexes, n, t, out, threads);
}

}  // extern "C"
