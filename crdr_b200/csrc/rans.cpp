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
//  * decoder: per row a 256-bucket index over the 16-bit cumulative range; the slot search starts at the bucket's
//    first slot and walks forward (Gaussian / logistic tables put a handful of slots in a bucket).
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
  std::vector<uint16_t> bucket;   // [n_cdf][kBuckets]: first slot whose interval reaches into the bucket
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
    uint16_t* bk = a->bucket.data() + (size_t)r * kBuckets;
    int32_t sidx = 0;
    for (int32_t b = 0; b < kBuckets; ++b) {
      const int32_t cum = b << kBucketShift;
      while (sidx + 2 < size && cdf[sidx + 1] <= cum) ++sidx;
      bk[b] = (uint16_t)sidx;
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

  // branch-free renormalisation: the low word is always stored, the pointer only advances when it was due
  inline void renorm(uint64_t x_max) {
    const bool emit = x >= x_max;
    *wp = (uint32_t)x;
    wp += emit;
    x = emit ? (x >> 32) : x;
  }
  inline void put(const EncSym& e) {
    renorm(((kRansL >> kPrecision) << 32) * (uint64_t)e.freq);
    const uint64_t q = (uint64_t)(((unsigned __int128)x * e.rcp_freq) >> 64) >> e.rcp_shift;   // == x / freq
    x = x + e.bias + q * e.cmpl_freq;
  }
  inline void put_bits(uint32_t val) {
    renorm(((kRansL >> 16) << 32) * (uint64_t)(1u << (16 - kBypassBits)));
    x = (x << kBypassBits) | val;
  }
};

template <class SymT, class IdxT>
int64_t encode_one(const SymT* symbols, const IdxT* indexes, int64_t n, const TableRef& tr, uint8_t* out,
                   int64_t out_cap) {
  Encoder enc;
  const crdr_cdf_tables* t = tr.t;
  const EncSym* esym = tr.accel->enc.data();
  // a table symbol emits at most one word; an escape adds at most 2 + 8 + 1 nibbles (raw < 2^32) = two more words.
  // Per-thread grow-only scratch: a fresh multi-megabyte allocation per stream is an mmap / munmap pair, and a dozen
  // coder threads doing that at once serialise on the process's address-space lock.
  static thread_local std::vector<uint32_t> scratch;
  if (scratch.size() < (size_t)(3 * n + 8)) scratch.resize((size_t)(3 * n + 8));
  uint32_t* const words = scratch.data();
  enc.wp = words;
  const int32_t n_cdf = t->n_cdf;
  const int32_t* offsets = t->offsets;
  const int32_t* sizes = t->cdf_sizes;
  const int64_t stride = t->cdf_stride;
  for (int64_t i = n - 1; i >= 0; --i) {
    const int32_t ci = (int32_t)indexes[i];
    if ((uint32_t)ci >= (uint32_t)n_cdf) return std::numeric_limits<int64_t>::min();
    const int32_t max_value = sizes[ci] - 2;
    int32_t value = (int32_t)symbols[i] - offsets[ci];
    if (value < 0 || value >= max_value) {
      const uint32_t raw = value < 0 ? (uint32_t)(-2 * (int64_t)value - 1) : (uint32_t)(2 * ((int64_t)value - max_value));
      value = max_value;
      int32_t nb = 0;
      while (nb < 8 && (raw >> (nb * kBypassBits)) != 0) ++nb;
      // forward order is: [run of 15s][remainder][nibble 0 .. nibble nb-1]; feed it reversed
      for (int32_t j = nb - 1; j >= 0; --j) enc.put_bits((raw >> (j * kBypassBits)) & kBypassMax);
      int32_t full = nb / kBypassMax, rem = nb % kBypassMax;
      enc.put_bits((uint32_t)rem);
      for (int32_t r = 0; r < full; ++r) enc.put_bits((uint32_t)kBypassMax);
    }
    enc.put(esym[(int64_t)ci * stride + value]);
  }
  const int64_t nwords = (int64_t)(enc.wp - words);
  const int64_t nbytes = 4 * (nwords + 2);
  if (nbytes > out_cap) return -nbytes;
  uint32_t w0 = (uint32_t)enc.x, w1 = (uint32_t)(enc.x >> 32);
  std::memcpy(out, &w0, 4);
  std::memcpy(out + 4, &w1, 4);
  size_t k = 8;
  for (int64_t j = nwords; j-- > 0; k += 4) std::memcpy(out + k, &words[(size_t)j], 4);
  return nbytes;
}

struct Decoder {
  std::vector<uint32_t> buf;
  size_t pos = 0;
  uint64_t x = 0;

  inline uint32_t next_word() { return pos < buf.size() ? buf[pos++] : 0u; }
  inline uint32_t get_bits() {
    uint32_t val = (uint32_t)(x & ((1u << kBypassBits) - 1));
    x >>= kBypassBits;
    if (x < kRansL) x = (x << 32) | next_word();
    return val;
  }
};

// Slot search: the bucket index gives the first candidate slot s0; the answer is s0 + #{j >= 1 : cdf[s0 + j] <= cum}.
// With 64 cumulative values per bucket a Gaussian / logistic row has a handful of slots per bucket, so one 8-wide
// compare (AVX2) or a short scalar walk resolves it; both continue in steps while a whole group compared <= cum.
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) static inline int32_t find_slot_avx2(const int32_t* row, int32_t s, int32_t cum) {
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

// returns 0 ok, 1 index out of range
template <bool AVX2, class IdxT, class OutT>
#if defined(__x86_64__)
__attribute__((target("avx2")))
#endif
int decode_loop(Decoder* d, const IdxT* indexes, int64_t n, const TableRef& tr, OutT* out) {
  const crdr_cdf_tables* t = tr.t;
  const uint16_t* bucket = tr.accel->bucket.data();
  const int32_t* cdf_pad = tr.accel->cdf_pad.data();
  const int64_t pstride = tr.accel->pad_stride;
  const int32_t n_cdf = t->n_cdf;
  const int32_t* sizes = t->cdf_sizes;
  const int32_t* offsets = t->offsets;
  uint64_t x = d->x;
  size_t pos = d->pos;
  const uint32_t* buf = d->buf.data();
  const size_t nwords = d->buf.size();
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = (int32_t)indexes[i];
    if ((uint32_t)ci >= (uint32_t)n_cdf) { d->x = x; d->pos = pos; return 1; }
    const int32_t* row = cdf_pad + (int64_t)ci * pstride;
    const int32_t max_value = sizes[ci] - 2;
    const int32_t cum = (int32_t)(x & ((1u << kPrecision) - 1));
    int32_t s = bucket[(int64_t)ci * kBuckets + (cum >> kBucketShift)];
#if defined(__x86_64__)
    s = AVX2 ? find_slot_avx2(row, s, cum) : find_slot_scalar(row, s, cum);
#else
    s = find_slot_scalar(row, s, cum);
#endif
    const uint32_t start = (uint32_t)row[s], freq = (uint32_t)(row[s + 1] - row[s]);
    x = (uint64_t)freq * (x >> kPrecision) + (uint64_t)cum - start;
    // branch-free renormalisation (the refill is due every few symbols, unpredictably)
    const bool need = x < kRansL;
    const uint32_t w = pos < nwords ? buf[pos] : 0u;
    x = need ? ((x << 32) | w) : x;
    pos += need;
    int32_t value = s;
    if (__builtin_expect(value == max_value, 0)) {
      d->x = x; d->pos = pos;
      int32_t val = (int32_t)d->get_bits();
      int32_t nb = val;
      while (val == kBypassMax) { val = (int32_t)d->get_bits(); nb += val; }
      int32_t raw = 0;
      for (int32_t j = 0; j < nb; ++j) raw |= (int32_t)d->get_bits() << (j * kBypassBits);
      value = raw >> 1;
      value = (raw & 1) ? -value - 1 : value + max_value;
      x = d->x; pos = d->pos;
    }
    out[i] = (OutT)(value + offsets[ci]);
  }
  d->x = x;
  d->pos = pos;
  return 0;
}

template <class IdxT, class OutT>
int decode_some(Decoder* d, const IdxT* indexes, int64_t n, const TableRef& tr, OutT* out) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) return decode_loop<true>(d, indexes, n, tr, out);
#endif
  return decode_loop<false>(d, indexes, n, tr, out);
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
  const TableRef tr(t);
  parallel_for(count, threads, [&](int32_t i) { lengths[i] = encode_one(symbols[i], indexes[i], n[i], tr, out[i], out_cap[i]); });
  for (int32_t i = 0; i < count; ++i)
    if (lengths[i] < 0) return 1;
  return 0;
}

int crdr_rans_encode_batch_i16u8(int32_t count, const int16_t* const* symbols, const uint8_t* const* indexes,
                                 const int64_t* n, const crdr_cdf_tables* t, uint8_t* const* out, const int64_t* out_cap,
                                 int64_t* lengths, int32_t threads) {
  const TableRef tr(t);
  parallel_for(count, threads, [&](int32_t i) { lengths[i] = encode_one(symbols[i], indexes[i], n[i], tr, out[i], out_cap[i]); });
  for (int32_t i = 0; i < count; ++i)
    if (lengths[i] < 0) return 1;
  return 0;
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
  d->x = (uint64_t)d->buf[0] | ((uint64_t)d->buf[1] << 32);
  d->pos = 2;
  return 0;
}

int crdr_rans_decoder_decode_stream(void* dec, const int32_t* indexes, int64_t n, const crdr_cdf_tables* t,
                                    int32_t* out) {
  const TableRef tr(t);
  return decode_some(static_cast<Decoder*>(dec), indexes, n, tr, out);
}

int crdr_rans_decode_batch(int32_t count, void* const* decoders, const int32_t* const* indexes, const int64_t* n,
                           const crdr_cdf_tables* t, int32_t* const* out, int32_t threads) {
  std::atomic<int> bad{0};
  const TableRef tr(t);
  parallel_for(count, threads, [&](int32_t i) {
    if (decode_some(static_cast<Decoder*>(decoders[i]), indexes[i], n[i], tr, out[i])) bad.store(1);
  });
  return bad.load();
}

int crdr_rans_decode_batch_u8(int32_t count, void* const* decoders, const uint8_t* const* indexes, const int64_t* n,
                              const crdr_cdf_tables* t, int32_t* const* out, int32_t threads) {
  std::atomic<int> bad{0};
  const TableRef tr(t);
  parallel_for(count, threads, [&](int32_t i) {
    if (decode_some(static_cast<Decoder*>(decoders[i]), indexes[i], n[i], tr, out[i])) bad.store(1);
  });
  return bad.load();
}

}  // extern "C"
