// Host side of K1 for PAGEABLE fp32 codes: what `inference_one_epoch` hands over is a `torch.cat` of `.cpu()`
// batches (trainers/base.py:291-304) -- ordinary pageable memory.  The GPU cannot DMA from it: the bytes have to
// pass through the host's cores once anyway (into a pinned bounce buffer).  Doing the sign test in that one pass
// turns a 4-byte-per-code copy into a 1-bit-per-code write: the cores read the codes once (the floor for pageable
// memory), and 1/32 of the bytes cross PCIe.  Bit-identical to the CUDA kernel by construction and by test
// (tests/test_gpu_parity.py::test_host_pack_equals_device_pack): bit = (x > 0), zeros and NaNs reported in the
// same flag bits.  Pinned fp32 host tensors take the same route (the cores read them ~4x faster than the DMA engine
// does); CUDA tensors and other host dtypes never come here (pack.cu: DMA + pack kernels).
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

void ch_host_pool_run(int which, int n, const std::function<void(int)>& fn);

namespace {

// ---- a small persistent pool (a std::thread per call would cost ~20 us each, a streamed gallery makes dozens
// of calls per evaluation) -----------------------------------------------------------------------------------
class Pool {
 public:
  // pool 0: the evaluating thread's calls; pool 1: the gallery loader (loader.cu), whose ONE run lasts for the whole
  // gallery -- the short packs of the evaluating thread (queries, labels, the row sample) must not queue behind it
  static Pool& get(int which = 0) {
    static Pool p[2];
    return p[which != 0];
  }
  // runs fn(t) for t in [0, n) on n threads (the caller is thread 0)
  void run(int n, const std::function<void(int)>& fn) {
    if (n <= 1) {
      fn(0);
      return;
    }
    std::lock_guard<std::mutex> serial(call_mu_);
    ensure(n - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      want_ = n - 1;
      pending_ = n - 1;
      ++epoch_;
    }
    cv_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  Pool() = default;
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void ensure(int n) {
    while (static_cast<int>(workers_.size()) < n) {
      const int id = static_cast<int>(workers_.size()) + 1;
      workers_.emplace_back([this, id] {
        uint64_t seen = 0;
        for (;;) {
          const std::function<void(int)>* fn = nullptr;
          {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return epoch_ != seen; });
            seen = epoch_;
            if (stop_) return;
            if (id <= want_) fn = fn_;
          }
          if (fn != nullptr) {
            (*fn)(id);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_one();
          }
        }
      });
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* fn_ = nullptr;
  int want_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

// one row: ncols floats -> words u32 (bit k % 32 of word k / 32 = x[k] > 0); returns flag bits (1 zero, 2 NaN)
inline uint32_t pack_row_scalar(const float* x, int k0, int ncols, uint32_t* out, int words) {
  uint32_t fl = 0;
  for (int w = k0 / 32; w < words; ++w) {
    uint32_t v = 0;
    for (int j = 0; j < 32; ++j) {
      const int k = w * 32 + j;
      if (k >= ncols) break;
      const float f = x[k];
      if (f > 0.0f) v |= 1u << j;
      if (f != f) fl |= 2u;
      else if (f == 0.0f) fl |= 1u;
    }
    out[w] = v;
  }
  return fl;
}

// software prefetch distance (bytes ahead of the row being packed; 0 = none) and hint (0 NTA, 1 T0, 2 T2): one core's
// read bandwidth is bounded by its outstanding misses, and the hardware prefetcher stops at 4 KB page boundaries
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// (16 threads on the B200 host, 512 MB of codes: 151 GB/s without, 122 with NTA @ 2 KB, 190 with T0 @ 8 KB, 198 with T2 @ 8 KB)
static const int g_pf_bytes = env_int("CH_PACK_PREFETCH", 8192);
static const int g_pf_hint = env_int("CH_PACK_PREFETCH_HINT", 2);

__attribute__((target("avx512f"))) uint32_t pack_rows_avx512(const float* src, int64_t r0, int64_t r1, int ncols,
                                                            int64_t rs, int words, uint32_t* out) {
  const __m512 zero = _mm512_setzero_ps();
  __mmask16 anyz = 0, anyn = 0;
  uint32_t fl = 0;
  const int full = ncols / 32;                       // words made of two whole 16-float vectors
  for (int64_t r = r0; r < r1; ++r) {
    const float* x = src + r * rs;
    uint32_t* o = out + r * words;
    if (g_pf_bytes > 0) {
      const char* p = reinterpret_cast<const char*>(x) + g_pf_bytes;
      for (int l = 0; l < ncols * 4; l += 64) {
        if (g_pf_hint == 0) _mm_prefetch(p + l, _MM_HINT_NTA);
        else if (g_pf_hint == 1) _mm_prefetch(p + l, _MM_HINT_T0);
        else _mm_prefetch(p + l, _MM_HINT_T2);
      }
    }
    for (int w = 0; w < full; ++w) {
      const __m512 a = _mm512_loadu_ps(x + 32 * w), b = _mm512_loadu_ps(x + 32 * w + 16);
      const __mmask16 pa = _mm512_cmp_ps_mask(a, zero, _CMP_GT_OQ), pb = _mm512_cmp_ps_mask(b, zero, _CMP_GT_OQ);
      anyz |= _mm512_cmp_ps_mask(a, zero, _CMP_EQ_OQ) | _mm512_cmp_ps_mask(b, zero, _CMP_EQ_OQ);
      anyn |= _mm512_cmp_ps_mask(a, a, _CMP_UNORD_Q) | _mm512_cmp_ps_mask(b, b, _CMP_UNORD_Q);
      o[w] = static_cast<uint32_t>(pa) | (static_cast<uint32_t>(pb) << 16);
    }
    if (full < words) {
      // the last, partial word (nbit = 16, 48, ...): masked loads instead of a scalar loop (Cars196-16: the scalar
      // tail was the whole row, 1.3 ms of a 1.8 ms evaluation).  Lanes that are masked off read as 0: they are not
      // "> 0" and are kept out of the zero test.
      const int r = ncols - 32 * full;                     // 1 .. 31 codes
      const __mmask16 ml = r >= 16 ? static_cast<__mmask16>(0xffff) : static_cast<__mmask16>((1u << r) - 1u);
      const __mmask16 mh = r > 16 ? static_cast<__mmask16>((1u << (r - 16)) - 1u) : static_cast<__mmask16>(0);
      const __m512 a = _mm512_maskz_loadu_ps(ml, x + 32 * full);
      const __m512 b = mh ? _mm512_maskz_loadu_ps(mh, x + 32 * full + 16) : zero;
      const __mmask16 pa = _mm512_cmp_ps_mask(a, zero, _CMP_GT_OQ), pb = _mm512_cmp_ps_mask(b, zero, _CMP_GT_OQ);
      anyz |= (_mm512_cmp_ps_mask(a, zero, _CMP_EQ_OQ) & ml) | (_mm512_cmp_ps_mask(b, zero, _CMP_EQ_OQ) & mh);
      anyn |= (_mm512_cmp_ps_mask(a, a, _CMP_UNORD_Q) & ml) | (_mm512_cmp_ps_mask(b, b, _CMP_UNORD_Q) & mh);
      o[full] = static_cast<uint32_t>(pa) | (static_cast<uint32_t>(pb) << 16);
      for (int w = full + 1; w < words; ++w) o[w] = 0u;    // (words is a power of two: nbit = 200 has 8)
    }
  }
  if (anyz) fl |= 1u;
  if (anyn) fl |= 2u;
  return fl;
}

__attribute__((target("avx2"))) uint32_t pack_rows_avx2(const float* src, int64_t r0, int64_t r1, int ncols, int64_t rs,
                                                        int words, uint32_t* out) {
  const __m256 zero = _mm256_setzero_ps();
  int anyz = 0, anyn = 0;
  uint32_t fl = 0;
  const int full = ncols / 32;
  for (int64_t r = r0; r < r1; ++r) {
    const float* x = src + r * rs;
    uint32_t* o = out + r * words;
    for (int w = 0; w < full; ++w) {
      uint32_t v = 0;
      for (int j = 0; j < 4; ++j) {
        const __m256 a = _mm256_loadu_ps(x + 32 * w + 8 * j);
        v |= static_cast<uint32_t>(_mm256_movemask_ps(_mm256_cmp_ps(a, zero, _CMP_GT_OQ))) << (8 * j);
        anyz |= _mm256_movemask_ps(_mm256_cmp_ps(a, zero, _CMP_EQ_OQ));
        anyn |= _mm256_movemask_ps(_mm256_cmp_ps(a, a, _CMP_UNORD_Q));
      }
      o[w] = v;
    }
    if (full < words) fl |= pack_row_scalar(x, full * 32, ncols, o, words);
  }
  if (anyz) fl |= 1u;
  if (anyn) fl |= 2u;
  return fl;
}

uint32_t pack_rows_plain(const float* src, int64_t r0, int64_t r1, int ncols, int64_t rs, int words, uint32_t* out) {
  uint32_t fl = 0;
  for (int64_t r = r0; r < r1; ++r) fl |= pack_row_scalar(src + r * rs, 0, ncols, out + r * words, words);
  return fl;
}

typedef uint32_t (*pack_fn_t)(const float*, int64_t, int64_t, int, int64_t, int, uint32_t*);

// The widest variant this host runs.  CH_HOST_PACK_ISA = avx2 | plain (read at every call: a few ns beside a piece of
// >= 256 KB) narrows it, so that the variants a host without AVX-512 would take are tested on one that has it; a
// variant the host cannot run is never chosen.
pack_fn_t pick_pack_fn() {
  static const int widest = [] {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") ? 2 : __builtin_cpu_supports("avx2") ? 1 : 0;
  }();
  int level = widest;
  if (const char* e = getenv("CH_HOST_PACK_ISA")) {
    const int want = !strcmp(e, "plain") ? 0 : !strcmp(e, "avx2") ? 1 : 2;
    if (want < level) level = want;
  }
  return level == 2 ? pack_rows_avx512 : level == 1 ? pack_rows_avx2 : pack_rows_plain;
}

}  // namespace

// rows [0, n) of `src` (row stride rs floats, unit column stride) -> out (n, words) u32; returns the flag bits
// (bit 0: some value is exactly 0, bit 1: NaN) exactly as the pack kernels raise them.
uint32_t ch_host_pack_f32(const float* src, int64_t n, int ncols, int64_t rs, int words, uint32_t* out, int nthreads) {
  if (n <= 0) return 0;
  const pack_fn_t fn = pick_pack_fn();
  // small inputs are not worth waking threads for (>= 1 MB of codes per thread)
  const int64_t bytes = n * static_cast<int64_t>(ncols) * 4;
  int nt = nthreads;
  if (nt > bytes / (1 << 20)) nt = static_cast<int>(bytes / (1 << 20));
  if (nt < 1) nt = 1;
  // rows are handed out in ~256 KB pieces from a shared counter: on a virtual machine some cores run late, and a
  // static split waits for the slowest (measured on the 16-vCPU B200 host: 106 -> ~120 GB/s)
  std::vector<uint32_t> flags(static_cast<size_t>(nt), 0u);
  int64_t piece = (256 * 1024) / (static_cast<int64_t>(ncols) * 4);
  if (piece < 16) piece = 16;
  std::atomic<int64_t> next(0);
  ch_host_pool_run(0, nt, [&](int t) {
    uint32_t f = 0;
    for (;;) {
      const int64_t a = next.fetch_add(piece, std::memory_order_relaxed);
      if (a >= n) break;
      f |= fn(src, a, a + piece < n ? a + piece : n, ncols, rs, words, out);
    }
    flags[static_cast<size_t>(t)] = f;
  });
  uint32_t fl = 0;
  for (uint32_t f : flags) fl |= f;
  return fl;
}

// Calls of the evaluating thread (pool 0) are short and on the critical path of the GPU's first work; the loader's run
// (pool 1, loader.cu) is long and holds every core.  Two pools oversubscribe the cores, and the scheduler does not
// hand a freshly woken worker a core for milliseconds when 16 busy threads hold them (measured: a 0.18 ms query pack
// took 3.7 ms); pausing the loader meanwhile leaves cores idle (0.4 ms of a 5.5 ms evaluation).  So while a loader
// runs, a call of the evaluating thread is published as a JOB that the loader's threads help with between two of
// their own pieces (<= ~30 us away): each helper takes a ticket t < nt and runs fn(t).  Every fn draws pieces from a
// shared counter until none is left, so the caller's own fn(0) completes the job even if nobody helps.
struct HelpJob {
  const std::function<void(int)>* fn;
  int nt;
  std::atomic<int> ticket{1};
};
static std::atomic<HelpJob*> g_job{nullptr};
static std::atomic<int> g_inside{0};          // helpers that may hold a pointer to the job
static std::atomic<int> g_loaders{0};         // running loaders (they call ch_host_help between pieces)

void ch_host_loader_running(int delta) { g_loaders.fetch_add(delta); }

void ch_host_help() {
  if (g_job.load(std::memory_order_acquire) == nullptr) return;
  g_inside.fetch_add(1);
  HelpJob* j = g_job.load();
  if (j != nullptr) {
    const int t = j->ticket.fetch_add(1);
    if (t < j->nt) (*j->fn)(t);
  }
  g_inside.fetch_sub(1);
}

void ch_host_pool_run(int which, int n, const std::function<void(int)>& fn) {
  if (which == 0 && n > 1 && g_loaders.load() > 0) {
    static std::mutex one_job;
    std::lock_guard<std::mutex> serial(one_job);
    HelpJob job{&fn, n};
    g_job.store(&job);
    fn(0);
    g_job.store(nullptr);
    while (g_inside.load() > 0) std::this_thread::yield();    // helpers still inside fn (or about to look at the job)
    return;
  }
  Pool::get(which).run(n, fn);
}

// rows [r0, r1) of `src` -> dst (r1 - r0, words) u32 on the calling thread; returns the flag bits (loader.cu)
uint32_t ch_host_pack_rows(const float* src, int64_t r0, int64_t r1, int ncols, int64_t rs, int words, uint32_t* dst) {
  return pick_pack_fn()(src + r0 * rs, 0, r1 - r0, ncols, rs, words, dst);
}

// memcpy of a large pageable block into a pinned bounce buffer by the same pool (~256 KB pieces): a pageable
// cudaMemcpyAsync is staged by ONE driver thread at ~10 GB/s -- 0.8 ms for the 8 MB of int64 labels of a 1M-row gallery
// non-temporal stores: the next reader of the bounce buffer is a DMA engine, and lines left dirty in 16 cores' caches
// cost it ~4x (8 MB of label ids: a 1.1 ms DMA after memcpy, measured)
__attribute__((target("avx2"))) static size_t copy_nt_avx2(char* d, const char* s, size_t len) {
  size_t i = 0;
  for (; i + 128 <= len; i += 128) {
    const __m256i v0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
    const __m256i v1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
    const __m256i v2 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 64));
    const __m256i v3 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), v0);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), v1);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 64), v2);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 96), v3);
  }
  _mm_sfence();
  return i;
}
static void copy_piece(char* d, const char* s, size_t len, bool nt_ok) {
  size_t i = 0;
  if (nt_ok && (reinterpret_cast<uintptr_t>(d) & 31) == 0) i = copy_nt_avx2(d, s, len);
  if (i < len) memcpy(d + i, s + i, len - i);
}

void ch_host_copy_piece(void* dst, const void* src, size_t bytes) {
  static const bool nt_ok = __builtin_cpu_supports("avx2");
  copy_piece(static_cast<char*>(dst), static_cast<const char*>(src), bytes, nt_ok);
}

void ch_host_parallel_copy(void* dst, const void* src, size_t bytes, int nthreads) {
  const bool nt_ok = __builtin_cpu_supports("avx2");
  const size_t piece = static_cast<size_t>(256) << 10;
  int nt = nthreads;
  if (static_cast<size_t>(nt) > bytes / piece) nt = static_cast<int>(bytes / piece);
  if (nt <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::atomic<size_t> next(0);
  ch_host_pool_run(0, nt, [&](int) {
    for (;;) {
      const size_t a = next.fetch_add(piece, std::memory_order_relaxed);
      if (a >= bytes) break;
      copy_piece(static_cast<char*>(dst) + a, static_cast<const char*>(src) + a, a + piece < bytes ? piece : bytes - a, nt_ok);
    }
  });
}
