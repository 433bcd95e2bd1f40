// Native gallery loader: the host half of K1 for a HOST-resident fp32 gallery, run on its own thread from the first
// moment of an evaluation (include/concepthash_b200.h: ch_host_loader_*).
//
// `inference_one_epoch` hands over a torch.cat of `.cpu()` batches (trainers/base.py:291-304).  The host's cores are the
// bottleneck of such an evaluation: they have to read every code once (512 MB for a 1M x 128 gallery, ~4 ms on the
// 16-vCPU B200 host), the GPU needs ~2 ms for the rest.  The evaluation is therefore as fast as (a) the cores start
// early, (b) never pause, and (c) little is left to do after their last byte.  One pool run covers the whole gallery:
// the threads draw ~256 KB pieces from a shared counter, sign/bit-pack them (AVX-512 / AVX2, host_pack.cpp) into a
// ring of pinned chunk buffers, and thread 0 -- between two pieces of its own -- sends every chunk that is complete,
// in order, with one cudaMemcpyAsync + event on the caller's side stream.  Consumers (the evaluating thread) wait for
// "rows [0, r) are on their way" and make their stream wait on the event of the chunk that holds row r - 1.  The
// evaluating thread's own host packs (queries, labels, the row sample) run on a SECOND pool meanwhile.
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

void ch_host_pool_run(int which, int n, const std::function<void(int)>& fn);                       // host_pack.cpp
void ch_host_help();                     // host_pack.cpp: lends this thread to a host pack of the evaluating thread
void ch_host_loader_running(int delta);  // host_pack.cpp
uint32_t ch_host_pack_rows(const float* src, int64_t r0, int64_t r1, int ncols, int64_t rs, int words,
                           uint32_t* dst);                                                         // host_pack.cpp
int ch_ws_loader_ring(ch_ws* ws, size_t bytes, void** ring, size_t* ring_bytes, cudaEvent_t** last);   // api.cu

namespace {

__global__ void loader_or_flags_kernel(uint32_t* __restrict__ flags, uint32_t bits) { atomicOr(flags, bits); }

struct Loader {
  ch_ws* ws = nullptr;
  const float* src = nullptr;
  int64_t n = 0, rs = 0, rows_pad = 0;
  int ncols = 0, words = 0;
  uint32_t* out_dev = nullptr;
  uint32_t* flags_dev = nullptr;
  cudaStream_t st = nullptr;
  int64_t piece = 0, chunk_rows = 0;
  int nchunks = 0, slots = 0;
  uint32_t* ring = nullptr;
  cudaEvent_t* last_event = nullptr;       // (workspace) the last copy that read the ring: waited for by the next loader
  std::vector<cudaEvent_t> ev;             // per chunk, recorded behind its copy
  std::vector<std::atomic<int>> left;      // pieces of the chunk still being packed
  std::atomic<int64_t> next{0};
  std::atomic<int> freed{0};               // chunks whose copy has completed (their ring slot may be refilled)
  std::atomic<uint32_t> fl{0};
  std::atomic<bool> failed{false};
  std::mutex mu;
  std::condition_variable cv;
  int sent = 0;                            // chunks whose copy + event have been enqueued (guarded by mu)
  bool done = false;
  char err[512] = "";
  std::thread thread;

  explicit Loader(int chunks) : left(static_cast<size_t>(chunks)) {}

  void fail(const char* what, cudaError_t e) {
    std::lock_guard<std::mutex> lk(mu);
    if (!failed.load()) snprintf(err, sizeof(err), "gallery loader: %s: %s", what, cudaGetErrorString(e));
    failed.store(true);
    cv.notify_all();
  }

  // thread 0: sends every chunk that has been packed completely (in order) and frees the slots of finished copies
  void drain(int& sent_local) {
    while (sent_local < nchunks && left[static_cast<size_t>(sent_local)].load(std::memory_order_acquire) == 0) {
      const int c = sent_local;
      const int64_t r0 = static_cast<int64_t>(c) * chunk_rows;
      const int64_t r1 = r0 + chunk_rows < n ? r0 + chunk_rows : n;
      const uint32_t* buf = ring + static_cast<size_t>(c % slots) * chunk_rows * words;
      if (c == nchunks - 1) {               // flags are final once the last chunk is packed; they travel before its event
        const uint32_t f = fl.load();
        if (f != 0u && flags_dev != nullptr) {
          cudaGetLastError();               // (a cudaEventQuery that answered "not ready" is this thread's last error)
          loader_or_flags_kernel<<<1, 1, 0, st>>>(flags_dev, f);
          cudaError_t e = cudaGetLastError();
          if (e != cudaSuccess) return fail("flag kernel", e);
          __atomic_fetch_add(&ws->launches, 1, __ATOMIC_RELAXED);
        }
      }
      cudaError_t e = cudaMemcpyAsync(out_dev + static_cast<size_t>(r0) * words, buf,
                                      static_cast<size_t>(r1 - r0) * words * 4, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaEventRecord(ev[static_cast<size_t>(c)], st);
      if (e != cudaSuccess) return fail("copy", e);
      ++sent_local;
      {
        std::lock_guard<std::mutex> lk(mu);
        sent = sent_local;
      }
      cv.notify_all();
    }
    int f = freed.load(std::memory_order_relaxed);
    while (f < sent_local && cudaEventQuery(ev[static_cast<size_t>(f)]) == cudaSuccess) ++f;
    freed.store(f, std::memory_order_release);
  }

  void run() {
    if (cudaSetDevice(ws->device) != cudaSuccess) {
      fail("cudaSetDevice", cudaGetLastError());
    } else {
      int sent_local = 0;
      ch_host_loader_running(+1);
      // (one core is left to the evaluating thread: it queues the sample passes and the select launches meanwhile)
      int spare = 1;
      if (const char* e = getenv("CH_LOADER_SPARE")) spare = atoi(e);
      int nt = ws->pack_threads > 3 ? ws->pack_threads - spare : ws->pack_threads;
      if (nt < 1) nt = 1;
      ch_host_pool_run(1, nt, [&](int t) {
        uint32_t f = 0;
        for (;;) {
          if (t == 0) drain(sent_local);
          if (failed.load(std::memory_order_relaxed)) break;
          ch_host_help();                     // the evaluating thread's short packs go first (they feed the GPU's first work)
          const int64_t a = next.fetch_add(piece, std::memory_order_relaxed);
          if (a >= n) break;
          const int c = static_cast<int>(a / chunk_rows);
          while (c - slots >= freed.load(std::memory_order_acquire) && !failed.load(std::memory_order_relaxed)) {
            ch_host_help();
            if (t == 0) drain(sent_local);
            else std::this_thread::yield();
          }
          const int64_t c0 = static_cast<int64_t>(c) * chunk_rows;
          int64_t b = a + piece;
          if (b > c0 + chunk_rows) b = c0 + chunk_rows;
          if (b > n) b = n;
          f |= ch_host_pack_rows(src, a, b, ncols, rs, words,
                                 ring + (static_cast<size_t>(c % slots) * chunk_rows + static_cast<size_t>(a - c0)) * words);
          if (f != 0u) fl.fetch_or(f, std::memory_order_relaxed);
          left[static_cast<size_t>(c)].fetch_sub(1, std::memory_order_release);
        }
      });
      ch_host_loader_running(-1);
      while (sent_local < nchunks && !failed.load()) {       // pieces still in other threads' hands
        drain(sent_local);
        if (sent_local < nchunks) std::this_thread::yield();
      }
      if (!failed.load() && last_event != nullptr && nchunks > 0) {
        cudaError_t e = cudaEventRecord(*last_event, st);
        if (e != cudaSuccess) fail("event", e);
      }
    }
    std::lock_guard<std::mutex> lk(mu);
    done = true;
    cv.notify_all();
  }
};

}  // namespace

struct ch_loader {
  Loader impl;
  explicit ch_loader(int chunks) : impl(chunks) {}
};

extern "C" int ch_host_pack_threads(ch_ws* ws) { return ws ? ws->pack_threads : 0; }

extern "C" int ch_host_loader_start(ch_ws* ws, const float* codes_host, int64_t n, int nbit, int64_t row_stride,
                                    uint32_t* out_bits_dev, uint32_t* flags_dev, void* stream, ch_loader** out) {
  if (out == nullptr) CH_FAIL("null out pointer");
  *out = nullptr;
  if (ws == nullptr) CH_FAIL("null workspace");
  const int words = ch_code_words(nbit);
  if (words == 0) CH_FAIL("nbit=%d unsupported (1..%d)", nbit, CH_MAX_NBIT);
  if (n <= 0 || codes_host == nullptr || out_bits_dev == nullptr || row_stride < nbit)
    CH_FAIL("bad arguments to ch_host_loader_start");
  if ((reinterpret_cast<uintptr_t>(codes_host) & 3) != 0) CH_FAIL("codes must be 4-byte aligned");
  if (ws->pack_threads < 1) CH_FAIL("host packing is disabled (CH_PACK_THREADS=0)");
  ChDeviceGuard g(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // pieces of ~256 KB of codes; chunks (the unit that travels) of >= 32 pieces, at most ~256 per gallery
  int64_t piece = (256 * 1024) / (static_cast<int64_t>(nbit) * 4);
  if (piece < 16) piece = 16;
  int64_t per_chunk = 32;
  while ((n + piece * per_chunk - 1) / (piece * per_chunk) > 256) per_chunk *= 2;
  const int64_t chunk_rows = piece * per_chunk;
  const int nchunks = static_cast<int>((n + chunk_rows - 1) / chunk_rows);
  const size_t chunk_bytes = static_cast<size_t>(chunk_rows) * words * 4;
  void* ring = nullptr;
  size_t ring_bytes = 0;
  cudaEvent_t* last = nullptr;
  size_t want = chunk_bytes * static_cast<size_t>(nchunks);
  size_t cap = static_cast<size_t>(64) << 20;
  if (const char* e = getenv("CH_LOADER_RING_BYTES")) {       // tests: a small ring, so that slots are refilled
    const long long v = atoll(e);
    if (v > 0) cap = static_cast<size_t>(v);
  }
  if (want > cap) want = cap < 4 * chunk_bytes ? 4 * chunk_bytes : cap;
  if (ch_ws_loader_ring(ws, want, &ring, &ring_bytes, &last)) return 1;
  CH_CUDA(cudaEventSynchronize(*last));        // the previous loader's copies have left the ring
  ch_loader* h = new ch_loader(nchunks);
  Loader& L = h->impl;
  L.ws = ws; L.src = codes_host; L.n = n; L.rs = row_stride; L.rows_pad = ch_padded_rows(n);
  L.ncols = nbit; L.words = words; L.out_dev = out_bits_dev; L.flags_dev = flags_dev; L.st = st;
  L.piece = piece; L.chunk_rows = chunk_rows; L.nchunks = nchunks;
  L.slots = static_cast<int>((want < ring_bytes ? want : ring_bytes) / chunk_bytes);
  if (L.slots > nchunks) L.slots = nchunks;
  L.ring = static_cast<uint32_t*>(ring);
  L.last_event = last;
  L.ev.resize(static_cast<size_t>(nchunks));
  for (int c = 0; c < nchunks; ++c) {
    const int64_t r0 = static_cast<int64_t>(c) * chunk_rows;
    const int64_t r1 = r0 + chunk_rows < n ? r0 + chunk_rows : n;
    L.left[static_cast<size_t>(c)].store(static_cast<int>((r1 - r0 + piece - 1) / piece));
    cudaError_t e = cudaEventCreateWithFlags(&L.ev[static_cast<size_t>(c)], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      for (int d = 0; d < c; ++d) cudaEventDestroy(L.ev[static_cast<size_t>(d)]);
      delete h;
      CH_FAIL("cudaEventCreate failed: %s", cudaGetErrorString(e));
    }
  }
  // the pad rows [n, rows_pad) are zero, as ch_pack_sign leaves them
  cudaError_t e = cudaMemsetAsync(out_bits_dev + static_cast<size_t>(n) * words, 0,
                                  static_cast<size_t>(L.rows_pad - n) * words * 4, st);
  if (e != cudaSuccess) {
    for (auto& v : L.ev) cudaEventDestroy(v);
    delete h;
    CH_FAIL("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
  }
  L.thread = std::thread([h] { h->impl.run(); });
  *out = h;
  return 0;
}

extern "C" int ch_host_loader_wait(ch_loader* h, int64_t rows, void* stream) {
  if (h == nullptr) CH_FAIL("null loader");
  Loader& L = h->impl;
  if (rows <= 0) return 0;
  if (rows > L.n) rows = L.n;
  const int c = static_cast<int>((rows - 1) / L.chunk_rows);
  {
    std::unique_lock<std::mutex> lk(L.mu);
    L.cv.wait(lk, [&] { return L.sent > c || L.failed.load() || L.done; });
    if (L.failed.load()) CH_FAIL("%s", L.err);
    if (L.sent <= c) CH_FAIL("gallery loader stopped early");
  }
  ChDeviceGuard g(L.ws->device);
  CH_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), L.ev[static_cast<size_t>(c)], 0));
  return 0;
}

extern "C" int ch_host_loader_join(ch_loader* h, uint32_t* flags_host) {
  if (h == nullptr) return 0;
  Loader& L = h->impl;
  if (L.thread.joinable()) L.thread.join();
  if (flags_host != nullptr) *flags_host = L.fl.load();
  const bool failed = L.failed.load();
  if (failed) ch_set_error("%s", L.err);
  {
    ChDeviceGuard g(L.ws->device);
    for (auto& v : L.ev) cudaEventDestroy(v);      // (destruction is deferred by the runtime while a stream still waits on it)
  }
  delete h;
  return failed ? 1 : 0;
}
