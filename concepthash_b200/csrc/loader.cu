// Native host loader: the host half of K1 for HOST-resident fp32 codes, run on its own thread from the first moment
// of an evaluation (include/concepthash_b200.h: ch_host_loader_*).
//
// `inference_one_epoch` hands over a torch.cat of `.cpu()` batches (trainers/base.py:291-304).  The host's cores are the
// bottleneck of such an evaluation: they have to read every code once (512 MB for a 1M x 128 gallery, ~3 ms on the
// 16-vCPU B200 host), the GPU needs ~3 ms for the rest and can start on the gallery only once the row sample has
// given it the thresholds.  The evaluation is therefore as fast as (a) the cores start early, (b) never pause, (c) the
// evaluating thread never waits for them (it queues the GPU's work ahead of the data), and (d) little is left to do
// after the last byte.  A loader takes up to CH_LOADER_MAX_JOBS arrays in order (queries, row sample, gallery).  ONE
// pool run covers them all: the threads draw ~256 KB pieces from a shared counter, sign/bit-pack them (AVX-512 /
// AVX2, host_pack.cpp) into a ring of pinned chunk buffers, and thread 0 -- between two pieces of its own -- sends
// every chunk that is complete, in order, with one cudaMemcpyAsync on the caller's side stream, followed by
// an event.  A consumer (ch_host_loader_wait) blocks until the copy of the chunk that holds its last row has been
// queued, then makes ITS stream wait for that event; the queries and the sample are done ~0.3 ms after the start, so
// the evaluating thread waits less than it used to spend packing them itself, and meanwhile queues the GPU's work.
// (Opt-in, CH_LOADER_STREAM_OPS=1: a stream-ordered progress word per job -- cuStreamWriteValue32 behind every
// copy, cuStreamWaitValue32 on the consumer's stream -- so that nothing blocks at all.  It deadlocked on the B200
// box when the two streams shared a hardware queue; see StreamOps.)  The evaluating thread's own host packs
// (labels) are helped by the loader's threads between two pieces (host_pack.cpp: ch_host_help).
#include <cuda.h>
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
void ch_host_copy_piece(void* dst, const void* src, size_t bytes);   // host_pack.cpp: non-temporal stores into pinned memory
int ch_ws_loader_ring(ch_ws* ws, size_t bytes, void** ring, size_t* ring_bytes, cudaEvent_t** last,
                      uint32_t** progress_dev, uint32_t** seq);                                    // api.cu

namespace {

__global__ void loader_or_flags_kernel(uint32_t* __restrict__ flags, uint32_t bits) { atomicOr(flags, bits); }

// stream-ordered memory operations of the driver API (looked up once; the library links the runtime only)
typedef CUresult (*stream_value_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
struct StreamOps {
  stream_value_fn wait = nullptr, write = nullptr;
  StreamOps() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      wait = reinterpret_cast<stream_value_fn>(p);
    p = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      write = reinterpret_cast<stream_value_fn>(p);
    cudaGetLastError();
    if (wait == nullptr || write == nullptr) wait = write = nullptr;
    // OFF unless CH_LOADER_STREAM_OPS=1.  The ordering a stream memory operation creates is invisible to CUDA's
    // scheduler: when the waiting stream and the loader's copy stream are mapped onto the same hardware queue, the
    // wait can sit in front of the very copy it waits for.  Measured on the B200 box: tests green, the 25 000-query
    // evaluation hung.  The default is the event path: the caller blocks until the chunk's copy has been queued.
    const char* e = getenv("CH_LOADER_STREAM_OPS");
    if (e == nullptr || atoi(e) != 1) wait = write = nullptr;
  }
};
const StreamOps& stream_ops() {
  static StreamOps ops;
  return ops;
}

struct Job {
  int kind = CH_LOADER_PACK;               // CH_LOADER_COPY: rows of `words` x 4 bytes are copied as they are
  const float* src = nullptr;
  int64_t n = 0, rs = 0, rows_pad = 0;
  int ncols = 0, words = 0;
  uint32_t* out_dev = nullptr;
  uint32_t* flags_dev = nullptr;
  int64_t piece = 0, chunk_rows = 0, npieces = 0, piece0 = 0;   // piece0: global index of the job's first piece
  int nchunks = 0, chunk0 = 0;                                   // chunk0: global index of the job's first chunk
  uint32_t base = 0;                                             // progress word = base + chunks of this job sent
  std::atomic<uint32_t> fl{0};
};

struct Loader {
  ch_ws* ws = nullptr;
  cudaStream_t st = nullptr;
  int njobs = 0;
  Job job[CH_LOADER_MAX_JOBS];
  int64_t npieces = 0;
  int nchunks = 0, slots = 0;
  size_t slot_words = 0;
  uint32_t* ring = nullptr;
  uint32_t* progress_dev = nullptr;        // (workspace) one word per job slot
  cudaEvent_t* last_event = nullptr;       // (workspace) the last copy that read the ring: waited for by the next loader
  std::vector<cudaEvent_t> ev;             // per chunk, recorded behind its copy
  std::vector<std::atomic<int>> left;      // pieces of the chunk still being packed
  std::atomic<int64_t> next{0};
  std::atomic<int> freed{0};               // chunks whose copy has completed (their ring slot may be refilled)
  std::atomic<bool> failed{false};
  std::mutex mu;
  std::condition_variable cv;
  int sent = 0;                            // chunks whose copy + event have been enqueued (guarded by mu)
  bool done = false;
  char err[512] = "";
  std::thread thread;

  explicit Loader(int chunks) : left(static_cast<size_t>(chunks)) {}

  int job_of_chunk(int c, int* local) const {
    int j = 0;
    while (j + 1 < njobs && c >= job[j + 1].chunk0) ++j;
    *local = c - job[j].chunk0;
    return j;
  }

  void fail(const char* what, const char* why) {
    std::lock_guard<std::mutex> lk(mu);
    if (!failed.load()) snprintf(err, sizeof(err), "host loader: %s: %s", what, why);
    failed.store(true);
    cv.notify_all();
  }

  // thread 0: sends every chunk that has been packed completely (in order) and frees the slots of finished copies
  void drain(int& sent_local) {
    while (sent_local < nchunks && left[static_cast<size_t>(sent_local)].load(std::memory_order_acquire) == 0) {
      const int c = sent_local;
      int lc = 0;
      const int j = job_of_chunk(c, &lc);
      const Job& J = job[j];
      const int64_t r0 = static_cast<int64_t>(lc) * J.chunk_rows;
      const int64_t r1 = r0 + J.chunk_rows < J.n ? r0 + J.chunk_rows : J.n;
      const uint32_t* buf = ring + static_cast<size_t>(c % slots) * slot_words;
      if (lc == J.nchunks - 1) {            // the job's flags are final once its last chunk is packed: they travel first
        const uint32_t f = J.fl.load();
        if (f != 0u && J.flags_dev != nullptr) {
          cudaGetLastError();               // (a cudaEventQuery that answered "not ready" is this thread's last error)
          loader_or_flags_kernel<<<1, 1, 0, st>>>(J.flags_dev, f);
          cudaError_t e = cudaGetLastError();
          if (e != cudaSuccess) return fail("flag kernel", cudaGetErrorString(e));
          __atomic_fetch_add(&ws->launches, 1, __ATOMIC_RELAXED);
        }
      }
      cudaError_t e = cudaMemcpyAsync(J.out_dev + static_cast<size_t>(r0) * J.words, buf,
                                      static_cast<size_t>(r1 - r0) * J.words * 4, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaEventRecord(ev[static_cast<size_t>(c)], st);
      if (e != cudaSuccess) return fail("copy", cudaGetErrorString(e));
      if (stream_ops().write != nullptr) {
        const CUresult r = stream_ops().write(reinterpret_cast<CUstream>(st),
                                              reinterpret_cast<CUdeviceptr>(progress_dev + j),
                                              J.base + static_cast<uint32_t>(lc) + 1u, CU_STREAM_WRITE_VALUE_DEFAULT);
        if (r != CUDA_SUCCESS) return fail("cuStreamWriteValue32", "driver error");
      }
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

  // on failure the waiting streams are released (their kernels read garbage; the error is raised at join)
  void release_waiters() {
    if (stream_ops().write == nullptr) return;
    for (int j = 0; j < njobs; ++j)
      stream_ops().write(reinterpret_cast<CUstream>(st), reinterpret_cast<CUdeviceptr>(progress_dev + j),
                         job[j].base + static_cast<uint32_t>(job[j].nchunks), CU_STREAM_WRITE_VALUE_DEFAULT);
  }

  void run() {
    if (cudaSetDevice(ws->device) != cudaSuccess) {
      fail("cudaSetDevice", cudaGetErrorString(cudaGetLastError()));
    } else {
      int sent_local = 0;
      ch_host_loader_running(+1);
      // (one core is left to the evaluating thread: it queues the GPU's work meanwhile -- with none to spare an
      // evaluation took 7.5 instead of 4.9 ms)
      int spare = 1;
      if (const char* e = getenv("CH_LOADER_SPARE")) spare = atoi(e);
      int nt = ws->pack_threads > 3 ? ws->pack_threads - spare : ws->pack_threads;
      if (nt < 1) nt = 1;
      ch_host_pool_run(1, nt, [&](int t) {
        for (;;) {
          if (t == 0) drain(sent_local);
          if (failed.load(std::memory_order_relaxed)) break;
          ch_host_help();                     // the evaluating thread's short packs go first (they feed the GPU's first work)
          const int64_t p = next.fetch_add(1, std::memory_order_relaxed);
          if (p >= npieces) {
            // no piece left for this thread -- but thread 0 is the one that sends chunks and frees ring slots: it
            // stays until the last chunk is on its way.  (It used to leave here; a thread still waiting for a ring
            // slot then waited for ever, and the pool run with it: a rare hang with a ring of 4 slots.)
            while (t == 0 && sent_local < nchunks && !failed.load(std::memory_order_relaxed)) {
              drain(sent_local);
              ch_host_help();
              if (sent_local < nchunks) std::this_thread::yield();
            }
            break;
          }
          int j = 0;
          while (j + 1 < njobs && p >= job[j + 1].piece0) ++j;
          Job& J = job[j];
          const int64_t a = (p - J.piece0) * J.piece;
          const int lc = static_cast<int>(a / J.chunk_rows);
          const int c = J.chunk0 + lc;
          while (c - slots >= freed.load(std::memory_order_acquire) && !failed.load(std::memory_order_relaxed)) {
            ch_host_help();
            if (t == 0) drain(sent_local);
            else std::this_thread::yield();
          }
          int64_t b = a + J.piece;
          if (b > J.n) b = J.n;
          uint32_t* dst = ring + static_cast<size_t>(c % slots) * slot_words +
                          static_cast<size_t>(a - static_cast<int64_t>(lc) * J.chunk_rows) * J.words;
          if (J.kind == CH_LOADER_COPY) {
            const char* srcb = reinterpret_cast<const char*>(J.src);
            const size_t rowb = static_cast<size_t>(J.words) * 4, pitch = static_cast<size_t>(J.rs);
            if (pitch == rowb) {
              ch_host_copy_piece(dst, srcb + static_cast<size_t>(a) * pitch, static_cast<size_t>(b - a) * rowb);
            } else {
              for (int64_t r = a; r < b; ++r)
                memcpy(reinterpret_cast<char*>(dst) + static_cast<size_t>(r - a) * rowb,
                       srcb + static_cast<size_t>(r) * pitch, rowb);
            }
          } else {
            const uint32_t f = ch_host_pack_rows(J.src, a, b, J.ncols, J.rs, J.words, dst);
            if (f != 0u) J.fl.fetch_or(f, std::memory_order_relaxed);
          }
          left[static_cast<size_t>(c)].fetch_sub(1, std::memory_order_release);
        }
      });
      ch_host_loader_running(-1);
      while (sent_local < nchunks && !failed.load()) {       // pieces still in other threads' hands
        drain(sent_local);
        if (sent_local < nchunks) std::this_thread::yield();
      }
      if (failed.load()) release_waiters();
      if (!failed.load() && last_event != nullptr && nchunks > 0) {
        cudaError_t e = cudaEventRecord(*last_event, st);
        if (e != cudaSuccess) fail("event", cudaGetErrorString(e));
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

extern "C" int ch_host_loader_start(ch_ws* ws, const ch_loader_job* jobs, int njobs, void* stream, ch_loader** out) {
  if (out == nullptr) CH_FAIL("null out pointer");
  *out = nullptr;
  if (ws == nullptr) CH_FAIL("null workspace");
  if (jobs == nullptr || njobs < 1 || njobs > CH_LOADER_MAX_JOBS) CH_FAIL("1..%d loader jobs", CH_LOADER_MAX_JOBS);
  if (ws->pack_threads < 1) CH_FAIL("host packing is disabled (CH_PACK_THREADS=0)");
  ChDeviceGuard g(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Job tmp[CH_LOADER_MAX_JOBS];
  int nchunks = 0;
  int64_t npieces = 0;
  size_t slot_words = 0;
  for (int j = 0; j < njobs; ++j) {
    const ch_loader_job& in = jobs[j];
    Job& J = tmp[j];
    J.kind = in.kind;
    if (in.kind != CH_LOADER_PACK && in.kind != CH_LOADER_COPY) CH_FAIL("loader job %d: unknown kind %d", j, in.kind);
    if (in.kind == CH_LOADER_COPY) {
      // raw rows of `nbit` BYTES (a multiple of 4; e.g. 8 = int64 class ids), pitch `row_stride` bytes
      if (in.nbit < 4 || (in.nbit & 3) != 0 || in.nbit > 4096) CH_FAIL("copy job %d: %d bytes per row", j, in.nbit);
      J.words = in.nbit / 4;
    } else {
      J.words = ch_code_words(in.nbit);
      if (J.words == 0) CH_FAIL("nbit=%d unsupported (1..%d)", in.nbit, CH_MAX_NBIT);
    }
    if (in.n <= 0 || in.codes_host == nullptr || in.out_bits_dev == nullptr || in.row_stride < in.nbit)
      CH_FAIL("bad loader job %d", j);
    if ((reinterpret_cast<uintptr_t>(in.codes_host) & 3) != 0) CH_FAIL("loader sources must be 4-byte aligned");
    J.src = static_cast<const float*>(in.codes_host); J.n = in.n; J.rs = in.row_stride;
    J.rows_pad = in.kind == CH_LOADER_COPY ? in.n : ch_padded_rows(in.n);
    J.ncols = in.nbit; J.out_dev = in.out_bits_dev; J.flags_dev = in.flags_dev;
    // pieces of ~256 KB of codes; chunks (the unit that travels) of 32 pieces -- 8 for the short jobs in front of the
    // last one, whose rows the GPU is waiting for -- and at most ~256 chunks per job
    J.piece = (256 * 1024) / (in.kind == CH_LOADER_COPY ? static_cast<int64_t>(in.nbit)
                                                         : static_cast<int64_t>(in.nbit) * 4);
    if (J.piece < 16) J.piece = 16;
    int64_t per_chunk = (j + 1 < njobs) ? 8 : 32;
    while ((J.n + J.piece * per_chunk - 1) / (J.piece * per_chunk) > 256) per_chunk *= 2;
    J.chunk_rows = J.piece * per_chunk;
    J.nchunks = static_cast<int>((J.n + J.chunk_rows - 1) / J.chunk_rows);
    J.npieces = (J.n + J.piece - 1) / J.piece;
    J.chunk0 = nchunks;
    J.piece0 = npieces;
    nchunks += J.nchunks;
    npieces += J.npieces;
    const size_t cw = static_cast<size_t>(J.chunk_rows) * J.words;
    if (cw > slot_words) slot_words = cw;
  }
  void* ring = nullptr;
  size_t ring_bytes = 0;
  cudaEvent_t* last = nullptr;
  uint32_t* progress = nullptr;
  uint32_t* seq = nullptr;
  const size_t slot_bytes = slot_words * 4;
  size_t want = slot_bytes * static_cast<size_t>(nchunks);
  size_t cap = static_cast<size_t>(64) << 20;
  if (const char* e = getenv("CH_LOADER_RING_BYTES")) {       // tests: a small ring, so that slots are refilled
    const long long v = atoll(e);
    if (v > 0) cap = static_cast<size_t>(v);
  }
  if (want > cap) want = cap < 4 * slot_bytes ? 4 * slot_bytes : cap;
  if (ch_ws_loader_ring(ws, want, &ring, &ring_bytes, &last, &progress, &seq)) return 1;
  CH_CUDA(cudaEventSynchronize(*last));        // the previous loader's copies have left the ring
  ch_loader* h = new ch_loader(nchunks);
  Loader& L = h->impl;
  L.ws = ws; L.st = st; L.njobs = njobs; L.npieces = npieces; L.nchunks = nchunks;
  L.slot_words = slot_words;
  L.slots = static_cast<int>((want < ring_bytes ? want : ring_bytes) / slot_bytes);
  if (L.slots > nchunks) L.slots = nchunks;
  L.ring = static_cast<uint32_t*>(ring);
  L.progress_dev = progress;
  L.last_event = last;
  for (int j = 0; j < njobs; ++j) {
    Job& J = L.job[j];
    const Job& T = tmp[j];
    J.kind = T.kind; J.src = T.src; J.n = T.n; J.rs = T.rs; J.rows_pad = T.rows_pad; J.ncols = T.ncols; J.words = T.words;
    J.out_dev = T.out_dev; J.flags_dev = T.flags_dev; J.piece = T.piece; J.chunk_rows = T.chunk_rows;
    J.npieces = T.npieces; J.piece0 = T.piece0; J.nchunks = T.nchunks; J.chunk0 = T.chunk0;
    J.base = *seq;                       // (monotonic over the workspace's life; compared cyclically)
    *seq += static_cast<uint32_t>(J.nchunks);
    for (int lc = 0; lc < J.nchunks; ++lc) {
      const int64_t r0 = static_cast<int64_t>(lc) * J.chunk_rows;
      const int64_t r1 = r0 + J.chunk_rows < J.n ? r0 + J.chunk_rows : J.n;
      L.left[static_cast<size_t>(J.chunk0 + lc)].store(static_cast<int>((r1 - r0 + J.piece - 1) / J.piece));
    }
  }
  L.ev.resize(static_cast<size_t>(nchunks));
  for (int c = 0; c < nchunks; ++c) {
    cudaError_t e = cudaEventCreateWithFlags(&L.ev[static_cast<size_t>(c)], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      for (int d = 0; d < c; ++d) cudaEventDestroy(L.ev[static_cast<size_t>(d)]);
      delete h;
      CH_FAIL("cudaEventCreate failed: %s", cudaGetErrorString(e));
    }
  }
  // the pad rows [n, rows_pad) are zero, as ch_pack_sign leaves them
  for (int j = 0; j < njobs; ++j) {
    const Job& J = L.job[j];
    if (J.rows_pad == J.n) continue;
    cudaError_t e = cudaMemsetAsync(J.out_dev + static_cast<size_t>(J.n) * J.words, 0,
                                    static_cast<size_t>(J.rows_pad - J.n) * J.words * 4, st);
    if (e != cudaSuccess) {
      for (auto& v : L.ev) cudaEventDestroy(v);
      delete h;
      CH_FAIL("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    }
  }
  L.thread = std::thread([h] { h->impl.run(); });
  *out = h;
  return 0;
}

extern "C" int ch_host_loader_wait(ch_loader* h, int job, int64_t rows, void* stream, int block) {
  if (h == nullptr) CH_FAIL("null loader");
  Loader& L = h->impl;
  if (job < 0 || job >= L.njobs) CH_FAIL("loader job %d out of range", job);
  const Job& J = L.job[job];
  if (rows <= 0) return 0;
  if (rows > J.n) rows = J.n;
  const int lc = static_cast<int>((rows - 1) / J.chunk_rows);
  const bool ops = stream_ops().wait != nullptr;
  if (block || !ops) {
    std::unique_lock<std::mutex> lk(L.mu);
    L.cv.wait(lk, [&] { return L.sent > J.chunk0 + lc || L.failed.load() || L.done; });
    if (L.failed.load()) CH_FAIL("%s", L.err);
    if (L.sent <= J.chunk0 + lc) CH_FAIL("host loader stopped early");
  }
  ChDeviceGuard g(L.ws->device);
  if (ops) {
    // `stream` proceeds once the progress word says "chunk lc of this job has been copied" -- queued NOW, whether
    // or not the chunk exists yet
    const CUresult r = stream_ops().wait(static_cast<CUstream>(stream),
                                         reinterpret_cast<CUdeviceptr>(L.progress_dev + job),
                                         J.base + static_cast<uint32_t>(lc) + 1u, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) CH_FAIL("cuStreamWaitValue32 failed (%d)", static_cast<int>(r));
  } else {
    CH_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), L.ev[static_cast<size_t>(J.chunk0 + lc)], 0));
  }
  return 0;
}

extern "C" int ch_host_loader_join(ch_loader* h, uint32_t* flags_host) {
  if (h == nullptr) return 0;
  Loader& L = h->impl;
  if (L.thread.joinable()) L.thread.join();
  if (flags_host != nullptr)
    for (int j = 0; j < L.njobs; ++j) flags_host[j] = L.job[j].fl.load();
  const bool failed = L.failed.load();
  if (failed) ch_set_error("%s", L.err);
  {
    ChDeviceGuard g(L.ws->device);
    for (auto& v : L.ev) cudaEventDestroy(v);      // (destruction is deferred by the runtime while a stream still waits on it)
  }
  delete h;
  return failed ? 1 : 0;
}
