// Workspace, error reporting and device queries of the C-ABI (include/concepthash_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include <thread>

#include "common.cuh"

#include <vector>

static thread_local char g_err[1024] = "";

void ch_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct ch_ws_priv {
  ch_ws pub;
  void* scratch;
  size_t scratch_bytes;
  std::vector<void*> retired;   // outgrown scratch buffers: kept until destroy (a captured CUDA graph may still name them)
  void* res_dev;                // results of ch_reduce_means_enqueue: device words and their pinned host mirror
  void* res_host;
  void* ring;                   // pinned chunk ring of the gallery loader (loader.cu) + the event behind its last copy
  size_t ring_bytes;
  cudaEvent_t ring_last;
  bool ring_event;
  uint32_t* progress_dev;       // progress words of the loader's jobs (stream-ordered writes / waits) + their sequence
  uint32_t progress_seq;
};

// grow-only device scratch shared by the small helper kernels (all calls are stream-ordered by the caller)
int ch_ws_scratch(ch_ws* ws, size_t bytes, void** out) {
  ch_ws_priv* p = reinterpret_cast<ch_ws_priv*>(ws);
  if (bytes > p->scratch_bytes) {
    if (p->scratch) p->retired.push_back(p->scratch);
    p->scratch = nullptr;
    p->scratch_bytes = 0;
    const size_t want = ch_round_up(static_cast<int64_t>(bytes), 1 << 20);
    CH_CUDA(cudaMalloc(&p->scratch, want));
    p->scratch_bytes = want;
  }
  *out = p->scratch;
  return 0;
}

// 1 KB of device words + a pinned host mirror for the results of an evaluation (allocated on first use)
int ch_ws_results(ch_ws* ws, void** dev, void** host) {
  ch_ws_priv* p = reinterpret_cast<ch_ws_priv*>(ws);
  if (p->res_dev == nullptr) CH_CUDA(cudaMalloc(&p->res_dev, 1024));
  if (p->res_host == nullptr) CH_CUDA(cudaHostAlloc(&p->res_host, 1024, cudaHostAllocDefault));
  *dev = p->res_dev;
  *host = p->res_host;
  return 0;
}

// pinned ring of the gallery loader: grow-only, at least `bytes`; *last = the event recorded behind the last copy that
// read the ring (never recorded at first: synchronising on it returns at once)
int ch_ws_loader_ring(ch_ws* ws, size_t bytes, void** ring, size_t* ring_bytes, cudaEvent_t** last,
                      uint32_t** progress_dev, uint32_t** seq) {
  ch_ws_priv* p = reinterpret_cast<ch_ws_priv*>(ws);
  if (!p->ring_event) {
    CH_CUDA(cudaEventCreateWithFlags(&p->ring_last, cudaEventDisableTiming));
    p->ring_event = true;
  }
  if (p->progress_dev == nullptr) {
    CH_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->progress_dev), 64));
    CH_CUDA(cudaMemset(p->progress_dev, 0, 64));
    CH_CUDA(cudaDeviceSynchronize());
  }
  *progress_dev = p->progress_dev;
  *seq = &p->progress_seq;
  if (bytes > p->ring_bytes) {
    if (p->ring != nullptr) {
      CH_CUDA(cudaEventSynchronize(p->ring_last));
      CH_CUDA(cudaFreeHost(p->ring));
      p->ring = nullptr;
      p->ring_bytes = 0;
    }
    const size_t want = ch_round_up(static_cast<int64_t>(bytes), 1 << 20);
    CH_CUDA(cudaHostAlloc(&p->ring, want, cudaHostAllocDefault));
    p->ring_bytes = want;
  }
  *ring = p->ring;
  *ring_bytes = p->ring_bytes;
  *last = &p->ring_last;
  return 0;
}

extern "C" int ch_abi_version(void) { return CH_ABI_VERSION; }
extern "C" const char* ch_last_error(void) { return g_err; }

extern "C" int ch_workspace_create(int device, ch_ws** out) {
  if (out == nullptr) CH_FAIL("null out pointer");
  *out = nullptr;
  int count = 0;
  CH_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) CH_FAIL("device %d out of range (found %d CUDA devices)", device, count);
  cudaDeviceProp prop;
  CH_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    CH_FAIL("concepthash_b200 is built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major, prop.minor,
            prop.name);
  ChDeviceGuard guard(device);
  ch_ws_priv* p = new ch_ws_priv();   // value-initialised: every plain member is zero
  ch_ws* ws = &p->pub;
  ws->device = device;
  ws->sm_count = prop.multiProcessorCount;
  ws->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  ws->l2_bytes = prop.l2CacheSize;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  ws->clock_khz = khz;
  ws->stage_bytes = static_cast<size_t>(64) << 20;
  {
    unsigned hc = std::thread::hardware_concurrency();
    ws->host_threads = hc >= 16 ? 8 : (hc >= 4 ? static_cast<int>(hc / 2) : 1);
    if (const char* e = getenv("CH_HOST_THREADS")) {          // tuning knob for the pageable-source copies
      const int v = atoi(e);
      if (v >= 1 && v <= 64) ws->host_threads = v;
    }
    ws->pack_threads = hc >= 2 ? static_cast<int>(hc > 32 ? 32 : hc) : 1;   // read-only pass: every core helps
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {         // one process per GPU (torchrun): the ranks share the cores
      const int lw = atoi(e);
      if (lw > 1 && hc >= 2) {
        int v = static_cast<int>(hc) / lw;
        ws->pack_threads = v < 2 ? 2 : (v > 32 ? 32 : v);
      }
    }
    if (const char* e = getenv("CH_PACK_THREADS")) {          // 0 = keep the bounce-copy + GPU pack path
      const int v = atoi(e);
      if (v >= 0 && v <= 128) ws->pack_threads = v;
    }
  }
  for (int i = 0; i < 2; ++i) {
    ws->stage[i] = nullptr;
    ws->bounce[i] = nullptr;
    CH_CUDA(cudaEventCreateWithFlags(&ws->ev_copied[i], cudaEventDisableTiming));
    CH_CUDA(cudaEventCreateWithFlags(&ws->ev_consumed[i], cudaEventDisableTiming));
  }
  CH_CUDA(cudaStreamCreateWithFlags(&ws->copy_stream, cudaStreamNonBlocking));
  *out = ws;
  return 0;
}

// staging buffers are allocated on first host-side use
int ch_ws_ensure_stage(ch_ws* ws) {
  for (int i = 0; i < 2; ++i)
    if (ws->stage[i] == nullptr) CH_CUDA(cudaMalloc(&ws->stage[i], ws->stage_bytes));
  return 0;
}

int ch_ws_ensure_bounce(ch_ws* ws) {
  for (int i = 0; i < 2; ++i)
    if (ws->bounce[i] == nullptr) CH_CUDA(cudaHostAlloc(&ws->bounce[i], ws->stage_bytes, cudaHostAllocDefault));
  return 0;
}

extern "C" int ch_workspace_destroy(ch_ws* ws) {
  if (ws == nullptr) return 0;
  ch_ws_priv* p = reinterpret_cast<ch_ws_priv*>(ws);
  ChDeviceGuard guard(ws->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < 2; ++i) {
    if (ws->stage[i]) cudaFree(ws->stage[i]);
    if (ws->bounce[i]) cudaFreeHost(ws->bounce[i]);
    cudaEventDestroy(ws->ev_copied[i]);
    cudaEventDestroy(ws->ev_consumed[i]);
  }
  cudaStreamDestroy(ws->copy_stream);
  if (p->scratch) cudaFree(p->scratch);
  for (void* r : p->retired) cudaFree(r);
  if (p->res_dev) cudaFree(p->res_dev);
  if (p->res_host) cudaFreeHost(p->res_host);
  if (p->ring) cudaFreeHost(p->ring);
  if (p->ring_event) cudaEventDestroy(p->ring_last);
  if (p->progress_dev) cudaFree(p->progress_dev);
  delete p;
  return 0;
}

extern "C" int ch_device_info(ch_ws* ws, int* sm_count, int* max_smem_optin, int* l2_bytes, int* clock_khz) {
  if (ws == nullptr) CH_FAIL("null workspace");
  if (sm_count) *sm_count = ws->sm_count;
  if (max_smem_optin) *max_smem_optin = ws->max_smem_optin;
  if (l2_bytes) *l2_bytes = ws->l2_bytes;
  if (clock_khz) *clock_khz = ws->clock_khz;
  return 0;
}

extern "C" int64_t ch_launch_count(ch_ws* ws) { return ws ? ws->launches : 0; }
