// K2: Hamming distances (XOR + POPC on the integer pipes) fused with the per-query key histogram and
// the stable in-bucket prefixes the exact top-R selection / AP need.  Nothing of size (nq, ndb) is
// ever written.
//
// Mapping.  One THREAD owns one query (its packed code stays in registers); one CTA owns `threads`
// queries x one gallery stripe (a contiguous block of gallery rows).  The stripe is streamed through
// shared memory in tiles by 1-D bulk async copies (cp.async.bulk + mbarrier, the TMA engine; SASS
// UBLKCP), double-buffered; every thread reads the same gallery word at the same time (shared-memory
// broadcast) and walks the tile IN ROW ORDER.  The thread's private histogram lives in shared memory,
// laid out [key][thread] so that lane l always hits bank l (conflict-free without atomics); one u32
// per key holds {lo16: items seen, hi16: relevant items seen}.  Because the walk is in row order the
// value read back BEFORE the increment is exactly the stable in-bucket prefix
//     p(j) = #{k < j in this stripe : key_k == key_j}
// which is what "ties by ascending gallery row index" needs.  For the rare relevant (or candidate)
// pair the thread appends a 16-byte record {key, p_all, p_rel, row} to its own pre-sized slice of the
// record buffer: no atomics and no divergence-heavy code in the loop.  Counters are flushed into the
// (stripe, key, query) slabs in global memory before any u16 can overflow.
//
// Per pair (64-bit codes): 2 LOP3 + 2 POPC + 1 IADD3 (distance), 1 ISETP + 1 SEL (relevance), 1 IMAD
// (address), LDS + IADD + STS (histogram) and 1/4 LDS.128 x2 (gallery code + label broadcast).
// The limiter is the POPC pipe (16 lanes / clk / SM): algorithmic popc32 ops per pair = words.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kStages = 2;

// compile-time tile geometry of the code / id stages (static shared memory: the compiler then KNOWS the
// gallery tiles do not alias the histogram, so tile loads are hoisted above histogram stores)
template <int NW, bool TERN>
struct TileCfg {
  static constexpr int kRows = (NW <= 2 ? 256 : (NW == 4 ? 128 : 64)) / (TERN ? 2 : 1);
  static constexpr int kGroup = (NW * (TERN ? 2 : 1) >= 8) ? 2 : 4;   // rows per register group
  static constexpr int kPad = 2 * kGroup;                             // rows the prefetch may over-read
};

__host__ __device__ inline int tile_rows_for(int nw, bool tern, int label_mode, int lw) {
  int t = nw <= 2 ? 256 : (nw == 4 ? 128 : 64);
  if (tern) t /= 2;
  if (label_mode == CH_LAB_MASK)
    while (t > 32 && t * lw * 4 > 4096) t /= 2;
  return t;
}

struct SmemPlan {
  size_t hist, qmask, stage_mask, total_dynamic, total_static, total;
};

// dynamic part: histogram, query masks, gallery mask stages (CH_LAB_MASK only);
// static part: code / non-zero / id stages + barriers (see the kernel)
inline SmemPlan smem_plan(int nbins, int threads, int nw, bool tern, int label_mode, int lw, int tile) {
  SmemPlan p;
  p.hist = (static_cast<size_t>(nbins) * threads * 4 + 15) / 16 * 16;
  p.qmask = label_mode == CH_LAB_MASK ? (static_cast<size_t>(lw) * threads * 4 + 15) / 16 * 16 : 0;
  p.stage_mask = label_mode == CH_LAB_MASK ? static_cast<size_t>(tile) * lw * 4 : 0;
  p.total_dynamic = p.hist + p.qmask + kStages * p.stage_mask;
  const int rows = (nw <= 2 ? 256 : (nw == 4 ? 128 : 64)) / (tern ? 2 : 1);
  const int pad = 2 * ((nw * (tern ? 2 : 1) >= 8) ? 2 : 4);
  p.total_static = static_cast<size_t>(kStages) * (rows + pad) * nw * 4 * (tern ? 2 : 1) +
                   (label_mode == CH_LAB_ID ? static_cast<size_t>(kStages) * (rows + pad) * 4 : 0) + 64;
  p.total = p.total_dynamic + p.total_static;
  return p;
}

template <int NW, bool TERN>
__device__ __forceinline__ uint32_t pair_key(const uint32_t (&qb)[NW], const uint32_t (&qz)[NW],
                                             const uint32_t* gb, const uint32_t* gz, int nbit) {
  if constexpr (!TERN) {
    uint32_t d = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) d += __popc(qb[w] ^ gb[w]);
    return d;
  } else {
    // sign(0) = 0: only positions non-zero on both sides contribute +-1 to the inner product
    // key = 2 * dist = nbit - popc(M) + 2 * popc((sq ^ sg) & M),  M = nzq & nzg
    uint32_t both = 0, dis = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const uint32_t m = qz[w] & gz[w];
      both += __popc(m);
      dis += __popc((qb[w] ^ gb[w]) & m);
    }
    return static_cast<uint32_t>(nbit) - both + 2u * dis;
  }
}

// G consecutive gallery rows held in registers
template <int NW, bool TERN, int LAB, int G>
struct RowGroup {
  uint32_t g[G * NW];
  uint32_t z[TERN ? G * NW : 1];
  uint32_t lab[LAB == CH_LAB_ID ? G : 1];
  // i is a multiple of G; G * NW is a multiple of 4 -> all loads are aligned 128-bit broadcasts
  __device__ __forceinline__ void load(const uint32_t* gb, const uint32_t* gz, const uint32_t* gl, int i) {
#pragma unroll
    for (int v = 0; v < G * NW / 4; ++v) {
      const uint4 x = *reinterpret_cast<const uint4*>(gb + i * NW + 4 * v);
      g[4 * v] = x.x; g[4 * v + 1] = x.y; g[4 * v + 2] = x.z; g[4 * v + 3] = x.w;
      if constexpr (TERN) {
        const uint4 y = *reinterpret_cast<const uint4*>(gz + i * NW + 4 * v);
        z[4 * v] = y.x; z[4 * v + 1] = y.y; z[4 * v + 2] = y.z; z[4 * v + 3] = y.w;
      }
    }
    if constexpr (LAB == CH_LAB_ID) {
      if constexpr (G == 4) {
        const uint4 x = *reinterpret_cast<const uint4*>(gl + i);
        lab[0] = x.x; lab[1] = x.y; lab[2] = x.z; lab[3] = x.w;
      } else {
        const uint2 x = *reinterpret_cast<const uint2*>(gl + i);
        lab[0] = x.x; lab[1] = x.y;
      }
    }
  }
};

template <int NW, bool TERN, int LAB, bool THRESH>
__global__ void __launch_bounds__(256, (NW <= 2 && !TERN) ? 3 : 1) hamming_hist_kernel(const HistDev a) {
  typedef TileCfg<NW, TERN> Cfg;
  constexpr int G = Cfg::kGroup;
  constexpr int kStageRows = Cfg::kRows + Cfg::kPad;
  // static: gallery code / non-zero / id stages and their barriers
  __shared__ __align__(16) uint32_t s_bits[kStages][kStageRows * NW];
  __shared__ __align__(16) uint32_t s_nz[TERN ? kStages : 1][TERN ? kStageRows * NW : 4];
  __shared__ __align__(16) uint32_t s_ids[LAB == CH_LAB_ID ? kStages : 1][LAB == CH_LAB_ID ? kStageRows : 4];
  __shared__ __align__(8) uint64_t bars[kStages];
  // dynamic: histogram [key][thread], query masks, gallery mask stages
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const int TILE = a.tile_rows;                      // <= Cfg::kRows (smaller only for wide label masks)
  unsigned char* hist_b = smem_raw + tid * 4;        // this thread's column, byte-addressed
  const uint32_t T4 = static_cast<uint32_t>(T) * 4u;
  size_t off = (static_cast<size_t>(a.nbins) * T * 4 + 15) / 16 * 16;
  uint32_t* qmask = reinterpret_cast<uint32_t*>(smem_raw + off);
  if (LAB == CH_LAB_MASK) off += (static_cast<size_t>(a.lw) * T * 4 + 15) / 16 * 16;
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem_raw + off);   // [kStages][TILE * lw]
  const int lw = LAB == CH_LAB_MASK ? a.lw : 0;

  const int qtile = blockIdx.x % a.nqtiles;
  const int stripe = blockIdx.x / a.nqtiles;
  const long long q = static_cast<long long>(qtile) * T + tid;
  const bool active = q < a.nq;

  // ---- this thread's query ----
  uint32_t qb[NW], qz[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    qb[w] = active ? a.q_bits[q * NW + w] : 0u;
    qz[w] = (TERN && active) ? a.q_nz[q * NW + w] : 0u;
  }
  uint32_t qid = CH_QUERY_NOLABEL;
  if (LAB == CH_LAB_ID && active) qid = a.q_lab[q];
  if (LAB == CH_LAB_MASK)
    for (int w = 0; w < lw; ++w) qmask[w * T + tid] = active ? a.q_lab[q * lw + w] : 0u;
  uint32_t thr = 0;
  // nbins may have been narrowed from a HINT of the largest threshold (key_limit): a threshold beyond it is flagged
  // by exscan_check_kernel (status bit 3) and the evaluation is repeated -- this launch must only stay in bounds
  if (THRESH && active) thr = min(a.thresh[q], static_cast<uint32_t>(a.nbins - 1));

  for (int b = 0; b < a.nbins; ++b) *reinterpret_cast<uint32_t*>(hist_b + b * T4) = 0u;

  const size_t sq = static_cast<size_t>(stripe) * a.nq_pad + q;  // (stripe, query) slot
  uint32_t rptr = 0, rstart = 0, rend = 0;
  const int emit_mode = active ? a.emit_mode : CH_EMIT_NONE;
  if (emit_mode != CH_EMIT_NONE) {
    rstart = a.rec_off[sq];
    rend = rstart + a.rec_cap[sq];
    rptr = rstart;
  }
  bool overflow = false;

  // ---- stripe geometry ----
  const long long row_begin = static_cast<long long>(stripe) * a.rows_per_stripe;
  long long row_end = row_begin + a.rows_per_stripe;
  if (row_end > a.ndb) row_end = a.ndb;
  const int ntiles = row_end > row_begin ? static_cast<int>((row_end - row_begin + TILE - 1) / TILE) : 0;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](int k) {
    const long long r0 = row_begin + static_cast<long long>(k) * TILE;
    long long rows = row_end - r0;
    if (rows > TILE) rows = TILE;
    const uint32_t rows4 = static_cast<uint32_t>((rows + 3) & ~3ll);  // pad rows exist (ch_padded_rows)
    const int s = k % kStages;
    uint64_t* bar = &bars[s];
    const uint32_t bytes_bits = rows4 * NW * 4u;
    const uint32_t bytes_ids = LAB == CH_LAB_ID ? rows4 * 4u : 0u;
    const uint32_t bytes_mask = rows4 * static_cast<uint32_t>(lw) * 4u;
    mbar_arrive_expect_tx(bar, bytes_bits * (TERN ? 2u : 1u) + bytes_ids + bytes_mask);
    bulk_g2s(&s_bits[s][0], a.g_bits + r0 * NW, bytes_bits, bar);
    if constexpr (TERN) bulk_g2s(&s_nz[s][0], a.g_nz + r0 * NW, bytes_bits, bar);
    if constexpr (LAB == CH_LAB_ID) bulk_g2s(&s_ids[s][0], a.g_lab + r0, bytes_ids, bar);
    if constexpr (LAB == CH_LAB_MASK)
      bulk_g2s(s_mask + static_cast<size_t>(s) * TILE * lw, a.g_lab + r0 * lw, bytes_mask, bar);
  };

  uint32_t epoch = 0;  // number of flushes so far (uniform over the CTA)
  auto flush = [&]() {
    if (epoch == 0u) {
      // first (for stripes of <= 65535 rows: only) flush: nobody else ever writes this (stripe, key, query) entry
      // and the caller handed the slabs over zeroed, so plain stores do -- the read-modify-write form below is a
      // chain of dependent global round trips (one per occupied key), which made the flush of a short stripe
      // cost as much as its pair loop
#pragma unroll 4
      for (int b = 0; b < a.nbins; ++b) {
        uint32_t* h = reinterpret_cast<uint32_t*>(hist_b + b * T4);
        const uint32_t v = *h;
        *h = 0u;
        if (active && v != 0u) {
          const size_t o = (static_cast<size_t>(stripe) * a.nbins + b) * a.nq_pad + q;
          a.slab_all[o] = v & 0xffffu;
          if (LAB != CH_LAB_NONE) a.slab_rel[o] = v >> 16;
        }
      }
    } else {
      for (int b = 0; b < a.nbins; ++b) {
        uint32_t* h = reinterpret_cast<uint32_t*>(hist_b + b * T4);
        const uint32_t v = *h;
        if (v != 0u) {
          *h = 0u;
          if (active) {
            const size_t o = (static_cast<size_t>(stripe) * a.nbins + b) * a.nq_pad + q;
            a.slab_all[o] += v & 0xffffu;
            if (LAB != CH_LAB_NONE) a.slab_rel[o] += v >> 16;
          }
        }
      }
    }
    ++epoch;
  };

  // rare path: append one record to this thread's own slice
  auto emit = [&](uint32_t key, bool rel, uint32_t old, uint32_t shard_row) {
    uint32_t base_all = 0, base_rel = 0;
    if (epoch != 0u) {  // counts already flushed to the slab belong to the in-stripe prefix as well
      const size_t o = (static_cast<size_t>(stripe) * a.nbins + key) * a.nq_pad + q;
      base_all = a.slab_all[o];
      if (LAB != CH_LAB_NONE) base_rel = a.slab_rel[o];
    }
    if (rptr < rend)
      a.recs[rptr] = make_uint4(key | (rel ? 0x80000000u : 0u), base_all + (old & 0xffffu), base_rel + (old >> 16),
                                shard_row);
    else
      overflow = true;
    ++rptr;
  };

  auto relevant = [&](uint32_t glab, const uint32_t* gmask) -> bool {
    if constexpr (LAB == CH_LAB_ID) return glab == qid;
    if constexpr (LAB == CH_LAB_MASK) {
      uint32_t any = 0;
      for (int w = 0; w < lw; ++w) any |= qmask[w * T + tid] & gmask[w];
      return any != 0u;
    }
    return false;
  };

  // N consecutive rows starting at tile row i
  auto process = [&](const uint32_t* g, const uint32_t* z, const uint32_t* lab, const uint32_t* gmask_tile, int i,
                     uint32_t shard_row, auto n_tag) {
    constexpr int N = decltype(n_tag)::value;
    uint32_t key[N];
#pragma unroll
    for (int j = 0; j < N; ++j) key[j] = pair_key<NW, TERN>(qb, qz, g + j * NW, z + (TERN ? j * NW : 0), a.nbit);
    if constexpr (!THRESH) {
      uint32_t old[N];
      bool rel[N];
      bool any = emit_mode == CH_EMIT_CANDIDATES;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        rel[j] = relevant(LAB == CH_LAB_ID ? lab[j] : 0u, gmask_tile + static_cast<size_t>(i + j) * lw);
        uint32_t* h = reinterpret_cast<uint32_t*>(hist_b + key[j] * T4);
        old[j] = *h;
        *h = old[j] + (rel[j] ? 0x10001u : 1u);
        any |= rel[j];
      }
      if (emit_mode != CH_EMIT_NONE && any) {
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (emit_mode == CH_EMIT_CANDIDATES || rel[j]) emit(key[j], rel[j], old[j], shard_row + j);
      }
    } else {
      bool any = false;
#pragma unroll
      for (int j = 0; j < N; ++j) any |= key[j] <= thr;
      if (any) {  // candidates are rare: everything below is off the hot path
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (key[j] <= thr) {
            const bool rel = relevant(LAB == CH_LAB_ID ? lab[j] : 0u, gmask_tile + static_cast<size_t>(i + j) * lw);
            uint32_t* h = reinterpret_cast<uint32_t*>(hist_b + key[j] * T4);
            const uint32_t old = *h;
            *h = old + (rel ? 0x10001u : 1u);
            if (emit_mode == CH_EMIT_CANDIDATES || (emit_mode == CH_EMIT_RELEVANT && rel))
              emit(key[j], rel, old, shard_row + j);
          }
      }
    }
  };

  if (ntiles > 0 && tid == 0) issue(0);
  for (int k = 0; k < ntiles; ++k) {
    if (tid == 0 && k + 1 < ntiles) issue(k + 1);  // slot (k+1)%2 was released by the barrier ending tile k-1
    const int s = k % kStages;
    mbar_wait(&bars[s], static_cast<uint32_t>((k / kStages) & 1));

    const uint32_t* gb = &s_bits[s][0];
    const uint32_t* gz = TERN ? &s_nz[s][0] : &s_nz[0][0];
    const uint32_t* gl = LAB == CH_LAB_ID ? &s_ids[s][0] : &s_ids[0][0];
    const uint32_t* gm = s_mask + static_cast<size_t>(s) * TILE * lw;
    long long rows_ll = row_end - (row_begin + static_cast<long long>(k) * TILE);
    const int rows = rows_ll > TILE ? TILE : static_cast<int>(rows_ll);
    const uint32_t shard_row0 = static_cast<uint32_t>(a.row_base + row_begin) + static_cast<uint32_t>(k) * TILE;

    // register double-buffered walk over groups of G rows (prefetch may over-read into the kPad rows)
    const int nfull = rows - rows % G;
    RowGroup<NW, TERN, LAB, G> A, B;
    int i = 0;
    if (nfull > 0) A.load(gb, gz, gl, 0);
    for (; i < nfull; i += 2 * G) {
      B.load(gb, gz, gl, i + G);
      process(A.g, A.z, A.lab, gm, i, shard_row0 + i, std::integral_constant<int, G>());
      A.load(gb, gz, gl, i + 2 * G);
      if (i + G < nfull) process(B.g, B.z, B.lab, gm, i + G, shard_row0 + i + G, std::integral_constant<int, G>());
    }
    // tail rows of the stripe's last tile
    for (i = nfull; i < rows; ++i) {
      uint32_t g1[NW], z1[TERN ? NW : 1], lab1[1] = {0u};
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        g1[w] = gb[i * NW + w];
        if constexpr (TERN) z1[w] = gz[i * NW + w];
      }
      if constexpr (LAB == CH_LAB_ID) lab1[0] = gl[i];
      process(g1, z1, lab1, gm, i, shard_row0 + i, std::integral_constant<int, 1>());
    }
    __syncthreads();  // everyone is done with this stage before it is refilled
    if ((k + 1) % a.flush_tiles == 0 && k + 1 < ntiles) flush();
  }
  flush();

  if (a.emit_mode != CH_EMIT_NONE && active) {
    a.rec_cnt[sq] = (rptr < rend ? rptr : rend) - rstart;
    if (overflow) atomicOr(a.err_flag, 1u);
  }
}

// ---- slab reductions --------------------------------------------------------------------------------
__global__ void slab_totals_kernel(const uint32_t* __restrict__ slab, int nstripes, long long plane,
                                   uint32_t* __restrict__ tot) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= plane) return;
  uint32_t s = 0;
  for (int k = 0; k < nstripes; ++k) s += slab[static_cast<size_t>(k) * plane + i];
  tot[i] = s;
}
__global__ void slab_exscan_kernel(uint32_t* __restrict__ slab, int nstripes, long long plane) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= plane) return;
  uint32_t run = 0;
  for (int k = 0; k < nstripes; ++k) {
    const size_t o = static_cast<size_t>(k) * plane + i;
    const uint32_t v = slab[o];
    slab[o] = run;
    run += v;
  }
}

// Both slabs of an evaluation in ONE launch: slab (nslabs, nstripes, plane) -> in-place exclusive scan over the
// stripes (= the stripe bases) and the per-rank totals tot (nslabs, plane).  One thread per (slab, key, query);
// the stripe values are loaded in independent batches of 8 (the in-place store after every load had made the
// one-stripe-at-a-time loop a chain of dependent memory round trips).
__global__ void slab_scan_kernel(uint32_t* __restrict__ slab, int nslabs, int nstripes, long long plane,
                                 uint32_t* __restrict__ tot) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= plane * nslabs) return;
  const long long which = i / plane, e = i - which * plane;
  uint32_t* base = slab + static_cast<size_t>(which) * nstripes * plane + e;
  uint32_t run = 0;
  constexpr int B = 8;
  for (int k0 = 0; k0 < nstripes; k0 += B) {
    uint32_t v[B];
#pragma unroll
    for (int j = 0; j < B; ++j) v[j] = (k0 + j < nstripes) ? base[static_cast<size_t>(k0 + j) * plane] : 0u;
#pragma unroll
    for (int j = 0; j < B; ++j) {
      if (k0 + j < nstripes) base[static_cast<size_t>(k0 + j) * plane] = run;
      run += v[j];
    }
  }
  tot[i] = run;
}

// per-stripe class histogram of single-label gallery ids
__global__ void class_counts_kernel(const uint32_t* __restrict__ ids, long long ndb, int rows_per_stripe,
                                    int nclass, uint32_t* __restrict__ cls) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= ndb) return;
  const uint32_t id = ids[r];
  if (id < static_cast<uint32_t>(nclass)) atomicAdd(cls + static_cast<size_t>(r / rows_per_stripe) * nclass + id, 1u);
}

// ---- host side ---------------------------------------------------------------------------------------
typedef void (*hist_fn_t)(const HistDev);

template <int NW, bool TERN, int LAB>
hist_fn_t pick_thresh(bool thresh) {
  return thresh ? hamming_hist_kernel<NW, TERN, LAB, true> : hamming_hist_kernel<NW, TERN, LAB, false>;
}
template <int NW, bool TERN>
hist_fn_t pick_lab(int lab, bool thresh) {
  switch (lab) {
    case CH_LAB_NONE: return pick_thresh<NW, TERN, CH_LAB_NONE>(thresh);
    case CH_LAB_ID: return pick_thresh<NW, TERN, CH_LAB_ID>(thresh);
    default: return pick_thresh<NW, TERN, CH_LAB_MASK>(thresh);
  }
}
template <int NW>
hist_fn_t pick_tern(bool tern, int lab, bool thresh) {
  return tern ? pick_lab<NW, true>(lab, thresh) : pick_lab<NW, false>(lab, thresh);
}
hist_fn_t pick_kernel(int nw, bool tern, int lab, bool thresh) {
  switch (nw) {
    case 1: return pick_tern<1>(tern, lab, thresh);
    case 2: return pick_tern<2>(tern, lab, thresh);
    case 4: return pick_tern<4>(tern, lab, thresh);
    case 8: return pick_tern<8>(tern, lab, thresh);
    default: return nullptr;
  }
}

int pick_threads(int nbins, long long nq) {
  int t = 256;
  while (t > 32 && static_cast<size_t>(nbins) * t * 4 > 68 * 1024) t /= 2;
  while (t > 32 && nq <= t / 2) t /= 2;  // few queries: smaller CTAs, more stripes
  return t;
}

struct Geometry {
  int threads, nstripes, rows_per_stripe, tile, flush_tiles, nqtiles;
  long long nq_pad;
  size_t smem, smem_total;
};

// nq_pad: whole CTAs, and -- unless the query set is tiny -- whole 128-query tiles, so that the tensor-core select
// pass (128 TMEM lanes = 128 queries) can share the arrays of the POPC passes whatever CTA size those use
long long padded_queries(long long nq, int threads) {
  const long long n = nq > 0 ? nq : 1;
  const int unit = (n > 64 && threads < 128) ? 128 : threads;
  return ch_round_up(n, unit);
}

// Modelled time of one full-ranking pass for a candidate (threads, stripes): the pair loop (wave-quantised, scaled
// by the occupancy the CTA size allows) + the slab traffic every stripe costs (a (nbins x nq_pad) plane per slab
// that is zeroed, flushed by its CTAs, summed and scanned) + a fixed cost per CTA wave.
double model_cost(ch_ws* ws, long long ndb, int nw, int nbins, int threads, long long per_sm, long long nq_pad,
                  long long nstripes, long long rps) {
  const long long nqtiles = nq_pad / threads;
  const long long ctas = nstripes * nqtiles;
  const long long slots = per_sm * ws->sm_count;
  const long long waves = (ctas + slots - 1) / slots;
  long long resident = (ctas + static_cast<long long>(ws->sm_count) * waves - 1) / (ws->sm_count * waves);
  if (resident > per_sm) resident = per_sm;
  if (resident < 1) resident = 1;
  double occ = static_cast<double>(resident * threads / 32) / 16.0;       // warps per SM the POPC loop needs: ~16
  if (occ > 1.0) occ = 1.0;
  const double rate_sm = 0.87 * 16.0 * (ws->clock_khz > 0 ? ws->clock_khz * 1e3 : 1.9e9) / nw;   // pairs / s / SM
  const double t_pairs = static_cast<double>(waves) * resident * rps * threads / (rate_sm * occ);
  const double t_slab = static_cast<double>(nstripes) * nbins * nq_pad * 4.0 * 2.0 * 4.0 / 2.5e12;
  (void)ndb;
  return t_pairs + t_slab + 3e-6 * waves;
}

int make_geometry(ch_ws* ws, long long nq, long long ndb, int nbit, bool tern, int lab, int lw, int forced_threads,
                  int forced_stripes, int forced_rows_per_stripe, Geometry* g) {
  const int nw = ch_code_words(nbit);
  if (nw == 0) CH_FAIL("nbit=%d unsupported (1..%d)", nbit, CH_MAX_NBIT);
  const int nbins = (tern ? 2 * nbit : nbit) + 1;
  g->tile = tile_rows_for(nw, tern, lab, lw);
  g->flush_tiles = 65535 / g->tile;
  auto per_sm_of = [&](int threads, SmemPlan* plan) {
    *plan = smem_plan(nbins, threads, nw, tern, lab, lw, g->tile);
    // resident CTAs per SM by shared memory (1 KB reserved per CTA) and threads
    long long per_sm = (228ll * 1024) / static_cast<long long>(plan->total + 1024);  // static + dynamic
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 32) per_sm = 32;
    return per_sm;
  };
  auto set_threads = [&](int threads) {
    SmemPlan p;
    per_sm_of(threads, &p);
    g->threads = threads;
    g->smem = p.total_dynamic;
    g->smem_total = p.total;
    g->nq_pad = padded_queries(nq, threads);
    g->nqtiles = static_cast<int>(g->nq_pad / threads);
    return p.total;
  };
  if (forced_stripes > 0 || forced_threads > 0) {
    const size_t total = set_threads(forced_threads > 0 ? forced_threads : pick_threads(nbins, nq));
    if (total > static_cast<size_t>(ws->max_smem_optin))
      CH_FAIL("histogram needs %zu bytes of shared memory per CTA (nbit=%d ternary=%d classes/32=%d), device has %d",
              total, nbit, tern ? 1 : 0, lw, ws->max_smem_optin);
  }
  if (forced_stripes > 0) {
    if (forced_rows_per_stripe <= 0 || forced_rows_per_stripe % 256 != 0 ||
        static_cast<long long>(forced_stripes) * forced_rows_per_stripe < ndb)
      CH_FAIL("bad stripe geometry: %d stripes x %d rows for %lld gallery rows", forced_stripes,
              forced_rows_per_stripe, ndb);
    g->nstripes = forced_stripes;
    g->rows_per_stripe = forced_rows_per_stripe;
    return 0;
  }
  // free choice: the (CTA size, stripe count) with the smallest modelled time.  CTA sizes: the largest the
  // histogram allows and its halves down to 64 queries (more query tiles = fewer, longer stripes for the same
  // number of CTAs = smaller slabs and less flush traffic -- what small galleries need).
  const int t_max = forced_threads > 0 ? forced_threads : pick_threads(nbins, nq);
  int best_t = t_max;
  long long best_s = 1, best_rps = 256;
  double best_cost = 1e300;
  for (int t = t_max; t >= (forced_threads > 0 ? t_max : (t_max < 64 ? t_max : 64)); t /= 2) {
    SmemPlan p;
    const long long per_sm = per_sm_of(t, &p);
    if (p.total > static_cast<size_t>(ws->max_smem_optin)) continue;
    const long long nq_pad = padded_queries(nq, t);
    const long long nqtiles = nq_pad / t;
    const long long slots = per_sm * ws->sm_count;
    long long hi = (8 * slots + nqtiles - 1) / nqtiles;
    const long long max_stripes = ndb > 0 ? (ndb + 255) / 256 : 1;
    if (hi > max_stripes) hi = max_stripes;
    if (hi > 4096) hi = 4096;
    if (hi < 1) hi = 1;
    long long last_s = 0;
    for (long long s0 = 1; s0 <= hi; ++s0) {
      long long rps = ndb > 0 ? (ndb + s0 - 1) / s0 : 256;
      rps = ch_round_up(rps, 256);
      const long long s_eff = ndb > 0 ? (ndb + rps - 1) / rps : 1;
      if (s_eff == last_s) continue;
      last_s = s_eff;
      // (the measured 83-88 % of the POPC peak belongs to the largest CTA size: a smaller one must win clearly)
      const double cost = model_cost(ws, ndb, nw, nbins, t, per_sm, nq_pad, s_eff, rps) * (t == t_max ? 1.0 : 1.1);
      if (cost < best_cost * (1.0 - 1e-9)) {
        best_cost = cost;
        best_t = t;
        best_s = s_eff;
        best_rps = rps;
      }
    }
  }
  const size_t total = set_threads(best_t);
  if (total > static_cast<size_t>(ws->max_smem_optin))
    CH_FAIL("histogram needs %zu bytes of shared memory per CTA (nbit=%d ternary=%d classes/32=%d), device has %d",
            total, nbit, tern ? 1 : 0, lw, ws->max_smem_optin);
  // stripe boundaries are multiples of 256 rows (>= every tile size, so every pass of one evaluation --
  // whatever its label mode -- sees the same stripes and every bulk copy starts 16-byte aligned)
  if (best_rps > 0x7fffff00ll) CH_FAIL("gallery shard too large for one stripe");
  g->rows_per_stripe = static_cast<int>(best_rps);
  g->nstripes = static_cast<int>(best_s);
  return 0;
}

}  // namespace

extern "C" int ch_hist_geometry(ch_ws* ws, int64_t nq, int64_t ndb, int nbit, int ternary, int label_mode,
                                int mask_words, int32_t* threads, int64_t* nq_pad, int32_t* nstripes,
                                int32_t* rows_per_stripe) {
  if (ws == nullptr) CH_FAIL("null workspace");
  Geometry g;
  if (make_geometry(ws, nq, ndb, nbit, ternary != 0, label_mode, mask_words, 0, 0, 0, &g)) return 1;
  // computed once per evaluation (with the richest label mode) and handed back to every
  // ch_hamming_hist call, so all passes share the same (threads, stripes, rows per stripe)
  if (threads) *threads = g.threads;
  if (nq_pad) *nq_pad = g.nq_pad;
  if (nstripes) *nstripes = g.nstripes;
  if (rows_per_stripe) *rows_per_stripe = g.rows_per_stripe;
  return 0;
}

extern "C" int ch_hamming_hist(ch_ws* ws, const ch_hist_args* a, void* stream) {
  if (ws == nullptr || a == nullptr) CH_FAIL("null argument to ch_hamming_hist");
  if (a->nq <= 0 || a->ndb < 0) CH_FAIL("bad sizes nq=%lld ndb=%lld", (long long)a->nq, (long long)a->ndb);
  if (a->q_bits == nullptr || a->g_bits == nullptr || a->slab_all == nullptr) CH_FAIL("null packed input / slab");
  if (a->ternary && (a->q_nz == nullptr || a->g_nz == nullptr)) CH_FAIL("ternary codes need the non-zero planes");
  if (a->label_mode != CH_LAB_NONE && (a->q_lab == nullptr || a->g_lab == nullptr || a->slab_rel == nullptr))
    CH_FAIL("label mode %d needs labels and the relevant slab", a->label_mode);
  if (a->emit_mode != CH_EMIT_NONE &&
      (a->rec_off == nullptr || a->rec_cap == nullptr || a->rec_cnt == nullptr || a->recs == nullptr ||
       a->err_flag == nullptr))
    CH_FAIL("record emission needs rec_off / rec_cap / rec_cnt / recs / err_flag");
  if (a->emit_mode == CH_EMIT_RELEVANT && a->label_mode == CH_LAB_NONE) CH_FAIL("relevant records need labels");
  if ((reinterpret_cast<uintptr_t>(a->g_bits) | reinterpret_cast<uintptr_t>(a->g_nz) |
       reinterpret_cast<uintptr_t>(a->g_lab)) & 15)
    CH_FAIL("gallery arrays must be 16-byte aligned");
  ChDeviceGuard guard(ws->device);
  Geometry g;
  if (a->threads <= 0 || a->threads > 256 || (a->threads & 31) || a->nstripes <= 0)
    CH_FAIL("bad geometry threads=%d nstripes=%d (use ch_hist_geometry)", a->threads, a->nstripes);
  if (make_geometry(ws, a->nq, a->ndb, a->nbit, a->ternary != 0, a->label_mode, a->mask_words, a->threads,
                    a->nstripes, a->rows_per_stripe, &g))
    return 1;
  if (g.nq_pad != a->nq_pad) CH_FAIL("nq_pad mismatch: got %lld, geometry says %lld", (long long)a->nq_pad, g.nq_pad);
  const int nw = ch_code_words(a->nbit);
  hist_fn_t fn = pick_kernel(nw, a->ternary != 0, a->label_mode, a->thresh != nullptr);
  if (fn == nullptr) CH_FAIL("no kernel for words=%d", nw);
  HistDev d;
  d.q_bits = a->q_bits; d.q_nz = a->q_nz; d.g_bits = a->g_bits; d.g_nz = a->g_nz;
  d.q_lab = a->q_lab; d.g_lab = a->g_lab; d.slab_all = a->slab_all; d.slab_rel = a->slab_rel;
  d.thresh = a->thresh; d.rec_off = a->rec_off; d.rec_cap = a->rec_cap; d.rec_cnt = a->rec_cnt;
  d.recs = static_cast<uint4*>(a->recs); d.err_flag = a->err_flag;
  d.nq = a->nq; d.nq_pad = a->nq_pad; d.ndb = a->ndb; d.row_base = a->row_base;
  d.nbit = a->nbit; d.nbins = (a->ternary ? 2 * a->nbit : a->nbit) + 1; d.lw = a->mask_words;
  // select pass: keys above max(thresh) never occur -> the caller may keep narrower slabs / histograms
  if (a->thresh != nullptr && a->key_limit > 0 && a->key_limit < d.nbins) d.nbins = a->key_limit;
  d.emit_mode = a->emit_mode;
  d.nqtiles = g.nqtiles; d.rows_per_stripe = g.rows_per_stripe; d.tile_rows = g.tile; d.flush_tiles = g.flush_tiles;
  CH_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(g.smem)));
  const long long ctas = static_cast<long long>(g.nqtiles) * g.nstripes;
  if (ctas > 0x7fffffffll) CH_FAIL("grid too large");
  fn<<<static_cast<unsigned>(ctas), g.threads, g.smem, static_cast<cudaStream_t>(stream)>>>(d);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_slab_totals(ch_ws* ws, const uint32_t* slab, int nstripes, int nbins, int64_t nq_pad,
                              uint32_t* tot_dev, void* stream) {
  if (ws == nullptr || slab == nullptr || tot_dev == nullptr) CH_FAIL("null argument to ch_slab_totals");
  ChDeviceGuard guard(ws->device);
  const long long plane = static_cast<long long>(nbins) * nq_pad;
  slab_totals_kernel<<<static_cast<unsigned>((plane + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      slab, nstripes, plane, tot_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_slab_exscan(ch_ws* ws, uint32_t* slab, int nstripes, int nbins, int64_t nq_pad, void* stream) {
  if (ws == nullptr || slab == nullptr) CH_FAIL("null argument to ch_slab_exscan");
  ChDeviceGuard guard(ws->device);
  const long long plane = static_cast<long long>(nbins) * nq_pad;
  slab_exscan_kernel<<<static_cast<unsigned>((plane + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      slab, nstripes, plane);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_slab_scan(ch_ws* ws, uint32_t* slab, int nslabs, int nstripes, int nbins, int64_t nq_pad,
                            uint32_t* tot_dev, void* stream) {
  if (ws == nullptr || slab == nullptr || tot_dev == nullptr) CH_FAIL("null argument to ch_slab_scan");
  if (nslabs < 1 || nslabs > 2 || nstripes < 1) CH_FAIL("bad arguments to ch_slab_scan");
  ChDeviceGuard guard(ws->device);
  const long long plane = static_cast<long long>(nbins) * nq_pad;
  slab_scan_kernel<<<static_cast<unsigned>((plane * nslabs + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      slab, nslabs, nstripes, plane, tot_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_class_counts(ch_ws* ws, const uint32_t* g_ids, int64_t ndb, int rows_per_stripe, int nclass,
                               uint32_t* cls_dev, void* stream) {
  if (ws == nullptr || g_ids == nullptr || cls_dev == nullptr) CH_FAIL("null argument to ch_class_counts");
  if (ndb <= 0) return 0;
  ChDeviceGuard guard(ws->device);
  class_counts_kernel<<<static_cast<unsigned>((ndb + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g_ids, ndb, rows_per_stripe, nclass, cls_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}
