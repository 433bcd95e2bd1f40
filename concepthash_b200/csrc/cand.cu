// K3/K4 on candidate lists (the output of the tensor-core select pass, select_tc.cu).
//
// A candidate list holds, per (stripe, query), the shard-local row indices of the pairs with key <= thresh[q], in
// ascending row order; concatenated over the stripes it is the query's candidate set in global tie order.  Because
// EVERY item with key <= thresh is in the list, the canonical rank of a candidate is a count over candidates only:
//
//     rank(j) = #{cand : key < key_j} (all ranks) + #{cand on lower ranks : key = key_j} + #{cand before j : key = key_j}
//
// ch_cand_hist      one warp per query: key (XOR + POPC on the packed codes) and label match of every candidate;
//                   per-query key histograms {all, relevant} -> (nbins, nq_pad) totals (all-gathered by the caller,
//                   turned into bases by ch_scan_bases); the key and the relevance bit are kept per candidate.
// ch_cand_finalize  one warp per query walks the list in order, 32 candidates at a time: stable in-bucket prefixes
//                   by __match_any_sync + running per-key counters in shared memory; then either the AP / P@k sums
//                   (mode 0), the relevance of the rank-0 item (mode 1) or the ranked id list (mode 2).
#include "common.cuh"

namespace {

struct CandDev {
  const uint32_t* cand_off; const uint32_t* cand_cnt;
  uint32_t* cand_rows; uint16_t* cand_key;
  const uint32_t* q_bits; const uint32_t* g_bits;
  const uint32_t* q_nz; const uint32_t* g_nz;   // non-zero planes (ternary codes: keys on the doubled scale) or NULL
  const uint32_t* g_plane;   // (rows, PW) [code words | class id | pad]: one sector per candidate; or NULL
  const uint32_t* q_lab; const uint32_t* g_lab;
  uint32_t* tot_all; uint32_t* tot_rel;
  const uint32_t* base0_all; const uint32_t* base0_rel;
  const uint32_t* first_rel; uint32_t* first_rel_out;
  const uint32_t* key_max;   // (nq_pad) or NULL: candidates with a larger key cannot rank below rmax
  double* cols; long long* ids; int* keys;
  uint32_t* err_flag;
  long long nq, nq_pad, R, row_offset;
  int nstripes, nbins, nbit, label_mode, lw, remove_first, nR, nPR, mode;
  long long r_eff[CH_MAX_R];
  long long pr_k[CH_MAX_PR];
};

constexpr int kCandWarps = 8;

template <int W>
__device__ __forceinline__ uint32_t key_of(const uint32_t (&qw)[W], const uint32_t* __restrict__ g_bits, uint32_t row) {
  uint32_t gw[W];
  if constexpr (W == 4) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(g_bits) + row);
    gw[0] = v.x; gw[1] = v.y; gw[2] = v.z; gw[3] = v.w;
  } else if constexpr (W == 2) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(g_bits) + row);
    gw[0] = v.x; gw[1] = v.y;
  } else {
#pragma unroll
    for (int w = 0; w < W; ++w) gw[w] = __ldg(g_bits + static_cast<size_t>(row) * W + w);
  }
  uint32_t key = 0;
#pragma unroll
  for (int w = 0; w < W; ++w) key += __popc(qw[w] ^ gw[w]);
  return key;
}

// ternary codes: key = 2 x distance = nbit - #(both non-zero) + 2 #(both non-zero and signs differ)
template <int W>
__device__ __forceinline__ uint32_t key_of_ternary(const uint32_t (&qw)[W], const uint32_t (&qz)[W],
                                                   const uint32_t* __restrict__ g_bits,
                                                   const uint32_t* __restrict__ g_nz, uint32_t row, int nbit) {
  uint32_t both = 0, dis = 0;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const uint32_t gb = __ldg(g_bits + static_cast<size_t>(row) * W + w);
    const uint32_t gz = __ldg(g_nz + static_cast<size_t>(row) * W + w);
    const uint32_t m = qz[w] & gz;
    both += __popc(m);
    dis += __popc((qw[w] ^ gb) & m);
  }
  return static_cast<uint32_t>(nbit) - both + 2u * dis;
}

// words per row of the gather plane: the code words + the class id, rounded up to a power of two (8 .. 64 bytes)
__host__ __device__ constexpr int plane_words(int W) { return W == 1 ? 2 : W == 2 ? 4 : W == 4 ? 8 : 16; }

// one row of the gather plane ([code words | class id | pad]): ONE 32-byte sector for codes of <= 128 bits
template <int W>
__device__ __forceinline__ void load_plane_row(const uint32_t* __restrict__ plane, uint32_t row,
                                               uint32_t (&g)[plane_words(W)]) {
  constexpr int PW = plane_words(W);
  if constexpr (PW == 2) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(plane) + row);
    g[0] = v.x; g[1] = v.y;
  } else if constexpr (PW == 4) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(plane) + row);
    g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
  } else {
    // 32-byte rows: ONE 256-bit load (LDG.E.256, sm_100) instead of two 128-bit ones (a random gather costs the L1
    // one wavefront per lane and instruction)
#pragma unroll
    for (int v8 = 0; v8 < PW / 8; ++v8) {
      const uint32_t* src = plane + static_cast<size_t>(row) * PW + 8 * v8;
      asm("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
          : "=r"(g[8 * v8]), "=r"(g[8 * v8 + 1]), "=r"(g[8 * v8 + 2]), "=r"(g[8 * v8 + 3]), "=r"(g[8 * v8 + 4]),
            "=r"(g[8 * v8 + 5]), "=r"(g[8 * v8 + 6]), "=r"(g[8 * v8 + 7])
          : "l"(src));
    }
  }
}

__global__ void gather_plane_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ ids, long long rows,
                                    int W, int PW, uint32_t* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * PW) return;
  const long long r = i / PW;
  const int w = static_cast<int>(i - r * PW);
  out[i] = w < W ? bits[r * W + w] : (w == W ? ids[r] : 0u);
}

// keys + label matches of the candidates of query q (one warp): cand_key / the relevance bit of cand_rows are written,
// the key histograms {all, relevant} are left in shared memory (h_rel = h_all + nbins).  Returns "some key >= nbins".
template <int W>
__device__ __forceinline__ bool cand_hist_body(const CandDev& a, long long q, int lane, uint32_t* h_all) {
  uint32_t* h_rel = h_all + a.nbins;
  for (int b = lane; b < 2 * a.nbins; b += 32) h_all[b] = 0u;
  uint32_t qw[W], qz[W];
  const bool tern = a.q_nz != nullptr;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    qw[w] = a.q_bits[q * W + w];
    qz[w] = tern ? a.q_nz[q * W + w] : 0xffffffffu;
  }
  const uint32_t qid = a.label_mode == CH_LAB_ID ? a.q_lab[q] : 0u;
  const uint32_t* qm = a.label_mode == CH_LAB_MASK ? a.q_lab + q * a.lw : nullptr;
  bool bad = false;
  __syncwarp();
  // candidates per lane in flight.  The gathers of all U candidates are issued back to back, UNCONDITIONALLY (lanes
  // past the end of the slice read row 0), before the first one is used: with the load inside a predicated block per
  // candidate the compiler serialised them, and the kernel sat in long-scoreboard stalls (63 % of its samples) with
  // one random sector in flight per lane.
  constexpr int U = plane_words(W) <= 4 ? 8 : 4;
  // The slices of a query (one per stripe) are walked as ONE list, 32 stripes at a time: lane s holds (offset, count) of
  // stripe s0 + s and the exclusive prefix of the counts; position p of the concatenation belongs to the last stripe
  // whose prefix is <= p (a 5-step search over the lanes' registers).  A row-sharded or streamed gallery gives a query
  // 30-60 candidates per slice: walked slice by slice, a 32 U-wide step was mostly idle lanes (cand_hist took 0.40 ms
  // for a quarter of the candidates that the whole list takes 0.58 ms for, walk included).
  for (int s0 = 0; s0 < a.nstripes; s0 += 32) {
    const int s = s0 + lane;
    uint32_t my_off = 0u, my_n = 0u;
    if (s < a.nstripes) {
      const size_t sq = static_cast<size_t>(s) * a.nq_pad + q;
      my_off = a.cand_off[sq];
      my_n = a.cand_cnt[sq];
    }
    uint32_t inc = my_n;                                   // inclusive prefix of the counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const uint32_t my_pre = inc - my_n;
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    for (uint32_t i0 = 0; i0 < total; i0 += 32 * U) {
      uint32_t row[U], key[U], addr[U];
      bool ok[U], rel[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t p = i0 + u * 32 + lane;
        ok[u] = p < total;
        // last lane j with pre_j <= p  (pre is non-decreasing; empty slices share a prefix with their successor, the
        // search lands on the LAST of them, whose count covers p)
        int j = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const uint32_t pj = __shfl_sync(0xffffffffu, my_pre, (j + step) & 31);
          if (j + step < 32 && pj <= p) j += step;
        }
        const uint32_t off_j = __shfl_sync(0xffffffffu, my_off, j), pre_j = __shfl_sync(0xffffffffu, my_pre, j);
        addr[u] = off_j + (p - pre_j);
        row[u] = ok[u] ? (a.cand_rows[addr[u]] & 0x7fffffffu) : 0u;
      }
      if (a.g_plane != nullptr) {              // single-label: code and class id come in one sector
        uint32_t g[U][plane_words(W)];
#pragma unroll
        for (int u = 0; u < U; ++u) load_plane_row<W>(a.g_plane, row[u], g[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          uint32_t k = 0;
#pragma unroll
          for (int w = 0; w < W; ++w) k += __popc(qw[w] ^ g[u][w]);
          key[u] = k;
          rel[u] = ok[u] && g[u][W] == qid;
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          rel[u] = false;
          if (tern) key[u] = ok[u] ? key_of_ternary<W>(qw, qz, a.g_bits, a.g_nz, row[u], a.nbit) : 0u;
          else key[u] = ok[u] ? key_of<W>(qw, a.g_bits, row[u]) : 0u;
          if (ok[u]) {
            if (a.label_mode == CH_LAB_ID) {
              rel[u] = __ldg(a.g_lab + row[u]) == qid;
            } else if (a.label_mode == CH_LAB_MASK) {
              uint32_t any = 0;
              for (int w = 0; w < a.lw; ++w) any |= qm[w] & __ldg(a.g_lab + static_cast<size_t>(row[u]) * a.lw + w);
              rel[u] = any != 0u;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        if (key[u] < static_cast<uint32_t>(a.nbins)) {
          atomicAdd(&h_all[key[u]], 1u);
          if (rel[u]) atomicAdd(&h_rel[key[u]], 1u);
        } else {
          bad = true;
        }
        a.cand_key[addr[u]] = static_cast<uint16_t>(key[u]);
        if (rel[u]) a.cand_rows[addr[u]] = row[u] | 0x80000000u;
      }
    }
  }
  __syncwarp();
  return bad;
}

template <int W>
__global__ void __launch_bounds__(kCandWarps * 32) cand_hist_kernel(const CandDev a) {
  extern __shared__ uint32_t sh[];
  const int wip = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = static_cast<long long>(blockIdx.x) * kCandWarps + wip;
  if (q >= a.nq) return;                       // whole warps leave; only __syncwarp below
  uint32_t* h_all = sh + static_cast<size_t>(wip) * 2 * a.nbins;
  uint32_t* h_rel = h_all + a.nbins;
  const bool bad = cand_hist_body<W>(a, q, lane, h_all);
  // accumulated: the caller zero-initialises the totals and may histogram a list in several calls (row blocks of
  // a streamed gallery); one warp owns a query, calls are stream-ordered -> no atomics
  for (int b = lane; b < a.nbins; b += 32) {
    if (h_all[b] != 0u) a.tot_all[static_cast<size_t>(b) * a.nq_pad + q] += h_all[b];
    if (a.tot_rel != nullptr && h_rel[b] != 0u) a.tot_rel[static_cast<size_t>(b) * a.nq_pad + q] += h_rel[b];
  }
  if (bad) atomicOr(a.err_flag, 2u);
}


// 1 / x for x = 1 .. 2^32: MUFU.RCP64H seed + two Newton steps (relative error < 2^-52; the AP sums only need
// ~1e-12, a correctly rounded division costs ~40 instructions per relevant candidate)
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}

// the in-order walk of query q's candidate list (one warp): run_all / run_rel (shared memory, run_rel = run_all + nbins)
// hold the bases on entry; qrow = 128 words of shared memory for the compaction queue
__device__ __forceinline__ void cand_final_body(const CandDev& a, long long q, int lane, uint32_t* run_all,
                                                uint32_t* qrow, uint32_t kmax) {
  uint32_t* run_rel = run_all + a.nbins;
  const int ncols = 2 * a.nR + a.nPR;
  double acc[2 * CH_MAX_R + CH_MAX_PR];
  if (a.mode == 0)
    for (int c = 0; c < ncols; ++c) acc[c] = 0.0;
  const long long shift = a.remove_first ? 1 : 0;
  const long long frel = (a.remove_first && a.first_rel != nullptr) ? a.first_rel[q] : 0;
  const uint32_t lt = lanemask_lt();
  // Only candidates with key <= kmax[q] (the smallest key at which the list holds rmax items) can rank below rmax:
  // the others are dropped while the list is read, the survivors are compacted (in order) through a small
  // shared-memory queue and ranked 32 at a time.
  uint32_t* qkey = qrow + 64;
  uint32_t queued = 0;                                            // warp-uniform
  // ranks the candidates held by the lanes (key == 0xffffffff: none) -- one step of the in-order walk
  auto step = [&](uint32_t rowv, uint32_t key) {
    const bool valid = key != 0xffffffffu;
    const bool rel = valid && (rowv >> 31) != 0u;
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const uint32_t relm = __ballot_sync(0xffffffffu, rel);
    long long rank = 0, relrank = 0;
    if (valid) {
      rank = static_cast<long long>(run_all[key]) + __popc(peers & lt);
      relrank = static_cast<long long>(run_rel[key]) + __popc(peers & relm & lt);
    }
    __syncwarp();
    if (valid && (peers >> lane) == 1u) {         // highest lane of the group publishes the new running counts
      run_all[key] += __popc(peers);
      run_rel[key] += __popc(peers & relm);
    }
    __syncwarp();
    if (!valid) return;
    if (a.mode == 1) {
      if (rank == 0 && rel) a.first_rel_out[q] = 1u;
      return;
    }
    if (a.remove_first) {
      if (rank == 0) return;                      // the dropped self-retrieval
      rank -= shift;
      relrank -= frel;
    }
    if (a.mode == 2) {
      if (rank < a.R) {
        a.ids[q * a.R + rank] = a.row_offset + static_cast<long long>(rowv & 0x7fffffffu);
        if (a.keys != nullptr) a.keys[q * a.R + rank] = static_cast<int>(key);
      }
      return;
    }
    if (!rel) return;
    const double prec = static_cast<double>(relrank + 1) * fast_rcp(static_cast<double>(rank + 1));
    for (int j = 0; j < a.nR; ++j)
      if (rank < a.r_eff[j]) {
        acc[2 * j] += prec;
        acc[2 * j + 1] += 1.0;
      }
    for (int j = 0; j < a.nPR; ++j)
      if (rank < a.pr_k[j]) acc[2 * a.nR + j] += 1.0;
  };
  __syncwarp();
  for (int s = 0; s < a.nstripes; ++s) {
    const size_t sq = static_cast<size_t>(s) * a.nq_pad + q;
    const uint32_t off = a.cand_off[sq], n = a.cand_cnt[sq];
    // (the loads of the next 32 candidates are issued before the current 32 are ranked: the walk is a serial chain
    // of ballots / match_any / shared-memory counters and would otherwise expose one global round trip per step)
    uint32_t nrow = 0u, nkey = 0xffffffffu;
    if (lane < n) {
      nrow = a.cand_rows[off + lane];
      nkey = a.cand_key[off + lane];
    }
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
      const uint32_t i = i0 + lane;
      const uint32_t rowv = nrow, key = nkey;
      if (i + 32 < n) {
        nrow = a.cand_rows[off + i + 32];
        nkey = a.cand_key[off + i + 32];
      } else {
        nrow = 0u;
        nkey = 0xffffffffu;
      }
      // keys >= nbins were flagged by ch_cand_hist; keys > kmax cannot rank below rmax
      const bool keep = i < n && key < static_cast<uint32_t>(a.nbins) && key <= kmax;
      const uint32_t km = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const uint32_t pos = queued + __popc(km & lt);
        qrow[pos] = rowv;
        qkey[pos] = key;
      }
      queued += __popc(km);
      __syncwarp();
      if (queued >= 32u) {
        const uint32_t r0 = qrow[lane], k0 = qkey[lane];
        const uint32_t rest = queued - 32u;
        uint32_t r1 = 0u, k1 = 0u;
        if (lane < rest) {
          r1 = qrow[32 + lane];
          k1 = qkey[32 + lane];
        }
        __syncwarp();
        if (lane < rest) {
          qrow[lane] = r1;
          qkey[lane] = k1;
        }
        queued = rest;
        __syncwarp();
        step(r0, k0);
      }
    }
  }
  if (queued > 0u) step(lane < queued ? qrow[lane] : 0u, lane < queued ? qkey[lane] : 0xffffffffu);
  if (a.mode == 0) {
    for (int c = 0; c < ncols; ++c) {
      double v = acc[c];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // fixed tree: deterministic
      if (lane == 0) a.cols[q * ncols + c] = v;
    }
  }
}

__global__ void __launch_bounds__(kCandWarps * 32) cand_final_kernel(const CandDev a) {
  extern __shared__ uint32_t sh[];
  const int wip = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = static_cast<long long>(blockIdx.x) * kCandWarps + wip;
  if (q >= a.nq) return;
  uint32_t* run_all = sh + static_cast<size_t>(wip) * 2 * a.nbins;
  uint32_t* run_rel = run_all + a.nbins;
  for (int b = lane; b < a.nbins; b += 32) {
    run_all[b] = a.base0_all[static_cast<size_t>(b) * a.nq_pad + q];
    run_rel[b] = a.base0_rel != nullptr ? a.base0_rel[static_cast<size_t>(b) * a.nq_pad + q] : 0u;
  }
  const uint32_t kmax = a.key_max != nullptr ? a.key_max[q] : 0xfffffffeu;
  cand_final_body(a, q, lane, run_all, sh + static_cast<size_t>(kCandWarps) * 2 * a.nbins + static_cast<size_t>(wip) * 128,
                  kmax);
}

// ch_cand_hist + ch_scan_bases_pair + ch_cand_finalize (mode 0) of ONE rank in one kernel: a warp histograms the keys
// of its query's candidates, turns the histogram into the bases by a shared-memory scan (the totals of a single rank
// need no exchange), verifies the list (>= need candidates, else bad[q] / status) and walks it.  The two phases have
// complementary bounds -- random L2 gathers vs. a serial, issue-bound walk -- and the warps of an SM are in
// different phases at any moment, so the kernel takes about as long as the slower phase alone.
template <int W>
__global__ void __launch_bounds__(kCandWarps * 32) cand_rank_kernel(const CandDev a, long long rmax, long long need,
                                                                    uint32_t* __restrict__ status,
                                                                    uint32_t* __restrict__ bad_q) {
  extern __shared__ uint32_t sh[];
  const int wip = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = static_cast<long long>(blockIdx.x) * kCandWarps + wip;
  if (q >= a.nq) return;
  uint32_t* h_all = sh + static_cast<size_t>(wip) * 2 * a.nbins;
  uint32_t* h_rel = h_all + a.nbins;
  if (cand_hist_body<W>(a, q, lane, h_all)) atomicOr(a.err_flag, 2u);
  // histogram -> exclusive bases, 32 keys at a time; kmax = the smallest key at which the list holds rmax items
  uint32_t carry_a = 0u, carry_r = 0u, kmax = static_cast<uint32_t>(a.nbins - 1);
  bool found = false;
  for (int b0 = 0; b0 < a.nbins; b0 += 32) {
    const int b = b0 + lane;
    const uint32_t va = b < a.nbins ? h_all[b] : 0u, vr = b < a.nbins ? h_rel[b] : 0u;
    uint32_t ia = va, ir = vr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o), tr = __shfl_up_sync(0xffffffffu, ir, o);
      if (lane >= o) {
        ia += ta;
        ir += tr;
      }
    }
    __syncwarp();
    if (b < a.nbins) {
      h_all[b] = carry_a + ia - va;
      h_rel[b] = carry_r + ir - vr;
    }
    const uint32_t hit = __ballot_sync(0xffffffffu, b < a.nbins && rmax >= 0 &&
                                                         static_cast<long long>(carry_a + ia) >= rmax);
    if (!found && hit != 0u) {
      kmax = static_cast<uint32_t>(b0 + __ffs(static_cast<int>(hit)) - 1);
      found = true;
    }
    carry_a += __shfl_sync(0xffffffffu, ia, 31);
    carry_r += __shfl_sync(0xffffffffu, ir, 31);
  }
  __syncwarp();
  if (need > 0 && static_cast<long long>(carry_a) < need && lane == 0) {
    if (bad_q != nullptr) bad_q[q] = 1u;
    if (status != nullptr) atomicOr(status, 1u);
  }
  cand_final_body(a, q, lane, h_all, sh + static_cast<size_t>(kCandWarps) * 2 * a.nbins + static_cast<size_t>(wip) * 128,
                  kmax);
}

// capacities of the full pass from a candidate list of the ROW SAMPLE (rows = sample row indices):
// cap[s][q] = stride * (k + 6 sqrt(k + 1) + 9), k = #sample candidates of query q with key <= thresh[q] whose row
// lies in stripe s of the full pass (sample rows [s * rows_per_stripe, (s + 1) * rows_per_stripe)) -- the bound of
// record_caps_kernel.  One warp per query, per-stripe counters in shared memory.
// scut != NULL: the threshold of a query is refined to a (key, stripe) pair.  Candidates are a PREFIX of the canonical
// (key, row) order, and stripes are row blocks, so "key < t, or key == t in a stripe below scut" is a prefix too:
// scut[q] = the smallest stripe count at which the sample holds >= m items of that prefix (nstripes: all of them are
// needed).  Stripes >= scut[q] then run with the threshold t - 1 -- the items with key == t are a large part of a
// key-level candidate list, and only those of the first few stripes can reach the top R.
__global__ void __launch_bounds__(kCandWarps * 32) cand_caps_kernel(
    const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ rows,
    const uint16_t* __restrict__ key, const uint32_t* __restrict__ thresh, int list_stripes, int rows_per_stripe,
    int nstripes, long long nq, long long nq_pad, int stride, uint32_t* __restrict__ cap, int m,
    uint32_t* __restrict__ scut) {
  extern __shared__ uint32_t sh[];                       // per warp: nstripes counters (key < t), nstripes (key == t)
  const int wip = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = static_cast<long long>(blockIdx.x) * kCandWarps + wip;
  if (q >= nq_pad) return;                               // whole warps leave; only __syncwarp below
  uint32_t* c = sh + static_cast<size_t>(wip) * 2 * nstripes;
  uint32_t* ce = c + nstripes;
  for (int s = lane; s < 2 * nstripes; s += 32) c[s] = 0u;
  __syncwarp();
  if (q < nq) {
    const uint32_t t = thresh[q];
    for (int ls = 0; ls < list_stripes; ++ls) {
      const size_t i = static_cast<size_t>(ls) * nq_pad + q;
      const uint32_t o = off[i], n = cnt[i];
      for (uint32_t j = lane; j < n; j += 32) {
        const uint32_t k = key[o + j];
        if (k > t) continue;
        int s = static_cast<int>((rows[o + j] & 0x7fffffffu) / static_cast<uint32_t>(rows_per_stripe));
        if (s >= nstripes) s = nstripes - 1;
        atomicAdd(k == t ? &ce[s] : &c[s], 1u);
      }
    }
  }
  __syncwarp();
  int cut = nstripes;
  if (scut != nullptr) {
    if (q < nq) {
      // (every lane walks the few stripes itself: no divergence, no shuffles)
      uint32_t have = 0u;
      for (int s = 0; s < nstripes; ++s) have += c[s];
      cut = 0;
      while (cut < nstripes && have < static_cast<uint32_t>(m)) have += ce[cut++];
      if (have < static_cast<uint32_t>(m)) cut = nstripes;
    }
    if (lane == 0) scut[q] = static_cast<uint32_t>(cut);
  }
  for (int s = lane; s < nstripes; s += 32) {
    uint32_t v = 0u;
    if (q < nq) {
      const float kf = static_cast<float>(c[s] + (s < cut ? ce[s] : 0u));
      v = static_cast<uint32_t>((kf + 6.0f * sqrtf(kf + 1.0f) + 9.0f) * static_cast<float>(stride));
    }
    cap[static_cast<size_t>(s) * nq_pad + q] = v;
  }
}

int to_dev(const ch_cand_args* a, CandDev* d) {
  if (a == nullptr) CH_FAIL("null ch_cand_args");
  if (a->cand_off == nullptr || a->cand_cnt == nullptr || a->cand_rows == nullptr || a->cand_key == nullptr)
    CH_FAIL("null candidate arrays");
  if (a->nR < 0 || a->nR > CH_MAX_R || a->nPR < 0 || a->nPR > CH_MAX_PR) CH_FAIL("too many R / PRs entries");
  if (a->nbins <= 0 || a->nbins > 2 * CH_MAX_NBIT + 1) CH_FAIL("bad nbins %d", a->nbins);
  if (a->nq <= 0 || a->nq_pad < a->nq || a->nstripes <= 0) CH_FAIL("bad candidate geometry");
  d->cand_off = a->cand_off; d->cand_cnt = a->cand_cnt; d->cand_rows = a->cand_rows; d->cand_key = a->cand_key;
  d->q_bits = a->q_bits; d->g_bits = a->g_bits; d->q_lab = a->q_lab; d->g_lab = a->g_lab;
  d->q_nz = a->q_nz; d->g_nz = a->q_nz != nullptr ? a->g_nz : nullptr;
  if (a->q_nz != nullptr && a->g_nz == nullptr) CH_FAIL("ternary keys need both non-zero planes");
  d->nbit = a->nbit;
  // the gather plane holds [code | class id]: binary single-label shards only
  d->g_plane = (a->label_mode == CH_LAB_ID && a->q_nz == nullptr) ? a->g_plane : nullptr;
  d->tot_all = a->tot_all; d->tot_rel = a->tot_rel; d->base0_all = a->base0_all; d->base0_rel = a->base0_rel;
  d->first_rel = a->first_rel; d->first_rel_out = a->first_rel_out; d->cols = a->cols;
  d->key_max = a->key_max;
  d->ids = reinterpret_cast<long long*>(a->ids); d->keys = a->keys; d->err_flag = a->err_flag;
  d->nq = a->nq; d->nq_pad = a->nq_pad; d->R = a->R; d->row_offset = a->row_offset;
  d->nstripes = a->nstripes; d->nbins = a->nbins; d->label_mode = a->label_mode; d->lw = a->mask_words;
  d->remove_first = a->remove_first; d->nR = a->nR; d->nPR = a->nPR; d->mode = a->mode;
  for (int i = 0; i < CH_MAX_R; ++i) d->r_eff[i] = i < a->nR ? a->r_eff[i] : 0;
  for (int i = 0; i < CH_MAX_PR; ++i) d->pr_k[i] = i < a->nPR ? a->pr_k[i] : 0;
  return 0;
}

}  // namespace

extern "C" int ch_gather_plane_words(int nbit) {
  const int w = ch_code_words(nbit);
  return w == 0 ? 0 : plane_words(w);
}

extern "C" int ch_gather_plane(ch_ws* ws, const uint32_t* bits_dev, const uint32_t* ids_dev, int64_t rows, int nbit,
                               uint32_t* out_dev, void* stream) {
  if (ws == nullptr || bits_dev == nullptr || ids_dev == nullptr || out_dev == nullptr)
    CH_FAIL("null argument to ch_gather_plane");
  const int W = ch_code_words(nbit);
  if (W == 0 || rows < 0) CH_FAIL("bad arguments to ch_gather_plane");
  if (rows == 0) return 0;
  ChDeviceGuard guard(ws->device);
  const int PW = plane_words(W);
  const long long n = rows * PW;
  gather_plane_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bits_dev, ids_dev, rows, W, PW, out_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_cand_hist(ch_ws* ws, const ch_cand_args* a, void* stream) {
  if (ws == nullptr) CH_FAIL("null workspace");
  CandDev d;
  if (to_dev(a, &d)) return 1;
  if (a->q_bits == nullptr || a->g_bits == nullptr || a->tot_all == nullptr || a->err_flag == nullptr)
    CH_FAIL("null argument to ch_cand_hist");
  if (a->label_mode != CH_LAB_NONE && (a->q_lab == nullptr || a->g_lab == nullptr || a->tot_rel == nullptr))
    CH_FAIL("labels / relevant totals missing");
  if (a->label_mode == CH_LAB_NONE) d.tot_rel = nullptr;
  const int words = ch_code_words(a->nbit);
  if (words == 0) CH_FAIL("nbit=%d unsupported", a->nbit);
  ChDeviceGuard guard(ws->device);
  const unsigned blocks = static_cast<unsigned>((a->nq + kCandWarps - 1) / kCandWarps);
  const size_t smem = static_cast<size_t>(kCandWarps) * 2 * a->nbins * sizeof(uint32_t);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (words) {
    case 1: cand_hist_kernel<1><<<blocks, kCandWarps * 32, smem, st>>>(d); break;
    case 2: cand_hist_kernel<2><<<blocks, kCandWarps * 32, smem, st>>>(d); break;
    case 4: cand_hist_kernel<4><<<blocks, kCandWarps * 32, smem, st>>>(d); break;
    default: cand_hist_kernel<8><<<blocks, kCandWarps * 32, smem, st>>>(d); break;
  }
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_cand_rank(ch_ws* ws, const ch_cand_args* a, int64_t rmax, int64_t need, uint32_t* status_dev,
                            uint32_t* bad_dev, void* stream) {
  if (ws == nullptr) CH_FAIL("null workspace");
  CandDev d;
  if (to_dev(a, &d)) return 1;
  if (a->q_bits == nullptr || a->g_bits == nullptr || a->err_flag == nullptr || a->cols == nullptr)
    CH_FAIL("null argument to ch_cand_rank");
  if (a->label_mode != CH_LAB_NONE && (a->q_lab == nullptr || a->g_lab == nullptr)) CH_FAIL("labels missing");
  if (a->mode != 0 || a->remove_first) CH_FAIL("ch_cand_rank computes the AP sums (mode 0) without remove_first");
  if (need > 0 && status_dev == nullptr) CH_FAIL("verification needs a status word");
  const int words = ch_code_words(a->nbit);
  if (words == 0) CH_FAIL("nbit=%d unsupported", a->nbit);
  ChDeviceGuard guard(ws->device);
  d.tot_all = d.tot_rel = nullptr;
  d.base0_all = d.base0_rel = nullptr;
  d.key_max = nullptr;
  const unsigned blocks = static_cast<unsigned>((a->nq + kCandWarps - 1) / kCandWarps);
  const size_t smem = static_cast<size_t>(kCandWarps) * (2 * a->nbins + 128) * sizeof(uint32_t);   // counters + queue
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (words) {
    case 1: cand_rank_kernel<1><<<blocks, kCandWarps * 32, smem, st>>>(d, rmax, need, status_dev, bad_dev); break;
    case 2: cand_rank_kernel<2><<<blocks, kCandWarps * 32, smem, st>>>(d, rmax, need, status_dev, bad_dev); break;
    case 4: cand_rank_kernel<4><<<blocks, kCandWarps * 32, smem, st>>>(d, rmax, need, status_dev, bad_dev); break;
    default: cand_rank_kernel<8><<<blocks, kCandWarps * 32, smem, st>>>(d, rmax, need, status_dev, bad_dev); break;
  }
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_cand_caps(ch_ws* ws, const uint32_t* cand_off, const uint32_t* cand_cnt, const uint32_t* cand_rows,
                            const uint16_t* cand_key, const uint32_t* thresh, int list_stripes, int rows_per_stripe,
                            int nstripes, int64_t nq, int64_t nq_pad, int sample_stride, uint32_t* cap_dev,
                            int m, uint32_t* scut_dev, void* stream) {
  if (ws == nullptr || cand_off == nullptr || cand_cnt == nullptr || cand_rows == nullptr || cand_key == nullptr ||
      thresh == nullptr || cap_dev == nullptr)
    CH_FAIL("null argument to ch_cand_caps");
  if (list_stripes <= 0 || nstripes <= 0 || rows_per_stripe <= 0 || nq <= 0 || nq_pad < nq || sample_stride < 1)
    CH_FAIL("bad arguments to ch_cand_caps");
  ChDeviceGuard guard(ws->device);
  cand_caps_kernel<<<static_cast<unsigned>((nq_pad + kCandWarps - 1) / kCandWarps), kCandWarps * 32,
                     static_cast<size_t>(kCandWarps) * 2 * nstripes * sizeof(uint32_t),
                     static_cast<cudaStream_t>(stream)>>>(
      cand_off, cand_cnt, cand_rows, cand_key, thresh, list_stripes, rows_per_stripe, nstripes, nq, nq_pad,
      sample_stride, cap_dev, m, scut_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_cand_finalize(ch_ws* ws, const ch_cand_args* a, void* stream) {
  if (ws == nullptr) CH_FAIL("null workspace");
  CandDev d;
  if (to_dev(a, &d)) return 1;
  if (a->base0_all == nullptr) CH_FAIL("null bases");
  if (a->mode == 0 && (a->cols == nullptr || a->base0_rel == nullptr)) CH_FAIL("mode 0 needs cols and base0_rel");
  if (a->mode == 1 && a->first_rel_out == nullptr) CH_FAIL("mode 1 needs first_rel_out");
  if (a->mode == 2 && (a->ids == nullptr || a->R <= 0)) CH_FAIL("mode 2 needs ids and R > 0");
  if (a->mode < 0 || a->mode > 2) CH_FAIL("bad mode %d", a->mode);
  ChDeviceGuard guard(ws->device);
  const unsigned blocks = static_cast<unsigned>((a->nq + kCandWarps - 1) / kCandWarps);
  const size_t smem = static_cast<size_t>(kCandWarps) * (2 * a->nbins + 128) * sizeof(uint32_t);   // counters + queue
  cand_final_kernel<<<blocks, kCandWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(d);
  CH_LAUNCH_CHECK(ws);
  return 0;
}
