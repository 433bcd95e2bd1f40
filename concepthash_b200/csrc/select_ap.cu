// K3 (exact, sort-free top-R selection from the key histograms) and K4 (label match / precision@k /
// AP reduction), plus the small helpers around them.  Everything here touches O(nq * nbins) or
// O(#records) data -- tiny next to the nq * ndb pair loop of K2.
#include "common.cuh"

namespace {

// ---- K3: bases + thresholds ---------------------------------------------------------------------
// tot_all: (world, nbins, nq_pad) per-rank key totals.  A query's keys are split among kScanGroups threads (one warp
// = 32 consecutive queries of one key group: coalesced rows): every thread sums its keys, the groups' sums are
// prefixed through shared memory, and a second walk over the same keys (L1 / L2 hits) writes the bases.  One thread
// per query walking all keys left ~1.3 warps per SM: 74 us for 129 keys x 25 000 queries, all of it load latency.
constexpr int kScanGroups = 8;
__global__ void __launch_bounds__(32 * kScanGroups)
scan_bases_kernel(const uint32_t* __restrict__ tot_all, int world, int rank, int nbins, long long nq, long long nq_pad,
                  long long rmax, uint32_t* __restrict__ base0, uint32_t* __restrict__ thresh,
                  uint32_t* __restrict__ total) {
  __shared__ unsigned long long gsum[kScanGroups][32];
  __shared__ uint32_t gfirst[kScanGroups][32];
  const int lane = threadIdx.x & 31, kg = threadIdx.x >> 5;
  const long long q = static_cast<long long>(blockIdx.x) * 32 + lane;
  const bool live = q < nq;                      // (pad queries: zero bases, threshold nbins - 1, like before)
  const int chunk = (nbins + kScanGroups - 1) / kScanGroups;
  const int k0 = kg * chunk, k1 = min(nbins, k0 + chunk);
  constexpr int B = 8;                            // loads of a batch are independent
  unsigned long long sum = 0;
  if (live && q < nq_pad) {
    for (int key0 = k0; key0 < k1; key0 += B) {
      uint32_t v[B];
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
        v[jj] = 0u;
        if (key0 + jj < k1)
          for (int g = 0; g < world; ++g)
            v[jj] += __ldg(tot_all + (static_cast<size_t>(g) * nbins + key0 + jj) * nq_pad + q);
      }
#pragma unroll
      for (int jj = 0; jj < B; ++jj) sum += v[jj];
    }
  }
  gsum[kg][lane] = sum;
  __syncthreads();
  unsigned long long cum = 0;
  for (int g2 = 0; g2 < kg; ++g2) cum += gsum[g2][lane];
  uint32_t first = 0xffffffffu;                   // the first key of this group at which cum >= rmax
  if (q < nq_pad) {
    for (int key0 = k0; key0 < k1; key0 += B) {
      unsigned long long lower[B], tot[B];
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
        lower[jj] = tot[jj] = 0;
        if (live && key0 + jj < k1) {
          for (int g = 0; g < world; ++g) {
            const uint32_t v = __ldg(tot_all + (static_cast<size_t>(g) * nbins + key0 + jj) * nq_pad + q);
            if (g < rank) lower[jj] += v;
            tot[jj] += v;
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
        const int key = key0 + jj;
        if (key >= k1) break;
        base0[static_cast<size_t>(key) * nq_pad + q] = static_cast<uint32_t>(cum + lower[jj]);
        cum += tot[jj];
        if (first == 0xffffffffu && rmax >= 0 && cum >= static_cast<unsigned long long>(rmax))
          first = static_cast<uint32_t>(key);
      }
    }
  }
  gfirst[kg][lane] = first;
  __syncthreads();
  if (kg == 0 && q < nq_pad) {
    uint32_t t = static_cast<uint32_t>(nbins - 1);
    unsigned long long all = 0;
    for (int g2 = kScanGroups - 1; g2 >= 0; --g2) {
      if (gfirst[g2][lane] != 0xffffffffu) t = gfirst[g2][lane];
      all += gsum[g2][lane];
    }
    if (thresh != nullptr) thresh[q] = t;
    if (total != nullptr) total[q] = static_cast<uint32_t>(all);
  }
}

// ---- record capacities / offsets ------------------------------------------------------------------
__global__ void record_caps_kernel(int source, const uint32_t* __restrict__ a0, const uint32_t* __restrict__ a1,
                                   int nstripes, int nb, long long nq, long long nq_pad, int min_with_prev,
                                   int sample_stride, int src_stripes, uint32_t* __restrict__ cap) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(nstripes) * nq_pad) return;
  // src_stripes == 1 < nstripes: the one source stripe is replicated into every output stripe
  const int s = src_stripes == 1 ? 0 : static_cast<int>(i / nq_pad);
  const long long q = i % nq_pad;
  uint32_t c = 0;
  if (q < nq) {
    if (source == 0 || source == 1) {
      const int last = source == 0 ? static_cast<int>(a1[q]) : nb - 1;
      for (int key = 0; key <= last && key < nb; ++key) c += a0[(static_cast<size_t>(s) * nb + key) * nq_pad + q];
    } else {
      const uint32_t id = a1[q];
      if (id < static_cast<uint32_t>(nb)) c = a0[static_cast<size_t>(s) * nb + id];
    }
  }
  if (sample_stride > 1) {
    // c candidates seen in a 1-in-stride row sample -> bound on the real count (Poisson tail, checked later:
    // an overflowing slice raises the error flag and the caller redoes the evaluation exactly)
    const float k = static_cast<float>(c);
    c = static_cast<uint32_t>((k + 6.0f * sqrtf(k + 1.0f) + 9.0f) * static_cast<float>(sample_stride));
  }
  if (min_with_prev) {
    const uint32_t p = cap[i];
    c = p < c ? p : c;
  }
  cap[i] = c;
}

__global__ void row_totals_kernel(const uint32_t* __restrict__ cap, int nstripes, long long nq_pad,
                                  unsigned long long* __restrict__ rowtot) {
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= nq_pad) return;
  unsigned long long s = 0;
  for (int k = 0; k < nstripes; ++k) s += cap[static_cast<size_t>(k) * nq_pad + q];
  rowtot[q] = s;
}

// single-CTA exclusive scan (in place) of n u64 values; result[n] = total
__global__ void __launch_bounds__(1024) exscan_u64_kernel(unsigned long long* __restrict__ v, long long n) {
  __shared__ unsigned long long part[1024];
  const int t = threadIdx.x;
  const long long chunk = (n + 1023) / 1024;
  const long long b = t * chunk;
  long long e = b + chunk;
  if (e > n) e = n;
  unsigned long long s = 0;
  for (long long i = b; i < e; ++i) s += v[i];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; ++i) {
      const unsigned long long x = part[i];
      part[i] = run;
      run += x;
    }
    v[n] = run;
  }
  __syncthreads();
  unsigned long long run = part[t];
  for (long long i = b; i < e; ++i) {
    const unsigned long long x = v[i];
    v[i] = run;
    run += x;
  }
}

// max over the first n entries (single CTA), written next to the scan total
__global__ void __launch_bounds__(1024) max_u32_kernel(const uint32_t* __restrict__ v, long long n,
                                                       unsigned long long* __restrict__ out) {
  __shared__ uint32_t sh[1024];
  uint32_t m = 0;
  for (long long i = threadIdx.x; i < n; i += 1024) m = max(m, v[i]);
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] = max(sh[threadIdx.x], sh[threadIdx.x + w]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

__global__ void record_offsets_kernel(const uint32_t* __restrict__ cap, const unsigned long long* __restrict__ start,
                                      int nstripes, long long nq_pad, uint32_t* __restrict__ off) {
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= nq_pad) return;
  unsigned long long run = start[q];
  for (int k = 0; k < nstripes; ++k) {
    const size_t o = static_cast<size_t>(k) * nq_pad + q;
    off[o] = static_cast<uint32_t>(run);
    run += cap[o];
  }
}

// {all, relevant} totals of the candidate lists in ONE walk: tot (world, 2, nbins, nq_pad).  Besides the two base
// arrays: key_max[q] = smallest key at which the global list holds rmax items (0xfffffffe-like "no limit" = nbins-1
// when it never does), optional total_rel[q], and the verification of a sampled threshold -- status |= 1 when some
// query counted fewer than `need` candidates (replaces a separate ch_check_counts launch).
// (key groups as in scan_bases_kernel: with 8 ranks a thread of the one-thread-per-query form issued 2 x 8 dependent
// rounds of loads per key batch -- 82 us of a 1.6 ms step on 8 GPUs)
__global__ void __launch_bounds__(32 * kScanGroups)
scan_bases_pair_kernel(const uint32_t* __restrict__ tot, int world, int rank, int nbins, long long nq,
                       long long nq_pad, long long rmax, long long need, uint32_t* __restrict__ base0_all,
                       uint32_t* __restrict__ base0_rel, uint32_t* __restrict__ key_max,
                       uint32_t* __restrict__ total_rel, uint32_t* __restrict__ status, uint32_t* __restrict__ bad) {
  __shared__ unsigned long long gsum_a[kScanGroups][32], gsum_r[kScanGroups][32];
  __shared__ uint32_t gfirst[kScanGroups][32];
  const int lane = threadIdx.x & 31, kg = threadIdx.x >> 5;
  const long long q = static_cast<long long>(blockIdx.x) * 32 + lane;
  const bool inside = q < nq_pad, live = q < nq;
  const bool with_rel = base0_rel != nullptr;
  const size_t plane = static_cast<size_t>(nbins) * nq_pad;
  const int chunk = (nbins + kScanGroups - 1) / kScanGroups;
  const int k0 = kg * chunk, k1 = min(nbins, k0 + chunk);
  constexpr int B = 4;
  unsigned long long sum_a = 0, sum_r = 0;
  if (live) {
    for (int key0 = k0; key0 < k1; key0 += B) {
      uint32_t va[B], vr[B];
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
        va[jj] = vr[jj] = 0u;
        if (key0 + jj < k1) {
          for (int g = 0; g < world; ++g) {
            const size_t o = ((static_cast<size_t>(g) * 2) * nbins + key0 + jj) * nq_pad + q;
            va[jj] += __ldg(tot + o);
            if (with_rel) vr[jj] += __ldg(tot + o + plane);
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
        sum_a += va[jj];
        sum_r += vr[jj];
      }
    }
  }
  gsum_a[kg][lane] = sum_a;
  gsum_r[kg][lane] = sum_r;
  __syncthreads();
  unsigned long long cum_a = 0, cum_r = 0;
  for (int g2 = 0; g2 < kg; ++g2) {
    cum_a += gsum_a[g2][lane];
    cum_r += gsum_r[g2][lane];
  }
  uint32_t first = 0xffffffffu;
  if (inside) {
    for (int key0 = k0; key0 < k1; key0 += B) {
      unsigned long long la[B], ta[B], lr[B], tr[B];
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
        la[jj] = ta[jj] = lr[jj] = tr[jj] = 0;
        if (live && key0 + jj < k1) {
          for (int g = 0; g < world; ++g) {
            const size_t o = ((static_cast<size_t>(g) * 2) * nbins + key0 + jj) * nq_pad + q;
            const uint32_t va = __ldg(tot + o);
            const uint32_t vr = with_rel ? __ldg(tot + o + plane) : 0u;
            if (g < rank) {
              la[jj] += va;
              lr[jj] += vr;
            }
            ta[jj] += va;
            tr[jj] += vr;
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
        const int key = key0 + jj;
        if (key >= k1) break;
        const size_t o = static_cast<size_t>(key) * nq_pad + q;
        base0_all[o] = static_cast<uint32_t>(cum_a + la[jj]);
        if (with_rel) base0_rel[o] = static_cast<uint32_t>(cum_r + lr[jj]);
        cum_a += ta[jj];
        cum_r += tr[jj];
        if (first == 0xffffffffu && rmax >= 0 && cum_a >= static_cast<unsigned long long>(rmax))
          first = static_cast<uint32_t>(key);
      }
    }
  }
  gfirst[kg][lane] = first;
  __syncthreads();
  bool short_list = false;
  if (kg == 0 && inside) {
    uint32_t t = static_cast<uint32_t>(nbins - 1);
    unsigned long long all_a = 0, all_r = 0;
    for (int g2 = kScanGroups - 1; g2 >= 0; --g2) {
      if (gfirst[g2][lane] != 0xffffffffu) t = gfirst[g2][lane];
      all_a += gsum_a[g2][lane];
      all_r += gsum_r[g2][lane];
    }
    if (key_max != nullptr) key_max[q] = t;
    if (total_rel != nullptr) total_rel[q] = static_cast<uint32_t>(all_r);
    short_list = need > 0 && live && all_a < static_cast<unsigned long long>(need);
    if (short_list && bad != nullptr) bad[q] = 1u;
  }
  if (status != nullptr && __ballot_sync(0xffffffffu, short_list) != 0u && lane == 0) atomicOr(status, 1u);
}

// every stride-th row of a packed bit plane (rows, words) -> (rows_out_pad, words), pad rows zero
__global__ void gather_rows_kernel(const uint32_t* __restrict__ src, long long n_out, long long rows_out_pad,
                                   long long stride, int words, uint32_t* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows_out_pad * words) return;
  const long long r = i / words;
  const int w = static_cast<int>(i - r * words);
  out[i] = r < n_out ? __ldg(src + (r * stride) * words + w) : 0u;
}

// single-CTA exclusive scan like exscan_u64_kernel, plus what the host used to do after its round trip:
// info[0] = total slots (saturated), info[1] = max(thresh[0..nq)); status |= 4 when the total exceeds `limit`
// (the arena the caller allocated from a hint), |= 8 when a threshold reaches `key_limit` (the narrowed key range).
__global__ void __launch_bounds__(1024) exscan_check_kernel(unsigned long long* __restrict__ v, long long n,
                                                            const uint32_t* __restrict__ thresh, long long nq,
                                                            unsigned long long limit, uint32_t key_limit,
                                                            uint32_t* __restrict__ info, uint32_t* __restrict__ status) {
  __shared__ unsigned long long part[1024];
  __shared__ uint32_t mx[1024];
  const int t = threadIdx.x;
  const long long chunk = (n + 1023) / 1024;
  const long long b = t * chunk;
  long long e = b + chunk;
  if (e > n) e = n;
  unsigned long long s = 0;
  for (long long i = b; i < e; ++i) s += v[i];
  part[t] = s;
  uint32_t m = 0;
  if (thresh != nullptr)
    for (long long i = t; i < nq; i += 1024) m = max(m, thresh[i]);
  mx[t] = m;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (t < w) mx[t] = max(mx[t], mx[t + w]);
    __syncthreads();
  }
  // exclusive scan of the 1024 partial sums: shuffles within a warp, the 32 warp totals by warp 0 (a single thread
  // walking all 1024 took ~16 us of this kernel's 33)
  __shared__ unsigned long long wsum[33];
  {
    const int lane = t & 31, wp = t >> 5;
    const unsigned long long x = part[t];
    unsigned long long inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    if (lane == 31) wsum[wp] = inc;
    __syncthreads();
    if (wp == 0) {
      const unsigned long long w = wsum[lane];
      unsigned long long winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += y;
      }
      wsum[lane] = winc - w;                     // exclusive prefix of the warp totals
      if (lane == 31) wsum[32] = winc;           // the grand total
    }
    __syncthreads();
    part[t] = wsum[wp] + inc - x;
  }
  if (t == 0) {
    const unsigned long long run = wsum[32];
    v[n] = run;
    if (info != nullptr) {
      info[0] = run > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(run);
      info[1] = mx[0];
    }
    uint32_t fl = 0;
    if (run > limit) fl |= 4u;
    if (thresh != nullptr && key_limit != 0u && mx[0] >= key_limit) fl |= 8u;
    if (fl != 0u && status != nullptr) atomicOr(status, fl);
  }
  __syncthreads();
  unsigned long long run = part[t];
  // (batches of 8: a load of v[i + 1] may not pass the store to v[i] otherwise, and 25 dependent round trips of one
  // thread were half of this single-CTA kernel's time)
  for (long long i = b; i < e; i += 8) {
    unsigned long long x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = i + j < e ? v[i + j] : 0ull;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (i + j < e) v[i + j] = run;
      run += x[j];
    }
  }
}

// offsets as record_offsets_kernel; slices that would leave the arena get capacity 0 (the overflow is already
// flagged: nothing may be written there)
__global__ void record_offsets_clamp_kernel(uint32_t* __restrict__ cap, const unsigned long long* __restrict__ start,
                                            int nstripes, long long nq_pad, unsigned long long limit,
                                            uint32_t* __restrict__ off) {
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= nq_pad) return;
  unsigned long long run = start[q];
  for (int k = 0; k < nstripes; ++k) {
    const size_t o = static_cast<size_t>(k) * nq_pad + q;
    const uint32_t c = cap[o];
    if (run + c > limit) {
      cap[o] = 0u;
      off[o] = 0u;
    } else {
      off[o] = static_cast<uint32_t>(run);
      run += c;
    }
  }
}

// ---- K4: records -> AP sums / hit counts ------------------------------------------------------------
struct FinalDev {
  const uint4* recs; const uint32_t* rec_off; const uint32_t* rec_cnt;
  const uint32_t* base0_all; const uint32_t* base0_rel;
  const uint32_t* sbase_all; const uint32_t* sbase_rel;
  const uint32_t* first_rel;
  double* partial;
  long long nq, nq_pad;
  int nstripes, nbins, remove_first, nR, nPR;
  long long r_eff[CH_MAX_R];
  long long pr_k[CH_MAX_PR];
};

__device__ __forceinline__ bool record_rank(const FinalDev& a, const uint4 r, int s, long long q, bool need_rel,
                                            long long* rank, long long* relrank) {
  const uint32_t key = r.x & 0x7fffffffu;
  const size_t o0 = static_cast<size_t>(key) * a.nq_pad + q;
  const size_t os = (static_cast<size_t>(s) * a.nbins + key) * a.nq_pad + q;
  *rank = static_cast<long long>(a.base0_all[o0]) + a.sbase_all[os] + r.y;
  if (need_rel) *relrank = static_cast<long long>(a.base0_rel[o0]) + a.sbase_rel[os] + r.z;
  return (r.x >> 31) != 0u;
}

__global__ void finalize_kernel(const FinalDev a) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(a.nstripes) * a.nq_pad) return;
  const int s = static_cast<int>(i / a.nq_pad);
  const long long q = i - static_cast<long long>(s) * a.nq_pad;
  if (q >= a.nq) return;
  const int ncols = 2 * a.nR + a.nPR;
  double acc[2 * CH_MAX_R + CH_MAX_PR];
  for (int c = 0; c < ncols; ++c) acc[c] = 0.0;
  const uint32_t b = a.rec_off[i], n = a.rec_cnt[i];
  const long long shift = a.remove_first ? 1 : 0;
  const long long frel = (a.remove_first && a.first_rel != nullptr) ? a.first_rel[q] : 0;
  for (uint32_t k = 0; k < n; ++k) {
    const uint4 r = a.recs[b + k];
    long long rank, relrank = 0;
    if (!record_rank(a, r, s, q, true, &rank, &relrank)) continue;  // candidate records may be non-relevant
    if (a.remove_first) {
      if (rank == 0) continue;  // the dropped self-retrieval
      rank -= shift;
      relrank -= frel;
    }
    const double prec = static_cast<double>(relrank + 1) / static_cast<double>(rank + 1);
    for (int j = 0; j < a.nR; ++j)
      if (rank < a.r_eff[j]) {
        acc[2 * j] += prec;
        acc[2 * j + 1] += 1.0;
      }
    for (int j = 0; j < a.nPR; ++j)
      if (rank < a.pr_k[j]) acc[2 * a.nR + j] += 1.0;
  }
  double* out = a.partial + static_cast<size_t>(i) * ncols;
  for (int c = 0; c < ncols; ++c) out[c] = acc[c];
}

// cols[q][c] = sum over stripes (fixed order -> deterministic)
__global__ void reduce_stripes_kernel(const double* __restrict__ partial, int nstripes, long long nq,
                                      long long nq_pad, int ncols, double* __restrict__ cols) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nq * ncols) return;
  const long long q = i / ncols;
  const int c = static_cast<int>(i - q * ncols);
  double s = 0.0;
  for (int k = 0; k < nstripes; ++k) s += partial[(static_cast<size_t>(k) * nq_pad + q) * ncols + c];
  cols[i] = s;
}

__global__ void first_relevant_kernel(const FinalDev a, uint32_t* __restrict__ first_rel) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(a.nstripes) * a.nq_pad) return;
  const int s = static_cast<int>(i / a.nq_pad);
  const long long q = i - static_cast<long long>(s) * a.nq_pad;
  if (q >= a.nq) return;
  const uint32_t b = a.rec_off[i], n = a.rec_cnt[i];
  for (uint32_t k = 0; k < n; ++k) {
    long long rank, relrank;
    if (record_rank(a, a.recs[b + k], s, q, false, &rank, &relrank) && rank == 0) first_rel[q] = 1u;
  }
}

__global__ void scatter_ranked_kernel(const FinalDev a, long long R, long long row_offset, long long* __restrict__ ids,
                                      int* __restrict__ keys) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(a.nstripes) * a.nq_pad) return;
  const int s = static_cast<int>(i / a.nq_pad);
  const long long q = i - static_cast<long long>(s) * a.nq_pad;
  if (q >= a.nq) return;
  const uint32_t b = a.rec_off[i], n = a.rec_cnt[i];
  for (uint32_t k = 0; k < n; ++k) {
    const uint4 r = a.recs[b + k];
    long long rank, relrank;
    record_rank(a, r, s, q, false, &rank, &relrank);
    if (a.remove_first) {
      if (rank == 0) continue;
      rank -= 1;
    }
    if (rank < R) {
      ids[q * R + rank] = row_offset + r.w;
      if (keys != nullptr) keys[q * R + rank] = static_cast<int>(r.x & 0x7fffffffu);
    }
  }
}

// verification of the sampled threshold: flag bit 0 if some query counted fewer than `need` candidates
__global__ void check_counts_kernel(const uint32_t* __restrict__ total, long long nq, uint32_t need,
                                    uint32_t* __restrict__ flags) {
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool bad = q < nq && total[q] < need;
  if (__ballot_sync(0xffffffffu, bad) != 0u && (threadIdx.x & 31) == 0) atomicOr(flags, 1u);
}

// deterministic block sum of per-query values; one block per output scalar
__device__ double block_sum_1024(double v, double* sh) {
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (t < w) sh[t] += sh[t + w];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

struct PrK {
  long long k[CH_MAX_PR];   // by value in the kernel parameters: nothing to copy, nothing a captured graph could lose
};

// out[0..nR) = mAP_i, out[nR..nR+nPR) = recall@k, out[nR+nPR..nR+2nPR) = precision@k
__global__ void __launch_bounds__(1024) reduce_means_kernel(const double* __restrict__ cols,
                                                            const uint32_t* __restrict__ total_rel,
                                                            const uint32_t* __restrict__ first_rel, long long nq,
                                                            int nR, int nPR, const PrK pr_k,
                                                            double* __restrict__ ap_out, double* __restrict__ out) {
  __shared__ double sh[1024];
  const int ncols = 2 * nR + nPR;
  const int which = blockIdx.x;
  double s = 0.0;
  for (long long q = threadIdx.x; q < nq; q += 1024) {
    const double* c = cols + q * ncols;
    double v;
    if (which < nR) {
      v = c[2 * which + 1] > 0.0 ? c[2 * which] / c[2 * which + 1] : 0.0;
      if (ap_out != nullptr) ap_out[static_cast<size_t>(which) * nq + q] = v;
    } else if (which < nR + nPR) {
      const int j = which - nR;
      double tr = total_rel != nullptr ? static_cast<double>(total_rel[q]) : 0.0;
      if (first_rel != nullptr) tr -= static_cast<double>(first_rel[q]);
      v = c[2 * nR + j] / (tr > 1.0 ? tr : 1.0);
    } else {
      const int j = which - nR - nPR;
      v = c[2 * nR + j] / static_cast<double>(pr_k.k[j]);
    }
    s += v;
  }
  const double tot = block_sum_1024(s, sh);
  if (threadIdx.x == 0) out[which] = nq > 0 ? tot / static_cast<double>(nq) : 0.0;
}

// ---- K4 (list form): AP straight from a ranked id list, one warp per query -------------------------
__global__ void ap_from_ranked_kernel(const long long* __restrict__ ids, long long nq, long long R,
                                      const uint32_t* __restrict__ q_lab, const uint32_t* __restrict__ g_lab,
                                      int label_mode, int lw, int nPR, const long long* __restrict__ pr_k,
                                      double* __restrict__ cols) {
  const int lane = threadIdx.x & 31;
  const long long q = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  const int ncols = 2 + nPR;
  double sum = 0.0;
  uint32_t run = 0;
  uint32_t hits[CH_MAX_PR];
  for (int j = 0; j < nPR; ++j) hits[j] = 0;
  const uint32_t qid = label_mode == CH_LAB_ID ? q_lab[q] : 0u;
  for (long long k0 = 0; k0 < R; k0 += 32) {
    const long long k = k0 + lane;
    bool rel = false;
    if (k < R) {
      const long long id = ids[q * R + k];
      if (id >= 0) {
        if (label_mode == CH_LAB_ID) {
          rel = g_lab[id] == qid;
        } else {
          uint32_t any = 0;
          for (int w = 0; w < lw; ++w) any |= q_lab[q * lw + w] & g_lab[id * lw + w];
          rel = any != 0u;
        }
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, rel);
    if (rel) {
      const uint32_t before = run + __popc(m & lanemask_lt());
      sum += static_cast<double>(before + 1) / static_cast<double>(k + 1);
      for (int j = 0; j < nPR; ++j)
        if (k < pr_k[j]) hits[j] += 1;
    }
    run += __popc(m);
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int j = 0; j < nPR; ++j) hits[j] = __reduce_add_sync(0xffffffffu, hits[j]);
  if (lane == 0) {
    double* c = cols + q * ncols;
    c[0] = sum;
    c[1] = static_cast<double>(run);
    for (int j = 0; j < nPR; ++j) c[2 + j] = static_cast<double>(hits[j]);
  }
}

// ---- dense key matrix (tests, get_hamm_dist) ---------------------------------------------------------
__global__ void hamming_matrix_kernel(const uint32_t* __restrict__ qb, const uint32_t* __restrict__ qz,
                                      const uint32_t* __restrict__ gb, const uint32_t* __restrict__ gz, long long nq,
                                      long long ndb, int nbit, int nw, int tern, uint16_t* __restrict__ out) {
  const long long total = nq * ndb;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long q = i / ndb, j = i - q * ndb;
    uint32_t key;
    if (!tern) {
      key = 0;
      for (int w = 0; w < nw; ++w) key += __popc(qb[q * nw + w] ^ gb[j * nw + w]);
    } else {
      uint32_t both = 0, dis = 0;
      for (int w = 0; w < nw; ++w) {
        const uint32_t m = qz[q * nw + w] & gz[j * nw + w];
        both += __popc(m);
        dis += __popc((qb[q * nw + w] ^ gb[j * nw + w]) & m);
      }
      key = static_cast<uint32_t>(nbit) - both + 2u * dis;
    }
    out[i] = static_cast<uint16_t>(key);
  }
}

// ---- integer-pipe micro-benchmark ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) popc_peak_kernel(uint32_t* __restrict__ out, int iters, uint32_t seed) {
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u;
  uint32_t a4 = a0 * 11u, a5 = a0 * 13u, a6 = a0 * 17u, a7 = a0 * 19u;
  const uint32_t c = seed * 0x9E3779B9u + blockIdx.x;
  for (int i = 0; i < iters; ++i) {
    // 8 independent chains; each step = 1 LOP3 (alu pipe) + 1 POPC
    a0 = __popc(a0 ^ c) ^ a1; a1 = __popc(a1 ^ c) ^ a2; a2 = __popc(a2 ^ c) ^ a3; a3 = __popc(a3 ^ c) ^ a4;
    a4 = __popc(a4 ^ c) ^ a5; a5 = __popc(a5 ^ c) ^ a6; a6 = __popc(a6 ^ c) ^ a7; a7 = __popc(a7 ^ c) ^ a0;
  }
  const uint32_t r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 0xdeadbeefu) out[0] = r;  // never true in practice; keeps the loop alive
}

FinalDev to_dev(const ch_final_args* a) {
  FinalDev d;
  d.recs = static_cast<const uint4*>(a->recs); d.rec_off = a->rec_off; d.rec_cnt = a->rec_cnt;
  d.base0_all = a->base0_all; d.base0_rel = a->base0_rel; d.sbase_all = a->sbase_all; d.sbase_rel = a->sbase_rel;
  d.first_rel = a->first_rel; d.partial = a->partial; d.nq = a->nq; d.nq_pad = a->nq_pad;
  d.nstripes = a->nstripes; d.nbins = a->nbins; d.remove_first = a->remove_first; d.nR = a->nR; d.nPR = a->nPR;
  for (int i = 0; i < CH_MAX_R; ++i) d.r_eff[i] = i < a->nR ? a->r_eff[i] : 0;
  for (int i = 0; i < CH_MAX_PR; ++i) d.pr_k[i] = i < a->nPR ? a->pr_k[i] : 0;
  return d;
}

int check_final(const ch_final_args* a, bool need_rel) {
  if (a == nullptr) CH_FAIL("null ch_final_args");
  if (a->nR < 0 || a->nR > CH_MAX_R || a->nPR < 0 || a->nPR > CH_MAX_PR) CH_FAIL("too many R / PRs entries");
  if (a->recs == nullptr || a->rec_off == nullptr || a->rec_cnt == nullptr || a->base0_all == nullptr ||
      a->sbase_all == nullptr)
    CH_FAIL("null record / base arrays");
  if (need_rel && (a->base0_rel == nullptr || a->sbase_rel == nullptr)) CH_FAIL("null relevant base arrays");
  return 0;
}

unsigned blocks_for(long long n, int t) { return static_cast<unsigned>((n + t - 1) / t); }

}  // namespace

extern "C" int ch_scan_bases(ch_ws* ws, const uint32_t* tot_all_dev, int world, int rank, int nbins, int64_t nq,
                             int64_t nq_pad, int64_t rmax, uint32_t* base0_dev, uint32_t* thresh_out_dev,
                             uint32_t* total_out_dev, void* stream) {
  if (ws == nullptr || tot_all_dev == nullptr || base0_dev == nullptr) CH_FAIL("null argument to ch_scan_bases");
  if (world < 1 || rank < 0 || rank >= world) CH_FAIL("bad world/rank %d/%d", world, rank);
  ChDeviceGuard guard(ws->device);
  scan_bases_kernel<<<blocks_for(nq_pad, 32), 32 * kScanGroups, 0, static_cast<cudaStream_t>(stream)>>>(
      tot_all_dev, world, rank, nbins, nq, nq_pad, rmax, base0_dev, thresh_out_dev, total_out_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_record_caps(ch_ws* ws, int source, const uint32_t* slab_or_cls, const uint32_t* thresh_or_qids,
                              int nstripes, int nbins_or_nclass, int64_t nq, int64_t nq_pad, int min_with_prev,
                              int sample_stride, int src_stripes, uint32_t* cap_dev, void* stream) {
  if (ws == nullptr || slab_or_cls == nullptr || cap_dev == nullptr) CH_FAIL("null argument to ch_record_caps");
  if (source < 0 || source > 2) CH_FAIL("bad capacity source %d", source);
  if (source != 1 && thresh_or_qids == nullptr) CH_FAIL("capacity source %d needs thresholds / query ids", source);
  if (src_stripes != nstripes && src_stripes != 1) CH_FAIL("src_stripes must be nstripes or 1 (replicate)");
  ChDeviceGuard guard(ws->device);
  record_caps_kernel<<<blocks_for(static_cast<long long>(nstripes) * nq_pad, 256), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(source, slab_or_cls, thresh_or_qids, nstripes,
                                                            nbins_or_nclass, nq, nq_pad, min_with_prev, sample_stride,
                                                            src_stripes, cap_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

int ch_ws_scratch(ch_ws* ws, size_t bytes, void** out);  // api.cu

extern "C" int ch_record_offsets(ch_ws* ws, const uint32_t* cap_dev, int nstripes, int64_t nq, int64_t nq_pad,
                                 uint32_t* off_dev, uint64_t* total_host, const uint32_t* thresh_dev,
                                 uint32_t* thresh_max_host, void* stream) {
  if (ws == nullptr || cap_dev == nullptr || off_dev == nullptr || total_host == nullptr)
    CH_FAIL("null argument to ch_record_offsets");
  (void)nq;
  ChDeviceGuard guard(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* scratch = nullptr;
  if (ch_ws_scratch(ws, static_cast<size_t>(nq_pad + 2) * 8, &scratch)) return 1;
  unsigned long long* rowtot = static_cast<unsigned long long*>(scratch);
  row_totals_kernel<<<blocks_for(nq_pad, 256), 256, 0, st>>>(cap_dev, nstripes, nq_pad, rowtot);
  CH_LAUNCH_CHECK(ws);
  exscan_u64_kernel<<<1, 1024, 0, st>>>(rowtot, nq_pad);
  CH_LAUNCH_CHECK(ws);
  record_offsets_kernel<<<blocks_for(nq_pad, 256), 256, 0, st>>>(cap_dev, rowtot, nstripes, nq_pad, off_dev);
  CH_LAUNCH_CHECK(ws);
  if (thresh_dev != nullptr) {   // the same round trip also brings back max(thresh) (-> key_limit)
    max_u32_kernel<<<1, 1024, 0, st>>>(thresh_dev, nq, rowtot + nq_pad + 1);
    CH_LAUNCH_CHECK(ws);
  }
  unsigned long long both[2] = {0, 0};
  CH_CUDA(cudaMemcpyAsync(both, rowtot + nq_pad, thresh_dev != nullptr ? 16 : 8, cudaMemcpyDeviceToHost, st));
  CH_CUDA(cudaStreamSynchronize(st));
  const unsigned long long total = both[0];
  *total_host = total;
  if (thresh_max_host != nullptr) *thresh_max_host = static_cast<uint32_t>(both[1]);
  // (a total beyond the 32-bit slot index is the CALLER's to handle -- it evaluates the query set in chunks; the
  // offsets written above are meaningless then and must not be used)
  return 0;
}

extern "C" int ch_record_offsets_async(ch_ws* ws, uint32_t* cap_dev, int nstripes, int64_t nq, int64_t nq_pad,
                                       uint32_t* off_dev, const uint32_t* thresh_dev, uint64_t limit_slots,
                                       uint32_t key_limit, uint32_t* info_dev, uint32_t* status_dev, void* stream) {
  if (ws == nullptr || cap_dev == nullptr || off_dev == nullptr || status_dev == nullptr)
    CH_FAIL("null argument to ch_record_offsets_async");
  if (limit_slots == 0 || limit_slots >= 0xffffffffull) CH_FAIL("limit_slots must be in 1 .. 2^32 - 2");
  ChDeviceGuard guard(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* scratch = nullptr;
  if (ch_ws_scratch(ws, static_cast<size_t>(nq_pad + 2) * 8, &scratch)) return 1;
  unsigned long long* rowtot = static_cast<unsigned long long*>(scratch);
  row_totals_kernel<<<blocks_for(nq_pad, 256), 256, 0, st>>>(cap_dev, nstripes, nq_pad, rowtot);
  CH_LAUNCH_CHECK(ws);
  exscan_check_kernel<<<1, 1024, 0, st>>>(rowtot, nq_pad, thresh_dev, nq, limit_slots, key_limit, info_dev, status_dev);
  CH_LAUNCH_CHECK(ws);
  record_offsets_clamp_kernel<<<blocks_for(nq_pad, 256), 256, 0, st>>>(cap_dev, rowtot, nstripes, nq_pad, limit_slots,
                                                                       off_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_scan_bases_pair(ch_ws* ws, const uint32_t* tot_dev, int world, int rank, int nbins, int64_t nq,
                                  int64_t nq_pad, int64_t rmax, int64_t need, uint32_t* base0_all_dev,
                                  uint32_t* base0_rel_dev, uint32_t* key_max_dev, uint32_t* total_rel_dev,
                                  uint32_t* status_dev, uint32_t* bad_dev, void* stream) {
  if (ws == nullptr || tot_dev == nullptr || base0_all_dev == nullptr) CH_FAIL("null argument to ch_scan_bases_pair");
  if (world < 1 || rank < 0 || rank >= world) CH_FAIL("bad world/rank %d/%d", world, rank);
  if (need > 0 && status_dev == nullptr) CH_FAIL("verification needs a status word");
  ChDeviceGuard guard(ws->device);
  scan_bases_pair_kernel<<<blocks_for(nq_pad, 32), 32 * kScanGroups, 0, static_cast<cudaStream_t>(stream)>>>(
      tot_dev, world, rank, nbins, nq, nq_pad, rmax, need, base0_all_dev, base0_rel_dev, key_max_dev, total_rel_dev,
      status_dev, bad_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_gather_rows(ch_ws* ws, const uint32_t* bits_dev, int64_t n_src, int nbit, int64_t stride,
                              uint32_t* out_dev, int64_t rows_out_pad, void* stream) {
  if (ws == nullptr || bits_dev == nullptr || out_dev == nullptr) CH_FAIL("null argument to ch_gather_rows");
  const int words = ch_code_words(nbit);
  if (words == 0 || stride < 1 || n_src < 0) CH_FAIL("bad arguments to ch_gather_rows");
  const int64_t n_out = (n_src + stride - 1) / stride;
  if (rows_out_pad < n_out) CH_FAIL("output holds %lld rows, the sample has %lld", (long long)rows_out_pad, (long long)n_out);
  if (rows_out_pad == 0) return 0;
  ChDeviceGuard guard(ws->device);
  gather_rows_kernel<<<blocks_for(rows_out_pad * words, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bits_dev, n_out, rows_out_pad, stride, words, out_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_finalize_records(ch_ws* ws, const ch_final_args* a, void* stream) {
  if (ws == nullptr) CH_FAIL("null workspace");
  if (check_final(a, true)) return 1;
  if (a->partial == nullptr || a->cols == nullptr) CH_FAIL("null partial / cols");
  ChDeviceGuard guard(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const FinalDev d = to_dev(a);
  const int ncols = 2 * a->nR + a->nPR;
  // (no memset of `partial`: finalize_kernel writes every (stripe, query < nq) row, reduce_stripes reads only those)
  finalize_kernel<<<blocks_for(static_cast<long long>(a->nstripes) * a->nq_pad, 128), 128, 0, st>>>(d);
  CH_LAUNCH_CHECK(ws);
  reduce_stripes_kernel<<<blocks_for(a->nq * ncols, 256), 256, 0, st>>>(a->partial, a->nstripes, a->nq, a->nq_pad,
                                                                         ncols, a->cols);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_first_relevant(ch_ws* ws, const ch_final_args* a, uint32_t* first_rel_dev, void* stream) {
  if (ws == nullptr || first_rel_dev == nullptr) CH_FAIL("null argument to ch_first_relevant");
  if (check_final(a, false)) return 1;
  ChDeviceGuard guard(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const FinalDev d = to_dev(a);
  first_relevant_kernel<<<blocks_for(static_cast<long long>(a->nstripes) * a->nq_pad, 128), 128, 0, st>>>(
      d, first_rel_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_check_counts(ch_ws* ws, const uint32_t* total_dev, int64_t nq, int64_t need, uint32_t* flags_dev,
                               void* stream) {
  if (ws == nullptr || total_dev == nullptr || flags_dev == nullptr) CH_FAIL("null argument to ch_check_counts");
  ChDeviceGuard guard(ws->device);
  check_counts_kernel<<<blocks_for(nq, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      total_dev, nq, static_cast<uint32_t>(need < 0 ? 0 : need), flags_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

int ch_ws_results(ch_ws* ws, void** dev, void** host);  // api.cu

// The means and the status block are written to the workspace's result words and copied to their pinned host mirror,
// all on `stream` and without waiting: the sequence can be captured in a CUDA graph.  ch_reduce_means_fetch waits for
// the stream and hands the values out.  Result layout: doubles [0, nR + 2 nPR), then u32 flags from byte 512.
extern "C" int ch_reduce_means_enqueue(ch_ws* ws, const double* cols_dev, const uint32_t* total_rel_dev,
                                       const uint32_t* first_rel_dev, int64_t nq, int nR, int nPR, const int64_t* pr_k,
                                       double* ap_out_dev, const uint32_t* flags_dev, int nflags, void* stream) {
  if (ws == nullptr || cols_dev == nullptr) CH_FAIL("null argument to ch_reduce_means");
  if (nR < 0 || nR > CH_MAX_R || nPR < 0 || nPR > CH_MAX_PR) CH_FAIL("too many R / PRs entries");
  if (nflags < 0 || nflags > 128) CH_FAIL("at most 128 status words");
  const int nout = nR + 2 * nPR;
  ChDeviceGuard guard(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void *dev = nullptr, *host = nullptr;
  if (ch_ws_results(ws, &dev, &host)) return 1;
  // the caller's status block (overflow / verification flags, packing statistics, slot totals) rides on the same
  // host sync
  if (flags_dev != nullptr && nflags > 0)
    CH_CUDA(cudaMemcpyAsync(static_cast<char*>(host) + 512, flags_dev, static_cast<size_t>(nflags) * 4,
                            cudaMemcpyDeviceToHost, st));
  if (nout == 0) return 0;
  PrK prk;
  for (int i = 0; i < CH_MAX_PR; ++i) prk.k[i] = i < nPR ? pr_k[i] : 1;
  double* out_dev = static_cast<double*>(dev);
  reduce_means_kernel<<<nout, 1024, 0, st>>>(cols_dev, total_rel_dev, first_rel_dev, nq, nR, nPR, prk, ap_out_dev,
                                             out_dev);
  CH_LAUNCH_CHECK(ws);
  CH_CUDA(cudaMemcpyAsync(host, out_dev, static_cast<size_t>(nout) * 8, cudaMemcpyDeviceToHost, st));
  return 0;
}

extern "C" int ch_reduce_means_fetch(ch_ws* ws, double* out_host, int nout, uint32_t* flags_host, int nflags,
                                     void* stream) {
  if (ws == nullptr) CH_FAIL("null workspace");
  if (nout < 0 || nout > CH_MAX_R + 2 * CH_MAX_PR || nflags < 0 || nflags > 128) CH_FAIL("bad result counts");
  ChDeviceGuard guard(ws->device);
  void *dev = nullptr, *host = nullptr;
  if (ch_ws_results(ws, &dev, &host)) return 1;
  CH_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  if (out_host != nullptr && nout > 0) memcpy(out_host, host, static_cast<size_t>(nout) * 8);
  if (flags_host != nullptr && nflags > 0) memcpy(flags_host, static_cast<char*>(host) + 512, static_cast<size_t>(nflags) * 4);
  return 0;
}

extern "C" int ch_reduce_means(ch_ws* ws, const double* cols_dev, const uint32_t* total_rel_dev,
                               const uint32_t* first_rel_dev, int64_t nq, int nR, int nPR, const int64_t* pr_k,
                               double* ap_out_dev, double* out_host, const uint32_t* flags_dev, uint32_t* flags_host,
                               int nflags, void* stream) {
  if (out_host == nullptr) CH_FAIL("null argument to ch_reduce_means");
  if (ch_reduce_means_enqueue(ws, cols_dev, total_rel_dev, first_rel_dev, nq, nR, nPR, pr_k, ap_out_dev, flags_dev,
                              flags_host != nullptr ? nflags : 0, stream))
    return 1;
  return ch_reduce_means_fetch(ws, out_host, nR + 2 * nPR, flags_host, flags_host != nullptr && flags_dev != nullptr ? nflags : 0,
                               stream);
}

extern "C" int ch_scatter_ranked(ch_ws* ws, const ch_final_args* a, int64_t R, int64_t row_offset, int64_t* ids_dev,
                                 int32_t* keys_dev, void* stream) {
  if (ws == nullptr || ids_dev == nullptr) CH_FAIL("null argument to ch_scatter_ranked");
  if (check_final(a, false)) return 1;
  ChDeviceGuard guard(ws->device);
  const FinalDev d = to_dev(a);
  scatter_ranked_kernel<<<blocks_for(static_cast<long long>(a->nstripes) * a->nq_pad, 128), 128, 0,
                          static_cast<cudaStream_t>(stream)>>>(d, R, row_offset,
                                                               reinterpret_cast<long long*>(ids_dev), keys_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_ap_from_ranked(ch_ws* ws, const int64_t* ids_dev, int64_t nq, int64_t R, const uint32_t* q_lab,
                                 const uint32_t* g_lab, int label_mode, int mask_words, int nPR,
                                 const int64_t* pr_k, double* cols_dev, void* stream) {
  if (ws == nullptr || ids_dev == nullptr || q_lab == nullptr || g_lab == nullptr || cols_dev == nullptr)
    CH_FAIL("null argument to ch_ap_from_ranked");
  if (label_mode != CH_LAB_ID && label_mode != CH_LAB_MASK) CH_FAIL("bad label mode %d", label_mode);
  if (nPR < 0 || nPR > CH_MAX_PR) CH_FAIL("too many PRs entries");
  ChDeviceGuard guard(ws->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* scratch = nullptr;
  if (ch_ws_scratch(ws, CH_MAX_PR * 8, &scratch)) return 1;
  long long* prk_dev = static_cast<long long*>(scratch);
  long long prk_host[CH_MAX_PR];
  for (int i = 0; i < CH_MAX_PR; ++i) prk_host[i] = i < nPR ? pr_k[i] : 1;
  CH_CUDA(cudaMemcpyAsync(prk_dev, prk_host, sizeof(prk_host), cudaMemcpyHostToDevice, st));
  ap_from_ranked_kernel<<<blocks_for(nq * 32, 256), 256, 0, st>>>(reinterpret_cast<const long long*>(ids_dev), nq, R,
                                                                  q_lab, g_lab, label_mode, mask_words, nPR, prk_dev,
                                                                  cols_dev);
  CH_LAUNCH_CHECK(ws);
  // prk_host lives on this stack frame: make sure the copy has been consumed
  CH_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int ch_hamming_matrix(ch_ws* ws, const uint32_t* q_bits, const uint32_t* q_nz, const uint32_t* g_bits,
                                 const uint32_t* g_nz, int64_t nq, int64_t ndb, int nbit, int ternary,
                                 uint16_t* out_dev, void* stream) {
  if (ws == nullptr || q_bits == nullptr || g_bits == nullptr || out_dev == nullptr)
    CH_FAIL("null argument to ch_hamming_matrix");
  if (ternary && (q_nz == nullptr || g_nz == nullptr)) CH_FAIL("ternary codes need the non-zero planes");
  const int nw = ch_code_words(nbit);
  if (nw == 0) CH_FAIL("nbit=%d unsupported", nbit);
  if (nq * ndb == 0) return 0;
  ChDeviceGuard guard(ws->device);
  long long blocks = (nq * ndb + 255) / 256;
  if (blocks > ws->sm_count * 32ll) blocks = ws->sm_count * 32ll;
  hamming_matrix_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      q_bits, q_nz, g_bits, g_nz, nq, ndb, nbit, nw, ternary, out_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_popc_peak(ch_ws* ws, double* popc32_per_s, double* elapsed_ms) {
  if (ws == nullptr || popc32_per_s == nullptr) CH_FAIL("null argument to ch_popc_peak");
  ChDeviceGuard guard(ws->device);
  void* scratch = nullptr;
  if (ch_ws_scratch(ws, 64, &scratch)) return 1;
  cudaEvent_t e0, e1;
  CH_CUDA(cudaEventCreate(&e0));
  CH_CUDA(cudaEventCreate(&e1));
  const int iters = 4096, blocks = ws->sm_count * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CH_CUDA(cudaEventRecord(e0, 0));
    popc_peak_kernel<<<blocks, 256>>>(static_cast<uint32_t*>(scratch), iters, 12345u + rep);
    CH_LAUNCH_CHECK(ws);
    CH_CUDA(cudaEventRecord(e1, 0));
    CH_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    CH_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double ops = static_cast<double>(blocks) * 256.0 * iters * 8.0;
  *popc32_per_s = ops / (best * 1e-3);
  if (elapsed_ms) *elapsed_ms = best;
  return 0;
}
