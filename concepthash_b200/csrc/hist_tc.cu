// K2, tensor-core form of the SELECT pass (pairs with key <= threshold are counted / matched / recorded).
//
// For +-1 codes held as int8, <q, g> = nbit - 2 * hamming(q, g) exactly, so one tcgen05.mma kind::i8
// (UTCIMMA, int32 accumulators in TMEM) produces the distances of a 128-query x 256-gallery-row tile at a cost
// of K/32 instructions -- the 2 LOP3 + 2 POPC + IADD per 64 bits of the integer-pipe kernel disappear, and with
// them the POPC-pipe roofline.  What remains per pair in the select pass is ONE comparison against the query's
// threshold (dot >= nbit - 2 * thresh); the rare candidates then take exactly the same path as in hist.cu
// (label match, private shared-memory histogram in gallery-row order, 16-byte record), so the outputs are
// bit-identical to the integer-pipe kernel.
//
// Warp roles (256 threads, one CTA per SM, persistent over the tiles of one (query tile, stripe)):
//   warp 0   producer: 1-D bulk async copies (UBLKCP) of gallery tiles + their class ids into a 4-stage ring
//   warp 1   MMA issuer: one elected thread issues K/32 UTCIMMA per tile into one of two TMEM accumulators
//   warp 2   TMEM allocator (512 columns = 2 x (128 lanes x 256 columns x s32))
//   warps 4-7 epilogue: tcgen05.ld 32 columns at a time; thread = TMEM lane = query; rows consumed in order
// Operands live in shared memory in the canonical NO-SWIZZLE K-major core-matrix layout (8 rows x 16 bytes =
// 128 contiguous bytes; next 16-byte K chunk at +128 B (LBO); next 8-row group at +8*KB (SBO)).  The int8 planes
// are stored in HBM already in that order (expand_i8_tiled_kernel), so a tile is one contiguous bulk copy.
#include "common.cuh"

namespace {

constexpr int kTileN = 128;   // gallery rows per MMA tile (TMEM columns per accumulator)
constexpr int kTileM = 128;   // queries per CTA

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version 1 (Blackwell); SWIZZLE_NONE, base offset 0
  return d;
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// single-thread waiters (producer, MMA issuer) back off so that their spin does not steal issue slots from
// the epilogue warp sharing the scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(128);
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      "tcgen05.wait::ld.sync.aligned;\n"     // same asm statement: no use of r[] can be scheduled before the wait
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

struct TcDev {
  HistDev h;
  const int8_t* q_i8;   // tiled (>= nq_pad rows, KB)
  const int8_t* g_i8;   // tiled (rows_pad, KB)
  int kbins;            // histogram keys actually reachable in this pass (max threshold + 1)
  int nqtiles128;       // 128-query tiles in total
  int nqgroups;         // CTAs along the query axis (each owns QT consecutive query tiles)
};

// shared-memory carve-up (bytes), all offsets multiples of 128
template <int KB, int QT, int kTcStages>
struct TcSmem {
  static constexpr int kA = kTileM * KB;             // one query tile
  static constexpr int kB = kTileN * KB;             // one gallery tile
  static constexpr int kLab = kTileN * 4;
  static constexpr int kQueue = 8 * kTileM * 2;      // u16 [8][128] candidate queue per query tile
  static constexpr int offA = 0;
  static constexpr int offB = offA + QT * kA;
  static constexpr int offLab = offB + kTcStages * kB;
  static constexpr int offQueue = offLab + kTcStages * kLab;
  static constexpr int offHist = offQueue + QT * kQueue;
  static size_t total(int kbins) { return offHist + static_cast<size_t>(QT) * kbins * kTileM * 4; }
};

// One CTA = QT consecutive 128-query tiles x one gallery stripe.  Every gallery tile (128 rows) is multiplied
// against all QT query tiles (QT accumulators of 128 TMEM columns each, 512 columns in total for QT = 4); each
// query tile has its own epilogue warpgroup (4 warps), so 4 * QT warps hide each other's latencies while every
// query still sees its gallery rows strictly in row order (which the stable prefixes need).
template <int KB, int LAB, int QT, int kTcStages>
__global__ void __launch_bounds__(128 + 128 * QT, 1) hamming_select_tc_kernel(const TcDev t) {
  typedef TcSmem<KB, QT, kTcStages> S;
  const HistDev& a = t.h;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_a, bar_full[kTcStages], bar_empty[kTcStages], bar_tfull[QT], bar_tempty[QT];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int qgroup = blockIdx.x % t.nqgroups;
  const int stripe = blockIdx.x / t.nqgroups;
  const int qtile0 = qgroup * QT;
  int nvalid = t.nqtiles128 - qtile0;
  if (nvalid > QT) nvalid = QT;
  const long long row_begin = static_cast<long long>(stripe) * a.rows_per_stripe;
  long long row_end = row_begin + a.rows_per_stripe;
  if (row_end > a.ndb) row_end = a.ndb;
  const int ntiles = row_end > row_begin ? static_cast<int>((row_end - row_begin + kTileN - 1) / kTileN) : 0;

  if (tid == 0) {
    mbar_init(&bar_a, 1);
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1 + 4 * nvalid);   // MMA commit + every epilogue warp (class ids live in the stage)
    }
    for (int i = 0; i < QT; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0 && ntiles > 0) {
      mbar_arrive_expect_tx(&bar_a, nvalid * S::kA);
      bulk_g2s(smem + S::offA, t.q_i8 + static_cast<size_t>(qtile0) * kTileM * KB, nvalid * S::kA, &bar_a);
      for (int k = 0; k < ntiles; ++k) {
        const int s = k % kTcStages;
        mbar_wait_backoff(&bar_empty[s], static_cast<uint32_t>(((k / kTcStages) & 1) ^ 1));
        const long long r0 = row_begin + static_cast<long long>(k) * kTileN;
        long long rows = row_end - r0;
        if (rows > kTileN) rows = kTileN;
        const uint32_t rows8 = static_cast<uint32_t>((rows + 7) & ~7ll);   // pad rows exist (ch_padded_rows)
        const uint32_t bytes_b = rows8 * KB;
        const uint32_t bytes_l = LAB == CH_LAB_ID ? rows8 * 4u : 0u;
        mbar_arrive_expect_tx(&bar_full[s], bytes_b + bytes_l);
        bulk_g2s(smem + S::offB + s * S::kB, t.g_i8 + static_cast<size_t>(r0) * KB, bytes_b, &bar_full[s]);
        if (LAB == CH_LAB_ID) bulk_g2s(smem + S::offLab + s * S::kLab, a.g_lab + r0, bytes_l, &bar_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && ntiles > 0) {
      // s32 accumulate, s8 x s8, both K-major, N = kTileN, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kTileN >> 3) << 17) |
                             (static_cast<uint32_t>(kTileM >> 4) << 24);
      const uint32_t sbo = 8 * KB, lbo = 128;
      const uint32_t a_addr = smem_u32(smem + S::offA);
      mbar_wait_backoff(&bar_a, 0);
      for (int k = 0; k < ntiles; ++k) {
        const int s = k % kTcStages;
        mbar_wait_backoff(&bar_full[s], static_cast<uint32_t>((k / kTcStages) & 1));
        const uint32_t b_addr = smem_u32(smem + S::offB + s * S::kB);
        for (int i = 0; i < nvalid; ++i) {
          mbar_wait_backoff(&bar_tempty[i], static_cast<uint32_t>((k & 1) ^ 1));   // accumulator i drained
          tc_fence_after();
          const uint32_t d_addr = tmem_base + static_cast<uint32_t>(i) * kTileN;
#pragma unroll
          for (int kk = 0; kk < KB / 32; ++kk)
            umma_i8(d_addr, umma_desc(a_addr + i * S::kA + kk * 256, lbo, sbo), umma_desc(b_addr + kk * 256, lbo, sbo),
                    idesc, kk > 0 ? 1u : 0u);
          umma_commit(&bar_tfull[i]);   // accumulator i holds tile k
        }
        umma_commit(&bar_empty[s]);     // the stage's code bytes may be overwritten once these MMAs retire
      }
    }
  } else if (warp >= 4 && ((warp - 4) >> 2) < nvalid) {
    // ===================== epilogue: thread = TMEM lane = query =====================
    const int qt = (warp - 4) >> 2;                // query tile of this warpgroup
    const int e = (tid - 128) & 127;               // 0..127 within the query tile
    const int ewarp = warp & 3;                    // TMEM lanes 32 * ewarp ..
    unsigned char* hist_b = smem + S::offHist + static_cast<size_t>(qt) * t.kbins * kTileM * 4 + e * 4;
    unsigned short* queue = reinterpret_cast<unsigned short*>(smem + S::offQueue + qt * S::kQueue) + e;
    constexpr uint32_t T4 = kTileM * 4u;
    const long long q = static_cast<long long>(qtile0 + qt) * kTileM + e;
    const bool active = q < a.nq;
    uint32_t qid = CH_QUERY_NOLABEL;
    if (LAB == CH_LAB_ID && active) qid = a.q_lab[q];
    // key <= thresh  <=>  dot >= nbit - 2 * thresh
    const int tau = active ? a.nbit - 2 * static_cast<int>(a.thresh[q]) : 0x7fffffff;
    for (int b = 0; b < t.kbins; ++b) *reinterpret_cast<uint32_t*>(hist_b + b * T4) = 0u;

    const size_t sq = static_cast<size_t>(stripe) * a.nq_pad + q;
    uint32_t rptr = 0, rstart = 0, rend = 0;
    const int emit_mode = active ? a.emit_mode : CH_EMIT_NONE;
    if (emit_mode != CH_EMIT_NONE) {
      rstart = a.rec_off[sq];
      rend = rstart + a.rec_cap[sq];
      rptr = rstart;
    }
    bool overflow = false;
    uint32_t epoch = 0;
    auto flush = [&]() {
      for (int b = 0; b < t.kbins; ++b) {
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
      ++epoch;
    };
    // cold path: the queued candidates of one 8-column group, in column order
    auto drain = [&](uint32_t n, int c0, int rows, const uint32_t* labs, uint32_t shard_row0) {
      for (uint32_t i = 0; i < n; ++i) {
        const uint32_t ent = queue[i * kTileM];
        const uint32_t key = ent >> 5;
        const int col = c0 + static_cast<int>(ent & 31u);
        if (col >= rows) continue;
        bool rel = false;
        if (LAB == CH_LAB_ID) rel = labs[col] == qid;
        uint32_t* h = reinterpret_cast<uint32_t*>(hist_b + key * T4);
        const uint32_t old = *h;
        *h = old + (rel ? 0x10001u : 1u);
        if (emit_mode == CH_EMIT_CANDIDATES || (emit_mode == CH_EMIT_RELEVANT && rel)) {
          uint32_t base_all = 0, base_rel = 0;
          if (epoch != 0u) {
            const size_t o = (static_cast<size_t>(stripe) * a.nbins + key) * a.nq_pad + q;
            base_all = a.slab_all[o];
            if (LAB != CH_LAB_NONE) base_rel = a.slab_rel[o];
          }
          if (rptr < rend)
            a.recs[rptr] = make_uint4(key | (rel ? 0x80000000u : 0u), base_all + (old & 0xffffu),
                                      base_rel + (old >> 16), shard_row0 + col);
          else
            overflow = true;
          ++rptr;
        }
      }
    };

    for (int k = 0; k < ntiles; ++k) {
      const int s = k % kTcStages;
      long long rows_ll = row_end - (row_begin + static_cast<long long>(k) * kTileN);
      const int rows = rows_ll > kTileN ? kTileN : static_cast<int>(rows_ll);
      const uint32_t shard_row0 = static_cast<uint32_t>(a.row_base + row_begin) + static_cast<uint32_t>(k) * kTileN;
      const uint32_t* labs = reinterpret_cast<const uint32_t*>(smem + S::offLab + s * S::kLab);
      mbar_wait(&bar_full[s], static_cast<uint32_t>((k / kTcStages) & 1));   // class ids of the stage are visible
      mbar_wait(&bar_tfull[qt], static_cast<uint32_t>(k & 1));
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + static_cast<uint32_t>(qt) * kTileN;
      for (int c0 = 0; c0 < rows; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr0 + c0, r);
        // hot path: is any of the 32 dots >= tau?  (4 sub-maxima of 8, 3-input max tree; garbage columns past
        // `rows` in the stripe's last tile can only cause a spurious visit of the cold path)
        int mg[4];
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          const int o = g4 * 8;
          int x = max(max(static_cast<int>(r[o]), static_cast<int>(r[o + 1])), static_cast<int>(r[o + 2]));
          x = max(max(x, static_cast<int>(r[o + 3])), static_cast<int>(r[o + 4]));
          x = max(max(x, static_cast<int>(r[o + 5])), static_cast<int>(r[o + 6]));
          mg[g4] = max(x, static_cast<int>(r[o + 7]));
        }
        if (max(max(mg[0], mg[1]), max(mg[2], mg[3])) >= tau) {
          // cold path (a few lanes): per 8-column group queue the candidates, then handle them in a rolled loop
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            if (mg[g4] >= tau) {
              uint32_t n = 0;
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const int j = g4 * 8 + jj;
                const int dot = static_cast<int>(r[j]);
                if (dot >= tau) {
                  queue[n * kTileM] = static_cast<unsigned short>(((static_cast<uint32_t>(a.nbit - dot) >> 1) << 5) | j);
                  ++n;
                }
              }
              drain(n, c0, rows, labs, shard_row0);
            }
          }
        }
      }
      // this warp is done with the accumulator and with the stage's class ids
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bar_tempty[qt]);
        mbar_arrive(&bar_empty[s]);
      }
      if ((k + 1) % a.flush_tiles == 0 && k + 1 < ntiles) flush();
    }
    flush();
    if (a.emit_mode != CH_EMIT_NONE && active) {
      a.rec_cnt[sq] = (rptr < rend ? rptr : rend) - rstart;
      if (overflow) atomicOr(a.err_flag, 1u);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// packed sign bits -> +-1 int8 in the tiled core-matrix order; one thread per (row, 16-byte K chunk)
__global__ void expand_i8_tiled_kernel(const uint32_t* __restrict__ bits, long long rows, int words, int nbit, int kb,
                                       int8_t* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int chunks = kb / 16;
  if (i >= rows * chunks) return;
  // consecutive threads -> consecutive 16-byte slots of the OUTPUT (coalesced 128-bit stores)
  const long long group = i / (8ll * chunks);
  const int within = static_cast<int>(i - group * 8ll * chunks);
  const int chunk = within / 8, r8 = within % 8;
  const long long row = group * 8 + r8;
  const int k0 = chunk * 16;
  uint32_t w = 0;
  if (k0 < words * 32) w = bits[row * words + (k0 >> 5)] >> (k0 & 31);
  uint32_t o[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + v * 4 + j;
      const uint32_t byte = (k < nbit) ? (((w >> (v * 4 + j)) & 1u) ? 0x01u : 0xffu) : 0u;
      x |= byte << (8 * j);
    }
    o[v] = x;
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

typedef void (*tc_fn_t)(const TcDev);
template <int KB, int QT, int ST>
tc_fn_t pick_lab_tc(int lab, int kbins, size_t* smem) {
  *smem = TcSmem<KB, QT, ST>::total(kbins);
  return lab == CH_LAB_ID ? hamming_select_tc_kernel<KB, CH_LAB_ID, QT, ST>
                          : hamming_select_tc_kernel<KB, CH_LAB_NONE, QT, ST>;
}
template <int KB>
tc_fn_t pick_qt_tc(int qt, int st, int lab, int kbins, size_t* smem) {
  if (st == 4) {
    switch (qt) {
      case 4: return pick_lab_tc<KB, 4, 4>(lab, kbins, smem);
      case 3: return pick_lab_tc<KB, 3, 4>(lab, kbins, smem);
      case 2: return pick_lab_tc<KB, 2, 4>(lab, kbins, smem);
      default: return pick_lab_tc<KB, 1, 4>(lab, kbins, smem);
    }
  }
  switch (qt) {
    case 4: return pick_lab_tc<KB, 4, 3>(lab, kbins, smem);
    case 3: return pick_lab_tc<KB, 3, 3>(lab, kbins, smem);
    case 2: return pick_lab_tc<KB, 2, 3>(lab, kbins, smem);
    default: return pick_lab_tc<KB, 1, 3>(lab, kbins, smem);
  }
}
tc_fn_t pick_tc(int kb, int qt, int st, int lab, int kbins, size_t* smem) {
  switch (kb) {
    case 32: return pick_qt_tc<32>(qt, st, lab, kbins, smem);
    case 64: return pick_qt_tc<64>(qt, st, lab, kbins, smem);
    case 96: return pick_qt_tc<96>(qt, st, lab, kbins, smem);
    default: return pick_qt_tc<128>(qt, st, lab, kbins, smem);
  }
}

}  // namespace

extern "C" int ch_tc_code_bytes(int nbit) {
  if (nbit <= 0 || nbit > 128) return 0;
  return (nbit + 31) / 32 * 32;
}

extern "C" int ch_expand_i8(ch_ws* ws, const uint32_t* bits_dev, int64_t rows_pad, int nbit, int8_t* out_dev,
                            void* stream) {
  if (ws == nullptr || bits_dev == nullptr || out_dev == nullptr) CH_FAIL("null argument to ch_expand_i8");
  const int kb = ch_tc_code_bytes(nbit);
  if (kb == 0) CH_FAIL("nbit=%d unsupported by the tensor-core path (1..128)", nbit);
  if (rows_pad % 8) CH_FAIL("rows_pad must be a multiple of 8");
  ChDeviceGuard guard(ws->device);
  const long long n = rows_pad * (kb / 16);
  expand_i8_tiled_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bits_dev, rows_pad, ch_code_words(nbit), nbit, kb, out_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_hamming_select_tc(ch_ws* ws, const ch_hist_args* a, const int8_t* q_i8, const int8_t* g_i8,
                                    void* stream) {
  if (ws == nullptr || a == nullptr || q_i8 == nullptr || g_i8 == nullptr) CH_FAIL("null argument to ch_hamming_select_tc");
  if (a->thresh == nullptr) CH_FAIL("the tensor-core kernel implements the select pass only (thresh required)");
  if (a->ternary) CH_FAIL("the tensor-core kernel needs +-1 codes (no zeros)");
  if (a->label_mode != CH_LAB_NONE && a->label_mode != CH_LAB_ID) CH_FAIL("label mode %d unsupported", a->label_mode);
  const int kb = ch_tc_code_bytes(a->nbit);
  if (kb == 0) CH_FAIL("nbit=%d unsupported by the tensor-core path (1..128)", a->nbit);
  if (a->nq <= 0 || a->ndb < 0 || a->slab_all == nullptr) CH_FAIL("bad arguments");
  if (a->label_mode == CH_LAB_ID && (a->q_lab == nullptr || a->g_lab == nullptr || a->slab_rel == nullptr))
    CH_FAIL("labels / relevant slab missing");
  if (a->emit_mode != CH_EMIT_NONE &&
      (a->rec_off == nullptr || a->rec_cap == nullptr || a->rec_cnt == nullptr || a->recs == nullptr ||
       a->err_flag == nullptr))
    CH_FAIL("record emission needs rec_off / rec_cap / rec_cnt / recs / err_flag");
  if (a->nq_pad % kTileM) CH_FAIL("nq_pad must be a multiple of %d", kTileM);
  if (a->nstripes <= 0 || a->rows_per_stripe <= 0 || a->rows_per_stripe % 256 ||
      static_cast<long long>(a->nstripes) * a->rows_per_stripe < a->ndb)
    CH_FAIL("bad stripe geometry");
  if ((reinterpret_cast<uintptr_t>(q_i8) | reinterpret_cast<uintptr_t>(g_i8) | reinterpret_cast<uintptr_t>(a->g_lab)) & 15)
    CH_FAIL("operands must be 16-byte aligned");
  ChDeviceGuard guard(ws->device);
  TcDev d;
  HistDev& h = d.h;
  h.q_bits = nullptr; h.q_nz = nullptr; h.g_bits = nullptr; h.g_nz = nullptr;
  h.q_lab = a->q_lab; h.g_lab = a->g_lab; h.slab_all = a->slab_all; h.slab_rel = a->slab_rel;
  h.thresh = a->thresh; h.rec_off = a->rec_off; h.rec_cap = a->rec_cap; h.rec_cnt = a->rec_cnt;
  h.recs = static_cast<uint4*>(a->recs); h.err_flag = a->err_flag;
  h.nq = a->nq; h.nq_pad = a->nq_pad; h.ndb = a->ndb; h.row_base = a->row_base;
  h.nbit = a->nbit; h.nbins = a->nbit + 1; h.lw = 0; h.emit_mode = a->emit_mode;
  // keys above max(thresh) never occur: with key_limit the slabs have key_limit rows per stripe
  if (a->key_limit > 0 && a->key_limit < h.nbins) h.nbins = a->key_limit;
  h.nqtiles = static_cast<int>(a->nq_pad / kTileM); h.rows_per_stripe = a->rows_per_stripe; h.tile_rows = kTileN;
  h.flush_tiles = 65535 / kTileN;
  d.q_i8 = q_i8; d.g_i8 = g_i8;
  // only keys <= max threshold are reachable: the caller may pass that bound (key_limit = max thresh + 1)
  d.kbins = h.nbins;
  d.nqtiles128 = static_cast<int>(a->nq_pad / kTileM);
  // as many query tiles per CTA as shared memory allows (more epilogue warps = better latency hiding);
  // never more than the problem has; 4 pipeline stages if they fit, else 3
  const size_t smem_max = static_cast<size_t>(ws->max_smem_optin) - 1024;
  tc_fn_t fn = nullptr;
  size_t smem = 0;
  int qt = d.nqtiles128 < 4 ? d.nqtiles128 : 4;
  int st = 4;
  for (;; --qt) {
    st = 4;
    fn = pick_tc(kb, qt, st, a->label_mode, d.kbins, &smem);
    if (smem <= smem_max) break;
    st = 3;
    fn = pick_tc(kb, qt, st, a->label_mode, d.kbins, &smem);
    if (smem <= smem_max || qt == 1) break;
  }
  if (smem > smem_max) CH_FAIL("tensor-core kernel needs %zu bytes of shared memory (key_limit=%d)", smem, d.kbins);
  d.nqgroups = (d.nqtiles128 + qt - 1) / qt;
  if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: each CTA owns all 512 TMEM columns
  CH_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const long long ctas = static_cast<long long>(d.nqgroups) * a->nstripes;
  if (ctas > 0x7fffffffll) CH_FAIL("grid too large");
  fn<<<static_cast<unsigned>(ctas), 128 + 128 * qt, smem, static_cast<cudaStream_t>(stream)>>>(d);
  CH_LAUNCH_CHECK(ws);
  return 0;
}
