// K2, tensor-core form of the SELECT pass: which (query, gallery row) pairs have key <= the query's threshold?
//
// For +-1 codes held as int8, <q, g> = nbit - 2 * hamming(q, g) exactly, so tcgen05.mma kind::i8 (UTCIMMA, int32
// accumulators in TMEM) produces the distances of a 128-query x 128-gallery-row tile at a cost of KB/32
// instructions.  The per-query threshold rides in the contraction as well: two extra K slots hold
// (a, b), a + b = -(nbit - 2 * thresh[q]), on the query side and (1, 1) on the gallery side, so the accumulator is
//
//     D[q][j] = <q, g_j> - (nbit - 2 * thresh[q])        and        key(q, j) <= thresh[q]  <=>  D[q][j] >= 0.
//
// What is left per pair is the SIGN BIT of one TMEM word.  The epilogue reads the accumulator with
// tcgen05.ld ... .pack::16b (LDTM.PACK16BIT: the low 16 bits of two adjacent columns per register; |D| < 2^15), so a
// whole 128-column accumulator fits in 64 registers and is handed back to the tensor core immediately.  Per 32
// columns, 8 PRMT in sign-replicate mode turn the 32 sign bits into 32 sign BYTES and 8 LOP3 fold them into one
// 32-bit mask -- in the "sparse" variant only after an 8-instruction AND tree said that the chunk holds a
// candidate at all.  The gallery plane stores every 32-row block in the row order that makes bit 31 - t of that
// mask the sign of row t (kRowOfColumn), so a leading-zero count walks the candidates in row order.  The shard-local row index of every candidate is appended to the
// (stripe, query) slice of the candidate list, in ascending row order (thread = TMEM lane = query; tiles, chunks
// and bits are visited in row order).  Nothing else happens here: keys, label matches, stable ranks and AP are
// the business of cand.cu, which only ever sees the candidates (~0.01-0.3 % of the pairs).
//
// Warp roles (640 threads, one CTA per SM, persistent over the tiles of one (4 query tiles, stripe)):
//   warps 0-3   MMA issuers, one thread per query tile: KB/32 UTCIMMA (M = N = 128) per gallery tile into that
//               query tile's TMEM accumulator (4 x 128 columns = all 512 columns)
//   warp 0      its thread is also the producer: 1-D bulk async copies (UBLKCP) of gallery tiles into a 4-stage ring
//   warp 2      also the TMEM allocator
//   warps 4-19  epilogue: one warpgroup per query tile; the accumulator is read with two packed tcgen05.ld.x32 and
//               released to the MMA issuer BEFORE its sign bits are examined, so the refill overlaps the examination
// Operands live in shared memory in the canonical NO-SWIZZLE K-major core-matrix layout (8 rows x 16 bytes =
// 128 contiguous bytes; next 16-byte K chunk at +128 B (LBO); next 8-row group at +8*KB (SBO)).  The int8 planes
// are stored in HBM already in that order (expand_i8_tiled_kernel), so a tile is one contiguous bulk copy.
#include "common.cuh"

namespace {

constexpr int kTileN = 128;   // gallery rows per MMA tile (TMEM columns per accumulator, rows per shared-memory stage)
constexpr int kTileM = 128;   // queries per query tile
constexpr int kQT = 4;        // query tiles per CTA (4 accumulators x 128 columns = all 512 TMEM columns)
// gallery-tile ring: as deep as the 227 KB of shared memory allow beside the four resident query tiles
// (KB <= 192: 4 stages; KB = 224: 3; KB = 256 / 288, i.e. 225 ... 256-bit codes: 2)
__host__ __device__ constexpr int stages_for(int kb) { return kb <= 192 ? 4 : (kb <= 224 ? 3 : 2); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version 1 (Blackwell); SWIZZLE_NONE, base offset 0
  return d;
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

#define CH_R32(r)                                                                                                  \
  r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15], r[16],     \
      r[17], r[18], r[19], r[20], r[21], r[22], r[23], r[24], r[25], r[26], r[27], r[28], r[29], r[30], r[31]
#define CH_OUT32(r)                                                                                                \
  "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define CH_INOUT32(r)                                                                                              \
  "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),      \
      "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),       \
      "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),      \
      "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])

// 64 columns -> 32 registers: r[j] = lo16(column 2j) | lo16(column 2j + 1) << 16.
// issue only: the registers are NOT valid until tmem_wait(r) has been executed
__device__ __forceinline__ void tmem_ld64p_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : CH_OUT32(r)
      : "r"(taddr)
      : "memory");
}
// the wait names the registers as in/out operands: no use of r[] can be scheduled above it
__device__ __forceinline__ void tmem_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" : CH_INOUT32(r)::"memory");
}

struct SelDev {
  const int8_t* q_i8;          // tiled (>= nq_pad rows, KB); threshold slots filled by expand_i8_tiled_kernel
  const int8_t* g_i8;          // tiled (rows_pad, KB)
  const uint32_t* cand_off;    // (nstripes, nq_pad)
  const uint32_t* cand_cap;    // (nstripes, nq_pad)
  uint32_t* cand_cnt;          // (nstripes, nq_pad)
  uint32_t* cand_rows;         // shard-local row index of every candidate
  uint32_t* err_flag;
  uint32_t* bad;               // (nq_pad) or NULL: set to 1 for a query whose slice overflowed
  const uint32_t* thresh;      // per-query threshold keys when the comparison happens in the epilogue, else NULL
  int nbit, ternary;
  long long nq, nq_pad, ndb, row_base;
  int rows_per_stripe;
  long long q_stripe_bytes;    // distance of the per-stripe query planes (0: one plane)
  int nqtiles128;              // 128-query tiles in total
  int nqgroups;                // CTAs along the query axis (each owns kQT consecutive query tiles)
};

template <int KB>
struct SelSmem {
  static constexpr int kA = kTileM * KB;   // one query tile
  static constexpr int kB = kTileN * KB;   // one gallery tile
  static constexpr int offA = 0;
  static constexpr int offB = offA + kQT * kA;
  static constexpr int total = offB + stages_for(KB) * kB;
};

// Column c of a 32-column block holds gallery row kRowOfColumn(c) of the 32-row block: with the PRMT / LOP3
// folding below, bit 31 - t of the mask then is the sign of row t (leading-zero count = first candidate row).
__host__ __device__ constexpr int kRowOfColumn(int c) { return 31 - (8 * (c & 3) + (c >> 2)); }

// 16 packed registers = 32 columns -> mask, bit 31 - t = 1 iff the accumulator of row t is negative (no candidate).
// PRMT selector nibbles 9, b, d, f: bytes 1, 3 (sign bytes of the halves of x) and 5, 7 (of y), sign-replicated.
__device__ __forceinline__ uint32_t sign_mask32p(const uint32_t* r) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t p;
    asm("prmt.b32 %0, %1, %2, 0xfdb9;" : "=r"(p) : "r"(r[2 * k]), "r"(r[2 * k + 1]));
    m |= p & (0x01010101u << k);
  }
  return m;
}
// true iff some accumulator of the 32 columns is >= 0: 3-input AND tree over the packed registers
__device__ __forceinline__ bool any_candidate32p(const uint32_t* r) {
  const uint32_t a0 = r[0] & r[1] & r[2], a1 = r[3] & r[4] & r[5], a2 = r[6] & r[7] & r[8];
  const uint32_t a3 = r[9] & r[10] & r[11], a4 = r[12] & r[13] & r[14];
  const uint32_t x = (a0 & a1 & a2) & (a3 & a4 & r[15]);
  return (x & 0x80008000u) != 0x80008000u;
}

// ---- the comparison in the epilogue (sparse passes: the threshold block of the contraction is dropped) ----------
// The accumulator then is the plain inner product <q, g>, and key <= thresh <=> <q, g> >= tau, tau = nbit - 2 thresh
// (binary) or nbit - thresh (ternary, doubled key scale).  t2 = tau in both 16-bit halves.  Whether a 32-column chunk
// holds a candidate is a packed 16-bit MAX tree (VIMNMX3.S16x2: 8 instructions for 16 registers, as many as the AND
// tree of sign bits above) and one packed subtraction; only chunks that do hold one pay for the 16 packed
// subtractions that turn their registers into the D >= 0 form sign_mask32p expects.
__device__ __forceinline__ bool any_candidate32p_tau(const uint32_t* r, uint32_t t2) {
  uint32_t m = __vmaxs2(__vmaxs2(r[0], r[1]), r[2]);
#pragma unroll
  for (int j = 3; j + 1 < 16; j += 2) m = __vmaxs2(__vmaxs2(m, r[j]), r[j + 1]);
  m = __vmaxs2(m, r[15]);
  return (__vsub2(m, t2) & 0x80008000u) != 0x80008000u;
}
__device__ __forceinline__ uint32_t sign_mask32p_tau(const uint32_t* r, uint32_t t2) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t p;
    asm("prmt.b32 %0, %1, %2, 0xfdb9;" : "=r"(p) : "r"(__vsub2(r[2 * k], t2)), "r"(__vsub2(r[2 * k + 1], t2)));
    m |= p & (0x01010101u << k);
  }
  return m;
}

// ---- two gallery rows per accumulator cell (PAIR; binary codes of <= 128 bits, ternary codes of <= 64) ----------
// A 16-bit accumulator value has room for TWO comparisons.  With f = thresh - key in [-128, 127] (key <= 128), the
// paired planes (expand_i8_pair_kernel) make the contraction produce
//     E = (f_a + 128) + 256 * (-f_b - 1)            in [-32768, 32767]
// for the two gallery rows a, b that share a plane row: the query side holds [q, 64 q, slots], the gallery side
// [a, -4 b, slot constants]; for binary codes the gallery side is the {0, 1} form of the code, so that
// <q, g01> = (#ones of q) - hamming(q, g) -- a difference of DISTANCES, not of inner products, which is what makes
// 129 values fit a byte.  Bit 7 of E is "row a is a candidate", bit 15 "row b is a candidate": every byte of a packed
// register pair carries one sign bit, twice as many as in the one-row form, at the same PRMT + LOP3 per register --
// and a 128-column accumulator covers 256 gallery rows, so there are half as many accumulator hand-overs, TMEM
// loads and barrier round trips per pair.  (What bounds the one-row kernel is that chain, not the MMA: DESIGN.md 5.)
// Column c of a 32-column block holds rows kPairRow(c, 0) and kPairRow(c, 1) of the 64-row block: with the folding
// below, bit 31 - t of mask h (registers 8 h .. 8 h + 7) is row 32 h + t.
__host__ __device__ constexpr int kPairRow(int c, int f) {
  return 32 * (c >> 4) + 31 - 8 * (2 * (c & 1) + f) - ((c >> 1) & 7);
}
constexpr int kPairSlots = 5;   // query side (x0, x1, y1, y2, y3) . gallery side (2, 1, -128, -128, -128)
// 8 packed registers = 16 columns x 2 rows -> mask, bit 31 - t = 1 iff row t of the 32 is a candidate
__device__ __forceinline__ uint32_t pair_mask32(const uint32_t* r) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t p;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(p) : "r"(r[k]));
    m |= p & (0x01010101u << k);
  }
  return m;
}
__device__ __forceinline__ bool pair_any32(const uint32_t* r) {
  const uint32_t x = (r[0] | r[1] | r[2]) | (r[3] | r[4] | r[5]) | (r[6] | r[7]);
  return (x & 0x80808080u) != 0u;
}

// the MMA with a compile-time accumulate flag (no predicate set-up on the single issuing thread)
template <bool ACC>
__device__ __forceinline__ void umma_i8_imm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
  if (ACC)
    asm volatile(
        "{\n.reg .pred p;\nsetp.eq.u32 p, 1, 1;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\nsetp.eq.u32 p, 1, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
        : "memory");
}
// one lane of a converged warp (the warp stays converged: values computed around it can live in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0u;
}
// one non-blocking probe
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0u;
}
// tight poll for the threads whose latency is on the critical path of every tile (the MMA issuers)
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// One CTA = kQT consecutive 128-query tiles x one gallery stripe.
// TAU: the threshold comparison happens in the epilogue (planes without threshold slots, KB = the code bytes only)
// PAIR: two gallery rows per accumulator cell (planes of expand_i8_pair_kernel): a tile is 256 rows
template <int KB, bool DENSE, bool TAU, bool PAIR>
__global__ void __launch_bounds__(128 + 128 * kQT, 1) hamming_select_tc_kernel(const SelDev a) {
  typedef SelSmem<KB> S;
  constexpr int kStages = stages_for(KB);
  constexpr int kTileRows = PAIR ? 2 * kTileN : kTileN;   // gallery rows per tile
  static_assert(!(PAIR && TAU), "the paired form carries its thresholds in the contraction");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_a, bar_full[kStages], bar_empty[kStages], bar_tfull[kQT], bar_tempty[kQT];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int qgroup = blockIdx.x % a.nqgroups;
  const int stripe = blockIdx.x / a.nqgroups;
  const int qtile0 = qgroup * kQT;
  int nvalid = a.nqtiles128 - qtile0;
  if (nvalid > kQT) nvalid = kQT;
  const long long row_begin = static_cast<long long>(stripe) * a.rows_per_stripe;
  long long row_end = row_begin + a.rows_per_stripe;
  if (row_end > a.ndb) row_end = a.ndb;
  const int ntiles = row_end > row_begin ? static_cast<int>((row_end - row_begin + kTileRows - 1) / kTileRows) : 0;

  if (tid == 0) {
    mbar_init(&bar_a, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 4 * nvalid);   // every epilogue warp, once it has seen its accumulator complete
    }
    for (int i = 0; i < kQT; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], 4);    // the four epilogue warps of the query tile
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

  if (warp <= 3) {
    // ===================== MMA issuers (one thread per accumulator) + producer (the thread of warp 0) ==========
    // A tcgen05.mma is not queued deeply: the issuing thread stalls ~85 clk per instruction, and every commit /
    // barrier poll of that thread would leave the tensor core idle (measured with one issuer: ~240 clk per
    // accumulator and tile).  With one issuing thread per accumulator those gaps are filled by the others' MMAs.
    // The thread of warp 0 also feeds the gallery-tile ring; it never blocks on a stage unless the very next
    // tile is missing.
    const int i = warp;                                       // accumulator / query tile of this warp
    // The whole warp runs the loop (converged, warp-uniform values -> descriptors stay in uniform registers);
    // only the tcgen05 / bulk-copy instructions themselves are issued by one elected lane.
    if (ntiles > 0 && (i < nvalid || warp == 0)) {
      const bool issues = i < nvalid;
      // s32 accumulate, s8 x s8, both K-major, N = kTileN, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kTileN >> 3) << 17) |
                             (static_cast<uint32_t>(kTileM >> 4) << 24);
      // descriptors differ only in their 14-bit start-address field (units of 16 B): build one per operand once,
      // then add constants
      const uint64_t da = umma_desc(smem_u32(smem + S::offA + i * S::kA), 128, 8 * KB);
      const uint64_t db0 = umma_desc(smem_u32(smem + S::offB), 128, 8 * KB);
      const uint32_t d_addr = tmem_base + static_cast<uint32_t>(i * kTileN);
      int next_load = 0;                                      // producer state (warp 0 only)
      auto load_tile = [&](int t) {
        const int st = t % kStages;
        const long long r0 = row_begin + static_cast<long long>(t) * kTileRows;
        long long rows = row_end - r0;
        if (rows > kTileRows) rows = kTileRows;
        // whole (permuted) blocks of 32 plane rows (= 32 gallery rows, or 64 in the paired form); pad rows exist
        // (ch_padded_rows)
        const uint32_t prows = PAIR ? static_cast<uint32_t>((rows + 63) >> 6) * 32u
                                    : static_cast<uint32_t>((rows + 31) & ~31ll);
        const uint32_t bytes_b = prows * KB;
        const size_t prow0 = static_cast<size_t>(PAIR ? (r0 >> 1) : r0);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar_full[st], bytes_b);
          bulk_g2s(smem + S::offB + st * S::kB, a.g_i8 + prow0 * KB, bytes_b, &bar_full[st]);
        }
      };
      if (warp == 0) {
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar_a, nvalid * S::kA);
          bulk_g2s(smem + S::offA,
                   a.q_i8 + static_cast<size_t>(stripe) * a.q_stripe_bytes + static_cast<size_t>(qtile0) * kTileM * KB,
                   nvalid * S::kA, &bar_a);
        }
        for (; next_load < kStages && next_load < ntiles; ++next_load) load_tile(next_load);
      }
      if (issues) mbar_wait_spin(&bar_a, 0);
      int s = 0;
      uint32_t ph_full = 0;
      for (int k = 0; k < ntiles; ++k) {
        if (issues) {
          mbar_wait_spin(&bar_full[s], ph_full);
          mbar_wait_spin(&bar_tempty[i], static_cast<uint32_t>((k & 1) ^ 1));   // read out by its epilogue warps
          tc_fence_after();
          const uint64_t db = db0 + static_cast<uint64_t>((s * S::kB) >> 4);
          // Long batches (>= 4 MMAs) are issued one at a time, not interleaved: interleaved, all four accumulators
          // complete together and the tensor core idles through all four hand-overs at once.  (Measured: -7 % at
          // KB = 160, +12 % at KB = 96 -- short batches are issued by whoever is ready.)
          // The turn is a token that goes round the issuing warps through NAMED BARRIERS (ids 1 + accumulator): no
          // shared-memory traffic -- a shared-memory CAS takes 200-300 clk while the MMA's operand fetch owns the
          // port -- and it is handed on BEFORE the last MMA of the batch is issued, so the next warp's first MMA
          // is queued right behind this batch instead of after a round trip (2-3 % against the CAS lock of round 1).
          constexpr bool kOrdered = KB >= 128;
          constexpr int kMmas = KB / 32;
          constexpr int kHead = kMmas > 1 ? kMmas - 1 : 1;
          const bool tok = kOrdered && nvalid > 1;
          if (tok && !(k == 0 && i == 0))
            asm volatile("barrier.sync.aligned %0, 64;" ::"r"(1 + i) : "memory");
          if (elect_one()) {
            umma_i8_imm<false>(d_addr, da, db, idesc);
#pragma unroll
            for (int kk = 1; kk < kHead; ++kk)
              umma_i8_imm<true>(d_addr, da + static_cast<uint64_t>((kk * 256) >> 4),
                                db + static_cast<uint64_t>((kk * 256) >> 4), idesc);
          }
          __syncwarp();
          // (the last arrival of the kernel would find nobody waiting: it is skipped)
          if (tok && !(k == ntiles - 1 && i == nvalid - 1))
            asm volatile("barrier.arrive.aligned %0, 64;" ::"r"(1 + (i + 1 == nvalid ? 0 : i + 1)) : "memory");
          if (elect_one()) {    // (elect.sync picks the same lane for the same mask: the commit covers both blocks)
#pragma unroll
            for (int kk = kHead; kk < kMmas; ++kk)
              umma_i8_imm<true>(d_addr, da + static_cast<uint64_t>((kk * 256) >> 4),
                                db + static_cast<uint64_t>((kk * 256) >> 4), idesc);
            umma_commit(&bar_tfull[i]);           // accumulator i holds tile k (and is done reading stage s)
          }
          __syncwarp();
        }
        if (warp == 0) {
          // refill every stage whose tile all epilogue warps have seen complete; block only for tile k + 1
          while (next_load < ntiles && next_load <= k + kStages) {
            uint64_t* eb = &bar_empty[next_load % kStages];
            const uint32_t par = static_cast<uint32_t>(((next_load / kStages) & 1) ^ 1);
            if (next_load == k + 1) mbar_wait_spin(eb, par);
            else if (!mbar_test(eb, par)) break;
            load_tile(next_load++);
          }
        }
        if (++s == kStages) {
          s = 0;
          ph_full ^= 1u;
        }
      }
    }
  } else if (warp >= 4 && ((warp - 4) >> 2) < nvalid) {
    // ===================== epilogue: thread = TMEM lane = query =====================
    const int qt = (warp - 4) >> 2;                // query tile of this warpgroup
    const int e = (tid - 128) & 127;               // 0..127 within the query tile
    const int ewarp = warp & 3;                    // TMEM lanes 32 * ewarp ..
    const long long q = static_cast<long long>(qtile0 + qt) * kTileM + e;
    const bool active = q < a.nq;
    const size_t sq = static_cast<size_t>(stripe) * a.nq_pad + q;
    uint32_t n = 0, cap = 0;
    uint32_t* out = a.cand_rows;
    if (active) {
      out += a.cand_off[sq];
      cap = a.cand_cap[sq];
    }
    // rows >= row_lim (pad rows of the shard's last tile, stale columns of a short tile) are never candidates;
    // an inactive lane (query padding) accepts nothing
    const uint32_t row_lim = active ? static_cast<uint32_t>(a.row_base + row_end) : 0u;
    const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + static_cast<uint32_t>(qt) * kTileN;
    uint32_t t2 = 0x7fff7fffu;                     // (inactive lanes: nothing reaches the threshold)
    if (TAU && active) {
      const int tau = a.nbit - (a.ternary ? 1 : 2) * static_cast<int>(a.thresh[q]);
      t2 = (static_cast<uint32_t>(tau) & 0xffffu) * 0x10001u;
    }

    // candidates of one 32-row block (bit 31 - t = row row0 + t), appended in ascending row order
    auto emit = [&](uint32_t cand, uint32_t row0) {
      while (cand != 0u) {
        const uint32_t z = static_cast<uint32_t>(__clz(static_cast<int>(cand)));
        cand &= ~(0x80000000u >> z);
        const uint32_t row = row0 + z;
        if (row < row_lim) {
          if (n < cap) out[n] = row;
          ++n;
        }
      }
    };

    uint32_t ra[32], rb[32];
    for (int k = 0; k < ntiles; ++k) {
      const uint32_t row0 = static_cast<uint32_t>(a.row_base + row_begin) + static_cast<uint32_t>(k) * kTileRows;
      mbar_wait(&bar_tfull[qt], static_cast<uint32_t>(k & 1));
      tc_fence_after();
      if (lane == 0) mbar_arrive(&bar_empty[k % kStages]);   // this query tile's MMAs are done with the stage
      tmem_ld64p_issue(taddr0, ra);
      tmem_ld64p_issue(taddr0 + 64, rb);
      tmem_wait(ra);
      tmem_wait(rb);
      // the whole accumulator is in registers: the MMA refills it while its sign bits are examined
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[qt]);
      if (PAIR) {
        // eight masks of 32 rows each (bit set = candidate), then the candidate loops
        uint32_t m[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const uint32_t* r = (h < 4 ? ra : rb) + 8 * (h & 3);
          m[h] = (DENSE || pair_any32(r)) ? pair_mask32(r) : 0u;
        }
#pragma unroll
        for (int h = 0; h < 8; ++h) emit(m[h], row0 + 32u * h);
        continue;
      }
      // all four masks first (independent instruction chains), then the (rare, divergent) candidate loops
      uint32_t m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
      if (DENSE) {
        m0 = ~sign_mask32p(ra);
        m1 = ~sign_mask32p(ra + 16);
        m2 = ~sign_mask32p(rb);
        m3 = ~sign_mask32p(rb + 16);
      } else if (TAU) {
        const bool h0 = any_candidate32p_tau(ra, t2), h1 = any_candidate32p_tau(ra + 16, t2);
        const bool h2 = any_candidate32p_tau(rb, t2), h3 = any_candidate32p_tau(rb + 16, t2);
        if (h0) m0 = ~sign_mask32p_tau(ra, t2);
        if (h1) m1 = ~sign_mask32p_tau(ra + 16, t2);
        if (h2) m2 = ~sign_mask32p_tau(rb, t2);
        if (h3) m3 = ~sign_mask32p_tau(rb + 16, t2);
      } else {
        const bool h0 = any_candidate32p(ra), h1 = any_candidate32p(ra + 16);
        const bool h2 = any_candidate32p(rb), h3 = any_candidate32p(rb + 16);
        if (h0) m0 = ~sign_mask32p(ra);            // no candidate among the 32 rows is the usual case
        if (h1) m1 = ~sign_mask32p(ra + 16);
        if (h2) m2 = ~sign_mask32p(rb);
        if (h3) m3 = ~sign_mask32p(rb + 16);
      }
      emit(m0, row0);
      emit(m1, row0 + 32);
      emit(m2, row0 + 64);
      emit(m3, row0 + 96);
    }
    if (active) {
      a.cand_cnt[sq] = n < cap ? n : cap;
      if (n > cap) {
        atomicOr(a.err_flag, 1u);
        if (a.bad != nullptr) a.bad[q] = 1u;     // (the caller re-ranks just these queries by the exact path)
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// packed sign bits -> {-1, 0, +1} int8 in the tiled core-matrix order; one thread per (row, 16-byte K chunk).
// A clear bit of the non-zero plane `nz` (ternary codes: sign(0) = 0) gives 0 -- the contraction then is the
// ternary inner product and key = 2 x distance = nbit - <q, g> (oracle/map_oracle.py hamming_distance_matrix).
// Bytes nbit .. nbit + slots - 1 of a row are the threshold slots: all 1 on the gallery side; on the query side
// (thresh != NULL) they sum to -tau, tau = nbit - 2 thresh[row] (binary keys) or nbit - thresh[row] (ternary, keys on
// the doubled scale), so that D = <q, g> - tau >= 0 <=> key <= thresh; padding queries get the most negative sum.
__global__ void expand_i8_tiled_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ nz,
                                       long long rows_bits, long long rows_out, int words, int nbit, int kb, int slots,
                                       int ternary, int query, const uint32_t* __restrict__ thresh, long long nq,
                                       int8_t* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int chunks = kb / 16;
  if (i >= rows_out * chunks) return;
  // consecutive threads -> consecutive 16-byte slots of the OUTPUT (coalesced 128-bit stores)
  const long long group = i / (8ll * chunks);
  const int within = static_cast<int>(i - group * 8ll * chunks);
  const int chunk = within / 8, r8 = within % 8;
  const long long row = group * 8 + r8;        // row of the PLANE = TMEM lane (queries) / TMEM column (gallery)
  // gallery planes: column c of a 32-row block holds row kRowOfColumn(c) of that block (see sign_mask32p)
  const long long src = query ? row : (row & ~31ll) + kRowOfColumn(static_cast<int>(row & 31));
  const int k0 = chunk * 16;
  uint32_t w = 0, z = 0xffffffffu;
  const bool has_bits = src < rows_bits;
  if (has_bits && k0 < words * 32) {
    w = bits[src * words + (k0 >> 5)] >> (k0 & 31);
    if (nz != nullptr) z = nz[src * words + (k0 >> 5)] >> (k0 & 31);
  }
  int rest = 0;                                 // what the remaining threshold slots still have to sum to
  bool query_pad = false;
  if (thresh != nullptr) {
    if (row < nq) rest = (ternary ? 1 : 2) * static_cast<int>(thresh[row]) - nbit;     // -tau, in [-nbit, nbit]
    else query_pad = true;
  }
  uint32_t o[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + v * 4 + j;
      uint32_t byte = 0u;
      if (k < nbit) {
        if (has_bits && ((z >> (v * 4 + j)) & 1u) != 0u) byte = (((w >> (v * 4 + j)) & 1u) != 0u) ? 0x01u : 0xffu;
      } else if (k < nbit + slots) {
        int sv = 1;
        if (thresh != nullptr) {
          // slot k - nbit takes as much of -tau as an int8 holds; the slots before it took theirs
          int r = rest;
          for (int t = 0; t < k - nbit; ++t) r -= (r < -128 ? -128 : (r > 127 ? 127 : r));
          sv = query_pad ? -128 : (r < -128 ? -128 : (r > 127 ? 127 : r));
        }
        byte = static_cast<uint32_t>(sv) & 0xffu;
      }
      x |= byte << (8 * j);
    }
    o[v] = x;
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// The paired planes (see kPairRow): one thread per (plane row, 16-byte K chunk); bytes 0 .. nbit - 1 = field a,
// nbit .. 2 nbit - 1 = field b, then kPairSlots threshold slots.
//   gallery (query == 0): plane row p of the 32-row block p / 32 holds gallery rows 64 (p / 32) + kPairRow(p % 32, f);
//     field a = the code as {0, 1} (binary) or {-1, 0, +1} (ternary), field b = -4 x that; slots (2, 1, -128, -128, -128)
//   query: plane row = query; field a = the code as {-1, 0, +1}, field b = 64 x that; slots (x0, x1, y1, y2, y3) with
//     2 x0 + x1 = c0 = thresh - P + 128 and -128 (y1 + y2 + y3) = 256 c1, c1 = P - thresh - 1, where P = the number of
//     one bits of the query (binary: <q, g01> = P - hamming) or nbit (ternary: <q, g> = nbit - key).  Then
//     E = <row_q, row_g> = (thresh - key_a + 128) + 256 (key_b - thresh - 1).  thresh >= 128 (only possible when it is
//     the largest key there is: every row is a candidate) is encoded as a zero code with c0 = 128, c1 = -1.
__global__ void expand_i8_pair_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ nz,
                                      long long rows_bits, long long rows_out, int words, int nbit, int kb,
                                      int ternary, int query, const uint32_t* __restrict__ thresh, long long nq,
                                      int8_t* __restrict__ out, const uint32_t* __restrict__ scut, int stripe0) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int chunks = kb / 16;
  if (i >= rows_out * chunks) return;
  out += static_cast<size_t>(blockIdx.y) * rows_out * kb;      // one query plane per stripe (blockIdx.y)
  const long long group = i / (8ll * chunks);
  const int within = static_cast<int>(i - group * 8ll * chunks);
  const int chunk = within / 8, r8 = within % 8;
  const long long prow = group * 8 + r8;       // row of the PLANE = TMEM lane (queries) / TMEM column (gallery)
  long long src0 = prow, src1 = prow;
  if (!query) {
    const long long base = (prow >> 5) * 64;
    const int c = static_cast<int>(prow & 31);
    src0 = base + kPairRow(c, 0);
    src1 = base + kPairRow(c, 1);
  }
  const int k0 = chunk * 16;
  // a query row whose code is written as zeros: padding, or thresh >= 128 (every row is a candidate)
  uint32_t t = 0u;
  bool zero_code = false, none = false;
  if (query) {
    zero_code = !(prow < nq && prow < rows_bits);
    if (!zero_code) {
      t = thresh[prow];
      if (scut != nullptr && static_cast<uint32_t>(stripe0) + blockIdx.y >= scut[prow]) {
        // this stripe lies beyond the query's cut: threshold t - 1 (below 0: no candidates, like a padding query)
        if (t == 0u) none = true;
        else if (t <= 128u) --t;           // (t > 128: every row qualifies with t - 1 as well)
      }
      zero_code = none || t > 127u;
    }
  }
  uint32_t o[4];
  if (nz == nullptr && (nbit & 15) == 0 && k0 + 16 <= 2 * nbit) {
    // binary codes, the usual widths: the 16 bytes are 16 consecutive bits of ONE field of one source row
    const int f = k0 >= nbit ? 1 : 0;
    const int bit0 = k0 - f * nbit;
    const long long r = f ? src1 : src0;
    const bool live = r < rows_bits && !zero_code;
    uint32_t b16 = 0u;
    if (live) b16 = (bits[r * words + (bit0 >> 5)] >> (bit0 & 31)) & 0xffffu;
    // byte = base ^ (bit * mul): query +-1 / +-64, gallery {0, 1} / {0, -4}
    const uint32_t mul = query ? (f ? 0x80u : 0xfeu) : (f ? 0xfcu : 0x01u);
    const uint32_t base = (query && live) ? (f ? 0xc0c0c0c0u : 0xffffffffu) : 0u;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint32_t x = (((b16 >> (4 * v)) & 0xfu) * 0x00204081u) & 0x01010101u;   // bit j -> byte j
      o[v] = live ? (base ^ (x * mul)) : 0u;
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    return;
  }
  // slot values of this row (only the chunks that reach the slots need them)
  int slot[kPairSlots] = {2, 1, -128, -128, -128};
  if (query) {
    int c0 = 0, c1 = 0;                          // padding queries: E = 0, nothing is a candidate
    if (prow < nq && prow < rows_bits && !none && k0 + 16 > 2 * nbit) {
      if (t > 127u) {                            // every row is a candidate: E = 128 - 256 whatever the row
        c0 = 128;
        c1 = -1;
      } else {
        int P = nbit;
        if (!ternary) {
          P = 0;
          for (int w = 0; w < words; ++w) P += __popc(bits[prow * words + w]);
        }
        c0 = static_cast<int>(t) - P + 128;
        c1 = P - static_cast<int>(t) - 1;
      }
    }
    slot[0] = c0 >> 1;
    slot[1] = c0 & 1;
    int v = -2 * c1;
    for (int s = 2; s < kPairSlots; ++s) {
      const int y = v < -128 ? -128 : (v > 127 ? 127 : v);
      slot[s] = y;
      v -= y;
    }
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + v * 4 + j;
      int val = 0;
      if (k < 2 * nbit) {
        const int f = k < nbit ? 0 : 1;
        const int bit = k - f * nbit;
        const long long r = f ? src1 : src0;
        if (r < rows_bits && !zero_code) {
          const uint32_t w = bits[r * words + (bit >> 5)] >> (bit & 31);
          const bool nonzero = nz == nullptr || ((nz[r * words + (bit >> 5)] >> (bit & 31)) & 1u) != 0u;
          int code;                              // the code value of this operand form
          if (query || ternary) code = nonzero ? ((w & 1u) ? 1 : -1) : 0;
          else code = static_cast<int>(w & 1u);  // binary gallery: {0, 1}
          val = code * (query ? (f ? 64 : 1) : (f ? -4 : 1));
        }
      } else if (k < 2 * nbit + kPairSlots) {
        const int sidx = k - 2 * nbit;           // (selects, not an indexed load: slot[] stays in registers)
        val = sidx == 0 ? slot[0] : sidx == 1 ? slot[1] : sidx == 2 ? slot[2] : sidx == 3 ? slot[3] : slot[4];
      }
      x |= (static_cast<uint32_t>(val) & 0xffu) << (8 * j);
    }
    o[v] = x;
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

typedef void (*sel_fn_t)(const SelDev);
template <int KB>
sel_fn_t pick_dense(int dense, int tau, int pair, size_t* smem) {
  *smem = SelSmem<KB>::total;
  if (pair)
    return dense ? hamming_select_tc_kernel<KB, true, false, true> : hamming_select_tc_kernel<KB, false, false, true>;
  if (tau) return hamming_select_tc_kernel<KB, false, true, false>;
  return dense ? hamming_select_tc_kernel<KB, true, false, false> : hamming_select_tc_kernel<KB, false, false, false>;
}
sel_fn_t pick_sel(int kb, int dense, int tau, int pair, size_t* smem) {
  switch (kb) {
    case 32: return pick_dense<32>(dense, tau, pair, smem);
    case 64: return pick_dense<64>(dense, tau, pair, smem);
    case 96: return pick_dense<96>(dense, tau, pair, smem);
    case 128: return pick_dense<128>(dense, tau, pair, smem);
    case 160: return pick_dense<160>(dense, tau, pair, smem);
    case 192: return pick_dense<192>(dense, tau, pair, smem);
    case 224: return pick_dense<224>(dense, tau, pair, smem);
    case 256: return pick_dense<256>(dense, tau, pair, smem);
    default: return pick_dense<288>(dense, tau, pair, smem);
  }
}

// threshold slots: |tau| <= nbit must be a sum of int8 values
int thresh_slots(int nbit) { return nbit <= 254 ? 2 : 4; }

}  // namespace

extern "C" int ch_tc_queries_per_cta(void) { return kQT * kTileM; }

extern "C" int ch_tc_code_bytes(int nbit) {
  if (nbit <= 0 || nbit > CH_MAX_NBIT) return 0;
  return (nbit + thresh_slots(nbit) + 31) / 32 * 32;   // the codes + the threshold slots, in whole 32-byte K blocks
}

extern "C" int ch_tc_code_bytes_bare(int nbit) {
  if (nbit <= 0 || nbit > CH_MAX_NBIT) return 0;
  return (nbit + 31) / 32 * 32;                        // no threshold slots: the comparison happens in the epilogue
}

extern "C" int ch_tc_code_bytes_pair(int nbit, int ternary) {
  // thresh - key must fit a signed byte: keys 0 .. 128
  if (nbit <= 0 || nbit > (ternary ? 64 : 128)) return 0;
  return (2 * nbit + kPairSlots + 31) / 32 * 32;       // both fields + the threshold slots, in whole 32-byte K blocks
}

extern "C" int ch_tc_tile_rows(int pair) { return pair ? 2 * kTileN : kTileN; }

extern "C" int ch_expand_i8(ch_ws* ws, const uint32_t* bits_dev, const uint32_t* nz_dev, int64_t rows_bits, int nbit,
                            int ternary, int bare, int8_t* out_dev, int64_t rows_out, const uint32_t* thresh_dev,
                            int64_t nq, void* stream) {
  if (ws == nullptr || bits_dev == nullptr || out_dev == nullptr) CH_FAIL("null argument to ch_expand_i8");
  if (ternary && nz_dev == nullptr) CH_FAIL("ternary codes need the non-zero plane");
  if (bare == 3 || bare == 4) {
    // paired planes: two gallery rows per plane row (3), or the query plane that goes with them (4)
    const int kbp = ch_tc_code_bytes_pair(nbit, ternary);
    if (kbp == 0) CH_FAIL("nbit=%d ternary=%d has no paired form (binary <= 128 bits, ternary <= 64)", nbit, ternary);
    if (bare == 4 && thresh_dev == nullptr) CH_FAIL("the paired query plane needs the thresholds");
    if (bare == 3 && thresh_dev != nullptr) CH_FAIL("a gallery plane takes no thresholds");
    if (rows_out % 32 || rows_bits < 0 || (bare == 3 ? 2 * rows_out < rows_bits : rows_out < rows_bits))
      CH_FAIL("rows_out must be a multiple of 32 and cover rows_bits (gallery: two rows per plane row)");
    if (rows_out == 0) return 0;
    ChDeviceGuard guard(ws->device);
    const long long n = rows_out * (kbp / 16);
    expand_i8_pair_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        bits_dev, ternary ? nz_dev : nullptr, rows_bits, rows_out, ch_code_words(nbit), nbit, kbp, ternary ? 1 : 0,
        bare == 4 ? 1 : 0, thresh_dev, nq, out_dev, nullptr, 0);
    CH_LAUNCH_CHECK(ws);
    return 0;
  }
  if (bare && thresh_dev != nullptr) CH_FAIL("a plane without threshold slots takes no thresholds");
  if (bare < 0 || bare > 2)
    CH_FAIL("bare: 0 = threshold slots, 1 = bare gallery plane, 2 = bare query plane, 3 / 4 = paired gallery / query");
  const int kb = bare ? ch_tc_code_bytes_bare(nbit) : ch_tc_code_bytes(nbit);
  if (kb == 0) CH_FAIL("nbit=%d unsupported by the tensor-core path (1..%d)", nbit, CH_MAX_NBIT);
  if (rows_out % 32 || rows_bits < 0 || rows_out < rows_bits)
    CH_FAIL("rows_out must be a multiple of 32 (whole permuted row blocks) and >= rows_bits");
  if (rows_out == 0) return 0;
  ChDeviceGuard guard(ws->device);
  const long long n = rows_out * (kb / 16);
  expand_i8_tiled_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bits_dev, ternary ? nz_dev : nullptr, rows_bits, rows_out, ch_code_words(nbit), nbit, kb,
      bare ? 0 : thresh_slots(nbit), ternary ? 1 : 0, (thresh_dev != nullptr || bare == 2) ? 1 : 0, thresh_dev, nq,
      out_dev);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_expand_i8_query_stripes(ch_ws* ws, const uint32_t* bits_dev, const uint32_t* nz_dev, int64_t rows_bits,
                                          int nbit, int ternary, int8_t* out_dev, int64_t rows_out,
                                          const uint32_t* thresh_dev, const uint32_t* scut_dev, int stripe0, int nstripes,
                                          int64_t nq, void* stream) {
  if (ws == nullptr || bits_dev == nullptr || out_dev == nullptr || thresh_dev == nullptr || scut_dev == nullptr)
    CH_FAIL("null argument to ch_expand_i8_query_stripes");
  if (ternary && nz_dev == nullptr) CH_FAIL("ternary codes need the non-zero plane");
  const int kbp = ch_tc_code_bytes_pair(nbit, ternary);
  if (kbp == 0) CH_FAIL("nbit=%d ternary=%d has no paired form (binary <= 128 bits, ternary <= 64)", nbit, ternary);
  if (rows_out % 32 || rows_bits < 0 || rows_out < rows_bits || rows_out == 0)
    CH_FAIL("rows_out must be a positive multiple of 32 and >= rows_bits");
  if (nstripes <= 0 || nstripes > 65535 || stripe0 < 0) CH_FAIL("bad stripe range");
  ChDeviceGuard guard(ws->device);
  const long long n = rows_out * (kbp / 16);
  const dim3 grid(static_cast<unsigned>((n + 255) / 256), static_cast<unsigned>(nstripes));
  expand_i8_pair_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bits_dev, ternary ? nz_dev : nullptr, rows_bits, rows_out, ch_code_words(nbit), nbit, kbp, ternary ? 1 : 0, 1,
      thresh_dev, nq, out_dev, scut_dev, stripe0);
  CH_LAUNCH_CHECK(ws);
  return 0;
}

extern "C" int ch_hamming_select_tc(ch_ws* ws, const ch_select_args* a, void* stream) {
  if (ws == nullptr || a == nullptr) CH_FAIL("null argument to ch_hamming_select_tc");
  if (a->q_i8 == nullptr || a->g_i8 == nullptr || a->cand_off == nullptr || a->cand_cap == nullptr ||
      a->cand_cnt == nullptr || a->cand_rows == nullptr || a->err_flag == nullptr)
    CH_FAIL("null array in ch_select_args");
  const int tau = a->thresh != nullptr ? 1 : 0;       // comparison in the epilogue: planes without threshold slots
  const int pair = a->pair != 0 ? 1 : 0;              // two gallery rows per plane row (ch_expand_i8 forms 3 / 4)
  if (pair && tau) CH_FAIL("the paired form carries its thresholds in the contraction (thresh must be NULL)");
  const int kb = pair ? ch_tc_code_bytes_pair(a->nbit, a->ternary)
                      : (tau ? ch_tc_code_bytes_bare(a->nbit) : ch_tc_code_bytes(a->nbit));
  if (kb == 0) CH_FAIL("nbit=%d unsupported by the tensor-core path (1..%d; paired: binary <= 128, ternary <= 64)",
                       a->nbit, CH_MAX_NBIT);
  if (tau && a->dense) CH_FAIL("the epilogue-side comparison exists for the sparse epilogue only");
  if (a->nq <= 0 || a->ndb < 0) CH_FAIL("bad arguments");
  if (a->nq_pad % kTileM || a->nq_pad < a->nq) CH_FAIL("nq_pad must be a multiple of %d and >= nq", kTileM);
  if (a->nstripes <= 0 || a->rows_per_stripe <= 0 || a->rows_per_stripe % (pair ? 2 * kTileN : kTileN) ||
      static_cast<long long>(a->nstripes) * a->rows_per_stripe < a->ndb)
    CH_FAIL("bad stripe geometry");
  if (a->row_base < 0 || a->row_base + a->ndb > 0x7fffffffll) CH_FAIL("shard-local row indices must fit 31 bits");
  if ((reinterpret_cast<uintptr_t>(a->q_i8) | reinterpret_cast<uintptr_t>(a->g_i8)) & 15)
    CH_FAIL("operands must be 16-byte aligned");
  ChDeviceGuard guard(ws->device);
  SelDev d;
  d.q_i8 = a->q_i8; d.g_i8 = a->g_i8;
  d.cand_off = a->cand_off; d.cand_cap = a->cand_cap; d.cand_cnt = a->cand_cnt; d.cand_rows = a->cand_rows;
  d.err_flag = a->err_flag;
  d.thresh = a->thresh; d.nbit = a->nbit; d.ternary = a->ternary; d.bad = a->bad;
  d.nq = a->nq; d.nq_pad = a->nq_pad; d.ndb = a->ndb; d.row_base = a->row_base;
  d.rows_per_stripe = a->rows_per_stripe;
  d.q_stripe_bytes = a->q_stripe_bytes;
  if (a->q_stripe_bytes < 0 || (a->q_stripe_bytes & 15)) CH_FAIL("bad q_stripe_bytes");
  d.nqtiles128 = static_cast<int>(a->nq_pad / kTileM);
  d.nqgroups = (d.nqtiles128 + kQT - 1) / kQT;
  size_t smem = 0;
  sel_fn_t fn = pick_sel(kb, a->dense != 0, tau, pair, &smem);
  if (smem > static_cast<size_t>(ws->max_smem_optin)) CH_FAIL("tensor-core kernel needs %zu bytes of shared memory", smem);
  if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: each CTA owns all 512 TMEM columns
  CH_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const long long ctas = static_cast<long long>(d.nqgroups) * a->nstripes;
  if (ctas > 0x7fffffffll) CH_FAIL("grid too large");
  fn<<<static_cast<unsigned>(ctas), 128 + 128 * kQT, smem, static_cast<cudaStream_t>(stream)>>>(d);
  CH_LAUNCH_CHECK(ws);
  return 0;
}
