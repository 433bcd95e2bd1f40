/*
 * concepthash_b200 -- C-ABI of the B200-native retrieval-evaluation hot path
 * (real-valued hash codes -> sign/bit-pack -> Hamming ranking -> exact top-R -> mAP@R / P@k / R@k).
 *
 * Drop-in boundary.  The reference (kamwoh/concepthash) is pure Python; the interface this
 * library sits under is the import
 *     from utils.hashing import calculate_mAP, calculate_pr_curve      experiments/test_hashing.py:15
 *     from utils.hashing import calculate_mAP                          experiments/train_helper.py:18
 * and the two call sites experiments/test_hashing.py:106-119 and experiments/train_helper.py:228-234.
 * `utils/hashing.py` in this repository is the ctypes binding a maintainer drops into the reference's
 * `utils/` directory.  Every entry point below is one step of that function;
 * the reference-side statement of the step it replaces is cited next to it.
 *
 * Conventions
 *   - plain C, `extern "C"`, no torch / C++ types in any signature;
 *   - every function returns 0 on success, non-zero on failure; `ch_last_error()` gives the message
 *     (thread-local); nothing aborts, nothing falls back to the CPU;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - all `*_dev` pointers are device pointers owned by the caller (the Python side allocates them as
 *     torch tensors: torch is plumbing for memory / streams / NCCL only);
 *   - "rows_pad" = ch_padded_rows(n): packed arrays are allocated with that many rows, pad rows are
 *     written as zero by the pack kernels (the streaming kernels bulk-copy whole 16-byte groups);
 *   - sort key of a (query, gallery) pair: key = Hamming distance (binary codes) or
 *     key = 2 * distance (ternary codes: sign(0) = 0 gives half-integer distances);
 *     nbins = nbit + 1 (binary) or 2 * nbit + 1 (ternary);
 *   - canonical ranking: ascending (key, global gallery row index)  (SURVEY.md F4).
 */
#ifndef CONCEPTHASH_B200_H
#define CONCEPTHASH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CH_ABI_VERSION 5
#define CH_MAX_NBIT 256          /* words per code: 1, 2, 4 or 8 x u32 */
#define CH_MAX_R 8               /* length of an `R` list (test_hashing.py:124-128) */
#define CH_MAX_PR 32             /* length of `PRs` */

/* element types of caller tensors */
enum { CH_F32 = 0, CH_F16 = 1, CH_BF16 = 2, CH_F64 = 3, CH_I64 = 4, CH_I32 = 5, CH_U8 = 6, CH_I16 = 7, CH_I8 = 8 };
/* where a caller buffer lives */
enum { CH_MEM_DEVICE = 0, CH_MEM_HOST = 1 };
/* label representation after ch_pack_labels */
enum { CH_LAB_NONE = 0, CH_LAB_ID = 1, CH_LAB_MASK = 2 };
/* what the histogram pass records per (query, stripe) */
enum { CH_EMIT_NONE = 0, CH_EMIT_RELEVANT = 1, CH_EMIT_CANDIDATES = 2 };

#define CH_QUERY_NOLABEL   0xFFFFFFFFu   /* id of a query row without any positive class  */
#define CH_GALLERY_NOLABEL 0xFFFFFFFEu   /* id of a gallery row without any positive class */

typedef struct ch_ws ch_ws;     /* per-process / per-GPU workspace: device props, H2D staging, copy stream */

int         ch_abi_version(void);
const char* ch_last_error(void);
int         ch_workspace_create(int device, ch_ws** out);
int         ch_workspace_destroy(ch_ws* ws);
int         ch_device_info(ch_ws* ws, int* sm_count, int* max_smem_optin, int* l2_bytes, int* clock_khz);

/* rows a packed array must be allocated with */
int64_t ch_padded_rows(int64_t n);
/* u32 words per packed code row for `nbit` (1, 2, 4 or 8); 0 if unsupported */
int     ch_code_words(int nbit);

/* ---- K1: sign + bit-pack ---------------------------------------------------------------------
 * Replaces `torch.sign(codes)` (+ `codes[|codes| < threshold] = 0` when threshold != 0) of the
 * missing utils.hashing.calculate_mAP; in-tree statements: models/layers/signhash.py:11,
 * trainers/orthohash.py:78; threshold: experiments/test_hashing.py:109, configs/val.yaml:12.
 *   codes   : (n, nbit) real-valued, element (i, k) at codes[i * row_stride + k * col_stride]
 *             (strides in elements; non-contiguous slices of test_hashing.py:91-98 are fine);
 *             CH_MEM_HOST buffers are staged through the workspace in pipelined chunks.
 *   threshold: already rounded to the dtype of `codes` by the caller (torch compares in that dtype)
 *   col_sub : NULL, or device f64[nbit] subtracted from every row before the sign (`zero_mean_eval`,
 *             experiments/train_helper.py:223-226, test_hashing.py:100-103: both sets minus the gallery's column
 *             mean); the difference is taken in fp32 (fp64 for fp64 codes), like `codes - mean` in torch, so the
 *             sign is that of the exact difference.  Values already rounded to the dtype of `codes`.
 *   out_bits: (rows_pad, words) u32, bit (k % 32) of word k / 32 = codes[i, k] > 0
 *   out_nz  : same shape, bit = sign(codes[i, k]) != 0; may be NULL only if the caller knows there
 *             are no zeros (it is still checked: flags bit 0)
 *   flags   : device u32, OR-ed: bit 0 = some sign is 0 (ternary needed), bit 1 = NaN seen
 */
int ch_pack_sign(ch_ws* ws, const void* codes, int mem, int dtype, int64_t n, int nbit,
                 int64_t row_stride, int64_t col_stride, double threshold, const double* col_sub_dev,
                 uint32_t* out_bits_dev, uint32_t* out_nz_dev, uint32_t* flags_dev, void* stream);
/* The host half of ch_pack_sign for PAGEABLE fp32 codes -- what `inference_one_epoch` returns, a torch.cat of
 * `.cpu()` batches (trainers/base.py:291-304).  Nothing can DMA from pageable memory, the host's cores have to
 * touch every byte once anyway; they apply the sign test in that pass (AVX-512 / AVX2, `threads` threads) and only
 * the bits cross PCIe.  ch_pack_sign does this internally for CH_MEM_HOST (pageable or pinned) + CH_F32 +
 * threshold 0 + no col_sub + no non-zero plane; this entry point exposes the same routine with host output
 * (unit tests; callers that pack while the loader is still producing batches).  out_bits_host: (n, words) u32,
 * same bit order as ch_pack_sign; *flags_host |= 1 (a zero) / 2 (NaN).  Bit-identical to the kernels. */
int ch_host_pack_sign(const float* codes_host, int64_t n, int nbit, int64_t row_stride,
                      uint32_t* out_bits_host, uint32_t* flags_host, int threads);
/* Host loader: ch_host_pack_sign for whole HOST-resident fp32 arrays on a thread of its own, started before anything
 * else of the evaluation (the host's cores are the bottleneck of an evaluation whose codes arrive as the `torch.cat`
 * of `.cpu()` batches of trainers/base.py:291-304: they read 4 bytes per code, the GPU is done in less).  Up to
 * CH_LOADER_MAX_JOBS arrays are packed in order (the evaluator passes queries, row sample, gallery) in ~256 KB
 * pieces by `ch_host_pack_threads(ws)` - 1 threads into a ring of pinned chunks; every complete chunk is copied to
 * its job's out_bits_dev (rows_pad x words u32, the layout of ch_pack_sign; pad rows zeroed) on `stream`, in order,
 * followed by an event.  Zeros / NaNs of a job are OR-ed into its *flags_dev
 * (bits as in ch_pack_sign) on `stream` before the job's last chunk.  The arrays (pageable or pinned) must stay
 * valid until ch_host_loader_join.
 *   ch_host_loader_wait : makes `stream` (NOT the loader's own stream) wait until rows [0, rows) of `job` have landed:
 *                         blocks the calling thread until their copy has been queued, then waits for its event.
 *                         (Only with CH_LOADER_STREAM_OPS=1 and block = 0 it returns at once -- a cuStreamWaitValue32
 *                         on the job's progress word; off by default, such waits can deadlock, see loader.cu.)
 *   ch_host_loader_join : waits for the thread, frees the handle; flags_host[j] = the flag bits of job j (host side) */
#define CH_LOADER_MAX_JOBS 6
typedef struct ch_loader ch_loader;
enum { CH_LOADER_PACK = 0, CH_LOADER_COPY = 1 };
typedef struct ch_loader_job {
  const void* codes_host;    /* PACK: (n, nbit) fp32, unit column stride, 4-byte aligned.  COPY: n rows of raw bytes */
  int64_t n;
  int32_t nbit;              /* PACK: bits per row.  COPY: BYTES per row (a multiple of 4, e.g. 8 = int64 class ids) */
  int32_t kind;              /* CH_LOADER_PACK | CH_LOADER_COPY (rows travel as they are: the label ids of a side) */
  int64_t row_stride;        /* PACK: in elements, >= nbit.  COPY: in bytes, >= nbit */
  uint32_t* out_bits_dev;    /* PACK: (ch_padded_rows(n), ch_code_words(nbit)) u32.  COPY: n x nbit bytes, dense */
  uint32_t* flags_dev;       /* PACK: device u32, OR-ed; may be NULL.  COPY: unused */
} ch_loader_job;
int ch_host_pack_threads(ch_ws* ws);
int ch_host_loader_start(ch_ws* ws, const ch_loader_job* jobs, int njobs, void* stream, ch_loader** out);
int ch_host_loader_wait(ch_loader* loader, int job, int64_t rows, void* stream, int block);
int ch_host_loader_join(ch_loader* loader, uint32_t* flags_host);
/* column sums (fp64, deterministic) of DEVICE codes (n, ncols): the numerator of `db_codes.mean(dim=0)`
 * (zero_mean_eval); with a row-sharded gallery the caller all-reduces the sums over ranks. */
int ch_column_sums(ch_ws* ws, const void* codes_dev, int dtype, int64_t n, int ncols, int64_t row_stride,
                   int64_t col_stride, double* sums_dev, void* stream);

/* ---- labels ----------------------------------------------------------------------------------
 * Relevance in the reference = "share >= 1 positive class" on (N, C) one-/multi-hot labels
 * (missing utils.hashing; same notion as models/loss/dpsh.py:58, trainers/adsh.py:144).
 *   labels (n, C): positives are entries > 0.  C == 0 means 1-D integer class ids (n,).
 *   out_ids  : (rows_pad) u32 = index of the first positive (or `nolabel`)
 *   out_masks: (rows_pad, mask_words) u32 bitmask of positives, mask_words = ceil(C / 32); may be
 *              NULL when C == 0
 *   info     : device u32[4]: [0] max positives per row (atomicMax), [1] max id + 1,
 *              [2] rows with no positive, [3] reserved
 */
int ch_pack_labels(ch_ws* ws, const void* labels, int mem, int dtype, int64_t n, int C,
                   int64_t row_stride, int64_t col_stride, uint32_t nolabel,
                   uint32_t* out_ids_dev, uint32_t* out_masks_dev, uint32_t* info_dev, void* stream);

/* ---- K2: Hamming + per-query key histogram (+ stable in-bucket prefixes) ---------------------
 * Replaces the 32-row chunked `0.5 * (nbit - q @ chunk.T)` GEMM loop + the dense (nq, ndb) matrix
 * (identity: trainers/orthohash.py:263-264 `get_hd`, trainers/orthohash.py:49) and is the counting
 * half of `torch.topk(dist, R, largest=False)` (sibling in tree: utils/metrics.py:20).
 * One thread owns one query; a CTA owns `threads` queries x one gallery stripe (contiguous rows,
 * streamed through shared memory by bulk async copies); the per-thread histogram lives in shared
 * memory as {lo16 = all items, hi16 = relevant items} per key and is flushed to the stripe slabs.
 */
typedef struct ch_hist_args {
  /* packed inputs */
  const uint32_t* q_bits;  const uint32_t* q_nz;   /* (nq_pad', words); q_nz NULL unless ternary */
  const uint32_t* g_bits;  const uint32_t* g_nz;   /* (rows_pad, words) local gallery shard       */
  const uint32_t* q_lab;   const uint32_t* g_lab;  /* ids (n) or masks (n, mask_words); NULL if CH_LAB_NONE */
  /* outputs: stripe slabs, zero-initialised by the caller; ONE call writes every (stripe, key, query) entry it
   * owns (a call does not accumulate into the counts of an earlier call) */
  uint32_t* slab_all;      /* (nstripes, nbins, nq_pad) */
  uint32_t* slab_rel;      /* same; NULL if label_mode == CH_LAB_NONE */
  /* thresholded ("select") mode: only pairs with key <= thresh[q] are counted / recorded */
  const uint32_t* thresh;  /* (nq_pad) or NULL = count every pair */
  /* records (16 B each: {key | rel << 31, in-stripe prefix all, in-stripe prefix relevant, local row}) */
  const uint32_t* rec_off; /* (nstripes, nq_pad) first record slot of (stripe, query) */
  const uint32_t* rec_cap; /* (nstripes, nq_pad) capacity of (stripe, query)          */
  uint32_t* rec_cnt;       /* (nstripes, nq_pad) records written                      */
  void*     recs;          /* uint4[]                                                 */
  uint32_t* err_flag;      /* device u32, set non-zero on record overflow             */
  int64_t nq, nq_pad, ndb;             /* ndb = rows of the local shard */
  int32_t nbit, ternary, label_mode, mask_words, emit_mode;
  int32_t nstripes, threads, rows_per_stripe;   /* geometry from ch_hist_geometry */
  int64_t row_base;        /* added to the local row index stored in records (streamed galleries: the shard is
                              processed in row blocks, one call per block) */
  int32_t key_limit;       /* select pass only: max(thresh) + 1 if the caller knows it, else 0.  When set, keys >= key_limit
                              cannot occur and the slabs are (nstripes, key_limit, nq_pad) instead of (.., nbins, ..) */
} ch_hist_args;

/* chooses threads per CTA (= queries per CTA), nq_pad, stripes and rows per stripe (a multiple of 256)
 * for a problem on this device; call it once per evaluation with the richest label mode used. */
int ch_hist_geometry(ch_ws* ws, int64_t nq, int64_t ndb, int nbit, int ternary, int label_mode,
                     int mask_words, int32_t* threads, int64_t* nq_pad, int32_t* nstripes,
                     int32_t* rows_per_stripe);
int ch_hamming_hist(ch_ws* ws, const ch_hist_args* a, void* stream);

/* ---- K2, tensor-core form of the select pass -------------------------------------------------------
 * Same question as ch_hamming_hist with `thresh` set -- which pairs have key <= thresh[q]? -- answered by
 * tcgen05.mma kind::i8 (UTCIMMA) with the accumulators in TMEM: for +-1 codes held as int8,
 * <q, g> = nbit - 2 * hamming; for ternary codes ({-1, 0, +1}: `threshold` of configs/val.yaml:12 /
 * experiments/test_hashing.py:109, or an exact zero) the same contraction gives key = 2 * distance = nbit - <q, g>.
 * nbit <= 256.  The int8 planes are made from the packed bits (and the non-zero plane of ternary codes)
 * by ch_expand_i8 in the shared-memory operand order (8-row x 16-byte core matrices), so the kernel streams
 * them with plain 1-D bulk copies.  ch_tc_code_bytes(nbit) = bytes per plane row: the codes plus the
 * "threshold slots" (2 bytes from nbit on; 4 for nbit > 254), rounded up to whole 32-byte K blocks; 0 if nbit is
 * unsupported.  The slots hold 1 on the gallery side and int8 values that sum to -tau on the query side
 * (tau = nbit - 2 * thresh[q], or nbit - thresh[q] on the doubled key scale of ternary codes), so
 * that the accumulator is  <q, g> - (nbit - 2 * thresh[q])  and  key <= thresh  <=>  accumulator >= 0: the
 * epilogue only looks at sign bits.
 *   ch_expand_i8: bits (rows_bits, words) -> out (rows_out, kb) int8, rows_out >= rows_bits, rows_out % 32 == 0
 *                 (rows past rows_bits get zero codes); thresh_dev == NULL: gallery plane, every 32-row block
 *                 stored in the row order the kernel's sign-bit folding wants (plane row c of a block = block
 *                 row 8 (c & 3) + (c >> 2)); else query plane (thresh (>= nq) u32; rows >= nq are padding
 *                 queries that accept nothing).
 * Output = candidate lists: for every (stripe, query) the shard-local row indices (row_base + row) of the
 * pairs with key <= thresh[q], ascending, in the slice [cand_off, cand_off + cand_cap) of cand_rows;
 * cand_cnt = number written; err_flag |= 1 if a slice overflowed (capacities: ch_record_caps source 0,
 * offsets: ch_record_offsets).  Keys, label matches, ranks and AP follow in ch_cand_hist / ch_cand_finalize. */
typedef struct ch_select_args {
  const int8_t* q_i8;         /* query plane  (>= nq_pad rows) */
  const int8_t* g_i8;         /* gallery plane (rows_pad rows)  */
  const uint32_t* cand_off;   /* (nstripes, nq_pad) */
  const uint32_t* cand_cap;   /* (nstripes, nq_pad) */
  uint32_t* cand_cnt;         /* (nstripes, nq_pad) */
  uint32_t* cand_rows;        /* u32[] */
  uint32_t* err_flag;
  const uint32_t* thresh;     /* NULL: the thresholds ride in the contraction (planes with threshold slots,
                                 ch_tc_code_bytes).  Else (nq_pad) keys: the planes carry NO slots
                                 (ch_tc_code_bytes_bare: one K block less at nbit = 64 / 128 / ...), the accumulator
                                 is the plain <q, g> and the epilogue compares it with tau = nbit - 2 thresh[q]
                                 (ternary: nbit - thresh[q]) by packed 16-bit max / subtract; sparse epilogue only */
  int64_t nq, nq_pad, ndb, row_base;   /* nq_pad % 128 == 0; ndb = rows of this call's row block */
  int32_t nbit, nstripes, rows_per_stripe;   /* rows_per_stripe % 128 == 0 */
  int32_t dense;              /* != 0: most 32-row chunks of a warp hold a candidate -> skip the max-tree filter */
  int32_t ternary;            /* keys on the doubled scale (read when thresh != NULL or pair != 0) */
  int32_t pair;               /* != 0: the planes are the PAIRED forms of ch_expand_i8 (3 / 4): a plane row of the gallery
                                 holds two gallery rows, a 16-bit accumulator value both comparisons (bit 7 / bit 15),
                                 a tile is 256 rows; rows_per_stripe % 256 == 0, thresh must be NULL */
  uint32_t* bad;              /* (nq_pad) or NULL: bad[q] = 1 for every query one of whose slices overflowed, so that
                                 the caller can re-rank just those queries by the exact path */
  int64_t q_stripe_bytes;     /* 0: one query plane for all stripes; else q_i8 holds one plane per stripe of this launch
                                 (ch_expand_i8_query_stripes), this many bytes apart */
} ch_select_args;
int ch_tc_code_bytes(int nbit);
/* queries one CTA of the select kernel owns (its grid is ceil(nq_pad / this) x nstripes, one CTA per SM) */
int ch_tc_queries_per_cta(void);
int ch_tc_code_bytes_bare(int nbit);   /* bytes per plane row without threshold slots (see ch_select_args.thresh) */
/* Paired planes: TWO gallery rows per plane row, so an accumulator (128 TMEM columns) covers 256 gallery rows and the
 * select kernel has half as many accumulator hand-overs, TMEM loads and barrier round trips per pair.  With
 * f = thresh - key in [-128, 127] (keys <= 128: binary codes of <= 128 bits, ternary codes of <= 64) the contraction
 * yields E = (f_a + 128) + 256 (-f_b - 1) in int16 range for the rows a, b of a plane row: bit 7 of the packed 16-bit
 * accumulator says "a is a candidate", bit 15 "b is".  Query plane row = [q, 64 q, 5 slots], gallery plane row =
 * [a, -4 b, (2, 1, -128, -128, -128)], a / b as {0, 1} for binary codes (<q, g01> = ones(q) - hamming) and {-1, 0, +1}
 * for ternary ones; plane row c of a 32-row block holds gallery rows 64 (block) + 32 (c >> 4) + 31 - 8 (2 (c & 1) + f)
 * - ((c >> 1) & 7), f = 0 / 1.  ch_tc_code_bytes_pair = bytes per plane row (0: no paired form for this nbit). */
int ch_tc_code_bytes_pair(int nbit, int ternary);
int ch_tc_tile_rows(int pair);         /* gallery rows per tile of the select kernel: 128, paired 256 */
int ch_expand_i8(ch_ws* ws, const uint32_t* bits_dev, const uint32_t* nz_dev /* ternary codes, else NULL */,
                 int64_t rows_bits, int nbit, int ternary,
                 int bare /* 0: threshold slots; 1: gallery plane without; 2: query plane without;
                             3: paired gallery plane (rows_out = PLANE rows >= rows_bits / 2);
                             4: paired query plane (thresh_dev required) */,
                 int8_t* out_dev, int64_t rows_out, const uint32_t* thresh_dev /* or NULL */, int64_t nq,
                 void* stream);
/* the paired query plane (form 4) once per stripe: out (nstripes, rows_out, kb); plane s carries thresh[q] for the
 * stripes stripe0 + s < scut[q] and thresh[q] - 1 (below 0: nothing is a candidate) for the others.  The select
 * kernel reads plane `stripe` of a launch when ch_select_args.q_stripe_bytes = rows_out * kb is set. */
int ch_expand_i8_query_stripes(ch_ws* ws, const uint32_t* bits_dev, const uint32_t* nz_dev, int64_t rows_bits, int nbit,
                               int ternary, int8_t* out_dev, int64_t rows_out, const uint32_t* thresh_dev,
                               const uint32_t* scut_dev, int stripe0, int nstripes, int64_t nq, void* stream);
int ch_hamming_select_tc(ch_ws* ws, const ch_select_args* a, void* stream);

/* ---- K3/K4 on candidate lists ---------------------------------------------------------------------
 * The rest of `torch.topk` + the per-query numpy AP loop of the missing calculate_mAP, on the candidates only.
 * Every item with key <= thresh[q] is a candidate, so canonical ranks are counts over candidates:
 *   ch_cand_hist     per candidate: key = popcount(q ^ g) on the packed codes, relevance by the labels (ids or
 *                    masks); writes cand_key, ORs bit 31 of cand_rows for relevant candidates, and the per-query
 *                    key totals tot_all / tot_rel (nbins, nq_pad) of this rank are ACCUMULATED into the
 *                    caller-zeroed arrays (a list may be histogrammed in several calls, one per row block).
 *                    err_flag |= 2 if a candidate has key >= nbins.
 *   (caller: all-gather the totals over ranks, ch_scan_bases -> base0_all / base0_rel)
 *   ch_cand_finalize walks each query's list in order (stripe-major = ascending row): rank = base0_all[key] +
 *                    #earlier candidates of the same key; relrank likewise on relevant ones.
 *                    mode 0: cols (nq, 2 nR + nPR) AP sums / counts / hits as in ch_finalize_records;
 *                    mode 1: first_rel_out[q] = 1 iff the rank-0 item is relevant (remove_first_retrieved);
 *                    mode 2: ids (nq, R) int64 = row_offset + row at its rank, keys (nq, R) int32 (pre-filled -1). */
typedef struct ch_cand_args {
  const uint32_t* cand_off;  const uint32_t* cand_cnt;    /* (nstripes, nq_pad) */
  uint32_t* cand_rows;       uint16_t* cand_key;          /* per candidate slot */
  const uint32_t* q_bits;    const uint32_t* g_bits;      /* packed codes (hist only) */
  const uint32_t* q_nz;      const uint32_t* g_nz;        /* non-zero planes of ternary codes (keys = 2 x distance,
                                                             0 .. 2 nbit), else NULL */
  const uint32_t* g_plane;   /* optional (CH_LAB_ID): ch_gather_plane output -- code + class id in ONE sector */
  const uint32_t* q_lab;     const uint32_t* g_lab;       /* ids or masks; NULL if CH_LAB_NONE (hist only) */
  uint32_t* tot_all;         uint32_t* tot_rel;           /* (nbins, nq_pad) out of ch_cand_hist */
  const uint32_t* base0_all; const uint32_t* base0_rel;   /* (nbins, nq_pad) in of ch_cand_finalize */
  const uint32_t* first_rel; uint32_t* first_rel_out;     /* (nq_pad) */
  const uint32_t* key_max;   /* (nq_pad) or NULL (ch_cand_finalize): thresh output of ch_scan_bases for rmax = the
                                longest list asked for -- candidates with a larger key are skipped */
  double*  cols;             /* (nq, ncols) */
  int64_t* ids;  int32_t* keys;   /* (nq, R) */
  uint32_t* err_flag;
  int64_t nq, nq_pad, R, row_offset;
  int32_t nstripes, nbins, nbit, label_mode, mask_words, remove_first, nR, nPR, mode;
  int64_t r_eff[CH_MAX_R];
  int64_t pr_k[CH_MAX_PR];
} ch_cand_args;
int ch_cand_hist(ch_ws* ws, const ch_cand_args* a, void* stream);
/* gather plane of a single-label gallery shard: out (rows, ch_gather_plane_words(nbit)) u32 = [code words | class
 * id | zero pad], 8 .. 64 bytes per row, so that the key + label match of a candidate costs one memory sector
 * instead of two random gathers (ch_cand_hist is bound by L2 sectors) */
int ch_gather_plane_words(int nbit);
int ch_gather_plane(ch_ws* ws, const uint32_t* bits_dev, const uint32_t* ids_dev, int64_t rows, int nbit,
                    uint32_t* out_dev, void* stream);
int ch_cand_finalize(ch_ws* ws, const ch_cand_args* a, void* stream);
/* ch_cand_hist + ch_scan_bases_pair + ch_cand_finalize (mode 0, no remove_first) of a SINGLE rank in one kernel: a warp
 * histograms the keys of its query's candidates (cand_key and the relevance bits are written as by ch_cand_hist), scans
 * the histogram into the bases in shared memory, verifies the list -- fewer than `need` candidates: bad[q] = 1,
 * status |= 1 -- and walks it (candidates beyond the key at which the list holds `rmax` items are skipped).
 * tot_*, base0_*, key_max of the arguments are not used; cols (nq, 2 nR + nPR) as in ch_cand_finalize. */
int ch_cand_rank(ch_ws* ws, const ch_cand_args* a, int64_t rmax, int64_t need, uint32_t* status_dev, uint32_t* bad_dev,
                 void* stream);
/* Two-level threshold sampling: the candidate list (list_stripes slices per query, rows = SAMPLE row indices, keys
 * already written by ch_cand_hist) came from a select pass over a 1-in-`sample_stride` row sample.
 * cap[s][q] = sample_stride * (k + 6 sqrt(k + 1) + 9), k = #sample candidates of query q with key <= thresh[q]
 * whose sample row lies in [s * rows_per_stripe, (s + 1) * rows_per_stripe): the capacity of slice (s, q) of the
 * full pass (same bound as ch_record_caps with sample_stride).  cap_dev: (nstripes, nq_pad).
 * scut_dev != NULL (nq_pad): the threshold is refined to a (key, stripe) pair -- candidates are a prefix of the
 * canonical (key, row) order and stripes are row blocks, so "key < thresh, or key == thresh in a stripe < scut" is a
 * prefix as well: scut[q] = the smallest stripe count at which the sample holds >= m items of that prefix (nstripes
 * if it never does); stripes >= scut[q] are to be selected with thresh[q] - 1 (ch_expand_i8_query_stripes) and their
 * capacities count keys < thresh only. */
int ch_cand_caps(ch_ws* ws, const uint32_t* cand_off, const uint32_t* cand_cnt, const uint32_t* cand_rows,
                 const uint16_t* cand_key, const uint32_t* thresh, int list_stripes, int rows_per_stripe,
                 int nstripes, int64_t nq, int64_t nq_pad, int sample_stride, uint32_t* cap_dev,
                 int m, uint32_t* scut_dev, void* stream);

/* slab reductions: totals over stripes -> tot (nbins, nq_pad); exclusive scan over stripes in place */
int ch_slab_totals(ch_ws* ws, const uint32_t* slab, int nstripes, int nbins, int64_t nq_pad,
                   uint32_t* tot_dev, void* stream);
int ch_slab_exscan(ch_ws* ws, uint32_t* slab, int nstripes, int nbins, int64_t nq_pad, void* stream);
/* both steps for the {all, relevant} slabs of a full-ranking pass in one launch: slab (nslabs, nstripes, nbins,
 * nq_pad) is scanned in place over the stripes, tot (nslabs, nbins, nq_pad) receives the per-rank totals */
int ch_slab_scan(ch_ws* ws, uint32_t* slab, int nslabs, int nstripes, int nbins, int64_t nq_pad,
                 uint32_t* tot_dev, void* stream);

/* ---- K3: exact top-R selection ----------------------------------------------------------------
 * Counting half of `torch.topk` finished: from the per-rank totals (all-gathered over ranks, laid out
 * (world, nbins, nq_pad)) derive per query
 *   base0[key][q] = #items with smaller key on all ranks + #items with equal key on lower ranks,
 *   thresh[q]     = smallest key t with #(key <= t) >= min(rmax, ndb_total)   (if thresh_out != NULL),
 *   total[q]      = number of counted items over all ranks                     (if total_out != NULL).
 * Gallery rows are partitioned contiguously in rank order, so (rank, stripe, in-stripe prefix) is the
 * ascending-global-row-index tie order.
 */
int ch_scan_bases(ch_ws* ws, const uint32_t* tot_all_dev, int world, int rank, int nbins,
                  int64_t nq, int64_t nq_pad, int64_t rmax, uint32_t* base0_dev,
                  uint32_t* thresh_out_dev, uint32_t* total_out_dev, void* stream);

/* record capacities per (stripe, query) and their offsets.
 *   source 0: candidates  = sum_{key <= thresh[q]} slab_all[s][key][q]   (un-scanned slab)
 *   source 1: relevant    = sum_key slab_rel[s][key][q]                  (un-scanned slab)
 *   source 2: class count = cls_cnt[s][q_ids[q]]                          (single-label fast path)
 *   min_with_prev != 0 keeps min(existing cap, new cap).
 *   sample_stride > 1: the counts come from a 1-in-stride row sample; they are scaled to a high-probability
 *   bound stride * (k + 6 sqrt(k + 1) + 9) (an overflow is detected by the pass and reported in err_flag).
 */
int ch_record_caps(ch_ws* ws, int source, const uint32_t* slab_or_cls, const uint32_t* thresh_or_qids,
                   int nstripes, int nbins_or_nclass, int64_t nq, int64_t nq_pad, int min_with_prev,
                   int sample_stride, int src_stripes /* nstripes, or 1 = replicate the one source stripe */,
                   uint32_t* cap_dev, void* stream);
/* off[s][q] = start[q] + sum_{s' < s} cap[s'][q]; start = exclusive scan of per-query totals;
 * total_host receives the total number of record slots (the call synchronises the stream).  Offsets are 32-bit:
 * a total >= 2^32 - 16 is reported, not an error -- the caller splits the query set (AP is per-query). */
int ch_record_offsets(ch_ws* ws, const uint32_t* cap_dev, int nstripes, int64_t nq, int64_t nq_pad,
                      uint32_t* off_dev, uint64_t* total_host, const uint32_t* thresh_dev /* or NULL */,
                      uint32_t* thresh_max_host /* max(thresh[0..nq)) rides on the same sync */, void* stream);
/* The same offsets WITHOUT a host round trip (steady state of a repeated evaluation: the caller allocated
 * `limit_slots` slots from the total of the previous evaluation of this shape).  info_dev[0] = total slots
 * (saturated to 2^32 - 1), info_dev[1] = max(thresh[0..nq)) -- read back with the results; status_dev |= 4 when
 * the total exceeds limit_slots (slices that would leave the arena get capacity 0: cap_dev is updated in place),
 * |= 8 when a threshold reaches key_limit (the narrowed key range the caller sized its arrays for; 0 = no check).
 * The caller re-runs the evaluation through ch_record_offsets when either bit comes back set. */
int ch_record_offsets_async(ch_ws* ws, uint32_t* cap_dev, int nstripes, int64_t nq, int64_t nq_pad,
                            uint32_t* off_dev, const uint32_t* thresh_dev /* or NULL */, uint64_t limit_slots,
                            uint32_t key_limit, uint32_t* info_dev /* u32[2] or NULL */, uint32_t* status_dev,
                            void* stream);
/* ch_scan_bases for the {all, relevant} totals of the candidate lists in one launch: tot (world, 2, nbins,
 * nq_pad) -> base0_all, base0_rel (NULL: no labels), key_max[q] = smallest key at which the global list holds
 * rmax items, total_rel[q] (optional), and the verification of a sampled threshold (torch.topk must see at least
 * R items): status |= 1 if some query holds fewer than `need` candidates (need <= 0: no check). */
int ch_scan_bases_pair(ch_ws* ws, const uint32_t* tot_dev, int world, int rank, int nbins, int64_t nq,
                       int64_t nq_pad, int64_t rmax, int64_t need, uint32_t* base0_all_dev,
                       uint32_t* base0_rel_dev, uint32_t* key_max_dev, uint32_t* total_rel_dev,
                       uint32_t* status_dev, uint32_t* bad_dev /* (nq_pad) or NULL: 1 for every short query */,
                       void* stream);
/* every stride-th row of a packed bit plane (the row sample that picks the thresholds): out (rows_out_pad, words),
 * rows beyond ceil(n_src / stride) are zeroed */
int ch_gather_rows(ch_ws* ws, const uint32_t* bits_dev, int64_t n_src, int nbit, int64_t stride,
                   uint32_t* out_dev, int64_t rows_out_pad, void* stream);
/* per-stripe class histogram of single-label gallery ids: cls (nstripes, nclass), zeroed by caller */
int ch_class_counts(ch_ws* ws, const uint32_t* g_ids, int64_t ndb, int rows_per_stripe, int nclass,
                    uint32_t* cls_dev, void* stream);

/* ---- K4: label match, precision@k, AP ---------------------------------------------------------
 * Replaces the per-query numpy loop of the missing calculate_mAP (97 % of its wall time):
 * `imatch -> cumsum -> Px -> sum(Px * imatch) / sum(imatch)`.  Works on the records of the relevant
 * retrieved items: rank = base0_all + stripe base + prefix, relrank likewise on the relevant counts.
 * cols (nq, ncols) fp64 per rank, ncols = 2 * nR + nPR: [sum_i, cnt_i]*nR, hits_k*nPR -- these are
 * SUMMED over ranks (NCCL all-reduce) before ch_reduce_means.
 */
typedef struct ch_final_args {
  const void*     recs;  const uint32_t* rec_off;  const uint32_t* rec_cnt;
  const uint32_t* base0_all;   const uint32_t* base0_rel;      /* (nbins, nq_pad) */
  const uint32_t* sbase_all;   const uint32_t* sbase_rel;      /* scanned slabs (nstripes, nbins, nq_pad) */
  const uint32_t* first_rel;   /* (nq_pad) 1 if the rank-0 item is relevant (remove_first only) */
  double*   partial;           /* (nstripes, nq_pad, ncols) scratch */
  double*   cols;              /* (nq, ncols) out */
  int64_t nq, nq_pad;
  int32_t nstripes, nbins, remove_first, nR, nPR;
  int64_t r_eff[CH_MAX_R];     /* list lengths, already clamped to the (post-removal) gallery size */
  int64_t pr_k[CH_MAX_PR];
} ch_final_args;
int ch_finalize_records(ch_ws* ws, const ch_final_args* a, void* stream);
/* remove_first_retrieved (test_hashing.py:105-112): first_rel[q] = 1 iff a relevant record has rank 0 */
int ch_first_relevant(ch_ws* ws, const ch_final_args* a, uint32_t* first_rel_dev, void* stream);
/* cols (summed over ranks) -> out_host[nR + 2 * nPR] = mAP_i..., recall_k..., precision_k...
 * total_rel (nq_pad) = relevant items in the whole gallery per query (u32, already global). */
int ch_reduce_means(ch_ws* ws, const double* cols_dev, const uint32_t* total_rel_dev,
                    const uint32_t* first_rel_dev, int64_t nq, int nR, int nPR, const int64_t* pr_k,
                    double* ap_out_dev /* (nR, nq) or NULL */, double* out_host,
                    const uint32_t* flags_dev /* u32[nflags] or NULL */, uint32_t* flags_host /* same sync */,
                    int nflags, void* stream);
/* the same in two halves: _enqueue launches the reduction and the copies into the workspace's pinned result words on
 * `stream` without waiting (capturable in a CUDA graph; at most 128 status words), _fetch waits for the stream and hands
 * the values out */
int ch_reduce_means_enqueue(ch_ws* ws, const double* cols_dev, const uint32_t* total_rel_dev,
                            const uint32_t* first_rel_dev, int64_t nq, int nR, int nPR, const int64_t* pr_k,
                            double* ap_out_dev, const uint32_t* flags_dev, int nflags, void* stream);
int ch_reduce_means_fetch(ch_ws* ws, double* out_host, int nout, uint32_t* flags_host, int nflags, void* stream);
/* flags_dev[0] |= 1 if some total_dev[q] < need  (verification of a sampled threshold) */
int ch_check_counts(ch_ws* ws, const uint32_t* total_dev, int64_t nq, int64_t need, uint32_t* flags_dev,
                    void* stream);

/* ranked id list from CH_EMIT_CANDIDATES records: ids (nq, R) int64 / keys (nq, R) int32, pre-filled
 * by the caller (-1); every rank writes only its own rows' slots. */
int ch_scatter_ranked(ch_ws* ws, const ch_final_args* a, int64_t R, int64_t row_offset,
                      int64_t* ids_dev, int32_t* keys_dev, void* stream);
/* AP / hits straight from a ranked id list (one warp per query): the literal
 * `imatch / cumsum / Px` loop.  labels as ids or masks.  cols layout as above with nR = 1. */
int ch_ap_from_ranked(ch_ws* ws, const int64_t* ids_dev, int64_t nq, int64_t R,
                      const uint32_t* q_lab, const uint32_t* g_lab, int label_mode, int mask_words,
                      int nPR, const int64_t* pr_k, double* cols_dev, void* stream);

/* ---- small / debug ----------------------------------------------------------------------------
 * dense key matrix (nq, ndb) u16 for parity tests and for get_hamm_dist (trainers/orthohash.py:362,
 * trainers/dpn.py:30; identity trainers/orthohash.py:263-264). */
int ch_hamming_matrix(ch_ws* ws, const uint32_t* q_bits, const uint32_t* q_nz, const uint32_t* g_bits,
                      const uint32_t* g_nz, int64_t nq, int64_t ndb, int nbit, int ternary,
                      uint16_t* out_dev, void* stream);
/* integer-pipe micro-benchmark: measured POPC.32 throughput of this GPU (ops / s) */
int ch_popc_peak(ch_ws* ws, double* popc32_per_s, double* elapsed_ms);
/* number of kernels this library launched since the workspace was created */
int64_t ch_launch_count(ch_ws* ws);

#ifdef __cplusplus
}
#endif
#endif /* CONCEPTHASH_B200_H */
