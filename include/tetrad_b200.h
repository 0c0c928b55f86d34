/*
 * tetrad_b200 -- C ABI of the B200-native quartet-resolution hot path.
 *
 * This is the drop-in boundary for ONE path of eaton-lab/tetrad: the per-quartet loop
 *   sample one SNP per locus -> 3 x (16x16) site-pattern counts -> SVDquartets score -> best split
 * i.e. reference `new_infer_resolved_quartets` (tetrad/src/resolve_quartets.py:191-265) with its
 * two numba kernels (`subsample_chunk_to_matrices` :42-73, `full_chunk_to_matrices` :76-104) and its
 * numpy.linalg svd / matrix_rank calls (:242-243).  The reference has no FFI of its own (it is pure
 * Python); INTEGRATION.md shows the ctypes stub a maintainer adds to route that function here.
 *
 * Conventions
 *   - plain C, no exceptions; every call returns 0 on success or a negative TQ_E* code and
 *     records a message retrievable with tq_last_error() (thread-local).
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensors); pointers
 *     named h_* are HOST pointers.  `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream).  Device-buffer calls only enqueue work: no hidden synchronisation.
 *   - codes in tmparr: 0,1,2,3 = A,C,G,T; anything else is missing (the reference writes 78 = 'N').
 *   - a quartet is four ascending tip indices; split k of the result means
 *       0: t0 t1 | t2 t3     1: t0 t2 | t1 t3     2: t0 t3 | t1 t2
 *     (reference tetrad/src/run_inference.py:41-45).
 */
#ifndef TETRAD_B200_H
#define TETRAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQ_VERSION 100          /* 0.1.0 */

#define TQ_OK            0
#define TQ_EINVAL       -1      /* bad argument (NULL pointer, negative size, tip index >= ntips) */
#define TQ_ELOCUS       -2      /* locus ids do not form one contiguous run per id */
#define TQ_ECUDA        -3      /* a CUDA runtime call failed; see tq_last_error() */
#define TQ_ENOMEM       -4

int         tq_version(void);
const char *tq_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Locus plan: the per-replicate locus structure shared by every quartet.
 *
 * Replaces the role of `tmpmap[:, 0]` inside subsample_chunk_to_matrices
 * (resolve_quartets.py:58-64): the `last_loc` state machine keeps the first usable column of
 * every run of equal locus id.  The plan sorts the locus runs by length (longest first), cuts
 * them into tiles of 1024 loci and records, per tile, how many within-locus ranks it has; the
 * packed matrix then stores "rank k of locus l" at bit l of plane k, so "first usable column
 * of a locus" becomes "lowest k with the bit set" -- a running OR, with no cross-word scan.
 *
 * h_locus: tmpmap[:,0] on the host, `stride` = element stride (2 for the reference's u32[S,2]).
 * Fails with TQ_ELOCUS when an id occurs in two separate runs (the reference's result would
 * then depend on which columns are missing in between; its own get_spans assumes sorted ids).
 * The plan uploads its small tables to `device` on `stream` and owns them.  device = -1 builds
 * a host-only plan (layout queries via tq_plan_info / tq_plan_tables; no kernels can use it).
 */
typedef struct tq_plan tq_plan;

int  tq_plan_create(const uint32_t *h_locus, int64_t ncols, int64_t stride,
                    int device, void *stream, tq_plan **out);
/* Same plan from the run lengths alone (run i covers the next h_run_len[i] columns): used for
 * bootstrap replicates, whose instance lengths are known before any column is materialised. */
int  tq_plan_create_runs(const uint32_t *h_run_len, int64_t nruns, int device, void *stream, tq_plan **out);
/* The device tables of a plan are stream-ordered memory (cudaMallocAsync) of the stream given at creation:
 * tq_plan_destroy frees them in that stream's order without synchronising the device, so work already enqueued
 * on that stream may still use the plan; work on OTHER streams must have completed before the call. */
void tq_plan_destroy(tq_plan *plan);

/* info[0]=ncols  [1]=nloci  [2]=ntiles  [3]=row_words (u32 words per tip row of the packed matrix)
 * info[4]=max locus length  [5]=padded columns (sum over tiles of ranks*1024)  [6]=device  [7]=0 */
int  tq_plan_info(const tq_plan *plan, int64_t info[8]);

/* Copies of the plan's tables (any pointer may be NULL): tile_off u32[ntiles+1] (offset of each
 * tile inside a tip row, in 16-byte units), tile_k u32[ntiles] (ranks per tile), loc_start /
 * loc_len u32[ntiles*1024] (first column and length of each locus in sorted order, 0-padded),
 * loc_run u32[ntiles*1024] (index of that locus run in column order, before the length sort). */
int  tq_plan_tables(const tq_plan *plan, uint32_t *tile_off, uint32_t *tile_k,
                    uint32_t *loc_start, uint32_t *loc_len, uint32_t *loc_run);

/* ------------------------------------------------------------------------------------------
 * K1: pack.  d_tmparr u8[ntips][row_stride] (device)  ->  d_planes u32[ntips][row_words] (device):
 * one-hot A/C/G/T bit-planes (missing = no bit set), interleaved as uint4 {A,C,G,T} per 32 loci.
 * Replaces the row gather + `seqs >= 78` pass of resolve_quartets.py:212-216 (done once per
 * replicate instead of once per quartet).  d_planes must be 16-byte aligned.
 */
int tq_pack(const tq_plan *plan, const uint8_t *d_tmparr, int64_t ntips, int64_t row_stride,
            uint32_t *d_planes, void *stream);
/* Only tip rows [tip_lo, tip_hi) of the same matrices (d_tmparr and d_planes are the bases of the full matrices):
 * under one process per GPU every rank packs its share of the rows and the planes are all-gathered. */
int tq_pack_rows(const tq_plan *plan, const uint8_t *d_tmparr, int64_t ntips, int64_t row_stride,
                 int64_t tip_lo, int64_t tip_hi, uint32_t *d_planes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Quartet unranking on the device.  Replaces _index_to_combination / iter_chunks_full
 * (tetrad/src/combinations.py:82-106): lexicographic rank in itertools.combinations(range(ntips), 4)
 * -> four ascending tip ids, written to d_quartets u32[m][4].  d_ranks i64[m] (device) gives the ranks
 * (e.g. the reference's rng.choice(C(n,4), size, replace=False) sample); with d_ranks = NULL the
 * ranks are the contiguous range [start, start + m) -- the exhaustive enumeration, no upload at all.
 * Ranks must be < C(ntips, 4) (checked for the range form only).
 */
int tq_unrank(int device, const int64_t *d_ranks, int64_t start, int64_t m, int64_t ntips,
              uint32_t *d_quartets, void *stream);

/* Random quartet sample without a host draw (scales with the number of GPUs; the reference's host draw,
 * rng.choice(C(n,4), Q, replace=False) at tetrad/src/combinations.py:109-114, is serial and costs 0.13 s per million
 * quartets): quartet i of the sample is unrank(pi_key(i)) with pi_key a keyed pseudo-random permutation of
 * [0, C(ntips,4)) (6-round Feistel network, cycle-walked).  Writes quartets start .. start + m - 1 of the sample to
 * d_quartets u32[m][4]; distinct indices give distinct quartets.  Same distribution as the reference's draw, another
 * stream (like the within-locus shuffles); tq_unrank with uploaded ranks reproduces the reference's. */
int tq_sample_quartets(int device, int64_t start, int64_t m, int64_t ntips, uint64_t key, uint32_t *d_quartets, void *stream);
/* the ranks pi_key(start .. start + m - 1) themselves, computed on the host (tests; no kernel) */
int tq_sample_ranks_host(int64_t start, int64_t m, int64_t ntips, uint64_t key, int64_t *h_ranks);

/* ------------------------------------------------------------------------------------------
 * K2/K5: bootstrap replicate packed on the device.  Replaces resample_tmp_database + jit_resample +
 * jit_resolve_ambigs (run_inference.py:99-143, tetrad/jit/resample.py:20-64,
 * tetrad/jit/resolve_ambigs.py:12-36) for replicates >= 1: `plan` comes from
 * tq_plan_create_runs(lengths of the drawn loci); d_src_start u32[nruns] is the first column, in
 * d_seqarr, of the locus drawn for instance i (spans[lidxs[i], 0]); d_seqarr u8[ntips][row_stride]
 * holds ASCII codes (A C G T, N / - = missing, R K S Y W M = two-base ambiguity) or codes 0..3.
 * Inside every instance the column order is a Philox4x32-10 permutation keyed by seed_shuffle; every
 * ambiguity cell takes one of its two bases by a Philox coin keyed by seed_ambig (the reference
 * uses numba's private generator for both, so only the distribution can match -- SURVEY.md H6).
 * d_srccol is caller-owned scratch of plan info[5] (padded columns) u32.
 */
int tq_boot_pack(const tq_plan *plan, const uint8_t *d_seqarr, int64_t ntips, int64_t row_stride,
                 const uint32_t *d_src_start, uint32_t seed_shuffle, uint32_t seed_ambig,
                 uint32_t *d_srccol, uint32_t *d_planes, void *stream);
/* The same replicate, planes of tip rows [tip_lo, tip_hi) only (the column orders in d_srccol are built in full by
 * every caller; coins are keyed by the global tip index, so the rows equal those of the unsharded call bit for bit). */
int tq_boot_pack_rows(const tq_plan *plan, const uint8_t *d_seqarr, int64_t ntips, int64_t row_stride,
                      int64_t tip_lo, int64_t tip_hi, const uint32_t *d_src_start, uint32_t seed_shuffle,
                      uint32_t seed_ambig, uint32_t *d_srccol, uint32_t *d_planes, void *stream);

/* ------------------------------------------------------------------------------------------
 * K3+K4: resolve.  For each of the m quartets in d_quartets u32[m][4]:
 *   counts (resolve_quartets.py:42-104), singular values + ranks (:242-243), minrank (:246),
 *   tail scores (:248), first argmin (:251), nsnps (:226,:264); zero-SNP quartets get scores
 *   0.001 (:232) and a split drawn from a counter-based hash of (seed, tip ids) in place of the
 *   reference's unseeded np.random.randint(3) (:231).
 * Outputs (device): d_rscor f64[m][3], d_rstat u32[m][2] = (best split, nsnps),
 *   d_counts u32[m][256] or NULL = mats[0] row-major (index 64a+16b+4c+d),
 *   d_minrank u32[m] or NULL.
 * subsample != 0 -> one SNP per locus (subsample_chunk_to_matrices; a large launch first converts the packed matrix into
 * {hi, lo, valid} words in stream-ordered scratch of the matrix's own size, freed behind the launch); 0 -> all SNPs
 * (full_chunk_to_matrices; chunks of the list whose quartets fill the tiles of a tip-pair grid densely -- exhaustive
 * lists and contiguous ranges of them -- go through one u8 GEMM over tip pairs on tcgen05.mma, decided on the device).  Tip indices must be < ntips (checked on the device: offending
 * quartets get best = 0xFFFFFFFF and word 0 of the call's status block, if given, is set non-zero).
 * d_status: NULL or u32[4], zeroed by the caller.  [0] bit 0 = some quartet had a tip index out of range; [1..3] are
 *   written only after tq_set_option("stack_stats", 1): diagnostics of the tensor-core kernel in subsample mode, summed
 *   over the quartets (wrapping): tile products taken, tiles stacked on earlier ones, and loci counted one by one (the
 *   busiest lane's share) -- see "stack_fill" below.
 */
int tq_resolve(const tq_plan *plan, const uint32_t *d_planes, int64_t ntips,
               const uint32_t *d_quartets, int64_t m, int subsample, uint64_t seed,
               double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts, uint32_t *d_minrank,
               uint32_t *d_status, void *stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer entry: the whole of new_infer_resolved_quartets(tmparr, tmpmap, quartets,
 * subsample_snps) for numpy callers.  Copies inputs to the device, plans, packs, resolves,
 * copies results back and synchronises.  `ctx` keeps device/pinned staging buffers between
 * calls (the reference re-reads tmparr per chunk, resolve_quartets.py:33-35; with a ctx the
 * packed matrix is reused when the same tmparr/tmpmap pointers+sizes+`epoch` come again).
 */
typedef struct tq_ctx tq_ctx;

int  tq_ctx_create(int device, tq_ctx **out);
void tq_ctx_destroy(tq_ctx *ctx);

int tq_resolve_host(tq_ctx *ctx,
                    const uint8_t *h_tmparr, int64_t ntips, int64_t ncols, int64_t row_stride,
                    const uint32_t *h_locus, int64_t locus_stride,
                    const uint32_t *h_quartets, int64_t m, int subsample, uint64_t seed,
                    uint64_t epoch,
                    double *h_rscor, uint32_t *h_rstat, uint32_t *h_counts, uint32_t *h_minrank);

/* bytes copied by the last tq_resolve_host call on this ctx: out[0]=host->device, out[1]=device->host */
int tq_ctx_last_traffic(const tq_ctx *ctx, int64_t out[2]);

/* ------------------------------------------------------------------------------------------
 * Result sink (host only, no CUDA): the text the reference writes with pandas / per-line Python.
 * tq_format_tsv: rows "t0 t1 t2 t3 s0 s1 s2 best nsnps" (tab separated, scores %.6f, '\n' ended) of
 *   {name}.quartets_{rep}.tsv (run_inference.py:232-234) into `out`; returns the number of bytes
 *   written, TQ_ENOMEM if `cap` (>= 160 bytes per row suffices) is too small.
 * tq_round_scores: every score as it reads back from that text (the reference derives its wQMC weights
 *   from the parsed TSV, run_inference.py:262-297).
 * tq_format_qmc: lines "a,b|c,d:%.5f" (run_inference.py:305) for n quartets whose tips are already in
 *   split order; returns bytes written (>= 128 bytes per line suffices).
 * tq_qmc_weights: the quartets that pass the filters (nsnps >= max(1, min_snps), ratio >= min_ratio;
 *   run_inference.py:258,275,300), their tips in split order a,b|c,d and their wQMC weight (strategy 0: 1,
 *   1: mean of the two largest scores, 2: that mean over the smallest, 3: 1 - smallest / sum; :280-297), computed
 *   from the scores as they read back from the TSV.  tips_out u32[m][4] and weight_out f64[m] receive the kept
 *   rows in order; returns their number, TQ_EINVAL for a score that is not a finite non-negative number.
 */
int64_t tq_format_tsv(const uint32_t *quartets, const uint32_t *rstat, const double *rscor, int64_t m,
                      char *out, int64_t cap);
int     tq_round_scores(const double *rscor, int64_t n, double *out);
int64_t tq_format_qmc(const uint32_t *tips, const double *weight, int64_t n, char *out, int64_t cap);
/* the same lines for rows rows[0..n) of tips / weight (the shuffled order the wQMC input is written in, :308-327),
 * without materialising the permuted arrays */
int64_t tq_format_qmc_rows(const uint32_t *tips, const double *weight, const int64_t *rows, int64_t n, char *out, int64_t cap);
int64_t tq_qmc_weights(const uint32_t *quartets, const uint32_t *rstat, const double *rscor, int64_t m, int weights,
                       int64_t min_snps, double min_ratio, uint32_t *tips_out, double *weight_out);

/* ------------------------------------------------------------------------------------------
 * Measurement helpers (bench.py): pipe micro-benchmarks on `device`.
 * kind 0: POPC  (result: popc32 thread-instructions per second)
 * kind 1: LOP3  (and/or)      kind 2: DFMA        kind 3: device copy (bytes/s, read+write)
 * kind 4: IMMA.16832.U8.U8 (mma.sync m16n8k32 u8), result: warp-level instructions per second
 * kind 10..14: POPC rate at 16 warps/SM when mixed with nothing / SHFL / LDS.128 / DFMA / ALU work
 */
int tq_microbench(int device, int kind, double *result);

/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t tq_launch_count(void);

/* name of the resolve kernel variant the last tq_resolve / tq_resolve_host launched
 * ("tq::resolve_mma_kernel<true>", ...); a static string, never NULL */
const char *tq_last_kernel(void);

/* count warps per CTA (of 16; the others score) of the last tensor-core resolve launch: the launcher picks the mix from
 * the plan's tiles and rank chunks per tip row (8 + 8 at the C2 shape, 12 + 4 at C3 / C4 / C5); 0 before the first launch */
int tq_last_count_warps(void);

/* Full mode routes every chunk (2^18 quartets) of a list on the device: to the tip-pair GEMM when its tiles are dense
 * (>= 30 quartets per 128-quartet tile, all tips ascending), to the per-quartet kernel otherwise.  out[0] / out[1] =
 * chunks routed to the GEMM / to the per-quartet kernel since the library was loaded (read after synchronising). */
int tq_pairgemm_stats(int64_t out[2]);

/* Tuning / test hooks; process-wide, initialised from the environment (TQ_KERNEL, TQ_MMA_FLUSH).
 *   "kernel"           0 = chosen by launch size and mode (default), 1 = one-warp-per-quartet POPC kernel,
 *                      3 = tensor-core (IMMA) count kernel, 5 = tip-pair GEMM on tcgen05.mma (full mode; a subsample
 *                      launch runs 3)
 *   "mma_flush_every"  > 0: tile products between accumulator flushes of the tensor-core kernel
 *                      (default 16384; tests lower it to reach the flush path), 0 = default
 *   "stack_fill"       subsample mode stacks the merged planes of consecutive sparse tiles in one tile copy and takes ONE
 *                      tile product when this percentage of its 1024 bit slots is filled: -1 = default (100: the
 *                      overflow rule "stack_ovf" alone decides), 0 = one product per tile (never stack), 1..99
 *   "stack_ovf"        a tile is stacked on the previous ones only if no lane has more than this many selected loci
 *                      whose bit slot is already taken (those are counted one by one, by shared-memory atomics);
 *                      otherwise the stack is multiplied out first.  0..32, default 8
 *   "stack_stats"      != 0: the tensor-core kernel adds its stacking diagnostics to words 1..3 of the status block
 *                      (default 0: off)
 *   "count_warps"      count warps per CTA of the tensor-core kernel (of 16; the others score): 0 = chosen from the
 *                      plan (default: 8 at the C2 shape, 12 where a tip row has a hundred tiles or more), 8, 12
 * Returns TQ_EINVAL for an unknown name or an out-of-range value. */
int tq_set_option(const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* TETRAD_B200_H */
