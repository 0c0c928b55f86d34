// Full mode (subsample_snps = False, the reference CLI's default: tetrad/src/cli_init.py:61) for dense quartet lists as
// ONE u8 GEMM over tip PAIRS on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in tensor memory).
//
// full_chunk_to_matrices (tetrad/src/resolve_quartets.py:76-104) counts EVERY column whose four cells are bases:
//
//     T_q[ab][cd] = sum over columns s of  X^{(t0,t1)}_{ab}[s] * Y^{(t2,t3)}_{cd}[s],     X_ab = A_a(t0) & B_b(t1),
//                                                                                          Y_cd = C_c(t2) & D_d(t3)
//
// (a missing cell has no plane bit, so incomplete columns drop out by themselves; the four invariant patterns aaaa are
// zeroed afterwards, :88-95 / :218).  Unlike subsample mode there is no per-locus selection, so the operands depend on
// a tip PAIR only: all quartets that share (t0, t1) share X, all that share (t2, t3) share Y, and the counts of every
// quartet are 16 x 16 blocks of the product of the pair-plane matrix with itself.  The per-quartet tensor-core kernel
// (tq_resolve.cu) unpacks both operands once per QUARTET (206 tile products of 16 x 16 x 1024 at the C2 shape); here
// they are unpacked once per PAIR and tile:
//
//   tile        8 row pairs (t0 < t1) x 16 column pairs (t2 < t3) = 128 x 256 outputs = 128 candidate quartets;
//               pairs are numbered lexicographically, so the tiles that hold wanted quartets (t1 < t2) form a staircase
//               -- a device planner buckets the quartet list by tile, only non-empty tiles are computed, and the planner
//               also DECIDES, on the device, whether the chunk is dense enough for the GEMM (>= 30 quartets per tile)
//               or goes to the per-quartet kernel: no host synchronisation, and a rank's contiguous share of an
//               exhaustive list stays on the GEMM however many ranks there are
//   K           all 32-column words of the packed rows; one tcgen05.mma (M = 128, N = 256, K = 32 bytes) per word
//   producers   12 warps, thread = one operand row (pair, ab): two 4-byte plane loads, one AND, and the 32 columns of
//               the word become 32 bytes {0, 1} -- (x >> v) & 0x01010101 for v = 0..7 -- written as two 16-byte K chunks
//               into the stage's K-major core-matrix layout (no swizzle: 8 rows x 16 B = one 128-byte core matrix)
//   MMA         one thread issues 8 MMAs per stage (kKW = 8 words), tcgen05.commit releases the stage / publishes the tile
//   epilogue    4 warps (one per tensor-memory lane quadrant) read the wanted 16 x 16 blocks with tcgen05.ld and write the
//               quartet's 256 counts (1 KB) to global memory; two accumulator buffers (2 x 256 TMEM columns) overlap the
//               epilogue of tile i with the main loop of tile i + 1
//   scores      score_counts_kernel (the same svd_scores4 as every other kernel) after the GEMM, 16 warps per SM.  (Running
//               it beside the GEMM on a second stream, one group of tiles behind, was measured first: 6 score warps fit
//               next to a GEMM CTA, and the pair took as long as one after the other -- 20.8 ms against 10.8 + 12.5 under
//               ncu -- because both live off the same issue slots; alone, each gets the whole SM.)
//
// Exactness: products are 0/1, accumulation is s32, K < 2^31 columns: counts are exact integers, bit-identical to the
// other kernels'.  Applies when the list holds >= 30 % of all C(n,4) quartets (exhaustive lists in practice; a sparse
// random sample would compute 128 candidates to use one or two -- the per-quartet kernel keeps those).
#include "tq_kernels.cuh"
#include <cstdlib>
#include "tq_device.cuh"
#include "tq_svd.cuh"

#include <cub/device/device_scan.cuh>

#include <mutex>

namespace tq {

static constexpr unsigned kFull = 0xFFFFFFFFu;
static constexpr int kRP = 8;                     // row pairs per tile    -> M = 128
static constexpr int kCP = 16;                    // column pairs per tile -> N = 256
static constexpr int kKW = 8;                     // plane words per stage -> 8 MMAs of K = 32 bytes (96 KB of operands:
static constexpr int kGStages = 2;                // the per-stage wait / fence / arrive is paid once per 256 columns)
static constexpr int kProdWarps = 12;             // 384 producer threads = 128 A rows + 256 B rows
static constexpr int kMmaWarp = kProdWarps;       // warp 12
                                                  // warps 13..16: epilogue (their index % 4 covers the four TMEM lane quadrants)
static constexpr int kGemmThreads = (kProdWarps + 5) * 32;
static constexpr uint32_t kABytes = kKW * 2 * 128 * 16;      // 32 KB: [word][K chunk][row][16 B]
static constexpr uint32_t kBBytes = kKW * 2 * 256 * 16;      // 64 KB
static constexpr uint32_t kStageBytes = kABytes + kBBytes;
static constexpr uint32_t kTmemColsGemm = 512;               // two accumulators of 256 columns
static constexpr int kScoreWarps = 16;                       // 512 threads x 128 registers = the SM's register file

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 at bit 4, a/b_format U8 = 0, K-major both,
// n_dim = N >> 3 at bit 17, m_dim = M >> 4 at bit 24
static constexpr uint32_t kIdesc = (2u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, no swizzle: start address >> 4, leading byte
// offset (between the two K chunks of a K = 32 step) >> 4 at bit 16, stride byte offset (between 8-row groups) >> 4 at
// bit 32, version 1 at bit 46, layout type 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

struct PairGemmArgs {
    const uint32_t *planes;        // u32 view of the packed planes: word w of tip t, base x at ((t * row_u4 + w) * 4 + x)
    uint32_t row_u4;
    uint32_t ntips;
    uint32_t npairs;               // n (n - 1) / 2
    uint32_t ntc;                  // column tiles per tile row
    const uint32_t *tile_list;     // non-empty cells, ascending
    const uint32_t *ntiles;        // number of non-empty cells (device: the planner's result)
    const uint32_t *cell_off;      // first entry of every cell
    const uint32_t *cell_count;
    const uint2 *entries;          // (quartet index, local row pair << 8 | local column pair), bucketed by cell
    uint32_t *counts;              // [m][256]
    const uint32_t *flags;         // flags[0] != 0: this chunk takes the GEMM route (else the kernels return at once)
    const uint32_t *rowmode;       // rowmode[0]: numbering of the row pairs the planner chose for this chunk (kRowsLex / kRowsT1)
};

// Pair numbering.  Column pairs (t2 < t3): lexicographic, index = off(t2) + t3 - t2 - 1.  Row pairs (t0 < t1): the planner
// counts the non-empty tiles under TWO numberings and takes the one with fewer, per chunk, on the device:
//   kRowsLex  lexicographic like the columns.  A contiguous range of a lexicographic quartet list (one rank's share, one
//             chunk of a long list) is a contiguous range of row pairs, so its tiles stay dense; but a row block holds eight
//             different t1, and every t2 among them gives half-empty tiles (C2, whole list: 2 879 tiles, 62 % wanted).
//   kRowsT1   by (t1, t0): a row block shares its t1, so a tile is wanted (t1 < t2) as a whole (C2, whole list: 1 958
//             tiles, 92 % wanted -- 16.7 instead of 20.9 ms); thin ranges of the list scatter over all row blocks (an
//             eighth of the C2 list: 987 tiles against 300).
static constexpr uint32_t kRowsLex = 0u, kRowsT1 = 1u;
__host__ __device__ __forceinline__ uint32_t pair_offset(uint32_t u, uint32_t n) { return u * (2 * n - 1 - u) / 2; }
__host__ __device__ __forceinline__ uint32_t pair_index(uint32_t u, uint32_t v, uint32_t n) { return pair_offset(u, n) + v - u - 1; }

__device__ __forceinline__ uint2 unrank_pair(uint32_t idx, uint32_t n)
{
    uint32_t lo = 0, hi = n - 2;                     // largest u with off(u) <= idx
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (pair_offset(mid, n) <= idx) lo = mid; else hi = mid - 1;
    }
    return make_uint2(lo, lo + 1 + (idx - pair_offset(lo, n)));
}

// ------------------------------------------------------------------------------------------
// planner: bucket the quartets by tile
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 unrank_rowpair_t1(uint32_t rp)
{
    uint32_t t1 = (uint32_t)((1.0 + sqrt(1.0 + 8.0 * (double)rp)) * 0.5);
    while (t1 * (t1 - 1) / 2 > rp) --t1;
    while (t1 * (t1 + 1) / 2 <= rp) ++t1;
    return make_uint2(rp - t1 * (t1 - 1) / 2, t1);
}
__device__ __forceinline__ uint32_t rowpair_index(uint32_t t0, uint32_t t1, uint32_t n, uint32_t rowmode)
{
    return rowmode == kRowsT1 ? t1 * (t1 - 1) / 2 + t0 : pair_index(t0, t1, n);
}
__device__ __forceinline__ bool quartet_cell(uint4 t, uint32_t n, uint32_t ntc, uint32_t rowmode, uint32_t &cell, uint32_t &local)
{
    if (!(t.x < t.y && t.y < t.z && t.z < t.w && t.w < n)) return false;      // out of range, or not ascending
    const uint32_t rp = rowpair_index(t.x, t.y, n, rowmode), cp = pair_index(t.z, t.w, n);
    cell = (rp / kRP) * ntc + cp / kCP;
    local = ((rp % kRP) << 8) | (cp % kCP);
    return true;
}

// stats[0]: quartets with a tip out of range; stats[1]: quartets in range whose tips are not ascending (the GEMM needs
// ascending tips; a chunk that holds one goes to the per-quartet kernel, which does not care)
__global__ void __launch_bounds__(256) pg_count_kernel(const uint4 *__restrict__ quartets, int64_t m, uint32_t n, uint32_t ntc,
                                                       uint32_t *__restrict__ cell_count, uint32_t *__restrict__ cell_count_t1,
                                                       uint32_t *__restrict__ stats)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const uint4 t = __ldg(quartets + q);
    uint32_t cell, local;
    if (quartet_cell(t, n, ntc, kRowsLex, cell, local)) {
        atomicAdd(cell_count + cell, 1u);
        quartet_cell(t, n, ntc, kRowsT1, cell, local);
        atomicAdd(cell_count_t1 + cell, 1u);
    } else if (t.x >= n || t.y >= n || t.z >= n || t.w >= n) atomicAdd(stats, 1u);
    else atomicAdd(stats + 1, 1u);
}

// nonempty[0 / 1] = non-empty cells (= tiles to compute) under the lexicographic / the (t1, t0) row numbering
__global__ void __launch_bounds__(256) pg_nonempty_kernel(const uint32_t *__restrict__ cell_count, const uint32_t *__restrict__ cell_count_t1,
                                                          uint32_t ncells, uint32_t *__restrict__ nonempty)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = __syncthreads_count(i < ncells && cell_count[i] != 0u);
    const int b = __syncthreads_count(i < ncells && cell_count_t1[i] != 0u);
    if (threadIdx.x == 0) {
        if (a) atomicAdd(nonempty, (uint32_t)a);
        if (b) atomicAdd(nonempty + 1, (uint32_t)b);
    }
}

// The numbering with fewer tiles wins (ties: lexicographic); its counts become THE cell counts of the chunk.
__global__ void __launch_bounds__(256) pg_pick_kernel(uint32_t *__restrict__ cell_count, const uint32_t *__restrict__ cell_count_t1,
                                                      uint32_t ncells, const uint32_t *__restrict__ nonempty, uint32_t *__restrict__ rowmode)
{
    const uint32_t mode = nonempty[1] < nonempty[0] ? kRowsT1 : kRowsLex;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) rowmode[0] = mode;
    if (mode == kRowsT1 && i < ncells) cell_count[i] = cell_count_t1[i];
}

// The chunk's route, decided where the planner's result lives.  flags[0] = 1: tip-pair GEMM + score kernel; 0: the
// per-quartet tensor-core kernel (which then skips nothing).  Every tile costs the same and about as much as 30 quartets
// of the per-quartet kernel (C2: 4.95 us per tile against 0.198 - 0.030 us per quartet; both scale with the columns).
// g_route counts the decisions (host-mapped: tq_pairgemm_stats).
__global__ void pg_decide_kernel(const uint32_t *__restrict__ ntiles, const uint32_t *__restrict__ nentries,
                                 const uint32_t *__restrict__ stats, int force, uint32_t *__restrict__ flags,
                                 unsigned long long *__restrict__ route)
{
    const bool dense = (uint64_t)ntiles[0] * 30ull <= (uint64_t)nentries[0];
    const bool gemm = stats[1] == 0u && nentries[0] > 0u && (force == 5 || dense);
    flags[0] = gemm ? 1u : 0u;
    atomicAdd(route + (gemm ? 0 : 1), 1ull);
}

// GEMM route only: quartets with a tip out of range never reach a tile; they get the error marks here
__global__ void __launch_bounds__(256) pg_mark_bad_kernel(const uint4 *__restrict__ quartets, int64_t m, uint32_t n,
                                                          const uint32_t *__restrict__ flags, double *__restrict__ rscor,
                                                          uint2 *__restrict__ rstat, uint32_t *__restrict__ status)
{
    if (!flags[0]) return;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const uint4 t = __ldg(quartets + q);
    if (t.x >= n || t.y >= n || t.z >= n || t.w >= n) {
        rstat[q] = make_uint2(0xFFFFFFFFu, 0u);
        rscor[3 * q] = rscor[3 * q + 1] = rscor[3 * q + 2] = 0.0;
        if (status) atomicOr(status, 1u);
    }
}

__global__ void __launch_bounds__(256) pg_flag_kernel(const uint32_t *__restrict__ cell_count, uint32_t ncells, uint32_t *__restrict__ flag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncells) flag[i] = cell_count[i] != 0u;
}

// tile_list[flag_scan[cell]] = cell for every non-empty cell (ascending); ntiles[0] = their number
__global__ void __launch_bounds__(256) pg_compact_kernel(const uint32_t *__restrict__ cell_count, const uint32_t *__restrict__ flag_scan,
                                                         uint32_t ncells, uint32_t *__restrict__ tile_list, uint32_t *__restrict__ ntiles)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncells && cell_count[i] != 0u) tile_list[flag_scan[i]] = i;
    if (i == 0) ntiles[0] = flag_scan[ncells];
}

__global__ void __launch_bounds__(256) pg_fill_kernel(const uint4 *__restrict__ quartets, int64_t m, uint32_t n, uint32_t ntc,
                                                      const uint32_t *__restrict__ rowmode, const uint32_t *__restrict__ cell_off,
                                                      uint32_t *__restrict__ cell_fill, uint2 *__restrict__ entries)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    uint32_t cell, local;
    if (!quartet_cell(__ldg(quartets + q), n, ntc, rowmode[0], cell, local)) return;
    const uint32_t pos = cell_off[cell] + atomicAdd(cell_fill + cell, 1u);
    entries[pos] = make_uint2((uint32_t)q, local);
}

// ------------------------------------------------------------------------------------------
// the GEMM
// ------------------------------------------------------------------------------------------
struct __align__(16) PairGemmShared {
    unsigned long long full[kGStages], empty[kGStages];     // stage filled by the producers / released by the MMAs
    unsigned long long tfull[2], tempty[2];                 // accumulator published / drained
    uint32_t tmem_base;
};

// One producer thread = one operand row (pair, ab) of kRows (128: A, 256: B).  Per stage: kKW plane words -> kKW x 32 bytes
// {0, 1}, two 16-byte K chunks per word at compile-time offsets from the row's slot in the stage.  Plane words are loaded
// TWO stages ahead and right AFTER a stage's fence: fence.proxy.async is a full memory barrier (SASS MEMBAR.ALL.CTA), so
// loads still in flight when it executes are waited for -- issued after it they have a whole stage of work to land.
template <uint32_t kRows>
__device__ __forceinline__ void produce_tile(const uint32_t *__restrict__ p0, const uint32_t *__restrict__ p1, uint32_t nst,
                                             unsigned char *row_base, unsigned long long *full, unsigned long long *empty,
                                             uint32_t &it, int lane)
{
    uint32_t x[kKW], na[kKW], nb[kKW];               // na & nb: the two planes of the next stage, ANDed when used
#pragma unroll
    for (int w = 0; w < kKW; ++w) {
        x[w] = __ldg(p0 + 4 * w) & __ldg(p1 + 4 * w);
        na[w] = nst > 1 ? __ldg(p0 + 4 * (kKW + w)) : 0u;
        nb[w] = nst > 1 ? __ldg(p1 + 4 * (kKW + w)) : 0u;
    }
    for (uint32_t st = 0; st < nst; ++st, ++it) {
        const uint32_t s = it % kGStages;
        mbar_wait_short(&empty[s], ((it / kGStages) & 1u) ^ 1u);      // the MMAs that read this slot are done
        unsigned char *base = row_base + s * kStageBytes;
#pragma unroll
        for (int w = 0; w < kKW; ++w) {
            uint32_t o[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) o[v] = (x[w] >> v) & 0x01010101u;   // byte j of o[v] = column 8 j + v of the word
            *reinterpret_cast<uint4 *>(base + w * kRows * 32) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4 *>(base + w * kRows * 32 + kRows * 16) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        fence_proxy_async_smem();                    // my stores are visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
#pragma unroll
        for (int w = 0; w < kKW; ++w) x[w] = na[w] & nb[w];
        if (st + 2 < nst) {
            const uint32_t *q0 = p0 + (size_t)(st + 2) * kKW * 4, *q1 = p1 + (size_t)(st + 2) * kKW * 4;
#pragma unroll
            for (int w = 0; w < kKW; ++w) {
                na[w] = __ldg(q0 + 4 * w);
                nb[w] = __ldg(q1 + 4 * w);
            }
        }
    }
}

__global__ void __launch_bounds__(kGemmThreads, 1)
pair_gemm_kernel(PairGemmArgs args)
{
    if (!args.flags[0]) return;                           // the planner sent this chunk to the per-quartet kernel
    extern __shared__ __align__(1024) unsigned char pg_smem[];
    unsigned char *stages = pg_smem;
    PairGemmShared &sh = *reinterpret_cast<PairGemmShared *>(pg_smem + kGStages * kStageBytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGStages; ++s) { mbar_init(&sh.full[s], kProdWarps); mbar_init(&sh.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sh.tfull[b], 1); mbar_init(&sh.tempty[b], 4); }
        mbar_init_fence();
    }
    if (warp == kMmaWarp) tmem_alloc(&sh.tmem_base, kTmemColsGemm);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = sh.tmem_base;

    const uint32_t tile_lo = 0u, tile_hi = args.ntiles[0], rowmode = args.rowmode[0];
    const uint32_t nst = args.row_u4 / kKW;                  // stages per tile (row_u4 is a multiple of 32)
    const uint32_t stage0 = smem_u32(stages);

    if (warp < kProdWarps) {
        // ------------------------------ producers: bits -> bytes ------------------------------
        const uint32_t row = threadIdx.x;                    // 0..127: A rows, 128..383: B rows
        const bool is_a = row < 128u;
        const uint32_t r = is_a ? row : row - 128u;          // row inside its operand
        const uint32_t dst0 = (is_a ? 0u : kABytes) + r * 16u;          // + word * nrows * 32 + chunk * nrows * 16
        uint32_t it = 0;                                      // stages produced so far (ring position)
        for (uint32_t tile = tile_lo + blockIdx.x; tile < tile_hi; tile += gridDim.x) {
            const uint32_t cell = __ldg(args.tile_list + tile);
            const uint32_t tr = cell / args.ntc, tc = cell % args.ntc;
            uint2 pr;                                        // my pair: (tip of the first plane, tip of the second plane)
            if (is_a) {
                const uint32_t rp = tr * kRP + (r >> 4), rq = rp < args.npairs ? rp : args.npairs - 1;
                pr = rowmode == kRowsT1 ? unrank_rowpair_t1(rq) : unrank_pair(rq, args.ntips);
            }
            else { const uint32_t cp = tc * kCP + (r >> 4); pr = unrank_pair(cp < args.npairs ? cp : args.npairs - 1, args.ntips); }
            const uint32_t *p0 = args.planes + (size_t)pr.x * args.row_u4 * 4 + ((r >> 2) & 3u);
            const uint32_t *p1 = args.planes + (size_t)pr.y * args.row_u4 * 4 + (r & 3u);
            if (is_a) produce_tile<128u>(p0, p1, nst, stages + dst0, sh.full, sh.empty, it, lane);
            else produce_tile<256u>(p0, p1, nst, stages + dst0, sh.full, sh.empty, it, lane);
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------ MMA issuer ------------------------------
        // K-major, no swizzle: the two 16-byte K chunks of a K = 32 step are rows x 16 B apart (leading byte offset), the
        // 8-row core matrices 128 B apart (stride byte offset)
        const uint32_t a_lbo = 128u * 16u, b_lbo = 256u * 16u, a_sbo = 128u, b_sbo = 128u;
        uint32_t it = 0, tcount = 0;
        for (uint32_t tile = tile_lo + blockIdx.x; tile < tile_hi; tile += gridDim.x, ++tcount) {
            const uint32_t buf = tcount & 1u;
            mbar_wait(&sh.tempty[buf], ((tcount >> 1) & 1u) ^ 1u);              // the epilogue has drained this accumulator
            tcgen05_fence_after();
            for (uint32_t st = 0; st < nst; ++st, ++it) {
                const uint32_t s = it % kGStages;
                mbar_wait_short(&sh.full[s], (it / kGStages) & 1u);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t sa = stage0 + s * kStageBytes, sb = sa + kABytes;
#pragma unroll
                    for (int w = 0; w < kKW; ++w)
                        umma_i8(tmem_base + buf * 256u, make_desc(sa + (uint32_t)w * 128u * 32u, a_lbo, a_sbo),
                                make_desc(sb + (uint32_t)w * 256u * 32u, b_lbo, b_sbo), kIdesc, (st | (uint32_t)w) != 0u);
                    umma_commit(&sh.empty[s]);                // slot free once these MMAs have read it
                    if (st + 1 == nst) umma_commit(&sh.tfull[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------ epilogue: TMEM -> counts ------------------------------
        const uint32_t quad = (uint32_t)warp & 3u;
        uint32_t tcount = 0;
        for (uint32_t tile = tile_lo + blockIdx.x; tile < tile_hi; tile += gridDim.x, ++tcount) {
            const uint32_t buf = tcount & 1u;
            const uint32_t cell = __ldg(args.tile_list + tile);
            const uint32_t e_lo = __ldg(args.cell_off + cell), e_hi = e_lo + __ldg(args.cell_count + cell);
            mbar_wait(&sh.tfull[buf], (tcount >> 1) & 1u);
            tcgen05_fence_after();
            for (uint32_t e = e_lo; e < e_hi; ++e) {
                const uint2 ent = __ldg(args.entries + e);
                const uint32_t lr = ent.y >> 8, lc = ent.y & 0xFFu;
                if ((lr >> 1) != quad) continue;              // rows 16 lr .. 16 lr + 15 live in lane quadrant lr / 2
                uint32_t v[16];
                tmem_ld16(tmem_base + ((quad * 32u) << 16) + buf * 256u + lc * 16u, v);
                tmem_wait_ld();
                if ((uint32_t)(lane >> 4) == (lr & 1u)) {
                    const uint32_t ab = (uint32_t)lane & 15u;
                    uint4 *dst = reinterpret_cast<uint4 *>(args.counts + (size_t)ent.x * 256 + ab * 16);
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if ((uint32_t)k == ab && ab % 5u == 0u) v[k] = 0u;       // invariant patterns aaaa (ab = cd = 5 a), :218
                    dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
                    dst[2] = make_uint4(v[8], v[9], v[10], v[11]);
                    dst[3] = make_uint4(v[12], v[13], v[14], v[15]);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.tempty[buf]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemColsGemm);
}

// ------------------------------------------------------------------------------------------
// scores from finished count tensors (global memory), the quartets named by a range of planner entries
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScoreWarps * 32, 1)
score_counts_kernel(const uint32_t *__restrict__ counts, const uint4 *__restrict__ quartets, const uint2 *__restrict__ entries,
                    const uint32_t *__restrict__ nentries, const uint32_t *__restrict__ flags, uint64_t seed,
                    double *__restrict__ rscor, uint2 *__restrict__ rstat, uint32_t *__restrict__ minrank_out)
{
    if (!flags[0]) return;
    extern __shared__ __align__(1024) unsigned char pg_smem[];
    SvdScratch4 *scratch = reinterpret_cast<SvdScratch4 *>(pg_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SvdScratch4 &sv = scratch[warp];
    const uint32_t e_lo = 0u, e_hi = nentries[0];          // entries = the quartets with valid tips, in tile order
    const uint32_t nwarps = gridDim.x * kScoreWarps;
    for (uint32_t base = e_lo + 4u * (blockIdx.x * kScoreWarps + warp); base < e_hi; base += 4u * nwarps) {
        const int nq = (int)min(4u, e_hi - base);
        uint32_t actmask = 0u, my_nsnps = 0u, my_best = 0u;
        long long my_q = -1;
        uint32_t qx[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= nq) continue;
            const uint32_t q = __ldg(entries + base + j).x;
            const uint32_t *T = counts + (size_t)q * 256;
            qx[j] = q;
            uint32_t nsnps = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) nsnps += T[lane + 32 * i];
            nsnps = __reduce_add_sync(kFull, nsnps);                        // :226
            uint32_t best = 0;
            if (nsnps == 0) best = tiebreak_split(seed, __ldg(quartets + q));   // stands in for :231
            else actmask |= 1u << j;
            if (lane == j) { my_q = q; my_nsnps = nsnps; my_best = best; }
        }
        double score[3] = {0.001, 0.001, 0.001};                            // :232
        uint32_t minrank = 0;
        if (actmask) {
            const uint32_t q0 = qx[0], q1 = qx[1], q2 = qx[2], q3 = qx[3];
            svd_scores4_impl([=](int qi) { return counts + (size_t)(qi == 0 ? q0 : qi == 1 ? q1 : qi == 2 ? q2 : q3) * 256; },
                             actmask, sv, lane, score, my_best, minrank);
        }
        if (lane < 4 && my_q >= 0) {
            rscor[3 * my_q] = score[0]; rscor[3 * my_q + 1] = score[1]; rscor[3 * my_q + 2] = score[2];
            rstat[my_q] = make_uint2(my_best, my_nsnps);                    // :251, :264
            if (minrank_out) minrank_out[my_q] = minrank;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// Whether a full-mode launch is worth planning at all (the decision proper is taken per chunk on the device): enough
// quartets for a persistent kernel, a pair x pair tile grid that fits a few million cells.
bool pairgemm_applies(int64_t ntips, int64_t m, int num_sms)
{
    if (ntips < 5 || ntips > 2048 || m < (int64_t)num_sms * 16) return false;
    const uint64_t np = (uint64_t)ntips * (ntips - 1) / 2;
    const uint64_t ncells = ((np + kRP - 1) / kRP) * ((np + kCP - 1) / kCP);
    return ncells <= (4u << 20);
}

namespace {
std::mutex g_attr_mu;
uint64_t g_attr_done = 0;
// chunks routed to the GEMM / to the per-quartet kernel since load: host-mapped pinned memory the decide kernel bumps
unsigned long long *g_route_host = nullptr;
std::mutex g_route_mu;
unsigned long long *route_counters()
{
    std::lock_guard<std::mutex> lk(g_route_mu);
    if (!g_route_host) {
        if (cudaHostAlloc(reinterpret_cast<void **>(&g_route_host), 2 * sizeof(unsigned long long),
                          cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        g_route_host[0] = g_route_host[1] = 0ull;
    }
    return g_route_host;
}
}  // namespace

void pairgemm_stats(int64_t out[2])
{
    unsigned long long *r = route_counters();
    out[0] = r ? (int64_t)r[0] : 0;
    out[1] = r ? (int64_t)r[1] : 0;
}

#define PG_CHECK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cudaFreeAsync(scratch, stream); return e__; } } while (0)

cudaError_t launch_resolve_pairgemm(const PlanDev &plan, const uint32_t *d_planes, int64_t ntips, const uint32_t *d_quartets,
                                    int64_t m, uint64_t seed, double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts,
                                    uint32_t *d_minrank, uint32_t *d_status, int num_sms, int force, cudaStream_t stream)
{
    if (m == 0) return cudaSuccess;
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t gemm_smem = (size_t)kGStages * kStageBytes + sizeof(PairGemmShared);
    const size_t score_smem = sizeof(SvdScratch4) * kScoreWarps;
    {
        std::lock_guard<std::mutex> lk(g_attr_mu);
        const uint64_t bit = 1ull << (dev & 63);
        if (!(g_attr_done & bit)) {
            cudaError_t e = cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(score_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)score_smem);
            if (e != cudaSuccess) return e;
            g_attr_done |= bit;
        }
    }
    unsigned long long *route_host = route_counters(), *route_dev = nullptr;
    if (!route_host || cudaHostGetDevicePointer(reinterpret_cast<void **>(&route_dev), route_host, 0) != cudaSuccess)
        return cudaErrorMemoryAllocation;

    const uint32_t n = (uint32_t)ntips, np = n * (n - 1) / 2;
    const uint32_t ntr = (np + kRP - 1) / kRP, ntc = (np + kCP - 1) / kCP, ncells = ntr * ntc;
    // Chunks of the list: 1 KB of counts per quartet lives in scratch (unless the caller wants the counts anyway), and
    // every chunk is routed on its own.
    const int64_t chunk_cap = (int64_t)1 << 18;
    const uint4 *quartets = reinterpret_cast<const uint4 *>(d_quartets);
    uint2 *rstat = reinterpret_cast<uint2 *>(d_rstat);
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)ncells + 1, stream);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const int64_t mcap = m < chunk_cap ? m : chunk_cap;
    // scratch: cell_count | cell_count_t1 | cell_off (+1) | flag | flag_scan (+1) | cell_fill | tile_list | small words | entries | counts | cub temp
    const size_t cb = up(((size_t)ncells + 1) * 4);
    const size_t o_count = 0, o_count2 = o_count + cb, o_off = o_count2 + cb, o_flag = o_off + cb, o_fscan = o_flag + cb,
                 o_fill = o_fscan + cb, o_list = o_fill + cb, o_small = o_list + cb, o_entries = o_small + 256, o_counts = o_entries + up((size_t)mcap * 8),
                 o_temp = o_counts + (d_counts ? 0 : up((size_t)mcap * 1024)), total = o_temp + up(temp_bytes);
    char *scratch = nullptr;
    cudaError_t e = cudaMallocAsync(&scratch, total, stream);
    if (e != cudaSuccess) return e;
    uint32_t *cell_count = (uint32_t *)(scratch + o_count), *cell_count_t1 = (uint32_t *)(scratch + o_count2);
    uint32_t *cell_off = (uint32_t *)(scratch + o_off);
    uint32_t *flag = (uint32_t *)(scratch + o_flag), *fscan = (uint32_t *)(scratch + o_fscan);
    uint32_t *cell_fill = (uint32_t *)(scratch + o_fill), *tile_list = (uint32_t *)(scratch + o_list);
    uint32_t *small = (uint32_t *)(scratch + o_small);          // [0] tiles, [1..2] stats, [4] route flag, [5] row numbering, [6..7] tiles under either
    uint32_t *ntiles = small, *stats = small + 1, *flags = small + 4, *rowmode = small + 5, *nonempty = small + 6;
    uint2 *entries = (uint2 *)(scratch + o_entries);

    for (int64_t c0 = 0; c0 < m; c0 += chunk_cap) {
        const int64_t mc = (m - c0 < chunk_cap) ? m - c0 : chunk_cap;
        uint32_t *counts = d_counts ? d_counts + (size_t)c0 * 256 : (uint32_t *)(scratch + o_counts);
        PG_CHECK(cudaMemsetAsync(scratch, 0, o_entries, stream));         // counts, offsets, flags, scans, fills, small words
        const unsigned qb = (unsigned)((mc + 255) / 256), cbk = (ncells + 1 + 255) / 256;
        pg_count_kernel<<<qb, 256, 0, stream>>>(quartets + c0, mc, n, ntc, cell_count, cell_count_t1, stats);
        pg_nonempty_kernel<<<cbk, 256, 0, stream>>>(cell_count, cell_count_t1, ncells, nonempty);
        pg_pick_kernel<<<cbk, 256, 0, stream>>>(cell_count, cell_count_t1, ncells, nonempty, rowmode);
        PG_CHECK(cub::DeviceScan::ExclusiveSum(scratch + o_temp, temp_bytes, cell_count, cell_off, (int)ncells + 1, stream));
        pg_flag_kernel<<<cbk, 256, 0, stream>>>(cell_count, ncells, flag);
        PG_CHECK(cub::DeviceScan::ExclusiveSum(scratch + o_temp, temp_bytes, flag, fscan, (int)ncells + 1, stream));
        pg_compact_kernel<<<cbk, 256, 0, stream>>>(cell_count, fscan, ncells, tile_list, ntiles);
        // cell_off[ncells] = number of bucketed quartets (the scan runs over ncells + 1 elements)
        pg_decide_kernel<<<1, 1, 0, stream>>>(ntiles, cell_off + ncells, stats, force, flags, route_dev);
        pg_fill_kernel<<<qb, 256, 0, stream>>>(quartets + c0, mc, n, ntc, rowmode, cell_off, cell_fill, entries);
        pg_mark_bad_kernel<<<qb, 256, 0, stream>>>(quartets + c0, mc, n, flags, d_rscor + 3 * c0, rstat + c0, d_status);
        g_launches += 11;
        PG_CHECK(cudaGetLastError());

        PairGemmArgs args;
        args.planes = d_planes; args.row_u4 = plan.row_u4; args.ntips = n; args.npairs = np; args.ntc = ntc;
        args.tile_list = tile_list; args.ntiles = ntiles; args.cell_off = cell_off; args.cell_count = cell_count;
        args.entries = entries; args.counts = counts; args.flags = flags;
        args.rowmode = rowmode;
        pair_gemm_kernel<<<num_sms, kGemmThreads, gemm_smem, stream>>>(args);
        score_counts_kernel<<<num_sms, kScoreWarps * 32, score_smem, stream>>>(counts, quartets + c0, entries, cell_off + ncells, flags,
                                                                             seed, d_rscor + 3 * c0, rstat + c0,
                                                                             d_minrank ? d_minrank + c0 : nullptr);
        g_launches += 2;
        PG_CHECK(cudaGetLastError());
        // the other route: the per-quartet kernel returns at once when flags[0] says the GEMM took the chunk
        PG_CHECK(launch_resolve_mma_full(plan, d_planes, ntips, d_quartets + 4 * c0, mc, seed, d_rscor + 3 * c0, d_rstat + 2 * c0,
                                         d_counts ? d_counts + (size_t)c0 * 256 : nullptr, d_minrank ? d_minrank + c0 : nullptr,
                                         d_status, num_sms, flags, stream));
    }
    cudaFreeAsync(scratch, stream);
    g_last_kernel = "tq::pair_gemm_kernel (tcgen05.mma kind::i8) + tq::score_counts_kernel | tq::resolve_mma_kernel<false>, routed per chunk on the device";
    return cudaSuccess;
}

}  // namespace tq
