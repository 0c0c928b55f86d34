// Hand-written sm_100a kernels of the quartet-resolution hot path.
//
//   pack_kernel            K1     u8 tmparr -> locus-tiled one-hot A/C/G/T bit-planes
//   boot_columns_kernel,   K2/K5  bootstrap replicates on the device: Philox column orders per locus instance, planes packed
//   pack_gather_kernel            straight from the resident ASCII seqarr with IUPAC cells resolved by Philox coins
//   unrank_kernel                 lexicographic rank -> quartet (combinations.py:94-106)
//   resolve_mma_kernel     K3+K4  the kernel large launches run: count warps select one SNP per locus (stage 1) and form the
//                                 256 site-pattern counts on the integer tensor cores (stage 2: IMMA.16832.U8.U8 on bit planes
//                                 unpacked by one LOP3 per four loci); score warps take four quartets at a time through three
//                                 FP64 QR + one-sided-Jacobi passes of four 16x16 flattenings; rank / minrank / tail score /
//                                 argmin; only the result leaves the SM
//   resolve_kernel                same result, one warp per quartet, counts by AND + POPC with the counters in tensor memory
//                                 (tcgen05.alloc/ld/st): what small launches run
//   resolve_ws_kernel             warp-specialised POPC variant (round-1 v5), selectable
//   pipe_kernel, mix_kernel, imma_rate_kernel, copy_kernel   roofline denominators measured on the box (tq_microbench)
//
// launch_resolve picks the resolve kernel by launch size; tq_set_option("kernel", ..) / TQ_KERNEL force one.  All three
// give bit-identical results.  Build switches (all default to the shipped configuration; profiles/ cites the measurements):
//   TQ_MMA_NC / TQ_MMA_NS / TQ_MMA_CTAS  warp mix of resolve_mma_kernel       TQ_MMA_STAGES  cp.async ring depth
//   TQ_MMA_TMA      ring filled by TMA bulk copies        TQ_MMA_SCALES  4 | 8 mask scales in stage 2
//   TQ_MMA_SPLIT, TQ_MMA_ROLEMASK  placement of the roles over the four schedulers
//   TQ_MMA_TIMERS   clock64 phase accumulators           TQ_DBG_NO_LOAD / _NO_S1 / _NO_S2, TQ_NO_SVD, TQ_COUNT_TILES  knock-outs
//   TQ_STOP2, TQ_TAU2  Jacobi stopping / rotation thresholds       TQ_TMEM, TQ_WS_DEC/INC, TQ_MIN_CTAS  the POPC kernels
//
// Reference semantics restated here (paths relative to the reference repo):
//   tetrad/src/resolve_quartets.py:212-221  gather, missing mask, invariant mask
//   tetrad/src/resolve_quartets.py:42-73    subsample_chunk_to_matrices (first usable SNP/locus)
//   tetrad/src/resolve_quartets.py:76-104   full_chunk_to_matrices
//   tetrad/src/resolve_quartets.py:226-264  nsnps, svd, matrix_rank, minrank, score, argmin
//
// Data layout: locus runs are sorted by length (longest first) and cut into
// tiles of 1024 loci; a tile with K ranks stores, for every tip, K x 32 uint4 {A,C,G,T}:
// word `lane` of rank k holds, at bit b, the k-th column of sorted locus 1024*tile+32*lane+b.
// "First usable column of a locus" (the reference's last_loc state machine) is then
// "lowest rank k whose bit survives the quartet's mask" -- a running OR per lane.
#include "tq_kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace tq {

std::atomic<int64_t> g_launches{0};
std::atomic<const char *> g_last_kernel{"none"};
// Kernel choice and test hooks (tq_set_option in the C ABI; the environment gives the initial values).
//   g_kernel_choice: 0 = by launch size, 1 = fused (one warp per quartet, POPC), 2 = warp-specialised POPC, 3 = tensor-core count
static int initial_kernel_choice()
{
    const char *e = getenv("TQ_KERNEL");
    if (!e) return 0;
    return !strcmp(e, "fused") ? 1 : !strcmp(e, "ws") ? 2 : !strcmp(e, "mma") ? 3 : 0;
}
static int64_t initial_mma_flush()
{
    const char *e = getenv("TQ_MMA_FLUSH");
    return e ? atol(e) : 0;
}
std::atomic<int> g_kernel_choice{initial_kernel_choice()};
std::atomic<int64_t> g_mma_flush{initial_mma_flush()};

static constexpr unsigned kFull = 0xFFFFFFFFu;
#ifndef TQ_TMEM
#define TQ_TMEM 1                             // 1: packed pattern counters live in tensor memory (tcgen05.ld/st)
#endif
#ifndef TQ_WS_DEC
#define TQ_WS_DEC 72                          // registers of a count warp in the warp-specialised kernel
#define TQ_WS_INC 96                          // registers of a score warp; 256*DEC + 128*INC must equal 384*80
#endif
#ifndef TQ_ROWS_IN_FLIGHT
#define TQ_ROWS_IN_FLIGHT 2                    // pattern rows (16 POPCs each) kept in flight in stage 2
#endif
#ifndef TQ_DEFAULT_WS
#define TQ_DEFAULT_WS 1                       // 1: large launches use the warp-specialised kernel (env TQ_KERNEL=ws|fused overrides)
#endif
#ifndef TQ_DEFAULT_MMA
#define TQ_DEFAULT_MMA 1                      // 1: large launches use the tensor-core count kernel (env TQ_KERNEL=mma|ws|fused overrides)
#endif
// Warp mix of the tensor-core kernel, one CTA per SM.  Measured at C2 (M quartets/s) with score warps that take four
// quartets at a time: 8+8 10.88, 7+9 10.42, 8+7 10.18, 7+8 10.05, 9+7 9.70, 10+6 9.67, 10+7 9.50, 6+9 9.28, 9+8 9.22
// (three-matrix score warps before that: 8+11 9.83, 4+6 x2 CTAs 9.32-9.52; profiles/r01_v6_resolve_ncu.md).
#ifndef TQ_MMA_NC
#define TQ_MMA_NC 8                           // count warps per CTA of the tensor-core kernel
#endif
#ifndef TQ_MMA_NS
#define TQ_MMA_NS 8                           // score warps per CTA
#endif
#ifndef TQ_MMA_CTAS
#define TQ_MMA_CTAS 1                         // resident CTAs per SM
#endif
#ifndef TQ_MIN_CTAS
#define TQ_MIN_CTAS (TQ_TMEM ? 5 : 4)         // resident CTAs per SM the register budget is sized for
#endif
static constexpr int kWarps = 4;              // warps (= quartets in flight) per CTA
static constexpr int kLdm = 18;               // padded column length of the smem matrices (doubles)
static constexpr double kEps = 2.220446049250313e-16;
static constexpr int kMaxSweeps = 30;

// ------------------------------------------------------------------------------------------
// K3: counts
// ------------------------------------------------------------------------------------------
// One warp resolves one quartet.  Lane l owns locus word l of every tile (32 loci).
//
// stage 1 (subsample mode): the tip rows are streamed rank-chunk by rank-chunk (512 B per tip per
//   chunk) through a per-warp cp.async ring in shared memory; for every chunk the lane forms
//   sel = complete & segregating & ~taken and ORs the selected columns' one-hot bits into 16
//   merged planes M (one SNP per locus, resolve_quartets.py:58-64).
// stage 2: the 256 pattern counts of a tile are popc(A_a & B_b & C_c & D_d) on the merged planes.
//   To halve the accumulator registers the two half-warps split the patterns: lanes 0-15 count
//   a in {0,1}, lanes 16-31 count a in {2,3}, each for its own word and for the word of its
//   partner lane (l ^ 16), whose planes arrive by 14 warp shuffles.
//   acc[i] packs two 16-bit counters: patterns (a, b, cd=2h) and (a, b, cd=2h+1) with
//   i = ((a & 1) * 4 + b) * 8 + h.  A lane adds at most 64 per call, the 16-lane butterfly of the
//   flush multiplies by 16, so flushing every 63 calls keeps every field below 65536.
#ifdef TQ_PHASE_TIMERS
__device__ unsigned long long g_phase_clk[8];
#define TQ_T0() long long tq_t0 = clock64()
#define TQ_ACC(i) do { long long tq_t1 = clock64(); ph[i] += tq_t1 - tq_t0; tq_t0 = tq_t1; } while (0)
#else
#define TQ_T0()
#define TQ_ACC(i)
#endif
static constexpr int kStages = 4;               // cp.async ring depth (rank-chunks in flight); 6 was slower (L1 shrinks)
static constexpr uint32_t kFlushEvery = 63;

struct __align__(16) SvdScratch {
    double M[3][16][kLdm];                      // flattenings M[m][col][row]; later W (R factor) / exchange buffer
    double nrm[3][16];                          // singular values
    double srt[3][16];                          // sorted singular values; pivot-column buffer during the QR
};

struct __align__(16) WarpScratch {
    union {
        uint4 ring[kStages][4][32];             // stage-1 staging: [stage][tip][lane]      8 KB
        SvdScratch svd;                         // the Jacobi reuses the ring once counting is over
    };
    uint32_t T[256];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#if TQ_TMEM
// Tensor memory as an accumulator file.  Each CTA (4 warps) allocates 64 TMEM columns; warp w owns
// TMEM lanes 32w..32w+31 of those columns (a warp may only touch its own lane quadrant), i.e. every
// thread has 64 private 32-bit cells = its 64 packed counters, which therefore do not occupy
// registers across tiles.  The 32x32b shape moves N consecutive columns of the thread's own lane.
static constexpr uint32_t kTmemCols = 64;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// One (a,b) row = 16 patterns.  The ANDs and the POPCs are volatile asm so that ptxas keeps them in
// the order written: a software pipeline in which every POPC has ~16 independent POPCs and the next
// row's ANDs between its issue and the add that consumes it (left alone, ptxas puts the consumer 1-3
// instructions behind the POPC and the warp sits on the XU latency).
__device__ __forceinline__ void row_and(uint32_t ab, const uint32_t (&C)[4], const uint32_t (&D)[4], uint32_t (&v)[16])
{
#pragma unroll
    for (int cd = 0; cd < 16; ++cd)
        asm volatile("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(v[cd]) : "r"(ab), "r"(C[cd >> 2]), "r"(D[cd & 3]));
}
__device__ __forceinline__ void row_popc(uint32_t (&v)[16])
{
#pragma unroll
    for (int cd = 0; cd < 16; ++cd) asm volatile("popc.b32 %0, %0;" : "+r"(v[cd]));
}
__device__ __forceinline__ void row_add(const uint32_t (&v)[16], uint32_t *acc8)
{
#pragma unroll
    for (int h = 0; h < 8; ++h) acc8[h] += v[2 * h] + (v[2 * h + 1] << 16);
}
#endif

__device__ __forceinline__ void accumulate128(uint32_t A0, uint32_t A1, const uint32_t (&B)[4],
                                              const uint32_t (&C)[4], const uint32_t (&D)[4],
                                              uint32_t (&acc)[64])
{
#pragma unroll
    for (int ai = 0; ai < 2; ++ai)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t ab = (ai ? A1 : A0) & B[b];
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const uint32_t p0 = __popc(ab & C[h >> 1] & D[2 * (h & 1)]);
                const uint32_t p1 = __popc(ab & C[h >> 1] & D[2 * (h & 1) + 1]);
                acc[(ai * 4 + b) * 8 + h] += p0 + (p1 << 16);
            }
        }
}

// 256 pattern counts of one tile word pair (own word + partner lane's word), split by half-warp.
#if TQ_TMEM
__device__ __forceinline__ void count_tile(const uint32_t (&MA)[4], const uint32_t (&MB)[4],
                                           const uint32_t (&MC)[4], const uint32_t (&MD)[4],
                                           bool upper, uint32_t tacc)
{
    // planes the partner needs from me: the a-planes of ITS half, and all of B, C, D
    const uint32_t sendA0 = upper ? MA[0] : MA[2], sendA1 = upper ? MA[1] : MA[3];
    const uint32_t pA0 = __shfl_xor_sync(kFull, sendA0, 16), pA1 = __shfl_xor_sync(kFull, sendA1, 16);
    uint32_t pB[4], pC[4], pD[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pB[i] = __shfl_xor_sync(kFull, MB[i], 16);
        pC[i] = __shfl_xor_sync(kFull, MC[i], 16);
        pD[i] = __shfl_xor_sync(kFull, MD[i], 16);
    }
    const uint32_t A0 = upper ? MA[2] : MA[0], A1 = upper ? MA[3] : MA[1];
    tmem_wait_st();                                  // the previous tile's stores have landed
    // counters 16g .. 16g+15 hold rows (ai, b) = (g / 2, 2 (g % 2) + {0, 1}).  A runtime loop: the
    // counters are addressed in tensor memory, so nothing forces the 256 patterns to be unrolled,
    // and the whole body (64 AND + 64 POPC + 64 add) stays resident in the instruction cache.
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        uint32_t acc[16];
        tmem_ld16(tacc + 16 * g, acc);
        const bool hi = (g & 1) != 0;
        const uint32_t a_own = (g >> 1) ? A1 : A0, a_par = (g >> 1) ? pA1 : pA0;
        const uint32_t o0 = a_own & (hi ? MB[2] : MB[0]), o1 = a_own & (hi ? MB[3] : MB[1]);
        const uint32_t q0 = a_par & (hi ? pB[2] : pB[0]), q1 = a_par & (hi ? pB[3] : pB[1]);
#if TQ_ROWS_IN_FLIGHT == 2
        uint32_t va[16], vb[16];
        row_and(o0, MC, MD, va);
        row_popc(va);
        row_and(o1, MC, MD, vb);
        tmem_wait_ld();
        row_add(va, acc);
        row_popc(vb);
        row_and(q0, pC, pD, va);
        row_add(vb, acc + 8);
        row_popc(va);
        row_and(q1, pC, pD, vb);
        row_add(va, acc);
        row_popc(vb);
        row_add(vb, acc + 8);
#else
        uint32_t va[16];
        row_and(o0, MC, MD, va); row_popc(va);
        tmem_wait_ld();
        row_add(va, acc);
        row_and(o1, MC, MD, va); row_popc(va); row_add(va, acc + 8);
        row_and(q0, pC, pD, va); row_popc(va); row_add(va, acc);
        row_and(q1, pC, pD, va); row_popc(va); row_add(va, acc + 8);
#endif
        tmem_st16(tacc + 16 * g, acc);
    }
}

__device__ __forceinline__ void zero_acc(uint32_t tacc)
{
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    tmem_wait_st();
#pragma unroll
    for (int g = 0; g < 4; ++g) tmem_st16(tacc + 16 * g, z);
}
#else
__device__ __forceinline__ void count_tile(const uint32_t (&MA)[4], const uint32_t (&MB)[4],
                                           const uint32_t (&MC)[4], const uint32_t (&MD)[4],
                                           bool upper, uint32_t (&acc)[64])
{
    // planes the partner needs from me: the a-planes of ITS half, and all of B, C, D
    const uint32_t sendA0 = upper ? MA[0] : MA[2], sendA1 = upper ? MA[1] : MA[3];
    const uint32_t pA0 = __shfl_xor_sync(kFull, sendA0, 16), pA1 = __shfl_xor_sync(kFull, sendA1, 16);
    uint32_t pB[4], pC[4], pD[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pB[i] = __shfl_xor_sync(kFull, MB[i], 16);
        pC[i] = __shfl_xor_sync(kFull, MC[i], 16);
        pD[i] = __shfl_xor_sync(kFull, MD[i], 16);
    }
    accumulate128(upper ? MA[2] : MA[0], upper ? MA[3] : MA[1], MB, MC, MD, acc);
    accumulate128(pA0, pA1, pB, pC, pD, acc);
}
#endif

// Butterfly-add the packed counters over the 16 lanes of each half-warp, add into T, clear.
__device__ __forceinline__ void flush_acc(uint32_t (&acc)[64], uint32_t *T, int lane)
{
    const int abase = (lane >> 4) * 2;           // first a-value of this half-warp
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        uint32_t v = acc[i];
        v += __shfl_xor_sync(kFull, v, 8);
        v += __shfl_xor_sync(kFull, v, 4);
        v += __shfl_xor_sync(kFull, v, 2);
        v += __shfl_xor_sync(kFull, v, 1);
        if ((lane & 15) == (i & 15)) {
            const int ai = i >> 5, b = (i >> 3) & 3, h = i & 7;
            uint2 *t = reinterpret_cast<uint2 *>(T + ((abase + ai) * 4 + b) * 16 + 2 * h);
            uint2 cur = *t;
            cur.x += v & 0xFFFFu;
            cur.y += v >> 16;
            *t = cur;
        }
        acc[i] = 0u;
    }
}

// Issue the four 16-byte cp.async of rank-chunk `nu` (one per tip) into ring stage nu % kStages.
// All global addressing is a 32-bit uint4 index off `planes` (the host checks ntips*row_u4 < 2^32).
__device__ __forceinline__ void issue_chunk(uint32_t ring_s, const uint4 *__restrict__ planes, uint32_t lane_nu,
                                            uint32_t stage, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3)
{
    const uint32_t s = ring_s + stage * (4 * 32 * 16);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(planes + (t0 + lane_nu)) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s + 512), "l"(planes + (t1 + lane_nu)) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s + 1024), "l"(planes + (t2 + lane_nu)) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s + 1536), "l"(planes + (t3 + lane_nu)) : "memory");
}

#if TQ_TMEM
// Read the 64 packed counters back from tensor memory, butterfly-add them over the 16 lanes of
// each half-warp, add into T, and clear them.  Runtime loop over the four groups (code size).
__device__ __noinline__ void flush_tmem(uint32_t tacc, uint32_t *T, int lane)
{
    const int abase = (lane >> 4) * 2;           // first a-value of this half-warp
    tmem_wait_st();
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        uint32_t part[16];
        tmem_ld16(tacc + 16 * g, part);
        tmem_wait_ld();
        uint32_t *row = T + ((abase + (g >> 1)) * 4 + 2 * (g & 1)) * 16;      // pattern (a, b = 2(g%2), cd = 0)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t v = part[k];
            v += __shfl_xor_sync(kFull, v, 8);
            v += __shfl_xor_sync(kFull, v, 4);
            v += __shfl_xor_sync(kFull, v, 2);
            v += __shfl_xor_sync(kFull, v, 1);
            if ((lane & 15) == k) {                                            // k = 8 (b & 1) + h
                uint2 *t = reinterpret_cast<uint2 *>(row + (k >> 3) * 16 + 2 * (k & 7));
                uint2 cur = *t;
                cur.x += v & 0xFFFFu;
                cur.y += v >> 16;
                *t = cur;
            }
            part[k] = 0u;
        }
        tmem_st16(tacc + 16 * g, part);
    }
}
#endif

template <bool kSubsample>
__device__ __forceinline__ void count_quartet(const PlanDev &plan, const uint4 *__restrict__ planes,
                                              uint4 tips, const uint4 *ring, uint32_t *T, int lane, uint32_t debug, uint32_t tacc
#ifdef TQ_PHASE_TIMERS
                                              , long long (&ph)[8]
#endif
                                              )
{
    TQ_T0();
#if TQ_TMEM
    const uint32_t acc = tacc;                     // handle of this thread's 64 counters in tensor memory
#define TQ_FLUSH() flush_tmem(acc, T, lane)
#else
    uint32_t acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0u;
#define TQ_FLUSH() flush_acc(acc, T, lane)
#endif
    const bool upper = lane >= 16;
    uint32_t pending = 0;
    const uint32_t nchunks = plan.row_u4 / kTileWords;
    const uint32_t t0 = tips.x * plan.row_u4, t1 = tips.y * plan.row_u4;    // row starts, uint4 units
    const uint32_t t2 = tips.z * plan.row_u4, t3 = tips.w * plan.row_u4;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring + lane);

    // prologue: chunks 0 .. kStages-2 in flight
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if ((uint32_t)s < nchunks) issue_chunk(ring_s, planes, s * kTileWords + lane, s, t0, t1, t2, t3);
        cp_async_commit();
    }

    uint32_t u = 0;                                   // rank-chunk index along the rows
    for (int tile = 0; tile < plan.ntiles; ++tile) {
        const uint32_t K = __ldg(plan.tile_k + tile);
        uint32_t taken = 0u;
        uint32_t MA[4] = {0u, 0u, 0u, 0u}, MB[4] = {0u, 0u, 0u, 0u};
        uint32_t MC[4] = {0u, 0u, 0u, 0u}, MD[4] = {0u, 0u, 0u, 0u};
        for (uint32_t k = 0; k < K; ++k, ++u) {
            {   // keep kStages-1 chunks in flight
                const uint32_t nu = u + kStages - 1;
                if (nu < nchunks) issue_chunk(ring_s, planes, nu * kTileWords + lane, nu % kStages, t0, t1, t2, t3);
                cp_async_commit();
            }
            cp_async_wait<kStages - 1>();
            const int s = u % kStages;
            const uint4 *st = ring + s * (4 * 32) + lane;
            const uint4 a = st[0], b = st[32], c = st[64], d = st[96];
            if (kSubsample && (debug & 4u)) {
                MA[0] |= a.x; MB[0] |= b.y; MC[0] |= c.z; MD[0] |= d.w;
            } else if (kSubsample) {
                const uint32_t valid = (a.x | a.y | a.z | a.w) & (b.x | b.y | b.z | b.w) &
                                       (c.x | c.y | c.z | c.w) & (d.x | d.y | d.z | d.w);
                const uint32_t same = (a.x & b.x & c.x & d.x) | (a.y & b.y & c.y & d.y) |
                                      (a.z & b.z & c.z & d.z) | (a.w & b.w & c.w & d.w);
                const uint32_t sel = valid & ~same & ~taken;
                taken |= sel;
                MA[0] |= a.x & sel; MA[1] |= a.y & sel; MA[2] |= a.z & sel; MA[3] |= a.w & sel;
                MB[0] |= b.x & sel; MB[1] |= b.y & sel; MB[2] |= b.z & sel; MB[3] |= b.w & sel;
                MC[0] |= c.x & sel; MC[1] |= c.y & sel; MC[2] |= c.z & sel; MC[3] |= c.w & sel;
                MD[0] |= d.x & sel; MD[1] |= d.y & sel; MD[2] |= d.z & sel; MD[3] |= d.w & sel;
            } else {
                // full mode (resolve_quartets.py:92-95): every complete column counts; the four
                // invariant patterns aaaa are zeroed by the caller
                const uint32_t A[4] = {a.x, a.y, a.z, a.w}, B[4] = {b.x, b.y, b.z, b.w};
                const uint32_t C[4] = {c.x, c.y, c.z, c.w}, D[4] = {d.x, d.y, d.z, d.w};
                count_tile(A, B, C, D, upper, acc);
                if (++pending == kFlushEvery) { TQ_FLUSH(); pending = 0; }
            }
        }
        TQ_ACC(0);
        if (kSubsample && !(debug & 2u)) {
            count_tile(MA, MB, MC, MD, upper, acc);
            if (++pending == kFlushEvery) { TQ_FLUSH(); pending = 0; }
        }
        TQ_ACC(1);
    }
    cp_async_wait<0>();
    TQ_FLUSH();
    TQ_ACC(2);
#undef TQ_FLUSH
}

// ------------------------------------------------------------------------------------------
// K4: scores.  One-sided (Hestenes) Jacobi on the three 16x16 flattenings, columns held in
// REGISTERS.  Lane (m, s) = (lane / 8, lane % 8) owns two columns P and Q of matrix m (lanes
// 24-31 idle on zero columns).  A sweep is a recursive-halving tournament of 8 + 4 + 2 + 1 = 15
// steps that meets every pair of the 16 columns once: inside a group of g lanes the Q columns
// circulate by one lane per step (34 warp shuffles), and between phases the lower half of each
// group trades its Q for its partner's P through a small shared-memory buffer.  Shared-memory
// traffic per step is ~4x lower than keeping the matrices in shared memory, which is what bounded
// the first versions of this kernel (profiles/).  The rotation tangent uses approximate
// reciprocal / rsqrt (a slightly-off angle only slows convergence); cos is refined to full FP64.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

#ifndef TQ_TAU2
#define TQ_TAU2 1e-18
#endif
static constexpr double kTau2 = TQ_TAU2;  // rotate when gamma^2 > tau^2 * |x|^2 |y|^2  (tau = 1e-9)
#ifndef TQ_STOP2
#define TQ_STOP2 1e-8
#endif
static constexpr double kStop2 = TQ_STOP2;    // a sweep whose largest relative gamma was < 1e-4 is the last

// Self-scaling ("fast") rotations: a column is stored as s * v with a scalar s, its reciprocal r = 1 / s travels along.
// Rotating x = sP P, y = sQ Q by (cos, sin) is exactly
//     x' = (cos sP) [P - tan (sQ / sP) Q],     y' = (cos sQ) [Q + tan (sP / sQ) P],
// so the 16 rows cost two FMAs each instead of two multiplies and two FMAs, and the factor cos goes into the scales
// (FP64 instructions are what the score warps wait for: profiles/r01_v7_resolve_ncu.md).  The scales are folded
// back into the columns once per sweep (fold_scales), so they stay within a few orders of magnitude of 1.
__device__ __forceinline__ void fold_scales(double (&P)[16], double (&Q)[16], double &sP, double &rP, double &sQ, double &rQ)
{
#pragma unroll
    for (int r = 0; r < 16; ++r) { P[r] *= sP; Q[r] *= sQ; }
    sP = 1.0; rP = 1.0; sQ = 1.0; rQ = 1.0;
}

__device__ __forceinline__ void jacobi_rotate(double (&P)[16], double (&Q)[16], double &sP, double &rP, double &sQ, double &rQ,
                                              double &nP, double &nQ, double tiny, bool &big, bool force = false,
                                              bool enable = true)
{
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll
    for (int r = 0; r < 16; r += 4) {
        g0 = fma(P[r], Q[r], g0); g1 = fma(P[r + 1], Q[r + 1], g1);
        g2 = fma(P[r + 2], Q[r + 2], g2); g3 = fma(P[r + 3], Q[r + 3], g3);
    }
    const double ga = (sP * sQ) * ((g0 + g1) + (g2 + g3));
    const double ab = nP * nQ, gg = ga * ga;
    if (force || (enable && nP > tiny && nQ > tiny && gg > kTau2 * ab)) {
        big |= gg > kStop2 * ab;
        // Rotation that zeroes gamma:  with a = |Q|^2 - |P|^2, b = 2 gamma, r = sqrt(a^2 + b^2):
        //   cos^2 = (1 + |a| / r) / 2,   sin = sign(a) * b / (2 r cos),   tan = sin / cos.
        // Only two special-function ops (approximate FP64 rsqrt) sit on the dependency chain; the
        // pair (cos, sin) is then renormalised so that cos^2 + sin^2 = 1 to FP64 accuracy (the
        // angle itself may be ~1e-6 off, which only slows convergence).
        const double a = nQ - nP, b2 = ga + ga;
        const double q = fma(a, a, b2 * b2);
        double ir;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(ir) : "d"(q));       // ~ 1 / r
        const double c2 = fma(0.5 * fabs(a), ir, 0.5);                   // cos^2 in [0.5, 1]
        double rc;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(c2));      // ~ 1 / cos
        double cs = c2 * rc;
        double sn = (a < 0.0 ? -0.5 : 0.5) * b2 * ir * rc;
        const double e = fma(cs, cs, fma(sn, sn, -1.0));                 // cos^2 + sin^2 - 1, ~1e-6
        const double f = fma(e, fma(0.375, e, -0.5), 1.0);               // (1 + e)^(-1/2) to O(e^3)
        cs *= f;
        sn *= f;
        double ic = rc;                                                  // 1 / cos: two Newton steps from the approximation
        ic = ic * fma(-cs, ic, 2.0);
        ic = ic * fma(-cs, ic, 2.0);
        const double t = sn * ic;                                        // tan
        const double al = t * (sQ * rP), be = t * (sP * rQ);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const double p = P[r], qq = Q[r];
            P[r] = fma(-al, qq, p);
            Q[r] = fma(be, p, qq);
        }
        sP *= cs; sQ *= cs; rP *= ic; rQ *= ic;
        const double tg = t * ga;
        nP -= tg;
        nQ += tg;
    }
}

__device__ __noinline__ void svd_scores(const uint32_t *T, SvdScratch &sv, int lane, double *score_out, uint32_t *best_out,
                                        uint32_t *minrank_out, uint32_t debug)
{
    double v[8];
    double fro2 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = (double)T[lane + 32 * i];
        fro2 = fma(v[i], v[i], fro2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fro2 += __shfl_xor_sync(kFull, fro2, o);
    const double tiny = (kEps * kEps) * fro2;       // columns this small are null directions
    __syncwarp();
    // the three flattenings (resolve_quartets.py:63,67-72), stored column-major
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int p = lane + 32 * i;
        const int a = p >> 6, b = (p >> 4) & 3, c = (p >> 2) & 3, d = p & 3;
        sv.M[0][4 * c + d][4 * a + b] = v[i];       // rows (a,b), cols (c,d)
        sv.M[1][4 * b + d][4 * a + c] = v[i];       // rows (a,c), cols (b,d)
        sv.M[2][4 * b + c][4 * a + d] = v[i];       // rows (a,d), cols (b,c)
    }
    __syncwarp();

    const int m = lane >> 3, s = lane & 7;
    double P[16], Q[16];
    if (lane < 24) {
        const double2 *cp = reinterpret_cast<const double2 *>(&sv.M[m][s][0]);
        const double2 *cq = reinterpret_cast<const double2 *>(&sv.M[m][s + 8][0]);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const double2 x = cp[r], y = cq[r];
            P[2 * r] = x.x; P[2 * r + 1] = x.y; Q[2 * r] = y.x; Q[2 * r + 1] = y.y;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) { P[r] = 0.0; Q[r] = 0.0; }
    }
    __syncwarp();                                    // M is free from here on: reuse it below

    // Preconditioner (Drmac-Veselic): A P = Q R by modified Gram-Schmidt with column pivoting, then
    // Jacobi on the ROWS of R instead of the columns of A -- same singular values, ~40 % fewer sweeps
    // (numpy model: 9.4 -> 5.5 per quartet).  Lane (m, s) still owns columns s and s+8 of A; step k
    // picks the remaining column of largest norm, broadcasts it inside the 8-lane group, and every lane
    // writes r_kj = q_k . a_j for its two columns to W[m][k][j] (j = ORIGINAL column index: a column
    // permutation of R does not change its singular values) and removes that component.
    if (!(debug & 128u)) {
        double *W = &sv.M[0][0][0];                  // [3][16 rows k][16 cols j], 6 KB
        bool doneP = lane >= 24, doneQ = lane >= 24;
        double mP = 0.0, mQ = 0.0;                   // squared norms of the remaining parts of P, Q
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            if ((k & 3) == 0) {                      // exact every 4th step, downdated in between
                mP = 0.0; mQ = 0.0;
#pragma unroll
                for (int r = 0; r < 16; ++r) { mP = fma(P[r], P[r], mP); mQ = fma(Q[r], Q[r], mQ); }
            }
            const double nP = doneP ? -1.0 : mP, nQ = doneQ ? -1.0 : mQ;
            double best = (nP >= nQ) ? nP : nQ;
            int idx = (nP >= nQ) ? s : s + 8;
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) {
                const double ob = __shfl_xor_sync(kFull, best, off);
                const int oi = __shfl_xor_sync(kFull, idx, off);
                if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
            }
            const bool any = best >= 0.0;            // a column of this matrix is still unprocessed
            // the pivot column goes through shared memory (one writer per 8-lane group)
            double2 *pv = reinterpret_cast<double2 *>(&sv.srt[lane < 24 ? m : 0][0]);
            if (lane < 24 && idx == s) {
#pragma unroll
                for (int r = 0; r < 8; ++r) pv[r] = make_double2(P[2 * r], P[2 * r + 1]);
            } else if (lane < 24 && idx == s + 8) {
#pragma unroll
                for (int r = 0; r < 8; ++r) pv[r] = make_double2(Q[2 * r], Q[2 * r + 1]);
            }
            __syncwarp();
            double dP = 0.0, dQ = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 x = pv[r];
                dP = fma(x.x, P[2 * r], dP); dP = fma(x.y, P[2 * r + 1], dP);
                dQ = fma(x.x, Q[2 * r], dQ); dQ = fma(x.y, Q[2 * r + 1], dQ);
            }
            // 1/sqrt(|a_p|^2) from the pivot's own dot product (exact norm, not the downdated one)
            const double pp = __shfl_sync(kFull, (idx == s) ? dP : dQ, (lane & ~7) | (idx & 7));
            const bool live = any && pp > tiny;      // else the pivot is a null direction: retire it with a zero row
            const double safe = live ? pp : 1.0;
            double rinv;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rinv) : "d"(safe));
            rinv = rinv * fma(-0.5 * safe, rinv * rinv, 1.5);
            rinv = rinv * fma(-0.5 * safe, rinv * rinv, 1.5);
            const double inv = rinv * rinv;
            const bool useP = live && !doneP, useQ = live && !doneQ;
            const double rP = useP ? dP * rinv : 0.0, rQ = useQ ? dQ * rinv : 0.0;
            if (lane < 24) {
                W[(m * 16 + k) * 16 + s] = rP;
                W[(m * 16 + k) * 16 + s + 8] = rQ;
            }
            mP = fmax(fma(-rP, rP, mP), 0.0);
            mQ = fmax(fma(-rQ, rQ, mQ), 0.0);
            const double cP = useP ? -dP * inv : 0.0, cQ = useQ ? -dQ * inv : 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 x = pv[r];
                P[2 * r] = fma(cP, x.x, P[2 * r]); P[2 * r + 1] = fma(cP, x.y, P[2 * r + 1]);
                Q[2 * r] = fma(cQ, x.x, Q[2 * r]); Q[2 * r + 1] = fma(cQ, x.y, Q[2 * r + 1]);
            }
            __syncwarp();
            if (any && idx == s) doneP = true;
            if (any && idx == s + 8) doneQ = true;
        }
        __syncwarp();
        if (lane < 24) {                             // rows s and s+8 of R become this lane's columns
            const double2 *cp = reinterpret_cast<const double2 *>(W + (m * 16 + s) * 16);
            const double2 *cq = reinterpret_cast<const double2 *>(W + (m * 16 + s + 8) * 16);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 x = cp[r], y = cq[r];
                P[2 * r] = x.x; P[2 * r + 1] = x.y; Q[2 * r] = y.x; Q[2 * r + 1] = y.y;
            }
        }
        __syncwarp();
    }
    double2 *xbuf = reinterpret_cast<double2 *>(&sv.M[0][0][0]);   // [32 lanes][11 double2] (176 B stride: conflict-free)

    // A matrix stops rotating after ITS last sweep, whatever the other two still need: every flattening then goes
    // through exactly the same arithmetic here and in svd_scores4, and the three kernel variants agree bit for bit.
    bool gdone = lane >= 24;
    double sP = 1.0, rP = 1.0, sQ = 1.0, rQ = 1.0;   // column scales of the self-scaling rotations and their reciprocals
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        fold_scales(P, Q, sP, rP, sQ, rQ);
        double nP = 0.0, nQ = 0.0;                   // refresh the cached squared norms
#pragma unroll
        for (int r = 0; r < 16; ++r) { nP = fma(P[r], P[r], nP); nQ = fma(Q[r], Q[r], nQ); }
        bool big = false;
#pragma unroll 1
        for (int g = 8; g >= 1; g >>= 1) {
            for (int r = 0; r < g; ++r) {
                jacobi_rotate(P, Q, sP, rP, sQ, rQ, nP, nQ, tiny, big, (debug & 64u) != 0, !gdone);
                if (r + 1 < g && !(debug & 16u)) {   // Q moves one lane down inside its group of g
                    const int src = (lane & ~(g - 1)) | ((lane + 1) & (g - 1));
#pragma unroll
                    for (int k = 0; k < 16; ++k) Q[k] = __shfl_sync(kFull, Q[k], src);
                    nQ = __shfl_sync(kFull, nQ, src); sQ = __shfl_sync(kFull, sQ, src); rQ = __shfl_sync(kFull, rQ, src);
                }
            }
            if (g > 1 && !(debug & 16u)) {
                // split every group in two: lower lanes keep P and take the partner's P as Q,
                // upper lanes keep Q and take the partner's Q as P
                const int h = g >> 1;
                const bool lower = (lane & (g - 1)) < h;
                double2 *mine = xbuf + lane * 11;
                if (lower) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) mine[k] = make_double2(Q[2 * k], Q[2 * k + 1]);
                    mine[8] = make_double2(nQ, sQ); mine[9] = make_double2(rQ, 0.0);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) mine[k] = make_double2(P[2 * k], P[2 * k + 1]);
                    mine[8] = make_double2(nP, sP); mine[9] = make_double2(rP, 0.0);
                }
                __syncwarp();
                const double2 *theirs = xbuf + (lower ? lane + h : lane - h) * 11;
                if (lower) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { const double2 x = theirs[k]; Q[2 * k] = x.x; Q[2 * k + 1] = x.y; }
                    nQ = theirs[8].x; sQ = theirs[8].y; rQ = theirs[9].x;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { const double2 x = theirs[k]; P[2 * k] = x.x; P[2 * k + 1] = x.y; }
                    nP = theirs[8].x; sP = theirs[8].y; rP = theirs[9].x;
                }
                __syncwarp();
            }
        }
        if (debug & 32u) { if (sweep == 8) break; else continue; }
        const uint32_t bigs = __ballot_sync(kFull, big);
        if (((bigs >> (lane & ~7)) & 0xFFu) == 0u) gdone = true;
        if (__all_sync(kFull, gdone)) break;
    }

    // singular values = column norms; sort descending per matrix by rank counting
    {
        fold_scales(P, Q, sP, rP, sQ, rQ);
        double nP = 0.0, nQ = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) { nP = fma(P[r], P[r], nP); nQ = fma(Q[r], Q[r], nQ); }
        __syncwarp();
        if (lane < 24) { sv.nrm[m][s] = sqrt(nP); sv.nrm[m][s + 8] = sqrt(nQ); }
    }
    __syncwarp();
    for (int c = lane; c < 48; c += 32) {
        const int mm = c >> 4, col = c & 15;
        const double val = sv.nrm[mm][col];
        int pos = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const double u = sv.nrm[mm][j];
            pos += (u > val || (u == val && j < col)) ? 1 : 0;
        }
        sv.srt[mm][pos] = val;
    }
    __syncwarp();
    // rank = #{sigma > sigma_max * 16 * eps} (numpy.linalg.matrix_rank), :243
    int rank = 16;
    if (lane < 3) {
        const double tol = sv.srt[lane][0] * (16.0 * kEps);
        rank = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) rank += (sv.srt[lane][j] > tol) ? 1 : 0;
    }
    int mr = min(__shfl_sync(kFull, rank, 0), min(__shfl_sync(kFull, rank, 1), __shfl_sync(kFull, rank, 2)));
    mr = min(mr, 10);                                                  // :246
    double sc = 0.0;
    if (lane < 3) {
        for (int j = mr; j < 16; ++j) sc = fma(sv.srt[lane][j], sv.srt[lane][j], sc);
        sc = sqrt(sc);                                                 // :248
    }
    const double s0 = __shfl_sync(kFull, sc, 0), s1 = __shfl_sync(kFull, sc, 1), s2 = __shfl_sync(kFull, sc, 2);
    uint32_t best = 0;                                                 // first argmin, :251
    double lo = s0;
    if (s1 < lo) { best = 1; lo = s1; }
    if (s2 < lo) { best = 2; }
    score_out[0] = s0; score_out[1] = s1; score_out[2] = s2;
    *best_out = best;
    *minrank_out = (uint32_t)mr;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Scores of up to FOUR quartets at a time (tensor-core kernel).  svd_scores above keeps lanes 24-31 idle: three
// matrices of 8 lanes each.  Here the 12 flattenings of four quartets go through the same QR + Jacobi code in three
// passes of four matrices, all 32 lanes busy: a quarter fewer passes per quartet.  Pass r, lane group m = lane / 8
// works on matrix 4r + m = flattening (4r + m) % 3 of quartet (4r + m) / 3; the sorted singular values of the 12
// matrices are parked in shared memory and lanes 0..3 finish one quartet each (rank, minrank, tail norms, argmin).
// ------------------------------------------------------------------------------------------
struct __align__(16) SvdScratch4 {
    double M[4][16][kLdm];                      // flattenings M[m][col][row]; later W (R factors) / exchange buffer
    double piv[4][16];                          // pivot-column buffer during the QR; unsorted singular values
    double sig[12][16];                         // sorted singular values of the 12 matrices
};

// Tbase: the CTA's slot array (256 counts per slot); slots: slot index of quartet j in byte j; actmask bit j: quartet j
// is present and has SNPs.  Lane j < 4 receives the result of quartet j.
__device__ __noinline__ void svd_scores4(const uint32_t *Tbase, uint32_t slots, uint32_t actmask, SvdScratch4 &sv, int lane,
                                         double (&score_out)[3], uint32_t &best_out, uint32_t &minrank_out)
{
    const int m = lane >> 3, s = lane & 7;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int mi = 4 * pass + m, qi = mi / 3, f = mi - 3 * qi;
        const bool act = (actmask >> qi) & 1u;
        if (!__any_sync(kFull, act)) continue;      // none of this pass's four matrices exists (fewer than four quartets)
        const uint32_t *T = Tbase + ((slots >> (8 * qi)) & 0xFFu) * 256u;
        __syncwarp();
        // this group's flattening (resolve_quartets.py:63,67-72), stored column-major; 32 of its 256 entries per lane
        double fro2 = 0.0;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int p = s + 8 * i;
            const int a = p >> 6, b = (p >> 4) & 3, c = (p >> 2) & 3, d = p & 3;
            const double v = act ? (double)T[p] : 0.0;
            const int col = (f == 0) ? 4 * c + d : (f == 1) ? 4 * b + d : 4 * b + c;
            const int row = (f == 0) ? 4 * a + b : (f == 1) ? 4 * a + c : 4 * a + d;
            sv.M[m][col][row] = v;
            fro2 = fma(v, v, fro2);
        }
        fro2 += __shfl_xor_sync(kFull, fro2, 4);
        fro2 += __shfl_xor_sync(kFull, fro2, 2);
        fro2 += __shfl_xor_sync(kFull, fro2, 1);
        const double tiny = (kEps * kEps) * fro2;       // columns this small are null directions
        __syncwarp();
        double P[16], Q[16];
        {
            const double2 *cp = reinterpret_cast<const double2 *>(&sv.M[m][s][0]);
            const double2 *cq = reinterpret_cast<const double2 *>(&sv.M[m][s + 8][0]);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 x = cp[r], y = cq[r];
                P[2 * r] = x.x; P[2 * r + 1] = x.y; Q[2 * r] = y.x; Q[2 * r + 1] = y.y;
            }
        }
        __syncwarp();                                    // M is free from here on: reuse it below

        // QR preconditioner, as in svd_scores: pivoted modified Gram-Schmidt inside the 8-lane group, then the Jacobi
        // runs on the rows of R.
        {
            double *W = &sv.M[0][0][0];                  // [4][16 rows k][16 cols j], 8 KB
            bool doneP = !act, doneQ = !act;
            double mP = 0.0, mQ = 0.0;                   // squared norms of the remaining parts of P, Q
#pragma unroll 1
            for (int k = 0; k < 16; ++k) {
                if ((k & 3) == 0) {                      // exact every 4th step, downdated in between
                    mP = 0.0; mQ = 0.0;
#pragma unroll
                    for (int r = 0; r < 16; ++r) { mP = fma(P[r], P[r], mP); mQ = fma(Q[r], Q[r], mQ); }
                }
                const double nP = doneP ? -1.0 : mP, nQ = doneQ ? -1.0 : mQ;
                double best = (nP >= nQ) ? nP : nQ;
                int idx = (nP >= nQ) ? s : s + 8;
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) {
                    const double ob = __shfl_xor_sync(kFull, best, off);
                    const int oi = __shfl_xor_sync(kFull, idx, off);
                    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
                }
                const bool any = best >= 0.0;            // a column of this matrix is still unprocessed
                double2 *pv = reinterpret_cast<double2 *>(&sv.piv[m][0]);
                if (idx == s) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) pv[r] = make_double2(P[2 * r], P[2 * r + 1]);
                } else if (idx == s + 8) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) pv[r] = make_double2(Q[2 * r], Q[2 * r + 1]);
                }
                __syncwarp();
                double dP = 0.0, dQ = 0.0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double2 x = pv[r];
                    dP = fma(x.x, P[2 * r], dP); dP = fma(x.y, P[2 * r + 1], dP);
                    dQ = fma(x.x, Q[2 * r], dQ); dQ = fma(x.y, Q[2 * r + 1], dQ);
                }
                const double pp = __shfl_sync(kFull, (idx == s) ? dP : dQ, (lane & ~7) | (idx & 7));
                const bool live = any && pp > tiny;      // else the pivot is a null direction: retire it with a zero row
                const double safe = live ? pp : 1.0;
                double rinv;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rinv) : "d"(safe));
                rinv = rinv * fma(-0.5 * safe, rinv * rinv, 1.5);
                rinv = rinv * fma(-0.5 * safe, rinv * rinv, 1.5);
                const double inv = rinv * rinv;
                const bool useP = live && !doneP, useQ = live && !doneQ;
                const double rP = useP ? dP * rinv : 0.0, rQ = useQ ? dQ * rinv : 0.0;
                W[(m * 16 + k) * 16 + s] = rP;
                W[(m * 16 + k) * 16 + s + 8] = rQ;
                mP = fmax(fma(-rP, rP, mP), 0.0);
                mQ = fmax(fma(-rQ, rQ, mQ), 0.0);
                const double cP = useP ? -dP * inv : 0.0, cQ = useQ ? -dQ * inv : 0.0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double2 x = pv[r];
                    P[2 * r] = fma(cP, x.x, P[2 * r]); P[2 * r + 1] = fma(cP, x.y, P[2 * r + 1]);
                    Q[2 * r] = fma(cQ, x.x, Q[2 * r]); Q[2 * r + 1] = fma(cQ, x.y, Q[2 * r + 1]);
                }
                __syncwarp();
                if (any && idx == s) doneP = true;
                if (any && idx == s + 8) doneQ = true;
            }
            __syncwarp();
            {                                            // rows s and s+8 of R become this lane's columns
                const double2 *cp = reinterpret_cast<const double2 *>(W + (m * 16 + s) * 16);
                const double2 *cq = reinterpret_cast<const double2 *>(W + (m * 16 + s + 8) * 16);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const double2 x = cp[r], y = cq[r];
                    P[2 * r] = x.x; P[2 * r + 1] = x.y; Q[2 * r] = y.x; Q[2 * r + 1] = y.y;
                }
            }
            __syncwarp();
        }
        double2 *xbuf = reinterpret_cast<double2 *>(&sv.M[0][0][0]);   // [32 lanes][11 double2] (176 B stride: conflict-free)

        // A matrix stops rotating after ITS last sweep, whatever the other three of the pass still need: the result of a
        // quartet must not depend on which quartets happen to share its batch (run-to-run reproducibility).
        bool gdone = !act;
        double sP = 1.0, rP = 1.0, sQ = 1.0, rQ = 1.0;   // column scales of the self-scaling rotations and their reciprocals
        for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
            fold_scales(P, Q, sP, rP, sQ, rQ);
            double nP = 0.0, nQ = 0.0;                   // refresh the cached squared norms
#pragma unroll
            for (int r = 0; r < 16; ++r) { nP = fma(P[r], P[r], nP); nQ = fma(Q[r], Q[r], nQ); }
            bool big = false;
#pragma unroll 1
            for (int g = 8; g >= 1; g >>= 1) {
                for (int r = 0; r < g; ++r) {
                    jacobi_rotate(P, Q, sP, rP, sQ, rQ, nP, nQ, tiny, big, false, !gdone);
                    if (r + 1 < g) {                     // Q moves one lane down inside its group of g
                        const int src = (lane & ~(g - 1)) | ((lane + 1) & (g - 1));
#pragma unroll
                        for (int k = 0; k < 16; ++k) Q[k] = __shfl_sync(kFull, Q[k], src);
                        nQ = __shfl_sync(kFull, nQ, src); sQ = __shfl_sync(kFull, sQ, src); rQ = __shfl_sync(kFull, rQ, src);
                    }
                }
                if (g > 1) {
                    const int h = g >> 1;
                    const bool lower = (lane & (g - 1)) < h;
                    double2 *mine = xbuf + lane * 11;
                    if (lower) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) mine[k] = make_double2(Q[2 * k], Q[2 * k + 1]);
                        mine[8] = make_double2(nQ, sQ); mine[9] = make_double2(rQ, 0.0);
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) mine[k] = make_double2(P[2 * k], P[2 * k + 1]);
                        mine[8] = make_double2(nP, sP); mine[9] = make_double2(rP, 0.0);
                    }
                    __syncwarp();
                    const double2 *theirs = xbuf + (lower ? lane + h : lane - h) * 11;
                    if (lower) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) { const double2 x = theirs[k]; Q[2 * k] = x.x; Q[2 * k + 1] = x.y; }
                        nQ = theirs[8].x; sQ = theirs[8].y; rQ = theirs[9].x;
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) { const double2 x = theirs[k]; P[2 * k] = x.x; P[2 * k + 1] = x.y; }
                        nP = theirs[8].x; sP = theirs[8].y; rP = theirs[9].x;
                    }
                    __syncwarp();
                }
            }
            const uint32_t bigs = __ballot_sync(kFull, big);
            if (((bigs >> (lane & ~7)) & 0xFFu) == 0u) gdone = true;
            if (__all_sync(kFull, gdone)) break;
        }

        // singular values = column norms; sort descending per matrix by rank counting
        {
            fold_scales(P, Q, sP, rP, sQ, rQ);
            double nP = 0.0, nQ = 0.0;
#pragma unroll
            for (int r = 0; r < 16; ++r) { nP = fma(P[r], P[r], nP); nQ = fma(Q[r], Q[r], nQ); }
            __syncwarp();
            sv.piv[m][s] = sqrt(nP); sv.piv[m][s + 8] = sqrt(nQ);
        }
        __syncwarp();
        for (int c = lane; c < 64; c += 32) {
            const int mm = c >> 4, col = c & 15;
            const double val = sv.piv[mm][col];
            int pos = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double u = sv.piv[mm][j];
                pos += (u > val || (u == val && j < col)) ? 1 : 0;
            }
            sv.sig[4 * pass + mm][pos] = val;
        }
        __syncwarp();
    }
    __syncwarp();
    // lane j < 4 finishes quartet j: rank = #{sigma > sigma_max * 16 * eps} (numpy.linalg.matrix_rank, :243),
    // minrank (:246), tail norms (:248), first argmin (:251)
    if (lane < 4 && ((actmask >> lane) & 1u)) {
        int mr = 16;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const double *sg = sv.sig[3 * lane + f];
            const double tol = sg[0] * (16.0 * kEps);
            int rank = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) rank += (sg[j] > tol) ? 1 : 0;
            mr = min(mr, rank);
        }
        mr = min(mr, 10);
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const double *sg = sv.sig[3 * lane + f];
            double sc = 0.0;
            for (int j = mr; j < 16; ++j) sc = fma(sg[j], sg[j], sc);
            score_out[f] = sqrt(sc);
        }
        uint32_t best = 0;
        double lo = score_out[0];
        if (score_out[1] < lo) { best = 1; lo = score_out[1]; }
        if (score_out[2] < lo) { best = 2; }
        best_out = best;
        minrank_out = (uint32_t)mr;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// fused resolve kernel: one warp per quartet, persistent grid-stride over groups of kWarps
// consecutive quartets (neighbours in lexicographic order share tips -> L1 hits)
// ------------------------------------------------------------------------------------------
template <bool kSubsample>
__global__ void __launch_bounds__(kWarps * 32, TQ_MIN_CTAS)
resolve_kernel(PlanDev plan, const uint4 *__restrict__ planes, uint32_t ntips,
               const uint4 *__restrict__ quartets, int64_t m, uint64_t seed,
               double *__restrict__ rscor, uint2 *__restrict__ rstat,
               uint32_t *__restrict__ counts, uint32_t *__restrict__ minrank_out,
               uint32_t *__restrict__ status, uint32_t debug, uint32_t num_sms, uint32_t stagger_clk)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch *scratch = reinterpret_cast<WarpScratch *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpScratch &ws = scratch[warp];
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
#ifdef TQ_PHASE_TIMERS
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tq_start = clock64();
#endif
#if TQ_TMEM
    __shared__ uint32_t tmem_base_slot;
    if (warp == 0) {                                  // one warp allocates (and later frees) the CTA's columns
        const uint32_t slot = (uint32_t)__cvta_generic_to_shared(&tmem_base_slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's lane quadrant
    zero_acc(tacc);
#else
    const uint32_t tacc = 0u;
#endif

    // Every quartet costs the same, so co-resident warps would run their count and SVD phases in
    // lock-step and leave the POPC pipe idle during the SVD phases.  Delay the k-th resident CTA of
    // an SM by k quarter-periods once, at start: afterwards one warp per scheduler is in its SVD
    // while the other three count.
    if (stagger_clk) {
        const uint32_t k = (blockIdx.x / num_sms) & 3u;
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)k * stagger_clk) __nanosleep(2000);
    }

    for (int64_t q = (int64_t)blockIdx.x * kWarps + warp; q < m; q += nwarps) {
        const uint4 t = __ldg(quartets + q);
        if (t.x >= ntips || t.y >= ntips || t.z >= ntips || t.w >= ntips) {
            if (lane == 0) {
                rstat[q] = make_uint2(0xFFFFFFFFu, 0u);
                rscor[3 * q] = rscor[3 * q + 1] = rscor[3 * q + 2] = 0.0;
                if (status) atomicOr(status, 1u);
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) ws.T[lane + 32 * i] = 0u;
        __syncwarp();
#ifdef TQ_PHASE_TIMERS
        count_quartet<kSubsample>(plan, planes, t, &ws.ring[0][0][0], ws.T, lane, debug, tacc, ph);
        long long tq_t0 = clock64();
#else
        count_quartet<kSubsample>(plan, planes, t, &ws.ring[0][0][0], ws.T, lane, debug, tacc);
#endif
        __syncwarp();
        if (!kSubsample && lane < 4) ws.T[85u * lane] = 0u;   // invariant patterns aaaa, :218
        __syncwarp();
        uint32_t nsnps = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t v = ws.T[lane + 32 * i];
            nsnps += v;
            if (counts) counts[q * 256 + lane + 32 * i] = v;
        }
        nsnps = __reduce_add_sync(kFull, nsnps);                        // :226

        double score[3] = {0.001, 0.001, 0.001};                        // :232
        uint32_t best, minrank = 0;
#ifdef TQ_NO_SVD
        if (true) {
#else
        if (nsnps == 0 || (debug & 1u)) {
#endif
            uint64_t h = seed;
            h = splitmix64(h ^ t.x); h = splitmix64(h ^ t.y);
            h = splitmix64(h ^ t.z); h = splitmix64(h ^ t.w);
            best = (uint32_t)(h % 3ull);                                // stands in for :231
        } else {
#ifndef TQ_NO_SVD
            svd_scores(ws.T, ws.svd, lane, score, &best, &minrank, debug);
#endif
        }
        TQ_ACC(3);
        if (lane == 0) {
            rscor[3 * q] = score[0]; rscor[3 * q + 1] = score[1]; rscor[3 * q + 2] = score[2];
            rstat[q] = make_uint2(best, nsnps);                         // :251, :264
            if (minrank_out) minrank_out[q] = minrank;
        }
        __syncwarp();
    }
#if TQ_TMEM
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
#endif
#ifdef TQ_PHASE_TIMERS
    if (lane == 0) {
        for (int i = 0; i < 4; ++i) atomicAdd(&g_phase_clk[i], (unsigned long long)ph[i]);
        atomicAdd(&g_phase_clk[4], (unsigned long long)(clock64() - tq_start));
        atomicAdd(&g_phase_clk[5], 1ull);
    }
#endif
}

// ------------------------------------------------------------------------------------------
// Warp-specialised variant of the fused kernel.  A CTA has 12 warps: warps 0-7 only count
// (stage 1 + stage 2, counters in tensor memory, registers trimmed with setmaxnreg), warps 8-11
// only score (QR + Jacobi, registers raised with setmaxnreg).  Count warp c hands the 256 counts
// of a quartet to score warp 8 + c/2 through one of two shared-memory slots guarded by a
// full/empty mbarrier pair, so the POPC-bound and the FP64-latency-bound halves of a quartet run
// concurrently on different warps instead of back to back on one.  Still one launch; still only
// scores and the chosen split leave the SM.
// ------------------------------------------------------------------------------------------
static constexpr int kWsCount = 8, kWsScore = 4, kWsWarps = kWsCount + kWsScore;

struct __align__(16) WsSlotMeta { long long q; uint4 tips; };

struct __align__(16) WsShared {
    uint4 ring[kWsCount][kStages][4][32];
    uint32_t T[kWsCount][2][256];
    WsSlotMeta meta[kWsCount][2];
    SvdScratch svd[kWsScore];
    unsigned long long full[kWsCount][2], empty[kWsCount][2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n}\n"
                 ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// Wait for the phase with the given parity to complete.  try_wait suspends the warp in hardware for up to
// the hinted time (it resumes as soon as the phase completes), so a waiting warp costs a few instructions per
// millisecond instead of a spin loop that competes for issue slots with the warps it is waiting for
// (profiles/r01_v6_resolve_ncu.md).  A wait that lasts seconds means a protocol bug: trap (launch error)
// rather than hang the GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t tries = 0;; ++tries) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(1000000u) : "memory");
        if (ok) break;
        if ((tries & 1023u) == 1023u) {                                 // ~10 s at 2 GHz: a protocol bug, not a wait
            const long long now = clock64();
            if (t0 == 0) t0 = now; else if (now - t0 > 20000000000ll) __trap();
        }
    }
}

template <bool kSubsample>
__global__ void __launch_bounds__(kWsWarps * 32, 2)
resolve_ws_kernel(PlanDev plan, const uint4 *__restrict__ planes, uint32_t ntips,
                  const uint4 *__restrict__ quartets, int64_t m, uint64_t seed,
                  double *__restrict__ rscor, uint2 *__restrict__ rstat,
                  uint32_t *__restrict__ counts, uint32_t *__restrict__ minrank_out,
                  uint32_t *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WsShared &sh = *reinterpret_cast<WsShared *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int c = 0; c < kWsCount; ++c)
            for (int b = 0; b < 2; ++b) { mbar_init(&sh.full[c][b], 1); mbar_init(&sh.empty[c][b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        const uint32_t slot = (uint32_t)__cvta_generic_to_shared(&sh.tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(slot), "r"(2 * kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = sh.tmem_base;

    const int64_t ncount = (int64_t)gridDim.x * kWsCount;               // count warps in the grid
    auto quartets_of = [&](int c) -> int64_t {                           // how many quartets count warp c of this CTA resolves
        const int64_t first = (int64_t)blockIdx.x * kWsCount + c;
        return first < m ? (m - first + ncount - 1) / ncount : 0;
    };

    if (warp < kWsCount) {
        // ------------------------------ count warps ------------------------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(TQ_WS_DEC));
        const int c = warp;
        const uint32_t tacc = tmem_base + ((uint32_t)((c & 3) * 32) << 16) + (uint32_t)(c >> 2) * kTmemCols;
        zero_acc(tacc);
        const int64_t n = quartets_of(c);
        int64_t q = (int64_t)blockIdx.x * kWsCount + c;
        for (int64_t it = 0; it < n; ++it, q += ncount) {
            const int b = (int)(it & 1);
            mbar_wait(&sh.empty[c][b], (uint32_t)(((it >> 1) & 1) ^ 1));   // slot free (passes at once the first time)
            uint32_t *T = sh.T[c][b];
            const uint4 t = __ldg(quartets + q);
            const bool bad = t.x >= ntips || t.y >= ntips || t.z >= ntips || t.w >= ntips;
#pragma unroll
            for (int i = 0; i < 8; ++i) T[lane + 32 * i] = 0u;
            __syncwarp();
#ifdef TQ_PHASE_TIMERS
            long long ph_unused[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // the phase timers report the homogeneous kernel only
            if (!bad) count_quartet<kSubsample>(plan, planes, t, &sh.ring[c][0][0][0], T, lane, 0u, tacc, ph_unused);
#else
            if (!bad) count_quartet<kSubsample>(plan, planes, t, &sh.ring[c][0][0][0], T, lane, 0u, tacc);
#endif
            if (lane == 0) { sh.meta[c][b].q = bad ? -1 - q : q; sh.meta[c][b].tips = t; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.full[c][b]);
        }
        tmem_wait_st();
    } else {
        // ------------------------------ score warps ------------------------------
        // pool = 384 threads x 80 registers (ptxas, launch bounds (384, 2)); 256 x 72 + 128 x 96 uses it exactly.
        // Asking for more than the pool holds would wait forever.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(TQ_WS_INC));
        const int sidx = warp - kWsCount;
        SvdScratch &sv = sh.svd[sidx];
        const int c0 = 2 * sidx;
        const int64_t n0 = quartets_of(c0), n1 = quartets_of(c0 + 1);
        const int64_t nmax = n0 > n1 ? n0 : n1;
        for (int64_t it = 0; it < nmax; ++it) {
            const int b = (int)(it & 1);
            const uint32_t parity = (uint32_t)((it >> 1) & 1);
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                const int c = c0 + k;
                if (it >= (k ? n1 : n0)) continue;
                mbar_wait(&sh.full[c][b], parity);
                uint32_t *T = sh.T[c][b];
                const long long mq = sh.meta[c][b].q;
                const uint4 t = sh.meta[c][b].tips;
                if (mq < 0) {                                              // tip index out of range
                    const long long q = -1 - mq;
                    if (lane == 0) {
                        rstat[q] = make_uint2(0xFFFFFFFFu, 0u);
                        rscor[3 * q] = rscor[3 * q + 1] = rscor[3 * q + 2] = 0.0;
                        if (status) atomicOr(status, 1u);
                    }
                } else {
                    const long long q = mq;
                    if (!kSubsample && lane < 4) T[85u * lane] = 0u;      // invariant patterns aaaa, :218
                    __syncwarp();
                    uint32_t nsnps = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t v = T[lane + 32 * i];
                        nsnps += v;
                        if (counts) counts[q * 256 + lane + 32 * i] = v;
                    }
                    nsnps = __reduce_add_sync(kFull, nsnps);                // :226
                    double score[3] = {0.001, 0.001, 0.001};                // :232
                    uint32_t best, minrank = 0;
                    if (nsnps == 0) {
                        uint64_t h = seed;
                        h = splitmix64(h ^ t.x); h = splitmix64(h ^ t.y);
                        h = splitmix64(h ^ t.z); h = splitmix64(h ^ t.w);
                        best = (uint32_t)(h % 3ull);                        // stands in for :231
                    } else {
                        svd_scores(T, sv, lane, score, &best, &minrank, 0u);
                    }
                    if (lane == 0) {
                        rscor[3 * q] = score[0]; rscor[3 * q + 1] = score[1]; rscor[3 * q + 2] = score[2];
                        rstat[q] = make_uint2(best, nsnps);                 // :251, :264
                        if (minrank_out) minrank_out[q] = minrank;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.empty[c][b]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(2 * kTmemCols) : "memory");
}

// ------------------------------------------------------------------------------------------
// Stage 2 on the integer tensor cores.
//
//   T[ab][cd] = sum over the tile's 1024 loci of  X_ab[l] * Y_cd[l],   X_ab = A_a & B_b,  Y_cd = C_c & D_d
//
// is a 16 x 16 x 1024 product of 0/1 matrices whose rows ARE the merged bit planes.  sm_100a has no
// native 1-bit MMA (ptxas expands mma.b1 into IMMA + ~100 unpack instructions), but the legacy u8
// instruction mma.sync.m16n8k32 (SASS IMMA.16832.U8.U8) is native, and a bit plane becomes an u8
// operand with ONE LOP3 per four loci: word & (0x01010101 << v) holds the bytes {0, 2^v}.  Both
// operands use the same mask, so all products of one instruction are {0, 4^v}; v = 0..3 on the word
// and on the word >> 4 cover the eight bits of every byte, and each v has its own s32 accumulator
// fragment: count = sum_v acc_v >> 2v.  The K order of an MMA is free as long as A and B agree, so lane
// (g = lane / 4, j = lane % 4) simply feeds bytes j of ITS OWN word pair -- the data is never
// transposed: the lane reads planes A_p, A_p+2, B_r (rows g and g + 8 of the A operand, p = g / 4,
// r = g % 4) and C_p, C_p+2, D_r (column g of the two B operands) of words 8 j .. 8 j + 7 from a
// per-warp shared-memory copy of the tile's 16 merged planes (row stride 34 words: conflict-free
// 8-byte reads).  Per tile and warp: 16 STS + 24 LDS.64 + ~320 LOP3/SHF + 64 IMMA, against
// 256 LOP3 + 256 POPC + 128 IADD + the TMEM round trips of the POPC path; nothing runs on the XU pipe,
// the MMA sums over the whole tile (no cross-lane butterfly), and the counters are 32 registers
// (measured stand-alone: profiles/probes/imma_count_probe.cu, ~1000 clocks per tile and scheduler
// against >= 2048 for 256 POPCs).
// ------------------------------------------------------------------------------------------
#ifdef TQ_MMA_TIMERS                                  // profiling only: clock64 phase accumulators of the tensor-core kernel
__device__ unsigned long long g_mma_clk[8];         // count: slot wait, stage 2, total, warps; score: full wait, svd, total, warps
#define TQM_DECL(x) long long x = 0
#define TQM_T(x) const long long x = clock64()
#define TQM_ADD(acc, t0) acc += clock64() - (t0)
#else
#define TQM_DECL(x)
#define TQM_T(x)
#define TQM_ADD(acc, t0)
#endif
#ifndef TQ_MMA_SPLIT
#define TQ_MMA_SPLIT 0
#endif

#ifndef TQ_MMA_CPASYNC
#define TQ_MMA_CPASYNC "cp.async.cg.shared.global"     // .cg bypasses L1; .ca (rows of consecutive quartets shared through L1) measured 3 % slower
#endif
#ifndef TQ_MMA_UNROLL
#define TQ_MMA_UNROLL 1                       // unroll factor of the rank-chunk loop of stage 1
#endif
#ifndef TQ_MMA_TMA
#define TQ_MMA_TMA 0                          // 1: count warps fill their ring with TMA bulk copies (cp.async.bulk + mbarrier); measured
                                              // slower than cp.async at this depth (7.76 vs 9.81 M quartets/s at C2), kept for reference
#endif
#ifndef TQ_MMA_STAGES
#define TQ_MMA_STAGES 4                       // cp.async ring depth of the tensor-core kernel's count warps
#endif
static constexpr int kMStages = TQ_MMA_STAGES;
static constexpr int kMUnroll = TQ_MMA_UNROLL;
#ifndef TQ_MMA_SCALES
#define TQ_MMA_SCALES 4                       // 4: masks 0x01010101 << v on the word and on the word >> 4 (32 accumulator registers);
                                              // 8: masks << 0..7 straight from the planes, no shift (64 accumulator registers)
#endif
static constexpr int kScales = TQ_MMA_SCALES;
static constexpr int kPStride = 34;                 // words per plane row of the per-warp tile copy
static constexpr uint32_t kMmaFlushEvery = kScales == 8 ? 512 : 16384;   // tile products between accumulator flushes: 16384 * 1024 * 4^3
                                                    // (512 * 128 * 4^7 with eight scales) < 2^31
                                                    // (env TQ_MMA_FLUSH lowers it so that tests reach the flush path)

__device__ __forceinline__ void imma_u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// acc[v][h][i]: scale 4^v, column half h (cd = 8h ..), fragment register i
struct MmaLanePtrs {              // per-lane views of the warp's tile copy P, set up once per quartet
    uint32_t *st;                 // P + lane: where this lane stores its word of every plane
    const uint32_t *pa;           // A_p words 8j..: A_p+2 two rows on, C_p eight rows on, C_p+2 ten rows on
    const uint32_t *pb;           // B_r words 8j..: D_r eight rows on
};
__device__ __forceinline__ MmaLanePtrs mma_lane_ptrs(uint32_t *P, int lane)
{
    const int g = lane >> 2, j = lane & 3;
    MmaLanePtrs lp;
    lp.st = P + lane;
    lp.pa = P + (g >> 2) * kPStride + 8 * j;
    lp.pb = P + (4 + (g & 3)) * kPStride + 8 * j;
    return lp;
}

__device__ __forceinline__ void count_tile_mma(const uint32_t (&MA)[4], const uint32_t (&MB)[4],
                                               const uint32_t (&MC)[4], const uint32_t (&MD)[4],
                                               const MmaLanePtrs &lp, int (&acc)[kScales][2][4])
{
    __syncwarp();                                    // the previous tile's reads are over
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        lp.st[(0 + k) * kPStride] = MA[k];
        lp.st[(4 + k) * kPStride] = MB[k];
        lp.st[(8 + k) * kPStride] = MC[k];
        lp.st[(12 + k) * kPStride] = MD[k];
    }
    __syncwarp();
    const uint32_t *pa = lp.pa, *pb = lp.pb;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint2 a0 = *reinterpret_cast<const uint2 *>(pa + 2 * i);
        const uint2 a1 = *reinterpret_cast<const uint2 *>(pa + 2 * kPStride + 2 * i);
        const uint2 b = *reinterpret_cast<const uint2 *>(pb + 2 * i);
        const uint2 c0 = *reinterpret_cast<const uint2 *>(pa + 8 * kPStride + 2 * i);
        const uint2 c1 = *reinterpret_cast<const uint2 *>(pa + 10 * kPStride + 2 * i);
        const uint2 d = *reinterpret_cast<const uint2 *>(pb + 8 * kPStride + 2 * i);
#if TQ_MMA_SCALES == 8
#pragma unroll
        for (int v = 0; v < 8; ++v) {                 // three-input LOP3 straight from the planes: a & b & mask
            const uint32_t msk = 0x01010101u << v;
            const uint32_t f0 = a0.x & b.x & msk, f1 = a1.x & b.x & msk, f2 = a0.y & b.y & msk, f3 = a1.y & b.y & msk;
            imma_u8(acc[v][0], f0, f1, f2, f3, c0.x & d.x & msk, c0.y & d.y & msk);
            imma_u8(acc[v][1], f0, f1, f2, f3, c1.x & d.x & msk, c1.y & d.y & msk);
        }
#else
        uint32_t x[2][4], y[2][4];    // x: rows g / g+8 of word W, then of word W';  y: column g of half 0 (W, W'), of half 1 (W, W')
        x[0][0] = a0.x & b.x; x[0][1] = a1.x & b.x; x[0][2] = a0.y & b.y; x[0][3] = a1.y & b.y;
        y[0][0] = c0.x & d.x; y[0][1] = c0.y & d.y; y[0][2] = c1.x & d.x; y[0][3] = c1.y & d.y;
#pragma unroll
        for (int q = 0; q < 4; ++q) { x[1][q] = x[0][q] >> 4; y[1][q] = y[0][q] >> 4; }
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint32_t msk = 0x01010101u << v;
                const uint32_t f0 = x[s][0] & msk, f1 = x[s][1] & msk, f2 = x[s][2] & msk, f3 = x[s][3] & msk;
                imma_u8(acc[v][0], f0, f1, f2, f3, y[s][0] & msk, y[s][1] & msk);
                imma_u8(acc[v][1], f0, f1, f2, f3, y[s][2] & msk, y[s][3] & msk);
            }
#endif
    }
}

// Add the accumulated counts into T and clear the accumulators.  Lane (g, j) holds rows g and g + 8,
// columns 2j and 2j + 1 of each column half: no cross-lane step.
__device__ __forceinline__ void flush_mma(int (&acc)[kScales][2][4], uint32_t *T, int lane)
{
    const int g = lane >> 2, j = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            uint32_t c0 = 0u, c1 = 0u;
#pragma unroll
            for (int v = 0; v < kScales; ++v) {
                c0 += (uint32_t)acc[v][h][2 * rr] >> (2 * v);
                c1 += (uint32_t)acc[v][h][2 * rr + 1] >> (2 * v);
                acc[v][h][2 * rr] = 0; acc[v][h][2 * rr + 1] = 0;
            }
            uint2 *t = reinterpret_cast<uint2 *>(T + (g + 8 * rr) * 16 + 8 * h + 2 * j);
            uint2 cur = *t;
            cur.x += c0; cur.y += c1;
            *t = cur;
        }
}

// Stage 1 as in count_quartet (cp.async ring, first usable column per locus), stage 2 by count_tile_mma.
// Differences that matter at two count warps per scheduler (profiles/r01_v6_resolve_ncu.md): the tile's rank
// count is fetched one tile ahead (it was a global-load latency on the critical path of every tile), the
// ring reads are issued before the next chunk's address arithmetic and cp.async so that their latency is
// covered, and the global addresses are four 64-bit row pointers plus one running 32-bit offset.
template <bool kSubsample>
__device__ __forceinline__ void count_quartet_mma(const PlanDev &plan, const uint4 *__restrict__ planes,
                                                  uint4 tips, const uint4 *ring, unsigned long long *ring_full, uint32_t &ring_pos,
                                                  uint32_t *P, uint32_t *T, int lane, uint32_t flush_every
#ifdef TQ_MMA_TIMERS
                                                  , long long &clk_stage2
#endif
                                                  )
{
    static_assert((kMStages & (kMStages - 1)) == 0, "ring depth must be a power of two");
    int acc[kScales][2][4];
#pragma unroll
    for (int v = 0; v < kScales; ++v)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[v][h][i] = 0;
    const MmaLanePtrs lp = mma_lane_ptrs(P, lane);
    uint32_t pending = 0;
    const uint32_t nchunks = plan.row_u4 / kTileWords;
#if TQ_MMA_TMA
    // TMA bulk copies: lanes 0..3 each own one tip row and copy its 512-byte rank chunk with one cp.async.bulk that
    // completes on the stage's mbarrier; lane 0 arms the barrier with the stage's 2 KB.  The other 28 lanes issue
    // nothing and compute no addresses.  `ring_pos` runs on across quartets (stage = pos % depth, parity = pos / depth).
    const uint32_t my_tip = (lane & 3) == 0 ? tips.x : (lane & 3) == 1 ? tips.y : (lane & 3) == 2 ? tips.z : tips.w;
    const uint4 *grow = planes + (size_t)my_tip * plan.row_u4;
    const uint32_t ring_row_s = (uint32_t)__cvta_generic_to_shared(ring) + (uint32_t)(lane & 3) * 512u;
    const uint32_t full_s = (uint32_t)__cvta_generic_to_shared(ring_full);
    const uint32_t pos0 = ring_pos;
    auto issue = [&](uint32_t chunk) {
        const uint32_t stage = (pos0 + chunk) & (kMStages - 1);
        const uint32_t bar = full_s + stage * 8u;
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(2048u) : "memory");
        if (lane < 4)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(ring_row_s + stage * 2048u), "l"(grow + (size_t)chunk * kTileWords), "r"(512u), "r"(bar) : "memory");
    };
#ifndef TQ_DBG_NO_LOAD
    __syncwarp();                                     // the previous quartet's ring reads are over
#pragma unroll
    for (int s = 0; s < kMStages - 1; ++s)
        if ((uint32_t)s < nchunks) issue(s);
#endif
    ring_pos = pos0 + nchunks;
#else
    const uint4 *g0 = planes + (size_t)tips.x * plan.row_u4 + lane, *g1 = planes + (size_t)tips.y * plan.row_u4 + lane;
    const uint4 *g2 = planes + (size_t)tips.z * plan.row_u4 + lane, *g3 = planes + (size_t)tips.w * plan.row_u4 + lane;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring + lane);
    const uint32_t pos0 = 0;
    auto issue = [&](uint32_t chunk) {
        const uint32_t s = ring_s + (chunk & (kMStages - 1)) * (4 * 32 * 16), off = chunk * kTileWords;
        asm volatile(TQ_MMA_CPASYNC " [%0], [%1], 16;\n" ::"r"(s), "l"(g0 + off) : "memory");
        asm volatile(TQ_MMA_CPASYNC " [%0], [%1], 16;\n" ::"r"(s + 512), "l"(g1 + off) : "memory");
        asm volatile(TQ_MMA_CPASYNC " [%0], [%1], 16;\n" ::"r"(s + 1024), "l"(g2 + off) : "memory");
        asm volatile(TQ_MMA_CPASYNC " [%0], [%1], 16;\n" ::"r"(s + 1536), "l"(g3 + off) : "memory");
    };
#ifndef TQ_DBG_NO_LOAD
#pragma unroll
    for (int s = 0; s < kMStages - 1; ++s) {
        if ((uint32_t)s < nchunks) issue(s);
        cp_async_commit();
    }
#endif
#endif
    uint32_t u = 0;                                   // rank-chunk index along the rows
#ifdef TQ_COUNT_TILES                                 // profiling only: count the first tiles, to see the score warps' ceiling
    const int ntiles = plan.ntiles < TQ_COUNT_TILES ? plan.ntiles : TQ_COUNT_TILES;
#else
    const int ntiles = plan.ntiles;
#endif
    uint32_t Knext = ntiles > 0 ? __ldg(plan.tile_k) : 0u;
    for (int tile = 0; tile < ntiles; ++tile) {
        const uint32_t K = Knext;
        if (tile + 1 < ntiles) Knext = __ldg(plan.tile_k + tile + 1);
        uint32_t taken = 0u;
        uint32_t MA[4] = {0u, 0u, 0u, 0u}, MB[4] = {0u, 0u, 0u, 0u};
        uint32_t MC[4] = {0u, 0u, 0u, 0u}, MD[4] = {0u, 0u, 0u, 0u};
#pragma unroll kMUnroll
        for (uint32_t k = 0; k < K; ++k, ++u) {
#ifndef TQ_DBG_NO_LOAD
#if TQ_MMA_TMA
            mbar_wait(ring_full + ((pos0 + u) & (kMStages - 1)), ((pos0 + u) / kMStages) & 1u);   // chunk u has landed
#else
            cp_async_wait<kMStages - 2>();            // chunk u has landed (chunks u+1 .. u+kMStages-2 may be in flight)
#endif
#endif
            const uint4 *st = ring + ((pos0 + u) & (kMStages - 1)) * (4 * 32) + lane;
            const uint4 a = st[0], b = st[32], c = st[64], d = st[96];
#ifndef TQ_DBG_NO_LOAD
            {   // refill the stage read in the previous iteration; keeps kMStages-1 chunks in flight
                const uint32_t nu = u + kMStages - 1;
#if TQ_MMA_TMA
                __syncwarp();                         // every lane's reads of that stage have returned
                if (nu < nchunks) issue(nu);
#else
                if (nu < nchunks) issue(nu);
                cp_async_commit();
#endif
            }
#endif
#ifdef TQ_DBG_NO_S1
            if (kSubsample) { MA[0] |= a.x; MB[0] |= b.y; MC[0] |= c.z; MD[0] |= d.w; } else
#endif
            if (kSubsample) {
                const uint32_t valid = (a.x | a.y | a.z | a.w) & (b.x | b.y | b.z | b.w) &
                                       (c.x | c.y | c.z | c.w) & (d.x | d.y | d.z | d.w);
                const uint32_t same = (a.x & b.x & c.x & d.x) | (a.y & b.y & c.y & d.y) |
                                      (a.z & b.z & c.z & d.z) | (a.w & b.w & c.w & d.w);
                const uint32_t sel = valid & ~same & ~taken;
                taken |= sel;
                // a selected column is complete, so the T plane of every tip is what the other three leave of `taken`:
                // three merged planes per tip are updated per chunk, the fourth is derived once per tile
                MA[0] |= a.x & sel; MA[1] |= a.y & sel; MA[2] |= a.z & sel;
                MB[0] |= b.x & sel; MB[1] |= b.y & sel; MB[2] |= b.z & sel;
                MC[0] |= c.x & sel; MC[1] |= c.y & sel; MC[2] |= c.z & sel;
                MD[0] |= d.x & sel; MD[1] |= d.y & sel; MD[2] |= d.z & sel;
            } else {
                // full mode (resolve_quartets.py:92-95): every complete column counts; the four
                // invariant patterns aaaa are zeroed by the caller
                const uint32_t A[4] = {a.x, a.y, a.z, a.w}, B[4] = {b.x, b.y, b.z, b.w};
                const uint32_t C[4] = {c.x, c.y, c.z, c.w}, D[4] = {d.x, d.y, d.z, d.w};
                count_tile_mma(A, B, C, D, lp, acc);
                if (++pending == flush_every) { flush_mma(acc, T, lane); pending = 0; }
            }
        }
        if (kSubsample) {
#ifndef TQ_DBG_NO_S1
            MA[3] = taken & ~(MA[0] | MA[1] | MA[2]); MB[3] = taken & ~(MB[0] | MB[1] | MB[2]);
            MC[3] = taken & ~(MC[0] | MC[1] | MC[2]); MD[3] = taken & ~(MD[0] | MD[1] | MD[2]);
#endif
            TQM_T(ts2);
#ifndef TQ_DBG_NO_S2
            count_tile_mma(MA, MB, MC, MD, lp, acc);
#else
            acc[0][0][0] += MA[0] ^ MA[1] ^ MA[2] ^ MA[3] ^ MB[0] ^ MB[1] ^ MB[2] ^ MB[3] ^ MC[0] ^ MC[1] ^ MC[2] ^ MC[3] ^ MD[0] ^ MD[1] ^ MD[2] ^ MD[3];
#endif
            TQM_ADD(clk_stage2, ts2);
            if (++pending == flush_every) { flush_mma(acc, T, lane); pending = 0; }
        }
    }
#if !TQ_MMA_TMA
    cp_async_wait<0>();
#endif
    flush_mma(acc, T, lane);
}

// ------------------------------------------------------------------------------------------
// Warp-specialised kernel around the tensor-core count path.  A CTA has NC count warps and NS score
// warps.  The count path has no XU bottleneck and plenty of independent work per warp, so few
// count warps keep the SM's integer pipes busy; the FP64 Jacobi is bound by its own dependency chains
// (a score warp issues about one instruction in ten cycles), so most of the register file goes to score
// warps.  Hand-off through a CTA-wide ring of NC + NS slots of 256 counts: producers and consumers
// take tickets (one shared-memory atomic each), slot = ticket % slots, full / empty mbarrier per
// slot, so ANY score warp takes the next finished quartet -- no fixed count -> score pairing.
// ------------------------------------------------------------------------------------------
template <int NC, int NS>
struct __align__(16) MmaShared {
    static constexpr int kSlots = NC + 4 * NS;      // every count warp fills one slot, every score warp holds up to four
    uint4 ring[NC][kMStages][4][32];
    unsigned long long ring_full[NC][kMStages];     // TMA completion barrier of every ring stage (TQ_MMA_TMA)
    uint32_t P[NC][16 * kPStride];
    uint32_t T[kSlots][256];
    WsSlotMeta meta[kSlots];
    SvdScratch4 svd[NS];
    unsigned long long full[kSlots], empty[kSlots];
    uint32_t produce, consume;
};

template <int NC, int NS>
constexpr int mma_launch_warps() { return NC + NS; }

template <bool kSubsample, int NC, int NS, int kCtas>
__global__ void __launch_bounds__(mma_launch_warps<NC, NS>() * 32, kCtas)
resolve_mma_kernel(PlanDev plan, const uint4 *__restrict__ planes, uint32_t ntips,
                   const uint4 *__restrict__ quartets, int64_t m, uint64_t seed,
                   double *__restrict__ rscor, uint2 *__restrict__ rstat,
                   uint32_t *__restrict__ counts, uint32_t *__restrict__ minrank_out,
                   uint32_t *__restrict__ status, uint32_t flush_every)
{
    using Shared = MmaShared<NC, NS>;
    constexpr int kSlots = Shared::kSlots;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared &sh = *reinterpret_cast<Shared *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1); }
        for (int c = 0; c < NC; ++c)
            for (int s = 0; s < kMStages; ++s) mbar_init(&sh.ring_full[c][s], 1);
        sh.produce = 0u; sh.consume = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int64_t ncount = (int64_t)gridDim.x * NC;                     // count warps in the grid
    auto quartets_of = [&](int c) -> int64_t {                           // how many quartets count warp c of this CTA resolves
        const int64_t first = (int64_t)blockIdx.x * NC + c;
        return first < m ? (m - first + ncount - 1) / ncount : 0;
    };

    // Role of a warp.  Warp w runs on scheduler w % 4, and a score warp that shares a scheduler with count warps is starved
    // by them (score 190 k clocks per quartet against 75 k on a scheduler of its own: the count warps are nearly always
    // ready to issue).  TQ_MMA_SPLIT 0: warps 0..NC-1 count, the rest score, every scheduler gets both roles.
    // 1: schedulers 0-1 only count, 2-3 only score (NC == NS, both even).  2: schedulers 0-2 only count (NS warps each),
    // scheduler 3 only scores (NS warps).  (setmaxnreg works on groups of four consecutive warps, i.e. one warp of every
    // scheduler, so idle filler warps cannot hand their registers to the others: the mixes are limited to these.)
#ifdef TQ_MMA_ROLEMASK                               // bit w set: warp w counts (profiling: any placement of NC count warps)
    static_assert(__builtin_popcount(TQ_MMA_ROLEMASK) == NC, "TQ_MMA_ROLEMASK must have NC bits set");
    const bool is_count = (TQ_MMA_ROLEMASK >> warp) & 1u;
    const uint32_t below = (1u << warp) - 1u;
    const int role_idx = is_count ? __popc(TQ_MMA_ROLEMASK & below) : __popc(~(uint32_t)TQ_MMA_ROLEMASK & below);
#elif TQ_MMA_SPLIT == 2
    static_assert(NC == 3 * NS, "TQ_MMA_SPLIT=2: three count warps and one score warp in every group of four warps");
    const bool is_count = (warp & 3) < 3;
    const int role_idx = is_count ? (warp >> 2) * 3 + (warp & 3) : (warp >> 2);
#elif TQ_MMA_SPLIT == 1
    static_assert(NC == NS && NC % 2 == 0, "TQ_MMA_SPLIT=1 needs as many count as score warps");
    const bool is_count = (warp & 3) < 2;
    const int role_idx = (warp >> 2) * 2 + (warp & 1);
#else
    const bool is_count = warp < NC;
    const int role_idx = is_count ? warp : warp - NC;
#endif
    if (is_count) {
        // ------------------------------ count warps ------------------------------
        const int c = role_idx;
        const int64_t n = quartets_of(c);
        int64_t q = (int64_t)blockIdx.x * NC + c;
        TQM_DECL(clk_wait); TQM_DECL(clk_s2); TQM_T(t_begin);
        uint32_t ring_pos = 0;
        for (int64_t it = 0; it < n; ++it, q += ncount) {
            uint32_t ticket = 0;
            if (lane == 0) ticket = atomicAdd(&sh.produce, 1u);
            ticket = __shfl_sync(kFull, ticket, 0);
            const uint32_t slot = ticket % kSlots, use = ticket / kSlots;
            TQM_T(tw);
            mbar_wait(&sh.empty[slot], (use & 1u) ^ 1u);                  // slot free (passes at once on its first use)
            TQM_ADD(clk_wait, tw);
            uint32_t *T = sh.T[slot];
            const uint4 t = __ldg(quartets + q);
            const bool bad = t.x >= ntips || t.y >= ntips || t.z >= ntips || t.w >= ntips;
#pragma unroll
            for (int i = 0; i < 8; ++i) T[lane + 32 * i] = 0u;
            __syncwarp();
#ifdef TQ_MMA_TIMERS
            if (!bad) count_quartet_mma<kSubsample>(plan, planes, t, &sh.ring[c][0][0][0], sh.ring_full[c], ring_pos, sh.P[c], T, lane, flush_every, clk_s2);
#else
            if (!bad) count_quartet_mma<kSubsample>(plan, planes, t, &sh.ring[c][0][0][0], sh.ring_full[c], ring_pos, sh.P[c], T, lane, flush_every);
#endif
            if (lane == 0) { sh.meta[slot].q = bad ? -1 - q : q; sh.meta[slot].tips = t; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.full[slot]);
        }
#ifdef TQ_MMA_TIMERS
        if (lane == 0) {
            atomicAdd(&g_mma_clk[0], (unsigned long long)clk_wait); atomicAdd(&g_mma_clk[1], (unsigned long long)clk_s2);
            atomicAdd(&g_mma_clk[2], (unsigned long long)(clock64() - t_begin)); atomicAdd(&g_mma_clk[3], 1ull);
        }
#endif
    } else {
        // ------------------------------ score warps ------------------------------
        // A score warp takes up to four finished quartets at a time (svd_scores4: 12 matrices in three passes of four).
        SvdScratch4 &sv = sh.svd[role_idx];
        int64_t total = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) total += quartets_of(c);
        TQM_DECL(clk_wait); TQM_DECL(clk_svd); TQM_T(t_begin);
        while (true) {
            uint32_t ticket0 = 0;
            if (lane == 0) ticket0 = atomicAdd(&sh.consume, 4u);
            ticket0 = __shfl_sync(kFull, ticket0, 0);
            if ((int64_t)ticket0 >= total) break;
            const int nq = (int)((total - (int64_t)ticket0) < 4 ? (total - (int64_t)ticket0) : 4);
            uint32_t slots = 0u, actmask = 0u, my_nsnps = 0u, my_best = 0u;
            long long my_q = -1;                                           // lane j < 4: quartet j of this batch
            TQM_T(tw);
#pragma unroll 1
            for (int j = 0; j < nq; ++j) {
                const uint32_t ticket = ticket0 + (uint32_t)j;
                const uint32_t slot = ticket % kSlots, use = ticket / kSlots;
                mbar_wait(&sh.full[slot], use & 1u);
                slots |= slot << (8 * j);
                uint32_t *T = sh.T[slot];
                const long long mq = sh.meta[slot].q;
                const uint4 t = sh.meta[slot].tips;
                if (mq < 0) {                                              // tip index out of range
                    const long long q = -1 - mq;
                    if (lane == 0) {
                        rstat[q] = make_uint2(0xFFFFFFFFu, 0u);
                        rscor[3 * q] = rscor[3 * q + 1] = rscor[3 * q + 2] = 0.0;
                        if (status) atomicOr(status, 1u);
                    }
                    continue;
                }
                if (!kSubsample && lane < 4) T[85u * lane] = 0u;          // invariant patterns aaaa, :218
                __syncwarp();
                uint32_t nsnps = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t v = T[lane + 32 * i];
                    nsnps += v;
                    if (counts) counts[mq * 256 + lane + 32 * i] = v;
                }
                nsnps = __reduce_add_sync(kFull, nsnps);                    // :226
                uint32_t best = 0;
                if (nsnps == 0) {
                    uint64_t h = seed;
                    h = splitmix64(h ^ t.x); h = splitmix64(h ^ t.y);
                    h = splitmix64(h ^ t.z); h = splitmix64(h ^ t.w);
                    best = (uint32_t)(h % 3ull);                            // stands in for :231
                } else {
                    actmask |= 1u << j;
                }
                if (lane == j) { my_q = mq; my_nsnps = nsnps; my_best = best; }
            }
            TQM_ADD(clk_wait, tw);
            TQM_T(tsv);
            static_assert(kSlots <= 256, "slot indices travel in one byte each");
            double score[3] = {0.001, 0.001, 0.001};                        // :232
            uint32_t minrank = 0;
#ifndef TQ_NO_SVD
            if (actmask) svd_scores4(&sh.T[0][0], slots, actmask, sv, lane, score, my_best, minrank);
#endif
            if (lane < 4 && my_q >= 0) {
                rscor[3 * my_q] = score[0]; rscor[3 * my_q + 1] = score[1]; rscor[3 * my_q + 2] = score[2];
                rstat[my_q] = make_uint2(my_best, my_nsnps);                // :251, :264
                if (minrank_out) minrank_out[my_q] = minrank;
            }
            __syncwarp();
            TQM_ADD(clk_svd, tsv);
            if (lane == 0)
                for (int j = 0; j < nq; ++j) mbar_arrive(&sh.empty[(ticket0 + (uint32_t)j) % kSlots]);
        }
#ifdef TQ_MMA_TIMERS
        if (lane == 0) {
            atomicAdd(&g_mma_clk[4], (unsigned long long)clk_wait); atomicAdd(&g_mma_clk[5], (unsigned long long)clk_svd);
            atomicAdd(&g_mma_clk[6], (unsigned long long)(clock64() - t_begin)); atomicAdd(&g_mma_clk[7], 1ull);
        }
#endif
    }
}

// cudaFuncSetAttribute is per device: remember which devices of this process have the attribute (one bit each).
static bool attribute_needed(std::atomic<uint64_t> &done)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    return (done.fetch_or(bit) & bit) == 0;
}

template <bool kSubsample>
static cudaError_t launch_resolve_mma(const PlanDev &plan, const uint4 *planes, uint32_t ntips, const uint4 *quartets,
                                      int64_t m, uint64_t seed, double *d_rscor, uint2 *rstat, uint32_t *d_counts,
                                      uint32_t *d_minrank, uint32_t *d_status, int num_sms, cudaStream_t stream)
{
    constexpr int NC = TQ_MMA_NC, NS = TQ_MMA_NS, kCtas = TQ_MMA_CTAS;
    auto kern = resolve_mma_kernel<kSubsample, NC, NS, kCtas>;
    const size_t smem = sizeof(MmaShared<NC, NS>);
    static std::atomic<uint64_t> attr{0};
    if (attribute_needed(attr)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { attr = 0; return e; }
    }
    const int64_t want = (m + NC - 1) / NC;
    const int64_t cap = (int64_t)num_sms * kCtas;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    const int64_t fv = g_mma_flush.load();
    const uint32_t flush_every = (fv >= 1 && fv < (int64_t)kMmaFlushEvery) ? (uint32_t)fv : kMmaFlushEvery;
    kern<<<grid, mma_launch_warps<NC, NS>() * 32, smem, stream>>>(plan, planes, ntips, quartets, m, seed, d_rscor, rstat, d_counts, d_minrank, d_status,
                                                 flush_every);
    ++g_launches;
    g_last_kernel = kSubsample ? "tq::resolve_mma_kernel<true>" : "tq::resolve_mma_kernel<false>";
#ifdef TQ_MMA_TIMERS
    {
        cudaDeviceSynchronize();
        unsigned long long h[8];
        cudaMemcpyFromSymbol(h, g_mma_clk, sizeof(h));
        const double cw = (double)h[3], sw = (double)h[7], qn = (double)m;
        fprintf(stderr, "[mma clk] %d+%d warps x %d CTAs/SM, %.0f count warps, %.0f score warps, %.1f quartets per count warp\n"
                        "  count warp, per quartet: slot wait %.0f, stage 2 %.0f, rest (stage 1 + flush) %.0f, total %.0f\n"
                        "  score warp, per quartet: wait for counts %.0f, score %.0f; per warp total %.0f of which waiting %.0f\n",
                NC, NS, kCtas, cw, sw, qn / cw, h[0] / qn, h[1] / qn, (h[2] - h[0] - h[1]) / qn, h[2] / qn,
                h[4] / qn, h[5] / qn, h[6] / sw, h[4] / sw);
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_mma_clk, z, sizeof(z));
    }
#endif
    return cudaGetLastError();
}

cudaError_t launch_resolve(const PlanDev &plan, const uint32_t *d_planes, int64_t ntips,
                           const uint32_t *d_quartets, int64_t m, bool subsample, uint64_t seed,
                           double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts,
                           uint32_t *d_minrank, uint32_t *d_status, int num_sms,
                           cudaStream_t stream)
{
    if (m == 0) return cudaSuccess;
    static const uint32_t debug = getenv("TQ_DEBUG") ? (uint32_t)atoi(getenv("TQ_DEBUG")) : 0u;   // phase knock-outs for profiling only
    const int64_t want = (m + kWarps - 1) / kWarps;
    // quarter of the estimated per-quartet period of a warp (clocks): 4 warps share one POPC pipe
    // (256 POPC x ~8.5 clk per tile each) plus ~100k clocks of SVD; only worth it with >= 8 quartets/warp
    uint32_t stagger = 0;
    if (m >= (int64_t)num_sms * 16 * 8 && !(debug & 8u))
        stagger = (uint32_t)(((int64_t)plan.ntiles * 10200 + 120000) / 4);
    if (getenv("TQ_STAGGER")) stagger = (uint32_t)atoi(getenv("TQ_STAGGER"));
    const int64_t cap = (int64_t)num_sms * TQ_MIN_CTAS;   // resident CTAs per SM (launch bounds)
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    const uint4 *planes = reinterpret_cast<const uint4 *>(d_planes);
    const uint4 *quartets = reinterpret_cast<const uint4 *>(d_quartets);
    uint2 *rstat = reinterpret_cast<uint2 *>(d_rstat);
    // Three variants of the same device code.  On tiny launches (C1: 495 quartets) a count -> score hand-off is pure
    // latency, so the homogeneous kernel runs; from 16 quartets per SM the tensor-core kernel; the warp-specialised
    // POPC kernel only when forced.  tq_set_option("kernel", ..) / TQ_KERNEL=mma|ws|fused override.
    const int force = g_kernel_choice.load();
    // Size rule (profiles/tools/time_small.py, C2 matrix, M quartets/s fused / ws / mma): 1 184 quartets 4.5 / 3.7 / 4.1,
    // 2 368: 6.0 / 5.4 / 6.2, 4 736: 7.1 / 6.3 / 8.0, 37 888: 7.4 / 7.6 / 10.7 -- the tensor-core kernel from 16 quartets per SM.
    const bool use_mma = force ? (force == 3) : (TQ_DEFAULT_MMA != 0 && m >= (int64_t)num_sms * 16);
    if (use_mma) {
        return subsample ? launch_resolve_mma<true>(plan, planes, (uint32_t)ntips, quartets, m, seed, d_rscor, rstat, d_counts,
                                                    d_minrank, d_status, num_sms, stream)
                         : launch_resolve_mma<false>(plan, planes, (uint32_t)ntips, quartets, m, seed, d_rscor, rstat, d_counts,
                                                     d_minrank, d_status, num_sms, stream);
    }
    const bool use_ws = force ? (force == 2) : (TQ_DEFAULT_WS != 0 && m >= (int64_t)num_sms * 64);
    if (use_ws && TQ_TMEM) {
        const size_t wsmem = sizeof(WsShared);
        static std::atomic<uint64_t> ws_attr{0};
        if (attribute_needed(ws_attr)) {
            cudaFuncSetAttribute(resolve_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
            cudaFuncSetAttribute(resolve_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
        }
        const int64_t wwant = (m + kWsCount - 1) / kWsCount;
        const int64_t wcap = (int64_t)num_sms * 2;
        const unsigned wgrid = (unsigned)(wwant < wcap ? wwant : wcap);
        if (subsample)
            resolve_ws_kernel<true><<<wgrid, kWsWarps * 32, wsmem, stream>>>(plan, planes, (uint32_t)ntips, quartets, m, seed,
                                                                            d_rscor, rstat, d_counts, d_minrank, d_status);
        else
            resolve_ws_kernel<false><<<wgrid, kWsWarps * 32, wsmem, stream>>>(plan, planes, (uint32_t)ntips, quartets, m, seed,
                                                                             d_rscor, rstat, d_counts, d_minrank, d_status);
        ++g_launches;
        g_last_kernel = subsample ? "tq::resolve_ws_kernel<true>" : "tq::resolve_ws_kernel<false>";
        return cudaGetLastError();
    }
    g_last_kernel = subsample ? "tq::resolve_kernel<true>" : "tq::resolve_kernel<false>";
    const size_t smem = sizeof(WarpScratch) * kWarps;
    static std::atomic<uint64_t> attr_set{0};
    if (attribute_needed(attr_set)) {
        cudaFuncSetAttribute(resolve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(resolve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    if (subsample)
        resolve_kernel<true><<<grid, kWarps * 32, smem, stream>>>(plan, planes, (uint32_t)ntips, quartets,
                                                               m, seed, d_rscor, rstat, d_counts,
                                                               d_minrank, d_status, debug, (uint32_t)num_sms, stagger);
    else
        resolve_kernel<false><<<grid, kWarps * 32, smem, stream>>>(plan, planes, (uint32_t)ntips, quartets,
                                                                m, seed, d_rscor, rstat, d_counts,
                                                                d_minrank, d_status, debug, (uint32_t)num_sms, stagger);
    ++g_launches;
#ifdef TQ_PHASE_TIMERS
    {
        cudaDeviceSynchronize();
        unsigned long long h[8];
        cudaMemcpyFromSymbol(h, g_phase_clk, sizeof(h));
        const double w = (double)h[5];
        fprintf(stderr, "[phase clk per warp] stage1 %.0f stage2 %.0f flush %.0f svd+out %.0f total %.0f (warps %.0f)\n",
                h[0] / w, h[1] / w, h[2] / w, h[3] / w, h[4] / w, w);
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z));
    }
#endif
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// pipe micro-benchmarks (roofline denominators measured on the box)
// ------------------------------------------------------------------------------------------
template <int kKind>
__global__ void __launch_bounds__(256) pipe_kernel(uint32_t *out, int iters, uint32_t seed)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (kKind == 2) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = 1.0 + 1e-9 * (double)(tid + u);
        const double a = 1.0 + 1e-12 * (double)seed, b = 1e-15;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = fma(v[u], a, b);
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
        if (s == 123.456) out[0] = 1u;
    } else {
        uint32_t v[8], acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { v[u] = tid * 2654435761u + seed + u; acc[u] = u; }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (kKind == 0) acc[u] += __popc(v[u] ^ acc[u]);
                else acc[u] = (acc[u] & v[u]) | (acc[u] >> 1);      // LOP3 + SHF
            }
        uint32_t s = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += acc[u];
        if (s == 0x12345u) out[0] = s;
    }
}

// POPC mixed with other MIO-queue instructions: kMix 0 = POPC only, 1 = + 4 SHFL per 8 POPC,
// 2 = + 2 LDS.128 per 8 POPC, 3 = + 8 independent DFMA per 8 POPC, 4 = + 16 LOP3/IADD per 8 POPC
template <int kMix>
__global__ void __launch_bounds__(128) mix_kernel(uint32_t *out, int iters, uint32_t seed)
{
    __shared__ uint4 sm[128];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    sm[threadIdx.x] = make_uint4(tid, seed, tid ^ seed, 7u);
    __syncthreads();
    uint32_t v[8], acc[8], extra = seed;
    double d[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { v[u] = tid * 2654435761u + seed + u; acc[u] = u; d[u] = 1.0 + u; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += __popc(v[u] ^ acc[u]);
        if (kMix == 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] ^= __shfl_xor_sync(0xFFFFFFFFu, v[u + 4], 16);
        } else if (kMix == 2) {
            const uint4 a = sm[(threadIdx.x + it) & 127], b = sm[(threadIdx.x + 2 * it + 1) & 127];
            extra += a.x ^ a.w ^ b.y ^ b.z;
        } else if (kMix == 3) {
#pragma unroll
            for (int u = 0; u < 8; ++u) d[u] = fma(d[u], 1.0000001, 1e-9);
        } else if (kMix == 4) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { v[u] = (v[u] & 0x7FFFFFFFu) + (v[u] >> 3); extra ^= v[u] | acc[u]; }
        }
    }
    uint32_t s = extra;
    double ds = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) { s += acc[u] + v[u]; ds += d[u]; }
    if (s == 0x12345u || ds == 1.5) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                   size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// kind 4: issue rate of the legacy u8 tensor-core instruction (IMMA.16832.U8.U8), 6 independent accumulator
// fragments per warp, operands one bit per byte as in count_tile_mma
__global__ void __launch_bounds__(512) imma_rate_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (threadIdx.x * 2654435761u + seed * (i + 1)) & 0x04040404u;
    b[0] = (threadIdx.x * 40503u + seed) & 0x04040404u; b[1] = b[0] >> 1;
    int c[6][4];
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[k][i] = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 6; ++k) imma_u8(c[k], a[0], a[1], a[2], a[3], b[0], b[1]);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += (uint32_t)(c[k][0] + c[k][1] + c[k][2] + c[k][3]);
    if (s == 0x12345u) out[0] = s;
}

cudaError_t run_microbench(int kind, double *result)
{
    cudaError_t err;
    cudaEvent_t e0, e1;
    if ((err = cudaEventCreate(&e0)) != cudaSuccess) return err;
    if ((err = cudaEventCreate(&e1)) != cudaSuccess) return err;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float best_ms = 1e30f;
    double work = 0.0;
    if (kind >= 10) {   // POPC + other instruction mixes at the resolve kernel's occupancy (16 warps/SM)
        uint32_t *out = nullptr;
        if ((err = cudaMalloc(&out, 64)) != cudaSuccess) return err;
        const int iters = 4096, blocks = sms * 4, threads = 128;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            switch (kind) {
            case 10: mix_kernel<0><<<blocks, threads>>>(out, iters, rep); break;
            case 11: mix_kernel<1><<<blocks, threads>>>(out, iters, rep); break;
            case 12: mix_kernel<2><<<blocks, threads>>>(out, iters, rep); break;
            case 13: mix_kernel<3><<<blocks, threads>>>(out, iters, rep); break;
            default: mix_kernel<4><<<blocks, threads>>>(out, iters, rep); break;
            }
            ++g_launches;
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms) best_ms = ms;
        }
        cudaFree(out);
        work = (double)blocks * threads * (double)iters * 8.0;
    } else if (kind == 4) {
        uint32_t *out = nullptr;
        if ((err = cudaMalloc(&out, 64)) != cudaSuccess) return err;
        const int iters = 8192, blocks = sms, threads = 512;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            imma_rate_kernel<<<blocks, threads>>>(out, iters, rep);
            ++g_launches;
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms) best_ms = ms;
        }
        cudaFree(out);
        work = (double)blocks * (threads / 32) * (double)iters * 6.0;   // warp-level IMMA instructions
    } else if (kind == 3) {
        const size_t bytes = (size_t)1 << 30;
        uint4 *a = nullptr, *b = nullptr;
        if ((err = cudaMalloc(&a, bytes)) != cudaSuccess) return err;
        if ((err = cudaMalloc(&b, bytes)) != cudaSuccess) { cudaFree(a); return err; }
        cudaMemset(a, 1, bytes);
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            copy_kernel<<<sms * 16, 256>>>(a, b, bytes / 16);
            ++g_launches;
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms) best_ms = ms;
        }
        cudaFree(a); cudaFree(b);
        work = 2.0 * (double)bytes;
    } else {
        uint32_t *out = nullptr;
        if ((err = cudaMalloc(&out, 64)) != cudaSuccess) return err;
        const int iters = 4096, blocks = sms * 8, threads = 256;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            if (kind == 0) pipe_kernel<0><<<blocks, threads>>>(out, iters, rep);
            else if (kind == 1) pipe_kernel<1><<<blocks, threads>>>(out, iters, rep);
            else pipe_kernel<2><<<blocks, threads>>>(out, iters, rep);
            ++g_launches;
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms) best_ms = ms;
        }
        cudaFree(out);
        work = (double)blocks * threads * (double)iters * 8.0;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    *result = work / ((double)best_ms * 1e-3);
    return cudaSuccess;
}

}  // namespace tq
