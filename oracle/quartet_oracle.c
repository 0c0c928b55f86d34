/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the quartet-resolution hot path.
 *
 * Plain-C restatement of the integer half of the reference's per-quartet loop
 * (eaton-lab/tetrad, reference paths relative to /root/reference):
 *
 *   tetrad/src/resolve_quartets.py:212      seqs = tmparr[sidx, :]            (row gather)
 *   tetrad/src/resolve_quartets.py:216      nmask0 = sum(seqs >= 78, axis=0)  (missing)
 *   tetrad/src/resolve_quartets.py:218      nmask1 = all four equal           (invariant)
 *   tetrad/src/resolve_quartets.py:42-73    subsample_chunk_to_matrices       (selection + count)
 *   tetrad/src/resolve_quartets.py:76-104   full_chunk_to_matrices            (count, no selection)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load this library; the product path (tetrad_b200/) never does.
 *
 * One deliberate narrowing: a cell whose code is neither 0..3 nor >= 78 (an unresolved
 * byte such as 'B'=66) makes the reference index its 16x16 matrix out of bounds
 * (undefined behaviour under numba).  Here, and in the CUDA path,
 * such a cell counts as missing.  Inputs produced by init_database / resample_tmp_database
 * from A/C/G/T/N/-/RKSYWM data never contain such cells.
 */
#include <stdint.h>
#include <string.h>

/* mats[3][16][16] for one quartet; sel (may be NULL) gets 1 for every counted column. */
static void count_one(const uint8_t *r0, const uint8_t *r1, const uint8_t *r2, const uint8_t *r3,
                      const uint32_t *locus, int64_t locus_stride, int64_t ncols, int subsample,
                      uint32_t *mats, uint8_t *sel)
{
    uint32_t *m0 = mats, *m1 = mats + 256, *m2 = mats + 512;
    memset(mats, 0, 3 * 256 * sizeof(uint32_t));
    if (sel) memset(sel, 0, (size_t)ncols);
    uint32_t last_loc = 0xFFFFFFFFu;              /* np.uint32(-1), resolve_quartets.py:57 */
    for (int64_t c = 0; c < ncols; c++) {
        uint8_t a = r0[c], b = r1[c], g = r2[c], d = r3[c];
        if (a > 3 || b > 3 || g > 3 || d > 3) continue;       /* :216 missing (see header) */
        if (a == b && a == g && a == d) continue;             /* :218 invariant */
        if (subsample) {
            uint32_t loc = locus[c * locus_stride];
            if (loc == last_loc) continue;                     /* :60 */
            last_loc = loc;                                    /* :64 */
        }
        m0[(4 * a + b) * 16 + (4 * g + d)] += 1;              /* :63 / :95 */
        if (sel) sel[c] = 1;
    }
    /* the two alternative flattenings, :67-72 / :98-103:
       row x of mats[0] (x = 4a+b), reshaped 4x4 over (g,d), becomes block (y,z) = (a,b)
       of mats[1] and its transpose becomes the same block of mats[2]. */
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++)
            for (int g = 0; g < 4; g++)
                for (int d = 0; d < 4; d++) {
                    uint32_t v = m0[(4 * a + b) * 16 + (4 * g + d)];
                    m1[(4 * a + g) * 16 + (4 * b + d)] = v;
                    m2[(4 * a + d) * 16 + (4 * b + g)] = v;
                }
}

/*
 * tmparr  u8[ntips][row_stride]   codes 0..3 / 78
 * locus   tmpmap[:,0] with element stride locus_stride (2 for the [S,2] array)
 * quartets u32[m][4]
 * mats    u32[m][3][16][16]
 * selmask u8[m][ncols] or NULL
 * returns 0, or -1 if a tip index is out of range.
 */
int tq_oracle_counts(const uint8_t *tmparr, int64_t ntips, int64_t ncols, int64_t row_stride,
                     const uint32_t *locus, int64_t locus_stride,
                     const uint32_t *quartets, int64_t m, int subsample,
                     uint32_t *mats, uint8_t *selmask)
{
    for (int64_t q = 0; q < m; q++) {
        const uint32_t *t = quartets + 4 * q;
        for (int i = 0; i < 4; i++)
            if ((int64_t)t[i] >= ntips) return -1;
        count_one(tmparr + (int64_t)t[0] * row_stride, tmparr + (int64_t)t[1] * row_stride,
                  tmparr + (int64_t)t[2] * row_stride, tmparr + (int64_t)t[3] * row_stride,
                  locus, locus_stride, ncols, subsample,
                  mats + q * 768, selmask ? selmask + q * ncols : (uint8_t *)0);
    }
    return 0;
}
