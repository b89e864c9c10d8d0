// speck_b200/csrc/sort_numeric_impl.cuh -- numeric phase of the sort classes for one value type.
// u32 keys when col_bits + log2(N) <= 32, else u64 keys (wideKeys).
#pragma once
#include "sort_cta.cuh"

namespace sb {

template <typename T>
void launch_sort_numeric(const LaunchCtx &lc, int sortClass, bool wideKeys, const u32 *perm, u32 count,
                         const u32 *aRp, const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi,
                         const T *bV, const u32 *rowOps, const u32 *cRp, u32 *cCi, T *cV)
{
    if (count == 0) return;
    u32 *rp = const_cast<u32 *>(cRp);
#define SB_NUM(G, E)                                                                                         \
    do {                                                                                                     \
        if (wideKeys)                                                                                        \
            launch_sort_rows<G, E, u64, T, SORT_NUMERIC>(lc, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, rp, cCi, cV); \
        else                                                                                                 \
            launch_sort_rows<G, E, u32, T, SORT_NUMERIC>(lc, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, rp, cCi, cV); \
    } while (0)
#define SB_NUM_CTA(WMAX)                                                                                     \
    do {                                                                                                     \
        if (wideKeys)                                                                                        \
            launch_sort_rows_cta<WMAX, 16, u64, T, true>(lc, warps, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, rp, cCi, cV); \
        else                                                                                                 \
            launch_sort_rows_cta<WMAX, 16, u32, T, true>(lc, warps, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, rp, cCi, cV); \
    } while (0)
    switch (sortClass) {
        case 0: SB_NUM(4, 1); break;
        case 1: SB_NUM(8, 1); break;
        case 2: SB_NUM(16, 1); break;
        case 3: SB_NUM(32, 1); break;
        case 4: SB_NUM(32, 2); break;
        case 5: SB_NUM(32, 4); break;
        case 6: SB_NUM(32, 8); break;
        case 7: SB_NUM(32, 16); break;
        default: {
            const int warps = cta_class_warps(sortClass - NUM_WARP_SORT);
            if (warps <= 2) SB_NUM_CTA(2);
            else if (warps <= 4) SB_NUM_CTA(4);
            else if (warps <= 8) SB_NUM_CTA(8);
            else SB_NUM_CTA(16);
            break;
        }
    }
#undef SB_NUM
#undef SB_NUM_CTA
}

template <typename T>
void launch_map_numeric(const LaunchCtx &lc, int sortClass, const RowDesc *desc, u32 count, const uint2 *aSeg,
                        const T *aV, const u32 *bCi, const T *bV, const unsigned short *rankMap, u32 *cCi, T *cV)
{
    if (count == 0) return;
#define SB_MAP(G, N) launch_map_rows<G, N, T>(lc, desc, count, aSeg, aV, bCi, bV, rankMap, cCi, cV)
    switch (sortClass) {
        case 0: SB_MAP(4, 4); break;
        case 1: if (lc.narrow) SB_MAP(4, 8); else SB_MAP(8, 8); break;
        case 2: if (lc.narrow) SB_MAP(4, 16); else SB_MAP(16, 16); break;
        case 3: if (lc.narrow) SB_MAP(8, 32); else SB_MAP(32, 32); break;
        case 4: if (lc.narrow) SB_MAP(16, 64); else SB_MAP(32, 64); break;
        case 5: if (lc.narrow > 1) SB_MAP(16, 128); else SB_MAP(32, 128); break;
        case 6: if (lc.narrow > 2) SB_MAP(16, 256); else SB_MAP(32, 256); break;
        default: SB_MAP(32, 512); break;
    }
#undef SB_MAP
}

}  // namespace sb
