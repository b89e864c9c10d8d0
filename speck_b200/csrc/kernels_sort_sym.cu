// speck_b200/csrc/kernels_sort_sym.cu -- symbolic phase of the sort classes (u32 column keys).
#include "sort_cta.cuh"

namespace sb {

void launch_sort_symbolic(const LaunchCtx &lc, int sortClass, bool wideKeys, const u32 *perm, u32 count,
                          const u32 *aRp, const u32 *aCi, const u32 *bRp, const u32 *bCi, const u32 *rowOps,
                          u32 *rowNnz, const RowDesc *desc, const uint2 *aSeg, unsigned short *rankMap)
{
    if (count == 0) return;
    const float *nv = nullptr;
    const bool map = desc && aSeg && rankMap && sortClass < NUM_WARP_SORT;
#define SB_SYM(G, E)                                                                                                      \
    do {                                                                                                                  \
        if (!map)                                                                                                         \
            launch_sort_rows<G, E, u32, float, SORT_COUNT>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowNnz,   \
                                                           nullptr, nullptr);                                             \
        else if (wideKeys)                                                                                                \
            launch_sort_rows<G, E, u64, float, SORT_MAP>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowNnz,     \
                                                         nullptr, nullptr, desc, aSeg, rankMap);                          \
        else                                                                                                              \
            launch_sort_rows<G, E, u32, float, SORT_MAP>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowNnz,     \
                                                         nullptr, nullptr, desc, aSeg, rankMap);                          \
    } while (0)
    switch (sortClass) {
        // narrow: several rows per warp (4 / 8 / 16 lanes with 2..8 keys each): the per-row fixed work (descriptor,
        // scans, searches) is shared by the rows of a warp and more of the network runs inside a lane
        case 0: SB_SYM(4, 1); break;
        case 1: if (lc.narrow) SB_SYM(4, 2); else SB_SYM(8, 1); break;
        case 2: if (lc.narrow) SB_SYM(4, 4); else SB_SYM(16, 1); break;
        case 3: if (lc.narrow) SB_SYM(8, 4); else SB_SYM(32, 1); break;
        case 4: if (lc.narrow) SB_SYM(16, 4); else SB_SYM(32, 2); break;
        case 5: if (lc.narrow > 1) SB_SYM(16, 8); else SB_SYM(32, 4); break;
        case 6: if (lc.narrow > 2) SB_SYM(16, 16); else SB_SYM(32, 8); break;
        case 7: SB_SYM(32, 16); break;
        default: {
            const int warps = cta_class_warps(sortClass - NUM_WARP_SORT);
#define SB_SYM_CTA(WMAX) launch_sort_rows_cta<WMAX, 16, u32, float, false>(lc, warps, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowNnz, nullptr, nullptr)
            if (warps <= 2) SB_SYM_CTA(2);
            else if (warps <= 4) SB_SYM_CTA(4);
            else if (warps <= 8) SB_SYM_CTA(8);
            else SB_SYM_CTA(16);
#undef SB_SYM_CTA
            break;
        }
    }
#undef SB_SYM
}

}  // namespace sb
