// speck_b200/csrc/kernels_sort_sym.cu -- symbolic phase of the sort classes (u32 column keys).
#include "sort_cta.cuh"

namespace sb {

void launch_sort_symbolic(const LaunchCtx &lc, int sortClass, const u32 *perm, u32 count, const u32 *aRp,
                          const u32 *aCi, const u32 *bRp, const u32 *bCi, const u32 *rowOps, u32 *rowNnz)
{
    if (count == 0) return;
    const float *nv = nullptr;
#define SB_SYM(G, E) \
    launch_sort_rows<G, E, u32, float, false>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowNnz, nullptr, nullptr)
    switch (sortClass) {
        case 0: SB_SYM(4, 1); break;
        case 1: SB_SYM(8, 1); break;
        case 2: SB_SYM(16, 1); break;
        case 3: SB_SYM(32, 1); break;
        case 4: SB_SYM(32, 2); break;
        case 5: SB_SYM(32, 4); break;
        case 6: SB_SYM(32, 8); break;
        case 7: SB_SYM(32, 16); break;
        case 8: launch_sort_rows_cta<2, 16, u32, float, false>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowNnz, nullptr, nullptr); break;
        case 9: launch_sort_rows_cta<4, 16, u32, float, false>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowNnz, nullptr, nullptr); break;
        case 10: launch_sort_rows_cta<8, 16, u32, float, false>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowNnz, nullptr, nullptr); break;
        case 11: launch_sort_rows_cta<16, 16, u32, float, false>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowNnz, nullptr, nullptr); break;
        default: break;
    }
#undef SB_SYM
}

}  // namespace sb
