// speck_b200/csrc/kernels_rank.cu -- launchers of the rank row classes (rank_cta.cuh).
// Launch shapes: 128 / 256 / 512 / 1024 threads with 8 product slots per thread (4 and 16 slots per thread
// were measured slower on the same rows, profiles/r1_notes.md), and 1024 x 16 for rows of 8193..16384
// products.  `capProducts` is the largest product count among the rows of the launch.
#include "rank_cta.cuh"
#include "map_split.cuh"

namespace sb {

constexpr int RANK_E = 8;

#define SB_RANK_SHAPES8(CALL)                               \
    do {                                                    \
        if (capProducts <= 128 * RANK_E) CALL(128, RANK_E);       \
        else if (capProducts <= 256 * RANK_E) CALL(256, RANK_E);  \
        else if (capProducts <= 512 * RANK_E) CALL(512, RANK_E);  \
        else CALL(1024, RANK_E);                            \
    } while (0)
// ... plus the 16384-product shape, which exists for the mapped kernels only
#define SB_RANK_SHAPES(CALL)                                        \
    do {                                                            \
        if (capProducts > 1024 * RANK_E) CALL(1024, 2 * RANK_E);    \
        else SB_RANK_SHAPES8(CALL);                                 \
    } while (0)

void launch_rank_symbolic(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                          const u32 *aCi, const u32 *bRp, const u32 *bCi, const u32 *rowOps, const u32 *rowMin,
                          const u32 *rowMax, u32 *rowNnz, const RowDesc *desc, const uint2 *aSeg,
                          unsigned short *rankMap, int levels)
{
    if (count == 0) return;
    const float *nv = nullptr;
#define SB_RANK_CNT3(TH, E)                                                                                             \
    launch_rank_rows<TH, E, float, RANK_COUNT, 3>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowMin, rowMax,  \
                                                  nullptr, nullptr, nullptr, rowNnz, nullptr, nullptr)
#define SB_RANK_MAP3(TH, E)                                                                                             \
    launch_rank_rows<TH, E, float, RANK_MAP, 3>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowMin, rowMax,    \
                                                desc, aSeg, rankMap, rowNnz, nullptr, nullptr)
    if (levels == 3) {   // cols(B) in (2^20, 2^25]
        // lane-group classes of 256 / 512 products whose bitonic sort would need 64-bit keys (capi.cu)
        if (desc && aSeg && rankMap && capProducts <= 32 * RANK_E) SB_RANK_MAP3(32, RANK_E);
        else if (desc && aSeg && rankMap && capProducts <= 64 * RANK_E) SB_RANK_MAP3(64, RANK_E);
        else if (desc && aSeg && rankMap) SB_RANK_SHAPES(SB_RANK_MAP3);
        else if (capProducts <= SORT_MAX_PRODUCTS) SB_RANK_SHAPES8(SB_RANK_CNT3);
        return;
    }
#define SB_RANK_CNT(TH, E)                                                                                              \
    launch_rank_rows<TH, E, float, RANK_COUNT>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowMin, rowMax, \
                                                    nullptr, nullptr, nullptr, rowNnz, nullptr, nullptr)
#define SB_RANK_MAP(TH, E)                                                                                              \
    launch_rank_rows<TH, E, float, RANK_MAP>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowMin, rowMax,   \
                                                  desc, aSeg, rankMap, rowNnz, nullptr, nullptr)
    if (desc && aSeg && rankMap) SB_RANK_SHAPES(SB_RANK_MAP);
    else if (capProducts <= SORT_MAX_PRODUCTS) SB_RANK_SHAPES8(SB_RANK_CNT);   // larger rows: map only (capi.cu)
#undef SB_RANK_CNT
#undef SB_RANK_MAP
#undef SB_RANK_CNT3
#undef SB_RANK_MAP3
}

template <typename T>
void launch_rank_numeric(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                         const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi, const T *bV,
                         const u32 *rowOps, const u32 *rowMin, const u32 *rowMax, const u32 *cRp, u32 *cCi, T *cV)
{
    if (count == 0) return;
    u32 *rp = const_cast<u32 *>(cRp);
#define SB_RANK_NUM(TH, E)                                                                                               \
    launch_rank_rows<TH, E, T, RANK_NUMERIC>(lc, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, rowMin, rowMax,   \
                                                  nullptr, nullptr, nullptr, rp, cCi, cV)
    if (capProducts <= SORT_MAX_PRODUCTS) SB_RANK_SHAPES8(SB_RANK_NUM);   // larger rows: map only (capi.cu)
#undef SB_RANK_NUM
}

template <typename T>
void launch_map_numeric_cta(const LaunchCtx &lc, u32 capProducts, const RowDesc *desc, u32 count, const uint2 *aSeg,
                            const T *aV, const u32 *bCi, const T *bV, const unsigned short *rankMap, u32 *cCi,
                            T *cV, int colDirect, int bigSplit)
{
    if (count == 0) return;
#define SB_MAP_SPLIT(TH, E, S) launch_map_rows_split<TH, E, T, S>(lc, desc, count, aSeg, aV, bCi, bV, rankMap, cCi, cV)
    // rows of 4097 .. 16384 products: several CTAs per row, each staging one range of the row's positions, instead of
    // one 1024-thread CTA per SM (map_split.cuh).  1: 512 x 16 slots (two CTAs per SM), the 16384 class split in two;
    // 2: 512 x 8 slots (three CTAs per SM), split in two / four; 3: only the 16384 class (512 x 16, split in two)
    if (bigSplit && !colDirect && capProducts > 512 * RANK_E) {
        const bool top = capProducts > 1024 * RANK_E;
        if (bigSplit == 1 || (bigSplit == 3 && top)) {
            if (top) SB_MAP_SPLIT(512, 2 * RANK_E, 2);
            else launch_map_rows_cta<512, 2 * RANK_E, T>(lc, desc, count, aSeg, aV, bCi, bV, rankMap, cCi, cV);
            return;
        }
        if (bigSplit == 2) {
            if (top) SB_MAP_SPLIT(512, RANK_E, 4);
            else SB_MAP_SPLIT(512, RANK_E, 2);
            return;
        }
    }
#undef SB_MAP_SPLIT
#define SB_MAP_NUM(TH, E) launch_map_rows_cta<TH, E, T>(lc, desc, count, aSeg, aV, bCi, bV, rankMap, cCi, cV)
#define SB_MAP_NUM_CD(TH, E) launch_map_rows_cta<TH, E, T, true>(lc, desc, count, aSeg, aV, bCi, bV, rankMap, cCi, cV)
    if (capProducts <= 32 * RANK_E) SB_MAP_NUM(32, RANK_E);
    else if (capProducts <= 64 * RANK_E) SB_MAP_NUM(64, RANK_E);
    else if (colDirect && capProducts > 1024 * RANK_E) SB_MAP_NUM_CD(1024, 2 * RANK_E);   // 8193..16384 products
    else if (colDirect == 3 && capProducts > 512 * RANK_E) SB_MAP_NUM_CD(512, 2 * RANK_E);   // 4097..8192 as 512 x 16
    else if (colDirect && capProducts > 512 * RANK_E) SB_MAP_NUM_CD(1024, RANK_E);       // 4097..8192: two CTAs per SM
    else if (colDirect > 1 && capProducts > 256 * RANK_E) SB_MAP_NUM_CD(512, RANK_E);
    else SB_RANK_SHAPES(SB_MAP_NUM);
#undef SB_MAP_NUM
#undef SB_MAP_NUM_CD
}

#define SB_INST(T)                                                                                                     \
    template void launch_rank_numeric<T>(const LaunchCtx &, u32, const u32 *, u32, const u32 *, const u32 *, const T *, \
                                         const u32 *, const u32 *, const T *, const u32 *, const u32 *, const u32 *,   \
                                         const u32 *, u32 *, T *);                                                     \
    template void launch_map_numeric_cta<T>(const LaunchCtx &, u32, const RowDesc *, u32, const uint2 *, const T *,    \
                                            const u32 *, const T *, const unsigned short *, u32 *, T *, int, int);
SB_INST(double)
SB_INST(float)
#undef SB_INST

}  // namespace sb
