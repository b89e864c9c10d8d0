// speck_b200/csrc/kernels_flat.cu -- launchers of the flat (staged) symbolic rank kernel and of the hash-count
// experiment (rank_flat.cuh).
#include "rank_flat.cuh"
#include "map_seg.cuh"

namespace sb {

void launch_rank_flat(const LaunchCtx &lc, u32 capProducts, int perThread, const RowDesc *desc, u32 count,
                      const uint2 *aSeg, const u32 *bCi, unsigned short *rankMap, u32 *cRp)
{
    if (count == 0) return;
#define SB_FLAT(TH, E) launch_rank_flat_shape<TH, E>(lc, desc, count, aSeg, bCi, rankMap, cRp)
    if (perThread >= 16) {
        if (capProducts <= 1024) SB_FLAT(64, 16);
        else if (capProducts <= 2048) SB_FLAT(128, 16);
        else if (capProducts <= 4096) SB_FLAT(256, 16);
        else if (capProducts <= 8192) SB_FLAT(512, 16);
        else SB_FLAT(1024, 16);
    } else {
        if (capProducts <= 256) SB_FLAT(32, 8);         // lane-group classes routed here by "flat_min_class"
        else if (capProducts <= 512) SB_FLAT(64, 8);
        else if (capProducts <= 1024) SB_FLAT(128, 8);
        else if (capProducts <= 2048) SB_FLAT(256, 8);
        else if (capProducts <= 4096) SB_FLAT(512, 8);
        else if (capProducts <= 8192) SB_FLAT(1024, 8);
        else SB_FLAT(1024, 16);
    }
#undef SB_FLAT
}

void launch_hash_count(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                       const u32 *aCi, const u32 *bRp, const u32 *bCi, u32 *cRp)
{
    if (count == 0) return;
#define SB_HASH(TH, LOG) launch_hash_count_shape<TH, LOG>(lc, perm, count, aRp, aCi, bRp, bCi, cRp)
    if (capProducts <= 1024) SB_HASH(128, 11);
    else if (capProducts <= 2048) SB_HASH(256, 12);
    else if (capProducts <= 4096) SB_HASH(512, 13);
    else if (capProducts <= 8192) SB_HASH(1024, 14);
    else SB_HASH(1024, 15);
#undef SB_HASH
}

template <typename T>
void launch_map_seg(const LaunchCtx &lc, u32 capProducts, const RowDesc *desc, u32 count, const uint2 *aSeg,
                    const u32 *aOff, const T *aV, const u32 *bCi, const T *bV, const unsigned short *rankMap,
                    u32 *cCi, T *cV)
{
    if (count == 0) return;
#define SB_SEG(TH, E) launch_map_seg_shape<TH, E, T>(lc, desc, count, aSeg, aOff, aV, bCi, bV, rankMap, cCi, cV)
    if (capProducts <= 512) SB_SEG(64, 8);
    else if (capProducts <= 1024) SB_SEG(128, 8);
    else if (capProducts <= 2048) SB_SEG(256, 8);
    else if (capProducts <= 4096) SB_SEG(512, 8);
    else if (capProducts <= 8192) SB_SEG(1024, 8);
    else SB_SEG(1024, 16);
#undef SB_SEG
}
template void launch_map_seg<double>(const LaunchCtx &, u32, const RowDesc *, u32, const uint2 *, const u32 *,
                                     const double *, const u32 *, const double *, const unsigned short *, u32 *,
                                     double *);
template void launch_map_seg<float>(const LaunchCtx &, u32, const RowDesc *, u32, const uint2 *, const u32 *,
                                    const float *, const u32 *, const float *, const unsigned short *, u32 *, float *);

}  // namespace sb
