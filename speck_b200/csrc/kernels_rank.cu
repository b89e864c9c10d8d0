// speck_b200/csrc/kernels_rank.cu -- launchers of the rank row classes (rank_cta.cuh).
// Four launch shapes (128 / 256 / 512 / 1024 threads, 8 product slots per thread): `capProducts` is the
// largest product count among the rows of the launch.
#include "rank_cta.cuh"

namespace sb {

// product slots per thread: 8 (128..1024 threads) or 4 (256..1024 threads, 8 for the largest shape)
static int g_rankE = 8;
void set_rank_slots(int e) { g_rankE = (e == 4 || e == 16) ? e : 8; }

void launch_rank_symbolic(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                          const u32 *aCi, const u32 *bRp, const u32 *bCi, const u32 *rowOps, const u32 *rowMin,
                          const u32 *rowMax, u32 *rowNnz)
{
    if (count == 0) return;
    const float *nv = nullptr;
#define SB_RANK_SYM(TH, E) \
    launch_rank_rows<TH, E, float, false>(lc, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowMin, rowMax, rowNnz, nullptr, nullptr)
    if (g_rankE == 4) {
        if (capProducts <= 1024) SB_RANK_SYM(256, 4);
        else if (capProducts <= 2048) SB_RANK_SYM(512, 4);
        else if (capProducts <= 4096) SB_RANK_SYM(1024, 4);
        else SB_RANK_SYM(1024, 8);
    } else if (g_rankE == 16) {
        if (capProducts <= 1024) SB_RANK_SYM(64, 16);
        else if (capProducts <= 2048) SB_RANK_SYM(128, 16);
        else if (capProducts <= 4096) SB_RANK_SYM(256, 16);
        else SB_RANK_SYM(512, 16);
    } else {
        if (capProducts <= 1024) SB_RANK_SYM(128, 8);
        else if (capProducts <= 2048) SB_RANK_SYM(256, 8);
        else if (capProducts <= 4096) SB_RANK_SYM(512, 8);
        else SB_RANK_SYM(1024, 8);
    }
#undef SB_RANK_SYM
}

template <typename T>
void launch_rank_numeric(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                         const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi, const T *bV,
                         const u32 *rowOps, const u32 *rowMin, const u32 *rowMax, const u32 *cRp, u32 *cCi, T *cV)
{
    if (count == 0) return;
    u32 *rp = const_cast<u32 *>(cRp);
#define SB_RANK_NUM(TH, E) \
    launch_rank_rows<TH, E, T, true>(lc, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, rowMin, rowMax, rp, cCi, cV)
    if (g_rankE == 4) {
        if (capProducts <= 1024) SB_RANK_NUM(256, 4);
        else if (capProducts <= 2048) SB_RANK_NUM(512, 4);
        else if (capProducts <= 4096) SB_RANK_NUM(1024, 4);
        else SB_RANK_NUM(1024, 8);
    } else if (g_rankE == 16) {
        if (capProducts <= 1024) SB_RANK_NUM(64, 16);
        else if (capProducts <= 2048) SB_RANK_NUM(128, 16);
        else if (capProducts <= 4096) SB_RANK_NUM(256, 16);
        else SB_RANK_NUM(512, 16);
    } else {
        if (capProducts <= 1024) SB_RANK_NUM(128, 8);
        else if (capProducts <= 2048) SB_RANK_NUM(256, 8);
        else if (capProducts <= 4096) SB_RANK_NUM(512, 8);
        else SB_RANK_NUM(1024, 8);
    }
#undef SB_RANK_NUM
}
template void launch_rank_numeric<double>(const LaunchCtx &, u32, const u32 *, u32, const u32 *, const u32 *,
                                          const double *, const u32 *, const u32 *, const double *, const u32 *,
                                          const u32 *, const u32 *, const u32 *, u32 *, double *);
template void launch_rank_numeric<float>(const LaunchCtx &, u32, const u32 *, u32, const u32 *, const u32 *,
                                         const float *, const u32 *, const u32 *, const float *, const u32 *,
                                         const u32 *, const u32 *, const u32 *, u32 *, float *);

}  // namespace sb
