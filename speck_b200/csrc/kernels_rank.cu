// speck_b200/csrc/kernels_rank.cu -- launchers of the rank row classes (rank_cta.cuh).
#include "rank_cta.cuh"

namespace sb {

constexpr int RANK_E = 8;   // product slots per thread: a class of 512*(c+2) products runs 64*(c+2) threads

void launch_rank_symbolic(const LaunchCtx &lc, int ctaClass, const u32 *perm, u32 count, const u32 *aRp,
                          const u32 *aCi, const u32 *bRp, const u32 *bCi, const u32 *rowOps, const u32 *rowMin,
                          const u32 *rowMax, u32 *rowNnz)
{
    if (count == 0) return;
    const float *nv = nullptr;
    const int threads = 512 * (ctaClass + 2) / RANK_E;
    launch_rank_rows<RANK_E, 1024, float, false>(lc, threads, perm, count, aRp, aCi, nv, bRp, bCi, nv, rowOps, rowMin, rowMax,
                                                 rowNnz, nullptr, nullptr);
}

template <typename T>
void launch_rank_numeric(const LaunchCtx &lc, int ctaClass, const u32 *perm, u32 count, const u32 *aRp,
                         const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi, const T *bV,
                         const u32 *rowOps, const u32 *rowMin, const u32 *rowMax, const u32 *cRp, u32 *cCi, T *cV)
{
    if (count == 0) return;
    const int threads = 512 * (ctaClass + 2) / RANK_E;
    launch_rank_rows<RANK_E, 1024, T, true>(lc, threads, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, rowMin, rowMax,
                                            const_cast<u32 *>(cRp), cCi, cV);
}
template void launch_rank_numeric<double>(const LaunchCtx &, int, const u32 *, u32, const u32 *, const u32 *,
                                          const double *, const u32 *, const u32 *, const double *, const u32 *,
                                          const u32 *, const u32 *, const u32 *, u32 *, double *);
template void launch_rank_numeric<float>(const LaunchCtx &, int, const u32 *, u32, const u32 *, const u32 *,
                                         const float *, const u32 *, const u32 *, const float *, const u32 *,
                                         const u32 *, const u32 *, const u32 *, u32 *, float *);

}  // namespace sb
