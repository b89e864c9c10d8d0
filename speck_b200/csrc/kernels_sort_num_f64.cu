// numeric sort classes, fp64 values
#include "sort_numeric_impl.cuh"
namespace sb {
template void launch_sort_numeric<double>(const LaunchCtx &, int, bool, const u32 *, u32, const u32 *,
                                          const u32 *, const double *, const u32 *, const u32 *,
                                          const double *, const u32 *, const u32 *, u32 *, double *);
}
