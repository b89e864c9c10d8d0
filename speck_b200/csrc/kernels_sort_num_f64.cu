// numeric sort classes, fp64 values
#include "sort_numeric_impl.cuh"
namespace sb {
template void launch_sort_numeric<double>(const LaunchCtx &, int, bool, const u32 *, u32, const u32 *,
                                          const u32 *, const double *, const u32 *, const u32 *,
                                          const double *, const u32 *, const u32 *, u32 *, double *);
template void launch_map_numeric<double>(const LaunchCtx &, int, const RowDesc *, u32, const uint2 *, const double *,
                                         const u32 *, const double *, const unsigned short *, u32 *, double *);
}
