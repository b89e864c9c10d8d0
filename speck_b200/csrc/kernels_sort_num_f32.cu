// numeric sort classes, fp32 values
#include "sort_numeric_impl.cuh"
namespace sb {
template void launch_sort_numeric<float>(const LaunchCtx &, int, bool, const u32 *, u32, const u32 *,
                                         const u32 *, const float *, const u32 *, const u32 *,
                                         const float *, const u32 *, const u32 *, u32 *, float *);
template void launch_map_numeric<float>(const LaunchCtx &, int, const RowDesc *, u32, const uint2 *, const float *,
                                         const u32 *, const float *, const unsigned short *, u32 *, float *);
}
