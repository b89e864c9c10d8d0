// speck_b200/host/Compare.cpp -- spECK::Compare over the C ABI (reference source/GPU/Compare.cu:66-82).
// Values, when requested, are compared at the reference's 1 % relative tolerance (Compare.cu:49-58).
#include <cstdio>
#include <type_traits>
#include <cuda_runtime.h>
#include "Compare.h"
#include "speck_b200.h"

namespace spECK {

template <typename DataType>
bool Compare(const dCSR<DataType> &ref, const dCSR<DataType> &cmp, bool compare_data)
{
    static speck_ctx *ctx = nullptr;   // lazily created on the current device
    if (!ctx) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (speck_b200_create(dev, &ctx) != SPECK_OK) {
            printf("ERROR: %s\n", speck_b200_last_error());
            return false;
        }
    }
    speck_csr a{ref.rows, ref.cols, ref.nnz, ref.data, ref.row_offsets, ref.col_ids};
    speck_csr b{cmp.rows, cmp.cols, cmp.nnz, cmp.data, cmp.row_offsets, cmp.col_ids};
    const int rc = std::is_same<DataType, float>::value ? speck_b200_compare_f32(ctx, &a, &b, compare_data, 0.01)
                                                        : speck_b200_compare_f64(ctx, &a, &b, compare_data, 0.01);
    if (rc < 0) printf("ERROR: %s\n", speck_b200_last_error());
    return rc == 1;
}

template bool Compare<float>(const dCSR<float> &, const dCSR<float> &, bool);
template bool Compare<double>(const dCSR<double> &, const dCSR<double> &, bool);

}  // namespace spECK
