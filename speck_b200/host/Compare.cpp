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
    speck_mismatch mm{};
    const int rc = std::is_same<DataType, float>::value ? speck_b200_compare_report_f32(ctx, &a, &b, compare_data, 0.01, &mm)
                                                        : speck_b200_compare_report_f64(ctx, &a, &b, compare_data, 0.01, &mm);
    if (rc < 0) printf("ERROR: %s\n", speck_b200_last_error());
    if (rc == 0) {   // what the reference's d_compare never tells: where the first difference is
        if (mm.kind == 3)
            printf("Compare: matrices differ in shape or nnz (%zux%zu nnz %zu vs %zux%zu nnz %zu)\n", ref.rows, ref.cols, ref.nnz, cmp.rows, cmp.cols, cmp.nnz);
        else if (mm.kind == 0)
            printf("Compare: first difference in row %llu: %u entries vs %u\n", (unsigned long long)mm.row, mm.ref_len, mm.cmp_len);
        else if (mm.kind == 1)
            printf("Compare: first difference in row %llu, entry %u: column %u vs %u\n", (unsigned long long)mm.row, mm.index_in_row, mm.ref_col, mm.cmp_col);
        else
            printf("Compare: first difference in row %llu, entry %u (column %u): value %.17g vs %.17g\n", (unsigned long long)mm.row, mm.index_in_row, mm.ref_col, mm.ref_val, mm.cmp_val);
    }
    return rc == 1;
}

template bool Compare<float>(const dCSR<float> &, const dCSR<float> &, bool);
template bool Compare<double>(const dCSR<double> &, const dCSR<double> &, bool);

}  // namespace spECK
