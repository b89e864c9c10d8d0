// speck_b200/host/Multiply.cpp -- spECK::MultiplyspECK<> over the C ABI (include/speck_b200.h).
// Same explicit instantiations as the reference (source/GPU/Multiply.cu:1124-1131) and the same
// error conventions (source/GPU/Multiply.cu:57-97, 594-599): precondition failures print
// "ERROR: ..." and return with matOut untouched; CUDA failures throw std::exception like
// HANDLE_ERROR (include/common.h:19-31).
#include <cstdio>
#include <exception>
#include <type_traits>
#include "Multiply.h"
#include "speck_b200.h"

namespace spECK {

template <typename T>
static speck_csr view_of(const dCSR<T> &m)
{
    speck_csr v;
    v.rows = m.rows; v.cols = m.cols; v.nnz = m.nnz;
    v.data = m.data; v.row_offsets = m.row_offsets; v.col_ids = m.col_ids;
    return v;
}

template <typename DataType, int BLOCKS_PER_SM, int THREADS_PER_BLOCK, int MAX_DYNAMIC_SHARED, int MAX_STATIC_SHARED>
void MultiplyspECKImplementation(const dCSR<DataType> &A, const dCSR<DataType> &B, dCSR<DataType> &matOut,
                                 spECKConfig &config, Timings &timings)
{
    if (!config.b200) {
        if (speck_b200_create(config.device, &config.b200) != SPECK_OK) {
            printf("ERROR: %s\n", speck_b200_last_error());
            return;
        }
    }
    const speck_csr a = view_of(A), b = view_of(B);
    speck_csr c = view_of(matOut);
    speck_timings t{};
    t.measure_all = timings.measureAll;
    t.measure_complete = timings.measureCompleteTime;
    int rc;
    if (std::is_same<DataType, float>::value) rc = speck_b200_spgemm_f32(config.b200, &a, &b, &c, &t);
    else rc = speck_b200_spgemm_f64(config.b200, &a, &b, &c, &t);
    // the callee may have re-allocated the arrays: write them back even on late failures
    matOut.rows = c.rows; matOut.cols = c.cols; matOut.nnz = c.nnz;
    matOut.data = static_cast<DataType *>(c.data); matOut.row_offsets = c.row_offsets; matOut.col_ids = c.col_ids;
    if (rc == SPECK_ERR_CUDA) {
        printf("%s\n", speck_b200_last_error());
        throw std::exception();
    }
    if (rc != SPECK_OK) {
        printf("ERROR: %s\n", speck_b200_last_error());
        return;
    }
    const float s[Timings::kStages] = {t.init, t.count_products, t.load_balance_counting, t.global_maps_counting,
                                      t.spgemm_counting, t.alloc_c, t.load_balance_numeric, t.global_maps_numeric,
                                      t.spgemm_numeric, t.sorting, t.cleanup, t.complete};
    if (timings.measureAll)
        for (int i = 0; i < Timings::kStages - 1; ++i) *timings.stage(i) = s[i];
    if (timings.measureAll || timings.measureCompleteTime) timings.complete = t.complete;
    if (timings.measureAll) {  // per-iteration stage table, as the reference prints it (Multiply.cu:1097-1113)
        static const char *names[Timings::kStages] = {"init", "countProducts", "loadBalanceCounting", "globalMapsCounting",
                                                      "spGEMMCounting", "allocC", "loadBalanceNumeric", "globalMapsNumeric",
                                                      "spGEMMNumeric", "sorting", "cleanup", "complete"};
        printf("spECK     %s\n", "timings (ms)");
        for (int i = 0; i < Timings::kStages; ++i) printf("  %-22s %10.4f\n", names[i], s[i]);
    }
}

template <typename DataType, int BLOCKS_PER_SM, int THREADS_PER_BLOCK, int MAX_DYNAMIC_SHARED, int MAX_STATIC_SHARED>
void MultiplyspECK(const dCSR<DataType> &A, const dCSR<DataType> &B, dCSR<DataType> &matOut, spECKConfig &config,
                   Timings &timings)
{
    MultiplyspECKImplementation<DataType, BLOCKS_PER_SM, THREADS_PER_BLOCK, MAX_DYNAMIC_SHARED, MAX_STATIC_SHARED>(
        A, B, matOut, config, timings);
}

template void MultiplyspECK<float, 4, 1024, spECK_DYNAMIC_MEM_PER_BLOCK, spECK_STATIC_MEM_PER_BLOCK>(
    const dCSR<float> &, const dCSR<float> &, dCSR<float> &, spECKConfig &, Timings &);
template void MultiplyspECK<double, 4, 1024, spECK_DYNAMIC_MEM_PER_BLOCK, spECK_STATIC_MEM_PER_BLOCK>(
    const dCSR<double> &, const dCSR<double> &, dCSR<double> &, spECKConfig &, Timings &);
template void MultiplyspECKImplementation<float, 4, 1024, spECK_DYNAMIC_MEM_PER_BLOCK, spECK_STATIC_MEM_PER_BLOCK>(
    const dCSR<float> &, const dCSR<float> &, dCSR<float> &, spECKConfig &, Timings &);
template void MultiplyspECKImplementation<double, 4, 1024, spECK_DYNAMIC_MEM_PER_BLOCK, spECK_STATIC_MEM_PER_BLOCK>(
    const dCSR<double> &, const dCSR<double> &, dCSR<double> &, spECKConfig &, Timings &);

}  // namespace spECK
