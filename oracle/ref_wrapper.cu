// oracle/ref_wrapper.cu -- C entry points over the UNMODIFIED reference implementation.
//
// TEST / BASELINE INFRASTRUCTURE ONLY.  This file is compiled by oracle/build_ref.sh together
// with the reference's own sources, read in place from /root/reference (never copied into this
// repository); the outputs go to oracle/_ref/ (git-ignored).  It drives the reference exactly
// like its Executor does (source/Executor.cpp:43-72): spECKConfig::initialize(0), warm-up
// iterations, timed iterations with Timings::measureCompleteTime, C re-used across iterations.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "CSR.h"
#include "Compare.h"
#include "Multiply.h"
#include "dCSR.h"

namespace {
template <typename T>
void fill(CSR<T> &m, size_t rows, size_t cols, size_t nnz, const uint32_t *rp, const uint32_t *ci, const T *v)
{
    m.alloc(rows, cols, nnz);
    memcpy(m.row_offsets.get(), rp, (rows + 1) * sizeof(uint32_t));
    if (nnz) {
        memcpy(m.col_ids.get(), ci, nnz * sizeof(uint32_t));
        memcpy(m.data.get(), v, nnz * sizeof(T));
    }
}

// C = A.B with reference spECK.  b* may alias a*.  Host output arrays are malloc'ed; free with
// ref_speck_free.  times_ms (>= iters floats) receives timings.complete of every timed iteration;
// stage_ms (12 floats) the mean per-stage times when track_stages != 0 (a second set of
// iterations, as TrackIndividualTimes perturbs the total).  Returns 0 on success.
template <typename T>
int multiply_impl(size_t rowsA, size_t colsA, size_t nnzA, const uint32_t *aRp, const uint32_t *aCi, const T *aV,
                  size_t rowsB, size_t colsB, size_t nnzB, const uint32_t *bRp, const uint32_t *bCi, const T *bV,
                  int warmup, int iters, int track_stages, uint64_t *nnzC, uint32_t **cRp, uint32_t **cCi, T **cV,
                  float *times_ms, float *stage_ms)
{
    try {
        CSR<T> hA, hB;
        fill(hA, rowsA, colsA, nnzA, aRp, aCi, aV);
        dCSR<T> dA, dB, dC;
        convert(dA, hA, 0);
        const bool alias = (aRp == bRp && aCi == bCi && aV == bV);
        if (alias) {
            convert(dB, hA, 0);  // the reference driver uploads A twice as well (DataLoader.cpp:70-73)
        } else {
            fill(hB, rowsB, colsB, nnzB, bRp, bCi, bV);
            convert(dB, hB, 0);
        }
        auto config = spECK::spECKConfig::initialize(0);
        Timings t;
        for (int i = 0; i < warmup; ++i) {
            t = Timings();
            t.measureCompleteTime = true;
            spECK::MultiplyspECK<T, 4, 1024, spECK_DYNAMIC_MEM_PER_BLOCK, spECK_STATIC_MEM_PER_BLOCK>(dA, dB, dC, config, t);
        }
        for (int i = 0; i < iters; ++i) {
            t = Timings();
            t.measureCompleteTime = true;
            spECK::MultiplyspECK<T, 4, 1024, spECK_DYNAMIC_MEM_PER_BLOCK, spECK_STATIC_MEM_PER_BLOCK>(dA, dB, dC, config, t);
            if (times_ms) times_ms[i] = t.complete;
        }
        if (track_stages && stage_ms) {
            Timings acc;
            const int n = iters > 0 ? iters : 1;
            for (int i = 0; i < n; ++i) {
                t = Timings();
                t.measureAll = true;
                t.measureCompleteTime = true;
                spECK::MultiplyspECK<T, 4, 1024, spECK_DYNAMIC_MEM_PER_BLOCK, spECK_STATIC_MEM_PER_BLOCK>(dA, dB, dC, config, t);
                acc += t;
            }
            acc /= float(n);
            const float s[12] = {acc.init, acc.countProducts, acc.loadBalanceCounting, acc.globalMapsCounting,
                                 acc.spGEMMCounting, acc.allocC, acc.loadBalanceNumeric, acc.globalMapsNumeric,
                                 acc.spGEMMNumeric, acc.sorting, acc.cleanup, acc.complete};
            memcpy(stage_ms, s, sizeof(s));
        }
        cudaDeviceSynchronize();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            fprintf(stderr, "ref_speck: CUDA error %s\n", cudaGetErrorString(e));
            return 2;
        }
        *nnzC = dC.nnz;
        *cRp = (uint32_t *)calloc(rowsA + 1, sizeof(uint32_t));
        *cCi = (uint32_t *)malloc((dC.nnz ? dC.nnz : 1) * sizeof(uint32_t));
        *cV = (T *)malloc((dC.nnz ? dC.nnz : 1) * sizeof(T));
        if (dC.row_offsets) cudaMemcpy(*cRp, dC.row_offsets, (rowsA + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        if (dC.nnz && dC.col_ids) cudaMemcpy(*cCi, dC.col_ids, dC.nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        if (dC.nnz && dC.data) cudaMemcpy(*cV, dC.data, dC.nnz * sizeof(T), cudaMemcpyDeviceToHost);
        config.cleanup();
        return 0;
    } catch (...) {
        fprintf(stderr, "ref_speck: exception inside the reference\n");
        return 1;
    }
}
}  // namespace

extern "C" {

int ref_speck_static_smem(void) { return spECK_STATIC_MEM_PER_BLOCK; }
int ref_speck_dynamic_smem(void) { return spECK_DYNAMIC_MEM_PER_BLOCK; }

// fp64 / fp32 instantiations of the reference (source/GPU/Multiply.cu:1130-1131)
int ref_speck_multiply_f64(size_t rowsA, size_t colsA, size_t nnzA, const uint32_t *aRp, const uint32_t *aCi,
                           const double *aV, size_t rowsB, size_t colsB, size_t nnzB, const uint32_t *bRp,
                           const uint32_t *bCi, const double *bV, int warmup, int iters, int track_stages,
                           uint64_t *nnzC, uint32_t **cRp, uint32_t **cCi, double **cV, float *times_ms,
                           float *stage_ms)
{
    return multiply_impl<double>(rowsA, colsA, nnzA, aRp, aCi, aV, rowsB, colsB, nnzB, bRp, bCi, bV, warmup, iters,
                                 track_stages, nnzC, cRp, cCi, cV, times_ms, stage_ms);
}
int ref_speck_multiply_f32(size_t rowsA, size_t colsA, size_t nnzA, const uint32_t *aRp, const uint32_t *aCi,
                           const float *aV, size_t rowsB, size_t colsB, size_t nnzB, const uint32_t *bRp,
                           const uint32_t *bCi, const float *bV, int warmup, int iters, int track_stages,
                           uint64_t *nnzC, uint32_t **cRp, uint32_t **cCi, float **cV, float *times_ms,
                           float *stage_ms)
{
    return multiply_impl<float>(rowsA, colsA, nnzA, aRp, aCi, aV, rowsB, colsB, nnzB, bRp, bCi, bV, warmup, iters,
                                track_stages, nnzC, cRp, cCi, cV, times_ms, stage_ms);
}

void ref_speck_free(void *p) { free(p); }

}  // extern "C"
