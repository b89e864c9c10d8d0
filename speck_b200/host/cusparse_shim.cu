// speck_b200/host/cusparse_shim.cu -- cuSPARSE::CuSparseTest<T> (the reference's CompareResult
// partner and transposer, externals/cusparse/source/cuSparseMultiply.cu:46-138) on the CUDA 12
// generic API: cusparseSpGEMM_{workEstimation,compute,copy} + cusparseCsr2cscEx2.  cuSPARSE's
// SpGEMM does not promise column-sorted rows, the reference's compare is positional
// (source/GPU/Compare.cu:27-47), so rows are sorted afterwards with cusparseXcsrsort.
#include <cstdio>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>
#include <cusparse.h>
#include "cusparse/include/cuSparseMultiply.h"

namespace {
void ck(cusparseStatus_t s, const char *what)
{
    if (s != CUSPARSE_STATUS_SUCCESS) {
        printf("CuSparse error: %s (%s)\n", what, cusparseGetErrorString(s));
        throw std::runtime_error(std::string("cuSPARSE: ") + what);
    }
}
void cu(cudaError_t e, const char *what)
{
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
template <typename T> cudaDataType dtype();
template <> cudaDataType dtype<float>() { return CUDA_R_32F; }
template <> cudaDataType dtype<double>() { return CUDA_R_64F; }
}  // namespace

namespace cuSPARSE {

template <typename T>
CuSparseTest<T>::CuSparseTest()
{
    cusparseHandle_t h = nullptr;
    ck(cusparseCreate(&h), "init failed");
    handle = h;
}

template <typename T>
CuSparseTest<T>::~CuSparseTest()
{
    if (handle) cusparseDestroy((cusparseHandle_t)handle);
}

template <typename T>
float CuSparseTest<T>::Multiply(const dCSR<T> &A, const dCSR<T> &B, dCSR<T> &C, uint32_t &nnzOut)
{
    cusparseHandle_t h = (cusparseHandle_t)handle;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    C.reset();
    C.rows = A.rows;
    C.cols = B.cols;
    cu(cudaMalloc(&C.row_offsets, (A.rows + 1) * sizeof(unsigned)), "cudaMalloc C.row_offsets");

    cusparseSpMatDescr_t mA, mB, mC;
    ck(cusparseCreateCsr(&mA, A.rows, A.cols, A.nnz, A.row_offsets, A.col_ids, A.data, CUSPARSE_INDEX_32I,
                         CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dtype<T>()), "create A");
    ck(cusparseCreateCsr(&mB, B.rows, B.cols, B.nnz, B.row_offsets, B.col_ids, B.data, CUSPARSE_INDEX_32I,
                         CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dtype<T>()), "create B");
    ck(cusparseCreateCsr(&mC, A.rows, B.cols, 0, C.row_offsets, nullptr, nullptr, CUSPARSE_INDEX_32I,
                         CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dtype<T>()), "create C");
    cusparseSpGEMMDescr_t desc;
    ck(cusparseSpGEMM_createDescr(&desc), "SpGEMM descr");
    const T alpha = (T)1, beta = (T)0;
    const cusparseOperation_t op = CUSPARSE_OPERATION_NON_TRANSPOSE;
    size_t s1 = 0, s2 = 0;
    void *b1 = nullptr, *b2 = nullptr;
    ck(cusparseSpGEMM_workEstimation(h, op, op, &alpha, mA, mB, &beta, mC, dtype<T>(), CUSPARSE_SPGEMM_DEFAULT, desc,
                                     &s1, nullptr), "workEstimation (size)");
    cu(cudaMalloc(&b1, s1 ? s1 : 4), "cudaMalloc SpGEMM buffer 1");
    ck(cusparseSpGEMM_workEstimation(h, op, op, &alpha, mA, mB, &beta, mC, dtype<T>(), CUSPARSE_SPGEMM_DEFAULT, desc,
                                     &s1, b1), "workEstimation");
    ck(cusparseSpGEMM_compute(h, op, op, &alpha, mA, mB, &beta, mC, dtype<T>(), CUSPARSE_SPGEMM_DEFAULT, desc, &s2,
                              nullptr), "compute (size)");
    cu(cudaMalloc(&b2, s2 ? s2 : 4), "cudaMalloc SpGEMM buffer 2");
    ck(cusparseSpGEMM_compute(h, op, op, &alpha, mA, mB, &beta, mC, dtype<T>(), CUSPARSE_SPGEMM_DEFAULT, desc, &s2,
                              b2), "compute");
    int64_t r = 0, c = 0, nnz = 0;
    ck(cusparseSpMatGetSize(mC, &r, &c, &nnz), "get size");
    C.nnz = (size_t)nnz;
    nnzOut = (uint32_t)nnz;
    cu(cudaMalloc(&C.col_ids, (nnz ? nnz : 1) * sizeof(unsigned)), "cudaMalloc C.col_ids");
    cu(cudaMalloc(&C.data, (nnz ? nnz : 1) * sizeof(T)), "cudaMalloc C.data");
    ck(cusparseCsrSetPointers(mC, C.row_offsets, C.col_ids, C.data), "set pointers");
    ck(cusparseSpGEMM_copy(h, op, op, &alpha, mA, mB, &beta, mC, dtype<T>(), CUSPARSE_SPGEMM_DEFAULT, desc), "copy");

    // sort the columns of every row (values permuted along)
    if (nnz > 0) {
        size_t sb = 0;
        void *buf = nullptr;
        int *permIdx = nullptr;
        T *sorted = nullptr;
        cusparseMatDescr_t md;
        ck(cusparseCreateMatDescr(&md), "mat descr");
        ck(cusparseXcsrsort_bufferSizeExt(h, (int)A.rows, (int)B.cols, (int)nnz, (const int *)C.row_offsets,
                                          (const int *)C.col_ids, &sb), "csrsort size");
        cu(cudaMalloc(&buf, sb ? sb : 4), "cudaMalloc sort buffer");
        cu(cudaMalloc(&permIdx, nnz * sizeof(int)), "cudaMalloc permutation");
        cu(cudaMalloc(&sorted, nnz * sizeof(T)), "cudaMalloc sorted values");
        ck(cusparseCreateIdentityPermutation(h, (int)nnz, permIdx), "identity permutation");
        ck(cusparseXcsrsort(h, (int)A.rows, (int)B.cols, (int)nnz, md, (const int *)C.row_offsets, (int *)C.col_ids,
                            permIdx, buf), "csrsort");
        cusparseDnVecDescr_t vOut;
        cusparseSpVecDescr_t vIn;
        ck(cusparseCreateSpVec(&vIn, nnz, nnz, permIdx, sorted, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dtype<T>()),
           "spvec");
        ck(cusparseCreateDnVec(&vOut, nnz, C.data, dtype<T>()), "dnvec");
        ck(cusparseGather(h, vOut, vIn), "gather");
        cu(cudaMemcpy(C.data, sorted, nnz * sizeof(T), cudaMemcpyDeviceToDevice), "copy sorted values");
        cusparseDestroySpVec(vIn);
        cusparseDestroyDnVec(vOut);
        cusparseDestroyMatDescr(md);
        cudaFree(buf);
        cudaFree(permIdx);
        cudaFree(sorted);
    }
    cusparseSpGEMM_destroyDescr(desc);
    cusparseDestroySpMat(mA);
    cusparseDestroySpMat(mB);
    cusparseDestroySpMat(mC);
    cudaFree(b1);
    cudaFree(b2);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

template <typename T>
void CuSparseTest<T>::Transpose(const dCSR<T> &A, dCSR<T> &AT)
{
    cusparseHandle_t h = (cusparseHandle_t)handle;
    AT.alloc(A.cols, A.rows, A.nnz);
    size_t sb = 0;
    void *buf = nullptr;
    ck(cusparseCsr2cscEx2_bufferSize(h, (int)A.rows, (int)A.cols, (int)A.nnz, A.data, (const int *)A.row_offsets,
                                     (const int *)A.col_ids, AT.data, (int *)AT.row_offsets, (int *)AT.col_ids, dtype<T>(),
                                     CUSPARSE_ACTION_NUMERIC, CUSPARSE_INDEX_BASE_ZERO, CUSPARSE_CSR2CSC_ALG1, &sb),
       "csr2csc size");
    cu(cudaMalloc(&buf, sb ? sb : 4), "cudaMalloc csr2csc buffer");
    ck(cusparseCsr2cscEx2(h, (int)A.rows, (int)A.cols, (int)A.nnz, A.data, (const int *)A.row_offsets,
                          (const int *)A.col_ids, AT.data, (int *)AT.row_offsets, (int *)AT.col_ids, dtype<T>(),
                          CUSPARSE_ACTION_NUMERIC, CUSPARSE_INDEX_BASE_ZERO, CUSPARSE_CSR2CSC_ALG1, buf),
       "csr2csc");
    cudaDeviceSynchronize();
    cudaFree(buf);
}

template class CuSparseTest<float>;
template class CuSparseTest<double>;

}  // namespace cuSPARSE
