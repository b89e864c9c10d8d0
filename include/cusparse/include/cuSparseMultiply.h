// include/cusparse/include/cuSparseMultiply.h -- the reference's cuSPARSE comparator interface
// (externals/cusparse/include/cuSparseMultiply.h:10-88) on the CUDA 12 generic API: the legacy
// cusparse?csrgemm / csr2csc entry points it used were removed in CUDA 11 (SURVEY section 0,
// fact 7).  Implementation: speck_b200/host/cusparse_shim.cu.
#pragma once
#include <cstdint>
#include "dCSR.h"

namespace cuSPARSE {

template <typename DataType>
class CuSparseTest {
public:
    CuSparseTest();
    ~CuSparseTest();
    // C = A.B with cusparseSpGEMM; rows of C column-sorted (the reference compares positionally,
    // source/GPU/Compare.cu:27-47).  Returns the elapsed milliseconds.
    float Multiply(const dCSR<DataType> &A, const dCSR<DataType> &B, dCSR<DataType> &matOut, uint32_t &cusparse_nnz);
    // AT = transpose(A) with cusparseCsr2cscEx2
    void Transpose(const dCSR<DataType> &A, dCSR<DataType> &AT);

private:
    void *handle = nullptr;   // cusparseHandle_t
};

}  // namespace cuSPARSE
