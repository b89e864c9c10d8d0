// include/COO.h -- host COO container + MatrixMarket reader (reference include/COO.h, source/COO.cpp).
#pragma once
#include <cstddef>
#include <memory>
#include "Vector.h"

template <typename T>
struct COO {
    size_t rows = 0, cols = 0, nnz = 0;
    std::unique_ptr<T[]> data;
    std::unique_ptr<unsigned int[]> row_ids;
    std::unique_ptr<unsigned int[]> col_ids;
    void alloc(size_t rows, size_t cols, size_t nnz);
};

// MatrixMarket "coordinate" files: real|integer|double|pattern(->1)|complex(real part),
// general|symmetric|hermitian (off-diagonals mirrored); skew-symmetric throws; duplicates kept
// (semantics of the reference's loadMTX, source/COO.cpp:52-164).
template <typename T> COO<T> loadMTX(const char *file);
template <typename T> void spmv(DenseVector<T> &res, const COO<T> &m, const DenseVector<T> &v, bool transpose = false);
