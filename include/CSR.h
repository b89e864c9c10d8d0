// include/CSR.h -- host CSR container + .hicsr cache I/O (reference include/CSR.h, source/CSR.cpp).
#pragma once
#include <cstddef>
#include <memory>

template <typename T> struct COO;
template <typename T> struct DenseVector;

template <typename T>
struct CSR {
    struct Statistics { double mean, std_dev; size_t max, min; };

    size_t rows = 0, cols = 0, nnz = 0;
    std::unique_ptr<T[]> data;
    std::unique_ptr<unsigned int[]> row_offsets;
    std::unique_ptr<unsigned int[]> col_ids;

    void alloc(size_t rows, size_t cols, size_t nnz);
    // row-length statistics (Welford), as CSR<T>::computeStatistics in the reference
    void computeStatistics(double &mean, double &std_dev, size_t &max, size_t &min) const;
    Statistics rowStatistics() const
    {
        Statistics s;
        computeStatistics(s.mean, s.std_dev, s.max, s.min);
        return s;
    }
};

// binary ".hicsr" cache, byte-compatible with the reference (source/CSR.cpp:27-137)
template <typename T> CSR<T> loadCSR(const char *file);
template <typename T> void storeCSR(const CSR<T> &mat, const char *file);
// COO -> CSR: sort by (row, col); duplicates are kept (source/CSR.cpp:173-212)
template <typename T> void convert(CSR<T> &dst, const COO<T> &src);
template <typename T> void spmv(DenseVector<T> &res, const CSR<T> &m, const DenseVector<T> &v, bool transpose = false);
