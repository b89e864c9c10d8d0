// speck_b200/host/dCSR.cpp -- device CSR container (reference source/dCSR.cpp): owning
// cudaMalloc'ed arrays, host<->device converts.  Every CUDA call is checked (the reference
// ignores cudaMalloc failures, dCSR.cpp:32-35).
#include <cstring>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>
#include "CSR.h"
#include "dCSR.h"

namespace {
void cu(cudaError_t e, const char *what)
{
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
template <typename T>
void release(dCSR<T> &m)
{
    cudaFree(m.col_ids);
    cudaFree(m.data);
    cudaFree(m.row_offsets);
    m.col_ids = nullptr;
    m.data = nullptr;
    m.row_offsets = nullptr;
    m.rows = m.nnz = 0;
}
template <typename T>
void copy3(T *dData, unsigned *dCols, unsigned *dOffs, const T *sData, const unsigned *sCols, const unsigned *sOffs,
           size_t rows, size_t nnz, cudaMemcpyKind kind)
{
    if (nnz) {
        cu(cudaMemcpy(dData, sData, nnz * sizeof(T), kind), "copy values");
        cu(cudaMemcpy(dCols, sCols, nnz * sizeof(unsigned), kind), "copy col_ids");
    }
    cu(cudaMemcpy(dOffs, sOffs, (rows + 1) * sizeof(unsigned), kind), "copy row_offsets");
}
}  // namespace

template <typename T>
void dCSR<T>::alloc(size_t r, size_t c, size_t n, bool allocOffsets)
{
    release(*this);
    rows = r;
    cols = c;
    nnz = n;
    if (n) {
        cu(cudaMalloc(&data, n * sizeof(T)), "cudaMalloc values");
        cu(cudaMalloc(&col_ids, n * sizeof(unsigned)), "cudaMalloc col_ids");
    }
    if (allocOffsets) cu(cudaMalloc(&row_offsets, (r + 1) * sizeof(unsigned)), "cudaMalloc row_offsets");
}
template <typename T> dCSR<T>::~dCSR() { release(*this); }
template <typename T> void dCSR<T>::reset() { release(*this); }

template <typename T>
void convert(dCSR<T> &dst, const CSR<T> &src, unsigned int padding)
{
    dst.alloc(src.rows + padding, src.cols, src.nnz + 8 * padding);
    dst.rows = src.rows;
    dst.nnz = src.nnz;
    copy3(dst.data, dst.col_ids, dst.row_offsets, src.data.get(), src.col_ids.get(), src.row_offsets.get(), src.rows,
          src.nnz, cudaMemcpyHostToDevice);
    if (padding) {
        cudaMemset(dst.data + src.nnz, 0, 8 * padding * sizeof(T));
        cudaMemset(dst.col_ids + src.nnz, 0, 8 * padding * sizeof(unsigned));
        cudaMemset(dst.row_offsets + src.rows + 1, 0, padding * sizeof(unsigned));
    }
}
template <typename T>
void convert(CSR<T> &dst, const dCSR<T> &src, unsigned int padding)
{
    dst.alloc(src.rows + padding, src.cols, src.nnz + 8 * padding);
    dst.rows = src.rows;
    dst.nnz = src.nnz;
    if (src.row_offsets)
        copy3(dst.data.get(), dst.col_ids.get(), dst.row_offsets.get(), src.data, src.col_ids, src.row_offsets, src.rows,
              src.nnz, cudaMemcpyDeviceToHost);
}
template <typename T>
void convert(dCSR<T> &dst, const dCSR<T> &src, unsigned int padding)
{
    dst.alloc(src.rows + padding, src.cols, src.nnz + 8 * padding);
    dst.rows = src.rows;
    dst.nnz = src.nnz;
    copy3(dst.data, dst.col_ids, dst.row_offsets, src.data, src.col_ids, src.row_offsets, src.rows, src.nnz,
          cudaMemcpyDeviceToDevice);
}
template <typename T>
void convert(CSR<T> &dst, const CSR<T> &src, unsigned int padding)
{
    dst.alloc(src.rows + padding, src.cols, src.nnz + 8 * padding);
    dst.rows = src.rows;
    dst.nnz = src.nnz;
    std::memcpy(dst.data.get(), src.data.get(), src.nnz * sizeof(T));
    std::memcpy(dst.col_ids.get(), src.col_ids.get(), src.nnz * sizeof(unsigned));
    std::memcpy(dst.row_offsets.get(), src.row_offsets.get(), (src.rows + 1) * sizeof(unsigned));
}

#define SPECK_INSTANTIATE(T)                                          \
    template struct dCSR<T>;                                          \
    template void convert(dCSR<T> &, const CSR<T> &, unsigned int);   \
    template void convert(CSR<T> &, const dCSR<T> &, unsigned int);   \
    template void convert(dCSR<T> &, const dCSR<T> &, unsigned int);  \
    template void convert(CSR<T> &, const CSR<T> &, unsigned int);
SPECK_INSTANTIATE(float)
SPECK_INSTANTIATE(double)
