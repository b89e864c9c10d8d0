// include/dCSR.h -- device CSR container of the spECK API, ABI-compatible with the reference's
// include/dCSR.h:9-47 (same member order: rows, cols, nnz, data, row_offsets, col_ids; virtual
// destructor, so the vptr comes first).  Arrays are cudaMalloc'ed and owned by the object; the
// multiply (re)allocates them with the reuse rules of source/GPU/Multiply.cu:155-165, 589-592.
#pragma once
#include <cstddef>

template <typename T> struct CSR;

template <typename T>
struct dCSR {
    size_t rows = 0, cols = 0, nnz = 0;
    T *data = nullptr;
    unsigned int *row_offsets = nullptr;
    unsigned int *col_ids = nullptr;

    dCSR() = default;
    dCSR(const dCSR &) = delete;
    dCSR &operator=(const dCSR &) = delete;
    // frees everything, then allocates data/col_ids (and row_offsets when allocOffsets)
    void alloc(size_t rows, size_t cols, size_t nnz, bool allocOffsets = true);
    void reset();
    virtual ~dCSR();
};

// Non-owning POD view (reference: dCSRNoDealloc, include/dCSR.h:24-35).  Its six data members
// have the layout of the C ABI's speck_csr (include/speck_b200.h).
template <typename T>
struct dCSRNoDealloc {
    size_t rows, cols, nnz;
    T *data;
    unsigned int *row_offsets;
    unsigned int *col_ids;

    dCSRNoDealloc() = default;
    dCSRNoDealloc(const dCSR<T> &m)
        : rows(m.rows), cols(m.cols), nnz(m.nnz), data(m.data), row_offsets(m.row_offsets), col_ids(m.col_ids) {}
};

// host <-> device copies (reference source/dCSR.cpp:50-96)
template <typename T> void convert(dCSR<T> &dst, const CSR<T> &src, unsigned int padding = 0);
template <typename T> void convert(dCSR<T> &dst, const dCSR<T> &src, unsigned int padding = 0);
template <typename T> void convert(CSR<T> &dst, const dCSR<T> &src, unsigned int padding = 0);
template <typename T> void convert(CSR<T> &dst, const CSR<T> &src, unsigned int padding = 0);
