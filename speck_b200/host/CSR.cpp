// speck_b200/host/CSR.cpp -- host CSR container, .hicsr cache (byte layout of the reference,
// source/CSR.cpp:27-137: 80-byte header, 2*sizeof(T)-byte state block, values, col_ids,
// row_offsets), COO -> CSR conversion (sort by (row, col), duplicates kept, :173-212).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>
#include "COO.h"
#include "CSR.h"
#include "Vector.h"

namespace {
constexpr size_t kHeaderBytes = 80;
const unsigned char kMagic[9] = {'H', 'i', 1, 'C', 'o', 'm', 'p', 's', 'd'};
struct HiCsrHeader {  // explicit offsets instead of a reinterpret_cast'ed struct
    uint64_t typesize, compresseddir, indexsize, fixedoffset, offsetsize, rows, cols, nnz;
    void encode(unsigned char *b) const
    {
        std::memset(b, 0, kHeaderBytes);
        std::memcpy(b, kMagic, sizeof(kMagic));
        const uint64_t f[8] = {typesize, compresseddir, indexsize, fixedoffset, offsetsize, rows, cols, nnz};
        std::memcpy(b + 16, f, sizeof(f));
    }
    bool decode(const unsigned char *b)
    {
        if (std::memcmp(b, kMagic, sizeof(kMagic)) != 0) return false;
        uint64_t f[8];
        std::memcpy(f, b + 16, sizeof(f));
        typesize = f[0]; compresseddir = f[1]; indexsize = f[2]; fixedoffset = f[3]; offsetsize = f[4];
        rows = f[5]; cols = f[6]; nnz = f[7];
        return true;
    }
};
}  // namespace

template <typename T>
void CSR<T>::alloc(size_t r, size_t c, size_t n)
{
    rows = r;
    cols = c;
    nnz = n;
    data.reset(new T[n ? n : 1]());
    col_ids.reset(new unsigned int[n ? n : 1]());
    row_offsets.reset(new unsigned int[r + 1]());
}

template <typename T>
void CSR<T>::computeStatistics(double &mean, double &std_dev, size_t &max, size_t &min) const
{
    mean = 0;
    double m2 = 0;
    max = 0;
    min = cols;
    for (size_t i = 0; i < rows; ++i) {
        const size_t len = row_offsets[i + 1] - row_offsets[i];
        min = std::min(min, len);
        max = std::max(max, len);
        const double d = (double)len - mean;
        mean += d / (double)(i + 1);
        m2 += d * ((double)len - mean);
    }
    std_dev = rows < 2 ? 0.0 : std::sqrt(m2 / (double)(rows - 1));
}

template <typename T>
CSR<T> loadCSR(const char *file)
{
    std::ifstream in(file, std::ios::binary);
    if (!in.is_open()) throw std::runtime_error(std::string("could not open \"") + file + "\"");
    unsigned char hb[kHeaderBytes];
    in.read(reinterpret_cast<char *>(hb), kHeaderBytes);
    if (!in.good()) throw std::runtime_error("Could not read CSR header");
    HiCsrHeader h;
    if (!h.decode(hb)) throw std::runtime_error("File does not appear to be a CSR Matrix");
    unsigned char state[2 * sizeof(T)];
    in.read(reinterpret_cast<char *>(state), sizeof(state));
    if (!in.good()) throw std::runtime_error("Could not read CompressedMatrix state");
    if (h.typesize != sizeof(T)) throw std::runtime_error("File does not contain a CSR matrix with matching type");
    CSR<T> m;
    m.alloc(h.rows, h.cols, h.nnz);
    in.read(reinterpret_cast<char *>(m.data.get()), m.nnz * sizeof(T));
    in.read(reinterpret_cast<char *>(m.col_ids.get()), m.nnz * sizeof(unsigned int));
    in.read(reinterpret_cast<char *>(m.row_offsets.get()), (m.rows + 1) * sizeof(unsigned int));
    if (!in.good()) throw std::runtime_error("Could not read CSR matrix data");
    return m;
}

template <typename T>
void storeCSR(const CSR<T> &m, const char *file)
{
    std::ofstream out(file, std::ios::binary);
    if (!out.is_open()) throw std::runtime_error(std::string("could not open \"") + file + "\"");
    HiCsrHeader h{sizeof(T), 0, sizeof(uint32_t), 0, sizeof(uint32_t), m.rows, m.cols, m.nnz};
    unsigned char hb[kHeaderBytes];
    h.encode(hb);
    out.write(reinterpret_cast<const char *>(hb), kHeaderBytes);
    unsigned char state[2 * sizeof(T)] = {};   // {T scaling = 1; bool transpose = false; padding}
    const T one = (T)1;
    std::memcpy(state, &one, sizeof(T));
    out.write(reinterpret_cast<const char *>(state), sizeof(state));
    out.write(reinterpret_cast<const char *>(m.data.get()), m.nnz * sizeof(T));
    out.write(reinterpret_cast<const char *>(m.col_ids.get()), m.nnz * sizeof(unsigned int));
    out.write(reinterpret_cast<const char *>(m.row_offsets.get()), (m.rows + 1) * sizeof(unsigned int));
}

template <typename T>
void convert(CSR<T> &dst, const COO<T> &src)
{
    std::vector<size_t> order(src.nnz);
    std::iota(order.begin(), order.end(), (size_t)0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        return src.row_ids[a] != src.row_ids[b] ? src.row_ids[a] < src.row_ids[b] : src.col_ids[a] < src.col_ids[b];
    });
    dst.alloc(src.rows, src.cols, src.nnz);
    std::cout << src.nnz << std::endl;   // the reference prints the nnz here (CSR.cpp:189)
    for (size_t i = 0; i < src.nnz; ++i) {
        const size_t s = order[i];
        dst.data[i] = src.data[s];
        dst.col_ids[i] = src.col_ids[s];
        dst.row_offsets[src.row_ids[s] + 1]++;
    }
    for (size_t r = 0; r < src.rows; ++r) dst.row_offsets[r + 1] += dst.row_offsets[r];
}

template <typename T>
void spmv(DenseVector<T> &res, const CSR<T> &m, const DenseVector<T> &v, bool transpose)
{
    if ((transpose ? m.rows : m.cols) != v.size) throw std::runtime_error("SPMV dimensions mismatch");
    res.alloc(transpose ? m.cols : m.rows);
    for (size_t r = 0; r < m.rows; ++r)
        for (unsigned p = m.row_offsets[r]; p < m.row_offsets[r + 1]; ++p) {
            if (transpose) res.data[m.col_ids[p]] += m.data[p] * v.data[r];
            else res.data[r] += m.data[p] * v.data[m.col_ids[p]];
        }
}

#define SPECK_INSTANTIATE(T)                                                              \
    template struct CSR<T>;                                                               \
    template CSR<T> loadCSR<T>(const char *);                                             \
    template void storeCSR<T>(const CSR<T> &, const char *);                              \
    template void convert(CSR<T> &, const COO<T> &);                                      \
    template void spmv(DenseVector<T> &, const CSR<T> &, const DenseVector<T> &, bool);
SPECK_INSTANTIATE(float)
SPECK_INSTANTIATE(double)
