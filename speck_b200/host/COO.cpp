// speck_b200/host/COO.cpp -- COO container + MatrixMarket reader with the semantics of the
// reference's loadMTX (source/COO.cpp:52-164).
#include <algorithm>
#include <cctype>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "COO.h"

template <typename T>
void COO<T>::alloc(size_t r, size_t c, size_t n)
{
    rows = r;
    cols = c;
    nnz = n;
    data.reset(new T[n ? n : 1]());
    row_ids.reset(new unsigned int[n ? n : 1]());
    col_ids.reset(new unsigned int[n ? n : 1]());
}

namespace {
std::string lower(std::string s)
{
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}
}  // namespace

template <typename T>
COO<T> loadMTX(const char *file)
{
    std::ifstream in(file);
    if (!in.is_open()) throw std::runtime_error(std::string("could not open \"") + file + "\"");
    std::string line;
    if (!std::getline(in, line)) throw std::runtime_error("empty MatrixMarket file");
    std::istringstream banner(line);
    std::string tag, object, format, field, symmetry;
    banner >> tag >> object >> format >> field >> symmetry;
    if (tag != "%%MatrixMarket" || lower(object) != "matrix") throw std::runtime_error("not a MatrixMarket matrix file");
    if (lower(format) != "coordinate") throw std::runtime_error("only coordinate MatrixMarket files are supported");
    field = lower(field);
    symmetry = lower(symmetry);
    const bool pattern = field == "pattern";
    const bool complex = field == "complex";
    if (!(pattern || complex || field == "real" || field == "integer" || field == "double"))
        throw std::runtime_error("unsupported MatrixMarket field: " + field);
    bool mirror;
    if (symmetry == "general") mirror = false;
    else if (symmetry == "symmetric" || symmetry == "hermitian") mirror = true;
    else throw std::runtime_error("unsupported MatrixMarket symmetry: " + symmetry);

    do {
        if (!std::getline(in, line)) throw std::runtime_error("MatrixMarket size line missing");
    } while (line.empty() || line[0] == '%');
    size_t rows = 0, cols = 0, entries = 0;
    {
        std::istringstream sz(line);
        sz >> rows >> cols >> entries;
        if (sz.fail()) throw std::runtime_error("malformed MatrixMarket size line");
    }
    std::vector<unsigned> r, c;
    std::vector<T> v;
    r.reserve(mirror ? 2 * entries : entries);
    c.reserve(r.capacity());
    v.reserve(r.capacity());
    for (size_t i = 0; i < entries; ++i) {
        do {
            if (!std::getline(in, line)) throw std::runtime_error("unexpected end of MatrixMarket file");
        } while (line.empty() || line[0] == '%');
        std::istringstream e(line);
        long long ri = 0, ci = 0;
        double val = 1.0, imag = 0.0;
        e >> ri >> ci;
        if (!pattern) e >> val;
        if (complex) e >> imag;  // only the real part is kept
        if (e.fail()) throw std::runtime_error("malformed MatrixMarket entry: " + line);
        if (ri < 1 || ci < 1 || (size_t)ri > rows || (size_t)ci > cols)
            throw std::runtime_error("MatrixMarket entry out of bounds: " + line);
        r.push_back((unsigned)(ri - 1));
        c.push_back((unsigned)(ci - 1));
        v.push_back((T)val);
        if (mirror && ri != ci) {
            r.push_back((unsigned)(ci - 1));
            c.push_back((unsigned)(ri - 1));
            v.push_back((T)val);
        }
    }
    COO<T> m;
    m.alloc(rows, cols, v.size());
    std::copy(r.begin(), r.end(), m.row_ids.get());
    std::copy(c.begin(), c.end(), m.col_ids.get());
    std::copy(v.begin(), v.end(), m.data.get());
    return m;
}

template <typename T>
void spmv(DenseVector<T> &res, const COO<T> &m, const DenseVector<T> &v, bool transpose)
{
    if ((transpose ? m.rows : m.cols) != v.size) throw std::runtime_error("SPMV dimensions mismatch");
    res.alloc(transpose ? m.cols : m.rows);
    for (size_t i = 0; i < m.nnz; ++i) {
        if (transpose) res.data[m.col_ids[i]] += m.data[i] * v.data[m.row_ids[i]];
        else res.data[m.row_ids[i]] += m.data[i] * v.data[m.col_ids[i]];
    }
}

#define SPECK_INSTANTIATE(T)                                                              \
    template struct COO<T>;                                                               \
    template COO<T> loadMTX<T>(const char *);                                             \
    template void spmv(DenseVector<T> &, const COO<T> &, const DenseVector<T> &, bool);
SPECK_INSTANTIATE(float)
SPECK_INSTANTIATE(double)
