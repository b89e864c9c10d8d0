// include/Vector.h -- dense host vector used by the spmv helpers (reference include/Vector.h).
#pragma once
#include <cstddef>
#include <memory>

template <typename T>
struct DenseVector {
    size_t size = 0;
    std::unique_ptr<T[]> data;
    void alloc(size_t n)
    {
        data.reset(new T[n]());
        size = n;
    }
};
