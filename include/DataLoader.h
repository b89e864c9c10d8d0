// include/DataLoader.h -- loads <matrix>.mtx (or its .hicsr cache), uploads A, builds B = A
// (square) or B = A^T (reference include/DataLoader.h, source/DataLoader.cpp:24-75).
#pragma once
#include <string>
#include "CSR.h"
#include "dCSR.h"

template <typename ValueType>
struct Matrices {
    CSR<ValueType> cpuA, cpuB;
    dCSR<ValueType> gpuA, gpuB;
};

template <typename ValueType>
class DataLoader {
public:
    explicit DataLoader(std::string path);
    Matrices<ValueType> matrices;
};
