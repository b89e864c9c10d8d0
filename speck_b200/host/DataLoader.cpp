// speck_b200/host/DataLoader.cpp -- matrix loading for runspECK (reference source/DataLoader.cpp:24-75):
// try "<path>d_.hicsr" (fp64) / "<path>.hicsr" (fp32), else parse the .mtx, convert, write the cache;
// upload A; B = A for square inputs, B = A^T (cuSPARSE csr2csc) otherwise.  Same stdout lines.
#include <exception>
#include <iostream>
#include <type_traits>
#include "COO.h"
#include "DataLoader.h"
#include "cusparse/include/cuSparseMultiply.h"

template <typename ValueType>
DataLoader<ValueType>::DataLoader(std::string path)
{
    const std::string csrPath = path + (std::is_same<ValueType, double>::value ? "d_" : "") + ".hicsr";
    try {
        std::cout << "trying to load csr file \"" << csrPath << "\"\n";
        matrices.cpuA = loadCSR<ValueType>(csrPath.c_str());
        std::cout << "successfully loaded: \"" << csrPath << "\"\n";
    } catch (std::exception &ex) {
        std::cout << "could not load csr file:\n\t" << ex.what() << "\n";
        try {
            std::cout << "trying to load mtx file \"" << path << "\"\n";
            COO<ValueType> coo = loadMTX<ValueType>(path.c_str());
            convert(matrices.cpuA, coo);
            std::cout << "successfully loaded and converted: \"" << csrPath << "\"\n";
        } catch (std::exception &ex2) {
            std::cout << ex2.what() << std::endl;
            std::cout << "could not load mtx file: \"" << path << "\"\n";
            throw "could not load mtx file";
        }
        try {
            std::cout << "write csr file for future use\n";
            storeCSR(matrices.cpuA, csrPath.c_str());
        } catch (std::exception &ex3) {
            std::cout << ex3.what() << std::endl;
        }
    }
    convert(matrices.gpuA, matrices.cpuA, 0);
    if (matrices.gpuA.rows != matrices.gpuA.cols) {
        cuSPARSE::CuSparseTest<ValueType> cusparse;
        cusparse.Transpose(matrices.gpuA, matrices.gpuB);
        convert(matrices.cpuB, matrices.gpuB);
    } else {
        convert(matrices.gpuB, matrices.cpuA, 0);
        convert(matrices.cpuB, matrices.cpuA, 0);
    }
}

template class DataLoader<float>;
template class DataLoader<double>;
