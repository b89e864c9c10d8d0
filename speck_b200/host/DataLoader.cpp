// speck_b200/host/DataLoader.cpp -- matrix loading for runspECK (reference source/DataLoader.cpp:24-75):
// try "<path>d_.hicsr" (fp64) / "<path>.hicsr" (fp32), else parse the .mtx, convert, write the cache;
// upload A; B = A for square inputs, B = A^T (cuSPARSE csr2csc) otherwise.  Same stdout lines.
#include <exception>
#include <iostream>
#include <type_traits>
#include <cuda_runtime.h>
#include "COO.h"
#include "Config.h"
#include "DataLoader.h"
#include "speck_b200.h"
#include "cusparse/include/cuSparseMultiply.h"

namespace {

int coo_to_csr_dev(speck_ctx *c, size_t r, size_t cl, size_t n, const unsigned *dr, const unsigned *dc, const float *dv, speck_csr *o)
{
    return speck_b200_coo_to_csr_f32(c, r, cl, n, dr, dc, dv, SPECK_COO_KEEP, o);
}
int coo_to_csr_dev(speck_ctx *c, size_t r, size_t cl, size_t n, const unsigned *dr, const unsigned *dc, const double *dv, speck_csr *o)
{
    return speck_b200_coo_to_csr_f64(c, r, cl, n, dr, dc, dv, SPECK_COO_KEEP, o);
}

// GpuConvert=true: the (row, column) sort of the freshly parsed triplets runs on the device
// (speck_b200_coo_to_csr_*, duplicates kept like the host conversion) and the CSR comes back for the .hicsr cache.
template <typename T>
bool convert_on_gpu(CSR<T> &dst, const COO<T> &src)
{
    speck_ctx *ctx = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    if (speck_b200_create(dev, &ctx) != SPECK_OK) return false;
    unsigned *dr = nullptr, *dc = nullptr;
    T *dv = nullptr;
    speck_csr out{};
    bool ok = cudaMalloc((void **)&dr, (src.nnz + 1) * 4) == cudaSuccess && cudaMalloc((void **)&dc, (src.nnz + 1) * 4) == cudaSuccess &&
              cudaMalloc((void **)&dv, (src.nnz + 1) * sizeof(T)) == cudaSuccess;
    if (ok && src.nnz) {
        cudaMemcpy(dr, src.row_ids.get(), src.nnz * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dc, src.col_ids.get(), src.nnz * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dv, src.data.get(), src.nnz * sizeof(T), cudaMemcpyHostToDevice);
    }
    ok = ok && coo_to_csr_dev(ctx, src.rows, src.cols, src.nnz, dr, dc, dv, &out) == SPECK_OK;
    if (ok) {
        dst.alloc(src.rows, src.cols, out.nnz);
        cudaMemcpy(dst.row_offsets.get(), out.row_offsets, (src.rows + 1) * 4, cudaMemcpyDeviceToHost);
        if (out.nnz) {
            cudaMemcpy(dst.col_ids.get(), out.col_ids, out.nnz * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(dst.data.get(), out.data, out.nnz * sizeof(T), cudaMemcpyDeviceToHost);
        }
        speck_b200_free_csr(ctx, &out);
    }
    cudaFree(dr); cudaFree(dc); cudaFree(dv);
    speck_b200_destroy(ctx);
    return ok;
}

}  // namespace

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
            if (Config::getBool(Config::GpuConvert, false) && convert_on_gpu(matrices.cpuA, coo))
                std::cout << coo.nnz << std::endl;   // the line the host conversion prints (CSR.cpp:189)
            else
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
