// speck_b200/host/Executor.cpp -- warm-up + timed loops of runspECK with the reference's stdout
// contract (source/Executor.cpp:13-81): "Matrix: RxC: N nonzeros", optional cuSPARSE index check
// ("Error: Matrix incorrect"), " var-SpGEMM -> NNZ: n", " var-SpGEMM SpGEMM: t ms".
// New, after those lines: a GFLOPS line (2*P/t), which the reference never computes.
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <vector>
#include <cuda_runtime.h>
#include "Compare.h"
#include "Config.h"
#include "DataLoader.h"
#include "Executor.h"
#include "Multiply.h"
#include "cusparse/include/cuSparseMultiply.h"

namespace {

std::vector<int> parse_devices(const std::string &s)
{
    std::vector<int> out;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ','))
        if (!tok.empty()) out.push_back(std::atoi(tok.c_str()));
    return out;
}

speck_csr host_view(const CSR<float> &m) { return speck_csr{m.rows, m.cols, m.nnz, m.data.get(), m.row_offsets.get(), m.col_ids.get()}; }
speck_csr host_view(const CSR<double> &m) { return speck_csr{m.rows, m.cols, m.nnz, m.data.get(), m.row_offsets.get(), m.col_ids.get()}; }
int sharded_create(speck_ctx **c, int n, const speck_csr *a, const speck_csr *b, speck_shard_plan **p, float) { return speck_b200_sharded_create_f32(c, n, a, b, p); }
int sharded_create(speck_ctx **c, int n, const speck_csr *a, const speck_csr *b, speck_shard_plan **p, double) { return speck_b200_sharded_create_f64(c, n, a, b, p); }

// Devices=d0,d1,...: A is cut into product-balanced row slabs, one per device, B is replicated, every device
// multiplies its slab concurrently and the slabs are concatenated on the first device.  Same stdout
// lines as the single-device run, then one line per slab.
template <typename ValueType>
int run_sharded(const std::vector<int> &devices, Matrices<ValueType> &m, int warmup, int iterations, bool check)
{
    const int n = (int)devices.size();
    std::vector<speck_ctx *> ctx(n, nullptr);
    auto cleanup = [&](speck_shard_plan *plan) {
        speck_b200_sharded_destroy(plan);
        for (auto c : ctx) speck_b200_destroy(c);
    };
    for (int g = 0; g < n; ++g)
        if (speck_b200_create(devices[g], &ctx[g]) != SPECK_OK) {
            printf("ERROR: %s\n", speck_b200_last_error());
            cleanup(nullptr);
            return 1;
        }
    const speck_csr hA = host_view(m.cpuA), hB = host_view(m.cpuB);
    speck_shard_plan *plan = nullptr;
    if (sharded_create(ctx.data(), n, &hA, &hB, &plan, ValueType()) != SPECK_OK) {
        printf("ERROR: %s\n", speck_b200_last_error());
        cleanup(nullptr);
        return 1;
    }
    speck_shard_info info{};
    double sumMs = 0.0, sumConcat = 0.0;
    dCSR<ValueType> result, reference;
    if (check) {
        cudaSetDevice(devices[0]);
        uint32_t refNnz = 0;
        cuSPARSE::CuSparseTest<ValueType> cusparse;
        cusparse.Multiply(m.gpuA, m.gpuB, reference, refNnz);
        cudaFree(reference.data);
        reference.data = nullptr;
    }
    bool concatOk = true;
    for (int i = 0; i < warmup + iterations; ++i) {
        if (speck_b200_sharded_multiply(plan, &info) != SPECK_OK) {
            printf("ERROR: %s\n", speck_b200_last_error());
            cleanup(plan);
            return 1;
        }
        speck_csr c{result.rows, result.cols, result.nnz, result.data, result.row_offsets, result.col_ids};
        const int rc = speck_b200_sharded_concat(plan, &c, &info);
        if (rc == SPECK_OK) {
            result.rows = c.rows; result.cols = c.cols; result.nnz = c.nnz;
            result.data = (ValueType *)c.data; result.row_offsets = c.row_offsets; result.col_ids = c.col_ids;
            if (check && !spECK::Compare(reference, result, false)) printf("Error: Matrix incorrect\n");
        } else {
            concatOk = false;   // nnz(C) >= 2^32: C stays distributed
        }
        if (i >= warmup) { sumMs += info.ms_multiply; sumConcat += info.ms_concat; }
    }
    uint64_t nnz = 0, P = 0;
    for (int g = 0; g < n; ++g) { nnz += info.nnz_c[g]; P += info.products[g]; }
    const double ms = iterations > 0 ? sumMs / iterations : 0.0;
    std::cout << std::setw(20) << "var-SpGEMM -> NNZ: " << nnz << std::endl;
    std::cout << std::setw(20) << "var-SpGEMM SpGEMM: " << ms << " ms" << std::endl;
    if (ms > 0) std::cout << std::setw(20) << "var-SpGEMM GFLOPS: " << 2.0 * (double)P / (ms * 1e-3) / 1e9 << " (P = " << P << ", " << n << " devices)" << std::endl;
    if (concatOk && iterations > 0) std::cout << std::setw(20) << "var-SpGEMM concat: " << sumConcat / iterations << " ms (slabs -> device " << devices[0] << ")" << std::endl;
    else if (!concatOk) std::cout << "var-SpGEMM concat: skipped, nnz(C) does not fit u32 row_offsets; C stays distributed\n";
    for (int g = 0; g < n; ++g)
        std::cout << "  device " << devices[g] << ": rows [" << info.cuts[g] << ", " << info.cuts[g + 1] << ")  P = " << info.products[g]
                  << "  nnz(C) = " << info.nnz_c[g] << "  " << info.ms_device[g] << " ms" << std::endl;
    cudaSetDevice(devices[0]);
    cleanup(plan);
    return 0;
}

}  // namespace

template <typename ValueType>
int Executor<ValueType>::run()
{
    iterationsWarmup = Config::getInt(Config::IterationsWarmUp, 5);
    iterationsExecution = Config::getInt(Config::IterationsExecution, 10);
    const std::vector<int> devices = parse_devices(Config::getString(Config::Devices, ""));
    const int device = devices.empty() ? Config::getInt(Config::Device, 0) : devices[0];
    cudaSetDevice(device);
    DataLoader<ValueType> loader(runConfig.filePath);
    auto &m = loader.matrices;
    std::cout << "Matrix: " << m.cpuA.rows << "x" << m.cpuA.cols << ": " << m.cpuA.nnz << " nonzeros\n";
    if (devices.size() > 1)
        return run_sharded<ValueType>(devices, m, iterationsWarmup, iterationsExecution, Config::getBool(Config::CompareResult, false));

    dCSR<ValueType> result, reference;
    const bool trackStages = Config::getBool(Config::TrackIndividualTimes, false);
    const bool trackComplete = Config::getBool(Config::TrackCompleteTimes, true);
    const bool check = Config::getBool(Config::CompareResult, false);
    auto config = spECK::spECKConfig::initialize(device);

    if (check) {
        uint32_t refNnz = 0;
        cuSPARSE::CuSparseTest<ValueType> cusparse;
        cusparse.Multiply(m.gpuA, m.gpuB, reference, refNnz);
        cudaFree(reference.data);   // indices only, as in the reference (Executor.cpp:35-39)
        reference.data = nullptr;
    }

    Timings sum;
    auto iterate = [&](int n, bool accumulate) {
        for (int i = 0; i < n; ++i) {
            Timings t;
            t.measureAll = trackStages;
            t.measureCompleteTime = trackComplete;
            spECK::MultiplyspECK<ValueType, 4, 1024, spECK_DYNAMIC_MEM_PER_BLOCK, spECK_STATIC_MEM_PER_BLOCK>(
                m.gpuA, m.gpuB, result, config, t);
            if (accumulate) sum += t;
            if (check && result.data != nullptr && result.col_ids != nullptr && !spECK::Compare(reference, result, false))
                printf("Error: Matrix incorrect\n");
        }
    };
    iterate(iterationsWarmup, false);
    iterate(iterationsExecution, true);
    if (iterationsExecution > 0) sum /= (float)iterationsExecution;

    std::cout << std::setw(20) << "var-SpGEMM -> NNZ: " << result.nnz << std::endl;
    std::cout << std::setw(20) << "var-SpGEMM SpGEMM: " << sum.complete << " ms" << std::endl;
    speck_stats st{};
    if (config.b200 && speck_b200_get_stats(config.b200, &st) == SPECK_OK && sum.complete > 0)
        std::cout << std::setw(20) << "var-SpGEMM GFLOPS: " << 2.0 * (double)st.products / (sum.complete * 1e-3) / 1e9
                  << " (P = " << st.products << ")" << std::endl;
    config.cleanup();
    return 0;
}

template int Executor<double>::run();
template int Executor<float>::run();
