// speck_b200/host/Executor.cpp -- warm-up + timed loops of runspECK with the reference's stdout
// contract (source/Executor.cpp:13-81): "Matrix: RxC: N nonzeros", optional cuSPARSE index check
// ("Error: Matrix incorrect"), " var-SpGEMM -> NNZ: n", " var-SpGEMM SpGEMM: t ms".
// New, after those lines: a GFLOPS line (2*P/t), which the reference never computes.
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <cuda_runtime.h>
#include "Compare.h"
#include "Config.h"
#include "DataLoader.h"
#include "Executor.h"
#include "Multiply.h"
#include "cusparse/include/cuSparseMultiply.h"

template <typename ValueType>
int Executor<ValueType>::run()
{
    iterationsWarmup = Config::getInt(Config::IterationsWarmUp, 5);
    iterationsExecution = Config::getInt(Config::IterationsExecution, 10);
    const int device = Config::getInt(Config::Device, 0);
    cudaSetDevice(device);
    DataLoader<ValueType> loader(runConfig.filePath);
    auto &m = loader.matrices;
    std::cout << "Matrix: " << m.cpuA.rows << "x" << m.cpuA.cols << ": " << m.cpuA.nnz << " nonzeros\n";

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
