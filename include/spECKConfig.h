// include/spECKConfig.h -- per-device configuration object of the spECK API (reference
// include/spECKConfig.h:8-53): device limits, six streams, four events.  New: the handle of the
// B200 library context that owns the pooled workspace; it is created by initialize() and
// released by cleanup().
#pragma once
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#include "speck_b200.h"

namespace spECK {

struct spECKConfig {
    int sm = 0;
    int maxStaticSharedMemoryPerBlock = 0;
    int maxDynamicSharedMemoryPerBlock = 0;
    std::vector<cudaStream_t> streams;
    cudaEvent_t completeStart = 0, completeEnd = 0, individualStart = 0, individualEnd = 0;
    speck_ctx *b200 = nullptr;   // library context (streams, pooled workspace) for this device
    int device = 0;

    static spECKConfig initialize(int cudaDeviceNumber)
    {
        spECKConfig c;
        c.device = cudaDeviceNumber;
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, cudaDeviceNumber);
        c.sm = prop.multiProcessorCount;
        c.maxStaticSharedMemoryPerBlock = (int)prop.sharedMemPerBlock;
        c.maxDynamicSharedMemoryPerBlock = (int)std::max(prop.sharedMemPerBlockOptin, prop.sharedMemPerBlock);
        c.streams.resize(6, nullptr);
        for (auto &s : c.streams) cudaStreamCreate(&s);
        cudaEventCreate(&c.completeStart);
        cudaEventCreate(&c.completeEnd);
        cudaEventCreate(&c.individualStart);
        cudaEventCreate(&c.individualEnd);
        speck_b200_create(cudaDeviceNumber, &c.b200);
        return c;
    }

    void cleanup()
    {
        for (auto s : streams) cudaStreamDestroy(s);
        streams.clear();
        cudaEventDestroy(completeStart);
        cudaEventDestroy(completeEnd);
        cudaEventDestroy(individualStart);
        cudaEventDestroy(individualEnd);
        if (b200) speck_b200_destroy(b200);
        b200 = nullptr;
    }

private:
    spECKConfig() = default;
};

}  // namespace spECK
