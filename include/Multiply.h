// include/Multiply.h -- public entry point of the spECK API, same signature as the reference's
// include/Multiply.h:10-19.  The implementation (speck_b200/host/Multiply.cpp) forwards to the
// C ABI speck_b200_spgemm_{f32,f64}; the library exports the same explicit instantiations as the
// reference (source/GPU/Multiply.cu:1130-1131): <float|double, 4, 1024, DYN, STATIC>.
#pragma once
#include "Timings.h"
#include "dCSR.h"
#include "spECKConfig.h"

// Kept for source compatibility with callers that spell the template arguments (Executor.cpp:48).
// The B200 kernels size their own shared memory (up to the 227 KB opt-in limit); these two
// constants no longer select anything and never trigger the reference's per-call WARNING.
static constexpr int spECK_STATIC_MEM_PER_BLOCK{49152};
static constexpr int spECK_DYNAMIC_MEM_PER_BLOCK{49152};

namespace spECK {

template <typename DataType, int BLOCKS_PER_SM, int THREADS_PER_BLOCK, int MAX_DYNAMIC_SHARED, int MAX_STATIC_SHARED>
void MultiplyspECK(const dCSR<DataType> &A, const dCSR<DataType> &B, dCSR<DataType> &matOut, spECKConfig &config,
                   Timings &timings);

template <typename DataType, int BLOCKS_PER_SM, int THREADS_PER_BLOCK, int MAX_DYNAMIC_SHARED, int MAX_STATIC_SHARED>
void MultiplyspECKImplementation(const dCSR<DataType> &A, const dCSR<DataType> &B, dCSR<DataType> &matOut,
                                 spECKConfig &config, Timings &timings);

}  // namespace spECK
