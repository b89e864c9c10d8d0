// profiles/micro/gather_bw.cu -- what does B200 HBM/L2 deliver for the SpGEMM numeric access pattern?
// Random segments (SEG bytes, 16-byte aligned starts) gathered from a table of TABLE_MB, every byte read once per
// "product"; warps read whole segments coalesced.  Prints GB/s of gathered bytes (useful bytes).
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bw gather_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__global__ void k_gather(const uint4 *__restrict__ table, const uint32_t *__restrict__ segStart, uint32_t nSeg,
                         int vecPerSeg, unsigned long long *sink)
{
    // one sub-warp group of vecPerSeg lanes per segment (vecPerSeg in {2,4,8,16,32}: 32..512 bytes)
    const uint32_t lanesPer = vecPerSeg;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t groups = (gridDim.x * blockDim.x) / lanesPer;
    const uint32_t g = gtid / lanesPer, l = gtid % lanesPer;
    unsigned long long acc = 0;
    // 4 segments in flight per group
    for (uint32_t s = g; s < nSeg; s += 4 * groups) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t si = s + u * groups;
            v[u] = si < nSeg ? __ldg(table + segStart[si] + l) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x123456789abcdefull) *sink = acc;
}

int main(int argc, char **argv)
{
    const size_t tableMB[] = {64, 200, 400, 2048};
    const int vecs[] = {2, 4, 8, 16, 32};
    const uint32_t nSeg = 1u << 26;   // 64 M segments
    uint32_t *dSeg;
    unsigned long long *dSink;
    cudaMalloc(&dSeg, (size_t)nSeg * 4);
    cudaMalloc(&dSink, 8);
    std::vector<uint32_t> h(nSeg);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (size_t mb : tableMB) {
        uint4 *dT;
        const size_t vecsInTable = mb * 1024 * 1024 / 16;
        cudaMalloc(&dT, vecsInTable * 16);
        cudaMemset(dT, 1, vecsInTable * 16);
        for (int vp : vecs) {
            uint64_t x = 88172645463325252ull;
            for (uint32_t i = 0; i < nSeg; ++i) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                h[i] = (uint32_t)(x % (vecsInTable - vp));
            }
            const uint32_t use = (uint32_t)((size_t)nSeg * 2 / vp < nSeg ? (size_t)nSeg * 2 / vp : nSeg);   // ~2 GB of gathered bytes
            cudaMemcpy(dSeg, h.data(), (size_t)use * 4, cudaMemcpyHostToDevice);
            for (int threads : {256}) {
                const int blocks = 148 * (2048 / threads);
                k_gather<<<blocks, threads>>>(dT, dSeg, use, vp, dSink);
                cudaEventRecord(e0);
                for (int r = 0; r < 3; ++r) k_gather<<<blocks, threads>>>(dT, dSeg, use, vp, dSink);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                ms /= 3;
                const double bytes = (double)use * vp * 16;
                printf("table %5zu MB  segment %4d B  %8.1f GB/s useful  (%.3f ms, %u segments, +%.1f GB/s of start indices)\n", mb, vp * 16,
                       bytes / ms / 1e6, ms, use, (double)use * 4 / ms / 1e6);
            }
        }
        cudaFree(dT);
    }
    return 0;
}
