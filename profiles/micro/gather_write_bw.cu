// profiles/micro/gather_write_bw.cu -- ceiling of the numeric phase's traffic mix on B200: random segments gathered
// from a 200 MB table (the size of B in config #2: L2 holds part of it) while the same kernel streams its results to
// a large output array (C, write-once).  Modes: gather only, write only, gather + write, gather + write(.cs).
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_write_bw gather_write_bw.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

template <int MODE>   // 0 gather, 1 write, 2 gather + write, 3 gather + write.cs
__global__ void k_mix(const uint4 *__restrict__ table, const uint32_t *__restrict__ segStart, uint32_t nSeg, int vecPerSeg,
                      uint4 *__restrict__ out, unsigned long long *sink)
{
    const uint32_t lanesPer = vecPerSeg;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t groups = (gridDim.x * blockDim.x) / lanesPer;
    const uint32_t g = gtid / lanesPer, l = gtid % lanesPer;
    unsigned long long acc = 0;
    for (uint32_t s = g; s < nSeg; s += 4 * groups) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t si = s + u * groups;
            v[u] = make_uint4(si, l, 0, 0);
            if (MODE != 1 && si < nSeg) v[u] = __ldg(table + segStart[si] + l);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t si = s + u * groups;
            if (MODE == 0) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
            else if (si < nSeg) {
                uint4 *dst = out + (size_t)si * lanesPer + l;   // consecutive groups write consecutive 16-byte vectors
                if (MODE == 3) __stcs(dst, v[u]); else *dst = v[u];
            }
        }
    }
    if (MODE == 0 && acc == 0x123456789abcdefull) *sink = acc;
}

int main()
{
    const size_t tableMB = 200;
    const size_t vecsInTable = tableMB * 1024 * 1024 / 16;
    uint4 *dT, *dOut;
    uint32_t *dSeg;
    unsigned long long *dSink;
    cudaMalloc(&dT, vecsInTable * 16);
    cudaMemset(dT, 1, vecsInTable * 16);
    const size_t outBytes = (size_t)4 << 30;
    cudaMalloc(&dOut, outBytes);
    cudaMalloc(&dSink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int vp : {4, 8, 16}) {
        const uint32_t nSeg = (uint32_t)(outBytes / ((size_t)vp * 16));
        std::vector<uint32_t> h(nSeg);
        uint64_t x = 88172645463325252ull;
        for (uint32_t i = 0; i < nSeg; ++i) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            h[i] = (uint32_t)(x % (vecsInTable - vp));
        }
        cudaMalloc(&dSeg, (size_t)nSeg * 4);
        cudaMemcpy(dSeg, h.data(), (size_t)nSeg * 4, cudaMemcpyHostToDevice);
        const int threads = 256, blocks = 148 * 8;
        for (int mode = 0; mode < 4; ++mode) {
            auto run = [&]() {
                if (mode == 0) k_mix<0><<<blocks, threads>>>(dT, dSeg, nSeg, vp, dOut, dSink);
                if (mode == 1) k_mix<1><<<blocks, threads>>>(dT, dSeg, nSeg, vp, dOut, dSink);
                if (mode == 2) k_mix<2><<<blocks, threads>>>(dT, dSeg, nSeg, vp, dOut, dSink);
                if (mode == 3) k_mix<3><<<blocks, threads>>>(dT, dSeg, nSeg, vp, dOut, dSink);
            };
            run();
            cudaEventRecord(e0);
            for (int r = 0; r < 3; ++r) run();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            ms /= 3;
            const double gb = (double)nSeg * vp * 16 / 1e9;
            const char *names[] = {"gather only        ", "write only         ", "gather + write     ", "gather + write(.cs)"};
            const double moved = (mode == 0 || mode == 1) ? gb : 2 * gb;
            printf("segment %4d B  %s  %7.3f ms  %8.1f GB/s (gathered + written bytes)\n", vp * 16, names[mode], ms, moved / ms * 1e3);
        }
        cudaFree(dSeg);
    }
    return 0;
}
