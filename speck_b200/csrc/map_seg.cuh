// speck_b200/csrc/map_seg.cuh -- segment-major numeric kernel of the mapped CTA classes (round 2).
//
// Replaces k_map_rows_cta (rank_cta.cuh) for rows of 513..16384 products; reference counterpart: the numeric
// hash accumulation + sorting, include/GPU/spECK_HashSpGEMM.cuh:591-866, 1856-1925.
//
// Why: k_map_rows_cta gives thread t the products [t*EP, (t+1)*EP) of the row, so one warp load touches 32
// different 32-byte sectors: ncu (profiles/r2_notes.md) shows 5.4 GB of L2->SM read traffic for 1.8 GB of DRAM reads
// in the 1024-product class -- the kernel is bound by sector traffic between L2 and the SMs, not by HBM.  Here a
// group of SG consecutive lanes walks ONE B-row segment (one A entry), so every load instruction of the group reads
// consecutive columns / values / rank codes; the A value is uniform in the group (no per-product owner search) and
// the entry's first product index comes from the analysis (aOff), so the kernel has no block scan and no tables:
// one barrier before the coalesced write-out.  Segments longer than SEG_LONG are processed by the whole CTA
// afterwards (a hub row of B would otherwise serialise one group).
// Products that are not the first of their column (rank-map bit 15) are parked in a small shared list and added
// once every first product is in place; rows that fold heavily (more parked products than the list holds) add
// every product into a zeroed staging row instead.
#pragma once
#include "common.cuh"

namespace sb {

constexpr int SEG_SG = 8;          // lanes per B-row segment
constexpr u32 SEG_LONG = 64;       // longer segments: whole CTA

template <int THREADS, int E, typename T>
struct MapSegLayout {
    static constexpr size_t al(size_t b) { return (b + 15) / 16 * 16; }
    static constexpr size_t CAP = (size_t)THREADS * E;
    static constexpr size_t DUPCAP = CAP / 8;           // parked products
    static constexpr size_t LONGCAP = CAP / SEG_LONG;   // long segments of one row (each has > SEG_LONG products)
    static constexpr size_t OUTVAL = 0;
    static constexpr size_t OUTCOL = OUTVAL + al(CAP * sizeof(T));
    static constexpr size_t DUPVAL = OUTCOL + al(CAP * 4);
    static constexpr size_t DUPRANK = DUPVAL + al(DUPCAP * sizeof(T));
    static constexpr size_t LONGLIST = DUPRANK + al(DUPCAP * 2);
    static constexpr size_t SMEM = LONGLIST + al(LONGCAP * 4);
};

template <int THREADS, int E, typename T>
__global__ void __launch_bounds__(THREADS, (1536 / THREADS) > 0 ? (1536 / THREADS) : 1)
k_map_seg(const RowDesc *__restrict__ desc, const uint2 *__restrict__ aSeg, const u32 *__restrict__ aOff,
          const T *__restrict__ aV, const u32 *__restrict__ bCi, const T *__restrict__ bV,
          const unsigned short *__restrict__ rankMap, u32 *__restrict__ cCi, T *__restrict__ cV)
{
    using L = MapSegLayout<THREADS, E, T>;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    T *outVal = reinterpret_cast<T *>(smemRaw + L::OUTVAL);
    u32 *outCol = reinterpret_cast<u32 *>(smemRaw + L::OUTCOL);
    T *dupVal = reinterpret_cast<T *>(smemRaw + L::DUPVAL);
    unsigned short *dupRank = reinterpret_cast<unsigned short *>(smemRaw + L::DUPRANK);
    u32 *longList = reinterpret_cast<u32 *>(smemRaw + L::LONGLIST);
    __shared__ u32 sDupCnt, sLongCnt;

    const u32 tid = threadIdx.x;
    const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(desc + blockIdx.x));
    const uint4 d1 = __ldg(reinterpret_cast<const uint4 *>(desc + blockIdx.x) + 1);
    const u32 aBeg = d0.x, aEnd = d0.x + d0.y, n = d0.z;
    const u32 cBase = d1.x, nnzRow = d1.y;
    const unsigned short *map = rankMap + (((u64)d1.w << 32) | d1.z);
    // parked-list mode while the row's repeated products fit the list, else every product is added into a zeroed row
    const bool addAll = (n - nnzRow) > (u32)L::DUPCAP;
    if (tid == 0) { sDupCnt = 0; sLongCnt = 0; }
    if (addAll)
        for (u32 j = tid; j < nnzRow; j += THREADS) outVal[j] = (T)0;
    __syncthreads();

    auto emit = [&](u32 code, u32 col, T pr) {
        const u32 r = code & MAP_RANK_MASK;
        if (addAll) {
            if (!(code & MAP_DUP)) outCol[r] = col;
            atomicAdd(&outVal[r], pr);
        } else if (code & MAP_DUP) {
            const u32 slot = atomicAdd(&sDupCnt, 1u);
            dupVal[slot] = pr;
            dupRank[slot] = (unsigned short)r;
        } else {
            outVal[r] = pr;
            outCol[r] = col;
        }
    };

    // ---------------------------------------------------------------- short segments: SG lanes per A entry
    constexpr u32 NG = THREADS / SEG_SG;
    const u32 g = tid / SEG_SG, gl = tid % SEG_SG;
#pragma unroll 1
    for (u32 e = aBeg + g; e < aEnd; e += NG) {
        const uint2 seg = __ldg(aSeg + e);
        const u32 len = seg.y - seg.x;
        if (len == 0) continue;
        if (len > SEG_LONG) {
            if (gl == 0) longList[atomicAdd(&sLongCnt, 1u)] = e;
            continue;
        }
        const T av = __ldg(aV + e);
        const u32 off = __ldg(aOff + e);
        const u32 *pc = bCi + seg.x;
        const T *pv = bV + seg.x;
        const unsigned short *pm = map + off;
        constexpr int U = 4;   // products per lane in flight
#pragma unroll 1
        for (u32 j0 = gl; j0 < len; j0 += U * SEG_SG) {
            u32 col[U], code[U];
            T bv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const u32 j = j0 + u * SEG_SG;
                const bool ok = j < len;
                col[u] = ok ? __ldg(pc + j) : 0u;
                bv[u] = ok ? __ldg(pv + j) : (T)0;
                code[u] = ok ? (u32)pm[j] : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (j0 + u * SEG_SG < len) emit(code[u], col[u], av * bv[u]);
        }
    }
    __syncthreads();

    // ---------------------------------------------------------------- long segments: the whole CTA per A entry
    const u32 nLong = sLongCnt;
#pragma unroll 1
    for (u32 li = 0; li < nLong; ++li) {
        const u32 e = longList[li];
        const uint2 seg = __ldg(aSeg + e);
        const u32 len = seg.y - seg.x;
        const T av = __ldg(aV + e);
        const u32 off = __ldg(aOff + e);
#pragma unroll 1
        for (u32 j = tid; j < len; j += THREADS) emit((u32)map[off + j], __ldg(bCi + seg.x + j), av * __ldg(bV + seg.x + j));
    }
    if (nLong) __syncthreads();   // CTA-uniform

    // ---------------------------------------------------------------- parked products
    if (!addAll) {
        const u32 nd = sDupCnt;
        if (nd) {
            for (u32 j = tid; j < nd; j += THREADS) atomicAdd(&outVal[dupRank[j]], dupVal[j]);
            __syncthreads();
        }
    }
#pragma unroll 1
    for (u32 j = tid; j < nnzRow; j += THREADS) {
        cCi[cBase + j] = outCol[j];
        cV[cBase + j] = outVal[j];
    }
}

template <int THREADS, int E, typename T>
void launch_map_seg_shape(const LaunchCtx &lc, const RowDesc *desc, u32 count, const uint2 *aSeg, const u32 *aOff,
                          const T *aV, const u32 *bCi, const T *bV, const unsigned short *rankMap, u32 *cCi, T *cV)
{
    using L = MapSegLayout<THREADS, E, T>;
    auto kern = k_map_seg<THREADS, E, T>;
    if (L::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    kern<<<count, THREADS, L::SMEM, lc.stream>>>(desc, aSeg, aOff, aV, bCi, bV, rankMap, cCi, cV);
    ++*lc.launches;
}

}  // namespace sb
