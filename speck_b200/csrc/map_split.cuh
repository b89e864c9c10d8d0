// speck_b200/csrc/map_split.cuh -- numeric phase of the LARGE mapped rows (4097 .. 16384 products), several CTAs
// per row (round 2).
//
// k_map_rows_cta stages a whole row of C in shared memory: 12 bytes per slot, so the 1024-thread shapes of these
// rows hold 98 / 196 KB and run ONE CTA per SM -- while that CTA scans, waits at a barrier or writes its row out,
// nothing on the SM is gathering (ncu: 1.1 - 2.3 TB/s of DRAM traffic against 3.4 - 4.3 TB/s for the small shapes,
// profiles/r2/final_launches_rmat20.txt).  Here SPLIT consecutive CTAs share a row: CTA h stages the output range
// [h * CAP, (h + 1) * CAP) of the row only (CAP = THREADS * E slots).  Every CTA walks the row's rank codes (2 bytes
// per product, the second CTA finds them in L2), keeps the products whose recorded position falls into its range in
// a bit mask, and gathers / multiplies / scatters only those, GU at a time, so the loads in flight per thread do not
// drop with SPLIT.  The smaller staging rows let 2 - 3 CTAs of different rows overlap their phases on one SM.
// Replaces the same reference code as k_map_rows_cta (include/GPU/spECK_HashSpGEMM.cuh:591-866).
#pragma once
#include "rank_cta.cuh"

namespace sb {

template <int THREADS, int E, typename T, int SPLIT>
struct MapSplitLayout {
    static constexpr size_t al(size_t b) { return (b + 15) / 16 * 16; }
    static constexpr size_t CAP = (size_t)THREADS * E;       // staged positions per CTA
    static constexpr size_t NMAX = CAP * SPLIT;               // products of a row
    static constexpr size_t OUTVAL = 0;
    static constexpr size_t SAV = OUTVAL + al(CAP * sizeof(T));
    static constexpr size_t OUTCOL = SAV + al(THREADS * sizeof(T));
    static constexpr size_t SINCL = OUTCOL + al(CAP * 4);
    static constexpr size_t SBS = SINCL + al(THREADS * 4);
    static constexpr size_t STAB = SBS + al(THREADS * 4);
    static constexpr size_t SMEM = STAB + al(NMAX / 32 * 2);
    static constexpr int CTAS_PER_SM = (int)((227 * 1024) / (SMEM + 1024)) < 1536 / THREADS
                                           ? (int)((227 * 1024) / (SMEM + 1024)) : 1536 / THREADS;
};

template <int THREADS, int E, typename T, int SPLIT>
__global__ void __launch_bounds__(THREADS, MapSplitLayout<THREADS, E, T, SPLIT>::CTAS_PER_SM)
k_map_rows_split(const RowDesc *__restrict__ desc, const uint2 *__restrict__ aSeg, const T *__restrict__ aV,
                 const u32 *__restrict__ bCi, const T *__restrict__ bV, const unsigned short *__restrict__ rankMap,
                 u32 *__restrict__ cCi, T *__restrict__ cV)
{
    using L = MapSplitLayout<THREADS, E, T, SPLIT>;
    static_assert(E * SPLIT <= 32, "a thread's products must fit one 32-bit mask");
    constexpr u32 NONE = 0xffffffffu;
    constexpr u32 CAP = (u32)L::CAP;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    T *outVal = reinterpret_cast<T *>(smemRaw + L::OUTVAL);
    T *sAv = reinterpret_cast<T *>(smemRaw + L::SAV);
    u32 *outCol = reinterpret_cast<u32 *>(smemRaw + L::OUTCOL);
    u32 *sIncl = reinterpret_cast<u32 *>(smemRaw + L::SINCL);
    u32 *sBs = reinterpret_cast<u32 *>(smemRaw + L::SBS);
    unsigned short *sTab = reinterpret_cast<unsigned short *>(smemRaw + L::STAB);
    __shared__ u32 sWarp[33];

    const u32 tid = threadIdx.x;
    const u32 rowIdx = blockIdx.x / SPLIT, h = blockIdx.x % SPLIT;
    const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(desc + rowIdx));
    const uint4 d1 = __ldg(reinterpret_cast<const uint4 *>(desc + rowIdx) + 1);
    const u32 aBeg = d0.x, aEnd = d0.x + d0.y, n = d0.z;
    const u32 nnzRow = d1.y;
    const u32 rLo = h * CAP;                      // first position of the row staged by this CTA
    if (rLo >= nnzRow) return;                    // CTA-uniform: short rows of the class need fewer CTAs
    const u32 rCnt = min(nnzRow - rLo, CAP);
    const u32 cBase = d1.x + rLo;
    const unsigned short *map = rankMap + (((u64)d1.w << 32) | d1.z);
    const u32 EP = (n + THREADS - 1) / THREADS;   // products per thread (<= E * SPLIT), blocked assignment
    const u32 myFirst = tid * EP;

    // as in k_map_rows_cta: rows whose A entries fit one batch place the first product of a column with plain stores
    // and add the others afterwards; rows with several batches add everything into a zeroed staging row
    const bool multi = (aEnd - aBeg) > (u32)THREADS;
    if (multi)
        for (u32 j = tid; j < rCnt; j += THREADS) outVal[j] = (T)0;
    u32 dup = 0;
    u32 base = 0;
#pragma unroll 1
    for (u32 ab = aBeg; ab < aEnd; ab += THREADS) {
        const u32 nb = min((u32)THREADS, aEnd - ab);
        u32 bs = 0, len = 0;
        if (tid < nb) {
            const uint2 seg = __ldg(aSeg + ab + tid);
            bs = seg.x;
            len = seg.y - seg.x;
            sAv[tid] = __ldg(aV + ab + tid);
        }
        u32 total;
        const u32 excl = cta_exclusive_scan<THREADS>(len, sWarp, &total);
        sIncl[tid] = excl + len;
        sBs[tid] = bs - excl;  // q = sBs[owner] + p
        if (len) {             // owner table: sTab[b] = entry owning product 32*b of this batch
            const u32 bLast = (excl + len - 1) >> 5;
            for (u32 b = (excl + 31) >> 5; b <= bLast; ++b) sTab[b] = (unsigned short)tid;
        }
        __syncthreads();
        constexpr int GU = 4;   // products per thread in flight
        const u32 first = max(myFirst, base), last = min(min(myFirst + EP, n), base + total);
        // this thread's products of the batch whose position lies in [rLo, rLo + rCnt)
        u32 want = 0;
        if (first < last) {
#pragma unroll 4
            for (u32 gp = first; gp < last; ++gp) {
                const u32 r = (u32)map[gp] & MAP_RANK_MASK;
                want |= (r - rLo < rCnt ? 1u : 0u) << (gp - myFirst);
            }
        }
        SegWalk<T, true> walk;
        if (want) walk.start(myFirst + (u32)(__ffs(want) - 1) - base, sTab, sIncl, sBs, sAv);
        while (want) {
            u32 q[GU], cc[GU], code[GU], idx[GU];
            T av[GU], bv[GU];
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                q[u] = NONE;
                av[u] = (T)0;
                code[u] = 0;
                idx[u] = 0;
                if (want) {
                    idx[u] = (u32)(__ffs(want) - 1);
                    want &= want - 1;
                    const u32 gp = myFirst + idx[u];
                    code[u] = map[gp];
                    q[u] = walk.locate(gp - base, sIncl, sBs, sAv);
                    av[u] = walk.av;
                }
            }
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                cc[u] = q[u] != NONE ? __ldg(bCi + q[u]) : 0u;
                bv[u] = q[u] != NONE ? __ldg(bV + q[u]) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                if (q[u] != NONE) {
                    const u32 r = (code[u] & MAP_RANK_MASK) - rLo;
                    const T pr = av[u] * bv[u];
                    if (multi) {
                        if (!(code[u] & MAP_DUP)) outCol[r] = cc[u];
                        atomicAdd(&outVal[r], pr);
                    } else if (code[u] & MAP_DUP) {
                        dup |= 1u << idx[u];
                    } else {
                        outVal[r] = pr;
                        outCol[r] = cc[u];
                    }
                }
            }
        }
        base += total;
        __syncthreads();
    }
    if (__syncthreads_or(dup != 0)) {   // single batch: its tables are still in place
        while (dup) {
            const u32 gp = myFirst + (u32)(__ffs(dup) - 1);
            dup &= dup - 1;
            SegWalk<T, true> walk;
            walk.start(gp, sTab, sIncl, sBs, sAv);
            atomicAdd(&outVal[((u32)map[gp] & MAP_RANK_MASK) - rLo], walk.av * __ldg(bV + walk.segBs + gp));
        }
        __syncthreads();
    }
#pragma unroll 1
    for (u32 j = tid; j < rCnt; j += THREADS) {
        cCi[cBase + j] = outCol[j];
        cV[cBase + j] = outVal[j];
    }
}

template <int THREADS, int E, typename T, int SPLIT>
void launch_map_rows_split(const LaunchCtx &lc, const RowDesc *desc, u32 count, const uint2 *aSeg, const T *aV,
                           const u32 *bCi, const T *bV, const unsigned short *rankMap, u32 *cCi, T *cV)
{
    using L = MapSplitLayout<THREADS, E, T, SPLIT>;
    auto kern = k_map_rows_split<THREADS, E, T, SPLIT>;
    if (L::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    kern<<<count * SPLIT, THREADS, L::SMEM, lc.stream>>>(desc, aSeg, aV, bCi, bV, rankMap, cCi, cV);
    ++*lc.launches;
}

}  // namespace sb
