// speck_b200/csrc/rank_cta.cuh -- "rank" row classes: one CTA per row of C, sorted output positions
// from a two-level column bitmap instead of a sort.
//
// Replaces, for rows of 513 .. 8192 products, the shared-memory hash accumulation + rank/radix sort
// of the reference (hashSpGEMMNumericImplementation + hashSpGEMMSortingKernel,
// include/GPU/spECK_HashSpGEMM.cuh:591-866, 1856-1925) and its symbolic twin (:919-1119).
//
// The row's products stay in registers (E per thread, flat enumeration as in sort_cta.cuh).  A
// column, relative to the row's smallest column, is split as
//     c = [ top word (<= 10 bits) | bit in top word (5) | bit in leaf word (5) ]      extent <= 2^20
//   top level  : dense array of <= 1024 words; bit set <=> that 32-column chunk holds a product
//   leaf level : one 32-column word per TOUCHED chunk only, at slot = popcount-rank of the chunk's
//                top-level bit (prefix popcounts from a block scan): storage is O(products), not
//                O(extent), and so is the clearing / scanning work
// After the leaf words are scanned, the sorted position of a product is
//     leafPre[slot] + popc(leaf[slot] below the product's bit)
// i.e. O(1) per product instead of the O(log^2 n) compare-exchanges of the bitonic classes (~350
// instructions per product on rows of this size, profiles/r1_notes.md).  The product that sets a
// column's leaf bit first owns the output slot (plain stores of column and value into a shared staging
// row); further products of the same column are added afterwards (shared-memory atomicAdd, rare when the
// row does not compress); the staged row is then written to C coalesced.
#pragma once
#include "common.cuh"

namespace sb {

constexpr int RANK_TOP_WORDS = 1024;
constexpr u32 RANK_MAX_EXTENT = 1u << 20;   // RANK_TOP_WORDS * 32 * 32 columns

// exclusive scan of one u32 per thread over a CTA of THREADS threads; sWarp: 33 words.
// Contains two barriers; *total = block sum.
template <int THREADS>
__device__ __forceinline__ u32 cta_exclusive_scan(u32 v, u32 *sWarp, u32 *total)
{
    constexpr u32 NW = THREADS / 32;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (u32)d) incl += t;
    }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const u32 wv = lane < NW ? sWarp[lane] : 0u;
        u32 winc = wv;
#pragma unroll
        for (int d = 1; d < (int)NW; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (u32)d) winc += t;
        }
        if (lane < NW) sWarp[lane] = winc - wv;
        if (lane == NW - 1) sWarp[32] = winc;
    }
    __syncthreads();
    *total = sWarp[32];
    return sWarp[warp] + incl - v;
}

// Shared memory of a CTA of THREADS threads (CAP = THREADS * E product slots), 16-byte aligned pieces, all
// offsets compile-time:
//   outVal[CAP] T | sAv[THREADS] T | top[1024] | topPre[1024] u16 | leaf[CAP] | leafPre[CAP] u16 |
//   leafId[CAP] u16 | sIncl[THREADS] | sBs[THREADS] | sTab[CAP/32] u16 | outCol[CAP] (numeric; aliased onto
//   top/topPre when it fits: both are dead once the leaf words are filled)
template <int THREADS, int E, typename T, bool NUMERIC>
struct RankLayout {
    static constexpr size_t al(size_t b) { return (b + 15) / 16 * 16; }
    static constexpr size_t CAP = (size_t)THREADS * E;
    static constexpr bool COL_ALIASED = CAP * 4 <= RANK_TOP_WORDS * 6;
    static constexpr size_t OUTVAL = 0;
    static constexpr size_t SAV = OUTVAL + (NUMERIC ? al(CAP * sizeof(T)) : 0);
    static constexpr size_t TOP = SAV + (NUMERIC ? al(THREADS * sizeof(T)) : 0);
    static constexpr size_t TOPPRE = TOP + RANK_TOP_WORDS * 4;
    static constexpr size_t LEAF = TOPPRE + RANK_TOP_WORDS * 2;
    static constexpr size_t LEAFPRE = LEAF + al(CAP * 4);
    static constexpr size_t LEAFID = LEAFPRE + al(CAP * 2);
    static constexpr size_t SINCL = LEAFID + (NUMERIC ? al(CAP * 2) : 0);
    static constexpr size_t SBS = SINCL + al(THREADS * 4);
    static constexpr size_t STAB = SBS + al(THREADS * 4);
    static constexpr size_t OUTCOL = (NUMERIC && COL_ALIASED) ? TOP : STAB + al(CAP / 32 * 2);
    static constexpr size_t SMEM = STAB + al(CAP / 32 * 2) + ((NUMERIC && !COL_ALIASED) ? CAP * 4 : 0);
};

// popcount prefix of a level: words[0..n) (padded with zero words to a multiple of 4 by the caller)
// -> pre[j] = number of set bits in words[0..j), returns the total.  Threads own consecutive groups of four
// words (one 16-byte load, one 8-byte store of four packed u16 prefixes).
template <int THREADS, bool WRITE_PREFIX>
__device__ __forceinline__ u32 rank_scan_level(const u32 *bits, unsigned short *pre, u32 words, u32 *sWarp)
{
    const u32 tid = threadIdx.x;
    const u32 groups = (words + 3) >> 2;
    const u32 per = (groups + THREADS - 1) / THREADS;
    const u32 g0 = min(groups, tid * per), g1 = min(groups, g0 + per);
    u32 s = 0;
#pragma unroll 1
    for (u32 g = g0; g < g1; ++g) {
        const uint4 v = reinterpret_cast<const uint4 *>(bits)[g];
        s += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    u32 tot;
    u32 run = cta_exclusive_scan<THREADS>(s, sWarp, &tot);
    if (WRITE_PREFIX) {
#pragma unroll 1
        for (u32 g = g0; g < g1; ++g) {
            const uint4 v = reinterpret_cast<const uint4 *>(bits)[g];
            const u32 p1 = run + __popc(v.x), p2 = p1 + __popc(v.y), p3 = p2 + __popc(v.z);
            reinterpret_cast<uint2 *>(pre)[g] = make_uint2(run | (p1 << 16), p2 | (p3 << 16));
            run = p3 + __popc(v.w);
        }
    }
    return tot;
}

template <int THREADS, int E, typename T, bool NUMERIC>
__global__ void __launch_bounds__(THREADS)
k_rank_rows(const u32 *__restrict__ perm, const u32 *__restrict__ aRp, const u32 *__restrict__ aCi,
            const T *__restrict__ aV, const u32 *__restrict__ bRp, const u32 *__restrict__ bCi,
            const T *__restrict__ bV, const u32 *__restrict__ rowOps, const u32 *__restrict__ rowMin,
            const u32 *__restrict__ rowMax, u32 *cRp, u32 *__restrict__ cCi, T *__restrict__ cV)
{
    using L = RankLayout<THREADS, E, T, NUMERIC>;
    constexpr u32 NONE = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    T *outVal = reinterpret_cast<T *>(smemRaw + L::OUTVAL);
    T *sAv = reinterpret_cast<T *>(smemRaw + L::SAV);
    u32 *top = reinterpret_cast<u32 *>(smemRaw + L::TOP);
    unsigned short *topPre = reinterpret_cast<unsigned short *>(smemRaw + L::TOPPRE);
    u32 *leaf = reinterpret_cast<u32 *>(smemRaw + L::LEAF);
    unsigned short *leafPre = reinterpret_cast<unsigned short *>(smemRaw + L::LEAFPRE);
    unsigned short *leafId = reinterpret_cast<unsigned short *>(smemRaw + L::LEAFID);
    u32 *sIncl = reinterpret_cast<u32 *>(smemRaw + L::SINCL);
    u32 *sBs = reinterpret_cast<u32 *>(smemRaw + L::SBS);
    unsigned short *sTab = reinterpret_cast<unsigned short *>(smemRaw + L::STAB);
    u32 *outCol = reinterpret_cast<u32 *>(smemRaw + L::OUTCOL);
    __shared__ u32 sWarp[33];

    const u32 tid = threadIdx.x;
    const u32 row = perm[blockIdx.x];
    const u32 n = rowOps[row];                               // products of the row, <= CAP
    const u32 aBeg = aRp[row], aEnd = aRp[row + 1];
    const u32 cmin = rowMin[row];
    const u32 topWords = ((rowMax[row] - cmin) >> 10) + 1;   // <= RANK_TOP_WORDS: the host checks cols(B)
    u32 cBase = 0, nnzRow = 0;
    if (NUMERIC) {
        cBase = cRp[row];
        nnzRow = cRp[row + 1] - cBase;
    }
#pragma unroll 1
    for (u32 j = tid; j < (topWords + 3) >> 2; j += THREADS) reinterpret_cast<uint4 *>(top)[j] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    // ---------------------------------------------------------------- gather (flat over the CTA)
    // product gp (row-wide index: k ascending, then B-row order) lives in thread gp % THREADS, slot gp / THREADS;
    // slots i with i * THREADS >= n are empty in every thread: all slot loops stop there (CTA-uniform)
    u32 col[E];
    T prod[E];
#pragma unroll
    for (int i = 0; i < E; ++i) { col[i] = NONE; prod[i] = (T)0; }
    u32 base = 0;
#pragma unroll 1
    for (u32 ab = aBeg; ab < aEnd; ab += THREADS) {
        const u32 nb = min((u32)THREADS, aEnd - ab);
        u32 bs = 0, len = 0;
        if (tid < nb) {
            const u32 k = __ldg(aCi + ab + tid);
            bs = __ldg(bRp + k);
            len = __ldg(bRp + k + 1) - bs;
            if (NUMERIC) sAv[tid] = __ldg(aV + ab + tid);
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
        constexpr int GU = E < 4 ? E : 4;   // products per thread in flight
#pragma unroll
        for (int i0 = 0; i0 < E; i0 += GU) {
            if ((u32)i0 * THREADS >= base + total) break;           // CTA-uniform
            if ((u32)(i0 + GU) * THREADS <= base) continue;         // CTA-uniform (earlier batch)
            u32 q[GU], cc[GU];
            T av[GU], bv[GU];
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                const u32 gp = (u32)(i0 + u) * THREADS + tid;
                const u32 p = gp - base;
                q[u] = NONE;
                av[u] = (T)0;
                if (p < total) {   // gp < base wraps p beyond any total
                    u32 lo = sTab[p >> 5];
                    while (sIncl[lo] <= p) ++lo;
                    q[u] = sBs[lo] + p;
                    if (NUMERIC) av[u] = sAv[lo];
                }
            }
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                if ((u32)(i0 + u) * THREADS >= base + total) break;
                cc[u] = q[u] != NONE ? __ldg(bCi + q[u]) : 0u;
                bv[u] = (NUMERIC && q[u] != NONE) ? __ldg(bV + q[u]) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                if ((u32)(i0 + u) * THREADS >= base + total) break;
                if (q[u] != NONE) {
                    const u32 c = cc[u] - cmin;
                    col[i0 + u] = c;
                    if (NUMERIC) prod[i0 + u] = av[u] * bv[u];
                    atomicOr(&top[c >> 10], 1u << ((c >> 5) & 31));
                }
            }
        }
        base += total;
        __syncthreads();
    }

    // ---------------------------------------------------------------- leaf words of the touched chunks
    const u32 leaves = rank_scan_level<THREADS, true>(top, topPre, topWords, sWarp);
#pragma unroll 1
    for (u32 j = tid; j < (leaves + 3) >> 2; j += THREADS) reinterpret_cast<uint4 *>(leaf)[j] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    u32 dup = 0;   // bit i: slot i is not the first product of its column
#pragma unroll
    for (int i = 0; i < E; ++i) {
        if ((u32)i * THREADS >= n) break;
        if (col[i] != NONE) {
            const u32 c = col[i];
            const u32 tw = c >> 10, tb = (c >> 5) & 31;
            const u32 s = topPre[tw] + __popc(top[tw] & ((1u << tb) - 1u));
            const u32 bit = 1u << (c & 31);
            const u32 old = atomicOr(&leaf[s], bit);
            if (NUMERIC) {
                if (old & bit) dup |= 1u << i;
                leafId[s] = (unsigned short)(c >> 5);   // every product of the word stores the same id
            }
            col[i] = (s << 5) | (c & 31u);              // from here on: leaf slot and bit
        }
    }
    __syncthreads();

    if (!NUMERIC) {
        const u32 distinct = rank_scan_level<THREADS, false>(leaf, nullptr, leaves, sWarp);
        if (tid == 0) cRp[row] = distinct;
        return;
    } else {
        rank_scan_level<THREADS, true>(leaf, leafPre, leaves, sWarp);
        __syncthreads();
        // ------------------------------------------------------------ first product of a column: plain stores
#pragma unroll
        for (int i = 0; i < E; ++i) {
            if ((u32)i * THREADS >= n) break;
            if (col[i] != NONE) {
                const u32 s = col[i] >> 5, b = col[i] & 31;
                const u32 rank = leafPre[s] + __popc(leaf[s] & ((1u << b) - 1u));
                if (!((dup >> i) & 1u)) {
                    outVal[rank] = prod[i];
                    outCol[rank] = cmin + ((u32)leafId[s] << 5) + b;
                }
                col[i] = rank;
            }
        }
        // ------------------------------------------------------------ further products of a column: add
        if (__syncthreads_or(dup != 0)) {
#pragma unroll
            for (int i = 0; i < E; ++i)
                if ((dup >> i) & 1u) atomicAdd(&outVal[col[i]], prod[i]);
            __syncthreads();
        }
#pragma unroll 1
        for (u32 j = tid; j < nnzRow; j += THREADS) {
            cCi[cBase + j] = outCol[j];
            cV[cBase + j] = outVal[j];
        }
    }
}

template <int THREADS, int E, typename T, bool NUMERIC>
void launch_rank_rows(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi,
                      const T *aV, const u32 *bRp, const u32 *bCi, const T *bV, const u32 *rowOps, const u32 *rowMin,
                      const u32 *rowMax, u32 *cRp, u32 *cCi, T *cV)
{
    using L = RankLayout<THREADS, E, T, NUMERIC>;
    auto kern = k_rank_rows<THREADS, E, T, NUMERIC>;
    if (L::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    kern<<<count, THREADS, L::SMEM, lc.stream>>>(perm, aRp, aCi, aV, bRp, bCi, bV, rowOps, rowMin, rowMax, cRp, cCi, cV);
    ++*lc.launches;
}

}  // namespace sb
