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
constexpr u32 RANK_MAX_EXTENT = 1u << 20;    // two levels: RANK_TOP_WORDS * 32 * 32 columns
constexpr u32 RANK_MAX_EXTENT3 = 1u << 25;   // three levels

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

// Blocked product assignment of the CTA kernels: thread t owns the row's products gp in [t * EP, (t + 1) * EP),
// EP = ceil(n / THREADS) <= E (gp = index in the row's flat enumeration: A entries ascending, then B-row order).
// A thread locates the A entry of its first product through the owner table and then walks forward: consecutive
// products mostly stay inside one B row, so the per-product cost is one compare.  (The strided assignment used
// before needed a table lookup + walk per product: ~25 of the ~50 gather instructions per product.)
template <typename T, bool WITH_VALUE>
struct SegWalk {
    u32 lo, segEnd, segBs;
    T av;
    __device__ __forceinline__ void start(u32 p, const unsigned short *sTab, const u32 *sIncl, const u32 *sBs, const T *sAv)
    {
        lo = sTab[p >> 5];
        while ((segEnd = sIncl[lo]) <= p) ++lo;
        segBs = sBs[lo];
        if (WITH_VALUE) av = sAv[lo];
    }
    // position in B (col_ids / data index) of product p of the batch, p not below the previous call's
    __device__ __forceinline__ u32 locate(u32 p, const u32 *sIncl, const u32 *sBs, const T *sAv)
    {
        if (p >= segEnd) {
            do { ++lo; segEnd = sIncl[lo]; } while (p >= segEnd);
            segBs = sBs[lo];
            if (WITH_VALUE) av = sAv[lo];
        }
        return segBs + p;
    }
};

// Shared memory of a CTA of THREADS threads (CAP = THREADS * E product slots), 16-byte aligned pieces, all
// offsets compile-time:
//   outVal[CAP] T | sAv[THREADS] T | top[1024] | topPre[1024] u16 | leaf[CAP] | leafPre[CAP] u16 |
//   leafId[CAP] u16 | sIncl[THREADS] | sBs[THREADS] | sTab[CAP/32] u16 | outCol[CAP] (numeric; aliased onto
//   top/topPre when it fits: both are dead once the leaf words are filled)
template <int THREADS, int E, typename T, bool NUMERIC, int LEVELS = 2>
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
    static constexpr size_t MID = LEAFID + (NUMERIC ? al(CAP * 2) : 0);        // LEVELS == 3: mid words + prefixes
    static constexpr size_t MIDPRE = MID + (LEVELS == 3 ? al(CAP * 4) : 0);
    static constexpr size_t SINCL = MIDPRE + (LEVELS == 3 ? al(CAP * 2) : 0);
    static constexpr size_t SBS = SINCL + al(THREADS * 4);
    static constexpr size_t STAB = SBS + al(THREADS * 4);
    static constexpr size_t OUTCOL = (NUMERIC && COL_ALIASED) ? TOP : STAB + al(CAP / 32 * 2);
    static constexpr size_t SMEM = STAB + al(CAP / 32 * 2) + ((NUMERIC && !COL_ALIASED) ? CAP * 4 : 0);
};

// popcount prefix of a level: words[0..n) (padded with zero words to a multiple of 4 by the caller)
// -> pre[j] = number of set bits in words[0..j), returns the total.  Threads own consecutive groups of four
// words (one 16-byte load, one 8-byte store of four packed u16 prefixes).
// Thread t owns the MAXG consecutive groups t * MAXG ..; MAXG is a compile-time bound (1024 top words / 4 / THREADS
// or E / 4), so both loops are fully unrolled.
template <int THREADS, int MAXG, bool WRITE_PREFIX>
__device__ __forceinline__ u32 rank_scan_level(const u32 *bits, unsigned short *pre, u32 words, u32 *sWarp)
{
    const u32 groups = (words + 3) >> 2;
    const u32 g0 = threadIdx.x * MAXG;
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
        if (g0 + k < groups) {
            const uint4 w = reinterpret_cast<const uint4 *>(bits)[g0 + k];
            s += __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
        }
    }
    u32 tot;
    u32 run = cta_exclusive_scan<THREADS>(s, sWarp, &tot);
    if (WRITE_PREFIX) {   // the words are read again: keeping 4 * MAXG popcounts across the scan spills at 32 registers
#pragma unroll
        for (int k = 0; k < MAXG; ++k) {
            if (g0 + k < groups) {
                const uint4 w = reinterpret_cast<const uint4 *>(bits)[g0 + k];
                const u32 p1 = run + __popc(w.x), p2 = p1 + __popc(w.y), p3 = p2 + __popc(w.z);
                reinterpret_cast<uint2 *>(pre)[g0 + k] = make_uint2(run | (p1 << 16), p2 | (p3 << 16));
                run = p3 + __popc(w.w);
            }
        }
    }
    return tot;
}

// clears the first `words` words (rounded up to groups of four) of a level; same compile-time bound
template <int THREADS, int MAXG>
__device__ __forceinline__ void rank_clear_level(u32 *bits, u32 words)
{
    const u32 groups = (words + 3) >> 2;
#pragma unroll
    for (int k = 0; k < MAXG; ++k) {
        const u32 g = threadIdx.x + k * THREADS;
        if (g < groups) reinterpret_cast<uint4 *>(bits)[g] = make_uint4(0, 0, 0, 0);
    }
}

// MODE: RANK_COUNT   symbolic, distinct columns per row only
//       RANK_NUMERIC self-contained numeric phase (no rank map available)
//       RANK_MAP     symbolic + rank map: every product's sorted position (and whether it is the first
//                    product of its column) is recorded, so the numeric phase needs no bitmaps at all
//                    (k_map_rows_cta below)
enum { RANK_COUNT = 0, RANK_NUMERIC = 1, RANK_MAP = 2 };

// the mapped symbolic kernel is instruction-bound: full occupancy (32 registers) measured best
template <int THREADS, int E, typename T, int MODE, int LEVELS>
__global__ void __launch_bounds__(THREADS, (MODE == RANK_MAP && E <= 8) ? (2048 / THREADS > 32 ? 32 : 2048 / THREADS) : 1)
k_rank_rows(const u32 *__restrict__ perm, const u32 *__restrict__ aRp, const u32 *__restrict__ aCi,
            const T *__restrict__ aV, const u32 *__restrict__ bRp, const u32 *__restrict__ bCi,
            const T *__restrict__ bV, const u32 *__restrict__ rowOps, const u32 *__restrict__ rowMin,
            const u32 *__restrict__ rowMax, const RowDesc *__restrict__ desc, const uint2 *__restrict__ aSeg,
            unsigned short *__restrict__ rankMap, u32 *cRp, u32 *__restrict__ cCi, T *__restrict__ cV)
{
    constexpr bool NUMERIC = MODE == RANK_NUMERIC;
    constexpr bool DESC = MODE == RANK_MAP;   // row parameters from the descriptor, B-row bounds from aSeg
    static_assert(LEVELS == 2 || (LEVELS == 3 && !NUMERIC), "three levels: symbolic modes only");
    constexpr int TOPSHIFT = 5 * LEVELS;      // column bits below a top word
    constexpr int TOPG = (RANK_TOP_WORDS / 4 + THREADS - 1) / THREADS;   // groups of four words per thread: top level
    constexpr int LEAFG = (E + 3) / 4;                                   // ... compact levels (<= CAP words)
    using L = RankLayout<THREADS, E, T, NUMERIC, LEVELS>;
    constexpr u32 NONE = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    T *outVal = reinterpret_cast<T *>(smemRaw + L::OUTVAL);
    T *sAv = reinterpret_cast<T *>(smemRaw + L::SAV);
    u32 *top = reinterpret_cast<u32 *>(smemRaw + L::TOP);
    unsigned short *topPre = reinterpret_cast<unsigned short *>(smemRaw + L::TOPPRE);
    u32 *leaf = reinterpret_cast<u32 *>(smemRaw + L::LEAF);
    unsigned short *leafPre = reinterpret_cast<unsigned short *>(smemRaw + L::LEAFPRE);
    unsigned short *leafId = reinterpret_cast<unsigned short *>(smemRaw + L::LEAFID);
    u32 *mid = reinterpret_cast<u32 *>(smemRaw + L::MID);
    unsigned short *midPre = reinterpret_cast<unsigned short *>(smemRaw + L::MIDPRE);
    u32 *sIncl = reinterpret_cast<u32 *>(smemRaw + L::SINCL);
    u32 *sBs = reinterpret_cast<u32 *>(smemRaw + L::SBS);
    unsigned short *sTab = reinterpret_cast<unsigned short *>(smemRaw + L::STAB);
    u32 *outCol = reinterpret_cast<u32 *>(smemRaw + L::OUTCOL);
    __shared__ u32 sWarp[33];

    const u32 tid = threadIdx.x;
    u32 row, n, aBeg, aEnd, cmin, cmax;
    u64 mapOff = 0;
    if (DESC) {
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(desc + blockIdx.x));
        const uint4 d1 = __ldg(reinterpret_cast<const uint4 *>(desc + blockIdx.x) + 1);
        aBeg = d0.x; aEnd = d0.x + d0.y; n = d0.z; row = d0.w;
        cmin = d1.x; cmax = d1.y;
        mapOff = ((u64)d1.w << 32) | d1.z;
    } else {
        row = perm[blockIdx.x];
        n = rowOps[row];                                     // products of the row, <= CAP
        aBeg = aRp[row]; aEnd = aRp[row + 1];
        cmin = rowMin[row]; cmax = rowMax[row];
    }
    const u32 topWords = ((cmax - cmin) >> TOPSHIFT) + 1;    // <= RANK_TOP_WORDS: the host checks cols(B)
    u32 cBase = 0, nnzRow = 0;
    if (NUMERIC) {
        cBase = cRp[row];
        nnzRow = cRp[row + 1] - cBase;
    }
    rank_clear_level<THREADS, TOPG>(top, topWords);
    __syncthreads();

    // ---------------------------------------------------------------- gather (flat over the CTA)
    // product gp (row-wide index: k ascending, then B-row order) lives in thread gp / EP, slot gp % EP (SegWalk);
    // slots i >= EP are empty in every thread: all slot loops stop there (CTA-uniform)
    const u32 EP = (n + THREADS - 1) / THREADS;
    const u32 myFirst = tid * EP;
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
            if (DESC) {
                const uint2 seg = __ldg(aSeg + ab + tid);
                bs = seg.x;
                len = seg.y - seg.x;
            } else {
                const u32 k = __ldg(aCi + ab + tid);
                bs = __ldg(bRp + k);
                len = __ldg(bRp + k + 1) - bs;
            }
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
        const u32 first = max(myFirst, base), last = min(min(myFirst + EP, n), base + total);
        const u32 mine = first < last ? last - first : 0u;   // this thread's products in the batch
        SegWalk<T, NUMERIC> walk;
        if (mine) walk.start(first - base, sTab, sIncl, sBs, sAv);
#pragma unroll
        for (int i0 = 0; i0 < E; i0 += GU) {
            if ((u32)i0 >= EP) break;           // CTA-uniform
            u32 q[GU], cc[GU];
            T av[GU], bv[GU];
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                const u32 gp = myFirst + (u32)(i0 + u);
                q[u] = NONE;
                av[u] = (T)0;
                if (gp - first < mine) {   // unsigned: also false for gp < first
                    q[u] = walk.locate(gp - base, sIncl, sBs, sAv);
                    if (NUMERIC) av[u] = walk.av;
                }
            }
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                cc[u] = q[u] != NONE ? __ldg(bCi + q[u]) : 0u;
                bv[u] = (NUMERIC && q[u] != NONE) ? __ldg(bV + q[u]) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                if (q[u] != NONE) {
                    const u32 c = cc[u] - cmin;
                    col[i0 + u] = c;
                    if (NUMERIC) prod[i0 + u] = av[u] * bv[u];
                    atomicOr(&top[c >> TOPSHIFT], 1u << ((c >> (TOPSHIFT - 5)) & 31));
                }
            }
        }
        base += total;
        __syncthreads();
    }

    // ---------------------------------------------------------------- (three levels) mid words of the touched
    // 1024-column blocks: same step as the leaf level below, one digit higher
    u32 upperWords = topWords;   // words of the level above the leaves
    if (LEVELS == 3) {
        const u32 mids = rank_scan_level<THREADS, TOPG, true>(top, topPre, topWords, sWarp);
        rank_clear_level<THREADS, LEAFG>(mid, mids);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < E; ++i) {
            if ((u32)i >= EP) break;
            if (col[i] != NONE) {
                const u32 c = col[i];
                const u32 tw = c >> 15, tb = (c >> 10) & 31;
                const u32 ms = topPre[tw] + __popc(top[tw] & ((1u << tb) - 1u));
                atomicOr(&mid[ms], 1u << ((c >> 5) & 31));
                col[i] = (ms << 10) | (c & 1023u);   // mid slot (< 2^14) and the two low digits
            }
        }
        __syncthreads();
        upperWords = mids;
    }
    const u32 *upper = LEVELS == 3 ? mid : top;
    unsigned short *upperPre = LEVELS == 3 ? midPre : topPre;

    // ---------------------------------------------------------------- leaf words of the touched chunks
    const u32 leaves = LEVELS == 3 ? rank_scan_level<THREADS, LEAFG, true>(upper, upperPre, upperWords, sWarp)
                                   : rank_scan_level<THREADS, TOPG, true>(upper, upperPre, upperWords, sWarp);
    rank_clear_level<THREADS, LEAFG>(leaf, leaves);
    __syncthreads();
    u32 dup = 0;   // bit i: slot i is not the first product of its column
#pragma unroll
    for (int i = 0; i < E; ++i) {
        if ((u32)i >= EP) break;
        if (col[i] != NONE) {
            const u32 c = col[i];
            const u32 tw = c >> 10, tb = (c >> 5) & 31;   // word / bit in the level above (three levels: mid slot)
            const u32 s = upperPre[tw] + __popc(upper[tw] & ((1u << tb) - 1u));
            const u32 bit = 1u << (c & 31);
            const u32 old = atomicOr(&leaf[s], bit);
            if (MODE != RANK_COUNT && (old & bit)) dup |= 1u << i;
            if (NUMERIC) leafId[s] = (unsigned short)(c >> 5);   // every product of the word stores the same id
            col[i] = (s << 5) | (c & 31u);              // from here on: leaf slot and bit
        }
    }
    __syncthreads();

    if (MODE == RANK_COUNT) {
        const u32 distinct = rank_scan_level<THREADS, LEAFG, false>(leaf, nullptr, leaves, sWarp);
        if (tid == 0) cRp[row] = distinct;
        return;
    } else if (MODE == RANK_MAP) {
        const u32 distinct = rank_scan_level<THREADS, LEAFG, true>(leaf, leafPre, leaves, sWarp);
        if (tid == 0) cRp[row] = distinct;
        __syncthreads();
        unsigned short *map = rankMap + mapOff;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            if ((u32)i >= EP) break;
            if (col[i] != NONE) {
                const u32 s = col[i] >> 5, b = col[i] & 31;
                const u32 rank = leafPre[s] + __popc(leaf[s] & ((1u << b) - 1u));
                map[myFirst + (u32)i] = (unsigned short)(rank | (((dup >> i) & 1u) ? MAP_DUP : 0u));
            }
        }
        return;
    } else {
        rank_scan_level<THREADS, LEAFG, true>(leaf, leafPre, leaves, sWarp);
        __syncthreads();
        // ------------------------------------------------------------ first product of a column: plain stores
#pragma unroll
        for (int i = 0; i < E; ++i) {
            if ((u32)i >= EP) break;
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

template <int THREADS, int E, typename T, int MODE, int LEVELS = 2>
void launch_rank_rows(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi,
                      const T *aV, const u32 *bRp, const u32 *bCi, const T *bV, const u32 *rowOps, const u32 *rowMin,
                      const u32 *rowMax, const RowDesc *desc, const uint2 *aSeg, unsigned short *rankMap, u32 *cRp,
                      u32 *cCi, T *cV)
{
    using L = RankLayout<THREADS, E, T, MODE == RANK_NUMERIC, LEVELS>;
    auto kern = k_rank_rows<THREADS, E, T, MODE, LEVELS>;
    if (L::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    kern<<<count, THREADS, L::SMEM, lc.stream>>>(perm, aRp, aCi, aV, bRp, bCi, bV, rowOps, rowMin, rowMax, desc, aSeg,
                                                 rankMap, cRp, cCi, cV);
    ++*lc.launches;
}

// ------------------------------------------------------------------------------------------------
// Numeric phase of a mapped row (CTA per row): gather, multiply, scatter by the recorded rank into a
// shared staging row, write the row to C coalesced.  No bitmaps, no scans beyond the B-row lengths.
// ------------------------------------------------------------------------------------------------
// COLDIRECT: the column ids of the first products go straight to C (scattered 4-byte stores inside the row's range,
// merged by the L2 before they reach DRAM) and only the values are staged: two thirds of the shared memory, which
// lets the 1024-thread shapes keep two CTAs per SM instead of one.
template <int THREADS, int E, typename T, bool COLDIRECT = false>
struct MapCtaLayout {
    static constexpr size_t al(size_t b) { return (b + 15) / 16 * 16; }
    static constexpr size_t CAP = (size_t)THREADS * E;
    static constexpr size_t OUTVAL = 0;
    static constexpr size_t SAV = OUTVAL + al(CAP * sizeof(T));
    static constexpr size_t OUTCOL = SAV + al(THREADS * sizeof(T));
    static constexpr size_t SINCL = OUTCOL + (COLDIRECT ? 0 : al(CAP * 4));
    static constexpr size_t SBS = SINCL + al(THREADS * 4);
    static constexpr size_t STAB = SBS + al(THREADS * 4);
    static constexpr size_t SMEM = STAB + al(CAP / 32 * 2);
};

#ifndef SB_MAP_GU
#define SB_MAP_GU 4
#endif
#ifndef SB_MAP_CTA_THREADS_PER_SM
#define SB_MAP_CTA_THREADS_PER_SM 1536
#endif
template <int THREADS, int E, typename T, bool COLDIRECT>
__global__ void __launch_bounds__(THREADS, (COLDIRECT && THREADS == 1024 && E == 8) ? 2 : ((SB_MAP_CTA_THREADS_PER_SM / THREADS) > 0 ? (SB_MAP_CTA_THREADS_PER_SM / THREADS) : 1))
k_map_rows_cta(const RowDesc *__restrict__ desc, const uint2 *__restrict__ aSeg, const T *__restrict__ aV,
               const u32 *__restrict__ bCi, const T *__restrict__ bV, const unsigned short *__restrict__ rankMap,
               u32 *__restrict__ cCi, T *__restrict__ cV)
{
    using L = MapCtaLayout<THREADS, E, T, COLDIRECT>;
    constexpr u32 NONE = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    T *outVal = reinterpret_cast<T *>(smemRaw + L::OUTVAL);
    T *sAv = reinterpret_cast<T *>(smemRaw + L::SAV);
    u32 *outCol = reinterpret_cast<u32 *>(smemRaw + L::OUTCOL);
    u32 *sIncl = reinterpret_cast<u32 *>(smemRaw + L::SINCL);
    u32 *sBs = reinterpret_cast<u32 *>(smemRaw + L::SBS);
    unsigned short *sTab = reinterpret_cast<unsigned short *>(smemRaw + L::STAB);
    __shared__ u32 sWarp[33];

    const u32 tid = threadIdx.x;
    const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(desc + blockIdx.x));
    const uint4 d1 = __ldg(reinterpret_cast<const uint4 *>(desc + blockIdx.x) + 1);
    const u32 aBeg = d0.x, aEnd = d0.x + d0.y, n = d0.z;
    const u32 cBase = d1.x, nnzRow = d1.y;
    const unsigned short *map = rankMap + (((u64)d1.w << 32) | d1.z);
    const u32 EP = (n + THREADS - 1) / THREADS;   // products per thread, blocked assignment (SegWalk)
    const u32 myFirst = tid * EP;

    // Products that are not the first of their column are added once every first product is in place: their
    // slots are remembered in a bit mask and re-located from the (still valid) tables after a barrier.  Rows
    // whose A entries need several batches (rare) add every product into a zeroed staging row instead.
    const bool multi = (aEnd - aBeg) > (u32)THREADS;
    if (multi)
        for (u32 j = tid; j < nnzRow; j += THREADS) outVal[j] = (T)0;
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
        constexpr int GU = E < SB_MAP_GU ? E : SB_MAP_GU;   // products per thread in flight
        const u32 first = max(myFirst, base), last = min(min(myFirst + EP, n), base + total);
        const u32 mine = first < last ? last - first : 0u;   // this thread's products in the batch
        SegWalk<T, true> walk;
        if (mine) walk.start(first - base, sTab, sIncl, sBs, sAv);
#pragma unroll
        for (int i0 = 0; i0 < E; i0 += GU) {
            if ((u32)i0 >= EP) break;           // CTA-uniform
            u32 q[GU], cc[GU], code[GU];
            T av[GU], bv[GU];
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                const u32 gp = myFirst + (u32)(i0 + u);
                q[u] = NONE;
                av[u] = (T)0;
                code[u] = 0;
                if (gp - first < mine) {   // unsigned: also false for gp < first
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
                    const u32 r = code[u] & MAP_RANK_MASK;
                    const T pr = av[u] * bv[u];
                    if (multi) {
                        if (!(code[u] & MAP_DUP)) {
                            if (COLDIRECT) cCi[cBase + r] = cc[u]; else outCol[r] = cc[u];
                        }
                        atomicAdd(&outVal[r], pr);
                    } else if (code[u] & MAP_DUP) {
                        dup |= 1u << (i0 + u);
                    } else {
                        outVal[r] = pr;
                        if (COLDIRECT) cCi[cBase + r] = cc[u]; else outCol[r] = cc[u];
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
            atomicAdd(&outVal[map[gp] & MAP_RANK_MASK], walk.av * __ldg(bV + walk.segBs + gp));
        }
        __syncthreads();
    }
#pragma unroll 1
    for (u32 j = tid; j < nnzRow; j += THREADS) {
        if (!COLDIRECT) cCi[cBase + j] = outCol[j];
        cV[cBase + j] = outVal[j];
    }
}

template <int THREADS, int E, typename T, bool COLDIRECT = false>
void launch_map_rows_cta(const LaunchCtx &lc, const RowDesc *desc, u32 count, const uint2 *aSeg, const T *aV,
                         const u32 *bCi, const T *bV, const unsigned short *rankMap, u32 *cCi, T *cV)
{
    using L = MapCtaLayout<THREADS, E, T, COLDIRECT>;
    auto kern = k_map_rows_cta<THREADS, E, T, COLDIRECT>;
    if (L::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    kern<<<count, THREADS, L::SMEM, lc.stream>>>(desc, aSeg, aV, bCi, bV, rankMap, cCi, cV);
    ++*lc.launches;
}

}  // namespace sb
