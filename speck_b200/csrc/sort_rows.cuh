// speck_b200/csrc/sort_rows.cuh -- "sort" row classes: one lane group (G lanes, E products per
// lane, N = G*E) per row of C.
//
//   gather : the group enumerates the row's products p = 0..ops-1 in (k ascending, B-row order);
//            consecutive lanes read consecutive entries of a B row (coalesced) and the owner A
//            entry of each product is found by a shuffle binary search over the inclusive scan
//            of the B-row lengths (no per-product linear search as in readOperations /
//            iterateMatrixNumeric of the reference, spECK_HashSpGEMM.cuh:130-215).
//   sort   : keys (col << IDXBITS | p) are sorted by a bitonic network held in registers
//            (blocked layout: lane l owns logical elements l*E .. l*E+E-1), shuffles only for
//            partner distances >= E.  No shared-memory atomics, no hash table, no O(n^2) rank
//            sort (reference :715-728, :837-850).
//   emit   : symbolic -> number of distinct columns; numeric -> heads of equal-column runs are
//            ranked with ballot/popc, the run's products are added in ascending p (= ascending k,
//            the order of a sequential Gustavson loop) and (col, val) is written to consecutive positions of C.
#pragma once
#include "common.cuh"

namespace sb {

constexpr int SORT_BLOCK = 128;

template <int N> struct Log2 { static constexpr int value = 1 + Log2<N / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

template <typename K> __device__ __forceinline__ K key_min(K a, K b) { return a < b ? a : b; }
template <typename K> __device__ __forceinline__ K key_max(K a, K b) { return a < b ? b : a; }

// half-cleaner stages j = jStart, jStart/2, ..., 1 (partner = idx ^ j, ascending)
template <int G, int E, typename KeyT>
__device__ __forceinline__ void bitonic_half_cleaners(KeyT (&reg)[E], const u32 l, const u32 gmask, const int jStart)
{
#pragma unroll
    for (int j = (G * E) >> 1; j > 0; j >>= 1) {
        if (j > jStart) continue;  // resolved at compile time once the caller's loop is unrolled
        if (j >= E) {
            const int lm = j / E;
            const bool lower = (l & lm) == 0;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const KeyT o = __shfl_xor_sync(gmask, reg[r], lm, G);
                reg[r] = lower ? key_min(reg[r], o) : key_max(reg[r], o);
            }
        } else {
#pragma unroll
            for (int r = 0; r < E; ++r) {
                if ((r & j) == 0) {
                    const KeyT a = reg[r], b = reg[r | j];
                    reg[r] = key_min(a, b);
                    reg[r | j] = key_max(a, b);
                }
            }
        }
    }
}

// Ascending bitonic sort of N = G*E keys, blocked layout (logical index = l*E + r).
// "Mirrored" formulation: each merge of width k starts with partner = idx ^ (k-1), followed by
// half-cleaners partner = idx ^ j (j = k/4 .. 1).  Every compare-exchange is ascending (lower index
// keeps the minimum), so no direction selects are needed: 2 min/max per intra-lane exchange, one
// shuffle + one predicated min/max per inter-lane exchange.
template <int G, int E, typename KeyT>
__device__ __forceinline__ void bitonic_sort_regs(KeyT (&reg)[E], const u32 l, const u32 gmask)
{
    constexpr int N = G * E;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
        // ---- mirrored stage
        if (k <= E) {
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int r2 = r ^ (k - 1);
                if (r < r2) {
                    const KeyT a = reg[r], b = reg[r2];
                    reg[r] = key_min(a, b);
                    reg[r2] = key_max(a, b);
                }
            }
        } else {
            const int lm = k / E - 1;                 // partner lane = l ^ lm, partner register = E-1-r
            const bool lower = (l & (k / (2 * E))) == 0;
#pragma unroll
            for (int r = 0; r < (E + 1) / 2; ++r) {
                const int r2 = E - 1 - r;
                const KeyT o1 = __shfl_xor_sync(gmask, reg[r2], lm, G);
                if (r2 != r) {
                    const KeyT o2 = __shfl_xor_sync(gmask, reg[r], lm, G);
                    reg[r2] = lower ? key_min(reg[r2], o2) : key_max(reg[r2], o2);
                }
                reg[r] = lower ? key_min(reg[r], o1) : key_max(reg[r], o1);
            }
        }
        // ---- half cleaners
        bitonic_half_cleaners<G, E, KeyT>(reg, l, gmask, k >> 2);
    }
}

// MODE: SORT_COUNT   symbolic, distinct columns only (keys = columns)
//       SORT_NUMERIC self-contained numeric phase (keys = col << IDXBITS | product index, values staged)
//       SORT_MAP     symbolic + rank map: sorted position of every product and its first-of-column flag
//                    are written to the rank map (common.cuh); the numeric phase is then k_map_rows
enum { SORT_COUNT = 0, SORT_NUMERIC = 1, SORT_MAP = 2 };

template <int G, int E, typename KeyT, typename T, int MODE>
struct SortLayout {
    static constexpr int N = G * E;
    static constexpr int NPAD = N + N / 32;            // one pad word per 32: conflict-free blocked write-back
    static constexpr int GROUPS = SORT_BLOCK / G;      // groups (rows) per block
    static constexpr size_t KEY_BYTES = (size_t)GROUPS * NPAD * sizeof(KeyT);
    // second array per group: staged values (numeric) or the row's rank codes in product order (map)
    static constexpr size_t VAL_BYTES = MODE == SORT_NUMERIC ? (size_t)GROUPS * N * sizeof(T)
                                        : (MODE == SORT_MAP ? (size_t)GROUPS * N * sizeof(unsigned short) : 0);
    static constexpr size_t SMEM = ((KEY_BYTES + 15) / 16) * 16 + VAL_BYTES;
};

template <int G, int E, typename KeyT, typename T, int MODE>
__device__ __forceinline__ void
sort_rows_body(const u32 blk, const u32 *__restrict__ perm, const u32 count, const u32 *__restrict__ aRp,
            const u32 *__restrict__ aCi, const T *__restrict__ aV, const u32 *__restrict__ bRp,
            const u32 *__restrict__ bCi, const T *__restrict__ bV, const u32 *__restrict__ rowOps,
            u32 *cRp /* symbolic: counts out; numeric: offsets in */, u32 *__restrict__ cCi,
            T *__restrict__ cV, const RowDesc *__restrict__ desc, const uint2 *__restrict__ aSeg,
            unsigned short *__restrict__ rankMap)
{
    constexpr bool NUMERIC = MODE == SORT_NUMERIC;
    constexpr bool DESC = MODE == SORT_MAP;        // row parameters from the descriptor, B-row bounds from aSeg
    constexpr bool IDXKEYS = MODE != SORT_COUNT;   // keys carry the product index
    using L = SortLayout<G, E, KeyT, T, MODE>;
    constexpr int N = L::N;
    constexpr int IDXBITS = Log2<N>::value;
    constexpr KeyT SENT = ~(KeyT)0;
    extern __shared__ __align__(16) unsigned char smemRaw[];

    const u32 laneW = threadIdx.x & 31;
    const u32 l = threadIdx.x % G;
    const u32 grp = threadIdx.x / G;
    const u32 gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (laneW - l));
    KeyT *keys = reinterpret_cast<KeyT *>(smemRaw) + (size_t)grp * L::NPAD;
    T *vals = reinterpret_cast<T *>(smemRaw + ((L::KEY_BYTES + 15) / 16) * 16) + (size_t)grp * N;
    unsigned short *codes = reinterpret_cast<unsigned short *>(smemRaw + ((L::KEY_BYTES + 15) / 16) * 16) + (size_t)grp * N;

    const u32 gidx = blk * L::GROUPS + grp;
    const bool active = gidx < count;
    u32 row = 0, ops = 0, aBeg = 0, aEnd = 0;
    u64 mapOff = 0;
    if (active) {
        if (DESC) {
            const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(desc + gidx));
            const uint4 d1 = __ldg(reinterpret_cast<const uint4 *>(desc + gidx) + 1);
            aBeg = d0.x; aEnd = d0.x + d0.y; ops = d0.z; row = d0.w;
            mapOff = ((u64)d1.w << 32) | d1.z;
        } else {
            row = perm[gidx];
            ops = rowOps[row];
            aBeg = aRp[row];
            aEnd = aRp[row + 1];
        }
    }

    // ---------------------------------------------------------------- gather
    u32 base = 0;
    for (u32 ab = aBeg; ab < aEnd; ab += G) {
        const u32 ai = ab + l;
        u32 bs = 0, len = 0;
        T av = (T)0;
        if (ai < aEnd) {
            if (DESC) {
                const uint2 seg = __ldg(aSeg + ai);
                bs = seg.x;
                len = seg.y - seg.x;
            } else {
                const u32 k = __ldg(aCi + ai);
                bs = __ldg(bRp + k);
                len = __ldg(bRp + k + 1) - bs;
            }
            if (NUMERIC) av = __ldg(aV + ai);
        }
        u32 incl = len;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const u32 t = __shfl_up_sync(gmask, incl, d, G);
            if ((int)l >= d) incl += t;
        }
        const u32 total = __shfl_sync(gmask, incl, G - 1, G);
        const u32 qb = bs - (incl - len);   // position in B of product p of this entry = qb + p
        // U products per lane and iteration (whole-warp groups only): owners first, then all loads, then the stores
        constexpr int U = E >= 4 ? 4 : 1;
        for (u32 p0 = 0; p0 < total; p0 += U * G) {
            u32 q[U], col[U];
            T oAv[U], bv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u > 0 && p0 + u * G >= total) break;   // group-uniform
                const u32 p = p0 + u * G + l;
                u32 lo = 0;
#pragma unroll
                for (int s = G / 2; s >= 1; s >>= 1) {
                    const u32 v = __shfl_sync(gmask, incl, lo + s - 1, G);
                    if (v <= p) lo += s;
                }
                q[u] = __shfl_sync(gmask, qb, lo, G) + p;
                if (NUMERIC) oAv[u] = __shfl_sync(gmask, av, lo, G);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u > 0 && p0 + u * G >= total) break;
                const bool ok = p0 + u * G + l < total;
                col[u] = ok ? __ldg(bCi + q[u]) : 0u;
                if (NUMERIC) bv[u] = ok ? __ldg(bV + q[u]) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u > 0 && p0 + u * G >= total) break;
                const u32 p = p0 + u * G + l;
                if (p < total) {
                    const u32 gp = base + p;
                    if (IDXKEYS) {
                        keys[gp] = ((KeyT)col[u] << IDXBITS) | (KeyT)gp;
                        if (NUMERIC) vals[gp] = oAv[u] * bv[u];
                    } else {
                        keys[gp] = (KeyT)col[u];
                    }
                }
            }
        }
        base += total;
    }
    __syncwarp(gmask);

    // ---------------------------------------------------------------- sort
    KeyT reg[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const u32 idx = r * G + l;  // conflict-free striped read; the logical index is l*E + r
        reg[r] = idx < ops ? keys[idx] : SENT;
    }
    __syncwarp(gmask);
    bitonic_sort_regs<G, E, KeyT>(reg, l, gmask);

    if (MODE == SORT_COUNT) {
        // ------------------------------------------------------------ symbolic: distinct columns
        KeyT prevLast = __shfl_up_sync(gmask, reg[E - 1], 1, G);
        u32 cnt = 0;
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const KeyT prev = (r == 0) ? prevLast : reg[r - 1];
            const bool valid = reg[r] != SENT;
            const bool first = (r == 0) && (l == 0);
            cnt += (valid && (first || reg[r] != prev)) ? 1u : 0u;
        }
#pragma unroll
        for (int d = G / 2; d >= 1; d >>= 1) cnt += __shfl_xor_sync(gmask, cnt, d, G);
        if (active && l == 0) cRp[row] = cnt;
        return;
    } else if (MODE == SORT_MAP) {
        // ------------------------------------------------------------ symbolic + rank map
        // lane l holds the sorted elements l*E .. l*E+E-1: heads of equal-column runs are found in registers, a
        // group scan of the per-lane head counts gives every element its position; the codes go to shared memory
        // in product order (they are stored to the map coalesced below)
        constexpr KeyT IDXMASK = ((KeyT)1 << IDXBITS) - 1;
        const KeyT prevLast = __shfl_up_sync(gmask, reg[E - 1], 1, G);
        u32 headBits = 0, cnt = 0;
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const KeyT prev = (r == 0) ? prevLast : reg[r - 1];
            const bool first = (r == 0) && (l == 0);
            // element l*E + r is a product iff it lies below ops (sentinels sort last; a key may equal the sentinel)
            const bool head = l * E + r < ops && (first || (u32)(reg[r] >> IDXBITS) != (u32)(prev >> IDXBITS));
            headBits |= (head ? 1u : 0u) << r;
            cnt += head ? 1u : 0u;
        }
        u32 incl = cnt;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const u32 t = __shfl_up_sync(gmask, incl, d, G);
            if ((int)l >= d) incl += t;
        }
        const u32 running = __shfl_sync(gmask, incl, G - 1, G);   // distinct columns of the row
        u32 pos = incl - cnt;                                      // heads in the lanes before this one
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const bool head = (headBits >> r) & 1u;
            pos += head ? 1u : 0u;
            if (l * E + r < ops)   // position = heads at or before this product - 1
                codes[(u32)(reg[r] & IDXMASK)] = (unsigned short)((pos - 1u) | (head ? 0u : MAP_DUP));
        }
        __syncwarp(gmask);
        if (active) {
            unsigned short *map = rankMap + mapOff;
            for (u32 j = l; j < ops; j += G) map[j] = codes[j];
            if (l == 0) cRp[row] = running;
        }
        return;
    } else {
        // ------------------------------------------------------------ numeric: fold runs, write C
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const u32 idx = l * E + r;
            keys[idx + (idx >> 5)] = reg[r];
        }
        __syncwarp(gmask);
        const u32 cBase = active ? cRp[row] : 0u;
        constexpr KeyT IDXMASK = ((KeyT)1 << IDXBITS) - 1;
        u32 running = 0;
        for (u32 i0 = 0; i0 < ops; i0 += G) {
            const u32 i = i0 + l;
            const bool valid = i < ops;
            KeyT key = 0;
            u32 col = 0;
            bool head = false;
            if (valid) {
                key = keys[i + (i >> 5)];
                col = (u32)(key >> IDXBITS);
                if (i == 0)
                    head = true;
                else {
                    const KeyT pk = keys[(i - 1) + ((i - 1) >> 5)];
                    head = (u32)(pk >> IDXBITS) != col;
                }
            }
            const u32 bal = __ballot_sync(gmask, head);
            const u32 gb = (G == 32) ? bal : ((bal >> (laneW - l)) & ((1u << (G & 31)) - 1u));
            if (head) {
                T sum = vals[(u32)(key & IDXMASK)];
                for (u32 j = i + 1; j < ops; ++j) {
                    const KeyT k2 = keys[j + (j >> 5)];
                    if ((u32)(k2 >> IDXBITS) != col) break;
                    sum = sum + vals[(u32)(k2 & IDXMASK)];
                }
                const u32 o = cBase + running + __popc(gb & ((1u << l) - 1u));
                cCi[o] = col;
                cV[o] = sum;
            }
            running += __popc(gb);
        }
    }
}

template <int G, int E, typename KeyT, typename T, int MODE>
__global__ void __launch_bounds__(SORT_BLOCK)
k_sort_rows(const u32 *__restrict__ perm, const u32 count, const u32 *__restrict__ aRp,
            const u32 *__restrict__ aCi, const T *__restrict__ aV, const u32 *__restrict__ bRp,
            const u32 *__restrict__ bCi, const T *__restrict__ bV, const u32 *__restrict__ rowOps,
            u32 *cRp, u32 *__restrict__ cCi, T *__restrict__ cV, const RowDesc *__restrict__ desc,
            const uint2 *__restrict__ aSeg, unsigned short *__restrict__ rankMap)
{
    sort_rows_body<G, E, KeyT, T, MODE>(blockIdx.x, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, cRp, cCi, cV, desc, aSeg,
                                        rankMap);
}

// The same with a capped grid (LaunchCtx::gridCap): every CTA loops over groups of SORT_BLOCK / G rows.  A separate
// kernel: the loop costs the one-shot kernel 8 registers per thread.
template <int G, int E, typename KeyT, typename T, int MODE>
__global__ void __launch_bounds__(SORT_BLOCK)
k_sort_rows_loop(const u32 *__restrict__ perm, const u32 count, const u32 *__restrict__ aRp,
                 const u32 *__restrict__ aCi, const T *__restrict__ aV, const u32 *__restrict__ bRp,
                 const u32 *__restrict__ bCi, const T *__restrict__ bV, const u32 *__restrict__ rowOps,
                 u32 *cRp, u32 *__restrict__ cCi, T *__restrict__ cV, const RowDesc *__restrict__ desc,
                 const uint2 *__restrict__ aSeg, unsigned short *__restrict__ rankMap)
{
    constexpr u32 GROUPS = SortLayout<G, E, KeyT, T, MODE>::GROUPS;
#pragma unroll 1
    for (u32 blk = blockIdx.x; blk * GROUPS < count; blk += gridDim.x) {
        sort_rows_body<G, E, KeyT, T, MODE>(blk, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, cRp, cCi, cV, desc, aSeg,
                                            rankMap);
        __syncwarp();   // the next row group reuses the lane groups' shared-memory regions
    }
}

template <int G, int E, typename KeyT, typename T, int MODE>
void launch_sort_rows(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi,
                      const T *aV, const u32 *bRp, const u32 *bCi, const T *bV, const u32 *rowOps, u32 *cRp,
                      u32 *cCi, T *cV, const RowDesc *desc = nullptr, const uint2 *aSeg = nullptr,
                      unsigned short *rankMap = nullptr)
{
    using L = SortLayout<G, E, KeyT, T, MODE>;
    auto kern = k_sort_rows<G, E, KeyT, T, MODE>;
    if (L::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    const u32 grid = (count + L::GROUPS - 1) / L::GROUPS;
    if constexpr (MODE == SORT_MAP && G == 32 && E >= 4 && sizeof(KeyT) == 4) {
        if (lc.gridCap && grid > lc.gridCap) {
            auto loop = k_sort_rows_loop<G, E, KeyT, T, MODE>;
            if (L::SMEM > 48 * 1024)
                cudaFuncSetAttribute(loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
            loop<<<lc.gridCap, SORT_BLOCK, L::SMEM, lc.stream>>>(perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, cRp, cCi, cV,
                                                                 desc, aSeg, rankMap);
            ++*lc.launches;
            return;
        }
    }
    kern<<<grid, SORT_BLOCK, L::SMEM, lc.stream>>>(perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, cRp, cCi, cV, desc, aSeg,
                                                   rankMap);
    ++*lc.launches;
}

// ------------------------------------------------------------------------------------------------
// Numeric phase of mapped rows, lane-group classes: one group of G lanes per row of <= N products.
// Gather as above; every product is scattered to its recorded rank in a shared staging row (plain
// stores when the row has no repeated column, otherwise shared-memory atomicAdd into a zeroed row),
// then the row is written to C coalesced.  No sort.
// ------------------------------------------------------------------------------------------------
template <int G, int N, typename T>
struct MapLayout {
    static constexpr int GROUPS = SORT_BLOCK / G;
    static constexpr size_t VAL_BYTES = (size_t)GROUPS * N * sizeof(T);
    static constexpr size_t SMEM = VAL_BYTES + (size_t)GROUPS * N * sizeof(u32);
};

template <int G, int N, typename T>
__global__ void __launch_bounds__(SORT_BLOCK)
k_map_rows(const RowDesc *__restrict__ desc, const u32 count, const uint2 *__restrict__ aSeg,
           const T *__restrict__ aV, const u32 *__restrict__ bCi, const T *__restrict__ bV,
           const unsigned short *__restrict__ rankMap, u32 *__restrict__ cCi, T *__restrict__ cV)
{
    using L = MapLayout<G, N, T>;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const u32 laneW = threadIdx.x & 31;
    const u32 l = threadIdx.x % G;
    const u32 grp = threadIdx.x / G;
    const u32 gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (laneW - l));
    T *outVal = reinterpret_cast<T *>(smemRaw) + (size_t)grp * N;
    u32 *outCol = reinterpret_cast<u32 *>(smemRaw + L::VAL_BYTES) + (size_t)grp * N;

    const u32 gidx = blockIdx.x * L::GROUPS + grp;
    const bool active = gidx < count;
    u32 aBeg = 0, aEnd = 0, cBase = 0, nnzRow = 0;
    bool folds = false;   // some column receives more than one product
    const unsigned short *map = rankMap;
    if (active) {
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(desc + gidx));
        const uint4 d1 = __ldg(reinterpret_cast<const uint4 *>(desc + gidx) + 1);
        aBeg = d0.x;
        aEnd = d0.x + d0.y;
        cBase = d1.x;
        nnzRow = d1.y;
        folds = nnzRow < d0.z;
        map = rankMap + (((u64)d1.w << 32) | d1.z);
    }
    if (folds) {
        for (u32 j = l; j < nnzRow; j += G) outVal[j] = (T)0;
        __syncwarp(gmask);
    }
    u32 base = 0;
    for (u32 ab = aBeg; ab < aEnd; ab += G) {
        const u32 ai = ab + l;
        u32 bs = 0, len = 0;
        T av = (T)0;
        if (ai < aEnd) {
            const uint2 seg = __ldg(aSeg + ai);
            bs = seg.x;
            len = seg.y - seg.x;
            av = __ldg(aV + ai);
        }
        u32 incl = len;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const u32 t = __shfl_up_sync(gmask, incl, d, G);
            if ((int)l >= d) incl += t;
        }
        const u32 total = __shfl_sync(gmask, incl, G - 1, G);
        const u32 qb = bs - (incl - len);   // position in B of product p of this entry = qb + p
        // U products per lane and iteration (whole-warp groups only): owners first, then all loads, then the stores
        constexpr int U = (N / G >= 4) ? 4 : 1;
        for (u32 p0 = 0; p0 < total; p0 += U * G) {
            u32 q[U], col[U], code[U];
            T oAv[U], bv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u > 0 && p0 + u * G >= total) break;   // group-uniform
                const u32 p = p0 + u * G + l;
                u32 lo = 0;
#pragma unroll
                for (int s = G / 2; s >= 1; s >>= 1) {
                    const u32 v = __shfl_sync(gmask, incl, lo + s - 1, G);
                    if (v <= p) lo += s;
                }
                q[u] = __shfl_sync(gmask, qb, lo, G) + p;
                oAv[u] = __shfl_sync(gmask, av, lo, G);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u > 0 && p0 + u * G >= total) break;
                const u32 p = p0 + u * G + l;
                const bool ok = p < total;
                code[u] = ok ? (u32)map[base + p] : 0u;
                col[u] = ok ? __ldg(bCi + q[u]) : 0u;
                bv[u] = ok ? __ldg(bV + q[u]) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u > 0 && p0 + u * G >= total) break;
                if (p0 + u * G + l < total) {
                    const T pr = oAv[u] * bv[u];
                    const u32 r = code[u] & MAP_RANK_MASK;
                    if (!folds) {
                        outVal[r] = pr;
                        outCol[r] = col[u];
                    } else {
                        if (!(code[u] & MAP_DUP)) outCol[r] = col[u];
                        atomicAdd(&outVal[r], pr);
                    }
                }
            }
        }
        base += total;
    }
    __syncwarp(gmask);
    for (u32 j = l; j < nnzRow; j += G) {
        cCi[cBase + j] = outCol[j];
        cV[cBase + j] = outVal[j];
    }
}

template <int G, int N, typename T>
void launch_map_rows(const LaunchCtx &lc, const RowDesc *desc, u32 count, const uint2 *aSeg, const T *aV,
                     const u32 *bCi, const T *bV, const unsigned short *rankMap, u32 *cCi, T *cV)
{
    using L = MapLayout<G, N, T>;
    auto kern = k_map_rows<G, N, T>;
    if (L::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    const u32 grid = (count + L::GROUPS - 1) / L::GROUPS;
    kern<<<grid, SORT_BLOCK, L::SMEM, lc.stream>>>(desc, count, aSeg, aV, bCi, bV, rankMap, cCi, cV);
    ++*lc.launches;
}

}  // namespace sb
