// speck_b200/csrc/rank_flat.cuh -- "flat" symbolic kernel of the rank row classes (round 2).
//
// Same two-level column bitmap as rank_cta.cuh (replaces the symbolic hash phase of the reference,
// include/GPU/spECK_HashSpGEMM.cuh:919-1119), different data movement:
//   * the row's B.col_ids segments are STAGED into shared memory in product order by asynchronous copies
//     (cp.async, 4 bytes per lane, fire and forget: no registers hold products, every B segment of the row is in
//     flight at once) instead of per-thread blocked register slots filled through a per-product search (SegWalk);
//   * every later pass is a flat strided loop over the staged row: consecutive lanes own consecutive products, no
//     owner lookups, no predicated register slots, and the rank map is stored coalesced.
// Passes: (1) top-level bit per touched 32-column chunk, scan; (2) leaf bit per column (the product whose atomicOr
// sets it first owns the output slot, later ones are flagged), scan; (3) rank = leaf prefix + popcount below.
#pragma once
#include "rank_cta.cuh"

namespace sb {

__device__ __forceinline__ void cp_async_u32(u32 *smemDst, const u32 *gsrc)
{
    const u32 d = (u32)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <int THREADS, int E>
struct FlatLayout {
    static constexpr size_t al(size_t b) { return (b + 15) / 16 * 16; }
    static constexpr size_t CAP = (size_t)THREADS * E;
    static constexpr size_t STASH = 0;
    static constexpr size_t TOP = STASH + al(CAP * 4);
    static constexpr size_t TOPPRE = TOP + RANK_TOP_WORDS * 4;
    static constexpr size_t LEAF = TOPPRE + RANK_TOP_WORDS * 2;
    static constexpr size_t LEAFPRE = LEAF + al(CAP * 4);
    static constexpr size_t SBS = LEAFPRE + al(CAP * 2);
    static constexpr size_t SOFF = SBS + al(THREADS * 4);
    static constexpr size_t SLEN = SOFF + al(THREADS * 4);
    static constexpr size_t SMEM = SLEN + al(THREADS * 4);
};

// lanes per B segment while staging: short B rows dominate by count, eight lanes (one 32-byte sector) keep most
// lanes busy on them
#ifndef SB_FLAT_SG
#define SB_FLAT_SG 8
#endif
constexpr int FLAT_SG = SB_FLAT_SG;

template <int THREADS, int E>
__global__ void __launch_bounds__(THREADS)
k_rank_flat(const RowDesc *__restrict__ desc, const uint2 *__restrict__ aSeg, const u32 *__restrict__ bCi,
            unsigned short *__restrict__ rankMap, u32 *__restrict__ cRp)
{
    using L = FlatLayout<THREADS, E>;
    constexpr int TOPG = (RANK_TOP_WORDS / 4 + THREADS - 1) / THREADS;
    constexpr int LEAFG = (E + 3) / 4;
    constexpr u32 DUPBIT = 0x80000000u;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    u32 *stash = reinterpret_cast<u32 *>(smemRaw + L::STASH);
    u32 *top = reinterpret_cast<u32 *>(smemRaw + L::TOP);
    unsigned short *topPre = reinterpret_cast<unsigned short *>(smemRaw + L::TOPPRE);
    u32 *leaf = reinterpret_cast<u32 *>(smemRaw + L::LEAF);
    unsigned short *leafPre = reinterpret_cast<unsigned short *>(smemRaw + L::LEAFPRE);
    u32 *sBs = reinterpret_cast<u32 *>(smemRaw + L::SBS);
    u32 *sOff = reinterpret_cast<u32 *>(smemRaw + L::SOFF);
    u32 *sLen = reinterpret_cast<u32 *>(smemRaw + L::SLEN);
    __shared__ u32 sWarp[33];

    const u32 tid = threadIdx.x;
    const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(desc + blockIdx.x));
    const uint4 d1 = __ldg(reinterpret_cast<const uint4 *>(desc + blockIdx.x) + 1);
    const u32 aBeg = d0.x, aEnd = d0.x + d0.y, n = d0.z, row = d0.w;
    const u32 cmin = d1.x, cmax = d1.y;
    unsigned short *map = rankMap + (((u64)d1.w << 32) | d1.z);
    const u32 topWords = ((cmax - cmin) >> 10) + 1;   // <= RANK_TOP_WORDS: the host checks cols(B)

    // ---------------------------------------------------------------- stage the row's columns, product order
    u32 base = 0;
#pragma unroll 1
    for (u32 ab = aBeg; ab < aEnd; ab += THREADS) {
        const u32 nb = min((u32)THREADS, aEnd - ab);
        u32 bs = 0, len = 0;
        if (tid < nb) {
            const uint2 seg = __ldg(aSeg + ab + tid);
            bs = seg.x;
            len = seg.y - seg.x;
        }
        u32 total;
        const u32 excl = cta_exclusive_scan<THREADS>(len, sWarp, &total);
        sBs[tid] = bs;
        sOff[tid] = base + excl;
        sLen[tid] = len;
        if (ab == aBeg) rank_clear_level<THREADS, TOPG>(top, topWords);
        __syncthreads();
        const u32 g = tid / FLAT_SG, gl = tid % FLAT_SG;
#pragma unroll 1
        for (u32 e = g; e < nb; e += THREADS / FLAT_SG) {
            const u32 b = sBs[e], l = sLen[e], o = sOff[e];
            for (u32 j = gl; j < l; j += FLAT_SG) cp_async_u32(stash + o + j, bCi + b + j);
        }
        base += total;
        if (aEnd - ab > (u32)THREADS) __syncthreads();   // the segment tables are rewritten by the next batch
    }
    cp_async_wait_all();
    __syncthreads();

    // ---------------------------------------------------------------- pass 1: touched 32-column chunks
#pragma unroll
    for (int k = 0; k < E; ++k) {
        if ((u32)k * THREADS >= n) break;   // CTA-uniform
        const u32 i = (u32)k * THREADS + tid;
        if (i < n) {
            const u32 c = stash[i] - cmin;
            atomicOr(&top[c >> 10], 1u << ((c >> 5) & 31));
        }
    }
    __syncthreads();
    const u32 leaves = rank_scan_level<THREADS, TOPG, true>(top, topPre, topWords, sWarp);
    rank_clear_level<THREADS, LEAFG>(leaf, leaves);
    __syncthreads();

    // ---------------------------------------------------------------- pass 2: leaf bits, first-of-column flags
#pragma unroll
    for (int k = 0; k < E; ++k) {
        if ((u32)k * THREADS >= n) break;
        const u32 i = (u32)k * THREADS + tid;
        if (i < n) {
            const u32 c = stash[i] - cmin;
            const u32 tw = c >> 10, tb = (c >> 5) & 31;
            const u32 s = topPre[tw] + __popc(top[tw] & ((1u << tb) - 1u));
            const u32 bit = 1u << (c & 31);
            const u32 old = atomicOr(&leaf[s], bit);
            stash[i] = (s << 5) | (c & 31u) | ((old & bit) ? DUPBIT : 0u);
        }
    }
    __syncthreads();
    const u32 distinct = rank_scan_level<THREADS, LEAFG, true>(leaf, leafPre, leaves, sWarp);
    if (tid == 0) cRp[row] = distinct;
    __syncthreads();

    // ---------------------------------------------------------------- pass 3: ranks -> map (coalesced)
#pragma unroll
    for (int k = 0; k < E; ++k) {
        if ((u32)k * THREADS >= n) break;
        const u32 i = (u32)k * THREADS + tid;
        if (i < n) {
            const u32 v = stash[i];
            const u32 s = (v & ~DUPBIT) >> 5, b = v & 31u;
            const u32 rank = leafPre[s] + __popc(leaf[s] & ((1u << b) - 1u));
            map[i] = (unsigned short)(rank | ((v & DUPBIT) ? MAP_DUP : 0u));
        }
    }
}

template <int THREADS, int E>
void launch_rank_flat_shape(const LaunchCtx &lc, const RowDesc *desc, u32 count, const uint2 *aSeg, const u32 *bCi,
                            unsigned short *rankMap, u32 *cRp)
{
    using L = FlatLayout<THREADS, E>;
    auto kern = k_rank_flat<THREADS, E>;
    if (L::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    kern<<<count, THREADS, L::SMEM, lc.stream>>>(desc, aSeg, bCi, rankMap, cRp);
    ++*lc.launches;
}

// ------------------------------------------------------------------------------------------------
// Count-only symbolic by hashing (experiment, option "hash_count"): open-addressed set of the row's
// columns in shared memory, one atomicCAS per product -- the reference's HashMapNoValue
// (include/HashMap.cuh:136-230) without its per-row occupancy counters.  No order information.
// ------------------------------------------------------------------------------------------------
template <int THREADS, int LOGCAP>
__global__ void __launch_bounds__(THREADS)
k_hash_count(const u32 *__restrict__ perm, const u32 *__restrict__ aRp, const u32 *__restrict__ aCi,
             const u32 *__restrict__ bRp, const u32 *__restrict__ bCi, u32 *__restrict__ cRp)
{
    constexpr u32 SLOTS = 1u << LOGCAP;
    constexpr u32 EMPTY = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    u32 *table = reinterpret_cast<u32 *>(smemRaw);
    __shared__ u32 sCnt;
    const u32 tid = threadIdx.x;
    const u32 row = perm[blockIdx.x];
    for (u32 j = tid; j < SLOTS / 4; j += THREADS) reinterpret_cast<uint4 *>(table)[j] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
    if (tid == 0) sCnt = 0;
    __syncthreads();
    const u32 aBeg = aRp[row], aEnd = aRp[row + 1];
    const u32 g = tid / FLAT_SG, gl = tid % FLAT_SG;
    u32 cnt = 0;
#pragma unroll 1
    for (u32 e = aBeg + g; e < aEnd; e += THREADS / FLAT_SG) {
        const u32 k = __ldg(aCi + e);
        const u32 bs = __ldg(bRp + k), be = __ldg(bRp + k + 1);
        for (u32 q = bs + gl; q < be; q += FLAT_SG) {
            const u32 c = __ldg(bCi + q);
            u32 h = (c * 2654435761u) >> (32 - LOGCAP);
            while (true) {
                const u32 old = atomicCAS(&table[h], EMPTY, c);
                if (old == EMPTY) { ++cnt; break; }
                if (old == c) break;
                h = (h + 1) & (SLOTS - 1);
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((tid & 31) == 0 && cnt) atomicAdd(&sCnt, cnt);
    __syncthreads();
    if (tid == 0) cRp[row] = sCnt;
}

template <int THREADS, int LOGCAP>
void launch_hash_count_shape(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi,
                             const u32 *bRp, const u32 *bCi, u32 *cRp)
{
    auto kern = k_hash_count<THREADS, LOGCAP>;
    const size_t smem = (size_t)4 << LOGCAP;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<count, THREADS, smem, lc.stream>>>(perm, aRp, aCi, bRp, bCi, cRp);
    ++*lc.launches;
}

}  // namespace sb
