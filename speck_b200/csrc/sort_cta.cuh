// speck_b200/csrc/sort_cta.cuh -- multi-warp sort classes: one CTA of WARPS warps per row of C,
// N = WARPS * 32 * E products (E = 16 keys per lane).  Same three steps as sort_rows.cuh (gather / bitonic sort / fold +
// emit); each warp keeps 32*E keys in registers, merge stages whose partner distance
// reaches 32*E exchange keys through shared memory, everything below runs on shuffles/registers.
// Products are enumerated flat over the CTA (block scan of B-row lengths + binary search).
#pragma once
#include "sort_rows.cuh"

namespace sb {

template <int WARPS, int E, typename KeyT, typename T, bool NUMERIC>
struct CtaSortLayout {
    static constexpr int THREADS = WARPS * 32;
    static constexpr int WN = 32 * E;          // keys per warp
    static constexpr int N = WARPS * WN;
    static constexpr int NPAD = N + N / 32;
    static constexpr size_t KEY_BYTES = ((size_t)NPAD * sizeof(KeyT) + 15) / 16 * 16;
    static constexpr size_t VAL_BYTES = NUMERIC ? (size_t)N * sizeof(T) : 0;
    static constexpr size_t BATCH_BYTES = (size_t)THREADS * (8 + (NUMERIC ? sizeof(T) : 0));
    static constexpr size_t TAB_BYTES = (size_t)(N / 32) * sizeof(unsigned short);
    static constexpr size_t SMEM = KEY_BYTES + VAL_BYTES + BATCH_BYTES + TAB_BYTES;
};

template <int WARPS, int E, typename KeyT, typename T, bool NUMERIC>
__global__ void __launch_bounds__(WARPS * 32)
k_sort_rows_cta(const u32 *__restrict__ perm, const u32 count, const u32 *__restrict__ aRp,
                const u32 *__restrict__ aCi, const T *__restrict__ aV, const u32 *__restrict__ bRp,
                const u32 *__restrict__ bCi, const T *__restrict__ bV, const u32 *__restrict__ rowOps,
                u32 *cRp, u32 *__restrict__ cCi, T *__restrict__ cV)
{
    using L = CtaSortLayout<WARPS, E, KeyT, T, NUMERIC>;
    constexpr int THREADS = L::THREADS;
    constexpr int N = L::N;
    constexpr int WN = L::WN;
    constexpr int IDXBITS = Log2<N>::value;
    constexpr KeyT SENT = ~(KeyT)0;
    constexpr u32 FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    KeyT *keys = reinterpret_cast<KeyT *>(smemRaw);
    T *vals = reinterpret_cast<T *>(smemRaw + L::KEY_BYTES);
    u32 *sIncl = reinterpret_cast<u32 *>(smemRaw + L::KEY_BYTES + L::VAL_BYTES);
    u32 *sBs = sIncl + THREADS;
    T *sAv = reinterpret_cast<T *>(sBs + THREADS);
    unsigned short *sTab = reinterpret_cast<unsigned short *>(smemRaw + L::KEY_BYTES + L::VAL_BYTES + L::BATCH_BYTES);
    __shared__ u32 sWarp[32];
    __shared__ KeyT sLast[WARPS];

    const u32 tid = threadIdx.x;
    const u32 l = tid & 31, w = tid >> 5;
    const u32 row = perm[blockIdx.x];
    const u32 ops = rowOps[row];
    const u32 aBeg = aRp[row], aEnd = aRp[row + 1];

    // ---------------------------------------------------------------- gather (flat over the CTA)
    u32 base = 0;
    for (u32 ab = aBeg; ab < aEnd; ab += THREADS) {
        const u32 nb = min((u32)THREADS, aEnd - ab);
        u32 bs = 0, len = 0;
        if (tid < nb) {
            const u32 k = __ldg(aCi + ab + tid);
            bs = __ldg(bRp + k);
            len = __ldg(bRp + k + 1) - bs;
            if (NUMERIC) sAv[tid] = __ldg(aV + ab + tid);
        }
        // block inclusive scan of len
        u32 incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(FULL, incl, d);
            if (l >= (u32)d) incl += t;
        }
        if (l == 31) sWarp[w] = incl;
        __syncthreads();
        u32 warpBase = 0, total = 0;
#pragma unroll
        for (int i = 0; i < WARPS; ++i) {
            const u32 s = sWarp[i];
            if (i < (int)w) warpBase += s;
            total += s;
        }
        incl += warpBase;
        sIncl[tid] = incl;
        sBs[tid] = bs - (incl - len);  // q = sBs[owner] + p
        if (len) {  // owner table: sTab[b] = entry owning product 32*b (total <= N, so N/32 entries suffice)
            const u32 excl = incl - len;
            const u32 bLast = (incl - 1) >> 5;
            for (u32 b = (excl + 31) >> 5; b <= bLast; ++b) sTab[b] = (unsigned short)tid;
        }
        __syncthreads();
        // four products per thread and iteration: owners first, then all loads, then the stores
        for (u32 p0 = tid; p0 < total; p0 += 4 * THREADS) {
            u32 q[4], col[4];
            T av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const u32 p = p0 + u * THREADS;
                q[u] = 0xffffffffu;
                av[u] = (T)0;
                if (p < total) {
                    u32 lo = sTab[p >> 5];
                    while (sIncl[lo] <= p) ++lo;
                    q[u] = sBs[lo] + p;
                    if (NUMERIC) av[u] = sAv[lo];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                col[u] = q[u] != 0xffffffffu ? __ldg(bCi + q[u]) : 0u;
                bv[u] = (NUMERIC && q[u] != 0xffffffffu) ? __ldg(bV + q[u]) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (q[u] == 0xffffffffu) continue;
                const u32 gp = base + p0 + u * THREADS;
                if (NUMERIC) {
                    keys[gp] = ((KeyT)col[u] << IDXBITS) | (KeyT)gp;
                    vals[gp] = av[u] * bv[u];
                } else {
                    keys[gp] = (KeyT)col[u];
                }
            }
        }
        base += total;
        __syncthreads();
    }

    // ---------------------------------------------------------------- sort
    // logical index of (warp w, lane l, register r) = w*WN + l*E + r
    KeyT reg[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const u32 idx = w * WN + r * 32 + l;  // conflict-free read; any bijection is fine before sorting
        reg[r] = idx < ops ? keys[idx] : SENT;
    }
    __syncthreads();
    bitonic_sort_regs<32, E, KeyT>(reg, l, FULL);  // every warp: its WN keys ascending
    const u32 myBase = w * WN + l * E;
    // merge levels across warps.  The loops are NOT unrolled: one copy of the exchange code and of the
    // in-warp half cleaners keeps the kernel inside the instruction cache (a fully unrolled 1024-key
    // network stalls ~60 % of its issue slots on instruction fetch, profiles/r1_icache.md).
#pragma unroll 1
    for (int k = 2 * WN; k <= N; k <<= 1) {
        // mirrored stage (encoded as j == k), then cross-warp half cleaners j = k/4 ... WN
#pragma unroll 1
        for (int j = k; j >= WN; j >>= 1) {
            if (j == k >> 1) continue;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const u32 idx = myBase + r;
                keys[idx + (idx >> 5)] = reg[r];
            }
            __syncthreads();
            const u32 x = (j == k) ? (u32)(k - 1) : (u32)j;
            const bool lower = (j == k) ? ((w & (k / (2 * WN))) == 0) : ((w & (j / WN)) == 0);
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const u32 pidx = (myBase + r) ^ x;
                const KeyT o = keys[pidx + (pidx >> 5)];
                reg[r] = lower ? key_min(reg[r], o) : key_max(reg[r], o);
            }
            __syncthreads();
        }
        bitonic_half_cleaners<32, E, KeyT>(reg, l, FULL, WN / 2);
    }

    if (!NUMERIC) {
        // ------------------------------------------------------------ symbolic: distinct columns
        if (l == 31) sLast[w] = reg[E - 1];
        __syncthreads();
        KeyT prevLast = __shfl_up_sync(FULL, reg[E - 1], 1);
        if (l == 0 && w > 0) prevLast = sLast[w - 1];
        u32 cnt = 0;
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const KeyT prev = (r == 0) ? prevLast : reg[r - 1];
            const bool valid = reg[r] != SENT;
            const bool first = (r == 0) && (tid == 0);
            cnt += (valid && (first || reg[r] != prev)) ? 1u : 0u;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
        if (l == 0) sWarp[w] = cnt;
        __syncthreads();
        if (tid == 0) {
            u32 t = 0;
#pragma unroll
            for (int i = 0; i < WARPS; ++i) t += sWarp[i];
            cRp[row] = t;
        }
        return;
    } else {
        // ------------------------------------------------------------ numeric: fold runs, write C
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const u32 idx = myBase + r;
            keys[idx + (idx >> 5)] = reg[r];
        }
        __syncthreads();
        constexpr KeyT IDXMASK = ((KeyT)1 << IDXBITS) - 1;
        const u32 segBeg = w * WN;
        const u32 segEnd = min(ops, segBeg + WN);
        // pass 1: heads in this warp's segment
        u32 heads = 0;
        for (u32 i0 = segBeg; i0 < segEnd; i0 += 32) {
            const u32 i = i0 + l;
            bool head = false;
            if (i < segEnd) {
                const u32 col = (u32)(keys[i + (i >> 5)] >> IDXBITS);
                head = (i == 0) || ((u32)(keys[(i - 1) + ((i - 1) >> 5)] >> IDXBITS) != col);
            }
            heads += __popc(__ballot_sync(FULL, head));
        }
        if (l == 0) sWarp[w] = heads;
        __syncthreads();
        u32 running = 0;
#pragma unroll
        for (int i = 0; i < WARPS; ++i)
            if (i < (int)w) running += sWarp[i];
        const u32 cBase = cRp[row];
        // pass 2: emit
        for (u32 i0 = segBeg; i0 < segEnd; i0 += 32) {
            const u32 i = i0 + l;
            KeyT key = 0;
            u32 col = 0;
            bool head = false;
            if (i < segEnd) {
                key = keys[i + (i >> 5)];
                col = (u32)(key >> IDXBITS);
                head = (i == 0) || ((u32)(keys[(i - 1) + ((i - 1) >> 5)] >> IDXBITS) != col);
            }
            const u32 bal = __ballot_sync(FULL, head);
            if (head) {
                T sum = vals[(u32)(key & IDXMASK)];
                for (u32 j = i + 1; j < ops; ++j) {
                    const KeyT k2 = keys[j + (j >> 5)];
                    if ((u32)(k2 >> IDXBITS) != col) break;
                    sum = sum + vals[(u32)(k2 & IDXMASK)];
                }
                const u32 o = cBase + running + __popc(bal & ((1u << l) - 1u));
                cCi[o] = col;
                cV[o] = sum;
            }
            running += __popc(bal);
        }
    }
}

template <int WARPS, int E, typename KeyT, typename T, bool NUMERIC>
void launch_sort_rows_cta(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi,
                          const T *aV, const u32 *bRp, const u32 *bCi, const T *bV, const u32 *rowOps, u32 *cRp,
                          u32 *cCi, T *cV)
{
    using L = CtaSortLayout<WARPS, E, KeyT, T, NUMERIC>;
    auto kern = k_sort_rows_cta<WARPS, E, KeyT, T, NUMERIC>;
    if (L::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    kern<<<count, L::THREADS, L::SMEM, lc.stream>>>(perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, cRp, cCi, cV);
    ++*lc.launches;
}

}  // namespace sb
