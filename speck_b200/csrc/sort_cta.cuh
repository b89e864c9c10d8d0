// speck_b200/csrc/sort_cta.cuh -- multi-warp sort classes: one CTA of WARPS warps per row of C,
// N = WARPS * 32 * E products (E = 16 keys per lane).  Same three steps as sort_rows.cuh (gather / bitonic sort / fold +
// emit); each warp keeps 32*E keys in registers, merge stages whose partner distance
// reaches 32*E exchange keys through shared memory, everything below runs on shuffles/registers.
// Products are enumerated flat over the CTA (block scan of B-row lengths + binary search).
#pragma once
#include "sort_rows.cuh"

namespace sb {

// Shared-memory layout for a CTA of `warps` warps (runtime, WMAX/2 < warps <= WMAX):
//   keys[n + n/32] | vals[n] (numeric) | sIncl[threads] | sBs[threads] | sAv[threads] (numeric) | sTab[n/32]
template <int E, typename KeyT, typename T, bool NUMERIC>
struct CtaSortLayout {
    static constexpr int WN = 32 * E;          // keys per warp
    __host__ __device__ static size_t key_bytes(int warps) { return ((size_t)(warps * WN + warps * WN / 32) * sizeof(KeyT) + 15) / 16 * 16; }
    __host__ __device__ static size_t val_bytes(int warps) { return NUMERIC ? (size_t)warps * WN * sizeof(T) : 0; }
    __host__ __device__ static size_t batch_bytes(int warps) { return (size_t)warps * 32 * (8 + (NUMERIC ? sizeof(T) : 0)); }
    __host__ __device__ static size_t tab_bytes(int warps) { return (size_t)(warps * WN / 32) * sizeof(unsigned short); }
    __host__ __device__ static size_t smem(int warps) { return key_bytes(warps) + val_bytes(warps) + batch_bytes(warps) + tab_bytes(warps); }
};

// WMAX = power of two >= the CTA's warp count: fixes the index bits of the keys and the depth of the
// merge network; warps beyond blockDim.x/32 are virtual (all +inf) and every exchange with them is a
// no-op, so a row with 1300 products is sorted by 3 warps (1536 slots) instead of 4 (2048).
template <int WMAX, int E, typename KeyT, typename T, bool NUMERIC>
__global__ void __launch_bounds__(WMAX * 32)
k_sort_rows_cta(const u32 *__restrict__ perm, const u32 count, const u32 *__restrict__ aRp,
                const u32 *__restrict__ aCi, const T *__restrict__ aV, const u32 *__restrict__ bRp,
                const u32 *__restrict__ bCi, const T *__restrict__ bV, const u32 *__restrict__ rowOps,
                u32 *cRp, u32 *__restrict__ cCi, T *__restrict__ cV)
{
    using L = CtaSortLayout<E, KeyT, T, NUMERIC>;
    const int WARPS = blockDim.x >> 5;
    const u32 THREADS = blockDim.x;
    constexpr int WN = L::WN;
    constexpr int NPOW2 = WMAX * WN;
    constexpr int IDXBITS = Log2<NPOW2>::value;
    constexpr KeyT SENT = ~(KeyT)0;
    constexpr u32 FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    KeyT *keys = reinterpret_cast<KeyT *>(smemRaw);
    T *vals = reinterpret_cast<T *>(smemRaw + L::key_bytes(WARPS));
    u32 *sIncl = reinterpret_cast<u32 *>(smemRaw + L::key_bytes(WARPS) + L::val_bytes(WARPS));
    u32 *sBs = sIncl + THREADS;
    T *sAv = reinterpret_cast<T *>(sBs + THREADS);
    unsigned short *sTab = reinterpret_cast<unsigned short *>(smemRaw + L::key_bytes(WARPS) + L::val_bytes(WARPS) + L::batch_bytes(WARPS));
    __shared__ u32 sWarp[32];
    __shared__ KeyT sLast[WMAX];

    const u32 tid = threadIdx.x;
    const u32 l = tid & 31, w = tid >> 5;
    const u32 row = perm[blockIdx.x];
    const u32 ops = rowOps[row];
    const u32 aBeg = aRp[row], aEnd = aRp[row + 1];

    // ---------------------------------------------------------------- gather (flat over the CTA)
    u32 base = 0;
    for (u32 ab = aBeg; ab < aEnd; ab += THREADS) {
        const u32 nb = min(THREADS, aEnd - ab);
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
        for (int i = 0; i < WMAX; ++i) {
            const u32 s = i < WARPS ? sWarp[i] : 0u;
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
    for (int k = 2 * WN; k <= NPOW2; k <<= 1) {
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
            // partner warp (same for all registers of a lane: x flips whole-warp bits, and for the mirror
            // also the in-warp bits); a virtual partner holds +inf and always has the higher index
            if ((((myBase) ^ x) / WN) < (u32)WARPS) {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const u32 pidx = (myBase + r) ^ x;
                    const KeyT o = keys[pidx + (pidx >> 5)];
                    reg[r] = lower ? key_min(reg[r], o) : key_max(reg[r], o);
                }
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
        for (int i = 0; i < (int)w; ++i) running += sWarp[i];
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

template <int WMAX, int E, typename KeyT, typename T, bool NUMERIC>
void launch_sort_rows_cta(const LaunchCtx &lc, int warps, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi,
                          const T *aV, const u32 *bRp, const u32 *bCi, const T *bV, const u32 *rowOps, u32 *cRp,
                          u32 *cCi, T *cV)
{
    using L = CtaSortLayout<E, KeyT, T, NUMERIC>;
    auto kern = k_sort_rows_cta<WMAX, E, KeyT, T, NUMERIC>;
    const size_t smem = L::smem(warps);
    if (L::smem(WMAX) > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem(WMAX));
    kern<<<count, warps * 32, smem, lc.stream>>>(perm, count, aRp, aCi, aV, bRp, bCi, bV, rowOps, cRp, cCi, cV);
    ++*lc.launches;
}

// CTA sort class index (0..14) -> warps (2..16) and the power-of-two network size
inline int cta_class_warps(int ctaClass) { return ctaClass + 2; }

}  // namespace sb
