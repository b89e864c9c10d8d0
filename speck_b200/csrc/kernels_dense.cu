// speck_b200/csrc/kernels_dense.cu -- "dense" row class: one CTA per row of C, column bitmap in
// shared memory.
//
// Replaces denseSpGEMMCount / denseSpGEMMNumeric and the class-0/1/2 hash kernels of the
// reference (include/GPU/spECK_HashSpGEMM.cuh:1300-1711, 591-866) for rows with more products
// than the sort classes take.  Differences by design (B200: 227 KB shared memory per CTA):
//   * the bitmap covers a window of up to 2^20 columns (128 KB) instead of a ~5.8 k-column dense
//     value window, so most matrices need ONE pass per row instead of range/5.8k passes;
//   * the bitmap is cleared sparsely: a byte per 128-column chunk (plain idempotent stores, no
//     atomics) records which chunks were touched; only those are counted / cleared;
//   * the products of a row are enumerated flat over the whole CTA (block scan of the B-row
//     lengths + binary search), so lanes stay busy whatever the B-row lengths are;
//   * sorted output comes from popcount ranks (chunk prefix + in-chunk popc), values are
//     accumulated with fp RED (red.global.add) straight into the zero-initialised C row, so no
//     value window, no shared-memory CAS loops and no separate sorting pass are needed.
// Rows are pulled from a device-side queue (atomic counter), CTAs are persistent.
#include "common.cuh"
#include "dense_seq.cuh"

namespace sb {

constexpr int CHUNK_WORDS = 4;       // 128 columns per chunk = one 16-byte shared load
constexpr int DENSE_MAX_WIN_BITS = 20;
constexpr int DENSE_MIN_WIN_BITS = 14;  // 16384 columns

int dense_window_bits(u64 colsB)
{
    int b = DENSE_MIN_WIN_BITS;
    while (b < DENSE_MAX_WIN_BITS && (1ull << b) < colsB) ++b;
    return b;
}

static size_t dense_smem_bytes(int winBits, int threads, size_t valBytes)
{
    const size_t words = (size_t)1 << (winBits - 5);
    const size_t chunks = words / CHUNK_WORDS;
    return (words + chunks) * sizeof(u32) + chunks /* touched bytes */ + (size_t)threads * (8 + valBytes);
}

__device__ __forceinline__ u32 lower_bound_dev(const u32 *__restrict__ a, u32 lo, u32 hi, u32 key)
{
    while (lo < hi) {
        const u32 mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// exclusive scan of one u32 per thread; returns the exclusive prefix, *total = block sum.
template <int THREADS>
__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32 *sWarp, u32 *total)
{
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (u32)d) incl += t;
    }
    __syncthreads();  // sWarp may still be read from a previous use
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    constexpr int NW = THREADS / 32;
    u32 wv = (lane < NW) ? sWarp[lane] : 0u;
    u32 winc = wv;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= (u32)d) winc += t;
    }
    const u32 warpBase = __shfl_sync(0xffffffffu, winc - wv, warp);
    *total = __shfl_sync(0xffffffffu, winc, NW - 1);
    return warpBase + incl - v;
}

// Shared-memory layout of one CTA:
//   bitmap[WWORDS] u32 | chunkPrefix[NCHUNK] u32 | touched[NCHUNK] u8 | sIncl[THREADS] u32 | sBs[THREADS] u32 |
//   sAv[THREADS] T | svals[SVALS] T
// Windows are relative to the row: the first one starts at the row's smallest column (rounded down to
// a chunk), so a banded row needs one small window wherever its band lies.
// SVALS > 0: rows (windows) with at most SVALS distinct columns accumulate their values in shared
// memory (atomicAdd on svals[rank]) and are written out coalesced; this is the path of high-compression
// rows (FEM-like matrices), where a RED + column store per product would multiply the traffic to C.
template <int THREADS, typename T, bool NUMERIC, int SVALS>
__global__ void __launch_bounds__(THREADS)
k_dense_rows(const u32 *__restrict__ perm, const u32 count, u32 *rowCounter, const u32 *__restrict__ aRp,
             const u32 *__restrict__ aCi, const T *__restrict__ aV, const u32 *__restrict__ bRp,
             const u32 *__restrict__ bCi, const T *__restrict__ bV, const int winBits,
             const u32 *__restrict__ rowMin, const u32 *__restrict__ rowMax, u32 *bitmapStore, u32 *cRp,
             u32 *__restrict__ cCi, T *cV, const u32 seqMax, const u32 *__restrict__ rowOps, u32 *seqRows, const u32 fold,
             const bool testSet)
{
    extern __shared__ __align__(16) u32 dsm[];
    const u32 W = 1u << winBits;
    const u32 WWORDS = W >> 5;
    const u32 NCHUNK = WWORDS / CHUNK_WORDS;
    u32 *bitmap = dsm;
    u32 *chunkPrefix = bitmap + WWORDS;
    unsigned char *touched = reinterpret_cast<unsigned char *>(chunkPrefix + NCHUNK);
    u32 *sIncl = reinterpret_cast<u32 *>(touched + NCHUNK);
    u32 *sBs = sIncl + THREADS;
    T *sAv = reinterpret_cast<T *>(sBs + THREADS);
    T *svals = sAv + THREADS;
    constexpr u32 TAB = 2 * THREADS;   // owner table: one entry per 32 products of a batch
    __shared__ unsigned short sTab[TAB];
    __shared__ u32 sWarp[32];
    __shared__ u32 sRow;

    const u32 tid = threadIdx.x;
    for (u32 i = tid; i < WWORDS; i += THREADS) bitmap[i] = 0;
    for (u32 i = tid; i < NCHUNK / 4; i += THREADS) reinterpret_cast<u32 *>(touched)[i] = 0;
    const u32 CPT = (NCHUNK + THREADS - 1) / THREADS;  // consecutive chunks per thread
    const u32 chBeg = tid * CPT;
    const u32 chEnd = min(NCHUNK, chBeg + CPT);

    // one batch = up to THREADS entries of the A row: B-row bounds (trimmed to the window when the row
    // needs several), inclusive scan of the lengths; products of the batch are then enumerated flat
    // (p -> owner by binary search), so every thread gets the same number of products.
    auto load_batch = [&](u32 ab, u32 aEnd, bool trim, u32 winLo, u32 winHi, u32 &nb) -> u32 {
        nb = min((u32)THREADS, aEnd - ab);
        u32 bs = 0, len = 0;
        if (tid < nb) {
            const u32 k = __ldg(aCi + ab + tid);
            bs = __ldg(bRp + k);
            u32 be = __ldg(bRp + k + 1);
            if (trim) {
                bs = lower_bound_dev(bCi, bs, be, winLo);
                be = lower_bound_dev(bCi, bs, be, winHi);
            }
            len = be - bs;
            if (NUMERIC) sAv[tid] = __ldg(aV + ab + tid);
        }
        u32 total;
        const u32 excl = block_exclusive_scan<THREADS>(len, sWarp, &total);
        sIncl[tid] = excl + len;
        sBs[tid] = bs - excl;  // q = sBs[owner] + p
        // owner table: sTab[b] = entry that owns product 32*b (each entry fills the blocks it starts)
        if (len && total <= 32u * TAB) {
            const u32 bLast = (excl + len - 1) >> 5;
            for (u32 b = (excl + 31) >> 5; b <= bLast; ++b) sTab[b] = (unsigned short)tid;
        }
        __syncthreads();
        return total;
    };
    // owner of product p: table entry, then a short forward walk (<= 32 steps); binary search only for
    // batches with more than 32*TAB products
    auto owner_of = [&](u32 p, u32 nb, u32 total) -> u32 {
        if (total <= 32u * TAB) {
            u32 o = sTab[p >> 5];
            while (sIncl[o] <= p) ++o;
            return o;
        }
        u32 lo = 0, hi = nb;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (sIncl[mid] <= p) lo = mid + 1; else hi = mid;
        }
        return lo;
    };

    while (true) {
        __syncthreads();
        if (tid == 0) sRow = atomicAdd(rowCounter, 1u);
        __syncthreads();
        const u32 ri = sRow;
        if (ri >= count) break;
        const u32 row = perm[ri];
        if (NUMERIC && seqMax && dense_seq_takes(rowOps[row], cRp[row + 1] - cRp[row], seqMax, fold)) continue;   // k_dense_seq's row
        const u32 aBeg = aRp[row], aEnd = aRp[row + 1];
        const bool oneBatch = (aEnd - aBeg) <= (u32)THREADS;
        const u32 colMin = rowMin[row], colMax = rowMax[row];
        const u32 base0 = colMin & ~(u32)(CHUNK_WORDS * 32 - 1);
        const bool trim = (colMax - base0) >= W;  // more than one window
        const u32 rowStart = NUMERIC ? cRp[row] : 0u;
        u32 winBase = 0;

        for (u32 winLo = base0; winLo <= colMax; winLo += W) {
            const u32 winHi = winLo + W;
            u32 nb = 0, total = 0;
            const u32 extWords = ((min(colMax, winHi - 1) - winLo) >> 5) + 1;
            u32 *store = (bitmapStore && !trim) ? bitmapStore + (size_t)ri * WWORDS : nullptr;
            if (NUMERIC && store) {
                // ---------------------------------------------- the symbolic phase kept this row's bitmap
                for (u32 i = tid; i < extWords; i += THREADS) {
                    const u32 wv = store[i];
                    bitmap[i] = wv;
                    if (wv) touched[i >> 2] = 1;
                }
                if (oneBatch) total = load_batch(aBeg, aEnd, trim, winLo, winHi, nb);
                __syncthreads();
            } else {
                // ---------------------------------------------- pass A: set column bits
                // (a variant with 16 lanes per B row and no owner search was measured on the cant-shaped matrix: symbolic
                // 0.86 -> 1.01 ms -- consecutive columns of a FEM cluster hit the same bitmap word and the atomics serialise)
                for (u32 ab = aBeg; ab < aEnd; ab += THREADS) {
                    total = load_batch(ab, aEnd, trim, winLo, winHi, nb);
                    // four products per thread and iteration: the four column loads are in flight together
                    for (u32 p0 = tid; p0 < total; p0 += 4 * THREADS) {
                        u32 q[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const u32 p = p0 + u * THREADS;
                            q[u] = p < total ? sBs[owner_of(p, nb, total)] + p : 0xffffffffu;
                        }
                        u32 cc[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) cc[u] = q[u] != 0xffffffffu ? __ldg(bCi + q[u]) : 0u;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (q[u] != 0xffffffffu) {
                                const u32 c = cc[u] - winLo;
                                // test before set: in rows that fold (FEM-like: ~18 products per column) most bits are
                                // set already and a plain shared load is far cheaper than the atomic (option test_set)
                                const u32 bit = 1u << (c & 31);
                                if (!testSet || !(bitmap[c >> 5] & bit)) {
                                    atomicOr(&bitmap[c >> 5], bit);
                                    touched[c >> 7] = 1;
                                }
                            }
                        }
                    }
                    __syncthreads();
                }
                if (!NUMERIC && store) {  // keep the bitmap for the numeric phase
                    for (u32 i = tid; i < extWords; i += THREADS) store[i] = bitmap[i];
                    __syncthreads();
                }
            }
            // ------------------------------------------------ chunk counts (touched chunks only)
            u32 tsum = 0;
            for (u32 ch = chBeg; ch < chEnd; ++ch) {
                if (touched[ch]) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(&bitmap[ch * CHUNK_WORDS]);
                    if (NUMERIC) chunkPrefix[ch] = tsum;
                    tsum += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
                    if (!NUMERIC) {  // symbolic: this thread is the only reader left -> clear now
                        *reinterpret_cast<uint4 *>(&bitmap[ch * CHUNK_WORDS]) = make_uint4(0, 0, 0, 0);
                        touched[ch] = 0;
                    }
                }
            }
            u32 winTotal;
            const u32 texcl = block_exclusive_scan<THREADS>(tsum, sWarp, &winTotal);

            if (NUMERIC) {
                const bool local = SVALS > 0 && winTotal <= (u32)SVALS;
                const u32 outBase = rowStart + winBase;
                for (u32 ch = chBeg; ch < chEnd; ++ch) {
                    if (touched[ch]) {
                        const u32 pfx = chunkPrefix[ch] + texcl;   // window-local rank of the chunk's first column
                        chunkPrefix[ch] = pfx;
                        if (local) {  // sorted column ids straight from the bitmap
                            const uint4 v = *reinterpret_cast<const uint4 *>(&bitmap[ch * CHUNK_WORDS]);
                            const u32 wv[4] = {v.x, v.y, v.z, v.w};
                            u32 pos = outBase + pfx;
                            const u32 colBase = winLo + ch * (CHUNK_WORDS * 32);
#pragma unroll
                            for (int wi = 0; wi < 4; ++wi) {
                                u32 bits = wv[wi];
                                while (bits) {
                                    const u32 b = __ffs(bits) - 1;
                                    bits &= bits - 1;
                                    cCi[pos++] = colBase + wi * 32 + b;
                                }
                            }
                        }
                    }
                }
                // zero the accumulators of this window (shared or, for big windows, C itself)
                if (local) {
                    for (u32 j = tid; j < winTotal; j += THREADS) svals[j] = (T)0;
                } else {
                    for (u32 j = tid; j < winTotal; j += THREADS) cV[outBase + j] = (T)0;
                    __threadfence_block();
                }
                __syncthreads();
                // -------------------------------------------- pass B: rank -> accumulate the product
                for (u32 ab = aBeg; ab < aEnd; ab += THREADS) {
                    if (!oneBatch) total = load_batch(ab, aEnd, trim, winLo, winHi, nb);
                    for (u32 p0 = tid; p0 < total; p0 += 4 * THREADS) {
                      u32 qq[4], colv[4];
                      T avv[4], bvv[4];
#pragma unroll
                      for (int u = 0; u < 4; ++u) {
                          const u32 p = p0 + u * THREADS;
                          qq[u] = 0xffffffffu;
                          avv[u] = (T)0;
                          if (p < total) {
                              const u32 o = owner_of(p, nb, total);
                              qq[u] = sBs[o] + p;
                              avv[u] = sAv[o];
                          }
                      }
#pragma unroll
                      for (int u = 0; u < 4; ++u) {
                          colv[u] = qq[u] != 0xffffffffu ? __ldg(bCi + qq[u]) : 0u;
                          bvv[u] = qq[u] != 0xffffffffu ? __ldg(bV + qq[u]) : (T)0;
                      }
#pragma unroll
                      for (int u = 0; u < 4; ++u) {
                        if (qq[u] == 0xffffffffu) continue;
                        const u32 col = colv[u];
                        const T prod = avv[u] * bvv[u];
                        const u32 c = col - winLo;
                        const u32 w = c >> 5;
                        const u32 ch = w >> 2;
                        const u32 wi = w & 3;
                        const uint4 v = *reinterpret_cast<const uint4 *>(&bitmap[ch * CHUNK_WORDS]);
                        const u32 below = (wi > 0 ? __popc(v.x) : 0) + (wi > 1 ? __popc(v.y) : 0) +
                                          (wi > 2 ? __popc(v.z) : 0);
                        const u32 word = wi == 0 ? v.x : (wi == 1 ? v.y : (wi == 2 ? v.z : v.w));
                        const u32 rank = chunkPrefix[ch] + below + __popc(word & ((1u << (c & 31)) - 1u));
                        if (local) {
                            atomicAdd(&svals[rank], prod);
                        } else {
                            cCi[outBase + rank] = col;  // every product of a column writes the same value
                            atomicAdd(&cV[outBase + rank], prod);
                        }
                      }
                    }
                    __syncthreads();
                }
                if (local)
                    for (u32 j = tid; j < winTotal; j += THREADS) cV[outBase + j] = svals[j];
                // -------------------------------------------- sparse clear
                for (u32 ch = chBeg; ch < chEnd; ++ch) {
                    if (touched[ch]) {
                        *reinterpret_cast<uint4 *>(&bitmap[ch * CHUNK_WORDS]) = make_uint4(0, 0, 0, 0);
                        touched[ch] = 0;
                    }
                }
            }
            __syncthreads();
            winBase += winTotal;
        }
        if (!NUMERIC && tid == 0) {
            cRp[row] = winBase;
            if (seqRows && dense_seq_takes(rowOps[row], winBase, (u32)DENSE_SEQ_MAX, fold)) atomicAdd(&seqRows[winBase <= 512u ? 0 : 1], 1u);
        }
    }
}

template <int THREADS, typename T, bool NUMERIC, int SVALS>
static void launch_dense_t(const LaunchCtx &lc, int winBits, const u32 *perm, u32 count, u32 *rowCounter,
                           const u32 *aRp, const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi, const T *bV,
                           const u32 *rowMin, const u32 *rowMax, u32 *bitmapStore, u32 *cRp, u32 *cCi, T *cV,
                           u32 seqMax = 0, const u32 *rowOps = nullptr, u32 *seqRows = nullptr, u32 fold = DENSE_SEQ_FOLD,
                           bool testSet = false)
{
    const size_t smem = dense_smem_bytes(winBits, THREADS, sizeof(T)) + (size_t)SVALS * sizeof(T);
    auto kern = k_dense_rows<THREADS, T, NUMERIC, SVALS>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int perSm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, THREADS, smem);
    if (perSm < 1) perSm = 1;
    u32 grid = (u32)(lc.smCount * perSm);
    if (grid > count) grid = count;
    kern<<<grid, THREADS, smem, lc.stream>>>(perm, count, rowCounter, aRp, aCi, aV, bRp, bCi, bV, winBits, rowMin,
                                             rowMax, bitmapStore, cRp, cCi, cV, seqMax, rowOps, seqRows, fold, testSet);
    ++*lc.launches;
}

// Two launch shapes:
//   local : rows whose column extent fits DENSE_LOCAL_COLS -> 16 k-column window (2 KB bitmap), 256 threads,
//           2048-entry shared value accumulator, ~8 CTAs per SM
//   wide  : everything else -> window of up to 2^20 columns, 1024 threads (one or two CTAs per SM),
//           4096-entry shared value accumulator
size_t dense_local_store_bytes(u32 count) { return (size_t)count * ((size_t)1 << (DENSE_LOCAL_BITS - 5)) * sizeof(u32); }

void launch_dense_symbolic(const LaunchCtx &lc, bool local, const u32 *perm, u32 count, u32 *rowCounter,
                           const u32 *aRp, const u32 *aCi, const u32 *bRp, const u32 *bCi, u32 colsB,
                           const u32 *rowMin, const u32 *rowMax, u32 *bitmapStore, u32 *rowNnz, const u32 *rowOps,
                           u32 *seqRows, bool everyRow, bool testSet)
{
    if (count == 0) return;
    const float *nv = nullptr;
    if (local)
        launch_dense_t<256, float, false, 0>(lc, DENSE_LOCAL_BITS, perm, count, rowCounter, aRp, aCi, nv, bRp, bCi, nv,
                                             rowMin, rowMax, bitmapStore, rowNnz, nullptr, nullptr, 0u,
                                             rowOps, (bitmapStore && rowOps) ? seqRows : nullptr,
                                             everyRow ? 0u : DENSE_SEQ_FOLD, testSet);
    else
        launch_dense_t<1024, float, false, 0>(lc, dense_window_bits(colsB), perm, count, rowCounter, aRp, aCi, nv, bRp,
                                              bCi, nv, rowMin, rowMax, nullptr, rowNnz, nullptr, nullptr, 0u, rowOps);
}

template <typename T>
void launch_dense_numeric(const LaunchCtx &lc, bool local, const u32 *perm, u32 count, u32 *rowCounter,
                          const u32 *aRp, const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi, const T *bV,
                          u32 colsB, const u32 *rowMin, const u32 *rowMax, u32 *bitmapStore, const u32 *cRp,
                          u32 *cCi, T *cV, int seq, const u32 *rowOps)
{
    const u32 fold = (seq & 16) ? 0u : DENSE_SEQ_FOLD;   // deterministic mode: every row that fits the accumulator
    if (count == 0) return;
    u32 *rp = const_cast<u32 *>(cRp);
    // sequential-k kernel with TMA-staged B segments (dense_seq.cuh): local rows with a kept bitmap whose distinct
    // columns fit its shared accumulator; bulk copies need 16-byte aligned B arrays
    const bool tma = (seq & 3) == 2 && ((reinterpret_cast<uintptr_t>(bCi) | reinterpret_cast<uintptr_t>(bV)) & 15u) == 0;
    const bool small = seq & 4, large = seq & 8;   // which shapes have rows (counted by the symbolic kernel)
    const bool useSeq = (seq & 3) && (small || large) && local && bitmapStore && rowOps;
    if (useSeq && tma) {
        if (small) launch_dense_seq_t<T, 512, true>(lc, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowMin, rowMax, bitmapStore, cRp, cCi, cV, 0u, rowOps, fold);
        if (large) launch_dense_seq_t<T, DENSE_SEQ_MAX, true>(lc, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowMin, rowMax, bitmapStore, cRp, cCi, cV, 512u, rowOps, fold);
    } else if (useSeq) {
        if (small) launch_dense_seq_t<T, 512, false>(lc, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowMin, rowMax, bitmapStore, cRp, cCi, cV, 0u, rowOps, fold);
        if (large) launch_dense_seq_t<T, DENSE_SEQ_MAX, false>(lc, perm, count, aRp, aCi, aV, bRp, bCi, bV, rowMin, rowMax, bitmapStore, cRp, cCi, cV, 512u, rowOps, fold);
    }
    if (local)
        launch_dense_t<256, T, true, 2048>(lc, DENSE_LOCAL_BITS, perm, count, rowCounter, aRp, aCi, aV, bRp, bCi, bV,
                                           rowMin, rowMax, bitmapStore, rp, cCi, cV, useSeq ? (u32)DENSE_SEQ_MAX : 0u, rowOps, nullptr, fold);
    else
        launch_dense_t<1024, T, true, 4096>(lc, dense_window_bits(colsB), perm, count, rowCounter, aRp, aCi, aV, bRp,
                                            bCi, bV, rowMin, rowMax, nullptr, rp, cCi, cV);
}
template <typename T>
void launch_det_rows(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi, const T *aV,
                     const u32 *bRp, const u32 *bCi, const T *bV, const u32 *cRp, const u32 *cCi, T *cV,
                     const u32 *rowOps, bool skipSeqRows)
{
    launch_det_rows_t<T>(lc, perm, count, aRp, aCi, aV, bRp, bCi, bV, cRp, cCi, cV, rowOps,
                         skipSeqRows ? (u32)DENSE_SEQ_MAX : 0u, 0u);
}
template void launch_det_rows<double>(const LaunchCtx &, const u32 *, u32, const u32 *, const u32 *, const double *,
                                      const u32 *, const u32 *, const double *, const u32 *, const u32 *, double *,
                                      const u32 *, bool);
template void launch_det_rows<float>(const LaunchCtx &, const u32 *, u32, const u32 *, const u32 *, const float *,
                                     const u32 *, const u32 *, const float *, const u32 *, const u32 *, float *,
                                     const u32 *, bool);

template void launch_dense_numeric<double>(const LaunchCtx &, bool, const u32 *, u32, u32 *, const u32 *, const u32 *,
                                           const double *, const u32 *, const u32 *, const double *, u32,
                                           const u32 *, const u32 *, u32 *, const u32 *, u32 *, double *, int, const u32 *);
template void launch_dense_numeric<float>(const LaunchCtx &, bool, const u32 *, u32, u32 *, const u32 *, const u32 *,
                                          const float *, const u32 *, const u32 *, const float *, u32,
                                          const u32 *, const u32 *, u32 *, const u32 *, u32 *, float *, int, const u32 *);

}  // namespace sb
