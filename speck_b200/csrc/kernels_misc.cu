// speck_b200/csrc/kernels_misc.cu -- row analysis, binning, row_ptr scan, direct rows, compare.
#include "common.cuh"

namespace sb {

// ------------------------------------------------------------------------------------------
// Row analysis.  Replaces readOperations (reference include/common.cuh:321-459):
//   rowOps[i] = sum_{k in A_i} nnz(B_k)   (upper bound of nnz(C_i), exact product count)
//   P (u64 atomic), max over rows, and -- new -- the bin histogram, so that no host
//   round trip is needed to classify (reference: Multiply.cu:263-345 on the host).
// LA lanes cooperate on one row; A.col_idx is read coalesced across the lane group.
// Rows without products get rowNnz = 0 here; one-entry rows get rowNnz = nnz(B_k) (they
// skip the symbolic phase, as directSpGEMMCount does, spECK_HashSpGEMM.cuh:572-589).
// ------------------------------------------------------------------------------------------
// Per-row summary of B: (begin, end, first column, last column) in one 16-byte entry, so that the analysis
// gathers ONE sector per A entry instead of three (B.row_offsets pair, first and last column of the row).
__global__ void __launch_bounds__(256) k_row_info(u32 rowsB, const u32 *__restrict__ bRp, const u32 *__restrict__ bCi,
                                                  uint4 *__restrict__ rowInfo)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= rowsB) return;
    const u32 bs = bRp[k], be = bRp[k + 1];
    uint4 ri = make_uint4(bs, be, 0xffffffffu, 0u);
    if (be > bs) {
        ri.z = __ldg(bCi + bs);
        ri.w = __ldg(bCi + be - 1);
    }
    rowInfo[k] = ri;
}

void launch_row_info(const LaunchCtx &lc, u32 rowsB, const u32 *bRp, const u32 *bCi, uint4 *rowInfo)
{
    if (rowsB == 0) return;
    k_row_info<<<(rowsB + 255) / 256, 256, 0, lc.stream>>>(rowsB, bRp, bCi, rowInfo);
    ++*lc.launches;
}

// B-row summary of one A entry: (begin, end, first column, last column); empty rows carry (0xffffffff, 0)
__device__ __forceinline__ uint4 fetch_row_summary(const uint4 *__restrict__ rowInfo, const u32 *__restrict__ bRp,
                                                   const u32 *__restrict__ bCi, u32 k, bool tiered)
{
    if (rowInfo) return __ldg(rowInfo + k);
    uint4 ri = make_uint4(__ldg(bRp + k), __ldg(bRp + k + 1), 0xffffffffu, 0u);
    if (!tiered && ri.y > ri.x) {
        ri.z = __ldg(bCi + ri.x);
        ri.w = __ldg(bCi + ri.y - 1);
    }
    return ri;
}

constexpr u32 ANALYZE_LONG_ROW = 1024;   // A rows with at least this many entries: one CTA per row (k_analyze_long)

template <int LA>
__global__ void __launch_bounds__(256) k_analyze(u32 rows, const u32 *__restrict__ aRp,
                                                 const u32 *__restrict__ aCi,
                                                 const u32 *__restrict__ bRp, const u32 *__restrict__ bCi,
                                                 u32 *__restrict__ rowOps, u32 *__restrict__ rowMin,
                                                 u32 *__restrict__ rowMax, u32 *__restrict__ rowNnz, Scalars *sc,
                                                 u32 sortMax, uint2 *__restrict__ aSeg,
                                                 const uint4 *__restrict__ rowInfo, u32 *__restrict__ aOff,
                                                 u32 *__restrict__ mapLen, bool mapCta, int mapMinClass,
                                                 u32 extentMinOps, u32 *__restrict__ longRows)
{
    __shared__ u32 sBin[NUM_BINS];
    __shared__ unsigned long long sProd;
    __shared__ u32 sMax;
    if (threadIdx.x < NUM_BINS) sBin[threadIdx.x] = 0;
    if (threadIdx.x == 0) { sProd = 0; sMax = 0; }
    __syncthreads();

    const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 row = gtid / LA;
    const u32 lane = threadIdx.x % LA;
    u64 ops64 = 0;
    u32 aLen = 0;
    u32 cmin = 0xffffffffu, cmax = 0u;  // column extent of the row's products: B rows are sorted, so the
                                        // first / last entry of each referenced B row bound it (the
                                        // reference's rowColMinMax, common.cuh:395-400)
    u32 beg = row < rows ? aRp[row] : 0u, end = row < rows ? aRp[row + 1] : 0u;
    aLen = end - beg;
    // hub rows of A (web graphs: thousands of entries) would serialise hundreds of dependent gathers on one lane
    // group: they are queued for k_analyze_long (one CTA per row) and skipped here
    const bool queued = longRows && aLen >= ANALYZE_LONG_ROW;
    if (queued) {
        if (lane == 0) longRows[atomicAdd(&sc->longCount, 1u)] = row;
        end = beg;
    }
    // B-row summary of one A entry: (begin, end, first column, last column); empty rows carry (0xffffffff, 0)
    // extentMinOps > 0 (no rowInfo table: B has so many rows that the 16-byte summaries would miss the L2): the first
    // pass gathers only the row_offsets pair (4 B per row of B, L2-resident); the first / last columns are gathered in a
    // second pass for the rows with at least extentMinOps products -- the only ones whose extent is ever used
    // (banded test of classify_row, rank and bitmap kernels); the others keep the placeholder extent [0, 0].
    const bool tiered = !rowInfo && extentMinOps > 0;
    auto fetch = [&](u32 k) -> uint4 { return fetch_row_summary(rowInfo, bRp, bCi, k, tiered); };
    const u32 gmask = LA == 32 ? 0xffffffffu : (((1u << (LA & 31)) - 1u) << ((threadIdx.x & 31) - lane));
    if (!aOff) {
        // four entries per lane and iteration: their summary gathers are in flight together (a hub row of A
        // otherwise serialises hundreds of dependent column -> summary loads on one lane group)
        for (u32 p0 = beg + lane; p0 < end; p0 += 4 * LA) {
            u32 kk[4];
            uint4 ri[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) kk[u] = (p0 + u * LA < end) ? __ldg(aCi + p0 + u * LA) : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < 4; ++u) ri[u] = kk[u] != 0xffffffffu ? fetch(kk[u]) : make_uint4(0u, 0u, 0xffffffffu, 0u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (kk[u] == 0xffffffffu) continue;
                cmin = min(cmin, ri[u].z);
                cmax = max(cmax, ri[u].w);
                ops64 += (u64)(ri[u].y - ri[u].x);
                if (aSeg) aSeg[p0 + u * LA] = make_uint2(ri[u].x, ri[u].y);
            }
        }
    } else {
        // all LA lanes of the group run the same number of iterations (the per-entry product offsets need a group scan)
        for (u32 p0 = beg; p0 < end; p0 += LA) {
            const u32 p = p0 + lane;
            u32 len = 0, bs = 0, be = 0;
            if (p < end) {
                const uint4 ri = fetch(__ldg(aCi + p));
                bs = ri.x; be = ri.y;
                cmin = min(cmin, ri.z);
                cmax = max(cmax, ri.w);
                len = be - bs;
            }
            // index of the entry's first product in the row's flat enumeration (u32: rows beyond 2^32 products are not mapped)
            u32 incl = len;
#pragma unroll
            for (int d = 1; d < LA; d <<= 1) {
                const u32 t = __shfl_up_sync(gmask, incl, d, LA);
                if ((int)lane >= d) incl += t;
            }
            if (p < end) aOff[p] = (u32)ops64 + incl - len;
            ops64 += (u64)__shfl_sync(gmask, incl, LA - 1, LA);   // every lane carries the running row total
            if (aSeg && p < end) aSeg[p] = make_uint2(bs, be);
        }
    }
    if (!aOff) {
#pragma unroll
        for (int d = LA / 2; d >= 1; d >>= 1) ops64 += __shfl_xor_sync(0xffffffffu, ops64, d);
    }
    if (tiered) {
        cmin = 0xffffffffu; cmax = 0u;
        if (ops64 >= (u64)extentMinOps) {   // uniform in the lane group: every lane holds the row total
            for (u32 p0 = beg + lane; p0 < end; p0 += 2 * LA) {
                u32 bs[2], be[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    bs[u] = be[u] = 0;
                    if (p0 + u * LA < end) {
                        const u32 k = __ldg(aCi + p0 + u * LA);
                        bs[u] = __ldg(bRp + k);
                        be[u] = __ldg(bRp + k + 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (be[u] > bs[u]) {
                        cmin = min(cmin, __ldg(bCi + bs[u]));
                        cmax = max(cmax, __ldg(bCi + be[u] - 1));
                    }
            }
        } else {
            cmin = 0u;
        }
    }
#pragma unroll
    for (int d = LA / 2; d >= 1; d >>= 1) {
        cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, d));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
    }
    const u32 ops = ops64 > 0xffffffffull ? 0xffffffffu : (u32)ops64;

    u64 myProd = 0;
    u32 myMax = 0;
    if (row < rows && lane == 0 && !queued) {
        rowOps[row] = ops;
        rowMin[row] = cmin;
        rowMax[row] = cmax;
        const int bin = classify_row(ops, aLen, ops ? cmax - cmin + 1u : 0u, sortMax);
        if (mapLen) {   // products of the row when its class records a rank map in the symbolic phase
            const bool mapped = bin >= BIN_SORT0 + mapMinClass && (bin < BIN_SORT0 + NUM_WARP_SORT || (mapCta && bin < BIN_DENSE_LOCAL));
            mapLen[row] = mapped ? ops : 0u;
        }
        if (bin < 0)
            rowNnz[row] = 0;
        else {
            if (bin == BIN_DIRECT) rowNnz[row] = ops;
            atomicAdd(&sBin[bin], 1u);
        }
        myProd = ops64;
        myMax = ops;
    }
    // warp reduce, then one shared atomic per warp, one global atomic per block
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        myProd += __shfl_xor_sync(0xffffffffu, myProd, d);
        myMax = max(myMax, __shfl_xor_sync(0xffffffffu, myMax, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (myProd) atomicAdd(&sProd, (unsigned long long)myProd);
        if (myMax) atomicMax(&sMax, myMax);
    }
    __syncthreads();
    if (threadIdx.x < NUM_BINS && sBin[threadIdx.x]) atomicAdd(&sc->binCount[threadIdx.x], sBin[threadIdx.x]);
    if (threadIdx.x == 0) {
        if (sProd) atomicAdd((unsigned long long *)&sc->products, sProd);
        if (sMax) atomicMax(&sc->maxRowProducts, sMax);
    }
}

// One CTA per queued hub row of A: same outputs as k_analyze, 1024 entries per iteration.
__global__ void __launch_bounds__(256) k_analyze_long(const u32 *__restrict__ longRows, const u32 *__restrict__ aRp,
                                                      const u32 *__restrict__ aCi, const u32 *__restrict__ bRp,
                                                      const u32 *__restrict__ bCi, u32 *__restrict__ rowOps,
                                                      u32 *__restrict__ rowMin, u32 *__restrict__ rowMax,
                                                      u32 *__restrict__ rowNnz, Scalars *sc, u32 sortMax,
                                                      uint2 *__restrict__ aSeg, const uint4 *__restrict__ rowInfo,
                                                      u32 *__restrict__ mapLen, bool mapCta, int mapMinClass, u32 extentMinOps)
{
    __shared__ unsigned long long sOps[8];
    __shared__ u32 sMin[8], sMax[8];
    const u32 n = sc->longCount;
    for (u32 i = blockIdx.x; i < n; i += gridDim.x) {
        const u32 row = longRows[i];
        const u32 beg = aRp[row], end = aRp[row + 1];
        u64 ops64 = 0;
        u32 cmin = 0xffffffffu, cmax = 0u;
        for (u32 p0 = beg + threadIdx.x; p0 < end; p0 += 4 * 256) {
            u32 kk[4];
            uint4 ri[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) kk[u] = (p0 + u * 256 < end) ? __ldg(aCi + p0 + u * 256) : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ri[u] = make_uint4(0u, 0u, 0xffffffffu, 0u);
                if (kk[u] != 0xffffffffu) {
                    ri[u] = fetch_row_summary(rowInfo, bRp, bCi, kk[u], false);   // a hub row always needs its extent
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (kk[u] == 0xffffffffu) continue;
                cmin = min(cmin, ri[u].z);
                cmax = max(cmax, ri[u].w);
                ops64 += (u64)(ri[u].y - ri[u].x);
                if (aSeg) aSeg[p0 + u * 256] = make_uint2(ri[u].x, ri[u].y);
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            ops64 += __shfl_xor_sync(0xffffffffu, ops64, d);
            cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, d));
            cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
        }
        __syncthreads();   // the arrays may still be read by thread 0 of the previous row
        if ((threadIdx.x & 31) == 0) {
            sOps[threadIdx.x >> 5] = ops64;
            sMin[threadIdx.x >> 5] = cmin;
            sMax[threadIdx.x >> 5] = cmax;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) {
                ops64 += sOps[w];
                cmin = min(cmin, sMin[w]);
                cmax = max(cmax, sMax[w]);
            }
            const u32 ops = ops64 > 0xffffffffull ? 0xffffffffu : (u32)ops64;
            rowOps[row] = ops;
            rowMin[row] = cmin;
            rowMax[row] = cmax;
            const int bin = classify_row(ops, end - beg, ops ? cmax - cmin + 1u : 0u, sortMax);
            if (mapLen) {
                const bool mapped = bin >= BIN_SORT0 + mapMinClass && (bin < BIN_SORT0 + NUM_WARP_SORT || (mapCta && bin < BIN_DENSE_LOCAL));
                mapLen[row] = mapped ? ops : 0u;
            }
            if (bin < 0)
                rowNnz[row] = 0;
            else
                atomicAdd(&sc->binCount[bin], 1u);
            if (ops64) atomicAdd((unsigned long long *)&sc->products, (unsigned long long)ops64);
            if (ops) atomicMax(&sc->maxRowProducts, ops);
        }
    }
}

void launch_analyze(const LaunchCtx &lc, u32 rows, u64 nnzA, const u32 *aRp, const u32 *aCi, const u32 *bRp,
                    const u32 *bCi, u32 *rowOps, u32 *rowMin, u32 *rowMax, u32 *rowNnz, Scalars *sc, u32 sortMax,
                    uint2 *aSeg, const uint4 *rowInfo, u32 *aOff, u32 *mapLen, bool mapCta, int mapMinClass,
                    u32 extentMinOps, u32 *longRows)
{
    if (rows == 0) return;
    const double avg = (double)nnzA / (double)rows;
    const int threads = 256;
#define SB_ANALYZE(LA)                                                                                                     \
    k_analyze<LA><<<(u32)(((u64)rows * LA + threads - 1) / threads), threads, 0, lc.stream>>>(                             \
        rows, aRp, aCi, bRp, bCi, rowOps, rowMin, rowMax, rowNnz, sc, sortMax, aSeg, rowInfo, aOff, mapLen, mapCta, mapMinClass,  \
        extentMinOps, aOff ? nullptr : longRows)
    if (avg <= 3.0) SB_ANALYZE(2);
    else if (avg <= 6.0) SB_ANALYZE(4);
    else if (avg <= 24.0) SB_ANALYZE(8);
    else SB_ANALYZE(32);
#undef SB_ANALYZE
    ++*lc.launches;
    if (longRows && !aOff) {   // the queue is usually empty: the CTAs read one counter and leave
        k_analyze_long<<<lc.smCount, 256, 0, lc.stream>>>(longRows, aRp, aCi, bRp, bCi, rowOps, rowMin, rowMax, rowNnz, sc, sortMax,
                                                          aSeg, rowInfo, mapLen, mapCta, mapMinClass, extentMinOps);
        ++*lc.launches;
    }
}

// ------------------------------------------------------------------------------------------
// Binning: scatter row ids into one permutation array ordered by bin.  Replaces the load
// balancer (reference spECK_HashLoadBalancer.cuh:265-347 + scan_largearray_kernel.cuh:182-281
// + the <=6 D2D memcpys): one pass, one global atomic per (block, bin).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bin_scatter(u32 rows, const u32 *__restrict__ aRp,
                                                     const u32 *__restrict__ rowOps,
                                                     const u32 *__restrict__ rowMin,
                                                     const u32 *__restrict__ rowMax, u32 *__restrict__ perm,
                                                     Scalars *sc, u32 sortMax, const u64 *__restrict__ mapBase,
                                                     RowDesc *__restrict__ desc)
{
    __shared__ u32 sCnt[NUM_BINS];
    __shared__ u32 sBase[NUM_BINS];
    if (threadIdx.x < NUM_BINS) sCnt[threadIdx.x] = 0;
    __syncthreads();
    const u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    int bin = -1;
    u32 rank = 0;
    if (row < rows) {
        const u32 ops = rowOps[row];
        bin = classify_row(ops, aRp[row + 1] - aRp[row], ops ? rowMax[row] - rowMin[row] + 1u : 0u, sortMax);
        if (bin >= 0) rank = atomicAdd(&sCnt[bin], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NUM_BINS) {
        u32 start = 0;
        for (int b = 0; b < (int)threadIdx.x; ++b) start += sc->binCount[b];
        const u32 c = sCnt[threadIdx.x];
        sBase[threadIdx.x] = start + (c ? atomicAdd(&sc->binCursor[threadIdx.x], c) : 0u);
    }
    __syncthreads();
    if (bin >= 0) {
        const u32 pos = sBase[bin] + rank;
        perm[pos] = row;
        if (desc) {   // row descriptor in bin order (common.cuh: RowDesc): every source is read coalesced by row here
            RowDesc d;
            d.aBeg = aRp[row];
            d.aLen = aRp[row + 1] - d.aBeg;
            d.n = rowOps[row];
            d.row = row;
            d.c0 = rowMin[row];
            d.c1 = rowMax[row];
            d.mapOff = mapBase[row];
            desc[pos] = d;
        }
    }
}

void launch_bin_scatter(const LaunchCtx &lc, u32 rows, const u32 *aRp, const u32 *rowOps, const u32 *rowMin,
                        const u32 *rowMax, u32 *perm, Scalars *sc, u32 sortMax, const u64 *mapBase, RowDesc *desc)
{
    if (rows == 0) return;
    k_bin_scatter<<<(rows + 255) / 256, 256, 0, lc.stream>>>(rows, aRp, rowOps, rowMin, rowMax, perm, sc, sortMax, mapBase, desc);
    ++*lc.launches;
}

// ------------------------------------------------------------------------------------------
// Row descriptors of the mapped classes (common.cuh: RowDesc), in perm order: written by k_bin_scatter, switched to
// the numeric flavour (position / length in C) here once row_offsets are scanned.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_desc_numeric(u32 count, const u32 *__restrict__ cRp, RowDesc *desc)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const u32 row = desc[j].row;
    const u32 b = cRp[row];
    desc[j].c0 = b;
    desc[j].c1 = cRp[row + 1] - b;
}

void launch_desc_numeric(const LaunchCtx &lc, u32 count, const u32 *cRp, RowDesc *desc)
{
    if (count == 0) return;
    k_desc_numeric<<<(count + 255) / 256, 256, 0, lc.stream>>>(count, cRp, desc);
    ++*lc.launches;
}

// ------------------------------------------------------------------------------------------
// row_ptr scan: single-pass exclusive scan with decoupled look-back, in place.
// Replaces cub::DeviceScan::ExclusiveSum (reference Multiply.cu:570).  data[n-1] is the
// trailing slot (its input is ignored and treated as 0, so it receives the total = nnz(C)).
// Tile state word: bits 63..62 = flag (0 empty, 1 aggregate, 2 inclusive prefix), rest value.
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr u64 FLAG_AGG = 1ull << 62;
constexpr u64 FLAG_PFX = 2ull << 62;
constexpr u64 VALUE_MASK = (1ull << 62) - 1;

size_t scan_tile_state_bytes(u32 n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE + 1) * sizeof(u64); }

template <typename OutT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(const u32 *data, OutT *out, u32 n, volatile u64 *tileState,
                                                       u32 *tileCounter, u64 *totalOut)
{
    __shared__ u32 sTile;
    __shared__ u64 sWarpSum[SCAN_THREADS / 32];
    __shared__ u64 sExclusive;
    if (threadIdx.x == 0) sTile = atomicAdd(tileCounter, 1u);
    __syncthreads();
    const u32 tile = sTile;
    const u32 base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;

    u32 items[SCAN_ITEMS];
    u64 tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const u32 idx = base + i;
        items[i] = (idx < n - 1) ? data[idx] : 0u;  // trailing slot and out-of-range count as 0
        tsum += items[i];
    }
    // block inclusive scan of thread sums
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) sWarpSum[warp] = incl;
    __syncthreads();
    u64 warpBase = 0, blockAgg = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < (int)warp) warpBase += sWarpSum[w];
        blockAgg += sWarpSum[w];
    }
    const u64 threadExcl = warpBase + incl - tsum;

    // publish aggregate, look back (warp 0)
    if (warp == 0) {
        if (lane == 0) {
            tileState[tile] = (tile == 0 ? FLAG_PFX : FLAG_AGG) | blockAgg;
            __threadfence();
        }
        u64 exclusive = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            while (true) {
                const int idx = t - (int)lane;
                u64 st = FLAG_PFX;  // lanes before tile 0 read as "prefix 0"
                if (idx >= 0) {
                    do { st = tileState[idx]; } while ((st >> 62) == 0);
                }
                const u32 pfxMask = __ballot_sync(0xffffffffu, (st >> 62) == 2);
                // nearest tile with an inclusive prefix = lowest lane set in pfxMask
                const int first = __ffs(pfxMask) - 1;  // always >= 0 eventually (tile 0 or virtual)
                u64 contrib = ((int)lane <= first || first < 0) ? (st & VALUE_MASK) : 0;
                if (idx < 0) contrib = 0;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                exclusive += contrib;
                if (first >= 0) break;
                t -= 32;
            }
            if (lane == 0) {
                tileState[tile] = FLAG_PFX | (exclusive + blockAgg);
                __threadfence();
            }
        }
        if (lane == 0) sExclusive = exclusive;
    }
    __syncthreads();
    u64 run = sExclusive + threadExcl;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const u32 idx = base + i;
        if (idx < n) out[idx] = (OutT)run;
        if (idx == n - 1) *totalOut = run;  // 64-bit total: overflow of the u32 API is detectable
        run += items[i];
    }
}

void launch_scan(const LaunchCtx &lc, u32 *data, u32 n, u64 *tileState, Scalars *sc)
{
    if (n == 0) return;
    const u32 tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    cudaMemsetAsync(tileState, 0, (size_t)tiles * sizeof(u64), lc.stream);
    k_scan<u32><<<tiles, SCAN_THREADS, 0, lc.stream>>>(data, data, n, tileState, &sc->tileCounter, &sc->nnzC);
    ++*lc.launches;
}

void launch_scan_map(const LaunchCtx &lc, const u32 *in, u64 *out, u32 n, u64 *tileState, Scalars *sc)
{
    if (n == 0) return;
    const u32 tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    cudaMemsetAsync(tileState, 0, (size_t)tiles * sizeof(u64), lc.stream);
    k_scan<u64><<<tiles, SCAN_THREADS, 0, lc.stream>>>(in, out, n, tileState, &sc->mapTileCounter, &sc->mapTotal);
    ++*lc.launches;
}

// ------------------------------------------------------------------------------------------
// Scalars -> mapped pinned host memory, then a sequence number: the host polls the sequence word instead of going
// through cudaStreamSynchronize (the two mid-pipeline read-backs of one multiply: bin counts, nnz(C)).  Replaces
// the reference's blocking 4/8-byte cudaMemcpy's (source/GPU/Multiply.cu:250, 573, 615).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_publish(const Scalars *__restrict__ dSc, Scalars *hSc, volatile u32 *hSeq, u32 seq)
{
    const u32 *src = reinterpret_cast<const u32 *>(dSc);
    volatile u32 *dst = reinterpret_cast<volatile u32 *>(hSc);
    for (u32 i = threadIdx.x; i < sizeof(Scalars) / 4; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *hSeq = seq;
}

void launch_publish(const LaunchCtx &lc, const Scalars *dSc, Scalars *hSc, volatile u32 *hSeq, u32 seq)
{
    k_publish<<<1, 128, 0, lc.stream>>>(dSc, hSc, hSeq, seq);
    ++*lc.launches;
}

// ------------------------------------------------------------------------------------------
// Multi-GPU helpers (no reference counterpart: the reference is single-device).
//   k_find_cuts:   contiguous row cuts balanced by products: cut g = first row whose exclusive product prefix
//                  reaches g * P / parts (prefix = the u64 scan of rowOps), cuts[0] = 0, cuts[parts] = rows
//   k_offset_rows: row_offsets of slab g of the concatenated C: out[i] = in[i] + base
// ------------------------------------------------------------------------------------------
__global__ void k_find_cuts(const u64 *__restrict__ prefix, u32 rows, u32 parts, u32 *__restrict__ cuts,
                            u64 *__restrict__ partProducts)
{
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > parts) return;
    const u64 total = prefix[rows];
    auto cut_of = [&](u32 part) -> u32 {
        if (part == 0) return 0u;
        if (part >= parts) return rows;
        const u64 target = (u64)(((unsigned __int128)total * part) / parts);
        u32 lo = 0, hi = rows;   // first r with prefix[r] >= target
        while (lo < hi) {
            const u32 mid = lo + (hi - lo) / 2;
            if (prefix[mid] < target) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const u32 c = cut_of(g);
    cuts[g] = c;
    if (g < parts && partProducts) partProducts[g] = prefix[cut_of(g + 1)] - prefix[c];
}

// cost of a row for the partition = products + wEntry * entries of the A row + wRow (saturating u32)
__global__ void __launch_bounds__(256) k_row_cost(u32 rows, const u32 *__restrict__ aRp, u32 *__restrict__ rowOps, u32 wRow, u32 wEntry)
{
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const u64 c = (u64)rowOps[r] + (u64)wEntry * (aRp[r + 1] - aRp[r]) + wRow;
    rowOps[r] = c > 0xffffffffull ? 0xffffffffu : (u32)c;
}

void launch_row_cost(const LaunchCtx &lc, u32 rows, const u32 *aRp, u32 *rowOps, u32 wRow, u32 wEntry)
{
    if (rows == 0 || (wRow == 0 && wEntry == 0)) return;
    k_row_cost<<<(rows + 255) / 256, 256, 0, lc.stream>>>(rows, aRp, rowOps, wRow, wEntry);
    ++*lc.launches;
}

void launch_find_cuts(const LaunchCtx &lc, const u64 *prefix, u32 rows, u32 parts, u32 *cuts, u64 *partProducts)
{
    k_find_cuts<<<(parts + 1 + 63) / 64, 64, 0, lc.stream>>>(prefix, rows, parts, cuts, partProducts);
    ++*lc.launches;
}

__global__ void __launch_bounds__(256) k_offset_rows(const u32 *__restrict__ in, u32 n, u32 base, u32 *__restrict__ out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + base;
}

void launch_offset_rows(const LaunchCtx &lc, const u32 *in, u32 n, u32 base, u32 *out)
{
    if (n == 0) return;
    k_offset_rows<<<(n + 255) / 256, 256, 0, lc.stream>>>(in, n, base, out);
    ++*lc.launches;
}

// ------------------------------------------------------------------------------------------
// Push of one slab of C into the concatenated C (one process per GPU: the destination is peer memory opened through
// CUDA IPC, the stores travel over NVLink / NVSwitch; the gathering device pushes its own slab with the same kernel).
// The transfer and the row_offsets fix-up are one kernel: every thread moves 16-byte pieces (destination-aligned, the
// source is read element-wise when the two are not aligned alike), row offsets get the slab's nnz offset added on
// the way.  Stores, not loads, cross the link: a store needs no round trip.
// ------------------------------------------------------------------------------------------
template <typename E>
__device__ __forceinline__ void push_elems(E *__restrict__ dst, const E *__restrict__ src, u64 n, u64 gtid, u64 gsize)
{
    constexpr u32 V = 16 / sizeof(E);
    u64 head = ((16u - (u32)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) / sizeof(E);
    if (head > n) head = n;
    for (u64 i = gtid; i < head; i += gsize) dst[i] = src[i];
    const u64 body = (n - head) / V;
    const E *s0 = src + head;
    uint4 *d0 = reinterpret_cast<uint4 *>(dst + head);
    if ((reinterpret_cast<uintptr_t>(s0) & 15u) == 0) {
        const uint4 *sv = reinterpret_cast<const uint4 *>(s0);
        u64 j = gtid;
        for (; j + 3 * gsize < body; j += 4 * gsize) {   // four 16-byte pieces in flight per thread
            const uint4 a = sv[j], b = sv[j + gsize], c = sv[j + 2 * gsize], d = sv[j + 3 * gsize];
            d0[j] = a; d0[j + gsize] = b; d0[j + 2 * gsize] = c; d0[j + 3 * gsize] = d;
        }
        for (; j < body; j += gsize) d0[j] = sv[j];
    } else {
        for (u64 j = gtid; j < body; j += gsize) {
            union { uint4 v; E e[V]; } t;
#pragma unroll
            for (u32 k = 0; k < V; ++k) t.e[k] = s0[j * V + k];
            d0[j] = t.v;
        }
    }
    for (u64 i = head + body * V + gtid; i < n; i += gsize) dst[i] = src[i];
}

template <typename T>
__global__ void __launch_bounds__(256) k_push_slab(const u32 *__restrict__ srcRp, const u32 *__restrict__ srcCi,
                                                   const T *__restrict__ srcV, u32 rowsOut, u64 nnz, u32 nnzBase,
                                                   u32 *__restrict__ dstRp, u32 *__restrict__ dstCi, T *__restrict__ dstV)
{
    const u64 gtid = (u64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (u64)gridDim.x * blockDim.x;
    for (u64 i = gtid; i < rowsOut; i += gsize) dstRp[i] = srcRp[i] + nnzBase;
    push_elems<u32>(dstCi, srcCi, nnz, gtid, gsize);
    push_elems<T>(dstV, srcV, nnz, gtid, gsize);
}

template <typename T>
void launch_push_slab(const LaunchCtx &lc, const u32 *srcRp, const u32 *srcCi, const T *srcV, u32 rowsOut, u64 nnz,
                      u64 nnzBase, u32 rowBase, u32 *dstRp, u32 *dstCi, T *dstV)
{
    if (rowsOut == 0 && nnz == 0) return;
    k_push_slab<T><<<lc.smCount * 8, 256, 0, lc.stream>>>(srcRp, srcCi, srcV, rowsOut, nnz, (u32)nnzBase, dstRp + rowBase,
                                                          dstCi + nnzBase, dstV + nnzBase);
    ++*lc.launches;
}
template void launch_push_slab<double>(const LaunchCtx &, const u32 *, const u32 *, const double *, u32, u64, u64, u32, u32 *,
                                       u32 *, double *);
template void launch_push_slab<float>(const LaunchCtx &, const u32 *, const u32 *, const float *, u32, u64, u64, u32, u32 *,
                                      u32 *, float *);

// ------------------------------------------------------------------------------------------
// Direct rows: A row with one entry -> C row = a_ik * B_k, already sorted.
// (reference: directSpGEMMNumericImplementation, spECK_HashSpGEMM.cuh:543-569)
// ------------------------------------------------------------------------------------------
template <typename T, int LPR>
__global__ void __launch_bounds__(256) k_direct(const u32 *__restrict__ perm, u32 count,
                                                const u32 *__restrict__ aRp, const u32 *__restrict__ aCi,
                                                const T *__restrict__ aV, const u32 *__restrict__ bRp,
                                                const u32 *__restrict__ bCi, const T *__restrict__ bV,
                                                const u32 *__restrict__ cRp, u32 *__restrict__ cCi,
                                                T *__restrict__ cV)
{
    const u32 g = (blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const u32 l = threadIdx.x % LPR;
    if (g >= count) return;
    const u32 row = perm[g];
    const u32 a = aRp[row];
    const u32 k = aCi[a];
    const T av = aV[a];
    const u32 bs = bRp[k], len = bRp[k + 1] - bs;
    const u32 o = cRp[row];
    for (u32 j = l; j < len; j += LPR) {
        cCi[o + j] = __ldg(bCi + bs + j);
        cV[o + j] = av * __ldg(bV + bs + j);
    }
}

template <typename T>
void launch_direct_numeric(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi,
                           const T *aV, const u32 *bRp, const u32 *bCi, const T *bV, const u32 *cRp,
                           u32 *cCi, T *cV)
{
    if (count == 0) return;
    constexpr int LPR = 8;
    const u32 grid = (u32)(((u64)count * LPR + 255) / 256);
    k_direct<T, LPR><<<grid, 256, 0, lc.stream>>>(perm, count, aRp, aCi, aV, bRp, bCi, bV, cRp, cCi, cV);
    ++*lc.launches;
}
template void launch_direct_numeric<double>(const LaunchCtx &, const u32 *, u32, const u32 *, const u32 *,
                                            const double *, const u32 *, const u32 *, const double *,
                                            const u32 *, u32 *, double *);
template void launch_direct_numeric<float>(const LaunchCtx &, const u32 *, u32, const u32 *, const u32 *,
                                           const float *, const u32 *, const u32 *, const float *,
                                           const u32 *, u32 *, float *);

// ------------------------------------------------------------------------------------------
// Compare (reference source/GPU/Compare.cu:11-62): row lengths, positional column ids,
// optionally values with a relative tolerance.  One warp per row.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_compare(u32 rows, const u32 *__restrict__ rpA,
                                                 const u32 *__restrict__ ciA, const T *__restrict__ vA,
                                                 const u32 *__restrict__ rpB, const u32 *__restrict__ ciB,
                                                 const T *__restrict__ vB, bool compareData, double relTol,
                                                 Scalars *sc)
{
    const u32 row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    if (row >= rows) return;
    const u32 a0 = rpA[row], a1 = rpA[row + 1], b0 = rpB[row], b1 = rpB[row + 1];
    u64 first = ~0ull;   // smallest differing (row, kind, position) seen by this lane
    if ((a1 - a0) != (b1 - b0)) {
        first = (u64)row << 32;
    } else {
        for (u32 j = lane; j < a1 - a0 && first == ~0ull; j += 32) {
            const u32 pos = j < (1u << 28) ? j : (1u << 28) - 1u;
            if (ciA[a0 + j] != ciB[b0 + j]) first = ((u64)row << 32) | (1u << 28) | pos;
            else if (compareData) {
                const double x = (double)vA[a0 + j], y = (double)vB[b0 + j];
                const double d = fabs(x - y), s = fmax(fabs(x), fabs(y));
                if (d > relTol * s && d > 1e-300) first = ((u64)row << 32) | (2u << 28) | pos;
            }
        }
    }
    if (first != ~0ull) {
        sc->compareFlag = 1u;
        atomicMin((unsigned long long *)&sc->compareFirst, (unsigned long long)first);
    }
}

template <typename T>
void launch_compare(const LaunchCtx &lc, u32 rows, const u32 *rpA, const u32 *ciA, const T *vA, const u32 *rpB,
                    const u32 *ciB, const T *vB, bool compareData, double relTol, Scalars *sc)
{
    if (rows == 0) return;
    const u32 grid = (u32)(((u64)rows * 32 + 255) / 256);
    k_compare<T><<<grid, 256, 0, lc.stream>>>(rows, rpA, ciA, vA, rpB, ciB, vB, compareData, relTol, sc);
    ++*lc.launches;
}
template void launch_compare<double>(const LaunchCtx &, u32, const u32 *, const u32 *, const double *,
                                     const u32 *, const u32 *, const double *, bool, double, Scalars *);
template void launch_compare<float>(const LaunchCtx &, u32, const u32 *, const u32 *, const float *,
                                    const u32 *, const u32 *, const float *, bool, double, Scalars *);

}  // namespace sb
