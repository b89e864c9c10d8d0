// speck_b200/csrc/capi.cu -- C ABI (include/speck_b200.h) and host orchestration of one multiply.
//
// Stage order and the reference stage each one replaces (source/GPU/Multiply.cu):
//   analyze + bin      <- countProducts (:237-273) + loadBalanceCounting (:279-351)
//   symbolic           <- globalMapsCounting + spGEMMCounting (:357-582), no global hash maps; also records the
//                         rank map (sorted position of every product, 2 B each) for the numeric phase
//   scan               <- cub::DeviceScan::ExclusiveSum (:570) + nnz read-back (:571-575)
//   alloc C            <- allocC (:588-608), same reuse rules
//   numeric            <- loadBalanceNumeric .. sorting (:614-1050); rows are written sorted, there is no
//                         separate sorting stage and (with the rank map) no hashing or sorting at all
// Host<->device traffic per call: two 4-byte-class read-backs of one pinned Scalars struct
// (after binning, after the scan) instead of the reference's 3-8 blocking copies, and no
// cudaMalloc/cudaFree in the steady state (pooled workspace).
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/speck_b200.h"
#include "common.cuh"

using namespace sb;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(e__ == cudaErrorMemoryAllocation ? SPECK_ERR_OOM : SPECK_ERR_CUDA,            \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

constexpr int NSIDE = 4;      // side streams of the default launch plan
constexpr int NSTREAMS = 8;   // ... plus four that only the "sym_mix" plan forks

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct speck_ctx {
    int device = 0;
    int smCount = 0;
    cudaStream_t main = nullptr;
    cudaStream_t side[NSTREAMS] = {};
    cudaEvent_t evFork = nullptr, evCopy = nullptr, evJoin[NSTREAMS] = {};
    cudaEvent_t evStage[8] = {};   // 0..5 stage boundaries, 6..7 around the descriptor update that runs during the second read-back
    bool descTimed = false;
    Scalars *dSc = nullptr;
    Scalars *hSc = nullptr;  // pinned, mapped: written by k_publish (or by a plain D2H copy when spinWait is off)
    volatile u32 *hSeq = nullptr;   // sequence word next to the mirror, polled by the host
    u32 seq = 0;
    bool spinWait = true;    // mid-pipeline read-backs: poll the mapped mirror instead of cudaStreamSynchronize
    DevBuf rowOps, rowMin, rowMax, perm, tileState, bitmapStore, mapLen, mapBase, rankMap, aSeg, aOff, desc, rowInfo;
    DevBuf stage[6];          // device staging of the *_host entry points (A: rp, ci, v; B: rp, ci, v)
    void *hostOut[3] = {};    // pinned output buffers of the *_host entry points
    size_t hostOutCap[3] = {};
    speck_csr hostC = {};     // device C kept across *_host calls (reuse rules)
    size_t hostCValBytes = 0; // element size of hostC.data (f32 and f64 host calls share the context)
    u32 sortMax = RANK_MAX_PRODUCTS;  // rows with more products take the bitmap path (clamped per multiply)
    bool rankPath = true;     // rows of 513..8192 products: rank classes instead of the CTA sort classes
    int symStreams = NSIDE;   // side streams used by the symbolic phase (instruction-bound kernels overlap well)
    int numStreams = 0;       // ... and by the numeric phase besides the bitmap kernel's own stream; 0 = automatic:
                              // one stream for large multiplies (the mapped numeric kernels are memory-latency
                              // bound and run faster one after the other: R-MAT scale 20 5.0 vs 6.2 ms), all
                              // streams for small ones (launch-latency bound: webbase-like 0.24 vs 0.39 ms)
    size_t rankMapMaxBytes = ~(size_t)0;  // test hook: larger maps are treated as "does not fit" (exercises the fallbacks)
    int mapMinClass = 0;      // lane-group classes below this one (rows of <= 2 << class products) stay unmapped
    int mapCtaMin = NUM_WARP_SORT;        // first lane-group class whose mapped numeric phase uses the CTA kernel (NUM_WARP_SORT = never)
    u32 partRowCost = 0, partEntryCost = 0;   // partition_rows: cost of a row = products + partEntryCost * entries + partRowCost
    const void *hubKeyPtr = nullptr;   // A of the last multiply (row_offsets pointer, rows, nnz) and whether it had hub rows:
    size_t hubKeyRows = 0, hubKeyNnz = 0;   // the queue + k_analyze_long launch is skipped for a matrix known to have none
    bool hubRows = true;
    int tieredAnalysis = 0;   // 1: two-pass analysis (row_offsets pairs first, column extents only for rows that use them);
                              // measured slower even when B's 16-byte row summaries miss the L2 (R-MAT scale 24: analysis
                              // 1.70 -> 1.97 ms), so it is off; kept as a tested option
    bool deterministic = false;   // bit-reproducible values in sequential ascending-k summation order (slower): sort classes
                              // up to 8192 products, sequential-k kernel for every local bitmap row that fits, the
                              // remaining bitmap rows recomputed by k_det_rows
    bool testSet = true;      // local bitmap rows, symbolic: read the bitmap word before the atomicOr
    int denseSeq = 1;         // banded / high-compression rows: sequential-k numeric kernel (dense_seq.cuh): 0 = off,
                              // 1 = B segments loaded by the lanes, 2 = staged by TMA bulk copies
    int flatMinClass = NUM_WARP_SORT;   // lane-group classes >= this (6 = 256, 7 = 512 products) rank with the flat bitmap
                              // kernel instead of the register bitonic sort (mapped, two-level matrices only)
    int narrow = 2;           // lane-group classes of <= 64 (1) / <= 128 (2) products: several rows per warp, 0 = one row
                              // per warp from 17 products (config-5 matrix: 16.2 -> 15.6 ms, profiles/r2_notes.md)
    int narrowNum = -1;       // the same for the numeric kernels alone (-1: follow narrow)
    int symMix = 0;           // symbolic phase: > 0 = the lane-group sort kernels of the 128 / 256 / 512-product classes
                              // (instruction-bound) run with about this many CTAs per SM, looping over their rows, next
                              // to the bitmap rank kernels (shared-memory-bound) instead of after them
    int numPlan = 0;          // numeric phase of large multiplies (one stream): 1 = the one-CTA-per-SM kernels (bitmap rows,
                              // 4097..16384-product classes) on a second stream next to the small shapes
    int bigSplit = 0;         // mapped numeric kernel of rows of 4097 .. 16384 products: 1..3 = several CTAs per row
                              // (map_split.cuh), 0 = one 1024-thread CTA per row
    int colDirect = 0;        // mapped numeric CTA kernels: 1..3 = the large shapes stage values only and write column ids
                              // straight to C (rank_cta.cuh: COLDIRECT); measured slower (profiles/r2_notes.md), off
    int segNum = 0;           // mapped numeric CTA classes: 1 = segment-major kernel (map_seg.cuh; measured slower: fewer
                              // loads in flight per SM, profiles/r2_notes.md), 0 = k_map_rows_cta
    int flatSym = 1;          // mapped two-level symbolic rank kernel: 1 = flat staged variant (rank_flat.cuh), 0 = rank_cta.cuh
    int flatE = 8;            // ... product slots per thread of the flat variant (8 or 16)
    bool hashCount = false;   // experiment: count-only symbolic by hashing when no rank map is recorded
    bool rankMapOn = true;    // symbolic phase records every product's sorted position (2 B per product)
    u32 launches = 0;
    speck_stats stats = {};
    bool stageTimesPending = false;   // the stage times of `stats` are still in the events (read on demand: five
                                      // cudaEventElapsedTime calls cost ~8 us, a third of the host overhead of a multiply)
};

namespace {

int ensure(DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return SPECK_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SPECK_ERR_OOM, "workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return SPECK_OK;
}

void release(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// sort classes: u32 keys while (col << log2 N) fits, N = power-of-two size of the class's sorting network
bool sort_keys_wide(int sc, u32 colsB)
{
    u32 npow2 = 4u << sc;  // lane-group classes
    if (sc >= NUM_WARP_SORT) {
        npow2 = 1024;
        while (npow2 < 512u * (u32)(sc - NUM_WARP_SORT + 2)) npow2 <<= 1;
    }
    return ((u64)colsB * npow2) > (1ull << 32);
}

// Scalars to the host in the middle of a multiply.  Polling a word in mapped pinned memory costs 2-3 us against
// ~15-20 us for cudaMemcpyAsync + cudaStreamSynchronize; the stream is queried now and then so that a failed
// kernel cannot hang the host.
// publish_scalars starts the read-back, wait_scalars completes it: kernels launched in between run on the device
// while the scalars travel to the host
void publish_scalars(speck_ctx *c, LaunchCtx &lc)
{
    if (!c->spinWait) {
        cudaMemcpyAsync(c->hSc, c->dSc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->main);
        cudaEventRecord(c->evCopy, c->main);
        return;
    }
    launch_publish(lc, c->dSc, c->hSc, c->hSeq, ++c->seq);
}

int wait_scalars(speck_ctx *c)
{
    if (!c->spinWait) {
        CU_TRY(cudaEventSynchronize(c->evCopy));
        return SPECK_OK;
    }
    const u32 seq = c->seq;
    for (u64 spins = 0;; ++spins) {
        if (*c->hSeq == seq) break;
        if ((spins & 0x3fff) == 0x3fff) {
            const cudaError_t q = cudaStreamQuery(c->main);
            if (q != cudaErrorNotReady && *c->hSeq != seq) {   // the stream drained (or failed) without publishing
                if (q != cudaSuccess) return fail(SPECK_ERR_CUDA, "kernel failure before the scalar read-back: %s", cudaGetErrorString(q));
                CU_TRY(cudaStreamSynchronize(c->main));
                if (*c->hSeq == seq) break;
                return fail(SPECK_ERR_CUDA, "scalar read-back was not published");
            }
        }
    }
    return SPECK_OK;
}

int read_scalars(speck_ctx *c, LaunchCtx &lc)
{
    publish_scalars(c, lc);
    return wait_scalars(c);
}

void fork_streams(speck_ctx *c, int n = NSIDE)
{
    cudaEventRecord(c->evFork, c->main);
    for (int i = 0; i < n; ++i) cudaStreamWaitEvent(c->side[i], c->evFork, 0);
}

void join_streams(speck_ctx *c, int n = NSIDE)
{
    for (int i = 0; i < n; ++i) {
        cudaEventRecord(c->evJoin[i], c->side[i]);
        cudaStreamWaitEvent(c->main, c->evJoin[i], 0);
    }
}

// stage times of the last multiply, from its events (they stay valid until the next multiply records them again)
void finish_stage_times(speck_ctx *c)
{
    if (!c->stageTimesPending) return;
    c->stageTimesPending = false;
    speck_stats &st = c->stats;
    cudaEventElapsedTime(&st.ms_analysis, c->evStage[0], c->evStage[1]);
    cudaEventElapsedTime(&st.ms_symbolic, c->evStage[1], c->evStage[2]);
    cudaEventElapsedTime(&st.ms_scan, c->evStage[2], c->evStage[3]);
    cudaEventElapsedTime(&st.ms_numeric, c->evStage[4], c->evStage[5]);
    if (c->descTimed) {
        float d = 0.f;
        cudaEventElapsedTime(&d, c->evStage[6], c->evStage[7]);
        st.ms_numeric += d;
    }
    cudaEventElapsedTime(&st.ms_total, c->evStage[0], c->evStage[5]);
}

template <typename T>
int spgemm_impl(speck_ctx *c, const speck_csr *A, const speck_csr *B, speck_csr *C, speck_timings *tm)
{
    if (!c || !A || !B || !C) return fail(SPECK_ERR_INVALID, "null argument");
    c->stageTimesPending = false;   // the events are about to be recorded again
    // guards of the reference, source/GPU/Multiply.cu:57-70
    if (B->cols > (1u << 27)) return fail(SPECK_ERR_TOO_LARGE, "matrix B has more than %d columns (%zu)", 1 << 27, B->cols);
    if (A->rows > (1u << 27)) return fail(SPECK_ERR_TOO_LARGE, "matrix A has more than %d rows (%zu)", 1 << 27, A->rows);
    if (A->nnz == 0 || B->nnz == 0) {
        C->nnz = 0;
        c->stats = speck_stats{};
        return SPECK_OK;
    }
    if (A->cols != B->rows) return fail(SPECK_ERR_INVALID, "shape mismatch: A is %zux%zu, B is %zux%zu", A->rows, A->cols, B->rows, B->cols);
    if (!A->row_offsets || !A->col_ids || !A->data || !B->row_offsets || !B->col_ids || !B->data)
        return fail(SPECK_ERR_INVALID, "null CSR array");

    CU_TRY(cudaSetDevice(c->device));
    const u32 rows = (u32)A->rows;
    const u32 colsB = (u32)B->cols;
    const u32 *aRp = A->row_offsets, *aCi = A->col_ids, *bRp = B->row_offsets, *bCi = B->col_ids;
    const T *aV = (const T *)A->data, *bV = (const T *)B->data;
    c->launches = 0;
    LaunchCtx lc{c->main, c->smCount, &c->launches};
    // CTA-level sort classes (> 1024 products) are used only while (col << log2 N) fits a u32 key;
    // wider matrices send those rows to the bitmap path instead of sorting u64 keys.
    u32 sortMax = c->sortMax;
    const bool det = c->deterministic;
    const bool wantMap = c->rankMapOn && !det;   // mapped numeric kernels add repeated columns with shared atomics
    // rank classes (no key-width limit): two bitmap levels up to 2^20 columns, three up to 2^25 (mapped only)
    const int rankLevels = colsB <= RANK_EXTENT_LIMIT ? 2 : 3;
    const bool useRank = !det && c->rankPath && (rankLevels == 2 || (colsB <= RANK_EXTENT_LIMIT3 && wantMap));
    if (!(useRank && wantMap) && sortMax > SORT_MAX_PRODUCTS) sortMax = SORT_MAX_PRODUCTS;  // 8193..16384: mapped rank kernels only
    if (sortMax > 1024 && !useRank) {  // largest power-of-two network whose keys fit 32 bits: cols * N <= 2^32
        u32 fit = 8192;
        while (fit > 1024 && ((u64)colsB * fit) > (1ull << 32)) fit >>= 1;
        if (sortMax > fit) sortMax = fit;
    }

    // ---- init: workspace, C.row_offsets (reuse rule of Multiply.cu:155-165)
    cudaEventRecord(c->evStage[0], c->main);
    int rc;
    if ((rc = ensure(c->rowOps, (size_t)rows * 4))) return rc;
    if ((rc = ensure(c->perm, (size_t)rows * 4))) return rc;
    if ((rc = ensure(c->rowMin, (size_t)rows * 4))) return rc;
    if ((rc = ensure(c->rowMax, (size_t)rows * 4))) return rc;
    if ((rc = ensure(c->tileState, scan_tile_state_bytes(rows + 1)))) return rc;
    if (wantMap) {
        if ((rc = ensure(c->mapLen, (size_t)(rows + 1) * 4))) return rc;
        if ((rc = ensure(c->mapBase, (size_t)(rows + 1) * 8))) return rc;
        if ((rc = ensure(c->aSeg, (size_t)A->nnz * sizeof(uint2)))) return rc;
        if (c->segNum && (rc = ensure(c->aOff, (size_t)A->nnz * sizeof(u32)))) return rc;
        if ((rc = ensure(c->desc, (size_t)rows * sizeof(RowDesc)))) return rc;
    }
    if (c->tieredAnalysis != 1 && (rc = ensure(c->rowInfo, (size_t)B->rows * sizeof(uint4)))) return rc;
    u32 *cRp = C->row_offsets;
    if (!(C->rows == A->rows && cRp != nullptr)) {
        if (cRp) cudaFree(cRp);
        C->row_offsets = nullptr;
        CU_TRY(cudaMalloc(&cRp, (size_t)(rows + 1) * 4));
        C->row_offsets = cRp;
        C->rows = A->rows;
    }
    u32 *rowOps = (u32 *)c->rowOps.p, *perm = (u32 *)c->perm.p;
    u32 *rowMin = (u32 *)c->rowMin.p, *rowMax = (u32 *)c->rowMax.p;
    CU_TRY(cudaMemsetAsync(c->dSc, 0, sizeof(Scalars), c->main));

    // ---- analysis + binning
    // B-row summaries: 16 B per row of B, one gather per A entry (option tiered_analysis = 1: row_offsets pairs only,
    // column extents in a second pass for the rows that use them)
    const bool tiered = c->tieredAnalysis == 1;
    // hub rows of A (>= 1024 entries) are queued for one CTA each; a matrix seen before without any skips that launch
    const bool sameA = c->hubKeyPtr == (const void *)aRp && c->hubKeyRows == A->rows && c->hubKeyNnz == A->nnz;
    const bool useHubQueue = !sameA || c->hubRows;
    if (!tiered) launch_row_info(lc, (u32)B->rows, bRp, bCi, (uint4 *)c->rowInfo.p);
    launch_analyze(lc, rows, A->nnz, aRp, aCi, bRp, bCi, rowOps, rowMin, rowMax, cRp, c->dSc, sortMax,
                   wantMap ? (uint2 *)c->aSeg.p : nullptr, tiered ? nullptr : (const uint4 *)c->rowInfo.p,
                   wantMap && c->segNum ? (u32 *)c->aOff.p : nullptr, wantMap ? (u32 *)c->mapLen.p : nullptr, useRank,
                   c->mapMinClass, tiered ? min(128u, sortMax + 1u) : 0u, useHubQueue ? perm /* free until k_bin_scatter */ : nullptr);
    if (wantMap) launch_scan_map(lc, (const u32 *)c->mapLen.p, (u64 *)c->mapBase.p, rows + 1, (u64 *)c->tileState.p, c->dSc);
    launch_bin_scatter(lc, rows, aRp, rowOps, rowMin, rowMax, perm, c->dSc, sortMax, wantMap ? (const u64 *)c->mapBase.p : nullptr,
                       wantMap ? (RowDesc *)c->desc.p : nullptr);
    cudaEventRecord(c->evStage[1], c->main);
    if ((rc = read_scalars(c, lc))) return rc;
    const Scalars s1 = *c->hSc;
    c->hubKeyPtr = aRp; c->hubKeyRows = A->rows; c->hubKeyNnz = A->nnz;
    if (useHubQueue) c->hubRows = s1.longCount != 0;
    if (s1.products == 0) {  // Multiply.cu:256-261: alloc(rows, cols, 0, false)
        if (C->data) cudaFree(C->data);
        if (C->col_ids) cudaFree(C->col_ids);
        if (C->row_offsets) cudaFree(C->row_offsets);
        C->data = nullptr; C->col_ids = nullptr; C->row_offsets = nullptr;
        C->rows = A->rows; C->cols = B->cols; C->nnz = 0;
        c->stats = speck_stats{};
        return SPECK_OK;
    }
    u32 binStart[NUM_BINS + 1];
    binStart[0] = 0;
    for (int b = 0; b < NUM_BINS; ++b) binStart[b + 1] = binStart[b] + s1.binCount[b];

    // first bin of each rank launch group: CTA class c holds rows of <= 512 * (c + 2) products
    const int cta0 = BIN_SORT0 + NUM_WARP_SORT;
    const int rankGroupFirst[6] = {cta0, cta0 + 1, cta0 + 3, cta0 + 7, cta0 + 15, cta0 + NUM_CTA_SORT};
    constexpr int RANK_GROUPS = 5;   // products <= 1024 << g

    // bitmaps of the local-dense rows are kept from the symbolic to the numeric phase (2 KB per row)
    u32 *bitmapStore = nullptr;
    {
        const size_t need = dense_local_store_bytes(s1.binCount[BIN_DENSE_LOCAL]);
        if (need && need <= ((size_t)8 << 30) && ensure(c->bitmapStore, need) == SPECK_OK) bitmapStore = (u32 *)c->bitmapStore.p;
    }
    // rank map (2 B per product of the mapped rows): optional -- without it (switched off, or no memory) the
    // numeric phase recomputes the positions with the self-contained kernels
    unsigned short *rankMap = nullptr;
    RowDesc *desc = nullptr;
    const uint2 *aSeg = nullptr;
    if (wantMap && s1.mapTotal) {
        const size_t need = (size_t)s1.mapTotal * 2;
        bool fits = need <= c->rankMap.cap && need <= c->rankMapMaxBytes;
        if (!fits && need <= c->rankMapMaxBytes) {  // grow only while it leaves at least half of the free memory to C
            size_t freeB = 0, totalB = 0;
            cudaMemGetInfo(&freeB, &totalB);
            fits = need < (freeB + c->rankMap.cap) / 2;
        }
        if (fits && ensure(c->rankMap, need) == SPECK_OK) {
            rankMap = (unsigned short *)c->rankMap.p;
            desc = (RowDesc *)c->desc.p;
            aSeg = (const uint2 *)c->aSeg.p;
            // descriptors of all binned rows are in place, perm order (desc + binStart[b] = first row of bin b)
        }
    }
    // ---- symbolic: dense rows first (longest), then the sort classes from large to small
    const bool mix = c->symMix > 0 && useRank && desc && rankLevels == 2;
    const int symForked = mix ? NSTREAMS : NSIDE;
    fork_streams(c, symForked);
    int sidx = 0;
    bool sortDone[NUM_SORT] = {};
    if (mix) {
        // "sym_mix": the three largest lane-group classes first, each on its own stream with a capped grid (shares of
        // sym_mix CTAs per SM in proportion to their products), so that the rank kernels launched next fill the rest of
        // every SM: the bitonic sort is bound by instruction issue, the bitmap kernels by the shared-memory pipe
        u64 work[NUM_WARP_SORT] = {}, total = 0;
        for (int sc = NUM_WARP_SORT - 3; sc < NUM_WARP_SORT; ++sc)
            if (sc >= c->mapMinClass) total += work[sc] = (u64)s1.binCount[BIN_SORT0 + sc] << (sc + 2);
        for (int sc = NUM_WARP_SORT - 1; sc >= NUM_WARP_SORT - 3 && total; --sc) {
            const u32 cnt = s1.binCount[BIN_SORT0 + sc];
            if (!cnt || !work[sc] || (rankLevels == 3 && sort_keys_wide(sc, colsB))) continue;
            LaunchCtx ls{c->side[NSIDE + (NUM_WARP_SORT - 1 - sc)], c->smCount, &c->launches};
            const u64 share = ((u64)c->smCount * (u64)c->symMix * work[sc] + total - 1) / total;
            ls.gridCap = (u32)max((u64)c->smCount, share);
            launch_sort_symbolic(ls, sc, sort_keys_wide(sc, colsB), perm + binStart[BIN_SORT0 + sc], cnt, aRp, aCi, bRp, bCi,
                                 rowOps, cRp, desc + binStart[BIN_SORT0 + sc], aSeg, rankMap);
            sortDone[sc] = true;
        }
    }
    for (int loc = 0; loc < 2; ++loc) {
        const int bin = loc ? BIN_DENSE_LOCAL : BIN_DENSE;
        if (!s1.binCount[bin]) continue;
        LaunchCtx ls{c->side[sidx++ % c->symStreams], c->smCount, &c->launches};
        launch_dense_symbolic(ls, loc != 0, perm + binStart[bin], s1.binCount[bin], &c->dSc->denseCounter[loc], aRp,
                              aCi, bRp, bCi, colsB, rowMin, rowMax, loc ? bitmapStore : nullptr, cRp,
                              rowOps, (loc && (c->denseSeq || det)) ? c->dSc->seqRows : nullptr, det, loc && c->testSet);
    }
    // rank classes: the CTA sort bins grouped by launch shape (products <= 8192 / 4096 / 2048 / 1024)
    if (useRank) {
        for (int g = RANK_GROUPS - 1; g >= 0; --g) {
            const int b0 = rankGroupFirst[g], b1 = rankGroupFirst[g + 1];
            const u32 cnt = binStart[b1] - binStart[b0];
            if (!cnt) continue;
            LaunchCtx ls{c->side[sidx++ % c->symStreams], c->smCount, &c->launches};
            if (g == RANK_GROUPS - 1 && !rankMap) {  // no rank map could be allocated: the bitmap kernel takes these rows
                launch_dense_symbolic(ls, false, perm + binStart[b0], cnt, &c->dSc->denseCounter[4], aRp, aCi, bRp, bCi, colsB,
                                      rowMin, rowMax, nullptr, cRp);
                continue;
            }
            if (rankMap && rankLevels == 2 && c->flatSym)
                launch_rank_flat(ls, 1024u << g, c->flatE, desc + binStart[b0], cnt, aSeg, bCi, rankMap, cRp);
            else if (!rankMap && c->hashCount && g < RANK_GROUPS - 1)
                launch_hash_count(ls, 1024u << g, perm + binStart[b0], cnt, aRp, aCi, bRp, bCi, cRp);
            else
                launch_rank_symbolic(ls, 1024u << g, perm + binStart[b0], cnt, aRp, aCi, bRp, bCi, rowOps, rowMin, rowMax, cRp,
                                     desc ? desc + binStart[b0] : nullptr, aSeg, rankMap, rankLevels);
        }
    }
    for (int sc = (useRank ? NUM_WARP_SORT : NUM_SORT) - 1; sc >= 0; --sc) {
        const u32 cnt = s1.binCount[BIN_SORT0 + sc];
        if (!cnt || sortDone[sc]) continue;
        LaunchCtx ls{c->side[sidx++ % c->symStreams], c->smCount, &c->launches};
        const bool mapped = desc && sc >= c->mapMinClass;
        ls.narrow = mapped ? c->narrow : 0;
        // 64-bit sort keys cost 2-3x (R-MAT scale 24: 28 ps per product in the 512 class): wide matrices rank these
        // rows with the three-level bitmap kernel (u32 throughout) instead
        if (mapped && useRank && rankLevels == 3 && sc >= 6 && sc < NUM_WARP_SORT && sort_keys_wide(sc, colsB))
            launch_rank_symbolic(ls, 4u << sc, perm + binStart[BIN_SORT0 + sc], cnt, aRp, aCi, bRp, bCi, rowOps, rowMin, rowMax,
                                 cRp, desc + binStart[BIN_SORT0 + sc], aSeg, rankMap, 3);
        else if (mapped && useRank && rankLevels == 2 && c->flatSym && sc >= c->flatMinClass && sc >= 6)
            // 129..512 products: the flat bitmap kernel on 32 / 64 threads instead of the register bitonic sort
            launch_rank_flat(ls, 4u << sc, 8, desc + binStart[BIN_SORT0 + sc], cnt, aSeg, bCi, rankMap, cRp);
        else
            launch_sort_symbolic(ls, sc, sort_keys_wide(sc, colsB), perm + binStart[BIN_SORT0 + sc], cnt, aRp, aCi, bRp, bCi,
                                 rowOps, cRp, mapped ? desc + binStart[BIN_SORT0 + sc] : nullptr, aSeg, rankMap);
    }
    join_streams(c, symForked);
    CU_TRY(cudaGetLastError());   // launch / attribute errors of the symbolic kernels (side streams are joined)
    cudaEventRecord(c->evStage[2], c->main);

    // ---- scan + nnz read-back
    launch_scan(lc, cRp, rows + 1, (u64 *)c->tileState.p, c->dSc);
    cudaEventRecord(c->evStage[3], c->main);
    publish_scalars(c, lc);
    // the descriptors take their C offsets while nnz(C) travels to the host (timed with the numeric phase)
    c->descTimed = rankMap != nullptr;
    if (rankMap) {
        cudaEventRecord(c->evStage[6], c->main);
        launch_desc_numeric(lc, binStart[NUM_BINS], cRp, desc);
        cudaEventRecord(c->evStage[7], c->main);
    }
    if ((rc = wait_scalars(c))) return rc;
    CU_TRY(cudaGetLastError());
    const u64 nnzC = c->hSc->nnzC;
    const int seqKind = c->denseSeq ? c->denseSeq : (det ? 1 : 0);
    const int denseSeq = seqKind ? (seqKind | (c->hSc->seqRows[0] ? 4 : 0) | (c->hSc->seqRows[1] ? 8 : 0) | (det ? 16 : 0)) : 0;
    if (nnzC > 0xffffffffull) return fail(SPECK_ERR_OVERFLOW, "nnz(C) = %llu does not fit the u32 row_offsets of the spECK API", (unsigned long long)nnzC);

    // ---- alloc C (Multiply.cu:589-602): only when nnz changed
    if (C->nnz != nnzC || !C->data || !C->col_ids) {
        if (C->data) cudaFree(C->data);
        if (C->col_ids) cudaFree(C->col_ids);
        C->data = nullptr; C->col_ids = nullptr; C->nnz = 0;
        cudaError_t e1 = cudaMalloc(&C->data, nnzC * sizeof(T));
        cudaError_t e2 = cudaMalloc((void **)&C->col_ids, nnzC * sizeof(u32));
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            cudaGetLastError();
            if (C->data) cudaFree(C->data);
            if (C->col_ids) cudaFree(C->col_ids);
            C->data = nullptr; C->col_ids = nullptr;
            return fail(SPECK_ERR_OOM, "out of memory allocating C with %llu non-zeros", (unsigned long long)nnzC);
        }
    }
    C->nnz = nnzC;
    C->cols = B->cols;
    u32 *cCi = C->col_ids;
    T *cV = (T *)C->data;

    // ---- numeric
    const int numStreams = c->numStreams ? c->numStreams : (s1.products >= (1ull << 27) ? 1 : NSIDE);
    const bool bigApart = c->numPlan == 1 && numStreams == 1;
    cudaEventRecord(c->evStage[4], c->main);
    fork_streams(c);
    sidx = 0;
    for (int loc = 0; loc < 2; ++loc) {
        const int bin = loc ? BIN_DENSE_LOCAL : BIN_DENSE;
        if (!s1.binCount[bin]) continue;
        LaunchCtx ls{c->side[bigApart ? 1 : NSIDE - 1], c->smCount, &c->launches};
        launch_dense_numeric<T>(ls, loc != 0, perm + binStart[bin], s1.binCount[bin], &c->dSc->denseCounter[2 + loc],
                                aRp, aCi, aV, bRp, bCi, bV, colsB, rowMin, rowMax, loc ? bitmapStore : nullptr, cRp, cCi, cV,
                                denseSeq, rowOps);
        if (det)   // same stream: the bitmap kernel has written the row's columns
            launch_det_rows<T>(ls, perm + binStart[bin], s1.binCount[bin], aRp, aCi, aV, bRp, bCi, bV, cRp, cCi, cV, rowOps,
                               loc && bitmapStore && (denseSeq & 12));
    }
    if (useRank) {
        for (int g = RANK_GROUPS - 1; g >= 0; --g) {
            const int b0 = rankGroupFirst[g], b1 = rankGroupFirst[g + 1];
            const u32 cnt = binStart[b1] - binStart[b0];
            if (!cnt) continue;
            LaunchCtx ls{c->side[bigApart ? (g >= 3 ? 1 : 0) : sidx++ % numStreams], c->smCount, &c->launches};
            if (g == RANK_GROUPS - 1 && !rankMap) {
                launch_dense_numeric<T>(ls, false, perm + binStart[b0], cnt, &c->dSc->denseCounter[5], aRp, aCi, aV, bRp, bCi,
                                        bV, colsB, rowMin, rowMax, nullptr, cRp, cCi, cV);
                continue;
            }
            if (rankMap && c->segNum)
                launch_map_seg<T>(ls, 1024u << g, desc + binStart[b0], cnt, aSeg, (const u32 *)c->aOff.p, aV, bCi, bV, rankMap, cCi, cV);
            else if (rankMap)
                launch_map_numeric_cta<T>(ls, 1024u << g, desc + binStart[b0], cnt, aSeg, aV, bCi, bV, rankMap, cCi, cV, c->colDirect, c->bigSplit);
            else if (rankLevels == 2)
                launch_rank_numeric<T>(ls, 1024u << g, perm + binStart[b0], cnt, aRp, aCi, aV, bRp, bCi, bV, rowOps, rowMin,
                                       rowMax, cRp, cCi, cV);
            else  // wide matrix and no rank map: CTA bitonic classes with 64-bit keys
                for (int b = b0; b < b1; ++b)
                    launch_sort_numeric<T>(ls, b - BIN_SORT0, sort_keys_wide(b - BIN_SORT0, colsB), perm + binStart[b],
                                           s1.binCount[b], aRp, aCi, aV, bRp, bCi, bV, rowOps, cRp, cCi, cV);
        }
    }
    for (int sc = (useRank ? NUM_WARP_SORT : NUM_SORT) - 1; sc >= 0; --sc) {
        const u32 cnt = s1.binCount[BIN_SORT0 + sc];
        if (!cnt) continue;
        LaunchCtx ls{c->side[sidx++ % numStreams], c->smCount, &c->launches};
        const bool mapped = rankMap && sc < NUM_WARP_SORT && sc >= c->mapMinClass;
        ls.narrow = mapped ? (c->narrowNum >= 0 ? c->narrowNum : c->narrow) : 0;
        if (mapped && sc >= c->mapCtaMin)   // rows of <= 4 << sc products: 32 / 64 threads x 8 slots
            launch_map_numeric_cta<T>(ls, 4u << sc, desc + binStart[BIN_SORT0 + sc], cnt, aSeg, aV, bCi, bV, rankMap, cCi, cV);
        else if (mapped)
            launch_map_numeric<T>(ls, sc, desc + binStart[BIN_SORT0 + sc], cnt, aSeg, aV, bCi, bV, rankMap, cCi, cV);
        else
            launch_sort_numeric<T>(ls, sc, sort_keys_wide(sc, colsB), perm + binStart[BIN_SORT0 + sc], cnt, aRp, aCi, aV,
                                   bRp, bCi, bV, rowOps, cRp, cCi, cV);
    }
    {
        LaunchCtx ls{c->side[sidx++ % numStreams], c->smCount, &c->launches};
        launch_direct_numeric<T>(ls, perm + binStart[BIN_DIRECT], s1.binCount[BIN_DIRECT], aRp, aCi, aV, bRp, bCi, bV,
                                 cRp, cCi, cV);
    }
    join_streams(c);
    cudaEventRecord(c->evStage[5], c->main);
    CU_TRY(cudaStreamSynchronize(c->main));
    CU_TRY(cudaGetLastError());

    // ---- stats / timings
    speck_stats &st = c->stats;
    st = speck_stats{};
    st.products = s1.products;
    st.nnz_c = nnzC;
    st.max_row_products = s1.maxRowProducts;
    for (int b = 0; b < NUM_BINS; ++b) st.class_rows[b] = s1.binCount[b];
    st.kernel_launches = c->launches;
    c->stageTimesPending = true;   // read by speck_b200_get_stats / the sharded summary / below when timings are asked for
    st.workspace_bytes = c->rowOps.cap + c->perm.cap + c->rowMin.cap + c->rowMax.cap + c->tileState.cap + c->bitmapStore.cap +
                         c->mapLen.cap + c->mapBase.cap + c->rankMap.cap + c->aSeg.cap + c->aOff.cap + c->desc.cap + c->rowInfo.cap;
    if (tm) {
        finish_stage_times(c);
        float allocMs = 0.f;
        cudaEventElapsedTime(&allocMs, c->evStage[3], c->evStage[4]);
        tm->init = 0.f;
        tm->count_products = st.ms_analysis;
        tm->load_balance_counting = 0.f;   // binning is fused into the analysis stage
        tm->global_maps_counting = 0.f;    // no global hash maps exist
        tm->spgemm_counting = st.ms_symbolic + st.ms_scan;
        tm->alloc_c = allocMs;
        tm->load_balance_numeric = 0.f;    // the same binning serves both phases
        tm->global_maps_numeric = 0.f;
        tm->spgemm_numeric = st.ms_numeric;
        tm->sorting = 0.f;                 // rows are emitted sorted by the numeric kernels
        tm->cleanup = 0.f;                 // pooled workspace, nothing to free
        tm->complete = st.ms_total;
    }
    return SPECK_OK;
}

int ensure_host(speck_ctx *c, int i, size_t bytes)
{
    if (bytes <= c->hostOutCap[i]) return SPECK_OK;
    if (c->hostOut[i]) cudaFreeHost(c->hostOut[i]);
    c->hostOut[i] = nullptr;
    c->hostOutCap[i] = 0;
    size_t want = bytes + bytes / 16 + 4096;
    cudaError_t e = cudaMallocHost(&c->hostOut[i], want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SPECK_ERR_OOM, "pinned host allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    c->hostOutCap[i] = want;
    return SPECK_OK;
}

template <typename T>
int spgemm_host_impl(speck_ctx *c, const speck_csr *A, const speck_csr *B, speck_csr *C, uint64_t *h2d, uint64_t *d2h)
{
    if (!c || !A || !B || !C) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    const bool alias = (A->row_offsets == B->row_offsets && A->col_ids == B->col_ids && A->data == B->data &&
                        A->rows == B->rows && A->nnz == B->nnz);
    uint64_t up = 0;
    speck_csr dA = *A, dB = *B;
    const speck_csr *src[2] = {A, B};
    speck_csr *dst[2] = {&dA, &dB};
    for (int m = 0; m < (alias ? 1 : 2); ++m) {
        const speck_csr *h = src[m];
        const size_t bytes[3] = {(h->rows + 1) * 4, h->nnz * 4, h->nnz * sizeof(T)};
        const void *hp[3] = {h->row_offsets, h->col_ids, h->data};
        for (int j = 0; j < 3; ++j) {
            int rc = ensure(c->stage[m * 3 + j], bytes[j] ? bytes[j] : 4);
            if (rc) return rc;
            if (bytes[j]) CU_TRY(cudaMemcpyAsync(c->stage[m * 3 + j].p, hp[j], bytes[j], cudaMemcpyHostToDevice, c->main));
            up += bytes[j];
        }
        dst[m]->row_offsets = (u32 *)c->stage[m * 3 + 0].p;
        dst[m]->col_ids = (u32 *)c->stage[m * 3 + 1].p;
        dst[m]->data = c->stage[m * 3 + 2].p;
    }
    if (alias) { dB = dA; dB.rows = B->rows; dB.cols = B->cols; }
    if (c->hostCValBytes != sizeof(T) && c->hostC.data) {   // the kept value buffer was sized for the other type
        cudaFree(c->hostC.data);
        c->hostC.data = nullptr;
        c->hostC.nnz = 0;
    }
    c->hostCValBytes = sizeof(T);
    int rc = spgemm_impl<T>(c, &dA, &dB, &c->hostC, nullptr);
    if (rc) return rc;
    const speck_csr &dC = c->hostC;
    uint64_t down = 0;
    C->rows = A->rows; C->cols = B->cols; C->nnz = dC.nnz;
    C->row_offsets = nullptr; C->col_ids = nullptr; C->data = nullptr;
    if (dC.nnz == 0) {
        // empty product (A.nnz == 0, B.nnz == 0 or P == 0): the device conventions of the reference leave
        // row_offsets stale or null (Multiply.cu:67-70, 256-261); the host entry returns a well-formed empty CSR
        const size_t bytes = (A->rows + 1) * 4;
        if ((rc = ensure_host(c, 0, bytes))) return rc;
        memset(c->hostOut[0], 0, bytes);
        C->row_offsets = (u32 *)c->hostOut[0];
    } else if (dC.row_offsets) {
        const size_t bytes[3] = {(dC.rows + 1) * 4, dC.nnz * 4, dC.nnz * sizeof(T)};
        const void *dp[3] = {dC.row_offsets, dC.col_ids, dC.data};
        for (int j = 0; j < 3; ++j) {
            if ((rc = ensure_host(c, j, bytes[j] ? bytes[j] : 4))) return rc;
            if (bytes[j]) CU_TRY(cudaMemcpyAsync(c->hostOut[j], dp[j], bytes[j], cudaMemcpyDeviceToHost, c->main));
            down += bytes[j];
        }
        CU_TRY(cudaStreamSynchronize(c->main));
        C->row_offsets = (u32 *)c->hostOut[0];
        C->col_ids = (u32 *)c->hostOut[1];
        C->data = c->hostOut[2];
    }
    if (h2d) *h2d = up;
    if (d2h) *d2h = down;
    return SPECK_OK;
}

template <typename T>
int compare_impl(speck_ctx *c, const speck_csr *ref, const speck_csr *cmp, int compareData, double relTol,
                 speck_mismatch *first = nullptr)
{
    if (!c || !ref || !cmp) return fail(SPECK_ERR_INVALID, "null argument");
    if (first) *first = speck_mismatch{};
    if (ref->rows != cmp->rows || ref->cols != cmp->cols || ref->nnz != cmp->nnz) {
        if (first) first->kind = 3;
        if (ref->rows != cmp->rows || ref->cols != cmp->cols || !first) return 0;
    }
    if (ref->nnz == 0 && cmp->nnz == 0) return 1;
    if (!ref->row_offsets || !cmp->row_offsets || !ref->col_ids || !cmp->col_ids) {
        if (first) first->kind = 3;
        return 0;
    }
    if (compareData && (!ref->data || !cmp->data)) return fail(SPECK_ERR_INVALID, "compare_data set but a value array is null");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaMemsetAsync(&c->dSc->compareFlag, 0, 4, c->main));
    CU_TRY(cudaMemsetAsync(&c->dSc->compareFirst, 0xff, 8, c->main));
    u32 n = 0;
    LaunchCtx lc{c->main, c->smCount, &n};
    launch_compare<T>(lc, (u32)ref->rows, ref->row_offsets, ref->col_ids, (const T *)ref->data, cmp->row_offsets,
                      cmp->col_ids, (const T *)cmp->data, compareData != 0, relTol, c->dSc);
    CU_TRY(cudaMemcpyAsync(&c->hSc->compareFlag, &c->dSc->compareFlag, 4, cudaMemcpyDeviceToHost, c->main));
    CU_TRY(cudaMemcpyAsync(&c->hSc->compareFirst, &c->dSc->compareFirst, 8, cudaMemcpyDeviceToHost, c->main));
    CU_TRY(cudaStreamSynchronize(c->main));
    CU_TRY(cudaGetLastError());
    if (c->hSc->compareFlag == 0) return ref->nnz == cmp->nnz ? 1 : 0;
    if (first) {   // details of the first difference: a handful of 4- and 8-byte reads
        const u64 key = c->hSc->compareFirst;
        const u32 row = (u32)(key >> 32), kind = (u32)(key >> 28) & 0xfu, pos = (u32)key & ((1u << 28) - 1u);
        first->row = row;
        first->kind = kind;
        first->index_in_row = pos;
        u32 ra[2] = {}, rb[2] = {};
        CU_TRY(cudaMemcpy(ra, ref->row_offsets + row, 8, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(rb, cmp->row_offsets + row, 8, cudaMemcpyDeviceToHost));
        first->ref_len = ra[1] - ra[0];
        first->cmp_len = rb[1] - rb[0];
        if (kind >= 1) {
            CU_TRY(cudaMemcpy(&first->ref_col, ref->col_ids + ra[0] + pos, 4, cudaMemcpyDeviceToHost));
            CU_TRY(cudaMemcpy(&first->cmp_col, cmp->col_ids + rb[0] + pos, 4, cudaMemcpyDeviceToHost));
            T va = 0, vb = 0;
            if (ref->data) CU_TRY(cudaMemcpy(&va, (const T *)ref->data + ra[0] + pos, sizeof(T), cudaMemcpyDeviceToHost));
            if (cmp->data) CU_TRY(cudaMemcpy(&vb, (const T *)cmp->data + rb[0] + pos, sizeof(T), cudaMemcpyDeviceToHost));
            first->ref_val = (double)va;
            first->cmp_val = (double)vb;
        }
    }
    return 0;
}

}  // namespace

template <typename T>
int push_slab_impl(speck_ctx *c, const speck_csr *S, uint64_t nnzBase, uint32_t rowBase, int last, uint32_t *dstRp,
                   uint32_t *dstCi, T *dstV, float *deviceMs)
{
    if (!c || !S || !dstRp) return fail(SPECK_ERR_INVALID, "null argument");
    if (deviceMs) *deviceMs = 0.f;
    if (S->nnz && (!S->col_ids || !S->data || !S->row_offsets || !dstCi || !dstV)) return fail(SPECK_ERR_INVALID, "null CSR array");
    if (nnzBase + S->nnz > 0xffffffffull)
        return fail(SPECK_ERR_OVERFLOW, "concatenated nnz(C) does not fit the u32 row_offsets of the spECK API: keep C distributed");
    CU_TRY(cudaSetDevice(c->device));
    const u32 rowsOut = (u32)S->rows + (last ? 1u : 0u);
    u32 launches = 0;
    LaunchCtx lc{c->main, c->smCount, &launches};
    finish_stage_times(c);   // two of the stage events are borrowed below: read the last multiply's times first
    cudaEventRecord(c->evStage[6], c->main);
    if (S->nnz && S->row_offsets) {
        launch_push_slab<T>(lc, S->row_offsets, S->col_ids, (const T *)S->data, rowsOut, S->nnz, nnzBase, rowBase, dstRp,
                            dstCi, dstV);
    } else if (rowsOut) {   // empty slab (the device conventions leave its row_offsets stale or null): every row starts at nnzBase
        std::vector<u32> fill((size_t)rowsOut, (u32)nnzBase);
        CU_TRY(cudaMemcpyAsync(dstRp + rowBase, fill.data(), fill.size() * 4, cudaMemcpyHostToDevice, c->main));
        CU_TRY(cudaStreamSynchronize(c->main));
    }
    cudaEventRecord(c->evStage[7], c->main);
    CU_TRY(cudaStreamSynchronize(c->main));
    CU_TRY(cudaGetLastError());
    if (deviceMs) cudaEventElapsedTime(deviceMs, c->evStage[6], c->evStage[7]);
    c->descTimed = false;   // the two events were borrowed from the stage timers
    return SPECK_OK;
}

extern "C" {

int speck_b200_abi_version(void) { return SPECK_B200_ABI_VERSION; }
const char *speck_b200_last_error(void) { return g_err; }

int speck_b200_destroy(speck_ctx *c);

int speck_b200_create(int device, speck_ctx **out)
{
    if (!out) return fail(SPECK_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SPECK_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(SPECK_ERR_NO_DEVICE, "device %d out of range (%d devices)", device, n);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(SPECK_ERR_NO_DEVICE, "device %d is sm_%d%d; this build targets sm_100a (B200) only", device, prop.major, prop.minor);
    CU_TRY(cudaSetDevice(device));
    speck_ctx *c = new (std::nothrow) speck_ctx();
    if (!c) return fail(SPECK_ERR_OOM, "host allocation failed");
    c->device = device;
    c->smCount = prop.multiProcessorCount;
    // The main stream is a blocking stream: uploads made by the host layer with plain cudaMemcpy / cudaMemset on
    // the legacy default stream (host/dCSR.cpp) are ordered before the multiply without an explicit
    // synchronisation.  The side streams only ever run between a fork from and a join into the main stream.
    const int rc = [&]() -> int {
        CU_TRY(cudaStreamCreate(&c->main));
        for (int i = 0; i < NSTREAMS; ++i) {
            CU_TRY(cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
            CU_TRY(cudaEventCreateWithFlags(&c->evJoin[i], cudaEventDisableTiming));
        }
        CU_TRY(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->evCopy, cudaEventDisableTiming));
        for (auto &e : c->evStage) CU_TRY(cudaEventCreate(&e));
        CU_TRY(cudaMalloc(&c->dSc, sizeof(Scalars)));
        CU_TRY(cudaHostAlloc(&c->hSc, sizeof(Scalars) + 64, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(c->hSc, 0, sizeof(Scalars) + 64);
        c->hSeq = reinterpret_cast<volatile u32 *>(reinterpret_cast<char *>(c->hSc) + ((sizeof(Scalars) + 15) / 16) * 16);
        return SPECK_OK;
    }();
    if (rc != SPECK_OK) {   // do not leak the partially built context
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        speck_b200_destroy(c);
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    *out = c;
    return SPECK_OK;
}

int speck_b200_destroy(speck_ctx *c)
{
    if (!c) return SPECK_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    release(c->rowOps); release(c->perm); release(c->rowMin); release(c->rowMax); release(c->tileState); release(c->bitmapStore); release(c->mapLen); release(c->mapBase); release(c->rankMap); release(c->aSeg); release(c->aOff); release(c->desc); release(c->rowInfo);
    for (auto &b : c->stage) release(b);
    for (auto &h : c->hostOut) if (h) cudaFreeHost(h);
    if (c->hostC.data) cudaFree(c->hostC.data);
    if (c->hostC.col_ids) cudaFree(c->hostC.col_ids);
    if (c->hostC.row_offsets) cudaFree(c->hostC.row_offsets);
    if (c->dSc) cudaFree(c->dSc);
    if (c->hSc) cudaFreeHost(c->hSc);
    for (auto &e : c->evStage) if (e) cudaEventDestroy(e);
    for (int i = 0; i < NSTREAMS; ++i) {
        if (c->evJoin[i]) cudaEventDestroy(c->evJoin[i]);
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
    }
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evCopy) cudaEventDestroy(c->evCopy);
    if (c->main) cudaStreamDestroy(c->main);
    delete c;
    return SPECK_OK;
}

int speck_b200_sm_count(const speck_ctx *c) { return c ? c->smCount : 0; }

int speck_b200_spgemm_f64(speck_ctx *c, const speck_csr *A, const speck_csr *B, speck_csr *C, speck_timings *t)
{
    return spgemm_impl<double>(c, A, B, C, t);
}
int speck_b200_spgemm_f32(speck_ctx *c, const speck_csr *A, const speck_csr *B, speck_csr *C, speck_timings *t)
{
    return spgemm_impl<float>(c, A, B, C, t);
}
int speck_b200_spgemm_host_f64(speck_ctx *c, const speck_csr *A, const speck_csr *B, speck_csr *C, uint64_t *h2d, uint64_t *d2h)
{
    return spgemm_host_impl<double>(c, A, B, C, h2d, d2h);
}
int speck_b200_spgemm_host_f32(speck_ctx *c, const speck_csr *A, const speck_csr *B, speck_csr *C, uint64_t *h2d, uint64_t *d2h)
{
    return spgemm_host_impl<float>(c, A, B, C, h2d, d2h);
}

int speck_b200_get_stats(const speck_ctx *c, speck_stats *out)
{
    if (!c || !out) return fail(SPECK_ERR_INVALID, "null argument");
    cudaSetDevice(c->device);
    finish_stage_times(const_cast<speck_ctx *>(c));   // the stage times are read from the events on first use
    *out = c->stats;
    return SPECK_OK;
}

int speck_b200_row_products(speck_ctx *c, const speck_csr *A, const speck_csr *B, uint32_t *dRowOps, uint64_t *products, uint32_t *maxRow)
{
    if (!c || !A || !B) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    if (products) *products = 0;
    if (maxRow) *maxRow = 0;
    if (A->rows == 0 || A->nnz == 0 || B->nnz == 0) {
        if (dRowOps && A->rows) CU_TRY(cudaMemsetAsync(dRowOps, 0, A->rows * 4, c->main));
        CU_TRY(cudaStreamSynchronize(c->main));
        return SPECK_OK;
    }
    const u32 rows = (u32)A->rows;
    int rc;
    if ((rc = ensure(c->rowOps, (size_t)rows * 4))) return rc;
    if ((rc = ensure(c->perm, (size_t)rows * 4))) return rc;  // scratch for the rowNnz side output
    if ((rc = ensure(c->rowMin, (size_t)rows * 4))) return rc;
    if ((rc = ensure(c->rowMax, (size_t)rows * 4))) return rc;
    CU_TRY(cudaMemsetAsync(c->dSc, 0, sizeof(Scalars), c->main));
    u32 n = 0;
    LaunchCtx lc{c->main, c->smCount, &n};
    launch_analyze(lc, rows, A->nnz, A->row_offsets, A->col_ids, B->row_offsets, B->col_ids, (u32 *)c->rowOps.p,
                   (u32 *)c->rowMin.p, (u32 *)c->rowMax.p, (u32 *)c->perm.p, c->dSc, c->sortMax, nullptr, nullptr);
    if (dRowOps) CU_TRY(cudaMemcpyAsync(dRowOps, c->rowOps.p, (size_t)rows * 4, cudaMemcpyDeviceToDevice, c->main));
    CU_TRY(cudaMemcpyAsync(c->hSc, c->dSc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->main));
    CU_TRY(cudaStreamSynchronize(c->main));
    CU_TRY(cudaGetLastError());
    if (products) *products = c->hSc->products;
    if (maxRow) *maxRow = c->hSc->maxRowProducts;
    return SPECK_OK;
}

int speck_b200_compare_f64(speck_ctx *c, const speck_csr *r, const speck_csr *m, int cd, double tol)
{
    return compare_impl<double>(c, r, m, cd, tol);
}
int speck_b200_compare_f32(speck_ctx *c, const speck_csr *r, const speck_csr *m, int cd, double tol)
{
    return compare_impl<float>(c, r, m, cd, tol);
}

int speck_b200_compare_report_f64(speck_ctx *c, const speck_csr *r, const speck_csr *m, int cd, double tol, speck_mismatch *first)
{
    return compare_impl<double>(c, r, m, cd, tol, first);
}
int speck_b200_compare_report_f32(speck_ctx *c, const speck_csr *r, const speck_csr *m, int cd, double tol, speck_mismatch *first)
{
    return compare_impl<float>(c, r, m, cd, tol, first);
}

int speck_b200_malloc(speck_ctx *c, void **dptr, size_t bytes)
{
    if (!c || !dptr) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    *dptr = nullptr;
    CU_TRY(cudaMalloc(dptr, bytes ? bytes : 4));
    return SPECK_OK;
}
int speck_b200_free(speck_ctx *c, void *dptr)
{
    if (!c) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    if (dptr) CU_TRY(cudaFree(dptr));
    return SPECK_OK;
}
int speck_b200_memcpy_h2d(speck_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (!c) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    if (bytes) CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->main));
    CU_TRY(cudaStreamSynchronize(c->main));
    return SPECK_OK;
}
int speck_b200_memcpy_d2h(speck_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (!c) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    if (bytes) CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->main));
    CU_TRY(cudaStreamSynchronize(c->main));
    return SPECK_OK;
}
int speck_b200_free_csr(speck_ctx *c, speck_csr *C)
{
    if (!c || !C) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    if (C->data) cudaFree(C->data);
    if (C->col_ids) cudaFree(C->col_ids);
    if (C->row_offsets) cudaFree(C->row_offsets);
    memset(C, 0, sizeof(*C));
    return SPECK_OK;
}
int speck_b200_synchronize(speck_ctx *c)
{
    if (!c) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaDeviceSynchronize());
    return SPECK_OK;
}
void *speck_b200_stream(speck_ctx *c) { return c ? (void *)c->main : nullptr; }

int speck_b200_partition_rows(speck_ctx *c, const speck_csr *A, const speck_csr *B, int parts, uint32_t *cuts,
                              uint64_t *part_products)
{
    if (!c || !A || !B || !cuts) return fail(SPECK_ERR_INVALID, "null argument");
    if (parts < 1 || parts > SPECK_MAX_SHARDS) return fail(SPECK_ERR_INVALID, "parts must be in [1, %d]", SPECK_MAX_SHARDS);
    if (A->cols != B->rows) return fail(SPECK_ERR_INVALID, "shape mismatch: A is %zux%zu, B is %zux%zu", A->rows, A->cols, B->rows, B->cols);
    CU_TRY(cudaSetDevice(c->device));
    const u32 rows = (u32)A->rows;
    cuts[0] = 0;
    for (int g = 1; g <= parts; ++g) cuts[g] = rows;
    if (part_products) memset(part_products, 0, sizeof(uint64_t) * parts);
    if (rows == 0 || A->nnz == 0 || B->nnz == 0) {
        for (int g = 1; g < parts; ++g) cuts[g] = (u32)(((u64)rows * g) / parts);
        return SPECK_OK;
    }
    int rc;
    if ((rc = ensure(c->rowOps, (size_t)(rows + 1) * 4))) return rc;
    if ((rc = ensure(c->perm, (size_t)rows * 4))) return rc;   // scratch for the rowNnz side output
    if ((rc = ensure(c->rowMin, (size_t)rows * 4))) return rc;
    if ((rc = ensure(c->rowMax, (size_t)rows * 4))) return rc;
    if ((rc = ensure(c->mapBase, (size_t)(rows + 1) * 8 + (size_t)(parts + 1) * 4 + (size_t)parts * 8 + 16))) return rc;
    if ((rc = ensure(c->tileState, scan_tile_state_bytes(rows + 1)))) return rc;
    CU_TRY(cudaMemsetAsync(c->dSc, 0, sizeof(Scalars), c->main));
    u32 n = 0;
    LaunchCtx lc{c->main, c->smCount, &n};
    launch_analyze(lc, rows, A->nnz, A->row_offsets, A->col_ids, B->row_offsets, B->col_ids, (u32 *)c->rowOps.p,
                   (u32 *)c->rowMin.p, (u32 *)c->rowMax.p, (u32 *)c->perm.p, c->dSc, c->sortMax, nullptr, nullptr);
    launch_row_cost(lc, rows, A->row_offsets, (u32 *)c->rowOps.p, c->partRowCost, c->partEntryCost);
    u64 *prefix = (u64 *)c->mapBase.p;
    launch_scan_map(lc, (const u32 *)c->rowOps.p, prefix, rows + 1, (u64 *)c->tileState.p, c->dSc);
    u64 *dPart = prefix + rows + 1;
    u32 *dCuts = (u32 *)(dPart + parts);
    launch_find_cuts(lc, prefix, rows, (u32)parts, dCuts, dPart);
    CU_TRY(cudaMemcpyAsync(cuts, dCuts, (size_t)(parts + 1) * 4, cudaMemcpyDeviceToHost, c->main));
    if (part_products) CU_TRY(cudaMemcpyAsync(part_products, dPart, (size_t)parts * 8, cudaMemcpyDeviceToHost, c->main));
    CU_TRY(cudaStreamSynchronize(c->main));
    CU_TRY(cudaGetLastError());
    return SPECK_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// Row-sharded multiply on several devices of one box.
// ------------------------------------------------------------------------------------------------------------
struct speck_shard_plan {
    int n = 0;
    size_t valBytes = 8;
    speck_ctx *ctx[SPECK_MAX_SHARDS] = {};
    speck_csr dA[SPECK_MAX_SHARDS] = {};   // slab of A on device g (row_offsets re-based)
    speck_csr dB[SPECK_MAX_SHARDS] = {};   // B on device g
    speck_csr dC[SPECK_MAX_SHARDS] = {};   // slab of C of the last multiply (reused)
    u32 cuts[SPECK_MAX_SHARDS + 1] = {};
    u64 products[SPECK_MAX_SHARDS] = {};
    float msSetup = 0.f;
    size_t rows = 0, colsB = 0;
};

namespace {

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void free_dev_csr(speck_csr &m)
{
    if (m.data) cudaFree(m.data);
    if (m.col_ids) cudaFree(m.col_ids);
    if (m.row_offsets) cudaFree(m.row_offsets);
    m = speck_csr{};
}

int alloc_dev_csr(speck_csr &m, size_t rows, size_t cols, size_t nnz, size_t valBytes)
{
    m = speck_csr{};
    m.rows = rows; m.cols = cols; m.nnz = nnz;
    CU_TRY(cudaMalloc((void **)&m.row_offsets, (rows + 1) * 4));
    CU_TRY(cudaMalloc((void **)&m.col_ids, (nnz ? nnz : 1) * 4));
    CU_TRY(cudaMalloc(&m.data, (nnz ? nnz : 1) * valBytes));
    return SPECK_OK;
}

template <typename T>
int sharded_create_impl(speck_ctx **ctxs, int n, const speck_csr *A, const speck_csr *B, speck_shard_plan **out)
{
    if (!ctxs || !A || !B || !out) return fail(SPECK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n < 1 || n > SPECK_MAX_SHARDS) return fail(SPECK_ERR_INVALID, "number of devices must be in [1, %d]", SPECK_MAX_SHARDS);
    for (int g = 0; g < n; ++g) {   // several contexts may share a device (that is how the one-GPU tests run the slabs)
        if (!ctxs[g]) return fail(SPECK_ERR_INVALID, "null context %d", g);
        for (int h = 0; h < g; ++h)
            if (ctxs[h] == ctxs[g]) return fail(SPECK_ERR_INVALID, "context %d is listed twice", g);
    }
    if (A->cols != B->rows) return fail(SPECK_ERR_INVALID, "shape mismatch: A is %zux%zu, B is %zux%zu", A->rows, A->cols, B->rows, B->cols);
    if (!A->row_offsets || !B->row_offsets || (A->nnz && (!A->col_ids || !A->data)) || (B->nnz && (!B->col_ids || !B->data)))
        return fail(SPECK_ERR_INVALID, "null CSR array");
    speck_shard_plan *p = new (std::nothrow) speck_shard_plan();
    if (!p) return fail(SPECK_ERR_OOM, "host allocation failed");
    p->n = n;
    p->valBytes = sizeof(T);
    p->rows = A->rows;
    p->colsB = B->cols;
    for (int g = 0; g < n; ++g) p->ctx[g] = ctxs[g];
    const double t0 = now_ms();
    const int rc = [&]() -> int {
        // B and (for the partition) the whole of A on the first device
        speck_ctx *c0 = ctxs[0];
        CU_TRY(cudaSetDevice(c0->device));
        int rc;
        if ((rc = alloc_dev_csr(p->dB[0], B->rows, B->cols, B->nnz, sizeof(T)))) return rc;
        CU_TRY(cudaMemcpyAsync(p->dB[0].row_offsets, B->row_offsets, (B->rows + 1) * 4, cudaMemcpyHostToDevice, c0->main));
        if (B->nnz) {
            CU_TRY(cudaMemcpyAsync(p->dB[0].col_ids, B->col_ids, B->nnz * 4, cudaMemcpyHostToDevice, c0->main));
            CU_TRY(cudaMemcpyAsync(p->dB[0].data, B->data, B->nnz * sizeof(T), cudaMemcpyHostToDevice, c0->main));
        }
        speck_csr fullA;
        if ((rc = alloc_dev_csr(fullA, A->rows, A->cols, A->nnz, 1))) return rc;   // the partition reads indices only
        CU_TRY(cudaMemcpyAsync(fullA.row_offsets, A->row_offsets, (A->rows + 1) * 4, cudaMemcpyHostToDevice, c0->main));
        if (A->nnz) CU_TRY(cudaMemcpyAsync(fullA.col_ids, A->col_ids, A->nnz * 4, cudaMemcpyHostToDevice, c0->main));
        rc = speck_b200_partition_rows(c0, &fullA, &p->dB[0], n, p->cuts, p->products);
        free_dev_csr(fullA);
        if (rc) return rc;
        // B to the peers: device-to-device over NVLink (peer access is switched on when the pair supports it)
        for (int g = 1; g < n; ++g) {
            speck_ctx *cg = ctxs[g];
            CU_TRY(cudaSetDevice(cg->device));
            int can = 0;
            if (cg->device != c0->device) cudaDeviceCanAccessPeer(&can, cg->device, c0->device);
            if (can && cudaDeviceEnablePeerAccess(c0->device, 0) != cudaSuccess) cudaGetLastError();   // already enabled is fine
            if ((rc = alloc_dev_csr(p->dB[g], B->rows, B->cols, B->nnz, sizeof(T)))) return rc;
            CU_TRY(cudaMemcpyPeerAsync(p->dB[g].row_offsets, cg->device, p->dB[0].row_offsets, c0->device, (B->rows + 1) * 4, cg->main));
            if (B->nnz) {
                CU_TRY(cudaMemcpyPeerAsync(p->dB[g].col_ids, cg->device, p->dB[0].col_ids, c0->device, B->nnz * 4, cg->main));
                CU_TRY(cudaMemcpyPeerAsync(p->dB[g].data, cg->device, p->dB[0].data, c0->device, B->nnz * sizeof(T), cg->main));
            }
        }
        // slabs of A: row_offsets re-based on the host, column ids / values are contiguous slices
        std::vector<u32> rp;
        for (int g = 0; g < n; ++g) {
            speck_ctx *cg = ctxs[g];
            CU_TRY(cudaSetDevice(cg->device));
            const u32 r0 = p->cuts[g], r1 = p->cuts[g + 1];
            const u32 e0 = A->row_offsets[r0], e1 = A->row_offsets[r1];
            if ((rc = alloc_dev_csr(p->dA[g], r1 - r0, A->cols, e1 - e0, sizeof(T)))) return rc;
            rp.resize((size_t)(r1 - r0) + 1);
            for (u32 i = 0; i <= r1 - r0; ++i) rp[i] = A->row_offsets[r0 + i] - e0;
            CU_TRY(cudaMemcpyAsync(p->dA[g].row_offsets, rp.data(), rp.size() * 4, cudaMemcpyHostToDevice, cg->main));
            if (e1 > e0) {
                CU_TRY(cudaMemcpyAsync(p->dA[g].col_ids, A->col_ids + e0, (size_t)(e1 - e0) * 4, cudaMemcpyHostToDevice, cg->main));
                CU_TRY(cudaMemcpyAsync(p->dA[g].data, (const T *)A->data + e0, (size_t)(e1 - e0) * sizeof(T), cudaMemcpyHostToDevice, cg->main));
            }
            CU_TRY(cudaStreamSynchronize(cg->main));   // rp is reused by the next slab
        }
        for (int g = 0; g < n; ++g) {
            CU_TRY(cudaSetDevice(ctxs[g]->device));
            CU_TRY(cudaStreamSynchronize(ctxs[g]->main));
        }
        return SPECK_OK;
    }();
    if (rc != SPECK_OK) {
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        speck_b200_sharded_destroy(p);
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    p->msSetup = (float)(now_ms() - t0);
    *out = p;
    return SPECK_OK;
}

}  // namespace

extern "C" {

int speck_b200_sharded_create_f64(speck_ctx **ctxs, int n, const speck_csr *A, const speck_csr *B, speck_shard_plan **plan)
{
    return sharded_create_impl<double>(ctxs, n, A, B, plan);
}
int speck_b200_sharded_create_f32(speck_ctx **ctxs, int n, const speck_csr *A, const speck_csr *B, speck_shard_plan **plan)
{
    return sharded_create_impl<float>(ctxs, n, A, B, plan);
}

int speck_b200_sharded_multiply(speck_shard_plan *p, speck_shard_info *info)
{
    if (!p) return fail(SPECK_ERR_INVALID, "null plan");
    int rcs[SPECK_MAX_SHARDS] = {};
    std::string errs[SPECK_MAX_SHARDS];
    const double t0 = now_ms();
    auto work = [&](int g) {
        rcs[g] = p->valBytes == 4 ? spgemm_impl<float>(p->ctx[g], &p->dA[g], &p->dB[g], &p->dC[g], nullptr)
                                  : spgemm_impl<double>(p->ctx[g], &p->dA[g], &p->dB[g], &p->dC[g], nullptr);
        if (rcs[g]) errs[g] = g_err;   // the message lives in the worker thread's buffer
    };
    std::vector<std::thread> th;
    for (int g = 1; g < p->n; ++g) th.emplace_back(work, g);   // one host thread per device, the caller's thread drives device 0
    work(0);
    for (auto &t : th) t.join();
    const double t1 = now_ms();
    for (int g = 0; g < p->n; ++g)
        if (rcs[g]) return fail(rcs[g], "shard %d (device %d): %s", g, p->ctx[g]->device, errs[g].c_str());
    if (info) {
        *info = speck_shard_info{};
        info->shards = p->n;
        for (int g = 0; g <= p->n; ++g) info->cuts[g] = p->cuts[g];
        for (int g = 0; g < p->n; ++g) {
            cudaSetDevice(p->ctx[g]->device);
            finish_stage_times(p->ctx[g]);
            info->products[g] = p->ctx[g]->stats.products;
            info->nnz_c[g] = p->dC[g].nnz;
            info->ms_device[g] = p->ctx[g]->stats.ms_total;
        }
        info->ms_setup = p->msSetup;
        info->ms_multiply = (float)(t1 - t0);
    }
    return SPECK_OK;
}

int speck_b200_sharded_concat(speck_shard_plan *p, speck_csr *C, speck_shard_info *info)
{
    if (!p || !C) return fail(SPECK_ERR_INVALID, "null argument");
    const double t0 = now_ms();
    speck_ctx *c0 = p->ctx[0];
    u64 total = 0;
    for (int g = 0; g < p->n; ++g) total += p->dC[g].nnz;
    if (total > 0xffffffffull)
        return fail(SPECK_ERR_OVERFLOW, "concatenated nnz(C) = %llu does not fit the u32 row_offsets of the spECK API: keep C distributed",
                    (unsigned long long)total);
    CU_TRY(cudaSetDevice(c0->device));
    if (!(C->rows == p->rows && C->row_offsets)) {   // same reuse rules as spgemm_impl
        if (C->row_offsets) cudaFree(C->row_offsets);
        C->row_offsets = nullptr;
        CU_TRY(cudaMalloc((void **)&C->row_offsets, (p->rows + 1) * 4));
        C->rows = p->rows;
    }
    if (C->nnz != total || !C->data || !C->col_ids) {
        if (C->data) cudaFree(C->data);
        if (C->col_ids) cudaFree(C->col_ids);
        C->data = nullptr; C->col_ids = nullptr; C->nnz = 0;
        CU_TRY(cudaMalloc(&C->data, (total ? total : 1) * p->valBytes));
        CU_TRY(cudaMalloc((void **)&C->col_ids, (total ? total : 1) * 4));
    }
    C->nnz = total;
    C->cols = p->colsB;
    int rc;
    u32 maxRows = 0;
    for (int g = 0; g < p->n; ++g) maxRows = max(maxRows, p->cuts[g + 1] - p->cuts[g]);
    if ((rc = ensure(c0->stage[0], (size_t)(maxRows + 1) * 4))) return rc;   // landing buffer of a peer's row_offsets
    u32 launches = 0;
    LaunchCtx lc{c0->main, c0->smCount, &launches};
    u64 base = 0;
    for (int g = 0; g < p->n; ++g) {
        const speck_csr &s = p->dC[g];
        const u32 r0 = p->cuts[g], nr = p->cuts[g + 1] - p->cuts[g];
        const int dev = p->ctx[g]->device;
        if (s.nnz) {
            CU_TRY(cudaMemcpyPeerAsync(C->col_ids + base, c0->device, s.col_ids, dev, s.nnz * 4, c0->main));
            CU_TRY(cudaMemcpyPeerAsync((char *)C->data + base * p->valBytes, c0->device, s.data, dev, s.nnz * p->valBytes, c0->main));
        }
        if (s.nnz && s.row_offsets) {
            const u32 *src = s.row_offsets;
            if (g > 0) {
                CU_TRY(cudaMemcpyPeerAsync(c0->stage[0].p, c0->device, s.row_offsets, dev, (size_t)(nr + 1) * 4, c0->main));
                src = (const u32 *)c0->stage[0].p;
            }
            launch_offset_rows(lc, src, nr + (g == p->n - 1 ? 1u : 0u), (u32)base, C->row_offsets + r0);
        } else {   // empty slab: every row starts (and the slab ends) at base
            std::vector<u32> fill((size_t)nr + 1, (u32)base);
            CU_TRY(cudaMemcpyAsync(C->row_offsets + r0, fill.data(), fill.size() * 4, cudaMemcpyHostToDevice, c0->main));
            CU_TRY(cudaStreamSynchronize(c0->main));
        }
        base += s.nnz;
    }
    // the last slab wrote row_offsets[rows] = total only when it was not empty
    const u32 tot32 = (u32)total;
    CU_TRY(cudaMemcpyAsync(C->row_offsets + p->rows, &tot32, 4, cudaMemcpyHostToDevice, c0->main));
    CU_TRY(cudaStreamSynchronize(c0->main));
    CU_TRY(cudaGetLastError());
    if (info) {
        info->concatenated = 1;
        info->ms_concat = (float)(now_ms() - t0);
    }
    return SPECK_OK;
}

// ---- one process per GPU: CUDA IPC handles of the concatenated C and the push of a slab into it
int speck_b200_ipc_export(speck_ctx *c, const void *dptr, unsigned char handle[SPECK_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == SPECK_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    if (!c || !dptr || !handle) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, const_cast<void *>(dptr)));
    memcpy(handle, &h, sizeof(h));
    return SPECK_OK;
}

int speck_b200_ipc_open(speck_ctx *c, const unsigned char handle[SPECK_IPC_HANDLE_BYTES], void **dptr)
{
    if (!c || !dptr || !handle) return fail(SPECK_ERR_INVALID, "null argument");
    CU_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    *dptr = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));   // maps the peer allocation, enables peer access
    return SPECK_OK;
}

int speck_b200_ipc_close(speck_ctx *c, void *dptr)
{
    if (!c) return fail(SPECK_ERR_INVALID, "null argument");
    if (!dptr) return SPECK_OK;
    CU_TRY(cudaSetDevice(c->device));
    CU_TRY(cudaIpcCloseMemHandle(dptr));
    return SPECK_OK;
}

int speck_b200_push_slab_f64(speck_ctx *c, const speck_csr *S, uint64_t nnzBase, uint32_t rowBase, int last, uint32_t *dstRp,
                             uint32_t *dstCi, double *dstV, float *deviceMs)
{
    return push_slab_impl<double>(c, S, nnzBase, rowBase, last, dstRp, dstCi, dstV, deviceMs);
}

int speck_b200_push_slab_f32(speck_ctx *c, const speck_csr *S, uint64_t nnzBase, uint32_t rowBase, int last, uint32_t *dstRp,
                             uint32_t *dstCi, float *dstV, float *deviceMs)
{
    return push_slab_impl<float>(c, S, nnzBase, rowBase, last, dstRp, dstCi, dstV, deviceMs);
}

int speck_b200_sharded_slab(speck_shard_plan *p, int g, speck_csr *A_slab, speck_csr *C_slab)
{
    if (!p || g < 0 || g >= p->n) return fail(SPECK_ERR_INVALID, "bad plan or slab index");
    if (A_slab) *A_slab = p->dA[g];
    if (C_slab) *C_slab = p->dC[g];
    return SPECK_OK;
}

int speck_b200_sharded_destroy(speck_shard_plan *p)
{
    if (!p) return SPECK_OK;
    for (int g = 0; g < p->n; ++g) {
        if (!p->ctx[g]) continue;
        cudaSetDevice(p->ctx[g]->device);
        cudaDeviceSynchronize();
        free_dev_csr(p->dA[g]);
        free_dev_csr(p->dB[g]);
        free_dev_csr(p->dC[g]);
    }
    delete p;
    return SPECK_OK;
}

int speck_b200_set_option(speck_ctx *c, const char *key, long long value)
{
    if (!c || !key) return fail(SPECK_ERR_INVALID, "null argument");
    if (!strcmp(key, "sort_max")) {
        if (value < 4 || value > (long long)RANK_MAX_PRODUCTS || ((value & (value - 1)) && (value % 512 || value > SORT_MAX_PRODUCTS)))
            return fail(SPECK_ERR_INVALID, "sort_max must be a power of two or a multiple of 512 (up to %u) in [4, %u]",
                        SORT_MAX_PRODUCTS, RANK_MAX_PRODUCTS);
        c->sortMax = (u32)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "sym_streams") || !strcmp(key, "num_streams")) {
        if (value < (key[0] == 's' ? 1 : 0) || value > NSIDE) return fail(SPECK_ERR_INVALID, "%s must be in [1, %d]", key, NSIDE);
        (key[0] == 's' ? c->symStreams : c->numStreams) = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "rank_map_max_bytes")) {
        c->rankMapMaxBytes = value < 0 ? ~(size_t)0 : (size_t)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "map_min_class")) {
        if (value < 0 || value > NUM_WARP_SORT) return fail(SPECK_ERR_INVALID, "map_min_class must be in [0, %d]", NUM_WARP_SORT);
        c->mapMinClass = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "map_cta_min")) {
        if (value < 4 || value > NUM_WARP_SORT) return fail(SPECK_ERR_INVALID, "map_cta_min must be in [4, %d]", NUM_WARP_SORT);
        c->mapCtaMin = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "rank_map")) {
        c->rankMapOn = value != 0;
        return SPECK_OK;
    }
    if (!strcmp(key, "partition_row_cost") || !strcmp(key, "partition_entry_cost")) {
        if (value < 0 || value > 1024) return fail(SPECK_ERR_INVALID, "%s must be in [0, 1024]", key);
        (key[10] == 'r' ? c->partRowCost : c->partEntryCost) = (u32)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "tiered_analysis")) {
        if (value < -1 || value > 1) return fail(SPECK_ERR_INVALID, "tiered_analysis must be -1, 0 or 1");
        c->tieredAnalysis = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "deterministic")) {
        c->deterministic = value != 0;
        return SPECK_OK;
    }
    if (!strcmp(key, "test_set")) {
        c->testSet = value != 0;
        return SPECK_OK;
    }
    if (!strcmp(key, "dense_seq")) {
        if (value < 0 || value > 2) return fail(SPECK_ERR_INVALID, "dense_seq must be 0, 1 or 2");
        c->denseSeq = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "col_direct")) {
        if (value < 0 || value > 3) return fail(SPECK_ERR_INVALID, "col_direct must be in [0, 3]");
        c->colDirect = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "flat_min_class")) {
        if (value < 6 || value > NUM_WARP_SORT) return fail(SPECK_ERR_INVALID, "flat_min_class must be in [6, %d]", NUM_WARP_SORT);
        c->flatMinClass = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "narrow_groups")) {
        if (value < 0 || value > 3) return fail(SPECK_ERR_INVALID, "narrow_groups must be in [0, 3]");
        c->narrow = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "narrow_numeric")) {
        if (value < -1 || value > 3) return fail(SPECK_ERR_INVALID, "narrow_numeric must be in [-1, 3]");
        c->narrowNum = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "sym_mix")) {
        if (value < 0 || value > 16) return fail(SPECK_ERR_INVALID, "sym_mix must be in [0, 16]");
        c->symMix = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "num_plan")) {
        if (value < 0 || value > 1) return fail(SPECK_ERR_INVALID, "num_plan must be 0 or 1");
        c->numPlan = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "big_split")) {
        if (value < 0 || value > 3) return fail(SPECK_ERR_INVALID, "big_split must be in [0, 3]");
        c->bigSplit = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "spin_wait")) {
        c->spinWait = value != 0;
        return SPECK_OK;
    }
    if (!strcmp(key, "seg_num")) {
        c->segNum = value != 0;
        return SPECK_OK;
    }
    if (!strcmp(key, "flat_sym")) {
        c->flatSym = value != 0;
        return SPECK_OK;
    }
    if (!strcmp(key, "flat_e")) {
        if (value != 8 && value != 16) return fail(SPECK_ERR_INVALID, "flat_e must be 8 or 16");
        c->flatE = (int)value;
        return SPECK_OK;
    }
    if (!strcmp(key, "hash_count")) {
        c->hashCount = value != 0;
        return SPECK_OK;
    }
    if (!strcmp(key, "rank_path")) {
        c->rankPath = value != 0;
        return SPECK_OK;
    }
    if (!strcmp(key, "release_workspace")) {
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        release(c->rowOps); release(c->perm); release(c->rowMin); release(c->rowMax); release(c->tileState); release(c->bitmapStore); release(c->mapLen); release(c->mapBase); release(c->rankMap); release(c->aSeg); release(c->aOff); release(c->desc); release(c->rowInfo);
        for (auto &b : c->stage) release(b);
        if (c->hostC.data) cudaFree(c->hostC.data);       // device C kept by the *_host entry points
        if (c->hostC.col_ids) cudaFree(c->hostC.col_ids);
        if (c->hostC.row_offsets) cudaFree(c->hostC.row_offsets);
        c->hostC = speck_csr{};
        return SPECK_OK;
    }
    return fail(SPECK_ERR_INVALID, "unknown option '%s'", key);
}

}  // extern "C"
