// speck_b200/csrc/common.cuh -- shared definitions of the sm_100a SpGEMM kernels.
//
// Pipeline (replaces source/GPU/Multiply.cu:99-1122 of the reference):
//   B row summary -> analyze rows -> bin rows by product count -> rank-map offsets + row descriptors ->
//   symbolic (exact nnz per row + the rank map: sorted position of every product) ->
//   decoupled look-back scan (row_ptr) -> numeric (gather, multiply, scatter by rank, coalesced row writes).
// Row classes ("bins"):
//   BIN_DIRECT        A row has exactly one entry: C row = scaled copy of one B row
//                     (reference: directSpGEMM*, spECK_HashSpGEMM.cuh:543-589)
//   BIN_SORT0 + c     c = 0..7: products <= 4<<c, one lane group per row, register bitonic sort (sort_rows.cuh);
//                     c = 8..23: 513..16384 products, one CTA per row: two/three-level column bitmap ranks
//                     (rank_cta.cuh); CTA bitonic sort (sort_cta.cuh) only as the fallback for matrices wider
//                     than the rank kernels take.  Mapped rows: numeric = k_map_rows / k_map_rows_cta
//                     (replaces the smem hash + O(n^2) rank sort / radix sort, :591-866, :1856-1925)
//   BIN_DENSE(_LOCAL) banded rows and rows with more products: CTA per row, sparse-cleared column bitmap in
//                     shared memory, popcount ranks give the sorted position of every product, values
//                     accumulate in shared memory or with fp RED into C
//                     (replaces denseSpGEMM{Count,Numeric}, :1300-1711)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef uint32_t u32;
typedef uint64_t u64;

namespace sb {

constexpr int NUM_WARP_SORT = 8;            // lane-group classes: 4, 8, ..., 512 products
constexpr int NUM_CTA_SORT = 16;            // CTA classes: 1024, 1536, ..., 8192 products (steps of 512) and 16384
constexpr int NUM_SORT = NUM_WARP_SORT + NUM_CTA_SORT;
constexpr int BIN_DIRECT = 0;
constexpr int BIN_SORT0 = 1;
constexpr int BIN_DENSE_LOCAL = BIN_SORT0 + NUM_SORT;   // 25: bitmap path, column extent <= DENSE_LOCAL_COLS
constexpr int BIN_DENSE = BIN_DENSE_LOCAL + 1;          // 26: bitmap path, wide rows
constexpr int NUM_BINS = BIN_DENSE + 1;                 // 27
constexpr int DENSE_LOCAL_BITS = 14;
constexpr u32 DENSE_LOCAL_COLS = (1u << DENSE_LOCAL_BITS) - 128u;  // window starts chunk-aligned below the row minimum
constexpr u32 SORT_MAX_PRODUCTS = 8192;     // largest row of the CTA sort kernels (sort_cta.cuh)
constexpr u32 RANK_MAX_PRODUCTS = 16384;    // largest row of the rank kernels (rank_cta.cuh): 1024 threads x 16 slots

// Device-resident scalars of one multiply; mirrored into pinned host memory.
struct Scalars {
    u64 products;             // P (u64: the reference's u32 sumProducts overflows, Multiply.cu:237-252)
    u64 nnzC;                 // total of the row_ptr scan
    u32 maxRowProducts;
    u32 binCount[NUM_BINS];   // rows per bin (written by k_analyze)
    u32 binCursor[NUM_BINS];  // scatter cursors (k_bin_scatter)
    u32 denseCounter[6];      // dynamic row queues of the dense kernels (symbolic/numeric x local/wide, + the
                              // largest rank bin when no rank map could be allocated)
    u32 tileCounter;          // dynamic tile ids of the row_ptr scan
    u32 mapTileCounter;       // dynamic tile ids of the rank-map offset scan
    u64 mapTotal;             // entries of the rank map (products of all mapped rows)
    u32 longCount;            // hub rows of A queued by k_analyze for k_analyze_long
    u32 seqRows[2];           // local bitmap rows the sequential-k numeric kernel takes (<= 512 / <= 2048 entries), counted
                              // by the symbolic kernel so that the host launches those kernels only when needed
    u32 compareFlag;          // k_compare: 0 = equal
    u64 compareFirst;         // k_compare: smallest (row << 32 | kind << 28 | position in row) that differs; kind 0 = row
                              // length, 1 = column id, 2 = value (first-mismatch report)
};

// Row descriptor of the mapped classes, one per binned row in perm order (built once per multiply): a CTA /
// lane group reads its row's parameters with two 16-byte loads instead of chasing perm -> row -> five arrays.
struct __align__(16) RowDesc {
    u32 aBeg, aLen;   // A entries of the row
    u32 n;            // products of the row
    u32 row;
    u32 c0;           // symbolic: smallest column of the row; numeric: first position of the row in C
    u32 c1;           // symbolic: largest column;              numeric: entries of the row in C
    u64 mapOff;       // first rank-map entry of the row
};

struct CsrView {
    const u32 *rp;
    const u32 *ci;
    const void *v;
};

// bin of a row; -1 = no products at all.  extent = maxCol - minCol + 1 of the row's products.
// Rows whose column extent is at most four times their product count compress or are nearly dense: they take the bitmap path, which does not pay for the duplicates the way sorting does.
__host__ __device__ __forceinline__ int classify_row(u32 ops, u32 aLen, u32 extent, u32 sortMax)
{
    if (ops == 0) return -1;
    if (aLen == 1) return BIN_DIRECT;
    const bool banded = ops >= 128u && extent <= 4u * ops;
    if (ops > sortMax || banded) return extent <= DENSE_LOCAL_COLS ? BIN_DENSE_LOCAL : BIN_DENSE;
    if (ops > SORT_MAX_PRODUCTS) return BIN_SORT0 + NUM_SORT - 1;                       // 8193..16384 (rank kernels only)
    if (ops > 512u) return BIN_SORT0 + NUM_WARP_SORT + (int)((ops + 511u) / 512u) - 2;  // 2..16 warps
    int c = 0;
    while ((4u << c) < ops) ++c;
    return BIN_SORT0 + c;
}

// ---------------------------------------------------------------- launch API (host)
struct LaunchCtx {
    cudaStream_t stream;
    int smCount;
    u32 *launches;   // incremented per kernel launch
    int narrow = 0;  // lane-group classes of <= 64 (1) / <= 128 (2) products: fewer lanes per row, more products per lane
    u32 gridCap = 0; // lane-group sort kernels: at most this many CTAs, each looping over row groups (0: one CTA per group
                     // of rows); lets a long instruction-bound kernel share the SMs with kernels of other streams
};

// aOff (optional, nnz(A) entries): index of every A entry's first product in its row's flat product enumeration
// (exclusive prefix of the B-row lengths inside the row), so that the segment-major numeric kernel needs no scan.
// aSeg (optional, nnz(A) entries): (begin, end) of the B row referenced by every A entry, so that the row
// kernels need one load level (aSeg) instead of two (A.col_ids -> B.row_offsets)
void launch_analyze(const LaunchCtx &lc, u32 rows, u64 nnzA, const u32 *aRp, const u32 *aCi, const u32 *bRp,
                    const u32 *bCi, u32 *rowOps, u32 *rowMin, u32 *rowMax, u32 *rowNnz, Scalars *sc, u32 sortMax,
                    uint2 *aSeg, const uint4 *rowInfo, u32 *aOff = nullptr, u32 *mapLen = nullptr, bool mapCta = false,
                    int mapMinClass = 0, u32 extentMinOps = 0 /* > 0 and no rowInfo: column extents only for rows with at
                    least this many products (two-pass analysis for matrices whose B summaries would miss the L2) */,
                    u32 *longRows = nullptr /* scratch of `rows` entries: queue of the hub rows of A (>= 1024 entries), which
                    one CTA each analyses afterwards */);
// rowInfo[k] = (begin, end, first column, last column) of B row k: one gather per A entry in the analysis
void launch_row_info(const LaunchCtx &lc, u32 rowsB, const u32 *bRp, const u32 *bCi, uint4 *rowInfo);
// descriptors (written by launch_bin_scatter, symbolic flavour: c0/c1 = column extent) switched to the numeric
// flavour (c0/c1 = position / length in C) once row_offsets are scanned
void launch_desc_numeric(const LaunchCtx &lc, u32 count, const u32 *cRp, RowDesc *desc);
// mapLen (launch_analyze; optional, rows + 1 entries): products of the row when its class records a rank map in
// the symbolic phase (lane-group classes from mapMinClass on, CTA classes when mapCta), else 0.
// launch_bin_scatter: one permutation ordered by bin and, when desc != nullptr, the row descriptors in that order
// (mapBase = exclusive scan of mapLen)
void launch_bin_scatter(const LaunchCtx &lc, u32 rows, const u32 *aRp, const u32 *rowOps, const u32 *rowMin,
                        const u32 *rowMax, u32 *perm, Scalars *sc, u32 sortMax, const u64 *mapBase, RowDesc *desc);
void launch_scan(const LaunchCtx &lc, u32 *data, u32 n /* entries incl. the trailing total slot */,
                 u64 *tileState, Scalars *sc);
// exclusive scan of in[0..n-1) into 64-bit out[0..n) (out[n-1] = total, also stored in sc->mapTotal)
void launch_scan_map(const LaunchCtx &lc, const u32 *in, u64 *out, u32 n, u64 *tileState, Scalars *sc);

// device scalars -> mapped pinned host mirror + sequence word (polled by the host)
void launch_publish(const LaunchCtx &lc, const Scalars *dSc, Scalars *hSc, volatile u32 *hSeq, u32 seq);

// multi-GPU helpers: product-balanced row cuts from the u64 prefix of rowOps (prefix[rows] = P); row_offsets of a
// slab shifted by the slab's first position in the concatenated C
void launch_row_cost(const LaunchCtx &lc, u32 rows, const u32 *aRp, u32 *rowOps, u32 wRow, u32 wEntry);
void launch_find_cuts(const LaunchCtx &lc, const u64 *prefix, u32 rows, u32 parts, u32 *cuts, u64 *partProducts);
void launch_offset_rows(const LaunchCtx &lc, const u32 *in, u32 n, u32 base, u32 *out);
// one slab of C -> its place in the concatenated C (possibly peer memory): columns and values at nnzBase, row offsets
// with nnzBase added at rowBase; 16-byte stores
template <typename T>
void launch_push_slab(const LaunchCtx &lc, const u32 *srcRp, const u32 *srcCi, const T *srcV, u32 rowsOut, u64 nnz,
                      u64 nnzBase, u32 rowBase, u32 *dstRp, u32 *dstCi, T *dstV);

// Rank map: one u16 per product of a mapped row, written by the symbolic phase at mapBase[row] + (index of the
// product in the row's flat enumeration: A entries ascending, then B-row order):
//   bits 0..14 = position of the product's column in the sorted row of C, bit 15 = not the first product of
//   that column.  The numeric phase then only gathers, multiplies and scatters by rank.
constexpr u32 MAP_DUP = 0x8000u;
constexpr u32 MAP_RANK_MASK = 0x7fffu;
size_t scan_tile_state_bytes(u32 n);

// mapBase / rankMap == nullptr: count only; wideKeys as in launch_sort_numeric (needed for the map only)
void launch_sort_symbolic(const LaunchCtx &lc, int sortClass, bool wideKeys, const u32 *perm, u32 count,
                          const u32 *aRp, const u32 *aCi, const u32 *bRp, const u32 *bCi, const u32 *rowOps,
                          u32 *rowNnz, const RowDesc *desc, const uint2 *aSeg, unsigned short *rankMap);
// numeric phase of mapped rows: lane-group classes (sortClass < NUM_WARP_SORT) ...
template <typename T>
void launch_map_numeric(const LaunchCtx &lc, int sortClass, const RowDesc *desc, u32 count, const uint2 *aSeg,
                        const T *aV, const u32 *bCi, const T *bV, const unsigned short *rankMap, u32 *cCi, T *cV);
// ... and CTA classes (rows of <= capProducts products)
template <typename T>
void launch_map_numeric_cta(const LaunchCtx &lc, u32 capProducts, const RowDesc *desc, u32 count, const uint2 *aSeg,
                            const T *aV, const u32 *bCi, const T *bV, const unsigned short *rankMap, u32 *cCi,
                            T *cV, int colDirect = 0 /* 1: the 1024-thread shapes write column ids straight to C and
                            stage only the values (two CTAs per SM), 2: the 512-thread shape as well */,
                            int bigSplit = 0 /* rows of 4097 .. 16384 products: several CTAs per row (map_split.cuh) */);
template <typename T>
void launch_sort_numeric(const LaunchCtx &lc, int sortClass, bool wideKeys, const u32 *perm, u32 count,
                         const u32 *aRp, const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi,
                         const T *bV, const u32 *rowOps, const u32 *cRp, u32 *cCi, T *cV);
template <typename T>
void launch_direct_numeric(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi,
                           const T *aV, const u32 *bRp, const u32 *bCi, const T *bV, const u32 *cRp,
                           u32 *cCi, T *cV);

// rank classes (rank_cta.cuh): the CTA sort classes' rows when cols(B) <= RANK_EXTENT_LIMIT (two bitmap levels)
// or <= RANK_EXTENT_LIMIT3 (three levels; symbolic kernels only, i.e. together with the rank map)
constexpr u32 RANK_EXTENT_LIMIT = 1u << 20;
constexpr u32 RANK_EXTENT_LIMIT3 = 1u << 25;
void launch_rank_symbolic(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                          const u32 *aCi, const u32 *bRp, const u32 *bCi, const u32 *rowOps, const u32 *rowMin,
                          const u32 *rowMax, u32 *rowNnz, const RowDesc *desc, const uint2 *aSeg,
                          unsigned short *rankMap, int levels /* 2: cols(B) <= 2^20, 3: <= 2^25 */);
// flat (staged) variant of the mapped two-level symbolic kernel (rank_flat.cuh); perThread = product slots per
// thread (8 or 16)
void launch_rank_flat(const LaunchCtx &lc, u32 capProducts, int perThread, const RowDesc *desc, u32 count,
                      const uint2 *aSeg, const u32 *bCi, unsigned short *rankMap, u32 *cRp);
// segment-major numeric kernel of the mapped CTA classes (map_seg.cuh); aOff from launch_analyze
template <typename T>
void launch_map_seg(const LaunchCtx &lc, u32 capProducts, const RowDesc *desc, u32 count, const uint2 *aSeg,
                    const u32 *aOff, const T *aV, const u32 *bCi, const T *bV, const unsigned short *rankMap,
                    u32 *cCi, T *cV);
// count-only symbolic by shared-memory hashing (experiment)
void launch_hash_count(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                       const u32 *aCi, const u32 *bRp, const u32 *bCi, u32 *cRp);
template <typename T>
void launch_rank_numeric(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                         const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi, const T *bV,
                         const u32 *rowOps, const u32 *rowMin, const u32 *rowMax, const u32 *cRp, u32 *cCi, T *cV);

// dense (bitmap) path; winBits = log2 of the column window held in shared memory
int dense_window_bits(u64 colsB);
void launch_dense_symbolic(const LaunchCtx &lc, bool local, const u32 *perm, u32 count, u32 *rowCounter,
                           const u32 *aRp, const u32 *aCi, const u32 *bRp, const u32 *bCi, u32 colsB,
                           const u32 *rowMin, const u32 *rowMax, u32 *bitmapStore, u32 *rowNnz,
                           const u32 *rowOps = nullptr, u32 *seqRows = nullptr /* Scalars::seqRows, counted for local rows */,
                           bool everyRow = false /* deterministic mode: count every row that fits, whatever its fold */,
                           bool testSet = false /* local rows: read the bitmap word before the atomicOr */);
size_t dense_local_store_bytes(u32 count);
template <typename T>
void launch_dense_numeric(const LaunchCtx &lc, bool local, const u32 *perm, u32 count, u32 *rowCounter,
                          const u32 *aRp, const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi, const T *bV,
                          u32 colsB, const u32 *rowMin, const u32 *rowMax, u32 *bitmapStore, const u32 *cRp,
                          u32 *cCi, T *cV,
                          int seq = 0 /* local rows, sequential-k kernel (dense_seq.cuh): 1 = lane loads, 2 = TMA-staged B segments;
                                         +4: rows of <= 512 entries exist, +8: rows of 513..2048 entries exist,
                                         +16: deterministic mode (every row that fits the accumulator) */,
                          const u32 *rowOps = nullptr /* products per row: the sequential-k kernel takes the rows that fold */);
// deterministic mode: values of the rows of a bitmap bin recomputed in sequential ascending-k order (dense_seq.cuh: k_det_rows);
// skipSeqRows: rows the sequential-k kernel wrote are left alone
template <typename T>
void launch_det_rows(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi, const T *aV,
                     const u32 *bRp, const u32 *bCi, const T *bV, const u32 *cRp, const u32 *cCi, T *cV,
                     const u32 *rowOps, bool skipSeqRows);
constexpr int DENSE_SEQ_MAX = 2048;   // distinct columns of a row the sequential-k kernel accumulates in shared memory

template <typename T>
void launch_compare(const LaunchCtx &lc, u32 rows, const u32 *rpA, const u32 *ciA, const T *vA, const u32 *rpB,
                    const u32 *ciB, const T *vB, bool compareData, double relTol, Scalars *sc);

}  // namespace sb
