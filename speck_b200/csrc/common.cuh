// speck_b200/csrc/common.cuh -- shared definitions of the sm_100a SpGEMM kernels.
//
// Pipeline (replaces source/GPU/Multiply.cu:99-1122 of the reference):
//   analyze rows -> bin rows by product count -> symbolic (exact nnz per row) ->
//   decoupled look-back scan (row_ptr) -> numeric (sorted col/val assembly).
// Row classes ("bins"):
//   BIN_DIRECT        A row has exactly one entry: C row = scaled copy of one B row
//                     (reference: directSpGEMM*, spECK_HashSpGEMM.cuh:543-589)
//   BIN_SORT0 + c     products <= 4<<c (c = 0..7): register bitonic sort of the row's products by a
//                     lane group; c = 8..22: by a CTA of 2..16 warps (512 keys per warp),
//                     duplicates folded after the sort
//                     (replaces the smem hash + O(n^2) rank sort, :591-866)
//   BIN_DENSE         more products: CTA per row, sparse-cleared column bitmap in
//                     shared memory, popcount ranks give the sorted position of every
//                     product, values accumulate with fp RED into C
//                     (replaces denseSpGEMM{Count,Numeric}, :1300-1711)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef uint32_t u32;
typedef uint64_t u64;

namespace sb {

constexpr int NUM_WARP_SORT = 8;            // lane-group classes: 4, 8, ..., 512 products
constexpr int NUM_CTA_SORT = 15;            // CTA classes: 2..16 warps x 512 keys = 1024, 1536, ..., 8192 products
constexpr int NUM_SORT = NUM_WARP_SORT + NUM_CTA_SORT;
constexpr int BIN_DIRECT = 0;
constexpr int BIN_SORT0 = 1;
constexpr int BIN_DENSE_LOCAL = BIN_SORT0 + NUM_SORT;   // 24: bitmap path, column extent <= DENSE_LOCAL_COLS
constexpr int BIN_DENSE = BIN_DENSE_LOCAL + 1;          // 25: bitmap path, wide rows
constexpr int NUM_BINS = BIN_DENSE + 1;                 // 26
constexpr int DENSE_LOCAL_BITS = 14;
constexpr u32 DENSE_LOCAL_COLS = (1u << DENSE_LOCAL_BITS) - 128u;  // window starts chunk-aligned below the row minimum
constexpr u32 SORT_MAX_PRODUCTS = 8192;

// Device-resident scalars of one multiply; mirrored into pinned host memory.
struct Scalars {
    u64 products;             // P (u64: the reference's u32 sumProducts overflows, Multiply.cu:237-252)
    u64 nnzC;                 // total of the row_ptr scan
    u32 maxRowProducts;
    u32 binCount[NUM_BINS];   // rows per bin (written by k_analyze)
    u32 binCursor[NUM_BINS];  // scatter cursors (k_bin_scatter)
    u32 denseCounter[4];      // dynamic row queues of the dense kernels (symbolic/numeric x local/wide)
    u32 tileCounter;          // dynamic tile ids of the scan
    u32 compareFlag;          // k_compare: 0 = equal
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
};

void launch_analyze(const LaunchCtx &lc, u32 rows, u64 nnzA, const u32 *aRp, const u32 *aCi, const u32 *bRp,
                    const u32 *bCi, u32 *rowOps, u32 *rowMin, u32 *rowMax, u32 *rowNnz, Scalars *sc, u32 sortMax);
void launch_bin_scatter(const LaunchCtx &lc, u32 rows, const u32 *aRp, const u32 *rowOps, const u32 *rowMin,
                        const u32 *rowMax, u32 *perm, Scalars *sc, u32 sortMax);
void launch_scan(const LaunchCtx &lc, u32 *data, u32 n /* entries incl. the trailing total slot */,
                 u64 *tileState, Scalars *sc);
size_t scan_tile_state_bytes(u32 n);

void launch_sort_symbolic(const LaunchCtx &lc, int sortClass, const u32 *perm, u32 count, const u32 *aRp,
                          const u32 *aCi, const u32 *bRp, const u32 *bCi, const u32 *rowOps, u32 *rowNnz);
template <typename T>
void launch_sort_numeric(const LaunchCtx &lc, int sortClass, bool wideKeys, const u32 *perm, u32 count,
                         const u32 *aRp, const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi,
                         const T *bV, const u32 *rowOps, const u32 *cRp, u32 *cCi, T *cV);
template <typename T>
void launch_direct_numeric(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi,
                           const T *aV, const u32 *bRp, const u32 *bCi, const T *bV, const u32 *cRp,
                           u32 *cCi, T *cV);

// rank classes (rank_cta.cuh): the CTA sort classes' rows when cols(B) <= RANK_EXTENT_LIMIT
constexpr u32 RANK_EXTENT_LIMIT = 1u << 20;
void set_rank_slots(int e);   // experiment switch: 4 or 8 product slots per thread
void launch_rank_symbolic(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                          const u32 *aCi, const u32 *bRp, const u32 *bCi, const u32 *rowOps, const u32 *rowMin,
                          const u32 *rowMax, u32 *rowNnz);
template <typename T>
void launch_rank_numeric(const LaunchCtx &lc, u32 capProducts, const u32 *perm, u32 count, const u32 *aRp,
                         const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi, const T *bV,
                         const u32 *rowOps, const u32 *rowMin, const u32 *rowMax, const u32 *cRp, u32 *cCi, T *cV);

// dense (bitmap) path; winBits = log2 of the column window held in shared memory
int dense_window_bits(u64 colsB);
void launch_dense_symbolic(const LaunchCtx &lc, bool local, const u32 *perm, u32 count, u32 *rowCounter,
                           const u32 *aRp, const u32 *aCi, const u32 *bRp, const u32 *bCi, u32 colsB,
                           const u32 *rowMin, const u32 *rowMax, u32 *bitmapStore, u32 *rowNnz);
size_t dense_local_store_bytes(u32 count);
template <typename T>
void launch_dense_numeric(const LaunchCtx &lc, bool local, const u32 *perm, u32 count, u32 *rowCounter,
                          const u32 *aRp, const u32 *aCi, const T *aV, const u32 *bRp, const u32 *bCi, const T *bV,
                          u32 colsB, const u32 *rowMin, const u32 *rowMax, u32 *bitmapStore, const u32 *cRp,
                          u32 *cCi, T *cV);

template <typename T>
void launch_compare(const LaunchCtx &lc, u32 rows, const u32 *rpA, const u32 *ciA, const T *vA, const u32 *rpB,
                    const u32 *ciB, const T *vB, bool compareData, double relTol, Scalars *sc);

}  // namespace sb
