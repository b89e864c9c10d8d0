/*
 * speck_b200.h -- C ABI of the B200-native CSR x CSR SpGEMM that replaces the
 * hot path behind spECK::MultiplyspECK<>() (GPUPeople/spECK).
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 * The C++ template spECK::MultiplyspECK<T,4,1024,DYN,STATIC> in include/Multiply.h
 * (same signature as the reference's include/Multiply.h:15-16) is a thin wrapper
 * over speck_b200_spgemm_{f32,f64}; INTEGRATION.md shows the binding a reference
 * maintainer would add.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   speck_b200_spgemm_f64/f32  <- spECK::MultiplyspECK / MultiplyspECKImplementation,
 *                                 include/Multiply.h:15-19, source/GPU/Multiply.cu:51-1128
 *   speck_csr                  <- dCSRNoDealloc<T>, include/dCSR.h:24-35 (same field order)
 *   speck_timings              <- Timings, include/Timings.h:4-18 (same fields)
 *   speck_b200_create/destroy  <- spECKConfig::initialize/cleanup, include/spECKConfig.h:15-45
 *   speck_b200_compare_f64/f32 <- spECK::Compare, source/GPU/Compare.cu:66-82
 *   speck_b200_row_products    <- readOperations, include/common.cuh:321-459
 *
 * Error convention: every call returns 0 on success or a negative speck_status;
 * speck_b200_last_error() returns the message.  The reference's own conventions
 * (printf("ERROR: ...") and return with C untouched, source/GPU/Multiply.cu:57-97)
 * are reproduced by the C++ wrapper from these codes.
 */
#ifndef SPECK_B200_H
#define SPECK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPECK_B200_ABI_VERSION 1

typedef enum speck_status {
    SPECK_OK = 0,
    SPECK_ERR_INVALID = -1,      /* null pointers, shape mismatch (A.cols != B.rows)          */
    SPECK_ERR_TOO_LARGE = -2,    /* rows(A) or cols(B) > 2^27 (Multiply.cu:57-66)             */
    SPECK_ERR_OVERFLOW = -3,     /* nnz(C) does not fit the u32 row_offsets of the API        */
    SPECK_ERR_CUDA = -4,         /* a CUDA runtime call failed                                */
    SPECK_ERR_OOM = -5,          /* out of device memory for C or the workspace               */
    SPECK_ERR_NO_DEVICE = -6     /* no CUDA device / not an sm_100 device                     */
} speck_status;

/* Device (or, for the *_host entry points, host) CSR view.  0-based, row_offsets has
 * rows+1 entries, columns ascending and duplicate-free inside each row (precondition of
 * the reference as well, include/common.cuh:395-400). */
typedef struct speck_csr {
    size_t rows, cols, nnz;
    void *data;              /* float* or double* */
    uint32_t *row_offsets;
    uint32_t *col_ids;
} speck_csr;

/* Field-for-field mirror of the reference's Timings (milliseconds). */
typedef struct speck_timings {
    int measure_all;
    int measure_complete;
    float init, count_products, load_balance_counting, global_maps_counting, spgemm_counting,
        alloc_c, load_balance_numeric, global_maps_numeric, spgemm_numeric, sorting, cleanup,
        complete;
} speck_timings;

#define SPECK_NUM_CLASSES 32

/* What the last multiply did (for bench.py's roofline arithmetic and the tests). */
typedef struct speck_stats {
    uint64_t products;                        /* P = sum_i sum_k nnz(B_k), u64               */
    uint64_t nnz_c;
    uint32_t max_row_products;
    uint32_t class_rows[SPECK_NUM_CLASSES];   /* rows per size class         */
    uint32_t kernel_launches;                 /* kernels of this library launched by the call */
    float ms_analysis, ms_symbolic, ms_scan, ms_numeric, ms_total; /* CUDA-event ms           */
    uint64_t workspace_bytes;
} speck_stats;

typedef struct speck_ctx speck_ctx;

int speck_b200_abi_version(void);
const char *speck_b200_last_error(void);

/* Context = device id, streams, events, pooled workspace (replaces spECKConfig). */
int speck_b200_create(int device, speck_ctx **out);
int speck_b200_destroy(speck_ctx *ctx);
int speck_b200_sm_count(const speck_ctx *ctx);

/* C = A . B on the context's device.  A and B are borrowed device views.  C is in/out with
 * the reference's ownership rules (source/GPU/Multiply.cu:155-165, 589-592):
 *   - if C->rows == A->rows and C->row_offsets != NULL the row_offsets buffer is reused,
 *     otherwise the old one is cudaFree'd and a new one cudaMalloc'ed;
 *   - if C->nnz != nnz(C) (or data/col_ids are NULL) data and col_ids are cudaFree'd and
 *     re-allocated with cudaMalloc, so the caller's dCSR destructor can cudaFree them;
 *   - A->nnz == 0 or B->nnz == 0: only C->nnz = 0 is written (Multiply.cu:67-70);
 *   - P == 0: rows/cols set, nnz = 0, all three arrays freed and NULL (Multiply.cu:256-261).
 * The call synchronises the context's stream before returning. */
int speck_b200_spgemm_f64(speck_ctx *ctx, const speck_csr *A, const speck_csr *B, speck_csr *C,
                          speck_timings *timings /* may be NULL */);
int speck_b200_spgemm_f32(speck_ctx *ctx, const speck_csr *A, const speck_csr *B, speck_csr *C,
                          speck_timings *timings /* may be NULL */);

/* End-to-end variant on HOST buffers: uploads A and B (B may alias A), multiplies, downloads
 * C.  C->row_offsets/col_ids/data are written into pinned host buffers owned by the context
 * (valid until the next *_host call or speck_b200_destroy).  h2d/d2h byte counts returned. */
int speck_b200_spgemm_host_f64(speck_ctx *ctx, const speck_csr *A, const speck_csr *B,
                               speck_csr *C, uint64_t *h2d_bytes, uint64_t *d2h_bytes);
int speck_b200_spgemm_host_f32(speck_ctx *ctx, const speck_csr *A, const speck_csr *B,
                               speck_csr *C, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

int speck_b200_get_stats(const speck_ctx *ctx, speck_stats *out);

/* Analysis only: rowOperations (u32[rows], device pointer, may be NULL), P and the row maximum. */
int speck_b200_row_products(speck_ctx *ctx, const speck_csr *A, const speck_csr *B,
                            uint32_t *d_row_ops, uint64_t *products, uint32_t *max_row_products);

/* Device-side comparison of two CSR matrices: 1 = equal, 0 = different, <0 = error.
 * Row lengths and positional column ids always; values with relative tolerance rel_tol when
 * compare_data != 0 (the reference hard-codes 1 %, Compare.cu:49-58). */
int speck_b200_compare_f64(speck_ctx *ctx, const speck_csr *ref, const speck_csr *cmp,
                           int compare_data, double rel_tol);
int speck_b200_compare_f32(speck_ctx *ctx, const speck_csr *ref, const speck_csr *cmp,
                           int compare_data, double rel_tol);

/* The same comparison with a report of the FIRST difference (smallest row, then kind, then position), which the
 * reference's d_compare cannot give (it only raises a flag, Compare.cu:27-58).  Returns 1 / 0 / <0 like above. */
typedef struct speck_mismatch {
    uint64_t row;            /* first row that differs                                                  */
    uint32_t kind;           /* 0 = row length, 1 = column id, 2 = value, 3 = shape / nnz of the matrices */
    uint32_t index_in_row;   /* position inside the row (kinds 1 and 2)                                  */
    uint32_t ref_len, cmp_len;
    uint32_t ref_col, cmp_col;
    double ref_val, cmp_val; /* values at that position (kind 2; 0 when a value array is missing)        */
} speck_mismatch;
int speck_b200_compare_report_f64(speck_ctx *ctx, const speck_csr *ref, const speck_csr *cmp, int compare_data,
                                  double rel_tol, speck_mismatch *first);
int speck_b200_compare_report_f32(speck_ctx *ctx, const speck_csr *ref, const speck_csr *cmp, int compare_data,
                                  double rel_tol, speck_mismatch *first);

/* GPU-side COO -> CSR for the loader path (replaces the host conversion, reference source/CSR.cpp:173-212): the
 * (row, column) pairs are sorted on the device, stable; duplicate_policy SPECK_COO_KEEP keeps repeated (row, column)
 * entries next to each other in input order (what the reference's loader does), SPECK_COO_SUM folds them into one
 * entry (the duplicate-free form the multiply requires).  Inputs are device arrays of nnz entries; `out` receives
 * cudaMalloc'ed arrays (free with speck_b200_free_csr).  Returns a speck_status (no message is recorded). */
#define SPECK_COO_KEEP 0
#define SPECK_COO_SUM 1
int speck_b200_coo_to_csr_f64(speck_ctx *ctx, size_t rows, size_t cols, size_t nnz, const uint32_t *d_row_ids,
                              const uint32_t *d_col_ids, const double *d_values, int duplicate_policy, speck_csr *out);
int speck_b200_coo_to_csr_f32(speck_ctx *ctx, size_t rows, size_t cols, size_t nnz, const uint32_t *d_row_ids,
                              const uint32_t *d_col_ids, const float *d_values, int duplicate_policy, speck_csr *out);

/* Plumbing for bindings without a CUDA runtime of their own (ctypes / cgo / JNI). */
int speck_b200_malloc(speck_ctx *ctx, void **dptr, size_t bytes);
int speck_b200_free(speck_ctx *ctx, void *dptr);
int speck_b200_memcpy_h2d(speck_ctx *ctx, void *dst, const void *src, size_t bytes);
int speck_b200_memcpy_d2h(speck_ctx *ctx, void *dst, const void *src, size_t bytes);
int speck_b200_free_csr(speck_ctx *ctx, speck_csr *C); /* cudaFree the three arrays, zero C */
int speck_b200_synchronize(speck_ctx *ctx);
/* The context's main stream as a cudaStream_t (void* to keep this header CUDA-free). */
void *speck_b200_stream(speck_ctx *ctx);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU (the reference is single-device, source/Executor.cpp:25 hard-codes device 0).
 * Rows of C depend only on the matching rows of A and on all of B: A is cut into contiguous row slabs balanced by
 * intermediate products, B is replicated once at setup (uploaded to the first device, copied to its peers over
 * NVLink), every device multiplies its slab -- no exchange in the per-multiply path -- and the slabs of C are
 * optionally concatenated on the first device (peer copies + a row_offsets fix-up).
 * ------------------------------------------------------------------------------------------ */
#define SPECK_MAX_SHARDS 16

/* Row cuts of A balanced by products, computed on the device: analysis (row products) -> 64-bit scan -> search.
 * cuts: host array of parts + 1 row indices (cuts[0] = 0, cuts[parts] = A->rows); part_products (may be NULL):
 * host array of parts product counts (of the balanced cost when the partition_*_cost options are set).  A and B are device views on ctx's device. */
int speck_b200_partition_rows(speck_ctx *ctx, const speck_csr *A, const speck_csr *B, int parts, uint32_t *cuts,
                              uint64_t *part_products);

typedef struct speck_shard_info {
    int shards;
    int concatenated;
    uint32_t cuts[SPECK_MAX_SHARDS + 1];
    uint64_t products[SPECK_MAX_SHARDS];
    uint64_t nnz_c[SPECK_MAX_SHARDS];
    float ms_device[SPECK_MAX_SHARDS];   /* CUDA-event ms of each slab's multiply on its device            */
    float ms_setup;                      /* uploads, peer broadcast of B, partition (host wall clock)      */
    float ms_multiply;                   /* host wall clock around the concurrent slab multiplies          */
    float ms_concat;                     /* peer copies + offset fix-up into one CSR on the first device   */
} speck_shard_info;

typedef struct speck_shard_plan speck_shard_plan;

/* Setup: HOST CSR A and B -> slabs of A on ctxs[0..n), B on every device.  One context per device. */
int speck_b200_sharded_create_f64(speck_ctx **ctxs, int n, const speck_csr *A_host, const speck_csr *B_host,
                                  speck_shard_plan **plan);
int speck_b200_sharded_create_f32(speck_ctx **ctxs, int n, const speck_csr *A_host, const speck_csr *B_host,
                                  speck_shard_plan **plan);
/* One multiply: every device multiplies its slab concurrently (one host thread per device); the slabs of C stay
 * on their devices and are reused by the next call (the reference's C-reuse rule per slab). */
int speck_b200_sharded_multiply(speck_shard_plan *plan, speck_shard_info *info);
/* Concatenate the slabs of the last multiply into one device CSR on ctxs[0] (C in/out, same ownership rules as
 * speck_b200_spgemm_*).  SPECK_ERR_OVERFLOW when the total nnz does not fit u32 row_offsets: keep C distributed. */
int speck_b200_sharded_concat(speck_shard_plan *plan, speck_csr *C, speck_shard_info *info);
/* Borrowed views of slab g (device pointers on ctxs[g]'s device): its rows of A and of the last C. */
int speck_b200_sharded_slab(speck_shard_plan *plan, int g, speck_csr *A_slab, speck_csr *C_slab);
int speck_b200_sharded_destroy(speck_shard_plan *plan);

/* One process per GPU (torchrun / MPI launches; bench.py --gpus N): concatenation of the slabs of C on one device WITHOUT
 * a collective library.  The gathering process allocates the three arrays of the concatenated C (speck_b200_malloc) and
 * exports them (CUDA IPC, 64-byte handles that travel over any host channel); every other process opens them and PUSHES
 * its slab with one kernel of 16-byte peer stores over NVLink / NVSwitch: col_ids and values at the slab's nnz offset,
 * row_offsets with that offset added (the fix-up is fused into the transfer).  The gathering process pushes its own
 * slab the same way (local stores).  `last` != 0: the slab also writes the final row_offsets[rows] entry.
 * device_ms (may be NULL): CUDA-event time of the push kernel.  The call returns when the stores are complete. */
#define SPECK_IPC_HANDLE_BYTES 64
int speck_b200_ipc_export(speck_ctx *ctx, const void *dptr, unsigned char handle[SPECK_IPC_HANDLE_BYTES]);
int speck_b200_ipc_open(speck_ctx *ctx, const unsigned char handle[SPECK_IPC_HANDLE_BYTES], void **dptr);
int speck_b200_ipc_close(speck_ctx *ctx, void *dptr);
int speck_b200_push_slab_f64(speck_ctx *ctx, const speck_csr *C_slab, uint64_t nnz_base, uint32_t row_base, int last,
                             uint32_t *dst_row_offsets, uint32_t *dst_col_ids, double *dst_data, float *device_ms);
int speck_b200_push_slab_f32(speck_ctx *ctx, const speck_csr *C_slab, uint64_t nnz_base, uint32_t row_base, int last,
                             uint32_t *dst_row_offsets, uint32_t *dst_col_ids, float *dst_data, float *device_ms);

/* Tuning knobs (integers; the defaults are the measured best, the switches exist so that the
 * tests can drive every kernel family over the same inputs):
 *   "sort_max"          largest row-product count handled by the sort / rank classes (power of two or
 *                       multiple of 512 in [4, 16384]; rows with more products take the bitmap path)
 *   "rank_path"         1 (default): rows of 513..16384 products use the bitmap-rank kernels,
 *                       0: the CTA bitonic-sort kernels (always used when cols(B) > 2^20)
 *   "rank_map"          1 (default): the symbolic phase records every product's sorted position
 *                       (2 B per product of workspace) and the numeric phase only gathers and scatters,
 *                       0: self-contained numeric kernels
 *   "sym_streams", "num_streams"   side streams of the two phases (1..4; num_streams 0 = automatic)
 *   "map_min_class", "map_cta_min" which lane-group classes use the rank map / the CTA map kernel
 *   "flat_sym"          1 (default): mapped two-level symbolic kernel = flat variant with cp.async-staged columns
 *                       (rank_flat.cuh), 0: register-slot variant (rank_cta.cuh); "flat_e" 8 / 16 slots per thread
 *   "dense_seq"         banded / high-compression rows: 1 (default) sequential-k numeric kernel with lane loads,
 *                       2: the same with TMA (cp.async.bulk) staged B segments, 0: product-parallel kernel only
 *   "test_set"          1 (default): the symbolic bitmap kernel of the banded / high-compression rows reads the bitmap
 *                       word before its atomicOr (most bits are set already when many products fold into a column)
 *   "deterministic"     1: values bit-reproducible, summed in sequential order (ascending k, products
 *                       rounded before they are added); slower: no rank map, sort classes, rows that accumulate
 *                       with atomics are recomputed.  Default 0 (like the reference: "not bit stable")
 *   "tiered_analysis"   1: the analysis gathers B's row_offsets only and fetches column extents in a second pass for
 *                       rows with >= 128 products; 0 (default): one 16-byte row summary per A entry (measured faster)
 *   "spin_wait"         1 (default): the two mid-pipeline scalar read-backs poll mapped pinned memory, 0: they use
 *                       cudaMemcpyAsync + cudaStreamSynchronize
 *   "partition_row_cost", "partition_entry_cost"   speck_b200_partition_rows balances products + entry_cost * nnz(A row)
 *                       + row_cost per row instead of products alone (default 0, 0).  Measured on the config-5
 *                       matrix: a row costs about as much as 7 products, an entry of A about 1.5 (profiles/r2_notes.md)
 *   "col_direct"        1..3: large mapped numeric shapes stage values only and write column ids straight to C
 *                       (measured slower; default 0)
 *   "narrow_groups"     2 (default): mapped rows of <= 128 products share a warp (4 / 8 / 16 lanes per row with 2..8
 *                       products per lane) in the symbolic sort and the numeric kernels; 1: rows of <= 64 products only;
 *                       0: one row per warp from 17 products; 3: the 256-product class as well (measured slower);
 *                       "narrow_numeric" -1 (default: follow narrow_groups) .. 3 sets the numeric kernels alone
 *   "big_split"         1..3: rows of 4097..16384 products take several CTAs per row in the numeric phase
 *                       (measured slower; default 0)
 *   "flat_min_class"    6 / 7: the 129..256 / 257..512-product classes rank with the flat bitmap kernel instead of the
 *                       register bitonic sort (measured slower; default 8 = off)
 *   "sym_mix"           N > 0: the sort kernels of the three largest lane-group classes run as capped grids (about N CTAs
 *                       per SM) next to the bitmap rank kernels instead of after them (measured slower; default 0)
 *   "num_plan"          1: numeric phase of large multiplies (which use one stream): the one-CTA-per-SM kernels on a
 *                       second stream next to the small shapes (experiment; default 0)
 *   "seg_num", "hash_count"        experiments kept for the record (profiles/r2_notes.md), off by default
 *   "rank_map_max_bytes"  upper bound of the rank-map workspace (-1 = no bound); a multiply whose map would be larger
 *                       runs the self-contained numeric kernels instead (tests use it to force that fallback)
 *   "release_workspace" 1 frees the pooled workspace now. */
int speck_b200_set_option(speck_ctx *ctx, const char *key, long long value);

#ifdef __cplusplus
}
#endif
#endif /* SPECK_B200_H */
