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

/* Plumbing for bindings without a CUDA runtime of their own (ctypes / cgo / JNI). */
int speck_b200_malloc(speck_ctx *ctx, void **dptr, size_t bytes);
int speck_b200_free(speck_ctx *ctx, void *dptr);
int speck_b200_memcpy_h2d(speck_ctx *ctx, void *dst, const void *src, size_t bytes);
int speck_b200_memcpy_d2h(speck_ctx *ctx, void *dst, const void *src, size_t bytes);
int speck_b200_free_csr(speck_ctx *ctx, speck_csr *C); /* cudaFree the three arrays, zero C */
int speck_b200_synchronize(speck_ctx *ctx);
/* The context's main stream as a cudaStream_t (void* to keep this header CUDA-free). */
void *speck_b200_stream(speck_ctx *ctx);

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
 *   "release_workspace" 1 frees the pooled workspace now. */
int speck_b200_set_option(speck_ctx *ctx, const char *key, long long value);

#ifdef __cplusplus
}
#endif
#endif /* SPECK_B200_H */
