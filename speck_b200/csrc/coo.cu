// speck_b200/csrc/coo.cu -- GPU-side COO -> CSR.
//
// Replaces, on the device, the loader's host conversion (reference source/CSR.cpp:173-212: std::sort of 16-byte
// entries by (row, column), then a histogram + scan), which dominates the start-up for webbase-sized files.
//   1. keys (row << 32 | col) + entry index sorted by an LSD radix sort (cub::DeviceRadixSort, toolkit CUB: this
//      is the loader, not the SpGEMM hot path; stable, so equal (row, col) keep their input order, which is what
//      the host layer's std::stable_sort gives),
//   2. duplicate policy: KEEP (the reference keeps duplicates, adjacent) or SUM (runs of equal (row, col) folded in
//      input order into their first entry -- the duplicate-free form the multiply requires),
//   3. row_offsets from the sorted keys (each entry that starts a new row fills the offsets of the rows it skips).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/speck_b200.h"
#include "common.cuh"

using namespace sb;

namespace {

__global__ void __launch_bounds__(256) k_coo_keys(u64 n, const u32 *__restrict__ r, const u32 *__restrict__ c,
                                                  u64 *__restrict__ keys, u32 *__restrict__ idx)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ((u64)r[i] << 32) | c[i];
    idx[i] = (u32)i;
}

// head[i] = 1 when entry i of the sorted list starts a new (row, col) (always, under the KEEP policy)
__global__ void __launch_bounds__(256) k_coo_heads(u64 n, const u64 *__restrict__ keys, u32 *__restrict__ head, int sum)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (!sum || i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// pos = exclusive scan of head.  Heads write their column, fold their run's values in input order and fill the
// row_offsets of every row in (previous entry's row, own row].
template <typename T>
__global__ void __launch_bounds__(256) k_coo_emit(u64 n, u32 rows, const u64 *__restrict__ keys, const u32 *__restrict__ idx,
                                                  const u32 *__restrict__ head, const u32 *__restrict__ pos,
                                                  const T *__restrict__ vIn, u32 *__restrict__ rp, u32 *__restrict__ ci,
                                                  T *__restrict__ v)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!head[i]) return;
    const u64 key = keys[i];
    const u32 o = pos[i];
    T sum = vIn[idx[i]];
    for (u64 j = i + 1; j < n && !head[j]; ++j) sum += vIn[idx[j]];
    ci[o] = (u32)key;
    v[o] = sum;
    const u32 row = (u32)(key >> 32);
    const u32 prevRow = i == 0 ? 0xffffffffu : (u32)(keys[i - 1] >> 32);
    if (i == 0 || prevRow != row)
        for (u32 r = (i == 0 ? 0u : prevRow + 1u); r <= row; ++r) rp[r] = o;
}

// rows after the last entry's row are empty; row_offsets[rows] = number of entries written
__global__ void __launch_bounds__(256) k_coo_tail(u64 n, u32 rows, const u64 *__restrict__ keys, const u32 *__restrict__ head,
                                                  const u32 *__restrict__ pos, u32 *__restrict__ rp, u32 *__restrict__ nnzOut)
{
    const u32 total = pos[n - 1] + head[n - 1];
    const u32 lastRow = (u32)(keys[n - 1] >> 32);
    for (u32 r = lastRow + 1 + threadIdx.x; r <= rows; r += blockDim.x) rp[r] = total;
    if (threadIdx.x == 0) *nnzOut = total;
}

thread_local char g_cooErr[256] = "";

template <typename T>
int coo_impl(speck_ctx *ctx, size_t rows, size_t cols, size_t nnz, const u32 *dRow, const u32 *dCol, const T *dVal,
             int policy, speck_csr *out)
{
    if (!ctx || !out || (nnz && (!dRow || !dCol || !dVal))) return SPECK_ERR_INVALID;
    if (rows > 0xffffffffull || cols > 0xffffffffull || nnz > 0xffffffffull) return SPECK_ERR_TOO_LARGE;
    cudaStream_t st = (cudaStream_t)speck_b200_stream(ctx);
    *out = speck_csr{};
    out->rows = rows;
    out->cols = cols;
    if (cudaMalloc((void **)&out->row_offsets, (rows + 1) * 4) != cudaSuccess) return SPECK_ERR_OOM;
    if (nnz == 0 || rows == 0) {
        cudaMemsetAsync(out->row_offsets, 0, (rows + 1) * 4, st);
        cudaMalloc((void **)&out->col_ids, 4);
        cudaMalloc(&out->data, sizeof(T));
        return cudaStreamSynchronize(st) == cudaSuccess ? SPECK_OK : SPECK_ERR_CUDA;
    }
    u64 *keys = nullptr, *keysSorted = nullptr;
    u32 *idx = nullptr, *idxSorted = nullptr, *head = nullptr, *pos = nullptr, *dNnz = nullptr;
    void *tmp = nullptr;
    size_t tmpBytes = 0, scanBytes = 0;
    int rc = SPECK_OK;
    auto cleanup = [&]() {
        for (void *p : {(void *)keys, (void *)keysSorted, (void *)idx, (void *)idxSorted, (void *)head, (void *)pos, (void *)dNnz, tmp})
            if (p) cudaFree(p);
    };
    int rowBits = 1;   // only the row bits that can be set take part in the sort (fewer radix passes for small matrices)
    while (rowBits < 32 && (rows - 1) >> rowBits) ++rowBits;
    cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keysSorted, idx, idxSorted, (int)nnz, 0, 32 + rowBits, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, head, pos, (int)nnz, st);
    if (scanBytes > tmpBytes) tmpBytes = scanBytes;
    bool ok = cudaMalloc((void **)&keys, nnz * 8) == cudaSuccess && cudaMalloc((void **)&keysSorted, nnz * 8) == cudaSuccess &&
              cudaMalloc((void **)&idx, nnz * 4) == cudaSuccess && cudaMalloc((void **)&idxSorted, nnz * 4) == cudaSuccess &&
              cudaMalloc((void **)&head, nnz * 4) == cudaSuccess && cudaMalloc((void **)&pos, nnz * 4) == cudaSuccess &&
              cudaMalloc((void **)&dNnz, 4) == cudaSuccess && cudaMalloc(&tmp, tmpBytes ? tmpBytes : 4) == cudaSuccess &&
              cudaMalloc((void **)&out->col_ids, nnz * 4) == cudaSuccess && cudaMalloc(&out->data, nnz * sizeof(T)) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        cleanup();
        speck_b200_free_csr(ctx, out);
        return SPECK_ERR_OOM;
    }
    const u32 grid = (u32)((nnz + 255) / 256);
    k_coo_keys<<<grid, 256, 0, st>>>(nnz, dRow, dCol, keys, idx);
    cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keysSorted, idx, idxSorted, (int)nnz, 0, 32 + rowBits, st);
    k_coo_heads<<<grid, 256, 0, st>>>(nnz, keysSorted, head, policy == SPECK_COO_SUM);
    cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, head, pos, (int)nnz, st);
    k_coo_emit<T><<<grid, 256, 0, st>>>(nnz, (u32)rows, keysSorted, idxSorted, head, pos, dVal, out->row_offsets, out->col_ids,
                                        (T *)out->data);
    k_coo_tail<<<1, 256, 0, st>>>(nnz, (u32)rows, keysSorted, head, pos, out->row_offsets, dNnz);
    u32 nnzOut = 0;
    cudaMemcpyAsync(&nnzOut, dNnz, 4, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = SPECK_ERR_CUDA;
    out->nnz = nnzOut;
    cleanup();
    if (rc != SPECK_OK) speck_b200_free_csr(ctx, out);
    return rc;
}

}  // namespace

extern "C" {

int speck_b200_coo_to_csr_f64(speck_ctx *ctx, size_t rows, size_t cols, size_t nnz, const uint32_t *d_row_ids,
                              const uint32_t *d_col_ids, const double *d_values, int duplicate_policy, speck_csr *out)
{
    return coo_impl<double>(ctx, rows, cols, nnz, d_row_ids, d_col_ids, d_values, duplicate_policy, out);
}
int speck_b200_coo_to_csr_f32(speck_ctx *ctx, size_t rows, size_t cols, size_t nnz, const uint32_t *d_row_ids,
                              const uint32_t *d_col_ids, const float *d_values, int duplicate_policy, speck_csr *out)
{
    return coo_impl<float>(ctx, rows, cols, nnz, d_row_ids, d_col_ids, d_values, duplicate_policy, out);
}

}  // extern "C"
