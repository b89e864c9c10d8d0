// speck_b200/csrc/dense_seq.cuh -- numeric phase of the banded / high-compression rows ("dense_local" bin) with
// TMA-staged B-row segments and sequential-k accumulation (round 2).
//
// Replaces, for rows whose distinct columns fit the shared accumulator, the product-parallel pass B of
// k_dense_rows (kernels_dense.cu) and through it denseSpGEMMNumeric of the reference
// (include/GPU/spECK_HashSpGEMM.cuh:1300-1472).  On FEM-like matrices ~18 products fold into every entry of C;
// accumulating them product-parallel needs one shared-memory fp64 atomicAdd per product, which sm_100 executes
// as a compare-and-swap loop (ATOMS.CAST.SPIN), contended because 256 threads hit ~330 accumulators.
// Here ONE WARP owns the row and walks its A entries in ascending k:
//   * the columns of one B row are distinct, so the lanes of a step never touch the same accumulator:
//     plain shared-memory read-modify-write, no atomics;
//   * the order of additions per entry of C is ascending k with separately rounded products (__dmul_rn /
//     __dadd_rn) -- the order of a sequential Gustavson loop, so these rows are bit-reproducible;
//   * the B-row segments (columns and values) are staged into a shared-memory ring by bulk asynchronous copies
//     (cp.async.bulk + mbarrier, issued by one lane, RING segments ahead of the consumer), which is what the
//     B segments of these matrices suit: tens to hundreds of contiguous entries each.  Segments are fetched as
//     16-byte-aligned supersets (<= 3 entries of over-fetch at either end, masked by index).
// The sorted column ids come straight from the row's bitmap, kept by the symbolic phase.
#pragma once
#include "common.cuh"

namespace sb {

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA), completion reported to the mbarrier as transferred bytes
__device__ __forceinline__ void bulk_g2s(void *smemDst, const void *gsrc, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smemDst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }

// Rows the sequential-k kernel takes: distinct columns fit its accumulator and at least DENSE_SEQ_FOLD products
// fold into every entry of C on average.  Measured (profiles/r2_notes.md): at 17.7 products per entry (cant-like)
// it runs 1.4x faster than the product-parallel kernel, whose shared fp64 atomics then contend; at 2.2 (banded
// generator) it is 1.4x slower (short B rows leave most lanes of a step idle).
constexpr u32 DENSE_SEQ_FOLD = 6;   // fold = 0 (deterministic mode): every row that fits
__host__ __device__ __forceinline__ bool dense_seq_takes(u32 products, u32 nnzRow, u32 seqMax, u32 fold)
{
    return nnzRow <= seqMax && (u64)products >= (u64)fold * nnzRow;
}

constexpr int SEQ_WORDS = 1 << (DENSE_LOCAL_BITS - 5);   // bitmap words of a local row (512)
constexpr int SEQ_MAXSEG = 128;                          // entries per staged piece of a B row
constexpr int SEQ_RING = 4;                              // pieces in flight

template <typename T, int SV, bool TMA>
struct DenseSeqSmem {
    alignas(16) u32 bitmap[SEQ_WORDS];
    alignas(16) u32 ringC[TMA ? SEQ_RING : 1][TMA ? SEQ_MAXSEG : 4];   // staging ring of the TMA variant only
    alignas(16) T ringV[TMA ? SEQ_RING : 1][TMA ? SEQ_MAXSEG : 4];
    alignas(16) T svals[SV];
    unsigned short wordPre[SEQ_WORDS];
    alignas(8) u64 bar[SEQ_RING];
};

// walks the pieces (A entry ascending, then 16-byte-aligned pieces of its B row) of one row; warp-uniform
struct PieceCursor {
    u32 ab;        // first A entry of the loaded batch of 32
    u32 e;         // entry inside the batch
    u32 ps;        // next piece start (element index in B, multiple of 4)
    u32 bs, be;    // batch registers: B-row bounds of entry `lane`
};

// TMA = true : B segments staged by cp.async.bulk into the shared ring (16-byte aligned pieces)
// TMA = false: B segments loaded by the lanes themselves (coalesced, one piece of <= 128 entries prefetched into
//              registers while the previous one is accumulated)
template <typename T, int SV, bool TMA>
__global__ void __launch_bounds__(32)
k_dense_seq(const u32 *__restrict__ perm, const u32 count, const u32 *__restrict__ aRp, const u32 *__restrict__ aCi,
            const T *__restrict__ aV, const u32 *__restrict__ bRp, const u32 *__restrict__ bCi, const T *__restrict__ bV,
            const u32 *__restrict__ rowMin, const u32 *__restrict__ rowMax, const u32 *__restrict__ bitmapStore,
            const u32 *__restrict__ cRp, u32 *__restrict__ cCi, T *__restrict__ cV, const u32 minNnz,
            const u32 *__restrict__ rowOps, const u32 fold)
{
    __shared__ DenseSeqSmem<T, SV, TMA> sm;
    const u32 lane = threadIdx.x;
    const u32 ri = blockIdx.x;
    if (ri >= count) return;
    const u32 row = perm[ri];
    const u32 cBase = cRp[row], nnzRow = cRp[row + 1] - cBase;
    // another shape of this kernel, or k_dense_rows, takes the row (same test there: dense_seq_takes)
    if (nnzRow > (u32)SV || nnzRow <= minNnz || !dense_seq_takes(rowOps[row], nnzRow, (u32)DENSE_SEQ_MAX, fold)) return;
    const u32 aBeg = aRp[row], aEnd = aRp[row + 1];
    const u32 base0 = rowMin[row] & ~127u;
    const u32 extWords = ((rowMax[row] - base0) >> 5) + 1;   // <= SEQ_WORDS for rows of this bin

    if (TMA) {
        if (lane < SEQ_RING) mbar_init(&sm.bar[lane], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }

    // ---------------------------------------------------------------- bitmap kept by the symbolic phase, word prefixes
    const u32 *store = bitmapStore + (size_t)ri * SEQ_WORDS;
    u32 cnt = 0;
    u32 wv[SEQ_WORDS / 32];
#pragma unroll
    for (int k = 0; k < SEQ_WORDS / 32; ++k) {   // lane owns the 16 consecutive words lane*16 ..
        const u32 w = lane * (SEQ_WORDS / 32) + k;
        wv[k] = w < extWords ? __ldg(store + w) : 0u;
        sm.bitmap[w] = wv[k];
        cnt += __popc(wv[k]);
    }
    u32 incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (u32)d) incl += t;
    }
    u32 run = incl - cnt;
#pragma unroll
    for (int k = 0; k < SEQ_WORDS / 32; ++k) {
        const u32 w = lane * (SEQ_WORDS / 32) + k;
        sm.wordPre[w] = (unsigned short)run;
        u32 bits = wv[k];
        while (bits) {   // sorted column ids of the row straight from the bitmap
            const u32 b = __ffs(bits) - 1;
            bits &= bits - 1;
            cCi[cBase + run++] = base0 + w * 32 + b;
        }
    }
    for (u32 j = lane; j < nnzRow; j += 32) sm.svals[j] = (T)0;
    __syncwarp();

    // ---------------------------------------------------------------- producer / consumer over the row's pieces
    auto load_batch = [&](PieceCursor &c) {
        const u32 i = c.ab + lane;
        c.bs = c.be = 0;
        if (i < aEnd) {
            const u32 k = __ldg(aCi + i);
            c.bs = __ldg(bRp + k);
            c.be = __ldg(bRp + k + 1);
        }
        c.e = 0;
    };
    // positions the cursor on the next piece; false when the row is exhausted.  *pBs/*pBe: the entry's B-row bounds,
    // *pPs/*pPe: the aligned piece, *entry: index of the A entry
    auto next_piece = [&](PieceCursor &c, u32 &pBs, u32 &pBe, u32 &pPs, u32 &pPe, u32 &entry) -> bool {
        while (true) {
            if (c.ab >= aEnd) return false;
            if (c.e >= 32 || c.ab + c.e >= aEnd) {
                c.ab += 32;
                if (c.ab >= aEnd) return false;
                load_batch(c);
                c.ps = 0xffffffffu;
            }
            const u32 bs = __shfl_sync(0xffffffffu, c.bs, c.e), be = __shfl_sync(0xffffffffu, c.be, c.e);
            if (c.ps == 0xffffffffu) c.ps = TMA ? (bs & ~3u) : bs;
            if (be > bs && c.ps < be) {
                pBs = bs; pBe = be; pPs = c.ps;
                pPe = min(TMA ? ((be + 3u) & ~3u) : be, c.ps + (u32)SEQ_MAXSEG);
                entry = c.ab + c.e;
                c.ps = pPe;
                if (c.ps >= be) { ++c.e; c.ps = 0xffffffffu; }
                return true;
            }
            ++c.e;
            c.ps = 0xffffffffu;
        }
    };
    auto issue = [&](PieceCursor &c, u32 step) -> bool {
        u32 bs, be, ps, pe, entry;
        if (!next_piece(c, bs, be, ps, pe, entry)) return false;
        if (lane == 0) {
            const u32 slot = step % SEQ_RING, n = pe - ps;
            mbar_arrive_expect_tx(&sm.bar[slot], n * (4u + (u32)sizeof(T)));
            bulk_g2s(sm.ringC[slot], bCi + ps, n * 4u, &sm.bar[slot]);
            bulk_g2s(sm.ringV[slot], bV + ps, n * (u32)sizeof(T), &sm.bar[slot]);
        }
        return true;
    };

    auto accumulate = [&](u32 col, T bv, T av) {
        const u32 c = col - base0;
        const u32 w = c >> 5;
        const u32 rank = sm.wordPre[w] + __popc(sm.bitmap[w] & ((1u << (c & 31)) - 1u));
        sm.svals[rank] = add_rn<T>(sm.svals[rank], mul_rn<T>(av, bv));
    };
    u32 bs, be, ps, pe, entry;
    if (TMA) {
        PieceCursor prod{aBeg, 0, 0xffffffffu, 0, 0}, cons{aBeg, 0, 0xffffffffu, 0, 0};
        if (aBeg < aEnd) { load_batch(prod); cons.bs = prod.bs; cons.be = prod.be; }
        u32 issued = 0;
        bool more = aBeg < aEnd;
        for (; more && issued < (u32)SEQ_RING; ++issued) more = issue(prod, issued);
        if (!more && issued) --issued;   // the last call found nothing to issue
        u32 step = 0;
        while (aBeg < aEnd && next_piece(cons, bs, be, ps, pe, entry)) {
            const u32 slot = step % SEQ_RING;
            const T av = __ldg(aV + entry);
            mbar_wait(&sm.bar[slot], (step / SEQ_RING) & 1u);
            const u32 *pc = sm.ringC[slot];
            const T *pv = sm.ringV[slot];
#pragma unroll 2
            for (u32 j = lane; j < pe - ps; j += 32) {
                const u32 idx = ps + j;
                if (idx >= bs && idx < be) accumulate(pc[j], pv[j], av);   // the aligned piece may over-fetch <= 3 entries per end
            }
            __syncwarp();
            ++step;
            if (more) {   // refill the slot that was just consumed
                fence_proxy_async();
                more = issue(prod, issued);
                if (more) ++issued;
            }
        }
    } else {
        constexpr int U = SEQ_MAXSEG / 32;
        PieceCursor cons{aBeg, 0, 0xffffffffu, 0, 0};
        if (aBeg < aEnd) load_batch(cons);
        u32 nc[U], cc[U];
        T nv[U], cv[U];
        T nav = (T)0;
        auto fetch = [&]() {   // the next piece into registers
            nav = __ldg(aV + entry);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const u32 idx = ps + lane + 32u * u;
                const bool ok = idx < pe;
                nc[u] = ok ? __ldg(bCi + idx) : 0xffffffffu;
                nv[u] = ok ? __ldg(bV + idx) : (T)0;
            }
        };
        bool have = aBeg < aEnd && next_piece(cons, bs, be, ps, pe, entry);
        if (have) fetch();
        while (have) {
            const T av = nav;
#pragma unroll
            for (int u = 0; u < U; ++u) { cc[u] = nc[u]; cv[u] = nv[u]; }
            have = next_piece(cons, bs, be, ps, pe, entry);
            if (have) fetch();          // in flight while the current piece is accumulated
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (cc[u] != 0xffffffffu) accumulate(cc[u], cv[u], av);
            __syncwarp();
        }
    }
    __syncwarp();
    for (u32 j = lane; j < nnzRow; j += 32) cV[cBase + j] = sm.svals[j];
}

template <typename T, int SV, bool TMA>
void launch_dense_seq_t(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi, const T *aV,
                        const u32 *bRp, const u32 *bCi, const T *bV, const u32 *rowMin, const u32 *rowMax,
                        const u32 *bitmapStore, const u32 *cRp, u32 *cCi, T *cV, u32 minNnz, const u32 *rowOps, u32 fold)
{
    k_dense_seq<T, SV, TMA><<<count, 32, 0, lc.stream>>>(perm, count, aRp, aCi, aV, bRp, bCi, bV, rowMin, rowMax, bitmapStore, cRp, cCi, cV, minNnz, rowOps, fold);
    ++*lc.launches;
}

// ------------------------------------------------------------------------------------------------
// Deterministic mode (option "deterministic"; the reference is "not bit stable", config.ini:8-9): values of the
// rows whose numeric kernel accumulates with atomics (rows beyond the sort classes, wide bitmap rows) are
// recomputed in sequential ascending-k order.  The columns of the row are already in C (sorted); one CTA walks the A entries
// in ascending k, the products of one B row go to distinct entries of C (found by binary search), so every step is
// a plain read-modify-write; products are rounded before they are added.  Slow (a barrier per A entry), used
// for the few rows no deterministic kernel takes.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_det_rows(const u32 *__restrict__ perm, const u32 count, const u32 *__restrict__ aRp, const u32 *__restrict__ aCi,
           const T *__restrict__ aV, const u32 *__restrict__ bRp, const u32 *__restrict__ bCi, const T *__restrict__ bV,
           const u32 *__restrict__ cRp, const u32 *__restrict__ cCi, T *cV, const u32 *__restrict__ rowOps,
           const u32 seqMax, const u32 fold)
{
    for (u32 ri = blockIdx.x; ri < count; ri += gridDim.x) {
        const u32 row = perm[ri];
        const u32 cBase = cRp[row], nnzRow = cRp[row + 1] - cBase;
        if (seqMax && dense_seq_takes(rowOps[row], nnzRow, seqMax, fold)) continue;   // k_dense_seq wrote this row
        volatile T *out = cV + cBase;
        for (u32 j = threadIdx.x; j < nnzRow; j += 256) out[j] = (T)0;
        __syncthreads();
        const u32 aBeg = aRp[row], aEnd = aRp[row + 1];
        for (u32 e = aBeg; e < aEnd; ++e) {
            const u32 k = __ldg(aCi + e);
            const T av = __ldg(aV + e);
            const u32 bs = __ldg(bRp + k), be = __ldg(bRp + k + 1);
            for (u32 q = bs + threadIdx.x; q < be; q += 256) {
                const u32 col = __ldg(bCi + q);
                u32 lo = 0, hi = nnzRow;   // position of col in the row of C
                while (lo < hi) {
                    const u32 mid = (lo + hi) >> 1;
                    if (__ldg(cCi + cBase + mid) < col) lo = mid + 1; else hi = mid;
                }
                out[lo] = add_rn<T>(out[lo], mul_rn<T>(av, __ldg(bV + q)));
            }
            __syncthreads();
        }
    }
}

template <typename T>
void launch_det_rows_t(const LaunchCtx &lc, const u32 *perm, u32 count, const u32 *aRp, const u32 *aCi, const T *aV,
                       const u32 *bRp, const u32 *bCi, const T *bV, const u32 *cRp, const u32 *cCi, T *cV,
                       const u32 *rowOps, u32 seqMax, u32 fold)
{
    if (count == 0) return;
    u32 grid = (u32)lc.smCount * 8u;
    if (grid > count) grid = count;
    k_det_rows<T><<<grid, 256, 0, lc.stream>>>(perm, count, aRp, aCi, aV, bRp, bCi, bV, cRp, cCi, cV, rowOps, seqMax, fold);
    ++*lc.launches;
}

}  // namespace sb
