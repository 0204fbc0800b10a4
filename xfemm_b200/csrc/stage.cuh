// Shared-memory staging of the SELL matrix stream for the persistent solver kernels.
//
// A block's chunk of slices is contiguous in sell_val / sell_col.  Read straight from global
// memory, the bytes in flight are bounded by (resident warps) x (loads a warp can hold in
// registers): with 64 registers/thread the solver kernels keep 32 warps per SM and reach about
// half of the HBM bandwidth.  Here the stream is cut into slice-aligned tiles of <= TILE_ENT
// entries that the TMA engine copies into a ring of shared-memory stages (cp.async.bulk,
// completion on an mbarrier, L2 evict-first); the warps take slices through the block's ticket
// counter, wait for the tile of their slice, and read values/columns from shared memory, so
// the only global loads left on a warp's critical path are the gathers of x.  The warp that
// finishes the last slice of a tile re-arms the barrier and issues the copy of the tile
// STAGES ahead — no producer warp, no polling.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xb {

#ifndef XB_STAGE_TILE_ENT
#define XB_STAGE_TILE_ENT 2048
#endif
#ifndef XB_STAGE_COUNT
#define XB_STAGE_COUNT 6
#endif
constexpr int STAGE_TILE_ENT = XB_STAGE_TILE_ENT; // entries per tile (2048: 16 KB of values + 8 KB of columns)
constexpr int STAGE_COUNT = XB_STAGE_COUNT;
constexpr size_t STAGE_SMEM_BYTES = (size_t)STAGE_COUNT * STAGE_TILE_ENT * 12 + 128;
static_assert(STAGE_COUNT * (8 + 4 + 4) <= 128, "barriers and counters must fit the tail of the ring");

// per-block view of the ring in dynamic shared memory
struct StageSmem
{
    double *val;   // [STAGE_COUNT][STAGE_TILE_ENT]
    int *col;      // [STAGE_COUNT][STAGE_TILE_ENT]
    const int4 *tile; // this block's tiles in global memory: {first slice, end slice, first entry, end entry}
    uint64_t *full; // [STAGE_COUNT]
    int *remaining; // [STAGE_COUNT] slices of the resident tile not yet finished
    int *current;   // [STAGE_COUNT] tile resident (or on its way) in each stage
    int ntiles;
};

__device__ __forceinline__ StageSmem stage_carve(unsigned char *base)
{
    StageSmem S;
    S.val = (double *)base;
    S.col = (int *)(base + (size_t)STAGE_COUNT * STAGE_TILE_ENT * 8);
    S.tile = nullptr;
    S.full = (uint64_t *)(base + (size_t)STAGE_COUNT * STAGE_TILE_ENT * 12);
    S.remaining = (int *)(S.full + STAGE_COUNT);
    S.current = S.remaining + STAGE_COUNT;
    S.ntiles = 0;
    return S;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on `bar`, streamed data: L2 evict-first
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// one thread: arm the barrier of tile k's stage and start its two copies
__device__ __forceinline__ void stage_issue(const StageSmem &S, int k, const double *sell_val, const int *sell_col)
{
    const int st = k % STAGE_COUNT;
    const int4 t = S.tile[k];
    const uint32_t ent = (uint32_t)(t.w - t.z);
    S.remaining[st] = t.y - t.x;
    *(volatile int *)&S.current[st] = k;
    if (ent == 0) { mbar_arrive(&S.full[st]); return; } // slices without SELL entries (wide / padding rows only)
    mbar_arrive_expect_tx(&S.full[st], ent * 12u);
    bulk_g2s(S.val + (size_t)st * STAGE_TILE_ENT, sell_val + t.z, ent * 8u, &S.full[st]);
    bulk_g2s(S.col + (size_t)st * STAGE_TILE_ENT, sell_col + t.z, ent * 4u, &S.full[st]);
}

__device__ __forceinline__ void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// Whole block.  For every slice of [s0,s1) (tiles S.tile[0..ntiles)): acc = sum_k val*x[col] over the
// slice's SELL entries, then epi(slice, row, acc, pre(row)).  pre(row) runs as soon as the warp has its
// ticket — before it waits for the tile — so the row-indexed loads of the epilogue (diagonal, x[row], right-hand
// side) are in flight during the wait and the gathers.  `s_next` is the block's ticket counter.
// before(): runs once per warp after the ring has been started and before the warp takes its first ticket — the
// place for work that would otherwise hold up the whole block in front of the stream (the wide rows, wide_rows.cuh).
struct StageNoBefore { __device__ __forceinline__ void operator()() const {} };
template <class Pre, class Epi, class Before = StageNoBefore>
__device__ __forceinline__ void staged_slices(StageSmem &S, const int *__restrict__ slice_ptr, const double *sell_val,
                                              const int *sell_col, int s0, int s1, const double *x, int *s_next, Pre pre,
                                              Epi epi, Before before = Before())
{
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
    {
        for (int st = 0; st < STAGE_COUNT; st++) { mbar_init(&S.full[st], 1); S.remaining[st] = 0; S.current[st] = -1; }
        fence_mbar_init();
        fence_proxy_async();
        *s_next = s0;
        for (int k = 0; k < S.ntiles && k < STAGE_COUNT; k++) stage_issue(S, k, sell_val, sell_col);
    }
    __syncthreads();
    before();
    int k = 0;
    while (true)
    {
        int s = 0;
        if (lane == 0) s = atomicAdd(s_next, 1);
        s = __shfl_sync(0xffffffffu, s, 0);
        if (s >= s1) break;
        const int base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5; // in flight while we wait for the tile
        const auto pv = pre(s * 32 + lane);
        while (s >= S.tile[k].y) k++;
        const int st = k % STAGE_COUNT;
        // tickets can run several ring cycles ahead of the copies: first make sure the stage holds
        // (or is receiving) tile k, then the phase parity of its barrier is unambiguous
        while (*(volatile int *)&S.current[st] != k) {}
        mbar_wait(&S.full[st], (uint32_t)(k / STAGE_COUNT) & 1u);
        const int off = base - S.tile[k].z + lane;
        const double *v = S.val + (size_t)st * STAGE_TILE_ENT + off;
        const int *c = S.col + (size_t)st * STAGE_TILE_ENT + off;
        double acc = 0.0;
#pragma unroll 8
        for (int q = 0; q < w; q++) acc = fma(v[q * 32], x[c[q * 32]], acc);
        epi(s, s * 32 + lane, acc, pv);
        __syncwarp();
        if (lane == 0)
        {
            const int left = atomicSub(&S.remaining[st], 1) - 1;
            if (left == 0 && k + STAGE_COUNT < S.ntiles)
            {
                fence_proxy_async(); // the generic-proxy reads of this stage are done before the async write
                stage_issue(S, k + STAGE_COUNT, sell_val, sell_col);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int st = 0; st < STAGE_COUNT; st++) mbar_inval(&S.full[st]);
    __syncthreads();
}

} // namespace xb
