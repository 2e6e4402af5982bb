// ts2d_pipe.cuh -- producer/consumer staging of per-tile triangle lists (sm_90+/sm_100a).
//
// The composite kernels walk each tile's slice of the sorted instance list.  Staging that slice
// cooperatively (every thread loads one record, __syncthreads, everybody consumes) couples the 8 pixel
// warps of a tile at every batch: ncu attributed 22 % of all warp-stall samples of the backward kernel to
// those CTA barriers (warps cover sub-tiles with different amounts of surviving work).  Here a 9th warp
// is a dedicated PRODUCER: its lanes read the list entries and issue asynchronous 16-byte copies
// (cp.async / LDGSTS, completion counted on an mbarrier) of the 48-byte / 32-byte raster records
// into a ring of shared-memory slots; the 8 CONSUMER warps wait on a slot's "full" mbarrier, work through
// it at their own pace and arrive on its "empty" mbarrier.  No __syncthreads in the steady state; a fast
// warp can run ahead of a slow one by the depth of the ring.
#pragma once
#include "ts2d_fast.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// Ampere-style asynchronous 16-byte copy global -> shared (LDGSTS, L2-only caching) and its mbarrier hook: the
// arrive fires once all earlier cp.async of the executing thread have landed (.noinc: the arrival is pre-counted in
// the barrier's init count).  Measured on B200: for 48/32-byte raster records the LDGSTS path sustains the ring while
// one cp.async.bulk (TMA) per record does not -- three CTAs per SM x 256 tiny bulk copies per batch serialise in the
// TMA unit and the consumer warps starve (profiles/README.md, "TMA vs LDGSTS staging").
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA 1-D bulk copy global -> shared; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Out-of-line copy of the exact pair evaluation for the rare in-band fallback of the fast kernels (powf + expf + two IEEE
// divides are ~250 instructions: inlined they would sit in the middle of the hot loop's instruction-cache footprint).
static __device__ __noinline__ bool eval_exact_slow(const float4 e1, float v3x, float v3y, float area2, float op, float two_gamma, float px,
                                                    float py, PairEval &e)
{
    return eval_exact(e1.x, e1.y, e1.z, e1.w, v3x, v3y, area2, op, two_gamma, px, py, e);
}

// Conservative coverage test of one triangle against ONE sub-tile rectangle [x0, x0+7] x [y0, y0+3] (pixel offsets
// from the tile origin (ox, oy)); same margins as subtile_mask() in ts2d_fast.cuh.  r0 = {v1, v2}, r1 = {v3, 1/area2, op}.
static __device__ __noinline__ bool subtile_covers(const float4 r0, const float4 r1, float ox, float oy, float x0, float y0, const GammaK gk)
{
    const float inv = r1.z, op = r1.w;
    const float p1x = r0.x - ox, p1y = r0.y - oy, p2x = r0.z - ox, p2y = r0.w - oy, p3x = r1.x - ox, p3y = r1.y - oy;
    const float a10 = (p2x * p3y - p2y * p3x) * inv;
    const float a20 = (p3x * p1y - p3y * p1x) * inv;
    const float A1 = (r0.w - r1.y) * inv, B1 = (r1.x - r0.z) * inv;
    const float A2 = (r1.y - r0.y) * inv, B2 = (r0.x - r1.x) * inv;
    const float ext = fmaxf(fmaxf(fmaxf(fabsf(p1x), fabsf(p1y)), fmaxf(fabsf(p2x), fabsf(p2y))), fmaxf(fabsf(p3x), fabsf(p3y))) + 16.0f;
    const float err_a = 1.0e-6f * (ext * ext * fabsf(inv)) + 1.0e-6f;
    const float L = 2.0f * __logf(255.0f * op);
    if (!(L > -1.0e-3f) || !(fabsf(inv) < 3.0e37f)) return false;
    const float Lp = fmaxf(L, 1.0e-6f);
    float E = gk.is_one ? sqrtf(Lp) : exp2f(__log2f(Lp) * gk.inv_two_gamma);
    E = fminf(E, 10.0f);
    const float thr = (1.0f - (E * 1.002f + 2.0e-3f + 12.0f * err_a)) * (1.0f / 3.0f);
    const float A3 = -A1 - A2, B3 = -B1 - B2, a30 = 1.0f - a10 - a20;
    const float x1 = x0 + 7.0f, y1 = y0 + 3.0f;
    const float m1 = a10 + fmaxf(A1 * x0, A1 * x1) + fmaxf(B1 * y0, B1 * y1);
    const float m2 = a20 + fmaxf(A2 * x0, A2 * x1) + fmaxf(B2 * y0, B2 * y1);
    const float m3 = a30 + fmaxf(A3 * x0, A3 * x1) + fmaxf(B3 * y0, B3 * y1);
    return (m1 >= thr) && (m2 >= thr) && (m3 >= thr);
}
