// ts2d_fast.cuh -- shared pieces of the fast composite kernels (forward and backward).
//
// Fast path.  Per pixel the barycentrics are formed exactly like the reference (pixel-relative
// cross products, R2D/src/forward.cu:299-305) but with one multiply by a per-triangle reciprocal
// instead of two IEEE divides, ecc^(2 gamma) by ecc*ecc (gamma == 1) or MUFU lg2/ex2, and exp by
// MUFU ex2.  That keeps alpha within a few 1e-6 (relative) of the reference's value; every
// comparison the reference makes (ecc range, alpha < 1/255, op*G < 0.99, T <= 1e-4) is taken from the
// fast value only when it is further from the threshold than a rigorous rounding bound ("decision
// band"); inside the band the pair -- or the pixel's transmittance -- is re-evaluated with the
// op-for-op mirror of the reference arithmetic (eval_exact, ts2d_common.cuh).  So the set of
// contributing (pixel, triangle) pairs, n_contrib and the early-termination point are the
// reference's, while values differ by fp32 rounding only.
//
// Sub-tile culling.  Each warp owns an 8x4-pixel sub-tile.  At emission (ts2d_binning.cu) every instance's
// alpha >= 1/255 footprint (the triangle scaled about its centroid by E = (2 ln(255 op))^(1/(2 gamma)),
// clipped to ecc <= 10) is tested against the tile's 8 sub-tile rectangles with a conservative 3-edge + bounding-box
// test on affine edge functions; the 8-bit mask travels in the low bits of the instance key through the tile sort,
// and a warp only gathers list entries whose bit is set.  Margins cover the rounding of both the affine form
// and the reference's form, so culled pairs are exactly pairs the reference would have skipped.
#pragma once
#include <string.h>

#include "ts2d_common.cuh"

#define TS2D_LOG2E 1.4426950408889634f

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Shared memory through explicit 32-bit shared-space addresses.  ptxas (sm_100a) lowers "address of a __shared__ symbol" to
// S2R SR_CgaCtaId + MOV + LEA (the shared::cluster window) and, under a tight register bound, re-materialises that triple in
// front of every access inside the walk loops (3 issue slots + an S2R scoreboard wait per access group).  smem_base() pins
// the address in one register; all hot accesses are ld/st.shared relative to it.
__device__ __forceinline__ uint32_t smem_base(const void *p)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("mov.u32 %0, %0;" : "+r"(a));  // opaque: cannot be re-derived from the symbol
    return a;
}
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds32f(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// streaming (evict-first) 16-byte store
__device__ __forceinline__ void st_stream128(float4 *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Packed fp32: sm_100 executes add/mul/fma.rn.f32x2 as ONE FADD2 / FMUL2 / FFMA2 issue slot for two independent IEEE fp32
// operations; an operand may be a register pair, one register broadcast to both halves (bc()), or an immediate -- ptxas folds
// the pack / broadcast into the operand form, no MOVs.  Used where two sums share their multiplier (the issue-bound kernels).
struct v2 {
    float a, b;
};
__device__ __forceinline__ v2 mk2v(float a, float b)
{
    v2 r;
    r.a = a;
    r.b = b;
    return r;
}
__device__ __forceinline__ v2 bc(float x) { return mk2v(x, x); }
__device__ __forceinline__ v2 fma2(v2 x, v2 y, v2 z)
{
    v2 d;
    asm("{.reg .b64 rx, ry, rz, rd;\n mov.b64 rx, {%2, %3};\n mov.b64 ry, {%4, %5};\n mov.b64 rz, {%6, %7};\n"
        " fma.rn.f32x2 rd, rx, ry, rz;\n mov.b64 {%0, %1}, rd;}"
        : "=f"(d.a), "=f"(d.b)
        : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b), "f"(z.a), "f"(z.b));
    return d;
}
__device__ __forceinline__ v2 mul2(v2 x, v2 y)
{
    v2 d;
    asm("{.reg .b64 rx, ry, rd;\n mov.b64 rx, {%2, %3};\n mov.b64 ry, {%4, %5};\n mul.rn.f32x2 rd, rx, ry;\n mov.b64 {%0, %1}, rd;}"
        : "=f"(d.a), "=f"(d.b)
        : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b));
    return d;
}
__device__ __forceinline__ v2 add2(v2 x, v2 y)
{
    v2 d;
    asm("{.reg .b64 rx, ry, rd;\n mov.b64 rx, {%2, %3};\n mov.b64 ry, {%4, %5};\n add.rn.f32x2 rd, rx, ry;\n mov.b64 {%0, %1}, rd;}"
        : "=f"(d.a), "=f"(d.b)
        : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b));
    return d;
}
__device__ __forceinline__ v2 sub2(v2 x, v2 y)
{
    v2 d;
    asm("{.reg .b64 rx, ry, rd;\n mov.b64 rx, {%2, %3};\n mov.b64 ry, {%4, %5};\n sub.rn.f32x2 rd, rx, ry;\n mov.b64 {%0, %1}, rd;}"
        : "=f"(d.a), "=f"(d.b)
        : "f"(x.a), "f"(x.b), "f"(y.a), "f"(y.b));
    return d;
}

// gamma-dependent constants (kernel-uniform)
#define TS2D_ARGMIN_TIE 2.7e-6f  // 2^-22 * (1 + sum |a_i|), sum |a_i| <= 10: see the arg-min note in ts2d_render_bwd_fast.cu

struct GammaK {
    float gamma, two_gamma, inv_two_gamma;
    float band;      // relative half-width of the alpha decision band
    float terr_c0;   // rel. error bound of a fast alpha = terr_c0 + terr_c1 * |power|
    float terr_c1;
    bool is_one;     // gamma == 1: ecc^(2 gamma) = ecc*ecc
};

// Error model (eps = 2^-24 = 6e-8).  The fast a_i differ from the reference's by <= 1.5 ulp (reciprocal-multiply
// vs IEEE divide) => |d ecc| <= 3 * 1.5 eps * (1 + |a1| + |a2|) <= 1.2e-6 wherever a pair can contribute
// (all a_i in [-0.8, 2.6]).  d ln(alpha) = gamma * ecc^(2 gamma - 1) * d ecc <= gamma * (1 + 2|power|) * d ecc
// (gamma >= 0.5).  powf is within 4 ulp of ecc^(2 gamma), ecc*ecc within 0.5 ulp => d power <= 2.4e-7 |power|;
// the lg2/ex2 route adds 2 gamma * 2^-22 * ln 2 relative on the power; expf vs ex2.approx (+ argument rounding)
// adds <= 4e-7.  band is the bound evaluated at the alpha = 1/255 threshold with op = 1 (|power| = 5.54).
__device__ __forceinline__ GammaK make_gamma(float gamma)
{
    GammaK g;
    g.gamma = gamma;
    g.two_gamma = 2.0f * gamma;
    g.inv_two_gamma = gamma > 0.0f ? 1.0f / (2.0f * gamma) : 0.0f;
    g.is_one = (gamma == 1.0f);
    const float lg = g.is_one ? 0.0f : 2.0e-7f * gamma;  // log2f (1 ulp of |log2 ecc| <= 3.4) * 2 gamma * ln 2 + exp2f (2 ulp)
    g.terr_c0 = 1.8e-6f * gamma + 4.0e-7f;
    g.terr_c1 = 3.6e-6f * gamma + 2.4e-7f + lg;
    g.band = 1.5f * (g.terr_c0 + g.terr_c1 * 5.6f) + 2.0e-6f;
    return g;
}

// Sub-tile coverage mask of one triangle in the tile with origin (ox, oy).  r0 = {v1, v2}, r1 = {v3, area2, op}.
// warp w <-> sub-tile (sx = w & 1, sy = w >> 1), pixels dx in [8sx, 8sx+7], dy in [4sy, 4sy+3].
__device__ __forceinline__ uint32_t subtile_mask(const float4 r0, const float4 r1, float inv, float ox, float oy, const GammaK gk)
{
    const float op = r1.w;
    const float p1x = r0.x - ox, p1y = r0.y - oy, p2x = r0.z - ox, p2y = r0.w - oy, p3x = r1.x - ox, p3y = r1.y - oy;
    const float a10 = (p2x * p3y - p2y * p3x) * inv;
    const float a20 = (p3x * p1y - p3y * p1x) * inv;
    const float A1 = (r0.w - r1.y) * inv, B1 = (r1.x - r0.z) * inv;  // d a1 / d(px, py)
    const float A2 = (r1.y - r0.y) * inv, B2 = (r0.x - r1.x) * inv;  // d a2 / d(px, py)
    // rounding bound on any barycentric inside this tile, for the affine form and for the reference's form
    const float ext = fmaxf(fmaxf(fmaxf(fabsf(p1x), fabsf(p1y)), fmaxf(fabsf(p2x), fabsf(p2y))), fmaxf(fabsf(p3x), fabsf(p3y))) + 16.0f;
    const float err_a = 1.0e-6f * (ext * ext * fabsf(inv)) + 1.0e-6f;  // ~8 ulp of 2 ext^2 / |area2|
    // alpha >= 1/255  <=>  ecc^(2 gamma) <= L = 2 ln(255 op)
    const float L = 2.0f * __logf(255.0f * op);
    if (!(L > -1.0e-3f) || !(fabsf(inv) < 3.0e37f)) return 0u;  // opacity below 1/255 (with margin): contributes nowhere
    const float Lp = fmaxf(L, 1.0e-6f);
    float E = gk.is_one ? sqrtf(Lp) : exp2f(__log2f(Lp) * gk.inv_two_gamma);
    E = fminf(E, 10.0f);
    const float Ec = E * 1.002f + 2.0e-3f + 12.0f * err_a;  // conservative footprint radius in ecc units
    const float thr = (1.0f - Ec) * (1.0f / 3.0f);           // need min(a1,a2,a3) >= thr somewhere in the rect
    const float A3 = -A1 - A2, B3 = -B1 - B2, a30 = 1.0f - a10 - a20;
    float mx1[2], mx2[2], mx3[2], my1[4], my2[4], my3[4];
#pragma unroll
    for (int sx = 0; sx < 2; sx++) {
        const float lo = 8.0f * sx, hi = lo + 7.0f;
        mx1[sx] = fmaxf(A1 * lo, A1 * hi);
        mx2[sx] = fmaxf(A2 * lo, A2 * hi);
        mx3[sx] = fmaxf(A3 * lo, A3 * hi);
    }
#pragma unroll
    for (int sy = 0; sy < 4; sy++) {
        const float lo = 4.0f * sy, hi = lo + 3.0f;
        my1[sy] = fmaxf(B1 * lo, B1 * hi);
        my2[sy] = fmaxf(B2 * lo, B2 * hi);
        my3[sy] = fmaxf(B3 * lo, B3 * hi);
    }
    // The footprint {min a_i >= thr} is the triangle scaled by Ec about its centroid c: vertices c + Ec (v_i - c).  The three edge
    // tests above separate a rectangle from it only along the edge normals; its bounding box adds the two axis-aligned separating
    // directions (rectangles that sit beyond an acute corner pass all three edge tests but can never be touched).
    const float cx = (p1x + p2x + p3x) * (1.0f / 3.0f), cy = (p1y + p2y + p3y) * (1.0f / 3.0f);
    const float bx0 = fmaf(Ec, fminf(fminf(p1x, p2x), p3x) - cx, cx) - 0.02f, bx1 = fmaf(Ec, fmaxf(fmaxf(p1x, p2x), p3x) - cx, cx) + 0.02f;
    const float by0 = fmaf(Ec, fminf(fminf(p1y, p2y), p3y) - cy, cy) - 0.02f, by1 = fmaf(Ec, fmaxf(fmaxf(p1y, p2y), p3y) - cy, cy) + 0.02f;
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int sx = w & 1, sy = w >> 1;
        const bool ok = (a10 + mx1[sx] + my1[sy] >= thr) && (a20 + mx2[sx] + my2[sy] >= thr) && (a30 + mx3[sx] + my3[sy] >= thr) &&
                        (bx0 <= 8.0f * sx + 7.0f) && (bx1 >= 8.0f * sx) && (by0 <= 4.0f * sy + 3.0f) && (by1 >= 4.0f * sy);
        m |= ok ? (1u << w) : 0u;
    }
    return m;
}

// Fast per-pair evaluation with absolute pixel coordinates.  e1 = {v1.x, v1.y, v2.x, v2.y}, e2 = {v3.x, v3.y, 1/area2, op}.
// Returns whether the pair contributes according to the fast value; `uncertain` says the reference's decision
// cannot be inferred from it (then the caller uses eval_exact).
struct FastPair {
    float a1, a2, a3, ecc, pw, power, G, og, alpha;  // pw = ecc^(2 gamma), power = -pw / 2
    float pv1x, pv1y, pv2x, pv2y, pv3x, pv3y;
};

__device__ __forceinline__ bool eval_fast(const float4 e1, const float4 e2, float px, float py, const GammaK gk, FastPair &f, bool &uncertain)
{
    {   // the three vertex - pixel differences: each (x, y) pair is one packed subtract
        const v2 p = mk2v(px, py);
        const v2 d1 = sub2(mk2v(e1.x, e1.y), p), d2 = sub2(mk2v(e1.z, e1.w), p), d3 = sub2(mk2v(e2.x, e2.y), p);
        f.pv1x = d1.a; f.pv1y = d1.b;
        f.pv2x = d2.a; f.pv2y = d2.b;
        f.pv3x = d3.a; f.pv3y = d3.b;
    }
    f.a1 = fmaf(f.pv2x, f.pv3y, -(f.pv2y * f.pv3x)) * e2.z;
    f.a2 = fmaf(f.pv3x, f.pv1y, -(f.pv3y * f.pv1x)) * e2.z;
    f.a3 = 1.0f - f.a1 - f.a2;
    f.ecc = fmaf(fminf(fminf(f.a1, f.a2), f.a3), -3.0f, 1.0f);
    float pw;
    if (gk.is_one)
        pw = f.ecc * f.ecc;
    else
        pw = exp2f(gk.two_gamma * log2f(fmaxf(f.ecc, 1.0e-30f)));  // full-precision log2f/exp2f: the error is multiplied by 2 gamma |power|
    f.pw = pw;
    f.power = -0.5f * pw;  // (dead code where the caller works from pw)
    // exp(power) = 2^(pw * (-log2(e) / 2)): the same bits as 2^(power * log2(e)) -- the factor 1/2 is a power of two, it commutes with
    // the rounding of the product -- in one multiply instead of two
    f.G = ex2_approx(pw * (-0.5f * TS2D_LOG2E));
    f.og = e2.w * f.G;
    f.alpha = fminf(0.99f, f.og);
    const float d = fmaf(f.alpha, 255.0f, -1.0f);  // alpha * 255 - 1
    uncertain = (fabsf(d) <= gk.band) || (f.ecc < 1.0e-4f);
    // The reference's ecc > 10 cut needs no test here: the fast kernels only run for gamma >= 0.6 (ts2d_use_fast), where
    // ecc >= 9.9 implies alpha <= exp(-0.5 * 9.9^1.2) << 1/255, i.e. d < -band: skipped by both.
    return d >= 0.0f;
}
