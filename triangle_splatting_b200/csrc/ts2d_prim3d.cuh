// ts2d_prim3d.cuh -- per-pair arithmetic of the 3D primitive (ray / triangle-plane intersection in view space), shared by the
// mirror kernels (ts2d_prim3d.cu), the fast kernels (ts2d_prim3d_fast.cu) and the emission kernel (ts2d_binning.cu).
//
// Reference: R3D = submodules/diff-triangle-rasterization-3D; R3D/src/forward.cu:243-276, R3D/src/backward.cu:330-352.
//
// Conditioning.  The reference forms the barycentrics from view-space differences pv_k = v_k - depth * ray: with view-space
// coordinates of size Z and triangles of size s the subtraction cancels log2(Z / s) bits, so the reference's own a_i carry a
// relative noise of ~20 eps Z / s (1e-5 .. 1e-3 in practice) -- far above the 1e-5 pixel bar.  A differently rounded formula
// can therefore NOT reproduce the reference's pixels; the only way to meet the bar is to execute the reference's geometry
// arithmetic operation for operation.  geo3() does that (same expression trees as the reference, so the same FMA contraction;
// the mirror kernels, whose forward outputs are bit-identical to the reference's, are built on the very same function), with
// the two per-triangle subexpressions dot(v1, n) and 1 / dot(n, n) taken from the raster record (computed once per triangle
// in K1 by the same expressions).  Only what FOLLOWS ecc (pow, exp) is replaced by fast arithmetic in the fast kernels, under
// the decision bands of ts2d_fast.cuh.
#pragma once
#include "ts2d_fast.cuh"

// R3D/src/auxiliary.h:35-43
__device__ __forceinline__ float proj_to_pix(float v, int S) { return (v + 1.0f) * S * 0.5f - 0.5f; }
__device__ __forceinline__ float pix_to_proj(float v, int S) { return (2.0f * v - S + 1.0f) / (float)(S); }

// Raster record of the 3D primitive (geometry state; 80 B of the 5 float4 per triangle):
//   rec0[3i+0] = {v1v.x v1v.y v1v.z v2v.x}   rec0[3i+1] = {v2v.y v2v.z v3v.x v3v.y}   rec0[3i+2] = {v3v.z n.x n.y n.z}
//   rec1[2i+0] = {r g b opacity}              rec1[2i+1] = {K_fwd, 1 / dot(n, n), K_bwd, -}   (K = dot(v1v, n) in the two roundings below)
// with v_k_view = W2C v_k and n = (v2v - v1v) x (v3v - v1v), NOT normalised (R3D/src/forward.cu:94).
struct Tri3 {
    f3 v1, v2, v3, n;
};
__device__ __forceinline__ Tri3 unpack3(const float4 a, const float4 b, const float4 c)
{
    Tri3 t;
    t.v1 = mk3(a.x, a.y, a.z);
    t.v2 = mk3(a.w, b.x, b.y);
    t.v3 = mk3(b.z, b.w, c.x);
    t.n = mk3(c.y, c.z, c.w);
    return t;
}
// The two per-triangle subexpressions of the per-pair arithmetic.  nvcc's FMA contraction of a 3-term dot product depends on
// the surrounding code: the reference's sm_100 build computes dot(v1, n) as fma(n.z, v1.z, fma(n.y, v1.y, n.x * v1.x)) in the
// forward kernel but as fma(v1.z, n.z, fma(v1.x, n.x, v1.y * n.y)) in the backward kernel, and dot(n, n) as
// fma(n.z, n.z, fma(n.x, n.x, n.y * n.y)) in both (cuobjdump -sass of the reference's own sm_100 build, FORWARD::renderCUDA /
// BACKWARD::renderCUDA).  Everything here is therefore written with explicit round-to-nearest intrinsics.
template <bool BWD>
__device__ __forceinline__ float tri3_K(const Tri3 &t)
{
    if (BWD) return __fmaf_rn(t.v1.z, t.n.z, __fmaf_rn(t.v1.x, t.n.x, __fmul_rn(t.v1.y, t.n.y)));
    return __fmaf_rn(t.n.z, t.v1.z, __fmaf_rn(t.n.y, t.v1.y, __fmul_rn(t.n.x, t.v1.x)));
}
__device__ __forceinline__ float tri3_inv_nn(const Tri3 &t)
{
    return __frcp_rn(__fmaf_rn(t.n.z, t.n.z, __fmaf_rn(t.n.x, t.n.x, __fmul_rn(t.n.y, t.n.y))));
}

struct Pair3 {
    float depth, inv_pn, inv_nn, a1, a2, a3, ecc, power, G, alpha;
    f3 pv1, pv2, pv3;
};

// dot(c, n) as the reference's build contracts it: fma(n.z, c.z, fma(n.x, c.x, n.y * c.y))
__device__ __forceinline__ float dot3_ref(f3 c, f3 n) { return __fmaf_rn(n.z, c.z, __fmaf_rn(n.x, c.x, __fmul_rn(n.y, c.y))); }
// cross(a, b) as the reference's build contracts it: every component fma(first product, -(second product))
__device__ __forceinline__ f3 cross3_ref(f3 a, f3 b)
{
    return mk3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

// Geometry half of a pair: everything up to ecc, in the reference's arithmetic (R3D/src/forward.cu:243-270,
// R3D/src/backward.cu:330-346) with the contraction of its sm_100 SASS.  K = tri3_K<BWD>(t), inv_nn = tri3_inv_nn(t).
// BWD: the backward kernel divides differently (depth = K * (1 / pn), R3D/src/backward.cu:334-335, vs K / pn in forward.cu:246).
// ray.z must be 1 (R3D/src/forward.cu:187).  Returns false where the reference `continue`s (|ray . n| < eps, ecc outside [0, 10]).
template <bool BWD>
__device__ __forceinline__ bool geo3(const Tri3 &t, float K, float inv_nn, f3 ray, Pair3 &e)
{
    const float pn = __fadd_rn(t.n.z, __fmaf_rn(ray.x, t.n.x, __fmul_rn(ray.y, t.n.y)));
    if (fabsf(pn) < TS2D_EPS) return false;
    if (BWD) {
        e.inv_pn = __frcp_rn(pn);
        e.depth = __fmul_rn(K, e.inv_pn);
    } else {
        e.depth = __fdiv_rn(K, pn);
    }
    e.pv1 = mk3(__fmaf_rn(-ray.x, e.depth, t.v1.x), __fmaf_rn(-ray.y, e.depth, t.v1.y), __fsub_rn(t.v1.z, e.depth));
    e.pv2 = mk3(__fmaf_rn(-ray.x, e.depth, t.v2.x), __fmaf_rn(-ray.y, e.depth, t.v2.y), __fsub_rn(t.v2.z, e.depth));
    e.pv3 = mk3(__fmaf_rn(-ray.x, e.depth, t.v3.x), __fmaf_rn(-ray.y, e.depth, t.v3.y), __fsub_rn(t.v3.z, e.depth));
    e.inv_nn = inv_nn;
    e.a1 = __fmul_rn(dot3_ref(cross3_ref(e.pv2, e.pv3), t.n), inv_nn);
    e.a2 = __fmul_rn(dot3_ref(cross3_ref(e.pv3, e.pv1), t.n), inv_nn);
    e.a3 = __fsub_rn(__fsub_rn(1.0f, e.a1), e.a2);
    e.ecc = __fmaf_rn(fminf(fminf(e.a1, e.a2), e.a3), -3.0f, 1.0f);
    return !(e.ecc < 0.0f || e.ecc > 10.0f);
}

// Opacity half in the reference's arithmetic (libdevice powf / expf).  Forward skips on alpha < 1/255, backward on G < 1/255.
template <bool BWD>
__device__ __forceinline__ bool alpha3_exact(float op, float two_gamma, Pair3 &e)
{
    e.power = __fmul_rn(-0.5f, powf(e.ecc, two_gamma));
    e.G = expf(e.power);
    e.alpha = fminf(0.99f, __fmul_rn(op, e.G));
    if (BWD) return !(e.G < 1.0f / 255.0f);
    return !(e.alpha < 1.0f / 255.0f);
}

// The whole pair, op for op (R3D/src/forward.cu:243-276 / R3D/src/backward.cu:330-352).
template <bool BWD>
__device__ __forceinline__ bool eval_pair3(const Tri3 &t, float K, float inv_nn, float op, float two_gamma, f3 ray, Pair3 &e)
{
    if (!geo3<BWD>(t, K, inv_nn, ray, e)) return false;
    return alpha3_exact<BWD>(op, two_gamma, e);
}

// Fast opacity half: ecc^(2 gamma) by ecc * ecc (gamma == 1) or log2f / exp2f, exp by MUFU.EX2.  ecc is the reference's own
// value here, so the fast alpha differs from the reference's only by the pow / exp implementations: relative error
// <= c0 + c1 |power| with c0 = 4e-7 (expf vs ex2.approx incl. argument rounding) and c1 = 2.4e-7 (powf <= 4 ulp) + the
// log2f / exp2f route's 2e-7 gamma.  `unc`: a reference decision (the 1/255 cut on alpha [forward] or on G [backward]) cannot
// be inferred from the fast value; the caller then runs alpha3_exact().
struct Gamma3 {
    float two_gamma, c0, c1, band;
    bool is_one;
};
__device__ __forceinline__ Gamma3 make_gamma3(float gamma, bool is_one)
{
    Gamma3 g;
    g.two_gamma = 2.0f * gamma;
    g.is_one = is_one;
    g.c0 = 4.0e-7f;
    g.c1 = 2.4e-7f + (is_one ? 0.0f : 2.0e-7f * gamma);
    g.band = 2.0f * (g.c0 + g.c1 * 5.6f) + 1.0e-6f;  // relative half-width at the 1/255 cut (|power| <= 5.55 there), x2 margin
    return g;
}
template <bool BWD>
__device__ __forceinline__ bool alpha3_fast(float op, const Gamma3 gk, Pair3 &e, float &og, bool &unc)
{
    float pw;
    if (gk.is_one)
        pw = e.ecc * e.ecc;
    else
        pw = exp2f(gk.two_gamma * log2f(fmaxf(e.ecc, 1.0e-30f)));
    e.power = -0.5f * pw;
    e.G = ex2_approx(e.power * TS2D_LOG2E);
    og = op * e.G;
    e.alpha = fminf(0.99f, og);
    const float d = fmaf(BWD ? e.G : e.alpha, 255.0f, -1.0f);
    unc = (fabsf(d) <= gk.band) || (e.ecc < 1.0e-4f);  // ecc -> 0: pow / log2 of a denormal-sized argument, leave it to libdevice
    return d >= 0.0f;
}

// ------------------------------------------------------------------------------------------------ sub-tile coverage (3D)
// Conservative 8-bit coverage mask of one instance, same role as subtile_mask() of the 2D primitive.
//
// A pair can be visited by either pass only if G >= 1/255 (the backward's cut; the forward's alpha = min(0.99, op G) >= 1/255
// implies it), i.e. ecc <= E = (2 ln 255)^(1 / (2 gamma)) in the REFERENCE's arithmetic, i.e. min_i a_i >= (1 - E) / 3.
// The barycentrics of the ray / plane intersection are projective in the pixel: a_i = N_i / D with N_i = ray . m_i,
// D = ray . n, m_2 = e3 x v1, m_3 = v1 x e2, m_1 = n - m_2 - m_3 (e_k = v_k - v1).  Where D keeps one sign s over the tile,
//   a_i >= thr  <=>  s * ray . (m_i - thr n) >= 0,
// an affine function of the pixel whose maximum over a sub-tile rectangle sits at a corner -- the same corner test as in 2D.
// Margins: the affine forms are evaluated with error <= 8 eps |ray| (|v||e| + |n|) each; the reference's a_i (cancelling
// subtraction, see the header) carry err_ref (below), which widens the footprint: Ec = E + 12 err_ref.  Where D changes sign
// inside the tile (the triangle's plane passes through the eye inside this tile's cone) every sub-tile is kept.
__device__ __forceinline__ uint32_t subtile_mask3d(const Tri3 &t, float ox, float oy, int W, int H, float tfx, float tfy, float gamma, bool gamma_is_one)
{
    const f3 e2 = t.v2 - t.v1, e3 = t.v3 - t.v1;
    const f3 m2 = cross3(e3, t.v1), m3 = cross3(t.v1, e2);
    const float n_len = len3(t.n);
    if (!(n_len > 0.0f) || !(n_len < 3.0e37f)) return 0xFFu;
    const float Lsq = 2.0f * 5.5412635f;  // 2 ln 255
    float E = gamma_is_one ? sqrtf(Lsq) : exp2f(__log2f(Lsq) * (0.5f / gamma));
    E = fminf(E, 10.0f);
    const float zm = fmaxf(fmaxf(fabsf(t.v1.x) + fabsf(t.v1.y) + fabsf(t.v1.z), fabsf(t.v2.x) + fabsf(t.v2.y) + fabsf(t.v2.z)),
                           fabsf(t.v3.x) + fabsf(t.v3.y) + fabsf(t.v3.z));
    const f3 e23 = e3 - e2;
    const float emax = sqrtf(fmaxf(fmaxf(dot3(e2, e2), dot3(e3, e3)), dot3(e23, e23)));
    const float L = (1.0f + E) * emax;
    // ray at the tile origin and its per-pixel steps (ray.z = 1)
    const float sx = 2.0f * tfx / (float)W, sy = 2.0f * tfy / (float)H;
    const float rx0 = tfx * pix_to_proj(ox, W), ry0 = tfy * pix_to_proj(oy, H);
    const float rmax = 1.0f + fabsf(rx0) + fabsf(ry0) + 16.0f * (fabsf(sx) + fabsf(sy));
    const float nl1 = fabsf(t.n.x) + fabsf(t.n.y) + fabsf(t.n.z);
    // D over the tile
    const float D0 = fmaf(rx0, t.n.x, fmaf(ry0, t.n.y, t.n.z)), Dx = sx * t.n.x, Dy = sy * t.n.y;
    const float Dmin = D0 + fminf(0.0f, 15.0f * Dx) + fminf(0.0f, 15.0f * Dy), Dmax = D0 + fmaxf(0.0f, 15.0f * Dx) + fmaxf(0.0f, 15.0f * Dy);
    const float err_D = 8.0f * 6.0e-8f * rmax * nl1;
    if (!(Dmin > err_D) && !(Dmax < -err_D)) return 0xFFu;  // sign of D not definite over the tile (also catches NaN)
    const float s = Dmin > 0.0f ? 1.0f : -1.0f;
    const float Dabs = fminf(fabsf(Dmin), fabsf(Dmax)) - err_D;
    // error of the reference's a_i: (i) rounding of pv_k = v_k - depth ray and of the cross / dot products, (ii) the error of
    // depth = dot(v1, n) / (ray . n) -- dot(v1, n) cancels for planes seen at a grazing angle -- which slides the
    // intersection point along the ray: |d a| <= |d depth| |e| |ray| / |n|
    const float err_ref = 64.0f * 6.0e-8f * (zm + L) * L / n_len + 24.0f * 6.0e-8f * zm * emax * rmax * (1.0f + rmax) / Dabs;
    const float Ec = E * 1.002f + 2.0e-3f + 12.0f * err_ref;
    const float thr = (1.0f - Ec) * (1.0f / 3.0f);
    const float el1 = fabsf(e2.x) + fabsf(e2.y) + fabsf(e2.z) + fabsf(e3.x) + fabsf(e3.y) + fabsf(e3.z);
    const float err_g = 16.0f * 6.0e-8f * rmax * (zm * el1 + (1.0f + fabsf(thr)) * nl1) + 1.0e-37f;
    // c_i = s (m_i - thr n)
    const f3 c2 = s * (m2 - thr * t.n), c3 = s * (m3 - thr * t.n);
    const f3 c1 = s * ((1.0f - thr) * t.n) - s * m2 - s * m3;
    const float g10 = fmaf(rx0, c1.x, fmaf(ry0, c1.y, c1.z)), A1 = sx * c1.x, B1 = sy * c1.y;
    const float g20 = fmaf(rx0, c2.x, fmaf(ry0, c2.y, c2.z)), A2 = sx * c2.x, B2 = sy * c2.y;
    const float g30 = fmaf(rx0, c3.x, fmaf(ry0, c3.y, c3.z)), A3 = sx * c3.x, B3 = sy * c3.y;
    float mx1[2], mx2[2], mx3[2], my1[4], my2[4], my3[4];
#pragma unroll
    for (int ix = 0; ix < 2; ix++) {
        const float lo = 8.0f * ix, hi = lo + 7.0f;
        mx1[ix] = fmaxf(A1 * lo, A1 * hi);
        mx2[ix] = fmaxf(A2 * lo, A2 * hi);
        mx3[ix] = fmaxf(A3 * lo, A3 * hi);
    }
#pragma unroll
    for (int iy = 0; iy < 4; iy++) {
        const float lo = 4.0f * iy, hi = lo + 3.0f;
        my1[iy] = fmaxf(B1 * lo, B1 * hi);
        my2[iy] = fmaxf(B2 * lo, B2 * hi);
        my3[iy] = fmaxf(B3 * lo, B3 * hi);
    }
    // bounding box of the footprint's projection (the triangle scaled by Ec about its centroid), when it lies in front of the eye
    float bx0 = -3.0e38f, bx1 = 3.0e38f, by0 = -3.0e38f, by1 = 3.0e38f;
    {
        const f3 c = (t.v1 + t.v2 + t.v3) * (1.0f / 3.0f);
        const f3 q1 = c + Ec * (t.v1 - c), q2 = c + Ec * (t.v2 - c), q3 = c + Ec * (t.v3 - c);
        const float zmin = fminf(fminf(q1.z, q2.z), q3.z);
        if (zmin > 1.0e-3f * zm) {
            const float i1 = 1.0f / q1.z, i2 = 1.0f / q2.z, i3 = 1.0f / q3.z;
            // pixel of a view-space direction (x/z, y/z): px = ((x/z) / tfx * W + W - 1) / 2, relative to the tile origin
            const float kx = 0.5f * (float)W / tfx, ky = 0.5f * (float)H / tfy, cx0 = 0.5f * ((float)W - 1.0f) - ox, cy0 = 0.5f * ((float)H - 1.0f) - oy;
            const float x1 = fmaf(q1.x * i1, kx, cx0), x2 = fmaf(q2.x * i2, kx, cx0), x3 = fmaf(q3.x * i3, kx, cx0);
            const float y1 = fmaf(q1.y * i1, ky, cy0), y2 = fmaf(q2.y * i2, ky, cy0), y3 = fmaf(q3.y * i3, ky, cy0);
            const float mg = 0.05f + 2.0e-6f * ((float)W + (float)H) + 1.0e-5f * (fabsf(x1) + fabsf(x2) + fabsf(x3) + fabsf(y1) + fabsf(y2) + fabsf(y3));
            bx0 = fminf(fminf(x1, x2), x3) - mg;
            bx1 = fmaxf(fmaxf(x1, x2), x3) + mg;
            by0 = fminf(fminf(y1, y2), y3) - mg;
            by1 = fmaxf(fmaxf(y1, y2), y3) + mg;
        }
    }
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int ix = w & 1, iy = w >> 1;
        const bool ok = (g10 + mx1[ix] + my1[iy] >= -err_g) && (g20 + mx2[ix] + my2[iy] >= -err_g) && (g30 + mx3[ix] + my3[iy] >= -err_g) &&
                        (bx0 <= 8.0f * ix + 7.0f) && (bx1 >= 8.0f * ix) && (by0 <= 4.0f * iy + 3.0f) && (by1 >= 4.0f * iy);
        m |= ok ? (1u << w) : 0u;
    }
    return m;
}
