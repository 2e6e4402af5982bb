// ts2d_sh.cuh -- spherical-harmonics colour (forward + backward) and the warp-cooperative row staging shared by the
// per-triangle kernels of both primitives (ts2d_preprocess.cu: 2D, ts2d_prim3d.cu: 3D).
// Replaces computeRGBFromSH / computeRGBFromSHBackward (R2D/src/forward.cu:9-59, R2D/src/backward.cu:9-119; the 3D package
// carries byte-identical copies, R3D/src/forward.cu:9-59, R3D/src/backward.cu:9-119).
#pragma once
#include "ts2d_common.cuh"

static __constant__ float kC0 = 0.28209479177387814f;
static __constant__ float kC1 = 0.4886025119029199f;
static __constant__ float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f};
static __constant__ float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                             -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

__device__ __forceinline__ f3 ld3(const float *p) { return mk3(p[0], p[1], p[2]); }

// ---- warp-cooperative staging of per-triangle rows (SH coefficients / SH gradients) ----------------------------
// A thread owns one triangle but its row is 3M floats (192 B at M = 16): per-thread scalar access puts 32 different
// cache lines behind every load/store instruction.  Instead the warp moves its 32 rows -- one contiguous 32*3M-float
// block -- with coalesced 16-byte accesses through a shared-memory tile whose row stride (3M + 4 floats) makes the
// per-thread float4 accesses bank-conflict free.  Used when 3M is a multiple of 4 (M = 4, 8, 12, 16).
__device__ __forceinline__ void warp_rows_load(const float *__restrict__ g, float *tile, int q /*float4 per row*/, int rs4 /*tile row stride in float4*/,
                                               int nrows, int lane)
{
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    float4 *t4 = reinterpret_cast<float4 *>(tile);
    int r = lane / q, c = lane - r * q;
    const int dr = 32 / q, dc = 32 - dr * q;
    for (int e = lane; e < nrows * q; e += 32) {
        t4[r * rs4 + c] = __ldg(g4 + e);
        r += dr;
        c += dc;
        if (c >= q) { c -= q; r++; }
    }
}
__device__ __forceinline__ void warp_rows_store(float *__restrict__ g, const float *tile, int q, int rs4, int nrows, int lane)
{
    float4 *g4 = reinterpret_cast<float4 *>(g);
    const float4 *t4 = reinterpret_cast<const float4 *>(tile);
    int r = lane / q, c = lane - r * q;
    const int dr = 32 / q, dc = 32 - dr * q;
    for (int e = lane; e < nrows * q; e += 32) {
        g4[e] = t4[r * rs4 + c];
        r += dr;
        c += dc;
        if (c >= q) { c -= q; r++; }
    }
}
// ---- parameter-space inputs (ts2d_model_inputs): the model's Python preamble done in registers -------------------
// Every helper keeps the operation order of the torch kernels the preamble launches (src/diff_recon/models/VanillaTS_model.py),
// with explicit round-to-nearest intrinsics so that no neighbouring code can be contracted into it: the values that reach the
// reference-order geometry code are bit-identical to the tensors the reference would have materialised.

// _rescale_triangles (:431-447): t_center = vertex.mean(dim=1, keepdim=True) -- torch's reduce kernel sums the three vertices in
// order in one thread and multiplies by float(1/3); (vertex - t_center) * ratio + t_center is three separate elementwise kernels.
__device__ __forceinline__ float rescale1(float v, float c, float ratio) { return __fadd_rn(__fmul_rn(__fsub_rn(v, c), ratio), c); }
template <bool MODEL>
__device__ __forceinline__ void load_tri(const float *vp, const ModelIn &mi, f3 &a, f3 &b, f3 &c)
{
    a = ld3(vp);
    b = ld3(vp + 3);
    c = ld3(vp + 6);
    if (MODEL && mi.ratio != 1.0f) {
        const float third = 1.0f / 3.0f;
        const f3 m = mk3(__fmul_rn(__fadd_rn(__fadd_rn(a.x, b.x), c.x), third), __fmul_rn(__fadd_rn(__fadd_rn(a.y, b.y), c.y), third),
                         __fmul_rn(__fadd_rn(__fadd_rn(a.z, b.z), c.z), third));
        a = mk3(rescale1(a.x, m.x, mi.ratio), rescale1(a.y, m.y, mi.ratio), rescale1(a.z, m.z, mi.ratio));
        b = mk3(rescale1(b.x, m.x, mi.ratio), rescale1(b.y, m.y, mi.ratio), rescale1(b.z, m.z, mi.ratio));
        c = mk3(rescale1(c.x, m.x, mi.ratio), rescale1(c.y, m.y, mi.ratio), rescale1(c.z, m.z, mi.ratio));
    }
}
// adjoint of the rescale map: g_i -> ratio * g_i + (1 - ratio) / 3 * (g_1 + g_2 + g_3)
__device__ __forceinline__ void rescale_bwd(const ModelIn &mi, f3 &g1, f3 &g2, f3 &g3)
{
    if (mi.ratio != 1.0f) {
        const f3 sum = (g1 + g2) + g3;
        const f3 gc = (sum - mi.ratio * sum) / 3.0f;
        g1 = mi.ratio * g1 + gc;
        g2 = mi.ratio * g2 + gc;
        g3 = mi.ratio * g3 + gc;
    }
}
// get_opacity (:84) = torch.sigmoid: 1 / (1 + exp(-x)) in fp32 with the accurate expf and an IEEE divide;
// straight-through binarisation (:620-621): ((opacity > thr).float() - opacity).detach() + opacity.
__device__ __forceinline__ float sigmoid_rn(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
__device__ __forceinline__ float model_opacity(const ModelIn &mi, int idx, float &sig)
{
    sig = sigmoid_rn(mi.logit[idx]);
    if (mi.ste_thr >= 0.0f) return __fadd_rn(__fsub_rn(sig > mi.ste_thr ? 1.0f : 0.0f, sig), sig);
    return sig;
}
// bg_depth = (camera_center - vertex).norm(dim=-1).max() (:623) over the UN-rescaled vertices of every triangle (culled or not).
// One atomic per warp at most, and only while the running maximum is still below the warp's.
__device__ __forceinline__ void bg_depth_max(const ModelIn &mi, const float *vp, bool valid, f3 cam)
{
    float m = 0.0f;
    if (valid) {
        const f3 a = cam - ld3(vp), b = cam - ld3(vp + 3), c = cam - ld3(vp + 6);
        m = fmaxf(fmaxf(len3(a), len3(b)), len3(c));
    }
    const uint32_t w = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if ((threadIdx.x & 31) == 0 && w > *(volatile uint32_t *)mi.bg_bits) atomicMax(mi.bg_bits, w);
}

// SH rows from the split parameters: tile row r = [f_dc[r] (3 floats) | f_rest[r] (3(M-1) floats)], row stride rsf floats (odd,
// so the per-thread scalar reads of sh_colour are bank-conflict free).  Both sources are contiguous 32-row blocks; the f_rest
// block starts 16-byte aligned (row0 is a multiple of 32) and is moved with 16-byte loads, scattered element-wise.
__device__ __forceinline__ void warp_rows_load_split(const float *__restrict__ f_dc, const float *__restrict__ f_rest, float *tile, int M,
                                                     int rsf, int row0, int nrows, int lane)
{
    for (int i = lane; i < 3 * nrows; i += 32) tile[(i / 3) * rsf + (i % 3)] = __ldg(f_dc + 3 * (size_t)row0 + i);
    const int L = 3 * (M - 1), n = nrows * L;
    if (L == 0) return;
    const float *src = f_rest + (size_t)row0 * L;
    const int n4 = (((uintptr_t)src & 15) == 0) ? n / 4 : 0;
    int r = (4 * lane) / L, c = 4 * lane - r * L;
    const int dr = 128 / L, dc = 128 - dr * L;
    for (int e = lane; e < n4; e += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + e);
        int rr = r, cc = c;
        tile[rr * rsf + 3 + cc] = v.x; if (++cc == L) { cc = 0; rr++; }
        tile[rr * rsf + 3 + cc] = v.y; if (++cc == L) { cc = 0; rr++; }
        tile[rr * rsf + 3 + cc] = v.z; if (++cc == L) { cc = 0; rr++; }
        tile[rr * rsf + 3 + cc] = v.w;
        r += dr;
        c += dc;
        if (c >= L) { c -= L; r++; }
    }
    for (int i = 4 * n4 + lane; i < n; i += 32) tile[(i / L) * rsf + 3 + (i % L)] = __ldg(src + i);
}
__device__ __forceinline__ void warp_rows_store_split(float *__restrict__ g_dc, float *__restrict__ g_rest, const float *tile, int M, int rsf,
                                                      int row0, int nrows, int lane)
{
    for (int i = lane; i < 3 * nrows; i += 32) g_dc[3 * (size_t)row0 + i] = tile[(i / 3) * rsf + (i % 3)];
    const int L = 3 * (M - 1), n = nrows * L;
    if (L == 0) return;
    float *dst = g_rest + (size_t)row0 * L;
    const int n4 = (((uintptr_t)dst & 15) == 0) ? n / 4 : 0;
    int r = (4 * lane) / L, c = 4 * lane - r * L;
    const int dr = 128 / L, dc = 128 - dr * L;
    for (int e = lane; e < n4; e += 32) {
        float4 v;
        int rr = r, cc = c;
        v.x = tile[rr * rsf + 3 + cc]; if (++cc == L) { cc = 0; rr++; }
        v.y = tile[rr * rsf + 3 + cc]; if (++cc == L) { cc = 0; rr++; }
        v.z = tile[rr * rsf + 3 + cc]; if (++cc == L) { cc = 0; rr++; }
        v.w = tile[rr * rsf + 3 + cc];
        reinterpret_cast<float4 *>(dst)[e] = v;
        r += dr;
        c += dc;
        if (c >= L) { c -= L; r++; }
    }
    for (int i = 4 * n4 + lane; i < n; i += 32) dst[i] = tile[(i / L) * rsf + 3 + (i % L)];
}
static inline int ts2d_split_row_stride(int M) { return (3 * M) | 1; }

// VanillaTS_model.py:347-363 (_training_statistic) for one visible triangle, fused into the K9 tail.  The model builds
// visible_mask from the radii it RETURNS, i.e. after `radii // render_up_scale` (:651, :674): a triangle whose full-resolution
// radius is below the up-scale factor is rendered (and gets gradients) but is not "visible" for the statistics.
__device__ __forceinline__ void model_statistics(const ModelOut &mo, int idx, float gcx, float gcy, int radius)
{
    if (radius / mo.radii_div <= 0) return;
    if (mo.grad_accum) mo.grad_accum[idx] += sqrtf(gcx * gcx + gcy * gcy);
    if (mo.grad_denom) mo.grad_denom[idx] += 1.0f;
    if (mo.csum) mo.csum[idx] = fmaxf(mo.csum[idx], mo.fwd_csum[idx]);
    if (mo.cmax) mo.cmax[idx] = fmaxf(mo.cmax[idx], mo.fwd_cmax[idx]);
    if (mo.cdenom) mo.cdenom[idx] += 1.0f;
    if (mo.max_radii) mo.max_radii[idx] = fmaxf(mo.max_radii[idx], (float)(radius / mo.radii_div));
}

static inline bool ts2d_rows_tileable(int M, const void *a, const void *b)
{
    return M > 0 && (3 * M) % 4 == 0 && ((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0;
}

// Real SH basis (degree <= 3) times per-triangle coefficients, +0.5, clamp at 0 with mask.
__device__ __forceinline__ f3 sh_colour(int deg, const float *sh, f3 pos, f3 cam, uint8_t &mask)
{
    f3 dir = pos - cam;
    dir = dir / len3(dir);
    f3 rgb = kC0 * ld3(sh);
    if (deg > 0) {
        const float x = dir.x, y = dir.y, z = dir.z;
        rgb = rgb - kC1 * y * ld3(sh + 3) + kC1 * z * ld3(sh + 6) - kC1 * x * ld3(sh + 9);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            rgb = rgb + kC2[0] * xy * ld3(sh + 12) + kC2[1] * yz * ld3(sh + 15) + kC2[2] * (2.0f * zz - xx - yy) * ld3(sh + 18) +
                  kC2[3] * xz * ld3(sh + 21) + kC2[4] * (xx - yy) * ld3(sh + 24);
            if (deg > 2) {
                rgb = rgb + kC3[0] * y * (3.0f * xx - yy) * ld3(sh + 27) + kC3[1] * xy * z * ld3(sh + 30) +
                      kC3[2] * y * (4.0f * zz - xx - yy) * ld3(sh + 33) + kC3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * ld3(sh + 36) +
                      kC3[4] * x * (4.0f * zz - xx - yy) * ld3(sh + 39) + kC3[5] * z * (xx - yy) * ld3(sh + 42) +
                      kC3[6] * x * (xx - 3.0f * yy) * ld3(sh + 45);
            }
        }
    }
    rgb = rgb + 0.5f;
    mask = (uint8_t)((rgb.x < 0 ? 1 : 0) | (rgb.y < 0 ? 2 : 0) | (rgb.z < 0 ? 4 : 0));
    return mk3(fmaxf(rgb.x, 0.0f), fmaxf(rgb.y, 0.0f), fmaxf(rgb.z, 0.0f));
}

// Back-propagation through u = v / |v|: the Jacobian is the projector onto the plane normal to u, scaled by 1 / |v|.
__device__ __forceinline__ f3 grad_norm3(f3 v, f3 dv)
{
    const float inv_n = 1.0f / len3(v);
    const f3 u = v * inv_n;
    return (dv - dot3(u, dv) * u) * inv_n;
}

__device__ __forceinline__ void st3(float *p, f3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// SH backward (the adjoint of sh_colour; backward.cu:9-119 computes the same two results): dL/dsh_k = b_k(dir) g for the active
// coefficients (zero for the inactive ones), and dL/dcentre through the view direction,  d rgb / d dir = sum_k sh_k (grad b_k)^T,
// so  dL/ddir = sum_k <g, sh_k> grad b_k.  One pass over the coefficients: the scalar c_k = <g, sh_k> is taken before out_k is
// written, so `out` may alias `sh` (the tiled kernel back-propagates in place in shared memory).  b_k are the homogeneous
// polynomials of sh_colour, their gradients written out by hand:
//   k  b_k / constant                grad b_k / constant
//   1  y                             (0, 1, 0)                         9   y (3xx - yy)        (6xy, 3(xx - yy), 0)
//   2  z                             (0, 0, 1)                         10  xyz                 (yz, xz, xy)
//   3  x                             (1, 0, 0)                         11  y (4zz - xx - yy)   (-2xy, 4zz - xx - 3yy, 8yz)
//   4  xy                            (y, x, 0)                         12  z (2zz - 3xx - 3yy) (-6xz, -6yz, 3(2zz - xx - yy))
//   5  yz                            (0, z, y)                         13  x (4zz - xx - yy)   (4zz - 3xx - yy, -2xy, 8xz)
//   6  2zz - xx - yy                 (-2x, -2y, 4z)                    14  z (xx - yy)         (2xz, -2yz, xx - yy)
//   7  xz                            (z, 0, x)                         15  x (xx - 3yy)        (3(xx - yy), -6xy, 0)
//   8  xx - yy                       (2x, -2y, 0)
__device__ __forceinline__ f3 sh_colour_bwd(int deg, int M, const float *sh, f3 pos, f3 cam, uint8_t mask, f3 g, float *out)
{
    const f3 view = pos - cam;
    const f3 u = view / len3(view);
    if (mask & 1) g.x = 0.0f;  // clamped channels pass no gradient
    if (mask & 2) g.y = 0.0f;
    if (mask & 4) g.z = 0.0f;
    f3 gd = mk3(0, 0, 0);
    auto term = [&](int k, float konst, float b, float bx, float by, float bz) {
        const float c = konst * dot3(g, ld3(sh + 3 * k));
        st3(out + 3 * k, (konst * b) * g);
        gd.x = fmaf(c, bx, gd.x);
        gd.y = fmaf(c, by, gd.y);
        gd.z = fmaf(c, bz, gd.z);
    };
    const float x = u.x, y = u.y, z = u.z;
    st3(out, kC0 * g);
    int active = 1;
    if (deg > 0) {
        term(1, -kC1, y, 0.0f, 1.0f, 0.0f);
        term(2, kC1, z, 0.0f, 0.0f, 1.0f);
        term(3, -kC1, x, 1.0f, 0.0f, 0.0f);
        active = 4;
    }
    if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z;
        const float d = xx - yy, p = 2.0f * zz - xx - yy;
        term(4, kC2[0], x * y, y, x, 0.0f);
        term(5, kC2[1], y * z, 0.0f, z, y);
        term(6, kC2[2], p, -2.0f * x, -2.0f * y, 4.0f * z);
        term(7, kC2[3], x * z, z, 0.0f, x);
        term(8, kC2[4], d, 2.0f * x, -2.0f * y, 0.0f);
        active = 9;
        if (deg > 2) {
            const float xy = x * y, yz = y * z, xz = x * z;
            const float r = 4.0f * zz - xx - yy;  // shared by k = 11 and k = 13
            term(9, kC3[0], y * (3.0f * xx - yy), 6.0f * xy, 3.0f * d, 0.0f);
            term(10, kC3[1], xy * z, yz, xz, xy);
            term(11, kC3[2], y * r, -2.0f * xy, r - 2.0f * yy, 8.0f * yz);
            term(12, kC3[3], z * (p - 2.0f * (xx + yy)), -6.0f * xz, -6.0f * yz, 3.0f * p);
            term(13, kC3[4], x * r, r - 2.0f * xx, -2.0f * xy, 8.0f * xz);
            term(14, kC3[5], z * d, 2.0f * xz, -2.0f * yz, d);
            term(15, kC3[6], x * (xx - 3.0f * yy), 3.0f * d, -6.0f * xy, 0.0f);
            active = 16;
        }
    }
    for (int k = 3 * active; k < 3 * M; k++) out[k] = 0.0f;
    return grad_norm3(view, gd);
}
