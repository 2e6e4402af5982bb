/*
 * ts2d_oracle.c -- CPU restatement of the reference 2D triangle-splatting rasterizer (and, in its last section,
 * of the reference's 3D-primitive package on the same pipeline).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under triangle_splatting_b200/ may include, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU baseline -- never as the product path.
 *
 * Parity status: the reference ships no golden vectors.  This oracle is pinned
 * against outputs of the reference's own CUDA extension (oracle/_ref, built by oracle/build_ref.py)
 * run on a B200 on seeded scenes; those outputs are committed as tests/golden/*.npz together with
 * the generating script tests/golden/make_golden.py, and tests/test_oracle_golden.py checks them.
 *
 * It follows, stage by stage (R2D = /root/reference/submodules/diff-triangle-rasterization-2D):
 *   sh_to_rgb            R2D/src/forward.cu:9-59      (computeRGBFromSH)
 *   ts2d_oracle_preprocess   R2D/src/forward.cu:61-193    (FORWARD::preprocessCUDA), helpers auxiliary.h:35-118
 *   ts2d_oracle_bin          R2D/src/rasterizer.cu:37-99, 186-231 (scan, duplicateWithKeys, SortPairs, identifyTileRanges)
 *   ts2d_oracle_render       R2D/src/forward.cu:198-355   (FORWARD::renderCUDA)
 *   ts2d_oracle_render_bwd   R2D/src/backward.cu:265-493  (BACKWARD::renderCUDA)
 *   ts2d_oracle_preprocess_bwd R2D/src/backward.cu:9-263  (SH backward, projection backward, BACKWARD::preprocessCUDA)
 *
 * Compiled twice: -DREAL=float (an fp32 "mirror": same operation order as the reference, but no
 * FMA contraction and glibc powf/expf, so float results agree to a few ulp, not bit-for-bit) and
 * -DREAL=double ("truth" for accuracy comparisons).  Per-triangle gradient sums are accumulated in
 * double in both builds (the reference sums them with fp32 atomics in a non-deterministic order).
 * The one fp64 hop of the reference (ndc2Pix, auxiliary.h:35-38) is kept in double in both builds.
 *
 * Third build, -DREAL=double -DDECIDE_F32 ("f64d": truth GIVEN the reference's decisions).  At the benchmarked sizes a pure
 * fp64 evaluation takes a handful of different per-pair / per-pixel decisions (alpha < 1/255, T <= 1e-4) than any fp32
 * implementation, so its outputs are not comparable entry by entry.  In this build every decision is taken from an fp32
 * shadow evaluation with the reference's arithmetic (fmaf where the reference's sm_100 SASS contracts, IEEE divide, powf,
 * expf) on the fp32 per-triangle state handed in by the caller, the early termination is FORCED to the n_contrib handed in
 * (the reference's own, bit-exact on the GPU), and every value is computed in double: the exact-arithmetic result of the very
 * computation the reference performs.  tests/ use it to hold gradient accuracy at C2/C3 scale (ours vs truth <= k * reference
 * vs truth).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
typedef REAL real;

#define TILE 16
#define EPSF ((real)(float)1e-8) /* auxiliary.h:8  EPS = (float)(1e-8) */

#define R_IS_FLOAT (sizeof(real) == sizeof(float))
static inline real r_sqrt(real x) { return R_IS_FLOAT ? (real)sqrtf((float)x) : (real)sqrt((double)x); }
static inline real r_pow(real x, real y) { return R_IS_FLOAT ? (real)powf((float)x, (float)y) : (real)pow((double)x, (double)y); }
static inline real r_exp(real x) { return R_IS_FLOAT ? (real)expf((float)x) : (real)exp((double)x); }
static inline real r_abs(real x) { return x < 0 ? -x : x; }
static inline real r_ceil(real x) { return R_IS_FLOAT ? (real)ceilf((float)x) : (real)ceil((double)x); }
static inline real r_min(real a, real b) { return a < b ? a : b; } /* no NaNs on this path */
static inline real r_max(real a, real b) { return a > b ? a : b; }

/* float -> int32 with CUDA's cvt.rzi.s32 semantics (saturating, NaN -> 0). */
static inline int f2i(real v)
{
    if (!(v == v)) return 0;
    if (v >= (real)2147483647.0) return 2147483647;
    if (v <= (real)-2147483648.0) return (-2147483647 - 1);
    return (int)v;
}

static const real SH_C0 = (real)0.28209479177387814f;
static const real SH_C1 = (real)0.4886025119029199f;
static const real SH_C2[5] = {(real)1.0925484305920792f, (real)-1.0925484305920792f, (real)0.31539156525252005f,
                              (real)-1.0925484305920792f, (real)0.5462742152960396f};
static const real SH_C3[7] = {(real)-0.5900435899266435f, (real)2.890611442640554f, (real)-0.4570457994644658f,
                              (real)0.3731763325901154f, (real)-0.4570457994644658f, (real)1.445305721320277f,
                              (real)-0.5900435899266435f};

typedef struct { real x, y; } v2;
typedef struct { real x, y, z; } v3;

static inline v3 v3_make(real x, real y, real z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, real s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_div(v3 a, real s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static inline real v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline real v3_norm(v3 a) { return r_sqrt(v3_dot(a, a)); }
static inline v3 v3_cross(v3 a, v3 b) { return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline v2 v2_make(real x, real y) { v2 r = {x, y}; return r; }
static inline v2 v2_add(v2 a, v2 b) { return v2_make(a.x + b.x, a.y + b.y); }
static inline v2 v2_sub(v2 a, v2 b) { return v2_make(a.x - b.x, a.y - b.y); }
static inline v2 v2_scale(v2 a, real s) { return v2_make(a.x * s, a.y * s); }
static inline v2 v2_mul(v2 a, v2 b) { return v2_make(a.x * b.x, a.y * b.y); }
static inline real v2_cross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline v2 v2_perp(v2 a) { return v2_make(a.y, -a.x); } /* auxiliary.h:184-187 cross(float2) */
static inline real v2_norm(v2 a) { return r_sqrt(a.x * a.x + a.y * a.y); }

/* auxiliary.h:40-48 (column-major 4x4 as uploaded by the Python side: m[col*4+row]) */
static inline v3 xform_point43(v3 p, const real *m)
{
    return v3_make(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                   m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
/* auxiliary.h:69-77 */
static inline v3 xform_vec43(v3 p, const real *m)
{
    return v3_make(m[0] * p.x + m[4] * p.y + m[8] * p.z, m[1] * p.x + m[5] * p.y + m[9] * p.z, m[2] * p.x + m[6] * p.y + m[10] * p.z);
}
/* auxiliary.h:79-87 */
static inline v3 xform_vec43_T(v3 p, const real *m)
{
    return v3_make(m[0] * p.x + m[1] * p.y + m[2] * p.z, m[4] * p.x + m[5] * p.y + m[6] * p.z, m[8] * p.x + m[9] * p.y + m[10] * p.z);
}
/* auxiliary.h:89-95: homogeneous projection with 1/(|w|+eps). */
static inline v3 project_point(v3 p, const real *m)
{
    real hx = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
    real hy = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
    real hz = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
    real hw = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15];
    real winv = (real)1.0 / (r_abs(hw) + EPSF);
    return v3_make(hx * winv, hy * winv, hz * winv);
}
/* auxiliary.h:97-118: first-order projection of a view-space vector at p_view. */
static inline v2 project_vec_approx(v3 p, v3 d, real tfx, real tfy)
{
    return v2_make((d.x - d.z * p.x / p.z) / (p.z * tfx), (d.y - d.z * p.y / p.z) / (p.z * tfy));
}
/* auxiliary.h:35-38: the reference evaluates this in double and rounds to float. */
static inline real ndc2pix(real v, int S) { return (real)((((double)v + 1.0) * (double)S - 1.0) * 0.5); }

/* auxiliary.h:128-138 */
static inline v2 dnorm2(v2 v, v2 dv)
{
    real sum2 = v.x * v.x + v.y * v.y;
    real n = r_sqrt(sum2);
    real inv = (real)1.0 / (n * n * n);
    return v2_make(((sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y) * inv, (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y) * inv);
}
/* auxiliary.h:140-151 */
static inline v3 dnorm3(v3 v, v3 dv)
{
    real sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    real n = r_sqrt(sum2);
    real inv = (real)1.0 / (n * n * n);
    v3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * inv;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * inv;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * inv;
    return r;
}

/* ---------------------------------------------------------------- SH -> RGB (forward.cu:9-59) */
static v3 sh_to_rgb(int deg, int M, v3 pos, v3 campos, const real *sh_all, int idx, uint8_t *clamped)
{
    v3 dir = v3_sub(pos, campos);
    dir = v3_div(dir, v3_norm(dir));
    const real *s = sh_all + (size_t)idx * M * 3;
#define SH(i) v3_make(s[3 * (i)], s[3 * (i) + 1], s[3 * (i) + 2])
    v3 rgb = v3_scale(SH(0), SH_C0);
    if (deg > 0) {
        real x = dir.x, y = dir.y, z = dir.z;
        rgb = v3_sub(rgb, v3_scale(SH(1), SH_C1 * y));
        rgb = v3_add(rgb, v3_scale(SH(2), SH_C1 * z));
        rgb = v3_sub(rgb, v3_scale(SH(3), SH_C1 * x));
        if (deg > 1) {
            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            rgb = v3_add(rgb, v3_scale(SH(4), SH_C2[0] * xy));
            rgb = v3_add(rgb, v3_scale(SH(5), SH_C2[1] * yz));
            rgb = v3_add(rgb, v3_scale(SH(6), SH_C2[2] * ((real)2.0 * zz - xx - yy)));
            rgb = v3_add(rgb, v3_scale(SH(7), SH_C2[3] * xz));
            rgb = v3_add(rgb, v3_scale(SH(8), SH_C2[4] * (xx - yy)));
            if (deg > 2) {
                rgb = v3_add(rgb, v3_scale(SH(9), SH_C3[0] * y * ((real)3.0 * xx - yy)));
                rgb = v3_add(rgb, v3_scale(SH(10), SH_C3[1] * xy * z));
                rgb = v3_add(rgb, v3_scale(SH(11), SH_C3[2] * y * ((real)4.0 * zz - xx - yy)));
                rgb = v3_add(rgb, v3_scale(SH(12), SH_C3[3] * z * ((real)2.0 * zz - (real)3.0 * xx - (real)3.0 * yy)));
                rgb = v3_add(rgb, v3_scale(SH(13), SH_C3[4] * x * ((real)4.0 * zz - xx - yy)));
                rgb = v3_add(rgb, v3_scale(SH(14), SH_C3[5] * z * (xx - yy)));
                rgb = v3_add(rgb, v3_scale(SH(15), SH_C3[6] * x * (xx - (real)3.0 * yy)));
            }
        }
    }
#undef SH
    rgb.x += (real)0.5; rgb.y += (real)0.5; rgb.z += (real)0.5;
    clamped[3 * idx + 0] = rgb.x < 0;
    clamped[3 * idx + 1] = rgb.y < 0;
    clamped[3 * idx + 2] = rgb.z < 0;
    return v3_make(r_max(rgb.x, 0), r_max(rgb.y, 0), r_max(rgb.z, 0));
}

/* ------------------------------------------------------ preprocess (forward.cu:61-193)
 * All per-triangle outputs are SoA arrays of length P (zero-initialised here, like the reference's
 * fill_(0) of its state buffer, param_struct.h:26-34). */
int ts2d_oracle_preprocess(int W, int H, int P, int D, int M, int rich_info, int use_shs, int back_culling, real tfx, real tfy,
                           const real *viewmatrix, const real *projmatrix, const real *campos, const real *vertex, const real *shs,
                           int32_t *radii, real *v2d /*[P][3][2]*/, real *area2_out, real *normal_view /*[P][3]*/, real *v_depth /*[P][3]*/,
                           real *depth, real *rgb /*[P][3]*/, uint8_t *clamped /*[P][3]*/, uint32_t *tiles_touched,
                           uint32_t *rect_min /*[P][2]*/, uint32_t *rect_max /*[P][2]*/)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const v3 cam = v3_make(campos[0], campos[1], campos[2]);
    memset(radii, 0, sizeof(int32_t) * P);
    memset(v2d, 0, sizeof(real) * 6 * P);
    memset(area2_out, 0, sizeof(real) * P);
    memset(normal_view, 0, sizeof(real) * 3 * P);
    memset(v_depth, 0, sizeof(real) * 3 * P);
    memset(depth, 0, sizeof(real) * P);
    memset(rgb, 0, sizeof(real) * 3 * P);
    memset(clamped, 0, 3 * (size_t)P);
    memset(tiles_touched, 0, sizeof(uint32_t) * P);
    memset(rect_min, 0, sizeof(uint32_t) * 2 * P);
    memset(rect_max, 0, sizeof(uint32_t) * 2 * P);

    for (int i = 0; i < P; i++) {
        const real *vp = vertex + 9 * (size_t)i;
        const v3 a = v3_make(vp[0], vp[1], vp[2]), b = v3_make(vp[3], vp[4], vp[5]), c = v3_make(vp[6], vp[7], vp[8]);
        const v3 center = v3_div(v3_add(v3_add(a, b), c), (real)3.0);
        const v3 cproj = project_point(center, projmatrix);
        if (cproj.z <= 0) continue; /* near culling, forward.cu:98 */

        const v3 cview = xform_point43(center, viewmatrix);
        const real limx = (real)1.3f * tfx * cview.z, limy = (real)1.3f * tfy * cview.z;
        const v3 cclip = v3_make(r_min(r_max(-limx, cview.x), limx), r_min(r_max(-limy, cview.y), limy), cview.z);

        const v3 r1 = v3_sub(a, center), r2 = v3_sub(b, center), r3 = v3_sub(c, center);
        const v3 r1v = xform_vec43(r1, viewmatrix), r2v = xform_vec43(r2, viewmatrix);
        if (v3_norm(v3_cross(r1v, r2v)) < EPSF) continue; /* degenerate, :113 */
        const v3 r3v = xform_vec43(r3, viewmatrix);
        const v2 p1 = project_vec_approx(cclip, r1v, tfx, tfy), p2 = project_vec_approx(cclip, r2v, tfx, tfy), p3 = project_vec_approx(cclip, r3v, tfx, tfy);
        const real n1 = v2_norm(p1), n2 = v2_norm(p2), n3 = v2_norm(p3);
        if (n1 < EPSF || n2 < EPSF || n3 < EPSF) continue; /* :124 */

        const v2 scaling = v2_make((real)0.5f * (real)W, (real)0.5f * (real)H);
        const real ks = (real)0.5f; /* low-pass growth, :128 */
        const v2 q1 = v2_mul(p1, v2_make(scaling.x + ks / n1, scaling.y + ks / n1));
        const v2 q2 = v2_mul(p2, v2_make(scaling.x + ks / n2, scaling.y + ks / n2));
        const v2 q3 = v2_mul(p3, v2_make(scaling.x + ks / n3, scaling.y + ks / n3));
        const v2 c2d = v2_make(ndc2pix(cproj.x, W), ndc2pix(cproj.y, H));
        const v2 s1 = v2_add(c2d, q1), s2 = v2_add(c2d, q2), s3 = v2_add(c2d, q3);
        const real area2 = v2_cross(v2_sub(s2, s1), v2_sub(s3, s1));
        if (back_culling) {
            if (area2 >= -EPSF) continue; /* :142 */
        } else {
            if (r_abs(area2) < EPSF) continue; /* :147 */
        }
        const real dil = (real)3.0f;
        const v2 d1 = v2_add(c2d, v2_scale(q1, dil)), d2 = v2_add(c2d, v2_scale(q2, dil)), d3 = v2_add(c2d, v2_scale(q3, dil));
        const v2 vmin = v2_make(r_min(r_min(d1.x, d2.x), d3.x), r_min(r_min(d1.y, d2.y), d3.y));
        const v2 vmax = v2_make(r_max(r_max(d1.x, d2.x), d3.x), r_max(r_max(d1.y, d2.y), d3.y));

        /* :158-161  min(grid, max(0, (int)(v/16))) -- grid is unsigned, the int is >= 0 after max. */
        int ix0 = f2i(vmin.x / (real)TILE), iy0 = f2i(vmin.y / (real)TILE);
        int ix1 = f2i((vmax.x + (real)(TILE - 1)) / (real)TILE), iy1 = f2i((vmax.y + (real)(TILE - 1)) / (real)TILE);
        uint32_t rx0 = (uint32_t)(ix0 < 0 ? 0 : ix0), ry0 = (uint32_t)(iy0 < 0 ? 0 : iy0);
        uint32_t rx1 = (uint32_t)(ix1 < 0 ? 0 : ix1), ry1 = (uint32_t)(iy1 < 0 ? 0 : iy1);
        if (rx0 > (uint32_t)gx) rx0 = gx;
        if (ry0 > (uint32_t)gy) ry0 = gy;
        if (rx1 > (uint32_t)gx) rx1 = gx;
        if (ry1 > (uint32_t)gy) ry1 = gy;
        if (rx1 <= rx0 || ry1 <= ry0) continue; /* :162 */

        if (use_shs) {
            v3 col = sh_to_rgb(D, M, center, cam, shs, i, clamped);
            rgb[3 * i] = col.x; rgb[3 * i + 1] = col.y; rgb[3 * i + 2] = col.z;
        }
        if (rich_info) {
            v3 n = v3_cross(r1v, r2v);
            n = v3_div(n, v3_norm(n));
            normal_view[3 * i] = n.x; normal_view[3 * i + 1] = n.y; normal_view[3 * i + 2] = n.z;
            v_depth[3 * i] = r1v.z + cview.z; v_depth[3 * i + 1] = r2v.z + cview.z; v_depth[3 * i + 2] = r3v.z + cview.z;
        }
        v2d[6 * i + 0] = s1.x; v2d[6 * i + 1] = s1.y; v2d[6 * i + 2] = s2.x; v2d[6 * i + 3] = s2.y; v2d[6 * i + 4] = s3.x; v2d[6 * i + 5] = s3.y;
        area2_out[i] = area2;
        depth[i] = cview.z;
        tiles_touched[i] = (rx1 - rx0) * (ry1 - ry0);
        rect_min[2 * i] = rx0; rect_min[2 * i + 1] = ry0;
        rect_max[2 * i] = rx1; rect_max[2 * i + 1] = ry1;
        radii[i] = (int32_t)r_max(r_ceil((vmax.x - vmin.x) * (real)0.5f), r_ceil((vmax.y - vmin.y) * (real)0.5f)); /* :192 */
    }
    return 0;
}

/* ------------------------------------------------ binning (rasterizer.cu:37-99,186-231)
 * depth32[] are the fp32 bit patterns of the per-triangle depth (the reference sorts raw bits).
 * Returns R (= num_rendered).  keys/list may be NULL to only count. */
typedef struct { uint64_t key; uint32_t seq; uint32_t val; } inst_t;
static int inst_cmp(const void *a, const void *b)
{
    const inst_t *x = (const inst_t *)a, *y = (const inst_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0); /* stable: emission order */
}
int64_t ts2d_oracle_bin(int W, int H, int P, const uint32_t *tiles_touched, const uint32_t *rect_min, const uint32_t *rect_max,
                        const uint32_t *depth32, uint32_t *point_offsets /*[P] inclusive scan*/, uint64_t *keys_sorted, uint32_t *list_sorted,
                        uint32_t *ranges /*[tiles][2]*/)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    uint64_t run = 0;
    for (int i = 0; i < P; i++) { run += tiles_touched[i]; point_offsets[i] = (uint32_t)run; }
    const int64_t R = (int64_t)run;
    if (!keys_sorted || !list_sorted) return R;
    inst_t *inst = (inst_t *)malloc(sizeof(inst_t) * (size_t)(R ? R : 1));
    size_t off = 0;
    for (int i = 0; i < P; i++) {
        if (tiles_touched[i] == 0) continue;
        for (uint32_t y = rect_min[2 * i + 1]; y < rect_max[2 * i + 1]; y++)
            for (uint32_t x = rect_min[2 * i]; x < rect_max[2 * i]; x++) {
                inst[off].key = ((uint64_t)(y * (uint32_t)gx + x) << 32) | depth32[i];
                inst[off].seq = (uint32_t)off;
                inst[off].val = (uint32_t)i;
                off++;
            }
    }
    qsort(inst, (size_t)R, sizeof(inst_t), inst_cmp);
    for (int64_t k = 0; k < R; k++) { keys_sorted[k] = inst[k].key; list_sorted[k] = inst[k].val; }
    free(inst);
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    for (int64_t k = 0; k < R; k++) {
        uint32_t t = (uint32_t)(keys_sorted[k] >> 32);
        if (k == 0) ranges[2 * t] = 0;
        else {
            uint32_t pt = (uint32_t)(keys_sorted[k - 1] >> 32);
            if (pt != t) { ranges[2 * pt + 1] = (uint32_t)k; ranges[2 * t] = (uint32_t)k; }
        }
        if (k == R - 1) ranges[2 * t + 1] = (uint32_t)R;
    }
    return R;
}

/* ---------------------------------------------------------- per-pair alpha (shared by fwd/bwd)
 * forward.cu:299-314 == backward.cu:382-401.  Returns 0 if the pair is skipped. */
typedef struct { real a1, a2, a3, ecc, power, G, alpha; v2 pv1, pv2, pv3; } pair_t;
static inline int eval_pair(const real *v, real area2, real op, real gamma, real px, real py, pair_t *o)
{
    o->pv1 = v2_make(v[0] - px, v[1] - py);
    o->pv2 = v2_make(v[2] - px, v[3] - py);
    o->pv3 = v2_make(v[4] - px, v[5] - py);
    o->a1 = v2_cross(o->pv2, o->pv3) / area2;
    o->a2 = v2_cross(o->pv3, o->pv1) / area2;
    o->a3 = (real)1.0f - o->a1 - o->a2;
    o->ecc = (real)1.0f - (real)3.0f * r_min(r_min(o->a1, o->a2), o->a3);
    if (o->ecc < 0 || o->ecc > (real)10.0f) return 0;
    o->power = (real)-0.5f * r_pow(o->ecc, (real)2.0f * gamma);
    o->G = r_exp(o->power);
    o->alpha = r_min((real)0.99f, op * o->G);
    if (o->alpha < (real)1.0f / (real)255.0f) return 0;
    return 1;
}

#ifdef DECIDE_F32
/* fp32 shadow of forward.cu:299-314 with the contraction of the reference's sm_100 build (cross products as
 * fma(a, b, -(c * d)), ecc as fma(min, -3, 1)).  *unclamped: op * G < 0.99 (backward.cu:442). */
static inline int decide_pair(const real *v, real area2, real op, real gamma, real px, real py, int *unclamped)
{
    const float fx = (float)px, fy = (float)py, A = (float)area2;
    const float p1x = (float)v[0] - fx, p1y = (float)v[1] - fy, p2x = (float)v[2] - fx, p2y = (float)v[3] - fy, p3x = (float)v[4] - fx,
                p3y = (float)v[5] - fy;
    const float a1 = fmaf(p2x, p3y, -(p2y * p3x)) / A, a2 = fmaf(p3x, p1y, -(p3y * p1x)) / A, a3 = (1.0f - a1) - a2;
    const float ecc = fmaf(fminf(fminf(a1, a2), a3), -3.0f, 1.0f);
    if (ecc < 0.0f || ecc > 10.0f) return 0;
    const float G = expf(-0.5f * powf(ecc, 2.0f * (float)gamma));
    const float og = (float)op * G;
    *unclamped = og < 0.99f;
    return !(fminf(0.99f, og) < 1.0f / 255.0f);
}
#endif

/* values of one pair that the shadow decided to keep (no tests; ecc is clamped at 0 where the double value dips below it) */
static inline void eval_pair_values(const real *v, real area2, real op, real gamma, real px, real py, pair_t *o)
{
    o->pv1 = v2_make(v[0] - px, v[1] - py);
    o->pv2 = v2_make(v[2] - px, v[3] - py);
    o->pv3 = v2_make(v[4] - px, v[5] - py);
    o->a1 = v2_cross(o->pv2, o->pv3) / area2;
    o->a2 = v2_cross(o->pv3, o->pv1) / area2;
    o->a3 = (real)1.0f - o->a1 - o->a2;
    o->ecc = r_max((real)0, (real)1.0f - (real)3.0f * r_min(r_min(o->a1, o->a2), o->a3));
    o->power = (real)-0.5f * r_pow(o->ecc, (real)2.0f * gamma);
    o->G = r_exp(o->power);
    o->alpha = r_min((real)0.99f, op * o->G);
}

/* ------------------------------------------------------- forward composite (forward.cu:198-355)
 * feature: [P][C] (the SH colours when use_shs).  out_feature planar [C][H][W].  contrib_* may be
 * NULL unless rich_info.  contrib_sum is accumulated in double internally. */
int ts2d_oracle_render(int W, int H, int C, real gamma, int rich_info, const uint32_t *ranges, const uint32_t *list, const real *v2d,
                       const real *area2, const real *normal_view, const real *v_depth, const real *feature, const real *opacity,
                       real background_depth, const real *background, real *final_T, uint32_t *n_contrib, real *out_feature,
                       real *out_depth, real *out_normal, real *contrib_sum, real *contrib_max, int P, int tile_step, int tile_offset)
{
    /* tile_step/tile_offset: composite only tiles with tile_id % tile_step == tile_offset (1/0 = all).  Used for the
     * bounded cpu_baseline sample and to emulate image-space tile sharding in the CPU (gloo) tests. */
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    double *csum = NULL;
    uint32_t *cmax_bits = NULL;
    if (rich_info) {
        csum = (double *)calloc((size_t)(P ? P : 1), sizeof(double));
        cmax_bits = (uint32_t *)calloc((size_t)(P ? P : 1), sizeof(uint32_t));
        for (int i = 0; i < P; i++) contrib_max[i] = 0;
    }
    if (tile_step < 1) tile_step = 1;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int ty = tile / gx, tx = tile % gx;
        if (tile % tile_step != tile_offset) continue;
        {
            const uint32_t beg = ranges[2 * (ty * gx + tx)], end = ranges[2 * (ty * gx + tx) + 1];
            for (int ly = 0; ly < TILE; ly++)
                for (int lx = 0; lx < TILE; lx++) {
                    const int px = tx * TILE + lx, py = ty * TILE + ly;
                    if (px >= W || py >= H) continue;
                    const size_t pix = (size_t)W * py + px;
                    real T = 1, acc[3] = {0, 0, 0}, accd = 0;
                    v3 accn = v3_make(0, 0, 0);
                    uint32_t contributor = 0, last = 0;
                    int done = 0;
#ifdef DECIDE_F32
                    const uint32_t forced = n_contrib[pix]; /* INPUT in this build: the reference's own stopping point */
#endif
                    for (uint32_t k = beg; k < end && !done; k++) {
#ifdef DECIDE_F32
                        if (contributor >= forced) break;
#endif
                        contributor++;
                        last = contributor;
                        const uint32_t id = list[k];
                        pair_t pr;
#ifdef DECIDE_F32
                        int unclamped_;
                        if (!decide_pair(v2d + 6 * (size_t)id, area2[id], opacity[id], gamma, (real)px, (real)py, &unclamped_)) continue;
                        eval_pair_values(v2d + 6 * (size_t)id, area2[id], opacity[id], gamma, (real)px, (real)py, &pr);
#else
                        if (!eval_pair(v2d + 6 * (size_t)id, area2[id], opacity[id], gamma, (real)px, (real)py, &pr)) continue;
#endif
                        const real contrib = pr.alpha * T;
                        for (int ch = 0; ch < C; ch++) acc[ch] += feature[(size_t)id * C + ch] * contrib;
                        if (rich_info) {
#pragma omp atomic
                            csum[id] += (double)contrib;
                            { /* lock-free max: contrib > 0, so the bit patterns of fp32 values order like the values */
                                float cf = (float)contrib;
                                uint32_t nb, ob;
                                memcpy(&nb, &cf, 4);
                                uint32_t *slot = cmax_bits + id;
                                ob = __atomic_load_n(slot, __ATOMIC_RELAXED);
                                while (nb > ob && !__atomic_compare_exchange_n(slot, &ob, nb, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                            }
                            accn.x += normal_view[3 * id] * contrib; accn.y += normal_view[3 * id + 1] * contrib; accn.z += normal_view[3 * id + 2] * contrib;
                            const real d = v_depth[3 * id] * pr.a1 + v_depth[3 * id + 1] * pr.a2 + v_depth[3 * id + 2] * pr.a3;
                            accd += d * contrib;
                        }
                        T *= ((real)1.0f - pr.alpha);
#ifndef DECIDE_F32
                        if (T <= (real)0.0001f) done = 1; /* the terminating triangle IS blended, forward.cu:332-334 */
#endif
                    }
                    final_T[pix] = T;
#ifndef DECIDE_F32
                    n_contrib[pix] = last;
#else
                    (void)last;
#endif
                    for (int ch = 0; ch < C; ch++) out_feature[(size_t)ch * H * W + pix] = acc[ch] + T * background[ch];
                    if (rich_info) {
                        out_depth[pix] = accd + T * background_depth;
                        out_normal[pix] = accn.x; out_normal[(size_t)H * W + pix] = accn.y; out_normal[2 * (size_t)H * W + pix] = accn.z;
                    }
                }
        }
    }
    if (rich_info) {
        for (int i = 0; i < P; i++) { contrib_sum[i] = (real)csum[i]; float mf; memcpy(&mf, cmax_bits + i, 4); contrib_max[i] = (real)mf; }
        free(cmax_bits);
        free(csum);
    }
    return 0;
}

/* ------------------------------------------------------ backward composite (backward.cu:265-493)
 * All g_* outputs are double[P][...] accumulators, zeroed here. */
int ts2d_oracle_render_bwd(int W, int H, int C, real gamma, int rich_info, const uint32_t *ranges, const uint32_t *list, const real *v2d,
                           const real *area2, const real *normal_view, const real *v_depth, const real *feature, const real *opacity,
                           real background_depth, const real *background, const real *final_T, const uint32_t *n_contrib,
                           const real *dL_dout_feature, const real *dL_dout_depth, const real *dL_dout_normal, int P,
                           double *g_v2d /*[P][3][2]*/, double *g_normal /*[P][3]*/, double *g_vdepth /*[P][3]*/, double *g_feature /*[P][C]*/,
                           double *g_opacity /*[P]*/, int tile_step, int tile_offset)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    if (tile_step < 1) tile_step = 1;
    memset(g_v2d, 0, sizeof(double) * 6 * P);
    memset(g_normal, 0, sizeof(double) * 3 * P);
    memset(g_vdepth, 0, sizeof(double) * 3 * P);
    memset(g_feature, 0, sizeof(double) * (size_t)C * P);
    memset(g_opacity, 0, sizeof(double) * P);
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int ty = tile / gx, tx = tile % gx;
        if (tile % tile_step != tile_offset) continue;
        {
            const uint32_t beg = ranges[2 * (ty * gx + tx)], end = ranges[2 * (ty * gx + tx) + 1];
            for (int ly = 0; ly < TILE; ly++)
                for (int lx = 0; lx < TILE; lx++) {
                    const int px = tx * TILE + lx, py = ty * TILE + ly;
                    if (px >= W || py >= H) continue;
                    const size_t pix = (size_t)W * py + px;
                    real T = final_T[pix];
                    const uint32_t last = n_contrib[pix];
                    uint32_t contributor = end - beg;
                    real acc[3] = {0, 0, 0}, gpix[3] = {0, 0, 0};
                    for (int ch = 0; ch < C; ch++) { acc[ch] = background[ch]; gpix[ch] = dL_dout_feature[(size_t)ch * H * W + pix]; }
                    v3 accn = v3_make(0, 0, 0), gn = v3_make(0, 0, 0);
                    real accd = background_depth, gd = 0;
                    if (rich_info) {
                        gn = v3_make(dL_dout_normal[pix], dL_dout_normal[(size_t)W * H + pix], dL_dout_normal[2 * (size_t)W * H + pix]);
                        gd = dL_dout_depth[pix];
                    }
                    for (uint32_t k = end; k-- > beg;) {
                        contributor--;
                        if (contributor >= last) continue; /* behind the last visited entry, :377-379 */
                        const uint32_t id = list[k];
                        const real *v = v2d + 6 * (size_t)id;
                        pair_t pr;
#ifdef DECIDE_F32
                        int unclamped_;
                        if (!decide_pair(v, area2[id], opacity[id], gamma, (real)px, (real)py, &unclamped_)) continue;
                        eval_pair_values(v, area2[id], opacity[id], gamma, (real)px, (real)py, &pr);
#else
                        if (!eval_pair(v, area2[id], opacity[id], gamma, (real)px, (real)py, &pr)) continue;
#endif
                        const real op = opacity[id];
                        T /= ((real)1.0f - pr.alpha);
                        const real contrib = pr.alpha * T;
                        real dL_dcontrib = 0;
                        v3 dL_da = v3_make(0, 0, 0);
                        for (int ch = 0; ch < C; ch++) {
#pragma omp atomic
                            g_feature[(size_t)id * C + ch] += (double)(gpix[ch] * contrib);
                            const real feat = feature[(size_t)id * C + ch];
                            dL_dcontrib += gpix[ch] * (feat - acc[ch]);
                            acc[ch] = pr.alpha * feat + ((real)1.0f - pr.alpha) * acc[ch];
                        }
                        if (rich_info) {
#pragma omp atomic
                            g_normal[3 * id] += (double)(gn.x * contrib);
#pragma omp atomic
                            g_normal[3 * id + 1] += (double)(gn.y * contrib);
#pragma omp atomic
                            g_normal[3 * id + 2] += (double)(gn.z * contrib);
                            const v3 nrm = v3_make(normal_view[3 * id], normal_view[3 * id + 1], normal_view[3 * id + 2]);
                            dL_dcontrib += v3_dot(gn, v3_sub(nrm, accn));
                            accn = v3_add(v3_scale(nrm, pr.alpha), v3_scale(accn, (real)1.0f - pr.alpha));
                            const real dL_ddepth = gd * contrib;
#pragma omp atomic
                            g_vdepth[3 * id] += (double)(dL_ddepth * pr.a1);
#pragma omp atomic
                            g_vdepth[3 * id + 1] += (double)(dL_ddepth * pr.a2);
#pragma omp atomic
                            g_vdepth[3 * id + 2] += (double)(dL_ddepth * pr.a3);
                            const v3 vd = v3_make(v_depth[3 * id], v_depth[3 * id + 1], v_depth[3 * id + 2]);
                            dL_da = v3_add(dL_da, v3_scale(vd, dL_ddepth));
                            const real dep = vd.x * pr.a1 + vd.y * pr.a2 + vd.z * pr.a3;
                            dL_dcontrib += gd * (dep - accd);
                            accd = pr.alpha * dep + ((real)1.0f - pr.alpha) * accd;
                        }
                        const real dL_dalpha = dL_dcontrib * T;
                        real dL_dpower = 0;
#ifdef DECIDE_F32
                        if (unclamped_) dL_dpower = dL_dalpha * pr.alpha;
#else
                        if (op * pr.G < (real)0.99f) dL_dpower = dL_dalpha * pr.alpha; /* :442-446 */
#endif
                        const real dL_decc = dL_dpower * 2 * gamma * pr.power / (pr.ecc + EPSF);
                        /* sub-gradient of min: first arg-min in order a1,a2,a3 with <=, :449-461 */
                        if (pr.a1 <= pr.a2 && pr.a1 <= pr.a3) dL_da.x += dL_decc * (real)-3.0f;
                        else if (pr.a2 <= pr.a1 && pr.a2 <= pr.a3) dL_da.y += dL_decc * (real)-3.0f;
                        else dL_da.z += dL_decc * (real)-3.0f;

                        const v2 s1 = v2_make(v[0], v[1]), s2 = v2_make(v[2], v[3]), s3 = v2_make(v[4], v[5]);
                        const v2 e12 = v2_sub(s2, s1), e23 = v2_sub(s3, s2), e31 = v2_sub(s1, s3);
                        const real inv = (real)1.0f / area2[id];
                        const v2 da1_d1 = v2_scale(v2_perp(v2_scale(e23, pr.a1)), inv);
                        const v2 da1_d2 = v2_scale(v2_perp(v2_add(v2_scale(e31, pr.a1), pr.pv3)), inv);
                        const v2 da1_d3 = v2_scale(v2_perp(v2_sub(v2_scale(e12, pr.a1), pr.pv2)), inv);
                        const v2 da2_d1 = v2_scale(v2_perp(v2_sub(v2_scale(e23, pr.a2), pr.pv3)), inv);
                        const v2 da2_d2 = v2_scale(v2_perp(v2_scale(e31, pr.a2)), inv);
                        const v2 da2_d3 = v2_scale(v2_perp(v2_add(v2_scale(e12, pr.a2), pr.pv1)), inv);
                        const v2 da3_d1 = v2_scale(v2_perp(v2_add(v2_scale(e23, pr.a3), pr.pv2)), inv);
                        const v2 da3_d2 = v2_scale(v2_perp(v2_sub(v2_scale(e31, pr.a3), pr.pv1)), inv);
                        const v2 da3_d3 = v2_scale(v2_perp(v2_scale(e12, pr.a3)), inv);
                        const v2 g1 = v2_add(v2_add(v2_scale(da1_d1, dL_da.x), v2_scale(da2_d1, dL_da.y)), v2_scale(da3_d1, dL_da.z));
                        const v2 g2 = v2_add(v2_add(v2_scale(da1_d2, dL_da.x), v2_scale(da2_d2, dL_da.y)), v2_scale(da3_d2, dL_da.z));
                        const v2 g3 = v2_add(v2_add(v2_scale(da1_d3, dL_da.x), v2_scale(da2_d3, dL_da.y)), v2_scale(da3_d3, dL_da.z));
                        double *gv = g_v2d + 6 * (size_t)id;
#pragma omp atomic
                        gv[0] += (double)g1.x;
#pragma omp atomic
                        gv[1] += (double)g1.y;
#pragma omp atomic
                        gv[2] += (double)g2.x;
#pragma omp atomic
                        gv[3] += (double)g2.y;
#pragma omp atomic
                        gv[4] += (double)g3.x;
#pragma omp atomic
                        gv[5] += (double)g3.y;
#pragma omp atomic
                        g_opacity[id] += (double)(dL_dalpha * pr.G); /* unconditional, :490 */
                    }
                }
        }
    }
    return 0;
}

/* computeRGBFromSHBackward (backward.cu:9-119; R3D/src/backward.cu:9-119 is byte-identical): writes dL/dsh of triangle i,
 * returns dL/d(centre) through the view direction. */
static v3 sh_backward(int D, int M, v3 center, v3 cam, const real *shs, int i, const uint8_t *clamped, const double *g_rgb, real *dL_dshs)
{
        const v3 dir_orig = v3_sub(center, cam);
        const v3 dir = v3_div(dir_orig, v3_norm(dir_orig));
        const real *s = shs + (size_t)i * M * 3;
        real *o = dL_dshs + (size_t)i * M * 3;
#define SH(k) v3_make(s[3 * (k)], s[3 * (k) + 1], s[3 * (k) + 2])
#define OUT(k, w) do { o[3 * (k)] = (w) * g.x; o[3 * (k) + 1] = (w) * g.y; o[3 * (k) + 2] = (w) * g.z; } while (0)
        v3 g = v3_make((real)g_rgb[3 * i], (real)g_rgb[3 * i + 1], (real)g_rgb[3 * i + 2]);
        g.x *= clamped[3 * i] ? 0 : 1; g.y *= clamped[3 * i + 1] ? 0 : 1; g.z *= clamped[3 * i + 2] ? 0 : 1;
        v3 dx = v3_make(0, 0, 0), dy = v3_make(0, 0, 0), dz = v3_make(0, 0, 0);
        const real x = dir.x, y = dir.y, z = dir.z;
        OUT(0, SH_C0);
        if (D > 0) {
            OUT(1, -SH_C1 * y); OUT(2, SH_C1 * z); OUT(3, -SH_C1 * x);
            dx = v3_scale(SH(3), -SH_C1); dy = v3_scale(SH(1), -SH_C1); dz = v3_scale(SH(2), SH_C1);
            if (D > 1) {
                const real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                OUT(4, SH_C2[0] * xy); OUT(5, SH_C2[1] * yz); OUT(6, SH_C2[2] * ((real)2.f * zz - xx - yy)); OUT(7, SH_C2[3] * xz); OUT(8, SH_C2[4] * (xx - yy));
                dx = v3_add(dx, v3_add(v3_add(v3_add(v3_scale(SH(4), SH_C2[0] * y), v3_scale(SH(6), SH_C2[2] * (real)2.f * -x)), v3_scale(SH(7), SH_C2[3] * z)), v3_scale(SH(8), SH_C2[4] * (real)2.f * x)));
                dy = v3_add(dy, v3_add(v3_add(v3_add(v3_scale(SH(4), SH_C2[0] * x), v3_scale(SH(5), SH_C2[1] * z)), v3_scale(SH(6), SH_C2[2] * (real)2.f * -y)), v3_scale(SH(8), SH_C2[4] * (real)2.f * -y)));
                dz = v3_add(dz, v3_add(v3_add(v3_scale(SH(5), SH_C2[1] * y), v3_scale(SH(6), SH_C2[2] * (real)2.f * (real)2.f * z)), v3_scale(SH(7), SH_C2[3] * x)));
                if (D > 2) {
                    OUT(9, SH_C3[0] * y * ((real)3.f * xx - yy)); OUT(10, SH_C3[1] * xy * z); OUT(11, SH_C3[2] * y * ((real)4.f * zz - xx - yy));
                    OUT(12, SH_C3[3] * z * ((real)2.f * zz - (real)3.f * xx - (real)3.f * yy)); OUT(13, SH_C3[4] * x * ((real)4.f * zz - xx - yy));
                    OUT(14, SH_C3[5] * z * (xx - yy)); OUT(15, SH_C3[6] * x * (xx - (real)3.f * yy));
                    v3 t = v3_scale(SH(9), SH_C3[0] * (real)3.f * (real)2.f * xy);
                    t = v3_add(t, v3_scale(SH(10), SH_C3[1] * yz));
                    t = v3_add(t, v3_scale(SH(11), SH_C3[2] * (real)-2.f * xy));
                    t = v3_add(t, v3_scale(SH(12), SH_C3[3] * (real)-3.f * (real)2.f * xz));
                    t = v3_add(t, v3_scale(SH(13), SH_C3[4] * ((real)-3.f * xx + (real)4.f * zz - yy)));
                    t = v3_add(t, v3_scale(SH(14), SH_C3[5] * (real)2.f * xz));
                    t = v3_add(t, v3_scale(SH(15), SH_C3[6] * (real)3.f * (xx - yy)));
                    dx = v3_add(dx, t);
                    t = v3_scale(SH(9), SH_C3[0] * (real)3.f * (xx - yy));
                    t = v3_add(t, v3_scale(SH(10), SH_C3[1] * xz));
                    t = v3_add(t, v3_scale(SH(11), SH_C3[2] * ((real)-3.f * yy + (real)4.f * zz - xx)));
                    t = v3_add(t, v3_scale(SH(12), SH_C3[3] * (real)-3.f * (real)2.f * yz));
                    t = v3_add(t, v3_scale(SH(13), SH_C3[4] * (real)-2.f * xy));
                    t = v3_add(t, v3_scale(SH(14), SH_C3[5] * (real)-2.f * yz));
                    t = v3_add(t, v3_scale(SH(15), SH_C3[6] * (real)-3.f * (real)2.f * xy));
                    dy = v3_add(dy, t);
                    t = v3_scale(SH(10), SH_C3[1] * xy);
                    t = v3_add(t, v3_scale(SH(11), SH_C3[2] * (real)4.f * (real)2.f * yz));
                    t = v3_add(t, v3_scale(SH(12), SH_C3[3] * (real)3.f * ((real)2.f * zz - xx - yy)));
                    t = v3_add(t, v3_scale(SH(13), SH_C3[4] * (real)4.f * (real)2.f * xz));
                    t = v3_add(t, v3_scale(SH(14), SH_C3[5] * (xx - yy)));
                    dz = v3_add(dz, t);
                }
            }
        }
#undef SH
#undef OUT
        const v3 gdir = v3_make(v3_dot(g, dx), v3_dot(g, dy), v3_dot(g, dz));
    return dnorm3(dir_orig, gdir);
}
/* ------------------------------------------- preprocess backward (backward.cu:9-263)
 * Inputs g_* are the per-triangle screen-space sums from ts2d_oracle_render_bwd (rounded to REAL
 * here, as the reference holds them in fp32).  g_rgb is dL_drgb (== dL_dfeature scratch in SH mode,
 * extension_interface.cu:232 / backward.cu:243).  Outputs are zero-initialised (torch::zeros). */
int ts2d_oracle_preprocess_bwd(int W, int H, int P, int D, int M, int use_shs, int rich_info, real tfx, real tfy, const real *viewmatrix,
                               const real *projmatrix, const real *campos, const real *vertex, const real *shs, const int32_t *radii,
                               const uint8_t *clamped, const double *g_v2d, const double *g_normal, const double *g_vdepth,
                               const double *g_rgb, real *dL_dvertex /*[P][9]*/, real *dL_dcenter2D /*[P][2]*/, real *dL_dshs /*[P][M][3]*/)
{
    memset(dL_dvertex, 0, sizeof(real) * 9 * P);
    memset(dL_dcenter2D, 0, sizeof(real) * 2 * P);
    if (M > 0) memset(dL_dshs, 0, sizeof(real) * 3 * (size_t)M * P);
    const v3 cam = v3_make(campos[0], campos[1], campos[2]);
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        const real *vp = vertex + 9 * (size_t)i;
        const v3 a = v3_make(vp[0], vp[1], vp[2]), b = v3_make(vp[3], vp[4], vp[5]), c = v3_make(vp[6], vp[7], vp[8]);
        const v3 center = v3_div(v3_add(v3_add(a, b), c), (real)3.0);
        const v3 cview = xform_point43(center, viewmatrix);
        const real limx = (real)1.3f * tfx * cview.z, limy = (real)1.3f * tfy * cview.z;
        const v3 cclip = v3_make(r_min(r_max(-limx, cview.x), limx), r_min(r_max(-limy, cview.y), limy), cview.z);
        const v3 r1 = v3_sub(a, center), r2 = v3_sub(b, center), r3 = v3_sub(c, center);
        const v3 r1v = xform_vec43(r1, viewmatrix), r2v = xform_vec43(r2, viewmatrix), r3v = xform_vec43(r3, viewmatrix);
        const v2 p1 = project_vec_approx(cclip, r1v, tfx, tfy), p2 = project_vec_approx(cclip, r2v, tfx, tfy), p3 = project_vec_approx(cclip, r3v, tfx, tfy);

        const v2 g1 = v2_make((real)g_v2d[6 * i], (real)g_v2d[6 * i + 1]), g2 = v2_make((real)g_v2d[6 * i + 2], (real)g_v2d[6 * i + 3]),
                 g3 = v2_make((real)g_v2d[6 * i + 4], (real)g_v2d[6 * i + 5]);
        const v2 gc2d = v2_add(v2_add(g1, g2), g3);
        const v2 scaling = v2_make((real)0.5f * (real)W, (real)0.5f * (real)H);
        const real ks = (real)0.5f;
        const v2 gp1 = v2_add(v2_mul(scaling, g1), v2_scale(dnorm2(p1, g1), ks));
        const v2 gp2 = v2_add(v2_mul(scaling, g2), v2_scale(dnorm2(p2, g2), ks));
        const v2 gp3 = v2_add(v2_mul(scaling, g3), v2_scale(dnorm2(p3, g3), ks));
        const v2 gcproj = v2_mul(scaling, gc2d);

        /* projectVecApproxBackward, backward.cu:131-142 */
        v3 grv[3], gcv = v3_make(0, 0, 0);
        const v3 rv[3] = {r1v, r2v, r3v};
        const v2 gp[3] = {gp1, gp2, gp3};
        for (int k = 0; k < 3; k++) {
            const real px_pz = cclip.x / cclip.z, py_pz = cclip.y / cclip.z;
            const real vx_pz = rv[k].x / cclip.z, vy_pz = rv[k].y / cclip.z, vz_pz = rv[k].z / cclip.z;
            const v2 dv = v2_make(gp[k].x / (cclip.z * tfx), gp[k].y / (cclip.z * tfy));
            grv[k] = v3_make(dv.x, dv.y, -dv.x * px_pz - dv.y * py_pz);
            const v3 dp = v3_make(-dv.x * vz_pz, -dv.y * vz_pz, dv.x * ((real)2.0f * vz_pz * px_pz - vx_pz) + dv.y * ((real)2.0f * vz_pz * py_pz - vy_pz));
            gcv = v3_add(gcv, dp);
        }
        if (cview.x < -limx || cview.x > limx) gcv.x = 0; /* clip was active, :209-216 */
        if (cview.y < -limy || cview.y > limy) gcv.y = 0;

        if (rich_info) {
            const v3 gn = v3_make((real)g_normal[3 * i], (real)g_normal[3 * i + 1], (real)g_normal[3 * i + 2]);
            const v3 gvd = v3_make((real)g_vdepth[3 * i], (real)g_vdepth[3 * i + 1], (real)g_vdepth[3 * i + 2]);
            const v3 cr = v3_cross(r1v, r2v);
            const v3 gcr = dnorm3(cr, gn);
            grv[0] = v3_add(grv[0], v3_add(v3_cross(r2v, gcr), v3_make(0, 0, gvd.x)));
            grv[1] = v3_add(grv[1], v3_add(v3_cross(gcr, r1v), v3_make(0, 0, gvd.y)));
            grv[2] = v3_add(grv[2], v3_make(0, 0, gvd.z));
            gcv = v3_add(gcv, v3_make(0, 0, gvd.x + gvd.y + gvd.z));
        }

        /* projectPointBackward with zero z-gradient, backward.cu:121-129,231 */
        v3 gcenter;
        {
            const real *m = projmatrix;
            real hx = m[0] * center.x + m[4] * center.y + m[8] * center.z + m[12];
            real hy = m[1] * center.x + m[5] * center.y + m[9] * center.z + m[13];
            real hz = m[2] * center.x + m[6] * center.y + m[10] * center.z + m[14];
            real hw = m[3] * center.x + m[7] * center.y + m[11] * center.z + m[15];
            real winv = (real)1.0f / (r_abs(hw) + EPSF);
            v3 pp = v3_make(hx * winv, hy * winv, hz * winv);
            v3 gpp = v3_make(gcproj.x, gcproj.y, 0);
            real s = r_abs(winv);
            real gh[4] = {s * gpp.x, s * gpp.y, s * gpp.z, s * (-v3_dot(gpp, pp))};
            gcenter = v3_make(m[0] * gh[0] + m[1] * gh[1] + m[2] * gh[2] + m[3] * gh[3], m[4] * gh[0] + m[5] * gh[1] + m[6] * gh[2] + m[7] * gh[3],
                              m[8] * gh[0] + m[9] * gh[1] + m[10] * gh[2] + m[11] * gh[3]);
        }
        gcenter = v3_add(gcenter, xform_vec43_T(gcv, viewmatrix));
        const v3 gr1 = xform_vec43_T(grv[0], viewmatrix), gr2 = xform_vec43_T(grv[1], viewmatrix), gr3 = xform_vec43_T(grv[2], viewmatrix);

        if (use_shs) gcenter = v3_add(gcenter, sh_backward(D, M, center, cam, shs, i, clamped, g_rgb, dL_dshs));
        const v3 gv1 = v3_div(v3_add(v3_sub(v3_sub(v3_scale(gr1, 2), gr2), gr3), gcenter), (real)3.0f);
        const v3 gv2 = v3_div(v3_add(v3_sub(v3_sub(v3_scale(gr2, 2), gr1), gr3), gcenter), (real)3.0f);
        const v3 gv3 = v3_div(v3_add(v3_sub(v3_sub(v3_scale(gr3, 2), gr1), gr2), gcenter), (real)3.0f);
        real *ov = dL_dvertex + 9 * (size_t)i;
        ov[0] = gv1.x; ov[1] = gv1.y; ov[2] = gv1.z; ov[3] = gv2.x; ov[4] = gv2.y; ov[5] = gv2.z; ov[6] = gv3.x; ov[7] = gv3.y; ov[8] = gv3.z;
        dL_dcenter2D[2 * i] = gc2d.x; dL_dcenter2D[2 * i + 1] = gc2d.y;
    }
    return 0;
}


/* =====================================================================================================================
 * 3D primitive: the reference's second rasterizer package, R3D = /root/reference/submodules/diff-triangle-rasterization-3D
 * (same binning: ts2d_oracle_bin above restates R3D/src/rasterizer.cu:37-99 too -- the files are identical there).
 *   ts3d_oracle_preprocess       R3D/src/forward.cu:61-146    (FORWARD::preprocessCUDA), helpers R3D/src/auxiliary.h:35-100
 *   ts3d_oracle_render           R3D/src/forward.cu:151-306   (FORWARD::renderCUDA)
 *   ts3d_oracle_render_bwd       R3D/src/backward.cu:215-454  (BACKWARD::renderCUDA)
 *   ts3d_oracle_preprocess_bwd   R3D/src/backward.cu:144-213  (BACKWARD::preprocessCUDA)
 * Pinned the same way as the 2D part: against outputs of the reference's own CUDA build (oracle/_ref/ts3d_ref_C*.so)
 * on seeded scenes, committed as tests/golden/r3d_*.npz.
 * ===================================================================================================================== */
/* R3D/src/auxiliary.h:35-43 (fp32, unlike the 2D package's fp64 ndc2Pix) */
static inline real proj_to_pix(real v, int S) { return (v + (real)1.0f) * (real)S * (real)0.5f - (real)0.5f; }
static inline real pix_to_proj(real v, int S) { return ((real)2.0f * v - (real)S + (real)1.0f) / (real)S; }

int ts3d_oracle_preprocess(int W, int H, int P, int D, int M, int use_shs, int back_culling, const real *viewmatrix, const real *projmatrix,
                           const real *campos, const real *vertex, const real *shs, int32_t *radii, real *v_view /*[P][3][3]*/,
                           real *normal_view /*[P][3], un-normalised*/, real *depth, real *rgb /*[P][3]*/, uint8_t *clamped /*[P][3]*/,
                           uint32_t *tiles_touched, uint32_t *rect_min /*[P][2]*/, uint32_t *rect_max /*[P][2]*/)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const v3 cam = v3_make(campos[0], campos[1], campos[2]);
    memset(radii, 0, sizeof(int32_t) * P);
    memset(v_view, 0, sizeof(real) * 9 * P);
    memset(normal_view, 0, sizeof(real) * 3 * P);
    memset(depth, 0, sizeof(real) * P);
    memset(rgb, 0, sizeof(real) * 3 * P);
    memset(clamped, 0, 3 * (size_t)P);
    memset(tiles_touched, 0, sizeof(uint32_t) * P);
    memset(rect_min, 0, sizeof(uint32_t) * 2 * P);
    memset(rect_max, 0, sizeof(uint32_t) * 2 * P);
    for (int i = 0; i < P; i++) {
        const real *vp = vertex + 9 * (size_t)i;
        const v3 a = v3_make(vp[0], vp[1], vp[2]), b = v3_make(vp[3], vp[4], vp[5]), c = v3_make(vp[6], vp[7], vp[8]);
        const v3 av = xform_point43(a, viewmatrix), bv = xform_point43(b, viewmatrix), cv = xform_point43(c, viewmatrix);
        const v3 cview = v3_div(v3_add(v3_add(av, bv), cv), (real)3.0f);
        const v3 n = v3_cross(v3_sub(bv, av), v3_sub(cv, av));
        if (v3_norm(n) < EPSF) continue;          /* degenerate, :96 */
        if (back_culling && n.z >= 0) continue;   /* :98 */
        const real dil = (real)3.0f;
        const v3 center = v3_div(v3_add(v3_add(a, b), c), (real)3.0f);
        const v3 d1 = v3_add(center, v3_scale(v3_sub(a, center), dil)), d2 = v3_add(center, v3_scale(v3_sub(b, center), dil)),
                 d3 = v3_add(center, v3_scale(v3_sub(c, center), dil));
        const v3 p1 = project_point(d1, projmatrix), p2 = project_point(d2, projmatrix), p3 = project_point(d3, projmatrix);
        if (p1.z <= 0 || p2.z <= 0 || p3.z <= 0) continue; /* near culling on the dilated vertices, :111 */
        const v2 s1 = v2_make(proj_to_pix(p1.x, W), proj_to_pix(p1.y, H)), s2 = v2_make(proj_to_pix(p2.x, W), proj_to_pix(p2.y, H)),
                 s3 = v2_make(proj_to_pix(p3.x, W), proj_to_pix(p3.y, H));
        const v2 vmin = v2_make(r_min(r_min(s1.x, s2.x), s3.x), r_min(r_min(s1.y, s2.y), s3.y));
        const v2 vmax = v2_make(r_max(r_max(s1.x, s2.x), s3.x), r_max(r_max(s1.y, s2.y), s3.y));
        int ix0 = f2i(vmin.x / (real)TILE), iy0 = f2i(vmin.y / (real)TILE);
        int ix1 = f2i((vmax.x + (real)(TILE - 1)) / (real)TILE), iy1 = f2i((vmax.y + (real)(TILE - 1)) / (real)TILE);
        uint32_t rx0 = (uint32_t)(ix0 < 0 ? 0 : ix0), ry0 = (uint32_t)(iy0 < 0 ? 0 : iy0);
        uint32_t rx1 = (uint32_t)(ix1 < 0 ? 0 : ix1), ry1 = (uint32_t)(iy1 < 0 ? 0 : iy1);
        if (rx0 > (uint32_t)gx) rx0 = gx;
        if (ry0 > (uint32_t)gy) ry0 = gy;
        if (rx1 > (uint32_t)gx) rx1 = gx;
        if (ry1 > (uint32_t)gy) ry1 = gy;
        if (rx1 <= rx0 || ry1 <= ry0) continue; /* :124 */
        if (use_shs) {
            v3 col = sh_to_rgb(D, M, center, cam, shs, i, clamped);
            rgb[3 * i] = col.x; rgb[3 * i + 1] = col.y; rgb[3 * i + 2] = col.z;
        }
        real *o = v_view + 9 * (size_t)i;
        o[0] = av.x; o[1] = av.y; o[2] = av.z; o[3] = bv.x; o[4] = bv.y; o[5] = bv.z; o[6] = cv.x; o[7] = cv.y; o[8] = cv.z;
        normal_view[3 * i] = n.x; normal_view[3 * i + 1] = n.y; normal_view[3 * i + 2] = n.z;
        depth[i] = cview.z;
        tiles_touched[i] = (rx1 - rx0) * (ry1 - ry0);
        rect_min[2 * i] = rx0; rect_min[2 * i + 1] = ry0;
        rect_max[2 * i] = rx1; rect_max[2 * i + 1] = ry1;
        radii[i] = (int32_t)r_max(r_ceil((vmax.x - vmin.x) * (real)0.5f), r_ceil((vmax.y - vmin.y) * (real)0.5f)); /* :145 */
    }
    return 0;
}

/* per-pair evaluation, R3D/src/forward.cu:243-276 (bwd == 0) / R3D/src/backward.cu:330-352 (bwd == 1).  The two differ:
 * forward divides by ray.n and skips on alpha < 1/255; backward multiplies by 1/(ray.n) and skips on G < 1/255. */
typedef struct { real depth, inv_pn, inv_nn, a1, a2, a3, ecc, power, G, alpha; v3 pv1, pv2, pv3; } pair3_t;
static inline int eval_pair3(const real *vv, const real *nv, real op, real gamma, v3 ray, int bwd, pair3_t *o)
{
    const v3 v1 = v3_make(vv[0], vv[1], vv[2]), v2 = v3_make(vv[3], vv[4], vv[5]), v3v = v3_make(vv[6], vv[7], vv[8]);
    const v3 n = v3_make(nv[0], nv[1], nv[2]);
    const real pn = v3_dot(ray, n);
    if (r_abs(pn) < EPSF) return 0;
    if (bwd) { o->inv_pn = (real)1.0f / pn; o->depth = v3_dot(v1, n) * o->inv_pn; }
    else { o->inv_pn = (real)1.0f / pn; o->depth = v3_dot(v1, n) / pn; }
    const v3 pview = v3_scale(ray, o->depth);
    o->pv1 = v3_sub(v1, pview); o->pv2 = v3_sub(v2, pview); o->pv3 = v3_sub(v3v, pview);
    o->inv_nn = (real)1.0f / v3_dot(n, n);
    o->a1 = v3_dot(v3_cross(o->pv2, o->pv3), n) * o->inv_nn;
    o->a2 = v3_dot(v3_cross(o->pv3, o->pv1), n) * o->inv_nn;
    o->a3 = (real)1.0f - o->a1 - o->a2;
    o->ecc = (real)1.0f - (real)3.0f * r_min(r_min(o->a1, o->a2), o->a3);
    if (o->ecc < 0 || o->ecc > (real)10.0f) return 0;
    o->power = (real)-0.5f * r_pow(o->ecc, (real)2.0f * gamma);
    o->G = r_exp(o->power);
    o->alpha = r_min((real)0.99f, op * o->G);
    if (bwd) return !(o->G < (real)1.0f / (real)255.0f);
    return !(o->alpha < (real)1.0f / (real)255.0f);
}

#ifdef DECIDE_F32
/* fp32 shadow of the 3D per-pair decisions (R3D/src/forward.cu:243-276, backward.cu:330-352), plain fp32 without contraction:
 * a pair within an ulp of a threshold may be decided differently from the GPU build; the tests that use this build compare
 * error quantiles, not single entries. */
static inline int decide_pair3(const real *vv, const real *nv, real op, real gamma, v3 ray, int bwd, int *unclamped)
{
    const float r[3] = {(float)ray.x, (float)ray.y, (float)ray.z}, n[3] = {(float)nv[0], (float)nv[1], (float)nv[2]};
    float v[9];
    for (int i = 0; i < 9; i++) v[i] = (float)vv[i];
    const float pn = r[0] * n[0] + r[1] * n[1] + r[2] * n[2];
    if (fabsf(pn) < (float)1e-8) return 0;
    const float v1n = v[0] * n[0] + v[1] * n[1] + v[2] * n[2];
    const float depth = bwd ? v1n * (1.0f / pn) : v1n / pn;
    float pv[9];
    for (int k = 0; k < 3; k++)
        for (int c = 0; c < 3; c++) pv[3 * k + c] = v[3 * k + c] - r[c] * depth;
    const float inv_nn = 1.0f / (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
#define CR(a, b, o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
    float c1[3], c2[3];
    CR((pv + 3), (pv + 6), c1);
    CR((pv + 6), (pv + 0), c2);
#undef CR
    const float a1 = (c1[0] * n[0] + c1[1] * n[1] + c1[2] * n[2]) * inv_nn, a2 = (c2[0] * n[0] + c2[1] * n[1] + c2[2] * n[2]) * inv_nn;
    const float a3 = 1.0f - a1 - a2;
    const float ecc = 1.0f - 3.0f * fminf(fminf(a1, a2), a3);
    if (ecc < 0.0f || ecc > 10.0f) return 0;
    const float G = expf(-0.5f * powf(ecc, 2.0f * (float)gamma));
    const float og = (float)op * G;
    *unclamped = og < 0.99f;
    if (bwd) return !(G < 1.0f / 255.0f);
    return !(fminf(0.99f, og) < 1.0f / 255.0f);
}
static inline void eval_pair3_values(const real *vv, const real *nv, real op, real gamma, v3 ray, pair3_t *o)
{
    const v3 v1 = v3_make(vv[0], vv[1], vv[2]), v2 = v3_make(vv[3], vv[4], vv[5]), v3v = v3_make(vv[6], vv[7], vv[8]);
    const v3 n = v3_make(nv[0], nv[1], nv[2]);
    const real pn = v3_dot(ray, n);
    o->inv_pn = (real)1.0f / pn;
    o->depth = v3_dot(v1, n) / pn;
    const v3 pview = v3_scale(ray, o->depth);
    o->pv1 = v3_sub(v1, pview); o->pv2 = v3_sub(v2, pview); o->pv3 = v3_sub(v3v, pview);
    o->inv_nn = (real)1.0f / v3_dot(n, n);
    o->a1 = v3_dot(v3_cross(o->pv2, o->pv3), n) * o->inv_nn;
    o->a2 = v3_dot(v3_cross(o->pv3, o->pv1), n) * o->inv_nn;
    o->a3 = (real)1.0f - o->a1 - o->a2;
    o->ecc = r_max((real)0, (real)1.0f - (real)3.0f * r_min(r_min(o->a1, o->a2), o->a3));
    o->power = (real)-0.5f * r_pow(o->ecc, (real)2.0f * gamma);
    o->G = r_exp(o->power);
    o->alpha = r_min((real)0.99f, op * o->G);
}
#endif

int ts3d_oracle_render(int W, int H, int C, real gamma, int rich_info, real tfx, real tfy, const uint32_t *ranges, const uint32_t *list,
                       const real *v_view, const real *normal_view, const real *feature, const real *opacity, real background_depth,
                       const real *background, real *final_T, uint32_t *n_contrib, real *out_feature, real *out_depth, real *out_normal,
                       real *contrib_sum, real *contrib_max, int P, int tile_step, int tile_offset)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    double *csum = NULL;
    uint32_t *cmax_bits = NULL;
    if (rich_info) {
        csum = (double *)calloc((size_t)(P ? P : 1), sizeof(double));
        cmax_bits = (uint32_t *)calloc((size_t)(P ? P : 1), sizeof(uint32_t));
    }
    if (tile_step < 1) tile_step = 1;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int ty = tile / gx, tx = tile % gx;
        if (tile % tile_step != tile_offset) continue;
        const uint32_t beg = ranges[2 * tile], end = ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                const size_t pix = (size_t)W * py + px;
                const v3 ray = v3_make(tfx * pix_to_proj((real)px, W), tfy * pix_to_proj((real)py, H), (real)1.0f);
                real T = 1, acc[3] = {0, 0, 0}, accd = 0;
                v3 accn = v3_make(0, 0, 0);
                uint32_t contributor = 0, last = 0;
                int done = 0;
#ifdef DECIDE_F32
                const uint32_t forced = n_contrib[pix];
#endif
                for (uint32_t k = beg; k < end && !done; k++) {
#ifdef DECIDE_F32
                    if (contributor >= forced) break;
#endif
                    contributor++;
                    last = contributor;
                    const uint32_t id = list[k];
                    pair3_t pr;
#ifdef DECIDE_F32
                    int unclamped_;
                    if (!decide_pair3(v_view + 9 * (size_t)id, normal_view + 3 * (size_t)id, opacity[id], gamma, ray, 0, &unclamped_)) continue;
                    eval_pair3_values(v_view + 9 * (size_t)id, normal_view + 3 * (size_t)id, opacity[id], gamma, ray, &pr);
#else
                    if (!eval_pair3(v_view + 9 * (size_t)id, normal_view + 3 * (size_t)id, opacity[id], gamma, ray, 0, &pr)) continue;
#endif
                    const real contrib = pr.alpha * T;
                    T *= ((real)1.0f - pr.alpha);
                    for (int ch = 0; ch < C; ch++) acc[ch] += feature[(size_t)id * C + ch] * contrib;
                    if (rich_info) {
#pragma omp atomic
                        csum[id] += (double)contrib;
                        {
                            float cf = (float)contrib;
                            uint32_t nb, ob;
                            memcpy(&nb, &cf, 4);
                            uint32_t *slot = cmax_bits + id;
                            ob = __atomic_load_n(slot, __ATOMIC_RELAXED);
                            while (nb > ob && !__atomic_compare_exchange_n(slot, &ob, nb, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                        }
                        accn.x += normal_view[3 * id] * contrib; accn.y += normal_view[3 * id + 1] * contrib; accn.z += normal_view[3 * id + 2] * contrib;
                        accd += pr.depth * contrib;
                    }
#ifndef DECIDE_F32
                    if (T <= (real)0.0001f) done = 1;
#endif
                }
                final_T[pix] = T;
#ifndef DECIDE_F32
                n_contrib[pix] = last;
#else
                (void)last;
#endif
                for (int ch = 0; ch < C; ch++) out_feature[(size_t)ch * H * W + pix] = acc[ch] + T * background[ch];
                if (rich_info) {
                    out_depth[pix] = accd + T * background_depth;
                    out_normal[pix] = accn.x; out_normal[(size_t)H * W + pix] = accn.y; out_normal[2 * (size_t)H * W + pix] = accn.z;
                }
            }
    }
    if (rich_info) {
        for (int i = 0; i < P; i++) { contrib_sum[i] = (real)csum[i]; float mf; memcpy(&mf, cmax_bits + i, 4); contrib_max[i] = (real)mf; }
        free(cmax_bits);
        free(csum);
    }
    return 0;
}

static inline void atomic_add3(double *dst, v3 g)
{
#pragma omp atomic
    dst[0] += (double)g.x;
#pragma omp atomic
    dst[1] += (double)g.y;
#pragma omp atomic
    dst[2] += (double)g.z;
}
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline v3 v3_comb3(real a, v3 x, real b, v3 y, real c, v3 z) { return v3_add(v3_add(v3_scale(x, a), v3_scale(y, b)), v3_scale(z, c)); }

int ts3d_oracle_render_bwd(int W, int H, int C, real gamma, int rich_info, real tfx, real tfy, const uint32_t *ranges, const uint32_t *list,
                           const real *v_view, const real *normal_view, const real *feature, const real *opacity, real background_depth,
                           const real *background, const real *final_T, const uint32_t *n_contrib, const real *dL_dout_feature,
                           const real *dL_dout_depth, const real *dL_dout_normal, int P, double *g_vview /*[P][3][3]*/,
                           double *g_normal /*[P][3]*/, double *g_feature /*[P][C]*/, double *g_opacity /*[P]*/, int tile_step, int tile_offset)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    if (tile_step < 1) tile_step = 1;
    memset(g_vview, 0, sizeof(double) * 9 * P);
    memset(g_normal, 0, sizeof(double) * 3 * P);
    memset(g_feature, 0, sizeof(double) * (size_t)C * P);
    memset(g_opacity, 0, sizeof(double) * P);
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int ty = tile / gx, tx = tile % gx;
        if (tile % tile_step != tile_offset) continue;
        const uint32_t beg = ranges[2 * tile], end = ranges[2 * tile + 1];
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx * TILE + lx, py = ty * TILE + ly;
                if (px >= W || py >= H) continue;
                const size_t pix = (size_t)W * py + px;
                const v3 ray = v3_make(tfx * pix_to_proj((real)px, W), tfy * pix_to_proj((real)py, H), (real)1.0f);
                real T = final_T[pix];
                const uint32_t last = n_contrib[pix];
                uint32_t contributor = end - beg;
                real acc[3] = {0, 0, 0}, gpix[3] = {0, 0, 0};
                for (int ch = 0; ch < C; ch++) { acc[ch] = background[ch]; gpix[ch] = dL_dout_feature[(size_t)ch * H * W + pix]; }
                v3 accn = v3_make(0, 0, 0), gn = v3_make(0, 0, 0);
                real accd = background_depth, gd = 0;
                if (rich_info) {
                    gn = v3_make(dL_dout_normal[pix], dL_dout_normal[(size_t)W * H + pix], dL_dout_normal[2 * (size_t)W * H + pix]);
                    gd = dL_dout_depth[pix];
                }
                for (uint32_t k = end; k-- > beg;) {
                    contributor--;
                    if (contributor >= last) continue;
                    const uint32_t id = list[k];
                    const real *vv = v_view + 9 * (size_t)id, *nv = normal_view + 3 * (size_t)id;
                    pair3_t pr;
                    const real op = opacity[id];
#ifdef DECIDE_F32
                    int unclamped_;
                    if (!decide_pair3(vv, nv, op, gamma, ray, 1, &unclamped_)) continue;
                    eval_pair3_values(vv, nv, op, gamma, ray, &pr);
#else
                    if (!eval_pair3(vv, nv, op, gamma, ray, 1, &pr)) continue;
#endif
                    const v3 v1 = v3_make(vv[0], vv[1], vv[2]), v2 = v3_make(vv[3], vv[4], vv[5]), v3v = v3_make(vv[6], vv[7], vv[8]);
                    const v3 n = v3_make(nv[0], nv[1], nv[2]);
                    T /= ((real)1.0f - pr.alpha);
                    const real contrib = pr.alpha * T;
                    real dL_dcontrib = 0, dL_ddepth = 0;
                    v3 dL_dnormal = v3_make(0, 0, 0);
                    for (int ch = 0; ch < C; ch++) {
#pragma omp atomic
                        g_feature[(size_t)id * C + ch] += (double)(gpix[ch] * contrib);
                        const real feat = feature[(size_t)id * C + ch];
                        dL_dcontrib += gpix[ch] * (feat - acc[ch]);
                        acc[ch] = pr.alpha * feat + ((real)1.0f - pr.alpha) * acc[ch];
                    }
                    if (rich_info) {
                        dL_dnormal = v3_add(dL_dnormal, v3_scale(gn, contrib));
                        dL_dcontrib += v3_dot(gn, v3_sub(n, accn));
                        accn = v3_add(v3_scale(n, pr.alpha), v3_scale(accn, (real)1.0f - pr.alpha));
                        dL_ddepth += gd * contrib;
                        dL_dcontrib += gd * (pr.depth - accd);
                        accd = pr.alpha * pr.depth + ((real)1.0f - pr.alpha) * accd;
                    }
                    const real dL_dalpha = dL_dcontrib * T;
#ifdef DECIDE_F32
                    const real dL_dpower = unclamped_ ? dL_dalpha * pr.alpha : 0;
#else
                    const real dL_dpower = (op * pr.G < (real)0.99f) ? dL_dalpha * pr.alpha : 0;
#endif
                    const real dL_decc = dL_dpower * 2 * gamma * pr.power / (pr.ecc + EPSF);
                    v3 dda = v3_make(0, 0, 0);
                    if (pr.a1 <= pr.a2 && pr.a1 <= pr.a3) dda.x = (real)-3.0f;
                    else if (pr.a2 <= pr.a1 && pr.a2 <= pr.a3) dda.y = (real)-3.0f;
                    else dda.z = (real)-3.0f;
                    const v3 dL_da = v3_scale(dda, dL_decc);
                    /* R3D/src/backward.cu:401-428 */
                    const v3 z3 = v3_make(0, 0, 0);
                    const v3 da1_dv1 = z3;
                    const v3 da1_dv2 = v3_scale(v3_cross(pr.pv3, n), pr.inv_nn);
                    const v3 da1_dv3 = v3_scale(v3_cross(n, pr.pv2), pr.inv_nn);
                    const v3 da1_dn = v3_scale(v3_sub(v3_cross(pr.pv2, pr.pv3), v3_scale(n, (real)2.0f * pr.a1)), pr.inv_nn);
                    const real da1_dd = v3_dot(n, v3_cross(v3_sub(v3v, v2), ray)) * pr.inv_nn;
                    const v3 da2_dv1 = v3_scale(v3_cross(n, pr.pv3), pr.inv_nn);
                    const v3 da2_dv2 = z3;
                    const v3 da2_dv3 = v3_scale(v3_cross(pr.pv1, n), pr.inv_nn);
                    const v3 da2_dn = v3_scale(v3_sub(v3_cross(pr.pv3, pr.pv1), v3_scale(n, (real)2.0f * pr.a2)), pr.inv_nn);
                    const real da2_dd = v3_dot(n, v3_cross(v3_sub(v1, v3v), ray)) * pr.inv_nn;
                    const v3 da3_dv1 = v3_sub(v3_neg(da1_dv1), da2_dv1), da3_dv2 = v3_sub(v3_neg(da1_dv2), da2_dv2),
                             da3_dv3 = v3_sub(v3_neg(da1_dv3), da2_dv3), da3_dn = v3_sub(v3_neg(da1_dn), da2_dn);
                    const real da3_dd = -da1_dd - da2_dd;
                    dL_ddepth += dL_da.x * da1_dd + dL_da.y * da2_dd + dL_da.z * da3_dd;
                    const v3 dd_dv1 = v3_scale(n, pr.inv_pn);
                    const v3 dd_dn = v3_scale(v3_sub(v1, v3_scale(ray, pr.depth)), pr.inv_pn);
                    const v3 g1 = v3_add(v3_comb3(dL_da.x, da1_dv1, dL_da.y, da2_dv1, dL_da.z, da3_dv1), v3_scale(dd_dv1, dL_ddepth));
                    const v3 g2 = v3_comb3(dL_da.x, da1_dv2, dL_da.y, da2_dv2, dL_da.z, da3_dv2);
                    const v3 g3 = v3_comb3(dL_da.x, da1_dv3, dL_da.y, da2_dv3, dL_da.z, da3_dv3);
                    dL_dnormal = v3_add(dL_dnormal, v3_add(v3_comb3(dL_da.x, da1_dn, dL_da.y, da2_dn, dL_da.z, da3_dn), v3_scale(dd_dn, dL_ddepth)));
                    atomic_add3(g_vview + 9 * (size_t)id, g1);
                    atomic_add3(g_vview + 9 * (size_t)id + 3, g2);
                    atomic_add3(g_vview + 9 * (size_t)id + 6, g3);
                    atomic_add3(g_normal + 3 * (size_t)id, dL_dnormal);
#pragma omp atomic
                    g_opacity[id] += (double)(dL_dalpha * pr.G);
                }
            }
    }
    return 0;
}

int ts3d_oracle_preprocess_bwd(int P, int D, int M, int use_shs, const real *viewmatrix, const real *campos, const real *vertex, const real *shs,
                               const int32_t *radii, const uint8_t *clamped, const real *v_view, const double *g_vview, const double *g_normal,
                               const double *g_rgb, real *dL_dvertex /*[P][9]*/, real *dL_dcenter2D /*[P][2]*/, real *dL_dshs /*[P][M][3]*/)
{
    memset(dL_dvertex, 0, sizeof(real) * 9 * P);
    memset(dL_dcenter2D, 0, sizeof(real) * 2 * P);
    if (M > 0) memset(dL_dshs, 0, sizeof(real) * 3 * (size_t)M * P);
    const v3 cam = v3_make(campos[0], campos[1], campos[2]);
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        const real *vv = v_view + 9 * (size_t)i;
        const v3 av = v3_make(vv[0], vv[1], vv[2]), bv = v3_make(vv[3], vv[4], vv[5]), cv = v3_make(vv[6], vv[7], vv[8]);
        const double *g = g_vview + 9 * (size_t)i;
        v3 g1v = v3_make((real)g[0], (real)g[1], (real)g[2]), g2v = v3_make((real)g[3], (real)g[4], (real)g[5]),
           g3v = v3_make((real)g[6], (real)g[7], (real)g[8]);
        const v3 gN = v3_make((real)g_normal[3 * i], (real)g_normal[3 * i + 1], (real)g_normal[3 * i + 2]);
        g1v = v3_add(g1v, v3_cross(v3_sub(bv, cv), gN)); /* :179-181 */
        g2v = v3_add(g2v, v3_cross(v3_sub(cv, av), gN));
        g3v = v3_add(g3v, v3_cross(v3_sub(av, bv), gN));
        v3 g1 = xform_vec43_T(g1v, viewmatrix), g2 = xform_vec43_T(g2v, viewmatrix), g3 = xform_vec43_T(g3v, viewmatrix);
        if (use_shs) {
            const real *vp = vertex + 9 * (size_t)i;
            const v3 a = v3_make(vp[0], vp[1], vp[2]), b = v3_make(vp[3], vp[4], vp[5]), c = v3_make(vp[6], vp[7], vp[8]);
            const v3 center = v3_div(v3_add(v3_add(a, b), c), (real)3.0f);
            const v3 gsh = sh_backward(D, M, center, cam, shs, i, clamped, g_rgb, dL_dshs);
            g1 = v3_add(g1, v3_div(gsh, (real)3.0f));
            g2 = v3_add(g2, v3_div(gsh, (real)3.0f));
            g3 = v3_add(g3, v3_div(gsh, (real)3.0f));
        }
        real *ov = dL_dvertex + 9 * (size_t)i;
        ov[0] = g1.x; ov[1] = g1.y; ov[2] = g1.z; ov[3] = g2.x; ov[4] = g2.y; ov[5] = g2.z; ov[6] = g3.x; ov[7] = g3.y; ov[8] = g3.z;
        const v3 gcv = xform_vec43(v3_add(v3_add(g1, g2), g3), viewmatrix); /* :211-213 */
        dL_dcenter2D[2 * i] = gcv.x; dL_dcenter2D[2 * i + 1] = gcv.y;
    }
    return 0;
}

int ts2d_oracle_sizeof_real(void) { return (int)sizeof(real); }
