// ts2d_preprocess.cu -- per-triangle stages: forward preprocess + SH colour (K1) and their backward (K9).
//
// Replaces R2D/src/forward.cu:9-193 (computeRGBFromSH, FORWARD::preprocessCUDA) and
// R2D/src/backward.cu:9-263 (computeRGBFromSHBackward, projectPointBackward,
// projectVecApproxBackward, BACKWARD::preprocessCUDA).
//
// Parity rule: every quantity that feeds a comparison or a float->int cast (cull tests, tile rect,
// radii, depth bits) keeps the reference's expression tree, is compiled WITHOUT --use_fast_math and
// with the default FMA contraction, so the integer outputs are bit-identical.
// Layout is new: one 48 B (+32 B rich) record per triangle instead of 13 SoA arrays, SH
// clamp mask packed to one byte, tile rect packed to 4 x u16.
#include "ts2d_common.cuh"
#include "ts2d_sh.cuh"

// Shared geometric prefix of K1 and K9: centre, view-space offsets, projected offsets.
struct TriGeom {
    f3 center, center_view, center_clip, r1v, r2v, r3v;
    f2 r1p, r2p, r3p;
    float limx, limy;
};

__device__ __forceinline__ void tri_geom_view(f3 a, f3 b, f3 c, const float *view, float tfx, float tfy, TriGeom &t, f3 &r1, f3 &r2, f3 &r3)
{
    t.center = (a + b + c) / 3.0f;
    t.center_view = xf_point(view, t.center);
    t.limx = 1.3f * tfx * t.center_view.z;
    t.limy = 1.3f * tfy * t.center_view.z;
    t.center_clip = mk3(fminf(fmaxf(-t.limx, t.center_view.x), t.limx), fminf(fmaxf(-t.limy, t.center_view.y), t.limy), t.center_view.z);
    r1 = a - t.center;
    r2 = b - t.center;
    r3 = c - t.center;
}

// TILED: the warp stages its 32 SH rows through shared memory with coalesced 16-byte loads (see warp_rows_load);
// only the data movement changes -- the arithmetic on the coefficients is the same expression tree either way.
// MODEL: parameter-space inputs (ts2d_model_inputs) -- raw vertices rescaled in registers, sigmoid / STE opacity, SH rows staged
// from the split f_dc / f_rest tensors, background depth maxed into the header.  Same expression trees downstream.
template <bool TILED, bool MODEL>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_preprocess(int W, int H, int P, int D, int M, int C, bool rich, bool use_shs, int gx, int gy, bool back_culling, float tfx, float tfy,
             int shard_rank, int shard_world, const float *__restrict__ view, const float *__restrict__ proj, const float *__restrict__ campos,
             const float *__restrict__ vertex, const float *__restrict__ shs, const float *__restrict__ feature,
             const float *__restrict__ opacity, int32_t *__restrict__ radii, float4 *__restrict__ rec0, float4 *__restrict__ rec1,
             uint32_t *__restrict__ dkey, uint32_t *__restrict__ ids, uint32_t *__restrict__ tiles, ushort4 *__restrict__ rect,
             uint8_t *__restrict__ clamp, ModelIn mi)
{
    extern __shared__ __align__(16) float s_rows[];
    ts2d_grid_chain();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const float *sh_row = shs + (size_t)idx * M * 3;
    if (MODEL) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int row0 = idx - lane, nrows = min(32, P - row0);
        if (nrows <= 0) return;
        const int rsf = (3 * M) | 1;
        float *tile = s_rows + (size_t)warp * 32 * rsf;
        warp_rows_load_split(mi.f_dc, mi.f_rest, tile, M, rsf, row0, nrows, lane);
        if (mi.bg_bits) bg_depth_max(mi, vertex + 9 * (size_t)min(idx, P - 1), idx < P, ld3(campos));
        __syncwarp();
        sh_row = tile + lane * rsf;
    } else if (TILED) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int row0 = idx - lane, nrows = min(32, P - row0);
        if (nrows <= 0) return;
        const int q = (3 * M) / 4, rs4 = q + 1;
        float *tile = s_rows + (size_t)warp * 32 * rs4 * 4;
        warp_rows_load(shs + (size_t)row0 * M * 3, tile, q, rs4, nrows, lane);
        __syncwarp();
        sh_row = tile + lane * rs4 * 4;
    }
    if (idx >= P) return;

    int out_radius = 0;
    uint32_t out_tiles = 0, out_key = 0xFFFFFFFFu;
    ushort4 out_rect = make_ushort4(0, 0, 0, 0);

    do {
        TriGeom t;
        f3 r1, r2, r3;
        const float *vp = vertex + 9 * (size_t)idx;
        {
            // near culling on the projected centre happens before the view transform in the reference;
            // order of evaluation does not change values.
            f3 a, b, c;
            load_tri<MODEL>(vp, mi, a, b, c);
            const f3 center = (a + b + c) / 3.0f;
            const f3 cproj = project_center(proj, center);
            if (cproj.z <= 0) break;
            tri_geom_view(a, b, c, view, tfx, tfy, t, r1, r2, r3);
            t.r1v = xf_vec(view, r1);
            t.r2v = xf_vec(view, r2);
            if (len3(cross3(t.r1v, t.r2v)) < TS2D_EPS) break;
            t.r3v = xf_vec(view, r3);
            t.r1p = project_offset(t.center_clip, t.r1v, tfx, tfy);
            t.r2p = project_offset(t.center_clip, t.r2v, tfx, tfy);
            t.r3p = project_offset(t.center_clip, t.r3v, tfx, tfy);
            const float n1 = len2(t.r1p), n2 = len2(t.r2p), n3 = len2(t.r3p);
            if (n1 < TS2D_EPS || n2 < TS2D_EPS || n3 < TS2D_EPS) break;

            const f2 scaling = mk2(0.5f * W, 0.5f * H);
            const float kernel_size = 0.5f;
            const f2 q1 = t.r1p * (scaling + kernel_size / n1);
            const f2 q2 = t.r2p * (scaling + kernel_size / n2);
            const f2 q3 = t.r3p * (scaling + kernel_size / n3);
            const f2 c2d = mk2(ndc_to_pix(cproj.x, W), ndc_to_pix(cproj.y, H));
            const f2 s1 = c2d + q1, s2 = c2d + q2, s3 = c2d + q3;
            const float area2 = cross2(s2 - s1, s3 - s1);
            if (back_culling) {
                if (area2 >= -TS2D_EPS) break;
            } else {
                if (fabsf(area2) < TS2D_EPS) break;
            }
            const float dilation = 3.0f;
            const f2 d1 = c2d + dilation * q1, d2 = c2d + dilation * q2, d3 = c2d + dilation * q3;
            const f2 vmin = mk2(fminf(fminf(d1.x, d2.x), d3.x), fminf(fminf(d1.y, d2.y), d3.y));
            const f2 vmax = mk2(fmaxf(fmaxf(d1.x, d2.x), d3.x), fmaxf(fmaxf(d1.y, d2.y), d3.y));
            const uint32_t rx0 = min((uint32_t)gx, (uint32_t)max(0, (int)(vmin.x / TS2D_TILE)));
            const uint32_t ry0 = min((uint32_t)gy, (uint32_t)max(0, (int)(vmin.y / TS2D_TILE)));
            const uint32_t rx1 = min((uint32_t)gx, (uint32_t)max(0, (int)((vmax.x + TS2D_TILE - 1) / TS2D_TILE)));
            const uint32_t ry1 = min((uint32_t)gy, (uint32_t)max(0, (int)((vmax.y + TS2D_TILE - 1) / TS2D_TILE)));
            if (rx1 <= rx0 || ry1 <= ry0) break;

            f3 rgb;
            uint8_t mask = 0;
            if (use_shs) {
                rgb = sh_colour(D, sh_row, center, ld3(campos), mask);
            } else {
                const float *fp = feature + (size_t)idx * C;
                rgb = mk3(fp[0], C > 1 ? fp[1] : 0.0f, C > 2 ? fp[2] : 0.0f);
            }
            clamp[idx] = mask;
            const float depth = t.center_view.z;
            rec0[3 * (size_t)idx + 0] = make_float4(s1.x, s1.y, s2.x, s2.y);
            float sig;
            const float op = MODEL ? model_opacity(mi, idx, sig) : opacity[idx];
            rec0[3 * (size_t)idx + 1] = make_float4(s3.x, s3.y, 1.0f / area2, op);  // reciprocal once per triangle, not per (tile, triangle)
            rec0[3 * (size_t)idx + 2] = make_float4(rgb.x, rgb.y, rgb.z, area2);
            if (rich) {
                f3 n = cross3(t.r1v, t.r2v);
                n = n / len3(n);
                const f3 vd = mk3(t.r1v.z, t.r2v.z, t.r3v.z) + t.center_view.z;
                rec1[2 * (size_t)idx + 0] = make_float4(n.x, n.y, n.z, vd.x);
                rec1[2 * (size_t)idx + 1] = make_float4(vd.y, vd.z, 0.0f, 0.0f);
            }
            out_key = __float_as_uint(depth);
            out_rect = make_ushort4((unsigned short)rx0, (unsigned short)ry0, (unsigned short)rx1, (unsigned short)ry1);
            if (shard_world == 1) {
                out_tiles = (rx1 - rx0) * (ry1 - ry0);
            } else {
                // tiles of the rect this rank owns (tile_id % world == rank): per row, every world-th tile from a closed-form first column
                const uint32_t world = (uint32_t)shard_world, rank = (uint32_t)shard_rank;
                uint32_t n = 0;
                for (uint32_t y = ry0; y < ry1; y++) {
                    const uint32_t x0 = rx0 + (rank + world - (y * (uint32_t)gx + rx0) % world) % world;
                    n += x0 < rx1 ? (rx1 - x0 + world - 1) / world : 0u;
                }
                out_tiles = n;
            }
            out_radius = max(ceilf((vmax.x - vmin.x) * 0.5f), ceilf((vmax.y - vmin.y) * 0.5f));
        }
    } while (0);

    radii[idx] = out_radius;
    tiles[idx] = out_tiles;
    rect[idx] = out_rect;
    dkey[idx] = out_key;
}

int ts2d_launch_preprocess(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, int32_t *radii, GeomState gs, cudaStream_t s)
{
    const int P = g->P;
    const int gx = (cam->width + TS2D_TILE - 1) / TS2D_TILE, gy = (cam->height + TS2D_TILE - 1) / TS2D_TILE;
#define TS2D_K1_ARGS                                                                                                                          \
    cam->width, cam->height, P, g->sh_degree, g->M, g->C, f->rich_info != 0, g->use_shs != 0, gx, gy, f->back_culling != 0, cam->tan_fovx,    \
        cam->tan_fovy, f->shard_rank, f->shard_world, cam->viewmatrix, cam->projmatrix, cam->campos, g->vertex, g->shs, g->feature, g->opacity, \
        radii, gs.rec0, gs.rec1, gs.dkey, gs.ids, gs.tiles, gs.rect, gs.clamp, mi
    // stage whole SH rows only when at least half of each row is actually read ((D+1)^2 of M coefficients)
    const int K = (g->sh_degree + 1) * (g->sh_degree + 1);
    const ModelIn mi = ts2d_model_in(g, gs.hdr);
    if (mi.on) {
        if (mi.bg_bits) TS2D_CUDA_TRY(cudaMemsetAsync(mi.bg_bits, 0, sizeof(uint32_t), s));
        const size_t smem = (size_t)(TS2D_BLOCK / 32) * 32 * ts2d_split_row_stride(g->M) * sizeof(float);
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess<true, true>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, smem, s, TS2D_K1_ARGS));
    } else if (g->use_shs && ts2d_rows_tileable(g->M, g->shs, g->shs) && 2 * K >= g->M) {
        const size_t smem = (size_t)(TS2D_BLOCK / 32) * 32 * ((3 * g->M) / 4 + 1) * 16;
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess<true, false>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, smem, s, TS2D_K1_ARGS));
    } else {
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess<false, false>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, 0, s, TS2D_K1_ARGS));
    }
#undef TS2D_K1_ARGS
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K9: per-triangle backward.  gacc holds, per triangle, the 16 screen-space sums produced by the
// composite backward: [0..5] dL/d(v1,v2,v3)_2D, [6] dL/d opacity, [7] pad, [8..10] dL/d rgb, [11] pad... see GACC_* below.
// ------------------------------------------------------------------------------------------------
#ifndef TS2D_K9_MINB
#define TS2D_K9_MINB 3
#endif
__device__ __forceinline__ f2 grad_norm2(f2 v, f2 dv)
{
    const float sum2 = v.x * v.x + v.y * v.y;
    const float n = sqrtf(sum2);
    const float inv = 1.0f / (n * n * n);
    return mk2(((sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y) * inv, (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y) * inv);
}
__device__ __forceinline__ void project_offset_bwd(f3 p, f3 d, float tfx, float tfy, f2 g, f3 &gp, f3 &gd)
{
    const float px_pz = p.x / p.z, py_pz = p.y / p.z;
    const float vx_pz = d.x / p.z, vy_pz = d.y / p.z, vz_pz = d.z / p.z;
    const f2 gv = mk2(g.x / (p.z * tfx), g.y / (p.z * tfy));
    gd = mk3(gv.x, gv.y, -gv.x * px_pz - gv.y * py_pz);
    gp = mk3(-gv.x * vz_pz, -gv.y * vz_pz, gv.x * (2.0f * vz_pz * px_pz - vx_pz) + gv.y * (2.0f * vz_pz * py_pz - vy_pz));
}



template <bool TILED, bool MODEL>
__global__ void __launch_bounds__(TS2D_BLOCK, TS2D_K9_MINB)
k_preprocess_bwd(int W, int H, int P, int D, int M, int C, bool use_shs, bool rich, float tfx, float tfy, const float *__restrict__ view,
                 const float *__restrict__ proj, const float *__restrict__ campos, const float *__restrict__ vertex,
                 const float *__restrict__ shs, const int32_t *__restrict__ radii, const uint8_t *__restrict__ clamp,
                 const float4 *__restrict__ gacc, bool moments, const float4 *__restrict__ rec0, float *__restrict__ dL_dvertex, float *__restrict__ dL_dcenter2D, float *__restrict__ dL_dshs,
                 float *__restrict__ dL_dfeature, float *__restrict__ dL_dopacity, ModelIn mi, ModelOut mo)
{
    extern __shared__ __align__(16) float s_rows[];
    ts2d_grid_chain();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row0 = idx - lane;                       // first triangle of this warp
    const int nrows = min(32, P - row0);               // <= 0: the whole warp is past the end
    const int q = (3 * M) / 4, rs4 = q + 1;            // float4 per row / tile row stride
    const int rsf = (3 * M) | 1;                       // MODEL: row stride in floats
    float *tile_in = s_rows + (size_t)warp * 32 * (MODEL ? rsf : rs4 * 4), *tile_out = tile_in;  // one tile: SH rows in, their gradients out, in place
    const float *sh_row = shs + (size_t)idx * M * 3;
    float *gsh_row = dL_dshs + (size_t)idx * M * 3;
    if (MODEL) {
        if (nrows <= 0) return;
        if (D > 0) warp_rows_load_split(mi.f_dc, mi.f_rest, tile_in, M, rsf, row0, nrows, lane);
        __syncwarp();
        sh_row = tile_in + lane * rsf;
        gsh_row = tile_out + lane * rsf;
    } else if (TILED) {
        if (nrows <= 0) return;
        if (use_shs && D > 0) warp_rows_load(shs + (size_t)row0 * M * 3, tile_in, q, rs4, nrows, lane);
        __syncwarp();
        sh_row = tile_in + lane * rs4 * 4;
        gsh_row = tile_out + lane * rs4 * 4;
    }
    const bool live = idx < P && radii[idx] > 0;
    if (!live) {  // reference leaves its zero-initialised outputs untouched here (backward.cu:165)
        if (idx < P) {
            float *ov = dL_dvertex + 9 * (size_t)idx;
            for (int k = 0; k < 9; k++) ov[k] = 0.0f;
            dL_dcenter2D[2 * idx] = 0.0f;
            dL_dcenter2D[2 * idx + 1] = 0.0f;
            for (int k = 0; k < 3 * M; k++) gsh_row[k] = 0.0f;
            for (int k = 0; k < C; k++) dL_dfeature[(size_t)idx * C + k] = 0.0f;
            dL_dopacity[idx] = 0.0f;
        }
    } else {
    float *ov = dL_dvertex + 9 * (size_t)idx;
    const float4 A0 = gacc[4 * (size_t)idx + 0], A1 = gacc[4 * (size_t)idx + 1], A2 = gacc[4 * (size_t)idx + 2], A3 = gacc[4 * (size_t)idx + 3];
    f2 g1 = mk2(A0.x, A0.y), g2 = mk2(A0.z, A0.w), g3 = mk2(A1.x, A1.y);
    f2 gc2d = g1 + g2 + g3;  // backward.cu:191 (dL_dcenter_2D = sum of the three vertex gradients)
    if (moments) {
        // The fast composite backward accumulates S_k = sum ga_k, Q_k = sum ga_k (p - v1) (k = 1, 2; ga_k = dL/da_k - dL/da_3)
        // instead of the vertex gradients; with a1 = 1 + cross(e23, q)/A, a2 = cross(e31, q)/A (q = p - v1) the reference's
        // Jacobians (backward.cu:464-479) summed over pixels collapse to the fixed 6 -> 6 map below.
        const float4 r0 = rec0[3 * (size_t)idx], r1 = rec0[3 * (size_t)idx + 1];
        const f2 s1 = mk2(r0.x, r0.y), s2 = mk2(r0.z, r0.w), s3 = mk2(r1.x, r1.y);
        const float inv = r1.z;
        const f2 e12 = s2 - s1, e23 = s3 - s2, e31 = s1 - s3, w = s3 - s1;
        const float S1 = A0.x, S2 = A0.w;
        const f2 Q1 = mk2(A0.y, A0.z), Q2 = mk2(A1.x, A1.y);
        // D = sum ga1 (a1 - 1) + sum ga2 a2: how far the weighted pixels sit from v1 in barycentric terms.  Keeping it apart from S1
        // matters: with T = S1 + D the terms e31 T + w S1 of the vertex-2 gradient cancel down to e31 D, and for a triangle whose
        // contributing pixels lie near v1 (|D| << |S1|) the rounded cancellation used to cost vertex 2 (and 3) most of their digits
        // (measured at C3: single entries 0.7 off where the reference is exact to 1e-6).
        const float D = (cross2(e23, Q1) + cross2(e31, Q2)) * inv;
        const f2 gc = perp2(e23 * S1 + e31 * S2) * inv;  // response to a rigid translation = g1 + g2 + g3 (the Q and D terms cancel identically)
        g1 = gc + perp2(e23 * D + Q2) * inv;
        g2 = perp2(e31 * D - Q1) * inv;
        g3 = perp2(e12 * D + (Q1 - Q2)) * inv;
        (void)w;
        gc2d = gc;  // dL_dcenter2D (backward.cu:191) in closed form rather than as the sum of the three rounded vectors
    }
    const float g_op = A1.z;
    const f3 g_rgb = mk3(A2.x, A2.y, A2.z);
    const f3 g_n = mk3(A1.w, A2.w, A3.x);
    const f3 g_vd = mk3(A3.y, A3.z, A3.w);

    TriGeom t;
    f3 r1, r2, r3;
    const float *vp = vertex + 9 * (size_t)idx;
    {
        f3 a, b, c;
        load_tri<MODEL>(vp, mi, a, b, c);
        tri_geom_view(a, b, c, view, tfx, tfy, t, r1, r2, r3);
    }
    t.r1v = xf_vec(view, r1);
    t.r2v = xf_vec(view, r2);
    t.r3v = xf_vec(view, r3);
    t.r1p = project_offset(t.center_clip, t.r1v, tfx, tfy);
    t.r2p = project_offset(t.center_clip, t.r2v, tfx, tfy);
    t.r3p = project_offset(t.center_clip, t.r3v, tfx, tfy);

    const f2 scaling = mk2(0.5f * W, 0.5f * H);
    const float kernel_size = 0.5f;
    const f2 gp1 = scaling * g1 + kernel_size * grad_norm2(t.r1p, g1);
    const f2 gp2 = scaling * g2 + kernel_size * grad_norm2(t.r2p, g2);
    const f2 gp3 = scaling * g3 + kernel_size * grad_norm2(t.r3p, g3);
    const f2 gcproj = scaling * gc2d;

    f3 gr1v, gr2v, gr3v, gcc, gcv = mk3(0, 0, 0);
    project_offset_bwd(t.center_clip, t.r1v, tfx, tfy, gp1, gcc, gr1v);
    gcv = gcv + gcc;
    project_offset_bwd(t.center_clip, t.r2v, tfx, tfy, gp2, gcc, gr2v);
    gcv = gcv + gcc;
    project_offset_bwd(t.center_clip, t.r3v, tfx, tfy, gp3, gcc, gr3v);
    gcv = gcv + gcc;
    if (t.center_view.x < -t.limx || t.center_view.x > t.limx) gcv.x = 0;
    if (t.center_view.y < -t.limy || t.center_view.y > t.limy) gcv.y = 0;

    if (rich) {
        const f3 cr = cross3(t.r1v, t.r2v);
        const f3 gcr = grad_norm3(cr, g_n);
        gr1v = gr1v + (cross3(t.r2v, gcr) + mk3(0, 0, g_vd.x));
        gr2v = gr2v + (cross3(gcr, t.r1v) + mk3(0, 0, g_vd.y));
        gr3v = gr3v + mk3(0, 0, g_vd.z);
        gcv = gcv + mk3(0, 0, g_vd.x + g_vd.y + g_vd.z);
    }

    f3 gcenter;
    {
        const float4 h = xf_hom(proj, t.center);
        const float winv = 1.0f / (fabsf(h.w) + TS2D_EPS);
        const f3 pp = mk3(h.x * winv, h.y * winv, h.z * winv);
        const f3 gpp = mk3(gcproj.x, gcproj.y, 0);
        const float s = fabsf(winv);
        const float hx = s * gpp.x, hy = s * gpp.y, hz = s * gpp.z, hw = s * (-dot3(gpp, pp));
        gcenter = mk3(proj[0] * hx + proj[1] * hy + proj[2] * hz + proj[3] * hw, proj[4] * hx + proj[5] * hy + proj[6] * hz + proj[7] * hw,
                      proj[8] * hx + proj[9] * hy + proj[10] * hz + proj[11] * hw);
    }
    gcenter = gcenter + xf_vec_T(view, gcv);
    const f3 gr1 = xf_vec_T(view, gr1v), gr2 = xf_vec_T(view, gr2v), gr3 = xf_vec_T(view, gr3v);

    if (use_shs) {
        const f3 gsh = sh_colour_bwd(D, M, sh_row, t.center, ld3(campos), clamp[idx], g_rgb, gsh_row);
        gcenter = gcenter + gsh;
    } else {
        for (int k = 0; k < 3 * M; k++) gsh_row[k] = 0.0f;
    }
    {
        f3 gv1 = (2 * gr1 - gr2 - gr3 + gcenter) / 3.0f, gv2 = (2 * gr2 - gr1 - gr3 + gcenter) / 3.0f, gv3 = (2 * gr3 - gr1 - gr2 + gcenter) / 3.0f;
        if (MODEL) rescale_bwd(mi, gv1, gv2, gv3);
        st3(ov, gv1);
        st3(ov + 3, gv2);
        st3(ov + 6, gv3);
    }
    dL_dcenter2D[2 * idx] = gc2d.x;
    dL_dcenter2D[2 * idx + 1] = gc2d.y;
    dL_dfeature[(size_t)idx * C + 0] = g_rgb.x;
    if (C > 1) dL_dfeature[(size_t)idx * C + 1] = g_rgb.y;
    if (C > 2) dL_dfeature[(size_t)idx * C + 2] = g_rgb.z;
    if (MODEL) {  // SigmoidBackward: grad * (1 - y) * y; the straight-through term passes the gradient unchanged
        const float y = sigmoid_rn(mi.logit[idx]);
        dL_dopacity[idx] = g_op * (1.0f - y) * y;
        model_statistics(mo, idx, gc2d.x, gc2d.y, radii[idx]);
    } else {
        dL_dopacity[idx] = g_op;
    }
    }  // live
    if (MODEL) {
        __syncwarp();
        warp_rows_store_split(mo.g_dc, mo.g_rest, tile_out, M, rsf, row0, nrows, lane);
    } else if (TILED) {
        __syncwarp();
        warp_rows_store(dL_dshs + (size_t)row0 * M * 3, tile_out, q, rs4, nrows, lane);
    }
}

int ts2d_launch_preprocess_bwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, const int32_t *radii, GeomState gs,
                               const float *gacc, const ts2d_backward_out *out, cudaStream_t s)
{
    const int P = g->P;
    const ModelIn mi = ts2d_model_in(g, gs.hdr);
    const ModelOut mo = ts2d_model_out(out);
    const bool tiled = !mi.on && ts2d_rows_tileable(g->M, g->shs ? (const void *)g->shs : (const void *)out->dL_dshs, out->dL_dshs);
#define TS2D_K9_ARGS                                                                                                                         \
    cam->width, cam->height, P, g->sh_degree, g->M, g->C, g->use_shs != 0, f->rich_info != 0, cam->tan_fovx, cam->tan_fovy, cam->viewmatrix, \
        cam->projmatrix, cam->campos, g->vertex, g->shs, radii, gs.clamp, (const float4 *)gacc, ts2d_use_fast(g, f), gs.rec0, out->dL_dvertex, \
        out->dL_dcenter2D, out->dL_dshs, out->dL_dfeature, out->dL_dopacity, mi, mo
    if (mi.on) {
        const size_t smem = (size_t)(TS2D_BLOCK / 32) * 32 * ts2d_split_row_stride(g->M) * sizeof(float);
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess_bwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess_bwd<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess_bwd<true, true>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, smem, s, TS2D_K9_ARGS));
    } else if (tiled) {
        const size_t smem = (size_t)(TS2D_BLOCK / 32) * 32 * ((3 * g->M) / 4 + 1) * 16;
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess_bwd<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess_bwd<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess_bwd<true, false>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, smem, s, TS2D_K9_ARGS));
    } else {
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess_bwd<false, false>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, 0, s, TS2D_K9_ARGS));
    }
#undef TS2D_K9_ARGS
    return (int)cudaGetLastError();
}
