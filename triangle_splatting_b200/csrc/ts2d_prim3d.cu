// ts2d_prim3d.cu -- the reference's second per-pixel primitive ("3D": ray / triangle-plane intersection in view space)
// on the same tile pipeline (flags.primitive == TS2D_PRIMITIVE_3D).
//
// Replaces, for R3D = submodules/diff-triangle-rasterization-3D (same pybind API and the same binning / sort host
// code as R2D -- `diff R2D/src/rasterizer.cu R3D/src/rasterizer.cu` differs only in which per-triangle arrays are passed):
//   k_preprocess3d       R3D/src/forward.cu:61-146    FORWARD::preprocessCUDA
//   k_render3d_fwd       R3D/src/forward.cu:151-306   FORWARD::renderCUDA
//   k_render3d_bwd       R3D/src/backward.cu:215-454  BACKWARD::renderCUDA
//   k_preprocess3d_bwd   R3D/src/backward.cu:144-213  BACKWARD::preprocessCUDA
// Depth ordering, key emission, the tile sort and the tile ranges are the primitive-independent kernels of ts2d_binning.cu.
//
// Record layout (geometry state; 64 B per triangle in the two record arrays of GeomState):
//   rec0[3i+0] = {v1v.x v1v.y v1v.z v2v.x}   rec0[3i+1] = {v2v.y v2v.z v3v.x v3v.y}   rec0[3i+2] = {v3v.z n.x n.y n.z}
//   rec1[2i+0] = {r g b opacity}              rec1[2i+1] = {K_fwd, 1 / dot(n, n), K_bwd, -}  (fast kernels only; ts2d_prim3d.cuh)
// with v_k_view = W2C v_k and n = (v2v - v1v) x (v3v - v1v), NOT normalised (R3D/src/forward.cu:94).
//
// Quirks of the reference that are kept on purpose (a drop-in must reproduce the reference's numbers):
//   * forward skips a pair when alpha = min(0.99, op G) < 1/255, backward when G < 1/255 (R3D/src/forward.cu:274-276 vs
//     R3D/src/backward.cu:348-352): the reverse walk un-blends pairs the forward walk never blended when op < 1;
//   * forward computes depth = dot(v1, n) / (ray . n), backward dot(v1, n) * (1 / (ray . n)) (:246 vs :334-335);
//   * out_normal accumulates the un-normalised plane normal (:283).
// The composite backward uses the same reduction as the 2D mirror kernel: the 16 per-pair gradient components
// (9 view-space vertex, 3 normal, 3 colour, 1 opacity -- exactly one 64 B accumulator line) are summed over the warp's
// 32 pixels with a recursive-halving butterfly and leave as one 64 B RED burst per (warp, triangle) instead of the
// reference's 16 scalar atomics per (pixel, triangle) (R3D/src/backward.cu:365,430-451).
#include "ts2d_prim3d.cuh"
#include "ts2d_sh.cuh"

namespace {

// tiles of the rect [rx0, rx1) x [ry0, ry1) this rank owns (tile_id % world == rank)
__device__ __forceinline__ uint32_t owned_tiles(uint32_t rx0, uint32_t ry0, uint32_t rx1, uint32_t ry1, int gx, int shard_rank, int shard_world)
{
    if (shard_world == 1) return (rx1 - rx0) * (ry1 - ry0);
    const uint32_t world = (uint32_t)shard_world, rank = (uint32_t)shard_rank;
    uint32_t n = 0;
    for (uint32_t y = ry0; y < ry1; y++) {
        const uint32_t x0 = rx0 + (rank + world - (y * (uint32_t)gx + rx0) % world) % world;
        n += x0 < rx1 ? (rx1 - x0 + world - 1) / world : 0u;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------ K1 (3D)
template <bool TILED, bool MODEL>  // MODEL: parameter-space inputs, see ts2d_preprocess.cu / ts2d_sh.cuh
__global__ void __launch_bounds__(TS2D_BLOCK)
k_preprocess3d(int W, int H, int P, int D, int M, int C, bool use_shs, int gx, int gy, bool back_culling, int shard_rank, int shard_world,
               const float *__restrict__ view, const float *__restrict__ proj, const float *__restrict__ campos,
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
        const float *vp = vertex + 9 * (size_t)idx;
        f3 v1, v2, v3;
        load_tri<MODEL>(vp, mi, v1, v2, v3);
        const f3 v1v = xf_point(view, v1), v2v = xf_point(view, v2), v3v = xf_point(view, v3);
        const f3 center_view = (v1v + v2v + v3v) / 3.0f;
        const f3 n = cross3(v2v - v1v, v3v - v1v);
        if (len3(n) < TS2D_EPS) break;            // degenerate
        if (back_culling && n.z >= 0) break;      // back face
        const float dilation = 3.0f;
        const f3 center = (v1 + v2 + v3) / 3.0f;
        const f3 d1 = center + dilation * (v1 - center), d2 = center + dilation * (v2 - center), d3 = center + dilation * (v3 - center);
        const f3 p1 = project_center(proj, d1), p2 = project_center(proj, d2), p3 = project_center(proj, d3);
        if (p1.z <= 0 || p2.z <= 0 || p3.z <= 0) break;  // near culling on the dilated vertices
        const f2 s1 = mk2(proj_to_pix(p1.x, W), proj_to_pix(p1.y, H));
        const f2 s2 = mk2(proj_to_pix(p2.x, W), proj_to_pix(p2.y, H));
        const f2 s3 = mk2(proj_to_pix(p3.x, W), proj_to_pix(p3.y, H));
        const f2 vmin = mk2(fminf(fminf(s1.x, s2.x), s3.x), fminf(fminf(s1.y, s2.y), s3.y));
        const f2 vmax = mk2(fmaxf(fmaxf(s1.x, s2.x), s3.x), fmaxf(fmaxf(s1.y, s2.y), s3.y));
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
        rec0[3 * (size_t)idx + 0] = make_float4(v1v.x, v1v.y, v1v.z, v2v.x);
        rec0[3 * (size_t)idx + 1] = make_float4(v2v.y, v2v.z, v3v.x, v3v.y);
        rec0[3 * (size_t)idx + 2] = make_float4(v3v.z, n.x, n.y, n.z);
        float sig;
        rec1[2 * (size_t)idx + 0] = make_float4(rgb.x, rgb.y, rgb.z, MODEL ? model_opacity(mi, idx, sig) : opacity[idx]);
        {   // the two per-triangle subexpressions of the per-pair arithmetic (ts2d_prim3d.cuh: geo3), from the STORED values
            Tri3 t;
            t.v1 = v1v; t.v2 = v2v; t.v3 = v3v; t.n = n;
            rec1[2 * (size_t)idx + 1] = make_float4(tri3_K<false>(t), tri3_inv_nn(t), tri3_K<true>(t), 0.0f);
        }
        out_key = __float_as_uint(center_view.z);
        out_rect = make_ushort4((unsigned short)rx0, (unsigned short)ry0, (unsigned short)rx1, (unsigned short)ry1);
        out_tiles = owned_tiles(rx0, ry0, rx1, ry1, gx, shard_rank, shard_world);
        out_radius = max(ceilf((vmax.x - vmin.x) * 0.5f), ceilf((vmax.y - vmin.y) * 0.5f));
    } while (0);

    radii[idx] = out_radius;
    tiles[idx] = out_tiles;
    rect[idx] = out_rect;
    dkey[idx] = out_key;
}

// ------------------------------------------------------------------------------------------------ K7 (3D)
template <bool RICH>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_render3d_fwd(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float two_gamma, float tfx, float tfy,
               const uint2 *__restrict__ ranges, const uint32_t *__restrict__ list, const float4 *__restrict__ rec0,
               const float4 *__restrict__ rec1, float bg_depth, const float *__restrict__ bg_ptr, const float *__restrict__ background, float *__restrict__ final_T,
               uint32_t *__restrict__ n_contrib, float *__restrict__ out_feature, float *__restrict__ out_depth,
               float *__restrict__ out_normal, float *__restrict__ contrib_sum, float *__restrict__ contrib_max)
{
    if (bg_ptr) bg_depth = __ldg(bg_ptr);  // model inputs: background depth computed on the device by K1
    __shared__ float4 s_rec[TS2D_BLOCK * 4];
    __shared__ uint32_t s_id[TS2D_BLOCK];

    const int tile = blockIdx.x * shard_world + shard_rank;
    if (tile >= n_tiles) return;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int px = tile_x * TS2D_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TS2D_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const size_t pix = (size_t)W * py + px;
    const f3 ray = mk3(tfx * pix_to_proj((float)px, W), tfy * pix_to_proj((float)py, H), 1.0f);

    const uint2 range = ranges[tile];
    float T = 1.0f;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f;
    f3 accn = mk3(0.f, 0.f, 0.f);
    uint32_t last = 0;
    bool done = !inside;

    for (uint32_t base = range.x; base < range.y; base += TS2D_BLOCK) {
        if (__syncthreads_and(done)) break;
        const int n = min((uint32_t)TS2D_BLOCK, range.y - base);
        if (tid < n) {
            const uint32_t id = list[base + tid];
            s_id[tid] = id;
            const float4 *r = rec0 + 3 * (size_t)id;
            s_rec[4 * tid + 0] = __ldg(r);
            s_rec[4 * tid + 1] = __ldg(r + 1);
            s_rec[4 * tid + 2] = __ldg(r + 2);
            s_rec[4 * tid + 3] = __ldg(rec1 + 2 * (size_t)id);
        }
        __syncthreads();

        for (int j = 0; j < n; j++) {
            if (__all_sync(0xffffffffu, done)) break;
            bool hit = false;
            float contrib = 0.0f;
            if (!done) {
                last = base - range.x + j + 1;
                const Tri3 t = unpack3(s_rec[4 * j], s_rec[4 * j + 1], s_rec[4 * j + 2]);
                const float4 col = s_rec[4 * j + 3];
                Pair3 e;
                if (eval_pair3<false>(t, tri3_K<false>(t), tri3_inv_nn(t), col.w, two_gamma, ray, e)) {
                    hit = true;
                    contrib = e.alpha * T;
                    T *= (1.0f - e.alpha);
                    acc0 += col.x * contrib;
                    acc1 += col.y * contrib;
                    acc2 += col.z * contrib;
                    if (RICH) {
                        accn = accn + t.n * contrib;
                        accd += e.depth * contrib;
                    }
                    if (T <= 0.0001f) done = true;
                }
            }
            if (RICH) {
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    const float s = warp_sum(contrib);
                    const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(contrib));  // contrib >= 0: bit order == value order
                    if (lane == 0) {
                        const uint32_t id = s_id[j];
                        atomicAdd(contrib_sum + id, s);
                        atomicMax((unsigned int *)contrib_max + id, mx);
                    }
                }
            }
        }
    }

    if (inside) {
        const float bg0 = background[0], bg1 = C > 1 ? background[1] : 0.f, bg2 = C > 2 ? background[2] : 0.f;
        final_T[pix] = T;
        n_contrib[pix] = last;
        const size_t HW = (size_t)H * W;
        out_feature[pix] = acc0 + T * bg0;
        if (C > 1) out_feature[HW + pix] = acc1 + T * bg1;
        if (C > 2) out_feature[2 * HW + pix] = acc2 + T * bg2;
        if (RICH) {
            out_depth[pix] = accd + T * bg_depth;
            out_normal[pix] = accn.x;
            out_normal[HW + pix] = accn.y;
            out_normal[2 * HW + pix] = accn.z;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K8 (3D)
// Sum v[0..15] over the warp; on return lane L holds component (L >> 1) & 15 (same butterfly as ts2d_render_bwd.cu).
__device__ __forceinline__ float warp_reduce16_3d(float (&v)[16], int lane)
{
    float w8[8], w4[4], w2[2];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float keep = up ? v[i + 8] : v[i];
            const float send = up ? v[i] : v[i + 8];
            w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float keep = up ? w8[i + 4] : w8[i];
            const float send = up ? w8[i] : w8[i + 4];
            w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float keep = up ? w4[i + 2] : w4[i];
            const float send = up ? w4[i] : w4[i + 2];
            w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    float r;
    {
        const bool up = lane & 2;
        const float keep = up ? w2[1] : w2[0];
        const float send = up ? w2[0] : w2[1];
        r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// Accumulator line of the 3D primitive (GACC_STRIDE floats per triangle):
//   [0..2] dL/d v1_view   [3..5] dL/d v2_view   [6..8] dL/d v3_view   [9..11] dL/d normal_view   [12..14] dL/d rgb   [15] dL/d opacity
template <bool RICH>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_render3d_bwd(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, float tfx, float tfy,
               const uint2 *__restrict__ ranges, const uint32_t *__restrict__ list, const float4 *__restrict__ rec0,
               const float4 *__restrict__ rec1, float bg_depth, const float *__restrict__ bg_ptr, const float *__restrict__ background, const float *__restrict__ final_T,
               const uint32_t *__restrict__ n_contrib, const float *__restrict__ dL_dout_feature, const float *__restrict__ dL_dout_depth,
               const float *__restrict__ dL_dout_normal, float *__restrict__ gacc)
{
    if (bg_ptr) bg_depth = __ldg(bg_ptr);  // model inputs: background depth computed on the device by K1
    __shared__ float4 s_rec[TS2D_BLOCK * 4];
    __shared__ uint32_t s_id[TS2D_BLOCK];

    const int tile = blockIdx.x * shard_world + shard_rank;
    if (tile >= n_tiles) return;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int px = tile_x * TS2D_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TS2D_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const size_t pix = (size_t)W * py + px;
    const size_t HW = (size_t)H * W;
    const float two_gamma = 2.0f * gamma;
    const f3 ray = mk3(tfx * pix_to_proj((float)px, W), tfy * pix_to_proj((float)py, H), 1.0f);

    const uint2 range = ranges[tile];
    const uint32_t len = range.y - range.x;
    float T = inside ? final_T[pix] : 0.0f;
    const uint32_t last = inside ? n_contrib[pix] : 0u;

    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = bg_depth;
    f3 accn = mk3(0.f, 0.f, 0.f);
    float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f, gd = 0.f;
    f3 gn = mk3(0.f, 0.f, 0.f);
    if (inside) {
        acc0 = background[0];
        gp0 = dL_dout_feature[pix];
        if (C > 1) { acc1 = background[1]; gp1 = dL_dout_feature[HW + pix]; }
        if (C > 2) { acc2 = background[2]; gp2 = dL_dout_feature[2 * HW + pix]; }
        if (RICH) {
            gn = mk3(dL_dout_normal[pix], dL_dout_normal[HW + pix], dL_dout_normal[2 * HW + pix]);
            gd = dL_dout_depth[pix];
        }
    }
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);

    for (uint32_t done_cnt = 0; done_cnt < len; done_cnt += TS2D_BLOCK) {
        __syncthreads();
        const int n = min((uint32_t)TS2D_BLOCK, len - done_cnt);
        if (tid < n) {
            const uint32_t id = list[range.y - 1 - done_cnt - tid];  // reversed order
            s_id[tid] = id;
            const float4 *r = rec0 + 3 * (size_t)id;
            s_rec[4 * tid + 0] = __ldg(r);
            s_rec[4 * tid + 1] = __ldg(r + 1);
            s_rec[4 * tid + 2] = __ldg(r + 2);
            s_rec[4 * tid + 3] = __ldg(rec1 + 2 * (size_t)id);
        }
        __syncthreads();

        for (int j = 0; j < n; j++) {
            const uint32_t pos = len - 1 - done_cnt - j;  // 0-based list position of this entry
            if (pos >= warp_last) continue;                // warp-uniform
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = 0.0f;
            bool hit = false;
            if (pos < last) {
                const Tri3 t = unpack3(s_rec[4 * j], s_rec[4 * j + 1], s_rec[4 * j + 2]);
                const float4 col = s_rec[4 * j + 3];
                const float op = col.w;
                Pair3 e;
                if (eval_pair3<true>(t, tri3_K<true>(t), tri3_inv_nn(t), op, two_gamma, ray, e)) {
                    hit = true;
                    T /= (1.0f - e.alpha);
                    const float contrib = e.alpha * T;
                    float dL_dcontrib = 0.0f;
                    f3 dL_dnormal = mk3(0.f, 0.f, 0.f);
                    float dL_ddepth = 0.0f;

                    v[12] = gp0 * contrib;
                    dL_dcontrib += gp0 * (col.x - acc0);
                    acc0 = e.alpha * col.x + (1.0f - e.alpha) * acc0;
                    if (C > 1) {
                        v[13] = gp1 * contrib;
                        dL_dcontrib += gp1 * (col.y - acc1);
                        acc1 = e.alpha * col.y + (1.0f - e.alpha) * acc1;
                    }
                    if (C > 2) {
                        v[14] = gp2 * contrib;
                        dL_dcontrib += gp2 * (col.z - acc2);
                        acc2 = e.alpha * col.z + (1.0f - e.alpha) * acc2;
                    }
                    if (RICH) {
                        dL_dnormal = dL_dnormal + gn * contrib;
                        dL_dcontrib += dot3(gn, t.n - accn);
                        accn = e.alpha * t.n + (1.0f - e.alpha) * accn;
                        dL_ddepth += gd * contrib;
                        dL_dcontrib += gd * (e.depth - accd);
                        accd = e.alpha * e.depth + (1.0f - e.alpha) * accd;
                    }

                    const float dL_dalpha = dL_dcontrib * T;
                    const float dL_dpower = (op * e.G < 0.99f) ? (dL_dalpha * e.alpha) : 0.0f;
                    const float dL_decc = dL_dpower * 2 * gamma * e.power / (e.ecc + TS2D_EPS);
                    f3 decc_da = mk3(0.f, 0.f, 0.f);
                    if (e.a1 <= e.a2 && e.a1 <= e.a3) decc_da.x = -3.0f;
                    else if (e.a2 <= e.a1 && e.a2 <= e.a3) decc_da.y = -3.0f;
                    else decc_da.z = -3.0f;
                    const f3 dL_da = dL_decc * decc_da;

                    // R3D/src/backward.cu:401-428
                    const f3 z3 = mk3(0.f, 0.f, 0.f);
                    const f3 da1_dv1 = z3;
                    const f3 da1_dv2 = cross3(e.pv3, t.n) * e.inv_nn;
                    const f3 da1_dv3 = cross3(t.n, e.pv2) * e.inv_nn;
                    const f3 da1_dn = (cross3(e.pv2, e.pv3) - (2.0f * e.a1) * t.n) * e.inv_nn;
                    const float da1_dd = dot3(t.n, cross3(t.v3 - t.v2, ray)) * e.inv_nn;

                    const f3 da2_dv1 = cross3(t.n, e.pv3) * e.inv_nn;
                    const f3 da2_dv2 = z3;
                    const f3 da2_dv3 = cross3(e.pv1, t.n) * e.inv_nn;
                    const f3 da2_dn = (cross3(e.pv3, e.pv1) - (2.0f * e.a2) * t.n) * e.inv_nn;
                    const float da2_dd = dot3(t.n, cross3(t.v1 - t.v3, ray)) * e.inv_nn;

                    const f3 da3_dv1 = -da1_dv1 - da2_dv1;
                    const f3 da3_dv2 = -da1_dv2 - da2_dv2;
                    const f3 da3_dv3 = -da1_dv3 - da2_dv3;
                    const f3 da3_dn = -da1_dn - da2_dn;
                    const float da3_dd = -da1_dd - da2_dd;

                    dL_ddepth += dL_da.x * da1_dd + dL_da.y * da2_dd + dL_da.z * da3_dd;
                    const f3 dd_dv1 = t.n * e.inv_pn;
                    const f3 dd_dn = (t.v1 - e.depth * ray) * e.inv_pn;

                    const f3 g1 = dL_da.x * da1_dv1 + dL_da.y * da2_dv1 + dL_da.z * da3_dv1 + dL_ddepth * dd_dv1;
                    const f3 g2 = dL_da.x * da1_dv2 + dL_da.y * da2_dv2 + dL_da.z * da3_dv2;
                    const f3 g3 = dL_da.x * da1_dv3 + dL_da.y * da2_dv3 + dL_da.z * da3_dv3;
                    dL_dnormal = dL_dnormal + (dL_da.x * da1_dn + dL_da.y * da2_dn + dL_da.z * da3_dn + dL_ddepth * dd_dn);

                    v[0] = g1.x; v[1] = g1.y; v[2] = g1.z;
                    v[3] = g2.x; v[4] = g2.y; v[5] = g2.z;
                    v[6] = g3.x; v[7] = g3.y; v[8] = g3.z;
                    v[9] = dL_dnormal.x; v[10] = dL_dnormal.y; v[11] = dL_dnormal.z;
                    v[15] = dL_dalpha * e.G;
                }
            }
            if (__ballot_sync(0xffffffffu, hit) == 0u) continue;
            const float r = warp_reduce16_3d(v, lane);
            if ((lane & 1) == 0) atomicAdd(gacc + (size_t)s_id[j] * GACC_STRIDE + (lane >> 1), r);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K9 (3D)
template <bool TILED, bool MODEL>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_preprocess3d_bwd(int P, int D, int M, int C, bool use_shs, const float *__restrict__ view, const float *__restrict__ campos,
                   const float *__restrict__ vertex, const float *__restrict__ shs, const int32_t *__restrict__ radii,
                   const uint8_t *__restrict__ clamp, const float4 *__restrict__ gacc, const float4 *__restrict__ rec0,
                   float *__restrict__ dL_dvertex, float *__restrict__ dL_dcenter2D, float *__restrict__ dL_dshs,
                   float *__restrict__ dL_dfeature, float *__restrict__ dL_dopacity, ModelIn mi, ModelOut mo)
{
    extern __shared__ __align__(16) float s_rows[];
    ts2d_grid_chain();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row0 = idx - lane;
    const int nrows = min(32, P - row0);
    const int q = (3 * M) / 4, rs4 = q + 1;
    const int rsf = (3 * M) | 1;
    float *tile = s_rows + (size_t)warp * 32 * (MODEL ? rsf : rs4 * 4);  // SH rows in, their gradients out, in place
    const float *sh_row = shs + (size_t)idx * M * 3;
    float *gsh_row = dL_dshs + (size_t)idx * M * 3;
    if (MODEL) {
        if (nrows <= 0) return;
        if (D > 0) warp_rows_load_split(mi.f_dc, mi.f_rest, tile, M, rsf, row0, nrows, lane);
        __syncwarp();
        sh_row = tile + lane * rsf;
        gsh_row = tile + lane * rsf;
    } else if (TILED) {
        if (nrows <= 0) return;
        if (use_shs && D > 0) warp_rows_load(shs + (size_t)row0 * M * 3, tile, q, rs4, nrows, lane);
        __syncwarp();
        sh_row = tile + lane * rs4 * 4;
        gsh_row = tile + lane * rs4 * 4;
    }
    const bool live = idx < P && radii[idx] > 0;
    if (!live) {  // the reference leaves its zero-initialised outputs untouched (R3D/src/backward.cu:168-169)
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
        const float4 A0 = gacc[4 * (size_t)idx + 0], A1 = gacc[4 * (size_t)idx + 1], A2 = gacc[4 * (size_t)idx + 2], A3 = gacc[4 * (size_t)idx + 3];
        const Tri3 t = unpack3(rec0[3 * (size_t)idx], rec0[3 * (size_t)idx + 1], rec0[3 * (size_t)idx + 2]);
        f3 g1v = mk3(A0.x, A0.y, A0.z), g2v = mk3(A0.w, A1.x, A1.y), g3v = mk3(A1.z, A1.w, A2.x);
        const f3 gN = mk3(A2.y, A2.z, A2.w);
        const f3 g_rgb = mk3(A3.x, A3.y, A3.z);
        const float g_op = A3.w;
        g1v = g1v + cross3(t.v2 - t.v3, gN);
        g2v = g2v + cross3(t.v3 - t.v1, gN);
        g3v = g3v + cross3(t.v1 - t.v2, gN);
        f3 g1 = xf_vec_T(view, g1v), g2 = xf_vec_T(view, g2v), g3 = xf_vec_T(view, g3v);
        if (use_shs) {
            const float *vp = vertex + 9 * (size_t)idx;
            f3 w1, w2, w3;
            load_tri<MODEL>(vp, mi, w1, w2, w3);
            const f3 center = (w1 + w2 + w3) / 3.0f;
            const f3 gsh = sh_colour_bwd(D, M, sh_row, center, ld3(campos), clamp[idx], g_rgb, gsh_row);
            g1 = g1 + gsh / 3.0f;
            g2 = g2 + gsh / 3.0f;
            g3 = g3 + gsh / 3.0f;
        } else {
            for (int k = 0; k < 3 * M; k++) gsh_row[k] = 0.0f;
        }
        float *ov = dL_dvertex + 9 * (size_t)idx;
        const f3 gcv = xf_vec(view, g1 + g2 + g3);  // w.r.t. the vertices the rasterizer saw (the rescaled ones with model inputs)
        if (MODEL) rescale_bwd(mi, g1, g2, g3);
        st3(ov, g1);
        st3(ov + 3, g2);
        st3(ov + 6, g3);
        dL_dcenter2D[2 * idx] = gcv.x;
        dL_dcenter2D[2 * idx + 1] = gcv.y;
        dL_dfeature[(size_t)idx * C + 0] = g_rgb.x;
        if (C > 1) dL_dfeature[(size_t)idx * C + 1] = g_rgb.y;
        if (C > 2) dL_dfeature[(size_t)idx * C + 2] = g_rgb.z;
        if (MODEL) {
            const float y = sigmoid_rn(mi.logit[idx]);
            dL_dopacity[idx] = g_op * (1.0f - y) * y;
            model_statistics(mo, idx, gcv.x, gcv.y, radii[idx]);
        } else {
            dL_dopacity[idx] = g_op;
        }
    }
    if (MODEL) {
        __syncwarp();
        warp_rows_store_split(mo.g_dc, mo.g_rest, tile, M, rsf, row0, nrows, lane);
    } else if (TILED) {
        __syncwarp();
        warp_rows_store(dL_dshs + (size_t)row0 * M * 3, tile, q, rs4, nrows, lane);
    }
}

__global__ void k_export_geometry3d(int P, const float4 *rec0, const float4 *rec1, const uint32_t *dkey, const uint32_t *tiles, const ushort4 *rect,
                                    const uint8_t *clamp, float *v_view, float *normal_view, float *depth, float *rgb, uint8_t *clamped,
                                    uint32_t *tiles_touched, uint32_t *rect_min, uint32_t *rect_max)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool vis = dkey[i] != 0xFFFFFFFFu;
    float4 a = make_float4(0, 0, 0, 0), b = a, c = a, q = a;
    if (vis) {
        a = rec0[3 * (size_t)i];
        b = rec0[3 * (size_t)i + 1];
        c = rec0[3 * (size_t)i + 2];
        q = rec1[2 * (size_t)i];
    }
    if (v_view) {
        float *o = v_view + 9 * (size_t)i;
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w; o[8] = c.x;
    }
    if (normal_view) { normal_view[3 * i] = c.y; normal_view[3 * i + 1] = c.z; normal_view[3 * i + 2] = c.w; }
    if (depth) depth[i] = vis ? __uint_as_float(dkey[i]) : 0.0f;
    if (rgb) { rgb[3 * i] = q.x; rgb[3 * i + 1] = q.y; rgb[3 * i + 2] = q.z; }
    const uint8_t m = vis ? clamp[i] : 0;
    if (clamped) { clamped[3 * i] = m & 1; clamped[3 * i + 1] = (m >> 1) & 1; clamped[3 * i + 2] = (m >> 2) & 1; }
    if (tiles_touched) tiles_touched[i] = tiles[i];
    const ushort4 r = rect[i];
    if (rect_min) { rect_min[2 * i] = r.x; rect_min[2 * i + 1] = r.y; }
    if (rect_max) { rect_max[2 * i] = r.z; rect_max[2 * i + 1] = r.w; }
}

}  // namespace

int ts2d_launch_preprocess3d(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, int32_t *radii, GeomState gs, cudaStream_t s)
{
    const int P = g->P;
    const int gx = (cam->width + TS2D_TILE - 1) / TS2D_TILE, gy = (cam->height + TS2D_TILE - 1) / TS2D_TILE;
#define TS2D_K1_ARGS                                                                                                                      \
    cam->width, cam->height, P, g->sh_degree, g->M, g->C, g->use_shs != 0, gx, gy, f->back_culling != 0, f->shard_rank, f->shard_world,    \
        cam->viewmatrix, cam->projmatrix, cam->campos, g->vertex, g->shs, g->feature, g->opacity, radii, gs.rec0, gs.rec1, gs.dkey, gs.ids, \
        gs.tiles, gs.rect, gs.clamp, mi
    const int K = (g->sh_degree + 1) * (g->sh_degree + 1);
    const ModelIn mi = ts2d_model_in(g, gs.hdr);
    if (mi.on) {
        if (mi.bg_bits) TS2D_CUDA_TRY(cudaMemsetAsync(mi.bg_bits, 0, sizeof(uint32_t), s));
        const size_t smem = (size_t)(TS2D_BLOCK / 32) * 32 * ts2d_split_row_stride(g->M) * sizeof(float);
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess3d<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess3d<true, true>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, smem, s, TS2D_K1_ARGS));
    } else if (g->use_shs && ts2d_rows_tileable(g->M, g->shs, g->shs) && 2 * K >= g->M) {
        const size_t smem = (size_t)(TS2D_BLOCK / 32) * 32 * ((3 * g->M) / 4 + 1) * 16;
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess3d<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess3d<true, false>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, smem, s, TS2D_K1_ARGS));
    } else {
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess3d<false, false>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, 0, s, TS2D_K1_ARGS));
    }
#undef TS2D_K1_ARGS
    return (int)cudaGetLastError();
}

int ts2d_launch_render3d_fwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                             ImageState is, const ts2d_forward_out *out, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    if (owned <= 0) return 0;
    const float two_gamma = 2.0f * g->gamma;
#define TS2D_F3_ARGS                                                                                                                        \
    W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, two_gamma, cam->tan_fovx, cam->tan_fovy, is.ranges, list, gs.rec0, gs.rec1,      \
        g->background_depth, ts2d_bg_ptr(g, gs), g->background, is.final_T, is.n_contrib, out->out_feature
    if (f->rich_info) {
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_sum, 0, sizeof(float) * (size_t)g->P, s));
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_max, 0, sizeof(float) * (size_t)g->P, s));
        k_render3d_fwd<true><<<owned, TS2D_BLOCK, 0, s>>>(TS2D_F3_ARGS, out->depth, out->normal, out->contrib_sum, out->contrib_max);
    } else {
        k_render3d_fwd<false><<<owned, TS2D_BLOCK, 0, s>>>(TS2D_F3_ARGS, nullptr, nullptr, nullptr, nullptr);
    }
#undef TS2D_F3_ARGS
    return (int)cudaGetLastError();
}

int ts2d_launch_render3d_bwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                             ImageState is, const ts2d_loss_in *loss, float *gacc, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    TS2D_CUDA_TRY(cudaMemsetAsync(gacc, 0, sizeof(float) * GACC_STRIDE * (size_t)g->P, s));
    if (owned <= 0) return 0;
#define TS2D_B3_ARGS                                                                                                                        \
    W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, cam->tan_fovx, cam->tan_fovy, is.ranges, list, gs.rec0, gs.rec1,       \
        g->background_depth, ts2d_bg_ptr(g, gs), g->background, is.final_T, is.n_contrib, loss->dL_dout_feature
    if (f->rich_info) k_render3d_bwd<true><<<owned, TS2D_BLOCK, 0, s>>>(TS2D_B3_ARGS, loss->dL_dout_depth, loss->dL_dout_normal, gacc);
    else k_render3d_bwd<false><<<owned, TS2D_BLOCK, 0, s>>>(TS2D_B3_ARGS, nullptr, nullptr, gacc);
#undef TS2D_B3_ARGS
    return (int)cudaGetLastError();
}

int ts2d_launch_preprocess3d_bwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, const int32_t *radii, GeomState gs,
                                 const float *gacc, const ts2d_backward_out *out, cudaStream_t s)
{
    (void)f;
    const int P = g->P;
    const ModelIn mi = ts2d_model_in(g, gs.hdr);
    const ModelOut mo = ts2d_model_out(out);
    const bool tiled = !mi.on && ts2d_rows_tileable(g->M, g->shs ? (const void *)g->shs : (const void *)out->dL_dshs, out->dL_dshs);
#define TS2D_K9_ARGS                                                                                                                   \
    P, g->sh_degree, g->M, g->C, g->use_shs != 0, cam->viewmatrix, cam->campos, g->vertex, g->shs, radii, gs.clamp, (const float4 *)gacc, \
        gs.rec0, out->dL_dvertex, out->dL_dcenter2D, out->dL_dshs, out->dL_dfeature, out->dL_dopacity, mi, mo
    if (mi.on) {
        const size_t smem = (size_t)(TS2D_BLOCK / 32) * 32 * ts2d_split_row_stride(g->M) * sizeof(float);
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess3d_bwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess3d_bwd<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess3d_bwd<true, true>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, smem, s, TS2D_K9_ARGS));
    } else if (tiled) {
        const size_t smem = (size_t)(TS2D_BLOCK / 32) * 32 * ((3 * g->M) / 4 + 1) * 16;
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess3d_bwd<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_preprocess3d_bwd<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess3d_bwd<true, false>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, smem, s, TS2D_K9_ARGS));
    } else {
        TS2D_CUDA_TRY(ts2d_launch(k_preprocess3d_bwd<false, false>, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, 0, s, TS2D_K9_ARGS));
    }
#undef TS2D_K9_ARGS
    return (int)cudaGetLastError();
}

int ts2d_launch_export_geometry3d(int P, GeomState gs, float *v_view, float *normal_view, float *depth, float *rgb, uint8_t *clamped,
                                  uint32_t *tiles_touched, uint32_t *rect_min, uint32_t *rect_max, cudaStream_t s)
{
    k_export_geometry3d<<<(P + 255) / 256, 256, 0, s>>>(P, gs.rec0, gs.rec1, gs.dkey, gs.tiles, gs.rect, gs.clamp, v_view, normal_view, depth, rgb,
                                                        clamped, tiles_touched, rect_min, rect_max);
    return (int)cudaGetLastError();
}
