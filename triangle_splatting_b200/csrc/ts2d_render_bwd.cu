// ts2d_render_bwd.cu -- per-tile reverse-walk gradient accumulation (K8).
//
// Replaces R2D/src/backward.cu:265-493 (BACKWARD::renderCUDA).
//
// Same tile/sub-tile decomposition and record staging as the forward composite, walking each
// tile's list back to front from the pixel's n_contrib.  The reference issues 10 (16 with
// rich_info) global float atomics per contributing (pixel, triangle) pair; here the 16 per-pair
// gradient components are first reduced across the 32 pixels of the warp with a recursive-halving
// butterfly (16 SHFL + 16 FADD for all 16 components instead of 80 + 80 for 16 independent
// butterflies), after which 16 lanes issue ONE coalesced 64 B RED burst into the triangle's
// accumulator line: 32x fewer atomics, all landing in a single L2 sector pair.
#include "ts2d_common.cuh"

// Sum v[0..15] over the warp; on return lane L holds component (L >> 1) & 15 in the return value.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane)
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

template <bool RICH>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_render_bwd(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, const uint2 *__restrict__ ranges,
             const uint32_t *__restrict__ list, const float4 *__restrict__ rec0, const float4 *__restrict__ rec1, float bg_depth, const float *__restrict__ bg_ptr,
             const float *__restrict__ background, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
             const float *__restrict__ dL_dout_feature, const float *__restrict__ dL_dout_depth, const float *__restrict__ dL_dout_normal,
             float *__restrict__ gacc)
{
    if (bg_ptr) bg_depth = __ldg(bg_ptr);  // model inputs: background depth computed on the device by K1
    __shared__ float4 s_rec0[TS2D_BLOCK * 3];
    __shared__ float4 s_rec1[RICH ? TS2D_BLOCK * 2 : 1];
    __shared__ uint32_t s_id[TS2D_BLOCK];

    const int tile = blockIdx.x * shard_world + shard_rank;
    if (tile >= n_tiles) return;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int px = tile_x * TS2D_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TS2D_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t pix = (size_t)W * py + px;
    const size_t HW = (size_t)H * W;
    const float two_gamma = 2.0f * gamma;

    const uint2 range = ranges[tile];
    const uint32_t len = range.y - range.x;
    float T = inside ? final_T[pix] : 0.0f;
    const uint32_t last = inside ? n_contrib[pix] : 0u;

    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accn0 = 0.f, accn1 = 0.f, accn2 = 0.f, accd = bg_depth;
    float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f, gn0 = 0.f, gn1 = 0.f, gn2 = 0.f, gd = 0.f;
    if (inside) {
        acc0 = background[0];
        gp0 = dL_dout_feature[pix];
        if (C > 1) { acc1 = background[1]; gp1 = dL_dout_feature[HW + pix]; }
        if (C > 2) { acc2 = background[2]; gp2 = dL_dout_feature[2 * HW + pix]; }
        if (RICH) {
            gn0 = dL_dout_normal[pix];
            gn1 = dL_dout_normal[HW + pix];
            gn2 = dL_dout_normal[2 * HW + pix];
            gd = dL_dout_depth[pix];
        }
    }
    // nothing in this warp's pixels was ever visited beyond max(last): skip whole batches behind it
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);

    for (uint32_t done_cnt = 0; done_cnt < len; done_cnt += TS2D_BLOCK) {
        __syncthreads();
        const int n = min((uint32_t)TS2D_BLOCK, len - done_cnt);
        if (tid < n) {
            const uint32_t id = list[range.y - 1 - done_cnt - tid];  // reversed order
            s_id[tid] = id;
            const float4 *r = rec0 + 3 * (size_t)id;
            s_rec0[3 * tid + 0] = __ldg(r);
            s_rec0[3 * tid + 1] = __ldg(r + 1);
            s_rec0[3 * tid + 2] = __ldg(r + 2);
            if (RICH) {
                const float4 *q = rec1 + 2 * (size_t)id;
                s_rec1[2 * tid + 0] = __ldg(q);
                s_rec1[2 * tid + 1] = __ldg(q + 1);
            }
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
                const float4 r0 = s_rec0[3 * j], r1 = s_rec0[3 * j + 1], r2 = s_rec0[3 * j + 2];
                PairEval e;
                if (eval_exact(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r2.w, r1.w, two_gamma, pxf, pyf, e)) {
                    hit = true;
                    const float area2 = r2.w, op = r1.w;
                    T = T / (1.0f - e.alpha);
                    const float contrib = e.alpha * T;
                    const float om = 1.0f - e.alpha;
                    float dL_dcontrib = 0.0f;
                    float da1 = 0.f, da2 = 0.f, da3 = 0.f;  // dL/d(a1,a2,a3)

                    v[8] = gp0 * contrib;
                    dL_dcontrib += gp0 * (r2.x - acc0);
                    acc0 = e.alpha * r2.x + om * acc0;
                    v[9] = gp1 * contrib;
                    dL_dcontrib += gp1 * (r2.y - acc1);
                    acc1 = e.alpha * r2.y + om * acc1;
                    v[10] = gp2 * contrib;
                    dL_dcontrib += gp2 * (r2.z - acc2);
                    acc2 = e.alpha * r2.z + om * acc2;

                    if (RICH) {
                        const float4 q0 = s_rec1[2 * j], q1 = s_rec1[2 * j + 1];
                        v[7] = gn0 * contrib;
                        v[11] = gn1 * contrib;
                        v[12] = gn2 * contrib;
                        dL_dcontrib += gn0 * (q0.x - accn0) + gn1 * (q0.y - accn1) + gn2 * (q0.z - accn2);
                        accn0 = e.alpha * q0.x + om * accn0;
                        accn1 = e.alpha * q0.y + om * accn1;
                        accn2 = e.alpha * q0.z + om * accn2;
                        const float dL_ddepth = gd * contrib;
                        v[13] = dL_ddepth * e.a1;
                        v[14] = dL_ddepth * e.a2;
                        v[15] = dL_ddepth * e.a3;
                        da1 = dL_ddepth * q0.w;
                        da2 = dL_ddepth * q1.x;
                        da3 = dL_ddepth * q1.y;
                        const float depth = q0.w * e.a1 + q1.x * e.a2 + q1.y * e.a3;
                        dL_dcontrib += gd * (depth - accd);
                        accd = e.alpha * depth + om * accd;
                    }

                    const float dL_dalpha = dL_dcontrib * T;
                    float dL_dpower = 0.0f;
                    if (op * e.G < 0.99f) dL_dpower = dL_dalpha * e.alpha;
                    const float dL_decc = dL_dpower * 2 * gamma * e.power / (e.ecc + TS2D_EPS);
                    if (e.a1 <= e.a2 && e.a1 <= e.a3) da1 += dL_decc * -3.0f;
                    else if (e.a2 <= e.a1 && e.a2 <= e.a3) da2 += dL_decc * -3.0f;
                    else da3 += dL_decc * -3.0f;

                    // barycentric Jacobians (backward.cu:464-479); perp(u) = (u.y, -u.x)
                    const f2 e12 = mk2(r0.z - r0.x, r0.w - r0.y);  // v2 - v1
                    const f2 e23 = mk2(r1.x - r0.z, r1.y - r0.w);  // v3 - v2
                    const f2 e31 = mk2(r0.x - r1.x, r0.y - r1.y);  // v1 - v3
                    const f2 pv1 = mk2(e.pv1x, e.pv1y), pv2 = mk2(e.pv2x, e.pv2y), pv3 = mk2(e.pv3x, e.pv3y);
                    const float inv = 1.0f / area2;
                    const f2 j11 = perp2(e23 * e.a1) * inv, j12 = perp2(e31 * e.a1 + pv3) * inv, j13 = perp2(e12 * e.a1 - pv2) * inv;
                    const f2 j21 = perp2(e23 * e.a2 - pv3) * inv, j22 = perp2(e31 * e.a2) * inv, j23 = perp2(e12 * e.a2 + pv1) * inv;
                    const f2 j31 = perp2(e23 * e.a3 + pv2) * inv, j32 = perp2(e31 * e.a3 - pv1) * inv, j33 = perp2(e12 * e.a3) * inv;
                    const f2 gv1 = da1 * j11 + da2 * j21 + da3 * j31;
                    const f2 gv2 = da1 * j12 + da2 * j22 + da3 * j32;
                    const f2 gv3 = da1 * j13 + da2 * j23 + da3 * j33;
                    v[0] = gv1.x; v[1] = gv1.y; v[2] = gv2.x; v[3] = gv2.y; v[4] = gv3.x; v[5] = gv3.y;
                    v[6] = dL_dalpha * e.G;  // unconditional (backward.cu:490)
                }
            }
            if (__ballot_sync(0xffffffffu, hit) == 0u) continue;
            const float r = warp_reduce16(v, lane);
            if ((lane & 1) == 0) atomicAdd(gacc + (size_t)s_id[j] * GACC_STRIDE + (lane >> 1), r);
        }
    }
}

int ts2d_launch_render_bwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                           ImageState is, const ts2d_loss_in *loss, float *gacc, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    TS2D_CUDA_TRY(cudaMemsetAsync(gacc, 0, sizeof(float) * GACC_STRIDE * (size_t)g->P, s));
    if (owned <= 0) return 0;
    if (f->rich_info) {
        k_render_bwd<true><<<owned, TS2D_BLOCK, 0, s>>>(W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, is.ranges, list, gs.rec0,
                                                        gs.rec1, g->background_depth, ts2d_bg_ptr(g, gs), g->background, is.final_T, is.n_contrib,
                                                        loss->dL_dout_feature, loss->dL_dout_depth, loss->dL_dout_normal, gacc);
    } else {
        k_render_bwd<false><<<owned, TS2D_BLOCK, 0, s>>>(W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, is.ranges, list, gs.rec0,
                                                         gs.rec1, g->background_depth, ts2d_bg_ptr(g, gs), g->background, is.final_T, is.n_contrib,
                                                         loss->dL_dout_feature, nullptr, nullptr, gacc);
    }
    return (int)cudaGetLastError();
}
