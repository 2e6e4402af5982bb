// ts2d_render_bwd_fast.cu -- fast reverse-walk gradient accumulation (K8, flags.exact == 0).
//
// Same contract as k_render_bwd (ts2d_render_bwd.cu, the mirror of R2D/src/backward.cu:265-493) with the
// forward fast kernel's machinery (sub-tile masks, pixel-relative fast barycentrics, decision bands with
// eval_exact fallback -- so the set of contributing pairs is the one the forward pass blended) plus:
//
//   * Vertex gradients as MOMENTS.  The reference evaluates nine 2-vector Jacobians d a_i / d v_j per
//     (pixel, triangle) pair (backward.cu:464-479).  a_1, a_2 are affine in the pixel position, so
//     sum_p g_i(p) d a_i(p)/d v_j is a fixed linear map of the six moments
//         S_k = sum ga_k,  Q_k = sum ga_k * (p - v1),   ga_k = dL/da_k - dL/da_3,  k = 1, 2
//     taken relative to the triangle's own vertex v1 (well conditioned, tile independent, and p - v1 is
//     already in registers).  A pixel only forms 6 products; the 6 -> 6 map is applied once per triangle
//     in the preprocess-backward kernel (moments_to_vertex_grads, ts2d_preprocess.cu).
//   * 16 components reduced over the warp by a recursive-halving butterfly (16 SHFL + 16 FADD), then one
//     coalesced 64 B RED burst per (warp, triangle) into the triangle's accumulator line.
#include "ts2d_fast.cuh"

namespace {

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
    return r;  // lane L holds component (L >> 1) & 15
}

template <bool RICH>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_render_bwd_fast(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, const uint2 *__restrict__ ranges,
                  const uint32_t *__restrict__ list, const float4 *__restrict__ rec0, const float4 *__restrict__ rec1, float bg_depth,
                  const float *__restrict__ background, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
                  const float *__restrict__ dL_dout_feature, const float *__restrict__ dL_dout_depth,
                  const float *__restrict__ dL_dout_normal, float *__restrict__ gacc)
{
    __shared__ float4 s_e1[TS2D_BLOCK];   // v1.x, v1.y, v2.x, v2.y
    __shared__ float4 s_e2[TS2D_BLOCK];   // v3.x, v3.y, 1/area2, opacity
    __shared__ float4 s_col[TS2D_BLOCK];  // r, g, b, triangle id (bits)
    __shared__ float4 s_q0[RICH ? TS2D_BLOCK : 1];
    __shared__ float4 s_q1[RICH ? TS2D_BLOCK : 1];
    __shared__ uint8_t s_mask[TS2D_BLOCK];

    const int tile = blockIdx.x * shard_world + shard_rank;
    if (tile >= n_tiles) return;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int px = tile_x * TS2D_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TS2D_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float ox = (float)(tile_x * TS2D_TILE), oy = (float)(tile_y * TS2D_TILE);
    const size_t pix = (size_t)W * py + px;
    const size_t HW = (size_t)H * W;
    const GammaK gk = make_gamma(gamma);

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
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
    // list positions >= tile_last were visited by no pixel of the tile: skip their batches entirely
    __shared__ uint32_t s_tile_last;
    if (tid == 0) s_tile_last = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&s_tile_last, warp_last);
    __syncthreads();
    const uint32_t tile_last = s_tile_last;

    // batches are staged in REVERSE list order: staged slot t of the batch starting at `top` is list position top - t
    for (uint32_t done_cnt = len - tile_last; done_cnt < len; done_cnt += TS2D_BLOCK) {
        __syncthreads();
        const int n = min((uint32_t)TS2D_BLOCK, len - done_cnt);
        if (tid < n) {
            const uint32_t id = list[range.y - 1 - done_cnt - tid];
            const float4 *r = rec0 + 3 * (size_t)id;
            const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
            const float inv = 1.0f / r1.z;
            s_e1[tid] = r0;
            s_e2[tid] = make_float4(r1.x, r1.y, inv, r1.w);
            s_col[tid] = make_float4(r2.x, r2.y, r2.z, __uint_as_float(id));
            s_mask[tid] = (uint8_t)subtile_mask(r0, r1, inv, ox, oy, gk);
            if (RICH) {
                const float4 *q = rec1 + 2 * (size_t)id;
                s_q0[tid] = __ldg(q);
                s_q1[tid] = __ldg(q + 1);
            }
        }
        __syncthreads();

        for (int c = 0; c * 32 < n; c++) {
            const int idx = c * 32 + lane;
            const uint32_t mine = (idx < n) ? (uint32_t)s_mask[idx] : 0u;
            uint32_t bits = __ballot_sync(0xffffffffu, (mine >> warp) & 1u);
            while (bits) {
                const int j = c * 32 + (__ffs(bits) - 1);
                bits &= bits - 1;
                const uint32_t pos = len - 1 - done_cnt - j;  // 0-based list position
                if (pos >= warp_last) continue;                // warp-uniform
                const float4 e1 = s_e1[j], e2 = s_e2[j];
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; k++) v[k] = 0.0f;
                bool hit = false;
                if (pos < last) {
                    FastPair f;
                    bool unc;
                    hit = eval_fast(e1, e2, pxf, pyf, gk, f, unc);
                    // one more reference decision lives in the backward pass: op*G < 0.99 (clamp not active)
                    unc = unc || (fabsf(f.og - 0.99f) <= 0.99f * gk.band);
                    if (unc) {
                        const uint32_t id = __float_as_uint(s_col[j].w);
                        const float area2 = __ldg(&rec0[3 * (size_t)id + 1].z);
                        PairEval e;
                        hit = eval_exact(e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, area2, e2.w, gk.two_gamma, pxf, pyf, e);
                        f.a1 = e.a1; f.a2 = e.a2; f.a3 = e.a3; f.ecc = e.ecc;
                        if (hit) { f.power = e.power; f.G = e.G; f.og = __fmul_rn(e2.w, e.G); f.alpha = e.alpha; }
                    }
                    if (hit) {
                        const float4 col = s_col[j];
                        const float om = 1.0f - f.alpha;
                        T = T * rcp_approx(om);
                        const float contrib = f.alpha * T;
                        float dL_dcontrib;
                        v[8] = gp0 * contrib;
                        v[9] = gp1 * contrib;
                        v[10] = gp2 * contrib;
                        dL_dcontrib = gp0 * (col.x - acc0);
                        dL_dcontrib = fmaf(gp1, col.y - acc1, dL_dcontrib);
                        dL_dcontrib = fmaf(gp2, col.z - acc2, dL_dcontrib);
                        acc0 = fmaf(f.alpha, col.x, om * acc0);
                        acc1 = fmaf(f.alpha, col.y, om * acc1);
                        acc2 = fmaf(f.alpha, col.z, om * acc2);
                        float da1 = 0.f, da2 = 0.f, da3 = 0.f;
                        if (RICH) {
                            const float4 q0 = s_q0[j], q1 = s_q1[j];
                            v[7] = gn0 * contrib;
                            v[11] = gn1 * contrib;
                            v[12] = gn2 * contrib;
                            dL_dcontrib = fmaf(gn0, q0.x - accn0, dL_dcontrib);
                            dL_dcontrib = fmaf(gn1, q0.y - accn1, dL_dcontrib);
                            dL_dcontrib = fmaf(gn2, q0.z - accn2, dL_dcontrib);
                            accn0 = fmaf(f.alpha, q0.x, om * accn0);
                            accn1 = fmaf(f.alpha, q0.y, om * accn1);
                            accn2 = fmaf(f.alpha, q0.z, om * accn2);
                            const float dL_ddepth = gd * contrib;
                            v[13] = dL_ddepth * f.a1;
                            v[14] = dL_ddepth * f.a2;
                            v[15] = dL_ddepth * f.a3;
                            da1 = dL_ddepth * q0.w;
                            da2 = dL_ddepth * q1.x;
                            da3 = dL_ddepth * q1.y;
                            const float depth = fmaf(q1.y, f.a3, fmaf(q0.w, f.a1, q1.x * f.a2));
                            dL_dcontrib = fmaf(gd, depth - accd, dL_dcontrib);
                            accd = fmaf(f.alpha, depth, om * accd);
                        }
                        const float dL_dalpha = dL_dcontrib * T;
                        v[6] = dL_dalpha * f.G;  // unconditional (backward.cu:490)
                        const float dL_dpower = (f.og < 0.99f) ? dL_dalpha * f.alpha : 0.0f;
                        const float dL_decc3 = -3.0f * dL_dpower * gk.two_gamma * f.power * rcp_approx(f.ecc + TS2D_EPS);
                        // sub-gradient of min: first arg-min in the order a1, a2, a3 (backward.cu:449-461)
                        if (f.a1 <= f.a2 && f.a1 <= f.a3) da1 += dL_decc3;
                        else if (f.a2 <= f.a1 && f.a2 <= f.a3) da2 += dL_decc3;
                        else da3 += dL_decc3;
                        const float ga1 = da1 - da3, ga2 = da2 - da3;
                        // moments about v1: q = p - v1 = -pv1
                        v[0] = ga1;
                        v[1] = -ga1 * f.pv1x;
                        v[2] = -ga1 * f.pv1y;
                        v[3] = ga2;
                        v[4] = -ga2 * f.pv1x;
                        v[5] = -ga2 * f.pv1y;
                    }
                }
                if (__ballot_sync(0xffffffffu, hit) == 0u) continue;
                const float r = warp_reduce16(v, lane);
                if ((lane & 1) == 0) atomicAdd(gacc + (size_t)__float_as_uint(s_col[j].w) * GACC_STRIDE + (lane >> 1), r);
            }
        }
    }
}

}  // namespace

int ts2d_launch_render_bwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                                ImageState is, const ts2d_loss_in *loss, float *gacc, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    TS2D_CUDA_TRY(cudaMemsetAsync(gacc, 0, sizeof(float) * GACC_STRIDE * (size_t)g->P, s));
    if (owned <= 0) return 0;
    if (f->rich_info) {
        k_render_bwd_fast<true><<<owned, TS2D_BLOCK, 0, s>>>(W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, is.ranges, list,
                                                             gs.rec0, gs.rec1, g->background_depth, g->background, is.final_T, is.n_contrib,
                                                             loss->dL_dout_feature, loss->dL_dout_depth, loss->dL_dout_normal, gacc);
    } else {
        k_render_bwd_fast<false><<<owned, TS2D_BLOCK, 0, s>>>(W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, is.ranges, list,
                                                              gs.rec0, gs.rec1, g->background_depth, g->background, is.final_T, is.n_contrib,
                                                              loss->dL_dout_feature, nullptr, nullptr, gacc);
    }
    return (int)cudaGetLastError();
}
