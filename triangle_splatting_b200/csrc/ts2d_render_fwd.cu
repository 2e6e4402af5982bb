// ts2d_render_fwd.cu -- per-tile front-to-back alpha composite (K7).
//
// Replaces R2D/src/forward.cu:198-355 (FORWARD::renderCUDA).
//
// One CTA per 16x16 tile; the 8 warps own 8x4-pixel sub-tiles (lane = pixel).  The tile's slice of
// the sorted instance list is staged 256 entries at a time into shared memory as 48 B (+32 B rich)
// raster records (3 + 2 LDG.128 per entry instead of the reference's 9-15 scalar gathers).
// A warp stops as soon as its 32 pixels are saturated (the reference keeps all 256 threads
// iterating until the whole tile is done); the CTA stops when all warps are done.
// contrib_sum / contrib_max are warp-aggregated: one RED per (warp, contributing triangle) instead of
// one per (pixel, triangle) (forward.cu:323-324).
//
// Semantics kept exactly: integer pixel centres, skip tests
// ecc<0||ecc>10 and alpha<1/255, the terminating triangle IS blended, n_contrib = 1-based list
// position of the last entry visited before the pixel saturated (or the list length).
#include "ts2d_common.cuh"

template <bool RICH>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_render_fwd(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float two_gamma, const uint2 *__restrict__ ranges,
             const uint32_t *__restrict__ list, const float4 *__restrict__ rec0, const float4 *__restrict__ rec1, float bg_depth, const float *__restrict__ bg_ptr,
             const float *__restrict__ background, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib, float *__restrict__ out_feature,
             float *__restrict__ out_depth, float *__restrict__ out_normal, float *__restrict__ contrib_sum, float *__restrict__ contrib_max)
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

    const uint2 range = ranges[tile];
    float T = 1.0f;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f, accn0 = 0.f, accn1 = 0.f, accn2 = 0.f;
    uint32_t last = 0;
    bool done = !inside;

    for (uint32_t base = range.x; base < range.y; base += TS2D_BLOCK) {
        // barrier protects the staging buffers of the previous batch; the vote ends the tile early.
        if (__syncthreads_and(done)) break;
        const int n = min((uint32_t)TS2D_BLOCK, range.y - base);
        if (tid < n) {
            const uint32_t id = list[base + tid];
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
            if (__all_sync(0xffffffffu, done)) break;
            bool hit = false;
            float contrib = 0.0f;
            if (!done) {
                last = base - range.x + j + 1;
                const float4 r0 = s_rec0[3 * j], r1 = s_rec0[3 * j + 1], r2 = s_rec0[3 * j + 2];  // r1.z = 1/area2 (fast kernels), r2.w = area2
                PairEval e;
                if (eval_exact(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r2.w, r1.w, two_gamma, pxf, pyf, e)) {
                    hit = true;
                    contrib = __fmul_rn(e.alpha, T);
                    acc0 = __fmaf_rn(contrib, r2.x, acc0);
                    acc1 = __fmaf_rn(contrib, r2.y, acc1);
                    acc2 = __fmaf_rn(contrib, r2.z, acc2);
                    if (RICH) {
                        const float4 q0 = s_rec1[2 * j], q1 = s_rec1[2 * j + 1];
                        accn0 = __fmaf_rn(contrib, q0.x, accn0);
                        accn1 = __fmaf_rn(contrib, q0.y, accn1);
                        accn2 = __fmaf_rn(contrib, q0.z, accn2);
                        // d = vd.x*a1 + vd.y*a2 + vd.z*a3 with the reference build's contraction
                        const float d = __fmaf_rn(e.a3, q1.y, __fmaf_rn(q0.w, e.a1, __fmul_rn(q1.x, e.a2)));
                        accd = __fmaf_rn(contrib, d, accd);
                    }
                    T = __fmul_rn(T, __fsub_rn(1.0f, e.alpha));
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
        out_feature[pix] = __fmaf_rn(T, bg0, acc0);
        if (C > 1) out_feature[HW + pix] = __fmaf_rn(T, bg1, acc1);
        if (C > 2) out_feature[2 * HW + pix] = __fmaf_rn(T, bg2, acc2);
        if (RICH) {
            out_depth[pix] = __fmaf_rn(T, bg_depth, accd);
            out_normal[pix] = accn0;
            out_normal[HW + pix] = accn1;
            out_normal[2 * HW + pix] = accn2;
        }
    }
}

int ts2d_launch_render_fwd(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                           ImageState is, const ts2d_forward_out *out, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    if (owned <= 0) return 0;
    const float two_gamma = 2.0f * g->gamma;
    if (f->rich_info) {
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_sum, 0, sizeof(float) * (size_t)g->P, s));
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_max, 0, sizeof(float) * (size_t)g->P, s));
        k_render_fwd<true><<<owned, TS2D_BLOCK, 0, s>>>(W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, two_gamma, is.ranges, list,
                                                        gs.rec0, gs.rec1, g->background_depth, ts2d_bg_ptr(g, gs), g->background, is.final_T, is.n_contrib,
                                                        out->out_feature, out->depth, out->normal, out->contrib_sum, out->contrib_max);
    } else {
        k_render_fwd<false><<<owned, TS2D_BLOCK, 0, s>>>(W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, two_gamma, is.ranges, list,
                                                         gs.rec0, gs.rec1, g->background_depth, ts2d_bg_ptr(g, gs), g->background, is.final_T, is.n_contrib,
                                                         out->out_feature, nullptr, nullptr, nullptr, nullptr);
    }
    return (int)cudaGetLastError();
}
