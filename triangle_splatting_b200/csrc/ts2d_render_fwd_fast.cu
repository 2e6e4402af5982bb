// ts2d_render_fwd_fast.cu -- fast front-to-back composite (K7, flags.exact == 0).
//
// Same contract as k_render_fwd (ts2d_render_fwd.cu) -- which stays as the op-for-op mirror of
// R2D/src/forward.cu:198-355 -- but organised for throughput on sm_100a:
//   * staging thread = one list entry: 3(+2) LDG.128 of the raster record, per-(tile, triangle) affine
//     barycentric setup, 8-bit sub-tile coverage mask (ts2d_fast.cuh);
//   * each warp walks only the entries whose mask bit is set for its 8x4 sub-tile (ballot + ffs);
//   * per pixel: 4 FFMA + min3 + 1 MUFU.EX2 to the alpha decision; decisions inside the rounding band
//     are re-taken with eval_exact(); the T <= 1e-4 cut is re-taken with an exact transmittance
//     re-walk done cooperatively by the warp (exact_T_upto) when T lands inside its error band;
//   * contrib_sum / contrib_max: one REDUX (fixed-point 2^-26) + one REDUX.MAX per (warp, triangle),
//     then a single lane issues the two REDs.
#include "ts2d_fast.cuh"

namespace {

// Exact transmittance of pixel (px, py) after visiting list positions [start, upto] (inclusive), computed with the
// reference's arithmetic and in the reference's order.  Called by the WHOLE warp for the pixel of lane `src`:
// lanes evaluate 32 consecutive entries in parallel, the product is then chained sequentially through shuffles so
// every rounding matches the reference's serial T *= (1 - alpha).
__device__ __noinline__ float exact_T_upto(const uint32_t *__restrict__ list, const float4 *__restrict__ rec0, uint32_t start, uint32_t upto,
                                           float px, float py, float two_gamma, int lane)
{
    float T = 1.0f;
    for (uint32_t k0 = start; k0 <= upto; k0 += 32) {
        const uint32_t k = k0 + lane;
        float f = 1.0f;
        if (k <= upto) {
            const uint32_t id = list[k];
            const float4 r0 = __ldg(rec0 + 3 * (size_t)id), r1 = __ldg(rec0 + 3 * (size_t)id + 1);
            PairEval e;
            if (eval_exact(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, two_gamma, px, py, e)) f = __fsub_rn(1.0f, e.alpha);
        }
#pragma unroll
        for (int l = 0; l < 32; l++) T = __fmul_rn(T, __shfl_sync(0xffffffffu, f, l));  // x 1.0f is exact
    }
    return T;
}

template <bool RICH>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_render_fwd_fast(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, const uint2 *__restrict__ ranges,
                  const uint32_t *__restrict__ list, const float4 *__restrict__ rec0, const float4 *__restrict__ rec1, float bg_depth,
                  const float *__restrict__ background, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
                  float *__restrict__ out_feature, float *__restrict__ out_depth, float *__restrict__ out_normal,
                  float *__restrict__ contrib_sum, float *__restrict__ contrib_max)
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
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    const int px = tile_x * TS2D_TILE + lx, py = tile_y * TS2D_TILE + ly;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float ox = (float)(tile_x * TS2D_TILE), oy = (float)(tile_y * TS2D_TILE);
    const size_t pix = (size_t)W * py + px;
    const GammaK gk = make_gamma(gamma);

    const uint2 range = ranges[tile];
    const uint32_t len = range.y - range.x;
    float T = 1.0f, terr = 0.0f;  // terr: bound on the relative error of T w.r.t. the reference's value
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f, accn0 = 0.f, accn1 = 0.f, accn2 = 0.f;
    uint32_t last = 0;
    bool done = !inside, saturated = false;

    for (uint32_t base = range.x; base < range.y; base += TS2D_BLOCK) {
        if (__syncthreads_and(done)) break;
        const int n = min((uint32_t)TS2D_BLOCK, range.y - base);
        if (tid < n) {
            const uint32_t id = list[base + tid];
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
            if (__all_sync(0xffffffffu, done)) break;
            const int idx = c * 32 + lane;
            const uint32_t mine = (idx < n) ? (uint32_t)s_mask[idx] : 0u;
            uint32_t bits = __ballot_sync(0xffffffffu, (mine >> warp) & 1u);
            while (bits) {
                const int j = c * 32 + (__ffs(bits) - 1);
                bits &= bits - 1;
                const float4 e1 = s_e1[j], e2 = s_e2[j];
                bool hit = false;
                float contrib = 0.0f;
                bool tband = false;
                FastPair f;
                if (!done) {
                    bool unc;
                    hit = eval_fast(e1, e2, pxf, pyf, gk, f, unc);
                    if (unc) {  // rare: take the reference's own decision and value for this pair
                        const uint32_t id = __float_as_uint(s_col[j].w);
                        const float area2 = __ldg(&rec0[3 * (size_t)id + 1].z);
                        PairEval e;
                        hit = eval_exact(e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, area2, e2.w, gk.two_gamma, pxf, pyf, e);
                        f.alpha = e.alpha;
                        f.power = e.power;
                        f.a1 = e.a1; f.a2 = e.a2; f.a3 = e.a3;
                    }
                    if (hit) {
                        contrib = f.alpha * T;
                        const float4 col = s_col[j];
                        acc0 = fmaf(contrib, col.x, acc0);
                        acc1 = fmaf(contrib, col.y, acc1);
                        acc2 = fmaf(contrib, col.z, acc2);
                        if (RICH) {
                            const float4 q0 = s_q0[j], q1 = s_q1[j];
                            accn0 = fmaf(contrib, q0.x, accn0);
                            accn1 = fmaf(contrib, q0.y, accn1);
                            accn2 = fmaf(contrib, q0.z, accn2);
                            const float d = fmaf(f.a3, q1.y, fmaf(q0.w, f.a1, q1.x * f.a2));
                            accd = fmaf(contrib, d, accd);
                        }
                        const float om = 1.0f - f.alpha;
                        T *= om;
                        // relative error of T grows by (abs error of alpha) / (1 - alpha)
                        terr = fmaf(f.alpha * (unc ? 0.0f : fmaf(gk.terr_c1, fabsf(f.power), gk.terr_c0)), rcp_approx(om) * 1.0001f, terr + 1.3e-7f);
                        const float dT = T - 0.0001f;
                        tband = fabsf(dT) <= 0.0001f * terr;
                        if (dT <= 0.0f) done = true;  // provisional; re-decided below when tband
                    }
                }
                uint32_t need = __ballot_sync(0xffffffffu, tband);
                while (need) {  // rare: exact transmittance re-walk for one pixel at a time, whole warp cooperating
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const float spx = __shfl_sync(0xffffffffu, pxf, src), spy = __shfl_sync(0xffffffffu, pyf, src);
                    const float Te = exact_T_upto(list, rec0, range.x, base + j, spx, spy, gk.two_gamma, lane);
                    if (lane == src) {
                        done = (Te <= 0.0001f);
                        T = Te;
                        terr = 0.0f;
                    }
                }
                if (done && !saturated && inside) {
                    saturated = true;
                    last = base - range.x + j + 1;
                }
                if (RICH) {
                    if (__ballot_sync(0xffffffffu, hit)) {
                        const uint32_t q = hit ? __float2uint_rn(contrib * 67108864.0f) : 0u;  // 2^26 fixed point, contrib < 1
                        const uint32_t ssum = __reduce_add_sync(0xffffffffu, q);
                        const uint32_t smax = __reduce_max_sync(0xffffffffu, __float_as_uint(contrib));
                        if (lane == 0) {
                            const uint32_t id = __float_as_uint(s_col[j].w);
                            atomicAdd(contrib_sum + id, (float)ssum * (1.0f / 67108864.0f));
                            atomicMax((unsigned int *)contrib_max + id, smax);
                        }
                    }
                }
            }
        }
    }

    if (inside) {
        if (!saturated) last = len;
        const float bg0 = background[0], bg1 = C > 1 ? background[1] : 0.f, bg2 = C > 2 ? background[2] : 0.f;
        const size_t HW = (size_t)H * W;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_feature[pix] = fmaf(T, bg0, acc0);
        if (C > 1) out_feature[HW + pix] = fmaf(T, bg1, acc1);
        if (C > 2) out_feature[2 * HW + pix] = fmaf(T, bg2, acc2);
        if (RICH) {
            out_depth[pix] = fmaf(T, bg_depth, accd);
            out_normal[pix] = accn0;
            out_normal[HW + pix] = accn1;
            out_normal[2 * HW + pix] = accn2;
        }
    }
}

}  // namespace

int ts2d_launch_render_fwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *list,
                                ImageState is, const ts2d_forward_out *out, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    if (owned <= 0) return 0;
    if (f->rich_info) {
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_sum, 0, sizeof(float) * (size_t)g->P, s));
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_max, 0, sizeof(float) * (size_t)g->P, s));
        k_render_fwd_fast<true><<<owned, TS2D_BLOCK, 0, s>>>(W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, is.ranges, list,
                                                             gs.rec0, gs.rec1, g->background_depth, g->background, is.final_T, is.n_contrib,
                                                             out->out_feature, out->depth, out->normal, out->contrib_sum, out->contrib_max);
    } else {
        k_render_fwd_fast<false><<<owned, TS2D_BLOCK, 0, s>>>(W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, is.ranges, list,
                                                              gs.rec0, gs.rec1, g->background_depth, g->background, is.final_T, is.n_contrib,
                                                              out->out_feature, nullptr, nullptr, nullptr, nullptr);
    }
    return (int)cudaGetLastError();
}
