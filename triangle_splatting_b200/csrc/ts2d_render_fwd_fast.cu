// ts2d_render_fwd_fast.cu -- fast front-to-back composite (K7, flags.exact == 0).
//
// Same contract as k_render_fwd (ts2d_render_fwd.cu) -- which stays as the op-for-op mirror of
// R2D/src/forward.cu:198-355 -- but organised for instruction throughput on sm_100a (the kernel is
// FP32-issue bound, not HBM bound: ncu shows ~80 % issue-slot utilisation at 2 % DRAM throughput):
//   * staging thread = one list entry: 3(+2) LDG.128 of the raster record, reciprocal of area2,
//     8-bit sub-tile coverage mask (ts2d_fast.cuh), one 80-byte shared-memory entry;
//   * each warp walks only the entries whose mask bit is set for its 8x4 sub-tile (ballot + ffs);
//   * per pixel: reference-shaped barycentrics with one reciprocal multiply, min3, 1 MUFU.EX2;
//     decisions inside the rounding band are re-taken with eval_exact(); the T <= 1e-4 cut is re-taken
//     with an exact transmittance re-walk done cooperatively by the warp (exact_T_upto) when T lands
//     inside its running error bound;
//   * contrib_sum / contrib_max: one REDUX.SUM (fixed-point 2^-26) + one REDUX.MAX per (warp, triangle),
//     then a single lane issues the two REDs.
#include "ts2d_fast.cuh"

namespace {

constexpr int FW_SLOTS = 16;  // triangles per contrib-statistics panel

template <bool RICH>
struct __align__(16) FwdEntry {
    float4 e1;   // v1.x, v1.y, v2.x, v2.y
    float4 e2;   // v3.x, v3.y, 1/area2, opacity
    float4 col;  // r, g, b, triangle id (bits)
    float4 q0;   // n.x, n.y, n.z, vd1   (rich)
    float4 q1;   // vd2, vd3, -, -       (rich)
};
template <>
struct __align__(16) FwdEntry<false> {
    float4 e1, e2, col;
};

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
            if (eval_exact(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, __ldg(&rec0[3 * (size_t)id + 2].w), r1.w, two_gamma, px, py, e)) f = __fsub_rn(1.0f, e.alpha);
        }
#pragma unroll
        for (int l = 0; l < 32; l++) T = __fmul_rn(T, __shfl_sync(0xffffffffu, f, l));  // x 1.0f is exact
    }
    return T;
}

// Out-of-line slow path: the reference's own decision and alpha for one pair.
__device__ __forceinline__ bool exact_pair(const float4 e1, const float4 e2, const float4 *__restrict__ rec0, uint32_t id, float two_gamma,
                                        float px, float py, float &alpha, float &power, float &a1, float &a2, float &a3)
{
    const float area2 = __ldg(&rec0[3 * (size_t)id + 2].w);
    PairEval e;
    const bool hit = eval_exact(e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, area2, e2.w, two_gamma, px, py, e);
    alpha = e.alpha;
    power = e.power;
    a1 = e.a1;
    a2 = e.a2;
    a3 = e.a3;
    return hit;
}

template <bool RICH, bool GAMMA1>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_render_fwd_fast(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, const uint2 *__restrict__ ranges,
                  const uint32_t *__restrict__ list, const float4 *__restrict__ rec0, const float4 *__restrict__ rec1, float bg_depth,
                  const float *__restrict__ background, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
                  float *__restrict__ out_feature, float *__restrict__ out_depth, float *__restrict__ out_normal,
                  float *__restrict__ contrib_sum, float *__restrict__ contrib_max)
{
    __shared__ FwdEntry<RICH> s_ent[TS2D_BLOCK];
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
    GammaK gk = make_gamma(gamma);
    gk.is_one = GAMMA1;

    const uint2 range = ranges[tile];
    float T = 1.0f, Terr = 0.0f;  // Terr: bound on |T - (the reference's T)|
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f, accn0 = 0.f, accn1 = 0.f, accn2 = 0.f;
    uint32_t last = range.y - range.x;  // n_contrib if the pixel never saturates
    bool done = !inside;

    // contrib statistics panel (RICH): s_panel[warp][slot][pixel]; 33-word rows are conflict-free for both access patterns
    __shared__ float s_panel[RICH ? 8 : 1][RICH ? FW_SLOTS : 1][RICH ? 33 : 1];
    int slot = 0;
    uint32_t my_id = 0;
    const int k = lane & 15, half = lane >> 4;
    auto flush_panel = [&](int filled) {
        __syncwarp();
        float s = 0.0f, m = 0.0f;
        if (k < filled) {
            const float *row = s_panel[warp][k] + half * 16;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float v = row[i];
                s += v;
                m = fmaxf(m, v);
            }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        if (k < filled && half == 0) {
            atomicAdd(contrib_sum + my_id, s);
            atomicMax((unsigned int *)contrib_max + my_id, __float_as_uint(m));  // contrib >= 0: bit order == value order
        }
        __syncwarp();
    };

    for (uint32_t base = range.x; base < range.y; base += TS2D_BLOCK) {
        if (__syncthreads_and(done)) break;
        const int n = min((uint32_t)TS2D_BLOCK, range.y - base);
        if (tid < n) {
            const uint32_t id = list[base + tid];
            const float4 *r = rec0 + 3 * (size_t)id;
            const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
            const float inv = r1.z;  // the record carries 1/area2 (K1 computes it once per triangle)
            FwdEntry<RICH> &E = s_ent[tid];
            E.e1 = r0;
            E.e2 = r1;
            E.col = make_float4(r2.x, r2.y, r2.z, __uint_as_float(id));
            if constexpr (RICH) {
                const float4 *q = rec1 + 2 * (size_t)id;
                E.q0 = __ldg(q);
                E.q1 = __ldg(q + 1);
            }
            s_mask[tid] = (uint8_t)subtile_mask(r0, r1, inv, ox, oy, gk);
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
                const FwdEntry<RICH> &E = s_ent[j];
                const float4 e1 = E.e1, e2 = E.e2;
                float contrib = 0.0f;  // > 0 <=> this lane blended the triangle
                bool tband = false;
                if (!done) {
                    FastPair f;
                    bool unc;
                    bool hit = eval_fast(e1, e2, pxf, pyf, gk, f, unc);
                    if (unc) hit = exact_pair(e1, e2, rec0, __float_as_uint(E.col.w), gk.two_gamma, pxf, pyf, f.alpha, f.power, f.a1, f.a2, f.a3);
                    if (hit) {
                        contrib = f.alpha * T;
                        const float4 col = E.col;
                        acc0 = fmaf(contrib, col.x, acc0);
                        acc1 = fmaf(contrib, col.y, acc1);
                        acc2 = fmaf(contrib, col.z, acc2);
                        if constexpr (RICH) {
                            const float4 q0 = E.q0, q1 = E.q1;
                            accn0 = fmaf(contrib, q0.x, accn0);
                            accn1 = fmaf(contrib, q0.y, accn1);
                            accn2 = fmaf(contrib, q0.z, accn2);
                            accd = fmaf(contrib, fmaf(f.a3, q1.y, fmaf(q0.w, f.a1, q1.x * f.a2)), accd);
                        }
                        const float om = 1.0f - f.alpha;
                        // |T_new - T_ref_new| <= |T - T_ref| * om + T * |d alpha| + rounding of the two product chains
                        Terr = fmaf(Terr, om, contrib * (unc ? 0.0f : fmaf(gk.terr_c1, fabsf(f.power), gk.terr_c0)));
                        T *= om;
                        Terr = fmaf(T, 1.3e-7f, Terr);
                        const float dT = T - 0.0001f;
                        tband = fabsf(dT) <= Terr;
                        if (dT <= 0.0f) {  // provisional when tband: re-decided below on the exact transmittance
                            done = true;
                            last = base - range.x + j + 1;
                        }
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
                        last = done ? (base - range.x + j + 1) : (range.y - range.x);
                        T = Te;
                        Terr = 0.0f;
                    }
                }
                if constexpr (RICH) {
                    // contrib_sum / contrib_max (forward.cu:323-324): park this pair-row in the warp's panel; every 16 rows the
                    // lanes switch roles (lane = triangle slot x pixel half) and reduce with plain FADD / FMNMX
                    if (__ballot_sync(0xffffffffu, contrib > 0.0f)) {
                        s_panel[warp][slot][lane] = contrib;
                        if (k == slot) my_id = __float_as_uint(E.col.w);
                        if (++slot == FW_SLOTS) {
                            flush_panel(FW_SLOTS);
                            slot = 0;
                        }
                    }
                }
            }
        }
    }
    if constexpr (RICH) {
        if (slot) flush_panel(slot);
    }

    if (inside) {
        const size_t pix = (size_t)W * py + px;
        const float bg0 = background[0], bg1 = C > 1 ? background[1] : 0.f, bg2 = C > 2 ? background[2] : 0.f;
        const size_t HW = (size_t)H * W;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_feature[pix] = fmaf(T, bg0, acc0);
        if (C > 1) out_feature[HW + pix] = fmaf(T, bg1, acc1);
        if (C > 2) out_feature[2 * HW + pix] = fmaf(T, bg2, acc2);
        if constexpr (RICH) {
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
    const bool g1 = g->gamma == 1.0f;
#define TS2D_FWD_ARGS                                                                                                                      \
    W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, is.ranges, list, gs.rec0, gs.rec1, g->background_depth, g->background, \
        is.final_T, is.n_contrib, out->out_feature
    if (f->rich_info) {
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_sum, 0, sizeof(float) * (size_t)g->P, s));
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_max, 0, sizeof(float) * (size_t)g->P, s));
        if (g1)
            k_render_fwd_fast<true, true><<<owned, TS2D_BLOCK, 0, s>>>(TS2D_FWD_ARGS, out->depth, out->normal, out->contrib_sum, out->contrib_max);
        else
            k_render_fwd_fast<true, false><<<owned, TS2D_BLOCK, 0, s>>>(TS2D_FWD_ARGS, out->depth, out->normal, out->contrib_sum, out->contrib_max);
    } else {
        if (g1)
            k_render_fwd_fast<false, true><<<owned, TS2D_BLOCK, 0, s>>>(TS2D_FWD_ARGS, nullptr, nullptr, nullptr, nullptr);
        else
            k_render_fwd_fast<false, false><<<owned, TS2D_BLOCK, 0, s>>>(TS2D_FWD_ARGS, nullptr, nullptr, nullptr, nullptr);
    }
#undef TS2D_FWD_ARGS
    return (int)cudaGetLastError();
}
