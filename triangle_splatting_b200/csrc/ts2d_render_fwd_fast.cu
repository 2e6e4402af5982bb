// ts2d_render_fwd_fast.cu -- fast front-to-back composite (K7, flags.exact == 0).
//
// Same contract as k_render_fwd (ts2d_render_fwd.cu) -- which stays as the op-for-op mirror of
// R2D/src/forward.cu:198-355 -- but organised for instruction throughput on sm_100a (the kernel is
// FP32-issue bound, not HBM bound: ncu shows ~85 % issue-slot utilisation at 2 % DRAM throughput):
//   * CTA = one 16x16 tile = 8 consumer warps (8x4-pixel sub-tiles, lane = pixel) + 1 producer warp that streams
//     the tile's list into a 3-slot shared-memory ring with TMA bulk copies and mbarriers (ts2d_pipe.cuh):
//     no CTA barrier in the steady state;
//   * a consumer warp first tests the slot's entries against its own sub-tile (lane = entry, conservative
//     3-edge test) and then walks only the covered ones (ballot + ffs);
//   * per pixel: reference-shaped barycentrics with one reciprocal multiply, min3, 1 MUFU.EX2;
//     decisions inside the rounding band are re-taken with eval_exact(); the T <= 1e-4 cut is re-taken
//     with an exact transmittance re-walk done cooperatively by the warp (exact_T_upto) when T lands
//     inside its running error bound;
//   * contrib_sum / contrib_max: per-warp 16-slot panel in shared memory, reduced with lanes re-mapped to
//     (triangle, pixel half): plain FADD/FMNMX, two REDs per (warp, triangle).
#include "ts2d_pipe.cuh"

namespace {

constexpr int FW_SLOTS = 16;    // triangles per contrib-statistics panel
constexpr int FW_NB = 128;      // list entries per ring slot
constexpr int FW_NS = 3;        // ring depth
constexpr int FW_THREADS = 288; // 8 consumer warps + 1 producer warp

template <bool RICH>
struct __align__(16) FwdSmem {
    float4 rec0[FW_NS][FW_NB][3];             // {v1 v2} {v3 1/area2 op} {rgb area2}
    float4 rec1[FW_NS][RICH ? FW_NB : 1][2];  // {n vd1} {vd2 vd3 - -}
    uint32_t id[FW_NS][FW_NB];
    float panel[RICH ? 8 : 1][RICH ? FW_SLOTS : 1][RICH ? 33 : 1];
    uint64_t full[FW_NS], empty[FW_NS];
    int n_finished;  // consumer warps whose 32 pixels are all saturated
    int stop_at;     // first batch the producer did NOT stage (set once every consumer is finished)
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
            const float area2 = __ldg(&rec0[3 * (size_t)id + 2].w);
            PairEval e;
            if (eval_exact(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, area2, r1.w, two_gamma, px, py, e)) f = __fsub_rn(1.0f, e.alpha);
        }
#pragma unroll
        for (int l = 0; l < 32; l++) T = __fmul_rn(T, __shfl_sync(0xffffffffu, f, l));  // x 1.0f is exact
    }
    return T;
}

template <bool RICH, bool GAMMA1>
__global__ void __launch_bounds__(FW_THREADS)
k_render_fwd_fast(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, const uint2 *__restrict__ ranges,
                  const uint32_t *__restrict__ list, const float4 *__restrict__ rec0, const float4 *__restrict__ rec1, float bg_depth,
                  const float *__restrict__ background, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib,
                  float *__restrict__ out_feature, float *__restrict__ out_depth, float *__restrict__ out_normal,
                  float *__restrict__ contrib_sum, float *__restrict__ contrib_max)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FwdSmem<RICH> &S = *reinterpret_cast<FwdSmem<RICH> *>(smem_raw);

    const int tile = blockIdx.x * shard_world + shard_rank;
    if (tile >= n_tiles) return;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint2 range = ranges[tile];
    const uint32_t len = range.y - range.x;
    const int nb = (int)((len + FW_NB - 1) / FW_NB);

    if (tid == 0) {
        for (int s = 0; s < FW_NS; s++) {
            mbar_init(&S.full[s], 32);  // every producer lane arrives once per batch (with its share of the tx bytes)
            mbar_init(&S.empty[s], 8);  // lane 0 of every consumer warp
        }
        S.n_finished = 0;
        S.stop_at = 0x7fffffff;
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 8) {
        // ------------------------------------------------------------------ producer warp
        constexpr uint32_t kBytes = 48u + (RICH ? 32u : 0u);
        for (int b = 0; b < nb; b++) {
            const int slot = b % FW_NS;
            if (b >= FW_NS) mbar_wait(&S.empty[slot], ((b / FW_NS) - 1) & 1);
            int nf = 0;
            if (lane == 0) nf = *(volatile int *)&S.n_finished;
            nf = __shfl_sync(0xffffffffu, nf, 0);  // one read, warp-uniform decision
            if (nf == 8) {                          // all pixels of the tile are saturated: stop streaming
                if (lane == 0) *(volatile int *)&S.stop_at = b;
                break;
            }
            const int n = min((uint32_t)FW_NB, len - (uint32_t)b * FW_NB);
            uint32_t ids[FW_NB / 32];
            uint32_t cnt = 0;
#pragma unroll
            for (int i = 0; i < FW_NB / 32; i++) {
                const int t = i * 32 + lane;
                if (t < n) {
                    ids[i] = list[range.x + (uint32_t)b * FW_NB + t];
                    S.id[slot][t] = ids[i];
                    cnt++;
                }
            }
            mbar_arrive_expect_tx(&S.full[slot], cnt * kBytes);  // release: the id stores above are visible to the waiters
#pragma unroll
            for (int i = 0; i < FW_NB / 32; i++) {
                const int t = i * 32 + lane;
                if (t < n) {
                    bulk_g2s(&S.rec0[slot][t][0], rec0 + 3 * (size_t)ids[i], 48, &S.full[slot]);
                    if (RICH) bulk_g2s(&S.rec1[slot][t][0], rec1 + 2 * (size_t)ids[i], 32, &S.full[slot]);
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    const int px = tile_x * TS2D_TILE + lx, py = tile_y * TS2D_TILE + ly;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float ox = (float)(tile_x * TS2D_TILE), oy = (float)(tile_y * TS2D_TILE);
    const float sub_x0 = (float)((warp & 1) * 8), sub_y0 = (float)((warp >> 1) * 4);
    GammaK gk = make_gamma(gamma);
    gk.is_one = GAMMA1;

    float T = 1.0f, Terr = 0.0f;  // Terr: bound on |T - (the reference's T)|
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f, accn0 = 0.f, accn1 = 0.f, accn2 = 0.f;
    uint32_t last = len;  // n_contrib if the pixel never saturates
    bool done = !inside;

    // contrib statistics panel (RICH): panel[warp][slot][pixel]; 33-word rows are conflict-free for both access patterns
    int pslot = 0;
    uint32_t my_id = 0;
    const int k = lane & 15, half = lane >> 4;
    auto flush_panel = [&](int filled) {
        __syncwarp();
        float s = 0.0f, m = 0.0f;
        if (k < filled) {
            const float *row = S.panel[warp][k] + half * 16;
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

    bool finished = false;  // warp-uniform: all 32 pixels saturated
    for (int b = 0; b < nb; b++) {
        const int slot = b % FW_NS;
        const uint32_t parity = (b / FW_NS) & 1;
        if (finished) {  // drain: keep the ring protocol going until the producer stops
            bool stopped = false;
            while (!mbar_try_wait(&S.full[slot], parity)) {
                if (*(volatile int *)&S.stop_at == b) { stopped = true; break; }
            }
            if (stopped) break;
        } else {
            mbar_wait(&S.full[slot], parity);
            const uint32_t base = (uint32_t)b * FW_NB;  // list position of the slot's first entry, relative to range.x
            const int n = min((uint32_t)FW_NB, len - base);
            // which entries of the slot can touch this warp's 8x4 sub-tile (lane = entry)
            uint32_t bits[FW_NB / 32];
#pragma unroll
            for (int g = 0; g < FW_NB / 32; g++) {
                const int t = g * 32 + lane;
                bool c = false;
                if (t < n) c = subtile_covers(S.rec0[slot][t][0], S.rec0[slot][t][1], ox, oy, sub_x0, sub_y0, gk);
                bits[g] = __ballot_sync(0xffffffffu, c);
            }
#pragma unroll
            for (int g = 0; g < FW_NB / 32; g++) {
                uint32_t rem = bits[g];
                if (__all_sync(0xffffffffu, done)) rem = 0;
                while (rem) {
                    const int j = g * 32 + (__ffs(rem) - 1);
                    rem &= rem - 1;
                    const float4 *E = S.rec0[slot][j];
                    const float4 e1 = E[0], e2 = E[1];
                    float contrib = 0.0f;  // > 0 <=> this lane blended the triangle
                    bool tband = false;
                    if (!done) {
                        FastPair f;
                        bool unc;
                        bool hit = eval_fast(e1, e2, pxf, pyf, gk, f, unc);
                        if (unc) {  // rare: the reference's own decision and alpha for this pair
                            PairEval e;
                            hit = eval_exact(e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, E[2].w, e2.w, gk.two_gamma, pxf, pyf, e);
                            f.alpha = e.alpha; f.power = e.power; f.a1 = e.a1; f.a2 = e.a2; f.a3 = e.a3;
                        }
                        if (hit) {
                            contrib = f.alpha * T;
                            const float4 col = E[2];
                            acc0 = fmaf(contrib, col.x, acc0);
                            acc1 = fmaf(contrib, col.y, acc1);
                            acc2 = fmaf(contrib, col.z, acc2);
                            if constexpr (RICH) {
                                const float4 q0 = S.rec1[slot][j][0], q1 = S.rec1[slot][j][1];
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
                                last = base + j + 1;
                            }
                        }
                    }
                    uint32_t need = __ballot_sync(0xffffffffu, tband);
                    while (need) {  // rare: exact transmittance re-walk for one pixel at a time, whole warp cooperating
                        const int src = __ffs(need) - 1;
                        need &= need - 1;
                        const float spx = __shfl_sync(0xffffffffu, pxf, src), spy = __shfl_sync(0xffffffffu, pyf, src);
                        const float Te = exact_T_upto(list, rec0, range.x, range.x + base + j, spx, spy, gk.two_gamma, lane);
                        if (lane == src) {
                            done = (Te <= 0.0001f);
                            last = done ? (base + j + 1) : len;
                            T = Te;
                            Terr = 0.0f;
                        }
                    }
                    if constexpr (RICH) {
                        // contrib_sum / contrib_max (forward.cu:323-324): park this pair-row in the warp's panel; every 16 rows
                        // the lanes switch roles (lane = triangle slot x pixel half) and reduce with plain FADD / FMNMX
                        if (__ballot_sync(0xffffffffu, contrib > 0.0f)) {
                            S.panel[warp][pslot][lane] = contrib;
                            if (k == pslot) my_id = S.id[slot][j];
                            if (++pslot == FW_SLOTS) {
                                flush_panel(FW_SLOTS);
                                pslot = 0;
                            }
                        }
                    }
                }
            }
            if (__all_sync(0xffffffffu, done)) {
                finished = true;
                if (lane == 0) atomicAdd(&S.n_finished, 1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty[slot]);
    }
    if constexpr (RICH) {
        if (pslot) flush_panel(pslot);
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
#define TS2D_FWD_LAUNCH(R, G, ...)                                                                                                \
    do {                                                                                                                          \
        const size_t smem = sizeof(FwdSmem<R>);                                                                                   \
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_render_fwd_fast<R, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        k_render_fwd_fast<R, G><<<owned, FW_THREADS, smem, s>>>(TS2D_FWD_ARGS, __VA_ARGS__);                                      \
    } while (0)
    if (f->rich_info) {
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_sum, 0, sizeof(float) * (size_t)g->P, s));
        TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_max, 0, sizeof(float) * (size_t)g->P, s));
        if (g1) TS2D_FWD_LAUNCH(true, true, out->depth, out->normal, out->contrib_sum, out->contrib_max);
        else TS2D_FWD_LAUNCH(true, false, out->depth, out->normal, out->contrib_sum, out->contrib_max);
    } else {
        if (g1) TS2D_FWD_LAUNCH(false, true, nullptr, nullptr, nullptr, nullptr);
        else TS2D_FWD_LAUNCH(false, false, nullptr, nullptr, nullptr, nullptr);
    }
#undef TS2D_FWD_LAUNCH
#undef TS2D_FWD_ARGS
    return (int)cudaGetLastError();
}
