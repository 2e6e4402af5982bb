// ts2d_render_fwd_fast.cu -- fast front-to-back composite (K7, flags.exact == 0).
//
// Same contract as k_render_fwd (ts2d_render_fwd.cu) -- which stays as the op-for-op mirror of
// R2D/src/forward.cu:198-355 -- but organised for instruction throughput on sm_100a (the kernel is
// FP32-issue bound, not HBM bound: ncu shows ~80 % issue-slot utilisation at 2 % DRAM throughput).
//
// Warp-autonomous walk.  Each warp owns one 8x4-pixel sub-tile of a 16x16 tile and never synchronises with the
// other warps of its CTA (the cooperative-staging version of this kernel lost 20 % of its warp time at the
// per-batch __syncthreads: the number of list entries that touch a sub-tile differs from warp to warp):
//   gather   the warp scans the tile's instance keys 32 at a time (one coalesced load, prefetched one chunk ahead),
//            keeps the entries whose sub-tile bit is set (ts2d_binning.cu computes the 8-bit coverage mask once per
//            instance) and compacts their list positions until up to 32 are collected;
//   stage    lane i loads the raster record of collected entry i (3(+2) LDG.128) into the warp's private
//            shared-memory buffer -- only records that the sub-tile really needs are read;
//   walk     a plain counted loop over the staged entries: reference-shaped barycentrics with one reciprocal
//            multiply, min3, 1 MUFU.EX2 per pixel; decisions inside the rounding band are re-taken with
//            eval_exact(); the T <= 1e-4 cut is re-taken with an exact transmittance re-walk done cooperatively
//            by the warp (exact_T_upto) when T lands inside its running error bound.
//   contrib_sum / contrib_max: each visited entry leaves its 32 per-pixel contrib values in one row of a 32-row panel
//            (a single STS in the walk); after the walk the lanes switch roles (lane = entry), reduce their row with
//            plain FADD / FMNMX and issue one RED pair per triangle that was blended anywhere in the sub-tile.
#include "ts2d_fast.cuh"

namespace {

constexpr int FW_PROW = 33;    // panel row stride in words: conflict-free for both access patterns

// Per-warp shared-memory block (byte offsets from the warp's base address):
//   entry j at j * EB:  {v1.x v1.y v2.x v2.y} {v3.x v3.y 1/area2 op} {r g b id} [RICH: {n.x n.y n.z vd1} {vd2 vd3 pos -}]
//   non-RICH: list positions in a separate u32[32] after the entries;
//   RICH: contrib panel [32 entries][33] floats after that (row j = the 32 pixels' contrib of staged entry j).
template <bool RICH>
struct FwdLayout {
    static constexpr int EB = RICH ? 80 : 48;
    static constexpr int POS = RICH ? 72 : 32 * 48;      // position of entry j: POS + j * POS_STRIDE
    static constexpr int POS_STRIDE = RICH ? 80 : 4;
    static constexpr int PANEL = RICH ? 32 * 80 : 0;
    static constexpr int BYTES = RICH ? 32 * 80 + 32 * FW_PROW * 4 : 32 * 48 + 128;
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

// The reference's own decision and alpha for one pair.
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

#ifndef TS2D_FWD_MINB
#define TS2D_FWD_MINB 4  // resident CTAs per SM the register allocation targets (4 -> 64 registers; 5 -> 48 spills the accumulators)
#endif

// CW = warps per CTA: a CTA owns CW of the tile's eight 8x4 sub-tiles (blockIdx = owned-tile index * (8 / CW) + part).  The warps
// never synchronise with each other, so the CTA is only a unit of residency: a CTA stays on its SM until its slowest warp is done,
// and with 8 warps of unequal work ~15 % of the warp slots idled (ncu: 42 % active warps of a 50 % ceiling).
template <bool RICH, bool GAMMA1, int CW>
__global__ void __launch_bounds__(32 * CW, TS2D_FWD_MINB * 8 / CW)
k_render_fwd_fast(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, const uint2 *__restrict__ ranges,
                  const uint32_t *keys, const uint32_t *__restrict__ list, const float4 *__restrict__ rec0,
                  const float4 *__restrict__ rec1, float bg_depth, const float *__restrict__ bg_ptr, const float *__restrict__ background, float *__restrict__ final_T,
                  uint32_t *__restrict__ n_contrib, float *__restrict__ out_feature, float *__restrict__ out_depth,
                  float *__restrict__ out_normal, float *__restrict__ contrib_sum, float *__restrict__ contrib_max,
                  uint32_t *__restrict__ lastw, unsigned long long *__restrict__ bwd_rows, unsigned long long *__restrict__ csum64)
{
    ts2d_grid_chain();
    if (bg_ptr) bg_depth = __ldg(bg_ptr);  // model inputs: background depth computed on the device by K1
    using L = FwdLayout<RICH>;
    extern __shared__ __align__(16) unsigned char s_raw[];

    constexpr int PARTS = 8 / CW;
    const int tile = (blockIdx.x / PARTS) * shard_world + shard_rank;
    if (tile >= n_tiles) return;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int tid = threadIdx.x, lwarp = tid >> 5, warp = (blockIdx.x % PARTS) * CW + lwarp, lane = tid & 31;
    const int px = tile_x * TS2D_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TS2D_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    float pxf = (float)px, pyf = (float)py;
    asm volatile("" : "+f"(pxf), "+f"(pyf));  // opaque: keep them in registers (ptxas re-materialised two I2FP per walk iteration)
    GammaK gk = make_gamma(GAMMA1 ? 1.0f : gamma);  // gamma == 1: every constant of the error model folds to an immediate
    gk.is_one = GAMMA1;
    const uint32_t sb = smem_base(s_raw + lwarp * L::BYTES);  // this warp's block
    uint32_t *keys_rw = const_cast<uint32_t *>(keys);
    const uint32_t lt_mask = (1u << lane) - 1u;

    const uint2 range = ranges[tile];
    float T = 1.0f, Terr = 0.0f;  // Terr: bound on |T - (the reference's T)|
    v2 acc01 = bc(0.f), accn01 = bc(0.f);  // {channel 0, channel 1} and {n.x, n.y}: one FFMA2 each per blend
    float acc2 = 0.f, accd = 0.f, accn2 = 0.f;
    uint32_t last = range.y - range.x;  // n_contrib if the pixel never saturates
    bool done = !inside;
    uint32_t rows = 0;  // (sub-tile, entry) pairs the backward pass will visit: one 64 B row each (ts2d_bwd_reduce.cu)

    // contrib_sum / contrib_max (forward.cu:323-324), RICH: every visited entry j leaves its 32 per-pixel contrib values in panel
    // row j (zeros where the pixel skipped it); after the walk lane j reduces row j with plain FADD / FMNMX and issues the
    // triangle's RED pair -- one flush per round of up to 32 entries, nothing but a single STS inside the walk.
    auto flush_panel = [&](int visited) {
        __syncwarp();
        bool kept = false;
        if (lane < visited) {
            const uint32_t row = sb + L::PANEL + lane * (FW_PROW * 4);
            float s = 0.0f, m = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const float v = lds32f(row + 4 * i);
                s += v;
                m = fmaxf(m, v);
            }
            if (m > 0.0f) {
                kept = true;
                const uint32_t id = lds32(sb + lane * L::EB + 44);
                atomicAdd(csum64 + id, __float2ull_rn(s * 4294967296.0f));  // 2^-32 fixed point: integer sums do not depend on the order of the additions
                atomicMax((unsigned int *)contrib_max + id, __float_as_uint(m));  // contrib >= 0: bit order == value order
            } else {
                // No pixel of this sub-tile blended the entry (footprint between pixel centres, or every pixel under it already
                // saturated): clear the sub-tile's coverage bit in the instance key, so the backward pass -- whose per-pair decisions
                // are the same arithmetic -- does not even stage it.  Other warps only ever look at their own bit of the word.
                atomicAnd(keys_rw + lds32(sb + L::POS + lane * L::POS_STRIDE), ~(1u << warp));
            }
        }
        rows += __popc(__ballot_sync(0xffffffffu, kept));
    };

    uint32_t cur = range.x;  // next list position to scan
    uint32_t kreg = (cur + lane < range.y) ? __ldg(keys + cur + lane) : 0u;  // keys[cur + lane], always one chunk ahead
    while (true) {
        if (__all_sync(0xffffffffu, done)) break;
        // ---- gather: list positions of up to 32 entries that cover this sub-tile
        int count = 0;
        while (cur < range.y) {
            const bool cov = (kreg >> warp) & 1u;
            const uint32_t b = __ballot_sync(0xffffffffu, cov);
            const int n = __popc(b);
            if (count + n > 32) break;  // this chunk opens the next round (kreg still holds it)
            if (cov) sts32(sb + L::POS + (count + __popc(b & lt_mask)) * L::POS_STRIDE, cur + lane);
            count += n;
            cur += 32;
            kreg = (cur + lane < range.y) ? __ldg(keys + cur + lane) : 0u;
        }
        if (count == 0) break;  // list exhausted
        __syncwarp();
        // ---- stage: lane i fetches entry i
        if (lane < count) {
            const uint32_t ea = sb + lane * L::EB;
            const uint32_t id = list[lds32(sb + L::POS + lane * L::POS_STRIDE)];
            const float4 *r = rec0 + 3 * (size_t)id;
            const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
            sts128(ea, r0);
            sts128(ea + 16, r1);
            sts128(ea + 32, make_float4(r2.x, r2.y, r2.z, __uint_as_float(id)));
            if constexpr (RICH) {
                const float4 *q = rec1 + 2 * (size_t)id;
                const float4 q0 = __ldg(q), q1 = __ldg(q + 1);
                sts128(ea + 48, q0);
                sts32f(ea + 64, q1.x);  // +72 holds the list position
                sts32f(ea + 68, q1.y);
            }
        }
        __syncwarp();
        // ---- walk
        // software-pipelined by one entry: the vertex words of entry j + 1 are requested before the panel store of entry j (a shared
        // load cannot be moved above a store that may alias it, so at the top of the iteration it would sit behind that store with
        // its whole latency in front of the pair evaluation; one entry past the staged ones is read and never used)
        uint32_t ea = sb, prow = sb + L::PANEL + lane * 4;  // this lane's word of the entry's panel row
        int visited = count;
        float4 e1 = lds128(ea), e2 = lds128(ea + 16);
        for (int j = 0; j < count; j++, ea += L::EB, prow += FW_PROW * 4) {
            float contrib = 0.0f;  // > 0 <=> this lane blended the triangle
            bool evt = false;      // this lane's transmittance reached the 1e-4 cut or its error band: handled out of line below
            if (!done) {
                FastPair f;
                bool unc;
                bool hit = eval_fast(e1, e2, pxf, pyf, gk, f, unc);
                if (unc) hit = exact_pair(e1, e2, rec0, lds32(ea + 44), gk.two_gamma, pxf, pyf, f.alpha, f.power, f.a1, f.a2, f.a3);
                if (hit) {
                    // the entry's colour / normal / depth words are requested first and consumed last: the shared-memory helpers are
                    // volatile asm, so the compiler leaves the loads where they are written and the transmittance update runs under them
                    const float4 col = lds128(ea + 32);
                    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f);
                    float2 q1 = make_float2(0.f, 0.f);
                    if constexpr (RICH) {
                        q0 = lds128(ea + 48);
                        q1 = lds64(ea + 64);
                    }
                    contrib = f.alpha * T;
                    const float om = 1.0f - f.alpha;
                    // |T_new - T_ref_new| <= |T - T_ref| * om + T * |d alpha| + rounding of the two product chains
                    // (terr_c1 |power| written as (terr_c1 / 2) pw: the same product, so the same rounded sum, without forming power)
                    Terr = fmaf(Terr, om, contrib * (unc ? 0.0f : fmaf(0.5f * gk.terr_c1, f.pw, gk.terr_c0)));
                    T *= om;
                    Terr = fmaf(T, 1.3e-7f, Terr);
                    evt = (T - 0.0001f) <= Terr;  // saturated (T <= 1e-4) or within the error band of the cut
                    acc01 = fma2(bc(contrib), mk2v(col.x, col.y), acc01);
                    acc2 = fmaf(contrib, col.z, acc2);
                    if constexpr (RICH) {
                        accn01 = fma2(bc(contrib), mk2v(q0.x, q0.y), accn01);
                        accn2 = fmaf(contrib, q0.z, accn2);
                        accd = fmaf(contrib, fmaf(f.a3, q1.y, fmaf(q0.w, f.a1, q1.x * f.a2)), accd);
                    }
                }
            }
            const float4 n1 = lds128(ea + L::EB), n2 = lds128(ea + L::EB + 16);
            if constexpr (RICH) sts32f(prow, contrib);
            e1 = n1;
            e2 = n2;
            if (__any_sync(0xffffffffu, evt)) {  // at most a few times per pixel: everything about saturation lives here
                const uint32_t pos = lds32(sb + L::POS + j * L::POS_STRIDE);
                const float dT = T - 0.0001f;
                if (evt && dT <= 0.0f) {  // provisional when inside the band: re-decided below on the exact transmittance
                    done = true;
                    last = pos - range.x + 1;
                }
                uint32_t need = __ballot_sync(0xffffffffu, evt && fabsf(dT) <= Terr);
                while (need) {  // rare: exact transmittance re-walk for one pixel at a time, whole warp cooperating
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const float spx = __shfl_sync(0xffffffffu, pxf, src), spy = __shfl_sync(0xffffffffu, pyf, src);
                    const float Te = exact_T_upto(list, rec0, range.x, pos, spx, spy, gk.two_gamma, lane);
                    if (lane == src) {
                        done = (Te <= 0.0001f);
                        last = done ? (pos - range.x + 1) : (range.y - range.x);
                        T = Te;
                        Terr = 0.0f;
                    }
                }
                if (__all_sync(0xffffffffu, done)) {  // every pixel of the sub-tile has saturated
                    visited = j + 1;
                    break;
                }
            }
        }
        if constexpr (RICH) flush_panel(visited);
        else rows += (uint32_t)visited;  // no per-entry blend information without the panel: every visited entry gets a row
        __syncwarp();  // the next gather overwrites positions / entries / panel rows
    }
    {   // where the backward walk of this sub-tile starts, and how many rows it will write
        const uint32_t wl = __reduce_max_sync(0xffffffffu, inside ? last : 0u);
        if (lane == 0) {
            lastw[8 * (size_t)tile + warp] = wl;
            if (rows) atomicAdd(bwd_rows, (unsigned long long)rows);
        }
    }

    if (inside) {
        const size_t pix = (size_t)W * py + px;
        const float bg0 = background[0], bg1 = C > 1 ? background[1] : 0.f, bg2 = C > 2 ? background[2] : 0.f;
        const size_t HW = (size_t)H * W;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_feature[pix] = fmaf(T, bg0, acc01.a);
        if (C > 1) out_feature[HW + pix] = fmaf(T, bg1, acc01.b);
        if (C > 2) out_feature[2 * HW + pix] = fmaf(T, bg2, acc2);
        if constexpr (RICH) {
            out_depth[pix] = fmaf(T, bg_depth, accd);
            out_normal[pix] = accn01.a;
            out_normal[HW + pix] = accn01.b;
            out_normal[2 * HW + pix] = accn2;
        }
    }
}

}  // namespace

int ts2d_launch_render_fwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *keys,
                                const uint32_t *list, ImageState is, const ts2d_forward_out *out, bool pre_cleared, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    if (owned <= 0) return 0;
    const bool g1 = g->gamma == 1.0f;
#define TS2D_FWD_ARGS                                                                                                                       \
    W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, is.ranges, keys, list, gs.rec0, gs.rec1, g->background_depth, ts2d_bg_ptr(g, gs),         \
        g->background, is.final_T, is.n_contrib, o_feature
    float *o_feature = out->out_feature, *o_depth = out->depth, *o_normal = out->normal, *o_csum = out->contrib_sum, *o_cmax = out->contrib_max;
    unsigned long long *rows_ctr = reinterpret_cast<unsigned long long *>(&gs.hdr->render.bwd_rows);
    // contrib_sum is accumulated in 2^-32 fixed point (bit-reproducible) and converted once at the end
    unsigned long long *csum64 = f->rich_info ? gs.csum64 : nullptr;
#define TS2D_FWD_LAUNCH_CW(R, G, CW, ...)                                                                                              \
    do {                                                                                                                               \
        const size_t smem = CW * (size_t)FwdLayout<R>::BYTES;                                                                          \
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_render_fwd_fast<R, G, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_render_fwd_fast<R, G, CW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));         \
        TS2D_CUDA_TRY(ts2d_launch(k_render_fwd_fast<R, G, CW>, owned * (8 / CW), 32 * CW, smem, s, TS2D_FWD_ARGS, __VA_ARGS__));         \
    } while (0)
#define TS2D_FWD_LAUNCH(R, G, ...)                                                                                                     \
    do {                                                                                                                               \
        switch (ts2d_cta_warps()) {                                                                                                    \
        case 8: TS2D_FWD_LAUNCH_CW(R, G, 8, __VA_ARGS__); break;                                                                       \
        default: TS2D_FWD_LAUNCH_CW(R, G, 1, __VA_ARGS__); break;                                                                      \
        }                                                                                                                              \
    } while (0)
    if (f->rich_info) {
        if (!pre_cleared) {  // (the one-enqueue forward clears both at the start of the frame: no memset node between the chain's kernels)
            TS2D_CUDA_TRY(cudaMemsetAsync(gs.csum64, 0, sizeof(unsigned long long) * (size_t)g->P, s));
            TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_max, 0, sizeof(float) * (size_t)g->P, s));
        }
        if (g1) TS2D_FWD_LAUNCH(true, true, o_depth, o_normal, o_csum, o_cmax, is.lastw, rows_ctr, csum64);
        else TS2D_FWD_LAUNCH(true, false, o_depth, o_normal, o_csum, o_cmax, is.lastw, rows_ctr, csum64);
    } else {
        if (g1) TS2D_FWD_LAUNCH(false, true, nullptr, nullptr, o_csum, o_cmax, is.lastw, rows_ctr, csum64);
        else TS2D_FWD_LAUNCH(false, false, nullptr, nullptr, o_csum, o_cmax, is.lastw, rows_ctr, csum64);
    }
#undef TS2D_FWD_LAUNCH
#undef TS2D_FWD_LAUNCH_CW
#undef TS2D_FWD_ARGS
    TS2D_CUDA_TRY(cudaGetLastError());
    if (csum64) return ts2d_launch_contrib_finish(g->P, csum64, out->contrib_sum, s);
    return 0;
}
