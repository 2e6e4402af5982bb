// ts2d_prim3d_fast.cu -- fast composite kernels of the 3D primitive (flags.primitive == TS2D_PRIMITIVE_3D, flags.exact == 0).
//
// Same contracts as k_render3d_fwd / k_render3d_bwd (ts2d_prim3d.cu, the op-for-op mirrors of R3D/src/forward.cu:151-306 and
// R3D/src/backward.cu:215-454) on the machinery of the 2D fast kernels (ts2d_render_{fwd,bwd}_fast.cu):
//   * sub-tile coverage masks computed once per instance at emission (subtile_mask3d) and carried in the instance keys;
//   * warp-autonomous gather / stage / walk, one warp per 8x4-pixel sub-tile, no CTA barriers;
//   * the geometry half of every pair in the REFERENCE's arithmetic (geo3, ts2d_prim3d.cuh -- the reference's barycentrics are
//     ill-conditioned, so only its own operation sequence reproduces them), with dot(v1, n) and 1 / dot(n, n) read from the
//     raster record instead of being recomputed per pair (two IEEE divides and two dot products less);
//   * the opacity half (pow, exp) on the MUFU pipe inside decision bands; libdevice powf / expf only where a reference
//     decision (alpha or G against 1/255, op G against 0.99, T against 1e-4) is within rounding of its threshold.
//
// Backward: the reference issues 16 scalar atomics per contributing (pixel, triangle) pair (R3D/src/backward.cu:365,430-451).
// Here each pair parks FIVE scalars (contrib, dL/dop term, D = dL/d ecc routed to the arg-min barycentric, a1, a2) in a
// shared-memory panel; every 8 triangles the lanes switch roles (lane = (triangle, quarter of the pixels)) and sum moments of
// those scalars.  All sixteen per-triangle outputs are linear in the moments, because on the triangle's plane
//   pv_1 = -(a2 e2 + a3 e3),  pv_2 = pv_1 + e2,  pv_3 = pv_1 + e3,  cross(pv_j, pv_k) = a_i n,  depth = v1.z + a2 e2.z + a3 e3.z
// (e_k = v_k - v1), so every Jacobian of R3D/src/backward.cu:401-428 is affine in (a1, a2) with per-triangle vector
// coefficients f_k = (e_k x n) / |n|^2:
//   da1/dv2 = -a2 f2 + (a1 + a2) f3      da1/dv3 = -(a1 + a3) f2 + a3 f3      da1/dn = -a1 n / |n|^2
//   da2/dv1 =  a2 f2 - (a1 + a2) f3      da2/dv3 = -a2 f2 - a3 f3             da2/dn = -a2 n / |n|^2
//   da1/ddepth = ray . (f2 - f3) = (k2 - k3 - a2 - a3) / depth        da2/ddepth = ray . f3 = (k3 + a2) / depth      (k_j = v1 . f_j)
//   ddepth/dv1 = n / (ray . n) = n depth / K                            ddepth/dn = pv_1 depth / K                     (K = v1 . n)
// These forms are also better conditioned than the reference's per-pair cross products of cancelling differences, so the
// gradients sit closer to the fp64 truth than the reference's own (tests/test_gpu_parity.py: test_3d_gradient_accuracy_vs_truth).
// One 64 B RED burst per (warp, triangle) replaces up to 32 x 16 scalar atomics.
#include "ts2d_prim3d.cuh"

namespace {

constexpr int F3_EB = 80;       // staged entry: {v1 v2.x} {v2.yz v3.xy} {v3.z n} {r g b op} {K 1/nn id pos}  (K in the rounding of the pass)
constexpr int FW3_PROW = 33;    // forward contrib panel row stride (words)

template <bool RICH>
struct Fwd3Layout {
    static constexpr int PANEL = 32 * F3_EB;
    static constexpr int BYTES = RICH ? 32 * F3_EB + 32 * FW3_PROW * 4 : 32 * F3_EB;
};

__device__ __forceinline__ f3 pixel_ray(int px, int py, int W, int H, float tfx, float tfy)
{
    return mk3(tfx * pix_to_proj((float)px, W), tfy * pix_to_proj((float)py, H), 1.0f);  // R3D/src/forward.cu:187
}

// Gather step shared by both kernels: collect up to CAP list positions whose sub-tile bit is set.  `pend` holds the covered,
// not yet collected lanes of the chunk that is being consumed (a chunk may be split between two rounds); next() loads the
// following chunk's ballot and returns the list position lane l looked at (position = origin(l)).
template <int CAP, int POS_OFF, typename NextChunk>
__device__ __forceinline__ int gather_round(uint32_t sb, uint32_t &pend, uint32_t &pend_pos, int lane, uint32_t lt_mask, NextChunk next)
{
    int count = 0;
    while (count < CAP) {
        if (pend == 0u) {
            if (!next(pend, pend_pos)) break;
            if (pend == 0u) continue;
        }
        const int n = __popc(pend), take = min(n, CAP - count);
        const bool mine = (pend >> lane) & 1u;
        const int rank = __popc(pend & lt_mask);
        const bool sel = mine && rank < take;
        if (sel) sts32(sb + POS_OFF + (count + rank) * F3_EB, pend_pos);
        pend &= ~__ballot_sync(0xffffffffu, sel);
        count += take;
    }
    return count;
}

// Stage: lane i fetches the record of collected entry i (`base`: list index of tile-relative position 0).
// BWD: the gather left (position | live sub-tile bits << 24) in the position word; the row this (sub-tile, entry) pair owns
// (ts2d_bwd_reduce.cu) goes to slot_base + 4 * lane.
template <bool BWD>
__device__ __forceinline__ void stage_entry(uint32_t sb, int lane, int count, const uint32_t *__restrict__ list, uint32_t base,
                                            const float4 *__restrict__ rec0, const float4 *__restrict__ rec1, int warp = 0,
                                            const uint32_t *__restrict__ ei_of = nullptr, const uint32_t *__restrict__ sbase = nullptr,
                                            uint32_t slot_base = 0)
{
    if (lane < count) {
        const uint32_t ea = sb + lane * F3_EB;
        uint32_t pos = lds32(ea + 76);
        if (BWD) {
            const uint32_t bits = pos >> 24;
            pos &= 0xFFFFFFu;
            sts32(slot_base + 4 * lane, __ldg(sbase + __ldg(ei_of + base + pos)) + __popc(bits & ((1u << warp) - 1u)));
        }
        const uint32_t id = list[base + pos];
        const float4 *r = rec0 + 3 * (size_t)id;
        const float4 *q = rec1 + 2 * (size_t)id;
        const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), q0 = __ldg(q);
        const float4 q1v = __ldg(q + 1);  // {K_fwd, 1/nn, K_bwd, -}
        const float2 q1 = make_float2(BWD ? q1v.z : q1v.x, q1v.y);
        sts128(ea, r0);
        sts128(ea + 16, r1);
        sts128(ea + 32, r2);
        sts128(ea + 48, q0);
        sts128(ea + 64, make_float4(q1.x, q1.y, __uint_as_float(id), __uint_as_float(pos)));
    }
}

// Exact transmittance of the pixel with ray `ray` after visiting list positions [start, upto] (inclusive) in the reference's
// arithmetic and order (whole warp cooperates: 32 entries evaluated in parallel, the product chained serially).
__device__ __noinline__ float exact_T_upto3(const uint32_t *__restrict__ list, const float4 *__restrict__ rec0, const float4 *__restrict__ rec1,
                                            uint32_t start, uint32_t upto, float rx, float ry, float two_gamma, int lane)
{
    const f3 ray = mk3(rx, ry, 1.0f);
    float T = 1.0f;
    for (uint32_t k0 = start; k0 <= upto; k0 += 32) {
        const uint32_t k = k0 + lane;
        float f = 1.0f;
        if (k <= upto) {
            const uint32_t id = list[k];
            const Tri3 t = unpack3(__ldg(rec0 + 3 * (size_t)id), __ldg(rec0 + 3 * (size_t)id + 1), __ldg(rec0 + 3 * (size_t)id + 2));
            const float op = __ldg(&rec1[2 * (size_t)id].w);
            const float2 kq = __ldg(reinterpret_cast<const float2 *>(rec1 + 2 * (size_t)id + 1));
            Pair3 e;
            if (eval_pair3<false>(t, kq.x, kq.y, op, two_gamma, ray, e)) f = 1.0f - e.alpha;
        }
#pragma unroll
        for (int l = 0; l < 32; l++) T = __fmul_rn(T, __shfl_sync(0xffffffffu, f, l));  // x 1.0f is exact
    }
    return T;
}

// ------------------------------------------------------------------------------------------------ K7 (3D, fast)
template <bool RICH, bool GAMMA1, int CW>  // CW = warps per CTA (see k_render_fwd_fast): a CTA is only a unit of residency here
__global__ void __launch_bounds__(32 * CW, 24 / CW)
k_render3d_fwd_fast(int W, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, float tfx, float tfy,
                    const uint2 *__restrict__ ranges, const uint32_t *keys, const uint32_t *__restrict__ list,
                    const float4 *__restrict__ rec0, const float4 *__restrict__ rec1, float bg_depth, const float *__restrict__ bg_ptr, const float *__restrict__ background,
                    float *__restrict__ final_T, uint32_t *__restrict__ n_contrib, float *__restrict__ out_feature, float *__restrict__ out_depth,
                    float *__restrict__ out_normal, float *__restrict__ contrib_sum, float *__restrict__ contrib_max,
                    uint32_t *__restrict__ lastw, unsigned long long *__restrict__ bwd_rows, unsigned long long *__restrict__ csum64)
{
    ts2d_grid_chain();
    if (bg_ptr) bg_depth = __ldg(bg_ptr);  // model inputs: background depth computed on the device by K1
    using L = Fwd3Layout<RICH>;
    extern __shared__ __align__(16) unsigned char s_raw[];

    constexpr int PARTS = 8 / CW;
    const int tile = (blockIdx.x / PARTS) * shard_world + shard_rank;
    if (tile >= n_tiles) return;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int tid = threadIdx.x, lwarp = tid >> 5, warp = (blockIdx.x % PARTS) * CW + lwarp, lane = tid & 31;
    const int px = tile_x * TS2D_TILE + (warp & 1) * 8 + (lane & 7);
    const int py = tile_y * TS2D_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W && py < H;
    const f3 ray = pixel_ray(px, py, W, H, tfx, tfy);
    const Gamma3 gk = make_gamma3(GAMMA1 ? 1.0f : gamma, GAMMA1);
    const uint32_t sb = smem_base(s_raw + lwarp * L::BYTES);
    const uint32_t lt_mask = (1u << lane) - 1u;

    const uint2 range = ranges[tile];
    float T = 1.0f, Terr = 0.0f;  // Terr: bound on |T - (the reference's T)|
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f, accn0 = 0.f, accn1 = 0.f, accn2 = 0.f;
    uint32_t last = range.y - range.x;  // n_contrib if the pixel never saturates
    bool done = !inside;
    uint32_t rows = 0;  // (sub-tile, entry) pairs the backward pass will visit: one 64 B row each (ts2d_bwd_reduce.cu)

    auto flush_panel = [&](int visited) {
        __syncwarp();
        if (lane < visited) {
            const uint32_t row = sb + L::PANEL + lane * (FW3_PROW * 4);
            float s = 0.0f, m = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const float v = lds32f(row + 4 * i);
                s += v;
                m = fmaxf(m, v);
            }
            if (m > 0.0f) {
                const uint32_t id = lds32(sb + lane * F3_EB + 72);
                atomicAdd(csum64 + id, __float2ull_rn(s * 4294967296.0f));  // 2^-32 fixed point: integer sums do not depend on the order of the additions
                atomicMax((unsigned int *)contrib_max + id, __float_as_uint(m));  // contrib >= 0: bit order == value order
            }
            // (No coverage-bit clearing here, unlike the 2D kernel: the 3D reference's backward takes its skip decision on
            // G < 1/255 instead of alpha < 1/255 (R3D/src/backward.cu:351), so it visits pairs the forward never blended.)
        }
    };

    uint32_t cur = range.x;  // next list position to scan
    uint32_t kreg = (cur + lane < range.y) ? __ldg(keys + cur + lane) : 0u;  // keys[cur + lane], always one chunk ahead
    uint32_t pend = 0u, pend_pos = 0u;
    auto next_chunk = [&](uint32_t &p, uint32_t &ppos) {
        if (cur >= range.y) return false;
        p = __ballot_sync(0xffffffffu, (kreg >> warp) & 1u);
        ppos = cur + lane - range.x;  // tile-relative list position this lane looked at
        cur += 32;
        kreg = (cur + lane < range.y) ? __ldg(keys + cur + lane) : 0u;
        return true;
    };
    while (true) {
        if (__all_sync(0xffffffffu, done)) break;
        const int count = gather_round<32, 76>(sb, pend, pend_pos, lane, lt_mask, next_chunk);
        if (count == 0) break;  // list exhausted
        __syncwarp();
        stage_entry<false>(sb, lane, count, list, range.x, rec0, rec1);
        __syncwarp();
        // ---- walk
        uint32_t ea = sb, prow = sb + L::PANEL;
        int visited = count;
        for (int j = 0; j < count; j++, ea += F3_EB, prow += FW3_PROW * 4) {
            float contrib = 0.0f;  // > 0 <=> this lane blended the triangle
            bool evt = false;      // this lane's transmittance reached the 1e-4 cut or its error band: handled out of line below
            if (!done) {
                const Tri3 t = unpack3(lds128(ea), lds128(ea + 16), lds128(ea + 32));
                const float2 kq = lds64(ea + 64);
                Pair3 e;
                if (geo3<false>(t, kq.x, kq.y, ray, e)) {
                    const float4 col = lds128(ea + 48);
                    float og;
                    bool unc;
                    bool hit = alpha3_fast<false>(col.w, gk, e, og, unc);
                    if (unc) hit = alpha3_exact<false>(col.w, gk.two_gamma, e);
                    if (hit) {
                        contrib = e.alpha * T;
                        acc0 = fmaf(contrib, col.x, acc0);
                        acc1 = fmaf(contrib, col.y, acc1);
                        acc2 = fmaf(contrib, col.z, acc2);
                        if constexpr (RICH) {
                            accn0 = fmaf(contrib, t.n.x, accn0);
                            accn1 = fmaf(contrib, t.n.y, accn1);
                            accn2 = fmaf(contrib, t.n.z, accn2);
                            accd = fmaf(contrib, e.depth, accd);
                        }
                        const float om = 1.0f - e.alpha;
                        // |T_new - T_ref_new| <= |T - T_ref| * om + T * |d alpha| + rounding of the two product chains
                        Terr = fmaf(Terr, om, contrib * (unc ? 0.0f : fmaf(gk.c1, fabsf(e.power), gk.c0)));
                        T *= om;
                        Terr = fmaf(T, 1.3e-7f, Terr);
                        evt = (T - 0.0001f) <= Terr;  // saturated (T <= 1e-4) or within the error band of the cut
                    }
                }
            }
            if constexpr (RICH) sts32f(prow + lane * 4, contrib);
            if (__any_sync(0xffffffffu, evt)) {  // at most a few times per pixel: everything about saturation lives here
                const uint32_t pos = lds32(ea + 76);
                const float dT = T - 0.0001f;
                if (evt && dT <= 0.0f) {  // provisional when inside the band: re-decided below on the exact transmittance
                    done = true;
                    last = pos + 1;
                }
                uint32_t need = __ballot_sync(0xffffffffu, evt && fabsf(dT) <= Terr);
                while (need) {  // rare: exact transmittance re-walk for one pixel at a time, whole warp cooperating
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const float srx = __shfl_sync(0xffffffffu, ray.x, src), sry = __shfl_sync(0xffffffffu, ray.y, src);
                    const float Te = exact_T_upto3(list, rec0, rec1, range.x, range.x + pos, srx, sry, gk.two_gamma, lane);
                    if (lane == src) {
                        done = (Te <= 0.0001f);
                        last = done ? (pos + 1) : (range.y - range.x);
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
        rows += (uint32_t)visited;  // no coverage-bit clearing in 3D (see flush_panel): every visited entry gets a row in the backward pass
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

// ------------------------------------------------------------------------------------------------ K8 (3D, fast)
constexpr int B3_CAP = 16;      // entries staged per round
constexpr int B3_ROWS = 8;      // triangles per phase-2 panel
constexpr int B3_NS = 5;        // parked scalars per pair: contrib, dL/dop term, D | arg-min, a1, a2
constexpr int B3_WROW = B3_NS * 32 + 1;

// Per-warp shared-memory block:
//   ENT   B3_CAP staged entries (F3_EB bytes each)
//   F     per pixel {gp0 gp1 gp2 gd} {gn0 gn1 gn2 -}; pixel p at p * 32 + (p >> 3) * 16 (the four quarters of phase 2 read
//         four different rows with one LDS.128: the 16 B skew per quarter puts them on disjoint banks)
//   W     panel [B3_ROWS][B3_WROW]: [scalar * 32 + pixel]
//   INFO  per panel row {v1 v2.x}{v2.yz v3.xy}{v3.z n}{K 1/nn id row}   (64 B used, 80 B stride: conflict-free LDS.128 across rows)
constexpr int B3_INFO = 80;
struct Bwd3Layout {
    static constexpr int SLOT = B3_CAP * F3_EB;      // u32[B3_CAP]: row index of (this sub-tile, staged entry j)
    static constexpr int F = B3_CAP * F3_EB + B3_CAP * 4;
    static constexpr int W = F + 32 * 32 + 64;
    static constexpr int INFO = W + ((B3_ROWS * B3_WROW * 4 + 15) / 16) * 16;
    static constexpr int BYTES = INFO + B3_ROWS * B3_INFO;
};
__device__ __forceinline__ uint32_t f_row(int p) { return (uint32_t)(p * 32 + (p >> 3) * 16); }

// Phase 2 (out of line).  Lane (k, quarter) sums panel row k over pixels quarter*8 .. quarter*8+7, the four quarters are
// combined with two xor-shuffles, every lane maps the moments to the 16 accumulator components and issues one 16 B RED.
// Accumulator line (GACC_STRIDE floats): [0..2] dL/dv1_view [3..5] dL/dv2_view [6..8] dL/dv3_view [9..11] dL/dn [12..14] dL/drgb [15] dL/dop
template <bool GEO>
static __device__ __noinline__ void bwd3_flush_panel(uint32_t wb, uint32_t ib, uint32_t fb, float4 *__restrict__ rows, uint32_t rows_cap, int filled, int lane)
{
    const int k = lane & 7, quarter = lane >> 3;
    __syncwarp();
    float s_c0 = 0.f, s_c1 = 0.f, s_c2 = 0.f, s_op = 0.f, s_n0 = 0.f, s_n1 = 0.f, s_n2 = 0.f;
    float U1 = 0.f, U1a1 = 0.f, U1a2 = 0.f, U2 = 0.f, U2a1 = 0.f, U2a2 = 0.f, E0 = 0.f, Ea2 = 0.f, Ea3 = 0.f;
    Tri3 t;
    t.v1 = t.v2 = t.v3 = t.n = mk3(0.f, 0.f, 0.f);
    float inv_nn = 0.f, invK = 0.f;
    uint32_t slot = 0;
    f3 e2 = mk3(0.f, 0.f, 0.f), e3 = e2, f2 = e2, f3v = e2;
    if (k < filled) {
        const uint32_t ia = ib + B3_INFO * k;
        t = unpack3(lds128(ia), lds128(ia + 16), lds128(ia + 32));
        const float4 kq = lds128(ia + 48);
        inv_nn = kq.y;
        invK = fabsf(kq.x) > 1.0e-30f ? 1.0f / kq.x : 0.0f;
        slot = __float_as_uint(kq.w);
        e2 = t.v2 - t.v1;
        e3 = t.v3 - t.v1;
        f2 = cross3(e2, t.n) * inv_nn;
        f3v = cross3(e3, t.n) * inv_nn;
        const float k2 = dot3(t.v1, f2), k3 = dot3(t.v1, f3v), k23 = k2 - k3;
        const uint32_t row = wb + (k * B3_WROW + quarter * 8) * 4;
        const uint32_t frow = fb + f_row(quarter * 8);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float c = lds32f(row + 4 * i), w1 = lds32f(row + 4 * (32 + i)), Dp = lds32f(row + 4 * (64 + i));
            const float a1 = lds32f(row + 4 * (96 + i)), a2 = lds32f(row + 4 * (128 + i));
            const float4 f0 = lds128(frow + 32 * i);
            s_c0 = fmaf(c, f0.x, s_c0);
            s_c1 = fmaf(c, f0.y, s_c1);
            s_c2 = fmaf(c, f0.z, s_c2);
            s_op += w1;
            // arg-min barycentric (1, 2, 3) in the two LSBs of D: (u1, u2) = (D, 0), (0, D), (-D, -D)  [da3 = -da1 - da2]
            const uint32_t db = __float_as_uint(Dp);
            const bool to1 = (db & 1u) != 0u, to2 = (db & 2u) != 0u;
            const float Ds = (to1 && to2) ? -Dp : Dp;
            const float u1 = to1 ? Ds : 0.0f, u2 = to2 ? Ds : 0.0f;
            const float a3 = 1.0f - a1 - a2;
            U1 += u1;
            U1a1 = fmaf(u1, a1, U1a1);
            U1a2 = fmaf(u1, a2, U1a2);
            U2 += u2;
            U2a1 = fmaf(u2, a1, U2a1);
            U2a2 = fmaf(u2, a2, U2a2);
            // E = dL/ddepth / (ray . n);  dL/ddepth = gd contrib + (u1 (k23 - a2 - a3) + u2 (k3 + a2)) / depth,  1 / (ray . n) = depth / K
            float E = (u1 * (k23 - a2 - a3) + u2 * (k3 + a2)) * invK;
            if (GEO) {
                const float4 f1 = lds128(frow + 32 * i + 16);
                s_n0 = fmaf(c, f1.x, s_n0);
                s_n1 = fmaf(c, f1.y, s_n1);
                s_n2 = fmaf(c, f1.z, s_n2);
                const float depth = fmaf(a2, e2.z, fmaf(a3, e3.z, t.v1.z));
                E = fmaf(f0.w * c, depth * invK, E);
            }
            E0 += E;
            Ea2 = fmaf(E, a2, Ea2);
            Ea3 = fmaf(E, a3, Ea3);
        }
    }
#define XQ(v) v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16)
    XQ(s_c0); XQ(s_c1); XQ(s_c2); XQ(s_op); XQ(U1); XQ(U1a1); XQ(U1a2); XQ(U2); XQ(U2a1); XQ(U2a2); XQ(E0); XQ(Ea2); XQ(Ea3);
    if (GEO) { XQ(s_n0); XQ(s_n1); XQ(s_n2); }
#undef XQ
    if (k < filled) {
        const float cA2 = -U1a2, cB2 = U1a1 + U1a2;                                       // G2 = cA2 f2 + cB2 f3
        const float cA3 = -(U1 - U1a2 + U2a2), cB3 = U1 - U1a1 - U1a2 - U2 + U2a1 + U2a2; // G3
        const float cA1 = U2a2, cB1 = -(U2a1 + U2a2);                                      // G1 (+ E0 n)
        float4 q;  // the lane's quarter of the accumulator line; one 16-byte store into the pair's own row, no atomics
        if (quarter == 0) {
            const f3 G1 = cA1 * f2 + cB1 * f3v + E0 * t.n;
            q = make_float4(G1.x, G1.y, G1.z, fmaf(cA2, f2.x, cB2 * f3v.x));
        } else if (quarter == 1) {
            q = make_float4(fmaf(cA2, f2.y, cB2 * f3v.y), fmaf(cA2, f2.z, cB2 * f3v.z), fmaf(cA3, f2.x, cB3 * f3v.x), fmaf(cA3, f2.y, cB3 * f3v.y));
        } else if (quarter == 2) {
            const float cn = -inv_nn * (U1a1 + U2a2);
            const f3 GN = mk3(s_n0, s_n1, s_n2) + cn * t.n - (Ea2 * e2 + Ea3 * e3);
            q = make_float4(fmaf(cA3, f2.z, cB3 * f3v.z), GN.x, GN.y, GN.z);
        } else {
            q = make_float4(s_c0, s_c1, s_c2, s_op);
        }
        if (slot < rows_cap) st_stream128(rows + 4 * (size_t)slot + quarter, q);  // written once, read once by the row reduction: keep it out of the way of the raster records in L2
    }
    __syncwarp();
}

template <bool RICH, bool GAMMA1, int CW>
__global__ void __launch_bounds__(32 * CW, 24 / CW)
k_render3d_bwd_fast(int W_, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, float tfx, float tfy,
                    const uint2 *__restrict__ ranges, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ list,
                    const float4 *__restrict__ rec0, const float4 *__restrict__ rec1, float bg_depth, const float *__restrict__ bg_ptr, const float *__restrict__ background,
                    const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib, const float *__restrict__ dL_dout_feature,
                    const float *__restrict__ dL_dout_depth, const float *__restrict__ dL_dout_normal, const uint32_t *__restrict__ ei_of,
                    const uint32_t *__restrict__ sbase, float4 *__restrict__ rows, uint32_t rows_cap)
{
    ts2d_grid_chain();
    if (bg_ptr) bg_depth = __ldg(bg_ptr);  // model inputs: background depth computed on the device by K1
    using L = Bwd3Layout;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    constexpr int PARTS = 8 / CW;
    const int tile = (blockIdx.x / PARTS) * shard_world + shard_rank;
    if (tile >= n_tiles) return;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int tid = threadIdx.x, lwarp = tid >> 5, warp = (blockIdx.x % PARTS) * CW + lwarp, lane = tid & 31;
    const int px = tile_x * TS2D_TILE + (warp & 1) * 8 + (lane & 7), py = tile_y * TS2D_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < W_ && py < H;
    const size_t pix = (size_t)W_ * py + px;
    const size_t HW = (size_t)H * W_;
    const f3 ray = pixel_ray(px, py, W_, H, tfx, tfy);
    const Gamma3 gk = make_gamma3(GAMMA1 ? 1.0f : gamma, GAMMA1);
    const uint32_t sb = smem_base(smem_raw + lwarp * L::BYTES);
    const uint32_t lt_mask = (1u << lane) - 1u;

    const uint2 range = ranges[tile];
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
    sts128(sb + L::F + f_row(lane), make_float4(gp0, gp1, gp2, gd));  // per-pixel table for phase 2
    sts128(sb + L::F + f_row(lane) + 16, make_float4(gn0, gn1, gn2, 0.0f));
    // upstream normal / depth gradients all zero in this sub-tile: their terms are exactly zero for the reference too
    const bool geo = RICH && __any_sync(0xffffffffu, gd != 0.0f || gn0 != 0.0f || gn1 != 0.0f || gn2 != 0.0f);
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
    __syncwarp();

    int prow = 0;  // next free panel row (rows persist across rounds)
    auto flush_panel = [&](int filled) {
        if (geo) bwd3_flush_panel<true>(sb + L::W, sb + L::INFO, sb + L::F, rows, rows_cap, filled, lane);
        else bwd3_flush_panel<false>(sb + L::W, sb + L::INFO, sb + L::F, rows, rows_cap, filled, lane);
    };

    // Back to front: `rem` list positions [0, rem) (tile-relative) are still to be scanned; a chunk is the 32 positions below
    // rem, lane l looking at position rem - 1 - l (so ballot order == visiting order).
    uint32_t rem = warp_last;
    uint32_t kreg = (lane < rem) ? __ldg(keys + range.x + rem - 1 - lane) : 0u;  // always one chunk ahead
    uint32_t pend = 0u, pend_pos = 0u;
    auto next_chunk = [&](uint32_t &p, uint32_t &ppos) {
        if (rem == 0u) return false;
        p = __ballot_sync(0xffffffffu, (kreg >> warp) & 1u);
        ppos = (rem - 1u - lane) | (kreg << 24);  // position | live sub-tile bits; garbage on lanes >= rem, whose ballot bit is 0
        rem -= min(rem, 32u);
        kreg = (lane < rem) ? __ldg(keys + range.x + rem - 1 - lane) : 0u;
        return true;
    };
    while (true) {
        const int count = gather_round<B3_CAP, 76>(sb, pend, pend_pos, lane, lt_mask, next_chunk);
        if (count == 0) break;
        __syncwarp();
        stage_entry<true>(sb, lane, count, list, range.x, rec0, rec1, warp, ei_of, sbase, sb + L::SLOT);
        __syncwarp();
        // ---- walk
        uint32_t ea = sb;
        for (int j = 0; j < count; j++, ea += F3_EB) {
            const float4 e0 = lds128(ea), e1 = lds128(ea + 16), e2 = lds128(ea + 32), e4 = lds128(ea + 64);
            const uint32_t pos = __float_as_uint(e4.w);
            const uint32_t info_slot = lds32(sb + L::SLOT + 4 * j);  // for the panel's INFO row: requested early, used after the pair evaluation
            float w_c = 0.0f, w_op = 0.0f, w_D = 0.0f, w_a1 = 0.0f, w_a2 = 0.0f;
            bool visited = false;  // NOT "contrib != 0": the backward's cut is on G, so a pair with op == 0 still feeds dL/dop
            if (pos < last) {
                const Tri3 t = unpack3(e0, e1, e2);
                Pair3 e;
                if (geo3<true>(t, e4.x, e4.y, ray, e)) {
                    const float4 col = lds128(ea + 48);
                    float og;
                    bool unc;
                    bool hit = alpha3_fast<true>(col.w, gk, e, og, unc);
                    unc = unc || (fabsf(og - 0.99f) <= 0.99f * gk.band);  // the clamp decision of dL/dpower (op G < 0.99)
                    if (unc) {
                        hit = alpha3_exact<true>(col.w, gk.two_gamma, e);
                        og = col.w * e.G;
                    }
                    if (hit) {
                        visited = true;
                        const float om = 1.0f - e.alpha;
                        T = T * rcp_approx(om);
                        w_c = e.alpha * T;
                        float dL_dcontrib = fmaf(gp2, col.z - acc2, fmaf(gp1, col.y - acc1, gp0 * (col.x - acc0)));
                        acc0 = fmaf(e.alpha, col.x, om * acc0);
                        acc1 = fmaf(e.alpha, col.y, om * acc1);
                        acc2 = fmaf(e.alpha, col.z, om * acc2);
                        if (geo) {
                            dL_dcontrib = fmaf(gn0, t.n.x - accn0, dL_dcontrib);
                            dL_dcontrib = fmaf(gn1, t.n.y - accn1, dL_dcontrib);
                            dL_dcontrib = fmaf(gn2, t.n.z - accn2, dL_dcontrib);
                            accn0 = fmaf(e.alpha, t.n.x, om * accn0);
                            accn1 = fmaf(e.alpha, t.n.y, om * accn1);
                            accn2 = fmaf(e.alpha, t.n.z, om * accn2);
                            dL_dcontrib = fmaf(gd, e.depth - accd, dL_dcontrib);
                            accd = fmaf(e.alpha, e.depth, om * accd);
                        }
                        const float dL_dalpha = dL_dcontrib * T;
                        w_op = dL_dalpha * e.G;  // unconditional (R3D/src/backward.cu:451)
                        const float dL_dpower = (og < 0.99f) ? dL_dalpha * e.alpha : 0.0f;
                        const float D = -3.0f * dL_dpower * gk.two_gamma * e.power * rcp_approx(e.ecc + TS2D_EPS);
                        // sub-gradient of min: first arg-min in the order a1, a2, a3 (R3D/src/backward.cu:389-399)
                        const uint32_t sel = (e.a1 <= e.a2 && e.a1 <= e.a3) ? 1u : ((e.a2 <= e.a1 && e.a2 <= e.a3) ? 2u : 3u);
                        w_D = __uint_as_float((__float_as_uint(D) & ~3u) | sel);
                        w_a1 = e.a1;
                        w_a2 = e.a2;
                    }
                }
            }
            (void)visited;  // every staged entry owns a row and is flushed (zeros when no pixel of the sub-tile visited the pair)
            const uint32_t row = sb + L::W + (prow * B3_WROW + lane) * 4;
            sts32f(row, w_c);
            sts32f(row + 128, w_op);
            sts32f(row + 256, w_D);
            sts32f(row + 384, w_a1);
            sts32f(row + 512, w_a2);
            {   // row info for phase 2; every lane stores the same words
                const uint32_t ia = sb + L::INFO + prow * B3_INFO;
                sts128(ia, e0);
                sts128(ia + 16, e1);
                sts128(ia + 32, e2);
                sts128(ia + 48, make_float4(e4.x, e4.y, e4.z, __uint_as_float(info_slot)));
            }
            if (++prow == B3_ROWS) {
                flush_panel(B3_ROWS);
                prow = 0;
            }
        }
        __syncwarp();  // the next gather overwrites positions / entries
    }
    if (prow) flush_panel(prow);
}

}  // namespace

int ts2d_launch_render3d_fwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *keys,
                                  const uint32_t *list, ImageState is, const ts2d_forward_out *out, bool pre_cleared, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    if (owned <= 0) return 0;
    const bool g1 = g->gamma == 1.0f;
#define TS2D_F3_ARGS                                                                                                                     \
    W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, cam->tan_fovx, cam->tan_fovy, is.ranges, keys, list, gs.rec0,       \
        gs.rec1, g->background_depth, ts2d_bg_ptr(g, gs), g->background, is.final_T, is.n_contrib, o_feature
    float *o_feature = out->out_feature, *o_depth = out->depth, *o_normal = out->normal, *o_csum = out->contrib_sum, *o_cmax = out->contrib_max;
    unsigned long long *rows_ctr = reinterpret_cast<unsigned long long *>(&gs.hdr->render.bwd_rows);
    // contrib_sum is accumulated in 2^-32 fixed point (bit-reproducible) and converted once at the end
    unsigned long long *csum64 = f->rich_info ? gs.csum64 : nullptr;
#define TS2D_F3_LAUNCH_CW(R, G, CW, ...)                                                                                               \
    do {                                                                                                                               \
        const size_t smem = CW * (size_t)Fwd3Layout<R>::BYTES;                                                                         \
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_render3d_fwd_fast<R, G, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_render3d_fwd_fast<R, G, CW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));       \
        TS2D_CUDA_TRY(ts2d_launch(k_render3d_fwd_fast<R, G, CW>, owned * (8 / CW), 32 * CW, smem, s, TS2D_F3_ARGS, __VA_ARGS__));        \
    } while (0)
#define TS2D_F3_LAUNCH(R, G, ...)                                                                                                      \
    do {                                                                                                                               \
        if (ts2d_cta_warps() == 8) TS2D_F3_LAUNCH_CW(R, G, 8, __VA_ARGS__);                                                            \
        else TS2D_F3_LAUNCH_CW(R, G, 1, __VA_ARGS__);                                                                                  \
    } while (0)
    if (f->rich_info) {
        if (!pre_cleared) {
            TS2D_CUDA_TRY(cudaMemsetAsync(gs.csum64, 0, sizeof(unsigned long long) * (size_t)g->P, s));
            TS2D_CUDA_TRY(cudaMemsetAsync(out->contrib_max, 0, sizeof(float) * (size_t)g->P, s));
        }
        if (g1) TS2D_F3_LAUNCH(true, true, o_depth, o_normal, o_csum, o_cmax, is.lastw, rows_ctr, csum64);
        else TS2D_F3_LAUNCH(true, false, o_depth, o_normal, o_csum, o_cmax, is.lastw, rows_ctr, csum64);
    } else {
        if (g1) TS2D_F3_LAUNCH(false, true, nullptr, nullptr, o_csum, o_cmax, is.lastw, rows_ctr, csum64);
        else TS2D_F3_LAUNCH(false, false, nullptr, nullptr, o_csum, o_cmax, is.lastw, rows_ctr, csum64);
    }
#undef TS2D_F3_LAUNCH
#undef TS2D_F3_LAUNCH_CW
#undef TS2D_F3_ARGS
    TS2D_CUDA_TRY(cudaGetLastError());
    if (csum64) return ts2d_launch_contrib_finish(g->P, csum64, out->contrib_sum, s);
    return 0;
}

int ts2d_launch_render3d_bwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *keys,
                                  const uint32_t *list, ImageState is, const ts2d_loss_in *loss, BwdScratch sc, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    if (owned <= 0) return 0;
    const bool g1 = g->gamma == 1.0f;
    const uint32_t rows_cap = (uint32_t)(sc.rows_cap < 0xFFFFFFFFll ? sc.rows_cap : 0xFFFFFFFFll);
#define TS2D_B3_ARGS                                                                                                                     \
    W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, cam->tan_fovx, cam->tan_fovy, is.ranges, keys, list, gs.rec0,       \
        gs.rec1, g->background_depth, ts2d_bg_ptr(g, gs), g->background, is.final_T, is.n_contrib, loss->dL_dout_feature
#define TS2D_B3_LAUNCH_CW(R, G, CW, ...)                                                                                               \
    do {                                                                                                                               \
        const size_t smem = CW * (size_t)Bwd3Layout::BYTES;                                                                            \
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_render3d_bwd_fast<R, G, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_render3d_bwd_fast<R, G, CW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));       \
        TS2D_CUDA_TRY(ts2d_launch(k_render3d_bwd_fast<R, G, CW>, owned * (8 / CW), 32 * CW, smem, s, TS2D_B3_ARGS, __VA_ARGS__, sc.ei, sc.sbase, sc.rows, rows_cap)); \
    } while (0)
#define TS2D_B3_LAUNCH(R, G, ...)                                                                                                      \
    do {                                                                                                                               \
        if (ts2d_cta_warps() == 8) TS2D_B3_LAUNCH_CW(R, G, 8, __VA_ARGS__);                                                            \
        else TS2D_B3_LAUNCH_CW(R, G, 1, __VA_ARGS__);                                                                                  \
    } while (0)
    if (f->rich_info) {
        if (g1) TS2D_B3_LAUNCH(true, true, loss->dL_dout_depth, loss->dL_dout_normal);
        else TS2D_B3_LAUNCH(true, false, loss->dL_dout_depth, loss->dL_dout_normal);
    } else {
        if (g1) TS2D_B3_LAUNCH(false, true, nullptr, nullptr);
        else TS2D_B3_LAUNCH(false, false, nullptr, nullptr);
    }
#undef TS2D_B3_LAUNCH
#undef TS2D_B3_LAUNCH_CW
#undef TS2D_B3_ARGS
    return (int)cudaGetLastError();
}
