// ts2d_render_bwd_fast.cu -- fast reverse-walk gradient accumulation (K8, flags.exact == 0).
//
// Same contract as k_render_bwd (ts2d_render_bwd.cu, the mirror of R2D/src/backward.cu:265-493) with the
// forward fast kernel's machinery (warp-autonomous gather / stage / walk over the instance keys' sub-tile masks,
// reference-shaped fast barycentrics, decision bands with eval_exact fallback -- so the set of contributing pairs
// is exactly the one the forward pass blended).  The walk runs back to front and starts at the sub-tile's own
// last contributor (max n_contrib over the warp's 32 pixels), not at the end of the tile's list.
//
// What is different from the reference is how the per-pair contributions are summed over pixels.  The reference
// issues 10-16 global atomics per contributing (pixel, triangle) pair.  Here every one of the 16 per-triangle
// outputs is written as  sum_p w(p) * f(p)  where w is one of only THREE per-pair scalars that depend on the
// sequential walk (contrib = alpha T, dL/dalpha * G, and D = dL/d ecc routed to the arg-min barycentric) and
// f is a per-pixel constant (upstream gradients, pixel offsets) -- the barycentrics being affine in the pixel,
// their Jacobians reduce to first moments (see ts2d_preprocess.cu: the moments -> vertex-gradient map).
//   phase 1 (lane = pixel):    walk the staged entries, run the T / colour recurrences, park the three scalars of
//                              each pair in an 8-row shared-memory panel W[row][pixel];
//   phase 2 (lane = triangle): every 8 rows, each lane owns (row, quarter of the 32 pixels) and accumulates the
//                              16 sums with plain FFMAs from W and the per-pixel table F -- no per-pair shuffles,
//                              no selects -- then two xor-combines and one 16-byte STORE per lane (out of line) into the
//                              64 B row this (sub-tile, list entry) pair owns: no atomics anywhere, the per-triangle sums are
//                              finished by a sequential reduction in a fixed order (ts2d_bwd_reduce.cu) -- bit-reproducible.
// This replaces a 16-value warp butterfly per pair-iteration (16 SHFL + 16 FADD + 30 SEL) by 3 STS + ~30
// amortised instructions, and moves all Jacobian arithmetic out of the per-pixel loop.
#include "ts2d_fast.cuh"

namespace {

constexpr int BW_ROWS = 8;      // triangles per phase-2 panel
constexpr int BW_WROW = 99;     // 3 scalars x 32 pixels + the two arg-min ballots + 1 pad word (any odd stride is conflict-free for both phases)
constexpr int BW_BAL = 96 * 4;  // byte offset of the row's ballot words {to a1, to a2}

// Per-warp shared-memory block (byte offsets from the warp's base address):
//   ENT   entry j at j * EB: {v1.x v1.y v2.x v2.y} {v3.x v3.y 1/area2 op} {r g b id} [RICH: {n.x n.y n.z vd1} {vd2 vd3 pos -}]
//   POS   non-RICH only: u32[32] list positions (RICH keeps them in the entry's spare word)
//   F     per pixel {gp0 gp1 gp2 gd} {gn0 gn1 gn2 -}
//   W     panel [8 rows][99]: [scalar * 32 + pixel], then the two ballots that route D (bit p of word 96: pixel p's D goes to a1, word 97: to a2)
//   INFO  per panel row {v1 v2} {v3 1/area2 op} {id, row index} (48 B stride): what phase 2 needs to know about the triangle -- written only
//         for rows that outlive their gather round (the staged entries are overwritten by the next one); rows of the current round are
//         read from the staged entries themselves
template <bool RICH>
struct BwdLayout {
    static constexpr int EB = RICH ? 80 : 48;
    static constexpr int POS = RICH ? 72 : 32 * 48;        // tile-relative list position of entry j: POS + j * POS_STRIDE
    static constexpr int SLOT = RICH ? 76 : 32 * 48 + 128; // row index of (this sub-tile, entry j): SLOT + j * POS_STRIDE
    static constexpr int POS_STRIDE = RICH ? 80 : 4;
    static constexpr int F = RICH ? 32 * 80 : 32 * 48 + 256;
    static constexpr int W = F + 32 * 32 + 64;  // F rows are skewed by 16 B per group of 8 pixels, see f_row()
    static constexpr int INFO = W + ((BW_ROWS * BW_WROW * 4 + 15) / 16) * 16;
    static constexpr int BYTES = INFO + BW_ROWS * 48;
};

// F table: pixel p at p * 32 + (p >> 3) * 16.  In phase 2 the four pixel-quarters of a warp read four different rows with one
// LDS.128; rows 8 apart would sit 256 B apart = on the same banks (4-way conflict), the 16 B skew per quarter separates them.
__device__ __forceinline__ uint32_t f_row(int p) { return (uint32_t)(p * 32 + (p >> 3) * 16); }

// phase 2 (out of line: one copy keeps the kernel inside the instruction cache and out of the walk loop's register budget).
// lane (k, quarter) sums panel row k over pixels quarter*8 .. quarter*8+7, then flushes the triangle.
// wb / ib / fb: shared-space addresses of the warp's W panel, row ids and F table.
// GEO: some pixel of the sub-tile has a non-zero normal / depth upstream gradient (a template parameter, so that the common
// colour-only case does not even issue the geometry terms as predicated-off instructions).
// jbase: index (among the entries staged in the current round, at `eb`) of the entry panel row 0 belongs to; rows with jbase + k < 0
// were carried over from an earlier round and have their copy in INFO.
template <bool RICH, bool GEO>
static __device__ __noinline__ void bwd_flush_panel(uint32_t wb, uint32_t ib, uint32_t fb, uint32_t eb, int jbase, const float4 *__restrict__ rec1,
                                                    float4 *__restrict__ rows, uint32_t rows_cap, float ox, float oy,
                                                    float sub_x0, float sub_y0, int filled, int lane)
{
    using L = BwdLayout<RICH>;
    constexpr bool geo = GEO;
    const int k = lane & 7, quarter = lane >> 3;
    __syncwarp();
    // where this row's triangle record lives: the staged entry (current round) or its INFO copy (carried rows)
    const int ent = jbase + k;
    const uint32_t src = ent >= 0 ? eb + (uint32_t)ent * L::EB : ib + 48 * k;
    // sums that share their multiplier ride in pairs on the packed pipe: {c0, c1}, {n0, n1}, {u1, u2} and their x-moments;
    // the y-moments need no loop term at all: the lane's 8 pixels share one row, M_y = dy * S
    v2 s_c01 = bc(0.f), s_n01 = bc(0.f), U0 = bc(0.f), Ux = bc(0.f);
    float s_c2 = 0.f, s_n2 = 0.f, m0 = 0.f, m1 = 0.f, s_op = 0.f;
    const float dyp = sub_y0 + (float)quarter;
    uint32_t id = 0;
    float vd1 = 0.f, vd2 = 0.f, vd3 = 0.f;
    if (k < filled) {
        id = lds32(src + (ent >= 0 ? 44 : 32));
        if (geo) {  // vertex depths: re-read from L1/L2, in flight during the panel sums
            vd1 = __ldg(&rec1[2 * (size_t)id].w);
            const float2 q = __ldg(reinterpret_cast<const float2 *>(rec1 + 2 * (size_t)id + 1));
            vd2 = q.x;
            vd3 = q.y;
        }
        const uint32_t row = wb + (k * BW_WROW + quarter * 8) * 4;
        const uint32_t frow = fb + f_row(quarter * 8);
        // arg-min routing of D: bit i of bq1 / bq2 says that pixel quarter * 8 + i sends its D to a1 / a2 (both: the arg-min was a3 and
        // phase 1 stored -D)
        const uint32_t bq1 = lds32(wb + k * (BW_WROW * 4) + BW_BAL) >> (quarter * 8), bq2 = lds32(wb + k * (BW_WROW * 4) + BW_BAL + 4) >> (quarter * 8);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            // pixel p = quarter * 8 + i (lane index of phase 1) inside the sub-tile: x = p & 7 = i, y = p >> 3 = quarter
            const float dxp = (float)i;  // x offset inside the sub-tile (an immediate); sub_x0 is added once after the loop
            const float c = lds32f(row + 4 * i), w1 = lds32f(row + 4 * (32 + i)), Dp = lds32f(row + 4 * (64 + i));
            const float4 f0 = lds128(frow + 32 * i);
            s_c01 = fma2(bc(c), mk2v(f0.x, f0.y), s_c01);
            s_c2 = fmaf(c, f0.z, s_c2);
            s_op += w1;
            const v2 u = mk2v(((bq1 >> i) & 1u) ? Dp : 0.0f, ((bq2 >> i) & 1u) ? Dp : 0.0f);
            if (geo) {
                const float4 f1 = lds128(frow + 32 * i + 16);
                const float cg = c * f0.w;                       // contrib * gd
                s_n01 = fma2(bc(c), mk2v(f1.x, f1.y), s_n01);
                s_n2 = fmaf(c, f1.z, s_n2);
                m0 += cg;
                m1 = fmaf(cg, dxp, m1);
            }
            U0 = add2(U0, u);
            Ux = fma2(u, bc(dxp), Ux);
        }
    }
    m1 = fmaf(sub_x0, m0, m1);
    float s_c0 = s_c01.a, s_c1 = s_c01.b, s_n0 = s_n01.a, s_n1 = s_n01.b, m2 = dyp * m0;
    float u10 = U0.a, u20 = U0.b, u1x = fmaf(sub_x0, U0.a, Ux.a), u2x = fmaf(sub_x0, U0.b, Ux.b), u1y = dyp * U0.a, u2y = dyp * U0.b;
#define XQ(v) v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16)
    XQ(s_c0); XQ(s_c1); XQ(s_c2); XQ(s_op); XQ(u10); XQ(u1x); XQ(u1y); XQ(u20); XQ(u2x); XQ(u2y);
    if (geo) { XQ(s_n0); XQ(s_n1); XQ(s_n2); XQ(m0); XQ(m1); XQ(m2); }
#undef XQ
    if (k < filled) {
        const float4 e1 = lds128(src), e2 = lds128(src + 16);
        // the row of this (sub-tile, entry) pair
        const uint32_t slot = lds32(ent >= 0 ? (RICH ? src + 76 : eb + L::SLOT + 4 * (uint32_t)ent) : src + 36);
        const float inv = e2.z;
        const float p1x = e1.x - ox, p1y = e1.y - oy;
        float S1 = u10, M1x = u1x, M1y = u1y, S2 = u20, M2x = u2x, M2y = u2y;
        float gv0 = 0.f, gv1 = 0.f, gv2 = 0.f;
        if (geo) {
            // sum gd contrib a_i with the barycentrics expanded about v1, where a = (1, 0, 0): a_i(p) = a_i(v1) + grad a_i . (p - v1).
            // (Expanding about the tile origin instead needs a_i(origin), a cross product of vertex offsets of up to a tile plus the
            // dilated triangle: for a triangle of a few pixels that constant alone carries 1e-5 of rounding.)
            const float A1 = (e1.w - e2.y) * inv, B1 = (e2.x - e1.z) * inv;  // grad a_1
            const float A2 = (e2.y - e1.y) * inv, B2 = (e1.x - e2.x) * inv;  // grad a_2
            const float mqx = fmaf(-p1x, m0, m1), mqy = fmaf(-p1y, m0, m2);  // first moments of gd contrib about v1
            const float t1 = fmaf(B1, mqy, A1 * mqx), t2 = fmaf(B2, mqy, A2 * mqx);
            gv0 = m0 + t1;
            gv1 = t2;
            gv2 = -t1 - t2;  // a_3 = 1 - a_1 - a_2
            const float d13 = vd1 - vd3, d23 = vd2 - vd3;   // depth term of ga_k = (vd_k - vd_3) gd contrib
            S1 = fmaf(d13, m0, S1); M1x = fmaf(d13, m1, M1x); M1y = fmaf(d13, m2, M1y);
            S2 = fmaf(d23, m0, S2); M2x = fmaf(d23, m1, M2x); M2y = fmaf(d23, m2, M2y);
        }
        // moments about v1: q = p - v1 = d - (v1 - o)
        const float Q1x = fmaf(-p1x, S1, M1x), Q1y = fmaf(-p1y, S1, M1y), Q2x = fmaf(-p1x, S2, M2x), Q2y = fmaf(-p1y, S2, M2y);
        float4 q;  // the lane's quarter of the 16-float accumulator layout (GACC_STRIDE, ts2d_common.cuh)
        if (quarter == 0) q = make_float4(S1, Q1x, Q1y, S2);
        else if (quarter == 1) q = make_float4(Q2x, Q2y, s_op, s_n0);
        else if (quarter == 2) q = make_float4(s_c0, s_c1, s_c2, s_n1);
        else q = make_float4(s_n2, gv0, gv1, gv2);  // zeros without geometry gradients
        if (slot < rows_cap) st_stream128(rows + 4 * (size_t)slot + quarter, q);  // written once, read once by the row reduction: keep it out of the way of the raster records in L2
    }
    __syncwarp();
}

template <bool RICH, bool GAMMA1, int CW>  // CW = warps per CTA, see k_render_fwd_fast
__global__ void __launch_bounds__(32 * CW, 32 / CW)
k_render_bwd_fast(int W_, int H, int C, int gx, int shard_rank, int shard_world, int n_tiles, float gamma, const uint2 *__restrict__ ranges,
                  const uint32_t *__restrict__ keys, const uint32_t *__restrict__ list, const float4 *__restrict__ rec0,
                  const float4 *__restrict__ rec1, float bg_depth, const float *__restrict__ bg_ptr, const float *__restrict__ background, const float *__restrict__ final_T,
                  const uint32_t *__restrict__ n_contrib, const float *__restrict__ dL_dout_feature, const float *__restrict__ dL_dout_depth,
                  const float *__restrict__ dL_dout_normal, const uint32_t *__restrict__ ei_of, const uint32_t *__restrict__ sbase,
                  float4 *__restrict__ rows, uint32_t rows_cap)
{
    ts2d_grid_chain();
    if (bg_ptr) bg_depth = __ldg(bg_ptr);  // model inputs: background depth computed on the device by K1
    using L = BwdLayout<RICH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    constexpr int PARTS = 8 / CW;
    const int tile = (blockIdx.x / PARTS) * shard_world + shard_rank;
    if (tile >= n_tiles) return;
    const int tile_x = tile % gx, tile_y = tile / gx;
    const int tid = threadIdx.x, lwarp = tid >> 5, warp = (blockIdx.x % PARTS) * CW + lwarp, lane = tid & 31;
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    const int px = tile_x * TS2D_TILE + lx, py = tile_y * TS2D_TILE + ly;
    const bool inside = px < W_ && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float ox = (float)(tile_x * TS2D_TILE), oy = (float)(tile_y * TS2D_TILE);
    const size_t pix = (size_t)W_ * py + px;
    const size_t HW = (size_t)H * W_;
    GammaK gk = make_gamma(GAMMA1 ? 1.0f : gamma);  // gamma == 1: every constant of the error model folds to an immediate
    gk.is_one = GAMMA1;
    const uint32_t sb = smem_base(smem_raw + lwarp * L::BYTES);  // this warp's block
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t lane4 = 4u * lane;
    asm volatile("" : "+r"(lane4));  // opaque: keep the lane's panel offset in a register (ptxas re-read SR_TID in every walk iteration)

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
    }
    v2 acc01 = mk2v(acc0, acc1);       // running colour behind the current entry, channels {0, 1}
    const v2 gp01 = mk2v(gp0, gp1);
    if (inside) {
        if (RICH) {
            gn0 = dL_dout_normal[pix];
            gn1 = dL_dout_normal[HW + pix];
            gn2 = dL_dout_normal[2 * HW + pix];
            gd = dL_dout_depth[pix];
        }
    }
    sts128(sb + L::F + f_row(lane), make_float4(gp0, gp1, gp2, gd));  // per-pixel table for phase 2
    sts128(sb + L::F + f_row(lane) + 16, make_float4(gn0, gn1, gn2, 0.0f));
    // geometry upstream gradients all zero in this sub-tile (w_geometry = 0 training configs): the normal / depth
    // terms are exactly zero for the reference too, so skipping them changes no bit of the result
    const bool geo = RICH && __any_sync(0xffffffffu, gd != 0.0f || gn0 != 0.0f || gn1 != 0.0f || gn2 != 0.0f);
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
    __syncwarp();

    int prow = 0;   // next free panel row (rows persist across rounds)
    uint32_t urow = sb + L::W;  // ... and its shared-space address
    int jbase = 0;  // staged-entry index of panel row 0 (negative: the first -jbase rows were carried over from earlier rounds)
    const float sub_x0 = (float)((warp & 1) * 8), sub_y0 = (float)((warp >> 1) * 4);
    auto flush_panel = [&](int filled) {
        if (geo) bwd_flush_panel<RICH, RICH>(sb + L::W, sb + L::INFO, sb + L::F, sb, jbase, rec1, rows, rows_cap, ox, oy, sub_x0, sub_y0, filled, lane);
        else bwd_flush_panel<RICH, false>(sb + L::W, sb + L::INFO, sb + L::F, sb, jbase, rec1, rows, rows_cap, ox, oy, sub_x0, sub_y0, filled, lane);
    };

    // Back to front: `rem` list positions [range.x, range.x + rem) are still to be scanned; a chunk is the 32 positions
    // below range.x + rem, lane l looking at position range.x + rem - 1 - l (so ballot order == visiting order).
    uint32_t rem = warp_last;
    uint32_t kreg = (lane < rem) ? __ldg(keys + range.x + rem - 1 - lane) : 0u;  // always one chunk ahead
    while (rem > 0) {
        // ---- gather
        int count = 0;
        while (rem > 0) {
            const bool cov = (kreg >> warp) & 1u;
            const uint32_t b = __ballot_sync(0xffffffffu, cov);
            const int n = __popc(b);
            if (count + n > 32) break;  // this chunk opens the next round (kreg still holds it)
            // tile-relative position | live sub-tile bits << 24 (ts2d_bwd_reduce.cu: k_bwd_rows_mark left only the live ones in the key)
            if (cov) sts32(sb + L::POS + (count + __popc(b & lt_mask)) * L::POS_STRIDE, (rem - 1 - lane) | (kreg << 24));
            count += n;
            rem -= min(rem, 32u);
            kreg = (lane < rem) ? __ldg(keys + range.x + rem - 1 - lane) : 0u;
        }
        if (count == 0) break;
        __syncwarp();
        // ---- stage
        if (lane < count) {
            const uint32_t ea = sb + lane * L::EB;
            const uint32_t packed = lds32(sb + L::POS + lane * L::POS_STRIDE);
            const uint32_t prel = packed & 0xFFFFFFu, bits = packed >> 24;
            const uint32_t id = list[range.x + prel];
            // row of (this sub-tile, this entry): rows of an instance are numbered by sub-tile among its live bits
            const uint32_t slot = __ldg(sbase + __ldg(ei_of + range.x + prel)) + __popc(bits & ((1u << warp) - 1u));
            const float4 *r = rec0 + 3 * (size_t)id;
            const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
            sts128(ea, r0);
            sts128(ea + 16, r1);
            sts128(ea + 32, make_float4(r2.x, r2.y, r2.z, __uint_as_float(id)));
            if (RICH) {
                const float4 *q = rec1 + 2 * (size_t)id;
                const float4 q0 = __ldg(q), q1 = __ldg(q + 1);
                sts128(ea + 48, q0);
                sts32f(ea + 64, q1.x);  // +72 holds the list position, +76 the row index
                sts32f(ea + 68, q1.y);
            }
            sts32(sb + L::POS + lane * L::POS_STRIDE, prel);
            sts32(sb + L::SLOT + lane * L::POS_STRIDE, slot);
        }
        __syncwarp();
        // ---- walk
        uint32_t ea = sb;
        for (int j = 0; j < count; j++, ea += L::EB) {
            const float4 e1 = lds128(ea), e2 = lds128(ea + 16);
            const uint32_t pos = lds32(sb + L::POS + j * L::POS_STRIDE);
            float w_c = 0.0f, w_op = 0.0f, w_D = 0.0f;
            bool not1 = false, not2 = false;  // D does not go to a1 / a2
            {
                // evaluated by every lane, also those whose own last contributor lies in front of this entry (they are masked out
                // of the result): a branch around the evaluation would put the entry loads behind the position load and compare
                FastPair f;
                bool unc;
                bool hit = eval_fast(e1, e2, pxf, pyf, gk, f, unc);
                const bool live = pos < last;
                hit = hit && live;
                // one more reference decision lives in the backward pass: op*G < 0.99 (clamp not active)
                unc = unc || (fabsf(f.og - 0.99f) <= 0.99f * gk.band);
                {   // ... and so does the arg-min of (a1, a2, a3): dL/d min(a) goes to ONE barycentric, and near a corner of a thin
                    // triangle the choice changes the vertex gradient by its own size.  The fast a1, a2 are c * rn(1 / area2) where
                    // the reference divides (same c, bit for bit): |a_fast - a_ref| <= 1.5 * 2^-23 |a| for a1 and a2, their sum plus
                    // two roundings at 1 for a3 = 1 - a1 - a2; with every |a_i| <= 5.4 where a pair can contribute (ecc <= 7.4 at
                    // gamma = 0.6) the two smallest can swap only when they are within TS2D_ARGMIN_TIE of each other.
                    const float lo12 = fminf(f.a1, f.a2), hi12 = fmaxf(f.a1, f.a2);
                    unc = unc || (fmaxf(lo12, fminf(hi12, f.a3)) - fminf(lo12, f.a3) <= TS2D_ARGMIN_TIE);
                }
                if (unc && live) {
                    const float area2 = __ldg(&rec0[3 * (size_t)lds32(ea + 44) + 2].w);
                    PairEval e;
                    hit = eval_exact(e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, area2, e2.w, gk.two_gamma, pxf, pyf, e);
                    f.a1 = e.a1; f.a2 = e.a2; f.a3 = e.a3; f.ecc = e.ecc;
                    if (hit) { f.power = e.power; f.pw = -2.0f * e.power; f.G = e.G; f.og = __fmul_rn(e2.w, e.G); f.alpha = e.alpha; }
                }
                if (hit) {
                    const float4 col = lds128(ea + 32);
                    const float om = 1.0f - f.alpha;
                    T = T * rcp_approx(om);
                    w_c = f.alpha * T;
                    // channels 0 and 1 ride in one packed instruction each
                    const v2 col01 = mk2v(col.x, col.y);
                    const v2 t01 = mul2(gp01, sub2(col01, acc01));
                    float dL_dcontrib = fmaf(gp2, col.z - acc2, t01.a + t01.b);
                    acc01 = fma2(bc(f.alpha), col01, mul2(bc(om), acc01));
                    acc2 = fmaf(f.alpha, col.z, om * acc2);
                    if (geo) {
                        const float4 q0 = lds128(ea + 48);
                        const float2 q1 = lds64(ea + 64);
                        dL_dcontrib = fmaf(gn0, q0.x - accn0, dL_dcontrib);
                        dL_dcontrib = fmaf(gn1, q0.y - accn1, dL_dcontrib);
                        dL_dcontrib = fmaf(gn2, q0.z - accn2, dL_dcontrib);
                        accn0 = fmaf(f.alpha, q0.x, om * accn0);
                        accn1 = fmaf(f.alpha, q0.y, om * accn1);
                        accn2 = fmaf(f.alpha, q0.z, om * accn2);
                        const float depth = fmaf(q1.y, f.a3, fmaf(q0.w, f.a1, q1.x * f.a2));
                        dL_dcontrib = fmaf(gd, depth - accd, dL_dcontrib);
                        accd = fmaf(f.alpha, depth, om * accd);
                    }
                    const float dL_dalpha = dL_dcontrib * T;
                    w_op = dL_dalpha * f.G;  // unconditional (backward.cu:490)
                    const float dL_dpower = (f.og < 0.99f) ? dL_dalpha * f.alpha : 0.0f;
                    // -3 dL_dpower * 2 gamma * power / (ecc + eps).  gamma == 1: -3 * 2 * (-pw / 2) = 3 pw with the factors of two exact, so
                    // (3 dL_dpower) pw carries the same roundings as the general expression in two multiplies instead of four
                    const float rec = rcp_approx(f.ecc + TS2D_EPS);
                    const float D = gk.is_one ? ((3.0f * dL_dpower) * f.pw) * rec : -3.0f * dL_dpower * gk.two_gamma * f.power * rec;
                    // sub-gradient of min: first arg-min in the order a1, a2, a3 (backward.cu:449-461).  a3 = 1 - a1 - a2, so D to a3 is -D
                    // to both a1 and a2: the panel gets the signed value, the choice travels as two warp ballots
                    not2 = f.a1 <= f.a2 && f.a1 <= f.a3;           // arg-min a1
                    not1 = !not2 && f.a2 <= f.a1 && f.a2 <= f.a3;  // arg-min a2
                    w_D = (not1 || not2) ? D : -D;
                }
            }
            // every staged entry owns a row (its live bit is set), so every staged entry is flushed -- also when no pixel of the
            // sub-tile turns out to contribute (non-rich forward: conservative coverage bits): the row is then written as zeros
            const uint32_t to1 = __ballot_sync(0xffffffffu, !not1), to2 = __ballot_sync(0xffffffffu, !not2);
            sts32f(urow + lane4, w_c);
            sts32f(urow + lane4 + 128, w_op);
            sts32f(urow + lane4 + 256, w_D);
            sts32(urow + BW_BAL, to1);  // every lane stores the same words (cheaper than electing one: no predicate)
            sts32(urow + BW_BAL + 4, to2);
            urow += BW_WROW * 4;
            if (++prow == BW_ROWS) {
                flush_panel(BW_ROWS);
                prow = 0;
                urow = sb + L::W;
                jbase = j + 1;
            }
        }
        // rows that stay in the panel: the staged entries they point to are overwritten by the next gather, keep a copy of what
        // phase 2 needs (at most 7 rows per round instead of a copy per visited entry)
        {
            const int ent = jbase + lane;
            if (lane < prow && ent >= 0) {
                const uint32_t src = sb + (uint32_t)ent * L::EB, dst = sb + L::INFO + 48 * lane;
                const float4 c1 = lds128(src), c2 = lds128(src + 16);
                const uint32_t cid = lds32(src + 44), cslot = lds32(sb + L::SLOT + ent * L::POS_STRIDE);
                sts128(dst, c1);
                sts128(dst + 16, c2);
                sts32(dst + 32, cid);
                sts32(dst + 36, cslot);
            }
            jbase = -prow;
        }
        __syncwarp();  // the next gather overwrites positions / entries
    }
    if (prow) flush_panel(prow);
}

}  // namespace

int ts2d_launch_render_bwd_fast(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, GeomState gs, const uint32_t *keys,
                                const uint32_t *list, ImageState is, const ts2d_loss_in *loss, BwdScratch sc, cudaStream_t s)
{
    const int W = cam->width, H = cam->height;
    const int gx = (W + TS2D_TILE - 1) / TS2D_TILE, gy = (H + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int owned = (n_tiles - f->shard_rank + f->shard_world - 1) / f->shard_world;
    if (owned <= 0) return 0;
    const bool g1 = g->gamma == 1.0f;
    const uint32_t rows_cap = (uint32_t)(sc.rows_cap < 0xFFFFFFFFll ? sc.rows_cap : 0xFFFFFFFFll);
#define TS2D_BWD_ARGS                                                                                                                 \
    W, H, g->C, gx, f->shard_rank, f->shard_world, n_tiles, g->gamma, is.ranges, keys, list, gs.rec0, gs.rec1, g->background_depth, ts2d_bg_ptr(g, gs),   \
        g->background, is.final_T, is.n_contrib, loss->dL_dout_feature
#define TS2D_BWD_LAUNCH_CW(R, G, CW, ...)                                                                                               \
    do {                                                                                                                                \
        const size_t smem = CW * (size_t)BwdLayout<R>::BYTES;                                                                           \
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_render_bwd_fast<R, G, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        TS2D_CUDA_TRY(cudaFuncSetAttribute(k_render_bwd_fast<R, G, CW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));          \
        TS2D_CUDA_TRY(ts2d_launch(k_render_bwd_fast<R, G, CW>, owned * (8 / CW), 32 * CW, smem, s, TS2D_BWD_ARGS, __VA_ARGS__, sc.ei, sc.sbase, sc.rows, rows_cap)); \
    } while (0)
#define TS2D_BWD_LAUNCH(R, G, ...)                                                                                                      \
    do {                                                                                                                                \
        switch (ts2d_cta_warps()) {                                                                                                     \
        case 8: TS2D_BWD_LAUNCH_CW(R, G, 8, __VA_ARGS__); break;                                                                        \
        default: TS2D_BWD_LAUNCH_CW(R, G, 1, __VA_ARGS__); break;                                                                       \
        }                                                                                                                               \
    } while (0)
    if (f->rich_info) {
        if (g1) TS2D_BWD_LAUNCH(true, true, loss->dL_dout_depth, loss->dL_dout_normal);
        else TS2D_BWD_LAUNCH(true, false, loss->dL_dout_depth, loss->dL_dout_normal);
    } else {
        if (g1) TS2D_BWD_LAUNCH(false, true, nullptr, nullptr);
        else TS2D_BWD_LAUNCH(false, false, nullptr, nullptr);
    }
#undef TS2D_BWD_LAUNCH
#undef TS2D_BWD_LAUNCH_CW
#undef TS2D_BWD_ARGS
    return (int)cudaGetLastError();
}
