// ts2d_bwd_reduce.cu -- atomics-free, deterministic write-back of the composite backward (fast kernels).
//
// The reference sums every per-pixel gradient contribution into its P-sized arrays with fp32 atomics
// (R2D/src/backward.cu:412,482-490; R3D/src/backward.cu:365,430-451): the order of the additions, and with it the low bits of
// every gradient, changes from run to run.  The fast composite backward (ts2d_render_bwd_fast.cu, ts2d_prim3d_fast.cu) already
// reduces the 32 pixels of an 8x4 sub-tile inside its warp; what is left per triangle are a handful of partial sums -- one per
// (sub-tile, list entry) pair the forward pass blended.  Instead of RED-ing those into the triangle's accumulator line, each pair
// owns one 64 B ROW of a scratch array, and the rows are laid out in EMISSION order (depth rank, then tile of the rect, then
// sub-tile): all rows of a triangle are consecutive, so its sums are one sequential reduction in a fixed order.
//
//   k_bwd_rows_mark    per sorted list position: live sub-tile bits (coverage bit still set after the forward pass AND position
//                      below the sub-tile's own last contributor `lastw`), written back into the instance key (the backward
//                      kernel reads them there); emission index ei of the instance (from estart[id] and the tile's place in
//                      the triangle's rect) -> ei[pos], popcount -> cnt[ei]
//   k_scan_u8          exclusive scan of cnt over emission indices -> sbase (row of emission index e starts at sbase[e])
//   composite backward row of (pos, sub-tile w) = sbase[ei[pos]] + popc(live bits below w): one 16-byte store per quarter-lane
//   k_bwd_rows_reduce  one thread per depth rank: sums rows [sbase[start], sbase[end]) of its triangle into the 64 B accumulator
//                      line K9 reads (zeros for triangles without rows: no memset of the accumulators either)
#include "ts2d_sort.cuh"

namespace {

__global__ void __launch_bounds__(TS2D_BLOCK)
k_bwd_rows_mark(const int64_t *__restrict__ n_dev, int64_t cap, int gx, int shard_rank, int shard_world, uint32_t *tkey, const uint32_t *__restrict__ list,
                const uint2 *__restrict__ ranges, const uint32_t *__restrict__ lastw, const ushort4 *__restrict__ rect,
                const uint32_t *__restrict__ estart, uint32_t *__restrict__ ei_out, uint8_t *__restrict__ cnt)
{
    const int64_t R = rs_count(n_dev, cap);
    const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= R) return;
    const uint32_t key = tkey[pos];
    const uint32_t tile = key >> TS2D_MASK_BITS;
    const uint32_t rel = (uint32_t)pos - ranges[tile].x;
    const uint4 l0 = __ldg(reinterpret_cast<const uint4 *>(lastw + 8 * (size_t)tile)), l1 = __ldg(reinterpret_cast<const uint4 *>(lastw + 8 * (size_t)tile) + 1);
    uint32_t live = key & 0xFFu;
    live &= (rel < l0.x ? 1u : 0u) | (rel < l0.y ? 2u : 0u) | (rel < l0.z ? 4u : 0u) | (rel < l0.w ? 8u : 0u) | (rel < l1.x ? 16u : 0u) |
            (rel < l1.y ? 32u : 0u) | (rel < l1.z ? 64u : 0u) | (rel < l1.w ? 128u : 0u);
    if (live != (key & 0xFFu)) tkey[pos] = (key & ~0xFFu) | live;
    const uint32_t id = list[pos];
    const ushort4 rc = rect[id];
    const uint32_t ty = tile / (uint32_t)gx, tx = tile - ty * (uint32_t)gx;
    uint32_t k;
    if (shard_world == 1) {
        k = (ty - rc.y) * (uint32_t)(rc.z - rc.x) + (tx - rc.x);
    } else {  // k-th OWNED tile of the rect, row-major (the order k_emit_warp<.., true> deals them in)
        const uint32_t world = (uint32_t)shard_world, rank = (uint32_t)shard_rank;
        k = 0;
        for (uint32_t y = rc.y; y < ty; y++) {
            const uint32_t x0 = rc.x + (rank + world - (y * (uint32_t)gx + rc.x) % world) % world;
            k += x0 < rc.z ? (rc.z - x0 + world - 1) / world : 0u;
        }
        const uint32_t x0 = rc.x + (rank + world - (ty * (uint32_t)gx + rc.x) % world) % world;
        k += (tx - x0) / world;
    }
    const uint32_t e = estart[id] + k;
    ei_out[pos] = e;
    if (e < cap) cnt[e] = (uint8_t)__popc(live);
}

// exclusive scan of n (device-side) bytes -> out[0..n], out[n] = total; same chained look-back as k_scan_gather
__global__ void __launch_bounds__(RS_THREADS)
k_scan_u8(const uint8_t *__restrict__ src, uint32_t *__restrict__ out, const int64_t *__restrict__ n_dev, int64_t cap, unsigned long long *status,
          uint32_t *ticket)
{
    __shared__ uint32_t s_w[RS_WARPS];
    __shared__ uint32_t s_tile, s_excl;
    const int64_t n = rs_count(n_dev, cap);
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * SC_TILE;
    if (base >= n) return;
    uint32_t v[SC_ITEMS], sum = 0;
    const int64_t e0 = base + (int64_t)tid * SC_ITEMS;
    if (e0 + SC_ITEMS <= n) {  // 8 consecutive bytes, 8-byte aligned (cnt is 256-byte aligned, SC_ITEMS == 8)
        const uint2 w = *reinterpret_cast<const uint2 *>(src + e0);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            v[i] = (w.x >> (8 * i)) & 0xFFu;
            v[4 + i] = (w.y >> (8 * i)) & 0xFFu;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) v[i] = (e0 + i < n) ? src[e0 + i] : 0u;
    }
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) sum += v[i];
    uint32_t total;
    uint32_t excl = block_excl_scan256(sum, s_w, &total);
    if (tid == 0) {
        const unsigned long long tagP = 1ull << 32, tagI = 2ull << 32;
        uint32_t look = 0;
        if (tile == 0) {
            st_relaxed_u64(status, tagI | total);
        } else {
            st_relaxed_u64(status + tile, tagP | total);
            int64_t j = (int64_t)tile - 1;
            while (true) {
                const unsigned long long w = ld_relaxed_u64(status + j);
                const unsigned long long tg = w & 0xFFFFFFFF00000000ull;
                if (tg == tagI) { look += (uint32_t)w; break; }
                if (tg == tagP) { look += (uint32_t)w; j--; continue; }
                __nanosleep(20);
            }
            st_relaxed_u64(status + tile, tagI | (unsigned long long)(look + total));
        }
        s_excl = look;
        if (base + SC_TILE >= n) out[n] = look + total;
    }
    __syncthreads();
    excl += s_excl;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
        if (e0 + i < n) out[e0 + i] = excl;
        excl += v[i];
    }
}

// one thread per depth rank; rows of adjacent ranks are adjacent in memory
__global__ void __launch_bounds__(TS2D_BLOCK)
k_bwd_rows_reduce(int P, const int64_t *__restrict__ n_dev, int64_t cap, int64_t rows_cap, const uint32_t *__restrict__ order, const uint32_t *__restrict__ tiles,
                  const uint32_t *__restrict__ offs, const uint32_t *__restrict__ sbase, const float4 *__restrict__ rows, float4 *__restrict__ gacc)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P) return;
    const int64_t R = rs_count(n_dev, cap);
    const uint32_t id = order[r];
    const uint32_t n = tiles[id];
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    if (n) {
        int64_t end = offs[r], start = end - n;
        end = end < R ? end : R;
        start = start < R ? start : R;
        int64_t b0 = sbase[start], b1 = sbase[end];
        b1 = b1 < rows_cap ? b1 : rows_cap;
        for (int64_t s = b0; s < b1; s++) {
            const float4 *q = rows + 4 * s;
            const float4 x0 = __ldg(q), x1 = __ldg(q + 1), x2 = __ldg(q + 2), x3 = __ldg(q + 3);
            a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
            a1.x += x1.x; a1.y += x1.y; a1.z += x1.z; a1.w += x1.w;
            a2.x += x2.x; a2.y += x2.y; a2.z += x2.z; a2.w += x2.w;
            a3.x += x3.x; a3.y += x3.y; a3.z += x3.z; a3.w += x3.w;
        }
    }
    float4 *g = gacc + 4 * (size_t)id;
    g[0] = a0;
    g[1] = a1;
    g[2] = a2;
    g[3] = a3;
}

}  // namespace

int ts2d_launch_bwd_rows_prepare(const ts2d_camera *cam, const ts2d_flags *f, GeomState gs, BinState bs, ImageState is, BwdScratch sc, cudaStream_t s)
{
    const int gx = (cam->width + TS2D_TILE - 1) / TS2D_TILE, gy = (cam->height + TS2D_TILE - 1) / TS2D_TILE;
    const int sbuf = ts2d_sorted_buf(gx * gy);
    const int64_t *n_dev = &gs.hdr->num_rendered;
    if (bs.cap <= 0) return 0;
    TS2D_CUDA_TRY(cudaMemsetAsync(sc.sstatus, 0, sc.sstatus_bytes, s));
    TS2D_CUDA_TRY(cudaMemsetAsync(&gs.hdr->render.tickets[TS2D_TICKET_BWD_SCAN], 0, sizeof(uint32_t), s));
    k_bwd_rows_mark<<<(unsigned)((bs.cap + TS2D_BLOCK - 1) / TS2D_BLOCK), TS2D_BLOCK, 0, s>>>(n_dev, bs.cap, gx, f->shard_rank, f->shard_world, bs.tkey[sbuf],
                                                                                             bs.tval[sbuf], is.ranges, is.lastw, gs.rect, gs.estart,
                                                                                             sc.ei, sc.cnt);
    k_scan_u8<<<(unsigned)sc_tiles(bs.cap), RS_THREADS, 0, s>>>(sc.cnt, sc.sbase, n_dev, bs.cap, sc.sstatus, &gs.hdr->render.tickets[TS2D_TICKET_BWD_SCAN]);
    return (int)cudaGetLastError();
}

int ts2d_launch_bwd_rows_reduce(int32_t P, GeomState gs, BwdScratch sc, cudaStream_t s)
{
    if (P <= 0) return 0;
    k_bwd_rows_reduce<<<(P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, 0, s>>>(P, &gs.hdr->num_rendered, sc.cap, sc.rows_cap, gs.ids2, gs.tiles, gs.offs,
                                                                            sc.sbase, sc.rows, reinterpret_cast<float4 *>(sc.gacc));
    return (int)cudaGetLastError();
}
