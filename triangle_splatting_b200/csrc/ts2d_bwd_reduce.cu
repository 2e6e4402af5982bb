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
//   k_bwd_rows_reduce  64 depth ranks per block: streams their (contiguous) rows through shared memory, 4 lanes per triangle sum its
//                      rows in row order into the 64 B accumulator line K9 reads (zeros for triangles without rows: no memset either)
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
    if (e0 + SC_ITEMS <= n) {  // 16 consecutive bytes, 16-byte aligned (cnt is 256-byte aligned, SC_ITEMS == 16)
        const uint4 w = *reinterpret_cast<const uint4 *>(src + e0);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = (ww[i >> 2] >> (8 * (i & 3))) & 0xFFu;
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
            look = lookback_sum(status, (int64_t)tile - 1, 1, tagP, tagI);
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

// Row reduction.  A block owns RR_RANKS consecutive depth ranks; their rows are ONE contiguous span of the row array (rows are laid
// out in emission order and the instance ranges of consecutive ranks are consecutive).  The block streams the span through shared
// memory in chunks with fully coalesced 16-byte loads; 4 lanes per rank (one per float4 column of a row) then add the rows of their
// rank that sit in the chunk, in row order: every accumulator sees its addends in one fixed order, whatever the launch looks like.
constexpr int RR_RANKS = 64;    // ranks per block = quads per block (256 threads)
constexpr int RR_CHUNK = 256;   // rows staged per step (16 KB)

__device__ __forceinline__ float4 ldg_stream128(const float4 *p)
{
    float4 v;  // read once: do not let the rows displace the raster records from L1 / L2
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(4 * RR_RANKS)
k_bwd_rows_reduce(int P, const int64_t *__restrict__ n_dev, int64_t cap, int64_t rows_cap, const uint32_t *__restrict__ order, const uint32_t *__restrict__ tiles,
                  const uint32_t *__restrict__ offs, const uint32_t *__restrict__ sbase, const float4 *__restrict__ rows, float4 *__restrict__ gacc)
{
    __shared__ float4 s_rows[RR_CHUNK * 4];
    __shared__ uint32_t s_b[RR_RANKS + 1];  // first row of each rank of the block; [RR_RANKS] = end of the span
    __shared__ uint32_t s_id[RR_RANKS];
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * RR_RANKS;
    const int64_t R = rs_count(n_dev, cap);
    if (tid <= RR_RANKS) {
        // instance ranges of consecutive ranks are consecutive: rank r starts at offs[r - 1] (inclusive scan) and ends at offs[r]
        const int r = r0 + tid;
        int64_t e = 0;
        if (r > 0) e = offs[(r - 1 < P ? r - 1 : P - 1)];
        e = e < R ? e : R;
        int64_t b = R > 0 ? sbase[e] : 0;  // no instances: the scan wrote nothing
        s_b[tid] = (uint32_t)(b < rows_cap ? b : rows_cap);
        if (tid < RR_RANKS) s_id[tid] = r < P ? order[r] : 0xFFFFFFFFu;
    }
    __syncthreads();
    const int q = tid >> 2, c = tid & 3;
    const uint32_t lo = s_b[q], hi = s_b[q + 1];
    const uint32_t span0 = s_b[0], span1 = s_b[RR_RANKS];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t c0 = span0; c0 < span1; c0 += RR_CHUNK) {
        const uint32_t n4 = 4u * min((uint32_t)RR_CHUNK, span1 - c0);
        const float4 *src = rows + 4 * (size_t)c0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i = (uint32_t)tid + 256u * k;
            if (i < n4) s_rows[i] = ldg_stream128(src + i);
        }
        __syncthreads();
        const uint32_t a = max(lo, c0), b = min(hi, c0 + (uint32_t)RR_CHUNK);
        for (uint32_t s = a; s < b; s++) {
            const float4 x = s_rows[4 * (s - c0) + c];
            acc.x += x.x;
            acc.y += x.y;
            acc.z += x.z;
            acc.w += x.w;
        }
        __syncthreads();
    }
    const uint32_t id = s_id[q];
    if (id != 0xFFFFFFFFu) gacc[4 * (size_t)id + c] = acc;  // zeros for triangles without rows: the accumulators need no memset
    (void)tiles;
}

__global__ void __launch_bounds__(TS2D_BLOCK) k_contrib_finish(int P, const unsigned long long *__restrict__ csum64, float *__restrict__ contrib_sum)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) contrib_sum[i] = (float)csum64[i] * (1.0f / 4294967296.0f);
}

}  // namespace

// contrib_sum (forward.cu:323) is summed over (sub-tile, entry) pairs as 2^-32 fixed-point integers by the fast forward kernels:
// the result does not depend on the order of the atomics.  This turns the sums into the fp32 output.
int ts2d_launch_contrib_finish(int32_t P, const unsigned long long *csum64, float *contrib_sum, cudaStream_t s)
{
    if (P <= 0) return 0;
    k_contrib_finish<<<(P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, 0, s>>>(P, csum64, contrib_sum);
    return (int)cudaGetLastError();
}

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
    k_bwd_rows_reduce<<<(P + RR_RANKS - 1) / RR_RANKS, 4 * RR_RANKS, 0, s>>>(P, &gs.hdr->num_rendered, sc.cap, sc.rows_cap, gs.ids2, gs.tiles, gs.offs,
                                                                            sc.sbase, sc.rows, reinterpret_cast<float4 *>(sc.gacc));
    return (int)cudaGetLastError();
}
