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
//                      kernel reads them there); emission index ei of the instance (from binrec[id] and the tile's place in
//                      the triangle's rect) -> ei[pos], popcount -> cnt[ei]
//   ts2d_scan          exclusive scan of cnt over emission indices -> sbase (row of emission index e starts at sbase[e])
//   composite backward row of (pos, sub-tile w) = sbase[ei[pos]] + popc(live bits below w): one 16-byte store per quarter-lane
//   k_bwd_rows_reduce  64 depth ranks per block: streams their (contiguous) rows through shared memory, 4 lanes per triangle sum its
//                      rows in row order into the 64 B accumulator line K9 reads (zeros for triangles without rows: no memset either)
#include "ts2d_sort.cuh"

namespace {

// MK positions per thread.  The kernel is a chain of gathers (key -> tile -> ranges / lastw, list -> the triangle's 16-byte bin record) whose limit is the
// L2 transaction rate (ncu: l1tex 77 %, lts 59 % of peak), not latency: 4 positions per thread cost 92 registers and two thirds of
// the occupancy for nothing (129 us against 118 us with one position and 31 registers).
constexpr int MK = 1;
__global__ void __launch_bounds__(TS2D_BLOCK)
k_bwd_rows_mark(const int64_t *__restrict__ n_dev, int64_t cap, int gx, int shard_rank, int shard_world, uint32_t *tkey, const uint32_t *__restrict__ list,
                const uint2 *__restrict__ ranges, const uint32_t *__restrict__ lastw, const uint4 *__restrict__ binrec,
                uint32_t *__restrict__ ei_out, uint8_t *__restrict__ cnt)
{
    ts2d_grid_chain();
    const int64_t R = rs_count(n_dev, cap);
    const int64_t p0 = (int64_t)blockIdx.x * (blockDim.x * MK) + threadIdx.x;
    uint32_t key[MK], id[MK];
    bool ok[MK];
#pragma unroll
    for (int u = 0; u < MK; u++) {
        const int64_t pos = p0 + (int64_t)u * blockDim.x;
        ok[u] = pos < R;
        key[u] = ok[u] ? tkey[pos] : 0u;
        id[u] = ok[u] ? list[pos] : 0u;
    }
    uint32_t first[MK], est[MK];
    uint4 l0[MK], l1[MK];
    uint4 br[MK];
#pragma unroll
    for (int u = 0; u < MK; u++) {
        const uint32_t tile = key[u] >> TS2D_MASK_BITS;
        first[u] = ok[u] ? ranges[tile].x : 0u;
        l0[u] = ok[u] ? __ldg(reinterpret_cast<const uint4 *>(lastw + 8 * (size_t)tile)) : make_uint4(0, 0, 0, 0);
        l1[u] = ok[u] ? __ldg(reinterpret_cast<const uint4 *>(lastw + 8 * (size_t)tile) + 1) : make_uint4(0, 0, 0, 0);
        br[u] = ok[u] ? __ldg(binrec + id[u]) : make_uint4(0u, 0x00010001u, 0u, 0u);
        est[u] = br[u].z;
    }
#pragma unroll
    for (int u = 0; u < MK; u++) {
        if (!ok[u]) continue;
        const int64_t pos = p0 + (int64_t)u * blockDim.x;
        const uint32_t tile = key[u] >> TS2D_MASK_BITS;
        const uint32_t rel = (uint32_t)pos - first[u];
        uint32_t live = key[u] & 0xFFu;
        live &= (rel < l0[u].x ? 1u : 0u) | (rel < l0[u].y ? 2u : 0u) | (rel < l0[u].z ? 4u : 0u) | (rel < l0[u].w ? 8u : 0u) | (rel < l1[u].x ? 16u : 0u) |
                (rel < l1[u].y ? 32u : 0u) | (rel < l1[u].z ? 64u : 0u) | (rel < l1[u].w ? 128u : 0u);
        if (live != (key[u] & 0xFFu)) tkey[pos] = (key[u] & ~0xFFu) | live;
        const uint32_t ty = tile / (uint32_t)gx, tx = tile - ty * (uint32_t)gx;
        const ushort4 r = make_ushort4((unsigned short)(br[u].x & 0xffffu), (unsigned short)(br[u].x >> 16), (unsigned short)(br[u].y & 0xffffu),
                                       (unsigned short)(br[u].y >> 16));
        uint32_t k;
        if (shard_world == 1) {
            k = (ty - r.y) * (uint32_t)(r.z - r.x) + (tx - r.x);
        } else {  // k-th OWNED tile of the rect, row-major (the order k_emit_warp<.., true> deals them in)
            const uint32_t world = (uint32_t)shard_world, rank = (uint32_t)shard_rank;
            k = 0;
            for (uint32_t y = r.y; y < ty; y++) {
                const uint32_t x0 = r.x + (rank + world - (y * (uint32_t)gx + r.x) % world) % world;
                k += x0 < r.z ? (r.z - x0 + world - 1) / world : 0u;
            }
            const uint32_t x0 = r.x + (rank + world - (ty * (uint32_t)gx + r.x) % world) % world;
            k += (tx - x0) / world;
        }
        const uint32_t e = est[u] + k;
        ei_out[pos] = e;
        if (e < cap) cnt[e] = (uint8_t)__popc(live);
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
    ts2d_grid_chain();
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
    ts2d_grid_chain();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) contrib_sum[i] = (float)csum64[i] * (1.0f / 4294967296.0f);
}

}  // namespace

// contrib_sum (forward.cu:323) is summed over (sub-tile, entry) pairs as 2^-32 fixed-point integers by the fast forward kernels:
// the result does not depend on the order of the atomics.  This turns the sums into the fp32 output.
int ts2d_launch_contrib_finish(int32_t P, const unsigned long long *csum64, float *contrib_sum, cudaStream_t s)
{
    if (P <= 0) return 0;
    return (int)ts2d_launch(k_contrib_finish, (P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, 0, s, P, csum64, contrib_sum);
}

int ts2d_launch_bwd_rows_prepare(const ts2d_camera *cam, const ts2d_flags *f, GeomState gs, BinState bs, ImageState is, BwdScratch sc, cudaStream_t s)
{
    const int gx = (cam->width + TS2D_TILE - 1) / TS2D_TILE, gy = (cam->height + TS2D_TILE - 1) / TS2D_TILE;
    const int sbuf = ts2d_sorted_buf(gx * gy);
    const int64_t *n_dev = &gs.hdr->num_rendered;
    if (bs.cap <= 0) return 0;
    TS2D_CUDA_TRY(ts2d_launch(k_bwd_rows_mark, (unsigned)((bs.cap + TS2D_BLOCK * MK - 1) / (TS2D_BLOCK * MK)), TS2D_BLOCK, 0, s, n_dev, bs.cap, gx, f->shard_rank,
                              f->shard_world, bs.tkey[sbuf], bs.tval[sbuf], is.ranges, is.lastw, gs.binrec, sc.ei, sc.cnt));
    // sbase = exclusive scan of cnt over emission indices, sbase[R] = number of rows
    return (int)ts2d_scan<LoadU8, false>(LoadU8{sc.cnt}, n_dev, bs.cap, reinterpret_cast<uint32_t *>(sc.sstatus), sc.sbase, nullptr, true, s);
}

int ts2d_launch_bwd_rows_reduce(int32_t P, GeomState gs, BwdScratch sc, cudaStream_t s)
{
    if (P <= 0) return 0;
    return (int)ts2d_launch(k_bwd_rows_reduce, (P + RR_RANKS - 1) / RR_RANKS, 4 * RR_RANKS, 0, s, P, &gs.hdr->num_rendered, sc.cap, sc.rows_cap, gs.ids2,
                            gs.tiles, gs.offs, sc.sbase, sc.rows, reinterpret_cast<float4 *>(sc.gacc));
}
