// ts2d_binning.cu -- ordering, key emission, tile binning, tile ranges (K2-K6).
//
// Replaces R2D/src/rasterizer.cu:37-99,186-231 (InclusiveSum, duplicateWithKeys, 64-bit
// DeviceRadixSort::SortPairs over R instances, identifyTileRanges).
//
// The reference sorts R (tile<<32 | depth_bits) keys with a 45-bit LSD radix sort: 6 passes over
// 12 B/instance.  Here the same total order -- (tile, depth bits, triangle id) -- is produced in
// two cheaper steps:
//   1. sort the P triangles once by their 32-bit depth pattern (stable in triangle id);
//   2. emit instances in that depth-rank order and bin them with a STABLE radix sort over only the
//      ceil(log2(tiles)) tile-id bits (2 passes over 8 B/instance at 1080p).
// Stability of step 2 preserves (depth, id) order inside every tile, so the resulting list is
// bit-identical to the reference's point_list (ts2d_export_binning rebuilds the 64-bit keys).
// The 8 key bits below the tile id carry each instance's sub-tile coverage mask through the sort.
//
// Every kernel here is hand-written (ts2d_sort.cuh holds the radix passes and the scan) and takes the number of instances R
// from DEVICE memory: nothing in this file needs R on the host, so a whole forward pass can be enqueued without a
// synchronisation (the reference blocks on a cudaMemcpy of R, rasterizer.cu:190-193).
#include "ts2d_prim3d.cuh"
#include "ts2d_sort.cuh"

namespace {

// Instance key = (tile id << 8) | sub-tile coverage mask.  The tile sort orders on the tile bits only and carries the mask
// along for free; the fast composite kernels read it instead of re-deriving coverage from the raster record (once per
// instance here instead of once per instance per pass there).  With MASKS == false (mirror kernels) the mask byte is 0xFF.
// MASKS: 0 = none (mirror kernels), 1 = 2D primitive (subtile_mask, ts2d_fast.cuh), 2 = 3D primitive (subtile_mask3d, ts2d_prim3d.cuh)
struct EmitCam {
    int W, H;
    float tfx, tfy;
};
template <int MASKS>
__device__ __forceinline__ uint32_t instance_key(uint32_t tile, int gx, uint32_t id, const float4 *__restrict__ rec0, const GammaK gk, const EmitCam cam)
{
    if (MASKS == 0) return (tile << TS2D_MASK_BITS) | 0xFFu;
    const float4 r0 = __ldg(rec0 + 3 * (size_t)id), r1 = __ldg(rec0 + 3 * (size_t)id + 1);
    const uint32_t ty = tile / (uint32_t)gx, tx = tile - ty * (uint32_t)gx;
    if (MASKS == 2) {
        const Tri3 t = unpack3(r0, r1, __ldg(rec0 + 3 * (size_t)id + 2));
        return (tile << TS2D_MASK_BITS) |
               subtile_mask3d(t, (float)(tx * TS2D_TILE), (float)(ty * TS2D_TILE), cam.W, cam.H, cam.tfx, cam.tfy, gk.gamma, gk.is_one);
    }
    return (tile << TS2D_MASK_BITS) | subtile_mask(r0, r1, r1.z, (float)(tx * TS2D_TILE), (float)(ty * TS2D_TILE), gk);
}

// Emission, one warp per 32 consecutive depth ranks: the warp's instances [start of rank r0, end of rank r0+31) are
// dealt to lanes round-robin, each lane finding its (triangle, k-th tile of the rect, row-major) by a 5-step search over the
// 32 scan values held in the warp.  Same output as rasterizer.cu:63-74 in depth order, with coalesced stores and no
// divergence on the rect size.  SHARDED: only the tiles this rank owns (tile % world == rank) count and are written; the
// k-th owned tile is found row by row (each row holds every world-th tile starting at a closed-form first column).
// `cap`: instances the output arrays hold; if the frame has more (R > cap: the caller sized the binning state too small)
// the excess is dropped here -- the caller finds R > cap in the frame's counters and repeats the render.
template <int MASKS, bool SHARDED>
__global__ void __launch_bounds__(TS2D_BLOCK)
k_emit_warp(int P, int gx, float gamma, EmitCam cam, int shard_rank, int shard_world, const uint32_t *__restrict__ order, const uint32_t *__restrict__ tiles,
            const ushort4 *__restrict__ rect, const uint32_t *__restrict__ offs, const float4 *__restrict__ rec0, uint32_t cap, uint32_t *__restrict__ tkey,
            uint32_t *__restrict__ tval, uint4 *__restrict__ binrec, uint2 *__restrict__ ranges)
{
    ts2d_grid_chain();
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x);
    const GammaK gk = make_gamma(gamma);
    uint32_t id = 0, n = 0, end = 0;
    uint32_t rc_lo = 0, rc_hi = 0;  // {min.x | min.y << 16}, {max.x | max.y << 16}
    if (r < P) {
        id = order[r];
        n = tiles[id];
        end = offs[r];
        const ushort4 rc = rect[id];
        rc_lo = (uint32_t)rc.x | ((uint32_t)rc.y << 16);
        rc_hi = (uint32_t)rc.z | ((uint32_t)rc.w << 16);
        if (n) binrec[id] = make_uint4(rc_lo, rc_hi, end - n, n);  // one record per triangle for the backward's row marking
    } else {
        end = offs[P - 1];
    }
    const uint32_t start = end - n;
    const uint32_t w_start = __shfl_sync(0xffffffffu, start, 0), w_end = min(__shfl_sync(0xffffffffu, end, 31), cap);
    for (uint32_t base = w_start; base < w_end; base += 32) {
        const uint32_t i = base + lane;
        // t = number of lanes whose end <= i (ends are non-decreasing)
        int t = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const uint32_t e = __shfl_sync(0xffffffffu, end, t + step - 1);
            if (e <= i) t += step;
        }
        t = min(t, 31);  // lanes past w_end
        const uint32_t t_start = __shfl_sync(0xffffffffu, start, t), t_id = __shfl_sync(0xffffffffu, id, t);
        const uint32_t lo = __shfl_sync(0xffffffffu, rc_lo, t), hi = __shfl_sync(0xffffffffu, rc_hi, t);
        if (i < w_end) {
            uint32_t k = i - t_start;
            const uint32_t x_lo = lo & 0xffffu, x_hi = hi & 0xffffu;
            uint32_t tile;
            if (!SHARDED) {
                const uint32_t w = x_hi - x_lo;
                const uint32_t dy = k / w, dx = k - dy * w;
                tile = ((lo >> 16) + dy) * (uint32_t)gx + x_lo + dx;
            } else {
                const uint32_t world = (uint32_t)shard_world, rank = (uint32_t)shard_rank;
                uint32_t y = lo >> 16;
                for (;; y++) {  // the instance exists, so the loop ends inside the rect
                    const uint32_t first = y * (uint32_t)gx + x_lo;                   // first tile of the row
                    const uint32_t x0 = x_lo + (rank + world - first % world) % world;  // first owned column
                    const uint32_t cnt = x0 < x_hi ? (x_hi - x0 + world - 1) / world : 0u;
                    if (k < cnt) {
                        tile = y * (uint32_t)gx + x0 + k * world;
                        break;
                    }
                    k -= cnt;
                }
            }
            tkey[i] = instance_key<MASKS>(tile, gx, t_id, rec0, gk, cam);
            tval[i] = t_id;
            atomicAdd(&ranges[tile].y, 1u);  // instances per tile (the array is zero on entry): k_tile_tables turns the counts into ranges
        }
    }
}

// Tile ranges and the digit histograms of the tile sort from the per-tile instance counts the emission left in ranges[].y -- one
// block, before the sort runs (the ranges depend on the counts only): replaces identifyTileRanges (rasterizer.cu:79-99, a pass over
// the R sorted keys) and a histogram pass over the R unsorted keys.  Tiles without instances keep {0, 0} like the reference's
// zero-initialised array.  hist[d][b] = instances whose d-th 8-bit digit of the tile id is b (all 256 bins are written).
__global__ void __launch_bounds__(1024) k_tile_tables(int n_tiles, int np, uint2 *__restrict__ ranges, uint32_t *__restrict__ hist0, uint32_t *__restrict__ hist1,
                                                     uint32_t *__restrict__ hist2)
{
    __shared__ uint32_t s_part[1024];
    __shared__ uint32_t s_h[3][RS_BINS];
    ts2d_grid_chain();
    const int tid = threadIdx.x;
    for (int i = tid; i < 3 * RS_BINS; i += 1024) (&s_h[0][0])[i] = 0;
    const int per = (n_tiles + 1023) / 1024;
    const int t0 = tid * per, t1 = min(n_tiles, t0 + per);
    uint32_t sum = 0;
    for (int t = t0; t < t1; t++) sum += ranges[t].y;
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partial sums
        const uint32_t v = tid >= o ? s_part[tid - o] : 0u;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint32_t run = s_part[tid] - sum;
    for (int t = t0; t < t1; t++) {
        const uint32_t c = ranges[t].y;
        if (c) {
            ranges[t] = make_uint2(run, run + c);
            for (int d = 0; d < np; d++) atomicAdd(&s_h[d][((uint32_t)t >> (8 * d)) & 0xFFu], c);
            run += c;
        }
    }
    __syncthreads();
    if (tid < RS_BINS) {
        hist0[tid] = s_h[0][tid];
        if (np > 1) hist1[tid] = s_h[1][tid];
        if (np > 2) hist2[tid] = s_h[2][tid];
    }
}

int hist_blocks(int64_t n) { return (int)((n + 4095) / 4096 < 592 ? ((n + 4095) / 4096 > 0 ? (n + 4095) / 4096 : 1) : 592); }

}  // namespace

size_t ts2d_sort_status_bytes(int64_t n_cap) { return ts2d_align_up(rs_status_bytes(n_cap), 256); }
size_t ts2d_scan_status_bytes(int64_t n_cap) { return ts2d_align_up((size_t)sc_tiles(n_cap > 0 ? n_cap : 1) * sizeof(unsigned long long), 256); }  // block sums of ts2d_scan (u32; sized generously)

// K2/K3: depth order of the triangles (stable LSD radix sort of the 32-bit depth patterns, triangle id as the payload), scan of
// tiles-touched in that order (-> offs, and R in the header).
int ts2d_launch_order_and_scan(int32_t P, GeomState gs, int64_t *R_host, cudaStream_t s)
{
    RadixHistArgs h = {};
    h.keys = gs.dkey;
    h.n_dev = nullptr;
    h.n_cap = P;
    h.ndigits = 4;
    for (int d = 0; d < 4; d++) {
        h.shift[d] = 8 * d;
        h.mask[d] = 0xFFu;
        h.hist[d] = gs.hdr->hist[d];
    }
    TS2D_CUDA_TRY(ts2d_launch(k_radix_hist, hist_blocks(P), RS_THREADS, 0, s, h));
    // dkey (kept: the export calls and the 64-bit key reconstruction read it) -> (tmpk, ids) -> (dkey2, ids2) -> (tmpk, ids) -> (dkey2, ids2)
    const uint32_t *kin[4] = {gs.dkey, gs.tmpk, gs.dkey2, gs.tmpk}, *vin[4] = {nullptr, gs.ids, gs.ids2, gs.ids};
    uint32_t *kout[4] = {gs.tmpk, gs.dkey2, gs.tmpk, gs.dkey2}, *vout[4] = {gs.ids, gs.ids2, gs.ids, gs.ids2};
    for (int d = 0; d < 4; d++) {
        RadixPassArgs a = {};
        a.kin = kin[d];
        a.kout = kout[d];
        a.vin = vin[d];
        a.vout = vout[d];
        a.hist = gs.hdr->hist[d];
        a.status = gs.status;
        a.ticket = &gs.hdr->tickets[TS2D_TICKET_DEPTH0 + d];
        a.n_dev = nullptr;
        a.n_cap = P;
        a.shift = 8 * d;
        a.mask = 0xFFu;
        a.pass_uid = (uint32_t)(d + 1);
        TS2D_CUDA_TRY(ts2d_launch_chained(k_radix_pass, (unsigned)rs_tiles(P), RS_THREADS, 0, s, a));
    }
    TS2D_CUDA_TRY((ts2d_scan<LoadGatherU32, true>(LoadGatherU32{gs.ids2, gs.tiles}, nullptr, P, reinterpret_cast<uint32_t *>(gs.sstatus), gs.offs,
                                                   &gs.hdr->num_rendered, false, s)));
    if (R_host) {
        int64_t R = 0;
        TS2D_CUDA_TRY(cudaMemcpyAsync(&R, &gs.hdr->num_rendered, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        TS2D_CUDA_TRY(cudaStreamSynchronize(s));
        *R_host = R;
    }
    return 0;
}

// K4-K6.
int ts2d_launch_binning(const ts2d_camera *cam, const ts2d_geometry *g, const ts2d_flags *f, int64_t R_host, GeomState gs, BinState bs, ImageState is,
                        bool pre_cleared, cudaStream_t s)
{
    const int gx = (cam->width + TS2D_TILE - 1) / TS2D_TILE, gy = (cam->height + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    const int P = g->P;
    if (!pre_cleared) TS2D_CUDA_TRY(cudaMemsetAsync(is.ranges, 0, sizeof(uint2) * (size_t)n_tiles, s));
    if (R_host == 0) return 0;
    const int64_t n_launch = R_host > 0 ? (R_host < bs.cap ? R_host : bs.cap) : bs.cap;  // what the grids are sized for
    if (n_launch <= 0) return 0;
    if (!pre_cleared) TS2D_CUDA_TRY(cudaMemsetAsync(bs.status, 0, ts2d_sort_status_bytes(n_launch), s));
    const int masks = !ts2d_use_fast(g, f) ? 0 : (f->primitive == TS2D_PRIMITIVE_3D ? 2 : 1);
    const int blocks = (P + TS2D_BLOCK - 1) / TS2D_BLOCK;
    const EmitCam ec = {cam->width, cam->height, cam->tan_fovx, cam->tan_fovy};
    const uint32_t cap32 = (uint32_t)(bs.cap < 0xFFFFFFFFll ? bs.cap : 0xFFFFFFFFll);
#define TS2D_EMIT_ARGS P, gx, g->gamma, ec, f->shard_rank, f->shard_world, gs.ids2, gs.tiles, gs.rect, gs.offs, gs.rec0, cap32, bs.tkey[0], bs.tval[0], gs.binrec, is.ranges
    if (f->shard_world > 1) {
        if (masks == 2) TS2D_CUDA_TRY(ts2d_launch(k_emit_warp<2, true>, blocks, TS2D_BLOCK, 0, s, TS2D_EMIT_ARGS));
        else if (masks == 1) TS2D_CUDA_TRY(ts2d_launch(k_emit_warp<1, true>, blocks, TS2D_BLOCK, 0, s, TS2D_EMIT_ARGS));
        else TS2D_CUDA_TRY(ts2d_launch(k_emit_warp<0, true>, blocks, TS2D_BLOCK, 0, s, TS2D_EMIT_ARGS));
    } else {
        if (masks == 2) TS2D_CUDA_TRY(ts2d_launch(k_emit_warp<2, false>, blocks, TS2D_BLOCK, 0, s, TS2D_EMIT_ARGS));
        else if (masks == 1) TS2D_CUDA_TRY(ts2d_launch(k_emit_warp<1, false>, blocks, TS2D_BLOCK, 0, s, TS2D_EMIT_ARGS));
        else TS2D_CUDA_TRY(ts2d_launch(k_emit_warp<0, false>, blocks, TS2D_BLOCK, 0, s, TS2D_EMIT_ARGS));
    }
#undef TS2D_EMIT_ARGS
    TS2D_CUDA_TRY(cudaGetLastError());
    // stable sort on the tile bits: digits of 8 bits from bit TS2D_MASK_BITS up (the last one narrower)
    const int tb = ts2d_tile_bits(n_tiles), np = ts2d_tile_sort_passes(n_tiles);
    const int64_t *n_dev = &gs.hdr->num_rendered;
    static_assert(TS2D_MASK_BITS == 8, "the digits of the tile sort are the bytes of the tile id");
    TS2D_CUDA_TRY(ts2d_launch_chained(k_tile_tables, 1, 1024, 0, s, n_tiles, np, is.ranges, gs.hdr->render.hist[0], gs.hdr->render.hist[1], gs.hdr->render.hist[2]));
    for (int d = 0; d < np; d++) {
        RadixPassArgs a = {};
        a.kin = bs.tkey[d & 1];
        a.kout = bs.tkey[(d & 1) ^ 1];
        a.vin = bs.tval[d & 1];
        a.vout = bs.tval[(d & 1) ^ 1];
        a.hist = gs.hdr->render.hist[d];
        a.status = bs.status;
        a.ticket = &gs.hdr->render.tickets[TS2D_TICKET_TILE0 + d];
        a.n_dev = n_dev;
        a.n_cap = bs.cap;
        a.shift = TS2D_MASK_BITS + 8 * d;
        a.mask = (d == np - 1) ? ((1u << (tb - 8 * d)) - 1u) : 0xFFu;
        a.pass_uid = (uint32_t)(8 + d);
        TS2D_CUDA_TRY(ts2d_launch_chained(k_radix_pass, (unsigned)rs_tiles(n_launch), RS_THREADS, 0, s, a));
    }
    return (int)cudaGetLastError();
}
