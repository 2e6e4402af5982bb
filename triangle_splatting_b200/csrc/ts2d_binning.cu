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
//
// The radix passes and the scan are CUB device primitives from the CUDA toolkit (the reference uses
// the same library for its sort/scan, rasterizer.cu:186,211); everything else is hand-written.
#include <cub/cub.cuh>

#include "ts2d_common.cuh"

namespace {

struct GatherTiles {
    const uint32_t *tiles;
    __host__ __device__ __forceinline__ uint32_t operator()(uint32_t id) const { return tiles[id]; }
};

// One thread per depth rank: write the (owned) tiles of the triangle's rect, row-major, like
// rasterizer.cu:63-74 but in depth order and with the tile id alone as key.
__global__ void __launch_bounds__(TS2D_BLOCK)
k_emit(int P, int gx, int shard_rank, int shard_world, const uint32_t *__restrict__ order, const uint32_t *__restrict__ tiles,
       const ushort4 *__restrict__ rect, const uint32_t *__restrict__ offs, uint32_t *__restrict__ tkey, uint32_t *__restrict__ tval)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P) return;
    const uint32_t id = order[r];
    if (tiles[id] == 0) return;
    uint32_t off = (r == 0) ? 0u : offs[r - 1];
    const ushort4 rc = rect[id];
    for (uint32_t y = rc.y; y < rc.w; y++)
        for (uint32_t x = rc.x; x < rc.z; x++) {
            const uint32_t t = y * (uint32_t)gx + x;
            if (shard_world > 1 && (t % (uint32_t)shard_world) != (uint32_t)shard_rank) continue;
            tkey[off] = t;
            tval[off] = id;
            off++;
        }
}

// rasterizer.cu:79-99 on 32-bit tile keys.
__global__ void __launch_bounds__(TS2D_BLOCK) k_ranges(int64_t R, const uint32_t *__restrict__ tkey, uint2 *__restrict__ ranges)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t cur = tkey[i];
    if (i == 0)
        ranges[cur].x = 0;
    else {
        const uint32_t prev = tkey[i - 1];
        if (cur != prev) {
            ranges[prev].y = (uint32_t)i;
            ranges[cur].x = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[cur].y = (uint32_t)R;
}

__global__ void k_set_header(GeomHeader *h, const uint32_t *offs, int P)
{
    h->num_rendered = (P > 0) ? (int64_t)offs[P - 1] : 0;
}

}  // namespace

size_t ts2d_depth_sort_temp_bytes(int32_t P)
{
    size_t sort_b = 0, scan_b = 0;
    const int n = P > 0 ? P : 1;
    uint32_t *np = nullptr;
    cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const uint32_t *)np, np, (const uint32_t *)np, np, n, 0, 32);
    cub::TransformInputIterator<uint32_t, GatherTiles, const uint32_t *> it(np, GatherTiles{np});
    cudaError_t e2 = cub::DeviceScan::InclusiveSum(nullptr, scan_b, it, np, n);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        // No device (CPU-only box): CUB cannot size its temp storage.  Bound it: an alternate key+value buffer plus histograms.
        cudaGetLastError();
        return ts2d_align_up((size_t)n * 8 + (4u << 20), 256);
    }
    return ts2d_align_up(sort_b > scan_b ? sort_b : scan_b, 256);
}

size_t ts2d_tile_sort_temp_bytes(int64_t R)
{
    size_t sort_b = 0;
    const int64_t n = R > 0 ? R : 1;
    uint32_t *np = nullptr;
    cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const uint32_t *)np, np, (const uint32_t *)np, np, n, 0, 32);
    if (e1 != cudaSuccess) {
        cudaGetLastError();
        return ts2d_align_up((size_t)n * 8 + (4u << 20), 256);
    }
    return ts2d_align_up(sort_b, 256);
}

// K2/K3: depth order of the triangles, scan of tiles-touched in that order, R to the host.
int ts2d_launch_order_and_scan(int32_t P, GeomState gs, int64_t *R_host, cudaStream_t s)
{
    size_t tb = gs.cub_temp_bytes;
    TS2D_CUDA_TRY(cub::DeviceRadixSort::SortPairs(gs.cub_temp, tb, (const uint32_t *)gs.dkey, gs.dkey2, (const uint32_t *)gs.ids, gs.ids2, P, 0,
                                                  32, s));
    cub::TransformInputIterator<uint32_t, GatherTiles, const uint32_t *> it(gs.ids2, GatherTiles{gs.tiles});
    tb = gs.cub_temp_bytes;
    TS2D_CUDA_TRY(cub::DeviceScan::InclusiveSum(gs.cub_temp, tb, it, gs.offs, P, s));
    k_set_header<<<1, 1, 0, s>>>(gs.hdr, gs.offs, P);
    TS2D_CUDA_TRY(cudaGetLastError());
    int64_t R = 0;
    TS2D_CUDA_TRY(cudaMemcpyAsync(&R, &gs.hdr->num_rendered, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    TS2D_CUDA_TRY(cudaStreamSynchronize(s));
    *R_host = R;
    return 0;
}

static int bits_for(uint32_t n_tiles)
{
    int b = 1;
    while (b < 32 && (1u << b) < n_tiles) b++;
    return b;
}

// K4-K6.
int ts2d_launch_binning(const ts2d_camera *cam, const ts2d_flags *f, int32_t P, int64_t R, GeomState gs, BinState bs, ImageState is,
                        cudaStream_t s)
{
    const int gx = (cam->width + TS2D_TILE - 1) / TS2D_TILE, gy = (cam->height + TS2D_TILE - 1) / TS2D_TILE;
    const int n_tiles = gx * gy;
    TS2D_CUDA_TRY(cudaMemsetAsync(is.ranges, 0, sizeof(uint2) * (size_t)n_tiles, s));
    if (R == 0) return 0;
    k_emit<<<(P + TS2D_BLOCK - 1) / TS2D_BLOCK, TS2D_BLOCK, 0, s>>>(P, gx, f->shard_rank, f->shard_world, gs.ids2, gs.tiles, gs.rect, gs.offs,
                                                                  bs.tkey[0], bs.tval[0]);
    TS2D_CUDA_TRY(cudaGetLastError());
    size_t tb = bs.cub_temp_bytes;
    TS2D_CUDA_TRY(cub::DeviceRadixSort::SortPairs(bs.cub_temp, tb, (const uint32_t *)bs.tkey[0], bs.tkey[1], (const uint32_t *)bs.tval[0],
                                                  bs.tval[1], R, 0, bits_for((uint32_t)n_tiles), s));
    k_ranges<<<(unsigned)((R + TS2D_BLOCK - 1) / TS2D_BLOCK), TS2D_BLOCK, 0, s>>>(R, bs.tkey[1], is.ranges);
    return (int)cudaGetLastError();
}
