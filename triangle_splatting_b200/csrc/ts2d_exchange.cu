// Exchange kernels of a tile-sharded render over NVLink peer memory (include/ts2d.h: "Multi-GPU exchange").
// No counterpart in the reference (it has no distributed code).
//
// The composite kernels know nothing about other GPUs: a rank renders the tiles it owns (tile % world == rank) into its own
// replica of a SYMMETRIC buffer (same layout on every rank, mapped by every peer and behind one NVSwitch multicast alias), and
// the two kernels here complete the frame and the gradient sums on every rank:
//   k_exchange_tiles      the 64-byte rows of the owned tiles, replica -> every replica, one multimem.st.v4 per 16 bytes
//                         (the switch fans the store out; disjoint tiles, so every rank ends up with the same frame)
//   k_exchange_allreduce  a rank owns the slice [first, first + count) of an array: multimem.ld_reduce pulls the slice from every
//                         replica and combines it inside the switch, multimem.st writes the result back to every replica.  One
//                         rank computes each element, so all ranks read the same bits; the switch combines in a fixed order,
//                         so two frames give the same bits.  (fp32 add for the per-triangle sums, u32 max for contrib_max, whose
//                         values are non-negative floats: bit order == value order.)
// The caller brackets them with two rendezvous of the ranks: after every rank's local results are complete and before anybody
// publishes (also: nobody still reads the previous frame), and after every rank has published.
#include "ts2d_common.cuh"

namespace {

__device__ __forceinline__ void mc_st128(float *p, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st32(float *p, float v) { asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// One thread per (owned tile, plane, row, 4-pixel group): 16 rows x 4 groups = 64 threads per tile and plane.
__global__ void __launch_bounds__(256) k_exchange_tiles(const float *__restrict__ local, float *mc, int n_planes, int W, int H, int gx, int owned,
                                                       int rank, int world)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per_plane = (int64_t)owned * 64;
    if (t >= per_plane * n_planes) return;
    const int plane = (int)(t / per_plane);
    const int r = (int)(t - plane * per_plane);
    const int tile = rank + (r >> 6) * world;
    const int row = (r >> 2) & 15, grp = r & 3;
    const int x = (tile % gx) * TS2D_TILE + 4 * grp, y = (tile / gx) * TS2D_TILE + row;
    if (y >= H || x >= W) return;
    const size_t off = ((size_t)plane * H + y) * W + x;
    if (x + 3 < W && (off & 3) == 0) {
        mc_st128(mc + off, *reinterpret_cast<const float4 *>(local + off));
    } else {
        for (int i = 0; i < 4 && x + i < W; i++) mc_st32(mc + off + i, local[off + i]);
    }
}

template <int OP>  // 0: fp32 add, four elements per thread; 1: u32 max, one element per thread
__global__ void __launch_bounds__(256) k_exchange_allreduce(void *mc, int64_t first, int64_t count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (OP == 0) {
        if (4 * i >= count) return;
        float *p = reinterpret_cast<float *>(mc) + first + 4 * i;
        float4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
        mc_st128(p, v);
    } else {
        if (i >= count) return;
        uint32_t *p = reinterpret_cast<uint32_t *>(mc) + first + i;
        uint32_t v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.max.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
}

}  // namespace

extern "C" {

int ts2d_exchange_tiles(const float *local, float *multicast, int32_t n_planes, int32_t width, int32_t height, int32_t rank, int32_t world,
                        void *stream)
{
    if (!local || !multicast) return TS2D_E_NULL;
    if (n_planes < 1 || width < 1 || height < 1) return TS2D_E_SIZE;
    if (world < 2 || world > TS2D_MAX_RANKS || rank < 0 || rank >= world || ((uintptr_t)local & 15) || ((uintptr_t)multicast & 15)) return TS2D_E_FABRIC;
    const int gx = (width + TS2D_TILE - 1) / TS2D_TILE, gy = (height + TS2D_TILE - 1) / TS2D_TILE;
    const int owned = (gx * gy - rank + world - 1) / world;
    if (owned <= 0) return 0;
    const int64_t threads = (int64_t)owned * 64 * n_planes;
    k_exchange_tiles<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(local, multicast, n_planes, width, height, gx, owned, rank, world);
    return (int)cudaGetLastError();
}

int ts2d_exchange_allreduce(void *multicast, int64_t first, int64_t count, int32_t op, void *stream)
{
    if (count <= 0) return 0;
    if (!multicast) return TS2D_E_NULL;
    if (first < 0 || (op != TS2D_EXCHANGE_ADD_F32 && op != TS2D_EXCHANGE_MAX_U32)) return TS2D_E_FABRIC;
    if (op == TS2D_EXCHANGE_ADD_F32) {
        if ((first & 3) || (count & 3) || ((uintptr_t)multicast & 15)) return TS2D_E_FABRIC;
        k_exchange_allreduce<0><<<(unsigned)((count / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(multicast, first, count);
    } else {
        k_exchange_allreduce<1><<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(multicast, first, count);
    }
    return (int)cudaGetLastError();
}

}  // extern "C"
