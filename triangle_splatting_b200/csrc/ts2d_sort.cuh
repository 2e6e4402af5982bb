// ts2d_sort.cuh -- hand-written device-wide primitives of the binning chain (sm_100a): a stable LSD radix sort (8-bit digits,
// one kernel per pass, chained look-back between tiles) and a chained inclusive scan.  They replace cub::DeviceRadixSort /
// cub::DeviceScan (which the reference calls at R2D/src/rasterizer.cu:186,211): the problem sizes, the payloads and the element
// count are specific to this pipeline --
//   * the element count may live on the DEVICE (`n_dev`): the host enqueues the whole frame without knowing the number of
//     instances R, grids are sized for the capacity of the caller's buffers and blocks past the end retire immediately;
//   * pass 0 of the tile sort reads only the bits that are sorted on and carries the sub-tile coverage mask in the low key bits.
//
// One pass = one kernel: a block takes a ticket (tile index = ticket, so a block only ever waits for tiles whose blocks are already
// running), ranks its 3072 keys by digit (warp-level match, keys of a warp are consecutive in memory so the ranking is stable),
// publishes its 256 digit counts, looks back over the preceding tiles' published counts (partial | inclusive, one 64-bit word per
// (tile, digit) tagged with the pass id so that one memset per frame serves every pass), and scatters through shared memory so
// that every run of equal digits is written with consecutive addresses.
#pragma once
#include "ts2d_common.cuh"

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 12;                     // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 3072 keys per block
constexpr int RS_BINS = 256;

static inline int64_t rs_tiles(int64_t n) { return (n + RS_TILE - 1) / RS_TILE; }
static inline size_t rs_status_bytes(int64_t n_cap) { return (size_t)(rs_tiles(n_cap > 0 ? n_cap : 1)) * RS_BINS * sizeof(unsigned long long); }

__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// element count of a launch: the device-side count clamped to the capacity the buffers were sized for, or the host's number
__device__ __forceinline__ int64_t rs_count(const int64_t *n_dev, int64_t n_cap)
{
    if (!n_dev) return n_cap;
    const int64_t n = *n_dev;
    return n < n_cap ? n : n_cap;
}

// Chained look-back over the status words of the preceding tiles (stride `stride` words apart, the first at `first`): sum of their
// counts back to the nearest tile that has published an INCLUSIVE prefix.  When a whole wave of blocks starts at once every block has
// to walk back over the partial counts of the blocks in front of it, so the walk issues its loads eight at a time (independent
// loads: one L2 round trip per eight tiles instead of one per tile); words that turn out to lie beyond the stopping point were
// read for nothing, words that are not published yet are simply read again.
__device__ __forceinline__ uint32_t lookback_sum(const unsigned long long *status, int64_t j, size_t stride, unsigned long long tagP,
                                                 unsigned long long tagI)
{
    constexpr int LB = 8;
    uint32_t excl = 0;
    while (true) {
        unsigned long long w[LB];
#pragma unroll
        for (int b = 0; b < LB; b++) w[b] = (j - b >= 0) ? ld_relaxed_u64(status + (size_t)(j - b) * stride) : tagI;  // below tile 0: stop, add 0
        int used = LB;
        bool done = false;
#pragma unroll
        for (int b = 0; b < LB; b++) {
            if (done || used != LB) continue;
            const unsigned long long tg = w[b] & 0xFFFFFFFF00000000ull;
            if (tg == tagI) {
                excl += (uint32_t)w[b];
                done = true;
            } else if (tg == tagP) {
                excl += (uint32_t)w[b];
            } else {
                used = b;  // not published yet (zero, or a word of an earlier pass): resume from here
            }
        }
        if (done) return excl;
        j -= used;
        if (used != LB) __nanosleep(20);
    }
}

// exclusive scan of one value per thread over a 256-thread block; `total` (optional) receives the sum.  s_w: 8 words of scratch.
__device__ __forceinline__ uint32_t block_excl_scan256(uint32_t v, uint32_t *s_w, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    uint32_t base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
        const uint32_t c = s_w[w];
        if (w < warp) base += c;
        sum += c;
    }
    if (total) *total = sum;
    __syncthreads();  // s_w may be reused by the caller
    return base + x - v;
}

// ---- digit histograms of up to four 8-bit digits, one sweep over the keys ---------------------------------------------------
struct RadixHistArgs {
    const uint32_t *keys;
    const int64_t *n_dev;
    int64_t n_cap;
    int ndigits;
    int shift[4];
    uint32_t mask[4];
    uint32_t *hist[4];  // [256] each, zeroed by the caller (stream-ordered)
};

static __global__ void __launch_bounds__(RS_THREADS) k_radix_hist(RadixHistArgs a)
{
    __shared__ uint32_t s_h[4][RS_BINS];
    ts2d_grid_chain();
    const int64_t n = rs_count(a.n_dev, a.n_cap);
    for (int d = 0; d < a.ndigits; d++) s_h[d][threadIdx.x] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t k = a.keys[i];
        // (warp-aggregating equal digits with match.any before the atomic was measured: 2.5x slower than the plain shared atomics)
        for (int d = 0; d < a.ndigits; d++) atomicAdd(&s_h[d][(k >> a.shift[d]) & a.mask[d]], 1u);
    }
    __syncthreads();
    for (int d = 0; d < a.ndigits; d++) {
        const uint32_t c = s_h[d][threadIdx.x];
        if (c) atomicAdd(a.hist[d] + threadIdx.x, c);
    }
}

// ---- one radix pass ------------------------------------------------------------------------------------------------------
struct RadixPassArgs {
    const uint32_t *kin;
    uint32_t *kout;
    const uint32_t *vin;   // NULL: the value of element i is i (first pass over an iota payload)
    uint32_t *vout;
    const uint32_t *hist;  // [256] digit counts of this pass (k_radix_hist)
    unsigned long long *status;  // [tiles][256], zero or stale (other pass ids) on entry
    uint32_t *ticket;      // zero on entry
    const int64_t *n_dev;
    int64_t n_cap;
    int shift;
    uint32_t mask;
    uint32_t pass_uid;     // unique (non-zero) per pass between two memsets of `status`
};

static __global__ void __launch_bounds__(RS_THREADS, 4) k_radix_pass(RadixPassArgs a)
{
    __shared__ uint32_t s_cnt[RS_WARPS][RS_BINS];  // per-warp digit counters, then per-warp exclusive offsets
    __shared__ uint32_t s_texcl[RS_BINS];          // first position of digit d in the tile's sorted order
    __shared__ uint32_t s_gbase[RS_BINS];          // global position of tile-sorted element t with digit d: s_gbase[d] + t
    __shared__ uint32_t s_key[RS_TILE], s_val[RS_TILE];
    __shared__ uint32_t s_w[RS_WARPS];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    ts2d_grid_chain();
    const int64_t n = rs_count(a.n_dev, a.n_cap);
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) s_cnt[w][tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * RS_TILE;
    if (base >= n) return;  // past the end (grids are sized for the capacity): nobody waits for this tile
    const int count = (int)((n - base) < (int64_t)RS_TILE ? (n - base) : (int64_t)RS_TILE);

    // global exclusive digit offsets of this pass
    const uint32_t gexcl = block_excl_scan256(a.hist[tid], s_w, nullptr);

    // load (warp w: 32 * RS_ITEMS consecutive keys; item i of lane l = element w * 384 + i * 32 + l) and rank
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int e = warp * (32 * RS_ITEMS) + i * 32 + lane;
        const bool valid = e < count;
        key[i] = valid ? a.kin[base + e] : 0u;
        val[i] = valid ? (a.vin ? a.vin[base + e] : (uint32_t)(base + e)) : 0u;
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int e = warp * (32 * RS_ITEMS) + i * 32 + lane;
        const bool valid = e < count;
        const uint32_t d = (key[i] >> a.shift) & a.mask;
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : (0x10000u | (uint32_t)lane));  // past-the-end lanes match only themselves
        const uint32_t lower = peers & lt_mask;
        uint32_t prev = 0;
        if (valid) prev = s_cnt[warp][d];
        __syncwarp();
        if (valid && lower == 0u) s_cnt[warp][d] = prev + __popc(peers);
        __syncwarp();
        rank[i] = prev + __popc(lower);
    }
    __syncthreads();

    // digit `tid`: counts of the 8 warps -> exclusive offsets per warp, tile total
    uint32_t tcount = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
        const uint32_t c = s_cnt[w][tid];
        s_cnt[w][tid] = tcount;
        tcount += c;
    }
    const uint32_t texcl = block_excl_scan256(tcount, s_w, nullptr);
    s_texcl[tid] = texcl;

    // chained look-back over the preceding tiles' counts of digit `tid`
    {
        const unsigned long long tagP = ((unsigned long long)((a.pass_uid << 2) | 1u)) << 32, tagI = ((unsigned long long)((a.pass_uid << 2) | 2u)) << 32;
        unsigned long long *mine = a.status + (size_t)tile * RS_BINS + tid;
        uint32_t excl = 0;
        if (tile == 0) {
            st_relaxed_u64(mine, tagI | tcount);
        } else {
            st_relaxed_u64(mine, tagP | tcount);
            excl = lookback_sum(a.status + tid, (int64_t)tile - 1, RS_BINS, tagP, tagI);
            st_relaxed_u64(mine, tagI | (unsigned long long)(excl + tcount));
        }
        s_gbase[tid] = gexcl + excl - texcl;
    }
    __syncthreads();

    // tile-local sorted order in shared memory
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int e = warp * (32 * RS_ITEMS) + i * 32 + lane;
        if (e < count) {
            const uint32_t d = (key[i] >> a.shift) & a.mask;
            const uint32_t pos = s_texcl[d] + s_cnt[warp][d] + rank[i];
            s_key[pos] = key[i];
            s_val[pos] = val[i];
        }
    }
    __syncthreads();
    for (int t = tid; t < count; t += RS_THREADS) {
        const uint32_t k = s_key[t];
        const uint32_t dst = s_gbase[(k >> a.shift) & a.mask] + (uint32_t)t;
        a.kout[dst] = k;
        a.vout[dst] = s_val[t];
    }
}

// ---- device-wide scans: block sums -> scan of the block sums (one block) -> apply ------------------------------------------
// Three short kernels instead of one chained pass: a chain of look-backs over a few thousand tiles that all start in the same
// wave costs more than re-reading a few MB (measured: 55 us chained vs < 20 us for the 7 MB row-count scan of the backward).
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = RS_THREADS * SC_ITEMS;  // 4096 elements per block
static inline int64_t sc_tiles(int64_t n) { return (n + SC_TILE - 1) / SC_TILE; }

// A Load gives f(i) for the strided block-sum pass and block16() -- 16 consecutive elements starting at a multiple of 16, all below n --
// for the thread-blocked apply pass (vector loads: one 16-byte load for bytes, four for the gathered indices).
struct LoadGatherU32 {  // f(i) = src[order[i]]: tiles-touched in depth-rank order (rasterizer.cu:186)
    const uint32_t *order, *src;
    __device__ __forceinline__ uint32_t operator()(int64_t i) const { return src[order[i]]; }
    __device__ __forceinline__ void block16(int64_t base, uint32_t (&v)[16]) const
    {
        const uint4 *o = reinterpret_cast<const uint4 *>(order + base);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint4 w = __ldg(o + q);
            v[4 * q + 0] = src[w.x];
            v[4 * q + 1] = src[w.y];
            v[4 * q + 2] = src[w.z];
            v[4 * q + 3] = src[w.w];
        }
    }
};
struct LoadU8 {
    const uint8_t *src;
    __device__ __forceinline__ uint32_t operator()(int64_t i) const { return src[i]; }
    __device__ __forceinline__ void block16(int64_t base, uint32_t (&v)[16]) const
    {
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(src + base));
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int b = 0; b < 4; b++) v[4 * q + b] = (ws[q] >> (8 * b)) & 0xffu;
    }
};

template <class Load>
static __global__ void __launch_bounds__(RS_THREADS) k_scan_sums(Load f, const int64_t *n_dev, int64_t n_cap, uint32_t *__restrict__ sums)
{
    __shared__ uint32_t s_w[RS_WARPS];
    ts2d_grid_chain();
    const int64_t n = rs_count(n_dev, n_cap);
    const int64_t base = (int64_t)blockIdx.x * SC_TILE;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {  // strided: coalesced loads, the order inside a block does not matter for its sum
        const int64_t e = base + (int64_t)i * RS_THREADS + threadIdx.x;
        if (e < n) sum += f(e);
    }
    uint32_t total;
    block_excl_scan256(sum, s_w, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// exclusive scan of nb block sums in place (one block of 1024 threads); the grand total goes to total64 and / or total32
static __global__ void __launch_bounds__(1024) k_scan_block_sums(uint32_t *sums, int nb, int64_t *total64, uint32_t *total32, const int64_t *n_dev, int64_t n_cap)
{
    __shared__ uint32_t s_part[1024];
    ts2d_grid_chain();
    const int tid = threadIdx.x;
    const int per = (nb + 1023) / 1024;
    const int b0 = tid * per, b1 = min(nb, b0 + per);
    uint32_t sum = 0;
    for (int b = b0; b < b1; b++) sum += sums[b];
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partial sums
        const uint32_t v = tid >= o ? s_part[tid - o] : 0u;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint32_t run = s_part[tid] - sum;
    for (int b = b0; b < b1; b++) {
        const uint32_t c = sums[b];
        sums[b] = run;
        run += c;
    }
    if (tid == 1023) {
        if (total64) *total64 = (int64_t)s_part[1023];
        if (total32) total32[rs_count(n_dev, n_cap)] = s_part[1023];  // exclusive-scan output of n elements: out[n] = total
    }
}

template <class Load, bool INCLUSIVE>
static __global__ void __launch_bounds__(RS_THREADS)
k_scan_apply(Load f, const int64_t *n_dev, int64_t n_cap, const uint32_t *__restrict__ sums, uint32_t *__restrict__ out)
{
    __shared__ uint32_t s_w[RS_WARPS];
    ts2d_grid_chain();
    const int64_t n = rs_count(n_dev, n_cap);
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;  // thread t owns SC_ITEMS consecutive elements
    if ((int64_t)blockIdx.x * SC_TILE >= n) return;
    static_assert(SC_ITEMS == 16, "block16");
    uint32_t v[SC_ITEMS], sum = 0;
    const bool full = base + SC_ITEMS <= n;
    if (full) {
        f.block16(base, v);
    } else {
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) v[i] = (base + i < n) ? f(base + i) : 0u;
    }
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) sum += v[i];
    uint32_t run = block_excl_scan256(sum, s_w, nullptr) + sums[blockIdx.x];
    if (full) {
        uint4 *o = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t r[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                if (INCLUSIVE) run += v[4 * q + b];
                r[b] = run;
                if (!INCLUSIVE) run += v[4 * q + b];
            }
            o[q] = make_uint4(r[0], r[1], r[2], r[3]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < SC_ITEMS; i++) {
            if (INCLUSIVE) run += v[i];
            if (base + i < n) out[base + i] = run;
            if (!INCLUSIVE) run += v[i];
        }
    }
}

// host side: scan of n (host or device count, <= n_cap) elements; `sums` holds sc_tiles(n_cap) words of scratch
template <class Load, bool INCLUSIVE>
static inline cudaError_t ts2d_scan(Load f, const int64_t *n_dev, int64_t n_cap, uint32_t *sums, uint32_t *out, int64_t *total64, bool total_at_end, cudaStream_t s)
{
    const int nb = (int)sc_tiles(n_cap > 0 ? n_cap : 1);
    cudaError_t e = ts2d_launch_chained(k_scan_sums<Load>, nb, RS_THREADS, 0, s, f, n_dev, n_cap, sums);
    if (e == cudaSuccess) e = ts2d_launch_chained(k_scan_block_sums, 1, 1024, 0, s, sums, nb, total64, total_at_end ? out : (uint32_t *)nullptr, n_dev, n_cap);
    if (e == cudaSuccess) e = ts2d_launch_chained(k_scan_apply<Load, INCLUSIVE>, nb, RS_THREADS, 0, s, f, n_dev, n_cap, (const uint32_t *)sums, out);
    return e;
}
