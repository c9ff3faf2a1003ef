// Device-wide prefix sums (reduce-then-scan, three launches per level) used for
//   * rank of every valid source point      (numpy boolean compaction x[valid], kd_tree.py:469-470)
//   * cell start offsets of the counting sort (index build)
//   * row compaction / scatter by valid_output_index (kd_tree.py:530-531, 719-723)
#pragma once
#include "prs_common.cuh"

namespace prs {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename In> __device__ __forceinline__ uint32_t scan_load(const In *in, int64_t i, int64_t n)
{
    return i < n ? (uint32_t)(in[i] != 0 ? (sizeof(In) == 1 ? 1u : (uint32_t)in[i]) : 0u) : 0u;
}

// The SCAN_ITEMS consecutive items of a thread; 32-bit inputs at a 16-byte aligned address go as two 128-bit loads.
template <typename In>
__device__ __forceinline__ void scan_load_items(const In *in, int64_t base, int64_t n, uint32_t v[SCAN_ITEMS])
{
    if (sizeof(In) == 4 && base + SCAN_ITEMS <= n && (reinterpret_cast<uintptr_t>(in + base) & 15) == 0) {
        const uint4 a = *reinterpret_cast<const uint4 *>(in + base), b = *reinterpret_cast<const uint4 *>(in + base + 4);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
        return;
    }
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) v[j] = scan_load(in, base + j, n);
}

// n_dev (optional): the length lives in device memory (index build: the size of the cell table is decided by a
// kernel); the grid is then sized for the capacity and blocks stride over the tiles that exist.
template <typename In>
__global__ void scan_tile_sums(const In *__restrict__ in, int64_t n, const int64_t *__restrict__ n_dev,
                               uint32_t *__restrict__ sums)
{
    static_assert(SCAN_ITEMS == 8, "scan_load_items unpacks two uint4");
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    if (n_dev) n = *n_dev;
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int64_t base = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        scan_load_items(in, base, n, v);
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) s += v[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < SCAN_THREADS / 32; ++w) t += warp_sums[w];
            sums[tile] = t;
        }
        __syncthreads();
    }
}

// out[i] = offset[tile] + exclusive (or inclusive) prefix within the tile.  in may alias out (uint32 in place).
template <typename In>
__global__ void scan_tile_apply(const In *in, int64_t n, const int64_t *__restrict__ n_dev,
                                const uint32_t *__restrict__ tile_offsets, uint32_t *out, int inclusive)
{
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    if (n_dev) n = *n_dev;
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int64_t base = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        scan_load_items(in, base, n, v);
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) s += v[j];
        uint32_t incl = s;
        int lane = threadIdx.x & 31;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_sums[w];
        uint32_t run = tile_offsets[tile] + woff + (incl - s);
        uint32_t o[SCAN_ITEMS];
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            o[j] = inclusive ? run + v[j] : run;
            run += v[j];
        }
        if (base + SCAN_ITEMS <= n && (reinterpret_cast<uintptr_t>(out + base) & 15) == 0) {
            *reinterpret_cast<uint4 *>(out + base) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4 *>(out + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j)
                if (base + j < n) out[base + j] = o[j];
        }
        __syncthreads();
    }
}

// One block scans up to a few million items in place, chunk by chunk with a running carry: used for the second
// level of the device-wide scan and for small tables, where a launch costs more than the work.
constexpr int SMALL_SCAN_THREADS = 1024;
static __global__ void __launch_bounds__(SMALL_SCAN_THREADS)
    scan_small_kernel(uint32_t *data, int64_t n, int inclusive, uint32_t *total_out,
                      const int64_t *__restrict__ n_dev = nullptr, int64_t per_item = 1)
{
    if (n_dev) n = (*n_dev + per_item - 1) / per_item;  // e.g. the tiles of a device-sized array
    __shared__ uint32_t warp_sums[SMALL_SCAN_THREADS / 32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += SMALL_SCAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < n ? data[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t before = carry + (warp ? warp_sums[warp - 1] : 0u) + (incl - v);
        if (i < n) data[i] = inclusive ? before + v : before;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_sums[SMALL_SCAN_THREADS / 32 - 1];
        __syncthreads();
    }
    if (total_out && threadIdx.x == 0) *total_out = carry_s;
}
constexpr int64_t SMALL_SCAN_MAX = 1 << 14;

// Exclusive/inclusive scan of n items into out (uint32).  total_dev (optional) receives the grand total.
// Temporary storage comes from the stream-ordered allocator.
template <typename In>
int scan_u32(const In *in, uint32_t *out, int64_t n, int inclusive, uint32_t *total_dev, cudaStream_t st)
{
    if (n <= 0) {
        if (total_dev) PRS_CK(cudaMemsetAsync(total_dev, 0, sizeof(uint32_t), st));
        return PRS_OK;
    }
    if (sizeof(In) == 4 && (const void *)in == (const void *)out && n <= SMALL_SCAN_MAX) {
        scan_small_kernel<<<1, SMALL_SCAN_THREADS, 0, st>>>(out, n, inclusive, total_dev);
        PRS_CKL();
        return PRS_OK;
    }
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t *sums = nullptr;
    PRS_CK(cudaMallocAsync((void **)&sums, sizeof(uint32_t) * (size_t)(tiles + 1), st));
    scan_tile_sums<In><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, nullptr, sums);
    PRS_CKL();
    // exclusive scan of the tile sums (recursive); sums[tiles] gets the total
    uint32_t *tot = sums + tiles;
    if (tiles <= SMALL_SCAN_MAX) {
        scan_small_kernel<<<1, SMALL_SCAN_THREADS, 0, st>>>(sums, tiles, 0, tot);
        PRS_CKL();
    } else {
        int rc = scan_u32<uint32_t>(sums, sums, tiles, 0, tot, st);
        if (rc != PRS_OK) return rc;
    }
    scan_tile_apply<In><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, nullptr, sums, out, inclusive);
    PRS_CKL();
    if (total_dev) PRS_CK(cudaMemcpyAsync(total_dev, tot, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    PRS_CK(cudaFreeAsync(sums, st));
    return PRS_OK;
}

// In-place scan of data[0 .. *n_dev) whose length is only known on the device (*n_dev <= cap).  Three launches with
// fixed grids; the second level is one block over the tile sums that exist.
inline int scan_u32_dev(uint32_t *data, int64_t cap, const int64_t *n_dev, int inclusive, cudaStream_t st)
{
    if (cap <= 0) return PRS_OK;
    const int64_t tiles_cap = (cap + SCAN_TILE - 1) / SCAN_TILE;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        PRS_CK(cudaGetDevice(&dev));
        PRS_CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const unsigned grid = (unsigned)(tiles_cap < (int64_t)n_sm * 8 ? tiles_cap : (int64_t)n_sm * 8);
    uint32_t *sums = nullptr;
    PRS_CK(cudaMallocAsync((void **)&sums, sizeof(uint32_t) * (size_t)(tiles_cap + 1), st));
    scan_tile_sums<uint32_t><<<grid, SCAN_THREADS, 0, st>>>(data, 0, n_dev, sums);
    PRS_CKL();
    scan_small_kernel<<<1, SMALL_SCAN_THREADS, 0, st>>>(sums, 0, 0, nullptr, n_dev, SCAN_TILE);
    PRS_CKL();
    scan_tile_apply<uint32_t><<<grid, SCAN_THREADS, 0, st>>>(data, 0, n_dev, sums, data, inclusive);
    PRS_CKL();
    PRS_CK(cudaFreeAsync(sums, st));
    return PRS_OK;
}

}  // namespace prs
