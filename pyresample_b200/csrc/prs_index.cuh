// Spatial index over the valid source points: cube-face cell grid built by a counting sort
// (one-digit radix sort on the cell key), replacing pykdtree.KDTree(input_coords) (kd_tree.py:464-489).
//
// Layout in HBM
//   recs[n_indexed]      Rec<T> = ECEF xyz in the tree dtype + raveled source index, sorted by cell key
//   cell_start[M + 1]    prefix sums over the cell table.  The table only spans, per cube face, the bounding
//                        rectangle of the cells that can hold a point (from an occupancy probe), row-major:
//                        key = face_base[f] + (iv - fv0[f]) * fw[f] + (iu - fu0[f])
//   coarse_start[Mc + 1] the same for COARSE x COARSE blocks of cells (L2-resident; answers "is there any
//                        source point within the radius at all" for the many target pixels far from the swath)
//   rank[n_source]       rank of every source point among the valid ones (omitted when all are valid)
#pragma once
#include "prs_common.cuh"
#include "prs_scan.cuh"
#include <chrono>
#include <mutex>
#include <utility>
#include <vector>

namespace prs {

constexpr int COARSE = 16;   // fine cells per coarse cell side
constexpr int OCC_G = 128;   // resolution of the occupancy probe used to size the cells and the table
constexpr uint32_t CELL_SKIP = 0xFFFFFFFFu;

// Per-face table geometry (fine and coarse).  A face without points has fw == 0.
struct FaceTable {
    int fu0[6], fv0[6], fw[6], fh[6];
    int64_t base[6];
};

}  // namespace prs

struct prs_index {
    int tree_dtype;      // PRS_F32 / PRS_F64
    int64_t n_source;    // raveled source size
    int64_t n_valid;     // pykdtree's .n : number of valid source points (rank space size)
    int64_t n_indexed;   // valid and not excluded
    int G;               // fine cells per face side (multiple of COARSE)
    int Gc;              // coarse cells per face side
    int all_valid;       // rank == source index
    uint32_t first_valid_source;  // raveled index of rank 0 (kd_tree.py:754-759 gathers new_data[0] for missing)
    void *recs;
    uint32_t *cell_start;    // = cell_alloc + 3
    uint32_t *coarse_start;  // = coarse_alloc + 3
    uint32_t *cell_alloc, *coarse_alloc;
    uint32_t *rank;
    prs::FaceTable fine, coarse;
    int64_t n_cells, n_coarse;
    size_t bytes;
    cudaStream_t stream;  // stream the memory was allocated on (stream-ordered free)
    // queries launched on OTHER streams: one event per stream, recorded behind the stream's latest query, so that
    // prs_index_destroy can order the stream-ordered free behind them (see index_note_use / prs_index_destroy)
    struct prs_index_uses *uses;
};

struct prs_index_uses {
    std::mutex lock;
    std::vector<std::pair<cudaStream_t, cudaEvent_t>> events;
};

namespace prs {

// A query on a stream other than the one the index lives on: remember it, so that the free waits for it.
inline void index_note_use(const prs_index *cix, cudaStream_t st)
{
    prs_index *ix = const_cast<prs_index *>(cix);
    if (!ix || !ix->uses || st == ix->stream) return;
    std::lock_guard<std::mutex> guard(ix->uses->lock);
    for (auto &u : ix->uses->events)
        if (u.first == st) {
            cudaEventRecord(u.second, st);
            return;
        }
    cudaEvent_t ev = nullptr;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    ix->uses->events.emplace_back(st, ev);
}

// Stream-ordered allocations come from the device's default pool; keep freed blocks cached in the pool
// instead of returning them to the driver (cudaMalloc/cudaFree of the 100 MB-class tables costs tens of ms).
inline int ensure_pool_configured()
{
    static std::mutex lock;
    static bool done[64] = {false};
    int dev = 0;
    PRS_CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> guard(lock);
    const int slot = dev < 0 || dev >= 64 ? 0 : dev;
    if (done[slot] && dev >= 0 && dev < 64) return PRS_OK;
    cudaMemPool_t pool;
    PRS_CK(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long thr = ~0ull;
    PRS_CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    done[slot] = true;
    return PRS_OK;
}

__device__ __forceinline__ void face_cell_of(float x, float y, float z, int G, int &face, int &iu, int &iv)
{
    float u, v;
    face_uv(x, y, z, face, u, v);
    iu = cell_coord(warp_uv(u), G);
    iv = cell_coord(warp_uv(v), G);
}

// Cell of a point on a GIVEN face (not necessarily its dominant one): used for points that sit on a cube edge,
// where the last bit of the coordinates decides the face.  u / v beyond the edge clamp to the edge cells.
__device__ __forceinline__ void cell_on_face(int face, float x, float y, float z, int G, int &iu, int &iv)
{
    const int axis = face >> 1;
    const float a = axis == 0 ? x : (axis == 1 ? y : z);
    const float inv = 1.0f / fmaxf(fabsf(a), 1e-30f);
    const float u = (axis == 0 ? y : x) * inv, v = (axis == 2 ? y : z) * inv;
    iu = cell_coord(warp_uv(u), G);
    iv = cell_coord(warp_uv(v), G);
}

template <typename T> __device__ __forceinline__ void unit_f32(T x, T y, T z, float &fx, float &fy, float &fz)
{
    const float invR = (float)(1.0 / PRS_R);
    fx = (float)x * invR;
    fy = (float)y * invR;
    fz = (float)z * invR;
}

// Read-back of a few words: cudaMemcpyAsync + cudaStreamSynchronize by default.  Optional low-latency variant
// (PRS_MAILBOX=1): a one-warp kernel copies them into mapped pinned host memory and raises a sequence flag, the
// host spins on the flag.
static __global__ void publish_kernel(const uint32_t *__restrict__ src, volatile uint32_t *dst, int n_words, uint32_t seq)
{
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[1 + i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) dst[0] = seq;
}

inline int read_back_small(const void *dev_src, void *host_dst, size_t bytes, cudaStream_t st)
{
    struct Mailbox {
        volatile uint32_t *host = nullptr;
        uint32_t *dev = nullptr;
        uint32_t seq = 0;
    };
    static thread_local Mailbox mb;
    const int n_words = (int)((bytes + 3) / 4);
    // opt-in (PRS_MAILBOX=1): it saves ~10 us per build, but with NVML clock queries running next to it (bench.py's
    // sampler, any monitoring agent) the flag write was seen to arrive tens of milliseconds late
    static const bool mailbox_enabled = [] {
        const char *e = getenv("PRS_MAILBOX");
        return e && e[0] == '1';
    }();
    const bool use_mailbox = n_words <= 255 && mailbox_enabled;
    if (use_mailbox && !mb.host) {
        void *h = nullptr;
        if (cudaHostAlloc(&h, 1024, cudaHostAllocMapped) == cudaSuccess &&
            cudaHostGetDevicePointer((void **)&mb.dev, h, 0) == cudaSuccess) {
            mb.host = (volatile uint32_t *)h;
            mb.host[0] = 0;
        } else {
            cudaGetLastError();
        }
    }
    if (use_mailbox && mb.host) {
        const uint32_t seq = ++mb.seq ? mb.seq : ++mb.seq;  // never 0
        publish_kernel<<<1, 64, 0, st>>>((const uint32_t *)dev_src, mb.dev, n_words, seq);
        PRS_CKL();
        // spin (bounded), then fall back to a blocking wait so that an error cannot hang the host
        for (long spins = 0; mb.host[0] != seq; ++spins) {
            if (spins > 20000000L) {
                PRS_CK(cudaStreamSynchronize(st));
                if (mb.host[0] != seq) return fail(PRS_ECUDA, "index build: read-back flag never arrived");
            }
        }
        memcpy(host_dst, (const void *)(mb.host + 1), bytes);
        return PRS_OK;
    }
    PRS_CK(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, st));
    PRS_CK(cudaStreamSynchronize(st));
    return PRS_OK;
}

struct BuildCounters {
    uint32_t n_valid;       // valid source points (pykdtree's .n)
    uint32_t n_skipped;     // valid but excluded (query mask) or outside this rank's target coverage
    uint32_t first_source;  // raveled index of the first valid point (rank 0)
    uint32_t n_occ;         // occupied probe cells
};

// _get_valid_input_index (kd_tree.py:406, 426-428) and the reduce_data box of data_reduce.py:282-305.  numpy promotes
// float32 coordinates to float64 against the float64 bounds; float64 conversions and compares are slow on the
// GPU, so for float32 coordinates the bounds are rounded on the host to the float32 value that gives the identical
// decision ((double)x >= b  <=>  x >= roundup32(b);  (double)x <= b  <=>  x <= rounddown32(b)).
template <typename T> struct BoxT {
    int kind;
    T lat_min, lat_max, lon_min, lon_max;
};
inline float round_up_f32(double b)
{
    float f = (float)b;
    if ((double)f < b) f = nextafterf(f, INFINITY);
    return f;
}
inline float round_down_f32(double b)
{
    float f = (float)b;
    if ((double)f > b) f = nextafterf(f, -INFINITY);
    return f;
}
inline void make_box(const prs_reduce_box &b, BoxT<float> &o)
{
    o.kind = b.kind;
    o.lat_min = round_up_f32(b.lat_min);
    o.lon_min = round_up_f32(b.lon_min);
    o.lat_max = round_down_f32(b.lat_max);
    o.lon_max = round_down_f32(b.lon_max);
}
inline void make_box(const prs_reduce_box &b, BoxT<double> &o)
{
    o.kind = b.kind;
    o.lat_min = b.lat_min;
    o.lon_min = b.lon_min;
    o.lat_max = b.lat_max;
    o.lon_max = b.lon_max;
}
template <typename T>
__device__ __forceinline__ bool valid_predicate(T lo, T la, const BoxT<T> &box, bool premasked)
{
    bool ok = lonlat_in_range<T>(lo, la) && !premasked;
    switch (box.kind) {
    case PRS_REDUCE_LAT_GE:
        ok = ok && (la >= box.lat_min);
        break;
    case PRS_REDUCE_LAT_LE:
        ok = ok && (la <= box.lat_max);
        break;
    case PRS_REDUCE_BOX:
        ok = ok && (la >= box.lat_min) && (la <= box.lat_max) && (lo >= box.lon_min) && (lo <= box.lon_max);
        break;
    case PRS_REDUCE_BOX_DATELINE:
        ok = ok && (la >= box.lat_min) && (la <= box.lat_max) &&
             (((lo >= box.lon_min) && (lo <= (T)180)) || ((lo <= box.lon_max) && (lo >= (T)-180)));
        break;
    default:
        break;
    }
    return ok;
}

// Temporaries of the build (scratch counters, unsorted records, keep flags) live in one grow-only buffer per
// device that stays allocated between builds.  Taking 100 MB-class temporaries from the stream-ordered pool on
// every build made the pool re-map memory for the long-lived arrays every few calls (1.5 ms inside cudaMallocAsync,
// seen on C3); with the temporaries out of the way the pool only ever sees the same three long-lived sizes.
// A build holds the buffer (mutex) while it queues its kernels; the next user waits on the event recorded behind
// the last kernel that touched it, so builds on different streams stay correctly ordered.
struct BuildWorkspace {
    std::mutex lock;
    void *buf = nullptr;
    size_t capacity = 0;
    cudaEvent_t last_use = nullptr;
    bool used = false;
};
inline BuildWorkspace &build_workspace(int dev)
{
    static BuildWorkspace per_device[64];
    return per_device[dev < 0 || dev >= 64 ? 0 : dev];
}
inline int release_build_workspace()
{
    int dev = 0;
    PRS_CK(cudaGetDevice(&dev));
    BuildWorkspace &ws = build_workspace(dev);
    std::lock_guard<std::mutex> guard(ws.lock);
    PRS_CK(cudaDeviceSynchronize());
    if (ws.buf) PRS_CK(cudaFree(ws.buf));
    ws.buf = nullptr;
    ws.capacity = 0;
    ws.used = false;
    cudaMemPool_t pool;
    PRS_CK(cudaDeviceGetDefaultMemPool(&pool, dev));
    PRS_CK(cudaMemPoolTrimTo(pool, 0));
    return PRS_OK;
}

// RAII: acquire() takes the mutex and makes `st` wait for the previous user; the destructor records the event.
struct WorkspaceLease {
    BuildWorkspace *ws = nullptr;
    cudaStream_t st = nullptr;
    int acquire(size_t bytes, cudaStream_t stream, void **out)
    {
        int dev = 0;
        PRS_CK(cudaGetDevice(&dev));
        ws = &build_workspace(dev);
        ws->lock.lock();
        st = stream;
        if (!ws->last_use) PRS_CK(cudaEventCreateWithFlags(&ws->last_use, cudaEventDisableTiming));
        if (ws->used) PRS_CK(cudaStreamWaitEvent(st, ws->last_use, 0));
        if (ws->capacity < bytes) {
            if (ws->buf) {
                PRS_CK(cudaEventSynchronize(ws->last_use));
                PRS_CK(cudaFree(ws->buf));
                ws->buf = nullptr;
                ws->capacity = 0;
            }
            const size_t want = bytes + bytes / 8;  // a little headroom: sizes of successive swaths differ slightly
            PRS_CK(cudaMalloc(&ws->buf, want));
            ws->capacity = want;
        }
        *out = ws->buf;
        return PRS_OK;
    }
    ~WorkspaceLease()
    {
        if (!ws) return;
        if (ws->last_use && cudaEventRecord(ws->last_use, st) == cudaSuccess) ws->used = true;
        ws->lock.unlock();
    }
};

// PRS_BUILD_TIMING=1: host-side time stamps of the build on stderr (where does the host wait?)
struct BuildTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    BuildTimer() : t0(std::chrono::steady_clock::now())
    {
        static const bool enabled = getenv("PRS_BUILD_TIMING") != nullptr;
        on = enabled;
    }
    void mark(const char *what)
    {
        if (!on) return;
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[prs build] %-28s %9.1f us\n", what, us);
    }
};

#ifndef PRS_SELECT_MIN_BLOCKS
#define PRS_SELECT_MIN_BLOCKS 8  // 32 registers, no spills: full occupancy for the issue-bound selection pass
#endif
constexpr int BUILD_TPB = 256;
constexpr int BUILD_ITEMS = 4;  // points per thread in the build passes

// Pass 1: which points are valid, which of them this index needs, and an occupancy probe at OCC_G resolution.
//   valid (in or out)        -> 0: not part of the tree at all
//   exclude[i] (query mask)  -> keeps its rank, never returned
//   cover (optional)         -> points whose coverage cell is 0 are beyond the radius of every target point of
//                               this rank
// Only approximate coordinates (fast sin/cos intrinsics) are needed here: probe and coverage cells are tens of
// kilometres wide and both masks carry a whole cell of slack.  The exact ECEF transform runs in pass 2, for the
// points that are kept, once the cell size is known.
template <typename T>
__global__ void __launch_bounds__(BUILD_TPB, PRS_SELECT_MIN_BLOCKS)
    build_select_kernel(const T *__restrict__ lon, const T *__restrict__ lat, int64_t n, uint8_t *__restrict__ valid,
                        int compute_valid, const BoxT<T> box, const uint8_t *__restrict__ premask,
                        const uint8_t *__restrict__ exclude, const uint8_t *__restrict__ cover,
                        uint8_t *__restrict__ keep, uint8_t *__restrict__ occ, BuildCounters *counters,
                        uint2 *__restrict__ block_counts)
{
    // a warp works on 32 consecutive points at a time (their cells mostly coincide, so its occupancy stores and
    // the histogram atomics of pass 2 collapse into one or two transactions); all loads are issued first
    const int64_t base = (int64_t)blockIdx.x * (BUILD_TPB * BUILD_ITEMS) + threadIdx.x;
    int nv = 0, ns = 0;
    uint32_t first = 0xFFFFFFFFu;
    T lo[BUILD_ITEMS], la[BUILD_ITEMS];
    uint8_t vin[BUILD_ITEMS], ex[BUILD_ITEMS];
#pragma unroll
    for (int j = 0; j < BUILD_ITEMS; ++j) {
        const int64_t i = base + (int64_t)j * BUILD_TPB;
        lo[j] = la[j] = 0;
        vin[j] = ex[j] = 0;
        if (i < n) {
            lo[j] = lon[i];
            la[j] = lat[i];
            vin[j] = compute_valid ? (premask ? premask[i] : (uint8_t)0) : valid[i];
            ex[j] = exclude ? exclude[i] : (uint8_t)0;
        }
    }
    int last_cell = -1;  // probe cells are tens of km wide: a thread's points mostly share one
    __shared__ int s_cells[64];
    if (threadIdx.x < 64) s_cells[threadIdx.x] = -1;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < BUILD_ITEMS; ++j) {
        const int64_t i = base + (int64_t)j * BUILD_TPB;
        if (i >= n) continue;
        const bool v = compute_valid ? valid_predicate<T>(lo[j], la[j], box, vin[j] != 0) : vin[j] != 0;
        if (compute_valid && valid) valid[i] = v ? 1 : 0;
        bool kept = false;
        if (v) {
            ++nv;
            first = min(first, (uint32_t)i);
            bool skip = ex[j] != 0;
            if (!skip) {
                float slo, clo, sla, cla;
                __sincosf((float)lo[j] * 0.017453292f, &slo, &clo);
                __sincosf((float)la[j] * 0.017453292f, &sla, &cla);
                const float fx = cla * clo, fy = cla * slo, fz = sla;
                int face, iu, iv;
                if (cover) {
                    face_cell_of(fx, fy, fz, PRS_COVER_CELLS, face, iu, iv);
                    skip = !cover[((size_t)face * PRS_COVER_CELLS + iv) * PRS_COVER_CELLS + iu];
                }
                if (!skip) {
                    face_cell_of(fx, fy, fz, OCC_G, face, iu, iv);
                    const int cell = (face * OCC_G + iv) * OCC_G + iu;
                    if (cell != last_cell) {
                        // every block sets the same few bytes, and a cache line that the whole grid writes is a
                        // serial resource: announce each cell once per block (a race only costs a repeated store)
                        if (s_cells[cell & 63] != cell) {
                            s_cells[cell & 63] = cell;
                            occ[cell] = 1;
                        }
                        last_cell = cell;
                    }
                    // on a cube edge (or corner) the exact coordinates of pass 2 may pick another face than these
                    // approximate ones: the probe cell is marked on every face that can win, so the table of
                    // whichever face pass 2 chooses spans the point
                    const float ax = fabsf(fx), ay = fabsf(fy), az = fabsf(fz);
                    const float lim = fmaxf(ax, fmaxf(ay, az)) * (1.0f - 1e-4f);
#pragma unroll
                    for (int axis = 0; axis < 3; ++axis) {
                        const float a = axis == 0 ? ax : (axis == 1 ? ay : az);
                        const float sv = axis == 0 ? fx : (axis == 1 ? fy : fz);
                        const int f2 = 2 * axis + (sv < 0.f ? 1 : 0);
                        if (a >= lim && f2 != face) {
                            int ju, jv;
                            cell_on_face(f2, fx, fy, fz, OCC_G, ju, jv);
                            occ[(f2 * OCC_G + jv) * OCC_G + ju] = 1;
                        }
                    }
                    kept = true;
                }
            }
            if (!kept) ++ns;
        }
        keep[i] = kept ? 1 : 0;
    }
    // block totals (one record per block, reduced by build_reduce_kernel: no same-address atomics storm)
    __shared__ int s_nv[BUILD_TPB / 32], s_ns[BUILD_TPB / 32];
    __shared__ uint32_t s_first[BUILD_TPB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        ns += __shfl_xor_sync(0xffffffffu, ns, o);
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_nv[threadIdx.x >> 5] = nv;
        s_ns[threadIdx.x >> 5] = ns;
        s_first[threadIdx.x >> 5] = first;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tv = 0, ts = 0;
        uint32_t tf = 0xFFFFFFFFu;
        for (int w = 0; w < BUILD_TPB / 32; ++w) {
            tv += s_nv[w];
            ts += s_ns[w];
            tf = min(tf, s_first[w]);
        }
        block_counts[blockIdx.x] = make_uint2((unsigned)tv, (unsigned)ts);
        // rank 0's source index: blocks run roughly in order, so almost all of them skip the atomic
        if (tf < *(volatile uint32_t *)&counters->first_source) atomicMin(&counters->first_source, tf);
    }
}

// Reduce the per-block counts and the occupancy probe: number of valid / skipped points, bounding rectangle (in
// probe cells) of the occupied probe cells of each face (rect[f] = {u0, v0, u1, v1}) and their number.
static __global__ void build_reduce_kernel(const uint8_t *__restrict__ occ, const uint2 *__restrict__ block_counts,
                                           int64_t n_blocks, int *__restrict__ rect, BuildCounters *counters)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    unsigned nv = 0, ns = 0, no = 0;
    for (int64_t b = tid; b < n_blocks; b += nth) {
        uint2 c = block_counts[b];
        nv += c.x;
        ns += c.y;
    }
    // four probe cells per thread (one 32-bit load); a warp covers one row of OCC_G = 128 cells of one face, so
    // its bounding box is reduced in registers and costs at most four atomics
    static_assert(OCC_G == 128, "a warp reads exactly one row of probe cells");
    const uint32_t *occ4 = reinterpret_cast<const uint32_t *>(occ);
    for (int w = tid; w < 6 * OCC_G * OCC_G / 4; w += nth) {
        const uint32_t bits = occ4[w];
        const int i = w * 4, f = i / (OCC_G * OCC_G), iv = (i / OCC_G) % OCC_G, iu = i % OCC_G;
        int u0 = 1 << 30, u1 = -1;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((bits >> (8 * k)) & 0xFFu) {
                u0 = min(u0, iu + k);
                u1 = max(u1, iu + k);
                ++no;
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            u0 = min(u0, __shfl_xor_sync(0xffffffffu, u0, o));
            u1 = max(u1, __shfl_xor_sync(0xffffffffu, u1, o));
        }
        if ((threadIdx.x & 31) == 0 && u1 >= 0) {
            atomicMin(&rect[4 * f + 0], u0);
            atomicMin(&rect[4 * f + 1], iv);
            atomicMax(&rect[4 * f + 2], u1);
            atomicMax(&rect[4 * f + 3], iv);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        ns += __shfl_xor_sync(0xffffffffu, ns, o);
        no += __shfl_xor_sync(0xffffffffu, no, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nv) atomicAdd(&counters->n_valid, nv);
        if (ns) atomicAdd(&counters->n_skipped, ns);
        if (no) atomicAdd(&counters->n_occ, no);
    }
}

__device__ __forceinline__ int64_t table_key(const FaceTable &t, int face, int iu, int iv)
{
    return t.base[face] + (int64_t)(iv - t.fv0[face]) * t.fw[face] + (iu - t.fu0[face]);
}

// Pass 2: exact ECEF coordinates of the kept points, their cell key and the histogram into table[key + 1].
// tmp[i] = (x, y, z, cell key) - the record's source index is its position i, so the fourth word carries the key
// to pass 3.
// SPARSE (a rank of a sharded run keeps a fraction of the points): the coordinates are only loaded for kept points;
// otherwise all loads are issued up front.
template <typename T, bool SPARSE>
__global__ void __launch_bounds__(BUILD_TPB)
    build_xyz_kernel(const T *__restrict__ lon, const T *__restrict__ lat, int64_t n, const uint8_t *__restrict__ keep,
                     int G, const FaceTable ft, Rec<T> *__restrict__ tmp, uint32_t *__restrict__ table)
{
    const int64_t base = (int64_t)blockIdx.x * (BUILD_TPB * BUILD_ITEMS) + threadIdx.x;
    T lo[BUILD_ITEMS], la[BUILD_ITEMS];
    uint8_t kp[BUILD_ITEMS];
#pragma unroll
    for (int j = 0; j < BUILD_ITEMS; ++j) {
        const int64_t i = base + (int64_t)j * BUILD_TPB;
        kp[j] = i < n ? keep[i] : (uint8_t)0;
    }
#pragma unroll
    for (int j = 0; j < BUILD_ITEMS; ++j) {
        const int64_t i = base + (int64_t)j * BUILD_TPB;
        lo[j] = la[j] = 0;
        if (i < n && (!SPARSE || kp[j])) {
            lo[j] = lon[i];
            la[j] = lat[i];
        }
    }
#pragma unroll
    for (int j = 0; j < BUILD_ITEMS; ++j) {
        const int64_t i = base + (int64_t)j * BUILD_TPB;
        if (!kp[j]) {
            // dense builds mark the few skipped points in tmp, so that pass 3 needs nothing but tmp
            if (!SPARSE && i < n) store_rec(tmp + i, (T)0, (T)0, (T)0, CELL_SKIP);
            continue;
        }
        T x, y, z;
        lonlat_to_xyz(lo[j], la[j], x, y, z);
        float fx, fy, fz;
        unit_f32<T>(x, y, z, fx, fy, fz);
        int face, iu, iv;
        face_cell_of(fx, fy, fz, G, face, iu, iv);
        // The table rectangle of the face is a superset of the probe cells pass 1 marked for this point - also when
        // the point sits on a cube edge and pass 1 saw it on the neighbouring face (it marks both).  Should the face
        // still have no table (cannot happen by construction), file the point on the nearest face that has one:
        // a query that reaches the point's position also reaches the edge cells of that face.
        if (ft.fw[face] == 0) {
            const float ax = fabsf(fx), ay = fabsf(fy), az = fabsf(fz);
            float best = -1.f;
            int bf = -1;
#pragma unroll
            for (int axis = 0; axis < 3; ++axis) {
                const float a = axis == 0 ? ax : (axis == 1 ? ay : az);
                const float sv = axis == 0 ? fx : (axis == 1 ? fy : fz);
                const int f2 = 2 * axis + (sv < 0.f ? 1 : 0);
                if (f2 != face && ft.fw[f2] != 0 && a > best) {
                    best = a;
                    bf = f2;
                }
            }
            if (bf < 0) {  // no table at all on the faces this point touches: drop it from the index
                store_rec(tmp + i, (T)0, (T)0, (T)0, CELL_SKIP);
                continue;
            }
            face = bf;
            cell_on_face(face, fx, fy, fz, G, iu, iv);
        }
        iu = min(max(iu, ft.fu0[face]), ft.fu0[face] + ft.fw[face] - 1);
        iv = min(max(iv, ft.fv0[face]), ft.fv0[face] + ft.fh[face] - 1);
        const uint32_t c = (uint32_t)table_key(ft, face, iu, iv);
        atomicAdd(&table[c + 1], 1u);
        store_rec(tmp + i, x, y, z, c);
    }
}

// Pass 3: scatter.  table[c + 1] holds start(c) after the exclusive scan and ends as start(c + 1).
constexpr int SCATTER_ITEMS = 4;
template <typename T, bool SPARSE>
__global__ void __launch_bounds__(BUILD_TPB)
    build_scatter_kernel(const Rec<T> *__restrict__ tmp, const uint8_t *__restrict__ keep, int64_t n,
                         uint32_t *__restrict__ table, Rec<T> *__restrict__ recs)
{
    const int64_t base = (int64_t)blockIdx.x * (BUILD_TPB * SCATTER_ITEMS) + threadIdx.x;
    Rec<T> r[SCATTER_ITEMS];
    uint8_t kp[SCATTER_ITEMS];
    // dense: every record was written by pass 2 (CELL_SKIP marks the skipped ones); sparse: only the records of
    // kept points exist, the keep flags say which
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j) {
        const int64_t i = base + (int64_t)j * BUILD_TPB;
        r[j].idx = CELL_SKIP;
        kp[j] = 0;
        if (i < n) {
            if (SPARSE)
                kp[j] = keep[i];
            else
                r[j] = load_rec(tmp + i);
        }
    }
    if (SPARSE) {
#pragma unroll
        for (int j = 0; j < SCATTER_ITEMS; ++j)
            if (kp[j]) r[j] = load_rec(tmp + base + (int64_t)j * BUILD_TPB);
    }
    // consecutive source points mostly share a cell: one atomic per (warp, cell) instead of one per point
    const unsigned lane = threadIdx.x & 31u;
    uint32_t pos[SCATTER_ITEMS];
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j) {
        const uint32_t c = r[j].idx;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        const unsigned leader = __ffs(peers) - 1;
        uint32_t first = 0;
        if (lane == leader && c != CELL_SKIP) first = atomicAdd(&table[c + 1], (uint32_t)__popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        pos[j] = first + __popc(peers & ((1u << lane) - 1u));
    }
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j)
        if (r[j].idx != CELL_SKIP)
            store_rec(recs + pos[j], r[j].x, r[j].y, r[j].z, (uint32_t)(base + (int64_t)j * BUILD_TPB));
}

// Coarse table: number of points in each COARSE x COARSE block of fine cells, from the fine row prefix sums.
static __global__ void build_coarse_kernel(const uint32_t *__restrict__ cell_start, const FaceTable fine,
                                           const FaceTable coarse, int64_t n_coarse, uint32_t *__restrict__ ctable)
{
    int64_t cc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cc >= n_coarse) return;
    int face = 5;
#pragma unroll
    for (int f = 4; f >= 0; --f)
        if (cc < coarse.base[f + 1]) face = f;
    int64_t local = cc - coarse.base[face];
    int cu = coarse.fu0[face] + (int)(local % coarse.fw[face]);
    int cv = coarse.fv0[face] + (int)(local / coarse.fw[face]);
    uint32_t tot = 0;
    for (int r = 0; r < COARSE; ++r) {
        int64_t k0 = table_key(fine, face, cu * COARSE, cv * COARSE + r);
        tot += cell_start[k0 + COARSE] - cell_start[k0];
    }
    ctable[cc + 1] = tot;
}

template <typename T>
int index_build_t(const T *lon, const T *lat, int64_t n, uint8_t *valid, const prs_reduce_box *box,
                  const uint8_t *premask, const uint8_t *exclude, const uint8_t *cover, int flags, int k_hint,
                  double radius_hint, prs_index *ix, cudaStream_t st)
{
    (void)radius_hint;
    const int TPB = 256;
    BuildTimer timer;
    int rc = ensure_pool_configured();
    if (rc != PRS_OK) return rc;
    ix->stream = st;
    const int64_t n_blocks = (n + BUILD_TPB * BUILD_ITEMS - 1) / (BUILD_TPB * BUILD_ITEMS);
    const int compute_valid = (flags & PRS_BUILD_COMPUTE_VALID) != 0;
    prs_reduce_box bx0;
    memset(&bx0, 0, sizeof(bx0));
    if (box) bx0 = *box;
    BoxT<T> bx;
    make_box(bx0, bx);
    const size_t occ_n = 6ull * OCC_G * OCC_G;
    // workspace: counters | rect[24] | occ[occ_n] | block_counts[n_blocks] || unsorted records || keep flags
    const size_t off_counts = (256 + occ_n + 15) & ~(size_t)15;
    const size_t scratch_bytes = (off_counts + sizeof(uint2) * (size_t)n_blocks + 255) & ~(size_t)255;
    const size_t tmp_bytes = (sizeof(Rec<T>) * (size_t)n + 255) & ~(size_t)255;
    WorkspaceLease lease;
    void *ws_base = nullptr;
    rc = lease.acquire(scratch_bytes + tmp_bytes + (size_t)n, st, &ws_base);
    if (rc != PRS_OK) return rc;
    uint8_t *scratch = (uint8_t *)ws_base;
    PRS_CK(cudaMemsetAsync(scratch, 0, off_counts, st));
    BuildCounters *counters = (BuildCounters *)scratch;
    int *rect_dev = (int *)(scratch + 32);
    uint8_t *occ = scratch + 256;
    uint2 *block_counts = (uint2 *)(scratch + off_counts);
    {
        struct {
            BuildCounters c;
            uint32_t pad[4];
            int rect[24];
        } init;
        memset(&init, 0, sizeof(init));
        init.c.first_source = 0xFFFFFFFFu;
        for (int f = 0; f < 6; ++f) {
            init.rect[4 * f] = init.rect[4 * f + 1] = 1 << 30;
            init.rect[4 * f + 2] = init.rect[4 * f + 3] = -1;
        }
        PRS_CK(cudaMemcpyAsync(scratch, &init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    // pass 1: validity, selection, occupancy probe (records are addressed by source index: no rank table needed)
    Rec<T> *tmp = (Rec<T> *)(scratch + scratch_bytes);
    uint8_t *keep = scratch + scratch_bytes + tmp_bytes;
    build_select_kernel<T><<<(unsigned)n_blocks, BUILD_TPB, 0, st>>>(lon, lat, n, valid, compute_valid, bx, premask, exclude,
                                                                     cover, keep, occ, counters, block_counts);
    PRS_CKL();
    build_reduce_kernel<<<96, TPB, 0, st>>>(occ, block_counts, n_blocks, rect_dev, counters);
    PRS_CKL();
    timer.mark("pass 1 queued");
    // the one host round trip of the build: sizes the cell grid
    struct {
        BuildCounters c;
        uint32_t pad[4];
        int rect[24];
    } h;
    rc = read_back_small(scratch, &h, sizeof(h), st);
    if (rc != PRS_OK) return rc;
    timer.mark("counters read back");
    ix->n_valid = h.c.n_valid;
    ix->n_indexed = (int64_t)h.c.n_valid - (int64_t)h.c.n_skipped;
    ix->first_valid_source = h.c.first_source == 0xFFFFFFFFu ? 0u : h.c.first_source;
    ix->all_valid = (int64_t)h.c.n_valid == n;
    if ((flags & PRS_BUILD_RANKS) && !ix->all_valid && ix->n_valid > 0) {
        // rank of every source point among the valid ones = the index space of index_array (kd_tree.py:635)
        PRS_CK(cudaMallocAsync((void **)&ix->rank, sizeof(uint32_t) * (size_t)n, st));
        ix->bytes += sizeof(uint32_t) * (size_t)n;
        rc = scan_u32<uint8_t>(valid, ix->rank, n, 0, nullptr, st);
        if (rc != PRS_OK) return rc;
    }
    if (ix->n_indexed <= 0) {
        ix->n_indexed = 0;
        ix->G = COARSE;
        ix->Gc = 1;
        return PRS_OK;
    }
    // target points per cell: the hot path of the query scans the 3x3 block around a point's cell (9*lambda
    // candidates) and that block should hold the k nearest with room to spare
    double lambda = k_hint > 4 ? 0.5 * (double)k_hint : 2.0;
    if (const char *env = getenv("PRS_CELL_OCCUPANCY")) {
        double v = atof(env);
        if (v > 0) lambda = v;
    }
    uint32_t n_occ = h.c.n_occ ? h.c.n_occ : 1;
    double g = (double)OCC_G * sqrt((double)ix->n_indexed / ((double)n_occ * lambda));
    int G = (int)ceil(g / COARSE) * COARSE;
    int Gmax = 16384;
    if (const char *env = getenv("PRS_MAX_CELLS_PER_SIDE")) {
        int v = atoi(env);
        if (v >= COARSE) Gmax = (v / COARSE) * COARSE;
    }
    if (G < COARSE) G = COARSE;
    if (G > Gmax) G = Gmax;
    ix->G = G;
    ix->Gc = G / COARSE;
    // table rectangles: probe rect -> coarse cells (conservative, +1 probe cell of slack) -> fine cells
    int64_t nf = 0, nc = 0;
    for (int f = 0; f < 6; ++f) {
        const int *r = h.rect + 4 * f;
        ix->fine.base[f] = nf;
        ix->coarse.base[f] = nc;
        if (r[2] < r[0]) {
            ix->fine.fu0[f] = ix->fine.fv0[f] = ix->fine.fw[f] = ix->fine.fh[f] = 0;
            ix->coarse.fu0[f] = ix->coarse.fv0[f] = ix->coarse.fw[f] = ix->coarse.fh[f] = 0;
            continue;
        }
        auto lo = [&](int p) { long v = (long)floor((double)(p - 1) * ix->Gc / OCC_G); return (int)(v < 0 ? 0 : v); };
        auto hi = [&](int p) {
            long v = (long)ceil((double)(p + 2) * ix->Gc / OCC_G);
            return (int)(v > ix->Gc ? ix->Gc : v);
        };
        int cu0 = lo(r[0]), cv0 = lo(r[1]), cu1 = hi(r[2]), cv1 = hi(r[3]);  // [cu0, cu1) coarse cells
        if (cu1 <= cu0) cu1 = cu0 + 1;
        if (cv1 <= cv0) cv1 = cv0 + 1;
        ix->coarse.fu0[f] = cu0;
        ix->coarse.fv0[f] = cv0;
        ix->coarse.fw[f] = cu1 - cu0;
        ix->coarse.fh[f] = cv1 - cv0;
        ix->fine.fu0[f] = cu0 * COARSE;
        ix->fine.fv0[f] = cv0 * COARSE;
        ix->fine.fw[f] = (cu1 - cu0) * COARSE;
        ix->fine.fh[f] = (cv1 - cv0) * COARSE;
        nc += (int64_t)ix->coarse.fw[f] * ix->coarse.fh[f];
        nf += (int64_t)ix->fine.fw[f] * ix->fine.fh[f];
    }
    ix->n_cells = nf;
    ix->n_coarse = nc;
    if (nf >= 0xFFFFFFF0LL) return fail(PRS_ENOSUP, "cell table too large");
    // pass 2 + scan + pass 3
    // three leading pad words: cell_start + 1 (the array the scan walks) is then 16-byte aligned for 128-bit accesses
    PRS_CK(cudaMallocAsync((void **)&ix->cell_alloc, sizeof(uint32_t) * (size_t)(nf + 4), st));
    ix->cell_start = ix->cell_alloc + 3;
    ix->bytes += sizeof(uint32_t) * (size_t)(nf + 4);
    PRS_CK(cudaMemsetAsync(ix->cell_alloc, 0, sizeof(uint32_t) * (size_t)(nf + 4), st));
    timer.mark("cell table allocated");
    if (ix->n_indexed * 2 < n)  // a sharded rank (or a heavily masked source): most points are skipped
        build_xyz_kernel<T, true><<<(unsigned)n_blocks, BUILD_TPB, 0, st>>>(lon, lat, n, keep, G, ix->fine, tmp, ix->cell_start);
    else
        build_xyz_kernel<T, false><<<(unsigned)n_blocks, BUILD_TPB, 0, st>>>(lon, lat, n, keep, G, ix->fine, tmp, ix->cell_start);
    PRS_CKL();
    rc = scan_u32<uint32_t>(ix->cell_start + 1, ix->cell_start + 1, nf, 0, nullptr, st);
    if (rc != PRS_OK) return rc;
    size_t rec_bytes = sizeof(Rec<T>) * (size_t)ix->n_indexed;
    PRS_CK(cudaMallocAsync((void **)&ix->recs, rec_bytes, st));
    timer.mark("records allocated");
    ix->bytes += rec_bytes;
    const unsigned scatter_blocks = (unsigned)((n + BUILD_TPB * SCATTER_ITEMS - 1) / (BUILD_TPB * SCATTER_ITEMS));
    if (ix->n_indexed * 2 < n)
        build_scatter_kernel<T, true><<<scatter_blocks, BUILD_TPB, 0, st>>>(tmp, keep, n, ix->cell_start, (Rec<T> *)ix->recs);
    else
        build_scatter_kernel<T, false><<<scatter_blocks, BUILD_TPB, 0, st>>>(tmp, keep, n, ix->cell_start, (Rec<T> *)ix->recs);
    PRS_CKL();
    // coarse table
    PRS_CK(cudaMallocAsync((void **)&ix->coarse_alloc, sizeof(uint32_t) * (size_t)(nc + 4), st));
    ix->coarse_start = ix->coarse_alloc + 3;
    ix->bytes += sizeof(uint32_t) * (size_t)(nc + 4);
    PRS_CK(cudaMemsetAsync(ix->coarse_start, 0, sizeof(uint32_t), st));
    build_coarse_kernel<<<(unsigned)((nc + TPB - 1) / TPB), TPB, 0, st>>>(ix->cell_start, ix->fine, ix->coarse, nc,
                                                                         ix->coarse_start);
    PRS_CKL();
    rc = scan_u32<uint32_t>(ix->coarse_start + 1, ix->coarse_start + 1, nc, 1, nullptr, st);
    if (rc != PRS_OK) return rc;
    timer.mark("all queued");
    return PRS_OK;
}

}  // namespace prs
