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

// Description of a built index.  It lives in DEVICE memory (written by build_plan_kernel, read by every later build
// pass and by the query kernels), so that the build is queued without a host round trip; the host fetches a copy
// only when somebody asks (prs_index_info).
struct IndexMeta {
    uint32_t n_valid;             // valid source points (pykdtree's .n)
    uint32_t n_skipped;           // valid but excluded (query mask) or outside this rank's target coverage
    uint32_t first_valid_source;  // raveled index of rank 0 (kd_tree.py:754-759 gathers new_data[0] for missing)
    uint32_t n_occ;               // occupied probe cells
    int32_t G, Gc;                // fine / coarse cells per face side (G a multiple of COARSE)
    int32_t all_valid;            // rank == source index
    int32_t status;               // 0 ok; 1: the cell table did not fit its allocation even at the coarsest grid
    int64_t n_indexed;            // valid and not skipped
    int64_t n_cells, n_coarse;
    int64_t n_source;
    FaceTable fine, coarse;
};

}  // namespace prs

struct prs_index_uses;

struct prs_index {
    int tree_dtype;      // PRS_F32 / PRS_F64
    int64_t n_source;    // raveled source size
    void *recs;          // capacity n_source records
    uint32_t *cell_start;    // = cell_alloc + 3
    uint32_t *coarse_start;  // = coarse_alloc + 3
    uint32_t *cell_alloc, *coarse_alloc;
    uint32_t *rank;      // only with PRS_BUILD_RANKS
    int64_t cap_cells, cap_coarse;
    prs::IndexMeta *meta_dev;
    prs::IndexMeta meta;  // host copy, valid once `fetched`
    int fetched;
    cudaEvent_t planned;  // recorded behind build_plan_kernel: meta_dev is final from there on
    size_t bytes;
    cudaStream_t stream;  // stream the memory was allocated on (stream-ordered free)
    // queries launched on OTHER streams: one event per stream, recorded behind the stream's latest query, so that
    // prs_index_destroy can order the stream-ordered free behind them (see index_note_use / prs_index_destroy)
    prs_index_uses *uses;
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

#ifdef PRS_WITH_INDEX_BUILD  // the build itself is only compiled into prs_b200.cu
struct BuildCounters {   // zeroed before the build; every field is raised from 0 with atomicAdd / atomicMax
    uint32_t n_valid;    // totals of pass 1 (reduced from the per-block records by build_plan_kernel)
    uint32_t n_skipped;
    uint32_t first_inv;  // ~(raveled index of the first valid point), 0 = none
    uint32_t n_occ;      // occupied probe cells
    uint32_t rect[24];   // per face: OCC_G - u0, OCC_G - v0, u1 + 1, v1 + 1 of the occupied probe cells (0 = none)
    uint32_t done;       // blocks of build_plan_kernel that have finished their share
};

// Host copy of the index description, fetched on first use: a copy on a side stream that only waits for the plan
// kernel (not for whatever was queued behind the build).
inline int index_fetch_meta(const prs_index *cix)
{
    prs_index *ix = const_cast<prs_index *>(cix);
    if (!ix) return fail(PRS_EINVAL, "null index");
    if (ix->fetched) return PRS_OK;
    static std::mutex lock;
    static cudaStream_t side[64] = {nullptr};
    std::lock_guard<std::mutex> guard(lock);
    if (ix->fetched) return PRS_OK;
    int dev = 0;
    PRS_CK(cudaGetDevice(&dev));
    const int slot = dev < 0 || dev >= 64 ? 0 : dev;
    if (!side[slot]) PRS_CK(cudaStreamCreateWithFlags(&side[slot], cudaStreamNonBlocking));
    PRS_CK(cudaStreamWaitEvent(side[slot], ix->planned, 0));
    PRS_CK(cudaMemcpyAsync(&ix->meta, ix->meta_dev, sizeof(IndexMeta), cudaMemcpyDeviceToHost, side[slot]));
    PRS_CK(cudaStreamSynchronize(side[slot]));
    ix->fetched = 1;
    if (ix->meta.status != 0) return fail(PRS_ENOSUP, "cell table too large for its allocation");
    return PRS_OK;
}

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
// points that are kept, once the cell size is known.  Tin: dtype of the lon/lat arrays.
template <typename Tin>
__global__ void __launch_bounds__(BUILD_TPB, PRS_SELECT_MIN_BLOCKS)
    build_select_kernel(const Tin *__restrict__ lon, const Tin *__restrict__ lat, int64_t n, uint8_t *__restrict__ valid,
                        int compute_valid, const BoxT<Tin> box, const uint8_t *__restrict__ premask,
                        const uint8_t *__restrict__ exclude, const uint8_t *__restrict__ cover,
                        uint8_t *__restrict__ keep, uint8_t *__restrict__ occ, BuildCounters *counters,
                        uint2 *__restrict__ block_counts)
{
    // a warp works on 32 consecutive points at a time (their cells mostly coincide, so its occupancy stores and
    // the histogram atomics of pass 2 collapse into one or two transactions); all loads are issued first
    const int64_t base = (int64_t)blockIdx.x * (BUILD_TPB * BUILD_ITEMS) + threadIdx.x;
    int nv = 0, ns = 0;
    uint32_t first = 0xFFFFFFFFu;
    Tin lo[BUILD_ITEMS], la[BUILD_ITEMS];
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
        const bool v = compute_valid ? valid_predicate<Tin>(lo[j], la[j], box, vin[j] != 0) : vin[j] != 0;
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
                    if ((int)(ax >= lim) + (int)(ay >= lim) + (int)(az >= lim) > 1) {  // rare: within 1e-4 of an edge
#pragma unroll 1
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
                    }
                    kept = true;
                }
            }
            if (!kept) ++ns;
        }
        keep[i] = kept ? 1 : 0;
    }
    // block totals (one record per block, reduced by build_plan_kernel: no same-address atomics storm)
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
        if (tf != 0xFFFFFFFFu && ~tf > *(volatile uint32_t *)&counters->first_inv) atomicMax(&counters->first_inv, ~tf);
    }
}

struct PlanParams {
    int64_t n, n_blocks, cap_cells;
    double lambda;  // target points per cell
    int g_max;      // upper limit of cells per face side (a multiple of COARSE)
};

// Table geometry for G cells per side from the per-face bounding rectangles (in probe cells) of the occupied probe
// cells: probe rect -> coarse cells (conservative, +1 probe cell of slack) -> fine cells.  Returns the number of
// fine cells.
__device__ inline int64_t plan_tables(int G, const int rect[24], FaceTable &fine, FaceTable &coarse, int64_t &n_coarse)
{
    const int Gc = G / COARSE;
    int64_t nf = 0, nc = 0;
    for (int f = 0; f < 6; ++f) {
        const int *r = rect + 4 * f;
        fine.base[f] = nf;
        coarse.base[f] = nc;
        if (r[2] < r[0]) {
            fine.fu0[f] = fine.fv0[f] = fine.fw[f] = fine.fh[f] = 0;
            coarse.fu0[f] = coarse.fv0[f] = coarse.fw[f] = coarse.fh[f] = 0;
            continue;
        }
        // [cu0, cu1) coarse cells
        int cu0 = max(0, (int)floor((double)(r[0] - 1) * Gc / OCC_G)), cv0 = max(0, (int)floor((double)(r[1] - 1) * Gc / OCC_G));
        int cu1 = min(Gc, (int)ceil((double)(r[2] + 2) * Gc / OCC_G)), cv1 = min(Gc, (int)ceil((double)(r[3] + 2) * Gc / OCC_G));
        if (cu1 <= cu0) cu1 = cu0 + 1;
        if (cv1 <= cv0) cv1 = cv0 + 1;
        coarse.fu0[f] = cu0;
        coarse.fv0[f] = cv0;
        coarse.fw[f] = cu1 - cu0;
        coarse.fh[f] = cv1 - cv0;
        fine.fu0[f] = cu0 * COARSE;
        fine.fv0[f] = cv0 * COARSE;
        fine.fw[f] = (cu1 - cu0) * COARSE;
        fine.fh[f] = (cv1 - cv0) * COARSE;
        nc += (int64_t)(coarse.fw[f] + 1) * coarse.fh[f];  // a row holds fw + 1 row-local prefix sums
        nf += (int64_t)fine.fw[f] * fine.fh[f];
    }
    n_coarse = nc;
    return nf;
}

// The plan: totals of pass 1, bounding rectangles of the occupancy probe (reduced by all blocks), then - in the block
// that finishes last - the cell size (a cell should hold ~lambda points, so that the 3x3 block a query scans first
// holds the k nearest with room to spare) and the table geometry.  What index_build used to do on the host after a
// read-back.
static __global__ void __launch_bounds__(256)
    build_plan_kernel(const uint8_t *__restrict__ occ, const uint2 *__restrict__ block_counts, BuildCounters *counters,
                      const PlanParams pp, IndexMeta *__restrict__ meta)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    unsigned nv = 0, ns = 0, no = 0;
    for (int64_t b = tid; b < pp.n_blocks; b += nth) {
        const uint2 c = block_counts[b];
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
            atomicMax(&counters->rect[4 * f + 0], (uint32_t)(OCC_G - u0));
            atomicMax(&counters->rect[4 * f + 1], (uint32_t)(OCC_G - iv));
            atomicMax(&counters->rect[4 * f + 2], (uint32_t)(u1 + 1));
            atomicMax(&counters->rect[4 * f + 3], (uint32_t)(iv + 1));
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
    // the block that finishes last sees everybody's contributions
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&counters->done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    volatile BuildCounters *vc = counters;
    IndexMeta m;
    m.n_valid = vc->n_valid;
    m.n_skipped = vc->n_skipped;
    m.first_valid_source = vc->first_inv ? ~vc->first_inv : 0u;
    m.n_occ = vc->n_occ;
    m.all_valid = (int64_t)m.n_valid == pp.n;
    m.status = 0;
    m.n_source = pp.n;
    m.n_indexed = (int64_t)m.n_valid - (int64_t)m.n_skipped;
    int rect[24];
    for (int f = 0; f < 6; ++f) {
        const uint32_t e0 = vc->rect[4 * f], e1 = vc->rect[4 * f + 1], e2 = vc->rect[4 * f + 2], e3 = vc->rect[4 * f + 3];
        const bool any = e2 != 0 && m.n_indexed > 0;
        rect[4 * f + 0] = any ? OCC_G - (int)e0 : 1;
        rect[4 * f + 1] = any ? OCC_G - (int)e1 : 1;
        rect[4 * f + 2] = any ? (int)e2 - 1 : 0;
        rect[4 * f + 3] = any ? (int)e3 - 1 : 0;
    }
    if (m.n_indexed < 0) m.n_indexed = 0;
    const double n_occ = m.n_occ ? (double)m.n_occ : 1.0;
    const double g = (double)OCC_G * sqrt((double)m.n_indexed / (n_occ * pp.lambda));
    int G = (int)ceil(g / COARSE) * COARSE;
    if (G < COARSE) G = COARSE;
    if (G > pp.g_max) G = pp.g_max;
    int64_t nc = 0;
    int64_t nf = plan_tables(G, rect, m.fine, m.coarse, nc);
    // a source made of clusters far apart on one face spans a rectangle that is mostly empty: coarsen the grid
    // until the table fits its allocation (4 entries per source point)
    while (nf > pp.cap_cells && G > COARSE) {
        G = max(COARSE, (G / 2 / COARSE) * COARSE);
        nf = plan_tables(G, rect, m.fine, m.coarse, nc);
    }
    if (nf > pp.cap_cells) {  // cannot happen for cap_cells >= 6 * COARSE^2: reported when the host asks
        m.status = 1;
        m.n_indexed = 0;
        for (int i = 0; i < 24; ++i) rect[i] = (i & 3) < 2 ? 1 : 0;
        nf = plan_tables(G, rect, m.fine, m.coarse, nc);
    }
    m.G = G;
    m.Gc = G / COARSE;
    m.n_cells = nf;
    m.n_coarse = nc;
    *meta = m;
}

// zero table[0 .. n_cells + 4): the size is only known on the device
static __global__ void build_clear_kernel(const IndexMeta *__restrict__ meta, uint32_t *__restrict__ cell_alloc)
{
    const int64_t nf = meta->n_cells + 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    // cell_alloc is 16-byte aligned: whole uint4 stores, the tail word by word
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < nf; i += stride) {
        if (i + 4 <= nf)
            *reinterpret_cast<uint4 *>(cell_alloc + i) = make_uint4(0, 0, 0, 0);
        else
            for (int64_t j = i; j < nf; ++j) cell_alloc[j] = 0;
    }
}

__device__ __forceinline__ int64_t table_key(const FaceTable &t, int face, int iu, int iv)
{
    return t.base[face] + (int64_t)(iv - t.fv0[face]) * t.fw[face] + (iu - t.fu0[face]);
}

// ECEF of one source point in the tree dtype T from lon/lat of dtype Tin.  Tin == T: the reference's chain in that
// dtype.  float32 lon/lat into a float64 tree (XArrayResamplerNN, kd_tree.py:970-981): the reference transforms in
// float32 (lonlat2xyz on float32 arrays) and only then widens (`astype(np.float64)`) - so does this.
template <typename Tin, typename T>
__device__ __forceinline__ void source_xyz(Tin lo, Tin la, T &x, T &y, T &z)
{
    Tin xi, yi, zi;
    lonlat_to_xyz(lo, la, xi, yi, zi);
    x = (T)xi, y = (T)yi, z = (T)zi;
}

// Pass 2: exact ECEF coordinates of the kept points, their cell key and the histogram into table[key + 1].
// tmp[i] = (x, y, z, cell key) - the record's source index is its position i, so the fourth word carries the key
// to pass 3.  sparse (a rank of a sharded run keeps a fraction of the points; decided on the host: a coverage mask
// was given): the coordinates are only loaded for kept points and tmp is only written for them; otherwise all loads
// are issued up front.
template <typename Tin, typename T, bool sparse>
__global__ void __launch_bounds__(BUILD_TPB)
    build_xyz_kernel(const Tin *__restrict__ lon, const Tin *__restrict__ lat, int64_t n,
                     const uint8_t *__restrict__ keep, const IndexMeta *__restrict__ meta,
                     Rec<T> *__restrict__ tmp, uint32_t *__restrict__ table)
{
    // table geometry: one word per thread, one round trip to L2 per block
    __shared__ FaceTable ft;
    static_assert(sizeof(FaceTable) % 4 == 0 && sizeof(FaceTable) / 4 <= BUILD_TPB, "copied word by word");
    if (threadIdx.x < sizeof(FaceTable) / 4)
        reinterpret_cast<uint32_t *>(&ft)[threadIdx.x] = reinterpret_cast<const uint32_t *>(&meta->fine)[threadIdx.x];
    const int G = meta->G;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (BUILD_TPB * BUILD_ITEMS) + threadIdx.x;
    Tin lo[BUILD_ITEMS], la[BUILD_ITEMS];
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
        if (i < n && (!sparse || kp[j])) {
            lo[j] = lon[i];
            la[j] = lat[i];
        }
    }
#pragma unroll
    for (int j = 0; j < BUILD_ITEMS; ++j) {
        const int64_t i = base + (int64_t)j * BUILD_TPB;
        if (!kp[j]) {
            // dense builds mark the few skipped points in tmp, so that pass 3 needs nothing but tmp
            if (!sparse && i < n) store_rec(tmp + i, (T)0, (T)0, (T)0, CELL_SKIP);
            continue;
        }
        T x, y, z;
        source_xyz<Tin, T>(lo[j], la[j], x, y, z);
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
template <typename T, bool sparse>
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
            if (sparse)
                kp[j] = keep[i];
            else
                r[j] = load_rec(tmp + i);
        }
    }
    if (sparse) {
#pragma unroll
        for (int j = 0; j < SCATTER_ITEMS; ++j)
            if (kp[j]) r[j] = load_rec(tmp + base + (int64_t)j * BUILD_TPB);
    }
    // consecutive source points mostly share a cell: one atomic per (warp, cell) instead of one per point
    // (the atomics of the thread's items are issued before any result is used: their round trips overlap)
    const unsigned lane = threadIdx.x & 31u;
    uint32_t pos[SCATTER_ITEMS], first[SCATTER_ITEMS];
    unsigned peers[SCATTER_ITEMS];
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j) peers[j] = __match_any_sync(0xffffffffu, r[j].idx);
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j) {
        first[j] = 0;
        if (lane == (unsigned)(__ffs(peers[j]) - 1) && r[j].idx != CELL_SKIP)
            first[j] = atomicAdd(&table[r[j].idx + 1], (uint32_t)__popc(peers[j]));
    }
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j)
        pos[j] = __shfl_sync(0xffffffffu, first[j], __ffs(peers[j]) - 1) + __popc(peers[j] & ((1u << lane) - 1u));
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j)
        if (r[j].idx != CELL_SKIP)
            store_rec(recs + pos[j], r[j].x, r[j].y, r[j].z, (uint32_t)(base + (int64_t)j * BUILD_TPB));
}

// Coarse table: number of points in each COARSE x COARSE block of fine cells, stored as ROW-LOCAL prefix sums
// (fw + 1 entries per row of coarse cells), so that one block per row computes counts and prefix sums in one pass
// and no device-wide scan is needed.  Count of the coarse cells [u0, u1] of a row: row[u1 + 1] - row[u0].
constexpr int COARSE_TPB = 256;
static __global__ void __launch_bounds__(COARSE_TPB)
    build_coarse_kernel(const uint32_t *__restrict__ cell_start, const IndexMeta *__restrict__ meta,
                        uint32_t *__restrict__ ctable)
{
    __shared__ FaceTable fine, coarse;
    __shared__ uint32_t warp_tot[COARSE_TPB / 32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x < sizeof(FaceTable) / 4) {
        reinterpret_cast<uint32_t *>(&fine)[threadIdx.x] = reinterpret_cast<const uint32_t *>(&meta->fine)[threadIdx.x];
        reinterpret_cast<uint32_t *>(&coarse)[threadIdx.x] = reinterpret_cast<const uint32_t *>(&meta->coarse)[threadIdx.x];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int rows_total = 0;
#pragma unroll
    for (int f = 0; f < 6; ++f) rows_total += coarse.fh[f];
    for (int r = blockIdx.x; r < rows_total; r += gridDim.x) {
        int face = 0, local = r;
#pragma unroll
        for (int f = 0; f < 5; ++f)
            if (face == f && local >= coarse.fh[f]) {
                local -= coarse.fh[f];
                face = f + 1;
            }
        const int fw = coarse.fw[face];
        uint32_t *row = ctable + coarse.base[face] + (int64_t)local * (fw + 1);
        const int cv = coarse.fv0[face] + local;
        if (threadIdx.x == 0) {
            row[0] = 0;
            carry_s = 0;
        }
        __syncthreads();
        for (int i0 = 0; i0 < fw; i0 += COARSE_TPB) {
            const int i = i0 + threadIdx.x;
            uint32_t tot = 0;
            if (i < fw) {
                const int cu = coarse.fu0[face] + i;
                int64_t k0 = table_key(fine, face, cu * COARSE, cv * COARSE);
#pragma unroll 4
                for (int q = 0; q < COARSE; ++q, k0 += fine.fw[face]) tot += cell_start[k0 + COARSE] - cell_start[k0];
            }
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            uint32_t before = carry_s;
            for (int w = 0; w < warp; ++w) before += warp_tot[w];
            if (i < fw) row[i + 1] = before + incl;
            __syncthreads();
            if (threadIdx.x == COARSE_TPB - 1) carry_s = before + incl;
            __syncthreads();
        }
    }
}

inline int64_t index_cap_cells(int64_t n) { return 4 * n + 65536; }

// Queue the whole build on `st`.  Nothing here waits for the device: the cell size and the table geometry are
// decided by build_plan_kernel and stay in device memory (IndexMeta); the arrays are allocated for the largest
// size they can take (records: one per source point; cell table: 4 entries per source point, the plan coarsens
// the grid to fit), so their sizes - and the stream-ordered pool's reuse of them - do not depend on the data.
template <typename Tin, typename T>
int index_build_t(const Tin *lon, const Tin *lat, int64_t n, uint8_t *valid, const prs_reduce_box *box,
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
    BoxT<Tin> bx;
    make_box(bx0, bx);
    const size_t occ_n = 6ull * OCC_G * OCC_G;
    // workspace: counters | occ[occ_n] | block_counts[n_blocks] || unsorted records || keep flags
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
    uint8_t *occ = scratch + 256;
    uint2 *block_counts = (uint2 *)(scratch + off_counts);
    Rec<T> *tmp = (Rec<T> *)(scratch + scratch_bytes);
    uint8_t *keep = scratch + scratch_bytes + tmp_bytes;
    // long-lived arrays, sized for the worst case
    ix->cap_cells = index_cap_cells(n);
    ix->cap_coarse = ix->cap_cells / (COARSE * COARSE) + 6 * 1024 + 8;  // + one entry per row of coarse cells
    PRS_CK(cudaMallocAsync((void **)&ix->meta_dev, sizeof(IndexMeta), st));
    PRS_CK(cudaMallocAsync((void **)&ix->cell_alloc, sizeof(uint32_t) * (size_t)(ix->cap_cells + 8), st));
    PRS_CK(cudaMallocAsync((void **)&ix->coarse_alloc, sizeof(uint32_t) * (size_t)(ix->cap_coarse + 8), st));
    PRS_CK(cudaMallocAsync((void **)&ix->recs, sizeof(Rec<T>) * (size_t)n, st));
    // three leading pad words: cell_start + 1 (the array the scan walks) is then 16-byte aligned for 128-bit accesses
    ix->cell_start = ix->cell_alloc + 3;
    ix->coarse_start = ix->coarse_alloc + 3;
    ix->bytes = sizeof(IndexMeta) + sizeof(uint32_t) * (size_t)(ix->cap_cells + ix->cap_coarse + 16) + sizeof(Rec<T>) * (size_t)n;
    if (flags & PRS_BUILD_RANKS) {
        PRS_CK(cudaMallocAsync((void **)&ix->rank, sizeof(uint32_t) * (size_t)n, st));
        ix->bytes += sizeof(uint32_t) * (size_t)n;
    }
    timer.mark("allocated");
    // pass 1: validity, selection, occupancy probe (records are addressed by source index: no rank table needed)
    build_select_kernel<Tin><<<(unsigned)n_blocks, BUILD_TPB, 0, st>>>(lon, lat, n, valid, compute_valid, bx, premask, exclude,
                                                                       cover, keep, occ, counters, block_counts);
    PRS_CKL();
    // the plan: cell size and table geometry, on the device
    PlanParams pp;
    pp.n = n;
    pp.n_blocks = n_blocks;
    pp.cap_cells = ix->cap_cells;
    // target points per cell: the hot path of the query scans the 3x3 block around a point's cell (9*lambda
    // candidates) and that block should hold the k nearest with room to spare
    // measured (profiles/r2_experiments.txt): k = 8 is fastest at 6 points per cell (C3 4.84 ms vs 5.17 at 4 and 4.95 at
    // 8), k = 16 at 8 (C4 53.3 ms vs 54.8 at 12), k = 1 at 2 (C5 6.6 ms vs 7.1 at 3)
    pp.lambda = k_hint > 8 ? 0.5 * (double)k_hint : (k_hint > 4 ? 0.75 * (double)k_hint : 2.0);
    if (const char *env = getenv("PRS_CELL_OCCUPANCY")) {
        double v = atof(env);
        if (v > 0) pp.lambda = v;
    }
    pp.g_max = 16384;
    if (const char *env = getenv("PRS_MAX_CELLS_PER_SIDE")) {
        int v = atoi(env);
        if (v >= COARSE) pp.g_max = (v / COARSE) * COARSE;
    }
    build_plan_kernel<<<96, 256, 0, st>>>(occ, block_counts, counters, pp, ix->meta_dev);
    PRS_CKL();
    PRS_CK(cudaEventRecord(ix->planned, st));
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        PRS_CK(cudaGetDevice(&dev));
        PRS_CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    build_clear_kernel<<<n_sm * 8, TPB, 0, st>>>(ix->meta_dev, ix->cell_alloc);
    PRS_CKL();
    if (ix->rank) {
        // rank of every source point among the valid ones = the index space of index_array (kd_tree.py:635)
        rc = scan_u32<uint8_t>(valid, ix->rank, n, 0, nullptr, st);
        if (rc != PRS_OK) return rc;
    }
    // pass 2 + scan + pass 3
    const bool sparse = cover != nullptr;  // a rank of a sharded run: most points are skipped
    if (sparse)
        build_xyz_kernel<Tin, T, true><<<(unsigned)n_blocks, BUILD_TPB, 0, st>>>(lon, lat, n, keep, ix->meta_dev, tmp, ix->cell_start);
    else
        build_xyz_kernel<Tin, T, false><<<(unsigned)n_blocks, BUILD_TPB, 0, st>>>(lon, lat, n, keep, ix->meta_dev, tmp, ix->cell_start);
    PRS_CKL();
    rc = scan_u32_dev(ix->cell_start + 1, ix->cap_cells, &ix->meta_dev->n_cells, 0, st);
    if (rc != PRS_OK) return rc;
    const unsigned scatter_blocks = (unsigned)((n + BUILD_TPB * SCATTER_ITEMS - 1) / (BUILD_TPB * SCATTER_ITEMS));
    if (sparse)
        build_scatter_kernel<T, true><<<scatter_blocks, BUILD_TPB, 0, st>>>(tmp, keep, n, ix->cell_start, (Rec<T> *)ix->recs);
    else
        build_scatter_kernel<T, false><<<scatter_blocks, BUILD_TPB, 0, st>>>(tmp, keep, n, ix->cell_start, (Rec<T> *)ix->recs);
    PRS_CKL();
    // coarse table (counts + row-local prefix sums in one kernel)
    build_coarse_kernel<<<n_sm * 8, COARSE_TPB, 0, st>>>(ix->cell_start, ix->meta_dev, ix->coarse_start);
    PRS_CKL();
    timer.mark("all queued");
    return PRS_OK;
}

#endif  // PRS_WITH_INDEX_BUILD

}  // namespace prs
