// Bounded exact k-NN against the cube-face cell index, with the sample gather / weighting fused into the
// epilogue.  Replaces KDTree.query (kd_tree.py:545-548) + get_sample_from_neighbour_info (kd_tree.py:566-818).
//
// Work decomposition: the target is cut into tiles of 32 columns x QROWS rows.
//   A. classify_kernel, one warp per tile: bound the tile's points by a spherical cap and ask the L2-resident
//      coarse table whether ANY source point lies within (tile cap + radius).  Wide target grids are mostly far
//      from the swath: such tiles get their final result (fill / sentinels) streamed out right away;
//   B. query_kernel, one block per remaining tile, one thread per point: the block stages the candidate records of
//      the tile in shared memory with bulk copies (cp.async.bulk + mbarrier), every thread scans the 3x3 block of
//      cells around its own point from there, then proves the result or visits exactly the cells that can still
//      beat the k-th distance (general search, global memory);
//   C. epilogue: neighbour info, or the gathered / weighted samples (nearest rows go through shared memory
//      so that the warp writes contiguous 128-byte segments).
#pragma once
#include "prs_common.cuh"
#include "prs_index.cuh"
#include "prs_proj.cuh"
#include <type_traits>

namespace prs {

// PRS_STATS: event counters for kernel experiments (tools/repro_stats.py reads them through prs_debug_stats)
#ifdef PRS_STATS
__device__ unsigned long long g_stats[32];
#define PRS_STAT(i, v) atomicAdd(&g_stats[i], (unsigned long long)(v))
#else
#define PRS_STAT(i, v) \
    do {               \
    } while (0)
#endif

#ifndef PRS_QROWS
#define PRS_QROWS 8
#endif
constexpr int QROWS = PRS_QROWS;    // target rows per tile
constexpr int QWARPS = 8;           // warps per block of the classification kernel
constexpr int STAGE_ROW_BYTES = 64; // nearest rows up to this size are staged through shared memory
// Resident blocks per SM asked of the compiler for query_kernel (32 * QROWS threads): the search is latency bound,
// so occupancy beats registers up to k = 8 (64 registers, measured on C2/C3/C5); wider top-k lists get more.
#ifndef PRS_QUERY_MIN_BLOCKS
#define PRS_QUERY_MIN_BLOCKS(K) (((K) <= 4 ? 32 : (K) <= 8 ? 24 : 16) / PRS_QROWS)
#endif

// Records fetched together in the candidate loop of the hot path.
#ifndef PRS_SCAN_BATCH
#define PRS_SCAN_BATCH(K) ((K) <= 4 ? 4 : 2)
#endif

#ifndef PRS_TOPK_BUBBLE
#define PRS_TOPK_BUBBLE 0
#endif

// ------------------------------------------------------------------------------------------ top-k list
// Sorted ascending by (d2, source index): ties go to the lowest source index (BASELINE north_star).
template <typename T, int K> struct TopK {
    T d[K];
    uint32_t id[K];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            d[i] = Ar<T>::inf();
            id[i] = PRS_NONE;
        }
    }
    // (a, ia) before (b, ib) in the list order.  float: squared distances are >= +0 and never NaN, so their bit
    // patterns order like unsigned integers and (bits << 32 | index) compares both fields at once.
    static __device__ __forceinline__ bool before(T a, uint32_t ia, T b, uint32_t ib)
    {
        if (sizeof(T) == 4) {
            const unsigned long long ka = ((unsigned long long)__float_as_uint((float)a) << 32) | ia;
            const unsigned long long kb = ((unsigned long long)__float_as_uint((float)b) << 32) | ib;
            return ka < kb;
        }
        return a < b || (a == b && ia < ib);
    }
    // CHECK_DUP = false where every record is seen at most once (the first walk over the 3x3 block)
    template <bool CHECK_DUP = true> __device__ __forceinline__ void offer(T d2, uint32_t idx, T dub)
    {
        if (!(d2 < dub)) return;  // strict radius test in the tree dtype (pykdtree)
        if (before(d2, idx, d[K - 1], id[K - 1])) {
            if (K > 1 && CHECK_DUP) {  // a cell may be visited twice (overlapping phases): never list a point twice
                bool dup = false;
#pragma unroll
                for (int j = 0; j < K - 1; ++j) dup = dup || (id[j] == idx);
                if (dup) return;
            }
#if PRS_TOPK_BUBBLE
            d[K - 1] = d2;
            id[K - 1] = idx;
#pragma unroll
            for (int j = K - 1; j > 0; --j) {
                bool sw = before(d[j], id[j], d[j - 1], id[j - 1]);
                T td = d[j - 1];
                uint32_t ti = id[j - 1];
                d[j - 1] = sw ? d[j] : td;
                id[j - 1] = sw ? id[j] : ti;
                d[j] = sw ? td : d[j];
                id[j] = sw ? ti : id[j];
            }
#else
            // every slot decides for itself (no chain of dependent swaps): slot j takes its upper neighbour when
            // the candidate goes above that neighbour, the candidate when it goes exactly here, else keeps its entry
            bool lt[K];
#pragma unroll
            for (int j = 0; j < K; ++j) lt[j] = before(d2, idx, d[j], id[j]);
#pragma unroll
            for (int j = K - 1; j > 0; --j) {
                const T nd = lt[j] ? d2 : d[j];
                const uint32_t ni = lt[j] ? idx : id[j];
                d[j] = lt[j - 1] ? d[j - 1] : nd;
                id[j] = lt[j - 1] ? id[j - 1] : ni;
            }
            d[0] = lt[0] ? d2 : d[0];
            id[0] = lt[0] ? idx : id[0];
#endif
        }
    }
    __device__ __forceinline__ bool full() const { return id[K - 1] != PRS_NONE; }

    // ---- branch-free bulk insertion (wide lists): a chunk of CH candidates, none of them listed yet, is sorted by a
    // bitonic network and merged into the list by a bitonic merge.  Every lane of a warp does the same work whatever
    // its data, whereas offer() makes the whole warp walk the insertion code whenever ANY lane accepts a candidate -
    // for k = 16 that was 60 % of the search's instructions at a third of the lanes active.
    // (a, ia) / (b, ib) -> ordered so that the first is before the second when up, the reverse otherwise
    static __device__ __forceinline__ void cex(T &a, uint32_t &ia, T &b, uint32_t &ib, bool up)
    {
        const bool sw = up ? before(b, ib, a, ia) : before(a, ia, b, ib);
        const T ta = sw ? b : a, tb = sw ? a : b;
        const uint32_t tia = sw ? ib : ia, tib = sw ? ia : ib;
        a = ta, b = tb, ia = tia, ib = tib;
    }
    // cd / ci: CH candidates (CH a power of two <= K); entries that are no candidates hold (inf, PRS_NONE)
    template <int CH> __device__ __forceinline__ void merge_chunk(T (&cd)[CH], uint32_t (&ci)[CH])
    {
        static_assert(CH <= K && (CH & (CH - 1)) == 0 && (K & (K - 1)) == 0, "power-of-two sizes");
        // bitonic sort of the chunk, DESCENDING
#pragma unroll
        for (int k = 2; k <= CH; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int l = i ^ j;
                    if (l > i) cex(cd[i], ci[i], cd[l], ci[l], (i & k) != 0);
                }
            }
        }
        // list ascending, chunk descending (aligned with the END of the list): the element-wise minimum of the two holds
        // the K smallest of their union as a bitonic sequence ...
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int t = K - CH + i;
            const bool take = before(cd[i], ci[i], d[t], id[t]);
            d[t] = take ? cd[i] : d[t];
            id[t] = take ? ci[i] : id[t];
        }
        // ... which a bitonic merge sorts (ascending)
#pragma unroll
        for (int j = K >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const int l = i ^ j;
                if (l > i) cex(d[i], id[i], d[l], id[l], true);
            }
        }
    }
};

// What the kernels know about the index: the device-resident description (IndexMeta, copied into shared memory by
// every block - the host never needs to have seen it) plus the arrays.
template <typename T> struct IndexView : IndexMeta {
    const Rec<T> *recs;
    const uint32_t *cell_start;
    const uint32_t *coarse_start;
    const uint32_t *rank;  // NULL when the index was built without PRS_BUILD_RANKS
};
template <typename T> struct IndexRef {  // kernel parameter
    const Rec<T> *recs;
    const uint32_t *cell_start;
    const uint32_t *coarse_start;
    const uint32_t *rank;
    const IndexMeta *meta;
};
// All threads of the block call this (it ends with a barrier).
template <typename T> __device__ __forceinline__ void load_index_view(IndexView<T> &ix, const IndexRef<T> &ref)
{
    const uint32_t *src = reinterpret_cast<const uint32_t *>(ref.meta);
    uint32_t *dst = reinterpret_cast<uint32_t *>(static_cast<IndexMeta *>(&ix));
    static_assert(sizeof(IndexMeta) % 4 == 0, "IndexMeta is copied word by word");
    for (int i = threadIdx.x; i < (int)(sizeof(IndexMeta) / 4); i += blockDim.x) dst[i] = __ldg(src + i);
    if (threadIdx.x == 0) {
        ix.recs = ref.recs;
        ix.cell_start = ref.cell_start;
        ix.coarse_start = ref.coarse_start;
        ix.rank = ref.rank;
    }
    __syncthreads();
}

struct CellRect {
    int u0, u1, v0, v1;
};

// Intersect with the part of the face the table spans; false when nothing is left.
__device__ __forceinline__ bool clip_rect(const FaceTable &t, int face, CellRect &r)
{
    if (t.fw[face] == 0) return false;
    r.u0 = max(r.u0, t.fu0[face]);
    r.u1 = min(r.u1, t.fu0[face] + t.fw[face] - 1);
    r.v0 = max(r.v0, t.fv0[face]);
    r.v1 = min(r.v1, t.fv0[face] + t.fh[face] - 1);
    return r.u0 <= r.u1 && r.v0 <= r.v1;
}

// Gnomonic bounding box (u0, u1, v0, v1) on `face` of the cap around the unit vector (fx, fy, fz) whose angular
// radius theta satisfies sin^2(theta) <= s2 (s = sqrt(s2)).  Returns false when the cap cannot touch the face.
// Closed form: the gnomonic image of a spherical cap is the conic w^T (q q^T - cos^2 I) w = 0, w = (1, u, v), with
//   u = (qa*qb -/+ s*sqrt(qa^2 + qb^2 - s^2)) / (qa^2 - s^2)      (same for v with qc),
// qa the component along the face normal, qb / qc along the face's u / v axes.  Bounds carry float32 slack
// (cell keys were computed from float32 coordinates, <= 1e-7 relative) and are NOT clamped to [-1, 1].
__device__ __forceinline__ bool face_bounds(int face, float fx, float fy, float fz, float s2, float s, float b[4])
{
    float qa, qb, qc;
    int axis = face >> 1;
    float sign = (face & 1) ? -1.0f : 1.0f;
    if (axis == 0) {
        qa = sign * fx;
        qb = fy;
        qc = fz;
    } else if (axis == 1) {
        qa = sign * fy;
        qb = fx;
        qc = fz;
    } else {
        qa = sign * fz;
        qb = fx;
        qc = fy;
    }
    // a point of the cap on this face has its dominant component >= 1/sqrt(3); theta <= 1.2*s for s <= 0.8
    if (qa < 0.57735f - 1.2f * s - 1e-3f) return false;
    float den = qa * qa - s2;
    if (den > 0.02f) {
        float inv = 1.0f / den;
        float ru = s * sqrtf(fmaxf(qa * qa + qb * qb - s2, 0.f));
        float rv = s * sqrtf(fmaxf(qa * qa + qc * qc - s2, 0.f));
        float u0 = (qa * qb - ru) * inv, u1 = (qa * qb + ru) * inv;
        float v0 = (qa * qc - rv) * inv, v1 = (qa * qc + rv) * inv;
        b[0] = u0 - 3e-6f * (1.0f + fabsf(u0));
        b[1] = u1 + 3e-6f * (1.0f + fabsf(u1));
        b[2] = v0 - 3e-6f * (1.0f + fabsf(v0));
        b[3] = v1 + 3e-6f * (1.0f + fabsf(v1));
        return !(b[1] < -1.f || b[0] > 1.f || b[3] < -1.f || b[2] > 1.f);
    }
    b[0] = b[2] = -2.f;  // cap reaches the horizon of this chart: take the whole face
    b[1] = b[3] = 2.f;
    return true;
}
__device__ __forceinline__ CellRect bounds_to_cells(const float b[4], int G)
{
    CellRect r;
    r.u0 = cell_coord(warp_uv(b[0]) - 2e-6f, G);
    r.u1 = cell_coord(warp_uv(b[1]) + 2e-6f, G);
    r.v0 = cell_coord(warp_uv(b[2]) - 2e-6f, G);
    r.v1 = cell_coord(warp_uv(b[3]) + 2e-6f, G);
    return r;
}
__device__ __forceinline__ bool bounds_inside_face(const float b[4])
{
    return b[0] > -1.f && b[1] < 1.f && b[2] > -1.f && b[3] < 1.f;
}

// sin^2(theta) <= (d/R)^2 for chord d, inflated so that float32 rounding can never exclude a candidate
template <typename T> __device__ __forceinline__ float cap_s2(T d2)
{
    const float invR2 = (float)(1.0 / (PRS_R * PRS_R));
    return (float)d2 * invR2 * 1.0002f + 1e-13f;
}

__device__ __forceinline__ int home_face(float fx, float fy, float fz)
{
    float ax = fabsf(fx), ay = fabsf(fy), az = fabsf(fz);
    if (ax >= ay && ax >= az) return fx < 0.f ? 1 : 0;
    if (ay >= az) return fy < 0.f ? 3 : 2;
    return fz < 0.f ? 5 : 4;
}

__device__ __forceinline__ uint32_t coarse_rect_count(const FaceTable &ct, const uint32_t *__restrict__ coarse_start,
                                                      int f, CellRect c)
{
    if (!clip_rect(ct, f, c)) return 0;
    uint32_t total = 0;
    // rows of fw + 1 row-local prefix sums (build_coarse_kernel)
    const int pitch = ct.fw[f] + 1;
    const uint32_t *row = coarse_start + ct.base[f] + (int64_t)(c.v0 - ct.fv0[f]) * pitch - ct.fu0[f];
    for (int iv = c.v0; iv <= c.v1; ++iv, row += pitch) total += __ldg(row + c.u1 + 1) - __ldg(row + c.u0);
    return total;
}

// Number of indexed points in the coarse cells touched by the cap: an upper bound of the points within the cap,
// so 0 proves there is none.  hb / single: bounds on the home face and whether the cap stays on that face.
__device__ __forceinline__ uint32_t coarse_count(const FaceTable &ct, const uint32_t *__restrict__ coarse_start, int Gc,
                                                 float fx, float fy, float fz, float s2, float s, int hface,
                                                 const float hb[4], bool single)
{
    uint32_t total = coarse_rect_count(ct, coarse_start, hface, bounds_to_cells(hb, Gc));
    if (!single) {
#pragma unroll 1
        for (int f = 0; f < 6; ++f) {
            float b[4];
            if (f == hface || !face_bounds(f, fx, fy, fz, s2, s, b)) continue;
            total += coarse_rect_count(ct, coarse_start, f, bounds_to_cells(b, Gc));
        }
    }
    return total;
}

// Scan the cells of `face` in rectangle c (already clipped to the table), skipping the visited block h when
// skip is set.  One row of cells = one contiguous run of records (two table reads); a row crossing the visited
// block splits into at most two runs.
template <typename T, int K>
__device__ __forceinline__ void scan_rect(const IndexView<T> &ix, int face, const CellRect c, bool skip, const CellRect h,
                                          T qx, T qy, T qz, T dub, TopK<T, K> &top)
{
    const FaceTable &ft = ix.fine;
    const uint32_t *row = ix.cell_start + ft.base[face] + (int64_t)(c.v0 - ft.fv0[face]) * ft.fw[face] - ft.fu0[face];
#pragma unroll 1
    for (int iv = c.v0; iv <= c.v1; ++iv, row += ft.fw[face]) {
        const bool split = skip && iv >= h.v0 && iv <= h.v1;
#pragma unroll 1
        for (int seg = 0; seg < 2; ++seg) {
            int a, b;
            if (!split) {
                if (seg) break;
                a = c.u0;
                b = c.u1;
            } else if (seg == 0) {
                a = c.u0;
                b = min(c.u1, h.u0 - 1);
            } else {
                a = max(c.u0, h.u1 + 1);
                b = c.u1;
            }
            if (a > b) continue;
            const uint32_t s = __ldg(row + a), e = __ldg(row + b + 1);
            // two records in flight per step (one load per iteration leaves the thread waiting a full memory round
            // trip per candidate)
#pragma unroll 1
            PRS_STAT(13, e - s);  // candidates offered by the general search
            for (uint32_t i = s; i < e; i += 2) {
                const Rec<T> r0 = load_rec(ix.recs + i), r1 = load_rec(ix.recs + min(i + 1, e - 1));
                top.offer(dist2<T>(qx, qy, qz, r0.x, r0.y, r0.z), r0.idx, dub);
                if (i + 1 < e) top.offer(dist2<T>(qx, qy, qz, r1.x, r1.y, r1.z), r1.idx, dub);
            }
        }
    }
}

// How phase 0 of the general search grows its block: 0 = rings of two cells walked as four strips, 1 = doubling
// (2m + 1), 2 = by half (m + m / 2 + 1), the last two scanned as one block with a hole
#ifndef PRS_GROWTH
#define PRS_GROWTH 1  // measured on C3 (search + epilogue, ms): strips 5.5, doubling 3.7, by half 3.9
#endif
// General search: continues after the home block `hb` around the home cell (cu, cv) has been scanned.
//   phase 0: grow the block ring by ring until it holds k candidates or covers the radius.  The rings are thin (two
//            cells): a point OUTSIDE the swath has its k nearest in the sliver where the block first cuts into the
//            swath, and a block that doubles each time would drag in thousands of candidates behind it (those points
//            were 4 % of C3's pixels and most of its search time).  Empty space is crossed 16 cells at a time after a
//            look at the coarse table;
//   phase 1: on the home face, visit every cell that can still beat (or tie) the k-th distance;
//   phase 2: the same on the other faces the cap reaches (only near cube edges).
// One loop with a single scan site.  This is the cold path (sparse data, swath edges, cube edges).
template <typename T, int K>
__device__ __noinline__ TopK<T, K> knn_search_general(const IndexView<T> &ix, T qx, T qy, T qz, T dub, float s2r, float sr,
                                                      TopK<T, K> top, CellRect hb, int hface, int cu, int cv)
{
    const float invR = (float)(1.0 / PRS_R);
    const float fx = (float)qx * invR, fy = (float)qy * invR, fz = (float)qz * invR;
    const int G = ix.G;
    float hbnd[4];
    face_bounds(hface, fx, fy, fz, s2r, sr, hbnd);  // the home face always contains part of the cap
    const bool single = bounds_inside_face(hbnd);
    // too few candidates around the point itself: is there any source point within the radius at all?
    if (top.id[0] == PRS_NONE &&
        coarse_count(ix.coarse, ix.coarse_start, ix.Gc, fx, fy, fz, s2r, sr, hface, hbnd, single) == 0)
        return top;
    CellRect rr = bounds_to_cells(hbnd, G);  // everything within the radius on the home face
    int phase = top.full() ? 1 : (clip_rect(ix.fine, hface, rr) ? 0 : 1);
    PRS_STAT(1, 1);                    // points entering the general search past the coarse test
    PRS_STAT(2, top.full() ? 1 : 0);   // ... with a full list (only the cap check failed)
    PRS_STAT(3, top.id[0] == PRS_NONE ? 1 : 0);  // ... with an empty list
    // half-width of what has been visited around (cu, cv)
    int m = hb.u0 <= hb.u1 ? max(max(cu - hb.u0, hb.u1 - cu), max(cv - hb.v0, hb.v1 - cv)) : 0;
    int f = 0;
    // phase 0 walks a ring as four strips around the visited block (sub = 0..3: above, below, left, right), each only
    // when the coarse table shows points under it; sub == 4: grow the block; sub == 5: the whole new block at once
    int sub = 4;
    CellRect cnew;
    cnew.u0 = cnew.v0 = 1, cnew.u1 = cnew.v1 = 0;
    float s2b = s2r, sb = sr;
    if (top.full()) {
        s2b = cap_s2<T>(top.d[K - 1]);
        sb = sqrtf(s2b);
    }
#pragma unroll 1
    for (;;) {
        CellRect c;
        c.u0 = c.v0 = 1;
        c.u1 = c.v1 = 0;
        int face = hface;
        bool skip = false, scan = false;
        if (phase == 0) {
            if (sub == 4) {
                // look ahead: when the coarse cells under the block 16 cells further out hold no point at all, that
                // whole block is crossed without a scan
                CellRect far;
                far.u0 = max(cu - m - COARSE, rr.u0), far.u1 = min(cu + m + COARSE, rr.u1);
                far.v0 = max(cv - m - COARSE, rr.v0), far.v1 = min(cv + m + COARSE, rr.v1);
                bool empty_ahead = false;
                if (far.u0 <= far.u1 && far.v0 <= far.v1) {
                    CellRect cc;
                    cc.u0 = far.u0 / COARSE, cc.u1 = far.u1 / COARSE, cc.v0 = far.v0 / COARSE, cc.v1 = far.v1 / COARSE;
                    empty_ahead = coarse_rect_count(ix.coarse, ix.coarse_start, hface, cc) == 0;
                }
#if PRS_GROWTH == 0
                m += empty_ahead ? COARSE : 2;
#elif PRS_GROWTH == 1
                m = empty_ahead ? m + COARSE : 2 * m + 1;
#else
                m = empty_ahead ? m + COARSE : m + (m >> 1) + 1;
#endif
                cnew.u0 = max(cu - m, rr.u0);
                cnew.u1 = min(cu + m, rr.u1);
                cnew.v0 = max(cv - m, rr.v0);
                cnew.v1 = min(cv + m, rr.v1);
                const bool nonempty = cnew.u0 <= cnew.u1 && cnew.v0 <= cnew.v1;
                const bool around = PRS_GROWTH == 0 && nonempty && hb.u0 <= hb.u1 && cnew.u0 <= hb.u0 &&
                                    cnew.u1 >= hb.u1 && cnew.v0 <= hb.v0 && cnew.v1 >= hb.v1;
                if (empty_ahead || !nonempty) {
                    sub = 5;  // nothing to scan
                } else if (around) {
                    sub = 0;  // four strips around the visited block
                } else {
                    sub = 5;  // no visited block to go around (or it sticks out of the radius): the whole block
                    c = cnew;
                    skip = hb.u0 <= hb.u1;
                    scan = true;
                }
            }
            if (sub < 4) {
                c = cnew;
                if (sub == 0)
                    c.v1 = hb.v0 - 1;
                else if (sub == 1)
                    c.v0 = hb.v1 + 1;
                else {
                    c.v0 = hb.v0, c.v1 = hb.v1;
                    if (sub == 2)
                        c.u1 = hb.u0 - 1;
                    else
                        c.u0 = hb.u1 + 1;
                }
                if (c.u0 <= c.u1 && c.v0 <= c.v1) {
                    CellRect cc;
                    cc.u0 = c.u0 / COARSE, cc.u1 = c.u1 / COARSE, cc.v0 = c.v0 / COARSE, cc.v1 = c.v1 / COARSE;
                    scan = coarse_rect_count(ix.coarse, ix.coarse_start, hface, cc) != 0;
                }
            }
        } else if (phase == 1) {
            float b[4];
            face_bounds(hface, fx, fy, fz, s2b, sb, b);
            c = bounds_to_cells(b, G);
            skip = hb.u0 <= hb.u1;
            scan = clip_rect(ix.fine, hface, c) &&
                   !(skip && c.u0 >= hb.u0 && c.u1 <= hb.u1 && c.v0 >= hb.v0 && c.v1 <= hb.v1);
        } else {
            float b[4];
            face = f;
            if (f != hface && face_bounds(f, fx, fy, fz, s2b, sb, b)) {
                c = bounds_to_cells(b, G);
                scan = clip_rect(ix.fine, f, c);
            }
        }
        PRS_STAT(4 + phase, 1);  // loop iterations per phase
        if (scan) {
            PRS_STAT(7 + phase, 1);  // scans per phase
            PRS_STAT(10 + phase, (long long)(c.u1 - c.u0 + 1) * (c.v1 - c.v0 + 1));  // cells per phase
        }
        if (scan) scan_rect<T, K>(ix, face, c, skip, hb, qx, qy, qz, dub, top);
        if (phase == 0) {
            if (sub < 3) {
                ++sub;
                continue;
            }
            // the ring is done.  The visited region is tracked as one rectangle; a block that does not contain the
            // previous one is simply not recorded (its cells may be offered again later - TopK::offer drops duplicates)
            if (cnew.u0 <= cnew.u1 && cnew.v0 <= cnew.v1 &&
                (hb.u0 > hb.u1 || (cnew.u0 <= hb.u0 && cnew.u1 >= hb.u1 && cnew.v0 <= hb.v0 && cnew.v1 >= hb.v1)))
                hb = cnew;
            sub = 4;
            if (top.full() || (cu - m <= rr.u0 && cu + m >= rr.u1 && cv - m <= rr.v0 && cv + m >= rr.v1)) {
                phase = 1;
                if (top.full()) {
                    s2b = cap_s2<T>(top.d[K - 1]);
                    sb = sqrtf(s2b);
                }
            }
        } else if (phase == 1) {
            if (single) break;
            phase = 2;
        } else {
            if (++f >= 6) break;
        }
    }
    return top;
}

// ------------------------------------------------------------------------------------------ bulk-copy staging
// Blackwell / Hopper asynchronous-copy primitives used by the search kernel: 1-D bulk copies global -> shared
// (cp.async.bulk, SASS UBLKCP) whose completion is counted in bytes on a shared-memory mbarrier.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// dst (shared) and src (global) 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// orders this thread's earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Candidate records of one tile, staged in shared memory.  The threads of a tile (32 x QROWS neighbouring target
// points) scan almost the same cells, and one row of cells is ONE contiguous run of records: the block takes, for
// every row of cells any of its points touches, the span of cells between the leftmost and the rightmost touched
// one (the "hull" of the tile's home blocks - tight also when the tile lies diagonally on the cube face) and fetches
// each row's run with one bulk copy.  The matching stretch of the cell table is staged next to it, so that a thread
// finds the bounds of its runs without going to global memory.
constexpr int HULL_ROWS = 64;  // rows of cells a staged tile may span (rows are slotted modulo this)
#ifndef PRS_HOME_MAX
#define PRS_HOME_MAX 4
#endif
constexpr int HOME_MAX = PRS_HOME_MAX;  // largest half-width of a home block: (2 * HOME_MAX + 1)^2 cells
// Shared memory per block: records + cell-table stretch.  A tile's hull holds roughly (points per cell) x (cells the
// tile covers + a margin of m cells), and cells are sized for 2 .. k / 2 points (index_build_t: 2 for k <= 4, 0.75 k up to 8, 0.5 k above).
#ifndef PRS_STAGE_BYTES
#define PRS_STAGE_BYTES(K) ((K) <= 4 ? 24576 : 57344)
#endif
constexpr int STAGE_CELLS = 2048;  // entries of the staged cell-table stretch (8 KB)
struct StageCtx {
    unsigned long long bar;          // mbarrier: 32 arrivals (warp 0) + the bytes of the row copies
    int umin[HULL_ROWS], umax[HULL_ROWS];
    uint32_t gstart[HULL_ROWS];      // index of the first staged record of the row in recs[]
    uint32_t sbase[HULL_ROWS];       // ... and its slot in the record buffer
    int csoff[HULL_ROWS];            // staged cell-table stretch: cell_start[row][u] sits at scs[csoff + u]
    int vlo, vhi;
    unsigned faces;
    int staged;                      // 1: the tile's candidates are (being) staged
};

// Home cell of a query point and the (clipped) block of (2m + 1)^2 cells around it.
struct Home {
    int face, cu, cv;
    CellRect hb;  // empty (u0 > u1) when the block lies outside the table
};
template <typename T> __device__ __forceinline__ Home home_cell(const IndexView<T> &ix, T qx, T qy, T qz)
{
    const float invR = (float)(1.0 / PRS_R);
    const float fx = (float)qx * invR, fy = (float)qy * invR, fz = (float)qz * invR;
    Home h;
    float hu, hv;
    face_uv(fx, fy, fz, h.face, hu, hv);
    h.cu = cell_coord(warp_uv(hu), ix.G);
    h.cv = cell_coord(warp_uv(hv), ix.G);
    h.hb.u0 = h.hb.v0 = 1;
    h.hb.u1 = h.hb.v1 = 0;
    return h;
}
template <typename T> __device__ __forceinline__ void home_block(const IndexView<T> &ix, Home &h, int m)
{
    const FaceTable &ft = ix.fine;
    h.hb.u0 = max(h.cu - m, ft.fu0[h.face]);
    h.hb.u1 = min(h.cu + m, ft.fu0[h.face] + ft.fw[h.face] - 1);
    h.hb.v0 = max(h.cv - m, ft.fv0[h.face]);
    h.hb.v1 = min(h.cv + m, ft.fv0[h.face] + ft.fh[h.face] - 1);
    if (ft.fw[h.face] == 0 || h.hb.u0 > h.hb.u1 || h.hb.v0 > h.hb.v1) {
        h.hb.u0 = h.hb.v0 = 1;
        h.hb.u1 = h.hb.v1 = 0;
    }
}

// Number of indexed points in the coarse cell (COARSE x COARSE fine cells) a home cell belongs to: the local density.
template <typename T> __device__ __forceinline__ uint32_t home_density(const IndexView<T> &ix, const Home &h)
{
    const FaceTable &ct = ix.coarse;
    const int cu = h.cu / COARSE, cv = h.cv / COARSE, f = h.face;
    if (ct.fw[f] == 0 || cu < ct.fu0[f] || cu >= ct.fu0[f] + ct.fw[f] || cv < ct.fv0[f] || cv >= ct.fv0[f] + ct.fh[f]) return 0;
    const uint32_t *p = ix.coarse_start + ct.base[f] + (int64_t)(cv - ct.fv0[f]) * (ct.fw[f] + 1) + (cu - ct.fu0[f]);
    return __ldg(p + 1) - __ldg(p);
}

// Half-width of the home block for k neighbours where a cell holds `lambda` points on average: the k-th neighbour
// lies ~sqrt(k / (pi lambda)) cells away, and the block must contain the cap of that radius around a point that may
// sit at the edge of its own cell.
#ifndef PRS_HOME_FACTOR
#define PRS_HOME_FACTOR 1.1f
#endif
__device__ __forceinline__ int home_halfwidth(int k, float lambda)
{
    if (!(lambda > 0.f)) return HOME_MAX;
    const float r = sqrtf((float)k / (3.14159265f * lambda));
    return min(HOME_MAX, max(1, (int)ceilf(PRS_HOME_FACTOR * r)));
}

// Bounds [s, e) in recs[] of the three rows of a 3x3 home block, the point's own row first: its candidates are the
// likeliest neighbours, so the k-th distance tightens early and fewer later candidates enter the list.
struct HomeRuns {
    uint32_t s[3], e[3];  // empty run (s == e) for a row outside the table
};
template <typename T> __device__ __forceinline__ HomeRuns home_runs(const IndexView<T> &ix, const Home &h)
{
    HomeRuns r;
    const int v[3] = {h.cv, h.cv - 1, h.cv + 1};
    const FaceTable &ft = ix.fine;
    const int fw = ft.fw[h.face];
    const uint32_t *row0 = ix.cell_start + ft.base[h.face] - (int64_t)ft.fv0[h.face] * fw - ft.fu0[h.face];
    // all run bounds first: six independent loads in flight together
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        r.s[j] = r.e[j] = 0;
        if (h.hb.u0 <= h.hb.u1 && v[j] >= h.hb.v0 && v[j] <= h.hb.v1) {
            const uint32_t *row = row0 + (int64_t)v[j] * fw;
            r.s[j] = __ldg(row + h.hb.u0);
            r.e[j] = __ldg(row + h.hb.u1 + 1);
        }
    }
    return r;
}

// Hot path, records from global memory: the runs are walked as ONE flattened loop so that the lanes of a warp stay
// busy; records are fetched PRS_SCAN_BATCH at a time before any is examined (one load per iteration left the
// thread waiting a full memory round trip per candidate).
template <typename T, int K>
__device__ __forceinline__ void scan_home_global(const IndexView<T> &ix, const HomeRuns &hr, T qx, T qy, T qz, T dub,
                                                 TopK<T, K> &top)
{
    const uint32_t n0 = hr.e[0] - hr.s[0], n01 = n0 + (hr.e[1] - hr.s[1]), total = n01 + (hr.e[2] - hr.s[2]);
    constexpr int B = PRS_SCAN_BATCH(K);
#pragma unroll 1
    for (uint32_t t0 = 0; t0 < total; t0 += B) {
        Rec<T> r[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const uint32_t t = min(t0 + u, total - 1);
            const uint32_t i = t < n0 ? hr.s[0] + t : (t < n01 ? hr.s[1] + (t - n0) : hr.s[2] + (t - n01));
            r[u] = load_rec(ix.recs + i);
        }
#pragma unroll
        for (int u = 0; u < B; ++u)
            if (t0 + u < total)
                top.template offer<false>(dist2<T>(qx, qy, qz, r[u].x, r[u].y, r[u].z), r[u].idx, dub);
    }
}

__device__ __forceinline__ Rec<float> lds_rec(const Rec<float> *p)
{
    const float4 v = *reinterpret_cast<const float4 *>(p);
    Rec<float> r;
    r.x = v.x, r.y = v.y, r.z = v.z, r.idx = __float_as_uint(v.w);
    return r;
}
__device__ __forceinline__ Rec<double> lds_rec(const Rec<double> *p)
{
    const double2 a = reinterpret_cast<const double2 *>(p)[0], b = reinterpret_cast<const double2 *>(p)[1];
    Rec<double> r;
    r.x = a.x, r.y = a.y, r.z = b.x, r.idx = (uint32_t)(__double_as_longlong(b.y) & 0xFFFFFFFFLL), r.pad = 0;
    return r;
}

// Hot path, records and cell bounds staged in shared memory: the rows of the home block, the point's own row first,
// then outwards.  Lists of 16 or more are filled chunk-wise through the sorting network (TopK::merge_chunk).
#ifndef PRS_CHUNK_MIN_K
#define PRS_CHUNK_MIN_K 16  // for k = 8 the network loses to offer() (C3: 5.6 vs 5.2 ms)
#endif
template <typename T, int K>
__device__ __forceinline__ void scan_home_staged(const StageCtx &sc, const Rec<T> *srecs, const uint32_t *scs,
                                                 const Home &h, T qx, T qy, T qz, T dub, TopK<T, K> &top)
{
    if (h.hb.u0 > h.hb.u1) return;  // home block outside the table
    const int m = max(h.cv - h.hb.v0, h.hb.v1 - h.cv);
    if constexpr (K >= PRS_CHUNK_MIN_K) {
        constexpr int CH = K < 16 ? K : 16;
        // a cursor over the candidates of all rows: row `step` (0: cv; 1: cv - 1; 2: cv + 1; ...), records [i, e)
        int step = -1;
        uint32_t i = 0, e = 0;
        const Rec<T> *run = srecs;
        bool more = true;
#pragma unroll 1
        while (more) {
            T cd[CH];
            uint32_t ci[CH];
            bool any = false;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                cd[j] = Ar<T>::inf();
                ci[j] = PRS_NONE;
                while (more && i >= e) {  // next row with records
                    ++step;
                    if (step > 2 * m) {
                        more = false;
                        break;
                    }
                    const int off = (step + 1) >> 1;
                    const int v = (step & 1) ? h.cv - off : h.cv + off;
                    if (v < h.hb.v0 || v > h.hb.v1) continue;
                    const int slot = v & (HULL_ROWS - 1);
                    const uint32_t *cs = scs + sc.csoff[slot];
                    i = cs[h.hb.u0];
                    e = cs[h.hb.u1 + 1];
                    run = srecs + sc.sbase[slot] - sc.gstart[slot];
                }
                if (more) {
                    const Rec<T> r = lds_rec(run + i);
                    ++i;
                    const T d2 = dist2<T>(qx, qy, qz, r.x, r.y, r.z);
                    // strict radius test in the tree dtype (pykdtree); what cannot enter the list is left out of the sort
                    if (d2 < dub && TopK<T, K>::before(d2, r.idx, top.d[K - 1], top.id[K - 1])) {
                        cd[j] = d2;
                        ci[j] = r.idx;
                        any = true;
                    }
                }
            }
            if (__any_sync(__activemask(), any)) top.template merge_chunk<CH>(cd, ci);
        }
    } else {
#pragma unroll 1
    for (int step = 0; step <= 2 * m; ++step) {
        // step 0: cv; 1: cv - 1; 2: cv + 1; 3: cv - 2; ...
        const int off = (step + 1) >> 1;
        const int v = (step & 1) ? h.cv - off : h.cv + off;
        if (v < h.hb.v0 || v > h.hb.v1) continue;
        const int slot = v & (HULL_ROWS - 1);
        const uint32_t *cs = scs + sc.csoff[slot];
        const uint32_t s = cs[h.hb.u0], e = cs[h.hb.u1 + 1];
        const Rec<T> *run = srecs + sc.sbase[slot] - sc.gstart[slot];
#pragma unroll 1
        for (uint32_t i = s; i < e; ++i) {
            const Rec<T> r = lds_rec(run + i);
            top.template offer<false>(dist2<T>(qx, qy, qz, r.x, r.y, r.z), r.idx, dub);
        }
    }
    }
}

// After the home block: done when the cap of the k-th distance provably stays inside the block (closed-form bounds,
// home face only); anything else continues in knn_search_general.
template <typename T, int K>
__device__ __forceinline__ void knn_finish(const IndexView<T> &ix, const Home &h, T qx, T qy, T qz, T dub, float s2r,
                                           float sr, TopK<T, K> &top)
{
    const float invR = (float)(1.0 / PRS_R);
    const float fx = (float)qx * invR, fy = (float)qy * invR, fz = (float)qz * invR;
    if (top.full()) {
        const float s2b = cap_s2<T>(top.d[K - 1]);
        float b[4];
        face_bounds(h.face, fx, fy, fz, s2b, sqrtf(s2b), b);
        if (bounds_inside_face(b)) {
            CellRect c = bounds_to_cells(b, ix.G);
            // cells of the bound cap outside the table hold no points; inside the visited block nothing is left
            if (!clip_rect(ix.fine, h.face, c) ||
                (c.u0 >= h.hb.u0 && c.u1 <= h.hb.u1 && c.v0 >= h.hb.v0 && c.v1 <= h.hb.v1))
                return;
        }
    }
    PRS_STAT(24, 1);                  // points sent to the general search
    PRS_STAT(25, top.full() ? 1 : 0);  // ... although their list was full
    top = knn_search_general<T, K>(ix, qx, qy, qz, dub, s2r, sr, top, h.hb, h.face, h.cu, h.cv);
}

// ------------------------------------------------------------------------------------------ targets
#define PRS_TK_LONLAT 0    // explicit lon/lat arrays
#define PRS_TK_AREA_SEP 1  // area, separable projection: per-row / per-column tables
#define PRS_TK_AREA_GEN 2  // area, per-pixel inverse projection

template <typename T> struct ColTab {  // per column: cos/sin of the longitude
    T c, s;
};
template <typename T> struct RowTab {  // per row: R*cos(lat), R*sin(lat)
    T rc, rz;
};

#define PRS_EPI_INFO 0
#define PRS_EPI_NEAREST 1
#define PRS_EPI_GAUSS 2     // QueryParams::mode; the search kernels are instantiated per accumulation / data dtype:
#define PRS_EPI_GAUSS_FF 2  //   float32 accumulation, float32 data (single channel, float32 tree)
#define PRS_EPI_GAUSS_DF 3  //   float64 accumulation, float32 data
#define PRS_EPI_GAUSS_DD 4  //   float64 accumulation, float64 data

struct QueryParams {
    // target
    int64_t n_t;
    const void *tlon, *tlat;
    const uint8_t *valid_extra;
    prs_proj_params proj;
    prs_area_grid grid;
    int64_t row0;
    int64_t row_block, row_stride;  // block-cyclic row ownership (row_block == 0: contiguous)
    int nrows;      // area targets: local rows of this call
    int tiles_x;    // tiles per tile-row
    int tk;         // PRS_TK_* (a template parameter of the classification, a run-time value for the search)
    const void *col_tab, *row_tab;
    const uint8_t *col_ok, *row_ok;
    // search
    double radius;
    int k;
    // epilogue
    int mode;
    uint32_t *idx_out;
    void *dist_out;
    uint8_t *valid_out;
    unsigned int *flags;  // [0] any k-th neighbour found, [1] number of invalid target points
    // nearest gather
    const void *data;
    const void *fill_row;
    void *out;
    int row_bytes;
    int vec;  // copy granularity in bytes (1, 2, 4, 8, 16)
    // gauss
    int channels;
    int data_f64;
    int acc_f64;
    const double *sigma2;  // device [channels]: float(sigma)**2
    int uniform_sigma;
    double fill;
    void *stddev_out, *count_out;
    // Targets whose coordinates are expensive (per-pixel inverse projection, explicit lon/lat): classification
    // computes every point anyway and leaves (x, y, z, valid) here for the search (4 values of the tree dtype).
    // Always set for PRS_TK_LONLAT / PRS_TK_AREA_GEN; separable areas use the tables instead.
    void *target_xyz;
    // staged execution (prs_resample_nearest_fill / _search): 0 = classify + fill + search in one call,
    // 1 = classify + fill only (live tiles listed in scratch), 2 = search the tiles listed in scratch
    int stage;
    uint32_t *scratch;       // caller-owned: [0] = number of live tiles, [1..] = their ids, then the coordinates
    uint8_t *tile_row_live;  // stage 1: [tile row] = 1 when the tile row holds a live tile
};

__device__ __forceinline__ void copy_row(void *dst, const void *src, int bytes, int vec)
{
    switch (vec) {
    case 16:
        for (int o = 0; o < bytes; o += 16)
            *reinterpret_cast<uint4 *>((char *)dst + o) = *reinterpret_cast<const uint4 *>((const char *)src + o);
        break;
    case 8:
        for (int o = 0; o < bytes; o += 8)
            *reinterpret_cast<uint2 *>((char *)dst + o) = *reinterpret_cast<const uint2 *>((const char *)src + o);
        break;
    case 4:
        for (int o = 0; o < bytes; o += 4)
            *reinterpret_cast<uint32_t *>((char *)dst + o) = *reinterpret_cast<const uint32_t *>((const char *)src + o);
        break;
    case 2:
        for (int o = 0; o < bytes; o += 2)
            *reinterpret_cast<uint16_t *>((char *)dst + o) = *reinterpret_cast<const uint16_t *>((const char *)src + o);
        break;
    default:
        for (int o = 0; o < bytes; ++o) ((uint8_t *)dst)[o] = ((const uint8_t *)src)[o];
    }
}

// Load up to 4 consecutive channels of one source row (vectorised when the row start is 16-byte aligned).
template <typename Td>
__device__ __forceinline__ void load_channels(const Td *__restrict__ p, int nc, bool vec_ok, Td v[4])
{
    if (vec_ok && nc == 4) {
        if (sizeof(Td) == 4) {
            float4 f = __ldg(reinterpret_cast<const float4 *>(p));
            v[0] = (Td)f.x;
            v[1] = (Td)f.y;
            v[2] = (Td)f.z;
            v[3] = (Td)f.w;
        } else {
            double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
            v[0] = (Td)a.x;
            v[1] = (Td)a.y;
            v[2] = (Td)b.x;
            v[3] = (Td)b.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = j < nc ? __ldg(p + j) : (Td)0;
    }
}

// _resample_with_weights (kd_tree.py:741-818) + _calculate_uncertainty (kd_tree.py:821-859) for one target point that
// has at least one neighbour, fused behind the search.  T: distance dtype (the Gaussian is evaluated in it,
// kd_tree.py:165), Td: data dtype, Ta: accumulation / output dtype (numpy promotion: float64 for multi-channel data,
// kd_tree.py:783).  Channels are processed four at a time so that a neighbour's row is read with 16-byte loads; per
// channel the neighbours are accumulated in order i = 0..k-1 exactly like the reference's Python loop.
template <typename T, typename Ta, typename Td, int K>
__device__ __forceinline__ void gauss_epilogue(const QueryParams &P, const IndexView<T> &ix, int64_t p,
                                               const TopK<T, K> &top)
{
    const int C = P.channels, k = P.k;
    const bool want_uncert = P.stddev_out != nullptr;
    Ta *out = (Ta *)P.out + (size_t)p * C;
    Ta *sd_out = want_uncert ? (Ta *)P.stddev_out + (size_t)p * C : nullptr;
    Ta *cnt_out = want_uncert ? (Ta *)P.count_out + (size_t)p * C : nullptr;
    const Td *data = (const Td *)P.data;
    const bool vec_ok = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(data) & 15) == 0);
    T dist2w[K];  // r ** 2 with r the rounded distance; missing neighbours: distance 1 (kd_tree.py:767-768)
    int count = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const bool missing = top.id[i] == PRS_NONE;
        const T r = missing ? (T)1 : Ar<T>::sqrt(top.d[i]);
        dist2w[i] = Ar<T>::mul(r, r);
        if (i < k && !missing) ++count;
    }
    Ta wu[K];  // weights when every channel has the same sigma
    if (P.uniform_sigma) {
        const T s2 = (T)P.sigma2[0];  // float(sigma)**2 as a weak Python scalar -> distance dtype
#pragma unroll
        for (int i = 0; i < K; ++i) wu[i] = (Ta)Ar<T>::exp(Ar<T>::div(-dist2w[i], s2));
    }
#pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 4) {
        const int nc = min(4, C - c0);
        Ta res[4] = {0, 0, 0, 0}, norm[4] = {0, 0, 0, 0};
        T s2c[4];
        if (!P.uniform_sigma) {
#pragma unroll
            for (int j = 0; j < 4; ++j) s2c[j] = j < nc ? (T)P.sigma2[c0 + j] : (T)1;
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
            if (i < k) {
                const bool missing = top.id[i] == PRS_NONE;
                // missing neighbours gather new_data[0] and get weight 0 * w (kd_tree.py:754-759, 802)
                const uint32_t row = missing ? ix.first_valid_source : top.id[i];
                Td v[4];
                load_channels<Td>(data + (size_t)row * C + c0, nc, vec_ok, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    Ta w = P.uniform_sigma ? wu[i] : (Ta)Ar<T>::exp(Ar<T>::div(-dist2w[i], s2c[j]));
                    Ta wt = missing ? Ar<Ta>::mul((Ta)0, w) : w;
                    res[j] = Ar<Ta>::add(res[j], Ar<Ta>::mul(wt, (Ta)v[j]));
                    norm[j] = Ar<Ta>::add(norm[j], wt);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // norm > 0 is False for NaN as well, like numpy's boolean mask
            res[j] = (norm[j] > 0) ? Ar<Ta>::div(res[j], norm[j]) : (Ta)P.fill;
            if (j < nc) out[c0 + j] = res[j];
        }
        if (want_uncert) {
            Ta nsq[4] = {0, 0, 0, 0}, sd[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < K; ++i) {
                if (i < k) {
                    const bool missing = top.id[i] == PRS_NONE;
                    const uint32_t row = missing ? ix.first_valid_source : top.id[i];
                    Td v[4];
                    load_channels<Td>(data + (size_t)row * C + c0, nc, vec_ok, v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        Ta w = P.uniform_sigma ? wu[i] : (Ta)Ar<T>::exp(Ar<T>::div(-dist2w[i], s2c[j]));
                        Ta wt = missing ? Ar<Ta>::mul((Ta)0, w) : w;
                        nsq[j] = Ar<Ta>::add(nsq[j], Ar<Ta>::mul(wt, wt));
                        Ta val = missing ? Ar<Ta>::mul((Ta)0, (Ta)v[j]) : (Ta)v[j];
                        Ta dv = Ar<Ta>::sub(val, res[j]);
                        sd[j] = Ar<Ta>::add(sd[j], Ar<Ta>::mul(wt, Ar<Ta>::mul(dv, dv)));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nc) {
                    sd_out[c0 + j] =
                        count > 1 ? Ar<Ta>::sqrt(Ar<Ta>::mul(
                                        Ar<Ta>::div(norm[j], Ar<Ta>::sub(Ar<Ta>::mul(norm[j], norm[j]), nsq[j])), sd[j]))
                                  : Ar<Ta>::nan();
                    cnt_out[c0 + j] = (Ta)count;
                }
            }
        }
    }
}

// A weighted pixel without any neighbour (or outside valid_output_index, kd_tree.py:719-723, 865-868): every weight is
// zero, so norm == 0 and the pixel is the fill value whatever new_data[0] holds (kd_tree.py:807-809).
template <typename Ta> __device__ __forceinline__ void gauss_empty(const QueryParams &P, int64_t p)
{
    const int C = P.channels;
    Ta *out = (Ta *)P.out + (size_t)p * C;
    for (int c = 0; c < C; ++c) out[c] = (Ta)P.fill;
    if (P.stddev_out) {
        Ta *sd = (Ta *)P.stddev_out + (size_t)p * C, *cnt = (Ta *)P.count_out + (size_t)p * C;
        for (int c = 0; c < C; ++c) {
            sd[c] = Ar<Ta>::nan();
            cnt[c] = 0;
        }
    }
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------ tiles
// A tile is 32 columns x QROWS rows of target points (arrays: 32*QROWS consecutive points, lane-strided).
// tile_origin: first row / this lane's column of the tile (area targets); tile_point: raveled index or -1.
__device__ __forceinline__ void tile_origin(const QueryParams &P, int TK, uint32_t tile, int lane, int &trow0, int &tcol)
{
    if (TK == PRS_TK_LONLAT) {
        trow0 = tcol = 0;
        return;
    }
    const uint32_t ty = tile / (uint32_t)P.tiles_x, tx = tile - ty * (uint32_t)P.tiles_x;
    trow0 = (int)ty * QROWS;
    tcol = (int)tx * 32 + lane;
}
__device__ __forceinline__ int64_t tile_point(const QueryParams &P, int TK, uint32_t tile, int trow0, int tcol, int j,
                                              int lane)
{
    if (TK == PRS_TK_LONLAT) {
        int64_t q = (int64_t)tile * (32 * QROWS) + j * 32 + lane;
        return q < P.n_t ? q : -1;
    }
    const int W = (int)P.grid.width;
    return (trow0 + j < P.nrows && tcol < W) ? (int64_t)(trow0 + j) * W + tcol : -1;
}

// Global row of local row l (see prs_target in include/prs_b200.h).
__device__ __forceinline__ int64_t global_row(int64_t row0, int64_t row_block, int64_t row_stride, int64_t l)
{
    if (row_block <= 0) return row0 + l;
    const int64_t s = l / row_block;
    return row0 + s * row_stride + (l - s * row_block);
}

// lon/lat of one pixel of a general (non-separable) area: ONE out-of-line copy of the projection code per kernel
// (inlined at every use the double-precision inverse projections made the kernels hundreds of KB of SASS).
template <typename T>
__device__ __noinline__ void area_point_lonlat(const QueryParams &P, int trow, int tcol, T &lon, T &lat)
{
    area_pixel_lonlat<T>(P.proj, P.grid, global_row(P.row0, P.row_block, P.row_stride, trow), tcol, lon, lat);
}

// Target coordinates of one point (kd_tree.py:513-534): validity + ECEF in the tree dtype.
template <typename T, int TK>
__device__ __forceinline__ bool target_point(const QueryParams &P, int64_t p, int trow, int tcol, T &qx, T &qy, T &qz)
{
    bool valid;
    qx = qy = qz = 0;
    if (TK == PRS_TK_AREA_SEP) {
        ColTab<T> ct = ((const ColTab<T> *)P.col_tab)[tcol];
        RowTab<T> rt = ((const RowTab<T> *)P.row_tab)[trow];
        valid = P.col_ok[tcol] && P.row_ok[trow];
        qx = Ar<T>::mul(rt.rc, ct.c);
        qy = Ar<T>::mul(rt.rc, ct.s);
        qz = rt.rz;
    } else {
        T lon, lat;
        if (TK == PRS_TK_LONLAT) {
            lon = ((const T *)P.tlon)[p];
            lat = ((const T *)P.tlat)[p];
        } else {
            area_point_lonlat<T>(P, trow, tcol, lon, lat);
        }
        valid = lonlat_in_range<T>(lon, lat);
        if (valid) lonlat_to_xyz(lon, lat, qx, qy, qz);
    }
    if (P.valid_extra) valid = valid && P.valid_extra[p];
    return valid;
}

// Target coordinates handed from classification to the search: (x, y, z, valid != 0) per point.
template <typename T> __device__ __forceinline__ void store_target_xyz(void *buf, int64_t p, T x, T y, T z, bool ok)
{
    if (sizeof(T) == 4) {
        reinterpret_cast<float4 *>(buf)[p] = make_float4((float)x, (float)y, (float)z, ok ? 1.f : 0.f);
    } else {
        double2 *q = reinterpret_cast<double2 *>(buf) + 2 * p;
        q[0] = make_double2((double)x, (double)y);
        q[1] = make_double2((double)z, ok ? 1.0 : 0.0);
    }
}
template <typename T> __device__ __forceinline__ bool load_target_xyz(const void *buf, int64_t p, T &x, T &y, T &z)
{
    if (sizeof(T) == 4) {
        const float4 v = reinterpret_cast<const float4 *>(buf)[p];
        x = (T)v.x, y = (T)v.y, z = (T)v.z;
        return v.w != 0.f;
    }
    const double2 *q = reinterpret_cast<const double2 *>(buf) + 2 * p;
    const double2 a = q[0], b = q[1];
    x = (T)a.x, y = (T)a.y, z = (T)b.x;
    return b.y != 0.0;
}

// The search's view of a target point: separable areas recompute it from the tables (two multiplies), everything
// else reads what classification left in target_xyz.
template <typename T>
__device__ __forceinline__ bool query_point(const QueryParams &P, int64_t p, int trow, int tcol, T &qx, T &qy, T &qz)
{
    if (P.tk == PRS_TK_AREA_SEP) return target_point<T, PRS_TK_AREA_SEP>(P, p, trow, tcol, qx, qy, qz);
    return load_target_xyz<T>(P.target_xyz, p, qx, qy, qz);
}

// ------------------------------------------------------------------------------------------ results
// Neighbour info of one point (kd_tree.py:545-548): ranks, distances, valid_output_index.  All 32 lanes call this.
template <typename T, int K>
__device__ __forceinline__ void emit_info(const QueryParams &P, const IndexView<T> &ix, int64_t p, int lane,
                                          const TopK<T, K> &top, bool valid)
{
    bool kth = false;
    if (p >= 0) {
        const int k = P.k;
        uint32_t *io = P.idx_out + (size_t)p * k;
        T *dout = P.dist_out ? (T *)P.dist_out + (size_t)p * k : nullptr;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            if (i < k) {
                uint32_t s = top.id[i];
                io[i] = s == PRS_NONE ? ix.n_valid : (ix.rank ? __ldg(ix.rank + s) : s);
                if (dout) dout[i] = s == PRS_NONE ? Ar<T>::inf() : Ar<T>::sqrt(top.d[i]);
                if (i == k - 1 && s != PRS_NONE) kth = true;
            }
        }
        if (P.valid_out) P.valid_out[p] = valid ? 1 : 0;
    }
    if (P.flags) {
        if (kth && P.flags[0] == 0) P.flags[0] = 1;
        unsigned bad = __ballot_sync(0xffffffffu, p >= 0 && !valid);
        if (lane == 0 && bad) atomicAdd(&P.flags[1], (unsigned)__popc(bad));
    }
}

// Nearest-neighbour row of one point (kd_tree.py:705-711).  All 32 lanes call this: rows of up to STAGE_ROW_BYTES
// are staged in shared memory and streamed out as contiguous 128-byte segments.  src_index == PRS_NONE: fill row.
__device__ __forceinline__ void emit_nearest(const QueryParams &P, int64_t p, int trow, int tcol, int lane,
                                             unsigned char *stage, uint32_t src_index)
{
    const void *src = src_index == PRS_NONE ? P.fill_row
                                            : (const void *)((const char *)P.data + (size_t)src_index * P.row_bytes);
    const bool staged = P.row_bytes <= STAGE_ROW_BYTES && P.tk != PRS_TK_LONLAT;
    if (staged) {
        const int rb = P.row_bytes;
        if (p >= 0) copy_row(stage + lane * rb, src, rb, P.vec);
        __syncwarp();
        if (trow < P.nrows) {
            const int W = (int)P.grid.width;
            const int c0 = tcol - lane;
            const int span = min(32, W - c0) * rb;
            char *dst = (char *)P.out + ((size_t)trow * W + c0) * rb;
            if ((rb & 3) == 0) {
                for (int o = lane * 4; o < span; o += 128)
                    *reinterpret_cast<uint32_t *>(dst + o) = *reinterpret_cast<const uint32_t *>(stage + o);
            } else {
                for (int o = lane; o < span; o += 32) dst[o] = stage[o];
            }
        }
        __syncwarp();
    } else if (p >= 0) {
        copy_row((char *)P.out + (size_t)p * P.row_bytes, src, P.row_bytes, P.vec);
    }
}

// Result of a point that has no neighbour (dead tiles; run-time k, no top-k list involved).
template <typename T>
__device__ __forceinline__ void emit_empty(const QueryParams &P, const IndexView<T> &ix, int64_t p, int trow, int tcol,
                                           int lane, unsigned char *stage, bool valid)
{
    if (P.mode == PRS_EPI_INFO) {
        if (p >= 0) {
            const int k = P.k;
            uint32_t *io = P.idx_out + (size_t)p * k;
            T *dout = P.dist_out ? (T *)P.dist_out + (size_t)p * k : nullptr;
            for (int i = 0; i < k; ++i) {
                io[i] = ix.n_valid;
                if (dout) dout[i] = Ar<T>::inf();
            }
            if (P.valid_out) P.valid_out[p] = valid ? 1 : 0;
        }
        if (P.flags) {
            unsigned bad = __ballot_sync(0xffffffffu, p >= 0 && !valid);
            if (lane == 0 && bad) atomicAdd(&P.flags[1], (unsigned)__popc(bad));
        }
    } else if (P.mode == PRS_EPI_NEAREST) {
        emit_nearest(P, p, trow, tcol, lane, stage, PRS_NONE);
    } else if (p >= 0) {
        if (P.acc_f64)
            gauss_empty<double>(P, p);
        else
            gauss_empty<float>(P, p);
    }
}

// Result of a searched point.  All 32 lanes of the warp call this together.
template <typename T, int K, int EPI>
__device__ __forceinline__ void emit_result(const QueryParams &P, const IndexView<T> &ix, int64_t p, int trow, int tcol,
                                            int lane, unsigned char *stage, const TopK<T, K> &top, bool valid)
{
    if (EPI == PRS_EPI_INFO) {
        emit_info<T, K>(P, ix, p, lane, top, valid);
    } else if (EPI == PRS_EPI_NEAREST) {
        emit_nearest(P, p, trow, tcol, lane, stage, top.id[0]);
    } else if (p >= 0) {
        if (P.flags) {
            bool kth = false;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if (i == P.k - 1 && top.id[i] != PRS_NONE) kth = true;
            if (kth && P.flags[0] == 0) P.flags[0] = 1;
        }
        typedef typename std::conditional<EPI == PRS_EPI_GAUSS_FF, float, double>::type Ta;
        typedef typename std::conditional<EPI == PRS_EPI_GAUSS_DD, double, float>::type Td;
        if (!valid || top.id[0] == PRS_NONE)
            gauss_empty<Ta>(P, p);
        else
            gauss_epilogue<T, Ta, Td, K>(P, ix, p, top);
    }
}

// ------------------------------------------------------------------------------------------ classification
// One warp per tile: bound the tile's points by a spherical cap and ask the L2-resident coarse table whether ANY
// source point lies within (cap + radius).  Tiles that see nothing ("dead") get their final result without a search
// (fill value / "no neighbour" sentinels) - on wide target grids that is most of the output, written as a stream;
// the other tiles are listed in live[] for the search kernel.

// Cap (unit centre c, squared chord d2m) around the valid points of a tile; false when the tile has none.
// Separable projections: the tile is the product of <= 32 longitudes (one per lane) and QROWS latitudes, and
//   u(j,l).c = sin(lat_j) sin(lat_c) + cos(lat_j) cos(lat_c) cos(lon_l - lon_c)
// grows with cos(lon_l - lon_c) (cos(lat) >= 0), so the farthest point of row j sits at the column with the smallest
// cos(dlon): two warp reductions instead of a transform per point.
template <typename T, int TK>
__device__ __forceinline__ bool tile_cap(const QueryParams &P, uint32_t tile, int trow0, int tcol, int lane,
                                         unsigned &vmask, float &cx, float &cy, float &cz, float &d2m)
{
    const float invR = (float)(1.0 / PRS_R);
    vmask = 0;
    if (TK == PRS_TK_AREA_SEP && !P.valid_extra) {
        const int W = (int)P.grid.width;
        float lc = 0.f, ls = 0.f;
        bool col_ok = false;
        if (tcol < W) {
            const ColTab<T> ct = ((const ColTab<T> *)P.col_tab)[tcol];
            lc = (float)ct.c;
            ls = (float)ct.s;
            col_ok = P.col_ok[tcol] != 0;
        }
        const int rows = min(QROWS, P.nrows - trow0);
        float rcos = 0.f, rsin = 0.f;  // lane j < rows holds tile row j
        bool row_ok = false;
        if (lane < rows) {
            const RowTab<T> rt = ((const RowTab<T> *)P.row_tab)[trow0 + lane];
            rcos = (float)rt.rc * invR;
            rsin = (float)rt.rz * invR;
            row_ok = P.row_ok[trow0 + lane] != 0;
        }
        const unsigned rmask = __ballot_sync(0xffffffffu, row_ok), cmask = __ballot_sync(0xffffffffu, col_ok);
        if (col_ok) vmask = rmask;
        if (rmask == 0 || cmask == 0) return false;
        const int jc = (rmask >> (QROWS / 2)) & 1u ? QROWS / 2 : __ffs(rmask) - 1;
        const int lcn = (cmask >> 16) & 1u ? 16 : __ffs(cmask) - 1;
        const float cc = __shfl_sync(0xffffffffu, lc, lcn), cs = __shfl_sync(0xffffffffu, ls, lcn);
        const float rcc = __shfl_sync(0xffffffffu, rcos, jc), rsc = __shfl_sync(0xffffffffu, rsin, jc);
        // chord^2 from coordinate DIFFERENCES (no cancellation: 2 - 2 u.c loses everything for a tile of a few km):
        //   |u - c|^2 = [(sin - sin_c)^2 + (cos - cos_c)^2]_lat + cos(lat) cos(lat_c) [(c - c_c)^2 + (s - s_c)^2]_lon
        // both brackets are >= 0 and cos(lat) >= 0, so the farthest point of row j is at the column with the largest
        // longitude bracket.  The table entries carry <= 1e-7 of absolute error each, i.e. <= 4e-7 on the chord;
        // tile_live adds 1e-6 to the chord before use.
        const float dlc = lc - cc, dls = ls - cs;
        const float blon = warp_max(col_ok ? dlc * dlc + dls * dls : 0.f);
        const float drs = rsin - rsc, drc = rcos - rcc;
        d2m = warp_max(row_ok ? (drs * drs + drc * drc) + rcos * rcc * blon : 0.f) * 1.0001f;
        cx = rcc * cc;
        cy = rcc * cs;
        cz = rsc;
        return true;
    }
    float ux[QROWS], uy[QROWS], uz[QROWS];
#pragma unroll
    for (int j = 0; j < QROWS; ++j) {
        const int64_t p = tile_point(P, TK, tile, trow0, tcol, j, lane);
        ux[j] = uy[j] = uz[j] = 0.f;
        if (p >= 0) {
            T qx, qy, qz;
            const bool ok = target_point<T, TK>(P, p, trow0 + j, tcol, qx, qy, qz);
            if (P.target_xyz) store_target_xyz<T>(P.target_xyz, p, qx, qy, qz, ok);
            if (ok) {
                vmask |= 1u << j;
                ux[j] = (float)qx * invR;
                uy[j] = (float)qy * invR;
                uz[j] = (float)qz * invR;
            }
        }
    }
    const unsigned has = __ballot_sync(0xffffffffu, vmask != 0);
    if (has == 0) return false;
    // cap centre: any valid point will do (the cap only has to contain the tile); prefer the middle of the tile
    const int src_lane = (has & (1u << 16)) ? 16 : __ffs(has) - 1;
    const int jj = (vmask & (1u << (QROWS / 2))) ? QROWS / 2 : __ffs(vmask) - 1;  // used on lanes with a point
    float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
    for (int j = 0; j < QROWS; ++j)
        if (j == jj) {
            mx = ux[j];
            my = uy[j];
            mz = uz[j];
        }
    cx = __shfl_sync(0xffffffffu, mx, src_lane);
    cy = __shfl_sync(0xffffffffu, my, src_lane);
    cz = __shfl_sync(0xffffffffu, mz, src_lane);
    d2m = 0.f;
#pragma unroll
    for (int j = 0; j < QROWS; ++j)
        if (vmask & (1u << j)) {
            float dx = ux[j] - cx, dy = uy[j] - cy, dz = uz[j] - cz;
            d2m = fmaxf(d2m, dx * dx + dy * dy + dz * dz);
        }
    d2m = warp_max(d2m);
    return true;
}

// Is any indexed point within `radius` of the cap (unit centre c, squared chord d2m)?  Conservative: true when unsure.
template <typename T>
__device__ __forceinline__ bool cap_has_points(const QueryParams &P, const IndexView<T> &ix, float cx, float cy, float cz,
                                               float d2m)
{
    const T dub = (T)(P.radius * P.radius);
    const float sr = sqrtf(cap_s2<T>(dub));
    // chord of the union cap on the unit sphere <= tile chord + radius chord (triangle inequality)
    const float st = (sqrtf(d2m) * 1.0001f + 1e-6f) + sr;
    const float s2t = st * st;
    if (!(s2t < 0.25f)) return true;
    const int hf = home_face(cx, cy, cz);
    float hb[4];
    face_bounds(hf, cx, cy, cz, s2t, st, hb);
    return coarse_count(ix.coarse, ix.coarse_start, ix.Gc, cx, cy, cz, s2t, st, hf, hb, bounds_inside_face(hb)) != 0;
}

// Per-pixel inverse projections are expensive (ellipsoidal polar stereographic: ~1600 instructions a pixel) and most
// tiles of a wide grid are nowhere near the swath.  Before transforming all 32 x QROWS points, bound the tile from
// five of them - its four corner pixels and the centre pixel: the tile is a small rectangle in projected
// coordinates, the map to the sphere is smooth, so over a tile of at most ~100 km the distance from the centre is
// largest at a corner up to second-order terms of relative size tile / Earth radius (< 2 %); the cap takes the
// largest corner distance plus 5 % plus 200 m.  All five valid => every pixel of the tile is valid (the valid
// region of each projection - plane, disc, rectangle - is convex in projected coordinates).  Returns true when the
// tile is certainly dead; anything doubtful (an invalid probe, a tile wider than ~120 km on the sphere, extra
// validity masks) is left to the full evaluation.
template <typename T>
__device__ __forceinline__ bool tile_dead_by_probe(const QueryParams &P, const IndexView<T> &ix, int trow0, int tcol,
                                                   int lane)
{
    if (P.valid_extra || ix.n_indexed <= 0) return false;
    const float invR = (float)(1.0 / PRS_R);
    const int W = (int)P.grid.width;
    const int c0 = tcol - lane, c1 = min(c0 + 31, W - 1);
    const int r0 = trow0, r1 = min(trow0 + QROWS, P.nrows) - 1;
    bool ok = false;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (lane < 5) {
        const int pr = lane == 4 ? (r0 + r1) >> 1 : ((lane & 2) ? r1 : r0);
        const int pc = lane == 4 ? (c0 + c1) >> 1 : ((lane & 1) ? c1 : c0);
        T qx, qy, qz;
        ok = target_point<T, PRS_TK_AREA_GEN>(P, (int64_t)pr * W + pc, pr, pc, qx, qy, qz);
        ux = (float)qx * invR, uy = (float)qy * invR, uz = (float)qz * invR;
    }
    if ((__ballot_sync(0xffffffffu, ok) & 0x1fu) != 0x1fu) return false;
    const float cx = __shfl_sync(0xffffffffu, ux, 4), cy = __shfl_sync(0xffffffffu, uy, 4), cz = __shfl_sync(0xffffffffu, uz, 4);
    const float dx = ux - cx, dy = uy - cy, dz = uz - cz;
    const float d2 = warp_max(lane < 4 ? dx * dx + dy * dy + dz * dz : 0.f);
    if (d2 > 4e-4f) return false;  // wider than ~127 km: not "small"
    const float chord = sqrtf(d2) * 1.05f + 3.2e-5f;
    return !cap_has_points<T>(P, ix, cx, cy, cz, chord * chord);
}

// tile_live: can any point of the tile have a neighbour?  vmask = this lane's valid points (bit j = tile row j).
template <typename T, int TK>
__device__ __forceinline__ bool tile_live(const QueryParams &P, const IndexView<T> &ix, uint32_t tile, int lane,
                                          unsigned &vmask_out)
{
    int trow0, tcol;
    tile_origin(P, TK, tile, lane, trow0, tcol);
    if (TK == PRS_TK_AREA_GEN && tile_dead_by_probe<T>(P, ix, trow0, tcol, lane)) {
        vmask_out = 0xFFFFFFFFu;  // every pixel of such a tile is valid
        return false;
    }
    unsigned vmask;
    float cx, cy, cz, d2m;
    bool is_live = tile_cap<T, TK>(P, tile, trow0, tcol, lane, vmask, cx, cy, cz, d2m) && ix.n_indexed > 0;
    if (is_live) is_live = cap_has_points<T>(P, ix, cx, cy, cz, d2m);
    vmask_out = vmask;
    return is_live;
}

__device__ __forceinline__ bool fill_is_pattern(const QueryParams &P, int TK)
{
    return P.mode == PRS_EPI_NEAREST && P.row_bytes <= STAGE_ROW_BYTES && (P.row_bytes & 3) == 0 && TK != PRS_TK_LONLAT;
}

// The warp writes `count` copies of `value` to dst (evict-first 16-byte stores on the aligned part).
template <typename V> __device__ __forceinline__ void warp_fill(V *dst, int64_t count, V value, int lane)
{
    constexpr int PER = 16 / (int)sizeof(V);
    const int head = (int)min((int64_t)(((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) / sizeof(V)), count);
    if (lane < head) dst[lane] = value;
    const int64_t nvec = (count - head) / PER;
    uint4 pack;
    if constexpr (sizeof(V) == 8) {
        const unsigned long long b = *reinterpret_cast<const unsigned long long *>(&value);
        pack = make_uint4((unsigned)b, (unsigned)(b >> 32), (unsigned)b, (unsigned)(b >> 32));
    } else {
        const unsigned b = *reinterpret_cast<const unsigned *>(&value);
        pack = make_uint4(b, b, b, b);
    }
    uint4 *mid = reinterpret_cast<uint4 *>(dst + head);
    for (int64_t i = lane; i < nvec; i += 32) __stcs(mid + i, pack);
    const int64_t done = head + nvec * PER;
    if (done + lane < count) dst[done + lane] = value;
}

// tile_fill: the result of a tile without neighbours.  vmask is only read when !fill_is_pattern().
template <typename T, int TK>
__device__ __forceinline__ void tile_fill(const QueryParams &P, const IndexView<T> &ix, uint32_t tile, int lane,
                                          unsigned char *stage, unsigned vmask)
{
    int trow0, tcol;
    tile_origin(P, TK, tile, lane, trow0, tcol);
    if (fill_is_pattern(P, TK)) {
        // fast path: the tile is rows of the fill pattern - stage 32 fill rows once, stream them out per tile row
        const int rb = P.row_bytes;
        copy_row(stage + lane * rb, P.fill_row, rb, P.vec);
        __syncwarp();
        const int W = (int)P.grid.width;
        const int c0 = tcol - lane;
        const int span = min(32, W - c0) * rb;
        char *dst0 = (char *)P.out + ((size_t)trow0 * W + c0) * rb;
        const size_t row_pitch = (size_t)W * rb;
        if (((reinterpret_cast<uintptr_t>(dst0) | row_pitch | (size_t)span) & 15) == 0) {
            // 16-byte stores: each lane keeps its (at most STAGE_ROW_BYTES / 16) pieces of the pattern in registers
            uint4 v[STAGE_ROW_BYTES / 16];
#pragma unroll
            for (int i = 0; i < STAGE_ROW_BYTES / 16; ++i)
                if (lane * 16 + i * 512 < span) v[i] = *reinterpret_cast<const uint4 *>(stage + lane * 16 + i * 512);
#pragma unroll
            for (int j = 0; j < QROWS; ++j) {
                if (trow0 + j < P.nrows) {
                    char *dst = dst0 + j * row_pitch;
#pragma unroll
                    for (int i = 0; i < STAGE_ROW_BYTES / 16; ++i)  // written once, never read here: evict first
                        if (lane * 16 + i * 512 < span) __stcs(reinterpret_cast<uint4 *>(dst + lane * 16 + i * 512), v[i]);
                }
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < QROWS; ++j) {
            if (trow0 + j < P.nrows) {
                char *dst = dst0 + j * row_pitch;
                for (int o = lane * 4; o < span; o += 128)
                    *reinterpret_cast<uint32_t *>(dst + o) = *reinterpret_cast<const uint32_t *>(stage + o);
            }
        }
        return;
    }
    if (TK != PRS_TK_LONLAT && P.mode != PRS_EPI_NEAREST) {
        // neighbour info / weighted modes on an area: a tile row is one contiguous stretch of every output array, and
        // a dead pixel's values are constants - the warp streams them out with 16-byte stores (per-thread stores of a
        // pixel's own 64..128 bytes touched 32 cache lines per instruction: C4's classification wrote 6.4 GB at 1.1 TB/s)
        const int W = (int)P.grid.width;
        const int c0 = tcol - lane;
        const int npx = min(32, W - c0);
        unsigned n_bad = 0;
#pragma unroll 1
        for (int j = 0; j < QROWS; ++j) {
            if (trow0 + j >= P.nrows) break;
            const size_t p0 = (size_t)(trow0 + j) * W + c0;
            if (P.mode == PRS_EPI_INFO) {
                warp_fill<uint32_t>(P.idx_out + p0 * P.k, (int64_t)npx * P.k, ix.n_valid, lane);
                if (P.dist_out) warp_fill<T>((T *)P.dist_out + p0 * P.k, (int64_t)npx * P.k, Ar<T>::inf(), lane);
                const bool valid = (vmask >> j) & 1u;
                if (P.valid_out && lane < npx) P.valid_out[p0 + lane] = valid ? 1 : 0;
                n_bad += __popc(__ballot_sync(0xffffffffu, lane < npx && !valid));
            } else if (P.acc_f64) {
                warp_fill<double>((double *)P.out + p0 * P.channels, (int64_t)npx * P.channels, P.fill, lane);
                if (P.stddev_out) {
                    warp_fill<double>((double *)P.stddev_out + p0 * P.channels, (int64_t)npx * P.channels, Ar<double>::nan(), lane);
                    warp_fill<double>((double *)P.count_out + p0 * P.channels, (int64_t)npx * P.channels, 0., lane);
                }
            } else {
                warp_fill<float>((float *)P.out + p0 * P.channels, (int64_t)npx * P.channels, (float)P.fill, lane);
                if (P.stddev_out) {
                    warp_fill<float>((float *)P.stddev_out + p0 * P.channels, (int64_t)npx * P.channels, Ar<float>::nan(), lane);
                    warp_fill<float>((float *)P.count_out + p0 * P.channels, (int64_t)npx * P.channels, 0.f, lane);
                }
            }
        }
        if (P.mode == PRS_EPI_INFO && P.flags && lane == 0 && n_bad) atomicAdd(&P.flags[1], n_bad);
        return;
    }
#pragma unroll 1
    for (int j = 0; j < QROWS; ++j) {
        const int64_t p = tile_point(P, TK, tile, trow0, tcol, j, lane);
        emit_empty<T>(P, ix, p, trow0 + j, tcol, lane, stage, (vmask >> j) & 1u);
    }
}

// Kernel A: classify each tile, fill the dead ones, list the live ones.  The fill stores (95 us of DRAM time on C2)
// overlap the classification arithmetic of other warps inside this kernel.  Tried and rejected (slower on C1, C2,
// C3 and C5, profiles/r1_overlap_experiments.txt): one persistent self-scheduling kernel doing classification and
// search, and classification / fill / search as three kernels with the fill stream on a second stream.
// Independent of the number of neighbours: a dead tile's result is sentinels / fill whatever k is.
template <typename T, int TK>
__global__ void __launch_bounds__(QWARPS * 32, 4)
    classify_kernel(const __grid_constant__ QueryParams P, const __grid_constant__ IndexRef<T> ref, uint32_t n_tiles,
                    uint32_t *__restrict__ live, uint32_t *__restrict__ n_live)
{
    __shared__ __align__(16) unsigned char stage_all[QWARPS][32 * STAGE_ROW_BYTES];
    __shared__ IndexView<T> ix;
    load_index_view<T>(ix, ref);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * QWARPS + warp;
    if (tile >= n_tiles) return;
    unsigned vmask;
    if (tile_live<T, TK>(P, ix, tile, lane, vmask)) {
        if (lane == 0) {
            live[atomicAdd(n_live, 1u)] = tile;
            if (P.tile_row_live && TK != PRS_TK_LONLAT) P.tile_row_live[tile / (uint32_t)P.tiles_x] = 1;
        }
    } else {
        tile_fill<T, TK>(P, ix, tile, lane, stage_all[warp], vmask);
    }
}

// ------------------------------------------------------------------------------------------ search
// Kernel B: the tiles that may have neighbours.  One block of QROWS warps per live tile, one point per thread
// (warp w takes row w of the tile).
//
// k >= 4 (STAGED), per tile:
//   1. every thread finds its point's home cell; every warp reads the local density of source points from the
//      coarse table and picks the half-width m of its home blocks ((2m + 1)^2 cells that should hold the k nearest:
//      cells are sized for the AVERAGE density, a swath is several times sparser at its edges than at nadir); the
//      block reduces, per row of cells, the span of cells its home blocks touch (StageCtx::umin / umax);
//   2. the rows are laid out back to back in the staging buffer and fetched with one bulk copy each (cp.async.bulk ->
//      mbarrier, issued by warp 0); meanwhile all warps copy the matching stretch of the cell table;
//   3. the threads scan their candidates from shared memory, prove the result with the closed-form cap bounds or
//      continue in the general search (global memory), and write their result.
// A tile whose points lie on different cube faces, span more than HULL_ROWS rows of cells or need more than the
// staging buffers hold is searched straight from global memory (3x3 home blocks, scan_home_global).
// k == 1: no staging (NOT STAGED) - a nearest-neighbour search reads a dozen candidates per point, and sparse targets
// like C2's 5 km grid over a 1 km swath share no cells between neighbouring points.
template <typename T, int K, int EPI>
__global__ void __launch_bounds__(QROWS * 32, PRS_QUERY_MIN_BLOCKS(K))
    query_kernel(const __grid_constant__ QueryParams P, const __grid_constant__ IndexRef<T> ref,
                 const uint32_t *__restrict__ live, const uint32_t *__restrict__ n_live)
{
    constexpr bool STAGED = K >= 4;
    extern __shared__ __align__(128) unsigned char dyn_smem[];  // staged records, cell-table stretch [nearest: row stage]
    __shared__ StageCtx sc;
    __shared__ IndexView<T> ix;
    load_index_view<T>(ix, ref);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    Rec<T> *srecs = reinterpret_cast<Rec<T> *>(dyn_smem);
    uint32_t *scs = reinterpret_cast<uint32_t *>(dyn_smem + PRS_STAGE_BYTES(K));
    unsigned char *stage = dyn_smem + warp * (32 * STAGE_ROW_BYTES);  // EPI_NEAREST (never STAGED)
    constexpr uint32_t CAP = PRS_STAGE_BYTES(K) / sizeof(Rec<T>);
    const uint32_t count = __ldg(n_live);
    const T dub = (T)(P.radius * P.radius);  // pykdtree: dtype(distance_upper_bound ** 2)
    const float s2r = cap_s2<T>(dub), sr = sqrtf(s2r);
    const int TKr = P.tk;
    if (STAGED && tid == 0) mbar_init(reinterpret_cast<uint64_t *>(&sc.bar), 32);
    uint32_t phase = 0;
#pragma unroll 1
    for (uint32_t t = blockIdx.x; t < count; t += gridDim.x) {
        const uint32_t tile = __ldg(live + t);
        int trow, tcol;
        tile_origin(P, TKr, tile, lane, trow, tcol);
        const int64_t p = tile_point(P, TKr, tile, trow, tcol, warp, lane);
        trow += warp;
        T qx = 0, qy = 0, qz = 0;
        bool valid = false;
        if (p >= 0) valid = query_point<T>(P, p, trow, tcol, qx, qy, qz);
        Home h;
        h.face = 0, h.cu = h.cv = 0;
        h.hb.u0 = h.hb.v0 = 1, h.hb.u1 = h.hb.v1 = 0;
        if (valid) h = home_cell<T>(ix, qx, qy, qz);
        TopK<T, K> top;
        top.init();
        if (!STAGED) {
            if (valid) {
                home_block<T>(ix, h, 1);
                const HomeRuns hr = home_runs<T>(ix, h);
                scan_home_global<T, K>(ix, hr, qx, qy, qz, dub, top);
                knn_finish<T, K>(ix, h, qx, qy, qz, dub, s2r, sr, top);
            }
            emit_result<T, K, EPI>(P, ix, p, trow, tcol, lane, stage, top, valid);
            continue;
        }
        // ---- 1. home blocks (half-width from the local density) and their hull
        {
            const uint32_t dens = valid ? home_density<T>(ix, h) : 0u;
            const unsigned nval = __popc(__ballot_sync(0xffffffffu, valid));
            const uint32_t dsum = __reduce_add_sync(0xffffffffu, dens);
            const int m = home_halfwidth(P.k, nval ? (float)dsum / (float)(nval * (COARSE * COARSE)) : 0.f);
            if (valid) home_block<T>(ix, h, m);
        }
        const bool has_block = valid && h.hb.u0 <= h.hb.u1;
        if (tid < HULL_ROWS) {
            sc.umin[tid] = 0x7fffffff;
            sc.umax[tid] = -1;
        }
        if (tid == 0) {
            sc.vlo = 0x7fffffff;
            sc.vhi = -1;
            sc.faces = 0;
        }
        __syncthreads();  // also: every thread is done with the previous tile's staged records
        {
            unsigned todo = __ballot_sync(0xffffffffu, has_block);
            if (todo) {
                const int wlo = __reduce_min_sync(0xffffffffu, has_block ? h.hb.v0 : 0x7fffffff);
                const int whi = __reduce_max_sync(0xffffffffu, has_block ? h.hb.v1 : -1);
                const unsigned wf = __reduce_or_sync(0xffffffffu, has_block ? 1u << h.face : 0u);
                if (lane == 0) {
                    atomicMin(&sc.vlo, wlo);
                    atomicMax(&sc.vhi, whi);
                    atomicOr(&sc.faces, wf);
                }
            }
            // lanes with the same home row contribute one span; rows v0..v1 of the block all take it
            while (todo) {
                const int src = __ffs(todo) - 1;
                const int v = __shfl_sync(0xffffffffu, h.cv, src);
                const bool mine = has_block && h.cv == v;
                const unsigned grp = __ballot_sync(0xffffffffu, mine);
                const int lo = __reduce_min_sync(0xffffffffu, mine ? h.hb.u0 : 0x7fffffff);
                const int hi = __reduce_max_sync(0xffffffffu, mine ? h.hb.u1 : -1);
                const int r0 = __shfl_sync(0xffffffffu, h.hb.v0, src), r1 = __shfl_sync(0xffffffffu, h.hb.v1, src);
                if (lane < 2 * HOME_MAX + 1 && r0 + lane <= r1) {
                    atomicMin(&sc.umin[(r0 + lane) & (HULL_ROWS - 1)], lo);
                    atomicMax(&sc.umax[(r0 + lane) & (HULL_ROWS - 1)], hi);
                }
                todo &= ~grp;
            }
        }
        __syncthreads();
        // ---- 2. layout (every warp computes it: lanes hold rows lane and lane + 32), bulk copies (warp 0), cell table
        {
            const int vlo = sc.vlo, vhi = sc.vhi;
            const unsigned faces = sc.faces;
            bool ok = vhi >= vlo && (vhi - vlo) < HULL_ROWS && __popc(faces) == 1;
            uint32_t gs[2] = {0, 0}, n[2] = {0, 0}, w[2] = {0, 0};
            const uint32_t *rowp[2] = {nullptr, nullptr};
            if (ok) {
                const int face = __ffs(faces) - 1;
                const FaceTable &ft = ix.fine;
                const int fw = ft.fw[face];
                const uint32_t *row0 = ix.cell_start + ft.base[face] - (int64_t)ft.fv0[face] * fw - ft.fu0[face];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int v = vlo + lane + 32 * j;
                    if (v <= vhi) {
                        const int slot = v & (HULL_ROWS - 1);
                        const int u0 = sc.umin[slot], u1 = sc.umax[slot];
                        if (u1 >= u0) {
                            rowp[j] = row0 + (int64_t)v * fw + u0;
                            w[j] = (uint32_t)(u1 - u0 + 2);
                            gs[j] = __ldg(rowp[j]);
                            n[j] = __ldg(rowp[j] + (u1 - u0 + 1)) - gs[j];
                        }
                    }
                }
            }
            // exclusive prefixes over the (up to 64) rows: records and cell-table entries
            uint32_t a = n[0], b = n[1], ca = w[0], cb = w[1];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
                const uint32_t tc = __shfl_up_sync(0xffffffffu, ca, o), td = __shfl_up_sync(0xffffffffu, cb, o);
                if (lane >= o) a += ta, b += tb, ca += tc, cb += td;
            }
            const uint32_t tot0 = __shfl_sync(0xffffffffu, a, 31), total = tot0 + __shfl_sync(0xffffffffu, b, 31);
            const uint32_t ctot0 = __shfl_sync(0xffffffffu, ca, 31), ctotal = ctot0 + __shfl_sync(0xffffffffu, cb, 31);
            ok = ok && total > 0 && total <= CAP && ctotal <= (uint32_t)STAGE_CELLS;
            if (ok) {
                const uint32_t base[2] = {a - n[0], tot0 + b - n[1]};
                const uint32_t cbase[2] = {ca - w[0], ctot0 + cb - w[1]};
                if (warp == 0) {
                    fence_proxy_async();  // the buffer was read (generic proxy) for the previous tile
                    mbar_arrive_expect_tx(reinterpret_cast<uint64_t *>(&sc.bar), (n[0] + n[1]) * (uint32_t)sizeof(Rec<T>));
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int v = vlo + lane + 32 * j;
                    if (v <= vhi) {
                        if (warp == 0) {
                            const int slot = v & (HULL_ROWS - 1);
                            // (umin / umax are reset for the next tile by whoever gets there first: what the scan
                            // needs of them is kept in fields that are only written here)
                            sc.gstart[slot] = gs[j];
                            sc.sbase[slot] = base[j];
                            sc.csoff[slot] = (int)cbase[j] - sc.umin[slot];
                            if (n[j])
                                bulk_copy_g2s(srecs + base[j], ix.recs + gs[j], n[j] * (uint32_t)sizeof(Rec<T>),
                                              reinterpret_cast<uint64_t *>(&sc.bar));
                        }
                    }
                }
                // the cell-table stretch: warp w copies rows w, w + QROWS, ... (row r is held by lane r % 32, half r / 32)
                const int nrows = vhi - vlo + 1;
                for (int r = warp; r < nrows; r += QROWS) {
                    const int src = r & 31, half = r >> 5;
                    const uint32_t *gp = reinterpret_cast<const uint32_t *>(
                        __shfl_sync(0xffffffffu, (unsigned long long)(half ? rowp[1] : rowp[0]), src));
                    const uint32_t wr = __shfl_sync(0xffffffffu, half ? w[1] : w[0], src);
                    const uint32_t cbr = __shfl_sync(0xffffffffu, half ? cbase[1] : cbase[0], src);
                    for (uint32_t j = lane; j < wr; j += 32) scs[cbr + j] = __ldg(gp + j);
                }
            }
            // written here and nowhere else: a thread that is slow to read it cannot see the next tile's value,
            // because nobody gets here again before two more block-wide barriers
            if (tid == 0) sc.staged = ok ? 1 : 0;
        }
        __syncthreads();
        // ---- 3. search
        const bool staged = sc.staged != 0;
        if (staged) {
            mbar_wait(reinterpret_cast<uint64_t *>(&sc.bar), phase);
            phase ^= 1u;
        }
        if (tid == 0) PRS_STAT(staged ? 14 : 15, 1);  // tiles staged / not staged
        if (valid) {
            PRS_STAT(0, 1);  // valid points searched
            PRS_STAT(16 + min(7, max(h.cv - h.hb.v0, h.hb.v1 - h.cv)), 1);  // half-width of the home block
            if (staged) {
                scan_home_staged<T, K>(sc, srecs, scs, h, qx, qy, qz, dub, top);
            } else {
                home_block<T>(ix, h, 1);
                const HomeRuns hr = home_runs<T>(ix, h);
                scan_home_global<T, K>(ix, hr, qx, qy, qz, dub, top);
            }
            knn_finish<T, K>(ix, h, qx, qy, qz, dub, s2r, sr, top);
        }
        emit_result<T, K, EPI>(P, ix, p, trow, tcol, lane, stage, top, valid);
    }
}

// Coverage of a target at PRS_COVER_CELLS^2 cells per cube face: marks every cell that the cap (point, radius)
// touches, dilated by one cell, so that the index build of this rank may drop source points elsewhere.
template <typename T, int TK>
__global__ void coverage_kernel(const __grid_constant__ QueryParams P, uint8_t *__restrict__ cover)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_t) return;
    int trow = 0, tcol = 0;
    if (TK != PRS_TK_LONLAT) {
        const int64_t W = P.grid.width;
        trow = (int)(p / W);
        tcol = (int)(p - (int64_t)trow * W);
    }
    T qx, qy, qz;
    if (!target_point<T, TK>(P, p, trow, tcol, qx, qy, qz)) return;
    const float invR = (float)(1.0 / PRS_R);
    const float fx = (float)qx * invR, fy = (float)qy * invR, fz = (float)qz * invR;
    const T dub = (T)(P.radius * P.radius);
    const float s2 = cap_s2<T>(dub), s = sqrtf(s2);
    const int CG = PRS_COVER_CELLS;
#pragma unroll 1
    for (int f = 0; f < 6; ++f) {
        float b[4];
        if (!face_bounds(f, fx, fy, fz, s2, s, b)) continue;
        CellRect c = bounds_to_cells(b, CG);
        c.u0 = max(c.u0 - 1, 0);
        c.v0 = max(c.v0 - 1, 0);
        c.u1 = min(c.u1 + 1, CG - 1);
        c.v1 = min(c.v1 + 1, CG - 1);
        for (int iv = c.v0; iv <= c.v1; ++iv)
            for (int iu = c.u0; iu <= c.u1; ++iu) {
                uint8_t *cell = cover + ((size_t)f * CG + iv) * CG + iu;
                if (!*cell) *cell = 1;
            }
    }
}

// Per-row / per-column tables of a separable projection (longlat, eqc, merc): lon depends on the column
// only and lat on the row only, so the per-pixel transform collapses to two multiplies.
template <typename T>
__global__ void area_tables_kernel(const __grid_constant__ prs_proj_params P, const __grid_constant__ prs_area_grid g, int64_t row0, int64_t row_block,
                                   int64_t row_stride, int64_t nrows, ColTab<T> *col_tab, uint8_t *col_ok,
                                   RowTab<T> *row_tab, uint8_t *row_ok)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const T NaN = Ar<T>::nan();
    if (i < g.width) {
        T lon, lat;
        area_pixel_lonlat<T>(P, g, row0, i, lon, lat);
        bool ok = (lon >= (T)-180) && (lon <= (T)180);
        T s = NaN, c = NaN;
        if (ok) {
            if (sizeof(T) == 4) {
                float fs, fc;
                trig_f32((float)lon, fs, fc);
                s = (T)fs;
                c = (T)fc;
            } else {
                double ds, dc;
                trig_f64((double)lon, ds, dc);
                s = (T)ds;
                c = (T)dc;
            }
        }
        col_tab[i].c = c;
        col_tab[i].s = s;
        col_ok[i] = ok;
    } else if (i < g.width + nrows) {
        int64_t r = i - g.width;
        T lon, lat;
        area_pixel_lonlat<T>(P, g, global_row(row0, row_block, row_stride, r), 0, lon, lat);
        bool ok = (lat <= (T)90) && (lat >= (T)-90);
        T rc = NaN, rz = NaN;
        if (ok) {
            if (sizeof(T) == 4) {
                float fs, fc;
                trig_f32((float)lat, fs, fc);
                rc = (T)__fmul_rn(6370997.0f, fc);
                rz = (T)__fmul_rn(6370997.0f, fs);
            } else {
                double ds, dc;
                trig_f64((double)lat, ds, dc);
                rc = (T)__dmul_rn(6370997.0, dc);
                rz = (T)__dmul_rn(6370997.0, ds);
            }
        }
        row_tab[r].rc = rc;
        row_tab[r].rz = rz;
        row_ok[r] = ok;
    }
}

}  // namespace prs
