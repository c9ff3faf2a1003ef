// Bounded exact k-NN against the cube-face cell index, with the sample gather / weighting fused into the
// epilogue.  Replaces KDTree.query (kd_tree.py:545-548) + get_sample_from_neighbour_info (kd_tree.py:566-818).
#pragma once
#include "prs_common.cuh"
#include "prs_index.cuh"
#include "prs_proj.cuh"

namespace prs {

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
    __device__ __forceinline__ void offer(T d2, uint32_t idx, T dub)
    {
        if (!(d2 < dub)) return;  // strict radius test in the tree dtype (pykdtree)
        if (d2 < d[K - 1] || (d2 == d[K - 1] && idx < id[K - 1])) {
            d[K - 1] = d2;
            id[K - 1] = idx;
#pragma unroll
            for (int j = K - 1; j > 0; --j) {
                bool sw = d[j] < d[j - 1] || (d[j] == d[j - 1] && id[j] < id[j - 1]);
                T td = d[j - 1];
                uint32_t ti = id[j - 1];
                d[j - 1] = sw ? d[j] : td;
                id[j - 1] = sw ? id[j] : ti;
                d[j] = sw ? td : d[j];
                id[j] = sw ? ti : id[j];
            }
        }
    }
    __device__ __forceinline__ bool full() const { return id[K - 1] != PRS_NONE; }
};

template <typename T> struct IndexView {
    const Rec<T> *recs;
    const uint32_t *cell_start;
    const uint32_t *coarse_start;
    const uint32_t *rank;  // NULL when rank == source index
    FaceTable fine, coarse;
    int G, Gc;
    uint32_t n_valid;
    uint32_t first_valid_source;
};

// Cell rectangle of one cube face that contains every point within the cap {p : |p - q|^2 <= bound}.
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

// (fx, fy, fz): unit query vector.  s2: upper bound of sin^2 of the cap's angular radius.
// Returns false when the cap cannot touch this face.  Closed form: the gnomonic image of a spherical cap is
// a conic  w^T (q q^T - cos^2 I) w = 0,  w = (1, u, v);  its axis-aligned bounding box is
//   u = (qa*qb -/+ s*sqrt(qa^2 + qb^2 - s^2)) / (qa^2 - s^2)      (same for v with qc).
__device__ __forceinline__ bool face_rect(int face, float fx, float fy, float fz, float s2, float s, int G,
                                          CellRect &r)
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
    // a point of the cap on this face has its dominant component >= 1/sqrt(3); the cap's angular radius
    // is <= 1.2*s for s <= 0.8
    if (qa < 0.57735f - 1.2f * s - 1e-3f) return false;
    float den = qa * qa - s2;
    float u0 = -1.f, u1 = 1.f, v0 = -1.f, v1 = 1.f;
    if (den > 0.02f) {
        float inv = 1.0f / den;
        float ru = s * sqrtf(fmaxf(qa * qa + qb * qb - s2, 0.f));
        float rv = s * sqrtf(fmaxf(qa * qa + qc * qc - s2, 0.f));
        u0 = (qa * qb - ru) * inv;
        u1 = (qa * qb + ru) * inv;
        v0 = (qa * qc - rv) * inv;
        v1 = (qa * qc + rv) * inv;
        // float32 slack: cell keys were computed from float32 coordinates (<= 1e-7 relative)
        u0 -= 3e-6f * (1.0f + fabsf(u0));
        u1 += 3e-6f * (1.0f + fabsf(u1));
        v0 -= 3e-6f * (1.0f + fabsf(v0));
        v1 += 3e-6f * (1.0f + fabsf(v1));
        if (u1 < -1.f || u0 > 1.f || v1 < -1.f || v0 > 1.f) return false;
    }
    r.u0 = cell_coord(warp_uv(u0) - 2e-6f, G);
    r.u1 = cell_coord(warp_uv(u1) + 2e-6f, G);
    r.v0 = cell_coord(warp_uv(v0) - 2e-6f, G);
    r.v1 = cell_coord(warp_uv(v1) + 2e-6f, G);
    return true;
}

// sin^2(theta) <= (d/R)^2 for chord d, inflated so that float32 rounding can never exclude a candidate
template <typename T> __device__ __forceinline__ float cap_s2(T d2)
{
    const float invR2 = (float)(1.0 / (PRS_R * PRS_R));
    return (float)d2 * invR2 * 1.0002f + 1e-13f;
}

// Number of indexed points inside the cap's cell rectangles of the COARSE table (an upper bound of the number of
// points within the cap; 0 proves there is none).
__device__ __forceinline__ uint32_t coarse_count(const FaceTable &ct, const uint32_t *__restrict__ coarse_start, int Gc,
                                                 float fx, float fy, float fz, float s2, float s)
{
    uint32_t total = 0;
#pragma unroll 1
    for (int f = 0; f < 6; ++f) {
        CellRect c;
        if (!face_rect(f, fx, fy, fz, s2, s, Gc, c)) continue;
        if (!clip_rect(ct, f, c)) continue;
        const uint32_t *row = coarse_start + ct.base[f] + (int64_t)(c.v0 - ct.fv0[f]) * ct.fw[f] - ct.fu0[f];
        for (int iv = c.v0; iv <= c.v1; ++iv, row += ct.fw[f]) total += __ldg(row + c.u1 + 1) - __ldg(row + c.u0);
    }
    return total;
}

template <typename T, int K>
__device__ __forceinline__ void scan_run(const Rec<T> *__restrict__ recs, uint32_t s, uint32_t e, T qx, T qy, T qz,
                                         T dub, TopK<T, K> &top)
{
    for (uint32_t i = s; i < e; ++i) {
        Rec<T> r = load_rec(recs + i);
        top.offer(dist2<T>(qx, qy, qz, r.x, r.y, r.z), r.idx, dub);
    }
}

// Scan the cells [u0,u1] x [v0,v1] of `face` (already clipped to the table), skipping the visited block
// [hu0,hu1] x [hv0,hv1] when skip is set.  One row = one contiguous run of records (two table reads).
template <typename T, int K>
__device__ __forceinline__ void scan_rect(const IndexView<T> &ix, int face, const CellRect &c, bool skip,
                                          const CellRect &h, T qx, T qy, T qz, T dub, TopK<T, K> &top)
{
    const FaceTable &ft = ix.fine;
    const uint32_t *row = ix.cell_start + ft.base[face] + (int64_t)(c.v0 - ft.fv0[face]) * ft.fw[face] - ft.fu0[face];
    for (int iv = c.v0; iv <= c.v1; ++iv, row += ft.fw[face]) {
        if (skip && iv >= h.v0 && iv <= h.v1) {
            if (c.u0 < h.u0) {
                int e1 = min(c.u1, h.u0 - 1);
                scan_run<T, K>(ix.recs, __ldg(row + c.u0), __ldg(row + e1 + 1), qx, qy, qz, dub, top);
            }
            if (c.u1 > h.u1) {
                int s1 = max(c.u0, h.u1 + 1);
                scan_run<T, K>(ix.recs, __ldg(row + s1), __ldg(row + c.u1 + 1), qx, qy, qz, dub, top);
            }
        } else {
            scan_run<T, K>(ix.recs, __ldg(row + c.u0), __ldg(row + c.u1 + 1), qx, qy, qz, dub, top);
        }
    }
}

// Exact bounded k-NN of one query point.  dub = (T)(radius^2); candidates need d2 < dub.
template <typename T, int K>
__device__ void knn_search(const IndexView<T> &ix, T qx, T qy, T qz, T dub, TopK<T, K> &top)
{
    const float invR = (float)(1.0 / PRS_R);
    const float fx = (float)qx * invR, fy = (float)qy * invR, fz = (float)qz * invR;
    const int G = ix.G;
    // ---- coarse emptiness test over the whole radius: most target pixels of a wide grid see no source at all
    const float s2r = cap_s2<T>(dub);
    const float sr = sqrtf(s2r);
    if (coarse_count(ix.coarse, ix.coarse_start, ix.Gc, fx, fy, fz, s2r, sr) == 0) return;
    // ---- phase A: grow a block of cells around the query's own cell until it holds k candidates
    int hface;
    float hu, hv;
    face_uv(fx, fy, fz, hface, hu, hv);
    const int cu = cell_coord(warp_uv(hu), G), cv = cell_coord(warp_uv(hv), G);
    CellRect hb;  // visited block on the home face (empty: u0 > u1)
    hb.u0 = hb.v0 = 1;
    hb.u1 = hb.v1 = 0;
    CellRect rr;  // everything within the radius on the home face, clipped to the table
    if (face_rect(hface, fx, fy, fz, s2r, sr, G, rr) && clip_rect(ix.fine, hface, rr)) {
        int m = (K == 1) ? 0 : 1;
        for (;;) {
            CellRect nb;
            nb.u0 = max(cu - m, rr.u0);
            nb.u1 = min(cu + m, rr.u1);
            nb.v0 = max(cv - m, rr.v0);
            nb.v1 = min(cv + m, rr.v1);
            if (nb.u0 <= nb.u1 && nb.v0 <= nb.v1) {
                scan_rect<T, K>(ix, hface, nb, hb.u0 <= hb.u1, hb, qx, qy, qz, dub, top);
                hb = nb;
            }
            if (top.full() || (nb.u0 == rr.u0 && nb.u1 == rr.u1 && nb.v0 == rr.v0 && nb.v1 == rr.v1)) break;
            if (cu - m <= rr.u0 && cu + m >= rr.u1 && cv - m <= rr.v0 && cv + m >= rr.v1) break;
            m = 2 * m + 1;
        }
    }
    // ---- phase B: everything that can still beat (or tie) the current k-th distance, on every face
    T bound = top.full() ? top.d[K - 1] : dub;
    const float s2b = cap_s2<T>(bound);
    const float sb = sqrtf(s2b);
#pragma unroll 1
    for (int f = 0; f < 6; ++f) {
        CellRect c;
        if (!face_rect(f, fx, fy, fz, s2b, sb, G, c)) continue;
        if (!clip_rect(ix.fine, f, c)) continue;
        bool home = (f == hface) && (hb.u0 <= hb.u1);
        if (home && c.u0 >= hb.u0 && c.u1 <= hb.u1 && c.v0 >= hb.v0 && c.v1 <= hb.v1) continue;
        scan_rect<T, K>(ix, f, c, home, hb, qx, qy, qz, dub, top);
    }
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
#define PRS_EPI_GAUSS 2

struct QueryParams {
    // target
    int64_t n_t;
    const void *tlon, *tlat;
    const uint8_t *valid_extra;
    prs_proj_params proj;
    prs_area_grid grid;
    int64_t row0;
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
    int64_t row_bytes;
    int vec;  // copy granularity in bytes (1, 2, 4, 8, 16)
    // gauss
    int channels;
    int data_f64;
    int acc_f64;
    const double *sigma2;  // device [channels]: float(sigma)**2
    int uniform_sigma;
    double fill;
    void *stddev_out, *count_out;
};

__device__ __forceinline__ void copy_row(void *dst, const void *src, int64_t bytes, int vec)
{
    switch (vec) {
    case 16:
        for (int64_t o = 0; o < bytes; o += 16)
            *reinterpret_cast<uint4 *>((char *)dst + o) = __ldg(reinterpret_cast<const uint4 *>((const char *)src + o));
        break;
    case 8:
        for (int64_t o = 0; o < bytes; o += 8)
            *reinterpret_cast<uint2 *>((char *)dst + o) = __ldg(reinterpret_cast<const uint2 *>((const char *)src + o));
        break;
    case 4:
        for (int64_t o = 0; o < bytes; o += 4)
            *reinterpret_cast<uint32_t *>((char *)dst + o) =
                __ldg(reinterpret_cast<const uint32_t *>((const char *)src + o));
        break;
    case 2:
        for (int64_t o = 0; o < bytes; o += 2)
            *reinterpret_cast<uint16_t *>((char *)dst + o) =
                __ldg(reinterpret_cast<const uint16_t *>((const char *)src + o));
        break;
    default:
        for (int64_t o = 0; o < bytes; ++o) ((uint8_t *)dst)[o] = __ldg((const uint8_t *)src + o);
    }
}

// _resample_with_weights (kd_tree.py:741-818) + _calculate_uncertainty (kd_tree.py:821-859) for one target
// point and one channel.  Tw: dtype the weight function is evaluated in (distance dtype);
// Ta: accumulation / output dtype.  Neighbour i: source row src[i] (PRS_NONE = missing), weight w[i].
template <typename Ta, typename Td, int K>
__device__ __forceinline__ void weighted_channel(const Td *__restrict__ data, int channels, int c, const uint32_t *src,
                                                 const Ta *w, int k, uint32_t first_valid_source, Ta fill, Ta &res_out,
                                                 Ta &std_out, int &count_out, bool want_uncert)
{
    Ta res = 0, norm = 0;
    Ta vals[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        if (i < k) {
            bool missing = src[i] == PRS_NONE;
            // missing neighbours gather new_data[0] and get weight 0 (kd_tree.py:754-759,802)
            uint32_t row = missing ? first_valid_source : src[i];
            Ta v = (Ta)__ldg(data + (size_t)row * channels + c);
            Ta wt = missing ? Ar<Ta>::mul((Ta)0, w[i]) : w[i];
            vals[i] = v;
            res = Ar<Ta>::add(res, Ar<Ta>::mul(wt, v));
            norm = Ar<Ta>::add(norm, wt);
        }
    }
    if (norm > 0)
        res = Ar<Ta>::div(res, norm);
    else
        res = fill;  // also reached for NaN norm, like numpy's boolean masks (norm > 0 is False)
    res_out = res;
    if (want_uncert) {
        int count = 0;
        Ta norm_sqr = 0, sd = 0;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            if (i < k) {
                bool missing = src[i] == PRS_NONE;
                Ta wt = missing ? Ar<Ta>::mul((Ta)0, w[i]) : w[i];
                count += missing ? 0 : 1;
                norm_sqr = Ar<Ta>::add(norm_sqr, Ar<Ta>::mul(wt, wt));
                Ta val = missing ? Ar<Ta>::mul((Ta)0, vals[i]) : vals[i];
                Ta dv = Ar<Ta>::sub(val, res);
                sd = Ar<Ta>::add(sd, Ar<Ta>::mul(wt, Ar<Ta>::mul(dv, dv)));
            }
        }
        count_out = count;
        if (count > 1)
            std_out = Ar<Ta>::sqrt(Ar<Ta>::mul(Ar<Ta>::div(norm, Ar<Ta>::sub(Ar<Ta>::mul(norm, norm), norm_sqr)), sd));
        else
            std_out = Ar<Ta>::nan();
    }
}

template <typename T, typename Ta, typename Td, int K>
__device__ __forceinline__ void gauss_epilogue(const QueryParams &P, const IndexView<T> &ix, int64_t p,
                                               const TopK<T, K> &top, bool valid)
{
    const int C = P.channels;
    const bool want_uncert = P.stddev_out != nullptr;
    Ta *out = (Ta *)P.out + (size_t)p * C;
    Ta *sd_out = want_uncert ? (Ta *)P.stddev_out + (size_t)p * C : nullptr;
    Ta *cnt_out = want_uncert ? (Ta *)P.count_out + (size_t)p * C : nullptr;
    if (!valid) {  // not in valid_output_index: fill (kd_tree.py:719-723), NaN / 0 uncertainty (kd_tree.py:865-868)
        for (int c = 0; c < C; ++c) {
            out[c] = (Ta)P.fill;
            if (want_uncert) {
                sd_out[c] = Ar<Ta>::nan();
                cnt_out[c] = 0;
            }
        }
        return;
    }
    T dist[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        // missing neighbours: distance 1 (kd_tree.py:767-768)
        dist[i] = top.id[i] == PRS_NONE ? (T)1 : Ar<T>::sqrt(top.d[i]);
    }
    Ta w[K];
    for (int c = 0; c < C; ++c) {
        if (c == 0 || !P.uniform_sigma) {
            T s2 = (T)P.sigma2[c];  // float(sigma)**2 as a weak Python scalar -> distance dtype
#pragma unroll
            for (int i = 0; i < K; ++i) {
                // np.exp(-r ** 2 / float(sigma) ** 2) evaluated in the distance dtype (kd_tree.py:165)
                T r2 = Ar<T>::mul(dist[i], dist[i]);
                w[i] = (Ta)Ar<T>::exp(Ar<T>::div(-r2, s2));
            }
        }
        Ta res, sd = 0;
        int cnt = 0;
        weighted_channel<Ta, Td, K>((const Td *)P.data, C, c, top.id, w, P.k, ix.first_valid_source, (Ta)P.fill, res,
                                    sd, cnt, want_uncert);
        out[c] = res;
        if (want_uncert) {
            sd_out[c] = sd;
            cnt_out[c] = (Ta)cnt;
        }
    }
}

template <typename T, int K, int TK>
__global__ void __launch_bounds__(256) query_kernel(const __grid_constant__ QueryParams P, const IndexView<T> ix)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_t) return;
    // ---- target coordinates (kd_tree.py:513-534)
    bool valid;
    T qx = 0, qy = 0, qz = 0;
    if (TK == PRS_TK_LONLAT) {
        T lon = ((const T *)P.tlon)[p], lat = ((const T *)P.tlat)[p];
        valid = lonlat_in_range<T>(lon, lat);
        if (valid) lonlat_to_xyz(lon, lat, qx, qy, qz);
    } else if (TK == PRS_TK_AREA_SEP) {
        const int64_t W = P.grid.width;
        int64_t row = p / W, col = p - row * W;
        valid = P.col_ok[col] && P.row_ok[row];
        ColTab<T> ct = ((const ColTab<T> *)P.col_tab)[col];
        RowTab<T> rt = ((const RowTab<T> *)P.row_tab)[row];
        qx = Ar<T>::mul(rt.rc, ct.c);
        qy = Ar<T>::mul(rt.rc, ct.s);
        qz = rt.rz;
    } else {
        const int64_t W = P.grid.width;
        int64_t row = p / W, col = p - row * W;
        T lon, lat;
        area_pixel_lonlat<T>(P.proj, P.grid, P.row0 + row, col, lon, lat);
        valid = lonlat_in_range<T>(lon, lat);
        if (valid) lonlat_to_xyz(lon, lat, qx, qy, qz);
    }
    if (P.valid_extra) valid = valid && P.valid_extra[p];
    // ---- search
    TopK<T, K> top;
    top.init();
    const T dub = (T)(P.radius * P.radius);  // pykdtree: dtype(distance_upper_bound ** 2)
    if (valid && ix.recs != nullptr) knn_search<T, K>(ix, qx, qy, qz, dub, top);
    // ---- epilogue
    if (P.mode == PRS_EPI_INFO) {
        const int k = P.k;
        uint32_t *io = P.idx_out + (size_t)p * k;
        T *dout = P.dist_out ? (T *)P.dist_out + (size_t)p * k : nullptr;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            if (i < k) {
                uint32_t s = top.id[i];
                uint32_t r = s == PRS_NONE ? ix.n_valid : (ix.rank ? __ldg(ix.rank + s) : s);
                io[i] = r;
                if (dout) dout[i] = s == PRS_NONE ? Ar<T>::inf() : Ar<T>::sqrt(top.d[i]);
            }
        }
        if (P.valid_out) P.valid_out[p] = valid ? 1 : 0;
        if (P.flags) {
            bool kth = false;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if (i == k - 1) kth = top.id[i] != PRS_NONE;
            if (kth && P.flags[0] == 0) P.flags[0] = 1;
            if (!valid) atomicAdd(&P.flags[1], 1u);
        }
    } else if (P.mode == PRS_EPI_NEAREST) {
        const void *src = top.id[0] == PRS_NONE ? P.fill_row : (const void *)((const char *)P.data + (size_t)top.id[0] * P.row_bytes);
        copy_row((char *)P.out + (size_t)p * P.row_bytes, src, P.row_bytes, P.vec);
    } else {
        if (P.flags) {
            bool kth = false;
#pragma unroll
            for (int i = 0; i < K; ++i)
                if (i == P.k - 1) kth = top.id[i] != PRS_NONE;
            if (kth && P.flags[0] == 0) P.flags[0] = 1;
        }
        if (P.acc_f64) {
            if (P.data_f64)
                gauss_epilogue<T, double, double, K>(P, ix, p, top, valid);
            else
                gauss_epilogue<T, double, float, K>(P, ix, p, top, valid);
        } else {
            gauss_epilogue<T, float, float, K>(P, ix, p, top, valid);
        }
    }
}

// Per-row / per-column tables of a separable projection (longlat, eqc, merc): lon depends on the column
// only and lat on the row only, so the per-pixel transform collapses to two multiplies.
template <typename T>
__global__ void area_tables_kernel(const prs_proj_params P, const prs_area_grid g, int64_t row0, int64_t nrows,
                                   ColTab<T> *col_tab, uint8_t *col_ok, RowTab<T> *row_tab, uint8_t *row_ok)
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
        area_pixel_lonlat<T>(P, g, row0 + r, 0, lon, lat);
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
