// C-ABI of the B200-native kd-tree neighbour resampling path (see include/prs_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#define PRS_WITH_INDEX_BUILD
#include "prs_common.cuh"
#include "prs_index.cuh"
#include "prs_proj.cuh"
#include "prs_launch.cuh"
#include "prs_query.cuh"
#include "prs_scan.cuh"
#include "prs_bilinear.cuh"

using namespace prs;

namespace {

constexpr int TPB = 256;
inline unsigned blocks_for(int64_t n, int tpb = TPB) { return (unsigned)((n + tpb - 1) / tpb); }

// ------------------------------------------------------------------------------------------ small kernels
template <typename T>
__global__ void lonlat_to_xyz_kernel(const T *__restrict__ lon, const T *__restrict__ lat, int64_t n, T *__restrict__ xyz)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T x, y, z;
    lonlat_to_xyz(lon[i], lat[i], x, y, z);
    xyz[3 * i] = x;
    xyz[3 * i + 1] = y;
    xyz[3 * i + 2] = z;
}

template <typename T>
__global__ void area_lonlats_kernel(const __grid_constant__ prs_proj_params P, const __grid_constant__ prs_area_grid g, int64_t row0, int64_t row_step,
                                    int64_t nrows, int64_t col0, int64_t col_step, int64_t ncols, T *lon_out, T *lat_out,
                                    T *xyz_out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * ncols) return;
    int64_t r = i / ncols, c = i - r * ncols;
    T lon, lat;
    area_pixel_lonlat<T>(P, g, row0 + r * row_step, col0 + c * col_step, lon, lat);
    if (lon_out) lon_out[i] = lon;
    if (lat_out) lat_out[i] = lat;
    if (xyz_out) {
        T x, y, z;
        lonlat_to_xyz(lon, lat, x, y, z);
        xyz_out[3 * i] = x;
        xyz_out[3 * i + 1] = y;
        xyz_out[3 * i + 2] = z;
    }
}

// get_lonlats of a whole area for ONE RANK of a row-sharded run: tiles of 16 x 16 source pixels whose surroundings
// hold no cell of the rank's coverage mask (prs_target_coverage) are not transformed - their pixels would be dropped
// by the index build anyway - and come out as NaN (= invalid).  The test is conservative: the tile's four corners and
// centre must all be valid and lie on one cube face, the bounding box of their coverage cells (~20 km each), grown by
// one cell, must stay inside the face and span at most 10 cells (a tile under ~160 km: the bulge of its edges beyond
// the corners' box is far below one cell), and no cell of that box may be set.
__global__ void __launch_bounds__(128) area_tile_keep_kernel(const __grid_constant__ prs_proj_params P,
                                                            const __grid_constant__ prs_area_grid g,
                                                            const uint8_t *__restrict__ cover, int tiles_x, int tiles_y,
                                                            uint8_t *__restrict__ tile_keep)
{
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= tiles_x * tiles_y) return;
    const int64_t W = g.width, H = g.height;
    const int64_t r0 = (int64_t)(tile / tiles_x) * 16, c0 = (int64_t)(tile % tiles_x) * 16;
    const int64_t r1 = min(r0 + 15, H - 1), c1 = min(c0 + 15, W - 1);
    bool keep = false;
    int face0 = -1, u0 = 0, u1 = 0, v0 = 0, v1 = 0;
#pragma unroll 1
    for (int i = 0; i < 5 && !keep; ++i) {
        const int64_t pr = i == 4 ? (r0 + r1) / 2 : ((i & 2) ? r1 : r0), pc = i == 4 ? (c0 + c1) / 2 : ((i & 1) ? c1 : c0);
        double lon, lat;
        area_pixel_lonlat<double>(P, g, pr, pc, lon, lat);
        if (!(lon >= -180.0 && lon <= 180.0 && lat >= -90.0 && lat <= 90.0)) {
            keep = true;
            break;
        }
        float slo, clo, sla, cla;
        __sincosf((float)lon * 0.017453292f, &slo, &clo);
        __sincosf((float)lat * 0.017453292f, &sla, &cla);
        int face, iu, iv;
        face_cell_of(cla * clo, cla * slo, sla, PRS_COVER_CELLS, face, iu, iv);
        if (i == 0) {
            face0 = face;
            u0 = u1 = iu;
            v0 = v1 = iv;
        } else {
            keep = face != face0;
            u0 = min(u0, iu), u1 = max(u1, iu), v0 = min(v0, iv), v1 = max(v1, iv);
        }
    }
    --u0, ++u1, --v0, ++v1;
    keep = keep || u0 < 0 || v0 < 0 || u1 >= PRS_COVER_CELLS || v1 >= PRS_COVER_CELLS || u1 - u0 >= 10 || v1 - v0 >= 10;
    if (!keep) {
        const uint8_t *c = cover + (size_t)face0 * PRS_COVER_CELLS * PRS_COVER_CELLS;
        for (int v = v0; v <= v1; ++v)
            for (int u = u0; u <= u1; ++u) keep = keep || c[(size_t)v * PRS_COVER_CELLS + u] != 0;
    }
    tile_keep[tile] = keep ? 1 : 0;
}

template <typename T>
__global__ void area_lonlats_masked_kernel(const __grid_constant__ prs_proj_params P, const __grid_constant__ prs_area_grid g,
                                           const uint8_t *__restrict__ tile_keep, int tiles_x, T *lon_out, T *lat_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t W = g.width;
    if (i >= W * g.height) return;
    const int64_t r = i / W, c = i - r * W;
    T lon = Ar<T>::nan(), lat = Ar<T>::nan();
    if (tile_keep[(r >> 4) * tiles_x + (c >> 4)]) area_pixel_lonlat<T>(P, g, r, c, lon, lat);
    lon_out[i] = lon;
    lat_out[i] = lat;
}

template <typename T>
__global__ void valid_mask_kernel(const T *__restrict__ lon, const T *__restrict__ lat, int64_t n, const BoxT<T> box,
                                  const uint8_t *__restrict__ premask, uint8_t *__restrict__ valid)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ok = valid_predicate<T>(lon[i], lat[i], box, premask && premask[i]);
    valid[i] = ok ? 1 : 0;
}

__global__ void gather_nearest_kernel(const uint32_t *__restrict__ index, int64_t n_out, uint32_t n_valid,
                                      const uint32_t *__restrict__ rank_to_source, const void *data, int64_t row_bytes,
                                      int vec, const void *fill_row, void *out)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_out) return;
    uint32_t r = index[p];
    const void *src;
    if (r >= n_valid) {
        src = fill_row;
    } else {
        uint32_t s = rank_to_source ? __ldg(rank_to_source + r) : r;
        src = (const char *)data + (size_t)s * row_bytes;
    }
    copy_row((char *)out + (size_t)p * row_bytes, src, (int)row_bytes, vec);
}

struct WeightedParams {
    const uint32_t *index;
    const void *dist;
    int64_t n_out;
    int k;
    uint32_t n_valid;
    const uint32_t *rank_to_source;
    const void *data;
    int channels;
    const void *weights;  // [n_planes][n_out*k] or NULL
    const int32_t *plane_of_channel;  // device
    const double *sigma2;             // device or NULL
    double fill;
    void *out, *stddev_out, *count_out;
    uint32_t first_valid_source;
};

// get_sample_from_neighbour_info('custom') on precomputed neighbour info, runtime k (kd_tree.py:741-859).
// Tw: dtype of distances / explicit weights; Td: data dtype; Ta: accumulation and output dtype.
// One thread per (target point, group of four channels): the threads of a point sit in neighbouring lanes, so a
// neighbour's row is read as one contiguous segment (16 bytes per lane), index and weight loads are broadcasts, and the
// output row is written contiguously.  Per channel the neighbours are accumulated in order i = 0..k-1 like the
// reference's Python loop; four neighbours are fetched at a time so that their loads are in flight together.
template <typename Tw, typename Td, typename Ta>
__global__ void __launch_bounds__(256, 2) gather_weighted_kernel(const WeightedParams P)
{
    const int k = P.k, C = P.channels;
    const int groups = (C + 3) >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p = t / groups;
    if (p >= P.n_out) return;
    const int c0 = (int)(t - p * groups) * 4;
    const int nc = min(4, C - c0);
    const uint32_t *idx = P.index + (size_t)p * k;
    const bool want_uncert = P.stddev_out != nullptr;
    const Td *data = (const Td *)P.data;
    const bool vec_ok = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(data) & 15) == 0);
    const Tw *dist = P.weights ? nullptr : (const Tw *)P.dist + (size_t)p * k;
    const Tw *wplane[4] = {nullptr, nullptr, nullptr, nullptr};
    Tw s2c[4] = {1, 1, 1, 1};
    bool same_w = true;  // the four channels share one weight (same plane / same sigma): evaluate it once
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + min(j, nc - 1);
        if (P.weights)
            wplane[j] = (const Tw *)P.weights + ((size_t)P.plane_of_channel[c] * P.n_out + p) * k;
        else
            s2c[j] = (Tw)P.sigma2[c];
        same_w = same_w && wplane[j] == wplane[0] && s2c[j] == s2c[0];
    }
    Ta res[4] = {0, 0, 0, 0}, norm[4] = {0, 0, 0, 0};
    int count = 0;
    for (int pass = 0; pass < (want_uncert ? 2 : 1); ++pass) {
        Ta nsq[4] = {0, 0, 0, 0}, sd[4] = {0, 0, 0, 0};
        for (int i0 = 0; i0 < k; i0 += 4) {
            uint32_t r[4];
            bool missing[4];
            Td v[4][4];
            Ta w0[4];   // weight of the group's first channel (of all four when they share it)
            Tw d2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = i0 + u < k ? idx[i0 + u] : 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                missing[u] = r[u] >= P.n_valid;
                const uint32_t s = missing[u] ? P.first_valid_source : (P.rank_to_source ? __ldg(P.rank_to_source + r[u]) : r[u]);
                d2[u] = 0;
                w0[u] = 0;
                if (i0 + u < k) {
                    load_channels<Td>(data + (size_t)s * C + c0, nc, vec_ok, v[u]);
                    if (!P.weights) {
                        const Tw d = missing[u] ? (Tw)1 : dist[i0 + u];
                        d2[u] = Ar<Tw>::mul(d, d);
                    }
                    w0[u] = P.weights ? (Ta)wplane[0][i0 + u] : (Ta)Ar<Tw>::exp(Ar<Tw>::div(-d2[u], s2c[0]));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u >= k) continue;
                if (pass == 0 && !missing[u]) ++count;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    Ta w = w0[u];
                    if (j > 0 && !same_w)
                        w = P.weights ? (Ta)wplane[j][i0 + u] : (Ta)Ar<Tw>::exp(Ar<Tw>::div(-d2[u], s2c[j]));
                    const Ta wt = missing[u] ? Ar<Ta>::mul((Ta)0, w) : w;
                    if (pass == 0) {
                        res[j] = Ar<Ta>::add(res[j], Ar<Ta>::mul(wt, (Ta)v[u][j]));
                        norm[j] = Ar<Ta>::add(norm[j], wt);
                    } else {
                        nsq[j] = Ar<Ta>::add(nsq[j], Ar<Ta>::mul(wt, wt));
                        const Ta val = missing[u] ? Ar<Ta>::mul((Ta)0, (Ta)v[u][j]) : (Ta)v[u][j];
                        const Ta dv = Ar<Ta>::sub(val, res[j]);
                        sd[j] = Ar<Ta>::add(sd[j], Ar<Ta>::mul(wt, Ar<Ta>::mul(dv, dv)));
                    }
                }
            }
        }
        if (pass == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // norm > 0 is False for NaN as well, like numpy's boolean mask (kd_tree.py:807-809)
                res[j] = (norm[j] > 0) ? Ar<Ta>::div(res[j], norm[j]) : (Ta)P.fill;
            }
            Ta *o = (Ta *)P.out + (size_t)p * C + c0;
            if (nc == 4 && sizeof(Ta) == 8 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                reinterpret_cast<double2 *>(o)[0] = make_double2((double)res[0], (double)res[1]);
                reinterpret_cast<double2 *>(o)[1] = make_double2((double)res[2], (double)res[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nc) o[j] = res[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nc) {
                    const Ta sdv =
                        count > 1 ? Ar<Ta>::sqrt(Ar<Ta>::mul(
                                        Ar<Ta>::div(norm[j], Ar<Ta>::sub(Ar<Ta>::mul(norm[j], norm[j]), nsq[j])), sd[j]))
                                  : Ar<Ta>::nan();
                    ((Ta *)P.stddev_out)[(size_t)p * C + c0 + j] = sdv;
                    ((Ta *)P.count_out)[(size_t)p * C + c0 + j] = (Ta)count;
                }
            }
        }
    }
}

// the distances a weight function is given (kd_tree.py:751-768): float64, 1 where the neighbour is missing
template <typename T>
__global__ void __launch_bounds__(256) weight_distances_kernel(const uint32_t *__restrict__ index, const T *__restrict__ dist,
                                                               int64_t n, uint32_t n_valid, double *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __ldcs(index + i) >= n_valid ? 1.0 : (double)__ldcs(dist + i);
}

// Row compaction / scatter: one thread per VEC-byte piece of a row, consecutive threads on consecutive pieces, so a
// warp reads and writes whole lines whatever the row length (a thread per row touched 32 lines per instruction: the
// scatter of C4's 128-byte rows ran at 2.1 TB/s).
template <int VEC> struct Piece;
template <> struct Piece<16> { typedef uint4 type; };
template <> struct Piece<8> { typedef uint2 type; };
template <> struct Piece<4> { typedef uint32_t type; };
template <> struct Piece<2> { typedef uint16_t type; };
template <> struct Piece<1> { typedef uint8_t type; };

template <int VEC>
__global__ void __launch_bounds__(256) compact_rows_kernel(const uint8_t *__restrict__ mask, const uint32_t *__restrict__ pos,
                                                           int64_t n, int pieces, const void *in, void *out)
{
    typedef typename Piece<VEC>::type V;
    const int64_t total = n * pieces;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / pieces;
        const int c = (int)(t - row * pieces);
        if (mask[row]) ((V *)out)[(int64_t)pos[row] * pieces + c] = __ldcs((const V *)in + t);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) scatter_rows_kernel(const uint8_t *__restrict__ mask, const uint32_t *__restrict__ pos,
                                                           int64_t n, int pieces, const void *in, const void *fill_row,
                                                           void *out)
{
    typedef typename Piece<VEC>::type V;
    const int64_t total = n * pieces;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / pieces;
        const int c = (int)(t - row * pieces);
        const V v = mask[row] ? __ldcs((const V *)in + (int64_t)pos[row] * pieces + c) : ((const V *)fill_row)[c];
        __stcs((V *)out + t, v);
    }
}

template <typename F> int by_vec(int vec, F &&f)
{
    switch (vec) {
    case 16: return f(std::integral_constant<int, 16>());
    case 8: return f(std::integral_constant<int, 8>());
    case 4: return f(std::integral_constant<int, 4>());
    case 2: return f(std::integral_constant<int, 2>());
    default: return f(std::integral_constant<int, 1>());
    }
}

// grid of a grid-stride kernel over `total` work items: four per thread, at most 32 blocks of 256 threads per SM
static int stride_blocks(int64_t total)
{
    const int64_t want = (total + 1023) / 1024;
    return (int)(want < 1 ? 1 : (want < 148 * 32 ? want : 148 * 32));
}

__global__ void rank_to_source_kernel(const uint8_t *__restrict__ valid, const uint32_t *__restrict__ pos, int64_t n,
                                      uint32_t *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !valid[i]) return;
    out[pos[i]] = (uint32_t)i;
}

int pick_vec(int64_t row_bytes, const void *a, const void *b, const void *c)
{
    uintptr_t bits = (uintptr_t)row_bytes | (uintptr_t)a | (uintptr_t)b | (uintptr_t)c;
    if ((bits & 15) == 0) return 16;
    if ((bits & 7) == 0) return 8;
    if ((bits & 3) == 0) return 4;
    if ((bits & 1) == 0) return 2;
    return 1;
}

// ------------------------------------------------------------------------------------------ query launch
struct TargetTables {
    void *col_tab = nullptr, *row_tab = nullptr;
    uint8_t *col_ok = nullptr, *row_ok = nullptr;
};

}  // namespace

namespace prs {
// launch_search<T, K, EPI> / launch_classify<T> are instantiated in prs_query_inst.cu / prs_classify_inst.cu, never here
#define PRS_EXTERN_SEARCH(T, K, EPI)                                                                                   \
    extern template int launch_search<T, K, EPI>(const QueryParams &, const IndexRef<T> &, int64_t, const uint32_t *, \
                                                 const uint32_t *, cudaStream_t);
#define PRS_EXTERN_T(T)                                                                                              \
    extern template int launch_classify<T>(const QueryParams &, const IndexRef<T> &, int64_t, uint32_t *, uint32_t *, \
                                           cudaStream_t);                                                            \
    extern template int launch_coverage_t<T>(const QueryParams &, int, uint8_t *, cudaStream_t);                     \
    PRS_EXTERN_SEARCH(T, 1, PRS_EPI_INFO)                                                                            \
    PRS_EXTERN_SEARCH(T, 4, PRS_EPI_INFO)                                                                            \
    PRS_EXTERN_SEARCH(T, 8, PRS_EPI_INFO)                                                                            \
    PRS_EXTERN_SEARCH(T, 16, PRS_EPI_INFO)                                                                           \
    PRS_EXTERN_SEARCH(T, 32, PRS_EPI_INFO)                                                                           \
    PRS_EXTERN_SEARCH(T, 1, PRS_EPI_NEAREST)                                                                         \
    PRS_EXTERN_SEARCH(T, 4, PRS_EPI_GAUSS_FF)                                                                        \
    PRS_EXTERN_SEARCH(T, 8, PRS_EPI_GAUSS_FF)                                                                        \
    PRS_EXTERN_SEARCH(T, 16, PRS_EPI_GAUSS_FF)                                                                       \
    PRS_EXTERN_SEARCH(T, 32, PRS_EPI_GAUSS_FF)                                                                       \
    PRS_EXTERN_SEARCH(T, 4, PRS_EPI_GAUSS_DF)                                                                        \
    PRS_EXTERN_SEARCH(T, 8, PRS_EPI_GAUSS_DF)                                                                        \
    PRS_EXTERN_SEARCH(T, 16, PRS_EPI_GAUSS_DF)                                                                       \
    PRS_EXTERN_SEARCH(T, 32, PRS_EPI_GAUSS_DF)                                                                       \
    PRS_EXTERN_SEARCH(T, 4, PRS_EPI_GAUSS_DD)                                                                        \
    PRS_EXTERN_SEARCH(T, 8, PRS_EPI_GAUSS_DD)                                                                        \
    PRS_EXTERN_SEARCH(T, 16, PRS_EPI_GAUSS_DD)                                                                       \
    PRS_EXTERN_SEARCH(T, 32, PRS_EPI_GAUSS_DD)
PRS_EXTERN_T(float)
PRS_EXTERN_T(double)

// Dispatch on (K, epilogue).  The top-k list width K is the smallest instantiated one >= k.
template <typename T>
int launch_search_k(const QueryParams &P, const IndexRef<T> &ix, int64_t tiles, const uint32_t *live,
                    const uint32_t *n_live, cudaStream_t st)
{
    const int k = P.k;
    if (k > 32) return fail(PRS_ENOSUP, "neighbours > 32 is not supported");
#define PRS_GO(K, EPI) return launch_search<T, K, EPI>(P, ix, tiles, live, n_live, st)
    switch (P.mode) {
    case PRS_EPI_NEAREST:
        PRS_GO(1, PRS_EPI_NEAREST);
    case PRS_EPI_GAUSS:
#define PRS_GO_K(EPI)             \
    if (k <= 4) PRS_GO(4, EPI);   \
    if (k <= 8) PRS_GO(8, EPI);   \
    if (k <= 16) PRS_GO(16, EPI); \
    PRS_GO(32, EPI)
        if (!P.acc_f64) {
            PRS_GO_K(PRS_EPI_GAUSS_FF);
        } else if (!P.data_f64) {
            PRS_GO_K(PRS_EPI_GAUSS_DF);
        } else {
            PRS_GO_K(PRS_EPI_GAUSS_DD);
        }
#undef PRS_GO_K
    default:
        if (k <= 1) PRS_GO(1, PRS_EPI_INFO);
        if (k <= 4) PRS_GO(4, PRS_EPI_INFO);
        if (k <= 8) PRS_GO(8, PRS_EPI_INFO);
        if (k <= 16) PRS_GO(16, PRS_EPI_INFO);
        PRS_GO(32, PRS_EPI_INFO);
    }
#undef PRS_GO
}
}  // namespace prs

namespace {

// Fill the target part of the kernel parameters; separable projections get their per-row / per-column tables.
template <typename T> int setup_target(QueryParams &P, const prs_target *tg, int &tk, TargetTables &tabs, cudaStream_t st)
{
    tk = PRS_TK_LONLAT;
    P.n_t = tg->n;
    P.valid_extra = tg->valid_extra;
    if (tg->kind == PRS_TARGET_AREA) {
        if (tg->n != tg->nrows * tg->grid.width) return fail(PRS_EINVAL, "target.n != nrows * width");
        P.proj = tg->proj;
        P.grid = tg->grid;
        P.row0 = tg->row0;
        P.row_block = tg->row_block;
        P.row_stride = tg->row_stride;
        if (tg->nrows > 0x7FFFFFF0LL || tg->grid.width > 0x7FFFFFF0LL) return fail(PRS_ENOSUP, "area too large");
        P.nrows = (int)tg->nrows;
        P.tiles_x = (int)((tg->grid.width + 31) / 32);
        bool separable = tg->proj.kind == PRS_PROJ_LONGLAT || tg->proj.kind == PRS_PROJ_EQC || tg->proj.kind == PRS_PROJ_MERC;
        if (separable && tg->n > 0) {
            tk = PRS_TK_AREA_SEP;
            int64_t W = tg->grid.width, H = tg->nrows;
            PRS_CK(cudaMallocAsync(&tabs.col_tab, sizeof(ColTab<T>) * (size_t)W, st));
            PRS_CK(cudaMallocAsync(&tabs.row_tab, sizeof(RowTab<T>) * (size_t)H, st));
            PRS_CK(cudaMallocAsync((void **)&tabs.col_ok, (size_t)W, st));
            PRS_CK(cudaMallocAsync((void **)&tabs.row_ok, (size_t)H, st));
            area_tables_kernel<T><<<blocks_for(W + H), TPB, 0, st>>>(tg->proj, tg->grid, tg->row0, tg->row_block,
                                                                      tg->row_stride, H, (ColTab<T> *)tabs.col_tab,
                                                                      tabs.col_ok, (RowTab<T> *)tabs.row_tab, tabs.row_ok);
            PRS_CKL();
            P.col_tab = tabs.col_tab;
            P.row_tab = tabs.row_tab;
            P.col_ok = tabs.col_ok;
            P.row_ok = tabs.row_ok;
        } else {
            tk = PRS_TK_AREA_GEN;
        }
    } else {
        P.tlon = tg->lon;
        P.tlat = tg->lat;
    }
    P.tk = tk;
    return PRS_OK;
}

int free_tables(TargetTables &tabs, cudaStream_t st)
{
    if (tabs.col_tab) {
        PRS_CK(cudaFreeAsync(tabs.col_tab, st));
        PRS_CK(cudaFreeAsync(tabs.row_tab, st));
        PRS_CK(cudaFreeAsync(tabs.col_ok, st));
        PRS_CK(cudaFreeAsync(tabs.row_ok, st));
        tabs.col_tab = nullptr;
    }
    return PRS_OK;
}

// Bytes of the live-tile list at the head of a staged call's scratch buffer (the target coordinates follow it).
inline size_t live_list_bytes(int64_t tiles) { return ((size_t)(tiles + 1) * sizeof(uint32_t) + 255) & ~(size_t)255; }

template <typename T> int launch_query_t(QueryParams &P, const prs_index *index, const prs_target *tg, cudaStream_t st)
{
    IndexRef<T> ix;
    ix.recs = (const Rec<T> *)index->recs;
    ix.cell_start = index->cell_start;
    ix.coarse_start = index->coarse_start;
    ix.rank = index->rank;
    ix.meta = index->meta_dev;
    int tk = PRS_TK_LONLAT;
    TargetTables tabs;
    int rc = setup_target<T>(P, tg, tk, tabs, st);
    if (rc != PRS_OK) return rc;
    if (P.k > 32) {
        free_tables(tabs, st);
        return fail(PRS_ENOSUP, "neighbours > 32 is not supported");
    }
    if (P.n_t <= 0) return free_tables(tabs, st);
    const int64_t tiles = tile_count(P);
    if (tiles > 0x7FFFFFF0LL) return fail(PRS_ENOSUP, "target too large for one launch");
    // work buffers: [0] = number of live tiles, [1..] = their ids; for targets whose coordinates are expensive
    // (per-pixel inverse projection, explicit lon/lat) the (x, y, z, valid) hand-over from classification to search
    const size_t xyz_bytes = tk == PRS_TK_AREA_SEP ? 0 : 4 * sizeof(T) * (size_t)P.n_t;
    unsigned char *buf = (unsigned char *)P.scratch;
    if (P.stage == 0) PRS_CK(cudaMallocAsync((void **)&buf, live_list_bytes(tiles) + xyz_bytes, st));
    if (!buf) return fail(PRS_EINVAL, "staged call without scratch");
    uint32_t *n_live = (uint32_t *)buf, *live = n_live + 1;
    if (xyz_bytes) P.target_xyz = buf + live_list_bytes(tiles);
    if (P.stage != 2) rc = launch_classify<T>(P, ix, tiles, live, n_live, st);
    if (rc == PRS_OK && P.stage != 1) rc = launch_search_k<T>(P, ix, tiles, live, n_live, st);
    if (P.stage == 0) PRS_CK(cudaFreeAsync(buf, st));
    int rc2 = free_tables(tabs, st);
    return rc != PRS_OK ? rc : rc2;
}

int launch_query(QueryParams &P, const prs_index *index, const prs_target *tg, cudaStream_t st)
{
    if (!index || !tg) return fail(PRS_EINVAL, "null index/target");
    if (tg->dtype != index->tree_dtype && tg->kind == PRS_TARGET_LONLAT)
        return fail(PRS_EINVAL, "target lon/lat dtype must equal the tree dtype (kd_tree.py:536-542)");
    if (P.k < 1) return fail(PRS_EINVAL, "neighbours must be >= 1");
    const int rc = index->tree_dtype == PRS_F32 ? launch_query_t<float>(P, index, tg, st)
                                                : launch_query_t<double>(P, index, tg, st);
    index_note_use(index, st);
    return rc;
}

int read_flags(unsigned int *flags_dev, int64_t *flags_host, cudaStream_t st)
{
    unsigned int h[2] = {0, 0};
    PRS_CK(cudaMemcpyAsync(h, flags_dev, sizeof(h), cudaMemcpyDeviceToHost, st));
    PRS_CK(cudaStreamSynchronize(st));
    PRS_CK(cudaFreeAsync(flags_dev, st));
    flags_host[0] = h[0];
    flags_host[1] = h[1];
    return PRS_OK;
}

}  // namespace

// ============================================================================================== C-ABI
extern "C" {

int prs_abi_version(void) { return PRS_ABI_VERSION; }
int prs_release_cached_memory(void) { return prs::release_build_workspace(); }
const char *prs_last_error(void) { return err_slot().c_str(); }
int64_t prs_launch_count(void) { return (int64_t)launch_counter(); }

int prs_lonlat_to_xyz(const void *lon, const void *lat, int64_t n, int dtype, void *xyz_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return PRS_OK;
    if (dtype == PRS_F32)
        lonlat_to_xyz_kernel<float><<<blocks_for(n), TPB, 0, st>>>((const float *)lon, (const float *)lat, n, (float *)xyz_out);
    else if (dtype == PRS_F64)
        lonlat_to_xyz_kernel<double><<<blocks_for(n), TPB, 0, st>>>((const double *)lon, (const double *)lat, n, (double *)xyz_out);
    else
        return fail(PRS_EINVAL, "dtype must be PRS_F32 or PRS_F64");
    PRS_CKL();
    return PRS_OK;
}

int prs_area_lonlats(const prs_proj_params *proj, const prs_area_grid *grid, int dtype, int64_t row0, int64_t row_step,
                     int64_t nrows, int64_t col0, int64_t col_step, int64_t ncols, void *lon_out, void *lat_out,
                     void *xyz_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!proj || !grid) return fail(PRS_EINVAL, "null proj/grid");
    int64_t n = nrows * ncols;
    if (n <= 0) return PRS_OK;
    if (dtype == PRS_F32)
        area_lonlats_kernel<float><<<blocks_for(n), TPB, 0, st>>>(*proj, *grid, row0, row_step, nrows, col0, col_step, ncols,
                                                                   (float *)lon_out, (float *)lat_out, (float *)xyz_out);
    else if (dtype == PRS_F64)
        area_lonlats_kernel<double><<<blocks_for(n), TPB, 0, st>>>(*proj, *grid, row0, row_step, nrows, col0, col_step, ncols,
                                                                    (double *)lon_out, (double *)lat_out, (double *)xyz_out);
    else
        return fail(PRS_EINVAL, "dtype must be PRS_F32 or PRS_F64");
    PRS_CKL();
    return PRS_OK;
}

int prs_area_lonlats_covered(const prs_proj_params *proj, const prs_area_grid *grid, int dtype, const uint8_t *cover,
                             void *lon_out, void *lat_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!proj || !grid || !cover || !lon_out || !lat_out) return fail(PRS_EINVAL, "null argument");
    if (grid->width <= 0 || grid->height <= 0) return PRS_OK;
    if (dtype != PRS_F32 && dtype != PRS_F64) return fail(PRS_EINVAL, "dtype must be PRS_F32 or PRS_F64");
    const int64_t tx = (grid->width + 15) / 16, ty = (grid->height + 15) / 16;
    if (tx * ty > 0x7fffffffLL) return fail(PRS_ENOSUP, "area too large");
    uint8_t *tile_keep = nullptr;
    PRS_CK(cudaMallocAsync((void **)&tile_keep, (size_t)(tx * ty), st));
    area_tile_keep_kernel<<<(unsigned)((tx * ty + 127) / 128), 128, 0, st>>>(*proj, *grid, cover, (int)tx, (int)ty, tile_keep);
    PRS_CKL();
    const int64_t n = grid->width * grid->height;
    if (dtype == PRS_F32)
        area_lonlats_masked_kernel<float><<<blocks_for(n), TPB, 0, st>>>(*proj, *grid, tile_keep, (int)tx, (float *)lon_out,
                                                                          (float *)lat_out);
    else
        area_lonlats_masked_kernel<double><<<blocks_for(n), TPB, 0, st>>>(*proj, *grid, tile_keep, (int)tx, (double *)lon_out,
                                                                           (double *)lat_out);
    PRS_CKL();
    PRS_CK(cudaFreeAsync(tile_keep, st));
    return PRS_OK;
}

int prs_valid_mask(const void *lon, const void *lat, int64_t n, int dtype, const prs_reduce_box *box, const uint8_t *premask,
                   uint8_t *valid_out, int64_t *n_valid_host, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    prs_reduce_box b;
    memset(&b, 0, sizeof(b));
    if (box) b = *box;
    if (n > 0) {
        if (dtype == PRS_F32)
        {
            BoxT<float> bf;
            make_box(b, bf);
            valid_mask_kernel<float><<<blocks_for(n), TPB, 0, st>>>((const float *)lon, (const float *)lat, n, bf, premask, valid_out);
        }
        else if (dtype == PRS_F64)
        {
            BoxT<double> bd;
            make_box(b, bd);
            valid_mask_kernel<double><<<blocks_for(n), TPB, 0, st>>>((const double *)lon, (const double *)lat, n, bd, premask, valid_out);
        }
        else
            return fail(PRS_EINVAL, "dtype must be PRS_F32 or PRS_F64");
        PRS_CKL();
    }
    if (n_valid_host) {
        *n_valid_host = 0;
        if (n > 0) {
            uint32_t *tmp = nullptr, *tot = nullptr;
            PRS_CK(cudaMallocAsync((void **)&tmp, sizeof(uint32_t) * (size_t)n, st));
            PRS_CK(cudaMallocAsync((void **)&tot, sizeof(uint32_t), st));
            int rc = scan_u32<uint8_t>(valid_out, tmp, n, 0, tot, st);
            if (rc != PRS_OK) return rc;
            uint32_t h = 0;
            PRS_CK(cudaMemcpyAsync(&h, tot, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            PRS_CK(cudaStreamSynchronize(st));
            PRS_CK(cudaFreeAsync(tmp, st));
            PRS_CK(cudaFreeAsync(tot, st));
            *n_valid_host = h;
        }
    }
    return PRS_OK;
}

int prs_index_build(const void *lon, const void *lat, int64_t n, int dtype, uint8_t *valid, const prs_reduce_box *box,
                    const uint8_t *premask, const uint8_t *exclude, int tree_dtype, int k_hint, double radius_hint,
                    const uint8_t *cover, int flags, prs_index **index_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!index_out) return fail(PRS_EINVAL, "index_out is null");
    if (n <= 0 || n >= 0xFFFFFFFFLL) return fail(PRS_EINVAL, "source size must be in [1, 2^32 - 2]");
    if (dtype != tree_dtype && !(dtype == PRS_F32 && tree_dtype == PRS_F64))
        return fail(PRS_ENOSUP, "lon/lat dtype must equal the tree dtype, or be float32 for a float64 tree");
    if (!valid && !(flags & PRS_BUILD_COMPUTE_VALID))
        return fail(PRS_EINVAL, "valid mask is required (prs_valid_mask) unless PRS_BUILD_COMPUTE_VALID is set");
    if ((flags & PRS_BUILD_RANKS) && !valid) return fail(PRS_EINVAL, "PRS_BUILD_RANKS needs the valid mask buffer");
    prs_index *ix = new prs_index();
    memset(ix, 0, sizeof(*ix));
    ix->uses = new prs_index_uses();
    ix->tree_dtype = tree_dtype;
    ix->n_source = n;
    int rc = PRS_OK;
    if (cudaEventCreateWithFlags(&ix->planned, cudaEventDisableTiming) != cudaSuccess)
        rc = fail(PRS_ECUDA, "cudaEventCreate failed");
    else if (tree_dtype == PRS_F32)
        rc = index_build_t<float, float>((const float *)lon, (const float *)lat, n, valid, box, premask, exclude, cover,
                                         flags, k_hint, radius_hint, ix, st);
    else if (tree_dtype == PRS_F64 && dtype == PRS_F64)
        rc = index_build_t<double, double>((const double *)lon, (const double *)lat, n, valid, box, premask, exclude,
                                           cover, flags, k_hint, radius_hint, ix, st);
    else if (tree_dtype == PRS_F64)
        rc = index_build_t<float, double>((const float *)lon, (const float *)lat, n, valid, box, premask, exclude,
                                          cover, flags, k_hint, radius_hint, ix, st);
    else
        rc = fail(PRS_EINVAL, "tree_dtype must be PRS_F32 or PRS_F64");
    if (rc != PRS_OK) {
        prs_index_destroy(ix);
        return rc;
    }
    *index_out = ix;
    return PRS_OK;
}

int prs_target_coverage(const prs_target *target, double radius, uint8_t *cover_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!target || !cover_out) return fail(PRS_EINVAL, "null argument");
    PRS_CK(cudaMemsetAsync(cover_out, 0, 6ull * PRS_COVER_CELLS * PRS_COVER_CELLS, st));
    QueryParams P;
    memset(&P, 0, sizeof(P));
    P.radius = radius;
    int tk = PRS_TK_LONLAT, rc;
    TargetTables tabs;
    if (target->dtype == PRS_F32) {
        rc = setup_target<float>(P, target, tk, tabs, st);
        if (rc == PRS_OK) rc = launch_coverage_t<float>(P, tk, cover_out, st);
    } else {
        rc = setup_target<double>(P, target, tk, tabs, st);
        if (rc == PRS_OK) rc = launch_coverage_t<double>(P, tk, cover_out, st);
    }
    int rc2 = free_tables(tabs, st);
    return rc != PRS_OK ? rc : rc2;
}

int prs_index_destroy(prs_index *ix)
{
    if (!ix) return PRS_OK;
    // stream-ordered: blocks return to the pool once the work already queued on the build stream has finished -
    // and the queries that other streams still have in flight (one event per such stream)
    if (ix->uses) {
        for (auto &u : ix->uses->events) {
            cudaStreamWaitEvent(ix->stream, u.second, 0);
            cudaEventDestroy(u.second);
        }
        delete ix->uses;
        ix->uses = nullptr;
    }
    if (ix->recs) cudaFreeAsync(ix->recs, ix->stream);
    if (ix->cell_alloc) cudaFreeAsync(ix->cell_alloc, ix->stream);
    if (ix->coarse_alloc) cudaFreeAsync(ix->coarse_alloc, ix->stream);
    if (ix->rank) cudaFreeAsync(ix->rank, ix->stream);
    if (ix->meta_dev) cudaFreeAsync(ix->meta_dev, ix->stream);
    if (ix->planned) cudaEventDestroy(ix->planned);
    delete ix;
    return PRS_OK;
}

int prs_index_info(const prs_index *ix, int64_t info[8])
{
    if (!ix || !info) return fail(PRS_EINVAL, "null argument");
    // the build never waits for the device: the numbers are fetched (a copy that only waits for the plan kernel)
    // the first time somebody asks
    int rc = index_fetch_meta(ix);
    if (rc != PRS_OK) return rc;
    info[0] = ix->n_source;
    info[1] = ix->meta.n_valid;
    info[2] = ix->meta.n_indexed;
    info[3] = ix->meta.G;
    info[4] = (int64_t)ix->bytes;
    info[5] = ix->tree_dtype;
    info[6] = ix->meta.Gc;
    info[7] = ix->meta.all_valid;
    return PRS_OK;
}

int prs_query(const prs_index *index, const prs_target *target, int k, double radius, uint32_t *idx_out, void *dist_out,
              uint8_t *valid_out, int64_t *flags_host, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    QueryParams P;
    memset(&P, 0, sizeof(P));
    P.radius = radius;
    P.k = k;
    P.mode = PRS_EPI_INFO;
    P.idx_out = idx_out;
    P.dist_out = dist_out;
    P.valid_out = valid_out;
    if (!idx_out) return fail(PRS_EINVAL, "idx_out is null");
    if (index && !index->rank) {  // without the rank table index_array is only right when every source point is valid
        int rc0 = index_fetch_meta(index);
        if (rc0 != PRS_OK) return rc0;
        if (!index->meta.all_valid && index->meta.n_valid > 0)
            return fail(PRS_EINVAL, "prs_query needs an index built with PRS_BUILD_RANKS (some source points are invalid)");
    }
    if (flags_host) {
        PRS_CK(cudaMallocAsync((void **)&P.flags, 2 * sizeof(unsigned int), st));
        PRS_CK(cudaMemsetAsync(P.flags, 0, 2 * sizeof(unsigned int), st));
    }
    int rc = launch_query(P, index, target, st);
    if (rc != PRS_OK) return rc;
    if (flags_host) return read_flags(P.flags, flags_host, st);
    return PRS_OK;
}

int prs_resample_nearest(const prs_index *index, const prs_target *target, double radius, const void *data,
                         int64_t row_bytes, const void *fill_row, void *out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!data || !fill_row || !out || row_bytes <= 0) return fail(PRS_EINVAL, "bad data/fill/out");
    QueryParams P;
    memset(&P, 0, sizeof(P));
    P.radius = radius;
    P.k = 1;
    P.mode = PRS_EPI_NEAREST;
    P.data = data;
    P.fill_row = fill_row;
    P.out = out;
    if (row_bytes > 0x7FFFFFF0LL) return fail(PRS_ENOSUP, "row too large");
    P.row_bytes = (int)row_bytes;
    P.vec = pick_vec(row_bytes, data, fill_row, out);
    return launch_query(P, index, target, st);
}

int64_t prs_resample_scratch_bytes(const prs_target *target)
{
    if (!target || target->kind != PRS_TARGET_AREA) return 0;
    const int64_t tiles = ((target->grid.width + 31) / 32) * ((target->nrows + PRS_TILE_ROWS - 1) / PRS_TILE_ROWS);
    const bool separable = target->proj.kind == PRS_PROJ_LONGLAT || target->proj.kind == PRS_PROJ_EQC ||
                           target->proj.kind == PRS_PROJ_MERC;
    const size_t coord = target->dtype == PRS_F64 ? sizeof(double) : sizeof(float);
    return (int64_t)(live_list_bytes(tiles) + (separable ? 0 : 4 * coord * (size_t)target->n));
}

static int resample_nearest_stage(int stage, const prs_index *index, const prs_target *target, double radius,
                                  const void *data, int64_t row_bytes, const void *fill_row, void *out, void *scratch,
                                  uint8_t *tile_row_live, cudaStream_t st)
{
    static_assert(PRS_TILE_ROWS == prs::QROWS, "PRS_TILE_ROWS of the header is the kernels' tile height");
    if (!target || target->kind != PRS_TARGET_AREA) return fail(PRS_ENOSUP, "staged resampling needs an area target");
    if (!fill_row || !out || !scratch || row_bytes <= 0 || row_bytes > 0x7FFFFFF0LL) return fail(PRS_EINVAL, "bad fill/out/scratch");
    if (stage == 2 && !data) return fail(PRS_EINVAL, "data is null");
    QueryParams P;
    memset(&P, 0, sizeof(P));
    P.radius = radius;
    P.k = 1;
    P.mode = PRS_EPI_NEAREST;
    P.data = data;
    P.fill_row = fill_row;
    P.out = out;
    P.row_bytes = (int)row_bytes;
    P.vec = pick_vec(row_bytes, data ? data : fill_row, fill_row, out);
    P.stage = stage;
    P.scratch = (uint32_t *)scratch;
    if (stage == 1 && tile_row_live) {
        P.tile_row_live = tile_row_live;
        PRS_CK(cudaMemsetAsync(tile_row_live, 0, (size_t)((target->nrows + PRS_TILE_ROWS - 1) / PRS_TILE_ROWS), st));
    }
    return launch_query(P, index, target, st);
}

int prs_resample_nearest_fill(const prs_index *index, const prs_target *target, double radius, int64_t row_bytes,
                              const void *fill_row, void *out, void *scratch, uint8_t *tile_row_live, void *stream)
{
    return resample_nearest_stage(1, index, target, radius, nullptr, row_bytes, fill_row, out, scratch, tile_row_live,
                                  (cudaStream_t)stream);
}

int prs_resample_nearest_search(const prs_index *index, const prs_target *target, double radius, const void *data,
                                int64_t row_bytes, const void *fill_row, void *out, void *scratch, void *stream)
{
    return resample_nearest_stage(2, index, target, radius, data, row_bytes, fill_row, out, scratch, nullptr,
                                  (cudaStream_t)stream);
}

int prs_resample_gauss(const prs_index *index, const prs_target *target, int k, double radius, const void *data,
                       int data_dtype, int channels, const double *sigmas_host, double fill, void *out, void *stddev_out,
                       void *count_out, int64_t *flags_host, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!index || !data || !out || channels < 1 || !sigmas_host) return fail(PRS_EINVAL, "bad argument");
    if ((stddev_out == nullptr) != (count_out == nullptr)) return fail(PRS_EINVAL, "stddev_out and count_out go together");
    QueryParams P;
    memset(&P, 0, sizeof(P));
    P.radius = radius;
    P.k = k;
    P.mode = PRS_EPI_GAUSS;
    P.data = data;
    P.out = out;
    P.channels = channels;
    P.data_f64 = data_dtype == PRS_F64;
    // numpy promotion in _resample_with_weights: multi-channel weights are float64 (np.ones(n) * w, kd_tree.py:783)
    P.acc_f64 = (channels > 1) || P.data_f64 || index->tree_dtype == PRS_F64;
    P.fill = fill;
    P.stddev_out = stddev_out;
    P.count_out = count_out;
    std::vector<double> s2(channels);
    bool uniform = true;
    for (int c = 0; c < channels; ++c) {
        s2[c] = sigmas_host[c] * sigmas_host[c];
        if (s2[c] != s2[0]) uniform = false;
    }
    P.uniform_sigma = uniform;
    double *s2_dev = nullptr;
    PRS_CK(cudaMallocAsync((void **)&s2_dev, sizeof(double) * channels, st));
    PRS_CK(cudaMemcpyAsync(s2_dev, s2.data(), sizeof(double) * channels, cudaMemcpyHostToDevice, st));
    PRS_CK(cudaStreamSynchronize(st));  // s2 is a stack-owned staging buffer
    P.sigma2 = s2_dev;
    if (flags_host) {
        PRS_CK(cudaMallocAsync((void **)&P.flags, 2 * sizeof(unsigned int), st));
        PRS_CK(cudaMemsetAsync(P.flags, 0, 2 * sizeof(unsigned int), st));
    }
    int rc = launch_query(P, index, target, st);
    PRS_CK(cudaFreeAsync(s2_dev, st));
    if (rc != PRS_OK) return rc;
    if (flags_host) return read_flags(P.flags, flags_host, st);
    return PRS_OK;
}

int prs_gather_nearest(const uint32_t *index, int64_t n_out, uint32_t n_valid, const uint32_t *rank_to_source,
                       const void *data, int64_t row_bytes, const void *fill_row, void *out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_out <= 0) return PRS_OK;
    if (!index || !data || !fill_row || !out || row_bytes <= 0) return fail(PRS_EINVAL, "bad argument");
    int vec = pick_vec(row_bytes, data, fill_row, out);
    gather_nearest_kernel<<<blocks_for(n_out), TPB, 0, st>>>(index, n_out, n_valid, rank_to_source, data, row_bytes, vec,
                                                            fill_row, out);
    PRS_CKL();
    return PRS_OK;
}

int prs_gather_weighted(const uint32_t *index, const void *dist, int64_t n_out, int k, uint32_t n_valid,
                        const uint32_t *rank_to_source, const void *data, int data_dtype, int channels, const void *weights,
                        int weight_dtype, int n_planes, const int32_t *plane_of_channel_host, const double *sigmas_host,
                        double fill, void *out, void *stddev_out, void *count_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_out <= 0) return PRS_OK;
    if (!index || !data || !out || channels < 1 || k < 1) return fail(PRS_EINVAL, "bad argument");
    if (!weights && !(sigmas_host && dist)) return fail(PRS_EINVAL, "need explicit weights or (dist, sigmas)");
    if ((stddev_out == nullptr) != (count_out == nullptr)) return fail(PRS_EINVAL, "stddev_out and count_out go together");
    WeightedParams P;
    memset(&P, 0, sizeof(P));
    P.index = index;
    P.dist = dist;
    P.n_out = n_out;
    P.k = k;
    P.n_valid = n_valid;
    P.rank_to_source = rank_to_source;
    P.data = data;
    P.channels = channels;
    P.weights = weights;
    P.fill = fill;
    P.out = out;
    P.stddev_out = stddev_out;
    P.count_out = count_out;
    // first valid source row (new_data[0]) for missing neighbours
    uint32_t fvs = 0;
    if (rank_to_source) {
        PRS_CK(cudaMemcpyAsync(&fvs, rank_to_source, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PRS_CK(cudaStreamSynchronize(st));
    }
    P.first_valid_source = fvs;
    int32_t *planes_dev = nullptr;
    double *s2_dev = nullptr;
    std::vector<int32_t> planes(channels, 0);
    std::vector<double> s2(channels, 1.0);
    if (weights) {
        for (int c = 0; c < channels; ++c) {
            planes[c] = plane_of_channel_host ? plane_of_channel_host[c] : 0;
            if (planes[c] < 0 || planes[c] >= n_planes) return fail(PRS_EINVAL, "plane_of_channel out of range");
        }
        PRS_CK(cudaMallocAsync((void **)&planes_dev, sizeof(int32_t) * channels, st));
        PRS_CK(cudaMemcpyAsync(planes_dev, planes.data(), sizeof(int32_t) * channels, cudaMemcpyHostToDevice, st));
    } else {
        for (int c = 0; c < channels; ++c) s2[c] = sigmas_host[c] * sigmas_host[c];
        PRS_CK(cudaMallocAsync((void **)&s2_dev, sizeof(double) * channels, st));
        PRS_CK(cudaMemcpyAsync(s2_dev, s2.data(), sizeof(double) * channels, cudaMemcpyHostToDevice, st));
    }
    PRS_CK(cudaStreamSynchronize(st));
    P.plane_of_channel = planes_dev;
    P.sigma2 = s2_dev;
    bool w64 = weight_dtype == PRS_F64, d64 = data_dtype == PRS_F64;
    bool acc64 = channels > 1 || w64 || d64;
    unsigned grid = blocks_for(n_out * ((channels + 3) / 4));
    if (!acc64)
        gather_weighted_kernel<float, float, float><<<grid, TPB, 0, st>>>(P);
    else if (w64 && d64)
        gather_weighted_kernel<double, double, double><<<grid, TPB, 0, st>>>(P);
    else if (w64)
        gather_weighted_kernel<double, float, double><<<grid, TPB, 0, st>>>(P);
    else if (d64)
        gather_weighted_kernel<float, double, double><<<grid, TPB, 0, st>>>(P);
    else
        gather_weighted_kernel<float, float, double><<<grid, TPB, 0, st>>>(P);
    PRS_CKL();
    if (planes_dev) PRS_CK(cudaFreeAsync(planes_dev, st));
    if (s2_dev) PRS_CK(cudaFreeAsync(s2_dev, st));
    return PRS_OK;
}

int prs_compact_rows(const uint8_t *mask, int64_t n, int64_t row_bytes, const void *in, void *out, int64_t *n_kept_host,
                     void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_kept_host) *n_kept_host = 0;
    if (n <= 0) return PRS_OK;
    uint32_t *pos = nullptr, *tot = nullptr;
    PRS_CK(cudaMallocAsync((void **)&pos, sizeof(uint32_t) * (size_t)n, st));
    PRS_CK(cudaMallocAsync((void **)&tot, sizeof(uint32_t), st));
    int rc = scan_u32<uint8_t>(mask, pos, n, 0, tot, st);
    if (rc != PRS_OK) return rc;
    if (in && out) {
        const int vec = pick_vec(row_bytes, in, out, nullptr);
        const int pieces = (int)(row_bytes / vec);
        by_vec(vec, [&](auto v) {
            compact_rows_kernel<decltype(v)::value><<<stride_blocks(n * pieces), 256, 0, st>>>(mask, pos, n, pieces, in, out);
            return 0;
        });
        PRS_CKL();
    }
    if (n_kept_host) {
        uint32_t h = 0;
        PRS_CK(cudaMemcpyAsync(&h, tot, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PRS_CK(cudaStreamSynchronize(st));
        *n_kept_host = h;
    }
    PRS_CK(cudaFreeAsync(pos, st));
    PRS_CK(cudaFreeAsync(tot, st));
    return PRS_OK;
}

int prs_weight_distances(const uint32_t *index, const void *dist, int dist_dtype, int64_t n, uint32_t n_valid, double *out,
                         void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return PRS_OK;
    if (!index || !dist || !out) return fail(PRS_EINVAL, "prs_weight_distances: null array");
    const int64_t want = (n + 1023) / 1024;  // four entries per thread
    const int blocks = (int)(want < 148 * 32 ? want : 148 * 32);
    if (dist_dtype == PRS_F32)
        weight_distances_kernel<float><<<blocks, 256, 0, st>>>(index, (const float *)dist, n, n_valid, out);
    else if (dist_dtype == PRS_F64)
        weight_distances_kernel<double><<<blocks, 256, 0, st>>>(index, (const double *)dist, n, n_valid, out);
    else
        return fail(PRS_EINVAL, "prs_weight_distances: distances are float32 or float64");
    PRS_CKL();
    return PRS_OK;
}

int prs_rank_to_source(const uint8_t *valid, int64_t n, uint32_t *out, int64_t *n_valid_host, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_valid_host) *n_valid_host = 0;
    if (n <= 0) return PRS_OK;
    uint32_t *pos = nullptr, *tot = nullptr;
    PRS_CK(cudaMallocAsync((void **)&pos, sizeof(uint32_t) * (size_t)n, st));
    PRS_CK(cudaMallocAsync((void **)&tot, sizeof(uint32_t), st));
    int rc = scan_u32<uint8_t>(valid, pos, n, 0, tot, st);
    if (rc != PRS_OK) return rc;
    if (out) {
        rank_to_source_kernel<<<blocks_for(n), TPB, 0, st>>>(valid, pos, n, out);
        PRS_CKL();
    }
    if (n_valid_host) {
        uint32_t h = 0;
        PRS_CK(cudaMemcpyAsync(&h, tot, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PRS_CK(cudaStreamSynchronize(st));
        *n_valid_host = h;
    }
    PRS_CK(cudaFreeAsync(pos, st));
    PRS_CK(cudaFreeAsync(tot, st));
    return PRS_OK;
}

int prs_scatter_rows(const uint8_t *mask, int64_t n, int64_t row_bytes, const void *in, const void *fill_row, void *out,
                     void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return PRS_OK;
    if (!mask || !in || !fill_row || !out || row_bytes <= 0) return fail(PRS_EINVAL, "bad argument");
    uint32_t *pos = nullptr;
    PRS_CK(cudaMallocAsync((void **)&pos, sizeof(uint32_t) * (size_t)n, st));
    int rc = scan_u32<uint8_t>(mask, pos, n, 0, nullptr, st);
    if (rc != PRS_OK) return rc;
    const int vec = pick_vec(row_bytes, in, out, fill_row);
    const int pieces = (int)(row_bytes / vec);
    by_vec(vec, [&](auto v) {
        scatter_rows_kernel<decltype(v)::value><<<stride_blocks(n * pieces), 256, 0, st>>>(mask, pos, n, pieces, in, fill_row,
                                                                                       out);
        return 0;
    });
    PRS_CKL();
    PRS_CK(cudaFreeAsync(pos, st));
    return PRS_OK;
}

/* ---- bilinear (pyresample/bilinear/_base.py) ------------------------------------------------------------- */
int prs_project_forward(const void *lon, const void *lat, int64_t n, int dtype, const prs_proj_params *proj, void *xy_out,
                        void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return PRS_OK;
    if (!lon || !lat || !proj || !xy_out) return fail(PRS_EINVAL, "null argument");
    if (dtype == PRS_F32)
        project_forward_kernel<float><<<blocks_for(n), TPB, 0, st>>>((const float *)lon, (const float *)lat, n, *proj, (double2 *)xy_out);
    else if (dtype == PRS_F64)
        project_forward_kernel<double><<<blocks_for(n), TPB, 0, st>>>((const double *)lon, (const double *)lat, n, *proj, (double2 *)xy_out);
    else
        return fail(PRS_EINVAL, "dtype must be PRS_F32 or PRS_F64");
    PRS_CKL();
    return PRS_OK;
}

int prs_bilinear_info(const uint32_t *index, int64_t n_t, int k, uint32_t n_valid, const uint32_t *rank_to_source,
                      const void *xy, const prs_area_grid *grid, double *t_out, double *s_out, uint32_t *idx4_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_t <= 0) return PRS_OK;
    if (!index || !xy || !grid || !t_out || !s_out || !idx4_out || k < 1) return fail(PRS_EINVAL, "bad argument");
    if (n_t != grid->width * grid->height) return fail(PRS_EINVAL, "n_t != width * height");
    BilInfoParams P;
    P.index = index;
    P.n_t = n_t;
    P.k = k;
    P.n_valid = n_valid;
    P.rank_to_source = rank_to_source;
    P.xy = (const double2 *)xy;
    P.grid = *grid;
    P.t_out = t_out;
    P.s_out = s_out;
    P.idx4_out = idx4_out;
    bilinear_info_kernel<<<blocks_for(n_t), TPB, 0, st>>>(P);
    PRS_CKL();
    return PRS_OK;
}

int prs_bilinear_sample(const uint32_t *idx4, const double *t, const double *s, int64_t n_out, const uint32_t *rank_to_source,
                        const void *data, int data_dtype, int channels, double fill, double *out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_out <= 0) return PRS_OK;
    if (!idx4 || !t || !s || !data || !out || channels < 1) return fail(PRS_EINVAL, "bad argument");
    BilSampleParams P;
    P.idx4 = idx4;
    P.t = t;
    P.s = s;
    P.n_out = n_out;
    P.rank_to_source = rank_to_source;
    P.data = data;
    P.channels = channels;
    P.fill = fill;
    P.out = out;
    const unsigned grid = blocks_for(n_out * channels);
    if (data_dtype == PRS_F32)
        bilinear_sample_kernel<float><<<grid, TPB, 0, st>>>(P);
    else if (data_dtype == PRS_F64)
        bilinear_sample_kernel<double><<<grid, TPB, 0, st>>>(P);
    else
        return fail(PRS_EINVAL, "data_dtype must be PRS_F32 or PRS_F64");
    PRS_CKL();
    return PRS_OK;
}

}  // extern "C"
