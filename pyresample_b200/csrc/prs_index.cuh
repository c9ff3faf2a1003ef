// Spatial index over the valid source points: cube-face cell grid built by a counting sort
// (one-digit radix sort on the cell key), replacing pykdtree.KDTree(input_coords) (kd_tree.py:464-489).
#pragma once
#include "prs_common.cuh"
#include "prs_scan.cuh"

struct prs_index {
    int tree_dtype;      // PRS_F32 / PRS_F64
    int64_t n_source;    // raveled source size
    int64_t n_valid;     // pykdtree's .n : number of valid source points (rank space size)
    int64_t n_indexed;   // valid and not excluded
    int G;               // fine cells per face side (multiple of COARSE)
    int Gc;              // coarse cells per face side
    int all_valid;       // rank == source index
    uint32_t first_valid_source;  // raveled index of rank 0 (kd_tree.py:754-759 gathers new_data[0] for missing)
    void *recs;          // Rec<T>[n_indexed], sorted by cell
    uint32_t *cell_start;    // [6*G*G + 1]
    uint32_t *coarse_start;  // [6*Gc*Gc + 1]
    uint32_t *rank;          // [n_source] rank of each source point (NULL when all_valid)
    size_t bytes;
};

namespace prs {

constexpr int COARSE = 16;       // fine cells per coarse cell side
constexpr int OCC_G = 128;       // resolution of the occupancy probe used to size the cells
constexpr uint32_t CELL_SKIP = 0xFFFFFFFFu;

template <typename T> __device__ __forceinline__ uint32_t cell_of(T x, T y, T z, int G)
{
    int face;
    float u, v;
    const float invR = (float)(1.0 / PRS_R);
    face_uv((float)x * invR, (float)y * invR, (float)z * invR, face, u, v);
    int iu = cell_coord(warp_uv(u), G), iv = cell_coord(warp_uv(v), G);
    return (uint32_t)((face * G + iv) * G + iu);
}

// Pass 1: ECEF of every valid point, written in rank order; occupancy probe at OCC_G resolution.
template <typename T>
__global__ void build_xyz_kernel(const T *__restrict__ lon, const T *__restrict__ lat, int64_t n,
                                 const uint8_t *__restrict__ valid, const uint8_t *__restrict__ exclude,
                                 const uint32_t *__restrict__ rank, Rec<T> *__restrict__ tmp,
                                 uint32_t *__restrict__ cellid, uint8_t *__restrict__ occ)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !valid[i]) return;
    uint32_t r = rank ? rank[i] : (uint32_t)i;
    T x, y, z;
    lonlat_to_xyz(lon[i], lat[i], x, y, z);
    store_rec(tmp + r, x, y, z, (uint32_t)i);
    bool skip = exclude && exclude[i];
    cellid[r] = skip ? CELL_SKIP : 0u;
    if (!skip) occ[cell_of(x, y, z, OCC_G)] = 1;
}

// Pass 2: cell key of every indexed point + histogram into table[cell + 1].
template <typename T>
__global__ void build_count_kernel(const Rec<T> *__restrict__ tmp, int64_t n_valid, int G,
                                   uint32_t *__restrict__ cellid, uint32_t *__restrict__ table)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_valid) return;
    if (cellid[j] == CELL_SKIP) return;
    Rec<T> r = load_rec(tmp + j);
    uint32_t c = cell_of(r.x, r.y, r.z, G);
    cellid[j] = c;
    atomicAdd(&table[c + 1], 1u);
}

// Pass 3: scatter.  table[c + 1] holds start(c) after the exclusive scan and ends as start(c + 1).
template <typename T>
__global__ void build_scatter_kernel(const Rec<T> *__restrict__ tmp, int64_t n_valid,
                                     const uint32_t *__restrict__ cellid, uint32_t *__restrict__ table,
                                     Rec<T> *__restrict__ recs)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_valid) return;
    uint32_t c = cellid[j];
    if (c == CELL_SKIP) return;
    uint32_t pos = atomicAdd(&table[c + 1], 1u);
    Rec<T> r = load_rec(tmp + j);
    store_rec(recs + pos, r.x, r.y, r.z, r.idx);
}

// Within a cell the scatter order depends on atomic arrival; sort each cell by source index so the
// layout (not just the result) is deterministic.  Cells hold a handful of points: insertion sort.
template <typename T>
__global__ void build_sort_cells_kernel(Rec<T> *__restrict__ recs, const uint32_t *__restrict__ cell_start,
                                        int64_t n_cells)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    uint32_t s = cell_start[c], e = cell_start[c + 1];
    for (uint32_t i = s + 1; i < e; ++i) {
        Rec<T> key = recs[i];
        uint32_t j = i;
        while (j > s && recs[j - 1].idx > key.idx) {
            recs[j] = recs[j - 1];
            --j;
        }
        recs[j] = key;
    }
}

// Coarse table: number of points in each COARSE x COARSE block of fine cells, from the fine row prefix sums.
__global__ void build_coarse_kernel(const uint32_t *__restrict__ cell_start, int G, int Gc,
                                    uint32_t *__restrict__ ctable)
{
    int64_t cc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t ncc = 6LL * Gc * Gc;
    if (cc >= ncc) return;
    int cu = (int)(cc % Gc);
    int cv = (int)((cc / Gc) % Gc);
    int face = (int)(cc / ((int64_t)Gc * Gc));
    uint32_t tot = 0;
    for (int r = 0; r < COARSE; ++r) {
        int64_t row = ((int64_t)face * G + (cv * COARSE + r)) * G;
        tot += cell_start[row + cu * COARSE + COARSE] - cell_start[row + cu * COARSE];
    }
    ctable[cc + 1] = tot;
}

template <typename T>
int index_build_t(const T *lon, const T *lat, int64_t n, const uint8_t *valid, const uint8_t *exclude, int k_hint,
                  double radius_hint, prs_index *ix, cudaStream_t st)
{
    (void)radius_hint;
    const int TPB = 256;
    // ranks of valid points
    uint32_t *rank = nullptr, *tot_dev = nullptr;
    PRS_CK(cudaMalloc((void **)&rank, sizeof(uint32_t) * (size_t)n));
    PRS_CK(cudaMallocAsync((void **)&tot_dev, sizeof(uint32_t) * 2, st));
    int rc = scan_u32<uint8_t>(valid, rank, n, 0, tot_dev, st);
    if (rc != PRS_OK) return rc;
    uint32_t n_valid_h = 0;
    PRS_CK(cudaMemcpyAsync(&n_valid_h, tot_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PRS_CK(cudaStreamSynchronize(st));
    ix->n_valid = n_valid_h;
    ix->all_valid = (int64_t)n_valid_h == n;
    ix->rank = rank;
    ix->bytes += sizeof(uint32_t) * (size_t)n;
    if (ix->all_valid) {
        PRS_CK(cudaFree(rank));
        ix->rank = rank = nullptr;
        ix->bytes -= sizeof(uint32_t) * (size_t)n;
    }
    if (n_valid_h == 0) {
        PRS_CK(cudaFreeAsync(tot_dev, st));
        ix->n_indexed = 0;
        ix->G = COARSE;
        ix->Gc = 1;
        return PRS_OK;
    }
    // pass 1
    Rec<T> *tmp = nullptr;
    uint32_t *cellid = nullptr;
    uint8_t *occ = nullptr;
    const size_t occ_n = 6ull * OCC_G * OCC_G;
    PRS_CK(cudaMallocAsync((void **)&tmp, sizeof(Rec<T>) * (size_t)n_valid_h, st));
    PRS_CK(cudaMallocAsync((void **)&cellid, sizeof(uint32_t) * (size_t)n_valid_h, st));
    PRS_CK(cudaMallocAsync((void **)&occ, occ_n, st));
    PRS_CK(cudaMemsetAsync(occ, 0, occ_n, st));
    build_xyz_kernel<T><<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, st>>>(lon, lat, n, valid, exclude, rank, tmp,
                                                                          cellid, occ);
    PRS_CKL();
    // first valid source index + occupancy -> cell resolution
    uint32_t *occ_sum = nullptr;
    PRS_CK(cudaMallocAsync((void **)&occ_sum, sizeof(uint32_t) * occ_n, st));
    rc = scan_u32<uint8_t>(occ, occ_sum, (int64_t)occ_n, 1, tot_dev + 1, st);
    if (rc != PRS_OK) return rc;
    uint32_t n_occ = 0;
    Rec<T> first;
    PRS_CK(cudaMemcpyAsync(&n_occ, tot_dev + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PRS_CK(cudaMemcpyAsync(&first, tmp, sizeof(Rec<T>), cudaMemcpyDeviceToHost, st));
    PRS_CK(cudaStreamSynchronize(st));
    PRS_CK(cudaFreeAsync(occ_sum, st));
    PRS_CK(cudaFreeAsync(occ, st));
    ix->first_valid_source = first.idx;
    double lambda = k_hint > 4 ? (double)k_hint : 4.0;  // target points per cell
    if (const char *env = getenv("PRS_CELL_OCCUPANCY")) {
        double v = atof(env);
        if (v > 0) lambda = v;
    }
    if (n_occ == 0) n_occ = 1;
    double g = (double)OCC_G * sqrt((double)n_valid_h / ((double)n_occ * lambda));
    int G = (int)ceil(g / COARSE) * COARSE;
    int Gmax = 8192;
    if (const char *env = getenv("PRS_MAX_CELLS_PER_SIDE")) {
        int v = atoi(env);
        if (v >= COARSE) Gmax = (v / COARSE) * COARSE;
    }
    if (G < COARSE) G = COARSE;
    if (G > Gmax) G = Gmax;
    ix->G = G;
    ix->Gc = G / COARSE;
    // pass 2 + scan + pass 3
    const int64_t n_cells = 6LL * G * G;
    PRS_CK(cudaMalloc((void **)&ix->cell_start, sizeof(uint32_t) * (size_t)(n_cells + 1)));
    ix->bytes += sizeof(uint32_t) * (size_t)(n_cells + 1);
    PRS_CK(cudaMemsetAsync(ix->cell_start, 0, sizeof(uint32_t) * (size_t)(n_cells + 1), st));
    unsigned gv = (unsigned)(((int64_t)n_valid_h + TPB - 1) / TPB);
    build_count_kernel<T><<<gv, TPB, 0, st>>>(tmp, n_valid_h, G, cellid, ix->cell_start);
    PRS_CKL();
    rc = scan_u32<uint32_t>(ix->cell_start + 1, ix->cell_start + 1, n_cells, 0, tot_dev, st);
    if (rc != PRS_OK) return rc;
    uint32_t n_idx_h = 0;
    PRS_CK(cudaMemcpyAsync(&n_idx_h, tot_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PRS_CK(cudaStreamSynchronize(st));
    ix->n_indexed = n_idx_h;
    size_t rec_bytes = sizeof(Rec<T>) * (size_t)(n_idx_h > 0 ? n_idx_h : 1);
    PRS_CK(cudaMalloc((void **)&ix->recs, rec_bytes));
    ix->bytes += rec_bytes;
    build_scatter_kernel<T><<<gv, TPB, 0, st>>>(tmp, n_valid_h, cellid, ix->cell_start, (Rec<T> *)ix->recs);
    PRS_CKL();
    build_sort_cells_kernel<T><<<(unsigned)((n_cells + TPB - 1) / TPB), TPB, 0, st>>>((Rec<T> *)ix->recs,
                                                                                      ix->cell_start, n_cells);
    PRS_CKL();
    // coarse table
    const int64_t n_cc = 6LL * ix->Gc * ix->Gc;
    PRS_CK(cudaMalloc((void **)&ix->coarse_start, sizeof(uint32_t) * (size_t)(n_cc + 1)));
    ix->bytes += sizeof(uint32_t) * (size_t)(n_cc + 1);
    PRS_CK(cudaMemsetAsync(ix->coarse_start, 0, sizeof(uint32_t), st));
    build_coarse_kernel<<<(unsigned)((n_cc + TPB - 1) / TPB), TPB, 0, st>>>(ix->cell_start, G, ix->Gc,
                                                                          ix->coarse_start);
    PRS_CKL();
    rc = scan_u32<uint32_t>(ix->coarse_start + 1, ix->coarse_start + 1, n_cc, 1, nullptr, st);
    if (rc != PRS_OK) return rc;
    PRS_CK(cudaFreeAsync(tmp, st));
    PRS_CK(cudaFreeAsync(cellid, st));
    PRS_CK(cudaFreeAsync(tot_dev, st));
    return PRS_OK;
}

}  // namespace prs
