// Host-side launch of the query kernels for one (tree dtype, K, target kind).  launch_query_k is instantiated
// explicitly in prs_query_inst.cu - one translation unit per (T, K), and per target kind for the wide top-k lists,
// whose unrolled insertion networks take minutes to compile - so that the units build in parallel.
#pragma once
#include "prs_query.cuh"

namespace prs {

template <typename T, int K, int TK> int launch_query_k(const QueryParams &P, const IndexView<T> &ix, cudaStream_t st)
{
    if (P.n_t <= 0) return PRS_OK;
    int64_t tiles;
    if (TK == PRS_TK_LONLAT)
        tiles = (P.n_t + 32 * QROWS - 1) / (32 * QROWS);
    else
        tiles = (int64_t)P.tiles_x * ((P.nrows + QROWS - 1) / QROWS);
    if (tiles > 0x7FFFFFF0LL) return fail(PRS_ENOSUP, "target too large for one launch");
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        PRS_CK(cudaGetDevice(&dev));
        PRS_CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    uint32_t *buf = P.scratch;  // [0] = number of live tiles, [1..] = their ids
    if (P.stage == 0) PRS_CK(cudaMallocAsync((void **)&buf, sizeof(uint32_t) * (size_t)(tiles + 1), st));
    if (!buf) return fail(PRS_EINVAL, "staged call without scratch");
    uint32_t *n_live = buf, *live = buf + 1;
    QueryParams Q = P;  // + the coordinate hand-over buffer where the target transform is expensive
    void *xyz = nullptr;
    if (TK != PRS_TK_AREA_SEP && P.stage == 0) {
        PRS_CK(cudaMallocAsync(&xyz, 4 * sizeof(T) * (size_t)P.n_t, st));
        Q.target_xyz = xyz;
    }
    if (P.stage != 2) {
        PRS_CK(cudaMemsetAsync(n_live, 0, sizeof(uint32_t), st));
        classify_kernel<T, K, TK><<<(unsigned)((tiles + QWARPS - 1) / QWARPS), QWARPS * 32, 0, st>>>(Q, ix, (uint32_t)tiles, live, n_live);
        PRS_CKL();
    }
    if (P.stage != 1) {
        // search: persistent-style grid, a multiple of the SM count, blocks stride over the live list
        int64_t qgrid = (int64_t)n_sm * 32;
        if (qgrid > tiles) qgrid = tiles;
        query_kernel<T, K, TK><<<(unsigned)qgrid, QROWS * 32, 0, st>>>(Q, ix, live, n_live);
        PRS_CKL();
    }
    if (xyz) PRS_CK(cudaFreeAsync(xyz, st));
    if (P.stage == 0) PRS_CK(cudaFreeAsync(buf, st));
    return PRS_OK;
}

template <typename T, int TK> inline int launch_coverage_k(const QueryParams &P, uint8_t *cover, cudaStream_t st)
{
    if (P.n_t <= 0) return PRS_OK;
    coverage_kernel<T, TK><<<(unsigned)((P.n_t + 255) / 256), 256, 0, st>>>(P, cover);
    PRS_CKL();
    return PRS_OK;
}

template <typename T> int launch_coverage_t(const QueryParams &P, int tk, uint8_t *cover, cudaStream_t st)
{
    switch (tk) {
    case PRS_TK_LONLAT:
        return launch_coverage_k<T, PRS_TK_LONLAT>(P, cover, st);
    case PRS_TK_AREA_SEP:
        return launch_coverage_k<T, PRS_TK_AREA_SEP>(P, cover, st);
    default:
        return launch_coverage_k<T, PRS_TK_AREA_GEN>(P, cover, st);
    }
}

}  // namespace prs
