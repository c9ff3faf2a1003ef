// Host-side launch of the classification and search kernels.  The templates are instantiated explicitly in
// prs_classify_inst.cu (tree dtype) and prs_query_inst.cu (tree dtype, K, epilogue) - one translation unit each, so
// that the wide top-k lists, whose unrolled insertion networks take minutes to compile, build in parallel.
#pragma once
#include "prs_query.cuh"

namespace prs {

inline int64_t tile_count(const QueryParams &P)
{
    if (P.tk == PRS_TK_LONLAT) return (P.n_t + 32 * QROWS - 1) / (32 * QROWS);
    return (int64_t)P.tiles_x * ((P.nrows + QROWS - 1) / QROWS);
}

// Classification + fill of the dead tiles (classify_kernel); live[] / n_live receive the tiles left to search.
template <typename T, int TK>
int launch_classify_tk(const QueryParams &P, const IndexRef<T> &ix, int64_t tiles, uint32_t *live, uint32_t *n_live,
                       cudaStream_t st)
{
    PRS_CK(cudaMemsetAsync(n_live, 0, sizeof(uint32_t), st));
    classify_kernel<T, TK><<<(unsigned)((tiles + QWARPS - 1) / QWARPS), QWARPS * 32, 0, st>>>(P, ix, (uint32_t)tiles, live, n_live);
    PRS_CKL();
    return PRS_OK;
}
template <typename T>
int launch_classify(const QueryParams &P, const IndexRef<T> &ix, int64_t tiles, uint32_t *live, uint32_t *n_live,
                    cudaStream_t st)
{
    switch (P.tk) {
    case PRS_TK_LONLAT:
        return launch_classify_tk<T, PRS_TK_LONLAT>(P, ix, tiles, live, n_live, st);
    case PRS_TK_AREA_SEP:
        return launch_classify_tk<T, PRS_TK_AREA_SEP>(P, ix, tiles, live, n_live, st);
    default:
        return launch_classify_tk<T, PRS_TK_AREA_GEN>(P, ix, tiles, live, n_live, st);
    }
}

// Search of the live tiles (query_kernel): persistent-style grid, a multiple of the SM count, blocks stride over
// the live list.  Dynamic shared memory: the record staging buffer (+ the row stage of the nearest epilogue).
template <typename T, int K, int EPI>
int launch_search(const QueryParams &P, const IndexRef<T> &ix, int64_t tiles, const uint32_t *live,
                  const uint32_t *n_live, cudaStream_t st)
{
    int dev = 0;
    PRS_CK(cudaGetDevice(&dev));
    static int n_sm[64] = {0};
    static bool configured[64] = {false};
    const int slot = dev < 0 || dev >= 64 ? 0 : dev;
    // k >= 4: staged records + cell-table stretch; nearest (k = 1, never staged): the row stage of its epilogue
    const size_t smem = K >= 4 ? PRS_STAGE_BYTES(K) + STAGE_CELLS * sizeof(uint32_t)
                               : (EPI == PRS_EPI_NEAREST ? QROWS * 32 * STAGE_ROW_BYTES : 0);
    if (!configured[slot] || slot != dev) {
        PRS_CK(cudaDeviceGetAttribute(&n_sm[slot], cudaDevAttrMultiProcessorCount, dev));
        PRS_CK(cudaFuncSetAttribute(query_kernel<T, K, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[slot] = true;
    }
    int64_t qgrid = (int64_t)n_sm[slot] * 32;
    if (qgrid > tiles) qgrid = tiles;
    query_kernel<T, K, EPI><<<(unsigned)qgrid, QROWS * 32, smem, st>>>(P, ix, live, n_live);
    PRS_CKL();
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
