// Explicit instantiations of the search launch (query_kernel), one translation unit per (tree dtype, K, epilogue):
//   nvcc -DPRS_INST_T=float -DPRS_INST_K=8 -DPRS_INST_EPI=2 -c prs_query_inst.cu
// PRS_INST_EPI: 0 neighbour info, 1 nearest gather (K = 1 only), 2 / 3 / 4 Gaussian weighting (f32, f64 over f32 data, f64).
#include "prs_launch.cuh"

#if !defined(PRS_INST_T) || !defined(PRS_INST_K) || !defined(PRS_INST_EPI)
#error "define PRS_INST_T (float|double), PRS_INST_K (1|4|8|16|32) and PRS_INST_EPI (0|1|2)"
#endif

namespace prs {
template int launch_search<PRS_INST_T, PRS_INST_K, PRS_INST_EPI>(const QueryParams &, const IndexRef<PRS_INST_T> &, int64_t,
                                                                 const uint32_t *, const uint32_t *, cudaStream_t);
}

#if defined(PRS_STATS)
// one copy of the counters per translation unit: the experiment reads the unit it exercises
extern "C" int PRS_STATS_NAME(unsigned long long *out)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, prs::g_stats, sizeof(prs::g_stats));
    unsigned long long zero[32] = {0};
    cudaMemcpyToSymbol(prs::g_stats, zero, sizeof(zero));
    return 0;
}
#endif
