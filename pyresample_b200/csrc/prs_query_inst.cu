// Explicit instantiations of the query launch (classify_kernel + query_kernel):
//   nvcc -DPRS_INST_T=float -DPRS_INST_K=8 -c prs_query_inst.cu                     all three target kinds
//   nvcc -DPRS_INST_T=double -DPRS_INST_K=32 -DPRS_INST_TK=2 -c prs_query_inst.cu   one target kind
#include "prs_launch.cuh"

#ifndef PRS_INST_T
#error "define PRS_INST_T (float|double), PRS_INST_K (1|4|8|16|32) and optionally PRS_INST_TK (0|1|2)"
#endif

namespace prs {
#define PRS_INSTANTIATE(TK) \
    template int launch_query_k<PRS_INST_T, PRS_INST_K, TK>(const QueryParams &, const IndexView<PRS_INST_T> &, cudaStream_t);
#ifdef PRS_INST_TK
PRS_INSTANTIATE(PRS_INST_TK)
#else
PRS_INSTANTIATE(PRS_TK_LONLAT)
PRS_INSTANTIATE(PRS_TK_AREA_SEP)
PRS_INSTANTIATE(PRS_TK_AREA_GEN)
#endif
}
