// One explicit instantiation of the query launch (classify_kernel + query_kernel for all target kinds) per
// translation unit:  nvcc -DPRS_INST_T=float -DPRS_INST_K=8 -c prs_query_inst.cu
#include "prs_launch.cuh"

#ifndef PRS_INST_T
#error "define PRS_INST_T (float|double) and PRS_INST_K (1|4|8|16|32)"
#endif

namespace prs {
template int launch_query_tk<PRS_INST_T, PRS_INST_K>(const QueryParams &, const IndexView<PRS_INST_T> &, int, cudaStream_t);
}
