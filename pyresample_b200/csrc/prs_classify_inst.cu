// Explicit instantiation of the classification launch (classify_kernel for the three target kinds) and of the
// coverage launch, one translation unit per tree dtype:  nvcc -DPRS_INST_T=float -c prs_classify_inst.cu
#include "prs_launch.cuh"

#ifndef PRS_INST_T
#error "define PRS_INST_T (float|double)"
#endif

namespace prs {
template int launch_classify<PRS_INST_T>(const QueryParams &, const IndexRef<PRS_INST_T> &, int64_t, uint32_t *, uint32_t *,
                                         cudaStream_t);
template int launch_coverage_t<PRS_INST_T>(const QueryParams &, int, uint8_t *, cudaStream_t);
}
