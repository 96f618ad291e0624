// bf16 tcgen05 GEMM path (placeholder until the tensor-core kernels land): fails loudly.
#include "common.cuh"
#include "../../include/sd_b200.h"
int sd_gemm_tc_dispatch(const sd_gemm_desc* d, void* stream) { (void)d; (void)stream; return SD_ERR_UNSUPPORTED; }
