// The second- and third-generation attention kernels (attention2.cuh, attention3.cuh) in a translation unit of their own:
// their ~0.5 MB of instantiations (four softmax variants of v3, two of v2) compile in parallel with the GEMM / engine
// objects, and the trace build folds this file back into host_utils.cu so that one module owns the trace buffer.
#include "host_utils.h"

namespace uvlt {

int attn23_init_attributes() {
  UVLT_CUDA_OK(cudaFuncSetAttribute(attention2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn2Smem::TOTAL));
  UVLT_CUDA_OK(cudaFuncSetAttribute(attention2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn2Smem::TOTAL));
#define UVLT_AT3_ATTR(VAR) \
  UVLT_CUDA_OK(cudaFuncSetAttribute(attention3_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Smem::TOTAL))
  UVLT_AT3_ATTR(0); UVLT_AT3_ATTR(2); UVLT_AT3_ATTR(3); UVLT_AT3_ATTR(4);
  return 0;
}

int attn23_launch(const AttnLaunch& a, cudaStream_t stream) {
  if (a.v3) {
#define UVLT_AT3_LAUNCH(VAR) \
  UVLT_LAUNCH(attention3_kernel<VAR>, dim3(a.grid3), dim3(AT3_THREADS), Attn3Smem::TOTAL, stream, a.tma_qkv, a.tma_o, a.p3)
    switch (a.var3) {
      case 0: UVLT_AT3_LAUNCH(0); break;
      case 3: UVLT_AT3_LAUNCH(3); break;
      case 4: UVLT_AT3_LAUNCH(4); break;
      default: UVLT_AT3_LAUNCH(2); break;
    }
    UVLT_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const int ntiles = (a.p2.n + AT2_BQ - 1) / AT2_BQ;
  dim3 grid2(a.p2.split_all ? ntiles : (ntiles + 1) / 2, a.p2.H, a.B);
  if (a.poly) UVLT_LAUNCH(attention2_kernel<true>, grid2, dim3(AT2_THREADS), Attn2Smem::TOTAL, stream, a.tma_qkv, a.tma_o, a.p2);
  else UVLT_LAUNCH(attention2_kernel<false>, grid2, dim3(AT2_THREADS), Attn2Smem::TOTAL, stream, a.tma_qkv, a.tma_o, a.p2);
  UVLT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace uvlt
