// cf_mlp_tc.cu -- tcgen05 version of K-4 (placeholder until the tensor-core kernel lands).
#include "cf_common.cuh"

namespace cf {
size_t fusion_tc_workspace_bytes(int32_t C, int32_t mode)
{
    (void)mode;
    return (size_t)2 * C * C * sizeof(float);
}
int fusion_tc(const float *, const float *, const int32_t *, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t,
              float, float, float, float, const float *, int32_t, const float *, const float *, const float *,
              const float *, float *, int32_t, void *, cudaStream_t)
{
    set_error("cf_fusion_fwd: tcgen05 path not built yet; use CF_MODE_FP32_SIMT");
    return CF_ERR_UNSUPPORTED;
}
}  // namespace cf
