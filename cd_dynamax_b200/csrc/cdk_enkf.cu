// cdk_enkf.cu -- CD-EnKF (placeholder until the ensemble kernel lands).
#include "cdk_common.cuh"

namespace cdk {
template <typename T>
int launch_enkf(const KArgs<T>&, cudaStream_t) {
  return CDK_E_UNSUPPORTED;
}
template int launch_enkf<double>(const KArgs<double>&, cudaStream_t);
template int launch_enkf<float>(const KArgs<float>&, cudaStream_t);
}  // namespace cdk
