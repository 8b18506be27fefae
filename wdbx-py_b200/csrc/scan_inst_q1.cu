// Instantiates the K1 scan kernels that score 1 query per pass (see scan_topk_kernel.cuh).
#include "scan_topk_kernel.cuh"

namespace wdbx {
cudaError_t launch_scan_qb1(const ScanParams& p, const ScanPlan& plan, bool bf16, cudaStream_t stream) {
  return scan::launch_qb<1>(p, plan, bf16, stream);
}
}  // namespace wdbx
