// Instantiates the K1 scan kernels that score 2 queries per pass (see scan_topk_kernel.cuh).
#include "scan_topk_kernel.cuh"

namespace wdbx {
cudaError_t launch_scan_qb2(const ScanParams& p, const ScanPlan& plan, bool bf16, cudaStream_t stream) {
  return scan::launch_qb<2>(p, plan, bf16, stream);
}
}  // namespace wdbx
