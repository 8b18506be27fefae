// scan_topk.cu -- launch planning and dispatch for K1 (kernel in scan_topk_kernel.cuh, instantiated
// per queries-per-pass in scan_inst_q{1,2,4,8}.cu so the translation units compile in parallel).
#include <algorithm>

#include "kernels.h"

namespace wdbx {

cudaError_t launch_scan_qb1(const ScanParams&, const ScanPlan&, bool, cudaStream_t);
cudaError_t launch_scan_qb2(const ScanParams&, const ScanPlan&, bool, cudaStream_t);
cudaError_t launch_scan_qb4(const ScanParams&, const ScanPlan&, bool, cudaStream_t);
cudaError_t launch_scan_qb8(const ScanParams&, const ScanPlan&, bool, cudaStream_t);

namespace {
size_t up128(size_t x) { return (x + 127) & ~static_cast<size_t>(127); }
}  // namespace

int scan_plan(int dim, int dpad, int elem_bytes, int k, int B, int sm_count, const ScanTuning& tune, ScanPlan* plan) {
  if (dim <= 0 || dpad < dim || k <= 0 || k > kMaxK || B <= 0) return -1;
  const int row_bytes = dpad * elem_bytes;
  if (row_bytes % 16 != 0) return -1;
  const int cpr = row_bytes / 16;          // 16-byte chunks per row
  int lpr = 32;                            // lanes per row: >= 3 chunks per lane when possible
  while (lpr > 1 && cpr < 3 * lpr) lpr >>= 1;
  int lpr_log2 = 0;
  while ((1 << lpr_log2) < lpr) ++lpr_log2;
  const int G = 32 / lpr;
  const size_t budget = 226 * 1024;

  int U = tune.rows_unroll > 0 ? tune.rows_unroll : 4;
  if (U != 1 && U != 2 && U != 4) U = 4;
  U = std::min(U, lpr);                    // the kernel spreads QB*U sums over the lpr lanes of a row
  int QB = 1;
  const int want = tune.queries_per_pass > 0 ? tune.queries_per_pass : 8;
  while (QB < want && QB < B && QB < 8 && QB * 2 * U <= lpr) QB <<= 1;
  while (QB > 1 && static_cast<size_t>(QB) * k > 2048) QB >>= 1;   // keep the running lists small

  for (;;) {
    // QB*U > 8 needs more than 128 registers/thread => at most 8 warps per CTA
    const int max_warps = (QB * U > 8) ? 8 : 16;
    // measured on B200 (tools/quick_perf.py): 16 warps x 1 stage beats 8 x 2 when the register block allows it
    int warps = tune.warps > 0 ? std::min(tune.warps, max_warps) : max_warps;
    // k > 32: list inserts are longer, a second stage per warp keeps the HBM stream busy meanwhile
    // (measured: 12.5M x 384 bf16 k=100: 1502 -> 1363 us; k=10: 1 stage is 6 % faster)
    int stages = tune.stages > 0 ? std::min(tune.stages, 8) : ((warps > 8 && k <= 32) ? 1 : 2);
    auto fixed = [&](int w, int st) {
      return up128(static_cast<size_t>(QB) * dpad * 4) + 128 + up128(static_cast<size_t>(w) * st * 8) +
             up128(static_cast<size_t>(w) * QB * k * 8);
    };
    auto stage_bytes_for = [&](int u) { return static_cast<size_t>(u) * G * row_bytes; };
    auto total = [&](int w, int st, int u) {
      return fixed(w, st) + std::max(stage_bytes_for(u) * w * st, static_cast<size_t>(w) * (k <= 32 ? 32 : (k <= 128 ? 128 : k)) * 8);
    };
    int u = U;
    while (stages > 2 && total(warps, stages, u) > budget) --stages;
    while (u > 1 && total(warps, stages, u) > budget) u >>= 1;
    while (warps > 1 && total(warps, stages, u) > budget) warps >>= 1;
    while (stages > 1 && total(warps, stages, u) > budget) --stages;
    if (total(warps, stages, u) > budget) {
      if (QB > 1) { QB >>= 1; continue; }   // retry with fewer queries per pass
      return -4;                            // a single row does not fit a stage
    }
    plan->lpr_log2 = lpr_log2;
    plan->nch = (cpr + lpr - 1) / lpr;
    plan->U = u;
    plan->KS = 0;
    plan->tile_rows = u * G;
    plan->stage_bytes = static_cast<int>(stage_bytes_for(u));
    plan->stages = stages;
    plan->warps = warps;
    plan->grid = tune.grid > 0 ? tune.grid : sm_count;
    plan->smem_bytes = total(warps, stages, u);
    plan->queries_per_block = QB;
    plan->pdl = 0;
    return 0;
  }
}

cudaError_t launch_scan_topk(const ScanParams& p, const ScanPlan& plan, bool bf16, cudaStream_t stream) {
  switch (plan.queries_per_block) {
    case 8: return launch_scan_qb8(p, plan, bf16, stream);
    case 4: return launch_scan_qb4(p, plan, bf16, stream);
    case 2: return launch_scan_qb2(p, plan, bf16, stream);
    default: return launch_scan_qb1(p, plan, bf16, stream);
  }
}

}  // namespace wdbx
