// tcgen05 / TMEM / bulk-TMA coupling-stack kernel (GBNF_GEMM_F16_TC).  STUB: replaced in the next milestone.
#pragma once
#include <string>
#include <vector>
#include "common.cuh"

namespace gbnf {

struct TcPlan {
  size_t smem_bytes = 0;
  int tmem_cols = 0;
};

inline bool tc_make_plan(const ModelDims&, const std::vector<StepDesc>&, TcPlan*, std::string* why) {
  *why = "not built yet";
  return false;
}
inline cudaError_t tc_configure(const TcPlan&) { return cudaSuccess; }
inline int tc_launch(const CouplingArgs&, const TcPlan&, int, cudaStream_t) { return -1; }

}  // namespace gbnf
