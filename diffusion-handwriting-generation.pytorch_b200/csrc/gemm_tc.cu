// placeholder, replaced below
#include "kernels.h"
#include <stdio.h>
namespace dhg {
struct TcGemmPlan { int dummy; };
TcGemmPlan* tc_gemm_plan_create(const bf16*, int, int, const bf16*, int, int, int, const Epilogue&, char* err, int errlen) {
  snprintf(err, errlen, "tcgen05 GEMM not built yet");
  return nullptr;
}
void tc_gemm_plan_destroy(TcGemmPlan* p) { delete p; }
int tc_gemm_launch(const TcGemmPlan*, const Epilogue&, cudaStream_t) { return 1; }
}
