// tcgen05 / TMEM implicit-GEMM engine -- placeholder until the UTCHMMA kernel lands.
#include "common.cuh"
namespace hi3d { int validate_gemm(const hi3d_gemm_params* p, const char* who); }
extern "C" int hi3d_gemm_tc5(const hi3d_gemm_params* p, void* stream) {
  (void)p; (void)stream;
  hi3d::set_error("hi3d_gemm_tc5: tcgen05 engine not built in this revision");
  return -38;
}
