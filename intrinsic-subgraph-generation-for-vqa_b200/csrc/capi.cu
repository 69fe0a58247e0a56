// capi.cu — version and error strings of the libisg.so C ABI (include/isg.h).
#include "common.cuh"

extern "C" int isg_version(void) { return 100; }  // 0.1.0

extern "C" const char* isg_error_string(int code) {
  switch (code) {
    case ISG_OK: return "ok";
    case ISG_EINVAL: return "ISG_EINVAL: null pointer, negative size or inconsistent arguments";
    case ISG_EUNSUPPORTED: return "ISG_EUNSUPPORTED: shape / dtype / mode outside what the sm_100a kernels are built for";
    case ISG_EWORKSPACE: return "ISG_EWORKSPACE: workspace missing or too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown isg error";
}
