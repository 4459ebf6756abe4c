#include "wxb_common.cuh"
extern "C" int wxb_gemm_bf16(wxb_ctx* ctx, const void*, const void*, const float*, void*, int, int, int, int, void*) {
  return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_gemm_bf16: not built yet");
}
