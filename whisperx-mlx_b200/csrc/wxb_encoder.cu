#include "wxb_common.cuh"
extern "C" int wxb_encode(wxb_ctx* ctx, const float*, int, void*, void*) {
  return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_encode: not built yet");
}
