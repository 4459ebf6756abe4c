#include "wxb_common.cuh"
extern "C" int wxb_decode_greedy(wxb_ctx* ctx, const void*, int, const int32_t*, int, const wxb_decode_opts*, int32_t*, int32_t*, float*, float*, void*) {
  return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_decode_greedy: not built yet");
}
extern "C" int wxb_decoder_logits(wxb_ctx* ctx, const void*, int, const int32_t*, int, float*, void*) {
  return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_decoder_logits: not built yet");
}
