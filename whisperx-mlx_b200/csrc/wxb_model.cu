// wxb_model.cu — model table (wxb_set_model); filled in with the encoder/decoder.
#include "wxb_common.cuh"
void wxb_model_free(wxb_ctx* ctx) { (void)ctx; }
extern "C" int wxb_set_model(wxb_ctx* ctx, const wxb_dims*, const char* const*, const void* const*, int) {
  return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_set_model: not built yet");
}
