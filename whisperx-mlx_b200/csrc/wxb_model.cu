// wxb_model.cu — wxb_set_model: records the borrowed device pointers by name.
#include "wxb_model.cuh"

void wxb_decoder_reset(wxb_ctx* ctx);  // wxb_decoder.cu: the cached layer / tensor-map tables hold weight pointers

void wxb_model_free(wxb_ctx* ctx) {
  wxb_decoder_reset(ctx);
  delete ctx->model;
  ctx->model = nullptr;
}

const void* wxb_weight(wxb_ctx* ctx, const std::string& name) {
  if (!ctx->model) {
    wxb_fail(ctx, WXB_ERR_STATE, "no model set (call wxb_set_model first)");
    return nullptr;
  }
  const void* p = ctx->model->get(name);
  if (!p) wxb_fail(ctx, WXB_ERR_STATE, "model tensor '%s' missing from wxb_set_model table", name.c_str());
  return p;
}

#define GETW(field, type, nm)                                         \
  do {                                                                \
    out->field = (type)wxb_weight(ctx, pre + nm);                     \
    if (!out->field) return WXB_ERR_STATE;                            \
  } while (0)

int wxb_enc_layer(wxb_ctx* ctx, int i, EncLayerW* out) {
  const std::string pre = "enc." + std::to_string(i) + ".";
  GETW(ln1_w, const float*, "ln1.w"); GETW(ln1_b, const float*, "ln1.b");
  GETW(qkv_w, const __nv_bfloat16*, "qkv.w"); GETW(qkv_b, const float*, "qkv.b");
  GETW(out_w, const __nv_bfloat16*, "out.w"); GETW(out_b, const float*, "out.b");
  GETW(ln2_w, const float*, "ln2.w"); GETW(ln2_b, const float*, "ln2.b");
  GETW(fc1_w, const __nv_bfloat16*, "fc1.w"); GETW(fc1_b, const float*, "fc1.b");
  GETW(fc2_w, const __nv_bfloat16*, "fc2.w"); GETW(fc2_b, const float*, "fc2.b");
  return WXB_OK;
}

int wxb_dec_layer(wxb_ctx* ctx, int i, DecLayerW* out) {
  const std::string pre = "dec." + std::to_string(i) + ".";
  GETW(ln1_w, const float*, "ln1.w"); GETW(ln1_b, const float*, "ln1.b");
  GETW(qkv_w, const __nv_bfloat16*, "qkv.w"); GETW(qkv_b, const float*, "qkv.b");
  GETW(out_w, const __nv_bfloat16*, "out.w"); GETW(out_b, const float*, "out.b");
  GETW(ln2_w, const float*, "ln2.w"); GETW(ln2_b, const float*, "ln2.b");
  GETW(cq_w, const __nv_bfloat16*, "cq.w"); GETW(cq_b, const float*, "cq.b");
  GETW(ckv_w, const __nv_bfloat16*, "ckv.w"); GETW(ckv_b, const float*, "ckv.b");
  GETW(cout_w, const __nv_bfloat16*, "cout.w"); GETW(cout_b, const float*, "cout.b");
  GETW(ln3_w, const float*, "ln3.w"); GETW(ln3_b, const float*, "ln3.b");
  GETW(fc1_w, const __nv_bfloat16*, "fc1.w"); GETW(fc1_b, const float*, "fc1.b");
  GETW(fc2_w, const __nv_bfloat16*, "fc2.w"); GETW(fc2_b, const float*, "fc2.b");
  return WXB_OK;
}

extern "C" int wxb_set_model(wxb_ctx* ctx, const wxb_dims* dims, const char* const* names,
                             const void* const* ptrs_dev, int n_tensors) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!dims || !names || !ptrs_dev || n_tensors <= 0) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_set_model: bad argument");
  const wxb_dims& d = *dims;
  if (d.n_audio_state % 64 || d.n_text_state % 64 || d.n_audio_state / d.n_audio_head != 64 ||
      d.n_text_state / d.n_text_head != 64)
    return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_set_model: head_dim must be 64 (d=%d heads=%d)", d.n_audio_state, d.n_audio_head);
  if (d.n_audio_ctx != 1500 || d.n_mels > 128 || d.n_mels % 8)
    return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_set_model: n_audio_ctx=%d n_mels=%d unsupported", d.n_audio_ctx, d.n_mels);
  wxb_model_free(ctx);
  ctx->model = new wxb_model();
  ctx->model->dims = d;
  for (int i = 0; i < n_tensors; ++i) {
    if (!names[i] || !ptrs_dev[i]) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_set_model: tensor %d is NULL", i);
    ctx->model->t[names[i]] = ptrs_dev[i];
  }
  return WXB_OK;
}
