// wxb_model.cuh — the borrowed weight table (kernel-layout names, see whisperx/backends/b200_weights.py
// `to_kernel_layout`) and the per-model scratch the encoder / decoder reuse between calls.
#pragma once
#include "wxb_common.cuh"

struct wxb_model {
  wxb_dims dims;
  std::map<std::string, const void*> t;
  const void* get(const std::string& k) const {
    auto it = t.find(k);
    return it == t.end() ? nullptr : it->second;
  }
};

struct EncLayerW {
  const float *ln1_w, *ln1_b, *qkv_b, *out_b, *ln2_w, *ln2_b, *fc1_b, *fc2_b;
  const __nv_bfloat16 *qkv_w, *out_w, *fc1_w, *fc2_w;
};

struct DecLayerW {
  const float *ln1_w, *ln1_b, *qkv_b, *out_b, *ln2_w, *ln2_b, *cq_b, *ckv_b, *cout_b, *ln3_w, *ln3_b, *fc1_b, *fc2_b;
  const __nv_bfloat16 *qkv_w, *out_w, *cq_w, *ckv_w, *cout_w, *fc1_w, *fc2_w;
};

// Resolve "prefix.name" or fail with a message; returns nullptr on failure.
const void* wxb_weight(wxb_ctx* ctx, const std::string& name);
int wxb_enc_layer(wxb_ctx* ctx, int i, EncLayerW* out);
int wxb_dec_layer(wxb_ctx* ctx, int i, DecLayerW* out);
