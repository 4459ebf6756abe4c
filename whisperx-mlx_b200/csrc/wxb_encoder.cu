// wxb_encoder.cu — K2: the Whisper audio encoder (conv stem, 2 x GELU, sinusoid positions,
// pre-LN transformer blocks, ln_post).  Architecture: OpenAI Whisper AudioEncoder as ported by
// mlx_whisper (the reference's dependency; container oracle transformers/models/whisper/
// modeling_whisper.py:541-648).
//
// Data layout in HBM (B chunks, T = 1500 positions, d = n_audio_state):
//   melT  bf16 [B*3002 + 2, n_mels]  frame-major mel with one zero frame before and after each chunk,
//                                    so row r of the conv1 im2col matrix is the 3*n_mels contiguous
//                                    elements starting at r*n_mels (a TMA view with OVERLAPPING rows)
//   h1    bf16 [B*3002 + 2, d]       GELU(conv1), same padding; conv2 (stride 2) im2col row q is the
//                                    3*d contiguous elements starting at q*2*d
//   x     f32  [B*T, d]              residual stream
//   xn    bf16 [B*T, d]              LayerNorm output (GEMM A operand)
//   qkv   bf16 [B*T, 3d]             fused Q|K|V projection
//   att   bf16 [B*T, d]              attention output
//   hid   bf16 [B*T, 4d]             GELU(fc1)
// All GEMMs run on wxb_gemm.cu (tcgen05/TMEM/TMA) with bias / GELU / residual / positional add fused
// into the epilogue.
#include "wxb_gemm.cuh"
#include "wxb_model.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int T_AUDIO = 1500;
constexpr int T_MEL = 3000;
constexpr int G1 = 3002;  // padded frames per chunk (melT / h1)

// ---------------------------------------------------------------------------------------------
// mel f32 [B, n_mels, 3000] -> melT bf16 rows (b*3002 + 1 + frame), via a 32x32 smem transpose
// ---------------------------------------------------------------------------------------------
__global__ void mel_transpose_kernel(const float* __restrict__ mel, __nv_bfloat16* __restrict__ melT, int n_mels) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int f0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const float* src = mel + (size_t)b * n_mels * T_MEL;
  for (int i = ty; i < 32; i += 8) {
    const int m = m0 + i, f = f0 + tx;
    tile[i][tx] = (m < n_mels && f < T_MEL) ? src[(size_t)m * T_MEL + f] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int f = f0 + i, m = m0 + tx;
    if (f < T_MEL && m < n_mels) melT[((size_t)b * G1 + 1 + f) * n_mels + m] = __float2bfloat16_rn(tile[tx][i]);
  }
}

// zero the padding rows (row 0 and 3001 of each chunk, plus the 2 tail rows) of a [B*3002+2, width] bf16 buffer
__global__ void zero_pad_rows_kernel(__nv_bfloat16* __restrict__ buf, int B, int width) {
  const int row_id = blockIdx.x;  // 0 .. 2B+1
  long long row;
  if (row_id < 2 * B) row = (long long)(row_id >> 1) * G1 + ((row_id & 1) ? (G1 - 1) : 0);
  else row = (long long)B * G1 + (row_id - 2 * B);
  __nv_bfloat16* p = buf + row * width;
  for (int i = threadIdx.x; i < width; i += blockDim.x) p[i] = __float2bfloat16_rn(0.f);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5): f32 [rows, d] -> bf16 [rows, d]; one warp per row, row kept in registers
// ---------------------------------------------------------------------------------------------
template <int MAXV>  // MAXV float4 per lane: d <= 128 * MAXV
__global__ void __launch_bounds__(256)
layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                      __nv_bfloat16* __restrict__ y, long long rows, int d) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  const int nv = d >> 2;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      v[i] = xr[idx];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += a * a + bb * bb + c * c + e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  uint2* yr = reinterpret_cast<uint2*>(y + row * d);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float4 ww = __ldg(w4 + idx), bb = __ldg(b4 + idx);
      __nv_bfloat162 lo = __floats2bfloat162_rn((v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y);
      __nv_bfloat162 hi = __floats2bfloat162_rn((v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      yr[idx] = pk;
    }
  }
}

int launch_layernorm(wxb_ctx* ctx, const float* x, const float* w, const float* b, __nv_bfloat16* y, long long rows, int d,
                     cudaStream_t st) {
  if (d % 4 || d > 128 * 10) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "layernorm: d=%d", d);
  const unsigned grid = (unsigned)ceil_div64(rows, 8);
  if (d <= 512) layernorm_bf16_kernel<4><<<grid, 256, 0, st>>>(x, w, b, y, rows, d);
  else layernorm_bf16_kernel<10><<<grid, 256, 0, st>>>(x, w, b, y, rows, d);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}

}  // namespace

int wxb_attention_tc(wxb_ctx* ctx, const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int d, int H, const int* lens_dev,
                     cudaStream_t st);  // wxb_attn.cu

// exported to wxb_decoder.cu
int wxb_launch_layernorm(wxb_ctx* ctx, const float* x, const float* w, const float* b, __nv_bfloat16* y, long long rows,
                         int d, cudaStream_t st) {
  return launch_layernorm(ctx, x, w, b, y, rows, d, st);
}

#ifndef WXB_ENC_GROUP
// default chunk-group size of the layer stack (0 = whole batch).  Measured on the 60-chunk large-v3 job (tools/ab_enc_group.sh, one
// gpurun call): whole batch 170.5 ms, groups of 6 / 8 / 10 169.3-170.3, 12 / 15 166.4, 20 168.0, 30 170.2 ms.
#define WXB_ENC_GROUP 15
#endif
int wxb_encode_impl(wxb_ctx* ctx, const float* mel_dev, int B, __nv_bfloat16* enc_out, cudaStream_t st) {
  if (!ctx->model) return wxb_fail(ctx, WXB_ERR_STATE, "wxb_encode: no model set");
  const wxb_dims& D = ctx->model->dims;
  const int d = D.n_audio_state, nm = D.n_mels, H = D.n_audio_head;
  const long long M = (long long)B * T_AUDIO;
  const long long rows1 = (long long)B * G1 + 2;
  __nv_bfloat16* melT = (__nv_bfloat16*)wxb_named(ctx, "enc.melT", rows1 * nm * 2);
  __nv_bfloat16* h1 = (__nv_bfloat16*)wxb_named(ctx, "enc.h1", rows1 * d * 2);
  // Chunk groups: the conv stem runs over the whole batch, the layer stack over groups of Bs chunks one after the other, so
  // that what one kernel writes (xn, qkv, att, hid of a group) is still L2-resident when the next kernel reads it.  Results do
  // not depend on the grouping (every row's arithmetic is the same).
  const int eg = ctx->enc_group >= 0 ? ctx->enc_group : WXB_ENC_GROUP;
  const int Bs = (eg > 0 && eg < B) ? eg : B;
  const long long Mg = (long long)Bs * T_AUDIO;
  float* x = (float*)wxb_named(ctx, "enc.x", M * d * 4);
  __nv_bfloat16* xn = (__nv_bfloat16*)wxb_named(ctx, "enc.xn", Mg * d * 2);
  __nv_bfloat16* qkv = (__nv_bfloat16*)wxb_named(ctx, "enc.qkv", Mg * 3 * d * 2);
  __nv_bfloat16* att = (__nv_bfloat16*)wxb_named(ctx, "enc.att", Mg * d * 2);
  __nv_bfloat16* hid = (__nv_bfloat16*)wxb_named(ctx, "enc.hid", Mg * 4 * d * 2);
  if (!melT || !h1 || !x || !xn || !qkv || !att || !hid) return WXB_ERR_CUDA;

  const __nv_bfloat16* c1w = (const __nv_bfloat16*)wxb_weight(ctx, "enc.conv1.w");
  const float* c1b = (const float*)wxb_weight(ctx, "enc.conv1.b");
  const __nv_bfloat16* c2w = (const __nv_bfloat16*)wxb_weight(ctx, "enc.conv2.w");
  const float* c2b = (const float*)wxb_weight(ctx, "enc.conv2.b");
  const float* pos = (const float*)wxb_weight(ctx, "enc.pos");
  const float* lnp_w = (const float*)wxb_weight(ctx, "enc.ln_post.w");
  const float* lnp_b = (const float*)wxb_weight(ctx, "enc.ln_post.b");
  if (!c1w || !c1b || !c2w || !c2b || !pos || !lnp_w || !lnp_b) return WXB_ERR_STATE;

  // --- mel -> frame-major bf16 with zero frames around each chunk (already there when K1 handed it over: mel_dev == NULL)
  zero_pad_rows_kernel<<<2 * B + 2, 128, 0, st>>>(melT, B, nm);
  WXB_LAUNCH_CHECK(ctx);
  zero_pad_rows_kernel<<<2 * B + 2, 128, 0, st>>>(h1, B, d);
  WXB_LAUNCH_CHECK(ctx);
  if (mel_dev) {
    mel_transpose_kernel<<<dim3(ceil_div(T_MEL, 32), ceil_div(nm, 32), B), dim3(32, 8), 0, st>>>(mel_dev, melT, nm);
    WXB_LAUNCH_CHECK(ctx);
  }
  int rc;
  {  // conv1 (k3, s1, p1) + GELU as one GEMM over overlapping rows
    GemmArgs a;
    a.A = melT; a.lda = nm; a.M = B * G1; a.W = c1w; a.N = d; a.K = 3 * nm;
    a.bias = c1b; a.gelu = 1; a.out = h1; a.out_f32 = 0; a.ldo = d;
    a.g_in = G1; a.g_valid = T_MEL; a.g_out = G1; a.out_off = 1;
    if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
  }
  {  // conv2 (k3, s2, p1) + GELU + positional embedding -> residual stream
    GemmArgs a;
    a.A = h1; a.lda = 2LL * d; a.M = B * (T_AUDIO + 1); a.W = c2w; a.N = d; a.K = 3 * d;
    a.bias = c2b; a.gelu = 1; a.residual = pos; a.res_mode = 2; a.ldr = d;
    a.out = x; a.out_f32 = 1; a.ldo = d;
    a.g_in = T_AUDIO + 1; a.g_valid = T_AUDIO; a.g_out = T_AUDIO; a.out_off = 0;
    if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
  }
  for (int b0 = 0; b0 < B; b0 += Bs) {
    const int nb = (B - b0 < Bs) ? B - b0 : Bs;
    const long long Mb = (long long)nb * T_AUDIO;
    float* xg = x + (size_t)b0 * T_AUDIO * d;
    for (int l = 0; l < D.n_audio_layer; ++l) {
      EncLayerW w;
      if ((rc = wxb_enc_layer(ctx, l, &w)) != WXB_OK) return rc;
      if ((rc = launch_layernorm(ctx, xg, w.ln1_w, w.ln1_b, xn, Mb, d, st)) != WXB_OK) return rc;
      {
        GemmArgs a;
        a.A = xn; a.lda = d; a.M = (int)Mb; a.W = w.qkv_w; a.N = 3 * d; a.K = d; a.bias = w.qkv_b;
        a.out = qkv; a.ldo = 3 * d;
        if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
      }
      if ((rc = wxb_attention_tc(ctx, qkv, att, nb, T_AUDIO, d, H, nullptr, st)) != WXB_OK) return rc;
      {
        GemmArgs a;
        a.A = att; a.lda = d; a.M = (int)Mb; a.W = w.out_w; a.N = d; a.K = d; a.bias = w.out_b;
        a.residual = xg; a.res_mode = 1; a.ldr = d; a.out = xg; a.out_f32 = 1; a.ldo = d;
        if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
      }
      if ((rc = launch_layernorm(ctx, xg, w.ln2_w, w.ln2_b, xn, Mb, d, st)) != WXB_OK) return rc;
      {
        GemmArgs a;
        a.A = xn; a.lda = d; a.M = (int)Mb; a.W = w.fc1_w; a.N = 4 * d; a.K = d; a.bias = w.fc1_b; a.gelu = 1;
        a.out = hid; a.ldo = 4 * d;
        if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
      }
      {
        GemmArgs a;
        a.A = hid; a.lda = 4 * d; a.M = (int)Mb; a.W = w.fc2_w; a.N = d; a.K = 4 * d; a.bias = w.fc2_b;
        a.residual = xg; a.res_mode = 1; a.ldr = d; a.out = xg; a.out_f32 = 1; a.ldo = d;
        if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
      }
    }
    if ((rc = launch_layernorm(ctx, xg, lnp_w, lnp_b, enc_out + (size_t)b0 * T_AUDIO * d, Mb, d, st)) != WXB_OK) return rc;
  }
  return WXB_OK;
}

extern "C" int wxb_encoder_attention(wxb_ctx* ctx, const void* qkv_dev, void* out_dev, int B, int T, int d, int H, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!qkv_dev || !out_dev || B <= 0 || T <= 0 || H <= 0 || d != 64 * H)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_encoder_attention: bad argument (head_dim must be 64)");
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  return wxb_attention_tc(ctx, (const __nv_bfloat16*)qkv_dev, (__nv_bfloat16*)out_dev, B, T, d, H, nullptr, (cudaStream_t)stream);
}

extern "C" int wxb_encode(wxb_ctx* ctx, const float* mel_dev, int B, void* enc_out_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!enc_out_dev || B <= 0) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_encode: bad argument");
  if (!mel_dev) {
    // the log-mel was left in the encoder's input buffer by wxb_logmel_features
    if (!ctx->model) return wxb_fail(ctx, WXB_ERR_STATE, "wxb_encode: no model set");
    if (ctx->melT_chunks != B || ctx->melT_mels != ctx->model->dims.n_mels)
      return wxb_fail(ctx, WXB_ERR_STATE, "wxb_encode(mel = NULL): wxb_logmel_features left %d chunks of %d mel bins, asked for %d of %d",
                      ctx->melT_chunks, ctx->melT_mels, B, ctx->model->dims.n_mels);
  }
  ctx->melT_chunks = 0;  // consumed (or about to be overwritten by the transpose of mel_dev)
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  return wxb_encode_impl(ctx, mel_dev, B, (__nv_bfloat16*)enc_out_dev, (cudaStream_t)stream);
}
