// wxb_w2v.cu — the alignment model's forward pass (wav2vec2-base CTC, torchaudio WAV2VEC2_ASR_BASE_960H architecture),
// batched over ALL segments of a transcript.  Replaces the per-segment `model(waveform_segment)` call of the reference,
// /root/reference/whisperx/alignment.py:240-258 ("TODO: batched inference"); layer list per
// /root/reference/whisperx/convert_alignment_models.py:31-70 and torchaudio/models/wav2vec2/components.py:
//
//   waveform -> conv(1->512, k10 s5) -> GroupNorm(512 groups: per channel over time) -> GELU
//            -> 4 x [conv(512->512, k3 s2) -> GELU] -> 2 x [conv(512->512, k2 s2) -> GELU]          (no conv biases)
//            -> LayerNorm(512) -> Linear(512->768)
//            -> LayerNorm(x + GELU(grouped conv(768->768, k128, 16 groups, pad 64, last frame dropped)))   (weight-norm folded;
//               torchaudio builds the base model's Transformer with layer_norm_first = NOT encoder_layer_norm_first, i.e. its
//               `layer_norm` sits right after the positional conv and there is no final LayerNorm: components.py:419-426, :756)
//            -> 12 x [x = LN(x + MHA(x)); x = LN(x + W2 GELU(W1 x))]   (post-LN, 12 heads x 64)
//            -> Linear(768->n_out) = emission logits [T, n_out], T = floor((S - 400) / 320) + 1
//
// Data layout (B segments, P = padded frames per segment, frame-major / channels-last everywhere):
//   c0 bf16 [B * 64P, 512], c1 [B * 32P, 512] ... c6 [B * P, 512]: conv stack outputs.  Segment b owns rows b * (P << (6 - l)) ..;
//       because the padded lengths halve exactly with every stride-2 layer, output row r of layer l is the k * 512 contiguous
//       elements starting at input row 2 r: every conv is ONE tcgen05 GEMM over a tensor map with overlapping rows (row
//       stride s * 512 < row length k * 512), no im2col buffer.  Rows past a segment's valid length hold finite garbage that no
//       valid row ever reads (a valid output frame only reads valid input frames).
//   x f32 [B * P, 768] residual stream, xn bf16 [B * P, 768] GEMM operand, qkv bf16 [B * P, 2304], hid bf16 [B * P, 3072]
//   xg bf16 [16 groups][B][P + 128][48]: group-major, zero-padded copy of x for the positional conv, so that the im2col row of
//       (group g, frame t) is again one contiguous run (128 taps x 48 channels) — 16 GEMMs with N = 48.
// conv layer 0 (C_in = 1, K = 10) runs on the CUDA cores in fp32 (the waveform is not rounded to bf16): pass 1 accumulates the
// GroupNorm statistics per (segment, channel) in fp64, pass 2 recomputes the conv, normalises, applies GELU and stores bf16.
#include "wxb_gemm.cuh"
#include "wxb_model.cuh"
#include <math.h>

int wxb_attention_tc(wxb_ctx* ctx, const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int d, int H, const int* lens_dev,
                     cudaStream_t st);  // wxb_attn.cu

namespace {

constexpr int C_CONV = 512;
constexpr int N_CONV = 7;
constexpr int CONV_K[N_CONV] = {10, 3, 3, 3, 3, 2, 2};
constexpr int CONV_S[N_CONV] = {5, 2, 2, 2, 2, 2, 2};
constexpr int POS_PAD = 128;  // zero rows around every segment in the group-major positional-conv input (64 before, >= 63 after)

// ---------------------------------------------------------------------------------------------
// conv layer 0: y[t, c] = sum_j w[c, j] x[5 t + j], j < 10.  One block = 512 frames of one segment, thread = 2 channels.
// ---------------------------------------------------------------------------------------------
constexpr int C0_FRAMES = 512;
constexpr int C0_THREADS = 256;

struct W2vSeg {
  long long off;  // first sample in the packed audio buffer
  int S;          // samples
  int T1;         // valid frames after conv layer 0
  int T;          // valid frames after the conv stack (emission frames)
};

__device__ __forceinline__ void c0_stage(const float* __restrict__ audio, const W2vSeg& sg, int t0, float* xs) {
  // samples 5 t0 .. 5 (t0 + C0_FRAMES) + 4 (zero past the segment)
  const int n = 5 * C0_FRAMES + 5;
  for (int i = threadIdx.x; i < n; i += C0_THREADS) {
    const int s = 5 * t0 + i;
    xs[i] = (s < sg.S) ? __ldg(audio + sg.off + s) : 0.f;
  }
}

__global__ void __launch_bounds__(C0_THREADS)
w2v_conv0_stats_kernel(const float* __restrict__ audio, const W2vSeg* __restrict__ segs, const float* __restrict__ w0,
                       double* __restrict__ stats /* [B][512][2] */) {
  __shared__ float xs[5 * C0_FRAMES + 8];
  const int b = blockIdx.y, t0 = blockIdx.x * C0_FRAMES;
  const W2vSeg sg = segs[b];
  if (t0 >= sg.T1) return;
  c0_stage(audio, sg, t0, xs);
  const int c = 2 * threadIdx.x;
  float wa[10], wb[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) { wa[j] = __ldg(w0 + c * 10 + j); wb[j] = __ldg(w0 + (c + 1) * 10 + j); }
  __syncthreads();
  const int nt = min(C0_FRAMES, sg.T1 - t0);
  float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
  for (int t = 0; t < nt; ++t) {
    float ya = 0.f, yb = 0.f;
#pragma unroll
    for (int j = 0; j < 10; ++j) { const float x = xs[5 * t + j]; ya = fmaf(wa[j], x, ya); yb = fmaf(wb[j], x, yb); }
    sa += ya; qa = fmaf(ya, ya, qa);
    sb += yb; qb = fmaf(yb, yb, qb);
  }
  double* st = stats + ((size_t)b * C_CONV + c) * 2;
  atomicAdd(st + 0, (double)sa); atomicAdd(st + 1, (double)qa);
  atomicAdd(st + 2, (double)sb); atomicAdd(st + 3, (double)qb);
}

__global__ void __launch_bounds__(C0_THREADS)
w2v_conv0_apply_kernel(const float* __restrict__ audio, const W2vSeg* __restrict__ segs, const float* __restrict__ w0,
                       const double* __restrict__ stats, const float* __restrict__ gn_w, const float* __restrict__ gn_b,
                       __nv_bfloat16* __restrict__ c0, long long rows_per_seg) {
  __shared__ float xs[5 * C0_FRAMES + 8];
  const int b = blockIdx.y, t0 = blockIdx.x * C0_FRAMES;
  const W2vSeg sg = segs[b];
  if (t0 >= rows_per_seg) return;
  const int c = 2 * threadIdx.x;
  __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(c0 + ((size_t)b * rows_per_seg + t0) * C_CONV + c);
  const int nrows = (int)min((long long)C0_FRAMES, rows_per_seg - t0);
  if (t0 >= sg.T1) {  // rows past the segment's valid length: zeros (never read by a valid frame; kept finite)
    for (int t = 0; t < nrows; ++t) out[(size_t)t * (C_CONV / 2)] = __floats2bfloat162_rn(0.f, 0.f);
    return;
  }
  c0_stage(audio, sg, t0, xs);
  float wa[10], wb[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) { wa[j] = __ldg(w0 + c * 10 + j); wb[j] = __ldg(w0 + (c + 1) * 10 + j); }
  // GroupNorm with one channel per group: biased variance over the segment's valid frames, eps 1e-5
  const double* st = stats + ((size_t)b * C_CONV + c) * 2;
  const double n = (double)sg.T1;
  const double ma = st[0] / n, mb = st[2] / n;
  const float ra = (float)(1.0 / sqrt(fmax(st[1] / n - ma * ma, 0.0) + 1e-5)), rb = (float)(1.0 / sqrt(fmax(st[3] / n - mb * mb, 0.0) + 1e-5));
  const float ga = __ldg(gn_w + c) * ra, gb = __ldg(gn_w + c + 1) * rb;
  const float ba = __ldg(gn_b + c) - (float)ma * ga, bb = __ldg(gn_b + c + 1) - (float)mb * gb;
  __syncthreads();
  const int nt = min(nrows, sg.T1 - t0);
  for (int t = 0; t < nt; ++t) {
    float ya = 0.f, yb = 0.f;
#pragma unroll
    for (int j = 0; j < 10; ++j) { const float x = xs[5 * t + j]; ya = fmaf(wa[j], x, ya); yb = fmaf(wb[j], x, yb); }
    out[(size_t)t * (C_CONV / 2)] = __floats2bfloat162_rn(gelu_fast(fmaf(ya, ga, ba)), gelu_fast(fmaf(yb, gb, bb)));
  }
  for (int t = nt; t < nrows; ++t) out[(size_t)t * (C_CONV / 2)] = __floats2bfloat162_rn(0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over rows of width d (eps 1e-5), one warp per row: input bf16 or f32; outputs f32 (optional, may alias the
// input) and bf16 (optional).  normalize = 0 only converts (f32 -> bf16 copy of the residual stream).
// ---------------------------------------------------------------------------------------------
template <typename Tin, int MAXV>  // MAXV groups of 4 elements per lane: d <= 128 * MAXV
__global__ void __launch_bounds__(256)
w2v_layernorm_kernel(const Tin* x, const float* __restrict__ w, const float* __restrict__ b, float* y32 /* may alias x */,
                     __nv_bfloat16* __restrict__ y16, long long rows, int d, int normalize) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = d >> 2;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      if (sizeof(Tin) == 4) {
        v[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + row * d)[idx];
      } else {
        const uint2 raw = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + row * d)[idx];
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        v[i] = make_float4(a.x, a.y, c.x, c.y);
      }
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (normalize) {
    mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
        q += a * a + bb * bb + c * c + e * e;
      }
    }
    rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      float4 o = v[i];
      if (normalize) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + idx), bb = __ldg(reinterpret_cast<const float4*>(b) + idx);
        o = make_float4((v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y,
                        (v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w);
      }
      if (y32) reinterpret_cast<float4*>(y32 + row * d)[idx] = o;
      if (y16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(y16 + row * d)[idx] = pk;
      }
    }
  }
}

// x f32 [B * P, 768] -> xg bf16 [G][B][P + POS_PAD][cg] (cg = 768 / G channels per group), row u = t + 64 holds frame t of the
// segment for t < T_b, zeros elsewhere (the conv's zero padding and the frames past the segment's end)
__global__ void __launch_bounds__(256)
w2v_group_major_kernel(const float* __restrict__ x, const W2vSeg* __restrict__ segs, __nv_bfloat16* __restrict__ xg, int B, int P,
                       int d, int G) {
  const int cg = d / G;
  const int Pg = P + POS_PAD;
  const long long total = (long long)B * Pg * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % d);
    const long long r = i / d;
    const int u = (int)(r % Pg), b = (int)(r / Pg);
    const int t = u - 64;
    float v = 0.f;
    if (t >= 0 && t < segs[b].T) v = x[((long long)b * P + t) * d + ch];
    const int g = ch / cg, j = ch - g * cg;
    xg[(((long long)g * B + b) * Pg + u) * cg + j] = __float2bfloat16_rn(v);
  }
}

// emis_pad f32 [B * P, V] -> emis f32 [sum T_b, V] (segments back to back, the layout K4 reads)
__global__ void w2v_pack_kernel(const float* __restrict__ src, const W2vSeg* __restrict__ segs, const int* __restrict__ t_off,
                                float* __restrict__ dst, int P, int V) {
  const int b = blockIdx.y;
  const int T = segs[b].T;
  const long long n = (long long)T * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[(long long)t_off[b] * V + i] = src[(long long)b * P * V + i];
}

template <typename Tin>
int launch_ln(wxb_ctx* ctx, const Tin* x, const float* w, const float* b, float* y32, __nv_bfloat16* y16, long long rows, int d,
              int normalize, cudaStream_t st) {
  if (d % 4 || d > 128 * 8) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "w2v layernorm: d=%d", d);
  const unsigned grid = (unsigned)ceil_div64(rows, 8);
  if (d <= 512) w2v_layernorm_kernel<Tin, 4><<<grid, 256, 0, st>>>(x, w, b, y32, y16, rows, d, normalize);
  else w2v_layernorm_kernel<Tin, 8><<<grid, 256, 0, st>>>(x, w, b, y32, y16, rows, d, normalize);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}

const void* aw(wxb_ctx* ctx, const std::string& name) {
  if (!ctx->align_model) {
    wxb_fail(ctx, WXB_ERR_STATE, "no alignment model set (call wxb_set_align_model first)");
    return nullptr;
  }
  const void* p = ctx->align_model->get(name);
  if (!p) wxb_fail(ctx, WXB_ERR_STATE, "alignment-model tensor '%s' missing from the wxb_set_align_model table", name.c_str());
  return p;
}

}  // namespace

void wxb_align_model_free(wxb_ctx* ctx) {
  delete ctx->align_model;
  ctx->align_model = nullptr;
}

extern "C" int wxb_set_align_model(wxb_ctx* ctx, const wxb_w2v_dims* dims, const char* const* names, const void* const* ptrs_dev,
                                   int n_tensors) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!dims || !names || !ptrs_dev || n_tensors <= 0) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_set_align_model: bad argument");
  const wxb_w2v_dims& d = *dims;
  if (d.conv_dim != C_CONV || d.embed_dim % 64 || d.embed_dim / d.n_heads != 64 || d.embed_dim > 1024 || d.pos_kernel != 128 ||
      d.pos_groups <= 0 || d.embed_dim % d.pos_groups || (d.embed_dim / d.pos_groups) % 8 || d.n_layers <= 0 || d.n_out <= 0 ||
      d.ff_dim % 8)
    return wxb_fail(ctx, WXB_ERR_UNSUPPORTED,
                    "wxb_set_align_model: only the wav2vec2-base family is supported (group-norm extractor of 512 channels, "
                    "head_dim 64, embed_dim <= 1024, positional conv k128); got conv %d embed %d heads %d pos k%d g%d",
                    d.conv_dim, d.embed_dim, d.n_heads, d.pos_kernel, d.pos_groups);
  wxb_align_model_free(ctx);
  ctx->align_model = new wxb_model();
  ctx->align_dims = d;
  for (int i = 0; i < n_tensors; ++i) {
    if (!names[i] || !ptrs_dev[i]) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_set_align_model: tensor %d is NULL", i);
    ctx->align_model->t[names[i]] = ptrs_dev[i];
  }
  return WXB_OK;
}

extern "C" int wxb_w2v_frames(int n_samples) {
  long long t = n_samples;
  for (int l = 0; l < N_CONV; ++l) t = (t >= CONV_K[l]) ? (t - CONV_K[l]) / CONV_S[l] + 1 : 0;
  return (int)t;
}

extern "C" int wxb_w2v_emissions(wxb_ctx* ctx, const float* audio_dev, const int64_t* seg_off_host, const int32_t* seg_len_host,
                                 int n_seg, float* emis_out_dev, const int32_t* t_off_host, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (n_seg == 0) return WXB_OK;
  if (!audio_dev || !seg_off_host || !seg_len_host || n_seg < 0 || !emis_out_dev || !t_off_host)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_w2v_emissions: bad argument");
  if (!ctx->align_model) return wxb_fail(ctx, WXB_ERR_STATE, "wxb_w2v_emissions: no alignment model set");
  cudaStream_t st = (cudaStream_t)stream;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  const wxb_w2v_dims& D = ctx->align_dims;
  const int d = D.embed_dim, H = D.n_heads, FF = D.ff_dim, V = D.n_out, G = D.pos_groups, cg = d / G;
  const int B = n_seg;
  // ---- per-segment geometry; P = padded emission frames per segment (every conv layer's padded length is P << (6 - l))
  std::vector<W2vSeg> segs(B);
  int Smax = 0;
  for (int b = 0; b < B; ++b) {
    const int S = seg_len_host[b];
    if (S < 400 || seg_off_host[b] < 0)
      return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_w2v_emissions: segment %d has %d samples (minimum 400: pad short segments)", b, S);
    segs[b].off = seg_off_host[b];
    segs[b].S = S;
    segs[b].T1 = (S - 10) / 5 + 1;
    segs[b].T = wxb_w2v_frames(S);
    if (t_off_host[b + 1] - t_off_host[b] != segs[b].T)
      return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_w2v_emissions: t_off of segment %d spans %d frames, the model emits %d", b,
                      t_off_host[b + 1] - t_off_host[b], segs[b].T);
    if (S > Smax) Smax = S;
  }
  const int P = std::max(2, ceil_div(Smax, 320));
  if ((long long)B * 64 * P >= (1LL << 31) / 2) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_w2v_emissions: batch too large (%d x %d samples)", B, Smax);
  const long long M = (long long)B * P;  // rows of the transformer part
  // ---- workspaces
  W2vSeg* d_segs = (W2vSeg*)wxb_named(ctx, "w2v.segs", sizeof(W2vSeg) * B);
  int* d_lens = (int*)wxb_named(ctx, "w2v.lens", (size_t)4 * B * 2);
  double* stats = (double*)wxb_named(ctx, "w2v.stats", (size_t)B * C_CONV * 2 * 8);
  __nv_bfloat16* c[N_CONV];
  for (int l = 0; l < N_CONV; ++l) {
    const std::string nm = "w2v.c" + std::to_string(l);
    // zero on allocation: the 8 pad rows behind the last segment are read by the next layer's last (invalid) rows and are
    // never written; everything a masked attention key or a dropped GEMM row can see must be FINITE (0 x NaN = NaN)
    c[l] = (__nv_bfloat16*)wxb_named(ctx, nm.c_str(), ((size_t)B * ((size_t)P << (6 - l)) + 8) * C_CONV * 2, true);
    if (!c[l]) return WXB_ERR_CUDA;
  }
  __nv_bfloat16* ln512 = (__nv_bfloat16*)wxb_named(ctx, "w2v.ln512", (size_t)M * C_CONV * 2);
  float* x = (float*)wxb_named(ctx, "w2v.x", (size_t)M * d * 4);
  __nv_bfloat16* xn = (__nv_bfloat16*)wxb_named(ctx, "w2v.xn", (size_t)M * d * 2);
  __nv_bfloat16* qkv = (__nv_bfloat16*)wxb_named(ctx, "w2v.qkv", (size_t)M * 3 * d * 2);
  __nv_bfloat16* att = (__nv_bfloat16*)wxb_named(ctx, "w2v.att", (size_t)M * d * 2);
  __nv_bfloat16* hid = (__nv_bfloat16*)wxb_named(ctx, "w2v.hid", (size_t)M * FF * 2);
  const int Pg = P + POS_PAD;
  __nv_bfloat16* xg = (__nv_bfloat16*)wxb_named(ctx, "w2v.xg", ((size_t)G * B * Pg + 256) * cg * 2);
  float* emis_pad = (float*)wxb_named(ctx, "w2v.emis", (size_t)M * V * 4);
  if (!d_segs || !d_lens || !stats || !ln512 || !x || !xn || !qkv || !att || !hid || !xg || !emis_pad) return WXB_ERR_CUDA;
  std::vector<int> lens_toff(2 * B);
  for (int b = 0; b < B; ++b) { lens_toff[b] = segs[b].T; lens_toff[B + b] = t_off_host[b]; }
  WXB_CUDA(ctx, cudaMemcpyAsync(d_segs, segs.data(), sizeof(W2vSeg) * B, cudaMemcpyHostToDevice, st));
  WXB_CUDA(ctx, cudaMemcpyAsync(d_lens, lens_toff.data(), (size_t)4 * 2 * B, cudaMemcpyHostToDevice, st));
  WXB_CUDA(ctx, cudaMemsetAsync(stats, 0, (size_t)B * C_CONV * 2 * 8, st));

  const float* w0 = (const float*)aw(ctx, "w2v.conv0.w");
  const float* gn_w = (const float*)aw(ctx, "w2v.gn.w");
  const float* gn_b = (const float*)aw(ctx, "w2v.gn.b");
  if (!w0 || !gn_w || !gn_b) return WXB_ERR_STATE;
  int rc;
  const int stop = ctx->w2v_stop;
  // ---- conv layer 0 + GroupNorm + GELU (fp32 CUDA cores)
  {
    const long long rows0 = (long long)P << 6;
    dim3 grid((unsigned)ceil_div64(rows0, C0_FRAMES), B);
    w2v_conv0_stats_kernel<<<grid, C0_THREADS, 0, st>>>(audio_dev, d_segs, w0, stats);
    WXB_LAUNCH_CHECK(ctx);
    w2v_conv0_apply_kernel<<<grid, C0_THREADS, 0, st>>>(audio_dev, d_segs, w0, stats, gn_w, gn_b, c[0], rows0);
    WXB_LAUNCH_CHECK(ctx);
  }
  if (stop == 0) return WXB_OK;
  // ---- conv layers 1 .. 6: one GEMM each over overlapping rows (row stride 2 * 512, row length k * 512), GELU, no bias
  for (int l = 1; l < N_CONV; ++l) {
    const __nv_bfloat16* w = (const __nv_bfloat16*)aw(ctx, "w2v.conv" + std::to_string(l) + ".w");
    if (!w) return WXB_ERR_STATE;
    GemmArgs a;
    a.A = c[l - 1]; a.lda = (long long)CONV_S[l] * C_CONV; a.M = (int)((long long)B * ((long long)P << (6 - l)));
    a.W = w; a.N = C_CONV; a.K = CONV_K[l] * C_CONV;
    a.gelu = 1; a.out = c[l]; a.out_f32 = 0; a.ldo = C_CONV;
    if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
    if (stop == l) return WXB_OK;
  }
  // ---- feature projection: LayerNorm(512) -> Linear(512 -> d)
  {
    const float* lw = (const float*)aw(ctx, "w2v.fp.ln.w");
    const float* lb = (const float*)aw(ctx, "w2v.fp.ln.b");
    const __nv_bfloat16* w = (const __nv_bfloat16*)aw(ctx, "w2v.fp.w");
    const float* bias = (const float*)aw(ctx, "w2v.fp.b");
    if (!lw || !lb || !w || !bias) return WXB_ERR_STATE;
    if ((rc = launch_ln<__nv_bfloat16>(ctx, c[6], lw, lb, nullptr, ln512, M, C_CONV, 1, st)) != WXB_OK) return rc;
    GemmArgs a;
    a.A = ln512; a.lda = C_CONV; a.M = (int)M; a.W = w; a.N = d; a.K = C_CONV; a.bias = bias;
    a.out = x; a.out_f32 = 1; a.ldo = d;
    if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
  }
  if (stop == 7) return WXB_OK;
  // ---- convolutional positional embedding: x += GELU(grouped conv(x) + bias)
  {
    const __nv_bfloat16* w = (const __nv_bfloat16*)aw(ctx, "w2v.pos.w");  // [d, 128 * cg]: row = output channel, column = tap * cg + input channel of the group
    const float* bias = (const float*)aw(ctx, "w2v.pos.b");
    if (!w || !bias) return WXB_ERR_STATE;
    const long long total = (long long)B * Pg * d;
    w2v_group_major_kernel<<<(unsigned)std::min<long long>(ceil_div64(total, 256), 148 * 16), 256, 0, st>>>(x, d_segs, xg, B, P, d, G);
    WXB_LAUNCH_CHECK(ctx);
    for (int g = 0; g < G; ++g) {
      GemmArgs a;
      a.A = xg + (size_t)g * B * Pg * cg; a.lda = cg; a.M = B * Pg;
      a.W = w + (size_t)g * cg * (D.pos_kernel * cg); a.N = cg; a.K = D.pos_kernel * cg;
      a.bias = bias + g * cg; a.gelu = 1;
      a.residual = x + g * cg; a.res_mode = 1; a.ldr = d;
      a.out = x + g * cg; a.out_f32 = 1; a.ldo = d;
      a.g_in = Pg; a.g_valid = P; a.g_out = P; a.out_off = 0;  // im2col row u of a segment's block = output frame u
      if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
    }
    const float* lw = (const float*)aw(ctx, "w2v.ln.w");
    const float* lb = (const float*)aw(ctx, "w2v.ln.b");
    if (!lw || !lb) return WXB_ERR_STATE;
    if ((rc = launch_ln<float>(ctx, x, lw, lb, x, xn, M, d, 1, st)) != WXB_OK) return rc;  // transformer.layer_norm
  }
  if (stop == 8) return WXB_OK;
  // ---- 12 post-LN transformer layers
  for (int l = 0; l < D.n_layers; ++l) {
    const std::string pre = "w2v." + std::to_string(l) + ".";
    const __nv_bfloat16* qkv_w = (const __nv_bfloat16*)aw(ctx, pre + "qkv.w");
    const float* qkv_b = (const float*)aw(ctx, pre + "qkv.b");
    const __nv_bfloat16* out_w = (const __nv_bfloat16*)aw(ctx, pre + "out.w");
    const float* out_b = (const float*)aw(ctx, pre + "out.b");
    const float* ln1_w = (const float*)aw(ctx, pre + "ln1.w");
    const float* ln1_b = (const float*)aw(ctx, pre + "ln1.b");
    const __nv_bfloat16* fc1_w = (const __nv_bfloat16*)aw(ctx, pre + "fc1.w");
    const float* fc1_b = (const float*)aw(ctx, pre + "fc1.b");
    const __nv_bfloat16* fc2_w = (const __nv_bfloat16*)aw(ctx, pre + "fc2.w");
    const float* fc2_b = (const float*)aw(ctx, pre + "fc2.b");
    const float* ln2_w = (const float*)aw(ctx, pre + "ln2.w");
    const float* ln2_b = (const float*)aw(ctx, pre + "ln2.b");
    if (!qkv_w || !qkv_b || !out_w || !out_b || !ln1_w || !ln1_b || !fc1_w || !fc1_b || !fc2_w || !fc2_b || !ln2_w || !ln2_b)
      return WXB_ERR_STATE;
    {
      GemmArgs a;
      a.A = xn; a.lda = d; a.M = (int)M; a.W = qkv_w; a.N = 3 * d; a.K = d; a.bias = qkv_b; a.out = qkv; a.ldo = 3 * d;
      if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
    }
    if ((rc = wxb_attention_tc(ctx, qkv, att, B, P, d, H, d_lens, st)) != WXB_OK) return rc;
    {
      GemmArgs a;
      a.A = att; a.lda = d; a.M = (int)M; a.W = out_w; a.N = d; a.K = d; a.bias = out_b;
      a.residual = x; a.res_mode = 1; a.ldr = d; a.out = x; a.out_f32 = 1; a.ldo = d;
      if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
    }
    if ((rc = launch_ln<float>(ctx, x, ln1_w, ln1_b, x, xn, M, d, 1, st)) != WXB_OK) return rc;
    {
      GemmArgs a;
      a.A = xn; a.lda = d; a.M = (int)M; a.W = fc1_w; a.N = FF; a.K = d; a.bias = fc1_b; a.gelu = 1; a.out = hid; a.ldo = FF;
      if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
    }
    {
      GemmArgs a;
      a.A = hid; a.lda = FF; a.M = (int)M; a.W = fc2_w; a.N = d; a.K = FF; a.bias = fc2_b;
      a.residual = x; a.res_mode = 1; a.ldr = d; a.out = x; a.out_f32 = 1; a.ldo = d;
      if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
    }
    if ((rc = launch_ln<float>(ctx, x, ln2_w, ln2_b, x, xn, M, d, 1, st)) != WXB_OK) return rc;
    if (stop == 9 + l) return WXB_OK;
  }
  // ---- Linear(d -> n_out) on the last layer's output (xn); frames of every segment packed back to back
  {
    const __nv_bfloat16* w = (const __nv_bfloat16*)aw(ctx, "w2v.aux.w");
    const float* bias = (const float*)aw(ctx, "w2v.aux.b");
    if (!w || !bias) return WXB_ERR_STATE;
    GemmArgs a;
    a.A = xn; a.lda = d; a.M = (int)M; a.W = w; a.N = V; a.K = d; a.bias = bias; a.out = emis_pad; a.out_f32 = 1; a.ldo = V;
    if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
    w2v_pack_kernel<<<dim3(16, B), 256, 0, st>>>(emis_pad, d_segs, d_lens + B, emis_out_dev, P, V);
    WXB_LAUNCH_CHECK(ctx);
  }
  return WXB_OK;
}
