// wxb_logmel.cu — K1: log-mel frontend, replaces whisperx/audio.py:112-159.
//
// Pass 1 (logmel_kernel): one CTA = 16 consecutive STFT frames of one chunk (LM_THREADS threads).  The 2800 input samples the frames
// cover are staged once in shared memory (16-byte loads where the tile lies inside the chunk; reflect padding at the chunk
// edges and zero padding beyond the chunk's valid length applied on the way in), two real frames share one 400-point complex
// FFT (radix 4,4,5,5 Stockham in shared memory, wxb_fft400.h), |X|^2 goes back to shared memory and the sparse (banded) mel
// filterbank — taps staged in shared memory as [n_mels][<= 16] — and log10 are applied from there.  Raw log10 values are
// written FRAME-MAJOR [chunk][frame][n_mels] (a warp writes 32 consecutive mel bins of one frame: 128-byte stores) and a
// per-chunk atomic max is folded, because the clamp `max(x, chunk_max - 8)` needs the chunk-wide maximum.
// Pass 2 (logmel_finalize_kernel): reads the raw tile (L2-resident), applies clamp and (x + 4) / 4, and writes what the caller
// asked for: the reference layout f32 [n_mels, n_frames] (transposed through shared memory) and / or the encoder's input
// melT bf16 [chunk * 3002 + 1 + frame][n_mels] directly — no f32 round trip and no separate transpose kernel on the hot path.
#include "wxb_common.cuh"
#include "wxb_fft400.h"
#include <math.h>

#define LM_FRAMES 16
#define LM_PAIRS 8
#ifndef LM_THREADS
#define LM_THREADS 320  // 10 warps per CTA, 3 CTAs per SM (shared memory): the kernel is latency-bound, 5-warp CTAs issued 1.2 instr/clk/SM
#endif
#define LM_HOP 160
#define LM_NFFT 400
#define LM_NBIN 201
#define LM_SAMPLES ((LM_FRAMES - 1) * LM_HOP + LM_NFFT)  // 2800

__device__ float g_lm_window[LM_NFFT];
__device__ float2 g_lm_tw[LM_NFFT];

// tap table of the mel rows in shared memory: row stride 15 floats for <= 80 rows (widest row of mel_80: 14 taps), 11 for
// <= 128 rows (mel_128: 9 taps) — odd strides, so the 32 lanes of a warp (32 different rows) hit 32 different banks.  Taps a
// wider custom filter row may have beyond the stride are read from global memory.
#define LM_COEF_FLOATS (128 * 11)

struct LmSmem {
  float samp[LM_SAMPLES];
  float win[LM_NFFT];
  cpx tw[LM_NFFT];
  cpx bufA[LM_PAIRS * LM_NFFT];
  cpx bufB[LM_PAIRS * LM_NFFT];
  float coef[LM_COEF_FLOATS];
  int2 band[128];
  float red[16];
};

// band[m] = (first non-zero bin, number of bins up to the last non-zero) of filter row m;
// also resets the per-chunk running max to -inf.
__global__ void logmel_setup_kernel(const float* __restrict__ filters, int n_mels, int2* __restrict__ band,
                                    float* __restrict__ chunk_max, int n_chunks) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < n_mels) {
    int first = -1, last = -1;
    for (int f = 0; f < LM_NBIN; ++f)
      if (filters[tid * LM_NBIN + f] != 0.f) {
        if (first < 0) first = f;
        last = f;
      }
    band[tid] = (first < 0) ? make_int2(0, 0) : make_int2(first, last - first + 1);
  }
  for (int c = tid; c < n_chunks; c += gridDim.x * blockDim.x) chunk_max[c] = -INFINITY;
}

__global__ void __launch_bounds__(LM_THREADS)
logmel_kernel(const float* __restrict__ audio, const long long* __restrict__ chunk_off,
              const int* __restrict__ chunk_len, int S, int n_frames, int n_mels,
              const float* __restrict__ filters, const int2* __restrict__ band,
              float* __restrict__ raw_out /* [chunk][frame][n_mels] */, float* __restrict__ chunk_max) {
  extern __shared__ __align__(16) unsigned char lm_smem_raw[];
  LmSmem& sm = *reinterpret_cast<LmSmem*>(lm_smem_raw);
  const int tid = threadIdx.x;
  const int chunk = blockIdx.y;
  const int frame0 = blockIdx.x * LM_FRAMES;
  const long long off = chunk_off[chunk];
  const int len = chunk_len[chunk];

  for (int i = tid; i < LM_NFFT; i += LM_THREADS) {
    sm.win[i] = g_lm_window[i];
    const float2 t = g_lm_tw[i];
    sm.tw[i] = cpx{t.x, t.y};
  }
  // filter taps of every mel row: coef[m][k] = F[m, band.x + k], k < band.y
  for (int m = tid; m < n_mels; m += LM_THREADS) sm.band[m] = __ldg(band + m);
  const int cld = n_mels > 80 ? 11 : 15;
  for (int i = tid; i < n_mels * cld; i += LM_THREADS) {
    const int m = i / cld, k = i - m * cld;
    const int2 bd = __ldg(band + m);
    sm.coef[i] = (k < bd.y) ? __ldg(filters + m * LM_NBIN + bd.x + k) : 0.f;
  }
  // stage samples: torch.stft(center=True, pad_mode="reflect") over the zero-padded chunk
  const int base_n = frame0 * LM_HOP - LM_NFFT / 2;
  if (base_n >= 0 && base_n + LM_SAMPLES <= len && ((off + base_n) & 3) == 0) {
    // the whole tile lies inside the chunk's valid samples: 16-byte loads
    const float4* src = reinterpret_cast<const float4*>(audio + off + base_n);
    for (int i = tid; i < LM_SAMPLES / 4; i += LM_THREADS) reinterpret_cast<float4*>(sm.samp)[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < LM_SAMPLES; i += LM_THREADS) {
      int n = base_n + i;
      if (n < 0) n = -n;
      if (n >= S) n = 2 * (S - 1) - n;
      float v = 0.f;
      if (n >= 0 && n < len) v = __ldg(audio + off + n);
      sm.samp[i] = v;
    }
  }
  __syncthreads();

  // pass 1 (radix 4, Ns = 1): window + pack two frames, no twiddles
  for (int b = tid; b < LM_PAIRS * 100; b += LM_THREADS) {
    const int p = b / 100, i = b - p * 100;
    const float* fa = sm.samp + (2 * p) * LM_HOP;
    const float* fb = fa + LM_HOP;
    cpx v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = i + 100 * q;
      const float w = sm.win[idx];
      v[q] = cpx{fa[idx] * w, fb[idx] * w};
    }
    dft4(v);
    cpx* out = sm.bufB + p * LM_NFFT + i * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) out[q] = v[q];
  }
  __syncthreads();
  // pass 2 (radix 4, Ns = 4): B -> A
  for (int b = tid; b < LM_PAIRS * 100; b += LM_THREADS) {
    const int p = b / 100, i = b - p * 100;
    fft400_butterfly<4, 4>(sm.bufB + p * LM_NFFT, sm.bufA + p * LM_NFFT, sm.tw, i);
  }
  __syncthreads();
  // pass 3 (radix 5, Ns = 16): A -> B
  for (int b = tid; b < LM_PAIRS * 80; b += LM_THREADS) {
    const int p = b / 80, i = b - p * 80;
    fft400_butterfly<5, 16>(sm.bufA + p * LM_NFFT, sm.bufB + p * LM_NFFT, sm.tw, i);
  }
  __syncthreads();
  // pass 4 (radix 5, Ns = 80): B -> A
  for (int b = tid; b < LM_PAIRS * 80; b += LM_THREADS) {
    const int p = b / 80, i = b - p * 80;
    fft400_butterfly<5, 80>(sm.bufB + p * LM_NFFT, sm.bufA + p * LM_NFFT, sm.tw, i);
  }
  __syncthreads();
  // separate the two real frames and take |X|^2 (bins 0..200); power rows overlay bufB
  float* s_pow = reinterpret_cast<float*>(sm.bufB);
  for (int idx = tid; idx < LM_PAIRS * LM_NBIN; idx += LM_THREADS) {
    const int p = idx / LM_NBIN, f = idx - p * LM_NBIN;
    const cpx z = sm.bufA[p * LM_NFFT + f];
    const cpx zc = sm.bufA[p * LM_NFFT + ((LM_NFFT - f) % LM_NFFT)];
    const float xar = 0.5f * (z.x + zc.x), xai = 0.5f * (z.y - zc.y);
    const float dr = z.x - zc.x, di = z.y + zc.y;
    s_pow[(2 * p) * LM_NBIN + f] = xar * xar + xai * xai;
    s_pow[(2 * p + 1) * LM_NBIN + f] = 0.25f * (dr * dr + di * di);
  }
  __syncthreads();
  // banded mel filterbank + log10; consecutive threads -> consecutive mel rows of one frame (frame-major stores)
  float lmax = -INFINITY;
  float* out_chunk = raw_out + (size_t)chunk * n_mels * n_frames;
  for (int idx = tid; idx < LM_FRAMES * n_mels; idx += LM_THREADS) {
    const int fr = idx / n_mels, m = idx - fr * n_mels;
    const int2 bd = sm.band[m];
    const float* crow = sm.coef + m * cld;
    const float* prow = s_pow + fr * LM_NBIN + bd.x;
    float acc = 0.f;
    const int nk = min(bd.y, cld);
    for (int k = 0; k < nk; ++k) acc = fmaf(crow[k], prow[k], acc);
    for (int k = nk; k < bd.y; ++k) acc = fmaf(__ldg(filters + m * LM_NBIN + bd.x + k), prow[k], acc);
    const float v = log10f(fmaxf(acc, 1e-10f));
    const int frame = frame0 + fr;
    if (frame < n_frames) {
      out_chunk[(size_t)frame * n_mels + m] = v;
      lmax = fmaxf(lmax, v);
    }
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) sm.red[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float m = sm.red[0];
    for (int w = 1; w < LM_THREADS / 32; ++w) m = fmaxf(m, sm.red[w]);
    atomic_max_float(chunk_max + chunk, m);
  }
}

// pass 2: v = (max(raw, chunk_max - 8) + 4) / 4 for a tile of 32 frames x n_mels of one chunk; written as f32 [n_mels, n_frames]
// (reference layout, through a shared-memory transpose) and / or as bf16 rows of the encoder's frame-major input
__global__ void __launch_bounds__(256)
logmel_finalize_kernel(const float* __restrict__ raw, const float* __restrict__ chunk_max, int n_frames, int n_mels,
                       float* __restrict__ mel_out, __nv_bfloat16* __restrict__ melT, int melT_rows_per_chunk) {
  __shared__ float tile[32][129];
  const int chunk = blockIdx.y, f0 = blockIdx.x * 32;
  const float floor_v = chunk_max[chunk] - 8.0f;
  const float* src = raw + ((size_t)chunk * n_frames + f0) * n_mels;
  const int nf = min(32, n_frames - f0);
  for (int i = threadIdx.x; i < nf * n_mels; i += 256) {
    const int fr = i / n_mels, m = i - fr * n_mels;
    const float v = (fmaxf(src[i], floor_v) + 4.0f) / 4.0f;
    tile[fr][m] = v;
    if (melT) melT[((size_t)chunk * melT_rows_per_chunk + 1 + f0 + fr) * n_mels + m] = __float2bfloat16_rn(v);
  }
  if (!mel_out) return;
  __syncthreads();
  float* dst = mel_out + (size_t)chunk * n_mels * n_frames + f0;
  for (int i = threadIdx.x; i < 32 * n_mels; i += 256) {
    const int m = i >> 5, fr = i & 31;
    if (fr < nf) dst[(size_t)m * n_frames + fr] = tile[fr][m];
  }
}

static int logmel_tables(wxb_ctx* ctx, cudaStream_t st) {
  if (ctx->lm_tables_ready) return WXB_OK;
  float win[LM_NFFT];
  float2 tw[LM_NFFT];
  for (int n = 0; n < LM_NFFT; ++n) {
    // torch.hann_window(400) is periodic: 0.5 - 0.5 cos(2 pi n / 400)
    win[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / LM_NFFT));
    const double a = -2.0 * M_PI * n / LM_NFFT;
    tw[n] = make_float2((float)cos(a), (float)sin(a));
  }
  WXB_CUDA(ctx, cudaMemcpyToSymbolAsync(g_lm_window, win, sizeof(win), 0, cudaMemcpyHostToDevice, st));
  WXB_CUDA(ctx, cudaMemcpyToSymbolAsync(g_lm_tw, tw, sizeof(tw), 0, cudaMemcpyHostToDevice, st));
  WXB_CUDA(ctx, cudaStreamSynchronize(st));  // win/tw are stack buffers
  ctx->lm_tables_ready = true;
  return WXB_OK;
}

// Phase 1 only: raw log10 mel [n_chunks, n_mels, n_frames] + per-chunk max (device pointers).
int wxb_logmel_raw(wxb_ctx* ctx, const float* audio_dev, const int64_t* chunk_off_host,
                   const int32_t* chunk_len_host, int n_chunks, int n_samples_padded, int n_mels,
                   const float* filters_dev, float* raw_out_dev, float** chunk_max_dev_out,
                   cudaStream_t st) {
  if (n_samples_padded < LM_NFFT)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_logmel: n_samples_padded=%d must be >= 400", n_samples_padded);
  if (n_mels <= 0 || n_mels > 128) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_logmel: n_mels=%d", n_mels);
  if (n_chunks > 65535) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_logmel: more than 65535 chunks per call");
  for (int c = 0; c < n_chunks; ++c)
    if (chunk_len_host[c] < 0 || chunk_len_host[c] > n_samples_padded || chunk_off_host[c] < 0)
      return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_logmel: chunk %d has len %d (padded %d)", c, chunk_len_host[c], n_samples_padded);
  int rc;
  if ((rc = logmel_tables(ctx, st)) != WXB_OK) return rc;
  const int n_frames = n_samples_padded / LM_HOP;
  // device meta: [chunk_off int64 x n][chunk_len int32 x n][chunk_max f32 x n]
  const size_t meta_bytes = (size_t)n_chunks * (8 + 4 + 4);
  if ((rc = wxb_reserve(ctx, ctx->ws_mel_max, meta_bytes)) != WXB_OK) return rc;
  if ((rc = wxb_reserve(ctx, ctx->ws_mel_band, sizeof(int2) * 128)) != WXB_OK) return rc;
  unsigned char* meta = (unsigned char*)ctx->ws_mel_max.p;
  long long* d_off = (long long*)meta;
  int* d_len = (int*)(meta + (size_t)n_chunks * 8);
  float* d_max = (float*)(meta + (size_t)n_chunks * 12);
  WXB_CUDA(ctx, cudaMemcpyAsync(d_off, chunk_off_host, (size_t)n_chunks * 8, cudaMemcpyHostToDevice, st));
  WXB_CUDA(ctx, cudaMemcpyAsync(d_len, chunk_len_host, (size_t)n_chunks * 4, cudaMemcpyHostToDevice, st));
  logmel_setup_kernel<<<1, 128, 0, st>>>(filters_dev, n_mels, (int2*)ctx->ws_mel_band.p, d_max, n_chunks);
  WXB_LAUNCH_CHECK(ctx);
  if ((rc = wxb_func_smem(ctx, logmel_kernel, (int)sizeof(LmSmem))) != WXB_OK) return rc;
  dim3 grid(ceil_div(n_frames, LM_FRAMES), n_chunks);
  logmel_kernel<<<grid, LM_THREADS, sizeof(LmSmem), st>>>(audio_dev, d_off, d_len, n_samples_padded, n_frames,
                                                         n_mels, filters_dev, (const int2*)ctx->ws_mel_band.p,
                                                         raw_out_dev, d_max);
  WXB_LAUNCH_CHECK(ctx);
  *chunk_max_dev_out = d_max;
  return WXB_OK;
}

// both passes; mel_out_dev (f32 [n_chunks, n_mels, n_frames]) and melT (bf16 encoder input) are each optional
static int logmel_run(wxb_ctx* ctx, const float* audio_dev, const int64_t* chunk_off_host, const int32_t* chunk_len_host, int n_chunks,
                      int n_samples_padded, int n_mels, const float* filters_dev, float* mel_out_dev, __nv_bfloat16* melT,
                      int melT_rows_per_chunk, cudaStream_t st) {
  const int n_frames = n_samples_padded / LM_HOP;
  float* raw = (float*)wxb_named(ctx, "mel.raw", (size_t)n_chunks * n_frames * n_mels * 4);
  if (!raw) return WXB_ERR_CUDA;
  float* d_max = nullptr;
  int rc = wxb_logmel_raw(ctx, audio_dev, chunk_off_host, chunk_len_host, n_chunks, n_samples_padded, n_mels, filters_dev, raw,
                          &d_max, st);
  if (rc != WXB_OK) return rc;
  logmel_finalize_kernel<<<dim3(ceil_div(n_frames, 32), n_chunks), 256, 0, st>>>(raw, d_max, n_frames, n_mels, mel_out_dev, melT,
                                                                                melT_rows_per_chunk);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}

extern "C" int wxb_logmel(wxb_ctx* ctx, const float* audio_dev, const int64_t* chunk_off_host,
                          const int32_t* chunk_len_host, int n_chunks, int n_samples_padded, int n_mels,
                          const float* filters_dev, float* mel_out_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (n_chunks == 0) return WXB_OK;
  if (!audio_dev || !chunk_off_host || !chunk_len_host || n_chunks < 0 || !filters_dev || !mel_out_dev)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_logmel: bad argument");
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  return logmel_run(ctx, audio_dev, chunk_off_host, chunk_len_host, n_chunks, n_samples_padded, n_mels, filters_dev, mel_out_dev,
                    nullptr, 0, (cudaStream_t)stream);
}

// K1 -> K2 hand-off on the device: the log-mel of 30 s chunks goes straight into the encoder's frame-major bf16 input
// (ctx workspace "enc.melT"); wxb_encode(ctx, NULL, n_chunks, ...) consumes it.  mel_out_dev may be NULL.
extern "C" int wxb_logmel_features(wxb_ctx* ctx, const float* audio_dev, const int64_t* chunk_off_host,
                                   const int32_t* chunk_len_host, int n_chunks, int n_mels, const float* filters_dev,
                                   float* mel_out_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!audio_dev || !chunk_off_host || !chunk_len_host || n_chunks <= 0 || !filters_dev)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_logmel_features: bad argument");
  if (n_mels % 8) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_logmel_features: n_mels=%d must be a multiple of 8", n_mels);
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int G1 = 3002;  // padded frames per chunk of the encoder input (wxb_encoder.cu)
  __nv_bfloat16* melT = (__nv_bfloat16*)wxb_named(ctx, "enc.melT", ((size_t)n_chunks * G1 + 2) * n_mels * 2);
  if (!melT) return WXB_ERR_CUDA;
  ctx->melT_chunks = 0;
  int rc = logmel_run(ctx, audio_dev, chunk_off_host, chunk_len_host, n_chunks, 480000, n_mels, filters_dev, mel_out_dev, melT, G1,
                      (cudaStream_t)stream);
  if (rc != WXB_OK) return rc;
  ctx->melT_chunks = n_chunks;
  ctx->melT_mels = n_mels;
  return WXB_OK;
}
