// wxb_logmel.cu — K1: log-mel frontend, replaces whisperx/audio.py:112-159.
//
// One CTA = 16 consecutive STFT frames of one chunk (160 threads).  The 2800 input samples the
// frames cover are staged once in shared memory (reflect padding at the chunk edges and zero
// padding beyond the chunk's valid length applied on the way in), two real frames share one
// 400-point complex FFT (radix 4,4,5,5 Stockham in shared memory, wxb_fft400.h), |X|^2 goes
// back to shared memory and the sparse (banded) mel filterbank + log10 are applied from there.
// Output frames are written frame-contiguous per mel row (reference layout [n_mels, n_frames]).
// The clamp `max(x, chunk_max - 8)` needs the chunk-wide max: phase 1 writes raw log10 values
// and folds a per-chunk atomic max; phase 2 (logmel_finalize_kernel, L2-resident re-read)
// applies the clamp and the (x+4)/4 scaling.
#include "wxb_common.cuh"
#include "wxb_fft400.h"
#include <math.h>

#define LM_FRAMES 16
#define LM_PAIRS 8
#define LM_THREADS 160
#define LM_HOP 160
#define LM_NFFT 400
#define LM_NBIN 201
#define LM_SAMPLES ((LM_FRAMES - 1) * LM_HOP + LM_NFFT)  // 2800

__device__ float g_lm_window[LM_NFFT];
__device__ float2 g_lm_tw[LM_NFFT];

struct LmSmem {
  float samp[LM_SAMPLES];
  float win[LM_NFFT];
  cpx tw[LM_NFFT];
  cpx bufA[LM_PAIRS * LM_NFFT];
  cpx bufB[LM_PAIRS * LM_NFFT];
  float red[8];
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
              float* __restrict__ raw_out, float* __restrict__ chunk_max) {
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
  // stage samples: torch.stft(center=True, pad_mode="reflect") over the zero-padded chunk
  const int base_n = frame0 * LM_HOP - LM_NFFT / 2;
  for (int i = tid; i < LM_SAMPLES; i += LM_THREADS) {
    int n = base_n + i;
    if (n < 0) n = -n;
    if (n >= S) n = 2 * (S - 1) - n;
    float v = 0.f;
    if (n >= 0 && n < len) v = __ldg(audio + off + n);
    sm.samp[i] = v;
  }
  __syncthreads();

  // pass 1 (radix 4, Ns = 1): window + pack two frames, no twiddles
#pragma unroll
  for (int r = 0; r < (LM_PAIRS * 100) / LM_THREADS; ++r) {
    const int b = tid + LM_THREADS * r;
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
#pragma unroll
  for (int r = 0; r < (LM_PAIRS * 100) / LM_THREADS; ++r) {
    const int b = tid + LM_THREADS * r;
    const int p = b / 100, i = b - p * 100;
    fft400_butterfly<4, 4>(sm.bufB + p * LM_NFFT, sm.bufA + p * LM_NFFT, sm.tw, i);
  }
  __syncthreads();
  // pass 3 (radix 5, Ns = 16): A -> B
#pragma unroll
  for (int r = 0; r < (LM_PAIRS * 80) / LM_THREADS; ++r) {
    const int b = tid + LM_THREADS * r;
    const int p = b / 80, i = b - p * 80;
    fft400_butterfly<5, 16>(sm.bufA + p * LM_NFFT, sm.bufB + p * LM_NFFT, sm.tw, i);
  }
  __syncthreads();
  // pass 4 (radix 5, Ns = 80): B -> A
#pragma unroll
  for (int r = 0; r < (LM_PAIRS * 80) / LM_THREADS; ++r) {
    const int b = tid + LM_THREADS * r;
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
  // banded mel filterbank + log10; consecutive threads -> consecutive frames of one mel row
  float lmax = -INFINITY;
  float* out_chunk = raw_out + (size_t)chunk * n_mels * n_frames;
  for (int idx = tid; idx < LM_FRAMES * n_mels; idx += LM_THREADS) {
    const int fr = idx & (LM_FRAMES - 1), m = idx >> 4;
    const int2 bd = __ldg(band + m);
    const float* frow = filters + m * LM_NBIN + bd.x;
    const float* prow = s_pow + fr * LM_NBIN + bd.x;
    float acc = 0.f;
    for (int k = 0; k < bd.y; ++k) acc = fmaf(__ldg(frow + k), prow[k], acc);
    const float v = log10f(fmaxf(acc, 1e-10f));
    const int frame = frame0 + fr;
    if (frame < n_frames) {
      out_chunk[(size_t)m * n_frames + frame] = v;
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

// phase 2: out = (max(raw, chunk_max - 8) + 4) / 4, in place
__global__ void logmel_finalize_kernel(float* __restrict__ x, const float* __restrict__ chunk_max,
                                       long long per_chunk, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const float floor_v = chunk_max[i / per_chunk] - 8.0f;
    x[i] = (fmaxf(x[i], floor_v) + 4.0f) / 4.0f;
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

extern "C" int wxb_logmel(wxb_ctx* ctx, const float* audio_dev, const int64_t* chunk_off_host,
                          const int32_t* chunk_len_host, int n_chunks, int n_samples_padded, int n_mels,
                          const float* filters_dev, float* mel_out_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (n_chunks == 0) return WXB_OK;
  if (!audio_dev || !chunk_off_host || !chunk_len_host || n_chunks < 0 || !filters_dev || !mel_out_dev)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_logmel: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  float* d_max = nullptr;
  int rc = wxb_logmel_raw(ctx, audio_dev, chunk_off_host, chunk_len_host, n_chunks, n_samples_padded, n_mels,
                          filters_dev, mel_out_dev, &d_max, st);
  if (rc != WXB_OK) return rc;
  const long long per_chunk = (long long)n_mels * (n_samples_padded / LM_HOP);
  const long long total = per_chunk * n_chunks;
  const int threads = 256;
  long long blocks = ceil_div64(total, threads * 4);
  if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  logmel_finalize_kernel<<<(unsigned)blocks, threads, 0, st>>>(mel_out_dev, d_max, per_chunk, total);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}
