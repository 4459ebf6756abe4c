// wxb_dtw.cu — word timing from the decoder's cross-attention.
//
// Reference: /root/reference/mlx_whisper_optimized_final.py
//   :37-125   CrossAttentionBatchInference keeps, for every decode forward, the pre-softmax cross-attention scores of the last
//             query token of every layer (on the host, as Python lists of MLX arrays)
//   :128-253  extract_words_with_dtw: mean over the alignment heads, softmax(10 x) over the frames, median filter 7
//             (/root/reference/median_filter_fix.py:6-22), zero mean / unit variance per token, dtw(-w^T)
//             (mlx_whisper.timing.dtw = OpenAI whisper/timing.py dtw_cpu + backtrace), word grouping on the host.
//
// Here nothing leaves the GPU until the path: the decode kernel (wxb_decoder.cu) logs the scaled cross-attention QUERY of
// every alignment head at every position (64 floats per head: 2.5 KB per token at 10 heads, instead of the reference's
// 32 layers x 20 heads x 1500 scores); the scores are rebuilt afterwards from the cross-K cache that is still resident:
//
//   dtw_scores_kernel   qk[b, s, t] = 1/A sum_a q[b, pos0 + s, a, :] . K[l_a][b][h_a][t][:]        fp32 FMA, K read as bf16
//   dtw_cost_kernel     per (b, s) row: softmax(temperature x), median filter (reflect), (w - mean) / (std + 1e-8), negated
//   dtw_path_kernel     one CTA per sequence: anti-diagonal wavefront over the [frames + 1, tokens + 1] cost table (thread j owns
//                       token column j; every cell does the reference's `x + min(c0, c1, c2)` with its tie rules, so the path
//                       is bit-exact for a given cost matrix), 2-bit trace packed in shared memory, backtrace by one thread.
#include "wxb_model.cuh"
#include <math.h>

namespace {

constexpr int DTW_T_MAX = 1536;     // frames (N of the DP)
constexpr int DTW_M_MAX = 448;      // tokens (M of the DP)
constexpr int SC_TILE = 64;         // scores kernel: 64 positions x 64 frames per CTA, 4 x 4 per thread
constexpr int SC_LD = SC_TILE + 4;  // padded leading dimension of the [dim][row] shared tiles (16-byte aligned rows)

// ---------------------------------------------------------------------------------------------------------------------
// qk[b, s, t] = 1/A sum_a sum_d q[b, pos0 + s, a, d] K_a[b, t, d]
// grid (frame tiles, position tiles, B), 256 threads; thread (ty, tx) owns positions 4 ty .. 4 ty + 3 x frames 4 tx .. 4 tx + 3
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dtw_scores_kernel(const float* __restrict__ qlog, const __nv_bfloat16* __restrict__ cross_kv,
                                                         const int* __restrict__ heads /*[A][2]*/, const int* __restrict__ n_rows,
                                                         const int* __restrict__ row_off, float* __restrict__ qk, int A, int B0, int H,
                                                         int TX, int T, int pos0) {
  const int b = blockIdx.z, s0 = blockIdx.y * SC_TILE, t0 = blockIdx.x * SC_TILE;
  const int nr = n_rows[b];
  if (s0 >= nr) return;
  __shared__ __align__(16) float sq[64 * SC_LD];  // [d][s]
  __shared__ __align__(16) float sk[64 * SC_LD];  // [d][t]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int a = 0; a < A; ++a) {
    const int l = heads[2 * a], h = heads[2 * a + 1];
    const __nv_bfloat16* Kb = cross_kv + (((size_t)(l * 2) * B0 + b) * H + h) * (size_t)T * 64;
    __syncthreads();
    // q tile: 64 positions x 64 dims (f32), transposed into [d][s]; rows past n_rows are zero
    for (int e = tid; e < 64 * 16; e += 256) {
      const int s = e >> 4, d4 = (e & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s0 + s < nr) v = *reinterpret_cast<const float4*>(qlog + (((size_t)b * TX + pos0 + s0 + s) * A + a) * 64 + d4);
      sq[(d4 + 0) * SC_LD + s] = v.x; sq[(d4 + 1) * SC_LD + s] = v.y; sq[(d4 + 2) * SC_LD + s] = v.z; sq[(d4 + 3) * SC_LD + s] = v.w;
    }
    // K tile: 64 frames x 64 dims (bf16 -> f32), transposed into [d][t]; frames past T are zero
    for (int e = tid; e < 64 * 8; e += 256) {
      const int t = e >> 3, d8 = (e & 7) * 8;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (t0 + t < T) v = *reinterpret_cast<const uint4*>(Kb + (size_t)(t0 + t) * 64 + d8);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        sk[(d8 + 2 * j) * SC_LD + t] = f.x;
        sk[(d8 + 2 * j + 1) * SC_LD + t] = f.y;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      const float4 qv = *reinterpret_cast<const float4*>(sq + d * SC_LD + 4 * ty);
      const float4 kv = *reinterpret_cast<const float4*>(sk + d * SC_LD + 4 * tx);
      const float qq[4] = {qv.x, qv.y, qv.z, qv.w}, kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(qq[i], kk[j], acc[i][j]);
    }
  }
  const float inv = 1.0f / (float)A;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int s = s0 + 4 * ty + i;
    if (s >= nr) continue;
    float* o = qk + ((size_t)row_off[b] + s) * T + t0 + 4 * tx;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (t0 + 4 * tx + j < T) o[j] = acc[i][j] * inv;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// per row: w = softmax(temperature x); w = medfilt(w, width) with reflect padding; cost = -(w - mean) / (std + 1e-8)
// one CTA (256 threads) per row, the row lives in shared memory
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}
__device__ __forceinline__ void cswap(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo; b = hi;
}
constexpr int MF_MAX = 9;  // widest median filter
__global__ void __launch_bounds__(256) dtw_cost_kernel(const float* __restrict__ qk, float* __restrict__ cost, int T, float temperature,
                                                       int width) {
  __shared__ float w[DTW_T_MAX];
  __shared__ float red[8];
  const float* x = qk + (size_t)blockIdx.x * T;
  float* o = cost + (size_t)blockIdx.x * T;
  const int tid = threadIdx.x;
  float mx = -INFINITY;
  for (int t = tid; t < T; t += 256) { const float v = x[t] * temperature; w[t] = v; mx = fmaxf(mx, v); }
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int t = tid; t < T; t += 256) { const float e = expf(w[t] - mx); w[t] = e; sum += e; }
  sum = block_reduce(sum, red, false);  // its barriers also publish w[]
  const int pad = width / 2;
  // median of `width` reflect-padded neighbours (median_filter_fix.py:9-10: rows no longer than the pad are left alone)
  constexpr int PER = (DTW_T_MAX + 255) / 256;
  float med[PER];  // fixed trip counts below keep it in registers
  float s1 = 0.f;
#pragma unroll
  for (int n = 0; n < PER; ++n) {
    const int t = tid + 256 * n;
    med[n] = 0.f;
    if (t >= T) continue;
    if (T > pad && width > 1) {
      float v[MF_MAX];
#pragma unroll
      for (int k = 0; k < MF_MAX; ++k) {
        int i = t + k - pad;
        i = i < 0 ? -i : (i >= T ? 2 * (T - 1) - i : i);
        v[k] = k < width ? w[i] / sum : INFINITY;  // unused taps sort to the top
      }
      // partial selection sort up to the median position (exact order statistics, no arithmetic)
#pragma unroll
      for (int a = 0; a <= MF_MAX / 2; ++a)
#pragma unroll
        for (int c = a + 1; c < MF_MAX; ++c) cswap(v[a], v[c]);
      float m = v[0];
#pragma unroll
      for (int a = 1; a <= MF_MAX / 2; ++a)
        if (a == pad) m = v[a];
      med[n] = m;
    } else {
      med[n] = w[t] / sum;
    }
    s1 += med[n];
  }
  const float mean = block_reduce(s1, red, false) / (float)T;
  float s2 = 0.f;
#pragma unroll
  for (int n = 0; n < PER; ++n)
    if (tid + 256 * n < T) { const float dlt = med[n] - mean; s2 += dlt * dlt; }
  const float sd = sqrtf(block_reduce(s2, red, false) / (float)T) + 1e-8f;
#pragma unroll
  for (int n = 0; n < PER; ++n)
    if (tid + 256 * n < T) o[tid + 256 * n] = -((med[n] - mean) / sd);
}

// ---------------------------------------------------------------------------------------------------------------------
// DTW (OpenAI whisper/timing.py dtw_cpu + backtrace, what mlx_whisper.timing.dtw ports) on x[i, j] = cost[j][i]:
//   c[i, j] = x[i-1, j-1] + min(c0 = c[i-1, j-1], c1 = c[i-1, j], c2 = c[i, j-1]);  trace 0 / 1 / 2 with the reference's tie rules
//   (c0 < c1 && c0 < c2 -> 0; c1 < c0 && c1 < c2 -> 1; else 2), c[0, 0] = 0, the rest of row / column 0 = inf;
//   backtrace from (N, M) with trace[0, :] = 2, trace[:, 0] = 1.
// One CTA per sequence, thread j owns token column j + 1 and walks down its column one frame per anti-diagonal step.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DTW_M_MAX) dtw_path_kernel(const float* __restrict__ cost, const int* __restrict__ n_rows,
                                                             const int* __restrict__ row_off, int T, int cap,
                                                             int* __restrict__ path_frames, int* __restrict__ path_tokens,
                                                             int* __restrict__ path_len) {
  extern __shared__ uint32_t dsm[];
  const int b = blockIdx.x, M = n_rows[b], N = T;
  const int j = threadIdx.x;
  if (M <= 0) {
    if (j == 0) path_len[b] = 0;
    return;
  }
  const int W = (N >> 4) + 1;               // trace words per column: 2 bits per frame row 0 .. N
  uint32_t* trace = dsm;                     // [M][W]
  float* edge = reinterpret_cast<float*>(dsm + (size_t)DTW_M_MAX * W);  // [2][DTW_M_MAX + 1]: the column values of the last two steps
  int* pf = reinterpret_cast<int*>(edge + 2 * (DTW_M_MAX + 1));         // reversed path, cap entries each
  int* pt = pf + cap;
  const float* xj = cost + ((size_t)row_off[b] + j) * T;  // x[., j] = this token's row of the cost matrix
  float own = INFINITY;    // c[i, j+1] of the previous frame row (c[0, j+1] = inf)
  float diag = (j == 0) ? 0.f : INFINITY;  // c[i, j] of the previous frame row: c[0, 0] = 0, c[0, j > 0] = inf
  uint32_t tw = 0;
  float xnext = (j < M) ? __ldg(xj) : 0.f;  // x[i, j] is fetched one step ahead of its use
  for (int step = 0; step < N + M - 1; ++step) {
    const int i = step - j;  // 0-based frame of this thread's cell in this step
    if (j < M && i >= 0 && i < N) {
      const float xv = xnext;
      if (i + 1 < N) xnext = __ldg(xj + i + 1);
      const float left = (j == 0) ? INFINITY : edge[((step - 1) & 1) * (DTW_M_MAX + 1) + j - 1];  // c[i+1, j]: the neighbour's cell of the previous step
      const float c0 = diag, c1 = own, c2 = left;
      float c;
      uint32_t t;
      if (c0 < c1 && c0 < c2) { c = c0; t = 0u; }
      else if (c1 < c0 && c1 < c2) { c = c1; t = 1u; }
      else { c = c2; t = 2u; }
      own = xv + c;
      diag = left;  // c[i+1, j] is the diagonal neighbour of the next row
      edge[(step & 1) * (DTW_M_MAX + 1) + j] = own;
      const int r = i + 1;  // 1-based row
      tw |= t << ((r & 15) * 2);
      if ((r & 15) == 15 || r == N) { trace[(size_t)j * W + (r >> 4)] = tw; tw = 0; }
    }
    __syncthreads();
  }
  if (j == 0) {
    int i = N, jj = M, n = 0;
    while ((i > 0 || jj > 0) && n < cap) {
      pf[n] = i - 1; pt[n] = jj - 1; ++n;
      uint32_t t;
      if (i == 0) t = 2u;
      else if (jj == 0) t = 1u;
      else t = (trace[(size_t)(jj - 1) * W + (i >> 4)] >> ((i & 15) * 2)) & 3u;
      if (t == 0u) { --i; --jj; }
      else if (t == 1u) --i;
      else --jj;
    }
    path_len[b] = n;
    edge[0] = __int_as_float(n);
  }
  __syncthreads();
  const int P = __float_as_int(edge[0]);
  for (int k = j; k < P; k += blockDim.x) {
    path_frames[(size_t)b * cap + k] = pf[P - 1 - k];
    path_tokens[(size_t)b * cap + k] = pt[P - 1 - k];
  }
}

size_t dtw_path_smem(int T, int cap) {
  const int W = (T >> 4) + 1;
  return (size_t)DTW_M_MAX * W * 4 + (size_t)2 * (DTW_M_MAX + 1) * 4 + (size_t)2 * cap * 4;
}

// small int tables (n_rows, row offsets) -> device workspace; returns row total through *total
int upload_rows(wxb_ctx* ctx, const int32_t* n_rows_host, int B, int** n_rows_dev, int** row_off_dev, int* total, cudaStream_t st) {
  std::vector<int> h(2 * (size_t)B);
  int sum = 0;
  for (int b = 0; b < B; ++b) {
    if (n_rows_host[b] < 0 || n_rows_host[b] > DTW_M_MAX) return wxb_fail(ctx, WXB_ERR_INVALID, "dtw: n_rows[%d] = %d outside 0..%d", b, n_rows_host[b], DTW_M_MAX);
    h[b] = n_rows_host[b];
    h[B + b] = sum;
    sum += n_rows_host[b];
  }
  int* d = (int*)wxb_named(ctx, "dtw.rows", (size_t)2 * 4096 * 4);
  if (!d) return WXB_ERR_CUDA;
  if (B > 4096) return wxb_fail(ctx, WXB_ERR_INVALID, "dtw: at most 4096 sequences per call");
  // the table is tiny and pageable: a synchronous copy keeps the host vector's lifetime trivial
  WXB_CUDA(ctx, cudaStreamSynchronize(st));
  WXB_CUDA(ctx, cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  *n_rows_dev = d; *row_off_dev = d + B; *total = sum;
  return WXB_OK;
}

}  // namespace

extern "C" int wxb_decode_collect_heads(wxb_ctx* ctx, const int32_t* layer_head_host, int n_heads) {
  if (!ctx) return WXB_ERR_INVALID;
  if (n_heads < 0 || n_heads > 127 || (n_heads > 0 && !layer_head_host)) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decode_collect_heads: 0..127 heads");
  ctx->align_heads.assign(layer_head_host, layer_head_host + 2 * (size_t)n_heads);
  ctx->align_heads_dirty = true;
  ctx->qlog_valid = false;
  return WXB_OK;
}

extern "C" int wxb_dtw_scores(wxb_ctx* ctx, int B, int pos0, const int32_t* n_rows_host, float* qk_out_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!ctx->model) return wxb_fail(ctx, WXB_ERR_STATE, "wxb_dtw_scores: no model set");
  if (!ctx->qlog_valid || ctx->align_heads.empty())
    return wxb_fail(ctx, WXB_ERR_STATE, "wxb_dtw_scores: no decode has logged alignment-head queries (wxb_decode_collect_heads, then decode)");
  const wxb_dims& D = ctx->model->dims;
  const int A = (int)ctx->align_heads.size() / 2, T = D.n_audio_ctx, TX = D.n_text_ctx;
  if (B <= 0 || B > ctx->qlog_B0 || !n_rows_host || !qk_out_dev || pos0 < 0)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_dtw_scores: bad argument (the last decode ran %d sequences)", ctx->qlog_B0);
  cudaStream_t st = (cudaStream_t)stream;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  int max_rows = 0;
  for (int b = 0; b < B; ++b) {
    if (n_rows_host[b] > 0 && pos0 + n_rows_host[b] > ctx->qlog_pos) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_dtw_scores: sequence %d asks for positions up to %d, the decode logged %d", b, pos0 + n_rows_host[b], ctx->qlog_pos);
    max_rows = n_rows_host[b] > max_rows ? n_rows_host[b] : max_rows;
  }
  int *nr, *ro, total, rc;
  if ((rc = upload_rows(ctx, n_rows_host, B, &nr, &ro, &total, st)) != WXB_OK) return rc;
  if (max_rows == 0) return WXB_OK;
  const float* qlog = (const float*)wxb_named(ctx, "dec.qlog", 1);
  const int* heads = (const int*)wxb_named(ctx, "dec.qheads", 1);
  const __nv_bfloat16* ckv = (const __nv_bfloat16*)wxb_named(ctx, "dec.cross_kv", 1);
  if (!qlog || !heads || !ckv) return WXB_ERR_STATE;
  dim3 grid(ceil_div(T, SC_TILE), ceil_div(max_rows, SC_TILE), B);
  dtw_scores_kernel<<<grid, 256, 0, st>>>(qlog, ckv, heads, nr, ro, qk_out_dev, A, ctx->qlog_B0, D.n_text_head, TX, T, pos0);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}

extern "C" int wxb_dtw_cost(wxb_ctx* ctx, const float* qk_dev, int64_t rows, int T, float temperature, int medfilt_width,
                            float* cost_out_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!qk_dev || !cost_out_dev || rows < 0 || T <= 0 || T > DTW_T_MAX || medfilt_width < 1 || medfilt_width > MF_MAX || !(medfilt_width & 1))
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_dtw_cost: bad argument (T <= %d, odd filter width <= %d)", DTW_T_MAX, MF_MAX);
  if (rows == 0) return WXB_OK;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  dtw_cost_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(qk_dev, cost_out_dev, T, temperature, medfilt_width);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}

extern "C" int wxb_dtw_path(wxb_ctx* ctx, const float* cost_dev, int B, const int32_t* n_rows_host, int T, int32_t* path_frames_dev,
                            int32_t* path_tokens_dev, int32_t* path_len_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!cost_dev || B <= 0 || !n_rows_host || T <= 0 || T > DTW_T_MAX || !path_frames_dev || !path_tokens_dev || !path_len_dev)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_dtw_path: bad argument (T <= %d)", DTW_T_MAX);
  cudaStream_t st = (cudaStream_t)stream;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  int *nr, *ro, total, rc;
  if ((rc = upload_rows(ctx, n_rows_host, B, &nr, &ro, &total, st)) != WXB_OK) return rc;
  const int cap = T + DTW_M_MAX;
  const size_t smem = dtw_path_smem(T, cap);
  if ((rc = wxb_func_smem(ctx, dtw_path_kernel, (int)smem)) != WXB_OK) return rc;
  dtw_path_kernel<<<B, DTW_M_MAX, smem, st>>>(cost_dev, nr, ro, T, cap, path_frames_dev, path_tokens_dev, path_len_dev);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}

extern "C" int wxb_dtw_path_capacity(int T) { return T + DTW_M_MAX; }
