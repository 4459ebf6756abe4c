// wxb_decoder.cu — K3: batched greedy KV-cache decoder (Whisper TextDecoder) for VAD-cut chunks.
//
// Behavioural spec: the reference's in-tree batched loop
//   /root/reference/mlx_whisper_batch_decoder.py:317-384 (_main_loop_batch), :267-303 (update),
//   :386-468 (run: EOT trimming, avg_logprob), filters per SURVEY A.3 (SuppressBlank, SuppressTokens).
//
// One decode step is bandwidth-bound (weights once per step + cross-KV once per sequence), so every
// kernel here is built around streaming HBM:
//   dec_gemv_kernel    y[b,n] = sum_k act[b,k] W[n,k]: mma.sync m16n8k16 with the BATCH as the M tile
//                      (<=16 rows, padded) and 8 weight rows as the N tile; weights are read straight
//                      from HBM into MMA B-fragments with 16-byte loads (a K-permutation shared by the
//                      A and B fragments makes natural row-major weights fragment-ready, no repack, no
//                      shared-memory staging); the first weight loads are issued BEFORE the
//                      programmatic-dependent-launch wait so they overlap the previous kernel's tail.
//                      LayerNorm of the residual stream is fused into the activation staging, and
//                      bias / GELU / residual add / QKV scatter into the KV cache into the epilogue.
//   dec_attn_kernel    one CTA per (split, head, sequence): 8 lanes per key (16-byte loads, 128 B per
//                      key row = fully coalesced), scores -> softmax -> P.V in fp32, split-KV partials
//                      merged by the last-arriving CTA (self-cleaning ticket).
//   dec_sample_kernel  logit filters + argmax (first max) + logsumexp + bookkeeping, one CTA per row.
// The per-step launch sequence is captured once in a CUDA graph (position read from device memory).
//
// HBM layout (L decoder layers, B sequences, H heads, d = 64 H):
//   self K/V  bf16 [L][2][B][H][448][64]      cross K/V bf16 [L][2][B][H][1500][64]
//   x f32 [B,d] residual; q f32 [B,d]; att f32 [B,d]; hid bf16 [B,4d]; logits f32 [B,V]
#include "wxb_gemm.cuh"
#include "wxb_model.cuh"
#include <math.h>
#include <stdlib.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

int wxb_launch_layernorm(wxb_ctx* ctx, const float* x, const float* w, const float* b, __nv_bfloat16* y, long long rows,
                         int d, cudaStream_t st);

namespace {

constexpr int T_AUDIO = 1500;
constexpr int GV_THREADS = 256;
constexpr int GV_U = 4;  // 64-wide K chunks per register group

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void mma_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

enum { EPI_F32 = 0, EPI_RESID = 1, EPI_GELU_BF16 = 2, EPI_QKV = 3 };

constexpr int GV_ROWS = 64;    // weight rows per CTA (8 warps x 8 rows)
constexpr int GV_KS_MAX = 8;   // split-K CTAs of a row block form one thread-block cluster (portable limit 8)

struct GemvParams {
  int B, N, K;
  int ks;                      // split-K factor across CTAs (gridDim.y); K / ks is a multiple of 64
  const __nv_bfloat16* in;     // activations bf16 [B, K]
  long long ld_in;
  const __nv_bfloat16* W;
  const float* bias;
  int epi;
  void* out;
  long long ldo;
  // EPI_QKV
  float* q_out;
  __nv_bfloat16 *kcache, *vcache;  // [B][H][tmax][64] of this layer
  const int* d_pos;
  int H, tmax;
};

// Weight chunk c (64 K-columns) of this lane: two 16-byte loads at k = 64c + 8t and 64c + 32 + 8t of row (n0 + g).
__device__ __forceinline__ void gv_load(uint4* w, const __nv_bfloat16* wrow, int c_first, int c_end) {
#pragma unroll
  for (int u = 0; u < GV_U; ++u) {
    const int c = c_first + u;
    if (c < c_end) {
      w[2 * u] = ldg_nc_v4(wrow + (size_t)c * 64);
      w[2 * u + 1] = ldg_nc_v4(wrow + (size_t)c * 64 + 32);
    }
  }
}
__device__ __forceinline__ void gv_compute(float* acc, const uint4* w, const unsigned char* act_lo, const unsigned char* act_hi,
                                           int c_first, int c_end) {
#pragma unroll
  for (int u = 0; u < GV_U; ++u) {
    const int c = c_first + u;
    if (c < c_end) {
      const uint4 al0 = *reinterpret_cast<const uint4*>(act_lo + c * 128);
      const uint4 ah0 = *reinterpret_cast<const uint4*>(act_hi + c * 128);
      const uint4 al1 = *reinterpret_cast<const uint4*>(act_lo + c * 128 + 64);
      const uint4 ah1 = *reinterpret_cast<const uint4*>(act_hi + c * 128 + 64);
      const uint4 w0 = w[2 * u], w1 = w[2 * u + 1];
      mma_16816(acc, al0.x, ah0.x, al0.y, ah0.y, w0.x, w0.y);
      mma_16816(acc, al0.z, ah0.z, al0.w, ah0.w, w0.z, w0.w);
      mma_16816(acc, al1.x, ah1.x, al1.y, ah1.y, w1.x, w1.y);
      mma_16816(acc, al1.z, ah1.z, al1.w, ah1.w, w1.z, w1.w);
    }
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}

// LayerNorm (eps 1e-5) of the residual stream rows: f32 [B, d] -> bf16 [B, d], one warp per row (d <= 1280)
__global__ void __launch_bounds__(256)
dec_ln_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              __nv_bfloat16* __restrict__ y, int B, int d) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int nv = d >> 2;
  float4 ww[10], bb[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) {  // parameters do not depend on the previous kernel
    const int idx = lane + 32 * i;
    if (idx < nv) { ww[i] = __ldg(reinterpret_cast<const float4*>(w) + idx); bb[i] = __ldg(reinterpret_cast<const float4*>(bias) + idx); }
  }
  pdl_wait();
  pdl_launch();
  if (row >= B) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * d);
  float4 v[10];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) { v[i] = xr[idx]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
  }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float a = v[i].x - mean, b2 = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += a * a + b2 * b2 + c * c + e * e;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
  uint2* yr = reinterpret_cast<uint2*>(y + (size_t)row * d);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      uint2 pk;
      pk.x = pack_bf16((v[i].x - mean) * rstd * ww[i].x + bb[i].x, (v[i].y - mean) * rstd * ww[i].y + bb[i].y);
      pk.y = pack_bf16((v[i].z - mean) * rstd * ww[i].z + bb[i].z, (v[i].w - mean) * rstd * ww[i].w + bb[i].w);
      yr[idx] = pk;
    }
  }
}

// Skinny GEMM y[B, N] = act[B, K] W[N, K]^T for B <= 64.  CTA = 64 weight rows x one K-slice; warp w owns
// rows 8w..8w+7 and keeps its weight fragments in registers while looping over the (<= 4) 16-row batch
// tiles, whose activations are cp.async'ed into shared memory.  gridDim.y K-slices are combined
// deterministically: every slice publishes fp32 partials and the last-arriving CTA of a row block sums
// them in slice order and runs the epilogue (bias / GELU / residual / QKV scatter into the KV cache).
__global__ void __launch_bounds__(GV_THREADS, 2)
dec_gemv_kernel(const GemvParams p) {
  extern __shared__ __align__(16) unsigned char gv_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Bp = (p.B + 15) & ~15;
  const int MT = Bp >> 4;
  const int kslice = p.K / p.ks;
  const int k_begin = blockIdx.y * kslice;
  const int chunks = kslice >> 6;
  const int nblk0 = blockIdx.x * GV_ROWS;
  const int n0 = nblk0 + warp * 8;
  const size_t row_bytes = (size_t)(kslice + 32) * 2;  // +64 B: rows g and g+1 land on different bank halves

  int nrow = n0 + g;
  if (nrow >= p.N) nrow = p.N - 1;
  const __nv_bfloat16* wrow = p.W + (size_t)nrow * p.K + k_begin + 8 * t;
  uint4 wa[2 * GV_U], wb[2 * GV_U];
  gv_load(wa, wrow, 0, chunks);  // weights do not depend on the previous kernel: request them before the wait
  pdl_wait();

  // ---- stage the activation tile (B batch rows x K-slice, bf16) with 16-byte async copies ----
  {
    const int nv = kslice >> 3;
    for (int idx = tid; idx < Bp * nv; idx += GV_THREADS) {
      const int r = idx / nv, c = idx - r * nv;
      unsigned char* dst = gv_smem + r * row_bytes + c * 16;
      if (r < p.B) cp_async16(dst, p.in + (size_t)r * p.ld_in + k_begin + c * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  pdl_launch();

  float acc[4][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;
  {
    const unsigned char* act_lo = gv_smem + g * row_bytes + 16 * t;
    const unsigned char* act_hi = act_lo + 8 * row_bytes;
    for (int c = 0; c < chunks; c += 2 * GV_U) {
      if (c + GV_U < chunks) gv_load(wb, wrow, c + GV_U, chunks);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
        if (mt < MT) gv_compute(acc[mt], wa, act_lo + mt * 16 * row_bytes, act_hi + mt * 16 * row_bytes, c, chunks);
      if (c + 2 * GV_U < chunks) gv_load(wa, wrow, c + 2 * GV_U, chunks);
      if (c + GV_U < chunks) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
          if (mt < MT) gv_compute(acc[mt], wb, act_lo + mt * 16 * row_bytes, act_hi + mt * 16 * row_bytes, c + GV_U, chunks);
      }
    }
  }
  // ---- every K-slice CTA parks its 64-column fp32 tile in its own shared memory; the ks CTAs of a row block
  //      form a thread-block cluster and each of them finishes a slice of the batch rows, summing the ks
  //      tiles in rank order through distributed shared memory (deterministic, no global partials) ----
  float* s_out = reinterpret_cast<float*>(gv_smem);  // [Bp][66] overlays the activation tile
  __syncthreads();  // everyone is done reading the activation tile
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
    if (mt < MT) {
      const int col = warp * 8 + 2 * t;
      *reinterpret_cast<float2*>(s_out + (mt * 16 + g) * 66 + col) = make_float2(acc[mt][0], acc[mt][1]);
      *reinterpret_cast<float2*>(s_out + (mt * 16 + g + 8) * 66 + col) = make_float2(acc[mt][2], acc[mt][3]);
    }
  cg::cluster_group cluster = cg::this_cluster();
  int rank = 0;
  if (p.ks > 1) {
    cluster.sync();
    rank = (int)cluster.block_rank();
  } else {
    __syncthreads();
  }
  const float* tiles[8];
#pragma unroll
  for (int s2 = 0; s2 < 8; ++s2) tiles[s2] = (p.ks > 1 && s2 < p.ks) ? cluster.map_shared_rank(s_out, s2) : s_out;
  int pos = 0;
  if (p.epi == EPI_QKV) pos = *p.d_pos;
  const int n = nblk0 + 2 * lane;
  const bool ok0 = n < p.N, ok1 = n + 1 < p.N;
  const float b0 = (p.bias && ok0) ? __ldg(p.bias + n) : 0.f;
  const float b1 = (p.bias && ok1) ? __ldg(p.bias + n + 1) : 0.f;
  const int rows_per = (p.B + p.ks - 1) / p.ks;
  const int row_end = min(p.B, (rank + 1) * rows_per);
  for (int b = rank * rows_per + warp; b < row_end; b += 8) {
    float2 pv[8];
#pragma unroll
    for (int s2 = 0; s2 < 8; ++s2)
      if (s2 < p.ks) pv[s2] = *reinterpret_cast<const float2*>(tiles[s2] + b * 66 + 2 * lane);
    float v0 = b0, v1 = b1;
#pragma unroll
    for (int s2 = 0; s2 < 8; ++s2)
      if (s2 < p.ks) { v0 += pv[s2].x; v1 += pv[s2].y; }
    if (p.epi == EPI_F32) {
      float* o = reinterpret_cast<float*>(p.out) + (size_t)b * p.ldo + n;
      if (ok0) o[0] = v0;
      if (ok1) o[1] = v1;
    } else if (p.epi == EPI_RESID) {
      float* o = reinterpret_cast<float*>(p.out) + (size_t)b * p.ldo + n;
      if (ok1) {
        float2 cur = *reinterpret_cast<float2*>(o);
        cur.x += v0; cur.y += v1;
        *reinterpret_cast<float2*>(o) = cur;
      } else if (ok0) {
        o[0] += v0;
      }
    } else if (p.epi == EPI_GELU_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)b * p.ldo + n;
      if (ok1) *reinterpret_cast<uint32_t*>(o) = pack_bf16(gelu_erf(v0), gelu_erf(v1));
      else if (ok0) o[0] = __float2bfloat16_rn(gelu_erf(v0));
    } else {
      const int d = p.N / 3;  // n and n+1 share a 64-wide head block (n even)
      if (ok0) {
        if (n < d) {
          *reinterpret_cast<float2*>(p.q_out + (size_t)b * d + n) = make_float2(v0, v1);
        } else {
          const int nn = (n < 2 * d) ? (n - d) : (n - 2 * d);
          __nv_bfloat16* cache = (n < 2 * d) ? p.kcache : p.vcache;
          const int h = nn >> 6, j = nn & 63;
          *reinterpret_cast<uint32_t*>(cache + (((size_t)b * p.H + h) * p.tmax + pos) * 64 + j) = pack_bf16(v0, v1);
        }
      }
    }
  }
  if (p.ks > 1) cluster.sync();  // peers may still be reading this CTA's tile
}

// ---------------------------------------------------------------------------------------------
// decode attention: q f32 [B,d] against K/V bf16 [B][H][tkv][64]
// ---------------------------------------------------------------------------------------------
constexpr int DA_THREADS = 256;

struct AttnParams {
  const float* q;               // [B, d]
  const __nv_bfloat16 *K, *V;   // [B][H][tkv][64]
  int tkv;                      // allocated keys per (b,h)
  int n_keys;                   // used when d_pos == nullptr
  const int* d_pos;             // if set: n_keys = *d_pos + 1 (self-attention)
  int splits, H, d, B;
  float scale;
  __nv_bfloat16* out;           // [B, d] bf16 (feeds the out-projection GEMV)
  float* part;                  // [B][H][splits][66]  (m, l, o[64])
  int* ticket;                  // [B*H], zero-initialised, self-cleaning
};

// Single pass over the keys with an online softmax per 8-lane key slot: every lane owns 8 of the 64
// dims of its slot's keys; K and V rows of 4 keys per slot are requested together and one iteration
// ahead (<= 16 x 16 B in flight per lane); the 32 slot states of the CTA are merged through shared
// memory at the end.  static_kv (cross-attention): K/V do not depend on the previous kernel, so the
// first loads are issued before the programmatic-dependent-launch wait.
__device__ __forceinline__ void da_load(uint4* kv, uint4* vv, const __nv_bfloat16* Kb, const __nv_bfloat16* Vb, int kb, int slot,
                                        int c8, int k1) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int key = kb + u * 32 + slot;
    if (key < k1) {
      kv[u] = ldg_nc_v4(Kb + (size_t)key * 64 + c8 * 8);
      vv[u] = ldg_nc_v4(Vb + (size_t)key * 64 + c8 * 8);
    }
  }
}

__global__ void __launch_bounds__(DA_THREADS)
dec_attn_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float da_smem[];
  float* sq = da_smem;            // [64] scaled q
  float* sst = sq + 64;           // [32 slots][66]: m, l, acc[64]
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slot = lane >> 3, c8 = lane & 7;
  const bool static_kv = (p.d_pos == nullptr);
  const int n_units = p.splits * p.H * p.B;
  bool first = true;
  // persistent over work units (split, head, sequence): the cross-attention launch uses one CTA per SM so
  // that the other batch group's GEMV chain can co-reside and overlap with this HBM stream
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
  const int split = unit % p.splits, h = (unit / p.splits) % p.H, b = unit / (p.splits * p.H);
  const size_t slab = ((size_t)b * p.H + h) * p.tkv * 64;
  const __nv_bfloat16* Kb = p.K + slab;
  const __nv_bfloat16* Vb = p.V + slab;
  uint4 kA[4], vA[4], kB[4], vB[4];
  int n_keys = p.n_keys, per = 0, k0 = 0, k1 = 0;
  if (static_kv) {
    per = (n_keys + p.splits - 1) / p.splits;
    k0 = split * per;
    k1 = min(n_keys, k0 + per);
    da_load(kA, vA, Kb, Vb, k0 + warp * 4, slot, c8, k1);
  }
  if (first) pdl_wait();
  if (!static_kv) {
    n_keys = *p.d_pos + 1;
    per = (n_keys + p.splits - 1) / p.splits;
    k0 = split * per;
    k1 = min(n_keys, k0 + per);
    da_load(kA, vA, Kb, Vb, k0 + warp * 4, slot, c8, k1);
  }
  if (tid < 64) sq[tid] = p.q[(size_t)b * p.d + h * 64 + tid] * p.scale;
  __syncthreads();
  if (first) pdl_launch();
  first = false;
  float qr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) qr[j] = sq[c8 * 8 + j];

  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;

  auto consume = [&](const uint4* kv, const uint4* vv, int kb) {
    float sc4[4];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = kb + u * 32 + slot;
      float sdot = 0.f;
      if (key < k1) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          sdot = fmaf(qr[2 * j], f.x, sdot);
          sdot = fmaf(qr[2 * j + 1], f.y, sdot);
        }
      }
      sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
      sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
      sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
      sc4[u] = (key < k1) ? sdot : -INFINITY;
      mx = fmaxf(mx, sc4[u]);
    }
    if (mx > -INFINITY) {  // uniform within the 8-lane slot
      const float mn = fmaxf(m, mx);
      const float alpha = __expf(m - mn);  // m = -inf -> 0
      m = mn;
      l *= alpha;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= alpha;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float pr = __expf(sc4[u] - mn);  // -inf -> 0
        l += pr;
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vv[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          acc[2 * j] = fmaf(pr, f.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(pr, f.y, acc[2 * j + 1]);
        }
      }
    }
  };
  // masked-out V registers may hold garbage (never loaded): zero them so 0 * garbage cannot make NaN
  auto clear = [&](uint4* vv, int kb) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (kb + u * 32 + slot >= k1) vv[u] = make_uint4(0u, 0u, 0u, 0u);
  };
  for (int kb = k0 + warp * 4; kb < k1; kb += 2 * 128) {
    if (kb + 128 < k1) da_load(kB, vB, Kb, Vb, kb + 128, slot, c8, k1);
    clear(vA, kb);
    consume(kA, vA, kb);
    if (kb + 256 < k1) da_load(kA, vA, Kb, Vb, kb + 256, slot, c8, k1);
    if (kb + 128 < k1) {
      clear(vB, kb + 128);
      consume(kB, vB, kb + 128);
    }
  }
  // ---- merge the 32 slot states ----
  {
    float* st = sst + (warp * 4 + slot) * 66;
    if (c8 == 0) { st[0] = m; st[1] = l; }
#pragma unroll
    for (int j = 0; j < 8; ++j) st[2 + c8 * 8 + j] = acc[j];
  }
  __syncthreads();
  float o = 0.f, L = 0.f, M = -INFINITY;
  if (tid < 64) {
#pragma unroll 4
    for (int i = 0; i < 32; ++i) M = fmaxf(M, sst[i * 66]);
    if (M > -INFINITY) {
      for (int i = 0; i < 32; ++i) {
        const float w = __expf(sst[i * 66] - M);
        L += w * sst[i * 66 + 1];
        o += w * sst[i * 66 + 2 + tid];
      }
    }
  }
  if (p.splits == 1) {
    if (tid < 64) p.out[(size_t)b * p.d + h * 64 + tid] = __float2bfloat16_rn(o / L);
  } else {
    // ---- split-KV: publish the partial, the last CTA of this (b,h) merges ----
    float* part = p.part + (((size_t)b * p.H + h) * p.splits) * 66;
    if (tid < 64) {
      part[split * 66 + 2 + tid] = o;
      if (tid == 0) { part[split * 66] = M; part[split * 66 + 1] = L; }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int prev = atomicAdd(p.ticket + b * p.H + h, 1);
      s_last = (prev == p.splits - 1);
      if (s_last) p.ticket[b * p.H + h] = 0;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (tid < 64) {
        float MM = -INFINITY;
        for (int s2 = 0; s2 < p.splits; ++s2) MM = fmaxf(MM, __ldcg(part + s2 * 66));
        float LL = 0.f, OO = 0.f;
        for (int s2 = 0; s2 < p.splits; ++s2) {
          const float w = __expf(__ldcg(part + s2 * 66) - MM);
          LL += w * __ldcg(part + s2 * 66 + 1);
          OO += w * __ldcg(part + s2 * 66 + 2 + tid);
        }
        p.out[(size_t)b * p.d + h * 64 + tid] = __float2bfloat16_rn(OO / LL);
      }
    }
  }
  __syncthreads();  // sq / sst are reused by the next unit
  }  // unit loop
}

// Causal self-attention over the <= 448 cached positions: one WARP per (sequence, head), no shared memory,
// no block barriers.  Same lane layout as dec_attn_kernel (4 key slots x 8 lanes x 8 dims); the 4 slot
// states are merged with shuffles at the end.
__global__ void __launch_bounds__(256)
dec_self_attn_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ K, const __nv_bfloat16* __restrict__ V,
                     const int* __restrict__ d_pos, int tkv, int H, int d, int n_bh, float scale, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x * 8 + warp;
  if (bh >= n_bh) return;
  const int b = bh / H, h = bh - b * H;
  const int slot = lane >> 3, c8 = lane & 7;
  const int n_keys = *d_pos + 1;
  const __nv_bfloat16* Kb = K + (size_t)bh * tkv * 64;
  const __nv_bfloat16* Vb = V + (size_t)bh * tkv * 64;
  float qr[8];
  {
    const float4 a = *reinterpret_cast<const float4*>(q + (size_t)b * d + h * 64 + c8 * 8);
    const float4 c = *reinterpret_cast<const float4*>(q + (size_t)b * d + h * 64 + c8 * 8 + 4);
    qr[0] = a.x * scale; qr[1] = a.y * scale; qr[2] = a.z * scale; qr[3] = a.w * scale;
    qr[4] = c.x * scale; qr[5] = c.y * scale; qr[6] = c.z * scale; qr[7] = c.w * scale;
  }
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int kb = 0; kb < n_keys; kb += 16) {
    uint4 kv[4], vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = kb + u * 4 + slot;
      vv[u] = make_uint4(0u, 0u, 0u, 0u);
      if (key < n_keys) {
        kv[u] = *reinterpret_cast<const uint4*>(Kb + (size_t)key * 64 + c8 * 8);
        vv[u] = *reinterpret_cast<const uint4*>(Vb + (size_t)key * 64 + c8 * 8);
      }
    }
    float sc4[4], mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int key = kb + u * 4 + slot;
      float sdot = 0.f;
      if (key < n_keys) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          sdot = fmaf(qr[2 * j], f.x, sdot);
          sdot = fmaf(qr[2 * j + 1], f.y, sdot);
        }
      }
      sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
      sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
      sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
      sc4[u] = (key < n_keys) ? sdot : -INFINITY;
      mx = fmaxf(mx, sc4[u]);
    }
    if (mx > -INFINITY) {
      const float mn = fmaxf(m, mx);
      const float alpha = __expf(m - mn);
      m = mn;
      l *= alpha;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= alpha;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float pr = __expf(sc4[u] - mn);
        l += pr;
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&vv[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          acc[2 * j] = fmaf(pr, f.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(pr, f.y, acc[2 * j + 1]);
        }
      }
    }
  }
  // merge the 4 slot states (lanes differing in bits 3 and 4)
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float mn = fmaxf(m, m2);
    const float w1 = (m > -INFINITY) ? __expf(m - mn) : 0.f;
    const float w2 = (m2 > -INFINITY) ? __expf(m2 - mn) : 0.f;
    l = l * w1 + l2 * w2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a2 = __shfl_xor_sync(0xffffffffu, acc[j], o);
      acc[j] = acc[j] * w1 + a2 * w2;
    }
    m = mn;
  }
  if (slot == 0) {
    const float inv = 1.f / l;
    uint4 pk;
    pk.x = pack_bf16(acc[0] * inv, acc[1] * inv);
    pk.y = pack_bf16(acc[2] * inv, acc[3] * inv);
    pk.z = pack_bf16(acc[4] * inv, acc[5] * inv);
    pk.w = pack_bf16(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(out + (size_t)b * d + h * 64 + c8 * 8) = pk;
  }
}

// ---------------------------------------------------------------------------------------------
// token embedding + learned position; sampling; bookkeeping
// ---------------------------------------------------------------------------------------------
// x[b,:] = emb[tok[b*stride + pos]] + pos_emb[pos], pos = *d_pos
__global__ void __launch_bounds__(256)
dec_embed_kernel(const int* __restrict__ tok, int stride, const int* __restrict__ d_pos,
                 const __nv_bfloat16* __restrict__ emb, const float* __restrict__ pos_emb,
                 float* __restrict__ x, int d, int n_vocab) {
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.x;
  const int pos = *d_pos;
  int token = tok[(size_t)b * stride + pos];
  token = min(max(token, 0), n_vocab - 1);
  for (int i = 2 * threadIdx.x; i < d; i += 2 * blockDim.x) {
    const __nv_bfloat162 e = *reinterpret_cast<const __nv_bfloat162*>(emb + (size_t)token * d + i);
    const float2 pe = *reinterpret_cast<const float2*>(pos_emb + (size_t)pos * d + i);
    *reinterpret_cast<float2*>(x + (size_t)b * d + i) = make_float2(__low2float(e) + pe.x, __high2float(e) + pe.y);
  }
}

__global__ void dec_advance_kernel(int* d_pos) {
  pdl_wait();
  if (threadIdx.x == 0) *d_pos += 1;
}

// softmax probability of `token` per row (no_speech_prob at the SOT position, unfiltered)
__global__ void dec_token_prob_kernel(const float* __restrict__ logits, int V, int token, float* __restrict__ out) {
  __shared__ float red[32];
  pdl_wait();
  const float* x = logits + (size_t)blockIdx.x * V;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) m = fmaxf(m, x[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(x[i] - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    out[blockIdx.x] = expf(x[token] - m) / t;
  }
}

struct SampleParams {
  float* logits;  // [B, V] (filters are applied in place)
  int V;
  int* tokens;    // [B, stride]: prompt + sampled tokens; this step writes column pos + 1
  int stride;
  int* d_pos;
  int prompt_len;
  int eot, suppress_blank, blank_token, n_suppress;
  const int* suppress;
  float* sum_logprob;  // [B]
  int* done;           // [B] 1 once the row has emitted EOT
};

// mlx_whisper_batch_decoder.py:267-303 for one row: filters, argmax, logprob accounting, EOT latch.
__global__ void __launch_bounds__(1024)
dec_sample_kernel(const SampleParams p) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  float* x = p.logits + (size_t)b * p.V;
  const int pos = *p.d_pos;
  for (int i = tid; i < p.n_suppress; i += blockDim.x) {
    const int id = p.suppress[i];
    if (id >= 0 && id < p.V) x[id] = -INFINITY;
  }
  if (p.suppress_blank && pos == p.prompt_len - 1 && tid == 0) {
    if (p.blank_token >= 0 && p.blank_token < p.V) x[p.blank_token] = -INFINITY;
    x[p.eot] = -INFINITY;
  }
  __syncthreads();
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = tid; i < p.V; i += blockDim.x) {
    const float v = x[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
  __syncthreads();
  best = s_val[0]; bi = s_idx[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (s_val[w] > best || (s_val[w] == best && s_idx[w] < bi)) { best = s_val[w]; bi = s_idx[w]; }
  __syncthreads();
  float s = 0.f;
  for (int i = tid; i < p.V; i += blockDim.x) s += expf(x[i] - best);
  s = warp_sum(s);
  if ((tid & 31) == 0) s_val[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_val[w];
    const float logprob = -logf(tot);  // x[bi] - (best + log(sum)) with x[bi] == best
    int* row = p.tokens + (size_t)b * p.stride;
    const int last = row[pos];
    const bool was_eot = (last == p.eot) && (pos >= p.prompt_len);  // prompt tokens never latch
    if (!was_eot) p.sum_logprob[b] += logprob;
    const int next = was_eot ? p.eot : bi;
    row[pos + 1] = next;
    if (next == p.eot) p.done[b] = 1;
  }
  if (b == 0 && tid == 0) {
    // every row of this step has read d_pos before any block can be this far? No: blocks are
    // independent, so the position is advanced by dec_advance_kernel in the next launch.
  }
}

// n_tokens[b] = sampled tokens before the first EOT; tokens_out[b, i] = sampled token i (EOT padded)
__global__ void dec_finalize_kernel(const int* __restrict__ tokens, int stride, int prompt_len, int n_sampled, int sample_len,
                                    int eot, int* __restrict__ tokens_out, int* __restrict__ n_tokens) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int n = n_sampled;
    for (int i = 0; i < n_sampled; ++i)
      if (tokens[(size_t)b * stride + prompt_len + i] == eot) { n = i; break; }
    n_tokens[b] = n;
  }
  for (int i = threadIdx.x; i < sample_len; i += blockDim.x)
    tokens_out[(size_t)b * sample_len + i] = (i < n_sampled) ? tokens[(size_t)b * stride + prompt_len + i] : eot;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct DecBuffers {
  float *x, *q, *logits, *part, *sum_lp, *gv_part;
  __nv_bfloat16 *att, *xn, *hid, *self_kv, *cross_kv;
  int *ticket, *gv_ticket, *d_pos, *tokens, *done;
  int B, tok_stride;
  size_t gv_part_floats;
};

// profiling aid only (results become meaningless): WXB_DEC_SKIP bitmask 1 = cross-attention, 2 = GEMV + LN, 4 = self-attention
int dec_skip_mask() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXB_DEC_SKIP");
    v = e ? atoi(e) : 0;
  }
  return v;
}

bool use_pdl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXB_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int g_cluster_y = 1;   // cluster dimension (y) of the next launch_k call
int g_low_prio = 0;    // 1: the next launch_k call is the bandwidth stream (cross-attention) -> lowest priority

template <typename... KArgs, typename... Args>
int launch_k(wxb_ctx* ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int na = 0;
  {  // chain kernels outrank the cross-attention stream of the other batch group
    attr[na].id = cudaLaunchAttributePriority;
    static int use_prio = -1;
    if (use_prio < 0) { const char* e = getenv("WXB_PRIO"); use_prio = e ? atoi(e) : 1; }
    attr[na].val.priority = (g_low_prio || !use_prio) ? 0 : -1;
    ++na;
    g_low_prio = 0;
  }
  if (use_pdl()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (g_cluster_y > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = g_cluster_y;
    attr[na].val.clusterDim.z = 1;
    ++na;
    g_cluster_y = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
  ctx->launches++;
  if (e != cudaSuccess) return wxb_fail(ctx, WXB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return WXB_OK;
}

constexpr size_t GV_SMEM_MAX = 200 * 1024;

size_t gemv_smem(int Bp, int kslice) {
  const size_t act = (size_t)Bp * (kslice + 32) * 2;
  const size_t outt = (size_t)Bp * 66 * 4;
  return act > outt ? act : outt;
}

// split-K factor = cluster size in {1,2,4,8}: the smallest that yields about one CTA per SM while the
// activation tile stays <= ~100 KB (two CTAs per SM); very wide N (logits) never splits.
int pick_ks(wxb_ctx* ctx, int Bp, int N, int K) {
  const int row_blocks = ceil_div(N, GV_ROWS), kc = K / 64;
  int best = 0;
  for (int ks = 1; ks <= GV_KS_MAX; ks *= 2) {
    if (kc % ks) break;
    const size_t sm = gemv_smem(Bp, K / ks);
    if (sm > GV_SMEM_MAX) continue;
    best = ks;
    if (row_blocks * ks >= ctx->sm_count && (sm <= 100 * 1024 || row_blocks >= 2 * ctx->sm_count)) break;
  }
  return best;
}

int launch_gemv(wxb_ctx* ctx, GemvParams p, const DecBuffers& buf, cudaStream_t st) {
  (void)buf;
  if (dec_skip_mask() & 2) return WXB_OK;
  if (p.K % 64) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "gemv: K=%d must be a multiple of 64", p.K);
  if (p.B > 64) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "gemv: batch %d > 64", p.B);
  const int Bp = (p.B + 15) & ~15;
  p.ks = pick_ks(ctx, Bp, p.N, p.K);
  if (p.ks == 0) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "gemv: no K split of K=%d fits shared memory", p.K);
  g_cluster_y = p.ks;
  return launch_k(ctx, dec_gemv_kernel, dim3(ceil_div(p.N, GV_ROWS), p.ks), dim3(GV_THREADS), gemv_smem(Bp, p.K / p.ks), st, p);
}

int launch_ln(wxb_ctx* ctx, const float* x, const float* w, const float* b, __nv_bfloat16* y, int B, int d, cudaStream_t st) {
  if (dec_skip_mask() & 2) return WXB_OK;
  if (d % 4 || d > 1280) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder layernorm: d=%d", d);
  return launch_k(ctx, dec_ln_kernel, dim3(ceil_div(B, 8)), dim3(256), 0, st, x, w, b, y, B, d);
}

int launch_attn(wxb_ctx* ctx, AttnParams p, int B, cudaStream_t st) {
  if (dec_skip_mask() & (p.d_pos ? 4 : 1)) return WXB_OK;
  const size_t smem = (size_t)(64 + 32 * 66) * 4;
  p.B = B;
  const int n_units = p.splits * p.H * B;
  int grid = n_units;
  if (!p.d_pos) {  // cross-attention = the bandwidth stream: lowest priority so the other group's chain kernels slip in
    static int persist = -1;
    if (persist < 0) { const char* e = getenv("WXB_ATTN_PERSIST"); persist = e ? atoi(e) : 0; }
    if (persist > 0) grid = n_units < persist * ctx->sm_count ? n_units : persist * ctx->sm_count;
    g_low_prio = 1;
  }
  return launch_k(ctx, dec_attn_kernel, dim3(grid), dim3(DA_THREADS), smem, st, p);
}

int alloc_buffers(wxb_ctx* ctx, int B, int tok_stride, int group, DecBuffers* o) {
  const std::string sfx = group ? (".g" + std::to_string(group)) : std::string();
  auto nm = [&](const char* base) { return std::string(base) + sfx; };
  const wxb_dims& D = ctx->model->dims;
  const int d = D.n_text_state, L = D.n_text_layer, H = D.n_text_head, V = D.n_vocab;
  o->B = B;
  o->tok_stride = tok_stride;
  if (B > 64) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder: batch %d > 64 sequences per call", B);
  // opt in to the largest dynamic shared memory the GEMV may ask for; set outside graph capture
  WXB_CUDA(ctx, cudaFuncSetAttribute(dec_gemv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GV_SMEM_MAX));
  o->x = (float*)wxb_named(ctx, nm("dec.x").c_str(), (size_t)B * d * 4);
  o->q = (float*)wxb_named(ctx, nm("dec.q").c_str(), (size_t)B * d * 4);
  o->att = (__nv_bfloat16*)wxb_named(ctx, nm("dec.att").c_str(), (size_t)B * d * 2);
  o->xn = (__nv_bfloat16*)wxb_named(ctx, nm("dec.xn").c_str(), (size_t)B * d * 2);
  o->gv_part_floats = (size_t)4 << 20;  // 16 MB of fp32 split-K partials
  o->gv_part = (float*)wxb_named(ctx, nm("dec.gv_part").c_str(), o->gv_part_floats * 4);
  o->gv_ticket = (int*)wxb_named(ctx, nm("dec.gv_ticket").c_str(), (size_t)(ceil_div(V, GV_ROWS) + 64) * 4, true);
  o->hid = (__nv_bfloat16*)wxb_named(ctx, nm("dec.hid").c_str(), (size_t)B * 4 * d * 2);
  o->logits = (float*)wxb_named(ctx, nm("dec.logits").c_str(), (size_t)B * V * 4);
  o->part = (float*)wxb_named(ctx, nm("dec.part").c_str(), (size_t)B * H * 8 * 66 * 4);
  o->sum_lp = (float*)wxb_named(ctx, nm("dec.sum_lp").c_str(), (size_t)B * 4);
  o->self_kv = (__nv_bfloat16*)wxb_named(ctx, nm("dec.self_kv").c_str(), (size_t)L * 2 * B * H * D.n_text_ctx * 64 * 2);
  o->cross_kv = (__nv_bfloat16*)wxb_named(ctx, nm("dec.cross_kv").c_str(), (size_t)L * 2 * B * H * T_AUDIO * 64 * 2);
  o->ticket = (int*)wxb_named(ctx, nm("dec.ticket").c_str(), (size_t)B * H * 4 + 64, true);
  o->d_pos = (int*)wxb_named(ctx, nm("dec.pos").c_str(), 64);
  o->tokens = (int*)wxb_named(ctx, nm("dec.tokens").c_str(), (size_t)B * tok_stride * 4);
  o->done = (int*)wxb_named(ctx, nm("dec.done").c_str(), (size_t)B * 4);
  if (!o->x || !o->q || !o->att || !o->hid || !o->logits || !o->part || !o->sum_lp || !o->self_kv || !o->cross_kv ||
      !o->ticket || !o->d_pos || !o->tokens || !o->done || !o->xn || !o->gv_part || !o->gv_ticket)
    return WXB_ERR_CUDA;
  return WXB_OK;
}

// cross K/V of every decoder layer from the encoder output (tcgen05 GEMM, head-major scatter)
int cross_kv_precompute(wxb_ctx* ctx, const __nv_bfloat16* enc_out, const DecBuffers& buf, cudaStream_t st) {
  const wxb_dims& D = ctx->model->dims;
  const int d = D.n_text_state, H = D.n_text_head, B = buf.B;
  for (int l = 0; l < D.n_text_layer; ++l) {
    DecLayerW w;
    int rc = wxb_dec_layer(ctx, l, &w);
    if (rc != WXB_OK) return rc;
    GemmArgs a;
    a.A = enc_out; a.lda = D.n_audio_state; a.M = B * T_AUDIO; a.W = w.ckv_w; a.N = 2 * d; a.K = D.n_audio_state;
    a.bias = w.ckv_b;
    a.out = buf.cross_kv + (size_t)l * 2 * B * H * T_AUDIO * 64;
    a.ldo = 2 * d; a.kv_mode = 1; a.kv_B = B; a.kv_H = H; a.kv_T = T_AUDIO;
    if ((rc = wxb_gemm_launch(ctx, a, st)) != WXB_OK) return rc;
  }
  return WXB_OK;
}

// One decoder step at position *d_pos over buf.tokens[:, pos]; logits (optional) to logits_out with row stride ldl.
int decoder_step(wxb_ctx* ctx, const DecBuffers& buf, float* logits_out, long long ldl, cudaStream_t st) {
  const wxb_dims& D = ctx->model->dims;
  const int d = D.n_text_state, H = D.n_text_head, B = buf.B, L = D.n_text_layer, TX = D.n_text_ctx;
  const __nv_bfloat16* emb = (const __nv_bfloat16*)wxb_weight(ctx, "dec.emb");
  const float* pos_emb = (const float*)wxb_weight(ctx, "dec.pos");
  const float* lnf_w = (const float*)wxb_weight(ctx, "dec.ln.w");
  const float* lnf_b = (const float*)wxb_weight(ctx, "dec.ln.b");
  if (!emb || !pos_emb || !lnf_w || !lnf_b) return WXB_ERR_STATE;
  int rc;
  if ((rc = launch_k(ctx, dec_embed_kernel, dim3(B), dim3(256), 0, st, (const int*)buf.tokens, buf.tok_stride,
                     (const int*)buf.d_pos, emb, pos_emb, buf.x, d, D.n_vocab)) != WXB_OK)
    return rc;
  const float scale = 1.0f / sqrtf(64.f);
  const int cross_splits = (B * H >= 4 * ctx->sm_count) ? 1 : ((B * H >= 2 * ctx->sm_count) ? 2 : 4);
  for (int l = 0; l < L; ++l) {
    DecLayerW w;
    if ((rc = wxb_dec_layer(ctx, l, &w)) != WXB_OK) return rc;
    __nv_bfloat16* sk = buf.self_kv + (size_t)l * 2 * B * H * TX * 64;
    __nv_bfloat16* sv = sk + (size_t)B * H * TX * 64;
    const __nv_bfloat16* ck = buf.cross_kv + (size_t)l * 2 * B * H * T_AUDIO * 64;
    const __nv_bfloat16* cv = ck + (size_t)B * H * T_AUDIO * 64;
    GemvParams g = {};
    g.B = B;
    // 1. LN1 + fused QKV, K/V appended to the self cache at pos
    if ((rc = launch_ln(ctx, buf.x, w.ln1_w, w.ln1_b, buf.xn, B, d, st)) != WXB_OK) return rc;
    g.N = 3 * d; g.K = d; g.in = buf.xn; g.ld_in = d;
    g.W = w.qkv_w; g.bias = w.qkv_b; g.epi = EPI_QKV; g.q_out = buf.q; g.kcache = sk; g.vcache = sv;
    g.d_pos = buf.d_pos; g.H = H; g.tmax = TX;
    if ((rc = launch_gemv(ctx, g, buf, st)) != WXB_OK) return rc;
    // 2. causal self-attention over pos+1 cached positions (one warp per (b, h))
    if (!(dec_skip_mask() & 4)) {
      if ((rc = launch_k(ctx, dec_self_attn_kernel, dim3(ceil_div(B * H, 8)), dim3(256), 0, st, (const float*)buf.q,
                         (const __nv_bfloat16*)sk, (const __nv_bfloat16*)sv, (const int*)buf.d_pos, TX, H, d, B * H, scale, buf.att)) != WXB_OK)
        return rc;
    }
    AttnParams a = {};
    // 3. out projection + residual
    g = GemvParams{};
    g.B = B; g.N = d; g.K = d; g.in = buf.att; g.ld_in = d; g.W = w.out_w; g.bias = w.out_b;
    g.epi = EPI_RESID; g.out = buf.x; g.ldo = d;
    if ((rc = launch_gemv(ctx, g, buf, st)) != WXB_OK) return rc;
    // 4. LN2 + cross query
    g = GemvParams{};
    if ((rc = launch_ln(ctx, buf.x, w.ln2_w, w.ln2_b, buf.xn, B, d, st)) != WXB_OK) return rc;
    g.B = B; g.N = d; g.K = d; g.in = buf.xn; g.ld_in = d;
    g.W = w.cq_w; g.bias = w.cq_b; g.epi = EPI_F32; g.out = buf.q; g.ldo = d;
    if ((rc = launch_gemv(ctx, g, buf, st)) != WXB_OK) return rc;
    // 5. cross-attention over the 1500 encoder positions
    a = AttnParams{};
    a.q = buf.q; a.K = ck; a.V = cv; a.tkv = T_AUDIO; a.n_keys = T_AUDIO; a.d_pos = nullptr; a.splits = cross_splits;
    a.H = H; a.d = d; a.scale = scale; a.out = buf.att; a.part = buf.part; a.ticket = buf.ticket;
    if ((rc = launch_attn(ctx, a, B, st)) != WXB_OK) return rc;
    // 6. cross out projection + residual
    g = GemvParams{};
    g.B = B; g.N = d; g.K = d; g.in = buf.att; g.ld_in = d; g.W = w.cout_w; g.bias = w.cout_b;
    g.epi = EPI_RESID; g.out = buf.x; g.ldo = d;
    if ((rc = launch_gemv(ctx, g, buf, st)) != WXB_OK) return rc;
    // 7. LN3 + fc1 + GELU
    g = GemvParams{};
    if ((rc = launch_ln(ctx, buf.x, w.ln3_w, w.ln3_b, buf.xn, B, d, st)) != WXB_OK) return rc;
    g.B = B; g.N = 4 * d; g.K = d; g.in = buf.xn; g.ld_in = d;
    g.W = w.fc1_w; g.bias = w.fc1_b; g.epi = EPI_GELU_BF16; g.out = buf.hid; g.ldo = 4 * d;
    if ((rc = launch_gemv(ctx, g, buf, st)) != WXB_OK) return rc;
    // 8. fc2 + residual
    g = GemvParams{};
    g.B = B; g.N = d; g.K = 4 * d; g.in = buf.hid; g.ld_in = 4 * d; g.W = w.fc2_w; g.bias = w.fc2_b;
    g.epi = EPI_RESID; g.out = buf.x; g.ldo = d;
    if ((rc = launch_gemv(ctx, g, buf, st)) != WXB_OK) return rc;
  }
  if (logits_out) {
    GemvParams g = {};
    if ((rc = launch_ln(ctx, buf.x, lnf_w, lnf_b, buf.xn, B, d, st)) != WXB_OK) return rc;
    g.B = B; g.N = D.n_vocab; g.K = d; g.in = buf.xn; g.ld_in = d;
    g.W = emb; g.bias = nullptr; g.epi = EPI_F32; g.out = logits_out; g.ldo = ldl;
    if ((rc = launch_gemv(ctx, g, buf, st)) != WXB_OK) return rc;
  }
  return WXB_OK;
}

struct StepGraph {
  cudaGraphExec_t exec = nullptr;
  // identity of what was captured
  const void* model = nullptr;
  int B = 0, tok_stride = 0, mode = 0;
  void* key_ptrs[4] = {nullptr, nullptr, nullptr, nullptr};
  SampleParams sp = {};
};
StepGraph g_graphs[2][3];  // [batch group][0: prefill (no logits), 1: prefill + logits (no sampling), 2: logits + sample]

bool use_graph() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXB_GRAPH");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int enqueue_step(wxb_ctx* ctx, const DecBuffers& buf, int mode, const SampleParams& sp, cudaStream_t st) {
  int rc;
  if ((rc = decoder_step(ctx, buf, mode >= 1 ? buf.logits : nullptr, ctx->model->dims.n_vocab, st)) != WXB_OK) return rc;
  if (mode == 2) {
    if ((rc = launch_k(ctx, dec_sample_kernel, dim3(buf.B), dim3(1024), 0, st, sp)) != WXB_OK) return rc;
  }
  return launch_k(ctx, dec_advance_kernel, dim3(1), dim3(32), 0, st, buf.d_pos);
}

// Run one step of `mode`, through a cached CUDA graph when enabled.
int run_step(wxb_ctx* ctx, const DecBuffers& buf, int mode, const SampleParams& sp, cudaStream_t st, int group) {
  if (!use_graph()) return enqueue_step(ctx, buf, mode, sp, st);
  StepGraph& G = g_graphs[group][mode];
  const bool same = G.exec && G.model == (const void*)ctx->model && G.B == buf.B && G.tok_stride == buf.tok_stride &&
                    G.key_ptrs[0] == buf.x && G.key_ptrs[1] == buf.self_kv && G.key_ptrs[2] == buf.cross_kv &&
                    G.key_ptrs[3] == buf.tokens && memcmp(&G.sp, &sp, sizeof(sp)) == 0;
  if (!same) {
    if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
    const int64_t launches_before = ctx->launches;
    cudaGraph_t graph = nullptr;
    // capture on a private stream (the caller's stream may be the legacy default stream, which
    // cannot be captured); the instantiated graph is then launched on the caller's stream
    if (!ctx->cap_stream) {
      cudaStream_t cs;
      WXB_CUDA(ctx, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      ctx->cap_stream = cs;
    }
    cudaStream_t cs = (cudaStream_t)ctx->cap_stream;
    WXB_CUDA(ctx, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_step(ctx, buf, mode, sp, cs);
    cudaError_t e = cudaStreamEndCapture(cs, &graph);
    ctx->launches = launches_before;
    if (rc != WXB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return wxb_fail(ctx, WXB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&G.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { G.exec = nullptr; return wxb_fail(ctx, WXB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e)); }
    G.model = ctx->model; G.B = buf.B; G.tok_stride = buf.tok_stride; G.mode = mode;
    G.key_ptrs[0] = buf.x; G.key_ptrs[1] = buf.self_kv; G.key_ptrs[2] = buf.cross_kv; G.key_ptrs[3] = buf.tokens;
    G.sp = sp;
  }
  WXB_CUDA(ctx, cudaGraphLaunch(G.exec, st));
  const int L = ctx->model->dims.n_text_layer;
  ctx->launches += 1 + 11 * L + (mode >= 1 ? 2 : 0) + (mode == 2 ? 1 : 0) + 1;
  return WXB_OK;
}

}  // namespace

void wxb_decoder_reset_graphs() {
  for (auto& row : g_graphs)
    for (auto& G : row)
      if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
}

extern "C" int wxb_decode_greedy(wxb_ctx* ctx, const void* enc_out_dev, int B, const int32_t* prompt_host, int prompt_len,
                                 const wxb_decode_opts* opts, int32_t* tokens_out_dev, int32_t* n_tokens_dev,
                                 float* sum_logprob_dev, float* no_speech_prob_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!ctx->model) return wxb_fail(ctx, WXB_ERR_STATE, "wxb_decode_greedy: no model set");
  if (!enc_out_dev || B <= 0 || !prompt_host || prompt_len <= 0 || !opts || !tokens_out_dev || !n_tokens_dev || !sum_logprob_dev)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decode_greedy: bad argument");
  const wxb_dims& D = ctx->model->dims;
  const int sample_len = opts->sample_len;
  if (sample_len <= 0 || prompt_len + sample_len > D.n_text_ctx)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decode_greedy: prompt_len %d + sample_len %d exceeds n_text_ctx %d", prompt_len,
                    sample_len, D.n_text_ctx);
  if (opts->eot < 0 || opts->eot >= D.n_vocab) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decode_greedy: eot out of range");
  cudaStream_t st = (cudaStream_t)stream;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int stride = D.n_text_ctx + 1;
  int rc;
  // Batch groups: with >= 32 sequences the batch is decoded as two independent halves on two private
  // streams (own buffers, own step graphs), so one half's bandwidth-bound cross-attention overlaps the
  // other half's latency-bound GEMV chain; the second reader of a weight matrix hits L2.
  int ng = (B >= 32) ? 2 : 1;
  if (const char* e = getenv("WXB_DEC_GROUPS")) ng = (atoi(e) >= 2 && B >= 2) ? 2 : 1;
  if (B > 64 * ng) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_decode_greedy: at most %d sequences per call", 64 * ng);
  cudaStream_t sg[2] = {st, st};
  if (ng == 2) {
    for (int g = 0; g < 2; ++g) {
      if (!ctx->dec_streams[g]) {
        cudaStream_t s2;
        WXB_CUDA(ctx, cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        ctx->dec_streams[g] = s2;
      }
      sg[g] = (cudaStream_t)ctx->dec_streams[g];
    }
    if (!ctx->dec_events[0])
      for (int i = 0; i < 3; ++i) {
        cudaEvent_t ev;
        WXB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->dec_events[i] = ev;
      }
  }
  int g0[3] = {0, (ng == 2) ? (B + 1) / 2 : B, B};
  if (ng == 1) g0[2] = B;
  DecBuffers buf[2];
  SampleParams sp[2];
  for (int g = 0; g < ng; ++g) {
    const int Bg = g0[g + 1] - g0[g];
    if ((rc = alloc_buffers(ctx, Bg, stride, g, &buf[g])) != WXB_OK) return rc;
  }
  wxb_dec_timing tm;
  WXB_CUDA(ctx, cudaEventCreate(&tm.e0));
  WXB_CUDA(ctx, cudaEventCreate(&tm.e1));
  WXB_CUDA(ctx, cudaEventCreate(&tm.e2));
  WXB_CUDA(ctx, cudaEventRecord(tm.e0, st));
  if (ng == 2) {
    WXB_CUDA(ctx, cudaEventRecord((cudaEvent_t)ctx->dec_events[0], st));
    for (int g = 0; g < 2; ++g) WXB_CUDA(ctx, cudaStreamWaitEvent(sg[g], (cudaEvent_t)ctx->dec_events[0], 0));
  }
  std::vector<int> init((size_t)B * stride, opts->eot);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < prompt_len; ++i) init[(size_t)b * stride + i] = prompt_host[i];
  const size_t enc_row = (size_t)T_AUDIO * D.n_audio_state;
  for (int g = 0; g < ng; ++g) {
    const int Bg = buf[g].B;
    // tokens[b, :prompt_len] = prompt; state reset
    WXB_CUDA(ctx, cudaMemcpyAsync(buf[g].tokens, init.data() + (size_t)g0[g] * stride, (size_t)Bg * stride * 4, cudaMemcpyHostToDevice, sg[g]));
    WXB_CUDA(ctx, cudaMemsetAsync(buf[g].d_pos, 0, 4, sg[g]));
    WXB_CUDA(ctx, cudaMemsetAsync(buf[g].done, 0, (size_t)Bg * 4, sg[g]));
    WXB_CUDA(ctx, cudaMemsetAsync(buf[g].sum_lp, 0, (size_t)Bg * 4, sg[g]));
    if ((rc = cross_kv_precompute(ctx, (const __nv_bfloat16*)enc_out_dev + (size_t)g0[g] * enc_row, buf[g], sg[g])) != WXB_OK) return rc;
    SampleParams& s1 = sp[g];
    s1 = SampleParams{};
    s1.logits = buf[g].logits; s1.V = D.n_vocab; s1.tokens = buf[g].tokens; s1.stride = stride; s1.d_pos = buf[g].d_pos;
    s1.prompt_len = prompt_len; s1.eot = opts->eot; s1.suppress_blank = opts->suppress_blank; s1.blank_token = opts->blank_token;
    s1.n_suppress = opts->n_suppress; s1.suppress = opts->suppress_dev; s1.sum_logprob = buf[g].sum_lp; s1.done = buf[g].done;
  }
  for (int g = 0; g < ng; ++g) WXB_CUDA(ctx, cudaStreamSynchronize(sg[g]));  // `init` is pageable host memory
  WXB_CUDA(ctx, cudaEventRecord(tm.e1, sg[0]));

  // prompt positions 0 .. prompt_len-2 (forced tokens); logits only at position 0 for no_speech_prob
  for (int pos = 0; pos < prompt_len - 1; ++pos) {
    const bool want_nsp = (pos == 0 && opts->no_speech >= 0 && no_speech_prob_dev);
    for (int g = 0; g < ng; ++g) {
      if ((rc = run_step(ctx, buf[g], want_nsp ? 1 : 0, sp[g], sg[g], g)) != WXB_OK) return rc;
      if (want_nsp) {
        dec_token_prob_kernel<<<buf[g].B, 1024, 0, sg[g]>>>(buf[g].logits, D.n_vocab, opts->no_speech, no_speech_prob_dev + g0[g]);
        WXB_LAUNCH_CHECK(ctx);
      }
    }
  }
  const bool nsp_at_last = (prompt_len == 1 && opts->no_speech >= 0 && no_speech_prob_dev);
  const int check_every = opts->check_every > 0 ? opts->check_every : 16;
  std::vector<int> done_host(B);
  int n_sampled = 0;
  for (int i = 0; i < sample_len; ++i) {
    for (int g = 0; g < ng; ++g) {
      if (i == 0 && nsp_at_last) {
        // single-token prompt: the SOT position is also the first sampling position
        if ((rc = decoder_step(ctx, buf[g], buf[g].logits, D.n_vocab, sg[g])) != WXB_OK) return rc;
        dec_token_prob_kernel<<<buf[g].B, 1024, 0, sg[g]>>>(buf[g].logits, D.n_vocab, opts->no_speech, no_speech_prob_dev + g0[g]);
        WXB_LAUNCH_CHECK(ctx);
        if ((rc = launch_k(ctx, dec_sample_kernel, dim3(buf[g].B), dim3(1024), 0, sg[g], sp[g])) != WXB_OK) return rc;
        if ((rc = launch_k(ctx, dec_advance_kernel, dim3(1), dim3(32), 0, sg[g], buf[g].d_pos)) != WXB_OK) return rc;
      } else {
        if ((rc = run_step(ctx, buf[g], 2, sp[g], sg[g], g)) != WXB_OK) return rc;
      }
    }
    n_sampled = i + 1;
    if ((i + 1) % check_every == 0 && i + 1 < sample_len) {
      for (int g = 0; g < ng; ++g)
        WXB_CUDA(ctx, cudaMemcpyAsync(done_host.data() + g0[g], buf[g].done, (size_t)buf[g].B * 4, cudaMemcpyDeviceToHost, sg[g]));
      for (int g = 0; g < ng; ++g) WXB_CUDA(ctx, cudaStreamSynchronize(sg[g]));
      bool all = true;
      for (int b = 0; b < B; ++b) all = all && done_host[b];
      if (all) break;  // mlx_whisper_batch_decoder.py:357
    }
  }
  for (int g = 0; g < ng; ++g) {
    dec_finalize_kernel<<<buf[g].B, 256, 0, sg[g]>>>(buf[g].tokens, stride, prompt_len, n_sampled, sample_len, opts->eot,
                                                   tokens_out_dev + (size_t)g0[g] * sample_len, n_tokens_dev + g0[g]);
    WXB_LAUNCH_CHECK(ctx);
    WXB_CUDA(ctx, cudaMemcpyAsync(sum_logprob_dev + g0[g], buf[g].sum_lp, (size_t)buf[g].B * 4, cudaMemcpyDeviceToDevice, sg[g]));
    if (ng == 2) {
      WXB_CUDA(ctx, cudaEventRecord((cudaEvent_t)ctx->dec_events[1 + g], sg[g]));
      WXB_CUDA(ctx, cudaStreamWaitEvent(st, (cudaEvent_t)ctx->dec_events[1 + g], 0));
    }
  }
  WXB_CUDA(ctx, cudaEventRecord(tm.e2, st));
  tm.steps = prompt_len - 1 + n_sampled;
  ctx->dec_timings.push_back(tm);
  return WXB_OK;
}

extern "C" int wxb_decode_stats(wxb_ctx* ctx, double* cross_kv_ms, double* steps_ms, int64_t* n_steps, int reset) {
  if (!ctx) return WXB_ERR_INVALID;
  double a = 0, b = 0;
  int64_t n = 0;
  for (auto& t : ctx->dec_timings) {
    WXB_CUDA(ctx, cudaEventSynchronize(t.e2));
    float x = 0.f, y = 0.f;
    WXB_CUDA(ctx, cudaEventElapsedTime(&x, t.e0, t.e1));
    WXB_CUDA(ctx, cudaEventElapsedTime(&y, t.e1, t.e2));
    a += x; b += y; n += t.steps;
  }
  if (cross_kv_ms) *cross_kv_ms = a;
  if (steps_ms) *steps_ms = b;
  if (n_steps) *n_steps = n;
  if (reset) {
    for (auto& t : ctx->dec_timings) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); cudaEventDestroy(t.e2); }
    ctx->dec_timings.clear();
  }
  return WXB_OK;
}

extern "C" int wxb_decoder_logits(wxb_ctx* ctx, const void* enc_out_dev, int B, const int32_t* tokens_host, int n_tok,
                                  float* logits_out_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!ctx->model) return wxb_fail(ctx, WXB_ERR_STATE, "wxb_decoder_logits: no model set");
  const wxb_dims& D = ctx->model->dims;
  if (!enc_out_dev || B <= 0 || !tokens_host || n_tok <= 0 || n_tok > D.n_text_ctx || !logits_out_dev)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decoder_logits: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  DecBuffers buf;
  int rc;
  if ((rc = alloc_buffers(ctx, B, n_tok, 0, &buf)) != WXB_OK) return rc;
  WXB_CUDA(ctx, cudaMemcpyAsync(buf.tokens, tokens_host, (size_t)B * n_tok * 4, cudaMemcpyHostToDevice, st));
  WXB_CUDA(ctx, cudaMemsetAsync(buf.d_pos, 0, 4, st));
  if ((rc = cross_kv_precompute(ctx, (const __nv_bfloat16*)enc_out_dev, buf, st)) != WXB_OK) return rc;
  for (int pos = 0; pos < n_tok; ++pos) {
    if ((rc = decoder_step(ctx, buf, logits_out_dev + (size_t)pos * D.n_vocab, (long long)n_tok * D.n_vocab, st)) != WXB_OK) return rc;
    if ((rc = launch_k(ctx, dec_advance_kernel, dim3(1), dim3(32), 0, st, buf.d_pos)) != WXB_OK) return rc;
  }
  return WXB_OK;
}
