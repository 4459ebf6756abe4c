// wxb_attn.cu — encoder self-attention (non-causal, head_dim 64, T = 1500) on the 5th-gen tensor cores.
//
// One CTA = one (128-query tile, head, chunk).  Per 128-key tile j:
//   S_j = Q K_j^T          tcgen05.mma M=128 N=128 K=64  (Q, K_j: TMA 128B-swizzled K-major tiles)  -> TMEM
//   P_j = exp2(S_j c - m_j c)   softmax warps: one thread per query row reads its S row with tcgen05.ld,
//                               keeps the running max / sum in registers, writes P_j (bf16) into shared
//                               memory in the canonical K-major swizzled layout (A operand of the next MMA)
//   O_j = P_j V_j          tcgen05.mma M=128 N=64 K=128 (V^T tiles from a pre-transposed copy)     -> TMEM
//   o   = (o + O_{j-1}) * exp2((m_{j-1} - m_j) c)   accumulated in registers (no TMEM read-modify-write)
// S and O are double-buffered in TMEM and K/V/P in shared memory, so the tensor core computes S_{j+1} and
// O_{j-1} while the softmax warps work on tile j.  Warp roles: 0 = TMA producer, 1 = MMA issuer,
// 2 = TMEM allocator, 4..7 = softmax / output.
#include "wxb_common.cuh"
#include "wxb_tc.cuh"
#include <math.h>

using namespace wxbtc;

namespace {

constexpr int BQ = 128, BKV = 128;
constexpr int AT_THREADS = 256;
constexpr int Q_BYTES = BQ * 64 * 2;          // 16 KB
constexpr int K_BYTES = BKV * 64 * 2;         // 16 KB
constexpr int VB_BYTES = 64 * 64 * 2;         // one [64 dims x 64 keys] K-block of V^T, 8 KB
constexpr int V_BYTES = 2 * VB_BYTES;         // 16 KB
constexpr int PB_BYTES = BQ * 64 * 2;         // one [128 q x 64 keys] K-block of P, 16 KB
constexpr int P_BYTES = 2 * PB_BYTES;         // 32 KB
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;
constexpr int OFF_V = OFF_K + 2 * K_BYTES;
constexpr int OFF_P = OFF_V + 2 * V_BYTES;
constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;  // 147456
constexpr int AT_SMEM = OFF_BAR + 16 * 8 + 16 + 1024;
constexpr int TM_S = 0, TM_O = 256;           // TMEM columns: S[2] at 0/128, O[2] at 256/320

struct AttnTcParams {
  int T, Tpad, d, H, n_kv_tiles;
  float scale_log2;
  __nv_bfloat16* out;
};

// V part of qkv [B*T, 3d] -> vT [(b*H + h)*64 + j][Tpad] (keys contiguous), zero in the T..Tpad-1 padding
__global__ void __launch_bounds__(256)
v_transpose_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ vT, int T, int Tpad, int d, int H) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int t0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < 64 * 32; idx += 256) {
    const int r = idx >> 5, c2 = idx & 31;  // row = time, c2 = pair of dims
    const int t = t0 + r;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (t < T) v = *reinterpret_cast<const __nv_bfloat162*>(qkv + ((size_t)b * T + t) * 3 * d + 2 * d + h * 64 + 2 * c2);
    tile[r][2 * c2] = __low2bfloat16(v);
    tile[r][2 * c2 + 1] = __high2bfloat16(v);
  }
  __syncthreads();
  for (int idx = tid; idx < 64 * 32; idx += 256) {
    const int j = idx >> 5, t2 = idx & 31;  // row = dim, t2 = pair of times
    const int t = t0 + 2 * t2;
    if (t < Tpad) {
      __nv_bfloat162 v;
      v.x = tile[2 * t2][j];
      v.y = tile[2 * t2 + 1][j];
      *reinterpret_cast<__nv_bfloat162*>(vT + ((size_t)(b * H + h) * 64 + j) * Tpad + t) = v;
    }
  }
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;     // [2]
  uint64_t* s_empty = bars + 7;    // [2]
  uint64_t* p_full = bars + 9;     // [2]
  uint64_t* o_full = bars + 11;    // [2]
  uint64_t* o_empty = bars + 13;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int n = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQK)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
      mbar_init(s_full + i, 1);
      mbar_init(s_empty + i, 4);
      mbar_init(p_full + i, 4);
      mbar_init(o_full + i, 1);
      mbar_init(o_empty + i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int row0 = b * p.T;
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(smem + OFF_Q, &tmQK, q_full, h * 64, row0 + q0);
      for (int j = 0; j < n; ++j) {
        const int st = j & 1;
        mbar_wait(kv_empty + st, ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full + st, K_BYTES + V_BYTES);
        tma_load_2d(smem + OFF_K + st * K_BYTES, &tmQK, kv_full + st, p.d + h * 64, row0 + j * BKV);
        tma_load_2d(smem + OFF_V + st * V_BYTES, &tmV, kv_full + st, j * BKV, (b * p.H + h) * 64);
        tma_load_2d(smem + OFF_V + st * V_BYTES + VB_BYTES, &tmV, kv_full + st, j * BKV + 64, (b * p.H + h) * 64);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idescS = make_idesc_bf16(128, 128), idescO = make_idesc_bf16(128, 64);
      const uint32_t sQ = smem_u32(smem + OFF_Q);
      auto issue_S = [&](int j) {
        const int st = j & 1;
        mbar_wait(kv_full + st, (j >> 1) & 1);
        mbar_wait(s_empty + st, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t adesc = make_sw128_desc(sQ);
        const uint64_t bdesc = make_sw128_desc(smem_u32(smem + OFF_K + st * K_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(tmem_base + TM_S + st * 128, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idescS, k != 0);
        tc_commit(s_full + st);
      };
      mbar_wait(q_full, 0);
      issue_S(0);
      for (int j = 0; j < n; ++j) {
        const int st = j & 1;
        if (j + 1 < n) issue_S(j + 1);
        mbar_wait(p_full + st, (j >> 1) & 1);
        mbar_wait(o_empty + st, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kb2 = 0; kb2 < 2; ++kb2) {
          const uint64_t adesc = make_sw128_desc(smem_u32(smem + OFF_P + st * P_BYTES + kb2 * PB_BYTES));
          const uint64_t bdesc = make_sw128_desc(smem_u32(smem + OFF_V + st * V_BYTES + kb2 * VB_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_base + TM_O + st * 64, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idescO, (kb2 | k) != 0);
        }
        tc_commit(o_full + st);
        tc_commit(kv_empty + st);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output
    const int ew = warp - 4;
    const int r = ew * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const float sl = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;

    for (int j = 0; j < n; ++j) {
      const int st = j & 1;
      const int valid = p.T - j * BKV;  // keys of this tile that exist (>= 128 except for the last tile)
      mbar_wait(s_full + st, (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_addr + TM_S + st * 128;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tc_ld_32x32(s_addr + c * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float f = (c * 32 + i < valid) ? __uint_as_float(v[i]) : -INFINITY;
          mx = fmaxf(mx, f);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * sl);
      const float off = m_new * sl;
      float rowsum = 0.f;
      uint8_t* prow = smem + OFF_P + st * P_BYTES + r * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tc_ld_32x32(s_addr + c * 32, v);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = (c * 32 + i < valid) ? exp2f(__uint_as_float(v[i]) * sl - off) : 0.f;
          const float p1 = (c * 32 + i + 1 < valid) ? exp2f(__uint_as_float(v[i + 1]) * sl - off) : 0.f;
          rowsum += p0 + p1;
          __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);
        }
        uint8_t* blk = prow + (c >> 1) * PB_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = (c & 1) * 4 + q;
          *reinterpret_cast<uint4*>(blk + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      // S_j fully read, P_j written: release the S buffer, publish P to the tensor core (async proxy)
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_empty + st);
        mbar_arrive(p_full + st);
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      if (j > 0) {
        const int sp = (j - 1) & 1;
        mbar_wait(o_full + sp, ((j - 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t o_addr = tmem_base + lane_addr + TM_O + sp * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tc_ld_32x32(o_addr + c * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = (o[c * 32 + i] + __uint_as_float(v[i])) * alpha;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty + sp);
      }
    }
    {
      const int sp = (n - 1) & 1;
      mbar_wait(o_full + sp, ((n - 1) >> 1) & 1);
      tc_fence_after();
      const uint32_t o_addr = tmem_base + lane_addr + TM_O + sp * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tc_ld_32x32(o_addr + c * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(v[i]);
      }
    }
    const int q = q0 + r;
    if (q < p.T) {
      const float inv = 1.f / l_run;
      __nv_bfloat16* dst = p.out + ((size_t)b * p.T + q) * p.d + h * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        uint4 pk;
        __nv_bfloat162 b0 = __floats2bfloat162_rn(o[i] * inv, o[i + 1] * inv);
        __nv_bfloat162 b1 = __floats2bfloat162_rn(o[i + 2] * inv, o[i + 3] * inv);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(o[i + 4] * inv, o[i + 5] * inv);
        __nv_bfloat162 b3 = __floats2bfloat162_rn(o[i + 6] * inv, o[i + 7] * inv);
        pk.x = *reinterpret_cast<uint32_t*>(&b0);
        pk.y = *reinterpret_cast<uint32_t*>(&b1);
        pk.z = *reinterpret_cast<uint32_t*>(&b2);
        pk.w = *reinterpret_cast<uint32_t*>(&b3);
        *reinterpret_cast<uint4*>(dst + i) = pk;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int wxb_make_tmap_bf16(wxb_ctx* ctx, CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_rows);

// qkv bf16 [B*T, 3d] -> out bf16 [B*T, d]; vT = scratch bf16 [B*H*64, Tpad]
int wxb_attention_tc(wxb_ctx* ctx, const __nv_bfloat16* qkv, __nv_bfloat16* vT, __nv_bfloat16* out, int B, int T, int d, int H,
                     cudaStream_t st) {
  const int Tpad = (T + 7) & ~7;
  v_transpose_kernel<<<dim3(ceil_div(Tpad, 64), H, B), 256, 0, st>>>(qkv, vT, T, Tpad, d, H);
  WXB_LAUNCH_CHECK(ctx);
  CUtensorMap tmQK, tmV;
  int rc;
  if ((rc = wxb_make_tmap_bf16(ctx, &tmQK, qkv, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d * 2, 64, 128)) != WXB_OK) return rc;
  if ((rc = wxb_make_tmap_bf16(ctx, &tmV, vT, (uint64_t)Tpad, (uint64_t)B * H * 64, (uint64_t)Tpad * 2, 64, 64)) != WXB_OK) return rc;
  static bool attr = false;
  if (!attr) {
    WXB_CUDA(ctx, cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr = true;
  }
  AttnTcParams p;
  p.T = T; p.Tpad = Tpad; p.d = d; p.H = H; p.n_kv_tiles = ceil_div(T, BKV);
  p.scale_log2 = (1.0f / sqrtf(64.f)) * 1.44269504088896341f;
  p.out = out;
  attention_tc_kernel<<<dim3(ceil_div(T, BQ), H, B), AT_THREADS, AT_SMEM, st>>>(tmQK, tmV, p);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}
