// wxb_gemm.cu — K2 building block: persistent, warp-specialised bf16 GEMM on the 5th-gen tensor
// cores.  D[M,N] = A[M,K] * W[N,K]^T, fp32 accumulation in TMEM, fused epilogue.
//
//   warp 0      TMA producer: cp.async.bulk.tensor 2D loads of a 128 x 64 A box and a BN x 64 W box
//               (128-byte swizzle) into a STAGES-deep shared-memory ring, mbarrier complete_tx.
//   warp 1      MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16)
//               four times per stage; tcgen05.commit releases the stage / publishes the accumulator.
//   warp 2      TMEM allocator (2 x BN columns: double-buffered accumulator so the epilogue of
//               tile i overlaps the main loop of tile i+1).
//   warps 4-11  epilogue: two warps per TMEM lane quarter, each half of the tile's columns; tcgen05.ld (32 lanes x
//               32 columns per instruction, the next chunk's load in flight) -> bias, GELU(erf), residual /
//               positional add, row remap (conv stem) -> bf16 or f32 global stores.
// Grid = min(#tiles, #SMs); tiles are walked m-fastest so concurrently running CTAs share W tiles.
// PAIR variant (default for large problems): clusters of two CTAs issue ONE tcgen05.mma.cta_group::2 of M = 256 over two
// m-adjacent tiles of the same n block.  Each CTA stages its own 128 A rows and only HALF of the W tile (the
// tensor cores of both SMs read both halves), so a k block costs 32 KB of shared-memory fill and 8 KB of operand reads
// per MMA and SM instead of 48 KB and 12 KB: with single-CTA 128 x 256 tiles the shared-memory port (TMA fill +
// operand reads ~ 132 B/clk at full rate), not the tensor pipe, is the limit.  The leader CTA (rank 0) issues the MMAs;
// both CTAs' TMA loads complete on the leader's full barrier, tcgen05.commit multicasts onto both CTAs' barriers, and
// the peer's epilogue warps release the accumulator on the leader's barrier through the cluster window.
#include "wxb_gemm.cuh"
#include "wxb_tc.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 384;
constexpr int EPI_WARPS = 8;

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  const float* bias;
  const float* residual;
  int res_mode;
  long long ldr;
  int gelu;
  void* out;
  int out_f32;
  long long ldo;
  int g_in, g_valid, g_out, out_off;
  int kv_mode, kv_B, kv_H, kv_T;
};

using namespace wxbtc;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, int STAGES, bool PAIR = false>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;  // a pair's CTA holds half of the W tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 8 * (2 * STAGES + 4) + 16 + 1024;  // + alignment slack
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// In the cluster window the shared-memory addresses of the two CTAs of a pair differ in bit 24; clearing it names the
// leader's copy of an object (same offset) from either CTA.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
// 2-D TMA load into this CTA's shared memory whose bytes complete on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// tcgen05.commit of the pair's MMAs arriving on the mbarrier at this offset in both CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the leader's copy of `bar` (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

template <int BN, int STAGES, bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using L = SmemLayout<BN, STAGES, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PAIR: a work item is a pair of m-adjacent tiles; this CTA takes m block 2 * mp + rank (a phantom block past the
  // matrix still loads its W half and keeps the barriers in step: TMA zero-fills its A rows, the epilogue masks them)
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int num_mp = PAIR ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles;
  const int num_tiles = num_mp * p.num_n_tiles;
  const int first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + a, 1);
      mbar_init(tempty_bar + a, PAIR ? 2 * EPI_WARPS : EPI_WARPS);  // one arrive per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR) {  // the same warp of both CTAs allocates for the pair
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (producer and issuer warps run converged, one elect.sync lane issues: under a divergent `lane == 0` branch ptxas
    //  wraps every TMA / tcgen05 instruction in an ELECT + R2UR + branch sequence of ~14 instructions)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first; tile < num_tiles; tile += step) {
        const int m_blk = PAIR ? 2 * (tile % num_mp) + rank : tile % num_mp, n_blk = tile / num_mp;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (elect_one()) {
            if (PAIR) {
              if (rank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * L::STAGE_BYTES);  // both CTAs' tiles
              tma_load_2d_pair(sa, &tmA, full_bar + stage, kb * BK, m_blk * BM);
              tma_load_2d_pair(sb, &tmB, full_bar + stage, kb * BK, n_blk * BN + rank * (BN / 2));
            } else {
              mbar_arrive_expect_tx(full_bar + stage, L::STAGE_BYTES);
              tma_load_2d(sa, &tmA, full_bar + stage, kb * BK, m_blk * BM);
              tma_load_2d(sb, &tmB, full_bar + stage, kb * BK, n_blk * BN);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (rank == 0) {
      // instruction descriptor: D=f32 (bit 4), A=bf16 (bit 7), B=bf16 (bit 10), K-major A and B,
      // N>>3 at bit 17, M>>4 at bit 24 (M = 256 for the pair's MMA)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first; tile < num_tiles; tile += step) {
        mbar_wait(tempty_bar + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = make_sw128_desc(sa);
          const uint64_t bdesc = make_sw128_desc(sa + L::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
              if (PAIR) tc_mma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              else tc_mma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            if (PAIR) tc_commit_pair(empty_bar + stage);  // frees the slot in both CTAs
            else tc_commit(empty_bar + stage);
            if (kb == p.num_k_blocks - 1) {
              if (PAIR) tc_commit_pair(tfull_bar + acc);  // both CTAs' epilogues read their 128 rows of the accumulator
              else tc_commit(tfull_bar + acc);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp & 3;          // TMEM lane quarter this warp may read (warp % 4)
    const int half = (warp - 4) >> 2;  // which half of the tile's 32-column chunks it handles
    constexpr int CHUNKS = BN / 64;    // chunks per warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first; tile < num_tiles; tile += step) {
      const int m_blk = PAIR ? 2 * (tile % num_mp) + rank : tile % num_mp, n_blk = tile / num_mp;
      const int r = m_blk * BM + ew * 32 + lane;  // GEMM row of this thread
      bool row_ok = r < p.M;
      long long out_row = r;
      int t_in_group = r;
      if (p.g_in > 0) {
        const int g = r / p.g_in;
        t_in_group = r - g * p.g_in;
        row_ok = row_ok && (t_in_group < p.g_valid);
        out_row = (long long)g * p.g_out + t_in_group + p.out_off;
      }
      const float* res_row = nullptr;
      if (p.res_mode == 1) res_row = p.residual + out_row * p.ldr;
      else if (p.res_mode == 2) res_row = p.residual + (long long)t_in_group * p.ldr;
      // the residual values this thread will add (CHUNKS x 128 B of its row) are requested into L2 now, while the tile's
      // products are still being accumulated: the epilogue's loads then find them on chip instead of waiting on HBM (the
      // out-projection, K = d, is otherwise paced by its residual read: 36 % tensor-pipe activity before)
      if (res_row && row_ok) {
        const float* pf = res_row + n_blk * BN + half * CHUNKS * 32;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c)
          if (n_blk * BN + (half * CHUNKS + c) * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + c * 32));
      }
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + half * CHUNKS * 32);
      uint32_t vn[32];
      tc_ld_32x32(taddr0, vn);
#pragma unroll 1
      for (int cc = 0; cc < CHUNKS; ++cc) {
        uint32_t v[32];
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = vn[i];
        if (cc + 1 < CHUNKS) tc_ld_32x32(taddr0 + (cc + 1) * 32, vn);  // in flight while this chunk is processed
        const int n0 = n_blk * BN + (half * CHUNKS + cc) * 32;
        if (row_ok && n0 < p.N) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          const bool full = (n0 + 32 <= p.N);
          if (p.bias) {
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
                f[i] += bb.x; f[i + 1] += bb.y; f[i + 2] += bb.z; f[i + 3] += bb.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n0 + i < p.N) f[i] += __ldg(p.bias + n0 + i);
            }
          }
          if (p.gelu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = gelu_fast(f[i]);
          }
          if (res_row) {
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 rr = *reinterpret_cast<const float4*>(res_row + n0 + i);
                f[i] += rr.x; f[i + 1] += rr.y; f[i + 2] += rr.z; f[i + 3] += rr.w;
              }
            } else {
              for (int i = 0; i < 32; ++i)
                if (n0 + i < p.N) f[i] += res_row[n0 + i];
            }
          }
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n0;
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
            } else {
              for (int i = 0; i < 32; ++i)
                if (n0 + i < p.N) o[i] = f[i];
            }
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + n0;
            if (p.kv_mode) {
              const int dm = p.N >> 1;
              const int kv = n0 / dm, rem = n0 - kv * dm;
              const int hh = rem >> 6, j0 = rem & 63;
              const int bb = r / p.kv_T, tt = r - bb * p.kv_T;
              o = reinterpret_cast<__nv_bfloat16*>(p.out) +
                  ((((long long)kv * p.kv_B + bb) * p.kv_H + hh) * p.kv_T + tt) * 64 + j0;
            }
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 pk;
                __nv_bfloat162 b0 = __floats2bfloat162_rn(f[i], f[i + 1]);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(f[i + 2], f[i + 3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[i + 4], f[i + 5]);
                __nv_bfloat162 b3 = __floats2bfloat162_rn(f[i + 6], f[i + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&b0);
                pk.y = *reinterpret_cast<uint32_t*>(&b1);
                pk.z = *reinterpret_cast<uint32_t*>(&b2);
                pk.w = *reinterpret_cast<uint32_t*>(&b3);
                *reinterpret_cast<uint4*>(o + i) = pk;
              }
            } else {
              for (int i = 0; i < 32; ++i)
                if (n0 + i < p.N) o[i] = __float2bfloat16_rn(f[i]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(tempty_bar + acc);
        else mbar_arrive(tempty_bar + acc);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA leaves while its peer may still multicast into its shared memory
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int wxb_make_tmap_bf16(wxb_ctx* ctx, CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_rows) {
  if (!ctx->encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
      return wxb_fail(ctx, WXB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    ctx->encode_tiled = fn;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_bytes & 15))
    return wxb_fail(ctx, WXB_ERR_INVALID, "TMA operand must be 16-byte aligned (base %p, row stride %llu B)", base,
                    (unsigned long long)row_stride_bytes);
  // encoded maps are cached per ctx: the encoder makes ~200 GEMM calls per pass over a handful of (buffer, shape) pairs
  char keybuf[160];
  snprintf(keybuf, sizeof(keybuf), "%p/%llu/%llu/%llu/%u/%u", base, (unsigned long long)inner, (unsigned long long)rows,
           (unsigned long long)row_stride_bytes, box_inner, box_rows);
  auto hit = ctx->tmap_cache.find(keybuf);
  if (hit != ctx->tmap_cache.end()) {
    memcpy(tm, hit->second.data(), sizeof(CUtensorMap));
    return WXB_OK;
  }
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)ctx->encode_tiled)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                                                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return wxb_fail(ctx, WXB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%llu rows=%llu stride=%llu", (int)r,
                    (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_bytes);
  if (ctx->tmap_cache.size() > 4096) ctx->tmap_cache.clear();
  std::vector<unsigned char>& slot = ctx->tmap_cache[keybuf];
  slot.resize(sizeof(CUtensorMap));
  memcpy(slot.data(), tm, sizeof(CUtensorMap));
  return WXB_OK;
}

namespace {

template <int BN, int STAGES, bool PAIR>
int launch_cfg(wxb_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES, PAIR>;
  auto kern = gemm_tc_kernel<BN, STAGES, PAIR>;
  {
    int rc = wxb_func_smem(ctx, kern, L::TOTAL);
    if (rc != WXB_OK) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (PAIR) {
    const int pairs = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    const int max_pairs = ctx->sm_count / 2;
    cfg.gridDim = dim3(2 * (pairs < max_pairs ? pairs : max_pairs));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    cfg.gridDim = dim3(tiles < ctx->sm_count ? tiles : ctx->sm_count);
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
  ctx->launches++;
  if (e != cudaSuccess) return wxb_fail(ctx, WXB_ERR_CUDA, "gemm launch failed: %s", cudaGetErrorString(e));
  return WXB_OK;
}

// WXB_GEMM=1cta keeps every GEMM on the single-CTA kernel (A/B timing)
bool gemm_pairs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXB_GEMM");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

int wxb_gemm_launch(wxb_ctx* ctx, const GemmArgs& a, cudaStream_t st) {
  if (!a.A || !a.W || !a.out || a.M <= 0 || a.N <= 0 || a.K <= 0)
    return wxb_fail(ctx, WXB_ERR_INVALID, "gemm: bad argument (M=%d N=%d K=%d)", a.M, a.N, a.K);
  if (a.K % 8 != 0) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "gemm: K=%d must be a multiple of 8", a.K);
  // tile width: 256 where it divides N (or N is large), else 128
  const int BN = (a.N % 256 == 0 || a.N >= 2048) ? 256 : 128;
  GemmParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.num_m_tiles = ceil_div(a.M, BM);
  p.num_n_tiles = ceil_div(a.N, BN);
  p.num_k_blocks = ceil_div(a.K, BK);
  p.bias = a.bias; p.residual = a.residual; p.res_mode = a.residual ? a.res_mode : 0; p.ldr = a.ldr;
  p.gelu = a.gelu; p.out = a.out; p.out_f32 = a.out_f32; p.ldo = a.ldo ? a.ldo : a.N;
  p.g_in = a.g_in; p.g_valid = a.g_valid; p.g_out = a.g_out; p.out_off = a.out_off;
  p.kv_mode = a.kv_mode; p.kv_B = a.kv_B; p.kv_H = a.kv_H; p.kv_T = a.kv_T;
  if (a.kv_mode && (a.out_f32 || (a.N % 128) != 0)) return wxb_fail(ctx, WXB_ERR_INVALID, "gemm: kv_mode needs bf16 output and N %% 128 == 0");
  CUtensorMap tmA, tmB;
  int rc;
  const long long lda = a.lda ? a.lda : a.K;
  if ((rc = wxb_make_tmap_bf16(ctx, &tmA, a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)lda * 2, BK, BM)) != WXB_OK) return rc;
  // pairs of CTAs sharing the W tile where there is enough work for every pair of SMs
  const bool pair = gemm_pairs_enabled() && BN == 256 && p.num_m_tiles * p.num_n_tiles >= 2 * ctx->sm_count;
  if ((rc = wxb_make_tmap_bf16(ctx, &tmB, a.W, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.K * 2, BK, pair ? BN / 2 : BN)) != WXB_OK) return rc;
  if (pair) return launch_cfg<256, 6, true>(ctx, tmA, tmB, p, st);
  if (BN == 256) return launch_cfg<256, 4, false>(ctx, tmA, tmB, p, st);
  return launch_cfg<128, 6, false>(ctx, tmA, tmB, p, st);
}

extern "C" int wxb_gemm_bf16(wxb_ctx* ctx, const void* A_dev, const void* W_dev, const float* bias_dev, void* D_dev,
                             int M, int N, int K, int flags, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  GemmArgs a;
  a.A = (const __nv_bfloat16*)A_dev; a.lda = K; a.M = M;
  a.W = (const __nv_bfloat16*)W_dev; a.N = N; a.K = K;
  a.bias = bias_dev; a.gelu = flags & 1; a.out = D_dev; a.out_f32 = (flags >> 1) & 1; a.ldo = N;
  return wxb_gemm_launch(ctx, a, (cudaStream_t)stream);
}
