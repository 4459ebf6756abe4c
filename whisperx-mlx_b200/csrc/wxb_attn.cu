// wxb_attn.cu — encoder self-attention (non-causal, head_dim 64, T = 1500) on the 5th-gen tensor cores.
//
// One CTA = two 128-query tiles A and B of one (head, chunk), worked on by two softmax warpgroups in ping-pong: while
// warpgroup A exponentiates S_A the tensor core computes S_B and P_B V, and vice versa.  Per 128-key tile j and q tile t:
//   S_t = Q_t K_j^T        tcgen05.mma M=128 N=128 K=64  (Q, K_j: TMA 128B-swizzled K-major tiles)  -> TMEM
//   P_t = exp2((S_t - m) c)    softmax warps: one thread per query row reads its S row with tcgen05.ld, keeps the
//                              max / sum in registers, writes P_t (bf16) into shared memory in the canonical K-major
//                              swizzled layout (A operand of the next MMA)
//   O_t += P_t V_j         tcgen05.mma M=128 N=64 K=128 (P from TMEM, V_j rows as an MN-major B operand), accumulated IN TMEM
// The offset m baked into O_t and the row sum is only moved when the running row maximum has grown by more than 2^8
// (P stays within bf16 / fp32 range), so the O_t rescale (tcgen05.ld, multiply, tcgen05.st) is rare after the first
// tiles and the softmax threads touch O only once more, for the final 1 / sum.
// Warp roles: 0..3 softmax A, 4..7 softmax B, 8 = TMA producer, 9 = MMA issuer, 10 = TMEM allocator.
//
// Measured and not kept in round 2 (same shape, A/B inside one gpurun call; git history has both kernels):
//   * S read from tensor memory ONCE per key tile (offset fixed by the first tile's exact row maximum, growth of the maximum
//     detected on the fly, tile redone on the rare overflow): parity-clean, 1.233 ms vs 1.160 ms - tensor-memory read
//     bandwidth is not the limiter, the extra max tracking in the exp pass costs more than the second tcgen05.ld;
//   * two threads per query row (16 softmax warps, half-row maxima / sums exchanged through shared memory behind named
//     barriers; 96 registers per thread because 18 warps put 5 on one scheduler): parity-clean, 1.252 ms vs 1.159 ms;
//   * other FMA-pipe / MUFU splits of the exponentials (ATTN_POLY): flat between a quarter and a half of the pairs.
#include "wxb_common.cuh"
#include "wxb_tc.cuh"
#include <math.h>

using namespace wxbtc;

namespace {

constexpr int BQ = 128, BKV = 128;
constexpr int AT_THREADS = 384;
constexpr int Q_BYTES = BQ * 64 * 2;          // 16 KB per q tile
constexpr int K_BYTES = BKV * 64 * 2;         // 16 KB
constexpr int VB_BYTES = 64 * 64 * 2;         // one [64 dims x 64 keys] K-block of V^T, 8 KB
constexpr int V_BYTES = 2 * VB_BYTES;         // 16 KB
constexpr int PB_BYTES = BQ * 64 * 2;         // one [128 q x 64 keys] K-block of P, 16 KB
constexpr int P_BYTES = 2 * PB_BYTES;         // 32 KB per q tile
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + 2 * Q_BYTES;
constexpr int OFF_V = OFF_K + 2 * K_BYTES;
constexpr int OFF_P = OFF_V + 2 * V_BYTES;
constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;  // 163840
constexpr int AT_SMEM = OFF_BAR + 16 * 8 + 16 + 1024;
constexpr int TM_S = 0, TM_O = 256, TM_P = 384;  // TMEM columns: S_A, S_B at 0 / 128, O_A, O_B at 256 / 320, P_A, P_B (bf16 pairs) at 384 / 448
#ifndef WXB_ATTN_P_TMEM
#define WXB_ATTN_P_TMEM 1
#endif
#ifndef WXB_ATTN_POLY
#define WXB_ATTN_POLY 2
#endif
// which pairs of a row's exponentials run on the FMA pipe instead of MUFU.EX2: 0 = none, 1 = every 4th pair, 2 = every other pair,
// 3 = three of four pairs (measured, 60 x 20 heads, T = 1500: 1.250 / 1.153 / 1.160 / 1.231 ms per layer)
constexpr int ATTN_POLY = WXB_ATTN_POLY;
constexpr float RESCALE_LOG2 = 8.f;           // move the softmax offset only when the row max grew by more than 2^8

struct AttnTcParams {
  int T;            // rows per sequence in qkv / out (row stride between sequences)
  int d, H;
  const int* lens;  // optional [B]: valid positions of every sequence (<= T); nullptr = T for all
  float scale_log2;
  __nv_bfloat16* out;
};

__device__ __forceinline__ void tc_st_32x32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (lane = row, one 32-bit column = two consecutive bf16 along K) never
// passes through shared memory, so the product is not paced by the 128 A-row reads of the shared-memory form
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2), 3-input max, bare MUFU.EX2 -------------------------------------------
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// p = 2^(s c - off) for 64 scores of one query row, rounded to bf16 into one 128-byte row of a swizzled P block; returns
// the row sum.  The MUFU unit delivers 4 exponentials per clock and scheduler, an eighth of the FMA rate, so every other
// PAIR of elements is exponentiated on the FMA pipe instead: round-to-nearest split x = n + f by the 1.5 * 2^23 trick,
// 2^f by a degree-3 minimax polynomial on [-0.5, 0.5] (relative error 1e-4, far below the bf16 rounding of P),
// 2^n by adding n to the exponent field.
template <bool FULL>
__device__ __forceinline__ float exp_block64(const uint32_t* sv, int col0, int valid, float sl, float off, uint8_t* prow_blk, int r, uint32_t p_taddr) {
  const uint64_t sl2 = pack2(sl, sl), noff2 = pack2(-off, -off);
  const uint64_t magic2 = pack2(12582912.f, 12582912.f), nmagic2 = pack2(-12582912.f, -12582912.f), neg1 = pack2(-1.f, -1.f);
  const uint64_t c3 = pack2(0.05500893f, 0.05500893f), c2 = pack2(0.24221097f, 0.24221097f), c1 = pack2(0.69328293f, 0.69328293f);
  const uint64_t one2 = pack2(1.f, 1.f);
  uint64_t acc = pack2(0.f, 0.f);
  uint32_t pk[32];
#pragma unroll
  for (int i = 0; i < 64; i += 2) {
    const uint64_t x2 = ffma2(pack2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2, noff2);
    float p0, p1;
    if (ATTN_POLY == 2 ? ((i >> 1) & 1) != 0 : ATTN_POLY == 1 ? ((i >> 1) & 3) == 3 : ATTN_POLY == 3 ? ((i >> 1) & 3) != 0 : false) {
      float x0, x1;
      unpack2(x2, x0, x1);
      const uint64_t xc = pack2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
      const uint64_t rr = fadd2(xc, magic2);                       // low mantissa bits = round(x)
      const uint64_t f2 = ffma2(fadd2(rr, nmagic2), neg1, xc);     // x - round(x)
      uint64_t q2 = ffma2(f2, c3, c2);
      q2 = ffma2(q2, f2, c1);
      q2 = ffma2(q2, f2, one2);
      float q0, q1, r0, r1;
      unpack2(q2, q0, q1);
      unpack2(rr, r0, r1);
      p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
      p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
    } else {
      float x0, x1;
      unpack2(x2, x0, x1);
      p0 = ex2(x0);
      p1 = ex2(x1);
    }
    if (!FULL) {
      if (col0 + i >= valid) p0 = 0.f;
      if (col0 + i + 1 >= valid) p1 = 0.f;
    }
    acc = fadd2(acc, pack2(p0, p1));
    __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
    pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);
  }
#if WXB_ATTN_P_TMEM
  tc_st_32x32(p_taddr, pk);  // 32 columns = this row's 64 probabilities as bf16 pairs
#else
#pragma unroll
  for (int ch = 0; ch < 8; ++ch)
    *reinterpret_cast<uint4*>(prow_blk + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
#endif
  float a0, a1;
  unpack2(acc, a0, a1);
  return a0 + a1;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const AttnTcParams p) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = at_smem_raw + ((1024u - (smem_u32(at_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2 stages]
  uint64_t* kv_empty = bars + 3;   // [2 stages]
  uint64_t* s_full = bars + 5;     // [2 q tiles]  MMA -> softmax: S_t(j) complete
  uint64_t* p_full = bars + 7;     // [2 q tiles]  softmax -> MMA: P_t(j) written, S_t(j) fully read, O_t rescaled if needed
  uint64_t* pv_done = bars + 9;    // [2 q tiles]  MMA -> softmax: O_t += P_t(j) V_j complete (P_t buffer free, O_t current)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ, h = blockIdx.y, b = blockIdx.z;
  const int Tb = p.lens ? min(__ldg(p.lens + b), p.T) : p.T;  // valid positions of this sequence (keys and queries)
  if (q0 >= Tb) return;  // whole CTA: nothing allocated or initialised yet
  const int n = (Tb + BKV - 1) / BKV;

  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQK)) : "memory");
  }
  if (warp == 9 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(kv_full + i, 1);
      mbar_init(kv_empty + i, 1);
      mbar_init(s_full + i, 1);
      mbar_init(p_full + i, 4);
      mbar_init(pv_done + i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int row0 = b * p.T;
      mbar_arrive_expect_tx(q_full, 2 * Q_BYTES);
      tma_load_2d(smem + OFF_Q, &tmQK, q_full, h * 64, row0 + q0);
      tma_load_2d(smem + OFF_Q + Q_BYTES, &tmQK, q_full, h * 64, row0 + q0 + BQ);
      for (int j = 0; j < n; ++j) {
        const int st = j & 1;
        mbar_wait(kv_empty + st, ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full + st, K_BYTES + V_BYTES);
        tma_load_2d(smem + OFF_K + st * K_BYTES, &tmQK, kv_full + st, p.d + h * 64, row0 + j * BKV);
        // V rows exactly as they sit in qkv ([key][64 dims], 128-byte swizzle): the PV product reads the tile as an
        // MN-major B operand, so no transposed copy of V exists (keys past T are masked by p = 0; they hold finite data
        // of the next chunk, or zeros past the tensor)
        tma_load_2d(smem + OFF_V + st * V_BYTES, &tmQK, kv_full + st, 2 * p.d + h * 64, row0 + j * BKV);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    // (the warp runs converged and one elect.sync lane issues: under a divergent `lane == 0` branch ptxas wraps every
    //  tcgen05 instruction in an ELECT + 5 x R2UR + branch sequence, ~14 instructions per product)
    {
      const uint32_t idescS = make_idesc_bf16(128, 128), idescO = make_idesc_bf16(128, 64) | (1u << 16);  // B = V is MN-major
      auto issue_S = [&](int t, int j) {  // S_t(j) = Q_t K_j^T
        const uint64_t adesc = make_sw128_desc(smem_u32(smem + OFF_Q + t * Q_BYTES));
        const uint64_t bdesc = make_sw128_desc(smem_u32(smem + OFF_K + (j & 1) * K_BYTES));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_base + TM_S + t * 128, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idescS, k != 0);
          tc_commit(s_full + t);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(kv_full + 0, 0);
      tc_fence_after();
      issue_S(0, 0);
      issue_S(1, 0);
      for (int j = 0; j < n; ++j) {
        const int st = j & 1;
        for (int t = 0; t < 2; ++t) {
          mbar_wait(p_full + t, j & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
          for (int kb2 = 0; kb2 < 2; ++kb2) {
            const uint64_t adesc = make_sw128_desc(smem_u32(smem + OFF_P + t * P_BYTES + kb2 * PB_BYTES));
            // MN-major B: 8-key groups 1024 B apart (SBO), one K = 16 step = 16 keys = 2048 B further into the tile
            const uint64_t bdesc = make_sw128_desc(smem_u32(smem + OFF_V + st * V_BYTES + kb2 * VB_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#if WXB_ATTN_P_TMEM
              (void)adesc;  // P comes from TMEM: 8 columns (16 keys) per step
              tc_mma_bf16_ts(tmem_base + TM_O + t * 64, tmem_base + TM_P + t * 64 + kb2 * 32 + k * 8, bdesc + (uint64_t)(k * (2048 >> 4)), idescO, (j | kb2 | k) != 0);
#else
              tc_mma_bf16(tmem_base + TM_O + t * 64, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * (2048 >> 4)), idescO, (j | kb2 | k) != 0);
#endif
            }
          }
          tc_commit(pv_done + t);
          if (t == 1) tc_commit(kv_empty + st);  // S_A, S_B, PV_A, PV_B of this stage are all behind this commit
          }
          __syncwarp();
          if (j + 1 < n) {
            if (t == 0) {
              mbar_wait(kv_full + ((j + 1) & 1), ((j + 1) >> 1) & 1);
              tc_fence_after();
            }
            issue_S(t, j + 1);  // S_t is free: P_t(j) is only published after S_t(j) has been read completely
          }
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ softmax warpgroups + output
    const int t = warp >> 2, ew = warp & 3;
    const int r = ew * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + TM_S + t * 128;
    const uint32_t o_addr = tmem_base + lane_addr + TM_O + t * 64;
    const uint32_t p_addr = tmem_base + lane_addr + TM_P + t * 64;
    const float sl = p.scale_log2;
    float m_used = -INFINITY, m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = smem + OFF_P + t * P_BYTES + r * 128;

    for (int j = 0; j < n; ++j) {
      const int valid = Tb - j * BKV;  // keys of this tile that exist (>= 128 except for the last tile)
      mbar_wait(s_full + t, j & 1);
      tc_fence_after();
      // columns 64..127 stay in registers, columns 0..63 are read twice (max, then exp)
      uint32_t hi[64];
      tc_ld_32x32(s_addr + 64, hi);
      tc_ld_32x32(s_addr + 96, hi + 32);
      float mx = -INFINITY;
      {
        uint32_t lo[64];
        tc_ld_32x32(s_addr, lo);
        tc_ld_32x32(s_addr + 32, lo + 32);
        tc_wait_ld();
        if (valid >= BKV) {
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            mx = max3(mx, __uint_as_float(lo[i]), __uint_as_float(lo[i + 1]));
            mx = max3(mx, __uint_as_float(hi[i]), __uint_as_float(hi[i + 1]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            if (i < valid) mx = fmaxf(mx, __uint_as_float(lo[i]));
            if (64 + i < valid) mx = fmaxf(mx, __uint_as_float(hi[i]));
          }
        }
      }
      m_run = fmaxf(m_run, mx);
      const bool need = (m_run - m_used) * sl > RESCALE_LOG2;  // always true for the first tile (m_used = -inf)
      if (j > 0) {
        // P_t is free and O_t holds all tiles < j once PV_t(j-1) has completed
        mbar_wait(pv_done + t, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? exp2f((m_used - m_run) * sl) : 1.f;
          l_run *= alpha;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tc_ld_32x32(o_addr + c * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tc_st_32x32(o_addr + c * 32, v);
          }
          tc_wait_st();
        }
      }
      if (need) m_used = m_run;
      const float off = m_used * sl;
      float rowsum;
      const bool full = valid >= BKV;
      // keys 64..127 -> P block 1
      rowsum = full ? exp_block64<true>(hi, 64, valid, sl, off, prow + PB_BYTES, r, p_addr + 32) : exp_block64<false>(hi, 64, valid, sl, off, prow + PB_BYTES, r, p_addr + 32);
      // keys 0..63 -> P block 0
      {
        uint32_t lo[64];
        tc_ld_32x32(s_addr, lo);
        tc_ld_32x32(s_addr + 32, lo + 32);
        tc_wait_ld();
        rowsum += full ? exp_block64<true>(lo, 0, valid, sl, off, prow, r, p_addr) : exp_block64<false>(lo, 0, valid, sl, off, prow, r, p_addr);
      }
      l_run += rowsum;
      // S_t(j) fully read, O_t consistent, P_t(j) written: publish to the tensor core
#if WXB_ATTN_P_TMEM
      tc_wait_st();
      tc_fence_before();
#else
      tc_fence_before();
      fence_proxy_async();
#endif
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + t);
    }
    // ---- output: O_t / l
    mbar_wait(pv_done + t, (n - 1) & 1);
    tc_fence_after();
    const int q = q0 + t * BQ + r;
    const float inv = 1.f / l_run;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tc_ld_32x32(o_addr + c * 32, v);
      tc_wait_ld();
      if (q < Tb) {
        __nv_bfloat16* dst = p.out + ((size_t)b * p.T + q) * p.d + h * 64 + c * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
          __nv_bfloat162 b1 = __floats2bfloat162_rn(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
          __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
          __nv_bfloat162 b3 = __floats2bfloat162_rn(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
          pk.x = *reinterpret_cast<uint32_t*>(&b0);
          pk.y = *reinterpret_cast<uint32_t*>(&b1);
          pk.z = *reinterpret_cast<uint32_t*>(&b2);
          pk.w = *reinterpret_cast<uint32_t*>(&b3);
          *reinterpret_cast<uint4*>(dst + i) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int wxb_make_tmap_bf16(wxb_ctx* ctx, CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_rows);

// qkv bf16 [B*T, 3d] -> out bf16 [B*T, d]; lens_dev (optional, device int[B]) = valid positions per sequence
int wxb_attention_tc(wxb_ctx* ctx, const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int d, int H, const int* lens_dev,
                     cudaStream_t st) {
  CUtensorMap tmQK;
  int rc;
  if ((rc = wxb_make_tmap_bf16(ctx, &tmQK, qkv, (uint64_t)3 * d, (uint64_t)B * T, (uint64_t)3 * d * 2, 64, 128)) != WXB_OK) return rc;
  if ((rc = wxb_func_smem(ctx, attention_tc_kernel, AT_SMEM)) != WXB_OK) return rc;
  AttnTcParams p;
  p.T = T; p.d = d; p.H = H; p.lens = lens_dev;
  p.scale_log2 = (1.0f / sqrtf(64.f)) * 1.44269504088896341f;
  p.out = out;
  attention_tc_kernel<<<dim3(ceil_div(T, 2 * BQ), H, B), AT_THREADS, AT_SMEM, st>>>(tmQK, p);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}
