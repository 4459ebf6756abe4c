// wxb_decoder.cu — K3: batched greedy KV-cache decoder (Whisper TextDecoder) for VAD-cut chunks.
//
// Behavioural spec: the reference's in-tree batched loop
//   /root/reference/mlx_whisper_batch_decoder.py:317-384 (_main_loop_batch), :267-303 (update),
//   :386-468 (run: EOT trimming, avg_logprob), :37-100 (active-sequence masking), filters per SURVEY A.3
//   (SuppressBlank, SuppressTokens, ApplyTimestampRules: /root/reference/mlx_ultra_optimized_batch.py:38-71).
//
// One decode step streams ~17 GB at large-v3 / batch 60 (weights once, cross-KV once per sequence) through
// ~350 dependent small operators.  Launch/dependency latency, not bandwidth, dominated a kernel-per-operator
// design (measured: 8.7 us per dependent kernel, chain 3.1 ms + attention 2.9 ms per step, not overlapping),
// so the whole step is ONE persistent cooperative kernel: one 12-warp CTA per SM, operators are phases separated by
// a grid barrier (release-add + acquire-poll on one L2 counter, ~1.3 us):
//
//   per layer   LN1 | QKV | self-attention | out | LN2 | cq | cross-attention | cout | LN3 | fc1 | fc2
//   then        final LN | logits | no_speech_prob + filters + argmax + logsumexp + EOT latch
//
//   GEMV phases   y[b,n] = sum_k act[b,k] W[n,k] on the 5th-gen tensor cores: tcgen05.mma with the WEIGHTS as the M
//                 operand (128 rows per tile) and the (live) batch rows as the N operand (16 MT <= 64), both staged by
//                 TMA (128-byte swizzle) into a 6-stage mbarrier ring, fp32 accumulators in TMEM, read back with tcgen05.ld.  A CTA tile is 128 weight rows x (K / gk slice).  Split-K tiles write fp32 partials
//                 [gk][B][N]; the CONSUMER phase sums them in slice order (deterministic) together with bias /
//                 residual / LayerNorm / q-scaling / KV-cache append, so no reduction phase exists.
//   self-attn     one warp per (sequence, head): 8 lanes x 16 B per key row, online softmax per 8-lane key slot,
//                 cached K/V rows staged a few 16-key chunks ahead with cp.async.
//   cross-attn    every CTA streams whole (sequence, head) slabs of 192 KB K + 192 KB V through a TMA ring (warp 0 =
//                 producer), 11 consumer warps do S = K q / O += V^T p with mma.sync m16n8k16; the remainder slabs are
//                 cut into pieces whose partial states are merged by the last-arriving CTA (self-cleaning ticket).
//
// Active-sequence compaction (reference :37-100): rows that have emitted EOT leave the batch at launch boundaries.  The
// host polls the done flags every `check_every` steps; the next launch carries the list of live rows (`rows`), so GEMV
// batch tiles, LayerNorm rows, attention units and cross-K/V slabs exist only for live rows; finished rows keep their
// latched EOT (written by the host-side finalize).
//
// HBM layout (L decoder layers, B sequences, H heads, d = 64 H):
//   self K/V  bf16 [L][2][B][H][448][64]      cross K/V bf16 [L][2][B][H][1500][64]
//   x f32 [B,d] residual; xn/att bf16 [B,d]; hid bf16 [B,4d]; part f32 [gk][B][N]; logits f32 [B,V]
#include "wxb_gemm.cuh"
#include "wxb_model.cuh"
#include "wxb_tc.cuh"
#include <math.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

int wxb_make_tmap_bf16(wxb_ctx* ctx, CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_rows);  // wxb_gemm.cu

namespace {

using namespace wxbtc;

constexpr int T_AUDIO = 1500;
// Timing probes that leave phases out (results become meaningless).  The mask is a kernel PARAMETER that only a
// -DWXB_PROBE build of the host code can set (dec_skip_mask(): the shipped library ignores WXB_DEC_SKIP and always passes 0).
// The tests stay in the kernel on purpose: compiling them out changes ptxas' register allocation of the persistent kernel
// enough to create an 8-byte stack frame, and any stack frame costs 4 % or more of every phase (DESIGN.md "Stack frames").
#define WXB_SKIP(mask, bit) (((mask) & (bit)) != 0)
#ifndef WXB_MK_THREADS
#define WXB_MK_THREADS 384
#endif
constexpr int MK_THREADS = WXB_MK_THREADS;  // 12 warps: 11 cross-attention consumer warps, one self-attention round at 1200 units
constexpr int MK_WARPS = MK_THREADS / 32;
// 6-warp CTAs are built so that TWO kernel instances share every SM (launch bounds, shared-memory plan): the batch is cut into
// sequence groups, one kernel instance per group on its own stream, and one group's latency-bound operator chain runs under the
// other group's cross-K/V stream (launch_groups).  The hardware co-schedules CTAs of several launches that allocate tensor memory on
// one SM as long as their columns fit (tools/probe/coresident_probe.cu) although the occupancy API reports 1 CTA per SM.
constexpr int MK_CTAS_PER_SM = MK_WARPS <= 6 ? 2 : 1;
constexpr int GV_MAX_MT = MK_CTAS_PER_SM == 2 ? 2 : 4;  // m16 batch tiles of one kernel instance (rows per instance <= 16 GV_MAX_MT)
constexpr int LN_V4 = 2;    // LayerNorm phase: float4 groups per thread, d <= 4 * 2 * 256
constexpr int GK_MAX = 10;  // largest split-K factor a GEMV plan may use
constexpr int MAX_LAYERS = 32;
constexpr int MAX_GROUP = 64;  // sequences per call (4 m16 batch tiles)
// original row of every live (compact) row of the launch; file-scope shared memory so that no pointer to it is carried in
// registers across the phases of the persistent kernel (the kernel has no register to spare: see DESIGN.md "Stack frames")
__shared__ int s_rows[MAX_GROUP];


// ---- shared-memory plan of the persistent kernel (dynamic, 1024-byte aligned base) -------------------------
// One region is time-shared by the two TMA rings (GEMV phases and cross-attention never overlap inside a CTA):
//   GEMV ring   GV_NST stages x (A: 128 weight rows x 64 k bf16 = 16 KB | B: Bp batch rows x 64 k), 128-byte swizzle
//   KV ring     XA_NST stages x (K: 112 keys x 64 dims bf16 = 14 KB | V: 14 KB), 128-byte swizzle
// followed by the attention scratch (warp states of two items, raw q rows of two items).
constexpr int GV_ROWS = 128;                 // weight rows per tile = UMMA M
constexpr int GV_BK = 64;                    // k per stage (one 128-byte swizzle row)
constexpr int GV_A_BYTES = GV_ROWS * GV_BK * 2;
constexpr int XA_CW = MK_WARPS - 1;          // cross-attention consumer warps (hardware warps 1 ..); warp 0 produces
constexpr int XA_KEYS = 16 * XA_CW;          // keys per stage: one m16 tile per consumer warp
constexpr int XA_HALF = XA_KEYS * 128;       // bytes of K (or V) per stage
constexpr int XA_TAIL = MK_WARPS == 8 ? 48 : MK_WARPS == 6 ? 64 : 96;  // rows of the short TMA box used when at most this many keys of an item remain (1500 = 8 x 176 + 92 = 18 x 80 + 60)
constexpr int SST_BYTES = (2 * XA_CW * 66 * 4 + 127) & ~127;  // [2 item parities][XA_CW warps][66] floats, padded
constexpr int QRAW_ROWS = GK_MAX + 1;                     // bias row + up to GK_MAX split-K partial rows of q
constexpr int QRAW_BYTES = 2 * QRAW_ROWS * 64 * 4;        // two items in flight
constexpr int SCRATCH_BYTES = SST_BYTES + QRAW_BYTES;     // attention scratch behind the ring
#ifndef WXB_XA_MG
#define WXB_XA_MG 16
#endif
constexpr int XA_MG = WXB_XA_MG;  // piece states merged per L2 round trip
constexpr int XA_NST = MK_WARPS == 8 ? 6 : MK_WARPS == 10 ? 5 : 4;  // K/V ring depth (stages of 2 x 16 XA_CW keys x 128 B)
constexpr int GV_NST = MK_CTAS_PER_SM == 2 ? 4 : 6;  // GEMV ring depth
#ifndef WXB_GV_NACC
#define WXB_GV_NACC 1
#endif
// Independent TMEM accumulators of a GEMV tile (compile-time experiment): with 4, the K = 16 products of one ring stage go
// to accumulators 0 .. 3 in turn so that back-to-back tcgen05.mma do not form one dependent accumulation chain, and the
// epilogue adds them in a fixed order.  Measured A/B at batch 60 (profiles/r2_*.md): no phase got faster (the ~150 cycles per
// small tcgen05.mma are not an accumulator dependency) and every GEMV phase lost 0.1-0.2 us to the extra tcgen05.ld: default 1.
constexpr int GV_NACC = WXB_GV_NACC;         // 1, 2 or 4
constexpr int GV_TMEM_COLS = 64 * GV_NACC < 32 ? 32 : 64 * GV_NACC;
static_assert(GV_NACC == 1 || GV_NACC == 2 || GV_NACC == 4, "GV_NACC");
constexpr int RING_BYTES = XA_NST * 2 * XA_HALF;
constexpr size_t MK_SMEM = RING_BYTES + SCRATCH_BYTES + 1024;
static_assert(GV_NST * (GV_A_BYTES + 16 * GV_MAX_MT * GV_BK * 2) <= RING_BYTES, "GEMV ring must fit the shared region");
static_assert(MK_CTAS_PER_SM * (MK_SMEM + 10 * 1024) <= 228 * 1024, "shared memory of the CTAs that share an SM (dynamic + ~9 KB static + 1 KB reserved each)");

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float ex2_approx(float x) {  // 2^x, -inf -> 0, no range fix-up (arguments are <= XA_LAZY)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Grid barrier of the cooperative launch: monotone arrival counter (zeroed by the host before the launch).
// Arrival is a release-add (orders this CTA's earlier writes, made visible to thread 0 by the block barrier),
// the wait an acquire-poll.  The proxy fence orders the generic-proxy global writes of a phase with the TMA
// (async-proxy) reads of the next one.  While thread 0 arrives and waits, thread 32 runs `pre`: work of the NEXT
// operator that does not depend on this one (requesting its weight / K/V tiles), so that HBM latency is spent
// while the barrier completes.  With `prof` set, CTA 0 records the global timer at every barrier exit.
template <class Pre>
__device__ __forceinline__ void grid_sync(unsigned* bar, const unsigned index, int nc, int cta, unsigned long long* prof, Pre pre) {
  // index: number of this barrier within the launch (1 ..); the counter has reached index * nc once every CTA has arrived.  The
  // target is derived from the schedule position instead of being carried in a register through every phase.
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned target = index * (unsigned)nc;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned v;
    unsigned spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++spins > (1u << 26)) __trap();  // ~30 s: a CTA is missing; fail the launch instead of hanging the GPU
    } while ((int)(v - target) < 0);
    fence_proxy_async_all();  // thread 0 is also the TMA producer of the next phase
    if (prof && cta == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      prof[index - 1u] = t;  // barrier index within the launch
    }
  } else if (threadIdx.x == 32) {
    pre();
  }
  __syncthreads();
}

enum { EPI_PART = 0, EPI_GELU_BF16 = 1, EPI_LOGITS = 2 };

// plan of one GEMV phase: CTA tile = 128 weight rows x (K / gk) columns; tiles = ceil(N / 128) * gk
struct MkGemv {
  int N, K, gk, tiles;
};

struct SampleParams {
  float* logits;  // [B, ldl] (static filters are applied in place), rows 16-byte aligned
  long long ldl;
  int V;
  int* tokens;    // [B, stride]: prompt + sampled tokens; a step at position pos writes column pos + 1
  int stride;
  int prompt_len;
  int eot, suppress_blank, blank_token, n_suppress;
  const int* suppress;
  float* sum_logprob;  // [B]
  int* done;           // [B] 1 once the row has emitted EOT
  float* nsp_out;      // if set: softmax prob of nsp_token from the UNFILTERED logits of this step
  int nsp_token;
  // ApplyTimestampRules (decoding with timestamps, `without_timestamps=False`): SURVEY A.3 filter 3; the last clause is the
  // reference's batch-safe patch /root/reference/mlx_ultra_optimized_batch.py:38-71
  int ts_rules;        // 0 = off
  int ts_begin;        // first timestamp token id
  int no_timestamps;   // <|notimestamps|> id, suppressed when the rules are on
  int max_initial_ts;  // max_initial_timestamp_index (50), < 0 = unlimited
  int* ts_last;        // [B] last sampled timestamp token of each row, -1 = none yet
  // rows shared by several CTAs (persistent kernel only; nullptr = one CTA per row): slice statistics [B][R][8] and a
  // zero-initialised, self-cleaning arrival counter per row
  float* part_stats;
  int* part_ticket;
};

// tensor-map table (device array): per layer {qkv, out, cq, cout, fc1, fc2} weight maps, then emb, xn, att, hid, cross K/V
enum { TM_QKV = 0, TM_OUT = 1, TM_CQ = 2, TM_COUT = 3, TM_FC1 = 4, TM_FC2 = 5, TM_PER_LAYER = 6 };
// table behind the per-layer weight maps: emb | (xn, att, hid) boxes of 16, 32, 48, 64 rows | cross K/V full and tail boxes
enum { TM_EMB = 0, TM_ACT = 1, TM_KV = 13, TM_TAIL_COUNT = 15 };

struct MkParams {
  int B, d, H, L, V, TX;  // B = LIVE rows of this launch (activation buffers are indexed by the compact row number)
  int B0;                 // rows the K/V caches, token table and per-row outputs were allocated for (original row numbers)
  const int* rows;        // [B] original row of compact row i (identity while every row is live)
  int mode;     // 0: no logits (forced prompt token); 1: logits; 2: logits + sampling
  int n_steps;  // consecutive positions decoded by this launch (> 1 only in mode 2)
  unsigned delay_ns;  // sequence group g > 0: the instance idles this long first, so that the groups' cross-attention phases interleave
  int skip;     // profiling aid (WXB_DEC_SKIP): 1 cross-attention, 2 GEMV, 4 self-attention, 8 LayerNorm, 16 cross math, 32 cross merge,
                //   64 every CTA streams the same 4 K/V slabs (all L2 hits)
  const DecLayerW* layers;  // device array [L]
  const CUtensorMap* maps;  // device array [6 L + 5]
  const __nv_bfloat16* emb;
  const float *pos_emb, *lnf_w, *lnf_b;
  int* tokens;
  int tok_stride;
  int* d_pos;
  float* x;
  __nv_bfloat16 *xn, *att, *hid;
  float* part;
  __nv_bfloat16* self_kv;
  const __nv_bfloat16* cross_kv;
  float* logits;
  long long ldl;
  float* apart;  // cross-attention piece states [pieces][66]
  int* ticket;   // [<= gridDim] zero-initialised, self-cleaning
  unsigned* bar;
  unsigned long long* prof;  // optional: barrier-exit timestamps of CTA 0 (WXB_DEC_PROF)
  MkGemv g_qkv, g_dd, g_fc1, g_fc2, g_logits;
  SampleParams sp;
  float scale;
  // cross-attention queries of the alignment heads, kept for the DTW word timing (wxb_dtw.cu): qlog f32 [B0][TX][qA][64]
  // (scaled q of head qhead^-1(a) at every decoded position), qhead int8 [L * H] = index a of (layer, head) or -1
  float* qlog;
  const signed char* qhead;
  int qA;
};

// mbarriers + ring cursors of the persistent kernel.  The barriers live in one shared array (fixed slots sized for
// the deepest rings) addressed as `bars + 8 * slot`, so the whole synchronisation state costs one register plus the
// cursors.  The cursors are advanced identically by every thread (all loop bounds are CTA-uniform), so each thread
// holds its own consistent copy.
enum {
  MB_GV_FULL = 0,    // [6] TMA -> MMA
  MB_GV_EMPTY = 6,   // [6] MMA (tcgen05.commit / 8 warp arrivals) -> TMA
  MB_ACC_FULL = 12,  // [1] MMA -> epilogue
  MB_XK_FULL = 13,   // [6] TMA -> attention warps: the K half of a stage has landed
  MB_XK_EMPTY = 19,  // [6] consumer-warp arrivals -> producer: the K half has been read
  MB_XV_FULL = 25,   // [6] the V half of a stage (read one iteration later than its K half, released on its own)
  MB_XV_EMPTY = 31,  // [6]
  MB_ST_FULL = 37,   // [2] warp states of an item deposited -> merging warp
  MB_ST_FREE = 39,   // [2] merging warp -> writers of the item after next
  MB_Q_FULL = 41,    // [2] raw q rows of an item landed (cp.async arrive-on of the 32 producer lanes)
  MB_Q_FREE = 43,    // [2] consumer warps have built their q fragments
  MB_COUNT = 45
};
struct MkSync {
  uint32_t bars;       // shared-memory address of the barrier array
  uint32_t gv_count;   // GEMV stages issued so far (slot = count % GV_NST, parity = (count / GV_NST) & 1)
  uint32_t acc_count;  // accumulator hand-offs so far
  uint32_t tmem;       // TMEM base address (GV_NACC accumulators of 64 fp32 columns x 128 lanes)
  int pre;             // thread 0: stages of the coming operator already requested before the barrier wait (grid_sync)
  int cta, nc;         // this CTA's index within its group, CTAs per group
  __device__ __forceinline__ uint32_t mb(int slot) const { return bars + 8u * (uint32_t)slot; }
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < MK_WARPS; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = red[0];
#pragma unroll
  for (int w = 1; w < MK_WARPS; ++w) s = fmaxf(s, red[w]);
  return s;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm phase (one CTA per sequence row).  The residual row is first brought up to date:
//   from_embed: x = emb[token] + pos_emb[pos]
//   else:       x += bias + sum over the gk split-K partials of the previous GEMV (slice order)
// then xn = LN(x) in bf16 (two-pass variance, eps 1e-5).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ln_phase(const MkParams& p, bool from_embed, int gk, const float* __restrict__ prev_bias,
                                         const float* __restrict__ lw, const float* __restrict__ lb, int pos, float* red,
                                         const MkSync& sy) {
  const int tid = threadIdx.x, d = p.d, nv = d >> 2;
  for (int b = sy.cta; b < p.B; b += sy.nc) {
    float4 v[LN_V4];
    float s = 0.f;
    int token = 0;
    if (from_embed) {
      token = __ldcg(p.tokens + (size_t)s_rows[b] * p.tok_stride + pos);
      token = min(max(token, 0), p.V - 1);
    }
#pragma unroll
    for (int i = 0; i < LN_V4; ++i) {
      const int c4 = tid + MK_THREADS * i;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < nv) {
        if (from_embed) {
          const uint2 e = __ldg(reinterpret_cast<const uint2*>(p.emb + (size_t)token * d) + c4);
          const float4 pe = __ldg(reinterpret_cast<const float4*>(p.pos_emb + (size_t)pos * d) + c4);
          const float2 e0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&e.x));
          const float2 e1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&e.y));
          v[i] = make_float4(e0.x + pe.x, e0.y + pe.y, e1.x + pe.z, e1.y + pe.w);
        } else {
          // all loads of this element group are independent: one L2 round trip, then a fixed-order sum
          float4 a = __ldcg(reinterpret_cast<const float4*>(p.x + (size_t)b * d) + c4);
          const float4 bb = __ldg(reinterpret_cast<const float4*>(prev_bias) + c4);
          a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
#pragma unroll
          for (int k0 = 0; k0 < GK_MAX; k0 += 5) {
            if (k0 < gk) {
              float4 pp[5];
#pragma unroll
              for (int ks = 0; ks < 5; ++ks)
                if (k0 + ks < gk) pp[ks] = __ldcg(reinterpret_cast<const float4*>(p.part + ((size_t)(k0 + ks) * p.B + b) * d) + c4);
#pragma unroll
              for (int ks = 0; ks < 5; ++ks)
                if (k0 + ks < gk) { a.x += pp[ks].x; a.y += pp[ks].y; a.z += pp[ks].z; a.w += pp[ks].w; }
            }
          }
          v[i] = a;
        }
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    float4 ww[LN_V4], wb[LN_V4];  // requested before the reductions so their latency is hidden
#pragma unroll
    for (int i = 0; i < LN_V4; ++i) {
      const int c4 = tid + MK_THREADS * i;
      if (c4 < nv) { ww[i] = __ldg(reinterpret_cast<const float4*>(lw) + c4); wb[i] = __ldg(reinterpret_cast<const float4*>(lb) + c4); }
    }
    const float mean = block_sum(s, red) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_V4; ++i) {
      const int c4 = tid + MK_THREADS * i;
      if (c4 < nv) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rstd = rsqrtf(block_sum(q, red) / d + 1e-5f);
#pragma unroll
    for (int i = 0; i < LN_V4; ++i) {
      const int c4 = tid + MK_THREADS * i;
      if (c4 < nv) {
        reinterpret_cast<float4*>(p.x + (size_t)b * d)[c4] = v[i];
        uint2 pk;
        pk.x = pack_bf16((v[i].x - mean) * rstd * ww[i].x + wb[i].x, (v[i].y - mean) * rstd * ww[i].y + wb[i].y);
        pk.y = pack_bf16((v[i].z - mean) * rstd * ww[i].z + wb[i].z, (v[i].w - mean) * rstd * ww[i].w + wb[i].w);
        reinterpret_cast<uint2*>(p.xn + (size_t)b * d)[c4] = pk;
      }
    }
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------
// GEMV phase on the 5th-gen tensor cores: D[128 weight rows, Bp batch rows] = W_tile[128, Ks] act[Bp, Ks]^T.
// The WEIGHTS are the M operand (UMMA M = 128), the batch the N operand (UMMA N = Bp = 16 MT), so a weight
// element crosses HBM -> L2 -> shared memory (TMA, 128-byte swizzle) -> tensor core exactly once and never
// touches a register; fp32 accumulation in TMEM.  Thread 0 is the TMA producer, lane 0 of warp 1 issues
// tcgen05.mma, all 8 warps read the accumulator back (tcgen05.ld, lane = weight row) for the epilogue.
// Split-K tiles write fp32 partials [ks][B][N]; their consumer phase performs the reduction.
// ---------------------------------------------------------------------------------------------
template <int MT>
__device__ __forceinline__ void gemv_phase(const MkParams& p, const MkGemv& g, const CUtensorMap* wmap, const CUtensorMap* amap,
                                           const float* __restrict__ bias, const int epi, uint8_t* ring, MkSync& sy) {
  constexpr int Bp = 16 * MT;
  constexpr int B_BYTES = Bp * GV_BK * 2;
  constexpr int STAGE = GV_A_BYTES + B_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = p.B;
  const int Ks = g.K / g.gk, nkb = Ks / GV_BK;
  for (int tile = sy.cta; tile < g.tiles; tile += sy.nc) {
    const int ks = tile % g.gk, rb = tile / g.gk;
    const int k0 = ks * Ks, row0 = rb * GV_ROWS;
    // Producer and issuer warps run converged (every lane polls the barriers, operands are warp-uniform) and ONE
    // elected lane issues: under a divergent `lane == 0` branch ptxas wraps every TMA / tcgen05 instruction in an
    // ELECT + 5 x R2UR + branch sequence (~14 instructions each), which is what the issuing thread's time went into.
    if (warp == 0) {
      // ---- TMA producer: the whole K-slice is requested as fast as ring slots free up ----
      uint32_t c = sy.gv_count;
      const int pre = (tile == sy.cta) ? sy.pre : 0;  // weight halves of the first stages were requested before the barrier wait
      for (int kb = 0; kb < nkb; ++kb, ++c) {
        const uint32_t slot = c % GV_NST, par = (c / GV_NST) & 1;
        uint8_t* sa = ring + slot * STAGE;
        if (kb >= pre) mbar_wait(sy.mb(MB_GV_EMPTY + slot), par ^ 1);
        if (elect_one()) {
          if (kb >= pre) {
            mbar_arrive_expect_tx(sy.mb(MB_GV_FULL + slot), STAGE);
            tma_load_2d(sa, wmap, sy.mb(MB_GV_FULL + slot), k0 + kb * GV_BK, row0);
          }
          tma_load_2d(sa + GV_A_BYTES, amap, sy.mb(MB_GV_FULL + slot), k0 + kb * GV_BK, 0);
        }
        __syncwarp();
      }
      sy.pre = 0;
    } else if (warp == 1) {
      // ---- MMA issuer ----
      const uint32_t idesc = make_idesc_bf16(GV_ROWS, Bp);
      uint32_t c = sy.gv_count;
      for (int kb = 0; kb < nkb; ++kb, ++c) {
        const uint32_t slot = c % GV_NST, par = (c / GV_NST) & 1;
        mbar_wait(sy.mb(MB_GV_FULL + slot), par);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + slot * STAGE);
        const uint64_t adesc = make_sw128_desc(sa);
        const uint64_t bdesc = make_sw128_desc(sa + GV_A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < GV_BK / 16; ++k)  // +32 bytes along K inside the swizzle row = +2 in the address field
            tc_mma_bf16(sy.tmem + (uint32_t)((k % GV_NACC) * 64), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                        (kb | (k / GV_NACC)) != 0);
          tc_commit(sy.mb(MB_GV_EMPTY + slot));
          if (kb == nkb - 1) tc_commit(sy.mb(MB_ACC_FULL));
        }
        __syncwarp();
      }
    }
    sy.gv_count += nkb;
    __syncwarp();
    mbar_wait(sy.mb(MB_ACC_FULL), sy.acc_count & 1);
    sy.acc_count++;
    tc_fence_after();
    // ---- epilogue: warp w reads TMEM lanes 32 (w & 3) .. +31 (weight rows), 16-column groups j = w >> 2, + 2, .. ----
    const int n = row0 + (warp & 3) * 32 + lane;
    // whole groups of 4 warps (one per TMEM lane quarter) share the 16-column groups; the warps of an incomplete last group skip
    constexpr int EPI_GROUPS = MK_WARPS / 4;
    for (int j = (warp >> 2) < EPI_GROUPS ? (warp >> 2) : MT; j < MT; j += EPI_GROUPS) {
      uint32_t v[16];
      const uint32_t taddr = sy.tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(j * 16);
      tc_ld_32x32_x16(taddr, v);
      if (GV_NACC > 1) {
        uint32_t w[16];
#pragma unroll
        for (int a = 1; a < GV_NACC; ++a) {
          tc_ld_32x32_x16(taddr + (uint32_t)(a * 64), w);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
        }
      } else {
        tc_wait_ld();
      }
      if (n < g.N) {
        if (epi == EPI_PART) {
          float* o = p.part + ((size_t)ks * B + j * 16) * g.N + n;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (j * 16 + i < B) o[(size_t)i * g.N] = __uint_as_float(v[i]);
        } else if (epi == EPI_GELU_BF16) {
          const float bn = __ldg(bias + n);
          __nv_bfloat16* o = p.hid + (size_t)(j * 16) * g.N + n;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (j * 16 + i < B) o[(size_t)i * g.N] = __float2bfloat16_rn(gelu_erf(__uint_as_float(v[i]) + bn));
        } else {
          float* o = p.logits + (size_t)(j * 16) * p.ldl + n;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (j * 16 + i < B) o[(size_t)i * p.ldl] = __uint_as_float(v[i]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // the next tile's first MMA overwrites the accumulator
  }
}

// ---------------------------------------------------------------------------------------------
// attention pieces.  Lane layout: slot = lane >> 3 (4 key slots per warp), c8 = lane & 7 owns dims 8 c8 .. 8 c8 + 7.
// One call consumes 4 keys per slot: key(u) = kb + u * KSTRIDE + slot, masked by key < k1.
// ---------------------------------------------------------------------------------------------
template <int KSTRIDE>
__device__ __forceinline__ void att_consume(const uint4* kv, const uint4* vv, int kb, int slot, int k1, const float* qr, float& m,
                                            float& l, float* acc) {
  // Branch-free: the four keys' dot-product chains (8 FMA + 3 shuffles each) and their P V updates are independent instruction
  // streams that the scheduler can interleave; a branch per key would serialise them.  Keys at or past k1 were not loaded: their
  // staging slots hold stale but finite bf16 values (the staging region is zeroed when the kernel starts), their scores are
  // masked to -inf and 0 x finite adds nothing.
  float sc4[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv[u]);
    float s0 = 0.f, s1 = 0.f;  // two partial chains per key
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h2[j]);
      s0 = fmaf(qr[2 * j], f.x, s0);
      s1 = fmaf(qr[2 * j + 1], f.y, s1);
    }
    sc4[u] = s0 + s1;
  }
#pragma unroll
  for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
    for (int u = 0; u < 4; ++u) sc4[u] += __shfl_xor_sync(0xffffffffu, sc4[u], o);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    sc4[u] = (kb + u * KSTRIDE + slot < k1) ? sc4[u] : -INFINITY;
    mx = fmaxf(mx, sc4[u]);
  }
  const float mn = fmaxf(m, mx);
  const float alpha = (m > -INFINITY) ? __expf(m - mn) : 0.f;  // nothing accumulated yet: l = acc = 0 anyway
  m = mn;
  l *= alpha;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] *= alpha;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float pr = (sc4[u] > -INFINITY) ? __expf(sc4[u] - mn) : 0.f;  // mn = -inf only while every key so far was masked
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

// 16-byte asynchronous copy / shared-memory load at a compile-time offset from a base (the offset is an immediate of the instruction)
template <int DST_OFF, int SRC_OFF>
__device__ __forceinline__ void cp_async16(uint32_t dst, const char* src) {
  asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%3], 16;" ::"r"(dst), "l"(src), "n"(DST_OFF), "n"(SRC_OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void lds128(uint32_t addr, uint4& v) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF));
}

// Causal self-attention.  A (sequence, head) unit is finished by one warp: it completes the QKV GEMV for its head
// (sum of split-K partials + bias), appends the new K/V row to the cache, and attends over cached keys plus the new one
// (taken from registers, rounded to bf16 like its cached copy).  Units that do not fill a whole round of warps are NOT given
// a second round of their own: all warps of a CTA share such a unit, each attending over its share of the cached keys, and
// merge their states through shared memory.  With few units (small batches) EVERY unit is shared by W warps (sa_share): the
// phase is a latency chain over the cached keys, and 160 units on 1776 warps would leave 91 % of them idle.
struct SelfUnit {
  float q8[8];   // scaled q, dims 8 c8 .. 8 c8 + 7
  uint4 kq, vq;  // the new key / value row (bf16), same dims
};
__device__ __forceinline__ SelfUnit self_unit_qkv(const MkParams& p, const float* __restrict__ qkv_b, int b, int h, int slot, int c8) {
  const int d = p.d, B = p.B, gk = p.g_qkv.gk, N3 = 3 * d;
  const int col = h * 64 + c8 * 8;
  float q8[8], k8[8], v8[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(qkv_b + col)), a1 = __ldg(reinterpret_cast<const float4*>(qkv_b + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(qkv_b + d + col)), b1 = __ldg(reinterpret_cast<const float4*>(qkv_b + d + col + 4));
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(qkv_b + 2 * d + col)), c1 = __ldg(reinterpret_cast<const float4*>(qkv_b + 2 * d + col + 4));
    q8[0] = a0.x; q8[1] = a0.y; q8[2] = a0.z; q8[3] = a0.w; q8[4] = a1.x; q8[5] = a1.y; q8[6] = a1.z; q8[7] = a1.w;
    k8[0] = b0.x; k8[1] = b0.y; k8[2] = b0.z; k8[3] = b0.w; k8[4] = b1.x; k8[5] = b1.y; k8[6] = b1.z; k8[7] = b1.w;
    v8[0] = c0.x; v8[1] = c0.y; v8[2] = c0.z; v8[3] = c0.w; v8[4] = c1.x; v8[5] = c1.y; v8[6] = c1.z; v8[7] = c1.w;
  }
  for (int ks0 = 0; ks0 < gk; ks0 += 4) {
    // slot s fetches slice ks0 + s (6 independent 16-byte loads), then the 4 slots are summed by shuffles:
    // one L2 round trip per 4 slices and a fixed summation order
    float t[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) t[j] = 0.f;
    if (ks0 + slot < gk) {
      const float* base = p.part + ((size_t)(ks0 + slot) * B + b) * N3 + col;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4 a0 = __ldcg(reinterpret_cast<const float4*>(base + j * d));
        const float4 a1 = __ldcg(reinterpret_cast<const float4*>(base + j * d + 4));
        t[8 * j + 0] = a0.x; t[8 * j + 1] = a0.y; t[8 * j + 2] = a0.z; t[8 * j + 3] = a0.w;
        t[8 * j + 4] = a1.x; t[8 * j + 5] = a1.y; t[8 * j + 6] = a1.z; t[8 * j + 7] = a1.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      t[j] += __shfl_xor_sync(0xffffffffu, t[j], 8);
      t[j] += __shfl_xor_sync(0xffffffffu, t[j], 16);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { q8[j] += t[j]; k8[j] += t[8 + j]; v8[j] += t[16 + j]; }
  }
  SelfUnit u;
  u.kq.x = pack_bf16(k8[0], k8[1]); u.kq.y = pack_bf16(k8[2], k8[3]); u.kq.z = pack_bf16(k8[4], k8[5]); u.kq.w = pack_bf16(k8[6], k8[7]);
  u.vq.x = pack_bf16(v8[0], v8[1]); u.vq.y = pack_bf16(v8[2], v8[3]); u.vq.z = pack_bf16(v8[4], v8[5]); u.vq.w = pack_bf16(v8[6], v8[7]);
#pragma unroll
  for (int j = 0; j < 8; ++j) u.q8[j] = q8[j] * p.scale;
  return u;
}
// online-softmax state of one warp over the cached keys [k_begin, k_end) (and the new key if with_new), merged over
// the warp's 4 key slots: on return every lane holds m, l and the 8 output dims of its c8
#ifndef WXB_SA_AHEAD
#define WXB_SA_AHEAD (MK_WARPS == 8 ? 4 : MK_WARPS == 10 ? 3 : 2)
#endif
constexpr int SA_AHEAD = WXB_SA_AHEAD;  // chunks of 16 keys requested ahead of the one being consumed
constexpr int SA_RING = SA_AHEAD + 1;  // 4 KB chunk slots per warp
#ifndef WXB_SA_EXPERIMENT
static_assert(MK_WARPS * SA_RING * 4096 <= RING_BYTES, "self-attention staging must fit the TMA ring region");
#endif
__device__ __forceinline__ void self_unit_attend(const SelfUnit& u, const __nv_bfloat16* Kb, const __nv_bfloat16* Vb, int k_begin, int k_end,
                                                 bool with_new, int slot, int c8, uint32_t stage, float& m, float& lsum, float* acc) {
  m = -INFINITY; lsum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (with_new) {  // warp-uniform
    const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&u.kq);
    const __nv_bfloat162* vh = reinterpret_cast<const __nv_bfloat162*>(&u.vq);
    float sdot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(kh[j]);
      sdot = fmaf(u.q8[2 * j], f.x, sdot);
      sdot = fmaf(u.q8[2 * j + 1], f.y, sdot);
    }
    sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
    sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
    sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
    if (slot == 0) {
      m = sdot;
      lsum = 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(vh[j]);
        acc[2 * j] = f.x;
        acc[2 * j + 1] = f.y;
      }
    }
  }
  // Cached keys (written by earlier steps), 16 per chunk.  The loop is bound by memory latency, not bandwidth, so the
  // rows of the next SA_AHEAD chunks are kept in flight with cp.async into this warp's slice of the (idle) TMA ring:
  // a lane copies exactly the 16-byte pieces it will read back itself, so shared memory only extends its registers.
  const int n = k_end;
  const int nch = (n > k_begin) ? (n - k_begin + 15) >> 4 : 0;
  const uint32_t lane_off = (uint32_t)((slot * 8 + c8) * 16);  // inside a 512-byte row group (4 keys x 128 B)
  // this lane's piece of key k_begin + slot; chunk c, key group i is 2048 c + 512 i bytes further (rows are 128 bytes)
  const char* kp0 = reinterpret_cast<const char*>(Kb + (size_t)(k_begin + slot) * 64 + c8 * 8);
  const char* vp0 = reinterpret_cast<const char*>(Vb + (size_t)(k_begin + slot) * 64 + c8 * 8);
  auto issue = [&](int c) {  // chunk c: keys k_begin + 16 c + 4 i + slot; K pieces at [c % SA_RING][i], V pieces 2 KB behind
    if (c < nch) {
      const uint32_t dst = stage + (uint32_t)((c % SA_RING) * 4096) + lane_off;
      const char* kp = kp0 + (size_t)c * 2048;
      const char* vp = vp0 + (size_t)c * 2048;
      const int left = n - k_begin - (c << 4) - slot;  // this lane's keys of the chunk: i with 4 i < left
      if (0 < left) { cp_async16<0, 0>(dst, kp); cp_async16<2048, 0>(dst, vp); }
      if (4 < left) { cp_async16<512, 512>(dst, kp); cp_async16<2048 + 512, 512>(dst, vp); }
      if (8 < left) { cp_async16<1024, 1024>(dst, kp); cp_async16<2048 + 1024, 1024>(dst, vp); }
      if (12 < left) { cp_async16<1536, 1536>(dst, kp); cp_async16<2048 + 1536, 1536>(dst, vp); }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per chunk index, empty or not: the wait below counts groups
  };
#pragma unroll
  for (int c = 0; c < SA_AHEAD; ++c) issue(c);
  for (int c = 0; c < nch; ++c) {
    issue(c + SA_AHEAD);
    asm volatile("cp.async.wait_group %0;" ::"n"(SA_AHEAD) : "memory");  // chunk c has landed
    const uint32_t src = stage + (uint32_t)((c % SA_RING) * 4096) + lane_off;
    uint4 kA[4], vA[4];
    lds128<0>(src, kA[0]); lds128<512>(src, kA[1]); lds128<1024>(src, kA[2]); lds128<1536>(src, kA[3]);
    lds128<2048>(src, vA[0]); lds128<2048 + 512>(src, vA[1]); lds128<2048 + 1024>(src, vA[2]); lds128<2048 + 1536>(src, vA[3]);
    att_consume<4>(kA, vA, k_begin + (c << 4), slot, n, u.q8, m, lsum, acc);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // merge the 4 slot states (lanes differing in bits 3 and 4)
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float l2 = __shfl_xor_sync(0xffffffffu, lsum, o);
    const float mn = fmaxf(m, m2);
    const float w1 = (m > -INFINITY) ? __expf(m - mn) : 0.f;
    const float w2 = (m2 > -INFINITY) ? __expf(m2 - mn) : 0.f;
    lsum = lsum * w1 + l2 * w2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a2 = __shfl_xor_sync(0xffffffffu, acc[j], o);
      acc[j] = acc[j] * w1 + a2 * w2;
    }
    m = mn;
  }
}
__device__ __forceinline__ void self_unit_store(const MkParams& p, int b, int h, int c8, float lsum, const float* acc) {
  const float inv = 1.f / lsum;
  uint4 pk;
  pk.x = pack_bf16(acc[0] * inv, acc[1] * inv);
  pk.y = pack_bf16(acc[2] * inv, acc[3] * inv);
  pk.z = pack_bf16(acc[4] * inv, acc[5] * inv);
  pk.w = pack_bf16(acc[6] * inv, acc[7] * inv);
  *reinterpret_cast<uint4*>(p.att + (size_t)b * p.d + h * 64 + c8 * 8) = pk;
}

// warps of a CTA that share one unit in the shared round of the self-attention phase: all of them when only the remainder units of a
// large batch are shared, else the largest divisor of the warp count that fits every unit into that one round
#ifndef WXB_SA_MIN_SHARE
#define WXB_SA_MIN_SHARE 2
#endif
constexpr int SA_MIN_SHARE = WXB_SA_MIN_SHARE;  // fewest warps per unit for which sharing every unit pays (A/B knob; MK_WARPS + 1 = never)
__device__ __forceinline__ int sa_share(int n_units, int n_warps) {
  int W = MK_WARPS;
  if (n_units * SA_MIN_SHARE <= n_warps) {
#pragma unroll
    for (int c = SA_MIN_SHARE; c <= MK_WARPS; ++c)
      if (MK_WARPS % c == 0 && n_units * c <= n_warps) W = c;
  }
  return W;
}

__device__ __forceinline__ void self_attn_phase(const MkParams& p, int l, const float* __restrict__ qkv_b, int pos, uint8_t* ring,
                                                float* scratch, const MkSync& sy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, slot = lane >> 3, c8 = lane & 7;
  const uint32_t stage = smem_u32(ring) + (uint32_t)(warp * SA_RING * 4096);  // this warp's K/V staging slots
  const int H = p.H, B = p.B, TX = p.TX;
  __nv_bfloat16* sk = p.self_kv + (size_t)l * 2 * p.B0 * H * TX * 64;
  __nv_bfloat16* sv = sk + (size_t)p.B0 * H * TX * 64;
  const int n_units = B * H, n_warps = sy.nc * MK_WARPS;
  // Work items of a warp: its solo units (one warp per unit, whole rounds of n_warps units), then one SHARED round in which W warps
  // of a CTA finish one unit together, each attending over 1 / W of the cached keys, states merged through shared memory:
  //   many units (batch 60: 1200 units on 1776 warps)  the n_units % n_warps units that would need a second round of their own
  //                                                    are shared by all warps of a CTA (W = MK_WARPS);
  //   few units (small batches: strong scaling over several GPUs, BASELINE config 3)  the phase is a latency chain over the cached
  //                                                    keys, so EVERY unit is shared by W warps, W = the largest divisor of the warp
  //                                                    count that fits all units into the one shared round.
  int n_solo = n_units;  // units [0, n_solo) get a warp each; units [n_solo, n_units) are shared by W warps of a CTA
  const int left = n_units % n_warps;
  if (left > 0 && left <= sy.nc && n_units > n_warps) n_solo = n_units - left;
  if (n_units * SA_MIN_SHARE <= n_warps) n_solo = 0;
  // work items of this warp: its solo units, then (CTAs with a shared unit) its share of a shared unit's cached keys
  const int n_rounds = (n_solo + n_warps - 1) / n_warps;
  const bool shared_unit = n_solo + sy.cta < n_units;  // CTA-uniform: warp group 0 of this CTA has a shared unit
  for (int round = 0; round < n_rounds + (shared_unit ? 1 : 0); ++round) {
    const bool coop = round == n_rounds;
    // The sharing factor W and this warp's share are derived twice (before and after the key loop, the second time from an
    // opaque copy of n_units so that the compiler does not keep them live across it): the persistent kernel has no register
    // to spare (DESIGN.md "Stack frames").
    // solo units are dealt round-robin over the CTAs (unit u of a round: CTA u % nc, warp u / nc): 1200 units occupy 8 - 9 warps
    // of EVERY SM instead of all 12 warps of the first 100, so the latency-bound key loops run on all SMs' load paths
    int u0 = round * n_warps + warp * sy.nc + sy.cta, k0 = 0, k1 = pos;
    bool first = true;
    if (coop) {
      const int W = sa_share(n_units, n_warps), wi = warp % W;
      // a warp group past the last unit repeats the last unit (identical values to identical addresses) so that the shared
      // round stays free of divergent paths around its block barriers
      u0 = min(n_units - 1, n_solo + (warp / W) * sy.nc + sy.cta);
      const int per = (pos + W - 1) / W;  // cached keys per warp
      k0 = min(pos, wi * per);
      k1 = min(pos, k0 + per);
      first = wi == 0;
    }
    if (u0 >= (coop ? n_units : n_solo)) continue;
    const int b = u0 / H, h = u0 - b * H;
    const SelfUnit u = self_unit_qkv(p, qkv_b, b, h, slot, c8);
    const size_t slab = ((size_t)s_rows[b] * H + h) * TX * 64;
    if (slot == 0 && first) {
      *reinterpret_cast<uint4*>(sk + slab + (size_t)pos * 64 + c8 * 8) = u.kq;
      *reinterpret_cast<uint4*>(sv + slab + (size_t)pos * 64 + c8 * 8) = u.vq;
    }
    float m, lsum, acc[8];
    self_unit_attend(u, sk + slab, sv + slab, k0, k1, first, slot, c8, stage, m, lsum, acc);
    if (!coop) {
      if (slot == 0) self_unit_store(p, b, h, c8, lsum, acc);
    } else {
      float* st = scratch + warp * 66;  // m, l, O[64] of this warp
      if (lane == 0) { st[0] = m; st[1] = lsum; }
      if (slot == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) st[2 + c8 * 8 + j] = acc[j];
      }
      __syncthreads();
      int nu = n_units;
      asm volatile("" : "+r"(nu));
      const int W = sa_share(nu, n_warps);
      if ((threadIdx.x >> 5) % W == 0 && slot == 0) {  // the group's first warp merges the W states that start at its own slot
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < MK_WARPS; ++w)
          if (w < W) M = fmaxf(M, st[w * 66]);
        float L = 0.f, o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
        for (int w = 0; w < MK_WARPS; ++w) {
          if (w < W) {
            const float mw = st[w * 66];
            const float wt = (mw > -INFINITY) ? __expf(mw - M) : 0.f;
            L += wt * st[w * 66 + 1];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += wt * st[w * 66 + 2 + c8 * 8 + j];
          }
        }
        self_unit_store(p, b, h, c8, L, o);
      }
      __syncthreads();  // the scratch is the attention scratch of the next cross-attention phase
    }
  }
}

// Cross-attention over the 1500 encoder positions.  n_slabs = B * H (sequence, head) slabs of 192 KB K + 192 KB V.
// Every CTA streams floor(n_slabs / G) whole slabs; the n_slabs % G remaining slabs are cut into pieces so that
// the tail is spread over all CTAs as well.  Warp 0 is a dedicated producer: K/V do not depend on q, so it keeps
// XA_NST stages of 112 keys (14 KB K + 14 KB V, 2-D TMA boxes with the 128-byte swizzle, L2 evict-first, completing
// on an mbarrier) in flight ACROSS work items and, for shallow rings, pulls the stages behind them into L2 with
// TMA prefetches; it also stages the raw q rows of upcoming items with cp.async.  The HBM stream never drains
// while a slab's states are merged or the next q is assembled.
//
// Math on the (otherwise idle) legacy tensor pipe, one query per head: consumer warp w owns keys 16 w .. 16 w + 15 of a stage.
//   S = K q      mma.m16n8k16: A = K rows (ldmatrix), B column 0 = bf16 hi part of the scaled q, column 1 = its
//                bf16 lo part (q - hi), so S = c0 + c1 carries ~16 mantissa bits of q; columns 2-7 are zero.
//   softmax      online over 16-key blocks (scores replicated per quad, max by shuffles over the quads).
//   O += V^T p   mma.m16n8k16: A = V^T (ldmatrix.trans, 16 dims x 16 keys), B column 0 / 1 = hi / lo part of p.
struct XaItem {
  int slab, k0, k1, piece, lj;
};
// Item `it` of CTA `cta`'s work list: its qw whole slabs, then its np = n_items - qw remainder pieces.  (Measured A/B, -DWXB_XA_PIECES_FIRST:
// with the pieces at the FRONT of the list their publish / ticket / merge round trips stall the merging consumer warp and with it the
// 4-stage ring while the stream should be running: 82.5 vs 77.4 us per phase at batch 60, 55.8 vs 45.6 at batch 30.)
__device__ __forceinline__ XaItem xa_item(int it, int cta, int qw, int G, int P, int plen, int np) {
  XaItem x;
#ifndef WXB_XA_PIECES_FIRST
  it = (it < qw) ? np + it : it - qw;  // position in the list -> (pieces 0 .. np-1, whole slabs np ..) numbering used below
#endif
  if (it >= np) {
    x.slab = (it - np) * G + cta; x.k0 = 0; x.k1 = T_AUDIO; x.piece = -1; x.lj = 0;
  } else {
    const int pc = it * G + cta;
    x.lj = pc / P; x.piece = pc - x.lj * P;
    x.slab = qw * G + x.lj; x.k0 = x.piece * plen; x.k1 = min(T_AUDIO, x.k0 + plen);
  }
  return x;
}
// K/V stages of CTA `cta`'s work list in one cross-attention phase.  Every phase of a launch has the same list, so the
// ring and item cursors of phase number xq (cross-attention phases completed since the kernel started) are xq times the
// per-phase counts: they are recomputed at the start of a phase instead of living in registers across all the others.
__device__ __forceinline__ uint32_t xa_stage_count(int cta, int qw, int G, int P, int plen, int n_items) {
  uint32_t n_st = (uint32_t)qw * ((T_AUDIO + XA_KEYS - 1) / XA_KEYS);
  for (int k = 0; k < n_items - qw; ++k) {
#ifdef WXB_XA_PIECES_FIRST
    const int it = k;
#else
    const int it = qw + k;  // the pieces sit behind the whole slabs
#endif
    const XaItem x = xa_item(it, cta, qw, G, P, plen, n_items - qw);
    n_st += (uint32_t)((x.k1 - x.k0 + XA_KEYS - 1) / XA_KEYS);
  }
  return n_st;
}
// bf16 hi / lo split of two floats, packed for an MMA B fragment: sel 0 -> (hi(x), hi(y)), 1 -> (lo(x), lo(y)), else 0.
// Branch-free (sel differs between the lanes of a warp).
__device__ __forceinline__ uint32_t split_pack(float x, float y, int sel) {
  const uint32_t hi = pack_bf16(x, y);  // x in the low half
  const float hx = __uint_as_float(hi << 16), hy = __uint_as_float(hi & 0xffff0000u);
  const uint32_t lo = pack_bf16(x - hx, y - hy);
  return sel == 0 ? hi : (sel == 1 ? lo : 0u);
}

// The warps never meet at a block barrier inside the phase: the producer warp stages the raw q rows (bias + split-K
// partials of the cq GEMV) of the next items, every consumer warp sums them for itself, deposits its (m, l, O) state of
// a finished item in a double-buffered shared-memory slot and moves straight on; consumer warp (item % 7) merges the
// 7 states once all have arrived (mbarrier) and writes the output.
__device__ __forceinline__ void cross_attn_phase(const MkParams& p, int l, int xq, int pos, const float* __restrict__ cq_b, uint8_t* ring,
                                                 float* scratch, MkSync& sy) {
  constexpr uint32_t STAGE = 2 * XA_HALF;
  float* sst = scratch;                                                                          // [2 item parities][7 warps][66]: m, l, O[64]
  float* qraw = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + SST_BYTES);      // [2 item parities][QRAW_ROWS][64]
  const float scale = p.scale;
  const int skip = p.skip;
  const float* __restrict__ part_q = p.part;
  __nv_bfloat16* __restrict__ att = p.att;
  float* __restrict__ apart = p.apart;
  int* __restrict__ ticket = p.ticket;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int B = p.B, H = p.H, d = p.d, G = sy.nc, cta = sy.cta, gk = p.g_dd.gk;
  const CUtensorMap* kvmap = p.maps + (size_t)p.L * TM_PER_LAYER + TM_KV;  // full-stage boxes; kvmap + 1: tail boxes
  const int n_slabs = B * H, qw = n_slabs / G, r = n_slabs - qw * G;
  // rows of the [rows, 64] K/V tensor: the cache holds B0 (allocated) sequences per layer, slabs are addressed by the ORIGINAL row
  const int krow0 = (l * 2) * p.B0 * H * T_AUDIO, vrow0 = (l * 2 + 1) * p.B0 * H * T_AUDIO;
  int P = 0, plen = T_AUDIO;
  if (r > 0) {
    const int want = (G + r - 1) / r;
    plen = (T_AUDIO + want - 1) / want;
    P = (T_AUDIO + plen - 1) / plen;
  }
  const int n_items = qw + ((r > 0 && cta < r * P) ? ((r * P - 1 - cta) / G + 1) : 0);
  const uint32_t xa_count0 = (uint32_t)xq * xa_stage_count(cta, qw, G, P, plen, n_items), xa_items0 = (uint32_t)xq * (uint32_t)n_items;
  if (warp == 0) {
    // ------------------------------- producer warp -------------------------------
    int it = 0, kk = 0;  // load cursor
    XaItem x = {};
    int oslab = 0;  // K/V slab of the item's ORIGINAL row
    auto orig_slab = [&](int slab) { const int b = slab / H; return s_rows[b] * H + (slab - b * H); };
    if (n_items > 0) { x = xa_item(0, cta, qw, G, P, plen, n_items - qw); kk = x.k0; oslab = orig_slab(x.slab); }
    uint32_t issued = xa_count0;
    auto stage_q = [&](int qi) {
      // raw q rows of item qi (row gk = bias, rows 0 .. gk-1 = split-K partials of the cq GEMV), 2 rows per pass
      const XaItem xq = xa_item(qi, cta, qw, G, P, plen, n_items - qw);
      const uint32_t gi = xa_items0 + (uint32_t)qi, qpar = gi & 1;
      mbar_wait(sy.mb(MB_Q_FREE + qpar), ((gi >> 1) & 1) ^ 1);  // the consumers have used the rows of item gi - 2
      float* dst = qraw + qpar * (QRAW_ROWS * 64);
      const int b = xq.slab / H, h = xq.slab - b * H;
      const int half = lane >> 4, l16 = lane & 15;
      for (int r0 = 0; r0 <= gk; r0 += 2) {
        const int row = r0 + half;
        if (row <= gk) {
          const float* src = (row == gk) ? (cq_b + h * 64) : (part_q + ((size_t)row * B + b) * d + h * 64);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + row * 64 + l16 * 4)), "l"(src + l16 * 4) : "memory");
        }
      }
      cp_async_mbar_arrive(sy.mb(MB_Q_FULL + qpar));
    };
    while (it < n_items) {
      if (kk == x.k0) stage_q(it);  // first stage of an item: its raw q rows
      if (lane == 0 && (int)(issued - xa_count0) >= sy.pre) {  // the first sy.pre stages were requested before the barrier wait
        const uint32_t sl = issued % XA_NST, par = (issued / XA_NST) & 1;
        const bool tail = (x.k1 - kk <= XA_TAIL);  // keys past the item (or the tensor: zero-filled) are masked by the consumer
        const CUtensorMap* m = tail ? kvmap + 1 : kvmap;
        uint8_t* dst = ring + (size_t)sl * STAGE;
        const int slab_ld = WXB_SKIP(skip, 64) ? (oslab & 3) : oslab;  // probe: every CTA streams the same 4 slabs (all L2 hits)
        const uint32_t half_tx = tail ? XA_TAIL * 128 : XA_HALF;
        mbar_wait(sy.mb(MB_XK_EMPTY + sl), par ^ 1);
        mbar_arrive_expect_tx(sy.mb(MB_XK_FULL + sl), half_tx);
        tma_load_2d_hint(dst, m, sy.mb(MB_XK_FULL + sl), 0, krow0 + slab_ld * T_AUDIO + kk, L2_EVICT_FIRST);
        mbar_wait(sy.mb(MB_XV_EMPTY + sl), par ^ 1);
        mbar_arrive_expect_tx(sy.mb(MB_XV_FULL + sl), half_tx);
        tma_load_2d_hint(dst + XA_HALF, m, sy.mb(MB_XV_FULL + sl), 0, vrow0 + slab_ld * T_AUDIO + kk, L2_EVICT_FIRST);
      }
      __syncwarp();
      ++issued;
      kk += XA_KEYS;
      if (kk >= x.k1) {
        ++it;
        if (it < n_items) { x = xa_item(it, cta, qw, G, P, plen, n_items - qw); kk = x.k0; oslab = orig_slab(x.slab); }
      }
    }
    sy.pre = 0;
  } else {
    // ------------------------------- consumer warps -------------------------------
    const int cw = warp - 1;
    // ldmatrix lane addressing inside a 112-row x 128-byte swizzled tile (chunk' = chunk ^ (row & 7)); this warp's rows 16 cw ..
    // K (non-transposed): row 16 cw + (lane & 7) + 8 ((lane >> 3) & 1), 16-byte chunk 2 j + (lane >> 4); V (transposed): row
    // 16 cw + (lane & 7) + 8 ((lane >> 4) & 1), chunk 2 mt + ((lane >> 3) & 1).  With sw = lane & 7 the swizzled chunk is
    // (2 j + c) ^ sw = ((j ^ (sw >> 1)) << 1) | (c ^ (sw & 1)): the lane's byte offset for j = 0 is kept, fragment j is at
    // (stage base + offset) ^ (j << 5) (stage bases are multiples of 1024, so bits 4 .. 6 come from the chunk alone).
    const int sw = lane & 7;
    const uint32_t laneK = (uint32_t)((cw * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * 128 + ((((lane >> 4) & 1) ^ sw) << 4));
    const uint32_t laneV = (uint32_t)((cw * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * 128 + ((((lane >> 3) & 1) ^ sw) << 4));
    uint32_t consumed = xa_count0;
    for (int it = 0; it < n_items; ++it) {
      float qn0, qn1;
      int xk0, xk1;  // key range of the item
      {
      const XaItem x = xa_item(it, cta, qw, G, P, plen, n_items - qw);
      xk0 = x.k0; xk1 = x.k1;
      const int b = x.slab / H, h = x.slab - b * H;
      const uint32_t gi = xa_items0 + (uint32_t)it;  // items since kernel start: parity and phase of the double-buffered slots
      const uint32_t ipar = gi & 1, iph = (gi >> 1) & 1;
      // scaled q = (bias + split-K partials in slice order) * scale; lane holds dims lane and lane + 32
      mbar_wait(sy.mb(MB_Q_FULL + ipar), iph);
      {
        const float* qr = qraw + ipar * (QRAW_ROWS * 64);
        float a0 = qr[gk * 64 + lane], a1 = qr[gk * 64 + lane + 32];
        for (int ks = 0; ks < gk; ++ks) { a0 += qr[ks * 64 + lane]; a1 += qr[ks * 64 + lane + 32]; }
        qn0 = a0 * scale;
        qn1 = a1 * scale;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sy.mb(MB_Q_FREE + ipar));
      if (p.qlog != nullptr && cw == 0 && x.piece <= 0) {
        // an alignment head: the scaled query of this position is what the word-timing DTW multiplies with the cross K later
        const int a = p.qhead[l * H + h];
        if (a >= 0) {
          float* ql = p.qlog + (((size_t)s_rows[b] * p.TX + pos) * p.qA + a) * 64;
          ql[lane] = qn0;
          ql[lane + 32] = qn1;
        }
      }
      }
      // B fragments of q: lane (g, t) holds elements 16 j + 2 t + {0, 1} and + {8, 9}; column g = 0 hi part, g = 1 lo part.
      // The scores are kept in the base-2 domain (q carries log2 e), so a probability is one ex2 of a difference.
      uint32_t qb[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float src = ((j < 2) ? qn0 : qn1) * 1.4426950408889634f;
        const int e = (16 * j + 2 * t) & 31;
        const float v0 = __shfl_sync(0xffffffffu, src, e), v1 = __shfl_sync(0xffffffffu, src, e + 1);
        const float v8 = __shfl_sync(0xffffffffu, src, e + 8), v9 = __shfl_sync(0xffffffffu, src, e + 9);
        qb[j][0] = split_pack(v0, v1, g);
        qb[j][1] = split_pack(v8, v9, g);
      }
      float m = -INFINITY, lsum = 0.f, o[4][4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f;
      // Two K/V stages per iteration (one for an odd last stage): the consumer is bound by the dependent-issue latency of ONE warp
      // per scheduler (ldmatrix -> HMMA -> vote -> ex2 -> shuffles -> pack -> HMMA), so the chains of two 16-key blocks are
      // interleaved instruction by instruction and the barrier waits, vote and loop bookkeeping are shared.  The K half of a stage
      // is released as soon as its fragments are in registers, the V half after its own ldmatrix (separate barriers): a warp
      // holds shared memory only for the duration of an ldmatrix, so the ring stays full although two stages are consumed at a time.
      // Scores / probabilities live in the t = 0 lane of every quad (B columns 2 .. 7 are zero: the other lanes compute 0 and are
      // masked out of the votes and reductions; the P fragment is gathered from the t = 0 lanes).
      // The running offset m is only moved when a score exceeds it by more than 2^XA_LAZY (warp vote): most iterations need no
      // max reduction and no rescale at all; probabilities are then at most 2^XA_LAZY.
      constexpr float XA_LAZY = 8.f;
      const uint32_t ring_a = smem_u32(ring);
      int kk = xk0;
      auto stages = [&](auto ns_tag) {
        constexpr int NS = decltype(ns_tag)::value;
        uint32_t sl[NS];
        float c[NS][4];
        // Blocks past the item's end (last stage only) may lie outside the short TMA box: what is read there are stale bf16 values
        // of earlier stages / operators (the ring is zeroed when the kernel starts, so never uninitialised bits): their scores are
        // masked and 0 x finite adds nothing to O.
        {
          uint32_t ka[NS][4][4];
#pragma unroll
          for (int n = 0; n < NS; ++n) {
            const uint32_t cn = consumed + (uint32_t)n;
            sl[n] = cn % XA_NST;
            mbar_wait(sy.mb(MB_XK_FULL + sl[n]), (cn / XA_NST) & 1);
            const uint32_t a0 = ring_a + sl[n] * STAGE + laneK;
#pragma unroll
            for (int j = 0; j < 4; ++j) ldsm_x4(ka[n][j], a0 ^ (uint32_t)(j << 5));
          }
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int n = 0; n < NS; ++n) mbar_arrive(sy.mb(MB_XK_EMPTY + sl[n]));  // this warp holds its K fragments
          }
#pragma unroll
          for (int n = 0; n < NS; ++n) c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.f;
          if (!WXB_SKIP(skip, 16)) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int n = 0; n < NS; ++n) mma_16816(c[n], ka[n][j], qb[j][0], qb[j][1]);
          }
        }
        // V fragments: requested before the softmax arithmetic so that their latency is covered by it
        uint32_t va[NS][4][4];
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          mbar_wait(sy.mb(MB_XV_FULL + sl[n]), ((consumed + (uint32_t)n) / XA_NST) & 1);
          const uint32_t a1 = ring_a + sl[n] * STAGE + XA_HALF + laneV;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) ldsm_x4_t(va[n][mt], a1 ^ (uint32_t)(mt << 5));
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int n = 0; n < NS; ++n) mbar_arrive(sy.mb(MB_XV_EMPTY + sl[n]));  // this warp is done with the stages
        }
        // scores of this warp's keys 16 cw + g (s0) and + 8 (s1) of every stage; keys past the item are masked
        float s0[NS], s1[NS];
        float mx = -INFINITY;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          const int lim = xk1 - (kk + n * XA_KEYS + cw * 16);  // valid keys of this warp's block (<= 0: none)
          const bool mine = (t == 0) && !WXB_SKIP(skip, 16);
          s0[n] = (mine && g < lim) ? c[n][0] + c[n][1] : -INFINITY;
          s1[n] = (mine && g + 8 < lim) ? c[n][2] + c[n][3] : -INFINITY;
          mx = fmaxf(mx, fmaxf(s0[n], s1[n]));
        }
        if (__any_sync(0xffffffffu, mx > m + XA_LAZY)) {  // m = -inf: any finite score
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
          mx = __shfl_sync(0xffffffffu, mx, 0);
          const float alpha = exp2f(m - mx);  // m = -inf -> 0
          lsum *= alpha;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) { o[mt][0] *= alpha; o[mt][1] *= alpha; o[mt][2] *= alpha; o[mt][3] *= alpha; }
          m = mx;
        }
        if (m > -INFINITY) {  // warp-uniform; false only while every key seen so far was masked
          uint32_t pb[NS][2];
#pragma unroll
          for (int n = 0; n < NS; ++n) {
            const float p0 = ex2_approx(s0[n] - m), p1 = ex2_approx(s1[n] - m);  // -inf -> 0
            lsum += p0 + p1;  // meaningful in the t = 0 lanes
            // B fragment of p: lane (g, t) needs keys 2t, 2t+1 (b0) and 2t+8, 2t+9 (b1): the t = 0 lanes of quads 2t and 2t+1
            const float x0 = __shfl_sync(0xffffffffu, p0, 8 * t), x1 = __shfl_sync(0xffffffffu, p0, 8 * t + 4);
            const float y0 = __shfl_sync(0xffffffffu, p1, 8 * t), y1 = __shfl_sync(0xffffffffu, p1, 8 * t + 4);
            pb[n][0] = split_pack(x0, x1, g);
            pb[n][1] = split_pack(y0, y1, g);
          }
          if (!WXB_SKIP(skip, 128)) {
#pragma unroll
          for (int n = 0; n < NS; ++n)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) mma_16816(o[mt], va[n][mt], pb[n][0], pb[n][1]);
          }
        }
        consumed += NS;
        kk += NS * XA_KEYS;
      };
      while (kk + XA_KEYS < xk1) stages(std::integral_constant<int, 2>{});
      if (kk < xk1) stages(std::integral_constant<int, 1>{});
      // ---- deposit this warp's state; consumer warp (item % XA_CW) merges the states and writes the output ----
      // (the item's coordinates are derived again from an opaque copy of its index: nothing but the softmax state is carried
      // through the key loop in registers, see DESIGN.md "Stack frames")
      int it2 = it;
      asm volatile("" : "+r"(it2));
      const XaItem x = xa_item(it2, cta, qw, G, P, plen, n_items - qw);
      const int b = x.slab / H, h = x.slab - b * H;
      const uint32_t gi = xa_items0 + (uint32_t)it2, ipar = gi & 1, iph = (gi >> 1) & 1;
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
      mbar_wait(sy.mb(MB_ST_FREE + ipar), iph ^ 1);  // the merge of item gi - 2 has released this slot
      {
        float* st = sst + (ipar * XA_CW + cw) * 66;
        if (lane == 0) { st[0] = m; st[1] = lsum; }
        if (t == 0) {
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            st[2 + 16 * mt + g] = o[mt][0] + o[mt][1];
            st[2 + 16 * mt + g + 8] = o[mt][2] + o[mt][3];
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sy.mb(MB_ST_FULL + ipar));
      if (cw == (int)(gi % XA_CW)) {
        mbar_wait(sy.mb(MB_ST_FULL + ipar), iph);
        const float* st = sst + (size_t)ipar * XA_CW * 66;
        float M = -INFINITY;
#pragma unroll
        for (int i = 0; i < XA_CW; ++i) M = fmaxf(M, st[i * 66]);
        float Ls = 0.f, o0 = 0.f, o1 = 0.f;  // dims lane and lane + 32
#pragma unroll
        for (int i = 0; i < XA_CW; ++i) {
          const float w = exp2f(st[i * 66] - M);  // (offsets are base-2) a warp that saw no key of the item: m = -inf -> 0
          Ls += w * st[i * 66 + 1];
          o0 += w * st[i * 66 + 2 + lane];
          o1 += w * st[i * 66 + 2 + lane + 32];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sy.mb(MB_ST_FREE + ipar));
        __nv_bfloat16* out = att + (size_t)b * d + h * 64;
        if (WXB_SKIP(skip, 32)) {
        } else if (x.piece < 0) {
          out[lane] = __float2bfloat16_rn(o0 / Ls);
          out[lane + 32] = __float2bfloat16_rn(o1 / Ls);
        } else {
          // ---- piece of a remainder slab: publish the state, the last-arriving piece merges all of them ----
          float* part = apart + (size_t)x.lj * P * 66;
          part[x.piece * 66 + 2 + lane] = o0;
          part[x.piece * 66 + 2 + lane + 32] = o1;
          if (lane == 0) { part[x.piece * 66] = M; part[x.piece * 66 + 1] = Ls; }
          __threadfence();
          __syncwarp();
          int last = 0;
          if (lane == 0) {
            const int prev = atomicAdd(ticket + x.lj, 1);
            last = (prev == P - 1);
            if (last) ticket[x.lj] = 0;
          }
          last = __shfl_sync(0xffffffffu, last, 0);
          if (last) {
            __threadfence();
            // the P piece states in groups of XA_MG pieces, every load of a group in flight at once and the running offset
            // carried across groups: ONE L2 round trip per group (P <= XA_MG: one in all).  A separate pass over the maxima
            // plus groups of four were 4-5 dependent round trips at the very end of the phase: 4.4 us (batch 60) to 5.1 us
            // (batch 8) of every layer (tools/xa_fixed_probe.sh).
            float MM = -INFINITY, LL = 0.f, O0 = 0.f, O1 = 0.f;
            for (int s0 = 0; s0 < P; s0 += XA_MG) {
              float mv[XA_MG], lv[XA_MG], a0[XA_MG], a1[XA_MG];
#pragma unroll
              for (int u = 0; u < XA_MG; ++u) {
                const bool in = s0 + u < P;
                const float* ps = part + (in ? s0 + u : s0) * 66;
                mv[u] = in ? __ldcg(ps) : -INFINITY;
                lv[u] = __ldcg(ps + 1);
                a0[u] = __ldcg(ps + 2 + lane);
                a1[u] = __ldcg(ps + 2 + lane + 32);
              }
              float gm = mv[0];
#pragma unroll
              for (int u = 1; u < XA_MG; ++u) gm = fmaxf(gm, mv[u]);
              const float Mn = fmaxf(MM, gm);
              const float sc = (Mn == MM) ? 1.f : exp2f(MM - Mn);  // first group: 2^-inf = 0 on zeros
              LL *= sc; O0 *= sc; O1 *= sc;
              MM = Mn;
#pragma unroll
              for (int u = 0; u < XA_MG; ++u) {
                const float w = (mv[u] > -INFINITY) ? exp2f(mv[u] - MM) : 0.f;  // a slot past P
                LL += w * lv[u];
                O0 += w * a0[u];
                O1 += w * a1[u];
              }
            }
            out[lane] = __float2bfloat16_rn(O0 / LL);
            out[lane + 32] = __float2bfloat16_rn(O1 / LL);
          }
        }
      }
    }
  }
}

// mlx_whisper_batch_decoder.py:267-303 for one row per CTA: (no_speech_prob from the unfiltered logits,) filters, argmax
// (first max), logprob accounting, EOT latch.  The logits row is read once with 16-byte loads (rows are 16-byte aligned:
// ldl is a multiple of 4), 4 loads in flight per thread.
//
// Timestamp rules (sp.ts_rules, decoding with `without_timestamps=False`), per row, from its sampled tokens seq:
//   <|notimestamps|> is suppressed; after a timestamp: a second one forbids timestamps, a single one forbids text below eot;
//   timestamps never decrease (the open half of a pair may repeat, a closed pair must move on); the first sampled token
//   must be a timestamp <= max_initial_ts; and if logsumexp(timestamp logprobs) > max(text logprob) only timestamps remain
//   (/root/reference/mlx_ultra_optimized_batch.py:38-71).  Range rules are predicates on the token id (nothing is written
//   back); the pass keeps (max, first argmax, sum of exponentials) separately for ids below / from ts_begin.
struct RowStat {
  float m;  // running maximum
  int i;    // first index of the maximum
  float s;  // sum of exp(x - m)
};
__device__ __forceinline__ void stat_add(RowStat& a, float v, int idx) {
  if (v > a.m) {
    a.s = a.s * __expf(a.m - v) + 1.f;  // a.m = -inf -> s = 0
    a.m = v; a.i = idx;
  } else if (v > -INFINITY) {
    a.s += __expf(v - a.m);
  }
}
__device__ __forceinline__ void stat_merge(RowStat& a, float om, int oi, float os) {
  if (om > a.m || (om == a.m && oi < a.i)) {
    a.s = (a.m > -INFINITY ? a.s * __expf(a.m - om) : 0.f) + os;
    a.m = om; a.i = oi;
  } else if (om > -INFINITY) {
    a.s += os * __expf(om - a.m);
  }
}
__device__ __forceinline__ void stat_block(RowStat& a, float* red, int* red_i, float* red_s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, a.m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
    const float os = __shfl_xor_sync(0xffffffffu, a.s, o);
    stat_merge(a, ov, oi, os);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = a.m; red_i[threadIdx.x >> 5] = a.i; red_s[threadIdx.x >> 5] = a.s; }
  __syncthreads();
  a.m = red[0]; a.i = red_i[0]; a.s = red_s[0];
#pragma unroll
  for (int w = 1; w < MK_WARPS; ++w) stat_merge(a, red[w], red_i[w], red_s[w]);
}

__device__ __forceinline__ void sample_phase(const SampleParams& p, int B, int pos, bool do_sample, float* red, int* red_i,
                                             const MkSync& sy) {
  __shared__ float red_s[MK_WARPS];
  const int tid = threadIdx.x;
  constexpr int U = 2;  // float4 loads in flight per thread (3 or more cost the persistent kernel a stack frame)
  const int V4 = (p.V + 3) >> 2;  // the padding elements of the last group are masked by index
  // A row is scanned by R CTAs, each over its own slice of the vocabulary (the scan is a latency-bound loop of L2 round trips:
  // 17 iterations for one CTA at V = 51866, whatever the batch; with 8 rows on 148 CTAs R = 18 and it is one).  Slice statistics
  // go through `part_stats`, the last CTA to arrive at the row's counter merges them (slice order) and samples.  The first step
  // (no_speech_prob needs the unfiltered row) and the stand-alone sampling kernel keep one CTA per row.
  const int R = (p.part_stats && do_sample && !p.nsp_out) ? max(1, sy.nc / B) : 1;
  for (int unit = sy.cta; unit < B * R; unit += sy.nc) {
    const int b = unit / R, part = unit - b * R;
    const int g0 = (int)((long long)V4 * part / R), g1 = (int)((long long)V4 * (part + 1) / R);  // float4 groups of this slice
    float* x = p.logits + (size_t)b * p.ldl;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int ob = s_rows[b];
    if (p.nsp_out) {
      RowStat a = {-INFINITY, 0x7fffffff, 0.f};
      for (int i0 = tid; i0 < V4; i0 += MK_THREADS * U) {
        float4 t[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { const int i = i0 + MK_THREADS * j; t[j] = i < V4 ? __ldcg(x4 + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); }
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int i = 4 * (i0 + MK_THREADS * j);
          stat_add(a, t[j].x, i);
          if (i + 1 < p.V) stat_add(a, t[j].y, i + 1);
          if (i + 2 < p.V) stat_add(a, t[j].z, i + 2);
          if (i + 3 < p.V) stat_add(a, t[j].w, i + 3);
        }
      }
      stat_block(a, red, red_i, red_s);
      if (tid == 0) p.nsp_out[ob] = __expf(__ldcg(x + p.nsp_token) - a.m) / a.s;
    }
    if (!do_sample) continue;
    __syncthreads();
    // static filters, applied in place by the CTA whose slice holds the id (no CTA reads another slice)
    const int e0 = 4 * g0, e1 = min(4 * g1, p.V);
    for (int i = tid; i < p.n_suppress; i += MK_THREADS) {
      const int id = p.suppress[i];
      if (id >= e0 && id < e1) x[id] = -INFINITY;
    }
    if (p.suppress_blank && pos == p.prompt_len - 1 && tid == 0) {
      if (p.blank_token >= e0 && p.blank_token < e1) x[p.blank_token] = -INFINITY;
      if (p.eot >= e0 && p.eot < e1) x[p.eot] = -INFINITY;
    }
    // range rules of this row: ids < lo_text are masked, timestamps in [ts_begin, ts_lo) and ids > ts_hi are masked
    int lo_text = 0, ts_lo = p.V, ts_hi = p.V - 1, tsb = p.V;
    const int* row_tok = p.tokens + (size_t)ob * p.stride;
    if (p.ts_rules) {
      tsb = p.ts_begin; ts_lo = tsb;
      if (tid == 0 && p.no_timestamps >= e0 && p.no_timestamps < e1) x[p.no_timestamps] = -INFINITY;
      const int n_seq = pos + 1 - p.prompt_len;  // sampled tokens so far
      const bool last_ts = n_seq >= 1 && __ldcg(row_tok + pos) >= tsb;
      const bool pen_ts = n_seq < 2 || __ldcg(row_tok + pos - 1) >= tsb;
      if (last_ts) {
        if (pen_ts) ts_lo = p.V;   // a closed pair: no timestamp may follow
        else lo_text = p.eot;      // an open timestamp: no text token below eot may follow
      }
      const int tl = __ldcg(p.ts_last + ob);
      if (tl >= 0) {
        const int first_ok = (last_ts && !pen_ts) ? tl : tl + 1;  // timestamps must not decrease
        ts_lo = max(ts_lo, first_ok);
      }
      if (n_seq == 0) {
        lo_text = tsb;  // the first sampled token is a timestamp ...
        if (p.max_initial_ts >= 0) ts_hi = min(ts_hi, tsb + p.max_initial_ts);  // ... not later than max_initial_timestamp
      }
    }
    __syncthreads();
    RowStat tx = {-INFINITY, 0x7fffffff, 0.f}, tt = {-INFINITY, 0x7fffffff, 0.f};  // ids below / from ts_begin
    if (!p.ts_rules) {
      // no range rule: one running (max, argmax, sum) over the slice
      for (int i0 = g0 + tid; i0 < g1; i0 += MK_THREADS * U) {
        float4 t[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { const int i = i0 + MK_THREADS * j; t[j] = i < g1 ? __ldcg(x4 + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); }
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int i = 4 * (i0 + MK_THREADS * j);  // increasing within the thread: strict > keeps the first maximum
          stat_add(tx, t[j].x, i);
          if (i + 1 < p.V) stat_add(tx, t[j].y, i + 1);
          if (i + 2 < p.V) stat_add(tx, t[j].z, i + 2);
          if (i + 3 < p.V) stat_add(tx, t[j].w, i + 3);
        }
      }
    } else {
      for (int i0 = g0 + tid; i0 < g1; i0 += MK_THREADS * U) {
        float4 t[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { const int i = i0 + MK_THREADS * j; t[j] = i < g1 ? __ldcg(x4 + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); }
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int i = 4 * (i0 + MK_THREADS * j);
          const float e[4] = {t[j].x, t[j].y, t[j].z, t[j].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int id = i + q;
            if (id < tsb) {
              if (id >= lo_text) stat_add(tx, e[q], id);
            } else if (id >= ts_lo && id <= ts_hi) {
              stat_add(tt, e[q], id);
            }
          }
        }
      }
    }
    stat_block(tx, red, red_i, red_s);
    if (p.ts_rules) stat_block(tt, red, red_i, red_s);
    if (tid == 0) {
      bool mine = true;  // this CTA samples the row
      if (R > 1) {
        float* ps = p.part_stats + (size_t)(b * R + part) * 8;
        ps[0] = tx.m; ps[1] = __int_as_float(tx.i); ps[2] = tx.s;
        ps[3] = tt.m; ps[4] = __int_as_float(tt.i); ps[5] = tt.s;
        __threadfence();
        mine = atomicAdd(p.part_ticket + b, 1) == R - 1;
        if (mine) {
          __threadfence();
          p.part_ticket[b] = 0;  // self-cleaning
          tx = {-INFINITY, 0x7fffffff, 0.f}; tt = {-INFINITY, 0x7fffffff, 0.f};
          for (int r = 0; r < R; ++r) {
            const float* q = p.part_stats + (size_t)(b * R + r) * 8;
            stat_merge(tx, __ldcg(q), __float_as_int(__ldcg(q + 1)), __ldcg(q + 2));
            if (p.ts_rules) stat_merge(tt, __ldcg(q + 3), __float_as_int(__ldcg(q + 4)), __ldcg(q + 5));
          }
        }
      }
      if (mine) {
        float best = tx.m, ssum = tx.s;
        int bi = tx.i;
        if (p.ts_rules && tt.m > -INFINITY) {
          // logsumexp over timestamps vs the best text logprob (the common normaliser cancels): timestamps win -> only they remain
          const bool ts_only = !(tx.m > -INFINITY) || (tt.m + __logf(tt.s) > tx.m);
          if (ts_only) { best = tt.m; bi = tt.i; ssum = tt.s; }
          else {
            RowStat all = tx;
            stat_merge(all, tt.m, tt.i, tt.s);
            best = all.m; bi = all.i; ssum = all.s;
          }
        }
        const float logprob = -logf(ssum);  // x[bi] - (best + log(sum)) with x[bi] == best
        int* row = p.tokens + (size_t)ob * p.stride;
        const int last = __ldcg(row + pos);
        const bool was_eot = (last == p.eot) && (pos >= p.prompt_len);  // prompt tokens never latch
        if (!was_eot) p.sum_logprob[ob] += logprob;
        const int next = was_eot ? p.eot : bi;
        row[pos + 1] = next;
        if (next == p.eot) p.done[ob] = 1;
        if (p.ts_rules && next >= p.ts_begin) p.ts_last[ob] = next;
        (void)best;
      }
    }
    __syncthreads();
  }
}

// phase kinds of the step schedule: 11 per layer, then final LN | logits | sampling
enum { PH_LN = 0, PH_GEMV = 1, PH_SELF = 2, PH_CROSS = 3, PH_SAMPLE = 4 };

__device__ __forceinline__ int op_kind(int k) {
  return (k == 0 || k == 4 || k == 8 || k == 11) ? PH_LN : (k == 2) ? PH_SELF : (k == 6) ? PH_CROSS : (k == 13) ? PH_SAMPLE : PH_GEMV;
}
// operands of GEMV operator k of layer l
struct GemvOp {
  const MkGemv* g;
  const CUtensorMap *wm, *xm;
  int epi;
};
template <int MT>
__device__ __forceinline__ GemvOp gemv_op(const MkParams& p, int l, int k) {
  const CUtensorMap* lm = p.maps + (size_t)l * TM_PER_LAYER;
  const CUtensorMap* am = p.maps + (size_t)p.L * TM_PER_LAYER + TM_ACT + (MT - 1) * 3 - 1;  // am[0] unused here, am[1..3] = xn, att, hid of this MT
  const CUtensorMap* emb = p.maps + (size_t)p.L * TM_PER_LAYER + TM_EMB;
  GemvOp o;
  o.g = (k == 1) ? &p.g_qkv : (k == 9) ? &p.g_fc1 : (k == 10) ? &p.g_fc2 : (k == 12) ? &p.g_logits : &p.g_dd;
  o.wm = (k == 1) ? lm + TM_QKV : (k == 3) ? lm + TM_OUT : (k == 5) ? lm + TM_CQ : (k == 7) ? lm + TM_COUT
         : (k == 9) ? lm + TM_FC1 : (k == 10) ? lm + TM_FC2 : emb;
  o.xm = (k == 3 || k == 7) ? am + 2 : (k == 10) ? am + 3 : am + 1;
  o.epi = (k == 9) ? EPI_GELU_BF16 : (k == 12) ? EPI_LOGITS : EPI_PART;
  return o;
}

// While a grid barrier completes (grid_sync): request the first tiles of operator (l, k) that do not depend on the
// operator just finished: the weight halves of a GEMV's first ring stages, or the first K/V stages of a
// cross-attention phase.  Every thread calls this and learns the number of stages (sy.pre, skipped by the producer
// loops of those phases); only `issue` threads touch the barriers and the TMA unit.
template <int MT>
__device__ __forceinline__ void pre_issue(const MkParams& p, int l, int k, int xq, uint8_t* ring, MkSync& sy,
                                          const bool issue) {
  const int kind = op_kind(k);
  sy.pre = 0;
  if (kind == PH_GEMV && !WXB_SKIP(p.skip, 2)) {
    constexpr int STAGE = GV_A_BYTES + 16 * MT * GV_BK * 2;
    const GemvOp o = gemv_op<MT>(p, l, k);
    const MkGemv& g = *o.g;
    const int tile = sy.cta;
    if (tile >= g.tiles) return;
    const int Ks = g.K / g.gk, nkb = Ks / GV_BK;
    const int k0 = (tile % g.gk) * Ks, row0 = (tile / g.gk) * GV_ROWS;
    const int n = nkb < GV_NST ? nkb : GV_NST;
    for (int kb = 0; issue && kb < n; ++kb) {
      const uint32_t c = sy.gv_count + (uint32_t)kb, slot = c % GV_NST, par = (c / GV_NST) & 1;
      mbar_wait(sy.mb(MB_GV_EMPTY + slot), par ^ 1);
      mbar_arrive_expect_tx(sy.mb(MB_GV_FULL + slot), STAGE);
      tma_load_2d(ring + slot * STAGE, o.wm, sy.mb(MB_GV_FULL + slot), k0 + kb * GV_BK, row0);
    }
    sy.pre = n;
  } else if (kind == PH_CROSS && !WXB_SKIP(p.skip, 1)) {
    constexpr uint32_t STAGE = 2 * XA_HALF;
    const int G = sy.nc, cta = sy.cta;
    const CUtensorMap* kvmap = p.maps + (size_t)p.L * TM_PER_LAYER + TM_KV;
    const int n_slabs = p.B * p.H, qw = n_slabs / G, r = n_slabs - qw * G;
    const int krow0 = (l * 2) * p.B0 * p.H * T_AUDIO, vrow0 = (l * 2 + 1) * p.B0 * p.H * T_AUDIO;
    int P = 0, plen = T_AUDIO;
    if (r > 0) {
      const int want = (G + r - 1) / r;
      plen = (T_AUDIO + want - 1) / want;
      P = (T_AUDIO + plen - 1) / plen;
    }
    const int n_items = qw + ((r > 0 && cta < r * P) ? ((r * P - 1 - cta) / G + 1) : 0);
    const uint32_t xa_count0 = (uint32_t)xq * xa_stage_count(cta, qw, G, P, plen, n_items);
    int it = 0, n = 0;
    while (it < n_items && n < XA_NST) {
      const XaItem x = xa_item(it, cta, qw, G, P, plen, n_items - qw);
      const int xb = x.slab / p.H, oslab = s_rows[xb] * p.H + (x.slab - xb * p.H);
      for (int kk = x.k0; kk < x.k1 && n < XA_NST; kk += XA_KEYS, ++n) {
        if (!issue) continue;
        const uint32_t c = xa_count0 + (uint32_t)n, sl = c % XA_NST, par = (c / XA_NST) & 1;
        const bool tail = (x.k1 - kk <= XA_TAIL);
        const CUtensorMap* m = tail ? kvmap + 1 : kvmap;
        uint8_t* dst = ring + (size_t)sl * STAGE;
        const uint32_t half_tx = tail ? XA_TAIL * 128 : XA_HALF;
        mbar_wait(sy.mb(MB_XK_EMPTY + sl), par ^ 1);
        mbar_arrive_expect_tx(sy.mb(MB_XK_FULL + sl), half_tx);
        tma_load_2d_hint(dst, m, sy.mb(MB_XK_FULL + sl), 0, krow0 + oslab * T_AUDIO + kk, L2_EVICT_FIRST);
        mbar_wait(sy.mb(MB_XV_EMPTY + sl), par ^ 1);
        mbar_arrive_expect_tx(sy.mb(MB_XV_FULL + sl), half_tx);
        tma_load_2d_hint(dst + XA_HALF, m, sy.mb(MB_XV_FULL + sl), 0, vrow0 + oslab * T_AUDIO + kk, L2_EVICT_FIRST);
      }
      ++it;
    }
    sy.pre = n;
  }
}


// One operator of the step schedule.
// k: 0 LN1 | 1 QKV | 2 self-attention | 3 out | 4 LN2 | 5 cq | 6 cross-attention | 7 cout | 8 LN3 | 9 fc1 | 10 fc2
//    11 final LN | 12 logits | 13 (no_speech_prob,) filters + sampling
template <int MT>
__device__ __forceinline__ void run_op(const MkParams& p, const DecLayerW* s_layers, int l, int k, int pos, int xq,
                                       uint8_t* ring, float* scratch, float* red, int* red_i, MkSync& sy) {
  const DecLayerW& w = s_layers[l];
  const int kind = op_kind(k);
#ifdef WXB_STUB
  constexpr int stub = WXB_STUB;
#else
  constexpr int stub = 0;
#endif
  if (kind == PH_LN) {
    if (!(stub & 8) && !WXB_SKIP(p.skip, 8)) {
      // the LayerNorm phase first folds the previous GEMV's split-K partials (+ bias) into the residual row
      const bool from_embed = (k == 0 && l == 0);
      const int gk = (k == 0 || k == 11) ? p.g_fc2.gk : p.g_dd.gk;
      const float* pb = (k == 0) ? (l > 0 ? s_layers[l - 1].fc2_b : nullptr) : (k == 4) ? w.out_b : (k == 8) ? w.cout_b : w.fc2_b;
      const float* lw = (k == 0) ? w.ln1_w : (k == 4) ? w.ln2_w : (k == 8) ? w.ln3_w : p.lnf_w;
      const float* lb = (k == 0) ? w.ln1_b : (k == 4) ? w.ln2_b : (k == 8) ? w.ln3_b : p.lnf_b;
      ln_phase(p, from_embed, gk, pb, lw, lb, pos, red, sy);
    }
  } else if (kind == PH_GEMV) {
    if (!(stub & 2) && !WXB_SKIP(p.skip, 2)) {
      const GemvOp o = gemv_op<MT>(p, l, k);
      gemv_phase<MT>(p, *o.g, o.wm, o.xm, w.fc1_b, o.epi, ring, sy);
    }
  } else if (kind == PH_SELF) {
    if (!(stub & 4) && !WXB_SKIP(p.skip, 4)) self_attn_phase(p, l, w.qkv_b, pos, ring, scratch, sy);
  } else if (kind == PH_CROSS) {
    if (!(stub & 1) && !WXB_SKIP(p.skip, 1)) cross_attn_phase(p, l, xq, pos, w.cq_b, ring, scratch, sy);
  } else {
    if (!(stub & 16)) sample_phase(p.sp, p.B, pos, p.mode == 2, red, red_i, sy);
  }
}

// operators per decode step
__device__ __forceinline__ int ops_per_step(const MkParams& p) {
  return 11 * p.L + (p.mode >= 1 ? 2 : 0) + ((p.mode == 2 || p.sp.nsp_out) ? 1 : 0);
}

// One CTA per SM (cooperative launch); every CTA walks the whole schedule, operators separated by grid barriers.
// (Running several CTAs per SM so that one sequence group's chain hides under another group's K/V stream was tried
// and measured slower; see DESIGN.md.  A kernel that allocates tensor memory is pinned to one CTA per SM anyway.)
template <int MT>
__global__ void __launch_bounds__(MK_THREADS, MK_CTAS_PER_SM) dec_step_kernel(const __grid_constant__ MkParams p) {
  extern __shared__ uint8_t mk_smem_raw[];
  uint8_t* ring = mk_smem_raw + ((1024u - (smem_u32(mk_smem_raw) & 1023u)) & 1023u);  // 1024-byte aligned (swizzle atoms)
  float* scratch = reinterpret_cast<float*>(ring + RING_BYTES);
  __shared__ float red[MK_WARPS];
  __shared__ int red_i[MK_WARPS];
  __shared__ __align__(8) uint64_t bars[MB_COUNT];
  __shared__ uint32_t tmem_slot;
  __shared__ DecLayerW s_layers[MAX_LAYERS];  // pointer table of every layer: no dependent global load per phase
  for (int i = threadIdx.x; i < p.L * (int)(sizeof(DecLayerW) / 8); i += MK_THREADS)
    reinterpret_cast<unsigned long long*>(s_layers)[i] = reinterpret_cast<const unsigned long long*>(p.layers)[i];
  for (int i = threadIdx.x; i < p.B; i += MK_THREADS) s_rows[i] = p.rows ? p.rows[i] : i;
  // the cross-attention consumers may read ring rows that no TMA box of the current stage has written (cross_attn_phase): such
  // rows must never hold uninitialised bits
  for (int i = threadIdx.x; i < RING_BYTES / 16; i += MK_THREADS) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
  const int warp = threadIdx.x >> 5;
  MkSync sy;
  sy.bars = smem_u32(bars);
  sy.gv_count = 0; sy.acc_count = 0; sy.pre = 0;
  sy.cta = blockIdx.x; sy.nc = gridDim.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < GV_NST; ++i) { mbar_init(sy.mb(MB_GV_FULL + i), 1); mbar_init(sy.mb(MB_GV_EMPTY + i), 1); }
    mbar_init(sy.mb(MB_ACC_FULL), 1);
    for (int i = 0; i < XA_NST; ++i) {
      mbar_init(sy.mb(MB_XK_FULL + i), 1); mbar_init(sy.mb(MB_XK_EMPTY + i), XA_CW);
      mbar_init(sy.mb(MB_XV_FULL + i), 1); mbar_init(sy.mb(MB_XV_EMPTY + i), XA_CW);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(sy.mb(MB_ST_FULL + i), XA_CW); mbar_init(sy.mb(MB_ST_FREE + i), 1);
      mbar_init(sy.mb(MB_Q_FULL + i), 32); mbar_init(sy.mb(MB_Q_FREE + i), XA_CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(&tmem_slot, GV_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  sy.tmem = tmem_slot;

  if (p.delay_ns) {  // CTA-uniform
    if (threadIdx.x == 0) {
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do { __nanosleep(500); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < (unsigned long long)p.delay_ns);
    }
    __syncthreads();
  }
  if (p.prof && sy.cta == 0 && threadIdx.x == 0) {  // tracing aid: where and when this instance's CTA 0 started
    unsigned smid;
    unsigned long long t;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.prof[(1 << 16) - 1] = smid;
    p.prof[(1 << 16) - 2] = t;
  }
  const int pos0 = *p.d_pos;  // written only after the last barrier of this launch
  const int n_ph = ops_per_step(p);
  for (int s = 0; s < p.n_steps; ++s) {
    const int pos = pos0 + s;
    for (int ph = 0; ph < n_ph; ++ph) {
      int l = ph / 11, k = ph - 11 * l;
      if (l >= p.L) { k = 11 + (ph - 11 * p.L); l = p.L - 1; }
      run_op<MT>(p, s_layers, l, k, pos, s * p.L + l, ring, scratch, red, red_i, sy);
      if (ph + 1 < n_ph || s + 1 < p.n_steps) {
        // the operator after the barrier: (l2, k2)
        const int ph2 = (ph + 1 < n_ph) ? ph + 1 : 0;
        int l2 = ph2 / 11, k2 = ph2 - 11 * l2;
        if (l2 >= p.L) { k2 = 11 + (ph2 - 11 * p.L); l2 = p.L - 1; }
        const int xq2 = (ph + 1 < n_ph ? s : s + 1) * p.L + l2;  // cross-attention phases completed before operator (l2, k2)
        grid_sync(p.bar, (unsigned)(s * n_ph + ph + 1), sy.nc, sy.cta, p.prof, [&]() { pre_issue<MT>(p, l2, k2, xq2, ring, sy, true); });
        if (warp == 0) pre_issue<MT>(p, l2, k2, xq2, ring, sy, false);  // the producer warp only needs the count
      }
    }
  }
  if (sy.cta == 0 && threadIdx.x == 0) *p.d_pos = pos0 + p.n_steps;
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(sy.tmem, GV_TMEM_COLS);
  }
}

// The sampling phase on its own (one CTA per row), for parity tests of the sampling rule on arbitrary logits
// (wxb_decoder_sample): exactly the code the persistent kernel runs after the logits GEMV.
__global__ void __launch_bounds__(MK_THREADS, 1) dec_sample_kernel(const SampleParams sp, int B, int pos, int do_sample) {
  __shared__ float red[MK_WARPS];
  __shared__ int red_i[MK_WARPS];
  for (int i = threadIdx.x; i < B && i < MAX_GROUP; i += MK_THREADS) s_rows[i] = i;
  __syncthreads();
  MkSync sy = {};
  sy.cta = blockIdx.x; sy.nc = gridDim.x;
  sample_phase(sp, B, pos, do_sample != 0, red, red_i, sy);
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
constexpr size_t PART_FLOATS = (size_t)4 << 20;  // 16 MB of fp32 split-K partials, shared out between the sequence groups
// Sequence groups.  The activation buffers hold MAX_GROUP rows cut into LAYOUT_GROUPS fixed regions; group g of a launch works in
// region g (compact rows of the group from the region's first row), with its own split-K partials, piece states, tickets, grid
// barrier counter, position counter, row list and activation tensor maps.  K/V caches, tokens and per-row outputs are indexed
// by the ORIGINAL row and shared.
constexpr int LAYOUT_GROUPS = MK_CTAS_PER_SM;
constexpr int GROUP_ROWS = MAX_GROUP / LAYOUT_GROUPS;  // rows of a region = 16 GV_MAX_MT
static_assert(GROUP_ROWS == 16 * GV_MAX_MT && LAYOUT_GROUPS <= WXB_MAX_DEC_GROUPS, "sequence-group layout");
constexpr size_t GROUP_PART_FLOATS = PART_FLOATS / LAYOUT_GROUPS;
constexpr int GROUP_PIECES = 1024, GROUP_TICKETS = 512;
#ifndef WXB_DEC_GROUP_MIN
#define WXB_DEC_GROUP_MIN 8
#endif
constexpr int DEC_GROUP_MIN_ROWS = WXB_DEC_GROUP_MIN;  // fewest rows per group for which a second group is started
#ifndef WXB_DEC_GROUP_DELAY_NS
#define WXB_DEC_GROUP_DELAY_NS 55000
#endif

struct DecBuffers {
  float *x, *logits, *part, *apart, *sum_lp;
  __nv_bfloat16 *att, *xn, *hid, *self_kv, *cross_kv;
  int *ticket, *d_pos, *tokens, *done, *rows, *ts_last;
  float* qlog;               // alignment-head query log (nullptr = off), see MkParams::qlog
  const signed char* qhead;
  int qA;
  unsigned* bar;
  const DecLayerW* layers;
  const CUtensorMap* maps;  // LAYOUT_GROUPS tables of n_maps entries
  size_t n_maps;
  int B, tok_stride;
  long long ldl;  // row stride of `logits` (n_vocab rounded up to 4 floats: 16-byte aligned rows)
};

// profiling aid of -DWXB_PROBE builds only (results become meaningless): WXB_DEC_SKIP bitmask, see MkParams::skip
int dec_skip_mask() {
#ifdef WXB_PROBE
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXB_DEC_SKIP");
    v = e ? atoi(e) : 0;
  }
  return v;
#else
  return 0;
#endif
}

// WXB_DEC_PROF=1: CTA 0 of the step kernel records the global timer at every grid barrier; after a decode the
// per-phase averages of the last launch are printed to stderr (tracing aid, no effect on results).
bool dec_prof_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WXB_DEC_PROF");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
constexpr size_t PROF_SLOTS = 1 << 16;  // the kernel keeps its start record in the last two slots
static_assert(PROF_SLOTS == (1 << 16), "dec_step_kernel writes prof[(1 << 16) - 1 / - 2]");
struct ProfLast { int mode = 0, n_steps = 0, L = 0, groups = 1; bool nsp = false; unsigned long long* dev = nullptr; } g_prof_last;

// Pick the split-K factor for y[B, N] = act[B, K] W[N, K]^T on G CTAs (tiles of 128 weight rows x K / gk).
// Cost model in microseconds: a CTA pulls its weight tile at ~60 KB/us and its activation slice from L2 at
// ~80 KB/us, every wave of tiles pays ~1 us of pipeline latency, and split-K partials are written once and
// read once through L2 (~10 MB/us chip-wide).
MkGemv plan_gemv(int N, int K, int B, int Bp, int G, bool full_k) {
  // (B = live rows of ONE sequence group; its partials live in GROUP_PART_FLOATS floats)
  MkGemv best = {N, K, 0, 0};
  double best_cost = 1e30;
  for (int gk = 1; gk <= K / GV_BK && gk <= GK_MAX; ++gk) {
    if (full_k && gk > 1) break;
    if (K % gk) continue;
    const int Ks = K / gk;
    if (Ks % GV_BK) continue;
    if ((size_t)gk * B * N > GROUP_PART_FLOATS) continue;
    const int tiles = ceil_div(N, GV_ROWS) * gk;
    const int waves = ceil_div(tiles, G);
    const double cost = waves * ((double)GV_ROWS * Ks * 2 / 60e3 + (double)Bp * Ks * 2 / 80e3 + 1.0) +
                        (gk > 1 ? 2.0 * gk * Bp * (double)N * 4 / 10e6 + 0.05 * gk : 0.0);
    if (cost < best_cost) { best_cost = cost; best.gk = gk; best.tiles = tiles; }
  }
  return best;
}

int alloc_buffers(wxb_ctx* ctx, int B, int tok_stride, DecBuffers* o) {
  auto nm = [&](const char* base) { return std::string(base); };
  const wxb_dims& D = ctx->model->dims;
  const int d = D.n_text_state, L = D.n_text_layer, H = D.n_text_head, V = D.n_vocab;
  o->B = B;
  o->tok_stride = tok_stride;
  o->ldl = ((long long)V + 3) & ~3LL;
  if (B > MAX_GROUP) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder: batch %d > %d sequences per call", B, MAX_GROUP);
  if (L > MAX_LAYERS) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder: %d layers > %d", L, MAX_LAYERS);
  if (d > 4 * LN_V4 * MK_THREADS || d % 64)
    return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder: n_text_state=%d unsupported (multiple of 64, <= %d)", d, 4 * LN_V4 * MK_THREADS);
  o->x = (float*)wxb_named(ctx, nm("dec.x").c_str(), (size_t)MAX_GROUP * d * 4);
  o->att = (__nv_bfloat16*)wxb_named(ctx, nm("dec.att").c_str(), (size_t)MAX_GROUP * d * 2);
  o->xn = (__nv_bfloat16*)wxb_named(ctx, nm("dec.xn").c_str(), (size_t)MAX_GROUP * d * 2);
  o->part = (float*)wxb_named(ctx, nm("dec.part").c_str(), PART_FLOATS * 4);
  o->hid = (__nv_bfloat16*)wxb_named(ctx, nm("dec.hid").c_str(), (size_t)MAX_GROUP * 4 * d * 2);
  o->logits = (float*)wxb_named(ctx, nm("dec.logits").c_str(), (size_t)MAX_GROUP * o->ldl * 4);
  o->apart = (float*)wxb_named(ctx, nm("dec.apart").c_str(), (size_t)LAYOUT_GROUPS * GROUP_PIECES * 66 * 4);
  o->sum_lp = (float*)wxb_named(ctx, nm("dec.sum_lp").c_str(), (size_t)B * 4);
  o->self_kv = (__nv_bfloat16*)wxb_named(ctx, nm("dec.self_kv").c_str(), (size_t)L * 2 * B * H * D.n_text_ctx * 64 * 2);
  o->cross_kv = (__nv_bfloat16*)wxb_named(ctx, nm("dec.cross_kv").c_str(), (size_t)L * 2 * B * H * T_AUDIO * 64 * 2);
  o->ticket = (int*)wxb_named(ctx, nm("dec.ticket").c_str(), (size_t)LAYOUT_GROUPS * GROUP_TICKETS * 4, true);
  o->d_pos = (int*)wxb_named(ctx, nm("dec.pos").c_str(), 64 * WXB_MAX_DEC_GROUPS);            // one 64-byte slot per group
  o->bar = (unsigned*)wxb_named(ctx, "dec.bar", 128 * WXB_MAX_DEC_GROUPS, true);  // grid-barrier counters (128 bytes apart), zeroed before every launch
  o->tokens = (int*)wxb_named(ctx, nm("dec.tokens").c_str(), (size_t)B * tok_stride * 4);
  o->done = (int*)wxb_named(ctx, nm("dec.done").c_str(), (size_t)B * 4);
  o->rows = (int*)wxb_named(ctx, "dec.rows", (size_t)2 * MAX_GROUP * 4);  // two row lists: a launch may still read the other one
  o->ts_last = (int*)wxb_named(ctx, "dec.ts_last", (size_t)MAX_GROUP * 4);
  DecLayerW* layers = (DecLayerW*)wxb_named(ctx, "dec.layers", (size_t)L * sizeof(DecLayerW));
  if (!o->x || !o->att || !o->hid || !o->logits || !o->part || !o->apart || !o->sum_lp || !o->self_kv || !o->cross_kv ||
      !o->ticket || !o->d_pos || !o->bar || !o->tokens || !o->done || !o->xn || !layers || !o->rows || !o->ts_last)
    return WXB_ERR_CUDA;
  if (ctx->dec_layers_model != (const void*)ctx->model) {
    std::vector<DecLayerW> h(L);
    for (int l = 0; l < L; ++l) {
      int rc = wxb_dec_layer(ctx, l, &h[l]);
      if (rc != WXB_OK) return rc;
    }
    WXB_CUDA(ctx, cudaMemcpy(layers, h.data(), (size_t)L * sizeof(DecLayerW), cudaMemcpyHostToDevice));
    ctx->dec_layers_model = ctx->model;
  }
  o->layers = layers;
  // alignment-head query log for the DTW word timing (wxb_decode_collect_heads)
  o->qlog = nullptr; o->qhead = nullptr; o->qA = 0;
  ctx->qlog_valid = false;
  if (!ctx->align_heads.empty()) {
    const int A = (int)ctx->align_heads.size() / 2;
    signed char* qh = (signed char*)wxb_named(ctx, "dec.qhead", (size_t)MAX_LAYERS * 64);
    int* qhs = (int*)wxb_named(ctx, "dec.qheads", (size_t)128 * 2 * 4);
    o->qlog = (float*)wxb_named(ctx, "dec.qlog", (size_t)B * D.n_text_ctx * A * 64 * 4);
    if (!qh || !qhs || !o->qlog) return WXB_ERR_CUDA;
    if (H > 64) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder: alignment-head logging takes at most 64 heads per layer");
    if (ctx->align_heads_dirty) {
      std::vector<signed char> h((size_t)MAX_LAYERS * 64, (signed char)-1);
      for (int a = 0; a < A; ++a) {
        const int l = ctx->align_heads[2 * a], hd = ctx->align_heads[2 * a + 1];
        if (l < 0 || l >= L || hd < 0 || hd >= H)
          return wxb_fail(ctx, WXB_ERR_INVALID, "alignment head (%d, %d) outside the decoder's %d layers x %d heads", l, hd, L, H);
        h[(size_t)l * H + hd] = (signed char)a;
      }
      WXB_CUDA(ctx, cudaDeviceSynchronize());  // a previous decode may still be reading the old tables
      WXB_CUDA(ctx, cudaMemcpy(qh, h.data(), h.size(), cudaMemcpyHostToDevice));
      WXB_CUDA(ctx, cudaMemcpy(qhs, ctx->align_heads.data(), ctx->align_heads.size() * 4, cudaMemcpyHostToDevice));
      ctx->align_heads_dirty = false;
    }
    o->qhead = qh; o->qA = A;
    ctx->qlog_B0 = B; ctx->qlog_pos = 0;
  }
  // tensor maps (128-byte swizzle, 64-element boxes): weights [N, K] in 128-row boxes, activations [B, K] in boxes of
  // 16 MT rows (one map per batch-tile count MT = 1 .. 4: a launch over fewer live rows uses the narrower box); rows >= B
  // are zero-filled by the TMA unit
  const size_t n_maps = (size_t)TM_PER_LAYER * L + TM_TAIL_COUNT;
  CUtensorMap* maps = (CUtensorMap*)wxb_named(ctx, nm("dec.maps").c_str(), LAYOUT_GROUPS * n_maps * sizeof(CUtensorMap));
  if (!maps) return WXB_ERR_CUDA;
  o->n_maps = n_maps;
  wxb_dec_maps_key& key = ctx->dec_maps_key;
  if (key.model != (const void*)ctx->model || key.xn != o->xn || key.att != o->att || key.hid != o->hid || key.ckv != o->cross_kv || key.B != B ||
      key.groups != LAYOUT_GROUPS) {
    std::vector<CUtensorMap> h(LAYOUT_GROUPS * n_maps);
    int rc;
    for (int l = 0; l < L; ++l) {
      DecLayerW w;
      if ((rc = wxb_dec_layer(ctx, l, &w)) != WXB_OK) return rc;
      const struct { const __nv_bfloat16* p; int N, K; } ws[TM_PER_LAYER] = {
          {w.qkv_w, 3 * d, d}, {w.out_w, d, d}, {w.cq_w, d, d}, {w.cout_w, d, d}, {w.fc1_w, 4 * d, d}, {w.fc2_w, d, 4 * d}};
      for (int i = 0; i < TM_PER_LAYER; ++i)
        if ((rc = wxb_make_tmap_bf16(ctx, &h[(size_t)l * TM_PER_LAYER + i], ws[i].p, (uint64_t)ws[i].K, (uint64_t)ws[i].N,
                                     (uint64_t)ws[i].K * 2, GV_BK, GV_ROWS)) != WXB_OK)
          return rc;
    }
    const void* emb = wxb_weight(ctx, "dec.emb");
    if (!emb) return WXB_ERR_STATE;
    CUtensorMap* am = &h[(size_t)TM_PER_LAYER * L];
    if ((rc = wxb_make_tmap_bf16(ctx, am + TM_EMB, emb, (uint64_t)d, (uint64_t)V, (uint64_t)d * 2, GV_BK, GV_ROWS)) != WXB_OK) return rc;
    for (int mt = 1; mt <= GV_MAX_MT; ++mt) {  // region 0; the other regions' activation maps are made below
      CUtensorMap* a3 = am + TM_ACT + (mt - 1) * 3;
      if ((rc = wxb_make_tmap_bf16(ctx, a3 + 0, o->xn, (uint64_t)d, (uint64_t)GROUP_ROWS, (uint64_t)d * 2, GV_BK, 16 * mt)) != WXB_OK) return rc;
      if ((rc = wxb_make_tmap_bf16(ctx, a3 + 1, o->att, (uint64_t)d, (uint64_t)GROUP_ROWS, (uint64_t)d * 2, GV_BK, 16 * mt)) != WXB_OK) return rc;
      if ((rc = wxb_make_tmap_bf16(ctx, a3 + 2, o->hid, (uint64_t)4 * d, (uint64_t)GROUP_ROWS, (uint64_t)4 * d * 2, GV_BK, 16 * mt)) != WXB_OK) return rc;
    }
    // cross K/V of all layers as one [rows, 64] tensor read in full-stage boxes (shorter boxes at the end of an item)
    if ((rc = wxb_make_tmap_bf16(ctx, am + TM_KV, o->cross_kv, 64, (uint64_t)L * 2 * B * H * T_AUDIO, 128, 64, XA_KEYS)) != WXB_OK) return rc;
    if ((rc = wxb_make_tmap_bf16(ctx, am + TM_KV + 1, o->cross_kv, 64, (uint64_t)L * 2 * B * H * T_AUDIO, 128, 64, XA_TAIL)) != WXB_OK) return rc;
    // the tables of regions 1 ..: the same weight / K/V maps, activation maps over the region's rows (rows past the region are
    // zero-filled by the TMA unit; rows past a launch's live rows hold stale values whose products are never stored)
    for (int g = 1; g < LAYOUT_GROUPS; ++g) {
      for (size_t i = 0; i < n_maps; ++i) h[g * n_maps + i] = h[i];
      CUtensorMap* amg = &h[g * n_maps + (size_t)TM_PER_LAYER * L];
      const size_t r0 = (size_t)g * GROUP_ROWS;
      for (int mt = 1; mt <= GV_MAX_MT; ++mt) {
        CUtensorMap* a3 = amg + TM_ACT + (mt - 1) * 3;
        if ((rc = wxb_make_tmap_bf16(ctx, a3 + 0, o->xn + r0 * d, (uint64_t)d, (uint64_t)GROUP_ROWS, (uint64_t)d * 2, GV_BK, 16 * mt)) != WXB_OK) return rc;
        if ((rc = wxb_make_tmap_bf16(ctx, a3 + 1, o->att + r0 * d, (uint64_t)d, (uint64_t)GROUP_ROWS, (uint64_t)d * 2, GV_BK, 16 * mt)) != WXB_OK) return rc;
        if ((rc = wxb_make_tmap_bf16(ctx, a3 + 2, o->hid + r0 * 4 * d, (uint64_t)4 * d, (uint64_t)GROUP_ROWS, (uint64_t)4 * d * 2, GV_BK, 16 * mt)) != WXB_OK) return rc;
      }
    }
    WXB_CUDA(ctx, cudaDeviceSynchronize());  // a previous decode may still be reading the old table
    WXB_CUDA(ctx, cudaMemcpy(maps, h.data(), LAYOUT_GROUPS * n_maps * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    key.model = ctx->model; key.xn = o->xn; key.att = o->att; key.hid = o->hid; key.ckv = o->cross_kv; key.B = B; key.groups = LAYOUT_GROUPS;
  }
  o->maps = maps;
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

// Sequence groups of a launch over n_live rows: group g takes rows [lo[g], lo[g + 1]) of the live list.
struct GroupSplit {
  int G;
  int lo[WXB_MAX_DEC_GROUPS + 1];
};
int split_groups(wxb_ctx* ctx, int n_live, GroupSplit* gs) {
  int G = (LAYOUT_GROUPS >= 2 && n_live >= 2 * DEC_GROUP_MIN_ROWS) ? 2 : 1;
  if (ctx->dec_groups_override > 0) G = ctx->dec_groups_override;
  G = std::max(1, std::min(std::min(G, LAYOUT_GROUPS), n_live));
  while (G < LAYOUT_GROUPS && ceil_div(n_live, G) > GROUP_ROWS) ++G;
  if (ceil_div(n_live, G) > GROUP_ROWS)
    return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder: %d rows do not fit %d sequence group(s) of %d rows", n_live, G, GROUP_ROWS);
  gs->G = G;
  for (int g = 0; g <= G; ++g) gs->lo[g] = (int)((long long)n_live * g / G);
  return WXB_OK;
}

// The row lists of a launch: group g's live rows (original row numbers) at rows_base + g * GROUP_ROWS.
int upload_rows(wxb_ctx* ctx, const GroupSplit& gs, const int* live_host, int* rows_base, cudaStream_t st) {
  for (int g = 0; g < gs.G; ++g)
    WXB_CUDA(ctx, cudaMemcpyAsync(rows_base + g * GROUP_ROWS, live_host + gs.lo[g], (size_t)(gs.lo[g + 1] - gs.lo[g]) * 4,
                                  cudaMemcpyHostToDevice, st));
  return WXB_OK;
}

// Launch the persistent step kernel for ONE sequence group: n_steps consecutive positions starting at *d_pos, over the B live
// rows listed in rows_dev (original row numbers), in activation region `region`.
int launch_group(wxb_ctx* ctx, const DecBuffers& buf, int region, int mode, int n_steps, const SampleParams& sp, float* logits_out,
                 long long ldl, int B, const int* rows_dev, unsigned delay_ns, cudaStream_t st) {
  const wxb_dims& D = ctx->model->dims;
  const int d = D.n_text_state;
  const int Bp = (B + 15) & ~15, MT = Bp / 16;
  int G = ctx->sm_count;
#ifdef WXB_PROBE
  if (const char* e = getenv("WXB_DEC_GRID")) { const int g = atoi(e); if (g > 0 && g < G) G = g; }  // timing probe: narrower grid
#endif
  const size_t r0 = (size_t)region * GROUP_ROWS;
  MkParams p = {};
  p.B = B; p.B0 = buf.B; p.rows = rows_dev;
  p.d = d; p.H = D.n_text_head; p.L = D.n_text_layer; p.V = D.n_vocab; p.TX = D.n_text_ctx;
  p.mode = mode; p.n_steps = n_steps; p.skip = dec_skip_mask(); p.delay_ns = delay_ns;
  p.layers = buf.layers; p.maps = buf.maps + (size_t)region * buf.n_maps;
  p.emb = (const __nv_bfloat16*)wxb_weight(ctx, "dec.emb");
  p.pos_emb = (const float*)wxb_weight(ctx, "dec.pos");
  p.lnf_w = (const float*)wxb_weight(ctx, "dec.ln.w");
  p.lnf_b = (const float*)wxb_weight(ctx, "dec.ln.b");
  if (!p.emb || !p.pos_emb || !p.lnf_w || !p.lnf_b) return WXB_ERR_STATE;
  p.tokens = buf.tokens; p.tok_stride = buf.tok_stride; p.d_pos = buf.d_pos + 16 * region;
  p.x = buf.x + r0 * d; p.xn = buf.xn + r0 * d; p.att = buf.att + r0 * d; p.hid = buf.hid + r0 * 4 * d;
  p.part = buf.part + (size_t)region * GROUP_PART_FLOATS;
  p.self_kv = buf.self_kv; p.cross_kv = buf.cross_kv;
  p.logits = logits_out; p.ldl = ldl;
  p.apart = buf.apart + (size_t)region * GROUP_PIECES * 66; p.ticket = buf.ticket + region * GROUP_TICKETS; p.bar = buf.bar + 32 * region;
  p.qlog = buf.qlog; p.qhead = buf.qhead; p.qA = buf.qA;
  if (dec_prof_enabled()) {
    unsigned long long* base = (unsigned long long*)wxb_named(ctx, "dec.prof", LAYOUT_GROUPS * PROF_SLOTS * 8);
    p.prof = base ? base + (size_t)region * PROF_SLOTS : nullptr;
    if ((size_t)n_steps * (11 * p.L + 4) + 2 > PROF_SLOTS) p.prof = nullptr;
    if (region == 0) {
      g_prof_last.mode = mode; g_prof_last.n_steps = n_steps; g_prof_last.L = p.L; g_prof_last.nsp = sp.nsp_out != nullptr;
      g_prof_last.dev = p.prof; g_prof_last.groups = 1;
    } else {
      g_prof_last.groups = region + 1;
    }
  }
  p.g_qkv = plan_gemv(3 * d, d, B, Bp, G, false);
  p.g_dd = plan_gemv(d, d, B, Bp, G, false);
  p.g_fc1 = plan_gemv(4 * d, d, B, Bp, G, true);
  p.g_fc2 = plan_gemv(d, 4 * d, B, Bp, G, false);
  p.g_logits = plan_gemv(D.n_vocab, d, B, Bp, G, true);
  if (!p.g_qkv.gk || !p.g_dd.gk || !p.g_fc1.gk || !p.g_fc2.gk || !p.g_logits.gk)
    return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder: no GEMV tiling fits (d=%d, batch %d)", d, B);
  p.sp = sp;
  if (p.sp.logits) p.sp.logits += r0 * (size_t)p.sp.ldl;  // the sampling phase reads the group's rows of the scratch logits
  p.sp.part_stats = p.apart;    // idle outside the cross-attention phases (<= 148 slices x 8 floats of the 1024 x 66)
  p.sp.part_ticket = p.ticket;  // all zero outside the cross-attention phases
  p.scale = 1.0f / sqrtf(64.f);
  if (MT > GV_MAX_MT) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder: %d rows in one sequence group (at most %d)", B, 16 * GV_MAX_MT);
  void (*kern)(const MkParams) = MT == 1 ? dec_step_kernel<1> : dec_step_kernel<2>;
  if (GV_MAX_MT > 2 && MT > 2) kern = MT == 3 ? dec_step_kernel<GV_MAX_MT < 3 ? 1 : 3> : dec_step_kernel<GV_MAX_MT < 4 ? 1 : 4>;
  if (ctx->func_smem.find(reinterpret_cast<const void*>(kern)) == ctx->func_smem.end()) {
    int rc = wxb_func_smem(ctx, kern, (int)MK_SMEM);
    if (rc != WXB_OK) return rc;
    if (MK_CTAS_PER_SM > 1)  // two instances share an SM: leave the whole configurable array to shared memory
      WXB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    WXB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MK_THREADS, MK_SMEM));
    if (per_sm < 1) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "decoder: persistent step kernel does not fit an SM");
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G);
  cfg.blockDim = dim3(MK_THREADS);
  cfg.dynamicSmemBytes = MK_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
#ifdef WXB_DEC_NO_COOP
  cfg.numAttrs = 0;
#else
  cfg.numAttrs = 1;
#endif
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  ctx->launches++;
  if (e != cudaSuccess) return wxb_fail(ctx, WXB_ERR_CUDA, "decoder step launch failed: %s", cudaGetErrorString(e));
  return WXB_OK;
}

// One launch round: every sequence group of the split runs n_steps positions, group 0 on the caller's stream, the others on side
// streams forked from it and joined back, so that the caller's stream order covers all of them.  rows_base: the row lists written
// by upload_rows.  by_row: logits_out is indexed by the ORIGINAL row (teacher-forced logits with all rows live); otherwise it is the
// scratch indexed by region row.
int launch_groups(wxb_ctx* ctx, const DecBuffers& buf, const GroupSplit& gs, int mode, int n_steps, const SampleParams& sp,
                  float* logits_out, long long ldl, bool by_row, const int* rows_base, cudaStream_t st) {
  WXB_CUDA(ctx, cudaMemsetAsync(buf.bar, 0, 128 * WXB_MAX_DEC_GROUPS, st));
  if (buf.qlog) { ctx->qlog_pos += n_steps; ctx->qlog_valid = true; }
  if (gs.G > 1) {
    if (!ctx->dec_fork) WXB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->dec_fork, cudaEventDisableTiming));
    WXB_CUDA(ctx, cudaEventRecord(ctx->dec_fork, st));
  }
  int rc;
  for (int g = 0; g < gs.G; ++g) {
    cudaStream_t sg = st;
    if (g > 0) {
      if (!ctx->dec_side[g - 1]) {
        WXB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->dec_side[g - 1], cudaStreamNonBlocking));
        WXB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->dec_join[g - 1], cudaEventDisableTiming));
      }
      sg = ctx->dec_side[g - 1];
      WXB_CUDA(ctx, cudaStreamWaitEvent(sg, ctx->dec_fork, 0));
    }
    const size_t lrow = by_row ? (size_t)gs.lo[g] : (size_t)g * GROUP_ROWS;
    const unsigned delay = (unsigned)((long long)ctx->dec_group_delay_ns * g / gs.G);
    if ((rc = launch_group(ctx, buf, g, mode, n_steps, sp, logits_out + lrow * (size_t)ldl, ldl, gs.lo[g + 1] - gs.lo[g],
                           rows_base + g * GROUP_ROWS, delay, sg)) != WXB_OK)
      return rc;
    if (g > 0) {
      WXB_CUDA(ctx, cudaEventRecord(ctx->dec_join[g - 1], sg));
      WXB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->dec_join[g - 1], 0));
    }
  }
  return WXB_OK;
}

int dump_prof(wxb_ctx* ctx) {
  const ProfLast& P = g_prof_last;
  if (!P.dev) return WXB_OK;
  WXB_CUDA(ctx, cudaDeviceSynchronize());
  const int tail = (P.mode >= 1 ? 2 : 0) + ((P.mode == 2 || P.nsp) ? 1 : 0);
  const int per_step = 11 * P.L + tail;  // phases = barriers per step (the last phase of the launch has none)
  const int n = per_step * P.n_steps - 1;
  std::vector<unsigned long long> t(n);
  WXB_CUDA(ctx, cudaMemcpy(t.data(), P.dev, (size_t)n * 8, cudaMemcpyDeviceToHost));
  static const char* names[11] = {"ln1", "qkv", "self", "out", "ln2", "cq", "cross", "cout", "ln3", "fc1", "fc2"};
  double sum[16] = {0};
  long cnt[16] = {0};
  for (int s = 0; s < P.n_steps; ++s)
    for (int i = 0; i < per_step; ++i) {
      const int k = s * per_step + i;
      if (k == 0 || k >= n) continue;
      const double us = (double)(t[k] - t[k - 1]) * 1e-3;
      const int slot = i < 11 * P.L ? i % 11 : 11 + (i - 11 * P.L);  // 11 lnf, 12 logits, 13 sample (barrier after it = step end)
      sum[slot] += us; cnt[slot]++;
    }
  fprintf(stderr, "[wxb dec prof] mode %d, %d steps/launch, per-phase mean us (phase + its barrier):", P.mode, P.n_steps);
  double layer = 0;
  for (int i = 0; i < 11; ++i) { if (cnt[i]) { fprintf(stderr, " %s %.2f", names[i], sum[i] / cnt[i]); layer += sum[i] / cnt[i]; } }
  fprintf(stderr, " | layer %.2f |", layer);
  static const char* tn[4] = {"lnf", "logits", "sample", "x"};
  for (int i = 11; i < 15; ++i) if (cnt[i]) fprintf(stderr, " %s %.2f", tn[i - 11], sum[i] / cnt[i]);
  fprintf(stderr, " | step %.1f us\n", (double)(t[n - 1] - t[0]) * 1e-3 / P.n_steps);
  for (int g = 1; g < P.groups; ++g) {
    // the other sequence groups' instances: how far behind group 0 they pass the same barriers, and where their CTA 0 ran
    std::vector<unsigned long long> u(n);
    unsigned long long meta[2][2];
    WXB_CUDA(ctx, cudaMemcpy(u.data(), P.dev + (size_t)g * PROF_SLOTS, (size_t)n * 8, cudaMemcpyDeviceToHost));
    WXB_CUDA(ctx, cudaMemcpy(meta[0], P.dev + PROF_SLOTS - 2, 16, cudaMemcpyDeviceToHost));
    WXB_CUDA(ctx, cudaMemcpy(meta[1], P.dev + (size_t)g * PROF_SLOTS + PROF_SLOTS - 2, 16, cudaMemcpyDeviceToHost));
    double lag_cross_in = 0, lag_cross_out = 0;
    long c = 0;
    for (int s = 0; s < P.n_steps; ++s)
      for (int l = 0; l < P.L; ++l) {
        const int k = s * per_step + l * 11 + 5;  // barrier after cq = start of the cross-attention phase
        if (k + 1 >= n) continue;
        lag_cross_in += (double)((long long)(u[k] - t[k])) * 1e-3;
        lag_cross_out += (double)((long long)(u[k + 1] - t[k + 1])) * 1e-3;
        ++c;
      }
    fprintf(stderr, "[wxb dec prof] group %d: CTA 0 on SM %llu (group 0: SM %llu), started %.1f us after group 0, enters / leaves cross-attention %.1f / %.1f us after group 0, ends %.1f us after\n",
            g, meta[1][1], meta[0][1], (double)((long long)(meta[1][0] - meta[0][0])) * 1e-3, c ? lag_cross_in / c : 0.0,
            c ? lag_cross_out / c : 0.0, (double)((long long)(u[n - 1] - t[n - 1])) * 1e-3);
  }
  return WXB_OK;
}

}  // namespace

void wxb_decoder_reset(wxb_ctx* ctx) {
  ctx->dec_layers_model = nullptr;
  ctx->dec_maps_key = wxb_dec_maps_key();
  ctx->align_heads_dirty = !ctx->align_heads.empty();  // re-validated against the new model's layers x heads
  ctx->qlog_valid = false;
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
  if (opts->no_speech >= D.n_vocab) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decode_greedy: no_speech out of range");
  if (opts->apply_timestamp_rules && (opts->timestamp_begin <= opts->eot || opts->timestamp_begin >= D.n_vocab))
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decode_greedy: timestamp_begin %d out of range (eot %d, n_vocab %d)",
                    opts->timestamp_begin, opts->eot, D.n_vocab);
  cudaStream_t st = (cudaStream_t)stream;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int stride = D.n_text_ctx + 1;
  int rc;
  DecBuffers buf;
  if ((rc = alloc_buffers(ctx, B, stride, &buf)) != WXB_OK) return rc;
  // device-side timing is opt-in (wxb_decode_stats(reset = 1) switches it on); the events belong to the ctx from the moment
  // they exist, so an error return below cannot leak them
  wxb_dec_timing* tm = nullptr;
  if (ctx->dec_timing_on) {
    if (ctx->dec_timings.size() >= WXB_MAX_DEC_TIMINGS) wxb_dec_timings_clear(ctx);
    wxb_dec_timing t = {};
    WXB_CUDA(ctx, cudaEventCreate(&t.e0));
    if (cudaEventCreate(&t.e1) != cudaSuccess) { cudaEventDestroy(t.e0); return wxb_fail(ctx, WXB_ERR_CUDA, "cudaEventCreate failed"); }
    if (cudaEventCreate(&t.e2) != cudaSuccess) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); return wxb_fail(ctx, WXB_ERR_CUDA, "cudaEventCreate failed"); }
    ctx->dec_timings.push_back(t);
    tm = &ctx->dec_timings.back();
    // e1 / e2 are recorded below; an early error return leaves them unrecorded, wxb_decode_stats skips such entries
    WXB_CUDA(ctx, cudaEventRecord(tm->e0, st));
  }
  // tokens[b, :prompt_len] = prompt; state reset
  std::vector<int> init((size_t)B * stride, opts->eot);
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < prompt_len; ++i) init[(size_t)b * stride + i] = prompt_host[i];
  WXB_CUDA(ctx, cudaMemcpyAsync(buf.tokens, init.data(), (size_t)B * stride * 4, cudaMemcpyHostToDevice, st));
  int pos_host[16 * WXB_MAX_DEC_GROUPS] = {};  // one 64-byte slot per region
  WXB_CUDA(ctx, cudaMemsetAsync(buf.d_pos, 0, 64 * WXB_MAX_DEC_GROUPS, st));
  WXB_CUDA(ctx, cudaMemsetAsync(buf.done, 0, (size_t)B * 4, st));
  WXB_CUDA(ctx, cudaMemsetAsync(buf.sum_lp, 0, (size_t)B * 4, st));
  WXB_CUDA(ctx, cudaMemsetAsync(buf.ts_last, 0xff, (size_t)MAX_GROUP * 4, st));  // -1: no timestamp sampled yet
  if ((rc = cross_kv_precompute(ctx, (const __nv_bfloat16*)enc_out_dev, buf, st)) != WXB_OK) return rc;
  SampleParams sp = {};
  sp.logits = buf.logits; sp.ldl = buf.ldl; sp.V = D.n_vocab; sp.tokens = buf.tokens; sp.stride = stride;
  sp.prompt_len = prompt_len; sp.eot = opts->eot; sp.suppress_blank = opts->suppress_blank; sp.blank_token = opts->blank_token;
  sp.n_suppress = opts->n_suppress; sp.suppress = opts->suppress_dev; sp.sum_logprob = buf.sum_lp; sp.done = buf.done;
  sp.nsp_out = nullptr; sp.nsp_token = opts->no_speech;
  sp.ts_rules = opts->apply_timestamp_rules ? 1 : 0; sp.ts_begin = opts->timestamp_begin; sp.no_timestamps = opts->no_timestamps;
  sp.max_initial_ts = opts->max_initial_timestamp_index; sp.ts_last = buf.ts_last;
  WXB_CUDA(ctx, cudaStreamSynchronize(st));  // `init` is pageable host memory
  if (tm) WXB_CUDA(ctx, cudaEventRecord(tm->e1, st));

  const bool want_nsp = (opts->no_speech >= 0 && no_speech_prob_dev);
  std::vector<int> done_host(B), live(B);
  for (int b = 0; b < B; ++b) live[b] = b;
  int n_live = B, n_sampled = 0, flip = 0;
  GroupSplit gs;
  if ((rc = split_groups(ctx, n_live, &gs)) != WXB_OK) return rc;
  int* rows_dev = buf.rows;
  if ((rc = upload_rows(ctx, gs, live.data(), rows_dev, st)) != WXB_OK) return rc;
  // prompt positions 0 .. prompt_len-2 (forced tokens); logits only at position 0 for no_speech_prob
  for (int pos = 0; pos < prompt_len - 1; ++pos) {
    SampleParams s1 = sp;
    const bool nsp = (pos == 0 && want_nsp);
    if (nsp) s1.nsp_out = no_speech_prob_dev;
    if ((rc = launch_groups(ctx, buf, gs, nsp ? 1 : 0, 1, s1, buf.logits, buf.ldl, false, rows_dev, st)) != WXB_OK) return rc;
  }
  // Sampling loop.  Every `check_every` positions the host reads the EOT flags (mlx_whisper_batch_decoder.py:357) and the
  // next launch runs over the rows that are still live (:37-100: finished sequences leave the batch; their K/V slabs are
  // no longer streamed and their GEMV / LayerNorm / attention work disappears); the live rows are dealt to the sequence
  // groups anew.
  const int check_every = opts->check_every > 0 ? opts->check_every : 16;
  while (n_sampled < sample_len) {
    // a single-token prompt makes the SOT position the first sampling position: that step also emits no_speech_prob
    const bool nsp_now = (n_sampled == 0 && prompt_len == 1 && want_nsp);
    const int n = nsp_now ? 1 : std::min(check_every, sample_len - n_sampled);
    SampleParams s1 = sp;
    if (nsp_now) s1.nsp_out = no_speech_prob_dev;
    if ((rc = launch_groups(ctx, buf, gs, 2, n, s1, buf.logits, buf.ldl, false, rows_dev, st)) != WXB_OK) return rc;
    n_sampled += n;
    if (n_sampled < sample_len && !nsp_now) {
      WXB_CUDA(ctx, cudaMemcpyAsync(done_host.data(), buf.done, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
      WXB_CUDA(ctx, cudaStreamSynchronize(st));
      int m = 0;
      for (int i = 0; i < n_live; ++i)
        if (!done_host[live[i]]) live[m++] = live[i];
      if (m == 0) break;  // every row has emitted EOT
      if (m < n_live && !opts->no_compaction) {
        n_live = m;
        // the position counter of every region a later split may use: all groups have advanced to the same position
        const int pos_now = prompt_len - 1 + n_sampled;
        for (int g = 0; g < LAYOUT_GROUPS; ++g) pos_host[16 * g] = pos_now;
        WXB_CUDA(ctx, cudaMemcpyAsync(buf.d_pos, pos_host, sizeof(pos_host), cudaMemcpyHostToDevice, st));
        if ((rc = split_groups(ctx, n_live, &gs)) != WXB_OK) return rc;
        flip ^= 1;
        rows_dev = buf.rows + (flip ? MAX_GROUP : 0);
        if ((rc = upload_rows(ctx, gs, live.data(), rows_dev, st)) != WXB_OK) return rc;
      }
    }
  }
  dec_finalize_kernel<<<B, 256, 0, st>>>(buf.tokens, stride, prompt_len, n_sampled, sample_len, opts->eot, tokens_out_dev, n_tokens_dev);
  WXB_LAUNCH_CHECK(ctx);
  WXB_CUDA(ctx, cudaMemcpyAsync(sum_logprob_dev, buf.sum_lp, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  if (tm) {
    WXB_CUDA(ctx, cudaEventRecord(tm->e2, st));
    tm->steps = prompt_len - 1 + n_sampled;
  }
  if (dec_prof_enabled()) return dump_prof(ctx);
  return WXB_OK;
}

extern "C" int wxb_decode_stats(wxb_ctx* ctx, double* cross_kv_ms, double* steps_ms, int64_t* n_steps, int reset) {
  if (!ctx) return WXB_ERR_INVALID;
  double a = 0, b = 0;
  int64_t n = 0;
  for (auto& t : ctx->dec_timings) {
    if (t.steps <= 0) continue;  // a call that failed before its last event was recorded
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
    wxb_dec_timings_clear(ctx);
    ctx->dec_timing_on = true;  // timing is opt-in: the first reset switches it on
  }
  return WXB_OK;
}

extern "C" int wxb_decoder_sample(wxb_ctx* ctx, float* logits_dev, int64_t ldl, int B, int n_vocab, int32_t* tokens_dev, int stride,
                                  int pos, int prompt_len, const wxb_decode_opts* opts, float* sum_logprob_dev, int32_t* done_dev,
                                  int32_t* ts_last_dev, float* no_speech_prob_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!logits_dev || B <= 0 || B > MAX_GROUP || n_vocab <= 0 || ldl < n_vocab || (ldl & 3) || !tokens_dev || pos < 0 || pos + 1 >= stride ||
      prompt_len <= 0 || !opts || !sum_logprob_dev || !done_dev || (reinterpret_cast<uintptr_t>(logits_dev) & 15))
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decoder_sample: bad argument (ldl must be a multiple of 4 floats, logits 16-byte aligned)");
  if (opts->eot < 0 || opts->eot >= n_vocab) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decoder_sample: eot out of range");
  if (opts->apply_timestamp_rules && (!ts_last_dev || opts->timestamp_begin <= opts->eot || opts->timestamp_begin >= n_vocab))
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_decoder_sample: timestamp rules need ts_last_dev and eot < timestamp_begin < n_vocab");
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  SampleParams sp = {};
  sp.logits = logits_dev; sp.ldl = ldl; sp.V = n_vocab; sp.tokens = tokens_dev; sp.stride = stride; sp.prompt_len = prompt_len;
  sp.eot = opts->eot; sp.suppress_blank = opts->suppress_blank; sp.blank_token = opts->blank_token;
  sp.n_suppress = opts->n_suppress; sp.suppress = opts->suppress_dev; sp.sum_logprob = sum_logprob_dev; sp.done = done_dev;
  sp.nsp_out = (opts->no_speech >= 0) ? no_speech_prob_dev : nullptr; sp.nsp_token = opts->no_speech;
  sp.ts_rules = opts->apply_timestamp_rules ? 1 : 0; sp.ts_begin = opts->timestamp_begin; sp.no_timestamps = opts->no_timestamps;
  sp.max_initial_ts = opts->max_initial_timestamp_index; sp.ts_last = ts_last_dev;
  dec_sample_kernel<<<B, MK_THREADS, 0, (cudaStream_t)stream>>>(sp, B, pos, 1);
  WXB_LAUNCH_CHECK(ctx);
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
  if ((rc = alloc_buffers(ctx, B, n_tok, &buf)) != WXB_OK) return rc;
  WXB_CUDA(ctx, cudaMemcpyAsync(buf.tokens, tokens_host, (size_t)B * n_tok * 4, cudaMemcpyHostToDevice, st));
  WXB_CUDA(ctx, cudaMemsetAsync(buf.d_pos, 0, 64 * WXB_MAX_DEC_GROUPS, st));
  if ((rc = cross_kv_precompute(ctx, (const __nv_bfloat16*)enc_out_dev, buf, st)) != WXB_OK) return rc;
  std::vector<int> live(B);
  for (int b = 0; b < B; ++b) live[b] = b;
  GroupSplit gs;
  if ((rc = split_groups(ctx, B, &gs)) != WXB_OK) return rc;
  if ((rc = upload_rows(ctx, gs, live.data(), buf.rows, st)) != WXB_OK) return rc;
  WXB_CUDA(ctx, cudaStreamSynchronize(st));  // `live` is pageable host memory
  SampleParams sp = {};
  for (int pos = 0; pos < n_tok; ++pos)
    if ((rc = launch_groups(ctx, buf, gs, 1, 1, sp, logits_out_dev + (size_t)pos * D.n_vocab, (long long)n_tok * D.n_vocab, true, buf.rows, st)) != WXB_OK) return rc;
  return WXB_OK;
}
