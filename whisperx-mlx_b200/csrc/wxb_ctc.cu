// wxb_ctc.cu — K4: CTC forced-alignment trellis + backtrack / beam-2, one warp per segment.
//
// Numeric contract (bit-exact on indices and trellis values against the reference):
//   whisperx/alignment.py:387-404  get_trellis        (fp32 add then max, no FMA; column 0 is a
//                                                      cumsum accumulated in fp64 and rounded to
//                                                      fp32 per element = torch CPU cumsum)
//   whisperx/alignment.py:407-437  get_wildcard_emission (token -1 -> max over non-blank labels)
//   whisperx/alignment.py:447-481  backtrack
//   whisperx/alignment.py:500-579  backtrack_beam(beam_width=2) (what align() calls, :269)
//
// Layout: segments are ragged; emissions [sumT, V] and tokens [sumN] are packed back to back and
// a small per-segment descriptor table gives the offsets.  Each warp owns one segment: lane l
// owns trellis columns l, l+32, ...; the previous row lives in shared memory (ping-pong), the
// emission rows stream through an 8-deep cp.async ring so the t-recurrence never waits on HBM.
#include "wxb_common.cuh"
#include <math.h>

struct CtcSeg {
  int t_off, T, n_off, N;
  long long tr_off;
};

#define CTC_WARPS 4
#define CTC_WMAX 8  // widest beam of the backtrack_beam walk (the reference's default is 5, align() uses 2)
#define CTC_RING 8
#define CTC_PD 6

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// nanmax (torch.maximum: the first NaN operand wins, else the larger) as selects: no branch for the compiler to build blocks from
__device__ __forceinline__ float nanmax_sel(float a, float b) {
  float r;
  asm("{\n\t.reg .pred p, q;\n\t.reg .f32 m;\n\t"
      "max.f32 m, %1, %2;\n\t"
      "setp.nan.f32 q, %2, %2;\n\t"
      "selp.f32 m, %2, m, q;\n\t"
      "setp.nan.f32 p, %1, %1;\n\t"
      "selp.f32 %0, %1, m, p;\n\t}"
      : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// max over v != blank of row[v]  (warp-uniform result)
__device__ __forceinline__ float wild_max_global(const float* row, int V, int blank, int lane) {
  float m = -INFINITY;
  for (int v = lane; v < V; v += 32)
    if (v != blank) m = fmaxf(m, __ldg(row + v));
  return warp_max(m);
}

// WMAX: beam arrays sized (and loops unrolled) for widths <= WMAX: 2 keeps align()'s walk in registers.
// KREG > 0: every segment of the launch has N <= 32 KREG tokens and the trellis recurrence keeps the current row in REGISTERS:
// lane l owns columns l, l + 32, .. (a warp's store of one register is 128 contiguous bytes of the trellis row), the left
// neighbour's value arrives by one rotate-shuffle per column, the emission offset of every column's token is a register too,
// and a step is KREG independent {SHFL, LDS, 2 FADD, max, STG} groups with no shared-memory row, no per-column branch and no
// store -> load ordering.  (Consecutive columns per lane need one shuffle per STEP, but then every store instruction touches 32
// different sectors: measured 3.44 ms vs 4.55 ms for the shared-memory rows at N = 1040, store-transaction bound.)
// KREG = 0 is the general path (rows ping-pong in shared memory, any N that fits).  Same adds, same max per cell: same bits.
template <int WMAX, int KREG>
__global__ void __launch_bounds__(CTC_WARPS * 32)
ctc_align_kernel(const float* __restrict__ emis, const int* __restrict__ tok,
                 const CtcSeg* __restrict__ segs, int n_seg, int V, int Vpad, int blank, int mode,
                 int nmax_pad, int wpb, int beam_w, float* __restrict__ trellis, int2* __restrict__ hist,
                 int* __restrict__ path_tok, float* __restrict__ path_lp,
                 float* __restrict__ path_prob, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg_id = blockIdx.x * wpb + warp;  // wpb = warps (segments) per block, chosen by the host so the shared memory fits
  if (seg_id >= n_seg) return;
  // per-warp carve-up
  const size_t per_warp = (size_t)nmax_pad * 12 + (size_t)CTC_RING * Vpad * 4;
  unsigned char* base = smem_raw + per_warp * warp;
  int* stok = reinterpret_cast<int*>(base);
  float* rowA = reinterpret_cast<float*>(base + (size_t)nmax_pad * 4);
  float* rowB = reinterpret_cast<float*>(base + (size_t)nmax_pad * 8);
  float* ering = reinterpret_cast<float*>(base + (size_t)nmax_pad * 12);

  const CtcSeg sg = segs[seg_id];
  const int T = sg.T, N = sg.N;
  const float* E = emis + (size_t)sg.t_off * V;
  const int* tk = tok + sg.n_off;
  float* TR = trellis + sg.tr_off;
  int* o_tok = path_tok + sg.t_off;
  float* o_lp = path_lp + sg.t_off;
  float* o_pr = path_prob + sg.t_off;

  if (T <= 0 || N <= 0) {
    if (lane == 0) status[seg_id] = 1;
    return;
  }

  // ------------------------------------------------------------------ trellis (get_trellis)
  bool has_wild = false;
  for (int j = lane; j < N; j += 32) {
    int v = tk[j];
    stok[j] = v;
    has_wild |= (v < 0);
  }
  has_wild = __any_sync(0xffffffffu, has_wild);
  // trellis[-N+1:, 0] = +inf  (Python slice: N==1 -> rows 0..T-1; N-1 >= T -> all rows)
  int inf_start = (N == 1) ? 0 : (T - N + 1);
  if (inf_start < 0) inf_start = 0;

  if (KREG > 0) {
    __syncwarp();  // stok is complete
    constexpr int KR = KREG > 0 ? KREG : 1;
    const int kl = (N - lane + 31) >> 5;  // this lane's columns lane + 32 k, k < kl, exist
    float cur[KR];
    int toff[KR];  // byte offset of the column's emission inside a ring row; wildcards read the spare slot V (the step's wc)
#pragma unroll
    for (int k = 0; k < KR; ++k) {
      const int j = lane + 32 * k;
      const int tkn = (j < N) ? stok[j] : 0;
      toff[k] = 4 * (tkn < 0 ? V : tkn);
      cur[k] = (j == 0) ? ((0 >= inf_start) ? INFINITY : 0.f) : -INFINITY;
      if (k < kl) TR[j] = cur[k];
    }
    for (int r = 0; r < CTC_PD; ++r) {
      if (r < T)
        for (int v = lane; v < V; v += 32) cp_async4(ering + (r % CTC_RING) * Vpad + v, E + (size_t)r * V + v);
      cp_async_commit();
    }
    double acc = 0.0;  // lane 0 only: fp64 accumulator of the blank column (torch CPU cumsum)
    for (int t = 0; t < T - 1; ++t) {
      {
        const int r = t + CTC_PD;
        if (r < T)
          for (int v = lane; v < V; v += 32) cp_async4(ering + (r % CTC_RING) * Vpad + v, E + (size_t)r * V + v);
        cp_async_commit();
      }
      cp_async_wait<CTC_PD - 1>();  // rows <= t+1 have landed
      __syncwarp();
      float* er = ering + (t % CTC_RING) * Vpad;
      const float* er1 = ering + ((t + 1) % CTC_RING) * Vpad;
      const float eb = er[blank];
      if (has_wild) {
        float wc = -INFINITY;
        for (int v = lane; v < V; v += 32)
          if (v != blank) wc = fmaxf(wc, er[v]);
        wc = warp_max(wc);
        if (lane == 0) er[V] = wc;  // spare slot behind the V labels (Vpad > V)
        __syncwarp();
      }
      float c0 = 0.f;
      if (lane == 0) {
        acc += (double)er1[blank];
        c0 = (t + 1 >= inf_start) ? INFINITY : (float)acc;
      }
      float* trow = TR + (size_t)(t + 1) * N + lane;
      const char* erb = reinterpret_cast<const char*>(er);
      const int src = (lane + 31) & 31;  // rotate: lane l reads lane l - 1, lane 0 reads lane 31
      // column j - 1 of the current row: lane l - 1's cur[k]; for lane 0 it is lane 31's cur[k - 1].  All rotates first, then a
      // branch-free body (nanmax_sel, predicated stores): one basic block, so the scheduler overlaps the columns' LDS / FADD
      // chains.  With a shuffle or a branch inside every column's code each column was its own block: ~125 cycles apiece.
      float rot[KR];
#pragma unroll
      for (int k = 0; k < KR; ++k) rot[k] = __shfl_sync(0xffffffffu, cur[k], src);
#pragma unroll
      for (int k = 0; k < KR; ++k) {
        const float left = (lane == 0) ? rot[k > 0 ? k - 1 : 0] : rot[k];
        const float w = *reinterpret_cast<const float*>(erb + toff[k]);
        float nv = nanmax_sel(__fadd_rn(cur[k], eb), __fadd_rn(left, w));
        if (k == 0) nv = (lane == 0) ? c0 : nv;
        cur[k] = nv;
        if (k < kl) trow[32 * k] = nv;
      }
      __syncwarp();  // every lane is done with ring row t before a later cp.async lands on its slot
    }
    cp_async_wait<0>();
    __syncwarp();
  } else {
  float* prev = rowA;
  float* next = rowB;
  for (int j = lane; j < N; j += 32) {
    float v = (j == 0) ? ((0 >= inf_start) ? INFINITY : 0.f) : -INFINITY;
    prev[j] = v;
    TR[j] = v;
  }
  // prime the emission ring: rows 0 .. PD-1
  for (int r = 0; r < CTC_PD; ++r) {
    if (r < T)
      for (int v = lane; v < V; v += 32) cp_async4(ering + (r % CTC_RING) * Vpad + v, E + (size_t)r * V + v);
    cp_async_commit();
  }
  double acc = 0.0;  // lane 0 only: fp64 accumulator of the blank column (torch CPU cumsum)
  for (int t = 0; t < T - 1; ++t) {
    {
      const int r = t + CTC_PD;
      if (r < T)
        for (int v = lane; v < V; v += 32)
          cp_async4(ering + (r % CTC_RING) * Vpad + v, E + (size_t)r * V + v);
      cp_async_commit();
    }
    cp_async_wait<CTC_PD - 1>();  // rows <= t+1 have landed
    __syncwarp();
    const float* er = ering + (t % CTC_RING) * Vpad;
    const float* er1 = ering + ((t + 1) % CTC_RING) * Vpad;
    const float eb = er[blank];
    float wc = -INFINITY;
    if (has_wild) {
      for (int v = lane; v < V; v += 32)
        if (v != blank) wc = fmaxf(wc, er[v]);
      wc = warp_max(wc);
    }
    float c0 = 0.f;
    if (lane == 0) {
      acc += (double)er1[blank];
      c0 = (t + 1 >= inf_start) ? INFINITY : (float)acc;
    }
    float* trow = TR + (size_t)(t + 1) * N;
    for (int j = lane; j < N; j += 32) {
      float nv;
      if (j == 0) {
        nv = c0;
      } else {
        const int tkn = stok[j];
        const float w = (tkn < 0) ? wc : er[tkn];
        nv = nanmax(__fadd_rn(prev[j], eb), __fadd_rn(prev[j - 1], w));
      }
      next[j] = nv;
      trow[j] = nv;
    }
    __syncwarp();
    float* tmp = prev;
    prev = next;
    next = tmp;
  }
  cp_async_wait<0>();
  __syncwarp();
  }
  if (mode == WXB_CTC_TRELLIS_ONLY) {
    if (lane == 0) status[seg_id] = 0;
    return;
  }
  __threadfence_block();
  __syncwarp();

  // Both walks below run warp-uniform (every lane executes the same scalar logic on the same
  // addresses = broadcast loads); lane 0 alone writes.
  if (mode == WXB_CTC_BACKTRACK) {
    // -------------------------------------------------------------- backtrack (:447-481)
    int t = T - 1, j = N - 1;
    {
      float lp = E[(size_t)t * V + blank];
      if (lane == 0) { o_tok[t] = j; o_lp[t] = lp; o_pr[t] = expf(lp); }
    }
    bool fail = false;
    while (j > 0) {
      if (t <= 0) { fail = true; break; }  // reference: assert t > 0
      const float* erow = E + (size_t)(t - 1) * V;
      const float p_stay = __ldg(erow + blank);
      const int tkn = stok[j];
      const float* trow = TR + (size_t)(t - 1) * N;
      const float tr_j = trow[j], tr_j1 = trow[j - 1];  // requested together with the emissions: one round trip per step
      const float p_change = (tkn < 0) ? wild_max_global(erow, V, blank, lane) : __ldg(erow + tkn);
      const float stayed = __fadd_rn(tr_j, p_stay);
      const float changed = __fadd_rn(tr_j1, p_change);
      t -= 1;
      const bool ch = changed > stayed;
      if (ch) j -= 1;
      const float lp = ch ? p_change : p_stay;
      if (lane == 0) { o_tok[t] = j; o_lp[t] = lp; o_pr[t] = expf(lp); }
    }
    if (fail) {
      if (lane == 0) status[seg_id] = 1;
      return;
    }
    while (t > 0) {
      const float lp = __ldg(E + (size_t)(t - 1) * V + blank);
      if (lane == 0) { o_tok[t - 1] = j; o_lp[t - 1] = lp; o_pr[t - 1] = expf(lp); }
      t -= 1;
    }
    if (lane == 0) status[seg_id] = 0;
    return;
  }

  // ------------------------------------------------------------------ beam search of width W (:500-579; align() uses W = 2)
  // All live beams share the same time index.  hist[k*W+slot] = (j | parent<<28, lp bits) is the point appended at step k
  // (time T-1-k) by the beam that ends up in `slot` after the sort.  Candidates are generated beam by beam as [stay, change]
  // and the first W of a STABLE descending sort on the predecessor cell's score survive (Python's sorted(..., reverse=True):
  // ties keep generation order; duplicates of one cell are not merged).
  const int W = beam_w;
  int2* H = hist + (size_t)sg.t_off * W;
  int bj[WMAX];
  int nb = 1;
  int t = T - 1;
  bj[0] = N - 1;
  const float lp0 = E[(size_t)t * V + blank];
  int K = 0;
  while (nb > 0 && bj[0] > 0) {
    // Candidates sit in FIXED slots (2 i = beam i stays, 2 i + 1 = beam i changes) with a validity mask, so every array index
    // below is a compile-time constant and the step's state lives in registers; compacting them with a running count put the
    // four arrays in local memory.  Slot order = the reference's generation order, which is what the stable sort keeps on ties.
    int cj[2 * WMAX];
    float cs[2 * WMAX], cl[2 * WMAX];
    unsigned valid = 0u;
#pragma unroll
    for (int i = 0; i < 2 * WMAX; ++i) { cj[i] = 0; cs[i] = 0.f; cl[i] = 0.f; }
    if (t > 0) {
      const float* erow = E + (size_t)(t - 1) * V;
      const float* trow = TR + (size_t)(t - 1) * N;
      const float p_stay = __ldg(erow + blank);
      // Everything a step reads depends only on the beams' positions: the predecessor cells of every beam and the emission
      // of its token are requested together (one L2 round trip per step).  Read one after the other behind the isinf tests,
      // they were three dependent round trips per beam, ~5000 cycles per step.
      float ss[WMAX], sc[WMAX], pc[WMAX];
#pragma unroll
      for (int i = 0; i < WMAX; ++i) {
        if (i < nb) {
          const int j = bj[i];
          const int tkn = stok[j];
          ss[i] = trow[j];
          sc[i] = trow[j > 0 ? j - 1 : 0];
          pc[i] = __ldg(erow + (tkn < 0 ? blank : tkn));
        }
      }
#pragma unroll
      for (int i = 0; i < WMAX; ++i) {
        if (i < nb) {
          const int j = bj[i];
          if (!isinf(ss[i])) { cj[2 * i] = j; cs[2 * i] = ss[i]; cl[2 * i] = p_stay; valid |= 1u << (2 * i); }
          if (j > 0 && !isinf(sc[i])) {
            const float p_change = (stok[j] < 0) ? wild_max_global(erow, V, blank, lane) : pc[i];
            cj[2 * i + 1] = j - 1; cs[2 * i + 1] = sc[i]; cl[2 * i + 1] = p_change; valid |= 1u << (2 * i + 1);
          }
        }
      }
    }
    // stable descending top-W: repeatedly take the first not-yet-taken candidate with the largest score
    const int nc = __popc(valid);
    nb = nc < W ? nc : W;
    t -= 1;
    K += 1;
#pragma unroll
    for (int s2 = 0; s2 < WMAX; ++s2) {
      if (s2 < nb) {
        int best = -1;
        float best_s = 0.f;
        int best_j = 0;
        float best_l = 0.f;
#pragma unroll
        for (int i = 0; i < 2 * WMAX; ++i)
          if (((valid >> i) & 1u) && (best < 0 || cs[i] > best_s)) { best = i; best_s = cs[i]; best_j = cj[i]; best_l = cl[i]; }
        valid &= ~(1u << best);
        bj[s2] = best_j;
        if (lane == 0) H[(size_t)K * W + s2] = make_int2(best_j | ((best >> 1) << 28), __float_as_int(best_l));
      }
    }
  }
  if (nb == 0) {
    if (lane == 0) status[seg_id] = 1;  // reference returns None -> "backtrack failed"
    return;
  }
  __syncwarp();
  // pad the best beam down to t = 0 with blank emissions (:571-576)
  {
    const int j = bj[0];
    for (int tt = t - 1 - lane; tt >= 0; tt -= 32) {
      const float lp = __ldg(E + (size_t)tt * V + blank);
      o_tok[tt] = j; o_lp[tt] = lp; o_pr[tt] = expf(lp);
    }
  }
  // walk the parent chain of the best beam back to the initial point
  {
    int slot = 0;
    for (int k = K; k >= 1; --k) {
      const int2 h = H[(size_t)k * W + slot];
      const int j = h.x & 0x0fffffff;
      const float lp = __int_as_float(h.y);
      const int tt = T - 1 - k;
      if (lane == 0) { o_tok[tt] = j; o_lp[tt] = lp; o_pr[tt] = expf(lp); }
      slot = (h.x >> 28) & 7;
    }
    if (lane == 0) { o_tok[T - 1] = N - 1; o_lp[T - 1] = lp0; o_pr[T - 1] = expf(lp0); status[seg_id] = 0; }
  }
}

__global__ void log_softmax_rows_kernel(float* __restrict__ x, long long rows, int V) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* p = x + row * V;
  float m = -INFINITY;
  for (int v = lane; v < V; v += 32) m = fmaxf(m, p[v]);
  m = warp_max(m);
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s += expf(p[v] - m);
  s = warp_sum(s);
  const float lse = m + logf(s);
  for (int v = lane; v < V; v += 32) p[v] = p[v] - lse;
}

extern "C" {

int wxb_ctc_align(wxb_ctx* ctx, const float* emis_dev, const int32_t* t_off_host,
                  const int32_t* tok_dev, const int32_t* n_off_host, int n_seg, int V, int blank,
                  int mode, float* trellis_dev, int32_t* path_tok_dev, float* path_lp_dev,
                  float* path_prob_dev, int32_t* status_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (n_seg == 0) return WXB_OK;
  if (!emis_dev || !t_off_host || !tok_dev || !n_off_host || n_seg < 0 || V <= 0 || blank < 0 ||
      blank >= V || !status_dev)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_ctc_align: bad argument");
  // mode = base | beam_width << 8; a beam walk with no width given is the width align() uses (2)
  int beam_w = (mode >> 8) & 0xff;
  mode &= 0xff;
  if (mode != WXB_CTC_BACKTRACK && mode != WXB_CTC_BEAM2 && mode != WXB_CTC_TRELLIS_ONLY)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_ctc_align: unknown mode %d", mode);
  if (mode == WXB_CTC_BEAM2 && beam_w == 0) beam_w = 2;
  if (mode == WXB_CTC_BEAM2 && (beam_w < 1 || beam_w > CTC_WMAX))
    return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_ctc_align: beam width %d (supported: 1..%d)", beam_w, CTC_WMAX);
  if (mode != WXB_CTC_BEAM2) beam_w = 1;
  if (mode != WXB_CTC_TRELLIS_ONLY && (!path_tok_dev || !path_lp_dev || !path_prob_dev))
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_ctc_align: path outputs are NULL");
  cudaStream_t st = (cudaStream_t)stream;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<CtcSeg> segs(n_seg);
  long long tr_total = 0;
  int nmax = 1;
  for (int s = 0; s < n_seg; ++s) {
    CtcSeg& g = segs[s];
    g.t_off = t_off_host[s];
    g.T = t_off_host[s + 1] - t_off_host[s];
    g.n_off = n_off_host[s];
    g.N = n_off_host[s + 1] - n_off_host[s];
    if (g.T < 0 || g.N < 0) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_ctc_align: offsets not monotone");
    if (g.N >= (1 << 28)) return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_ctc_align: N too large");
    g.tr_off = tr_total;
    tr_total += (long long)g.T * g.N;
    if (g.N > nmax) nmax = g.N;
  }
  const long long sumT = t_off_host[n_seg];
  const int nmax_pad = (nmax + 31) & ~31;
  const int Vpad = (V + 1 + 3) & ~3;  // ring row stride: the V labels + a spare slot for the step's wildcard emission
  // shared memory per warp: token ids + two trellis rows + the emission ring; large vocabularies (the ja / zh align models
  // have 2-3.5 k labels) get fewer warps per block instead of an error
  const size_t per_warp = (size_t)nmax_pad * 12 + (size_t)CTC_RING * Vpad * 4;
  int wpb = CTC_WARPS;
  while (wpb > 1 && per_warp * wpb > 220 * 1024) wpb >>= 1;
  const size_t smem = per_warp * wpb;
  if (smem > 220 * 1024)
    return wxb_fail(ctx, WXB_ERR_UNSUPPORTED, "wxb_ctc_align: a segment with %d tokens over %d labels needs %zu B of shared memory", nmax, V, smem);
  int rc;
  if ((rc = wxb_reserve(ctx, ctx->ws_ctc_meta, sizeof(CtcSeg) * n_seg)) != WXB_OK) return rc;
  if (!trellis_dev) {
    if ((rc = wxb_reserve(ctx, ctx->ws_ctc_trellis, sizeof(float) * (size_t)(tr_total > 0 ? tr_total : 1))) != WXB_OK) return rc;
    trellis_dev = (float*)ctx->ws_ctc_trellis.p;
  }
  if ((rc = wxb_reserve(ctx, ctx->ws_ctc_hist, sizeof(int2) * (size_t)beam_w * (size_t)(sumT + 1))) != WXB_OK) return rc;
  WXB_CUDA(ctx, cudaMemcpyAsync(ctx->ws_ctc_meta.p, segs.data(), sizeof(CtcSeg) * n_seg, cudaMemcpyHostToDevice, st));
  // the pageable source buffer `segs` dies at return: the copy above is staged synchronously by
  // the runtime for pageable memory, so this is safe.
  // register-resident recurrence when every segment fits 32 KREG columns (KREG in {8, 16, 34}), shared-memory rows otherwise
  const int kreg = nmax <= 256 ? 8 : nmax <= 512 ? 16 : nmax <= 1088 ? 34 : 0;
  auto kern = beam_w <= 2 ? (kreg == 8 ? ctc_align_kernel<2, 8> : kreg == 16 ? ctc_align_kernel<2, 16> : kreg == 34 ? ctc_align_kernel<2, 34> : ctc_align_kernel<2, 0>)
                          : (kreg == 8 ? ctc_align_kernel<CTC_WMAX, 8> : kreg == 16 ? ctc_align_kernel<CTC_WMAX, 16>
                             : kreg == 34 ? ctc_align_kernel<CTC_WMAX, 34> : ctc_align_kernel<CTC_WMAX, 0>);
#ifdef WXB_PROBE
  if (getenv("WXB_CTC_SMEM_ROWS")) kern = beam_w <= 2 ? ctc_align_kernel<2, 0> : ctc_align_kernel<CTC_WMAX, 0>;  // A/B: the general path
#endif
  if ((rc = wxb_func_smem(ctx, kern, (int)smem)) != WXB_OK) return rc;
  const int grid = ceil_div(n_seg, wpb);
  kern<<<grid, wpb * 32, smem, st>>>(
      emis_dev, tok_dev, (const CtcSeg*)ctx->ws_ctc_meta.p, n_seg, V, Vpad, blank, mode, nmax_pad, wpb, beam_w,
      trellis_dev, (int2*)ctx->ws_ctc_hist.p, path_tok_dev, path_lp_dev, path_prob_dev, status_dev);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}

int wxb_log_softmax_rows(wxb_ctx* ctx, float* x_dev, int64_t rows, int V, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (rows == 0) return WXB_OK;
  if (!x_dev || rows < 0 || V <= 0) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_log_softmax_rows: bad argument");
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int wpb = 8;
  log_softmax_rows_kernel<<<(unsigned)ceil_div64(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(x_dev, rows, V);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}

}  // extern "C"
