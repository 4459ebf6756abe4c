// wxb_vad.cu — VAD post-processing and chunking on the GPU: frame scores in HBM -> speech regions -> <= chunk_size
// chunks -> the (offset, length) table K1 consumes, without the scores ever visiting the host.
//
// Reference:
//   /root/reference/whisperx/vads/pyannote.py:134-216   Binarize.__call__ (hysteresis thresholding + WhisperX min-cut at max_duration)
//   /root/reference/whisperx/vads/pyannote.py:282-301   Pyannote.merge_chunks (Binarize(max_duration=chunk_size) -> timeline -> merge)
//   /root/reference/whisperx/vads/vad.py:20-53          Vad.merge_chunks (greedy merge of speech regions into chunks)
//   /root/reference/whisperx/asr.py:70-73               chunk audio = audio[int(start * 16000) : int(end * 16000)]
//
// The reference walks the frames one by one in Python.  Here one WARP owns a recording: 32 frames are tested per step against the
// predicate of the current state (active: "region too long" or "score < offset"; inactive: "score > onset"), a ballot finds
// the first frame where something happens, everything before it is skipped in one go.  The min-cut's argmin over the second
// half of the open region is a strided warp reduction that keeps numpy's first-minimum rule.  All time arithmetic is IEEE double
// with explicit round-to-nearest intrinsics (no FMA contraction), so region and chunk boundaries are bit-identical to the
// reference's Python floats; the score tests are float32 compares like numpy's.
//
// The reference's curr_scores / curr_timestamps lists are the frame set {head} + [lo, hi]: `head` is the stale first element the
// Python code leaves in the lists (frame 0, or the frame that closed the previous region), [lo, hi] the frames appended since.
#include "wxb_common.cuh"
#include <math.h>

namespace {

__device__ __forceinline__ double frame_mid(long long i, double start, double step, double duration) {
  const double s = __dadd_rn(start, __dmul_rn((double)i, step));   // SlidingWindow[i].start
  return __dmul_rn(0.5, __dadd_rn(s, __dadd_rn(s, duration)));     // Segment.middle
}

struct Cand {
  float v;
  int p;
};
// numpy argmin order: NaN beats everything, then the smaller value, then the smaller position
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {
  const bool an = a.v != a.v, bn = b.v != b.v;
  if (an != bn) return an;
  if (an) return a.p < b.p;
  return a.v < b.v || (a.v == b.v && a.p < b.p);
}

__global__ void __launch_bounds__(32) vad_chunks_kernel(const float* __restrict__ scores, const long long* __restrict__ score_off,
                                                        const long long* __restrict__ n_samples, wxb_vad_params prm, int max_regions,
                                                        int max_chunks, int sample_rate, int max_chunk_samples,
                                                        double* __restrict__ regions, int* __restrict__ n_regions,
                                                        double* __restrict__ chunks, int* __restrict__ chunk_first,
                                                        int* __restrict__ n_chunks, long long* __restrict__ chunk_off,
                                                        int* __restrict__ chunk_len) {
  const int rec = blockIdx.x, lane = threadIdx.x;
  const float* y = scores + score_off[rec];
  const long long n = score_off[rec + 1] - score_off[rec];
  double* reg = regions + (size_t)rec * max_regions * 2;
  const float onset = prm.onset, offset = prm.offset;
  const double maxd = prm.chunk_size, fs = prm.frame_start, fstep = prm.frame_step, fdur = prm.frame_duration;
  int nreg = 0;
  auto emit = [&](double a, double b) {
    if (__dsub_rn(b, a) > 1e-6) {  // pyannote: an empty segment is never stored
      if (lane == 0 && nreg < max_regions) { reg[2 * nreg] = a; reg[2 * nreg + 1] = b; }
      ++nreg;  // counted even past the capacity: the host sees the overflow
    }
  };
  if (n > 0) {
    double region_start = frame_mid(0, fs, fstep, fdur);
    bool active = y[0] > onset;
    bool has_head = true;
    long long head = 0, lo = 1, hi = 0;
    long long i = 1;
    while (i < n) {
      const long long ii = i + lane;
      bool pred = false;
      double t = 0.0;
      float yv = 0.f;
      if (ii < n) {
        t = frame_mid(ii, fs, fstep, fdur);
        yv = y[ii];
        pred = active ? (__dsub_rn(t, region_start) > maxd || yv < offset) : (yv > onset);
      }
      const unsigned m = __ballot_sync(0xffffffffu, pred);
      const int ev = m ? __ffs(m) - 1 : 32;  // lanes before the event change nothing but the open list
      const long long valid = (n - i) < 32 ? (n - i) : 32;
      const long long skipped = ev < valid ? ev : valid;
      if (active && skipped > 0) {  // frames i .. i + skipped - 1 are appended
        if (hi < lo) lo = i;
        hi = i + skipped - 1;
      }
      if (ev >= valid) { i += valid; continue; }
      const long long e = i + ev;
      const double te = __shfl_sync(0xffffffffu, t, ev);
      const float ye = __shfl_sync(0xffffffffu, yv, ev);
      if (active) {
        if (__dsub_rn(te, region_start) > maxd) {
          // min-cut: first minimum over the second half of the list {head} + [lo, hi]
          const long long len = (has_head ? 1 : 0) + (hi >= lo ? hi - lo + 1 : 0);
          const long long half = len / 2;
          Cand best = {INFINITY, 0x7fffffff};
          bool any = false;
          for (long long p = half + lane; p < len; p += 32) {
            const long long f = has_head ? (p == 0 ? head : lo + p - 1) : lo + p;
            const Cand c = {y[f], (int)(p - half)};
            if (!any || better(c, best)) { best = c; any = true; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            Cand c;
            c.v = __shfl_xor_sync(0xffffffffu, best.v, o);
            c.p = __shfl_xor_sync(0xffffffffu, best.p, o);
            const bool c_any = __shfl_xor_sync(0xffffffffu, (int)any, o) != 0;
            if (c_any && (!any || better(c, best))) { best = c; any = true; }
          }
          const long long cut = half + best.p;  // list position (len >= 1 here: the list is never empty while active)
          const long long fcut = has_head ? (cut == 0 ? head : lo + cut - 1) : lo + cut;
          const double tcut = frame_mid(fcut, fs, fstep, fdur);
          emit(region_start, tcut);
          region_start = tcut;
          if (!(has_head && cut == 0)) lo = fcut + 1;  // everything up to the cut leaves the list
          has_head = false;
        } else {  // ye < offset: the region closes at this frame
          emit(region_start, te);
          region_start = te;
          active = false;
          has_head = false;
          lo = e + 1; hi = e;  // empty; the append below turns the closing frame into the stale head
        }
        // the event frame is appended in both sub-cases
        if (!active) { has_head = true; head = e; lo = e + 1; hi = e; }
        else { if (hi < lo) lo = e; hi = e; }
      } else {  // ye > onset: a region opens; the lists keep their stale head and grow from the NEXT frame
        region_start = te;
        active = true;
        lo = e + 1; hi = e;
      }
      i = e + 1;
    }
    if (active) emit(region_start, n > 1 ? frame_mid(n - 1, fs, fstep, fdur) : region_start);
  }
  __syncwarp();
  // ---- Vad.merge_chunks over the regions (sequential: a few hundred entries), then the K1 table -------------------------
  if (lane == 0) {
    n_regions[rec] = nreg;
    const int nr = nreg < max_regions ? nreg : max_regions;
    double* ch = chunks + (size_t)rec * max_chunks * 2;
    int* first = chunk_first + (size_t)rec * (max_chunks + 1);
    int nch = 0;
    if (nr > 0) {
      double cur_start = reg[0], cur_end = 0.0;
      int first_member = 0;
      for (int k = 0; k < nr; ++k) {
        const double s = reg[2 * k], e = reg[2 * k + 1];
        if (__dsub_rn(e, cur_start) > maxd && __dsub_rn(cur_end, cur_start) > 0.0) {
          if (nch < max_chunks) { ch[2 * nch] = cur_start; ch[2 * nch + 1] = cur_end; first[nch] = first_member; }
          ++nch;
          cur_start = s;
          first_member = k;
        }
        cur_end = e;
      }
      if (nch < max_chunks) { ch[2 * nch] = cur_start; ch[2 * nch + 1] = cur_end; first[nch] = first_member; }
      ++nch;
    }
    n_chunks[rec] = nch;
    const int nc = nch < max_chunks ? nch : max_chunks;
    first[nc] = nr;
    for (int k = 0; k < nc; ++k) {
      // asr.py:70-73: audio[int(start * 16000) : int(end * 16000)] (Python slicing clips at the end of the recording)
      long long a = (long long)__dmul_rn(ch[2 * k], (double)sample_rate), b = (long long)__dmul_rn(ch[2 * k + 1], (double)sample_rate);
      const long long ns = n_samples[rec];
      a = a < 0 ? 0 : (a > ns ? ns : a);
      b = b < a ? a : (b > ns ? ns : b);
      long long len = b - a;
      if (len > max_chunk_samples) len = max_chunk_samples;  // K1 takes at most one 30 s window per chunk (pad_or_trim)
      chunk_off[(size_t)rec * max_chunks + k] = a;
      chunk_len[(size_t)rec * max_chunks + k] = (int)len;
    }
  }
}

// Stand-in frame scorer (no VAD checkpoint exists offline): score = sigmoid((10 log10(mean x^2 + 1e-10) - floor_db) / width_db)
// over frames of 400 samples every 160; one warp per frame, coalesced reads, double accumulation in a fixed order.
__global__ void __launch_bounds__(256) vad_energy_kernel(const float* __restrict__ audio, long long n_samples, long long n_frames,
                                                         float floor_db, float width_db, float* __restrict__ scores) {
  const long long frame = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (frame >= n_frames) return;
  const long long s0 = frame * 160;
  double acc = 0.0;
  for (int k = lane; k < 400; k += 32) {
    const long long s = s0 + k;
    const double v = s < n_samples ? (double)audio[s] : 0.0;
    acc = __dadd_rn(acc, __dmul_rn(v, v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  if (lane == 0) {
    const float db = (float)(10.0 * log10(acc / 400.0 + 1e-10));
    scores[frame] = 1.0f / (1.0f + expf(-(db - floor_db) / width_db));
  }
}

}  // namespace

extern "C" int wxb_vad_energy_scores(wxb_ctx* ctx, const float* audio_dev, int64_t n_samples, float floor_db, float width_db,
                                     float* scores_out_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!audio_dev || n_samples <= 0 || !scores_out_dev || !(width_db > 0.f)) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_vad_energy_scores: bad argument");
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  const long long nf = (n_samples + 159) / 160;
  vad_energy_kernel<<<(unsigned)((nf + 7) / 8), 256, 0, (cudaStream_t)stream>>>(audio_dev, n_samples, nf, floor_db, width_db, scores_out_dev);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}

extern "C" int64_t wxb_vad_energy_frames(int64_t n_samples) { return n_samples > 0 ? (n_samples + 159) / 160 : 0; }

extern "C" int wxb_vad_chunks(wxb_ctx* ctx, const float* scores_dev, const int64_t* score_off_host, const int64_t* n_samples_host,
                              int n_rec, const wxb_vad_params* prm, int max_regions, int max_chunks, double* regions_dev,
                              int32_t* n_regions_dev, double* chunks_dev, int32_t* chunk_first_dev, int32_t* n_chunks_dev,
                              int64_t* chunk_off_dev, int32_t* chunk_len_dev, void* stream) {
  if (!ctx) return WXB_ERR_INVALID;
  if (!scores_dev || !score_off_host || !n_samples_host || n_rec <= 0 || n_rec > 4096 || !prm || max_regions <= 0 || max_chunks <= 0 ||
      !regions_dev || !n_regions_dev || !chunks_dev || !chunk_first_dev || !n_chunks_dev || !chunk_off_dev || !chunk_len_dev)
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_vad_chunks: bad argument");
  if (!(prm->frame_step > 0.0) || !(prm->frame_duration > 0.0) || !(prm->chunk_size > 2.0 * prm->frame_step))
    return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_vad_chunks: frame_step / frame_duration must be positive and chunk_size > 2 frame steps");
  if (!(prm->onset > 0.f && prm->onset < 1.f)) return wxb_fail(ctx, WXB_ERR_INVALID, "vad_onset is a decimal value between 0 and 1.");
  cudaStream_t st = (cudaStream_t)stream;
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  long long* tab = (long long*)wxb_named(ctx, "vad.tab", (size_t)(2 * 4096 + 2) * 8);
  if (!tab) return WXB_ERR_CUDA;
  std::vector<long long> h((size_t)2 * n_rec + 1);
  for (int r = 0; r <= n_rec; ++r) h[r] = score_off_host[r];
  for (int r = 0; r < n_rec; ++r) {
    if (score_off_host[r + 1] < score_off_host[r] || n_samples_host[r] < 0) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_vad_chunks: bad offsets for recording %d", r);
    h[n_rec + 1 + r] = n_samples_host[r];
  }
  WXB_CUDA(ctx, cudaStreamSynchronize(st));  // the small table is pageable host memory
  WXB_CUDA(ctx, cudaMemcpy(tab, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  wxb_vad_params p = *prm;
  if (!(p.offset > 0.f)) p.offset = p.onset;  // Binarize: offset = offset or onset
  vad_chunks_kernel<<<n_rec, 32, 0, st>>>(scores_dev, tab, tab + n_rec + 1, p, max_regions, max_chunks, 16000, 480000, regions_dev,
                                          n_regions_dev, chunks_dev, chunk_first_dev, n_chunks_dev, (long long*)chunk_off_dev, chunk_len_dev);
  WXB_LAUNCH_CHECK(ctx);
  return WXB_OK;
}
