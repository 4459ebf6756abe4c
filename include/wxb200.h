/*
 * wxb200.h — C-ABI of the B200-native WhisperX hot path (libwxb200.so).
 *
 * The reference (sooth/whisperx-mlx) has no FFI: its plugin boundary is the duck-typed
 * Python backend interface (whisperx/backends/base.py:8-57, whisperx/asr.py:67-87) and the
 * module-level aligner functions (whisperx/alignment.py:113,387,447,500).  This header is the
 * C-ABI that sits directly under our Python mirror of that interface; every entry point names
 * the reference function whose numeric core it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary.
 *   - pointers named *_dev are device pointers on the ctx's GPU, *_host are host pointers.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*, NULL = legacy
 *     default stream) unless stated otherwise; buffers are caller-owned and must stay alive
 *     until the stream reaches the call.
 *   - return value 0 = ok, negative = error; wxb_last_error() gives the message.  Nothing
 *     throws across the ABI.  There is no CPU fallback: without a usable sm_100 GPU
 *     wxb_create() fails.
 *   - a wxb_ctx is bound to one device and is NOT thread-safe (one ctx per process per GPU).
 */
#ifndef WXB200_H
#define WXB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WXB_ABI_VERSION 3

typedef struct wxb_ctx wxb_ctx;

enum {
  WXB_OK = 0,
  WXB_ERR_INVALID = -1,   /* bad argument */
  WXB_ERR_CUDA = -2,      /* CUDA runtime / driver error */
  WXB_ERR_UNSUPPORTED = -3, /* device is not sm_100 or a shape is outside the kernel's range */
  WXB_ERR_STATE = -4      /* call made in the wrong order (e.g. encode before set_model) */
};

/* ------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------- */
int wxb_abi_version(void);
/* Create a context on CUDA device `device`.  Fails with WXB_ERR_UNSUPPORTED if the device's
 * compute capability is not 10.x. */
int wxb_create(int device, wxb_ctx** out);
void wxb_destroy(wxb_ctx* ctx);
/* Library-owned string, valid until the next call on ctx.  ctx may be NULL (global slot used
 * by wxb_create failures). */
const char* wxb_last_error(const wxb_ctx* ctx);
/* Number of kernel launches issued by this ctx since creation (bench.py's gpu_launches). */
int64_t wxb_launch_count(const wxb_ctx* ctx);

/* Bring-up aids (tools/w2v_debug.py; no effect on results): wxb_debug_set(ctx, "w2v_stop", s) makes wxb_w2v_emissions return
 * after stage s (0..6 conv layer, 7 feature projection, 8 positional conv, 9 + l transformer layer l; -1 = off);
 * wxb_debug_copy copies `bytes` of the ctx workspace called `name` (e.g. "w2v.c3", "w2v.x") to a device buffer. */
int wxb_debug_set(wxb_ctx* ctx, const char* key, int value);
int wxb_debug_copy(wxb_ctx* ctx, const char* name, void* dst_dev, int64_t offset, int64_t bytes);

/* ------------------------------------------------------------------------------------------
 * K1  log-mel frontend — replaces whisperx/audio.py:112-159 log_mel_spectrogram
 *     (reflect-padded STFT n_fft=400 hop=160 periodic Hann, |X|^2, mel filterbank,
 *      log10(max(.,1e-10)), max(., per-chunk max - 8), (.+4)/4).
 *
 * audio_dev      f32, all chunks back to back
 * chunk_off_host int64[n_chunks]  first sample of each chunk inside audio_dev
 * chunk_len_host int32[n_chunks]  valid samples of each chunk (<= n_samples_padded)
 * n_samples_padded  every chunk is treated as zero-padded on the right to this many samples
 *                   before the STFT (audio.py:147-148 `padding`); >= 400.
 *                   n_frames = n_samples_padded / 160 (floor; the reference drops the last frame).
 * filters_dev    f32 [n_mels, 201] (assets/mel_filters.npz), n_mels in {80,128} (any <= 128 ok)
 * mel_out_dev    f32 [n_chunks, n_mels, n_frames]   (reference layout, audio.py:159)
 * The max used by the clamp is PER CHUNK (the reference is called once per chunk).
 * ---------------------------------------------------------------------------------------- */
int wxb_logmel(wxb_ctx* ctx, const float* audio_dev, const int64_t* chunk_off_host,
               const int32_t* chunk_len_host, int n_chunks, int n_samples_padded, int n_mels,
               const float* filters_dev, float* mel_out_dev, void* stream);

/* K1 -> K2 hand-off on the device for 30 s chunks (n_samples_padded = 480000): same computation, but the result is ALSO left
 * in the context's encoder input buffer (bf16, frame-major) so that wxb_encode(ctx, NULL, n_chunks, ...) can consume it
 * without an f32 round trip through HBM.  mel_out_dev (f32 [n_chunks, n_mels, 3000]) may be NULL. */
int wxb_logmel_features(wxb_ctx* ctx, const float* audio_dev, const int64_t* chunk_off_host,
                        const int32_t* chunk_len_host, int n_chunks, int n_mels, const float* filters_dev,
                        float* mel_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4  CTC forced alignment — replaces whisperx/alignment.py:387-404 get_trellis,
 *     :407-437 get_wildcard_emission, :447-481 backtrack, :500-579 backtrack_beam(beam_width=2)
 *
 * emis_dev   f32 [sum T, V] log-probabilities (already log_softmax'ed, alignment.py:258),
 *            segments back to back
 * t_off_host int32[n_seg+1]   frame offsets of each segment inside emis_dev
 * tok_dev    int32[sum N]     token ids per segment back to back, -1 = wildcard
 * n_off_host int32[n_seg+1]   token offsets
 * mode       WXB_CTC_BACKTRACK (alignment.py:447), WXB_CTC_BEAM2 (alignment.py:500, width 2 as align() calls it) or WXB_CTC_BEAM(w)
 * trellis_dev   optional f32 buffer of sum(T_i*N_i) elements receiving every segment's trellis
 *               (row-major [T_i, N_i], back to back, offsets = prefix sums of T_i*N_i);
 *               NULL = use the ctx workspace.
 * path_tok_dev  int32[sum T]  token_index of the path point at each frame
 * path_lp_dev   f32[sum T]    log-prob of that point (exact emission value; score = exp(lp))
 * path_prob_dev f32[sum T]    expf(lp) computed on the device (may differ from torch's exp by
 *                             an ulp; indices and lp are bit-exact)
 * status_dev    int32[n_seg]  0 ok, 1 = "backtrack failed" (alignment.py:271 / assert :451)
 * ---------------------------------------------------------------------------------------- */
enum { WXB_CTC_BACKTRACK = 0, WXB_CTC_BEAM2 = 1, WXB_CTC_TRELLIS_ONLY = 2 };
/* backtrack_beam with another beam_width w in 1..8 (the reference's default is 5): mode = WXB_CTC_BEAM(w) */
#define WXB_CTC_BEAM(w) (WXB_CTC_BEAM2 | ((w) << 8))

int wxb_ctc_align(wxb_ctx* ctx, const float* emis_dev, const int32_t* t_off_host,
                  const int32_t* tok_dev, const int32_t* n_off_host, int n_seg, int V, int blank,
                  int mode, float* trellis_dev, int32_t* path_tok_dev, float* path_lp_dev,
                  float* path_prob_dev, int32_t* status_dev, void* stream);

/* In-place log_softmax over the last dim of f32 [rows, V] (alignment.py:258), one warp per row
 * with a warp-shuffle logsumexp.  Provided so emissions can stay on the device. */
int wxb_log_softmax_rows(wxb_ctx* ctx, float* x_dev, int64_t rows, int V, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2/K3  Whisper model — replaces the mlx_whisper model the reference delegates to
 *     (call sites: whisperx/backends/mlx_lightning.py:163-196, mlx_whisper_batch_decoder.py:
 *      37-100 logits, :267-303 update, :317-384 main loop).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_mels;        /* 80 | 128 */
  int32_t n_audio_ctx;   /* 1500 */
  int32_t n_audio_state; /* d */
  int32_t n_audio_head;
  int32_t n_audio_layer;
  int32_t n_vocab;
  int32_t n_text_ctx;    /* 448 */
  int32_t n_text_state;
  int32_t n_text_head;
  int32_t n_text_layer;
} wxb_dims;

/* Weight table: `names[i]` identifies tensor i (OpenAI-Whisper parameter names, e.g.
 * "encoder.blocks.0.attn.query.weight"); ptrs_dev[i] is a device pointer to a contiguous
 * tensor of the documented dtype (bf16 matrices, f32 biases / LN / positional tables; see
 * DESIGN.md).  The library BORROWS the pointers: the caller keeps the tensors alive until
 * wxb_destroy or the next wxb_set_model. */
int wxb_set_model(wxb_ctx* ctx, const wxb_dims* dims, const char* const* names,
                  const void* const* ptrs_dev, int n_tensors);

/* Encoder: mel f32 [B, n_mels, 3000] (output layout of wxb_logmel) -> enc_out bf16 [B,1500,d].
 * mel_dev = NULL: encode the B chunks whose log-mel the last wxb_logmel_features call left on the device. */
int wxb_encode(wxb_ctx* ctx, const float* mel_dev, int B, void* enc_out_dev, void* stream);

typedef struct {
  int32_t eot;             /* end-of-transcript token id */
  int32_t no_speech;       /* <|nospeech|> id, or -1 */
  int32_t sample_len;      /* max sampled tokens (reference: n_text_ctx // 2 = 224) */
  int32_t suppress_blank;  /* 1: at the first sampled position suppress blank_token and eot */
  int32_t blank_token;     /* id of encode(" ") (only used when suppress_blank) */
  int32_t n_suppress;      /* length of suppress_dev */
  const int32_t* suppress_dev; /* device int32 ids set to -inf every step */
  int32_t check_every;     /* host polls the EOT flags every this many steps (0 = 16); rows that have finished leave
                              the batch at these boundaries (active-sequence compaction,
                              mlx_whisper_batch_decoder.py:37-100) */
  int32_t no_compaction;   /* 1: finished rows keep riding along until every row is done (A/B timing, tests) */
  /* ApplyTimestampRules (decode with timestamps, mlx_lightning.py:187-193 `without_timestamps=False`); the last
   * clause is the reference's batch-safe patch mlx_ultra_optimized_batch.py:38-71 */
  int32_t apply_timestamp_rules;        /* 0 = off (prompt ends with <|notimestamps|>) */
  int32_t timestamp_begin;              /* id of <|0.00|> */
  int32_t no_timestamps;                /* id of <|notimestamps|>, suppressed when the rules are on (-1 = none) */
  int32_t max_initial_timestamp_index;  /* 50 = 1.0 s / 0.02 s; < 0 = unlimited */
} wxb_decode_opts;

/* Batched greedy KV-cache decode.  enc_out bf16 [B,1500,d]; prompt_host int32[prompt_len]
 * (same prompt for every row, mlx_whisper_batch_decoder.py:406-407).  Outputs (device):
 * tokens_out int32 [B, sample_len] sampled tokens (EOT-padded), n_tokens int32[B] tokens before
 * the first EOT, sum_logprob f32[B], no_speech_prob f32[B] (softmax prob of no_speech at the
 * SOT position, unfiltered).  This call synchronises `stream` internally when polling. */
int wxb_decode_greedy(wxb_ctx* ctx, const void* enc_out_dev, int B, const int32_t* prompt_host,
                      int prompt_len, const wxb_decode_opts* opts, int32_t* tokens_out_dev,
                      int32_t* n_tokens_dev, float* sum_logprob_dev, float* no_speech_prob_dev,
                      void* stream);

/* Device-side timing of the wxb_decode_greedy calls made since the last reset (CUDA events recorded
 * on the caller's stream; this call synchronises on them): total milliseconds spent in the cross-KV
 * projection GEMMs, in the per-token step loop, and the number of decoder steps run.  Timing is OPT-IN:
 * nothing is recorded until the first call with reset = 1 (a serving process that never asks pays for
 * no events); at most 4096 calls are kept between resets. */
int wxb_decode_stats(wxb_ctx* ctx, double* cross_kv_ms, double* steps_ms, int64_t* n_steps, int reset);

/* Teacher-forced logits for parity tests: runs the decoder over tokens_host int32 [B, n_tok]
 * and writes the f32 logits of every position to logits_out_dev [B, n_tok, n_vocab]. */
int wxb_decoder_logits(wxb_ctx* ctx, const void* enc_out_dev, int B, const int32_t* tokens_host,
                       int n_tok, float* logits_out_dev, void* stream);

/* The decoder's sampling phase on its own, for parity tests of the greedy update rule on arbitrary logits
 * (mlx_whisper_batch_decoder.py:267-303 + the filters of wxb_decode_opts): reads row b of logits_dev f32 [B, ldl]
 * (ldl % 4 == 0, 16-byte aligned; static filters are applied IN PLACE), the row's last token tokens_dev[b, pos],
 * writes tokens_dev[b, pos + 1], adds the token's log-prob to sum_logprob_dev[b] unless the row had already
 * emitted EOT, sets done_dev[b] on EOT, updates ts_last_dev[b] (last sampled timestamp, -1 = none; timestamp
 * rules only) and, if opts->no_speech >= 0 and no_speech_prob_dev != NULL, the unfiltered no-speech probability. */
int wxb_decoder_sample(wxb_ctx* ctx, float* logits_dev, int64_t ldl, int B, int n_vocab, int32_t* tokens_dev,
                       int stride, int pos, int prompt_len, const wxb_decode_opts* opts, float* sum_logprob_dev,
                       int32_t* done_dev, int32_t* ts_last_dev, float* no_speech_prob_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Alignment model — replaces the per-segment `model(waveform_segment)` forward of the reference's align()
 *     (whisperx/alignment.py:240-258, one B = 1 call per segment, "TODO: batched inference") with ONE batched pass
 *     over all segments: the wav2vec2-base CTC architecture (torchaudio WAV2VEC2_ASR_BASE_960H; layer list
 *     whisperx/convert_alignment_models.py:31-70).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t conv_dim;    /* 512: channels of the 7-layer conv feature extractor (k 10,3,3,3,3,2,2; s 5,2,2,2,2,2,2; group-norm mode) */
  int32_t embed_dim;   /* 768 */
  int32_t n_heads;     /* 12 (head_dim must be 64) */
  int32_t n_layers;    /* 12, post-LN */
  int32_t ff_dim;      /* 3072 */
  int32_t pos_kernel;  /* 128 */
  int32_t pos_groups;  /* 16 */
  int32_t n_out;       /* CTC labels (29) */
} wxb_w2v_dims;

/* Borrowed weight table like wxb_set_model: "w2v.conv0.w" f32 [512,10], "w2v.gn.w/b" f32 [512], "w2v.conv{1..6}.w" bf16
 * [512, k*512] (column = tap*512 + in-channel), "w2v.fp.ln.w/b", "w2v.fp.w" bf16 [d,512], "w2v.fp.b", "w2v.pos.w" bf16
 * [d, 128*d/groups] (weight-norm folded; column = tap*(d/groups) + in-channel of the group), "w2v.pos.b",
 * "w2v.{i}.qkv.w" bf16 [3d,d] (Q|K|V), ".qkv.b", ".out.w/.b", ".ln1.w/.b", ".fc1.w/.b", ".fc2.w/.b", ".ln2.w/.b",
 * "w2v.ln.w/b", "w2v.aux.w" bf16 [n_out,d], "w2v.aux.b". */
int wxb_set_align_model(wxb_ctx* ctx, const wxb_w2v_dims* dims, const char* const* names,
                        const void* const* ptrs_dev, int n_tensors);

/* Emission frames the model produces for n_samples input samples: floor((n_samples - 400) / 320) + 1 for >= 400. */
int wxb_w2v_frames(int n_samples);

/* Batched forward.  audio_dev f32, segment b = seg_len_host[b] (>= 400) samples starting at seg_off_host[b];
 * emis_out_dev f32 [sum T_b, n_out] receives the LOGITS of every segment back to back (t_off_host int32[n_seg+1] =
 * prefix sums of wxb_w2v_frames(seg_len)); apply wxb_log_softmax_rows and hand the buffer to wxb_ctc_align. */
int wxb_w2v_emissions(wxb_ctx* ctx, const float* audio_dev, const int64_t* seg_off_host,
                      const int32_t* seg_len_host, int n_seg, float* emis_out_dev,
                      const int32_t* t_off_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * Word timing from the decoder's cross-attention — replaces the numeric core of
 *     mlx_whisper_optimized_final.py:37-125 (per-step collection of the cross-attention QK of every layer) and :128-253
 *     extract_words_with_dtw (alignment-head mean, softmax(10 x), median filter 7 (median_filter_fix.py:6-22), per-token
 *     normalisation, mlx_whisper.timing.dtw = OpenAI whisper/timing.py dtw_cpu + backtrace).  Word grouping stays on the host.
 * ---------------------------------------------------------------------------------------- */
/* Select the (layer, head) pairs (layer_head_host int32 [n_heads][2], model.alignment_heads) whose scaled cross-attention
 * queries every later wxb_decode_greedy / wxb_decoder_logits call logs on the device (64 floats per head and position);
 * n_heads = 0 switches the logging off.  At most 127 heads. */
int wxb_decode_collect_heads(wxb_ctx* ctx, const int32_t* layer_head_host, int n_heads);

/* Pre-softmax cross-attention scores of the LAST decode, averaged over the selected heads: for sequence b, rows
 * s = 0 .. n_rows_host[b]-1 are the query positions pos0 + s (pos0 = prompt_len - 1: the forward that predicts sampled token s,
 * mlx_whisper_optimized_final.py:160-176).  qk_out_dev f32 [sum n_rows, n_audio_ctx], sequences back to back.  Needs the
 * cross-K cache of that decode to be still resident (call before the next decode). */
int wxb_dtw_scores(wxb_ctx* ctx, int B, int pos0, const int32_t* n_rows_host, float* qk_out_dev, void* stream);

/* rows x T scores -> the DTW cost matrix: cost = -normalise(medfilt(softmax(temperature x), width)) per row
 * (mlx_whisper_optimized_final.py:182-199; the reference uses temperature 10, width 7).  T <= 1536, odd width <= 9. */
int wxb_dtw_cost(wxb_ctx* ctx, const float* qk_dev, int64_t rows, int T, float temperature, int medfilt_width,
                 float* cost_out_dev, void* stream);

/* dtw(x) with x[i, j] = cost[row j of sequence b][frame i] for every sequence (n_rows_host[b] <= 448 token rows of T frames,
 * back to back in cost_dev): path_frames_dev / path_tokens_dev int32 [B, wxb_dtw_path_capacity(T)] receive the path from
 * (0, 0) to (T-1, n_rows-1) (row 0 / row 1 of the reference's result), path_len_dev int32 [B] its length (0 for n_rows = 0).
 * Bit-exact against the CPU algorithm for a given cost matrix (same additions, same tie rules). */
int wxb_dtw_path(wxb_ctx* ctx, const float* cost_dev, int B, const int32_t* n_rows_host, int T, int32_t* path_frames_dev,
                 int32_t* path_tokens_dev, int32_t* path_len_dev, void* stream);
int wxb_dtw_path_capacity(int T);

/* ------------------------------------------------------------------------------------------
 * VAD post-processing and chunking — replaces whisperx/vads/pyannote.py:134-216 Binarize.__call__ (hysteresis
 *     thresholding of frame scores + the WhisperX min-cut at max_duration = chunk_size), :282-301 Pyannote.merge_chunks,
 *     whisperx/vads/vad.py:20-53 Vad.merge_chunks and the chunk slicing of whisperx/asr.py:70-73, for frame scores that are
 *     already in HBM.  Region / chunk boundaries are IEEE doubles bit-identical to the reference's Python floats.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double frame_duration; /* pyannote SlidingWindow of the score track: frame i = [start + i step, + duration), time = its middle */
  double frame_step;
  double frame_start;
  double chunk_size;     /* seconds: max_duration of the min-cut and the merge limit */
  float onset;           /* a region opens when score > onset (0 < onset < 1) */
  float offset;          /* ... and closes when score < offset; <= 0 = use onset (`offset or onset`) */
} wxb_vad_params;

/* scores_dev f32: the score tracks of n_rec recordings back to back, recording r = [score_off_host[r], score_off_host[r+1]);
 * n_samples_host int64[n_rec] = samples of each recording (chunk slices are clipped to it like a Python slice).
 * Outputs (device), per recording r:
 *   regions_dev f64 [n_rec, max_regions, 2]   speech regions (start, end) in timeline order, n_regions_dev int32[n_rec]
 *   chunks_dev  f64 [n_rec, max_chunks, 2]    merged chunks (start, end),                     n_chunks_dev  int32[n_rec]
 *   chunk_first_dev int32 [n_rec, max_chunks + 1]  chunk k holds regions chunk_first[k] .. chunk_first[k+1]-1 ("segments")
 *   chunk_off_dev int64 / chunk_len_dev int32 [n_rec, max_chunks]  int(start * 16000) and the slice length, clipped to the
 *                 recording and to 480000 samples: the table wxb_logmel / wxb_logmel_features take
 * A count larger than its capacity means the arrays were truncated (call again with larger capacities). */
int wxb_vad_chunks(wxb_ctx* ctx, const float* scores_dev, const int64_t* score_off_host, const int64_t* n_samples_host,
                   int n_rec, const wxb_vad_params* prm, int max_regions, int max_chunks, double* regions_dev,
                   int32_t* n_regions_dev, double* chunks_dev, int32_t* chunk_first_dev, int32_t* n_chunks_dev,
                   int64_t* chunk_off_dev, int32_t* chunk_len_dev, void* stream);

/* Stand-in frame scorer (NOT in the reference: Silero comes from torch.hub and the pyannote checkpoint is not in the tree, so no
 * VAD model exists offline): score[i] = sigmoid((10 log10(mean x^2 over samples [160 i, 160 i + 400)) + 1e-10) - floor_db) /
 * width_db), wxb_vad_energy_frames(n_samples) = ceil(n_samples / 160) scores; frame clock duration 0.025, step 0.010, start 0. */
int wxb_vad_energy_scores(wxb_ctx* ctx, const float* audio_dev, int64_t n_samples, float floor_db, float width_db,
                          float* scores_out_dev, void* stream);
int64_t wxb_vad_energy_frames(int64_t n_samples);

/* Stand-alone bf16 GEMM used by the encoder (exposed for parity tests and roofline timing):
 * D[M,N] = A[M,K] * W[N,K]^T (+bias[N]) (GELU) ; A,W bf16 row-major, D bf16 or f32.
 * flags: bit0 = GELU, bit1 = output f32. */
int wxb_gemm_bf16(wxb_ctx* ctx, const void* A_dev, const void* W_dev, const float* bias_dev,
                  void* D_dev, int M, int N, int K, int flags, void* stream);

/* Stand-alone encoder self-attention (exposed for parity tests and roofline timing): non-causal softmax(Q K^T / 8) V
 * per head, head_dim 64.  qkv_dev bf16 [B*T, 3d] (Q | K | V column blocks, head h at columns 64h), out_dev bf16 [B*T, d]. */
int wxb_encoder_attention(wxb_ctx* ctx, const void* qkv_dev, void* out_dev, int B, int T, int d, int H, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WXB200_H */
