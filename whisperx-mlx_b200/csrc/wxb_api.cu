// wxb_api.cu — context lifetime, error slot, workspace management (C-ABI in include/wxb200.h).
#include "wxb_common.cuh"
#include <stdarg.h>
#include <string.h>

static std::string g_global_err;

int wxb_fail(wxb_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->err = buf;
  else
    g_global_err = buf;
  return code;
}

int wxb_reserve(wxb_ctx* ctx, wxb_buf& b, size_t bytes) {
  if (bytes <= b.cap) return WXB_OK;
  if (b.p) {
    // buffers may still be in use by queued work: a blocking free is the safe choice here
    cudaError_t e = cudaFree(b.p);
    if (e != cudaSuccess) return wxb_fail(ctx, WXB_ERR_CUDA, "cudaFree: %s", cudaGetErrorString(e));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess)
    return wxb_fail(ctx, WXB_ERR_CUDA, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
  b.cap = want;
  return WXB_OK;
}

void* wxb_named(wxb_ctx* ctx, const char* name, size_t bytes, bool zero_on_alloc) {
  wxb_buf& b = ctx->named[name];
  if (bytes <= b.cap) return b.p;
  size_t old = b.cap;
  if (wxb_reserve(ctx, b, bytes) != WXB_OK) return nullptr;
  if (zero_on_alloc && b.cap != old) cudaMemset(b.p, 0, b.cap);
  return b.p;
}

void wxb_dec_timings_clear(wxb_ctx* ctx) {
  for (auto& t : ctx->dec_timings) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); cudaEventDestroy(t.e2); }
  ctx->dec_timings.clear();
}

extern "C" {

int wxb_abi_version(void) { return WXB_ABI_VERSION; }

int wxb_create(int device, wxb_ctx** out) {
  if (!out) return wxb_fail(nullptr, WXB_ERR_INVALID, "wxb_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return wxb_fail(nullptr, WXB_ERR_CUDA, "wxb_create: no CUDA device (%s); there is no CPU fallback",
                    e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= n)
    return wxb_fail(nullptr, WXB_ERR_INVALID, "wxb_create: device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess)
    return wxb_fail(nullptr, WXB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return wxb_fail(nullptr, WXB_ERR_UNSUPPORTED,
                    "wxb_create: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return wxb_fail(nullptr, WXB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  wxb_ctx* c = new wxb_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return WXB_OK;
}

void wxb_destroy(wxb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  wxb_model_free(ctx);
  wxb_align_model_free(ctx);
  wxb_dec_timings_clear(ctx);
  for (int i = 0; i < WXB_MAX_DEC_GROUPS - 1; ++i) {
    if (ctx->dec_side[i]) cudaStreamDestroy(ctx->dec_side[i]);
    if (ctx->dec_join[i]) cudaEventDestroy(ctx->dec_join[i]);
  }
  if (ctx->dec_fork) cudaEventDestroy(ctx->dec_fork);
  wxb_buf* bufs[] = {&ctx->ws_ctc_trellis, &ctx->ws_ctc_hist, &ctx->ws_ctc_meta, &ctx->ws_mel_max,
                     &ctx->ws_mel_band};
  for (wxb_buf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (auto& kv : ctx->named)
    if (kv.second.p) cudaFree(kv.second.p);
  delete ctx;
}

const char* wxb_last_error(const wxb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_global_err.c_str(); }

int64_t wxb_launch_count(const wxb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int wxb_debug_set(wxb_ctx* ctx, const char* key, int value) {
  if (!ctx || !key) return WXB_ERR_INVALID;
  if (std::string(key) == "w2v_stop") { ctx->w2v_stop = value; return WXB_OK; }
  if (std::string(key) == "enc_group") {  // chunks per group of the encoder's layer stack (0 = whole batch, -1 = the library's default); results do not depend on it
    if (value < -1) return wxb_fail(ctx, WXB_ERR_INVALID, "enc_group: >= -1");
    ctx->enc_group = value;
    return WXB_OK;
  }
  if (std::string(key) == "dec_groups") {  // A/B aid: sequence groups per decode call (0 = automatic); results do not depend on it
    if (value < 0 || value > WXB_MAX_DEC_GROUPS) return wxb_fail(ctx, WXB_ERR_INVALID, "dec_groups: 0 .. %d", WXB_MAX_DEC_GROUPS);
    ctx->dec_groups_override = value;
    return WXB_OK;
  }
  if (std::string(key) == "dec_group_delay_ns") {  // A/B aid: start offset between the sequence groups' kernel instances
    if (value < 0 || value > 10000000) return wxb_fail(ctx, WXB_ERR_INVALID, "dec_group_delay_ns: 0 .. 10000000");
    ctx->dec_group_delay_ns = value;
    return WXB_OK;
  }
  return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_debug_set: unknown key '%s'", key);
}

int wxb_debug_copy(wxb_ctx* ctx, const char* name, void* dst_dev, int64_t offset, int64_t bytes) {
  if (!ctx || !name || !dst_dev || bytes < 0 || offset < 0) return WXB_ERR_INVALID;
  auto it = ctx->named.find(name);
  if (it == ctx->named.end() || !it->second.p) return wxb_fail(ctx, WXB_ERR_STATE, "wxb_debug_copy: no workspace named '%s'", name);
  if ((size_t)(offset + bytes) > it->second.cap) return wxb_fail(ctx, WXB_ERR_INVALID, "wxb_debug_copy: '%s' holds %zu bytes", name, it->second.cap);
  WXB_CUDA(ctx, cudaSetDevice(ctx->device));
  WXB_CUDA(ctx, cudaDeviceSynchronize());
  WXB_CUDA(ctx, cudaMemcpy(dst_dev, (const char*)it->second.p + offset, (size_t)bytes, cudaMemcpyDeviceToDevice));
  return WXB_OK;
}

}  // extern "C"
