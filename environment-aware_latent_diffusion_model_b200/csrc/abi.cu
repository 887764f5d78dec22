// Library-level entry points of the C ABI and the conv/linear dispatcher.
#include "common.cuh"

#include <atomic>
#include <stdlib.h>

namespace ealdm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// programmatic dependent launch between the kernels of one stream (common.cuh); EALDM_PDL=1 / ealdm_set_pdl(1) enable it
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("EALDM_PDL");
    v = (e && atoi(e) != 0) ? 1 : 0;   // opt-in: measured neutral on the graph-replayed forward
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

namespace tc {
bool supported(const ealdm_conv_args* a);
int launch(const ealdm_conv_args* a, cudaStream_t st);
int set_option(int option, int value);
int ln_parts(const ealdm_conv_args* a);
}  // namespace tc
namespace simt {
int launch(const ealdm_conv_args* a, cudaStream_t st);
}

}  // namespace ealdm

using namespace ealdm;

extern "C" int ealdm_abi_version(void) { return EALDM_ABI_VERSION; }
extern "C" const char* ealdm_last_error(void) { return g_err; }
extern "C" int64_t ealdm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int ealdm_tc_set_option(int option, int value) { return tc::set_option(option, value); }
extern "C" int ealdm_set_pdl(int enabled) {
  const int prev = pdl_enabled() ? 1 : 0;
  g_pdl.store(enabled ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

extern "C" int ealdm_device_check(void) {
  int dev = 0;
  EALDM_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  EALDM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  EALDM_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return set_error(EALDM_EUNSUPPORTED,
                     "device %d is sm_%d%d; libealdm_b200 contains sm_100a code only", dev, major,
                     minor);
  return 0;
}

extern "C" int64_t ealdm_conv_ln_parts(const ealdm_conv_args* a) {
  if (a == nullptr) return set_error(EALDM_EINVAL, "conv_ln_parts: null args");
  return tc::ln_parts(a);
}

extern "C" int ealdm_conv(const ealdm_conv_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a != nullptr, "conv: null args");
  EALDM_REQUIRE(a->dtype == EALDM_F32 || a->dtype == EALDM_BF16, "conv: bad dtype %d", a->dtype);
  EALDM_REQUIRE(a->n_src == 1 || a->n_src == 2, "conv: n_src must be 1 or 2");
  EALDM_REQUIRE(a->weight && a->out, "conv: null weight/out");
  EALDM_REQUIRE(a->n_out > 0 && a->k_total > 0 && a->h_out > 0 && a->w_out > 0, "conv: bad sizes");
  EALDM_REQUIRE(a->act >= EALDM_ACT_NONE && a->act <= EALDM_ACT_SOFTMAX4, "conv: bad act %d", a->act);
  EALDM_REQUIRE(a->act != EALDM_ACT_SOFTMAX4 || a->wi_tokens == 4, "conv: SOFTMAX4 needs per-image weights, 4 tokens");
  EALDM_REQUIRE(!(a->act == EALDM_ACT_GEGLU && (a->rowvec || a->out2)),
                "conv: GEGLU with rowvec / out2 unsupported");
  for (int s = 0; s < a->n_src; ++s) {
    const ealdm_conv_src& x = a->src[s];
    EALDM_REQUIRE(x.x != nullptr, "conv: src[%d].x is null", s);
    EALDM_REQUIRE(x.n > 0 && x.h > 0 && x.w > 0 && x.c > 0 && x.ld >= x.c, "conv: src[%d] bad dims", s);
    EALDM_REQUIRE(x.ksize == 1 || x.ksize == 3, "conv: src[%d].ksize must be 1 or 3", s);
    EALDM_REQUIRE(x.stride == 1 || x.stride == 2, "conv: src[%d].stride must be 1 or 2", s);
    EALDM_REQUIRE(x.pad >= 0 && x.pad <= 1, "conv: src[%d].pad must be 0 or 1", s);
    EALDM_REQUIRE(x.n == a->src[0].n, "conv: sources disagree on n");
    // the output extent must be reachable: last tap of the last output pixel may hang over by <= 1
    const long long hin = x.upsample ? x.h * 2 : x.h, win = x.upsample ? x.w * 2 : x.w;
    EALDM_REQUIRE((a->h_out - 1) * x.stride - x.pad < hin && (a->w_out - 1) * x.stride - x.pad < win,
                  "conv: src[%d] output extent %lldx%lld outside the input", s, (long long)a->h_out,
                  (long long)a->w_out);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->weight_adjoint && !a->wi_tokens) {
    EALDM_REQUIRE(a->impl != EALDM_IMPL_SIMT && a->dtype == EALDM_BF16 && tc::supported(a),
                  "conv: weight_adjoint needs the tcgen05 path (bf16, one source, c %% 64 == 0, n_out %% 64 == 0)");
    return tc::launch(a, st);
  }
  if (a->wi_tokens) {
    EALDM_REQUIRE(a->impl != EALDM_IMPL_SIMT && a->dtype == EALDM_BF16 && tc::supported(a),
                  "conv: per-image weights (wi_*) need the tcgen05 path: bf16, one 1x1 source, h*w %% 128 == 0, "
                  "wi_heads * wi_tokens in {32, 64, 96, 128}");
    return tc::launch(a, st);
  }
  EALDM_REQUIRE(!(a->ln_partial_out || a->ln_partial_in) ||
                    (a->impl != EALDM_IMPL_SIMT && a->dtype == EALDM_BF16 && tc::supported(a)),
                "conv: a folded LayerNorm (ln_partial_*) needs the tcgen05 path in linear geometry");
  if (a->impl == EALDM_IMPL_SIMT) return simt::launch(a, st);
  if (a->impl == EALDM_IMPL_TCGEN05) return tc::launch(a, st);
  if (a->dtype == EALDM_BF16 && tc::supported(a)) return tc::launch(a, st);
  return simt::launch(a, st);
}
