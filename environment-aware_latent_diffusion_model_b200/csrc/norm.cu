// GroupNorm (+SiLU) and LayerNorm over NHWC activations, plus a row softmax.  HBM-bound kernels:
// every pixel row is read with consecutive threads on consecutive channels (coalesced), statistics
// are accumulated fp32 per thread -> fp32 per CTA (shared atomics) -> fp64 per (image, group).
#include "common.cuh"

namespace ealdm {
namespace norm {

constexpr int NT = 256;
constexpr int MAX_GROUPS = 64;

// ---- GroupNorm statistics: stats[n][g] = {sum, sum of squares} ------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT)
gn_stats_kernel(const T* __restrict__ x, long long ld, int hw, int c, int groups, int pix_per_cta,
                double* __restrict__ stats) {
  __shared__ float s_sum[MAX_GROUPS], s_sq[MAX_GROUPS];
  const int t = threadIdx.x;
  const int n = blockIdx.y;
  if (t < groups) { s_sum[t] = 0.f; s_sq[t] = 0.f; }
  __syncthreads();
  const int vpp = c >> 2;  // vec4 per pixel
  const int cpg = c / groups;
  const int lanes_v = vpp < NT ? vpp : NT;
  const int pix_lanes = NT / lanes_v;
  const int tv = t % lanes_v, tp = t / lanes_v;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const T* base = x + static_cast<long long>(n) * hw * ld;
  if (tp < pix_lanes) {
    for (int v = tv; v < vpp; v += lanes_v) {
      float s = 0.f, ss = 0.f;
      for (int pix = p0 + tp; pix < p1; pix += pix_lanes) {
        Vec4<T> q;
        q.load(base + static_cast<long long>(pix) * ld + v * 4);
        float f[4];
        q.get(f);
#pragma unroll
        for (int j = 0; j < 4; ++j) { s += f[j]; ss = fmaf(f[j], f[j], ss); }
      }
      const int g = (v * 4) / cpg;
      atomicAdd(&s_sum[g], s);
      atomicAdd(&s_sq[g], ss);
    }
  }
  __syncthreads();
  if (t < groups) {
    double* d = stats + (static_cast<long long>(n) * groups + t) * 2;
    atomicAdd(d, static_cast<double>(s_sum[t]));
    atomicAdd(d + 1, static_cast<double>(s_sq[t]));
  }
}

// ---- GroupNorm apply: y = (x - mean) * rstd * gamma + beta, optional SiLU ---------------------------
template <typename T>
__global__ void __launch_bounds__(NT)
gn_apply_kernel(const T* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, int hw,
                int c, int groups, int pix_per_cta, const double* __restrict__ stats, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, int act) {
  __shared__ float s_mean[MAX_GROUPS], s_rstd[MAX_GROUPS];
  const int t = threadIdx.x;
  const int n = blockIdx.y;
  const int cpg = c / groups;
  if (t < groups) {
    const double* d = stats + (static_cast<long long>(n) * groups + t) * 2;
    const double cnt = static_cast<double>(hw) * cpg;
    const double mean = d[0] / cnt;
    double var = d[1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[t] = static_cast<float>(mean);
    s_rstd[t] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const int vpp = c >> 2;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const long long items = static_cast<long long>(p1 - p0) * vpp;
  const T* xb = x + static_cast<long long>(n) * hw * ld_x;
  T* yb = y + static_cast<long long>(n) * hw * ld_y;
  for (long long i = t; i < items; i += NT) {
    const int pix = p0 + static_cast<int>(i / vpp);
    const int v = static_cast<int>(i % vpp);
    const int ch = v * 4;
    Vec4<T> q;
    q.load(xb + static_cast<long long>(pix) * ld_x + ch);
    float f[4];
    q.get(f);
    const int g = ch / cpg;
    const float mean = s_mean[g], rstd = s_rstd[g];
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + ch));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + ch));
    const float gv[4] = {ga.x, ga.y, ga.z, ga.w};
    const float bv[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float o = (f[j] - mean) * rstd * gv[j] + bv[j];
      if (act == EALDM_ACT_SILU) o = silu_f(o);
      f[j] = o;
    }
    q.set(f);
    q.store(yb + static_cast<long long>(pix) * ld_y + ch);
  }
}

template <typename T>
static int group_norm_t(const ealdm_group_norm_args* a, cudaStream_t st) {
  const int hw = static_cast<int>(a->hw);
  const int n = static_cast<int>(a->n);
  // enough CTAs to fill 148 SMs a few times over, at least 16 pixels each
  long long chunks = ceil_div(1184, n);
  const long long max_chunks = ceil_div(hw, 16);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int ppc = static_cast<int>(ceil_div(hw, chunks));
  chunks = ceil_div(hw, ppc);
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(n));
  EALDM_CUDA(cudaMemsetAsync(a->stats, 0, sizeof(double) * 2 * n * a->groups, st));
  gn_stats_kernel<T><<<grid, NT, 0, st>>>(reinterpret_cast<const T*>(a->x), a->ld_x, hw,
                                          static_cast<int>(a->c), a->groups, ppc, a->stats);
  EALDM_LAUNCH_CHECK();
  gn_apply_kernel<T><<<grid, NT, 0, st>>>(reinterpret_cast<const T*>(a->x), a->ld_x,
                                          reinterpret_cast<T*>(a->y), a->ld_y, hw,
                                          static_cast<int>(a->c), a->groups, ppc, a->stats, a->eps,
                                          a->gamma, a->beta, a->act);
  EALDM_LAUNCH_CHECK();
  return 0;
}

// ---- LayerNorm: one warp per row, two-pass mean / centred variance ------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT)
layer_norm_kernel(const T* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y,
                  long long rows, int c, float eps, const float* __restrict__ gamma,
                  const float* __restrict__ beta) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (NT / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * ld_x;
  T* yr = y + row * ld_y;
  const int vpr = c >> 2;
  float s = 0.f;
  for (int v = lane; v < vpr; v += 32) {
    Vec4<T> q;
    q.load(xr + v * 4);
    float f[4];
    q.get(f);
    s += (f[0] + f[1]) + (f[2] + f[3]);
  }
  const float mean = warp_sum(s) / static_cast<float>(c);
  float ss = 0.f;
  for (int v = lane; v < vpr; v += 32) {
    Vec4<T> q;
    q.load(xr + v * 4);
    float f[4];
    q.get(f);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float d = f[j] - mean; ss = fmaf(d, d, ss); }
  }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(c) + eps);
  for (int v = lane; v < vpr; v += 32) {
    Vec4<T> q;
    q.load(xr + v * 4);
    float f[4];
    q.get(f);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + v * 4));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + v * 4));
    f[0] = (f[0] - mean) * rstd * ga.x + be.x;
    f[1] = (f[1] - mean) * rstd * ga.y + be.y;
    f[2] = (f[2] - mean) * rstd * ga.z + be.z;
    f[3] = (f[3] - mean) * rstd * ga.w + be.w;
    q.set(f);
    q.store(yr + v * 4);
  }
}

// ---- in-place row softmax(scale * x), one warp per row ---------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT)
softmax_rows_kernel(T* __restrict__ x, long long ld, long long rows, int c, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (NT / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  T* xr = x + row * ld;
  float m = -INFINITY;
  for (int i = lane; i < c; i += 32) m = fmaxf(m, to_f32(xr[i]) * scale);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < c; i += 32) s += expf(to_f32(xr[i]) * scale - m);
  s = warp_sum(s);
  const float inv = 1.0f / s;
  for (int i = lane; i < c; i += 32) xr[i] = from_f32<T>(expf(to_f32(xr[i]) * scale - m) * inv);
}

}  // namespace norm
}  // namespace ealdm

using namespace ealdm;

extern "C" int ealdm_group_norm(const ealdm_group_norm_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->x && a->y && a->stats && a->gamma && a->beta, "group_norm: null argument");
  EALDM_REQUIRE(a->groups > 0 && a->groups <= norm::MAX_GROUPS && a->c % a->groups == 0,
                "group_norm: c=%lld not divisible by groups=%d", (long long)a->c, a->groups);
  EALDM_REQUIRE((a->c / a->groups) % 4 == 0 && a->ld_x % 4 == 0 && a->ld_y % 4 == 0,
                "group_norm: channels per group, ld_x and ld_y must be multiples of 4");
  EALDM_REQUIRE(a->n > 0 && a->n <= 65535 && a->hw > 0, "group_norm: bad n/hw");
  EALDM_REQUIRE(a->act == EALDM_ACT_NONE || a->act == EALDM_ACT_SILU, "group_norm: bad act");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->dtype == EALDM_F32) return norm::group_norm_t<float>(a, st);
  if (a->dtype == EALDM_BF16) return norm::group_norm_t<bf16>(a, st);
  return set_error(EALDM_EINVAL, "group_norm: bad dtype %d", a->dtype);
}

extern "C" int ealdm_layer_norm(const ealdm_layer_norm_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->x && a->y && a->gamma && a->beta, "layer_norm: null argument");
  EALDM_REQUIRE(a->c % 4 == 0 && a->ld_x % 4 == 0 && a->ld_y % 4 == 0,
                "layer_norm: c, ld_x, ld_y must be multiples of 4");
  if (a->rows == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(ceil_div(a->rows, norm::NT / 32));
  if (a->dtype == EALDM_F32) {
    norm::layer_norm_kernel<float><<<grid, norm::NT, 0, st>>>(
        reinterpret_cast<const float*>(a->x), a->ld_x, reinterpret_cast<float*>(a->y), a->ld_y,
        a->rows, (int)a->c, a->eps, a->gamma, a->beta);
  } else if (a->dtype == EALDM_BF16) {
    norm::layer_norm_kernel<bf16><<<grid, norm::NT, 0, st>>>(
        reinterpret_cast<const bf16*>(a->x), a->ld_x, reinterpret_cast<bf16*>(a->y), a->ld_y,
        a->rows, (int)a->c, a->eps, a->gamma, a->beta);
  } else {
    return set_error(EALDM_EINVAL, "layer_norm: bad dtype %d", a->dtype);
  }
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_softmax_rows(void* x, int64_t ld, int32_t dtype, int64_t rows, int64_t c,
                                  float scale, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && rows >= 0 && c > 0, "softmax_rows: bad argument");
  if (rows == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, norm::NT / 32));
  if (dtype == EALDM_F32)
    norm::softmax_rows_kernel<float><<<grid, norm::NT, 0, st>>>(reinterpret_cast<float*>(x), ld, rows,
                                                                (int)c, scale);
  else if (dtype == EALDM_BF16)
    norm::softmax_rows_kernel<bf16><<<grid, norm::NT, 0, st>>>(reinterpret_cast<bf16*>(x), ld, rows,
                                                               (int)c, scale);
  else
    return set_error(EALDM_EINVAL, "softmax_rows: bad dtype %d", dtype);
  EALDM_LAUNCH_CHECK();
  return 0;
}
