// GroupNorm (+SiLU) and LayerNorm over NHWC activations, plus a row softmax.  HBM-bound kernels:
// every pixel row is read with consecutive threads on consecutive channels (coalesced); GroupNorm
// statistics go fp32 per thread -> fp32 per (CTA, 4-channel vector) -> fp64 per (image, group) in a
// fixed summation order (no atomics, bit-reproducible).
#include "common.cuh"

namespace ealdm {
namespace norm {

constexpr int NT = 256;
constexpr int MAX_GROUPS = 64;

// ---- GroupNorm statistics: part[n][chunk][v] = {sum, sum of squares} of 4 channels over the chunk ---
// No atomics: every (image, pixel chunk, 4-channel vector) partial is produced by exactly one
// thread (after a fixed-order shared-memory reduction over the CTA's pixel lanes), and the apply
// kernel combines the partials of a group in a fixed order in fp64 => bit-reproducible.
template <typename TX>
__global__ void __launch_bounds__(NT)
gn_stats_kernel(const TX* __restrict__ x, long long ld, int hw, int c, int pix_per_cta,
                float2* __restrict__ part) {
  __shared__ float2 red[NT];
  const int t = threadIdx.x;
  const int n = blockIdx.y;
  const int vpp = c >> 2;  // vec4 per pixel
  const int lanes_v = vpp < NT ? vpp : NT;
  const int pix_lanes = NT / lanes_v;
  const int tv = t % lanes_v, tp = t / lanes_v;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const TX* base = x + static_cast<long long>(n) * hw * ld;
  float2* out = part + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * vpp;
  for (int v0 = 0; v0 < vpp; v0 += lanes_v) {
    const int v = v0 + tv;
    float s = 0.f, ss = 0.f;
    if (tp < pix_lanes && v < vpp) {
      for (int pix = p0 + tp; pix < p1; pix += pix_lanes) {
        Vec4<TX> q;
        q.load(base + static_cast<long long>(pix) * ld + v * 4);
        float f[4];
        q.get(f);
#pragma unroll
        for (int j = 0; j < 4; ++j) { s += f[j]; ss = fmaf(f[j], f[j], ss); }
      }
    }
    if (pix_lanes > 1) {
      __syncthreads();
      red[t] = make_float2(s, ss);
      __syncthreads();
      if (tp == 0 && v < vpp) {
        for (int k = 1; k < pix_lanes; ++k) { s += red[k * lanes_v + tv].x; ss += red[k * lanes_v + tv].y; }
      }
    }
    if (tp == 0 && v < vpp) out[v] = make_float2(s, ss);
  }
}

// ---- GroupNorm apply: y = (x - mean) * rstd * gamma + beta, optional SiLU ---------------------------
template <typename TX, typename T>
__global__ void __launch_bounds__(NT)
gn_apply_kernel(const TX* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, int hw,
                int c, int groups, int pix_per_cta, const float2* __restrict__ part, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, int act) {
  __shared__ float s_mean[MAX_GROUPS], s_rstd[MAX_GROUPS];
  const int t = threadIdx.x;
  const int n = blockIdx.y;
  const int cpg = c / groups;
  const int vpp = c >> 2;
  {
    // 8 lanes per group sum the group's partials (chunks x cpg/4 vectors) in a fixed order
    const int chunks = gridDim.x;
    const int vpg = cpg >> 2;
    const int terms = chunks * vpg;
    const float2* pn = part + static_cast<long long>(n) * chunks * vpp;
    for (int g0 = 0; g0 < groups; g0 += NT / 8) {
      const int g = g0 + (t >> 3);
      const int sub = t & 7;
      double s = 0.0, ss = 0.0;
      if (g < groups) {
        for (int k = sub; k < terms; k += 8) {
          const int ch = k / vpg, vv = k - ch * vpg;
          const float2 p = pn[static_cast<long long>(ch) * vpp + g * vpg + vv];
          s += static_cast<double>(p.x);
          ss += static_cast<double>(p.y);
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
      }
      if (g < groups && sub == 0) {
        const double cnt = static_cast<double>(hw) * cpg;
        const double mean = s / cnt;
        double var = ss / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[g] = static_cast<float>(mean);
        s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      }
    }
  }
  __syncthreads();
  // same thread -> (4-channel vector, pixel lane) mapping as the statistics kernel: consecutive threads
  // on consecutive channels (coalesced), no integer divisions in the pixel loop
  const int lanes_v = vpp < NT ? vpp : NT;
  const int pix_lanes = NT / lanes_v;
  const int tv = t % lanes_v, tp = t / lanes_v;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const TX* xb = x + static_cast<long long>(n) * hw * ld_x;
  T* yb = y + static_cast<long long>(n) * hw * ld_y;
  if (tp < pix_lanes) {
    for (int v = tv; v < vpp; v += lanes_v) {
      const int ch = v * 4;
      const int g = ch / cpg;
      const float mean = s_mean[g], rstd = s_rstd[g];
      const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + ch));
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + ch));
      // y = x * a + b with a = rstd * gamma, b = beta - mean * a
      const float a0 = rstd * ga.x, a1 = rstd * ga.y, a2 = rstd * ga.z, a3 = rstd * ga.w;
      const float b0 = be.x - mean * a0, b1 = be.y - mean * a1, b2 = be.z - mean * a2, b3 = be.w - mean * a3;
#pragma unroll 4
      for (int pix = p0 + tp; pix < p1; pix += pix_lanes) {
        Vec4<TX> qx;
        qx.load(xb + static_cast<long long>(pix) * ld_x + ch);
        float f[4];
        qx.get(f);
        f[0] = fmaf(f[0], a0, b0); f[1] = fmaf(f[1], a1, b1); f[2] = fmaf(f[2], a2, b2); f[3] = fmaf(f[3], a3, b3);
        if (act == EALDM_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) f[j] = silu_f(f[j]);
        }
        Vec4<T> q;
        q.set(f);
        q.store(yb + static_cast<long long>(pix) * ld_y + ch);
      }
    }
  }
}

// pixel chunking shared by the workspace query and the launch
static void gn_chunking(long long n, long long hw, int* pix_per_cta, long long* chunks) {
  // enough CTAs to fill 148 SMs a few times over, at least 16 pixels each
  long long ch = ceil_div(1184, n);
  const long long max_chunks = ceil_div(hw, 16);
  if (ch > max_chunks) ch = max_chunks;
  if (ch < 1) ch = 1;
  const int ppc = static_cast<int>(ceil_div(hw, ch));
  *pix_per_cta = ppc;
  *chunks = ceil_div(hw, ppc);
}

template <typename TX, typename T>
static int group_norm_t(const ealdm_group_norm_args* a, cudaStream_t st) {
  const int hw = static_cast<int>(a->hw);
  const int n = static_cast<int>(a->n);
  int ppc;
  long long chunks;
  gn_chunking(n, hw, &ppc, &chunks);
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(n));
  float2* part = reinterpret_cast<float2*>(a->workspace);
  gn_stats_kernel<TX><<<grid, NT, 0, st>>>(reinterpret_cast<const TX*>(a->x), a->ld_x, hw,
                                           static_cast<int>(a->c), ppc, part);
  EALDM_LAUNCH_CHECK();
  gn_apply_kernel<TX, T><<<grid, NT, 0, st>>>(reinterpret_cast<const TX*>(a->x), a->ld_x,
                                          reinterpret_cast<T*>(a->y), a->ld_y, hw,
                                          static_cast<int>(a->c), a->groups, ppc, part, a->eps,
                                          a->gamma, a->beta, a->act);
  EALDM_LAUNCH_CHECK();
  return 0;
}

// ---- LayerNorm: one warp per row; the row lives in registers (one global read), two-pass variance ----
template <typename TX, typename T, int NV>   // NV = float4 per lane: c == NV * 128
__global__ void __launch_bounds__(NT)
layer_norm_reg_kernel(const TX* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y,
                      long long rows, float eps, const float* __restrict__ gamma,
                      const float* __restrict__ beta) {
  constexpr int C_ = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (NT / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TX* xr = x + row * ld_x;
  T* yr = y + row * ld_y;
  float f[NV][4];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    Vec4<TX> q;
    q.load(xr + (k * 32 + lane) * 4);
    q.get(f[k]);
    s += (f[k][0] + f[k][1]) + (f[k][2] + f[k][3]);
  }
  const float mean = warp_sum(s) / static_cast<float>(C_);
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float d = f[k][j] - mean; ss = fmaf(d, d, ss); }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(C_) + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * 4;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    float o[4];
    o[0] = (f[k][0] - mean) * rstd * ga.x + be.x;
    o[1] = (f[k][1] - mean) * rstd * ga.y + be.y;
    o[2] = (f[k][2] - mean) * rstd * ga.z + be.z;
    o[3] = (f[k][3] - mean) * rstd * ga.w + be.w;
    Vec4<T> q;
    q.set(o);
    q.store(yr + c);
  }
}

// generic fallback (any c % 4 == 0): three passes over the row, served from L1 after the first
template <typename TX, typename T>
__global__ void __launch_bounds__(NT)
layer_norm_kernel(const TX* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y,
                  long long rows, int c, float eps, const float* __restrict__ gamma,
                  const float* __restrict__ beta) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (NT / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TX* xr = x + row * ld_x;
  T* yr = y + row * ld_y;
  const int vpr = c >> 2;
  float s = 0.f;
  for (int v = lane; v < vpr; v += 32) {
    Vec4<TX> q;
    q.load(xr + v * 4);
    float f[4];
    q.get(f);
    s += (f[0] + f[1]) + (f[2] + f[3]);
  }
  const float mean = warp_sum(s) / static_cast<float>(c);
  float ss = 0.f;
  for (int v = lane; v < vpr; v += 32) {
    Vec4<TX> q;
    q.load(xr + v * 4);
    float f[4];
    q.get(f);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float d = f[j] - mean; ss = fmaf(d, d, ss); }
  }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(c) + eps);
  for (int v = lane; v < vpr; v += 32) {
    Vec4<TX> qx;
    qx.load(xr + v * 4);
    float f[4];
    qx.get(f);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + v * 4));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + v * 4));
    f[0] = (f[0] - mean) * rstd * ga.x + be.x;
    f[1] = (f[1] - mean) * rstd * ga.y + be.y;
    f[2] = (f[2] - mean) * rstd * ga.z + be.z;
    f[3] = (f[3] - mean) * rstd * ga.w + be.w;
    Vec4<T> q;
    q.set(f);
    q.store(yr + v * 4);
  }
}

template <typename TX, typename T>
static void layer_norm_launch(const ealdm_layer_norm_args* a, cudaStream_t st) {
  const unsigned grid = static_cast<unsigned>(ceil_div(a->rows, NT / 32));
  const TX* x = reinterpret_cast<const TX*>(a->x);
  T* y = reinterpret_cast<T*>(a->y);
#define EALDM_LN_REG(NV)                                                                         \
  layer_norm_reg_kernel<TX, T, NV><<<grid, NT, 0, st>>>(x, a->ld_x, y, a->ld_y, a->rows, a->eps, \
                                                        a->gamma, a->beta)
  switch (a->c) {
    case 128: EALDM_LN_REG(1); break;
    case 256: EALDM_LN_REG(2); break;
    case 512: EALDM_LN_REG(4); break;
    case 1024: EALDM_LN_REG(8); break;
    default:
      layer_norm_kernel<TX, T><<<grid, NT, 0, st>>>(x, a->ld_x, y, a->ld_y, a->rows, (int)a->c, a->eps,
                                                    a->gamma, a->beta);
  }
#undef EALDM_LN_REG
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

extern "C" int64_t ealdm_group_norm_workspace_bytes(int64_t n, int64_t hw, int64_t c) {
  if (n <= 0 || hw <= 0 || c <= 0) return 0;
  int ppc;
  long long chunks;
  norm::gn_chunking(n, hw, &ppc, &chunks);
  return n * chunks * (c / 4) * static_cast<int64_t>(sizeof(float2));
}

extern "C" int ealdm_group_norm(const ealdm_group_norm_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->x && a->y && a->workspace && a->gamma && a->beta, "group_norm: null argument");
  EALDM_REQUIRE(a->groups > 0 && a->groups <= norm::MAX_GROUPS && a->c % a->groups == 0,
                "group_norm: c=%lld not divisible by groups=%d", (long long)a->c, a->groups);
  EALDM_REQUIRE((a->c / a->groups) % 4 == 0 && a->ld_x % 4 == 0 && a->ld_y % 4 == 0,
                "group_norm: channels per group, ld_x and ld_y must be multiples of 4");
  EALDM_REQUIRE(a->n > 0 && a->n <= 65535 && a->hw > 0, "group_norm: bad n/hw");
  EALDM_REQUIRE(a->act == EALDM_ACT_NONE || a->act == EALDM_ACT_SILU, "group_norm: bad act");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->dtype == EALDM_F32) return norm::group_norm_t<float, float>(a, st);
  if (a->dtype == EALDM_BF16 && a->x_f32) return norm::group_norm_t<float, bf16>(a, st);
  if (a->dtype == EALDM_BF16) return norm::group_norm_t<bf16, bf16>(a, st);
  return set_error(EALDM_EINVAL, "group_norm: bad dtype %d", a->dtype);
}

extern "C" int ealdm_layer_norm(const ealdm_layer_norm_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->x && a->y && a->gamma && a->beta, "layer_norm: null argument");
  EALDM_REQUIRE(a->c % 4 == 0 && a->ld_x % 4 == 0 && a->ld_y % 4 == 0,
                "layer_norm: c, ld_x, ld_y must be multiples of 4");
  if (a->rows == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->dtype == EALDM_F32) {
    norm::layer_norm_launch<float, float>(a, st);
  } else if (a->dtype == EALDM_BF16 && a->x_f32) {
    norm::layer_norm_launch<float, bf16>(a, st);
  } else if (a->dtype == EALDM_BF16) {
    norm::layer_norm_launch<bf16, bf16>(a, st);
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
