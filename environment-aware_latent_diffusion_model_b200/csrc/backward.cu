// Backward kernels of the HBM-bound operators: GroupNorm(+SiLU), LayerNorm, GEGLU, SiLU, column sums
// (bias / per-image embedding gradients), and the two resampling adjoints (zero insertion for the
// stride-2 convolution's data gradient, 2x2 sum pooling for the nearest-2x upsampling).
// All reductions are done in a fixed order (thread -> CTA partial -> final pass): no atomics, the
// gradients are bit-reproducible run to run.  Parameter gradients ACCUMULATE into their destination
// (PyTorch .grad semantics).
#include "common.cuh"

#include <stdlib.h>

namespace ealdm {
namespace bwd {

constexpr int NT = 256;
constexpr int MAX_GROUPS = 64;

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }
// d/dz silu(z)
__device__ __forceinline__ float dsilu_f(float z) {
  const float s = sigmoid_f(z);
  return s * fmaf(z, 1.0f - s, 1.0f);
}
// d/dx gelu_erf(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float dgelu_erf_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return fmaf(x, pdf, cdf);
}

static void chunking(long long n, long long hw, int* pix_per_cta, long long* chunks) {
  long long ch = ceil_div(2368, n);
  const long long max_chunks = ceil_div(hw, 16);
  if (ch > max_chunks) ch = max_chunks;
  if (ch < 1) ch = 1;
  const int ppc = static_cast<int>(ceil_div(hw, ch));
  *pix_per_cta = ppc;
  *chunks = ceil_div(hw, ppc);
}

// ---- GroupNorm backward ---------------------------------------------------------------------------------
// y = act(z), z = gamma * xhat + beta, xhat = (x - mean) * rstd over (c/groups, hw) of one image.
//   dz    = dy * act'(z)
//   dx    = rstd * (gamma*dz - (s1 + xhat*s2)/m),  s1 = sum_g gamma*dz, s2 = sum_g gamma*dz*xhat, m = cpg*hw
//   dgamma[c] += sum_{n,p} dz*xhat,  dbeta[c] += sum_{n,p} dz
template <typename TX, typename TD>
__global__ void __launch_bounds__(NT)
gn_bwd_stats_kernel(const TX* __restrict__ x, long long ld_x, const TD* __restrict__ dy, long long ld_dy, int hw,
                    int c, int groups, int pix_per_cta, const float* __restrict__ stats,
                    const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                    float2* __restrict__ part) {
  __shared__ float red[NT][8];
  const int t = threadIdx.x;
  const int n = blockIdx.y;
  const int vpp = c >> 2;
  const int cpg = c / groups;
  const int lanes_v = vpp < NT ? vpp : NT;
  const int pix_lanes = NT / lanes_v;
  const int tv = t % lanes_v, tp = t / lanes_v;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const TX* xb = x + static_cast<long long>(n) * hw * ld_x;
  const TD* db = dy + static_cast<long long>(n) * hw * ld_dy;
  float2* out = part + (static_cast<long long>(n) * gridDim.x + blockIdx.x) * c;
  for (int v0 = 0; v0 < vpp; v0 += lanes_v) {
    const int v = v0 + tv;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (tp < pix_lanes && v < vpp) {
      const int ch = v * 4;
      const int g = ch / cpg;
      const float mean = stats[(static_cast<long long>(n) * groups + g) * 2];
      const float rstd = stats[(static_cast<long long>(n) * groups + g) * 2 + 1];
      const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + ch));
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + ch));
      const float gv[4] = {ga.x, ga.y, ga.z, ga.w}, bv[4] = {be.x, be.y, be.z, be.w};
      for (int pix = p0 + tp; pix < p1; pix += pix_lanes) {
        Vec4<TX> qx;
        Vec4<TD> qd;
        qx.load(xb + static_cast<long long>(pix) * ld_x + ch);
        qd.load(db + static_cast<long long>(pix) * ld_dy + ch);
        float fx[4], fd[4];
        qx.get(fx);
        qd.get(fd);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = (fx[j] - mean) * rstd;
          float dz = fd[j];
          if (act == EALDM_ACT_SILU) dz *= dsilu_f(fmaf(gv[j], xh, bv[j]));
          a[j] += dz;
          b[j] = fmaf(dz, xh, b[j]);
        }
      }
    }
    if (pix_lanes > 1) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) { red[t][j] = a[j]; red[t][4 + j] = b[j]; }
      __syncthreads();
      if (tp == 0 && v < vpp) {
        for (int k = 1; k < pix_lanes; ++k)
#pragma unroll
          for (int j = 0; j < 4; ++j) { a[j] += red[k * lanes_v + tv][j]; b[j] += red[k * lanes_v + tv][4 + j]; }
      }
    }
    if (tp == 0 && v < vpp) {
#pragma unroll
      for (int j = 0; j < 4; ++j) out[v * 4 + j] = make_float2(a[j], b[j]);
    }
  }
}

// one CTA per image: ab[n][c] = sum over chunks (fixed order); gs[n][g] = {s1, s2}
__global__ void __launch_bounds__(NT)
gn_bwd_finalize_kernel(const float2* __restrict__ part, int chunks, int c, int groups,
                       const float* __restrict__ gamma, float2* __restrict__ ab, float2* __restrict__ gs) {
  const int n = blockIdx.x;
  const float2* pn = part + static_cast<long long>(n) * chunks * c;
  float2* abn = ab + static_cast<long long>(n) * c;
  for (int ch = threadIdx.x; ch < c; ch += NT) {
    float sa = 0.f, sb = 0.f;
    for (int k = 0; k < chunks; ++k) {
      const float2 e = pn[static_cast<long long>(k) * c + ch];
      sa += e.x;
      sb += e.y;
    }
    abn[ch] = make_float2(sa, sb);
  }
  __syncthreads();
  const int cpg = c / groups;
  for (int g = threadIdx.x; g < groups; g += NT) {
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < cpg; ++j) {
      const float2 e = abn[g * cpg + j];
      const double gm = static_cast<double>(gamma[g * cpg + j]);
      s1 += gm * e.x;
      s2 += gm * e.y;
    }
    gs[static_cast<long long>(n) * groups + g] = make_float2(static_cast<float>(s1), static_cast<float>(s2));
  }
}

__global__ void __launch_bounds__(NT)
gn_bwd_param_kernel(const float2* __restrict__ ab, int n, int c, float* __restrict__ dgamma,
                    float* __restrict__ dbeta) {
  const int ch = blockIdx.x * NT + threadIdx.x;
  if (ch >= c) return;
  float sa = 0.f, sb = 0.f;
  for (int i = 0; i < n; ++i) {
    const float2 e = ab[static_cast<long long>(i) * c + ch];
    sa += e.x;
    sb += e.y;
  }
  if (dgamma) dgamma[ch] += sb;
  if (dbeta) dbeta[ch] += sa;
}

template <typename TX, typename TD, typename TO>
__global__ void __launch_bounds__(NT)
gn_bwd_apply_kernel(const TX* __restrict__ x, long long ld_x, const TD* __restrict__ dy, long long ld_dy, int hw,
                    int c, int groups, int pix_per_cta, const float* __restrict__ stats,
                    const float2* __restrict__ gs, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int act, const float* __restrict__ add, long long ld_add, const float* __restrict__ add2,
                    long long ld_add2, TO* __restrict__ dx, long long ld_dx, TD* __restrict__ dx2,
                    long long ld_dx2) {
  const int t = threadIdx.x;
  const int n = blockIdx.y;
  const int vpp = c >> 2;
  const int cpg = c / groups;
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const int lanes_v = vpp < NT ? vpp : NT;
  const int pix_lanes = NT / lanes_v;
  const int tv = t % lanes_v, tp = t / lanes_v;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const long long row0 = static_cast<long long>(n) * hw;
  if (tp >= pix_lanes) return;
  for (int v = tv; v < vpp; v += lanes_v) {
    const int ch = v * 4;
    const int g = ch / cpg;
    const float mean = stats[(static_cast<long long>(n) * groups + g) * 2];
    const float rstd = stats[(static_cast<long long>(n) * groups + g) * 2 + 1];
    const float2 s = gs[static_cast<long long>(n) * groups + g];
    const float k1 = s.x * inv_m, k2 = s.y * inv_m;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + ch));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + ch));
    const float gv[4] = {ga.x, ga.y, ga.z, ga.w}, bv[4] = {be.x, be.y, be.z, be.w};
    for (int pix = p0 + tp; pix < p1; pix += pix_lanes) {
      const long long r = row0 + pix;
      Vec4<TX> qx;
      Vec4<TD> qd;
      qx.load(x + r * ld_x + ch);
      qd.load(dy + r * ld_dy + ch);
      float fx[4], fd[4], o[4];
      qx.get(fx);
      qd.get(fd);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (fx[j] - mean) * rstd;
        float dz = fd[j];
        if (act == EALDM_ACT_SILU) dz *= dsilu_f(fmaf(gv[j], xh, bv[j]));
        o[j] = rstd * (gv[j] * dz - fmaf(xh, k2, k1));
      }
      if (add) {
        const float4 q = *reinterpret_cast<const float4*>(add + r * ld_add + ch);
        o[0] += q.x; o[1] += q.y; o[2] += q.z; o[3] += q.w;
      }
      if (add2) {
        const float4 q = *reinterpret_cast<const float4*>(add2 + r * ld_add2 + ch);
        o[0] += q.x; o[1] += q.y; o[2] += q.z; o[3] += q.w;
      }
      Vec4<TO> qo;
      qo.set(o);
      qo.store(dx + r * ld_dx + ch);
      if (dx2) {
        Vec4<TD> q2;
        q2.set(o);
        q2.store(dx2 + r * ld_dx2 + ch);
      }
    }
  }
}

// ---- LayerNorm backward: one warp per row, the row lives in registers ----------------------------------
template <typename TX, typename TD, typename TO, int NV>  // c == NV * 128
__global__ void __launch_bounds__(NT)
ln_bwd_kernel(const TX* __restrict__ x, long long ld_x, const TD* __restrict__ dy, long long ld_dy, long long rows,
              float eps, const float* __restrict__ gamma, const float* __restrict__ add, long long ld_add,
              TO* __restrict__ dx, long long ld_dx, TD* __restrict__ dx2, long long ld_dx2,
              float2* __restrict__ part) {
  constexpr int C_ = NV * 128;
  __shared__ float2 acc_s[C_];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gm[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(gamma + (k * 32 + lane) * 4));
    gm[k][0] = q.x; gm[k][1] = q.y; gm[k][2] = q.z; gm[k][3] = q.w;
  }
  float dg[NV][4], dbt[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) dg[k][j] = dbt[k][j] = 0.f;

  for (long long row = static_cast<long long>(blockIdx.x) * (NT / 32) + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * (NT / 32)) {
    float f[NV][4], d[NV][4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      Vec4<TX> q;
      q.load(x + row * ld_x + (k * 32 + lane) * 4);
      q.get(f[k]);
      Vec4<TD> qd;
      qd.load(dy + row * ld_dy + (k * 32 + lane) * 4);
      qd.get(d[k]);
      s += (f[k][0] + f[k][1]) + (f[k][2] + f[k][3]);
    }
    const float mean = warp_sum(s) / static_cast<float>(C_);
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) { f[k][j] -= mean; ss = fmaf(f[k][j], f[k][j], ss); }
    const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(C_) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[k][j] *= rstd;  // xhat
        dg[k][j] = fmaf(d[k][j], f[k][j], dg[k][j]);
        dbt[k][j] += d[k][j];
        d[k][j] *= gm[k][j];  // g = dy * gamma
        s1 += d[k][j];
        s2 = fmaf(d[k][j], f[k][j], s2);
      }
    s1 = warp_sum(s1) / static_cast<float>(C_);
    s2 = warp_sum(s2) / static_cast<float>(C_);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = rstd * (d[k][j] - fmaf(f[k][j], s2, s1));
      const int col = (k * 32 + lane) * 4;
      if (add) {
        const float4 q = *reinterpret_cast<const float4*>(add + row * ld_add + col);
        o[0] += q.x; o[1] += q.y; o[2] += q.z; o[3] += q.w;
      }
      Vec4<TO> qo;
      qo.set(o);
      qo.store(dx + row * ld_dx + col);
      if (dx2) {
        Vec4<TD> q2;
        q2.set(o);
        q2.store(dx2 + row * ld_dx2 + col);
      }
    }
  }
  // CTA partial of dgamma / dbeta: warps add their registers in warp order (fixed order)
  for (int w = 0; w < NT / 32; ++w) {
    if (warp == w) {
#pragma unroll
      for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = (k * 32 + lane) * 4 + j;
          float2 e = w == 0 ? make_float2(0.f, 0.f) : acc_s[col];
          e.x += dg[k][j];
          e.y += dbt[k][j];
          acc_s[col] = e;
        }
    }
    __syncthreads();
  }
  for (int col = threadIdx.x; col < C_; col += NT) part[static_cast<long long>(blockIdx.x) * C_ + col] = acc_s[col];
}

// 32 columns x 8 partial lanes per CTA, lanes combined in a fixed order
__global__ void __launch_bounds__(NT)
ln_bwd_param_kernel(const float2* __restrict__ part, int nparts, int c, float* __restrict__ dgamma,
                    float* __restrict__ dbeta) {
  __shared__ float2 red[8][32];
  const int tc = threadIdx.x & 31, tk = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + tc;
  float sg = 0.f, sb = 0.f;
  if (ch < c)
    for (int i = tk; i < nparts; i += 8) {
      const float2 e = part[static_cast<long long>(i) * c + ch];
      sg += e.x;
      sb += e.y;
    }
  red[tk][tc] = make_float2(sg, sb);
  __syncthreads();
  if (tk == 0 && ch < c) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { sg += red[k][tc].x; sb += red[k][tc].y; }
    if (dgamma) dgamma[ch] += sg;
    if (dbeta) dbeta[ch] += sb;
  }
}

// ---- GEGLU (natural [value | gate] layout, attention.py:37-44) ------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT)
geglu_fwd_kernel(const T* __restrict__ pre, long long ld_pre, long long rows, int inner, T* __restrict__ out,
                 long long ld_out) {
  const int vpr = inner >> 2;
  const long long total = rows * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < total;
       i += static_cast<long long>(NT) * gridDim.x) {
    const long long r = i / vpr;
    const int col = static_cast<int>(i - r * vpr) * 4;
    Vec4<T> qv, qg;
    qv.load(pre + r * ld_pre + col);
    qg.load(pre + r * ld_pre + inner + col);
    float v[4], g[4], o[4];
    qv.get(v);
    qg.get(g);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = v[j] * gelu_erf_f(g[j]);
    Vec4<T> qo;
    qo.set(o);
    qo.store(out + r * ld_out + col);
  }
}
template <typename T>
__global__ void __launch_bounds__(NT)
geglu_bwd_kernel(const T* __restrict__ pre, long long ld_pre, const T* __restrict__ dout, long long ld_dout,
                 long long rows, int inner, T* __restrict__ dpre, long long ld_dpre) {
  const int vpr = inner >> 2;
  const long long total = rows * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < total;
       i += static_cast<long long>(NT) * gridDim.x) {
    const long long r = i / vpr;
    const int col = static_cast<int>(i - r * vpr) * 4;
    Vec4<T> qv, qg, qd;
    qv.load(pre + r * ld_pre + col);
    qg.load(pre + r * ld_pre + inner + col);
    qd.load(dout + r * ld_dout + col);
    float v[4], g[4], d[4], dv[4], dg[4];
    qv.get(v);
    qg.get(g);
    qd.get(d);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dv[j] = d[j] * gelu_erf_f(g[j]);
      dg[j] = d[j] * v[j] * dgelu_erf_f(g[j]);
    }
    Vec4<T> q1, q2;
    q1.set(dv);
    q2.set(dg);
    q1.store(dpre + r * ld_dpre + col);
    q2.store(dpre + r * ld_dpre + inner + col);
  }
}

// ---- SiLU forward / backward on [rows, c] -----------------------------------------------------------------
template <typename TX, typename TY>
__global__ void __launch_bounds__(NT)
silu_fwd_kernel(const TX* __restrict__ x, long long ld_x, long long rows, int c, TY* __restrict__ y, long long ld_y) {
  const long long total = rows * c;
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < total;
       i += static_cast<long long>(NT) * gridDim.x) {
    const long long r = i / c;
    const int col = static_cast<int>(i - r * c);
    y[r * ld_y + col] = from_f32<TY>(silu_f(to_f32(x[r * ld_x + col])));
  }
}
// dx = dy * silu'(x)
template <typename TX, typename TD, typename TO>
__global__ void __launch_bounds__(NT)
silu_bwd_kernel(const TX* __restrict__ x, long long ld_x, const TD* __restrict__ dy, long long ld_dy, long long rows,
                int c, TO* __restrict__ dx, long long ld_dx) {
  const long long total = rows * c;
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < total;
       i += static_cast<long long>(NT) * gridDim.x) {
    const long long r = i / c;
    const int col = static_cast<int>(i - r * c);
    dx[r * ld_dx + col] = from_f32<TO>(to_f32(dy[r * ld_dy + col]) * dsilu_f(to_f32(x[r * ld_x + col])));
  }
}

// ---- column sums per segment: out[s][col] (+)= sum over the segment's rows of x[row][col] -------------------
static void colsum_chunking(long long segs, long long rows_per_seg, int* rows_per_cta, long long* chunks) {
  long long ch = ceil_div(592, segs);   // ~4 CTAs per SM over all segments
  const long long max_chunks = ceil_div(rows_per_seg, 32);
  if (ch > max_chunks) ch = max_chunks;
  if (ch < 1) ch = 1;
  const int rpc = static_cast<int>(ceil_div(rows_per_seg, ch));
  *rows_per_cta = rpc;
  *chunks = ceil_div(rows_per_seg, rpc);
}

// 4 consecutive columns per thread (16 B of fp32 / 8 B of bf16 per load), 8 independent loads in flight
template <typename T>
__global__ void __launch_bounds__(NT)
colsum_partial_kernel(const T* __restrict__ x, long long ld, long long rows_per_seg, int c, int rows_per_cta,
                      float* __restrict__ part) {
  __shared__ float4 red[NT];
  const int t = threadIdx.x;
  const int seg = blockIdx.y;
  const int vpr = (c + 3) >> 2;                 // vec4 per row (the last one may be ragged)
  const int lanes_v = vpr < NT ? vpr : NT;
  const int row_lanes = NT / lanes_v;
  const int tv = t % lanes_v, tr = t / lanes_v;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, rows_per_seg);
  const T* xb = x + static_cast<long long>(seg) * rows_per_seg * ld;
  float* out = part + (static_cast<long long>(seg) * gridDim.x + blockIdx.x) * c;
  const bool vec_ok = (c & 3) == 0;             // host guarantees ld % 4 == 0 and aligned base in that case
  for (int v0 = 0; v0 < vpr; v0 += lanes_v) {
    const int v = v0 + tv;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (tr < row_lanes && v < vpr) {
      const int col = v * 4;
      if (vec_ok) {
        constexpr int U = 8;
        for (long long r = r0 + tr; r < r1; r += static_cast<long long>(U) * row_lanes) {
          Vec4<T> q[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const long long rr = r + static_cast<long long>(u) * row_lanes;
            if (rr < r1) q[u].load(xb + rr * ld + col);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (r + static_cast<long long>(u) * row_lanes < r1) {
              float f[4];
              q[u].get(f);
              s[0] += f[0]; s[1] += f[1]; s[2] += f[2]; s[3] += f[3];
            }
          }
        }
      } else {
        for (long long r = r0 + tr; r < r1; r += row_lanes)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (col + j < c) s[j] += to_f32(xb[r * ld + col + j]);
      }
    }
    if (row_lanes > 1) {
      __syncthreads();
      red[t] = make_float4(s[0], s[1], s[2], s[3]);
      __syncthreads();
      if (tr == 0 && v < vpr)
        for (int k = 1; k < row_lanes; ++k) {
          const float4 e = red[k * lanes_v + tv];
          s[0] += e.x; s[1] += e.y; s[2] += e.z; s[3] += e.w;
        }
    }
    if (tr == 0 && v < vpr) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (v * 4 + j < c) out[v * 4 + j] = s[j];
    }
  }
}
// 32 columns x 8 chunk lanes per CTA; lanes are combined in a fixed order
__global__ void __launch_bounds__(NT)
colsum_final_kernel(const float* __restrict__ part, int chunks, int c, float* __restrict__ out, long long ld_out,
                    int accumulate) {
  __shared__ float red[8][32];
  const int tc = threadIdx.x & 31, tk = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tc;
  const int seg = blockIdx.y;
  const float* p = part + static_cast<long long>(seg) * chunks * c;
  float s = 0.f;
  if (col < c)
    for (int k = tk; k < chunks; k += 8) s += p[static_cast<long long>(k) * c + col];
  red[tk][tc] = s;
  __syncthreads();
  if (tk == 0 && col < c) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][tc];
    float* o = out + static_cast<long long>(seg) * ld_out + col;
    *o = accumulate ? *o + s : s;
  }
}

// ---- the same in ONE launch (c % 4 == 0): a CTA sums 128 columns (4 per lane) of one row chunk, 8 rows at a time,
// writes its partial, and the LAST CTA of a (segment, column block) to finish -- found with a ticket counter, no spinning
// -- adds the partials in chunk order (fixed, so the sums are reproducible) and writes / accumulates the result.  The
// 183 column sums of a training step were 2 launches each; the second one (~8 us of latency for a few KB) is gone.
template <typename T>
__global__ void __launch_bounds__(NT)
colsum_ticket_kernel(const T* __restrict__ x, long long ld, long long rows_per_seg, int c, int rows_per_cta,
                     float* __restrict__ part, unsigned int* __restrict__ tickets, float* __restrict__ out,
                     long long ld_out, int accumulate) {
  __shared__ float4 red[NT / 32][32];
  __shared__ unsigned int ticket_s;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int cb = blockIdx.x, chunk = blockIdx.y, seg = blockIdx.z;
  const int chunks = gridDim.y;
  const int col = cb * 128 + lane * 4;
  const bool live = col < c;
  const long long r0 = static_cast<long long>(chunk) * rows_per_cta;
  const long long r1 = min(r0 + rows_per_cta, rows_per_seg);
  const T* xb = x + static_cast<long long>(seg) * rows_per_seg * ld + col;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    constexpr int U = 8;
    for (long long r = r0 + wp; r < r1; r += static_cast<long long>(U) * (NT / 32)) {
      Vec4<T> q[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long rr = r + static_cast<long long>(u) * (NT / 32);
        if (rr < r1) q[u].load(xb + rr * ld);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + static_cast<long long>(u) * (NT / 32) < r1) {
          float f[4];
          q[u].get(f);
          s[0] += f[0]; s[1] += f[1]; s[2] += f[2]; s[3] += f[3];
        }
      }
    }
  }
  red[wp][lane] = make_float4(s[0], s[1], s[2], s[3]);
  __syncthreads();
  float* pbase = part + (static_cast<long long>(seg) * gridDim.x + cb) * chunks * 128;
  if (wp == 0) {
    float4 a = red[0][lane];
#pragma unroll
    for (int k = 1; k < NT / 32; ++k) {
      const float4 e = red[k][lane];
      a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
    }
    __stcg(reinterpret_cast<float4*>(pbase + static_cast<long long>(chunk) * 128) + lane, a);
    __threadfence();
    __syncwarp();
    if (lane == 0) ticket_s = atomicAdd(tickets + seg * gridDim.x + cb, 1u);
  }
  __syncthreads();
  if (ticket_s != static_cast<unsigned int>(chunks - 1)) return;
  __threadfence();   // the last CTA: every partial of this (segment, column block) is visible
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = wp; k < chunks; k += NT / 32) {
    const float4 e = __ldcg(reinterpret_cast<const float4*>(pbase + static_cast<long long>(k) * 128) + lane);
    a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
  }
  __syncthreads();
  red[wp][lane] = a;
  __syncthreads();
  if (wp == 0) {
    a = red[0][lane];
#pragma unroll
    for (int k = 1; k < NT / 32; ++k) {
      const float4 e = red[k][lane];
      a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
    }
    if (live) {
      float* o = out + static_cast<long long>(seg) * ld_out + col;
      if (accumulate) { o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w; }
      else { o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; }
    }
    if (lane == 0) tickets[seg * gridDim.x + cb] = 0u;   // re-armed for the next launch on this stream
  }
}
static void colsum_ticket_chunking(long long segs, long long rows_per_seg, long long c, int* rows_per_cta,
                                   long long* chunks) {
  const long long cbs = ceil_div(c, 128);
  long long ch = ceil_div(592, segs * cbs);   // ~4 CTAs per SM over all segments and column blocks
  const long long max_chunks = ceil_div(rows_per_seg, 64);
  if (ch > max_chunks) ch = max_chunks;
  if (ch < 1) ch = 1;
  const int rpc = static_cast<int>(ceil_div(rows_per_seg, ch));
  *rows_per_cta = rpc;
  *chunks = ceil_div(rows_per_seg, rpc);
}
constexpr int COLSUM_TICKETS = 1 << 16;
static StreamScratch g_colsum_tickets(COLSUM_TICKETS * sizeof(unsigned int));   // see common.cuh
static unsigned int* colsum_tickets(cudaStream_t st) { return static_cast<unsigned int*>(g_colsum_tickets.get(st)); }

// ---- resampling adjoints ------------------------------------------------------------------------------------
// z[n, 2*oh, 2*ow, :] = dy[n, oh, ow, :], zero elsewhere (adjoint of reading every second pixel)
template <typename T>
__global__ void __launch_bounds__(NT)
zero_insert2x_kernel(const T* __restrict__ dy, long long ld_dy, int n, int h, int w, int c, T* __restrict__ z,
                     long long ld_z) {
  const int vpp = c >> 2;
  const long long total = static_cast<long long>(n) * (2 * h) * (2 * w) * vpp;
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < total;
       i += static_cast<long long>(NT) * gridDim.x) {
    const int v = static_cast<int>(i % vpp);
    const long long pix = i / vpp;
    const int ow = static_cast<int>(pix % (2 * w));
    const int oh = static_cast<int>((pix / (2 * w)) % (2 * h));
    const int img = static_cast<int>(pix / (4LL * w * h));
    Vec4<T> q;
    if (((ow | oh) & 1) == 0) {
      q.load(dy + ((static_cast<long long>(img) * h + (oh >> 1)) * w + (ow >> 1)) * ld_dy + v * 4);
    } else {
      const float zf[4] = {0.f, 0.f, 0.f, 0.f};
      q.set(zf);
    }
    q.store(z + pix * ld_z + v * 4);
  }
}
// dx[n, h, w, :] = sum of the 2x2 block of dup (adjoint of nearest-2x upsampling) (+ add)
template <typename T, typename TO>
__global__ void __launch_bounds__(NT)
sumpool2x2_kernel(const T* __restrict__ dup, long long ld_dup, int n, int h, int w, int c,
                  const float* __restrict__ add, long long ld_add, TO* __restrict__ dx, long long ld_dx,
                  T* __restrict__ dx2, long long ld_dx2) {
  const int vpp = c >> 2;
  const long long total = static_cast<long long>(n) * h * w * vpp;
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < total;
       i += static_cast<long long>(NT) * gridDim.x) {
    const int v = static_cast<int>(i % vpp);
    const long long pix = i / vpp;
    const int ow = static_cast<int>(pix % w);
    const int oh = static_cast<int>((pix / w) % h);
    const int img = static_cast<int>(pix / (static_cast<long long>(w) * h));
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        Vec4<T> q;
        q.load(dup + ((static_cast<long long>(img) * 2 * h + 2 * oh + a) * (2 * w) + 2 * ow + b) * ld_dup + v * 4);
        float f[4];
        q.get(f);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] += f[j];
      }
    if (add) {
      const float4 q = *reinterpret_cast<const float4*>(add + pix * ld_add + v * 4);
      o[0] += q.x; o[1] += q.y; o[2] += q.z; o[3] += q.w;
    }
    Vec4<TO> qo;
    qo.set(o);
    qo.store(dx + pix * ld_dx + v * 4);
    if (dx2) {
      Vec4<T> q2;
      q2.set(o);
      q2.store(dx2 + pix * ld_dx2 + v * 4);
    }
  }
}

static int grid_for(long long total) {
  const long long b = ceil_div(total, NT);
  return static_cast<int>(b < 148 * 16 ? (b < 1 ? 1 : b) : 148 * 16);
}

template <typename TX, typename TD, typename TO>
static int group_norm_bwd_t(const ealdm_group_norm_bwd_args* a, cudaStream_t st) {
  int ppc;
  long long chunks;
  chunking(a->n, a->hw, &ppc, &chunks);
  const int hw = static_cast<int>(a->hw), c = static_cast<int>(a->c);
  float2* part = reinterpret_cast<float2*>(a->workspace);
  float2* ab = part + a->n * chunks * c;
  float2* gs = ab + a->n * c;
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(a->n));
  const TX* x = reinterpret_cast<const TX*>(a->x);
  const TD* dy = reinterpret_cast<const TD*>(a->dy);
  gn_bwd_stats_kernel<TX, TD><<<grid, NT, 0, st>>>(x, a->ld_x, dy, a->ld_dy, hw, c, a->groups, ppc, a->stats,
                                                   a->gamma, a->beta, a->act, part);
  EALDM_LAUNCH_CHECK();
  gn_bwd_finalize_kernel<<<static_cast<unsigned>(a->n), NT, 0, st>>>(part, static_cast<int>(chunks), c, a->groups,
                                                                     a->gamma, ab, gs);
  EALDM_LAUNCH_CHECK();
  if (a->dgamma || a->dbeta) {
    gn_bwd_param_kernel<<<static_cast<unsigned>(ceil_div(c, NT)), NT, 0, st>>>(ab, static_cast<int>(a->n), c,
                                                                               a->dgamma, a->dbeta);
    EALDM_LAUNCH_CHECK();
  }
  gn_bwd_apply_kernel<TX, TD, TO><<<grid, NT, 0, st>>>(
      x, a->ld_x, dy, a->ld_dy, hw, c, a->groups, ppc, a->stats, gs, a->gamma, a->beta, a->act, a->add, a->ld_add,
      a->add2, a->ld_add2, reinterpret_cast<TO*>(a->dx), a->ld_dx, reinterpret_cast<TD*>(a->dx2), a->ld_dx2);
  EALDM_LAUNCH_CHECK();
  return 0;
}

template <typename TX, typename TD, typename TO>
static int layer_norm_bwd_t(const ealdm_layer_norm_bwd_args* a, int nparts, cudaStream_t st) {
  float2* part = reinterpret_cast<float2*>(a->workspace);
  const TX* x = reinterpret_cast<const TX*>(a->x);
  const TD* dy = reinterpret_cast<const TD*>(a->dy);
  TO* dx = reinterpret_cast<TO*>(a->dx);
  TD* dx2 = reinterpret_cast<TD*>(a->dx2);
#define EALDM_LN_BWD(NV)                                                                                        \
  ln_bwd_kernel<TX, TD, TO, NV><<<nparts, NT, 0, st>>>(x, a->ld_x, dy, a->ld_dy, a->rows, a->eps, a->gamma, a->add, \
                                                       a->ld_add, dx, a->ld_dx, dx2, a->ld_dx2, part)
  switch (a->c / 128) {
    case 1: EALDM_LN_BWD(1); break;
    case 2: EALDM_LN_BWD(2); break;
    case 4: EALDM_LN_BWD(4); break;
    case 8: EALDM_LN_BWD(8); break;
    default: return set_error(EALDM_EUNSUPPORTED, "layer_norm_bwd: c must be 128, 256, 512 or 1024");
  }
#undef EALDM_LN_BWD
  EALDM_LAUNCH_CHECK();
  if (a->dgamma || a->dbeta) {
    ln_bwd_param_kernel<<<static_cast<unsigned>(ceil_div(a->c, 32)), NT, 0, st>>>(part, nparts, static_cast<int>(a->c),
                                                                                  a->dgamma, a->dbeta);
    EALDM_LAUNCH_CHECK();
  }
  return 0;
}

static int ln_parts(long long rows) {
  const long long want = ceil_div(rows, NT / 32);
  return static_cast<int>(want < 148 * 4 ? want : 148 * 4);
}

}  // namespace bwd
}  // namespace ealdm

using namespace ealdm;

extern "C" int64_t ealdm_group_norm_bwd_workspace_bytes(int64_t n, int64_t hw, int64_t c) {
  if (n <= 0 || hw <= 0 || c <= 0) return 0;
  int ppc;
  long long chunks;
  bwd::chunking(n, hw, &ppc, &chunks);
  return (n * chunks * c + n * c + n * bwd::MAX_GROUPS) * static_cast<int64_t>(sizeof(float2));
}

extern "C" int ealdm_group_norm_bwd(const ealdm_group_norm_bwd_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->x && a->dy && a->dx && a->stats && a->gamma && a->beta && a->workspace,
                "group_norm_bwd: null argument");
  EALDM_REQUIRE(a->groups > 0 && a->groups <= bwd::MAX_GROUPS && a->c % a->groups == 0 && (a->c / a->groups) % 4 == 0,
                "group_norm_bwd: bad groups / channels");
  EALDM_REQUIRE(a->ld_x % 4 == 0 && a->ld_dy % 4 == 0 && a->ld_dx % 4 == 0 && a->ld_add % 4 == 0 &&
                    a->ld_add2 % 4 == 0 && a->ld_dx2 % 4 == 0,
                "group_norm_bwd: pitches must be multiples of 4");
  EALDM_REQUIRE(a->n > 0 && a->n <= 65535 && a->hw > 0, "group_norm_bwd: bad n/hw");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->dtype == EALDM_F32) return bwd::group_norm_bwd_t<float, float, float>(a, st);
  EALDM_REQUIRE(a->dtype == EALDM_BF16, "group_norm_bwd: bad dtype");
  if (a->x_f32)
    return a->dx_f32 ? bwd::group_norm_bwd_t<float, bf16, float>(a, st) : bwd::group_norm_bwd_t<float, bf16, bf16>(a, st);
  return a->dx_f32 ? bwd::group_norm_bwd_t<bf16, bf16, float>(a, st) : bwd::group_norm_bwd_t<bf16, bf16, bf16>(a, st);
}

extern "C" int64_t ealdm_layer_norm_bwd_workspace_bytes(int64_t rows, int64_t c) {
  if (rows <= 0 || c <= 0) return 0;
  return static_cast<int64_t>(bwd::ln_parts(rows)) * c * static_cast<int64_t>(sizeof(float2));
}

extern "C" int ealdm_layer_norm_bwd(const ealdm_layer_norm_bwd_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->x && a->dy && a->dx && a->gamma && a->workspace, "layer_norm_bwd: null argument");
  EALDM_REQUIRE(a->rows > 0 && a->c > 0 && a->c % 128 == 0, "layer_norm_bwd: c must be a multiple of 128");
  EALDM_REQUIRE(a->ld_x % 4 == 0 && a->ld_dy % 4 == 0 && a->ld_dx % 4 == 0 && a->ld_add % 4 == 0 && a->ld_dx2 % 4 == 0,
                "layer_norm_bwd: pitches must be multiples of 4");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nparts = bwd::ln_parts(a->rows);
  if (a->dtype == EALDM_F32) return bwd::layer_norm_bwd_t<float, float, float>(a, nparts, st);
  EALDM_REQUIRE(a->dtype == EALDM_BF16, "layer_norm_bwd: bad dtype");
  EALDM_REQUIRE(a->x_f32 && a->dx_f32, "layer_norm_bwd: the bf16 path takes an fp32 x and writes an fp32 dx");
  return bwd::layer_norm_bwd_t<float, bf16, float>(a, nparts, st);
}

extern "C" int ealdm_geglu(const void* pre, int64_t ld_pre, int32_t dtype, int64_t rows, int64_t inner, void* out,
                           int64_t ld_out, ealdm_stream_t stream) {
  EALDM_REQUIRE(pre && out && rows > 0 && inner > 0 && inner % 4 == 0 && ld_pre % 4 == 0 && ld_out % 4 == 0,
                "geglu: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = bwd::grid_for(rows * (inner / 4));
  if (dtype == EALDM_F32)
    bwd::geglu_fwd_kernel<float><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const float*>(pre), ld_pre, rows,
                                                           (int)inner, reinterpret_cast<float*>(out), ld_out);
  else
    bwd::geglu_fwd_kernel<bf16><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const bf16*>(pre), ld_pre, rows, (int)inner,
                                                          reinterpret_cast<bf16*>(out), ld_out);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_geglu_bwd(const void* pre, int64_t ld_pre, const void* dout, int64_t ld_dout, int32_t dtype,
                               int64_t rows, int64_t inner, void* dpre, int64_t ld_dpre, ealdm_stream_t stream) {
  EALDM_REQUIRE(pre && dout && dpre && rows > 0 && inner > 0 && inner % 4 == 0 && ld_pre % 4 == 0 &&
                    ld_dout % 4 == 0 && ld_dpre % 4 == 0,
                "geglu_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = bwd::grid_for(rows * (inner / 4));
  if (dtype == EALDM_F32)
    bwd::geglu_bwd_kernel<float><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const float*>(pre), ld_pre,
                                                           reinterpret_cast<const float*>(dout), ld_dout, rows,
                                                           (int)inner, reinterpret_cast<float*>(dpre), ld_dpre);
  else
    bwd::geglu_bwd_kernel<bf16><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const bf16*>(pre), ld_pre,
                                                          reinterpret_cast<const bf16*>(dout), ld_dout, rows,
                                                          (int)inner, reinterpret_cast<bf16*>(dpre), ld_dpre);
  EALDM_LAUNCH_CHECK();
  return 0;
}

// y = silu(x): x fp32, y `dtype`
extern "C" int ealdm_silu(const float* x, int64_t ld_x, int64_t rows, int64_t c, int32_t dtype, void* y, int64_t ld_y,
                          ealdm_stream_t stream) {
  EALDM_REQUIRE(x && y && rows > 0 && c > 0, "silu: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = bwd::grid_for(rows * c);
  if (dtype == EALDM_F32)
    bwd::silu_fwd_kernel<float, float><<<grid, bwd::NT, 0, st>>>(x, ld_x, rows, (int)c, reinterpret_cast<float*>(y), ld_y);
  else
    bwd::silu_fwd_kernel<float, bf16><<<grid, bwd::NT, 0, st>>>(x, ld_x, rows, (int)c, reinterpret_cast<bf16*>(y), ld_y);
  EALDM_LAUNCH_CHECK();
  return 0;
}

// dx = dy * silu'(x): x fp32, dy `dtype`, dx `dtype`
extern "C" int ealdm_silu_bwd(const float* x, int64_t ld_x, const void* dy, int64_t ld_dy, int32_t dtype, int64_t rows,
                              int64_t c, void* dx, int64_t ld_dx, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && dy && dx && rows > 0 && c > 0, "silu_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = bwd::grid_for(rows * c);
  if (dtype == EALDM_F32)
    bwd::silu_bwd_kernel<float, float, float><<<grid, bwd::NT, 0, st>>>(x, ld_x, reinterpret_cast<const float*>(dy),
                                                                        ld_dy, rows, (int)c,
                                                                        reinterpret_cast<float*>(dx), ld_dx);
  else
    bwd::silu_bwd_kernel<float, bf16, bf16><<<grid, bwd::NT, 0, st>>>(x, ld_x, reinterpret_cast<const bf16*>(dy), ld_dy,
                                                                      rows, (int)c, reinterpret_cast<bf16*>(dx), ld_dx);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t ealdm_colsum_workspace_bytes(int64_t segs, int64_t rows_per_seg, int64_t c) {
  if (segs <= 0 || rows_per_seg <= 0 || c <= 0) return 0;
  int ppc;
  long long chunks, chunks_t;
  bwd::colsum_chunking(segs, rows_per_seg, &ppc, &chunks);
  bwd::colsum_ticket_chunking(segs, rows_per_seg, c, &ppc, &chunks_t);
  const int64_t two_kernel = segs * chunks * c * 4;
  const int64_t ticket = segs * ceil_div(c, 128) * chunks_t * 128 * 4;
  return two_kernel > ticket ? two_kernel : ticket;
}

extern "C" int ealdm_colsum(const void* x, int64_t ld_x, int32_t dtype, int64_t segs, int64_t rows_per_seg, int64_t c,
                            float* out, int64_t ld_out, int32_t accumulate, void* workspace, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && out && workspace && segs > 0 && segs <= 65535 && rows_per_seg > 0 && c > 0, "colsum: bad arguments");
  EALDM_REQUIRE(c % 4 != 0 || (ld_x % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & (dtype == EALDM_F32 ? 15 : 7)) == 0),
                "colsum: with c %% 4 == 0 the rows must be 4-element aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int ppc;
  long long chunks;
  float* part = reinterpret_cast<float*>(workspace);
  static const bool one_kernel = getenv("EALDM_COLSUM_TWO_KERNELS") == nullptr;
  const long long cbs = ceil_div(c, 128);
  if (one_kernel && c % 4 == 0 && segs * cbs <= bwd::COLSUM_TICKETS && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0) {
    if (unsigned int* tickets = bwd::colsum_tickets(st)) {
      bwd::colsum_ticket_chunking(segs, rows_per_seg, c, &ppc, &chunks);
      dim3 grid(static_cast<unsigned>(cbs), static_cast<unsigned>(chunks), static_cast<unsigned>(segs));
      if (dtype == EALDM_F32)
        bwd::colsum_ticket_kernel<float><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const float*>(x), ld_x, rows_per_seg,
                                                                   (int)c, ppc, part, tickets, out, ld_out, accumulate);
      else
        bwd::colsum_ticket_kernel<bf16><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const bf16*>(x), ld_x, rows_per_seg,
                                                                  (int)c, ppc, part, tickets, out, ld_out, accumulate);
      EALDM_LAUNCH_CHECK();
      return 0;
    }
  }
  bwd::colsum_chunking(segs, rows_per_seg, &ppc, &chunks);
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(segs));
  if (dtype == EALDM_F32)
    bwd::colsum_partial_kernel<float><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const float*>(x), ld_x, rows_per_seg,
                                                                (int)c, ppc, part);
  else
    bwd::colsum_partial_kernel<bf16><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const bf16*>(x), ld_x, rows_per_seg,
                                                               (int)c, ppc, part);
  EALDM_LAUNCH_CHECK();
  dim3 g2(static_cast<unsigned>(ceil_div(c, 32)), static_cast<unsigned>(segs));
  bwd::colsum_final_kernel<<<g2, bwd::NT, 0, st>>>(part, static_cast<int>(chunks), (int)c, out, ld_out, accumulate);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_zero_insert2x(const void* dy, int64_t ld_dy, int32_t dtype, int64_t n, int64_t h, int64_t w,
                                   int64_t c, void* z, int64_t ld_z, ealdm_stream_t stream) {
  EALDM_REQUIRE(dy && z && n > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0 && ld_dy % 4 == 0 && ld_z % 4 == 0,
                "zero_insert2x: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = bwd::grid_for(n * 4 * h * w * (c / 4));
  if (dtype == EALDM_F32)
    bwd::zero_insert2x_kernel<float><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const float*>(dy), ld_dy, (int)n, (int)h,
                                                               (int)w, (int)c, reinterpret_cast<float*>(z), ld_z);
  else
    bwd::zero_insert2x_kernel<bf16><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const bf16*>(dy), ld_dy, (int)n, (int)h,
                                                              (int)w, (int)c, reinterpret_cast<bf16*>(z), ld_z);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_sumpool2x2(const void* dup, int64_t ld_dup, int32_t dtype, int64_t n, int64_t h, int64_t w,
                                int64_t c, const float* add, int64_t ld_add, float* dx, int64_t ld_dx, void* dx2,
                                int64_t ld_dx2, ealdm_stream_t stream) {
  EALDM_REQUIRE(dup && dx && n > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0 && ld_dup % 4 == 0 && ld_dx % 4 == 0 &&
                    ld_add % 4 == 0 && ld_dx2 % 4 == 0,
                "sumpool2x2: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = bwd::grid_for(n * h * w * (c / 4));
  if (dtype == EALDM_F32)
    bwd::sumpool2x2_kernel<float, float><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const float*>(dup), ld_dup, (int)n,
                                                                   (int)h, (int)w, (int)c, add, ld_add, dx, ld_dx,
                                                                   reinterpret_cast<float*>(dx2), ld_dx2);
  else
    bwd::sumpool2x2_kernel<bf16, float><<<grid, bwd::NT, 0, st>>>(reinterpret_cast<const bf16*>(dup), ld_dup, (int)n,
                                                                  (int)h, (int)w, (int)c, add, ld_add, dx, ld_dx,
                                                                  reinterpret_cast<bf16*>(dx2), ld_dx2);
  EALDM_LAUNCH_CHECK();
  return 0;
}
