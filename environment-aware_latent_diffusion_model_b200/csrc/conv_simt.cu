// Generic implicit-GEMM convolution / linear on the FP32 FFMA pipes.
//
// This is the parity-mode (EALDM_F32) implementation and the catch-all for layers the tcgen05
// kernel does not take (4- and 3-channel inputs, fused nearest-2x upsampling): any kernel size in
// {1,3}, stride in {1,2}, asymmetric padding, up to two K-concatenated sources, the same epilogue.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles, fp32 accumulation in summation order
// k = (source, kh, kw, c).
#include "common.cuh"

namespace ealdm {
namespace simt {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct Src {
  const void* x;
  long long ld;
  int n, h, w, c;
  int ksize, stride, pad, upsample;
  int k_len;  // ksize*ksize*c
};

struct Params {
  int nsrc;
  Src src[2];
  const void* weight;
  long long k_total;
  int N;      // weight rows
  long long M;
  int Hout, Wout;
  int vec4;   // all channel counts and k_total are multiples of 4 and pointers are 16 B aligned
  Epilogue ep;
};

template <typename T>
__device__ __forceinline__ float load_a(const Params& p, int n, int oh, int ow, long long k) {
  int s = 0;
  if (p.nsrc > 1 && k >= p.src[0].k_len) { s = 1; k -= p.src[0].k_len; }
  const Src& x = p.src[s];
  const int tap = static_cast<int>(k / x.c);
  const int c = static_cast<int>(k - static_cast<long long>(tap) * x.c);
  const int kh = tap / x.ksize, kw = tap - kh * x.ksize;
  int ih = oh * x.stride + kh - x.pad;
  int iw = ow * x.stride + kw - x.pad;
  const int hh = x.upsample ? x.h * 2 : x.h, ww = x.upsample ? x.w * 2 : x.w;
  if (ih < 0 || ih >= hh || iw < 0 || iw >= ww) return 0.0f;
  if (x.upsample) { ih >>= 1; iw >>= 1; }
  const T* px = reinterpret_cast<const T*>(x.x) +
                ((static_cast<long long>(n) * x.h + ih) * x.w + iw) * x.ld + c;
  return to_f32(*px);
}

template <typename T>
__device__ __forceinline__ void load_a4(const Params& p, int n, int oh, int ow, long long k,
                                        float (&f)[4]) {
  int s = 0;
  if (p.nsrc > 1 && k >= p.src[0].k_len) { s = 1; k -= p.src[0].k_len; }
  const Src& x = p.src[s];
  const int tap = static_cast<int>(k / x.c);
  const int c = static_cast<int>(k - static_cast<long long>(tap) * x.c);
  const int kh = tap / x.ksize, kw = tap - kh * x.ksize;
  int ih = oh * x.stride + kh - x.pad;
  int iw = ow * x.stride + kw - x.pad;
  const int hh = x.upsample ? x.h * 2 : x.h, ww = x.upsample ? x.w * 2 : x.w;
  if (ih < 0 || ih >= hh || iw < 0 || iw >= ww) {
    f[0] = f[1] = f[2] = f[3] = 0.0f;
    return;
  }
  if (x.upsample) { ih >>= 1; iw >>= 1; }
  const T* px = reinterpret_cast<const T*>(x.x) +
                ((static_cast<long long>(n) * x.h + ih) * x.w + iw) * x.ld + c;
  Vec4<T> v;
  v.load(px);
  v.get(f);
}

template <typename T, typename TOut>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const Params p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int t = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * BM;
  const int n0 = blockIdx.y * BN;

  // loader mapping: 4 consecutive k of one row
  const int lrow = t >> 2;
  const int lk = (t & 3) * 4;
  const long long am = m0 + lrow;
  const bool a_valid = am < p.M;
  int an = 0, aoh = 0, aow = 0;
  if (a_valid) {
    const long long hw = static_cast<long long>(p.Hout) * p.Wout;
    an = static_cast<int>(am / hw);
    const int rem = static_cast<int>(am - an * hw);
    aoh = rem / p.Wout;
    aow = rem - aoh * p.Wout;
  }
  const int bn_row = n0 + lrow;
  const bool b_valid = bn_row < p.N;
  const T* wrow = reinterpret_cast<const T*>(p.weight) + static_cast<long long>(bn_row) * p.k_total;

  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (long long k0 = 0; k0 < p.k_total; k0 += BK) {
    float fa[4] = {0.f, 0.f, 0.f, 0.f}, fb[4] = {0.f, 0.f, 0.f, 0.f};
    const long long k = k0 + lk;
    if (p.vec4) {
      if (k < p.k_total) {
        if (a_valid) load_a4<T>(p, an, aoh, aow, k, fa);
        if (b_valid) {
          Vec4<T> v;
          v.load(wrow + k);
          v.get(fb);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k + j < p.k_total) {
          if (a_valid) fa[j] = load_a<T>(p, an, aoh, aow, k + j);
          if (b_valid) fb[j] = to_f32(wrow[k + j]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk + j][lrow] = fa[j];
      Bs[lk + j][lrow] = fb[j];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  const Epilogue& ep = p.ep;
  if (ep.act == EALDM_ACT_GEGLU) {
    // stage acc + bias through shared memory so value/gate pairs (16 columns apart) meet
    __shared__ float Cs[BM][BN + 1];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx * 4 + j;
        float v = acc[i][j];
        if (ep.bias && col < p.N) v += ep.bias[col];
        Cs[ty * 4 + i][tx * 4 + j] = v;
      }
    __syncthreads();
    for (int idx = t; idx < BM * (BN / 2); idx += NT) {
      const int rr = idx / (BN / 2), oc = idx % (BN / 2);
      const int blk = oc >> 4, j = oc & 15;
      const long long m = m0 + rr;
      const int col = n0 + blk * 32 + j;  // accumulator column of the value
      if (m < p.M && col < p.N) {
        float v = Cs[rr][blk * 32 + j] * gelu_erf_f(Cs[rr][blk * 32 + 16 + j]);
        const long long ocol = (n0 >> 1) + blk * 16 + j;
        if (ep.residual)
          v += ep.res_f32 ? reinterpret_cast<const float*>(ep.residual)[m * ep.ld_res + ocol]
                          : to_f32(reinterpret_cast<const T*>(ep.residual)[m * ep.ld_res + ocol]);
        reinterpret_cast<TOut*>(ep.out)[m * ep.ld_out + ocol] = from_f32<TOut>(v);
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const long long img = m / ep.rows_per_image;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= p.N) continue;
      float v = acc[i][j];
      if (ep.bias) v += ep.bias[col];
      if (ep.rowvec) v += ep.rowvec[img * ep.ld_rowvec + col];
      if (ep.act == EALDM_ACT_SILU) v = silu_f(v);
      else if (ep.act == EALDM_ACT_RELU) v = fmaxf(v, 0.f);
      if (ep.residual)
        v += ep.res_f32 ? reinterpret_cast<const float*>(ep.residual)[m * ep.ld_res + col]
                        : to_f32(reinterpret_cast<const T*>(ep.residual)[m * ep.ld_res + col]);
      reinterpret_cast<TOut*>(ep.out)[m * ep.ld_out + col] = from_f32<TOut>(v);
      if (ep.out2) reinterpret_cast<T*>(ep.out2)[m * ep.ld_out2 + col] = from_f32<T>(v);
    }
  }
}

int launch(const ealdm_conv_args* a, cudaStream_t st) {
  Params p;
  memset(&p, 0, sizeof(p));
  p.nsrc = a->n_src;
  const int esize = a->dtype == EALDM_BF16 ? 2 : 4;
  bool vec4 = (a->k_total % 4 == 0) && ((reinterpret_cast<uintptr_t>(a->weight) % (4 * esize)) == 0);
  long long ksum = 0;
  for (int s = 0; s < a->n_src; ++s) {
    const ealdm_conv_src& x = a->src[s];
    Src& d = p.src[s];
    d.x = x.x; d.ld = x.ld;
    d.n = (int)x.n; d.h = (int)x.h; d.w = (int)x.w; d.c = (int)x.c;
    d.ksize = x.ksize; d.stride = x.stride; d.pad = x.pad; d.upsample = x.upsample;
    d.k_len = x.ksize * x.ksize * (int)x.c;
    ksum += d.k_len;
    if (x.c % 4 != 0 || x.ld % 4 != 0 || (reinterpret_cast<uintptr_t>(x.x) % (4 * esize)) != 0)
      vec4 = false;
  }
  EALDM_REQUIRE(ksum == a->k_total, "k_total %lld does not match the sources (%lld)",
                (long long)a->k_total, ksum);
  p.weight = a->weight;
  p.k_total = a->k_total;
  p.N = (int)a->n_out;
  p.Hout = (int)a->h_out;
  p.Wout = (int)a->w_out;
  p.M = a->src[0].n * a->h_out * a->w_out;
  p.vec4 = vec4 ? 1 : 0;
  p.ep.bias = a->bias;
  p.ep.rowvec = a->rowvec;
  p.ep.ld_rowvec = a->ld_rowvec;
  p.ep.rows_per_image = a->h_out * a->w_out;
  p.ep.residual = a->residual;
  p.ep.ld_res = a->ld_res;
  p.ep.out = a->out;
  p.ep.ld_out = a->ld_out;
  p.ep.act = a->act;
  p.ep.out_f32 = a->out_f32;
  p.ep.res_f32 = a->res_f32;
  p.ep.out2 = a->out2;
  p.ep.ld_out2 = a->ld_out2;
  if (a->act == EALDM_ACT_GEGLU)
    EALDM_REQUIRE(a->n_out % 32 == 0, "GEGLU needs n_out %% 32 == 0 (got %lld)", (long long)a->n_out);

  dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(p.N, BN));
  EALDM_REQUIRE(grid.y <= 65535, "n_out too large");
  if (a->dtype == EALDM_F32) {
    conv_simt_kernel<float, float><<<grid, NT, 0, st>>>(p);
  } else if (a->out_f32) {
    conv_simt_kernel<bf16, float><<<grid, NT, 0, st>>>(p);
  } else {
    conv_simt_kernel<bf16, bf16><<<grid, NT, 0, st>>>(p);
  }
  EALDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace simt
}  // namespace ealdm
