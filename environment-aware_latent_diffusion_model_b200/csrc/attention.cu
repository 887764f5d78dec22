// Attention for the EALDM UNet.
//   * flash_simt_kernel  : online-softmax attention, one query per thread, K/V tiles in shared
//                          memory, fp32 math (parity mode and reference for the tensor-core kernel);
//   * flash_mma_kernel   : bf16 tensor-core flash attention for head_dim 32 (attention_mma.cu);
//   * smallkv_kernel     : cross-attention on the 4-token environment embedding (n_kv <= 16):
//                          one thread per (query, head), K/V served from L1/L2.
#include "common.cuh"

#include <stdlib.h>

namespace ealdm {
namespace attn {

int launch_flash_mma(const ealdm_attention_args* a, cudaStream_t st);  // attention_mma.cu

constexpr int QT = 128;  // queries per CTA (one per thread)
constexpr int KT = 32;   // keys per shared-memory tile

template <typename T, int D>
__global__ void __launch_bounds__(QT)
flash_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                  long long ld_q, long long ld_kv, long long hs_q, long long hs_kv, int n_q,
                  int n_kv, float scale, T* __restrict__ out, long long ld_out) {
  __shared__ __align__(16) float Ks[KT][D];
  __shared__ __align__(16) float Vs[KT][D];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * QT + threadIdx.x;
  const bool valid = qi < n_q;
  float qr[D], o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { qr[d] = 0.f; o[d] = 0.f; }
  if (valid) {
    const T* qp = q + (static_cast<long long>(b) * n_q + qi) * ld_q + h * hs_q;
#pragma unroll
    for (int d = 0; d < D; d += 4) {
      Vec4<T> t4;
      t4.load(qp + d);
      float f[4];
      t4.get(f);
#pragma unroll
      for (int j = 0; j < 4; ++j) qr[d + j] = f[j] * scale;
    }
  }
  float m = -INFINITY, l = 0.f;
  const T* kb = k + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  const T* vb = v + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  for (int j0 = 0; j0 < n_kv; j0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * (D / 4); i += QT) {
      const int r = i / (D / 4), c4 = (i % (D / 4)) * 4;
      float fk[4] = {0.f, 0.f, 0.f, 0.f}, fv[4] = {0.f, 0.f, 0.f, 0.f};
      if (j0 + r < n_kv) {
        Vec4<T> t4;
        t4.load(kb + static_cast<long long>(j0 + r) * ld_kv + c4);
        t4.get(fk);
        t4.load(vb + static_cast<long long>(j0 + r) * ld_kv + c4);
        t4.get(fv);
      }
      *reinterpret_cast<float4*>(&Ks[r][c4]) = make_float4(fk[0], fk[1], fk[2], fk[3]);
      *reinterpret_cast<float4*>(&Vs[r][c4]) = make_float4(fv[0], fv[1], fv[2], fv[3]);
    }
    __syncthreads();
    const int cnt = min(KT, n_kv - j0);
    float s[KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[j][d]);
        acc = fmaf(qr[d], kk.x, acc);
        acc = fmaf(qr[d + 1], kk.y, acc);
        acc = fmaf(qr[d + 2], kk.z, acc);
        acc = fmaf(qr[d + 3], kk.w, acc);
      }
      s[j] = (j < cnt) ? acc : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float m_new = fmaxf(m, tmax);
    const float corr = expf(m - m_new);
    l *= corr;
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float p = expf(s[j] - m_new);
      l += p;
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d]);
        o[d] = fmaf(p, vv.x, o[d]);
        o[d + 1] = fmaf(p, vv.y, o[d + 1]);
        o[d + 2] = fmaf(p, vv.z, o[d + 2]);
        o[d + 3] = fmaf(p, vv.w, o[d + 3]);
      }
    }
    m = m_new;
  }
  if (valid) {
    const float inv = 1.0f / l;
    T* op = out + (static_cast<long long>(b) * n_q + qi) * ld_out + h * D;
#pragma unroll
    for (int d = 0; d < D; d += 4) {
      float f[4] = {o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv};
      Vec4<T> t4;
      t4.set(f);
      t4.store(op + d);
    }
  }
}

constexpr int MAX_SMALL_KV = 16;

template <typename T, int D>
__global__ void __launch_bounds__(256)
smallkv_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
               long long ld_q, long long ld_kv, long long hs_q, long long hs_kv, long long batch,
               int heads, int n_q, int n_kv, float scale, T* __restrict__ out, long long ld_out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = batch * n_q * heads;
  if (idx >= total) return;
  const int h = static_cast<int>(idx % heads);
  const long long row = idx / heads;  // b*n_q + i
  const long long b = row / n_q;
  float qr[D];
  const T* qp = q + row * ld_q + h * hs_q;
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    Vec4<T> t4;
    t4.load(qp + d);
    float f[4];
    t4.get(f);
#pragma unroll
    for (int j = 0; j < 4; ++j) qr[d + j] = f[j];
  }
  float s[MAX_SMALL_KV];
  float m = -INFINITY;
  const T* kb = k + b * n_kv * ld_kv + h * hs_kv;
  const T* vb = v + b * n_kv * ld_kv + h * hs_kv;
#pragma unroll
  for (int j = 0; j < MAX_SMALL_KV; ++j) {
    if (j < n_kv) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        Vec4<T> t4;
        t4.load(kb + static_cast<long long>(j) * ld_kv + d);
        float f[4];
        t4.get(f);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc = fmaf(qr[d + e], f[e], acc);
      }
      s[j] = acc * scale;
      m = fmaxf(m, s[j]);
    } else {
      s[j] = -INFINITY;
    }
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < MAX_SMALL_KV; ++j) {
    s[j] = (j < n_kv) ? expf(s[j] - m) : 0.f;
    l += s[j];
  }
  const float inv = 1.0f / l;
  float o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < MAX_SMALL_KV; ++j) {
    if (j < n_kv) {
      const float p = s[j] * inv;
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        Vec4<T> t4;
        t4.load(vb + static_cast<long long>(j) * ld_kv + d);
        float f[4];
        t4.get(f);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[d + e] = fmaf(p, f[e], o[d + e]);
      }
    }
  }
  T* op = out + row * ld_out + h * D;
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    float f[4] = {o[d], o[d + 1], o[d + 2], o[d + 3]};
    Vec4<T> t4;
    t4.set(f);
    t4.store(op + d);
  }
}

// ---- cross-attention on a handful of context tokens, bf16, head_dim 32: HBM-bound --------------------
// Q and the output are streamed exactly once with fully coalesced 16-byte accesses: consecutive threads
// own consecutive 16-byte chunks (8 channels) of the token row, so 4 neighbouring lanes share one
// (token, head) and combine their partial dot products with two shuffles.  A thread keeps the SAME
// chunk position for all the rows it visits, so its slice of K and V (n_kv x 8 channels each) is
// loaded once into registers and reused for every row.
constexpr int XKV_MAX = 4;       // context tokens (register-resident K / V slices)
constexpr int XKV_THREADS = 256;
constexpr int XKV_ROWS = 8;      // rows per thread, all loads in flight before the first use

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}

template <int NKV>
__global__ void __launch_bounds__(XKV_THREADS)
xattn_bf16_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                  long long ld_q, long long ld_kv, int c, int n_q, float scale_log2,
                  bf16* __restrict__ out, long long ld_out) {
  pdl_wait();      // launched with programmatic serialization (common.cuh)
  pdl_trigger();
  const int b = blockIdx.y;
  const int cpr = c >> 3;                      // 16-byte chunks per token row; divides XKV_THREADS
  const int cc = threadIdx.x % cpr;            // this thread's chunk position
  const int rstep = XKV_THREADS / cpr;         // rows covered by one pass of the CTA
  const int row0 = blockIdx.x * (rstep * XKV_ROWS) + threadIdx.x / cpr;
  const bf16* qb = q + static_cast<long long>(b) * n_q * ld_q + cc * 8;
  bf16* ob = out + static_cast<long long>(b) * n_q * ld_out + cc * 8;
  uint4 qv[XKV_ROWS];
#pragma unroll
  for (int u = 0; u < XKV_ROWS; ++u) {
    const int row = row0 + u * rstep;
    qv[u] = row < n_q ? __ldg(reinterpret_cast<const uint4*>(qb + static_cast<long long>(row) * ld_q))
                      : make_uint4(0, 0, 0, 0);
  }
  float kf[NKV][8], vf[NKV][8];
  {
    const bf16* kb = k + static_cast<long long>(b) * NKV * ld_kv + cc * 8;
    const bf16* vb = v + static_cast<long long>(b) * NKV * ld_kv + cc * 8;
#pragma unroll
    for (int j = 0; j < NKV; ++j) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(kb + static_cast<long long>(j) * ld_kv)), kf[j]);
      unpack8(__ldg(reinterpret_cast<const uint4*>(vb + static_cast<long long>(j) * ld_kv)), vf[j]);
#pragma unroll
      for (int e = 0; e < 8; ++e) kf[j][e] *= scale_log2;   // fold softmax scale and log2(e) into K
    }
  }
#pragma unroll
  for (int u = 0; u < XKV_ROWS; ++u) {
    float qf[8];
    unpack8(qv[u], qf);
    float s[NKV];
#pragma unroll
    for (int j = 0; j < NKV; ++j) {
      float acc = qf[0] * kf[j][0];
#pragma unroll
      for (int e = 1; e < 8; ++e) acc = fmaf(qf[e], kf[j][e], acc);
      // the 4 lanes holding the 32 channels of this (token, head)
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      s[j] = acc;
    }
    float m = s[0];
#pragma unroll
    for (int j = 1; j < NKV; ++j) m = fmaxf(m, s[j]);
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < NKV; ++j) { s[j] = exp2f(s[j] - m); l += s[j]; }
    const float inv = __fdividef(1.0f, l);
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NKV; ++j) {
      const float pj = s[j] * inv;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf(pj, vf[j][e], o[e]);
    }
    const int row = row0 + u * rstep;
    if (row < n_q) {
      uint4 w;
      __nv_bfloat162 t;
      t = __floats2bfloat162_rn(o[0], o[1]); w.x = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(o[2], o[3]); w.y = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(o[4], o[5]); w.z = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(o[6], o[7]); w.w = *reinterpret_cast<uint32_t*>(&t);
      *reinterpret_cast<uint4*>(ob + static_cast<long long>(row) * ld_out) = w;
    }
  }
}

// usable when heads are packed (head h = columns [32h, 32h+32)), rows are 16-byte aligned and the
// chunks of a row tile the CTA evenly
static bool xattn_bf16_ok(const ealdm_attention_args* a) {
  const long long cpr = a->heads * 4;
  return a->dtype == EALDM_BF16 && a->head_dim == 32 && a->n_kv <= XKV_MAX && a->head_stride_q == 32 &&
         a->head_stride_kv == 32 && a->ld_q % 8 == 0 && a->ld_kv % 8 == 0 && a->ld_out % 8 == 0 &&
         ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
           reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->out)) & 15) == 0 &&
         cpr <= XKV_THREADS && XKV_THREADS % cpr == 0;
}

static int launch_xattn_bf16(const ealdm_attention_args* a, cudaStream_t st) {
  const int c = static_cast<int>(a->heads * 32);
  const int rows_per_cta = XKV_THREADS / (c / 8) * XKV_ROWS;
  dim3 grid(static_cast<unsigned>(ceil_div(a->n_q, rows_per_cta)), static_cast<unsigned>(a->batch));
  const float sl2 = a->scale * 1.4426950408889634f;
  const bf16* q = reinterpret_cast<const bf16*>(a->q);
  const bf16* k = reinterpret_cast<const bf16*>(a->k);
  const bf16* v = reinterpret_cast<const bf16*>(a->v);
  bf16* o = reinterpret_cast<bf16*>(a->out);
#define EALDM_XATTN(NKV)                                                                           \
  case NKV:                                                                                        \
    EALDM_CUDA(launch_pdl(xattn_bf16_kernel<NKV>, grid, dim3(XKV_THREADS), 0, st, q, k, v, a->ld_q, a->ld_kv, c, \
                          (int)a->n_q, sl2, o, a->ld_out));                                        \
    break
  switch (a->n_kv) {
    EALDM_XATTN(1); EALDM_XATTN(2); EALDM_XATTN(3); EALDM_XATTN(4);
    default: return set_error(EALDM_EINVAL, "xattn: n_kv %lld > %d", (long long)a->n_kv, XKV_MAX);
  }
#undef EALDM_XATTN
  EALDM_LAUNCH_CHECK();
  return 0;
}

template <typename T, int D>
static int launch_t(const ealdm_attention_args* a, cudaStream_t st) {
  const T* q = reinterpret_cast<const T*>(a->q);
  const T* k = reinterpret_cast<const T*>(a->k);
  const T* v = reinterpret_cast<const T*>(a->v);
  T* o = reinterpret_cast<T*>(a->out);
  if (a->n_kv <= MAX_SMALL_KV) {
    const long long total = a->batch * a->n_q * a->heads;
    smallkv_kernel<T, D><<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, st>>>(
        q, k, v, a->ld_q, a->ld_kv, a->head_stride_q, a->head_stride_kv, a->batch, (int)a->heads,
        (int)a->n_q, (int)a->n_kv, a->scale, o, a->ld_out);
  } else {
    dim3 grid(static_cast<unsigned>(ceil_div(a->n_q, QT)), static_cast<unsigned>(a->heads),
              static_cast<unsigned>(a->batch));
    flash_simt_kernel<T, D><<<grid, QT, 0, st>>>(q, k, v, a->ld_q, a->ld_kv, a->head_stride_q,
                                                  a->head_stride_kv, (int)a->n_q, (int)a->n_kv,
                                                  a->scale, o, a->ld_out);
  }
  EALDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace attn
}  // namespace ealdm

namespace ealdm {
namespace attn_tc {
bool supported(const ealdm_attention_args* a);                             // attention_tc.cu
int launch(const ealdm_attention_args* a, cudaStream_t st, bool wide);
}  // namespace attn_tc
}  // namespace ealdm

using namespace ealdm;

extern "C" int ealdm_attention(const ealdm_attention_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->q && a->k && a->v && a->out, "attention: null argument");
  EALDM_REQUIRE(a->batch > 0 && a->heads > 0 && a->n_q > 0 && a->n_kv > 0, "attention: bad sizes");
  EALDM_REQUIRE(a->batch <= 65535 && a->heads <= 65535, "attention: batch/heads too large");
  EALDM_REQUIRE(a->ld_q % 4 == 0 && a->ld_kv % 4 == 0 && a->ld_out % 4 == 0 &&
                    a->head_stride_q % 4 == 0 && a->head_stride_kv % 4 == 0,
                "attention: pitches must be multiples of 4 elements");
  EALDM_REQUIRE(a->head_dim == 32 || a->head_dim == 64,
                "attention: head_dim %lld unsupported (32 or 64)", (long long)a->head_dim);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    // tcgen05 / TMEM flash attention (n_q, n_kv multiples of 128); EALDM_ATTN_TC = off | wide | narrow (tuning)
    static const char* mode = getenv("EALDM_ATTN_TC");
    const bool off = mode != nullptr && mode[0] == 'o';
    const bool wide = mode != nullptr && mode[0] == 'w';
    if (a->impl == EALDM_IMPL_TCGEN05 || (a->impl == EALDM_IMPL_AUTO && !off)) {
      if (attn_tc::supported(a)) return attn_tc::launch(a, st, wide);
      EALDM_REQUIRE(a->impl != EALDM_IMPL_TCGEN05,
                    "attention: the tcgen05 path needs bf16, head_dim 32, n_q and n_kv multiples of 128");
    }
  }
  if (a->dtype == EALDM_BF16 && a->head_dim == 32 && a->n_kv > attn::MAX_SMALL_KV &&
      a->impl != EALDM_IMPL_SIMT)
    return attn::launch_flash_mma(a, st);
  EALDM_REQUIRE(a->lse == nullptr, "attention: lse is only produced by the bf16 tensor-core path (head_dim 32)");
  if (a->impl != EALDM_IMPL_SIMT && attn::xattn_bf16_ok(a)) return attn::launch_xattn_bf16(a, st);
  if (a->dtype == EALDM_F32) {
    return a->head_dim == 32 ? attn::launch_t<float, 32>(a, st) : attn::launch_t<float, 64>(a, st);
  } else if (a->dtype == EALDM_BF16) {
    return a->head_dim == 32 ? attn::launch_t<bf16, 32>(a, st) : attn::launch_t<bf16, 64>(a, st);
  }
  return set_error(EALDM_EINVAL, "attention: bad dtype %d", a->dtype);
}
