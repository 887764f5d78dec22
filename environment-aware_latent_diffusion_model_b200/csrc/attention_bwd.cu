// Attention backward (softmax(scale * q k^T) v), generic FP32-accumulate SIMT kernels for any key count
// (self-attention and the 4-key cross-attention on the STDiff conditioning) at head_dim 32 or 64.
//
//   P = softmax(S), S = scale * Q K^T;  D_i = dO_i . O_i
//   dV = P^T dO;  dP = dO V^T;  dS = P * (dP - D);  dQ = scale * dS K;  dK = scale * dS^T Q
//
// Two passes, no atomics: a query-stationary kernel (one thread per query row: log-sum-exp, D, dQ) and a
// key-stationary kernel (one thread per (key row, query chunk): partial dK / dV), then a fixed-order sum
// of the query-chunk partials.  Threads of a warp share (batch, head), so K/V (resp. Q/dO) row loads are
// warp-uniform broadcasts.
#include "common.cuh"

namespace ealdm {
namespace attn_bwd {

constexpr int NT = 128;

template <typename T, int HD>
__device__ __forceinline__ void load_row(const T* p, float (&f)[HD]) {
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    Vec4<T> q;
    q.load(p + d);
    float t[4];
    q.get(t);
    f[d] = t[0]; f[d + 1] = t[1]; f[d + 2] = t[2]; f[d + 3] = t[3];
  }
}
template <typename T, int HD>
__device__ __forceinline__ void store_row(T* p, const float (&f)[HD]) {
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    Vec4<T> q;
    const float t[4] = {f[d], f[d + 1], f[d + 2], f[d + 3]};
    q.set(t);
    q.store(p + d);
  }
}

template <typename T, int HD>
__global__ void __launch_bounds__(NT)
dq_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ o,
          const T* __restrict__ dout, long long ld_q, long long ld_kv, long long hs_q, long long hs_kv,
          long long ld_o, long long ld_do, int n_q, int n_kv, float scale, T* __restrict__ dq, long long ld_dq,
          long long hs_dq, float* __restrict__ lse, float* __restrict__ dsum) {
  const int i = blockIdx.x * NT + threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  if (i >= n_q) return;
  const long long rq = static_cast<long long>(b) * n_q + i;
  float qf[HD], df[HD], acc[HD];
  load_row<T, HD>(q + rq * ld_q + h * hs_q, qf);
  load_row<T, HD>(dout + rq * ld_do + h * HD, df);
  float D = 0.f;
  {
    float of[HD];
    load_row<T, HD>(o + rq * ld_o + h * HD, of);
#pragma unroll
    for (int d = 0; d < HD; ++d) D = fmaf(df[d], of[d], D);
  }
  const T* kb = k + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  const T* vb = v + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < n_kv; ++j) {
    float kf[HD];
    load_row<T, HD>(kb + static_cast<long long>(j) * ld_kv, kf);
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s = fmaf(qf[d], kf[d], s);
    s *= scale;
    const float mn = fmaxf(m, s);
    l = l * expf(m - mn) + expf(s - mn);
    m = mn;
  }
  const float L = m + logf(l);
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  for (int j = 0; j < n_kv; ++j) {
    float kf[HD], vf[HD];
    load_row<T, HD>(kb + static_cast<long long>(j) * ld_kv, kf);
    load_row<T, HD>(vb + static_cast<long long>(j) * ld_kv, vf);
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      s = fmaf(qf[d], kf[d], s);
      dp = fmaf(df[d], vf[d], dp);
    }
    const float p = expf(s * scale - L);
    const float ds = p * (dp - D) * scale;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = fmaf(ds, kf[d], acc[d]);
  }
  store_row<T, HD>(dq + rq * ld_dq + h * hs_dq, acc);
  const long long si = (static_cast<long long>(b) * gridDim.y + h) * n_q + i;
  lse[si] = L;
  dsum[si] = D;
}

template <typename T, int HD>
__global__ void __launch_bounds__(NT)
dkv_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ dout,
           long long ld_q, long long ld_kv, long long hs_q, long long hs_kv, long long ld_do, int n_q, int n_kv,
           int heads, int chunks, int q_per_chunk, float scale, const float* __restrict__ lse,
           const float* __restrict__ dsum, float* __restrict__ part_k, float* __restrict__ part_v) {
  const int idx = blockIdx.x * NT + threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  if (idx >= n_kv * chunks) return;
  const int j = idx % n_kv, ch = idx / n_kv;
  const long long rk = static_cast<long long>(b) * n_kv + j;
  float kf[HD], vf[HD], dk[HD], dv[HD];
  load_row<T, HD>(k + rk * ld_kv + h * hs_kv, kf);
  load_row<T, HD>(v + rk * ld_kv + h * hs_kv, vf);
#pragma unroll
  for (int d = 0; d < HD; ++d) dk[d] = dv[d] = 0.f;
  const int i0 = ch * q_per_chunk, i1 = min(n_q, i0 + q_per_chunk);
  const float* Lb = lse + (static_cast<long long>(b) * heads + h) * n_q;
  const float* Db = dsum + (static_cast<long long>(b) * heads + h) * n_q;
  for (int i = i0; i < i1; ++i) {
    const long long rq = static_cast<long long>(b) * n_q + i;
    float qf[HD], df[HD];
    load_row<T, HD>(q + rq * ld_q + h * hs_q, qf);
    load_row<T, HD>(dout + rq * ld_do + h * HD, df);
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      s = fmaf(qf[d], kf[d], s);
      dp = fmaf(df[d], vf[d], dp);
    }
    const float p = expf(s * scale - Lb[i]);
    const float ds = p * (dp - Db[i]) * scale;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      dv[d] = fmaf(p, df[d], dv[d]);
      dk[d] = fmaf(ds, qf[d], dk[d]);
    }
  }
  const long long o = ((static_cast<long long>(ch) * gridDim.z + b) * n_kv + j) * (heads * HD) + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    *reinterpret_cast<float4*>(part_k + o + d) = make_float4(dk[d], dk[d + 1], dk[d + 2], dk[d + 3]);
    *reinterpret_cast<float4*>(part_v + o + d) = make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]);
  }
}

// dk[(b*n_kv+j)*ld + h*hs + d] = sum over query chunks of the partials (fixed order)
template <typename T>
__global__ void __launch_bounds__(256)
dkv_finish_kernel(const float* __restrict__ part_k, const float* __restrict__ part_v, int chunks, long long rows,
                  int heads, int hd, T* __restrict__ dk, T* __restrict__ dv, long long ld_dkv, long long hs_dkv) {
  const long long cols = static_cast<long long>(heads) * hd;
  const long long total = rows * cols;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const int h = c / hd, d = c - h * hd;
    float sk = 0.f, sv = 0.f;
    for (int ch = 0; ch < chunks; ++ch) {
      sk += part_k[ch * total + i];
      sv += part_v[ch * total + i];
    }
    dk[r * ld_dkv + h * hs_dkv + d] = from_f32<T>(sk);
    dv[r * ld_dkv + h * hs_dkv + d] = from_f32<T>(sv);
  }
}

// ---- cross-attention on a handful of context tokens (n_kv <= 4, head_dim 32, contiguous heads) -----------------
// Same thread mapping as the forward xattn_bf16_kernel: a thread owns one 16-byte chunk (8 channels) of a token row,
// the 4 lanes of a (token, head) combine their dot products with two shuffles, K / V chunks live in registers.
// dQ is written straight out; dK / dV are accumulated over the CTA's rows in registers, folded across the CTA's
// row slots through shared memory in a fixed order, and written as one partial per CTA (no atomics).
constexpr int XB_THREADS = 256;
constexpr int XB_ROWS = 8;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2b(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int NKV>
__global__ void __launch_bounds__(XB_THREADS)
xattn_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                 const bf16* __restrict__ dout, long long ld_q, long long ld_kv, long long ld_do, int c, int n_q,
                 float scale, bf16* __restrict__ dq, long long ld_dq, float* __restrict__ part) {
  extern __shared__ float red[];                 // [rstep][2 * NKV * 8] per chunk column, see below
  const int b = blockIdx.y;
  const int cpr = c >> 3;
  const int cc = threadIdx.x % cpr;
  const int rs = threadIdx.x / cpr;
  const int rstep = XB_THREADS / cpr;
  const int row0 = blockIdx.x * (rstep * XB_ROWS) + rs;
  const float sl2 = scale * 1.4426950408889634f;
  float kf[NKV][8], vf[NKV][8];
  {
    const bf16* kb = k + static_cast<long long>(b) * NKV * ld_kv + cc * 8;
    const bf16* vb = v + static_cast<long long>(b) * NKV * ld_kv + cc * 8;
#pragma unroll
    for (int j = 0; j < NKV; ++j) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(kb + static_cast<long long>(j) * ld_kv)), kf[j]);
      unpack8(__ldg(reinterpret_cast<const uint4*>(vb + static_cast<long long>(j) * ld_kv)), vf[j]);
    }
  }
  float dk[NKV][8], dv[NKV][8];
#pragma unroll
  for (int j = 0; j < NKV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) dk[j][e] = dv[j][e] = 0.f;
  const bf16* qb = q + static_cast<long long>(b) * n_q * ld_q + cc * 8;
  const bf16* ob = dout + static_cast<long long>(b) * n_q * ld_do + cc * 8;
  bf16* dqb = dq + static_cast<long long>(b) * n_q * ld_dq + cc * 8;
  uint4 qv[XB_ROWS], ov[XB_ROWS];
#pragma unroll
  for (int u = 0; u < XB_ROWS; ++u) {
    const int row = row0 + u * rstep;
    const bool ok = row < n_q;
    qv[u] = ok ? __ldg(reinterpret_cast<const uint4*>(qb + static_cast<long long>(row) * ld_q)) : make_uint4(0, 0, 0, 0);
    ov[u] = ok ? __ldg(reinterpret_cast<const uint4*>(ob + static_cast<long long>(row) * ld_do)) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int u = 0; u < XB_ROWS; ++u) {
    float qf[8], of[8];
    unpack8(qv[u], qf);
    unpack8(ov[u], of);
    float s[NKV], dp[NKV];
#pragma unroll
    for (int j = 0; j < NKV; ++j) {
      float a = 0.f, d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { a = fmaf(qf[e], kf[j][e], a); d = fmaf(of[e], vf[j][e], d); }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      s[j] = a * sl2;
      dp[j] = d;
    }
    float m = s[0];
#pragma unroll
    for (int j = 1; j < NKV; ++j) m = fmaxf(m, s[j]);
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < NKV; ++j) { s[j] = exp2f(s[j] - m); l += s[j]; }
    const float inv = 1.0f / l;
    float D = 0.f;
#pragma unroll
    for (int j = 0; j < NKV; ++j) { s[j] *= inv; D = fmaf(s[j], dp[j], D); }
    float dqf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NKV; ++j) {
      const float ds = s[j] * (dp[j] - D) * scale;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dqf[e] = fmaf(ds, kf[j][e], dqf[e]);
        dk[j][e] = fmaf(ds, qf[e], dk[j][e]);     // rows beyond n_q contribute q = dO = 0
        dv[j][e] = fmaf(s[j], of[e], dv[j][e]);
      }
    }
    const int row = row0 + u * rstep;
    if (row < n_q) {
      uint4 w;
      w.x = pack2b(dqf[0], dqf[1]);
      w.y = pack2b(dqf[2], dqf[3]);
      w.z = pack2b(dqf[4], dqf[5]);
      w.w = pack2b(dqf[6], dqf[7]);
      *reinterpret_cast<uint4*>(dqb + static_cast<long long>(row) * ld_dq) = w;
    }
  }
  // fold the CTA's row slots: red[rs][cc][2*NKV*8]
  float* mine = red + (static_cast<long long>(rs) * cpr + cc) * (2 * NKV * 8);
#pragma unroll
  for (int j = 0; j < NKV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) { mine[j * 8 + e] = dk[j][e]; mine[NKV * 8 + j * 8 + e] = dv[j][e]; }
  __syncthreads();
  // part[(b * gridDim.x + cta)][2][NKV][c]
  float* out = part + (static_cast<long long>(b) * gridDim.x + blockIdx.x) * (2 * NKV * c);
  for (int i = threadIdx.x; i < 2 * NKV * c; i += XB_THREADS) {
    const int which = i / (NKV * c);             // 0: dK, 1: dV
    const int rem = i - which * NKV * c;
    const int j = rem / c, col = rem - j * c;
    const int ccol = col >> 3, e = col & 7;
    float acc = 0.f;
    for (int r = 0; r < rstep; ++r) acc += red[(static_cast<long long>(r) * cpr + ccol) * (2 * NKV * 8) + which * NKV * 8 + j * 8 + e];
    out[i] = acc;
  }
}

// dk[b, j, :] = sum over the CTAs of batch b (fixed order)
__global__ void __launch_bounds__(256)
xattn_bwd_finish_kernel(const float* __restrict__ part, int ctas, int nkv, int c, bf16* __restrict__ dk,
                        bf16* __restrict__ dv, long long ld_dkv) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 2 * nkv * c) return;
  const float* p = part + static_cast<long long>(b) * ctas * (2 * nkv * c) + i;
  float acc = 0.f;
  for (int k = 0; k < ctas; ++k) acc += p[static_cast<long long>(k) * (2 * nkv * c)];
  const int which = i / (nkv * c);
  const int rem = i - which * nkv * c;
  const int j = rem / c, col = rem - j * c;
  bf16* dst = (which == 0 ? dk : dv) + (static_cast<long long>(b) * nkv + j) * ld_dkv + col;
  *dst = __float2bfloat16_rn(acc);
}

static bool xattn_bwd_ok(const ealdm_attention_bwd_args* a) {
  const long long cpr = a->heads * 4;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return a->dtype == EALDM_BF16 && a->head_dim == 32 && a->n_kv <= 4 && a->impl != EALDM_IMPL_SIMT &&
         a->head_stride_q == 32 && a->head_stride_kv == 32 && a->head_stride_dq == 32 && a->head_stride_dkv == 32 &&
         a->ld_q % 8 == 0 && a->ld_kv % 8 == 0 && a->ld_dout % 8 == 0 && a->ld_dq % 8 == 0 && al16(a->q) &&
         al16(a->k) && al16(a->v) && al16(a->dout) && al16(a->dq) && cpr <= XB_THREADS && XB_THREADS % cpr == 0;
}

static int xattn_bwd_ctas(const ealdm_attention_bwd_args* a) {
  const int c = static_cast<int>(a->heads * 32);
  return static_cast<int>(ceil_div(a->n_q, XB_THREADS / (c / 8) * XB_ROWS));
}

static int launch_xattn_bwd(const ealdm_attention_bwd_args* a, cudaStream_t st) {
  const int c = static_cast<int>(a->heads * 32);
  const int ctas = xattn_bwd_ctas(a);
  const int nkv = static_cast<int>(a->n_kv);
  float* part = reinterpret_cast<float*>(a->workspace);
  dim3 grid(static_cast<unsigned>(ctas), static_cast<unsigned>(a->batch));
  const size_t smem = static_cast<size_t>(XB_THREADS) * 2 * nkv * 8 * sizeof(float);
  const bf16* q = reinterpret_cast<const bf16*>(a->q);
  const bf16* k = reinterpret_cast<const bf16*>(a->k);
  const bf16* v = reinterpret_cast<const bf16*>(a->v);
  const bf16* d_o = reinterpret_cast<const bf16*>(a->dout);
  bf16* dq = reinterpret_cast<bf16*>(a->dq);
#define EALDM_XB(NKV)                                                                                              \
  case NKV: {                                                                                                      \
    static DeviceOnce attr;                                                                                        \
    if (attr.pending()) {                                                                                          \
      EALDM_CUDA(cudaFuncSetAttribute(xattn_bwd_kernel<NKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)); \
      attr.done();                                                                                                 \
    }                                                                                                              \
    xattn_bwd_kernel<NKV><<<grid, XB_THREADS, smem, st>>>(q, k, v, d_o, a->ld_q, a->ld_kv, a->ld_dout, c,        \
                                                           (int)a->n_q, a->scale, dq, a->ld_dq, part);            \
  } break
  switch (nkv) {
    EALDM_XB(1); EALDM_XB(2); EALDM_XB(3); EALDM_XB(4);
    default: return set_error(EALDM_EINVAL, "xattn_bwd: n_kv %d > 4", nkv);
  }
#undef EALDM_XB
  EALDM_LAUNCH_CHECK();
  dim3 g2(static_cast<unsigned>(ceil_div(2LL * nkv * c, 256)), static_cast<unsigned>(a->batch));
  xattn_bwd_finish_kernel<<<g2, 256, 0, st>>>(part, ctas, nkv, c, reinterpret_cast<bf16*>(a->dk),
                                              reinterpret_cast<bf16*>(a->dv), a->ld_dkv);
  EALDM_LAUNCH_CHECK();
  return 0;
}

static int plan_chunks(const ealdm_attention_bwd_args* a, int* q_per_chunk) {
  const long long base = a->batch * a->heads * a->n_kv;
  long long ch = ceil_div(148LL * 512, base);
  const long long max_ch = a->n_q / 16 > 0 ? a->n_q / 16 : 1;
  if (ch > max_ch) ch = max_ch;
  if (ch < 1) ch = 1;
  *q_per_chunk = static_cast<int>(ceil_div(a->n_q, ch));
  return static_cast<int>(ceil_div(a->n_q, *q_per_chunk));
}

template <typename T, int HD>
static int run(const ealdm_attention_bwd_args* a, cudaStream_t st) {
  int qpc;
  const int chunks = plan_chunks(a, &qpc);
  float* lse = reinterpret_cast<float*>(a->workspace);
  float* dsum = lse + a->batch * a->heads * a->n_q;
  float* part_k = dsum + a->batch * a->heads * a->n_q;
  float* part_v = part_k + static_cast<long long>(chunks) * a->batch * a->n_kv * a->heads * HD;
  const T* q = reinterpret_cast<const T*>(a->q);
  const T* k = reinterpret_cast<const T*>(a->k);
  const T* v = reinterpret_cast<const T*>(a->v);
  const T* o = reinterpret_cast<const T*>(a->out);
  const T* d_o = reinterpret_cast<const T*>(a->dout);
  dim3 g1(static_cast<unsigned>(ceil_div(a->n_q, NT)), static_cast<unsigned>(a->heads), static_cast<unsigned>(a->batch));
  dq_kernel<T, HD><<<g1, NT, 0, st>>>(q, k, v, o, d_o, a->ld_q, a->ld_kv, a->head_stride_q, a->head_stride_kv, a->ld_out,
                                      a->ld_dout, (int)a->n_q, (int)a->n_kv, a->scale, reinterpret_cast<T*>(a->dq),
                                      a->ld_dq, a->head_stride_dq, lse, dsum);
  EALDM_LAUNCH_CHECK();
  dim3 g2(static_cast<unsigned>(ceil_div(a->n_kv * chunks, NT)), static_cast<unsigned>(a->heads),
          static_cast<unsigned>(a->batch));
  dkv_kernel<T, HD><<<g2, NT, 0, st>>>(q, k, v, d_o, a->ld_q, a->ld_kv, a->head_stride_q, a->head_stride_kv, a->ld_dout,
                                       (int)a->n_q, (int)a->n_kv, (int)a->heads, chunks, qpc, a->scale, lse, dsum, part_k,
                                       part_v);
  EALDM_LAUNCH_CHECK();
  const long long rows = a->batch * a->n_kv;
  const long long total = rows * a->heads * HD;
  const int blocks = static_cast<int>(ceil_div(total, 256) < 4096 ? ceil_div(total, 256) : 4096);
  dkv_finish_kernel<T><<<blocks, 256, 0, st>>>(part_k, part_v, chunks, rows, (int)a->heads, HD,
                                               reinterpret_cast<T*>(a->dk), reinterpret_cast<T*>(a->dv), a->ld_dkv,
                                               a->head_stride_dkv);
  EALDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace attn_bwd
}  // namespace ealdm

namespace ealdm {
namespace attn {
bool bwd_mma_ok(const ealdm_attention_bwd_args* a);                   // attention_bwd_mma.cu
int launch_bwd_mma(const ealdm_attention_bwd_args* a, cudaStream_t st);
}  // namespace attn
}  // namespace ealdm

using namespace ealdm;

extern "C" int64_t ealdm_attention_bwd_workspace_bytes(const ealdm_attention_bwd_args* a) {
  if (!a || a->batch <= 0 || a->heads <= 0 || a->n_q <= 0 || a->n_kv <= 0 || a->head_dim <= 0) return -1;
  int qpc;
  const int chunks = attn_bwd::plan_chunks(a, &qpc);
  const int64_t generic = 4 * (2 * a->batch * a->heads * a->n_q + 2LL * chunks * a->batch * a->n_kv * a->heads * a->head_dim);
  if (attn_bwd::xattn_bwd_ok(a)) {
    const int64_t x = 4LL * a->batch * attn_bwd::xattn_bwd_ctas(a) * 2 * a->n_kv * a->heads * 32;
    return x > generic ? x : generic;
  }
  return generic;
}

extern "C" int ealdm_attention_bwd(const ealdm_attention_bwd_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->q && a->k && a->v && a->out && a->dout && a->dq && a->dk && a->dv && a->workspace,
                "attention_bwd: null argument");
  EALDM_REQUIRE(a->batch > 0 && a->batch <= 65535 && a->heads > 0 && a->heads <= 65535 && a->n_q > 0 && a->n_kv > 0,
                "attention_bwd: bad sizes");
  EALDM_REQUIRE(a->head_dim == 32 || a->head_dim == 64, "attention_bwd: head_dim must be 32 or 64");
  EALDM_REQUIRE(a->ld_q % 4 == 0 && a->ld_kv % 4 == 0 && a->ld_out % 4 == 0 && a->ld_dout % 4 == 0 &&
                    a->ld_dq % 4 == 0 && a->ld_dkv % 4 == 0 && a->head_stride_q % 4 == 0 &&
                    a->head_stride_kv % 4 == 0 && a->head_stride_dq % 4 == 0 && a->head_stride_dkv % 4 == 0,
                "attention_bwd: pitches and head strides must be multiples of 4");
  EALDM_REQUIRE(a->workspace_bytes >= ealdm_attention_bwd_workspace_bytes(a), "attention_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (attn::bwd_mma_ok(a)) return attn::launch_bwd_mma(a, st);
  if (attn_bwd::xattn_bwd_ok(a)) return attn_bwd::launch_xattn_bwd(a, st);
  EALDM_REQUIRE(a->impl != EALDM_IMPL_TCGEN05,
                "attention_bwd: the tensor-core path needs bf16, head_dim 32, lse, n_q and n_kv multiples of 64");
  if (a->dtype == EALDM_F32)
    return a->head_dim == 32 ? attn_bwd::run<float, 32>(a, st) : attn_bwd::run<float, 64>(a, st);
  EALDM_REQUIRE(a->dtype == EALDM_BF16, "attention_bwd: bad dtype");
  return a->head_dim == 32 ? attn_bwd::run<bf16, 32>(a, st) : attn_bwd::run<bf16, 64>(a, st);
}
