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
  return 4 * (2 * a->batch * a->heads * a->n_q + 2LL * chunks * a->batch * a->n_kv * a->heads * a->head_dim);
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
  EALDM_REQUIRE(a->impl != EALDM_IMPL_TCGEN05,
                "attention_bwd: the tensor-core path needs bf16, head_dim 32, lse, n_q and n_kv multiples of 64");
  if (a->dtype == EALDM_F32)
    return a->head_dim == 32 ? attn_bwd::run<float, 32>(a, st) : attn_bwd::run<float, 64>(a, st);
  EALDM_REQUIRE(a->dtype == EALDM_BF16, "attention_bwd: bad dtype");
  return a->head_dim == 32 ? attn_bwd::run<bf16, 32>(a, st) : attn_bwd::run<bf16, 64>(a, st);
}
