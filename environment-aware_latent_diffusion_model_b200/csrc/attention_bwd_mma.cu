// bf16 tensor-core flash-attention BACKWARD for head_dim 32 (UNet self-attention, 1024 / 256 / 64 tokens).
//
//   P = 2^(S*c - lse),  S = Q K^T,  c = scale*log2(e),  D_i = dO_i . O_i
//   dV = P^T dO,  dP = dO V^T,  dS = P * (dP - D),  dQ = scale * dS K,  dK = scale * dS^T Q
//
// Three kernels, no atomics (bit-reproducible):
//   dsum_kernel  D = rowsum(dO * O)                                     (HBM-bound, coalesced 16-byte loads)
//   dkdv_kernel  key-stationary: one warp owns 16 key rows (K and V as register A-fragments) and streams
//                64-query tiles of Q / dO / lse / D through shared memory (cp.async double buffer):
//                S^T = K Q^T, P^T, dV += P^T dO, dP^T = V dO^T, dS^T, dK += dS^T Q  -- 64 mma per tile
//   dq_kernel    query-stationary: one warp owns 16 query rows (Q and dO fragments) and streams 64-key tiles
//                of K / V: S, P, dP = dO V^T, dS, dQ += dS K                   -- 48 mma per tile
// Nothing of size N^2 is ever written; the probabilities are recomputed from the forward's log-sum-exp.
#include "common.cuh"
#include "mma.cuh"

namespace ealdm {
namespace attn {

constexpr int TILE = 64;

// D[b, h, i] = sum_d dO[b, i, h, d] * O[b, i, h, d];  4 lanes x 8 elements per (row, head)
__global__ void __launch_bounds__(256)
dsum_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, long long ld_o, long long ld_do, int heads,
            int n_q, long long total_rows, float* __restrict__ dsum) {
  const long long idx = blockIdx.x * 256LL + threadIdx.x;  // (row, head, quarter)
  const int quarter = static_cast<int>(idx & 3);
  const long long rh = idx >> 2;
  const int h = static_cast<int>(rh % heads);
  const long long row = rh / heads;
  float s = 0.f;
  if (row < total_rows) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * ld_o + h * 32 + quarter * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(dout + row * ld_do + h * 32 + quarter * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s = fmaf(__uint_as_float(aw[j] << 16), __uint_as_float(bw[j] << 16), s);
      s = fmaf(__uint_as_float(aw[j] & 0xffff0000u), __uint_as_float(bw[j] & 0xffff0000u), s);
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (quarter == 0 && row < total_rows) {
    const long long b_ = row / n_q;
    const int i = static_cast<int>(row - b_ * n_q);
    dsum[(b_ * heads + h) * n_q + i] = s;
  }
}

// A fragment (16 rows x 32 d, two k-steps) straight from global memory: rows r0 / r1 = r0 + 8
__device__ __forceinline__ void load_a_frag(const bf16* base, long long ld, int r0, int cq, uint32_t (&a)[2][4]) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int c = ks * 16 + cq;
    a[ks][0] = *reinterpret_cast<const uint32_t*>(base + static_cast<long long>(r0) * ld + c);
    a[ks][1] = *reinterpret_cast<const uint32_t*>(base + static_cast<long long>(r0 + 8) * ld + c);
    a[ks][2] = *reinterpret_cast<const uint32_t*>(base + static_cast<long long>(r0) * ld + c + 8);
    a[ks][3] = *reinterpret_cast<const uint32_t*>(base + static_cast<long long>(r0 + 8) * ld + c + 8);
  }
}

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
dkdv_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
            const bf16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ dsum,
            long long ld_q, long long ld_kv, long long hs_q, long long hs_kv, long long ld_do, int n_q, int n_kv,
            float scale, float scale_log2, bf16* __restrict__ dk, bf16* __restrict__ dv, long long ld_dkv,
            long long hs_dkv) {
  __shared__ __align__(16) bf16 Qs[2][TILE][ROW_PAD];
  __shared__ __align__(16) bf16 Os[2][TILE][ROW_PAD];
  __shared__ __align__(16) float Ls[2][TILE];
  __shared__ __align__(16) float Ds[2][TILE];
  constexpr int NT = NWARPS * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int kv0 = blockIdx.x * (NWARPS * 16) + warp * 16;
  const int g = lane >> 2, cq = (lane & 3) * 2;

  uint32_t ka[2][4], va[2][4];
  load_a_frag(k + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv, ld_kv, kv0 + g, cq, ka);
  load_a_frag(v + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv, ld_kv, kv0 + g, cq, va);

  const bf16* qb = q + static_cast<long long>(b) * n_q * ld_q + h * hs_q;
  const bf16* ob = dout + static_cast<long long>(b) * n_q * ld_do + h * 32;
  const float* lb = lse + (static_cast<long long>(b) * gridDim.y + h) * n_q;
  const float* db = dsum + (static_cast<long long>(b) * gridDim.y + h) * n_q;
  auto prefetch = [&](int tile, int st) {
    for (int i = threadIdx.x; i < TILE * 4; i += NT) {
      const int r = i >> 2, c = (i & 3) * 8;
      cp_async16(&Qs[st][r][c], qb + static_cast<long long>(tile * TILE + r) * ld_q + c, true);
      cp_async16(&Os[st][r][c], ob + static_cast<long long>(tile * TILE + r) * ld_do + c, true);
    }
    for (int i = threadIdx.x; i < 32; i += NT) {
      if (i < 16) cp_async16(&Ls[st][i * 4], lb + tile * TILE + i * 4, true);
      else cp_async16(&Ds[st][(i - 16) * 4], db + tile * TILE + (i - 16) * 4, true);
    }
  };

  float dva[4][4], dka[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dva[i][j] = dka[i][j] = 0.f;

  const int ntiles = n_q / TILE;
  prefetch(0, 0);
  cp_async_commit();
  for (int t = 0; t < ntiles; ++t) {
    const int st = t & 1;
    if (t + 1 < ntiles) {
      prefetch(t + 1, st ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    float p[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      p[nt][0] = p[nt][1] = p[nt][2] = p[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      uint32_t f[4];
      ldmatrix_x4(f, &Qs[st][nt * 8 + (lane & 7)][(lane >> 3) * 8]);   // B = Q^T: n = query, k = d
      mma_bf16(p[nt], ka[0], f[0], f[1]);
      mma_bf16(p[nt], ka[1], f[2], f[3]);
      ldmatrix_x4(f, &Os[st][nt * 8 + (lane & 7)][(lane >> 3) * 8]);   // B = dO^T
      mma_bf16(dp[nt], va[0], f[0], f[1]);
      mma_bf16(dp[nt], va[1], f[2], f[3]);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 l2 = *reinterpret_cast<const float2*>(&Ls[st][nt * 8 + cq]);
      const float2 d2 = *reinterpret_cast<const float2*>(&Ds[st][nt * 8 + cq]);
      p[nt][0] = ex2_approx(fmaf(p[nt][0], scale_log2, -l2.x));
      p[nt][1] = ex2_approx(fmaf(p[nt][1], scale_log2, -l2.y));
      p[nt][2] = ex2_approx(fmaf(p[nt][2], scale_log2, -l2.x));
      p[nt][3] = ex2_approx(fmaf(p[nt][3], scale_log2, -l2.y));
      dp[nt][0] = p[nt][0] * (dp[nt][0] - d2.x);   // dS^T (without the softmax scale)
      dp[nt][1] = p[nt][1] * (dp[nt][1] - d2.y);
      dp[nt][2] = p[nt][2] * (dp[nt][2] - d2.x);
      dp[nt][3] = p[nt][3] * (dp[nt][3] - d2.y);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 16 queries per k-step
      uint32_t pa[4], sa[4];
      pa[0] = pack_bf16(p[2 * j][0], p[2 * j][1]);
      pa[1] = pack_bf16(p[2 * j][2], p[2 * j][3]);
      pa[2] = pack_bf16(p[2 * j + 1][0], p[2 * j + 1][1]);
      pa[3] = pack_bf16(p[2 * j + 1][2], p[2 * j + 1][3]);
      sa[0] = pack_bf16(dp[2 * j][0], dp[2 * j][1]);
      sa[1] = pack_bf16(dp[2 * j][2], dp[2 * j][3]);
      sa[2] = pack_bf16(dp[2 * j + 1][0], dp[2 * j + 1][1]);
      sa[3] = pack_bf16(dp[2 * j + 1][2], dp[2 * j + 1][3]);
#pragma unroll
      for (int dh = 0; dh < 2; ++dh) {
        uint32_t f[4];
        ldmatrix_x4_trans(f, &Os[st][j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][dh * 16 + (lane >> 4) * 8]);
        mma_bf16(dva[dh * 2], pa, f[0], f[1]);
        mma_bf16(dva[dh * 2 + 1], pa, f[2], f[3]);
        ldmatrix_x4_trans(f, &Qs[st][j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][dh * 16 + (lane >> 4) * 8]);
        mma_bf16(dka[dh * 2], sa, f[0], f[1]);
        mma_bf16(dka[dh * 2 + 1], sa, f[2], f[3]);
      }
    }
    __syncthreads();
  }
  bf16* dkb = dk + static_cast<long long>(b) * n_kv * ld_dkv + h * hs_dkv;
  bf16* dvb = dv + static_cast<long long>(b) * n_kv * ld_dkv + h * hs_dkv;
  const long long r0 = kv0 + g, r1 = r0 + 8;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int c = nt * 8 + cq;
    *reinterpret_cast<uint32_t*>(dkb + r0 * ld_dkv + c) = pack_bf16(dka[nt][0] * scale, dka[nt][1] * scale);
    *reinterpret_cast<uint32_t*>(dkb + r1 * ld_dkv + c) = pack_bf16(dka[nt][2] * scale, dka[nt][3] * scale);
    *reinterpret_cast<uint32_t*>(dvb + r0 * ld_dkv + c) = pack_bf16(dva[nt][0], dva[nt][1]);
    *reinterpret_cast<uint32_t*>(dvb + r1 * ld_dkv + c) = pack_bf16(dva[nt][2], dva[nt][3]);
  }
}

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
dq_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
          const bf16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ dsum,
          long long ld_q, long long ld_kv, long long hs_q, long long hs_kv, long long ld_do, int n_q, int n_kv,
          float scale, float scale_log2, bf16* __restrict__ dq, long long ld_dq, long long hs_dq) {
  __shared__ __align__(16) bf16 Ks[2][TILE][ROW_PAD];
  __shared__ __align__(16) bf16 Vs[2][TILE][ROW_PAD];
  constexpr int NT = NWARPS * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * (NWARPS * 16) + warp * 16;
  const int g = lane >> 2, cq = (lane & 3) * 2;
  const int r0 = q0 + g, r1 = r0 + 8;

  uint32_t qa[2][4], oa[2][4];
  load_a_frag(q + static_cast<long long>(b) * n_q * ld_q + h * hs_q, ld_q, r0, cq, qa);
  load_a_frag(dout + static_cast<long long>(b) * n_q * ld_do + h * 32, ld_do, r0, cq, oa);
  const long long sb = (static_cast<long long>(b) * gridDim.y + h) * n_q;
  const float nl0 = -lse[sb + r0], nl1 = -lse[sb + r1];
  const float d0 = dsum[sb + r0], d1 = dsum[sb + r1];

  const bf16* kb = k + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  const bf16* vb = v + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  auto prefetch = [&](int tile, int st) {
    for (int i = threadIdx.x; i < TILE * 4; i += NT) {
      const int r = i >> 2, c = (i & 3) * 8;
      const long long off = static_cast<long long>(tile * TILE + r) * ld_kv + c;
      cp_async16(&Ks[st][r][c], kb + off, true);
      cp_async16(&Vs[st][r][c], vb + off, true);
    }
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ntiles = n_kv / TILE;
  prefetch(0, 0);
  cp_async_commit();
  for (int t = 0; t < ntiles; ++t) {
    const int st = t & 1;
    if (t + 1 < ntiles) {
      prefetch(t + 1, st ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float p[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      p[nt][0] = p[nt][1] = p[nt][2] = p[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      uint32_t f[4];
      ldmatrix_x4(f, &Ks[st][nt * 8 + (lane & 7)][(lane >> 3) * 8]);   // B = K^T: n = key, k = d
      mma_bf16(p[nt], qa[0], f[0], f[1]);
      mma_bf16(p[nt], qa[1], f[2], f[3]);
      ldmatrix_x4(f, &Vs[st][nt * 8 + (lane & 7)][(lane >> 3) * 8]);   // B = V^T
      mma_bf16(dp[nt], oa[0], f[0], f[1]);
      mma_bf16(dp[nt], oa[1], f[2], f[3]);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      p[nt][0] = ex2_approx(fmaf(p[nt][0], scale_log2, nl0)) * (dp[nt][0] - d0);   // dS (without the scale)
      p[nt][1] = ex2_approx(fmaf(p[nt][1], scale_log2, nl0)) * (dp[nt][1] - d0);
      p[nt][2] = ex2_approx(fmaf(p[nt][2], scale_log2, nl1)) * (dp[nt][2] - d1);
      p[nt][3] = ex2_approx(fmaf(p[nt][3], scale_log2, nl1)) * (dp[nt][3] - d1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 16 keys per k-step
      uint32_t sa[4];
      sa[0] = pack_bf16(p[2 * j][0], p[2 * j][1]);
      sa[1] = pack_bf16(p[2 * j][2], p[2 * j][3]);
      sa[2] = pack_bf16(p[2 * j + 1][0], p[2 * j + 1][1]);
      sa[3] = pack_bf16(p[2 * j + 1][2], p[2 * j + 1][3]);
#pragma unroll
      for (int dh = 0; dh < 2; ++dh) {
        uint32_t f[4];
        ldmatrix_x4_trans(f, &Ks[st][j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][dh * 16 + (lane >> 4) * 8]);
        mma_bf16(acc[dh * 2], sa, f[0], f[1]);
        mma_bf16(acc[dh * 2 + 1], sa, f[2], f[3]);
      }
    }
    __syncthreads();
  }
  bf16* dqb = dq + static_cast<long long>(b) * n_q * ld_dq + h * hs_dq;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int c = nt * 8 + cq;
    *reinterpret_cast<uint32_t*>(dqb + static_cast<long long>(r0) * ld_dq + c) =
        pack_bf16(acc[nt][0] * scale, acc[nt][1] * scale);
    *reinterpret_cast<uint32_t*>(dqb + static_cast<long long>(r1) * ld_dq + c) =
        pack_bf16(acc[nt][2] * scale, acc[nt][3] * scale);
  }
}

bool bwd_mma_ok(const ealdm_attention_bwd_args* a) {
  if (a->dtype != EALDM_BF16 || a->head_dim != 32 || a->lse == nullptr || a->impl == EALDM_IMPL_SIMT) return false;
  if (a->n_q % TILE != 0 || a->n_kv % TILE != 0 || a->scale <= 0.f) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(a->q) || !al16(a->k) || !al16(a->v) || !al16(a->out) || !al16(a->dout)) return false;
  if (a->ld_q % 8 || a->ld_kv % 8 || a->ld_out % 8 || a->ld_dout % 8 || a->head_stride_q % 8 || a->head_stride_kv % 8)
    return false;
  if (a->ld_dq % 2 || a->ld_dkv % 2 || a->head_stride_dq % 2 || a->head_stride_dkv % 2) return false;
  return true;
}

int launch_bwd_mma(const ealdm_attention_bwd_args* a, cudaStream_t st) {
  const bf16* q = reinterpret_cast<const bf16*>(a->q);
  const bf16* k = reinterpret_cast<const bf16*>(a->k);
  const bf16* v = reinterpret_cast<const bf16*>(a->v);
  const bf16* o = reinterpret_cast<const bf16*>(a->out);
  const bf16* d_o = reinterpret_cast<const bf16*>(a->dout);
  float* dsum = reinterpret_cast<float*>(a->workspace);
  const float scale_log2 = a->scale * 1.4426950408889634f;
  const long long rows = a->batch * a->n_q;
  const long long threads = rows * a->heads * 4;
  dsum_kernel<<<static_cast<unsigned>(ceil_div(threads, 256)), 256, 0, st>>>(o, d_o, a->ld_out, a->ld_dout,
                                                                             (int)a->heads, (int)a->n_q, rows, dsum);
  EALDM_LAUNCH_CHECK();
  dim3 g1(static_cast<unsigned>(a->n_kv / 64), static_cast<unsigned>(a->heads), static_cast<unsigned>(a->batch));
  dkdv_kernel<4><<<g1, 128, 0, st>>>(q, k, v, d_o, a->lse, dsum, a->ld_q, a->ld_kv, a->head_stride_q,
                                     a->head_stride_kv, a->ld_dout, (int)a->n_q, (int)a->n_kv, a->scale, scale_log2,
                                     reinterpret_cast<bf16*>(a->dk), reinterpret_cast<bf16*>(a->dv), a->ld_dkv,
                                     a->head_stride_dkv);
  EALDM_LAUNCH_CHECK();
  dim3 g2(static_cast<unsigned>(a->n_q / 64), static_cast<unsigned>(a->heads), static_cast<unsigned>(a->batch));
  dq_kernel<4><<<g2, 128, 0, st>>>(q, k, v, d_o, a->lse, dsum, a->ld_q, a->ld_kv, a->head_stride_q,
                                   a->head_stride_kv, a->ld_dout, (int)a->n_q, (int)a->n_kv, a->scale, scale_log2,
                                   reinterpret_cast<bf16*>(a->dq), a->ld_dq, a->head_stride_dq);
  EALDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace attn
}  // namespace ealdm
