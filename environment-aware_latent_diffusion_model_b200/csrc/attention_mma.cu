// bf16 tensor-core flash attention for head_dim 32 (UNet self-attention: 1024 / 256 / 64 tokens).
// One warp owns 16 queries; K/V tiles of 64 keys are double-buffered in shared memory with
// cp.async; S = Q.K^T and O += P.V run on mma.sync m16n8k16 with fp32 accumulation and an online
// softmax in the exp2 domain.  At d = 32 the kernel is bound by exp throughput (128 tensor FLOPs per
// exp), not by the tensor pipe, which is why it does not use tcgen05.
#include "common.cuh"
#include "mma.cuh"

namespace ealdm {
namespace attn {

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
flash_mma_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                 long long ld_q, long long ld_kv, long long hs_q, long long hs_kv, int n_q, int n_kv,
                 float scale_log2, bf16* __restrict__ out, long long ld_out) {
  __shared__ __align__(16) bf16 Ks[2][KV_TILE][ROW_PAD];
  __shared__ __align__(16) bf16 Vs[2][KV_TILE][ROW_PAD];
  constexpr int NT = NWARPS * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * (NWARPS * 16) + warp * 16;
  const int r0 = q0 + (lane >> 2), r1 = r0 + 8;
  const int cq = (lane & 3) * 2;

  uint32_t qa[2][4];
  {
    const bf16* qb = q + static_cast<long long>(b) * n_q * ld_q + h * hs_q;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int c = ks * 16 + cq;
      qa[ks][0] = r0 < n_q ? *reinterpret_cast<const uint32_t*>(qb + r0 * ld_q + c) : 0u;
      qa[ks][1] = r1 < n_q ? *reinterpret_cast<const uint32_t*>(qb + r1 * ld_q + c) : 0u;
      qa[ks][2] = r0 < n_q ? *reinterpret_cast<const uint32_t*>(qb + r0 * ld_q + c + 8) : 0u;
      qa[ks][3] = r1 < n_q ? *reinterpret_cast<const uint32_t*>(qb + r1 * ld_q + c + 8) : 0u;
    }
  }

  const bf16* kb = k + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  const bf16* vb = v + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  auto prefetch = [&](int tile, int st) {
    // 64 rows x 4 chunks of 16 B for K and for V
    for (int i = threadIdx.x; i < KV_TILE * 4; i += NT) {
      const int r = i >> 2, c = (i & 3) * 8;
      const int key = tile * KV_TILE + r;
      const bool ok = key < n_kv;
      const long long off = static_cast<long long>(ok ? key : 0) * ld_kv + c;
      cp_async16(&Ks[st][r][c], kb + off, ok);
      cp_async16(&Vs[st][r][c], vb + off, ok);
    }
  };

  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int ntiles = (n_kv + KV_TILE - 1) / KV_TILE;
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

    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t kf[4];
      ldmatrix_x4(kf, &Ks[st][nt * 8 + (lane & 7)][(lane >> 3) * 8]);
      mma_bf16(s[nt], qa[0], kf[0], kf[1]);
      mma_bf16(s[nt], qa[1], kf[2], kf[3]);
    }
    const bool last = (t + 1) * KV_TILE > n_kv;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float val = s[nt][j] * scale_log2;
        if (last) {
          const int key = t * KV_TILE + nt * 8 + cq + (j & 1);
          if (key >= n_kv) val = -INFINITY;
        }
        s[nt][j] = val;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mn0);
      s[nt][1] = exp2f(s[nt][1] - mn0);
      s[nt][2] = exp2f(s[nt][2] - mn1);
      s[nt][3] = exp2f(s[nt][3] - mn1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 16 keys per k-step
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * j][0], s[2 * j][1]);
      pa[1] = pack_bf16(s[2 * j][2], s[2 * j][3]);
      pa[2] = pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[3] = pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, &Vs[st][j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8]
                                 [dp * 16 + (lane >> 4) * 8]);
        mma_bf16(o[dp * 2], pa, vf[0], vf[1]);
        mma_bf16(o[dp * 2 + 1], pa, vf[2], vf[3]);
      }
    }
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  bf16* ob = out + static_cast<long long>(b) * n_q * ld_out + h * 32;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int c = nt * 8 + cq;
    if (r0 < n_q)
      *reinterpret_cast<uint32_t*>(ob + r0 * ld_out + c) = pack_bf16(o[nt][0] * i0, o[nt][1] * i0);
    if (r1 < n_q)
      *reinterpret_cast<uint32_t*>(ob + r1 * ld_out + c) = pack_bf16(o[nt][2] * i1, o[nt][3] * i1);
  }
}

// ---- lean variant for n_kv % 64 == 0 (every UNet level) -----------------------------------------------
// At d = 32 the kernel is issue / MUFU bound, so every instruction per score counts:
//   * no tail masking;
//   * softmax scale, log2(e) and the running max folded into ONE FFMA feeding ex2: p = 2^(s*c - m*c);
//   * the row sums are produced by the tensor core: the P.V MMA gets a fifth 8-column tile whose
//     column 0 is all ones, so l accumulates (and is rescaled) exactly like O.

// The kernel is latency bound (MUFU and mma.sync dependency chains: XU pipe 41 %, legacy tensor pipe 43 %, issue
// 42 % at 4 warps per scheduler -- profiles/r01_attention_full.txt), so residency matters more than tile reuse:
// the register budget is capped at 80 and CTAs are 4 warps, 6 per SM (24 warps): 653 -> 556 us at level 0.
// (A packed ex2.approx.ftz.bf16x2 does NOT halve the MUFU work: ptxas splits it into two MUFU.EX2.BF16.)
template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 768 / (NWARPS * 32))
flash_mma_even_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                      long long ld_q, long long ld_kv, long long hs_q, long long hs_kv, int n_q, int n_kv,
                      float scale_log2, bf16* __restrict__ out, long long ld_out, float* __restrict__ lse) {
  __shared__ __align__(16) bf16 Ks[2][KV_TILE][ROW_PAD];
  __shared__ __align__(16) bf16 Vs[2][KV_TILE][ROW_PAD];
  constexpr int NT = NWARPS * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * (NWARPS * 16) + warp * 16;
  const int r0 = q0 + (lane >> 2), r1 = r0 + 8;
  const int cq = (lane & 3) * 2;

  pdl_wait();      // launched with programmatic serialization (common.cuh)
  pdl_trigger();
  uint32_t qa[2][4];
  {
    const bf16* qb = q + static_cast<long long>(b) * n_q * ld_q + h * hs_q;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int c = ks * 16 + cq;
      qa[ks][0] = r0 < n_q ? *reinterpret_cast<const uint32_t*>(qb + r0 * ld_q + c) : 0u;
      qa[ks][1] = r1 < n_q ? *reinterpret_cast<const uint32_t*>(qb + r1 * ld_q + c) : 0u;
      qa[ks][2] = r0 < n_q ? *reinterpret_cast<const uint32_t*>(qb + r0 * ld_q + c + 8) : 0u;
      qa[ks][3] = r1 < n_q ? *reinterpret_cast<const uint32_t*>(qb + r1 * ld_q + c + 8) : 0u;
    }
  }
  const bf16* kb = k + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  const bf16* vb = v + static_cast<long long>(b) * n_kv * ld_kv + h * hs_kv;
  auto prefetch = [&](int tile, int st) {
    for (int i = threadIdx.x; i < KV_TILE * 4; i += NT) {
      const int r = i >> 2, c = (i & 3) * 8;
      const long long off = static_cast<long long>(tile * KV_TILE + r) * ld_kv + c;
      cp_async16(&Ks[st][r][c], kb + off, true);
      cp_async16(&Vs[st][r][c], vb + off, true);
    }
  };
  // B fragment of the ones tile: B[k][0] = 1 for every k, other columns 0
  const uint32_t b_ones = (lane < 4) ? 0x3F803F80u : 0u;

  float o[5][4];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY;  // running row maxima of the RAW scores

  const int ntiles = n_kv / KV_TILE;
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

    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t kf[4];
      ldmatrix_x4(kf, &Ks[st][nt * 8 + (lane & 7)][(lane >> 3) * 8]);
      mma_bf16(s[nt], qa[0], kf[0], kf[1]);
      mma_bf16(s[nt], qa[1], kf[2], kf[3]);
    }
    float mx0 = fmaxf(s[0][0], s[0][1]), mx1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
    for (int nt = 1; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = ex2_approx((m0 - mn0) * scale_log2), c1 = ex2_approx((m1 - mn1) * scale_log2);
    m0 = mn0; m1 = mn1;
    const float ms0 = -mn0 * scale_log2, ms1 = -mn1 * scale_log2;
#pragma unroll
    for (int i = 0; i < 5; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 16 keys per k-step
      uint32_t pa[4];
      pa[0] = pack_bf16(ex2_approx(fmaf(s[2 * j][0], scale_log2, ms0)), ex2_approx(fmaf(s[2 * j][1], scale_log2, ms0)));
      pa[1] = pack_bf16(ex2_approx(fmaf(s[2 * j][2], scale_log2, ms1)), ex2_approx(fmaf(s[2 * j][3], scale_log2, ms1)));
      pa[2] = pack_bf16(ex2_approx(fmaf(s[2 * j + 1][0], scale_log2, ms0)),
                        ex2_approx(fmaf(s[2 * j + 1][1], scale_log2, ms0)));
      pa[3] = pack_bf16(ex2_approx(fmaf(s[2 * j + 1][2], scale_log2, ms1)),
                        ex2_approx(fmaf(s[2 * j + 1][3], scale_log2, ms1)));
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, &Vs[st][j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8]
                                 [dp * 16 + (lane >> 4) * 8]);
        mma_bf16(o[dp * 2], pa, vf[0], vf[1]);
        mma_bf16(o[dp * 2 + 1], pa, vf[2], vf[3]);
      }
      mma_bf16(o[4], pa, b_ones, b_ones);
    }
    __syncthreads();
  }
  // column 0 of the ones tile lives on the lane of each quad with (lane & 3) == 0
  const float l0 = __shfl_sync(0xffffffffu, o[4][0], lane & ~3);
  const float l1 = __shfl_sync(0xffffffffu, o[4][2], lane & ~3);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  if (lse != nullptr && (lane & 3) == 0) {
    // log2-domain log-sum-exp of the scaled scores, saved for the backward: p_ij = 2^(s_ij * c - lse_i)
    float* lp = lse + (static_cast<long long>(b) * gridDim.y + h) * n_q;
    if (r0 < n_q) lp[r0] = fmaf(m0, scale_log2, log2f(l0));
    if (r1 < n_q) lp[r1] = fmaf(m1, scale_log2, log2f(l1));
  }
  bf16* ob = out + static_cast<long long>(b) * n_q * ld_out + h * 32;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int c = nt * 8 + cq;
    if (r0 < n_q)
      *reinterpret_cast<uint32_t*>(ob + r0 * ld_out + c) = pack_bf16(o[nt][0] * i0, o[nt][1] * i0);
    if (r1 < n_q)
      *reinterpret_cast<uint32_t*>(ob + r1 * ld_out + c) = pack_bf16(o[nt][2] * i1, o[nt][3] * i1);
  }
}

int launch_flash_mma(const ealdm_attention_args* a, cudaStream_t st) {
  EALDM_REQUIRE(a->ld_kv % 8 == 0 && a->head_stride_kv % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(a->k) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a->v) & 15) == 0,
                "attention(mma): K/V need 16-byte aligned rows");
  EALDM_REQUIRE(a->ld_q % 2 == 0 && a->head_stride_q % 2 == 0 && a->ld_out % 2 == 0,
                "attention(mma): q/out pitches must be even");
  const float scale_log2 = a->scale * 1.4426950408889634f;
  const bf16* q = reinterpret_cast<const bf16*>(a->q);
  const bf16* k = reinterpret_cast<const bf16*>(a->k);
  const bf16* v = reinterpret_cast<const bf16*>(a->v);
  bf16* o = reinterpret_cast<bf16*>(a->out);
  EALDM_REQUIRE(a->lse == nullptr || (a->n_kv % KV_TILE == 0 && a->scale > 0.f),
                "attention(mma): lse output needs n_kv %% 64 == 0");
  if (a->n_kv % KV_TILE == 0 && a->scale > 0.f) {
    if (false) {   // 8-warp CTAs (128 queries): kept for reference, slower than 6 resident 4-warp CTAs
      dim3 grid(static_cast<unsigned>(ceil_div(a->n_q, 128)), static_cast<unsigned>(a->heads),
                static_cast<unsigned>(a->batch));
      flash_mma_even_kernel<8><<<grid, 256, 0, st>>>(q, k, v, a->ld_q, a->ld_kv, a->head_stride_q,
                                                     a->head_stride_kv, (int)a->n_q, (int)a->n_kv,
                                                     scale_log2, o, a->ld_out, a->lse);
    } else {
      dim3 grid(static_cast<unsigned>(ceil_div(a->n_q, 64)), static_cast<unsigned>(a->heads),
                static_cast<unsigned>(a->batch));
      EALDM_CUDA(launch_pdl(flash_mma_even_kernel<4>, grid, dim3(128), 0, st, q, k, v, a->ld_q, a->ld_kv,
                            a->head_stride_q, a->head_stride_kv, (int)a->n_q, (int)a->n_kv, scale_log2, o, a->ld_out,
                            a->lse));
    }
  } else if (a->n_q > 64) {
    dim3 grid(static_cast<unsigned>(ceil_div(a->n_q, 128)), static_cast<unsigned>(a->heads),
              static_cast<unsigned>(a->batch));
    flash_mma_kernel<8><<<grid, 256, 0, st>>>(q, k, v, a->ld_q, a->ld_kv, a->head_stride_q,
                                              a->head_stride_kv, (int)a->n_q, (int)a->n_kv,
                                              scale_log2, o, a->ld_out);
  } else {
    dim3 grid(static_cast<unsigned>(ceil_div(a->n_q, 64)), static_cast<unsigned>(a->heads),
              static_cast<unsigned>(a->batch));
    flash_mma_kernel<4><<<grid, 128, 0, st>>>(q, k, v, a->ld_q, a->ld_kv, a->head_stride_q,
                                              a->head_stride_kv, (int)a->n_q, (int)a->n_kv,
                                              scale_log2, o, a->ld_out);
  }
  EALDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace attn
}  // namespace ealdm
