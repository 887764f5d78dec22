// GroupNorm (+SiLU) and LayerNorm over NHWC activations, plus a row softmax.  HBM-bound kernels:
// every pixel row is read with consecutive threads on consecutive channels (coalesced); GroupNorm
// statistics go fp32 per thread -> fp32 per (CTA, 4-channel vector) -> fp64 per (image, group) in a
// fixed summation order (no atomics, bit-reproducible).
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

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
      constexpr int U = 8;
      for (int pix0 = p0 + tp; pix0 < p1; pix0 += U * pix_lanes) {
        Vec4<TX> q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pix = pix0 + u * pix_lanes;
          if (pix < p1) q[u].load(base + static_cast<long long>(pix) * ld + v * 4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (pix0 + u * pix_lanes < p1) {
            float f[4];
            q[u].get(f);
#pragma unroll
            for (int j = 0; j < 4; ++j) { s += f[j]; ss = fmaf(f[j], f[j], ss); }
          }
        }
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
                const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                float* __restrict__ stats_out) {
  // bf16 outputs: SiLU with the fast exp / divide intrinsics (error far below bf16 resolution)
  constexpr bool SILU_FAST = sizeof(T) == 2;
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
        if (stats_out != nullptr && blockIdx.x == 0) {
          stats_out[(static_cast<long long>(n) * groups + g) * 2] = s_mean[g];
          stats_out[(static_cast<long long>(n) * groups + g) * 2 + 1] = s_rstd[g];
        }
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
      // 8 independent 16-byte loads in flight per thread before any of them is consumed
      constexpr int U = 8;
      for (int pix0 = p0 + tp; pix0 < p1; pix0 += U * pix_lanes) {
        Vec4<TX> qx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pix = pix0 + u * pix_lanes;
          if (pix < p1) qx[u].load(xb + static_cast<long long>(pix) * ld_x + ch);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int pix = pix0 + u * pix_lanes;
          if (pix < p1) {
            float f[4];
            qx[u].get(f);
            f[0] = fmaf(f[0], a0, b0); f[1] = fmaf(f[1], a1, b1); f[2] = fmaf(f[2], a2, b2); f[3] = fmaf(f[3], a3, b3);
            if (act == EALDM_ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 4; ++j) f[j] = SILU_FAST ? __fdividef(f[j], 1.0f + __expf(-f[j])) : silu_f(f[j]);
            }
            Vec4<T> q;
            q.set(f);
            q.store(yb + static_cast<long long>(pix) * ld_y + ch);
          }
        }
      }
    }
  }
}

// ---- GroupNorm from partial statistics: a pure streaming pass ------------------------------------------------
// The producing convolution's epilogue (conv_tc.cu) or gn_partial_kernel has already written, for every
// (image, 32-pixel chunk, 8-channel octet), {sum, sum of squares}.  Every CTA folds the partials of its image in a
// fixed order (fp64: bit-reproducible), then streams its pixel rows once: no reduction pass over x, no barrier
// between a load and its store.
template <typename TX, typename T>
__global__ void __launch_bounds__(NT)
gn_apply_partial_kernel(const TX* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, int hw, int c,
                        int groups, int pix_per_cta, const float2* __restrict__ partial, long long pld, float eps,
                        const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                        float* __restrict__ stats_out, int unit_shift) {
  constexpr bool SILU_FAST = sizeof(T) == 2;
  __shared__ float s_mean[MAX_GROUPS], s_rstd[MAX_GROUPS];
  pdl_wait();      // launched with programmatic serialization (common.cuh): x / partial belong to the predecessor
  pdl_trigger();
  const int t = threadIdx.x;
  const int n = blockIdx.y;
  const int cpg = c / groups;
  const int vpp = c >> 2;
  {
    const int chunks = hw >> 5;
    const int opg = cpg >> unit_shift;   // partial entries (octets / quads) per group
    const int terms = chunks * opg;
    const float2* pn = partial + static_cast<long long>(n) * chunks * pld;
    for (int g0 = 0; g0 < groups; g0 += NT / 8) {
      const int g = g0 + (t >> 3);
      const int sub = t & 7;
      // Eight loads in flight per thread and round, added pairwise in fp32; the rounds are accumulated in fp64 (one
      // add per round and component).  A chain of dependent (load -> fp64 add) steps per term, and the fp64 reciprocal
      // square root behind it, made this prologue 5-12 us per CTA at the UNet's shapes (180 us at 256 x 256): fp64
      // arithmetic is slow on this part (clock64 trace of the same fold in the conv epilogue)
      double s = 0.0, ss = 0.0;
      if (g < groups) {
        for (int k0 = sub; k0 < terms; k0 += 64) {
          float2 q[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int k = k0 + 8 * u;
            const int ch = k / opg, oo = k - ch * opg;
            q[u] = k < terms ? pn[static_cast<long long>(ch) * pld + g * opg + oo] : make_float2(0.f, 0.f);
          }
          s += static_cast<double>(((q[0].x + q[1].x) + (q[2].x + q[3].x)) + ((q[4].x + q[5].x) + (q[6].x + q[7].x)));
          ss += static_cast<double>(((q[0].y + q[1].y) + (q[2].y + q[3].y)) + ((q[4].y + q[5].y) + (q[6].y + q[7].y)));
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
      }
      if (g < groups && sub == 0) {
        const double inv_cnt = 1.0 / (static_cast<double>(hw) * cpg);
        const double mean = s * inv_cnt;
        const float var = fmaxf(static_cast<float>(ss * inv_cnt - mean * mean), 0.f);
        s_mean[g] = static_cast<float>(mean);
        s_rstd[g] = rsqrtf(var + eps);
        if (stats_out != nullptr && blockIdx.x == 0) {
          stats_out[(static_cast<long long>(n) * groups + g) * 2] = s_mean[g];
          stats_out[(static_cast<long long>(n) * groups + g) * 2 + 1] = s_rstd[g];
        }
      }
    }
  }
  __syncthreads();
  const int lanes_v = vpp < NT ? vpp : NT;
  const int pix_lanes = NT / lanes_v;
  const int tv = t % lanes_v, tp = t / lanes_v;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const TX* xb = x + static_cast<long long>(n) * hw * ld_x;
  T* yb = y + static_cast<long long>(n) * hw * ld_y;
  if (tp >= pix_lanes) return;
  for (int v = tv; v < vpp; v += lanes_v) {
    const int ch = v * 4;
    const int g = ch / cpg;
    const float mean = s_mean[g], rstd = s_rstd[g];
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + ch));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + ch));
    const float a0 = rstd * ga.x, a1 = rstd * ga.y, a2 = rstd * ga.z, a3 = rstd * ga.w;
    const float b0 = be.x - mean * a0, b1 = be.y - mean * a1, b2 = be.z - mean * a2, b3 = be.w - mean * a3;
    constexpr int U = 8;
    for (int pix0 = p0 + tp; pix0 < p1; pix0 += U * pix_lanes) {
      Vec4<TX> qx[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * pix_lanes;
        if (pix < p1) qx[u].load(xb + static_cast<long long>(pix) * ld_x + ch);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * pix_lanes;
        if (pix < p1) {
          float f[4];
          qx[u].get(f);
          f[0] = fmaf(f[0], a0, b0); f[1] = fmaf(f[1], a1, b1); f[2] = fmaf(f[2], a2, b2); f[3] = fmaf(f[3], a3, b3);
          if (act == EALDM_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) f[j] = SILU_FAST ? __fdividef(f[j], 1.0f + __expf(-f[j])) : silu_f(f[j]);
          }
          Vec4<T> q;
          q.set(f);
          q.store(yb + static_cast<long long>(pix) * ld_y + ch);
        }
      }
    }
  }
}

// stand-alone producer of the partials: one CTA per (32-pixel chunk, image), one thread per octet
template <typename TX>
__global__ void __launch_bounds__(NT)
gn_partial_kernel(const TX* __restrict__ x, long long ld_x, int hw, int c, float2* __restrict__ partial,
                  long long pld) {
  const int n = blockIdx.y, chunk = blockIdx.x;
  const TX* xb = x + (static_cast<long long>(n) * hw + chunk * 32) * ld_x;
  float2* out = partial + (static_cast<long long>(n) * (hw >> 5) + chunk) * pld;
  for (int o = threadIdx.x; o < (c >> 3); o += NT) {
    float s = 0.f, ss = 0.f;
    for (int p = 0; p < 32; ++p) {
      Vec4<TX> q0, q1;
      q0.load(xb + static_cast<long long>(p) * ld_x + o * 8);
      q1.load(xb + static_cast<long long>(p) * ld_x + o * 8 + 4);
      float f[4], h[4];
      q0.get(f);
      q1.get(h);
#pragma unroll
      for (int j = 0; j < 4; ++j) { s += f[j]; ss = fmaf(f[j], f[j], ss); }
#pragma unroll
      for (int j = 0; j < 4; ++j) { s += h[j]; ss = fmaf(h[j], h[j], ss); }
    }
    out[o] = make_float2(s, ss);
  }
}

// pixel chunking shared by the workspace query and the launch
static void gn_chunking(long long n, long long hw, int* pix_per_cta, long long* chunks) {
  // enough CTAs to fill 148 SMs (8 resident CTAs each) twice over, at least 16 pixels each
  long long ch = ceil_div(2368, n);
  const long long max_chunks = ceil_div(hw, 16);
  if (ch > max_chunks) ch = max_chunks;
  if (ch < 1) ch = 1;
  const int ppc = static_cast<int>(ceil_div(hw, ch));
  *pix_per_cta = ppc;
  *chunks = ceil_div(hw, ppc);
}

template <typename TX, typename T>
static int gn_cluster_plan(const ealdm_group_norm_args* a, int* ppc_out, size_t* smem_out, int* nsplit_out);
template <typename TX, typename T>
static int group_norm_cluster(const ealdm_group_norm_args* a, int cl, int ppc, size_t smem, int nsplit,
                              cudaStream_t st, bool* launched);

template <typename TX, typename T>
static int group_norm_t(const ealdm_group_norm_args* a, cudaStream_t st) {
  if (a->partial != nullptr) {
    int ppc;
    long long chunks;
    {
      // fewer, fatter CTAs than the two-pass kernels: every CTA first folds the partial statistics of its image
      // (measured at batch 128: 296 CTAs 45 us, 1184 CTAs 49 us, 2368 CTAs 53 us, 4736 CTAs 63 us at level 0)
      const long long target = 296;
      long long ch = ceil_div(target, a->n);
      const long long max_chunks = ceil_div(a->hw, 16);
      if (ch > max_chunks) ch = max_chunks;
      if (ch < 1) ch = 1;
      ppc = static_cast<int>(ceil_div(a->hw, ch));
      chunks = ceil_div(a->hw, ppc);
    }
    dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(a->n));
    EALDM_CUDA(launch_pdl(gn_apply_partial_kernel<TX, T>, grid, dim3(NT), 0, st, reinterpret_cast<const TX*>(a->x),
                          a->ld_x, reinterpret_cast<T*>(a->y), a->ld_y, static_cast<int>(a->hw), static_cast<int>(a->c),
                          a->groups, ppc, reinterpret_cast<const float2*>(a->partial), a->partial_ld, a->eps, a->gamma,
                          a->beta, a->act, a->stats_out, a->partial_unit == 4 ? 2 : 3));
    EALDM_LAUNCH_CHECK();
    return 0;
  }
  {
    int ppc_c = 0, nsplit = 1;
    size_t smem = 0;
    const int cl = gn_cluster_plan<TX, T>(a, &ppc_c, &smem, &nsplit);
    if (cl > 0) {
      bool launched = false;
      if (int e = group_norm_cluster<TX, T>(a, cl, ppc_c, smem, nsplit, st, &launched)) return e;
      if (launched) return 0;
    }
  }
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
                                          a->gamma, a->beta, a->act, a->stats_out);
  EALDM_LAUNCH_CHECK();
  return 0;
}

// ---- single-pass GroupNorm on a thread-block cluster ---------------------------------------------------
// One cluster per image; CTA r of the cluster pulls its slab of pixel rows into shared memory with bulk
// TMA copies (one per row, all in flight at once), so x crosses HBM exactly once.  Statistics: fp32 per
// thread -> fp64 per (CTA, group) in a fixed order -> fp64 over the cluster's CTAs through distributed
// shared memory, rank 0 first (bit-reproducible, identical in every CTA).  The affine transform (+SiLU)
// is then applied from shared memory.
constexpr int CNT = 256;

__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :
      : "r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(ptx::smem_u32(bar))
      : "memory");
}

template <typename TX, typename T>
__global__ void __launch_bounds__(CNT)
gn_cluster_kernel(const TX* __restrict__ x, long long ld_x, T* __restrict__ y, long long ld_y, int hw,
                  int c, int groups, int ppc, float eps, const float* __restrict__ gamma,
                  const float* __restrict__ beta, int act, float* __restrict__ stats_out) {
  // blockIdx.z selects a channel range of `c` channels (`groups` whole groups) of the image
  x += static_cast<long long>(blockIdx.z) * c;
  y += static_cast<long long>(blockIdx.z) * c;
  gamma += static_cast<long long>(blockIdx.z) * c;
  beta += static_cast<long long>(blockIdx.z) * c;
  constexpr bool SILU_FAST = sizeof(T) == 2;
  extern __shared__ __align__(16) uint8_t gsm[];
  const int row_bytes = c * static_cast<int>(sizeof(TX));
  TX* slab = reinterpret_cast<TX*>(gsm);
  uint8_t* tail = gsm + static_cast<size_t>(ppc) * row_bytes;
  double* gpart = reinterpret_cast<double*>(tail);                       // [MAX_GROUPS][2]
  float2* red = reinterpret_cast<float2*>(tail + MAX_GROUPS * 16);       // [CNT]
  float* s_mean = reinterpret_cast<float*>(tail + MAX_GROUPS * 16 + CNT * 8);
  float* s_rstd = s_mean + MAX_GROUPS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_rstd + MAX_GROUPS);

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int nranks = static_cast<int>(cluster.num_blocks());
  const int t = threadIdx.x;
  const int n = blockIdx.y;
  const int p0 = rank * ppc;
  const TX* xb = x + (static_cast<long long>(n) * hw + p0) * ld_x;
  T* yb = y + (static_cast<long long>(n) * hw + p0) * ld_y;

  if (t == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive_expect_tx(bar, static_cast<uint32_t>(ppc) * row_bytes);
  }
  __syncthreads();
  for (int r = t; r < ppc; r += CNT)
    bulk_copy_g2s(reinterpret_cast<uint8_t*>(slab) + static_cast<size_t>(r) * row_bytes,
                  xb + static_cast<long long>(r) * ld_x, row_bytes, bar);
  ptx::mbar_wait(bar, 0);

  const int cpg = c / groups;
  const int vpg = cpg >> 2;
  const int vpp = c >> 2;
  const int lanes_v = vpp <= CNT ? vpp : (CNT / vpg) * vpg;
  const int pix_lanes = vpp <= CNT ? CNT / vpp : 1;
  const int tv = t % lanes_v, tp = t / lanes_v;
  for (int v0 = 0; v0 < vpp; v0 += lanes_v) {
    const int v = v0 + tv;
    float s = 0.f, ss = 0.f;
    if (tp < pix_lanes && v < vpp) {
#pragma unroll 4
      for (int pix = tp; pix < ppc; pix += pix_lanes) {
        Vec4<TX> q;
        q.load(slab + static_cast<size_t>(pix) * c + v * 4);
        float f[4];
        q.get(f);
#pragma unroll
        for (int j = 0; j < 4; ++j) { s += f[j]; ss = fmaf(f[j], f[j], ss); }
      }
    }
    red[t] = make_float2(s, ss);
    __syncthreads();
    const int nv = min(lanes_v, vpp - v0);
    if (t < nv / vpg) {
      double S = 0.0, SS = 0.0;
      for (int k = 0; k < pix_lanes; ++k)
        for (int vv = 0; vv < vpg; ++vv) {
          const float2 e = red[k * lanes_v + t * vpg + vv];
          S += static_cast<double>(e.x);
          SS += static_cast<double>(e.y);
        }
      gpart[2 * (v0 / vpg + t)] = S;
      gpart[2 * (v0 / vpg + t) + 1] = SS;
    }
    __syncthreads();
  }
  cluster.sync();
  if (t < groups) {
    // every rank's partials with independent remote loads (one DSMEM round trip), summed in rank order
    double2 rp[16];
#pragma unroll
    for (int r = 0; r < 16; ++r)
      rp[r] = r < nranks ? *reinterpret_cast<const double2*>(cluster.map_shared_rank(gpart, r) + 2 * t)
                         : make_double2(0.0, 0.0);
    double S = 0.0, SS = 0.0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      S += rp[r].x;
      SS += rp[r].y;
    }
    const double cnt = static_cast<double>(hw) * cpg;
    const double mean = S / cnt;
    double var = SS / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[t] = static_cast<float>(mean);
    s_rstd[t] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (stats_out != nullptr && rank == 0) {
      const long long gi = (static_cast<long long>(n) * gridDim.z + blockIdx.z) * groups + t;
      stats_out[gi * 2] = s_mean[t];
      stats_out[gi * 2 + 1] = s_rstd[t];
    }
  }
  __syncthreads();
  // this CTA is done reading its peers: arrive now (no global stores outstanding yet, so the release
  // fence is free) and wait only at the very end -- shared memory must outlive the peers' reads
  cluster.barrier_arrive();
  for (int v0 = 0; v0 < vpp; v0 += lanes_v) {
    const int v = v0 + tv;
    if (tp < pix_lanes && v < vpp) {
      const int ch = v * 4;
      const int g = ch / cpg;
      const float mean = s_mean[g], rstd = s_rstd[g];
      const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + ch));
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + ch));
      const float a0 = rstd * ga.x, a1 = rstd * ga.y, a2 = rstd * ga.z, a3 = rstd * ga.w;
      const float b0 = be.x - mean * a0, b1 = be.y - mean * a1, b2 = be.z - mean * a2, b3 = be.w - mean * a3;
#pragma unroll 4
      for (int pix = tp; pix < ppc; pix += pix_lanes) {
        Vec4<TX> qx;
        qx.load(slab + static_cast<size_t>(pix) * c + ch);
        float f[4];
        qx.get(f);
        f[0] = fmaf(f[0], a0, b0); f[1] = fmaf(f[1], a1, b1); f[2] = fmaf(f[2], a2, b2); f[3] = fmaf(f[3], a3, b3);
        if (act == EALDM_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) f[j] = SILU_FAST ? __fdividef(f[j], 1.0f + __expf(-f[j])) : silu_f(f[j]);
        }
        Vec4<T> q;
        q.set(f);
        q.store(yb + static_cast<long long>(pix) * ld_y + ch);
      }
    }
  }
  cluster.barrier_wait();
}

constexpr int GN_CLUSTER_TAIL = MAX_GROUPS * 16 + CNT * 8 + MAX_GROUPS * 8 + 16;

// (cluster size, channel split) for the single-pass kernel: the smallest launch whose per-CTA slab lets
// 3 CTAs share an SM (loads of one CTA overlap the reduction / stores of its neighbours); 0 if none.
template <typename TX, typename T>
static int gn_cluster_plan(const ealdm_group_norm_args* a, int* ppc_out, size_t* smem_out, int* nsplit_out) {
  if ((a->c * static_cast<long long>(sizeof(TX))) % 16 != 0 ||
      (a->ld_x * static_cast<long long>(sizeof(TX))) % 16 != 0 || (reinterpret_cast<uintptr_t>(a->x) & 15) != 0)
    return 0;
  const int cpg = static_cast<int>(a->c / a->groups);
  const int vpg = cpg >> 2;
  if (vpg < 1 || vpg > CNT) return 0;
  static const char* e_off = getenv("EALDM_GN_OFF");       // tuning switches
  static const char* e_cl = getenv("EALDM_GN_MAXCL");
  static const char* e_lim = getenv("EALDM_GN_LIMIT_KB");
  if (e_off != nullptr) return 0;
  const int max_cl = e_cl ? atoi(e_cl) : 16;
  const long long limit = (e_lim ? atoi(e_lim) : 74) * 1024LL;
  for (int ns = 1; ns <= 8; ns *= 2) {
    if (a->groups % ns != 0) break;
    const long long row_bytes = (a->c / ns) * static_cast<long long>(sizeof(TX));
    if (row_bytes % 16 != 0) break;
    for (int cl = 1; cl <= max_cl; cl *= 2) {
      if (a->hw % cl != 0) break;
      const long long ppc = a->hw / cl;
      const long long smem = ppc * row_bytes + GN_CLUSTER_TAIL;
      if (smem > limit) continue;
      *ppc_out = static_cast<int>(ppc);
      *smem_out = static_cast<size_t>(smem);
      *nsplit_out = ns;
      return cl;
    }
  }
  return 0;
}

template <typename TX, typename T>
static int group_norm_cluster(const ealdm_group_norm_args* a, int cl, int ppc, size_t smem, int nsplit,
                              cudaStream_t st, bool* launched) {
  *launched = false;
  auto kern = gn_cluster_kernel<TX, T>;
  // per device: opt in to the largest dynamic shared memory a CTA may use (once) and to non-portable cluster sizes
  static DeviceOnce smem_set, nonportable_set;
  if (smem_set.pending()) {
    int dev = 0, optin = 0;
    EALDM_CUDA(cudaGetDevice(&dev));
    EALDM_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    EALDM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    smem_set.done();
  }
  if (cl > 8 && nonportable_set.pending()) {
    EALDM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    nonportable_set.done();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(cl), static_cast<unsigned>(a->n), static_cast<unsigned>(nsplit));
  cfg.blockDim = dim3(CNT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cl);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // the cluster shape must be placeable on this device (checked once per shape)
  static int ok_cl[17] = {0};
  static size_t ok_smem[17] = {0};
  if (ok_cl[cl] == 0 || smem > ok_smem[cl]) {
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      ok_cl[cl] = -1;
    } else {
      ok_cl[cl] = 1;
    }
    ok_smem[cl] = smem;
  }
  if (ok_cl[cl] < 0) return 0;
  const TX* x = reinterpret_cast<const TX*>(a->x);
  T* y = reinterpret_cast<T*>(a->y);
  const int hw = static_cast<int>(a->hw), c = static_cast<int>(a->c / nsplit);
  EALDM_CUDA(cudaLaunchKernelEx(&cfg, kern, x, a->ld_x, y, a->ld_y, hw, c, a->groups / nsplit, ppc, a->eps,
                                a->gamma, a->beta, a->act, a->stats_out));
  count_launch();
  *launched = true;
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
  {
    const int unit = a->partial_unit == 4 ? 4 : 8;
    EALDM_REQUIRE(a->partial == nullptr || ((a->partial_unit == 0 || a->partial_unit == 4 || a->partial_unit == 8) &&
                                            a->hw % 32 == 0 && (a->c / a->groups) % unit == 0 &&
                                            a->partial_ld >= a->c / unit),
                  "group_norm: partial statistics need hw %% 32 == 0 and channels per group %% partial_unit == 0");
  }
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

extern "C" int ealdm_gn_partial(const void* x, int64_t ld_x, int32_t dtype, int64_t n, int64_t hw, int64_t c,
                                float* partial, int64_t partial_ld, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && partial && n > 0 && n <= 65535 && hw > 0 && hw % 32 == 0 && c > 0 && c % 8 == 0 &&
                    ld_x % 4 == 0 && partial_ld >= c / 8,
                "gn_partial: bad arguments (hw %% 32, c %% 8, ld_x %% 4)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>(hw / 32), static_cast<unsigned>(n));
  if (dtype == EALDM_F32)
    norm::gn_partial_kernel<float><<<grid, norm::NT, 0, st>>>(reinterpret_cast<const float*>(x), ld_x, (int)hw, (int)c,
                                                              reinterpret_cast<float2*>(partial), partial_ld);
  else
    norm::gn_partial_kernel<bf16><<<grid, norm::NT, 0, st>>>(reinterpret_cast<const bf16*>(x), ld_x, (int)hw, (int)c,
                                                             reinterpret_cast<float2*>(partial), partial_ld);
  EALDM_LAUNCH_CHECK();
  return 0;
}
