// Implicit-GEMM convolution / linear on the 5th-generation tensor cores (tcgen05) of sm_100a.
//
//   out[M, N] = epilogue( A[M, K] * W[N, K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// A is never materialised: for every 64-channel K block the producer thread issues ONE 4-D TMA box
// load over the NHWC activation tensor (C, W, H, N) whose spatial origin is shifted by the filter
// tap; out-of-range rows/columns are zero-filled by the TMA unit, which is exactly the conv zero
// padding.  The box lands in shared memory as 128 pixel rows x 128 B with the 128-byte swizzle the
// UMMA K-major descriptor expects.  Stride-2 convolutions use the tensor map's element strides.
// W tiles are 2-D TMA loads of the packed [N, K] weight matrix.
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0      TMA producer (one elected lane)         smem ring of STAGES x (A 16 KiB + B BN*128 B)
//   warp 1      TMEM allocator + tcgen05.mma issuer     2 accumulator stages of BN fp32 columns
//   warps 2..5  epilogue: tcgen05.ld -> bias / per-image row vector / SiLU / GEGLU / residual -> global
// so the epilogue of tile i overlaps the main loop of tile i+1.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace ealdm {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 per K block = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

struct Segment {
  int kblocks;  // taps * cblk
  int cblk;     // channels / 64
  int ksize;    // 1 or 3
  int pad;
  int stride;
  int bkoff;  // first K column of this segment inside W
};

struct Params {
  int nseg;
  Segment seg[2];
  int bw, bh, bn;  // output-space box: bw*bh*bn == 128
  int tiles_w, tiles_h;
  int Wout, Hout, Nimg;
  int N;  // logical accumulator columns
  int m_tiles, n_tiles;
  Epilogue ep;
};

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
};

// ---- epilogue for one row x 32 accumulator columns ------------------------------------------------
template <typename TOut>
__device__ __forceinline__ void store8(TOut* dst, const float (&o)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* dst, const float (&o)[8]) {
  reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
  reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
}
template <>
__device__ __forceinline__ void store8<bf16>(bf16* dst, const float (&o)[8]) {
  uint4 u;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(o[0], o[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(o[2], o[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(o[4], o[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(o[6], o[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
  *reinterpret_cast<uint4*>(dst) = u;
}

template <typename TOut>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&acc)[32], const Epilogue& ep,
                                               long long row, long long img, int col0, int N) {
  const bool vec_ok = (col0 + 32 <= N);
  if (ep.act == EALDM_ACT_GEGLU) {
    // columns [0,16) are values, [16,32) their gates; output column = col0/2 + j
    const int ocol0 = col0 >> 1;
    TOut* dst = reinterpret_cast<TOut*>(ep.out) + row * ep.ld_out + ocol0;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = __uint_as_float(acc[g * 8 + j]);
        float b = __uint_as_float(acc[16 + g * 8 + j]);
        if (ep.bias) {
          a += __ldg(ep.bias + col0 + g * 8 + j);
          b += __ldg(ep.bias + col0 + 16 + g * 8 + j);
        }
        o[j] = a * gelu_erf_f(b);
      }
      if (ep.residual) {
        if (ep.res_f32) {
          const float* r = reinterpret_cast<const float*>(ep.residual) + row * ep.ld_res + ocol0 + g * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        } else {
          const bf16* r = reinterpret_cast<const bf16*>(ep.residual) + row * ep.ld_res + ocol0 + g * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += __bfloat162float(r[j]);
        }
      }
      store8<TOut>(dst + g * 8, o);
    }
    return;
  }
  TOut* dst = reinterpret_cast<TOut*>(ep.out) + row * ep.ld_out + col0;
  const float* rv = ep.rowvec ? ep.rowvec + img * ep.ld_rowvec + col0 : nullptr;
  const bf16* res = (ep.residual && !ep.res_f32)
                        ? reinterpret_cast<const bf16*>(ep.residual) + row * ep.ld_res + col0
                        : nullptr;
  const float* resf = (ep.residual && ep.res_f32)
                          ? reinterpret_cast<const float*>(ep.residual) + row * ep.ld_res + col0
                          : nullptr;
  if (vec_ok) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(acc[g * 8 + j]);
      if (ep.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + g * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + g * 8 + 4));
        o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
        o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
      }
      if (rv) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(rv + g * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(rv + g * 8 + 4));
        o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
        o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
      }
      if (ep.act == EALDM_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = silu_f(o[j]);
      }
      if (res) {
        const uint4 u = *reinterpret_cast<const uint4*>(res + g * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[2 * j] += __low2float(h[j]);
          o[2 * j + 1] += __high2float(h[j]);
        }
      }
      if (resf) {
        const float4 r0 = *reinterpret_cast<const float4*>(resf + g * 8);
        const float4 r1 = *reinterpret_cast<const float4*>(resf + g * 8 + 4);
        o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w;
        o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
      }
      store8<TOut>(dst + g * 8, o);
      if (ep.out2) store8<bf16>(reinterpret_cast<bf16*>(ep.out2) + row * ep.ld_out2 + col0 + g * 8, o);
    }
  } else {
    for (int j = 0; j < 32; ++j) {
      if (col0 + j >= N) break;
      float o = __uint_as_float(acc[j]);
      if (ep.bias) o += __ldg(ep.bias + col0 + j);
      if (rv) o += __ldg(rv + j);
      if (ep.act == EALDM_ACT_SILU) o = silu_f(o);
      if (res) o += __bfloat162float(res[j]);
      if (resf) o += resf[j];
      dst[j] = from_f32<TOut>(o);
      if (ep.out2) reinterpret_cast<bf16*>(ep.out2)[row * ep.ld_out2 + col0 + j] = __float2bfloat16_rn(o);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA0);
    if (p.nseg > 1) ptx::prefetch_tensormap(&tmA1);
    ptx::prefetch_tensormap(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int kblocks_total = p.seg[0].kblocks + (p.nseg > 1 ? p.seg[1].kblocks : 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles;
        const int nt = tile - mt * p.n_tiles;
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        for (int s = 0; s < p.nseg; ++s) {
          const Segment sg = p.seg[s];
          const CUtensorMap* tmA = (s == 0) ? &tmA0 : &tmA1;
          int tap = 0, cb = 0;
          for (int kb = 0; kb < sg.kblocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            const int kh = tap / sg.ksize;
            const int kw = tap - kh * sg.ksize;
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            ptx::tma_load_4d(sa, tmA, &full_bar[stage], cb * BK, w0 * sg.stride + kw - sg.pad,
                             h0 * sg.stride + kh - sg.pad, n0);
            ptx::tma_load_2d(sa + C::A_BYTES, &tmB, &full_bar[stage], sg.bkoff + kb * BK, nt * BN);
            if (++cb == sg.cblk) { cb = 0; ++tap; }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < kblocks_total; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t adesc = ptx::make_sw128_kmajor_desc(sa);
          const uint64_t bdesc = ptx::make_sw128_kmajor_desc(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advancing K by 16 bf16 = 32 B inside the swizzle row = +2 in 16-byte units
            ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (kb == kblocks_total - 1) ptx::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;
    const int bw_i = r % p.bw;
    const int bh_i = (r / p.bw) % p.bh;
    const int bn_i = r / (p.bw * p.bh);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles;
      const int nt = tile - mt * p.n_tiles;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int w = tw * p.bw + bw_i, h = th * p.bh + bh_i, n = tn * p.bn + bn_i;
      const bool valid = (w < p.Wout) && (h < p.Hout) && (n < p.Nimg);
      const long long row = (static_cast<long long>(n) * p.Hout + h) * p.Wout + w;

      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr0 + ch * 32, v);
        ptx::tmem_ld_wait();
        const int col0 = nt * BN + ch * 32;
        if (valid && col0 < p.N) {
          if (p.ep.out_f32)
            epilogue_chunk<float>(v, p.ep, row, n, col0, p.N);
          else
            epilogue_chunk<bf16>(v, p.ep, row, n, col0, p.N);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---- host side --------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int pow2_ceil(long long v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int BN>
static int launch_bn(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                     const Params& p, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    EALDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  conv_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(a0, a1, b, p);
  EALDM_LAUNCH_CHECK();
  return 0;
}

// returns 1 if the tcgen05 path can run this problem
bool supported(const ealdm_conv_args* a) {
  if (a->dtype != EALDM_BF16) return false;
  if (a->n_src < 1 || a->n_src > 2) return false;
  for (int s = 0; s < a->n_src; ++s) {
    const ealdm_conv_src& x = a->src[s];
    if (x.c % BK != 0 || x.ld % 8 != 0) return false;
    if ((reinterpret_cast<uintptr_t>(x.x) & 15) != 0) return false;
    if (x.upsample) return false;
    if (x.ksize != 1 && x.ksize != 3) return false;
    if (x.stride != 1 && x.stride != 2) return false;
    if (x.n != a->src[0].n) return false;
  }
  if (a->k_total % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(a->weight) & 15) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(a->out) & 15) != 0 || a->ld_out % 8 != 0) return false;
  if (a->residual && ((reinterpret_cast<uintptr_t>(a->residual) & 15) != 0 || a->ld_res % 8 != 0))
    return false;
  if (a->act == EALDM_ACT_GEGLU && (a->rowvec || a->out2)) return false;
  if (a->out2 && ((reinterpret_cast<uintptr_t>(a->out2) & 15) != 0 || a->ld_out2 % 8 != 0)) return false;
  if (a->act == EALDM_ACT_GEGLU && a->n_out % 32 != 0) return false;
  return true;
}

int launch(const ealdm_conv_args* a, cudaStream_t st) {
  EALDM_REQUIRE(supported(a), "tcgen05 conv: unsupported shape/alignment (c%%64, ld%%8, 16 B pointers)");
  PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode();
  if (!encode) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled entry point not available");

  Params p;
  memset(&p, 0, sizeof(p));
  p.nseg = a->n_src;
  p.Wout = static_cast<int>(a->w_out);
  p.Hout = static_cast<int>(a->h_out);
  p.Nimg = static_cast<int>(a->src[0].n);
  p.N = static_cast<int>(a->n_out);
  p.bw = pow2_ceil(a->w_out) < BM ? pow2_ceil(a->w_out) : BM;
  p.bh = pow2_ceil(a->h_out) < BM / p.bw ? pow2_ceil(a->h_out) : BM / p.bw;
  p.bn = BM / (p.bw * p.bh);
  p.tiles_w = static_cast<int>(ceil_div(a->w_out, p.bw));
  p.tiles_h = static_cast<int>(ceil_div(a->h_out, p.bh));
  p.m_tiles = p.tiles_w * p.tiles_h * static_cast<int>(ceil_div(p.Nimg, p.bn));

  // choose the N tile: fewest (waves x tile cost)
  int BN;
  if (a->n_out <= 32) {
    BN = 32;
  } else if (a->n_out <= 128) {
    BN = 128;
  } else {
    const long long t256 = static_cast<long long>(p.m_tiles) * ceil_div(a->n_out, 256);
    const long long t128 = static_cast<long long>(p.m_tiles) * ceil_div(a->n_out, 128);
    const long long c256 = ceil_div(t256, num_sms()) * (256 + 48);
    const long long c128 = ceil_div(t128, num_sms()) * (128 + 48);
    BN = (c256 <= c128) ? 256 : 128;
  }
  p.n_tiles = static_cast<int>(ceil_div(a->n_out, BN));

  CUtensorMap tmA[2], tmB;
  memset(tmA, 0, sizeof(tmA));
  int koff = 0;
  for (int s = 0; s < a->n_src; ++s) {
    const ealdm_conv_src& x = a->src[s];
    Segment& sg = p.seg[s];
    sg.cblk = static_cast<int>(x.c / BK);
    sg.ksize = x.ksize;
    sg.kblocks = sg.cblk * x.ksize * x.ksize;
    sg.pad = x.pad;
    sg.stride = x.stride;
    sg.bkoff = koff;
    koff += static_cast<int>(x.c) * x.ksize * x.ksize;
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(x.c), static_cast<cuuint64_t>(x.w),
                          static_cast<cuuint64_t>(x.h), static_cast<cuuint64_t>(x.n)};
    cuuint64_t gstr[3] = {static_cast<cuuint64_t>(x.ld) * 2,
                          static_cast<cuuint64_t>(x.ld) * 2 * static_cast<cuuint64_t>(x.w),
                          static_cast<cuuint64_t>(x.ld) * 2 * static_cast<cuuint64_t>(x.w) *
                              static_cast<cuuint64_t>(x.h)};
    cuuint32_t box[4] = {BK, static_cast<cuuint32_t>(p.bw * x.stride),
                         static_cast<cuuint32_t>(p.bh * x.stride), static_cast<cuuint32_t>(p.bn)};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(x.stride), static_cast<cuuint32_t>(x.stride),
                          1};
    CUresult r = encode(&tmA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x.x), gdim,
                        gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(A%d) failed: %d", s, (int)r);
  }
  EALDM_REQUIRE(koff == a->k_total, "k_total %lld does not match the sources (%d)",
                (long long)a->k_total, koff);
  if (a->n_src == 1) tmA[1] = tmA[0];
  {
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(a->k_total), static_cast<cuuint64_t>(a->n_out)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(a->k_total) * 2};
    cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(BN)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->weight),
                        gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
  }

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

  switch (BN) {
    case 32: return launch_bn<32>(tmA[0], tmA[1], tmB, p, st);
    case 128: return launch_bn<128>(tmA[0], tmA[1], tmB, p, st);
    default: return launch_bn<256>(tmA[0], tmA[1], tmB, p, st);
  }
}

}  // namespace tc
}  // namespace ealdm
