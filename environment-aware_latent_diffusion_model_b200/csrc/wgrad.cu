// Weight gradient of an implicit-GEMM convolution / linear layer:
//
//   dW[j, (kh, kw, ci)] = sum over output pixels m = (n, oh, ow) of
//                         dY[m, j] * X[n, oh*stride + kh - pad, ow*stride + kw - pad, ci]
//
// i.e. a GEMM whose reduction dimension is the PIXEL index.  Both operands are NHWC, so the reduction
// index is the slow one in memory: exactly the "MN-major" operand form of tcgen05.  A [64 pixel x 64
// channel] TMA box (128-byte rows, 128-byte swizzle) lands in shared memory as the canonical MN-major
// SWIZZLE_128B atom: 64 channels contiguous, 8 pixel rows per 1024-byte group (SBO), the next 64
// channels LBO bytes further.  No transposed copies of activations or gradients are ever made.
// The X box is shifted by the filter tap and zero-filled by the TMA unit (= conv zero padding);
// stride-2 convolutions use the tensor map's element strides, as in the forward kernel.
//
// Work decomposition: tile = 128 output channels x (one tap, BN input channels); the pixel range is
// split `splits` ways so that tiles*splits covers the 148 SMs; every (tile, split) writes its fp32
// partial to the workspace with plain stores and a second kernel sums the splits in a fixed order
// (deterministic, no atomics) while permuting to the requested weight layout (packed K-major or
// PyTorch OIHW) and optionally accumulating into the destination (.grad semantics).
//
// Persistent warp-specialised CTA: warp 0 TMA producer, warp 1 MMA issuer (2 TMEM accumulator stages),
// warps 2..5 epilogue (TMEM lane quadrant each: one thread = one output-channel row).
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace ealdm {
namespace wgrad {

constexpr int BM = 128;   // output channels per tile
constexpr int KP = 64;    // pixels per k-block
constexpr int UMMA_K = 16;
constexpr int BOX_BYTES = KP * 128;  // one [64 px x 64 ch] bf16 box
constexpr int NUM_THREADS = 64 + 128;

struct Params {
  int bw, bh, bn;            // pixel box: bw*bh*bn == KP
  int tiles_w, tiles_h, tiles_n;
  int kblocks;               // tiles_w*tiles_h*tiles_n
  int splits, kb_per_split;
  int m_tiles;               // ceil(n_out / 128)
  int cblocks;               // ceil(c / BN)
  int taps, ksize, pad, stride;
  int n_out, c;
  long long ws_ld;           // taps * c
  float* ws;                 // [splits][n_out][ws_ld]
};

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = 2 * BOX_BYTES;
  static constexpr int B_BYTES = (BN / 64) * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN >= 256 ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256;
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

using ptx::make_sw128_mnmajor_desc;
// kind::f16, D = f32, A = B = bf16, both MN-major
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ Params p) {
  using C = Cfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("ealdm: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    ptx::prefetch_tensormap(&tmDy);
    ptx::prefetch_tensormap(&tmX);
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

  const int n_tiles = p.taps * p.cblocks;
  const int total = p.m_tiles * n_tiles * p.splits;
  // unit -> (split, m tile, tap, channel block); splits are the fastest index so that the CTAs working
  // on one weight tile run concurrently and share the dY / X tiles of neighbouring taps through L2
  auto decode = [&](int unit, int& sp, int& mt, int& tap, int& cb) {
    sp = unit % p.splits;
    const int tile = unit / p.splits;
    mt = tile / n_tiles;
    const int nt = tile - mt * n_tiles;
    tap = nt / p.cblocks;
    cb = nt - tap * p.cblocks;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
        int sp, mt, tap, cb;
        decode(unit, sp, mt, tap, cb);
        const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int tw = kb % p.tiles_w;
          const int th = (kb / p.tiles_w) % p.tiles_h;
          const int tn = kb / (p.tiles_w * p.tiles_h);
          const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
#pragma unroll
          for (int g = 0; g < 2; ++g)
            ptx::tma_load_4d(sa + g * BOX_BYTES, &tmDy, &full_bar[stage], mt * BM + g * 64, w0, h0, n0);
#pragma unroll
          for (int g = 0; g < BN / 64; ++g)
            ptx::tma_load_4d(sa + C::A_BYTES + g * BOX_BYTES, &tmX, &full_bar[stage], cb * BN + g * 64,
                             w0 * p.stride + kw - p.pad, h0 * p.stride + kh - p.pad, n0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16_mn(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
      int sp, mt, tap, cb;
      decode(unit, sp, mt, tap, cb);
      const int kb0 = sp * p.kb_per_split;
      const int nkb = min(kb0 + p.kb_per_split, p.kblocks) - kb0;
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t adesc = make_sw128_mnmajor_desc(sa, BOX_BYTES);
          const uint64_t bdesc = make_sw128_mnmajor_desc(sa + C::A_BYTES, BOX_BYTES);
#pragma unroll
          for (int k = 0; k < KP / UMMA_K; ++k) {
            // 16 pixel rows further = 16 * 128 B = 2048 B = +128 in 16-byte units
            ptx::umma_bf16(d_tmem, adesc + 128 * k, bdesc + 128 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (kb == nkb - 1) ptx::umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
      int sp, mt, tap, cb;
      decode(unit, sp, mt, tap, cb);
      const int row = mt * BM + quad * 32 + lane;
      float* dst = p.ws + (static_cast<long long>(sp) * p.n_out + row) * p.ws_ld +
                   static_cast<long long>(tap) * p.c + cb * BN;
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int ku = 0; ku < BN / 32; ++ku) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr0 + ku * 32, v);
        ptx::tmem_ld_wait();
        if (row < p.n_out) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = cb * BN + ku * 32 + 4 * j;
            if (col < p.c)  // c % 4 == 0
              *reinterpret_cast<float4*>(dst + ku * 32 + 4 * j) =
                  make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                              __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
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

// ---- split reduction + layout permutation -----------------------------------------------------------
// dst[j][perm(k)] (+)= sum_s ws[s][j][k], k = tap*c + ci.  layout 0: perm = identity (packed K-major,
// pitch ld_dw); layout 1: PyTorch OIHW, perm(k) = ci*taps + tap.
// One CTA per (output channel j, 64-channel block): the partials are read tap by tap along ci (coalesced),
// transposed through shared memory, and the [64 ci x taps] block of dw -- contiguous in OIHW -- is updated
// with coalesced accesses.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, int splits, long long n_out, long long ktot, int c, int taps,
                    float* __restrict__ dw, long long ld_dw, int layout, int accumulate) {
  __shared__ float tile[9][65];
  const long long j = blockIdx.y;
  const int ci0 = blockIdx.x * 64;
  const int nci = min(64, c - ci0);
  const long long total = n_out * ktot;
  const float* src = ws + j * ktot;
  for (int i = threadIdx.x; i < taps * 64; i += 256) {
    const int tap = i >> 6, cl = i & 63;
    float s = 0.f;
    if (cl < nci) {
      const long long o = static_cast<long long>(tap) * c + ci0 + cl;
      for (int sp = 0; sp < splits; ++sp) s += src[sp * total + o];
      if (layout == 0) {
        float* d = dw + j * ld_dw + o;
        *d = accumulate ? *d + s : s;
      }
    }
    tile[tap][cl] = s;
  }
  if (layout == 0) return;
  __syncthreads();
  float* d = dw + j * ld_dw + static_cast<long long>(ci0) * taps;
  for (int i = threadIdx.x; i < nci * taps; i += 256) {
    const int cl = i / taps, tap = i - cl * taps;
    d[i] = accumulate ? d[i] + tile[tap][cl] : tile[tap][cl];
  }
}

// identity permutation (1x1 / linear layers, or the packed layout): 4 consecutive k per thread
__global__ void __launch_bounds__(256)
wgrad_reduce_flat_kernel(const float* __restrict__ ws, int splits, long long n_out, long long ktot,
                         float* __restrict__ dw, long long ld_dw, int accumulate) {
  const long long total = n_out * ktot;
  const long long kv = ktot >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n_out * kv; i += 256LL * gridDim.x) {
    const long long j = i / kv;
    const long long k = (i - j * kv) << 2;
    float4 s = *reinterpret_cast<const float4*>(ws + j * ktot + k);
    for (int sp = 1; sp < splits; ++sp) {
      const float4 e = *reinterpret_cast<const float4*>(ws + sp * total + j * ktot + k);
      s.x += e.x; s.y += e.y; s.z += e.z; s.w += e.w;
    }
    float4* d = reinterpret_cast<float4*>(dw + j * ld_dw + k);
    if (accumulate) {
      const float4 o = *d;
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    *d = s;
  }
}

// ---- fp32 / generic SIMT weight gradient (parity mode and shapes the tensor-core kernel refuses) -----
// 32 x 32 output tile per CTA, 32 pixels per step through shared memory, fp32 accumulation.
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const T* __restrict__ x, long long ld_x, int n, int h, int w, int c, int ksize, int stride,
                  int pad, const T* __restrict__ dy, long long ld_dy, int n_out, int h_out, int w_out,
                  float* __restrict__ ws, long long ws_ld, int splits) {
  __shared__ float As[32][33];  // [pixel][co]
  __shared__ float Bs[32][33];  // [pixel][k col]
  const int t = threadIdx.x;
  const int tx = t & 31, ty = t >> 5;  // loader: row ty + 8*i, col tx
  const long long ktot = static_cast<long long>(ksize) * ksize * c;
  const int k0 = blockIdx.x * 32, j0 = blockIdx.y * 32, sp = blockIdx.z;
  const long long M = static_cast<long long>(n) * h_out * w_out;
  const long long per = (M + splits - 1) / splits;
  const long long m_begin = sp * per, m_end = min(M, m_begin + per);
  // column this thread loads for B
  const long long kcol = k0 + tx;
  const bool kvalid = kcol < ktot;
  int tap = 0, ci = 0, kh = 0, kw = 0;
  if (kvalid) {
    tap = static_cast<int>(kcol / c);
    ci = static_cast<int>(kcol - static_cast<long long>(tap) * c);
    kh = tap / ksize;
    kw = tap - kh * ksize;
  }
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // outputs (co = ty*4 + i, col = tx)
  for (long long m0 = m_begin; m0 < m_end; m0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      const long long m = m0 + r;
      float a = 0.f, b = 0.f;
      if (m < m_end) {
        if (j0 + tx < n_out) a = to_f32(dy[m * ld_dy + j0 + tx]);
        if (kvalid) {
          const int ow = static_cast<int>(m % w_out);
          const int oh = static_cast<int>((m / w_out) % h_out);
          const int img = static_cast<int>(m / (static_cast<long long>(w_out) * h_out));
          const int ih = oh * stride + kh - pad, iw = ow * stride + kw - pad;
          if (ih >= 0 && ih < h && iw >= 0 && iw < w)
            b = to_f32(x[((static_cast<long long>(img) * h + ih) * w + iw) * ld_x + ci]);
        }
      }
      As[r][tx] = a;
      Bs[r][tx] = b;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float b = Bs[r][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(As[r][ty * 4 + i], b, acc[i]);
    }
    __syncthreads();
  }
  if (kvalid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = j0 + ty * 4 + i;
      if (j < n_out) ws[(static_cast<long long>(sp) * n_out + j) * ws_ld + kcol] = acc[i];
    }
  }
}

// ---- host ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
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
static bool aligned16(const void* ptr, long long ld, int es) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * es) % 16 == 0;
}

static bool tc_supported(const ealdm_conv_wgrad_args* a) {
  if (a->dtype != EALDM_BF16) return false;
  const ealdm_conv_src& x = a->src;
  if (x.upsample) return false;
  if (x.c % 4 != 0) return false;
  return aligned16(x.x, x.ld, 2) && aligned16(a->dy, a->ld_dy, 2);
}

struct Plan {
  bool tc;
  int bn_tile;  // BN
  int splits;
  Params p;
};

static void make_plan(const ealdm_conv_wgrad_args* a, Plan* pl) {
  const ealdm_conv_src& x = a->src;
  Params& p = pl->p;
  memset(&p, 0, sizeof(p));
  pl->tc = (a->impl != EALDM_IMPL_SIMT) && tc_supported(a);
  p.ksize = x.ksize;
  p.taps = x.ksize * x.ksize;
  p.pad = x.pad;
  p.stride = x.stride;
  p.n_out = static_cast<int>(a->n_out);
  p.c = static_cast<int>(x.c);
  p.ws_ld = static_cast<long long>(p.taps) * p.c;
  const long long M = x.n * a->h_out * a->w_out;
  if (pl->tc) {
    p.bw = pow2_ceil(a->w_out) < KP ? pow2_ceil(a->w_out) : KP;
    p.bh = pow2_ceil(a->h_out) < KP / p.bw ? pow2_ceil(a->h_out) : KP / p.bw;
    p.bn = KP / (p.bw * p.bh);
    p.tiles_w = static_cast<int>(ceil_div(a->w_out, p.bw));
    p.tiles_h = static_cast<int>(ceil_div(a->h_out, p.bh));
    p.tiles_n = static_cast<int>(ceil_div(x.n, p.bn));
    p.kblocks = p.tiles_w * p.tiles_h * p.tiles_n;
    pl->bn_tile = p.c > 128 ? 256 : (p.c > 64 ? 128 : 64);
    p.m_tiles = static_cast<int>(ceil_div(a->n_out, BM));
    p.cblocks = static_cast<int>(ceil_div(p.c, pl->bn_tile));
    const long long tiles = static_cast<long long>(p.m_tiles) * p.taps * p.cblocks;
    // enough splits to cover the SMs, but at least 8 k-blocks of work per unit
    long long sp = ceil_div(num_sms(), tiles);
    const long long max_sp = p.kblocks / 8 > 0 ? p.kblocks / 8 : 1;
    if (sp > max_sp) sp = max_sp;
    if (sp < 1) sp = 1;
    p.kb_per_split = static_cast<int>(ceil_div(p.kblocks, sp));
    p.splits = static_cast<int>(ceil_div(p.kblocks, p.kb_per_split));
  } else {
    const long long tiles = ceil_div(p.ws_ld, 32) * ceil_div(a->n_out, 32);
    long long sp = ceil_div(4LL * num_sms(), tiles);
    const long long max_sp = M / 256 > 0 ? M / 256 : 1;
    if (sp > max_sp) sp = max_sp;
    if (sp < 1) sp = 1;
    p.splits = static_cast<int>(sp);
  }
  pl->splits = p.splits;
}

template <int BN>
static int launch_tc(const CUtensorMap& tmDy, const CUtensorMap& tmX, const Params& p, cudaStream_t st) {
  using C = Cfg<BN>;
  static DeviceOnce attr_set;
  if (attr_set.pending()) {
    EALDM_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES));
    attr_set.done();
  }
  const int total = p.m_tiles * p.taps * p.cblocks * p.splits;
  const int grid = total < num_sms() ? total : num_sms();
  wgrad_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(tmDy, tmX, p);
  EALDM_LAUNCH_CHECK();
  return 0;
}

static int encode_px_map(CUtensorMap* tm, const void* base, long long c, long long w, long long h, long long n,
                         long long ld, int bw, int bh, int bn, int stride) {
  PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode();
  if (!encode) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * gdim[1],
                        static_cast<cuuint64_t>(ld) * 2 * gdim[1] * gdim[2]};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw * stride), static_cast<cuuint32_t>(bh * stride),
                       static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(wgrad) failed: %d", (int)r);
  return 0;
}

}  // namespace wgrad
}  // namespace ealdm

using namespace ealdm;

static int wgrad_validate(const ealdm_conv_wgrad_args* a) {
  EALDM_REQUIRE(a != nullptr, "conv_wgrad: null args");
  EALDM_REQUIRE(a->dtype == EALDM_F32 || a->dtype == EALDM_BF16, "conv_wgrad: bad dtype %d", a->dtype);
  const ealdm_conv_src& x = a->src;
  EALDM_REQUIRE(x.n > 0 && x.h > 0 && x.w > 0 && x.c > 0 && x.ld >= x.c, "conv_wgrad: bad source dims");
  EALDM_REQUIRE(x.ksize == 1 || x.ksize == 3, "conv_wgrad: ksize must be 1 or 3");
  EALDM_REQUIRE(x.stride == 1 || x.stride == 2, "conv_wgrad: stride must be 1 or 2");
  EALDM_REQUIRE(x.pad >= 0 && x.pad <= 1 && x.upsample == 0, "conv_wgrad: pad must be 0/1, upsample 0");
  EALDM_REQUIRE(a->n_out > 0 && a->h_out > 0 && a->w_out > 0 && a->ld_dy >= a->n_out, "conv_wgrad: bad output dims");
  EALDM_REQUIRE(a->layout == EALDM_WGRAD_PACKED || a->layout == EALDM_WGRAD_OIHW, "conv_wgrad: bad layout");
  return 0;
}

extern "C" int64_t ealdm_conv_wgrad_workspace_bytes(const ealdm_conv_wgrad_args* a) {
  if (wgrad_validate(a) != 0) return -1;
  wgrad::Plan pl;
  wgrad::make_plan(a, &pl);
  return static_cast<int64_t>(pl.splits) * a->n_out * pl.p.ws_ld * 4;
}

extern "C" int ealdm_conv_wgrad(const ealdm_conv_wgrad_args* a, ealdm_stream_t stream) {
  if (int e = wgrad_validate(a)) return e;
  EALDM_REQUIRE(a->src.x && a->dy && a->dw && a->workspace, "conv_wgrad: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  wgrad::Plan pl;
  wgrad::make_plan(a, &pl);
  const ealdm_conv_src& x = a->src;
  const long long need = static_cast<long long>(pl.splits) * a->n_out * pl.p.ws_ld * 4;
  EALDM_REQUIRE(a->workspace_bytes >= need, "conv_wgrad: workspace too small (%lld < %lld)",
                (long long)a->workspace_bytes, need);
  EALDM_REQUIRE(a->impl != EALDM_IMPL_TCGEN05 || pl.tc, "conv_wgrad: tcgen05 path needs bf16, 16-byte rows, c %% 4 == 0");
  const long long ktot = pl.p.ws_ld;
  EALDM_REQUIRE(a->ld_dw >= ktot, "conv_wgrad: ld_dw smaller than ksize^2 * c");
  pl.p.ws = reinterpret_cast<float*>(a->workspace);
  if (pl.tc) {
    CUtensorMap tmDy, tmX;
    if (int e = wgrad::encode_px_map(&tmDy, a->dy, a->n_out, a->w_out, a->h_out, x.n, a->ld_dy, pl.p.bw, pl.p.bh,
                                     pl.p.bn, 1))
      return e;
    if (int e = wgrad::encode_px_map(&tmX, x.x, x.c, x.w, x.h, x.n, x.ld, pl.p.bw, pl.p.bh, pl.p.bn, x.stride))
      return e;
    int e;
    switch (pl.bn_tile) {
      case 64: e = wgrad::launch_tc<64>(tmDy, tmX, pl.p, st); break;
      case 128: e = wgrad::launch_tc<128>(tmDy, tmX, pl.p, st); break;
      default: e = wgrad::launch_tc<256>(tmDy, tmX, pl.p, st); break;
    }
    if (e) return e;
  } else {
    dim3 grid(static_cast<unsigned>(ceil_div(ktot, 32)), static_cast<unsigned>(ceil_div(a->n_out, 32)),
              static_cast<unsigned>(pl.splits));
    if (a->dtype == EALDM_F32)
      wgrad::wgrad_simt_kernel<float><<<grid, 256, 0, st>>>(
          reinterpret_cast<const float*>(x.x), x.ld, (int)x.n, (int)x.h, (int)x.w, (int)x.c, x.ksize, x.stride,
          x.pad, reinterpret_cast<const float*>(a->dy), a->ld_dy, (int)a->n_out, (int)a->h_out, (int)a->w_out,
          pl.p.ws, ktot, pl.splits);
    else
      wgrad::wgrad_simt_kernel<bf16><<<grid, 256, 0, st>>>(
          reinterpret_cast<const bf16*>(x.x), x.ld, (int)x.n, (int)x.h, (int)x.w, (int)x.c, x.ksize, x.stride,
          x.pad, reinterpret_cast<const bf16*>(a->dy), a->ld_dy, (int)a->n_out, (int)a->h_out, (int)a->w_out,
          pl.p.ws, ktot, pl.splits);
    EALDM_LAUNCH_CHECK();
  }
  if ((pl.p.taps == 1 || a->layout == EALDM_WGRAD_PACKED) && ktot % 4 == 0 && a->ld_dw % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(a->dw) & 15) == 0) {
    const long long work = a->n_out * (ktot / 4);
    const int blocks = static_cast<int>(ceil_div(work, 256) < 2368 ? ceil_div(work, 256) : 2368);
    wgrad::wgrad_reduce_flat_kernel<<<blocks, 256, 0, st>>>(pl.p.ws, pl.splits, a->n_out, ktot, a->dw, a->ld_dw,
                                                            a->accumulate);
    EALDM_LAUNCH_CHECK();
    return 0;
  }
  EALDM_REQUIRE(a->n_out <= 65535, "conv_wgrad: n_out too large");
  dim3 rgrid(static_cast<unsigned>(ceil_div(pl.p.c, 64)), static_cast<unsigned>(a->n_out));
  wgrad::wgrad_reduce_kernel<<<rgrid, 256, 0, st>>>(pl.p.ws, pl.splits, a->n_out, ktot, pl.p.c, pl.p.taps, a->dw,
                                                    a->ld_dw, a->layout, a->accumulate);
  EALDM_LAUNCH_CHECK();
  return 0;
}
