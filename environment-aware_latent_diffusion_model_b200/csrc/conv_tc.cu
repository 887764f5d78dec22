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
//   warps 2..9  epilogue: tcgen05.ld -> smem transpose -> bias / per-image row vector / SiLU / GEGLU /
//               (prefetched) residual -> coalesced global stores (+ optional bf16 shadow copy)
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
constexpr int EPI_WARPS = 8;                      // two per TMEM lane quadrant (even / odd 32-column chunks)
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int STG_STRIDE = 33;                    // floats per staged row (conflict-free transpose)
constexpr int STG_BYTES = 32 * STG_STRIDE * 4;    // per epilogue warp

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
  static constexpr int BAR_BYTES = 256;  // (2*STAGES + 4) mbarriers + the TMEM base slot
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_WARPS * STG_BYTES + 1024;
  static_assert((2 * STAGES + 4) * 8 + 4 <= BAR_BYTES, "barrier block too small");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB dynamic shared memory limit");
};

// ---- epilogue ----------------------------------------------------------------------------------------
// tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns).  Writing / reading
// global memory in that layout makes each warp-level access touch 32 different 128-byte lines, which
// is what bounded the small-K GEMMs of the transformer blocks (LSU wavefronts, not HBM).  The chunk is
// therefore transposed through a per-warp shared-memory tile so that 8 lanes cover 32 consecutive
// columns of one row: residual loads, output stores and the bf16 shadow stores are fully coalesced
// (4 rows x 128 B per warp instruction), and the residual of the NEXT chunk is prefetched into
// registers while the current one is processed (the first one before the accumulator is even ready).
// GELU for the bf16 tensor-core path: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below
// bf16 resolution) -- ~3x fewer instructions than erff in the epilogue's critical path.  The fp32
// parity path (conv_simt.cu) keeps the exact erff.
__device__ __forceinline__ float gelu_fast(float v) {
  const float x = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * __expf(-x * x);
  const float erf_v = copysignf(erf_abs, v);
  return 0.5f * v * (1.0f + erf_v);
}

struct RowSet {          // the 8 output rows this thread touches in the coalesced layout
  long long row[8];
  int img[8];
  unsigned valid;        // bit i: row i exists
};

// rare paths (N not a multiple of 4 at the right edge) are kept out of line to keep the hot code small
__device__ __noinline__ float4 ld_res4_edge(const Epilogue& ep, long long row, int col, int N) {
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  for (int e = 0; e < 4 && col + e < N; ++e)
    t[e] = ep.res_f32 ? reinterpret_cast<const float*>(ep.residual)[row * ep.ld_res + col + e]
                      : __bfloat162float(reinterpret_cast<const bf16*>(ep.residual)[row * ep.ld_res + col + e]);
  return make_float4(t[0], t[1], t[2], t[3]);
}

__device__ __forceinline__ float4 ld_res4(const Epilogue& ep, long long row, int col, int N, bool ok) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!ok || col >= N) return r;
  if (col + 4 > N) return ld_res4_edge(ep, row, col, N);
  if (ep.res_f32) {
    r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.residual) + row * ep.ld_res + col);
  } else {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(ep.residual) +
                                                    row * ep.ld_res + col);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    r = make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
  return r;
}

__device__ __forceinline__ uint2 pack4_bf16(const float4& v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  return u;
}

__device__ __noinline__ void st_out4_edge(const Epilogue& ep, long long row, int col, int N, float4 v) {
  const float t[4] = {v.x, v.y, v.z, v.w};
  for (int e = 0; e < 4 && col + e < N; ++e) {
    if (ep.out_f32)
      reinterpret_cast<float*>(ep.out)[row * ep.ld_out + col + e] = t[e];
    else
      reinterpret_cast<bf16*>(ep.out)[row * ep.ld_out + col + e] = __float2bfloat16_rn(t[e]);
    if (ep.out2) reinterpret_cast<bf16*>(ep.out2)[row * ep.ld_out2 + col + e] = __float2bfloat16_rn(t[e]);
  }
}

__device__ __forceinline__ void st_out4(const Epilogue& ep, long long row, int col, int N, const float4& v) {
  if (col + 4 > N) {
    st_out4_edge(ep, row, col, N, v);
    return;
  }
  if (ep.out_f32)
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + row * ep.ld_out + col) = v;
  else
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out) + row * ep.ld_out + col) = pack4_bf16(v);
  if (ep.out2)
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out2) + row * ep.ld_out2 + col) = pack4_bf16(v);
}

// 4 consecutive fp32 of a per-column vector (bias / per-image row vector), bounds-safe
__device__ __noinline__ float4 ld_vec4_edge(const float* p, int col, int N) {
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  for (int e = 0; e < 4 && col + e < N; ++e) t[e] = __ldg(p + col + e);
  return make_float4(t[0], t[1], t[2], t[3]);
}
__device__ __forceinline__ float4 ld_vec4(const float* p, int col, int N) {
  if (p == nullptr || col >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (col + 4 <= N) return __ldg(reinterpret_cast<const float4*>(p + col));
  return ld_vec4_edge(p, col, N);
}

// `b4` = bias (+ the per-image row vector when all rows of the warp belong to one image: rv_uniform)
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&acc)[32], float* stg, int lane,
                                               const Epilogue& ep, const RowSet& rs, const float4 (&res)[8],
                                               const float4& b4, bool rv_uniform, int col0, int N) {
#pragma unroll
  for (int j = 0; j < 32; ++j) stg[lane * STG_STRIDE + j] = __uint_as_float(acc[j]);
  __syncwarp();
  const int cc = (lane & 7) * 4;
  const int col = col0 + cc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + (lane >> 3);
    const float* sp = stg + rr * STG_STRIDE + cc;
    float4 v = make_float4(sp[0] + b4.x, sp[1] + b4.y, sp[2] + b4.z, sp[3] + b4.w);
    if (((rs.valid >> i) & 1u) && col < N) {
      if (ep.rowvec && !rv_uniform) {
        const float4 r4 = ld_vec4(ep.rowvec + static_cast<long long>(rs.img[i]) * ep.ld_rowvec, col, N);
        v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
      }
      if (ep.act == EALDM_ACT_SILU) {
        v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
      }
      v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w;
      st_out4(ep, rs.row[i], col, N, v);
    }
  }
  __syncwarp();
}

// GEGLU chunk: columns [0,16) are values, [16,32) their gates -> 16 output columns at col0/2.
// The raw accumulators are staged, then 4 lanes x 4 columns finish one row: bias pairs `bv`/`bg`
// (prefetched per tile) stay in registers.
__device__ __forceinline__ void epilogue_chunk_geglu(const uint32_t (&acc)[32], float* stg, int lane,
                                                     const Epilogue& ep, const long long (&row4)[4],
                                                     unsigned valid4, const float4& bv, const float4& bg,
                                                     int col0) {
#pragma unroll
  for (int j = 0; j < 32; ++j) stg[lane * STG_STRIDE + j] = __uint_as_float(acc[j]);
  __syncwarp();
  const int c4 = (lane & 3) * 4;
  const int oc = (col0 >> 1) + c4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = i * 8 + (lane >> 2);
    const float* sp = stg + rr * STG_STRIDE + c4;
    float4 v;
    v.x = (sp[0] + bv.x) * gelu_fast(sp[16] + bg.x);
    v.y = (sp[1] + bv.y) * gelu_fast(sp[17] + bg.y);
    v.z = (sp[2] + bv.z) * gelu_fast(sp[18] + bg.z);
    v.w = (sp[3] + bv.w) * gelu_fast(sp[19] + bg.w);
    if ((valid4 >> i) & 1u) {
      if (ep.residual) {
        const float4 r4 = ld_res4(ep, row4[i], oc, 1 << 30, true);
        v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
      }
      st_out4(ep, row4[i], oc, 1 << 30, v);
    }
  }
  __syncwarp();
}

template <int BN, bool GEGLU>
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
  uint8_t* stg_base = smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES;

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
      ptx::mbar_init(&tmem_empty[a], EPI_WARPS);
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
    // ===================== epilogue (warps 2..9) =====================
    const int quad = warp & 3;             // TMEM lane quadrant this warp may read
    const int part = (warp - 2) >> 2;      // 0: even 32-column chunks, 1: odd chunks
    float* stg = reinterpret_cast<float*>(stg_base + (warp - 2) * STG_BYTES);
    const bool has_res = p.ep.residual != nullptr;
    constexpr int NCH = (BN + 63) / 64;    // chunks per epilogue warp
    const int cc = (lane & 7) * 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles;
      const int nt = tile - mt * p.n_tiles;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      auto decode = [&](int r, long long& row, int& img) -> bool {
        const int w = tw * p.bw + r % p.bw;
        const int h = th * p.bh + (r / p.bw) % p.bh;
        const int n = tn * p.bn + r / (p.bw * p.bh);
        row = (static_cast<long long>(n) * p.Hout + h) * p.Wout + w;
        img = n;
        return (w < p.Wout) && (h < p.Hout) && (n < p.Nimg);
      };
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);

      if constexpr (GEGLU) {
        long long row4[4];
        unsigned valid4 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int img;
          if (decode(quad * 32 + i * 8 + (lane >> 2), row4[i], img)) valid4 |= 1u << i;
        }
        // (value, gate) bias pair of the first chunk, fetched before the accumulator is ready;
        // the next chunk's pair is prefetched while the current chunk is processed
        const int c4 = (lane & 3) * 4;
        float4 bv_n = ld_vec4(p.ep.bias, nt * BN + part * 32 + c4, p.N);
        float4 bg_n = ld_vec4(p.ep.bias, nt * BN + part * 32 + 16 + c4, p.N);
        ptx::mbar_wait(&tmem_full[acc], acc_phase);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int ch = part; ch < BN / 32; ch += 2) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr0 + ch * 32, v);
          const float4 bv = bv_n, bg = bg_n;
          const int col0 = nt * BN + ch * 32;
          if (ch + 2 < BN / 32) {
            bv_n = ld_vec4(p.ep.bias, col0 + 64 + c4, p.N);
            bg_n = ld_vec4(p.ep.bias, col0 + 64 + 16 + c4, p.N);
          }
          ptx::tmem_ld_wait();
          if (col0 < p.N) epilogue_chunk_geglu(v, stg, lane, p.ep, row4, valid4, bv, bg, col0);
        }
      } else {
        RowSet rs;
        rs.valid = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (decode(quad * 32 + i * 4 + (lane >> 3), rs.row[i], rs.img[i])) rs.valid |= 1u << i;
        // bias (+ per-image row vector when this thread's rows share one image) and residual of the
        // FIRST chunk are fetched before the accumulator is ready; those of the next chunk while the
        // current one is processed: no global-load latency between tcgen05.ld and the stores
        bool rv_uniform = false;
        const float* rv0 = nullptr;
        if (p.ep.rowvec) {
          rv_uniform = true;
          int img0 = 0;
          bool first = true;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if ((rs.valid >> i) & 1u) {
              if (first) { img0 = rs.img[i]; first = false; }
              else if (rs.img[i] != img0) rv_uniform = false;
            }
          }
          if (rv_uniform) rv0 = p.ep.rowvec + static_cast<long long>(img0) * p.ep.ld_rowvec;
        }
        auto load_bias = [&](int c0) -> float4 {
          float4 b = ld_vec4(p.ep.bias, c0, p.N);
          const float4 r4 = ld_vec4(rv0, c0, p.N);
          b.x += r4.x; b.y += r4.y; b.z += r4.z; b.w += r4.w;
          return b;
        };
        float4 b_n = load_bias(nt * BN + part * 32 + cc);
        float4 rnext[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rnext[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_res) {
          const int c0 = nt * BN + part * 32 + cc;
#pragma unroll
          for (int i = 0; i < 8; ++i) rnext[i] = ld_res4(p.ep, rs.row[i], c0, p.N, (rs.valid >> i) & 1u);
        }
        ptx::mbar_wait(&tmem_full[acc], acc_phase);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int ch = part; ch < BN / 32; ch += 2) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr0 + ch * 32, v);
          float4 rcur[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) rcur[i] = rnext[i];
          const float4 b4 = b_n;
          const int col0 = nt * BN + ch * 32;
          if (ch + 2 < BN / 32) {
            b_n = load_bias(col0 + 64 + cc);
            if (has_res) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                rnext[i] = ld_res4(p.ep, rs.row[i], col0 + 64 + cc, p.N, (rs.valid >> i) & 1u);
            }
          }
          ptx::tmem_ld_wait();
          if (col0 < p.N) epilogue_chunk(v, stg, lane, p.ep, rs, rcur, b4, rv_uniform, col0, p.N);
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

template <int BN, bool GEGLU>
static int launch_bn(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                     const Params& p, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    EALDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  conv_tc_kernel<BN, GEGLU><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(a0, a1, b, p);
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
  if (a->n_out <= 32 && a->act != EALDM_ACT_GEGLU) {
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

  const bool geglu = a->act == EALDM_ACT_GEGLU;
  switch (BN) {
    case 32: return launch_bn<32, false>(tmA[0], tmA[1], tmB, p, st);
    case 128:
      return geglu ? launch_bn<128, true>(tmA[0], tmA[1], tmB, p, st)
                   : launch_bn<128, false>(tmA[0], tmA[1], tmB, p, st);
    default:
      return geglu ? launch_bn<256, true>(tmA[0], tmA[1], tmB, p, st)
                   : launch_bn<256, false>(tmA[0], tmA[1], tmB, p, st);
  }
}

}  // namespace tc
}  // namespace ealdm
