// Fused GEGLU FeedForward of a transformer block at model width C = 256 (reference: FeedForward / GEGLU,
// ldm/modules/attention.py:37-64, and the residual add of BasicTransformerBlock._forward, attention.py:214):
//
//   out[M, 256] = ( value * gelu(gate) )[M, 1024] * W2^T + b2 + residual[M, 256],   [value | gate] = x[M, 256] * W1^T + b1
//
// as ONE kernel: the 8C-wide projection and the 4C-wide gated activation never leave the SM (unfused they are a
// 268 MB write plus a 268 MB read per layer at UNet batch 128, and two pipeline fills / drains).
//
// A CTA (one per SM, persistent over 128-row tiles) keeps its x tile [128 x 256] bf16 resident in shared memory and walks
// the hidden dimension in 16 chunks of 64:
//     D1[j & 1]  = x_tile * W1i[128 j .. 128 j + 128)^T          M=128, N=128, K=256   -> TMEM (two 128-column buffers)
//     H_s[j & 1] = GEGLU(D1 + b1) as bf16 [128 x 64]             epilogue warps, K-major SWIZZLE_128B = an A operand
//     D2        += H_s[j & 1] * W2[:, 64 j .. 64 j + 64)^T       M=128, N=2 x 128, K=64 -> TMEM (256 columns)
// and after the last chunk adds bias + residual to D2 and stores the tile.  W1 is the row-interleaved GEGLU matrix of
// packing.geglu_interleave (16 value rows then their 16 gate rows per block of 32), so that hidden column h = 16 b + i
// is accumulator column 32 b + i (value) and 32 b + 16 + i (gate), exactly as in conv_tc.cu's GEGLU epilogue.
//
// Warp roles: warp 0 = TMA producer (x tile; weights through a ring of 16 KB slots in the order the MMAs consume them),
// warp 1 = TMEM allocator + single-thread tcgen05.mma issuer (D1(j + 1) is issued BEFORE D2(j), so the tensor pipe works
// on the next chunk while the epilogue warps gate the current one), warps 2..9 = epilogue (two per TMEM lane quadrant).
// Arithmetic and accumulation order equal the unfused kernels' (same k-block order, same bias handling, same GEGLU
// polynomial), so the result is bit-identical to ealdm_conv(GEGLU) followed by ealdm_conv(+bias +residual).
#include "common.cuh"
#include "ptx.cuh"
#include "tc_math.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <cudaTypedefs.h>

namespace ealdm {
namespace tc {
int get_option(int option);
}
namespace ff {

constexpr int BM = 128;
constexpr int C = 256;          // model width (K of FF1, N of FF2)
constexpr int HID = 1024;       // 4 C
constexpr int NCHUNK = 16;      // chunks of 128 accumulator columns = 64 hidden columns
constexpr int SLOT = 16384;     // one [128 rows x 64 bf16] SWIZZLE_128B box
constexpr int NSLOT = 5;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;

constexpr int A_OFF = 0;                                  // x tile: 4 k-blocks
constexpr int RING_OFF = A_OFF + 4 * SLOT;                // weights
constexpr int H_OFF = RING_OFF + NSLOT * SLOT;            // gated hidden chunk, double buffered
constexpr int STG_OFF = H_OFF + 2 * SLOT;                 // per epilogue warp: 2 x [32 rows x 16 fp32] units
constexpr int STG_WARP = 2 * 2048;
constexpr int B1_OFF = STG_OFF + EPI_WARPS * STG_WARP;    // b1 (value halves pre-multiplied by 0.5), 2048 fp32
constexpr int B2_OFF = B1_OFF + 2 * HID * 4;              // b2, 256 fp32
constexpr int BAR_OFF = B2_OFF + C * 4;
constexpr int SMEM_BYTES = BAR_OFF + 512;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB dynamic shared memory limit");
static_assert(RING_OFF % 1024 == 0 && H_OFF % 1024 == 0 && STG_OFF % 1024 == 0 && BAR_OFF % 8 == 0, "alignment");

constexpr int TM_D2 = 0, TM_D1 = 256;   // TMEM columns

struct Params {
  int m_tiles;
  int out_f32;
  const float* b1;   // [2048] interleaved like W1
  const float* b2;   // [256]
};

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

// FORM: GELU form, tc_math.cuh: 1 tanh, 2 erf.  CL: thread-block cluster size.  The CTAs of a cluster work on different row
// tiles but consume the SAME weight slots in the same order, so every CTA fetches 1 / CL of each slot and multicasts it
// to the whole cluster (a slot is released by the MMAs of ALL its CTAs): per 128-row tile an SM pulls 64 KB of x +
// 1536 / CL KB of weights through its TMA path instead of 1600 KB, which is what bounds the CL = 1 kernel.
template <int FORM, int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ff_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmRes,
                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* a_full = bars;                 // [1]
  uint64_t* a_empty = bars + 1;            // [1]
  uint64_t* r_full = bars + 2;             // [NSLOT]
  uint64_t* r_empty = r_full + NSLOT;      // [NSLOT]
  uint64_t* d1_full = r_empty + NSLOT;     // [2]
  uint64_t* d1_empty = d1_full + 2;        // [2]
  uint64_t* h_full = d1_empty + 2;         // [2]
  uint64_t* h_empty = h_full + 2;          // [2]
  uint64_t* d2_full = h_empty + 2;         // [1]
  uint64_t* d2_empty = d2_full + 1;        // [1]
  uint64_t* res_bar = d2_empty + 1;        // [EPI_WARPS][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("ealdm: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW1);
    ptx::prefetch_tensormap(&tmW2);
    ptx::prefetch_tensormap(&tmRes);
    ptx::prefetch_tensormap(&tmOut);
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1);
    for (int s = 0; s < NSLOT; ++s) {
      ptx::mbar_init(&r_full[s], 1);
      ptx::mbar_init(&r_empty[s], CL);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&d1_full[b], 1);
      ptx::mbar_init(&d1_empty[b], EPI_WARPS);
      ptx::mbar_init(&h_full[b], EPI_WARPS);
      ptx::mbar_init(&h_empty[b], 1);
    }
    ptx::mbar_init(d2_full, 1);
    ptx::mbar_init(d2_empty, EPI_WARPS);
    for (int b = 0; b < 2 * EPI_WARPS; ++b) ptx::mbar_init(&res_bar[b], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // biases -> shared memory once per CTA (the value halves of b1 pre-multiplied by 0.5, see tc::geglu2)
  {
    float* b1s = reinterpret_cast<float*>(smem + B1_OFF);
    float* b2s = reinterpret_cast<float*>(smem + B2_OFF);
    for (int i = threadIdx.x; i < 2 * HID; i += NUM_THREADS) {
      float v = __ldg(p.b1 + i);
      if ((i & 16) == 0) v *= 0.5f;
      b1s[i] = v;
    }
    for (int i = threadIdx.x; i < C; i += NUM_THREADS) b2s[i] = __ldg(p.b2 + i);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) ptx::cluster_sync_all();   // every CTA's barriers exist before a peer multicasts to them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // every CTA of a cluster runs the same number of tiles (the weight ring is shared); tiles >= m_tiles are dummies whose
  // loads are zero-filled and whose stores are clipped by the TMA unit
  const int tile_first = static_cast<int>(blockIdx.x), tile_step = static_cast<int>(gridDim.x);
  const int tile_end = static_cast<int>((p.m_tiles + gridDim.x - 1) / gridDim.x) * tile_step;
  const uint32_t cta_rank = CL > 1 ? ptx::cluster_ctarank() : 0u;
  constexpr uint16_t CL_MASK = static_cast<uint16_t>((1u << CL) - 1u);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t slot = 0, sphase = 0, tcount = 0;
      auto load_w = [&](const CUtensorMap* tm, int c0, int c1) {
        ptx::mbar_wait(&r_empty[slot], sphase ^ 1u);
        ptx::mbar_arrive_expect_tx(&r_full[slot], SLOT);
        if constexpr (CL > 1)   // this CTA's 128 / CL rows of the box, to every CTA of the cluster
          ptx::tma_load_2d_mcast(smem + RING_OFF + slot * SLOT + cta_rank * (SLOT / CL), tm, &r_full[slot], c0,
                                 c1 + static_cast<int>(cta_rank) * (128 / CL), CL_MASK);
        else
          ptx::tma_load_2d(smem + RING_OFF + slot * SLOT, tm, &r_full[slot], c0, c1);
        if (++slot == NSLOT) { slot = 0; sphase ^= 1u; }
      };
      for (int tile = tile_first; tile < tile_end; tile += tile_step, ++tcount) {
        ptx::mbar_wait(a_empty, (tcount & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(a_full, 4 * SLOT);
        for (int kb = 0; kb < 4; ++kb) ptx::tma_load_2d(smem + A_OFF + kb * SLOT, &tmX, a_full, kb * 64, tile * BM);
        // the order the MMA warp consumes: W1(0), then W1(j + 1) followed by the two N halves of W2(j)
        for (int j = 0; j <= NCHUNK; ++j) {
          if (j < NCHUNK)
            for (int kb = 0; kb < 4; ++kb) load_w(&tmW1, kb * 64, j * 128);
          if (j >= 1)
            for (int half = 0; half < 2; ++half) load_w(&tmW2, (j - 1) * 64, half * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, 128);
    uint32_t slot = 0, sphase = 0, tcount = 0;
    uint32_t n_d1[2] = {0, 0}, n_h[2] = {0, 0};   // completed uses of each D1 / H buffer
    for (int tile = tile_first; tile < tile_end; tile += tile_step, ++tcount) {
      ptx::mbar_wait(a_full, tcount & 1u);
      for (int j = 0; j <= NCHUNK; ++j) {
        if (j < NCHUNK) {  // D1(j) = x_tile * W1 chunk j
          const int b = j & 1;
          ptx::mbar_wait(&d1_empty[b], (n_d1[b] & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d = tmem_base + TM_D1 + b * 128;
          for (int kb = 0; kb < 4; ++kb) {
            ptx::mbar_wait(&r_full[slot], sphase);
            ptx::tc_fence_after();
            if (lane == 0) {
              const uint64_t adesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + A_OFF + kb * SLOT));
              const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + RING_OFF + slot * SLOT));
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              if constexpr (CL > 1) ptx::umma_commit_mcast(&r_empty[slot], CL_MASK);
              else ptx::umma_commit(&r_empty[slot]);
              if (kb == 3) {
                ptx::umma_commit(&d1_full[b]);
                if (j == NCHUNK - 1) ptx::umma_commit(a_empty);   // every read of the x tile has completed
              }
            }
            __syncwarp();
            if (++slot == NSLOT) { slot = 0; sphase ^= 1u; }
          }
          ++n_d1[b];
        }
        if (j >= 1) {  // D2 += H(j - 1) * W2 chunk (j - 1)
          const int jj = j - 1, b = jj & 1;
          ptx::mbar_wait(&h_full[b], n_h[b] & 1u);
          if (jj == 0) ptx::mbar_wait(d2_empty, (tcount & 1u) ^ 1u);   // the previous tile's D2 has been drained
          ptx::tc_fence_after();
          for (int half = 0; half < 2; ++half) {
            ptx::mbar_wait(&r_full[slot], sphase);
            ptx::tc_fence_after();
            if (lane == 0) {
              const uint64_t adesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + H_OFF + b * SLOT));
              const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + RING_OFF + slot * SLOT));
              const uint32_t d = tmem_base + TM_D2 + half * 128;
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (jj | k) != 0 ? 1u : 0u);
              if constexpr (CL > 1) ptx::umma_commit_mcast(&r_empty[slot], CL_MASK);
              else ptx::umma_commit(&r_empty[slot]);
              if (half == 1) {
                ptx::umma_commit(&h_empty[b]);
                if (jj == NCHUNK - 1) ptx::umma_commit(d2_full);
              }
            }
            __syncwarp();
            if (++slot == NSLOT) { slot = 0; sphase ^= 1u; }
          }
          ++n_h[b];
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;   // TMEM lane quadrant this warp may read
    const int part = ew >> 2;    // which half of the columns of a chunk / of the output this warp owns
    const int row = quad * 32 + lane;
    const float* b1s = reinterpret_cast<const float*>(smem + B1_OFF);
    const float* b2s = reinterpret_cast<const float*>(smem + B2_OFF);
    uint8_t* stg = smem + STG_OFF + ew * STG_WARP;
    uint64_t* rbar = res_bar + 2 * ew;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    uint32_t n_d1[2] = {0, 0}, n_h[2] = {0, 0}, tcount = 0, n_res = 0;
    const uint64_t half2 = tc::pk2(0.5f, 0.5f);

    // residual unit u (16 fp32 columns) of this warp's 32 rows x 128 columns: TMA box into staging buffer u & 1
    auto issue_res = [&](int tile, int u) {   // lane 0 only
      ptx::mbar_arrive_expect_tx(&rbar[u & 1], 2048);
      ptx::tma_load_2d(stg + (u & 1) * 2048, &tmRes, &rbar[u & 1], part * 128 + u * 16, tile * BM + quad * 32);
    };

    for (int tile = tile_first; tile < tile_end; tile += tile_step, ++tcount) {
      if (lane == 0) {   // both staging buffers are free: the previous tile waited for its stores
        issue_res(tile, 0);
        issue_res(tile, 1);
      }
      for (int j = 0; j < NCHUNK; ++j) {
        const int b = j & 1;
        ptx::mbar_wait(&d1_full[b], n_d1[b] & 1u);
        ptx::tc_fence_after();
        uint32_t v[2][32];
        const uint32_t tcol = t_lane + TM_D1 + b * 128 + part * 64;
        ptx::tmem_ld_32x32(tcol, v[0]);
        ptx::tmem_ld_32x32(tcol + 32, v[1]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&d1_empty[b]);   // D1[b] is in registers: the MMA warp may overwrite it
        ++n_d1[b];
        ptx::mbar_wait(&h_empty[b], (n_h[b] & 1u) ^ 1u);  // D2(j - 2) has consumed this H buffer
        uint8_t* hb = smem + H_OFF + b * SLOT;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const float* bs = b1s + (j * 4 + part * 2 + half) * 32;
          const uint32_t(&vc)[32] = v[half];
          float o[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(bs + 4 * q);
            const float4 g = *reinterpret_cast<const float4*>(bs + 16 + 4 * q);
            const uint64_t val0 = tc::fma2(tc::pk2(__uint_as_float(vc[4 * q]), __uint_as_float(vc[4 * q + 1])), half2,
                                           tc::pk2(t.x, t.y));
            const uint64_t val1 = tc::fma2(tc::pk2(__uint_as_float(vc[4 * q + 2]), __uint_as_float(vc[4 * q + 3])), half2,
                                           tc::pk2(t.z, t.w));
            const uint64_t g0 = tc::add2(tc::pk2(__uint_as_float(vc[16 + 4 * q]), __uint_as_float(vc[17 + 4 * q])),
                                         tc::pk2(g.x, g.y));
            const uint64_t g1 = tc::add2(tc::pk2(__uint_as_float(vc[18 + 4 * q]), __uint_as_float(vc[19 + 4 * q])),
                                         tc::pk2(g.z, g.w));
            tc::upk2(tc::geglu2<FORM>(val0, g0), o[4 * q], o[4 * q + 1]);
            tc::upk2(tc::geglu2<FORM>(val1, g1), o[4 * q + 2], o[4 * q + 3]);
          }
          // hidden columns [32 part + 16 half, +16) of the chunk = 16-byte chunks 4 part + 2 half + {0, 1} of the row
          tc::sts_chunk_bf16_sw128(hb, row, 4 * part + 2 * half, &o[0]);
          tc::sts_chunk_bf16_sw128(hb, row, 4 * part + 2 * half + 1, &o[8]);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&h_full[b]);
        ++n_h[b];
      }

      // ---- final epilogue of the tile: out = (residual + b2) + D2, 8 units of 16 columns per warp ----
      ptx::mbar_wait(d2_full, tcount & 1u);
      ptx::tc_fence_after();
      for (int u = 0; u < 8; ++u) {
        uint8_t* sb = stg + (u & 1) * 2048;
        uint32_t acc[16];
        tmem_ld_32x16(t_lane + TM_D2 + part * 128 + u * 16, acc);
        ptx::mbar_wait(&rbar[u & 1], (n_res >> 1) & 1u);
        ++n_res;
        float r[16];
        {   // fp32 [32 rows x 64 B], SWIZZLE_64B: 16-byte chunk q of row l at q ^ ((l >> 1) & 3)
          const uint8_t* rp = sb + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(rp + ((q ^ sw) << 4));
            r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w;
          }
        }
        const float* bs = b2s + part * 128 + u * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(bs + 4 * q);
          r[4 * q] += t.x; r[4 * q + 1] += t.y; r[4 * q + 2] += t.z; r[4 * q + 3] += t.w;
        }
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) r[q] += __uint_as_float(acc[q]);
        if (u == 7) {   // D2 is in registers: the next tile's first W2 product may overwrite it
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(d2_empty);
        }
        if (p.out_f32) {
          uint8_t* wp = sb + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(wp + ((q ^ sw) << 4)) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        } else {        // bf16 [32 rows x 32 B], no swizzle
          uint4 lo, hi;
          lo.x = tc::pack2_bf16(r[0], r[1]); lo.y = tc::pack2_bf16(r[2], r[3]);
          lo.z = tc::pack2_bf16(r[4], r[5]); lo.w = tc::pack2_bf16(r[6], r[7]);
          hi.x = tc::pack2_bf16(r[8], r[9]); hi.y = tc::pack2_bf16(r[10], r[11]);
          hi.z = tc::pack2_bf16(r[12], r[13]); hi.w = tc::pack2_bf16(r[14], r[15]);
          __syncwarp();   // every lane has read its residual row before the buffer is overwritten at another pitch
          *reinterpret_cast<uint4*>(sb + lane * 32) = lo;
          *reinterpret_cast<uint4*>(sb + lane * 32 + 16) = hi;
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOut, sb, part * 128 + u * 16, tile * BM + quad * 32);
          ptx::bulk_commit();
          if (u + 2 < 8) {            // this buffer's next residual unit, once the store has read it
            ptx::bulk_wait_read<0>();
            issue_res(tile, u + 2);
          }
        }
      }
      if (lane == 0) ptx::bulk_wait_read<0>();   // both staging buffers are free for the next tile's prefetch
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) ptx::cluster_sync_all();   // no CTA leaves while a peer may still multicast into it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

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

static int encode2d(CUtensorMap* tm, CUtensorMapDataType dt, int es, const void* base, long long cols, long long rows,
                    long long ld, int box_cols, int box_rows, CUtensorMapSwizzle sw) {
  PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode();
  if (!encode) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tm, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(ff_fused) failed: %d", (int)r);
  return 0;
}

template <int FORM, int CL>
static int launch_cl(const CUtensorMap* tm, const Params& p, cudaStream_t st) {
  static DeviceOnce attr_set;
  static int max_clusters = 0;   // devices of one process are assumed to be the same model
  if (attr_set.pending()) {
    EALDM_CUDA(cudaFuncSetAttribute(ff_fused_kernel<FORM, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set.done();
  }
  int dev = 0, sms = 0;
  EALDM_CUDA(cudaGetDevice(&dev));
  EALDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  if (!max_clusters) {
    int n = sms / CL;
    if (CL > 1) {
      cfg.gridDim = dim3((sms / CL) * CL);
      EALDM_CUDA(cudaOccupancyMaxActiveClusters(&n, ff_fused_kernel<FORM, CL>, &cfg));
      EALDM_REQUIRE(n > 0, "ff_fused: no %d-CTA cluster fits on this device", CL);
      if (n > sms / CL) n = sms / CL;
    }
    max_clusters = n;
  }
  const int want = static_cast<int>(ceil_div(p.m_tiles, CL));
  cfg.gridDim = dim3(CL * (want < max_clusters ? want : max_clusters));
  EALDM_CUDA(cudaLaunchKernelEx(&cfg, ff_fused_kernel<FORM, CL>, tm[0], tm[1], tm[2], tm[3], tm[4], p));
  return 0;
}

int launch(const ealdm_ff_fused_args* a, cudaStream_t st) {
  auto al = [](const void* p, long long ld, int es) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * es) % 16 == 0;
  };
  EALDM_REQUIRE(a->c == C && a->hidden == HID, "ff_fused: built for c = 256, hidden = 1024 (got %lld, %lld)",
                (long long)a->c, (long long)a->hidden);
  EALDM_REQUIRE(a->x && a->w1 && a->b1 && a->w2 && a->b2 && a->residual && a->out, "ff_fused: null pointer");
  EALDM_REQUIRE(a->rows > 0 && a->ld_x >= C && a->ld_res >= C && a->ld_out >= C, "ff_fused: bad sizes");
  EALDM_REQUIRE(al(a->x, a->ld_x, 2) && al(a->w1, C, 2) && al(a->w2, HID, 2) && al(a->residual, a->ld_res, 4) &&
                    al(a->out, a->ld_out, a->out_f32 ? 4 : 2),
                "ff_fused: pointers and row pitches must be 16-byte aligned");
  CUtensorMap tm[5];
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const auto F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if (int e = encode2d(&tm[0], BF, 2, a->x, C, a->rows, a->ld_x, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  static const int cl = [] {   // EALDM_FF_CLUSTER = 1 | 2 | 4: CTAs sharing every weight slot by multicast
    const char* e = getenv("EALDM_FF_CLUSTER");
    const int v = e ? atoi(e) : 1;   // measured at M = 131072: 271 / 273 / 305 us for 1 / 2 / 4 -- the kernel is bound by
    return (v == 2 || v == 4) ? v : 1;   // the DEPTH of its 5-slot weight ring (80 KB in flight), not by L2 -> SM bandwidth
  }();
  if (int e = encode2d(&tm[1], BF, 2, a->w1, C, 2 * HID, C, 64, 128 / cl, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  if (int e = encode2d(&tm[2], BF, 2, a->w2, HID, C, HID, 64, 128 / cl, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  if (int e = encode2d(&tm[3], F32, 4, a->residual, C, a->rows, a->ld_res, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
  if (a->out_f32) {
    if (int e = encode2d(&tm[4], F32, 4, a->out, C, a->rows, a->ld_out, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return e;
  } else {
    if (int e = encode2d(&tm[4], BF, 2, a->out, C, a->rows, a->ld_out, 16, 32, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
  }
  Params p;
  p.m_tiles = static_cast<int>(ceil_div(a->rows, BM));
  p.out_f32 = a->out_f32;
  p.b1 = a->b1;
  p.b2 = a->b2;
  const bool erf = tc::get_option(EALDM_TC_OPT_GELU_ERF) != 0;
  switch (cl) {
    case 4: if (int e = erf ? launch_cl<2, 4>(tm, p, st) : launch_cl<1, 4>(tm, p, st)) return e; break;
    case 2: if (int e = erf ? launch_cl<2, 2>(tm, p, st) : launch_cl<1, 2>(tm, p, st)) return e; break;
    default: if (int e = erf ? launch_cl<2, 1>(tm, p, st) : launch_cl<1, 1>(tm, p, st)) return e; break;
  }
  EALDM_LAUNCH_CHECK();
  return 0;
}

}  // namespace ff
}  // namespace ealdm

extern "C" int ealdm_ff_geglu_fused(const ealdm_ff_fused_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a != nullptr, "ff_fused: null args");
  return ealdm::ff::launch(a, static_cast<cudaStream_t>(stream));
}
