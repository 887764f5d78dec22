// Flash attention forward on the 5th-generation tensor cores (tcgen05 + TMEM) for head_dim 32, the UNet's
// self-attention at 1024 / 256 tokens (levels 0 and 1; n_q and n_kv multiples of 128).
//
// One CTA = 128 query rows of one (batch, head); two CTAs share an SM (TMEM 2 x 256 columns).
//   warp 0      TMA producer: the Q tile once, then (K, V) tiles of 128 keys through a 2-stage ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer:
//                 S  = Q K^T      M=128, N=128, K=32   (both operands K-major)            -> TMEM cols [0,128)
//                 O_t = P V       M=128, N=32,  K=128  (A = P, K-major; B = V, MN-major)  -> TMEM cols [128+32*(t&1), +32)
//   warps 2..5  softmax: tcgen05.ld hands every thread ONE query row, so the row maximum and the row sum are
//               thread-local (no shuffles).  The 128 scores of the row are read from TMEM ONCE into registers
//               (the TMEM read port moves 16 fp32 per clock per SM -- the same rate as the MUFU -- so a second
//               pass over S would double the dominant cost) and S is released at once: the next Q K^T runs under
//               this tile's exponentials.  P is rounded to bf16 and written to shared memory in the K-major
//               SWIZZLE_128B layout the P V MMA reads.  O accumulates in TMEM across all key tiles; the online-
//               softmax rescaling is lazy (FlashAttention-4): the reference maximum of a row only moves when the
//               true maximum has grown by more than 2^8, so the TMEM read-modify-write of O almost never runs.
// Per 128 x 128 score tile the SM spends 1024 cycles in the MUFU (16 ex2 / clk), 1024 cycles of TMEM reads,
// 262 cycles in the tensor pipe and ~450 issue slots.
//
// Operand staging.  WIDE = true: every TMA box is 64 columns (this head and its right neighbour) so that all
// tiles use the SWIZZLE_128B layouts of conv_tc.cu / wgrad.cu; only the first 32 columns are multiplied.
// WIDE = false: 32-column boxes with SWIZZLE_64B (half the shared memory and L2 traffic).
#include "common.cuh"
#include "ptx.cuh"
#include "tc_math.cuh"

#include <stdlib.h>

#include <cuda.h>
#include <cudaTypedefs.h>

namespace ealdm {
namespace attn_tc {

constexpr int BM = 128;   // queries per CTA
constexpr int BN = 128;   // keys per tile
constexpr int NUM_THREADS = 64 + 128;
constexpr int TMEM_COLS = 256;
constexpr int O_COL = 128;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units

template <bool WIDE>
struct Cfg {
  static constexpr int COLS = WIDE ? 64 : 32;
  static constexpr int ROW_BYTES = COLS * 2;
  static constexpr int Q_BYTES = BM * ROW_BYTES;
  static constexpr int KV_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = 2 * KV_BYTES;
  static constexpr int P_BYTES = BM * BN * 2;
  static constexpr int Q_OFF = 0;                       // Q is double buffered (persistent CTAs)
  static constexpr int KV_OFF = Q_OFF + 2 * Q_BYTES;
  static constexpr int P_OFF = KV_OFF + 2 * STAGE_BYTES;
  static constexpr int BAR_OFF = P_OFF + 2 * P_BYTES;   // P is double buffered
  static constexpr int SMEM_BYTES = BAR_OFF + 128;
  static_assert(KV_OFF % 1024 == 0 && P_OFF % 1024 == 0, "alignment");
};

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor): layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16; b_mn = 1: B operand is MN-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// 2^x for a pair of arguments on the FMA pipe (FlashAttention-4's offload of part of the exponentials: the MUFU does 16
// ex2 per clock and SM, the FMA pipe 128 FMAs): Cody-Waite split x = n + f with the 1.5 * 2^23 rounding constant, a
// degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (relative error 7.5e-5, far below the bf16 rounding of P), and n
// added to the exponent field with one integer shift-add per element.
__device__ __forceinline__ uint64_t ex2_poly2(uint64_t x2) {
  using namespace tc;
  float x0, x1;
  upk2(x2, x0, x1);
  x2 = pk2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));          // below 2^-125 the exponent arithmetic would wrap
  const uint64_t magic = pk2(12582912.f, 12582912.f);
  const uint64_t t2 = add2(x2, magic);                       // the low mantissa bits of t hold round(x)
  const uint64_t f2 = sub2(x2, sub2(t2, magic));
  uint64_t p2 = fma2(f2, pk2(0.0551716685f, 0.0551716685f), pk2(0.2426111251f, 0.2426111251f));
  p2 = fma2(p2, f2, pk2(0.6932609677f, 0.6932609677f));
  p2 = fma2(p2, f2, pk2(0.9999280572f, 0.9999280572f));
  float t0, t1, p0, p1;
  upk2(t2, t0, t1);
  upk2(p2, p0, p1);
  const float r0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  const float r1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
  return pk2(r0, r1);
}

struct Params {
  int n_q, n_kv, heads, batch;
  int hs_q, hs_kv;      // head strides in elements
  float scale_log2;
  bf16* out;
  long long ld_out;
  float* lse;
};

// POLY: how many of every 8 exponentials run on the FMA pipe (0, 2 or 4)
//
// PERSISTENT: a CTA walks work items (128 queries of one (batch, head)) i = blockIdx.x, blockIdx.x + gridDim.x, ...; the
// shared-memory rings, the S / P / O hand-offs and their mbarrier parities run on ONE global key-tile counter g across the
// items, Q is double buffered, so the loads and the first Q K^T of item i + 1 run under the last exponentials, the O read
// and the output stores of item i -- the per-CTA set-up (TMEM allocation, barrier init, first DRAM round trips) is paid
// once per SM slot instead of once per 128 queries.
template <bool WIDE, int POLY>
__global__ void __launch_bounds__(NUM_THREADS, 2)
flash_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ Params p) {
  using C = Cfg<WIDE>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);   // [2]
  uint64_t* q_empty = q_full + 2;    // [2]
  uint64_t* kv_full = q_empty + 2;   // [2]
  uint64_t* kv_empty = kv_full + 2;  // [2]
  uint64_t* s_full = kv_empty + 2;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_ready = s_free + 1;
  uint64_t* pv_done = p_ready + 1;   // [2]
  uint64_t* o_free = pv_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = p.n_kv / BN;
  const int qtiles = p.n_q / BM;
  const int items = qtiles * p.heads * p.batch;
  // item -> (query tile, head, batch): the items of one (batch, head) are neighbours and share K / V through L2
  auto item_coords = [&](int it, int& q0, int& h, int& b) {
    const int qt = it % qtiles;
    const int bh = it / qtiles;
    q0 = qt * BM;
    h = bh % p.heads;
    b = bh / p.heads;
  };

  if (warp == 0 && lane == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("ealdm: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmK);
    ptx::prefetch_tensormap(&tmV);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&q_full[s], 1);
      ptx::mbar_init(&q_empty[s], 1);
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
      ptx::mbar_init(&pv_done[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_free, 4);
    ptx::mbar_init(p_ready, 4);
    ptx::mbar_init(o_free, 4);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // programmatic dependent launch: the set-up above ran under the predecessor's tail
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t g = 0;   // global key-tile counter
      int li = 0;       // local item counter
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++li) {
        int q0, h, b;
        item_coords(it, q0, h, b);
        const int qs = li & 1;
        ptx::mbar_wait(&q_empty[qs], ((li >> 1) & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&q_full[qs], C::Q_BYTES);
        ptx::tma_load_2d(smem + C::Q_OFF + qs * C::Q_BYTES, &tmQ, &q_full[qs], h * p.hs_q, b * p.n_q + q0);
        for (int t = 0; t < ntiles; ++t, ++g) {
          const int s = g & 1;
          ptx::mbar_wait(&kv_empty[s], ((g >> 1) & 1) ^ 1u);
          ptx::mbar_arrive_expect_tx(&kv_full[s], C::STAGE_BYTES);
          uint8_t* st = smem + C::KV_OFF + s * C::STAGE_BYTES;
          ptx::tma_load_2d(st, &tmK, &kv_full[s], h * p.hs_kv, b * p.n_kv + t * BN);
          ptx::tma_load_2d(st + C::KV_BYTES, &tmV, &kv_full[s], h * p.hs_kv, b * p.n_kv + t * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc(BM, BN, 0);
    constexpr uint32_t idesc_o = make_idesc(BM, 32, 1);
    constexpr uint32_t LT = WIDE ? 2u : 4u;
    constexpr uint32_t SBO = WIDE ? 1024u : 512u;
    const uint32_t q_addr = ptx::smem_u32(smem + C::Q_OFF);
    const uint32_t p_addr = ptx::smem_u32(smem + C::P_OFF);
    const int my_items = blockIdx.x < static_cast<unsigned>(items)
                             ? (items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
    const uint32_t total = static_cast<uint32_t>(my_items) * static_cast<uint32_t>(ntiles);
    // S(g) = Q(item of g) K(g)^T; `first` / `last`: g is the first / last key tile of its item
    auto issue_s = [&](uint32_t g, int li, bool first, bool last) {
      if (first) ptx::mbar_wait(&q_full[li & 1], (li >> 1) & 1);
      ptx::mbar_wait(&kv_full[g & 1], (g >> 1) & 1);
      if (g > 0) ptx::mbar_wait(s_free, (g - 1) & 1);   // the softmax warps have read S(g - 1)
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t k_addr = ptx::smem_u32(smem + C::KV_OFF + (g & 1) * C::STAGE_BYTES);
        const uint64_t adesc = make_desc(q_addr + (li & 1) * C::Q_BYTES, 16, SBO, LT);
        const uint64_t bdesc = make_desc(k_addr, 16, SBO, LT);
#pragma unroll
        for (int k = 0; k < 2; ++k)  // 16 elements = 32 B further inside the swizzle row
          ptx::umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
        if (last) ptx::umma_commit(&q_empty[li & 1]);   // this item's Q tile has been multiplied for the last time
      }
      __syncwarp();
    };
    if (total > 0) issue_s(0, 0, true, ntiles == 1);
    int li = 0, t = 0;   // item / key tile of g
    for (uint32_t g = 0; g < total; ++g) {
      if (g + 1 < total) {
        const bool nfirst = t + 1 == ntiles;
        const int nli = nfirst ? li + 1 : li;
        const int nt = nfirst ? 0 : t + 1;
        issue_s(g + 1, nli, nfirst, nt + 1 == ntiles);
      }
      ptx::mbar_wait(p_ready, g & 1);   // P(g) is in shared memory
      if (t == 0 && li > 0) ptx::mbar_wait(o_free, (li - 1) & 1);   // the previous item's O has been read out of TMEM
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t v_addr = ptx::smem_u32(smem + C::KV_OFF + (g & 1) * C::STAGE_BYTES + C::KV_BYTES);
        const uint32_t d_tmem = tmem_base + O_COL;
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          // A: P sub-tile (k >> 2) of [128 x 64] bf16 (16 KB), 32 B per k-step inside the swizzle row
          const uint64_t adesc = make_desc(p_addr + (g & 1) * C::P_BYTES + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, 2);
          // B: V rows 16k .. 16k+15 (MN-major: 16 rows of ROW_BYTES)
          const uint64_t bdesc = make_desc(v_addr + k * 16 * C::ROW_BYTES, 16, SBO, LT);
          ptx::umma_bf16(d_tmem, adesc, bdesc, idesc_o, (t | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&kv_empty[g & 1]);
        ptx::umma_commit(&pv_done[g & 1]);
      }
      __syncwarp();
      if (++t == ntiles) { t = 0; ++li; }
    }
  } else {
    // ===================== softmax (warps 2..5): one query row per thread =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float c = p.scale_log2;
    uint32_t g = 0;
    int li = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++li) {
    int q0, h, b;
    item_coords(it, q0, h, b);
    float m_ref = -INFINITY;   // reference maximum (raw score units) the probabilities of this row are relative to
    float l = 0.f;

    for (int t = 0; t < ntiles; ++t, ++g) {
      // P is double buffered: buffer g&1 was last read by the P V product of tile g-2, which precedes S(g) in the
      // tensor pipe, so it is free as soon as S(g) is
      uint8_t* pbuf = smem + C::P_OFF + (g & 1) * C::P_BYTES;
      ptx::mbar_wait(s_full, g & 1);
      ptx::tc_fence_after();
      uint32_t sv[2][32];
      ptx::tmem_ld_32x32(t_lane, sv[0]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        // the TMEM load of chunk ch+1 flies under the exponentials of chunk ch (S is read exactly once)
        ptx::tmem_ld_wait();
        if (ch + 1 < 4) {
          ptx::tmem_ld_32x32(t_lane + (ch + 1) * 32, sv[(ch + 1) & 1]);
        } else {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(s_free);   // S is consumed: the next Q K^T overlaps the rest of this tile
        }
        const uint32_t(&v)[32] = sv[ch & 1];
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(v[j]));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // lazy rescaling at chunk granularity: the reference only moves when a score outgrows it by more than 2^8
        const bool grow = (mx - m_ref) * c > RESCALE_THRESHOLD;   // always true for the first chunk (m_ref = -inf)
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx : m_ref;
          const float corr = ex2f((m_ref - m_new) * c);           // 1 for the rows that keep their reference
          m_ref = m_new;
          l *= corr;
          if (t > 0) {                                            // O (all previous tiles of this item) lives in TMEM
            ptx::mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
            ptx::tc_fence_after();
            uint32_t vo[32];
            ptx::tmem_ld_32x32(t_lane + O_COL, vo);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) vo[j] = __float_as_uint(__uint_as_float(vo[j]) * corr);
            ptx::tmem_st_32x32(t_lane + O_COL, vo);
            ptx::tmem_st_wait();
          }
          for (int pc = 0; pc < ch; ++pc) {                       // probabilities of this tile already written
            uint8_t* prow = pbuf + (pc >> 1) * 16384 + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4* q = reinterpret_cast<uint4*>(prow + ((((pc & 1) * 4 + j) ^ (row & 7)) << 4));
              uint4 u = *q;
              uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
                w[i] = pack2(__uint_as_float(w[i] << 16) * corr, __uint_as_float(w[i] & 0xffff0000u) * corr);
              *q = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
        // p = 2^(s c - m c) on packed f32x2 arithmetic (one FFMA2 / FADD2 per element pair); POLY of every 8
        // exponentials come from the FMA pipe (ex2_poly2), the others from the MUFU
        const float ms = -m_ref * c;
        const uint64_t c2 = tc::pk2(c, c), ms2 = tc::pk2(ms, ms);
        uint64_t sum2[2] = {0ull, 0ull};
        uint8_t* prow = pbuf + (ch >> 1) * 16384 + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint64_t x2 =
                tc::fma2(tc::pk2(__uint_as_float(v[8 * j + 2 * i]), __uint_as_float(v[8 * j + 2 * i + 1])), c2, ms2);
            uint64_t e2;
            if (2 * i < POLY) {
              e2 = ex2_poly2(x2);
            } else {
              float x0, x1;
              tc::upk2(x2, x0, x1);
              e2 = tc::pk2(ex2f(x0), ex2f(x1));
            }
            sum2[i & 1] = tc::add2(sum2[i & 1], e2);
            float e0, e1;
            tc::upk2(e2, e0, e1);
            w[i] = pack2(e0, e1);
          }
          const int chunk = (ch & 1) * 4 + j;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        {
          float a0, a1, b0, b1;
          tc::upk2(sum2[0], a0, a1);
          tc::upk2(sum2[1], b0, b1);
          l += (a0 + a1) + (b0 + b1);
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_ready);
    }
    // O = sum over the tiles, relative to m_ref
    float o[32];
    {
      const uint32_t gl = g - 1;
      ptx::mbar_wait(&pv_done[gl & 1], (gl >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld_32x32(t_lane + O_COL, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_free);   // the next item's first P V may overwrite O
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
    }
    const float m = m_ref;
    const float inv = 1.0f / l;
    const long long grow = static_cast<long long>(b) * p.n_q + q0 + row;
    bf16* orow = p.out + grow * p.ld_out + h * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u;
      u.x = pack2(o[8 * j] * inv, o[8 * j + 1] * inv);
      u.y = pack2(o[8 * j + 2] * inv, o[8 * j + 3] * inv);
      u.z = pack2(o[8 * j + 4] * inv, o[8 * j + 5] * inv);
      u.w = pack2(o[8 * j + 6] * inv, o[8 * j + 7] * inv);
      *reinterpret_cast<uint4*>(orow + 8 * j) = u;
    }
    if (p.lse != nullptr)
      p.lse[(static_cast<long long>(b) * p.heads + h) * p.n_q + q0 + row] = fmaf(m, c, log2f(l));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
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

// 2-D map over [rows, cols] bf16 with pitch ld: box = (box_cols, 128 rows)
static int encode2d(CUtensorMap* tm, const void* base, long long cols, long long rows, long long ld, int box_cols,
                    bool wide) {
  PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode();
  if (!encode) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(attention) failed: %d", (int)r);
  return 0;
}

bool supported(const ealdm_attention_args* a) {
  if (a->dtype != EALDM_BF16 || a->head_dim != 32 || a->scale <= 0.f) return false;
  if (a->n_q % BM != 0 || a->n_kv % BN != 0) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(a->q) || !al16(a->k) || !al16(a->v) || !al16(a->out)) return false;
  if (a->ld_q % 8 || a->ld_kv % 8 || a->ld_out % 8 || a->head_stride_q % 8 || a->head_stride_kv % 8) return false;
  return true;
}

template <bool WIDE, int POLY>
static int launch_t(const ealdm_attention_args* a, cudaStream_t st) {
  using C = Cfg<WIDE>;
  static DeviceOnce attr_set;
  if (attr_set.pending()) {
    EALDM_CUDA(cudaFuncSetAttribute(flash_tc_kernel<WIDE, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES));
    attr_set.done();
  }
  CUtensorMap tq, tk, tv;
  const long long qcols = (a->heads - 1) * a->head_stride_q + 32;
  const long long kcols = (a->heads - 1) * a->head_stride_kv + 32;
  if (int e = encode2d(&tq, a->q, qcols, a->batch * a->n_q, a->ld_q, C::COLS, WIDE)) return e;
  if (int e = encode2d(&tk, a->k, kcols, a->batch * a->n_kv, a->ld_kv, C::COLS, WIDE)) return e;
  if (int e = encode2d(&tv, a->v, kcols, a->batch * a->n_kv, a->ld_kv, C::COLS, WIDE)) return e;
  Params p;
  p.n_q = static_cast<int>(a->n_q);
  p.n_kv = static_cast<int>(a->n_kv);
  p.heads = static_cast<int>(a->heads);
  p.batch = static_cast<int>(a->batch);
  p.hs_q = static_cast<int>(a->head_stride_q);
  p.hs_kv = static_cast<int>(a->head_stride_kv);
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.ld_out = a->ld_out;
  p.lse = a->lse;
  // persistent: two CTAs per SM (EALDM_ATTN_PERSIST=0: one CTA per work item, for A/B measurements)
  const long long items = (a->n_q / BM) * a->heads * a->batch;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  static const char* e_persist = getenv("EALDM_ATTN_PERSIST");
  const long long slots = (e_persist && atoi(e_persist) == 0) ? items : 2LL * sms;
  dim3 grid(static_cast<unsigned>(items < slots ? items : slots));
  EALDM_CUDA(launch_pdl(flash_tc_kernel<WIDE, POLY>, grid, dim3(NUM_THREADS), C::SMEM_BYTES, st, tq, tk, tv, p));
  EALDM_LAUNCH_CHECK();
  return 0;
}

// EALDM_ATTN_POLY = 0 / 2 / 4: exponentials per 8 on the FMA pipe (A/B switch; results differ by the polynomial's 7.5e-5)
static int poly_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EALDM_ATTN_POLY");
    v = e ? atoi(e) : 2;
  }
  return v;
}

int launch(const ealdm_attention_args* a, cudaStream_t st, bool wide) {
  if (wide) return launch_t<true, 0>(a, st);
  switch (poly_mode()) {
    case 0: return launch_t<false, 0>(a, st);
    case 4: return launch_t<false, 4>(a, st);
    default: return launch_t<false, 2>(a, st);
  }
}

}  // namespace attn_tc
}  // namespace ealdm
