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
  static constexpr int Q_OFF = 0;
  static constexpr int KV_OFF = Q_OFF + Q_BYTES;
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

struct Params {
  int n_q, n_kv, heads;
  int hs_q, hs_kv;      // head strides in elements
  float scale_log2;
  bf16* out;
  long long ld_out;
  float* lse;
};

template <bool WIDE>
__global__ void __launch_bounds__(NUM_THREADS, 2)
flash_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ Params p) {
  using C = Cfg<WIDE>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* kv_full = q_full + 1;    // [2]
  uint64_t* kv_empty = kv_full + 2;  // [2]
  uint64_t* s_full = kv_empty + 2;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_ready = s_free + 1;
  uint64_t* pv_done = p_ready + 1;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int q0 = blockIdx.x * BM;
  const int ntiles = p.n_kv / BN;

  if (warp == 0 && lane == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("ealdm: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmK);
    ptx::prefetch_tensormap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
      ptx::mbar_init(&pv_done[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_free, 4);
    ptx::mbar_init(p_ready, 4);
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
      ptx::mbar_arrive_expect_tx(q_full, C::Q_BYTES);
      ptx::tma_load_2d(smem + C::Q_OFF, &tmQ, q_full, h * p.hs_q, b * p.n_q + q0);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        ptx::mbar_wait(&kv_empty[s], ((t >> 1) & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&kv_full[s], C::STAGE_BYTES);
        uint8_t* st = smem + C::KV_OFF + s * C::STAGE_BYTES;
        ptx::tma_load_2d(st, &tmK, &kv_full[s], h * p.hs_kv, b * p.n_kv + t * BN);
        ptx::tma_load_2d(st + C::KV_BYTES, &tmV, &kv_full[s], h * p.hs_kv, b * p.n_kv + t * BN);
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
    auto issue_s = [&](int t) {  // S = Q K_t^T
      const uint32_t k_addr = ptx::smem_u32(smem + C::KV_OFF + (t & 1) * C::STAGE_BYTES);
      const uint64_t adesc = make_desc(q_addr, 16, SBO, LT);
      const uint64_t bdesc = make_desc(k_addr, 16, SBO, LT);
#pragma unroll
      for (int k = 0; k < 2; ++k)  // 16 elements = 32 B further inside the swizzle row
        ptx::umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
      ptx::umma_commit(s_full);
    };
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(&kv_full[0], 0);
    ptx::tc_fence_after();
    if (lane == 0) issue_s(0);
    __syncwarp();
    for (int t = 0; t < ntiles; ++t) {
      if (t + 1 < ntiles) {
        ptx::mbar_wait(&kv_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
        ptx::mbar_wait(s_free, t & 1);  // the softmax warps have read S_t
        ptx::tc_fence_after();
        if (lane == 0) issue_s(t + 1);
        __syncwarp();
      }
      ptx::mbar_wait(p_ready, t & 1);   // P_t is in shared memory (and O_{t-1} has been read)
      ptx::tc_fence_after();
      if (lane == 0) {
        const uint32_t v_addr = ptx::smem_u32(smem + C::KV_OFF + (t & 1) * C::STAGE_BYTES + C::KV_BYTES);
        const uint32_t d_tmem = tmem_base + O_COL;
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          // A: P sub-tile (k >> 2) of [128 x 64] bf16 (16 KB), 32 B per k-step inside the swizzle row
          const uint64_t adesc = make_desc(p_addr + (t & 1) * C::P_BYTES + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, 2);
          // B: V rows 16k .. 16k+15 (MN-major: 16 rows of ROW_BYTES)
          const uint64_t bdesc = make_desc(v_addr + k * 16 * C::ROW_BYTES, 16, SBO, LT);
          ptx::umma_bf16(d_tmem, adesc, bdesc, idesc_o, (t | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&kv_empty[t & 1]);
        ptx::umma_commit(&pv_done[t & 1]);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax (warps 2..5): one query row per thread =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float c = p.scale_log2;
    float m_ref = -INFINITY;   // reference maximum (raw score units) the probabilities of this row are relative to
    float l = 0.f;

    for (int t = 0; t < ntiles; ++t) {
      // P is double buffered: buffer t&1 was last read by the P V product of tile t-2, which precedes S_t in the
      // tensor pipe, so it is free as soon as S_t is
      uint8_t* pbuf = smem + C::P_OFF + (t & 1) * C::P_BYTES;
      ptx::mbar_wait(s_full, t & 1);
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
          if (t > 0) {                                            // O (all previous tiles) lives in TMEM
            ptx::mbar_wait(&pv_done[(t - 1) & 1], ((t - 1) >> 1) & 1);
            ptx::tc_fence_after();
            uint32_t vo[32];
            ptx::tmem_ld_32x32(t_lane + O_COL, vo);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) vo[j] = __float_as_uint(__uint_as_float(vo[j]) * corr);
            ptx::tmem_st_32x32(t_lane + O_COL, vo);
            ptx::tmem_st_wait();
            if (ch + 1 < 4) {   // the wait above also retired the prefetch of the next chunk: nothing to redo
            }
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
        const float ms = -m_ref * c;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        uint8_t* prow = pbuf + (ch >> 1) * 16384 + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = ex2f(fmaf(__uint_as_float(v[8 * j + i]), c, ms));
            s4[i & 3] += e[i];
          }
          uint4 u;
          u.x = pack2(e[0], e[1]);
          u.y = pack2(e[2], e[3]);
          u.z = pack2(e[4], e[5]);
          u.w = pack2(e[6], e[7]);
          const int chunk = (ch & 1) * 4 + j;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = u;
        }
        l += (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_ready);
    }
    // O = sum over the tiles, relative to m_ref
    float o[32];
    {
      const int t = ntiles - 1;
      ptx::mbar_wait(&pv_done[t & 1], (t >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld_32x32(t_lane + O_COL, v);
      ptx::tmem_ld_wait();
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

template <bool WIDE>
static int launch_t(const ealdm_attention_args* a, cudaStream_t st) {
  using C = Cfg<WIDE>;
  static DeviceOnce attr_set;
  if (attr_set.pending()) {
    EALDM_CUDA(cudaFuncSetAttribute(flash_tc_kernel<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
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
  p.hs_q = static_cast<int>(a->head_stride_q);
  p.hs_kv = static_cast<int>(a->head_stride_kv);
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.ld_out = a->ld_out;
  p.lse = a->lse;
  dim3 grid(static_cast<unsigned>(a->n_q / BM), static_cast<unsigned>(a->heads), static_cast<unsigned>(a->batch));
  EALDM_CUDA(launch_pdl(flash_tc_kernel<WIDE>, grid, dim3(NUM_THREADS), C::SMEM_BYTES, st, tq, tk, tv, p));
  EALDM_LAUNCH_CHECK();
  return 0;
}

int launch(const ealdm_attention_args* a, cudaStream_t st, bool wide) {
  return wide ? launch_t<true>(a, st) : launch_t<false>(a, st);
}

}  // namespace attn_tc
}  // namespace ealdm
