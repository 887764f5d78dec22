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
//   warps 2..9  epilogue, two per TMEM lane quadrant (even / odd 32-column units, or -- "wide" passes, when there is
//               no residual and no shadow output -- the two 128-column halves with 128-byte rows per TMA store)
// so the epilogue of tile i overlaps the main loop of tile i+1.
//
// CTA pairs (template parameter CTA2, chosen for K >= 1024 and an even number of M tiles): the grid is a persistent
// set of 2-CTA clusters and a work item is a 256 x BN super-tile.  Each CTA stages its own 128 A rows and HALF of the B
// tile; the leader (cluster rank 0) issues `tcgen05.mma.cta_group::2` for both, its full barriers collect the TMA
// transactions of both CTAs (the peer's loads name the leader's mbarrier), `tcgen05.commit ... multicast::cluster`
// releases the stage and publishes the accumulator in both CTAs, and the peer's epilogue warps arrive remotely on the
// leader's tmem_empty barrier.  Each CTA's epilogue is unchanged: it drains its own 128 TMEM lanes.  Operand traffic
// per SM and k-block drops from 48 KB to 32 KB, which is what bounds the single-CTA kernel.
//
// Epilogue data movement is TMA in both directions: every epilogue warp owns [32 rows x 32 columns]
// units of the tile.  The residual unit is prefetched by a TMA box load (issued one unit ahead, across
// tile boundaries) into a swizzled shared-memory buffer; the thread that owns accumulator row r
// (tcgen05.ld hands every thread one row) reads row r of it, adds bias / per-image row vector /
// activation, writes the result back into the same buffer and one lane issues the TMA store (plus a
// second store of the bf16 operand shadow).  No thread computes a global address, every global access
// is a full 128-byte (fp32) or 64-byte (bf16) row segment, and M / N edges are clipped by the TMA unit.
#include "common.cuh"
#include "ptx.cuh"
#include "tc_math.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

namespace ealdm {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 per K block = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
// warp 0 TMA, warp 1 MMA, warps 2..9 epilogue, warp 10 row statistics of a folded LayerNorm (idle otherwise; the
// register file is allocated in units of 4 warps, so the 11th warp does not lower the 168-register ceiling)
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS + 32;
constexpr int EBUF_BYTES = 4096;                  // one [32 x 32] fp32 unit (bf16 units use half)
constexpr int O2BUF_BYTES = 2048;                 // bf16 shadow of a unit
constexpr int EPI_WARP_BYTES = 2 * EBUF_BYTES + O2BUF_BYTES;

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
  // epilogue
  const float* bias;
  const float* rowvec;
  long long ld_rowvec;
  int act;
  int out_f32;
  int has_res, res_f32;
  int has_out2;
  // GroupNorm partial statistics of the output (optional)
  float2* gn_partial;
  int gn_ld;       // entries (octets, or quads with gn_quads) per (image, chunk) row
  int gn_quads;    // partial entries are per 4 channels instead of 8 (GroupNorm groups of 4 channels)
  int gn_chunks;   // 32-pixel chunks per image
  // adjoint mode: B is read MN-major from the forward-packed matrix [K rows = src channels, taps * N columns]
  int b_mn;
  // relaxed epilogue store wait: one TMA store group may stay in flight when the next unit reuses no buffer
  int relaxed_wait;
  // wide epilogue passes (BN = 256, no residual, no shadow): every warp fills [32 rows x 128 B] boxes for 64 (fp32) or
  // 128 (bf16) columns and pays ONE wait / fence / store-issue sequence for them instead of one per 32 columns
  int wide;
  // nearest-2x upsampling folded into the convolution (phases = 4, else 1): output pixel (2y + py, 2x + px) of a 3x3
  // convolution over the upsampled image only sees input rows {y + py - 1, y + py} and columns {x + px - 1, x + px},
  // so each of the four output phases is a 2x2 convolution over the LOW-resolution input with pre-summed weights
  // (4/9 of the FLOPs, and the 4x larger upsampled tensor is never written or read).  A work item is (phase, M tile
  // of the low-resolution grid, N tile); the weights of phase p are the K columns [p * 4C, (p + 1) * 4C); the output
  // (and its shadow) is addressed through 5-D tensor maps (C, px, W/2, py, N * H/2).
  int phases;
  int gn_phase_chunks;  // 32-pixel chunks of the low-resolution grid per image: phase p fills chunks [p * this, ...)
  // LayerNorm folded into the GEMMs around it (linear geometry only: h_out = 1, one "image", row = output pixel).
  // PRODUCER side (ln_out): the epilogue that writes the fp32 token stream also writes, per row and per 32-column
  // unit, {sum, sum of squares} of its columns: ln_out[row * ln_out_parts + col / 32] (independent of the N tile).
  // CONSUMER side (ln_in): A is the RAW stream (its bf16 shadow), W was packed as W * gamma, `bias` holds
  // W beta (+ bias), ln_c1[col] = sum_k (W * gamma)[col, k]; with mu, rstd of the row from the partials
  //   out = rstd * acc - rstd * mu * c1[col] + bias[col]  ==  LayerNorm(x) W^T + bias
  // per-image B operand (ealdm_conv_args::wi_*): tmB is a 4-D (c, token, head, image) map, a tile lies in one image
  int b_img;
  // ... or in TWO (64-token images, 128 (head, key) logits each): the logits GEMM runs 128-column tiles, N tile nt of an
  // M tile multiplies by the operand of the tile's image nt and its epilogue zeroes the rows of the OTHER image, so every
  // row carries [its probabilities | zeros] (or the reverse) over 256 columns; the output GEMM then reads those 256
  // columns as K against [Zt of image 0; Zt of image 1] -- a block-diagonal product
  int b_img2;
  // LayerNorm APPLIED by the producing epilogue (ealdm_conv_args::ln_gamma; N == BN == 256, so a CTA holds whole rows):
  // pass 1 writes the fp32 result as usual, keeps per-row {sum, sum of squares} and stores the result back over its
  // TMEM accumulator; the two warps that share a row exchange their sums; pass 2 re-reads TMEM and writes
  // LayerNorm(result) * gamma + beta as bf16 units through the shadow-output map (tmOut2) -- staged in shared memory and
  // stored by TMA: direct 16-byte stores from registers (32 rows per instruction) measured 3x slower
  const float* ln_gamma;
  const float* ln_beta;
  float ln_apply_eps;
  float2* ln_out;
  int ln_out_parts;
  const float2* ln_in;
  int ln_in_parts;
  float ln_inv_c;   // 1 / (channels the LayerNorm runs over)
  float ln_eps;
  const float* ln_c1;
  // GroupNorm (+ SiLU) APPLIED by the producing epilogue (ealdm_conv_args::gn_gamma; 256-column tiles).  Pass 1 is the
  // usual epilogue: it writes the {sum, sum of squares} partials of its [32 pixels x 8 channels] blocks (gn_partial) and
  // -- unless gna_only -- the fp32 result, and stores the result back over its TMEM accumulator.  Every epilogue warp
  // then counts itself in at the (image, N tile) its 32 rows belong to and waits until the image's other tiles -- in
  // flight on neighbouring CTAs, or the next items of CTAs that do not wait for this one -- have counted in too
  // (gna_expected warps); pass 2 folds the image's partials (fp64, chunk order), and writes
  // act((x - mean) rstd gamma + beta) as bf16 through tmOut2 (gna_only: through tmOut -- the un-normalised tensor is
  // never written).  The counters return to zero with the last reader.  The stand-alone GroupNorm pass over the tensor
  // (and, with gna_only, the tensor itself) disappears.
  const float* gna_gamma;
  const float* gna_beta;
  float gna_eps;
  int gna_silu;
  int gna_only;
  int gna_debug;              // timing experiments (EALDM_GNA_DEBUG): 1 no wait, 2 no fold, 4 no pass 2, 8 no count-in
  int gna_octets;             // channel octets per group: 1, 2 or 4
  double gna_inv_count;       // 1 / (pixels per image * channels per group)
  unsigned int gna_expected;  // epilogue warps per (image, N tile): 2 * pixels per image / 32
  unsigned int* gna_counters; // [image][N tile]{counted in, read}
  // stream-K over the tail of the tile list (CTA pairs, 256-column tiles): the k-blocks of the LAST sk_tiles work items
  // are dealt evenly to the clusters, so that a launch of 1.73 waves costs 1.73 and not 2.  A cluster's share is
  // [end of a tile | whole tiles | beginning of a tile]; it computes the BEGINNING first and parks that accumulator in
  // `sk_ws` (fp32, exact), and the END last: its epilogue warps load the neighbour's parked accumulator into TMEM and
  // the MMAs continue on it, so every output is the same k-ordered fp32 sum as without the split (bit-identical).
  int sk_tiles;
  float4* sk_ws;          // [cluster][CTA of the pair][32-column unit][8][128 rows] float4
  unsigned int* sk_flags; // [cluster][CTA][epilogue warp]: 1 = parked (reset by the reader)
};

// CTA2: the tile is 256 x BN over a CTA pair (cta_group::2); each CTA stages its 128 A rows and BN/2 B rows
template <int BN, bool CTA2 = false>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (CTA2 ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = CTA2 ? (BN >= 256 ? 4 : 6) : ((BN >= 256) ? 3 : (BN >= 128 ? 4 : 6));
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES;                 // 1024-aligned
  static constexpr int BIAS_OFF = EPI_OFF + EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int BIAS_BYTES = 2 * BN * 4;                        // per accumulator stage
  static constexpr int BAR_OFF = BIAS_OFF + BIAS_BYTES;
  static constexpr int BAR_BYTES = 512;  // (2*STAGES + 8 + 2*EPI_WARPS) mbarriers + the TMEM base slot
  static constexpr int SMEM_BYTES = BAR_OFF + BAR_BYTES;
  static_assert((2 * STAGES + 9 + 2 * EPI_WARPS) * 8 + 4 <= BAR_BYTES, "barrier block too small");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB dynamic shared memory limit");
  static_assert(EPI_OFF % 1024 == 0 && EPI_WARP_BYTES % 1024 == 0 && BAR_OFF % 8 == 0, "alignment");
};

// GroupNorm partial statistics of one [32 pixels x 32 channels] unit held one row per lane (see Params::gn_partial)
__device__ __forceinline__ void gn_partial_unit(const Params& p, const float (&r)[32], int lane, bool valid, int n,
                                                int h, int w, int col0, int chunk0 = 0) {
  // {sum, sum of squares} of the 4 channel octets of this unit over the warp's 32 pixel rows (one 32-pixel
  // chunk of one image): 8 values per thread, folded across the lanes in a fixed order with 9 shuffles
  float v8[8];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { a += r[8 * o + j]; b = fmaf(r[8 * o + j], r[8 * o + j], b); }
    v8[2 * o] = valid ? a : 0.f;
    v8[2 * o + 1] = valid ? b : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {   // lanes 16..31 keep values 4..7
    const float send = (lane & 16) ? v8[i] : v8[i + 4];
    const float keep = (lane & 16) ? v8[i + 4] : v8[i];
    v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = (lane & 8) ? v8[i] : v8[i + 2];
    const float keep = (lane & 8) ? v8[i + 2] : v8[i];
    v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float send = (lane & 4) ? v8[0] : v8[1];
    const float keep = (lane & 4) ? v8[1] : v8[0];
    v8[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 2);
  v8[0] += __shfl_xor_sync(0xffffffffu, v8[0], 1);
  if ((lane & 3) == 0 && n < p.Nimg && col0 < p.N) {
    const int vi = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);  // value index 0..7
    const int chunk = chunk0 + ((h * p.Wout + w) >> 5);
    float* dst = reinterpret_cast<float*>(p.gn_partial + (static_cast<long long>(n) * p.gn_chunks + chunk) * p.gn_ld +
                                          (col0 >> 3) + (vi >> 1));
    dst[vi & 1] = v8[0];
  }
}

// the same per channel QUAD (GroupNorm groups of 4 channels): 16 values per thread, 15 exchange shuffles + 1
__device__ __forceinline__ void gn_partial_unit_quads(const Params& p, const float (&r)[32], int lane, bool valid, int n,
                                                      int h, int w, int col0, int chunk0 = 0) {
  float v16[16];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { a += r[4 * q + j]; b = fmaf(r[4 * q + j], r[4 * q + j], b); }
    v16[2 * q] = valid ? a : 0.f;
    v16[2 * q + 1] = valid ? b : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {   // lanes 16..31 keep values 8..15
    const float send = (lane & 16) ? v16[i] : v16[i + 8];
    const float keep = (lane & 16) ? v16[i + 8] : v16[i];
    v16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = (lane & 8) ? v16[i] : v16[i + 4];
    const float keep = (lane & 8) ? v16[i + 4] : v16[i];
    v16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = (lane & 4) ? v16[i] : v16[i + 2];
    const float keep = (lane & 4) ? v16[i + 2] : v16[i];
    v16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const float send = (lane & 2) ? v16[0] : v16[1];
    const float keep = (lane & 2) ? v16[1] : v16[0];
    v16[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v16[0] += __shfl_xor_sync(0xffffffffu, v16[0], 1);
  if ((lane & 1) == 0 && n < p.Nimg && col0 < p.N) {
    const int vi = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);  // 0..15
    const int chunk = chunk0 + ((h * p.Wout + w) >> 5);
    float* dst = reinterpret_cast<float*>(p.gn_partial + (static_cast<long long>(n) * p.gn_chunks + chunk) * p.gn_ld +
                                          (col0 >> 2) + (vi >> 1));
    dst[vi & 1] = v16[0];
  }
}

// output maps of upsampling phases 1..3 (phase 0 uses tmOut / tmOut2): the pixels (2y + py, 2x + px) of one phase form a
// regular sub-grid of the NHWC output, i.e. a 4-D (C, W/2, H/2, N) tensor with doubled pixel / row strides whose base
// is shifted by (py, px)
struct PhaseMaps {
  CUtensorMap out[3];
  CUtensorMap out2[3];
};

template <int BN, int GEGLU, bool CTA2>   // GEGLU: 0 none, 1 tanh-form GELU, 2 erf-form GELU (tc_math.cuh)
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmOut2, const __grid_constant__ CUtensorMap tmRes,
               const __grid_constant__ Params p, const __grid_constant__ PhaseMaps pm) {
  using C = Cfg<BN, CTA2>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;  // [EPI_WARPS][2]
  uint64_t* ln_full = res_bar + 2 * EPI_WARPS;   // [2] row statistics of a folded LayerNorm, per accumulator stage
  uint64_t* ln_empty = ln_full + 2;              // [2]
  uint64_t* seed_bar = ln_empty + 2;             // stream-K: the accumulator of the resumed tile is loaded
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(seed_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair: rank 0 (leader) owns the full barriers and issues the MMAs of both CTAs
  const uint32_t cta_rank = CTA2 ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) {
      printf("ealdm: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    ptx::prefetch_tensormap(&tmA0);
    if (p.nseg > 1) ptx::prefetch_tensormap(&tmA1);
    ptx::prefetch_tensormap(&tmB);
    ptx::prefetch_tensormap(&tmOut);
    if (p.has_out2) ptx::prefetch_tensormap(&tmOut2);
    if (p.has_res) ptx::prefetch_tensormap(&tmRes);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], CTA2 ? 2 * EPI_WARPS : EPI_WARPS);
    }
    for (int a = 0; a < 2 * EPI_WARPS; ++a) ptx::mbar_init(&res_bar[a], 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&ln_full[a], 1);
      ptx::mbar_init(&ln_empty[a], EPI_WARPS);
    }
    ptx::mbar_init(seed_bar, CTA2 ? 2 * EPI_WARPS : EPI_WARPS);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) {
      ptx::tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) ptx::cluster_sync_all();  // the peer's barriers exist before anything remote touches them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above ran under the predecessor's tail; global memory comes next
  pdl_wait();
  pdl_trigger();

  // work items: tiles (one CTA each) or, for CTA pairs, super-tiles of two consecutive M tiles and one N tile
  const int tile_first = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int total_tiles = (CTA2 ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles * p.phases;
  const int kblocks_total = p.seg[0].kblocks + (p.nseg > 1 ? p.seg[1].kblocks : 0);
  // tile -> (M tile, N tile, upsampling phase); the phases of one M tile are neighbours (they read the same input)
  auto tile_mnp = [&](int tile, int& mt, int& nt, int& phase) {
    int q = tile / p.n_tiles;
    nt = tile - q * p.n_tiles;
    phase = 0;
    if (p.phases > 1) {
      phase = q & 3;
      q >>= 2;
    }
    mt = CTA2 ? 2 * q + static_cast<int>(cta_rank) : q;  // mt == m_tiles (odd tail): every box is out of range
  };
  auto tile_mn = [&](int tile, int& mt, int& nt) {
    int phase;
    tile_mnp(tile, mt, nt, phase);
  };
  // work items of this CTA (pair), the same list in every role.  kind 0: a whole tile; stream-K (Params::sk_tiles):
  // kind 1 = the first k-blocks of a tile (item 0: the accumulator is parked for the next cluster), kind 2 = the
  // remaining k-blocks of the tile the previous cluster began (the last item: resumed from its parked accumulator)
  int n_items, n_dp, n_full = 0, full0 = 0, head_tile = -1, head_kb1 = 0, tail_tile = -1, tail_kb0 = 0;
  if (p.sk_tiles == 0) {
    n_dp = tile_first < total_tiles ? (total_tiles - tile_first + tile_step - 1) / tile_step : 0;
    n_items = n_dp;
  } else {
    const int dp_tiles = total_tiles - p.sk_tiles;   // a multiple of the number of clusters
    n_dp = dp_tiles / tile_step;
    const long long units = static_cast<long long>(p.sk_tiles) * kblocks_total;
    const long long u0 = units * tile_first / tile_step, u1 = units * (tile_first + 1) / tile_step;
    const int ta = static_cast<int>(u0 / kblocks_total), ka = static_cast<int>(u0 % kblocks_total);
    const int tb = static_cast<int>(u1 / kblocks_total), kb_end = static_cast<int>(u1 % kblocks_total);
    full0 = dp_tiles + ta;
    if (ka > 0) { tail_tile = dp_tiles + ta; tail_kb0 = ka; ++full0; }
    n_full = dp_tiles + tb - full0;
    if (kb_end > 0) { head_tile = dp_tiles + tb; head_kb1 = kb_end; }
    n_items = (head_tile >= 0 ? 1 : 0) + n_dp + n_full + (tail_tile >= 0 ? 1 : 0);
  }
  auto get_item = [&](int i, int& tile, int& kb0, int& kb1, int& kind) {
    kb0 = 0;
    kb1 = kblocks_total;
    kind = 0;
    if (head_tile >= 0) {
      if (i == 0) { tile = head_tile; kb1 = head_kb1; kind = 1; return; }
      --i;
    }
    if (i < n_dp) { tile = tile_first + i * tile_step; return; }
    i -= n_dp;
    if (i < n_full) { tile = full0 + i; return; }
    tile = tail_tile;
    kb0 = tail_kb0;
    kind = 2;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0b = CTA2 ? ptx::mapa_u32(&full_bar[0], 0) : 0u;  // the leader's full barriers
      for (int item = 0; item < n_items; ++item) {
        int tile, kb0, kb1, kind;
        get_item(item, tile, kb0, kb1, kind);
        int mt, nt, uph;   // (`phase` is the mbarrier parity of the ring in this role)
        tile_mnp(tile, mt, nt, uph);
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);
        // an upsampling phase shifts the 2x2 taps by (py, px) and selects its own K columns of W
        const int w0 = tw * p.bw + (uph & 1), h0 = th * p.bh + (uph >> 1), n0 = tn * p.bn;
        const int pkoff = uph * p.seg[0].kblocks * BK;
        // position of k-block kb0 in the walk over (segment, tap, channel block)
        int s = 0, kb = kb0;   // kb: k-block inside the segment
        if (kb >= p.seg[0].kblocks) { kb -= p.seg[0].kblocks; s = 1; }
        Segment sg = p.seg[s];
        const CUtensorMap* tmA = (s == 0) ? &tmA0 : &tmA1;
        int tap = kb / sg.cblk, cb = kb - tap * sg.cblk;
        for (int kbt = kb0; kbt < kb1; ++kbt, ++kb) {
          if (kb == sg.kblocks) {   // (only ever from segment 0 into segment 1)
            s = 1; sg = p.seg[1]; tmA = &tmA1; kb = 0; tap = 0; cb = 0;
          }
          {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            const int kh = tap / sg.ksize;
            const int kw = tap - kh * sg.ksize;
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            if constexpr (CTA2) {
              // both CTAs' boxes complete on the LEADER's full barrier, which expects the bytes of the pair
              if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
              const uint32_t fb = full0b + static_cast<uint32_t>(stage) * 8u;
              ptx::tma_load_4d_2sm(sa, tmA, fb, cb * BK, w0 * sg.stride + kw - sg.pad, h0 * sg.stride + kh - sg.pad,
                                   n0);
              if (p.b_mn) {
                // adjoint: this CTA's half of the N columns as BN/128 MN-major atoms of [64 K rows x 64 columns]
                const int tcol = (sg.ksize * sg.ksize - 1 - tap) * p.N + nt * BN + static_cast<int>(cta_rank) * (BN / 2);
#pragma unroll
                for (int g = 0; g < BN / 128; ++g)
                  ptx::tma_load_2d_2sm(sa + C::A_BYTES + g * 8192, &tmB, fb, tcol + g * 64, cb * BK);
              } else {
                ptx::tma_load_2d_2sm(sa + C::A_BYTES, &tmB, fb, sg.bkoff + pkoff + kb * BK,
                                     nt * BN + static_cast<int>(cta_rank) * (BN / 2));
              }
              if (++cb == sg.cblk) { cb = 0; ++tap; }
              if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
              continue;
            }
            ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            ptx::tma_load_4d(sa, tmA, &full_bar[stage], cb * BK, w0 * sg.stride + kw - sg.pad,
                             h0 * sg.stride + kh - sg.pad, n0);
            if (p.b_img) {
              // per-image B: the (token, head) rows of image n0 gathered by one 4-D box per [64 x 64] / [BN x 64] atom
              if (p.b_mn) {
#pragma unroll
                for (int g = 0; g < BN / 64; ++g)
                  ptx::tma_load_4d(sa + C::A_BYTES + g * 8192, &tmB, &full_bar[stage], nt * BN + g * 64, 0,
                                   (p.b_img2 ? (kb & 1) : kb) * (BK / p.b_img), n0 + (p.b_img2 ? (kb >> 1) : 0));
              } else {
                ptx::tma_load_4d(sa + C::A_BYTES, &tmB, &full_bar[stage], kb * BK, 0, 0, n0 + (p.b_img2 ? nt : 0));
              }
            } else if (p.b_mn) {
              // adjoint: rows = 64 source channels (K), columns = N of the flipped tap, 64 at a time (one SW128 atom)
              const int tcol = (sg.ksize * sg.ksize - 1 - tap) * p.N + nt * BN;
#pragma unroll
              for (int g = 0; g < BN / 64; ++g)
                ptx::tma_load_2d(sa + C::A_BYTES + g * 8192, &tmB, &full_bar[stage], tcol + g * 64, cb * BK);
            } else {
              ptx::tma_load_2d(sa + C::A_BYTES, &tmB, &full_bar[stage], sg.bkoff + pkoff + kb * BK, nt * BN);
            }
            if (++cb == sg.cblk) { cb = 0; ++tap; }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (CTA pairs: the leader only) =====================
    const uint32_t idesc =
        ptx::make_idesc_bf16(CTA2 ? 2 * BM : BM, BN) | (p.b_mn ? (1u << 16) : 0u);  // bit 16: B is MN-major
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = 0; item < n_items && leader; ++item) {
      int tile, kb0, kb1, kind;
      get_item(item, tile, kb0, kb1, kind);
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
      if (kind == 2) ptx::mbar_wait(seed_bar, 0);   // the parked accumulator is back in TMEM (both CTAs of a pair)
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      const uint32_t resumed = kind == 2 ? 1u : 0u;
      for (int kbt = kb0; kbt < kb1; ++kbt) {
        const int kb = (kbt - kb0) | static_cast<int>(resumed);   // 0 only for the MMA that starts a fresh accumulator
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t adesc = ptx::make_sw128_kmajor_desc(sa);
          if constexpr (CTA2) {
            const uint64_t bdesc = p.b_mn ? ptx::make_sw128_mnmajor_desc(sa + C::A_BYTES, 8192)
                                          : ptx::make_sw128_kmajor_desc(sa + C::A_BYTES);
            const uint64_t bstep = p.b_mn ? 128 : 2;  // 16 K rows of an MN-major atom = 2048 B, of a K-major one 32 B
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              ptx::umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + bstep * k, idesc, (kb | k) != 0 ? 1u : 0u);
            ptx::umma_commit_2sm(&empty_bar[stage], 3);  // frees this stage in both CTAs
            if (kbt == kb1 - 1) ptx::umma_commit_2sm(&tmem_full[acc], 3);
          } else if (p.b_mn) {
            // MN-major B: 64-column groups 8192 B apart (LBO), 8-row K groups 1024 B apart; 16 K rows = 2048 B
            const uint64_t bdesc = ptx::make_sw128_mnmajor_desc(sa + C::A_BYTES, 8192);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 128 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          } else {
            const uint64_t bdesc = ptx::make_sw128_kmajor_desc(sa + C::A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advancing K by 16 bf16 = 32 B inside the swizzle row = +2 in 16-byte units
              ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (!CTA2) {
            ptx::umma_commit(&empty_bar[stage]);
            if (kbt == kb1 - 1) ptx::umma_commit(&tmem_full[acc]);
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp == 2 + EPI_WARPS) {
    // ===================== row statistics of a folded LayerNorm (consumer side) =====================
    // For every tile, one or two tiles ahead of the epilogue: fold the producer's per-32-column {sum, sum of squares}
    // partials of the tile's 128 rows (four rows per lane, fixed order) into {rstd, -rstd * mean} in shared memory,
    // so that no epilogue thread waits for global memory at the top of a tile.
    if (p.ln_in != nullptr) {
      float2* rows_s = reinterpret_cast<float2*>(smem + C::EPI_OFF + EPI_WARP_BYTES + 2 * EBUF_BYTES);
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        int mt, nt;
        tile_mn(tile, mt, nt);
        ptx::mbar_wait(&ln_empty[acc], acc_phase ^ 1u);
#pragma unroll 1
        for (int j = 0; j < BM / 32; ++j) {
          int row = mt * BM + j * 32 + lane;   // linear geometry: tile rows are consecutive output pixels
          if (row >= p.Wout) row = p.Wout - 1;
          const float2* lp = p.ln_in + static_cast<long long>(row) * p.ln_in_parts;
          float s = 0.f, ss = 0.f;
          if ((p.ln_in_parts & 7) == 0) {
            const float4* lp4 = reinterpret_cast<const float4*>(lp);
            for (int i0 = 0; 2 * i0 < p.ln_in_parts; i0 += 4) {   // 8 partials per round, four 16-byte loads in flight
              const float4 q0 = __ldg(lp4 + i0), q1 = __ldg(lp4 + i0 + 1), q2 = __ldg(lp4 + i0 + 2),
                           q3 = __ldg(lp4 + i0 + 3);
              s += ((q0.x + q0.z) + (q1.x + q1.z)) + ((q2.x + q2.z) + (q3.x + q3.z));
              ss += ((q0.y + q0.w) + (q1.y + q1.w)) + ((q2.y + q2.w) + (q3.y + q3.w));
            }
          } else {
            for (int i = 0; i < p.ln_in_parts; ++i) {
              const float2 t = __ldg(lp + i);
              s += t.x;
              ss += t.y;
            }
          }
          const float mu = s * p.ln_inv_c;
          const float var = fmaxf(ss * p.ln_inv_c - mu * mu, 0.f);
          const float r = rsqrtf(var + p.ln_eps);
          rows_s[acc * BM + j * 32 + lane] = make_float2(r, -r * mu);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ln_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int part = ew >> 2;   // 0: even units, 1: odd units
    const int etid = threadIdx.x - 64;
    uint8_t* ebuf = smem + C::EPI_OFF + ew * EPI_WARP_BYTES;
    uint8_t* o2buf = ebuf + 2 * EBUF_BYTES;
    float* bias_s = reinterpret_cast<float*>(smem + C::BIAS_OFF);
    // c1 of a folded LayerNorm, staged like the bias; lives in epilogue warp 0's shadow buffer (ln_in excludes out2)
    float* c1_s = reinterpret_cast<float*>(smem + C::EPI_OFF + 2 * EBUF_BYTES);
    // ... and the rows' {rstd, -rstd * mean} [2][BM] in epilogue warp 1's
    const float2* ln_rows_s = reinterpret_cast<const float2*>(smem + C::EPI_OFF + EPI_WARP_BYTES + 2 * EBUF_BYTES);
    uint64_t* rbar = res_bar + 2 * ew;
    // a unit: 32 accumulator columns (GEGLU: 64 -> 32 output columns)
    constexpr int UNITS = GEGLU ? BN / 64 : BN / 32;
    constexpr int OUT_PER_TILE = GEGLU ? BN / 2 : BN;
    // this warp's 32 rows inside the (bw, bh, bn) output box of the tile
    const int r0 = quad * 32;
    const int sw0 = r0 % p.bw, sh0 = (r0 / p.bw) % p.bh, sn0 = r0 / (p.bw * p.bh);
    const int my_dn = (r0 + lane) / (p.bw * p.bh);
    const uint32_t res_bytes = p.res_f32 ? 4096u : 2048u;

    const uint32_t tmem_empty0 = CTA2 ? ptx::mapa_u32(&tmem_empty[0], 0) : 0u;  // the leader's barriers
    auto unit_origin = [&](int tile, int& nt, int& w, int& h, int& n) {
      int mt;
      tile_mn(tile, mt, nt);
      w = (mt % p.tiles_w) * p.bw + sw0;
      h = ((mt / p.tiles_w) % p.tiles_h) * p.bh + sh0;
      n = (mt / (p.tiles_w * p.tiles_h)) * p.bn + sn0;
    };
    auto issue_res = [&](int tile, int ku, int b) {  // lane 0 only
      int nt, w, h, n;
      unit_origin(tile, nt, w, h, n);
      ptx::mbar_arrive_expect_tx(&rbar[b], res_bytes);
      ptx::tma_load_4d(ebuf + b * EBUF_BYTES, &tmRes, &rbar[b], nt * OUT_PER_TILE + ku * 32, w, h, n);
    };

    // LayerNorm applied here (Params::ln_gamma): gamma | beta live in epilogue warp 4's shadow buffer, the row-sum
    // exchange of a warp pair in the shadow buffer of its part-0 warp (no shadow output in this mode)
    float* const lng_s = reinterpret_cast<float*>(smem + C::EPI_OFF + 4 * EPI_WARP_BYTES + 2 * EBUF_BYTES);
    float2* const lnx_s = reinterpret_cast<float2*>(smem + C::EPI_OFF + (ew & 3) * EPI_WARP_BYTES + 2 * EBUF_BYTES);
    if constexpr (BN == 256 && !GEGLU) {
      if (p.ln_gamma != nullptr) {
        lng_s[etid] = __ldg(p.ln_gamma + etid);          // 256 epilogue threads, 256 channels
        lng_s[256 + etid] = __ldg(p.ln_beta + etid);     // (visible after the first tile's named barrier)
      }
    }
    uint32_t tcount = 0;
    uint32_t it = 0;  // units processed by this warp: buffer = it & 1, mbarrier parity = (it >> 1) & 1
    const int first_epi = head_tile >= 0 ? 1 : 0;   // (a parked beginning of a tile has no epilogue)
    if (!GEGLU && p.has_res && lane == 0 && part < UNITS && first_epi < n_items) {
      int t0, a0, a1, a2;
      get_item(first_epi, t0, a0, a1, a2);
      issue_res(t0, part, 0);
    }

    // ---- stream-K: parking / resuming an accumulator (Params::sk_tiles) ----
    // unit ku of the tile as [8][128 rows] float4: the thread that owns accumulator row r moves float4 j of its 32
    // columns to / from entry j * 128 + r, so a warp instruction covers 512 contiguous bytes
    constexpr int SK_CTA_F4 = (BN / 32) * 8 * BM;
    auto sk_park = [&](uint32_t taddr) {   // this CTA's 128 x BN accumulator -> sk_ws slot of this cluster
      float4* dst = p.sk_ws + static_cast<size_t>(tile_first * 2 + static_cast<int>(cta_rank)) * SK_CTA_F4 + quad * 32 + lane;
#pragma unroll 1
      for (int ku = part; ku < BN / 32; ku += 2) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + ku * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          __stcg(dst + (ku * 8 + j) * BM, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                      __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        unsigned int* fl = p.sk_flags + (tile_first * 2 + static_cast<int>(cta_rank)) * EPI_WARPS + ew;
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(fl), "r"(1u) : "memory");
      }
    };
    auto sk_resume = [&](int item) {   // the previous cluster's parked accumulator -> TMEM stage of `item`
      const int slot = (tile_first - 1) * 2 + static_cast<int>(cta_rank);
      unsigned int* fl = p.sk_flags + slot * EPI_WARPS + ew;
      if (lane == 0) {
        unsigned int f;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(fl) : "memory");
        } while (f == 0u);
      }
      __syncwarp();
      const float4* src = p.sk_ws + static_cast<size_t>(slot) * SK_CTA_F4 + quad * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>((item & 1) * BN);
#pragma unroll 1
      for (int ku = part; ku < BN / 32; ku += 2) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldcg(src + (ku * 8 + j) * BM);
          v[4 * j] = __float_as_uint(t.x); v[4 * j + 1] = __float_as_uint(t.y);
          v[4 * j + 2] = __float_as_uint(t.z); v[4 * j + 3] = __float_as_uint(t.w);
        }
        ptx::tmem_st_32x32(taddr + ku * 32, v);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        *reinterpret_cast<volatile unsigned int*>(fl) = 0u;   // (the next launch parks here again)
        if constexpr (CTA2) ptx::mbar_arrive_cluster(ptx::mapa_u32(seed_bar, 0));
        else ptx::mbar_arrive(seed_bar);
      }
    };
    // the resumed tile is the LAST item; its TMEM stage is free once item (last - 2) has been drained, so the load
    // runs at the top of iteration (last - 1), under that item's main loop -- unless that item is the parked beginning
    // (two items in all): then right behind the parking, so that no cluster ever waits before it has parked its own
    const int last = n_items - 1;
    const int resume_top = tail_tile < 0 ? -1 : (last == 0 ? 0 : ((head_tile >= 0 && last == 1) ? -1 : last - 1));
    const bool resume_after_park = tail_tile >= 0 && head_tile >= 0 && last == 1;

    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = 0; item < n_items; ++item) {
      int tile, kb0_, kb1_, kind;
      get_item(item, tile, kb0_, kb1_, kind);
      if (kind == 1) {
        ptx::mbar_wait(&tmem_full[acc], acc_phase);
        ptx::tc_fence_after();
        sk_park(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN));
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTA2) ptx::mbar_arrive_cluster(tmem_empty0 + static_cast<uint32_t>(acc) * 8u);
          else ptx::mbar_arrive(&tmem_empty[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
        if (resume_after_park) sk_resume(last);
        continue;
      }
      int nt, w, h, n;
      unit_origin(tile, nt, w, h, n);
      int uph = 0;   // upsampling phase (py, px) of this work item
      if (p.phases > 1) uph = (tile / p.n_tiles) & 3;
      // bias of this tile's BN columns -> shared memory (one value per epilogue thread)
      if (etid < BN) {
        const int col = nt * BN + etid;
        float bv0 = (p.bias != nullptr && col < p.N) ? __ldg(p.bias + col) : 0.f;
        if (GEGLU && (etid & 16) == 0) bv0 *= 0.5f;  // value columns: the epilogue computes 0.5 v + 0.5 bias (geglu2)
        // (one image per tile, no residual: the per-image row vector joins the bias here -- (0 + bias) + rowvec, the
        // epilogue's own order -- instead of four L2 round trips per tile inside the unit loop)
        if (!GEGLU && p.rowvec != nullptr && p.bn == 1 && !p.has_res && p.act == EALDM_ACT_NONE && col < p.N && n < p.Nimg)
          bv0 += __ldg(p.rowvec + static_cast<long long>(n) * p.ld_rowvec + col);
        bias_s[acc * BN + etid] = bv0;
        if (p.ln_in != nullptr) {
          float cv = col < p.N ? __ldg(p.ln_c1 + col) : 0.f;
          if (GEGLU && (etid & 16) == 0) cv *= 0.5f;
          c1_s[acc * BN + etid] = cv;
        }
      }
      // LayerNorm folded into this GEMM: {rstd, -rstd * mean} of this thread's row, prepared by the statistics warp
      // one or two tiles ahead (out = ln_r * acc + ln_nm * c1 + bias)
      float ln_r = 1.f, ln_nm = 0.f;
      if (p.ln_in != nullptr) {
        ptx::mbar_wait(&ln_full[acc], acc_phase);
        const float2 st = ln_rows_s[acc * BM + quad * 32 + lane];
        ln_r = st.x;
        ln_nm = st.y;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ln_empty[acc]);
      }
      ptx::named_bar_sync(1, 32 * EPI_WARPS);
      // (behind the barrier: every epilogue warp has finished reading the stage's previous accumulator)
      if (item == resume_top) sk_resume(last);
      const float* bs = bias_s + acc * BN;
      const float* cs = c1_s + acc * BN;
      const float* rv = nullptr;
      if (!GEGLU && p.rowvec != nullptr && !(p.bn == 1 && !p.has_res && p.act == EALDM_ACT_NONE)) {
        int img = (n - sn0) + my_dn;
        if (img >= p.Nimg) img = p.Nimg - 1;
        rv = p.rowvec + static_cast<long long>(img) * p.ld_rowvec + nt * BN;
      }
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      // EALDM_GNA_DEBUG & 16: clock64 trace of this epilogue's phases (first CTA, first epilogue warp), printed per tile
      const bool trace = (p.gna_debug & 16) != 0 && blockIdx.x == 0 && ew == 0 && lane == 0;
      long long tq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (trace) tq[0] = clock64();

      bool wide_done = false;
      float lna_s = 0.f, lna_ss = 0.f;   // LayerNorm applied here: this thread's row sums over the units of its warp
      if constexpr (BN == 256) {
        if (p.wide) {
          wide_done = true;
          if constexpr (GEGLU) {
            // one pass per warp and tile: 128 accumulator columns -> 64 outputs = one [32 rows x 128 B] box
            if (lane == 0) ptx::bulk_wait_read<0>();
            __syncwarp();
            uint32_t v[2][32];
            ptx::tmem_ld_32x32(taddr0 + (part * 4) * 32, v[0]);
            // value = 0.5 (ln_r acc + ln_nm c1 + bias), gate = ln_r acc + ln_nm c1 + bias (without a folded LayerNorm
            // ln_r = 1 and the c1 term is skipped; the staged value-column constants are pre-halved)
            const bool ln = p.ln_in != nullptr;
            const uint64_t half2 = pk2(0.5f * ln_r, 0.5f * ln_r), one2 = pk2(ln_r, ln_r), nm2 = pk2(ln_nm, ln_nm);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int ch = part * 4 + c;
              ptx::tmem_ld_wait();
              if (c + 1 < 4) ptx::tmem_ld_32x32(taddr0 + (ch + 1) * 32, v[(c + 1) & 1]);  // flies under the math
              const uint32_t(&vc)[32] = v[c & 1];
              float o[16];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 t = *reinterpret_cast<const float4*>(bs + ch * 32 + 4 * j);
                const float4 g = *reinterpret_cast<const float4*>(bs + ch * 32 + 16 + 4 * j);
                uint64_t tv0 = pk2(t.x, t.y), tv1 = pk2(t.z, t.w), tg0 = pk2(g.x, g.y), tg1 = pk2(g.z, g.w);
                if (ln) {
                  const float4 ct = *reinterpret_cast<const float4*>(cs + ch * 32 + 4 * j);
                  const float4 cg = *reinterpret_cast<const float4*>(cs + ch * 32 + 16 + 4 * j);
                  tv0 = fma2(nm2, pk2(ct.x, ct.y), tv0);
                  tv1 = fma2(nm2, pk2(ct.z, ct.w), tv1);
                  tg0 = fma2(nm2, pk2(cg.x, cg.y), tg0);
                  tg1 = fma2(nm2, pk2(cg.z, cg.w), tg1);
                }
                const uint64_t val0 = fma2(pk2(__uint_as_float(vc[4 * j]), __uint_as_float(vc[4 * j + 1])), half2, tv0);
                const uint64_t val1 =
                    fma2(pk2(__uint_as_float(vc[4 * j + 2]), __uint_as_float(vc[4 * j + 3])), half2, tv1);
                const uint64_t g0 =
                    fma2(pk2(__uint_as_float(vc[16 + 4 * j]), __uint_as_float(vc[17 + 4 * j])), one2, tg0);
                const uint64_t g1 =
                    fma2(pk2(__uint_as_float(vc[18 + 4 * j]), __uint_as_float(vc[19 + 4 * j])), one2, tg1);
                upk2(geglu2<GEGLU>(val0, g0), o[4 * j], o[4 * j + 1]);
                upk2(geglu2<GEGLU>(val1, g1), o[4 * j + 2], o[4 * j + 3]);
              }
              sts_chunk_bf16_sw128(ebuf, lane, 2 * c, &o[0]);
              sts_chunk_bf16_sw128(ebuf, lane, 2 * c + 1, &o[8]);
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_4d(&tmOut, ebuf, nt * OUT_PER_TILE + part * 64, w, h, n);
              ptx::bulk_commit();
            }
          } else {
            // a pass = two [32 rows x 128 B] boxes: 2 x 32 fp32 columns or 2 x 64 bf16 columns
            const int cpp = p.out_f32 ? 2 : 4;  // 32-column accumulator chunks per pass
#pragma unroll 1
            for (int ps = part; ps < (BN / 32) / cpp; ps += 2) {
              if (lane == 0) ptx::bulk_wait_read<0>();
              __syncwarp();
              uint32_t v[2][32];
              ptx::tmem_ld_32x32(taddr0 + (ps * cpp) * 32, v[0]);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (c < cpp) {
                  const int ku = ps * cpp + c;
                  ptx::tmem_ld_wait();
                  if (c + 1 < cpp) ptx::tmem_ld_32x32(taddr0 + (ku + 1) * 32, v[(c + 1) & 1]);
                  const uint32_t(&vc)[32] = v[c & 1];
                  float r[32];
                  if (p.act == EALDM_ACT_NONE) {  // (bias + row vector) + accumulator: the narrow path's order
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const float4 t = *reinterpret_cast<const float4*>(bs + ku * 32 + 4 * j);
                      r[4 * j] = 0.f + t.x; r[4 * j + 1] = 0.f + t.y; r[4 * j + 2] = 0.f + t.z; r[4 * j + 3] = 0.f + t.w;
                    }
                    if (rv != nullptr) {
#pragma unroll
                      for (int j = 0; j < 8; ++j) {
                        if (nt * BN + ku * 32 + 4 * j < p.N) {
                          const float4 t = __ldg(reinterpret_cast<const float4*>(rv + ku * 32 + 4 * j));
                          r[4 * j] += t.x; r[4 * j + 1] += t.y; r[4 * j + 2] += t.z; r[4 * j + 3] += t.w;
                        }
                      }
                    }
                    if (p.ln_in != nullptr) {   // folded LayerNorm: ln_r * acc + ln_nm * c1 + bias
#pragma unroll
                      for (int j = 0; j < 8; ++j) {
                        const float4 t = *reinterpret_cast<const float4*>(cs + ku * 32 + 4 * j);
                        r[4 * j] = fmaf(ln_nm, t.x, r[4 * j]); r[4 * j + 1] = fmaf(ln_nm, t.y, r[4 * j + 1]);
                        r[4 * j + 2] = fmaf(ln_nm, t.z, r[4 * j + 2]); r[4 * j + 3] = fmaf(ln_nm, t.w, r[4 * j + 3]);
                      }
#pragma unroll
                      for (int j = 0; j < 32; ++j) r[j] = fmaf(__uint_as_float(vc[j]), ln_r, r[j]);
                    } else {
#pragma unroll
                      for (int j = 0; j < 32; ++j) r[j] += __uint_as_float(vc[j]);
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                      float f = __uint_as_float(vc[j]) + bs[ku * 32 + j];
                      if (rv != nullptr && nt * BN + ku * 32 + j < p.N) f += __ldg(rv + ku * 32 + j);
                      r[j] = p.act == EALDM_ACT_RELU ? fmaxf(f, 0.f) : silu_f(f);
                    }
                  }
                  if (p.gn_partial != nullptr) {
                    if (p.gn_quads) gn_partial_unit_quads(p, r, lane, (n - sn0) + my_dn < p.Nimg, n, h, w, nt * BN + ku * 32);
                    else gn_partial_unit(p, r, lane, (n - sn0) + my_dn < p.Nimg, n, h, w, nt * BN + ku * 32);
                  }
                  if (p.out_f32) {
                    sts_row_f32(ebuf + c * EBUF_BYTES, lane, r);
                  } else {
                    uint8_t* box = ebuf + (c >> 1) * EBUF_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; ++j) sts_chunk_bf16_sw128(box, lane, 4 * (c & 1) + j, &r[8 * j]);
                  }
                }
              }
              ptx::fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                const int c0 = nt * BN + ps * cpp * 32;
                ptx::tma_store_4d(&tmOut, ebuf, c0, w, h, n);
                ptx::tma_store_4d(&tmOut, ebuf + EBUF_BYTES, c0 + (p.out_f32 ? 32 : 64), w, h, n);
                ptx::bulk_commit();
              }
            }
          }
        }
      }

#pragma unroll 1
      for (int ku = part; ku < UNITS && !wide_done; ku += 2) {
        const int b = it & 1;
        uint8_t* eb = ebuf + b * EBUF_BYTES;
        if constexpr (GEGLU) {
          if (lane == 0) {  // buffer b was last read by the store issued two units ago
            if (p.relaxed_wait) ptx::bulk_wait_read<1>();
            else ptx::bulk_wait_read<0>();
          }
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int ch = 2 * ku + half;
            uint32_t v[32];
            ptx::tmem_ld_32x32(taddr0 + ch * 32, v);
            uint64_t bv[8], bg[8];
            const uint64_t nm2 = pk2(ln_nm, ln_nm);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 t = *reinterpret_cast<const float4*>(bs + ch * 32 + 4 * j);
              bv[2 * j] = pk2(t.x, t.y); bv[2 * j + 1] = pk2(t.z, t.w);
              const float4 g = *reinterpret_cast<const float4*>(bs + ch * 32 + 16 + 4 * j);
              bg[2 * j] = pk2(g.x, g.y); bg[2 * j + 1] = pk2(g.z, g.w);
              if (p.ln_in != nullptr) {   // folded LayerNorm: + ln_nm * c1 (value-column constants are pre-halved)
                const float4 ct = *reinterpret_cast<const float4*>(cs + ch * 32 + 4 * j);
                bv[2 * j] = fma2(nm2, pk2(ct.x, ct.y), bv[2 * j]); bv[2 * j + 1] = fma2(nm2, pk2(ct.z, ct.w), bv[2 * j + 1]);
                const float4 cg = *reinterpret_cast<const float4*>(cs + ch * 32 + 16 + 4 * j);
                bg[2 * j] = fma2(nm2, pk2(cg.x, cg.y), bg[2 * j]); bg[2 * j + 1] = fma2(nm2, pk2(cg.z, cg.w), bg[2 * j + 1]);
              }
            }
            ptx::tmem_ld_wait();
            float o[16];
            const uint64_t half2 = pk2(0.5f * ln_r, 0.5f * ln_r), one2 = pk2(ln_r, ln_r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint64_t val =
                  fma2(pk2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), half2, bv[j]);
              const uint64_t g = fma2(pk2(__uint_as_float(v[16 + 2 * j]), __uint_as_float(v[17 + 2 * j])), one2, bg[j]);
              upk2(geglu2<GEGLU>(val, g), o[2 * j], o[2 * j + 1]);
            }
            sts_chunk_bf16(eb, lane, 2 * half, &o[0]);
            sts_chunk_bf16(eb, lane, 2 * half + 1, &o[8]);
          }
        } else {
          float r[32];
          if (p.has_res) {
            ptx::mbar_wait(&rbar[b], (it >> 1) & 1u);
            if (p.res_f32) lds_row_f32(eb, lane, r);
            else lds_row_bf16(eb, lane, r);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0.f;
          }
          if (lane == 0) {
            // the other unit buffer (next residual prefetch) and the single shadow buffer are free once their stores
            // have read them; without either, only the store issued two units ago (same buffer) must be done
            if (p.relaxed_wait && !p.has_res && !p.has_out2) ptx::bulk_wait_read<1>();
            else ptx::bulk_wait_read<0>();
            if (p.has_res) {
              if (ku + 2 < UNITS) issue_res(tile, ku + 2, b ^ 1);
              else if (item + 1 < n_items && p.gna_gamma == nullptr) {   // (GroupNorm epilogue: issued after its pass 2)
                int t1, a0, a1, a2;
                get_item(item + 1, t1, a0, a1, a2);
                issue_res(t1, part, b ^ 1);
              }
            }
          }
          __syncwarp();
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr0 + ku * 32, v);
          if (p.act == EALDM_ACT_NONE) {  // bias (+ per-image row vector) folded into r while the load flies
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = *reinterpret_cast<const float4*>(bs + ku * 32 + 4 * j);
              r[4 * j] += t.x; r[4 * j + 1] += t.y; r[4 * j + 2] += t.z; r[4 * j + 3] += t.w;
            }
            if (rv != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (nt * BN + ku * 32 + 4 * j < p.N) {
                  const float4 t = __ldg(reinterpret_cast<const float4*>(rv + ku * 32 + 4 * j));
                  r[4 * j] += t.x; r[4 * j + 1] += t.y; r[4 * j + 2] += t.z; r[4 * j + 3] += t.w;
                }
              }
            }
            ptx::tmem_ld_wait();
            if (p.ln_in != nullptr) {   // folded LayerNorm: ln_r * acc + ln_nm * c1 + bias
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = *reinterpret_cast<const float4*>(cs + ku * 32 + 4 * j);
                r[4 * j] = fmaf(ln_nm, t.x, r[4 * j]); r[4 * j + 1] = fmaf(ln_nm, t.y, r[4 * j + 1]);
                r[4 * j + 2] = fmaf(ln_nm, t.z, r[4 * j + 2]); r[4 * j + 3] = fmaf(ln_nm, t.w, r[4 * j + 3]);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = fmaf(__uint_as_float(v[j]), ln_r, r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] += __uint_as_float(v[j]);
            }
            if (p.ln_out != nullptr) {  // producer of a folded LayerNorm: row sums of the fp32 result
              float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (nt * BN + ku * 32 + j < p.N) { a4[j & 3] += r[j]; b4[j & 3] = fmaf(r[j], r[j], b4[j & 3]); }
              }
              // one partial per 32-column unit, whatever the N tile: every schedule hands the consumer the same sums
              if (n < p.Nimg && w + lane < p.Wout)
                p.ln_out[static_cast<long long>(w + lane) * p.ln_out_parts + nt * (BN / 32) + ku] =
                    make_float2((a4[0] + a4[1]) + (a4[2] + a4[3]), (b4[0] + b4[1]) + (b4[2] + b4[3]));
            }
            if constexpr (BN == 256 && !GEGLU) {
              if (p.ln_gamma != nullptr) {   // LayerNorm applied here, pass 1: row sums, result back over the accumulator
                float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
                uint32_t u[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  a4[j & 3] += r[j];
                  b4[j & 3] = fmaf(r[j], r[j], b4[j & 3]);
                  u[j] = __float_as_uint(r[j]);
                }
                lna_s += (a4[0] + a4[1]) + (a4[2] + a4[3]);
                lna_ss += (b4[0] + b4[1]) + (b4[2] + b4[3]);
                ptx::tmem_st_32x32(taddr0 + ku * 32, u);
              }
            }
          } else if (BN <= 128 && p.act == EALDM_ACT_SOFTMAX4) {
            // collapsed cross-attention: the 32 columns are (head, key) base-2 logits; softmax over the 4 keys of a head
            ptx::tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float f0 = __uint_as_float(v[4 * g]) + bs[ku * 32 + 4 * g];
              const float f1 = __uint_as_float(v[4 * g + 1]) + bs[ku * 32 + 4 * g + 1];
              const float f2 = __uint_as_float(v[4 * g + 2]) + bs[ku * 32 + 4 * g + 2];
              const float f3 = __uint_as_float(v[4 * g + 3]) + bs[ku * 32 + 4 * g + 3];
              const float mx = fmaxf(fmaxf(f0, f1), fmaxf(f2, f3));
              const float e0 = ex2_approx(f0 - mx), e1 = ex2_approx(f1 - mx), e2 = ex2_approx(f2 - mx),
                          e3 = ex2_approx(f3 - mx);
              const float inv = __fdividef(1.0f, (e0 + e1) + (e2 + e3));
              r[4 * g] = e0 * inv; r[4 * g + 1] = e1 * inv; r[4 * g + 2] = e2 * inv; r[4 * g + 3] = e3 * inv;
            }
            if (p.b_img2 && (quad >> 1) != nt) {   // rows of the tile's other image: their half of the K range is zero
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = 0.f;
            }
          } else {  // SiLU (time-embedding MLP): act(acc + bias + rowvec) + residual
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float f = __uint_as_float(v[j]) + bs[ku * 32 + j];
              if (rv != nullptr && nt * BN + ku * 32 + j < p.N) f += __ldg(rv + ku * 32 + j);
              r[j] += p.act == EALDM_ACT_RELU ? fmaxf(f, 0.f) : silu_f(f);
            }
          }
          if (p.gn_partial != nullptr) {
            if (p.gn_quads)
              gn_partial_unit_quads(p, r, lane, (n - sn0) + my_dn < p.Nimg, n, h, w, nt * BN + ku * 32,
                                    uph * p.gn_phase_chunks);
            else
              gn_partial_unit(p, r, lane, (n - sn0) + my_dn < p.Nimg, n, h, w, nt * BN + ku * 32,
                              uph * p.gn_phase_chunks);
          }
          if constexpr (BN == 256 && !GEGLU) {
            if (p.gna_gamma != nullptr) {   // GroupNorm applied here, pass 1: the result back over the accumulator
              uint32_t u[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(r[j]);
              ptx::tmem_st_32x32(taddr0 + ku * 32, u);
              if (p.gna_only) {   // (no un-normalised output at all)
                ++it;
                continue;
              }
            }
          }
          if (p.out_f32) sts_row_f32(eb, lane, r);
          else sts_row_bf16(eb, lane, r);
          if (p.has_out2) sts_row_bf16(o2buf, lane, r);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          const int c0 = nt * OUT_PER_TILE + ku * 32;
          if (!GEGLU && uph != 0) {  // this unit's pixels of output phase (py, px): the phase's own strided map
            ptx::tma_store_4d(&pm.out[uph - 1], eb, c0, w, h, n);
            if (p.has_out2) ptx::tma_store_4d(&pm.out2[uph - 1], o2buf, c0, w, h, n);
          } else {
            ptx::tma_store_4d(&tmOut, eb, c0, w, h, n);
            if (!GEGLU && p.has_out2) ptx::tma_store_4d(&tmOut2, o2buf, c0, w, h, n);
          }
          ptx::bulk_commit();
        }
        ++it;
      }
      if constexpr (BN == 256 && !GEGLU) {
        if (p.ln_gamma != nullptr) {
          // ---- LayerNorm applied here, pass 2 ----
          ptx::tmem_st_wait();
          float2* const xs = lnx_s + (tcount & 1u) * 64;       // [tile parity][part][lane]
          xs[part * 32 + lane] = make_float2(lna_s, lna_ss);
          ptx::named_bar_sync(2 + (ew & 3), 64);               // the two warps that share these 32 rows
          const float2 o = xs[(part ^ 1) * 32 + lane];
          const float s_all = part == 0 ? lna_s + o.x : o.x + lna_s;      // part 0 + part 1, whichever warp adds
          const float ss_all = part == 0 ? lna_ss + o.y : o.y + lna_ss;
          const float mu = s_all * (1.0f / 256.0f);
          const float var = fmaxf(ss_all * (1.0f / 256.0f) - mu * mu, 0.f);
          const float rstd = rsqrtf(var + p.ln_apply_eps);
          const float nmu = -mu * rstd;
          // staging: the unit buffer of pass 1's last unit, as two bf16 [32 rows x 64 B] halves (the OTHER unit buffer
          // may already hold the next tile's first residual unit); free once every pass-1 store has read its box
          uint8_t* const fb = ebuf + ((it - 1u) & 1u) * EBUF_BYTES;
          if (lane == 0) ptx::bulk_wait_read<0>();
          __syncwarp();
          uint32_t vv[2][32];
          ptx::tmem_ld_32x32(taddr0 + part * 32, vv[0]);
#pragma unroll
          for (int ui = 0; ui < UNITS / 2; ++ui) {
            const int ku = part + 2 * ui;
            ptx::tmem_ld_wait();
            if (ui + 1 < UNITS / 2) ptx::tmem_ld_32x32(taddr0 + (ku + 2) * 32, vv[(ui + 1) & 1]);   // flies under the math
            const uint32_t(&v)[32] = vv[ui & 1];
            float y[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g4 = *reinterpret_cast<const float4*>(lng_s + ku * 32 + 4 * j);
              const float4 b4 = *reinterpret_cast<const float4*>(lng_s + 256 + ku * 32 + 4 * j);
              y[4 * j] = fmaf(fmaf(__uint_as_float(v[4 * j]), rstd, nmu), g4.x, b4.x);
              y[4 * j + 1] = fmaf(fmaf(__uint_as_float(v[4 * j + 1]), rstd, nmu), g4.y, b4.y);
              y[4 * j + 2] = fmaf(fmaf(__uint_as_float(v[4 * j + 2]), rstd, nmu), g4.z, b4.z);
              y[4 * j + 3] = fmaf(fmaf(__uint_as_float(v[4 * j + 3]), rstd, nmu), g4.w, b4.w);
            }
            uint8_t* const sb = fb + (ui & 1) * O2BUF_BYTES;
            if (ui >= 2) {   // the store issued two units ago has read this half
              if (lane == 0) ptx::bulk_wait_read<1>();
              __syncwarp();
            }
            sts_row_bf16(sb, lane, y);
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_4d(&tmOut2, sb, nt * BN + ku * 32, w, h, n);
              ptx::bulk_commit();
            }
          }
        }
      }
      if constexpr (BN == 256 && !GEGLU) {
        if (p.gna_gamma != nullptr) {
          // ---- GroupNorm applied here, pass 2 (Params::gna_gamma) ----
          if (trace) tq[1] = clock64();
          ptx::tmem_st_wait();
          ptx::tc_fence_before();   // (pass 2 reads columns that the other warp of this lane quadrant stored)
          __threadfence();          // this warp's partials are visible before it counts itself in
          __syncwarp();
          if (trace) tq[2] = clock64();
          if (n < p.Nimg) {     // (n: the image of this warp's 32 rows)
            unsigned int* cnt = p.gna_counters + 2 * (static_cast<size_t>(n) * p.n_tiles + nt);
            // lane = channel octet of this tile's 256 columns; its gamma / beta are requested before the wait
            const int oct = nt * (BN / 8) + lane;
            const bool live = oct * 8 < p.N;
            float4 gq[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)}, bq[2] = {gq[0], gq[0]};
            if (live) {
              gq[0] = __ldg(reinterpret_cast<const float4*>(p.gna_gamma + oct * 8));
              gq[1] = __ldg(reinterpret_cast<const float4*>(p.gna_gamma + oct * 8) + 1);
              bq[0] = __ldg(reinterpret_cast<const float4*>(p.gna_beta + oct * 8));
              bq[1] = __ldg(reinterpret_cast<const float4*>(p.gna_beta + oct * 8) + 1);
            }
            if (lane == 0 && !(p.gna_debug & 8)) {
              unsigned int f = atomicAdd(cnt, 1u) + 1u, spins = 0;
              while (f < p.gna_expected && !(p.gna_debug & 1)) {
                __nanosleep(200);   // (a busy spin measurably slows the MMA / TMA warps of the same CTA)
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(cnt) : "memory");
                if (++spins > (1u << 27)) {
                  printf("ealdm: GroupNorm epilogue waited too long for image %d\n", n);
                  __trap();
                }
              }
              __threadfence();   // (the counter was read with acquire semantics in the loop; this covers the atomic's value)
            }
            __syncwarp();
            if (trace) tq[3] = clock64();
            // fold the octet's partials over the image's 32-pixel chunks.  fp32 pairwise sums (fp64 adds and the fp64
            // reciprocal square root cost ~10,000 clk per tile here -- clock64 trace, DESIGN.md section 4 --, five times
            // the whole second pass); only mean^2 is subtracted in fp64
            float s = 0.f, ss = 0.f;
            if (live && !(p.gna_debug & 2)) {
              const float2* pp = p.gn_partial + static_cast<long long>(n) * p.gn_chunks * p.gn_ld + oct;
              // up to 32 loads in flight per round (one L2 round trip per round, not per chunk); fixed summation tree
              for (int c0 = 0; c0 < p.gn_chunks; c0 += 32) {
                float2 t[32];
#pragma unroll
                for (int u = 0; u < 32; ++u)
                  t[u] = c0 + u < p.gn_chunks ? __ldcg(pp + static_cast<long long>(c0 + u) * p.gn_ld) : make_float2(0.f, 0.f);
#pragma unroll
                for (int w2 = 16; w2 >= 1; w2 >>= 1) {
#pragma unroll
                  for (int u = 0; u < w2; ++u) {
                    t[u].x += t[u + w2].x;
                    t[u].y += t[u + w2].y;
                  }
                }
                s += t[0].x;
                ss += t[0].y;
              }
            }
            for (int d = 1; d < p.gna_octets; d <<= 1) {   // the octets of one group, lower octet first
              const float os = __shfl_xor_sync(0xffffffffu, s, d), oss = __shfl_xor_sync(0xffffffffu, ss, d);
              s = (lane & d) ? os + s : s + os;
              ss = (lane & d) ? oss + ss : ss + oss;
            }
            const double mean = static_cast<double>(s) * p.gna_inv_count;
            const float var = fmaxf(static_cast<float>(static_cast<double>(ss) * p.gna_inv_count - mean * mean), 0.f);
            const float rstd = rsqrtf(var + p.gna_eps);
            const float mu = static_cast<float>(mean);
            float2* const tab = reinterpret_cast<float2*>(o2buf);   // [256 columns]{scale, shift} (no shadow output here)
            if (live) {
              const float gv[8] = {gq[0].x, gq[0].y, gq[0].z, gq[0].w, gq[1].x, gq[1].y, gq[1].z, gq[1].w};
              const float bv[8] = {bq[0].x, bq[0].y, bq[0].z, bq[0].w, bq[1].x, bq[1].y, bq[1].z, bq[1].w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float a = rstd * gv[j];
                tab[lane * 8 + j] = make_float2(a, fmaf(-mu, a, bv[j]));
              }
            }
            __syncwarp();
            if (lane == 0 && !(p.gna_debug & 8) && atomicAdd(cnt + 1, 1u) == p.gna_expected - 1u) {   // the last reader re-arms the pair
              cnt[1] = 0u;
              __threadfence();
              cnt[0] = 0u;
            }
            if (trace) tq[4] = clock64();
            const CUtensorMap* const tmN = p.gna_only ? &tmOut : &tmOut2;
            // this warp normalises the 128 columns [part * 128, +128) of its 32 rows (pass 1 owned every second unit:
            // the other warp of the quadrant has counted in, so its TMEM stores are complete) and stores them as two
            // [32 rows x 64 columns] boxes of 128-byte rows, one from each unit buffer
            ptx::tc_fence_after();
            if (lane == 0) ptx::bulk_wait_read<0>();
            __syncwarp();
            uint32_t vv[2][32];
            ptx::tmem_ld_32x32(taddr0 + (part * 4) * 32, vv[0]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int ku = part * 4 + u;
              ptx::tmem_ld_wait();
              if (u + 1 < 4) ptx::tmem_ld_32x32(taddr0 + (ku + 1) * 32, vv[(u + 1) & 1]);
              if (p.gna_debug & 4) continue;
              const uint32_t(&v)[32] = vv[u & 1];
              float y[32];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 c2 = *reinterpret_cast<const float4*>(tab + ku * 32 + 2 * j);   // two columns' {scale, shift}
                y[2 * j] = fmaf(__uint_as_float(v[2 * j]), c2.x, c2.y);
                y[2 * j + 1] = fmaf(__uint_as_float(v[2 * j + 1]), c2.z, c2.w);
              }
              if (p.gna_silu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {   // x sigmoid(x) = 0.5 x (1 + tanh(0.5 x)): one MUFU per element
                  const float hx = 0.5f * y[j];
                  y[j] = fmaf(hx, tanh_approx(hx), hx);
                }
              }
              // the unit buffer the NEXT tile's first residual unit will use (it & 1) is filled and stored first
              uint8_t* const box = ebuf + (((it + (u >> 1)) & 1u)) * EBUF_BYTES;
#pragma unroll
              for (int j = 0; j < 4; ++j) sts_chunk_bf16_sw128(box, lane, 4 * (u & 1) + j, &y[8 * j]);
              if (u & 1) {
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  ptx::tma_store_4d(tmN, box, nt * BN + part * 128 + (u >> 1) * 64, w, h, n);
                  ptx::bulk_commit();
                }
              }
            }
          }
          if (trace) {
            tq[5] = clock64();
            printf("gna trace tile %d: pass1 %lld  st_wait+fence %lld  count+wait %lld  fold+table %lld  pass2 %lld clk\n", tile,
                   tq[1] - tq[0], tq[2] - tq[1], tq[3] - tq[2], tq[4] - tq[3], tq[5] - tq[4]);
          }
          // the next tile's first residual unit (pass 1 left it to us): its buffer is free once the FIRST box is read
          if (p.has_res && lane == 0 && item + 1 < n_items) {
            ptx::bulk_wait_read<1>();
            int t1, a0, a1, a2;
            get_item(item + 1, t1, a0, a1, a2);
            issue_res(t1, part, static_cast<int>(it & 1u));
          }
        }
      }
      if (p.gna_debug >= 1000) __nanosleep(static_cast<unsigned>(p.gna_debug));   // timing experiment: a longer epilogue
      ++tcount;
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTA2) ptx::mbar_arrive_cluster(tmem_empty0 + static_cast<uint32_t>(acc) * 8u);
        else ptx::mbar_arrive(&tmem_empty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) ptx::bulk_wait_read<0>();  // shared memory must outlive the last TMA stores
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) ptx::cluster_sync_all();  // neither CTA may leave (or free TMEM) while the pair still works
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (CTA2) ptx::tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
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

static int env_int(const char* name, int dflt);

template <int BN, int GEGLU>
static int launch_bn(const CUtensorMap* tm, const Params& p, const PhaseMaps& pm, cudaStream_t st) {
  using C = Cfg<BN>;
  static DeviceOnce attr_set;
  if (attr_set.pending()) {
    EALDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, GEGLU, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES));
    attr_set.done();
  }
  const int total = p.m_tiles * p.n_tiles * p.phases;
  const int grid = total < num_sms() ? total : num_sms();
  EALDM_CUDA(launch_pdl(conv_tc_kernel<BN, GEGLU, false>, dim3(grid), dim3(NUM_THREADS), C::SMEM_BYTES, st, tm[0], tm[1],
                        tm[2], tm[3], tm[4], tm[5], p, pm));
  EALDM_LAUNCH_CHECK();
  return 0;
}

// ---- stream-K (Params::sk_tiles) -----------------------------------------------------------------------
// Workspace for the parked accumulators and their flags: a StreamScratch slot (common.cuh) -- per stream, zeroed at
// creation and re-armed by the kernel; a launch that gets none runs the plain schedule (same bits).
constexpr int SK_MAX_CLUSTERS = 80;
constexpr size_t SK_WS_BYTES = static_cast<size_t>(SK_MAX_CLUSTERS) * 2 * (256 / 32) * 8 * BM * sizeof(float4);
constexpr size_t SK_FLAG_BYTES = static_cast<size_t>(SK_MAX_CLUSTERS) * 2 * EPI_WARPS * sizeof(unsigned int);
static StreamScratch g_sk_scratch(SK_WS_BYTES + SK_FLAG_BYTES, 4);   // (19 MB per slot)
static int g_opt_sk = -1;
static bool sk_workspace(int clusters, cudaStream_t st, float4** ws, unsigned int** flags) {
  if (clusters > SK_MAX_CLUSTERS) return false;
  uint8_t* base = static_cast<uint8_t*>(g_sk_scratch.get(st));
  if (!base) return false;
  *ws = reinterpret_cast<float4*>(base);
  *flags = reinterpret_cast<unsigned int*>(base + SK_WS_BYTES);
  return true;
}

// {counted in, read} counters of the GroupNorm epilogue (Params::gna_counters), per stream, zero between launches
constexpr size_t GNA_COUNTER_BYTES = 1 << 18;
static StreamScratch g_gna_counters(GNA_COUNTER_BYTES);

// decide the stream-K region of a CTA-pair launch of `total` super-tiles on `clusters` clusters
static void sk_plan(Params& p, int total, int clusters, cudaStream_t st) {
  if (g_opt_sk < 0) {
    const char* e = getenv("EALDM_TC_STREAMK");
    g_opt_sk = e ? atoi(e) : 0;
  }
  p.sk_tiles = 0;
  if (g_opt_sk == 2 && total > clusters && total % clusters != 0) {   // 2: whenever it applies (tests)
    const int kb2 = p.seg[0].kblocks + (p.nseg > 1 ? p.seg[1].kblocks : 0);
    float4* ws2 = nullptr;
    unsigned int* fl2 = nullptr;
    if (kb2 >= 2 && p.phases == 1 && !p.ln_in && !p.ln_out && !p.ln_gamma && !p.gna_gamma && !p.b_img && sk_workspace(clusters, st, &ws2, &fl2)) {
      p.sk_tiles = clusters + total % clusters;
      p.sk_ws = ws2;
      p.sk_flags = fl2;
    }
    return;
  }
  const int rem = total % clusters;
  const int kblocks = p.seg[0].kblocks + (p.nseg > 1 ? p.seg[1].kblocks : 0);
  // worth it when the last wave leaves a good part of the device idle and a tile is long enough to split
  // (measured at the 8x8 / 16x16 levels' shapes: timed alone, 3x3 convs of 144 k-blocks gain 3-4 % -- less than the
  // 13 % of idle clusters, because a partly idle last wave also runs faster -- and a K = 1024 GEMM of 16 k-blocks LOSES
  // 6 us to parking and resuming, hence the floor of 36 k-blocks = a 3x3 conv.  Inside the power-capped forward the
  // gain is gone: 68.34 / 68.39 samples/s with, 68.45 without, SM clock 1700 against 1728 MHz -- busy SMs in the last
  // wave are paid for in clock.  Hence opt-in: EALDM_TC_STREAMK=1.)
  if (!g_opt_sk || total <= clusters || rem == 0 || rem * 8 > clusters * 7 || kblocks < 36) return;
  if (p.phases != 1 || p.ln_in != nullptr || p.ln_out != nullptr || p.ln_gamma != nullptr || p.gna_gamma != nullptr || p.b_img) return;
  float4* ws = nullptr;
  unsigned int* flags = nullptr;
  if (!sk_workspace(clusters, st, &ws, &flags)) return;
  p.sk_tiles = clusters + rem;
  p.sk_ws = ws;
  p.sk_flags = flags;
}

// CTA pairs: a persistent grid of 2-CTA clusters, as many as the device can hold at once (one CTA per SM)
template <int BN, int GEGLU>
static int launch_pair(const CUtensorMap* tm, const Params& p, const PhaseMaps& pm, cudaStream_t st) {
  using C = Cfg<BN, true>;
  static int max_clusters = 0;   // devices of one process are assumed to be the same model
  static DeviceOnce attr_set;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;   // (the occupancy query below sees the cluster shape only)
  if (attr_set.pending()) {
    EALDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, GEGLU, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::SMEM_BYTES));
    attr_set.done();
  }
  if (!max_clusters) {
    cfg.gridDim = dim3(num_sms() & ~1);
    int n = 0;
    EALDM_CUDA(cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<BN, GEGLU, true>, &cfg));
    EALDM_REQUIRE(n > 0, "tcgen05 conv: no 2-CTA cluster fits on this device");
    max_clusters = n < num_sms() / 2 ? n : num_sms() / 2;
  }
  const int total = ((p.m_tiles + 1) / 2) * p.n_tiles * p.phases;
  const int clusters = total < max_clusters ? total : max_clusters;
  cfg.gridDim = dim3(2 * clusters);
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  Params ps = p;
  if constexpr (BN == 256 && GEGLU == 0) sk_plan(ps, total, clusters, st);
  EALDM_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, GEGLU, true>, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], ps, pm));
  EALDM_LAUNCH_CHECK();
  return 0;
}

// schedule switches (ealdm_tc_set_option); defaults from the environment, read once
static int g_opt[5] = {-1, -1, -1, -1, -1};
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static void init_options() {
  if (g_opt[0] >= 0) return;
  g_opt[EALDM_TC_OPT_CTA2] = env_int("EALDM_TC_CTA2", 1);
  g_opt[EALDM_TC_OPT_WIDE] = env_int("EALDM_TC_WIDE", 1);
  g_opt[EALDM_TC_OPT_RELAXED_WAIT] = env_int("EALDM_TC_RELAXED_WAIT", 1);
  g_opt[EALDM_TC_OPT_BN] = env_int("EALDM_TC_BN", 0);
  g_opt[EALDM_TC_OPT_GELU_ERF] = env_int("EALDM_TC_GELU_ERF", 0);
}
static void sk_init_option() {
  if (g_opt_sk < 0) g_opt_sk = env_int("EALDM_TC_STREAMK", 0);
}
int get_option(int option) {
  init_options();
  sk_init_option();
  if (option == EALDM_TC_OPT_STREAMK) return g_opt_sk;
  return (option >= 0 && option <= 4) ? g_opt[option] : 0;
}
int set_option(int option, int value) {
  init_options();
  sk_init_option();
  if (option == EALDM_TC_OPT_STREAMK) {
    const int prev = g_opt_sk;
    g_opt_sk = value;
    return prev;
  }
  if (option < 0 || option > 4) return set_error(EALDM_EINVAL, "tcgen05 conv: unknown option %d", option);
  const int prev = g_opt[option];
  g_opt[option] = value;
  return prev;
}
static int cta2_mode() {
  init_options();
  return g_opt[EALDM_TC_OPT_CTA2];
}

static bool aligned_2d(const void* ptr, long long ld, int elem_bytes) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * elem_bytes) % 16 == 0;
}

// returns 1 if the tcgen05 path can run this problem
bool supported(const ealdm_conv_args* a) {
  if (a->dtype != EALDM_BF16) return false;
  if (a->n_src < 1 || a->n_src > 2) return false;
  if (a->wi_tokens) {
    // per-image B: one 1x1 source, a 128-row tile inside one image, (head, token) = the 32 logits / probabilities
    const ealdm_conv_src& x = a->src[0];
    if (a->n_src != 1 || x.ksize != 1 || x.stride != 1 || x.pad != 0 || x.upsample || a->upsample_phases) return false;
    const bool two = x.h * x.w == BM / 2;   // two 64-token images per tile (Params::b_img2)
    if (((x.h * x.w) % BM != 0 && !two) || a->h_out != x.h || a->w_out != x.w) return false;
    const long long logits = static_cast<long long>(a->wi_tokens) * a->wi_heads;
    if (two && (logits != 128 || x.w != 8 || x.h != 8)) return false;
    if (a->wi_tokens < 1 || a->wi_heads < 1 || logits % 32 != 0 || logits > 128 || 64 % a->wi_tokens != 0) return false;
    if (!aligned_2d(a->weight, a->wi_ld, 2) || a->wi_head_stride % 8 != 0 || a->rowvec || a->gn_partial) return false;
    if (a->ln_partial_in || a->ln_partial_out) return false;
    if (a->weight_adjoint) {
      if (x.c != (two ? 2 : 1) * logits || a->n_out % 64 != 0 || a->n_out > a->wi_head_stride || a->act != EALDM_ACT_NONE)
        return false;
    } else {
      if (a->n_out != (two ? 2 : 1) * logits || x.c % BK != 0 || x.c > a->wi_head_stride || a->residual || a->out2 ||
          a->out_f32)
        return false;
      if (a->act != EALDM_ACT_SOFTMAX4 && a->act != EALDM_ACT_NONE) return false;
    }
  } else if (a->act == EALDM_ACT_SOFTMAX4) {
    return false;
  }
  for (int s = 0; s < a->n_src; ++s) {
    const ealdm_conv_src& x = a->src[s];
    if ((x.c % BK != 0 && !(a->wi_tokens && a->weight_adjoint)) || !aligned_2d(x.x, x.ld, 2)) return false;
    if (x.upsample && !a->upsample_phases) return false;
    if (x.ksize != 1 && x.ksize != 3) return false;
    if (x.stride != 1 && x.stride != 2) return false;
    if (x.n != a->src[0].n) return false;
  }
  if (a->upsample_phases) {
    // 3x3 / stride 1 / pad 1 over the nearest-2x upsampled source as four 2x2 phases over the source itself
    const ealdm_conv_src& x = a->src[0];
    if (a->n_src != 1 || !x.upsample || x.ksize != 3 || x.stride != 1 || x.pad != 1) return false;
    if (a->weight_adjoint || a->act == EALDM_ACT_GEGLU || a->residual) return false;
    if (a->h_out != 2 * x.h || a->w_out != 2 * x.w || a->k_total != 16 * x.c) return false;
    if (x.h * x.w < 32 || (x.w & (x.w - 1)) != 0 || (x.h & (x.h - 1)) != 0) return false;
  }
  if (a->wi_tokens) {
    // (checked above)
  } else if (a->weight_adjoint) {
    if (a->n_src != 1 || a->act == EALDM_ACT_GEGLU || a->n_out % 64 != 0) return false;
    const long long ldw = a->ld_weight ? a->ld_weight : a->n_out * a->src[0].ksize * a->src[0].ksize;
    if (!aligned_2d(a->weight, ldw, 2)) return false;
  } else if (!aligned_2d(a->weight, a->k_total, 2)) {
    return false;
  }
  if (!aligned_2d(a->out, a->ld_out, a->out_f32 ? 4 : 2)) return false;
  if (a->residual && !aligned_2d(a->residual, a->ld_res, a->res_f32 ? 4 : 2)) return false;
  if (a->out2 && !aligned_2d(a->out2, a->ld_out2, 2)) return false;
  if (a->rowvec && ((reinterpret_cast<uintptr_t>(a->rowvec) & 15) != 0 || a->ld_rowvec % 4 != 0 ||
                    a->n_out % 4 != 0))
    return false;
  if (a->act == EALDM_ACT_GEGLU && (a->rowvec || a->out2 || a->residual || a->n_out % 32 != 0)) return false;
  if (a->ln_partial_out || a->ln_partial_in) {
    // folded LayerNorm: linear geometry (rows = output pixels of ONE image row), narrow epilogue for the producer
    if (a->h_out != 1 || a->src[0].n != 1 || a->weight_adjoint || a->upsample_phases) return false;
    if (a->ln_partial_out && (a->act != EALDM_ACT_NONE || !a->out_f32)) return false;
    if (a->ln_partial_in && (a->out2 || a->ln_parts_in < 1 || a->ln_parts_in > 32 || !a->ln_c1 || a->ln_channels < 1 ||
                             (a->act != EALDM_ACT_NONE && a->act != EALDM_ACT_GEGLU)))
      return false;
  }
  if (a->ln_gamma) {
    // LayerNorm applied by the epilogue: whole rows in one 256-column tile, fp32 result, no shadow copy, narrow epilogue
    if (a->n_out != 256 || !a->ln_beta || !a->out2 || !a->out_f32 || a->act != EALDM_ACT_NONE || a->upsample_phases ||
        a->ln_partial_in || a->ln_partial_out || a->gn_partial || (a->wi_tokens && !a->weight_adjoint))
      return false;
  }
  if (a->gn_gamma) {
    // GroupNorm applied by the epilogue: 256-column tiles, groups of 8 / 16 / 32 channels inside one tile, the 32 rows
    // of an epilogue warp inside one image, partial statistics as channel octets
    if (!a->gn_beta || !a->gn_partial || (a->gn_unit != 0 && a->gn_unit != 8) || a->gn_groups < 1) return false;
    if (a->n_out < 256 || a->n_out % 64 != 0 || a->n_out % a->gn_groups != 0) return false;
    const long long cg = a->n_out / a->gn_groups;
    if (cg != 8 && cg != 16 && cg != 32) return false;
    if (a->act != EALDM_ACT_NONE || a->upsample_phases || a->wi_tokens || a->ln_gamma || a->ln_partial_in ||
        a->ln_partial_out || (a->h_out * a->w_out) % 32 != 0)
      return false;
    if (a->gn_only ? (a->out_f32 || a->out2) : (!a->out2)) return false;
    // the tiles of an image wait for each other while they hold their accumulators: they must be in flight together,
    // i.e. neighbours in one wave of the persistent schedule (or two adjacent waves) -- at most 16 tiles per image
    if (a->h_out * a->w_out > 16 * BM) return false;
  }
  if (a->gn_partial) {
    const long long hw = a->h_out * a->w_out;
    if (a->gn_unit != 0 && a->gn_unit != 8 && a->gn_unit != 4) return false;
    if (a->act == EALDM_ACT_GEGLU || hw % 32 != 0 || (a->w_out & (a->w_out - 1)) != 0 ||
        (a->h_out & (a->h_out - 1)) != 0 || a->n_out % 32 != 0)
      return false;
    if (a->w_out < 32 && (32 % a->w_out != 0 || a->h_out % (32 / a->w_out) != 0)) return false;
    if (a->upsample_phases) {  // the 32-pixel chunks are those of the low-resolution grid
      const long long wl = a->w_out / 2, hl = a->h_out / 2;
      if (wl < 32 && (32 % wl != 0 || hl % (32 / wl) != 0)) return false;
    }
  }
  return true;
}

// 4-D (C, W/2, H/2, N) view of the pixels (2y + py, 2x + px) of an NHWC tensor at the upsampled resolution: doubled
// pixel and row strides, base shifted by (py, px); a [32 columns x sub_w x sub_h x sub_n] unit of the LOW-resolution
// grid is one TMA box
static int encode_phase_map(PFN_cuTensorMapEncodeTiled_v12000 encode, CUtensorMap* tm, const void* base, bool f32,
                            long long cols, long long ld, const ealdm_conv_args* a, const int (&sub)[3], int py, int px) {
  const cuuint64_t es = f32 ? 4 : 2;
  const cuuint64_t pix = static_cast<cuuint64_t>(ld) * es;  // bytes between horizontally adjacent output pixels
  const cuuint64_t wout = static_cast<cuuint64_t>(a->w_out), hout = static_cast<cuuint64_t>(a->h_out);
  const uint8_t* origin = static_cast<const uint8_t*>(base) + (static_cast<cuuint64_t>(py) * wout + px) * pix;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cols), wout / 2, hout / 2, static_cast<cuuint64_t>(a->src[0].n)};
  cuuint64_t gstr[3] = {2 * pix, 2 * pix * wout, pix * wout * hout};
  cuuint32_t box[4] = {32u, static_cast<cuuint32_t>(sub[0]), static_cast<cuuint32_t>(sub[1]),
                       static_cast<cuuint32_t>(sub[2])};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                      const_cast<uint8_t*>(origin), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(upsampling phases) failed: %d", (int)r);
  return 0;
}

// 4-D (C, W, H, N) tensor map over an NHWC output-space tensor with a [32 columns x 32 rows] box
static int encode_unit_map(PFN_cuTensorMapEncodeTiled_v12000 encode, CUtensorMap* tm, const void* base,
                           bool f32, long long cols, long long ld, const ealdm_conv_args* a, const int (&sub)[3],
                           bool wide_bf16 = false) {
  const cuuint64_t es = f32 ? 4 : 2;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(a->w_out),
                        static_cast<cuuint64_t>(a->h_out), static_cast<cuuint64_t>(a->src[0].n)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(ld) * es, static_cast<cuuint64_t>(ld) * es * gdim[1],
                        static_cast<cuuint64_t>(ld) * es * gdim[1] * gdim[2]};
  // box rows are 128 B (32 fp32, or 64 bf16 in the wide epilogue) or 64 B (32 bf16)
  cuuint32_t box[4] = {wide_bf16 ? 64u : 32u, static_cast<cuuint32_t>(sub[0]), static_cast<cuuint32_t>(sub[1]),
                       static_cast<cuuint32_t>(sub[2])};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                      const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      (f32 || wide_bf16) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(epilogue) failed: %d", (int)r);
  return 0;
}

// the N tile: fewest (waves x tile cost)
static int choose_bn(const ealdm_conv_args* a, long long m_work) {
  init_options();
  const bool geglu = a->act == EALDM_ACT_GEGLU;
  if (a->ln_gamma) return 256;   // LayerNorm in the epilogue needs the whole row in one tile
  if (a->gn_gamma) return 256;   // GroupNorm in the epilogue: written for the 256-column tile
  if (a->wi_tokens && !a->weight_adjoint && a->src[0].h * a->src[0].w == BM / 2) return 128;   // one image's logits per N tile
  if (a->n_out <= 32 && !geglu && !a->weight_adjoint) return 32;
  if (a->n_out <= 128) return 128;
  const long long t256 = m_work * ceil_div(a->n_out, 256);
  const long long t128 = m_work * ceil_div(a->n_out, 128);
  const long long c256 = ceil_div(t256, num_sms()) * (256 + 48);
  const long long c128 = ceil_div(t128, num_sms()) * (128 + 48);
  int BN = (c256 <= c128) ? 256 : 128;
  if (g_opt[EALDM_TC_OPT_BN] == 128 || g_opt[EALDM_TC_OPT_BN] == 256) BN = g_opt[EALDM_TC_OPT_BN];
  return BN;
}

// number of {sum, sum of squares} partials per row a launch with ln_partial_out writes: one per 32 output columns
int ln_parts(const ealdm_conv_args* a) { return static_cast<int>(ceil_div(a->n_out, 32)); }

int launch(const ealdm_conv_args* a, cudaStream_t st) {
  EALDM_REQUIRE(supported(a), "tcgen05 conv: unsupported shape/alignment (c%%64, 16-byte rows and pointers)");
  init_options();
  PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode();
  if (!encode) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled entry point not available");

  Params p;
  memset(&p, 0, sizeof(p));
  const bool phased = a->upsample_phases != 0;  // tiles live on the LOW-resolution grid, four output phases each
  const long long w_grid = phased ? a->w_out / 2 : a->w_out, h_grid = phased ? a->h_out / 2 : a->h_out;
  p.nseg = a->n_src;
  p.phases = phased ? 4 : 1;
  p.Wout = static_cast<int>(w_grid);
  p.Hout = static_cast<int>(h_grid);
  p.Nimg = static_cast<int>(a->src[0].n);
  p.N = static_cast<int>(a->n_out);
  p.bw = pow2_ceil(w_grid) < BM ? pow2_ceil(w_grid) : BM;
  p.bh = pow2_ceil(h_grid) < BM / p.bw ? pow2_ceil(h_grid) : BM / p.bw;
  p.bn = BM / (p.bw * p.bh);
  p.tiles_w = static_cast<int>(ceil_div(w_grid, p.bw));
  p.tiles_h = static_cast<int>(ceil_div(h_grid, p.bh));
  p.m_tiles = p.tiles_w * p.tiles_h * static_cast<int>(ceil_div(p.Nimg, p.bn));
  // the 32 rows of one epilogue warp form a (sub_w, sub_h, sub_n) sub-box of the tile
  int sub[3];
  sub[0] = p.bw < 32 ? p.bw : 32;
  sub[1] = p.bh < 32 / sub[0] ? p.bh : 32 / sub[0];
  sub[2] = 32 / (sub[0] * sub[1]);

  const bool geglu = a->act == EALDM_ACT_GEGLU;
  const int BN = choose_bn(a, static_cast<long long>(p.m_tiles) * p.phases);
  p.n_tiles = static_cast<int>(ceil_div(a->n_out, BN));
  // CTA pairs (256 x 256 tiles, a third less operand traffic per SM) for every problem with an even number of M tiles
  // and a reduction long enough (K >= 1024) to amortise the pair's cross-CTA barrier latency (measured: K = 256 / 512
  // GEMMs lose 5-25 % as pairs, K >= 1024 GEMMs and all 3x3 convs gain 3-12 %); EALDM_TC_CTA2=2 pairs regardless of K
  const bool pair = !a->wi_tokens && (BN == 256 || (BN == 128 && !geglu && cta2_mode() >= 2)) && cta2_mode() != 0 && p.m_tiles >= 2 && p.m_tiles % 2 == 0 &&
                    ((phased ? a->k_total / 4 : a->k_total) >= 1024 || cta2_mode() == 2);  // 3: the K rule, 128-column tiles included

  CUtensorMap tm[6];  // A0, A1, W, out, out2, residual
  memset(tm, 0, sizeof(tm));
  int koff = 0;
  for (int s = 0; s < a->n_src; ++s) {
    const ealdm_conv_src& x = a->src[s];
    Segment& sg = p.seg[s];
    sg.cblk = static_cast<int>((x.c + BK - 1) / BK);  // (a 32-channel source of per-image weights: one zero-filled block)
    sg.ksize = phased ? 2 : x.ksize;  // a phase is a 2x2 convolution whose taps start at (py - 1, px - 1)
    sg.kblocks = sg.cblk * sg.ksize * sg.ksize;
    sg.pad = x.pad;
    sg.stride = x.stride;
    sg.bkoff = koff;
    koff += static_cast<int>(x.c) * (phased ? 16 : x.ksize * x.ksize);
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
    CUresult r = encode(&tm[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x.x), gdim,
                        gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(A%d) failed: %d", s, (int)r);
  }
  EALDM_REQUIRE(koff == a->k_total, "k_total %lld does not match the sources (%d)",
                (long long)a->k_total, koff);
  if (a->n_src == 1) tm[1] = tm[0];
  if (a->wi_tokens) {
    // 4-D (c, token, head, image) view of the context projection: box rows come out as (head, token), token fastest;
    // adjoint (MN-major atoms of 64 K rows): heads beyond wi_heads are out of range and arrive as zeros
    const cuuint64_t T = static_cast<cuuint64_t>(a->wi_tokens);
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(a->weight_adjoint ? a->n_out : a->k_total), T,
                          static_cast<cuuint64_t>(a->wi_heads), static_cast<cuuint64_t>(a->src[0].n)};
    cuuint64_t gstr[3] = {static_cast<cuuint64_t>(a->wi_ld) * 2, static_cast<cuuint64_t>(a->wi_head_stride) * 2,
                          static_cast<cuuint64_t>(a->wi_ld) * 2 * T};
    // (N tile of the logits GEMM: 32 or 128 rows; heads beyond wi_heads are out of range -> zero rows, clipped outputs)
    const int bn_logits = a->n_out <= 32 ? 32 : 128;
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(T),
                         static_cast<cuuint32_t>((a->weight_adjoint ? 64 : bn_logits) / a->wi_tokens), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm[2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->weight), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(per-image W) failed: %d", (int)r);
  } else if (a->weight_adjoint) {
    const ealdm_conv_src& x = a->src[0];
    const long long wcols = a->n_out * x.ksize * x.ksize;
    const long long ldw = a->ld_weight ? a->ld_weight : wcols;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(wcols), static_cast<cuuint64_t>(x.c)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ldw) * 2};
    cuuint32_t box[2] = {64, BK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm[2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->weight), gdim, gstr, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(W adjoint) failed: %d", (int)r);
  } else {
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(a->k_total), static_cast<cuuint64_t>(a->n_out)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(a->k_total) * 2};
    cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(pair ? BN / 2 : BN)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm[2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->weight),
                        gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return set_error(EALDM_ECUDA, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
  }
  const long long out_cols = geglu ? a->n_out / 2 : a->n_out;
  // (measured: the GEGLU pass gains 3 % at K = 256 and loses 4 % at K >= 512, where the stores of the narrow units
  // overlap the longer main loop better)
  const bool wide = g_opt[EALDM_TC_OPT_WIDE] != 0 && BN == 256 && !phased && !a->ln_partial_out && !a->ln_gamma && !a->gn_gamma &&
                    (geglu ? a->k_total <= 256 : (!a->residual && !a->out2));
  PhaseMaps pm;
  memset(&pm, 0, sizeof(pm));
  if (phased) {
    if (int e = encode_phase_map(encode, &tm[3], a->out, a->out_f32 != 0, out_cols, a->ld_out, a, sub, 0, 0)) return e;
    for (int ph = 1; ph < 4; ++ph)
      if (int e = encode_phase_map(encode, &pm.out[ph - 1], a->out, a->out_f32 != 0, out_cols, a->ld_out, a, sub,
                                   ph >> 1, ph & 1))
        return e;
  } else if (int e = encode_unit_map(encode, &tm[3], a->out, a->out_f32 != 0, out_cols, a->ld_out, a, sub,
                                     (wide || (a->gn_gamma && a->gn_only)) && !a->out_f32)) {
    return e;
  }
  tm[4] = tm[3];
  tm[5] = tm[3];
  if (a->out2) {
    if (phased) {
      if (int e = encode_phase_map(encode, &tm[4], a->out2, false, out_cols, a->ld_out2, a, sub, 0, 0)) return e;
      for (int ph = 1; ph < 4; ++ph)
        if (int e = encode_phase_map(encode, &pm.out2[ph - 1], a->out2, false, out_cols, a->ld_out2, a, sub, ph >> 1,
                                     ph & 1))
          return e;
    } else if (int e = encode_unit_map(encode, &tm[4], a->out2, false, out_cols, a->ld_out2, a, sub,
                                       a->gn_gamma != nullptr)) {   // (the GroupNorm epilogue stores 64-column boxes)
      return e;
    }
  }
  if (a->residual)
    if (int e = encode_unit_map(encode, &tm[5], a->residual, a->res_f32 != 0, out_cols, a->ld_res, a, sub)) return e;

  p.bias = a->bias;
  p.rowvec = a->rowvec;
  p.ld_rowvec = a->ld_rowvec;
  p.act = a->act;
  p.out_f32 = a->out_f32;
  p.has_res = a->residual != nullptr;
  p.res_f32 = a->res_f32;
  p.has_out2 = a->out2 != nullptr && !a->ln_gamma && !a->gn_gamma;
  p.gna_debug = env_int("EALDM_GNA_DEBUG", 0);
  if (a->gn_gamma) {
    const long long hw = a->h_out * a->w_out, cg = a->n_out / a->gn_groups;
    const long long entries = static_cast<long long>(p.Nimg) * p.n_tiles * 2;
    EALDM_REQUIRE(entries * 4 <= static_cast<long long>(GNA_COUNTER_BYTES), "tcgen05 conv: too many images for the GroupNorm epilogue");
    unsigned int* counters = static_cast<unsigned int*>(g_gna_counters.get(st));
    EALDM_REQUIRE(counters != nullptr, "tcgen05 conv: no counter scratch for the GroupNorm epilogue on this stream (first "
                                       "use inside a stream capture, or too many streams)");
    p.gna_gamma = a->gn_gamma;
    p.gna_beta = a->gn_beta;
    p.gna_eps = a->gn_eps;
    p.gna_silu = a->gn_silu;
    p.gna_only = a->gn_only;
    p.gna_debug = env_int("EALDM_GNA_DEBUG", 0);
    p.gna_octets = static_cast<int>(cg / 8);
    p.gna_inv_count = 1.0 / static_cast<double>(hw * cg);
    p.gna_expected = static_cast<unsigned int>(2 * hw / 32);
    p.gna_counters = counters;
  }
  p.ln_gamma = a->ln_gamma;
  p.ln_beta = a->ln_beta;
  p.ln_apply_eps = a->ln_eps;
  p.gn_partial = reinterpret_cast<float2*>(a->gn_partial);
  p.gn_ld = static_cast<int>(a->gn_ld);
  p.gn_quads = a->gn_unit == 4 ? 1 : 0;
  p.gn_chunks = static_cast<int>(a->h_out * a->w_out / 32);
  p.gn_phase_chunks = static_cast<int>(h_grid * w_grid / 32);
  p.ln_out = reinterpret_cast<float2*>(a->ln_partial_out);
  p.ln_out_parts = static_cast<int>(ceil_div(a->n_out, 32));
  p.ln_in = reinterpret_cast<const float2*>(a->ln_partial_in);
  p.ln_in_parts = static_cast<int>(a->ln_parts_in);
  p.ln_inv_c = a->ln_channels > 0 ? 1.0f / static_cast<float>(a->ln_channels) : 0.f;
  p.ln_eps = a->ln_eps;
  p.ln_c1 = a->ln_c1;
  p.b_mn = a->weight_adjoint ? 1 : 0;
  p.b_img = a->wi_tokens;
  p.b_img2 = (a->wi_tokens && a->src[0].h * a->src[0].w == BM / 2) ? 1 : 0;
  p.wide = wide ? 1 : 0;
  p.relaxed_wait = g_opt[EALDM_TC_OPT_RELAXED_WAIT];

  const bool erf = g_opt[EALDM_TC_OPT_GELU_ERF] != 0;
  switch (BN) {
    case 32: return launch_bn<32, 0>(tm, p, pm, st);
    case 128:
      if (pair) return launch_pair<128, 0>(tm, p, pm, st);
      if (!geglu) return launch_bn<128, 0>(tm, p, pm, st);
      return erf ? launch_bn<128, 2>(tm, p, pm, st) : launch_bn<128, 1>(tm, p, pm, st);
    default:
      if (pair) {
        if (!geglu) return launch_pair<256, 0>(tm, p, pm, st);
        return erf ? launch_pair<256, 2>(tm, p, pm, st) : launch_pair<256, 1>(tm, p, pm, st);
      }
      if (!geglu) return launch_bn<256, 0>(tm, p, pm, st);
      return erf ? launch_bn<256, 2>(tm, p, pm, st) : launch_bn<256, 1>(tm, p, pm, st);
  }
}

}  // namespace tc
}  // namespace ealdm
