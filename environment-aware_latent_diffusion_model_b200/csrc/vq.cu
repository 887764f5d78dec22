// Nearest-codebook vector quantisation (taming-transformers VectorQuantizer2.forward, the `quantize` of the
// reference's VQModel / VQModelInterface, ldm/models/autoencoder.py:39-41,274-282):
//   d[p, j] = |z_p|^2 + |e_j|^2 - 2 z_p . e_j,   idx[p] = argmin_j d[p, j] (first minimum),   z_q[p] = e[idx[p]]
// z is NCHW fp32 with e_dim channels (4 for vq-f8), the codebook is [n_e, e_dim] fp32 (16384 x 4).
// One thread per pixel; the codebook streams through shared memory in chunks that every thread of the CTA reads
// at the same address (broadcast), so the kernel is bound by FFMA issue: 6 instructions per (pixel, code).
#include "common.cuh"

namespace ealdm {
namespace vq {

constexpr int NT = 256;
constexpr int CHUNK_FLOATS = 10240;   // shared-memory budget: codes per chunk = CHUNK_FLOATS / (E + 1)

template <int E>
__global__ void __launch_bounds__(NT)
vq_nearest_kernel(const float* __restrict__ z, const float* __restrict__ codebook, int n_e, long long hw,
                  long long total, float* __restrict__ zq, long long* __restrict__ indices) {
  constexpr int CHUNK = CHUNK_FLOATS / (E + 1);
  __shared__ float cb[CHUNK][E];
  __shared__ float ee[CHUNK];
  const long long pix = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
  const bool valid = pix < total;
  const long long img = valid ? pix / hw : 0, p = valid ? pix - img * hw : 0;
  float zv[E];
  float zz = 0.f;
#pragma unroll
  for (int k = 0; k < E; ++k) {
    zv[k] = valid ? z[(img * E + k) * hw + p] : 0.f;
    zz = fmaf(zv[k], zv[k], zz);
  }
  float best = INFINITY;
  int best_j = 0;
  for (int c0 = 0; c0 < n_e; c0 += CHUNK) {
    const int nc = min(CHUNK, n_e - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < nc * E; i += NT) cb[i / E][i % E] = codebook[static_cast<long long>(c0) * E + i];
    __syncthreads();
    for (int i = threadIdx.x; i < nc; i += NT) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < E; ++k) s = fmaf(cb[i][k], cb[i][k], s);
      ee[i] = s;
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < nc; ++j) {
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < E; ++k) dot = fmaf(zv[k], cb[j][k], dot);
      const float d = __fsub_rn(__fadd_rn(zz, ee[j]), __fmul_rn(2.0f, dot));
      if (d < best) { best = d; best_j = c0 + j; }
    }
  }
  if (valid) {
    indices[pix] = best_j;
#pragma unroll
    for (int k = 0; k < E; ++k) zq[(img * E + k) * hw + p] = codebook[static_cast<long long>(best_j) * E + k];
  }
}

}  // namespace vq
}  // namespace ealdm

using namespace ealdm;

extern "C" int ealdm_vq_nearest(const float* z, int64_t n, int64_t e_dim, int64_t hw, const float* codebook,
                                int64_t n_e, float* zq, int64_t* indices, ealdm_stream_t stream) {
  EALDM_REQUIRE(z && codebook && zq && indices && n > 0 && hw > 0 && n_e > 0, "vq_nearest: bad arguments");
  const long long total = n * hw;
  const unsigned blocks = static_cast<unsigned>(ceil_div(total, vq::NT));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* idx = reinterpret_cast<long long*>(indices);
  switch (e_dim) {
    case 3: vq::vq_nearest_kernel<3><<<blocks, vq::NT, 0, st>>>(z, codebook, (int)n_e, hw, total, zq, idx); break;
    case 4: vq::vq_nearest_kernel<4><<<blocks, vq::NT, 0, st>>>(z, codebook, (int)n_e, hw, total, zq, idx); break;
    case 8: vq::vq_nearest_kernel<8><<<blocks, vq::NT, 0, st>>>(z, codebook, (int)n_e, hw, total, zq, idx); break;
    default: return set_error(EALDM_EUNSUPPORTED, "vq_nearest: e_dim %lld unsupported (3, 4 or 8)", (long long)e_dim);
  }
  EALDM_LAUNCH_CHECK();
  return 0;
}
