// Layout conversion at the module boundary (the reference API is fp32 NCHW), nearest-2x upsampling,
// strided copies/casts and the sinusoidal timestep embedding.
#include "common.cuh"

namespace ealdm {
namespace misc {

constexpr int NT = 256;

template <typename T>
__global__ void __launch_bounds__(NT)
nchw_to_nhwc_kernel(const float* __restrict__ x, long long total, int c, int hw, T* __restrict__ y,
                    long long ld_y) {
  // one thread per output element, channel fastest (coalesced writes; reads stride hw, tiny tensors)
  const long long i = static_cast<long long>(blockIdx.x) * NT + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c);
  const long long pix = i / c;  // n*hw + p
  const long long n = pix / hw;
  const int p = static_cast<int>(pix - n * hw);
  y[pix * ld_y + ch] = from_f32<T>(x[(n * c + ch) * hw + p]);
}

template <typename T>
__global__ void __launch_bounds__(NT)
nhwc_to_nchw_kernel(const T* __restrict__ x, long long ld_x, long long total, int c, int hw,
                    float* __restrict__ y) {
  // one thread per output element, pixel fastest (coalesced writes)
  const long long i = static_cast<long long>(blockIdx.x) * NT + threadIdx.x;
  if (i >= total) return;
  const int p = static_cast<int>(i % hw);
  const long long nc = i / hw;
  const int ch = static_cast<int>(nc % c);
  const long long n = nc / c;
  y[i] = to_f32(x[(n * hw + p) * ld_x + ch]);
}

template <typename T>
__global__ void __launch_bounds__(NT)
upsample2x_kernel(const T* __restrict__ x, long long ld_x, int h, int w, int c4, long long total,
                  T* __restrict__ y, long long ld_y) {
  // one thread per 4 output channels
  const long long i = static_cast<long long>(blockIdx.x) * NT + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % c4);
  long long pix = i / c4;
  const int ow = static_cast<int>(pix % (2 * w));
  pix /= (2 * w);
  const int oh = static_cast<int>(pix % (2 * h));
  const long long n = pix / (2 * h);
  Vec4<T> q;
  q.load(x + ((n * h + (oh >> 1)) * w + (ow >> 1)) * ld_x + v * 4);
  q.store(y + ((n * 2 * h + oh) * 2 * w + ow) * ld_y + v * 4);
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(NT)
copy2d_kernel(const TX* __restrict__ x, long long ld_x, TY* __restrict__ y, long long ld_y,
              long long total, int c) {
  const long long i = static_cast<long long>(blockIdx.x) * NT + threadIdx.x;
  if (i >= total) return;
  const int col = static_cast<int>(i % c);
  const long long r = i / c;
  y[r * ld_y + col] = from_f32<TY>(to_f32(x[r * ld_x + col]));
}

template <typename T>
__global__ void __launch_bounds__(NT)
timestep_embedding_kernel(const int64_t* __restrict__ t, long long n, int dim,
                          const float* __restrict__ freqs, T* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * NT + threadIdx.x;
  const int half = dim / 2;
  if (i >= n * half) return;
  const int j = static_cast<int>(i % half);
  const long long r = i / half;
  const float arg = static_cast<float>(t[r]) * freqs[j];
  out[r * dim + j] = from_f32<T>(cosf(arg));
  out[r * dim + half + j] = from_f32<T>(sinf(arg));
  if ((dim & 1) && j == 0) out[r * dim + dim - 1] = from_f32<T>(0.f);
}

}  // namespace misc
}  // namespace ealdm

using namespace ealdm;

static inline unsigned blocks_for(long long total) {
  return static_cast<unsigned>(ceil_div(total, misc::NT));
}

extern "C" int ealdm_nchw_to_nhwc(const float* x, int64_t n, int64_t c, int64_t h, int64_t w,
                                  int32_t dtype, void* y, int64_t ld_y, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && y && n > 0 && c > 0 && h > 0 && w > 0 && ld_y >= c, "nchw_to_nhwc: bad argument");
  const long long total = n * c * h * w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == EALDM_F32)
    misc::nchw_to_nhwc_kernel<float><<<blocks_for(total), misc::NT, 0, st>>>(
        x, total, (int)c, (int)(h * w), reinterpret_cast<float*>(y), ld_y);
  else if (dtype == EALDM_BF16)
    misc::nchw_to_nhwc_kernel<bf16><<<blocks_for(total), misc::NT, 0, st>>>(
        x, total, (int)c, (int)(h * w), reinterpret_cast<bf16*>(y), ld_y);
  else
    return set_error(EALDM_EINVAL, "nchw_to_nhwc: bad dtype %d", dtype);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_nhwc_to_nchw(const void* x, int64_t ld_x, int32_t dtype, int64_t n, int64_t c,
                                  int64_t h, int64_t w, float* y, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && y && n > 0 && c > 0 && h > 0 && w > 0 && ld_x >= c, "nhwc_to_nchw: bad argument");
  const long long total = n * c * h * w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == EALDM_F32)
    misc::nhwc_to_nchw_kernel<float><<<blocks_for(total), misc::NT, 0, st>>>(
        reinterpret_cast<const float*>(x), ld_x, total, (int)c, (int)(h * w), y);
  else if (dtype == EALDM_BF16)
    misc::nhwc_to_nchw_kernel<bf16><<<blocks_for(total), misc::NT, 0, st>>>(
        reinterpret_cast<const bf16*>(x), ld_x, total, (int)c, (int)(h * w), y);
  else
    return set_error(EALDM_EINVAL, "nhwc_to_nchw: bad dtype %d", dtype);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_upsample_nearest2x(const void* x, int64_t ld_x, int32_t dtype, int64_t n,
                                        int64_t h, int64_t w, int64_t c, void* y, int64_t ld_y,
                                        ealdm_stream_t stream) {
  EALDM_REQUIRE(x && y && n > 0 && c > 0 && h > 0 && w > 0, "upsample: bad argument");
  EALDM_REQUIRE(c % 4 == 0 && ld_x % 4 == 0 && ld_y % 4 == 0, "upsample: c, ld_x, ld_y must be multiples of 4");
  const long long total = n * 4 * h * w * (c / 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == EALDM_F32)
    misc::upsample2x_kernel<float><<<blocks_for(total), misc::NT, 0, st>>>(
        reinterpret_cast<const float*>(x), ld_x, (int)h, (int)w, (int)(c / 4), total,
        reinterpret_cast<float*>(y), ld_y);
  else if (dtype == EALDM_BF16)
    misc::upsample2x_kernel<bf16><<<blocks_for(total), misc::NT, 0, st>>>(
        reinterpret_cast<const bf16*>(x), ld_x, (int)h, (int)w, (int)(c / 4), total,
        reinterpret_cast<bf16*>(y), ld_y);
  else
    return set_error(EALDM_EINVAL, "upsample: bad dtype %d", dtype);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_copy2d(const void* x, int64_t ld_x, int32_t dtype_x, void* y, int64_t ld_y,
                            int32_t dtype_y, int64_t rows, int64_t c, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && y && rows >= 0 && c > 0, "copy2d: bad argument");
  if (rows == 0) return 0;
  const long long total = rows * c;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned g = blocks_for(total);
#define EALDM_COPY(TX, TY)                                                                       \
  misc::copy2d_kernel<TX, TY><<<g, misc::NT, 0, st>>>(reinterpret_cast<const TX*>(x), ld_x,      \
                                                      reinterpret_cast<TY*>(y), ld_y, total, (int)c)
  if (dtype_x == EALDM_F32 && dtype_y == EALDM_F32) EALDM_COPY(float, float);
  else if (dtype_x == EALDM_F32 && dtype_y == EALDM_BF16) EALDM_COPY(float, bf16);
  else if (dtype_x == EALDM_BF16 && dtype_y == EALDM_F32) EALDM_COPY(bf16, float);
  else if (dtype_x == EALDM_BF16 && dtype_y == EALDM_BF16) EALDM_COPY(bf16, bf16);
  else return set_error(EALDM_EINVAL, "copy2d: bad dtypes %d %d", dtype_x, dtype_y);
#undef EALDM_COPY
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_timestep_embedding(const int64_t* t, int64_t n, int32_t dim, const float* freqs,
                                        int32_t dtype, void* out, ealdm_stream_t stream) {
  EALDM_REQUIRE(t && freqs && out && n > 0 && dim >= 2, "timestep_embedding: bad argument");
  const long long total = n * (dim / 2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == EALDM_F32)
    misc::timestep_embedding_kernel<float><<<blocks_for(total), misc::NT, 0, st>>>(
        t, n, dim, freqs, reinterpret_cast<float*>(out));
  else if (dtype == EALDM_BF16)
    misc::timestep_embedding_kernel<bf16><<<blocks_for(total), misc::NT, 0, st>>>(
        t, n, dim, freqs, reinterpret_cast<bf16*>(out));
  else
    return set_error(EALDM_EINVAL, "timestep_embedding: bad dtype %d", dtype);
  EALDM_LAUNCH_CHECK();
  return 0;
}
