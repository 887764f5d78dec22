// EALDM conditioner (`UnetCond`, STDiff/models.py:411-539 of the reference) -- the small fp32 kernels around the
// first-stage encoder and the two out_layer GEMMs (which run on the conv / linear kernels):
//   fourier_style   ConditioningTransform + CondScale: fourier features of the time stamp times one bias-free matrix
//   lstm_cell       one torch.nn.LSTM step per frame (gate order i, f, g, o), input projection fused (in <= 64)
//   adain           InstanceNorm2d + x * (1 + gamma) + beta, written into a column window of the concat buffer
//   batch_norm_relu BatchNorm2d (running or batch statistics) + ReLU over a handful of channels
// Everything here is once-per-batch work on [T, 4, 32, 32] maps: one launch each, fixed-order reductions.
#include "common.cuh"

namespace ealdm {
namespace cond {

// out[t, j] = sum_k feat[t, k] * w[j, k] * gain,  feat = [cos0, sin0, cos1, sin1, ...] of 2 pi f_i time[t]
// (include_lin: pair 0 is (1, lin_lr * time[t])); float op order of models.py:221-233: (2 pi as float * f) * t.
__global__ void fourier_style_kernel(const float* __restrict__ time, int T, const float* __restrict__ freqs, int nf,
                                     int include_lin, float lin_lr, const float* __restrict__ w, int n_out, float gain,
                                     float* __restrict__ feat_out, float* __restrict__ out) {
  const int t = blockIdx.x;
  __shared__ float feat[32];
  if (static_cast<int>(threadIdx.x) < nf) {
    const int i = threadIdx.x;
    const float tt = time[t];
    const float arg = (6.283185307179586f * freqs[i]) * tt;
    float c = cosf(arg), s = sinf(arg);
    if (include_lin && i == 0) { c = 1.0f; s = lin_lr * tt; }
    feat[2 * i] = c;
    feat[2 * i + 1] = s;
    if (feat_out != nullptr) { feat_out[t * 2 * nf + 2 * i] = c; feat_out[t * 2 * nf + 2 * i + 1] = s; }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < 2 * nf; ++k) acc = fmaf(feat[k], w[j * 2 * nf + k] * gain, acc);
    out[static_cast<long long>(t) * n_out + j] = acc;
  }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// one LSTM step for B sequences: gates = W_ih x + b_ih + b_hh (+ rec = W_hh h_prev, computed by the linear kernel)
__global__ void lstm_cell_kernel(const float* __restrict__ x, long long ld_x, int n_in, const float* __restrict__ w_ih,
                                 const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                 const float* __restrict__ rec, const float* __restrict__ c_prev, int H,
                                 float* __restrict__ h_out, long long ld_h, float* __restrict__ c_out) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float xs[64];
  if (static_cast<int>(threadIdx.x) < n_in) xs[threadIdx.x] = x[b * ld_x + threadIdx.x];
  __syncthreads();
  if (j >= H) return;
  float g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int row = q * H + j;
    float acc = 0.f;
    for (int k = 0; k < n_in; ++k) acc = fmaf(xs[k], w_ih[static_cast<long long>(row) * n_in + k], acc);
    acc += b_ih[row];
    if (rec != nullptr) acc += rec[static_cast<long long>(b) * 4 * H + row];
    g[q] = acc + b_hh[row];
  }
  const float cp = c_prev != nullptr ? c_prev[static_cast<long long>(b) * H + j] : 0.f;
  const float c = sigmoid_f(g[1]) * cp + sigmoid_f(g[0]) * tanhf(g[2]);
  c_out[static_cast<long long>(b) * H + j] = c;
  h_out[b * ld_h + j] = sigmoid_f(g[3]) * tanhf(c);
}

// AdaIN over an NHWC map [B, hw, c] (c <= 32): one CTA per image, biased variance, fp64 fold of per-thread partials
// in a fixed order; out = (x - mean) * rstd * (1 + gamma) + beta with style = [gamma(c) | beta(c)] per image
__global__ void __launch_bounds__(256)
adain_kernel(const float* __restrict__ x, long long ld_x, int hw, int c, const float* __restrict__ style,
             long long ld_style, float eps, float* __restrict__ y, long long ld_y) {
  const int b = blockIdx.x;
  const float* xb = x + static_cast<long long>(b) * hw * ld_x;
  float* yb = y + static_cast<long long>(b) * hw * ld_y;
  __shared__ double red[256][2];
  __shared__ float s_scale[32], s_shift[32];
  for (int ch = 0; ch < c; ++ch) {
    double s = 0.0, ss = 0.0;
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
      const double v = static_cast<double>(xb[p * ld_x + ch]);
      s += v;
      ss += v * v;
    }
    red[threadIdx.x][0] = s;
    red[threadIdx.x][1] = ss;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (static_cast<int>(threadIdx.x) < o) {
        red[threadIdx.x][0] += red[threadIdx.x + o][0];
        red[threadIdx.x][1] += red[threadIdx.x + o][1];
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const double mean = red[0][0] / hw;
      double var = red[0][1] / hw - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      const float g1 = 1.0f + style[b * ld_style + ch];
      s_scale[ch] = rstd * g1;
      s_shift[ch] = style[b * ld_style + c + ch] - static_cast<float>(mean) * rstd * g1;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < hw * c; i += blockDim.x) {
    const int p = i / c, ch = i - p * c;
    yb[p * ld_y + ch] = fmaf(xb[p * ld_x + ch], s_scale[ch], s_shift[ch]);
  }
}

// BatchNorm2d (+ ReLU) over [rows, c] (c <= 32) in ONE CTA: training -> statistics of the batch (biased variance for
// the normalisation, both moments returned so that the host can update the running buffers), else running statistics
__global__ void __launch_bounds__(1024)
batch_norm_relu_kernel(const float* __restrict__ x, long long ld_x, long long rows, int c,
                       const float* __restrict__ gamma, const float* __restrict__ beta,
                       const float* __restrict__ running_mean, const float* __restrict__ running_var, int training,
                       float eps, int relu, float* __restrict__ y, long long ld_y, float* __restrict__ batch_stats) {
  __shared__ double red[1024][2];
  __shared__ float s_scale[32], s_shift[32];
  for (int ch = 0; ch < c; ++ch) {
    float mean, var;
    if (training) {
      double s = 0.0, ss = 0.0;
      for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
        const double v = static_cast<double>(x[r * ld_x + ch]);
        s += v;
        ss += v * v;
      }
      red[threadIdx.x][0] = s;
      red[threadIdx.x][1] = ss;
      __syncthreads();
      for (int o = 512; o > 0; o >>= 1) {
        if (static_cast<int>(threadIdx.x) < o) {
          red[threadIdx.x][0] += red[threadIdx.x + o][0];
          red[threadIdx.x][1] += red[threadIdx.x + o][1];
        }
        __syncthreads();
      }
      const double m = red[0][0] / rows;
      double v = red[0][1] / rows - m * m;
      if (v < 0.0) v = 0.0;
      mean = static_cast<float>(m);
      var = static_cast<float>(v);
      if (threadIdx.x == 0 && batch_stats != nullptr) { batch_stats[ch] = mean; batch_stats[c + ch] = var; }
      __syncthreads();
    } else {
      mean = running_mean[ch];
      var = running_var[ch];
    }
    if (threadIdx.x == 0) {
      const float sc = gamma[ch] * rsqrtf(var + eps);
      s_scale[ch] = sc;
      s_shift[ch] = beta[ch] - mean * sc;
    }
  }
  __syncthreads();
  for (long long i = threadIdx.x; i < rows * c; i += blockDim.x) {
    const long long r = i / c;
    const int ch = static_cast<int>(i - r * c);
    float v = fmaf(x[r * ld_x + ch], s_scale[ch], s_shift[ch]);
    if (relu) v = fmaxf(v, 0.f);
    y[r * ld_y + ch] = v;
  }
}

// ---- adjoints (the reference trains the conditioner with the UNet, ldm/models/diffusion/ddpm.py:1409-1415) -------------
// block-wide sum of NV doubles per thread, fixed order (tree over threadIdx); result valid in every thread
template <int NT, int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double (*red)[NV]) {
#pragma unroll
  for (int k = 0; k < NV; ++k) red[threadIdx.x][k] = v[k];
  __syncthreads();
  for (int o = NT / 2; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) {
#pragma unroll
      for (int k = 0; k < NV; ++k) red[threadIdx.x][k] += red[threadIdx.x + o][k];
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = red[0][k];
  __syncthreads();
}

// AdaIN adjoint with respect to the STYLE only (its input is the frozen encoder's output): out = xn (1 + gamma) + beta
//   dgamma[b, ch] = sum_p dy * xn = rstd (sum dy x - mean sum dy),   dbeta[b, ch] = sum_p dy
__global__ void __launch_bounds__(256)
adain_bwd_kernel(const float* __restrict__ x, long long ld_x, int hw, int c, const float* __restrict__ dy,
                 long long ld_dy, float eps, float* __restrict__ dstyle, long long ld_ds) {
  const int b = blockIdx.x;
  const float* xb = x + static_cast<long long>(b) * hw * ld_x;
  const float* db = dy + static_cast<long long>(b) * hw * ld_dy;
  __shared__ double red[256][4];
  for (int ch = 0; ch < c; ++ch) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};   // sum x, sum x^2, sum dy, sum dy x
    for (int p = threadIdx.x; p < hw; p += blockDim.x) {
      const double xv = static_cast<double>(xb[p * ld_x + ch]);
      const double dv = static_cast<double>(db[p * ld_dy + ch]);
      v[0] += xv; v[1] += xv * xv; v[2] += dv; v[3] += dv * xv;
    }
    block_sum<256, 4>(v, red);
    if (threadIdx.x == 0) {
      const double mean = v[0] / hw;
      double var = v[1] / hw - mean * mean;
      if (var < 0.0) var = 0.0;
      const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
      dstyle[b * ld_ds + ch] = static_cast<float>(rstd * (v[3] - mean * v[2]));
      dstyle[b * ld_ds + c + ch] = static_cast<float>(v[2]);
    }
  }
}

// BatchNorm2d + ReLU adjoint over [rows, c] in ONE CTA.  y is the forward OUTPUT (post ReLU): g = dy (y > 0).
//   dbeta = sum g, dgamma = sum g xh;  eval: dx = g gamma rstd;  training: dx = gamma rstd (g - mean(g) - xh mean(g xh))
__global__ void __launch_bounds__(1024)
batch_norm_relu_bwd_kernel(const float* __restrict__ x, long long ld_x, long long rows, int c,
                           const float* __restrict__ gamma, const float* __restrict__ running_mean,
                           const float* __restrict__ running_var, int training, float eps, int relu,
                           const float* __restrict__ y, long long ld_y, const float* __restrict__ dy, long long ld_dy,
                           float* __restrict__ dx, long long ld_dx, float* __restrict__ dgamma,
                           float* __restrict__ dbeta) {
  __shared__ double red[1024][4];
  __shared__ float s_mean[32], s_rstd[32], s_mg[32], s_mgx[32];
  for (int ch = 0; ch < c; ++ch) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};   // sum x, sum x^2 (training), then sum g, sum g xh
    if (training) {
      for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
        const double xv = static_cast<double>(x[r * ld_x + ch]);
        v[0] += xv; v[1] += xv * xv;
      }
      block_sum<1024, 4>(v, red);
    }
    float mean, var;
    if (training) {
      const double m = v[0] / rows;
      double vv = v[1] / rows - m * m;
      if (vv < 0.0) vv = 0.0;
      mean = static_cast<float>(m);
      var = static_cast<float>(vv);
    } else {
      mean = running_mean[ch];
      var = running_var[ch];
    }
    const float rstd = rsqrtf(var + eps);
    double w[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
      float g = dy[r * ld_dy + ch];
      if (relu && !(y[r * ld_y + ch] > 0.f)) g = 0.f;
      const float xh = (x[r * ld_x + ch] - mean) * rstd;
      w[0] += static_cast<double>(g);
      w[1] += static_cast<double>(g) * static_cast<double>(xh);
    }
    block_sum<1024, 4>(w, red);
    if (threadIdx.x == 0) {
      s_mean[ch] = mean;
      s_rstd[ch] = rstd;
      s_mg[ch] = training ? static_cast<float>(w[0] / rows) : 0.f;
      s_mgx[ch] = training ? static_cast<float>(w[1] / rows) : 0.f;
      if (dbeta != nullptr) dbeta[ch] += static_cast<float>(w[0]);
      if (dgamma != nullptr) dgamma[ch] += static_cast<float>(w[1]);
    }
  }
  __syncthreads();
  for (long long i = threadIdx.x; i < rows * c; i += blockDim.x) {
    const long long r = i / c;
    const int ch = static_cast<int>(i - r * c);
    float g = dy[r * ld_dy + ch];
    if (relu && !(y[r * ld_y + ch] > 0.f)) g = 0.f;
    const float xh = (x[r * ld_x + ch] - s_mean[ch]) * s_rstd[ch];
    dx[r * ld_dx + ch] = gamma[ch] * s_rstd[ch] * (g - s_mg[ch] - xh * s_mgx[ch]);
  }
}

// one LSTM step backwards: the gates are recomputed exactly as in lstm_cell_kernel, then
//   dc = dh o (1 - tanh(c)^2) + dc_next;  di = dc g i (1 - i);  df = dc c_prev f (1 - f);  dg = dc i (1 - g^2);
//   do = dh tanh(c) o (1 - o);  dgates = [di, df, dg, do] (pre-activation);  dc_prev = dc f
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ x, long long ld_x, int n_in,
                                     const float* __restrict__ w_ih, const float* __restrict__ b_ih,
                                     const float* __restrict__ b_hh, const float* __restrict__ rec,
                                     const float* __restrict__ c_prev, int H, const float* __restrict__ dh,
                                     long long ld_dh, const float* __restrict__ dc_next, float* __restrict__ dgates,
                                     float* __restrict__ dc_prev) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float xs[64];
  if (static_cast<int>(threadIdx.x) < n_in) xs[threadIdx.x] = x[b * ld_x + threadIdx.x];
  __syncthreads();
  if (j >= H) return;
  float g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int row = q * H + j;
    float acc = 0.f;
    for (int k = 0; k < n_in; ++k) acc = fmaf(xs[k], w_ih[static_cast<long long>(row) * n_in + k], acc);
    acc += b_ih[row];
    if (rec != nullptr) acc += rec[static_cast<long long>(b) * 4 * H + row];
    g[q] = acc + b_hh[row];
  }
  const float ig = sigmoid_f(g[0]), fg = sigmoid_f(g[1]), gg = tanhf(g[2]), og = sigmoid_f(g[3]);
  const float cp = c_prev != nullptr ? c_prev[static_cast<long long>(b) * H + j] : 0.f;
  const float c = fg * cp + ig * gg;
  const float tc = tanhf(c);
  const float dhv = dh[b * ld_dh + j];
  float dc = dhv * og * (1.f - tc * tc);
  if (dc_next != nullptr) dc += dc_next[static_cast<long long>(b) * H + j];
  float* dgb = dgates + static_cast<long long>(b) * 4 * H;
  dgb[j] = dc * gg * ig * (1.f - ig);
  dgb[H + j] = dc * cp * fg * (1.f - fg);
  dgb[2 * H + j] = dc * ig * (1.f - gg * gg);
  dgb[3 * H + j] = dhv * tc * og * (1.f - og);
  if (dc_prev != nullptr) dc_prev[static_cast<long long>(b) * H + j] = dc * fg;
}

// dx = dy (y > 0) [* mask]: the adjoint of ReLU (+ an inverted-dropout mask, already scaled by 1 / (1 - p))
__global__ void relu_bwd_kernel(const float* __restrict__ y, long long ld_y, const float* __restrict__ dy,
                                long long ld_dy, const float* __restrict__ mask, long long ld_mask, long long rows,
                                int c, float* __restrict__ dx, long long ld_dx) {
  const long long n = rows * c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c;
    const int ch = static_cast<int>(i - r * c);
    float g = y[r * ld_y + ch] > 0.f ? dy[r * ld_dy + ch] : 0.f;
    if (mask != nullptr) g *= mask[r * ld_mask + ch];
    dx[r * ld_dx + ch] = g;
  }
}

}  // namespace cond
}  // namespace ealdm

using namespace ealdm;

extern "C" int ealdm_fourier_style(const float* time, int64_t t_count, const float* freqs, int32_t n_freq,
                                   int32_t include_lin, float lin_lr, const float* weight, int64_t n_out, float gain,
                                   float* features, float* out, ealdm_stream_t stream) {
  EALDM_REQUIRE(time && freqs && weight && out, "fourier_style: null pointer");
  EALDM_REQUIRE(t_count > 0 && n_freq > 0 && n_freq <= 16 && n_out > 0, "fourier_style: bad sizes");
  cond::fourier_style_kernel<<<static_cast<unsigned>(t_count), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      time, static_cast<int>(t_count), freqs, n_freq, include_lin, lin_lr, weight, static_cast<int>(n_out), gain,
      features, out);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_lstm_cell(const float* x, int64_t ld_x, int64_t batch, int64_t n_in, const float* w_ih,
                               const float* b_ih, const float* b_hh, const float* rec, const float* c_prev,
                               int64_t hidden, float* h_out, int64_t ld_h, float* c_out, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && w_ih && b_ih && b_hh && h_out && c_out, "lstm_cell: null pointer");
  EALDM_REQUIRE(batch > 0 && n_in > 0 && n_in <= 64 && hidden > 0, "lstm_cell: input size must be 1..64");
  dim3 grid(static_cast<unsigned>(ceil_div(hidden, 128)), static_cast<unsigned>(batch));
  cond::lstm_cell_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ld_x, static_cast<int>(n_in), w_ih, b_ih, b_hh, rec, c_prev, static_cast<int>(hidden), h_out, ld_h, c_out);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_adain(const float* x, int64_t ld_x, int64_t n, int64_t hw, int64_t c, const float* style,
                           int64_t ld_style, float eps, float* y, int64_t ld_y, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && style && y, "adain: null pointer");
  EALDM_REQUIRE(n > 0 && hw > 0 && c > 0 && c <= 32, "adain: 1..32 channels");
  cond::adain_kernel<<<static_cast<unsigned>(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ld_x, static_cast<int>(hw), static_cast<int>(c), style, ld_style, eps, y, ld_y);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_batch_norm_relu(const float* x, int64_t ld_x, int64_t rows, int64_t c, const float* gamma,
                                     const float* beta, const float* running_mean, const float* running_var,
                                     int32_t training, float eps, int32_t relu, float* y, int64_t ld_y,
                                     float* batch_stats, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && gamma && beta && y, "batch_norm_relu: null pointer");
  EALDM_REQUIRE(training || (running_mean && running_var), "batch_norm_relu: eval mode needs the running statistics");
  EALDM_REQUIRE(rows > 0 && c > 0 && c <= 32, "batch_norm_relu: 1..32 channels");
  cond::batch_norm_relu_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ld_x, rows, static_cast<int>(c), gamma, beta, running_mean, running_var, training, eps, relu, y, ld_y,
      batch_stats);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_adain_bwd(const float* x, int64_t ld_x, int64_t n, int64_t hw, int64_t c, const float* dy,
                               int64_t ld_dy, float eps, float* dstyle, int64_t ld_dstyle, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && dy && dstyle, "adain_bwd: null pointer");
  EALDM_REQUIRE(n > 0 && hw > 0 && c > 0 && c <= 32, "adain_bwd: 1..32 channels");
  cond::adain_bwd_kernel<<<static_cast<unsigned>(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ld_x, static_cast<int>(hw), static_cast<int>(c), dy, ld_dy, eps, dstyle, ld_dstyle);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_batch_norm_relu_bwd(const float* x, int64_t ld_x, int64_t rows, int64_t c, const float* gamma,
                                         const float* running_mean, const float* running_var, int32_t training,
                                         float eps, int32_t relu, const float* y, int64_t ld_y, const float* dy,
                                         int64_t ld_dy, float* dx, int64_t ld_dx, float* dgamma, float* dbeta,
                                         ealdm_stream_t stream) {
  EALDM_REQUIRE(x && gamma && y && dy && dx, "batch_norm_relu_bwd: null pointer");
  EALDM_REQUIRE(training || (running_mean && running_var), "batch_norm_relu_bwd: eval mode needs the running statistics");
  EALDM_REQUIRE(rows > 0 && c > 0 && c <= 32, "batch_norm_relu_bwd: 1..32 channels");
  cond::batch_norm_relu_bwd_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ld_x, rows, static_cast<int>(c), gamma, running_mean, running_var, training, eps, relu, y, ld_y, dy, ld_dy, dx,
      ld_dx, dgamma, dbeta);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_lstm_cell_bwd(const float* x, int64_t ld_x, int64_t batch, int64_t n_in, const float* w_ih,
                                   const float* b_ih, const float* b_hh, const float* rec, const float* c_prev,
                                   int64_t hidden, const float* dh, int64_t ld_dh, const float* dc_next, float* dgates,
                                   float* dc_prev, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && w_ih && b_ih && b_hh && dh && dgates, "lstm_cell_bwd: null pointer");
  EALDM_REQUIRE(batch > 0 && n_in > 0 && n_in <= 64 && hidden > 0, "lstm_cell_bwd: input size must be 1..64");
  dim3 grid(static_cast<unsigned>(ceil_div(hidden, 128)), static_cast<unsigned>(batch));
  cond::lstm_cell_bwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ld_x, static_cast<int>(n_in), w_ih, b_ih, b_hh, rec, c_prev, static_cast<int>(hidden), dh, ld_dh, dc_next,
      dgates, dc_prev);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_relu_bwd(const float* y, int64_t ld_y, const float* dy, int64_t ld_dy, const float* mask,
                              int64_t ld_mask, int64_t rows, int64_t c, float* dx, int64_t ld_dx,
                              ealdm_stream_t stream) {
  EALDM_REQUIRE(y && dy && dx, "relu_bwd: null pointer");
  EALDM_REQUIRE(rows > 0 && c > 0, "relu_bwd: bad sizes");
  const long long n = rows * c;
  const unsigned grid = static_cast<unsigned>(ceil_div(n, 256) < 4096 ? ceil_div(n, 256) : 4096);
  cond::relu_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, ld_y, dy, ld_dy, mask, ld_mask, rows,
                                                                            static_cast<int>(c), dx, ld_dx);
  EALDM_LAUNCH_CHECK();
  return 0;
}
