// DDIM update, q_sample and the guided-eps MSE as single vectorised elementwise kernels (fp32).
// Every arithmetic step uses the round-to-nearest intrinsics (no FMA contraction) in exactly the
// order the reference's chain of torch ops applies them, so the results are bit-identical to the
// reference given the same inputs.  HBM-bound: 6 tensors x 4 B per latent element for the CFG form.
#include "common.cuh"

namespace ealdm {
namespace sampler {

constexpr int NT = 256;

struct DdimScalars {
  float cfg_scale, sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef, sigma_t, temperature;
};

__device__ __forceinline__ void ddim_one(float x, float eu, float ec, float nz, bool has_u,
                                         bool has_n, const DdimScalars& s, float& x_prev,
                                         float& pred, float& e) {
  // ddim.py:176  e_t = e_t_uncond + scale * (e_t - e_t_uncond)
  e = has_u ? __fadd_rn(eu, __fmul_rn(s.cfg_scale, __fsub_rn(ec, eu))) : ec;
  // ddim.py:195  pred_x0 = (x - sqrt_one_minus_at * e_t) / a_t.sqrt()
  pred = __fdiv_rn(__fsub_rn(x, __fmul_rn(s.sqrt_one_minus_at, e)), s.sqrt_at);
  // ddim.py:199  dir_xt = (1 - a_prev - sigma_t**2).sqrt() * e_t
  const float dir = __fmul_rn(s.dir_coef, e);
  // ddim.py:200  noise = sigma_t * randn * temperature
  const float nt = has_n ? __fmul_rn(__fmul_rn(s.sigma_t, nz), s.temperature) : 0.0f;
  // ddim.py:203  x_prev = a_prev.sqrt() * pred_x0 + dir_xt + noise
  x_prev = __fadd_rn(__fadd_rn(__fmul_rn(s.sqrt_a_prev, pred), dir), nt);
}

__global__ void __launch_bounds__(NT)
ddim_step_kernel(const float* __restrict__ x, const float* __restrict__ eu,
                 const float* __restrict__ ec, const float* __restrict__ noise,
                 float* __restrict__ x_prev, float* __restrict__ pred_x0, float* __restrict__ e_out,
                 long long numel, DdimScalars s) {
  const long long nvec = numel >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * NT;
  for (long long i = static_cast<long long>(blockIdx.x) * NT + threadIdx.x; i < nvec; i += stride) {
    const float4 vx = reinterpret_cast<const float4*>(x)[i];
    const float4 vc = reinterpret_cast<const float4*>(ec)[i];
    float4 vu = make_float4(0.f, 0.f, 0.f, 0.f), vn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (eu) vu = reinterpret_cast<const float4*>(eu)[i];
    if (noise) vn = reinterpret_cast<const float4*>(noise)[i];
    float4 xp, pr, ee;
    ddim_one(vx.x, vu.x, vc.x, vn.x, eu != nullptr, noise != nullptr, s, xp.x, pr.x, ee.x);
    ddim_one(vx.y, vu.y, vc.y, vn.y, eu != nullptr, noise != nullptr, s, xp.y, pr.y, ee.y);
    ddim_one(vx.z, vu.z, vc.z, vn.z, eu != nullptr, noise != nullptr, s, xp.z, pr.z, ee.z);
    ddim_one(vx.w, vu.w, vc.w, vn.w, eu != nullptr, noise != nullptr, s, xp.w, pr.w, ee.w);
    reinterpret_cast<float4*>(x_prev)[i] = xp;
    if (pred_x0) reinterpret_cast<float4*>(pred_x0)[i] = pr;
    if (e_out) reinterpret_cast<float4*>(e_out)[i] = ee;
  }
  // tail (numel % 4)
  for (long long i = (nvec << 2) + static_cast<long long>(blockIdx.x) * NT + threadIdx.x; i < numel;
       i += stride) {
    float xp, pr, ee;
    ddim_one(x[i], eu ? eu[i] : 0.f, ec[i], noise ? noise[i] : 0.f, eu != nullptr, noise != nullptr,
             s, xp, pr, ee);
    x_prev[i] = xp;
    if (pred_x0) pred_x0[i] = pr;
    if (e_out) e_out[i] = ee;
  }
}

__global__ void __launch_bounds__(NT)
q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                const int64_t* __restrict__ t, const float* __restrict__ sa,
                const float* __restrict__ s1a, long long per_sample, long long total,
                float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * NT;
  for (long long i = static_cast<long long>(blockIdx.x) * NT + threadIdx.x; i < total; i += stride) {
    const long long b = i / per_sample;
    const int64_t tb = t[b];
    // ddpm.py:278-279  sqrt_alphas_cumprod[t] * x_start + sqrt_one_minus_alphas_cumprod[t] * noise
    out[i] = __fadd_rn(__fmul_rn(sa[tb], x0[i]), __fmul_rn(s1a[tb], noise[i]));
  }
}

__global__ void __launch_bounds__(NT)
cfg_mse_kernel(const float* __restrict__ eu, const float* __restrict__ ec,
               const float* __restrict__ target, float cfg_scale, long long per_sample,
               float* __restrict__ loss_simple) {
  __shared__ float red[NT / 32];
  const long long b = blockIdx.x;
  const float* pu = eu ? eu + b * per_sample : nullptr;
  const float* pc = ec + b * per_sample;
  const float* pt = target + b * per_sample;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < per_sample; i += NT) {
    const float c = pc[i];
    // ddpm.py:1043  e_t_uncond + s * (e_t - e_t_uncond)
    const float g = pu ? __fadd_rn(pu[i], __fmul_rn(cfg_scale, __fsub_rn(c, pu[i]))) : c;
    const float d = __fsub_rn(pt[i], g);
    acc += __fmul_rn(d, d);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < NT / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss_simple[b] = v / static_cast<float>(per_sample);
  }
}

// DDPM ancestral step, ddpm.py:1081-1140 (eps parameterisation), per sample b with t = t[b], in the reference's
// fp32 operation order:
//   x0   = sqrt_recip_ac[t]*x - sqrt_recipm1_ac[t]*eps;  optional clamp to [-1, 1]
//   mean = coef1[t]*x0 + coef2[t]*x;   x_prev = mean + (t != 0) * exp(0.5*logvar[t]) * (noise*temperature)
__global__ void __launch_bounds__(NT)
ddpm_step_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                 const long long* __restrict__ t, const float* __restrict__ sqrt_recip_ac,
                 const float* __restrict__ sqrt_recipm1_ac, const float* __restrict__ coef1,
                 const float* __restrict__ coef2, const float* __restrict__ logvar, int clip, float temperature,
                 long long per_sample, long long total, float* __restrict__ x_prev, float* __restrict__ x0_out) {
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < total;
       i += static_cast<long long>(NT) * gridDim.x) {
    const long long tb = t[i / per_sample];
    float x0 = __fsub_rn(__fmul_rn(sqrt_recip_ac[tb], x[i]), __fmul_rn(sqrt_recipm1_ac[tb], eps[i]));
    if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    const float mean = __fadd_rn(__fmul_rn(coef1[tb], x0), __fmul_rn(coef2[tb], x[i]));
    const float mask = tb == 0 ? 0.0f : 1.0f;
    const float sd = expf(__fmul_rn(0.5f, logvar[tb]));
    const float nz = __fmul_rn(noise[i], temperature);
    x_prev[i] = __fadd_rn(mean, __fmul_rn(__fmul_rn(mask, sd), nz));
    if (x0_out) x0_out[i] = x0;
  }
}

// PLMS (pseudo linear multistep) eps combination, plms.py:213-231, in the reference's fp32 operation order:
//   e_t = e_uncond + s*(e_cond - e_uncond)   (or e_cond)
//   mode 1: (e_t + o1) / 2   mode 2: (3 e_t - o1) / 2   mode 3: (23 e_t - 16 o1 + 5 o2) / 12
//   mode 4: (55 e_t - 59 o1 + 37 o2 - 9 o3) / 24
__global__ void __launch_bounds__(NT)
plms_eps_kernel(const float* __restrict__ eu, const float* __restrict__ ec, float cfg_scale,
                const float* __restrict__ o1, const float* __restrict__ o2, const float* __restrict__ o3, int mode,
                float* __restrict__ e_t_out, float* __restrict__ e_prime_out, long long numel) {
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < numel;
       i += static_cast<long long>(NT) * gridDim.x) {
    const float c = ec[i];
    const float e = eu ? __fadd_rn(eu[i], __fmul_rn(cfg_scale, __fsub_rn(c, eu[i]))) : c;
    if (e_t_out) e_t_out[i] = e;
    float r = e;
    if (mode == 1) {
      r = __fdiv_rn(__fadd_rn(e, o1[i]), 2.0f);
    } else if (mode == 2) {
      r = __fdiv_rn(__fsub_rn(__fmul_rn(3.0f, e), o1[i]), 2.0f);
    } else if (mode == 3) {
      r = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.0f, e), __fmul_rn(16.0f, o1[i])), __fmul_rn(5.0f, o2[i])), 12.0f);
    } else if (mode == 4) {
      r = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(55.0f, e), __fmul_rn(59.0f, o1[i])),
                                        __fmul_rn(37.0f, o2[i])),
                              __fmul_rn(9.0f, o3[i])),
                    24.0f);
    }
    if (e_prime_out) e_prime_out[i] = r;
  }
}

// gradient of sum_b w[b] * mean_i (guided - target)^2 w.r.t. e_cond / e_uncond
__global__ void __launch_bounds__(NT)
cfg_mse_bwd_kernel(const float* __restrict__ eu, const float* __restrict__ ec, const float* __restrict__ target,
                   const float* __restrict__ w, float cfg_scale, long long per_sample, long long total,
                   float* __restrict__ deu, float* __restrict__ dec) {
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < total;
       i += static_cast<long long>(NT) * gridDim.x) {
    const long long b = i / per_sample;
    const float c = ec[i];
    const float g = eu ? __fadd_rn(eu[i], __fmul_rn(cfg_scale, __fsub_rn(c, eu[i]))) : c;
    const float d = 2.0f * w[b] / static_cast<float>(per_sample) * (g - target[i]);
    if (eu) {
      dec[i] = cfg_scale * d;
      deu[i] = (1.0f - cfg_scale) * d;
    } else {
      dec[i] = d;
    }
  }
}

}  // namespace sampler
}  // namespace ealdm

using namespace ealdm;

extern "C" int ealdm_ddim_step(const ealdm_ddim_step_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->x && a->e_cond && a->x_prev, "ddim_step: null argument");
  EALDM_REQUIRE(a->numel > 0, "ddim_step: numel must be positive");
  EALDM_REQUIRE(a->noise || a->sigma_t == 0.0f, "ddim_step: noise is required when sigma_t != 0");
  const uintptr_t align = reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->e_cond) |
                          reinterpret_cast<uintptr_t>(a->e_uncond) |
                          reinterpret_cast<uintptr_t>(a->noise) |
                          reinterpret_cast<uintptr_t>(a->x_prev) |
                          reinterpret_cast<uintptr_t>(a->pred_x0) |
                          reinterpret_cast<uintptr_t>(a->e_out);
  EALDM_REQUIRE((align & 15) == 0, "ddim_step: tensors must be 16-byte aligned");
  sampler::DdimScalars s{a->cfg_scale, a->sqrt_one_minus_at, a->sqrt_at, a->sqrt_a_prev,
                         a->dir_coef, a->sigma_t, a->temperature};
  long long blocks = ceil_div(ceil_div(a->numel, 4), sampler::NT);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  sampler::ddim_step_kernel<<<static_cast<unsigned>(blocks), sampler::NT, 0,
                              static_cast<cudaStream_t>(stream)>>>(
      a->x, a->e_uncond, a->e_cond, a->noise, a->x_prev, a->pred_x0, a->e_out, a->numel, s);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_q_sample(const float* x0, const float* noise, const int64_t* t,
                              const float* sqrt_alphas_cumprod,
                              const float* sqrt_one_minus_alphas_cumprod, int64_t batch,
                              int64_t per_sample, float* out, ealdm_stream_t stream) {
  EALDM_REQUIRE(x0 && noise && t && sqrt_alphas_cumprod && sqrt_one_minus_alphas_cumprod && out,
                "q_sample: null argument");
  EALDM_REQUIRE(batch > 0 && per_sample > 0, "q_sample: bad sizes");
  const long long total = batch * per_sample;
  long long blocks = ceil_div(total, sampler::NT);
  if (blocks > 148 * 8) blocks = 148 * 8;
  sampler::q_sample_kernel<<<static_cast<unsigned>(blocks), sampler::NT, 0,
                             static_cast<cudaStream_t>(stream)>>>(
      x0, noise, t, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, per_sample, total, out);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_cfg_mse(const float* e_uncond, const float* e_cond, const float* target,
                             float cfg_scale, int64_t batch, int64_t per_sample, float* loss_simple,
                             ealdm_stream_t stream) {
  EALDM_REQUIRE(e_cond && target && loss_simple, "cfg_mse: null argument");
  EALDM_REQUIRE(batch > 0 && per_sample > 0, "cfg_mse: bad sizes");
  sampler::cfg_mse_kernel<<<static_cast<unsigned>(batch), sampler::NT, 0,
                            static_cast<cudaStream_t>(stream)>>>(e_uncond, e_cond, target, cfg_scale,
                                                                 per_sample, loss_simple);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_cfg_mse_bwd(const float* e_uncond, const float* e_cond, const float* target, const float* w,
                                 float cfg_scale, int64_t batch, int64_t per_sample, float* de_uncond,
                                 float* de_cond, ealdm_stream_t stream) {
  EALDM_REQUIRE(e_cond && target && w && de_cond, "cfg_mse_bwd: null argument");
  EALDM_REQUIRE((e_uncond == nullptr) == (de_uncond == nullptr), "cfg_mse_bwd: e_uncond and de_uncond go together");
  EALDM_REQUIRE(batch > 0 && per_sample > 0, "cfg_mse_bwd: bad sizes");
  const long long total = batch * per_sample;
  const long long blocks = ceil_div(total, sampler::NT);
  sampler::cfg_mse_bwd_kernel<<<static_cast<unsigned>(blocks < 2368 ? blocks : 2368), sampler::NT, 0,
                                static_cast<cudaStream_t>(stream)>>>(e_uncond, e_cond, target, w, cfg_scale,
                                                                     per_sample, total, de_uncond, de_cond);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_plms_eps(const float* e_uncond, const float* e_cond, float cfg_scale, const float* old1,
                              const float* old2, const float* old3, int32_t mode, float* e_t_out, float* e_prime_out,
                              int64_t numel, ealdm_stream_t stream) {
  EALDM_REQUIRE(e_cond && numel > 0 && mode >= 0 && mode <= 4, "plms_eps: bad arguments");
  EALDM_REQUIRE((mode < 1 || old1) && (mode < 3 || old2) && (mode < 4 || old3), "plms_eps: missing eps history");
  const long long blocks = ceil_div(numel, sampler::NT);
  sampler::plms_eps_kernel<<<static_cast<unsigned>(blocks < 2368 ? blocks : 2368), sampler::NT, 0,
                             static_cast<cudaStream_t>(stream)>>>(e_uncond, e_cond, cfg_scale, old1, old2, old3, mode,
                                                                  e_t_out, e_prime_out, numel);
  EALDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int ealdm_ddpm_step(const float* x, const float* eps, const float* noise, const int64_t* t,
                               const float* sqrt_recip_alphas_cumprod, const float* sqrt_recipm1_alphas_cumprod,
                               const float* posterior_mean_coef1, const float* posterior_mean_coef2,
                               const float* posterior_log_variance_clipped, int32_t clip_denoised, float temperature,
                               int64_t batch, int64_t per_sample, float* x_prev, float* x0_out, ealdm_stream_t stream) {
  EALDM_REQUIRE(x && eps && noise && t && sqrt_recip_alphas_cumprod && sqrt_recipm1_alphas_cumprod &&
                    posterior_mean_coef1 && posterior_mean_coef2 && posterior_log_variance_clipped && x_prev,
                "ddpm_step: null argument");
  EALDM_REQUIRE(batch > 0 && per_sample > 0, "ddpm_step: bad sizes");
  const long long total = batch * per_sample;
  const long long blocks = ceil_div(total, sampler::NT);
  sampler::ddpm_step_kernel<<<static_cast<unsigned>(blocks < 2368 ? blocks : 2368), sampler::NT, 0,
                              static_cast<cudaStream_t>(stream)>>>(
      x, eps, noise, reinterpret_cast<const long long*>(t), sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod,
      posterior_mean_coef1, posterior_mean_coef2, posterior_log_variance_clipped, clip_denoised, temperature,
      per_sample, total, x_prev, x0_out);
  EALDM_LAUNCH_CHECK();
  return 0;
}
