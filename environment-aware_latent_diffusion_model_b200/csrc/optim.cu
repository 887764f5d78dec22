// Fused optimizer step: AdamW (torch.optim.AdamW, the reference's optimizer, ldm/models/diffusion/ddpm.py:1409-1431)
// + the EMA shadow update (LitEma.forward, ldm/modules/ema.py:25-44) + the bf16 copy of the new weights that the
// next step's GEMMs read, as ONE pass over flat fp32 buffers.  HBM-bound: 38 bytes per parameter
// (read p, g, m, v, ema; write p, m, v, ema, bf16) against > 1200 launches of the reference's per-tensor loops.
#include "common.cuh"

namespace ealdm {
namespace optim {

constexpr int NT = 256;

__global__ void __launch_bounds__(NT)
adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 float* __restrict__ ema, bf16* __restrict__ p16, long long n4, float lr, float beta1, float beta2,
                 float eps, float weight_decay, float bias_c1, float bias_c2_sqrt, float grad_scale,
                 float ema_one_minus_decay) {
  const float step_size = lr / bias_c1;
  const float decay_mul = 1.0f - lr * weight_decay;
  for (long long i = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; i < n4;
       i += static_cast<long long>(NT) * gridDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pp.x, pp.y, pp.z, pp.w};
    const float ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w};
    float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * grad_scale;
      pa[j] *= decay_mul;                                   // decoupled weight decay: p *= 1 - lr*wd
      ma[j] = fmaf(beta1, ma[j], (1.0f - beta1) * gr);      // exp_avg.lerp_(grad, 1 - beta1)
      va[j] = fmaf(beta2, va[j], (1.0f - beta2) * gr * gr); // exp_avg_sq = beta2*v + (1-beta2)*g*g
      const float denom = sqrtf(va[j]) / bias_c2_sqrt + eps;
      pa[j] -= step_size * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (ema != nullptr) {
      float4 ee = reinterpret_cast<float4*>(ema)[i];
      // shadow.sub_(one_minus_decay * (shadow - param))
      ee.x -= ema_one_minus_decay * (ee.x - pa[0]);
      ee.y -= ema_one_minus_decay * (ee.y - pa[1]);
      ee.z -= ema_one_minus_decay * (ee.z - pa[2]);
      ee.w -= ema_one_minus_decay * (ee.w - pa[3]);
      reinterpret_cast<float4*>(ema)[i] = ee;
    }
    if (p16 != nullptr) {
      Vec4<bf16> q;
      q.set(pa);
      q.store(p16 + 4 * i);
    }
  }
}

}  // namespace optim
}  // namespace ealdm

using namespace ealdm;

extern "C" int ealdm_adamw_ema_step(const ealdm_adamw_args* a, ealdm_stream_t stream) {
  EALDM_REQUIRE(a && a->param && a->grad && a->exp_avg && a->exp_avg_sq, "adamw_ema_step: null argument");
  EALDM_REQUIRE(a->numel > 0 && a->numel % 4 == 0, "adamw_ema_step: numel must be a positive multiple of 4");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  EALDM_REQUIRE(al16(a->param) && al16(a->grad) && al16(a->exp_avg) && al16(a->exp_avg_sq) && al16(a->ema) &&
                    (reinterpret_cast<uintptr_t>(a->param_bf16) & 7) == 0,
                "adamw_ema_step: buffers must be 16-byte aligned");
  EALDM_REQUIRE(a->step >= 1, "adamw_ema_step: step counts from 1");
  const double bc1 = 1.0 - pow(static_cast<double>(a->beta1), static_cast<double>(a->step));
  const double bc2 = 1.0 - pow(static_cast<double>(a->beta2), static_cast<double>(a->step));
  const long long n4 = a->numel / 4;
  const long long blocks = ceil_div(n4, optim::NT);
  optim::adamw_ema_kernel<<<static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16), optim::NT, 0,
                            static_cast<cudaStream_t>(stream)>>>(
      a->param, a->grad, a->exp_avg, a->exp_avg_sq, a->ema, reinterpret_cast<bf16*>(a->param_bf16), n4, a->lr, a->beta1,
      a->beta2, a->eps, a->weight_decay, static_cast<float>(bc1), static_cast<float>(sqrt(bc2)), a->grad_scale,
      1.0f - a->ema_decay);
  EALDM_LAUNCH_CHECK();
  return 0;
}
