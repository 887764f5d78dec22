// Shared host/device helpers for libealdm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>

#include "../../include/ealdm_b200.h"

#include <atomic>
#include <mutex>

namespace ealdm {

// ---- error plumbing (thread-local message, integer codes across the ABI) -----------------------
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define EALDM_REQUIRE(cond, ...)                                        \
  do {                                                                  \
    if (!(cond)) return ::ealdm::set_error(EALDM_EINVAL, __VA_ARGS__);  \
  } while (0)

#define EALDM_CUDA(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::ealdm::set_error(EALDM_ECUDA, "%s failed: %s (%s:%d)", #expr,              \
                                cudaGetErrorString(_e), __FILE__, __LINE__);              \
  } while (0)

#define EALDM_LAUNCH_CHECK()                                                              \
  do {                                                                                    \
    cudaError_t _e = cudaPeekAtLastError();                                               \
    if (_e != cudaSuccess)                                                                \
      return ::ealdm::set_error(EALDM_ECUDA, "kernel launch failed: %s (%s:%d)",          \
                                cudaGetErrorString(_e), __FILE__, __LINE__);              \
    ::ealdm::count_launch();                                                              \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// One-time per-DEVICE work (cudaFuncSetAttribute is per device; one process may drive several GPUs, from several
// threads): a bit per device ordinal, set after the work succeeded.  Doing the work twice in a race is harmless.
struct DeviceOnce {
  std::atomic<unsigned long long> bits{0};
  static int device() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 63;
  }
  bool pending() const { return ((bits.load(std::memory_order_acquire) >> device()) & 1ull) == 0; }
  void done() { bits.fetch_or(1ull << device(), std::memory_order_release); }
};

// ---- library-owned scratch that must start out zeroed (ticket counters, stream-K parking slots) ----------------------
// One allocation per device, made on the first request that arrives OUTSIDE a stream capture and cut into equal
// slots; a stream gets the next free slot the first time it asks (pure bookkeeping, so it also works while the stream
// is being captured -- a CUDA graph's kernels keep the slot of their capture stream).  Kernels that use a slot leave it
// zeroed again.  Launches on different streams never share a slot; two graphs captured on the SAME stream do, and must
// not be replayed concurrently.  nullptr (caller falls back to a path without scratch) if the pool cannot be created
// now (first request inside a capture) or more streams asked than there are slots (16 unless stated).
class StreamScratch {
 public:
  explicit StreamScratch(size_t slot_bytes, int slots = MAX_SLOTS)
      : slot_bytes_((slot_bytes + 255) & ~static_cast<size_t>(255)), slots_(slots < MAX_SLOTS ? slots : MAX_SLOTS) {}
  void* get(cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    std::lock_guard<std::mutex> lock(mu_);
    Pool& p = pool_[dev];
    if (!p.base) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return nullptr;
      }
      void* b = nullptr;
      if (cudaMalloc(&b, slot_bytes_ * slots_) != cudaSuccess || cudaMemset(b, 0, slot_bytes_ * slots_) != cudaSuccess) {
        cudaGetLastError();
        if (b) cudaFree(b);
        return nullptr;
      }
      p.base = static_cast<uint8_t*>(b);
    }
    for (int i = 0; i < p.used; ++i)
      if (p.owner[i] == st) return p.base + i * slot_bytes_;
    if (p.used == slots_) return nullptr;
    p.owner[p.used] = st;
    return p.base + (p.used++) * slot_bytes_;
  }

 private:
  static constexpr int MAX_SLOTS = 16;
  struct Pool {
    uint8_t* base = nullptr;
    int used = 0;
    cudaStream_t owner[MAX_SLOTS];
  };
  size_t slot_bytes_;
  int slots_;
  std::mutex mu_;
  Pool pool_[64];
};

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// The forward pass is a chain of ~260 dependent kernels.  A kernel launched with the programmatic-stream-serialization
// attribute may become resident while its predecessor still runs: its prologue (barrier init, TMEM allocation, tensor
// map prefetch, index arithmetic) then overlaps the predecessor's tail, and `pdl_wait()` (griddepcontrol.wait) holds it
// until the predecessor has completed and its writes are visible.  EVERY kernel launched through `launch_pdl` must call
// `pdl_wait()` before its first global-memory access; `pdl_trigger()` lets the NEXT kernel in the stream do the same.
// Both instructions are no-ops in a kernel launched without the attribute.  Opt-in (EALDM_PDL=1 / ealdm_set_pdl): on the
// CUDA-graph-replayed forward it measured neutral (63.3 vs 64.0 samples/s on one box) -- the large CTAs of consecutive
// kernels cannot share an SM, so only launch latency is left to hide and graph replay already hides it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- element type helpers ----------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
// exact-erf GELU (F.gelu default), attention.py:44
__device__ __forceinline__ float gelu_erf_f(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}

// 4-element vectors of T (16 B for float, 8 B for bf16)
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void get(float (&f)[4]) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
  __device__ __forceinline__ void set(const float (&f)[4]) { v = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec4<bf16> {
  uint2 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint2*>(p) = v; }
  __device__ __forceinline__ void get(float (&f)[4]) const {
    __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  }
  __device__ __forceinline__ void set(const float (&f)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
    v.x = *reinterpret_cast<uint32_t*>(&a);
    v.y = *reinterpret_cast<uint32_t*>(&b);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- epilogue shared by the SIMT and tcgen05 implicit-GEMM kernels ------------------------------
struct Epilogue {
  const float* bias;
  const float* rowvec;
  long long ld_rowvec;
  long long rows_per_image;
  const void* residual;
  long long ld_res;
  void* out;
  long long ld_out;
  int act;
  int out_f32;
  int res_f32;
  void* out2;        // optional bf16/T shadow copy of the output
  long long ld_out2;
};

}  // namespace ealdm
