// Shared device helpers for the samk kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define SAMK_OK 0
#define SAMK_ERR_ARG -1
#define SAMK_ERR_CUDA -2
#define SAMK_ERR_UNSUPPORTED -3

namespace samk {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with the stream-serialization attribute may start (block
// scheduling, barrier / TMEM set-up) while its predecessor in the stream drains; pdl_wait() returns once the
// predecessor has completed and its writes are visible, so it sits before the first global-memory access.
// pdl_release() lets the successor start its own preamble; it is placed after pdl_wait(), so at most one kernel
// runs ahead.  Both are no-ops for a kernel launched without the attribute.  SAMK_PDL: 0 off, 1 GEMMs, 2 + LayerNorm.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static inline int pdl_level() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAMK_PDL");
    v = e ? atoi(e) : 0;
    if (v < 0) v = 0;
  }
  return v;
}
template <class... KArgs, class... Args>
static inline cudaError_t launch_maybe_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                           bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-7 counter RNG (7 rounds: the fewest that pass BigCrush, Salmon et al. SC'11): the
// dropout keep-mask of element e is a pure function of (seed, offset, e), so forward and backward
// kernels regenerate identical masks.  One call yields 128 bits = 16 bits for each of 8 consecutive
// elements (group g = e >> 3); element kept iff its 16 bits >= thresh16 = round(p * 65536), i.e. the
// drop probability is p quantised to 2^-16 and the keep-scale is computed from the quantised value.
// ---------------------------------------------------------------------------------------------
constexpr int kPhiloxRounds = 7;
// Per-replay salt of every dropout stream, XORed into the Philox key.  A CUDA-graph replay re-issues the kernels with
// the (seed, offset) arguments they were captured with; samk_set_dropout_salt() (one tiny copy before each replay)
// is what makes every replay draw new masks.  Forward and backward of one step see the same salt.  One copy of the
// variable per translation unit (no relocatable device code), all set together by the entry point in api.cu.
static __constant__ unsigned long long g_drop_salt = 0ull;
static inline int set_drop_salt_tu(unsigned long long salt, cudaStream_t stream) {
  return cudaMemcpyToSymbolAsync(g_drop_salt, &salt, sizeof(salt), 0, cudaMemcpyHostToDevice, stream) == cudaSuccess ? 0 : -2;
}
// device-to-device form: capturable in a CUDA graph, and it does not queue behind bulk host->device input copies
static inline int set_drop_salt_from_device_tu(const unsigned long long* src, cudaStream_t stream) {
  return cudaMemcpyToSymbolAsync(g_drop_salt, src, sizeof(unsigned long long), 0, cudaMemcpyDeviceToDevice, stream) == cudaSuccess ? 0 : -2;
}

__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr_lo, uint32_t ctr_hi) {
  seed ^= g_drop_salt;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = ctr_hi, c3 = 0x5a17c0deu;
#pragma unroll
  for (int r = 0; r < kPhiloxRounds; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// thresh16 = round(p * 2^16) in [0, 65535]; 0 = no dropout
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = (double)p * 65536.0 + 0.5;
  if (t < 0) t = 0;
  if (t > 65535.0) t = 65535.0;
  return (uint32_t)t;
}
// 1 / P(keep) for the quantised probability
__host__ __device__ __forceinline__ float drop_keep_scale(float p) {
  const uint32_t t = drop_threshold(p);
  return t ? (float)(65536.0 / (65536.0 - (double)t)) : 1.0f;
}

// keep bits of the 8 elements of group g (bit k = element 8g+k kept)
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t offset, uint64_t g, uint32_t thresh16) {
  const uint4 r = philox4x32(seed, g, (uint32_t)offset);
  const uint32_t th = thresh16 << 16;     // hi16(x) >= t  <=>  x >= t << 16
  uint32_t m = 0;
  m |= ((r.x << 16) >= th ? 1u : 0u);
  m |= (r.x >= th ? 2u : 0u);
  m |= ((r.y << 16) >= th ? 4u : 0u);
  m |= (r.y >= th ? 8u : 0u);
  m |= ((r.z << 16) >= th ? 16u : 0u);
  m |= (r.z >= th ? 32u : 0u);
  m |= ((r.w << 16) >= th ? 64u : 0u);
  m |= (r.w >= th ? 128u : 0u);
  return m;
}
// v[k] = kept ? v[k] * scale : 0 for the 8 elements of group g
__device__ __forceinline__ void dropout_apply8(float* v, uint64_t seed, uint64_t offset, uint64_t g, uint32_t thresh16,
                                               float scale) {
  const uint4 r = philox4x32(seed, g, (uint32_t)offset);
  const uint32_t th = thresh16 << 16;
  v[0] = (r.x << 16) >= th ? v[0] * scale : 0.f;
  v[1] = r.x >= th ? v[1] * scale : 0.f;
  v[2] = (r.y << 16) >= th ? v[2] * scale : 0.f;
  v[3] = r.y >= th ? v[3] * scale : 0.f;
  v[4] = (r.z << 16) >= th ? v[4] * scale : 0.f;
  v[5] = r.z >= th ? v[5] * scale : 0.f;
  v[6] = (r.w << 16) >= th ? v[6] * scale : 0.f;
  v[7] = r.w >= th ? v[7] * scale : 0.f;
}
// the 4 elements 4*e4 .. 4*e4+3 (half of group e4 >> 1)
__device__ __forceinline__ void dropout_apply4(float* v, uint64_t seed, uint64_t offset, uint64_t e4, uint32_t thresh16,
                                               float scale) {
  const uint4 r = philox4x32(seed, e4 >> 1, (uint32_t)offset);
  const uint32_t a = (e4 & 1) ? r.z : r.x, b = (e4 & 1) ? r.w : r.y;
  const uint32_t th = thresh16 << 16;
  v[0] = (a << 16) >= th ? v[0] * scale : 0.f;
  v[1] = a >= th ? v[1] * scale : 0.f;
  v[2] = (b << 16) >= th ? v[2] * scale : 0.f;
  v[3] = b >= th ? v[3] * scale : 0.f;
}
// single element e
__device__ __forceinline__ bool dropout_keep1(uint64_t seed, uint64_t offset, uint64_t e, uint32_t thresh16) {
  return (dropout_keep8(seed, offset, e >> 3, thresh16) >> (uint32_t)(e & 7)) & 1u;
}

// erf-GELU (sa_m4c.py:985-991) and its derivative from ONE exponential: erf by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, below fp32 rounding of the surrounding arithmetic), whose exp(-z^2) with
// z = x/sqrt(2) is exactly the Gaussian of the pdf term.
__device__ __forceinline__ float ptx_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ptx_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 17 instructions per element (6 FMUL, 8 FFMA, 2 MUFU, 1 LOP3); the __expf / __fdividef forms carry range fix-ups
// (FSETP + scaling multiplies) that doubled the epilogue of the FFN1 GEMM
__device__ __forceinline__ void gelu_pair(float x, float& g, float& dg) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = ptx_rcp(fmaf(0.3275911f, z, 1.0f));
  const float e = ptx_ex2(x * x * -0.72134752044448170368f);          // exp(-x^2/2) = 2^(-x^2 log2(e) / 2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float cdf = fmaf(copysignf(0.5f, x), erf_abs, 0.5f);
  g = x * cdf;
  dg = fmaf(x * 0.39894228040143267794f, e, cdf);
}
// Two elements at a time on the packed fp32 pipe (fma.rn.f32x2 / mul.rn.f32x2, SASS FFMA2 / FMUL2 on sm_100a): the same
// operations in the same order as gelu_pair -- bit-identical results -- with 14 packed + 8 scalar instructions per
// pair instead of 36 scalar ones.  The GEMM epilogues that evaluate GELU are issue-bound.
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long f2_splat(float a) { return f2_pack(a, a); }

__device__ __forceinline__ void gelu_pair2(float x0, float x1, float& g0, float& g1, float& d0, float& d1) {
  const unsigned long long x = f2_pack(x0, x1);
  const unsigned long long z = f2_mul(f2_pack(fabsf(x0), fabsf(x1)), f2_splat(0.70710678118654752440f));
  float u0, u1;
  f2_unpack(f2_fma(f2_splat(0.3275911f), z, f2_splat(1.0f)), u0, u1);
  const unsigned long long t = f2_pack(ptx_rcp(u0), ptx_rcp(u1));
  float a0, a1;
  f2_unpack(f2_mul(f2_mul(x, x), f2_splat(-0.72134752044448170368f)), a0, a1);
  const unsigned long long e = f2_pack(ptx_ex2(a0), ptx_ex2(a1));
  // the polynomial with all coefficients negated: np = -poly, so that erf_abs = fma(np * t, e, 1) = fma(-poly * t, e, 1)
  unsigned long long np = f2_fma(t, f2_splat(-1.061405429f), f2_splat(1.453152027f));
  np = f2_fma(t, np, f2_splat(-1.421413741f));
  np = f2_fma(t, np, f2_splat(0.284496736f));
  np = f2_fma(t, np, f2_splat(-0.254829592f));
  const unsigned long long erf_abs = f2_fma(f2_mul(np, t), e, f2_splat(1.0f));
  const unsigned long long cdf = f2_fma(f2_pack(copysignf(0.5f, x0), copysignf(0.5f, x1)), erf_abs, f2_splat(0.5f));
  f2_unpack(f2_mul(x, cdf), g0, g1);
  f2_unpack(f2_fma(f2_mul(x, f2_splat(0.39894228040143267794f)), e, cdf), d0, d1);
}
__device__ __forceinline__ float gelu_erf(float x) { float g, d; gelu_pair(x, g, d); return g; }
__device__ __forceinline__ float dgelu_erf(float x) { float g, d; gelu_pair(x, g, d); return d; }

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

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// 16-bit storage formats of the tensor-core operands: F16 = IEEE half (forward activations and weights: 11 significant
// bits, logits within 1e-3 of the fp32 reference), BF16 = bfloat16 (gradients: fp32 exponent range, no loss scaling)
template <bool F16> __device__ __forceinline__ uint32_t pack_16(float a, float b) {
  if constexpr (F16) return pack_f16(a, b); else return pack_bf16(a, b);
}
__device__ __forceinline__ uint32_t pack_16(float a, float b, bool f16) { return f16 ? pack_f16(a, b) : pack_bf16(a, b); }
template <bool F16> __device__ __forceinline__ float2 unpack_16(uint32_t u) {
  if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  else return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ float2 unpack_16(uint32_t u, bool f16) { return f16 ? unpack_16<true>(u) : unpack_16<false>(u); }
__device__ __forceinline__ float ld_16(const void* p, bool f16) {
  return f16 ? __half2float(*reinterpret_cast<const __half*>(p)) : __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p));
}
__device__ __forceinline__ void st_16(void* p, float v, bool f16) {
  if (f16) *reinterpret_cast<__half*>(p) = __float2half_rn(v); else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float pow2_scale_for(float amax, int target_exp) {
  // S = 2^(target_exp - e) with amax = f * 2^e, f in [0.5, 1): amax * S in [2^(target_exp-1), 2^target_exp)
  if (!(amax > 0.f) || !isfinite(amax)) return 1.0f;
  int e;
  frexpf(amax, &e);
  int k = target_exp - e;
  k = k > 120 ? 120 : (k < -120 ? -120 : k);
  return ldexpf(1.0f, k);
}
__device__ __forceinline__ uint32_t pack_f16_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

template <class T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <> struct Elem<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
};

}  // namespace samk
