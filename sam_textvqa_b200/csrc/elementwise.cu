// HBM-bound row kernels of the SA-M4C path: operand casts / bf16x3 splits, L2 normalisation,
// TF-style LayerNorm forward/backward (fused with the dropout mask and the bias-gradient column
// sums of the dense layer in front of it), dropout-add, column sums, and the two embedding
// blocks (TextBert embeddings, PrevPredEmbeddings).  One warp owns one row of <= 1024 floats.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/samk.h"

namespace samk {

int sm_count();
// SMs the grids are sized for (the device's count minus samk_reserve_sms; 148 on a B200)
static int n_sms() { const int n = sm_count(); return n > 0 ? n : 148; }


constexpr int kRowThreads = 256;           // 8 warps per block
constexpr int kMaxVec = 8;                 // float4 per lane -> cols <= 1024

// `bf16` arguments below carry the SAMK_DT_* code of the buffer: 0 = fp32, 1 = bf16, 2 = f16
__device__ __forceinline__ void store_act(void* base, int bf16, size_t idx, float4 v) {
  if (bf16) {
    const bool h = bf16 == SAMK_DT_F16;
    uint2 u = make_uint2(pack_16(v.x, v.y, h), pack_16(v.z, v.w, h));
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = u;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = v;
  }
}
__device__ __forceinline__ float4 load_act(const void* base, int bf16, size_t idx) {
  if (bf16) {
    uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    const bool h = bf16 == SAMK_DT_F16;
    float2 a = unpack_16(u.x, h), b = unpack_16(u.y, h);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
}

template <int DT>
__device__ __forceinline__ float4 load_act_t(const void* base, size_t idx) {
  if constexpr (DT == SAMK_DT_F32) {
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
  } else {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + idx);
    const float2 a = unpack_16<DT == SAMK_DT_F16>(u.x), b = unpack_16<DT == SAMK_DT_F16>(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
}

__device__ __forceinline__ float4 drop4(float4 v, uint32_t thresh, float scale, uint64_t seed, uint64_t off,
                                        uint64_t ctr) {
  if (!thresh) return v;
  float t[4] = {v.x, v.y, v.z, v.w};
  dropout_apply4(t, seed, off, ctr, thresh, scale);   // ctr = index of this group of 4 elements
  return make_float4(t[0], t[1], t[2], t[3]);
}

// ---------------------------------------------------------------------------------------------
// casts
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p, int vec) {
  if (vec) return *reinterpret_cast<const float4*>(p);
  return make_float4(p[0], p[1], p[2], p[3]);
}

__global__ void cast_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                            int rows, int cols, int vec, int f16) {
  const int c4 = cols >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)rows * c4; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / c4), c = (int)(i % c4) * 4;
    float4 v = ld4(x + (size_t)r * ldx + c, vec);
    *reinterpret_cast<uint2*>(y + (size_t)r * ldy + c) = make_uint2(pack_16(v.x, v.y, f16 != 0), pack_16(v.z, v.w, f16 != 0));
  }
}

// one read of an fp32 matrix, two 16-bit copies: half (forward operand) and bf16 (dgrad operand) of a weight
__global__ void cast_dual_kernel(const float* __restrict__ x, long long ldx, __half* __restrict__ yh, __nv_bfloat16* __restrict__ yb,
                                 long long ldy, int rows, int cols, int vec) {
  const int c4 = cols >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)rows * c4; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / c4), c = (int)(i % c4) * 4;
    float4 v = ld4(x + (size_t)r * ldx + c, vec);
    *reinterpret_cast<uint2*>(yh + (size_t)r * ldy + c) = make_uint2(pack_f16(v.x, v.y), pack_f16(v.z, v.w));
    *reinterpret_cast<uint2*>(yb + (size_t)r * ldy + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// x fp32 [rows, cols] -> three bf16 planes.  hi = bf16(x), lo = bf16(x - hi).
// order 0: (hi,hi,lo)   order 1: (hi,lo,hi).   along_rows 0: y[r, s*cols + c]; 1: y[s*rows + r, c]
__global__ void split3_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                              int rows, int cols, int order, int along_rows, int vec) {
  const int c4 = cols >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)rows * c4; i += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(i / c4), c = (int)(i % c4) * 4;
    float4 v = ld4(x + (size_t)r * ldx + c, vec);
    float h0 = bf2f(__float2bfloat16_rn(v.x)), h1 = bf2f(__float2bfloat16_rn(v.y));
    float h2 = bf2f(__float2bfloat16_rn(v.z)), h3 = bf2f(__float2bfloat16_rn(v.w));
    uint2 hi = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
    uint2 lo = make_uint2(pack_bf16(v.x - h0, v.y - h1), pack_bf16(v.z - h2, v.w - h3));
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      bool is_lo = order == 0 ? (s == 2) : (s == 1);
      size_t o = along_rows ? ((size_t)(s * rows + r) * ldy + c) : ((size_t)r * ldy + (size_t)s * cols + c);
      *reinterpret_cast<uint2*>(y + o) = is_lo ? lo : hi;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// F.normalize(x, dim=-1) (sa_m4c.py:208-209, 224-238): y = x / max(||x||_2, 1e-12)
// ---------------------------------------------------------------------------------------------
__global__ void l2norm_kernel(const float* __restrict__ x, long long ldx, void* __restrict__ y, long long ldy,
                              int y_bf16, void* __restrict__ y2, long long ldy2, int y2_bf16, int rows, int cols, int normalize) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    const float* xr = x + (size_t)r * ldx;
    float ss = 0.f;
    for (int c = lane * 4; c < cols; c += 128) {
      float4 v = *reinterpret_cast<const float4*>(xr + c);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float inv = normalize ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
    for (int c = lane * 4; c < cols; c += 128) {
      float4 v = *reinterpret_cast<const float4*>(xr + c);
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      store_act(y, y_bf16, (size_t)r * ldy + c, v);
      if (y2) store_act(y2, y2_bf16, (size_t)r * ldy2 + c, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm, TF style (sa_m4c.py:1016-1028): y = g * (x-u)/sqrt(var+eps) + b, biased variance.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowThreads)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, float* __restrict__ y, void* __restrict__ y2, int y2_bf16, void* __restrict__ y3, int y3_bf16,
                     int rows, int cols) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  pdl_wait();
  pdl_release();
  for (int r = warp; r < rows; r += nwarps) {
    float4 v[kMaxVec];
    int nvec = 0;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      int c = 4 * (lane + 32 * i);
      if (c < cols) { v[i] = *reinterpret_cast<const float4*>(x + (size_t)r * cols + c); nvec = i + 1; }
      else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // lanes past the row end contribute zeros to the sums; the variance pass must skip them
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
    s = warp_sum(s);
    const float mean = s / cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) if (i < nvec) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
    q = warp_sum(q);
    const float rstd = 1.0f / sqrtf(q / cols + eps);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) if (i < nvec) {
      int c = 4 * (lane + 32 * i);
      float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
      float4 o;
      o.x = g.x * ((v[i].x - mean) * rstd) + b.x;
      o.y = g.y * ((v[i].y - mean) * rstd) + b.y;
      o.z = g.z * ((v[i].z - mean) * rstd) + b.z;
      o.w = g.w * ((v[i].w - mean) * rstd) + b.w;
      if (y) *reinterpret_cast<float4*>(y + (size_t)r * cols + c) = o;
      if (y2) store_act(y2, y2_bf16, (size_t)r * cols + c, o);
      if (y3) store_act(y3, y3_bf16, (size_t)r * cols + c, o);
    }
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)); dgamma += dy*xhat; dbeta += dy.
// Optional second output dxd = dropout_mask(dx) in the activation dtype (the gradient of the dense
// output that sits under dropout in BertSelfOutput / BertOutput) and dbias += colsum(dxd).
// The three column sums are accumulated in a per-WARP shared-memory slice with plain vector loads /
// stores (each lane owns its columns, so there are no conflicts and no atomics); the row loop keeps few
// registers, several blocks stay resident per SM and the kernel makes a single pass over memory.  Block
// partials go to `partials` (or global atomics when it is NULL).
constexpr int kLnBwdWarps = 4;

__global__ void __launch_bounds__(kLnBwdWarps * 32)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     float eps, float* __restrict__ dx, void* __restrict__ dxd, int dxd_bf16, uint32_t thresh,
                     float scale, unsigned long long seed, unsigned long long off, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dbias, float* __restrict__ partials, int rows,
                     int cols, unsigned int* __restrict__ dxd_amax) {
  extern __shared__ float4 acc4[];   // [warps][3][cols/4]
  float amax = 0.f;                  // max |dxd| seen by this lane (for the exactly scaled half copy, samk_cast_scaled_f16)
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kLnBwdWarps + wib;
  const int nwarps = gridDim.x * kLnBwdWarps;
  const int c4n = cols >> 2;
  float4* sg = acc4 + (size_t)wib * 3 * c4n;
  float4* sb = sg + c4n;
  float4* sd = sb + c4n;
  for (int c = lane; c < 3 * c4n; c += 32) sg[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  pdl_wait();
  pdl_release();
  const bool want_d = dxd != nullptr || dbias != nullptr;
  const uint64_t row_groups = (uint64_t)((cols + 3) >> 2);
  for (int r = warp; r < rows; r += nwarps) {
    float4 v[kMaxVec], d[kMaxVec];
    int nvec = 0;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      int c = 4 * (lane + 32 * i);
      if (c < cols) {
        v[i] = *reinterpret_cast<const float4*>(x + (size_t)r * cols + c);
        d[i] = *reinterpret_cast<const float4*>(dy + (size_t)r * cols + c);
        nvec = i + 1;
        s += v[i].x + v[i].y + v[i].z + v[i].w;
      }
    }
    s = warp_sum(s);
    const float mean = s / cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) if (i < nvec) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    q = warp_sum(q);
    const float rstd = 1.0f / sqrtf(q / cols + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) if (i < nvec) {
      const int c4 = lane + 32 * i;
      float4 gm = *reinterpret_cast<const float4*>(gamma + 4 * c4);
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;   // xhat
      if (dgamma) {
        float4 t = sg[c4];
        t.x += d[i].x * v[i].x; t.y += d[i].y * v[i].y; t.z += d[i].z * v[i].z; t.w += d[i].w * v[i].w;
        sg[c4] = t;
      }
      if (dbeta) {
        float4 t = sb[c4];
        t.x += d[i].x; t.y += d[i].y; t.z += d[i].z; t.w += d[i].w;
        sb[c4] = t;
      }
      d[i] = make_float4(gm.x * d[i].x, gm.y * d[i].y, gm.z * d[i].z, gm.w * d[i].w);
      s1 += d[i].x + d[i].y + d[i].z + d[i].w;
      s2 += d[i].x * v[i].x + d[i].y * v[i].y + d[i].z * v[i].z + d[i].w * v[i].w;
    }
    s1 = warp_sum(s1) / cols;
    s2 = warp_sum(s2) / cols;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) if (i < nvec) {
      const int c4 = lane + 32 * i, c = 4 * c4;
      float4 o;
      o.x = rstd * (d[i].x - s1 - v[i].x * s2);
      o.y = rstd * (d[i].y - s1 - v[i].y * s2);
      o.z = rstd * (d[i].z - s1 - v[i].z * s2);
      o.w = rstd * (d[i].w - s1 - v[i].w * s2);
      if (dx) *reinterpret_cast<float4*>(dx + (size_t)r * cols + c) = o;
      if (want_d) {
        float4 od = drop4(o, thresh, scale, seed, off, (uint64_t)r * row_groups + (uint64_t)c4);
        if (dxd) store_act(dxd, dxd_bf16, (size_t)r * cols + c, od);
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(od.x), fabsf(od.y)), fmaxf(fabsf(od.z), fabsf(od.w))));
        if (dbias) { float4 t = sd[c4]; t.x += od.x; t.y += od.y; t.z += od.z; t.w += od.w; sd[c4] = t; }
      }
    }
  }
  if (dxd_amax) {
    amax = warp_max(amax);
    if (lane == 0 && amax > 0.f) atomicMax(dxd_amax, __float_as_uint(amax));
  }
  __syncthreads();
  // reduce the per-warp slices and publish the block partials
  float* outs[3] = {dgamma, dbeta, dbias};
  const float* accf = reinterpret_cast<const float*>(acc4);
#pragma unroll 1
  for (int k = 0; k < 3; ++k) {
    if (!outs[k]) continue;
    for (int c = threadIdx.x; c < cols; c += kLnBwdWarps * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kLnBwdWarps; ++w) t += accf[((size_t)w * 3 + k) * cols + c];
      if (partials) partials[((size_t)blockIdx.x * 3 + k) * cols + c] = t;
      else atomicAdd(outs[k] + c, t);
    }
  }
}

// out_k[c] += sum over blocks of partials[blk][k][c]; block = 32 columns x 32 block-lanes (latency-bound: the more
// independent loads in flight the better)
__global__ void __launch_bounds__(1024)
ln_bwd_finalize_kernel(const float* __restrict__ partials, int nblk, int cols, float* dgamma, float* dbeta, float* dbias) {
  __shared__ float red[32][33];
  const int k = blockIdx.y;
  float* out = k == 0 ? dgamma : (k == 1 ? dbeta : dbias);
  if (!out) return;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
  float t0 = 0.f, t1 = 0.f;
  if (c < cols) {
    int b = g;
    for (; b + 32 < nblk; b += 64) {
      t0 += partials[((size_t)b * 3 + k) * cols + c];
      t1 += partials[((size_t)(b + 32) * 3 + k) * cols + c];
    }
    if (b < nblk) t0 += partials[((size_t)b * 3 + k) * cols + c];
  }
  red[g][threadIdx.x & 31] = t0 + t1;
  __syncthreads();
  if (g == 0 && c < cols) {
    float u = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) u += red[w][threadIdx.x];
    out[c] += u;
  }
}

// out = dropout(a (+ b)); same kernel gives the backward (a = dout, b = null).
__global__ void dropout_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                   void* __restrict__ out2, int out2_bf16, int rows, int cols, uint32_t thresh,
                                   float scale, unsigned long long seed, unsigned long long off) {
  const int c4 = cols >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)rows * c4; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / c4, c = (i % c4) * 4;
    float4 v = *reinterpret_cast<const float4*>(a + r * cols + c);
    if (b) {
      float4 w = *reinterpret_cast<const float4*>(b + r * cols + c);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    v = drop4(v, thresh, scale, seed, off, r * (uint64_t)c4 + (c >> 2));
    if (out) *reinterpret_cast<float4*>(out + r * cols + c) = v;
    if (out2) store_act(out2, out2_bf16, r * cols + c, v);
  }
}

// out[c] += sum_r x[r, c]
// part_cols > 0: the columns are n consecutive groups of part_cols with separate destinations out, out1, out2
// (bias gradients of the fused q|k|v projection are three separate parameters)
template <int DT>
__global__ void __launch_bounds__(1024) colsum_kernel(const void* __restrict__ x, long long ld, int rows, int cols,
                              float* __restrict__ out, float* __restrict__ out1, float* __restrict__ out2, int part_cols) {
  // block = 1024 threads = 64 column-quads x 16 row lanes; a warp holds 16 quads (128 contiguous bytes of a bf16 row)
  // x 2 row lanes, reduced with one shuffle.  No shared memory and about one block per SM: on the side branch of the
  // backward pass a block has to fit next to a resident GEMM CTA, which owns all but ~1.7 KB of the SM's shared
  // memory (1 KB is reserved per block), so one big block per SM is what keeps enough loads in flight.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cq = blockIdx.x * 64 + (warp & 3) * 16 + (lane & 15);
  const int rl = (warp >> 2) * 2 + (lane >> 4);
  const int nrl = blockDim.x >> 6;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cq * 4 < cols) {
    const int stride = gridDim.y * nrl;
    int r = blockIdx.y * nrl + rl;
    for (; r + 3 * stride < rows; r += 4 * stride) {   // 4 independent loads in flight per thread
      float4 v0 = load_act_t<DT>(x, (size_t)r * ld + cq * 4);
      float4 v1 = load_act_t<DT>(x, (size_t)(r + stride) * ld + cq * 4);
      float4 v2 = load_act_t<DT>(x, (size_t)(r + 2 * stride) * ld + cq * 4);
      float4 v3 = load_act_t<DT>(x, (size_t)(r + 3 * stride) * ld + cq * 4);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; r < rows; r += stride) {
      float4 v = load_act_t<DT>(x, (size_t)r * ld + cq * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
  acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
  if (lane < 16 && cq * 4 < cols) {
    float* dst = out + cq * 4;
    if (part_cols > 0) {
      const int part = (cq * 4) / part_cols;
      dst = (part == 0 ? out : (part == 1 ? out1 : out2)) + (cq * 4 - part * part_cols);
    }
    atomicAdd(dst + 0, acc.x); atomicAdd(dst + 1, acc.y);
    atomicAdd(dst + 2, acc.z); atomicAdd(dst + 3, acc.w);
  }
}

// generic (any cols / ld / alignment) column sum: one warp per column
__global__ void colsum_scalar_kernel(const void* __restrict__ x, int x_bf16, long long ld, int rows, int cols,
                                     float* __restrict__ out) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (col >= cols) return;
  float acc = 0.f;
  for (int r = lane; r < rows; r += 32) {
    const size_t i = (size_t)r * ld + col;
    acc += x_bf16 ? ld_16(reinterpret_cast<const uint16_t*>(x) + i, x_bf16 == SAMK_DT_F16) : reinterpret_cast<const float*>(x)[i];
  }
  acc = warp_sum(acc);
  if (lane == 0) atomicAdd(out + col, acc);
}

// ---------------------------------------------------------------------------------------------
// Row-LN helpers for the embedding blocks (warp per row, 768-wide typical)
// ---------------------------------------------------------------------------------------------
struct RowLN {
  float4 xh[kMaxVec];  // xhat
  float rstd;
  int nvec;
};

__device__ __forceinline__ void row_ln_forward(float4* v, int cols, int lane, float eps, RowLN& st) {
  float s = 0.f;
  st.nvec = 0;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (4 * (lane + 32 * i) < cols) { st.nvec = i + 1; s += v[i].x + v[i].y + v[i].z + v[i].w; }
  }
  s = warp_sum(s);
  const float mean = s / cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
    st.xh[i] = make_float4(v[i].x - mean, v[i].y - mean, v[i].z - mean, v[i].w - mean);
    q += st.xh[i].x * st.xh[i].x + st.xh[i].y * st.xh[i].y + st.xh[i].z * st.xh[i].z + st.xh[i].w * st.xh[i].w;
  }
  q = warp_sum(q);
  st.rstd = 1.0f / sqrtf(q / cols + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
    st.xh[i].x *= st.rstd; st.xh[i].y *= st.rstd; st.xh[i].z *= st.rstd; st.xh[i].w *= st.rstd;
  }
}

// given dy (per lane vectors) -> dx in place; accumulates dgamma/dbeta with atomics
__device__ __forceinline__ void row_ln_backward(float4* d, const RowLN& st, const float* gamma, float* dgamma,
                                                float* dbeta, int cols, int lane) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
    int c = 4 * (lane + 32 * i);
    float4 gm = *reinterpret_cast<const float4*>(gamma + c);
    atomicAdd(dgamma + c + 0, d[i].x * st.xh[i].x); atomicAdd(dgamma + c + 1, d[i].y * st.xh[i].y);
    atomicAdd(dgamma + c + 2, d[i].z * st.xh[i].z); atomicAdd(dgamma + c + 3, d[i].w * st.xh[i].w);
    atomicAdd(dbeta + c + 0, d[i].x); atomicAdd(dbeta + c + 1, d[i].y);
    atomicAdd(dbeta + c + 2, d[i].z); atomicAdd(dbeta + c + 3, d[i].w);
    d[i] = make_float4(gm.x * d[i].x, gm.y * d[i].y, gm.z * d[i].z, gm.w * d[i].w);
    s1 += d[i].x + d[i].y + d[i].z + d[i].w;
    s2 += d[i].x * st.xh[i].x + d[i].y * st.xh[i].y + d[i].z * st.xh[i].z + d[i].w * st.xh[i].w;
  }
  s1 = warp_sum(s1) / cols;
  s2 = warp_sum(s2) / cols;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
    d[i].x = st.rstd * (d[i].x - s1 - st.xh[i].x * s2);
    d[i].y = st.rstd * (d[i].y - s1 - st.xh[i].y * s2);
    d[i].z = st.rstd * (d[i].z - s1 - st.xh[i].z * s2);
    d[i].w = st.rstd * (d[i].w - s1 - st.xh[i].w * s2);
  }
}

// same, but dgamma / dbeta contributions are added to per-lane register accumulators (lane owns fixed columns), to be
// reduced per block by block_flush_columns: thousands of rows hammering the same 768 addresses with global atomics
// cost 260 us in the TextBert embedding backward
__device__ __forceinline__ void row_ln_backward_acc(float4* d, const RowLN& st, const float* gamma, float4* ag, float4* ab,
                                                    int cols, int lane) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
    int c = 4 * (lane + 32 * i);
    float4 gm = *reinterpret_cast<const float4*>(gamma + c);
    ag[i].x += d[i].x * st.xh[i].x; ag[i].y += d[i].y * st.xh[i].y; ag[i].z += d[i].z * st.xh[i].z; ag[i].w += d[i].w * st.xh[i].w;
    ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
    d[i] = make_float4(gm.x * d[i].x, gm.y * d[i].y, gm.z * d[i].z, gm.w * d[i].w);
    s1 += d[i].x + d[i].y + d[i].z + d[i].w;
    s2 += d[i].x * st.xh[i].x + d[i].y * st.xh[i].y + d[i].z * st.xh[i].z + d[i].w * st.xh[i].w;
  }
  s1 = warp_sum(s1) / cols;
  s2 = warp_sum(s2) / cols;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
    d[i].x = st.rstd * (d[i].x - s1 - st.xh[i].x * s2);
    d[i].y = st.rstd * (d[i].y - s1 - st.xh[i].y * s2);
    d[i].z = st.rstd * (d[i].z - s1 - st.xh[i].z * s2);
    d[i].w = st.rstd * (d[i].w - s1 - st.xh[i].w * s2);
  }
}

// same again, accumulating dgamma / dbeta into shared-memory arrays of the block (kernels with too many hot arrays
// for register accumulators)
__device__ __forceinline__ void row_ln_backward_smem(float4* d, const RowLN& st, const float* gamma, float* sg, float* sb,
                                                     int cols, int lane) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
    int c = 4 * (lane + 32 * i);
    float4 gm = *reinterpret_cast<const float4*>(gamma + c);
    atomicAdd(sg + c + 0, d[i].x * st.xh[i].x); atomicAdd(sg + c + 1, d[i].y * st.xh[i].y);
    atomicAdd(sg + c + 2, d[i].z * st.xh[i].z); atomicAdd(sg + c + 3, d[i].w * st.xh[i].w);
    atomicAdd(sb + c + 0, d[i].x); atomicAdd(sb + c + 1, d[i].y);
    atomicAdd(sb + c + 2, d[i].z); atomicAdd(sb + c + 3, d[i].w);
    d[i] = make_float4(gm.x * d[i].x, gm.y * d[i].y, gm.z * d[i].z, gm.w * d[i].w);
    s1 += d[i].x + d[i].y + d[i].z + d[i].w;
    s2 += d[i].x * st.xh[i].x + d[i].y * st.xh[i].y + d[i].z * st.xh[i].z + d[i].w * st.xh[i].w;
  }
  s1 = warp_sum(s1) / cols;
  s2 = warp_sum(s2) / cols;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
    d[i].x = st.rstd * (d[i].x - s1 - st.xh[i].x * s2);
    d[i].y = st.rstd * (d[i].y - s1 - st.xh[i].y * s2);
    d[i].z = st.rstd * (d[i].z - s1 - st.xh[i].z * s2);
    d[i].w = st.rstd * (d[i].w - s1 - st.xh[i].w * s2);
  }
}

// block-wide: sum the per-lane column accumulators of all warps in shared memory, then one global atomic per column
__device__ __forceinline__ void block_flush_columns(const float4* acc, float* sm /*[cols]*/, float* out, int cols, int lane) {
  for (int c = threadIdx.x; c < cols; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    int c = 4 * (lane + 32 * i);
    if (c < cols) {
      atomicAdd(sm + c + 0, acc[i].x); atomicAdd(sm + c + 1, acc[i].y);
      atomicAdd(sm + c + 2, acc[i].z); atomicAdd(sm + c + 3, acc[i].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) atomicAdd(out + c, sm[c]);
  __syncthreads();
}

__device__ __forceinline__ void atomic_add4(float* p, float4 v) {
  atomicAdd(p + 0, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
}

// TextBert embeddings (sa_m4c.py:383 -> BertEmbeddings): dropout(LN(word[id] + pos[t] + type[0]))
__global__ void __launch_bounds__(kRowThreads)
bert_embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
                      const float* __restrict__ type, const float* __restrict__ gamma, const float* __restrict__ beta,
                      float eps, float* __restrict__ out, void* __restrict__ out2, int out2_bf16, int rows, int T,
                      int cols, uint32_t thresh, float scale, unsigned long long seed, unsigned long long off) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    const long long id = ids[r];
    const int t = r % T;
    float4 v[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      int c = 4 * (lane + 32 * i);
      if (c < cols) {
        float4 a = *reinterpret_cast<const float4*>(word + (size_t)id * cols + c);
        float4 b = *reinterpret_cast<const float4*>(pos + (size_t)t * cols + c);
        float4 d = *reinterpret_cast<const float4*>(type + c);
        v[i] = make_float4(a.x + b.x + d.x, a.y + b.y + d.y, a.z + b.z + d.z, a.w + b.w + d.w);
      }
    }
    RowLN st;
    row_ln_forward(v, cols, lane, eps, st);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
      int c = 4 * (lane + 32 * i);
      float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
      float4 o = make_float4(g.x * st.xh[i].x + b.x, g.y * st.xh[i].y + b.y, g.z * st.xh[i].z + b.z, g.w * st.xh[i].w + b.w);
      o = drop4(o, thresh, scale, seed, off, (uint64_t)r * (cols >> 2) + (c >> 2));
      *reinterpret_cast<float4*>(out + (size_t)r * cols + c) = o;
      if (out2) store_act(out2, out2_bf16, (size_t)r * cols + c, o);
    }
  }
}

__global__ void __launch_bounds__(kRowThreads)
bert_embed_bwd_kernel(const float* __restrict__ dout, const long long* __restrict__ ids, const float* __restrict__ word,
                      const float* __restrict__ pos, const float* __restrict__ type, const float* __restrict__ gamma,
                      float eps, float* __restrict__ dword, float* __restrict__ dpos, float* __restrict__ dtype,
                      float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int T, int cols, uint32_t thresh,
                      float scale, unsigned long long seed, unsigned long long off) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  __shared__ float sm_cols[1024];
  float4 ag[kMaxVec], ab[kMaxVec], at[kMaxVec];      // dgamma, dbeta, d(token_type row 0): identical addresses for every row
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) ag[i] = ab[i] = at[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = warp; r < rows; r += nwarps) {
    const long long id = ids[r];
    const int t = r % T;
    float4 v[kMaxVec], d[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      int c = 4 * (lane + 32 * i);
      if (c < cols) {
        float4 a = *reinterpret_cast<const float4*>(word + (size_t)id * cols + c);
        float4 b = *reinterpret_cast<const float4*>(pos + (size_t)t * cols + c);
        float4 e = *reinterpret_cast<const float4*>(type + c);
        v[i] = make_float4(a.x + b.x + e.x, a.y + b.y + e.y, a.z + b.z + e.z, a.w + b.w + e.w);
        d[i] = drop4(*reinterpret_cast<const float4*>(dout + (size_t)r * cols + c), thresh, scale, seed, off,
                     (uint64_t)r * (cols >> 2) + (c >> 2));
      }
    }
    RowLN st;
    row_ln_forward(v, cols, lane, eps, st);
    row_ln_backward_acc(d, st, gamma, ag, ab, cols, lane);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) if (i < st.nvec) {
      int c = 4 * (lane + 32 * i);
      if (id != 0) atomic_add4(dword + (size_t)id * cols + c, d[i]);  // padding_idx = 0 gets no gradient
      atomic_add4(dpos + (size_t)t * cols + c, d[i]);
      at[i].x += d[i].x; at[i].y += d[i].y; at[i].z += d[i].z; at[i].w += d[i].w;
    }
  }
  block_flush_columns(ag, sm_cols, dgamma, cols, lane);
  block_flush_columns(ab, sm_cols, dbeta, cols, lane);
  block_flush_columns(at, sm_cols, dtype, cols, lane);
}

// PrevPredEmbeddings (sa_m4c.py:919-948): row (b,t): idx = prev[b,t];
//   raw = idx < V ? LN_ans(cls_w[idx]) : LN_ocr(ocr_in[b, idx-V]);  out = raw + dropout(LN_emb(pos[t] + type[idx>=V]))
struct PrevPredParams {
  const long long* prev; const float* cls_w; const float* ocr_in; const float* pos; const float* type;
  const float* ans_g; const float* ans_b; const float* ocr_g; const float* ocr_b; const float* emb_g; const float* emb_b;
  float eps; int B, D, V, R, cols; uint32_t thresh; float scale; unsigned long long seed, off;
};

__global__ void __launch_bounds__(kRowThreads)
prevpred_fwd_kernel(PrevPredParams p, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int cols = p.cols;
  for (int r = warp; r < p.B * p.D; r += nwarps) {
    const int b = r / p.D, t = r % p.D;
    long long idx = p.prev[r];
    idx = idx < 0 ? 0 : (idx >= p.V + p.R ? p.V + p.R - 1 : idx);   // reference would raise IndexError
    const bool is_ocr = idx >= p.V;
    const float* src = is_ocr ? p.ocr_in + ((size_t)b * p.R + (idx - p.V)) * cols : p.cls_w + (size_t)idx * cols;
    const float* g1 = is_ocr ? p.ocr_g : p.ans_g;
    const float* b1 = is_ocr ? p.ocr_b : p.ans_b;
    float4 v[kMaxVec], e[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      int c = 4 * (lane + 32 * i);
      if (c < cols) {
        v[i] = *reinterpret_cast<const float4*>(src + c);
        float4 a = *reinterpret_cast<const float4*>(p.pos + (size_t)t * cols + c);
        float4 d = *reinterpret_cast<const float4*>(p.type + (size_t)(is_ocr ? 1 : 0) * cols + c);
        e[i] = make_float4(a.x + d.x, a.y + d.y, a.z + d.z, a.w + d.w);
      }
    }
    RowLN s1, s2;
    row_ln_forward(v, cols, lane, p.eps, s1);
    row_ln_forward(e, cols, lane, p.eps, s2);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) if (i < s1.nvec) {
      int c = 4 * (lane + 32 * i);
      float4 ga = *reinterpret_cast<const float4*>(g1 + c), ba = *reinterpret_cast<const float4*>(b1 + c);
      float4 ge = *reinterpret_cast<const float4*>(p.emb_g + c), be = *reinterpret_cast<const float4*>(p.emb_b + c);
      float4 o2 = make_float4(ge.x * s2.xh[i].x + be.x, ge.y * s2.xh[i].y + be.y, ge.z * s2.xh[i].z + be.z, ge.w * s2.xh[i].w + be.w);
      o2 = drop4(o2, p.thresh, p.scale, p.seed, p.off, (uint64_t)r * (cols >> 2) + (c >> 2));
      float4 o = make_float4(ga.x * s1.xh[i].x + ba.x + o2.x, ga.y * s1.xh[i].y + ba.y + o2.y,
                             ga.z * s1.xh[i].z + ba.z + o2.z, ga.w * s1.xh[i].w + ba.w + o2.w);
      *reinterpret_cast<float4*>(out + (size_t)r * cols + c) = o;
    }
  }
}

struct PrevPredGrads {
  float* d_cls_w; float* d_ocr_in; float* d_pos; float* d_type;
  float* d_ans_g; float* d_ans_b; float* d_ocr_g; float* d_ocr_b; float* d_emb_g; float* d_emb_b;
};

__global__ void __launch_bounds__(kRowThreads)
prevpred_bwd_kernel(PrevPredParams p, const float* __restrict__ dout, PrevPredGrads g) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int cols = p.cols;
  // gradients that every row adds to the same few addresses (three LayerNorms, the two token-type rows) are summed
  // per block in shared memory first: [ans_g, ans_b, ocr_g, ocr_b, emb_g, emb_b, type0, type1][cols]
  extern __shared__ float sacc[];
  for (int c = threadIdx.x; c < 8 * cols; c += blockDim.x) sacc[c] = 0.f;
  __syncthreads();
  for (int r = warp; r < p.B * p.D; r += nwarps) {
    const int b = r / p.D, t = r % p.D;
    long long idx = p.prev[r];
    idx = idx < 0 ? 0 : (idx >= p.V + p.R ? p.V + p.R - 1 : idx);   // reference would raise IndexError
    const bool is_ocr = idx >= p.V;
    const size_t src_off = is_ocr ? ((size_t)b * p.R + (idx - p.V)) * cols : (size_t)idx * cols;
    const float* src = (is_ocr ? p.ocr_in : p.cls_w) + src_off;
    float4 v[kMaxVec], e[kMaxVec], d1[kMaxVec], d2[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      int c = 4 * (lane + 32 * i);
      if (c < cols) {
        v[i] = *reinterpret_cast<const float4*>(src + c);
        float4 a = *reinterpret_cast<const float4*>(p.pos + (size_t)t * cols + c);
        float4 d = *reinterpret_cast<const float4*>(p.type + (size_t)(is_ocr ? 1 : 0) * cols + c);
        e[i] = make_float4(a.x + d.x, a.y + d.y, a.z + d.z, a.w + d.w);
        d1[i] = *reinterpret_cast<const float4*>(dout + (size_t)r * cols + c);
        d2[i] = drop4(d1[i], p.thresh, p.scale, p.seed, p.off, (uint64_t)r * (cols >> 2) + (c >> 2));
      }
    }
    RowLN s1, s2;
    row_ln_forward(v, cols, lane, p.eps, s1);
    row_ln_forward(e, cols, lane, p.eps, s2);
    row_ln_backward_smem(d1, s1, is_ocr ? p.ocr_g : p.ans_g, sacc + (is_ocr ? 2 : 0) * cols, sacc + (is_ocr ? 3 : 1) * cols, cols, lane);
    row_ln_backward_smem(d2, s2, p.emb_g, sacc + 4 * cols, sacc + 5 * cols, cols, lane);
    float* dsrc = (is_ocr ? g.d_ocr_in : g.d_cls_w) + src_off;
    float* stype = sacc + (is_ocr ? 7 : 6) * cols;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) if (i < s1.nvec) {
      int c = 4 * (lane + 32 * i);
      atomic_add4(dsrc + c, d1[i]);
      atomic_add4(g.d_pos + (size_t)t * cols + c, d2[i]);
      atomic_add4(stype + c, d2[i]);
    }
  }
  __syncthreads();
  float* const dst[8] = {g.d_ans_g, g.d_ans_b, g.d_ocr_g, g.d_ocr_b, g.d_emb_g, g.d_emb_b, g.d_type, g.d_type + cols};
#pragma unroll
  for (int a = 0; a < 8; ++a)
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      const float x = sacc[a * cols + c];
      if (x != 0.f) atomicAdd(dst[a] + c, x);
    }
}

// ---------------------------------------------------------------------------------------------
// Pointer network scores (sa_m4c.py:878-897): out[b,t,V+r] = q[b,t,:].k[b,r,:]/sqrt(dq) + (1-mask[b,r])*-1e4
// ---------------------------------------------------------------------------------------------
__global__ void ptr_scores_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                      const long long* __restrict__ mask, float* __restrict__ out, long long ldo,
                                      int col_off, int B, int D, int R, int dq, float inv_sqrt) {
  // one warp per (b, t, r)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = B * D * R;
  if (warp >= total) return;
  const int b = warp / (D * R), t = (warp / R) % D, r = warp % R;
  const float* qr = q + ((size_t)b * D + t) * dq;
  const float* kr = k + ((size_t)b * R + r) * dq;
  float s = 0.f;
  for (int c = lane * 4; c < dq; c += 128) {
    float4 a = *reinterpret_cast<const float4*>(qr + c), w = *reinterpret_cast<const float4*>(kr + c);
    s += a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
  }
  s = warp_sum(s);
  if (lane == 0) {
    float m = (1.0f - (float)mask[(size_t)b * R + r]) * -10000.0f;
    out[((size_t)b * D + t) * ldo + col_off + r] = s * inv_sqrt + m;
  }
}

// dq[b,t,:] = inv * sum_r ds[b,t,r] k[b,r,:] ; dk[b,r,:] = inv * sum_t ds[b,t,r] q[b,t,:]
__global__ void ptr_scores_bwd_kernel(const float* __restrict__ ds, long long ldds, int col_off,
                                      const float* __restrict__ q, const float* __restrict__ k, float* __restrict__ dq_,
                                      float* __restrict__ dk_, int B, int D, int R, int dq, float inv_sqrt) {
  // one warp per output row: first B*D rows of dq then B*R rows of dk
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nq = B * D, nk = B * R;
  if (warp >= nq + nk) return;
  // all column chunks of the row are accumulated together: kMaxVec independent loads in flight per step of the
  // (short, latency-bound) contraction loop
  float4 acc[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp < nq) {
    const int b = warp / D;
    const float* dsr = ds + (size_t)warp * ldds + col_off;
    for (int r = 0; r < R; ++r) {
      const float w = dsr[r] * inv_sqrt;
      const float* kr = k + ((size_t)b * R + r) * dq;
#pragma unroll
      for (int i = 0; i < kMaxVec; ++i) {
        const int c = 4 * (lane + 32 * i);
        if (c < dq) {
          const float4 kv = *reinterpret_cast<const float4*>(kr + c);
          acc[i].x += w * kv.x; acc[i].y += w * kv.y; acc[i].z += w * kv.z; acc[i].w += w * kv.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < dq) *reinterpret_cast<float4*>(dq_ + (size_t)warp * dq + c) = acc[i];
    }
  } else {
    const int row = warp - nq, b = row / R, r = row % R;
    for (int t = 0; t < D; ++t) {
      const float w = ds[((size_t)b * D + t) * ldds + col_off + r] * inv_sqrt;
      const float* qr = q + ((size_t)b * D + t) * dq;
#pragma unroll
      for (int i = 0; i < kMaxVec; ++i) {
        const int c = 4 * (lane + 32 * i);
        if (c < dq) {
          const float4 qv = *reinterpret_cast<const float4*>(qr + c);
          acc[i].x += w * qv.x; acc[i].y += w * qv.y; acc[i].z += w * qv.z; acc[i].w += w * qv.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = 4 * (lane + 32 * i);
      if (c < dq) *reinterpret_cast<float4*>(dk_ + (size_t)row * dq + c) = acc[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Masked BCE-with-logits (sam/task_utils.py:19-30), forward value and d(loss)/d(scores) in one pass:
//   loss = sum_{b,t,v} mask[b,t] * bce(x, y) / max(sum(mask), 1)
// ---------------------------------------------------------------------------------------------
__global__ void bce_loss_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ mask,
                                float* __restrict__ dx, float* __restrict__ loss_sum, const float* __restrict__ mask_sum,
                                long long n, int ncls) {
  float acc = 0.f;
  const float inv_cnt = 1.0f / fmaxf(*mask_sum, 1.0f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * blockDim.x) {
    const float m = mask[i / ncls];
    const float xv = x[i], yv = y[i];
    const float l = fmaxf(xv, 0.f) - xv * yv + log1pf(__expf(-fabsf(xv)));
    acc += m * l;
    if (dx) dx[i] = m * (1.0f / (1.0f + __expf(-xv)) - yv) * inv_cnt;
  }
  acc = warp_sum(acc);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(loss_sum, t * inv_cnt);
  }
}

__global__ void sum_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * blockDim.x) acc += x[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

__global__ void scale_kernel(float* __restrict__ x, long long n, const float* __restrict__ s) {
  const float v = *s;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * blockDim.x) x[i] *= v;
}

// ---- optimizer step on flat buffers (sam/task_utils.py:33-34 clip_grad_norm_, :42 torch.optim.Adam defaults) ----
__global__ void sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) acc += x[i] * x[i];
  acc = warp_sum(acc);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, (double)t);
  }
}

// p, m, v updated in place; g read only.  clip: g *= min(max_norm / (||g_all|| + 1e-6), 1) with ||g_all||^2 in *sumsq
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float neg_step_size, float one_minus_beta1, float beta2,
                            float one_minus_beta2, float eps, float sqrt_bc2, const double* __restrict__ sumsq,
                            float max_norm) {
  float coef = 1.0f;
  if (sumsq) {
    const float norm = (float)sqrt(*sumsq);
    coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
  }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += stride) {
    const float gi = g[i] * coef;
    const float mi = m[i] + one_minus_beta1 * (gi - m[i]);           // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * beta2 + one_minus_beta2 * gi * gi;       // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrt_bc2 + eps;                  // (sqrt(v) / sqrt(1 - beta2^t)) + eps
    p[i] = p[i] + (neg_step_size * mi) / denom;                      // addcdiv_(exp_avg, denom, value = -lr / (1 - beta1^t))
  }
}

static inline int grid_for(long long work, int per_block) {
  long long g = (work + per_block - 1) / per_block;
  int cap = n_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
static inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace samk

using namespace samk;

// ---------------------------------------------------------------------------------------------
// Exactly scaled f16 copy of a gradient tensor: pass 1 finds max|x| (non-negative floats order like their bit
// patterns, so an unsigned atomicMax does it), pass 2 multiplies by the power of two that puts the maximum into
// [2^11, 2^12) and rounds to IEEE half (saturating).  scale2 = {S, 1/S}; 1/S is the alpha of the consuming GEMM.
// ---------------------------------------------------------------------------------------------
__global__ void amax_kernel(const void* __restrict__ x, int dt, long long n4, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = load_act(x, dt, (size_t)i * 4);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}
__global__ void cast_scaled_f16_kernel(const void* __restrict__ x, int dt, long long n4, const unsigned int* __restrict__ amax_bits,
                                       __half* __restrict__ y, float* __restrict__ scale2) {
  const float S = pow2_scale_for(__uint_as_float(*amax_bits), 12);
  if (blockIdx.x == 0 && threadIdx.x == 0) { scale2[0] = S; scale2[1] = 1.0f / S; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = load_act(x, dt, (size_t)i * 4);
    *reinterpret_cast<uint2*>(y + i * 4) = make_uint2(pack_f16_sat(v.x * S, v.y * S), pack_f16_sat(v.z * S, v.w * S));
  }
}

// flat copy between fp32 and a 16-bit format (either direction), 4 elements per thread + scalar tail
__global__ void cast_flat_kernel(const void* __restrict__ x, int xdt, void* __restrict__ y, int ydt, long long n) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    store_act(y, ydt, (size_t)i * 4, load_act(x, xdt, (size_t)i * 4));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    const float v = xdt ? ld_16(reinterpret_cast<const uint16_t*>(x) + i, xdt == SAMK_DT_F16) : reinterpret_cast<const float*>(x)[i];
    if (ydt) st_16(reinterpret_cast<uint16_t*>(y) + i, v, ydt == SAMK_DT_F16); else reinterpret_cast<float*>(y)[i] = v;
  }
}

// Up to 4 row segments of a [B, L, d] fp32 tensor <-> separate contiguous [B, rows_k, d] tensors, one launch.
//   to_joint = 1: joint[b, off_k + i, :] = seg_k[b, i, :]     (MMT input: [txt ; obj ; ocr ; dec], sa_m4c.py:790)
//   to_joint = 0: seg_k[b, i, :] = joint[b, off_k + i, :]     (its backward; decoder / OCR rows for the output heads)
struct RowSegments { float* ptr[4]; int rows[4]; int off[4]; int n; };
__global__ void row_segments_kernel(float* __restrict__ joint, RowSegments sg, int B, int L, int d4, int to_joint) {
  int total = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) total += k < sg.n ? sg.rows[k] : 0;
  const long long n = (long long)B * total * d4;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % d4);
    long long r = e / d4;
    const int b = (int)(r / total);
    int i = (int)(r - (long long)b * total);
    int k = 0;
    while (k + 1 < sg.n && i >= sg.rows[k]) { i -= sg.rows[k]; ++k; }
    float4* sp = reinterpret_cast<float4*>(sg.ptr[k]) + ((size_t)b * sg.rows[k] + i) * d4 + c;
    float4* jp = reinterpret_cast<float4*>(joint) + ((size_t)b * L + sg.off[k] + i) * d4 + c;
    if (to_joint) *jp = *sp; else *sp = *jp;
  }
}

// key_valid[b, :] = [question_mask ; obj_mask ; ocr_mask ; zeros(D)] != 0 as bytes (sa_m4c.py:793-795: the decoder
// part of the joint attention mask is zeros, causality is handled by the kernels)
__global__ void key_valid_kernel(const long long* __restrict__ q, const long long* __restrict__ o, const long long* __restrict__ r,
                                 uint8_t* __restrict__ out, int B, int T, int O, int R, int D) {
  const int L = T + O + R + D;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * L) return;
  const int b = e / L, j = e - b * L;
  long long v = 0;
  if (j < T) v = q[(size_t)b * T + j];
  else if (j < T + O) v = o[(size_t)b * O + (j - T)];
  else if (j < T + O + R) v = r[(size_t)b * R + (j - T - O)];
  out[e] = v != 0 ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Beam-search step (sam/beam_search.py:88-130): for every sample, the K best of the K x ncls candidates
//   score(k, c) = log sigmoid(scores[b K + k, t, c]) + beam_score[b K + k]
// with the reference's rules: a completed beam may only continue with EOS at no cost (:92-96); at t = 0 only beam 0
// of every sample counts (:101-108).  One block per sample: every thread keeps its K best in registers, then K
// rounds of a block-wide arg-max pop the winners in descending order (ties: lowest candidate index).
// ---------------------------------------------------------------------------------------------
constexpr int kBeamMax = 16;
__global__ void __launch_bounds__(256)
beam_step_kernel(const float* __restrict__ scores, long long row_stride, int ncls, const float* __restrict__ beam_scores,
                 const uint8_t* __restrict__ completed, int eos, int first_step, int K, long long* __restrict__ prev_pos,
                 long long* __restrict__ new_pos, float* __restrict__ new_scores) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_win;
  const int b = blockIdx.x, tid = threadIdx.x;
  float best[kBeamMax];
  int bidx[kBeamMax];
#pragma unroll
  for (int i = 0; i < kBeamMax; ++i) { best[i] = -INFINITY; bidx[i] = 0x7fffffff; }
  const int total = K * ncls;
  for (int e = tid; e < total; e += blockDim.x) {
    const int k = e / ncls, c = e - k * ncls;
    const int row = b * K + k;
    float v;
    if (first_step && k > 0) v = -INFINITY;
    else if (completed && completed[row]) v = (c == eos) ? beam_scores[row] : -INFINITY;
    else {
      const float x = scores[(size_t)row * row_stride + c];
      v = fminf(x, 0.f) - log1pf(__expf(-fabsf(x))) + beam_scores[row];
    }
    if (v > best[K - 1] || (v == best[K - 1] && e < bidx[K - 1])) {      // insert into the sorted local list
      int j = K - 1;
      while (j > 0 && (v > best[j - 1] || (v == best[j - 1] && e < bidx[j - 1]))) { best[j] = best[j - 1]; bidx[j] = bidx[j - 1]; --j; }
      best[j] = v; bidx[j] = e;
    }
  }
  int head = 0;                       // next unpopped entry of this thread's list
  for (int r = 0; r < K; ++r) {
    float v = head < K ? best[head] : -INFINITY;
    int ix = head < K ? bidx[head] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
      if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
    }
    if ((tid & 31) == 0) { s_val[tid >> 5] = v; s_idx[tid >> 5] = ix; }
    __syncthreads();
    if (tid == 0) {
      float bv = s_val[0]; int bi = s_idx[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
      s_win = bi;
      const int out = b * K + r;
      const int k = bi == 0x7fffffff ? 0 : bi / ncls;
      prev_pos[out] = (long long)b * K + k;
      new_pos[out] = bi == 0x7fffffff ? eos : bi - k * ncls;
      new_scores[out] = bv;
    }
    __syncthreads();
    if (head < K && bidx[head] == s_win) ++head;
    __syncthreads();
  }
}

// row-wise arg-max of a [rows, ncls] fp32 matrix (first maximum), and whether the target at that index is set:
// the prediction step of the greedy decoder (sa_m4c.py:299-301) and a token-level hit count without moving the
// [B, D, V+R] logits to the host (sam/datasets/metrics.py:26 reads only the arg-max)
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const float* __restrict__ x, long long ld, int ncls, const float* __restrict__ targets, long long ldt,
                   long long* __restrict__ idx_out, float* __restrict__ hit_out) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const long long row = blockIdx.x;
  const float* xr = x + row * ld;
  float v = -INFINITY;
  int ix = 0x7fffffff;
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) {
    const float u = xr[c];
    if (ix == 0x7fffffff || u > v) { v = u; ix = c; }        // c ascends per thread: the first maximum stays
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
    if (oi != 0x7fffffff && (ix == 0x7fffffff || ov > v || (ov == v && oi < ix))) { v = ov; ix = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = v; s_idx[threadIdx.x >> 5] = ix; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (s_idx[w] != 0x7fffffff && (ix == 0x7fffffff || s_val[w] > v || (s_val[w] == v && s_idx[w] < ix))) { v = s_val[w]; ix = s_idx[w]; }
    if (ix == 0x7fffffff) ix = 0;
    idx_out[row] = ix;
    if (hit_out) hit_out[row] = targets ? targets[row * ldt + ix] : 0.f;
  }
}

// out[0:n] = a, out[n:2n] = b, out[2n:3n] = c (the three biases of the fused q|k|v projection)
__global__ void concat3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                               float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * n) return;
  out[i] = i < n ? a[i] : (i < 2 * n ? b[i - n] : c[i - 2 * n]);
}

#define SAMK_REQUIRE(cond, msg) \
  do { if (!(cond)) { set_error("%s: %s", __func__, msg); return SAMK_ERR_ARG; } } while (0)

extern "C" {

int samk_cast_16(const float* x, long long ldx, void* y, long long ldy, int y_dtype, int rows, int cols, void* stream) {
  SAMK_REQUIRE(x && y && rows >= 0 && cols >= 0, "bad argument");
  SAMK_REQUIRE(y_dtype == SAMK_DT_BF16 || y_dtype == SAMK_DT_F16, "y_dtype must be a 16-bit format");
  SAMK_REQUIRE(cols % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)y & 7) == 0, "cols/ldy must be multiples of 4");
  if (!rows || !cols) return SAMK_OK;
  cast_kernel<<<grid_for((long long)rows * cols / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, (__nv_bfloat16*)y, ldy, rows, cols, (ldx % 4 == 0 && al16(x)) ? 1 : 0, y_dtype == SAMK_DT_F16 ? 1 : 0);
  return check_launch(__func__);
}

int samk_cast_dual(const float* x, long long ldx, void* y_f16, void* y_bf16, long long ldy, int rows, int cols, void* stream) {
  SAMK_REQUIRE(x && y_f16 && y_bf16 && rows >= 0 && cols >= 0, "bad argument");
  SAMK_REQUIRE(cols % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)y_f16 & 7) == 0 && ((uintptr_t)y_bf16 & 7) == 0, "cols/ldy must be multiples of 4");
  if (!rows || !cols) return SAMK_OK;
  cast_dual_kernel<<<grid_for((long long)rows * cols / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, (__half*)y_f16, (__nv_bfloat16*)y_bf16, ldy, rows, cols, (ldx % 4 == 0 && al16(x)) ? 1 : 0);
  return check_launch(__func__);
}

int samk_cast_bf16(const float* x, long long ldx, void* y, long long ldy, int rows, int cols, void* stream) {
  return samk_cast_16(x, ldx, y, ldy, SAMK_DT_BF16, rows, cols, stream);
}

int samk_split3_bf16(const float* x, long long ldx, void* y, long long ldy, int rows, int cols, int order,
                     int along_rows, void* stream) {
  SAMK_REQUIRE(x && y && rows >= 0 && cols >= 0, "bad argument");
  SAMK_REQUIRE(cols % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)y & 7) == 0, "cols/ldy must be multiples of 4");
  if (!rows || !cols) return SAMK_OK;
  split3_kernel<<<grid_for((long long)rows * cols / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, (__nv_bfloat16*)y, ldy, rows, cols, order, along_rows, (ldx % 4 == 0 && al16(x)) ? 1 : 0);
  return check_launch(__func__);
}

int samk_l2norm(const float* x, long long ldx, void* y, long long ldy, int y_dtype, int rows, int cols, int normalize,
                void* stream) {
  SAMK_REQUIRE(x && y && rows >= 0 && cols >= 0, "bad argument");
  SAMK_REQUIRE(cols % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16(x) && ((uintptr_t)y & 7) == 0, "cols/ld must be multiples of 4");
  if (!rows || !cols) return SAMK_OK;
  l2norm_kernel<<<grid_for(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, y_dtype, nullptr, 0, 0, rows, cols, normalize);
  return check_launch(__func__);
}

int samk_l2norm2(const float* x, long long ldx, void* y, long long ldy, int y_dtype, void* y2, long long ldy2, int y2_dtype,
                 int rows, int cols, int normalize, void* stream) {
  SAMK_REQUIRE(x && y && y2 && rows >= 0 && cols >= 0, "bad argument");
  SAMK_REQUIRE(cols % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ldy2 % 4 == 0 && al16(x) && ((uintptr_t)y & 7) == 0 &&
               ((uintptr_t)y2 & 7) == 0, "cols/ld must be multiples of 4");
  if (!rows || !cols) return SAMK_OK;
  l2norm_kernel<<<grid_for(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, y_dtype, y2, ldy2, y2_dtype, rows, cols, normalize);
  return check_launch(__func__);
}

int samk_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y, void* y2,
                       int y2_dtype, void* y3, int y3_dtype, int rows, int cols, void* stream) {
  SAMK_REQUIRE(x && gamma && beta && (y || y2 || y3) && rows >= 0, "bad argument");
  SAMK_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 1024, "cols must be a multiple of 4, <= 1024");
  if (!rows) return SAMK_OK;
  launch_maybe_pdl(layernorm_fwd_kernel, dim3(grid_for(rows, 8)), dim3(kRowThreads), 0, (cudaStream_t)stream,
                   pdl_level() >= 2, x, gamma, beta, eps, y, y2, y2_dtype, y3, y3_dtype, rows, cols);
  return check_launch(__func__);
}

static int ln_bwd_grid(int rows) {
  int grid = grid_for(rows, kLnBwdWarps * 4);
  if (grid > 592) grid = 592;   // partials workspace is sized for 592 blocks (samk_layernorm_bwd_partials)
  return grid;
}

static int ln_bwd_launch(const float* dy, const float* x, const float* gamma, float eps, float* dx, void* dxd,
                         int dxd_dtype, float drop_p, unsigned long long seed, unsigned long long offset, float* dgamma,
                         float* dbeta, float* dbias, float* partials, int rows, int cols, float* dxd_amax, void* stream,
                         bool finalize, const char* who) {
  if (!(dy && x && gamma && rows >= 0)) { set_error("%s: bad argument", who); return SAMK_ERR_ARG; }
  if (!(cols > 0 && cols % 4 == 0 && cols <= 1024)) { set_error("%s: cols must be a multiple of 4, <= 1024", who); return SAMK_ERR_ARG; }
  if (dxd_amax && cudaMemsetAsync(dxd_amax, 0, sizeof(float), (cudaStream_t)stream) != cudaSuccess) {
    set_error("%s: memset failed", who);
    return SAMK_ERR_CUDA;
  }
  if (!rows) return SAMK_OK;
  const int grid = ln_bwd_grid(rows);
  const int smem = kLnBwdWarps * 3 * cols * (int)sizeof(float);
  launch_maybe_pdl(layernorm_bwd_kernel, dim3(grid), dim3(kLnBwdWarps * 32), (size_t)smem, (cudaStream_t)stream,
                   pdl_level() >= 2, dy, x, gamma, eps, dx, dxd, dxd_dtype,
                   drop_p > 0.f ? drop_threshold(drop_p) : 0u, drop_keep_scale(drop_p), seed, offset, dgamma, dbeta,
                   dbias, partials, rows, cols, reinterpret_cast<unsigned int*>(dxd_amax));
  int rc = check_launch(who);
  if (rc || !partials || !finalize) return rc;
  ln_bwd_finalize_kernel<<<dim3((cols + 31) / 32, 3), 1024, 0, (cudaStream_t)stream>>>(partials, grid, cols, dgamma, dbeta, dbias);
  return check_launch(who);
}

int samk_layernorm_bwd(const float* dy, const float* x, const float* gamma, float eps, float* dx, void* dxd,
                       int dxd_dtype, float drop_p, unsigned long long seed, unsigned long long offset, float* dgamma,
                       float* dbeta, float* dbias, float* partials, int rows, int cols, float* dxd_amax, void* stream) {
  return ln_bwd_launch(dy, x, gamma, eps, dx, dxd, dxd_dtype, drop_p, seed, offset, dgamma, dbeta, dbias, partials, rows, cols,
                       dxd_amax, stream, true, __func__);
}

int samk_layernorm_bwd_main(const float* dy, const float* x, const float* gamma, float eps, float* dx, void* dxd,
                            int dxd_dtype, float drop_p, unsigned long long seed, unsigned long long offset, float* dgamma,
                            float* dbeta, float* dbias, float* partials, int rows, int cols, float* dxd_amax, void* stream) {
  if (!partials) { set_error("%s: needs the partials workspace", __func__); return SAMK_ERR_ARG; }
  return ln_bwd_launch(dy, x, gamma, eps, dx, dxd, dxd_dtype, drop_p, seed, offset, dgamma, dbeta, dbias, partials, rows, cols,
                       dxd_amax, stream, false, __func__);
}

int samk_layernorm_bwd_finalize(const float* partials, int rows, int cols, float* dgamma, float* dbeta, float* dbias,
                                void* stream) {
  SAMK_REQUIRE(partials && rows >= 0 && cols > 0 && cols % 4 == 0 && cols <= 1024, "bad argument");
  if (!rows) return SAMK_OK;
  ln_bwd_finalize_kernel<<<dim3((cols + 31) / 32, 3), 1024, 0, (cudaStream_t)stream>>>(partials, ln_bwd_grid(rows), cols, dgamma,
                                                                                      dbeta, dbias);
  return check_launch(__func__);
}

long long samk_layernorm_bwd_partials(int cols) { return 592LL * 3 * cols; }

int samk_dropout_add(const float* a, const float* b, float* out, void* out2, int out2_dtype, int rows, int cols,
                     float drop_p, unsigned long long seed, unsigned long long offset, void* stream) {
  SAMK_REQUIRE(a && (out || out2) && rows >= 0 && cols >= 0 && cols % 4 == 0, "bad argument");
  if (!rows || !cols) return SAMK_OK;
  dropout_add_kernel<<<grid_for((long long)rows * cols / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      a, b, out, out2, out2_dtype, rows, cols, drop_p > 0.f ? drop_threshold(drop_p) : 0u,
      drop_keep_scale(drop_p), seed, offset);
  return check_launch(__func__);
}

static int colsum_threads() {
  static int t = 0;
  if (!t) {
    const char* e = getenv("SAMK_COLSUM_THREADS");
    const int v = e ? atoi(e) : 256;
    t = (v == 1024 || v == 512) ? v : 256;
  }
  return t;
}

static void launch_colsum(int dt, dim3 grid, int threads, cudaStream_t st, const void* x, long long ld, int rows, int cols,
                          float* o0, float* o1, float* o2, int part_cols) {
  if (dt == SAMK_DT_F32) colsum_kernel<SAMK_DT_F32><<<grid, threads, 0, st>>>(x, ld, rows, cols, o0, o1, o2, part_cols);
  else if (dt == SAMK_DT_F16) colsum_kernel<SAMK_DT_F16><<<grid, threads, 0, st>>>(x, ld, rows, cols, o0, o1, o2, part_cols);
  else colsum_kernel<SAMK_DT_BF16><<<grid, threads, 0, st>>>(x, ld, rows, cols, o0, o1, o2, part_cols);
}

int samk_colsum(const void* x, int x_dtype, long long ld, int rows, int cols, float* out, void* stream) {
  SAMK_REQUIRE(x && out && rows >= 0 && cols >= 0, "bad argument");
  if (!rows || !cols) return SAMK_OK;
  const bool bf = x_dtype != SAMK_DT_F32;
  if (cols % 4 || ld % 4 || ((uintptr_t)x & (bf ? 7 : 15))) {
    colsum_scalar_kernel<<<(cols * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, x_dtype, ld, rows, cols, out);
    return check_launch(__func__);
  }
  const int gx = (cols / 4 + 63) / 64;
  const int threads = colsum_threads(), nrl = threads / 64;
  int gy = (n_sms() * 1024 / threads + gx - 1) / gx;   // 1024 threads per SM in all
  if (gy > (rows + nrl - 1) / nrl) gy = (rows + nrl - 1) / nrl;
  if (gy < 1) gy = 1;
  dim3 grid(gx, gy);
  launch_colsum(x_dtype, grid, threads, (cudaStream_t)stream, x, ld, rows, cols, out, nullptr, nullptr, 0);
  return check_launch(__func__);
}

int samk_colsum3(const void* x, int x_dtype, long long ld, int rows, int part_cols, float* out0, float* out1, float* out2,
                 void* stream) {
  SAMK_REQUIRE(x && out0 && out1 && out2 && rows >= 0 && part_cols > 0, "bad argument");
  const bool bf = x_dtype != SAMK_DT_F32;
  SAMK_REQUIRE(part_cols % 4 == 0 && ld % 4 == 0 && ((uintptr_t)x & (bf ? 7 : 15)) == 0, "needs 4-column aligned parts");
  if (!rows) return SAMK_OK;
  const int cols = 3 * part_cols;
  const int gx = (cols / 4 + 63) / 64;
  const int threads = colsum_threads(), nrl = threads / 64;
  int gy = (n_sms() * 1024 / threads + gx - 1) / gx;   // 1024 threads per SM in all
  if (gy > (rows + nrl - 1) / nrl) gy = (rows + nrl - 1) / nrl;
  if (gy < 1) gy = 1;
  dim3 grid(gx, gy);
  launch_colsum(x_dtype, grid, threads, (cudaStream_t)stream, x, ld, rows, cols, out0, out1, out2, part_cols);
  return check_launch(__func__);
}

int samk_bert_embed_fwd(const long long* ids, const float* word, const float* pos, const float* type,
                        const float* gamma, const float* beta, float eps, float* out, void* out2, int out2_dtype,
                        int rows, int T, int cols, float drop_p, unsigned long long seed, unsigned long long offset,
                        void* stream) {
  SAMK_REQUIRE(ids && word && pos && type && gamma && beta && out, "null pointer");
  SAMK_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 1024 && T > 0, "bad size");
  if (!rows) return SAMK_OK;
  bert_embed_fwd_kernel<<<grid_for(rows, 8), kRowThreads, 0, (cudaStream_t)stream>>>(
      ids, word, pos, type, gamma, beta, eps, out, out2, out2_dtype, rows, T, cols,
      drop_p > 0.f ? drop_threshold(drop_p) : 0u, drop_keep_scale(drop_p), seed, offset);
  return check_launch(__func__);
}

int samk_bert_embed_bwd(const float* dout, const long long* ids, const float* word, const float* pos, const float* type,
                        const float* gamma, float eps, float* dword, float* dpos, float* dtype, float* dgamma,
                        float* dbeta, int rows, int T, int cols, float drop_p, unsigned long long seed,
                        unsigned long long offset, void* stream) {
  SAMK_REQUIRE(dout && ids && word && pos && type && gamma && dword && dpos && dtype && dgamma && dbeta, "null pointer");
  SAMK_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 1024 && T > 0, "bad size");
  if (!rows) return SAMK_OK;
  // few blocks: each reduces its rows' hot-column gradients in shared memory before touching global memory
  bert_embed_bwd_kernel<<<grid_for(rows, 32) < n_sms() ? grid_for(rows, 32) : n_sms(), kRowThreads, 0, (cudaStream_t)stream>>>(
      dout, ids, word, pos, type, gamma, eps, dword, dpos, dtype, dgamma, dbeta, rows, T, cols,
      drop_p > 0.f ? drop_threshold(drop_p) : 0u, drop_keep_scale(drop_p), seed, offset);
  return check_launch(__func__);
}

static int fill_prevpred(PrevPredParams& p, const long long* prev, const float* cls_w, const float* ocr_in,
                         const float* pos, const float* type, const float* const* ln, float eps, int B, int D, int V,
                         int R, int cols, float drop_p, unsigned long long seed, unsigned long long offset) {
  p.prev = prev; p.cls_w = cls_w; p.ocr_in = ocr_in; p.pos = pos; p.type = type;
  p.ans_g = ln[0]; p.ans_b = ln[1]; p.ocr_g = ln[2]; p.ocr_b = ln[3]; p.emb_g = ln[4]; p.emb_b = ln[5];
  p.eps = eps; p.B = B; p.D = D; p.V = V; p.R = R; p.cols = cols;
  p.thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  p.scale = drop_keep_scale(drop_p);
  p.seed = seed; p.off = offset;
  return 0;
}

int samk_prevpred_fwd(const long long* prev, const float* cls_w, const float* ocr_in, const float* pos,
                      const float* type, const float* const* ln6, float eps, float* out, int B, int D, int V, int R,
                      int cols, float drop_p, unsigned long long seed, unsigned long long offset, void* stream) {
  SAMK_REQUIRE(prev && cls_w && ocr_in && pos && type && ln6 && out, "null pointer");
  SAMK_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 1024, "bad size");
  if (!B || !D) return SAMK_OK;
  PrevPredParams p;
  fill_prevpred(p, prev, cls_w, ocr_in, pos, type, ln6, eps, B, D, V, R, cols, drop_p, seed, offset);
  prevpred_fwd_kernel<<<grid_for((long long)B * D, 8), kRowThreads, 0, (cudaStream_t)stream>>>(p, out);
  return check_launch(__func__);
}

int samk_prevpred_bwd(const float* dout, const long long* prev, const float* cls_w, const float* ocr_in,
                      const float* pos, const float* type, const float* const* ln6, float eps, float* const* grads10,
                      int B, int D, int V, int R, int cols, float drop_p, unsigned long long seed,
                      unsigned long long offset, void* stream) {
  SAMK_REQUIRE(dout && prev && cls_w && ocr_in && pos && type && ln6 && grads10, "null pointer");
  SAMK_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 1024, "bad size");
  if (!B || !D) return SAMK_OK;
  PrevPredParams p;
  fill_prevpred(p, prev, cls_w, ocr_in, pos, type, ln6, eps, B, D, V, R, cols, drop_p, seed, offset);
  PrevPredGrads g;
  g.d_cls_w = grads10[0]; g.d_ocr_in = grads10[1]; g.d_pos = grads10[2]; g.d_type = grads10[3];
  g.d_ans_g = grads10[4]; g.d_ans_b = grads10[5]; g.d_ocr_g = grads10[6]; g.d_ocr_b = grads10[7];
  g.d_emb_g = grads10[8]; g.d_emb_b = grads10[9];
  const int grid = grid_for((long long)B * D, 16) < n_sms() ? grid_for((long long)B * D, 16) : n_sms();
  prevpred_bwd_kernel<<<grid, kRowThreads, 8 * cols * sizeof(float), (cudaStream_t)stream>>>(p, dout, g);
  return check_launch(__func__);
}

int samk_ptr_scores_fwd(const float* q, const float* k, const long long* ocr_mask, float* out, long long ldo,
                        int col_off, int B, int D, int R, int dq, void* stream) {
  SAMK_REQUIRE(q && k && ocr_mask && out && dq % 4 == 0, "bad argument");
  long long warps = (long long)B * D * R;
  if (!warps) return SAMK_OK;
  ptr_scores_fwd_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(q, k, ocr_mask, out, ldo, col_off, B, D, R, dq, 1.0f / sqrtf((float)dq));
  return check_launch(__func__);
}

int samk_ptr_scores_bwd(const float* dscores, long long ldds, int col_off, const float* q, const float* k, float* dq_,
                        float* dk_, int B, int D, int R, int dq, void* stream) {
  SAMK_REQUIRE(dscores && q && k && dq_ && dk_ && dq % 4 == 0 && dq <= 1024, "bad argument (ptr_query_size <= 1024)");
  long long warps = (long long)B * (D + R);
  if (!warps) return SAMK_OK;
  ptr_scores_bwd_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dscores, ldds, col_off, q, k, dq_, dk_, B, D, R, dq, 1.0f / sqrtf((float)dq));
  return check_launch(__func__);
}

int samk_bce_loss(const float* scores, const float* targets, const float* loss_mask, float* dscores, float* loss_out,
                  float* scratch, int rows, int ncls, void* stream) {
  SAMK_REQUIRE(scores && targets && loss_mask && loss_out && scratch, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(loss_out, 0, sizeof(float), s);
  cudaMemsetAsync(scratch, 0, sizeof(float), s);
  if (!rows || !ncls) return SAMK_OK;
  sum_kernel<<<grid_for(rows, 256), 256, 0, s>>>(loss_mask, rows, scratch);
  bce_loss_kernel<<<grid_for((long long)rows * ncls, 1024), 256, 0, s>>>(scores, targets, loss_mask, dscores, loss_out, scratch, (long long)rows * ncls, ncls);
  return check_launch(__func__);
}

int samk_sumsq(const float* x, long long n, double* out_accum, void* stream) {
  SAMK_REQUIRE(x && out_accum && n >= 0, "bad argument");
  if (!n) return SAMK_OK;
  sumsq_kernel<<<grid_for(n, 2048), 256, 0, (cudaStream_t)stream>>>(x, n, out_accum);
  return check_launch(__func__);
}

int samk_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1,
                   double beta2, double eps, int step, const double* grad_sumsq, double max_norm, void* stream) {
  SAMK_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "bad argument");
  if (!n) return SAMK_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  adam_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(-lr / bc1),
                                                                  (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                                  (float)eps, (float)sqrt(bc2), grad_sumsq, (float)max_norm);
  return check_launch(__func__);
}

int samk_cast_scaled_f16(const void* x, int x_dtype, long long n, const float* amax, void* y, float* scale2, void* stream) {
  SAMK_REQUIRE(x && y && scale2 && n >= 0, "bad argument");
  SAMK_REQUIRE(n % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 7) == 0, "n must be a multiple of 4, aligned buffers");
  cudaStream_t st = (cudaStream_t)stream;
  // scale2[0] doubles as the amax accumulator of pass 1 (bit pattern), overwritten with S by pass 2 ... no: a reader
  // of pass 2 in another block could see S instead of amax, so the accumulator is the third float of the workspace
  const int grid_c = grid_for(n / 4, 2048) > 1184 ? 1184 : grid_for(n / 4, 2048);
  if (amax) {      // max|x| already known (e.g. accumulated by samk_layernorm_bwd while it wrote x): one pass
    cast_scaled_f16_kernel<<<n ? grid_c : 1, n ? 256 : 32, 0, st>>>(x, x_dtype, n / 4, reinterpret_cast<const unsigned int*>(amax), (__half*)y, scale2);
    return check_launch(__func__);
  }
  if (cudaMemsetAsync(scale2 + 2, 0, sizeof(float), st) != cudaSuccess) { set_error("%s: memset failed", __func__); return SAMK_ERR_CUDA; }
  if (!n) { cast_scaled_f16_kernel<<<1, 32, 0, st>>>(x, x_dtype, 0, reinterpret_cast<unsigned int*>(scale2 + 2), (__half*)y, scale2); return check_launch(__func__); }
  const int grid = grid_for(n / 4, 2048) > 1184 ? 1184 : grid_for(n / 4, 2048);
  amax_kernel<<<grid, 256, 0, st>>>(x, x_dtype, n / 4, reinterpret_cast<unsigned int*>(scale2 + 2));
  cast_scaled_f16_kernel<<<grid, 256, 0, st>>>(x, x_dtype, n / 4, reinterpret_cast<const unsigned int*>(scale2 + 2), (__half*)y, scale2);
  return check_launch(__func__);
}

int samk_cast_flat(const void* x, int x_dtype, void* y, int y_dtype, long long n, void* stream) {
  SAMK_REQUIRE(x && y && n >= 0, "bad argument");
  SAMK_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "buffers must be 16-byte aligned");
  if (!n) return SAMK_OK;
  cast_flat_kernel<<<grid_for(n / 4 + 1, 1024), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, y, y_dtype, n);
  return check_launch(__func__);
}

int samk_row_segments_f32(float* joint, float* const* segs, const int* rows, const int* offs, int n_seg, int B, int L, int d,
                          int to_joint, void* stream) {
  SAMK_REQUIRE(joint && segs && rows && offs && n_seg >= 1 && n_seg <= 4 && B >= 0 && L >= 0, "bad argument");
  SAMK_REQUIRE(d > 0 && d % 4 == 0 && al16(joint), "d must be a multiple of 4, 16-byte aligned buffers");
  RowSegments sg;
  long long total = 0;
  for (int k = 0; k < 4; ++k) {
    sg.ptr[k] = k < n_seg ? segs[k] : nullptr; sg.rows[k] = k < n_seg ? rows[k] : 0; sg.off[k] = k < n_seg ? offs[k] : 0;
    if (k < n_seg) {
      SAMK_REQUIRE(rows[k] >= 0 && offs[k] >= 0 && offs[k] + rows[k] <= L && (rows[k] == 0 || (segs[k] && al16(segs[k]))), "bad segment");
      total += rows[k];
    }
  }
  sg.n = n_seg;
  if (!B || !total) return SAMK_OK;
  row_segments_kernel<<<grid_for((long long)B * total * (d / 4), 1024), 256, 0, (cudaStream_t)stream>>>(joint, sg, B, L, d / 4, to_joint);
  return check_launch(__func__);
}

int samk_key_valid(const long long* q_mask, const long long* obj_mask, const long long* ocr_mask, uint8_t* out, int B, int T,
                   int O, int R, int D, void* stream) {
  SAMK_REQUIRE(out && B >= 0 && T >= 0 && O >= 0 && R >= 0 && D >= 0, "bad argument");
  SAMK_REQUIRE((T == 0 || q_mask) && (O == 0 || obj_mask) && (R == 0 || ocr_mask), "null mask");
  const long long n = (long long)B * (T + O + R + D);
  if (!n) return SAMK_OK;
  key_valid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q_mask, obj_mask, ocr_mask, out, B, T, O, R, D);
  return check_launch(__func__);
}

int samk_memset0(void* p, long long bytes, void* stream) {
  SAMK_REQUIRE(p && bytes >= 0, "bad argument");
  if (!bytes) return SAMK_OK;
  if (cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream) != cudaSuccess) {
    set_error("%s: %s", __func__, cudaGetErrorString(cudaGetLastError()));
    return SAMK_ERR_CUDA;
  }
  return SAMK_OK;
}

int samk_beam_step(const float* scores, long long row_stride, int ncls, const float* beam_scores, const uint8_t* completed,
                   int eos, int first_step, int B, int K, long long* prev_pos, long long* new_pos, float* new_scores, void* stream) {
  SAMK_REQUIRE(scores && beam_scores && prev_pos && new_pos && new_scores && B >= 0 && ncls > 0, "bad argument");
  SAMK_REQUIRE(K >= 1 && K <= kBeamMax, "beam size must be in 1..16");
  SAMK_REQUIRE((long long)K * ncls < 0x7fffffffLL, "too many candidates per sample");
  if (!B) return SAMK_OK;
  beam_step_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(scores, row_stride, ncls, beam_scores, completed, eos, first_step, K,
                                                       prev_pos, new_pos, new_scores);
  return check_launch(__func__);
}

int samk_argmax_rows(const float* x, long long ld, long long rows, int ncls, const float* targets, long long ldt, long long* idx_out,
                     float* hit_out, void* stream) {
  SAMK_REQUIRE(x && idx_out && rows >= 0 && ncls > 0 && rows < 0x7fffffffLL, "bad argument");
  if (!rows) return SAMK_OK;
  argmax_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, ld, ncls, targets, ldt, idx_out, hit_out);
  return check_launch(__func__);
}

int samk_concat3_f32(const float* a, const float* b, const float* c, float* out, int n, void* stream) {
  SAMK_REQUIRE(a && b && c && out && n >= 0, "bad argument");
  if (!n) return SAMK_OK;
  concat3_kernel<<<(3 * n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, b, c, out, n);
  return check_launch(__func__);
}

int samk_scale_inplace(float* x, long long n, const float* scale_dev, void* stream) {
  SAMK_REQUIRE(x && scale_dev, "null pointer");
  if (!n) return SAMK_OK;
  scale_kernel<<<grid_for(n, 1024), 256, 0, (cudaStream_t)stream>>>(x, n, scale_dev);
  return check_launch(__func__);
}

}  // extern "C"

namespace samk { int set_drop_salt_elementwise(unsigned long long salt, cudaStream_t stream) { return set_drop_salt_tu(salt, stream); } }
namespace samk { int set_drop_salt_dev_elementwise(const unsigned long long* src, cudaStream_t stream) { return set_drop_salt_from_device_tu(src, stream); } }
