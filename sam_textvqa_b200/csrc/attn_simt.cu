// Exact-fp32 SIMT masked multi-head attention, forward and backward.
//
// Restates SpatialBertSelfAttention.forward steps (3)-(7) (/root/reference/sam/sa_m4c.py:562-598)
// and the plain BertSelfAttention of the 'n' layers / TextBert in one kernel family, with the
// masks of attn_mask.cuh evaluated on the fly (no [B,L,L,H] tensors).  All arithmetic is fp32
// on the CUDA cores: this is the PARITY-MODE attention (logits within 1e-3 of the fp32
// reference) and the on-device cross-check for the tensor-core attention kernel.
//
// Layout: qkv [B, L, 3*H*dh] (q | k | v, head-major inside each third), ctx [B, L, H*dh],
// lse [B, H, L] fp32 = log-sum-exp of the allowed scaled scores (+inf marks a dead row).
// dh = 64.  Forward / dQ kernels: block = 32 query rows (8 warps x 4 rows) of one (b,h), keys
// streamed through shared memory in tiles of 64 with an online softmax.  dK/dV kernel: block =
// 32 keys (8 warps x 4 keys), query rows streamed in tiles of 64.
#include "common.cuh"
#include "attn_mask.cuh"
#include "../../include/samk.h"

namespace samk {

constexpr int DH = 64;
constexpr int KT = 64;          // streamed tile (keys in fwd/dQ, rows in dKV)
constexpr int OWN = 32;         // rows (or keys) owned by a block
constexpr int kAttnThreads = 256;

struct AttnArgs {
  const void* qkv; void* ctx; float* lse;
  const void* dctx; void* dqkv; float* delta;
  int in_bf16, out_bf16;   // SAMK_DT_* codes: in = qkv / ctx (activations), out = dctx / dqkv (gradients)
  int B, H;
  float scale;
  uint32_t drop_thresh; float drop_scale; unsigned long long seed, off;
  AttnMask m;
  int q_blk0;   // forward: first 32-row block to compute
};

__device__ __forceinline__ float ld_elem(const void* p, int bf16, size_t i) {
  return bf16 ? ld_16(reinterpret_cast<const uint16_t*>(p) + i, bf16 == SAMK_DT_F16) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_elem(void* p, int bf16, size_t i, float v) {
  if (bf16) st_16(reinterpret_cast<uint16_t*>(p) + i, v, bf16 == SAMK_DT_F16);
  else reinterpret_cast<float*>(p)[i] = v;
}

// keep-multiplier of attention-probability element (b,h,i,j): 0 or 1/(1-p)
__device__ __forceinline__ float attn_keep(const AttnArgs& a, int b, int h, int i, int j) {
  if (!a.drop_thresh) return 1.0f;
  const uint64_t row = ((uint64_t)(b * a.H + h) * a.m.L + i) * (uint64_t)((a.m.L + 7) >> 3);   // groups of 8 keys
  return dropout_keep1(a.seed, a.off, row * 8 + (uint64_t)j, a.drop_thresh) ? a.drop_scale : 0.0f;
}

// load rows [r0, r0+nrows) x 64 of one head slice into smem[nrows][65] (zero rows past L)
__device__ __forceinline__ void load_tile(float (*dst)[DH + 1], const void* src, int bf16, size_t base, long long ld,
                                          int r0, int nrows, int L) {
  for (int idx = threadIdx.x; idx < nrows * (DH / 4); idx += blockDim.x) {
    const int r = idx / (DH / 4), c = (idx % (DH / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < L) {
      const size_t o = base + (size_t)(r0 + r) * ld + c;
      if (bf16) {
        uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(src) + o);
        float2 x = unpack_16(u.x, bf16 == SAMK_DT_F16), y = unpack_16(u.y, bf16 == SAMK_DT_F16);
        v = make_float4(x.x, x.y, y.x, y.y);
      } else {
        v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + o);
      }
    }
    dst[r][c] = v.x; dst[r][c + 1] = v.y; dst[r][c + 2] = v.z; dst[r][c + 3] = v.w;
  }
}

// ---------------------------------------------------------------------------------------------
// forward (kBwd = false) and dQ (kBwd = true) share the row-parallel structure
// ---------------------------------------------------------------------------------------------
template <bool kBwd>
__global__ void __launch_bounds__(kAttnThreads)
attn_rows_kernel(const AttnArgs a) {
  extern __shared__ float4 dyn_smem[];
  float4 (*Q4)[DH] = reinterpret_cast<float4 (*)[DH]>(dyn_smem);            // per warp: q of its 4 rows, by d
  float4 (*G4)[DH] = Q4 + 8;                                                // bwd: dO of its 4 rows
  float4 (*P4)[KT] = reinterpret_cast<float4 (*)[KT]>(G4 + 8);              // p (fwd) / ds (bwd), 4 rows x key tile
  float (*Ks)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(P4 + 8);
  float (*Vs)[DH + 1] = Ks + KT;
  __shared__ int sflag;

  const AttnMask& m = a.m;
  const int L = m.L;
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x + (kBwd ? 0 : a.q_blk0)) * OWN + warp * 4;
  const long long ld = 3LL * a.H * DH;
  const size_t qbase = (size_t)b * L * ld + (size_t)h * DH;
  const size_t kbase = qbase + (size_t)a.H * DH, vbase = kbase + (size_t)a.H * DH;
  const long long ldc = (long long)a.H * DH;
  const size_t cbase = (size_t)b * L * ldc + (size_t)h * DH;

  const bool any_valid = sample_any_valid(m, b, &sflag);

  // stage q (and dO) of this warp's rows, d-major float4 over the 4 rows
  for (int d = lane; d < DH; d += 32) {
    float q[4], g[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = row0 + r;
      q[r] = i < L ? ld_elem(a.qkv, a.in_bf16, qbase + (size_t)i * ld + d) : 0.f;
      g[r] = (kBwd && i < L) ? ld_elem(a.dctx, a.out_bf16, cbase + (size_t)i * ldc + d) : 0.f;
    }
    Q4[warp][d] = make_float4(q[0], q[1], q[2], q[3]);
    if (kBwd) G4[warp][d] = make_float4(g[0], g[1], g[2], g[3]);
  }

  float mrow[4], lrow[4], acc[4][2], lse_r[4], delta_r[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { mrow[r] = -INFINITY; lrow[r] = 0.f; acc[r][0] = acc[r][1] = 0.f; lse_r[r] = 0.f; delta_r[r] = 0.f; }
  if (kBwd) {
    // delta_i = dO_i . O_i ; lse_i from the forward pass
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = row0 + r;
      float t = 0.f;
      if (i < L) {
        for (int d = lane; d < DH; d += 32)
          t += ld_elem(a.dctx, a.out_bf16, cbase + (size_t)i * ldc + d) * ld_elem(a.ctx, a.in_bf16, cbase + (size_t)i * ldc + d);
      }
      delta_r[r] = warp_sum(t);
      lse_r[r] = i < L ? a.lse[((size_t)b * a.H + h) * L + i] : INFINITY;
      if (i < L && lane == 0) a.delta[((size_t)b * a.H + h) * L + i] = delta_r[r];
    }
  }

  for (int k0 = 0; k0 < L; k0 += KT) {
    __syncthreads();
    load_tile(Ks, a.qkv, a.in_bf16, kbase, ld, k0, KT, L);
    load_tile(Vs, a.qkv, a.in_bf16, vbase, ld, k0, KT, L);
    __syncthreads();
    // scores of 4 rows x 2 keys per lane (+ dO.V in the backward)
    float s[4][2], dp[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) { s[r][0] = s[r][1] = 0.f; dp[r][0] = dp[r][1] = 0.f; }
#pragma unroll 8
    for (int d = 0; d < DH; ++d) {
      const float4 q = Q4[warp][d];
      const float ka = Ks[lane][d], kb = Ks[lane + 32][d];
      s[0][0] = fmaf(q.x, ka, s[0][0]); s[0][1] = fmaf(q.x, kb, s[0][1]);
      s[1][0] = fmaf(q.y, ka, s[1][0]); s[1][1] = fmaf(q.y, kb, s[1][1]);
      s[2][0] = fmaf(q.z, ka, s[2][0]); s[2][1] = fmaf(q.z, kb, s[2][1]);
      s[3][0] = fmaf(q.w, ka, s[3][0]); s[3][1] = fmaf(q.w, kb, s[3][1]);
      if (kBwd) {
        const float4 g = G4[warp][d];
        const float va = Vs[lane][d], vb = Vs[lane + 32][d];
        dp[0][0] = fmaf(g.x, va, dp[0][0]); dp[0][1] = fmaf(g.x, vb, dp[0][1]);
        dp[1][0] = fmaf(g.y, va, dp[1][0]); dp[1][1] = fmaf(g.y, vb, dp[1][1]);
        dp[2][0] = fmaf(g.z, va, dp[2][0]); dp[2][1] = fmaf(g.z, vb, dp[2][1]);
        dp[3][0] = fmaf(g.w, va, dp[3][0]); dp[3][1] = fmaf(g.w, vb, dp[3][1]);
      }
    }
    float p[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = row0 + r;
      bool ok[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = k0 + lane + 32 * c;
        ok[c] = i < L && attn_allowed(m, b, h, i, j, any_valid);
        s[r][c] = ok[c] ? s[r][c] * a.scale : -INFINITY;
      }
      if (!kBwd) {
        const float tmax = warp_max(fmaxf(s[r][0], s[r][1]));
        const float mnew = fmaxf(mrow[r], tmax);
        const float corr = (mnew == -INFINITY) ? 1.f : expf(mrow[r] - mnew);
        float psum = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float e = ok[c] ? expf(s[r][c] - mnew) : 0.f;
          psum += e;
          p[r][c] = e * attn_keep(a, b, h, i, k0 + lane + 32 * c);
        }
        psum = warp_sum(psum);
        lrow[r] = lrow[r] * corr + psum;
        acc[r][0] *= corr; acc[r][1] *= corr;
        mrow[r] = mnew;
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float pr = ok[c] ? expf(s[r][c] - lse_r[r]) : 0.f;
          const float keep = attn_keep(a, b, h, i, k0 + lane + 32 * c);
          p[r][c] = pr * (dp[r][c] * keep - delta_r[r]) * a.scale;   // ds * scale
        }
      }
    }
    P4[warp][lane] = make_float4(p[0][0], p[1][0], p[2][0], p[3][0]);
    P4[warp][lane + 32] = make_float4(p[0][1], p[1][1], p[2][1], p[3][1]);
    __syncwarp();
    // fwd: acc += P V ; bwd: acc += dS K     (lane owns d = lane, lane+32)
    float (*Ms)[DH + 1] = kBwd ? Ks : Vs;
#pragma unroll 8
    for (int j = 0; j < KT; ++j) {
      const float4 w = P4[warp][j];
      const float x0 = Ms[j][lane], x1 = Ms[j][lane + 32];
      acc[0][0] = fmaf(w.x, x0, acc[0][0]); acc[0][1] = fmaf(w.x, x1, acc[0][1]);
      acc[1][0] = fmaf(w.y, x0, acc[1][0]); acc[1][1] = fmaf(w.y, x1, acc[1][1]);
      acc[2][0] = fmaf(w.z, x0, acc[2][0]); acc[2][1] = fmaf(w.z, x1, acc[2][1]);
      acc[3][0] = fmaf(w.w, x0, acc[3][0]); acc[3][1] = fmaf(w.w, x1, acc[3][1]);
    }
    __syncwarp();
  }

#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = row0 + r;
    if (i >= L) continue;
    if (!kBwd) {
      const float inv = lrow[r] > 0.f ? 1.0f / lrow[r] : 0.f;
      st_elem(a.ctx, a.in_bf16, cbase + (size_t)i * ldc + lane, acc[r][0] * inv);
      st_elem(a.ctx, a.in_bf16, cbase + (size_t)i * ldc + lane + 32, acc[r][1] * inv);
      if (lane == 0) a.lse[((size_t)b * a.H + h) * L + i] = lrow[r] > 0.f ? mrow[r] + logf(lrow[r]) : INFINITY;
    } else {
      st_elem(a.dqkv, a.out_bf16, qbase + (size_t)i * ld + lane, acc[r][0]);
      st_elem(a.dqkv, a.out_bf16, qbase + (size_t)i * ld + lane + 32, acc[r][1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dK / dV: block owns 32 keys of one (b,h); query rows streamed in tiles of 64
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads)
attn_dkv_kernel(const AttnArgs a) {
  extern __shared__ float4 dyn_smem[];
  float4 (*K4)[DH] = reinterpret_cast<float4 (*)[DH]>(dyn_smem);            // per warp: its 4 keys, d-major
  float4 (*V4)[DH] = K4 + 8;
  float4 (*W4)[KT] = reinterpret_cast<float4 (*)[KT]>(V4 + 8);              // p_drop of 4 keys x row tile
  float4 (*S4)[KT] = W4 + 8;                                                // ds*scale
  float (*Qs)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(S4 + 8);
  float (*Gs)[DH + 1] = Qs + KT;                                            // dO rows
  float* lse_s = reinterpret_cast<float*>(Gs + KT);
  float* delta_s = lse_s + KT;
  __shared__ int sflag;

  const AttnMask& m = a.m;
  const int L = m.L;
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int key0 = blockIdx.x * OWN + warp * 4;
  const long long ld = 3LL * a.H * DH;
  const size_t qbase = (size_t)b * L * ld + (size_t)h * DH;
  const size_t kbase = qbase + (size_t)a.H * DH, vbase = kbase + (size_t)a.H * DH;
  const long long ldc = (long long)a.H * DH;
  const size_t cbase = (size_t)b * L * ldc + (size_t)h * DH;

  const bool any_valid = sample_any_valid(m, b, &sflag);

  for (int d = lane; d < DH; d += 32) {
    float k[4], v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = key0 + r;
      k[r] = j < L ? ld_elem(a.qkv, a.in_bf16, kbase + (size_t)j * ld + d) : 0.f;
      v[r] = j < L ? ld_elem(a.qkv, a.in_bf16, vbase + (size_t)j * ld + d) : 0.f;
    }
    K4[warp][d] = make_float4(k[0], k[1], k[2], k[3]);
    V4[warp][d] = make_float4(v[0], v[1], v[2], v[3]);
  }
  float dk[4][2], dv[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) { dk[r][0] = dk[r][1] = dv[r][0] = dv[r][1] = 0.f; }

  for (int i0 = 0; i0 < L; i0 += KT) {
    __syncthreads();
    load_tile(Qs, a.qkv, a.in_bf16, qbase, ld, i0, KT, L);
    load_tile(Gs, a.dctx, a.out_bf16, cbase, ldc, i0, KT, L);
    for (int i = threadIdx.x; i < KT; i += blockDim.x) {
      lse_s[i] = (i0 + i < L) ? a.lse[((size_t)b * a.H + h) * L + i0 + i] : INFINITY;
      delta_s[i] = (i0 + i < L) ? a.delta[((size_t)b * a.H + h) * L + i0 + i] : 0.f;
    }
    __syncthreads();
    float s[4][2], dp[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) { s[r][0] = s[r][1] = dp[r][0] = dp[r][1] = 0.f; }
#pragma unroll 8
    for (int d = 0; d < DH; ++d) {
      const float4 k = K4[warp][d], v = V4[warp][d];
      const float qa = Qs[lane][d], qb = Qs[lane + 32][d];
      const float ga = Gs[lane][d], gb = Gs[lane + 32][d];
      s[0][0] = fmaf(k.x, qa, s[0][0]); s[0][1] = fmaf(k.x, qb, s[0][1]);
      s[1][0] = fmaf(k.y, qa, s[1][0]); s[1][1] = fmaf(k.y, qb, s[1][1]);
      s[2][0] = fmaf(k.z, qa, s[2][0]); s[2][1] = fmaf(k.z, qb, s[2][1]);
      s[3][0] = fmaf(k.w, qa, s[3][0]); s[3][1] = fmaf(k.w, qb, s[3][1]);
      dp[0][0] = fmaf(v.x, ga, dp[0][0]); dp[0][1] = fmaf(v.x, gb, dp[0][1]);
      dp[1][0] = fmaf(v.y, ga, dp[1][0]); dp[1][1] = fmaf(v.y, gb, dp[1][1]);
      dp[2][0] = fmaf(v.z, ga, dp[2][0]); dp[2][1] = fmaf(v.z, gb, dp[2][1]);
      dp[3][0] = fmaf(v.w, ga, dp[3][0]); dp[3][1] = fmaf(v.w, gb, dp[3][1]);
    }
    float w[4][2], t[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = key0 + r;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int il = lane + 32 * c, i = i0 + il;
        const bool ok = i < L && j < L && attn_allowed(m, b, h, i, j, any_valid);
        const float pr = ok ? expf(s[r][c] * a.scale - lse_s[il]) : 0.f;
        const float keep = ok ? attn_keep(a, b, h, i, j) : 0.f;
        w[r][c] = pr * keep;
        t[r][c] = pr * (dp[r][c] * keep - delta_s[il]) * a.scale;
      }
    }
    W4[warp][lane] = make_float4(w[0][0], w[1][0], w[2][0], w[3][0]);
    W4[warp][lane + 32] = make_float4(w[0][1], w[1][1], w[2][1], w[3][1]);
    S4[warp][lane] = make_float4(t[0][0], t[1][0], t[2][0], t[3][0]);
    S4[warp][lane + 32] = make_float4(t[0][1], t[1][1], t[2][1], t[3][1]);
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < KT; ++i) {
      const float4 pw = W4[warp][i], ds = S4[warp][i];
      const float g0 = Gs[i][lane], g1 = Gs[i][lane + 32];
      const float q0 = Qs[i][lane], q1 = Qs[i][lane + 32];
      dv[0][0] = fmaf(pw.x, g0, dv[0][0]); dv[0][1] = fmaf(pw.x, g1, dv[0][1]);
      dv[1][0] = fmaf(pw.y, g0, dv[1][0]); dv[1][1] = fmaf(pw.y, g1, dv[1][1]);
      dv[2][0] = fmaf(pw.z, g0, dv[2][0]); dv[2][1] = fmaf(pw.z, g1, dv[2][1]);
      dv[3][0] = fmaf(pw.w, g0, dv[3][0]); dv[3][1] = fmaf(pw.w, g1, dv[3][1]);
      dk[0][0] = fmaf(ds.x, q0, dk[0][0]); dk[0][1] = fmaf(ds.x, q1, dk[0][1]);
      dk[1][0] = fmaf(ds.y, q0, dk[1][0]); dk[1][1] = fmaf(ds.y, q1, dk[1][1]);
      dk[2][0] = fmaf(ds.z, q0, dk[2][0]); dk[2][1] = fmaf(ds.z, q1, dk[2][1]);
      dk[3][0] = fmaf(ds.w, q0, dk[3][0]); dk[3][1] = fmaf(ds.w, q1, dk[3][1]);
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = key0 + r;
    if (j >= L) continue;
    st_elem(a.dqkv, a.out_bf16, kbase + (size_t)j * ld + lane, dk[r][0]);
    st_elem(a.dqkv, a.out_bf16, kbase + (size_t)j * ld + lane + 32, dk[r][1]);
    st_elem(a.dqkv, a.out_bf16, vbase + (size_t)j * ld + lane, dv[r][0]);
    st_elem(a.dqkv, a.out_bf16, vbase + (size_t)j * ld + lane + 32, dv[r][1]);
  }
}

constexpr int kRowsSmem = (8 * DH * 2 + 8 * KT) * 16 + 2 * KT * (DH + 1) * 4;
constexpr int kDkvSmem = (8 * DH * 2 + 8 * KT * 2) * 16 + 2 * KT * (DH + 1) * 4 + 2 * KT * 4;

template <class K>
static int ensure_smem(K kern, int bytes) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%d) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return SAMK_ERR_CUDA;
  }
  return SAMK_OK;
}

static int fill_args(AttnArgs& a, const samk_attn_params* p) {
  if (!p || !p->qkv || !p->key_valid) { set_error("samk_attn: null pointer"); return SAMK_ERR_ARG; }
  if (p->head_dim != DH) { set_error("samk_attn: head_dim must be 64"); return SAMK_ERR_UNSUPPORTED; }
  if (p->T < 0 || p->A < 0 || p->D < 0 || p->B < 0 || p->H <= 0 || p->H > 16) { set_error("samk_attn: bad sizes"); return SAMK_ERR_ARG; }
  if (p->spatial && p->A > 0 && !p->rel_bits) { set_error("samk_attn: spatial layer needs rel_bits"); return SAMK_ERR_ARG; }
  a.qkv = p->qkv; a.ctx = p->ctx; a.lse = p->lse; a.dctx = p->dctx; a.dqkv = p->dqkv; a.delta = p->delta;
  a.in_bf16 = p->dtype; a.out_bf16 = p->grad_dtype;
  a.B = p->B; a.H = p->H; a.scale = p->scale;
  a.drop_thresh = p->drop_p > 0.f ? drop_threshold(p->drop_p) : 0u;
  a.drop_scale = drop_keep_scale(p->drop_p);
  a.seed = p->drop_seed; a.off = p->drop_offset;
  a.m.valid = p->key_valid; a.m.rel = p->spatial ? p->rel_bits : nullptr;
  a.m.T = p->T; a.m.A = p->A; a.m.D = p->D; a.m.L = p->T + p->A + p->D;
  a.m.quad_mask = p->spatial ? p->quadrant_mask : 0u; a.m.spatial = p->spatial ? 1 : 0;
  a.q_blk0 = p->q_begin > 0 ? p->q_begin / OWN : 0;
  return SAMK_OK;
}

int attn_simt_fwd(const samk_attn_params* p, cudaStream_t stream) {
  AttnArgs a;
  int rc = fill_args(a, p);
  if (rc) return rc;
  if (!p->ctx || !p->lse) { set_error("samk_attn_fwd: ctx/lse null"); return SAMK_ERR_ARG; }
  if (!a.B || !a.m.L) return SAMK_OK;
  dim3 grid((a.m.L + OWN - 1) / OWN - a.q_blk0, a.H, a.B);
  if ((rc = ensure_smem(attn_rows_kernel<false>, kRowsSmem))) return rc;
  attn_rows_kernel<false><<<grid, kAttnThreads, kRowsSmem, stream>>>(a);
  return check_launch("samk_attn_fwd(simt)");
}

int attn_simt_bwd(const samk_attn_params* p, cudaStream_t stream) {
  AttnArgs a;
  int rc = fill_args(a, p);
  if (rc) return rc;
  if (!p->ctx || !p->lse || !p->dctx || !p->dqkv || !p->delta) { set_error("samk_attn_bwd: null pointer"); return SAMK_ERR_ARG; }
  if (!a.B || !a.m.L) return SAMK_OK;
  if (p->bwd_phase == 2) return SAMK_OK;        // this path has no separate preparation
  dim3 grid((a.m.L + OWN - 1) / OWN, a.H, a.B);
  if ((rc = ensure_smem(attn_rows_kernel<true>, kRowsSmem))) return rc;
  if ((rc = ensure_smem(attn_dkv_kernel, kDkvSmem))) return rc;
  attn_rows_kernel<true><<<grid, kAttnThreads, kRowsSmem, stream>>>(a);
  rc = check_launch("samk_attn_bwd(simt,dq)");
  if (rc) return rc;
  attn_dkv_kernel<<<grid, kAttnThreads, kDkvSmem, stream>>>(a);
  return check_launch("samk_attn_bwd(simt,dkv)");
}

}  // namespace samk

namespace samk { int set_drop_salt_attn_simt(unsigned long long salt, cudaStream_t stream) { return set_drop_salt_tu(salt, stream); } }
namespace samk { int set_drop_salt_dev_attn_simt(const unsigned long long* src, cudaStream_t stream) { return set_drop_salt_from_device_tu(src, stream); } }
