// Fused masked multi-head attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), bf16 I/O.
//
// Replaces SpatialBertSelfAttention.forward steps (3)-(7) (/root/reference/sam/sa_m4c.py:562-598) and
// the plain BertSelfAttention of the 'n' layers / TextBert, forward and backward.
//
//   forward  CTA = (128 query rows, head, sample).  TMA stages Q [128x64], K,V [KVTx64] (128B swizzle);
//            S = Q K^T by tcgen05.mma into TMEM; 128 threads (one per row) read S with tcgen05.ld,
//            apply the packed allow-bits, exp2 softmax (online across key tiles), Philox dropout, and
//            write P (bf16) into a swizzled smem tile; O += P V by tcgen05.mma (V as MN-major B
//            operand, no transpose); epilogue O/l -> ctx, log-sum-exp -> lse.
//   backward CTA = (128 keys, head, sample), loops over query tiles.  S = Q K^T and dP = dO V^T into
//            TMEM; 256 threads recompute P = exp(S - lse), dS = P*(dP*keep - delta)*scale and write
//            P_drop, dS (bf16) to smem ONCE; the same tiles then serve as K-major A (dQ = dS K) and as
//            MN-major A (dV += P^T dO, dK += dS^T Q).  dK,dV accumulate in TMEM over the query loop;
//            dQ tiles are reduced across key tiles with red.global.add.v4.f32 into an fp32 buffer.
//
// The mask is a precomputed bit matrix allow[b, h|0, i, j/32] built once per step by
// attn_build_mask_kernel from key_valid + packed relation words (attn_mask.cuh), 6 words per row
// at L=182 instead of the reference's fp32 [B,L,L,12] tensor.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "attn_mask.cuh"
#include "../../include/samk.h"

namespace samk {

int make_tmap_bf16_2d(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld,
                      int box_cols, int box_rows);
int attn_simt_fwd(const samk_attn_params* p, cudaStream_t stream);
int attn_simt_bwd(const samk_attn_params* p, cudaStream_t stream);

constexpr int TDH = 64;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// allow-bit matrix: words[b][hm][i][w], bit (j&31) of word j>>5 = query i may attend key j
// ---------------------------------------------------------------------------------------------
__global__ void attn_build_mask_kernel(AttnMask m, int H, int Hm, int W, uint32_t* __restrict__ out) {
  __shared__ int sflag;
  const int b = blockIdx.z, hm = blockIdx.y;
  const bool any_valid = sample_any_valid(m, b, &sflag);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.L * W) return;
  const int i = e / W, w = e % W;
  uint32_t bits = 0;
  for (int k = 0; k < 32; ++k) {
    const int j = w * 32 + k;
    if (j < m.L && attn_allowed(m, b, hm, i, j, any_valid)) bits |= 1u << k;
  }
  out[(((size_t)b * Hm + hm) * m.L + i) * W + w] = bits;
}

struct TcArgs {
  void* ctx; float* lse;            // fwd outputs (bwd: lse input)
  const float* delta;               // bwd
  void* dqkv; float* dq_accum;      // bwd outputs
  const uint32_t* allow; int Hm, W;
  int B, H, L;
  int q_tile0;                      // forward: first 128-row query tile to compute
  float scale_log2;                 // scale * log2(e)
  float scale;
  uint32_t drop_thresh; float drop_scale; unsigned long long seed, off;
};

__device__ __forceinline__ uint32_t allow_word(const TcArgs& a, int b, int h, int i, int j0) {
  const int w = j0 >> 5;
  if (i >= a.L || w >= a.W) return 0u;
  const int hm = a.Hm == 1 ? 0 : h;
  return a.allow[(((size_t)b * a.Hm + hm) * a.L + i) * a.W + w];
}

// dropout keep flags for 32 consecutive keys j0..j0+31 of probability row (b,h,i): bit k = keep
__device__ __forceinline__ uint32_t keep_word(const TcArgs& a, int b, int h, int i, int j0) {
  if (!a.drop_thresh) return 0xffffffffu;
  const uint64_t row = ((uint64_t)(b * a.H + h) * a.L + i) * (uint64_t)((a.L + 7) >> 3);   // groups of 8 keys
  uint32_t kw = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) kw |= dropout_keep8(a.seed, a.off, row + (uint64_t)((j0 >> 3) + q), a.drop_thresh) << (8 * q);
  return kw;
}

// write 32 consecutive bf16 values (keys 32*c32 .. +31 of the tile) of row r into a [rows][64-key block]
// 128B-swizzled tile set: block kb = c32/2 (16 KB each, 128 rows x 128 B), chunk16 = (c32&1)*4 + q
__device__ __forceinline__ void store_p_chunk(uint8_t* tile_base, int r, int c32, const float* v) {
  uint8_t* rowp = tile_base + (c32 >> 1) * 16384 + r * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = ((c32 & 1) * 4 + q) ^ (r & 7);
    uint4 u = make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                         pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
    *reinterpret_cast<uint4*>(rowp + chunk * 16) = u;
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int KVT>
__global__ void __launch_bounds__(128, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const TcArgs a) {
  // two CTAs per SM (112 KB each at KVT=192): no alignment slack, the 1024-byte alignment the 128B
  // swizzle needs comes from the declaration and is checked below
  extern __shared__ __align__(1024) uint8_t smem_fwd[];
  uint8_t* smem = smem_fwd;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                       // 128 x 64 bf16
  uint8_t* sK = sQ + 16384;                 // KVT x 64
  uint8_t* sV = sK + KVT * 128;             // KVT x 64
  uint8_t* sP = sV + KVT * 128;             // (KVT/64) blocks of 128 x 64
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (KVT / 64) * 16384);
  uint64_t* tma_bar = bars; uint64_t* s_bar = bars + 1; uint64_t* o_bar = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  constexpr int kTmemCols = 256;            // S: KVT (<=192) columns, O: 64 columns at offset 192
  constexpr uint32_t kOcol = 192;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = (blockIdx.x + a.q_tile0) * 128, h = blockIdx.y, b = blockIdx.z;
  const int L = a.L, H = a.H;
  const int n_kv = (L + KVT - 1) / KVT;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmQ); ptx::prefetch_tensormap(&tmKV);
    ptx::mbar_init(tma_bar, 1); ptx::mbar_init(s_bar, 1); ptx::mbar_init(o_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);

  const int row = q0 + tid;                 // query row of this thread
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, KVT, 0, 0);
  constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, 64, 0, 1);
  float m_run = -INFINITY, l_run = 0.f;

  for (int t = 0; t < n_kv; ++t) {
    const int k0 = t * KVT;
    if (tid == 0) {
      ptx::mbar_arrive_expect_tx(tma_bar, (t == 0 ? 16384 : 0) + 2 * KVT * 128);
      if (t == 0) ptx::tma_load_2d(sQ, &tmQ, tma_bar, h * TDH, b * L + q0);
      ptx::tma_load_2d(sK, &tmKV, tma_bar, (H + h) * TDH, b * L + k0);
      ptx::tma_load_2d(sV, &tmKV, tma_bar, (2 * H + h) * TDH, b * L + k0);
    }
    // allow-bit words of this row for the whole key tile: issued before the waits so the global-load
    // latency hides behind TMA + MMA
    uint32_t aw_t[KVT / 32];
#pragma unroll
    for (int c = 0; c < KVT / 32; ++c) aw_t[c] = allow_word(a, b, h, row, k0 + c * 32);
    ptx::mbar_wait(tma_bar, t & 1);
    if (tid == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ptx::umma_f16(tmem, ptx::make_smem_desc_sw128(ptx::smem_u32(sQ) + k * 32, 16, 1024),
                      ptx::make_smem_desc_sw128(ptx::smem_u32(sK) + k * 32, 16, 1024), idesc_s, k > 0);
      ptx::umma_commit(s_bar);
    }
    ptx::mbar_wait(s_bar, t & 1);
    ptx::tc_fence_after();

    // ---- pass 1: masked row maximum of this key tile
    float tmax = -INFINITY;
#pragma unroll
    for (int c = 0; c < KVT / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(t_lane + c * 32, r);
      ptx::tmem_ld_wait();
      const uint32_t aw = aw_t[c];
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if ((aw >> i) & 1u) tmax = fmaxf(tmax, __uint_as_float(r[i]));
    }
    tmax *= a.scale_log2;                   // scale > 0: max commutes with the scaling
    const float m_new = fmaxf(m_run, tmax);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    if (t > 0) {
      // rescale the running sum and the O accumulator (whole CTA takes this path together)
      const float corr = (m_run == -INFINITY) ? 1.f : fast_exp2(m_run - m_use);
      l_run *= corr;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_lane + kOcol + c * 32, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
        tmem_st_32x32(t_lane + kOcol + c * 32, r);
      }
      tmem_st_wait();
    }
    m_run = m_new;
    // ---- pass 2: probabilities -> bf16 P tile (dropout applied), running sum from undropped p
#pragma unroll
    for (int c = 0; c < KVT / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(t_lane + c * 32, r);
      ptx::tmem_ld_wait();
      const uint32_t aw = aw_t[c];
      const uint32_t kw = aw ? keep_word(a, b, h, row, k0 + c * 32) : 0u;
      float p[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = ((aw >> i) & 1u) ? fast_exp2(__uint_as_float(r[i]) * a.scale_log2 - m_use) : 0.f;
        l_run += e;
        p[i] = ((kw >> i) & 1u) ? e * a.drop_scale : 0.f;
      }
      store_p_chunk(sP, tid, c, p);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int k = 0; k < KVT / 16; ++k)
        ptx::umma_f16(tmem + kOcol,
                      ptx::make_smem_desc_sw128(ptx::smem_u32(sP) + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                      ptx::make_smem_desc_sw128(ptx::smem_u32(sV) + k * 2048, 16384, 1024), idesc_o,
                      (t > 0 || k > 0) ? 1u : 0u);
      ptx::umma_commit(o_bar);
    }
    ptx::mbar_wait(o_bar, t & 1);
    ptx::tc_fence_after();
  }

  // ---- epilogue: ctx = O / l, lse = ln(sum exp)
  const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
  __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(a.ctx) + ((size_t)b * L + row) * (size_t)(H * TDH) + h * TDH;
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(t_lane + kOcol + c * 32, r);
    ptx::tmem_ld_wait();
    if (row < L) {
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        *reinterpret_cast<uint4*>(crow + c * 32 + i) =
            make_uint4(pack_bf16(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
                       pack_bf16(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
                       pack_bf16(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
                       pack_bf16(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv));
    }
  }
  if (row < L) a.lse[((size_t)b * H + h) * L + row] = l_run > 0.f ? (m_run + log2f(l_run)) * kLn2 : INFINITY;

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<kTmemCols>(tmem); }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ dctx, const __nv_bfloat16* __restrict__ ctx,
                                  float* __restrict__ delta, int B, int H, int L) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (row, head), 8 lanes... simple: 1 thread
  if (e >= B * L * H) return;
  const int h = e % H, rowi = e / H, b = rowi / L, i = rowi % L;
  const uint4* pa = reinterpret_cast<const uint4*>(dctx + (size_t)rowi * H * TDH + h * TDH);
  const uint4* pb = reinterpret_cast<const uint4*>(ctx + (size_t)rowi * H * TDH + h * TDH);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint4 x = pa[k], y = pb[k];
    const __nv_bfloat162* xa = reinterpret_cast<const __nv_bfloat162*>(&x);
    const __nv_bfloat162* ya = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 f = __bfloat1622float2(xa[q]), g = __bfloat1622float2(ya[q]);
      s += f.x * g.x + f.y * g.y;
    }
  }
  delta[((size_t)b * H + h) * L + i] = s;
}

__global__ void __launch_bounds__(256, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                 // 128 keys x 64
  uint8_t* sV = sK + 16384;
  uint8_t* sQ = sV + 16384;           // 128 rows x 64
  uint8_t* sDO = sQ + 16384;
  uint8_t* sP = sDO + 16384;          // P_drop : 2 blocks of [128 rows][64 keys]
  uint8_t* sDS = sP + 32768;          // dS*scale
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 32768);
  uint64_t* kv_bar = bars; uint64_t* q_bar = bars + 1; uint64_t* s_bar = bars + 2; uint64_t* g_bar = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  constexpr int kTmemCols = 512;
  constexpr uint32_t kScol = 0, kDPcol = 128, kDVcol = 256, kDKcol = 320, kDQcol = 384;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, half = warp >> 2;          // TMEM lane quarter, column half
  const int j0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int L = a.L, H = a.H;
  const int n_q = (L + 127) / 128;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmQKV); ptx::prefetch_tensormap(&tmDO);
    ptx::mbar_init(kv_bar, 1); ptx::mbar_init(q_bar, 1); ptx::mbar_init(s_bar, 1); ptx::mbar_init(g_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);
  const int r = lq * 32 + lane;                        // tile row handled by this thread (query row / key row)

  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, 128, 0, 0);    // S, dP
  constexpr uint32_t idesc_kv = ptx::make_idesc_bf16(128, 64, 1, 1);    // dV, dK : A = P^T / dS^T (MN-major), B MN-major
  constexpr uint32_t idesc_q = ptx::make_idesc_bf16(128, 64, 0, 1);     // dQ : A = dS (K-major), B = K (MN-major)

  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(kv_bar, 2 * 16384);
    ptx::tma_load_2d(sK, &tmQKV, kv_bar, (H + h) * TDH, b * L + j0);
    ptx::tma_load_2d(sV, &tmQKV, kv_bar, (2 * H + h) * TDH, b * L + j0);
  }

  for (int t = 0; t < n_q; ++t) {
    const int i0 = t * 128;
    if (tid == 0) {
      ptx::mbar_arrive_expect_tx(q_bar, 2 * 16384);
      ptx::tma_load_2d(sQ, &tmQKV, q_bar, h * TDH, b * L + i0);
      ptx::tma_load_2d(sDO, &tmDO, q_bar, h * TDH, b * L + i0);
      if (t == 0) ptx::mbar_wait(kv_bar, 0);
      ptx::mbar_wait(q_bar, t & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t dq = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ) + k * 32, 16, 1024);
        const uint64_t dk = ptx::make_smem_desc_sw128(ptx::smem_u32(sK) + k * 32, 16, 1024);
        ptx::umma_f16(tmem + kScol, dq, dk, idesc_s, k > 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t dg = ptx::make_smem_desc_sw128(ptx::smem_u32(sDO) + k * 32, 16, 1024);
        const uint64_t dv = ptx::make_smem_desc_sw128(ptx::smem_u32(sV) + k * 32, 16, 1024);
        ptx::umma_f16(tmem + kDPcol, dg, dv, idesc_s, k > 0);
      }
      ptx::umma_commit(s_bar);
    }
    // per-row scalars and allow words first: their global-load latency hides behind TMA + MMA
    const int i = i0 + r;
    const float lse2 = i < L ? a.lse[((size_t)b * H + h) * L + i] * kLog2e : INFINITY;
    const float dlt = i < L ? a.delta[((size_t)b * H + h) * L + i] : 0.f;
    uint32_t aw_t[2];
    aw_t[0] = allow_word(a, b, h, i, j0 + (half * 2) * 32);
    aw_t[1] = allow_word(a, b, h, i, j0 + (half * 2 + 1) * 32);
    ptx::mbar_wait(s_bar, t & 1);
    ptx::tc_fence_after();

#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;
      uint32_t s[32], d[32];
      ptx::tmem_ld_32x32(t_lane + kScol + c * 32, s);
      ptx::tmem_ld_32x32(t_lane + kDPcol + c * 32, d);
      ptx::tmem_ld_wait();
      const uint32_t aw = aw_t[cc];
      const uint32_t kw = aw ? keep_word(a, b, h, i, j0 + c * 32) : 0u;
      float pv[32], dsv[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float p = ((aw >> k) & 1u) ? fast_exp2(__uint_as_float(s[k]) * a.scale_log2 - lse2) : 0.f;
        const float keep = ((kw >> k) & 1u) ? a.drop_scale : 0.f;
        pv[k] = p * keep;
        dsv[k] = p * (__uint_as_float(d[k]) * keep - dlt) * a.scale;
      }
      store_p_chunk(sP, r, c, pv);
      store_p_chunk(sDS, r, c, dsv);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      // dV[j,d] += sum_i P[i,j] dO[i,d] ; dK[j,d] += sum_i dS[i,j] Q[i,d]   (contraction over the 128 rows)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t ap = ptx::make_smem_desc_sw128(ptx::smem_u32(sP) + k * 2048, 16384, 1024);
        const uint64_t bg = ptx::make_smem_desc_sw128(ptx::smem_u32(sDO) + k * 2048, 16384, 1024);
        ptx::umma_f16(tmem + kDVcol, ap, bg, idesc_kv, (t > 0 || k > 0) ? 1u : 0u);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t as = ptx::make_smem_desc_sw128(ptx::smem_u32(sDS) + k * 2048, 16384, 1024);
        const uint64_t bq = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ) + k * 2048, 16384, 1024);
        ptx::umma_f16(tmem + kDKcol, as, bq, idesc_kv, (t > 0 || k > 0) ? 1u : 0u);
      }
      // dQ[i,d] = sum_j dS[i,j] K[j,d]   (contraction over the 128 keys of this CTA)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint64_t as = ptx::make_smem_desc_sw128(ptx::smem_u32(sDS) + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
        const uint64_t bk = ptx::make_smem_desc_sw128(ptx::smem_u32(sK) + k * 2048, 16384, 1024);
        ptx::umma_f16(tmem + kDQcol, as, bk, idesc_q, k > 0);
      }
      ptx::umma_commit(g_bar);
    }
    ptx::mbar_wait(g_bar, t & 1);
    ptx::tc_fence_after();
    // ---- dQ tile -> fp32 accumulation buffer (each thread: its row, 32 of the 64 columns)
    {
      uint32_t q[32];
      ptx::tmem_ld_32x32(t_lane + kDQcol + half * 32, q);
      ptx::tmem_ld_wait();
      if (i < L) {
        float* dst = a.dq_accum + ((size_t)b * L + i) * (size_t)(H * TDH) + h * TDH + half * 32;
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + k), "f"(__uint_as_float(q[k])),
                       "f"(__uint_as_float(q[k + 1])), "f"(__uint_as_float(q[k + 2])), "f"(__uint_as_float(q[k + 3]))
                       : "memory");
      }
    }
    ptx::tc_fence_before();
    __syncthreads();   // all TMEM reads of this iteration done before the next S/dP MMAs overwrite them
  }

  // ---- dK / dV rows -> dqkv (bf16): warps 0-3 write dK, warps 4-7 write dV; thread = key row
  {
    const int j = j0 + r;
    const uint32_t col = half == 0 ? kDKcol : kDVcol;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.dqkv) + ((size_t)b * L + j) * (size_t)(3 * H * TDH) +
                         (size_t)((half == 0 ? H : 2 * H) + h) * TDH;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(t_lane + col + c * 32, v);
      ptx::tmem_ld_wait();
      if (j < L) {
#pragma unroll
        for (int k = 0; k < 32; k += 8)
          *reinterpret_cast<uint4*>(dst + c * 32 + k) =
              make_uint4(pack_bf16(__uint_as_float(v[k]), __uint_as_float(v[k + 1])),
                         pack_bf16(__uint_as_float(v[k + 2]), __uint_as_float(v[k + 3])),
                         pack_bf16(__uint_as_float(v[k + 4]), __uint_as_float(v[k + 5])),
                         pack_bf16(__uint_as_float(v[k + 6]), __uint_as_float(v[k + 7])));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<kTmemCols>(tmem); }
}

// dq_accum fp32 [rows, H*64] -> bf16 q-part of dqkv [rows, 3*H*64]
__global__ void attn_dq_store_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, size_t rows, int hd) {
  const size_t n4 = rows * (size_t)(hd / 4);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e / (hd / 4), c = (e % (hd / 4)) * 4;
    float4 v = *reinterpret_cast<const float4*>(acc + r * hd + c);
    *reinterpret_cast<uint2*>(dqkv + r * 3 * hd + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static int fill_tc(TcArgs& a, const samk_attn_params* p) {
  if (!p->allow_bits) { set_error("samk_attn: tensor-core path needs allow_bits (samk_attn_build_mask)"); return SAMK_ERR_ARG; }
  a.ctx = p->ctx; a.lse = p->lse; a.delta = p->delta; a.dqkv = p->dqkv; a.dq_accum = p->dq_accum;
  a.allow = p->allow_bits; a.Hm = p->spatial ? p->H : 1;
  a.B = p->B; a.H = p->H; a.L = p->T + p->A + p->D; a.W = (a.L + 31) / 32;
  a.scale = p->scale; a.scale_log2 = p->scale * kLog2e;
  a.drop_thresh = p->drop_p > 0.f ? drop_threshold(p->drop_p) : 0u;
  a.drop_scale = drop_keep_scale(p->drop_p);
  a.seed = p->drop_seed; a.off = p->drop_offset;
  a.q_tile0 = p->q_begin > 0 ? p->q_begin / 128 : 0;
  return SAMK_OK;
}

template <class K>
static int set_smem(K kern, int bytes) {
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%d) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return SAMK_ERR_CUDA;
  }
  return SAMK_OK;
}

template <int KVT>
static int launch_fwd(const samk_attn_params* p, const TcArgs& a, cudaStream_t stream) {
  const long long rows = (long long)a.B * a.L;
  const int hd3 = 3 * a.H * TDH;
  CUtensorMap tq, tkv;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, p->qkv, rows, hd3, hd3, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tkv, p->qkv, rows, hd3, hd3, 64, KVT))) return rc;
  constexpr int smem = 16384 + 2 * KVT * 128 + (KVT / 64) * 16384 + 64;
  if ((rc = set_smem(attn_fwd_tc_kernel<KVT>, smem))) return rc;
  dim3 grid((a.L + 127) / 128 - a.q_tile0, a.H, a.B);
  attn_fwd_tc_kernel<KVT><<<grid, 128, smem, stream>>>(tq, tkv, a);
  return check_launch("samk_attn_fwd(tc)");
}

int attn_tc_fwd(const samk_attn_params* p, cudaStream_t stream) {
  TcArgs a;
  int rc = fill_tc(a, p);
  if (rc) return rc;
  if (!p->qkv || !p->ctx || !p->lse) { set_error("samk_attn_fwd: null pointer"); return SAMK_ERR_ARG; }
  if (!a.B || !a.L) return SAMK_OK;
  // one 192-key tile covers the shipped L=182; otherwise stream 128-key tiles
  if (a.L > 128 && a.L <= 192) return launch_fwd<192>(p, a, stream);
  return launch_fwd<128>(p, a, stream);
}

int attn_tc_bwd(const samk_attn_params* p, cudaStream_t stream) {
  TcArgs a;
  int rc = fill_tc(a, p);
  if (rc) return rc;
  if (!p->qkv || !p->ctx || !p->lse || !p->dctx || !p->dqkv || !p->delta || !p->dq_accum) {
    set_error("samk_attn_bwd: null pointer (tensor-core path also needs dq_accum)");
    return SAMK_ERR_ARG;
  }
  if (!a.B || !a.L) return SAMK_OK;
  const long long rows = (long long)a.B * a.L;
  const int hd = a.H * TDH;
  attn_delta_kernel<<<(unsigned)((rows * a.H + 255) / 256), 256, 0, stream>>>(
      (const __nv_bfloat16*)p->dctx, (const __nv_bfloat16*)p->ctx, p->delta, a.B, a.H, a.L);
  if ((rc = check_launch("samk_attn_bwd(delta)"))) return rc;
  if (cudaMemsetAsync(p->dq_accum, 0, (size_t)rows * hd * sizeof(float), stream) != cudaSuccess) {
    set_error("samk_attn_bwd: memset failed");
    return SAMK_ERR_CUDA;
  }
  CUtensorMap tqkv, tdo;
  if ((rc = make_tmap_bf16_2d(&tqkv, p->qkv, rows, 3 * hd, 3 * hd, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tdo, p->dctx, rows, hd, hd, 64, 128))) return rc;
  constexpr int smem = 4 * 16384 + 2 * 32768 + 1024 + 64;
  if ((rc = set_smem(attn_bwd_tc_kernel, smem))) return rc;
  dim3 grid((a.L + 127) / 128, a.H, a.B);
  attn_bwd_tc_kernel<<<grid, 256, smem, stream>>>(tqkv, tdo, a);
  if ((rc = check_launch("samk_attn_bwd(tc)"))) return rc;
  attn_dq_store_kernel<<<148 * 8, 256, 0, stream>>>(p->dq_accum, (__nv_bfloat16*)p->dqkv, (size_t)rows, hd);
  return check_launch("samk_attn_bwd(dq store)");
}

}  // namespace samk

extern "C" {

long long samk_attn_mask_words(int B, int H, int T, int A, int D, int spatial) {
  const long long L = (long long)T + A + D;
  return (long long)B * (spatial ? H : 1) * L * ((L + 31) / 32);
}

int samk_attn_build_mask(const samk_attn_params* p, uint32_t* allow_bits, void* stream) {
  using namespace samk;
  if (!p || !p->key_valid || !allow_bits) { set_error("samk_attn_build_mask: null pointer"); return SAMK_ERR_ARG; }
  if (p->spatial && p->A > 0 && !p->rel_bits) { set_error("samk_attn_build_mask: spatial needs rel_bits"); return SAMK_ERR_ARG; }
  AttnMask m;
  m.valid = p->key_valid; m.rel = p->spatial ? p->rel_bits : nullptr;
  m.T = p->T; m.A = p->A; m.D = p->D; m.L = p->T + p->A + p->D;
  m.quad_mask = p->spatial ? p->quadrant_mask : 0u; m.spatial = p->spatial ? 1 : 0;
  if (!p->B || !m.L) return SAMK_OK;
  const int W = (m.L + 31) / 32, Hm = p->spatial ? p->H : 1;
  dim3 grid((m.L * W + 127) / 128, Hm, p->B);
  attn_build_mask_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(m, p->H, Hm, W, allow_bits);
  return check_launch("samk_attn_build_mask");
}

int samk_attn_fwd(const samk_attn_params* p, int impl, void* stream) {
  if (!p) { samk::set_error("samk_attn_fwd: null params"); return SAMK_ERR_ARG; }
  if (impl == 0 && p->dtype == SAMK_DT_BF16) return samk::attn_tc_fwd(p, (cudaStream_t)stream);
  return samk::attn_simt_fwd(p, (cudaStream_t)stream);
}

int samk_attn_bwd(const samk_attn_params* p, int impl, void* stream) {
  if (!p) { samk::set_error("samk_attn_bwd: null params"); return SAMK_ERR_ARG; }
  if (impl == 0 && p->dtype == SAMK_DT_BF16) return samk::attn_tc_bwd(p, (cudaStream_t)stream);
  return samk::attn_simt_bwd(p, (cudaStream_t)stream);
}

}  // extern "C"
