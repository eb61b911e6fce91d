// Fused masked multi-head attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
// Storage formats: q|k|v, P and ctx are IEEE half (11 significant bits: the forward pass stays within 1e-3 of the
// fp32 reference).  The gradients dctx (in) and dq|dk|dv (out) are bfloat16.  tcgen05.mma kind::f16 needs A and B in
// ONE format (mixed f16 x bf16 faults on sm_100a), so the backward runs every product in half: attn_bwd_prep_kernel
// re-expresses dO of each (sample, head) in half with an exactly computed power-of-two scale s_bh (max|dO| s in [8,16)),
// everything downstream (delta, dP, dS, dQ, dK, dV) lives in that scaled domain -- all contractions of the backward
// are within one (sample, head), so s_bh factors out -- and the drains multiply by 1/s_bh before rounding to bf16.
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
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "attn_mask.cuh"
#include "../../include/samk.h"

namespace samk {

int make_tmap_bf16_2d(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld,
                      int box_cols, int box_rows);
int attn_simt_fwd(const samk_attn_params* p, cudaStream_t stream);
int attn_simt_bwd(const samk_attn_params* p, cudaStream_t stream);
int sm_count();

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
// One thread builds the word (32 keys) of one query row for ALL heads: the packed relation word of a pair already
// holds every head's bit, so the pair is classified once (attn_mask.cuh rules) instead of once per head.
__global__ void attn_build_mask_kernel(AttnMask m, int H, int Hm, int W, uint32_t* __restrict__ out) {
  __shared__ int sflag;
  const int b = blockIdx.y;
  const bool any_valid = sample_any_valid(m, b, &sflag);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.L * W) return;
  const int i = e / W, w = e % W;
  const int si = seg_of(m, i);
  uint32_t words[16];
#pragma unroll
  for (int h = 0; h < 16; ++h) words[h] = 0u;
  for (int k = 0; k < 32; ++k) {
    const int j = w * 32 + k;
    if (j >= m.L) break;
    const int sj = seg_of(m, j);
    const bool ok = (sj == 2) ? (si == 2 && j <= i) : (m.valid[(size_t)b * m.L + j] != 0);
    uint32_t hb;                                   // bit h = head h may attend (i -> j)
    if (!m.spatial) {
      hb = ((!any_valid && si != 2) || ok) ? 0xFFFFu : 0u;
    } else if ((m.quad_mask >> (3 * si + sj)) & 1u) {
      hb = 0u;
    } else if (si == 1 && sj == 1) {
      hb = ok ? (uint32_t)m.rel[((size_t)b * m.A + (i - m.T)) * m.A + (j - m.T)] : 0u;
    } else {
      hb = ok ? 0xFFFFu : 0u;
    }
#pragma unroll
    for (int h = 0; h < 16; ++h) words[h] |= ((hb >> h) & 1u) << k;
  }
#pragma unroll
  for (int h = 0; h < 16; ++h)
    if (h < Hm) out[(((size_t)b * Hm + h) * m.L + i) * W + w] = words[h];
}

struct TcArgs {
  void* ctx; float* lse;            // fwd outputs (bwd: lse input)
  const float* delta;               // bwd
  void* dqkv; float* dq_accum;      // bwd outputs
  const uint32_t* allow; int Hm, W;
  const uint32_t* keep;             // dropout keep bits [B, H, L, W] (samk_attn_build_keep) or nullptr = keep all
  const float* inv_scale;           // bwd: 1 / s_bh per (sample, head), from attn_bwd_prep_kernel
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
  const int w = j0 >> 5;
  if (!a.keep || i >= a.L || w >= a.W) return 0xffffffffu;
  return a.keep[(((size_t)b * a.H + h) * a.L + i) * a.W + w];
}

// Dropout keep bits of the attention probabilities, one word per (b, h, query row, 32 keys): bit k = element
// (i, 32 w + k) is kept.  The same Philox stream as the exact-fp32 kernel (attn_simt.cu: group of 8 keys g = j >> 3 of
// row ((b H + h) L + i) ngrp), evaluated ONCE per layer and step by this kernel -- beside the q|k|v projection GEMM,
// whose tensor-bound CTAs leave the integer pipes idle -- instead of once in the forward and once in the backward
// attention kernel, where the 7 Philox rounds were a third of all issued instructions.
__global__ void attn_build_keep_kernel(uint32_t* __restrict__ out, long long n_words, int L, int W, uint32_t thresh,
                                       unsigned long long seed, unsigned long long off) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_words) return;
  const int w = (int)(e % W);
  const long long row = e / W;                    // (b H + h) L + i
  const uint32_t ngrp = (uint32_t)((L + 7) >> 3);
  uint32_t kw = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t g = (uint32_t)(4 * w + q);
    if (g < ngrp) kw |= dropout_keep8(seed, off, (uint64_t)row * ngrp + g, thresh) << (8 * q);
  }
  out[e] = kw;
}

// write 32 consecutive bf16 values (keys 32*c32 .. +31 of the tile) of row r into a [rows][64-key block]
// 128B-swizzled tile set: block kb = c32/2 (16 KB each, 128 rows x 128 B), chunk16 = (c32&1)*4 + q
template <bool F16>
__device__ __forceinline__ void store_p_chunk(uint8_t* tile_base, int r, int c32, const float* v) {
  uint8_t* rowp = tile_base + (c32 >> 1) * 16384 + r * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = ((c32 & 1) * 4 + q) ^ (r & 7);
    uint4 u = make_uint4(pack_16<F16>(v[8 * q], v[8 * q + 1]), pack_16<F16>(v[8 * q + 2], v[8 * q + 3]),
                         pack_16<F16>(v[8 * q + 4], v[8 * q + 5]), pack_16<F16>(v[8 * q + 6], v[8 * q + 7]));
    *reinterpret_cast<uint4*>(rowp + chunk * 16) = u;
  }
}

// ---------------------------------------------------------------------------------------------
// forward, persistent + warp-specialised (the product path)
//
//   warp 0      TMA producer: Q tile pair (2 stages) and K/V tiles (KS-stage ring)
//   warps 1,2   MMA issuers (one lane each), one per query tile g of the pair: S_g = Q_g K^T into TMEM,
//               O_g (+)= P_g V with P_g read from TMEM
//   warps 3-18  softmax: 8 warps per query tile = 4 TMEM lane quarters (32 query rows each) x 2 halves of the
//               key tile; a row's two threads exchange their partial maximum / sum through shared memory
//   work item   (sample, head, query-tile pair); a CTA walks items blockIdx.x, +gridDim.x, ...
//
// Per key tile: each thread reads its score row from TMEM once for the masked maximum and once for
// p = 2^(s*c - m); the probabilities go back into TMEM as packed bf16 over the columns of S (no
// shared-memory round trip: the MMA's shared-memory read port is the scarce resource on this part) and
// are the A operand of the P V product.  Dropout bits come straight from Philox comparisons as select
// predicates; the 1/(1-p) scale and the softmax denominator are applied once to the 64 output columns.
// TMEM: group g uses columns [256 g, 256 g + KVT) for S/P and [256 g + 192, 256 g + 256) for O.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

#ifdef SAMK_TIMELINE
// developer instrumentation (tools/attn_timeline.py): clock stamps of CTA 0's first items
__device__ long long g_timeline[4096];
#define TL_STAMP(slot, idx) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (idx) < 64) g_timeline[(slot) * 64 + (idx)] = clock64(); } while (0)
// converged-warp variant (MMA issuer warps: elect.sync needs the warp reconverged afterwards)
#define TL_STAMP_W(slot, idx) do { TL_STAMP(slot, idx); __syncwarp(); } while (0)
// any warp of CTA 0 (per-warp arrival skew)
#define TL_STAMP_ANY(slot, idx) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (idx) < 64 && (slot) < 64) g_timeline[(slot) * 64 + (idx)] = clock64(); } while (0)
#else
#define TL_STAMP(slot, idx) do { } while (0)
#define TL_STAMP_W(slot, idx) do { } while (0)
#define TL_STAMP_ANY(slot, idx) do { } while (0)
#endif

template <int KVT> struct Fwd2Cfg {
  static constexpr int KS = KVT == 192 ? 2 : 3;             // K/V ring depth
  static constexpr int kQBytes = 2 * 16384;                 // one stage: two 128 x 64 query tiles
  static constexpr int kKVBytes = 2 * KVT * 128;            // one stage: K tile + V tile
  static constexpr int kBarOff = 2 * kQBytes + KS * kKVBytes;
  static constexpr int kXchOff = kBarOff + 256;             // row max / row sum exchange between column halves
  static constexpr int kXchBytes = 2 /*kind*/ * 2 /*slot*/ * 2 /*tile*/ * 2 /*half*/ * 128 * 4;
  static constexpr int kOutOff = kXchOff + kXchBytes;       // per softmax warp: 32 rows x 32 bf16 output transpose
  static constexpr int kSmem = kOutOff + 16 * 2048 + 1024;
};
constexpr int kFwd2Threads = 96 + 16 * 32;     // producer, one MMA issuer per query tile, 16 softmax warps

// allow words are re-read by all 12 heads' CTAs (non-spatial) / once per head: keep them in L2, skip L1
__device__ __forceinline__ uint32_t ld_allow(const uint32_t* p) { return __ldg(p); }

__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!ptx::mbar_try_wait(bar, parity)) __nanosleep(64);
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// The two key halves of one TMEM lane quarter (warps w and w + 4) meet on their own 64-thread barrier
// `first + quarter` instead of a barrier across all softmax warps of the tile: a pair never waits for the slowest of
// four.  Immediate barrier ids so that ptxas reserves only the barriers that are used.
template <int kFirst>
__device__ __forceinline__ void pair_bar_sync(int quarter) {
  switch (quarter) {
    case 0: asm volatile("bar.sync %0, 64;" ::"n"(kFirst) : "memory"); break;
    case 1: asm volatile("bar.sync %0, 64;" ::"n"(kFirst + 1) : "memory"); break;
    case 2: asm volatile("bar.sync %0, 64;" ::"n"(kFirst + 2) : "memory"); break;
    default: asm volatile("bar.sync %0, 64;" ::"n"(kFirst + 3) : "memory"); break;
  }
}

__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// masked maximum of 16 scores (aw: allow bits of these 16 keys in the low half)
__device__ __forceinline__ float masked_max16(const uint32_t (&r)[16], uint32_t aw, float m) {
  float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    if ((aw >> i) & 1u) m0 = fmaxf(m0, __uint_as_float(r[i]));
    if ((aw >> (i + 1)) & 1u) m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
    if ((aw >> (i + 2)) & 1u) m2 = fmaxf(m2, __uint_as_float(r[i + 2]));
    if ((aw >> (i + 3)) & 1u) m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// 16 scores -> probabilities (masked, dropout applied) as 8 packed half pairs; returns the sum of the
// undropped probabilities.  aw / kw: allow and dropout-keep bits of these 16 keys in the low half.
__device__ __forceinline__ float softmax_chunk16(const uint32_t (&r)[16], uint32_t aw, uint32_t kw, float sl2, float m_use,
                                                 uint32_t (&pk)[8]) {
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
  kw &= aw;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float e[8], p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x = fast_exp2(fmaf(__uint_as_float(r[8 * q + i]), sl2, -m_use));
      e[i] = ((aw >> (8 * q + i)) & 1u) ? x : 0.f;
      p[i] = ((kw >> (8 * q + i)) & 1u) ? x : 0.f;
    }
    l0 += e[0] + e[4]; l1 += e[1] + e[5]; l2 += e[2] + e[6]; l3 += e[3] + e[7];
#pragma unroll
    for (int i = 0; i < 4; ++i) pk[4 * q + i] = pack_f16(p[2 * i], p[2 * i + 1]);
  }
  return (l0 + l1) + (l2 + l3);
}

template <int KVT>
__global__ void __launch_bounds__(kFwd2Threads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const TcArgs a,
                 int n_qp, int qp0, int n_items, uint32_t magic_qp, uint32_t magic_h) {
  using Cfg = Fwd2Cfg<KVT>;
  constexpr int KS = Cfg::KS;
  constexpr int NC = KVT / 32;                          // 16-key chunks per thread (its half of the key tile)
  extern __shared__ uint8_t smem_raw2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2][2][128 x 64]
  uint8_t* sKV = smem + 2 * Cfg::kQBytes;               // [KS][K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* kv_full = bars + 4;       // [KS]
  uint64_t* kv_empty = bars + 4 + KS; // [KS]
  uint64_t* s_full = bars + 4 + 2 * KS;   // [2]
  uint64_t* p_ready = s_full + 2;         // [2]
  uint64_t* o_done = s_full + 4;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 6);
  float* xch = reinterpret_cast<float*>(smem + Cfg::kXchOff);   // [kind][slot][tile][half][128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L, H = a.H;
  const int n_kv = (L + KVT - 1) / KVT;
  const int n_qt = (L + 127) / 128;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmQ); ptx::prefetch_tensormap(&tmKV);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&q_full[i], 1); ptx::mbar_init(&q_empty[i], 2); }
    for (int i = 0; i < KS; ++i) { ptx::mbar_init(&kv_full[i], 1); ptx::mbar_init(&kv_empty[i], 2); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&p_ready[i], 8); ptx::mbar_init(&o_done[i], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // item -> (b, h, first query tile); the last, partial query tile of a sample is loaded shifted up on odd
  // iterations (its rows then sit in the upper TMEM lanes) so that both halves of the SM's warp
  // schedulers get the short tile's work in turn
  auto decode = [&](int item, int iter, int& b, int& h, int& qs0, int& qs1, int& ql0, int& ql1) {
    // exact for item < 2^32 / divisor (magic = ceil(2^32 / divisor), host side)
    const int bh = n_qp == 1 ? item : (int)__umulhi((uint32_t)item, magic_qp);
    const int qp = qp0 + item - bh * n_qp;
    b = (int)__umulhi((uint32_t)bh, magic_h);
    h = bh - b * H;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int qt = 2 * qp + g;
      int st = qt * 128, lo = qt * 128;
      if (qt >= n_qt || qt < a.q_tile0) { st = -1; lo = 0; }
      else if ((iter & 1) && qt == n_qt - 1 && L >= 128 && L - qt * 128 <= 64) st = L - 128;
      if (g == 0) { qs0 = st; ql0 = lo; } else { qs1 = st; ql1 = lo; }
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int iter = 0; uint32_t kvc = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
        int b, h, qs0, qs1, ql0, ql1;
        decode(item, iter, b, h, qs0, qs1, ql0, ql1);
        const int qs = iter & 1;
        mbar_wait_backoff(&q_empty[qs], ((iter >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&q_full[qs], ((qs0 >= 0) + (qs1 >= 0)) * 16384);
        if (qs0 >= 0) ptx::tma_load_2d(sQ + qs * Cfg::kQBytes, &tmQ, &q_full[qs], h * TDH, b * L + qs0);
        if (qs1 >= 0) ptx::tma_load_2d(sQ + qs * Cfg::kQBytes + 16384, &tmQ, &q_full[qs], h * TDH, b * L + qs1);
        for (int t = 0; t < n_kv; ++t, ++kvc) {
          const int ks = kvc % KS;
          mbar_wait_backoff(&kv_empty[ks], ((kvc / KS) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&kv_full[ks], Cfg::kKVBytes);
          uint8_t* sk = sKV + ks * Cfg::kKVBytes;
          ptx::tma_load_2d(sk, &tmKV, &kv_full[ks], (H + h) * TDH, b * L + t * KVT);
          ptx::tma_load_2d(sk + KVT * 128, &tmKV, &kv_full[ks], (2 * H + h) * TDH, b * L + t * KVT);
        }
      }
    }
  } else if (warp <= 2) {
    // ===================== MMA issuers: warp 1 drives query tile 0, warp 2 query tile 1 =====================
    // (independent pipelines: the short last tile of a sample runs ahead instead of waiting for the full one).
    // The whole warp runs the loop converged and computes the warp-uniform descriptors; only the tcgen05
    // instructions sit under the elected-lane predicate, so they issue back to back from uniform registers.
    {
      constexpr uint32_t idesc_s = ptx::make_idesc_16(128, KVT, 0, 0, ptx::kFmtF16, ptx::kFmtF16);
      constexpr uint32_t idesc_o = ptx::make_idesc_16(128, 64, 0, 1, ptx::kFmtF16, ptx::kFmtF16);
      constexpr uint64_t kStepK = 32 >> 4, kStepMN = 2048 >> 4;   // k-step increments of the descriptor address field
      const int g = warp - 1;
      int iter = 0; uint32_t kvc = 0; uint32_t pc = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
        int b, h, qs0, qs1, ql0, ql1;
        decode(item, iter, b, h, qs0, qs1, ql0, ql1);
        const bool exists = (g ? qs1 : qs0) >= 0;
        const int qs = iter & 1;
        if (g == 0) TL_STAMP_W(0, iter);
        mbar_wait_backoff(&q_full[qs], (iter >> 1) & 1);
        for (int t = 0; t < n_kv; ++t, ++kvc) {
          const int ks = kvc % KS;
          mbar_wait_backoff(&kv_full[ks], (kvc / KS) & 1);
          ptx::tc_fence_after();
          if (g == 0) TL_STAMP_W(1, iter);
          const uint32_t sk = ptx::smem_u32(sKV + ks * Cfg::kKVBytes), sv = sk + KVT * 128;
          const uint64_t dq = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ + qs * Cfg::kQBytes + g * 16384), 16, 1024);
          const uint64_t dk = ptx::make_smem_desc_sw128(sk, 16, 1024);
          const uint64_t dv = ptx::make_smem_desc_sw128(sv, 16384, 1024);
          if (ptx::elect_one()) {
            if (exists) {
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem + g * 256, dq + k * kStepK, dk + k * kStepK, idesc_s, k > 0);
              ptx::umma_commit(&s_full[g]);
            }
            if (t == n_kv - 1) {                     // all S products of this item issued: Q stage may be refilled
              if (exists) ptx::umma_commit(&q_empty[qs]); else ptx::mbar_arrive(&q_empty[qs]);
            }
          }
          __syncwarp();
          if (exists) {
            if (g == 0) TL_STAMP_W(2, iter);
            ptx::mbar_wait(&p_ready[g], pc & 1); ++pc;
            ptx::tc_fence_after();
            if (g == 0) TL_STAMP_W(3, iter);
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < KVT / 16; ++k)   // P (packed bf16) sits at the start of each column half of S
                ptx::umma_f16_ts(tmem + g * 256 + 192, tmem + g * 256 + (k < KVT / 32 ? k * 8 : KVT / 2 + (k - KVT / 32) * 8),
                                 dv + k * kStepMN, idesc_o, (t > 0 || k > 0) ? 1u : 0u);
              ptx::umma_commit(&o_done[g]);
              ptx::umma_commit(&kv_empty[ks]);
            }
          } else {
            if (ptx::elect_one()) ptx::mbar_arrive(&kv_empty[ks]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== softmax warps =====================
    // 16 warps: query tile g = (warp-3)/8, column half hf = ((warp-3)/4)&1 (keys [hf*KVT/2, +KVT/2) of the key
    // tile and output columns [32 hf, +32)), TMEM lane quarter = warp % 4.  The two halves of a row meet
    // twice per key tile through shared memory (row maximum) and once per item (row sum).
    const int g = (warp - 3) >> 3, hf = ((warp - 3) >> 2) & 1, quarter = warp & 3;
    const int lrow = quarter * 32 + lane;
    const uint32_t t_s = tmem + g * 256 + ((uint32_t)(quarter * 32) << 16);   // S columns of this tile
    const uint32_t t_o = t_s + 192 + hf * 32;                                    // this half's 32 output columns
    const float sl2 = a.scale_log2;
    uint32_t sc = 0, oc = 0, xc = 0;
    int iter = 0;
    // allow-bit words of this thread's (row, key half), fetched one (item, key tile) ahead of their use
    uint32_t aw_nx[NC / 2], kw_nx[NC / 2];
    auto fetch_allow = [&](int item_, int iter_, int t_) {
#pragma unroll
      for (int c = 0; c < NC / 2; ++c) { aw_nx[c] = 0u; kw_nx[c] = 0xffffffffu; }
      if (item_ >= n_items) return;
      int b_, h_, s0, s1, l0, l1;
      decode(item_, iter_, b_, h_, s0, s1, l0, l1);
      const int st = g ? s1 : s0, lo = g ? l1 : l0;
      const int row_ = st + lrow;
      if (st < 0 || row_ < lo || row_ >= L) return;
      const uint32_t* ar = a.allow + (((size_t)b_ * a.Hm + (a.Hm == 1 ? 0 : h_)) * L + row_) * a.W;
      const int w0 = (t_ * KVT + hf * (KVT / 2)) >> 5;
#pragma unroll
      for (int c = 0; c < NC / 2; ++c) if (w0 + c < a.W) aw_nx[c] = ld_allow(ar + w0 + c);
      if (a.keep) {
        const uint32_t* kr = a.keep + (((size_t)b_ * H + h_) * L + row_) * a.W;
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) if (w0 + c < a.W) kw_nx[c] = ld_allow(kr + w0 + c);
      }
    };
    fetch_allow(blockIdx.x, 0, 0);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
      int b, h, qs0, qs1, ql0, ql1;
      decode(item, iter, b, h, qs0, qs1, ql0, ql1);
      const int my_start = g ? qs1 : qs0, my_lo = g ? ql1 : ql0;
      if (my_start < 0) { fetch_allow(item + gridDim.x, iter + 1, 0); continue; }
      const int row = my_start + lrow;
      const bool active = row >= my_lo && row < L;
      const bool warp_active = __any_sync(0xffffffffu, active);
      float m_run = -INFINITY, l_run = 0.f;
      for (int t = 0; t < n_kv; ++t) {
        uint32_t aw_t[NC / 2], kw_t[NC / 2];
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) { aw_t[c] = aw_nx[c]; kw_t[c] = kw_nx[c]; }
        if (t + 1 < n_kv) fetch_allow(item, iter, t + 1); else fetch_allow(item + gridDim.x, iter + 1, 0);
        if (warp == 3 && lane == 0) TL_STAMP(8, iter);
        ptx::mbar_wait(&s_full[g], sc & 1); ++sc;
        ptx::tc_fence_after();
        if (warp == 3 && lane == 0) TL_STAMP(9, iter);
        const uint32_t t_sh = t_s + hf * (KVT / 2);
        uint32_t rA[16], rB[16];
        // ---- pass 1: masked maximum of this half of the row (TMEM loads one chunk ahead of the math)
        float tmax = -INFINITY;
        if (warp_active) {
          ptx::tmem_ld_32x16(t_sh, rA);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (c + 1 < NC) ptx::tmem_ld_32x16(t_sh + (c + 1) * 16, (c & 1) ? rA : rB);
            tmax = masked_max16((c & 1) ? rB : rA, aw_t[c >> 1] >> (16 * (c & 1)), tmax);
            if (c + 1 < NC) ptx::tmem_ld_wait();
          }
          // first chunk of pass 2 is fetched while the halves exchange their maxima
          ptx::tmem_ld_32x16(t_sh, rA);
        }
        float* xm = xch + (((0 * 2 + (xc & 1)) * 2 + g) * 2) * 128;
        xm[hf * 128 + lrow] = tmax;
        if (warp == 3 && lane == 0) TL_STAMP(10, iter);
        if (g) pair_bar_sync<5>(quarter); else pair_bar_sync<1>(quarter);
        if (warp == 3 && lane == 0) TL_STAMP(11, iter);
        tmax = fmaxf(tmax, xm[(hf ^ 1) * 128 + lrow]);
        ++xc;
        if (warp_active) {
          tmax *= sl2;                               // scale > 0: max commutes with the scaling
          const float m_new = fmaxf(m_run, tmax);
          const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
          if (t > 0) {
            ptx::tmem_ld_wait();
            ptx::mbar_wait(&o_done[g], oc & 1); ++oc;     // P V of the previous key tile has landed in O
            ptx::tc_fence_after();
            const float corr = (m_run == -INFINITY) ? 1.f : fast_exp2(m_run - m_use);
            l_run *= corr;
            if (__any_sync(0xffffffffu, corr != 1.f)) {
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                ptx::tmem_ld_32x16(t_o + c * 16, rB);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) rB[i] = __float_as_uint(__uint_as_float(rB[i]) * corr);
                tmem_st_32x16(t_o + c * 16, rB);
              }
            }
          }
          m_run = m_new;
          // ---- pass 2: p = 2^(s c - m) on allowed keys; kept (dropout) values -> packed bf16 P in TMEM
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (c + 1 < NC) ptx::tmem_ld_32x16(t_sh + (c + 1) * 16, (c & 1) ? rA : rB);
            uint32_t pk[8];
            l_run += softmax_chunk16((c & 1) ? rB : rA, aw_t[c >> 1] >> (16 * (c & 1)), kw_t[c >> 1] >> (16 * (c & 1)), sl2,
                                     m_use, pk);
            if (c + 1 < NC) ptx::tmem_ld_wait();      // also orders the P store below after the S load of the same columns
            tmem_st_32x8(t_sh + c * 8, pk);   // P of this half overlays the first columns of its own S half
          }
          tmem_st_wait();
        } else if (t > 0) {
          ++oc;                                   // keep the O-phase count in step without touching TMEM
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_ready[g]);
        if (warp == 3 && lane == 0) TL_STAMP(12, iter);
      }
      // ---- epilogue: ctx = O * keep_scale / l, lse = ln(sum exp); the halves add their partial sums
      float* xl = xch + (((1 * 2 + (iter & 1)) * 2 + g) * 2) * 128;
      xl[hf * 128 + lrow] = l_run;
      if (g) pair_bar_sync<5>(quarter); else pair_bar_sync<1>(quarter);
      l_run += xl[(hf ^ 1) * 128 + lrow];
      if (warp == 3 && lane == 0) TL_STAMP(13, iter);
      if (warp_active) {
        ptx::mbar_wait(&o_done[g], oc & 1); ++oc;
        ptx::tc_fence_after();
        if (warp == 3 && lane == 0) TL_STAMP(14, iter);
        const float inv = l_run > 0.f ? a.drop_scale / l_run : 0.f;
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_o, r);
        ptx::tmem_ld_wait();
        // transpose through shared memory: a row-per-thread store would touch 32 cache lines per instruction
        const uint32_t st = ptx::smem_u32(smem + Cfg::kOutOff) + (warp - 3) * 2048;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = 8 * j;
          sts128u(st + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4),
                  make_uint4(pack_f16(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
                             pack_f16(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
                             pack_f16(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
                             pack_f16(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv)));
        }
        __syncwarp();
        const int row_w = my_start + quarter * 32;      // first row of this warp
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 8 + (lane >> 2), cj = lane & 3, orow = row_w + rr;
          const uint4 v = lds128u(st + rr * 64 + ((cj ^ ((rr >> 1) & 3)) << 4));
          if (orow >= my_lo && orow < L)
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(a.ctx) + ((size_t)b * L + orow) * (size_t)(H * TDH) + h * TDH +
                                      hf * 32 + cj * 8) = v;
        }
        __syncwarp();
        if (active && hf == 0) a.lse[((size_t)b * H + h) * L + row] = l_run > 0.f ? (m_run + log2f(l_run)) * kLn2 : INFINITY;
        if (warp == 3 && lane == 0) TL_STAMP(15, iter);
      } else {
        ++oc;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem); }
}

// ---------------------------------------------------------------------------------------------
// forward v3: one query tile per CTA, TWO CTAs per SM.  Same arithmetic and TMEM use as v2 per tile (S/P 192 columns +
// O 64 columns = 256 columns per CTA), but the two resident CTAs are in different phases of their items, so one
// CTA's softmax fills the issue slots while the other waits for its TMA / MMA / barriers.  Work items are ordered
// full tiles first, the short last tile of every sample afterwards (static round-robin then balances the CTAs).
//   warp 0 TMA producer (Q double-buffered, K/V single stage), warp 1 MMA issuer, warps 2-9 softmax
//   (TMEM lane quarter = warp % 4, key half = (warp-2)/4).
// ---------------------------------------------------------------------------------------------
template <int KVT> struct Fwd3Cfg {
  static constexpr int kKVOff = 2 * 16384;                  // after the two Q stages
  static constexpr int kKVBytes = 2 * KVT * 128;            // K tile + V tile
  static constexpr int kBarOff = kKVOff + kKVBytes;
  static constexpr int kXchOff = kBarOff + 256;             // [kind][slot][half][128] floats
  static constexpr int kXchBytes = 2 * 2 * 2 * 128 * 4;
  static constexpr int kOutOff = kXchOff + kXchBytes;       // per softmax warp: 32 rows x 32 bf16
  static constexpr int kSmem = kOutOff + 8 * 2048 + 1024;
};
constexpr int kFwd3Threads = 64 + 8 * 32;

template <int KVT>
__global__ void __launch_bounds__(kFwd3Threads, 2)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const TcArgs a,
                 int n_bh, int n_items, uint32_t magic_bh, uint32_t magic_h) {
  using Cfg = Fwd3Cfg<KVT>;
  constexpr int NC = KVT / 32;                          // 16-key chunks per thread (its half of the key tile)
  extern __shared__ uint8_t smem_raw4[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw4) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2][128 x 64]
  uint8_t* sKV = smem + Cfg::kKVOff;                    // K | V
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = bars + 2;       // [2]
  uint64_t* kv_full = bars + 4;
  uint64_t* kv_empty = bars + 5;
  uint64_t* s_full = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* o_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* xch = reinterpret_cast<float*>(smem + Cfg::kXchOff);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L, H = a.H;
  const int n_kv = (L + KVT - 1) / KVT;
  const int n_qt = (L + 127) / 128;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmQ); ptx::prefetch_tensormap(&tmKV);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&q_full[i], 1); ptx::mbar_init(&q_empty[i], 1); }
    ptx::mbar_init(kv_full, 1); ptx::mbar_init(kv_empty, 1);
    ptx::mbar_init(s_full, 1); ptx::mbar_init(p_ready, 8); ptx::mbar_init(o_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // item -> (query tile, b, h): tiles ascending (the short last tile of every sample comes last); that tile is loaded
  // shifted up on odd iterations so its rows alternate between the lower and the upper TMEM lane quarters
  // chunked (magic_bh == 0): the (sample, head) pairs are taken gridDim.x at a time -- tile 0 of the whole chunk, then
  // tile 1 of the same chunk, ... -- so a pair's K/V rows are read again from L2 (a chunk's K/V is ~14 MB), not from DRAM,
  // and with the static round-robin schedule every CTA still alternates between full and short tiles.
  auto decode = [&](int item, int iter, int& b, int& h, int& q_start, int& q_lo) {
    int qi, bh;
    if (magic_bh) {
      qi = (int)__umulhi((uint32_t)item, magic_bh);               // exact for item < 2^32 / n_bh
      bh = item - qi * n_bh;
    } else {
      const int G = (int)gridDim.x, per = G * (n_qt - a.q_tile0);
      const int c = item / per, r = item - c * per;
      const int m = min(G, n_bh - c * G);                         // pairs in this chunk (the last one may be short)
      qi = r / m;
      bh = c * G + (r - qi * m);
    }
    const int qt = a.q_tile0 + qi;
    b = (int)__umulhi((uint32_t)bh, magic_h);
    h = bh - b * H;
    q_lo = qt * 128;
    q_start = qt * 128;
    // (chunked order: a CTA meets the short tile every (n_qt - q_tile0)-th iteration, so the flip follows that count)
    const bool flip = magic_bh ? (iter & 1) : ((iter / (n_qt - a.q_tile0)) & 1);
    if (flip && qt == n_qt - 1 && L >= 128 && L - qt * 128 <= 64) q_start = L - 128;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int iter = 0; uint32_t kvc = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
        int b, h, q_start, q_lo;
        decode(item, iter, b, h, q_start, q_lo);
        const int qs = iter & 1;
        mbar_wait_backoff(&q_empty[qs], ((iter >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&q_full[qs], 16384);
        ptx::tma_load_2d(sQ + qs * 16384, &tmQ, &q_full[qs], h * TDH, b * L + q_start);
        for (int t = 0; t < n_kv; ++t, ++kvc) {
          mbar_wait_backoff(kv_empty, (kvc & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(kv_full, Cfg::kKVBytes);
          ptx::tma_load_2d(sKV, &tmKV, kv_full, (H + h) * TDH, b * L + t * KVT);
          ptx::tma_load_2d(sKV + KVT * 128, &tmKV, kv_full, (2 * H + h) * TDH, b * L + t * KVT);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, tcgen05 under the elected lane) =====================
    constexpr uint32_t idesc_s = ptx::make_idesc_16(128, KVT, 0, 0, ptx::kFmtF16, ptx::kFmtF16);
    constexpr uint32_t idesc_o = ptx::make_idesc_16(128, 64, 0, 1, ptx::kFmtF16, ptx::kFmtF16);
    constexpr uint64_t kStepK = 32 >> 4, kStepMN = 2048 >> 4;
    int iter = 0; uint32_t kvc = 0, pc = 0;
    const uint32_t sk = ptx::smem_u32(sKV), sv = sk + KVT * 128;
    const uint64_t dk = ptx::make_smem_desc_sw128(sk, 16, 1024);
    const uint64_t dv = ptx::make_smem_desc_sw128(sv, 16384, 1024);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
      const int qs = iter & 1;
      mbar_wait_backoff(&q_full[qs], (iter >> 1) & 1);
      const uint64_t dq = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ + qs * 16384), 16, 1024);
      for (int t = 0; t < n_kv; ++t, ++kvc) {
        mbar_wait_backoff(kv_full, kvc & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem, dq + k * kStepK, dk + k * kStepK, idesc_s, k > 0);
          ptx::umma_commit(s_full);
          if (t == n_kv - 1) ptx::umma_commit(&q_empty[qs]);     // all S products of this item issued
        }
        __syncwarp();
        ptx::mbar_wait(p_ready, pc & 1); ++pc;
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < KVT / 16; ++k)   // P (packed bf16) sits at the start of each column half of S
            ptx::umma_f16_ts(tmem + 192, tmem + (k < KVT / 32 ? k * 8 : KVT / 2 + (k - KVT / 32) * 8), dv + k * kStepMN, idesc_o,
                             (t > 0 || k > 0) ? 1u : 0u);
          ptx::umma_commit(o_done);
          ptx::umma_commit(kv_empty);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax warps =====================
    const int hf = (warp - 2) >> 2, quarter = warp & 3;
    const int lrow = quarter * 32 + lane;
    const uint32_t t_s = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t t_o = t_s + 192 + hf * 32;
    const float sl2 = a.scale_log2;
    uint32_t sc = 0, oc = 0, xc = 0;
    int iter = 0;
    uint32_t aw_nx[NC / 2], kw_nx[NC / 2];
    auto fetch_allow = [&](int item_, int iter_, int t_) {
#pragma unroll
      for (int c = 0; c < NC / 2; ++c) { aw_nx[c] = 0u; kw_nx[c] = 0xffffffffu; }
      if (item_ >= n_items) return;
      int b_, h_, st, lo;
      decode(item_, iter_, b_, h_, st, lo);
      const int row_ = st + lrow;
      if (row_ < lo || row_ >= L) return;
      const uint32_t* ar = a.allow + (((size_t)b_ * a.Hm + (a.Hm == 1 ? 0 : h_)) * L + row_) * a.W;
      const int w0 = (t_ * KVT + hf * (KVT / 2)) >> 5;
#pragma unroll
      for (int c = 0; c < NC / 2; ++c) if (w0 + c < a.W) aw_nx[c] = ld_allow(ar + w0 + c);
      if (a.keep) {
        const uint32_t* kr = a.keep + (((size_t)b_ * H + h_) * L + row_) * a.W;
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) if (w0 + c < a.W) kw_nx[c] = ld_allow(kr + w0 + c);
      }
    };
    fetch_allow(blockIdx.x, 0, 0);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
      int b, h, my_start, my_lo;
      decode(item, iter, b, h, my_start, my_lo);
      const int row = my_start + lrow;
      const bool active = row >= my_lo && row < L;
      const bool warp_active = __any_sync(0xffffffffu, active);
      float m_run = -INFINITY, l_run = 0.f;
      for (int t = 0; t < n_kv; ++t) {
        uint32_t aw_t[NC / 2], kw_t[NC / 2];
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) { aw_t[c] = aw_nx[c]; kw_t[c] = kw_nx[c]; }
        if (t + 1 < n_kv) fetch_allow(item, iter, t + 1); else fetch_allow(item + gridDim.x, iter + 1, 0);
        ptx::mbar_wait(s_full, sc & 1); ++sc;
        ptx::tc_fence_after();
        const uint32_t t_sh = t_s + hf * (KVT / 2);
        uint32_t rA[16], rB[16];
        float tmax = -INFINITY;
        if (warp_active) {
          ptx::tmem_ld_32x16(t_sh, rA);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (c + 1 < NC) ptx::tmem_ld_32x16(t_sh + (c + 1) * 16, (c & 1) ? rA : rB);
            tmax = masked_max16((c & 1) ? rB : rA, aw_t[c >> 1] >> (16 * (c & 1)), tmax);
            if (c + 1 < NC) ptx::tmem_ld_wait();
          }
          ptx::tmem_ld_32x16(t_sh, rA);            // first chunk of pass 2, in flight during the exchange
        }
        float* xm = xch + ((0 * 2 + (xc & 1)) * 2) * 128;
        xm[hf * 128 + lrow] = tmax;
        pair_bar_sync<1>(quarter);
        tmax = fmaxf(tmax, xm[(hf ^ 1) * 128 + lrow]);
        ++xc;
        if (warp_active) {
          tmax *= sl2;
          const float m_new = fmaxf(m_run, tmax);
          const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
          if (t > 0) {
            ptx::tmem_ld_wait();
            ptx::mbar_wait(o_done, oc & 1); ++oc;
            ptx::tc_fence_after();
            const float corr = (m_run == -INFINITY) ? 1.f : fast_exp2(m_run - m_use);
            l_run *= corr;
            if (__any_sync(0xffffffffu, corr != 1.f)) {
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                ptx::tmem_ld_32x16(t_o + c * 16, rB);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) rB[i] = __float_as_uint(__uint_as_float(rB[i]) * corr);
                tmem_st_32x16(t_o + c * 16, rB);
              }
            }
          }
          m_run = m_new;
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            if (c + 1 < NC) ptx::tmem_ld_32x16(t_sh + (c + 1) * 16, (c & 1) ? rA : rB);
            uint32_t pk[8];
            l_run += softmax_chunk16((c & 1) ? rB : rA, aw_t[c >> 1] >> (16 * (c & 1)), kw_t[c >> 1] >> (16 * (c & 1)), sl2,
                                     m_use, pk);
            if (c + 1 < NC) ptx::tmem_ld_wait();
            tmem_st_32x8(t_sh + c * 8, pk);
          }
          tmem_st_wait();
        } else if (t > 0) {
          ++oc;
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(p_ready);
      }
      // ---- epilogue
      float* xl = xch + ((1 * 2 + (iter & 1)) * 2) * 128;
      xl[hf * 128 + lrow] = l_run;
      pair_bar_sync<1>(quarter);
      l_run += xl[(hf ^ 1) * 128 + lrow];
      if (warp_active) {
        ptx::mbar_wait(o_done, oc & 1); ++oc;
        ptx::tc_fence_after();
        const float inv = l_run > 0.f ? a.drop_scale / l_run : 0.f;
        uint32_t r[32];
        ptx::tmem_ld_32x32(t_o, r);
        ptx::tmem_ld_wait();
        const uint32_t st = ptx::smem_u32(smem + Cfg::kOutOff) + (warp - 2) * 2048;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = 8 * j;
          sts128u(st + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4),
                  make_uint4(pack_f16(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
                             pack_f16(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
                             pack_f16(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
                             pack_f16(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv)));
        }
        __syncwarp();
        const int row_w = my_start + quarter * 32;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 8 + (lane >> 2), cj = lane & 3, orow = row_w + rr;
          const uint4 v = lds128u(st + rr * 64 + ((cj ^ ((rr >> 1) & 3)) << 4));
          if (orow >= my_lo && orow < L)
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(a.ctx) + ((size_t)b * L + orow) * (size_t)(H * TDH) + h * TDH +
                                      hf * 32 + cj * 8) = v;
        }
        __syncwarp();
        if (active && hf == 0) a.lse[((size_t)b * H + h) * L + row] = l_run > 0.f ? (m_run + log2f(l_run)) * kLn2 : INFINITY;
      } else {
        ++oc;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<256>(tmem); }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// Backward preparation, one CTA per (sample, head): max|dO| over the item -> s = 2^k with max|dO| s in [8, 16);
// dO16 = half(dO s) (what the tensor cores read); delta[b,h,i] = sum_d dO16[i,d] O[i,d] (the row term of the softmax
// backward, in the scaled domain, from the very values the MMAs see); inv_scale[b,h] = 1/s.  dO rows of one head are
// 128 contiguous bytes: 8 lanes x 16 bytes per row, both passes hit L1/L2 (an item is <= 33 KB at L = 256).
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ dctx, const __half* __restrict__ ctx, __half* __restrict__ do16,
                     float* __restrict__ delta, float* __restrict__ inv_scale, int H, int L) {
  __shared__ float red[8];
  const int item = blockIdx.x, b = item / H, h = item - b * H;
  const int t = threadIdx.x, k8 = t & 7;
  const size_t ld = (size_t)H * TDH;
  const size_t base = (size_t)b * L * ld + (size_t)h * TDH + (size_t)k8 * 8;
  float m = 0.f;
  for (int i = t >> 3; i < L; i += 32) {
    const uint4 x = *reinterpret_cast<const uint4*>(dctx + base + (size_t)i * ld);
    const __nv_bfloat162* xa = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(xa[q]); m = fmaxf(m, fmaxf(fabsf(f.x), fabsf(f.y))); }
  }
  m = warp_max(m);
  if ((t & 31) == 0) red[t >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  const float s = pow2_scale_for(m, 4);
  if (t == 0) inv_scale[item] = 1.0f / s;
  for (int i0 = 0; i0 < L; i0 += 32) {
    const int i = i0 + (t >> 3);
    float acc = 0.f;
    if (i < L) {
      const uint4 x = *reinterpret_cast<const uint4*>(dctx + base + (size_t)i * ld);
      const uint4 y = *reinterpret_cast<const uint4*>(ctx + base + (size_t)i * ld);
      const __nv_bfloat162* xa = reinterpret_cast<const __nv_bfloat162*>(&x);
      const __half2* ya = reinterpret_cast<const __half2*>(&y);
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(xa[q]);
        o[q] = pack_f16_sat(f.x * s, f.y * s);
        const float2 r = __half22float2(*reinterpret_cast<const __half2*>(&o[q])), g = __half22float2(ya[q]);
        acc += r.x * g.x + r.y * g.y;
      }
      *reinterpret_cast<uint4*>(do16 + base + (size_t)i * ld) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (i < L && k8 == 0) delta[((size_t)b * H + h) * L + i] = acc;
  }
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
  const float inv_s = __ldg(a.inv_scale + b * H + h);          // dO16 = dO * s_bh: the results leave the scaled domain here

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

  using ptx::kFmtF16;
  constexpr uint32_t idesc_s = ptx::make_idesc_16(128, 128, 0, 0, kFmtF16, kFmtF16);     // S = Q K^T (half x half)
  constexpr uint32_t idesc_dp = idesc_s;                                                  // dP = dO16 V^T
  constexpr uint32_t idesc_dv = ptx::make_idesc_16(128, 64, 1, 1, kFmtF16, kFmtF16);     // dV: A = Pd^T (MN-major), B = dO16 (MN-major)
  constexpr uint32_t idesc_dk = idesc_dv;                                                 // dK: A = dS^T (MN-major), B = Q (MN-major)
  constexpr uint32_t idesc_q = ptx::make_idesc_16(128, 64, 0, 1, kFmtF16, kFmtF16);      // dQ: A = dS (K-major), B = K (MN-major)

  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(kv_bar, 2 * 16384);
    ptx::tma_load_2d(sK, &tmQKV, kv_bar, (H + h) * TDH, b * L + j0);
    ptx::tma_load_2d(sV, &tmQKV, kv_bar, (2 * H + h) * TDH, b * L + j0);
  }

  for (int t = 0; t < n_q; ++t) {
    const int i0 = t * 128;
    if (warp == 0) {
      // warp 0 converged: descriptors in uniform registers, tcgen05 / TMA under the elected lane (see gemm.cu)
      constexpr uint64_t kStepK = 32 >> 4;
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(q_bar, 2 * 16384);
        ptx::tma_load_2d(sQ, &tmQKV, q_bar, h * TDH, b * L + i0);
        ptx::tma_load_2d(sDO, &tmDO, q_bar, h * TDH, b * L + i0);
      }
      __syncwarp();
      if (t == 0) ptx::mbar_wait(kv_bar, 0);
      ptx::mbar_wait(q_bar, t & 1);
      ptx::tc_fence_after();
      const uint64_t dq = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ), 16, 1024), dk = ptx::make_smem_desc_sw128(ptx::smem_u32(sK), 16, 1024);
      const uint64_t dg = ptx::make_smem_desc_sw128(ptx::smem_u32(sDO), 16, 1024), dv = ptx::make_smem_desc_sw128(ptx::smem_u32(sV), 16, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem + kScol, dq + k * kStepK, dk + k * kStepK, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem + kDPcol, dg + k * kStepK, dv + k * kStepK, idesc_dp, k > 0);
        ptx::umma_commit(s_bar);
      }
      __syncwarp();
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
      store_p_chunk<true>(sP, r, c, pv);
      store_p_chunk<true>(sDS, r, c, dsv);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      ptx::tc_fence_after();
      constexpr uint64_t kStepK = 32 >> 4, kStepMN = 2048 >> 4;
      const uint64_t ap = ptx::make_smem_desc_sw128(ptx::smem_u32(sP), 16384, 1024), bg = ptx::make_smem_desc_sw128(ptx::smem_u32(sDO), 16384, 1024);
      const uint64_t as_t = ptx::make_smem_desc_sw128(ptx::smem_u32(sDS), 16384, 1024), bq = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ), 16384, 1024);
      const uint64_t as0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sDS), 16, 1024), as1 = ptx::make_smem_desc_sw128(ptx::smem_u32(sDS) + 16384, 16, 1024);
      const uint64_t bk = ptx::make_smem_desc_sw128(ptx::smem_u32(sK), 16384, 1024);
      if (ptx::elect_one()) {
        // dV[j,d] += sum_i P[i,j] dO[i,d] ; dK[j,d] += sum_i dS[i,j] Q[i,d]   (contraction over the 128 rows)
#pragma unroll
        for (int k = 0; k < 8; ++k) ptx::umma_f16(tmem + kDVcol, ap + k * kStepMN, bg + k * kStepMN, idesc_dv, (t > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k) ptx::umma_f16(tmem + kDKcol, as_t + k * kStepMN, bq + k * kStepMN, idesc_dk, (t > 0 || k > 0) ? 1u : 0u);
        // dQ[i,d] = sum_j dS[i,j] K[j,d]   (contraction over the 128 keys of this CTA)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          ptx::umma_f16(tmem + kDQcol, (k < 4 ? as0 : as1) + (k & 3) * kStepK, bk + k * kStepMN, idesc_q, k > 0);
        ptx::umma_commit(g_bar);
      }
      __syncwarp();
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
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + k), "f"(__uint_as_float(q[k]) * inv_s),
                       "f"(__uint_as_float(q[k + 1]) * inv_s), "f"(__uint_as_float(q[k + 2]) * inv_s), "f"(__uint_as_float(q[k + 3]) * inv_s)
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
              make_uint4(pack_bf16(__uint_as_float(v[k]) * inv_s, __uint_as_float(v[k + 1]) * inv_s),
                         pack_bf16(__uint_as_float(v[k + 2]) * inv_s, __uint_as_float(v[k + 3]) * inv_s),
                         pack_bf16(__uint_as_float(v[k + 4]) * inv_s, __uint_as_float(v[k + 5]) * inv_s),
                         pack_bf16(__uint_as_float(v[k + 6]) * inv_s, __uint_as_float(v[k + 7]) * inv_s));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<kTmemCols>(tmem); }
}

// ---------------------------------------------------------------------------------------------
// backward, persistent + warp-specialised, for L <= 256 (every shipped geometry): one CTA owns a whole
// (sample, head) at a time, so dQ, dK and dV are complete in TMEM and are written once as bf16 -- no
// fp32 atomics, no accumulation buffer, no conversion pass.
//
//   warp 0      TMA producer: Q, dO (<= 2 tiles of 128 rows), K, V (64-row chunks) of the next item
//   warp 1      MMA issuer
//   warps 2-17  16 element-wise warps: TMEM lane quarter = warp % 4 (32 query rows), column slice = (warp-2)/4
//   blocks      (key tile jt, query tile it), jt outer, processed as 64-key sub-blocks whose S / dP ping-pong
//               between two TMEM buffers so the tensor cores run one sub-block ahead of the 16 warps.  Per block:
//                 S = Q_it K_jt^T, dP = dO_it V_jt^T                     (tensor cores -> TMEM)
//                 P = 2^(S c - lse), Pd = P keep/(1-p), dS = P (dP keep/(1-p) - delta)   (16 warps)
//                   Pd, dS -> bf16, 128B-swizzled shared memory tiles [128 query rows][key columns]
//                 dV_jt += Pd^T dO_it, dK_jt += dS^T Q_it, dQ_it += dS K_jt   (tensor cores; the staged tiles
//                   serve as MN-major A for the first two and as K-major A for the third)
//               after the last query tile of a key tile the element-wise warps drain dK_jt, dV_jt; after the last
//               block they drain dQ.  The softmax scale is applied when dQ and dK are drained.
//   TMEM        [S 64 | dP 64] x 2 buffers 0..255 | dV 256..319 | dK 320..383 | dQ_0 384..447 | dQ_1 448..511
// ---------------------------------------------------------------------------------------------
constexpr int kBwd2Threads = 64 + 16 * 32;
struct Bwd2Smem {
  static constexpr int kQ = 0;                    // [2][128 x 64]
  static constexpr int kDO = 32768;               // [2][128 x 64]
  static constexpr int kK = 65536;                // [4][64 x 64]
  static constexpr int kV = 98304;                // [4][64 x 64]
  static constexpr int kP = 131072;               // Pd : 2 blocks of [128 rows][64 keys]
  static constexpr int kDS = 163840;              // dS
  static constexpr int kOut = 196608;             // 16 warps x 2 KB drain transposes
  static constexpr int kBar = 196608 + 32768;
  static constexpr int kTotal = kBar + 128 + 1024;
};

// drain 32 TMEM columns of this warp's 32 lanes as bf16 (scaled) into out[row_base + lane][col0 .. col0+32),
// rows limited to [row_lo, row_hi); the store is transposed through a warp-private 2 KB tile so that one warp
// instruction covers 8 rows x 64 contiguous bytes
__device__ __forceinline__ void drain32_bf16(uint32_t taddr, float mul, uint32_t st, __nv_bfloat16* out, long long ld,
                                             long long row_base, int row_lo_rel, int row_hi_rel, int lane) {
  uint32_t r[32];
  ptx::tmem_ld_32x32(taddr, r);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = 8 * j;
    sts128u(st + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4),
            make_uint4(pack_bf16(__uint_as_float(r[i]) * mul, __uint_as_float(r[i + 1]) * mul),
                       pack_bf16(__uint_as_float(r[i + 2]) * mul, __uint_as_float(r[i + 3]) * mul),
                       pack_bf16(__uint_as_float(r[i + 4]) * mul, __uint_as_float(r[i + 5]) * mul),
                       pack_bf16(__uint_as_float(r[i + 6]) * mul, __uint_as_float(r[i + 7]) * mul)));
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + (lane >> 2), cj = lane & 3;
    const uint4 v = lds128u(st + rr * 64 + ((cj ^ ((rr >> 1) & 3)) << 4));
    if (rr >= row_lo_rel && rr < row_hi_rel) *reinterpret_cast<uint4*>(out + (row_base + rr) * ld + cj * 8) = v;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kBwd2Threads, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                 const __grid_constant__ CUtensorMap tmDO, const TcArgs a, int n_items, uint32_t magic_h) {
  using SM = Bwd2Smem;
  extern __shared__ uint8_t smem_raw3[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw3) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBar);
  // Inputs are tracked per REGION so that the next item's tiles land as soon as the current item has read their
  // buffers for the last time (the shared memory is full: no second set of input buffers).  Region 2 it = Q / dO16 of
  // query tile it, region 2 jt + 1 = the K / V chunks of key tile jt.
  uint64_t* reg_full = bars; uint64_t* reg_empty = bars + 4; uint64_t* s_full = bars + 8;   // reg_full[4], reg_empty[4], s_full[2]
  // sub_done[2], alternating per sub-block: S / dP of a sub-block are ready long before the slower warps finish the
  // previous one, so with a single barrier a fast warp's arrival for sub-block u+1 would complete the phase of u
  // early; with two, running two ahead is impossible (S / dP of u+2 are only produced after sub_done(u))
  uint64_t* sub_done = bars + 10; uint64_t* stage_free = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.L, H = a.H;
  const int n_t = (L + 127) / 128;                       // query tiles = key tiles (1 or 2)
  const int n_last = (L - (n_t - 1) * 128 <= 64) ? 1 : 2;      // 64-key sub-blocks of the last key tile
  const int kv_chunks = (n_t - 1) * 2 + n_last;          // 64-row chunks of K / V to fetch
  // sub-block u of an item = (key tile jt, query tile it, 64-key half jh), jt outermost, jh innermost
  const int subs_full = 2 * n_t;                         // sub-blocks in a full key tile
  const int n_sub = (n_t - 1) * subs_full + n_last * n_t;
  auto decode_sub = [&](int u, int& jt, int& it, int& jh, int& nh) {
    if (u < (n_t - 1) * subs_full) { jt = 0; nh = 2; it = u >> 1; jh = u & 1; }     // (n_t <= 2: a full tile can only be jt 0)
    else { const int v = u - (n_t - 1) * subs_full; jt = n_t - 1; nh = n_last; it = v / n_last; jh = v - it * n_last; }
  };

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmQKV128); ptx::prefetch_tensormap(&tmQKV64); ptx::prefetch_tensormap(&tmDO);
    for (int r = 0; r < 4; ++r) { ptx::mbar_init(&reg_full[r], 1); ptx::mbar_init(&reg_empty[r], 1); }
    ptx::mbar_init(&s_full[0], 1); ptx::mbar_init(&s_full[1], 1);
    ptx::mbar_init(&sub_done[0], 16); ptx::mbar_init(&sub_done[1], 16); ptx::mbar_init(stage_free, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // S / dP of sub-block u live in buffer u & 1: S at 128 (u&1), dP at 128 (u&1) + 64
  constexpr uint32_t kDVcol = 256, kDKcol = 320, kDQcol = 384;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int iter = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
        const int b = (int)__umulhi((uint32_t)item, magic_h), h = item - b * H;
        const uint32_t par = (uint32_t)(iter & 1) ^ 1u;
        // regions in the order the products need them; each waits only for ITS buffers (released by the MMA warp right
        // after the last product of the previous item that reads them)
        for (int t = 0; t < n_t; ++t) {
          const int rq = 2 * t, rk = 2 * t + 1;
          const int c0 = 2 * t, nc = t < n_t - 1 ? 2 : n_last;            // K / V chunks of key tile t
          mbar_wait_backoff(&reg_empty[rq], par);
          ptx::mbar_arrive_expect_tx(&reg_full[rq], 2 * 16384);
          ptx::tma_load_2d(smem + SM::kQ + t * 16384, &tmQKV128, &reg_full[rq], h * TDH, b * L + t * 128);
          ptx::tma_load_2d(smem + SM::kDO + t * 16384, &tmDO, &reg_full[rq], h * TDH, b * L + t * 128);
          mbar_wait_backoff(&reg_empty[rk], par);
          ptx::mbar_arrive_expect_tx(&reg_full[rk], nc * 2 * 8192);
          for (int c = c0; c < c0 + nc; ++c) ptx::tma_load_2d(smem + SM::kK + c * 8192, &tmQKV64, &reg_full[rk], (H + h) * TDH, b * L + c * 64);
          for (int c = c0; c < c0 + nc; ++c) ptx::tma_load_2d(smem + SM::kV + c * 8192, &tmQKV64, &reg_full[rk], (2 * H + h) * TDH, b * L + c * 64);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop converged and computes the (warp-uniform) descriptors; only the tcgen05
    // instructions sit under the elected-lane predicate.  (With the arithmetic inside an `if (lane == 0)` ptxas
    // wraps every UTCHMMA in an elect / R2UR.BROADCAST / branch loop: ~100 cycles per instruction.)
    {
      const uint32_t sQ = ptx::smem_u32(smem + SM::kQ), sDO = ptx::smem_u32(smem + SM::kDO);
      const uint32_t sK = ptx::smem_u32(smem + SM::kK), sV = ptx::smem_u32(smem + SM::kV);
      const uint32_t sP = ptx::smem_u32(smem + SM::kP), sDS = ptx::smem_u32(smem + SM::kDS);
      using ptx::kFmtF16;
      constexpr uint32_t idesc_s = ptx::make_idesc_16(128, 64, 0, 0, kFmtF16, kFmtF16);     // S = Q K^T of a 64-key sub-block
      constexpr uint32_t idesc_dp = idesc_s;                                                 // dP = dO16 V^T
      constexpr uint32_t idesc_dv = ptx::make_idesc_16(128, 64, 1, 1, kFmtF16, kFmtF16);    // dV: A = Pd^T (MN-major), B = dO16 (MN-major)
      constexpr uint32_t idesc_dk = idesc_dv;                                                // dK: A = dS^T (MN-major), B = Q (MN-major)
      constexpr uint32_t idesc_q = ptx::make_idesc_16(128, 64, 0, 1, kFmtF16, kFmtF16);     // dQ: A = dS (K-major), B = K (MN-major)
      // k-step increments of the descriptor start-address field (bytes >> 4)
      constexpr uint64_t kStepK = 32 >> 4, kStepMN = 2048 >> 4;
      int iter = 0; uint32_t cu = 0;
      uint32_t have = 0;                                  // regions of the current item whose loads have been awaited
      auto need = [&](int r) {
        if (!((have >> r) & 1u)) {
          mbar_wait_backoff(&reg_full[r], (uint32_t)(iter & 1));
          ptx::tc_fence_after();
          have |= 1u << r;
        }
      };
      // last sub-block that reads a region (then its buffers are released to the producer)
      const int last_q0 = (n_t - 1) * subs_full + n_last - 1, last_q1 = n_sub - 1;
      const int last_k0 = n_t > 1 ? subs_full - 1 : n_sub - 1, last_k1 = n_sub - 1;
      auto issue_sdp = [&](int u) {
        int jt, it, jh, nh;
        decode_sub(u, jt, it, jh, nh);
        need(2 * it);
        need(2 * jt + 1);
        const uint32_t col = (uint32_t)(u & 1) * 128u;
        const uint32_t kb = (uint32_t)(jt * 2 + jh) * 8192u;               // 64-row chunk of K / V
        const uint64_t dq = ptx::make_smem_desc_sw128(sQ + it * 16384, 16, 1024), dk = ptx::make_smem_desc_sw128(sK + kb, 16, 1024);
        const uint64_t dg = ptx::make_smem_desc_sw128(sDO + it * 16384, 16, 1024), dv = ptx::make_smem_desc_sw128(sV + kb, 16, 1024);
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {      // the two accumulation chains alternate: consecutive MMAs are independent
            ptx::umma_f16(tmem + col, dq + kk * kStepK, dk + kk * kStepK, idesc_s, kk > 0);
            ptx::umma_f16(tmem + col + 64, dg + kk * kStepK, dv + kk * kStepK, idesc_dp, kk > 0);
          }
          ptx::umma_commit(&s_full[u & 1]);
        }
        __syncwarp();
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
        have = 0;
        issue_sdp(0);
        if (n_sub > 1) issue_sdp(1);
        for (int u = 0; u < n_sub; ++u, ++cu) {
          int jt, it, jh, nh;
          decode_sub(u, jt, it, jh, nh);
          TL_STAMP_W(26, (int)cu);
          ptx::mbar_wait(&sub_done[cu & 1], (cu >> 1) & 1);   // Pd, dS of sub-block u staged; its S / dP buffer is free again
          ptx::tc_fence_after();
          TL_STAMP_W(27, (int)cu);
          // dQ_it[i,d] += sum_j dS[i,j] K[j,d]   (contraction over the 64 keys of this sub-block)
          const uint64_t a_ds = ptx::make_smem_desc_sw128(sDS + jh * 16384, 16, 1024);
          const uint64_t b_k = ptx::make_smem_desc_sw128(sK + (jt * 2 + jh) * 8192, 16384, 1024);
          const uint64_t a_p = ptx::make_smem_desc_sw128(sP, 16384, 1024), b_do = ptx::make_smem_desc_sw128(sDO + it * 16384, 16384, 1024);
          const uint64_t a_dst = ptx::make_smem_desc_sw128(sDS, 16384, 1024), b_q = ptx::make_smem_desc_sw128(sQ + it * 16384, 16384, 1024);
          const uint32_t first_q = (jt > 0 || jh > 0) ? 1u : 0u, first_kv = it > 0 ? 1u : 0u;
          const bool block_end = jh == nh - 1;
          // the products that release the staging tiles go first, the S / dP two sub-blocks ahead last
          if (ptx::elect_one()) {
            if (block_end) {
              // block (jt, it) complete: dV_jt[j,d] += sum_i Pd[i,j] dO[i,d] ; dK_jt[j,d] += sum_i dS[i,j] Q[i,d];
              // the three accumulation chains (dV, dK, dQ) are interleaved so that consecutive MMAs are independent
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                ptx::umma_f16(tmem + kDVcol, a_p + kk * kStepMN, b_do + kk * kStepMN, idesc_dv, kk > 0 ? 1u : first_kv);
                ptx::umma_f16(tmem + kDKcol, a_dst + kk * kStepMN, b_q + kk * kStepMN, idesc_dk, kk > 0 ? 1u : first_kv);
                if (kk < 4)
                  ptx::umma_f16(tmem + kDQcol + it * 64, a_ds + kk * kStepK, b_k + kk * kStepMN, idesc_q, kk > 0 ? 1u : first_q);
              }
              ptx::umma_commit(stage_free);
            } else {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                ptx::umma_f16(tmem + kDQcol + it * 64, a_ds + kk * kStepK, b_k + kk * kStepMN, idesc_q, kk > 0 ? 1u : first_q);
            }
            // buffers nobody reads any more in this item go back to the producer (next item's loads start now)
            if (u == last_q0) ptx::umma_commit(&reg_empty[0]);
            if (u == last_k0) ptx::umma_commit(&reg_empty[1]);
            if (n_t > 1 && u == last_q1) ptx::umma_commit(&reg_empty[2]);
            if (n_t > 1 && u == last_k1) ptx::umma_commit(&reg_empty[3]);
          }
          __syncwarp();
          if (u + 2 < n_sub) issue_sdp(u + 2);
          TL_STAMP_W(28, (int)cu);
        }
      }
    }
  } else {
    // ===================== element-wise warps =====================
    const int quarter = warp & 3, slice = (warp - 2) >> 2;
    const int lrow = quarter * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t sP = ptx::smem_u32(smem + SM::kP), sDS = ptx::smem_u32(smem + SM::kDS);
    const uint32_t st_out = ptx::smem_u32(smem + SM::kOut) + (warp - 2) * 2048;
    const float sl2 = a.scale_log2, ds = a.drop_scale;
    __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(a.dqkv);
    const long long ld3 = 3ll * H * TDH;

    // per-row scalars and allow bits of the next item are fetched one item ahead.  This thread's 16 key columns
    // of sub-block (jt, jh) are keys jt*128 + jh*64 + slice*16 .. +15: bits (jh*64 + slice*16) & 31 .. of word
    // 4 jt + 2 jh + slice/2 of its allow row; aw_nx[t][jt] holds them for jh = 0 (low half) and jh = 1 (high half)
    // Row scalars (lse, delta) of the next item are fetched one item ahead; the allow / keep words one SUB-BLOCK ahead
    // (two registers in flight instead of sixteen: an item-ahead prefetch of every word spilled, and a spill store
    // right behind a load waits for the load).  Nothing touches a fetched value before the following sub-block / item.
    float lse_raw[2], dlt_raw[2];
    auto fetch_rows = [&](int item_) {
#pragma unroll
      for (int t = 0; t < 2; ++t) { lse_raw[t] = INFINITY; dlt_raw[t] = 0.f; }
      if (item_ >= n_items) return;
      const int b_ = (int)__umulhi((uint32_t)item_, magic_h), h_ = item_ - b_ * H;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int i = t * 128 + lrow;
        if (t < n_t && i < L) {
          lse_raw[t] = __ldg(a.lse + ((size_t)b_ * H + h_) * L + i);
          dlt_raw[t] = __ldg(a.delta + ((size_t)b_ * H + h_) * L + i);
        }
      }
    };
    // this thread's 16 key columns of sub-block (jt, jh) are keys jt*128 + jh*64 + slice*16 .. +15: bits
    // (slice & 1) * 16 .. of word 4 jt + 2 jh + slice / 2 of the allow / keep row of query row it*128 + lrow
    uint32_t aw_nx = 0u, kw_nx = 0xffffffffu;
    auto fetch_bits = [&](int item_, int u_) {
      aw_nx = 0u; kw_nx = 0xffffffffu;
      if (item_ >= n_items) return;
      int jt_, it_, jh_, nh_;
      decode_sub(u_, jt_, it_, jh_, nh_);
      const int i = it_ * 128 + lrow, w = 4 * jt_ + 2 * jh_ + (slice >> 1);
      if (i >= L || w >= a.W) return;
      const int b_ = (int)__umulhi((uint32_t)item_, magic_h), h_ = item_ - b_ * H;
      aw_nx = __ldg(a.allow + (((size_t)b_ * a.Hm + (a.Hm == 1 ? 0 : h_)) * L + i) * a.W + w);
      if (a.keep) kw_nx = __ldg(a.keep + (((size_t)b_ * H + h_) * L + i) * a.W + w);
    };
    fetch_bits(blockIdx.x, 0);
    fetch_rows(blockIdx.x);
    uint32_t cs0 = 0, cs1 = 0, cf = 0, cg = 0;     // s_full[0/1], stage_free completion counters; sub-blocks done
    int iter = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
      const int b = (int)__umulhi((uint32_t)item, magic_h), h = item - b * H;
      const float inv_s = __ldg(a.inv_scale + item);           // item = b H + h: leave the scaled domain at the drains
      float lse2[2], dlt[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) { lse2[t] = lse_raw[t] * kLog2e; dlt[t] = dlt_raw[t]; }
      fetch_rows(item + gridDim.x);
      for (int u = 0; u < n_sub; ++u) {
        int jt, it, jh, nh;
        decode_sub(u, jt, it, jh, nh);
        const int i = it * 128 + lrow;
        const bool active = i < L;
        const bool warp_active = it * 128 + quarter * 32 < L;
        const float my_lse2 = it ? lse2[1] : lse2[0], my_dlt = it ? dlt[1] : dlt[0];
        const uint32_t aw_raw = aw_nx, kw_raw = kw_nx;          // fetched during the previous sub-block
        if (u + 1 < n_sub) fetch_bits(item, u + 1); else fetch_bits(item + gridDim.x, 0);
        const uint32_t awc = active ? ((aw_raw >> ((slice & 1) * 16)) & 0xffffu) : 0u;
        const uint32_t kwc = (kw_raw >> ((slice & 1) * 16)) & 0xffffu;
        if (warp == 2 && lane == 0) TL_STAMP(20, (int)(cs0 + cs1));
        if (u & 1) { ptx::mbar_wait(&s_full[1], cs1 & 1); ++cs1; } else { ptx::mbar_wait(&s_full[0], cs0 & 1); ++cs0; }
        ptx::tc_fence_after();
        if (warp == 2 && lane == 0) TL_STAMP(21, (int)(cs0 + cs1) - 1);
        TL_STAMP_ANY(30 + (warp - 2), (int)(cs0 + cs1) - 1);
        uint32_t pk[8], dk_[8];                               // packed Pd / dS of the 16 columns
        if (warp_active) {
          uint32_t rs[16], rd[16];
          const uint32_t tcol = t_lane + (uint32_t)(u & 1) * 128u + slice * 16;
          ptx::tmem_ld_32x16(tcol, rs);
          ptx::tmem_ld_32x16(tcol + 64, rd);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            float e[8], uu[8], kf[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const float arg = fmaf(__uint_as_float(rs[8 * q + x]), sl2, -my_lse2);
              e[x] = ((awc >> (8 * q + x)) & 1u) ? fast_exp2(arg) : 0.f;
            }
#pragma unroll
            for (int x = 0; x < 8; ++x) kf[x] = ((kwc >> (8 * q + x)) & 1u) ? ds : 0.f;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              uu[x] = e[x] * fmaf(__uint_as_float(rd[8 * q + x]), kf[x], -my_dlt);   // dS / scale
              e[x] *= kf[x];                                                            // Pd
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              pk[4 * q + x] = pack_f16(e[2 * x], e[2 * x + 1]);
              dk_[4 * q + x] = pack_f16_sat(uu[2 * x], uu[2 * x + 1]);
            }
          }
        } else {
#pragma unroll
          for (int x = 0; x < 8; ++x) { pk[x] = 0u; dk_[x] = 0u; }
        }
        // the staging tiles of the previous block are free once its gradient MMAs have retired
        if (warp == 2 && lane == 0) TL_STAMP(22, (int)(cs0 + cs1) - 1);
        if (jh == 0 && cf > 0) ptx::mbar_wait(stage_free, (cf - 1) & 1);
        if (warp == 2 && lane == 0) TL_STAMP(23, (int)(cs0 + cs1) - 1);
        {
          // row lrow, key columns jh*64 + slice*16 .. +15 of the [128 rows][2 x 64 keys] 128B-swizzled tile pair
          const uint32_t base = (uint32_t)jh * 16384u + (uint32_t)lrow * 128u;
          const int ch = slice * 2;                           // 16-byte chunk index within the 128-byte row
          sts128u(sP + base + (((ch) ^ (lrow & 7)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
          sts128u(sP + base + (((ch + 1) ^ (lrow & 7)) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
          sts128u(sDS + base + (((ch) ^ (lrow & 7)) << 4), make_uint4(dk_[0], dk_[1], dk_[2], dk_[3]));
          sts128u(sDS + base + (((ch + 1) ^ (lrow & 7)) << 4), make_uint4(dk_[4], dk_[5], dk_[6], dk_[7]));
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sub_done[cg & 1]);
        ++cg;
        if (warp == 2 && lane == 0) TL_STAMP(24, (int)(cs0 + cs1) - 1);
        TL_STAMP_ANY(46 + (warp - 2), (int)(cs0 + cs1) - 1);
        if (jh == nh - 1) {
          ++cf;                                               // one stage_free completion per block
          if (it == n_t - 1) {
            // ---- key tile finished: drain dK_jt (scaled) and dV_jt.  slices 0,1 -> dK halves, slices 2,3 -> dV halves
            ptx::mbar_wait(stage_free, (cf - 1) & 1);
            ptx::tc_fence_after();
            const int j0 = jt * 128 + quarter * 32;
            if (j0 < L) {
              const bool is_k = slice < 2;
              const int hc = (slice & 1) * 32;
              drain32_bf16(t_lane + (is_k ? kDKcol : kDVcol) + hc, is_k ? a.scale * inv_s : inv_s, st_out,
                           dqkv + (size_t)((is_k ? H : 2 * H) + h) * TDH + hc, ld3, (long long)b * L + j0, 0, min(32, L - j0), lane);
            }
            ptx::tc_fence_before();
            if (warp == 2 && lane == 0) TL_STAMP(25, (int)(cs0 + cs1) - 1);
          }
        }
      }
      // ---- item finished: drain dQ (scaled).  slices 0,1 -> query tile 0 halves, slices 2,3 -> query tile 1
      {
        ptx::tc_fence_after();
        const int t = slice >> 1, hc = (slice & 1) * 32;
        const int i0 = t * 128 + quarter * 32;
        if (t < n_t && i0 < L)
          drain32_bf16(t_lane + kDQcol + t * 64 + hc, a.scale * inv_s, st_out, dqkv + (size_t)h * TDH + hc, ld3, (long long)b * L + i0, 0,
                       min(32, L - i0), lane);
        ptx::tc_fence_before();
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem); }
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
  const bool prep_only = p->bwd_phase == 2;      // the preparation kernel reads dctx and ctx only
  if (!p->allow_bits && !prep_only) { set_error("samk_attn: tensor-core path needs allow_bits (samk_attn_build_mask)"); return SAMK_ERR_ARG; }
  a.ctx = p->ctx; a.lse = p->lse; a.delta = p->delta; a.dqkv = p->dqkv; a.dq_accum = p->dq_accum;
  a.inv_scale = p->do_inv_scale;
  a.allow = p->allow_bits; a.Hm = p->spatial ? p->H : 1;
  a.keep = p->drop_p > 0.f ? p->keep_bits : nullptr;
  if (p->drop_p > 0.f && !p->keep_bits && !prep_only) { set_error("samk_attn: tensor-core path with dropout needs keep_bits (samk_attn_build_keep)"); return SAMK_ERR_ARG; }
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

int sm_count();

template <int KVT>
static int launch_fwd2(const samk_attn_params* p, const TcArgs& a, cudaStream_t stream) {
  const long long rows = (long long)a.B * a.L;
  const int hd3 = 3 * a.H * TDH;
  CUtensorMap tq, tkv;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, p->qkv, rows, hd3, hd3, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tkv, p->qkv, rows, hd3, hd3, 64, KVT))) return rc;
  if ((rc = set_smem(attn_fwd2_kernel<KVT>, Fwd2Cfg<KVT>::kSmem))) return rc;
  const int n_qt = (a.L + 127) / 128;
  const int qp0 = a.q_tile0 / 2;
  const int n_qp = (n_qt + 1) / 2 - qp0;
  if (n_qp <= 0) return SAMK_OK;
  const long long n_items = (long long)a.B * a.H * n_qp;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int grid = (int)(n_items < sms ? n_items : sms);
  if (n_items >= (1ll << 24)) { set_error("samk_attn_fwd: too many work items"); return SAMK_ERR_UNSUPPORTED; }
  const uint32_t magic_qp = (uint32_t)(((1ull << 32) + n_qp - 1) / n_qp), magic_h = (uint32_t)(((1ull << 32) + a.H - 1) / a.H);
  attn_fwd2_kernel<KVT><<<grid, kFwd2Threads, Fwd2Cfg<KVT>::kSmem, stream>>>(tq, tkv, a, n_qp, qp0, (int)n_items, magic_qp,
                                                                            magic_h);
  return check_launch("samk_attn_fwd(tc v2)");
}

template <int KVT>
static int launch_fwd3(const samk_attn_params* p, const TcArgs& a, cudaStream_t stream) {
  const long long rows = (long long)a.B * a.L;
  const int hd3 = 3 * a.H * TDH;
  CUtensorMap tq, tkv;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tq, p->qkv, rows, hd3, hd3, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tkv, p->qkv, rows, hd3, hd3, 64, KVT))) return rc;
  if ((rc = set_smem(attn_fwd3_kernel<KVT>, Fwd3Cfg<KVT>::kSmem))) return rc;
  const int n_qt = (a.L + 127) / 128;
  const int n_bh = a.B * a.H;
  const long long n_items = (long long)(n_qt - a.q_tile0) * n_bh;
  if (n_items <= 0) return SAMK_OK;
  if (n_items >= (1ll << 24)) { set_error("samk_attn_fwd: too many work items"); return SAMK_ERR_UNSUPPORTED; }
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int grid = (int)(n_items < 2 * sms ? n_items : 2 * sms);
  static int order = -1;                       // SAMK_ATTN_ORDER: 1 = chunked (default), 0 = all tiles 0, then all tiles 1, ...
  if (order < 0) { const char* e = getenv("SAMK_ATTN_ORDER"); order = e ? atoi(e) : 1; }
  const uint32_t magic_bh = order == 1 ? 0u : (uint32_t)(((1ull << 32) + n_bh - 1) / n_bh);
  const uint32_t magic_h = (uint32_t)(((1ull << 32) + a.H - 1) / a.H);
  attn_fwd3_kernel<KVT><<<grid, kFwd3Threads, Fwd3Cfg<KVT>::kSmem, stream>>>(tq, tkv, a, n_bh, (int)n_items, magic_bh, magic_h);
  return check_launch("samk_attn_fwd(tc v3)");
}

static int attn_fwd_version() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("SAMK_ATTN_FWD_V");
    v = (s && s[0] >= '2' && s[0] <= '3') ? s[0] - '0' : 0;      // 0 = pick by sequence length
  }
  return v;
}

int attn_tc_fwd(const samk_attn_params* p, cudaStream_t stream) {
  TcArgs a;
  int rc = fill_tc(a, p);
  if (rc) return rc;
  if (!p->qkv || !p->ctx || !p->lse) { set_error("samk_attn_fwd: null pointer"); return SAMK_ERR_ARG; }
  if (!a.B || !a.L) return SAMK_OK;
  // measured (tools/attn_bench.py --sweep): two CTAs per SM win up to ~500 keys (L=118: 31 -> 21 us, 182: 90 -> 80,
  // 268: 162 -> 118), the two-tile CTA with its 3-stage K/V ring wins for long sequences (1036: 958 vs 1028 us)
  const int ver = attn_fwd_version() ? attn_fwd_version() : (a.L <= 512 ? 3 : 2);
  if (ver == 3) {
    if (a.L > 128 && a.L <= 192) return launch_fwd3<192>(p, a, stream);
    return launch_fwd3<128>(p, a, stream);
  }
  if (a.L > 128 && a.L <= 192) return launch_fwd2<192>(p, a, stream);
  return launch_fwd2<128>(p, a, stream);
}

static int attn_bwd_version() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("SAMK_ATTN_BWD_V");
    v = (s && s[0] == '1') ? 1 : 2;
  }
  return v;
}

int attn_tc_bwd(const samk_attn_params* p, cudaStream_t stream) {
  TcArgs a;
  int rc = fill_tc(a, p);
  if (rc) return rc;
  if (!p->qkv || !p->ctx || !p->lse || !p->dctx || !p->dqkv || !p->delta || !p->do_f16 || !p->do_inv_scale) {
    set_error("samk_attn_bwd: null pointer (the tensor-core path needs the do_f16 / do_inv_scale workspaces)");
    return SAMK_ERR_ARG;
  }
  if (!a.B || !a.L) return SAMK_OK;
  const long long rows = (long long)a.B * a.L;
  const int hd = a.H * TDH;
  // bwd_phase: 0 = whole backward, 1 = the preparation already ran on these workspaces, 2 = preparation only (a caller
  // may run it on a second stream, beside the weight-gradient product that sits between the two in a layer's backward)
  if (p->bwd_phase != 1) {
    attn_bwd_prep_kernel<<<a.B * a.H, 256, 0, stream>>>((const __nv_bfloat16*)p->dctx, (const __half*)p->ctx, (__half*)p->do_f16,
                                                       p->delta, p->do_inv_scale, a.H, a.L);
    if ((rc = check_launch("samk_attn_bwd(prep)"))) return rc;
  }
  if (p->bwd_phase == 2) return SAMK_OK;
  CUtensorMap tqkv, tdo;
  if ((rc = make_tmap_bf16_2d(&tqkv, p->qkv, rows, 3 * hd, 3 * hd, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tdo, p->do_f16, rows, hd, hd, 64, 128))) return rc;

  if (a.L <= 256 && attn_bwd_version() == 2) {
    // whole (sample, head) per CTA: dQ, dK, dV complete in TMEM, written once
    CUtensorMap tqkv64;
    if ((rc = make_tmap_bf16_2d(&tqkv64, p->qkv, rows, 3 * hd, 3 * hd, 64, 64))) return rc;
    if ((rc = set_smem(attn_bwd2_kernel, Bwd2Smem::kTotal))) return rc;
    const long long n_items = (long long)a.B * a.H;
    if (n_items >= (1ll << 24)) { set_error("samk_attn_bwd: too many work items"); return SAMK_ERR_UNSUPPORTED; }
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int grid = (int)(n_items < sms ? n_items : sms);
    const uint32_t magic_h = (uint32_t)(((1ull << 32) + a.H - 1) / a.H);
    attn_bwd2_kernel<<<grid, kBwd2Threads, Bwd2Smem::kTotal, stream>>>(tqkv, tqkv64, tdo, a, (int)n_items, magic_h);
    return check_launch("samk_attn_bwd(tc v2)");
  }

  // long sequences: key-tile CTAs streaming query tiles, dQ reduced with fp32 atomics
  if (!p->dq_accum) { set_error("samk_attn_bwd: L > 256 needs dq_accum"); return SAMK_ERR_ARG; }
  if (cudaMemsetAsync(p->dq_accum, 0, (size_t)rows * hd * sizeof(float), stream) != cudaSuccess) {
    set_error("samk_attn_bwd: memset failed");
    return SAMK_ERR_CUDA;
  }
  constexpr int smem = 4 * 16384 + 2 * 32768 + 1024 + 64;
  if ((rc = set_smem(attn_bwd_tc_kernel, smem))) return rc;
  dim3 grid((a.L + 127) / 128, a.H, a.B);
  attn_bwd_tc_kernel<<<grid, 256, smem, stream>>>(tqkv, tdo, a);
  if ((rc = check_launch("samk_attn_bwd(tc)"))) return rc;
  attn_dq_store_kernel<<<(sm_count() > 0 ? sm_count() : 148) * 8, 256, 0, stream>>>(p->dq_accum, (__nv_bfloat16*)p->dqkv, (size_t)rows, hd);
  return check_launch("samk_attn_bwd(dq store)");
}

}  // namespace samk

extern "C" {

#ifdef SAMK_TIMELINE
int samk_debug_timeline(long long* dst, int n) {
  return cudaMemcpyFromSymbol(dst, samk::g_timeline, sizeof(long long) * (size_t)(n < 4096 ? n : 4096)) == cudaSuccess ? 0 : -2;
}
#endif

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
  if (Hm > 16) { set_error("samk_attn_build_mask: at most 16 heads (relation words are 16 bits)"); return SAMK_ERR_UNSUPPORTED; }
  dim3 grid((m.L * W + 127) / 128, p->B);
  attn_build_mask_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(m, p->H, Hm, W, allow_bits);
  return check_launch("samk_attn_build_mask");
}

int samk_attn_build_keep(const samk_attn_params* p, uint32_t* keep_bits, void* stream) {
  using namespace samk;
  if (!p || !keep_bits) { set_error("samk_attn_build_keep: null pointer"); return SAMK_ERR_ARG; }
  const int L = p->T + p->A + p->D, W = (L + 31) / 32;
  const long long n = (long long)p->B * p->H * L * W;
  if (!n) return SAMK_OK;
  const uint32_t thresh = p->drop_p > 0.f ? drop_threshold(p->drop_p) : 0u;
  attn_build_keep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(keep_bits, n, L, W, thresh, p->drop_seed,
                                                                                       p->drop_offset);
  return check_launch("samk_attn_build_keep");
}

int samk_attn_fwd(const samk_attn_params* p, int impl, void* stream) {
  if (!p) { samk::set_error("samk_attn_fwd: null params"); return SAMK_ERR_ARG; }
  if (impl == 0 && p->dtype == SAMK_DT_F16) return samk::attn_tc_fwd(p, (cudaStream_t)stream);
  if (impl == 0 && p->dtype != SAMK_DT_F32) { samk::set_error("samk_attn_fwd: the tensor-core kernels take f16 q|k|v"); return SAMK_ERR_UNSUPPORTED; }
  return samk::attn_simt_fwd(p, (cudaStream_t)stream);
}

int samk_attn_bwd(const samk_attn_params* p, int impl, void* stream) {
  if (!p) { samk::set_error("samk_attn_bwd: null params"); return SAMK_ERR_ARG; }
  if (impl == 0 && p->dtype == SAMK_DT_F16 && p->grad_dtype == SAMK_DT_BF16) return samk::attn_tc_bwd(p, (cudaStream_t)stream);
  if (impl == 0 && p->dtype != SAMK_DT_F32) { samk::set_error("samk_attn_bwd: the tensor-core kernels take f16 q|k|v / ctx and bf16 gradients"); return SAMK_ERR_UNSUPPORTED; }
  return samk::attn_simt_bwd(p, (cudaStream_t)stream);
}

}  // extern "C"

namespace samk { int set_drop_salt_attn_tc(unsigned long long salt, cudaStream_t stream) { return set_drop_salt_tu(salt, stream); } }
namespace samk { int set_drop_salt_dev_attn_tc(const unsigned long long* src, cudaStream_t stream) { return set_drop_salt_from_device_tu(src, stream); } }
