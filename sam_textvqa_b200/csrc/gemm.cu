// tcgen05 GEMM with fused epilogues: C[M,N] = epi(alpha * A[M,K] * B[N,K]^T), 16-bit operands (IEEE half or
// bfloat16, the same for A and B) -> fp32 accumulation.
//
// Persistent, warp-specialised, one CTA per SM (a CTA pair per 256 x 256 tile for the large shapes, see CL below):
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B swizzle, kStages-deep mbarrier ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma, accumulators in TMEM,
//                               two accumulator buffers so the epilogue overlaps the next tile)
//   warps 2..   epilogue       (8 or 16 warps: tcgen05.ld TMEM -> registers -> shared-memory transpose -> bias / GELU /
//                               dropout / residual -> coalesced global stores; warp w owns TMEM lanes 32*(w%4)..+31 and
//                               one column part of the tile)
// Tile 128 x BN x 64 (BN = 128 or 256).  Either operand may be K-major (row-major [rows,K]) or
// MN-major (row-major [K,rows]); the second form serves dgrad (B = W as stored) and wgrad
// (A = dY, B = X, contraction over tokens) without materialising transposes.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/samk.h"

namespace samk {

constexpr int BM = 128;
constexpr int BK = 64;
// Epilogue warps per CTA (template parameter EW): 8 = two warps per TMEM lane quarter, each draining half of the tile's
// columns; 16 = four per quarter, a quarter of the columns each.  The epilogue of the K = 768 shapes is longer than their
// main loop and latency-bound (two warps per scheduler), so those shapes take 16 warps and give up one pipeline stage
// for the extra transpose tiles.
constexpr int gemm_threads(int ew) { return 64 + 32 * ew; }

// Epilogue feature bits.  A kernel instantiation carries either a compile-time mask (EPI >= 0: the hot
// combinations of the SA-M4C layers, dead branches removed -- a fully dynamic epilogue is ~100 KB of
// SASS and stalls on instruction fetch) or EPI_DYNAMIC (runtime mask, any combination / odd sizes).
enum : int {
  F_OUT_BF16 = 1 << 0,    // out is bf16 (else fp32)
  F_BIAS = 1 << 1,
  F_DROP = 1 << 2,
  F_RES = 1 << 3,         // += residual (fp32)
  F_GELU_PAIR = 1 << 4,   // act 3: out = gelu(v), pre = gelu'(v)
  F_PRE = 1 << 5,         // pre-activation copy (act != 3)
  F_PRE_BF16 = 1 << 6,
  F_MULAUX = 1 << 7,      // act 4: v *= aux
  F_DGELU = 1 << 8,       // act 2: v *= gelu'(aux)
  F_AUX_BF16 = 1 << 9,
  F_GELU = 1 << 10,       // act 1
  F_ATOMIC = 1 << 11,     // red.global.add into fp32 out
  F_ALPHA = 1 << 12,      // alpha != 1
  F_VEC = 1 << 13,        // all pointers / pitches vector friendly and N % 8 == 0: no scalar tails
  // the 16-bit buffers are IEEE half instead of bfloat16 (F_OUT_BF16 / F_PRE_BF16 / F_AUX_BF16 then mean "16-bit")
  F_OUT_F16 = 1 << 14,
  F_PRE_F16 = 1 << 15,
  F_AUX_F16 = 1 << 16,
  // fp32 output whose rows are only 8-byte aligned (pitch % 2 == 0, e.g. the [B*D, V+R = 5050] score buffer the classifier
  // writes its V columns into): everything else as F_VEC, the output goes out as four 8-byte stores per 8 columns
  F_OUT_V2 = 1 << 17,
};
constexpr int EPI_DYNAMIC = -1;

struct EpiArgs {
  void* out; long long ldo; float alpha;
  const float* alpha_dev;   // optional device scalar multiplied into alpha (1 / scale of an f16-scaled gradient operand)
  const float* bias;
  void* pre; long long ldpre;
  const void* aux; long long ldaux;
  uint32_t drop_thresh; float drop_scale; unsigned long long seed, offset;
  const float* residual; long long ldres;
  int flags;
  int part_rows; void* out1; void* out2;   // F_ATOMIC: row groups with separate destinations (0 = single output)
};

template <int EPI> __device__ __forceinline__ int epi_flags(const EpiArgs& ep) {
  if constexpr (EPI >= 0) return EPI; else return ep.flags;
}

// The epilogue of 8 consecutive columns [col, col+cnt) (cnt <= 8, col % 8 == 0) of one output row, split into
// load / compute / store phases so that a caller handling several rows can issue all global loads
// first (memory-level parallelism), then the math, then the stores.
struct EpiRow {
  float v[8];      // accumulator values -> final values
};
struct EpiIn {       // operands fetched from global memory ahead of the accumulator (kept raw: no use before compute)
  float res[8];    // residual
  float aux[8];    // aux operand (act 2 / 4), fp32 form
  uint4 auxp;      // aux operand, packed 16-bit form (vector path)
};

template <int EPI>
__device__ __forceinline__ void epi_load(const EpiArgs& ep, EpiIn& e, int row, int col, int cnt) {
  const int F = epi_flags<EPI>(ep);
  const bool full = (F & F_VEC) && cnt == 8;
  if (F & F_RES) {
    const float* p = ep.residual + (size_t)row * ep.ldres + col;
    if (full) {
      const float4 b0 = *reinterpret_cast<const float4*>(p), b1 = *reinterpret_cast<const float4*>(p + 4);
      e.res[0] = b0.x; e.res[1] = b0.y; e.res[2] = b0.z; e.res[3] = b0.w;
      e.res[4] = b1.x; e.res[5] = b1.y; e.res[6] = b1.z; e.res[7] = b1.w;
    } else {
      _Pragma("unroll") for (int i = 0; i < 8; ++i) e.res[i] = i < cnt ? p[i] : 0.f;
    }
  }
  if (F & (F_MULAUX | F_DGELU)) {
    if (F & F_AUX_BF16) {
      const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + (size_t)row * ep.ldaux + col;
      if (full) {
        e.auxp = *reinterpret_cast<const uint4*>(p);
      } else {
        _Pragma("unroll") for (int i = 0; i < 8; ++i) e.aux[i] = i < cnt ? ld_16(p + i, (F & F_AUX_F16) != 0) : 0.f;
      }
    } else {
      const float* p = reinterpret_cast<const float*>(ep.aux) + (size_t)row * ep.ldaux + col;
      _Pragma("unroll") for (int i = 0; i < 8; ++i) e.aux[i] = i < cnt ? p[i] : 0.f;
    }
  }
}

// w: second output (pre-activation copy, or gelu' when act == 3)
template <int EPI>
__device__ __forceinline__ void epi_compute(const EpiArgs& ep, EpiRow& e, const EpiIn& in, float* w, const float* bias8,
                                            int row, int col, int cnt, int N, float alpha) {
  const int F = epi_flags<EPI>(ep);
  const bool full = (F & F_VEC) && cnt == 8;
  float* v = e.v;
  if (F & F_ALPHA) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= alpha;
  }
  if (F & F_BIAS) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += bias8[i];
  }
  if (F & F_GELU_PAIR) {
    // out = gelu(v), pre = gelu'(v): one erf and one exp serve both (backward then only multiplies)
#pragma unroll
    for (int i = 0; i < 8; i += 2) gelu_pair2(v[i], v[i + 1], v[i], v[i + 1], w[i], w[i + 1]);
  } else if (F & F_PRE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = v[i];
  }
  if (F & F_GELU) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) { float d0, d1; gelu_pair2(v[i], v[i + 1], v[i], v[i + 1], d0, d1); }
  } else if (F & (F_MULAUX | F_DGELU)) {
    float x[8];
    if ((F & F_AUX_BF16) && full) {
      const uint32_t* h = reinterpret_cast<const uint32_t*>(&in.auxp);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_16(h[j], (F & F_AUX_F16) != 0);
        x[2 * j] = f.x; x[2 * j + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = in.aux[i];
    }
    if (F & F_MULAUX) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= x[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= dgelu_erf(x[i]);
    }
  }
  if (F & F_DROP) {
    // element (row, c) belongs to the 4-group e4 = row * ceil(N/4) + c/4 of this dropout stream (the same
    // index the LayerNorm-backward kernel regenerates the mask from); 8-groups are pairs of 4-groups
    const uint64_t e4 = (uint64_t)row * (uint64_t)((N + 3) >> 2) + (uint64_t)(col >> 2);
    if (full && ((F & F_VEC) || !(e4 & 1))) {        // F_VEC: N % 8 == 0, so e4 is even
      dropout_apply8(v, ep.seed, ep.offset, e4 >> 1, ep.drop_thresh, ep.drop_scale);
    } else {
      dropout_apply4(v, ep.seed, ep.offset, e4, ep.drop_thresh, ep.drop_scale);
      if (cnt > 4) dropout_apply4(v + 4, ep.seed, ep.offset, e4 + 1, ep.drop_thresh, ep.drop_scale);
    }
  }
  if (F & F_RES) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += in.res[i];
  }
}

template <int EPI>
__device__ __forceinline__ void epi_store(const EpiArgs& ep, const EpiRow& e, const float* w, int row, int col, int cnt) {
  const int F = epi_flags<EPI>(ep);
  const bool full = (F & F_VEC) && cnt == 8;
  const float* v = e.v;
  if (F & (F_PRE | F_GELU_PAIR)) {
    if (F & F_PRE_BF16) {
      __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(ep.pre) + (size_t)row * ep.ldpre + col;
      const bool hf = (F & F_PRE_F16) != 0;
      if (full) {
        *reinterpret_cast<uint4*>(p) = make_uint4(pack_16(w[0], w[1], hf), pack_16(w[2], w[3], hf),
                                                  pack_16(w[4], w[5], hf), pack_16(w[6], w[7], hf));
      } else {
        _Pragma("unroll") for (int i = 0; i < 8; ++i) if (i < cnt) st_16(p + i, w[i], hf);
      }
    } else {
      float* p = reinterpret_cast<float*>(ep.pre) + (size_t)row * ep.ldpre + col;
      if (full) {
        *reinterpret_cast<float4*>(p) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(w[4], w[5], w[6], w[7]);
      } else {
        _Pragma("unroll") for (int i = 0; i < 8; ++i) if (i < cnt) p[i] = w[i];
      }
    }
  }
  if (F & F_ATOMIC) {
    void* base = ep.out;
    int prow = row;
    if (ep.part_rows > 0) {
      const int part = row / ep.part_rows;
      prow = row - part * ep.part_rows;
      base = part == 0 ? ep.out : (part == 1 ? ep.out1 : ep.out2);
    }
    float* p = reinterpret_cast<float*>(base) + (size_t)prow * ep.ldo + col;
    if (full && (F & F_OUT_V2)) {
      _Pragma("unroll") for (int i = 0; i < 8; i += 2)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p + i), "f"(v[i]), "f"(v[i + 1]) : "memory");
    } else if (full) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
    } else {
      _Pragma("unroll") for (int i = 0; i < 8; ++i) if (i < cnt) atomicAdd(p + i, v[i]);
    }
  } else if (F & F_OUT_BF16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + col;
    const bool hf = (F & F_OUT_F16) != 0;
    if (full) {
      *reinterpret_cast<uint4*>(p) = make_uint4(pack_16(v[0], v[1], hf), pack_16(v[2], v[3], hf),
                                                pack_16(v[4], v[5], hf), pack_16(v[6], v[7], hf));
    } else {
      _Pragma("unroll") for (int i = 0; i < 8; ++i) if (i < cnt) st_16(p + i, v[i], hf);
    }
  } else {
    float* p = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col;
    if (full && (F & F_OUT_V2)) {
      _Pragma("unroll") for (int i = 0; i < 8; i += 2) *reinterpret_cast<float2*>(p + i) = make_float2(v[i], v[i + 1]);
    } else if (full) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      _Pragma("unroll") for (int i = 0; i < 8; ++i) if (i < cnt) p[i] = v[i];
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epi_load_bias(const EpiArgs& ep, float* bias8, int col, int cnt) {
  const int F = epi_flags<EPI>(ep);
  if (!(F & F_BIAS)) return;
  if ((F & F_VEC) && cnt == 8) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + col + 4));
    bias8[0] = b0.x; bias8[1] = b0.y; bias8[2] = b0.z; bias8[3] = b0.w;
    bias8[4] = b1.x; bias8[5] = b1.y; bias8[6] = b1.z; bias8[7] = b1.w;
  } else {
    _Pragma("unroll") for (int i = 0; i < 8; ++i) bias8[i] = i < cnt ? ep.bias[col + i] : 0.f;
  }
}

// whole epilogue for one row x 8 columns (SIMT cross-check kernel)
__device__ __forceinline__ void epilogue_8(const EpiArgs& ep, const float* acc, int row, int col, int cnt, int N) {
  EpiRow e;
  EpiIn in;
  float bias8[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) e.v[i] = acc[i];
  epi_load_bias<EPI_DYNAMIC>(ep, bias8, col, cnt);
  epi_load<EPI_DYNAMIC>(ep, in, row, col, cnt);
  epi_compute<EPI_DYNAMIC>(ep, e, in, w, bias8, row, col, cnt, N, ep.alpha_dev ? ep.alpha * *ep.alpha_dev : ep.alpha);
  epi_store<EPI_DYNAMIC>(ep, e, w, row, col, cnt);
}

__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// One warp's 32 rows x 32 columns of accumulator, as delivered by tcgen05.ld (thread = row, 32 consecutive
// columns in registers).
//   STAGED = true : transposed through a warp-private 4 KB shared-memory tile (16-byte slots XOR-swizzled by
//     row, conflict-free both ways) so that the epilogue math and every global access run in a coalesced
//     layout: lane = (row sub-index lane/4, 8-column group lane%4); one warp-wide access covers 8 rows x
//     128 contiguous bytes.  Costs shared-memory bandwidth, which the MMA operand reads also need.
//   STAGED = false: each thread finishes its own row (4 groups of 8 columns): no shared-memory traffic,
//     but a warp-wide access touches 32 different rows.
constexpr int kStageBytesPerWarp = 32 * 32 * 4;

// (row, first column) of the `it`-th 8-column group this lane finishes within a 32x32 warp chunk
template <bool STAGED>
__device__ __forceinline__ void chunk_coord(int lane, int it, int row0, int col0, int& row, int& col) {
  if constexpr (STAGED) { row = row0 + it * 8 + (lane >> 2); col = col0 + (lane & 3) * 8; }
  else { row = row0 + lane; col = col0 + it * 8; }
}

template <bool STAGED> struct ChunkIn {
  EpiIn in[4];
  float bias[STAGED ? 1 : 4][8];
};

// issue the global loads (bias, residual, aux) of one chunk; they are consumed one chunk later
template <int EPI, bool STAGED>
__device__ __forceinline__ void chunk_prefetch(const EpiArgs& ep, ChunkIn<STAGED>& c, int row0, int col0, int M, int N, int lane) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    int row, col;
    chunk_coord<STAGED>(lane, it, row0, col0, row, col);
    if (col < N) {
      if (!STAGED || it == 0) epi_load_bias<EPI>(ep, c.bias[STAGED ? 0 : it], col, min(8, N - col));
      if (row < M) epi_load<EPI>(ep, c.in[it], row, col, min(8, N - col));
    }
  }
}

template <int EPI, bool STAGED>
__device__ __forceinline__ void chunk_finish(const EpiArgs& ep, const ChunkIn<STAGED>& c, uint32_t stage, const uint32_t (&r)[32],
                                             bool zero, int row0, int col0, int M, int N, int lane, float alpha) {
  EpiRow e[4];
  if constexpr (STAGED) {
    const uint32_t srow = stage + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts128(srow + ((j ^ (lane & 7)) << 4),
             zero ? make_float4(0.f, 0.f, 0.f, 0.f)
                  : make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                __uint_as_float(r[4 * j + 3])));
    __syncwarp();
    const int cg = lane & 3, sub = lane >> 2;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = it * 8 + sub;
      const uint32_t base = stage + rr * 128;
      const float4 x0 = lds128(base + (((2 * cg) ^ (rr & 7)) << 4)), x1 = lds128(base + (((2 * cg + 1) ^ (rr & 7)) << 4));
      e[it].v[0] = x0.x; e[it].v[1] = x0.y; e[it].v[2] = x0.z; e[it].v[3] = x0.w;
      e[it].v[4] = x1.x; e[it].v[5] = x1.y; e[it].v[6] = x1.z; e[it].v[7] = x1.w;
    }
    __syncwarp();          // staging tile may be overwritten by the next chunk from here on
  } else {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) e[it].v[i] = zero ? 0.f : __uint_as_float(r[8 * it + i]);
    }
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    int row, col;
    chunk_coord<STAGED>(lane, it, row0, col0, row, col);
    if (row < M && col < N) {
      float w[8];
      const int cnt = min(8, N - col);
      epi_compute<EPI>(ep, e[it], c.in[it], w, c.bias[STAGED ? 0 : it], row, col, cnt, N, alpha);
      epi_store<EPI>(ep, e[it], w, row, col, cnt);
    }
  }
}

template <int BN, int CL = 1, int EW = 8> struct GemmCfg {
  static constexpr int kStages = ((BN == 256 && CL == 1) ? 4 : 6) - (EW > 8 ? 1 : 0);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / CL) * BK * 2;      // CTA pair: each CTA holds half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingOff = kStages * kStageBytes + 256;       // after the barriers
  static constexpr int kSmemBytes = kStagingOff + EW * kStageBytesPerWarp /*epilogue transposes*/ + 1024 /*align slack*/;
  static constexpr int kTmemCols = 2 * BN;  // 256 or 512: power of two
};

// CL = 1: one CTA per 128 x BN tile.
// CL = 2: CTA pair (tcgen05 cta_group::2): a cluster of two CTAs owns a 256 x BN tile.  Each CTA loads its own 128
//   rows of A and HALF of the B tile; the leader CTA issues one M=256 MMA that reads both shared memories and writes
//   the accumulator rows of each CTA into that CTA's TMEM.  Per MMA cycle every SM then reads 32 KB instead of 48 KB
//   of operands from shared memory (the port the staged epilogue also needs).  All TMA loads of a stage complete on
//   the leader's full barrier; MMA completion is multicast to both CTAs' empty / accumulator-full barriers; both
//   CTAs' epilogue warps release the accumulator on the leader's barrier.
template <int BN, int A_MN, int B_MN, int CL, int EPI, bool STAGED, int EW = 8>
__global__ void __launch_bounds__(gemm_threads(EW), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const EpiArgs ep, int M, int N, int K, int split_k, uint32_t ab_fmt) {
  using Cfg = GemmCfg<BN, CL, EW>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + BM - 1) / BM;
  const int m_groups = (m_tiles + CL - 1) / CL;     // a cluster takes CL vertically adjacent tiles
  const int n_tiles = (N + BN - 1) / BN;
  const int kb_total = (K + BK - 1) / BK;
  const int kb_per = (kb_total + split_k - 1) / split_k;
  const int num_work = m_groups * n_tiles * split_k;
  const int cta_rank = CL > 1 ? (int)ptx::cluster_ctarank() : 0;
  const int work0 = blockIdx.x / CL, work_stride = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], CL * EW); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    if (CL == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    else ptx::tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  if (CL > 1) ptx::cluster_sync_all();   // peer barriers must be initialised before remote arrivals / multicasts
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // everything above overlapped the previous kernel's tail (programmatic dependent launch)
  pdl_release();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = work0; w < num_work; w += work_stride) {
        const int split = w % split_k;
        const int tile = w / split_k;
        const int m0 = ((tile / n_tiles) * CL + cta_rank) * BM, n0 = (tile % n_tiles) * BN;
        const int kb0 = split * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (CL == 1) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (A_MN) {
#pragma unroll
              for (int g = 0; g < BM / 64; ++g) ptx::tma_load_2d(sa + g * (BK * 128), &tmA, &full_bar[stage], m0 + g * 64, kb * BK);
            } else {
              ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int g = 0; g < BN / 64; ++g) ptx::tma_load_2d(sb + g * (BK * 128), &tmB, &full_bar[stage], n0 + g * 64, kb * BK);
            } else {
              ptx::tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
            }
          } else {
            // the leader's barrier collects the bytes of both CTAs (its own expect_tx may come after the peer's
            // first bytes: the transaction count is signed)
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            if (A_MN) {
#pragma unroll
              for (int g = 0; g < BM / 64; ++g) ptx::tma_load_2d_2sm(sa + g * (BK * 128), &tmA, &full_bar[stage], m0 + g * 64, kb * BK);
            } else {
              ptx::tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BK, m0);
            }
            const int nh = n0 + cta_rank * (BN / 2);       // this CTA's half of the B tile
            if (B_MN) {
#pragma unroll
              for (int g = 0; g < BN / 128; ++g) ptx::tma_load_2d_2sm(sb + g * (BK * 128), &tmB, &full_bar[stage], nh + g * 64, kb * BK);
            } else {
              ptx::tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * BK, nh);
            }
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (CTA pair: the leader CTA only) =====================
    // ab_fmt: operand formats (bit 0: A is bf16, bit 1: B is bf16; 0 = IEEE half) -- a kernel argument, so uniform
    const uint32_t idesc = ptx::make_idesc_16(BM * CL, BN, A_MN, B_MN, ab_fmt & 1u, (ab_fmt >> 1) & 1u);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    if (CL == 1 || cta_rank == 0) {
      for (int w = work0; w < num_work; w += work_stride) {
        const int split = w % split_k;
        const int kb0 = split * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          // descriptors are computed by the whole (converged) warp so that they live in uniform registers; only
          // the tcgen05 instructions themselves sit under the elected-lane predicate (an `if (lane == 0)` around
          // the arithmetic makes ptxas wrap every UTCHMMA in an elect / R2UR.BROADCAST / branch loop, ~150 cycles)
          {
            const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
            const uint32_t sb = sa + Cfg::kABytes;
            const uint64_t adesc0 = A_MN ? ptx::make_smem_desc_sw128(sa, BK * 128, 1024) : ptx::make_smem_desc_sw128(sa, 16, 1024);
            const uint64_t bdesc0 = B_MN ? ptx::make_smem_desc_sw128(sb, BK * 128, 1024) : ptx::make_smem_desc_sw128(sb, 16, 1024);
            const bool leader = ptx::elect_one();
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advancing the start-address field (bytes >> 4) cannot carry out of its 14 bits for shared memory
              const uint64_t adesc = adesc0 + (uint64_t)(A_MN ? k * (2048 >> 4) : k * (32 >> 4));
              const uint64_t bdesc = bdesc0 + (uint64_t)(B_MN ? k * (2048 >> 4) : k * (32 >> 4));
              if (leader) {
                if (CL == 1) ptx::umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                else ptx::umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              }
            }
            if (leader) {
              if (CL == 1) {
                ptx::umma_commit(&empty_bar[stage]);                    // smem slot free once these MMAs retire
                if (kb == kb1 - 1) ptx::umma_commit(&tfull_bar[acc]);   // accumulator ready
              } else {
                ptx::umma_commit_2sm(&empty_bar[stage], 0b11);          // ... in both CTAs of the pair
                if (kb == kb1 - 1) ptx::umma_commit_2sm(&tfull_bar[acc], 0b11);
              }
            }
          }
          __syncwarp();
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0 && ptx::elect_one()) {                           // empty K range: still hand over
          if (CL == 1) ptx::umma_commit(&tfull_bar[acc]); else ptx::umma_commit_2sm(&tfull_bar[acc], 0b11);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int lane_grp = warp & 3;          // TMEM lane quarter this warp may access (warp id % 4)
    const int col_half = (warp - 2) >> 2;   // which part (1 / (EW/4)) of the tile's columns this warp drains
    const float alpha = ((epi_flags<EPI>(ep) & F_ALPHA) && ep.alpha_dev) ? ep.alpha * __ldg(ep.alpha_dev) : ep.alpha;
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = work0; w < num_work; w += work_stride) {
      const int split = w % split_k;
      const int tile = w / split_k;
      const int m0 = ((tile / n_tiles) * CL + cta_rank) * BM, n0 = (tile % n_tiles) * BN;
      const int kb0 = split * kb_per;
      const bool has_k = min(kb_total, kb0 + kb_per) > kb0;
      const int row0 = m0 + lane_grp * 32;
      const uint32_t stage = ptx::smem_u32(smem + Cfg::kStagingOff) + (warp - 2) * kStageBytesPerWarp;
      constexpr int kChunks = BN / (8 * EW);         // 32-column chunks per warp
      const int cbase = col_half * kChunks;
      const bool live = (has_k || !(epi_flags<EPI>(ep) & F_ATOMIC)) && row0 < M;
      // operands of chunk i+1 are fetched from global memory while chunk i is finished (the first one even
      // before the accumulator is ready)
      // (only pays off for the aux operand of the fused dgrad; measured slower for residual-only epilogues,
      // and the runtime-flag epilogue has no registers to spare)
      constexpr bool kPipe = EPI >= 0 && (EPI & (F_MULAUX | F_DGELU)) != 0;
      ChunkIn<STAGED> cin[kPipe ? 2 : 1];
      if (kPipe && live && n0 + cbase * 32 < N) chunk_prefetch<EPI, STAGED>(ep, cin[0], row0, n0 + cbase * 32, M, N, lane);
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(lane_grp * 32) << 16);
      if (live) {
        if constexpr (kPipe) {
#pragma unroll
          for (int i = 0; i < kChunks; ++i) {
            const int col0 = n0 + (cbase + i) * 32;
            if (col0 < N) {  // warp-uniform
              if (i + 1 < kChunks && col0 + 32 < N) chunk_prefetch<EPI, STAGED>(ep, cin[(i + 1) & 1], row0, col0 + 32, M, N, lane);
              uint32_t r[32];
              ptx::tmem_ld_32x32(taddr + (cbase + i) * 32, r);
              ptx::tmem_ld_wait();
              chunk_finish<EPI, STAGED>(ep, cin[i & 1], stage, r, !has_k, row0, col0, M, N, lane, alpha);
            }
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < kChunks; ++i) {
            const int col0 = n0 + (cbase + i) * 32;
            if (col0 >= N) break;  // warp-uniform
            chunk_prefetch<EPI, STAGED>(ep, cin[0], row0, col0, M, N, lane);
            uint32_t r[32];
            ptx::tmem_ld_32x32(taddr + (cbase + i) * 32, r);
            ptx::tmem_ld_wait();
            chunk_finish<EPI, STAGED>(ep, cin[0], stage, r, !has_k, row0, col0, M, N, lane, alpha);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 1) ptx::mbar_arrive(&tempty_bar[acc]);
        else ptx::mbar_arrive_leader(&tempty_bar[acc]);       // the leader's MMA warp waits for both CTAs' epilogues
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  if (CL > 1) ptx::cluster_sync_all();   // no CTA may exit while its peer can still signal its barriers
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (CL == 1) ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    else ptx::tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT reference kernel (unit tests only: cross-checks the tensor-core kernel on the device).
// ---------------------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, int a_mn, long long lda,
                                 const __nv_bfloat16* __restrict__ B, int b_mn, long long ldb,
                                 const EpiArgs ep, int M, int N, int K, uint32_t ab_fmt) {
  const bool a_f16 = !(ab_fmt & 1u), b_f16 = !(ab_fmt & 2u);
  // one thread = one row x 32-column chunk, so the epilogue code path is shared
  const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunks_per_row = (N + 31) / 32;
  const int row = chunk / chunks_per_row;
  if (row >= M) return;
  const int col0 = (chunk % chunks_per_row) * 32;
  const int cnt = min(32, N - col0);
  float v[32];
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
  for (int k = 0; k < K; ++k) {
    float a = ld_16(a_mn ? &A[(size_t)k * lda + row] : &A[(size_t)row * lda + k], a_f16);
    for (int i = 0; i < cnt; ++i) {
      float b = ld_16(b_mn ? &B[(size_t)k * ldb + col0 + i] : &B[(size_t)(col0 + i) * ldb + k], b_f16);
      v[i] = fmaf(a, b, v[i]);
    }
  }
  EpiArgs e2 = ep;
  e2.flags &= ~F_VEC;
  for (int c = 0; c < cnt; c += 8) epilogue_8(e2, v + c, row, col0 + c, min(8, cnt - c), N);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] array with pitch ld (elements), 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld,
                      int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not found"); return SAMK_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) base=%p rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r, base, rows,
              cols, ld, box_cols, box_rows);
    return SAMK_ERR_CUDA;
  }
  return SAMK_OK;
}

static int g_sm_count = 0;
static int g_sm_reserved = 0;
int sm_count_physical() {
  if (!g_sm_count) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    g_sm_count = n;
  }
  return g_sm_count;
}
// SMs the persistent kernels (GEMM, attention) size their grids for: the device's SMs minus the ones reserved for a
// concurrently running collective (samk_reserve_sms).  A persistent kernel has a STATIC tile schedule and one CTA per
// SM: a CTA that cannot become resident because an NCCL CTA holds its SM starts only when a whole wave has finished
// and doubles the kernel's duration; leaving those SMs out of the grid costs reserved/148 instead.
int sm_count() {
  const int n = sm_count_physical();
  if (n <= 0) return n;
  int r = g_sm_reserved;
  if (r < 0) r = 0;
  if (r > n - 2) r = n - 2;
  return (n - r) & ~1;        // CTA pairs: keep it even
}
void set_sm_reserved(int n) { g_sm_reserved = n; }

template <int BN, int A_MN, int B_MN, int CL, int EPI, bool STAGED, int EW = 8>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const EpiArgs& ep, int M, int N, int K, int split_k,
                     uint32_t ab_fmt, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CL, EW>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CL, EPI, STAGED, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes, cudaGetErrorString(cudaGetLastError()));
      return SAMK_ERR_CUDA;
    }
    attr_set = true;
  }
  const int m_groups = ((M + BM - 1) / BM + CL - 1) / CL, n_tiles = (N + BN - 1) / BN;
  const int work = m_groups * n_tiles * split_k;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int clusters = sms / CL;
  if (work < clusters) clusters = work;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(gemm_threads(EW));
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_level() >= 1 ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, ep, M, N, K, split_k, ab_fmt);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("samk_gemm: launch failed: %s", cudaGetErrorString(e));
    return SAMK_ERR_CUDA;
  }
  return check_launch("samk_gemm");
}

// CTA-pair tiles (cta_group::2): SAMK_GEMM_2CTA=0/1 overrides the default
static int g_gemm_2cta = -1;
static int gemm_2cta() {
  if (g_gemm_2cta < 0) {
    const char* s = getenv("SAMK_GEMM_2CTA");
    g_gemm_2cta = s ? (s[0] == '1' ? 1 : 0) : 1;     // measured faster on every layer shape (tools/gemm_bench.py)
  }
  return g_gemm_2cta;
}

// Which specialised epilogues run with 16 epilogue warps on CTA-pair tiles: bit i = entry i of kSpecs (samk_gemm_16).
// SAMK_GEMM_EW16=<mask> overrides the default (measured per shape with tools/gemm_bench.py).
constexpr int kDefaultEw16Mask = 0x45;   // q|k|v projection, FFN1 (train: GELU pair, inference: GELU)
static int g_gemm_ew16 = -1;
static int gemm_ew16_mask() {
  if (g_gemm_ew16 < 0) {
    const char* s = getenv("SAMK_GEMM_EW16");
    g_gemm_ew16 = s ? (int)strtol(s, nullptr, 0) : kDefaultEw16Mask;
  }
  return g_gemm_ew16;
}

// the epilogue combinations of the SA-M4C layers that get a compile-time specialised kernel
// (forward activations are stored as IEEE half, gradients as bfloat16: see ops.py "Precision modes")
constexpr int M_QKV = F_OUT_BF16 | F_OUT_F16 | F_BIAS | F_VEC;                            // fused q|k|v projection
constexpr int M_OUTPROJ = F_BIAS | F_DROP | F_RES | F_VEC;                                // W_o / FFN2 forward (train)
constexpr int M_FFN1 = F_OUT_BF16 | F_OUT_F16 | F_BIAS | F_GELU_PAIR | F_PRE_BF16 | F_VEC;   // FFN1 forward (train): gelu half, gelu' bf16
constexpr int M_BF16 = F_OUT_BF16 | F_VEC;                                                // W_o dgrad
constexpr int M_F32_BIAS = F_BIAS | F_VEC;                                                // input projections, heads
constexpr int M_F32_BIAS_RES = F_BIAS | F_RES | F_VEC;                                    // inference W_o / FFN2
constexpr int M_GELU = F_OUT_BF16 | F_OUT_F16 | F_BIAS | F_GELU | F_VEC;                  // inference FFN1
constexpr int M_MULAUX = F_OUT_BF16 | F_MULAUX | F_AUX_BF16 | F_VEC;                      // FFN2 dgrad * gelu' (bf16: unpacked by a shift)
constexpr int M_F32_RES = F_RES | F_VEC;                                                  // FFN1 / qkv dgrad + skip grad
constexpr int M_F32 = F_VEC;
constexpr int M_F32_BIAS_V2 = F_BIAS | F_VEC | F_OUT_V2;                                  // classifier columns of the score buffer
constexpr int M_ATOMIC = F_ATOMIC | F_ALPHA | F_VEC;                                      // wgrad accumulation (alpha: 1 / gradient scale)
constexpr int M_ATOMIC_V2 = M_ATOMIC | F_OUT_V2;                                          // ... into a gradient whose rows are 8-byte aligned (OCR projection: pitch 3002)

}  // namespace samk

extern "C" int samk_gemm_bf16(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, int M,
                              int N, int K, const samk_gemm_epilogue* e, int split_k, int impl, void* stream_) {
  return samk_gemm_16(A, SAMK_DT_BF16, a_mn, lda, B, SAMK_DT_BF16, b_mn, ldb, M, N, K, e, split_k, impl, stream_);
}

extern "C" int samk_gemm_16(const void* A, int a_dtype, int a_mn, long long lda, const void* B, int b_dtype, int b_mn,
                            long long ldb, int M, int N, int K, const samk_gemm_epilogue* e, int split_k, int impl,
                            void* stream_) {
  using namespace samk;
  cudaStream_t stream = (cudaStream_t)stream_;
  auto is16 = [](int dt) { return dt == SAMK_DT_BF16 || dt == SAMK_DT_F16; };
  if (!is16(a_dtype) || !is16(b_dtype)) { set_error("samk_gemm_16: operands must be bf16 or f16"); return SAMK_ERR_ARG; }
  if (a_dtype != b_dtype) {
    // measured on B200: tcgen05.mma kind::f16 with a_format != b_format raises an illegal-instruction fault
    set_error("samk_gemm_16: A and B must share one 16-bit format (mixed f16 x bf16 products fault on sm_100a)");
    return SAMK_ERR_UNSUPPORTED;
  }
  const uint32_t ab_fmt = (a_dtype == SAMK_DT_BF16 ? 1u : 0u) | (b_dtype == SAMK_DT_BF16 ? 2u : 0u);
  if (!A || !B || !e || !e->out) { set_error("samk_gemm: null pointer"); return SAMK_ERR_ARG; }
  if (M < 0 || N < 0 || K < 0) { set_error("samk_gemm: negative size"); return SAMK_ERR_ARG; }
  if (M == 0 || N == 0) return SAMK_OK;
  if (split_k < 1) split_k = 1;
  if (split_k > 1 && !e->atomic_add) { set_error("samk_gemm: split_k>1 needs atomic_add"); return SAMK_ERR_ARG; }
  if (e->atomic_add && e->out_dtype != SAMK_DT_F32) { set_error("samk_gemm: atomic_add needs fp32 out"); return SAMK_ERR_ARG; }
  if ((lda % 8) || (ldb % 8) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) {
    set_error("samk_gemm: operands need 16-byte aligned base and ld %% 8 == 0 (lda=%lld ldb=%lld)", lda, ldb);
    return SAMK_ERR_ARG;
  }
  if ((e->act == 2 || e->act == 4) && !e->aux) { set_error("samk_gemm: act=2/4 needs aux"); return SAMK_ERR_ARG; }
  if (e->act == 3 && !e->pre) { set_error("samk_gemm: act=3 needs pre"); return SAMK_ERR_ARG; }
  if (e->act < 0 || e->act > 4) { set_error("samk_gemm: unknown act %d", e->act); return SAMK_ERR_ARG; }
  EpiArgs ep;
  ep.out = e->out; ep.ldo = e->ldo; ep.alpha = e->alpha; ep.alpha_dev = e->alpha_dev; ep.bias = e->bias;
  ep.pre = e->pre; ep.ldpre = e->ldpre;
  ep.aux = e->aux; ep.ldaux = e->ldaux;
  ep.drop_thresh = e->drop_p > 0.f ? drop_threshold(e->drop_p) : 0u;
  ep.drop_scale = drop_keep_scale(e->drop_p);
  ep.seed = e->drop_seed; ep.offset = e->drop_offset;
  ep.residual = e->residual; ep.ldres = e->ldres;
  ep.part_rows = 0; ep.out1 = ep.out2 = nullptr;
  if (e->part_rows > 0) {
    if (!e->atomic_add || !e->out_part1 || !e->out_part2 || e->part_rows % 32 || M > 3 * e->part_rows) {
      set_error("samk_gemm: part_rows needs atomic_add, two more outputs, part_rows %% 32 == 0 and M <= 3*part_rows");
      return SAMK_ERR_ARG;
    }
    ep.part_rows = e->part_rows; ep.out1 = e->out_part1; ep.out2 = e->out_part2;
  }
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool rest_ok = (N % 8 == 0) && (!ep.out1 || (al16(ep.out1) && al16(ep.out2))) && (!ep.bias || al16(ep.bias)) &&
                       (!ep.pre || (al16(ep.pre) && ep.ldpre % 8 == 0)) && (!ep.aux || (al16(ep.aux) && ep.ldaux % 8 == 0)) &&
                       (!ep.residual || (al16(ep.residual) && ep.ldres % 4 == 0));
  const bool vec_ok = rest_ok && al16(ep.out) && (ep.ldo % 8 == 0);
  // 8-byte aligned fp32 rows: vector path with 8-byte output stores
  const bool vec2_ok = rest_ok && !vec_ok && e->out_dtype == SAMK_DT_F32 && !ep.out1 &&
                       ((uintptr_t)ep.out & 7) == 0 && (ep.ldo % 2 == 0);
  int flags = 0;
  if (e->out_dtype != SAMK_DT_F32) flags |= F_OUT_BF16 | (e->out_dtype == SAMK_DT_F16 ? F_OUT_F16 : 0);
  if (ep.bias) flags |= F_BIAS;
  if (ep.drop_thresh) flags |= F_DROP;
  if (ep.residual) flags |= F_RES;
  if (e->act == 3) flags |= F_GELU_PAIR;
  else if (ep.pre) flags |= F_PRE;
  if (ep.pre && e->pre_dtype != SAMK_DT_F32) flags |= F_PRE_BF16 | (e->pre_dtype == SAMK_DT_F16 ? F_PRE_F16 : 0);
  if (e->act == 4) flags |= F_MULAUX;
  if (e->act == 2) flags |= F_DGELU;
  if ((e->act == 2 || e->act == 4) && e->aux_dtype != SAMK_DT_F32) flags |= F_AUX_BF16 | (e->aux_dtype == SAMK_DT_F16 ? F_AUX_F16 : 0);
  if (e->act == 1) flags |= F_GELU;
  if (e->atomic_add) flags |= F_ATOMIC;
  if (ep.alpha != 1.0f || ep.alpha_dev || e->atomic_add) flags |= F_ALPHA;   // (accumulating launches always: one specialisation)
  if (vec_ok) flags |= F_VEC;
  if (vec2_ok) flags |= F_VEC | F_OUT_V2;
  ep.flags = flags;
  if (K == 0 && !e->atomic_add) split_k = 1;

  if (impl == 1) {
    long long chunks = (long long)M * ((N + 31) / 32);
    gemm_simt_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, stream>>>(
        (const __nv_bfloat16*)A, a_mn, lda, (const __nv_bfloat16*)B, b_mn, ldb, ep, M, N, K, ab_fmt);
    return check_launch("samk_gemm(simt)");
  }

  // tile width: 256 unless that leaves most of the machine idle
  int bn = 256;
  {
    const int m_tiles = (M + BM - 1) / BM;
    const long long w256 = (long long)m_tiles * ((N + 255) / 256) * split_k;
    const int sms_ = sm_count() > 0 ? sm_count() : 148;
    if (N <= 128 || w256 < sms_) bn = 128;
  }
  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = make_tmap_bf16_2d(&ta, A, K, M, lda, 64, BK);
  else rc = make_tmap_bf16_2d(&ta, A, M, K, lda, BK, BM);
  if (rc) return rc;
  // CTA pair: 256 x 256 tile per two SMs; needs at least a pair of row tiles per column tile to be worth it
  struct Spec { int am, bm, mask; };
  static const Spec kSpecs[] = {{0, 0, M_QKV}, {0, 0, M_OUTPROJ}, {0, 0, M_FFN1}, {0, 0, M_BF16}, {0, 0, M_F32_BIAS},
                                {0, 0, M_F32_BIAS_RES}, {0, 0, M_GELU}, {0, 0, M_F32}, {0, 1, M_BF16}, {0, 1, M_MULAUX},
                                {0, 1, M_F32_RES}, {0, 1, M_F32}, {1, 1, M_ATOMIC}, {0, 0, M_F32_BIAS_V2}, {1, 1, M_ATOMIC_V2}};   // keep in step with SAMK_SPEC below
  bool spec = false;
  int spec_idx = -1, si = 0;
  for (const Spec& sp : kSpecs) {
    if (sp.am == (a_mn ? 1 : 0) && sp.bm == (b_mn ? 1 : 0) && sp.mask == flags) { spec = true; spec_idx = si; }
    ++si;
  }
  const bool pair = gemm_2cta() && spec && bn == 256 && M >= 2 * BM;
  const bool ew16 = pair && ((gemm_ew16_mask() >> spec_idx) & 1);
  if (b_mn) rc = make_tmap_bf16_2d(&tb, B, K, N, ldb, 64, BK);
  else rc = make_tmap_bf16_2d(&tb, B, N, K, ldb, BK, pair ? bn / 2 : bn);
  if (rc) return rc;

  // specialised epilogues run in the coalesced (shared-memory transposed) layout: measured faster than the
  // row-per-thread layout for every combination below (tools/gemm_bench.py)
#define SAMK_SPEC(AM_, BM_, MASK_)                                                                        \
  if (a_mn == AM_ && b_mn == BM_ && flags == (MASK_)) {                                                    \
    if (ew16) return launch_tc<256, AM_, BM_, 2, MASK_, true, 16>(ta, tb, ep, M, N, K, split_k, ab_fmt, stream);  \
    if (pair) return launch_tc<256, AM_, BM_, 2, MASK_, true>(ta, tb, ep, M, N, K, split_k, ab_fmt, stream);      \
    if (bn == 256) return launch_tc<256, AM_, BM_, 1, MASK_, true>(ta, tb, ep, M, N, K, split_k, ab_fmt, stream); \
    return launch_tc<128, AM_, BM_, 1, MASK_, true>(ta, tb, ep, M, N, K, split_k, ab_fmt, stream);                \
  }
  SAMK_SPEC(0, 0, M_QKV)
  SAMK_SPEC(0, 0, M_OUTPROJ)
  SAMK_SPEC(0, 0, M_FFN1)
  SAMK_SPEC(0, 0, M_BF16)
  SAMK_SPEC(0, 0, M_F32_BIAS)
  SAMK_SPEC(0, 0, M_F32_BIAS_RES)
  SAMK_SPEC(0, 0, M_GELU)
  SAMK_SPEC(0, 0, M_F32)
  SAMK_SPEC(0, 1, M_BF16)
  SAMK_SPEC(0, 1, M_MULAUX)
  SAMK_SPEC(0, 1, M_F32_RES)
  SAMK_SPEC(0, 1, M_F32)
  SAMK_SPEC(1, 1, M_ATOMIC)
  SAMK_SPEC(0, 0, M_F32_BIAS_V2)
  SAMK_SPEC(1, 1, M_ATOMIC_V2)
#undef SAMK_SPEC

  // anything else: runtime-flag epilogue
#define SAMK_DYN(BN_, AM_, BM_) return launch_tc<BN_, AM_, BM_, 1, EPI_DYNAMIC, false>(ta, tb, ep, M, N, K, split_k, ab_fmt, stream)
  if (bn == 256) {
    if (!a_mn && !b_mn) SAMK_DYN(256, 0, 0);
    if (!a_mn && b_mn) SAMK_DYN(256, 0, 1);
    if (a_mn && !b_mn) SAMK_DYN(256, 1, 0);
    SAMK_DYN(256, 1, 1);
  } else {
    if (!a_mn && !b_mn) SAMK_DYN(128, 0, 0);
    if (!a_mn && b_mn) SAMK_DYN(128, 0, 1);
    if (a_mn && !b_mn) SAMK_DYN(128, 1, 0);
    SAMK_DYN(128, 1, 1);
  }
#undef SAMK_DYN
}

namespace samk { int set_drop_salt_gemm(unsigned long long salt, cudaStream_t stream) { return set_drop_salt_tu(salt, stream); } }
namespace samk { int set_drop_salt_dev_gemm(const unsigned long long* src, cudaStream_t stream) { return set_drop_salt_from_device_tu(src, stream); } }
