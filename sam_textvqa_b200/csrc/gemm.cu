// tcgen05 GEMM with fused epilogues: C[M,N] = epi(alpha * A[M,K] * B[N,K]^T), bf16 x bf16 -> fp32.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B swizzle, kStages-deep mbarrier ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma, accumulators in TMEM,
//                               two accumulator buffers so the epilogue overlaps the next tile)
//   warps 2..5  epilogue       (tcgen05.ld TMEM -> registers -> bias / GELU / dropout / residual ->
//                               global; warp w owns TMEM lanes 32*(w%4)..+31)
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
constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quarter, each takes half the columns
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;

struct EpiArgs {
  void* out; long long ldo; int out_bf16; int atomic_add; float alpha;
  const float* bias;
  void* pre; long long ldpre; int pre_bf16;
  int act; const void* aux; long long ldaux; int aux_bf16;
  uint32_t drop_thresh; float drop_scale; unsigned long long seed, offset;
  const float* residual; long long ldres;
  int vec_ok;  // all pitches / pointers 16B friendly and N % 4 == 0
};

// Apply the epilogue to `cnt` (<=32, multiple handled generally) consecutive columns of one row.
__device__ __forceinline__ void epilogue_row_chunk(const EpiArgs& ep, float* v, int row, int col0, int cnt, int N) {
  const bool full = (cnt == 32) && ep.vec_ok;
  float pre_v[32];
  float* const v_in = v;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] *= ep.alpha;
  if (ep.bias) {
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 b = *reinterpret_cast<const float4*>(ep.bias + col0 + i);
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    } else {
      _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) v[i] += ep.bias[col0 + i];
    }
  }
  if (ep.act == 3) {
    // out = gelu(v), pre = gelu'(v): one erf and one exp serve both (backward then only multiplies)
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      gelu_pair(v[i], v[i], pre_v[i]);
    }
  }
  if (ep.pre) {
    const float* v = (ep.act == 3) ? pre_v : v_in;
    if (ep.pre_bf16) {
      __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(ep.pre) + (size_t)row * ep.ldpre + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(p + i) = make_uint4(pack_bf16(v[i], v[i + 1]), pack_bf16(v[i + 2], v[i + 3]),
                                                        pack_bf16(v[i + 4], v[i + 5]), pack_bf16(v[i + 6], v[i + 7]));
      } else {
        _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) p[i] = __float2bfloat16_rn(v[i]);
      }
    } else {
      float* p = reinterpret_cast<float*>(ep.pre) + (size_t)row * ep.ldpre + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
        _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) p[i] = v[i];
      }
    }
  }
  if (ep.act == 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  } else if (ep.act == 4) {
    if (ep.aux_bf16) {
      const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + (size_t)row * ep.ldaux + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u = *reinterpret_cast<const uint4*>(p + i);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(h[j]);
            v[i + 2 * j] *= f.x;
            v[i + 2 * j + 1] *= f.y;
          }
        }
      } else {
        _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) v[i] *= __bfloat162float(p[i]);
      }
    } else {
      const float* p = reinterpret_cast<const float*>(ep.aux) + (size_t)row * ep.ldaux + col0;
      _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) v[i] *= p[i];
    }
  } else if (ep.act == 2) {
    if (ep.aux_bf16) {
      const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + (size_t)row * ep.ldaux + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u = *reinterpret_cast<const uint4*>(p + i);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(h[j]);
            v[i + 2 * j] *= dgelu_erf(f.x);
            v[i + 2 * j + 1] *= dgelu_erf(f.y);
          }
        }
      } else {
        _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) v[i] *= dgelu_erf(__bfloat162float(p[i]));
      }
    } else {
      const float* p = reinterpret_cast<const float*>(ep.aux) + (size_t)row * ep.ldaux + col0;
      _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) v[i] *= dgelu_erf(p[i]);
    }
  }
  if (ep.drop_thresh) {
    const uint64_t row_ctr = (uint64_t)row * (uint64_t)((N + 3) >> 2);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (i < cnt) {
        uint4 r = dropout_bits4(ep.seed, ep.offset, row_ctr + (uint64_t)((col0 + i) >> 2));
        v[i] = r.x >= ep.drop_thresh ? v[i] * ep.drop_scale : 0.f;
        v[i + 1] = r.y >= ep.drop_thresh ? v[i + 1] * ep.drop_scale : 0.f;
        v[i + 2] = r.z >= ep.drop_thresh ? v[i + 2] * ep.drop_scale : 0.f;
        v[i + 3] = r.w >= ep.drop_thresh ? v[i + 3] * ep.drop_scale : 0.f;
      }
    }
  }
  if (ep.residual) {
    const float* p = ep.residual + (size_t)row * ep.ldres + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 b = *reinterpret_cast<const float4*>(p + i);
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    } else {
      _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) v[i] += p[i];
    }
  }
  if (ep.atomic_add) {
    float* p = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3]) : "memory");
    } else {
      _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) atomicAdd(p + i, v[i]);
    }
  } else if (ep.out_bf16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * ep.ldo + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        *reinterpret_cast<uint4*>(p + i) = make_uint4(pack_bf16(v[i], v[i + 1]), pack_bf16(v[i + 2], v[i + 3]),
                                                      pack_bf16(v[i + 4], v[i + 5]), pack_bf16(v[i + 6], v[i + 7]));
    } else {
      _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) p[i] = __float2bfloat16_rn(v[i]);
    }
  } else {
    float* p = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
      _Pragma("unroll") for (int i = 0; i < 32; ++i) if (i < cnt) p[i] = v[i];
    }
  }
}

template <int BN> struct GemmCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;  // 256 or 512: power of two
};

// CL = cluster size along M (1 or 2).  With CL = 2 the two CTAs of a cluster work on vertically adjacent
// output tiles that need the same B tile: each CTA fetches half of it and TMA-multicasts it into both
// shared memories, cutting the L2->SM operand traffic per flop by a third (the kernel is L2-bandwidth
// bound at 128x256 tiles otherwise).  Slots are released to both producers with a multicast commit.
template <int BN, int A_MN, int B_MN, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const EpiArgs ep, int M, int N, int K, int split_k) {
  using Cfg = GemmCfg<BN>;
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
  constexpr uint16_t kMcastMask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], CL); }
      for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], kEpiWarps); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  if (CL > 1) ptx::cluster_sync_all();   // peer barriers must be initialised before remote arrivals / multicasts
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

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
          ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (A_MN) {
#pragma unroll
            for (int g = 0; g < BM / 64; ++g) ptx::tma_load_2d(sa + g * (BK * 128), &tmA, &full_bar[stage], m0 + g * 64, kb * BK);
          } else {
            ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          }
          if (CL == 1) {
            if (B_MN) {
#pragma unroll
              for (int g = 0; g < BN / 64; ++g) ptx::tma_load_2d(sb + g * (BK * 128), &tmB, &full_bar[stage], n0 + g * 64, kb * BK);
            } else {
              ptx::tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
            }
          } else {
            // this CTA fetches its 1/CL share of the B tile and multicasts it to the whole cluster
            if (B_MN) {
#pragma unroll
              for (int g = 0; g < BN / 64 / CL; ++g) {
                const int gg = cta_rank * (BN / 64 / CL) + g;
                ptx::tma_load_2d_mcast(sb + gg * (BK * 128), &tmB, &full_bar[stage], n0 + gg * 64, kb * BK, kMcastMask);
              }
            } else {
              ptx::tma_load_2d_mcast(sb + cta_rank * (BN / CL) * 128, &tmB, &full_bar[stage], kb * BK,
                                     n0 + cta_rank * (BN / CL), kMcastMask);
            }
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN, B_MN);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
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
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? ptx::make_smem_desc_sw128(sa + k * 2048, BK * 128, 1024)
                                        : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? ptx::make_smem_desc_sw128(sb + k * 2048, BK * 128, 1024)
                                        : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
            ptx::umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (CL == 1) ptx::umma_commit(&empty_bar[stage]);      // smem slot free once these MMAs retire
          else ptx::umma_commit_mcast(&empty_bar[stage], kMcastMask);   // ... in every CTA that multicasts into it
          if (kb == kb1 - 1) ptx::umma_commit(&tfull_bar[acc]);  // accumulator ready
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (kb1 <= kb0 && lane == 0) ptx::umma_commit(&tfull_bar[acc]);  // empty K range: still hand over
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue =====================
    const int lane_grp = warp & 3;          // TMEM lane quarter this warp may access (warp id % 4)
    const int col_half = (warp - 2) >> 2;   // which half of the tile's columns this warp drains
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = work0; w < num_work; w += work_stride) {
      const int split = w % split_k;
      const int tile = w / split_k;
      const int m0 = ((tile / n_tiles) * CL + cta_rank) * BM, n0 = (tile % n_tiles) * BN;
      const int kb0 = split * kb_per;
      const bool has_k = min(kb_total, kb0 + kb_per) > kb0;
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = m0 + lane_grp * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(lane_grp * 32) << 16);
#pragma unroll 1
      for (int c = col_half * (BN / 64); c < (col_half + 1) * (BN / 64); ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= N) break;  // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr + c * 32, r);
        ptx::tmem_ld_wait();
        if (row < M && (has_k || !ep.atomic_add)) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = has_k ? __uint_as_float(r[i]) : 0.f;
          epilogue_row_chunk(ep, v, row, col0, min(32, N - col0), N);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  if (CL > 1) ptx::cluster_sync_all();   // no CTA may exit while its peer can still signal its barriers
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT reference kernel (unit tests only: cross-checks the tensor-core kernel on the device).
// ---------------------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, int a_mn, long long lda,
                                 const __nv_bfloat16* __restrict__ B, int b_mn, long long ldb,
                                 const EpiArgs ep, int M, int N, int K) {
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
    float a = __bfloat162float(a_mn ? A[(size_t)k * lda + row] : A[(size_t)row * lda + k]);
    for (int i = 0; i < cnt; ++i) {
      float b = __bfloat162float(b_mn ? B[(size_t)k * ldb + col0 + i] : B[(size_t)(col0 + i) * ldb + k]);
      v[i] = fmaf(a, b, v[i]);
    }
  }
  EpiArgs e2 = ep;
  e2.vec_ok = 0;
  epilogue_row_chunk(e2, v, row, col0, cnt, N);
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
int sm_count() {
  if (!g_sm_count) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    g_sm_count = n;
  }
  return g_sm_count;
}

template <int BN, int A_MN, int B_MN, int CL>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const EpiArgs& ep, int M, int N, int K, int split_k,
                     cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CL>;
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
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, ep, M, N, K, split_k);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("samk_gemm_bf16: cluster launch failed: %s", cudaGetErrorString(e));
    return SAMK_ERR_CUDA;
  }
  return check_launch("samk_gemm_bf16");
}

static int g_gemm_cluster = -1;
static int gemm_cluster() {
  if (g_gemm_cluster < 0) {
    const char* s = getenv("SAMK_GEMM_CLUSTER");
    g_gemm_cluster = (s && s[0] == '2') ? 2 : 1;
  }
  return g_gemm_cluster;
}

}  // namespace samk

extern "C" int samk_gemm_bf16(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, int M,
                              int N, int K, const samk_gemm_epilogue* e, int split_k, int impl, void* stream_) {
  using namespace samk;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!A || !B || !e || !e->out) { set_error("samk_gemm_bf16: null pointer"); return SAMK_ERR_ARG; }
  if (M < 0 || N < 0 || K < 0) { set_error("samk_gemm_bf16: negative size"); return SAMK_ERR_ARG; }
  if (M == 0 || N == 0) return SAMK_OK;
  if (split_k < 1) split_k = 1;
  if (split_k > 1 && !e->atomic_add) { set_error("samk_gemm_bf16: split_k>1 needs atomic_add"); return SAMK_ERR_ARG; }
  if (e->atomic_add && e->out_dtype != SAMK_DT_F32) { set_error("samk_gemm_bf16: atomic_add needs fp32 out"); return SAMK_ERR_ARG; }
  if ((lda % 8) || (ldb % 8) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) {
    set_error("samk_gemm_bf16: operands need 16-byte aligned base and ld %% 8 == 0 (lda=%lld ldb=%lld)", lda, ldb);
    return SAMK_ERR_ARG;
  }
  EpiArgs ep;
  ep.out = e->out; ep.ldo = e->ldo; ep.out_bf16 = e->out_dtype == SAMK_DT_BF16; ep.atomic_add = e->atomic_add;
  ep.alpha = e->alpha; ep.bias = e->bias;
  ep.pre = e->pre; ep.ldpre = e->ldpre; ep.pre_bf16 = e->pre_dtype == SAMK_DT_BF16;
  ep.act = e->act; ep.aux = e->aux; ep.ldaux = e->ldaux; ep.aux_bf16 = e->aux_dtype == SAMK_DT_BF16;
  ep.drop_thresh = e->drop_p > 0.f ? drop_threshold(e->drop_p) : 0u;
  ep.drop_scale = e->drop_p > 0.f ? 1.0f / (1.0f - e->drop_p) : 1.0f;
  ep.seed = e->drop_seed; ep.offset = e->drop_offset;
  ep.residual = e->residual; ep.ldres = e->ldres;
  if ((ep.act == 2 || ep.act == 4) && !ep.aux) { set_error("samk_gemm_bf16: act=2/4 needs aux"); return SAMK_ERR_ARG; }
  if (ep.act == 3 && !ep.pre) { set_error("samk_gemm_bf16: act=3 needs pre"); return SAMK_ERR_ARG; }
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  ep.vec_ok = (N % 4 == 0) && al16(ep.out) && (ep.ldo % 8 == 0) && (!ep.bias || al16(ep.bias)) &&
              (!ep.pre || (al16(ep.pre) && ep.ldpre % 8 == 0)) && (!ep.aux || (al16(ep.aux) && ep.ldaux % 8 == 0)) &&
              (!ep.residual || (al16(ep.residual) && ep.ldres % 4 == 0));
  if (K == 0 && !ep.atomic_add) split_k = 1;

  if (impl == 1) {
    long long chunks = (long long)M * ((N + 31) / 32);
    gemm_simt_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, stream>>>(
        (const __nv_bfloat16*)A, a_mn, lda, (const __nv_bfloat16*)B, b_mn, ldb, ep, M, N, K);
    return check_launch("samk_gemm_bf16(simt)");
  }

  // tile width: 256 unless that leaves most of the machine idle
  int bn = 256;
  {
    const int m_tiles = (M + BM - 1) / BM;
    const long long w256 = (long long)m_tiles * ((N + 255) / 256) * split_k;
    if (N <= 128 || w256 < 148) bn = 128;
  }
  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = make_tmap_bf16_2d(&ta, A, K, M, lda, 64, BK);
  else rc = make_tmap_bf16_2d(&ta, A, M, K, lda, BK, BM);
  if (rc) return rc;
  const int cl = gemm_cluster();
  if (b_mn) rc = make_tmap_bf16_2d(&tb, B, K, N, ldb, 64, BK);
  else rc = make_tmap_bf16_2d(&tb, B, N, K, ldb, BK, bn / cl);     // each cluster CTA fetches bn/cl rows of B
  if (rc) return rc;

#define SAMK_LAUNCH(BN_, AM_, BM_)                                                        \
  do {                                                                                    \
    if (cl == 2) return launch_tc<BN_, AM_, BM_, 2>(ta, tb, ep, M, N, K, split_k, stream); \
    return launch_tc<BN_, AM_, BM_, 1>(ta, tb, ep, M, N, K, split_k, stream);              \
  } while (0)
  if (bn == 256) {
    if (!a_mn && !b_mn) SAMK_LAUNCH(256, 0, 0);
    if (!a_mn && b_mn) SAMK_LAUNCH(256, 0, 1);
    if (a_mn && !b_mn) SAMK_LAUNCH(256, 1, 0);
    SAMK_LAUNCH(256, 1, 1);
  } else {
    if (!a_mn && !b_mn) SAMK_LAUNCH(128, 0, 0);
    if (!a_mn && b_mn) SAMK_LAUNCH(128, 0, 1);
    if (a_mn && !b_mn) SAMK_LAUNCH(128, 1, 0);
    SAMK_LAUNCH(128, 1, 1);
  }
#undef SAMK_LAUNCH
}
