// Gradient exchange of the data-parallel step over NVLink peer memory (NVSwitch multicast when the allocation has
// one): SUM over the ranks of a range of the flat gradient buffer, in place.
//
// Replaces the gradient reduction of the reference's nn.DataParallel (/root/reference/train.py:111-112).  The library
// collective (ncclAllReduce) needs CTAs with tens of KB of shared memory and hundreds of registers per thread, which
// cannot share an SM with the persistent 226 KB GEMM CTAs of the backward pass -- run beside them it stalls them, run
// after them it is exposed.  The kernels here use no shared memory and <= 40 registers, so their blocks fit on the
// SMs the backward pass is using and the exchange of a finished bucket proceeds under the rest of the backward pass.
//
//   wire    one symmetric allocation per rank (same size everywhere), mapped into every rank's address space
//           (wire[r] = rank r's copy as seen from here) and, with NVLS, once more as a multicast address.
//   step 1  pack      wire_local[lo:hi] = to_wire(flat[lo:hi])                     (bf16 or fp32 on the wire)
//   step 2  barrier   every rank has packed                                           (flag pads in peer memory)
//   step 3  reduce    rank r owns 1/world of [lo, hi): NVLS multimem.ld_reduce pulls the sum of all ranks' copies
//                     through the switch (fp32 accumulation) and multimem.st writes it back to all of them;
//                     without multicast: peer loads from every rank, one sum, peer stores to every rank
//   step 4  barrier   every rank has written its shard everywhere
//   step 5  unpack    flat[lo:hi] = from_wire(wire_local[lo:hi])
// The barriers are one-warp kernels: thread t stores the new epoch into rank t's pad (st.release.sys) and spins on
// its own pad (ld.acquire.sys) with a clock bound -- a rank that never arrives sets *error instead of hanging the GPU.
#include "common.cuh"
#include "../../include/samk.h"

namespace samk {

constexpr int kMaxWorld = 8;
constexpr int kXchgThreads = 128;             // small blocks: <= 8 K registers each, they fit beside a GEMM CTA

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <bool BF16> __device__ __forceinline__ uint4 multimem_ld_reduce(const void* mc) {
  uint4 v;
  if constexpr (BF16) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  } else {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  }
  return v;
}
__device__ __forceinline__ void multimem_st(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- barrier over the ranks ------------------------------------------------------------------------------------
struct FlagPtrs { unsigned int* p[kMaxWorld]; };

__global__ void __launch_bounds__(32) xchg_barrier_kernel(FlagPtrs flags, unsigned int* __restrict__ epoch,
                                                          int rank, int world, int* __restrict__ error, long long timeout_clocks) {
  const int t = threadIdx.x;
  const unsigned int e = *epoch + 1u;
  __syncwarp();
  if (t < world) {
    __threadfence_system();                       // everything this GPU wrote before (previous kernels of the stream) first
    st_release_sys(flags.p[t] + rank, e);
    const unsigned int* mine = flags.p[rank] + t;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      // (once a rank has been missed the results are void anyway: later barriers do not wait again)
      if (*reinterpret_cast<volatile int*>(error) != 0 || clock64() - t0 > timeout_clocks) { atomicCAS(error, 0, 1 + t); break; }
      __nanosleep(200);
    }
  }
  __syncwarp();
  if (t == 0) *epoch = e;
}

// ---- reduce: this rank's shard [v0, v1) in 16-byte vectors ----------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kXchgThreads) xchg_reduce_mc_kernel(char* __restrict__ mc, long long v0, long long v1) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long v = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; v + 7 * stride < v1; v += 8 * stride) {      // eight switch round trips in flight per thread
    uint4 x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = multimem_ld_reduce<BF16>(mc + ((v + u * stride) << 4));
#pragma unroll
    for (int u = 0; u < 8; ++u) multimem_st(mc + ((v + u * stride) << 4), x[u]);
  }
  for (; v < v1; v += stride) multimem_st(mc + (v << 4), multimem_ld_reduce<BF16>(mc + (v << 4)));
  __threadfence_system();
}

struct PeerPtrs { char* p[kMaxWorld]; };

template <bool BF16> __device__ __forceinline__ void accumulate(float (&acc)[8], uint4 x) {
  if constexpr (BF16) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h[q]); acc[2 * q] += f.x; acc[2 * q + 1] += f.y; }
  } else {
    acc[0] += __uint_as_float(x.x); acc[1] += __uint_as_float(x.y); acc[2] += __uint_as_float(x.z); acc[3] += __uint_as_float(x.w);
  }
}

template <bool BF16>
__global__ void __launch_bounds__(kXchgThreads) xchg_reduce_p2p_kernel(PeerPtrs w, int rank, int world, long long v0, long long v1) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += stride) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    uint4 x[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)              // all peer loads in flight before the first add; fixed rank order:
      if (r < world) x[r] = ld_relaxed_sys_v4(w.p[r] + (v << 4));   // every rank computes bit-identical sums
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < world) accumulate<BF16>(acc, x[r]);
    uint4 o;
    if constexpr (BF16) {
      o = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
    } else {
      o = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
    }
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < world) st_relaxed_sys_v4(w.p[r] + (v << 4), o);
  }
  __threadfence_system();
}

// ---- pack / unpack between the fp32 buffer and the local wire copy (n8 groups of 8 elements) ------------------------
template <bool BF16, bool PACK>
__global__ void __launch_bounds__(kXchgThreads) xchg_copy_kernel(float* __restrict__ flat, void* __restrict__ wire, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
    if constexpr (PACK) {
      float4 f[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i0 + u * stride < n4) f[u] = __ldcs(reinterpret_cast<const float4*>(flat) + i0 + u * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i0 + u * stride < n4) {
        const long long i = i0 + u * stride;
        if constexpr (BF16) reinterpret_cast<uint2*>(wire)[i] = make_uint2(pack_bf16(f[u].x, f[u].y), pack_bf16(f[u].z, f[u].w));
        else reinterpret_cast<float4*>(wire)[i] = f[u];
      }
    } else if constexpr (BF16) {
      uint2 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i0 + u * stride < n4) x[u] = __ldcg(reinterpret_cast<const uint2*>(wire) + i0 + u * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i0 + u * stride < n4) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&x[u].x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&x[u].y));
        reinterpret_cast<float4*>(flat)[i0 + u * stride] = make_float4(a.x, a.y, b.x, b.y);
      }
    } else {
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i0 + u * stride < n4) x[u] = __ldcg(reinterpret_cast<const float4*>(wire) + i0 + u * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i0 + u * stride < n4) reinterpret_cast<float4*>(flat)[i0 + u * stride] = x[u];
    }
  }
}

static int xchg_blocks() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("SAMK_XCHG_BLOCKS");
    v = e ? atoi(e) : 0;
    if (v <= 0) v = 592;                       // four small blocks per SM; they share the SMs with the backward pass
  }
  return v;
}

}  // namespace samk

extern "C" int samk_exchange_sum(const samk_peer_wire* w, float* flat, long long lo, long long hi, void* stream_) {
  using namespace samk;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!w || !flat || !w->wire_peers || !w->flag_peers || !w->epoch || !w->error) { set_error("samk_exchange_sum: null pointer"); return SAMK_ERR_ARG; }
  if (w->world < 1 || w->world > kMaxWorld || w->rank < 0 || w->rank >= w->world) { set_error("samk_exchange_sum: bad rank / world (max %d ranks)", kMaxWorld); return SAMK_ERR_ARG; }
  const bool bf16 = w->wire_dtype == SAMK_DT_BF16;
  if (!bf16 && w->wire_dtype != SAMK_DT_F32) { set_error("samk_exchange_sum: wire dtype must be bf16 or f32"); return SAMK_ERR_ARG; }
  const int vec = bf16 ? 8 : 4;                // elements per 16-byte wire vector
  if (lo < 0 || hi < lo || (lo % vec) || (hi % vec)) { set_error("samk_exchange_sum: range must be aligned to %d elements", vec); return SAMK_ERR_ARG; }
  if (hi == lo || w->world == 1) return SAMK_OK;
  const long long n = hi - lo;
  const int esz = bf16 ? 2 : 4;
  char* wl = (char*)w->wire_peers[w->rank] + lo * esz;
  FlagPtrs fp;
  for (int r = 0; r < kMaxWorld; ++r) fp.p[r] = r < w->world ? w->flag_peers[r] : nullptr;
  const int blocks = xchg_blocks();
  const long long timeout = w->timeout_clocks > 0 ? w->timeout_clocks : 20000000000ll;     // ~10 s
  // 1: pack
  if (bf16) xchg_copy_kernel<true, true><<<blocks, kXchgThreads, 0, stream>>>(flat + lo, wl, n / 4);
  else xchg_copy_kernel<false, true><<<blocks, kXchgThreads, 0, stream>>>(flat + lo, wl, n / 4);
  // 2: every rank has packed
  xchg_barrier_kernel<<<1, 32, 0, stream>>>(fp, w->epoch, w->rank, w->world, w->error, timeout);
  // 3: reduce this rank's shard
  const long long nv = n / vec, per = (nv + w->world - 1) / w->world;
  const long long base = lo / vec;
  long long v0 = base + per * w->rank, v1 = v0 + per;
  if (v0 > base + nv) v0 = base + nv;
  if (v1 > base + nv) v1 = base + nv;
  if (w->wire_mc) {
    if (bf16) xchg_reduce_mc_kernel<true><<<blocks, kXchgThreads, 0, stream>>>((char*)w->wire_mc, v0, v1);
    else xchg_reduce_mc_kernel<false><<<blocks, kXchgThreads, 0, stream>>>((char*)w->wire_mc, v0, v1);
  } else {
    PeerPtrs pp;
    for (int r = 0; r < kMaxWorld; ++r) pp.p[r] = r < w->world ? (char*)w->wire_peers[r] : nullptr;
    if (bf16) xchg_reduce_p2p_kernel<true><<<blocks, kXchgThreads, 0, stream>>>(pp, w->rank, w->world, v0, v1);
    else xchg_reduce_p2p_kernel<false><<<blocks, kXchgThreads, 0, stream>>>(pp, w->rank, w->world, v0, v1);
  }
  // 4: every shard is everywhere
  xchg_barrier_kernel<<<1, 32, 0, stream>>>(fp, w->epoch, w->rank, w->world, w->error, timeout);
  // 5: unpack
  if (bf16) xchg_copy_kernel<true, false><<<blocks, kXchgThreads, 0, stream>>>(flat + lo, wl, n / 4);
  else xchg_copy_kernel<false, false><<<blocks, kXchgThreads, 0, stream>>>(flat + lo, wl, n / 4);
  return check_launch("samk_exchange_sum");
}
