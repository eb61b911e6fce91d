// Spatial-relation graph builder: CUDA restatement of
// /root/reference/sam/spatial_utils.py:92-218 (build_graph_using_normalized_boxes), :7-30 (IoU),
// :55-89 (shared sector shifts) and of the 12-head expansion
// /root/reference/sam/spatial_utils.py:33-52 + sam/datasets/textvqa_dataset.py:373-409.
//
// One thread per output element (r,c) of one sample.  The reference fills [i,j] and [j,i]
// from the single ordered pair i<j, and the two entries are NOT mirror images (label_j is
// label_i +- pi, rounded), so thread (r,c) always evaluates the pair (i,j)=(min,max) and then
// picks the entry it owns.  All arithmetic is IEEE float64 with explicit round-to-nearest
// intrinsics: the reference evaluates every add/mul as a separately rounded NumPy op, and an
// FMA contraction would change the last bit of diag / IoU at a decision boundary.
//
// The angle -> sector step is done WITHOUT asin/acos: sector = ceil(theta/(pi/4)) is a monotone
// step function of sin (quadrants 1,4) or cos (quadrants 2,3) with exactly two steps per
// (quadrant, role); the step positions of the reference's NumPy arcsin/arccos are passed in
// as a table (samk_graph_default_sectors holds the values derived from NumPy 2.3 / AVX-512
// SVML by bisection, tests/test_graph_sectors.py re-derives them on the box it runs on).
#include "common.cuh"
#include "../../include/samk.h"

namespace samk {

struct SectorTable {
  // index = quadrant*2 + role; quadrant order q1,q4,q2,q3 ; role 0 = [i,j] entry, 1 = [j,i]
  double t1[8], t2[8];
  int base[8], step[8];
};

// Derived from np.arcsin / np.arccos (NumPy 2.3.5, x86-64 AVX-512) -- see file header.
static const double kDefaultSectors[8][4] = {
    // base, step, t1, t2      value = base + step*((x>=t1) + (x>=t2))
    {3, +1, 0x0.0000000000001p-1022, 0x1.6a09e667f3bcdp-1},     // q1 i : x = sin
    {7, +1, 0x1.0000000000001p-52, 0x1.6a09e667f3bcfp-1},       // q1 j
    {9, +1, -0x1.fffffffffffffp-1, -0x1.6a09e667f3bc9p-1},      // q4 i : x = sin
    {5, +1, -0x1.fffffffffffffp-1, -0x1.6a09e667f3bc9p-1},      // q4 j
    {7, -1, -0x1.6a09e667f3bcdp-1, -0x1.cb3b399d747f4p-55},     // q2 i : x = cos
    {11, -1, -0x1.6a09e667f3bcdp-1, -0x1.1cb3b399d747fp-51},    // q2 j
    {7, +1, -0x1.fffffffffffffp-1, -0x1.6a09e667f3bcap-1},      // q3 i : x = cos
    {3, +1, -0x1.fffffffffffffp-1, -0x1.6a09e667f3bcap-1},      // q3 j
};

template <class T> struct Box4;
template <> struct Box4<float> {
  static __device__ __forceinline__ void load(const float* p, double& a, double& b, double& c, double& d) {
    float4 v = *reinterpret_cast<const float4*>(p);
    a = (double)v.x; b = (double)v.y; c = (double)v.z; d = (double)v.w;
  }
};
template <> struct Box4<double> {
  static __device__ __forceinline__ void load(const double* p, double& a, double& b, double& c, double& d) {
    double2 lo = *reinterpret_cast<const double2*>(p);
    double2 hi = *reinterpret_cast<const double2*>(p + 2);
    a = lo.x; b = lo.y; c = hi.x; d = hi.y;
  }
};

__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }

// relation type for entry (role 0: [i,j], role 1: [j,i]) of ordered pair i<j; *directional set
// when the pair went through the angular branch (only then the shared matrices are written).
__device__ __forceinline__ int classify_pair(const double* bi, const double* bj, int role, double thr,
                                             const SectorTable& st, bool* directional) {
  *directional = false;
  const double ix1 = bi[0], iy1 = bi[1], ix2 = bi[2], iy2 = bi[3];
  const double jx1 = bj[0], jy1 = bj[1], jx2 = bj[2], jy2 = bj[3];
  if (ix1 < jx1 && ix2 > jx2 && iy1 < jy1 && iy2 > jy2) return role == 0 ? 1 : 2;   // :143-150
  if (jx1 < ix1 && jx2 > ix2 && jy1 < iy1 && jy2 > iy2) return role == 0 ? 2 : 1;   // :152-159
  // IoU (:7-30)
  double xA = ix1 > jx1 ? ix1 : jx1, yA = iy1 > jy1 ? iy1 : jy1;
  double xB = ix2 < jx2 ? ix2 : jx2, yB = iy2 < jy2 ? iy2 : jy2;
  double w = dsub(xB, xA), h = dsub(yB, yA);
  w = w > 0.0 ? w : 0.0;
  h = h > 0.0 ? h : 0.0;
  double inter = __dmul_rn(w, h);
  double areaA = __dmul_rn(dsub(ix2, ix1), dsub(iy2, iy1));
  double areaB = __dmul_rn(dsub(jx2, jx1), dsub(jy2, jy1));
  double iou = __ddiv_rn(inter, dsub(__dadd_rn(areaA, areaB), inter));
  if (iou >= 0.5) return 3;                                                          // :161-166
  double cyi = __dmul_rn(0.5, __dadd_rn(iy1, iy2)), cyj = __dmul_rn(0.5, __dadd_rn(jy1, jy2));
  double cxi = __dmul_rn(0.5, __dadd_rn(ix1, ix2)), cxj = __dmul_rn(0.5, __dadd_rn(jx1, jx2));
  double yd = dsub(cyi, cyj), xd = dsub(cxi, cxj);
  double diag = __dsqrt_rn(__dadd_rn(__dmul_rn(yd, yd), __dmul_rn(xd, xd)));
  if (!(diag < thr)) return 0;                                                       // :171
  *directional = true;
  double s = __ddiv_rn(yd, diag), c = __ddiv_rn(xd, diag);
  if (s != s || c != c) return 4;                                                    // NaN rule :192-203
  int q;
  double x;
  if (s >= 0.0 && c >= 0.0) { q = 0; x = s; }
  else if (s < 0.0 && c >= 0.0) { q = 1; x = s; }
  else if (s >= 0.0 && c < 0.0) { q = 2; x = c; }
  else { q = 3; x = c; }
  int k = q * 2 + role;
  return st.base[k] + st.step[k] * ((x >= st.t1[k] ? 1 : 0) + (x >= st.t2[k] ? 1 : 0));
}

__device__ __forceinline__ uint16_t head_bits(int t, int radius) {
  if (t <= 0) return 0;
  if (t < 4 || t > 11) return (uint16_t)(1u << (t - 1));
  uint32_t v = 0;
  for (int d = -radius; d <= radius; ++d) v |= 1u << (3 + ((t - 4 + d) & 7));
  return (uint16_t)v;
}

constexpr int kGraphThreads = 256;
constexpr int kMaxBoxesSmem = 1024;

template <class T>
__global__ void __launch_bounds__(kGraphThreads)
graph_kernel(const T* __restrict__ boxes, int8_t* __restrict__ types, int8_t* __restrict__ shared,
             uint16_t* __restrict__ bits, int B, int N, double thr, int radius, SectorTable st) {
  extern __shared__ double sbox[];  // [N][4] + pad flags
  const int b = blockIdx.y;
  const T* bp = boxes + (size_t)b * N * 4;
  uint8_t* spad = reinterpret_cast<uint8_t*>(sbox + (size_t)N * 4);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double x1, y1, x2, y2;
    Box4<T>::load(bp + (size_t)i * 4, x1, y1, x2, y2);
    sbox[i * 4 + 0] = x1; sbox[i * 4 + 1] = y1; sbox[i * 4 + 2] = x2; sbox[i * 4 + 3] = y2;
    // Python sum(): (((0+x1)+y1)+x2)+y2 == 0  (:134)
    spad[i] = (__dadd_rn(__dadd_rn(__dadd_rn(x1, y1), x2), y2) == 0.0) ? 1 : 0;
  }
  __syncthreads();
  const size_t nn = (size_t)N * N;
  const size_t plane = (size_t)B * nn;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(e / N), c = (int)(e % N);
    int t = 0;
    bool directional = false;
    if (!spad[r] && !spad[c]) {
      if (r == c) t = 12;                                                            // :136
      else {
        int i = r < c ? r : c, j = r < c ? c : r;
        t = classify_pair(sbox + i * 4, sbox + j * 4, r < c ? 0 : 1, thr, st, &directional);
      }
    }
    size_t o = (size_t)b * nn + e;
    types[o] = (int8_t)t;
    if (shared) {
      // order 31,32,51,52,71,72,91,92 : sector shifted by +1,-1,+2,-2,+3,-3,+4,-4 (:68-87)
      const bool dir = directional && t >= 4 && t <= 11;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int sh = (k & 1) ? -(k / 2 + 1) : (k / 2 + 1);
        shared[(size_t)k * plane + o] = dir ? (int8_t)(((t - 4 + sh) & 7) + 4) : (int8_t)0;
      }
    }
    if (bits) bits[o] = head_bits(t, radius);
  }
}

// int8 [B,A,A,12] reference-layout head masks -> packed uint16 bits (bit h = head h allowed)
__global__ void pack_adj_kernel(const int8_t* __restrict__ adj, uint16_t* __restrict__ bits, size_t n, int H) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int8_t* p = adj + e * H;
  uint32_t v = 0;
  for (int h = 0; h < H; ++h) v |= (p[h] != 0 ? 1u : 0u) << h;
  bits[e] = (uint16_t)v;
}

// relation types -> packed head bits for a context (closed form of the dataset's max-chain)
__global__ void types_to_bits_kernel(const int8_t* __restrict__ types, uint16_t* __restrict__ bits, size_t n, int radius) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) bits[e] = head_bits(types[e], radius);
}

// packed bits -> int8 [.., 12] (for handing reference-layout masks back to callers)
__global__ void unpack_bits_kernel(const uint16_t* __restrict__ bits, int8_t* __restrict__ adj, size_t n, int H) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * H) return;
  adj[e] = (int8_t)((bits[e / H] >> (e % H)) & 1);
}

template <class T>
static int launch_graph(const T* boxes, int8_t* types, int8_t* shared, uint16_t* bits, int B, int N,
                        double dist_thr, int context, const double* sectors, cudaStream_t stream) {
  if (B < 0 || N < 0) { set_error("samk_build_graph: negative size"); return SAMK_ERR_ARG; }
  if (B == 0 || N == 0) return SAMK_OK;
  if (!boxes || !types) { set_error("samk_build_graph: null pointer"); return SAMK_ERR_ARG; }
  if (N > kMaxBoxesSmem) { set_error("samk_build_graph: N=%d exceeds %d", N, kMaxBoxesSmem); return SAMK_ERR_UNSUPPORTED; }
  if (context < 1 || context > 9 || !(context & 1)) { set_error("samk_build_graph: context must be 1,3,5,7,9"); return SAMK_ERR_ARG; }
  SectorTable st;
  const double (*src)[4] = sectors ? reinterpret_cast<const double (*)[4]>(sectors) : kDefaultSectors;
  for (int k = 0; k < 8; ++k) {
    st.base[k] = (int)src[k][0]; st.step[k] = (int)src[k][1]; st.t1[k] = src[k][2]; st.t2[k] = src[k][3];
  }
  size_t nn = (size_t)N * N;
  int gx = (int)((nn + kGraphThreads * 4 - 1) / (kGraphThreads * 4));
  if (gx < 1) gx = 1;
  dim3 grid(gx, B);
  size_t smem = (size_t)N * 4 * sizeof(double) + N;
  graph_kernel<T><<<grid, kGraphThreads, smem, stream>>>(boxes, types, shared, bits, B, N,
                                                          dist_thr * sqrt(2.0), (context - 1) / 2, st);
  return check_launch("samk_build_graph");
}

}  // namespace samk

extern "C" {

int samk_build_graph_f32(const float* boxes, int8_t* types, int8_t* shared, uint16_t* bits, int B, int N,
                         double distance_threshold, int context, const double* sectors, void* stream) {
  return samk::launch_graph<float>(boxes, types, shared, bits, B, N, distance_threshold, context, sectors,
                                   (cudaStream_t)stream);
}

int samk_build_graph_f64(const double* boxes, int8_t* types, int8_t* shared, uint16_t* bits, int B, int N,
                         double distance_threshold, int context, const double* sectors, void* stream) {
  return samk::launch_graph<double>(boxes, types, shared, bits, B, N, distance_threshold, context, sectors,
                                    (cudaStream_t)stream);
}

const double* samk_graph_default_sectors(void) { return &samk::kDefaultSectors[0][0]; }

int samk_pack_adj(const int8_t* adj, uint16_t* bits, long long n_pairs, int heads, void* stream) {
  if (!adj || !bits || heads < 1 || heads > 16) { samk::set_error("samk_pack_adj: bad argument"); return SAMK_ERR_ARG; }
  if (n_pairs <= 0) return SAMK_OK;
  samk::pack_adj_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(adj, bits, (size_t)n_pairs, heads);
  return samk::check_launch("samk_pack_adj");
}

int samk_types_to_bits(const int8_t* types, uint16_t* bits, long long n, int context, void* stream) {
  if (!types || !bits || context < 1 || context > 9 || !(context & 1)) { samk::set_error("samk_types_to_bits: bad argument"); return SAMK_ERR_ARG; }
  if (n <= 0) return SAMK_OK;
  samk::types_to_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(types, bits, (size_t)n, (context - 1) / 2);
  return samk::check_launch("samk_types_to_bits");
}

int samk_unpack_bits(const uint16_t* bits, int8_t* adj, long long n_pairs, int heads, void* stream) {
  if (!adj || !bits || heads < 1 || heads > 16) { samk::set_error("samk_unpack_bits: bad argument"); return SAMK_ERR_ARG; }
  if (n_pairs <= 0) return SAMK_OK;
  samk::unpack_bits_kernel<<<(unsigned)((n_pairs * heads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits, adj, (size_t)n_pairs, heads);
  return samk::check_launch("samk_unpack_bits");
}

}  // extern "C"
