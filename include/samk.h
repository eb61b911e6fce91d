/* samk -- C ABI of the B200-native SA-M4C hot path (libsamk.so).
 *
 * The reference (yashkant/sam-textvqa) has no FFI layer: its hot path is the Python nn.Module
 * `SAM4C` (/root/reference/sam/sa_m4c.py:20-371) and the NumPy graph builder
 * (/root/reference/sam/spatial_utils.py:92-218), both running on PyTorch / NumPy library ops.
 * This header is the boundary a maintainer binds (ctypes stub in INTEGRATION.md); every entry
 * point cites the reference code it replaces.
 *
 * Conventions: plain C types only; every tensor argument is a raw DEVICE pointer with explicit
 * sizes / leading dimensions (in elements); the last argument is a cudaStream_t passed as void*.
 * Entry points return 0 on success or a negative SAMK_ERR_* code, never throw, never
 * synchronise and never allocate device memory; samk_last_error() gives a thread-local message.
 * bf16 = IEEE bfloat16 bit pattern (uint16_t storage).
 */
#ifndef SAMK_H_
#define SAMK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAMK_DT_F32 0
#define SAMK_DT_BF16 1

int samk_version(void);
const char* samk_last_error(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int samk_sm_count(void);

/* ---- spatial graph ------------------------------------------------------------------------
 * Replaces build_graph_using_normalized_boxes (sam/spatial_utils.py:92-218) for a batch of box
 * lists, bit-exact.  boxes [B,N,4] (x1,y1,x2,y2; all-zero-sum rows are padding, :134).
 *   types  int8 [B,N,N]      relation type 0..12  (= the reference's matrix "1")
 *   shared int8 [8,B,N,N]    optional (NULL to skip): matrices "31","32","51","52","71","72","91","92"
 *   bits   uint16 [B,N,N]    optional: packed 12-head mask for `context` in {1,3,5,7,9}, i.e.
 *                            torch_broadcast_adj_matrix (:33-52) + the max-chain of
 *                            sam/datasets/textvqa_dataset.py:373-409; bit h = head h may attend
 *   sectors                  HOST pointer to 8x4 doubles (base, step, t1, t2) or NULL for the
 *                            built-in table (see samk_graph_default_sectors)
 */
int samk_build_graph_f32(const float* boxes, int8_t* types, int8_t* shared, uint16_t* bits, int B, int N,
                         double distance_threshold, int context, const double* sectors, void* stream);
int samk_build_graph_f64(const double* boxes, int8_t* types, int8_t* shared, uint16_t* bits, int B, int N,
                         double distance_threshold, int context, const double* sectors, void* stream);
const double* samk_graph_default_sectors(void);
/* int8 [n_pairs, heads] head masks (the layout SAM4C.forward receives, sa_m4c.py:457) <-> uint16 bits */
int samk_pack_adj(const int8_t* adj, uint16_t* bits, long long n_pairs, int heads, void* stream);
int samk_unpack_bits(const uint16_t* bits, int8_t* adj, long long n_pairs, int heads, void* stream);

/* ---- dense contractions (tcgen05) -----------------------------------------------------------
 * C[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) ), bf16 operands, fp32 accumulation in TMEM.
 * Replaces the nn.Linear / torch.matmul calls of sa_m4c.py:118-123,141-146,171,429-431,554-556,
 * 875-876 and of the third-party BertSelfOutput / BertIntermediate / BertOutput blocks
 * (sa_m4c.py:617-619, 663-668), forward and backward.
 *   a_mn_major = 0: A stored [M,K] row-major (lda = row pitch)   1: A stored [K,M] row-major
 *   b_mn_major = 0: B stored [N,K] row-major (ldb = row pitch)   1: B stored [K,N] row-major
 * Leading dimensions are in elements and must be multiples of 8; base pointers 16-byte aligned.
 * Epilogue, in this order, each step optional:
 *   v = alpha*acc; v += bias[n]; pre[m,n] = v; v = gelu(v) | v *= gelu'(aux[m,n]);
 *   v = dropout(v; p, seed, offset); v += residual[m,n]; out[m,n] = v  (or out[m,n] += v atomically)
 */
typedef struct samk_gemm_epilogue {
  void* out;             /* [M,N] */
  long long ldo;
  int out_dtype;         /* SAMK_DT_* ; must be F32 when atomic_add */
  int atomic_add;        /* 1: out += v with red.global.add.f32 (split-K / gradient accumulation) */
  float alpha;
  const float* bias;     /* [N] or NULL */
  void* pre;             /* [M,N] pre-activation copy or NULL */
  long long ldpre;
  int pre_dtype;
  int act;               /* 0 none, 1 erf-GELU (sa_m4c.py:985-991), 2 multiply by GELU'(aux) */
  const void* aux;       /* [M,N] for act 2 */
  long long ldaux;
  int aux_dtype;
  float drop_p;          /* dropout probability, 0 = off */
  unsigned long long drop_seed, drop_offset;
  const float* residual; /* [M,N] fp32 or NULL */
  long long ldres;
} samk_gemm_epilogue;

/* split_k >= 1 partitions K over CTAs (requires atomic_add and a pre-zeroed or accumulating out).
 * impl: 0 = tcgen05 tensor-core kernel (the product path), 1 = SIMT reference kernel used by the
 * GPU unit tests to cross-check the tensor-core kernel. */
int samk_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb,
                   int M, int N, int K, const samk_gemm_epilogue* ep, int split_k, int impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAMK_H_ */
