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
 * bf16 = bfloat16 bit pattern, f16 = IEEE binary16 bit pattern (both uint16_t storage).  The tensor-core path stores
 * forward activations and weight operands as f16 (11 significant bits: logits within 1e-3 of the fp32 reference,
 * sa_m4c.py:179-202) and gradients as bf16 (fp32 exponent range).  A tcgen05 product needs both operands in ONE
 * format, so where a gradient meets a saved f16 activation either the activation has a bf16 copy (written by the
 * LayerNorm that produced it) or the gradient is re-expressed in f16 with an exactly computed power-of-two scale
 * (samk_cast_scaled_f16; attention backward: one scale per (sample, head)).  There is no global loss scale.
 */
#ifndef SAMK_H_
#define SAMK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAMK_DT_F32 0
#define SAMK_DT_BF16 1
#define SAMK_DT_F16 2   /* IEEE half: forward activations and weight operands (11 significant bits) */

int samk_version(void);
const char* samk_last_error(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int samk_sm_count(void);
/* Leave n SMs out of the grids of the persistent kernels (GEMM, attention) launched from now on, for a collective
 * that runs concurrently (the overlapped gradient all-reduce: NCCL capped at <= n CTAs).  Returns the number of SMs
 * the persistent kernels will use (>= 2, even), or a negative error code.  n = 0 restores the default. */
int samk_reserve_sms(int n);
/* XORed into the key of every dropout stream of every kernel launched afterwards on this device (default 0).  Kernels
 * replayed from a CUDA graph keep the (seed, offset) they were captured with; setting a new salt before each replay
 * (stream-ordered, not itself captured) gives every replay fresh masks, identical in its forward and backward. */
int samk_set_dropout_salt(unsigned long long salt, void* stream);
/* Device-side form: a one-thread kernel advances an on-device counter, hashes it (splitmix64) and the result becomes
 * the salt through device-to-device copies.  Everything is stream-ordered and capturable, so a captured training
 * step that starts with this call draws new masks on every replay with no host involvement. */
int samk_advance_dropout_salt(void* stream);

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
/* relation types 0..12 -> packed head bits for context c (torch_broadcast_adj_matrix + dataset max-chain) */
int samk_types_to_bits(const int8_t* types, uint16_t* bits, long long n, int context, void* stream);

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
  int act;               /* 0 none, 1 erf-GELU (sa_m4c.py:985-991), 2 multiply by GELU'(aux),
                            3 out = GELU(v) and pre = GELU'(v) (forward FFN1), 4 multiply by aux (backward FFN2 dgrad) */
  const void* aux;       /* [M,N] for act 2 */
  long long ldaux;
  int aux_dtype;
  float drop_p;          /* dropout probability, 0 = off */
  unsigned long long drop_seed, drop_offset;
  const float* residual; /* [M,N] fp32 or NULL */
  long long ldres;
  /* atomic_add only: if part_rows > 0 the M rows are consecutive groups of part_rows rows with separate fp32
   * destinations out, out_part1, out_part2 (each [part_rows, N], pitch ldo): the weight gradients of the fused
   * q|k|v projection are three separate parameters, one launch computes all three. part_rows % 32 == 0. */
  int part_rows;
  void* out_part1;
  void* out_part2;
  const float* alpha_dev; /* optional DEVICE scalar multiplied into alpha by the kernel: 1 / scale of an operand that
                             samk_cast_scaled_f16 scaled into the f16 range (weight gradients from f16-scaled dY) */
} samk_gemm_epilogue;

/* split_k >= 1 partitions K over CTAs (requires atomic_add and a pre-zeroed or accumulating out).
 * impl: 0 = tcgen05 tensor-core kernel (the product path), 1 = SIMT reference kernel used by the
 * GPU unit tests to cross-check the tensor-core kernel. */
int samk_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb,
                   int M, int N, int K, const samk_gemm_epilogue* ep, int split_k, int impl, void* stream);
/* same with the 16-bit storage format of the operands given (SAMK_DT_BF16 or SAMK_DT_F16; A and B must agree --
 * tcgen05.mma kind::f16 faults on mixed formats, measured on B200).  Forward products are f16 x f16; dgrad is
 * bf16 dY x bf16 weight copy; wgrad is either bf16 dY x bf16 activation copy or f16-scaled dY x f16 activation. */
int samk_gemm_16(const void* A, int a_dtype, int a_mn_major, long long lda, const void* B, int b_dtype, int b_mn_major,
                 long long ldb, int M, int N, int K, const samk_gemm_epilogue* ep, int split_k, int impl, void* stream);

/* ---- HBM-bound row kernels ------------------------------------------------------------------
 * fp32 -> bf16 operand cast; fp32 -> three bf16 planes (hi,hi,lo | hi,lo,hi) so that a plain bf16
 * GEMM over the 3x longer K reproduces fp32 products to ~2^-16 ("bf16x3" parity mode). */
int samk_cast_bf16(const float* x, long long ldx, void* y, long long ldy, int rows, int cols, void* stream);
int samk_cast_16(const float* x, long long ldx, void* y, long long ldy, int y_dtype, int rows, int cols, void* stream);
/* both 16-bit copies of an fp32 matrix in one pass (a weight: half for the forward product, bf16 for dgrad) */
int samk_cast_dual(const float* x, long long ldx, void* y_f16, void* y_bf16, long long ldy, int rows, int cols, void* stream);
/* contiguous copy between fp32 and a 16-bit format (either direction; e.g. the bf16 wire format of the gradient
 * all-reduce that replaces nn.DataParallel's reduction, train.py:111-112) */
int samk_cast_flat(const void* x, int x_dtype, void* y, int y_dtype, long long n, void* stream);
/* y = half(x * S) with S = the power of two that puts max|x| into [2^11, 2^12) (computed on the device, two passes,
 * saturating conversion); x contiguous, n % 4 == 0, x_dtype F32 or BF16.  amax: optional device float holding max|x|
 * already (then one pass).  scale2: 3 floats of device workspace,
 * scale2[0] = S, scale2[1] = 1/S (pass as samk_gemm_epilogue.alpha_dev of the product that consumes y). */
int samk_cast_scaled_f16(const void* x, int x_dtype, long long n, const float* amax, void* y, float* scale2, void* stream);
int samk_split3_bf16(const float* x, long long ldx, void* y, long long ldy, int rows, int cols, int order,
                     int along_rows, void* stream);
/* F.normalize(x, dim=-1) of sa_m4c.py:208-209,224-238 (normalize=0: plain copy/cast) */
int samk_l2norm(const float* x, long long ldx, void* y, long long ldy, int y_dtype, int rows, int cols, int normalize,
                void* stream);
/* the same with two outputs from one pass over x (the half copy the feature projection reads and the bf16 copy its
 * weight-gradient product reads: the normalised fp32 features are then never materialised) */
int samk_l2norm2(const float* x, long long ldx, void* y, long long ldy, int y_dtype, void* y2, long long ldy2, int y2_dtype,
                 int rows, int cols, int normalize, void* stream);
/* BertLayerNorm (sa_m4c.py:1016-1028; eps inside the sqrt, biased variance).  y fp32 and/or y2 in
 * y2_dtype and/or y3 in y3_dtype (the f16 copy the next contraction reads and the bf16 copy its weight-gradient
 * product reads, written in the same pass).  Backward: dx fp32; optional dxd = dropout_mask(dx) in dxd_dtype (gradient of the dense
 * output under the dropout of BertSelfOutput/BertOutput); dgamma, dbeta, dbias (= colsum(dxd)) are
 * ACCUMULATED (+=) into fp32 [cols] buffers (each may be NULL).  dxd_amax (optional, one float): receives max|dxd|,
 * the input of samk_cast_scaled_f16 for the weight-gradient product that needs dxd in half. */
int samk_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y, void* y2,
                       int y2_dtype, void* y3, int y3_dtype, int rows, int cols, void* stream);
int samk_layernorm_bwd(const float* dy, const float* x, const float* gamma, float eps, float* dx, void* dxd,
                       int dxd_dtype, float drop_p, unsigned long long seed, unsigned long long offset, float* dgamma,
                       float* dbeta, float* dbias, float* partials, int rows, int cols, float* dxd_amax, void* stream);
/* floats of scratch for `partials` (NULL = reduce with atomics instead of the 2-stage reduction) */
long long samk_layernorm_bwd_partials(int cols);
/* The same backward in two calls: _main leaves the block partial sums of dgamma / dbeta / dbias in `partials` (required),
 * _finalize (same rows, cols) adds them into the three [cols] buffers.  Nothing inside a backward pass reads those sums, so
 * a caller may issue _finalize on a second stream, beside the GEMM that follows (each call then needs its own scratch). */
int samk_layernorm_bwd_main(const float* dy, const float* x, const float* gamma, float eps, float* dx, void* dxd,
                            int dxd_dtype, float drop_p, unsigned long long seed, unsigned long long offset, float* dgamma,
                            float* dbeta, float* dbias, float* partials, int rows, int cols, float* dxd_amax, void* stream);
int samk_layernorm_bwd_finalize(const float* partials, int rows, int cols, float* dgamma, float* dbeta, float* dbias,
                                void* stream);
/* out = dropout(a + b) (b may be NULL); also the dropout backward with a = dout */
int samk_dropout_add(const float* a, const float* b, float* out, void* out2, int out2_dtype, int rows, int cols,
                     float drop_p, unsigned long long seed, unsigned long long offset, void* stream);
/* out[c] += sum_r x[r,c] */
int samk_colsum(const void* x, int x_dtype, long long ld, int rows, int cols, float* out, void* stream);
/* x = [rows, 3*part_cols]: column sums of the three column groups added into out0 / out1 / out2 (bias gradients of
 * the fused q|k|v projection, sa_m4c.py:554-560 backward: three separate nn.Linear biases) in one pass. */
int samk_colsum3(const void* x, int x_dtype, long long ld, int rows, int part_cols, float* out0, float* out1, float* out2,
                 void* stream);
/* BertEmbeddings of TextBert (sa_m4c.py:383): dropout(LN(word[id] + pos[t] + type[0])); backward
 * accumulates into the five gradient buffers. */
int samk_bert_embed_fwd(const long long* ids, const float* word, const float* pos, const float* type,
                        const float* gamma, const float* beta, float eps, float* out, void* out2, int out2_dtype,
                        int rows, int T, int cols, float drop_p, unsigned long long seed, unsigned long long offset,
                        void* stream);
int samk_bert_embed_bwd(const float* dout, const long long* ids, const float* word, const float* pos, const float* type,
                        const float* gamma, float eps, float* dword, float* dpos, float* dtype, float* dgamma,
                        float* dbeta, int rows, int T, int cols, float drop_p, unsigned long long seed,
                        unsigned long long offset, void* stream);
/* PrevPredEmbeddings.forward (sa_m4c.py:919-948) without the [B,V+R,768] concat: ln6 = device
 * pointers {ans_g, ans_b, ocr_g, ocr_b, emb_g, emb_b} (HOST array of 6); grads10 = HOST array
 * {d_cls_w, d_ocr_in, d_pos, d_type, d_ans_g, d_ans_b, d_ocr_g, d_ocr_b, d_emb_g, d_emb_b}, accumulated. */
int samk_prevpred_fwd(const long long* prev, const float* cls_w, const float* ocr_in, const float* pos,
                      const float* type, const float* const* ln6, float eps, float* out, int B, int D, int V, int R,
                      int cols, float drop_p, unsigned long long seed, unsigned long long offset, void* stream);
int samk_prevpred_bwd(const float* dout, const long long* prev, const float* cls_w, const float* ocr_in,
                      const float* pos, const float* type, const float* const* ln6, float eps, float* const* grads10,
                      int B, int D, int V, int R, int cols, float drop_p, unsigned long long seed,
                      unsigned long long offset, void* stream);
/* OcrPtrNet scoring (sa_m4c.py:891-893): out[b,t,col_off+r] = q[b,t].k[b,r]/sqrt(dq) + (1-mask[b,r])*-1e4 */
int samk_ptr_scores_fwd(const float* q, const float* k, const long long* ocr_mask, float* out, long long ldo,
                        int col_off, int B, int D, int R, int dq, void* stream);
int samk_ptr_scores_bwd(const float* dscores, long long ldds, int col_off, const float* q, const float* k, float* dq_,
                        float* dk_, int B, int D, int R, int dq, void* stream);
/* M4CDecodingBCEWithMaskLoss (sam/task_utils.py:19-30): loss_out[0] = loss; dscores (optional) =
 * d loss / d scores; scratch = 1 float. rows = B*D, ncls = V+R. */
int samk_bce_loss(const float* scores, const float* targets, const float* loss_mask, float* dscores, float* loss_out,
                  float* scratch, int rows, int ncls, void* stream);
int samk_scale_inplace(float* x, long long n, const float* scale_dev, void* stream);
/* Row segments of a joint [B, L, d] fp32 tensor <-> separate contiguous [B, rows_k, d] tensors (n_seg <= 4; segs /
 * rows / offs are HOST arrays).  to_joint = 1 builds the MMT input [txt ; obj ; ocr ; dec] (sa_m4c.py:790) or scatters
 * gradients of row slices into a joint buffer; to_joint = 0 is the inverse (the backward of the concatenation; the
 * decoder / OCR rows the output heads read, sa_m4c.py:852-862). */
int samk_row_segments_f32(float* joint, float* const* segs, const int* rows, const int* offs, int n_seg, int B, int L, int d,
                          int to_joint, void* stream);
/* key-valid bytes [B, T+O+R+D] of MMT.forward (sa_m4c.py:793-795): masks != 0, decoder part zero */
int samk_key_valid(const long long* q_mask, const long long* obj_mask, const long long* ocr_mask, uint8_t* out, int B, int T,
                   int O, int R, int D, void* stream);
/* stream-ordered zero fill (gradient buffers) */
int samk_memset0(void* p, long long bytes, void* stream);
/* Greedy prediction without moving the logits: idx_out[r] = argmax_c x[r, c] (first maximum; sa_m4c.py:299-301,
 * sam/datasets/metrics.py:26); hit_out[r] (optional) = targets[r, idx] (token-level hit against the soft targets). */
int samk_argmax_rows(const float* x, long long ld, long long rows, int ncls, const float* targets, long long ldt, long long* idx_out,
                     float* hit_out, void* stream);
/* One step of beam search (sam/beam_search.py:88-130; the decoder the reference ships disabled, train.py:222): per sample
 * the K best of K x ncls candidates log sigmoid(scores[b K + k, c]) + beam_scores[b K + k], completed beams continue
 * with EOS only, at the first step only beam 0 counts.  scores points at step t of [B K, D, ncls] (row_stride = D ncls).
 * Outputs [B K]: source beam row, chosen class, accumulated score (descending per sample). */
int samk_beam_step(const float* scores, long long row_stride, int ncls, const float* beam_scores, const uint8_t* completed,
                   int eos, int first_step, int B, int K, long long* prev_pos, long long* new_pos, float* new_scores, void* stream);
/* out = [a ; b ; c], n floats each: the three nn.Linear biases of the fused q|k|v projection (sa_m4c.py:429-431) */
int samk_concat3_f32(const float* a, const float* b, const float* c, float* out, int n, void* stream);

/* ---- optimizer step on flat fp32 buffers ------------------------------------------------------
 * The caller's side of the path (train.py:139-143): clip_gradients (sam/task_utils.py:33-34 = clip_grad_norm_ over all
 * parameters, coefficient min(max_norm / (norm + 1e-6), 1)) fused into torch.optim.Adam's update (task_utils.py:42,
 * defaults betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad).
 * samk_sumsq:     *out_accum += sum x[i]^2 (double; zero it first; call once per gradient range).
 * samk_adam_step: one parameter range with one learning rate (a param group of get_optimizer_parameters,
 *                 sa_m4c.py:349-371); step = 1-based update count; grad_sumsq NULL = no clipping. grad is not modified. */
int samk_sumsq(const float* x, long long n, double* out_accum, void* stream);
int samk_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1,
                   double beta2, double eps, int step, const double* grad_sumsq, double max_norm, void* stream);

/* ---- timing events usable inside a stream capture ------------------------------------------------
 * samk_timing_event_record on a capturing stream records with cudaEventRecordExternal (an event-record node of the graph);
 * after a replay + synchronisation samk_timing_event_elapsed_ms gives the time between two of them: per-kernel
 * durations of the REPLAYED step (bench.py's roofline figures). */
int samk_timing_event_create(void** ev);
int samk_timing_event_record(void* ev, void* stream);
int samk_timing_event_elapsed_ms(void* start, void* end, float* ms);
int samk_timing_event_destroy(void* ev);

/* ---- gradient exchange over NVLink peer memory ---------------------------------------------------
 * Replaces the gradient reduction of nn.DataParallel (/root/reference/train.py:111-112): flat[lo:hi] = SUM over the ranks
 * of flat[lo:hi], through a symmetric "wire" allocation that every rank maps from every other rank (and, on NVSwitch,
 * once more as a multicast address: the switch then adds the ranks' copies, multimem.ld_reduce / multimem.st).
 * Five small launches on `stream` (pack | barrier | reduce this rank's shard | barrier | unpack); no shared memory and few
 * registers, so the blocks run on SMs whose shared memory the backward GEMMs own.  Every rank must issue the same calls
 * in the same order.  lo, hi: multiples of 8 elements (bf16 wire) / 4 (f32 wire); the wire holds element i of the flat
 * index space at offset i.  A rank that does not arrive within timeout_clocks sets *error (no hang). */
typedef struct samk_peer_wire {
  int rank, world;
  int wire_dtype;                     /* SAMK_DT_BF16 (sum accumulated in fp32, rounded once) or SAMK_DT_F32 */
  void* const* wire_peers;            /* HOST array [world]: the wire allocation of rank r as mapped here; [rank] = local copy */
  void* wire_mc;                      /* multicast mapping of the wire allocation, or NULL (peer loads / stores instead) */
  unsigned int* const* flag_peers;    /* HOST array [world]: a symmetric pad of >= world uint32 per rank, zeroed before the
                                         first call (by every rank, followed by a host-side barrier) */
  unsigned int* epoch;                /* local device uint32, zeroed before the first call */
  int* error;                         /* local device int, zeroed before the first call */
  long long timeout_clocks;           /* 0 = 2e10 SM clocks (~10 s); after the first miss no barrier waits again */
} samk_peer_wire;
int samk_exchange_sum(const samk_peer_wire* w, float* flat, long long lo, long long hi, void* stream);

/* ---- masked multi-head attention -------------------------------------------------------------
 * Replaces SpatialBertSelfAttention.forward steps (1),(3)-(7) (sa_m4c.py:475-552, 562-598) and
 * the BertSelfAttention of the 'n' layers / TextBert (sa_m4c.py:743, 391); the [B,L,L,H] masks are
 * never materialised (see csrc/attn_mask.cuh for the boolean restatement).
 *   qkv [B,L,3*H*64] (q|k|v), ctx [B,L,H*64], lse [B,H,L] fp32, key_valid uint8 [B,L],
 *   rel_bits uint16 [B,A,A] (bit h = head h may attend i->j; NULL when spatial == 0),
 *   quadrant_mask bit (3*seg_i+seg_j), seg 0/1/2 = text/entity/decoder (attention_mask_quadrants q
 *   of the yml maps to bit q-1).  Backward: dctx -> dqkv (same layout), delta [B,H,L] scratch.
 * impl 0 = product kernel for the dtype (f16 q|k|v + bf16 gradients: tensor cores; f32: exact fp32), 1 = force the
 * exact fp32-arithmetic SIMT kernel on the same storage formats (on-device cross-check of the tensor-core kernels). */
typedef struct samk_attn_params {
  const void* qkv; void* ctx; float* lse;
  const void* dctx; void* dqkv; float* delta;
  int dtype;                 /* SAMK_DT_* of qkv / ctx: F16 (tensor-core kernels) or F32 (exact kernel) */
  int grad_dtype;            /* SAMK_DT_* of dctx / dqkv: BF16 (tensor-core kernels) or F32 (exact kernel) */
  int B, H, head_dim;
  int T, A, D;
  const uint8_t* key_valid;
  const uint16_t* rel_bits;
  unsigned int quadrant_mask;
  int spatial;
  float scale;
  float drop_p;
  unsigned long long drop_seed, drop_offset;
  const uint32_t* allow_bits; /* tensor-core path: [B, spatial?H:1, L, ceil(L/32)] from samk_attn_build_mask */
  float* dq_accum;            /* tensor-core backward: fp32 [B*L, H*64] scratch for the dQ reduction */
  int q_begin;                /* forward only: compute query rows >= q_begin (rounded down to the kernel's
                                 row tile); 0 = all rows.  Used by the cached greedy decoder (decoder rows only). */
  int bwd_phase;              /* tensor-core backward: 0 = whole backward; 2 = only the preparation kernel (do_f16, do_inv_scale,
                                 delta from dctx and ctx); 1 = the rest, after a phase-2 call on the same workspaces */
  const uint32_t* keep_bits;  /* tensor-core path with drop_p > 0: [B, H, L, ceil(L/32)] dropout keep bits of the
                                 attention probabilities from samk_attn_build_keep (same (seed, offset) stream as the
                                 exact kernel draws inline); shared by the forward and the backward launch */
  void* do_f16;               /* tensor-core backward workspace: f16 [B*L, H*64], dctx of every (sample, head) scaled by
                                 an exact power of two into the half range (written by the first kernel of samk_attn_bwd) */
  float* do_inv_scale;        /* tensor-core backward workspace: [B*H] floats, 1 / scale per (sample, head) */
} samk_attn_params;
int samk_attn_fwd(const samk_attn_params* p, int impl, void* stream);
int samk_attn_bwd(const samk_attn_params* p, int impl, void* stream);
/* keep_bits[b, h, i, w] bit k = attention probability (i, 32 w + k) survives dropout(p->drop_p) of sa_m4c.py:588 for
 * the stream (p->drop_seed, p->drop_offset); uses B, H, T, A, D, drop_* of p only.  One launch per layer and step,
 * typically beside the q|k|v projection GEMM. */
int samk_attn_build_keep(const samk_attn_params* p, uint32_t* keep_bits, void* stream);
/* allow-bit matrix shared by all layers of one kind in a step: bit (j&31) of word [b][h|0][i][j>>5] set
 * iff query i may attend key j (key validity, decoder causality, quadrants, relation bits). */
long long samk_attn_mask_words(int B, int H, int T, int A, int D, int spatial);
int samk_attn_build_mask(const samk_attn_params* p, uint32_t* allow_bits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SAMK_H_ */
