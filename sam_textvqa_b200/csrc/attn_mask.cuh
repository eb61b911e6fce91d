// Boolean restatement of the SA-M4C attention masks (shared by every attention kernel).
//
// Reference: MMT.forward builds the additive key mask [B,1,L,L] (sa_m4c.py:805-844: valid keys of
// the txt/obj/ocr segments, zeros for the decoder columns, lower-triangular in the dec x dec
// block) and SpatialBertSelfAttention builds the per-head spatial mask [B,L,L,H]
// (sa_m4c.py:475-552: ones, entity block <- adj, configured quadrants zeroed), combines them with
// min() and zeroes rows that end up fully masked (:566-584).  Added -10000 terms underflow to an
// exact 0 probability whenever a row keeps at least one allowed key, so "allowed" is a boolean:
//
//   key_ok(i,j)  = j in DEC ? (i in DEC and j <= i) : valid[b,j]
//   sp_ok(i,j,h) = quadrant(i,j) not masked and (i,j both in ENT ? bit h of rel[b,i-T,j-T] : true)
//
// A row with no allowed key: spatial layer -> output row is exactly 0 (:574-584); plain BertLayer
// -> every score got the same -10000, i.e. an unmasked softmax over all L keys.
#pragma once
#include <stdint.h>

namespace samk {

struct AttnMask {
  const uint8_t* valid;    // [B, L] key validity (txt|obj|ocr masks, decoder part ignored)
  const uint16_t* rel;     // [B, A, A] packed head bits, or nullptr for a non-spatial layer
  int T, A, D, L;          // segment sizes: text, entities (obj+ocr), decoder; L = T+A+D
  uint32_t quad_mask;      // bit (3*seg_i + seg_j) set = quadrant zeroed (spatial layers only)
  int spatial;
};

__device__ __forceinline__ int seg_of(const AttnMask& m, int i) { return i < m.T ? 0 : (i < m.T + m.A ? 1 : 2); }

// any_valid: whether sample b has at least one valid encoder key (degenerate-row rule above)
__device__ __forceinline__ bool attn_allowed(const AttnMask& m, int b, int h, int i, int j, bool any_valid) {
  if (j >= m.L) return false;
  const int si = seg_of(m, i), sj = seg_of(m, j);
  bool ok = (sj == 2) ? (si == 2 && j <= i) : (m.valid[(size_t)b * m.L + j] != 0);
  if (!m.spatial) {
    if (!any_valid && si != 2) return true;
    return ok;
  }
  if ((m.quad_mask >> (3 * si + sj)) & 1u) return false;
  if (si == 1 && sj == 1)
    ok = ok && ((m.rel[((size_t)b * m.A + (i - m.T)) * m.A + (j - m.T)] >> h) & 1u);
  return ok;
}

__device__ __forceinline__ bool sample_any_valid(const AttnMask& m, int b, int* sflag) {
  // block-wide: does sample b have a valid encoder key
  if (threadIdx.x == 0) *sflag = 0;
  __syncthreads();
  int f = 0;
  for (int j = threadIdx.x; j < m.T + m.A; j += blockDim.x) f |= m.valid[(size_t)b * m.L + j] != 0;
  if (f) atomicOr(sflag, 1);
  __syncthreads();
  return *sflag != 0;
}

}  // namespace samk
