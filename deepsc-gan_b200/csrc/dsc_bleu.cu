// K16: integer BLEU n-gram counts, one warp per sentence pair, lane per token position.
// Restates SeqtoText.sequence_to_text + remove_tags + split + nltk modified_precision on ids:
// truncate before the first <END>=2, drop <PAD>=0 <START>=1 <UNK>=3 and the empty token 4, then for
// n = 1..4 the clipped n-gram matches and max(1, #hyp n-grams).  Bit-exact integer contract.
#include "dsc_common.cuh"

namespace dsc {

__device__ __forceinline__ int clean_into(const int32_t* __restrict__ src, int len, int lane, int* dst) {
  int id = (lane < len) ? __ldg(src + lane) : 2;
  unsigned endm = __ballot_sync(0xffffffffu, id == 2);
  int end = endm ? (__ffs(endm) - 1) : 32;
  bool keep = lane < end && !(id == 0 || id == 1 || id == 3 || id == 4);
  unsigned km = __ballot_sync(0xffffffffu, keep);
  if (keep) dst[__popc(km & ((1u << lane) - 1u))] = id;
  __syncwarp();
  return __popc(km);
}

__global__ void __launch_bounds__(256)
bleu_counts_kernel(const int32_t* __restrict__ ref, int ref_len, const int32_t* __restrict__ hyp, int hyp_len,
                   int32_t* __restrict__ counts, int n) {
  __shared__ int rs[8][32], hs[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 8 + warp;
  if (s >= n) return;
  int* r = rs[warp];
  int* h = hs[warp];
  const int rl = clean_into(ref + (int64_t)s * ref_len, ref_len, lane, r);
  const int hl = clean_into(hyp + (int64_t)s * hyp_len, hyp_len, lane, h);
  // Lane i owns hypothesis position i.  eq_h / eq_r: bit j set when hyp[j] / ref[j] equals hyp[i] (one pass over the
  // cleaned sentences, shared-memory broadcasts).  An n-gram match at (i, j) is the AND of the unigram masks of
  // positions i..i+n-1 shifted by 0..n-1, so every order costs one shuffle, one shift and one AND per side instead of
  // an O(31 * n) compare loop; counts, clipping and the first-occurrence rule are popcounts of those masks.
  const int mine = (lane < hl) ? h[lane] : -1;
  unsigned eq_h = 0, eq_r = 0;
  for (int j = 0; j < hl; ++j) eq_h |= (unsigned)(h[j] == mine) << j;
  for (int j = 0; j < rl; ++j) eq_r |= (unsigned)(r[j] == mine) << j;
  unsigned mh = eq_h, mr = eq_r;                   // n-gram masks of the current order
  int out_match[4], out_total[4];
#pragma unroll
  for (int g = 1; g <= 4; ++g) {
    const int nh = hl - g + 1, nr = rl - g + 1;     // number of n-grams (may be <= 0)
    if (g > 1) {
      // lane i + g - 1 holds the unigram masks of the n-gram's last token (lanes past the end read garbage: masked below)
      mh &= __shfl_down_sync(0xffffffffu, eq_h, g - 1) >> (g - 1);
      mr &= __shfl_down_sync(0xffffffffu, eq_r, g - 1) >> (g - 1);
    }
    int contrib = 0;
    if (lane < nh) {
      const unsigned vh = mh & ((nh >= 32) ? 0xffffffffu : ((1u << nh) - 1u));
      const unsigned vr = (nr > 0) ? (mr & ((nr >= 32) ? 0xffffffffu : ((1u << nr) - 1u))) : 0u;
      const bool first = (vh & ((1u << lane) - 1u)) == 0u;          // no earlier occurrence of this n-gram in hyp
      if (first) contrib = min(__popc(vh), __popc(vr));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    out_match[g - 1] = contrib;
    out_total[g - 1] = max(1, nh);
  }
  if (lane == 0) {
    int32_t* c = counts + (int64_t)s * 10;
#pragma unroll
    for (int g = 0; g < 4; ++g) { c[g] = out_match[g]; c[4 + g] = out_total[g]; }
    c[8] = hl;
    c[9] = rl;
  }
}

}  // namespace dsc

using namespace dsc;

extern "C" int dsc_bleu_counts(const int32_t* ref, int ref_len, const int32_t* hyp, int hyp_len,
                               int32_t* counts, int n, void* stream) {
  DSC_REQUIRE(ref && hyp && counts && n >= 0, "dsc_bleu_counts: bad argument");
  DSC_REQUIRE(ref_len > 0 && ref_len <= 32 && hyp_len > 0 && hyp_len <= 32, "dsc_bleu_counts: sequence length must be in 1..32");
  if (n == 0) return DSC_OK;
  bleu_counts_kernel<<<(n + 7) / 8, 256, 0, as_stream(stream)>>>(ref, ref_len, hyp, hyp_len, counts, n);
  return check_launch("dsc_bleu_counts");
}
