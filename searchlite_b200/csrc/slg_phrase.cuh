// slg_phrase.cuh — resident term positions and phrase matching.
//
// Reference:
//   positions on disk   searchlite-core/src/index/postings.rs:117-125 (varint count | count x varint delta)
//   matches_phrase      src/query/phrase.rs:4-48
//   phrase_matches      src/api/reader.rs:1584-1597 (any variant), :1485-1518 (every phrase of a query is required)
//
// Layout: pos_begin[n_post_padded + 1] (u64, indexed like post_doc: term_start[t] + i; padding slots hold
// zero positions) and pos[] (u32, absolute positions, ascending per posting).
//
// matches_phrase searches, depth first, for positions p_0 < p_1 < ... < p_{n-1} (one per phrase term, in
// phrase order) whose gaps sum to at most `slop`: sum_i (p_i - p_{i-1} - 1) = p_{n-1} - p_0 - (n-1).  For a
// fixed p_0 the smallest reachable p_{n-1} is the greedy chain (each term takes its first position after
// the previous one), so a doc matches iff some p_0 has greedy_end(p_0) - p_0 - (n-1) <= slop.  Greedy
// chains are monotone in p_0, so one forward cursor per term visits every position once.
// A phrase becomes a doc bitmap (like a root filter); the scoring kernels AND it into `accept`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slg_kernels.cuh"
#include "slg_postimage.cuh"

namespace slg {

constexpr uint32_t kMaxPhraseTerms = 16;

struct PhraseDev {
  uint32_t n;                           // terms, phrase order
  uint32_t slop;
  uint32_t driver;                      // index of the rarest term: its postings enumerate the candidates
  uint32_t df[kMaxPhraseTerms];
  uint64_t start[kMaxPhraseTerms];      // term_start of each term
};

// Positions of the postings of a `.post` image: one CTA per 128-posting block, one thread per posting.
// post_posbyte / post_npos were recorded by slg_decode_post_image_kernel's walk.
static __global__ void __launch_bounds__(128) slg_decode_positions_kernel(const uint8_t *img, const PostTermHeader *hdr, uint64_t n_terms,
                                                                    const uint64_t *term_start, const uint32_t *term_blk,
                                                                    const uint32_t *term_df, uint32_t n_blocks,
                                                                    const uint32_t *post_posbyte, const uint64_t *pos_begin,
                                                                    uint32_t *pos, uint32_t *err) {
  const uint32_t blk = blockIdx.x;
  if (blk >= n_blocks) return;
  __shared__ uint32_t s_term;
  if (threadIdx.x == 0) {
    uint64_t lo = 0, hi = n_terms;  // last term with term_blk[t] <= blk and df > 0
    while (lo + 1 < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (term_blk[mid] <= blk) lo = mid;
      else hi = mid;
    }
    s_term = (uint32_t)lo;
  }
  __syncthreads();
  const uint32_t term = s_term;
  const uint32_t i = (blk - term_blk[term]) * kBlock + threadIdx.x;
  if (i >= term_df[term]) return;
  const PostTermHeader h = hdr[term];
  const uint64_t idx = term_start[term] + i;
  uint64_t p = h.payload + post_posbyte[idx];
  const uint64_t o0 = pos_begin[idx], o1 = pos_begin[idx + 1];
  uint32_t acc = 0;
  for (uint64_t o = o0; o < o1; o++) {
    uint32_t d;
    if (!read_varint_seq(img, p, h.end, d)) {
      atomicMax(err, 1u);
      return;
    }
    acc += d;  // index/postings.rs:192-196
    pos[o] = acc;
  }
}

// Positions handed over as CSR (slg_load_positions): counts per padded posting slot.
static __global__ void __launch_bounds__(128) slg_csr_position_counts_kernel(const uint64_t *csr_off, const uint64_t *csr_pos_off,
                                                                       uint64_t n_terms, const uint64_t *term_start,
                                                                       const uint32_t *term_blk, uint32_t n_blocks,
                                                                       uint32_t *post_npos) {
  const uint32_t blk = blockIdx.x;
  if (blk >= n_blocks) return;
  __shared__ uint32_t s_term;
  if (threadIdx.x == 0) {
    uint64_t lo = 0, hi = n_terms;
    while (lo + 1 < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (term_blk[mid] <= blk) lo = mid;
      else hi = mid;
    }
    s_term = (uint32_t)lo;
  }
  __syncthreads();
  const uint32_t term = s_term;
  const uint64_t src0 = csr_off[term];
  const uint32_t df = (uint32_t)(csr_off[term + 1] - src0);
  const uint32_t i = (blk - term_blk[term]) * kBlock + threadIdx.x;
  if (i >= df) return;
  post_npos[term_start[term] + i] = (uint32_t)(csr_pos_off[src0 + i + 1] - csr_pos_off[src0 + i]);
}

static __global__ void __launch_bounds__(128) slg_csr_position_copy_kernel(const uint64_t *csr_off, const uint64_t *csr_pos_off,
                                                                     const uint32_t *csr_pos, uint64_t n_terms,
                                                                     const uint64_t *term_start, const uint32_t *term_blk,
                                                                     uint32_t n_blocks, const uint64_t *pos_begin, uint32_t *pos) {
  const uint32_t blk = blockIdx.x;
  if (blk >= n_blocks) return;
  __shared__ uint32_t s_term;
  if (threadIdx.x == 0) {
    uint64_t lo = 0, hi = n_terms;
    while (lo + 1 < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (term_blk[mid] <= blk) lo = mid;
      else hi = mid;
    }
    s_term = (uint32_t)lo;
  }
  __syncthreads();
  const uint32_t term = s_term;
  const uint64_t src0 = csr_off[term];
  const uint32_t df = (uint32_t)(csr_off[term + 1] - src0);
  const uint32_t i = (blk - term_blk[term]) * kBlock + threadIdx.x;
  if (i >= df) return;
  const uint64_t s = csr_pos_off[src0 + i], e = csr_pos_off[src0 + i + 1];
  uint64_t o = pos_begin[term_start[term] + i];
  for (uint64_t j = s; j < e; j++) pos[o++] = csr_pos[j];
}

// first index in [0, df) of docs[] with docs[i] >= doc
__device__ __forceinline__ uint32_t phrase_lower_bound(const uint32_t *docs, uint32_t df, uint32_t doc) {
  uint32_t lo = 0, hi = df;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (docs[mid] < doc) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// One thread per posting of a phrase's driver term; a batch of phrases shares one launch: CTA b serves
// phrase i with blk_off[i] <= b < blk_off[i+1] and writes row i (row_words words) of `bits`, which must be
// zeroed.  Live / deleted docs are not consulted here (accept() checks them separately, api/reader.rs:3010).
static __global__ void __launch_bounds__(256) slg_phrase_bitmap_kernel(const PhraseDev *phrases, const uint32_t *blk_off, uint32_t n_phrases,
                                                                 const uint32_t *post_doc, const uint64_t *pos_begin,
                                                                 const uint32_t *pos, uint32_t doc_count, uint32_t *bits,
                                                                 uint64_t row_words) {
  __shared__ PhraseDev ph;
  __shared__ uint32_t s_first;
  if (threadIdx.x == 0) {
    uint32_t lo = 0, hi = n_phrases;  // last phrase with blk_off[i] <= blockIdx.x (empty phrases share an offset)
    while (lo + 1 < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (blk_off[mid] <= blockIdx.x) lo = mid;
      else hi = mid;
    }
    s_first = lo;
  }
  __syncthreads();
  const uint32_t pi = s_first;
  for (uint32_t w = threadIdx.x; w < sizeof(PhraseDev) / 4; w += blockDim.x)
    reinterpret_cast<uint32_t *>(&ph)[w] = reinterpret_cast<const uint32_t *>(&phrases[pi])[w];
  __syncthreads();
  const uint64_t i0 = (uint64_t)(blockIdx.x - blk_off[pi]) * blockDim.x + threadIdx.x;
  if (ph.n == 0 || i0 >= ph.df[ph.driver]) return;
  const uint32_t doc = post_doc[ph.start[ph.driver] + i0];
  if (doc >= doc_count) return;
  uint64_t pb[kMaxPhraseTerms];  // current cursor into pos[]
  uint64_t pe[kMaxPhraseTerms];
  for (uint32_t j = 0; j < ph.n; j++) {
    uint32_t idx;
    if (j == ph.driver) {
      idx = (uint32_t)i0;
    } else {
      const uint32_t *docs = post_doc + ph.start[j];
      idx = phrase_lower_bound(docs, ph.df[j], doc);
      if (idx >= ph.df[j] || docs[idx] != doc) return;  // phrase.rs:10-14
    }
    pb[j] = pos_begin[ph.start[j] + idx];
    pe[j] = pos_begin[ph.start[j] + idx + 1];
    if (pb[j] == pe[j]) return;  // phrase.rs:16-18
  }
  bool hit = ph.n == 1;  // phrase.rs:19-21
  for (uint64_t a = pb[0]; a < pe[0] && !hit; a++) {
    const uint32_t p0 = pos[a];
    uint32_t prev = p0;
    bool chain = true;
    for (uint32_t j = 1; j < ph.n; j++) {
      uint64_t c = pb[j];
      while (c < pe[j] && pos[c] <= prev) c++;  // phrase.rs:27-29
      pb[j] = c;                                 // a later p0 never needs an earlier position
      if (c >= pe[j]) {
        chain = false;
        break;
      }
      prev = pos[c];
    }
    if (!chain) break;  // term j has no position after this chain's prefix: none for a later start either
    if ((uint64_t)prev - p0 - (ph.n - 1) <= (uint64_t)ph.slop) hit = true;
  }
  if (hit) atomicOr(&bits[(uint64_t)pi * row_words + (doc >> 5)], 1u << (doc & 31));
}

// out = a op b over bitmap words: 0 and, 1 or, 2 and-not.  Tail bits of a and b are zero, so are out's.
static __global__ void slg_bitmap_combine_kernel(const uint32_t *a, const uint32_t *b, uint32_t words, uint32_t op, uint32_t *out) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words) return;
  const uint32_t x = a[w], y = b[w];
  out[w] = op == 0 ? (x & y) : (op == 1 ? (x | y) : (x & ~y));
}

// the same for n pairs in one launch: row i of `out` = a[i] op b[i]
static __global__ void slg_bitmap_combine_batch_kernel(const uint32_t *const *a, const uint32_t *const *b, uint32_t words, uint32_t op,
                                                uint32_t *out, uint64_t row_words) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = blockIdx.y;
  if (w >= words) return;
  const uint32_t x = a[i][w], y = b[i][w];
  out[(uint64_t)i * row_words + w] = op == 0 ? (x & y) : (op == 1 ? (x | y) : (x & ~y));
}

}  // namespace slg
