// slg_postimage.cuh — K1: decode the reference's on-disk posting image on the device.
//
// Format (searchlite-core/src/index/postings.rs:78-129, little-endian):
//   u32 df | u8 has_positions | u32 block_count|0x8000_0000 | u32 max_doc | f32 max_tf
//   [ u32 block_size | u32 block_max_doc[bc] | f32 block_max_tf[bc] ]
//   df x { varint doc_id (ABSOLUTE) | varint tf | [ varint npos | npos x varint delta ] }
// varints are LEB128 (util/varint.rs:5-49).  The host parses the fixed header; the payload is
// decoded here: one warp per term, 32 bytes per step, every lane that holds a terminator byte
// assembles its own varint from the (at most 4) bytes before it.  Lists with positions have a
// data-dependent varint count per posting and are walked by one lane; when positions are kept the walk
// records, per posting, the position count and the byte offset of its deltas, and
// slg_decode_positions_kernel (slg_phrase.cuh) decodes them one thread per posting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slg_kernels.cuh"

namespace slg {

struct PostTermHeader {
  uint64_t payload;  // byte offset of the first posting varint
  uint64_t end;      // byte offset one past the list
  uint32_t df;
  uint32_t has_positions;
};

__device__ __forceinline__ bool read_varint_seq(const uint8_t *img, uint64_t &p, uint64_t end, uint32_t &out) {
  uint32_t shift = 0, value = 0;
  for (;;) {
    if (p >= end) return false;
    const uint8_t b = img[p++];
    value |= (uint32_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      out = value;
      return true;
    }
    shift += 7;
    if (shift > 28) return false;  // util/varint.rs:44-46
  }
}

constexpr uint32_t kStageBytes = 1024;  // bytes of a list with positions staged per warp and refill
constexpr uint32_t kStageWords = kStageBytes / 8;
constexpr uint32_t kStageSlack = 2;     // the window that ends at the stage's last byte starts in its last word

// 8 bytes at byte offset `off` of a staged window
__device__ __forceinline__ uint64_t stage_window(const unsigned long long *buf, uint32_t off) {
  const uint32_t idx = off >> 3, sh = (off & 7u) * 8u;
  const uint64_t w0 = buf[idx];
  if (sh == 0) return w0;
  return (w0 >> sh) | ((uint64_t)buf[idx + 1] << (64u - sh));
}

// 8 bytes of the image starting at byte p, from two aligned 64-bit loads (the image buffer carries 16 bytes of
// slack behind its last byte, so the second load never leaves it)
__device__ __forceinline__ uint64_t load_window(const uint8_t *img, uint64_t p) {
  const uint64_t a = p & ~7ull;
  const uint64_t w0 = __ldg(reinterpret_cast<const unsigned long long *>(img + a));
  const uint32_t sh = (uint32_t)(p & 7ull) * 8u;
  if (sh == 0) return w0;
  const uint64_t w1 = __ldg(reinterpret_cast<const unsigned long long *>(img + a + 8));
  return (w0 >> sh) | (w1 << (64u - sh));
}

// LEB128 u32 at the low end of window w, of which `avail` (1..8) bytes are valid: value and byte length.
// util/varint.rs:37-49: at most five bytes, high bits of the fifth dropped by the u32 shift.
__device__ __forceinline__ bool window_varint(uint64_t w, uint32_t avail, uint32_t &value, uint32_t &len) {
  unsigned long long term = ~w & 0x8080808080808080ull;  // bit 7 clear = last byte of a varint
  if (avail < 8) term &= (1ull << (8 * avail)) - 1ull;
  if (!term) return false;
  len = (uint32_t)__ffsll((long long)term) >> 3;  // terminator bit 8k+7 -> ffs = 8k+8 -> k+1 bytes
  if (len > 5) return false;
  const uint64_t x = w & ((1ull << (8 * len)) - 1ull);
  value = (uint32_t)((x & 0x7Full) | ((x >> 1) & 0x3F80ull) | ((x >> 2) & 0x1FC000ull) | ((x >> 3) & 0xFE00000ull) |
                     ((x >> 4) & 0x7F0000000ull));
  return true;
}

// A term frequency that does not fit the resident byte (tf >= 255: long documents, frequent tokens) is kept exactly in a
// side list: (term, posting number inside the term, tf).  The load path patches the block maxima and the term's tf_wide
// row from it (slg_wide_patch_*).  Entries past the capacity are only counted; the host then repeats the decode with room.
__device__ __forceinline__ void wide_tf_record(uint32_t *count, uint32_t cap, uint4 *ovf, uint32_t term, uint32_t index, uint32_t tf) {
  const uint32_t at = atomicAdd(count, 1u);
  if (at < cap) ovf[at] = make_uint4(term, index, tf, 0u);
}

// exact block maxima for blocks that hold a saturated posting (float bits of non-negative values order like ints)
static __global__ void slg_wide_patch_blockmax_kernel(const uint4 *ovf, uint32_t n, const uint32_t *term_blk, float *blk_max_tf) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint4 v = ovf[e];
  atomicMax(reinterpret_cast<int *>(blk_max_tf + term_blk[v.x] + v.y / kBlock), __float_as_int((float)v.z));
}

// tf_wide rows of the wide terms from the resident bytes, then the exact values of the saturated postings on top
static __global__ void slg_wide_from_bytes_kernel(const uint64_t *term_start, const uint32_t *term_df, const uint8_t *post_tf,
                                                  const uint32_t *wide_terms, const uint64_t *wide_off, uint32_t n_wide, uint32_t *tf_wide) {
  const uint32_t w = blockIdx.y;
  if (w >= n_wide) return;
  const uint32_t term = wide_terms[w];
  const uint64_t src0 = term_start[term];
  const uint32_t df = term_df[term];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < df; i += gridDim.x * blockDim.x) tf_wide[wide_off[w] + i] = post_tf[src0 + i];
}
static __global__ void slg_wide_patch_tf_kernel(const uint4 *ovf, uint32_t n, const uint64_t *term_wide, uint32_t *tf_wide) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint4 v = ovf[e];
  const uint64_t row = term_wide[v.x];
  if (row != ~0ull) tf_wide[row + v.y] = v.z;
}

static __global__ void __launch_bounds__(128) slg_decode_post_image_kernel(const uint8_t *img, uint64_t img_bytes,
                                                                     const PostTermHeader *hdr, uint64_t n_terms,
                                                                     const uint64_t *term_start, const uint32_t *term_blk,
                                                                     uint32_t *post_doc, uint8_t *post_tf,
                                                                     uint32_t *blk_max_doc, float *blk_max_tf,
                                                                     uint32_t *post_npos, uint32_t *post_posbyte,
                                                                     uint32_t *err, uint32_t *ovf_count, uint32_t ovf_cap,
                                                                     uint4 *ovf) {
  const uint64_t term = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (term >= n_terms) return;
  const PostTermHeader h = hdr[term];
  const uint64_t out0 = term_start[term];
  if (h.df == 0) return;
  if (h.has_positions) {
    // The number of varints per posting depends on the data (doc | tf | npos | npos deltas), so the roles of the
    // varints are only known by following the records from the start of the list.  That chain is kept, but it is
    // the ONLY serial part and it runs on decoded values in shared memory: per 1 KB stage of the list
    //   1. the warp stages the bytes (coalesced 64-bit loads); every lane finds the varint terminators (bit 7
    //      clear) of its 32 bytes with a movemask multiply, a warp scan numbers the stage's varints;
    //   2. every lane decodes the varints that end in its bytes (8-byte window ending at the terminator, start
    //      found with clz over the continuation bits) -> val[], off[] in shared memory;
    //   3. lane 0 follows record -> record through val[] (s -> s + 3 + val[s+2]: two shared loads per posting)
    //      and lists the record starts; a record that runs past the stage leaves a carry of deltas to skip;
    //   4. all lanes emit the listed postings (doc, tf, npos, byte offset of the deltas).
    // Measured on the 21 MB head list of a 1 M-doc index: byte-serial walk 1.67 s, one lane with 8-byte windows
    // 1.07 s from global / 0.87 s from shared memory (instruction-latency bound, ~1600 cycles per posting);
    // this version 0.18 s.
    __shared__ unsigned long long s_stage[4][kStageWords + kStageSlack];
    __shared__ uint32_t s_val[4][kStageBytes];
    __shared__ uint16_t s_off[4][kStageBytes];
    __shared__ uint16_t s_rs[4][kStageBytes / 3 + 3];
    const int wi = threadIdx.x >> 5;
    unsigned long long *buf = s_stage[wi];
    uint32_t *val = s_val[wi];
    uint16_t *off = s_off[wi];
    uint16_t *rs = s_rs[wi];
    uint64_t p = h.payload;  // uniform: first byte not parsed yet
    uint64_t carry = 0;      // uniform: delta varints of the last listed posting still to jump over
    uint32_t i = 0;          // uniform: postings emitted
    uint32_t bad = 0;
    while (i < h.df) {
      if (p >= h.end) {
        bad = 1;
        break;
      }
      const uint64_t base = p & ~7ull;
      const uint32_t nbytes = (uint32_t)min((uint64_t)kStageBytes, h.end - base);
      const uint32_t nwords = (nbytes + 7u) / 8u + 1u;
      for (uint32_t w = lane; w < nwords; w += 32) {
        const uint64_t a = base + 8ull * w;
        buf[w] = a + 8 <= img_bytes + 16 ? __ldg(reinterpret_cast<const unsigned long long *>(img + a)) : 0x8080808080808080ull;
      }
      __syncwarp();
      const uint32_t lo = (uint32_t)(p - base);  // bytes of the first word that belong to what came before
      // 1. terminators of my 32 bytes
      uint32_t mask = 0;
#pragma unroll
      for (uint32_t k = 0; k < 4; k++) {
        const uint32_t widx = lane * 4 + k;
        if (widx * 8 < nbytes) {
          const unsigned long long t = ~buf[widx] & 0x8080808080808080ull;
          mask |= (uint32_t)((((t >> 7) * 0x0102040810204080ull) >> 56) & 0xFFull) << (8 * k);
        }
      }
      {
        const uint32_t first_byte = lane * 32;
        const uint32_t valid = nbytes > first_byte ? min(nbytes - first_byte, 32u) : 0u;
        if (valid < 32) mask &= (1u << valid) - 1u;
        if (lo > first_byte) mask &= ~((1u << (lo - first_byte)) - 1u);  // lo <= 7: lane 0 only
      }
      const uint32_t cnt = (uint32_t)__popc(mask);
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
      }
      const uint32_t T = __shfl_sync(0xFFFFFFFFu, incl, 31);
      uint32_t end_last = mask ? lane * 32 + (31u - (uint32_t)__clz((int)mask)) + 1u : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) end_last = max(end_last, __shfl_xor_sync(0xFFFFFFFFu, end_last, o));
      if (T == 0) end_last = lo;
      // 2. decode the varints that end in my bytes
      {
        uint32_t m = mask, idx = incl - cnt;
        while (m) {
          const uint32_t t = lane * 32 + (uint32_t)__ffs((int)m) - 1u;
          m &= m - 1u;
          const uint64_t win = t >= 7 ? stage_window(buf, t - 7) : (stage_window(buf, 0) << (8 * (7 - t)));  // terminator = byte 7
          unsigned long long nterm = ~win & 0x0080808080800000ull;  // bytes 2..6 that END an earlier varint
          if (t < lo + 7) nterm |= 0x0080808080800000ull & ((1ull << (8 * (lo + 7 - t))) - 1ull);  // nothing before p belongs to it
          uint32_t len = 1;
          if (!nterm) bad = 1;  // five continuation bytes before the terminator: longer than a u32 varint (util/varint.rs:44-46)
          else len = 7u - ((63u - (uint32_t)__clzll((long long)nterm)) >> 3);
          const uint64_t x = win >> (8 * (8 - len));
          val[idx] = (uint32_t)((x & 0x7Full) | ((x >> 1) & 0x3F80ull) | ((x >> 2) & 0x1FC000ull) | ((x >> 3) & 0xFE00000ull) |
                                ((x >> 4) & 0x7F0000000ull));
          off[idx] = (uint16_t)(t + 1 - len);
          idx++;
        }
      }
      bad = __any_sync(0xFFFFFFFFu, bad != 0) ? 1u : 0u;
      __syncwarp();
      // 3. the chain of records
      uint32_t n_rec = 0;
      uint64_t p_next = p;
      if (lane == 0 && !bad) {
        uint32_t s = 0;
        const uint64_t carry0 = carry;
        if (carry) {
          const uint32_t sk = (uint32_t)min(carry, (uint64_t)T);
          s = sk;
          carry -= sk;
        }
        while (carry == 0 && i + n_rec < h.df && s + 2 < T) {
          rs[n_rec++] = (uint16_t)s;
          const uint64_t nxt = (uint64_t)s + 3ull + val[s + 2];
          if (nxt > T) {
            carry = nxt - T;
            s = T;
          } else {
            s = (uint32_t)nxt;
          }
        }
        p_next = (carry > 0 || s >= T) ? base + end_last : base + off[s];
        if (p_next == p && n_rec == 0 && carry == carry0) bad = 1;  // no progress: the list ends inside a posting
      }
      n_rec = __shfl_sync(0xFFFFFFFFu, n_rec, 0);
      p_next = __shfl_sync(0xFFFFFFFFu, p_next, 0);
      carry = __shfl_sync(0xFFFFFFFFu, carry, 0);
      bad = __shfl_sync(0xFFFFFFFFu, bad, 0);
      __syncwarp();
      if (bad) break;
      // 4. emit
      for (uint32_t j = lane; j < n_rec; j += 32) {
        const uint32_t s = rs[j];
        const uint32_t tf = val[s + 1];
        if (post_npos) {  // positions stay resident: where this posting's deltas start (slg_decode_positions_kernel)
          const uint64_t at = base + (s + 3 < T ? (uint32_t)off[s + 3] : end_last) - h.payload;
          if (at > 0xFFFFFFFFull) atomicMax(err, 3u);
          post_npos[out0 + i + j] = val[s + 2];
          post_posbyte[out0 + i + j] = (uint32_t)at;
        }
        if (tf >= 255u) wide_tf_record(ovf_count, ovf_cap, ovf, (uint32_t)term, i + j, tf);
        post_doc[out0 + i + j] = val[s];
        post_tf[out0 + i + j] = (uint8_t)min(tf, 255u);
      }
      i += n_rec;
      p = p_next;
      __syncwarp();  // the stage buffers are rewritten next
    }
    if (lane == 0 && bad) atomicMax(err, 1u);
  } else {
    // parallel path: varint v (0-based) is field v&1 of posting v>>1
    uint32_t n_done = 0;        // varints completed before this chunk
    uint32_t prev_b = 0;        // previous chunk's byte of this lane
    const uint64_t nbytes = h.end - h.payload;
    const uint32_t n_var = 2u * h.df;
    for (uint64_t c = 0; c < nbytes && n_done < n_var; c += 32) {
      const uint64_t pos = h.payload + c + lane;
      const uint32_t b = pos < h.end ? img[pos] : 0x80u;  // padding never terminates a varint
      const bool term_byte = !(b & 0x80u);
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, term_byte);
      // assemble: bytes at lane-1..lane-4 that belong to the same varint (i.e. no terminator between)
      uint32_t value = b & 0x7Fu;
      bool open = true;  // still walking backwards inside this varint
#pragma unroll
      for (int back = 1; back <= 5; back++) {
        const int src = lane - back;
        // byte `back` positions earlier: from this chunk (src >= 0) or the previous one
        const uint32_t cur = __shfl_sync(0xFFFFFFFFu, b, src & 31);
        const uint32_t old = __shfl_sync(0xFFFFFFFFu, prev_b, src & 31);
        const bool have = src >= 0 || c > 0;
        const uint32_t bb = src >= 0 ? cur : old;
        if (open && have && (bb & 0x80u)) {
          if (back == 5) {
            open = false;  // a 6-byte varint: malformed
            if (term_byte) atomicMax(err, 1u);
          } else {
            value = (value << 7) | (bb & 0x7Fu);
          }
        } else {
          open = false;
        }
      }
      if (term_byte) {
        // walking backwards shifted the later (more significant) groups up by 7 per earlier byte,
        // so `value` is already the LEB128 value
        const uint32_t v = value;
        const uint32_t vi = n_done + __popc(m & ((1u << lane) - 1u));
        if (vi < n_var) {
          const uint32_t p = vi >> 1;
          if (vi & 1u) {
            if (v >= 255u) wide_tf_record(ovf_count, ovf_cap, ovf, (uint32_t)term, p, v);
            post_tf[out0 + p] = (uint8_t)min(v, 255u);
          } else {
            post_doc[out0 + p] = v;
          }
        }
      }
      n_done += __popc(m);
      prev_b = b;
    }
    if (lane == 0 && n_done < n_var) atomicMax(err, 1u);
  }
  __syncwarp();
  __threadfence_block();
  // block-max tables from the decoded postings (equal to the stored ones for a well-formed file)
  const uint32_t nb = (h.df + kBlock - 1) / kBlock;
  for (uint32_t bk = 0; bk < nb; bk++) {
    const uint32_t s = bk * kBlock, e = min(h.df, s + kBlock);
    uint32_t mtf = 0;
    for (uint32_t i = s + lane; i < e; i += 32) mtf = max(mtf, (uint32_t)post_tf[out0 + i]);
    for (int o = 16; o > 0; o >>= 1) mtf = max(mtf, __shfl_xor_sync(0xFFFFFFFFu, mtf, o));
    if (lane == 0) {
      blk_max_tf[term_blk[term] + bk] = (float)mtf;
      blk_max_doc[term_blk[term] + bk] = post_doc[out0 + e - 1];
    }
  }
}

}  // namespace slg
