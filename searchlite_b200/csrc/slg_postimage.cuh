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

__global__ void __launch_bounds__(128) slg_decode_post_image_kernel(const uint8_t *img, uint64_t img_bytes,
                                                                     const PostTermHeader *hdr, uint64_t n_terms,
                                                                     const uint64_t *term_start, const uint32_t *term_blk,
                                                                     uint32_t *post_doc, uint8_t *post_tf,
                                                                     uint32_t *blk_max_doc, float *blk_max_tf,
                                                                     uint32_t *post_npos, uint32_t *post_posbyte,
                                                                     uint32_t *err) {
  const uint64_t term = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (term >= n_terms) return;
  const PostTermHeader h = hdr[term];
  const uint64_t out0 = term_start[term];
  if (h.df == 0) return;
  if (h.has_positions) {
    if (lane == 0) {
      uint64_t p = h.payload;
      for (uint32_t i = 0; i < h.df; i++) {
        uint32_t d, tf, np, x;
        if (!read_varint_seq(img, p, h.end, d) || !read_varint_seq(img, p, h.end, tf) || !read_varint_seq(img, p, h.end, np)) {
          atomicMax(err, 1u);
          break;
        }
        if (post_npos) {  // positions stay resident: where this posting's deltas start (slg_decode_positions_kernel)
          if (p - h.payload > 0xFFFFFFFFull) atomicMax(err, 3u);
          post_npos[out0 + i] = np;
          post_posbyte[out0 + i] = (uint32_t)(p - h.payload);
        }
        bool ok = true;
        for (uint32_t j = 0; j < np && ok; j++) ok = read_varint_seq(img, p, h.end, x);
        if (!ok) {
          atomicMax(err, 1u);
          break;
        }
        if (tf >= 255u) atomicMax(err, 2u);
        post_doc[out0 + i] = d;
        post_tf[out0 + i] = (uint8_t)min(tf, 255u);
      }
    }
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
            if (v >= 255u) atomicMax(err, 2u);
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
