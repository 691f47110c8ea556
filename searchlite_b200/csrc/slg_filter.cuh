// slg_filter.cuh — K4: fast-field filter -> bitmap (1 bit per doc).
//
// Evaluates the reference's Filter AST (searchlite-core/src/api/types.rs:670-680) with the
// semantics of query/filters.rs:84-149 over the columns of index/fastfields.rs:490-657:
// inclusive ranges, missing value => predicate false, Not inverts that, And of nothing = true,
// Or of nothing = false.  List columns (StrList / I64List / F64List, and the nested forms flattened at load):
// a predicate holds when ANY value of the doc satisfies it; a doc without values fails it.  Keyword predicates arrive as a bitset over the column's dictionary
// (the ASCII-case-insensitive string compare of fastfields.rs:475-481 is done once per dictionary
// entry on the host), so the device only compares ordinals.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slg {

constexpr uint32_t kMaxFilterNodes = 64;
constexpr uint32_t kMaxFilterDepth = 16;
constexpr uint32_t FOP_KEYWORD_EQ = 0, FOP_KEYWORD_IN = 1, FOP_I64_RANGE = 2, FOP_F64_RANGE = 3, FOP_AND = 4, FOP_OR = 5,
                   FOP_NOT = 6, FOP_KEYWORD_LIST = 7, FOP_I64_LIST = 8, FOP_F64_LIST = 9, FOP_FALSE = 100;

struct FilterNodeDev {
  uint32_t op, n_children;
  int64_t i_min, i_max;
  double f_min, f_max;
  const void *values;      // i64 / f64 / u32 ordinals, one per doc
  const uint8_t *present;  // numeric columns
  const uint32_t *offsets; // list columns: [doc_count + 1]
  uint32_t set_off, set_words;  // keyword: bitset over dictionary ordinals
};

// The program is in prefix order; walking it backwards with a value stack evaluates every
// node after its children (And/Or are commutative, so the reversed child order is immaterial).
static __global__ void slg_filter_bitmap_kernel(const FilterNodeDev *nodes, uint32_t n_nodes, const uint32_t *ordsets,
                                         uint32_t doc_count, uint32_t *bits) {
  const uint32_t doc = blockIdx.x * blockDim.x + threadIdx.x;
  bool result = false;
  if (doc < doc_count) {
    unsigned long long stack = 0;  // bit i = value at depth i
    int sp = 0;
    for (int i = (int)n_nodes - 1; i >= 0; i--) {
      const FilterNodeDev nd = nodes[i];
      bool v = false;
      switch (nd.op) {
        case FOP_KEYWORD_EQ:
        case FOP_KEYWORD_IN: {
          const uint32_t o = static_cast<const uint32_t *>(nd.values)[doc];
          v = o != 0xFFFFFFFFu && (o >> 5) < nd.set_words && ((ordsets[nd.set_off + (o >> 5)] >> (o & 31)) & 1u);
          break;
        }
        case FOP_I64_RANGE: {
          const long long x = static_cast<const long long *>(nd.values)[doc];
          v = nd.present[doc] && x >= nd.i_min && x <= nd.i_max;
          break;
        }
        case FOP_F64_RANGE: {
          const double x = static_cast<const double *>(nd.values)[doc];
          v = nd.present[doc] && x >= nd.f_min && x <= nd.f_max;
          break;
        }
        case FOP_KEYWORD_LIST: {
          const uint32_t *ords = static_cast<const uint32_t *>(nd.values);
          for (uint32_t i = nd.offsets[doc], e = nd.offsets[doc + 1]; i < e && !v; i++) {
            const uint32_t o = ords[i];
            v = (o >> 5) < nd.set_words && ((ordsets[nd.set_off + (o >> 5)] >> (o & 31)) & 1u);
          }
          break;
        }
        case FOP_I64_LIST: {
          const long long *xs = static_cast<const long long *>(nd.values);
          for (uint32_t i = nd.offsets[doc], e = nd.offsets[doc + 1]; i < e && !v; i++) v = xs[i] >= nd.i_min && xs[i] <= nd.i_max;
          break;
        }
        case FOP_F64_LIST: {
          const double *xs = static_cast<const double *>(nd.values);
          for (uint32_t i = nd.offsets[doc], e = nd.offsets[doc + 1]; i < e && !v; i++) v = xs[i] >= nd.f_min && xs[i] <= nd.f_max;
          break;
        }
        case FOP_AND: {
          v = true;
          for (uint32_t c = 0; c < nd.n_children; c++) {
            sp--;
            v = v && ((stack >> sp) & 1ull);
          }
          break;
        }
        case FOP_OR: {
          v = false;
          for (uint32_t c = 0; c < nd.n_children; c++) {
            sp--;
            v = v || ((stack >> sp) & 1ull);
          }
          break;
        }
        case FOP_NOT: {
          sp--;
          v = !((stack >> sp) & 1ull);
          break;
        }
        default:
          v = false;
      }
      stack = (stack & ~(1ull << sp)) | ((unsigned long long)v << sp);
      sp++;
    }
    result = (stack & 1ull) != 0;
  }
  const uint32_t word = __ballot_sync(0xFFFFFFFFu, result);
  if ((threadIdx.x & 31) == 0 && doc < doc_count) bits[doc >> 5] = word;
}

}  // namespace slg
