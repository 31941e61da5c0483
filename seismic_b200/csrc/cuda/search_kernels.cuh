// Instantiations of k_search live in their own translation units (inst_*.cu, compiled in parallel); the driver
// (sgpu_api.cu) picks one through these functions.
#pragma once
#include "search.cuh"

namespace sgpu {

typedef void (*kern_t)(const SearchArgs);
enum QueryKind { Q_DENSE = 0, Q_BYTE = 1, Q_HASH = 2, Q_RANK = 3, Q_SORTED = 4 };

// u16 components, f16 values (the benchmark layout): every query representation
kern_t pick_rec16(QueryKind q, bool small_k);
kern_t pick_rec16_tma(bool small_k);  // Q_BYTE with TMA-staged records (k_search<..., TMA = true>)
// u32 components, f16 values (SeismicIndexLV): Q_RANK, Q_SORTED
kern_t pick_rec32(QueryKind q, bool small_k);
// DotVByte: Q_BYTE, Q_SORTED
kern_t pick_vb(QueryKind q, bool small_k);
// u16 components, value_kind in {BF16, F32, FIXEDU8, FIXEDU16}: Q_BYTE, Q_SORTED
kern_t pick_rec16v(uint32_t value_kind, QueryKind q, bool small_k);
// u32 components, value_kind in {BF16, F32, FIXEDU8, FIXEDU16}: Q_RANK, Q_SORTED
kern_t pick_rec32v(uint32_t value_kind, QueryKind q, bool small_k);

// exact (brute-force) top-k: k_exact_partial<Q, R> of exact.cuh
struct ExactArgs;
typedef void (*exact_t)(const ExactArgs);
exact_t pick_exact_rec16(bool dense);          // u16 / f16: dense f32 query when it fits shared memory, else sorted query
exact_t pick_exact_rec32();                    // u32 / f16
exact_t pick_exact_rec16v(uint32_t value_kind);
exact_t pick_exact_rec32v(uint32_t value_kind);

#define SGPU_K(T, OCC, Q, R) (small_k ? (kern_t)k_search<T, OCC, 2, Q, RegHeap, R> : (kern_t)k_search<T, OCC, 2, Q, SmemHeap, R>)

}  // namespace sgpu
