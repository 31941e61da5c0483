// Instantiations of k_search live in their own translation units (inst_*.cu, compiled in parallel); the driver
// (sgpu_api.cu) picks one through these functions.
#pragma once
#include "search.cuh"

namespace sgpu {

typedef void (*kern_t)(const SearchArgs);
// (heap kinds: see SGPU_K below)
enum QueryKind { Q_DENSE = 0, Q_BYTE = 1, Q_HASH = 2, Q_RANK = 3, Q_SORTED = 4 };

// u16 components, f16 values (the benchmark layout): every query representation
kern_t pick_rec16(QueryKind q, int hk);
kern_t pick_rec16_tma(int hk);
kern_t pick_rec16_var(int hk, int var);  // Q_BYTE, var = 10 * (CTAs / SM) + documents in flight per group: 41, 51  // Q_BYTE with TMA-staged records (k_search<..., TMA = true>)
// u32 components, f16 values (SeismicIndexLV): Q_RANK, Q_SORTED
kern_t pick_rec32(QueryKind q, int hk, int occ);  // occ: register budget (CTAs / SM) of the Q_RANK kernel
// DotVByte: Q_BYTE, Q_SORTED
kern_t pick_vb(QueryKind q, int hk, int var);
// u16 components, value_kind in {BF16, F32, FIXEDU8, FIXEDU16}: Q_BYTE, Q_SORTED
kern_t pick_rec16v(uint32_t value_kind, QueryKind q, int hk);
// u32 components, value_kind in {BF16, F32, FIXEDU8, FIXEDU16}: Q_RANK, Q_SORTED
kern_t pick_rec32v(uint32_t value_kind, QueryKind q, int hk);

// exact (brute-force) top-k: k_exact_partial<Q, R> of exact.cuh
struct ExactArgs;
typedef void (*exact_t)(const ExactArgs);
exact_t pick_exact_rec16(bool dense);          // u16 / f16: dense f32 query when it fits shared memory, else sorted query
exact_t pick_exact_rec32();                    // u32 / f16
exact_t pick_exact_rec16v(uint32_t value_kind);
exact_t pick_exact_rec32v(uint32_t value_kind);

// hk = heap kind: 0 k <= 32 (RegHeap), 1 k <= 128 (WideHeap on the layouts instantiated with SGPU_K3, else SmemHeap), 2 SmemHeap
#define SGPU_K(T, OCC, Q, R) (hk == 0 ? (kern_t)k_search<T, OCC, 2, Q, RegHeap, R> : (kern_t)k_search<T, OCC, 2, Q, SmemHeap, R>)
// one document in flight per 8-lane group: the layouts whose two-document build spills 100-400 bytes per thread
// under the 64-register budget (u32 components, and the value encodings that go through the generic mac_f path)
#define SGPU_K1(T, OCC, Q, R) (hk == 0 ? (kern_t)k_search<T, OCC, 1, Q, RegHeap, R> : (kern_t)k_search<T, OCC, 1, Q, SmemHeap, R>)
#define SGPU_K3D(T, OCC, D, Q, R)                                                               \
    (hk == 0 ? (kern_t)k_search<T, OCC, D, Q, RegHeap, R>                                       \
             : (hk == 1 ? (kern_t)k_search<T, OCC, D, Q, WideHeap, R> : (kern_t)k_search<T, OCC, D, Q, SmemHeap, R>))
#define SGPU_K3(T, OCC, Q, R)                                                                   \
    (hk == 0 ? (kern_t)k_search<T, OCC, 2, Q, RegHeap, R>                                       \
             : (hk == 1 ? (kern_t)k_search<T, OCC, 2, Q, WideHeap, R> : (kern_t)k_search<T, OCC, 2, Q, SmemHeap, R>))

}  // namespace sgpu
