#include "search_kernels.cuh"
#include "exact.cuh"
namespace sgpu {
kern_t pick_rec32(QueryKind q, bool small_k) {
    switch (q) {
        case Q_RANK: return SGPU_K(256, 4, RankQuery, Rec32);
        case Q_SORTED: return SGPU_K(256, 4, SortedQuery, Rec32);
        default: return nullptr;
    }
}
exact_t pick_exact_rec32() { return (exact_t)k_exact_partial<SortedQuery, Rec32>; }
}  // namespace sgpu
