#include "search_kernels.cuh"
#include "exact.cuh"
namespace sgpu {
kern_t pick_rec32_occ(int hk, int occ);  // inst_rec32_b.cu
kern_t pick_rec32(QueryKind q, int hk, int occ) {
    switch (q) {
        case Q_RANK: return occ == 4 ? SGPU_K3(256, 4, RankQuery, Rec32) : pick_rec32_occ(hk, occ);
        case Q_SORTED: return SGPU_K1(256, 4, SortedQuery, Rec32);
        default: return nullptr;
    }
}
exact_t pick_exact_rec32() { return (exact_t)k_exact_partial<SortedQuery, Rec32>; }
}  // namespace sgpu
