#include "search_kernels.cuh"
#include "exact.cuh"
namespace sgpu {
kern_t pick_rec16(QueryKind q, int hk) {
    switch (q) {
        case Q_DENSE: return SGPU_K(DENSE_THREADS, 1, DenseQuery, Rec16);
        case Q_BYTE: return SGPU_K3(256, 4, ByteQuery, Rec16);
        case Q_HASH: return SGPU_K(256, 4, HashQuery, Rec16);
        case Q_RANK: return SGPU_K(256, 4, RankQuery, Rec16);
        default: return nullptr;
    }
}
kern_t pick_rec16_tma(int hk) {  // byte-index query, records staged by TMA; 3 CTAs / SM
    return hk == 0 ? (kern_t)k_search<256, 3, 2, ByteQuery, RegHeap, Rec16, true>
                   : (kern_t)k_search<256, 3, 2, ByteQuery, SmemHeap, Rec16, true>;
}
exact_t pick_exact_rec16(bool dense) {
    return dense ? (exact_t)k_exact_partial<DenseQuery, Rec16> : (exact_t)k_exact_partial<SortedQuery, Rec16>;
}
}  // namespace sgpu
