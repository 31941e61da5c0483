#include "search_kernels.cuh"
namespace sgpu {
kern_t pick_rec16(QueryKind q, bool small_k) {
    switch (q) {
        case Q_DENSE: return SGPU_K(DENSE_THREADS, 1, DenseQuery, Rec16);
        case Q_BYTE: return SGPU_K(256, 4, ByteQuery, Rec16);
        case Q_HASH: return SGPU_K(256, 4, HashQuery, Rec16);
        case Q_RANK: return SGPU_K(256, 4, RankQuery, Rec16);
        default: return nullptr;
    }
}
}  // namespace sgpu
