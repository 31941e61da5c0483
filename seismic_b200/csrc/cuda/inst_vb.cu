#include "search_kernels.cuh"
namespace sgpu {
kern_t pick_vb(QueryKind q, int hk) {
    switch (q) {
        case Q_BYTE: return SGPU_K(256, 4, ByteQuery, RecVB);
        case Q_SORTED: return SGPU_K(256, 4, SortedQuery, RecVB);
        default: return nullptr;
    }
}
}  // namespace sgpu
