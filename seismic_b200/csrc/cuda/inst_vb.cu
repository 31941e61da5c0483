#include "search_kernels.cuh"
namespace sgpu {
kern_t pick_vb(QueryKind q, int hk, int var) {  // var: 4 = two documents per group; 41, 51 = one (4 / 5 CTAs per SM)
    switch (q) {
        case Q_BYTE:
            if (var == 41) return SGPU_K1(256, 4, ByteQuery, RecVB);
            if (var == 51) return SGPU_K1(256, 5, ByteQuery, RecVB);
            return SGPU_K(256, 4, ByteQuery, RecVB);
        case Q_SORTED: return SGPU_K(256, 4, SortedQuery, RecVB);
        default: return nullptr;
    }
}
}  // namespace sgpu
