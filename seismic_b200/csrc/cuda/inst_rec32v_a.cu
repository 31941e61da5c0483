#include "../../../include/seismic_b200.h"
#include "search_kernels.cuh"
#include "exact.cuh"
namespace sgpu {
kern_t pick_rec32v_b(uint32_t value_kind, QueryKind q, int hk);
kern_t pick_rec32v(uint32_t value_kind, QueryKind q, int hk) {
    if (q != Q_RANK && q != Q_SORTED) return nullptr;
    const bool s = q == Q_SORTED;
    switch (value_kind) {
        case SGPU_VAL_BF16: return s ? SGPU_K1(256, 4, SortedQuery, Rec32V<1>) : SGPU_K1(256, 4, RankQuery, Rec32V<1>);
        case SGPU_VAL_FIXEDU16: return s ? SGPU_K1(256, 4, SortedQuery, Rec32V<4>) : SGPU_K1(256, 4, RankQuery, Rec32V<4>);
        default: return pick_rec32v_b(value_kind, q, hk);
    }
}
exact_t pick_exact_rec32v_b(uint32_t value_kind);
exact_t pick_exact_rec32v(uint32_t value_kind) {
    switch (value_kind) {
        case SGPU_VAL_BF16: return (exact_t)k_exact_partial<SortedQuery, Rec32V<1>>;
        case SGPU_VAL_FIXEDU16: return (exact_t)k_exact_partial<SortedQuery, Rec32V<4>>;
        default: return pick_exact_rec32v_b(value_kind);
    }
}
}  // namespace sgpu
