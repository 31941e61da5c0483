#include "../../../include/seismic_b200.h"
#include "search_kernels.cuh"
#include "exact.cuh"
namespace sgpu {
kern_t pick_rec16v_b(uint32_t value_kind, QueryKind q, int hk) {
    const bool s = q == Q_SORTED;
    switch (value_kind) {
        case SGPU_VAL_F32: return s ? SGPU_K1(256, 4, SortedQuery, Rec16F32) : SGPU_K1(256, 4, ByteQuery, Rec16F32);
        case SGPU_VAL_FIXEDU8: return s ? SGPU_K1(256, 4, SortedQuery, Rec16U8) : SGPU_K1(256, 4, ByteQuery, Rec16U8);
        default: return nullptr;
    }
}
exact_t pick_exact_rec16v_b(uint32_t value_kind) {
    switch (value_kind) {
        case SGPU_VAL_F32: return (exact_t)k_exact_partial<SortedQuery, Rec16F32>;
        case SGPU_VAL_FIXEDU8: return (exact_t)k_exact_partial<SortedQuery, Rec16U8>;
        default: return nullptr;
    }
}
}  // namespace sgpu
