#include "search_kernels.cuh"
namespace sgpu {
// the benchmark kernel with ONE document in flight per 8-lane group: 46-56 registers, so 5 CTAs fit an SM
kern_t pick_rec16_var(int hk, int var) {
    switch (var) {
        case 41: return SGPU_K3D(256, 4, 1, ByteQuery, Rec16);
        case 51: return SGPU_K3D(256, 5, 1, ByteQuery, Rec16);
        default: return nullptr;
    }
}
}  // namespace sgpu
