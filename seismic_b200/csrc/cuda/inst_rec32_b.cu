#include "search_kernels.cuh"
namespace sgpu {
// The large-vocabulary kernel under other register budgets.  With two documents in flight per 8-lane group (D = 2) the
// u32 records keep 2 x 3 x 128 bits of loads per lane and the 64-register build spills ~500 bytes per thread (measured:
// 18.6 ms per 10 k queries at k = 100, 1 M docs); with ONE document in flight it needs 48-63 registers and no stack:
// 12.0 ms at 4 CTAs / SM.  occ = 10 * (CTAs / SM) + D for the D = 1 builds.
kern_t pick_rec32_occ(int hk, int occ) {
    switch (occ) {
        case 41: return SGPU_K3D(256, 4, 1, RankQuery, Rec32);
        case 51: return SGPU_K3D(256, 5, 1, RankQuery, Rec32);
        case 3: return SGPU_K3(256, 3, RankQuery, Rec32);
        case 2: return SGPU_K3(256, 2, RankQuery, Rec32);
        default: return nullptr;
    }
}
}  // namespace sgpu
