// Microbenchmark for the question "would staging the forward-index records with TMA (cp.async.bulk, 1-D) + mbarrier
// beat the 128-bit L1::no_allocate gathers of k_search?" (VERDICT r1, item 4b).  Same access pattern as the scoring
// loop: random 32-byte aligned records of `chunks` 32-byte chunks, 8 lanes per record, two records per group in
// flight, 256-thread CTAs.
//   A  ldg     lane8 reads chunks lane8, lane8 + 8, ... with two ld.global.nc.L1::no_allocate.v4 (what k_search does)
//   B  tma     per warp a ring of STAGES x 8 records in shared memory; one elected lane issues one cp.async.bulk per
//              record (completion on the stage's mbarrier), the lanes then read their chunks with two ld.shared.v4
// Both consume the data the same way (xor).  Prints GB/s; run under ncu for the LSU wavefront counts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_tma gather_tma.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(256) k_ldg(const uint4* __restrict__ buf, const uint32_t* __restrict__ starts, uint32_t n_rec,
                                             uint32_t chunks, uint32_t* out) {
    const uint32_t lane8 = threadIdx.x & 7;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const uint32_t n_groups = (gridDim.x * blockDim.x) >> 3;
    uint32_t acc = 0;
    for (uint32_t r = group; r + n_groups < n_rec; r += 2 * n_groups) {
        const uint4* rec[2] = {buf + (uint64_t)starts[r] * 2, buf + (uint64_t)starts[r + n_groups] * 2};
        for (uint32_t m = lane8; m < chunks; m += 8) {
            uint4 c[2], v[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) { c[j] = ld_stream(rec[j] + 2 * m); v[j] = ld_stream(rec[j] + 2 * m + 1); }
#pragma unroll
            for (int j = 0; j < 2; ++j) acc ^= c[j].x ^ c[j].y ^ c[j].z ^ c[j].w ^ v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// per warp: ring of STAGES stages, a stage = 8 records (4 groups x 2) of up to REC_MAX bytes
template <int STAGES, int REC_MAX>
__global__ void __launch_bounds__(256) k_tma(const uint4* __restrict__ buf, const uint32_t* __restrict__ starts, uint32_t n_rec,
                                             uint32_t chunks, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[8][STAGES];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lane8 = lane & 7, grp = lane >> 3;
    const uint32_t gwarp = blockIdx.x * 8 + warp, n_warps = gridDim.x * 8;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem) + warp * STAGES * 8 * REC_MAX;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init((uint32_t)__cvta_generic_to_shared(&bars[warp][s]), 1);
    __syncwarp();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = chunks * 32;
    const uint32_t per_warp = 8;  // records per stage
    auto issue = [&](uint32_t it, uint32_t s) {  // records it * n_warps * 8 + gwarp * 8 + [0, 8)
        if (lane == 0) {
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[warp][s]);
            mbar_expect_tx(bar, bytes * per_warp);
            for (uint32_t j = 0; j < per_warp; ++j) {
                const uint32_t r = (it * n_warps + gwarp) * per_warp + j;
                tma_load_1d(ring + (s * 8 + j) * REC_MAX, buf + (uint64_t)starts[r] * 2, bytes, bar);
            }
        }
    };
    const uint32_t iters = n_rec / (n_warps * per_warp);
    for (uint32_t s = 0; s < STAGES && s < iters; ++s) issue(s, s);
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t s = it % STAGES, parity = (it / STAGES) & 1;
        mbar_wait((uint32_t)__cvta_generic_to_shared(&bars[warp][s]), parity);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t rec = ring + (s * 8 + grp * 2 + j) * REC_MAX;
            for (uint32_t m = lane8; m < chunks; m += 8) {
                const uint4 c = lds128(rec + 32 * m), v = lds128(rec + 32 * m + 16);
                acc ^= c.x ^ c.y ^ c.z ^ c.w ^ v.x ^ v.y ^ v.z ^ v.w;
            }
        }
        __syncwarp();  // everyone is done with the stage before it is refilled
        if (it + STAGES < iters) issue(it + STAGES, s);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main(int argc, char** argv) {
    const uint64_t buf_bytes = (argc > 1 ? atoll(argv[1]) : 8192ll) << 20;
    const uint32_t n_rec = 1u << 24;
    uint4* buf;
    cudaMalloc(&buf, buf_bytes);
    cudaMemset(buf, 1, buf_bytes);
    uint32_t* d_out;
    cudaMalloc(&d_out, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    constexpr int REC_MAX = 512;
    cudaFuncSetAttribute(k_tma<1, REC_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1 * 8 * REC_MAX);
    cudaFuncSetAttribute(k_tma<2, REC_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 8 * REC_MAX);
    for (uint32_t chunks : {8u, 15u, 16u}) {
        const uint64_t units = buf_bytes / 32;
        std::vector<uint32_t> h(n_rec);
        uint64_t s = 88172645463325252ull;
        for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s % (units - chunks)); }
        uint32_t* d_s;
        cudaMalloc(&d_s, n_rec * 4);
        cudaMemcpy(d_s, h.data(), n_rec * 4, cudaMemcpyHostToDevice);
        for (int bps : {2, 3, 4}) {
            float best[3] = {1e9f, 1e9f, 1e9f};
            for (int rep = 0; rep < 3; ++rep) {
                float ms;
                cudaEventRecord(e0);
                k_ldg<<<148 * bps, 256>>>(buf, d_s, n_rec, chunks, d_out);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best[0]) best[0] = ms;
                cudaEventRecord(e0);
                k_tma<1, REC_MAX><<<148 * bps, 256, 8 * 1 * 8 * REC_MAX>>>(buf, d_s, n_rec, chunks, d_out);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best[1]) best[1] = ms;
                cudaEventRecord(e0);
                k_tma<2, REC_MAX><<<148 * bps, 256, 8 * 2 * 8 * REC_MAX>>>(buf, d_s, n_rec, chunks, d_out);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best[2]) best[2] = ms;
            }
            cudaError_t err = cudaGetLastError();
            const double bytes = (double)n_rec * chunks * 32;
            printf("rec %4u B  256 threads x %d CTA/SM: ldg.128 %7.1f GB/s | tma ring 1 stage (%2d KB smem/CTA) %7.1f GB/s | 2 stages (%2d KB) %7.1f GB/s  %s\n",
                   chunks * 32, bps, bytes / best[0] / 1e6, 8 * 8 * REC_MAX / 1024, bytes / best[1] / 1e6,
                   16 * 8 * REC_MAX / 1024, bytes / best[2] / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
        cudaFree(d_s);
    }
    return 0;
}
