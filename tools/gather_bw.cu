// Microbenchmark: achievable HBM bandwidth for random gathers of `rec_bytes`-byte records (32-byte aligned)
// with the access pattern of k_search (8 lanes per record, 16- or 32-byte loads per lane).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu
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

template <int D>
__global__ void k_gather(const uint4* __restrict__ buf, const uint32_t* __restrict__ starts, uint32_t n_rec, uint32_t chunks,
                         uint32_t* out) {
    const uint32_t lane8 = threadIdx.x & 7;
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const uint32_t n_groups = (gridDim.x * blockDim.x) >> 3;
    uint32_t acc = 0;
    for (uint32_t r = group; r + (D - 1) * n_groups < n_rec; r += D * n_groups) {
        const uint4* rec[D];
#pragma unroll
        for (int j = 0; j < D; ++j) rec[j] = buf + (uint64_t)starts[r + j * n_groups] * 2;
        for (uint32_t m = lane8; m < chunks; m += 8) {
            uint4 c[D], v[D];
#pragma unroll
            for (int j = 0; j < D; ++j) { c[j] = ld_stream(rec[j] + 2 * m); v[j] = ld_stream(rec[j] + 2 * m + 1); }
#pragma unroll
            for (int j = 0; j < D; ++j) acc ^= c[j].x ^ c[j].y ^ c[j].z ^ c[j].w ^ v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
        }
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
    for (uint32_t chunks : {4u, 8u, 15u, 16u, 32u, 64u}) {
        const uint64_t units = buf_bytes / 32;
        std::vector<uint32_t> h(n_rec);
        uint64_t s = 88172645463325252ull;
        for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (uint32_t)(s % (units - chunks)); }
        uint32_t* d_s;
        cudaMalloc(&d_s, n_rec * 4);
        cudaMemcpy(d_s, h.data(), n_rec * 4, cudaMemcpyHostToDevice);
        for (int threads : {128, 256}) for (int bps : {4, 8, 16}) {
            if (threads * bps > 2048) continue;
            float best[2] = {1e9f, 1e9f};
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                k_gather<2><<<148 * bps, threads>>>(buf, d_s, n_rec, chunks, d_out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best[0]) best[0] = ms;
                cudaEventRecord(e0);
                k_gather<4><<<148 * bps, threads>>>(buf, d_s, n_rec, chunks, d_out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1); if (ms < best[1]) best[1] = ms;
            }
            const double bytes = (double)n_rec * chunks * 32;
            printf("rec %4u B  threads %3d x %2d CTA/SM (%2d warps/SM): D=2 %7.1f GB/s   D=4 %7.1f GB/s\n", chunks * 32, threads, bps,
                   threads * bps / 32, bytes / best[0] / 1e6, bytes / best[1] / 1e6);
        }
        cudaFree(d_s);
    }
    return 0;
}
