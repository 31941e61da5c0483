// Exact top-k by brute force over the forward index (ground truth for recall@k; the reference's
// SeismicDataset.search via FlatIndex, src/inverted_index_wrapper.rs:721-742).  Same record format, same
// 8-lane scoring routine and same (score desc, start asc) order as k_search, so exact and approximate
// scores of one document are bit-identical and recall is a pure set comparison.
//
// Grid: one CTA per (segment, query), segment-major, so the CTAs resident at one time stream the SAME
// ~60 MB slice of the record buffer for different queries and the slice is served from the 126 MB L2.
#pragma once
#include "search.cuh"

namespace sgpu {

constexpr int EXACT_THREADS = 1024;
constexpr int EXACT_CAND = 2048;  // candidate ring per CTA (score, key)

struct ExactArgs {
    const uint4* fwd;
    const uint32_t* rec_start;
    uint64_t n_docs;
    const uint64_t* q_off;
    const uint32_t* q_comps;
    const float* q_vals;
    uint32_t nq, k, seg_docs, n_seg, qd_words;
    uint32_t chunk_units;  // rec_start units per 8-component chunk
    float value_scale;
    uint32_t* part_keys;  // [nq][n_seg][k]
    float* part_scores;
};

// Q: DenseQuery (f32[dim] in shared memory, vocabularies up to ~50 k) or SortedQuery (any vocabulary / query length,
// binary search per component); R: any plain record layout.
template <class Q, class R>
__global__ void __launch_bounds__(EXACT_THREADS, 1) k_exact_partial(const ExactArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SearchArgs sa{};
    sa.qd_words = a.qd_words;
    Q query;
    query.template init<EXACT_THREADS>(smem_raw, sa, threadIdx.x);
    unsigned char* p = smem_raw + ((Q::bytes(sa) + 15) & ~(size_t)15);
    float* heap_s = reinterpret_cast<float*>(p);        p += ((a.k + 3) & ~3u) * 4;
    uint32_t* heap_k = reinterpret_cast<uint32_t*>(p);  p += ((a.k + 3) & ~3u) * 4;
    float* cand_s = reinterpret_cast<float*>(p);        p += EXACT_CAND * 4;
    uint32_t* cand_k = reinterpret_cast<uint32_t*>(p);
    __shared__ uint32_t s_ncand, s_full, s_wkey;
    __shared__ float s_theta;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lane8 = tid & 7;
    const uint32_t seg = blockIdx.x / a.nq, qi = blockIdx.x % a.nq;
    const uint64_t qo = a.q_off[qi];
    const uint32_t qn = (uint32_t)(a.q_off[qi + 1] - qo);
    if (tid == 0) s_ncand = 0, s_full = 0, s_theta = 0.f, s_wkey = 0;
    __syncthreads();
    const Batch bt{a.q_off, a.q_comps, a.q_vals, a.nq, 0};
    query.template stage<EXACT_THREADS>(bt, Scratch{}, qi, qo, qn, tid);
    __syncthreads();
    SmemHeap heap;
    heap.reset(a.k, heap_s, heap_k);
    const uint64_t d_lo = (uint64_t)seg * a.seg_docs;
    const uint64_t d_hi = d_lo + a.seg_docs < a.n_docs ? d_lo + a.seg_docs : a.n_docs;
    constexpr uint32_t PER_ROUND = EXACT_THREADS / 8;  // documents per round
    for (uint64_t base = d_lo; base < d_hi; base += PER_ROUND) {
        const uint64_t d = base + (tid >> 3);
        uint32_t r0 = 0, nch = 0;
        if (d < d_hi) {
            r0 = __ldg(a.rec_start + d);
            nch = (__ldg(a.rec_start + d + 1) - r0) / a.chunk_units;
        }
        float s = group_reduce(score_rec<R>(reinterpret_cast<const char*>(a.fwd) + (uint64_t)r0 * R::UNIT_BYTES, nch, lane8,
                                            query, a.value_scale));
        if (lane8 == 0 && nch > 0 && (!s_full || better(s, r0, s_theta, s_wkey))) {
            const uint32_t slot = atomicAdd(&s_ncand, 1u);
            cand_s[slot] = s;
            cand_k[slot] = r0;
        }
        __syncthreads();
        const uint32_t nc = s_ncand;
        const bool full_now = s_full != 0;
        __syncthreads();  // every warp has read the counter before anyone appends to it again (uniform decision below)
        const bool last = base + PER_ROUND >= d_hi;
        if (nc + PER_ROUND > EXACT_CAND || !full_now || last) {  // drain (uniform decision)
            if (warp == 0) {
                for (uint32_t i0 = 0; i0 < nc; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    const bool have = i < nc;
                    heap.offer(have, have ? cand_s[i] : 0.f, have ? cand_k[i] : 0u, lane);
                }
                if (lane == 0) s_ncand = 0, s_full = heap.full(), s_theta = heap.theta, s_wkey = heap.wkey;
            }
            __syncthreads();
        }
    }
    if (warp == 0) {
        const uint64_t o = ((uint64_t)qi * a.n_seg + seg) * a.k;
        heap.write_sorted(lane, a.part_keys + o, a.part_scores + o);
    }
}

// one warp per query: merge the per-segment partial top-k lists, map keys to doc ids
static __global__ void __launch_bounds__(32) k_exact_merge(const ExactArgs a, const void*, float* out_scores,
                                                    uint32_t* out_counts, uint64_t* out_ids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* heap_s = reinterpret_cast<float*>(smem_raw);
    uint32_t* heap_k = reinterpret_cast<uint32_t*>(smem_raw + ((a.k + 3) & ~3u) * 4);
    const uint32_t qi = blockIdx.x, lane = threadIdx.x;
    SmemHeap heap;
    heap.reset(a.k, heap_s, heap_k);
    const uint64_t o = (uint64_t)qi * a.n_seg * a.k;
    const uint32_t total = a.n_seg * a.k;
    for (uint32_t i0 = 0; i0 < total; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t key = i < total ? a.part_keys[o + i] : 0xffffffffu;
        const bool have = key != 0xffffffffu;
        heap.offer(have, have ? a.part_scores[o + i] : 0.f, key, lane);
    }
    __syncwarp();
    // sorted output: reuse the partial buffers of segment 0 as scratch for keys
    uint32_t* skeys = a.part_keys + o;
    heap.write_sorted(lane, skeys, out_scores + (uint64_t)qi * a.k);
    __syncwarp();
    for (uint32_t i = lane; i < a.k; i += 32) {
        uint64_t id = ~0ull;
        if (i < heap.n) {
            const uint32_t key = skeys[i];
            uint64_t lo = 0, hi = a.n_docs + 1;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (__ldg(a.rec_start + mid) <= key) lo = mid + 1;
                else hi = mid;
            }
            id = lo - 1;
        }
        out_ids[(uint64_t)qi * a.k + i] = id;
    }
    if (lane == 0) out_counts[qi] = heap.n;
}

}  // namespace sgpu
