// Device code of the Seismic query hot path for sm_100a (B200).
//
// Data layout in HBM (built once by sgpu_index_create, see sgpu_api.cu):
//   lists[dim]        ListHdr: bases of the list's slices in the arrays below
//   postings[P]       u64 (rec_start32 << 16) | nnz     — same packing as the reference's
//                     PackedPostingBlock (src/posting_list.rs:38-52) but `start` counts 32-byte
//                     units of the record buffer instead of elements
//   blk_post_off      u32, per list B+1 entries (block_offsets, relative to the list's postings)
//   blk_min/blk_quant f32 per block (QuantizedSummary::minimums / quants)
//   sc_comp           u32 ascending summary component ids per list (component_ids)
//   sc_run_off        u32, per list n_sc+1 run bounds into the entry arrays (offsets)
//   ent_blk/ent_code  u16 block id / u8 code per summary entry (summaries_ids / values)
//   fwd records       one record per document, 32-byte aligned, made of 32-byte chunks
//                     [8 x u16 component | 8 x f16 value]; the tail chunk is padded with
//                     (component 0, value +0.0).  A group of 8 lanes reads 8 consecutive chunks
//                     = 256 contiguous bytes with two 128-bit loads per lane; every fetched
//                     32-byte sector is fully used.
//   rec_start[N+1]    u32 record start (32-byte units) per doc, for key -> doc id mapping
//
// Kernels (all hand written; no tensor cores — this is a gather/reduce path bounded by HBM):
//   k_prep    one warp per query: validation + top-`query_cut` term selection
//             (reference src/inverted_index.rs:172-175,187-190)
//   k_est     one warp per (query, term): QuantizedSummary::distances (Loop A,
//             reference src/quantized_summary.rs:64-160), bit-exact accumulation order
//   k_order   one CTA per query: block order of the first list by estimate, descending
//             (reference sort_and_search, src/posting_list.rs:162-166)
//   k_search  persistent, one CTA per SM: block traversal with skip test, forward-index scoring,
//             bounded top-k (Loop B; reference src/posting_list.rs:115-215, src/utils.rs:12-66)
//   k_finish  key -> doc id (id_from_range, reference src/inverted_index.rs:227-233)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgpu {

struct ListHdr {
    uint64_t post_base;  // into postings
    uint64_t ent_base;   // into ent_blk / ent_code
    uint64_t sc_base;    // into sc_comp ; run offsets start at sc_base + list id
    uint64_t blk_base;   // into blk_min / blk_quant ; blk_post_off starts at blk_base + list id
    uint32_t n_blk;
    uint32_t n_sc;
    uint32_t n_post;
    uint32_t pad;
};

struct DevIndex {
    const ListHdr* lists;
    const uint64_t* postings;
    const uint32_t* blk_post_off;
    const float* blk_min;
    const float* blk_quant;
    const uint32_t* sc_comp;
    const uint32_t* sc_run_off;
    const uint16_t* ent_blk;
    const uint8_t* ent_code;
    const uint4* fwd;           // record buffer, 2 x uint4 per chunk
    const uint32_t* rec_start;  // [n_docs+1]
    uint64_t n_docs;
    uint32_t dim;
};

struct Batch {
    const uint64_t* q_off;
    const uint32_t* q_comps;
    const float* q_vals;
    uint32_t nq;      // queries in this chunk
    uint32_t q_base;  // first query of the chunk inside the caller's batch
};

struct Scratch {
    uint32_t* terms;     // [nq_chunk * cut_eff] list ids, best first
    uint32_t* nterms;    // [nq]
    uint32_t* status;    // [nq] 0 ok, 1 invalid
    float* est;          // [nq_chunk * cut_eff * est_stride]
    uint16_t* order;     // [nq_chunk * est_stride]
    uint32_t* counters;  // [0] work counter, [1] max nterms, [2] invalid queries
    uint32_t* out_keys;  // [nq_chunk * k]
    unsigned long long* stats;  // [4] docs_scored, blocks_scored, blocks_pushed, fwd_units
    uint32_t est_stride;
    uint32_t cut_eff;
};

__device__ __forceinline__ uint32_t total_key(float f) {  // f32::total_cmp as unsigned key
    uint32_t x = __float_as_uint(f);
    return (x & 0x80000000u) ? ~x : (x | 0x80000000u);
}

// ------------------------------------------------------------------------------------------
// k_prep: one warp per query.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_prep(Batch b, uint32_t dim, uint32_t query_cut, uint32_t* nterms,
                                              uint32_t* status, uint32_t* counters) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= b.nq) return;
    const uint64_t o = b.q_off[b.q_base + q];
    const uint64_t n = b.q_off[b.q_base + q + 1] - o;
    bool bad = false;
    for (uint64_t i = lane; i < n; i += 32) {
        uint32_t c = b.q_comps[o + i];
        if (c >= dim) bad = true;
        if (i > 0 && b.q_comps[o + i - 1] > c) bad = true;  // is_sorted(): non-decreasing
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        uint32_t nt = bad ? 0u : (uint32_t)(n < (uint64_t)query_cut ? n : (uint64_t)query_cut);
        nterms[q] = nt;
        status[q] = bad ? 1u : 0u;
        atomicMax(&counters[1], nt);
        if (bad) atomicAdd(&counters[2], 1u);
    }
}

// Term selection: descending value under total_cmp, ties by position (components are sorted, so
// this is "smaller component first").  One warp per query; `cut_eff` rounds of warp arg-max.
__global__ void __launch_bounds__(128) k_terms(Batch b, const uint32_t* nterms, uint32_t cut_eff, uint32_t* terms) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= b.nq) return;
    const uint32_t nt = nterms[q];
    const uint64_t o = b.q_off[b.q_base + q];
    const uint32_t n = (uint32_t)(b.q_off[b.q_base + q + 1] - o);
    // (key, pos) of the previously selected term; next term is the max strictly after it in the order
    uint64_t last = ~0ull;  // composite (key << 32) | (~pos): larger == earlier in the order
    for (uint32_t t = 0; t < nt; ++t) {
        uint64_t best = 0;
        bool have = false;
        for (uint32_t i = lane; i < n; i += 32) {
            uint64_t c = ((uint64_t)total_key(b.q_vals[o + i]) << 32) | (uint32_t)(~i);
            if ((t == 0 || c < last) && (!have || c > best)) best = c, have = true;
        }
        for (int s = 16; s > 0; s >>= 1) {
            uint64_t ob = __shfl_xor_sync(0xffffffffu, best, s);
            bool oh = __shfl_xor_sync(0xffffffffu, (int)have, s);
            if (oh && (!have || ob > best)) best = ob, have = true;
        }
        last = best;
        if (lane == 0) terms[(uint64_t)q * cut_eff + t] = b.q_comps[o + (uint32_t)(~(uint32_t)best)];
    }
}

// ------------------------------------------------------------------------------------------
// k_est: Loop A.  One warp per (query, term).  est[s] accumulates, for the query components present
// in the list's summaries IN ASCENDING COMPONENT ORDER, ((code * quant[s]) + min[s]) * qv with four
// separate roundings (Rust does not contract to FMA).  A summary id occurs at most once per component,
// so lanes of one component never collide; components are serialised with __syncwarp().
// The accumulators live in shared memory when the list has <= EST_SMEM blocks, else directly in the
// global scratch (same algorithm, L2-coherent accesses).
// ------------------------------------------------------------------------------------------
constexpr int EST_WARPS = 4;
constexpr int EST_SMEM = 2048;  // blocks per warp kept in shared memory (4 x 8 KB)

__global__ void __launch_bounds__(EST_WARPS * 32) k_est(DevIndex ix, Batch b, Scratch sc) {
    __shared__ float s_est[EST_WARPS][EST_SMEM];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t task = blockIdx.x * EST_WARPS + w;
    const uint32_t q = task / sc.cut_eff, t = task % sc.cut_eff;
    if (q >= b.nq || t >= sc.nterms[q]) return;
    const uint32_t l = sc.terms[(uint64_t)q * sc.cut_eff + t];
    const ListHdr h = ix.lists[l];
    const uint32_t B = h.n_blk;
    float* g_est = sc.est + ((uint64_t)q * sc.cut_eff + t) * sc.est_stride;
    const bool in_smem = B <= EST_SMEM;
    float* acc = in_smem ? s_est[w] : g_est;
    for (uint32_t i = lane; i < B; i += 32) acc[i] = 0.f;
    __syncwarp();
    if (!in_smem) __threadfence_block();
    const uint64_t o = b.q_off[b.q_base + q];
    const uint32_t n = (uint32_t)(b.q_off[b.q_base + q + 1] - o);
    const uint32_t* scomp = ix.sc_comp + h.sc_base;
    const uint32_t* run = ix.sc_run_off + h.sc_base + l;
    const uint16_t* eb = ix.ent_blk + h.ent_base;
    const uint8_t* ec = ix.ent_code + h.ent_base;
    const float* mins = ix.blk_min + h.blk_base;
    const float* quants = ix.blk_quant + h.blk_base;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        int32_t found = -1;
        float qv = 0.f;
        if (i < n) {
            const uint32_t c = b.q_comps[o + i];
            qv = b.q_vals[o + i];
            const bool dup = i > 0 && b.q_comps[o + i - 1] == c;  // the merge consumes the first duplicate only
            if (!dup) {
                uint32_t lo = 0, hi = h.n_sc;
                while (lo < hi) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (__ldg(scomp + mid) < c) lo = mid + 1;
                    else hi = mid;
                }
                if (lo < h.n_sc && __ldg(scomp + lo) == c) found = (int32_t)lo;
            }
        }
        uint32_t m = __ballot_sync(0xffffffffu, found >= 0);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t ci = (uint32_t)__shfl_sync(0xffffffffu, found, src);
            const float wq = __shfl_sync(0xffffffffu, qv, src);
            const uint32_t e0 = __ldg(run + ci), e1 = __ldg(run + ci + 1);
            for (uint32_t e = e0 + lane; e < e1; e += 32) {
                const uint32_t s = __ldg(eb + e);
                const float code = (float)__ldg(ec + e);
                const float deq = __fadd_rn(__fmul_rn(code, __ldg(quants + s)), __ldg(mins + s));
                const float add = __fmul_rn(deq, wq);
                if (in_smem) {
                    acc[s] = __fadd_rn(acc[s], add);
                } else {
                    float cur = __ldcg(acc + s);
                    __stcg(acc + s, __fadd_rn(cur, add));
                }
            }
            __syncwarp();
        }
    }
    if (in_smem)
        for (uint32_t i = lane; i < B; i += 32) g_est[i] = acc[i];
}

// ------------------------------------------------------------------------------------------
// k_order: blocks of the FIRST list of each query sorted by (estimate desc under total_cmp, block id asc).
// One CTA per query.  Bitonic sort of 64-bit composites in shared memory when B <= ORDER_SMEM, else a
// rank sort straight from global memory (correct for any B <= 65535, slow; never hit by sane configs).
// ------------------------------------------------------------------------------------------
constexpr int ORDER_THREADS = 256;
constexpr int ORDER_SMEM = 4096;

__global__ void __launch_bounds__(ORDER_THREADS) k_order(DevIndex ix, Batch b, Scratch sc) {
    __shared__ uint64_t s_key[ORDER_SMEM];
    const uint32_t q = blockIdx.x;
    if (q >= b.nq || sc.nterms[q] == 0) return;
    const uint32_t l = sc.terms[(uint64_t)q * sc.cut_eff];
    const uint32_t B = ix.lists[l].n_blk;
    const float* est = sc.est + (uint64_t)q * sc.cut_eff * sc.est_stride;
    uint16_t* out = sc.order + (uint64_t)q * sc.est_stride;
    if (B <= ORDER_SMEM) {
        uint32_t n2 = 1;
        while (n2 < B) n2 <<= 1;
        // ascending sort of ((~key) << 32 | id): smallest composite == largest estimate, then smallest id
        for (uint32_t i = threadIdx.x; i < n2; i += ORDER_THREADS)
            s_key[i] = i < B ? (((uint64_t)(~total_key(est[i])) << 32) | i) : ~0ull;
        __syncthreads();
        for (uint32_t ksz = 2; ksz <= n2; ksz <<= 1)
            for (uint32_t j = ksz >> 1; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < n2; i += ORDER_THREADS) {
                    uint32_t p = i ^ j;
                    if (p > i) {
                        uint64_t a = s_key[i], c = s_key[p];
                        bool up = (i & ksz) == 0;
                        if ((a > c) == up) s_key[i] = c, s_key[p] = a;
                    }
                }
                __syncthreads();
            }
        for (uint32_t i = threadIdx.x; i < B; i += ORDER_THREADS) out[i] = (uint16_t)(s_key[i] & 0xffffu);
    } else {
        for (uint32_t i = threadIdx.x; i < B; i += ORDER_THREADS) {
            const uint64_t mine = ((uint64_t)(~total_key(est[i])) << 32) | i;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < B; ++j) rank += ((((uint64_t)(~total_key(est[j])) << 32) | j) < mine);
            out[rank] = (uint16_t)i;
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_search: Loop B.  Persistent kernel, one CTA of SEARCH_THREADS per SM, queries fetched from an atomic
// counter.  Per query the dense f32 query vector lives in shared memory (scattered in, scattered out).
//
// Exactness.  The reference walks blocks sequentially and skips block b iff the heap is full and
// est[b] < heap_factor * theta, theta = current k-th best score (src/posting_list.rs:130-132).  theta never
// decreases, so a block that fails the test against the CURRENT theta is skipped for good; blocks that pass
// are scored speculatively in waves (all their documents in parallel), then warp 0 REPLAYS the wave in the
// reference's block order with the live heap: re-tests each block, and pushes its documents only if the
// reference would have evaluated it.  Scores of blocks that the replay skips are discarded, so the heap —
// and therefore every later decision — is bit-identical to the sequential algorithm.  The `visited` set of
// the reference only prevents re-scoring; for results it is equivalent to "never push a doc that is
// already in the heap" (a doc seen earlier is either still in the heap or has score <= theta and cannot
// re-enter, KHeap::push is strict — src/utils.rs:36-39), which is what the replay checks.
// ------------------------------------------------------------------------------------------
constexpr int SEARCH_THREADS = 1024;
constexpr int SEARCH_WARPS = SEARCH_THREADS / 32;
constexpr int GROUPS = SEARCH_THREADS / 8;  // 8-lane groups, one document each

struct SearchArgs {
    DevIndex ix;
    Batch b;
    Scratch sc;
    uint32_t k;
    float heap_factor;
    int first_sorted;
    uint32_t wave_docs;        // soft cap of documents per wave
    uint32_t first_wave_docs;  // soft cap for the first wave of a query (heap still empty)
    uint32_t buf_docs;         // capacity of the wave buffers (>= largest block, >= wave caps)
    uint32_t qd_words;         // floats reserved for the dense query (dim rounded up)
    uint64_t* g_docs;          // optional global wave buffers (when buf_docs does not fit in smem)
    float* g_scores;
    float* out_scores;         // [nq*k] (chunk-relative)
    uint32_t* out_counts;      // [nq]
};

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// acc += q[c] * v for the 8 (component, value) pairs of one chunk, ascending, mul then add (no FMA).
__device__ __forceinline__ float chunk_dot(float acc, const uint4 c, const uint4 v, const float* __restrict__ qd) {
    const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
    const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&vw[j]));
        acc = __fadd_rn(acc, __fmul_rn(qd[cw[j] & 0xffffu], f.x));
        acc = __fadd_rn(acc, __fmul_rn(qd[cw[j] >> 16], f.y));
    }
    return acc;
}

// Score one document record (nch 32-byte chunks at `rec`) with an 8-lane group; lane8 handles chunks
// lane8, lane8+8, ...  The caller reduces the 8 partial sums with xor-shuffles 4, 2, 1.
__device__ __forceinline__ float score_rec(const uint4* __restrict__ rec, uint32_t nch, uint32_t lane8,
                                           const float* __restrict__ qd) {
    float acc = 0.f;
    uint32_t m = lane8;
    // two chunks in flight per lane per trip (covers documents up to 128 components in one trip)
    for (; m + 8 < nch; m += 16) {
        const uint4 c0 = ld_stream(rec + 2 * m), v0 = ld_stream(rec + 2 * m + 1);
        const uint4 c1 = ld_stream(rec + 2 * (m + 8)), v1 = ld_stream(rec + 2 * (m + 8) + 1);
        acc = chunk_dot(acc, c0, v0, qd);
        acc = chunk_dot(acc, c1, v1, qd);
    }
    if (m < nch) {
        const uint4 c0 = ld_stream(rec + 2 * m), v0 = ld_stream(rec + 2 * m + 1);
        acc = chunk_dot(acc, c0, v0, qd);
    }
    return acc;
}
__device__ __forceinline__ float score_doc(const uint4* __restrict__ fwd, uint64_t posting, uint32_t lane8,
                                           const float* __restrict__ qd) {
    const uint32_t nnz = (uint32_t)(posting & 0xffffu);
    return score_rec(fwd + (posting >> 16) * 2, (nnz + 7) >> 3, lane8, qd);
}

__device__ __forceinline__ bool better(float s, uint32_t key, float ws, uint32_t wkey) {
    return s > ws || (s == ws && key < wkey);
}

// warp 0 only: recompute the worst retained item (lowest score, ties: largest key)
__device__ __forceinline__ void find_worst(const float* hs, const uint32_t* hk, uint32_t n, uint32_t lane, float& theta,
                                           uint32_t& wkey, uint32_t& widx) {
    float s = 0.f;
    uint32_t key = 0, idx = 0xffffffffu;
    for (uint32_t i = lane; i < n; i += 32) {
        const float si = hs[i];
        const uint32_t ki = hk[i];
        if (idx == 0xffffffffu || better(s, key, si, ki)) s = si, key = ki, idx = i;
    }
    for (int sh = 16; sh > 0; sh >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, s, sh);
        const uint32_t ok = __shfl_xor_sync(0xffffffffu, key, sh);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, idx, sh);
        if (oi != 0xffffffffu && (idx == 0xffffffffu || better(s, key, os, ok))) s = os, key = ok, idx = oi;
    }
    theta = s;
    wkey = key;
    widx = idx;
}

// Warp-cooperative KHeap::push (src/utils.rs:32-41) of up to 32 items, one per lane (`have`).  Items that are
// already retained (same key) are ignored — the `visited` equivalence explained above k_search.  The retained
// set after the call does not depend on the order in which lanes are served (total order on (score, key)).
__device__ __forceinline__ void heap_offer(bool have, const float sc, const uint32_t key, float* hs, uint32_t* hk,
                                           const uint32_t k, const uint32_t lane, uint32_t& heap_n, float& theta,
                                           uint32_t& wkey, uint32_t& widx) {
    for (;;) {
        const bool fl = heap_n == k;
        const bool c = have && (!fl || better(sc, key, theta, wkey));
        const uint32_t m = __ballot_sync(0xffffffffu, c);
        if (!m) break;
        const int src = __ffs(m) - 1;
        const float bs = __shfl_sync(0xffffffffu, sc, src);
        const uint32_t bk = __shfl_sync(0xffffffffu, key, src);
        if ((int)lane == src) have = false;
        bool dup = false;
        for (uint32_t hh = lane; hh < heap_n; hh += 32) dup |= hk[hh] == bk;
        if (__any_sync(0xffffffffu, dup)) continue;
        const uint32_t slot = fl ? widx : heap_n;
        if (lane == 0) hs[slot] = bs, hk[slot] = bk;
        if (!fl) ++heap_n;
        __syncwarp();
        if (heap_n == k) find_worst(hs, hk, heap_n, lane, theta, wkey, widx);
    }
}

// warp-cooperative: write the retained items best first (rank sort) and pad to k
__device__ __forceinline__ void heap_write_sorted(const float* hs, const uint32_t* hk, uint32_t heap_n, uint32_t k,
                                                  uint32_t lane, uint32_t* out_keys, float* out_scores) {
    for (uint32_t i = lane; i < heap_n; i += 32) {
        const float si = hs[i];
        const uint32_t ki = hk[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < heap_n; ++j) rank += better(hs[j], hk[j], si, ki);
        out_keys[rank] = ki;
        out_scores[rank] = si;
    }
    for (uint32_t i = heap_n + lane; i < k; i += 32) out_keys[i] = 0xffffffffu, out_scores[i] = -INFINITY;
}

__global__ void __launch_bounds__(SEARCH_THREADS, 1) k_search(const SearchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // carve shared memory
    float* qd = reinterpret_cast<float*>(smem_raw);
    unsigned char* p = smem_raw + (size_t)a.qd_words * 4;
    uint32_t* cand_blk = reinterpret_cast<uint32_t*>(p);  p += SEARCH_THREADS * 4;
    uint32_t* cand_end = reinterpret_cast<uint32_t*>(p);  p += SEARCH_THREADS * 4;
    float* cand_est = reinterpret_cast<float*>(p);        p += SEARCH_THREADS * 4;
    float* heap_s = reinterpret_cast<float*>(p);          p += ((a.k + 3) & ~3u) * 4;
    uint32_t* heap_k = reinterpret_cast<uint32_t*>(p);    p += ((a.k + 3) & ~3u) * 4;
    uint64_t* docs = a.g_docs ? a.g_docs + (size_t)blockIdx.x * a.buf_docs : reinterpret_cast<uint64_t*>(p);
    if (!a.g_docs) p += (size_t)a.buf_docs * 8;
    float* scores = a.g_scores ? a.g_scores + (size_t)blockIdx.x * a.buf_docs : reinterpret_cast<float*>(p);

    __shared__ uint32_t s_q;
    __shared__ uint32_t s_warp_docs[SEARCH_WARPS], s_warp_cnt[SEARCH_WARPS];
    __shared__ uint32_t s_first_rej, s_wave_docs, s_wave_cnt;
    __shared__ float s_theta;
    __shared__ uint32_t s_full;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lane8 = tid & 7;
    const uint32_t k = a.k;

    for (uint32_t i = tid; i < a.qd_words; i += SEARCH_THREADS) qd[i] = 0.f;

    // warp-0 private heap state (registers, warp-uniform)
    uint32_t heap_n = 0, wkey = 0, widx = 0;
    float theta = 0.f;
    unsigned long long st_docs = 0, st_blocks = 0, st_pushed = 0, st_units = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_q = atomicAdd(&a.sc.counters[0], 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= a.b.nq) break;
        const uint64_t qo = a.b.q_off[a.b.q_base + q];
        const uint32_t qn = (uint32_t)(a.b.q_off[a.b.q_base + q + 1] - qo);
        const uint32_t nt = a.sc.nterms[q];  // 0 for invalid queries
        if (nt > 0)
            for (uint32_t i = tid; i < qn; i += SEARCH_THREADS) {
                const uint32_t c = a.b.q_comps[qo + i];
                if (i + 1 == qn || a.b.q_comps[qo + i + 1] != c) qd[c] = a.b.q_vals[qo + i];  // last duplicate wins
            }
        heap_n = 0;
        if (tid == 0) s_full = 0, s_theta = 0.f;
        __syncthreads();

        bool first_wave = true;
        for (uint32_t t = 0; t < nt; ++t) {
            const uint32_t l = a.sc.terms[(uint64_t)q * a.sc.cut_eff + t];
            const ListHdr h = a.ix.lists[l];
            const uint32_t B = h.n_blk;
            const float* est = a.sc.est + ((uint64_t)q * a.sc.cut_eff + t) * a.sc.est_stride;
            const uint16_t* ord = (t == 0 && a.first_sorted) ? a.sc.order + (uint64_t)q * a.sc.est_stride : nullptr;
            const uint32_t* boff = a.ix.blk_post_off + h.blk_base + l;
            const uint64_t* posts = a.ix.postings + h.post_base;
            uint32_t pos0 = 0;
            while (pos0 < B) {
                // ---------------- phase 1: candidate selection over positions [pos0, pos0 + T)
                const bool full = s_full != 0;
                const float thr = __fmul_rn(a.heap_factor, s_theta);
                const uint32_t cap = first_wave ? a.first_wave_docs : a.wave_docs;
                const uint32_t pos = pos0 + tid;
                bool pass = false;
                uint32_t blk = 0, nd = 0, p0 = 0;
                float e = 0.f;
                if (pos < B) {
                    blk = ord ? (uint32_t)ord[pos] : pos;
                    e = est[blk];
                    pass = !full || !(e < thr);
                    if (pass) {
                        p0 = boff[blk];
                        nd = boff[blk + 1] - p0;
                    }
                }
                // block-wide inclusive scans of nd and pass
                uint32_t cd = nd, cc = pass ? 1u : 0u;
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const uint32_t od = __shfl_up_sync(0xffffffffu, cd, sft);
                    const uint32_t oc = __shfl_up_sync(0xffffffffu, cc, sft);
                    if (lane >= (uint32_t)sft) cd += od, cc += oc;
                }
                if (lane == 31) s_warp_docs[warp] = cd, s_warp_cnt[warp] = cc;
                if (tid == 0) s_first_rej = 0xffffffffu;
                __syncthreads();
                if (warp == 0) {
                    uint32_t wd = s_warp_docs[lane], wc = s_warp_cnt[lane];
#pragma unroll
                    for (int sft = 1; sft < 32; sft <<= 1) {
                        const uint32_t od = __shfl_up_sync(0xffffffffu, wd, sft);
                        const uint32_t oc = __shfl_up_sync(0xffffffffu, wc, sft);
                        if (lane >= (uint32_t)sft) wd += od, wc += oc;
                    }
                    s_warp_docs[lane] = wd;
                    s_warp_cnt[lane] = wc;
                }
                __syncthreads();
                if (warp > 0) cd += s_warp_docs[warp - 1], cc += s_warp_cnt[warp - 1];
                // accept while the wave stays within its soft cap; the first passing block is always accepted
                // (buf_docs >= largest block), nothing may exceed the buffer capacity
                const bool accepted = pass && (cc == 1 || cd <= cap) && cd <= a.buf_docs;
                if (pass && !accepted) atomicMin(&s_first_rej, pos);
                __syncthreads();
                const uint32_t first_rej = s_first_rej;
                const bool in_wave = pass && pos < first_rej;
                if (in_wave) {
                    cand_blk[cc - 1] = blk;
                    cand_end[cc - 1] = cd;
                    cand_est[cc - 1] = e;
                    // ---------------- phase 2: copy the block's postings into the wave buffer
                    const uint32_t s0 = cd - nd;
                    for (uint32_t i = 0; i < nd; ++i) docs[s0 + i] = posts[p0 + i];
                }
                // totals: last thread's inclusive scan restricted to the wave == values at first_rej - 1
                if (tid == 0) s_wave_docs = 0, s_wave_cnt = 0;
                __syncthreads();
                if (in_wave) {
                    atomicMax(&s_wave_docs, cd);
                    atomicMax(&s_wave_cnt, cc);
                }
                __syncthreads();
                const uint32_t n_docs = s_wave_docs, n_cand = s_wave_cnt;
                pos0 = first_rej != 0xffffffffu ? first_rej : pos0 + SEARCH_THREADS;
                if (n_cand == 0) continue;
                first_wave = false;
                // ---------------- phase 3: score, one document per 8-lane group, two documents in flight
                for (uint32_t dbase = warp * 4; dbase < n_docs; dbase += 2 * GROUPS) {  // warp-uniform trip count
                    const uint32_t d = dbase + (lane >> 3), d1 = d + GROUPS;
                    const uint64_t pa = d < n_docs ? docs[d] : 0ull;  // nnz 0 -> no loads, score unused
                    const uint64_t pb = d1 < n_docs ? docs[d1] : 0ull;
                    float sa = score_doc(a.ix.fwd, pa, lane8, qd);
                    float sb = score_doc(a.ix.fwd, pb, lane8, qd);
                    sa = __fadd_rn(sa, __shfl_xor_sync(0xffffffffu, sa, 4));
                    sb = __fadd_rn(sb, __shfl_xor_sync(0xffffffffu, sb, 4));
                    sa = __fadd_rn(sa, __shfl_xor_sync(0xffffffffu, sa, 2));
                    sb = __fadd_rn(sb, __shfl_xor_sync(0xffffffffu, sb, 2));
                    sa = __fadd_rn(sa, __shfl_xor_sync(0xffffffffu, sa, 1));
                    sb = __fadd_rn(sb, __shfl_xor_sync(0xffffffffu, sb, 1));
                    if (lane8 == 0) {
                        if (d < n_docs) scores[d] = sa;
                        if (d1 < n_docs) scores[d1] = sb;
                    }
                }
                __syncthreads();
                // ---------------- phase 4: exact replay by warp 0
                if (warp == 0) {
                    st_docs += n_docs;
                    st_blocks += n_cand;
                    for (uint32_t i = lane; i < n_docs; i += 32) st_units += ((uint32_t)(docs[i] & 0xffffu) + 7) >> 3;
                    for (uint32_t j = 0; j < n_cand; ++j) {
                        const float ej = cand_est[j];
                        if (heap_n == k && ej < __fmul_rn(a.heap_factor, theta)) continue;
                        ++st_pushed;
                        const uint32_t s0 = j ? cand_end[j - 1] : 0u, s1 = cand_end[j];
                        for (uint32_t base = s0; base < s1; base += 32) {
                            const uint32_t i = base + lane;
                            const bool have = i < s1;
                            const float sc_i = have ? scores[i] : 0.f;
                            const uint64_t pi = have ? docs[i] : 0ull;
                            const uint32_t key = (uint32_t)(pi >> 16);
                            heap_offer(have, sc_i, key, heap_s, heap_k, k, lane, heap_n, theta, wkey, widx);
                        }
                    }
                    if (lane == 0) s_full = heap_n == k, s_theta = theta;
                }
                __syncthreads();
            }
        }
        // ---------------- results: best first (rank sort by warp 0), padded
        if (warp == 0) {
            uint32_t* ok = a.sc.out_keys + (uint64_t)q * k;
            float* os = a.out_scores + (uint64_t)q * k;
            heap_write_sorted(heap_s, heap_k, heap_n, k, lane, ok, os);
            if (lane == 0) a.out_counts[q] = heap_n;
        }
        __syncthreads();
        if (nt > 0)
            for (uint32_t i = tid; i < qn; i += SEARCH_THREADS) qd[a.b.q_comps[qo + i]] = 0.f;
    }
    if (warp == 0)
        for (int sh = 16; sh > 0; sh >>= 1) st_units += __shfl_xor_sync(0xffffffffu, st_units, sh);
    if (tid == 0 && a.sc.stats) {
        atomicAdd(&a.sc.stats[0], st_docs);
        atomicAdd(&a.sc.stats[1], st_blocks);
        atomicAdd(&a.sc.stats[2], st_pushed);
        atomicAdd(&a.sc.stats[3], st_units);
    }
}

// ------------------------------------------------------------------------------------------
// k_finish: key (record start) -> document index.  upper_bound(rec_start, key) - 1, which resolves a start
// shared with preceding EMPTY documents to the non-empty one (pinned by reference src/inverted_index.rs:716-772).
// ------------------------------------------------------------------------------------------
__global__ void k_finish(const uint32_t* rec_start, uint64_t n_docs, const uint32_t* keys, const uint32_t* counts,
                         uint32_t k, uint32_t nq, uint64_t* out_ids) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)nq * k) return;
    const uint32_t q = (uint32_t)(i / k), r = (uint32_t)(i % k);
    if (r >= counts[q]) {
        out_ids[i] = ~0ull;
        return;
    }
    const uint32_t key = keys[i];
    uint64_t lo = 0, hi = n_docs + 1;  // first index with rec_start > key
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(rec_start + mid) <= key) lo = mid + 1;
        else hi = mid;
    }
    out_ids[i] = lo - 1;
}

// ------------------------------------------------------------------------------------------
// Image construction helpers (run once per index).
// ------------------------------------------------------------------------------------------
// one warp per document: element arrays -> 32-byte chunk records, tail padded with zeros
__global__ void k_pack_records(const uint64_t* fwd_off, const uint16_t* comps, const uint16_t* vals,
                               const uint32_t* rec_start, uint64_t doc0, uint64_t n_docs_chunk, uint64_t elem0,
                               uint16_t* records /* base of the whole record buffer */) {
    const uint64_t d = doc0 + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= doc0 + n_docs_chunk) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t e0 = fwd_off[d], len = fwd_off[d + 1] - e0;
    const uint32_t nch = (uint32_t)((len + 7) >> 3);
    uint16_t* rec = records + (uint64_t)rec_start[d] * 16;
    for (uint32_t i = lane; i < nch * 8; i += 32) {
        const uint32_t ch = i >> 3, j = i & 7;
        const bool ok = i < len;
        rec[ch * 16 + j] = ok ? comps[e0 - elem0 + i] : (uint16_t)0;
        rec[ch * 16 + 8 + j] = ok ? vals[e0 - elem0 + i] : (uint16_t)0;
    }
}

// posting (element start << 16 | len) -> (record start << 16 | len)
__global__ void k_translate_postings(const uint64_t* fwd_off, const uint32_t* rec_start, uint64_t n_docs,
                                     uint64_t* postings, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t pk = postings[i];
    const uint64_t start = pk >> 16;
    uint64_t lo = 0, hi = n_docs + 1;  // upper_bound(fwd_off, start)
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (fwd_off[mid] <= start) lo = mid + 1;
        else hi = mid;
    }
    const uint64_t doc = lo - 1;
    postings[i] = ((uint64_t)rec_start[doc] << 16) | (pk & 0xffffu);
}

}  // namespace sgpu
