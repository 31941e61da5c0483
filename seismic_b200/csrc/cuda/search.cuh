// k_search: Loop B of the Seismic query path — block traversal with the summary skip test, forward-index
// scoring and the bounded top-k (reference src/posting_list.rs:115-215, src/utils.rs:12-66,
// src/inverted_index.rs:180-234).  Persistent kernel; CTAs fetch query ids from an atomic counter.
//
// Two instantiations share all the code below:
//   k_search<128, HashQuery>   "hq": the query lives in a 4096-slot perfect-hash table in shared memory
//                              (u16/u32 tag + f32 value, collision-free multiplier found by k_terms), ~32 KB per
//                              CTA, so 7 CTAs = 7 independent queries are resident per SM and hide each other's
//                              dependent-load and replay latencies.  Used for queries with <= 128 components.
//   k_search<1024, DenseQuery> the dense f32 query vector (dim x 4 B) in shared memory, one CTA per SM.  Handles
//                              any query; used for the (rare) queries the hash path cannot take.
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
// re-enter, KHeap::push is strict — src/utils.rs:36-39), which is what heap_offer checks.
#pragma once
#include "kernels.cuh"

namespace sgpu {

constexpr int HQ_THREADS = 128;
constexpr int HQ_LOG2_SLOTS = 12;
constexpr int HQ_SLOTS = 1 << HQ_LOG2_SLOTS;
constexpr int HQ_MAX_NNZ = 128;  // queries with more components take the dense kernel
constexpr int HQ_TRIES = 64;
constexpr int DENSE_THREADS = 1024;

__device__ __forceinline__ uint32_t hq_mult(uint32_t attempt) { return (2u * attempt + 1u) * 0x9E3779B1u; }
__device__ __forceinline__ uint32_t hq_slot(uint32_t c, uint32_t mult) { return (c * mult) >> (32 - HQ_LOG2_SLOTS); }

struct SearchArgs {
    DevIndex ix;
    Batch b;
    Scratch sc;
    const uint32_t* qlist;     // query ids (chunk-relative) this launch processes
    const uint32_t* n_list;    // device pointer to the length of qlist (written by k_terms)
    uint32_t counter_idx;      // which sc.counters[] entry is this launch's work counter
    uint32_t k;
    float heap_factor;
    int first_sorted;
    uint32_t wave_docs;        // soft cap of documents per wave
    uint32_t first_wave_docs;  // soft cap for the first wave of a query (heap still empty)
    uint32_t buf_docs;         // capacity of the wave buffers (>= largest block, >= wave caps)
    uint32_t qd_words;         // dense kernel: floats reserved for the dense query (dim rounded up)
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

// ---- query representations in shared memory -------------------------------------------------------------
struct DenseQuery {
    float* qd;
    __device__ __forceinline__ float operator()(uint32_t c) const { return qd[c]; }
    static __device__ __forceinline__ size_t bytes(const SearchArgs& a) { return (size_t)a.qd_words * 4; }
    template <int T>
    __device__ __forceinline__ void init(unsigned char* base, const SearchArgs& a, uint32_t tid) {
        qd = reinterpret_cast<float*>(base);
        for (uint32_t i = tid; i < a.qd_words; i += T) qd[i] = 0.f;
    }
    template <int T>
    __device__ __forceinline__ void stage(const Batch& b, const Scratch&, uint32_t q, uint64_t qo, uint32_t qn,
                                          uint32_t tid) {
        for (uint32_t i = tid; i < qn; i += T) {
            const uint32_t c = b.q_comps[qo + i];
            if (i + 1 == qn || b.q_comps[qo + i + 1] != c) qd[c] = b.q_vals[qo + i];  // last duplicate wins
        }
    }
    template <int T>
    __device__ __forceinline__ void unstage(const Batch& b, uint64_t qo, uint32_t qn, uint32_t tid) {
        for (uint32_t i = tid; i < qn; i += T) qd[b.q_comps[qo + i]] = 0.f;
    }
};

struct HashQuery {
    uint16_t* tags;
    float* vals;
    uint32_t mult;
    __device__ __forceinline__ float operator()(uint32_t c) const {
        const uint32_t s = hq_slot(c, mult);
        const float v = vals[s];
        return tags[s] == (uint16_t)c ? v : 0.f;
    }
    static __device__ __forceinline__ size_t bytes(const SearchArgs&) { return (size_t)HQ_SLOTS * 6; }
    template <int T>
    __device__ __forceinline__ void init(unsigned char* base, const SearchArgs&, uint32_t tid) {
        vals = reinterpret_cast<float*>(base);
        tags = reinterpret_cast<uint16_t*>(base + (size_t)HQ_SLOTS * 4);
        for (uint32_t i = tid; i < HQ_SLOTS; i += T) vals[i] = 0.f, tags[i] = 0xffffu;
        mult = 1;
    }
    template <int T>
    __device__ __forceinline__ void stage(const Batch& b, const Scratch& sc, uint32_t q, uint64_t qo, uint32_t qn,
                                          uint32_t tid) {
        mult = sc.hmult[q];
        for (uint32_t i = tid; i < qn; i += T) {
            const uint32_t c = b.q_comps[qo + i];
            if (i + 1 == qn || b.q_comps[qo + i + 1] != c) {  // last duplicate wins
                const uint32_t s = hq_slot(c, mult);
                tags[s] = (uint16_t)c;
                vals[s] = b.q_vals[qo + i];
            }
        }
    }
    template <int T>
    __device__ __forceinline__ void unstage(const Batch& b, uint64_t qo, uint32_t qn, uint32_t tid) {
        for (uint32_t i = tid; i < qn; i += T) {
            const uint32_t s = hq_slot(b.q_comps[qo + i], mult);
            tags[s] = 0xffffu;
            vals[s] = 0.f;
        }
    }
};

// acc += q[c] * v for the 8 (component, value) pairs of one chunk, ascending, mul then add (no FMA).
template <class Q>
__device__ __forceinline__ float chunk_dot(float acc, const uint4 c, const uint4 v, const Q& q) {
    const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
    const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&vw[j]));
        acc = __fadd_rn(acc, __fmul_rn(q(cw[j] & 0xffffu), f.x));
        acc = __fadd_rn(acc, __fmul_rn(q(cw[j] >> 16), f.y));
    }
    return acc;
}

// Score one document record (nch 32-byte chunks at `rec`) with an 8-lane group; lane8 handles chunks
// lane8, lane8+8, ...  The caller reduces the 8 partial sums with xor-shuffles 4, 2, 1.
template <class Q>
__device__ __forceinline__ float score_rec(const uint4* __restrict__ rec, uint32_t nch, uint32_t lane8, const Q& q) {
    float acc = 0.f;
    uint32_t m = lane8;
    // two chunks in flight per lane per trip (covers documents up to 128 components in one trip)
    for (; m + 8 < nch; m += 16) {
        const uint4 c0 = ld_stream(rec + 2 * m), v0 = ld_stream(rec + 2 * m + 1);
        const uint4 c1 = ld_stream(rec + 2 * (m + 8)), v1 = ld_stream(rec + 2 * (m + 8) + 1);
        acc = chunk_dot(acc, c0, v0, q);
        acc = chunk_dot(acc, c1, v1, q);
    }
    if (m < nch) {
        const uint4 c0 = ld_stream(rec + 2 * m), v0 = ld_stream(rec + 2 * m + 1);
        acc = chunk_dot(acc, c0, v0, q);
    }
    return acc;
}
template <class Q>
__device__ __forceinline__ float score_doc(const uint4* __restrict__ fwd, uint64_t posting, uint32_t lane8,
                                           const Q& q) {
    const uint32_t nnz = (uint32_t)(posting & 0xffffu);
    return score_rec(fwd + (posting >> 16) * 2, (nnz + 7) >> 3, lane8, q);
}
__device__ __forceinline__ float group_reduce(float s) {
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
    return __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
}

// ---- bounded top-k kept by ONE warp in shared memory -------------------------------------------------------
__device__ __forceinline__ bool better(float s, uint32_t key, float ws, uint32_t wkey) {
    return s > ws || (s == ws && key < wkey);
}

// recompute the worst retained item (lowest score, ties: largest key)
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
// already retained (same key) are ignored — the `visited` equivalence explained at the top.  The retained
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

// -----------------------------------------------------------------------------------------------------------
template <int T, class Q>
__global__ void __launch_bounds__(T, (T >= 1024 ? 1 : 7)) k_search(const SearchArgs a) {
    constexpr int NW = T / 32;      // warps
    constexpr int GROUPS = T / 8;   // 8-lane groups, one document each
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Q query;
    query.template init<T>(smem_raw, a, threadIdx.x);
    unsigned char* p = smem_raw + Q::bytes(a);
    uint32_t* cand_blk = reinterpret_cast<uint32_t*>(p);  p += T * 4;
    uint32_t* cand_end = reinterpret_cast<uint32_t*>(p);  p += T * 4;
    float* cand_est = reinterpret_cast<float*>(p);        p += T * 4;
    float* heap_s = reinterpret_cast<float*>(p);          p += ((a.k + 3) & ~3u) * 4;
    uint32_t* heap_k = reinterpret_cast<uint32_t*>(p);    p += ((a.k + 3) & ~3u) * 4;
    uint64_t* docs = a.g_docs ? a.g_docs + (size_t)blockIdx.x * a.buf_docs : reinterpret_cast<uint64_t*>(p);
    if (!a.g_docs) p += (size_t)a.buf_docs * 8;
    float* scores = a.g_scores ? a.g_scores + (size_t)blockIdx.x * a.buf_docs : reinterpret_cast<float*>(p);

    __shared__ uint32_t s_q;
    __shared__ uint32_t s_warp_docs[32], s_warp_cnt[32];
    __shared__ uint32_t s_first_rej, s_wave_docs, s_wave_cnt;
    __shared__ float s_theta;
    __shared__ uint32_t s_full;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lane8 = tid & 7;
    const uint32_t k = a.k;

    // warp-0 private heap state (registers, warp-uniform)
    uint32_t heap_n = 0, wkey = 0, widx = 0;
    float theta = 0.f;
    unsigned long long st_docs = 0, st_blocks = 0, st_pushed = 0, st_units = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_q = atomicAdd(&a.sc.counters[a.counter_idx], 1u);
        __syncthreads();
        if (s_q >= *a.n_list) break;
        const uint32_t q = a.qlist[s_q];
        const uint64_t qo = a.b.q_off[a.b.q_base + q];
        const uint32_t qn = (uint32_t)(a.b.q_off[a.b.q_base + q + 1] - qo);
        const uint32_t nt = a.sc.nterms[q];  // 0 for invalid queries
        if (nt > 0) query.template stage<T>(a.b, a.sc, q, qo, qn, tid);
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
                if (tid == 0) s_wave_docs = 0, s_wave_cnt = 0;  // every thread has consumed the previous wave's totals
                if (warp == 0) {
                    uint32_t wd = lane < NW ? s_warp_docs[lane] : 0u, wc = lane < NW ? s_warp_cnt[lane] : 0u;
#pragma unroll
                    for (int sft = 1; sft < NW; sft <<= 1) {
                        const uint32_t od = __shfl_up_sync(0xffffffffu, wd, sft);
                        const uint32_t oc = __shfl_up_sync(0xffffffffu, wc, sft);
                        if (lane >= (uint32_t)sft) wd += od, wc += oc;
                    }
                    if (lane < NW) s_warp_docs[lane] = wd, s_warp_cnt[lane] = wc;
                }
                __syncthreads();
                if (warp > 0) cd += s_warp_docs[warp - 1], cc += s_warp_cnt[warp - 1];
                // accept while the wave stays within its soft cap; the first passing block is always accepted
                // (buf_docs >= largest block); nothing may exceed the buffer capacity
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
                    atomicMax(&s_wave_docs, cd);
                    atomicMax(&s_wave_cnt, cc);
                }
                __syncthreads();
                const uint32_t n_docs = s_wave_docs, n_cand = s_wave_cnt;
                pos0 = first_rej != 0xffffffffu ? first_rej : pos0 + T;
                if (n_cand == 0) continue;
                first_wave = false;
                // ---------------- phase 3: score, one document per 8-lane group, two documents in flight
                for (uint32_t dbase = warp * 4; dbase < n_docs; dbase += 2 * GROUPS) {  // warp-uniform trip count
                    const uint32_t d = dbase + (lane >> 3), d1 = d + GROUPS;
                    const uint64_t pa = d < n_docs ? docs[d] : 0ull;  // nnz 0 -> no loads, score unused
                    const uint64_t pb = d1 < n_docs ? docs[d1] : 0ull;
                    float sa = score_doc(a.ix.fwd, pa, lane8, query);
                    float sb = score_doc(a.ix.fwd, pb, lane8, query);
                    sa = group_reduce(sa);
                    sb = group_reduce(sb);
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
                            heap_offer(have, sc_i, (uint32_t)(pi >> 16), heap_s, heap_k, k, lane, heap_n, theta, wkey,
                                       widx);
                        }
                    }
                    if (lane == 0) s_full = heap_n == k, s_theta = theta;
                }
                __syncthreads();
            }
        }
        // ---------------- results: best first (rank sort by warp 0), padded
        if (warp == 0) {
            heap_write_sorted(heap_s, heap_k, heap_n, k, lane, a.sc.out_keys + (uint64_t)q * k,
                              a.out_scores + (uint64_t)q * k);
            if (lane == 0) a.out_counts[q] = heap_n;
        }
        __syncthreads();
        if (nt > 0) query.template unstage<T>(a.b, qo, qn, tid);
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

}  // namespace sgpu
