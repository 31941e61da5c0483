// k_search: Loop B of the Seismic query path — block traversal with the summary skip test, forward-index
// scoring and the bounded top-k (reference src/posting_list.rs:115-215, src/utils.rs:12-66,
// src/inverted_index.rs:180-234).  Persistent kernel; CTAs fetch query ids from an atomic counter.
//
// The kernel is a template over the shared-memory representation of the query:
//   ByteQuery   byte index per vocabulary entry + value array (dim + 1 KB): the default — two dependent loads per
//               component, the fewest instructions (4 CTAs x 256 threads per SM)
//   RankQuery   bitmap over the vocabulary + per-word rank + value array (dim/8 + dim/32 + 1 KB = 5.8 KB at
//               dim 30522): a forward-index component first tests one bit (96 % of the components of a document
//               are not in the query and stop there); hits fetch their value through popcount ranking.  The only
//               table that fits a 200 k vocabulary (Rec32 / SeismicIndexLV); 1.2-1.4x the instructions of ByteQuery
//   HashQuery   4096-slot perfect hash (tag + value, 24 KB), multiplier found by k_terms
//   DenseQuery  dense f32 vector (dim x 4 B), one 1024-thread CTA per SM; handles ANY query and is the
//               fallback for queries with more than 255 distinct components.
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
// re-enter, KHeap::push is strict — src/utils.rs:36-39), which is what the heaps' offer() checks.
//
// Arithmetic of a document score (pinned by oracle/oracle_search.cpp ORDER_LANES8): 8 lanes per document, lane
// g accumulates the 32-byte chunks g, g+8, ... in ascending component order with separate f32 mul and add
// (no FMA), then a butterfly reduction (xor 4, 2, 1).  Components that are not in the query contribute
// q = +0.0, so skipping them (RankQuery) leaves every partial sum bit-identical.
#pragma once
#include "types.cuh"

namespace sgpu {

constexpr int HQ_THREADS = 128;
constexpr int HQ_D = 2;    // documents in flight per 8-lane group
constexpr int HQ_LOG2_SLOTS = 12;
constexpr int HQ_SLOTS = 1 << HQ_LOG2_SLOTS;
constexpr int HQ_MAX_NNZ = 128;  // HashQuery: queries with more components take the dense kernel
constexpr int HQ_TRIES = 64;
constexpr int DENSE_THREADS = 1024;
#ifndef SGPU_DOT2X
#define SGPU_DOT2X 1
#endif
#ifndef SGPU_SEL_P
#define SGPU_SEL_P 2  // block positions per thread and selection pass of k_search
#endif
#ifndef SGPU_SKIP_MISS
#define SGPU_SKIP_MISS 1
#endif

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
    uint32_t n_knn;            // Knn::refine: neighbours per result document (0 = off; <= ix.knn_dim)
    uint32_t wave_docs;        // soft cap of documents per wave
    uint32_t first_wave_docs;  // soft cap for the first wave of a query (heap still empty)
    uint32_t buf_docs;         // capacity of the wave buffers (>= wave caps); larger blocks are split
    uint32_t cand_cap;         // capacity of the per-wave candidate-block arrays (<= threads per CTA)
    uint32_t qd_words;         // 32-bit words of the query table (meaning depends on the query type)
    uint32_t bucket;           // 1: score the documents of a wave longest first (uniform rounds per warp)
    float value_scale;         // DotVByte: value = code * value_scale
    float* out_scores;         // [nq*k] (chunk-relative)
    uint32_t* out_counts;      // [nq]
};

#ifndef SGPU_LD_MODE
#define SGPU_LD_MODE 2  // A/B builds (same box, 1 M docs, ms per 10 k queries): 0 = ld.global.nc.L1::no_allocate 4.99,
                        // 1 = ld.global.cg (L2 only) 4.98, 2 = ld.global.nc (default L1 policy) 4.92
#endif
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
#if SGPU_LD_MODE == 1
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
#elif SGPU_LD_MODE == 2
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
#endif
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// One 32-byte chunk of the u16 / f16 layout (8 components + 8 values) with ONE 256-bit load (ld.global.nc.v8, sm_100):
// 4.59 vs 4.89 ms per 10 k queries against two 128-bit loads of the same sector (profiles/r2_variants_ld256.log) — half
// the load instructions and no second lookup of every sector in L1.
// (SGPU_LD256 = 0, the previous layout: chunks 4..7 of every round of 8 stored [values | components] — `sw` = bit 2 of
// the chunk index — so that the 8 lanes of a group reading the same half of their chunks from SHARED memory in the
// TMA-staged variant fall into 8 distinct bank quads.)
__device__ __forceinline__ void ld_chunk(const uint4* p, uint32_t sw, uint4& c, uint4& v) {
#if SGPU_LD256
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w), "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
#else
    c = ld_stream(p + sw);
    v = ld_stream(p + (sw ^ 1u));
#endif
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// (One cp.async.bulk.prefetch.L2 per record instead of one prefetch.global.L2 per 128-byte line was measured: the
// instruction is warp-uniform (UBLKPF), so the lanes' records are issued one after the other — 5.05 vs 4.88 ms.)

// ---- TMA (bulk async copy, 1-D) + mbarrier: the staging ring of the TMA variant of k_search ----------------------
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {  // one arrival + `bytes` expected
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nMBW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra MBD;\nbra MBW;\nMBD:\n}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
constexpr int TMA_ROUND_BYTES = 256;                  // one round of one document: 8 chunks of 32 bytes
constexpr int TMA_WARP_BYTES = 8 * TMA_ROUND_BYTES;   // per warp: one round of its 4 x 2 documents
__device__ __forceinline__ float h_lo(uint32_t vw) { return __low2float(*reinterpret_cast<const __half2*>(&vw)); }
__device__ __forceinline__ float h_hi(uint32_t vw) { return __high2float(*reinterpret_cast<const __half2*>(&vw)); }

// ---- query representations in shared memory -------------------------------------------------------------
// Interface: bytes(a) shared memory, init (once per CTA), stage/unstage (per query),
// mac2(acc, cw, vw): acc += q[c_lo]*v_lo ; acc += q[c_hi]*v_hi for the two (u16 component, f16 value) pairs
// packed in cw / vw, in that order, mul then add.
struct DenseQuery {
    static constexpr bool HAS_DOT8 = false;
    float* qd;
    static __device__ __forceinline__ size_t bytes(const SearchArgs& a) { return (size_t)a.qd_words * 4; }
    __device__ __forceinline__ float mac2(float acc, uint32_t cw, uint32_t vw) const {
        acc = __fadd_rn(acc, __fmul_rn(qd[cw & 0xffffu], h_lo(vw)));
        return __fadd_rn(acc, __fmul_rn(qd[cw >> 16], h_hi(vw)));
    }
    __device__ __forceinline__ float mac_f(float acc, uint32_t c, float val) const {
        return __fadd_rn(acc, __fmul_rn(qd[c], val));
    }
    template <int T>
    __device__ __forceinline__ void init(unsigned char* base, const SearchArgs& a, uint32_t tid) {
        qd = reinterpret_cast<float*>(base);
        for (uint32_t i = tid; i < a.qd_words; i += T) qd[i] = 0.f;
    }
    template <int T>
    __device__ __forceinline__ void stage(const Batch& b, const Scratch&, uint32_t, uint64_t qo, uint32_t qn,
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

struct ByteQuery {  // qidx[c] = 1 + position of c in the query (0 = absent); vals[0] = +0.0
    static constexpr bool HAS_DOT8 = true;
    uint8_t* qidx;
    float* vals;
    uint32_t qidx_s, vals_s;  // the same two arrays as 32-bit shared-memory addresses
    // One chunk (8 components): the 8 index loads are issued back to back, then the 8 value loads, then the
    // arithmetic — a warp issues in order, so interleaving "load, test, load, multiply" per component (what the
    // compiler emits for mac2 under the 64-register budget) exposes the shared-memory latency 16 times per chunk.
    // Components that are not in the query read vals[0] = +0.0 and add q*v = +0.0: acc is unchanged bit for bit.
    __device__ __forceinline__ float dot8(float acc, const uint4 c, const uint4 v) const {
        uint32_t i0, i1, i2, i3, i4, i5, i6, i7;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i0) : "r"(qidx_s + (c.x & 0xffffu)));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i1) : "r"(qidx_s + (c.x >> 16)));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i2) : "r"(qidx_s + (c.y & 0xffffu)));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i3) : "r"(qidx_s + (c.y >> 16)));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i4) : "r"(qidx_s + (c.z & 0xffffu)));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i5) : "r"(qidx_s + (c.z >> 16)));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i6) : "r"(qidx_s + (c.w & 0xffffu)));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(i7) : "r"(qidx_s + (c.w >> 16)));
        float q0, q1, q2, q3, q4, q5, q6, q7;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q0) : "r"(vals_s + i0 * 4));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q1) : "r"(vals_s + i1 * 4));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q2) : "r"(vals_s + i2 * 4));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q3) : "r"(vals_s + i3 * 4));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q4) : "r"(vals_s + i4 * 4));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q5) : "r"(vals_s + i5 * 4));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q6) : "r"(vals_s + i6 * 4));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q7) : "r"(vals_s + i7 * 4));
        acc = __fadd_rn(acc, __fmul_rn(q0, h_lo(v.x)));
        acc = __fadd_rn(acc, __fmul_rn(q1, h_hi(v.x)));
        acc = __fadd_rn(acc, __fmul_rn(q2, h_lo(v.y)));
        acc = __fadd_rn(acc, __fmul_rn(q3, h_hi(v.y)));
        acc = __fadd_rn(acc, __fmul_rn(q4, h_lo(v.z)));
        acc = __fadd_rn(acc, __fmul_rn(q5, h_hi(v.z)));
        acc = __fadd_rn(acc, __fmul_rn(q6, h_lo(v.w)));
        return __fadd_rn(acc, __fmul_rn(q7, h_hi(v.w)));
    }
    // One chunk of each of two documents in a single basic block: 16 independent index loads, then 16 value loads,
    // then the two (ordered) multiply-add chains — the shared-memory latency is paid twice per call, not per chunk.
    // (The index loads are 67 % of the kernel's shared-memory wavefronts: 32 random bytes hit 2.9 banks-deep on average.
    // Issuing every one of them twice — +61 % wavefronts, same results — costs +23 % kernel time, i.e. removing ALL of
    // their bank conflicts would buy at most ~15 %: profiles/r2_exp_double_idx.log.)
    __device__ __forceinline__ void dot2x(float& a0, float& a1, const uint4 (&c)[2], const uint4 (&v)[2]) const {
        uint32_t idx[16];
        float q[16];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t cw[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(idx[8 * j + 2 * e]) : "r"(qidx_s + (cw[e] & 0xffffu)));
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(idx[8 * j + 2 * e + 1]) : "r"(qidx_s + (cw[e] >> 16)));
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q[i]) : "r"(vals_s + idx[i] * 4));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t v0 = e == 0 ? v[0].x : e == 1 ? v[0].y : e == 2 ? v[0].z : v[0].w;
            const uint32_t v1 = e == 0 ? v[1].x : e == 1 ? v[1].y : e == 2 ? v[1].z : v[1].w;
            a0 = __fadd_rn(a0, __fmul_rn(q[2 * e], h_lo(v0)));
            a1 = __fadd_rn(a1, __fmul_rn(q[8 + 2 * e], h_lo(v1)));
            a0 = __fadd_rn(a0, __fmul_rn(q[2 * e + 1], h_hi(v0)));
            a1 = __fadd_rn(a1, __fmul_rn(q[8 + 2 * e + 1], h_hi(v1)));
        }
    }
    // Eight decoded (component, f32 value) pairs of one lane (DotVByte): all index loads, then all value loads, then the
    // ordered multiply-add chain.  Misses read vals[0] = +0.0 and add +-0: bit-identical to skipping them (mac_f).
    __device__ __forceinline__ float dot8f(float acc, const uint32_t (&c)[8], const float (&v)[8]) const {
        uint32_t idx[8];
        float q[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(idx[e]) : "r"(qidx_s + c[e]));
#pragma unroll
        for (int e = 0; e < 8; ++e) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q[e]) : "r"(vals_s + idx[e] * 4));
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = __fadd_rn(acc, __fmul_rn(q[e], v[e]));
        return acc;
    }
    static __device__ __forceinline__ size_t bytes(const SearchArgs& a) { return 1024 + (size_t)a.qd_words * 4; }
    __device__ __forceinline__ float mac2(float acc, uint32_t cw, uint32_t vw) const {
#if SGPU_SKIP_MISS
        // components that are not in the query contribute +0.0: skipping them leaves acc bit-identical
        const uint32_t i0 = qidx[cw & 0xffffu], i1 = qidx[cw >> 16];
        if (i0) acc = __fadd_rn(acc, __fmul_rn(vals[i0], h_lo(vw)));
        if (i1) acc = __fadd_rn(acc, __fmul_rn(vals[i1], h_hi(vw)));
        return acc;
#else
        acc = __fadd_rn(acc, __fmul_rn(vals[qidx[cw & 0xffffu]], h_lo(vw)));
        return __fadd_rn(acc, __fmul_rn(vals[qidx[cw >> 16]], h_hi(vw)));
#endif
    }
    __device__ __forceinline__ float mac_f(float acc, uint32_t c, float val) const {
        const uint32_t i = qidx[c];
        return i ? __fadd_rn(acc, __fmul_rn(vals[i], val)) : acc;
    }
    template <int T>
    __device__ __forceinline__ void init(unsigned char* base, const SearchArgs& a, uint32_t tid) {
        vals = reinterpret_cast<float*>(base);
        qidx = base + 1024;
        vals_s = (uint32_t)__cvta_generic_to_shared(vals);
        qidx_s = (uint32_t)__cvta_generic_to_shared(qidx);
        for (uint32_t i = tid; i < 256; i += T) vals[i] = 0.f;
        uint32_t* w = reinterpret_cast<uint32_t*>(qidx);
        for (uint32_t i = tid; i < a.qd_words; i += T) w[i] = 0u;
    }
    template <int T>
    __device__ __forceinline__ void stage(const Batch& b, const Scratch&, uint32_t, uint64_t qo, uint32_t qn,
                                          uint32_t tid) {
        for (uint32_t i = tid; i < qn; i += T) {
            const uint32_t c = b.q_comps[qo + i];
            if (i + 1 == qn || b.q_comps[qo + i + 1] != c) {  // last duplicate wins
                qidx[c] = (uint8_t)(i + 1);
                vals[i + 1] = b.q_vals[qo + i];
            }
        }
    }
    template <int T>
    __device__ __forceinline__ void unstage(const Batch& b, uint64_t qo, uint32_t qn, uint32_t tid) {
        for (uint32_t i = tid; i < qn; i += T) qidx[b.q_comps[qo + i]] = 0;
    }
};

struct HashQuery {
    static constexpr bool HAS_DOT8 = false;
    uint16_t* tags;
    float* vals;
    uint32_t mult;
    static __device__ __forceinline__ size_t bytes(const SearchArgs&) { return (size_t)HQ_SLOTS * 6; }
    __device__ __forceinline__ float get(uint32_t c) const {
        const uint32_t s = hq_slot(c, mult);
        const float v = vals[s];
        return tags[s] == (uint16_t)c ? v : 0.f;
    }
    __device__ __forceinline__ float mac2(float acc, uint32_t cw, uint32_t vw) const {
        acc = __fadd_rn(acc, __fmul_rn(get(cw & 0xffffu), h_lo(vw)));
        return __fadd_rn(acc, __fmul_rn(get(cw >> 16), h_hi(vw)));
    }
    __device__ __forceinline__ float mac_f(float acc, uint32_t c, float val) const {
        return __fadd_rn(acc, __fmul_rn(get(c), val));
    }
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

// bm: one bit per vocabulary entry; pre[w]: rank (among the query's distinct components) of the first component
// of bitmap word w, valid for words that have a bit set; vals[rank].  <= 255 distinct components.
struct RankQuery {
    static constexpr bool HAS_DOT8 = false;
    uint32_t* bm;
    uint8_t* pre;
    float* vals;
    // qd_words = number of bitmap words (dim/32 rounded up to a multiple of 4)
    static __device__ __forceinline__ size_t bytes(const SearchArgs& a) { return 1024 + (size_t)a.qd_words * 5; }
    __device__ __forceinline__ float mac(float acc, uint32_t c, uint32_t vw, bool hi) const {
        const uint32_t w = bm[c >> 5];
        if ((w >> (c & 31)) & 1u) {  // ~4 % of a document's components
            const uint32_t r = pre[c >> 5] + __popc(w & ((1u << (c & 31)) - 1u));
            acc = __fadd_rn(acc, __fmul_rn(vals[r], hi ? h_hi(vw) : h_lo(vw)));
        }
        return acc;
    }
    __device__ __forceinline__ float mac2(float acc, uint32_t cw, uint32_t vw) const {
        acc = mac(acc, cw & 0xffffu, vw, false);
        return mac(acc, cw >> 16, vw, true);
    }
    // (Fetching the 8 bitmap words of a chunk back to back before the hit tests, like ByteQuery::dot8, was measured on
    // the large-vocabulary benchmark: 12.14 vs 11.97 ms per 10 k queries — no gain, dropped.)
    __device__ __forceinline__ float mac_f(float acc, uint32_t c, float val) const {
        const uint32_t w = bm[c >> 5];
        if ((w >> (c & 31)) & 1u) {
            const uint32_t r = pre[c >> 5] + __popc(w & ((1u << (c & 31)) - 1u));
            acc = __fadd_rn(acc, __fmul_rn(vals[r], val));
        }
        return acc;
    }
    template <int T>
    __device__ __forceinline__ void init(unsigned char* base, const SearchArgs& a, uint32_t tid) {
        vals = reinterpret_cast<float*>(base);
        bm = reinterpret_cast<uint32_t*>(base + 1024);
        pre = base + 1024 + (size_t)a.qd_words * 4;
        for (uint32_t i = tid; i < a.qd_words; i += T) bm[i] = 0u;
    }
    template <int T>
    __device__ __forceinline__ void stage(const Batch& b, const Scratch&, uint32_t, uint64_t qo, uint32_t qn,
                                          uint32_t tid) {
        const uint32_t* qc = b.q_comps + qo;
        for (uint32_t i = tid; i < qn; i += T) {
            const uint32_t c = qc[i];
            if (i + 1 == qn || qc[i + 1] != c) {  // last occurrence of c carries the value (last duplicate wins)
                uint32_t rank = 0, first_in_word = 1;
                for (uint32_t j = 0; j < i; ++j) {
                    const bool run_end = qc[j] != qc[j + 1];
                    rank += run_end;
                    if (run_end && (qc[j] >> 5) == (c >> 5)) first_in_word = 0;
                }
                atomicOr(&bm[c >> 5], 1u << (c & 31));
                vals[rank] = b.q_vals[qo + i];
                if (first_in_word) pre[c >> 5] = (uint8_t)rank;
            }
        }
    }
    template <int T>
    __device__ __forceinline__ void unstage(const Batch& b, uint64_t qo, uint32_t qn, uint32_t tid) {
        for (uint32_t i = tid; i < qn; i += T) bm[b.q_comps[qo + i] >> 5] = 0u;
    }
};

// The query as it arrives: sorted components (duplicates included) + values, looked up by binary search.  Slow per
// component (log2(nnz) dependent shared-memory loads) but it takes a query of ANY length on ANY record layout — the
// path for queries with more than 255 components on the layouts that have no dense-query kernel (u32 components,
// DotVByte, bf16 / f32 / fixed-point values); e.g. the document-as-query searches of Knn::new
// (src/inverted_index.rs:448-500).  A duplicated component resolves to its LAST occurrence, like the dense scatter.
struct SortedQuery {
    static constexpr bool HAS_DOT8 = false;
    uint32_t* qc;
    float* qv;
    uint32_t n, steps;
    // qd_words = capacity in components
    static __device__ __forceinline__ size_t bytes(const SearchArgs& a) { return (size_t)a.qd_words * 8 + 16; }
    __device__ __forceinline__ float mac_f(float acc, uint32_t c, float val) const {
        uint32_t lo = 0, len = n;  // upper_bound(qc, c): first position with qc > c
        for (uint32_t s = 0; s < steps; ++s) {
            const uint32_t half = len >> 1;
            const bool right = len > 0 && qc[lo + half] <= c;
            lo = right ? lo + half + 1 : lo;
            len = right ? len - half - 1 : half;
        }
        if (lo > 0 && qc[lo - 1] == c) acc = __fadd_rn(acc, __fmul_rn(qv[lo - 1], val));
        return acc;
    }
    __device__ __forceinline__ float mac(float acc, uint32_t c, uint32_t vw, bool hi) const {
        return mac_f(acc, c, hi ? h_hi(vw) : h_lo(vw));
    }
    __device__ __forceinline__ float mac2(float acc, uint32_t cw, uint32_t vw) const {
        acc = mac_f(acc, cw & 0xffffu, h_lo(vw));
        return mac_f(acc, cw >> 16, h_hi(vw));
    }
    template <int T>
    __device__ __forceinline__ void init(unsigned char* base, const SearchArgs& a, uint32_t) {
        qc = reinterpret_cast<uint32_t*>(base);
        qv = reinterpret_cast<float*>(base + (size_t)a.qd_words * 4);
        n = 0, steps = 0;
    }
    template <int T>
    __device__ __forceinline__ void stage(const Batch& b, const Scratch&, uint32_t, uint64_t qo, uint32_t qn,
                                          uint32_t tid) {
        for (uint32_t i = tid; i < qn; i += T) qc[i] = b.q_comps[qo + i], qv[i] = b.q_vals[qo + i];
        n = qn;
        steps = 32 - __clz(qn);  // halving a range of n needs floor(log2(n)) + 1 steps to reach length 0
    }
    template <int T>
    __device__ __forceinline__ void unstage(const Batch&, uint64_t, uint32_t, uint32_t) {}
};

// ---- forward-index record layouts ---------------------------------------------------------------------------
// A record is a sequence of chunks of 8 (component, value) pairs; the posting's start field counts UNIT-byte units.
struct Rec16 {  // u16 components: chunk = [8 x u16 | 8 x f16] = 32 bytes = 2 x uint4, unit 32 bytes
    static constexpr bool PLAIN_F16 = true;
    static constexpr int CHUNK_BYTES = 32;
    static constexpr int UNIT_BYTES = 32;  // unit of the posting's start field
    struct Chunk { uint4 c, v; };
    static __device__ __forceinline__ void load(const char* p, Chunk& k, uint32_t m) {
        ld_chunk(reinterpret_cast<const uint4*>(p), (m >> 2) & 1u, k.c, k.v);
    }
    template <class Q>
    static __device__ __forceinline__ float dot(float acc, const Chunk& k, const Q& q, float) {
        if constexpr (Q::HAS_DOT8) {
            return q.dot8(acc, k.c, k.v);
        } else {
            acc = q.mac2(acc, k.c.x, k.v.x);
            acc = q.mac2(acc, k.c.y, k.v.y);
            acc = q.mac2(acc, k.c.z, k.v.z);
            return q.mac2(acc, k.c.w, k.v.w);
        }
    }
};
struct Rec32 {  // u32 components (large vocabulary): chunk = [8 x u32 | 8 x f16] = 48 bytes = 3 x uint4, unit 16 bytes
    static constexpr bool PLAIN_F16 = false;
    static constexpr int CHUNK_BYTES = 48;
    static constexpr int UNIT_BYTES = 16;
    struct Chunk { uint4 c0, c1, v; };
    static __device__ __forceinline__ void load(const char* p, Chunk& k, uint32_t) {
        const uint4* p4 = reinterpret_cast<const uint4*>(p);
        k.c0 = ld_stream(p4);
        k.c1 = ld_stream(p4 + 1);
        k.v = ld_stream(p4 + 2);
    }
    template <class Q>
    static __device__ __forceinline__ float dot(float acc, const Chunk& k, const Q& q, float) {
        acc = q.mac(acc, k.c0.x, k.v.x, false);
        acc = q.mac(acc, k.c0.y, k.v.x, true);
        acc = q.mac(acc, k.c0.z, k.v.y, false);
        acc = q.mac(acc, k.c0.w, k.v.y, true);
        acc = q.mac(acc, k.c1.x, k.v.z, false);
        acc = q.mac(acc, k.c1.y, k.v.z, true);
        acc = q.mac(acc, k.c1.z, k.v.w, false);
        return q.mac(acc, k.c1.w, k.v.w, true);
    }
};

// ---- the other plain value encodings of the reference, u16 components.  value -> f32 exactly as the
// oracle's decode(): bf16 = bits << 16; f32 as is; fixedu8 / fixedu16 = (float)code * scale (own definition, the
// reference's FixedU8Q/FixedU16Q scaling lives in vectorium).  These go through the generic per-component
// q.mac_f path (they are not the benchmark encodings).
__device__ __forceinline__ float u_to_f32(uint32_t code) {  // exact for code < 2^23, no conversion pipe
    return __uint_as_float(0x4b000000u | code) - 8388608.f;
}
template <int KIND>  // 1 bf16, 4 fixedu16: same 32-byte chunks as Rec16
struct Rec16V2 {
    static constexpr bool PLAIN_F16 = false;
    static constexpr int CHUNK_BYTES = 32;
    static constexpr int UNIT_BYTES = 32;
    struct Chunk { uint4 c, v; };
    static __device__ __forceinline__ void load(const char* p, Chunk& k, uint32_t) {
        ld_chunk(reinterpret_cast<const uint4*>(p), 0u, k.c, k.v);
    }
    static __device__ __forceinline__ float val(uint32_t vw, bool hi, float scale) {
        if constexpr (KIND == 1) return __uint_as_float(hi ? (vw & 0xffff0000u) : (vw << 16));
        else return __fmul_rn(u_to_f32(hi ? (vw >> 16) : (vw & 0xffffu)), scale);
    }
    template <class Q>
    static __device__ __forceinline__ float dot(float acc, const Chunk& k, const Q& q, float scale) {
        const uint32_t cw[4] = {k.c.x, k.c.y, k.c.z, k.c.w}, vw[4] = {k.v.x, k.v.y, k.v.z, k.v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc = q.mac_f(acc, cw[j] & 0xffffu, val(vw[j], false, scale));
            acc = q.mac_f(acc, cw[j] >> 16, val(vw[j], true, scale));
        }
        return acc;
    }
};
struct Rec16F32 {  // chunk = [8 x u16 | 8 x f32] = 48 bytes, unit 16 bytes
    static constexpr bool PLAIN_F16 = false;
    static constexpr int CHUNK_BYTES = 48;
    static constexpr int UNIT_BYTES = 16;
    struct Chunk { uint4 c, v0, v1; };
    static __device__ __forceinline__ void load(const char* p, Chunk& k, uint32_t) {
        const uint4* p4 = reinterpret_cast<const uint4*>(p);
        k.c = ld_stream(p4);
        k.v0 = ld_stream(p4 + 1);
        k.v1 = ld_stream(p4 + 2);
    }
    template <class Q>
    static __device__ __forceinline__ float dot(float acc, const Chunk& k, const Q& q, float) {
        acc = q.mac_f(acc, k.c.x & 0xffffu, __uint_as_float(k.v0.x));
        acc = q.mac_f(acc, k.c.x >> 16, __uint_as_float(k.v0.y));
        acc = q.mac_f(acc, k.c.y & 0xffffu, __uint_as_float(k.v0.z));
        acc = q.mac_f(acc, k.c.y >> 16, __uint_as_float(k.v0.w));
        acc = q.mac_f(acc, k.c.z & 0xffffu, __uint_as_float(k.v1.x));
        acc = q.mac_f(acc, k.c.z >> 16, __uint_as_float(k.v1.y));
        acc = q.mac_f(acc, k.c.w & 0xffffu, __uint_as_float(k.v1.z));
        return q.mac_f(acc, k.c.w >> 16, __uint_as_float(k.v1.w));
    }
};
struct Rec16U8 {  // chunk = [8 x u16 | 8 x u8] = 24 bytes, unit 8 bytes (8-byte loads)
    static constexpr bool PLAIN_F16 = false;
    static constexpr int CHUNK_BYTES = 24;
    static constexpr int UNIT_BYTES = 8;
    struct Chunk { uint2 c0, c1, v; };
    static __device__ __forceinline__ void load(const char* p, Chunk& k, uint32_t) {
        const uint2* p2 = reinterpret_cast<const uint2*>(p);
        k.c0 = __ldg(p2);
        k.c1 = __ldg(p2 + 1);
        k.v = __ldg(p2 + 2);
    }
    template <class Q>
    static __device__ __forceinline__ float dot(float acc, const Chunk& k, const Q& q, float scale) {
        const uint32_t cw[4] = {k.c0.x, k.c0.y, k.c1.x, k.c1.y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t vw = j < 2 ? k.v.x : k.v.y;
            acc = q.mac_f(acc, cw[j] & 0xffffu, __fmul_rn(u_to_f32((vw >> (16 * (j & 1))) & 0xffu), scale));
            acc = q.mac_f(acc, cw[j] >> 16, __fmul_rn(u_to_f32((vw >> (16 * (j & 1) + 8)) & 0xffu), scale));
        }
        return acc;
    }
};

// u32 components x the other plain value encodings ("both u16 and u32"): chunk = [8 x u32 | 8 values].
// KIND = SGPU_VAL_*: 1 bf16, 2 f32, 3 fixedu8, 4 fixedu16.  Generic per-component q.mac_f path.
template <int KIND>
struct Rec32V {
    static constexpr bool PLAIN_F16 = false;
    static constexpr int VB = KIND == 2 ? 4 : (KIND == 3 ? 1 : 2);
    static constexpr int CHUNK_BYTES = 32 + 8 * VB;  // 48, 64 or 40
    static constexpr int UNIT_BYTES = CHUNK_BYTES % 32 == 0 ? 32 : (CHUNK_BYTES % 16 == 0 ? 16 : 8);
    static constexpr int WORDS = CHUNK_BYTES / 4;
    struct Chunk { uint32_t w[WORDS]; };
    static __device__ __forceinline__ void load(const char* p, Chunk& k, uint32_t) {
        if constexpr (CHUNK_BYTES % 16 == 0) {
            const uint4* p4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
            for (int i = 0; i < WORDS / 4; ++i) {
                const uint4 t = ld_stream(p4 + i);
                k.w[4 * i] = t.x, k.w[4 * i + 1] = t.y, k.w[4 * i + 2] = t.z, k.w[4 * i + 3] = t.w;
            }
        } else {
            const uint2* p2 = reinterpret_cast<const uint2*>(p);
#pragma unroll
            for (int i = 0; i < WORDS / 2; ++i) {
                const uint2 t = __ldg(p2 + i);
                k.w[2 * i] = t.x, k.w[2 * i + 1] = t.y;
            }
        }
    }
    static __device__ __forceinline__ float val(const Chunk& k, int e, float scale) {
        if constexpr (KIND == 2) return __uint_as_float(k.w[8 + e]);
        else if constexpr (KIND == 3) return __fmul_rn(u_to_f32((k.w[8 + (e >> 2)] >> (8 * (e & 3))) & 0xffu), scale);
        else {
            const uint32_t vw = k.w[8 + (e >> 1)];
            if constexpr (KIND == 1) return __uint_as_float((e & 1) ? (vw & 0xffff0000u) : (vw << 16));
            else return __fmul_rn(u_to_f32((e & 1) ? (vw >> 16) : (vw & 0xffffu)), scale);
        }
    }
    template <class Q>
    static __device__ __forceinline__ float dot(float acc, const Chunk& k, const Q& q, float scale) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = q.mac_f(acc, k.w[e], val(k, e, scale));
        return acc;
    }
};

// Score one document record (nch chunks at `rec`) with an 8-lane group; lane8 handles chunks
// lane8, lane8+8, ...  The caller reduces the 8 partial sums with group_reduce.
template <class R, class Q>
__device__ __forceinline__ float score_rec(const char* __restrict__ rec, uint32_t nch, uint32_t lane8, const Q& q,
                                           float scale) {
    float acc = 0.f;
    for (uint32_t m = lane8; m < nch; m += 8) {
        typename R::Chunk k;
        R::load(rec + (size_t)R::CHUNK_BYTES * m, k, m);
        acc = R::dot(acc, k, q, scale);
    }
    return acc;
}
__device__ __forceinline__ float group_reduce(float s) {
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
    return __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
}

// Score D documents with one 8-lane group, round by round: in round r lane8 handles chunk lane8 + 8r of every
// document; all loads of a round are issued before the first use.  `rounds` must be warp-uniform.
template <int D, class R, class Q>
__device__ __forceinline__ void score_docs(const uint4* __restrict__ fwd, const uint64_t (&post)[D], uint32_t lane8,
                                           uint32_t rounds, const Q& q, float scale, float (&acc)[D]) {
    const char* rec[D];
    uint32_t nch[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        rec[j] = reinterpret_cast<const char*>(fwd) + (post[j] >> 16) * R::UNIT_BYTES + R::CHUNK_BYTES * lane8;
        nch[j] = ((uint32_t)(post[j] & 0xffffu) + 7) >> 3;
        acc[j] = 0.f;
    }
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t m = lane8 + 8 * r;
        if constexpr (SGPU_DOT2X && D == 2 && Q::HAS_DOT8 && R::CHUNK_BYTES == 32 && R::PLAIN_F16) {
            // both documents in one basic block; a chunk past the end of a record is (0, +0.0) x 8: adds q * 0 = +-0
            uint4 c[2], v[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                c[j] = make_uint4(0, 0, 0, 0), v[j] = make_uint4(0, 0, 0, 0);
                if (m < nch[j]) ld_chunk(reinterpret_cast<const uint4*>(rec[j] + (size_t)256 * r), (lane8 >> 2) & 1u, c[j], v[j]);
            }
            q.dot2x(acc[0], acc[1], c, v);
        } else {
            typename R::Chunk k[D];
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (m < nch[j]) R::load(rec[j] + (size_t)R::CHUNK_BYTES * 8 * r, k[j], m);
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (m < nch[j]) acc[j] = R::dot(acc[j], k[j], q, scale);
        }
    }
}

// ---- DotVByte records: gap-coded u16 components (1 or 2 bytes per gap) + u8 values -----------
// Format: csrc/host/build.cpp (convert_dotvbyte).  The posting's start field counts 16-byte units of the byte stream.
// Per record: a 16-byte directory entry per super-round of 64 chunks {u64 mask of the WIDE chunks, u32 wide chunks
// before}, 16 bytes per chunk [8 low bytes of its gaps | 8 codes], 8 bytes per wide chunk [the high bytes of its gaps];
// the gaps are ONE chain over the record.  lane8 decodes chunk lane8 of every round from TWO loads — its 16 fixed bytes
// and, if its mask bit is set, its 8 high bytes (rank by popcount) — which is what the uncompressed layout issues too:
// with the previous layout (control bit per component, exception groups behind a u16 offset table: 5 loads per chunk)
// the decode arithmetic measured as free and the three extra loads as the whole 24 % gap to f16
// (profiles/r2_dotvbyte_loads_experiment.log).  Four byte-permutes pair low and high bytes into u16 gaps, a packed
// prefix sum ((a, b) * 0x10001 = (a, a + b)) chains them inside the chunk, an 8-lane shuffle scan of the chunk totals
// plus the carry of the earlier rounds gives the chunk's first component — and the result is exactly a chunk of the plain
// layout (4 words of two u16 components, 4 words of two f16 values), fed to the same lookup / multiply-add code.  A code
// byte c is read as the f16 SUBNORMAL c * 2^-24 (exact), so the f16 -> f32 conversion of the plain layout doubles as the
// integer -> float conversion; the factor scale * 2^24 is applied once per document (oracle: doc_score_vbyte).
struct RecVB {
    static constexpr bool VBYTE = true;
    static constexpr bool PLAIN_F16 = false;
    static constexpr int UNIT_BYTES = 16;
};
template <class R>
struct is_vbyte { static constexpr bool value = false; };
template <>
struct is_vbyte<RecVB> { static constexpr bool value = true; };

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {  // selector nibbles must be 0..7
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

template <int D, class Q>
__device__ __forceinline__ void score_docs_vb(const uint8_t* __restrict__ stream, const uint64_t (&post)[D],
                                              uint32_t lane8, uint32_t rounds, const Q& q, float (&acc)[D],
                                              uint32_t& bytes) {
    // all addresses are (stream as 16-byte units) + 32-bit unit index: one IMAD.WIDE each
    const uint4* s16 = reinterpret_cast<const uint4*>(stream);
    uint32_t r16[D], nch[D], f16[D], w16[D];  // record start, chunks, first fixed part, wide area (16-byte units)
    uint32_t mlo[D], mhi[D], wbef[D], carry[D];  // current directory entry; components of the earlier rounds summed
#pragma unroll
    for (int j = 0; j < D; ++j) {
        r16[j] = (uint32_t)(post[j] >> 16);
        nch[j] = ((uint32_t)(post[j] & 0xffffu) + 7) >> 3;
        const uint32_t ndir = (nch[j] + 63) >> 6;
        f16[j] = r16[j] + ndir;
        w16[j] = f16[j] + nch[j];
        mlo[j] = mhi[j] = wbef[j] = carry[j] = 0;
        acc[j] = 0.f;
        if (lane8 == 0) bytes += 16 * (ndir + nch[j]);
    }
    for (uint32_t r = 0; r < rounds; ++r) {  // `rounds` and therefore r are warp-uniform
        const uint32_t m = lane8 + 8 * r;
        if ((r & 7) == 0) {  // a new super-round: its directory entry (the 8 lanes of a group read the same 16 bytes)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                uint4 de = make_uint4(0, 0, 0, 0);
                if (8 * r < nch[j]) de = ld_stream(s16 + (r16[j] + (r >> 3)));
                mlo[j] = de.x, mhi[j] = de.y, wbef[j] = de.z;
            }
        }
        const uint32_t b = lane8 + 8 * (r & 7);  // bit of chunk m in the mask of its super-round
        uint4 c[D], v[D], fx[D];
        uint2 hi[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {  // both loads of every chunk of the round are issued before the first use
            uint32_t w, below;
            if ((r & 4) == 0) {  // b < 32 (uniform)
                w = mlo[j];
                below = __popc(mlo[j] & ((1u << b) - 1u));
            } else {
                w = mhi[j];
                below = __popc(mlo[j]) + __popc(mhi[j] & ((1u << (b & 31)) - 1u));
            }
            fx[j] = make_uint4(0, 0, 0, 0);
            hi[j] = make_uint2(0, 0);
            if (m < nch[j]) fx[j] = ld_stream(s16 + (f16[j] + m));
            if ((w >> (b & 31)) & 1u) {  // mask bits exist only for chunks of the record
                hi[j] = __ldg(reinterpret_cast<const uint2*>(s16 + w16[j]) + (wbef[j] + below));
                bytes += 8;
            }
        }
#pragma unroll
        for (int j = 0; j < D; ++j) {  // a chunk past the end of a record is all zero: (last component, +0.0) x 8
            // (low byte, high byte) pairs -> u16 gaps, two per word; packed prefix sum inside the chunk (the gaps of
            // a record sum to less than 2^16, so no carry crosses the halves)
            c[j].x = prmt(fx[j].x, hi[j].x, 0x5140) * 0x10001u;
            c[j].y = (prmt(fx[j].x, hi[j].x, 0x7362) + (c[j].x >> 16)) * 0x10001u;
            c[j].z = (prmt(fx[j].y, hi[j].y, 0x5140) + (c[j].y >> 16)) * 0x10001u;
            c[j].w = (prmt(fx[j].y, hi[j].y, 0x7362) + (c[j].z >> 16)) * 0x10001u;
            // codes -> f16 subnormals (code * 2^-24), two per word
            v[j].x = prmt(fx[j].z, 0u, 0x4140);
            v[j].y = prmt(fx[j].z, 0u, 0x4342);
            v[j].z = prmt(fx[j].w, 0u, 0x4140);
            v[j].w = prmt(fx[j].w, 0u, 0x4342);
        }
        // first component of every chunk = components of the earlier rounds + of the lower lanes of this round: an
        // 8-lane inclusive scan of the chunk totals.  Totals and their sums stay below 2^16, so the two documents of a
        // group are scanned as the halves of one register.
        if constexpr (D == 2) {
            const uint32_t tot = prmt(c[0].w, c[1].w, 0x7632);  // (total of document 0, total of document 1)
            uint32_t incl = tot;
#pragma unroll
            for (int sft = 1; sft < 8; sft <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, sft, 8);
                if (lane8 >= (uint32_t)sft) incl += up;
            }
            const uint32_t base = carry[0] + incl - tot;
            carry[0] += __shfl_sync(0xffffffffu, incl, 7, 8);
            const uint32_t b0 = prmt(base, 0u, 0x1010), b1 = prmt(base, 0u, 0x3232);  // each half copied into both halves
            c[0].x += b0, c[0].y += b0, c[0].z += b0, c[0].w += b0;
            c[1].x += b1, c[1].y += b1, c[1].z += b1, c[1].w += b1;
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const uint32_t tot = c[j].w >> 16;
                uint32_t incl = tot;
#pragma unroll
                for (int sft = 1; sft < 8; sft <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, sft, 8);
                    if (lane8 >= (uint32_t)sft) incl += up;
                }
                const uint32_t base2 = (carry[j] + incl - tot) * 0x10001u;
                carry[j] += __shfl_sync(0xffffffffu, incl, 7, 8);
                c[j].x += base2, c[j].y += base2, c[j].z += base2, c[j].w += base2;
            }
        }
        if constexpr (D == 2 && Q::HAS_DOT8) {
            q.dot2x(acc[0], acc[1], c, v);
        } else if constexpr (Q::HAS_DOT8) {
#pragma unroll
            for (int j = 0; j < D; ++j) acc[j] = q.dot8(acc[j], c[j], v[j]);
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                acc[j] = q.mac2(acc[j], c[j].x, v[j].x);
                acc[j] = q.mac2(acc[j], c[j].y, v[j].y);
                acc[j] = q.mac2(acc[j], c[j].z, v[j].z);
                acc[j] = q.mac2(acc[j], c[j].w, v[j].w);
            }
        }
    }
}

// ---- bounded top-k kept by ONE warp ------------------------------------------------------------------------
__device__ __forceinline__ bool better(float s, uint32_t key, float ws, uint32_t wkey) {
    return s > ws || (s == ws && key < wkey);
}

// KHeap (src/utils.rs:12-66) for k <= 32: lane i holds the i-th best retained item (sorted, registers only).
struct RegHeap {
    float s;        // per lane
    uint32_t key;   // per lane
    uint32_t n, k;  // warp-uniform
    float theta;    // worst retained score (valid when full)
    uint32_t wkey;
    float* xs;      // k-entry exchange buffers in shared memory (merge32)
    uint32_t* xk;
    __device__ __forceinline__ void reset(uint32_t kk, float* bs, uint32_t* bk) {
        n = 0, k = kk, theta = 0.f, wkey = 0, s = 0.f, key = 0, xs = bs, xk = bk;
    }
    __device__ __forceinline__ bool full() const { return n == k; }
    // Push the candidates of the lanes in `m` (each can enter the heap as it stands) in ONE step: the result of
    // pushing a set of items one by one is the k best of (heap U set) whatever the order (total order, strict
    // replacement), so every item's final position is its rank in the union — counted with pipelined broadcasts
    // instead of a dependent insertion per item.  Keys within one call must be distinct (a posting list holds a
    // document once, and one call never spans two lists).
    __device__ __forceinline__ void merge32(uint32_t m, const float sc, const uint32_t ky, uint32_t lane) {
        bool c = (m >> lane) & 1u;
        uint32_t hb = 0;  // retained items better than this lane's candidate
        bool dup = false;
        for (uint32_t i = 0; i < n; ++i) {
            const float hs = __shfl_sync(0xffffffffu, s, i);
            const uint32_t hk = __shfl_sync(0xffffffffu, key, i);
            dup |= hk == ky;
            hb += better(hs, hk, sc, ky);
        }
        c = c && !dup;
        m = __ballot_sync(0xffffffffu, c);
        uint32_t cb_h = 0, cb_c = 0;  // candidates better than this lane's retained item / candidate
        for (uint32_t mm = m; mm; mm &= mm - 1) {
            const int j = __ffs(mm) - 1;
            const float cs = __shfl_sync(0xffffffffu, sc, j);
            const uint32_t ck = __shfl_sync(0xffffffffu, ky, j);
            cb_h += better(cs, ck, s, key);
            cb_c += better(cs, ck, sc, ky);
        }
        const uint32_t rh = lane + cb_h, rc = hb + cb_c;
        if (lane < n && rh < k) xs[rh] = s, xk[rh] = key;
        if (c && rc < k) xs[rc] = sc, xk[rc] = ky;
        __syncwarp();
        n = min(k, n + (uint32_t)__popc(m));
        if (lane < n) s = xs[lane], key = xk[lane];
        __syncwarp();
        if (n == k) {
            theta = __shfl_sync(0xffffffffu, s, k - 1);
            wkey = __shfl_sync(0xffffffffu, key, k - 1);
        }
    }
    // push up to 32 items, one per lane; items already retained (same key) are ignored
    // `distinct`: the keys of one call are known to differ (always true inside one posting list)
    __device__ __forceinline__ void offer(bool have, const float sc, const uint32_t ky, uint32_t lane,
                                          bool distinct = true) {
        {
            const uint32_t m0 = __ballot_sync(0xffffffffu, have && (n < k || better(sc, ky, theta, wkey)));
            if (distinct && __popc(m0) > 2) {
                merge32(m0, sc, ky, lane);
                return;
            }
        }
        for (;;) {
            const bool c = have && (n < k || better(sc, ky, theta, wkey));
            const uint32_t m = __ballot_sync(0xffffffffu, c);
            if (!m) break;
            const int src = __ffs(m) - 1;
            const float bs = __shfl_sync(0xffffffffu, sc, src);
            const uint32_t bk = __shfl_sync(0xffffffffu, ky, src);
            if ((int)lane == src) have = false;
            if (__ballot_sync(0xffffffffu, lane < n && key == bk)) continue;
            const uint32_t pos = __popc(__ballot_sync(0xffffffffu, lane < n && better(s, key, bs, bk)));
            const float ps = __shfl_up_sync(0xffffffffu, s, 1);
            const uint32_t pk = __shfl_up_sync(0xffffffffu, key, 1);
            const uint32_t last = n < k ? n : k - 1;
            if (lane > pos && lane <= last) s = ps, key = pk;
            if (lane == pos) s = bs, key = bk;
            if (n < k) ++n;
            if (n == k) {
                theta = __shfl_sync(0xffffffffu, s, k - 1);
                wkey = __shfl_sync(0xffffffffu, key, k - 1);
            }
        }
    }
    __device__ __forceinline__ void store_keys(uint32_t* dst, uint32_t lane) const {
        if (lane < n) dst[lane] = key;
    }
    __device__ __forceinline__ void write_sorted(uint32_t lane, uint32_t* out_keys, float* out_scores) const {
        for (uint32_t i = lane; i < k; i += 32) {
            const bool ok = i < n;  // only i == lane < 32 can be valid
            out_keys[i] = ok ? key : 0xffffffffu;
            out_scores[i] = ok ? s : -INFINITY;
        }
    }
};

// KHeap for any k (<= 1024): the retained items sorted best-first in shared memory.  A call merges up to 32 candidates
// (one per lane) in one step, like RegHeap::merge32: the result of pushing a set of items one by one is the k best of
// (heap U set) whatever the order, so every item's final position is its rank in the union — a binary search of each
// candidate in the sorted array, a count among the candidates, and an in-place shift of the retained items from the
// worst chunk of 32 down to the best (an item only ever moves towards the worse end, by the number of candidates that
// beat it, which grows with its position — so a chunk's writes never land on an unread item).
struct SmemHeap {
    float* hs;
    uint32_t* hk;
    uint32_t n, k, wkey;
    float theta;
    __device__ __forceinline__ void reset(uint32_t kk, float* s, uint32_t* ky) {
        hs = s, hk = ky, n = 0, k = kk, theta = 0.f, wkey = 0;
    }
    __device__ __forceinline__ bool full() const { return n == k; }
    // candidates of the lanes in `m`: keys distinct among themselves; each beats the current worst or the heap has room
    __device__ __forceinline__ void merge(uint32_t m, const float sc, const uint32_t ky, uint32_t lane) {
        bool c = (m >> lane) & 1u;
        // retained items better than this lane's candidate = its lower bound in the sorted array (fixed trip count)
        uint32_t hb = 0, len = n;
        const uint32_t steps = 32 - __clz(n);
        for (uint32_t st = 0; st < steps; ++st) {
            const uint32_t half = len >> 1;
            const bool right = len > 0 && better(hs[hb + half], hk[hb + half], sc, ky);
            hb = right ? hb + half + 1 : hb;
            len = right ? len - half - 1 : half;
        }
        if (c && hb < n && hk[hb] == ky) c = false;  // already retained (same document, same score, same position)
        m = __ballot_sync(0xffffffffu, c);
        if (!m) return;
        uint32_t cb = 0;  // candidates better than this lane's candidate
        for (uint32_t mm = m; mm; mm &= mm - 1) {
            const int j = __ffs(mm) - 1;
            cb += better(__shfl_sync(0xffffffffu, sc, j), __shfl_sync(0xffffffffu, ky, j), sc, ky);
        }
        __syncwarp();
        for (int base = n ? (int)((n - 1) & ~31u) : -32; base >= 0; base -= 32) {
            const uint32_t i = (uint32_t)base + lane;
            const bool valid = i < n;
            const float si = valid ? hs[i] : 0.f;
            const uint32_t ki = valid ? hk[i] : 0u;
            uint32_t up = 0;  // candidates better than this retained item
            for (uint32_t mm = m; mm; mm &= mm - 1) {
                const int j = __ffs(mm) - 1;
                up += better(__shfl_sync(0xffffffffu, sc, j), __shfl_sync(0xffffffffu, ky, j), si, ki);
            }
            __syncwarp();
            if (valid && up && i + up < k) hs[i + up] = si, hk[i + up] = ki;
            __syncwarp();
        }
        const uint32_t rc = hb + cb;
        if (c && rc < k) hs[rc] = sc, hk[rc] = ky;
        __syncwarp();
        n = min(k, n + (uint32_t)__popc(m));
        if (n == k) theta = hs[k - 1], wkey = hk[k - 1];
    }
    __device__ __forceinline__ void offer(bool have, const float sc, const uint32_t ky, uint32_t lane, bool distinct = true) {
        const uint32_t m = __ballot_sync(0xffffffffu, have && (n < k || better(sc, ky, theta, wkey)));
        if (!m) return;
        if (distinct) {
            merge(m, sc, ky, lane);
        } else {  // keys may repeat inside the call (kNN refine): one candidate at a time, re-filtered by the live heap
            for (uint32_t mm = m; mm; mm &= mm - 1) {
                const uint32_t one = mm & (0u - mm);
                const uint32_t ok = __ballot_sync(0xffffffffu, (one >> lane) & 1u && (n < k || better(sc, ky, theta, wkey)));
                if (ok) merge(ok, sc, ky, lane);
            }
        }
    }
    __device__ __forceinline__ void store_keys(uint32_t* dst, uint32_t lane) const {
        for (uint32_t i = lane; i < n; i += 32) dst[i] = hk[i];
    }
    __device__ __forceinline__ void write_sorted(uint32_t lane, uint32_t* out_keys, float* out_scores) const {
        for (uint32_t i = lane; i < k; i += 32) {
            out_keys[i] = i < n ? hk[i] : 0xffffffffu;
            out_scores[i] = i < n ? hs[i] : -INFINITY;
        }
    }
};

// KHeap for 32 < k <= 128: the retained items UNSORTED in registers, four slots per lane (item i lives in slot i / 32 of
// lane i % 32), scores as total_cmp keys, plus the current worst (theta, wkey).  A push overwrites the worst item and
// finds the new worst with two warp reductions (minimum score key, then the largest record offset among the items that
// have it) — ~45 instructions per change whatever k, against the ~150 dependent ones of a sorted shared-memory array
// (at k = 100 a query changes its heap ~480 times).  Pushing a set of items one by one gives the k best of
// (heap U set) in any order, and the order among the retained items is only needed once, for the output.
// Scores are never -0.0 or NaN here (sums that start from +0.0), so the integer order equals `better`'s float order.
struct WideHeap {
    uint32_t tk[4], ky[4];  // per lane: total_key(score), key; empty slot = (0xffffffff, 0xffffffff)
    uint32_t n, k, wkey, wtk;
    float theta;
    float* xs;  // k-entry buffers in shared memory (output sort)
    uint32_t* xk;
    __device__ __forceinline__ void reset(uint32_t kk, float* bs, uint32_t* bk) {
        n = 0, k = kk, theta = 0.f, wkey = 0, wtk = 0, xs = bs, xk = bk;
#pragma unroll
        for (int r = 0; r < 4; ++r) tk[r] = 0xffffffffu, ky[r] = 0xffffffffu;
    }
    __device__ __forceinline__ bool full() const { return n == k; }
    static __device__ __forceinline__ bool btr(uint32_t t, uint32_t key, uint32_t wt, uint32_t wk) {
        return t > wt || (t == wt && key < wk);
    }
    static __device__ __forceinline__ float untotal(uint32_t t) {
        return __uint_as_float((t & 0x80000000u) ? (t & 0x7fffffffu) : ~t);
    }
    __device__ __forceinline__ void find_worst() {  // heap full: empty slots (k % 32 != 0) hold the largest key
        wtk = __reduce_min_sync(0xffffffffu, min(min(tk[0], tk[1]), min(tk[2], tk[3])));
        uint32_t lk = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) lk = max(lk, tk[r] == wtk ? ky[r] : 0u);
        wkey = __reduce_max_sync(0xffffffffu, lk);
        theta = untotal(wtk);
    }
    __device__ __forceinline__ void offer(bool have, const float sc, const uint32_t key, uint32_t lane, bool = true) {
        const uint32_t ct = total_key(sc);
        uint32_t m = __ballot_sync(0xffffffffu, have && (n < k || btr(ct, key, wtk, wkey)));
        while (m) {  // warp-uniform: one candidate at a time against the live worst
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t bt = __shfl_sync(0xffffffffu, ct, src), bk = __shfl_sync(0xffffffffu, key, src);
            if (n == k && !btr(bt, bk, wtk, wkey)) continue;
            if (__any_sync(0xffffffffu, ky[0] == bk || ky[1] == bk || ky[2] == bk || ky[3] == bk)) continue;  // retained
            if (n < k) {
                const uint32_t r = n >> 5;
                if (lane == (n & 31u)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (r == (uint32_t)j) tk[j] = bt, ky[j] = bk;
                }
                if (++n == k) find_worst();
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (tk[j] == wtk && ky[j] == wkey) tk[j] = bt, ky[j] = bk;
                find_worst();
            }
        }
    }
    __device__ __forceinline__ void store_keys(uint32_t* dst, uint32_t lane) const {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r * 32 + lane < n) dst[r * 32 + lane] = ky[r];
    }
    // best first: every item's position is the number of retained items that beat it (keys are distinct)
    __device__ __forceinline__ void write_sorted(uint32_t lane, uint32_t* out_keys, float* out_scores) const {
        uint32_t* xt = reinterpret_cast<uint32_t*>(xs);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r * 32 + lane < n) xt[r * 32 + lane] = tk[r], xk[r * 32 + lane] = ky[r];
        __syncwarp();
        uint32_t rank[4] = {0, 0, 0, 0};
        for (uint32_t j = 0; j < n; ++j) {
            const uint32_t jt = xt[j], jk = xk[j];
#pragma unroll
            for (int r = 0; r < 4; ++r) rank[r] += btr(jt, jk, tk[r], ky[r]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r * 32 + lane < n) out_keys[rank[r]] = ky[r], out_scores[rank[r]] = untotal(tk[r]);
        for (uint32_t i = n + lane; i < k; i += 32) out_keys[i] = 0xffffffffu, out_scores[i] = -INFINITY;
        __syncwarp();
    }
};

// -----------------------------------------------------------------------------------------------------------
// T threads per CTA, OCC = CTAs per SM the register allocation is budgeted for, D = documents per 8-lane group
// per scoring iteration, Q = query representation, H = heap (RegHeap for k <= 32, WideHeap for k <= 128 on the
// benchmark layouts, SmemHeap otherwise),
// R = record layout (Rec16: u16 components, Rec32: u32 components).
// TMA = true (u16 / f16 layout, byte-index query, D = 2 only): the records are not gathered with per-lane 128-bit loads
// but staged round by round into a per-warp shared-memory buffer by the TMA unit — one cp.async.bulk per document
// and round (<= 256 contiguous bytes), completion on the warp's mbarrier.  The copy of the NEXT round (or of the next
// documents' first round) is issued as soon as the lanes have moved the current round into registers, i.e. it runs
// under the ~130 lookup / multiply-add instructions of the current round: one 2 KB stage per warp is a full double
// buffer, and no warp ever waits for a global load with its registers tied up.
template <int T, int OCC, int D, class Q, class H, class R = Rec16, bool TMA = false>
__global__ void __launch_bounds__(T, OCC) k_search(const SearchArgs a) {
    static_assert(!TMA || (D == 2 && R::PLAIN_F16 && Q::HAS_DOT8), "TMA staging: u16/f16 records, byte-index query, D = 2");
    constexpr int NW = T / 32;      // warps
    constexpr int GROUPS = T / 8;   // 8-lane groups
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (*a.n_list == 0) return;  // nothing routed to this instantiation (uniform): skip the table initialisation
    Q query;
    query.template init<T>(smem_raw, a, threadIdx.x);
    unsigned char* p = smem_raw + ((Q::bytes(a) + 15) & ~(size_t)15);
    uint32_t* cand_end = reinterpret_cast<uint32_t*>(p);  p += a.cand_cap * 4;
    float* cand_est = reinterpret_cast<float*>(p);        p += a.cand_cap * 4;
    uint32_t* cand_p0 = reinterpret_cast<uint32_t*>(p);   p += a.cand_cap * 4;
    uint32_t* cand_mx = reinterpret_cast<uint32_t*>(p);   p += a.cand_cap * 4;  // total_key of the block's best survivor, 0 = none
    float* heap_s = reinterpret_cast<float*>(p);          p += ((a.k + 3) & ~3u) * 4;
    uint32_t* heap_k = reinterpret_cast<uint32_t*>(p);    p += ((a.k + 3) & ~3u) * 4;
    uint64_t* docs = reinterpret_cast<uint64_t*>(p);      p += (size_t)a.buf_docs * 8;
    float* scores = reinterpret_cast<float*>(p);          p += (size_t)a.buf_docs * 4;
    uint32_t* surv = reinterpret_cast<uint32_t*>(p);      p += (size_t)((a.buf_docs + 31) / 32) * 4;  // bit d: document d of the wave can still enter the heap
    uint16_t* perm = reinterpret_cast<uint16_t*>(p);  // scoring order -> wave slot (longest documents first)
    p += (size_t)a.buf_docs * 2;
    // TMA variant: per warp one 2 KB stage (128-byte aligned) and one mbarrier
    const uint32_t tma_ring_s = ((uint32_t)__cvta_generic_to_shared(p) + 127u) & ~127u;
    __shared__ __align__(8) unsigned long long s_tma_bar[TMA ? NW : 1];
    uint32_t tma_par = 0;  // parity of the warp's next mbarrier phase
    if constexpr (TMA) {
        if ((threadIdx.x & 31) == 0) mbar_init((uint32_t)__cvta_generic_to_shared(&s_tma_bar[threadIdx.x >> 5]), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    __shared__ uint32_t s_q;
    __shared__ uint32_t s_warp_docs[32], s_warp_cnt[32];
    __shared__ uint32_t s_first_rej, s_wave_docs, s_wave_cnt, s_big_nd, s_big_p0, s_snap_n;
    __shared__ float s_theta;
    __shared__ uint32_t s_full, s_wkey;
    // phase clocks (thread 0 only, kept in shared memory to spare registers):
    // 0 fetch+stage, 1 select, 2 gather postings, 3 score, 4 replay, 5 results
    __shared__ uint32_t s_ph[6];
    __shared__ uint32_t s_cnt[2];  // waves scored, selection passes

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lane8 = tid & 7, grp = tid >> 3;
    const uint32_t k = a.k;

    // bytes of a posting's record (DotVByte: as if every chunk were wide)
    auto rec_bytes = [&](uint64_t pst) -> uint32_t {
        const uint32_t nch = ((uint32_t)(pst & 0xffffu) + 7) >> 3;
        if constexpr (is_vbyte<R>::value) return 16 + 24 * nch;
        else return nch * R::CHUNK_BYTES;
    };
    H heap;  // live in warp 0 only
    heap.reset(k, heap_s, heap_k);
    uint32_t st_docs = 0, st_blocks = 0, st_pushed = 0, st_units = 0;  // per-CTA totals fit 32 bits
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) s_ph[i] = 0;
        s_cnt[0] = s_cnt[1] = 0;
    }

    // Phase clock of thread 0.  Deliberately branch-free (one predicated reduction): an `if (tid == 0)` block can
    // leave thread 0 split from its warp across the code that follows, and every shuffle / vote there then takes its
    // divergent slow path (seen with ncu on a variant of this kernel: 4x slower replay).
    uint32_t t_last = (uint32_t)clock();
    const uint32_t ph_s = (uint32_t)__cvta_generic_to_shared(s_ph);
    auto lap = [&](int i) {
        const uint32_t now = (uint32_t)clock();
        asm volatile("{ .reg .pred p; setp.eq.u32 p, %0, 0; @p red.shared.add.u32 [%1], %2; }" ::"r"(tid),
                     "r"(ph_s + 4 * i), "r"(now - t_last)
                     : "memory");
        t_last = now;
    };

    // score docs[0, n) of the wave buffer into scores[] and mark the documents that can still enter the heap
    // nc > 0: also track the best surviving score of each of the wave's nc candidate blocks (cand_mx)
    auto note_survivor = [&](uint32_t d, float sc, uint32_t nc) {
        atomicOr(&surv[d >> 5], 1u << (d & 31));
        if (nc) {
            uint32_t lo = 0, hi = nc - 1;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (cand_end[mid] <= d) lo = mid + 1;
                else hi = mid;
            }
            atomicMax(&cand_mx[lo], total_key(sc));
        }
    };
    // the TMA variant of score_wave (see the template comment): same results, records staged by bulk async copies
    auto score_wave_tma = [&](uint32_t n, uint32_t nc) {
        if constexpr (TMA) {
        const bool w_full = s_full != 0;
        const float w_theta = s_theta;
        const uint32_t w_wkey = s_wkey;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_tma_bar[TMA ? warp : 0]);
        const uint32_t stage = tma_ring_s + warp * TMA_WARP_BYTES + (lane >> 3) * 2 * TMA_ROUND_BYTES;  // this group's two rounds
        const uint32_t sw = (lane8 >> 2) & 1u;
        auto load_posts = [&](uint32_t dbase, uint64_t (&post)[2], uint32_t (&slot)[2]) -> uint32_t {
            uint32_t mx = 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t d = dbase + 2 * grp + j;
                slot[j] = d < n ? perm[d] : 0xffffffffu;
                post[j] = d < n ? docs[slot[j]] : 0ull;
                mx = max(mx, (uint32_t)(post[j] & 0xffffu));
            }
            return max(1u, (__reduce_max_sync(0xffffffffu, mx) + 63) >> 6);  // >= 1: every iteration consumes one phase
        };
        auto issue = [&](const uint64_t (&post)[2], uint32_t r) {  // round r of the group's two documents
            if (lane8 == 0) {
                uint32_t bytes[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t total = (((uint32_t)(post[j] & 0xffffu) + 7) >> 3) * 32;
                    bytes[j] = total > TMA_ROUND_BYTES * r ? min((uint32_t)TMA_ROUND_BYTES, total - TMA_ROUND_BYTES * r) : 0u;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was just read through the generic proxy
                mbar_expect_tx(bar, bytes[0] + bytes[1]);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (bytes[j])
                        tma_load_1d(stage + j * TMA_ROUND_BYTES,
                                    reinterpret_cast<const char*>(a.ix.fwd) + (post[j] >> 16) * 32 + (size_t)TMA_ROUND_BYTES * r,
                                    bytes[j], bar);
            }
        };
        uint64_t post[2];
        uint32_t slot[2];
        uint32_t rounds = load_posts(0, post, slot);
        issue(post, 0);
        for (uint32_t dbase = 0; dbase < n; dbase += 2 * GROUPS) {  // CTA-uniform trip count
            {   // the records after the next ones -> L2 (lane: document lane8 / 4, 128-byte lines lane8 % 4 and + 4)
                const uint32_t dn = dbase + 4 * GROUPS + 2 * grp + (lane8 >> 2);
                if (dn < n) {
                    const uint64_t pn = docs[perm[dn]];
                    const uint32_t bytes = rec_bytes(pn), ln = (lane8 & 3) * 128;
                    const char* base = reinterpret_cast<const char*>(a.ix.fwd) + (pn >> 16) * 32;
                    if (ln < bytes) prefetch_l2(base + ln);
                    if (ln + 512 < bytes) prefetch_l2(base + ln + 512);
                }
            }
            uint64_t post_n[2];
            uint32_t slot_n[2];
            const uint32_t rounds_n = load_posts(dbase + 2 * GROUPS, post_n, slot_n);
            const uint32_t nch0 = ((uint32_t)(post[0] & 0xffffu) + 7) >> 3, nch1 = ((uint32_t)(post[1] & 0xffffu) + 7) >> 3;
            float acc[2] = {0.f, 0.f};
            for (uint32_t r = 0; r < rounds; ++r) {
                const uint32_t m = lane8 + 8 * r;
                mbar_wait(bar, tma_par);
                tma_par ^= 1u;
                uint4 c[2], v[2];
                c[0] = c[1] = v[0] = v[1] = make_uint4(0, 0, 0, 0);
                const uint32_t at = stage + 32 * lane8 + (SGPU_LD256 ? 0u : 16 * sw);
                if (m < nch0) c[0] = lds128(at), v[0] = lds128(at ^ 16u);
                if (m < nch1) c[1] = lds128(at + TMA_ROUND_BYTES), v[1] = lds128((at + TMA_ROUND_BYTES) ^ 16u);
                __syncwarp();  // every lane holds its chunks in registers: the stage is free
                if (r + 1 < rounds) issue(post, r + 1);
                else if (dbase + 2 * GROUPS < n) issue(post_n, 0);
                query.dot2x(acc[0], acc[1], c, v);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float sc = group_reduce(acc[j]);
                const uint32_t d = slot[j];
                if (lane8 == 0 && d != 0xffffffffu && (post[j] & 0xffffu)) {
                    scores[d] = sc;
                    st_units += ((uint32_t)(post[j] & 0xffffu) + 7) >> 3;
                    if (!w_full || better(sc, (uint32_t)(post[j] >> 16), w_theta, w_wkey)) note_survivor(d, sc, nc);
                }
            }
            post[0] = post_n[0], post[1] = post_n[1], slot[0] = slot_n[0], slot[1] = slot_n[1], rounds = rounds_n;
        }
            }
    };
    auto score_wave_ldg = [&](uint32_t n, uint32_t nc) {
        const bool w_full = s_full != 0;
        const float w_theta = s_theta;
        const uint32_t w_wkey = s_wkey;
        for (uint32_t dbase = 0; dbase < n; dbase += D * GROUPS) {  // CTA-uniform trip count
            if constexpr (D == 2) {
                // the next iteration's records -> L2 (lane: document lane8 / 4, 128-byte lines lane8 % 4 and + 4)
                const uint32_t dn = dbase + D * GROUPS + D * grp + (lane8 >> 2);
                if (dn < n) {
                    const uint64_t pn = docs[perm[dn]];
                    const uint32_t bytes = rec_bytes(pn), ln = (lane8 & 3) * 128;
                    const char* base = reinterpret_cast<const char*>(a.ix.fwd) + (pn >> 16) * R::UNIT_BYTES;
                    if (ln < bytes) prefetch_l2(base + ln);
                    if (ln + 512 < bytes) prefetch_l2(base + ln + 512);
                }
            }
            if constexpr (D == 1) {  // same, one document per group: its 8 lanes take the lines lane8 and lane8 + 8
                const uint32_t dn = dbase + GROUPS + grp;
                if (dn < n) {
                    const uint64_t pn = docs[perm[dn]];
                    const uint32_t bytes = rec_bytes(pn), ln = lane8 * 128;
                    const char* base = reinterpret_cast<const char*>(a.ix.fwd) + (pn >> 16) * R::UNIT_BYTES;
                    if (ln < bytes) prefetch_l2(base + ln);
                    if (ln + 1024 < bytes) prefetch_l2(base + ln + 1024);
                }
            }
            uint64_t post[D];
            uint32_t slot[D];
            uint32_t mx = 0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const uint32_t d = dbase + D * grp + j;  // a warp takes 4 * D consecutive positions of the order
                slot[j] = d < n ? perm[d] : 0xffffffffu;
                post[j] = d < n ? docs[slot[j]] : 0ull;  // nnz 0 -> no loads, score unused
                mx = max(mx, (uint32_t)(post[j] & 0xffffu));
            }
            const uint32_t rounds = (__reduce_max_sync(0xffffffffu, mx) + 63) >> 6;
            float acc[D];
            if constexpr (is_vbyte<R>::value)
                score_docs_vb<D>(reinterpret_cast<const uint8_t*>(a.ix.fwd), post, lane8, rounds, query, acc, st_units);
            else
                score_docs<D, R>(a.ix.fwd, post, lane8, rounds, query, a.value_scale, acc);
#pragma unroll
            for (int j = 0; j < D; ++j) {
                float s = group_reduce(acc[j]);
                if constexpr (is_vbyte<R>::value) s = __fmul_rn(s, a.value_scale);  // scale * 2^24, once per document
                const uint32_t d = slot[j];
                if (lane8 == 0 && d != 0xffffffffu && (post[j] & 0xffffu)) {  // nnz 0: an absent kNN neighbour
                    scores[d] = s;
                    if constexpr (!is_vbyte<R>::value) st_units += ((uint32_t)(post[j] & 0xffffu) + 7) >> 3;
                    // theta only grows: a document that cannot enter the heap as of the wave start never will
                    if (!w_full || better(s, (uint32_t)(post[j] >> 16), w_theta, w_wkey)) note_survivor(d, s, nc);
                }
            }
        }
    };
    auto score_wave = [&](uint32_t n, uint32_t nc) {
        if constexpr (TMA) score_wave_tma(n, nc);
        else score_wave_ldg(n, nc);
    };
    // Fill the wave buffer with n postings (loader(i) = posting of slot i) and fix the scoring order.  A warp scores
    // 4 * D documents at a time (8 consecutive positions of the order) and every lane walks ceil(nnz / 64) rounds of
    // the LONGEST of them, so mixing lengths wastes lanes (60 % of the lane slots carry data when documents come in
    // posting order).  Each warp fills a contiguous segment of the slots and orders it by rounds (>= 4, 3, 2, 1) with
    // ballots only — no atomics, no barrier: neighbours in the order then have the same number of rounds except at
    // the few bucket boundaries.  Results do not depend on the order (scores / survivors are stored by wave slot).
    // The caller provides the barrier that makes docs[] / perm[] visible.
    auto rounds_bucket = [&](uint64_t pst) -> uint32_t {  // 0: >= 4 rounds ... 3: <= 1 round
        const uint32_t r = ((uint32_t)(pst & 0xffffu) + 63) >> 6;
        return 4u - min(max(r, 1u), 4u);
    };
    auto fill_wave = [&](uint32_t n, auto loader) {
        const uint32_t seg = (((n + NW - 1) / NW) + 31) & ~31u;  // slots per warp, a multiple of 32
        const uint32_t s0 = min(n, warp * seg), s1 = min(n, s0 + seg);
        uint32_t cnt0 = 0, cnt1 = 0, cnt2 = 0;
        for (uint32_t i0 = s0; i0 < s1; i0 += 32) {  // warp-uniform trip count
            const uint32_t i = i0 + lane;
            uint32_t bk = 4;
            if (i < s1) {
                const uint64_t pst = loader(i);
                docs[i] = pst;
                bk = rounds_bucket(pst);
                if (!a.bucket) perm[i] = (uint16_t)i;
            }
            cnt0 += __popc(__ballot_sync(0xffffffffu, bk == 0));
            cnt1 += __popc(__ballot_sync(0xffffffffu, bk == 1));
            cnt2 += __popc(__ballot_sync(0xffffffffu, bk == 2));
        }
        if (!a.bucket) return;
        __syncwarp();
        uint32_t at[4] = {s0, s0 + cnt0, s0 + cnt0 + cnt1, s0 + cnt0 + cnt1 + cnt2};  // next free position per bucket
        for (uint32_t i0 = s0; i0 < s1; i0 += 32) {
            const uint32_t i = i0 + lane;
            const uint64_t pst = i < s1 ? docs[i] : 0ull;
            const uint32_t bk = i < s1 ? rounds_bucket(pst) : 4u;
#pragma unroll
            for (uint32_t bb = 0; bb < 4; ++bb) {
                const uint32_t m = __ballot_sync(0xffffffffu, bk == bb);
                if (bk == bb) {
                    const uint32_t pos = at[bb] + __popc(m & ((1u << lane) - 1u));
                    perm[pos] = (uint16_t)i;
                    if (pos < T / 4) {  // what the first scoring step of every warp reads: start the DRAM access now
                        const uint32_t bytes = rec_bytes(pst);
                        const char* base = reinterpret_cast<const char*>(a.ix.fwd) + (pst >> 16) * R::UNIT_BYTES;
                        for (uint32_t o = 0; o < bytes && o < 1024; o += 128) prefetch_l2(base + o);
                    }
                }
                at[bb] += __popc(m);
            }
        }
    };
    // warp 0: offer the surviving documents of wave slots [c0, c1) to the heap
    auto push_range = [&](uint32_t c0, uint32_t c1, bool distinct = true) {
        for (uint32_t base = c0 & ~31u; base < c1; base += 32) {
            uint32_t w = surv[base >> 5];  // warp-uniform
            if (base < c0) w &= 0xffffffffu << (c0 - base);
            if (c1 - base < 32) w &= (1u << (c1 - base)) - 1u;
            if (!w) continue;  // no survivor of this range in these 32 slots (a block usually spans two such groups)
            const uint32_t i = base + lane;
            const bool have = (w >> lane) & 1u;
            heap.offer(have, have ? scores[i] : 0.f, have ? (uint32_t)(docs[i] >> 16) : 0u, lane, distinct);
        }
    };

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
        heap.reset(k, heap_s, heap_k);
        if (tid == 0) s_full = 0, s_theta = 0.f, s_wkey = 0;
        __syncthreads();
        // (Loop A — summary estimates and the first list's block order — was also run INSIDE this kernel's CTAs, one warp
        // per list in the idle wave buffers, to hide its DRAM round trips under the other CTAs' scoring: measured 6.06
        // vs 5.68 ms per 10 k queries at cut 3, 7.56 vs 6.65 at cut 5 — the CTA holds a quarter of the SM for the 21-28 %
        // of its time the dependent loads take.  Dropped; the stand-alone k_est / k_order stay.)
        lap(0);

        bool first_wave = true;
        for (uint32_t t = 0; t < nt; ++t) {
            const uint32_t l = a.sc.terms[(uint64_t)q * a.sc.cut_eff + t];
            const ListHdr h = a.ix.lists[l];
            const uint32_t B = h.n_blk;
            const float* est = a.sc.est + ((uint64_t)q * a.sc.cut_eff + t) * a.sc.est_stride;
            const uint4* sel = (t == 0 && a.first_sorted) ? a.sc.sel + (uint64_t)q * a.sc.est_stride : nullptr;
            const uint32_t* boff = a.ix.blk_post_off + h.blk_base + l;
            const uint64_t* posts = a.ix.postings + h.post_base;
            uint32_t pos0 = 0;
            while (pos0 < B) {
                // ---------------- phase 1: candidate selection over positions [pos0, pos0 + SP * T), SP consecutive
                // positions per thread (a list of the benchmark index has ~350 blocks: one pass instead of two)
                constexpr int SP = SGPU_SEL_P;
                if (tid == 0) ++s_cnt[1];
                const bool full = s_full != 0;
                const float thr = __fmul_rn(a.heap_factor, s_theta);
                const uint32_t cap = first_wave ? a.first_wave_docs : a.wave_docs;
                bool pass[SP];
                uint32_t nd[SP], p0[SP], cdl[SP], ccl[SP];
                float e[SP];
                uint32_t cd_t = 0, cc_t = 0;  // this thread's totals
#pragma unroll
                for (int i = 0; i < SP; ++i) {  // independent loads: one memory round trip per selection pass
                    const uint32_t pos = pos0 + SP * tid + i;
                    pass[i] = false, nd[i] = 0, p0[i] = 0, e[i] = 0.f;
                    if (pos < B) {
                        if (sel) {
                            const uint4 se = __ldcg(sel + pos);
                            e[i] = __uint_as_float(se.x), p0[i] = se.y, nd[i] = se.z;
                        } else {
                            e[i] = __ldcg(est + pos);
                            p0[i] = boff[pos];
                            nd[i] = boff[pos + 1] - p0[i];
                        }
                        pass[i] = !full || !(e[i] < thr);
                        if (!pass[i]) nd[i] = 0;
                    }
                }
#pragma unroll
                for (int i = 0; i < SP; ++i) {
                    cd_t += nd[i], cc_t += pass[i] ? 1u : 0u;
                    cdl[i] = cd_t, ccl[i] = cc_t;
                }
                // block-wide inclusive scans of the documents and of the passing blocks
                uint32_t cd = cd_t, cc = cc_t;
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const uint32_t od = __shfl_up_sync(0xffffffffu, cd, sft);
                    const uint32_t oc = __shfl_up_sync(0xffffffffu, cc, sft);
                    if (lane >= (uint32_t)sft) cd += od, cc += oc;
                }
                // these blocks will most likely be in the wave: pull their postings towards L2 ahead of phase 2
                if (warp == 0) {
#pragma unroll
                    for (int i = 0; i < SP; ++i)
                        if (pass[i] && nd[i] && cd - cd_t + cdl[i] <= cap) {
                            const char* pp = reinterpret_cast<const char*>(posts + p0[i]);
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + (size_t)nd[i] * 8 - 8));
                        }
                }
                if (lane == 31) s_warp_docs[warp] = cd, s_warp_cnt[warp] = cc;
                if (tid == 0) s_first_rej = 0xffffffffu, s_wave_docs = 0, s_wave_cnt = 0, s_big_nd = 0;  // the previous
                                                                                   // wave's totals are consumed
                __syncthreads();
                // exclusive prefix of this thread (documents / passing blocks before its first position): every warp adds
                // up the totals of the warps before it itself — one barrier less than a scan by warp 0
                uint32_t bd = cd - cd_t, bc = cc - cc_t;
                for (uint32_t w = 0; w < warp; ++w) bd += s_warp_docs[w], bc += s_warp_cnt[w];
                // accept while the wave stays within its soft cap; the first passing block is accepted whenever it
                // fits the buffer; a first passing block larger than the buffer is processed alone, in parts.  The running
                // sums only grow, so the accepted blocks are a prefix of the passing ones: a passing block is in the wave
                // iff it is accepted, and the first rejected position is only needed to restart the next pass.
#pragma unroll
                for (int i = 0; i < SP; ++i) {
                    const uint32_t cdi = bd + cdl[i], cci = bc + ccl[i];
                    const bool accepted = pass[i] && cdi <= (cci == 1 ? a.buf_docs : cap) && cci <= a.cand_cap;
                    if (pass[i] && !accepted) atomicMin(&s_first_rej, pos0 + SP * tid + i);
                    if (pass[i] && cci == 1 && !accepted) s_big_nd = nd[i], s_big_p0 = p0[i];
                    if (accepted) {
                        cand_end[cci - 1] = cdi;
                        cand_est[cci - 1] = e[i];
                        cand_p0[cci - 1] = p0[i];
                        cand_mx[cci - 1] = 0u;
                        atomicMax(&s_wave_docs, cdi);
                        atomicMax(&s_wave_cnt, cci);
                    }
                }
                __syncthreads();
                const uint32_t first_rej = s_first_rej;
                const uint32_t n_docs = s_wave_docs, n_cand = s_wave_cnt, big_nd = s_big_nd, big_p0 = s_big_p0;
                lap(1);
                if (n_cand == 0 && big_nd) {
                    // ---------------- oversized block: it passed the test against the live theta (no other block is
                    // in flight), so the reference evaluates it; score and push it part by part
                    for (uint32_t off = 0; off < big_nd; off += a.buf_docs) {
                        const uint32_t part = min(a.buf_docs, big_nd - off);
                        fill_wave(part, [&](uint32_t i) { return posts[big_p0 + off + i]; });
                        for (uint32_t i = tid; i < (part + 31) >> 5; i += T) surv[i] = 0u;
                        __syncthreads();
                        score_wave(part, 0);
                        __syncthreads();
                        if (warp == 0) {
                            st_docs += part;
                            push_range(0, part);
                            if (lane == 0) s_full = heap.full(), s_theta = heap.theta, s_wkey = heap.wkey;
                        }
                        __syncthreads();
                    }
                    if (warp == 0) ++st_blocks, ++st_pushed;
                    first_wave = false;
                    pos0 = first_rej + 1;
                    lap(3);
                    continue;
                }
                pos0 = first_rej != 0xffffffffu ? first_rej : pos0 + SP * T;
                if (n_cand == 0) continue;
                first_wave = false;
                // ---------------- phase 2: gather the postings of all candidate blocks into the wave buffer
                // (flat over document slots; the owning block is found by binary search over cand_end)
                for (uint32_t i = tid; i < (n_docs + 31) >> 5; i += T) surv[i] = 0u;
                fill_wave(n_docs, [&](uint32_t i) {
                    uint32_t lo = 0, hi = n_cand - 1;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (cand_end[mid] <= i) lo = mid + 1;
                        else hi = mid;
                    }
                    const uint32_t start = lo ? cand_end[lo - 1] : 0u;
                    return posts[cand_p0[lo] + (i - start)];
                });
                __syncthreads();
                lap(2);
                // ---------------- phase 3: score the wave
                score_wave(n_docs, n_cand);
                __syncthreads();
                lap(3);
                // ---------------- phase 4: exact replay by warp 0
                if (warp == 0) {
                    st_docs += n_docs;
                    st_blocks += n_cand;
                    if (lane == 0) ++s_cnt[0];
                    // 32 candidate blocks at a time: a block whose best surviving score cannot enter the live heap
                    // only needs the skip test (counted, heap untouched); the first block that can AND passes the
                    // test against the live theta is pushed, which may raise theta, so the scan restarts after it.
                    // (Pushing a whole group speculatively and rolling back when a later block no longer passes
                    // against the final theta was measured: the rollbacks of the first, estimate-sorted list cost more
                    // than the merged pushes save — 5.33 vs 5.18 ms at k = 10, 26.4 vs 19.2 ms at k = 100.)
                    // (lane <-> candidate is fixed per group of 32, so a block's bounds, estimate and best survivor are
                    // read once; theta only grows, so after a push the remaining lanes are simply re-tested.)
                    for (uint32_t j0 = 0; j0 < n_cand; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        const bool valid = j < n_cand;
                        const uint32_t s0 = valid && j ? cand_end[j - 1] : 0u, s1 = valid ? cand_end[j] : 0u;
                        const uint32_t mx = valid ? cand_mx[j] : 0u;
                        const float ce = valid ? cand_est[j] : 0.f;
                        uint32_t live = __ballot_sync(0xffffffffu, valid);  // blocks of the group not yet decided
                        while (live) {
                            const bool has = mx != 0u && (!heap.full() || mx >= total_key(heap.theta));
                            const bool passes = valid && !(heap.full() && ce < __fmul_rn(a.heap_factor, heap.theta));
                            const uint32_t pm = __ballot_sync(0xffffffffu, passes) & live;
                            const uint32_t hm = __ballot_sync(0xffffffffu, passes && has) & live;
                            if (!hm) {
                                st_pushed += __popc(pm);
                                break;
                            }
                            const int f = __ffs(hm) - 1;
                            const uint32_t upto = (2u << f) - 1u;  // lanes 0 .. f (f = 31: all)
                            st_pushed += __popc(pm & upto);
                            push_range(__shfl_sync(0xffffffffu, s0, f), __shfl_sync(0xffffffffu, s1, f));
                            live &= ~upto;
                        }
                    }
                    if (lane == 0) s_full = heap.full(), s_theta = heap.theta, s_wkey = heap.wkey;
                }
                __syncthreads();
                lap(4);
            }
        }
        if (a.n_knn > 0 && nt > 0) {
            // ---------------- Knn::refine (src/inverted_index.rs:551-593): the first n_knn graph neighbours of every
            // document retained so far are scored and pushed.  The reference walks a sorted snapshot of the heap and
            // skips visited documents; the result is the k best of (heap U neighbours) in any order, and a document
            // seen before is either still retained (offer() ignores it) or can no longer enter (see the header).
            uint32_t* snap = cand_end;  // the four candidate arrays are contiguous: 4 * cand_cap >= k entries (host)
            __syncthreads();  // a selection pass that found nothing leaves the list loop without a barrier
            if (warp == 0) {
                heap.store_keys(snap, lane);
                if (lane == 0) s_snap_n = heap.n;
            }
            __syncthreads();
            const uint32_t n_snap = s_snap_n;
            for (uint32_t e = tid; e < n_snap; e += T) {  // id_from_range: key (record start) -> document
                const uint32_t key = snap[e];
                uint64_t lo = 0, hi = a.ix.n_docs + 1;
                while (lo < hi) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if (__ldg(a.ix.rec_start + mid) <= key) lo = mid + 1;
                    else hi = mid;
                }
                snap[e] = (uint32_t)(lo - 1);
            }
            __syncthreads();
            const uint32_t total = n_snap * a.n_knn;
            for (uint32_t base = 0; base < total; base += a.buf_docs) {
                const uint32_t part = min(a.buf_docs, total - base);
                fill_wave(part, [&](uint32_t i) {
                    const uint32_t c = base + i;
                    const uint64_t pst = a.ix.knn_posts[(uint64_t)snap[c / a.n_knn] * a.ix.knn_dim + c % a.n_knn];
                    return pst == ~0ull ? 0ull : pst;
                });
                for (uint32_t i = tid; i < (part + 31) >> 5; i += T) surv[i] = 0u;
                __syncthreads();
                score_wave(part, 0);
                __syncthreads();
                if (warp == 0) {
                    st_docs += part;
                    push_range(0, part, false);  // two retained documents may share a neighbour
                    if (lane == 0) s_full = heap.full(), s_theta = heap.theta, s_wkey = heap.wkey;
                }
                __syncthreads();
            }
        }
        // ---------------- results: best first, padded
        if (warp == 0) {
            heap.write_sorted(lane, a.sc.out_keys + (uint64_t)q * k, a.out_scores + (uint64_t)q * k);
            if (lane == 0) a.out_counts[q] = heap.n;
        }
        __syncthreads();
        if (nt > 0) query.template unstage<T>(a.b, qo, qn, tid);
        lap(5);
    }
    for (int sh = 16; sh > 0; sh >>= 1) st_units += __shfl_xor_sync(0xffffffffu, st_units, sh);
    if (lane == 0 && warp != 0 && a.sc.stats) atomicAdd(&a.sc.stats[3], (unsigned long long)st_units);
    if (tid == 0 && a.sc.stats) {
        atomicAdd(&a.sc.stats[0], (unsigned long long)st_docs);
        atomicAdd(&a.sc.stats[1], (unsigned long long)st_blocks);
        atomicAdd(&a.sc.stats[2], (unsigned long long)st_pushed);
        atomicAdd(&a.sc.stats[3], (unsigned long long)st_units);
        for (int i = 0; i < 6; ++i) atomicAdd(&a.sc.stats[4 + i], (unsigned long long)s_ph[i]);
        atomicAdd(&a.sc.stats[10], (unsigned long long)s_cnt[0]);
        atomicAdd(&a.sc.stats[11], (unsigned long long)s_cnt[1]);
    }
}

}  // namespace sgpu
