// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in seismic_b200/ may include, link or call this file;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// CPU restatement (C++17, scalar) of Seismic's query-time hot path, following the reference at
// /root/reference (commit e616de4) function by function:
//
//   oracle search driver      InvertedIndexBase::search            src/inverted_index.rs:153-234
//   term selection            k_largest_by(query_cut, value)       src/inverted_index.rs:187-190
//   block loop                PostingList::search                  src/posting_list.rs:115-146
//   sorted block loop         PostingList::sort_and_search         src/posting_list.rs:149-185
//   block evaluation          evaluate_posting_block               src/posting_list.rs:188-215
//   summary estimates         QuantizedSummary::distances          src/quantized_summary.rs:64-160
//   bounded heap              KHeap::{push,peek,into_sorted_vec}   src/utils.rs:12-66
//   posting unpack            PackedPostingBlock::unpack           src/posting_list.rs:54-59
//   result mapping            id_from_range                        src/inverted_index.rs:227-233
//
// PARITY PINNING.  The Rust reference cannot be built or imported in this image (no cargo/rustc,
// nightly toolchain, un-vendored git crates `vectorium` and `toolkit` with no pinned revision), so the
// oracle is pinned against every known-answer test the reference holds for this path:
//   src/inverted_index.rs:716-772 (test_empty_vectors), src/quantized_summary.rs:519-598
//   (test_distances_iter), docs/RustUsage.md:140-157 — see tests/test_oracle_golden.py.
// The following arithmetic lives in `vectorium` (absent) and is therefore DEFINED here, not pinned:
//   * summation order of the doc score (QueryEvaluator::compute_distance).  Two orders are provided:
//       ORDER_LANES8 (default) — element i goes to partial p[(i/8)%8]; p += q[c]*v (mul then add,
//         no FMA); result ((p0+p4)+(p2+p6))+((p1+p5)+(p3+p7)).  This is what an 8-lane SIMD gather
//         loop with a butterfly horizontal sum computes, and exactly what the CUDA kernel computes.
//       ORDER_SEQ — one accumulator, ascending component order.
//     Scores of the two orders differ by <= ~1e-6 relative; tests bound the difference by 1e-4.
//   * ties: equal scores are ordered by the smaller forward-index start offset first (a total order),
//     equal query values by the smaller component first, equal summary estimates (sorted first list)
//     by the smaller block id first.  The reference leaves all three unspecified (unstable sorts).
//   * duplicate query components: dense scatter keeps the LAST value; the summary merge matches the
//     FIRST occurrence (that is what the two-pointer merge of quantized_summary.rs:77-118 does).
//
// Build: see oracle/Makefile (g++ -O3 -march=native -ffp-contract=off).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/seismic_b200.h"

extern "C" {
typedef struct OracleStats {
    uint64_t n_queries;
    uint64_t lists_visited;
    uint64_t blocks_total;      // blocks of the visited lists
    uint64_t blocks_evaluated;  // not skipped
    uint64_t postings_seen;     // postings in evaluated blocks (incl. already visited docs)
    uint64_t docs_scored;       // first visits
    uint64_t results;           // returned tuples
    uint64_t bytes_summaries;
    uint64_t bytes_postings;    //               B
    uint64_t bytes_forward;     //               C
    uint64_t bytes_query_out;   //               D
    uint64_t bytes_total;
    double seconds;             // wall time of the search loop
} OracleStats;
}

namespace {

enum { ORDER_LANES8 = 0, ORDER_SEQ = 1 };

#if defined(__F16C__)
#include <immintrin.h>
inline float f16_to_f32(uint16_t h) { return _cvtsh_ss(h); }  // vcvtph2ps: exact, what the `half` crate uses on x86
#else
inline float f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, mant = h & 0x3ffu, x;
    if (exp == 0) {
        if (!mant) x = sign;
        else {
            int e = -1;
            do { mant <<= 1; e++; } while (!(mant & 0x400u));
            x = sign | ((uint32_t)(112 - e) << 23) | ((mant & 0x3ffu) << 13);
        }
    } else if (exp == 31) x = sign | 0x7f800000u | (mant << 13);
    else x = sign | ((exp + 112) << 23) | (mant << 13);
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}
#endif
inline float bf16_to_f32(uint16_t h) {
    uint32_t x = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}
inline uint32_t total_key(float f) {  // f32::total_cmp as an unsigned key
    uint32_t x;
    std::memcpy(&x, &f, 4);
    return (x & 0x80000000u) ? ~x : (x | 0x80000000u);
}

struct Item {
    float score;
    uint64_t start;
    uint32_t len;
};
// Ord of ScoredRange<DotProduct> as defined above: a < b  <=>  a is the better result.
inline bool better(const Item& a, const Item& b) {
    return a.score > b.score || (a.score == b.score && a.start < b.start);
}

// KHeap (src/utils.rs:12-66): BinaryHeap max-heap in Ord => top is the WORST retained item.
struct KHeap {
    std::vector<Item> h;
    size_t k;
    explicit KHeap(size_t kk) : k(kk) { h.reserve(kk); }
    static bool ord_less(const Item& a, const Item& b) { return better(a, b); }
    void push(const Item& it) {
        if (h.size() < k) {
            h.push_back(it);
            std::push_heap(h.begin(), h.end(), ord_less);
        } else if (better(it, h.front())) {  // item < *max
            std::pop_heap(h.begin(), h.end(), ord_less);
            h.back() = it;
            std::push_heap(h.begin(), h.end(), ord_less);
        }
    }
    size_t len() const { return h.size(); }
    const Item& peek() const { return h.front(); }
    std::vector<Item> into_sorted_vec() {
        std::sort_heap(h.begin(), h.end(), ord_less);  // ascending in Ord == best first
        return h;
    }
};

// FxHashSet<usize> stand-in: open addressing with epoch stamps (cleared in O(1) per query).
struct Visited {
    std::vector<uint64_t> key;
    std::vector<uint32_t> stamp;
    uint32_t epoch = 0;
    uint64_t mask = 0, used = 0;
    void reset(size_t want) {
        size_t cap = 1024;
        while (cap < want * 2) cap <<= 1;
        if (cap > key.size()) {
            key.assign(cap, 0);
            stamp.assign(cap, 0);
            epoch = 0;
        }
        mask = key.size() - 1;
        used = 0;
        if (++epoch == 0) {
            std::fill(stamp.begin(), stamp.end(), 0);
            epoch = 1;
        }
    }
    void grow() {
        std::vector<uint64_t> ok;
        ok.reserve(used);
        for (size_t i = 0; i < key.size(); ++i)
            if (stamp[i] == epoch) ok.push_back(key[i]);
        key.assign(key.size() * 2, 0);
        stamp.assign(key.size(), 0);
        mask = key.size() - 1;
        epoch = 1;
        used = 0;
        for (uint64_t x : ok) insert(x);
    }
    bool contains(uint64_t x) const {
        uint64_t i = (x * 0x9E3779B97F4A7C15ull) >> 20 & mask;
        while (stamp[i] == epoch) {
            if (key[i] == x) return true;
            i = (i + 1) & mask;
        }
        return false;
    }
    bool insert(uint64_t x) {  // true if newly inserted
        if (used * 2 >= key.size()) grow();
        uint64_t i = (x * 0x9E3779B97F4A7C15ull) >> 20 & mask;
        while (stamp[i] == epoch) {
            if (key[i] == x) return false;
            i = (i + 1) & mask;
        }
        stamp[i] = epoch;
        key[i] = x;
        ++used;
        return true;
    }
};

struct Ctx {  // per-thread scratch
    std::vector<float> qdense;
    std::vector<float> est;
    std::vector<uint32_t> order;
    Visited visited;
    OracleStats st{};
};

inline float decode(const SgpuIndexView& v, uint64_t i) {
    switch (v.value_kind) {
        case SGPU_VAL_F16: return f16_to_f32(((const uint16_t*)v.fwd_values)[i]);
        case SGPU_VAL_BF16: return bf16_to_f32(((const uint16_t*)v.fwd_values)[i]);
        case SGPU_VAL_F32: return ((const float*)v.fwd_values)[i];
        case SGPU_VAL_FIXEDU8: return (float)((const uint8_t*)v.fwd_values)[i] * v.value_scale;
        case SGPU_VAL_FIXEDU16: return (float)((const uint16_t*)v.fwd_values)[i] * v.value_scale;
        default: return 0.f;
    }
}
inline uint32_t value_bytes(uint32_t kind) {
    return kind == SGPU_VAL_F32 ? 4 : (kind == SGPU_VAL_FIXEDU8 ? 1 : 2);
}

// vectorium QueryEvaluator::compute_distance stand-in (dense query lookup), see header comment.
template <int ORDER, class CT>
float doc_score_t(const SgpuIndexView& v, const float* q, uint64_t start, uint32_t len) {
    const CT* c = (const CT*)v.fwd_comps + start;
    if (ORDER == ORDER_SEQ) {
        float acc = 0.f;
        for (uint32_t i = 0; i < len; ++i) acc = acc + q[c[i]] * decode(v, start + i);
        return acc;
    }
    float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (v.value_kind == SGPU_VAL_F16) {  // fast path for the headline encoding
        const uint16_t* h = (const uint16_t*)v.fwd_values + start;
        uint32_t i = 0;
        for (uint32_t m = 0; i + 8 <= len; ++m, i += 8) {
            float a = p[m & 7];
            for (int j = 0; j < 8; ++j) a = a + q[c[i + j]] * f16_to_f32(h[i + j]);
            p[m & 7] = a;
        }
        if (i < len) {  // tail chunk (fewer than 8 elements)
            const uint32_t lane = (i >> 3) & 7;
            float a = p[lane];
            for (; i < len; ++i) a = a + q[c[i]] * f16_to_f32(h[i]);
            p[lane] = a;
        }
    } else {
        for (uint32_t i = 0; i < len; ++i) p[(i >> 3) & 7] = p[(i >> 3) & 7] + q[c[i]] * decode(v, start + i);
    }
    return ((p[0] + p[4]) + (p[2] + p[6])) + ((p[1] + p[5]) + (p[3] + p[7]));
}
// DotVByte record (format: seismic_b200/csrc/host/build.cpp, convert_dotvbyte): `start` counts 16-byte units of the
// packed stream, `len` is the number of components.  The score is the fixed-point dot product scaled ONCE per
// document — (sum of q[c] * code) * scale — with the same partial-sum order as the other encodings.  So that the
// CUDA kernel can turn a code into a float with the half-precision conversion unit (the byte is read as an f16
// subnormal = code * 2^-24, exact), the sum is carried at 2^-24 of its value and the final factor is scale * 2^24;
// powers of two commute with every rounding involved (no overflow / underflow for |q| in [1e-30, 1e30]).
constexpr uint32_t VB_UNIT = 16;
inline uint32_t vb_record_bytes(const SgpuIndexView& v, uint64_t start, uint32_t len) {
    const uint8_t* rec = (const uint8_t*)v.fwd_values + start * VB_UNIT;
    const uint32_t nch = (len + 7) >> 3, ndir = (nch + 63) >> 6;
    uint32_t bytes = 16 * ndir + 16 * nch;  // directory, fixed parts; + 8 per wide chunk
    for (uint32_t t = 0; t < ndir; ++t) {
        uint64_t mask;
        std::memcpy(&mask, rec + 16ull * t, 8);
        bytes += 8 * (uint32_t)__builtin_popcountll(mask);
    }
    return bytes;
}
template <int ORDER>
float doc_score_vbyte(const SgpuIndexView& v, const float* q, uint64_t start, uint32_t len) {
    const uint8_t* rec = (const uint8_t*)v.fwd_values + start * VB_UNIT;
    const uint32_t nch = (len + 7) >> 3, ndir = (nch + 63) >> 6;
    const uint8_t* fixed = rec + 16ull * ndir;
    const uint8_t* wide = fixed + 16ull * nch;
    const float c24 = 5.9604644775390625e-08f;  // 2^-24
    float p[8] = {0, 0, 0, 0, 0, 0, 0, 0}, seq = 0.f;
    uint32_t c = 0, n_wide = 0;  // running component (the gaps are one chain over the record), wide chunks so far
    for (uint32_t m = 0; m < nch; ++m) {
        const uint8_t* fx = fixed + 16ull * m;
        uint64_t mask;
        std::memcpy(&mask, rec + 16ull * (m >> 6), 8);
        const uint8_t* hi = ((mask >> (m & 63)) & 1u) ? wide + 8ull * n_wide++ : nullptr;
        for (uint32_t f = 0; f < 8; ++f) {
            c += (uint32_t)fx[f] | (hi ? (uint32_t)hi[f] << 8 : 0u);
            if (m * 8 + f >= len) continue;  // tail padding: gap 0, code 0
            const float val = (float)fx[8 + f] * c24;
            if (ORDER == ORDER_SEQ) seq = seq + q[c] * val;
            else p[m & 7] = p[m & 7] + q[c] * val;
        }
    }
    const float s24 = v.value_scale * 16777216.f;
    if (ORDER == ORDER_SEQ) return seq * s24;
    return (((p[0] + p[4]) + (p[2] + p[6])) + ((p[1] + p[5]) + (p[3] + p[7]))) * s24;
}
template <int ORDER>
inline float doc_score(const SgpuIndexView& v, const float* q, uint64_t start, uint32_t len) {
    if (v.value_kind == SGPU_VAL_DOTVBYTE) return doc_score_vbyte<ORDER>(v, q, start, len);
    return v.comp_bits == 16 ? doc_score_t<ORDER, uint16_t>(v, q, start, len)
                             : doc_score_t<ORDER, uint32_t>(v, q, start, len);
}
// forward-index position (fwd_offsets units: elements, or bytes for DotVByte) of a posting's start field
inline uint64_t fwd_pos(const SgpuIndexView& v, uint64_t start) {
    return v.value_kind == SGPU_VAL_DOTVBYTE ? start * VB_UNIT : start;
}

inline uint32_t bits_for(uint64_t n_values) {  // BitField width for values in [0, n_values)
    uint32_t w = 1;
    while ((1ull << w) < n_values) ++w;
    return w;
}
inline uint32_t ceil_log2(uint64_t n) {
    uint32_t w = 0;
    while ((1ull << w) < n) ++w;
    return w;
}

// QuantizedSummary::distances for list `l` (src/quantized_summary.rs:64-160).  `est` gets B entries.
void summary_distances(const SgpuIndexView& v, uint64_t l, const uint32_t* qc, const float* qv, uint64_t nq,
                       float* est, OracleStats* st) {
    const uint64_t B = v.list_blk_start[l + 1] - v.list_blk_start[l];
    const float* mins = v.blk_min + v.list_blk_start[l];
    const float* quants = v.blk_quant + v.list_blk_start[l];
    const uint64_t sc0 = v.list_sc_start[l], nsc = v.list_sc_start[l + 1] - sc0;
    const uint32_t* sc = v.sc_comp + sc0;
    const uint32_t* run = v.sc_run_off + sc0 + l;
    const uint16_t* eb = v.ent_blk + v.list_ent_start[l];
    const uint8_t* ec = v.ent_code + v.list_ent_start[l];
    for (uint64_t b = 0; b < B; ++b) est[b] = 0.f;
    const uint32_t cbytes = v.comp_bits / 8;
    const uint32_t idw = bits_for(B ? B : 1);
    if (st) st->bytes_summaries += 8 * B + 4 * (B + 1);
    for (uint64_t j = 0; j < nq; ++j) {
        if (j > 0 && qc[j] == qc[j - 1]) continue;  // merge consumed the first duplicate only
        if (st) st->bytes_summaries += (uint64_t)cbytes * ceil_log2(nsc ? nsc : 1);
        const uint32_t* it = std::lower_bound(sc, sc + nsc, qc[j]);
        if (it == sc + nsc || *it != qc[j]) continue;
        uint64_t i = it - sc;
        const float w = qv[j];
        uint32_t b0 = run[i], b1 = run[i + 1];
        if (st) st->bytes_summaries += 8 + ((uint64_t)(b1 - b0) * (8 + idw) + 7) / 8;
        for (uint32_t e = b0; e < b1; ++e) {
            uint32_t s = eb[e];
            float deq = (float)ec[e] * quants[s] + mins[s];  // mul, add (no FMA: -ffp-contract=off)
            est[s] += deq * w;
        }
    }
}

template <int ORDER>
void search_one(const SgpuIndexView& v, const uint32_t* qc, const float* qv, uint64_t nq, const SgpuSearchParams& p,
                Ctx& cx, uint64_t* out_ids, float* out_scores, uint32_t* out_count) {
    OracleStats& st = cx.st;
    const uint32_t cbytes = v.comp_bits / 8, vbytes = value_bytes(v.value_kind);
    const bool vbyte = v.value_kind == SGPU_VAL_DOTVBYTE;
    float* q = cx.qdense.data();
    for (uint64_t i = 0; i < nq; ++i) q[qc[i]] = qv[i];  // dense evaluator; last duplicate wins
    st.bytes_query_out += nq * (cbytes + 4);

    KHeap heap(p.k);
    cx.visited.reset(std::min<uint64_t>(p.query_cut, nq) * 5000);

    // k_largest_by(query_cut, value): descending value, ties by position (== smaller component first)
    std::vector<uint32_t> terms(nq);
    for (uint64_t i = 0; i < nq; ++i) terms[i] = (uint32_t)i;
    auto term_cmp = [&](uint32_t a, uint32_t b) {
        uint32_t ka = total_key(qv[a]), kb = total_key(qv[b]);
        return ka != kb ? ka > kb : a < b;
    };
    size_t cut = std::min<size_t>(p.query_cut, nq);
    std::partial_sort(terms.begin(), terms.begin() + cut, terms.end(), term_cmp);

    for (size_t t = 0; t < cut; ++t) {
        const uint64_t l = qc[terms[t]];
        const uint64_t B = v.list_blk_start[l + 1] - v.list_blk_start[l];
        st.lists_visited++;
        st.blocks_total += B;
        if (cx.est.size() < B) cx.est.resize(B);
        float* est = cx.est.data();
        summary_distances(v, l, qc, qv, nq, est, &st);
        const uint32_t* boff = v.blk_post_off + v.list_blk_start[l] + l;
        const uint64_t* posts = v.postings + v.list_post_start[l];
        const bool sorted = (t == 0 && p.first_sorted);
        if (sorted) {
            cx.order.resize(B);
            for (uint64_t b = 0; b < B; ++b) cx.order[b] = (uint32_t)b;
            std::sort(cx.order.begin(), cx.order.end(), [&](uint32_t a, uint32_t b) {
                uint32_t ka = total_key(est[a]), kb = total_key(est[b]);
                return ka != kb ? ka > kb : a < b;
            });
        }
        for (uint64_t bi = 0; bi < B; ++bi) {
            const uint64_t b = sorted ? cx.order[bi] : bi;
            const float dot = est[b];
            if (heap.len() == p.k && dot < p.heap_factor * heap.peek().score) continue;
            st.blocks_evaluated++;
            const uint64_t p0 = boff[b], p1 = boff[b + 1];
            st.postings_seen += p1 - p0;
            st.bytes_postings += 8 * (p1 - p0);
            for (uint64_t i = p0; i < p1; ++i) {  // prefetch pass (src/posting_list.rs:198-204)
                uint64_t start = posts[i] >> 16;
                if (cx.visited.contains(start)) continue;
                // prefetch_with_range: every cache line of the vector's components and values
                const uint32_t plen = (uint32_t)(posts[i] & 0xffff);
                if (vbyte) {
                    const uint8_t* pr = (const uint8_t*)v.fwd_values + start * VB_UNIT;
                    for (uint32_t x = 0; x < plen * 3 + 64; x += 64) __builtin_prefetch(pr + x);
                } else {
                    const uint8_t* pc = (const uint8_t*)v.fwd_comps + start * cbytes;
                    const uint8_t* pv = (const uint8_t*)v.fwd_values + start * vbytes;
                    for (uint32_t x = 0; x < plen * cbytes; x += 64) __builtin_prefetch(pc + x);
                    for (uint32_t x = 0; x < plen * vbytes; x += 64) __builtin_prefetch(pv + x);
                }
            }
            for (uint64_t i = p0; i < p1; ++i) {
                uint64_t start = posts[i] >> 16;
                uint32_t len = (uint32_t)(posts[i] & 0xffff);
                if (cx.visited.insert(start)) {
                    st.docs_scored++;
                    st.bytes_forward += vbyte ? vb_record_bytes(v, start, len) : (uint64_t)len * (cbytes + vbytes);
                    heap.push(Item{doc_score<ORDER>(v, q, start, len), start, len});
                }
            }
        }
    }
    if (p.n_knn > 0 && v.knn_neighbours && v.knn_dim) {  // `if n_knn > 0 && let Some(knn)`, src/inverted_index.rs:215-217
        // Knn::refine (src/inverted_index.rs:551-593): snapshot of the heap, best first; for each retained document
        // its first min(dim, n_knn) graph neighbours; unvisited ones are scored and pushed.
        const uint32_t n_knn = std::min<uint32_t>(v.knn_dim, p.n_knn);
        KHeap copy = heap;  // heap.clone().into_sorted_vec()
        const std::vector<Item> snap = copy.into_sorted_vec();
        for (const Item& it : snap) {
            const uint64_t* ub = std::upper_bound(v.fwd_offsets, v.fwd_offsets + v.n_docs + 1, fwd_pos(v, it.start));
            const uint64_t id = (uint64_t)(ub - v.fwd_offsets) - 1;  // id_from_range
            for (uint32_t i = 0; i < n_knn; ++i) {
                const uint64_t nb = v.knn_neighbours[id * v.knn_dim + i];
                if (nb >= v.n_docs) continue;  // SGPU_PAD_ID: no neighbour (see include/seismic_b200.h)
                const uint64_t start = v.fwd_offsets[nb];  // range_from_id
                const uint32_t len = (uint32_t)(v.fwd_offsets[nb + 1] - start);
                if (len == 0) continue;  // an empty document is never a search result, hence never a neighbour
                if (cx.visited.insert(start)) {
                    st.docs_scored++;
                    st.bytes_forward += (uint64_t)len * (cbytes + vbytes);
                    heap.push(Item{doc_score<ORDER>(v, q, start, len), start, len});
                }
            }
        }
    }
    std::vector<Item> res = heap.into_sorted_vec();
    *out_count = (uint32_t)res.size();
    st.results += res.size();
    st.bytes_query_out += 12ull * res.size();
    for (uint32_t i = 0; i < p.k; ++i) {
        if (i < res.size()) {
            // id_from_range: the doc whose range starts here; duplicate offsets (empty docs) resolve to
            // the non-empty one (pinned by src/inverted_index.rs:716-772)
            const uint64_t* ub = std::upper_bound(v.fwd_offsets, v.fwd_offsets + v.n_docs + 1, fwd_pos(v, res[i].start));
            out_ids[i] = (uint64_t)(ub - v.fwd_offsets) - 1;
            out_scores[i] = res[i].score;
        } else {
            out_ids[i] = SGPU_PAD_ID;
            out_scores[i] = -INFINITY;
        }
    }
    for (uint64_t i = 0; i < nq; ++i) q[qc[i]] = 0.f;
}

int validate(const SgpuIndexView* v, const SgpuQueryBatch* qb, const SgpuSearchParams* p) {
    if (!v || !qb || !p || p->k == 0) return SGPU_EINVAL;
    if (p->n_knn != 0 && v->knn_neighbours && v->value_kind == SGPU_VAL_DOTVBYTE) return SGPU_EUNSUPPORTED;  // no kNN on DotVByte
    if (v->value_kind == SGPU_VAL_DOTVBYTE && (v->comp_bits != 16 || !v->fwd_nnz)) return SGPU_EUNSUPPORTED;
    for (uint64_t qi = 0; qi < qb->n_queries; ++qi)
        for (uint64_t i = qb->offsets[qi]; i < qb->offsets[qi + 1]; ++i) {
            if (qb->comps[i] >= v->dim) return SGPU_EINVAL;
            if (i > qb->offsets[qi] && qb->comps[i] < qb->comps[i - 1]) return SGPU_EINVAL;
        }
    return SGPU_OK;
}

void add_stats(OracleStats& a, const OracleStats& b) {
    a.lists_visited += b.lists_visited;
    a.blocks_total += b.blocks_total;
    a.blocks_evaluated += b.blocks_evaluated;
    a.postings_seen += b.postings_seen;
    a.docs_scored += b.docs_scored;
    a.results += b.results;
    a.bytes_summaries += b.bytes_summaries;
    a.bytes_postings += b.bytes_postings;
    a.bytes_forward += b.bytes_forward;
    a.bytes_query_out += b.bytes_query_out;
}

}  // namespace

extern "C" {

// Batched search.  acc_order: 0 = ORDER_LANES8, 1 = ORDER_SEQ.  n_threads: 1 = the reference's
// perf_inverted_index protocol (sequential loop, src/bin/perf_inverted_index.rs:184-210);
// >1 = static query ranges per thread (rayon par_iter stand-in, src/pylib/mod.rs:1129-1145); 0 = all cores.
// n_runs repeats the whole batch (timing protocol); results are those of the last run.
int oracle_batch_search(const SgpuIndexView* v, const SgpuQueryBatch* qb, const SgpuSearchParams* p, int acc_order,
                        int n_threads, int n_runs, uint64_t* out_ids, float* out_scores, uint32_t* out_counts,
                        OracleStats* stats) {
    int rc = validate(v, qb, p);
    if (rc != SGPU_OK) return rc;
    unsigned T = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    if (T > qb->n_queries) T = (unsigned)std::max<uint64_t>(1, qb->n_queries);
    if (n_runs < 1) n_runs = 1;
    std::vector<Ctx> ctx(T);
    for (auto& c : ctx) c.qdense.assign(v->dim, 0.f);
    auto work = [&](unsigned t, uint64_t b, uint64_t e) {
        Ctx& cx = ctx[t];
        for (uint64_t qi = b; qi < e; ++qi) {
            const uint64_t o = qb->offsets[qi], n = qb->offsets[qi + 1] - o;
            if (acc_order == ORDER_SEQ)
                search_one<ORDER_SEQ>(*v, qb->comps + o, qb->values + o, n, *p, cx, out_ids + qi * p->k,
                                      out_scores + qi * p->k, out_counts + qi);
            else
                search_one<ORDER_LANES8>(*v, qb->comps + o, qb->values + o, n, *p, cx, out_ids + qi * p->k,
                                         out_scores + qi * p->k, out_counts + qi);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    for (int run = 0; run < n_runs; ++run) {
        for (auto& c : ctx) c.st = OracleStats{};
        if (T == 1) {
            work(0, 0, qb->n_queries);
        } else {
            // dynamic chunks of 16 queries: per-query cost varies ~10x
            std::atomic<uint64_t> next{0};
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; ++t)
                th.emplace_back([&, t]() {
                    for (;;) {
                        uint64_t b = next.fetch_add(16);
                        if (b >= qb->n_queries) break;
                        work(t, b, std::min<uint64_t>(qb->n_queries, b + 16));
                    }
                });
            for (auto& x : th) x.join();
        }
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) {
        OracleStats s{};
        for (auto& c : ctx) add_stats(s, c.st);
        s.n_queries = qb->n_queries;
        s.bytes_total = s.bytes_summaries + s.bytes_postings + s.bytes_forward + s.bytes_query_out;
        s.seconds = secs / n_runs;
        *stats = s;
    }
    return SGPU_OK;
}

// QuantizedSummary::distances of one list, for the known-answer test of quantized_summary.rs:519-598.
int oracle_summary_distances(const SgpuIndexView* v, uint64_t list, const uint32_t* comps, const float* values,
                             uint64_t nnz, float* out_est) {
    if (!v || list >= v->dim) return SGPU_EINVAL;
    summary_distances(*v, list, comps, values, nnz, out_est, nullptr);
    return SGPU_OK;
}

// Exact top-k by brute force over the forward index (ground truth for recall; FlatIndex stand-in,
// src/inverted_index_wrapper.rs:721-742).  Same score arithmetic and tie rule as the search.
int oracle_exact_search(const SgpuIndexView* v, const SgpuQueryBatch* qb, uint32_t k, int n_threads,
                        uint64_t* out_ids, float* out_scores, uint32_t* out_counts) {
    SgpuSearchParams p{k, 1, 1.f, 0, 0};
    int rc = validate(v, qb, &p);
    if (rc != SGPU_OK) return rc;
    unsigned T = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&]() {
            std::vector<float> q(v->dim, 0.f);
            for (;;) {
                uint64_t qi = next.fetch_add(1);
                if (qi >= qb->n_queries) break;
                const uint64_t o = qb->offsets[qi], n = qb->offsets[qi + 1] - o;
                for (uint64_t i = 0; i < n; ++i) q[qb->comps[o + i]] = qb->values[o + i];
                KHeap heap(k);
                for (uint64_t d = 0; d < v->n_docs; ++d) {
                    uint64_t s = v->fwd_offsets[d];
                    uint32_t len = (uint32_t)(v->fwd_offsets[d + 1] - s);
                    if (v->value_kind == SGPU_VAL_DOTVBYTE) s /= VB_UNIT, len = v->fwd_nnz[d];
                    if (!len) continue;
                    heap.push(Item{doc_score<ORDER_LANES8>(*v, q.data(), s, len), s, len});
                }
                std::vector<Item> res = heap.into_sorted_vec();
                out_counts[qi] = (uint32_t)res.size();
                for (uint32_t i = 0; i < k; ++i) {
                    if (i < res.size()) {
                        const uint64_t* ub = std::upper_bound(v->fwd_offsets, v->fwd_offsets + v->n_docs + 1, fwd_pos(*v, res[i].start));
                        out_ids[qi * k + i] = (uint64_t)(ub - v->fwd_offsets) - 1;
                        out_scores[qi * k + i] = res[i].score;
                    } else {
                        out_ids[qi * k + i] = SGPU_PAD_ID;
                        out_scores[qi * k + i] = -INFINITY;
                    }
                }
                for (uint64_t i = 0; i < n; ++i) q[qb->comps[o + i]] = 0.f;
            }
        });
    for (auto& x : th) x.join();
    return SGPU_OK;
}

const char* oracle_version(void) { return "seismic oracle (C++ restatement of reference e616de4)"; }

}  // extern "C"
