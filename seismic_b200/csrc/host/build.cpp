// CPU index construction.  "Index build stays on CPU" (BASELINE.json north_star); this file
// restates the reference's build so that synthetic indexes have reference-like block and
// summary statistics.  It is NOT bit-compatible with the Rust build: the k-means centroid
// sample uses our own RNG (the reference uses rand::StdRng::choose_multiple, src/utils.rs:163-168)
// and every "ties unspecified" ordering of the reference is given a deterministic rule here.
//
//   pruning      global_threshold_pruning   reference src/inverted_index.rs:354-389
//                fixed_pruning              reference src/inverted_index.rs:293-329
//   blocking     blocking_with_random_kmeans reference src/posting_list.rs:227-300
//                RandomKmeansInvertedIndexApprox reference src/utils.rs:106-237
//                fixed_size_blocking        reference src/posting_list.rs:217-225
//   summaries    energy_preserving_summary  reference src/posting_list.rs:329-368
//                fixed_size_summary         reference src/posting_list.rs:302-327
//   quantize     reference src/utils.rs:68-90
//   inversion    QuantizedSummary::from     reference src/quantized_summary.rs:289-406
//   postings     PackedPostingBlock::pack   reference src/posting_list.rs:38-52,440-443
#include <chrono>
#include <cstdlib>
#include <memory>
#include <numeric>

#include "index.hpp"

namespace shost {

float decode_value(uint32_t kind, float scale, const void* values, uint64_t i) {
    switch (kind) {
        case SGPU_VAL_F16: return f16_bits_to_f32(((const uint16_t*)values)[i]);
        case SGPU_VAL_BF16: return bf16_bits_to_f32(((const uint16_t*)values)[i]);
        case SGPU_VAL_F32: return ((const float*)values)[i];
        case SGPU_VAL_FIXEDU8: return (float)((const uint8_t*)values)[i] * scale;
        case SGPU_VAL_FIXEDU16: return (float)((const uint16_t*)values)[i] * scale;
        default: return 0.f;
    }
}

static inline uint32_t value_bytes(uint32_t kind) {
    switch (kind) {
        case SGPU_VAL_F32: return 4;
        case SGPU_VAL_FIXEDU8: return 1;
        default: return 2;
    }
}

namespace {

static std::atomic<uint64_t> g_prof[8];
static inline uint64_t now_ns() { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Entry {  // pruned posting: (sortable value key, doc)
    uint32_t key;
    uint32_t doc;
};

struct ListOut {
    std::vector<uint64_t> postings;
    std::vector<uint32_t> blk_off;  // B+1 (always at least {0})
    std::vector<float> blk_min, blk_quant;
    std::vector<uint32_t> sc_comp;
    std::vector<uint32_t> sc_run_off;  // n_sc+1
    std::vector<uint16_t> ent_blk;
    std::vector<uint8_t> ent_code;
};

struct Fwd {  // read-only view of the encoded forward index
    const uint64_t* off;
    const void* comps;  // u16 or u32 by comp_bits
    uint32_t comp_bits;
    const void* vals;
    uint32_t kind;
    float scale;
    // per-document top-`doc_cut` components by value (descending), precomputed once: [n_docs * cut]
    const uint32_t* top_comp = nullptr;
    const float* top_val = nullptr;
    const uint8_t* top_n = nullptr;
    uint32_t cut = 0;
    inline uint32_t comp(uint64_t i) const {
        return comp_bits == 16 ? (uint32_t)((const uint16_t*)comps)[i] : ((const uint32_t*)comps)[i];
    }
    inline float val(uint64_t i) const {
        return kind == SGPU_VAL_F16 ? f16_bits_to_f32(((const uint16_t*)vals)[i]) : decode_value(kind, scale, vals, i);
    }
};

struct Scratch {
    std::vector<int32_t> slot;   // dim, -1
    struct Cell {
        uint32_t epoch;
        float v;
    };
    std::vector<Cell> cell;      // dim: per-block running max, validated by epoch
    uint32_t epoch = 0;
    std::vector<uint32_t> touched;
    std::vector<float> scores;
    std::vector<uint64_t> heap;  // (value key << 32 | ~component) composites for summary selection
    explicit Scratch(uint64_t dim) : slot(dim, -1), cell(dim, Cell{0, 0.f}) {}
};

// Rust `(x / q).round() as u8`: round half away from zero, saturating cast, NaN -> 0.
inline uint8_t round_sat_u8(float x) {
    if (!(x == x)) return 0;
    float r = std::round(x);
    if (r <= 0.f) return 0;
    if (r >= 255.f) return 255;
    return (uint8_t)r;
}

// top-`cut` components of a doc by value: descending value, ties by position (k_largest_by, src/utils.rs:125-127)
void top_components(const Fwd& f, uint64_t doc, uint32_t cut, std::vector<std::pair<uint32_t, uint32_t>>& top,
                    uint32_t* out_comp, float* out_val, uint8_t* out_n) {
    uint64_t b = f.off[doc], e = f.off[doc + 1];
    top.clear();
    for (uint64_t i = b; i < e; ++i) top.emplace_back(f32_total_key(f.val(i)), (uint32_t)(i - b));
    auto cmp = [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& c) {
        return a.first != c.first ? a.first > c.first : a.second < c.second;
    };
    if (top.size() > cut) {
        std::partial_sort(top.begin(), top.begin() + cut, top.end(), cmp);
        top.resize(cut);
    } else {
        std::sort(top.begin(), top.end(), cmp);
    }
    *out_n = (uint8_t)top.size();
    for (size_t i = 0; i < top.size(); ++i) {
        out_comp[i] = f.comp(b + top[i].second);
        out_val[i] = f.val(b + top[i].second);
    }
}

struct CentroidIndex {
    std::vector<uint32_t> off;   // per distinct comp slot
    std::vector<uint32_t> cent;
    std::vector<float> val;
};

// reference compute_centroid_assignments_approx_dot_product, src/utils.rs:106-144
void assign_docs(const Fwd& f, const std::vector<uint32_t>& docs, const std::vector<uint32_t>& centroid_docs,
                 const CentroidIndex& ci, const std::vector<char>& removed, Scratch& s,
                 std::vector<std::pair<uint32_t, uint32_t>>& out) {
    const size_t nc = centroid_docs.size();
    s.scores.resize(nc);
    for (uint32_t doc : docs) {
        std::fill(s.scores.begin(), s.scores.end(), 0.f);
        const uint32_t* tc = f.top_comp + (uint64_t)doc * f.cut;
        const float* tv = f.top_val + (uint64_t)doc * f.cut;
        for (uint32_t j = 0, n = f.top_n[doc]; j < n; ++j) {
            int32_t sl = s.slot[tc[j]];
            if (sl < 0) continue;
            const float v = tv[j];
            for (uint32_t e = ci.off[sl]; e < ci.off[sl + 1]; ++e) s.scores[ci.cent[e]] += ci.val[e] * v;
        }
        // Rust max_by keeps the LAST maximum; candidates filtered by `to_avoid`.
        int64_t best = -1;
        uint32_t best_key = 0;
        for (size_t c = 0; c < nc; ++c) {
            if (removed[c]) continue;
            uint32_t k = f32_total_key(s.scores[c]);
            if (best < 0 || k >= best_key) best = (int64_t)c, best_key = k;
        }
        if (best < 0) best = 0;  // unwrap_or((&centroids_doc_ids[0], &0.0))
        out.emplace_back(centroid_docs[best], doc);
    }
}

void build_list(const Fwd& f, uint64_t dim, const ShostBuildConfig& cfg, const Entry* entries, size_t len,
                Scratch& s, ListOut& out, std::string& err) {
    out.blk_off.assign(1, 0);
    out.sc_run_off.assign(1, 0);
    if (len == 0) return;
    std::vector<uint32_t> pl(len);
    for (size_t i = 0; i < len; ++i) pl[i] = entries[i].doc;

    uint64_t t0 = now_ns();
    // ---------------- blocking ----------------
    std::vector<uint32_t> block_offsets;
    if (cfg.blocking == 1) {  // fixed_size_blocking (src/posting_list.rs:217-225, quirks kept)
        uint32_t bs = std::max<uint32_t>(1, cfg.block_size);
        for (size_t i = 0; i < len / bs; ++i) block_offsets.push_back((uint32_t)(i * bs));
        if (block_offsets.empty() || block_offsets.back() != len) block_offsets.push_back((uint32_t)len);
        if (block_offsets.size() == 1) {  // list shorter than block_size: zero blocks, postings unreachable
            pl.clear();
            len = 0;
            block_offsets.assign(1, 0);
        }
    } else {
        size_t n_centroids = std::max<size_t>(1, (size_t)(cfg.centroid_fraction * (float)len));
        if (n_centroids > 65535) {
            err = "number of centroids > u16::MAX; decrease centroid_fraction";
            return;
        }
        // sample centroids without replacement (own RNG; same seed for every list like the reference)
        Rng rng(cfg.kmeans_seed, 0);
        std::vector<uint32_t> perm(len);
        std::iota(perm.begin(), perm.end(), 0u);
        std::vector<uint32_t> centroid_docs(n_centroids);
        for (size_t i = 0; i < n_centroids; ++i) {
            size_t j = i + (size_t)rng.below(len - i);
            std::swap(perm[i], perm[j]);
            centroid_docs[i] = pl[perm[i]];
        }
        // inverted index over the centroid docs (src/utils.rs:171-178)
        CentroidIndex ci;
        s.touched.clear();
        std::vector<uint32_t> counts;
        for (uint32_t cd : centroid_docs)
            for (uint64_t i = f.off[cd]; i < f.off[cd + 1]; ++i) {
                uint32_t c = f.comp(i);
                if (s.slot[c] < 0) {
                    s.slot[c] = (int32_t)s.touched.size();
                    s.touched.push_back(c);
                    counts.push_back(0);
                }
                counts[s.slot[c]]++;
            }
        ci.off.assign(counts.size() + 1, 0);
        for (size_t i = 0; i < counts.size(); ++i) ci.off[i + 1] = ci.off[i] + counts[i];
        ci.cent.resize(ci.off.back());
        ci.val.resize(ci.off.back());
        std::vector<uint32_t> fill(ci.off.begin(), ci.off.end() - 1);
        for (size_t ce = 0; ce < centroid_docs.size(); ++ce) {
            uint32_t cd = centroid_docs[ce];
            for (uint64_t i = f.off[cd]; i < f.off[cd + 1]; ++i) {
                uint32_t p = fill[s.slot[f.comp(i)]]++;
                ci.cent[p] = (uint32_t)ce;
                ci.val[p] = f.val(i);
            }
        }
        std::vector<char> removed(n_centroids, 0);
        std::vector<std::pair<uint32_t, uint32_t>> assign;
        assign.reserve(len);
        assign_docs(f, pl, centroid_docs, ci, removed, s, assign);
        std::sort(assign.begin(), assign.end());
        // dissolve too-small clusters (src/utils.rs:189-226)
        std::vector<uint32_t> to_reassign;
        std::vector<std::pair<uint32_t, uint32_t>> final_assign;
        final_assign.reserve(len);
        // map centroid doc id -> centroid index
        std::vector<std::pair<uint32_t, uint32_t>> cd_index(n_centroids);
        for (size_t i = 0; i < n_centroids; ++i) cd_index[i] = {centroid_docs[i], (uint32_t)i};
        std::sort(cd_index.begin(), cd_index.end());
        for (size_t g = 0; g < assign.size();) {
            size_t h = g;
            while (h < assign.size() && assign[h].first == assign[g].first) ++h;
            if (h - g <= cfg.min_cluster_size) {
                for (size_t i = g; i < h; ++i) to_reassign.push_back(assign[i].second);
                auto it = std::lower_bound(cd_index.begin(), cd_index.end(),
                                           std::make_pair(assign[g].first, (uint32_t)0));
                removed[it->second] = 1;
            } else {
                final_assign.insert(final_assign.end(), assign.begin() + g, assign.begin() + h);
            }
            g = h;
        }
        assign_docs(f, to_reassign, centroid_docs, ci, removed, s, final_assign);
        std::sort(final_assign.begin(), final_assign.end());
        for (uint32_t c : s.touched) s.slot[c] = -1;
        // groups -> blocks (src/posting_list.rs:279-297)
        block_offsets.push_back(0);
        for (size_t g = 0; g < final_assign.size();) {
            size_t h = g;
            while (h < final_assign.size() && final_assign[h].first == final_assign[g].first) ++h;
            for (size_t i = g; i < h; ++i) pl[i] = final_assign[i].second;
            block_offsets.push_back((uint32_t)h);
            g = h;
        }
    }
    const size_t B = block_offsets.size() - 1;
    if (B > 65535) {
        err = "Number of summaries cannot be more than 2^16";
        return;
    }
    out.blk_off = block_offsets;

    uint64_t t1 = now_ns(); g_prof[0] += t1 - t0;
    // ---------------- summaries + quantization ----------------
    struct Triple {
        uint32_t comp;
        uint16_t blk;
        uint8_t code;
    };
    std::vector<Triple> triples;
    std::vector<std::pair<uint32_t, float>> cv;  // (comp, max value)
    out.blk_min.resize(B);
    out.blk_quant.resize(B);
    for (size_t b = 0; b < B; ++b) {
        uint64_t ta = now_ns();
        s.touched.clear();
        if (++s.epoch == 0) {
            for (auto& ce : s.cell) ce.epoch = 0;
            s.epoch = 1;
        }
        if (b + 1 < B)
            for (uint32_t p = block_offsets[b + 1]; p < block_offsets[b + 2]; ++p) {
                const uint64_t o = f.off[pl[p]], cb = f.comp_bits / 8;
                for (uint64_t x = 0; x < 256; x += 64) {
                    __builtin_prefetch((const char*)f.comps + o * cb + x);
                    __builtin_prefetch((const char*)f.vals + o * 2 + x);
                }
            }
        for (uint32_t p = block_offsets[b]; p < block_offsets[b + 1]; ++p) {
            uint32_t doc = pl[p];
            for (uint64_t i = f.off[doc]; i < f.off[doc + 1]; ++i) {
                uint32_t c = f.comp(i);
                float v = f.val(i);
                Scratch::Cell& ce = s.cell[c];  // one cache line per component: epoch stamp + running max
                if (ce.epoch != s.epoch) {
                    ce.epoch = s.epoch;
                    ce.v = v;
                    s.touched.push_back(c);
                } else if (ce.v < v) {
                    ce.v = v;
                }
            }
        }
        uint64_t tb = now_ns(); g_prof[3] += tb - ta;
        // selection of the summary components.  The reference sorts all components by value (descending, ties
        // unspecified) and keeps a prefix; we pop the same prefix from a max-heap of (value, smaller component
        // first) composites, which avoids sorting the ~10x larger tail.
        s.heap.clear();
        float total = 0.f;
        for (uint32_t c : s.touched) {
            const float mv = s.cell[c].v;
            total += mv;
            s.heap.push_back(((uint64_t)f32_total_key(mv) << 32) | (uint32_t)(0xffffffffu - c));
        }
        // descending prefix by rounds: nth_element moves the next `take` largest to the front, they are sorted and
        // consumed in order; the round size doubles until the stop condition is met (typically in round one)
        auto decode = [](uint64_t x) -> std::pair<uint32_t, float> {
            uint32_t k = (uint32_t)(x >> 32), bits = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
            float v;
            std::memcpy(&v, &bits, 4);
            return {0xffffffffu - (uint32_t)x, v};
        };
        cv.clear();
        const bool fixed = cfg.summarization == 1;  // fixed_size_summary: the n_components largest
        const float until = total * cfg.summary_energy;  // energy_preserving_summary: take_while_inclusive
        float acc = 0.f;
        size_t done = 0, round = 96;
        bool reached = s.heap.empty() || (fixed && cfg.n_components == 0);
        while (!reached && done < s.heap.size()) {
            const size_t take = std::min(round, s.heap.size() - done);
            if (done + take < s.heap.size())
                std::nth_element(s.heap.begin() + done, s.heap.begin() + done + take, s.heap.end(),
                                 std::greater<uint64_t>());
            std::sort(s.heap.begin() + done, s.heap.begin() + done + take, std::greater<uint64_t>());
            for (size_t i = done; i < done + take; ++i) {
                cv.push_back(decode(s.heap[i]));
                if (fixed) {
                    if (cv.size() >= cfg.n_components) { reached = true; break; }
                } else {
                    acc += cv.back().second;
                    if (!(acc < until)) { reached = true; break; }
                }
            }
            done += take;
            round *= 2;
        }
        std::sort(cv.begin(), cv.end(),
                  [](const std::pair<uint32_t, float>& a, const std::pair<uint32_t, float>& c) { return a.first < c.first; });
        uint64_t tc = now_ns(); g_prof[4] += tc - tb;
        // quantize (src/utils.rs:68-90); a block always has >= 1 doc; empty docs only give empty summaries
        float mn = 0.f, quant = 0.f;
        if (!cv.empty()) {
            uint32_t kmin = f32_total_key(cv[0].second), kmax = kmin;
            float mx = cv[0].second;
            mn = cv[0].second;
            for (auto& x : cv) {
                uint32_t k = f32_total_key(x.second);
                if (k < kmin) kmin = k, mn = x.second;
                if (k > kmax) kmax = k, mx = x.second;
            }
            quant = (mx - mn) / 255.0f;
        }
        out.blk_min[b] = mn;
        out.blk_quant[b] = quant;
        for (auto& x : cv) triples.push_back({x.first, (uint16_t)b, round_sat_u8((x.second - mn) / quant)});
    }
    uint64_t t2 = now_ns(); g_prof[1] += t2 - t1;
    // ---------------- inversion by component (src/quantized_summary.rs:303-402) ----------------
    s.touched.clear();
    for (auto& t : triples)
        if (s.slot[t.comp] < 0) {
            s.slot[t.comp] = 0;
            s.touched.push_back(t.comp);
        }
    std::sort(s.touched.begin(), s.touched.end());
    for (size_t i = 0; i < s.touched.size(); ++i) s.slot[s.touched[i]] = (int32_t)i;
    out.sc_comp = s.touched;
    out.sc_run_off.assign(s.touched.size() + 1, 0);
    for (auto& t : triples) out.sc_run_off[s.slot[t.comp] + 1]++;
    for (size_t i = 0; i < s.touched.size(); ++i) out.sc_run_off[i + 1] += out.sc_run_off[i];
    out.ent_blk.resize(triples.size());
    out.ent_code.resize(triples.size());
    {
        std::vector<uint32_t> fill(out.sc_run_off.begin(), out.sc_run_off.end() - 1);
        for (auto& t : triples) {  // triples are in ascending block order -> runs keep ascending summary id
            uint32_t p = fill[s.slot[t.comp]]++;
            out.ent_blk[p] = t.blk;
            out.ent_code[p] = t.code;
        }
    }
    for (uint32_t c : s.touched) s.slot[c] = -1;
    uint64_t t3 = now_ns(); g_prof[2] += t3 - t2;
    // ---------------- packed postings ----------------
    out.postings.resize(len);
    for (size_t i = 0; i < len; ++i) {
        uint64_t st = f.off[pl[i]], l = f.off[pl[i] + 1] - st;
        out.postings[i] = (st << 16) | l;
    }
    (void)dim;
}

}  // namespace

int build_index(const ShostDataset& ds, const ShostBuildConfig& cfg_in, ShostIndex** out) {
    ShostBuildConfig cfg = cfg_in;
    const uint64_t N = ds.n_vecs, dim = ds.dim, nnz = ds.offsets.empty() ? 0 : ds.offsets.back();
    const unsigned T = hw_threads(cfg.n_threads);
    if (cfg.comp_bits != 16 && cfg.comp_bits != 32) { set_error("comp_bits must be 16 or 32"); return SGPU_EINVAL; }
    if (cfg.comp_bits == 16 && dim > 65536) { set_error("dim > 65536 needs comp_bits=32 (SeismicIndexLV)"); return SGPU_EINVAL; }
    if (cfg.value_kind > SGPU_VAL_FIXEDU16) { set_error("build: unsupported value_kind (build f16 then convert)"); return SGPU_EUNSUPPORTED; }
    if (N >= (1ull << 32)) { set_error("more than 2^32 documents"); return SGPU_EUNSUPPORTED; }
    if (nnz >= (1ull << 48)) { set_error("range.start exceeds 48-bit packing limit"); return SGPU_EINVAL; }
    // validate
    std::atomic<int> bad{0};
    parallel_for(N, 65536, T, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t d = b; d < e; ++d) {
            uint64_t s0 = ds.offsets[d], s1 = ds.offsets[d + 1];
            if (s1 < s0 || s1 - s0 > 65535) { bad = 1; return; }
            for (uint64_t i = s0; i < s1; ++i) {
                if (ds.comps[i] >= dim) { bad = 2; return; }
                if (i > s0 && ds.comps[i] <= ds.comps[i - 1]) { bad = 3; return; }
            }
        }
    });
    if (bad == 1) { set_error("range length exceeds 16-bit packing limit"); return SGPU_EINVAL; }
    if (bad == 2) { set_error("component id >= dim"); return SGPU_EINVAL; }
    if (bad == 3) { set_error("document components must be strictly ascending"); return SGPU_EINVAL; }

    const bool prof = std::getenv("SHOST_PROFILE") != nullptr;
    uint64_t w0 = now_ns();
    for (auto& g : g_prof) g = 0;
    auto lap = [&](const char* what) {
        if (prof) {
            uint64_t w1 = now_ns();
            std::fprintf(stderr, "[build] %-22s %8.2f s\n", what, (w1 - w0) * 1e-9);
            w0 = w1;
        }
    };
    auto* idx = new ShostIndex();
    idx->comp_bits = cfg.comp_bits;
    idx->value_kind = cfg.value_kind;
    idx->n_docs = N;
    idx->dim = dim;
    idx->nnz = nnz;
    idx->config = cfg;
    idx->sec[SEC_FWD_OFFSETS].adopt(ds.offsets);
    if (cfg.comp_bits == 16) {
        uint16_t* c16 = idx->sec[SEC_FWD_COMPS].alloc<uint16_t>(nnz);
        const uint32_t* src = ds.comps.data();
        parallel_for(nnz, 1 << 20, T, [&](uint64_t b, uint64_t e, unsigned) {
            for (uint64_t i = b; i < e; ++i) c16[i] = (uint16_t)src[i];
        });
    } else {
        idx->sec[SEC_FWD_COMPS].adopt(ds.comps);
    }
    // encode values
    float scale = 1.f;
    if (cfg.value_kind == SGPU_VAL_FIXEDU8 || cfg.value_kind == SGPU_VAL_FIXEDU16) {
        float mx = 0.f;
        for (uint64_t i = 0; i < nnz; ++i) mx = std::max(mx, ds.values[i]);
        float levels = cfg.value_kind == SGPU_VAL_FIXEDU8 ? 255.f : 65535.f;
        scale = mx > 0.f ? mx / levels : 1.f;
    }
    idx->value_scale = scale;
    {
        uint32_t vb = value_bytes(cfg.value_kind);
        uint8_t* dst = idx->sec[SEC_FWD_VALUES].alloc<uint8_t>(nnz * vb);
        const float* src = ds.values.data();
        const uint32_t kind = cfg.value_kind;
        parallel_for(nnz, 1 << 20, T, [&](uint64_t b, uint64_t e, unsigned) {
            for (uint64_t i = b; i < e; ++i) {
                float v = src[i];
                switch (kind) {
                    case SGPU_VAL_F16: ((uint16_t*)dst)[i] = f32_to_f16_bits(v); break;
                    case SGPU_VAL_BF16: ((uint16_t*)dst)[i] = f32_to_bf16_bits(v); break;
                    case SGPU_VAL_F32: ((float*)dst)[i] = v; break;
                    case SGPU_VAL_FIXEDU8: {
                        float r = std::nearbyint(v / scale);
                        ((uint8_t*)dst)[i] = (uint8_t)std::min(255.f, std::max(0.f, r));
                        break;
                    }
                    default: {
                        float r = std::nearbyint(v / scale);
                        ((uint16_t*)dst)[i] = (uint16_t)std::min(65535.f, std::max(0.f, r));
                    }
                }
            }
        });
    }
    Fwd f{idx->sec[SEC_FWD_OFFSETS].as<uint64_t>(), idx->sec[SEC_FWD_COMPS].ptr, cfg.comp_bits,
          idx->sec[SEC_FWD_VALUES].ptr, cfg.value_kind, scale};
    // per-document top-doc_cut components (used by every list the document appears in)
    std::vector<uint32_t> top_comp;
    std::vector<float> top_val;
    std::vector<uint8_t> top_n;
    if (cfg.blocking == 0) {
        const uint32_t cut = std::min<uint32_t>(std::max<uint32_t>(cfg.doc_cut, 1), 255);
        top_comp.resize(N * cut);
        top_val.resize(N * cut);
        top_n.resize(N);
        parallel_for(N, 4096, T, [&](uint64_t b, uint64_t e, unsigned) {
            std::vector<std::pair<uint32_t, uint32_t>> tmp;
            for (uint64_t d = b; d < e; ++d)
                top_components(f, d, cut, tmp, &top_comp[d * cut], &top_val[d * cut], &top_n[d]);
        });
        f.top_comp = top_comp.data();
        f.top_val = top_val.data();
        f.top_n = top_n.data();
        f.cut = cut;
    }

    lap("encode+top-cut");
    // ------------------------------------------------------------------ pruning
    std::vector<uint64_t> list_start(dim + 1, 0);
    std::vector<Entry> entries;
    {
        const uint64_t tot = cfg.pruning == 0 ? dim * (uint64_t)cfg.n_postings : nnz;
        const size_t cap = cfg.pruning == 0 ? (size_t)((float)cfg.n_postings * cfg.max_fraction) : (size_t)cfg.n_postings;
        uint32_t thr = 0;        // select key > thr, plus the first need_eq entries (doc order) with key == thr
        uint64_t need_eq = 0;
        bool take_all = nnz <= tot;
        const unsigned P = T;
        if (!take_all) {
            // two-level radix select of the tot-th largest key
            std::vector<std::vector<uint64_t>> h(P, std::vector<uint64_t>(65536, 0));
            parallel_parts(N, P, [&](unsigned p, uint64_t b, uint64_t e) {
                auto& hh = h[p];
                for (uint64_t i = f.off[b]; i < f.off[e]; ++i) hh[f32_total_key(f.val(i)) >> 16]++;
            });
            uint64_t above = 0;
            int hi = 65535;
            for (; hi >= 0; --hi) {
                uint64_t c = 0;
                for (unsigned p = 0; p < P; ++p) c += h[p][hi];
                if (above + c >= tot) break;
                above += c;
            }
            for (auto& hh : h) std::fill(hh.begin(), hh.end(), 0);
            parallel_parts(N, P, [&](unsigned p, uint64_t b, uint64_t e) {
                auto& hh = h[p];
                for (uint64_t i = f.off[b]; i < f.off[e]; ++i) {
                    uint32_t k = f32_total_key(f.val(i));
                    if ((int)(k >> 16) == hi) hh[k & 0xffff]++;
                }
            });
            int lo = 65535;
            for (; lo >= 0; --lo) {
                uint64_t c = 0;
                for (unsigned p = 0; p < P; ++p) c += h[p][lo];
                if (above + c >= tot) break;
                above += c;
            }
            thr = ((uint32_t)hi << 16) | (uint32_t)lo;
            need_eq = tot - above;
        }
        // per-part per-component counts
        std::vector<std::vector<uint32_t>> cnt(P, std::vector<uint32_t>(dim, 0));
        std::vector<std::vector<uint32_t>> eq_comps(P);
        parallel_parts(N, P, [&](unsigned p, uint64_t b, uint64_t e) {
            auto& c = cnt[p];
            for (uint64_t i = f.off[b]; i < f.off[e]; ++i) {
                if (take_all) { c[f.comp(i)]++; continue; }
                uint32_t k = f32_total_key(f.val(i));
                if (k > thr) c[f.comp(i)]++;
                else if (k == thr) eq_comps[p].push_back(f.comp(i));
            }
        });
        std::vector<uint64_t> eq_quota(P, 0);
        {
            uint64_t left = need_eq;
            for (unsigned p = 0; p < P; ++p) {
                uint64_t q = std::min<uint64_t>(left, eq_comps[p].size());
                eq_quota[p] = q;
                left -= q;
                for (uint64_t i = 0; i < q; ++i) cnt[p][eq_comps[p][i]]++;
                std::vector<uint32_t>().swap(eq_comps[p]);
            }
        }
        for (uint64_t c = 0; c < dim; ++c) {
            uint64_t s = 0;
            for (unsigned p = 0; p < P; ++p) s += cnt[p][c];
            list_start[c + 1] = list_start[c] + s;
        }
        entries.resize(list_start[dim]);
        // write cursors: cnt[p][c] becomes the absolute start for part p
        for (uint64_t c = 0; c < dim; ++c) {
            uint64_t s = list_start[c];
            for (unsigned p = 0; p < P; ++p) {
                uint32_t n = cnt[p][c];
                cnt[p][c] = (uint32_t)(s - list_start[c]);
                s += n;
            }
        }
        parallel_parts(N, P, [&](unsigned p, uint64_t b, uint64_t e) {
            auto& c = cnt[p];
            uint64_t eq_seen = 0;
            for (uint64_t d = b; d < e; ++d)
                for (uint64_t i = f.off[d]; i < f.off[d + 1]; ++i) {
                    uint32_t k = f32_total_key(f.val(i));
                    bool sel = take_all || k > thr || (k == thr && eq_seen++ < eq_quota[p]);
                    if (sel) {
                        uint32_t comp = f.comp(i);
                        entries[list_start[comp] + c[comp]++] = Entry{k, (uint32_t)d};
                    }
                }
        });
        // per list: (value desc, doc asc), cap
        std::vector<uint64_t> new_len(dim);
        parallel_for(dim, 64, T, [&](uint64_t b, uint64_t e, unsigned) {
            for (uint64_t c = b; c < e; ++c) {
                Entry* s0 = entries.data() + list_start[c];
                size_t n = list_start[c + 1] - list_start[c];
                std::stable_sort(s0, s0 + n, [](const Entry& a, const Entry& x) { return a.key > x.key; });
                new_len[c] = std::min(n, cap);
            }
        });
        lap("pruning");
        // lists keep their slots in `entries`; only the first new_len[c] entries of a slot are used
        std::vector<ListOut> outs(dim);
        std::vector<std::string> errs(T);
        std::vector<std::unique_ptr<Scratch>> scratch(T);
        parallel_for(dim, 8, T, [&](uint64_t b, uint64_t e, unsigned t) {
            if (!scratch[t]) scratch[t].reset(new Scratch(dim));
            for (uint64_t c = b; c < e; ++c) {
                if (!errs[t].empty()) return;
                build_list(f, dim, cfg, entries.data() + list_start[c], new_len[c], *scratch[t], outs[c], errs[t]);
            }
        });
        for (auto& er : errs)
            if (!er.empty()) {
                set_error(er);
                delete idx;
                return SGPU_EINVAL;
            }
        std::vector<Entry>().swap(entries);
        lap("per-list build");
        if (prof)
            std::fprintf(stderr, "[build]   cpu-s: blocking %.1f summaries %.1f (reduce %.1f select %.1f) inversion %.1f\n",
                         g_prof[0] * 1e-9, g_prof[1] * 1e-9, g_prof[3] * 1e-9, g_prof[4] * 1e-9, g_prof[2] * 1e-9);
        // ------------------------------------------------------------------ concatenate
        std::vector<uint64_t> lps(dim + 1, 0), lbs(dim + 1, 0), lss(dim + 1, 0), les(dim + 1, 0);
        for (uint64_t c = 0; c < dim; ++c) {
            lps[c + 1] = lps[c] + outs[c].postings.size();
            lbs[c + 1] = lbs[c] + outs[c].blk_min.size();
            lss[c + 1] = lss[c] + outs[c].sc_comp.size();
            les[c + 1] = les[c] + outs[c].ent_blk.size();
        }
        idx->sec[SEC_LIST_POST_START].adopt(lps);
        idx->sec[SEC_LIST_BLK_START].adopt(lbs);
        idx->sec[SEC_LIST_SC_START].adopt(lss);
        idx->sec[SEC_LIST_ENT_START].adopt(les);
        uint64_t* postings = idx->sec[SEC_POSTINGS].alloc<uint64_t>(lps[dim]);
        uint32_t* blk_post_off = idx->sec[SEC_BLK_POST_OFF].alloc<uint32_t>(lbs[dim] + dim);
        float* blk_min = idx->sec[SEC_BLK_MIN].alloc<float>(lbs[dim]);
        float* blk_quant = idx->sec[SEC_BLK_QUANT].alloc<float>(lbs[dim]);
        uint32_t* sc_comp = idx->sec[SEC_SC_COMP].alloc<uint32_t>(lss[dim]);
        uint32_t* sc_run_off = idx->sec[SEC_SC_RUN_OFF].alloc<uint32_t>(lss[dim] + dim);
        uint16_t* ent_blk = idx->sec[SEC_ENT_BLK].alloc<uint16_t>(les[dim]);
        uint8_t* ent_code = idx->sec[SEC_ENT_CODE].alloc<uint8_t>(les[dim]);
        parallel_for(dim, 64, T, [&](uint64_t b, uint64_t e, unsigned) {
            for (uint64_t c = b; c < e; ++c) {
                ListOut& o = outs[c];
                auto cp = [](auto* dst, const auto& v) {
                    if (!v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
                };
                cp(postings + lps[c], o.postings);
                cp(blk_post_off + lbs[c] + c, o.blk_off);
                cp(blk_min + lbs[c], o.blk_min);
                cp(blk_quant + lbs[c], o.blk_quant);
                cp(sc_comp + lss[c], o.sc_comp);
                cp(sc_run_off + lss[c] + c, o.sc_run_off);
                cp(ent_blk + les[c], o.ent_blk);
                cp(ent_code + les[c], o.ent_code);
                o = ListOut();  // release the list's buffers as soon as they are copied: the sections are first-touched
                                // while the per-list copies go away, instead of both being resident at the end
            }
        });
    }
    lap("concatenate");
    *out = idx;
    return SGPU_OK;
}

void fill_view(const ShostIndex& idx, SgpuIndexView* v) {
    std::memset(v, 0, sizeof(*v));
    v->comp_bits = idx.comp_bits;
    v->value_kind = idx.value_kind;
    v->n_docs = idx.n_docs;
    v->dim = idx.dim;
    v->value_scale = idx.value_scale;
    v->fwd_offsets = idx.sec[SEC_FWD_OFFSETS].as<uint64_t>();
    v->fwd_comps = idx.sec[SEC_FWD_COMPS].bytes ? idx.sec[SEC_FWD_COMPS].ptr : nullptr;
    v->fwd_values = idx.sec[SEC_FWD_VALUES].ptr;
    v->fwd_nnz = idx.sec[SEC_FWD_NNZ].bytes ? idx.sec[SEC_FWD_NNZ].as<uint16_t>() : nullptr;
    v->list_post_start = idx.sec[SEC_LIST_POST_START].as<uint64_t>();
    v->postings = idx.sec[SEC_POSTINGS].as<uint64_t>();
    v->list_blk_start = idx.sec[SEC_LIST_BLK_START].as<uint64_t>();
    v->blk_post_off = idx.sec[SEC_BLK_POST_OFF].as<uint32_t>();
    v->blk_min = idx.sec[SEC_BLK_MIN].as<float>();
    v->blk_quant = idx.sec[SEC_BLK_QUANT].as<float>();
    v->list_sc_start = idx.sec[SEC_LIST_SC_START].as<uint64_t>();
    v->sc_comp = idx.sec[SEC_SC_COMP].as<uint32_t>();
    v->list_ent_start = idx.sec[SEC_LIST_ENT_START].as<uint64_t>();
    v->sc_run_off = idx.sec[SEC_SC_RUN_OFF].as<uint32_t>();
    v->ent_blk = idx.sec[SEC_ENT_BLK].as<uint16_t>();
    v->ent_code = idx.sec[SEC_ENT_CODE].as<uint8_t>();
}

}  // namespace shost

// ---------------------------------------------------------------------------------------------------------
// DotVByte conversion.  The reference builds a standard u16/f16 index and then converts the
// forward index to `PackedSparseDataset<DotVByteFixedU8Encoder>` (src/pylib/dotvbyte.rs:195-213), re-packing the
// postings to the packed storage's ranges (src/inverted_index.rs:237-275).  The byte format of that encoder lives
// in the absent `vectorium` crate, so this is OUR documented format (not byte-compatible): a variable-byte code of the
// component gaps (1 or 2 bytes per gap) + u8 fixed-point values, laid out so that a GPU lane decodes one chunk of 8
// components from TWO aligned loads.  (Measured on the previous layout — one control bit per component, per-chunk
// exception groups behind a u16 offset table, five loads per chunk: the decode arithmetic was free, the three extra
// loads were the whole gap to the uncompressed index.  Hence the control bit per CHUNK: all eight gaps of a chunk take
// one byte, or all take two.)
//
//   value    u8 code, value = code * scale, scale = (largest f16 value of the collection) / 255   (FixedU8)
//   gaps     gap_i = component_i - component_{i-1} (gap_0 = component_0): ONE chain over the whole record, < 2^16 each
//   record   16-byte aligned; components in chunks of 8 (nch = ceil(nnz/8)), chunks in super-rounds of 64:
//     dir    16 bytes per super-round: u64 mask (bit b = chunk 64 t + b is WIDE: one of its gaps is >= 256),
//            u32 number of wide chunks in the earlier super-rounds, u32 zero
//     fixed  16 bytes per chunk: lo[8] = LOW bytes of its gaps, then val[8] = its codes (tail of the last chunk: gap 0,
//            code 0)
//     wide   8 bytes per WIDE chunk, in chunk order: the HIGH bytes of its gaps
//     zero padding to a multiple of 16 bytes
//   fwd_offsets[i]  byte offset of record i;  fwd_nnz[i] number of components;  postings = (offset/16 << 16) | nnz
// A GPU lane handles chunk m = lane + 8 r in round r: its wide bit and the rank of its wide entry come from the
// directory entry (loaded once per 8 rounds) by popcount, its loads are the 16 fixed bytes and (if wide) the 8 high
// bytes; the chunk's first component is the running total of the record so far, an 8-lane shuffle scan per round.
namespace shost {

int convert_dotvbyte(const ShostIndex& in, ShostIndex** out) {
    if (in.comp_bits != 16 || in.value_kind != SGPU_VAL_F16) {
        set_error("DotVByte conversion needs a u16/f16 index");
        return SGPU_EUNSUPPORTED;
    }
    const uint64_t N = in.n_docs;
    const uint64_t* off = in.sec[SEC_FWD_OFFSETS].as<uint64_t>();
    const uint16_t* comps = in.sec[SEC_FWD_COMPS].as<uint16_t>();
    const uint16_t* vals = in.sec[SEC_FWD_VALUES].as<uint16_t>();
    const unsigned T = hw_threads(in.config.n_threads);
    float mx = 0.f;
    for (uint64_t i = 0; i < in.nnz; ++i) mx = std::max(mx, f16_bits_to_f32(vals[i]));
    const float scale = mx > 0.f ? mx / 255.f : 1.f;
    // gap i of a record: one chain over the whole record
    auto gap = [&](uint64_t s, uint64_t n, uint64_t i) -> uint32_t {
        if (i >= n) return 0u;
        return i == 0 ? (uint32_t)comps[s] : (uint32_t)comps[s + i] - (uint32_t)comps[s + i - 1];
    };
    auto chunk_wide = [&](uint64_t s, uint64_t n, uint32_t m) -> bool {
        for (uint32_t f = 0; f < 8; ++f)
            if (gap(s, n, (uint64_t)m * 8 + f) >= 256) return true;
        return false;
    };
    // pass 1: record sizes
    std::vector<uint64_t> boff(N + 1, 0);
    parallel_for(N, 8192, T, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t d = b; d < e; ++d) {
            const uint64_t s = off[d], n = off[d + 1] - s;
            const uint32_t nch = (uint32_t)((n + 7) >> 3), ndir = (nch + 63) >> 6;
            uint64_t bytes = 16ull * ndir + 16ull * nch;  // directory, fixed parts
            for (uint32_t m = 0; m < nch; ++m) bytes += chunk_wide(s, n, m) ? 8 : 0;
            boff[d + 1] = (bytes + 15) & ~15ull;
        }
    });
    for (uint64_t d = 0; d < N; ++d) boff[d + 1] += boff[d];
    if ((boff[N] >> 4) >= (1ull << 48)) { set_error("packed forward index too large"); return SGPU_EUNSUPPORTED; }
    auto* idx = new ShostIndex();
    idx->comp_bits = 16;
    idx->value_kind = SGPU_VAL_DOTVBYTE;
    idx->n_docs = N;
    idx->dim = in.dim;
    idx->nnz = in.nnz;
    idx->value_scale = scale;
    idx->config = in.config;
    idx->sec[SEC_FWD_OFFSETS].adopt(boff);
    uint8_t* stream = idx->sec[SEC_FWD_VALUES].alloc<uint8_t>(boff[N] + 32);  // 32 bytes of slack for wide loads
    std::memset(stream + boff[N], 0, 32);
    uint16_t* nnzs = idx->sec[SEC_FWD_NNZ].alloc<uint16_t>(N);
    parallel_for(N, 8192, T, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t d = b; d < e; ++d) {
            const uint64_t s = off[d], n = off[d + 1] - s;
            const uint32_t nch = (uint32_t)((n + 7) >> 3), ndir = (nch + 63) >> 6;
            nnzs[d] = (uint16_t)n;
            uint8_t* rec = stream + boff[d];
            std::memset(rec, 0, boff[d + 1] - boff[d]);
            uint8_t* fixed = rec + 16ull * ndir;
            uint8_t* wide = fixed + 16ull * nch;
            uint32_t n_wide = 0;
            for (uint32_t m = 0; m < nch; ++m) {
                if ((m & 63) == 0) std::memcpy(rec + 16ull * (m >> 6) + 8, &n_wide, 4);  // wide chunks before this super-round
                uint8_t* fx = fixed + 16ull * m;
                const bool w = chunk_wide(s, n, m);
                for (uint32_t f = 0; f < 8; ++f) {
                    const uint64_t i = (uint64_t)m * 8 + f;
                    const float r = i < n ? std::nearbyint(f16_bits_to_f32(vals[s + i]) / scale) : 0.f;
                    fx[8 + f] = (uint8_t)std::min(255.f, std::max(0.f, r));
                    const uint32_t g = gap(s, n, i);
                    fx[f] = (uint8_t)(g & 0xff);
                    if (w) wide[8ull * n_wide + f] = (uint8_t)(g >> 8);
                }
                if (w) {
                    uint64_t mask;
                    std::memcpy(&mask, rec + 16ull * (m >> 6), 8);
                    mask |= 1ull << (m & 63);
                    std::memcpy(rec + 16ull * (m >> 6), &mask, 8);
                    ++n_wide;
                }
            }
        }
    });
    // posting lists: same blocks and summaries; postings re-packed to (byte offset / 16, nnz)
    for (int sidx : {SEC_LIST_POST_START, SEC_LIST_BLK_START, SEC_BLK_POST_OFF, SEC_BLK_MIN, SEC_BLK_QUANT,
                     SEC_LIST_SC_START, SEC_SC_COMP, SEC_LIST_ENT_START, SEC_SC_RUN_OFF, SEC_ENT_BLK, SEC_ENT_CODE}) {
        idx->sec[sidx].own.assign(in.sec[sidx].ptr, in.sec[sidx].ptr + in.sec[sidx].bytes);
        idx->sec[sidx].ptr = idx->sec[sidx].own.data();
        idx->sec[sidx].bytes = in.sec[sidx].bytes;
    }
    const uint64_t P = in.sec[SEC_POSTINGS].count<uint64_t>();
    const uint64_t* pin = in.sec[SEC_POSTINGS].as<uint64_t>();
    uint64_t* pout = idx->sec[SEC_POSTINGS].alloc<uint64_t>(P);
    parallel_for(P, 1 << 16, T, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t i = b; i < e; ++i) {
            const uint64_t start = pin[i] >> 16;
            const uint64_t d = (uint64_t)(std::upper_bound(off, off + N + 1, start) - off) - 1;
            pout[i] = ((boff[d] >> 4) << 16) | (pin[i] & 0xffff);
        }
    });
    *out = idx;
    return SGPU_OK;
}

}  // namespace shost
