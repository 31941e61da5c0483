// Synthetic SPLADE-shaped corpus / query generator.  Deterministic: every vector is a
// pure function of (seed, vector id) through a counter-based RNG, so any thread count produces the
// same bytes and queries can re-derive the document they were sampled from.
//
// Topic model: T latent topics, each an ordered set of `topic_terms` vocabulary entries drawn from a
// global Zipf term popularity; a document mixes 1-3 topics, draws most of its terms from the topics
// (Zipf(1) over the topic's term positions, so every topic has a small "core" shared by its documents)
// and the rest from the global popularity.  Values are log-normal (fitted on the reference's
// examples/toy_dataset: log-mean -1.44, log-std 1.32, max 2.63) with a boost for topic-core terms,
// clipped to [1e-3, 3.5].  Components are strictly ascending inside a vector.
#include <memory>
#include <mutex>
#include <unordered_map>

#include "index.hpp"

namespace shost {
namespace {

struct Model {
    uint64_t dim = 0, seed = 0;
    uint32_t n_topics = 0, topic_terms = 0;
    std::vector<uint32_t> rank_to_term;   // popularity rank -> component id
    Alias global;                         // over popularity ranks
    Alias position;                       // Zipf(1) over topic positions
    std::vector<uint32_t> topic_table;    // n_topics * topic_terms component ids
    std::vector<float> lognorm;           // 65536 quantiles of exp(N(0,1))
};

std::shared_ptr<Model> get_model(const ShostSynthConfig& c) {
    static std::mutex mu;
    static std::unordered_map<std::string, std::shared_ptr<Model>> cache;
    char key[160];
    std::snprintf(key, sizeof key, "%llu/%llu/%u/%u", (unsigned long long)c.dim, (unsigned long long)c.seed,
                  c.n_topics, c.topic_terms);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto m = std::make_shared<Model>();
    m->dim = c.dim;
    m->seed = c.seed;
    m->n_topics = c.n_topics;
    m->topic_terms = (uint32_t)std::min<uint64_t>(c.topic_terms, c.dim);
    Rng r(c.seed, 0xA11CE);
    m->rank_to_term.resize(c.dim);
    for (uint64_t i = 0; i < c.dim; ++i) m->rank_to_term[i] = (uint32_t)i;
    for (uint64_t i = c.dim; i > 1; --i) std::swap(m->rank_to_term[i - 1], m->rank_to_term[r.below(i)]);
    std::vector<double> w(c.dim);
    for (uint64_t i = 0; i < c.dim; ++i) w[i] = std::pow((double)(i + 1), -1.1);
    m->global.build(w);
    std::vector<double> wp(m->topic_terms);
    for (uint32_t i = 0; i < m->topic_terms; ++i) wp[i] = 1.0 / (double)(i + 1);
    m->position.build(wp);
    // topic term sets: flatter popularity (rank^-0.6) so that topic cores are not all the same head terms
    Alias flat;
    for (uint64_t i = 0; i < c.dim; ++i) w[i] = std::pow((double)(i + 1), -0.6);
    flat.build(w);
    m->topic_table.resize((size_t)m->n_topics * m->topic_terms);
    std::vector<uint32_t> stamp(c.dim, 0);
    for (uint32_t t = 0; t < m->n_topics; ++t) {
        Rng rt(c.seed, 0x70000000ull + t);
        uint32_t got = 0;
        uint32_t* dst = &m->topic_table[(size_t)t * m->topic_terms];
        while (got < m->topic_terms) {
            uint32_t term = m->rank_to_term[flat.sample(rt)];
            if (stamp[term] == t + 1) continue;
            stamp[term] = t + 1;
            dst[got++] = term;
        }
    }
    m->lognorm.resize(65536);
    for (int i = 0; i < 65536; ++i) {
        // inverse normal CDF (Acklam) at (i+0.5)/65536
        double p = (i + 0.5) / 65536.0, x;
        static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                   1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
        static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                   6.680131188771972e+01, -1.328068155288572e+01};
        static const double cc[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                    -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
        static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                   3.754408661907416e+00};
        if (p < 0.02425) {
            double q = std::sqrt(-2 * std::log(p));
            x = (((((cc[0] * q + cc[1]) * q + cc[2]) * q + cc[3]) * q + cc[4]) * q + cc[5]) /
                ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
        } else if (p > 1 - 0.02425) {
            double q = std::sqrt(-2 * std::log(1 - p));
            x = -(((((cc[0] * q + cc[1]) * q + cc[2]) * q + cc[3]) * q + cc[4]) * q + cc[5]) /
                ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
        } else {
            double q = p - 0.5, rr = q * q;
            x = (((((a[0] * rr + a[1]) * rr + a[2]) * rr + a[3]) * rr + a[4]) * rr + a[5]) * q /
                (((((b[0] * rr + b[1]) * rr + b[2]) * rr + b[3]) * rr + b[4]) * rr + 1);
        }
        m->lognorm[i] = (float)x;  // store the normal quantile; exp applied with (mu, sigma) at use
    }
    cache[key] = m;
    return m;
}

struct DocGen {
    const Model& m;
    const ShostSynthConfig& c;
    std::vector<uint32_t> stamp;  // per-thread dedupe (dim)
    uint32_t epoch = 0;
    DocGen(const Model& mm, const ShostSynthConfig& cc) : m(mm), c(cc), stamp(mm.dim, 0) {}

    struct Item {
        uint32_t comp;
        float val;
    };
    struct Topics {
        uint32_t n;
        uint32_t id[3];
        float w[3];
    };

    inline float normal_q(Rng& r) const { return m.lognorm[r.next() >> 48]; }

    Topics doc_topics(Rng& r) const {
        Topics t;
        double u = r.uniform();
        t.n = u < 0.5 ? 1 : (u < 0.85 ? 2 : 3);
        double s = 0;
        for (uint32_t i = 0; i < t.n; ++i) {
            t.id[i] = (uint32_t)r.below(m.n_topics);
            // Dirichlet(0.3)-like: gamma(0.3) ~ u^(1/0.3) * exp-ish; a cheap skewed positive weight
            double g = std::pow(r.uniform(), 1.0 / 0.3) + 1e-3;
            t.w[i] = (float)g;
            s += g;
        }
        for (uint32_t i = 0; i < t.n; ++i) t.w[i] = (float)(t.w[i] / s);
        return t;
    }

    // one topic-or-background term draw; returns component and its value
    inline Item draw(Rng& r, const Topics& t) const {
        Item it;
        if (r.uniform() < 0.85) {
            double u = r.uniform();
            uint32_t ti = 0;
            float acc = t.w[0];
            while (ti + 1 < t.n && u > acc) acc += t.w[++ti];
            uint32_t pos = m.position.sample(r);
            it.comp = m.topic_table[(size_t)t.id[ti] * m.topic_terms + pos];
            float boost = 1.0f + 3.0f / (1.0f + (float)pos * 0.125f);
            float base = std::exp(-1.9f + 1.10f * normal_q(r));
            it.val = base * boost * (0.6f + 0.8f * t.w[ti]);
        } else {
            it.comp = m.rank_to_term[m.global.sample(r)];
            it.val = std::exp(-2.2f + 1.0f * normal_q(r));
        }
        it.val = std::min(3.5f, std::max(1e-3f, it.val));
        return it;
    }

    void gen_doc(uint64_t doc_id, std::vector<Item>& out, Topics* topics_out = nullptr) {
        Rng r(c.seed, doc_id * 2 + 1);
        Topics t = doc_topics(r);
        if (topics_out) *topics_out = t;
        double ln = std::log((double)c.doc_nnz_mean) + c.doc_nnz_sigma * normal_q(r);
        int64_t nnz = (int64_t)std::llround(std::exp(ln));
        nnz = std::min<int64_t>(400, std::max<int64_t>(16, nnz));
        nnz = std::min<int64_t>(nnz, (int64_t)m.dim);
        out.clear();
        if (++epoch == 0) {
            std::fill(stamp.begin(), stamp.end(), 0);
            epoch = 1;
        }
        int attempts = 0;
        while ((int64_t)out.size() < nnz && attempts < nnz * 12) {
            ++attempts;
            Item it = draw(r, t);
            if (stamp[it.comp] == epoch) continue;
            stamp[it.comp] = epoch;
            out.push_back(it);
        }
        std::sort(out.begin(), out.end(), [](const Item& a, const Item& b) { return a.comp < b.comp; });
    }

    void gen_query(uint64_t qid, std::vector<Item>& out, std::vector<Item>& doc_tmp) {
        Rng r(c.seed ^ 0x5EED5EED5EEDull, qid * 2 + 1);
        uint64_t d = r.below(c.n_docs);
        Topics t;
        gen_doc(d, doc_tmp, &t);
        double ln = std::log((double)c.query_nnz_mean) + c.query_nnz_sigma * normal_q(r);
        int64_t nnz = (int64_t)std::llround(std::exp(ln));
        nnz = std::min<int64_t>(128, std::max<int64_t>(4, nnz));
        nnz = std::min<int64_t>(nnz, (int64_t)m.dim);
        // half of the query: sampled from the source doc's top-weighted terms
        std::sort(doc_tmp.begin(), doc_tmp.end(), [](const Item& a, const Item& b) { return a.val > b.val; });
        if (++epoch == 0) {
            std::fill(stamp.begin(), stamp.end(), 0);
            epoch = 1;
        }
        out.clear();
        int64_t from_doc = std::min<int64_t>((nnz + 1) / 2, (int64_t)doc_tmp.size());
        int64_t pool = std::min<int64_t>((int64_t)doc_tmp.size(), from_doc * 2);
        for (int64_t i = 0; i < pool && (int64_t)out.size() < from_doc; ++i) {
            // keep each of the top `pool` terms with probability 1/2, best first
            if (r.uniform() < 0.5 || pool - i <= from_doc - (int64_t)out.size()) {
                Item it = doc_tmp[i];
                float qv = std::exp(-1.34f + 1.39f * normal_q(r));
                it.val = std::min(3.5f, std::max(1e-3f, qv * (0.4f + 0.9f * doc_tmp[i].val)));
                stamp[it.comp] = epoch;
                out.push_back(it);
            }
        }
        int attempts = 0;
        while ((int64_t)out.size() < nnz && attempts < nnz * 12) {
            ++attempts;
            Item it = draw(r, t);
            if (stamp[it.comp] == epoch) continue;
            stamp[it.comp] = epoch;
            it.val = std::min(3.5f, std::max(1e-3f, std::exp(-1.6f + 1.2f * normal_q(r))));
            out.push_back(it);
        }
        // boost the three largest terms so that they carry ~25-35 % of the L1 mass (toy queries: 0.23-0.36)
        std::sort(out.begin(), out.end(), [](const Item& a, const Item& b) { return a.val > b.val; });
        double l1 = 0;
        for (auto& x : out) l1 += x.val;
        double top3 = 0;
        size_t n3 = std::min<size_t>(3, out.size());
        for (size_t i = 0; i < n3; ++i) top3 += out[i].val;
        double target = 0.25 + 0.10 * r.uniform();
        if (top3 < target * l1 && l1 > top3) {
            double f = target * (l1 - top3) / ((1 - target) * top3);
            for (size_t i = 0; i < n3; ++i) out[i].val = std::min(3.5f, (float)(out[i].val * f));
        }
        std::sort(out.begin(), out.end(), [](const Item& a, const Item& b) { return a.comp < b.comp; });
    }
};

int generate(const ShostSynthConfig& cfg, uint64_t n, bool queries, ShostDataset** out) {
    if (cfg.dim == 0 || cfg.n_docs == 0 || cfg.n_topics == 0 || cfg.topic_terms == 0) {
        set_error("synth: dim, n_docs, n_topics, topic_terms must be > 0");
        return SGPU_EINVAL;
    }
    auto model = get_model(cfg);
    const unsigned T = hw_threads(cfg.n_threads);
    auto* ds = new ShostDataset();
    ds->n_vecs = n;
    ds->dim = cfg.dim;
    ds->offsets.assign(n + 1, 0);
    // pass 1: sizes ; pass 2: fill (generation is deterministic, so it is simply run twice in chunks
    // small enough to keep the per-chunk vectors: we buffer per chunk instead of regenerating)
    const uint64_t chunk = 4096;
    const uint64_t n_chunks = (n + chunk - 1) / chunk;
    std::vector<std::vector<uint32_t>> cc(n_chunks);
    std::vector<std::vector<float>> cv(n_chunks);
    parallel_for(n_chunks, 1, T, [&](uint64_t b, uint64_t e, unsigned) {
        DocGen g(*model, cfg);
        std::vector<DocGen::Item> items, tmp;
        for (uint64_t ch = b; ch < e; ++ch) {
            uint64_t lo = ch * chunk, hi = std::min(n, lo + chunk);
            for (uint64_t i = lo; i < hi; ++i) {
                if (queries) g.gen_query(i, items, tmp);
                else g.gen_doc(i, items);
                ds->offsets[i + 1] = items.size();
                for (auto& it : items) {
                    cc[ch].push_back(it.comp);
                    cv[ch].push_back(it.val);
                }
            }
        }
    });
    for (uint64_t i = 0; i < n; ++i) ds->offsets[i + 1] += ds->offsets[i];
    ds->comps.resize(ds->offsets[n]);
    ds->values.resize(ds->offsets[n]);
    parallel_for(n_chunks, 1, T, [&](uint64_t b, uint64_t e, unsigned) {
        for (uint64_t ch = b; ch < e; ++ch) {
            uint64_t at = ds->offsets[ch * chunk];
            if (!cc[ch].empty()) {
                std::memcpy(&ds->comps[at], cc[ch].data(), cc[ch].size() * 4);
                std::memcpy(&ds->values[at], cv[ch].data(), cv[ch].size() * 4);
            }
            std::vector<uint32_t>().swap(cc[ch]);
            std::vector<float>().swap(cv[ch]);
        }
    });
    *out = ds;
    return SGPU_OK;
}

}  // namespace

int synth_documents(const ShostSynthConfig& cfg, ShostDataset** out) { return generate(cfg, cfg.n_docs, false, out); }
int synth_queries(const ShostSynthConfig& cfg, uint64_t n_queries, ShostDataset** out) {
    return generate(cfg, n_queries, true, out);
}

}  // namespace shost
