// extern "C" surface of the host side (shost_*) + the shared thread-local error string.
#include "index.hpp"

namespace shost {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }
}  // namespace shost

using namespace shost;

extern "C" {

const char* sgpu_last_error(void) { return shost::last_error(); }

void shost_default_config(ShostBuildConfig* c) {
    std::memset(c, 0, sizeof(*c));
    c->pruning = 0;
    c->n_postings = 3500;
    c->max_fraction = 1.5f;
    c->blocking = 0;
    c->centroid_fraction = 0.1f;
    c->min_cluster_size = 2;
    c->doc_cut = 15;
    c->block_size = 10;
    c->summarization = 0;
    c->summary_energy = 0.4f;
    c->n_components = 50;
    c->comp_bits = 16;
    c->value_kind = SGPU_VAL_F16;
    c->n_threads = 0;
    c->kmeans_seed = 1142;
}

void shost_default_synth(ShostSynthConfig* c) {
    std::memset(c, 0, sizeof(*c));
    c->n_docs = 100000;
    c->dim = 30522;
    c->seed = 20260517;
    c->n_topics = 4096;
    c->topic_terms = 2000;
    c->doc_nnz_mean = 115.f;
    c->doc_nnz_sigma = 0.30f;
    c->query_nnz_mean = 38.f;
    c->query_nnz_sigma = 0.35f;
    c->n_threads = 0;
}

int shost_dataset_create(uint64_t n_vecs, uint64_t dim, const uint64_t* offsets, const uint32_t* comps,
                         const float* values, ShostDataset** out) {
    if (!out || !offsets || (offsets[n_vecs] && (!comps || !values))) {
        set_error("shost_dataset_create: null argument");
        return SGPU_EINVAL;
    }
    auto* ds = new ShostDataset();
    ds->n_vecs = n_vecs;
    ds->dim = dim;
    ds->offsets.assign(offsets, offsets + n_vecs + 1);
    ds->comps.assign(comps, comps + offsets[n_vecs]);
    ds->values.assign(values, values + offsets[n_vecs]);
    *out = ds;
    return SGPU_OK;
}
int shost_dataset_read_bin(const char* path, ShostDataset** out) { return read_bin(path, out); }
int shost_dataset_write_bin(const ShostDataset* ds, const char* path) { return write_bin(*ds, path); }
void shost_dataset_destroy(ShostDataset* ds) { delete ds; }
uint64_t shost_dataset_len(const ShostDataset* ds) { return ds->n_vecs; }
uint64_t shost_dataset_dim(const ShostDataset* ds) { return ds->dim; }
uint64_t shost_dataset_nnz(const ShostDataset* ds) { return ds->offsets.empty() ? 0 : ds->offsets.back(); }
const uint64_t* shost_dataset_offsets(const ShostDataset* ds) { return ds->offsets.data(); }
const uint32_t* shost_dataset_comps(const ShostDataset* ds) { return ds->comps.data(); }
const float* shost_dataset_values(const ShostDataset* ds) { return ds->values.data(); }

int shost_synth_documents(const ShostSynthConfig* cfg, ShostDataset** out) { return synth_documents(*cfg, out); }
int shost_synth_queries(const ShostSynthConfig* cfg, uint64_t n, ShostDataset** out) {
    return synth_queries(*cfg, n, out);
}

int shost_index_build(const ShostDataset* ds, const ShostBuildConfig* cfg, ShostIndex** out) {
    if (!ds || !cfg || !out) {
        set_error("shost_index_build: null argument");
        return SGPU_EINVAL;
    }
    try {
        return build_index(*ds, *cfg, out);
    } catch (const std::bad_alloc&) {
        set_error("out of host memory while building");
        return SGPU_ENOMEM;
    } catch (const std::exception& e) {
        set_error(e.what());
        return SGPU_EINVAL;
    }
}
int shost_index_convert_dotvbyte(const ShostIndex* idx, ShostIndex** out) {
    if (!idx || !out) {
        set_error("shost_index_convert_dotvbyte: null argument");
        return SGPU_EINVAL;
    }
    return convert_dotvbyte(*idx, out);
}
int shost_index_save(const ShostIndex* idx, const char* path) { return save_index(*idx, path); }
int shost_index_load(const char* path, ShostIndex** out) {
    try {
        return load_index(path, out);
    } catch (const std::bad_alloc&) {
        set_error("out of host memory");
        return SGPU_ENOMEM;
    }
}
void shost_index_destroy(ShostIndex* idx) { delete idx; }
int shost_index_view(const ShostIndex* idx, SgpuIndexView* out) {
    fill_view(*idx, out);
    return SGPU_OK;
}
uint64_t shost_index_nnz(const ShostIndex* idx) { return idx->nnz; }

int shost_index_space_usage(const ShostIndex* idx, uint64_t bytes[6]) {
    // sizes as the reference would hold them (component type C, not our u32 host copy)
    uint64_t cb = idx->comp_bits / 8;
    uint64_t fwd = idx->sec[SEC_FWD_OFFSETS].bytes + idx->sec[SEC_FWD_COMPS].bytes + idx->sec[SEC_FWD_VALUES].bytes;
    uint64_t post = idx->sec[SEC_POSTINGS].bytes;
    uint64_t blk = idx->sec[SEC_BLK_POST_OFF].count<uint32_t>() * 8;  // Box<[usize]>
    uint64_t summ = idx->sec[SEC_SC_COMP].count<uint32_t>() * cb + idx->sec[SEC_SC_RUN_OFF].bytes +
                    idx->sec[SEC_ENT_BLK].bytes + idx->sec[SEC_ENT_CODE].bytes + idx->sec[SEC_BLK_MIN].bytes +
                    idx->sec[SEC_BLK_QUANT].bytes;
    bytes[0] = fwd;
    bytes[1] = post;
    bytes[2] = blk;
    bytes[3] = summ;
    bytes[4] = 0;
    bytes[5] = fwd + post + blk + summ;
    return SGPU_OK;
}

int shost_index_get_doc(const ShostIndex* idx, uint64_t id, uint32_t* comps, float* values, uint32_t cap,
                        uint32_t* nnz) {
    if (id >= idx->n_docs) {
        set_error("doc id out of range");
        return SGPU_EINVAL;
    }
    if (idx->value_kind == SGPU_VAL_DOTVBYTE) {  // decode of the packed record (format: build.cpp, convert_dotvbyte)
        const uint64_t* boff = idx->sec[SEC_FWD_OFFSETS].as<uint64_t>();
        const uint8_t* rec = idx->sec[SEC_FWD_VALUES].as<uint8_t>() + boff[id];
        const uint32_t n = idx->sec[SEC_FWD_NNZ].as<uint16_t>()[id], nch = (n + 7) >> 3, ndir = (nch + 63) >> 6;
        const uint8_t* fixed = rec + 16ull * ndir;
        const uint8_t* wide = fixed + 16ull * nch;
        *nnz = n;
        uint32_t c = 0, n_wide = 0;
        for (uint32_t m = 0; m < nch; ++m) {
            const uint8_t* fx = fixed + 16ull * m;
            uint64_t mask;
            std::memcpy(&mask, rec + 16ull * (m >> 6), 8);
            const uint8_t* hi = ((mask >> (m & 63)) & 1u) ? wide + 8ull * n_wide++ : nullptr;
            for (uint32_t f = 0; f < 8; ++f) {
                c += (uint32_t)fx[f] | (hi ? (uint32_t)hi[f] << 8 : 0u);
                const uint32_t i = m * 8 + f;
                if (i < n && i < cap) comps[i] = c, values[i] = (float)fx[8 + f] * idx->value_scale;
            }
        }
        return SGPU_OK;
    }
    const uint64_t* off = idx->sec[SEC_FWD_OFFSETS].as<uint64_t>();
    uint64_t b = off[id], e = off[id + 1];
    *nnz = (uint32_t)(e - b);
    for (uint64_t i = b; i < e && i - b < cap; ++i) {
        comps[i - b] = idx->comp_bits == 16 ? (uint32_t)idx->sec[SEC_FWD_COMPS].as<uint16_t>()[i]
                                            : idx->sec[SEC_FWD_COMPS].as<uint32_t>()[i];
        values[i - b] = decode_value(idx->value_kind, idx->value_scale, idx->sec[SEC_FWD_VALUES].ptr, i);
    }
    return SGPU_OK;
}

}  // extern "C"
