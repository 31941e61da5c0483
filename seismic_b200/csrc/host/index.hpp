// Host-side containers: sparse dataset and the logical Seismic index
// (field-for-field the reference structs, reference src/inverted_index.rs:38-52,
// src/posting_list.rs:68-73, src/quantized_summary.rs:14-24; per-list arrays concatenated).
#pragma once
#include "common.hpp"

struct ShostDataset {
    uint64_t n_vecs = 0, dim = 0;
    std::vector<uint64_t> offsets;  // n_vecs+1
    std::vector<uint32_t> comps;
    std::vector<float> values;
};

// One typed array that is either owned (vector<uint8_t>) or borrowed from an mmap.
struct Section {
    std::vector<uint8_t> own;
    const uint8_t* ptr = nullptr;
    uint64_t bytes = 0;
    template <class T>
    const T* as() const { return reinterpret_cast<const T*>(ptr); }
    template <class T>
    uint64_t count() const { return bytes / sizeof(T); }
    template <class T>
    void adopt(const std::vector<T>& v) {
        own.resize(v.size() * sizeof(T));
        if (!v.empty()) std::memcpy(own.data(), v.data(), own.size());
        ptr = own.data();
        bytes = own.size();
    }
    template <class T>
    T* alloc(uint64_t n) {
        own.assign(n * sizeof(T), 0);
        ptr = own.data();
        bytes = own.size();
        return reinterpret_cast<T*>(own.data());
    }
};

enum SectionId {
    SEC_FWD_OFFSETS = 0,
    SEC_FWD_COMPS,
    SEC_FWD_VALUES,
    SEC_FWD_NNZ,
    SEC_LIST_POST_START,
    SEC_POSTINGS,
    SEC_LIST_BLK_START,
    SEC_BLK_POST_OFF,
    SEC_BLK_MIN,
    SEC_BLK_QUANT,
    SEC_LIST_SC_START,
    SEC_SC_COMP,
    SEC_LIST_ENT_START,
    SEC_SC_RUN_OFF,
    SEC_ENT_BLK,
    SEC_ENT_CODE,
    SEC_COUNT
};

struct ShostIndex {
    uint32_t comp_bits = 16;
    uint32_t value_kind = SGPU_VAL_F16;
    uint64_t n_docs = 0, dim = 0, nnz = 0;
    float value_scale = 1.f;
    ShostBuildConfig config{};
    Section sec[SEC_COUNT];
    // mmap backing (load)
    void* map_base = nullptr;
    uint64_t map_len = 0;
    ~ShostIndex();
};

namespace shost {
int build_index(const ShostDataset& ds, const ShostBuildConfig& cfg, ShostIndex** out);
int convert_dotvbyte(const ShostIndex& in, ShostIndex** out);
int save_index(const ShostIndex& idx, const char* path);
int load_index(const char* path, ShostIndex** out);
void fill_view(const ShostIndex& idx, SgpuIndexView* v);
int read_bin(const char* path, ShostDataset** out);
int write_bin(const ShostDataset& ds, const char* path);
int synth_documents(const ShostSynthConfig& cfg, ShostDataset** out);
int synth_queries(const ShostSynthConfig& cfg, uint64_t n_queries, ShostDataset** out);
// decode value `i` of a plain forward index to f32
float decode_value(uint32_t value_kind, float scale, const void* values, uint64_t i);
}  // namespace shost
