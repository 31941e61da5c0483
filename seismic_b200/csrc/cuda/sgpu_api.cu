// sgpu_* C ABI: HBM image construction and the batched search driver (see include/seismic_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/seismic_b200.h"
#include "exact.cuh"
#include "kernels.cuh"
#include "search_kernels.cuh"

namespace shost {
void set_error(const std::string& msg);
}

namespace {

using namespace sgpu;

#define CK(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            shost::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
            return _e == cudaErrorMemoryAllocation ? SGPU_ENOMEM : SGPU_ECUDA;                           \
        }                                                                                                \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t ensure(size_t n) {  // grow-only
        if (n <= bytes) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, n ? n : 1);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~PinnedBuf() {
        if (p) cudaFreeHost(p);
    }
    cudaError_t ensure(size_t n) {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMallocHost(&p, n ? n : 1);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// true for cudaMallocHost / cudaHostRegister memory (the DMA engine can read or write it directly)
static bool is_page_locked(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

}  // namespace

struct SgpuIndex {
    int device = 0;
    int n_sm = 0;
    size_t smem_optin = 0, smem_per_sm = 0;
    cudaStream_t stream = nullptr;      // stream all work is enqueued on
    cudaStream_t own_stream = nullptr;  // the library's private stream
    cudaEvent_t ev[8] = {};
    // image
    DevBuf lists, postings, blk_post_off, blk_min, blk_quant, sc_comp, sc_run_off, sc_skip, ent_blk, ent_code, fwd, rec_start, knn_posts;
    sgpu::DevIndex ix{};
    uint64_t image_bytes = 0;
    uint32_t max_blocks = 0;      // largest number of blocks of any list
    uint32_t max_block_docs = 0;  // largest block
    uint32_t max_list_post = 0;   // longest posting list
    // options
    uint32_t wave_docs = 2048, first_wave_docs = 256;       // dense-query kernel (1024 threads)
    uint32_t hq_wave_docs = 768, hq_first_wave_docs = 128;  // compact-query kernel
    int hq_enabled = 1, hq_ctas_per_sm = 0;
    int hq_mode = 1;  // compact query: 1 byte index, 2 perfect hash, 3 bitmap + rank
    int hq_cand_cap = 256;    // candidate blocks per wave of the compact kernel
    int hq_carveout_pct = 0;  // shared-memory carveout of the compact kernel in % of the SM maximum (0: smallest that fits)
    int bucket = 1;   // score the documents of a wave longest first (uniform rounds per warp)
    int tma = 0;      // u16/f16 layout, byte-index query: stage the records with TMA bulk copies (3 CTAs / SM)
    int order_warp = 1;  // first-list block order by one warp per query in registers (0: CTA-wide shared-memory sort)
    int wide_heap = 1;  // 32 < k <= 128: register heap (WideHeap) on the layouts that instantiate it (0: SmemHeap)
    int occ16 = 4;      // u16 / f16 layout, byte-index query: kernel build (4 = two documents per group; 41, 51 = one)
    int occvb = 4;      // DotVByte layout: same choice
    int occ32 = 4;      // u32 / f16 layout (SeismicIndexLV): CTAs per SM the kernel's register budget is compiled for (4, 3, 2)
    int ctas = 0;
    uint64_t scratch_bytes = 1ull << 30;
    // per-batch scratch (grow-only)
    DevBuf d_qoff, d_qcomps, d_qvals, d_nterms, d_status, d_counters, d_terms, d_est, d_order, d_keys, d_stats;
    DevBuf d_out_ids, d_out_scores, d_out_counts, d_hmult, d_qlist, d_prep;
    PinnedBuf h_in, h_out, h_ctl;
    ~SgpuIndex() {
        cudaSetDevice(device);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        if (own_stream) cudaStreamDestroy(own_stream);
    }
};

namespace {

template <class T>
int upload(DevBuf& dst, const T* src, size_t n, cudaStream_t st, uint64_t* total) {
    CK(dst.ensure(n * sizeof(T)));
    if (n) CK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
    *total += n * sizeof(T);
    return SGPU_OK;
}

// Upload a kNN graph (host array of doc ids) and turn it into postings of the record image.
int set_knn_impl(SgpuIndex* ix, const uint64_t* neighbours, uint32_t knn_dim) {
    if (!ix) return SGPU_EINVAL;
    CK(cudaSetDevice(ix->device));
    if (!neighbours || knn_dim == 0) {
        ix->ix.knn_posts = nullptr;
        ix->ix.knn_dim = 0;
        return SGPU_OK;
    }
    if (ix->ix.vbyte) {
        shost::set_error("kNN graphs are not available for DotVByte indexes (as in the reference)");
        return SGPU_EUNSUPPORTED;
    }
    const uint64_t n = ix->ix.n_docs * (uint64_t)knn_dim;
    cudaStream_t st = ix->stream;
    DevBuf d_ids;
    CK(d_ids.ensure(n * 8));
    CK(ix->knn_posts.ensure(n * 8));
    CK(cudaMemcpyAsync(d_ids.p, neighbours, n * 8, cudaMemcpyHostToDevice, st));
    k_knn_posts<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_ids.as<uint64_t>(), n, ix->ix.rec_start, ix->ix.n_docs,
                                                            ix->ix.rec_chunk_units, ix->knn_posts.as<uint64_t>());
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ix->ix.knn_posts = ix->knn_posts.as<uint64_t>();
    ix->ix.knn_dim = knn_dim;
    return SGPU_OK;
}

int create_impl(const SgpuIndexView* v, int device, SgpuIndex** out) {
    if (!v || !out) {
        shost::set_error("sgpu_index_create: null argument");
        return SGPU_EINVAL;
    }
    const bool vbyte = v->value_kind == SGPU_VAL_DOTVBYTE;
    if ((v->comp_bits != 16 && v->comp_bits != 32) || v->value_kind > SGPU_VAL_DOTVBYTE ||
        (v->comp_bits == 32 && vbyte) || (vbyte && !v->fwd_nnz)) {
        shost::set_error("sgpu_index_create: supported forward indexes are u16 / u32 components with f16 / bf16 / f32 / "
                         "fixedu8 / fixedu16 values and u16 components with DotVByte values");
        return SGPU_EUNSUPPORTED;
    }
    // record geometry of the plain layouts: chunk = 8 components + 8 values; unit of rec_start / posting starts
    const uint32_t kind = v->value_kind;
    const uint32_t val_bytes = kind == SGPU_VAL_F32 ? 4 : (kind == SGPU_VAL_FIXEDU8 ? 1 : 2);
    const uint32_t chunk_bytes = 8 * ((v->comp_bits == 32 ? 4 : 2) + val_bytes);
    const uint32_t unit_bytes = chunk_bytes % 32 == 0 ? 32 : (chunk_bytes % 16 == 0 ? 16 : 8);
    const uint32_t chunk_units = chunk_bytes / unit_bytes;
    const bool fast_pack = kind == SGPU_VAL_F16;  // dedicated pack kernels for the two benchmark layouts
    const bool comp32 = v->comp_bits == 32;
    if (v->dim == 0 || (!comp32 && v->dim > 65536) || v->dim > (1u << 20)) {
        shost::set_error("sgpu_index_create: dim must be in [1, 65536] for u16 and [1, 2^20] for u32 components");
        return SGPU_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        shost::set_error("no CUDA device available (seismic_b200 has no CPU fallback)");
        return SGPU_ECUDA;
    }
    if (device < 0 || device >= ndev) {
        shost::set_error("sgpu_index_create: bad device ordinal");
        return SGPU_EINVAL;
    }
    CK(cudaSetDevice(device));
    std::unique_ptr<SgpuIndex> ix(new SgpuIndex());
    ix->device = device;
    CK(cudaDeviceGetAttribute(&ix->n_sm, cudaDevAttrMultiProcessorCount, device));
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    ix->smem_optin = (size_t)optin;
    int per_sm = 0;
    CK(cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    ix->smem_per_sm = (size_t)per_sm;
    ix->ctas = ix->n_sm;
    if (comp32 && kind == SGPU_VAL_F16) {
        // SeismicIndexLV: one document in flight per 8-lane group (48 registers, no spills), 5 CTAs per SM with
        // slightly smaller waves (measured at 1 M docs, vocabulary 200 k, k = 100: 18.7 -> 11.4 ms per 10 k queries)
        ix->occ32 = 51;
        ix->hq_wave_docs = 640;
        ix->hq_cand_cap = 128;
    }
    CK(cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking));
    ix->stream = ix->own_stream;
    for (auto& e : ix->ev) CK(cudaEventCreate(&e));
    cudaStream_t st = ix->stream;
    const uint64_t N = v->n_docs, dim = v->dim;
    uint64_t total = 0;

    // ---- record layout
    std::vector<uint32_t> rec_start(N + 1);
    if (vbyte) {  // the packed byte stream is the record buffer; rec_start in 16-byte units
        const uint64_t bytes = v->fwd_offsets[N];
        if ((bytes >> 4) >= (1ull << 32)) {
            shost::set_error("forward index larger than 2^32 record units");
            return SGPU_EUNSUPPORTED;
        }
        for (uint64_t d = 0; d <= N; ++d) {
            if (v->fwd_offsets[d] & 15) {
                shost::set_error("DotVByte records must be 16-byte aligned");
                return SGPU_EINVAL;
            }
            rec_start[d] = (uint32_t)(v->fwd_offsets[d] >> 4);
        }
        CK(ix->fwd.ensure(bytes + 64));
        CK(cudaMemsetAsync((char*)ix->fwd.p + bytes, 0, 64, st));
        if (bytes) CK(cudaMemcpyAsync(ix->fwd.p, v->fwd_values, bytes, cudaMemcpyHostToDevice, st));
        total += bytes;
    } else {
        uint64_t units = 0;
        for (uint64_t d = 0; d < N; ++d) {
            rec_start[d] = (uint32_t)units;
            uint64_t len = v->fwd_offsets[d + 1] - v->fwd_offsets[d];
            if (len > 65535) {
                shost::set_error("document longer than 65535 components");
                return SGPU_EINVAL;
            }
            units += ((len + 7) >> 3) * chunk_units;
            if (units >= (1ull << 32)) {
                shost::set_error("forward index larger than 2^32 record units");
                return SGPU_EUNSUPPORTED;
            }
        }
        rec_start[N] = (uint32_t)units;
        CK(ix->fwd.ensure(std::max<uint64_t>(units, 1) * unit_bytes + 64));
        total += units * unit_bytes;
    }
    if (int rc = upload(ix->rec_start, rec_start.data(), N + 1, st, &total)) return rc;
    DevBuf d_fwd_off;
    uint64_t scratch_total = 0;
    if (!vbyte)
        if (int rc = upload(d_fwd_off, v->fwd_offsets, N + 1, st, &scratch_total)) return rc;
    if (!vbyte) {
        // pack records on the GPU, ~64 M elements per slice
        const uint64_t slice_elems = 64ull << 20;
        DevBuf d_c, d_v;
        uint64_t d0 = 0;
        while (d0 < N) {
            uint64_t d1 = d0, e0 = v->fwd_offsets[d0];
            while (d1 < N && v->fwd_offsets[d1 + 1] - e0 <= slice_elems) ++d1;
            if (d1 == d0) d1 = d0 + 1;
            const uint64_t ne = v->fwd_offsets[d1] - e0;
            CK(d_c.ensure(std::max<uint64_t>(ne, 1) * (comp32 ? 4 : 2)));
            CK(d_v.ensure(std::max<uint64_t>(ne, 1) * val_bytes));
            if (ne) {
                if (comp32) CK(cudaMemcpyAsync(d_c.p, (const uint32_t*)v->fwd_comps + e0, ne * 4, cudaMemcpyHostToDevice, st));
                else CK(cudaMemcpyAsync(d_c.p, (const uint16_t*)v->fwd_comps + e0, ne * 2, cudaMemcpyHostToDevice, st));
                CK(cudaMemcpyAsync(d_v.p, (const uint8_t*)v->fwd_values + e0 * val_bytes, ne * val_bytes,
                                   cudaMemcpyHostToDevice, st));
            }
            const uint64_t nd = d1 - d0;
            const unsigned blocks = (unsigned)((nd + 7) / 8);
            if (!fast_pack)
                k_pack_records_any<<<blocks, 256, 0, st>>>(d_fwd_off.as<uint64_t>(), d_c.as<uint8_t>(), d_v.as<uint8_t>(),
                                                           ix->rec_start.as<uint32_t>(), d0, nd, e0, ix->fwd.as<uint8_t>(),
                                                           comp32 ? 4u : 2u, val_bytes, unit_bytes);
            else if (comp32)
                k_pack_records32<<<blocks, 256, 0, st>>>(d_fwd_off.as<uint64_t>(), d_c.as<uint32_t>(), d_v.as<uint16_t>(),
                                                         ix->rec_start.as<uint32_t>(), d0, nd, e0, ix->fwd.as<uint32_t>());
            else
                k_pack_records<<<blocks, 256, 0, st>>>(d_fwd_off.as<uint64_t>(), d_c.as<uint16_t>(), d_v.as<uint16_t>(),
                                                       ix->rec_start.as<uint32_t>(), d0, nd, e0, ix->fwd.as<uint16_t>());
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));
            d0 = d1;
        }
    }
    // ---- posting lists
    const uint64_t P = v->list_post_start[dim], TB = v->list_blk_start[dim], TSC = v->list_sc_start[dim],
                   TE = v->list_ent_start[dim];
    std::vector<ListHdr> hdr(dim);
    uint32_t max_blocks = 0, max_block_docs = 0;
    uint64_t n_skip_total = 0;
    for (uint64_t l = 0; l < dim; ++l) {
        ListHdr& h = hdr[l];
        h.post_base = v->list_post_start[l];
        h.ent_base = v->list_ent_start[l];
        h.sc_base = v->list_sc_start[l];
        h.blk_base = v->list_blk_start[l];
        h.n_blk = (uint32_t)(v->list_blk_start[l + 1] - v->list_blk_start[l]);
        h.n_sc = (uint32_t)(v->list_sc_start[l + 1] - v->list_sc_start[l]);
        h.n_post = (uint32_t)(v->list_post_start[l + 1] - v->list_post_start[l]);
        h.skip_base = (uint32_t)n_skip_total;
        n_skip_total += (h.n_sc + 31) >> 5;
        if (n_skip_total >= (1ull << 32)) {
            shost::set_error("summary component directory larger than 2^32 entries");
            return SGPU_EUNSUPPORTED;
        }
        if (h.n_blk > 65535) {
            shost::set_error("list with more than 65535 blocks");
            return SGPU_EINVAL;
        }
        max_blocks = std::max(max_blocks, h.n_blk);
        ix->max_list_post = std::max(ix->max_list_post, h.n_post);
        const uint32_t* bo = v->blk_post_off + h.blk_base + l;
        for (uint32_t b = 0; b < h.n_blk; ++b) max_block_docs = std::max(max_block_docs, bo[b + 1] - bo[b]);
    }
    ix->max_blocks = max_blocks;
    ix->max_block_docs = max_block_docs;
    if (int rc = upload(ix->lists, hdr.data(), dim, st, &total)) return rc;
    if (int rc = upload(ix->postings, v->postings, P, st, &total)) return rc;
    if (P && !vbyte) {  // DotVByte postings already carry (byte offset / 4, nnz)
        k_translate_postings<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(
            d_fwd_off.as<uint64_t>(), ix->rec_start.as<uint32_t>(), N, ix->postings.as<uint64_t>(), P);
        CK(cudaGetLastError());
    }
    if (int rc = upload(ix->blk_post_off, v->blk_post_off, TB + dim, st, &total)) return rc;
    if (int rc = upload(ix->blk_min, v->blk_min, TB, st, &total)) return rc;
    if (int rc = upload(ix->blk_quant, v->blk_quant, TB, st, &total)) return rc;
    if (int rc = upload(ix->sc_comp, v->sc_comp, TSC, st, &total)) return rc;
    if (int rc = upload(ix->sc_run_off, v->sc_run_off, TSC + dim, st, &total)) return rc;
    CK(ix->sc_skip.ensure(std::max<uint64_t>(n_skip_total, 1) * 4));
    total += n_skip_total * 4;
    k_build_skip<<<(unsigned)((dim + 7) / 8), 256, 0, st>>>(ix->lists.as<ListHdr>(), (uint32_t)dim, ix->sc_comp.as<uint32_t>(),
                                                           ix->sc_skip.as<uint32_t>());
    CK(cudaGetLastError());
    if (int rc = upload(ix->ent_blk, v->ent_blk, TE, st, &total)) return rc;
    if (int rc = upload(ix->ent_code, v->ent_code, TE, st, &total)) return rc;
    CK(cudaStreamSynchronize(st));
    ix->image_bytes = total;
    DevIndex& d = ix->ix;
    d.lists = ix->lists.as<ListHdr>();
    d.postings = ix->postings.as<uint64_t>();
    d.blk_post_off = ix->blk_post_off.as<uint32_t>();
    d.blk_min = ix->blk_min.as<float>();
    d.blk_quant = ix->blk_quant.as<float>();
    d.sc_comp = ix->sc_comp.as<uint32_t>();
    d.sc_run_off = ix->sc_run_off.as<uint32_t>();
    d.sc_skip = ix->sc_skip.as<uint32_t>();
    d.ent_blk = ix->ent_blk.as<uint16_t>();
    d.ent_code = ix->ent_code.as<uint8_t>();
    d.fwd = ix->fwd.as<uint4>();
    d.rec_start = ix->rec_start.as<uint32_t>();
    d.n_docs = N;
    d.dim = (uint32_t)dim;
    d.comp32 = comp32 ? 1u : 0u;
    d.vbyte = vbyte ? 1u : 0u;
    d.value_kind = kind;
    d.value_scale = v->value_scale;
    d.knn_posts = nullptr;
    d.knn_dim = 0;
    d.rec_chunk_units = vbyte ? 1u : chunk_units;
    if (v->knn_neighbours && v->knn_dim)
        if (int rc = set_knn_impl(ix.get(), v->knn_neighbours, v->knn_dim)) return rc;
    *out = ix.release();
    return SGPU_OK;
}

// ---- batched search: everything is enqueued on the index's stream without a host round trip ------------------
// enqueue_search() launches the whole pipeline (k_prep .. k_finish); finish_search() synchronises once, reads the
// validation counters / statistics that the kernels left in device memory, and reports errors.  The only cases that
// need the host in the middle are (a) query_cut > FAST_CUT, where the scratch is sized by the largest number of terms
// any query really has, (b) batches whose scratch exceeds the budget (processed chunk by chunk) and (c) queries too
// long for the compact query table on layouts without the dense-query kernel (a second pass of the sorted-query
// kernel over just those queries).
constexpr uint32_t FAST_CUT = 16;

struct Pending {
    bool active = false;
    bool events_pending = false;  // single chunk: ev[2..6] are read in finish_search
    bool long_pass_pending = false;
    uint32_t launches = 0, ctas_per_sm = 0, nq = 0, k = 0, vkind = 0;
    float ms_terms = 0.f, ms_sum = 0.f, ms_search = 0.f, ms_fin = 0.f;
    // what the deferred long-query pass needs
    void (*k_long)(const SearchArgs) = nullptr;
    SearchArgs a_long{};
    int long_ctas = 0, long_threads = 0;
    size_t long_smem = 0;
    uint64_t* d_ids = nullptr;
};

struct HostCtl {  // pinned
    uint32_t prep[8];
    uint32_t counters[8];
    unsigned long long stats[12];
};

int collect_chunk_times(SgpuIndex* ix, Pending& pd) {
    float ms;
    CK(cudaEventElapsedTime(&ms, ix->ev[2], ix->ev[3])); pd.ms_terms += ms;
    CK(cudaEventElapsedTime(&ms, ix->ev[3], ix->ev[4])); pd.ms_sum += ms;
    CK(cudaEventElapsedTime(&ms, ix->ev[4], ix->ev[5])); pd.ms_search += ms;
    CK(cudaEventElapsedTime(&ms, ix->ev[5], ix->ev[6])); pd.ms_fin += ms;
    return SGPU_OK;
}

// second pass over the queries the compact query table cannot hold (layouts without the dense-query kernel)
int run_long_pass(SgpuIndex* ix, Pending& pd, uint32_t max_nnz, uint32_t n_chunk, uint64_t* d_ids_chunk) {
    cudaStream_t st = ix->stream;
    SearchArgs al = pd.a_long;
    al.qd_words = max_nnz;  // SortedQuery: capacity in components
    const size_t smem = pd.long_smem + (size_t)al.qd_words * 8 + 16;  // SortedQuery::bytes
    if (smem + 1024 > ix->smem_optin) {
        shost::set_error("a query has more components (" + std::to_string(max_nnz) +
                         ") than the sorted-query kernel can stage in shared memory");
        return SGPU_EUNSUPPORTED;
    }
    CK(cudaFuncSetAttribute(pd.k_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pd.k_long, pd.long_threads, smem));
    if (occ < 1) {
        shost::set_error("sorted-query kernel does not fit");
        return SGPU_EUNSUPPORTED;
    }
    pd.k_long<<<occ * ix->n_sm, pd.long_threads, smem, st>>>(al);
    CK(cudaGetLastError());
    const uint64_t tot = (uint64_t)n_chunk * pd.k;
    k_finish<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ix->ix.rec_start, ix->ix.n_docs, al.sc.out_keys, al.out_counts,
                                                            pd.k, n_chunk, d_ids_chunk);
    CK(cudaGetLastError());
    pd.launches += 2;
    return SGPU_OK;
}

// Device-resident batch search. All pointers are device pointers on ix->device.
int enqueue_search(SgpuIndex* ix, const SgpuQueryBatch* dq, const SgpuSearchParams* p, uint64_t* d_ids,
                   float* d_scores, uint32_t* d_counts, Pending& pd) {
    pd = Pending{};
    if (!ix || !dq || !p || !d_ids || !d_scores || !d_counts) {
        shost::set_error("sgpu_batch_search: null argument");
        return SGPU_EINVAL;
    }
    if (p->k == 0) {
        shost::set_error("k must be > 0 (KHeap::new asserts k > 0)");
        return SGPU_EINVAL;
    }
    if (p->k > 1024) {
        shost::set_error("k > 1024 is not supported");
        return SGPU_EUNSUPPORTED;
    }
    if (dq->n_queries >= (1ull << 31)) {
        shost::set_error("too many queries in one batch");
        return SGPU_EINVAL;
    }
    CK(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const uint32_t nq = (uint32_t)dq->n_queries;
    if (nq == 0) return SGPU_OK;
    const uint32_t k = p->k;
    pd.nq = nq;
    pd.k = k;
    pd.vkind = ix->ix.value_kind;
    pd.d_ids = d_ids;

    CK(ix->d_nterms.ensure((size_t)nq * 4));
    CK(ix->d_status.ensure((size_t)nq * 4));
    CK(ix->d_prep.ensure(32));
    CK(ix->d_counters.ensure(32));
    CK(ix->d_stats.ensure(12 * sizeof(unsigned long long)));
    CK(ix->h_ctl.ensure(sizeof(HostCtl)));
    HostCtl* hc = ix->h_ctl.as<HostCtl>();
    CK(cudaMemsetAsync(ix->d_prep.p, 0, 32, st));
    CK(cudaMemsetAsync(ix->d_stats.p, 0, 12 * sizeof(unsigned long long), st));

    // ---- launch plans: the compact-query kernel (many CTAs / SM; byte-indexed, perfect-hash or bitmap+rank query
    // table, <= 255 distinct components) and the kernel for longer queries (dense f32 query for the u16/f16 layout,
    // sorted-query binary search elsewhere)
    const size_t heap_bytes = 2 * (size_t)((k + 3) & ~3u) * 4;
    const int hk = k <= 32 ? 0 : (k <= 128 && ix->wide_heap ? 1 : 2);  // heap kind (search_kernels.cuh)
    SearchArgs ad{};
    ad.ix = ix->ix;
    ad.k = k;
    ad.heap_factor = p->heap_factor;
    ad.first_sorted = p->first_sorted ? 1 : 0;
    // no graph attached: the reference skips the refine (`if n_knn > 0 && let Some(knn)`, src/inverted_index.rs:215-217)
    ad.n_knn = ix->ix.knn_posts ? std::min(p->n_knn, ix->ix.knn_dim) : 0u;
    ad.wave_docs = std::max(1u, ix->wave_docs);
    ad.first_wave_docs = std::max(1u, ix->first_wave_docs);
    ad.buf_docs = std::max(ad.wave_docs, ad.first_wave_docs);
    ad.cand_cap = DENSE_THREADS;
    ad.qd_words = (ix->ix.dim + 31u) & ~31u;
    ad.counter_idx = 0;
    ad.bucket = ix->bucket ? 1u : 0u;
    const int ctas = std::max(1, ix->ctas);
    auto wave_bytes = [&](const SearchArgs& x) {
        return 4 * (size_t)x.cand_cap * 4 + heap_bytes + (size_t)x.buf_docs * 12 + ((x.buf_docs + 31) / 32) * 4 +
               (size_t)x.buf_docs * 2 + 16;
    };
    const size_t smem_d = (size_t)ad.qd_words * 4 + wave_bytes(ad);
    const bool comp32 = ix->ix.comp32 != 0;
    // DotVByte sums code * 2^-24 and scales once per document (search.cuh, vb_decode)
    ad.value_scale = ix->ix.vbyte ? ix->ix.value_scale * 16777216.f : ix->ix.value_scale;
    const uint32_t vkind = ix->ix.value_kind;
    const bool plain16 = !comp32 && vkind == SGPU_VAL_F16;  // the layout that also has the dense-query kernel
    const bool dense_ok = plain16 && smem_d + 1024 <= ix->smem_optin;
    kern_t kd = pick_rec16(Q_DENSE, hk);
    if (dense_ok) CK(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));

    SearchArgs ah = ad;
    ah.wave_docs = std::max(1u, ix->hq_wave_docs);
    ah.first_wave_docs = std::max(1u, ix->hq_first_wave_docs);
    ah.buf_docs = std::max(ah.wave_docs, ah.first_wave_docs);
    ah.counter_idx = 3;
    const int mode = comp32 ? 3 : (!plain16 ? 1 : ix->hq_mode);  // 1 byte index, 2 perfect hash, 3 bitmap + rank
    const int hq_threads = 256;
    size_t qbytes = 0;
    kern_t kh = nullptr, kl = nullptr;  // compact-query kernel, sorted-query kernel (layouts without the dense one)
    const QueryKind qk = mode == 1 ? Q_BYTE : (mode == 2 ? Q_HASH : Q_RANK);
    if (mode == 1) {
        ah.qd_words = ((ix->ix.dim + 15u) / 16u) * 4u;
        qbytes = 1024 + (size_t)ah.qd_words * 4;
    } else if (mode == 2) {
        qbytes = (size_t)HQ_SLOTS * 6;
    } else {
        ah.qd_words = ((ix->ix.dim + 127u) / 128u) * 4u;
        qbytes = 1024 + (size_t)ah.qd_words * 5;
    }
    const bool use_tma = plain16 && qk == Q_BYTE && ix->tma;
    if (use_tma) kh = pick_rec16_tma(hk);
    else if (plain16) kh = qk == Q_BYTE && ix->occ16 != 4 ? pick_rec16_var(hk, ix->occ16) : pick_rec16(qk, hk);
    else if (vkind == SGPU_VAL_DOTVBYTE) kh = pick_vb(qk, hk, ix->occvb), kl = pick_vb(Q_SORTED, hk, 4);
    else if (comp32 && vkind == SGPU_VAL_F16) kh = pick_rec32(qk, hk, ix->occ32), kl = pick_rec32(Q_SORTED, hk, 4);
    else if (comp32) kh = pick_rec32v(vkind, qk, hk), kl = pick_rec32v(vkind, Q_SORTED, hk);
    else kh = pick_rec16v(vkind, qk, hk), kl = pick_rec16v(vkind, Q_SORTED, hk);
    if (!kh) {
        shost::set_error("no search kernel for this index layout");
        return SGPU_EUNSUPPORTED;
    }
    ah.cand_cap = (uint32_t)std::min(hq_threads, std::max(32, ((ix->hq_cand_cap + 3) / 4) * 4));
    if (ad.n_knn > 0)  // Knn::refine snapshots the heap into the four candidate arrays
        ah.cand_cap = (uint32_t)std::min<uint32_t>(hq_threads, std::max<uint32_t>(ah.cand_cap, ((k + 15) / 16) * 4));
    const size_t smem_h = ((qbytes + 15) & ~(size_t)15) + wave_bytes(ah) + (use_tma ? (size_t)(hq_threads / 32) * TMA_WARP_BYTES + 128 : 0);
    bool hq_ok = (!plain16 || ix->hq_enabled) && smem_h + 1024 <= (comp32 || use_tma ? ix->smem_optin : ix->smem_optin / 2);
    int hq_ctas = 0;
    uint32_t ctas_per_sm = 0;
    if (hq_ok) {
        CK(cudaFuncSetAttribute(kh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
        int occ = 0;
        CK(cudaFuncSetAttribute(kh, cudaFuncAttributePreferredSharedMemoryCarveout, 100));  // not the previous call's
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kh, hq_threads, smem_h));
        if (ix->hq_ctas_per_sm > 0) occ = std::min(occ, ix->hq_ctas_per_sm);
        // Shared memory and L1 share one 256 KB array and the in-flight gathers live in L1: ask for the smallest
        // shared-memory carveout that still holds `occ` CTAs (measured: a 228 KB carveout costs 20 % on k_search)
        cudaFuncAttributes fa{};
        CK(cudaFuncGetAttributes(&fa, kh));
        const size_t need = (size_t)occ * (smem_h + fa.sharedSizeBytes + 1024);
        int pct = ix->hq_carveout_pct > 0 ? ix->hq_carveout_pct
                                          : (int)std::min<size_t>(100, (need * 100 + ix->smem_per_sm - 1) / ix->smem_per_sm);
        CK(cudaFuncSetAttribute(kh, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        hq_ok = occ >= 1;
        hq_ctas = occ * ix->n_sm;
        ctas_per_sm = (uint32_t)occ;
    }
    pd.ctas_per_sm = ctas_per_sm;
    if (ad.n_knn > 0 && ((hq_ok && k > 4 * ah.cand_cap) || (dense_ok && k > 4 * ad.cand_cap))) {
        shost::set_error("n_knn > 0: k exceeds the snapshot buffer of the search kernel");
        return SGPU_EUNSUPPORTED;
    }
    if (!hq_ok && !dense_ok) {
        shost::set_error("neither the compact-query nor the dense-query kernel fits this index in shared memory");
        return SGPU_EUNSUPPORTED;
    }
    // the sorted-query kernel shares the wave geometry of the compact kernel; its table is sized per batch
    const bool sorted_ok = !dense_ok && kl != nullptr;
    const size_t smem_l_base = wave_bytes(ah) + 32;
    const uint32_t sorted_cap = sorted_ok && ix->smem_optin > smem_l_base + 2048
                                    ? (uint32_t)((ix->smem_optin - smem_l_base - 2048) / 8) : 0u;
    // routing (k_terms): queries the compact table cannot take go to the dense / sorted kernel
    const uint32_t max_nnz_compact = mode == 2 ? (uint32_t)HQ_MAX_NNZ : 255u;
    const uint32_t max_nnz_ok = dense_ok ? 0xffffffffu : (sorted_ok ? std::max(sorted_cap, max_nnz_compact) : max_nnz_compact);
    const uint32_t hq_max_nnz = hq_ok ? max_nnz_compact : 0u;
    const uint32_t hq_tries = mode == 2 ? (uint32_t)HQ_TRIES : 0u;

    CK(cudaEventRecord(ix->ev[0], st));
    Batch all{dq->offsets, dq->comps, dq->values, nq, 0};
    k_prep<<<(nq + 3) / 4, 128, 0, st>>>(all, ix->ix.dim, p->query_cut, max_nnz_ok, ix->d_nterms.as<uint32_t>(),
                                         ix->d_status.as<uint32_t>(), ix->d_prep.as<uint32_t>());
    CK(cudaGetLastError());
    ++pd.launches;
    CK(cudaEventRecord(ix->ev[1], st));
    uint32_t cut_eff = std::max(1u, p->query_cut);
    if (p->query_cut > FAST_CUT) {  // size the scratch by what the queries really have
        CK(cudaMemcpyAsync(hc->prep, ix->d_prep.p, 32, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (hc->prep[2] != 0) {
            shost::set_error("Query components must be sorted in ascending order and be < dim (" +
                             std::to_string(hc->prep[2]) + " invalid queries)");
            return SGPU_EINVAL;
        }
        cut_eff = std::max(1u, hc->prep[1]);
    }
    const uint32_t est_stride = std::max(32u, (ix->max_blocks + 31u) & ~31u);
    // chunk the batch so that the estimate scratch stays within budget
    const uint64_t per_query = (uint64_t)cut_eff * est_stride * 4 + (uint64_t)est_stride * 16 + (uint64_t)cut_eff * 4 +
                               (uint64_t)k * 4 + 12;
    const uint32_t chunk = (uint32_t)std::min<uint64_t>(nq, std::max<uint64_t>(1, ix->scratch_bytes / per_query));
    const bool multi = chunk < nq;
    CK(ix->d_terms.ensure((size_t)chunk * cut_eff * 4));
    CK(ix->d_est.ensure((size_t)chunk * cut_eff * est_stride * 4));
    CK(ix->d_order.ensure((size_t)chunk * est_stride * 16));
    CK(ix->d_keys.ensure((size_t)chunk * k * 4));
    CK(ix->d_hmult.ensure((size_t)chunk * 8));
    CK(ix->d_qlist.ensure((size_t)chunk * 8));

    for (uint32_t q0 = 0; q0 < nq; q0 += chunk) {
        const uint32_t n = std::min(chunk, nq - q0);
        Batch b{dq->offsets, dq->comps, dq->values, n, q0};
        Scratch sc{};
        sc.terms = ix->d_terms.as<uint32_t>();
        sc.nterms = ix->d_nterms.as<uint32_t>() + q0;
        sc.status = ix->d_status.as<uint32_t>() + q0;
        sc.est = ix->d_est.as<float>();
        sc.sel = ix->d_order.as<uint4>();
        sc.counters = ix->d_counters.as<uint32_t>();
        sc.hmult = ix->d_hmult.as<uint32_t>();
        sc.cost = ix->d_hmult.as<uint32_t>() + chunk;
        sc.qlist_hq = ix->d_qlist.as<uint32_t>();
        sc.qlist_dense = ix->d_qlist.as<uint32_t>() + chunk;
        sc.out_keys = ix->d_keys.as<uint32_t>();
        sc.stats = ix->d_stats.as<unsigned long long>();
        sc.est_stride = est_stride;
        sc.cut_eff = cut_eff;
        CK(cudaMemsetAsync(ix->d_counters.p, 0, 32, st));
        CK(cudaEventRecord(ix->ev[2], st));
        k_terms<<<(n + 3) / 4, 128, 0, st>>>(b, sc, ix->ix.lists, hq_max_nnz, (uint32_t)HQ_LOG2_SLOTS, hq_tries);
        CK(cudaGetLastError());
        k_route<<<1, ROUTE_THREADS, 0, st>>>(sc, n, cut_eff * ix->max_list_post);
        CK(cudaGetLastError());
        pd.launches += 2;
        CK(cudaEventRecord(ix->ev[3], st));
        const uint64_t tasks = (uint64_t)n * cut_eff;
        const int skip_fused = 0;
        k_est<<<(unsigned)((tasks + EST_WARPS - 1) / EST_WARPS), EST_WARPS * 32, 0, st>>>(ix->ix, b, sc, skip_fused);
        CK(cudaGetLastError());
        ++pd.launches;
        if (ad.first_sorted) {
            if (ix->order_warp) {
                k_order_warp<false><<<(n + ORDER_WARPS - 1) / ORDER_WARPS, ORDER_WARPS * 32, 0, st>>>(ix->ix, b, sc);
                CK(cudaGetLastError());
                ++pd.launches;
                if (ix->max_blocks > 512) {
                    k_order_warp<true><<<(n + ORDER_WARPS - 1) / ORDER_WARPS, ORDER_WARPS * 32, 0, st>>>(ix->ix, b, sc);
                    CK(cudaGetLastError());
                    ++pd.launches;
                }
            }
            if (!ix->order_warp || ix->max_blocks > ORDER_WARP_MAX) {
                k_order<<<n, ORDER_THREADS, 0, st>>>(ix->ix, b, sc, ix->order_warp);
                CK(cudaGetLastError());
                ++pd.launches;
            }
        }
        CK(cudaEventRecord(ix->ev[4], st));
        ad.b = ah.b = b;
        ad.sc = ah.sc = sc;
        ad.out_scores = ah.out_scores = d_scores + (uint64_t)q0 * k;
        ad.out_counts = ah.out_counts = d_counts + q0;
        if (hq_ok) {
            ah.qlist = sc.qlist_hq;
            ah.n_list = sc.counters + 4;
            kh<<<hq_ctas, hq_threads, smem_h, st>>>(ah);
            CK(cudaGetLastError());
            ++pd.launches;
        }
        if (dense_ok) {
            ad.qlist = sc.qlist_dense;
            ad.n_list = sc.counters + 5;
            kd<<<ctas, DENSE_THREADS, smem_d, st>>>(ad);
            CK(cudaGetLastError());
            ++pd.launches;
        }
        CK(cudaEventRecord(ix->ev[5], st));
        const uint64_t tot = (uint64_t)n * k;
        k_finish<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ix->ix.rec_start, ix->ix.n_docs, sc.out_keys,
                                                                ad.out_counts, k, n, d_ids + (uint64_t)q0 * k);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ix->ev[6], st));
        ++pd.launches;
        if (sorted_ok) {  // what a long-query pass over this chunk would need
            pd.k_long = kl;
            pd.a_long = ah;
            pd.a_long.qlist = sc.qlist_dense;
            pd.a_long.n_list = sc.counters + 5;
            pd.a_long.counter_idx = 0;
            pd.long_threads = hq_threads;
            pd.long_smem = smem_l_base;
        }
        if (multi) {
            CK(cudaMemcpyAsync(hc->counters, ix->d_counters.p, 32, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(hc->prep, ix->d_prep.p, 32, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (int rc = collect_chunk_times(ix, pd)) return rc;
            if (sorted_ok && hc->counters[5] > 0) {
                if (int rc = run_long_pass(ix, pd, hc->prep[6], n, d_ids + (uint64_t)q0 * k)) return rc;
                CK(cudaStreamSynchronize(st));
            }
        } else {
            pd.events_pending = true;
            pd.long_pass_pending = sorted_ok;
        }
    }
    CK(cudaMemcpyAsync(hc->prep, ix->d_prep.p, 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hc->counters, ix->d_counters.p, 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hc->stats, ix->d_stats.p, 12 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    pd.active = true;
    return SGPU_OK;
}

// One synchronisation: validation result, statistics, and (rarely) the long-query pass.  *reran is set when results
// were (re)written after the caller may already have enqueued its device-to-host copies.
int finish_search(SgpuIndex* ix, Pending& pd, SgpuSearchStats* stats, bool* reran) {
    if (reran) *reran = false;
    if (!pd.active) return SGPU_OK;
    cudaStream_t st = ix->stream;
    CK(cudaStreamSynchronize(st));
    HostCtl* hc = ix->h_ctl.as<HostCtl>();
    if (hc->prep[2] != 0) {
        shost::set_error("Query components must be sorted in ascending order and be < dim (" +
                         std::to_string(hc->prep[2]) + " invalid queries)");
        return SGPU_EINVAL;
    }
    if (hc->prep[3] != 0) {
        shost::set_error(std::to_string(hc->prep[3]) + " queries have more components than the search kernels of this "
                         "index layout can stage in shared memory");
        return SGPU_EUNSUPPORTED;
    }
    if (pd.events_pending)
        if (int rc = collect_chunk_times(ix, pd)) return rc;
    if (pd.long_pass_pending && hc->counters[5] > 0) {
        CK(cudaEventRecord(ix->ev[4], st));
        if (int rc = run_long_pass(ix, pd, hc->prep[6], pd.nq, pd.d_ids)) return rc;
        CK(cudaEventRecord(ix->ev[5], st));
        CK(cudaMemcpyAsync(hc->stats, ix->d_stats.p, 12 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ix->ev[4], ix->ev[5]));
        pd.ms_search += ms;
        if (reran) *reran = true;
    }
    if (stats) {
        float ms_prep = 0.f;
        CK(cudaEventElapsedTime(&ms_prep, ix->ev[0], ix->ev[1]));
        const unsigned long long* hs = hc->stats;
        for (int i = 0; i < 6; ++i) stats->phase_cycles[i] = hs[4 + i];
        stats->ms_prep = ms_prep + pd.ms_terms;
        stats->ms_summary = pd.ms_sum;
        stats->ms_search = pd.ms_search;
        stats->ms_finish = pd.ms_fin;
        stats->ms_total = stats->ms_prep + pd.ms_sum + pd.ms_search + pd.ms_fin;
        stats->n_launches = pd.launches;
        stats->ctas_per_sm = pd.ctas_per_sm;
        stats->docs_scored = hs[0];
        stats->blocks_scored = hs[1];
        stats->blocks_pushed = hs[2];
        stats->waves = hs[10];
        stats->select_passes = hs[11];
        const uint32_t vkind = pd.vkind;
        const uint64_t cb = 8ull * ((ix->ix.comp32 ? 4 : 2) + (vkind == SGPU_VAL_F32 ? 4 : (vkind == SGPU_VAL_FIXEDU8 ? 1 : 2)));
        stats->fwd_bytes = hs[3] * (ix->ix.vbyte ? 1ull : cb);
    }
    pd.active = false;
    return SGPU_OK;
}

int search_device_impl(SgpuIndex* ix, const SgpuQueryBatch* dq, const SgpuSearchParams* p, uint64_t* d_ids,
                       float* d_scores, uint32_t* d_counts, SgpuSearchStats* stats) {
    if (stats) std::memset(stats, 0, sizeof(*stats));
    Pending pd;
    if (int rc = enqueue_search(ix, dq, p, d_ids, d_scores, d_counts, pd)) return rc;
    return finish_search(ix, pd, stats, nullptr);
}

}  // namespace

extern "C" {

int sgpu_index_create(const SgpuIndexView* view, int device, SgpuIndex** out) {
    try {
        return create_impl(view, device, out);
    } catch (const std::bad_alloc&) {
        shost::set_error("out of host memory");
        return SGPU_ENOMEM;
    }
}

void sgpu_index_destroy(SgpuIndex* index) { delete index; }

uint64_t sgpu_index_device_bytes(const SgpuIndex* index) { return index ? index->image_bytes : 0; }

int sgpu_index_set_stream(SgpuIndex* ix, void* cuda_stream) {
    if (!ix) return SGPU_EINVAL;
    ix->stream = cuda_stream ? (cudaStream_t)cuda_stream : ix->own_stream;
    return SGPU_OK;
}

int sgpu_index_set_knn(SgpuIndex* ix, const uint64_t* neighbours, uint32_t knn_dim) {
    return set_knn_impl(ix, neighbours, knn_dim);
}

int sgpu_index_set_option(SgpuIndex* ix, const char* name, int64_t value) {
    if (!ix || !name) return SGPU_EINVAL;
    std::string n(name);
    if (n == "hq") {  // 0: dense-query kernel only; compact query: 1 byte index, 2 perfect hash, 3 bitmap + rank
        ix->hq_enabled = value != 0;
        if (value >= 1 && value <= 3) ix->hq_mode = (int)value;
        return SGPU_OK;
    }
    if (n == "bucket") {
        ix->bucket = value != 0;
        return SGPU_OK;
    }
    if (n == "tma") {
        ix->tma = value != 0;
        return SGPU_OK;
    }
    if (n == "order_warp") {
        ix->order_warp = value != 0;
        return SGPU_OK;
    }
    if (n == "wide_heap") {
        ix->wide_heap = value != 0;
        return SGPU_OK;
    }
    if (n == "occ16") {
        ix->occ16 = (int)value;  // 4 (two documents per group), 41, 51 (one document per group, 4 / 5 CTAs per SM)
        return SGPU_OK;
    }
    if (n == "occvb") {
        ix->occvb = (int)value;
        return SGPU_OK;
    }
    if (n == "occ32") {
        ix->occ32 = (int)value;  // 4, 3, 2: CTAs / SM at two documents per group; 41, 51: one document per group
        return SGPU_OK;
    }

    if (n == "hq_carveout_pct") {  // 0 = automatic
        ix->hq_carveout_pct = (int)std::min<int64_t>(100, std::max<int64_t>(0, value));
        return SGPU_OK;
    }
    if (value <= 0) {
        shost::set_error("option values must be positive");
        return SGPU_EINVAL;
    }
    if (n == "hq_cand_cap") ix->hq_cand_cap = (int)value;
    else if (n == "hq_wave_docs") ix->hq_wave_docs = (uint32_t)value;
    else if (n == "hq_first_wave_docs") ix->hq_first_wave_docs = (uint32_t)value;
    else if (n == "hq_ctas_per_sm") ix->hq_ctas_per_sm = (int)value;
    else if (n == "wave_docs") ix->wave_docs = (uint32_t)value;
    else if (n == "first_wave_docs") ix->first_wave_docs = (uint32_t)value;
    else if (n == "ctas") ix->ctas = (int)value;
    else if (n == "scratch_mb") ix->scratch_bytes = (uint64_t)value << 20;
    else {
        shost::set_error("unknown option: " + n);
        return SGPU_EINVAL;
    }
    return SGPU_OK;
}

int sgpu_batch_search_device(SgpuIndex* index, const SgpuQueryBatch* d_queries, const SgpuSearchParams* params,
                             uint64_t* d_out_ids, float* d_out_scores, uint32_t* d_out_counts,
                             SgpuSearchStats* stats) {
    return search_device_impl(index, d_queries, params, d_out_ids, d_out_scores, d_out_counts, stats);
}

int sgpu_batch_search(SgpuIndex* ix, const SgpuQueryBatch* q, const SgpuSearchParams* p, uint64_t* out_ids,
                      float* out_scores, uint32_t* out_counts, SgpuSearchStats* stats) {
    if (!ix || !q || !p || !out_ids || !out_scores || !out_counts) {
        shost::set_error("sgpu_batch_search: null argument");
        return SGPU_EINVAL;
    }
    if (p->k == 0) {
        shost::set_error("k must be > 0 (KHeap::new asserts k > 0)");
        return SGPU_EINVAL;
    }
    CK(cudaSetDevice(ix->device));
    const uint64_t nq = q->n_queries;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (nq == 0) return SGPU_OK;
    const uint64_t nnz = q->offsets[nq];
    cudaStream_t st = ix->stream;
    // Inputs: page-locked caller buffers are copied by DMA as they are; pageable ones are staged through pinned memory
    // piece by piece, so that the DMA of one array runs under the host copy of the next.
    const size_t b_off = (nq + 1) * 8, b_c = nnz * 4, b_v = nnz * 4;
    CK(ix->h_in.ensure(b_off + b_c + b_v));
    uint8_t* hin = ix->h_in.as<uint8_t>();
    CK(ix->d_qoff.ensure(b_off));
    CK(ix->d_qcomps.ensure(std::max<size_t>(b_c, 4)));
    CK(ix->d_qvals.ensure(std::max<size_t>(b_v, 4)));
    auto upload = [&](void* dst, const void* src, uint8_t* stage, size_t bytes) -> int {
        if (!bytes) return SGPU_OK;
        if (!is_page_locked(src)) {
            std::memcpy(stage, src, bytes);
            src = stage;
        }
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return SGPU_OK;
    };
    if (int rc = upload(ix->d_qoff.p, q->offsets, hin, b_off)) return rc;
    if (int rc = upload(ix->d_qcomps.p, q->comps, hin + b_off, b_c)) return rc;
    if (int rc = upload(ix->d_qvals.p, q->values, hin + b_off + b_c, b_v)) return rc;
    const size_t o_ids = nq * p->k * 8, o_sc = nq * p->k * 4, o_cnt = nq * 4;
    CK(ix->d_out_ids.ensure(o_ids));
    CK(ix->d_out_scores.ensure(o_sc));
    CK(ix->d_out_counts.ensure(o_cnt));
    SgpuQueryBatch dq{nq, ix->d_qoff.as<uint64_t>(), ix->d_qcomps.as<uint32_t>(), ix->d_qvals.as<float>()};
    Pending pd;
    if (int rc = enqueue_search(ix, &dq, p, ix->d_out_ids.as<uint64_t>(), ix->d_out_scores.as<float>(),
                                ix->d_out_counts.as<uint32_t>(), pd))
        return rc;
    // Outputs: straight into page-locked caller buffers, else through the pinned staging buffer
    const bool direct_out = is_page_locked(out_ids) && is_page_locked(out_scores) && is_page_locked(out_counts);
    uint8_t* hout = nullptr;
    if (!direct_out) {
        CK(ix->h_out.ensure(o_ids + o_sc + o_cnt));
        hout = ix->h_out.as<uint8_t>();
    }
    auto copy_out = [&]() -> int {
        CK(cudaMemcpyAsync(direct_out ? (void*)out_ids : (void*)hout, ix->d_out_ids.p, o_ids, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(direct_out ? (void*)out_scores : (void*)(hout + o_ids), ix->d_out_scores.p, o_sc,
                           cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(direct_out ? (void*)out_counts : (void*)(hout + o_ids + o_sc), ix->d_out_counts.p, o_cnt,
                           cudaMemcpyDeviceToHost, st));
        return SGPU_OK;
    };
    if (int rc = copy_out()) return rc;
    bool reran = false;
    if (int rc = finish_search(ix, pd, stats, &reran)) return rc;  // the one synchronisation of the call
    if (reran) {
        if (int rc = copy_out()) return rc;
        CK(cudaStreamSynchronize(st));
    }
    if (!direct_out) {
        std::memcpy(out_ids, hout, o_ids);
        std::memcpy(out_scores, hout + o_ids, o_sc);
        std::memcpy(out_counts, hout + o_ids + o_sc, o_cnt);
    }
    return SGPU_OK;
}

int sgpu_host_alloc(uint64_t bytes, void** out) {
    if (!out) {
        shost::set_error("sgpu_host_alloc: null argument");
        return SGPU_EINVAL;
    }
    *out = nullptr;
    CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return SGPU_OK;
}

void sgpu_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int sgpu_exact_search(SgpuIndex* ix, const SgpuQueryBatch* q, uint32_t k, uint64_t* out_ids, float* out_scores,
                      uint32_t* out_counts, float* ms_kernel) {
    if (!ix || !q || !out_ids || !out_scores || !out_counts || k == 0 || k > 1024) {
        shost::set_error("sgpu_exact_search: bad argument");
        return SGPU_EINVAL;
    }
    if (ix->ix.vbyte) {
        shost::set_error("sgpu_exact_search: not available for DotVByte indexes (the reference has no FlatIndex over them)");
        return SGPU_EUNSUPPORTED;
    }
    CK(cudaSetDevice(ix->device));
    const uint64_t nq = q->n_queries;
    if (nq == 0) return SGPU_OK;
    const uint64_t nnz = q->offsets[nq];
    cudaStream_t st = ix->stream;
    uint64_t max_nnz = 0;
    for (uint64_t i = 0; i < nq; ++i) {
        max_nnz = std::max(max_nnz, q->offsets[i + 1] - q->offsets[i]);
        for (uint64_t j = q->offsets[i]; j < q->offsets[i + 1]; ++j)
            if (q->comps[j] >= ix->ix.dim || (j > q->offsets[i] && q->comps[j] < q->comps[j - 1])) {
                shost::set_error("Query components must be sorted in ascending order and be < dim");
                return SGPU_EINVAL;
            }
    }
    CK(ix->d_qoff.ensure((nq + 1) * 8));
    CK(ix->d_qcomps.ensure(std::max<size_t>(nnz * 4, 4)));
    CK(ix->d_qvals.ensure(std::max<size_t>(nnz * 4, 4)));
    CK(cudaMemcpyAsync(ix->d_qoff.p, q->offsets, (nq + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nnz) {
        CK(cudaMemcpyAsync(ix->d_qcomps.p, q->comps, nnz * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ix->d_qvals.p, q->values, nnz * 4, cudaMemcpyHostToDevice, st));
    }
    CK(ix->d_out_ids.ensure(nq * k * 8));
    CK(ix->d_out_scores.ensure(nq * k * 4));
    CK(ix->d_out_counts.ensure(nq * 4));
    // segments sized to stay L2 resident while many queries stream over them
    const uint64_t N = ix->ix.n_docs;
    const uint32_t seg_docs = 131072;
    const uint32_t n_seg = (uint32_t)std::max<uint64_t>(1, (N + seg_docs - 1) / seg_docs);
    DevBuf part_keys, part_scores;
    CK(part_keys.ensure((size_t)nq * n_seg * k * 4));
    CK(part_scores.ensure((size_t)nq * n_seg * k * 4));
    ExactArgs ea{};
    ea.fwd = ix->ix.fwd;
    ea.rec_start = ix->ix.rec_start;
    ea.n_docs = N;
    ea.q_off = ix->d_qoff.as<uint64_t>();
    ea.q_comps = ix->d_qcomps.as<uint32_t>();
    ea.q_vals = ix->d_qvals.as<float>();
    ea.nq = (uint32_t)nq;
    ea.k = k;
    ea.seg_docs = seg_docs;
    ea.n_seg = n_seg;
    ea.chunk_units = ix->ix.rec_chunk_units;
    ea.value_scale = ix->ix.value_scale;
    ea.part_keys = part_keys.as<uint32_t>();
    ea.part_scores = part_scores.as<float>();
    // dense f32 query for the u16/f16 layout when the vocabulary fits shared memory, else the sorted query (any length)
    const size_t tail = 2 * (size_t)((k + 3) & ~3u) * 4 + EXACT_CAND * 8 + 32;
    const uint32_t vkind = ix->ix.value_kind;
    const bool plain16 = !ix->ix.comp32 && vkind == SGPU_VAL_F16;
    const uint32_t dense_words = (ix->ix.dim + 31u) & ~31u;
    const bool dense = plain16 && (size_t)dense_words * 4 + tail + 1024 <= ix->smem_optin;
    ea.qd_words = dense ? dense_words : (uint32_t)std::max<uint64_t>(max_nnz, 1);
    const size_t smem = (dense ? (size_t)ea.qd_words * 4 : (size_t)ea.qd_words * 8 + 16) + tail;
    if (smem + 1024 > ix->smem_optin) {
        shost::set_error("sgpu_exact_search: a query does not fit in shared memory");
        return SGPU_EUNSUPPORTED;
    }
    exact_t ke = plain16 ? pick_exact_rec16(dense)
                         : (ix->ix.comp32 ? (vkind == SGPU_VAL_F16 ? pick_exact_rec32() : pick_exact_rec32v(vkind))
                                          : pick_exact_rec16v(vkind));
    if (!ke) {
        shost::set_error("sgpu_exact_search: no kernel for this index layout");
        return SGPU_EUNSUPPORTED;
    }
    CK(cudaFuncSetAttribute(ke, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaEventRecord(ix->ev[0], st));
    const uint64_t n_cta = (uint64_t)nq * n_seg;
    ke<<<(unsigned)n_cta, EXACT_THREADS, smem, st>>>(ea);
    CK(cudaGetLastError());
    k_exact_merge<<<(unsigned)nq, 32, 2 * (size_t)((k + 3) & ~3u) * 4, st>>>(
        ea, nullptr, ix->d_out_scores.as<float>(), ix->d_out_counts.as<uint32_t>(),
        ix->d_out_ids.as<uint64_t>());
    CK(cudaGetLastError());
    CK(cudaEventRecord(ix->ev[1], st));
    CK(cudaMemcpyAsync(out_ids, ix->d_out_ids.p, nq * k * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_scores, ix->d_out_scores.p, nq * k * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_counts, ix->d_out_counts.p, nq * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ms_kernel) CK(cudaEventElapsedTime(ms_kernel, ix->ev[0], ix->ev[1]));
    return SGPU_OK;
}

const char* sgpu_version(void) { return "seismic_b200 0.2.0 sm_100a"; }

// ---------------------------------------------------------------------------------------------------------
// Multi-GPU group: replicas + one NCCL gather (see include/seismic_b200.h)
// ---------------------------------------------------------------------------------------------------------
}  // extern "C"

#include <dlfcn.h>

namespace {

// the handful of NCCL entry points the gather needs, resolved from libnccl.so.2 at run time (if the process already
// holds an NCCL with that soname — e.g. torch's — dlopen returns it)
typedef struct ncclComm* nccl_comm_t;
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        Send = (decltype(Send))dlsym(lib, "ncclSend");
        Recv = (decltype(Recv))dlsym(lib, "ncclRecv");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Send && Recv && GetErrorString;
    }
};
NcclApi g_nccl;
constexpr int kNcclUint8 = 1;  // ncclUint8 (nccl.h)

#define NK(expr)                                                                                         \
    do {                                                                                                 \
        int _r = (expr);                                                                                 \
        if (_r != 0) {                                                                                   \
            shost::set_error(std::string(#expr) + ": " + g_nccl.GetErrorString(_r));                     \
            return SGPU_ECUDA;                                                                           \
        }                                                                                                \
    } while (0)

}  // namespace

struct SgpuGroup {
    std::vector<SgpuIndex*> idx;     // one replica per device
    std::vector<nccl_comm_t> comms;  // one communicator per device (ncclCommInitAll)
    std::vector<Pending> pend;
    cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr;
    ~SgpuGroup() {
        for (auto c : comms)
            if (c) g_nccl.CommDestroy(c);
        for (auto* i : idx) delete i;
        if (ev_g0) cudaEventDestroy(ev_g0);
        if (ev_g1) cudaEventDestroy(ev_g1);
    }
};

namespace {

int group_create_impl(const SgpuIndexView* view, const int* devices, int n, SgpuGroup** out) {
    if (!view || !devices || !out || n <= 0) {
        shost::set_error("sgpu_group_create: bad argument");
        return SGPU_EINVAL;
    }
    for (int a = 0; a < n; ++a)
        for (int b2 = a + 1; b2 < n; ++b2)
            if (devices[a] == devices[b2]) {
                shost::set_error("sgpu_group_create: duplicate device");
                return SGPU_EINVAL;
            }
    std::unique_ptr<SgpuGroup> g(new SgpuGroup());
    for (int r = 0; r < n; ++r) {
        SgpuIndex* ix = nullptr;
        if (int rc = create_impl(view, devices[r], &ix)) return rc;
        g->idx.push_back(ix);
    }
    g->pend.resize(n);
    if (n > 1) {
        if (!g_nccl.load()) {
            shost::set_error("sgpu_group_create: libnccl.so.2 not found (needed for the multi-GPU gather)");
            return SGPU_ECUDA;
        }
        g->comms.assign(n, nullptr);
        NK(g_nccl.CommInitAll(g->comms.data(), n, devices));
    }
    CK(cudaSetDevice(devices[0]));
    CK(cudaEventCreate(&g->ev_g0));
    CK(cudaEventCreate(&g->ev_g1));
    *out = g.release();
    return SGPU_OK;
}

// contiguous, balanced ranges: the first nq % n devices get one extra query
inline void shard_bounds(uint64_t nq, int r, int n, uint64_t* lo, uint64_t* hi) {
    const uint64_t base = nq / n, extra = nq % n;
    *lo = r * base + std::min<uint64_t>(r, extra);
    *hi = *lo + base + ((uint64_t)r < extra ? 1 : 0);
}

int group_search_impl(SgpuGroup* g, const SgpuQueryBatch* q, const SgpuSearchParams* p, uint64_t* out_ids,
                      float* out_scores, uint32_t* out_counts, SgpuSearchStats* stats, float* ms_gather) {
    if (!g || !q || !p || !out_ids || !out_scores || !out_counts) {
        shost::set_error("sgpu_group_batch_search: null argument");
        return SGPU_EINVAL;
    }
    if (p->k == 0) {
        shost::set_error("k must be > 0 (KHeap::new asserts k > 0)");
        return SGPU_EINVAL;
    }
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (ms_gather) *ms_gather = 0.f;
    const int n = (int)g->idx.size();
    const uint64_t nq = q->n_queries, k = p->k;
    if (nq == 0) return SGPU_OK;
    SgpuIndex* root = g->idx[0];
    // device 0 holds the gathered tuples of the whole batch
    CK(cudaSetDevice(root->device));
    CK(root->d_out_ids.ensure(nq * k * 8));
    CK(root->d_out_scores.ensure(nq * k * 4));
    CK(root->d_out_counts.ensure(nq * 4));
    // ---- every device: stage its slice, copy in, enqueue the search (nothing here waits for a device)
    for (int r = 0; r < n; ++r) {
        SgpuIndex* ix = g->idx[r];
        uint64_t lo, hi;
        shard_bounds(nq, r, n, &lo, &hi);
        g->pend[r] = Pending{};
        if (hi == lo) continue;
        CK(cudaSetDevice(ix->device));
        cudaStream_t st = ix->stream;
        const uint64_t nr = hi - lo, e0 = q->offsets[lo], nnz = q->offsets[hi] - e0;
        const size_t b_off = (nr + 1) * 8, b_c = nnz * 4;
        CK(ix->h_in.ensure(b_off + 2 * b_c));
        uint8_t* hin = ix->h_in.as<uint8_t>();
        uint64_t* ho = reinterpret_cast<uint64_t*>(hin);
        for (uint64_t i = 0; i <= nr; ++i) ho[i] = q->offsets[lo + i] - e0;
        if (nnz) {
            std::memcpy(hin + b_off, q->comps + e0, b_c);
            std::memcpy(hin + b_off + b_c, q->values + e0, b_c);
        }
        CK(ix->d_qoff.ensure(b_off));
        CK(ix->d_qcomps.ensure(std::max<size_t>(b_c, 4)));
        CK(ix->d_qvals.ensure(std::max<size_t>(b_c, 4)));
        CK(cudaMemcpyAsync(ix->d_qoff.p, hin, b_off, cudaMemcpyHostToDevice, st));
        if (nnz) {
            CK(cudaMemcpyAsync(ix->d_qcomps.p, hin + b_off, b_c, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(ix->d_qvals.p, hin + b_off + b_c, b_c, cudaMemcpyHostToDevice, st));
        }
        SgpuQueryBatch dq{nr, ix->d_qoff.as<uint64_t>(), ix->d_qcomps.as<uint32_t>(), ix->d_qvals.as<float>()};
        uint64_t* d_ids;
        float* d_sc;
        uint32_t* d_cnt;
        if (r == 0) {  // device 0 writes its share straight into the gathered buffers
            d_ids = root->d_out_ids.as<uint64_t>() + lo * k;
            d_sc = root->d_out_scores.as<float>() + lo * k;
            d_cnt = root->d_out_counts.as<uint32_t>() + lo;
        } else {
            CK(ix->d_out_ids.ensure(nr * k * 8));
            CK(ix->d_out_scores.ensure(nr * k * 4));
            CK(ix->d_out_counts.ensure(nr * 4));
            d_ids = ix->d_out_ids.as<uint64_t>(), d_sc = ix->d_out_scores.as<float>(), d_cnt = ix->d_out_counts.as<uint32_t>();
        }
        if (int rc = enqueue_search(ix, &dq, p, d_ids, d_sc, d_cnt, g->pend[r])) return rc;
    }
    // ---- validation / long-query passes of every device (one synchronisation each), then ONE fused gather
    for (int r = 0; r < n; ++r) {
        CK(cudaSetDevice(g->idx[r]->device));
        if (int rc = finish_search(g->idx[r], g->pend[r], r == 0 ? stats : nullptr, nullptr)) return rc;
    }
    CK(cudaSetDevice(root->device));
    CK(cudaEventRecord(g->ev_g0, root->stream));
    if (n > 1) {
        NK(g_nccl.GroupStart());
        for (int r = 1; r < n; ++r) {
            uint64_t lo, hi;
            shard_bounds(nq, r, n, &lo, &hi);
            const uint64_t nr = hi - lo;
            if (!nr) continue;
            SgpuIndex* ix = g->idx[r];
            NK(g_nccl.Send(ix->d_out_ids.p, nr * k * 8, kNcclUint8, 0, g->comms[r], ix->stream));
            NK(g_nccl.Send(ix->d_out_scores.p, nr * k * 4, kNcclUint8, 0, g->comms[r], ix->stream));
            NK(g_nccl.Send(ix->d_out_counts.p, nr * 4, kNcclUint8, 0, g->comms[r], ix->stream));
            NK(g_nccl.Recv(root->d_out_ids.as<uint64_t>() + lo * k, nr * k * 8, kNcclUint8, r, g->comms[0], root->stream));
            NK(g_nccl.Recv(root->d_out_scores.as<float>() + lo * k, nr * k * 4, kNcclUint8, r, g->comms[0], root->stream));
            NK(g_nccl.Recv(root->d_out_counts.as<uint32_t>() + lo, nr * 4, kNcclUint8, r, g->comms[0], root->stream));
        }
        NK(g_nccl.GroupEnd());
    }
    CK(cudaSetDevice(root->device));
    CK(cudaEventRecord(g->ev_g1, root->stream));
    const size_t o_ids = nq * k * 8, o_sc = nq * k * 4, o_cnt = nq * 4;
    CK(root->h_out.ensure(o_ids + o_sc + o_cnt));
    uint8_t* hout = root->h_out.as<uint8_t>();
    CK(cudaMemcpyAsync(hout, root->d_out_ids.p, o_ids, cudaMemcpyDeviceToHost, root->stream));
    CK(cudaMemcpyAsync(hout + o_ids, root->d_out_scores.p, o_sc, cudaMemcpyDeviceToHost, root->stream));
    CK(cudaMemcpyAsync(hout + o_ids + o_sc, root->d_out_counts.p, o_cnt, cudaMemcpyDeviceToHost, root->stream));
    CK(cudaStreamSynchronize(root->stream));
    for (int r = 1; r < n; ++r) {  // the senders' streams are done once the receiver is; keep the handles quiescent
        CK(cudaSetDevice(g->idx[r]->device));
        CK(cudaStreamSynchronize(g->idx[r]->stream));
    }
    if (ms_gather) {
        CK(cudaSetDevice(root->device));
        CK(cudaEventElapsedTime(ms_gather, g->ev_g0, g->ev_g1));
    }
    std::memcpy(out_ids, hout, o_ids);
    std::memcpy(out_scores, hout + o_ids, o_sc);
    std::memcpy(out_counts, hout + o_ids + o_sc, o_cnt);
    return SGPU_OK;
}

}  // namespace

extern "C" {

int sgpu_group_create(const SgpuIndexView* view, const int* devices, int n_devices, SgpuGroup** out) {
    try {
        return group_create_impl(view, devices, n_devices, out);
    } catch (const std::bad_alloc&) {
        shost::set_error("out of host memory");
        return SGPU_ENOMEM;
    }
}
void sgpu_group_destroy(SgpuGroup* group) { delete group; }
int sgpu_group_size(const SgpuGroup* group) { return group ? (int)group->idx.size() : 0; }
int sgpu_group_set_knn(SgpuGroup* group, const uint64_t* neighbours, uint32_t knn_dim) {
    if (!group) return SGPU_EINVAL;
    for (auto* ix : group->idx)
        if (int rc = set_knn_impl(ix, neighbours, knn_dim)) return rc;
    return SGPU_OK;
}
int sgpu_group_batch_search(SgpuGroup* group, const SgpuQueryBatch* queries, const SgpuSearchParams* params,
                            uint64_t* out_ids, float* out_scores, uint32_t* out_counts, SgpuSearchStats* stats,
                            float* ms_gather) {
    return group_search_impl(group, queries, params, out_ids, out_scores, out_counts, stats, ms_gather);
}

}  // extern "C"
