/*
 * seismic_b200.h — C ABI of the B200-native Seismic query hot path.
 *
 * Two groups of entry points, both plain C (pointers + sizes, no C++/torch types):
 *
 *  sgpu_*   THE DROP-IN BOUNDARY. What a Rust host (`seismic` crate) would bind with
 *           `extern "C"` to replace the body of
 *             InvertedIndexBase::search            (reference src/inverted_index.rs:153-234)
 *             PostingList::search/sort_and_search  (reference src/posting_list.rs:115-185)
 *             PostingList::evaluate_posting_block  (reference src/posting_list.rs:188-215)
 *             QuantizedSummary::distances          (reference src/quantized_summary.rs:64-160)
 *             KHeap push/peek/into_sorted_vec      (reference src/utils.rs:12-66)
 *           for a whole batch of queries (the rayon loops of src/pylib/mod.rs:629-652, :1129-1145).
 *           The host keeps owning the index; `SgpuIndexView` is a borrowed, read-only view of
 *           the reference's logical arrays, from which the library builds its HBM image once.
 *
 *  shost_*  Host-side stand-in for the Rust host code that cannot be compiled in this image
 *           (no cargo/rustc): dataset container, CPU index build (restating
 *           reference src/inverted_index.rs:354-389,603-686, src/posting_list.rs:227-450,
 *           src/quantized_summary.rs:289-406, src/utils.rs:68-237), index file I/O, the
 *           "seismic inner format" reader/writer (reference scripts/convert_json_to_inner_format.py:10-27)
 *           and the synthetic SPLADE-shaped corpus generator used by bench.py.
 *
 * All functions return 0 on success, a negative SGPU_E* code on failure; the message is
 * available through sgpu_last_error() (thread-local).  Nothing throws across this boundary.
 */
#ifndef SEISMIC_B200_H
#define SEISMIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGPU_OK 0
#define SGPU_EINVAL (-1)   /* bad argument (k==0, unsorted query, component >= dim, ...) */
#define SGPU_ECUDA (-2)    /* CUDA runtime failure / no device */
#define SGPU_ENOMEM (-3)
#define SGPU_EIO (-4)
#define SGPU_EUNSUPPORTED (-5)

/* forward-index value encodings (reference matrix: src/bin/perf_inverted_index.rs:95-139) */
#define SGPU_VAL_F16 0
#define SGPU_VAL_BF16 1
#define SGPU_VAL_F32 2
#define SGPU_VAL_FIXEDU8 3
#define SGPU_VAL_FIXEDU16 4
#define SGPU_VAL_DOTVBYTE 5 /* gap-coded components + u8 values (src/pylib/dotvbyte.rs:20-35) */

/* padding values written beyond out_counts[q] */
#define SGPU_PAD_ID UINT64_MAX

/*
 * Borrowed view of one logical Seismic index (all arrays host memory, little endian).
 * It mirrors the reference structs field by field:
 *   InvertedIndexBase{forward_index, posting_lists}           src/inverted_index.rs:38-52
 *   PostingList{packed_postings, block_offsets, summaries}    src/posting_list.rs:68-73
 *   QuantizedSummary{n_summaries, dim, component_ids, offsets,
 *                    summaries_ids, values, minimums, quants} src/quantized_summary.rs:14-24
 * with the per-list arrays of all `dim` lists concatenated (CSR of CSR) and the
 * storage-only compressions (Elias-Fano offsets, BitField ids, dense-vs-sparse offset
 * strategy, quantized_summary.rs:49-62) decoded to plain integers.
 */
typedef struct SgpuIndexView {
    uint32_t comp_bits;   /* 16 or 32: the reference's component type C (u16 | u32)            */
    uint32_t value_kind;  /* SGPU_VAL_*                                                        */
    uint64_t n_docs;      /* forward_index.len()                                               */
    uint64_t dim;         /* number of posting lists == forward_index.input_dim()              */
    float value_scale;    /* FIXEDU8/FIXEDU16/DOTVBYTE: value = code * value_scale; else 1     */
    uint32_t knn_dim;     /* Knn::dim: neighbours stored per document; 0 = no kNN graph        */

    /* forward index.  Units of fwd_offsets: elements (plain encodings) or BYTES of the packed
     * stream (DOTVBYTE).  Doc i = [fwd_offsets[i], fwd_offsets[i+1]).                          */
    const uint64_t* fwd_offsets; /* [n_docs+1]                                                 */
    const void* fwd_comps;       /* [nnz] u16 (comp_bits 16) or u32 (comp_bits 32) component ids, */
                                 /* ascending inside a doc; NULL for DOTVBYTE                    */
    const void* fwd_values;      /* [nnz] f16/bf16 bits (u16), f32, u8 or u16 codes; or the     */
                                 /* DOTVBYTE byte stream (then fwd_comps == NULL)               */
    const uint16_t* fwd_nnz;     /* DOTVBYTE only: [n_docs] number of components per doc        */

    /* posting lists */
    const uint64_t* list_post_start; /* [dim+1] into postings                                   */
    const uint64_t* postings;        /* packed (start<<16)|len, src/posting_list.rs:38-52       */
    const uint64_t* list_blk_start;  /* [dim+1] into blk_min/blk_quant; list l has              */
                                     /*   B_l = list_blk_start[l+1]-list_blk_start[l] blocks    */
    const uint32_t* blk_post_off;    /* [TB+dim]; list l owns B_l+1 entries starting at         */
                                     /*   list_blk_start[l]+l, relative to list_post_start[l]   */
    const float* blk_min;            /* [TB] QuantizedSummary::minimums                         */
    const float* blk_quant;          /* [TB] QuantizedSummary::quants                           */
    const uint64_t* list_sc_start;   /* [dim+1] into sc_comp (summary component ids per list)   */
    const uint32_t* sc_comp;         /* [TSC] ascending inside a list (component_ids)           */
    const uint64_t* list_ent_start;  /* [dim+1] into ent_blk/ent_code                           */
    const uint32_t* sc_run_off;      /* [TSC+dim]; list l owns n_sc_l+1 entries starting at     */
                                     /*   list_sc_start[l]+l, relative to list_ent_start[l]     */
    const uint16_t* ent_blk;         /* [TE] summaries_ids (block id inside its list)           */
    const uint8_t* ent_code;         /* [TE] values (u8 codes)                                  */

    /* optional kNN graph == Knn{n_vecs, dim, neighbours} src/inverted_index.rs:430-434 with the BitField
     * decoded: row d holds the knn_dim neighbours of document d, best first; SGPU_PAD_ID = no neighbour
     * (the reference cannot represent a short row: it concatenates the rows, :487-491).  NULL = none.  */
    const uint64_t* knn_neighbours;  /* [n_docs * knn_dim]                                      */
} SgpuIndexView;

/* A batch of sparse queries in CSR form.  Components must be non-decreasing inside a query
 * (the reference asserts is_sorted, src/inverted_index.rs:172-175) and < dim.                  */
typedef struct SgpuQueryBatch {
    uint64_t n_queries;
    const uint64_t* offsets; /* [n_queries+1]                                                   */
    const uint32_t* comps;   /* [offsets[n_queries]]                                            */
    const float* values;     /* [offsets[n_queries]]                                            */
} SgpuQueryBatch;

/* Search parameters == arguments of InvertedIndexBase::search (src/inverted_index.rs:153-161). */
typedef struct SgpuSearchParams {
    uint32_t k;
    uint32_t query_cut;
    float heap_factor;
    uint32_t n_knn;        /* > 0: Knn::refine with min(n_knn, knn_dim) neighbours; ignored without a graph */
    int32_t first_sorted;  /* Python default True (src/pylib/mod.rs:497-498), CLI default false */
} SgpuSearchParams;

/* Per-call device timing + work counters (optional, may be NULL). */
typedef struct SgpuSearchStats {
    float ms_total;          /* all kernels of the call, CUDA events on the library stream      */
    float ms_prep;           /* term selection + validation                                     */
    float ms_summary;        /* Loop A: quantized-summary estimates (+ first-list ordering)     */
    float ms_search;         /* Loop B: persistent traversal/scoring/top-k kernel               */
    float ms_finish;         /* key -> doc id mapping                                           */
    uint32_t n_launches;     /* kernels launched by the call                                    */
    uint32_t ctas_per_sm;    /* resident CTAs per SM of the compact-query k_search launch        */
    uint64_t docs_scored;    /* forward-index vectors actually read (incl. speculative ones)    */
    uint64_t blocks_scored;  /* blocks whose docs were read                                     */
    uint64_t blocks_pushed;  /* blocks that survived the exact replay (== reference evaluated)  */
    uint64_t fwd_bytes;      /* bytes of forward-index records read                             */
    uint64_t phase_cycles[6];/* SM clocks summed over CTAs: fetch+stage, select, gather postings,
                                score, replay, results (k_search's own phase profile)           */
    uint64_t waves;          /* waves of documents scored (k_search)                            */
    uint64_t select_passes;  /* candidate-selection passes (k_search)                           */
} SgpuSearchStats;

typedef struct SgpuIndex SgpuIndex; /* opaque: HBM image + scratch + stream on one device        */

/* Build the HBM image of `view` on CUDA device `device`.  The view is not retained. */
int sgpu_index_create(const SgpuIndexView* view, int device, SgpuIndex** out);
void sgpu_index_destroy(SgpuIndex* index);
/* Bytes of HBM held by the image (without per-batch scratch). */
uint64_t sgpu_index_device_bytes(const SgpuIndex* index);

/*
 * Batched InvertedIndexBase::search with HOST buffers (copies in and out are part of the call).
 *   out_ids    [n_queries*k] doc indices, best first, padded with SGPU_PAD_ID
 *   out_scores [n_queries*k] dot products,            padded with -inf
 *   out_counts [n_queries]   number of valid results (the reference may return < k,
 *                            src/bin/perf_inverted_index.rs:201-206)
 * Results are returned in input order.
 * Caller buffers that are page-locked (sgpu_host_alloc, cudaMallocHost, cudaHostRegister) are read / written by DMA
 * directly; pageable ones are staged through the library's pinned buffers.
 */
int sgpu_batch_search(SgpuIndex* index, const SgpuQueryBatch* queries, const SgpuSearchParams* params,
                      uint64_t* out_ids, float* out_scores, uint32_t* out_counts, SgpuSearchStats* stats);

/* Same, but every pointer in `queries` and the three outputs are DEVICE pointers on the
 * index's device; work is enqueued on the library stream and the call returns after it
 * completes (stats are then final).  Used for the HBM-resident measurement and by callers
 * that post-process results on the GPU (e.g. the NCCL gather of result tuples). */
int sgpu_batch_search_device(SgpuIndex* index, const SgpuQueryBatch* d_queries, const SgpuSearchParams* params,
                             uint64_t* d_out_ids, float* d_out_scores, uint32_t* d_out_counts,
                             SgpuSearchStats* stats);

/* Enqueue all work of this index on `cuda_stream` (a cudaStream_t of the index's device, e.g. the caller's
 * torch stream) instead of the library's private stream; NULL restores the private stream. */
int sgpu_index_set_stream(SgpuIndex* index, void* cuda_stream);

/* Attach, replace or (neighbours == NULL) drop the kNN graph used by Knn::refine (src/inverted_index.rs:551-593):
 * `neighbours` is a HOST array [n_docs * knn_dim] of document ids, row d = neighbours of document d, best first,
 * SGPU_PAD_ID = none.  Replaces Knn::new_from_serialized / InvertedIndex::add_knn (src/inverted_index.rs:494-549)
 * on the device side.  Not available for DotVByte indexes (the reference class has no kNN, src/pylib/dotvbyte.rs). */
int sgpu_index_set_knn(SgpuIndex* index, const uint64_t* neighbours, uint32_t knn_dim);

/* Tuning knobs (they never change results; tests/test_gpu_parity.py runs every one of them against the oracle):
 *   scheduler     "hq_wave_docs", "hq_first_wave_docs", "hq_cand_cap" (compact-query kernel), "wave_docs",
 *                 "first_wave_docs", "ctas" (long-query kernel), "hq_ctas_per_sm", "hq_carveout_pct", "bucket",
 *                 "scratch_mb" (per-batch scratch; larger batches run in chunks)
 *   kernel build  "hq" (query table: 0 long-query kernel only, 1 byte index, 2 perfect hash, 3 bitmap + rank), "tma" (records staged by bulk
 *                 copies), "occ16" / "occvb" / "occ32" (documents in flight per 8-lane group and CTAs per SM of the
 *                 u16 / DotVByte / u32 kernels), "wide_heap" (register heap for 32 < k <= 128), "order_warp"
 * Returns SGPU_EINVAL for unknown names. */
int sgpu_index_set_option(SgpuIndex* index, const char* name, int64_t value);

/* Exact top-k by brute force over the forward index (ground truth for recall; mirrors
 * SeismicDataset.search via FlatIndex, src/inverted_index_wrapper.rs:721-742). Host buffers. */
int sgpu_exact_search(SgpuIndex* index, const SgpuQueryBatch* queries, uint32_t k, uint64_t* out_ids,
                      float* out_scores, uint32_t* out_counts, float* ms_kernel);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: the index replicated on several devices of one box, ONE batch split across them (the
 * reference's own parallelism is rayon over the queries of one batch_search call, src/pylib/mod.rs:629-652).
 * Single process, no Python / torch needed: one SgpuIndex per device, the query batch cut into contiguous ranges,
 * every device searches its range on its own stream, and the result tuples are gathered on the first device with ONE
 * fused NCCL group of send/recv pairs over NVLink (ncclCommInitAll; libnccl.so.2 is opened at run time), then copied
 * to the caller's host buffers.  Results are in input order and identical to a single-device call.
 * ------------------------------------------------------------------------------------------ */
typedef struct SgpuGroup SgpuGroup;
/* devices[n_devices]: CUDA ordinals (distinct); devices[0] receives the gather. */
int sgpu_group_create(const SgpuIndexView* view, const int* devices, int n_devices, SgpuGroup** out);
void sgpu_group_destroy(SgpuGroup* group);
int sgpu_group_size(const SgpuGroup* group);
/* kNN graph on every replica (see sgpu_index_set_knn). */
int sgpu_group_set_knn(SgpuGroup* group, const uint64_t* neighbours, uint32_t knn_dim);
/* Same contract as sgpu_batch_search (host buffers).  stats: device-side timings of devices[0]'s share; ms_gather
 * (optional, may be NULL): device time of the NCCL gather on devices[0]. */
int sgpu_group_batch_search(SgpuGroup* group, const SgpuQueryBatch* queries, const SgpuSearchParams* params,
                            uint64_t* out_ids, float* out_scores, uint32_t* out_counts, SgpuSearchStats* stats,
                            float* ms_gather);

/* Page-locked host memory for query / result buffers (cudaHostAlloc); free with sgpu_host_free. */
int sgpu_host_alloc(uint64_t bytes, void** out);
void sgpu_host_free(void* p);

const char* sgpu_last_error(void);
/* "seismic_b200 <version> sm_100a" */
const char* sgpu_version(void);

/* ------------------------------------------------------------------------------------------
 * Host side (stand-in for the Rust host code; CPU only, no CUDA needed)
 * ------------------------------------------------------------------------------------------ */

typedef struct ShostDataset ShostDataset; /* sparse dataset: CSR, u32 comps, f32 values          */
typedef struct ShostIndex ShostIndex;     /* logical Seismic index (owns or mmaps its arrays)    */

/* Configuration (reference src/configurations.rs:15-129; Python defaults src/pylib/mod.rs:329). */
typedef struct ShostBuildConfig {
    uint32_t pruning;            /* 0 GlobalThreshold, 1 FixedSize                               */
    uint32_t n_postings;         /* 3500                                                         */
    float max_fraction;          /* 1.5                                                          */
    uint32_t blocking;           /* 0 RandomKmeans(InvertedIndexApprox), 1 FixedSize             */
    float centroid_fraction;     /* 0.1                                                          */
    uint32_t min_cluster_size;   /* 2                                                            */
    uint32_t doc_cut;            /* 15                                                           */
    uint32_t block_size;         /* FixedSize blocking                                           */
    uint32_t summarization;      /* 0 EnergyPreserving, 1 FixedSize                              */
    float summary_energy;        /* 0.4                                                          */
    uint32_t n_components;       /* FixedSize summaries                                          */
    uint32_t comp_bits;          /* 16 | 32 (SeismicIndex vs SeismicIndexLV)                     */
    uint32_t value_kind;         /* SGPU_VAL_*: encoding of the forward index                    */
    uint32_t n_threads;          /* 0 = all cores                                                */
    uint64_t kmeans_seed;        /* reference uses 1142 for every list (src/utils.rs:163)        */
} ShostBuildConfig;
void shost_default_config(ShostBuildConfig* cfg);

/* datasets */
int shost_dataset_create(uint64_t n_vecs, uint64_t dim, const uint64_t* offsets, const uint32_t* comps,
                         const float* values, ShostDataset** out); /* copies; sorts nothing       */
int shost_dataset_read_bin(const char* path, ShostDataset** out);  /* seismic inner format        */
int shost_dataset_write_bin(const ShostDataset* ds, const char* path);
void shost_dataset_destroy(ShostDataset* ds);
uint64_t shost_dataset_len(const ShostDataset* ds);
uint64_t shost_dataset_dim(const ShostDataset* ds);
uint64_t shost_dataset_nnz(const ShostDataset* ds);
const uint64_t* shost_dataset_offsets(const ShostDataset* ds);
const uint32_t* shost_dataset_comps(const ShostDataset* ds);
const float* shost_dataset_values(const ShostDataset* ds);

/* Synthetic SPLADE-shaped corpus / query generator (deterministic, counter-based RNG).
 * kind 0 = documents, 1 = queries derived from the documents of the same (n_docs, dim, seed). */
typedef struct ShostSynthConfig {
    uint64_t n_docs;        /* corpus size the topic model is defined over                       */
    uint64_t dim;           /* vocabulary                                                        */
    uint64_t seed;
    uint32_t n_topics;      /* 4096                                                              */
    uint32_t topic_terms;   /* 2000                                                              */
    float doc_nnz_mean;     /* 115 (median of the log-normal), clipped to [16,400]               */
    float doc_nnz_sigma;    /* 0.30                                                              */
    float query_nnz_mean;   /* 38, clipped to [4,128]                                            */
    float query_nnz_sigma;  /* 0.35                                                              */
    uint32_t n_threads;
    uint32_t reserved;
} ShostSynthConfig;
void shost_default_synth(ShostSynthConfig* cfg);
int shost_synth_documents(const ShostSynthConfig* cfg, ShostDataset** out);
int shost_synth_queries(const ShostSynthConfig* cfg, uint64_t n_queries, ShostDataset** out);

/* index build / persistence */
int shost_index_build(const ShostDataset* ds, const ShostBuildConfig* cfg, ShostIndex** out);
/* u16/f16 index -> same posting lists over a DotVByte forward index (gap-coded components, u8 values);
 * reference: src/pylib/dotvbyte.rs:195-213 + convert_dataset_from, src/inverted_index.rs:237-275 */
int shost_index_convert_dotvbyte(const ShostIndex* idx, ShostIndex** out);
int shost_index_save(const ShostIndex* idx, const char* path);
int shost_index_load(const char* path, ShostIndex** out); /* mmap, zero copy                      */
void shost_index_destroy(ShostIndex* idx);
int shost_index_view(const ShostIndex* idx, SgpuIndexView* out); /* borrowed pointers             */
uint64_t shost_index_nnz(const ShostIndex* idx);
/* space report in the format of print_space_usage_byte (src/inverted_index.rs:103-149);
 * fills bytes[0..5] = forward, packed_postings, block_offsets, summaries, knn, total           */
int shost_index_space_usage(const ShostIndex* idx, uint64_t bytes[6]);
/* decoded doc `id` (components + f32 values) of the forward index: SeismicIndex.get(id)        */
int shost_index_get_doc(const ShostIndex* idx, uint64_t id, uint32_t* comps, float* values, uint32_t cap,
                        uint32_t* nnz);

#ifdef __cplusplus
}
#endif
#endif /* SEISMIC_B200_H */
