// Device-side views of the HBM image, the query batch and the per-batch scratch (layout: kernels.cuh header).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef SGPU_LD256
#define SGPU_LD256 1  // u16 x 16-bit-value records: one 256-bit load per 32-byte chunk (0: two 128-bit loads, swapped layout)
#endif

namespace sgpu {

struct ListHdr {
    uint64_t post_base;  // into postings
    uint64_t ent_base;   // into ent_blk / ent_code
    uint64_t sc_base;    // into sc_comp ; run offsets start at sc_base + list id
    uint64_t blk_base;   // into blk_min / blk_quant ; blk_post_off starts at blk_base + list id
    uint32_t n_blk;
    uint32_t n_sc;
    uint32_t n_post;
    uint32_t skip_base;  // into sc_skip: ceil(n_sc / 32) entries, the LAST summary component of every group of 32
};

struct DevIndex {
    const ListHdr* lists;
    const uint64_t* postings;
    const uint32_t* blk_post_off;
    const float* blk_min;
    const float* blk_quant;
    const uint32_t* sc_comp;
    const uint32_t* sc_run_off;
    const uint32_t* sc_skip;    // per list: last component id of every 32 summary components (directory of sc_comp)
    const uint16_t* ent_blk;
    const uint8_t* ent_code;
    const uint4* fwd;           // record buffer, 2 x uint4 per chunk
    const uint32_t* rec_start;  // [n_docs+1]
    uint64_t n_docs;
    uint32_t dim;
    uint32_t comp32;  // 1: u32 components (Rec32 records, 16-byte units), 0: u16 components (Rec16, 32-byte units)
    uint32_t vbyte;   // 1: DotVByte byte stream (4-byte units), u16 components
    uint32_t value_kind;  // SGPU_VAL_* of the records
    float value_scale;
    const uint64_t* knn_posts;  // [n_docs * knn_dim] neighbours as postings (record start << 16 | padded nnz), ~0 = none
    uint32_t knn_dim;
    uint32_t rec_chunk_units;   // units of rec_start per 8-component chunk (plain layouts)
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
    uint4* sel;          // [nq_chunk * est_stride] first list in search order: {estimate bits, first posting, postings, block}
    uint32_t* counters;  // [0] dense work counter, [1] max nterms, [2] invalid queries, [3] hq work counter,
                         // [4] number of hq queries, [5] number of dense queries
    uint32_t* hmult;     // [nq] perfect-hash multiplier of the query (0: dense kernel)
    uint32_t* cost;      // [nq] scheduling cost proxy (postings of the query's lists)
    uint32_t* qlist_hq;  // [nq] chunk-relative ids of the queries taken by the hash-query kernel
    uint32_t* qlist_dense;
    uint32_t* out_keys;  // [nq_chunk * k]
    unsigned long long* stats;  // [0..3] docs_scored, blocks_scored, blocks_pushed, fwd_units; [4..9] phase clocks
    uint32_t est_stride;
    uint32_t cut_eff;
};

__device__ __forceinline__ uint32_t total_key(float f) {  // f32::total_cmp as unsigned key
    uint32_t x = __float_as_uint(f);
    return (x & 0x80000000u) ? ~x : (x | 0x80000000u);
}

}  // namespace sgpu
