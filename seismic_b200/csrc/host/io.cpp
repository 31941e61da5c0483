// Index file I/O (flat, mmap-able) and the "seismic inner format" dataset reader/writer
// (reference scripts/convert_json_to_inner_format.py:10-27: u32 n_vecs; per vector u32 nnz,
// nnz x u32 components, nnz x f32 values; little endian).
//
// The reference persists indexes through vectorium's IndexSerializer (byte format not in the
// tree).  We keep the LOGICAL layout (same arrays, same packed-posting encoding,
// same u8 summaries with per-summary min/quant) in this documented container:
//
//   0   char[8]  "SEISB200"
//   8   u32 version (=1) | u32 comp_bits | u32 value_kind | f32 value_scale
//   24  u64 n_docs | u64 dim | u64 nnz
//   48  ShostBuildConfig (64 bytes, zero padded)
//   112 u32 n_sections (=SEC_COUNT) | u32 pad
//   120 n_sections x { u64 file_offset, u64 bytes }      (section order = enum SectionId)
//   ... sections, each aligned to 64 bytes
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "index.hpp"

ShostIndex::~ShostIndex() {
    if (map_base) munmap(map_base, map_len);
}

namespace shost {

static const char kMagic[8] = {'S', 'E', 'I', 'S', 'B', '2', '0', '0'};
static const uint64_t kHeaderFixed = 120;

int save_index(const ShostIndex& idx, const char* path) {
    // The sections of a loaded index point into an mmap of its file, possibly `path` itself: write a temporary file in
    // the same directory and rename it over the target, so re-saving over the source is safe (as it is for the
    // reference, which deserialises into owned memory) and a failed save never leaves a truncated index behind.
    const std::string tmp = std::string(path) + ".tmp" + std::to_string((long)getpid());
    FILE* fp = std::fopen(tmp.c_str(), "wb");
    if (!fp) { set_error(std::string("cannot open for writing: ") + path); return SGPU_EIO; }
    std::vector<uint8_t> hdr(kHeaderFixed + SEC_COUNT * 16, 0);
    std::memcpy(hdr.data(), kMagic, 8);
    uint32_t version = 1;
    std::memcpy(&hdr[8], &version, 4);
    std::memcpy(&hdr[12], &idx.comp_bits, 4);
    std::memcpy(&hdr[16], &idx.value_kind, 4);
    std::memcpy(&hdr[20], &idx.value_scale, 4);
    std::memcpy(&hdr[24], &idx.n_docs, 8);
    std::memcpy(&hdr[32], &idx.dim, 8);
    std::memcpy(&hdr[40], &idx.nnz, 8);
    static_assert(sizeof(ShostBuildConfig) <= 64, "config grew past its header slot");
    std::memcpy(&hdr[48], &idx.config, sizeof(ShostBuildConfig));
    uint32_t nsec = SEC_COUNT;
    std::memcpy(&hdr[112], &nsec, 4);
    uint64_t pos = (hdr.size() + 63) & ~63ull;
    for (int s = 0; s < SEC_COUNT; ++s) {
        uint64_t bytes = idx.sec[s].bytes;
        std::memcpy(&hdr[kHeaderFixed + s * 16], &pos, 8);
        std::memcpy(&hdr[kHeaderFixed + s * 16 + 8], &bytes, 8);
        pos = (pos + bytes + 63) & ~63ull;
    }
    bool ok = std::fwrite(hdr.data(), 1, hdr.size(), fp) == hdr.size();
    uint64_t cur = hdr.size();
    static const uint8_t zeros[64] = {0};
    for (int s = 0; s < SEC_COUNT && ok; ++s) {
        uint64_t target;
        std::memcpy(&target, &hdr[kHeaderFixed + s * 16], 8);
        if (target > cur) ok = std::fwrite(zeros, 1, target - cur, fp) == target - cur, cur = target;
        if (ok && idx.sec[s].bytes) ok = std::fwrite(idx.sec[s].ptr, 1, idx.sec[s].bytes, fp) == idx.sec[s].bytes;
        cur += idx.sec[s].bytes;
    }
    if (std::fclose(fp) != 0) ok = false;
    if (ok && std::rename(tmp.c_str(), path) != 0) ok = false;
    if (!ok) {
        std::remove(tmp.c_str());
        set_error(std::string("short write: ") + path);
        return SGPU_EIO;
    }
    return SGPU_OK;
}

int load_index(const char* path, ShostIndex** out) {
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) { set_error(std::string("cannot open: ") + path); return SGPU_EIO; }
    struct stat st;
    if (fstat(fd, &st) != 0 || (uint64_t)st.st_size < kHeaderFixed) {
        ::close(fd);
        set_error(std::string("not a seismic_b200 index: ") + path);
        return SGPU_EIO;
    }
    void* base = mmap(nullptr, st.st_size, PROT_READ, MAP_SHARED, fd, 0);
    ::close(fd);
    if (base == MAP_FAILED) { set_error("mmap failed"); return SGPU_EIO; }
    const uint8_t* p = (const uint8_t*)base;
    uint32_t version, nsec;
    std::memcpy(&version, p + 8, 4);
    std::memcpy(&nsec, p + 112, 4);
    if (std::memcmp(p, kMagic, 8) != 0 || version != 1 || nsec != SEC_COUNT ||
        (uint64_t)st.st_size < kHeaderFixed + nsec * 16) {
        munmap(base, st.st_size);
        set_error(std::string("bad magic/version: ") + path);
        return SGPU_EIO;
    }
    auto* idx = new ShostIndex();
    idx->map_base = base;
    idx->map_len = st.st_size;
    std::memcpy(&idx->comp_bits, p + 12, 4);
    std::memcpy(&idx->value_kind, p + 16, 4);
    std::memcpy(&idx->value_scale, p + 20, 4);
    std::memcpy(&idx->n_docs, p + 24, 8);
    std::memcpy(&idx->dim, p + 32, 8);
    std::memcpy(&idx->nnz, p + 40, 8);
    std::memcpy(&idx->config, p + 48, sizeof(ShostBuildConfig));
    for (uint32_t s = 0; s < nsec; ++s) {
        uint64_t off, bytes;
        std::memcpy(&off, p + kHeaderFixed + s * 16, 8);
        std::memcpy(&bytes, p + kHeaderFixed + s * 16 + 8, 8);
        if (bytes > (uint64_t)st.st_size || off > (uint64_t)st.st_size - bytes) {  // no overflow in off + bytes
            delete idx;
            set_error(std::string("truncated index file: ") + path);
            return SGPU_EIO;
        }
        idx->sec[s].ptr = p + off;
        idx->sec[s].bytes = bytes;
    }
    // Cross-check every section length against the header and the start arrays: a corrupt file must raise, not crash.
    auto fail = [&](const char* what) {
        delete idx;
        set_error(std::string("corrupt index file (") + what + "): " + path);
        return SGPU_EIO;
    };
    const uint64_t N = idx->n_docs, D = idx->dim;
    if ((idx->comp_bits != 16 && idx->comp_bits != 32) || idx->value_kind > SGPU_VAL_DOTVBYTE) return fail("encoding");
    if (N >= (1ull << 40) || D == 0 || D > (1ull << 32)) return fail("n_docs / dim");
    auto& S = idx->sec;
    if (S[SEC_FWD_OFFSETS].bytes != (N + 1) * 8) return fail("fwd_offsets");
    for (int sid : {SEC_LIST_POST_START, SEC_LIST_BLK_START, SEC_LIST_SC_START, SEC_LIST_ENT_START})
        if (S[sid].bytes != (D + 1) * 8) return fail("list start arrays");
    const uint64_t* fo = S[SEC_FWD_OFFSETS].as<uint64_t>();
    for (uint64_t d = 0; d < N; ++d)
        if (fo[d] > fo[d + 1]) return fail("fwd_offsets not monotone");
    const bool vb = idx->value_kind == SGPU_VAL_DOTVBYTE;
    const uint64_t vbytes = idx->value_kind == SGPU_VAL_F32 ? 4 : (idx->value_kind == SGPU_VAL_FIXEDU8 ? 1 : 2);
    if (vb) {
        if (S[SEC_FWD_VALUES].bytes < fo[N] || S[SEC_FWD_NNZ].bytes != N * 2) return fail("DotVByte stream");
    } else {
        if (fo[N] != idx->nnz || S[SEC_FWD_COMPS].bytes != fo[N] * (idx->comp_bits / 8) ||
            S[SEC_FWD_VALUES].bytes != fo[N] * vbytes)
            return fail("forward index");
    }
    auto last = [&](int sid) { return S[sid].as<uint64_t>()[D]; };
    auto monotone = [&](int sid) {
        const uint64_t* a = S[sid].as<uint64_t>();
        for (uint64_t l = 0; l < D; ++l)
            if (a[l] > a[l + 1]) return false;
        return a[0] == 0;
    };
    for (int sid : {SEC_LIST_POST_START, SEC_LIST_BLK_START, SEC_LIST_SC_START, SEC_LIST_ENT_START})
        if (!monotone(sid)) return fail("list start arrays not monotone");
    const uint64_t P = last(SEC_LIST_POST_START), TB = last(SEC_LIST_BLK_START), TSC = last(SEC_LIST_SC_START),
                   TE = last(SEC_LIST_ENT_START);
    if (S[SEC_POSTINGS].bytes != P * 8 || S[SEC_BLK_POST_OFF].bytes != (TB + D) * 4 || S[SEC_BLK_MIN].bytes != TB * 4 ||
        S[SEC_BLK_QUANT].bytes != TB * 4 || S[SEC_SC_COMP].bytes != TSC * 4 || S[SEC_SC_RUN_OFF].bytes != (TSC + D) * 4 ||
        S[SEC_ENT_BLK].bytes != TE * 2 || S[SEC_ENT_CODE].bytes != TE)
        return fail("posting list sections");
    *out = idx;
    return SGPU_OK;
}

int read_bin(const char* path, ShostDataset** out) {
    FILE* fp = std::fopen(path, "rb");
    if (!fp) { set_error(std::string("cannot open: ") + path); return SGPU_EIO; }
    uint32_t n = 0;
    if (std::fread(&n, 4, 1, fp) != 1) { std::fclose(fp); set_error("empty file"); return SGPU_EIO; }
    auto* ds = new ShostDataset();
    ds->n_vecs = n;
    ds->offsets.assign(1, 0);
    ds->offsets.reserve((size_t)n + 1);
    uint64_t dim = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t nnz;
        if (std::fread(&nnz, 4, 1, fp) != 1) { std::fclose(fp); delete ds; set_error("truncated .bin"); return SGPU_EIO; }
        size_t at = ds->comps.size();
        ds->comps.resize(at + nnz);
        ds->values.resize(at + nnz);
        if (nnz && (std::fread(&ds->comps[at], 4, nnz, fp) != nnz || std::fread(&ds->values[at], 4, nnz, fp) != nnz)) {
            std::fclose(fp); delete ds; set_error("truncated .bin"); return SGPU_EIO;
        }
        for (uint32_t j = 0; j < nnz; ++j) dim = std::max<uint64_t>(dim, (uint64_t)ds->comps[at + j] + 1);
        ds->offsets.push_back(at + nnz);
    }
    std::fclose(fp);
    ds->dim = dim;
    *out = ds;
    return SGPU_OK;
}

int write_bin(const ShostDataset& ds, const char* path) {
    FILE* fp = std::fopen(path, "wb");
    if (!fp) { set_error(std::string("cannot open for writing: ") + path); return SGPU_EIO; }
    uint32_t n = (uint32_t)ds.n_vecs;
    bool ok = std::fwrite(&n, 4, 1, fp) == 1;
    for (uint64_t i = 0; i < ds.n_vecs && ok; ++i) {
        uint32_t nnz = (uint32_t)(ds.offsets[i + 1] - ds.offsets[i]);
        ok = std::fwrite(&nnz, 4, 1, fp) == 1;
        if (nnz && ok)
            ok = std::fwrite(&ds.comps[ds.offsets[i]], 4, nnz, fp) == nnz &&
                 std::fwrite(&ds.values[ds.offsets[i]], 4, nnz, fp) == nnz;
    }
    if (std::fclose(fp) != 0) ok = false;
    if (!ok) { set_error("short write"); return SGPU_EIO; }
    return SGPU_OK;
}

}  // namespace shost
