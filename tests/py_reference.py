"""A second, independent restatement of the reference's search path — pure Python / numpy float32, written from the
reference sources (not from oracle/oracle_search.cpp) and used only to cross-check the C++ oracle on small random
indexes (tests/test_oracle_crosscheck.py).  Slow by design.

  InvertedIndexBase::search        /root/reference/src/inverted_index.rs:153-234
  QuantizedSummary::distances      src/quantized_summary.rs:64-160
  PostingList::search / sort_and_search / evaluate_posting_block   src/posting_list.rs:115-215
  KHeap                            src/utils.rs:12-66
  Knn::refine                      src/inverted_index.rs:551-593
The choices the reference leaves to its absent dependencies are the oracle's: score = 8 partial sums
(element i -> partial (i / 8) % 8, mul then add in f32) reduced as ((p0+p4)+(p2+p6))+((p1+p5)+(p3+p7)); better = higher
score, then smaller forward offset; equal query values: earlier position; equal estimates: smaller block id."""
import numpy as np

F = np.float32
PAD = np.uint64(0xFFFFFFFFFFFFFFFF)


def _key(x):  # f32::total_cmp as an integer key
    b = int(np.float32(x).view(np.uint32))
    return (~b & 0xFFFFFFFF) if b & 0x80000000 else (b | 0x80000000)


class KHeap:
    """Bounded set of the k best items; push replaces the worst only if the new item is strictly better."""

    def __init__(self, k):
        self.k, self.items = k, []  # items: (score f32, start)

    @staticmethod
    def better(a, b):
        return a[0] > b[0] or (a[0] == b[0] and a[1] < b[1])

    def worst(self):
        w = self.items[0]
        for it in self.items[1:]:
            if self.better(w, it):
                w = it
        return w

    def push(self, item):
        if len(self.items) < self.k:
            self.items.append(item)
        else:
            w = self.worst()
            if self.better(item, w):
                self.items[self.items.index(w)] = item

    def sorted(self):
        out = list(self.items)
        for i in range(1, len(out)):  # insertion sort with the same order
            j = i
            while j > 0 and self.better(out[j], out[j - 1]):
                out[j], out[j - 1] = out[j - 1], out[j]
                j -= 1
        return out


class PyIndex:
    def __init__(self, host):
        a = host.arrays()
        self.a = a
        self.n_docs, self.dim = host.len, host.dim
        self.off, self.comps, self.vals = host.forward_csr()
        self.fo = a["fwd_offsets"].astype(np.int64)
        self.knn = host.knn

    def doc_of(self, start):  # id_from_range: the document whose range starts here (the non-empty one)
        return int(np.searchsorted(self.fo, start, side="right")) - 1

    def doc_score(self, q, start, ln):
        p = [F(0)] * 8
        for i in range(ln):
            c = int(self.comps[start + i])
            lane = (i // 8) % 8
            p[lane] = F(p[lane] + F(q[c] * self.vals[start + i]))
        return F(F(F(p[0] + p[4]) + F(p[2] + p[6])) + F(F(p[1] + p[5]) + F(p[3] + p[7])))

    def distances(self, l, qc, qv):
        a = self.a
        b0, b1 = int(a["list_blk_start"][l]), int(a["list_blk_start"][l + 1])
        B = b1 - b0
        est = np.zeros(B, dtype=F)
        mins, quants = a["blk_min"][b0:b1], a["blk_quant"][b0:b1]
        s0, s1 = int(a["list_sc_start"][l]), int(a["list_sc_start"][l + 1])
        sc = a["sc_comp"][s0:s1]
        run = a["sc_run_off"][s0 + l: s1 + l + 1]
        e0 = int(a["list_ent_start"][l])
        eb, ec = a["ent_blk"][e0:], a["ent_code"][e0:]
        for j in range(len(qc)):
            if j > 0 and qc[j] == qc[j - 1]:
                continue
            i = int(np.searchsorted(sc, qc[j]))
            if i == len(sc) or sc[i] != qc[j]:
                continue
            for e in range(int(run[i]), int(run[i + 1])):
                s = int(eb[e])
                deq = F(F(F(ec[e]) * quants[s]) + mins[s])
                est[s] = F(est[s] + F(deq * F(qv[j])))
        return est

    def search(self, qc, qv, k, query_cut, heap_factor, n_knn=0, first_sorted=True):
        a = self.a
        q = np.zeros(self.dim, dtype=F)
        for c, v in zip(qc, qv):
            q[int(c)] = F(v)
        heap, visited = KHeap(k), set()
        terms = sorted(range(len(qc)), key=lambda i: (-_key(qv[i]), i))[: min(query_cut, len(qc))]
        evaluated = 0
        for t, ti in enumerate(terms):
            l = int(qc[ti])
            est = self.distances(l, qc, qv)
            B = len(est)
            order = list(range(B))
            if t == 0 and first_sorted:
                order.sort(key=lambda b: (-_key(est[b]), b))
            boff = a["blk_post_off"][int(a["list_blk_start"][l]) + l:]
            posts = a["postings"][int(a["list_post_start"][l]):]
            for b in order:
                if len(heap.items) == k and est[b] < F(F(heap_factor) * heap.worst()[0]):
                    continue
                evaluated += 1
                for i in range(int(boff[b]), int(boff[b + 1])):
                    start, ln = int(posts[i]) >> 16, int(posts[i]) & 0xFFFF
                    if start not in visited:
                        visited.add(start)
                        heap.push((self.doc_score(q, start, ln), start))
        if n_knn > 0 and self.knn is not None:
            dim_knn = self.knn.shape[1]
            for _, start in heap.sorted():
                d = self.doc_of(start)
                for i in range(min(dim_knn, n_knn)):
                    nb = self.knn[d, i]
                    if nb == PAD:
                        continue
                    s2, e2 = int(self.fo[int(nb)]), int(self.fo[int(nb) + 1])
                    if e2 > s2 and s2 not in visited:
                        visited.add(s2)
                        heap.push((self.doc_score(q, s2, e2 - s2), s2))
        res = heap.sorted()
        ids = [self.doc_of(s) for _, s in res]
        return ids, [float(s) for s, _ in res], evaluated


def decode_dotvbyte(host):
    """Independent decoder of the DotVByte forward index (format: seismic_b200/csrc/host/build.cpp, convert_dotvbyte;
    the reference's own byte format lives in vectorium).  Per document, 16-byte aligned, nch = ceil(nnz / 8) chunks:
    [16 bytes per super-round of 64 chunks: u64 mask of the WIDE chunks, u32 wide chunks before, u32 zero]
    [16 bytes per chunk: 8 low bytes of its gaps | 8 u8 codes][8 bytes per wide chunk: the high bytes of its gaps].
    The gaps are one chain over the record (gap 0 = first component).  Returns CSR (offsets, components,
    values = code * scale, codes)."""
    from seismic_b200 import _native as N
    v = host.view
    n = host.len
    fo = N.np_view(v.fwd_offsets, n + 1, np.uint64).astype(np.int64)
    stream = N.np_view(v.fwd_values, int(fo[-1]), np.uint8)
    nnzs = N.np_view(v.fwd_nnz, n, np.uint16)
    scale = np.float32(v.value_scale)
    off, comps, vals, codes = [0], [], [], []
    for d in range(n):
        assert fo[d] % 16 == 0
        rec = stream[int(fo[d]):int(fo[d + 1])]
        ln = int(nnzs[d])
        nch = (ln + 7) // 8
        ndir = (nch + 63) // 64
        fixed = rec[16 * ndir: 16 * ndir + 16 * nch]
        wide = rec[16 * ndir + 16 * nch:]
        c, n_wide = 0, 0
        for m in range(nch):
            entry = rec[16 * (m // 64): 16 * (m // 64) + 16]
            mask = int(entry[:8].view(np.uint64)[0])
            if m % 64 == 0:
                assert int(entry[8:12].view(np.uint32)[0]) == n_wide and not entry[12:].any()
                assert mask >> min(64, nch - m) == 0, "no mask bits beyond the last chunk"
            fx = fixed[16 * m: 16 * m + 16]
            is_wide = (mask >> (m % 64)) & 1
            hi = wide[8 * n_wide: 8 * n_wide + 8] if is_wide else np.zeros(8, np.uint8)
            n_wide += is_wide
            gaps = [int(fx[f]) | (int(hi[f]) << 8) for f in range(8)]
            assert bool(is_wide) == (max(gaps) >= 256), "a chunk is wide iff one of its gaps needs two bytes"
            for f in range(8):
                c += gaps[f]
                if m * 8 + f < ln:
                    comps.append(c)
                    codes.append(int(fx[8 + f]))
                    vals.append(np.float32(np.float32(fx[8 + f]) * scale))
                else:
                    assert gaps[f] == 0 and fx[8 + f] == 0
        assert not wide[8 * n_wide:].any(), "padding must be zero"
        assert len(rec) == (16 * ndir + 16 * nch + 8 * n_wide + 15) // 16 * 16
        off.append(len(comps))
    return np.array(off, np.uint64), np.array(comps, np.uint32), np.array(vals, np.float32), np.array(codes, np.uint8)
