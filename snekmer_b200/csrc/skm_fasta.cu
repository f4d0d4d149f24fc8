// skm_fasta.cu — FASTA ingest on the host cores (no device code): text -> packed residues + offsets + ids,
// the input side of kernel (a).  Replaces the per-record Python objects of Bio.SeqIO.parse(f, "fasta") as the
// vectorize rule uses them (kmerize.smk:90-129: f.id, f.seq, len(f.seq)):
//   record  starts at a '>' in the first column; text before the first record is skipped;
//   id      the title (rest of the '>' line) up to its first whitespace ("" for an empty title);
//   seq     the following lines, each stripped of trailing whitespace, joined, blanks and CR removed.
// Two passes over the text, both split over `threads` byte ranges: a thread owns the records whose '>' lies in
// its range and follows its last record into the next range.  Pass 1 counts (records, residues, id bytes) per
// range; the prefix sums give every range its output position; pass 2 writes.  The output is laid out exactly
// like the device batch (residues back to back, int64 offsets), so it can be read straight into pinned memory.
#include <cstring>
#include <thread>
#include <vector>

#include "skm_common.cuh"

namespace skm {

struct FastaCount {
    int64_t nseq = 0, nres = 0, idbytes = 0;
};

static inline bool is_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// first record start (a '>' in column 0) at or after p, or n
static int64_t next_record(const uint8_t *t, int64_t n, int64_t p) {
    while (p < n) {
        if (t[p] == '>' && (p == 0 || t[p - 1] == '\n')) return p;
        const void *nl = memchr(t + p, '\n', size_t(n - p));
        if (!nl) return n;
        p = int64_t((const uint8_t *)nl - t) + 1;
    }
    return n;
}

// Walk the records whose start lies in [lo, hi).  WRITE = false: count only.
template <bool WRITE>
static FastaCount fasta_range(const uint8_t *t, int64_t n, int64_t lo, int64_t hi, uint8_t *res, int64_t *off,
                              uint8_t *ids, int64_t *id_off, int64_t seq0, int64_t res0, int64_t id0) {
    FastaCount c;
    int64_t p = next_record(t, n, lo);
    while (p < hi && p < n) {
        // ---- title line ----
        int64_t e = p + 1;
        const void *nl = memchr(t + e, '\n', size_t(n - e));
        const int64_t line_end = nl ? int64_t((const uint8_t *)nl - t) : n;
        int64_t a = e;
        while (a < line_end && is_space(t[a])) ++a;          // title.split(None, 1)[0] skips leading blanks
        int64_t b = a;
        while (b < line_end && !is_space(t[b])) ++b;
        if (WRITE) {
            id_off[seq0 + c.nseq] = id0 + c.idbytes;
            off[seq0 + c.nseq] = res0 + c.nres;
            memcpy(ids + id0 + c.idbytes, t + a, size_t(b - a));
        }
        c.idbytes += b - a;
        // ---- sequence lines up to the next record ----
        int64_t q = line_end < n ? line_end + 1 : n;
        while (q < n && t[q] != '>') {
            const void *nl2 = memchr(t + q, '\n', size_t(n - q));
            const int64_t le = nl2 ? int64_t((const uint8_t *)nl2 - t) : n;
            int64_t r = le;
            while (r > q && is_space(t[r - 1])) --r;         // line.rstrip()
            if (WRITE) {
                uint8_t *dst = res + res0 + c.nres;
                int64_t w = 0;
                for (int64_t i = q; i < r; ++i) {
                    const uint8_t ch = t[i];
                    if (ch != ' ' && ch != '\r') dst[w++] = ch;
                }
                c.nres += w;
            } else {
                int64_t w = r - q;
                for (int64_t i = q; i < r; ++i) w -= (t[i] == ' ' || t[i] == '\r');
                c.nres += w;
            }
            q = le < n ? le + 1 : n;
        }
        ++c.nseq;
        p = q;
    }
    return c;
}

static int fasta_threads(int threads, int64_t nbytes) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    const int64_t by_size = nbytes / (1 << 20) + 1;          // >= 1 MiB of text per thread
    if (threads > by_size) threads = (int)by_size;
    return threads > 256 ? 256 : threads;
}

template <typename Fn>
static void run_ranges(int threads, Fn &&fn) {
    if (threads == 1) { fn(0); return; }
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int i = 0; i < threads; ++i) pool.emplace_back([&fn, i] { fn(i); });
    for (auto &th : pool) th.join();
}

}  // namespace skm

extern "C" {

int skm_fasta_scan(const uint8_t *text, int64_t nbytes, int threads, int64_t *nseq_out, int64_t *nres_out,
                   int64_t *idbytes_out) {
    using namespace skm;
    if (nbytes < 0 || (nbytes > 0 && !text) || !nseq_out || !nres_out || !idbytes_out) { set_error("skm_fasta_scan: bad arguments"); return SKM_ERR_INVALID; }
    const int T = fasta_threads(threads, nbytes);
    std::vector<FastaCount> cnt(T);
    run_ranges(T, [&](int i) {
        const int64_t lo = nbytes * i / T, hi = nbytes * (i + 1) / T;
        cnt[i] = fasta_range<false>(text, nbytes, lo, hi, nullptr, nullptr, nullptr, nullptr, 0, 0, 0);
    });
    int64_t a = 0, b = 0, c = 0;
    for (const auto &x : cnt) { a += x.nseq; b += x.nres; c += x.idbytes; }
    *nseq_out = a; *nres_out = b; *idbytes_out = c;
    return SKM_OK;
}

int skm_fasta_pack(const uint8_t *text, int64_t nbytes, int threads, uint8_t *residues_out, int64_t *offsets_out,
                   uint8_t *ids_out, int64_t *id_offsets_out) {
    using namespace skm;
    if (nbytes < 0 || (nbytes > 0 && !text) || !offsets_out || !id_offsets_out) { set_error("skm_fasta_pack: bad arguments"); return SKM_ERR_INVALID; }
    const int T = fasta_threads(threads, nbytes);
    std::vector<FastaCount> cnt(T);
    run_ranges(T, [&](int i) {
        const int64_t lo = nbytes * i / T, hi = nbytes * (i + 1) / T;
        cnt[i] = fasta_range<false>(text, nbytes, lo, hi, nullptr, nullptr, nullptr, nullptr, 0, 0, 0);
    });
    std::vector<FastaCount> base(T + 1);
    for (int i = 0; i < T; ++i) {
        base[i + 1].nseq = base[i].nseq + cnt[i].nseq;
        base[i + 1].nres = base[i].nres + cnt[i].nres;
        base[i + 1].idbytes = base[i].idbytes + cnt[i].idbytes;
    }
    if ((base[T].nres > 0 && !residues_out) || (base[T].idbytes > 0 && !ids_out)) { set_error("skm_fasta_pack: NULL output buffer"); return SKM_ERR_INVALID; }
    run_ranges(T, [&](int i) {
        const int64_t lo = nbytes * i / T, hi = nbytes * (i + 1) / T;
        fasta_range<true>(text, nbytes, lo, hi, residues_out, offsets_out, ids_out, id_offsets_out, base[i].nseq, base[i].nres, base[i].idbytes);
    });
    offsets_out[base[T].nseq] = base[T].nres;
    id_offsets_out[base[T].nseq] = base[T].idbytes;
    return SKM_OK;
}

}  // extern "C"
