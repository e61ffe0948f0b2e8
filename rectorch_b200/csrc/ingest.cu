// ingest.cu -- CSV rating files -> canonical CSR (host side of the data path, no pandas).
//
// Replaces pd.read_csv + scipy.sparse.csr_matrix((values, (rows, cols))) of DataReader
// (rectorch/data.py:363-409): the pre-processed files train.csv / {validation,test}_{tr,te}.csv have a
// header line "uid,iid[,<value column>,...]" and one rating per line.  The file is mapped, cut into
// chunks at line boundaries and parsed by a pool of host threads; the COO triples (kept per chunk, never
// concatenated) are turned into CSR by a stable parallel counting sort over the rows (per-thread row
// histograms -> exclusive offsets -> scatter) followed by a parallel per-row column sort in which duplicate
// (row, col) entries are summed in file order -- the canonical form scipy produces (coo -> csr ->
// sum_duplicates).  Values stay float64, as the reference's matrices are (data.py:377); DeviceCSR converts to
// fp32 when it uploads.
//
// This file is plain host C++ (no kernels): text parsing is host work by nature.  The CSR it produces is
// what b200vae_bind_csr receives after the upload.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cstdlib>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b200 {

struct CsvChunk {
    std::vector<int64_t> uid;
    std::vector<int64_t> iid;
    std::vector<double> val;
    int64_t bad_line = -1;      // chunk-relative index of the first malformed line
    int64_t lines = 0;
    int64_t uid_min = INT64_MAX, uid_max = -1, iid_min = INT64_MAX, iid_max = -1;
};

struct Csv {
    std::vector<CsvChunk> chunks;   // in file order
    int64_t n_records = 0;
    int ncols = 0;
    int n_threads = 1;
    std::string third_name;
    int64_t uid_min = 0, uid_max = -1, iid_max = -1;
};

// run fn(t) for t in [0, n) on up to `threads` host threads (dynamic assignment)
template <typename F>
static void parallel_for(int n, int threads, F fn) {
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) {
        for (int t = 0; t < n; ++t) fn(t);
        return;
    }
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (int t = next.fetch_add(1); t < n; t = next.fetch_add(1)) fn(t);
    };
    std::vector<std::thread> th;
    for (int k = 1; k < threads; ++k) th.emplace_back(worker);
    worker();
    for (auto& x : th) x.join();
}

static inline const char* skip_blank(const char* p, const char* e) {
    while (p < e && (*p == ' ' || *p == '\t')) ++p;
    return p;
}

// one integer field (pandas writes ints; "3.0" style floats with a zero fraction are accepted too)
static inline bool parse_i64(const char*& p, const char* e, int64_t& out) {
    p = skip_blank(p, e);
    auto r = std::from_chars(p, e, out);
    if (r.ec != std::errc()) return false;
    p = r.ptr;
    if (p < e && *p == '.') {
        ++p;
        while (p < e && *p == '0') ++p;
        if (p < e && *p >= '1' && *p <= '9') return false;
    }
    p = skip_blank(p, e);
    return true;
}

static inline bool parse_f64(const char*& p, const char* e, double& out) {
    p = skip_blank(p, e);
    auto r = std::from_chars(p, e, out);
    if (r.ec != std::errc()) return false;
    p = skip_blank(r.ptr, e);
    return true;
}

// parse the lines of [b, e): "uid,iid" or "uid,iid,value[,...]"; blank lines are skipped (as pandas does)
static void parse_chunk(const char* b, const char* e, char sep, bool want_val, CsvChunk* out) {
    const size_t guess = (size_t)(e - b) / 8 + 16;      // "uid,iid\n" is rarely shorter than 8 bytes
    out->uid.reserve(guess);
    out->iid.reserve(guess);
    if (want_val) out->val.reserve(guess);
    const char* p = b;
    while (p < e) {
        const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(e - p)));
        const char* le = nl ? nl : e;
        const char* next = nl ? nl + 1 : e;
        if (le > p && le[-1] == '\r') --le;
        const char* q = skip_blank(p, le);
        if (q < le) {
            int64_t u, i;
            double v = 1.0;
            bool ok = parse_i64(q, le, u) && q < le && *q == sep;
            if (ok) { ++q; ok = parse_i64(q, le, i); }
            if (ok && want_val) {
                ok = q < le && *q == sep;
                if (ok) { ++q; ok = parse_f64(q, le, v); }
                ok = ok && (q == le || *q == sep);
            } else if (ok) {
                ok = (q == le || *q == sep);
            }
            if (!ok) {
                if (out->bad_line < 0) out->bad_line = out->lines;
            } else {
                out->uid.push_back(u);
                out->iid.push_back(i);
                if (want_val) out->val.push_back(v);
                out->uid_min = std::min(out->uid_min, u); out->uid_max = std::max(out->uid_max, u);
                out->iid_min = std::min(out->iid_min, i); out->iid_max = std::max(out->iid_max, i);
            }
            out->lines++;
        }
        p = next;
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200vae_csv_open(b200vae_csv** out, const char* path, char sep, int n_threads) {
    B200_REQUIRE(out && path, B200VAE_EINVAL, "null argument");
    *out = nullptr;
    if (sep == 0) sep = ',';
    int fd = open(path, O_RDONLY);
    B200_REQUIRE(fd >= 0, B200VAE_EINVAL, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); set_error("cannot stat %s", path); return B200VAE_EINVAL; }
    const size_t size = (size_t)st.st_size;
    const char* base = nullptr;
    if (size > 0) {
        void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); set_error("cannot map %s", path); return B200VAE_EINVAL; }
        base = static_cast<const char*>(m);
        madvise(m, size, MADV_SEQUENTIAL);
    }
    close(fd);
    Csv* c = new (std::nothrow) Csv();
    if (!c) { if (base) munmap(const_cast<char*>(base), size); set_error("out of host memory"); return B200VAE_EINVAL; }
    auto fail = [&](int code) { if (base) munmap(const_cast<char*>(base), size); delete c; return code; };
    // ---- header: column names ----
    const char* end = base + size;
    const char* nl = size ? static_cast<const char*>(memchr(base, '\n', size)) : nullptr;
    const char* hend = nl ? nl : end;
    const char* body = nl ? nl + 1 : end;
    if (hend > base && hend[-1] == '\r') --hend;
    if (hend == base) { set_error("%s: missing header line", path); return fail(B200VAE_EINVAL); }
    {
        int col = 0;
        const char* f = base;
        for (const char* p = base; p <= hend; ++p) {
            if (p == hend || *p == sep) {
                if (col == 2) c->third_name.assign(f, p);
                ++col;
                f = p + 1;
            }
        }
        c->ncols = col;
    }
    if (c->ncols < 2) { set_error("%s: expected at least the columns uid,iid (separator '%c')", path, sep); return fail(B200VAE_EINVAL); }
    const bool want_val = c->ncols >= 3;
    // ---- body: ~1 MB chunks cut at line boundaries, parsed by a pool of threads ----
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, 64));
    c->n_threads = nt;
    const size_t body_size = (size_t)(end - body);
    const int n_chunks = (int)std::max<size_t>(1, std::min<size_t>(4096, body_size / (1 << 20) + 1));
    std::vector<const char*> cut(n_chunks + 1);
    cut[0] = body;
    cut[n_chunks] = end;
    for (int t = 1; t < n_chunks; ++t) {
        const char* p = body + body_size * (size_t)t / (size_t)n_chunks;
        if (p < cut[t - 1]) p = cut[t - 1];
        const char* q = p < end ? static_cast<const char*>(memchr(p, '\n', (size_t)(end - p))) : nullptr;
        cut[t] = q ? q + 1 : end;
    }
    c->chunks.resize(n_chunks);
    parallel_for(n_chunks, nt, [&](int t) { parse_chunk(cut[t], cut[t + 1], sep, want_val, &c->chunks[t]); });
    int64_t line0 = 2, total = 0;     // 1-based file line of the first body line
    int64_t umin = INT64_MAX, umax = -1, imin = INT64_MAX, imax = -1;
    for (int t = 0; t < n_chunks; ++t) {
        const CsvChunk& ch = c->chunks[t];
        if (ch.bad_line >= 0) {
            set_error("%s: malformed record near data line %lld", path, (long long)(line0 + ch.bad_line));
            return fail(B200VAE_EINVAL);
        }
        line0 += ch.lines;
        total += (int64_t)ch.uid.size();
        umin = std::min(umin, ch.uid_min); umax = std::max(umax, ch.uid_max);
        imin = std::min(imin, ch.iid_min); imax = std::max(imax, ch.iid_max);
    }
    if (base) munmap(const_cast<char*>(base), size);
    base = nullptr;
    c->n_records = total;
    if (total > 0) {
        if (umin < 0 || imin < 0) { set_error("%s: negative ids", path); delete c; return B200VAE_EINVAL; }
        c->uid_min = umin; c->uid_max = umax; c->iid_max = imax;
    }
    *out = reinterpret_cast<b200vae_csv*>(c);
    return 0;
}

int b200vae_csv_info(const b200vae_csv* h, int64_t* n_records, int32_t* n_cols, int64_t* uid_min, int64_t* uid_max,
                     int64_t* iid_max) {
    const Csv* c = reinterpret_cast<const Csv*>(h);
    B200_REQUIRE(c, B200VAE_EINVAL, "null handle");
    if (n_records) *n_records = c->n_records;
    if (n_cols) *n_cols = c->ncols;
    if (uid_min) *uid_min = c->uid_min;
    if (uid_max) *uid_max = c->uid_max;
    if (iid_max) *iid_max = c->iid_max;
    return 0;
}

const char* b200vae_csv_value_column(const b200vae_csv* h) {
    const Csv* c = reinterpret_cast<const Csv*>(h);
    return c ? c->third_name.c_str() : "";
}

int b200vae_csv_to_csr(const b200vae_csv* h, int64_t uid_base, int64_t n_rows, int32_t n_cols, int use_values,
                       int64_t* indptr_host, int32_t* indices_host, double* values_host, int64_t* nnz_out) {
    const Csv* c = reinterpret_cast<const Csv*>(h);
    B200_REQUIRE(c && indptr_host && indices_host && values_host && nnz_out, B200VAE_EINVAL, "null argument");
    B200_REQUIRE(n_rows >= 0 && n_cols >= 0, B200VAE_EINVAL, "negative shape");
    B200_REQUIRE(!use_values || c->ncols >= 3, B200VAE_EINVAL, "the file has no value column");
    const int64_t n = c->n_records;
    if (n > 0) {      // ids are bounded by the per-file extrema gathered while parsing
        B200_REQUIRE(c->uid_min - uid_base >= 0 && c->uid_max - uid_base < n_rows, B200VAE_EINVAL,
                     "row index %lld outside [0, %lld)",
                     (long long)(c->uid_min - uid_base < 0 ? c->uid_min - uid_base : c->uid_max - uid_base), (long long)n_rows);
        B200_REQUIRE(c->iid_max < n_cols, B200VAE_EINVAL, "column index %lld outside [0, %d)", (long long)c->iid_max, n_cols);
    }
    const int nch = (int)c->chunks.size();
    // ---- stable parallel counting sort by row: the chunks are dealt to P groups of consecutive chunks ----
    const int P = std::max(1, std::min({c->n_threads, nch, (int)std::max<int64_t>(1, (int64_t)(1 << 26) / std::max<int64_t>(n_rows, 1))}));
    std::vector<int> g_lo(P + 1);
    for (int g = 0; g <= P; ++g) g_lo[g] = (int)((int64_t)nch * g / P);
    std::vector<std::vector<int64_t>> hist(P);
    parallel_for(P, P, [&](int g) {
        hist[g].assign((size_t)n_rows + 1, 0);
        for (int t = g_lo[g]; t < g_lo[g + 1]; ++t)
            for (int64_t u : c->chunks[t].uid) hist[g][(size_t)(u - uid_base)]++;
    });
    // exclusive offsets: start[r] = records of rows < r; hist[g][r] becomes the first slot of group g in row r
    std::vector<int64_t> start((size_t)n_rows + 1, 0);
    {
        int64_t run = 0;
        for (int64_t r = 0; r < n_rows; ++r) {
            start[(size_t)r] = run;
            for (int g = 0; g < P; ++g) { const int64_t k = hist[g][(size_t)r]; hist[g][(size_t)r] = run; run += k; }
        }
        start[(size_t)n_rows] = run;
    }
    std::vector<int32_t> col((size_t)n);
    std::vector<double> val((size_t)n);
    parallel_for(P, P, [&](int g) {
        std::vector<int64_t>& cur = hist[g];
        for (int t = g_lo[g]; t < g_lo[g + 1]; ++t) {
            const CsvChunk& ch = c->chunks[t];
            const size_t m = ch.uid.size();
            for (size_t k = 0; k < m; ++k) {
                const int64_t p = cur[(size_t)(ch.uid[k] - uid_base)]++;
                col[(size_t)p] = (int32_t)ch.iid[k];
                val[(size_t)p] = use_values ? ch.val[k] : 1.0;
            }
        }
    });
    // ---- per row: stable sort by column, duplicates summed in file order (in place), unique count ----
    const int R = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)c->n_threads * 8, n_rows));
    std::vector<int64_t> uniq((size_t)n_rows + 1, 0);
    parallel_for(R, c->n_threads, [&](int part) {
        std::vector<std::pair<int32_t, double>> tmp;
        const int64_t r0 = n_rows * part / R, r1 = n_rows * (part + 1) / R;
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t a = start[(size_t)r], b = start[(size_t)r + 1];
            bool sorted = true;
            for (int64_t k = a + 1; k < b && sorted; ++k) sorted = col[(size_t)k - 1] < col[(size_t)k];
            if (sorted) { uniq[(size_t)r] = b - a; continue; }
            tmp.clear();
            for (int64_t k = a; k < b; ++k) tmp.emplace_back(col[(size_t)k], val[(size_t)k]);
            std::stable_sort(tmp.begin(), tmp.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
            int64_t w = a;
            for (size_t k = 0; k < tmp.size(); ++k) {
                if (k > 0 && tmp[k].first == tmp[k - 1].first) {
                    val[(size_t)w - 1] += tmp[k].second;
                } else {
                    col[(size_t)w] = tmp[k].first;
                    val[(size_t)w] = tmp[k].second;
                    ++w;
                }
            }
            uniq[(size_t)r] = w - a;
        }
    });
    indptr_host[0] = 0;
    for (int64_t r = 0; r < n_rows; ++r) indptr_host[r + 1] = indptr_host[r] + uniq[(size_t)r];
    parallel_for(R, c->n_threads, [&](int part) {
        const int64_t r0 = n_rows * part / R, r1 = n_rows * (part + 1) / R;
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t a = start[(size_t)r], w = indptr_host[r], m = uniq[(size_t)r];
            memcpy(indices_host + w, col.data() + a, (size_t)m * sizeof(int32_t));
            memcpy(values_host + w, val.data() + a, (size_t)m * sizeof(double));
        }
    });
    *nnz_out = indptr_host[n_rows];
    return 0;
}

int b200vae_csv_close(b200vae_csv* h) {
    delete reinterpret_cast<Csv*>(h);
    return 0;
}

}  // extern "C"
