// ingest.cu -- CSV rating files -> canonical CSR (host side of the data path, no pandas).
//
// Replaces pd.read_csv + scipy.sparse.csr_matrix((values, (rows, cols))) of DataReader
// (rectorch/data.py:363-409): the pre-processed files train.csv / {validation,test}_{tr,te}.csv have a
// header line "uid,iid[,<value column>,...]" and one rating per line.  The file is mapped, cut into
// per-thread chunks at line boundaries and parsed in parallel; the COO triples are turned into CSR by a
// counting sort over the rows followed by a per-row column sort in which duplicate (row, col) entries are
// summed -- the canonical form scipy produces (coo -> csr -> sum_duplicates).  Values stay float64, as the
// reference's matrices are (data.py:377); DeviceCSR converts to fp32 when it uploads.
//
// This file is plain host C++ (no kernels): text parsing is host work by nature.  The CSR it produces is
// what b200vae_bind_csr receives after the upload.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
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
};

struct Csv {
    std::vector<int64_t> uid;
    std::vector<int64_t> iid;
    std::vector<double> val;    // empty when the file has only two columns
    int ncols = 0;
    std::string third_name;
    int64_t uid_min = 0, uid_max = -1, iid_max = -1;
};

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
    // ---- body: chunks cut at line boundaries ----
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, 64));
    const size_t body_size = (size_t)(end - body);
    nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)nt, body_size / (1 << 20) + 1));
    std::vector<const char*> cut(nt + 1);
    cut[0] = body;
    cut[nt] = end;
    for (int t = 1; t < nt; ++t) {
        const char* p = body + body_size * (size_t)t / (size_t)nt;
        if (p < cut[t - 1]) p = cut[t - 1];
        const char* q = p < end ? static_cast<const char*>(memchr(p, '\n', (size_t)(end - p))) : nullptr;
        cut[t] = q ? q + 1 : end;
    }
    std::vector<CsvChunk> chunks(nt);
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(parse_chunk, cut[t], cut[t + 1], sep, want_val, &chunks[t]);
        parse_chunk(cut[0], cut[1], sep, want_val, &chunks[0]);
        for (auto& x : th) x.join();
    }
    int64_t line0 = 2, total = 0;     // 1-based file line of the first body line
    for (int t = 0; t < nt; ++t) {
        if (chunks[t].bad_line >= 0) {
            set_error("%s: malformed record near data line %lld", path, (long long)(line0 + chunks[t].bad_line));
            return fail(B200VAE_EINVAL);
        }
        line0 += chunks[t].lines;
        total += (int64_t)chunks[t].uid.size();
    }
    c->uid.reserve(total);
    c->iid.reserve(total);
    if (want_val) c->val.reserve(total);
    for (auto& ch : chunks) {
        c->uid.insert(c->uid.end(), ch.uid.begin(), ch.uid.end());
        c->iid.insert(c->iid.end(), ch.iid.begin(), ch.iid.end());
        if (want_val) c->val.insert(c->val.end(), ch.val.begin(), ch.val.end());
    }
    if (base) munmap(const_cast<char*>(base), size);
    base = nullptr;
    if (total > 0) {
        c->uid_min = *std::min_element(c->uid.begin(), c->uid.end());
        c->uid_max = *std::max_element(c->uid.begin(), c->uid.end());
        c->iid_max = *std::max_element(c->iid.begin(), c->iid.end());
        const int64_t iid_min = *std::min_element(c->iid.begin(), c->iid.end());
        if (c->uid_min < 0 || iid_min < 0) { set_error("%s: negative ids", path); delete c; return B200VAE_EINVAL; }
    }
    *out = reinterpret_cast<b200vae_csv*>(c);
    return 0;
}

int b200vae_csv_info(const b200vae_csv* h, int64_t* n_records, int32_t* n_cols, int64_t* uid_min, int64_t* uid_max,
                     int64_t* iid_max) {
    const Csv* c = reinterpret_cast<const Csv*>(h);
    B200_REQUIRE(c, B200VAE_EINVAL, "null handle");
    if (n_records) *n_records = (int64_t)c->uid.size();
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
    const int64_t n = (int64_t)c->uid.size();
    for (int64_t k = 0; k < n; ++k) {
        const int64_t r = c->uid[k] - uid_base;
        B200_REQUIRE(r >= 0 && r < n_rows, B200VAE_EINVAL, "row index %lld outside [0, %lld)", (long long)r, (long long)n_rows);
        B200_REQUIRE(c->iid[k] < n_cols, B200VAE_EINVAL, "column index %lld outside [0, %d)", (long long)c->iid[k], n_cols);
    }
    // counting sort by row (stable: records keep file order inside a row)
    std::vector<int64_t> start((size_t)n_rows + 1, 0);
    for (int64_t k = 0; k < n; ++k) start[(size_t)(c->uid[k] - uid_base) + 1]++;
    for (int64_t r = 0; r < n_rows; ++r) start[(size_t)r + 1] += start[(size_t)r];
    std::vector<int32_t> col((size_t)n);
    std::vector<double> val((size_t)n);
    {
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int64_t k = 0; k < n; ++k) {
            const int64_t p = cur[(size_t)(c->uid[k] - uid_base)]++;
            col[(size_t)p] = (int32_t)c->iid[k];
            val[(size_t)p] = use_values ? c->val[(size_t)k] : 1.0;
        }
    }
    // per row: sort by column (stable), sum duplicates in file order
    int64_t w = 0;
    std::vector<std::pair<int32_t, double>> tmp;
    indptr_host[0] = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t a = start[(size_t)r], b = start[(size_t)r + 1];
        bool sorted = true;
        for (int64_t k = a + 1; k < b && sorted; ++k) sorted = col[(size_t)k - 1] < col[(size_t)k];
        if (sorted) {
            for (int64_t k = a; k < b; ++k) { indices_host[w] = col[(size_t)k]; values_host[w] = val[(size_t)k]; ++w; }
        } else {
            tmp.clear();
            for (int64_t k = a; k < b; ++k) tmp.emplace_back(col[(size_t)k], val[(size_t)k]);
            std::stable_sort(tmp.begin(), tmp.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
            for (size_t k = 0; k < tmp.size(); ++k) {
                if (k > 0 && tmp[k].first == tmp[k - 1].first) {
                    values_host[w - 1] += tmp[k].second;
                } else {
                    indices_host[w] = tmp[k].first;
                    values_host[w] = tmp[k].second;
                    ++w;
                }
            }
        }
        indptr_host[r + 1] = w;
    }
    *nnz_out = w;
    return 0;
}

int b200vae_csv_close(b200vae_csv* h) {
    delete reinterpret_cast<Csv*>(h);
    return 0;
}

}  // extern "C"
