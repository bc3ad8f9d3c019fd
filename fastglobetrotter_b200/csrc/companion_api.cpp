// companion_api.cpp -- exported C ABI (include/fgt_companion.h): the four R-callable entry points
// with the reference's file handling and fatal-error behaviour, plus the packed entry points.
//
// File layer (host, multithreaded): list files, recombination maps, *.samples.out[.gz] paintings
// (reference C:201-275, C:341-445).  Paintings are inflated with zlib and tokenised by a pool of
// threads straight into packed per-haplotype label arrays; parsed files are cached per process
// (the R driver calls data_read ~100 times on the same files).
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <thread>
#include <unordered_map>
#include <vector>

#include "engine.h"

using namespace fgt;

namespace {

std::mutex g_api_mu;   // the engine is a per-process singleton; R calls serially, other hosts are serialised here

[[noreturn]] void die(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void die(const char *fmt, ...) {
    // reference convention: printf(message); exit(1) -- this ends the R session (SURVEY 8b "Errors")
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    fflush(stdout);
    exit(1);
}

// ---- list files (reference C:201-229): one path per line, first token used; the number of entries
// is the number of newline-terminated lines (the reference's feof() loop drops an unterminated last line)
std::vector<std::string> read_list_file(const char *path) {
    FILE *f = fopen(path, "r");
    if (!f) die("error opening %s\n", path);
    std::string all;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) all.append(buf, n);
    fclose(f);
    std::vector<std::string> out;
    size_t p = 0;
    while (true) {
        size_t e = all.find('\n', p);
        if (e == std::string::npos) break;
        size_t a = p;
        while (a < e && isspace((unsigned char)all[a])) a++;
        size_t b = a;
        while (b < e && !isspace((unsigned char)all[b])) b++;
        out.emplace_back(all.substr(a, b - a));
        p = e + 1;
    }
    return out;
}

struct RecombMap { std::vector<double> pos, rate; };

// reference C:353-397: header, then "pos rate" rows; nsites = newline-terminated data rows
RecombMap read_recomb(const std::string &path) {
    FILE *f = fopen(path.c_str(), "r");
    if (!f) die("error opening %s\n", path.c_str());
    std::string all;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) all.append(buf, n);
    fclose(f);
    size_t p = all.find('\n');
    if (p == std::string::npos) die("Something wrong with %s. Exiting....\n", path.c_str());
    p++;
    RecombMap m;
    while (true) {
        size_t e = all.find('\n', p);
        if (e == std::string::npos) break;
        const char *s = all.c_str() + p;
        char *end1 = nullptr, *end2 = nullptr;
        double a = strtod(s, &end1);          // sscanf("%lf") and strtod agree on decimal text
        double b = strtod(end1, &end2);
        if (end1 == s || end2 == end1) { a = (end1 == s) ? 0.0 : a; b = 0.0; }
        m.pos.push_back(a);
        m.rate.push_back(b);
        p = e + 1;
    }
    if (m.pos.empty()) die("Something wrong with %s -- no SNPs! Exiting....\n", path.c_str());
    for (size_t i = 0; i < m.pos.size(); i++) {
        if (i > 0 && m.pos[i] < m.pos[i - 1])
            die("positions in %s are not listed in increasing order at basepairs %lf and %lf. Exiting....\n", path.c_str(),
                m.pos[i - 1], m.pos[i]);
        if (m.rate[i] < 0) die("recom rate must be > 0 (basepair %lf)!! Exiting....\n", m.pos[i]);
    }
    return m;
}

// ---- samples files --------------------------------------------------------------------------
struct SamplesText {           // whole inflated file + line index
    std::string text;
    std::vector<size_t> line_start;   // start of every newline-terminated line, plus a trailing sentinel
};

bool inflate_file(const std::string &path, SamplesText &out) {
    gzFile g = gzopen(path.c_str(), "r");
    if (!g) return false;
    gzbuffer(g, 1 << 20);
    std::string &t = out.text;
    t.clear();
    struct stat stt;
    if (stat(path.c_str(), &stt) == 0) t.reserve((size_t)stt.st_size * 8 + 4096);
    std::vector<char> buf(8 << 20);
    int n;
    while ((n = gzread(g, buf.data(), (unsigned)buf.size())) > 0) t.append(buf.data(), (size_t)n);
    gzclose(g);
    out.line_start.clear();
    out.line_start.push_back(0);
    const char *base = t.data();
    const char *p = base, *end = base + t.size();
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', end - p);
        if (!nl) break;
        p = nl + 1;
        out.line_start.push_back((size_t)(p - base));
    }
    return true;
}

inline const char *skip_ws(const char *p, const char *e) { while (p < e && isspace((unsigned char)*p)) p++; return p; }
inline const char *skip_tok(const char *p, const char *e) { while (p < e && !isspace((unsigned char)*p)) p++; return p; }

std::string token_at(const char *&p, const char *e) {
    p = skip_ws(p, e);
    const char *a = p;
    p = skip_tok(p, e);
    return std::string(a, p - a);
}

struct SamplesMeta {           // what the reference learns from scanning the FIRST samples file (C:230-297)
    int nsamples = 0;
    long count = 0;            // lines after the header, as the reference counts them (+1 for the failing read)
    std::vector<int> rec_first_index;   // file individual index of each recipient, in file order
};

SamplesMeta scan_first_file(const SamplesText &st, const std::string &path, int ploidy, int n_rec, char **rec_labels) {
    SamplesMeta m;
    if (st.line_start.size() < 2) die("Something wrong with %s. Exiting....\n", path.c_str());
    const char *base = st.text.data();
    {   // header: token after "nsamples" "=" (C:234-239)
        const char *p = base, *e = base + std::min<size_t>(st.line_start[1], 2047);
        std::string tok = token_at(p, e);
        while (tok != "nsamples") {
            if (p >= e) die("Something wrong with %s. Exiting....\n", path.c_str());
            tok = token_at(p, e);
        }
        token_at(p, e);
        std::string v = token_at(p, e);
        m.nsamples = atoi(v.c_str());
    }
    if (m.nsamples <= 0) die("No painting samples in %s! Exiting....\n", path.c_str());
    std::unordered_map<std::string, int> names;
    for (int i = 0; i < n_rec; i++) names.emplace(rec_labels[i], i);
    long rec_count = 0, hap_count = 0;
    m.rec_first_index.assign(n_rec, 0);
    const size_t nlines = st.line_start.size() - 1;    // newline-terminated lines incl. header
    for (size_t li = 1; li < nlines; li++) {
        const char *p = base + st.line_start[li], *e = base + st.line_start[li + 1];
        p = skip_ws(p, e);
        if (e - p >= 4 && p[0] == 'H' && p[1] == 'A' && p[2] == 'P' && isspace((unsigned char)p[3])) {
            p += 3;
            token_at(p, e);
            std::string nm = token_at(p, e);
            if (names.count(nm)) {
                if (rec_count >= (long)ploidy * n_rec)
                    die("There are more than %d recipient haplotypes in %s! Exiting....\n", n_rec * ploidy, path.c_str());
                m.rec_first_index[rec_count / ploidy] = (int)(hap_count / ploidy);
                rec_count++;
            }
            hap_count++;
        }
    }
    bool partial = st.line_start.back() < st.text.size();
    m.count = (long)(nlines - 1) + 1;                    // + the read that hits EOF (C:248-274)
    (void)partial;
    long nhaps = m.count / (m.nsamples + 1);
    if (nhaps != ((double)m.count) / (m.nsamples + 1) && nhaps != ((double)m.count - 1) / (m.nsamples + 1))
        die("something wrong with file %s. Exiting....\n", path.c_str());
    long ninds = nhaps / ploidy;
    if (ninds <= 0) die("No individuals in %s! Exiting....\n", path.c_str());
    if (rec_count != (long)n_rec * ploidy)
        die("Only %ld recipient haplotypes in %s, while expecting %d! Exiting....\n", rec_count, path.c_str(), n_rec * ploidy);
    return m;
}

template <typename T>
bool parse_ints(const char *p, const char *e, T *dst, int n, int &maxv) {
    // first token (sample index) is skipped like reading(&step,"%s",waste), then n integers (C:443-445)
    p = skip_ws(p, e);
    p = skip_tok(p, e);
    int mx = maxv;
    for (int j = 0; j < n; j++) {
        while (p < e && (unsigned char)*p <= ' ') p++;
        if (p >= e) return false;
        bool neg = false;
        if (*p == '-') { neg = true; p++; } else if (*p == '+') p++;
        unsigned v = 0;
        const char *a = p;
        while (p < e && (unsigned)(*p - '0') <= 9u) { v = v * 10u + (unsigned)(*p - '0'); p++; }
        if (p == a) return false;
        while (p < e && (unsigned char)*p > ' ') p++;          // rest of a malformed token is ignored like sscanf does
        int iv = neg ? -(int)v : (int)v;
        if (iv > mx) mx = iv;
        dst[j] = (T)iv;
    }
    maxv = mx;
    return true;
}

struct PackedFile {
    std::string path;
    long long fsize = 0, mtime_ns = 0;
    int nsites = 0, nsamples = 0, ploidy = 0, label_bytes = 2, max_label = 0;
    std::vector<int> rec_index;
    std::vector<uint16_t> lab16;
    std::vector<int32_t> lab32;
    void *dev = nullptr;       // device copy (owned), filled lazily
    size_t dev_bytes = 0;
};

int n_threads() {
    int t = Engine::get().opt.threads;
    if (t <= 0) t = (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(t, 64));
}

// decode the paintings of the recipients (file individual indices rec_index) into a packed matrix
void pack_file(const SamplesText &st, const std::string &path, int ploidy, int nsamples, int nsites,
               const std::vector<int> &rec_index, PackedFile &pf) {
    const size_t nlines = st.line_start.size() - 1;
    const int per_ind = ploidy * (nsamples + 1);
    const int n_rec = (int)rec_index.size();
    const size_t rows = (size_t)n_rec * ploidy * nsamples;
    for (int r = 0; r < n_rec; r++) {
        size_t last = 1 + (size_t)rec_index[r] * per_ind + per_ind;     // line index after this individual
        if (last > nlines) die("Something wrong with %s -- perhaps does not contain %d individuals? Exiting....\n", path.c_str(), rec_index[r] + 1);
    }
    pf.lab16.assign(rows * nsites, 0);
    std::atomic<int> overflow{0}, failed{-1}, gmax{0};
    auto work = [&](int tid, int nt, bool wide) {
        for (size_t row = tid; row < rows; row += nt) {
            int r = (int)(row / ((size_t)ploidy * nsamples));
            int rem = (int)(row % ((size_t)ploidy * nsamples));
            int p = rem / nsamples, s = rem % nsamples;
            size_t li = 1 + (size_t)rec_index[r] * per_ind + (size_t)p * (nsamples + 1) + 1 + s;
            const char *a = st.text.data() + st.line_start[li], *e = st.text.data() + st.line_start[li + 1];
            int mx = 0;
            bool ok;
            if (!wide) {
                ok = parse_ints<uint16_t>(a, e, pf.lab16.data() + row * nsites, nsites, mx);
                if (mx > 65535) overflow = 1;
            } else ok = parse_ints<int32_t>(a, e, pf.lab32.data() + row * nsites, nsites, mx);
            if (!ok) failed = (int)row;
            int seen = gmax.load();
            while (mx > seen && !gmax.compare_exchange_weak(seen, mx)) {}
        }
    };
    auto run = [&](bool wide) {
        int nt = n_threads();
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, t, nt, wide);
        for (auto &t : th) t.join();
    };
    run(false);
    pf.label_bytes = 2;
    if (overflow) {                 // more than 65535 donor haplotypes: redo with 32-bit labels
        pf.lab16.clear(); pf.lab16.shrink_to_fit();
        pf.lab32.assign(rows * nsites, 0);
        run(true);
        pf.label_bytes = 4;
    }
    if (failed >= 0) {
        int row = failed;
        int r = row / (ploidy * nsamples), s = row % nsamples;
        die("Something wrong with %s at sample %d of individual %d. Exiting....\n", path.c_str(), s + 1, rec_index[r] + 1);
    }
    pf.max_label = gmax.load();
    pf.nsites = nsites; pf.nsamples = nsamples; pf.ploidy = ploidy; pf.rec_index = rec_index; pf.path = path;
}

std::mutex g_cache_mu;
std::map<std::string, std::shared_ptr<PackedFile>> g_cache;

bool file_identity(const std::string &path, long long &size, long long &mtime_ns) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return false;
    size = (long long)st.st_size;
    mtime_ns = (long long)st.st_mtim.tv_sec * 1000000000LL + st.st_mtim.tv_nsec;
    return true;
}

std::string cache_key(const std::string &path, long long size, long long mt, int ploidy, int nsites,
                      const std::vector<int> &rec_index) {
    unsigned long long h = 1469598103934665603ULL;
    for (int v : rec_index) { h ^= (unsigned)v; h *= 1099511628211ULL; }
    char buf[256];
    snprintf(buf, sizeof buf, "|%lld|%lld|%d|%d|%zu|%llx", size, mt, ploidy, nsites, rec_index.size(), h);
    return path + buf;
}

// ---- optional on-disk cache of decoded paintings --------------------------------------------------
// With FGT_PACK_CACHE_DIR set, every decoded samples file (and what the scan of the first file learnt) is also
// written there as a binary file keyed by source path, size, mtime and recipient set; a later process reads it back
// instead of inflating and tokenising the text.  Off by default: the reference library writes no files.
const char *pack_dir() {
    const char *d = getenv("FGT_PACK_CACHE_DIR");
    return (d && *d) ? d : nullptr;
}

std::string pack_path(const char *dir, const std::string &key, const char *ext) {
    unsigned long long h = 1469598103934665603ULL;
    for (unsigned char c : key) { h ^= c; h *= 1099511628211ULL; }
    char buf[64];
    snprintf(buf, sizeof buf, "/%016llx.%s", h, ext);
    return std::string(dir) + buf;
}

struct PackHdr {
    char magic[8];
    uint32_t key_len, label_bytes;
    int32_t nsites, nsamples, ploidy, max_label;
    uint64_t n_rec, n_labels;
};

bool write_all(FILE *f, const void *p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
bool read_all(FILE *f, void *p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }

void disk_store(const char *dir, const std::string &key, const PackedFile &pf) {
    const std::string path = pack_path(dir, key, "fgtpack"), tmp = path + ".tmp" + std::to_string((long long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;                                   // the cache is best effort: failures never reach the caller
    PackHdr h{};
    memcpy(h.magic, "FGTPACK1", 8);
    h.key_len = (uint32_t)key.size(); h.label_bytes = (uint32_t)pf.label_bytes;
    h.nsites = pf.nsites; h.nsamples = pf.nsamples; h.ploidy = pf.ploidy; h.max_label = pf.max_label;
    h.n_rec = pf.rec_index.size();
    h.n_labels = pf.label_bytes == 2 ? pf.lab16.size() : pf.lab32.size();
    const void *lab = pf.label_bytes == 2 ? (const void *)pf.lab16.data() : (const void *)pf.lab32.data();
    bool ok = write_all(f, &h, sizeof h) && write_all(f, key.data(), key.size()) &&
              write_all(f, pf.rec_index.data(), pf.rec_index.size() * sizeof(int)) &&
              write_all(f, lab, (size_t)h.n_labels * pf.label_bytes);
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

bool disk_load(const char *dir, const std::string &key, const std::string &src_path, PackedFile &pf) {
    FILE *f = fopen(pack_path(dir, key, "fgtpack").c_str(), "rb");
    if (!f) return false;
    PackHdr h{};
    bool ok = read_all(f, &h, sizeof h) && memcmp(h.magic, "FGTPACK1", 8) == 0 && h.key_len == key.size() &&
              (h.label_bytes == 2 || h.label_bytes == 4) && h.n_rec < (1u << 28) && h.nsites > 0;
    if (ok) {
        std::string k(h.key_len, '\0');
        ok = read_all(f, &k[0], k.size()) && k == key;     // the key holds path, size, mtime, ploidy, sites, recipients
    }
    if (ok) {
        pf.rec_index.resize(h.n_rec);
        ok = read_all(f, pf.rec_index.data(), h.n_rec * sizeof(int)) &&
             h.n_labels == (uint64_t)h.n_rec * h.ploidy * h.nsamples * h.nsites;
    }
    if (ok) {
        if (h.label_bytes == 2) { pf.lab16.resize(h.n_labels); ok = read_all(f, pf.lab16.data(), h.n_labels * 2); }
        else { pf.lab32.resize(h.n_labels); ok = read_all(f, pf.lab32.data(), h.n_labels * 4); }
        char extra;
        ok = ok && fread(&extra, 1, 1, f) == 0;            // nothing may follow
    }
    fclose(f);
    if (!ok) { pf = PackedFile(); return false; }
    pf.label_bytes = (int)h.label_bytes; pf.nsites = h.nsites; pf.nsamples = h.nsamples; pf.ploidy = h.ploidy;
    pf.max_label = h.max_label; pf.path = src_path;
    return true;
}

void disk_store_meta(const char *dir, const std::string &key, const SamplesMeta &m) {
    const std::string path = pack_path(dir, key, "fgtmeta"), tmp = path + ".tmp" + std::to_string((long long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const uint64_t hdr[4] = {0x315441544d544746ULL /* "FGTMETA1" */, key.size(), (uint64_t)m.rec_first_index.size(), (uint64_t)m.nsamples};
    const long long cnt = m.count;
    bool ok = write_all(f, hdr, sizeof hdr) && write_all(f, &cnt, sizeof cnt) && write_all(f, key.data(), key.size()) &&
              write_all(f, m.rec_first_index.data(), m.rec_first_index.size() * sizeof(int));
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

bool disk_load_meta(const char *dir, const std::string &key, SamplesMeta &m) {
    FILE *f = fopen(pack_path(dir, key, "fgtmeta").c_str(), "rb");
    if (!f) return false;
    uint64_t hdr[4];
    long long cnt = 0;
    bool ok = read_all(f, hdr, sizeof hdr) && hdr[0] == 0x315441544d544746ULL && hdr[1] == key.size() && hdr[2] < (1u << 28) &&
              read_all(f, &cnt, sizeof cnt);
    if (ok) {
        std::string k(hdr[1], '\0');
        ok = read_all(f, &k[0], k.size()) && k == key;
    }
    if (ok) {
        m.rec_first_index.resize(hdr[2]);
        ok = read_all(f, m.rec_first_index.data(), hdr[2] * sizeof(int));
    }
    fclose(f);
    if (!ok) { m = SamplesMeta(); return false; }
    m.nsamples = (int)hdr[3]; m.count = (long)cnt;
    return true;
}

struct Loaded {                 // everything a .C call needs, decoded
    int num_chrom = 0, nsamples = 0;
    std::vector<RecombMap> maps;
    std::vector<std::shared_ptr<PackedFile>> files;
    std::vector<int> rec_index;
};

Loaded load_inputs(int ploidy, char **samp_list, char **recom_list, int n_rec, char **rec_labels, int n_rec_needed) {
    Loaded L;
    std::vector<std::string> sfiles = read_list_file(*samp_list), rfiles = read_list_file(*recom_list);
    if (sfiles.size() != rfiles.size())
        die("The number of files in %s and %s do not match. Exiting....\n", *samp_list, *recom_list);
    if (sfiles.empty()) die("Something wrong with %s. Exiting....\n", *samp_list);
    L.num_chrom = (int)sfiles.size();
    const bool use_cache = Engine::get().opt.cache != 0;

    using TextFuture = std::future<std::shared_ptr<SamplesText>>;
    auto start_inflate = [&](int h) {
        return std::async(std::launch::async, [&, h]() {
            auto t = std::make_shared<SamplesText>();
            if (!inflate_file(sfiles[h], *t)) return std::shared_ptr<SamplesText>();
            return t;
        });
    };
    // files being inflated / held as text at once: up to four, fewer when the texts are large (painting text inflates
    // about 30x; at most ~16 GB of it is kept in memory)
    long long first_size = 0, first_mtime = 0;
    file_identity(sfiles[0], first_size, first_mtime);
    auto inflate_window = [&]() {
        const long long est = std::max<long long>(1, first_size * 32);
        const long long fit = std::max<long long>(1, (16LL << 30) / est);
        return (int)std::max<long long>(1, std::min<long long>({(long long)n_threads(), 4LL, fit}));
    };
    std::map<int, TextFuture> early;            // inflates in flight, by chromosome

    // first file: learn nsamples and which file individuals are the recipients (C:230-297)
    std::shared_ptr<SamplesText> first_text;
    SamplesMeta meta;
    {
        long long sz = 0, mt = 0;
        if (!file_identity(sfiles[0], sz, mt)) die("error opening %s\n", sfiles[0].c_str());
        static std::mutex mu;
        static std::map<std::string, SamplesMeta> meta_cache;
        std::string mk = sfiles[0] + "|" + std::to_string(sz) + "|" + std::to_string(mt) + "|" + std::to_string(ploidy);
        for (int i = 0; i < n_rec; i++) { mk += "|"; mk += rec_labels[i]; }
        bool hit = false;
        if (use_cache) {
            std::lock_guard<std::mutex> lk(mu);
            auto it = meta_cache.find(mk);
            if (it != meta_cache.end()) { meta = it->second; hit = true; }
        }
        if (!hit && pack_dir() && disk_load_meta(pack_dir(), mk, meta) && (int)meta.rec_first_index.size() == n_rec) hit = true;
        if (!hit) {
            // cold start: nothing of this data set is decoded yet, so the next files start inflating right away (zlib is
            // sequential per file; several files at once is the only parallelism the format offers) while the first one
            // is inflated here and scanned
            for (int h = 1; h < L.num_chrom && (int)early.size() < inflate_window() - 1; h++) early[h] = start_inflate(h);
            first_text = std::make_shared<SamplesText>();
            if (!inflate_file(sfiles[0], *first_text)) die("error opening %s\n", sfiles[0].c_str());
            meta = scan_first_file(*first_text, sfiles[0], ploidy, n_rec, rec_labels);
            if (pack_dir()) disk_store_meta(pack_dir(), mk, meta);
        }
        if (use_cache) { std::lock_guard<std::mutex> lk(mu); meta_cache[mk] = meta; }
    }
    L.nsamples = meta.nsamples;
    L.rec_index = meta.rec_first_index;
    std::vector<int> need(meta.rec_first_index.begin(), meta.rec_first_index.begin() + std::min(n_rec, n_rec_needed));

    L.maps.resize(L.num_chrom);
    L.files.resize(L.num_chrom);
    for (int h = 0; h < L.num_chrom; h++) L.maps[h] = read_recomb(rfiles[h]);
    std::vector<std::string> keys(L.num_chrom);
    std::vector<char> cached(L.num_chrom, 0);
    for (int h = 0; h < L.num_chrom; h++) {
        long long sz = 0, mt = 0;
        if (!file_identity(sfiles[h], sz, mt)) die("error opening %s\n", sfiles[h].c_str());
        keys[h] = cache_key(sfiles[h], sz, mt, ploidy, (int)L.maps[h].pos.size(), need);
        if (use_cache) {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            auto it = g_cache.find(keys[h]);
            if (it != g_cache.end()) { L.files[h] = it->second; cached[h] = 1; }
        }
        if (!cached[h] && pack_dir()) {
            auto pf = std::make_shared<PackedFile>();
            if (disk_load(pack_dir(), keys[h], sfiles[h], *pf) && pf->nsamples == meta.nsamples && pf->rec_index == need) {
                pf->fsize = sz; pf->mtime_ns = mt;
                L.files[h] = pf; cached[h] = 1;
                if (use_cache) { std::lock_guard<std::mutex> lk(g_cache_mu); g_cache[keys[h]] = pf; }
            }
        }
    }
    // tokenise the files in order; up to inflate_window() later files are being inflated meanwhile
    std::vector<int> todo;
    for (int h = 0; h < L.num_chrom; h++) if (!cached[h]) todo.push_back(h);
    for (auto it = early.begin(); it != early.end();) {         // a speculative inflate of a file that turned out cached
        if (cached[it->first]) { it->second.wait(); it = early.erase(it); } else ++it;
    }
    size_t issued = 0;                                           // todo[0..issued) have an inflate started or a text in hand
    auto top_up = [&](size_t done) {
        while (issued < todo.size() && issued < done + (size_t)inflate_window()) {
            const int h = todo[issued++];
            if (h == 0 && first_text) continue;
            if (!early.count(h)) early[h] = start_inflate(h);
        }
    };
    for (size_t q = 0; q < todo.size(); q++) {
        top_up(q);
        const int cur = todo[q];
        std::shared_ptr<SamplesText> cur_text;
        if (cur == 0 && first_text) cur_text = first_text;
        else { cur_text = early[cur].get(); early.erase(cur); }
        if (!cur_text) die("error opening %s\n", sfiles[cur].c_str());
        top_up(q + 1);                                           // this text is about to be released: one more file may start
        auto pf = std::make_shared<PackedFile>();
        pack_file(*cur_text, sfiles[cur], ploidy, meta.nsamples, (int)L.maps[cur].pos.size(), need, *pf);
        L.files[cur] = pf;
        if (pack_dir()) disk_store(pack_dir(), keys[cur], *pf);
        if (use_cache) { std::lock_guard<std::mutex> lk(g_cache_mu); g_cache[keys[cur]] = pf; }
        if (cur == 0) first_text.reset();
    }
    for (auto &kv : early) kv.second.wait();
    first_text.reset();
    return L;
}

// device copy of a packed file (cached with the file)
const void *device_labels(PackedFile &pf) {
    if (pf.dev) return pf.dev;
    size_t bytes = pf.label_bytes == 2 ? pf.lab16.size() * 2 : pf.lab32.size() * 4;
    void *d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    const void *src = pf.label_bytes == 2 ? (const void *)pf.lab16.data() : (const void *)pf.lab32.data();
    if (cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); cudaGetLastError(); return nullptr; }
    pf.dev = d; pf.dev_bytes = bytes;
    return d;
}

void fatal_rc(int rc) {
    switch (rc) {
        case FGT_ERR_NO_DEVICE: die("fastGLOBETROTTERCompanion (B200 build): no CUDA device available and there is no CPU path. Exiting....\n");
        case FGT_ERR_MISSING_DONOR: die("Individual has been painted by missing donor -9, Exiting....\n");
        case FGT_ERR_LABEL_RANGE: die("The painting samples and ID file are not matched, please check number of individuals in the ID file. Exiting....\n");
        case FGT_ERR_MODE3_TARGET: die("Mode 3 met a chunk copied from the target population (undefined behaviour in the reference companion). Exiting....\n");
        case FGT_ERR_RAND: die("libc rand() state is not the default additive-feedback generator. Exiting....\n");
        default: die("fastGLOBETROTTERCompanion (B200 build): internal error %d. Exiting....\n", rc);
    }
}

void curves_from_files(int mode, int *ploidy, int *ind_id_vec, char **samp, char **recom, double *binwidth, int *num_bins,
                       int *grid_start, int *num_donors, int *donor_label_vec, int *num_inds_rec, char **rec_labels,
                       double *weight_min, double *cutoff, int *num_surrogates, double *temp_weights, double *means) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    const int N = *num_inds_rec;
    Loaded L = load_inputs(*ploidy, samp, recom, N, rec_labels, N);
    std::vector<fgt_chrom_t> ch(L.num_chrom);
    bool on_dev = Engine::get().opt.cache != 0;
    for (int h = 0; h < L.num_chrom; h++) {
        PackedFile &pf = *L.files[h];
        ch[h].nsites = pf.nsites;
        ch[h].pos = L.maps[h].pos.data();
        ch[h].rate = L.maps[h].rate.data();
        const void *d = on_dev ? device_labels(pf) : nullptr;
        ch[h].labels = d ? d : (pf.label_bytes == 2 ? (const void *)pf.lab16.data() : (const void *)pf.lab32.data());
        ch[h].labels_on_device = d ? 1 : 0;
        if (pf.label_bytes != L.files[0]->label_bytes) die("fastGLOBETROTTERCompanion (B200 build): mixed label widths across chromosomes. Exiting....\n");
    }
    for (int i = 0; i < N * L.num_chrom; i++)
        if (ind_id_vec[i] < 0 || ind_id_vec[i] >= N) die("Something wrong with the individual index vector. Exiting....\n");
    int max_hap = 0;
    // donor_label_vec has one entry per haplotype of the id file; its length is not passed through .C, so take
    // the largest haplotype id that occurs in the paintings as the bound we need
    for (int h = 0; h < L.num_chrom; h++) {
        PackedFile &pf = *L.files[h];
        max_hap = std::max(max_hap, pf.max_label);
    }
    fgt_problem_t pb{};
    pb.mode = mode; pb.ploidy = *ploidy; pb.nsamples = L.nsamples; pb.num_chrom = L.num_chrom;
    pb.n_packed_inds = N; pb.label_bytes = L.files[0]->label_bytes; pb.chrom = ch.data();
    pb.num_inds = N; pb.ind_rows = ind_id_vec; pb.num_inds_total = N; pb.pair_offset = nullptr;
    pb.binwidth = *binwidth; pb.num_bins = *num_bins; pb.grid_start = *grid_start; pb.num_donors = *num_donors;
    pb.donor_label_vec = donor_label_vec; pb.n_donor_labels = max_hap; pb.weight_min = *weight_min;
    pb.gridweight_max_cutoff = *cutoff; pb.num_surrogates = *num_surrogates; pb.temp_weights = temp_weights;
    int rc = Engine::get().data_read_packed(pb, means, nullptr, nullptr);
    if (rc) fatal_rc(rc);
    if (std::isnan(means[0])) die("NA values. Exiting....\n");     // reference C:660-663 / C:1620
}

}  // namespace

// =================================================================================================
extern "C" {

void mutiply_temp_weight(double *tempweights, double *weights_mat, int *num_surrogates, int *num_donors) {
    int rc = Engine::get().outer_weights(tempweights, weights_mat, *num_surrogates, *num_donors);
    if (rc) fatal_rc(rc);
}

void data_read(int *ploidy, int *ind_id_vec, char **filenameSAMPread, char **filenameRECOMread, double *binwidth,
               int *num_bins, int *grid_start, int *num_donors, int *donor_label_vec, int *num_inds_rec,
               char **recipient_label_vec, double *weight_min, double *gridweightMAX_cutoff, int *num_surrogates,
               double *temp_weights, double *means) {
    curves_from_files(1, ploidy, ind_id_vec, filenameSAMPread, filenameRECOMread, binwidth, num_bins, grid_start,
                      num_donors, donor_label_vec, num_inds_rec, recipient_label_vec, weight_min, gridweightMAX_cutoff,
                      num_surrogates, temp_weights, means);
}

void data_read_mode3(int *ploidy, int *ind_id_vec, char **filenameSAMPread, char **filenameRECOMread, double *binwidth,
                     int *num_bins, int *grid_start, int *num_donors, int *donor_label_vec, int *num_inds_rec,
                     char **recipient_label_vec, double *weight_min, double *gridweightMAX_cutoff, int *num_surrogates,
                     double *temp_weights, double *means) {
    curves_from_files(3, ploidy, ind_id_vec, filenameSAMPread, filenameRECOMread, binwidth, num_bins, grid_start,
                      num_donors, donor_label_vec, num_inds_rec, recipient_label_vec, weight_min, gridweightMAX_cutoff,
                      num_surrogates, temp_weights, means);
}

void data_read_null(int *ploidy, char **filenameSAMPread, char **filenameRECOMread, double *binwidth, int *num_bins,
                    int *grid_start, int *num_donors, int *donor_label_vec, int *num_inds_rec, char **recipient_label_vec,
                    double *weight_min, double *gridweightMAX_cutoff, double *chrom_ind_res) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    const int N = *num_inds_rec;
    const int Nc = std::min(N, 100);
    Loaded L = load_inputs(*ploidy, filenameSAMPread, filenameRECOMread, N, recipient_label_vec, Nc);
    std::vector<fgt_chrom_t> ch(L.num_chrom);
    int max_hap = 0;
    for (int h = 0; h < L.num_chrom; h++) {
        PackedFile &pf = *L.files[h];
        ch[h].nsites = pf.nsites;
        ch[h].pos = L.maps[h].pos.data();
        ch[h].rate = L.maps[h].rate.data();
        const void *d = Engine::get().opt.cache ? device_labels(pf) : nullptr;
        ch[h].labels = d ? d : (pf.label_bytes == 2 ? (const void *)pf.lab16.data() : (const void *)pf.lab32.data());
        ch[h].labels_on_device = d ? 1 : 0;
        max_hap = std::max(max_hap, pf.max_label);
    }
    std::vector<int32_t> rec_rows(Nc);
    for (int i = 0; i < Nc; i++) rec_rows[i] = i;
    fgt_problem_t pb{};
    pb.mode = 1; pb.ploidy = *ploidy; pb.nsamples = L.nsamples; pb.num_chrom = L.num_chrom; pb.n_packed_inds = Nc;
    pb.label_bytes = L.files[0]->label_bytes; pb.chrom = ch.data(); pb.num_inds = N; pb.num_inds_total = N;
    pb.binwidth = *binwidth; pb.num_bins = *num_bins; pb.grid_start = *grid_start; pb.num_donors = *num_donors;
    pb.donor_label_vec = donor_label_vec; pb.n_donor_labels = max_hap; pb.weight_min = *weight_min;
    pb.gridweight_max_cutoff = *gridweightMAX_cutoff; pb.num_surrogates = 1;
    int rc = Engine::get().data_read_null_packed(pb, rec_rows.data(), chrom_ind_res, nullptr);
    if (rc) fatal_rc(rc);
    if (std::isnan(chrom_ind_res[0])) die("NA values. Exiting....\n");    // reference C:1055
}

int fgt_data_read_packed(const fgt_problem_t *prob, double *means, fgt_stats_t *stats) {
    if (!prob || !means) return FGT_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_api_mu);
    return Engine::get().data_read_packed(*prob, means, stats, nullptr);
}

int fgt_count_pairs_packed(const fgt_problem_t *prob, int64_t *out) {
    if (!prob || !out) return FGT_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_api_mu);
    return Engine::get().data_read_packed(*prob, nullptr, nullptr, out);
}

int fgt_data_read_null_packed(const fgt_problem_t *prob, const int32_t *rec_rows, double *chrom_ind_res, fgt_stats_t *stats) {
    if (!prob || !rec_rows || !chrom_ind_res) return FGT_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_api_mu);
    return Engine::get().data_read_null_packed(*prob, rec_rows, chrom_ind_res, stats);
}

int64_t fgt_get_raw_counts(double *out, int64_t capacity) { return Engine::get().get_raw(out, capacity); }

int fgt_set_option(const char *key, double value) { return key ? Engine::get().set_option(key, value) : FGT_ERR_ARG; }
double fgt_get_option(const char *key) { return key ? Engine::get().get_option(key) : NAN; }
void fgt_srand(unsigned seed) { Engine::get().srand_private(seed); }
void fgt_last_stats(fgt_stats_t *out) { if (out) *out = Engine::get().last_stats; }
int fgt_rand_fill(int32_t *out, int64_t n) { return out ? Engine::get().rand_fill_host(out, n) : FGT_ERR_ARG; }

int fgt_chunk_one(const void *labels, int label_bytes, int nsites, const double *pos, const double *rate, double weight_min,
                  double cutoff, const int32_t *donor_label_vec, int n_donor_labels, double *cpos, double *csize,
                  double *cweight, double *cend, int32_t *cpop) {
    return Engine::get().chunk_one(labels, label_bytes, nsites, pos, rate, weight_min, cutoff, donor_label_vec,
                                   n_donor_labels, cpos, csize, cweight, cend, cpop);
}

int fgt_parse_probe(const char *samples_list, const char *recom_list, int ploidy, int n_rec, char **rec_labels,
                    int32_t *num_chrom, int32_t *nsamples, int32_t *nsites /*[num_chrom]*/, int32_t *rec_index /*[n_rec]*/,
                    int32_t *label_bytes, uint64_t *label_sum /*[num_chrom]*/, void **labels_out /*[num_chrom] or NULL*/,
                    double *pos_out /*concatenated or NULL*/, double *rate_out) {
    // host-only: run the text layer (list files, recombination maps, multithreaded gz decode) and report
    // what it decoded.  Lets the CPU test suite check the parser without a GPU.
    if (!samples_list || !recom_list || !num_chrom) return FGT_ERR_ARG;
    char *sl = const_cast<char *>(samples_list), *rl = const_cast<char *>(recom_list);
    Loaded L = load_inputs(ploidy, &sl, &rl, n_rec, rec_labels, n_rec);
    *num_chrom = L.num_chrom;
    if (nsamples) *nsamples = L.nsamples;
    if (rec_index) for (int i = 0; i < n_rec; i++) rec_index[i] = L.rec_index[i];
    size_t off = 0;
    for (int h = 0; h < L.num_chrom; h++) {
        PackedFile &pf = *L.files[h];
        if (nsites) nsites[h] = pf.nsites;
        if (label_bytes) *label_bytes = pf.label_bytes;
        uint64_t sum = 0;
        if (pf.label_bytes == 2) for (uint16_t v : pf.lab16) sum += v;
        else for (int32_t v : pf.lab32) sum += (uint64_t)(uint32_t)v;
        if (label_sum) label_sum[h] = sum;
        if (labels_out && labels_out[h]) {
            if (pf.label_bytes == 2) std::memcpy(labels_out[h], pf.lab16.data(), pf.lab16.size() * 2);
            else std::memcpy(labels_out[h], pf.lab32.data(), pf.lab32.size() * 4);
        }
        if (pos_out) std::memcpy(pos_out + off, L.maps[h].pos.data(), pf.nsites * sizeof(double));
        if (rate_out) std::memcpy(rate_out + off, L.maps[h].rate.data(), pf.nsites * sizeof(double));
        off += pf.nsites;
    }
    return FGT_OK;
}

int fgt_rand_host_jump(unsigned seed, uint64_t n_draws, uint32_t *hist_out) {
    // host-only: history (oldest first) of the glibc stream `n_draws` after srand(seed), by polynomial jump-ahead
    if (!hist_out) return FGT_ERR_ARG;
    History h = apply_jump(poly_pow_E(n_draws), seeded_history(seed));
    std::memcpy(hist_out, h.x, sizeof h.x);
    return FGT_OK;
}

const char *fgt_version(void) { return "fastglobetrotter_b200 0.1 (sm_100a)"; }

}  // extern "C"
