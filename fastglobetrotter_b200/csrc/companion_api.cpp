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
#include <condition_variable>
#include <cstring>
#include <deque>
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
// A samples file is streamed: one thread inflates it into blocks of whole lines (a few MB each), a pool of
// threads tokenises the painting lines of each block straight into run-length-encoded rows (fgt_run_t: the maximal
// runs of equal donor haplotype the chunk builder needs, reference C:446-455) -- so neither the inflated text nor a
// dense label matrix is ever held in memory, and reading stops after the last recipient that is needed.
struct LineBlock {
    std::vector<char> buf;
    size_t len = 0;
    long first_line = 0;                 // index (in the file) of the block's first line; line 0 is the header
    std::vector<size_t> starts;          // starts[i] .. starts[i+1]: line i of the block (newline included)
};

class LineReader {
  public:
    explicit LineReader(const std::string &path) : path_(path) {
        g_ = gzopen(path.c_str(), "r");
        if (g_) gzbuffer(g_, 1 << 20);
    }
    ~LineReader() { if (g_) gzclose(g_); }
    bool ok() const { return g_ != nullptr; }
    // next block of newline-terminated lines (at least one, about `target` bytes); false at the end of the file.
    // An unterminated tail is kept in tail() like the reference's gzgets would deliver it last.
    bool next(LineBlock &blk, size_t target) {
        blk.starts.clear();
        blk.len = 0;
        blk.first_line = next_line_;
        if (eof_ && carry_.empty()) return false;
        std::vector<char> &b = blk.buf;
        b.assign(carry_.begin(), carry_.end());
        carry_.clear();
        size_t last_nl = std::string::npos;
        for (size_t i = b.size(); i-- > 0;) if (b[i] == '\n') { last_nl = i; break; }
        while (!eof_ && (last_nl == std::string::npos || b.size() < target)) {
            const size_t old = b.size();
            const size_t want = std::max<size_t>(target, 1 << 20);
            b.resize(old + want);
            const int n = gzread(g_, b.data() + old, (unsigned)want);
            if (n < 0) {
                int errnum = 0;
                const char *msg = gzerror(g_, &errnum);
                die("Something wrong with %s (%s). Exiting....\n", path_.c_str(), msg ? msg : "read error");
            }
            b.resize(old + (size_t)n);
            if ((size_t)n < want && (n == 0 || gzeof(g_))) {
                eof_ = true;
                int errnum = 0;
                gzerror(g_, &errnum);                      // a truncated or corrupt stream ends "early" with an error code set
                if (errnum < 0)
                    die("Something wrong with %s (corrupt or truncated gzip stream). Exiting....\n", path_.c_str());
            }
            for (size_t i = b.size(); i-- > old;) if (b[i] == '\n') { last_nl = i; break; }
        }
        if (last_nl == std::string::npos) {               // only an unterminated tail is left
            tail_.assign(b.begin(), b.end());
            b.clear();
            return false;
        }
        carry_.assign(b.begin() + last_nl + 1, b.end());
        if (eof_) { tail_ = carry_; carry_.clear(); }
        blk.len = last_nl + 1;
        blk.starts.push_back(0);
        const char *base = b.data(), *p = base, *end = base + blk.len;
        while (p < end) {
            const char *nl = (const char *)memchr(p, '\n', end - p);
            p = nl + 1;
            blk.starts.push_back((size_t)(p - base));
        }
        next_line_ += (long)blk.starts.size() - 1;
        return true;
    }
    long lines_seen() const { return next_line_; }
    const std::string &tail() const { return tail_; }

  private:
    std::string path_, carry_, tail_;
    gzFile g_ = nullptr;
    bool eof_ = false;
    long next_line_ = 0;
};

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

// one painting line -> maximal runs.  The first token (sample index) is skipped like reading(&step,"%s",waste), then
// nsites integers follow (C:443-445).  Returns false if the line holds fewer than nsites integers.
bool parse_runs(const char *p, const char *e, int nsites, std::vector<fgt_run_t> &out, int &maxv) {
    p = skip_ws(p, e);
    p = skip_tok(p, e);
    int mx = maxv;
    int prev = 0;
    out.clear();
    for (int j = 0; j < nsites; j++) {
        while (p < e && (unsigned char)*p <= ' ') p++;
        if (p >= e) return false;
        bool neg = false;
        if (*p == '-') { neg = true; p++; } else if (*p == '+') p++;
        unsigned v = 0;
        const char *a = p;
        while (p < e && (unsigned)(*p - '0') <= 9u) { v = v * 10u + (unsigned)(*p - '0'); p++; }
        if (p == a) return false;
        while (p < e && (unsigned char)*p > ' ') p++;          // rest of a malformed token is ignored like sscanf does
        const int iv = neg ? -(int)v : (int)v;
        if (j == 0 || iv != prev) {
            fgt_run_t r;
            r.start = (uint32_t)j; r.hap = (uint32_t)iv;
            out.push_back(r);
            if (iv > mx) mx = iv;
            prev = iv;
        }
    }
    maxv = mx;
    return true;
}

int n_threads() {
    int t = Engine::get().opt.threads;
    if (t <= 0) t = (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(t, 64));
}

struct PackedFile {            // the recipients' paintings of one samples file, run-length encoded
    std::string path;
    long long fsize = 0, mtime_ns = 0;
    int nsites = 0, nsamples = 0, ploidy = 0, max_label = 0;
    std::vector<int> rec_index;
    std::vector<int64_t> run_off;        // [rows + 1], rows = recipients * ploidy * nsamples
    std::vector<fgt_run_t> runs;
    bool pinned = false;                 // `runs` is page-locked (cudaHostRegister) for asynchronous uploads
    size_t bytes() const { return runs.size() * sizeof(fgt_run_t) + run_off.size() * sizeof(int64_t); }
    ~PackedFile() { if (pinned) cudaHostUnregister(runs.data()); }
};

// everything found in a samples file when it is read without knowing the recipients yet (cold first call)
struct ParsedAll {
    long lines = 0;                                  // newline-terminated lines, header included
    int nsamples = 0;
    std::vector<std::pair<long, std::string>> haps;  // (line index, third token) of every line whose first token is HAP
    std::vector<std::vector<fgt_run_t>> rows;        // indexed by (line - 1); empty for HAP lines and skipped lines
    int max_label = 0;
};

int header_nsamples(const char *p, const char *e, const std::string &path) {
    // header: token after "nsamples" "=" (C:234-239); the reference reads at most 2047 characters of it
    if (e - p > 2047) e = p + 2047;
    std::string tok = token_at(p, e);
    while (tok != "nsamples") {
        if (p >= e) die("Something wrong with %s. Exiting....\n", path.c_str());
        tok = token_at(p, e);
    }
    token_at(p, e);
    const std::string v = token_at(p, e);
    return atoi(v.c_str());
}

// Stream one samples file.  `select` == nullptr: tokenise every painting line and note every HAP line (all-mode);
// otherwise only the lines of the file individuals in `select` (their position follows from ploidy and nsamples,
// reference C:418-445) and reading stops behind the last of them.
void stream_samples(const std::string &path, int ploidy, int nsamples_known, int nsites, const std::vector<int> *select,
                    int workers, ParsedAll &out) {
    LineReader rd(path);
    if (!rd.ok()) die("error opening %s\n", path.c_str());
    std::mutex mu;
    std::condition_variable cv_push, cv_pop;
    std::deque<std::shared_ptr<LineBlock>> q;
    bool done = false;
    const size_t q_cap = (size_t)workers * 2 + 2;
    std::atomic<long> failed_line{-1};
    std::atomic<int> gmax{0};
    int nsamples = nsamples_known;
    long last_needed = -1;                              // exclusive line bound in select mode
    std::unordered_map<long, int> wanted;               // file individual -> 1
    struct Piece { long line; std::vector<fgt_run_t> runs; };
    std::vector<std::vector<Piece>> found(workers);
    std::vector<std::vector<std::pair<long, std::string>>> haps(workers);
    auto per_ind = [&]() { return (long)ploidy * (nsamples + 1); };
    auto set_select = [&]() {
        if (!select) return;
        long mx = -1;
        for (int v : *select) { wanted.emplace((long)v, 1); mx = std::max<long>(mx, v); }
        last_needed = 1 + (mx + 1) * per_ind();
    };
    if (nsamples > 0) set_select();
    auto work = [&](int tid) {
        std::vector<fgt_run_t> tmp;
        int mx = 0;
        while (true) {
            std::shared_ptr<LineBlock> blk;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_pop.wait(lk, [&] { return !q.empty() || done; });
                if (q.empty()) break;
                blk = q.front();
                q.pop_front();
            }
            cv_push.notify_one();
            const char *base = blk->buf.data();
            const long nl = (long)blk->starts.size() - 1;
            for (long i = 0; i < nl; i++) {
                const long li = blk->first_line + i;
                if (li == 0) continue;
                const char *a = base + blk->starts[i], *e = base + blk->starts[i + 1];
                if (!select) {
                    const char *p = skip_ws(a, e);
                    if (e - p >= 4 && p[0] == 'H' && p[1] == 'A' && p[2] == 'P' && isspace((unsigned char)p[3])) {
                        p += 3;
                        token_at(p, e);
                        haps[tid].emplace_back(li, token_at(p, e));
                    }
                }
                const long rem = (li - 1) % per_ind();
                if (rem % (nsamples + 1) == 0) continue;                     // the HAP line of a haplotype
                if (select) {
                    if (li >= last_needed || !wanted.count((li - 1) / per_ind())) continue;
                }
                if (!parse_runs(a, e, nsites, tmp, mx)) {
                    if (select) { long exp = -1; failed_line.compare_exchange_strong(exp, li); }
                    continue;                                                // all-mode: a short line only matters if it is needed
                }
                found[tid].push_back(Piece{li, tmp});
            }
        }
        int seen = gmax.load();
        while (mx > seen && !gmax.compare_exchange_weak(seen, mx)) {}
    };
    std::vector<std::thread> th;
    bool started = false;
    auto start_workers = [&]() { for (int t = 0; t < workers; t++) th.emplace_back(work, t); started = true; };
    const size_t target = 4u << 20;
    while (true) {
        auto blk = std::make_shared<LineBlock>();
        if (!rd.next(*blk, target)) break;
        if (blk->first_line == 0) {
            const int ns = header_nsamples(blk->buf.data(), blk->buf.data() + blk->starts[1], path);
            if (ns <= 0) die("No painting samples in %s! Exiting....\n", path.c_str());
            if (nsamples <= 0) { nsamples = ns; set_select(); }
            else if (ns != nsamples) die("Something wrong with %s. Exiting....\n", path.c_str());
        }
        if (!started) start_workers();
        const long first = blk->first_line;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_push.wait(lk, [&] { return q.size() < q_cap; });
            q.push_back(std::move(blk));
        }
        cv_pop.notify_one();
        if (select && last_needed >= 0 && first + 1 >= last_needed) break;        // (conservative: this block was still queued)
        if (select && last_needed >= 0 && rd.lines_seen() >= last_needed) break;  // nothing behind this point is needed
    }
    { std::lock_guard<std::mutex> lk(mu); done = true; }
    cv_pop.notify_all();
    for (auto &t : th) t.join();
    if (rd.lines_seen() == 0) die("Something wrong with %s. Exiting....\n", path.c_str());
    out.lines = rd.lines_seen();
    out.nsamples = nsamples;
    out.max_label = gmax.load();
    if (failed_line >= 0) {
        const long li = failed_line, ind = (li - 1) / per_ind();
        const int s = (int)((li - 1) % per_ind() % (nsamples + 1)) - 1;
        die("Something wrong with %s at sample %d of individual %ld. Exiting....\n", path.c_str(), s + 1, ind + 1);
    }
    for (auto &hv : haps) out.haps.insert(out.haps.end(), hv.begin(), hv.end());
    std::sort(out.haps.begin(), out.haps.end());
    out.rows.clear();
    out.rows.resize((size_t)std::max<long>(0, out.lines - 1));
    for (auto &fv : found)
        for (auto &pc : fv)
            if (pc.line >= 1 && pc.line - 1 < (long)out.rows.size()) out.rows[pc.line - 1] = std::move(pc.runs);
}

// what the reference's scan of the first file concludes (C:244-297), from the HAP lines of an all-mode pass
SamplesMeta meta_from(const ParsedAll &pa, const std::string &path, int ploidy, int n_rec, char **rec_labels) {
    SamplesMeta m;
    m.nsamples = pa.nsamples;
    std::unordered_map<std::string, int> names;
    for (int i = 0; i < n_rec; i++) names.emplace(rec_labels[i], i);
    long rec_count = 0, hap_count = 0;
    m.rec_first_index.assign(n_rec, 0);
    for (const auto &hp : pa.haps) {
        if (names.count(hp.second)) {
            if (rec_count >= (long)ploidy * n_rec)
                die("There are more than %d recipient haplotypes in %s! Exiting....\n", n_rec * ploidy, path.c_str());
            m.rec_first_index[rec_count / ploidy] = (int)(hap_count / ploidy);
            rec_count++;
        }
        hap_count++;
    }
    m.count = (pa.lines - 1) + 1;                        // + the read that hits EOF (C:248-274)
    const long nhaps = m.count / (m.nsamples + 1);
    if (nhaps != ((double)m.count) / (m.nsamples + 1) && nhaps != ((double)m.count - 1) / (m.nsamples + 1))
        die("something wrong with file %s. Exiting....\n", path.c_str());
    const long ninds = nhaps / ploidy;
    if (ninds <= 0) die("No individuals in %s! Exiting....\n", path.c_str());
    if (rec_count != (long)n_rec * ploidy)
        die("Only %ld recipient haplotypes in %s, while expecting %d! Exiting....\n", rec_count, path.c_str(), n_rec * ploidy);
    return m;
}

// the recipients' rows (file individuals rec_index, in that order) of a parsed file -> one packed, contiguous file
void pack_rows(ParsedAll &pa, const std::string &path, int ploidy, int nsamples, int nsites, const std::vector<int> &rec_index,
               PackedFile &pf) {
    const long per_ind = (long)ploidy * (nsamples + 1);
    const size_t rows = rec_index.size() * (size_t)ploidy * nsamples;
    pf.run_off.assign(rows + 1, 0);
    size_t total = 0, row = 0;
    for (size_t r = 0; r < rec_index.size(); r++)
        for (int p = 0; p < ploidy; p++)
            for (int sidx = 0; sidx < nsamples; sidx++, row++) {
                const long li = 1 + (long)rec_index[r] * per_ind + (long)p * (nsamples + 1) + 1 + sidx;
                if (li - 1 >= (long)pa.rows.size() || li >= pa.lines)
                    die("Something wrong with %s -- perhaps does not contain %d individuals? Exiting....\n", path.c_str(), rec_index[r] + 1);
                if (pa.rows[li - 1].empty())
                    die("Something wrong with %s at sample %d of individual %d. Exiting....\n", path.c_str(), sidx + 1, rec_index[r] + 1);
                total += pa.rows[li - 1].size();
                pf.run_off[row + 1] = (int64_t)total;
            }
    pf.runs.resize(total);
    row = 0;
    int mx = 0;
    for (size_t r = 0; r < rec_index.size(); r++)
        for (int p = 0; p < ploidy; p++)
            for (int sidx = 0; sidx < nsamples; sidx++, row++) {
                const long li = 1 + (long)rec_index[r] * per_ind + (long)p * (nsamples + 1) + 1 + sidx;
                const auto &v = pa.rows[li - 1];
                std::copy(v.begin(), v.end(), pf.runs.begin() + pf.run_off[row]);
                for (const fgt_run_t &x : v) if ((int)x.hap > mx) mx = (int)x.hap;
            }
    pf.max_label = mx;
    pf.nsites = nsites; pf.nsamples = nsamples; pf.ploidy = ploidy; pf.rec_index = rec_index; pf.path = path;
}

// ---- per-process cache of decoded files, least recently used entries dropped beyond a byte budget ("cache_mb") ----
std::mutex g_cache_mu;
struct CacheEnt { std::shared_ptr<PackedFile> pf; unsigned long long tick; };
std::map<std::string, CacheEnt> g_cache;
unsigned long long g_cache_tick = 0;

std::shared_ptr<PackedFile> cache_get(const std::string &key) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) return nullptr;
    it->second.tick = ++g_cache_tick;
    return it->second.pf;
}

void cache_put(const std::string &key, const std::shared_ptr<PackedFile> &pf) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache[key] = CacheEnt{pf, ++g_cache_tick};
    const double budget = std::max(0.0, Engine::get().opt.cache_mb) * 1048576.0;
    double used = 0;
    for (auto &kv : g_cache) used += (double)kv.second.pf->bytes();
    while (used > budget && g_cache.size() > 1) {             // entries still referenced by a running call stay alive through
        auto victim = g_cache.end();                           // their shared_ptr; the cache only drops its own reference
        for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
            if (it->first != key && (victim == g_cache.end() || it->second.tick < victim->second.tick)) victim = it;
        if (victim == g_cache.end()) break;
        used -= (double)victim->second.pf->bytes();
        g_cache.erase(victim);
    }
}

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

// ---- on-disk, chunk-level painting format ----------------------------------------------------------
// (1) cache: with FGT_PACK_CACHE_DIR set, every decoded samples file (and what the scan of the first file learnt) is
//     also written there as a binary run-length-encoded file keyed by source path, size, mtime and recipient set; a
//     later process reads it back instead of inflating and tokenising the text.  Off by default: the reference
//     library writes no files.
// (2) interchange: fgt_pack_samples() converts a whole *.samples.out[.gz] into a self-contained "*.fgtrle" file (all
//     haplotypes with their names); a samples list may name such files directly, so biobank-scale inputs skip text.
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
    uint32_t key_len, reserved;
    int32_t nsites, nsamples, ploidy, max_label;
    uint64_t n_rec, n_rows, n_runs;
};

bool write_all(FILE *f, const void *p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
bool read_all(FILE *f, void *p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }

void disk_store(const char *dir, const std::string &key, const PackedFile &pf) {
    const std::string path = pack_path(dir, key, "fgtpack"), tmp = path + ".tmp" + std::to_string((long long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;                                   // the cache is best effort: failures never reach the caller
    PackHdr h{};
    memcpy(h.magic, "FGTPACK2", 8);
    h.key_len = (uint32_t)key.size();
    h.nsites = pf.nsites; h.nsamples = pf.nsamples; h.ploidy = pf.ploidy; h.max_label = pf.max_label;
    h.n_rec = pf.rec_index.size(); h.n_rows = pf.run_off.size() - 1; h.n_runs = pf.runs.size();
    bool ok = write_all(f, &h, sizeof h) && write_all(f, key.data(), key.size()) &&
              write_all(f, pf.rec_index.data(), pf.rec_index.size() * sizeof(int)) &&
              write_all(f, pf.run_off.data(), pf.run_off.size() * sizeof(int64_t)) &&
              write_all(f, pf.runs.data(), pf.runs.size() * sizeof(fgt_run_t));
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

bool runs_well_formed(const PackedFile &pf) {
    const size_t rows = pf.run_off.size() - 1;
    if (pf.run_off[0] != 0 || (size_t)pf.run_off[rows] != pf.runs.size()) return false;
    for (size_t r = 0; r < rows; r++) {
        const int64_t a = pf.run_off[r], b = pf.run_off[r + 1];
        if (b <= a || pf.runs[a].start != 0) return false;
        for (int64_t i = a + 1; i < b; i++)
            if (pf.runs[i].start <= pf.runs[i - 1].start || pf.runs[i].start >= (uint32_t)pf.nsites || pf.runs[i].hap == pf.runs[i - 1].hap)
                return false;
    }
    return true;
}

bool disk_load(const char *dir, const std::string &key, const std::string &src_path, PackedFile &pf) {
    FILE *f = fopen(pack_path(dir, key, "fgtpack").c_str(), "rb");
    if (!f) return false;
    PackHdr h{};
    bool ok = read_all(f, &h, sizeof h) && memcmp(h.magic, "FGTPACK2", 8) == 0 && h.key_len == key.size() &&
              h.n_rec < (1u << 28) && h.nsites > 0 && h.nsamples > 0 && h.ploidy > 0 &&
              h.n_rows == h.n_rec * (uint64_t)h.ploidy * (uint64_t)h.nsamples && h.n_runs >= h.n_rows &&
              h.n_runs <= h.n_rows * (uint64_t)h.nsites;
    if (ok) {
        std::string k(h.key_len, '\0');
        ok = read_all(f, &k[0], k.size()) && k == key;     // the key holds path, size, mtime, ploidy, sites, recipients
    }
    if (ok) {
        pf.rec_index.resize(h.n_rec);
        pf.run_off.resize(h.n_rows + 1);
        pf.runs.resize(h.n_runs);
        ok = read_all(f, pf.rec_index.data(), h.n_rec * sizeof(int)) && read_all(f, pf.run_off.data(), pf.run_off.size() * sizeof(int64_t)) &&
             read_all(f, pf.runs.data(), pf.runs.size() * sizeof(fgt_run_t));
        char extra;
        ok = ok && fread(&extra, 1, 1, f) == 0;            // nothing may follow
    }
    fclose(f);
    if (ok) { pf.nsites = h.nsites; ok = runs_well_formed(pf); }
    if (!ok) { pf.rec_index.clear(); pf.run_off.clear(); pf.runs.clear(); return false; }
    pf.nsamples = h.nsamples; pf.ploidy = h.ploidy;
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

// ---- self-contained run-length-encoded samples file ("*.fgtrle") ----------------------------------------------
struct RleHdr {
    char magic[8];                 // "FGTRLE01"
    int32_t nsites, nsamples;
    uint64_t n_haps, n_lines, n_runs, names_bytes;
};

bool has_suffix(const std::string &s, const char *suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// rows of a *.fgtrle file are indexed like the text file's lines: hap h, sample s -> row h * nsamples + s
bool rle_read(const std::string &path, int &nsites, ParsedAll &pa) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    RleHdr h{};
    bool ok = read_all(f, &h, sizeof h) && memcmp(h.magic, "FGTRLE01", 8) == 0 && h.nsites > 0 && h.nsamples > 0 &&
              h.n_haps < (1ull << 31) && h.n_runs < (1ull << 40) && h.names_bytes < (1ull << 32);
    std::vector<char> names;
    std::vector<int64_t> off;
    std::vector<fgt_run_t> runs;
    if (ok) {
        names.resize(h.names_bytes);
        off.resize(h.n_haps * (uint64_t)h.nsamples + 1);
        runs.resize(h.n_runs);
        ok = read_all(f, names.data(), names.size()) && read_all(f, off.data(), off.size() * sizeof(int64_t)) &&
             read_all(f, runs.data(), runs.size() * sizeof(fgt_run_t));
    }
    fclose(f);
    if (!ok) return false;
    nsites = h.nsites;
    pa.nsamples = h.nsamples;
    pa.lines = (long)h.n_lines;
    pa.haps.clear();
    const char *p = names.data(), *e = p + names.size();
    for (uint64_t i = 0; i < h.n_haps && p < e; i++) {
        const char *z = (const char *)memchr(p, 0, e - p);
        if (!z) return false;
        pa.haps.emplace_back(1 + (long)i * (h.nsamples + 1), std::string(p, z - p));
        p = z + 1;
    }
    if (pa.haps.size() != h.n_haps) return false;
    pa.rows.assign((size_t)std::max<long>(0, pa.lines - 1), {});
    int mx = 0;
    for (uint64_t hp = 0; hp < h.n_haps; hp++)
        for (int sidx = 0; sidx < h.nsamples; sidx++) {
            const uint64_t row = hp * h.nsamples + sidx;
            const long li = 1 + (long)hp * (h.nsamples + 1) + 1 + sidx;
            if (off[row + 1] < off[row] || (uint64_t)off[row + 1] > h.n_runs || li - 1 >= (long)pa.rows.size()) return false;
            pa.rows[li - 1].assign(runs.begin() + off[row], runs.begin() + off[row + 1]);
            for (const fgt_run_t &x : pa.rows[li - 1]) if ((int)x.hap > mx) mx = (int)x.hap;
        }
    pa.max_label = mx;
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
    L.maps.resize(L.num_chrom);
    L.files.resize(L.num_chrom);
    {
        // recombination maps are parsed once per (path, size, mtime): a 100k-row map costs ~10 ms of strtod per call otherwise
        static std::mutex mu;
        static std::map<std::string, std::shared_ptr<RecombMap>> map_cache;
        for (int h = 0; h < L.num_chrom; h++) {
            long long sz = 0, mt = 0;
            std::string key;
            if (use_cache && file_identity(rfiles[h], sz, mt)) key = rfiles[h] + "|" + std::to_string(sz) + "|" + std::to_string(mt);
            std::shared_ptr<RecombMap> m;
            if (!key.empty()) {
                std::lock_guard<std::mutex> lk(mu);
                auto it = map_cache.find(key);
                if (it != map_cache.end()) m = it->second;
            }
            if (!m) {
                m = std::make_shared<RecombMap>(read_recomb(rfiles[h]));
                if (!key.empty()) {
                    std::lock_guard<std::mutex> lk(mu);
                    if (map_cache.size() > 256) map_cache.clear();
                    map_cache[key] = m;
                }
            }
            L.maps[h] = *m;
        }
    }

    // up to four files are streamed at once (zlib is sequential per file), the tokeniser threads are shared out among them
    const int nt = n_threads();
    auto parse_all = [&](int h, int workers) {
        auto pa = std::make_shared<ParsedAll>();
        int nsites = (int)L.maps[h].pos.size();
        if (has_suffix(sfiles[h], ".fgtrle")) {
            int ns_file = 0;
            if (!rle_read(sfiles[h], ns_file, *pa)) die("error opening %s\n", sfiles[h].c_str());
            if (ns_file != nsites) die("Something wrong with %s at sample 1 of individual 1. Exiting....\n", sfiles[h].c_str());
        } else stream_samples(sfiles[h], ploidy, 0, nsites, nullptr, workers, *pa);
        return pa;
    };

    // first file: learn nsamples and which file individuals are the recipients (C:230-297)
    SamplesMeta meta;
    std::shared_ptr<ParsedAll> first_all;
    std::map<int, std::future<std::shared_ptr<ParsedAll>>> early;       // cold start: later files read in all-mode meanwhile
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
            // cold start: the recipients' positions are only known once the first file has been read to its end, so it is
            // tokenised whole (run-length-encoded rows are small) and up to three later files are read the same way
            // at the same time
            const int fan = std::max(1, std::min({nt, 4, L.num_chrom}));
            const int per = std::max(1, nt / fan);
            for (int h = 1; h < fan; h++) early[h] = std::async(std::launch::async, parse_all, h, per);
            first_all = parse_all(0, per);
            meta = meta_from(*first_all, sfiles[0], ploidy, n_rec, rec_labels);
            if (pack_dir()) disk_store_meta(pack_dir(), mk, meta);
        }
        if (use_cache) { std::lock_guard<std::mutex> lk(mu); meta_cache[mk] = meta; }
    }
    L.nsamples = meta.nsamples;
    L.rec_index = meta.rec_first_index;
    std::vector<int> need(meta.rec_first_index.begin(), meta.rec_first_index.begin() + std::min(n_rec, n_rec_needed));

    std::vector<std::string> keys(L.num_chrom);
    std::vector<char> cached(L.num_chrom, 0);
    std::vector<std::pair<long long, long long>> ident(L.num_chrom);
    for (int h = 0; h < L.num_chrom; h++) {
        long long sz = 0, mt = 0;
        if (!file_identity(sfiles[h], sz, mt)) die("error opening %s\n", sfiles[h].c_str());
        ident[h] = {sz, mt};
        keys[h] = cache_key(sfiles[h], sz, mt, ploidy, (int)L.maps[h].pos.size(), need);
        if (use_cache) {
            auto pf = cache_get(keys[h]);
            if (pf) { L.files[h] = pf; cached[h] = 1; }
        }
        if (!cached[h] && pack_dir()) {
            auto pf = std::make_shared<PackedFile>();
            if (disk_load(pack_dir(), keys[h], sfiles[h], *pf) && pf->nsamples == meta.nsamples && pf->rec_index == need) {
                pf->fsize = sz; pf->mtime_ns = mt;
                L.files[h] = pf; cached[h] = 1;
                if (use_cache) cache_put(keys[h], pf);
            }
        }
    }
    auto finish = [&](int h, ParsedAll &pa) {
        if (pa.nsamples != meta.nsamples) die("Something wrong with %s. Exiting....\n", sfiles[h].c_str());
        auto pf = std::make_shared<PackedFile>();
        pack_rows(pa, sfiles[h], ploidy, meta.nsamples, (int)L.maps[h].pos.size(), need, *pf);
        pf->fsize = ident[h].first; pf->mtime_ns = ident[h].second;
        L.files[h] = pf;
        if (pack_dir()) disk_store(pack_dir(), keys[h], *pf);
        if (use_cache) cache_put(keys[h], pf);
    };
    if (first_all) { if (!cached[0]) finish(0, *first_all); first_all.reset(); }
    for (auto &kv : early) {
        auto pa = kv.second.get();
        if (!cached[kv.first]) finish(kv.first, *pa);
    }
    early.clear();
    // the remaining files: only the recipients' lines are tokenised, reading stops behind the last of them
    std::vector<int> todo;
    for (int h = 0; h < L.num_chrom; h++) if (!L.files[h]) todo.push_back(h);
    if (!todo.empty()) {
        const int fan = std::max(1, std::min<int>({nt, 4, (int)todo.size()}));
        const int per = std::max(1, nt / fan);
        std::atomic<size_t> next{0};
        auto file_worker = [&]() {
            while (true) {
                const size_t q = next.fetch_add(1);
                if (q >= todo.size()) break;
                const int h = todo[q];
                ParsedAll pa;
                const int nsites = (int)L.maps[h].pos.size();
                if (has_suffix(sfiles[h], ".fgtrle")) {
                    int ns_file = 0;
                    if (!rle_read(sfiles[h], ns_file, pa)) die("error opening %s\n", sfiles[h].c_str());
                    if (ns_file != nsites) die("Something wrong with %s at sample 1 of individual 1. Exiting....\n", sfiles[h].c_str());
                } else stream_samples(sfiles[h], ploidy, meta.nsamples, nsites, &need, per, pa);
                finish(h, pa);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < fan; t++) th.emplace_back(file_worker);
        file_worker();
        for (auto &t : th) t.join();
    }
    return L;
}

// page-lock a cached file's runs once so that every call's upload is an asynchronous DMA (needs a CUDA context)
void pin_runs(PackedFile &pf) {
    if (pf.pinned || pf.runs.empty()) return;
    if (cudaHostRegister(pf.runs.data(), pf.runs.size() * sizeof(fgt_run_t), cudaHostRegisterDefault) == cudaSuccess) pf.pinned = true;
    else cudaGetLastError();
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

// the packed problem of a .C call: run-length-encoded chromosomes straight from the (cached) decoded files
void chroms_of(Loaded &L, std::vector<fgt_chrom_t> &ch, int &max_hap) {
    ch.assign(L.num_chrom, fgt_chrom_t{});
    max_hap = 0;
    for (int h = 0; h < L.num_chrom; h++) {
        PackedFile &pf = *L.files[h];
        pin_runs(pf);
        ch[h].nsites = pf.nsites;
        ch[h].pos = L.maps[h].pos.data();
        ch[h].rate = L.maps[h].rate.data();
        ch[h].run_off = pf.run_off.data();
        ch[h].runs = pf.runs.data();
        ch[h].runs_on_device = 0;
        // donor_label_vec has one entry per haplotype of the id file; its length is not passed through .C, so the
        // largest haplotype id that occurs in the paintings is the bound we need
        max_hap = std::max(max_hap, pf.max_label);
    }
}

void curves_from_files(int mode, int *ploidy, int *ind_id_vec, char **samp, char **recom, double *binwidth, int *num_bins,
                       int *grid_start, int *num_donors, int *donor_label_vec, int *num_inds_rec, char **rec_labels,
                       double *weight_min, double *cutoff, int *num_surrogates, double *temp_weights, double *means) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    const int N = *num_inds_rec;
    Loaded L = load_inputs(*ploidy, samp, recom, N, rec_labels, N);
    for (int i = 0; i < N * L.num_chrom; i++)
        if (ind_id_vec[i] < 0 || ind_id_vec[i] >= N) die("Something wrong with the individual index vector. Exiting....\n");
    int rc0 = Engine::get().prepare_device();              // page-locking the decoded runs needs the CUDA context
    if (rc0) fatal_rc(rc0);
    std::vector<fgt_chrom_t> ch;
    int max_hap = 0;
    chroms_of(L, ch, max_hap);
    fgt_problem_t pb{};
    pb.mode = mode; pb.ploidy = *ploidy; pb.nsamples = L.nsamples; pb.num_chrom = L.num_chrom;
    pb.n_packed_inds = N; pb.label_bytes = 4; pb.chrom = ch.data();
    pb.num_inds = N; pb.ind_rows = ind_id_vec; pb.num_inds_total = N; pb.pair_offset = nullptr;
    pb.binwidth = *binwidth; pb.num_bins = *num_bins; pb.grid_start = *grid_start; pb.num_donors = *num_donors;
    pb.donor_label_vec = donor_label_vec; pb.n_donor_labels = max_hap; pb.weight_min = *weight_min;
    pb.gridweight_max_cutoff = *cutoff; pb.num_surrogates = *num_surrogates; pb.temp_weights = temp_weights;
    int rc = Engine::data_read_multi(pb, means, nullptr);           // one device, or the "devices" option's worth of them
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
    const int Nc = std::min(N, 100);                        // reference C:687, C:704-705
    // all recipients are decoded (not just the first 100): data_read on the same files then finds them in the cache
    Loaded L = load_inputs(*ploidy, filenameSAMPread, filenameRECOMread, N, recipient_label_vec, N);
    int rc0 = Engine::get().prepare_device();
    if (rc0) fatal_rc(rc0);
    std::vector<fgt_chrom_t> ch;
    int max_hap = 0;
    chroms_of(L, ch, max_hap);
    std::vector<int32_t> rec_rows(Nc);
    for (int i = 0; i < Nc; i++) rec_rows[i] = i;
    fgt_problem_t pb{};
    pb.mode = 1; pb.ploidy = *ploidy; pb.nsamples = L.nsamples; pb.num_chrom = L.num_chrom; pb.n_packed_inds = N;
    pb.label_bytes = 4; pb.chrom = ch.data(); pb.num_inds = N; pb.num_inds_total = N;
    pb.binwidth = *binwidth; pb.num_bins = *num_bins; pb.grid_start = *grid_start; pb.num_donors = *num_donors;
    pb.donor_label_vec = donor_label_vec; pb.n_donor_labels = max_hap; pb.weight_min = *weight_min;
    pb.gridweight_max_cutoff = *gridweightMAX_cutoff; pb.num_surrogates = 1;
    int rc = Engine::data_read_null_multi(pb, rec_rows.data(), chrom_ind_res, nullptr);
    if (rc) fatal_rc(rc);
    if (std::isnan(chrom_ind_res[0])) die("NA values. Exiting....\n");    // reference C:1055
}

namespace {
struct ProcDist { bool on = false; int rank = 0, world = 1; NcclComm comm = nullptr; } g_proc;   // one process per device (torchrun)
}

int fgt_data_read_packed(const fgt_problem_t *prob, double *means, fgt_stats_t *stats) {
    if (!prob || !means) return FGT_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_api_mu);
    if (g_proc.on && !prob->pair_offset && prob->num_inds_total > prob->num_inds) {
        // this process is one shard of a multi-process call: pair counts and curves are exchanged inside the library
        DistCtx dc;
        dc.rank = g_proc.rank; dc.world = g_proc.world; dc.comm = g_proc.comm; dc.h0 = nullptr; dc.store_state = true;
        return Engine::get().data_read_packed(*prob, means, stats, nullptr, &dc);
    }
    return Engine::data_read_multi(*prob, means, stats);
}

int fgt_dist_unique_id(char *out128) {
    if (!out128) return FGT_ERR_ARG;
    if (!nccl().loaded) { fprintf(stderr, "fastGLOBETROTTERCompanion(B200): NCCL unavailable (%s)\n", nccl().why); return FGT_ERR_ARG; }
    NcclUniqueId id;
    if (nccl().GetUniqueId(&id) != 0) return FGT_ERR_CUDA;
    std::memcpy(out128, id.internal, sizeof id.internal);
    return FGT_OK;
}

int fgt_dist_init(int rank, int world, const char *id128) {
    if (!id128 || world < 1 || rank < 0 || rank >= world) return FGT_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_api_mu);
    if (!nccl().loaded) { fprintf(stderr, "fastGLOBETROTTERCompanion(B200): NCCL unavailable (%s)\n", nccl().why); return FGT_ERR_ARG; }
    int rc = Engine::get().prepare_device();            // the communicator lives on the configured device ("device" option)
    if (rc) return rc;
    if (g_proc.on && g_proc.comm) { nccl().CommDestroy(g_proc.comm); g_proc = ProcDist{}; }
    NcclUniqueId id;
    std::memcpy(id.internal, id128, sizeof id.internal);
    NcclComm c = nullptr;
    const int nrc = nccl().CommInitRank(&c, world, id, rank);
    if (nrc != 0) {
        fprintf(stderr, "fastGLOBETROTTERCompanion(B200): ncclCommInitRank failed (%s)\n", nccl().GetErrorString ? nccl().GetErrorString(nrc) : "?");
        return FGT_ERR_CUDA;
    }
    g_proc.on = world > 1; g_proc.rank = rank; g_proc.world = world; g_proc.comm = c;
    return FGT_OK;
}

void fgt_dist_finalize(void) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    if (g_proc.comm && nccl().loaded) nccl().CommDestroy(g_proc.comm);
    g_proc = ProcDist{};
}

int fgt_count_pairs_packed(const fgt_problem_t *prob, int64_t *out) {
    if (!prob || !out) return FGT_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_api_mu);
    return Engine::get().data_read_packed(*prob, nullptr, nullptr, out);
}

int fgt_data_read_null_packed(const fgt_problem_t *prob, const int32_t *rec_rows, double *chrom_ind_res, fgt_stats_t *stats) {
    if (!prob || !rec_rows || !chrom_ind_res) return FGT_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_api_mu);
    if (g_proc.on)      // one process per GPU: this rank folds its rows of the tensor, the partial tensors are all-reduced inside the library
        return Engine::get().data_read_null_packed(*prob, rec_rows, chrom_ind_res, stats, g_proc.world, g_proc.rank, g_proc.comm);
    return Engine::data_read_null_multi(*prob, rec_rows, chrom_ind_res, stats);
}

// ---- callers of the path: opt-in entry points -----------------------------------------------------------
int fgt_null_normalise(const double *temp_weights, const double *null_mat, int num_surrogates, int num_donors, int num_bins,
                       double *means, double *null_curves_out) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    return Engine::get().null_normalise(temp_weights, null_mat, num_surrogates, num_donors, num_bins, means, null_curves_out);
}

void null_normalise(double *temp_weights, double *null_mat, int *num_surrogates, int *num_donors, int *num_bins, double *means) {
    // .C flavour of the above: every argument a pointer, `means` replaced in place, fatal errors end the process
    int rc = fgt_null_normalise(temp_weights, null_mat, *num_surrogates, *num_donors, *num_bins, means, nullptr);
    if (rc) fatal_rc(rc);
}

int fgt_getlhood_grid(const double *means, int n_curves, int n_times, const double *times, const double *admixtimes, int n_cand,
                      int nparam, const double *popfracs, double likpenalty, double *lhood_out) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    return Engine::get().getlhood_grid(means, n_curves, n_times, times, admixtimes, n_cand, nparam, popfracs, likpenalty, lhood_out);
}

int fgt_getmax(const double *means, int n_curves, int n_times, const double *times, int nparam, const double *popfracs, double *par_out,
               double *value_out, double *likpenalty_out) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    return Engine::get().getmax(means, n_curves, n_times, times, nparam, popfracs, par_out, value_out, likpenalty_out);
}

int fgt_bootstrap_means(const int32_t *ind_id_vec, int n_replicates, double *means_out) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    return Engine::get().bootstrap_means(ind_id_vec, n_replicates, means_out);
}

int64_t fgt_get_raw_counts(double *out, int64_t capacity) { return Engine::get_raw_all(out, capacity); }

void fgt_set_option_R(char **key, double *value) {          // .C flavour: .C("fgt_set_option_R", "devices", 8)
    if (key && *key && value) Engine::get().set_option(*key, *value);
}

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
    // what it decoded.  Lets the CPU test suite check the parser without a GPU.  The decoded form is run-length
    // encoded; labels_out receives it expanded to a dense matrix (uint16 if every id fits, else int32).
    if (!samples_list || !recom_list || !num_chrom) return FGT_ERR_ARG;
    char *sl = const_cast<char *>(samples_list), *rl = const_cast<char *>(recom_list);
    Loaded L = load_inputs(ploidy, &sl, &rl, n_rec, rec_labels, n_rec);
    *num_chrom = L.num_chrom;
    if (nsamples) *nsamples = L.nsamples;
    if (rec_index) for (int i = 0; i < n_rec; i++) rec_index[i] = L.rec_index[i];
    int mx = 0;
    for (int h = 0; h < L.num_chrom; h++) mx = std::max(mx, L.files[h]->max_label);
    const int lb = mx > 65535 ? 4 : 2;
    if (label_bytes) *label_bytes = lb;
    size_t off = 0;
    for (int h = 0; h < L.num_chrom; h++) {
        PackedFile &pf = *L.files[h];
        if (nsites) nsites[h] = pf.nsites;
        const size_t rows = pf.run_off.size() - 1;
        uint64_t sum = 0;
        for (size_t r = 0; r < rows; r++)
            for (int64_t i = pf.run_off[r]; i < pf.run_off[r + 1]; i++) {
                const uint32_t end = (i + 1 < pf.run_off[r + 1]) ? pf.runs[i + 1].start : (uint32_t)pf.nsites;
                sum += (uint64_t)pf.runs[i].hap * (end - pf.runs[i].start);
                if (labels_out && labels_out[h]) {
                    for (uint32_t j = pf.runs[i].start; j < end; j++) {
                        if (lb == 2) ((uint16_t *)labels_out[h])[r * pf.nsites + j] = (uint16_t)pf.runs[i].hap;
                        else ((int32_t *)labels_out[h])[r * pf.nsites + j] = (int32_t)pf.runs[i].hap;
                    }
                }
            }
        if (label_sum) label_sum[h] = sum;
        if (pos_out) std::memcpy(pos_out + off, L.maps[h].pos.data(), pf.nsites * sizeof(double));
        if (rate_out) std::memcpy(rate_out + off, L.maps[h].rate.data(), pf.nsites * sizeof(double));
        off += pf.nsites;
    }
    return FGT_OK;
}

int fgt_rle_encode(const void *labels, int label_bytes, int64_t rows, int nsites, int64_t *run_off, fgt_run_t *runs, int64_t capacity) {
    // host-only: dense label matrix -> run-length-encoded rows.  Returns the number of runs (needed capacity) or an error;
    // with runs == NULL only run_off (if given) and the count are produced.
    if (!labels || (label_bytes != 2 && label_bytes != 4) || rows < 0 || nsites <= 0) return FGT_ERR_ARG;
    int64_t n = 0;
    if (run_off) run_off[0] = 0;
    for (int64_t r = 0; r < rows; r++) {
        uint32_t prev = 0;
        for (int j = 0; j < nsites; j++) {
            const uint32_t v = label_bytes == 2 ? ((const uint16_t *)labels)[r * nsites + j] : (uint32_t)((const int32_t *)labels)[r * nsites + j];
            if (j == 0 || v != prev) {
                if (runs) {
                    if (n >= capacity) return FGT_ERR_ARG;
                    runs[n].start = (uint32_t)j; runs[n].hap = v;
                }
                n++;
                prev = v;
            }
        }
        if (run_off) run_off[r + 1] = n;
    }
    if (n > 0x7fffffffffLL) return FGT_ERR_ARG;
    return (int)std::min<int64_t>(n, 0x7fffffff);
}

int64_t fgt_pack_samples(const char *samples_path, int nsites, const char *out_path) {
    // host-only: *.samples.out[.gz] -> self-contained run-length-encoded "*.fgtrle" (every haplotype of the file with its
    // name).  A samples list may then name the .fgtrle file instead of the text file.  Returns the number of runs.
    if (!samples_path || !out_path || nsites <= 0) return FGT_ERR_ARG;
    ParsedAll pa;
    stream_samples(samples_path, 1, 0, nsites, nullptr, n_threads(), pa);
    RleHdr h{};
    memcpy(h.magic, "FGTRLE01", 8);
    h.nsites = nsites; h.nsamples = pa.nsamples;
    h.n_haps = pa.haps.size(); h.n_lines = (uint64_t)pa.lines;
    std::string names;
    std::vector<int64_t> off(1, 0);
    std::vector<fgt_run_t> runs;
    for (size_t i = 0; i < pa.haps.size(); i++) {
        names += pa.haps[i].second;
        names.push_back('\0');
        if (pa.haps[i].first != 1 + (long)i * (pa.nsamples + 1)) die("something wrong with file %s. Exiting....\n", samples_path);
        for (int sidx = 0; sidx < pa.nsamples; sidx++) {
            const long li = pa.haps[i].first + 1 + sidx;
            if (li - 1 >= (long)pa.rows.size() || pa.rows[li - 1].empty())
                die("Something wrong with %s at sample %d of individual %zu. Exiting....\n", samples_path, sidx + 1, i + 1);
            runs.insert(runs.end(), pa.rows[li - 1].begin(), pa.rows[li - 1].end());
            off.push_back((int64_t)runs.size());
        }
    }
    h.n_runs = runs.size(); h.names_bytes = names.size();
    const std::string tmp = std::string(out_path) + ".tmp" + std::to_string((long long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return FGT_ERR_ARG;
    bool ok = write_all(f, &h, sizeof h) && write_all(f, names.data(), names.size()) &&
              write_all(f, off.data(), off.size() * sizeof(int64_t)) && write_all(f, runs.data(), runs.size() * sizeof(fgt_run_t));
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), out_path) != 0) { remove(tmp.c_str()); return FGT_ERR_ARG; }
    return (int64_t)runs.size();
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
