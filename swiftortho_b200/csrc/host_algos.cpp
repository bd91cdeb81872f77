// Host-side steps of the search path that the reference also runs on the CPU: parameter tables,
// the query low-complexity mask, the deterministic quicksort order, bit score / e-value text.
// Each function names the reference lines whose behaviour it reproduces.
#include <algorithm>
#include <cmath>

#include "common.h"

namespace so {

static thread_local char g_err[512];
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

// Standard BLOSUM62 for the 23 residue letters the reference table holds (lib/fsearch.py:330).
const signed char kB62[23][23] = {
    {4, -1, -2, -2, 0, -1, -1, 0, -2, -1, -1, -1, -1, -2, -1, 1, 0, -3, -2, 0, -2, -1, 0},
    {-1, 5, 0, -2, -3, 1, 0, -2, 0, -3, -2, 2, -1, -3, -2, -1, -1, -3, -2, -3, -1, 0, -1},
    {-2, 0, 6, 1, -3, 0, 0, 0, 1, -3, -3, 0, -2, -3, -2, 1, 0, -4, -2, -3, 3, 0, -1},
    {-2, -2, 1, 6, -3, 0, 2, -1, -1, -3, -4, -1, -3, -3, -1, 0, -1, -4, -3, -3, 4, 1, -1},
    {0, -3, -3, -3, 9, -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1, -3, -3, -2},
    {-1, 1, 0, 0, -3, 5, 2, -2, 0, -3, -2, 1, 0, -3, -1, 0, -1, -2, -1, -2, 0, 3, -1},
    {-1, 0, 0, 2, -4, 2, 5, -2, 0, -3, -3, 1, -2, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1},
    {0, -2, 0, -1, -3, -2, -2, 6, -2, -4, -4, -2, -3, -3, -2, 0, -2, -2, -3, -3, -1, -2, -1},
    {-2, 0, 1, -1, -3, 0, 0, -2, 8, -3, -3, -1, -2, -1, -2, -1, -2, -2, 2, -3, 0, 0, -1},
    {-1, -3, -3, -3, -1, -3, -3, -4, -3, 4, 2, -3, 1, 0, -3, -2, -1, -3, -1, 3, -3, -3, -1},
    {-1, -2, -3, -4, -1, -2, -3, -4, -3, 2, 4, -2, 2, 0, -3, -2, -1, -2, -1, 1, -4, -3, -1},
    {-1, 2, 0, -1, -3, 1, 1, -2, -1, -3, -2, 5, -1, -3, -1, 0, -1, -3, -2, -2, 0, 1, -1},
    {-1, -1, -2, -3, -1, 0, -2, -3, -2, 1, 2, -1, 5, 0, -2, -1, -1, -1, -1, 1, -3, -1, -1},
    {-2, -3, -3, -3, -2, -3, -3, -3, -1, 0, 0, -3, 0, 6, -4, -2, -2, 1, 3, -1, -3, -3, -1},
    {-1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7, -1, -1, -4, -3, -2, -2, -1, -2},
    {1, -1, 1, 0, -1, 0, 0, 0, -1, -2, -2, 0, -1, -2, -1, 4, 1, -3, -2, -2, 0, 0, 0},
    {0, -1, 0, -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1, 5, -2, -2, 0, -1, -1, 0},
    {-3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1, -4, -3, -2, 11, 2, -3, -4, -3, -2},
    {-2, -2, -2, -3, -2, -1, -2, -3, 2, -1, -1, -2, -1, 3, -3, -2, -2, 2, 7, -1, -3, -2, -1},
    {0, -3, -3, -3, -1, -2, -2, -3, -3, 3, 1, -2, 1, -1, -2, -2, 0, -3, -1, 4, -3, -2, -1},
    {-2, -1, 3, 4, -3, 0, 1, -1, 0, -3, -4, 0, -3, -3, -2, 0, -1, -4, -3, -3, 4, 1, -1},
    {-1, 0, 0, 1, -3, 3, 4, -2, 0, -3, -3, 1, -1, -3, -1, 0, -1, -3, -2, -2, 1, 4, -1},
    {0, -1, -1, -1, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -2, 0, 0, -2, -1, -1, -1, -1, -1},
};

void make_code_table(uint8_t code[256]) {
    for (int i = 0; i < 256; i++) code[i] = kOther;
    for (int k = 0; k < 23; k++) {
        code[(unsigned char)kB62Letters[k]] = (uint8_t)k;
        code[(unsigned char)(kB62Letters[k] + 32)] = (uint8_t)k;
    }
}

void make_score_table(int8_t tbl[kClasses * kClasses]) {
    for (int a = 0; a < kClasses; a++)
        for (int b = 0; b < kClasses; b++) tbl[a * kClasses + b] = (a < 23 && b < 23) ? kB62[a][b] : -4;
}

int score_bytes(uint8_t a, uint8_t b) {
    static uint8_t code[256];
    static bool ready = false;
    if (!ready) {
        make_code_table(code);
        ready = true;
    }
    int x = code[a], y = code[b];
    return (x < 23 && y < 23) ? kB62[x][y] : -4;
}

static std::vector<std::string> split(const std::string &s, char c) {
    std::vector<std::string> out;
    size_t st = 0;
    for (;;) {
        size_t p = s.find(c, st);
        if (p == std::string::npos) {
            out.push_back(s.substr(st));
            return out;
        }
        out.push_back(s.substr(st, p - st));
        st = p + 1;
    }
}

// generate_nr_tbl (lib/fsearch.py:406-422): identity, then every letter of a group (both cases)
// maps to the smallest upper-case code of the group.
int parse_params(const so_params *p, Params &o) {
    if (!p || !p->seeds || !p->alphabets) {
        set_error("so_params: seeds/alphabets missing");
        return SO_EINVAL;
    }
    o.patterns = split(p->seeds, ',');
    o.mink = 1 << 30;
    o.maxk = 0;
    for (auto &s : o.patterns) {
        if (s.empty() || s.size() > 32) {
            set_error("seed pattern '%s': span must be 1..32", s.c_str());
            return SO_ELIMIT;
        }
        o.mink = std::min<int>(o.mink, (int)s.size());
        o.maxk = std::max<int>(o.maxk, (int)s.size());
    }
    if (o.patterns.size() > 16) {
        set_error("at most 16 seed patterns");
        return SO_ELIMIT;
    }
    o.alphabets.clear();
    for (auto &a : split(p->alphabets, '/')) {
        std::vector<uint16_t> t(256);
        for (int i = 0; i < 256; i++) t[i] = (uint16_t)i;
        std::string up = a;
        for (auto &ch : up) ch = (char)toupper((unsigned char)ch);
        for (auto &grp : split(up, ',')) {
            int lo = 1024;
            for (unsigned char ch : grp) lo = std::min<int>(lo, ch);
            for (unsigned char ch : grp) {
                t[ch] = (uint16_t)lo;
                t[(unsigned char)tolower(ch)] = (uint16_t)lo;
            }
        }
        o.alphabets.push_back(t);
    }
    if (o.alphabets.size() > 4) {
        set_error("at most 4 alphabets");
        return SO_ELIMIT;
    }
    if (p->n_buckets < 2) {
        set_error("-M (n_buckets) must be >= 2");
        return SO_EINVAL;
    }
    o.nc = p->n_buckets;
    o.step = p->step < 1 ? 1 : p->step;
    o.expect = p->expect;
    o.v = p->max_hits;
    o.max_miss = p->max_miss;
    o.thr = p->threshold;
    o.flt = p->filter_query != 0;
    o.chunk = p->chunk < 1 ? 50000 : p->chunk;
    o.rst = p->ref_start;
    o.red = p->ref_end;
    return SO_OK;
}

// ---------------------------------------------------------------------------------------------
// seg (lib/fsearch.py:2872-2928) with entropy (2854-2868) and the Counter quirk (157-177):
// the first-window tallies are 2*occurrences-1, the entropy is summed over letters in order of
// first appearance, and the sliding update keeps the reference's `b != 0 and X or Y` fall-through
// (a zero X selects Y).  All arithmetic in double, libm log, same operation order.
// ---------------------------------------------------------------------------------------------
// (k / 12) * log(k / 12) for integer-valued k in [0, 64), computed once with the expressions seg_mask uses
struct XLogXTable {
    double v[64];
    XLogXTable() {
        for (int k = 0; k < 64; k++) {
            double a = (double)k / 12.;
            v[k] = a * std::log(a);
        }
    }
};
static const XLogXTable kXLogX;
static inline double xlogx(double count, double a) {
    const int k = (int)count;
    if (k >= 0 && k < 64 && (double)k == count) return kXLogX.v[k];
    return a * std::log(a);
}

void seg_mask(const uint8_t *in, i64 n, uint8_t *out) {
    if (n <= 0) return;
    const double W = 12., MINENT = 2.2;
    std::vector<uint8_t> s((size_t)n);
    for (i64 i = 0; i < n; i++) s[(size_t)i] = (in[i] >= 'a' && in[i] <= 'z') ? (uint8_t)(in[i] - 32) : in[i];
    const double ln2 = std::log(2.0);
    double tally[256];
    bool known[256];
    for (int i = 0; i < 256; i++) tally[i] = 0, known[i] = false;
    uint8_t first_seen[12];
    int nfirst = 0;
    const i64 w0 = std::min<i64>(n, 12);
    for (i64 i = 0; i < w0; i++) {
        uint8_t c = s[(size_t)i];
        if (!known[c]) {
            known[c] = true;
            tally[c] = 0;  // Counter: first sight stores 0, later sights add 1
            first_seen[nfirst++] = c;
        } else
            tally[c] += 1;
    }
    for (i64 i = 0; i < w0; i++) tally[s[(size_t)i]] += 1.;
    double ent = 0;
    for (int k = 0; k < nfirst; k++) {
        double f = tally[first_seen[k]] / (double)w0;
        ent -= f * std::log(f);
    }
    ent /= std::log(2.0);
    std::vector<uint8_t> mask((size_t)n, 0);
    if (ent < MINENT) mask[0] = 1;
    for (i64 i = 1; i + 12 <= n; i++) {
        uint8_t out_c = s[(size_t)i - 1], in_c = s[(size_t)i + 11];
        if (out_c == in_c) {
            mask[(size_t)i] = mask[(size_t)i - 1];
            continue;
        }
        double before = tally[out_c];
        tally[out_c] -= 1;
        double in_before = known[in_c] ? tally[in_c] : 0.0;
        if (!known[in_c]) known[in_c] = true, tally[in_c] = 0;
        tally[in_c] += 1;
        // a * log(a) for a = k / 12 comes from a table filled with the same expression (same libm call, same
        // bits); tallies are small integers by construction
        double a = before / W, b = tally[out_c] / W;
        double alt = xlogx(before, a) / ln2;
        double delta = alt;
        if (b != 0) {
            double x = (xlogx(before, a) - xlogx(tally[out_c], b)) / ln2;
            if (x != 0) delta = x;
        }
        ent += delta;
        a = in_before / W;
        b = tally[in_c] / W;
        alt = -xlogx(tally[in_c], b) / ln2;
        delta = alt;
        if (a != 0) {
            double x = (xlogx(in_before, a) - xlogx(tally[in_c], b)) / ln2;
            if (x != 0) delta = x;
        }
        ent += delta;
        if (ent < MINENT) mask[(size_t)i] = 1;
    }
    i64 tail = std::max<i64>(0, n - 12);
    if (mask[(size_t)tail])
        for (i64 i = tail; i < n; i++) mask[(size_t)i] = 1;
    // unmasked: copy one residue; masked: emit 12 'x' and jump 12 (lib/fsearch.py:2918-2926)
    i64 st = 0, w = 0;
    while (st < n && w < n) {
        if (!mask[(size_t)st]) {
            out[w++] = s[(size_t)st];
            st += 1;
        } else {
            for (int k = 0; k < 12 && w < n; k++) out[w++] = 'x';
            st += 12;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The reference quicksort (lib/fsearch.py:260-327).  Rand.init_genrand(42) runs at the top of every
// quicksort() call, so random() is the constant below; ranges shorter than 7 use a stable
// insertion sort, a range of exactly 7 pivots on its middle, longer ranges on l+int(C*gap).
// Elements are packed (key << 32 | payload) and compared on the upper 32 bits only.
// ---------------------------------------------------------------------------------------------
static const double kPivotFrac = 0.3745401188473625;
static inline uint32_t keyof(uint64_t e) { return (uint32_t)(e >> 32); }

static void small_sort(uint64_t *x, i64 l, i64 r) {  // insort on [l, r]
    for (i64 i = l; i <= r; i++) {
        uint64_t v = x[i];
        uint32_t kv = keyof(v);
        i64 j = i - 1;
        while (j >= l && keyof(x[j]) > kv) {
            x[j + 1] = x[j];
            j--;
        }
        x[j + 1] = v;
    }
}

static i64 hoare(uint64_t *x, i64 l, i64 r) {  // partition(): pivot already at x[l]
    uint32_t pv = keyof(x[l]);
    i64 i = l, j = r + 1;
    for (;;) {
        do i++;
        while (i <= r && keyof(x[i]) < pv);
        do j--;
        while (keyof(x[j]) > pv);
        if (i > j) break;
        std::swap(x[i], x[j]);
    }
    std::swap(x[l], x[j]);
    return j;
}

static void quicksort_pruned(uint64_t *x, i64 n, i64 need) {
    struct Range {
        i64 l, r;
    };
    std::vector<Range> stack;
    stack.push_back(Range{0, n - 1});
    while (!stack.empty()) {
        Range t = stack.back();
        stack.pop_back();
        i64 l = t.l, r = t.r;
        if (r <= l || l >= need) continue;
        i64 gap = r - l + 1;
        if (gap < 7) {
            small_sort(x, l, r);
            continue;
        }
        i64 m = gap == 7 ? l + 3 : l + (i64)(kPivotFrac * (double)gap);
        std::swap(x[l], x[m]);
        i64 med = hoare(x, l, r);
        // the reference recurses left first, then right; the two sides are disjoint so the
        // processing order does not matter, only which sides are refined.
        stack.push_back(Range{med + 1, r});
        stack.push_back(Range{l, med - 1});
    }
}

void qsort_prefix(std::vector<uint64_t> &packed, i64 need) {
    quicksort_pruned(packed.data(), (i64)packed.size(), need);
}

void qsort_perm(const i64 *keys, i64 n, int32_t *perm) {
    if (n <= 0) return;
    // rank-compress the keys to 32 bits (order preserving) so the packed quicksort applies
    std::vector<i64> sorted(keys, keys + n);
    std::sort(sorted.begin(), sorted.end());
    sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    std::vector<uint64_t> x((size_t)n);
    for (i64 i = 0; i < n; i++) {
        uint64_t rk = (uint64_t)(std::lower_bound(sorted.begin(), sorted.end(), keys[i]) - sorted.begin());
        x[(size_t)i] = (rk << 32) | (uint32_t)i;
    }
    quicksort_pruned(x.data(), n, n);
    for (i64 i = 0; i < n; i++) perm[i] = (int32_t)(uint32_t)x[(size_t)i];
}

// score2bit / bit2e / f2s (lib/fsearch.py:1066-1071, 1086, 43-61)
i64 score2bit(i64 raw) { return (i64)((.267 * (double)raw + 3.1941832122778293) / 0.69314718055994529); }

double bit2e(i64 D, i64 ql, i64 tl, i64 bit) { return (double)(D * ql * tl) * std::pow(2, -(double)bit); }

// RPython str(float) is '%.6f'.  The *_to variants write into a caller buffer (>= 400 bytes) and return the length:
// so_write_rows formats ~10^5 rows per call and must not allocate per field.
static int six_to(char *b, double x) { return snprintf(b, 400, "%.6f", x); }

int f2s_to(char *out, double e) {
    if (e <= 0) {
        out[0] = '0', out[1] = 0;
        return 1;
    }
    if (e >= 1e-3) return six_to(out, e);
    double frac = std::log10(e);
    frac -= (double)(i64)frac;
    if (frac < 0) frac = 1 + frac;
    const double mant = std::pow(10, frac);
    char m[400], ex[400];
    int lm = six_to(m, mant), le = six_to(ex, std::log10(e / mant));
    const char *dm = (const char *)memchr(m, '.', (size_t)lm);
    const char *de = (const char *)memchr(ex, '.', (size_t)le);
    const int nm = dm ? std::min(lm, (int)(dm - m) + 3) : std::min(lm, 2);   // mantissa truncated to 2 decimals
    const int ne = de ? (int)(de - ex) : 0;                                   // integer part of the exponent text
    memcpy(out, m, (size_t)nm);
    out[nm] = 'e';
    memcpy(out + nm + 1, ex, (size_t)ne);
    out[nm + 1 + ne] = 0;
    return nm + 1 + ne;
}

int fmt_identity_to(char *out, double idy) {  // lib/fsearch.py:3235-3237
    int n;
    if (std::isnan(idy)) {
        memcpy(out, "nan", 4);
        n = 3;
    } else
        n = six_to(out, idy);
    const char *dot = (const char *)memchr(out, '.', (size_t)n);
    n = std::min(n, dot ? (int)(dot - out) + 3 : 2);
    out[n] = 0;
    return n;
}

std::string f2s(double e) {
    char b[400];
    const int n = f2s_to(b, e);
    return std::string(b, (size_t)n);
}

std::string fmt_identity(double idy) {
    char b[400];
    const int n = fmt_identity_to(b, idy);
    return std::string(b, (size_t)n);
}

}  // namespace so
