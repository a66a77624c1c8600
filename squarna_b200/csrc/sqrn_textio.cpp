// sqrn_textio.cpp -- bulk text lane of the CLI's single-sequence mode (host only, no CUDA).
//
// At >= 10^7 sequences/s on the device, reading the input and writing SQUARNA's text block per
// sequence in Python is the end-to-end limiter.  These two functions do what
// SQUARNA.py:80-256 (ParseDefaultInput / ParseFasta) and SQRNdbnseq.py:1301-1406 (the text of
// RunSQRNdbnseq) do for the PLAIN shape of an input -- name line + sequence, no reactivities,
// restraints or reference, one predicted structure per sequence -- on whole buffers.  Anything
// else is reported as SQRN_E_UNSUPPORTED and the caller takes the general (per-entry) path, so
// the text produced is the same either way (tests/test_host_python.py compares both).
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <system_error>
#include <thread>
#include <vector>
#include "../../include/sqrn.h"

// fn(0) .. fn(nt - 1), one host thread each.  A thread that cannot be started (std::system_error) does not
// abort the call: its share, and every later one, runs on the calling thread instead.
template <class F>
static void run_threads(int nt, F &fn)
{
    if (nt <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    int started = 0;
    try {
        th.reserve((size_t)nt);
        for (; started < nt; started++) th.emplace_back([&fn, started] { fn(started); });
    } catch (const std::system_error &) {
    } catch (const std::bad_alloc &) {
    }
    for (int t = started; t < nt; t++) fn(t);
    for (auto &x : th) x.join();
}

namespace {

// str.strip() / str.split() whitespace, ASCII part
inline bool is_ws(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }
inline bool is_gap(unsigned char c) { return c == '-' || c == '.' || c == '~'; }      // seq.py:12

// repr(round(x, 3)) for the magnitudes scores have: the decimal with at most three fraction digits,
// trailing zeros dropped but one kept ("12.0", "0.5", "187.935", "-0.0").  x is already a rounded value
// k / 1000, so k = llround(1000 x) is exact below 2^52 / 1000 and the digits come from integer arithmetic;
// anything else (huge, nan, inf) takes printf.
inline int fmt3(char *dst, double x)
{
    if (!(x > -4.0e12 && x < 4.0e12)) {
        int n = snprintf(dst, 48, "%.3f", x);
        while (n > 1 && dst[n - 1] == '0' && dst[n - 2] != '.') n--;
        return n;
    }
    int n = 0;
    if (x < 0.0 || (x == 0.0 && 1.0 / x < 0.0)) { dst[n++] = '-'; x = -x; }
    const long long k = (long long)(x * 1000.0 + 0.5);
    long long ip = k / 1000;
    const int fp = (int)(k % 1000);
    char tmp[24]; int t = 0;
    do { tmp[t++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
    while (t) dst[n++] = tmp[--t];
    dst[n++] = '.';
    dst[n++] = (char)('0' + fp / 100);
    if (fp % 100) { dst[n++] = (char)('0' + fp / 10 % 10); if (fp % 10) dst[n++] = (char)('0' + fp % 10); }
    return n;
}

}  // namespace

// One pass over the text.  multiline != 0: plain FASTA (a sequence may span lines, whole stripped
// lines are joined, SQUARNA.py:239-256); 0: SQUARNA's default format with the sequence on the
// first line after the name and nothing but blank lines after it (SQUARNA.py:80-203 with
// inputformat "q..."), the sequence being the first whitespace-separated token.
// Writes the counts always; fills the arrays when the capacities suffice, else SQRN_E_CAPACITY.
// name_begin/name_len: the stripped '>' line; seq_offsets[n+1] + seq: the sequence tokens.
static int parse_segment(const char *text, int64_t len, int multiline, int64_t *n_entries, int64_t *total_seq,
                         int64_t cap_entries, int64_t cap_seq, int64_t *name_begin, int32_t *name_len,
                         int64_t *seq_offsets, uint8_t *seq)
{
    if (!text || len < 0 || !n_entries || !total_seq) return SQRN_E_BADARG;
    const bool fill = name_begin && name_len && seq_offsets && seq;
    int64_t n = 0, tot = 0;
    bool overflow = false;
    int data_lines = 0;              // non-blank data lines of the current entry
    int raw_lines = 0;               // all its data lines, blank ones included
    bool in_entry = false;
    int64_t p = 0;
    while (p < len) {
        // one line [p, e), newline at e (or end of text)
        const char *nl = (const char *)memchr(text + p, '\n', (size_t)(len - p));
        const int64_t e = nl ? (int64_t)(nl - text) : len;
        int64_t a = p, b = e;
        for (int64_t q = p; q < e; q++) {
            const unsigned char c = (unsigned char)text[q];
            if (c >= 0x80) return SQRN_E_UNSUPPORTED;                 // not ASCII: characters != bytes
            if (c == '\r' && q + 1 != e) return SQRN_E_UNSUPPORTED;   // a lone CR is a line break for Python
        }
        while (a < b && is_ws((unsigned char)text[a])) a++;
        while (b > a && is_ws((unsigned char)text[b - 1])) b--;
        if (p < e && text[p] == '>') {                                // startswith('>') on the raw line
            if (in_entry && !multiline && data_lines != 1) return SQRN_E_UNSUPPORTED;
            if (n < cap_entries && fill) { name_begin[n] = a; name_len[n] = (int32_t)(b - a); seq_offsets[n] = tot; }
            else overflow = true;
            n++;
            in_entry = true; data_lines = 0; raw_lines = 0;
        } else if (a < b) {                                           // a non-blank data line
            if (!in_entry) {
                if (multiline) { p = e + 1; continue; }               // FASTA: text before the first '>' is dropped
                return SQRN_E_UNSUPPORTED;                            // default reactivities / restraints / reference lines
            }
            int64_t tb = b;
            if (!multiline) {
                // the sequence is the FIRST line after the name (a blank one there is an error in the reference);
                // later non-blank lines are reactivities / restraints / reference
                if (raw_lines >= 1) return SQRN_E_UNSUPPORTED;
                tb = a;
                while (tb < b && !is_ws((unsigned char)text[tb])) tb++;   // first token; the rest is a comment
            }
            if (tot + (tb - a) <= cap_seq && fill && !overflow) memcpy(seq + tot, text + a, (size_t)(tb - a));
            else overflow = true;
            tot += tb - a;
            data_lines++; raw_lines++;
        } else if (in_entry) raw_lines++;
        p = e + 1;
    }
    if (in_entry && !multiline && data_lines != 1) return SQRN_E_UNSUPPORTED;
    *n_entries = n; *total_seq = tot;
    if (!fill || overflow || n > cap_entries) return SQRN_E_CAPACITY;
    seq_offsets[n] = tot;
    return SQRN_OK;
}


// The public entry: long texts are cut at entry boundaries ("\n>") into a few segments that are counted, then
// filled, by one host thread each.
extern "C" int sqrn_text_parse(const char *text, int64_t len, int multiline, int64_t *n_entries, int64_t *total_seq,
                               int64_t cap_entries, int64_t cap_seq, int64_t *name_begin, int32_t *name_len,
                               int64_t *seq_offsets, uint8_t *seq)
{
    if (!text || len < 0 || !n_entries || !total_seq) return SQRN_E_BADARG;
    try {
    int nt = 1;
    if (len >= (4ll << 20)) {
        nt = (int)std::thread::hardware_concurrency();
        nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
    }
    std::vector<int64_t> cut((size_t)nt + 1, len);
    cut[0] = 0;
    for (int t = 1; t < nt; t++) {
        int64_t p = std::max(len * t / nt, cut[(size_t)t - 1]);
        const char *q = p < len ? (const char *)memchr(text + p, '\n', (size_t)(len - p)) : nullptr;
        while (q && q + 1 < text + len && q[1] != '>') q = (const char *)memchr(q + 1, '\n', (size_t)(text + len - (q + 1)));
        cut[(size_t)t] = (q && q + 1 < text + len) ? (int64_t)(q + 1 - text) : len;
    }
    std::vector<int64_t> cnt((size_t)nt, 0), tot((size_t)nt, 0);
    std::vector<int> rc((size_t)nt, SQRN_OK);
    auto run = [&](auto &&fn) { run_threads(nt, fn); };
    // pass 1: counts (and the shape check) per segment
    run([&](int t) {
        const int r = parse_segment(text + cut[(size_t)t], cut[(size_t)t + 1] - cut[(size_t)t], multiline, &cnt[(size_t)t], &tot[(size_t)t],
                                    0, 0, nullptr, nullptr, nullptr, nullptr);
        rc[(size_t)t] = (r == SQRN_E_CAPACITY) ? SQRN_OK : r;
    });
    int64_t n = 0, total = 0;
    std::vector<int64_t> n0((size_t)nt + 1, 0), s0((size_t)nt + 1, 0);
    for (int t = 0; t < nt; t++) {
        if (rc[(size_t)t] != SQRN_OK) return rc[(size_t)t];
        n += cnt[(size_t)t]; total += tot[(size_t)t];
        n0[(size_t)t + 1] = n; s0[(size_t)t + 1] = total;
    }
    if (n == 0) return SQRN_E_UNSUPPORTED;
    *n_entries = n; *total_seq = total;
    if (!(name_begin && name_len && seq_offsets && seq) || n > cap_entries || total > cap_seq) return SQRN_E_CAPACITY;
    // pass 2: fill, every segment into its own slice of the arrays
    run([&](int t) {
        if (!cnt[(size_t)t]) return;
        int64_t c = 0, s_ = 0;
        int64_t *so = seq_offsets + n0[(size_t)t];
        std::vector<int64_t> local((size_t)cnt[(size_t)t] + 1);
        rc[(size_t)t] = parse_segment(text + cut[(size_t)t], cut[(size_t)t + 1] - cut[(size_t)t], multiline, &c, &s_, cnt[(size_t)t], tot[(size_t)t],
                                      name_begin + n0[(size_t)t], name_len + n0[(size_t)t], local.data(), seq + s0[(size_t)t]);
        for (int64_t k = 0; k < cnt[(size_t)t]; k++) {
            name_begin[n0[(size_t)t] + k] += cut[(size_t)t];
            so[k] = local[(size_t)k] + s0[(size_t)t];
        }
    });
    for (int t = 0; t < nt; t++) if (rc[(size_t)t] != SQRN_OK) return rc[(size_t)t];
    seq_offsets[n] = total;
    return SQRN_OK;
    } catch (...) { return SQRN_E_NOMEM; }      // out of host memory: nothing crosses the C boundary
}

// UnAlign (seq.py:236-255) of every parsed sequence in one pass: the symbols without the gap characters
// "-.~" and their CSR offsets (what the prediction runs on).  sym must hold seq_offsets[n] bytes.
extern "C" int sqrn_text_ungap(int64_t n, const int64_t *seq_offsets, const uint8_t *seq, int64_t *sym_offsets, uint8_t *sym)
{
    if (n < 0 || !seq_offsets || !seq || !sym_offsets || !sym) return SQRN_E_BADARG;
    try {
    // ranges of entries, one host thread each: count the symbols that stay, then copy them to where the range starts
    int nt = 1;
    if (seq_offsets[n] >= (4ll << 20)) {
        nt = (int)std::thread::hardware_concurrency();
        nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
    }
    std::vector<int64_t> kept((size_t)nt + 1, 0);
    auto count = [&](int t) {
        const int64_t k0 = n * t / nt, k1 = n * (t + 1) / nt;
        int64_t c = 0;
        for (int64_t q = seq_offsets[k0]; q < seq_offsets[k1]; q++) c += !is_gap(seq[q]);
        kept[(size_t)t + 1] = c;
    };
    run_threads(nt, count);
    for (int t = 0; t < nt; t++) kept[(size_t)t + 1] += kept[(size_t)t];
    auto fill = [&](int t) {
        const int64_t k0 = n * t / nt, k1 = n * (t + 1) / nt;
        int64_t w = kept[(size_t)t];
        const bool no_gaps = kept[(size_t)t + 1] - kept[(size_t)t] == seq_offsets[k1] - seq_offsets[k0];
        if (no_gaps && k1 > k0) memcpy(sym + w, seq + seq_offsets[k0], (size_t)(seq_offsets[k1] - seq_offsets[k0]));
        for (int64_t k = k0; k < k1; k++) {
            if (no_gaps) w += seq_offsets[k + 1] - seq_offsets[k];
            else
                for (int64_t c = seq_offsets[k]; c < seq_offsets[k + 1]; c++)
                    if (!is_gap(seq[c])) sym[w++] = seq[c];
            sym_offsets[k + 1] = w;
        }
    };
    sym_offsets[0] = 0;
    run_threads(nt, fill);
    return SQRN_OK;
    } catch (...) { return SQRN_E_NOMEM; }
}

namespace {

struct FormatJob {
    const char *text; const int64_t *name_begin; const int32_t *name_len; const int64_t *seq_offsets; const uint8_t *seq;
    const int64_t *sym_offsets; const uint8_t *dbn; const double *scores; const char *psname; size_t pl;
    const char *cons; int cl;
};

// exact size of one entry's block
inline int64_t block_size(const FormatJob &J, int64_t k)
{
    char tmp[48];
    const int64_t L = J.seq_offsets[k + 1] - J.seq_offsets[k];
    const double st = J.scores[3 * k + 1];
    // name \n seq \n ____ \n dbn <cons> ==== \n dbn \t#1\t total \t struct \t react \t psname \n
    return J.name_len[k] + 1 + 5 * L + 3 + J.cl + 4 + fmt3(tmp, J.scores[3 * k]) + 1 +
           (st == 0.0 ? 1 : fmt3(tmp, st)) + 1 + fmt3(tmp, J.scores[3 * k + 2]) + 1 + (int64_t)J.pl + 1;
}

// writes the block of entry k at out; returns its size, or -1 if the dot-bracket string does not fit the sequence
inline int64_t write_block(const FormatJob &J, int64_t k, char *out)
{
    const int64_t L = J.seq_offsets[k + 1] - J.seq_offsets[k];
    const uint8_t *s = J.seq + J.seq_offsets[k];
    const uint8_t *d = J.dbn + J.sym_offsets[k];
    const int64_t nd = J.sym_offsets[k + 1] - J.sym_offsets[k];
    int64_t w = 0;
    memcpy(out + w, J.text + J.name_begin[k], (size_t)J.name_len[k]); w += J.name_len[k]; out[w++] = '\n';
    memcpy(out + w, s, (size_t)L); w += L; out[w++] = '\n';
    memset(out + w, '_', (size_t)L); w += L; out[w++] = '\n';
    char *line = out + w;                       // the re-gapped dot-bracket line, written once and copied
    int64_t q = 0;
    for (int64_t c = 0; c < L; c++) {
        if (is_gap(s[c])) line[c] = '.';
        else { if (q >= nd) return -1; line[c] = (char)d[q++]; }
    }
    if (q != nd) return -1;
    w += L;
    memcpy(out + w, J.cons, (size_t)J.cl); w += J.cl;
    memset(out + w, '=', (size_t)L); w += L; out[w++] = '\n';
    memcpy(out + w, line, (size_t)L); w += L;
    memcpy(out + w, "\t#1\t", 4); w += 4;
    const double total = J.scores[3 * k], st = J.scores[3 * k + 1], re = J.scores[3 * k + 2];
    w += fmt3(out + w, total); out[w++] = '\t';
    if (st == 0.0) out[w++] = '0'; else w += fmt3(out + w, st);
    out[w++] = '\t';
    w += fmt3(out + w, re); out[w++] = '\t';
    memcpy(out + w, J.psname, J.pl); w += (int64_t)J.pl; out[w++] = '\n';
    return w;
}

}  // namespace

// The text block of RunSQRNdbnseq (SQRNdbnseq.py:1301-1406) for entries [first, first + count) of a
// parsed input whose prediction has ONE structure per sequence:
//   name / sequence / '_' * L / dbn \t top-<conslim>_consensus / '=' * L / dbn \t #1 \t total \t struct \t react \t <psname>
// seq_offsets + seq: the sequence tokens as parsed (gaps included); sym_offsets + dbn: the ungapped
// CSR the prediction ran on and its dot-bracket bytes (gap columns print '.', ReAlign, seq.py:210-233;
// separators were already put back by the kernel).  scores: 3 per sequence, already round(x, 3);
// a structure score of exactly 0 prints as the int 0 (seq.py:871).  Two passes (exact block sizes, then the
// text), each spread over a few host threads when there are many entries.
extern "C" int sqrn_text_format(int64_t first, int64_t count, const char *text, const int64_t *name_begin,
                                const int32_t *name_len, const int64_t *seq_offsets, const uint8_t *seq,
                                const int64_t *sym_offsets, const uint8_t *dbn, const double *scores, int conslim,
                                const char *psname, char *out, int64_t cap, int64_t *written)
{
    if (!text || !name_begin || !name_len || !seq_offsets || !seq || !sym_offsets || !dbn || !scores || !psname || !written ||
        count < 0)
        return SQRN_E_BADARG;
    try {
    char cons[48];
    FormatJob J{text, name_begin, name_len, seq_offsets, seq, sym_offsets, dbn, scores, psname, strlen(psname), cons,
                snprintf(cons, sizeof cons, "\ttop-%d_consensus\n", conslim)};
    int nthreads = 1;
    if (count >= 8192) {
        nthreads = (int)std::thread::hardware_concurrency();
        nthreads = nthreads < 1 ? 1 : (nthreads > 16 ? 16 : nthreads);
    }
    auto range = [&](int t, int64_t &lo, int64_t &hi) { lo = first + count * t / nthreads; hi = first + count * (t + 1) / nthreads; };
    std::vector<int64_t> start((size_t)count + 1);
    auto run = [&](auto &&fn) { run_threads(nthreads, fn); };
    run([&](int t) { int64_t lo, hi; range(t, lo, hi); for (int64_t k = lo; k < hi; k++) start[(size_t)(k - first) + 1] = block_size(J, k); });
    start[0] = 0;
    for (int64_t k = 0; k < count; k++) start[(size_t)k + 1] += start[(size_t)k];
    *written = start[(size_t)count];
    if (!out || cap < start[(size_t)count]) return SQRN_E_CAPACITY;
    std::vector<int> bad((size_t)nthreads, 0);
    run([&](int t) {
        int64_t lo, hi; range(t, lo, hi);
        for (int64_t k = lo; k < hi; k++)
            if (write_block(J, k, out + start[(size_t)(k - first)]) != start[(size_t)(k - first) + 1] - start[(size_t)(k - first)]) { bad[(size_t)t] = 1; return; }
    });
    for (int b : bad) if (b) return SQRN_E_BADARG;
    return SQRN_OK;
    } catch (...) { return SQRN_E_NOMEM; }
}

// ---------------------------------------------------------------- packed boundary format
// Host helpers of sqrn_fast_predict_packed_host (include/sqrn.h): 2-bit base codes in, 4-bit bracket codes out.
static int host_threads(int64_t work, int64_t per_thread)
{
    int nt = (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 16) nt = 16;
    const int64_t want = work / per_thread + 1;
    return (int)std::min<int64_t>(nt, want);
}

extern "C" int sqrn_pack_symbols(int64_t n_total, const uint8_t *symbols, uint8_t *packed, int64_t *n_other)
{
    if (n_total < 0 || (n_total && (!symbols || !packed))) return SQRN_E_BADARG;
    uint8_t tab[256];
    memset(tab, 4, sizeof tab);
    tab['A'] = tab['a'] = 0; tab['C'] = tab['c'] = 1; tab['G'] = tab['g'] = 2;
    tab['U'] = tab['u'] = tab['T'] = tab['t'] = 3;                 // T -> U as SQRNdbnseq does (seq.py:1004-1010)
    const int64_t nbytes = (n_total + 3) / 4;
    const int nt = host_threads(nbytes, 1 << 20);
    std::vector<int64_t> bad((size_t)nt, 0);
    auto fn = [&](int t) {
        const int64_t b0 = nbytes * t / nt, b1 = nbytes * (t + 1) / nt;
        int64_t nb = 0;
        for (int64_t b = b0; b < b1; b++) {
            unsigned v = 0;
            for (int q = 0; q < 4; q++) {
                const int64_t k = 4 * b + q;
                unsigned c = k < n_total ? tab[symbols[k]] : 0u;
                if (c > 3) { nb++; c = 0; }
                v |= c << (2 * q);
            }
            packed[b] = (uint8_t)v;
        }
        bad[(size_t)t] = nb;
    };
    run_threads(nt, fn);
    int64_t tot = 0;
    for (int64_t x : bad) tot += x;
    if (n_other) *n_other = tot;
    return SQRN_OK;
}

// DBNToPairs (SQRNdbnseq.py:172-207): one stack per bracket kind, closing brackets without a partner are ignored,
// pairs sorted.  text = n code points; open_cp / close_cp = the n_kinds bracket glyphs of PairsToDBN's alphabet (passed in:
// the alphabet lives in the Python mirror).  pairs must hold n / 2 pairs.
extern "C" int sqrn_dbn_pairs(int64_t n, const uint32_t *text, int32_t n_kinds, const uint32_t *open_cp, const uint32_t *close_cp,
                              int32_t *pairs, int64_t *n_pairs)
{
    if (n < 0 || n > 0x7fffffff || n_kinds < 0 || n_kinds > 127 || !n_pairs || (n && (!text || !pairs)) || (n_kinds && (!open_cp || !close_cp)))
        return SQRN_E_BADARG;
    try {
        int8_t ascii[128];                          // +k+1: opening bracket of kind k, -(k+1): closing
        memset(ascii, 0, sizeof ascii);
        for (int k = 0; k < n_kinds; k++) {
            if (open_cp[k] < 128) ascii[open_cp[k]] = (int8_t)(k + 1);
            if (close_cp[k] < 128) ascii[close_cp[k]] = (int8_t)-(k + 1);
        }
        std::vector<std::vector<int32_t>> stack((size_t)n_kinds);
        std::vector<std::pair<int32_t, int32_t>> found;
        for (int64_t p = 0; p < n; p++) {
            const uint32_t c = text[p];
            int kind = 0;
            if (c < 128) kind = ascii[c];
            else for (int k = 0; k < n_kinds; k++) {
                if (c == open_cp[k]) { kind = k + 1; break; }
                if (c == close_cp[k]) { kind = -(k + 1); break; }
            }
            if (kind > 0) stack[(size_t)kind - 1].push_back((int32_t)p);
            else if (kind < 0 && !stack[(size_t)(-kind) - 1].empty()) {
                found.emplace_back(stack[(size_t)(-kind) - 1].back(), (int32_t)p);
                stack[(size_t)(-kind) - 1].pop_back();
            }
        }
        std::sort(found.begin(), found.end());
        for (size_t k = 0; k < found.size(); k++) { pairs[2 * k] = found[k].first; pairs[2 * k + 1] = found[k].second; }
        *n_pairs = (int64_t)found.size();
    } catch (...) { return SQRN_E_NOMEM; }
    return SQRN_OK;
}

extern "C" int sqrn_codes_to_ascii(int64_t n, const int8_t *codes, uint8_t *ascii)
{
    if (n < 0 || (n && (!codes || !ascii))) return SQRN_E_BADARG;
    // PairsToDBN's bracket alphabet (SQRNdbnseq.py:142-143): level L opens with op[L-1] and closes with cl[L-1].  Its
    // levels 31..49 are Cyrillic letters: those come out as byte 0 and the caller redoes that text in a wider
    // encoding; levels beyond the alphabet print as '.'
    static const char op[] = "([{<ABCDEFGHIJKLMNOPQRSTUVWXYZ", cl[] = ")]}>abcdefghijklmnopqrstuvwxyz";
    uint8_t tab[256];
    memset(tab, '.', sizeof tab);
    for (int l = 1; l <= 49; l++) { tab[l] = l <= 30 ? (uint8_t)op[l - 1] : 0; tab[256 - l] = l <= 30 ? (uint8_t)cl[l - 1] : 0; }
    const int nt = host_threads(n, 1 << 20);
    auto fn = [&](int t) {
        const int64_t k0 = n * t / nt, k1 = n * (t + 1) / nt;
        for (int64_t k = k0; k < k1; k++) ascii[k] = tab[(uint8_t)codes[k]];
    };
    run_threads(nt, fn);
    return SQRN_OK;
}

extern "C" int sqrn_unpack_dbn(int64_t n_seqs, const uint32_t *offsets, const uint8_t *dbn_nib, uint8_t *dbn_ascii)
{
    if (n_seqs < 0 || (n_seqs && (!offsets || !dbn_nib || !dbn_ascii))) return SQRN_E_BADARG;
    static const char op[] = ".([{<ABC", cl[] = ".)]}>abc";
    const int nt = host_threads(n_seqs, 1 << 16);
    auto fn = [&](int t) {
        const int64_t s0 = n_seqs * t / nt, s1 = n_seqs * (t + 1) / nt;
        for (int64_t b = s0; b < s1; b++) {
            const int64_t o = offsets[b], n = (int64_t)offsets[b + 1] - o;
            const uint8_t *src = dbn_nib + (o >> 1) + b;
            uint8_t *dst = dbn_ascii + o;
            for (int64_t p = 0; p < n; p++) {
                const unsigned v = (src[p >> 1] >> (4 * (p & 1))) & 15u;
                dst[p] = (uint8_t)((v & 8) ? cl[v & 7] : op[v & 7]);
            }
        }
    };
    run_threads(nt, fn);
    return SQRN_OK;
}
