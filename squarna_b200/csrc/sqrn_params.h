// sqrn_params.h -- host-side digest of a parameter set (plain C++, no CUDA):
// symbol codes, pairing masks, weight table and the pow() look-up tables.
//
// Every pow() of the scoring formulas is evaluated HERE with the host libm, the
// same function CPython's float.__pow__ calls, so the device never has to
// reproduce glibc's rounding:
//   sdf_lut[k]  = (1 / (1 + k)) ** distcoef        SQRNdbnseq.py:726, k = |dots + bw*brackets - ideal|
//   of_lut[o]   = (1 / (1 + o)) ** orderpenalty    SQRNdbnseq.py:729
//   pw17_lut[k] = (0.5 k) ** 1.7                   SQRNdbnseq.py:884 (bp scores are multiples of 0.5)
#pragma once
#include <math.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/sqrn.h"
#include "sqrn_device.cuh"

namespace sqrn {

struct HostParams {
    DevParams p;                  // LUT pointers are left NULL: the caller patches them
    std::vector<double> lut;      // sdf | of | pw17, in this order
    size_t sdf_off, of_off, pw17_off;
};

inline bool build_host_params(const sqrn_paramset &ps, int nmax, HostParams &H, std::string &err)
{
    DevParams &P = H.p;
    memset(&P, 0, sizeof(P));
    if (ps.n_bp < 0 || ps.n_bp > SQRN_MAX_BPKEYS) { err = "bad number of bpweights entries"; return false; }
    for (int c = 0; c < 256; c++) P.code_table[c] = CODE_OTHER;
    const char *acgu = "ACGU";
    for (int k = 0; k < 4; k++) {
        P.code_table[(unsigned char)acgu[k]] = (uint8_t)k;
        P.code_table[(unsigned char)(acgu[k] + 32)] = (uint8_t)k;
    }
    P.code_table[(unsigned char)'T'] = CODE_U;      // seq.upper().replace("T", "U"), SQRNdbnseq.py:1004
    P.code_table[(unsigned char)'t'] = CODE_U;
    P.code_table[(unsigned char)';'] = CODE_SEP;    // SEPS, SQRNdbnseq.py:14
    P.code_table[(unsigned char)'&'] = CODE_SEP;
    int K = 6;
    // a key symbol can match the normalised sequence only if it is not a lower-case
    // letter and not 'T' (both are rewritten before the lookup)
    auto matchable = [](unsigned char c) { return !(c >= 'a' && c <= 'z') && c != 'T'; };
    for (int k = 0; k < ps.n_bp; k++) {
        unsigned char a = ps.bp_keys[2 * k], b = ps.bp_keys[2 * k + 1];
        if (!matchable(a) || !matchable(b)) continue;
        unsigned char ab[2] = { a, b };
        for (int q = 0; q < 2; q++) {
            unsigned char c = ab[q];
            if (c == ';' || c == '&') { err = "chain separators in bpweights are not supported"; return false; }
            if (P.code_table[c] == CODE_OTHER) {
                if (K >= MAXK) { err = "too many distinct symbols in bpweights"; return false; }
                P.code_table[c] = (uint8_t)K;
                if (c >= 'A' && c <= 'Z') P.code_table[c + 32] = (uint8_t)K;
                K++;
            }
        }
        int ca = P.code_table[a], cb = P.code_table[b];
        // both orientations, later entries overwrite (SQRNdbnseq.py:282-284)
        P.pairmask[ca] |= 1u << cb; P.pairmask[cb] |= 1u << ca;
        P.weight[ca * MAXK + cb] = ps.bp_vals[k];
        P.weight[cb * MAXK + ca] = ps.bp_vals[k];
    }
    P.K = K;
    P.npc = 0;
    for (int c = 0; c < K; c++)
        if (P.pairmask[c]) {
            if (P.npc >= MAXPC) { err = "too many pairing symbols in bpweights"; return false; }
            P.pc_code[P.npc++] = c;
        }
    P.std_pairs = (K == 6 && P.npc == 4 && P.pairmask[CODE_A] == (1u << CODE_U) && P.pairmask[CODE_C] == (1u << CODE_G) &&
                   P.pairmask[CODE_G] == ((1u << CODE_C) | (1u << CODE_U)) && P.pairmask[CODE_U] == ((1u << CODE_A) | (1u << CODE_G)));
    double m = ceil(ps.minlen);
    P.m = m < 1 ? 1 : (m > 32 ? 32 : (int)m);
    P.minlen = ps.minlen; P.minbpscore = ps.minbpscore;
    P.minfinscore = ps.minbpscore * ps.minfinscorefactor;          // SQRNdbnseq.py:1073
    P.loopbonus = ps.loopbonus; P.maxstemnum = ps.maxstemnum; P.bracketweight = ps.bracketweight;
    P.distcoef = ps.distcoef; P.orderpenalty = ps.orderpenalty;
    P.bw_is_int = (ps.bracketweight == rint(ps.bracketweight) && fabs(ps.bracketweight) <= 64.0);
    P.bw_int = P.bw_is_int ? (int)ps.bracketweight : 0;
    int absbw = P.bw_int < 0 ? -P.bw_int : P.bw_int;
    P.sdf_n = P.bw_is_int ? (1 + absbw) * nmax + 8 : 0;
    P.of_n = 66;
    P.pw17_n = 4 * nmax + 8;
    H.lut.clear();
    H.sdf_off = 0;
    for (int k = 0; k < P.sdf_n; k++) H.lut.push_back(pow(1.0 / (1.0 + fabs((double)k)), ps.distcoef));
    H.of_off = H.lut.size();
    for (int o = 0; o < P.of_n; o++) H.lut.push_back(pow(1.0 / (double)(1 + o), ps.orderpenalty));
    H.pw17_off = H.lut.size();
    for (int k = 0; k < P.pw17_n; k++) H.lut.push_back(pow(0.5 * k, 1.7));
    // factor maxima for score_bound(): valid when every factor comes from the tables above
    P.ub_ok = P.bw_is_int ? 1 : 0;
    P.sdf_max = 1.0;                                   // "between chains" gives exactly 1 (SQRNdbnseq.py:723)
    for (int k = 0; k < P.sdf_n; k++) if (H.lut[H.sdf_off + k] > P.sdf_max) P.sdf_max = H.lut[H.sdf_off + k];
    P.of_max = 0.0;
    for (int o = 0; o < P.of_n; o++) if (H.lut[H.of_off + o] > P.of_max) P.of_max = H.lut[H.of_off + o];
    if (!(P.sdf_max < 1e300) || !(P.of_max < 1e300) || !(P.of_max > 0.0)) P.ub_ok = 0;
    // loopfactor = 1 + lb*g1*(2 - d1/2) + lb*g2*(2 - d2/2), g in {0,1}, (2 - d/2) in {2, 1.5, 1} (SQRNdbnseq.py:715)
    {
        double lb = ps.loopbonus, best = 1.0;
        const double cs[4] = { 0.0, 2.0, 1.5, 1.0 };   // 0.0 = no good loop
        for (int x = 0; x < 4; x++)
            for (int y = 0; y < 4; y++) {
                double lf = 1.0 + (lb * (x ? 1.0 : 0.0)) * (x ? cs[x] : 2.0);
                lf = lf + (lb * (y ? 1.0 : 0.0)) * (y ? cs[y] : 2.0);
                if (lf > best) best = lf;
            }
        P.lf_max = best;
        if (!(best < 1e300)) P.ub_ok = 0;
    }
    return true;
}

// Python round(x, 3): the correctly rounded 3-decimal string, read back as a
// double (float.__round__ uses dtoa mode 3).  Fast path when x*1000 is far
// from a rounding tie; snprintf/strtod otherwise.
inline double pyround3(double x)
{
    if (!(fabs(x) < 1e9)) {
        if (!isfinite(x)) return x;
        char buf[400]; snprintf(buf, sizeof buf, "%.3f", x); return strtod(buf, nullptr);
    }
    // y carries a relative error <= 2^-53: k is the integer nearest to the exact product whenever
    // y is further than that from the half-way point (same test as round3_fast on the device)
    double y = x * 1000.0, k = rint(y);
    if (fabs(y - k) < 0.4999999 - fabs(y) * 1e-15) return k / 1000.0;
    char buf[64]; snprintf(buf, sizeof buf, "%.3f", x); return strtod(buf, nullptr);
}

// reactivity factors of BPMatrix (SQRNdbnseq.py:333-336) over the distinct
// processed reactivities of a batch: pos[a*R+b] = ((1-(ra+rb)/2)*2)**0.5 and
// neg[a*R+b] = 1/max(pos, 0.01) (used when the pair weight is <= 0).
inline void build_react_lut(const double *rvals, int R, std::vector<double> &pos, std::vector<double> &neg)
{
    pos.resize((size_t)R * R); neg.resize((size_t)R * R);
    for (int a = 0; a < R; a++)
        for (int b = 0; b < R; b++) {
            double rf = pow((1.0 - (rvals[a] + rvals[b]) / 2.0) * 2.0, 0.5);
            pos[(size_t)a * R + b] = rf;
            neg[(size_t)a * R + b] = 1.0 / (rf > 0.01 ? rf : 0.01);
        }
}

// restraint pairs of every sequence sorted by (v + w, v): the order in which the
// anti-diagonal walk meets them
inline void sort_rbps(int64_t n_seqs, const int64_t *rbp_off, const int32_t *rbp, std::vector<int32_t> &out)
{
    out.assign(rbp, rbp + 2 * rbp_off[n_seqs]);
    for (int64_t b = 0; b < n_seqs; b++) {
        int64_t lo = rbp_off[b], n = rbp_off[b + 1] - lo;
        int32_t *p = out.data() + 2 * lo;
        for (int64_t a = 1; a < n; a++) {             // insertion sort: lists are short
            int32_t v = p[2 * a], w = p[2 * a + 1];
            int64_t c = a - 1;
            while (c >= 0 && (p[2 * c] + p[2 * c + 1] > v + w || (p[2 * c] + p[2 * c + 1] == v + w && p[2 * c] > v))) {
                p[2 * c + 2] = p[2 * c]; p[2 * c + 3] = p[2 * c + 1]; c--;
            }
            p[2 * c + 2] = v; p[2 * c + 3] = w;
        }
    }
}

}  // namespace sqrn
