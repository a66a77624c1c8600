// sqrn_emu.cpp -- TEST INFRASTRUCTURE: single-thread host build of the device
// functions in squarna_b200/csrc/sqrn_device.cuh (SQRN_HOST_EMU, team size 1).
// It lets `pytest -m "not gpu"` exercise the bit-mask enumeration, the scoring
// and the level logic against the oracle on a box without a GPU.  It is not
// linked into libsqrn_b200.so and nothing in squarna_b200/ loads it; the
// parallel behaviour (ballots, shuffles, atomics, barriers) is only covered by
// the `-m gpu` tests.
#define SQRN_HOST_EMU 1
#include <stdlib.h>
#include "../../squarna_b200/csrc/sqrn_params.h"

using namespace sqrn;

extern "C" int emu_run(const sqrn_paramset *ps, int64_t n_seqs, const int64_t *off, const uint8_t *sym,
                       const uint16_t *rcode, const double *rvals, int R, int react_comp,
                       const uint8_t *rclass, const int64_t *rbp_off, const int32_t *rbp,
                       const double *smat, int L, const int32_t *cols, int interchainonly,
                       int mode, int n_items, const int32_t *item_seq, const int64_t *init_off,
                       const int32_t *init_stems, const double *item_subopt,
                       const int64_t *out_off, int32_t *out_stems, int32_t *out_nstems, double *out_stemfin,
                       double *out_raw, uint8_t *out_flags, const int64_t *dbn_off, uint8_t *dbn_ascii,
                       int8_t *dbn_code, int ccap, unsigned long long *n_calls, int region_mode, int flavour, int pcap)
{
    int nmax = 1, rbmax = 0;
    for (int64_t b = 0; b < n_seqs; b++) {
        int n = (int)(off[b + 1] - off[b]);
        if (n > nmax) nmax = n;
        if (rbp_off) { int q = (int)(rbp_off[b + 1] - rbp_off[b]); if (q > rbmax) rbmax = q; }
    }
    HostParams H; std::string err;
    if (!build_host_params(*ps, nmax, H, err)) return -1;
    H.p.sdf_lut = H.lut.data() + H.sdf_off;
    H.p.of_lut = H.lut.data() + H.of_off;
    H.p.pw17_lut = H.lut.data() + H.pw17_off;
    std::vector<double> rpos, rneg;
    std::vector<int32_t> rb_sorted;
    DevBatch B; memset(&B, 0, sizeof B);
    B.n_seqs = n_seqs; B.off = off; B.sym = sym;
    if (rcode) { build_react_lut(rvals, R, rpos, rneg); B.rcode = rcode; B.rf_pos = rpos.data(); B.rf_neg = rneg.data(); B.rvals = rvals; B.R = R; }
    B.react_comp = react_comp;
    B.rclass = rclass;
    if (rbp_off) { sort_rbps(n_seqs, rbp_off, rbp, rb_sorted); B.rbp_off = rbp_off; B.rbp = rb_sorted.data(); }
    B.smat = smat; B.L = L; B.cols = cols; B.interchainonly = interchainonly;
    DevWork W; memset(&W, 0, sizeof W);
    W.n_items = n_items; W.mode = mode; W.item_seq = item_seq; W.init_off = init_off; W.init_stems = init_stems;
    W.item_subopt = item_subopt; W.out_off = out_off; W.out_stems = out_stems; W.out_nstems = out_nstems;
    W.out_stemfin = out_stemfin; W.out_raw = out_raw; W.out_flags = out_flags; W.dbn_off = dbn_off;
    W.out_dbn_ascii = dbn_ascii; W.out_dbn_code = dbn_code; W.n_calls = n_calls; W.region_mode = region_mode;
    Layout Lay = make_layout(nmax, rbmax, ccap, H.p.npc, 0, 0, 1, 1, 0, (flavour == 3 || flavour == 4) ? pcap : ((flavour == 5 || flavour == 8) ? -1 : 0));
    std::vector<GEnt> ge; std::vector<double> gb; std::vector<uint8_t> gq;
    if (flavour == 5 || flavour == 8) {            // global persistent list with cached scores (what CTA teams run), capacity pcap (two halves)
        ge.resize(2 * (size_t)pcap + 2); gb.resize(2 * (size_t)pcap + 2); gq.resize(2 * (size_t)pcap + 2);
        W.g_ent = ge.data(); W.g_bps = gb.data(); W.g_qb = gq.data(); W.g_cap = pcap;
        W.g_rebuild = sqrn::g_emu_gl_rebuild_every;
    }
    unsigned char *smem = (unsigned char *)aligned_alloc(16, (size_t)Lay.total + 16);
    // flavour 6: the pool-round path of CTA teams -- a base list per sequence (MODE_BASE), then the items sweep it
    std::vector<BEnt> be; std::vector<int64_t> boff; std::vector<int32_t> bn, bb;
    if (flavour == 6) {
        boff.resize((size_t)n_seqs + 1);
        for (int64_t b = 0; b <= n_seqs; b++) boff[(size_t)b] = b * (int64_t)pcap;
        be.resize((size_t)n_seqs * pcap + 1); bn.assign((size_t)n_seqs, -1); bb.assign((size_t)n_seqs * (GL_NBIN + 1), 0);
        W.base_ent = be.data(); W.base_off = boff.data(); W.base_n = bn.data(); W.base_bend = bb.data();
        DevWork Wb = W; Wb.mode = MODE_BASE; Wb.item_seq = nullptr; Wb.init_off = nullptr; Wb.n_items = (int)n_seqs;
        for (int64_t b = 0; b < n_seqs; b++) {
            State S = bind_state(smem, Lay);
            team_run_item<Cfg<0>>(S, H.p, B, Wb, Lay, (int)b);
        }
        flavour = 0;
    }
    // PERSIST flavours park the items whose run list overflowed; they are redone by the rescanning flavour
    std::vector<int32_t> ovf((size_t)n_items + 1); int n_ovf = 0;
    W.ovf_list = ovf.data(); W.ovf_count = &n_ovf;
    for (int pass = 0; pass < 2; pass++) {
    const int todo = pass == 0 ? n_items : n_ovf;
    if (pass == 1) {
        // (the product redoes overflowed items with another kernel and that kernel's own layout: the global-list layout
        //  has no room for the lists of the rescanning pass)
        if (flavour == 5 || flavour == 8) { Lay = make_layout(nmax, rbmax, ccap, H.p.npc, 0, 0, 1, 1, 0, 0); free(smem); smem = (unsigned char *)aligned_alloc(16, (size_t)Lay.total + 16); }
        if (flavour == 3) flavour = 1; else if (flavour == 4 || flavour == 5 || flavour == 8) flavour = 2;
    }
    for (int q = 0; q < todo; q++) {
        const int item = pass == 0 ? q : ovf[q];
        State S = bind_state(smem, Lay);
        if (flavour == 1) {            // the fast-lane flavour (plain batch, standard pairing table, MODE_TAIL)
            if (!H.p.std_pairs || rcode || rclass || rbp_off || smat || interchainonly || mode != MODE_TAIL) { free(smem); return -2; }
            team_run_item<Cfg<0, true, true, MODE_TAIL, 1>>(S, H.p, B, W, Lay, item);
        } else if (flavour == 3) {     // the fast-lane flavour with the persistent run list (what k_fast runs)
            if (!H.p.std_pairs || rcode || rclass || rbp_off || smat || interchainonly || mode != MODE_TAIL) { free(smem); return -2; }
            team_run_item<Cfg<0, true, true, MODE_TAIL, 1, false, true>>(S, H.p, B, W, Lay, item);
        } else if (flavour == 4) team_run_item<Cfg<0, false, false, -1, 1, false, true>>(S, H.p, B, W, Lay, item);   // general flavour, persistent list
        else if (flavour == 5) team_run_item<Cfg<0, false, false, -1, 0, false, true, 1>>(S, H.p, B, W, Lay, item);   // global list
        else if (flavour == 8) team_run_item<Cfg<0, false, false, -1, 0, false, true, 2>>(S, H.p, B, W, Lay, item);   // global list, whole-list sweeps
        else if (flavour == 2) team_run_item<Cfg<0, false, false, -1, 1>>(S, H.p, B, W, Lay, item);   // run-list scan
        else team_run_item<Cfg<0>>(S, H.p, B, W, Lay, item);                                              // per-thread rounds
    }
    if (flavour != 3 && flavour != 4 && flavour != 5 && flavour != 8) break;
    }
    free(smem);
    return 0;
}

extern "C" long emu_persist_steps(void) { return sqrn::g_emu_persist_steps; }
extern "C" long emu_base_sweeps(void) { return sqrn::g_emu_base_sweeps; }
extern "C" long emu_gl_rebuilds(void) { return sqrn::g_emu_gl_rebuilds; }
extern "C" long emu_gl_catchups(void) { return sqrn::g_emu_gl_catchups; }
extern "C" void emu_gl_set_rebuild(int every) { sqrn::g_emu_gl_rebuild_every = every; }

extern "C" double emu_pyround3(double x) { return pyround3(x); }
