// sqrn_device.cuh -- device side of libsqrn_b200: the greedy stem-selection
// hot path of SQUARNA on sm_100a.
//
// What the reference does per OptimalStems call (SQRNdbnseq.py:792-833) with two
// dense N x N float64 matrices, this code does from per-symbol bit masks held in
// shared memory; no matrix is ever written to HBM:
//
//   pairability (BPMatrix, seq.py:258-339)   = AND of a forward mask of symbol c
//       with the REVERSED mask of c's partners, shifted so that bit i lines up
//       with j = s - i on anti-diagonal s = i + j;
//   stems (AnnotateStems, seq.py:427-495)    = maximal runs of set bits, found
//       with ctz/funnel-shift word tricks, outermost cell first;
//   ScoreStems (seq.py:607-751)              = per-candidate evaluation of the
//       region confined by the innermost pair -- either the reference's own
//       position scan or an equivalent walk over the SELECTED STEMS in 5'
//       order (a few stems instead of hundreds of positions); pow() terms come
//       from host-built tables;
//   ChooseStems (seq.py:754-789)             = team-wide arg-max with the key
//       (score desc, i+j asc, i asc), which is the order the reference's stable
//       sort leaves ties in.
//
// One RESCANNING OptimalStems pass (team_scan) is a three-phase pipeline so that every
// phase keeps the lanes of a warp busy with the same kind of work:
//   1  lane per anti-diagonal:  enumerate runs >= minlen        -> run list
//   2a lane per run:            sum the bp scores (seq.py:416)  -> survivor list
//   2b lane per survivor:       ScoreStems factor product       -> arg-max
// Sequences that run to completion (MODE_TAIL) do not rescan: between two greedy steps the
// masked matrix only loses cells, so the list of maximal runs is built once and then only
// cut where the selected stem touches it -- in shared memory for warp teams (persist_build /
// persist_step), in global memory with cached adjusted scores for CTA teams and thread-block
// clusters (gl_build / gl_step).  DESIGN.md 3.1-3.2.
//
// A "team" is the group of threads that owns one (sequence, partial structure)
// work item: one warp for short sequences, one CTA for long ones.
#pragma once
#include <stdint.h>
#ifndef SQRN_HOST_EMU
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define SQRN_NOINLINE __noinline__
#else
// Single-thread "team" build used only by tests/emu (g++): the same device
// functions with T = 1, so the bit tricks and scoring logic can be debugged
// against the oracle on a box without a GPU.  Never part of libsqrn_b200.so.
#include <math.h>
#include <string.h>
#define __CUDACC__ 1
#define __device__
#define __host__
#define __forceinline__ inline
#define SQRN_NOINLINE
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline uint32_t __brev(uint32_t x) { uint32_t y = 0; for (int b = 0; b < 32; b++) if (x >> b & 1) y |= 1u << (31 - b); return y; }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __double2loint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)u; }
static inline int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x; }
static inline unsigned atomicMax(unsigned *p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
static inline uint32_t atomicAnd(uint32_t *p, uint32_t v) { uint32_t o = *p; *p &= v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
#endif

namespace sqrn {

constexpr int MAXK  = 16;   // symbol codes
constexpr int MAXPC = 8;    // codes that can pair (have masks)
constexpr int CODE_A = 0, CODE_C = 1, CODE_G = 2, CODE_U = 3, CODE_SEP = 4, CODE_OTHER = 5;

constexpr int RC_X = 1, RC_NOLEFT = 2, RC_NORIGHT = 4, RC_RBPOS = 8;

constexpr int MODE_TAIL = 0;   // run the single-path greedy to completion
constexpr int MODE_STEP = 1;   // one OptimalStems call, return the ChooseStems list
constexpr int MODE_YIELD = 2;  // AnnotateStems only, stems in reference order
constexpr int MODE_FINAL = 3;  // ScoreStruct + dbn of the given stems, no selection
constexpr int MODE_BASE = 4;   // build the base list of the sequence (DevWork::base_*)

constexpr int REGION_AUTO = 0, REGION_SCAN = 1, REGION_STEMS = 2;   // ScoreStems region evaluation

// result flags per item
constexpr int FLAG_INT0 = 1;       // struct score is the int 0 (prints "0"), seq.py:871
constexpr int FLAG_LEVELS = 2;     // more than 30 pseudoknot levels: ASCII glyphs exhausted
constexpr int FLAG_CAPACITY = 4;
constexpr int FLAG_ROUND = 8;      // device round(x, 3) was too close to a tie: host redoes it from out_raw2

// ---------------------------------------------------------------- parameters
struct DevParams {
    int      K, npc;
    int      std_pairs;             // the pairing table is exactly {GC, AU, GU} over ACGU
    int      pc_code[MAXPC];        // symbol code of pairing slot c
    uint32_t pairmask[MAXK];        // bit d: code pairs with code d
    double   weight[MAXK * MAXK];
    int      m;                     // prefilter run length = clamp(ceil(minlen), 1, 32)
    double   minlen, minbpscore, minfinscore, loopbonus, maxstemnum, bracketweight;
    int      bw_is_int, bw_int;
    double   distcoef, orderpenalty;
    const double *sdf_lut; int sdf_n;    // (1/(1+k))**distcoef      seq.py:726
    const double *of_lut;  int of_n;     // (1/(1+o))**orderpenalty  seq.py:729
    const double *pw17_lut; int pw17_n;  // (0.5k)**1.7              seq.py:884
    int      ub_ok;                 // the factor maxima below are valid (all pow() terms come from tables)
    double   sdf_max, of_max, lf_max;    // largest stem-distance / order / loop factor (score_bound)
    uint8_t  code_table[256];
};

struct DevBatch {
    int64_t        n_seqs;
    const int64_t *off;
    const uint8_t *sym;
    int            sym_packed;    // sym is a 2-bit stream: base k of the batch at bits 2 (k & 3) of byte k >> 2, A C G U = 0 1 2 3
    const uint16_t *rcode; const double *rf_pos; const double *rf_neg; const double *rvals; int R;
    int            react_comp;
    const uint8_t *rclass; const int64_t *rbp_off; const int32_t *rbp;
    const double  *smat;   int L;  const int32_t *cols;
    int            interchainonly;
    const double  *bpp; const int64_t *bpp_off; int bpp_mode;      // per-sequence N x N term, 1 additive / 2 multiplicative
};

struct BEnt;

struct DevWork {
    int            n_items;
    int            mode;
    int            item_base;    // first item of this launch (chunked launches share one CSR)
    int            region_mode;  // REGION_*
    int            round3;       // TAIL/FINAL: write round(x, 3) scores (ScoreStruct, seq.py:899)
    const int32_t *order;        // processing order (item ids), NULL = identity
    const int32_t *item_seq;     // item -> sequence, NULL = identity
    const int64_t *init_off;     // [n_items+1] CSR of initial stems, NULL = none
    const int32_t *init_stems;   // i, j, len
    const double  *item_subopt;  // MODE_STEP: cursubopt per item
    int           *counter;      // global work counter (zeroed before launch)
    const int     *n_items_dev;  // if set, the item count is read from here (a list an earlier kernel of the stream filled)
    void          *g_ent;        // Cfg::GLIST: persistent candidate lists (16-byte GEnt records), slot b = [b * g_cap, (b + 1) * g_cap)
    double        *g_bps;
    uint8_t       *g_qb;         //   bin of every record (gl_bin of its static bound)
    long long      g_cap;        //   records per HALF slot: a slot is two halves (the rebuild copies from one into the other)
    int            g_rebuild;    //   rebuild period in passes (0: GL_REBUILD)
    int           *g_cnt;        // cluster flavour: GL_CNT_INTS ints per slot -- [0] records in the list, [1] chunk counter of the sweeps, [16..] histograms
    unsigned long long *g_stat;      // optional counters: [0] entries swept, [1] ScoreStems evaluations, [2] cache resets,
                                     // [3] steps, [4] steps with a level change, [5] cuts
    // base lists (pool rounds): records of sequence b at base_ent + base_off[b], capacity base_off[b + 1] - base_off[b];
    // base_n[b] = records (-1: none / overflow), base_bend + 257 b = its bin ends
    BEnt          *base_ent;
    const int64_t *base_off;
    int32_t       *base_n;
    int32_t       *base_bend;
    int           *ovf_count;    // Cfg::PERSIST kernels: items whose run list overflowed are appended here and
    int32_t       *ovf_list;     //   left to a rescanning kernel launched behind them (order = ovf_list)
    // outputs
    const int64_t *out_off;      // [n_items+1] capacity CSR for stems
    int32_t       *out_stems;    // i, j, len
    int32_t       *out_nstems;   // TAIL: total stems; STEP: number of chosen stems
    double        *out_stemfin;  // optional: adjusted score per output stem
    double        *out_raw;      // TAIL: thescore*reactscore, thescore, reactscore per item
    uint8_t       *out_flags;    // FLAG_*
    const int64_t *dbn_off;      // [n_items] offset of the item's dbn
    uint8_t       *out_dbn_ascii;
    int8_t        *out_dbn_code;
    // compact outputs of the packed fast lane (sqrn_fast_predict_packed_host)
    uint8_t       *out_dbn_nib;  // two positions per byte, item b from byte (dbn_off[b] >> 1) + b: 0 '.', L opening / 8 | L closing bracket of level L <= 7
    int32_t       *out_milli;    // 2 per item: round(total, 3) and round(structscore, 3) in units of 0.001
    uint16_t      *out_ns16;     // number of stems
    int           *rnd_count;    // items whose rounding has to be redone by the host (FLAG_ROUND): count, and
    double        *rnd_list;     //   (item, total, structscore, reactscore) unrounded, rnd_cap entries
    int            rnd_cap;
    unsigned long long *n_calls; // OptimalStems-equivalent calls performed
};

// -------------------------------------------------------------- smem layout
struct Layout {
    int Ncap, W, WR, Scap, RBcap, Rcap, Ccap, Ocap, npc, Pcap;
    int o_pbps, o_pkey;
    int o_code, o_rcode, o_rcl, o_partner, o_owner, o_sepcnt, o_M, o_PR, o_rowok, o_colokR, o_Ub, o_Ubase,
        o_sti, o_stj, o_stl, o_stlev, o_evpos, o_evid, o_cc, o_perm, o_grp, o_gsz,
        o_rbv, o_rbw, o_rkey, o_rlen, o_ckey, o_clen, o_cbps, o_cfin, o_red, o_xchg, o_misc, o_stlev2, o_evjump, o_glx, o_stage, total;
};

__host__ __device__ constexpr int align_up(int x, int a) { return (x + a - 1) / a * a; }

// extras = 1: the batch carries reactivities (rcode array needed), 0: it does not, -1: plain batch
// (no restraint classes either); with_fin = the
// mode keeps the adjusted score of every survivor (MODE_STEP)
// scap = most stems one structure can hold (0: the N/2 + 1 upper bound)
// pcap = entries of the PERSISTENT run list (Cfg::PERSIST; 0: none).  That list replaces the run
//        and survivor lists of the rescanning path and shares their space -- a sequence uses one
//        path or the other -- but it is live between scans, so the level scratch gets its own room.
// planes = bit-mask planes per direction (0: npc; the 2-bit-plane flavour needs 2)
__host__ __device__ constexpr Layout make_layout(int Nmax, int RBmax, int Ccap, int npc, int tw, int Rcap = 0,
                                              int extras = 1, int with_fin = 1, int scap = 0, int pcap = 0,
                                              int planes = 0)
{
    Layout L{};
    L.Ncap = align_up(Nmax > 0 ? Nmax : 1, 32);
    L.W = L.Ncap / 32;
    L.WR = L.W + 3;
    L.Scap = L.Ncap / 2 + 1;
    if (scap > 0 && scap < L.Scap) L.Scap = scap;
    L.RBcap = RBmax;
    L.Ccap = Ccap;
    L.Rcap = Rcap > 0 ? Rcap : Ccap;
    L.Pcap = pcap;
    L.npc = npc;
    if (planes <= 0) planes = npc;
    int o = 0;
    // ---- the lists (doubles first: 8-byte aligned)
    // (global-list kernels, pcap < 0: cbps only holds one int per thread -- the record that was its best -- and ckey the
    //  per-warp screen / evaluate lists; clen and the staging ring of the base-list sweep are not used)
    const bool gl = pcap < 0;
    L.o_cbps = o;    o += gl ? align_up(4 * 32 * (tw > 0 ? tw : 1), 8) : 8 * Ccap;
    L.o_ckey = o;    o += 4 * Ccap;
    // run list; without a persistent list the level scratch (cc, gsz, perm, grp) is only live between
    // two scans and shares its space
    int lev_bytes = 4 * L.Scap + 4 * L.Scap + 2 * L.Scap + 2 * L.Scap;
    int run_bytes = 4 * L.Rcap + 2 * L.Rcap;
    int uni = align_up((lev_bytes > run_bytes && pcap <= 0) ? lev_bytes : run_bytes, 4);
    L.Ocap = uni / 2;                                     // int16 entries team_choose can rank in this space
    L.o_rkey = o;    L.o_rlen = o + 4 * L.Rcap;
    L.o_cc = o;      L.o_gsz = o + 4 * L.Scap;  L.o_perm = o + 8 * L.Scap;  L.o_grp = o + 10 * L.Scap;
    o += uni;
    L.o_clen = o;    o += gl ? 0 : 2 * Ccap;
    L.o_pbps = 0;    L.o_pkey = pcap > 0 ? 8 * pcap : 0;
    if (pcap > 0) {
        if (o < 12 * pcap) o = 12 * pcap;
        o = align_up(o, 4);
        L.o_cc = o;  L.o_gsz = o + 4 * L.Scap;  L.o_perm = o + 8 * L.Scap;  L.o_grp = o + 10 * L.Scap;
        o += align_up(lev_bytes, 4);
    }
    o = align_up(o, 8);
    L.o_cfin = o;    o += with_fin ? 8 * Ccap : 0;
    L.o_red = o;     o += (tw > 1) ? 32 * tw : 0;        // cross-warp reduction scratch
    L.o_xchg = o;    o += (tw > 1) ? 2 * 16 * 16 : 0;     // cluster exchange: 2 parities x 16 ranks x (fin, key, len)
    L.o_misc = o;    o += 64;
    L.o_M = o;       o += 4 * planes * L.W;
    L.o_PR = o;      o += 4 * planes * L.WR;
    L.o_rowok = o;   o += 4 * L.W;
    L.o_colokR = o;  o += 4 * L.WR;
    L.o_Ub = o;      o += 4 * L.W;
    L.o_Ubase = o;   o += 4 * (L.W + 1);
    L.o_partner = o; o += 2 * L.Ncap;
    L.o_owner = o;   o += 2 * L.Ncap;
    L.o_sepcnt = o;  o += 2 * (L.Ncap + 2);
    L.o_sti = o;     o += 2 * L.Scap;
    L.o_stj = o;     o += 2 * L.Scap;
    L.o_stl = o;     o += 2 * L.Scap;
    L.o_evpos = o;   o += 4 * L.Scap;                     // two arm events per stem
    L.o_evid = o;    o += 4 * L.Scap;
    L.o_rbv = o;     o += 2 * (RBmax + 1);
    L.o_rbw = o;     o += 2 * (RBmax + 1);
    L.o_rcode = o;   o += extras > 0 ? 2 * L.Ncap : 0;
    L.o_code = o;    o += L.Ncap;
    L.o_rcl = o;     o += extras >= 0 ? L.Ncap : 0;
    L.o_stlev = o;   o += L.Scap;
    L.o_stlev2 = o;  o += pcap < 0 ? L.Scap : 0;      // pcap < 0: global persistent list (levels of the previous step)
    o = align_up(o, 2);
    L.o_evjump = tw != 1 ? o : -1;                    // CTA teams (and the host emulation, tw = 0): where the arm walk of
    o += tw != 1 ? 4 * L.Scap : 0;                    // ScoreStems may jump to (team_apply_stem)
    o = align_up(o, 4);
    L.o_glx = o;     o += (pcap < 0 || tw != 1) ? 4 * (3 * 256 + 4) : 0;   // binned lists: histogram, bin ends, scatter cursors (gl_make_bins)
    o = align_up(o, 16);
    L.o_stage = o;   o += (tw > 1 && !gl) ? 16 + 2 * 16 * 32 * tw : 0;             // base-list sweep: two mbarriers + two tiles of T 16-byte records (cp.async.bulk)
    L.total = align_up(o, 16);
    return L;
}

#ifdef __CUDACC__

struct Best {
    double   fin;
    uint32_t key;     // (i+j) << 16 | i : smaller = earlier in the reference's enumeration order
    int      len;
};

__device__ __forceinline__ bool better(double fa, uint32_t ka, double fb, uint32_t kb)
{
    return fa > fb || (fa == fb && ka < kb);
}

// ------------------------------------------------------------------- team
// TW == 1: one warp per work item, several teams per CTA.
// TW  > 1: the whole CTA (TW warps) is one team.
// TW == 0: a single thread (host emulation build only).
template <int TW> struct Team {
    static constexpr int T = TW * 32;
#ifndef SQRN_HOST_EMU
    __device__ static __forceinline__ int rank() { return TW == 1 ? (threadIdx.x & 31) : threadIdx.x; }
    __device__ static __forceinline__ void sync()
    {
        if (TW == 1) __syncwarp(); else __syncthreads();
    }
    // both are barriers (memory included) for the team
    __device__ static __forceinline__ bool any(bool p)
    {
        if (TW == 1) { bool a = __any_sync(0xffffffffu, p); __syncwarp(); return a; }
        return __syncthreads_or(p) != 0;
    }
    // number of threads with p set (barrier)
    __device__ static __forceinline__ int count(bool p)
    {
        if (TW == 1) { int c = __popc(__ballot_sync(0xffffffffu, p)); __syncwarp(); return c; }
        return __syncthreads_count(p);
    }
    // Threads with p set claim consecutive slots starting at `base` (uniform); returns how many did.
    // TW > 1 uses the team-shared `counter`, which the caller keeps equal to `base` between calls.
    __device__ static __forceinline__ int claim(bool p, int *counter, int base, int &slot)
    {
        if (TW == 1) {
            uint32_t b = __ballot_sync(0xffffffffu, p);
            slot = base + __popc(b & ((1u << (threadIdx.x & 31)) - 1u));
            __syncwarp();
            return __popc(b);
        }
        slot = p ? atomicAdd(counter, 1) : 0;
        return __syncthreads_count(p);
    }
#endif
};
#ifdef SQRN_HOST_EMU
template <> struct Team<0> {
    static constexpr int T = 1;
    static inline int rank() { return 0; }
    static inline void sync() {}
    static inline bool any(bool p) { return p; }
    static inline int count(bool p) { return p ? 1 : 0; }
    static inline int claim(bool p, int *, int base, int &slot) { slot = base; return p ? 1 : 0; }
};
#endif

// Compile-time flavour of the work kernel.
//   PLAIN: the batch has no reactivities, restraints, alignment weights or interchainonly flag
//          (the `byseq` fast lane): all of that code is compiled out;
//   STDP:  the pairing table is exactly {GC, AU, GU} over ACGU: pairability comes from the two
//          bit planes of the 2-bit base codes, x = (b0 ^ r0) & (b1 | r1);
//   MODE:  a fixed MODE_* or -1 (taken from DevWork at run time).
//   RUNLIST: phase 1 collects runs in a shared list before they are scored (warp teams: keeps
//          the lanes of phase 2a dense); off, every thread scores the runs of its own diagonal
//          in team-wide rounds (CTA teams).  -1: on for warp teams only.
//   CLUSTER: the team is a thread-block CLUSTER: every CTA keeps a full replica of the (small)
//          sequence state in its own shared memory and scans every CS-th anti-diagonal; the
//          CTA-local winners meet in rank 0's shared memory through DSMEM once per greedy step.
//   PERSIST: MODE_TAIL keeps the list of maximal runs ACROSS greedy steps (persist_build /
//          persist_step): the anti-diagonals are enumerated once, afterwards only the runs that
//          touch the stem just selected are cut into their surviving pieces.  Needs Layout::Pcap.
//          A PERSIST kernel carries no rescanning code at all (its hot instruction range has to
//          fit the SM's instruction cache): an item whose list overflows is appended to
//          DevWork::ovf_list and redone from scratch by a non-PERSIST kernel.
//   GLIST: the persistent list lives in GLOBAL memory (DevWork::g_*; one slot per resident CTA) and
//          caches the adjusted score of every candidate between steps (gl_build / gl_step): for CTA
//          teams (long sequences, 10^5 .. 10^6 runs).  -1: on for CTA teams.
//   IO:    boundary format of the fast lane: 1 bytes only (ASCII symbols in, ASCII dot-bracket + float64 scores out),
//          2 packed only (2-bit codes in, 4-bit bracket codes + thousandths out), 0 decided at run time (DevBatch::
//          sym_packed, DevWork::out_*).  The fast kernels carry one format each: their hot code has to stay small.
template <int TW_, bool PLAIN_ = false, bool STDP_ = false, int MODE_ = -1, int RUNLIST_ = -1, bool CLUSTER_ = false,
          bool PERSIST_ = false, int GLIST_ = -1, int IO_ = 0>
struct Cfg {
    static constexpr int TW = TW_, MODE = MODE_, IO = IO_;
    static constexpr bool PLAIN = PLAIN_, STDP = STDP_, CLUSTER = CLUSTER_, PERSIST = PERSIST_;
    static constexpr bool GLIST = PERSIST_ && (GLIST_ < 0 ? (TW_ > 1) : (GLIST_ != 0));
    // GLIST_ == 2: every pass sweeps the whole list (no prefix rounds, no catch-up, no rebuild): for lists of a few thousand
    // records, where that machinery buys nothing and its code costs instruction-cache hits (four CTAs per SM)
    static constexpr bool GL_SIMPLE = GLIST_ == 2;
    static constexpr bool RUNLIST = RUNLIST_ < 0 ? (TW_ == 1) : (RUNLIST_ != 0);
};

struct State {
    int N, W, WR, nst, nrb, has_sep, has_react, has_smat, default_reacts, region_mode;
    int bpp_mode; const double *bpp;   // base-pair-probability term of this sequence (seq.py:341-365), 0 / NULL: none
    int dstride, doffset;            // this team scans the anti-diagonals 4 + doffset + dstride * q (cluster: rank, size)
    uint8_t  *code, *rcl, *stlev, *stlev2;
    uint16_t *rcode;
    int16_t  *partner, *owner, *sepcnt, *sti, *stj, *stl, *evpos, *evid, *evjump, *perm, *grp, *rbv, *rbw;
    uint32_t *M, *PR, *rowok, *colokR, *Ub, *ckey, *rkey, *pkey;
    int32_t  *cc, *gsz, *Ubase, *ghist, *gbend, *gcur;
    uint16_t *clen, *rlen;
    double   *cbps, *cfin, *red, *pbps;
    unsigned char *xchg, *stage;
    uint32_t bphase[2];  // phase parity of the two staging barriers (base-list sweep)
    int      *misc;      // [0] run count  [1] next item  [2..6] scratch  [7] survivor count  [8] persistent entries
    const int32_t *cols; // global, per sequence
};

__device__ __forceinline__ State bind_state(unsigned char *base, const Layout &L)
{
    State s;
    s.code = base + L.o_code;  s.rcode = (uint16_t *)(base + L.o_rcode);  s.rcl = base + L.o_rcl;
    s.stlev = base + L.o_stlev;  s.stlev2 = base + L.o_stlev2;
    s.partner = (int16_t *)(base + L.o_partner);  s.owner = (int16_t *)(base + L.o_owner);
    s.sepcnt = (int16_t *)(base + L.o_sepcnt);
    s.sti = (int16_t *)(base + L.o_sti);  s.stj = (int16_t *)(base + L.o_stj);  s.stl = (int16_t *)(base + L.o_stl);
    s.evpos = (int16_t *)(base + L.o_evpos);  s.evid = (int16_t *)(base + L.o_evid);
    s.evjump = L.o_evjump >= 0 ? (int16_t *)(base + L.o_evjump) : nullptr;
    s.perm = (int16_t *)(base + L.o_perm);  s.grp = (int16_t *)(base + L.o_grp);
    s.rbv = (int16_t *)(base + L.o_rbv);  s.rbw = (int16_t *)(base + L.o_rbw);
    s.M = (uint32_t *)(base + L.o_M);  s.PR = (uint32_t *)(base + L.o_PR);
    s.rowok = (uint32_t *)(base + L.o_rowok);  s.colokR = (uint32_t *)(base + L.o_colokR);
    s.Ub = (uint32_t *)(base + L.o_Ub);  s.Ubase = (int32_t *)(base + L.o_Ubase);
    s.ckey = (uint32_t *)(base + L.o_ckey);  s.rkey = (uint32_t *)(base + L.o_rkey);
    s.cc = (int32_t *)(base + L.o_cc);  s.gsz = (int32_t *)(base + L.o_gsz);
    s.clen = (uint16_t *)(base + L.o_clen);  s.rlen = (uint16_t *)(base + L.o_rlen);
    s.cbps = (double *)(base + L.o_cbps);
    s.pbps = (double *)(base + L.o_pbps);  s.pkey = (uint32_t *)(base + L.o_pkey);
    s.cfin = (double *)(base + L.o_cfin);
    s.red = (double *)(base + L.o_red);
    s.xchg = base + L.o_xchg;
    s.misc = (int *)(base + L.o_misc);
    s.ghist = (int32_t *)(base + L.o_glx);  s.gbend = s.ghist + 256;  s.gcur = s.gbend + 260;
    s.stage = base + L.o_stage;
    s.W = L.W; s.WR = L.WR;
    s.N = 0; s.nst = 0; s.nrb = 0; s.has_sep = 0; s.has_react = 0; s.has_smat = 0; s.default_reacts = 1;
    s.region_mode = REGION_AUTO;
    s.bpp_mode = 0; s.bpp = nullptr;
    s.dstride = 1; s.doffset = 0;
    s.cols = nullptr;
    s.bphase[0] = s.bphase[1] = 0;
    return s;
}

// bit b of dst[k] = pred(32 k + b): one ballot per word
template <int TW, class F>
__device__ __forceinline__ void team_mask(uint32_t *dst, int nwords, F &&pred)
{
#ifdef SQRN_HOST_EMU
    for (int k = 0; k < nwords; k++) {
        uint32_t m = 0;
        #pragma unroll 1
        for (int b = 0; b < 32; b++) if (pred(32 * k + b)) m |= 1u << b;
        dst[k] = m;
    }
#else
    const int lane = threadIdx.x & 31;
    const int w0 = (TW == 1) ? 0 : (int)(threadIdx.x >> 5), nw = (TW == 1) ? 1 : TW;
    #pragma unroll 1
    for (int k = w0; k < nwords; k += nw) {
        uint32_t m = __ballot_sync(0xffffffffu, pred(32 * k + lane));
        if (lane == 0) dst[k] = m;
    }
#endif
}

// exclusive prefix sum over the team; total returned through `total`
template <int TW>
__device__ __forceinline__ int team_exscan(State &S, int v, int &total)
{
#ifdef SQRN_HOST_EMU
    total = v; return 0;
#else
    int lane = threadIdx.x & 31, x = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    int wtot = __shfl_sync(0xffffffffu, x, 31);
    int excl = x - v;
    if (TW == 1) { total = wtot; return excl; }
    int *sc = (int *)S.red;
    int w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 31) sc[w] = wtot;
    __syncthreads();
    int base = 0, tot = 0;
    #pragma unroll 1
    for (int q = 0; q < TW; q++) { int t = sc[q]; if (q < w) base += t; tot += t; }
    __syncthreads();
    total = tot;
    return base + excl;
#endif
}

// minimum hairpin rule, seq.py:293-297: j >= i + inc4(i)
__device__ __forceinline__ int inc4_of(const State &S, int i)
{
    int inc = 4;
    if (i + 1 < S.N && S.code[i + 1] == CODE_SEP) inc = 2;
    if (i + 2 < S.N && S.code[i + 2] == CODE_SEP) inc = 3;
    return inc;
}

// number of unpaired positions in [0, p)
__device__ __forceinline__ int unpaired_before(const State &S, int p)
{
    return S.Ubase[p >> 5] + __popc(S.Ub[p >> 5] & ((1u << (p & 31)) - 1u));
}

// bits p0 .. p0 + 4 of the unpaired mask (positions outside the sequence read as 0)
__device__ __forceinline__ uint32_t unpaired_bits5(const State &S, int p0)
{
    const int w = p0 >> 5, sh = p0 & 31;
    const uint32_t lo = (w >= 0 && w < S.W) ? S.Ub[w] : 0u, hi = (w + 1 >= 0 && w + 1 < S.W) ? S.Ub[w + 1] : 0u;
    return __funnelshift_r(lo, hi, sh) & 31u;
}

// word prefix of the unpaired mask (after Ub changed)
template <class C>
__device__ void team_unpaired_prefix(State &S)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    int carry = 0;
    #pragma unroll 1
    for (int k0 = 0; k0 < S.W; k0 += T) {
        int k = k0 + r, v = (k < S.W) ? __popc(S.Ub[k]) : 0, tot;
        int ex = team_exscan<TW>(S, v, tot);
        if (k < S.W) S.Ubase[k] = carry + ex;
        carry += tot;
    }
    if (r == 0) S.Ubase[S.W] = carry;
    Team<TW>::sync();
}

// ------------------------------------------------------------- load a sequence
template <class C>
__device__ void team_load(State &S, const DevBatch &B, const DevParams &P, int seq)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    const int64_t o = B.off[seq];
    const int N = (int)(B.off[seq + 1] - o);
    S.N = N; S.nst = 0;
#ifndef SQRN_HOST_EMU
    if (C::PLAIN && C::STDP && TW == 1) {
        // fast lane: the forward masks come from ballots over the symbols as they are read (one
        // coalesced pass), the reversed ones from the forward words by a funnel shift and a bit reversal
        S.has_react = 0; S.has_smat = 0; S.cols = nullptr; S.default_reacts = 1; S.nrb = 0;
        S.bpp_mode = 0; S.bpp = nullptr;
        bool sep = false;
        #pragma unroll 1
        for (int k = 0; k < S.W; k++) {
            const int p = 32 * k + r;
            uint8_t c = CODE_OTHER;
            if (p < N) c = (C::IO == 2 || (C::IO == 0 && B.sym_packed)) ? (uint8_t)((B.sym[(o + p) >> 2] >> (2 * (int)((o + p) & 3))) & 3)
                                                                         : P.code_table[B.sym[o + p]];
            S.code[p] = c;
            S.partner[p] = -1;
            sep |= c == CODE_SEP;
            const uint32_t in = __ballot_sync(0xffffffffu, p < N);
            const uint32_t b0 = __ballot_sync(0xffffffffu, p < N && (c & 1));
            const uint32_t b1 = __ballot_sync(0xffffffffu, p < N && (c & 2));
            const uint32_t okb = __ballot_sync(0xffffffffu, p < N && c < 4);
            if (r == 0) {
                S.M[k] = b0; S.M[S.W + k] = b1; S.rowok[k] = okb; S.Ub[k] = in;
                S.Ubase[k] = (32 * k < N) ? 32 * k : N;
            }
        }
        if (r == 0) S.Ubase[S.W] = N;
        S.has_sep = __any_sync(0xffffffffu, sep);
        __syncwarp();
        // bit b of reversed word q is position N + 31 - 32 q - b: the forward window starting at N - 32 q, reversed
        #pragma unroll 1
        for (int q = r; q < S.WR; q += 32) {
            const int p0 = N - 32 * q, w = p0 >> 5, sh = p0 & 31;
            const bool lo_ok = w >= 0 && w < S.W, hi_ok = w + 1 >= 0 && w + 1 < S.W;
            S.PR[q] = __brev(__funnelshift_r(lo_ok ? S.M[w] : 0u, hi_ok ? S.M[w + 1] : 0u, sh));
            S.PR[S.WR + q] = __brev(__funnelshift_r(lo_ok ? S.M[S.W + w] : 0u, hi_ok ? S.M[S.W + w + 1] : 0u, sh));
            S.colokR[q] = __brev(__funnelshift_r(lo_ok ? S.rowok[w] : 0u, hi_ok ? S.rowok[w + 1] : 0u, sh));
        }
        if (S.has_sep && r == 0) {
            int c = 0;
            #pragma unroll 1
            for (int p = 0; p < N; p++) { S.sepcnt[p] = (int16_t)c; if (S.code[p] == CODE_SEP) c++; }
            S.sepcnt[N] = (int16_t)c;
        }
        __syncwarp();
        return;
    }
#endif
    S.has_react = !C::PLAIN && B.rcode != nullptr;
    S.has_smat = !C::PLAIN && B.smat != nullptr;
    S.cols = (!C::PLAIN && B.cols) ? B.cols + o : nullptr;
    S.bpp_mode = (!C::PLAIN && B.bpp) ? B.bpp_mode : 0;
    S.bpp = S.bpp_mode ? B.bpp + B.bpp_off[seq] : nullptr;
    const int Nw = S.W * 32;
    bool sep = false, nondef = false;
    #pragma unroll 1
    for (int p = r; p < Nw; p += T) {
        uint8_t c = CODE_OTHER, cl = 0;
        if (p < N) {
            c = (C::IO == 2 || (C::IO == 0 && B.sym_packed)) ? (uint8_t)((B.sym[(o + p) >> 2] >> (2 * (int)((o + p) & 3))) & 3)
                                                             : P.code_table[B.sym[o + p]];
            if (!C::PLAIN) {
                if (B.rcode) { uint16_t rc = B.rcode[o + p]; S.rcode[p] = rc; if (__ldg(&B.rvals[rc]) != 0.5) nondef = true; }
                if (B.rclass) cl = B.rclass[o + p] & 7;
            }
            if (c == CODE_SEP) sep = true;
        } else if (!C::PLAIN && B.rcode) S.rcode[p] = 0;
        S.code[p] = c;
        if (!C::PLAIN) S.rcl[p] = cl;
        S.partner[p] = -1;
        if (!C::PLAIN) S.owner[p] = -1;
    }
    S.has_sep = Team<TW>::any(sep);
    // "default reacts" switch, seq.py:273: every processed reactivity == 0.5
    S.default_reacts = C::PLAIN ? 1 : !Team<TW>::any(nondef);
    S.nrb = 0;
    int nrb_all = 0;
    const int32_t *rb = nullptr;
    if (!C::PLAIN) {
        // restraint pairs (sorted by (v+w, v) by the host); mark their positions
        if (B.rbp_off) { nrb_all = (int)(B.rbp_off[seq + 1] - B.rbp_off[seq]); rb = B.rbp + 2 * B.rbp_off[seq]; }
        #pragma unroll 1
        for (int k = r; k < nrb_all; k += T) {
            S.rcl[rb[2 * k]] |= RC_RBPOS;       // distinct positions: no write conflicts on the same byte
            S.rcl[rb[2 * k + 1]] |= RC_RBPOS;
        }
    }
    if (S.has_sep || nrb_all) {
        if (r == 0) {
            // prefix count of separators (rare: only multi-chain inputs)
            if (S.has_sep) {
                int c = 0;
                #pragma unroll 1
                for (int p = 0; p < N; p++) { S.sepcnt[p] = (int16_t)c; if (S.code[p] == CODE_SEP) c++; }
                S.sepcnt[N] = (int16_t)c;
            }
            // statically valid restraint cells: boolmat[v,w] != 0 (seq.py:443) and on a walked diagonal
            int n = 0;
            #pragma unroll 1
            for (int k = 0; k < nrb_all; k++) {
                int v = rb[2 * k], w = rb[2 * k + 1], s = v + w;
                if (s < 4 || s > 2 * N - 6) continue;
                if (!(P.pairmask[S.code[v]] >> S.code[w] & 1)) continue;
                int inc = 4;
                if (v + 1 < N && S.code[v + 1] == CODE_SEP) inc = 2;
                if (v + 2 < N && S.code[v + 2] == CODE_SEP) inc = 3;
                if (w < v + inc) continue;
                if (B.interchainonly && (!S.has_sep || S.sepcnt[w] - S.sepcnt[v] <= 0)) continue;
                S.rbv[n] = (int16_t)v; S.rbw[n] = (int16_t)w; n++;
            }
            S.misc[3] = n;
        }
        Team<TW>::sync();
        S.nrb = S.misc[3];
    }
    Team<TW>::sync();
    // bit q of the reversed arrays is position j = N - 1 - (q - 32)
    if (C::STDP) {
        // the two bit planes of the 2-bit base codes, forward and reversed; positions that hold
        // no base (separators, other symbols) are excluded through rowok / colokR
        team_mask<TW>(S.M, S.W, [&](int p) { return (S.code[p] & 1) != 0; });
        team_mask<TW>(S.M + S.W, S.W, [&](int p) { return (S.code[p] & 2) != 0; });
        team_mask<TW>(S.PR, S.WR, [&](int q) { int j = N - 1 - (q - 32); return j >= 0 && j < N && (S.code[j] & 1); });
        team_mask<TW>(S.PR + S.WR, S.WR, [&](int q) { int j = N - 1 - (q - 32); return j >= 0 && j < N && (S.code[j] & 2); });
    } else {
        // forward masks of each pairing symbol, reversed masks of its partners
        #pragma unroll 1
        for (int c = 0; c < P.npc; c++) {
            const int code = P.pc_code[c];
            const uint32_t pm = P.pairmask[code];
            team_mask<TW>(S.M + c * S.W, S.W, [&](int p) { return p < N && S.code[p] == code; });
            team_mask<TW>(S.PR + c * S.WR, S.WR, [&](int q) { int j = N - 1 - (q - 32); return j >= 0 && j < N && (pm >> S.code[j] & 1); });
        }
    }
    const int cmax = C::STDP ? 4 : MAXK;          // codes that hold a base
    if (C::PLAIN) {
        team_mask<TW>(S.rowok, S.W, [&](int p) { return p < N && S.code[p] < cmax; });
        team_mask<TW>(S.colokR, S.WR, [&](int q) { int j = N - 1 - (q - 32); return j >= 0 && j < N && S.code[j] < cmax; });
    } else {
        team_mask<TW>(S.rowok, S.W, [&](int p) { return p < N && S.code[p] < cmax && !(S.rcl[p] & (RC_X | RC_NORIGHT | RC_RBPOS)); });
        team_mask<TW>(S.colokR, S.WR, [&](int q) { int j = N - 1 - (q - 32); return j >= 0 && j < N && S.code[j] < cmax && !(S.rcl[j] & (RC_X | RC_NOLEFT | RC_RBPOS)); });
    }
    team_mask<TW>(S.Ub, S.W, [&](int p) { return p < N; });
    #pragma unroll 1
    for (int k = r; k <= S.W; k += T) S.Ubase[k] = (32 * k < N) ? 32 * k : N;
    Team<TW>::sync();
}

// add a selected stem to the structure: partners, owner, row/column masks
// (AnnotateStems zeroes the rows and columns of every selected position, seq.py:446-451),
// the unpaired mask and the 5'-sorted copy of the stem list
// light: the structure is only going to be scored and printed (MODE_FINAL): no arm events, nothing ScoreStems walks
template <class C>
__device__ void team_apply_stem(State &S, int i, int j, int len, bool refresh = true, bool light = false)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    const int idx = S.nst;
    #pragma unroll 1
    for (int k = r; k < len; k += T) {
        int v = i + k, w = j - k;
        S.partner[v] = (int16_t)w; S.partner[w] = (int16_t)v;
        S.owner[v] = (int16_t)idx; S.owner[w] = (int16_t)idx;
        atomicAnd(&S.rowok[v >> 5], ~(1u << (v & 31)));
        atomicAnd(&S.rowok[w >> 5], ~(1u << (w & 31)));
        atomicAnd(&S.Ub[v >> 5], ~(1u << (v & 31)));
        atomicAnd(&S.Ub[w >> 5], ~(1u << (w & 31)));
        int rv = 32 + S.N - 1 - v, rw = 32 + S.N - 1 - w;
        atomicAnd(&S.colokR[rv >> 5], ~(1u << (rv & 31)));
        atomicAnd(&S.colokR[rw >> 5], ~(1u << (rw & 31)));
    }
    if (light) {
        if (r == 0) { S.sti[idx] = (int16_t)i; S.stj[idx] = (int16_t)j; S.stl[idx] = (int16_t)len; }
        S.nst = idx + 1;
        return;                                  // (the caller synchronises once after the last stem)
    }
    // two arm events (first position of the 5' arm and of the 3' arm) go into the position-sorted
    // event list that ScoreStems walks: entries behind them move up by one or two
    const int ne = 2 * idx, p1 = i, p2 = j - len + 1;
    int r1, r2;
    {
        int lo_ = 0, hi_ = ne;
        #pragma unroll 1
        while (lo_ < hi_) { int mid = (lo_ + hi_) >> 1; if (S.evpos[mid] < p1) lo_ = mid + 1; else hi_ = mid; }
        r1 = lo_; hi_ = ne;
        #pragma unroll 1
        while (lo_ < hi_) { int mid = (lo_ + hi_) >> 1; if (S.evpos[mid] < p2) lo_ = mid + 1; else hi_ = mid; }
        r2 = lo_;
    }
    Team<TW>::sync();
    #pragma unroll 1
    for (int q0 = 0; q0 < ne; q0 += T) {             // from the top, so reads precede overwrites
        const int q = ne - 1 - q0 - r;
        int16_t vp = 0, vi = 0; const bool mv = q >= r1;
        if (mv) { vp = S.evpos[q]; vi = S.evid[q]; }
        Team<TW>::sync();
        if (mv) { int d = q >= r2 ? 2 : 1; S.evpos[q + d] = vp; S.evid[q + d] = vi; }
        Team<TW>::sync();
    }
    if (r == 0) {
        S.sti[idx] = (int16_t)i; S.stj[idx] = (int16_t)j; S.stl[idx] = (int16_t)len;
        S.evpos[r1] = (int16_t)p1; S.evid[r1] = (int16_t)(2 * idx);
        S.evpos[r2 + 1] = (int16_t)p2; S.evid[r2 + 1] = (int16_t)(2 * idx + 1);
    }
    S.nst = idx + 1;
    Team<TW>::sync();
    if (TW != 1 && S.evjump) {
        // A selected stem t = (i, j) is CLOSED when no selected stem starts inside (i, j) and ends beyond j.
        // The arm walk of ScoreStems (region_stems) that meets the 5' arm of a closed stem lying inside the
        // candidate's region may jump to the first arm behind j: everything in between is under the block
        // the stem opens (`inblockend` >= j, seq.py:672-689) and can neither count nor extend it.
        const int ne2 = 2 * S.nst;
        #pragma unroll 1
        for (int q = r; q < ne2; q += T) {
            const int id = S.evid[q];
            int jmp = q + 1;
            if (!(id & 1)) {
                const int t = id >> 1, ti = S.sti[t], tj = S.stj[t];
                bool closed = true;
                #pragma unroll 1
                for (int u = 0; u < S.nst && closed; u++) { const int ui_ = S.sti[u]; if (ui_ > ti && ui_ < tj && S.stj[u] > tj) closed = false; }
                if (closed) {
                    int lo_ = q + 1, hi_ = ne2;
                    #pragma unroll 1
                    while (lo_ < hi_) { int mid = (lo_ + hi_) >> 1; if (S.evpos[mid] <= tj) lo_ = mid + 1; else hi_ = mid; }
                    jmp = lo_;
                }
            }
            S.evjump[q] = (int16_t)jmp;
        }
        Team<TW>::sync();
    }
    if (refresh) team_unpaired_prefix<C>(S);
}

// ------------------------------------------------ pseudoknot levels per stem
// PairsToDBN(returnlevels=True), seq.py:119-150, restated per STEM: two
// position-disjoint stems cross all-or-none, so the crossing test runs on the
// outermost pairs, cross_count is the summed length of the crossing stems, the
// first-fit order is (cross_count, outer i) and groups are ranked by their
// number of pairs (stable).  tests/test_emu_vs_oracle.py checks this against the
// oracle's per-pair restatement.
__device__ __forceinline__ bool stems_cross(int i, int j, int k, int l)
{
    return (i < k && k < j && j < l) || (k < i && i < l && l < j);
}

// added >= 0: stlev[] is current except for the stem with this index, which was just appended.  A stem
// that crosses nothing joins the first-created group whatever the others do (its cross_count is 0 and
// it cannot clash with anyone), the cross_counts and the first-fit of the others do not change, and the
// first group only grows -- so if that group already ranks first (misc[13]), every level stays and
// the new stem is on level 1.  Anything else is recomputed in full.
template <class C>
__device__ int team_levels(State &S, int added = -1)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    const int n = S.nst;
    if (n == 0) { if (r == 0) S.misc[13] = 1; Team<TW>::sync(); return 0; }
    if (added >= 0 && S.misc[13] == 1) {
        const int i = S.sti[added], j = S.stj[added];
        bool x = false;
        #pragma unroll 1
        for (int u = r; u < n; u += T) if (u != added && stems_cross(i, j, S.sti[u], S.stj[u])) x = true;
        if (!Team<TW>::any(x)) {
            if (r == 0) S.stlev[added] = 1;
            Team<TW>::sync();
            return 1;
        }
    }
    bool crossing = false;
    #pragma unroll 1
    for (int t = r; t < n; t += T) {
        int i = S.sti[t], j = S.stj[t], c = 0;
        #pragma unroll 1
        for (int u = 0; u < n; u++)
            if (u != t && stems_cross(i, j, S.sti[u], S.stj[u])) c += S.stl[u];
        S.cc[t] = c;
        if (c) crossing = true;
    }
    crossing = Team<TW>::any(crossing);
    int ng = 1;
    if (!crossing) {
        // nothing crosses: one group, every stem on level 1
        #pragma unroll 1
        for (int t = r; t < n; t += T) S.stlev[t] = 1;
        if (r == 0) S.misc[13] = 1;
        Team<TW>::sync();
    } else {
        Team<TW>::sync();
        #pragma unroll 1
        for (int t = r; t < n; t += T) {
            int c = S.cc[t], i = S.sti[t], rank = 0;
            #pragma unroll 1
            for (int u = 0; u < n; u++) {
                int cu = S.cc[u];
                if (cu < c || (cu == c && S.sti[u] < i)) rank++;
            }
            S.perm[rank] = (int16_t)t;
        }
        Team<TW>::sync();
        bool serial = true;
#ifndef SQRN_HOST_EMU
        if (TW > 1) {
            // CTA teams (hundreds of stems): perm[] is sorted by cross_count, so the stems that cross nothing come
            // first and all join group 0; the others are placed one after the other (the order matters) by warp 0,
            // a lane per earlier crossing stem: bit h of the clash mask = group h holds a stem that crosses this
            // one, first-fit = its lowest clear bit.  More than 64 groups: the serial loop below.
            if (r < 32) {
                const int lane = r;
                int n0 = 0, sz0 = 0;
                #pragma unroll 1
                for (int a0 = 0; a0 < n; a0 += 32) {
                    const int a = a0 + lane;
                    bool z = false; int l = 0;
                    if (a < n) { const int t = S.perm[a]; z = S.cc[t] == 0; if (z) { S.grp[t] = 0; l = S.stl[t]; } }
                    const int nz = __popc(__ballot_sync(0xffffffffu, z));
                    n0 += nz;
                    sz0 += __reduce_add_sync(0xffffffffu, l);
                    if (nz < 32) break;
                }
                int ngw = 0, over = 0;
                if (n0 > 0) { ngw = 1; if (lane == 0) S.gsz[0] = sz0; }
                __syncwarp();
                #pragma unroll 1
                for (int a = n0; a < n; a++) {
                    const int t = S.perm[a], i = S.sti[t], j = S.stj[t];
                    uint32_t mlo = 0, mhi = 0;
                    #pragma unroll 1
                    for (int b = n0 + lane; b < a; b += 32) {
                        const int u = S.perm[b];
                        if (stems_cross(i, j, S.sti[u], S.stj[u])) { const int h = S.grp[u]; if (h < 32) mlo |= 1u << h; else mhi |= 1u << (h - 32); }
                    }
                    mlo = __reduce_or_sync(0xffffffffu, mlo); mhi = __reduce_or_sync(0xffffffffu, mhi);
                    const int g = ~mlo ? __ffs(~mlo) - 1 : (~mhi ? 32 + __ffs(~mhi) - 1 : 64);
                    if (g >= 64) { over = 1; break; }
                    if (lane == 0) { S.grp[t] = (int16_t)g; S.gsz[g] = (g == ngw ? 0 : S.gsz[g]) + S.stl[t]; }
                    if (g == ngw) ngw++;
                    __syncwarp();
                }
                // does the first-created group (home of every stem that crosses nothing) rank first?
                bool bigger = false;
                #pragma unroll 1
                for (int h = 1 + lane; h < ngw; h += 32) if (S.gsz[h] > S.gsz[0]) bigger = true;
                bigger = __any_sync(0xffffffffu, bigger);
                if (lane == 0) { S.misc[5] = ngw; S.misc[13] = bigger ? 0 : 1; S.misc[6] = over; }
            }
            __syncthreads();
            serial = S.misc[6] != 0;
            if (!serial) {
                // groups.sort(key=len, reverse=True) is stable: level = 1 + #groups that come first
                const int ngw = S.misc[5];
                #pragma unroll 1
                for (int t = r; t < n; t += T) {
                    const int g = S.grp[t], sz = S.gsz[g];
                    int lev = 1;
                    #pragma unroll 1
                    for (int h = 0; h < ngw; h++)
                        if (S.gsz[h] > sz || (S.gsz[h] == sz && h < g)) lev++;
                    S.stlev[t] = (uint8_t)(lev > 255 ? 255 : lev);
                }
            }
        }
#endif
        if (serial && r == 0) {
            ng = 0;
            #pragma unroll 1
            for (int a = 0; a < n; a++) {
                int t = S.perm[a], g = -1;
                if (S.cc[t] == 0) {
                    g = 0;                                 // crosses nothing: fits the first group
                    if (ng == 0) { ng = 1; S.gsz[0] = 0; }
                } else {
                    int i = S.sti[t], j = S.stj[t];
                    #pragma unroll 1
                    for (int h = 0; h < ng && g < 0; h++) {
                        bool clash = false;
                        #pragma unroll 1
                        for (int b = 0; b < a && !clash; b++) {
                            int u = S.perm[b];
                            if (S.grp[u] == h && stems_cross(i, j, S.sti[u], S.stj[u])) clash = true;
                        }
                        if (!clash) g = h;
                    }
                    if (g < 0) { g = ng++; S.gsz[g] = 0; }
                }
                S.grp[t] = (int16_t)g;
                S.gsz[g] += S.stl[t];
            }
            // groups.sort(key=len, reverse=True) is stable: level = 1 + #groups that come first
            #pragma unroll 1
            for (int t = 0; t < n; t++) {
                int g = S.grp[t], sz = S.gsz[g], lev = 1;
                #pragma unroll 1
                for (int h = 0; h < ng; h++)
                    if (S.gsz[h] > sz || (S.gsz[h] == sz && h < g)) lev++;
                S.stlev[t] = (uint8_t)(lev > 255 ? 255 : lev);
            }
            S.misc[5] = ng;
            // does the first-created group (home of every stem that crosses nothing) rank first?
            int first = 1;
            #pragma unroll 1
            for (int h = 1; h < ng; h++) if (S.gsz[h] > S.gsz[0]) first = 0;
            S.misc[13] = first;
        }
        Team<TW>::sync();
        ng = S.misc[5];
    }
    return ng;
}

// ------------------------------------------------------------ cell score
// scoremat[i,j] of BPMatrix (seq.py:329-338), times the alignment weight
// (seq.py:1084-1085).  Only called for cells of candidate stems.
__device__ __forceinline__ double cell_score(const State &S, const DevParams &P, const DevBatch &B, int i, int j)
{
    double w = P.weight[S.code[i] * MAXK + S.code[j]];
    if (S.has_react && !S.default_reacts) {
        int a = S.rcode[i], b = S.rcode[j];
        double rf;
        if (B.rf_pos) rf = (w <= 0.0) ? __ldg(&B.rf_neg[(size_t)a * B.R + b]) : __ldg(&B.rf_pos[(size_t)a * B.R + b]);
        else {
            // too many distinct reactivities for a table: sqrt instead of libm pow(x, 0.5)
            // (differs from the reference by at most 1 ulp on ~0.1 % of inputs; documented in DESIGN.md)
            rf = sqrt(__dmul_rn(__dsub_rn(1.0, __ddiv_rn(__dadd_rn(__ldg(&B.rvals[a]), __ldg(&B.rvals[b])), 2.0)), 2.0));
            if (w <= 0.0) rf = __ddiv_rn(1.0, rf > 0.01 ? rf : 0.01);
        }
        w = __dmul_rn(w, rf);
    }
    if (S.bpp_mode) {                                   // seq.py:353-356: scoremat += term, or scoremat *= term
        const double t = __ldg(&S.bpp[(int64_t)i * S.N + j]);
        w = S.bpp_mode == 1 ? __dadd_rn(w, t) : __dmul_rn(w, t);
    }
    if (S.has_smat) w = __dmul_rn(w, __ldg(&B.smat[(int64_t)S.cols[i] * B.L + S.cols[j]]));
    return w;
}

// raw bp score of the run (a .. a+len-1) on diagonal s: Python sum(), left to right from 0 (seq.py:416)
template <class C>
__device__ __forceinline__ double run_score(const State &S, const DevParams &P, const DevBatch &B, int s, int a, int len)
{
    double sc = 0.0;
    if (C::PLAIN || (!(S.has_react && !S.default_reacts) && !S.has_smat && !S.bpp_mode)) {
        #pragma unroll 1
        for (int q = 0; q < len; q++) sc = __dadd_rn(sc, P.weight[S.code[a + q] * MAXK + S.code[s - a - q]]);
    } else {
        #pragma unroll 1
        for (int q = 0; q < len; q++) sc = __dadd_rn(sc, cell_score(S, P, B, a + q, s - a - q));
    }
    return sc;
}

// ---------------------------------------------------- one word of a diagonal
// remaining restraint pairs keep their own cell (seq.py:443) if both ends are still free
__device__ __forceinline__ uint32_t restraint_cells(const State &S, int s, int k)
{
    uint32_t x = 0;
    int a = 0, b = S.nrb;
    #pragma unroll 1
    while (a < b) { int mid = (a + b) >> 1; if (S.rbv[mid] + S.rbw[mid] < s) a = mid + 1; else b = mid; }
    #pragma unroll 1
    for (; a < S.nrb && S.rbv[a] + S.rbw[a] == s; a++) {
        int v = S.rbv[a];
        if ((v >> 5) == k && S.partner[v] < 0 && S.partner[S.rbw[a]] < 0) x |= 1u << (v & 31);
    }
    return x;
}

__device__ __forceinline__ uint32_t range_mask(int k, int lo, int hi)
{
    uint32_t x = 0xffffffffu;
    int b0 = lo - 32 * k, b1 = hi - 32 * k;
    if (b0 > 0) x &= (b0 >= 32) ? 0u : (0xffffffffu << b0);
    if (b1 < 31) x &= (b1 < 0) ? 0u : (0xffffffffu >> (31 - b1));
    return x;
}

// bit t of the result: cell (i = 32k + t, j = s - i) is a live base pair of the
// masked bool matrix AnnotateStems walks (seq.py:431-451), restricted to the
// walked range lo <= i <= hi of the diagonal.  Stateless form (fresh loads).
template <class C>
__device__ __forceinline__ uint32_t diag_word(const State &S, const DevParams &P, int s, int k, int lo, int hi)
{
    const int bo = 32 * k + (S.N - 1 - s) + 32;     // bit offset into the reversed arrays
    const int wo = bo >> 5, sh = bo & 31;
    uint32_t x = 0;
    if (C::STDP) {
        uint32_t r0 = __funnelshift_r(S.PR[wo], S.PR[wo + 1], sh);
        uint32_t r1 = __funnelshift_r(S.PR[S.WR + wo], S.PR[S.WR + wo + 1], sh);
        x = (S.M[k] ^ r0) & (S.M[S.W + k] | r1);
    } else {
        #pragma unroll 1
        for (int c = 0; c < P.npc; c++) {
            uint32_t f = S.M[c * S.W + k];
            uint32_t g = __funnelshift_r(S.PR[c * S.WR + wo], S.PR[c * S.WR + wo + 1], sh);
            x |= f & g;
        }
    }
    x &= S.rowok[k] & __funnelshift_r(S.colokR[wo], S.colokR[wo + 1], sh);
    x &= range_mask(k, lo, hi);
    if (!C::PLAIN && S.nrb) x |= restraint_cells(S, s, k);
    return x;
}

// Sliding form: walks the words k0, k0+1, ... of one diagonal and keeps the low
// halves of the funnel shifts in registers, so each word costs one new load per
// reversed array instead of two.
struct DiagWalk {
    int k, wo, sh;
    uint32_t plo[4], clo;
};

template <class C>
__device__ __forceinline__ void walk_begin(DiagWalk &it, const State &S, const DevParams &P, int s, int k0)
{
    const int bo = 32 * k0 + (S.N - 1 - s) + 32;
    it.k = k0; it.wo = bo >> 5; it.sh = bo & 31;
    if (C::STDP) {
        it.plo[0] = S.PR[it.wo]; it.plo[1] = S.PR[S.WR + it.wo]; it.plo[2] = it.plo[3] = 0u;
    } else {
        #pragma unroll
        for (int c = 0; c < 4; c++) it.plo[c] = (c < P.npc) ? S.PR[c * S.WR + it.wo] : 0u;
    }
    it.clo = S.colokR[it.wo];
}

template <class C>
__device__ __forceinline__ uint32_t walk_next(DiagWalk &it, const State &S, const DevParams &P, int s, int lo, int hi)
{
    const int k = it.k, wo = it.wo;
    uint32_t x = 0;
    if (C::STDP) {
        uint32_t h0 = S.PR[wo + 1], h1 = S.PR[S.WR + wo + 1];
        uint32_t r0 = __funnelshift_r(it.plo[0], h0, it.sh), r1 = __funnelshift_r(it.plo[1], h1, it.sh);
        x = (S.M[k] ^ r0) & (S.M[S.W + k] | r1);
        it.plo[0] = h0; it.plo[1] = h1;
    } else {
        #pragma unroll
        for (int c = 0; c < 4; c++)
            if (c < P.npc) {
                uint32_t phi = S.PR[c * S.WR + wo + 1];
                x |= S.M[c * S.W + k] & __funnelshift_r(it.plo[c], phi, it.sh);
                it.plo[c] = phi;
            }
        #pragma unroll 1
        for (int c = 4; c < P.npc; c++)
            x |= S.M[c * S.W + k] & __funnelshift_r(S.PR[c * S.WR + wo], S.PR[c * S.WR + wo + 1], it.sh);
    }
    uint32_t chi = S.colokR[wo + 1];
    x &= S.rowok[k] & __funnelshift_r(it.clo, chi, it.sh);
    it.clo = chi;
    x &= range_mask(k, lo, hi);
    if (!C::PLAIN && S.nrb) x |= restraint_cells(S, s, k);
    it.k = k + 1; it.wo = wo + 1;
    return x;
}

// walked range of diagonal s: lo..hi (inclusive) in i; false if empty
template <class C>
__device__ __forceinline__ bool diag_range(const State &S, const DevBatch &B, int s, int &lo, int &hi)
{
    const int N = S.N;
    lo = s - (N - 1); if (lo < 0) lo = 0;
    hi = (s - 4) >> 1;
    if (S.has_sep) {
        // innermost extra cell with j - i in {2, 3} when a separator follows i (seq.py:293-297)
        int i1 = hi + 1, d = s - 2 * i1;
        if (i1 >= lo && d >= 2 && inc4_of(S, i1) <= d) hi = i1;
        if (!C::PLAIN && B.interchainonly) {
            // chains differ iff a separator lies between i and j; the set of such cells is a prefix
            if (S.sepcnt[s - lo] - S.sepcnt[lo] <= 0) return false;
            int a = lo, b = hi;          // largest i in [lo, hi] with a separator in (i, s-i)
            #pragma unroll 1
            while (a < b) {
                int mid = (a + b + 1) >> 1;
                if (S.sepcnt[s - mid] - S.sepcnt[mid] > 0) a = mid; else b = mid - 1;
            }
            hi = a;
        }
    } else if (!C::PLAIN && B.interchainonly) return false;
    return hi >= lo;
}

// ---------------------------------------------------- enumerate one diagonal
// bits of x where a maximal run of at least m cells starts (xn = next word, prev_top = top bit of the previous one)
__device__ __forceinline__ uint32_t run_starts(uint32_t x, uint32_t xn, uint32_t prev_top, int m)
{
    if (!x) return 0u;
    // bit b of y: cells b .. b + have - 1 of the (x, xn) pair are all set; `have` doubles per step
    uint32_t y = x, yn = xn;
    #pragma unroll 1
    for (int have = 1; have < m;) {
        const int step = have < m - have ? have : m - have;
        y &= __funnelshift_r(y, yn, step);
        yn &= yn >> step;
        have += step;
    }
    return y & ~((x << 1) | prev_top);
}

// Lane-per-diagonal walk (warp teams, YieldStems): calls emit(a, e) for every maximal run [a, e]
// (in i) of at least P.m cells of diagonal s, outermost first -- the order of seq.py:486-493.
// The word loop is uniform across the lanes of a warp (every lane steps its own diagonal one word
// at a time).
template <class C, class F>
__device__ __forceinline__ void enum_diag(const State &S, const DevParams &P, const DevBatch &B, int s, F &&emit)
{
    int lo, hi;
    if (!diag_range<C>(S, B, s, lo, hi)) return;
    const int k0 = lo >> 5, k1 = hi >> 5;
    DiagWalk it;
    walk_begin<C>(it, S, P, s, k0);
    uint32_t x = walk_next<C>(it, S, P, s, lo, hi);
    uint32_t prev_top = 0;
    int skip_until = -1;                   // runs already emitted extend up to here
    #pragma unroll 1
    for (int k = k0; k <= k1; k++) {
        uint32_t xn = (k < k1) ? walk_next<C>(it, S, P, s, lo, hi) : 0u;
        uint32_t starts = run_starts(x, xn, prev_top, P.m);
        #pragma unroll 1
        while (starts) {
            int b = __ffs(starts) - 1;
            starts &= starts - 1;
            int a = 32 * k + b;
            if (a <= skip_until) continue;
            // find the end of the run
            int e;
            uint32_t inv = ~(x >> b);            // bit 0 is 0; first set bit = run length
            int t = inv ? __ffs(inv) - 1 : 32;
            if (b + t < 32) e = a + t - 1;
            else {
                int kk = k + 1; uint32_t w = xn;
                #pragma unroll 1
                while (kk <= k1 && w == 0xffffffffu) { kk++; w = (kk <= k1) ? diag_word<C>(S, P, s, kk, lo, hi) : 0u; }
                e = (kk <= k1) ? 32 * kk + (__ffs(~w) - 1) - 1 : 32 * (k1 + 1) - 1;
                if (e > hi) e = hi;
            }
            skip_until = e;
            emit(a, e);
        }
        prev_top = x >> 31;
        x = xn;
    }
}

// ------------------------------------------------------------- ScoreStems
// What the scan of the region confined by the innermost pair (ss, se) yields
// (seq.py:665-689).
struct Region {
    int dots, br, nedges, e0, e1, inblockend;
    bool between;
    unsigned long long levmask;
};

// the reference's own position-by-position scan
__device__ __forceinline__ void region_scan(const State &S, int ss, int se, Region &R)
{
    int dots = 0, br = 0, nedges = 0, e0 = -1, e1 = -1, inblockend = -1;
    bool between = false;
    unsigned long long levmask = 0;
    #pragma unroll 1
    for (int pos = ss + 1; pos < se; pos++) {         // seq.py:665-689
        int pr = S.partner[pos];
        if (pr < 0) {
            if (pos > inblockend) dots++;
            if (S.code[pos] == CODE_SEP) between = true;
        } else if (pr < ss || pr > se) {
            if (pos > inblockend) {
                br++;
                int lv = S.stlev[S.owner[pos]];
                levmask |= 1ull << (lv > 63 ? 63 : lv - 1);
            }
        } else if (pos < pr && pr > inblockend) {
            inblockend = pr;
            if (nedges == 0) { e0 = pos; e1 = pr; }
            nedges++;
        }
    }
    R.dots = dots; R.br = br; R.nedges = nedges; R.e0 = e0; R.e1 = e1; R.inblockend = inblockend;
    R.between = between; R.levmask = levmask;
}

// The same quantities from a walk over the ARMS of the selected stems that lie inside the
// region, in position order (binary search to the first one).  ss and se are unpaired, so each
// arm of a selected stem lies entirely inside or entirely outside (ss, se):
//   5' arm inside, 3' arm inside  -> an "inner" stem: its outermost pair is the only one that can
//                        raise inblockend (its inner pairs close earlier); the merged spans of
//                        inner stems are the "blocks";
//   exactly one arm inside        -> `len` bracket positions, counted (with the stem's level)
//                        unless a block covers the arm: inblockend at the arm's first position
//                        is the largest 3' end among the inner stems opened before it -- all of
//                        them have been met by then, because events come in position order;
//   dots              =  unpaired positions of the region outside the blocks (prefix popcounts of
//                        the unpaired mask).
// An arm (a contiguous range of paired positions of ONE stem) cannot straddle a block boundary (a
// paired position of ANOTHER stem), so testing its first position is enough.
__device__ __forceinline__ void region_stems(const State &S, int ss, int se, Region &R)
{
    const int ne = 2 * S.nst;
    int br = 0, nedges = 0, e0 = -1, e1 = -1, ibe = -1, blo = 0, covU = 0;
    unsigned long long levmask = 0;
    int q = 0;
    {
        int hi_ = ne;
        #pragma unroll 1
        while (q < hi_) { int mid = (q + hi_) >> 1; if (S.evpos[mid] <= ss) q = mid + 1; else hi_ = mid; }
    }
    #pragma unroll 1
    for (; q < ne; q++) {
        const int x = S.evpos[q];
        if (x > se) break;
        const int id = S.evid[q], t = id >> 1;
        const int i = S.sti[t], j = S.stj[t];
        if (!(id & 1)) {                                // 5' arm, x == i
            if (j < se) {                               // inner stem
                if (j > ibe) {
                    if (nedges == 0) { e0 = i; e1 = j; }
                    nedges++;
                    if (i > ibe) {                      // a new block starts: close the previous one
                        if (ibe >= 0) covU += unpaired_before(S, ibe + 1) - unpaired_before(S, blo);
                        blo = i;
                    }
                    ibe = j;
                }
                if (S.evjump) q = S.evjump[q] - 1;      // a closed stem: on to the first arm behind it
                continue;
            }
        } else if (i > ss) continue;                    // 3' arm of an inner stem
        // one-armed: 5' arm with the partner beyond se, or 3' arm with the partner left of ss
        if (x > ibe) {
            br += S.stl[t];
            int lv = S.stlev[t];
            levmask |= 1ull << (lv > 63 ? 63 : lv - 1);
        }
    }
    if (ibe >= 0) covU += unpaired_before(S, ibe + 1) - unpaired_before(S, blo);
    R.dots = unpaired_before(S, se) - unpaired_before(S, ss + 1) - covU;
    R.br = br; R.nedges = nedges; R.e0 = e0; R.e1 = e1; R.inblockend = ibe;
    R.between = S.has_sep && (S.sepcnt[se] - S.sepcnt[ss + 1] > 0);   // separators never pair
    R.levmask = levmask;
}

// pow() outside the host-built tables (never on the shipped parameter sets): kept out of line
__device__ SQRN_NOINLINE double slow_pow(double a, double b) { return pow(a, b); }

template <class C>
__device__ __forceinline__ double score_candidate(const State &S, const DevParams &P, int s, int a, int len, double bps,
                                                  int *order_out = nullptr)
{
    const int N = S.N;
    const int oi = a, oj = s - a;
    const int ss = a + len - 1, se = oj - len + 1;    // innermost pair, seq.py:655
    Region R;
    if (S.nst == 0) {
        R.dots = se - ss - 1; R.br = 0; R.nedges = 0; R.e0 = R.e1 = -1; R.inblockend = -1; R.levmask = 0;
        R.between = S.has_sep && (S.sepcnt[se] - S.sepcnt[ss + 1] > 0);
    } else {
        // the fast-lane flavour carries only the stem walk (smaller code); otherwise the cheaper
        // of the two is picked per candidate unless a test forces one
        bool by_stems = C::PLAIN || S.region_mode != REGION_SCAN;
        if (by_stems) region_stems(S, ss, se, R); else region_scan(S, ss, se, R);
    }
    // good loops = {0..4}^2 with |x - y| <= 2 (the 19 entries of seq.py:615-622)
    bool goodloop = false; int diff1 = 0;
    if (R.nedges == 1) {
        int x = R.e0 - ss - 1, y = se - R.e1 - 1, d = x > y ? x - y : y - x;
        if (x <= 4 && y <= 4 && d <= 2) { goodloop = true; diff1 = d; }
    }
    bool goodout = false; int diff2 = 0;
    if (S.nst) {                                           // seq.py:700-711
        // the nearest paired position within five on either side, from the unpaired mask: the
        // reference walks outwards over at most five unpaired positions (a sixth one, or the end of
        // the sequence, cannot close a good loop: x, y <= 4)
        const uint32_t pl = ~unpaired_bits5(S, oi - 5) & 31u;      // bit t: position oi - 5 + t is paired (or < 0)
        const uint32_t pr = ~unpaired_bits5(S, oj + 1) & 31u;      // bit t: position oj + 1 + t is paired (or >= N)
        if (pl && pr) {
            const int vv = oi - 5 + (31 - __clz(pl)), ww = oj + 1 + (__ffs(pr) - 1);
            if (vv >= 0 && ww < N && S.partner[vv] == ww) {
                int x = oi - vv - 1, y = ww - oj - 1, d = x > y ? x - y : y - x;
                if (d <= 2) { goodout = true; diff2 = d; }
            }
        }
    }
    if (!goodloop && !goodout && len < 3) return -1.0;     // seq.py:744-745
    // loopfactor, seq.py:715 (left to right, no contraction)
    double lf = __dadd_rn(1.0, __dmul_rn(__dmul_rn(P.loopbonus, goodloop ? 1.0 : 0.0), 2.0 - diff1 * 0.5));
    lf = __dadd_rn(lf, __dmul_rn(__dmul_rn(P.loopbonus, goodout ? 1.0 : 0.0), 2.0 - diff2 * 0.5));
    // tetraloop bonus, seq.py:598-604, 718
    double tf = 1.0;
    if (se - ss - 1 == 4 && S.code[ss + 1] == CODE_G &&
        (S.code[ss + 3] == CODE_G || S.code[ss + 3] == CODE_A) && S.code[ss + 4] == CODE_A) tf = 1.25;
    // stem distance factor, seq.py:721-726
    double sdf = 1.0;
    if (!R.between) {
        int ideal = R.inblockend == -1 ? 4 : 2;
        if (P.bw_is_int) {
            int k = R.dots + P.bw_int * R.br - ideal; if (k < 0) k = -k;
            sdf = (k < P.sdf_n) ? __ldg(&P.sdf_lut[k]) : slow_pow(1.0 / (1.0 + (double)k), P.distcoef);
        } else {
            double x = fabs(__dadd_rn((double)R.dots, __dmul_rn(P.bracketweight, (double)R.br)) - (double)ideal);
            sdf = slow_pow(1.0 / (1.0 + x), P.distcoef);   // documented <= 1 ulp deviation
        }
    }
    int order = __popcll(R.levmask);
    if (order_out) *order_out = order;          // 0: no wing of a selected stem counts, the levels do not enter the score
    double of = (order < P.of_n) ? __ldg(&P.of_lut[order]) : slow_pow(1.0 / (1.0 + order), P.orderpenalty);
    // seq.py:732
    double fin = __dmul_rn(bps, sdf);
    fin = __dmul_rn(fin, of);
    fin = __dmul_rn(fin, lf);
    fin = __dmul_rn(fin, tf);
    return fin;
}

// -------------------------------------------- one OptimalStems pass (arg-max)
// team-wide reductions --------------------------------------------------------
template <class C>
__device__ __forceinline__ Best team_argmax(State &S, Best best)
{
    constexpr int TW = C::TW;
#ifndef SQRN_HOST_EMU
    if (TW == 1) {
        // three warp reductions on an order-preserving integer image of the score (high word, low word)
        // and the enumeration key, instead of five shuffle rounds of (double, key, len)
        unsigned long long u = (unsigned long long)__double_as_longlong(__dadd_rn(best.fin, 0.0));      // -0.0 -> +0.0
        u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
        const uint32_t hi = (uint32_t)(u >> 32), lo = (uint32_t)u;
        const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
        bool in = hi == mh;
        const uint32_t ml = __reduce_max_sync(0xffffffffu, in ? lo : 0u);
        in = in && lo == ml;
        const uint32_t mk = __reduce_min_sync(0xffffffffu, in ? best.key : 0xffffffffu);
        in = in && best.key == mk;
        const int src = __ffs(__ballot_sync(0xffffffffu, in)) - 1;
        best.fin = __shfl_sync(0xffffffffu, best.fin, src);
        best.len = __shfl_sync(0xffffffffu, best.len, src);
        best.key = mk;
        return best;
    }
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        double f = __shfl_xor_sync(0xffffffffu, best.fin, d);
        uint32_t k = __shfl_xor_sync(0xffffffffu, best.key, d);
        int l = __shfl_xor_sync(0xffffffffu, best.len, d);
        if (better(f, k, best.fin, best.key)) { best.fin = f; best.key = k; best.len = l; }
    }
    if (TW > 1) {
        // cross-warp through the scratch words of the team
        double *sd = S.red;                              // 1 double + 2 words per warp
        uint32_t *sk = (uint32_t *)(S.red + TW);
        __syncthreads();
        int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) { sd[w] = best.fin; sk[2 * w] = best.key; sk[2 * w + 1] = (uint32_t)best.len; }
        __syncthreads();
        Best b2; b2.fin = -1e300; b2.key = 0xffffffffu; b2.len = 0;
        #pragma unroll 1
        for (int q = 0; q < TW; q++) {
            double f = sd[q]; uint32_t k = sk[2 * q];
            if (better(f, k, b2.fin, b2.key)) { b2.fin = f; b2.key = k; b2.len = (int)sk[2 * q + 1]; }
        }
        best = b2;
        __syncthreads();
    }
#endif
    return best;
}

// Upper bound of the adjusted score of ANY candidate with raw score bps: every factor of
// seq.py:732 replaced by its largest possible value, multiplied in the same order with the same
// rounding.  Rounding is monotone, so fin <= bound holds exactly (not just approximately), which
// makes skipping candidates whose bound is below the current best bit-exact.
__device__ __forceinline__ double score_bound(const DevParams &P, double bps)
{
    if (!P.ub_ok || !(bps > 0.0)) return 1e300;
    double u = __dmul_rn(bps, P.sdf_max);
    u = __dmul_rn(u, P.of_max);
    u = __dmul_rn(u, P.lf_max);
    return __dmul_rn(u, 1.25);
}

// ---- shared by the binned run lists (the persistent list of k_long and the per-sequence base list of the pool rounds)
constexpr int GL_NBIN = 256;
// monotone map of a score (bound or floor) to a bin: linear between minfinscore and five times that, clamped
__device__ __forceinline__ int gl_bin(const DevParams &P, double u)
{
    const double lo = P.minfinscore > 0.0 ? P.minfinscore : 1.0;
    if (!(u > lo)) return 0;
    const double x = (u - lo) * ((GL_NBIN - 2) / (4.0 * lo));
    return x >= (double)(GL_NBIN - 2) ? GL_NBIN - 1 : (int)x;
}

// 32 bits of the unpaired mask starting at position p0 (positions outside the sequence read as 0)
__device__ __forceinline__ uint32_t unpaired_bits32(const State &S, int p0)
{
    const int w = p0 >> 5, sh = p0 & 31;
    const uint32_t lo = (w >= 0 && w < S.W) ? S.Ub[w] : 0u, hi = (w + 1 >= 0 && w + 1 < S.W) ? S.Ub[w + 1] : 0u;
    return __funnelshift_r(lo, hi, sh);
}
__device__ __forceinline__ bool unpaired_at(const State &S, int p) { return (S.Ub[p >> 5] >> (p & 31)) & 1u; }


// ---- bulk asynchronous copies (TMA unit, cp.async.bulk) with mbarrier completion: tiles of the base list are staged in
// shared memory one tile ahead of the threads that look at them.  One elected thread arms the barrier with the byte
// count and issues the copy; every thread waits on the barrier's phase before it reads the tile.
#ifndef SQRN_HOST_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// bytes: a multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, void *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// Base list of a sequence: its maximal runs under the EMPTY structure (bp score cached), binned by the static bound
// gl_ub0 like the persistent list of k_long, built once per (sequence, parameter set) by MODE_BASE and then shared,
// read-only, by every work item of that sequence -- the partial structures of its pool (MODE_STEP) and their tails.
// An item does not enumerate anti-diagonals: it sweeps the prefix of the list whose bounds reach its score window and
// cuts each run on the fly by its own unpaired mask (team_scan).
struct alignas(16) BEnt { uint32_t key, meta; double bps; };      // key = (i + j) << 16 | i, meta = len | candidate << 31
struct BaseView { const BEnt *ent; int n; };                       // the bin ends are in State::gbend
#ifdef SQRN_HOST_EMU
static long g_emu_base_sweeps = 0;        // tests assert that the base-list path really ran
#endif

// Phase 2b for one survivor: ScoreStems' adjusted score; folds it into the lane's running best.
// Candidates that cannot reach `floor` (the best score so far, or the lower end of the subopt
// range) are skipped without evaluating their region.  Returns the adjusted score (-1e300 if it
// is below minfinscore or was skipped).
template <class C>
__device__ __forceinline__ double consider(const State &S, const DevParams &P, uint32_t key, int len, double bps,
                                           double floor, Best &best)
{
    double ub = score_bound(P, bps);
    if (ub < floor || ub < P.minfinscore) return -1e300;
    if (!C::PERSIST && P.ub_ok && bps > 0.0 && P.loopbonus >= 0.0) {
        // second, tighter bound for the candidates the first one lets through (teams that scan many
        // candidates per step): the tetraloop factor is cheap to get exactly, and a good loop needs a
        // paired position within five of the stem's inner / outer ends -- read from the unpaired mask.
        // Same multiplication order and rounding as seq.py:732 with factors >= the true ones.
        const int a = (int)(key & 0xffff), oj = (int)(key >> 16) - a;
        const int ss = a + len - 1, se = oj - len + 1;
        double tf = 1.0;
        if (se - ss - 1 == 4 && S.code[ss + 1] == CODE_G &&
            (S.code[ss + 3] == CODE_G || S.code[ss + 3] == CODE_A) && S.code[ss + 4] == CODE_A) tf = 1.25;
        bool in_ok = false, out_ok = false;
        if (S.nst) {
            in_ok = (~unpaired_bits5(S, ss + 1) & 31u) && (~unpaired_bits5(S, se - 5) & 31u);
            out_ok = (~unpaired_bits5(S, a - 5) & 31u) && (~unpaired_bits5(S, oj + 1) & 31u);
        }
        double lf = __dadd_rn(1.0, __dmul_rn(__dmul_rn(P.loopbonus, in_ok ? 1.0 : 0.0), 2.0));
        lf = __dadd_rn(lf, __dmul_rn(__dmul_rn(P.loopbonus, out_ok ? 1.0 : 0.0), 2.0));
        double u = __dmul_rn(bps, P.sdf_max);
        u = __dmul_rn(u, P.of_max);
        u = __dmul_rn(u, lf);
        u = __dmul_rn(u, tf);
        if (u < floor || u < P.minfinscore) return -1e300;
    }
    double fin = score_candidate<C>(S, P, (int)(key >> 16), (int)(key & 0xffff), len, bps);
    if (!(fin >= P.minfinscore)) return -1e300;
    if (better(fin, key, best.fin, best.key)) { best.fin = fin; best.key = key; best.len = len; }
    return fin;
}

// Enumerates every candidate stem of the current structure (AnnotateStems),
// scores it (ScoreStems) and returns the team-wide best survivor: the head of
// ChooseStems' stable sort.  fin = -1e300 when nothing reaches minfinscore.
//
// Survivors of the bp-score filter collect in a list that is flushed whenever it
// fills up.  A flush scores the new entries (skipping those whose upper bound
// cannot matter), shares the team-wide best, and then
//   subopt < 0 (TAIL): empties the list -- only the best is wanted;
//   subopt >= 0 (STEP): keeps the entries with fin >= subopt * best so far.  The
//     final range subopt * best can only be higher, so everything ChooseStems can
//     return survives; misc[7] = entries kept, cfin[] their adjusted scores
//     (misc[7] > Ccap: the in-range candidates alone overflow the list and the
//     caller retries with a larger one).
template <class C>
__device__ Best team_scan(State &S, const DevParams &P, const DevBatch &B, const Layout &L, const double subopt,
                          const BaseView *base = nullptr)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    const bool keep = subopt >= 0.0;
    Best best; best.fin = -1e300; best.key = 0xffffffffu; best.len = 0;
    if (r == 0) { S.misc[0] = 0; S.misc[6] = 0; S.misc[7] = 0; }
    Team<TW>::sync();
    const int smax = 2 * S.N - 6;
    const int Rcap = L.Rcap, Ccap = L.Ccap;
    const int rflush = Rcap - (Rcap >> 2);
    int nsurv = 0;                         // entries in the survivor list (uniform; mirrored in misc[7] for TW > 1)
    int nkept = 0;                         // STEP: the first nkept entries are already scored
    bool overflow = false;

    // phase 2b over the survivor list (team-uniform call)
    auto flush_survivors = [&]() {
        Team<TW>::sync();                  // the survivors stored by the other threads are visible
        if (!C::RUNLIST) {
            // the shared counter also counts threads that found the list full and parked with their run
            nsurv = S.misc[7];
            if (nsurv > Ccap && !overflow) nsurv = Ccap;
        }
        int ns = nsurv < Ccap ? nsurv : Ccap;
        if (nsurv > Ccap) overflow = true;
        // TAIL: a candidate whose bound is below the best so far cannot win.  STEP: nor can it enter the
        // subopt range (only used when the range lies below the best: subopt <= 1 and best > 0).
        auto floor_of = [&](const Best &b) {
            return !keep ? b.fin : (b.fin > 0.0 ? __dmul_rn(subopt < 1.0 ? subopt : 1.0, b.fin) : -1e300);
        };
        double first = -1e300;             // entries with at least this raw score are scored in a first pass
        if (!C::RUNLIST && best.fin <= -1e300 && ns - nkept > T) {
            // nothing to bound against yet: score the entries with the highest raw scores first, so that
            // the bound can dismiss most of the others (each entry is still scored at most once)
            Best q; q.fin = -1e300; q.key = 0; q.len = 0;
            #pragma unroll 1
            for (int c = nkept + r; c < ns; c += T) { double b = S.cbps[c]; if (b > q.fin) q.fin = b; }
            q = team_argmax<C>(S, q);
            first = q.fin > 0.0 ? __dmul_rn(q.fin, 0.7) : -1e300;
            #pragma unroll 1
            for (int c = nkept + r; c < ns; c += T) {
                double b = S.cbps[c];
                if (!(b >= first)) continue;
                double fin = consider<C>(S, P, S.ckey[c], S.clen[c], b, floor_of(best), best);
                if (keep) S.cfin[c] = fin;
            }
            best = team_argmax<C>(S, best);
        }
        #pragma unroll 1
        for (int c = nkept + r; c < ns; c += T) {
            double b = S.cbps[c];
            if (first > -1e300 && b >= first) continue;          // done in the first pass
            double fin = consider<C>(S, P, S.ckey[c], S.clen[c], b, floor_of(best), best);
            if (keep) S.cfin[c] = fin;
        }
        best = team_argmax<C>(S, best);    // every thread continues with the team-wide best (barrier)
        if (!keep) {
            nsurv = 0;
            if (TW > 1 || !C::RUNLIST) { Team<TW>::sync(); if (r == 0) S.misc[7] = 0; Team<TW>::sync(); }
            return;
        }
        // compact: keep what is still inside the subopt range (in place, rounds of T entries)
        const double range = best.fin > -1e300 ? __dmul_rn(subopt, best.fin) : -1e300;
        int nk = 0;
        if (TW > 1) { if (r == 0) S.misc[6] = 0; Team<TW>::sync(); }
        #pragma unroll 1
        for (int c0 = 0; c0 < ns; c0 += T) {
            const int c = c0 + r;
            bool kp = false; uint32_t key = 0; uint16_t len = 0; double sc = 0.0, fin = 0.0;
            if (c < ns) {
                fin = S.cfin[c];
                // the top candidate is always taken (seq.py:766), the others must be inside the range
                kp = fin > -1e300 && (fin >= range || (fin == best.fin && S.ckey[c] == best.key));
                if (kp) { key = S.ckey[c]; len = S.clen[c]; sc = S.cbps[c]; }
            }
            int slot;
            int cnt = Team<TW>::claim(kp, &S.misc[6], nk, slot);     // barrier: this round's reads are done
            if (kp) { S.ckey[slot] = key; S.clen[slot] = len; S.cbps[slot] = sc; S.cfin[slot] = fin; }
            nk += cnt;
        }
        Team<TW>::sync();
        nkept = nk; nsurv = nk;
        if (TW > 1 || !C::RUNLIST) { if (r == 0) S.misc[7] = nk; Team<TW>::sync(); }
        if (nk + T > Ccap) overflow = true;        // the in-range entries alone (nearly) fill the list
    };
    // one survivor of phase 2a per thread (team-uniform call)
    auto add_survivors = [&](bool push, uint32_t key, int len, double sc) {
        int slot;
        int cnt = Team<TW>::claim(push, &S.misc[7], nsurv, slot);
        if (push && slot < Ccap) { S.ckey[slot] = key; S.clen[slot] = (uint16_t)len; S.cbps[slot] = sc; }
        nsurv += cnt;
    };

    if (TW != 1 && base) {
        // The sequence has a base list: no enumeration.  Rounds over the prefix of the binned list whose static bounds
        // reach the score window known so far (the best of a round is the floor of the next; at most four times the
        // records of the rounds before).  A thread per record: runs untouched by the structure keep their cached bp
        // score, the others are cut by the unpaired mask into their live pieces, which are re-summed.  Candidates go
        // through the same survivor list / flush as the enumerating paths.
        const int nb = base->n;
        int done = 0;
#ifdef SQRN_HOST_EMU
        g_emu_base_sweeps++;
#else
        // tiles of T records go through a two-stage ring in shared memory, filled by cp.async.bulk one tile ahead
        unsigned long long *bar = (unsigned long long *)S.stage;
        uint4 *tile = (uint4 *)(S.stage + 16);
        uint32_t phase0 = S.bphase[0], phase1 = S.bphase[1];        // (the barriers are initialised once per kernel: work_loop)
        auto fetch = [&](int c0, int stage, int stop) {            // (thread 0) records [c0, min(c0 + T, stop)) -> tile[stage]
            const int cnt = stop - c0 < T ? stop - c0 : T;
            bulk_load(tile + stage * T, &base->ent[c0], (uint32_t)cnt * 16u, &bar[stage]);
        };
#endif
        #pragma unroll 1
        for (;;) {
            int target = nb;
            {
                const double f = !keep ? best.fin : (best.fin > 0.0 ? __dmul_rn(subopt < 1.0 ? subopt : 1.0, best.fin) : -1e300);
                if (f > -1e300) target = S.gbend[gl_bin(P, f)];
                const int want = done > 0 ? 4 * done : 4 * T;
                if (target > want) {
                    int q = 0, qh = GL_NBIN - 1;               // the highest bin q with gbend[q] >= want (gbend falls with q)
                    #pragma unroll 1
                    while (q < qh) { const int mid = (q + qh + 1) >> 1; if (S.gbend[mid] >= want) q = mid; else qh = mid - 1; }
                    if (S.gbend[q] < target) target = S.gbend[q];
                }
                if (target > nb) target = nb;
            }
            if (target <= done) break;
#ifndef SQRN_HOST_EMU
            int kt = 0;                                  // tile number within the round (its stage: kt & 1)
            if (r == 0) fetch(done, 0, target);         // (every thread is past the barriers of the last flush: the ring is free)
#endif
            #pragma unroll 1
            for (int c0 = done; c0 < target; c0 += T) {
                const int c = c0 + r;
                bool pending = c < target, whole = false;
                uint32_t ekey = 0, live = 0; int len = 0, a = 0, s = 0, t = 0, q = 0; double ebps = 0.0; bool cand = false;
#ifndef SQRN_HOST_EMU
                const int stage = kt & 1;
                // the tile after this one: its stage was read during the tile before this one, and every thread has
                // passed a team barrier since then (the claim of the survivor list)
                if (r == 0 && c0 + T < target) fetch(c0 + T, stage ^ 1, target);
                mbar_wait(&bar[stage], stage ? phase1 : phase0);
                if (stage) phase1 ^= 1u; else phase0 ^= 1u;
                kt++;
#endif
                if (pending) {
#ifdef SQRN_HOST_EMU
                    const BEnt e = base->ent[c];
                    ekey = e.key; len = (int)(e.meta & 0xffffu); cand = (e.meta >> 31) != 0; ebps = e.bps;
#else
                    const uint4 w = tile[stage * T + r];
                    ekey = w.x; len = (int)(w.y & 0xffffu); cand = (w.y >> 31) != 0; ebps = __hiloint2double((int)w.w, (int)w.z);
#endif
                    a = (int)(ekey & 0xffffu); s = (int)(ekey >> 16); t = s - a;
                    if (len <= 32) {
                        const uint32_t full = 0xffffffffu >> (32 - len);
                        live = unpaired_bits32(S, a) & __brev(unpaired_bits32(S, t - 31)) & full;
                        whole = live == full;
                    } else {
                        whole = true;
                        #pragma unroll 1
                        for (int k = 0; k < len && whole; k++) whole = unpaired_at(S, a + k) && unpaired_at(S, t - k);
                    }
                }
                #pragma unroll 1
                for (;;) {
                    bool push = false; uint32_t key = 0; int plen = 0; double sc = 0.0;
                    if (pending) {
                        if (whole) { push = cand; key = ekey; plen = len; sc = ebps; pending = false; }
                        else {
                            int p0 = 0, pl = 0;
                            if (len <= 32) {
                                if (live) {
                                    p0 = __ffs(live) - 1;
                                    const uint32_t inv = ~(live >> p0);
                                    pl = inv ? __ffs(inv) - 1 : 32;
                                    live = (p0 + pl >= 32) ? 0u : (live >> (p0 + pl)) << (p0 + pl);
                                }
                            } else {
                                #pragma unroll 1
                                while (q < len && !(unpaired_at(S, a + q) && unpaired_at(S, t - q))) q++;
                                p0 = q;
                                #pragma unroll 1
                                while (q < len && unpaired_at(S, a + q) && unpaired_at(S, t - q)) q++;
                                pl = q - p0;
                            }
                            if (pl == 0) pending = false;
                            else if ((double)pl >= P.minlen) {
                                sc = run_score<C>(S, P, B, s, a + p0, pl);            // re-summed from the piece's own outermost cell (seq.py:416)
                                push = sc >= P.minbpscore;
                                key = ((uint32_t)s << 16) | (uint32_t)(a + p0); plen = pl;
                            }
                        }
                    }
                    if (nsurv + T > Ccap && !overflow) flush_survivors();
                    add_survivors(push, key, plen, sc);
#ifdef SQRN_HOST_EMU
                    S.misc[7] = nsurv;                 // (the one-thread team's claim() does not use the shared counter)
#endif
                    if (!Team<TW>::any(pending)) break;
                }
            }
            if (!overflow) flush_survivors();
            done = target;
            if (done >= nb) break;
        }
#ifndef SQRN_HOST_EMU
        S.bphase[0] = phase0; S.bphase[1] = phase1;
#endif
    } else if (!C::RUNLIST) {
        // CTA teams (long sequences: hundreds of runs per diagonal): a WARP walks one anti-diagonal,
        // 32 words (1024 cells) per step, one word per lane, so the lanes do the same work at the same
        // time: build the word, find the run starts in it, score those runs, append the ones that pass
        // the bp-score filter to the survivor list.  When the list is full a lane parks (keeping the run
        // it could not store); the team meets at a barrier only when every warp is parked or done,
        // flushes the list and resumes where it stopped.
#ifdef SQRN_HOST_EMU
        constexpr int WL = 1;
        const int G = 1, lane = 0;
#else
        constexpr int WL = 32;
        // short diagonals: a sub-group of G = 4..32 lanes per diagonal (G >= words of the longest one,
        // when that is <= 32), so a warp walks 32 / G diagonals at once
        int G = 32;
        { const int wd = (S.N / 2 + 31) / 32 + 1; while (G > 4 && (G >> 1) >= wd) G >>= 1; }
        const int lane = (threadIdx.x & 31) & (G - 1);         // lane within the group
#endif
        const int gid = r / G, ng = T / G;             // this group, groups in the team
        int q = gid;                                   // index of this group's diagonal within the team's share
        int sd = 0, lo = 0, hi = 0, k1 = -1, kb = 0;   // group-uniform: diagonal, walked range, last word, chunk base
        bool have_diag = false, warp_done = false, need_chunk = true;
        uint32_t x = 0, xn = 0, starts = 0, carry_top = 0;
        int k = 0;
        bool carry = false;
        uint32_t ckey_ = 0; int clen_ = 0; double csc_ = 0.0;

        // next chunk of the current diagonal, or the first chunk of the group's next diagonal
        // (uniform within a group; the groups of a warp run it in lockstep).  false: nothing left.
        auto advance = [&]() -> bool {
            bool live = true;
            #pragma unroll 1
            for (;;) {
                if (have_diag) {
                    kb += G;
                    if (kb <= k1) break;
                    have_diag = false; q += ng;
                }
                sd = 4 + S.doffset + S.dstride * q;
                if (sd > smax) { live = false; break; }
                if (!diag_range<C>(S, B, sd, lo, hi)) { q += ng; continue; }
                have_diag = true; kb = lo >> 5; k1 = hi >> 5; carry_top = 0;
                break;
            }
            k = kb + lane;
            x = (live && k <= k1) ? diag_word<C>(S, P, sd, k, lo, hi) : 0u;
#ifdef SQRN_HOST_EMU
            xn = (live && kb + G <= k1) ? diag_word<C>(S, P, sd, kb + G, lo, hi) : 0u;
            uint32_t pt = carry_top;
            carry_top = x >> 31;
#else
            xn = __shfl_down_sync(0xffffffffu, x, 1, G);
            if (lane == G - 1) xn = (live && kb + G <= k1) ? diag_word<C>(S, P, sd, kb + G, lo, hi) : 0u;
            uint32_t xp = __shfl_up_sync(0xffffffffu, x, 1, G);
            uint32_t pt = lane == 0 ? carry_top : (xp >> 31);
            carry_top = __shfl_sync(0xffffffffu, x, G - 1, G) >> 31;
#endif
            starts = live ? run_starts(x, xn, pt, P.m) : 0u;
            return live;
        };

        #pragma unroll 1
        for (;;) {
            if (!warp_done) {
                #pragma unroll 1
                for (;;) {
                    if (need_chunk) {
                        // a group without diagonals left idles (no starts) until the whole warp is done
                        bool live = advance();
#ifndef SQRN_HOST_EMU
                        live = __any_sync(0xffffffffu, live);
#endif
                        if (!live) { warp_done = true; break; }
                        need_chunk = false;
                    }
                    bool parked = false;
                    #pragma unroll 1
                    while (starts || carry) {
                        if (!carry) {
                            if (*(volatile int *)&S.misc[7] >= Ccap && !overflow) { parked = true; break; }   // list full
                            const int b = __ffs(starts) - 1;
                            starts &= starts - 1;
                            const int a = 32 * k + b;
                            int e;
                            uint32_t inv = ~(x >> b);            // bit 0 is 0; first set bit = run length
                            int t = inv ? __ffs(inv) - 1 : 32;
                            if (b + t < 32) e = a + t - 1;
                            else {
                                int kk = k + 1; uint32_t w = xn;
                                #pragma unroll 1
                                while (kk <= k1 && w == 0xffffffffu) { kk++; w = (kk <= k1) ? diag_word<C>(S, P, sd, kk, lo, hi) : 0u; }
                                e = (kk <= k1) ? 32 * kk + (__ffs(~w) - 1) - 1 : 32 * (k1 + 1) - 1;
                                if (e > hi) e = hi;
                            }
                            const int len = e - a + 1;
                            if ((double)len < P.minlen) continue;
                            const double sc = run_score<C>(S, P, B, sd, a, len);
                            if (!(sc >= P.minbpscore)) continue;
                            ckey_ = ((uint32_t)sd << 16) | (uint32_t)a; clen_ = len; csc_ = sc;
                        }
                        int slot = atomicAdd(&S.misc[7], 1);
                        if (slot < Ccap) { S.ckey[slot] = ckey_; S.clen[slot] = (uint16_t)clen_; S.cbps[slot] = csc_; carry = false; }
                        else if (overflow) carry = false;      // STEP list overflow: the result is discarded anyway
                        else { carry = true; parked = true; break; }   // lost the race for the last slots: park with the run
                    }
#ifndef SQRN_HOST_EMU
                    parked = __any_sync(0xffffffffu, parked);
#endif
                    if (parked) break;
                    need_chunk = true;
                }
            }
            const int nd = Team<TW>::count(warp_done);           // barrier
            flush_survivors();
            if (nd == T) break;
        }
    } else {
        // warp teams: phase 1 fills a run list, phase 2a scores it with dense lanes
        auto flush_runs = [&](int nr) {
            #pragma unroll 1
            for (int c0 = 0; c0 < nr; c0 += T) {
                if (nsurv + T > Ccap && !overflow) flush_survivors();
                const int c = c0 + r;
                bool push = false; uint32_t key = 0; int len = 0; double sc = 0.0;
                if (c < nr) {
                    key = S.rkey[c]; len = S.rlen[c];
                    sc = run_score<C>(S, P, B, (int)(key >> 16), (int)(key & 0xffff), len);
                    push = sc >= P.minbpscore;
                }
                add_survivors(push, key, len, sc);
            }
            Team<TW>::sync();
            if (r == 0) S.misc[0] = 0;
            Team<TW>::sync();
        };
        #pragma unroll 1
        for (int q0 = 0; 4 + S.doffset + S.dstride * q0 <= smax; q0 += T) {
            const int s = 4 + S.doffset + S.dstride * (q0 + r);
            const bool last = 4 + S.doffset + S.dstride * (q0 + T) > smax;
            int done = -1;                     // runs of this lane's diagonal starting at a <= done are in the list
            #pragma unroll 1
            for (;;) {
                bool pending = false;          // the run list filled up before this lane's diagonal was finished
                if (s <= smax)
                    enum_diag<C>(S, P, B, s, [&](int a, int e) {
                        if (a <= done || pending) return;
                        int len = e - a + 1;
                        if ((double)len < P.minlen) return;
                        int slot = atomicAdd(&S.misc[0], 1);
                        if (slot < Rcap) { S.rkey[slot] = ((uint32_t)s << 16) | (uint32_t)a; S.rlen[slot] = (uint16_t)len; done = a; }
                        else pending = true;
                    });
                const bool anyp = Team<TW>::any(pending);
                int nr = S.misc[0]; if (nr > Rcap) nr = Rcap;
                Team<TW>::sync();             // every thread has read the count before anyone changes it again
                if (nr >= rflush || last || anyp) flush_runs(nr);
                if (!anyp) break;
            }
        }
    }
    if (!overflow) flush_survivors();
    if (r == 0) S.misc[7] = overflow ? 0x7fffffff : nsurv;
    Team<TW>::sync();
    return best;
}

// ------------------------------------------ persistent run list (MODE_TAIL)
// AnnotateStems (seq.py:427-495) is a function of the masked bool matrix, and between two greedy
// steps that matrix only LOSES cells: the rows and columns of the positions of the stem just
// selected (seq.py:446-451; restraint cells die the same way, seq.py:438-443).  So every maximal
// run of the next step is a piece of a maximal run of this one.  The list below holds the maximal
// runs of at least minlen cells; a step cuts the runs that touch the new stem into their surviving
// pieces, re-sums those (outer -> inner from 0.0, seq.py:416) and leaves the rest alone -- the
// anti-diagonals are enumerated once per sequence instead of once per step.
//
// A run whose POSITIVE cell scores sum to less than minbpscore is dropped for good: the score of
// any piece of it is <= that sum (also in floating point: rounding is monotone and every partial
// sum of the positive terms dominates the matching partial sum of the piece), so no piece can pass
// the bp-score filter of seq.py:492.
//
// entry = cand << 31 | s << 17 | a << 8 | len (s = i + j, a = outermost i), 0 = dead; cand: the run
// passes minbpscore as it stands; pbps = its bp score.
constexpr uint32_t PK_CAND = 0x80000000u;
#ifdef SQRN_HOST_EMU
static long g_emu_persist_steps = 0;      // tests assert that the persistent path really ran
#endif
__device__ __forceinline__ uint32_t pk_pack(int s, int a, int len) { return ((uint32_t)s << 17) | ((uint32_t)a << 8) | (uint32_t)len; }
__device__ __forceinline__ int pk_len(uint32_t e) { return (int)(e & 255u); }
__device__ __forceinline__ int pk_a(uint32_t e) { return (int)((e >> 8) & 511u); }
__device__ __forceinline__ int pk_s(uint32_t e) { return (int)((e >> 17) & 0x3fffu); }
__device__ __forceinline__ uint32_t pk_key(uint32_t e) { return ((uint32_t)pk_s(e) << 16) | (uint32_t)pk_a(e); }

// bits lo .. hi (clipped to 0 .. 31) set
__device__ __forceinline__ uint32_t bits_between(int lo, int hi)
{
    if (lo < 0) lo = 0;
    if (hi > 31) hi = 31;
    return lo <= hi ? (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo) : 0u;
}

// bp score of a run and the sum of its positive cells
template <class C>
__device__ __forceinline__ double run_score_pos(const State &S, const DevParams &P, const DevBatch &B, int s, int a, int len,
                                                double &pos)
{
    double sc = 0.0, ps = 0.0;
    const bool simple = C::PLAIN || (!(S.has_react && !S.default_reacts) && !S.has_smat && !S.bpp_mode);
    #pragma unroll 1
    for (int q = 0; q < len; q++) {
        double w = simple ? P.weight[S.code[a + q] * MAXK + S.code[s - a - q]] : cell_score(S, P, B, a + q, s - a - q);
        sc = __dadd_rn(sc, w);
        if (w > 0.0) ps = __dadd_rn(ps, w);
    }
    pos = ps;
    return sc;
}

// worth trying: the expected number of runs fits the list (the list falls back when it overflows)
__device__ __forceinline__ bool persist_wanted(int N, const DevParams &P, const Layout &L)
{
    if (L.Pcap <= 0 || N > 511) return false;
    const double dens = P.m >= 4 ? 0.003 : (P.m == 3 ? 0.008 : 0.02);
    return dens * N * (double)N <= 0.75 * L.Pcap;
}

// Enumerates the maximal runs of the current structure state into the persistent list.
// false: the list overflowed (the caller rescans every step, as before).
template <class C>
__device__ bool persist_build(State &S, const DevParams &P, const DevBatch &B, const Layout &L)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    const int Pcap = L.Pcap;
    const int smax = 2 * S.N - 6;
    if (r == 0) S.misc[8] = 0;
    Team<TW>::sync();
    int n = 0;                                   // entries [0, n) are scored and filtered
    #pragma unroll 1
    for (int s0 = 4; s0 <= smax; s0 += T) {
        const int s = s0 + r;
        if (s <= smax)
            enum_diag<C>(S, P, B, s, [&](int a, int e) {
                const int len = e - a + 1;
                if ((double)len < P.minlen) return;
                if (len > 255) { atomicAdd(&S.misc[8], 1 << 20); return; }
                const int slot = atomicAdd(&S.misc[8], 1);
                if (slot < Pcap) S.pkey[slot] = pk_pack(s, a, len);
            });
        Team<TW>::sync();
        const int nr = S.misc[8];
        if (nr > Pcap) return false;             // uniform
        const bool last = s0 + T > smax;
        if (nr - n < T && !last && nr + 2 * T <= Pcap) continue;
        // score the new runs with dense lanes; keep those whose positive part can reach minbpscore
        int nk = n;
        #pragma unroll 1
        for (int c0 = n; c0 < nr; c0 += T) {
            const int c = c0 + r;
            bool keep = false; uint32_t e = 0; double sc = 0.0;
            if (c < nr) {
                e = S.pkey[c];
                double pos;
                sc = run_score_pos<C>(S, P, B, pk_s(e), pk_a(e), pk_len(e), pos);
                keep = pos >= P.minbpscore;
                if (sc >= P.minbpscore) e |= PK_CAND;
            }
            int slot;
            const int cnt = Team<TW>::claim(keep, &S.misc[6], nk, slot);      // barrier: this round's reads are done
            if (keep) { S.pkey[slot] = e; S.pbps[slot] = sc; }
            nk += cnt;
        }
        n = nk;
        Team<TW>::sync();
        if (r == 0) S.misc[8] = n;
        Team<TW>::sync();
    }
    return true;
}

// One OptimalStems pass over the persistent list.  (ui, uj, ul) = the stem applied since the last
// pass (ul = 0: none): runs touching it are cut first.  ok = false: the list overflowed while
// pieces were appended (it is abandoned; the caller rescans).
template <class C>
__device__ Best persist_step(State &S, const DevParams &P, const DevBatch &B, const Layout &L, int ui, int uj, int ul, bool &ok)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    const int Pcap = L.Pcap;
    int n = S.misc[8];
    ok = true;
#ifdef SQRN_HOST_EMU
    g_emu_persist_steps++;
#endif
    if (ul > 0) {
        const int u1 = ui + ul - 1, v0 = uj - ul + 1;      // dead positions: [ui, u1] and [v0, uj]
        #pragma unroll 1
        for (int c = r; c < n; c += T) {
            const uint32_t e = S.pkey[c];
            if (!e) continue;
            const int len = pk_len(e), a = pk_a(e), s = pk_s(e), t = s - a;
            // cell q of the run is (a + q, t - q): the q-ranges whose row or column died
            const int A0 = ui - a, A1 = u1 - a, B0 = v0 - a, B1 = uj - a;
            const int C0 = t - u1, C1 = t - ui, D0 = t - uj, D1 = t - v0;
            const int top = len - 1;
            const bool hit = (A0 <= top && A1 >= 0) || (B0 <= top && B1 >= 0) || (C0 <= top && C1 >= 0) || (D0 <= top && D1 >= 0);
            if (!hit) continue;
            // the surviving pieces = runs of live cells: from a bit mask (the usual case, len <= 32) or a walk
            uint32_t live = len <= 32 ? (0xffffffffu >> (32 - len)) & ~(bits_between(A0, A1) | bits_between(B0, B1) |
                                                                        bits_between(C0, C1) | bits_between(D0, D1)) : 0u;
            auto dead = [&](int q) { return (q >= A0 && q <= A1) || (q >= B0 && q <= B1) || (q >= C0 && q <= C1) || (q >= D0 && q <= D1); };
            bool first = true;
            int q = 0;
            #pragma unroll 1
            for (;;) {
                int st, pl;
                if (len <= 32) {
                    if (!live) break;
                    st = __ffs(live) - 1;
                    const uint32_t inv = ~(live >> st);
                    pl = inv ? __ffs(inv) - 1 : 32;
                    live = (st + pl >= 32) ? 0u : (live >> (st + pl)) << (st + pl);
                } else {
                    #pragma unroll 1
                    while (q < len && dead(q)) q++;
                    if (q >= len) break;
                    st = q;
                    #pragma unroll 1
                    while (q < len && !dead(q)) q++;
                    pl = q - st;
                }
                if ((double)pl < P.minlen) continue;
                // re-summed from the piece's own outermost cell (seq.py:416)
                double pos;
                const double sc = run_score_pos<C>(S, P, B, s, a + st, pl, pos);
                if (!(pos >= P.minbpscore)) continue;
                const uint32_t ne = pk_pack(s, a + st, pl) | (sc >= P.minbpscore ? PK_CAND : 0u);
                const int slot = first ? c : atomicAdd(&S.misc[8], 1);
                if (slot < Pcap) { S.pkey[slot] = ne; S.pbps[slot] = sc; }
                first = false;
            }
            if (first) S.pkey[c] = 0u;
        }
        Team<TW>::sync();
        n = S.misc[8];
        if (n > Pcap) { ok = false; n = Pcap; }
        Team<TW>::sync();
        if (!ok) return Best{-1e300, 0xffffffffu, 0};
    }
    Best best; best.fin = -1e300; best.key = 0xffffffffu; best.len = 0;
    double floor = -1e300;
    #pragma unroll 1
    for (int c0 = 0; c0 < n; c0 += T) {
        const int c = c0 + r;
        if (c < n) {
            const uint32_t e = S.pkey[c];
            if (e & PK_CAND)
                consider<C>(S, P, pk_key(e), pk_len(e), S.pbps[c], best.fin > floor ? best.fin : floor, best);
        }
#ifndef SQRN_HOST_EMU
        if (TW == 1 && c0 + T < n) {
            // share a lower bound of the best score so far (its high word): what cannot reach it is skipped
            unsigned hi = best.fin > 0.0 ? (unsigned)__double2hiint(best.fin) : 0u;
            hi = __reduce_max_sync(0xffffffffu, hi);
            const double f = __hiloint2double((int)hi, 0);
            if (hi && f > floor) floor = f;
        }
#endif
    }
    return team_argmax<C>(S, best);
}

// ------------------------- persistent candidate list in global memory (CTA teams)
// Long sequences have 10^5 .. 10^6 maximal runs and hundreds of greedy steps; rescanning them
// costs O(N^2) cell visits plus a ScoreStems region walk per surviving candidate per step.  Here
// the runs are enumerated ONCE into a per-CTA list in global memory (key, length, bp score) and
// every record CACHES what the last evaluation found:
//   FRESH   nothing known;
//   EVAL    v = the adjusted score (ScoreStems) under the structure it was evaluated with;
//   PRUNED  v = an upper bound of the adjusted score (tight_bound) that was below the best score of
//           the step that looked at it;  PRUNED1: the same with score_bound, which does not depend on
//           the structure and therefore never goes stale;
//   BELOW   the run's bp score is under minbpscore: no candidate as it stands, kept because a piece of
//           it may be one after a cut.
//
// The list is BINNED by a static upper bound.  gl_ub0 bounds the adjusted score of a run AND of every
// piece a later cut can leave of it (sum of the positive cells x the largest value of every factor of
// seq.py:732, the tetraloop factor exact); gl_bin maps it monotonically to one of 256 bins and the
// records lie in descending bin order (two enumeration passes: histogram, then scatter).  A greedy
// step only has to look at the records whose bound reaches its best score -- a PREFIX of the list:
//   * the prefix [0, done) is found by rounds: the floor known so far (last step's runner-ups, then
//     the best of the round) gives the bin the next round has to reach;
//   * records behind the prefix are not touched at all -- not even cut.  When the floor drops and
//     they enter the prefix they are CAUGHT UP from the unpaired mask (a cell is live iff both of
//     its positions are still unpaired: the masked matrix only loses cells, seq.py:446-451) and
//     start FRESH;
//   * records that were in the prefix of the previous step are up to date but for the stem T just
//     selected: they are revised incrementally (cut if T meets them; cached value dropped if T may
//     have changed it -- see below);
//   * pieces that cuts append live in an unsorted tail behind the bins and are always swept;
//   * every GL_REBUILD steps the whole list is caught up, the dead records are dropped and the
//     survivors are re-binned into the other half of the slot (their bounds only shrink).
// When does the adjusted score of a candidate c = (oi, oj, len), innermost pair (ss, se), change?
// ScoreStems reads (seq.py:607-751) the partners of the positions in (ss, se) and within five of
// the outermost pair, and the levels of the selected stems with a wing inside (ss, se).  So T can
// only matter if one of its arms meets [oi - 5, oj + 5] -- and not even then if T lies inside a
// selected stem B that itself lies inside (ss, se): positions under B are behind `inblockend`
// (seq.py:672-689) both before and after.  The host of T ("encloser": the selected stem with the
// largest i that encloses T) is found once per step.  Levels: if the level of any OLD stem
// changed when T was added, every EVAL record that counted a wing goes back to FRESH.
#ifdef SQRN_HOST_EMU
static inline long long gl_clock() { return 0; }
static long g_emu_gl_rebuilds = 0, g_emu_gl_catchups = 0;      // tests assert that these paths really ran
static int g_emu_gl_rebuild_every = 0;                          // tests: rebuild period (0: default)
#else
__device__ __forceinline__ long long gl_clock() { return clock64(); }
#endif
constexpr uint32_t GK_DEAD = 0xffffffffu;
constexpr uint32_t GS_FRESH = 0u, GS_EVAL = 1u, GS_PRUNED = 2u, GS_BELOW = 3u, GS_PRUNED1 = 4u;
constexpr uint32_t GS_MASK = 7u, GS_LEVELS = 8u;      // GS_LEVELS (with EVAL): the score involves pseudoknot levels
constexpr int GL_REBUILD = 32;    // default rebuild period (DevWork::g_rebuild overrides)

// one record = one 16-byte load: key = (i + j) << 16 | i, meta = len | state << 16 | stamp << 20, v = the cached score / bound
struct alignas(16) GEnt { uint32_t key, meta; double v; };
// a slot holds two halves of `cap` records (the rebuild copies from one into the other)
struct GList { GEnt *ent; double *bps; uint8_t *qb; int cap; unsigned long long *stat; int *cnt; int rebuild; };
// what the team remembers about its list between two passes (uniform over the team / cluster)
struct GState { int n_sorted, n_inc, half, since; };
constexpr int GL_BATCH = 8;       // records a thread loads back to back before it looks at any (memory-level parallelism)
constexpr int GL_CNT_INTS = 16 + 16 * GL_NBIN;      // cluster flavour: ints per slot in DevWork::g_cnt ([0] records, [1] chunk counter, [16 + 256 rank + q] histograms)

#ifdef SQRN_HOST_EMU
constexpr int GL_WL = 1;
static inline uint32_t gl_ballot(bool p) { return p ? 1u : 0u; }
static inline int gl_bcast0(int v) { return v; }
static inline void gl_wsync() {}
static inline int gl_lane() { return 0; }
static inline int gl_wid() { return 0; }
#else
constexpr int GL_WL = 32;
__device__ __forceinline__ uint32_t gl_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
__device__ __forceinline__ int gl_bcast0(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void gl_wsync() { __syncwarp(); }
__device__ __forceinline__ int gl_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ int gl_wid() { return threadIdx.x >> 5; }
#endif

__device__ __forceinline__ GEnt gl_load(const GEnt *p)
{
#ifdef SQRN_HOST_EMU
    return *p;
#else
    const uint4 q = *reinterpret_cast<const uint4 *>(p);
    GEnt e; e.key = q.x; e.meta = q.y; e.v = __hiloint2double((int)q.w, (int)q.z);
    return e;
#endif
}
__device__ __forceinline__ void gl_store(GEnt *p, uint32_t key, uint32_t meta, double v)
{
#ifdef SQRN_HOST_EMU
    p->key = key; p->meta = meta; p->v = v;
#else
    uint4 q; q.x = key; q.y = meta; q.z = (uint32_t)__double2loint(v); q.w = (uint32_t)__double2hiint(v);
    *reinterpret_cast<uint4 *>(p) = q;
#endif
}

// the second, tighter bound of consider() as a function: exact tetraloop factor, loop bonuses only
// where a paired position within five makes a good loop possible
__device__ __forceinline__ double tight_bound(const State &S, const DevParams &P, uint32_t key, int len, double bps)
{
    const int a = (int)(key & 0xffff), oj = (int)(key >> 16) - a;
    const int ss = a + len - 1, se = oj - len + 1;
    double tf = 1.0;
    if (se - ss - 1 == 4 && S.code[ss + 1] == CODE_G &&
        (S.code[ss + 3] == CODE_G || S.code[ss + 3] == CODE_A) && S.code[ss + 4] == CODE_A) tf = 1.25;
    bool in_ok = false, out_ok = false;
    if (S.nst) {
        in_ok = (~unpaired_bits5(S, ss + 1) & 31u) && (~unpaired_bits5(S, se - 5) & 31u);
        out_ok = (~unpaired_bits5(S, a - 5) & 31u) && (~unpaired_bits5(S, oj + 1) & 31u);
    }
    double lf = __dadd_rn(1.0, __dmul_rn(__dmul_rn(P.loopbonus, in_ok ? 1.0 : 0.0), 2.0));
    lf = __dadd_rn(lf, __dmul_rn(__dmul_rn(P.loopbonus, out_ok ? 1.0 : 0.0), 2.0));
    double u = __dmul_rn(bps, P.sdf_max);
    u = __dmul_rn(u, P.of_max);
    u = __dmul_rn(u, lf);
    return __dmul_rn(u, tf);
}

// Static bound of the run (s, a, len) and of every piece of it: `pos` (the sum of its positive cells)
// bounds the bp score of any piece, every factor of seq.py:732 takes its largest value in the same
// order with the same rounding (score_bound), and the tetraloop factor is 1.25 only if the run holds
// the one cell of its diagonal that closes a four-position hairpin with the GNRA pattern (a piece
// can only have it as its innermost cell).
__device__ __forceinline__ double gl_ub0(const State &S, const DevParams &P, int s, int a, int len, double pos)
{
    if (!(pos > 0.0)) return 1e300;
    double tf = 1.0;
    const int t = s - a, d = t - a - 5;
    if (d >= 0 && !(d & 1) && (d >> 1) < len) {
        const int ss = a + (d >> 1);
        if (S.code[ss + 1] == CODE_G && (S.code[ss + 3] == CODE_G || S.code[ss + 3] == CODE_A) && S.code[ss + 4] == CODE_A) tf = 1.25;
    }
    double u = __dmul_rn(pos, P.sdf_max);
    u = __dmul_rn(u, P.of_max);
    u = __dmul_rn(u, P.lf_max);
    return __dmul_rn(u, tf);
}

__device__ __forceinline__ bool gl_wanted(int N, const DevParams &P, long long cap)
{
    if (cap <= 0 || N > 32767 || !P.ub_ok || !(P.loopbonus >= 0.0)) return false;
    const double dens = P.m >= 4 ? 0.003 : (P.m == 3 ? 0.008 : 0.02);
    return dens * N * (double)N <= 0.8 * (double)cap;
}

template <class C> __device__ __forceinline__ Best cluster_best(State &S, Best b, int step);

// barrier of everything that works on the item: the CTA, or the thread-block cluster (whose CTAs share
// one list; its counters then live in global memory, DevWork::g_cnt)
template <class C>
__device__ __forceinline__ void gl_sync()
{
#ifndef SQRN_HOST_EMU
    if (C::CLUSTER) { __threadfence(); cooperative_groups::this_cluster().sync(); return; }
#endif
    Team<C::TW>::sync();
}
template <class C>
__device__ __forceinline__ bool gl_leader(const State &S) { return Team<C::TW>::rank() == 0 && (!C::CLUSTER || S.doffset == 0); }

// Calls emit(s, a, len) for every maximal run of at least P.m cells of the current masked matrix: a warp
// per anti-diagonal, a word per lane.  Diagonals are dealt statically (two calls visit the same runs in
// the same threads).
template <class C, class F>
__device__ __forceinline__ void gl_enum(const State &S, const DevParams &P, const DevBatch &B, F &&emit)
{
    constexpr int TW = C::TW;
    const int lane = gl_lane(), wid = gl_wid();
    const int nw = TW > 0 ? TW : 1;
    const int smax = 2 * S.N - 6;
    const int gw = C::CLUSTER ? S.doffset * nw + wid : wid, gnw = C::CLUSTER ? S.dstride * nw : nw;
    #pragma unroll 1
    for (int s = 4 + gw; s <= smax; s += gnw) {
        int lo, hi;
        if (!diag_range<C>(S, B, s, lo, hi)) continue;
        const int k1 = hi >> 5;
        uint32_t carry_top = 0;
        #pragma unroll 1
        for (int kb = lo >> 5; kb <= k1; kb += GL_WL) {
            const int k = kb + lane;
            const uint32_t x = k <= k1 ? diag_word<C>(S, P, s, k, lo, hi) : 0u;
#ifdef SQRN_HOST_EMU
            const uint32_t xn = k + 1 <= k1 ? diag_word<C>(S, P, s, k + 1, lo, hi) : 0u;
            const uint32_t pt = carry_top;
            carry_top = x >> 31;
#else
            uint32_t xn = __shfl_down_sync(0xffffffffu, x, 1);
            if (lane == 31) xn = kb + 32 <= k1 ? diag_word<C>(S, P, s, kb + 32, lo, hi) : 0u;
            const uint32_t xp = __shfl_up_sync(0xffffffffu, x, 1);
            const uint32_t pt = lane == 0 ? carry_top : (xp >> 31);
            carry_top = __shfl_sync(0xffffffffu, x, 31) >> 31;
#endif
            uint32_t starts = run_starts(x, xn, pt, P.m);
            #pragma unroll 1
            while (starts) {
                const int b = __ffs(starts) - 1;
                starts &= starts - 1;
                const int a = 32 * k + b;
                int e;
                const uint32_t inv = ~(x >> b);
                const int t = inv ? __ffs(inv) - 1 : 32;
                if (b + t < 32) e = a + t - 1;
                else {
                    int kk = k + 1; uint32_t w = xn;
                    #pragma unroll 1
                    while (kk <= k1 && w == 0xffffffffu) { kk++; w = (kk <= k1) ? diag_word<C>(S, P, s, kk, lo, hi) : 0u; }
                    e = (kk <= k1) ? 32 * kk + (__ffs(~w) - 1) - 1 : 32 * (k1 + 1) - 1;
                    if (e > hi) e = hi;
                }
                const int len = e - a + 1;
                if ((double)len < P.minlen) continue;
                emit(s, a, len);
            }
        }
    }
}

// From the team's histogram S.ghist[] (records per bin counted by THIS CTA): S.gbend[q] = records in the bins
// >= q (bins lie in descending order: bin q occupies [gbend[q + 1], gbend[q])), S.gcur[q] = where this CTA's
// records of bin q start.  Returns the number of records.  A cluster adds up the histograms of its CTAs
// through global memory; each CTA's records of a bin follow those of the lower ranks.
template <class C>
__device__ int gl_make_bins(State &S, const GList &g)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    Team<TW>::sync();             // the histogram is complete
#ifndef SQRN_HOST_EMU
    if (C::CLUSTER) {
        int *gh = g.cnt + 16;
        #pragma unroll 1
        for (int q = r; q < GL_NBIN; q += T) gh[GL_NBIN * S.doffset + q] = S.ghist[q];
        gl_sync<C>();
        #pragma unroll 1
        for (int q = r; q < GL_NBIN; q += T) {
            int tot = 0, before = 0;
            #pragma unroll 1
            for (int k = 0; k < S.dstride; k++) { const int h = *(volatile int *)&gh[GL_NBIN * k + q]; tot += h; if (k < S.doffset) before += h; }
            S.ghist[q] = tot; S.gcur[q] = before;
        }
        Team<TW>::sync();
    } else
#endif
    {
        #pragma unroll 1
        for (int q = r; q < GL_NBIN; q += T) S.gcur[q] = 0;
        Team<TW>::sync();
    }
    if (r == 0) {
        int acc = 0;
        S.gbend[GL_NBIN] = 0;
        #pragma unroll 1
        for (int q = GL_NBIN - 1; q >= 0; q--) { S.gcur[q] += acc; acc += S.ghist[q]; S.gbend[q] = acc; }
    }
    gl_sync<C>();                 // (cluster: nobody overwrites its histogram in global memory before everyone has read it)
    return S.gbend[0];
}

// Enumerates the maximal runs of the current structure state into the global list, binned by their static
// bound: ONE enumeration writes the records, unsorted, into the second half of the slot; a count pass and a scatter
// pass over those records (chunks dealt statically, so that a cluster's CTAs scatter what they counted) put them in
// bin order into the first half.  false: overflow.
template <class C>
__device__ bool gl_build(State &S, const DevParams &P, const DevBatch &B, const GList &g, GState &gs)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    int *np = C::CLUSTER ? &g.cnt[0] : &S.misc[8];        // records in the list
    GEnt *ent2 = g.ent + g.cap; double *bps2 = g.bps + g.cap; uint8_t *qb2 = g.qb + g.cap;
    if (gl_leader<C>(S)) *np = 0;
    #pragma unroll 1
    for (int q = r; q < GL_NBIN; q += T) S.ghist[q] = 0;
    gl_sync<C>();
    gl_enum<C>(S, P, B, [&](int s, int a, int len) {
        double pos;
        const double sc = run_score_pos<C>(S, P, B, s, a, len, pos);
        if (!(pos >= P.minbpscore)) return;
        const int slot = atomicAdd(np, 1);
        if (slot < g.cap) {
            gl_store(&ent2[slot], ((uint32_t)s << 16) | (uint32_t)a, (uint32_t)len | ((sc >= P.minbpscore ? GS_FRESH : GS_BELOW) << 16), 0.0);
            bps2[slot] = sc;
            qb2[slot] = (uint8_t)gl_bin(P, gl_ub0(S, P, s, a, len, pos));
        }
    });
    gl_sync<C>();
    const int n0 = *(volatile int *)np;
    gs.n_sorted = n0; gs.n_inc = n0; gs.half = 0; gs.since = 0;
    ((int *)S.cbps)[r] = -1;                          // gl_step: no remembered best yet
    if (n0 > g.cap) { gl_sync<C>(); return false; }
    const int lane = gl_lane();
    const int nw = TW > 0 ? TW : 1;
    const int gw = C::CLUSTER ? S.doffset * nw + gl_wid() : gl_wid(), gnw = C::CLUSTER ? S.dstride * nw : nw;
    #pragma unroll 1
    for (int c0 = gw * GL_WL; c0 < n0; c0 += gnw * GL_WL) {
        const int c = c0 + lane;
        if (c < n0) atomicAdd(&S.ghist[qb2[c]], 1);
    }
    gl_make_bins<C>(S, g);
    #pragma unroll 1
    for (int c0 = gw * GL_WL; c0 < n0; c0 += gnw * GL_WL) {
        const int c = c0 + lane;
        if (c < n0) {
            const GEnt e = gl_load(&ent2[c]);
            const int q = qb2[c];
            const int slot = atomicAdd(&S.gcur[q], 1);
            gl_store(&g.ent[slot], e.key, e.meta, e.v);
            g.bps[slot] = bps2[c]; g.qb[slot] = (uint8_t)q;
        }
    }
    gl_sync<C>();
    return true;
}

// The base list of a sequence (BEnt, above): the same two enumeration passes, records and bin ends written to the
// sequence's slot in global memory.  *n_out = records, or -1 when the slot is too small (its items then enumerate).
template <class C>
__device__ void base_build(State &S, const DevParams &P, const DevBatch &B, BEnt *ent, long long cap, int32_t *bend_out, int32_t *n_out)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    #pragma unroll 1
    for (int q = r; q < GL_NBIN; q += T) S.ghist[q] = 0;
    Team<TW>::sync();
    gl_enum<C>(S, P, B, [&](int s, int a, int len) {
        double pos;
        run_score_pos<C>(S, P, B, s, a, len, pos);
        if (!(pos >= P.minbpscore)) return;
        atomicAdd(&S.ghist[gl_bin(P, gl_ub0(S, P, s, a, len, pos))], 1);
    });
    GList none{};
    const int n = gl_make_bins<C>(S, none);
    if ((long long)n > cap || !P.ub_ok || !(P.loopbonus >= 0.0)) { if (r == 0) *n_out = -1; Team<TW>::sync(); return; }
    gl_enum<C>(S, P, B, [&](int s, int a, int len) {
        double pos;
        const double sc = run_score_pos<C>(S, P, B, s, a, len, pos);
        if (!(pos >= P.minbpscore)) return;
        const int slot = atomicAdd(&S.gcur[gl_bin(P, gl_ub0(S, P, s, a, len, pos))], 1);
        BEnt e; e.key = ((uint32_t)s << 16) | (uint32_t)a; e.meta = (uint32_t)len | (sc >= P.minbpscore ? 0x80000000u : 0u); e.bps = sc;
        ent[slot] = e;
    });
    #pragma unroll 1
    for (int q = r; q <= GL_NBIN; q += T) bend_out[q] = S.gbend[q];
    if (r == 0) *n_out = n;
    Team<TW>::sync();
}

// One OptimalStems pass over the global list.  (ui, uj, ul): the stem T applied since the last pass (ul = 0:
// none); (ei, ej): the selected stem enclosing T with the largest i (ei < 0: none); relevel: the level of some
// older stem changed when T was added; step: number of this pass (stamps the records it has brought up to
// date, so that nothing is revised twice).  ok = false: overflow.
//
//   rebuild   (every g.rebuild passes) the whole list is brought up to date, dead records are dropped and the
//             rest is re-binned into the other half of the slot;
//   prologue  every thread re-checks the record that was ITS best in the previous pass; what is still a valid
//             cached score gives the floor before the sweep starts (usually last pass's runner-up);
//   rounds    the prefix of the binned list whose bounds reach the floor (plus the tail of appended pieces) is
//             swept -- per record: cut / invalidate against T (`revise`), arg-max of the cached scores that
//             hold; FRESH records or PRUNED ones whose bound reaches the floor go to a per-warp list that is
//             SCREENED 32 at a time (bounds); what passes goes to a second list EVALUATED 32 at a time.  The
//             best of a round is the floor of the next one, until the prefix stops growing;
//   epilogue  the pieces appended by the cuts of this pass (behind the old end of the list) are screened too.
template <class C>
__device__ Best gl_step(State &S, const DevParams &P, const DevBatch &B, const GList &g, GState &gs, int ui, int uj, int ul,
                        int ei, int ej, bool relevel, bool &ok, int &xc, int step)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    ok = true;
    int *np = C::CLUSTER ? &g.cnt[0] : &S.misc[8];        // records in the list
    int *cc = C::CLUSTER ? &g.cnt[1] : &S.misc[12];       // next chunk of the sweep
    int *top = (int *)S.cbps;                             // per thread: the record that was its best in the last pass
    GEnt *ent = g.ent + (size_t)gs.half * g.cap;
    double *bps = g.bps + (size_t)gs.half * g.cap;
    uint8_t *qb = g.qb + (size_t)gs.half * g.cap;
    int n = *(volatile int *)np;
    const uint32_t stamp = (uint32_t)(step % 4095 + 1) << 20;       // never 0 (= not stamped); the caller stops at 4094 passes
    Best best; best.fin = -1e300; best.key = 0xffffffffu; best.len = 0;
    int best_idx = -1;
    const int u1 = ui + ul - 1, v0 = uj - ul + 1;          // the arms of T: [ui, u1] and [v0, uj]
    unsigned n_eval = 0, n_reset = 0, n_cut = 0, n_swept = 0;
    GEnt buf[GL_BATCH];
    const int lane = gl_lane();
    constexpr int CH = GL_WL * GL_BATCH;                  // records per chunk
    // warps take chunks of consecutive records from a shared counter (cuts and evaluations are unevenly spread
    // over the list; fixed strides would leave most warps waiting at the barriers)
    auto grab = [&]() {
        int c0 = 0;
        if (lane == 0) c0 = atomicAdd(cc, CH);
        return gl_bcast0(c0);
    };
    unsigned *fl_hi = (unsigned *)&S.misc[9];     // high word of the best score found so far (a lower bound of it)
    double floor = -1e300;
    auto refresh_floor = [&]() {
        const unsigned h = *(volatile unsigned *)fl_hi;
        const double f = __hiloint2double((int)h, 0);
        if (h && f > floor) floor = f;
    };
    auto offer = [&](double fin, uint32_t key, int len, int c) {          // a valid adjusted score
        if (fin >= P.minfinscore && better(fin, key, best.fin, best.key)) {
            best.fin = fin; best.key = key; best.len = len; best_idx = c;
            if (fin > floor) { floor = fin; if (fin > 0.0) atomicMax(fl_hi, (unsigned)__double2hiint(fin)); }
        }
    };
    // Replaces record c = (s, a, len), some of whose cells died, by its surviving pieces (the first one stays in
    // the slot, the others are appended).  Pieces are FRESH (or BELOW) and carry this pass's stamp.
    auto cut = [&](int c, GEnt &e, int s, int a, int len) -> uint32_t {
        const int t = s - a;
        uint32_t live = 0u;
        if (len <= 32) live = unpaired_bits32(S, a) & __brev(unpaired_bits32(S, t - 31)) & (0xffffffffu >> (32 - len));
        auto dead = [&](int q) { return !(unpaired_at(S, a + q) && unpaired_at(S, t - q)); };
        bool first = true;
        int q = 0;
        #pragma unroll 1
        for (;;) {
            int p0, pl;
            if (len <= 32) {
                if (!live) break;
                p0 = __ffs(live) - 1;
                const uint32_t inv = ~(live >> p0);
                pl = inv ? __ffs(inv) - 1 : 32;
                live = (p0 + pl >= 32) ? 0u : (live >> (p0 + pl)) << (p0 + pl);
            } else {
                #pragma unroll 1
                while (q < len && dead(q)) q++;
                if (q >= len) break;
                p0 = q;
                #pragma unroll 1
                while (q < len && !dead(q)) q++;
                pl = q - p0;
            }
            if ((double)pl < P.minlen) continue;
            double pos;
            const double sc = run_score_pos<C>(S, P, B, s, a + p0, pl, pos);
            if (!(pos >= P.minbpscore)) continue;
            const int slot = first ? c : atomicAdd(np, 1);
            const uint32_t nm = (uint32_t)pl | ((sc >= P.minbpscore ? GS_FRESH : GS_BELOW) << 16) | stamp;
            if (slot < g.cap) {
                gl_store(&ent[slot], ((uint32_t)s << 16) | (uint32_t)(a + p0), nm, 0.0);
                bps[slot] = sc;
                qb[slot] = (uint8_t)gl_bin(P, gl_ub0(S, P, s, a + p0, pl, pos));
            }
            if (first) { e.key = ((uint32_t)s << 16) | (uint32_t)(a + p0); e.meta = nm; }
            first = false;
        }
        n_cut++;
        if (first) { ent[c].key = GK_DEAD; e.key = GK_DEAD; return GS_BELOW; }
        return (e.meta >> 16) & GS_MASK;          // the piece that stayed in this slot
    };
    // Brings record c (loaded into e) up to date.  `stale` = false: it is up to date but for T -- it is cut if T
    // meets it, and its cached value goes back to FRESH when T may have changed it.  `stale` = true: it has not
    // been looked at for several passes -- it is cut by the unpaired mask and loses any cached value that depends
    // on the structure.  Returns the state it is in now (GK_DEAD records report GS_BELOW: nothing more to do with
    // them in this pass).
    auto revise = [&](int c, GEnt &e, bool stale) -> uint32_t {
        if (e.key == GK_DEAD) return GS_BELOW;
        const uint32_t meta = e.meta;
        uint32_t st = (meta >> 16) & GS_MASK;
        if ((meta & 0xfff00000u) == stamp) return st;                     // already revised in this pass
        const int len = (int)(meta & 0xffffu), a = (int)(e.key & 0xffffu), s = (int)(e.key >> 16), t = s - a;
        if (stale) {
#ifdef SQRN_HOST_EMU
            g_emu_gl_catchups++;
#endif
            bool hit;
            if (len <= 32) {
                const uint32_t full = 0xffffffffu >> (32 - len);
                hit = (unpaired_bits32(S, a) & __brev(unpaired_bits32(S, t - 31)) & full) != full;
            } else {
                hit = false;
                #pragma unroll 1
                for (int q = 0; q < len && !hit; q++) hit = !(unpaired_at(S, a + q) && unpaired_at(S, t - q));
            }
            if (hit) return cut(c, e, s, a, len);
            if (st == GS_EVAL || st == GS_PRUNED) { st = GS_FRESH; e.meta = (uint32_t)len | stamp; ent[c].meta = e.meta; n_reset++; }
            return st;
        }
        if (ul <= 0) return st;
        // T can cut the run or change its cached score only if one of its arms meets [a - 5, t + 5]
        // (the run's rows and columns all lie in [a, t]); most records fail this test and are done
        if (!((u1 < a - 5 || ui > t + 5) && (uj < a - 5 || v0 > t + 5))) {
            const int A0 = ui - a, A1 = u1 - a, B0 = v0 - a, B1 = uj - a;
            const int C0 = t - u1, C1 = t - ui, D0 = t - uj, D1 = t - v0;
            const int topq = len - 1;
            if ((A0 <= topq && A1 >= 0) || (B0 <= topq && B1 >= 0) || (C0 <= topq && C1 >= 0) || (D0 <= topq && D1 >= 0))
                return cut(c, e, s, a, len);
            if (st == GS_EVAL || st == GS_PRUNED) {
                // T meets the window the cached value was computed from: stale unless T is shielded
                const int ss = a + len - 1, se = t - len + 1;
                const bool shielded = ei >= 0 && ss < ei && ej < se;
                if (!shielded) { st = GS_FRESH; e.meta = (uint32_t)len | stamp; ent[c].meta = e.meta; n_reset++; return st; }
            }
        }
        // a level change only matters to scores that counted the wing of a selected stem
        if (relevel && st == GS_EVAL && (meta >> 16 & GS_LEVELS)) { st = GS_FRESH; e.meta = (uint32_t)len | stamp; ent[c].meta = e.meta; n_reset++; }
        return st;
    };
    auto evaluate = [&](int c) {
        const GEnt e = gl_load(&ent[c]);
        const int len = (int)(e.meta & 0xffffu);
        const double sc = bps[c];
        int order = 0;
        const double fin = score_candidate<C>(S, P, (int)(e.key >> 16), (int)(e.key & 0xffffu), len, sc, &order);
        gl_store(&ent[c], e.key, (uint32_t)len | ((GS_EVAL | (order ? GS_LEVELS : 0u)) << 16) | stamp, fin);
        n_eval++;
#ifdef SQRN_EMU_DEBUG
        printf("  eval nst=%d (%d,%d,%d) bps=%g fin=%.17g\n", S.nst, (int)(e.key & 0xffff), (int)(e.key >> 16) - (int)(e.key & 0xffff), len, sc, fin);
#endif
        offer(fin, e.key, len, c);
    };
    // true: record c (already revised) has to be evaluated: its bounds reach the floor
    auto screen = [&](int c) -> bool {
        const GEnt e = gl_load(&ent[c]);
        if (e.key == GK_DEAD) return false;
        const uint32_t st = (e.meta >> 16) & GS_MASK;
        if (st == GS_EVAL || st == GS_BELOW) return false;
        if ((st == GS_PRUNED || st == GS_PRUNED1) && e.v < floor) return false;
        const int len = (int)(e.meta & 0xffffu);
        const double sc = bps[c];
        double ub = score_bound(P, sc);
        if (ub < floor || ub < P.minfinscore) { if (st != GS_PRUNED1) gl_store(&ent[c], e.key, (uint32_t)len | (GS_PRUNED1 << 16) | stamp, ub); return false; }
        ub = tight_bound(S, P, e.key, len, sc);
        if (ub < floor || ub < P.minfinscore) {
            // (a record that was PRUNED with this very bound stays as it is: no store, no dirty sector)
            if (!(st == GS_PRUNED && ub == e.v)) gl_store(&ent[c], e.key, (uint32_t)len | (GS_PRUNED << 16) | stamp, ub);
            return false;
        }
        return true;
    };
    auto wants_look = [&](uint32_t st, const GEnt &e) {
        return st == GS_FRESH || ((st == GS_PRUNED || st == GS_PRUNED1) && !(e.v < floor));
    };
    // per-warp lists: `wa` records to screen, `wb` records to evaluate (Layout::Ccap >= 4 GL_WL per warp)
    int *wa = (int *)S.ckey + 4 * GL_WL * gl_wid(), *wb = wa + 2 * GL_WL;
    int na = 0, nb = 0;
    auto push = [&](int *list, int &cnt, bool p, int c) {
        const uint32_t bal = gl_ballot(p);
        if (p) list[cnt + __popc(bal & ((1u << lane) - 1u))] = c;
        cnt += __popc(bal);
        gl_wsync();
    };
    auto drain = [&](bool all) {
        #pragma unroll 1
        while (na >= GL_WL || (all && na > 0)) {
            const int take = na >= GL_WL ? GL_WL : na;
            na -= take;
            refresh_floor();
            const int c = lane < take ? wa[na + lane] : 0;
            const bool need = lane < take && screen(c);
            gl_wsync();
            push(wb, nb, need, c);
            if (nb >= GL_WL) { nb -= GL_WL; evaluate(wb[nb + lane]); gl_wsync(); }
        }
        if (all && nb > 0) { if (lane < nb) evaluate(wb[lane]); nb = 0; gl_wsync(); }
    };
    const long long t_a = g.stat ? gl_clock() : 0;

    // ---- rebuild: everything up to date, dead records dropped, survivors re-binned into the other half
    const int period = g.rebuild > 0 ? g.rebuild : GL_REBUILD;
    if (!C::GL_SIMPLE && ul > 0 && gs.since >= period) {
#ifdef SQRN_HOST_EMU
        g_emu_gl_rebuilds++;
#endif
        GEnt *ent2 = g.ent + (size_t)(gs.half ^ 1) * g.cap;
        double *bps2 = g.bps + (size_t)(gs.half ^ 1) * g.cap;
        uint8_t *qb2 = g.qb + (size_t)(gs.half ^ 1) * g.cap;
        #pragma unroll 1
        for (int q = r; q < GL_NBIN; q += T) S.ghist[q] = 0;
        Team<TW>::sync();
        // chunks are dealt statically (a cluster's CTAs scatter what they counted)
        const int nw = TW > 0 ? TW : 1;
        const int gw = C::CLUSTER ? S.doffset * nw + gl_wid() : gl_wid(), gnw = C::CLUSTER ? S.dstride * nw : nw;
        #pragma unroll 1
        for (int c0 = gw * CH; c0 < n; c0 += gnw * CH) {
            #pragma unroll
            for (int u = 0; u < GL_BATCH; u++) {
                const int c = c0 + u * GL_WL + lane;
                if (c < n) buf[u] = gl_load(&ent[c]); else buf[u].key = GK_DEAD;
            }
            #pragma unroll 1
            for (int u = 0; u < GL_BATCH; u++) {
                const int c = c0 + u * GL_WL + lane;
                GEnt e = buf[u];
                revise(c, e, c >= gs.n_inc && c < gs.n_sorted);
                if (e.key != GK_DEAD) atomicAdd(&S.ghist[qb[c]], 1);
            }
        }
        gl_sync<C>();
        const int n2 = *(volatile int *)np;
        if (n2 > g.cap) { ok = false; return best; }
        #pragma unroll 1
        for (int c0 = n + gw * GL_WL; c0 < n2; c0 += gnw * GL_WL) {          // the pieces these cuts appended
            const int c = c0 + lane;
            if (c < n2) atomicAdd(&S.ghist[qb[c]], 1);
        }
        const int nn = gl_make_bins<C>(S, g);
        #pragma unroll 1
        for (int c0 = gw * CH; c0 < n; c0 += gnw * CH) {
            #pragma unroll
            for (int u = 0; u < GL_BATCH; u++) {
                const int c = c0 + u * GL_WL + lane;
                if (c < n) buf[u] = gl_load(&ent[c]); else buf[u].key = GK_DEAD;
            }
            #pragma unroll 1
            for (int u = 0; u < GL_BATCH; u++) {
                const int c = c0 + u * GL_WL + lane;
                if (buf[u].key == GK_DEAD) continue;
                const int q = qb[c];
                const int slot = atomicAdd(&S.gcur[q], 1);
                gl_store(&ent2[slot], buf[u].key, buf[u].meta, buf[u].v);
                bps2[slot] = bps[c]; qb2[slot] = (uint8_t)q;
            }
        }
        #pragma unroll 1
        for (int c0 = n + gw * GL_WL; c0 < n2; c0 += gnw * GL_WL) {
            const int c = c0 + lane;
            if (c < n2) {
                const GEnt e = gl_load(&ent[c]);
                const int q = qb[c];
                const int slot = atomicAdd(&S.gcur[q], 1);
                gl_store(&ent2[slot], e.key, e.meta, e.v);
                bps2[slot] = bps[c]; qb2[slot] = (uint8_t)q;
            }
        }
        top[r] = -1;
        gl_sync<C>();
        if (gl_leader<C>(S)) *np = nn;
        gs.half ^= 1; gs.n_sorted = nn; gs.n_inc = nn; gs.since = 0;
        ent = ent2; bps = bps2; qb = qb2;
        n = nn;
        ul = 0;                                    // every record has seen T
        if (g.stat && gl_leader<C>(S)) atomicAdd(&g.stat[12], 1ull);
    }
    gs.since++;

    // ---- prologue: the floor from the records that were the threads' bests in the last pass
    if (r == 0) *fl_hi = 0u;
    gl_sync<C>();
    Best fb; fb.fin = -1e300; fb.key = 0xffffffffu; fb.len = 0;
    if (n >= 8 * T || C::CLUSTER) {                // (short lists: a sweep is a few records per thread, a floor buys nothing)
        const int c = top[r];
        if (c >= 0 && c < n) {
            GEnt e = gl_load(&ent[c]);
            if (revise(c, e, false) == GS_EVAL) offer(e.v, e.key, (int)(e.meta & 0xffffu), c);
        }
        fb = team_argmax<C>(S, best);
        if (C::CLUSTER) fb = cluster_best<C>(S, fb, xc++);
        if (fb.fin > floor) floor = fb.fin;        // every thread starts from the team-wide (cluster-wide) floor
    }
    // ---- rounds over the prefix of the binned list that the floor reaches (+ the tail, in the first one)
    const int ns = gs.n_sorted, ntail = n - ns;
    int done = 0;
    bool first_round = true;
    #pragma unroll 1
    for (;;) {
        // as far as the floor known so far asks for, but no further than four times what has been swept (a weak
        // early floor must not drag the whole list in: the next round will know better); whole bins
        int target = (fb.fin > -1e300 && !C::GL_SIMPLE) ? S.gbend[gl_bin(P, fb.fin)] : ns;
        if (!C::GL_SIMPLE) {
            const int want = done > 0 ? 4 * done : (gs.n_inc > 16 * T && gs.n_inc < ns ? gs.n_inc : 16 * T);     // (first round: what the last pass needed)
            if (target > want) {
                int q = 0, qh = GL_NBIN - 1;               // the highest bin q with gbend[q] >= want (gbend falls with q)
                #pragma unroll 1
                while (q < qh) { const int mid = (q + qh + 1) >> 1; if (S.gbend[mid] >= want) q = mid; else qh = mid - 1; }
                if (S.gbend[q] < target) target = S.gbend[q];
            }
        }
        if (target > ns) target = ns;
        if (target <= done && !first_round) break;
        if (target < done) target = done;
        // virtual index space of the round: [done, target) of the bins, then (first round) the tail [ns, n)
        const int span = target - done, vtot = span + (first_round ? ntail : 0);
        if (gl_leader<C>(S)) *cc = 0;
        gl_sync<C>();
        #pragma unroll 1
        for (int v0_ = grab(); v0_ < vtot; v0_ = grab()) {
            #pragma unroll
            for (int u = 0; u < GL_BATCH; u++) {
                const int v = v0_ + u * GL_WL + lane;
                const int c = v < span ? done + v : ns + (v - span);
                if (v < vtot) buf[u] = gl_load(&ent[c]); else buf[u].key = GK_DEAD;
            }
            refresh_floor();
            #pragma unroll 1
            for (int u = 0; u < GL_BATCH; u++) {
                const int v = v0_ + u * GL_WL + lane;
                const int c = v < span ? done + v : ns + (v - span);
                GEnt e = buf[u];
                const uint32_t st = revise(c, e, !C::GL_SIMPLE && v < span && c >= gs.n_inc);
                if (st == GS_EVAL) offer(e.v, e.key, (int)(e.meta & 0xffffu), c);
                push(wa, na, wants_look(st, e), c);
                if (na >= GL_WL) drain(false);
            }
            n_swept += GL_BATCH;
        }
        drain(true);
        done = target; first_round = false;
        fb = team_argmax<C>(S, best);
        if (C::CLUSTER) fb = cluster_best<C>(S, fb, xc++);
        if (fb.fin > floor) floor = fb.fin;
        if (done >= ns) break;
    }
    // what is up to date with every selected stem now: the records this pass revised; the others still are
    // if this pass had no new stem to show them
    gs.n_inc = ul > 0 ? done : (done > gs.n_inc ? done : gs.n_inc);
    // ---- epilogue: the pieces this pass appended
    if (gl_leader<C>(S)) *cc = n;
    gl_sync<C>();
    const long long t_b = g.stat ? gl_clock() : 0;
    const int n2 = *(volatile int *)np;
    if (n2 > g.cap) { ok = false; return best; }
    if (n2 > n) {
        #pragma unroll 1
        for (int c0 = grab(); c0 < n2; c0 = grab()) {
            #pragma unroll 1
            for (int u = 0; u < GL_BATCH; u++) {
                const int c = c0 + u * GL_WL + lane;
                push(wa, na, c < n2, c);
                if (na >= GL_WL) drain(false);
            }
        }
        drain(true);
    }
    top[r] = best_idx;
    if (g.stat) {
        Team<TW>::sync();
        if (r == 0) { const long long t_c = gl_clock(); atomicAdd(&g.stat[6], (unsigned long long)(t_b - t_a)); atomicAdd(&g.stat[7], (unsigned long long)(t_c - t_b)); }
        if (n_eval) atomicAdd(&g.stat[1], (unsigned long long)n_eval);
        if (n_reset) atomicAdd(&g.stat[2], (unsigned long long)n_reset);
        if (n_cut) atomicAdd(&g.stat[5], (unsigned long long)n_cut);
        if (n_swept && lane == 0) atomicAdd(&g.stat[0], (unsigned long long)n_swept * GL_WL);
        if (gl_leader<C>(S)) { atomicAdd(&g.stat[3], 1ull); if (relevel) atomicAdd(&g.stat[4], 1ull); }
    }
    if (n2 > n) {
        best = team_argmax<C>(S, best);
        if (C::CLUSTER) best = cluster_best<C>(S, best, xc++);
        return best;
    }
    return fb;
}

// two stems are "in conflict" when they share a paired position (seq.py:783-786)
__device__ __forceinline__ bool iv_overlap(int a0, int a1, int b0, int b1) { return a0 <= b1 && b0 <= a1; }
__device__ __forceinline__ bool stems_share(int i1, int j1, int l1, int i2, int j2, int l2)
{
    return iv_overlap(i1, i1 + l1 - 1, i2, i2 + l2 - 1) || iv_overlap(i1, i1 + l1 - 1, j2 - l2 + 1, j2) ||
           iv_overlap(j1 - l1 + 1, j1, i2, i2 + l2 - 1) || iv_overlap(j1 - l1 + 1, j1, j2 - l2 + 1, j2);
}

// ChooseStems (seq.py:754-789) after a team_scan(keep_all): every candidate with
// fin >= subopt * best that conflicts with all the better chosen ones, in the
// order of the reference's stable sort.  Writes (i, j, len) triples; returns
// the number chosen, or -1 if the candidate list overflowed (caller retries
// with a larger list).
template <class C>
__device__ int team_choose(State &S, const Layout &L, const Best &best, double subopt,
                           int32_t *out, double *outfin, int cap)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    int ntot = S.misc[7];
    if (ntot > L.Ccap) return -1;
    if (best.fin <= -1e300) return 0;
    const double range = __dmul_rn(subopt, best.fin);
    // rank the in-range candidates by (fin desc, enumeration order)
    int16_t *ord = (int16_t *)S.rkey;      // the run list is dead after the scan: Ocap entries fit in its space
    if (r == 0) S.misc[6] = 0;
    Team<TW>::sync();
    #pragma unroll 1
    for (int c = r; c < ntot; c += T) {
        double f = S.cfin[c];
        uint32_t k = S.ckey[c];
        if (f <= -1e300 || (f < range && !(f == best.fin && k == best.key))) continue;
        int rank = 0;
        #pragma unroll 1
        for (int d = 0; d < ntot; d++) {
            double g = S.cfin[d];
            if (g <= -1e300 || (g < range && !(g == best.fin && S.ckey[d] == best.key))) continue;
            if (better(g, S.ckey[d], f, k)) rank++;
        }
        if (rank < L.Ocap) ord[rank] = (int16_t)c;
        atomicAdd(&S.misc[6], 1);
    }
    Team<TW>::sync();
    int nin = S.misc[6];
    Team<TW>::sync();                      // everyone has read the count before thread 0 reuses the word below
    if (nin > L.Ocap || nin > 32767) return -1;
    if (r == 0) {
        int n = 0;
        #pragma unroll 1
        for (int a = 0; a < nin; a++) {
            int c = ord[a];
            uint32_t k = S.ckey[c];
            int i = k & 0xffff, j = (int)(k >> 16) - i, len = S.clen[c];
            bool ok = true;
            #pragma unroll 1
            for (int q = 0; q < n && ok; q++)
                if (!stems_share(i, j, len, out[3 * q], out[3 * q + 1], out[3 * q + 2])) ok = false;
            if (a == 0) ok = true;
            if (ok) {
                if (n < cap) { out[3 * n] = i; out[3 * n + 1] = j; out[3 * n + 2] = len; if (outfin) outfin[n] = S.cfin[c]; }
                n++;
            }
        }
        S.misc[6] = n;
    }
    Team<TW>::sync();
    return S.misc[6];
}

// AnnotateStems output in reference order (YieldStems, ali.py:86-101): stems of
// the current structure state, written as (i, j, len) + bp score.  Returns the
// number of stems (may exceed cap: the caller then re-allocates and re-runs).
template <class C>
__device__ int team_yield(State &S, const DevParams &P, const DevBatch &B, int32_t *out, double *outsc, int64_t cap)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    const int smax = 2 * S.N - 6;
    int base = 0;
    #pragma unroll 1
    for (int s0 = 4; s0 <= smax; s0 += T) {
        int s = s0 + r, cnt = 0;
        // pass 0 counts the stems of this lane's diagonal, pass 1 writes them at the scanned offset
        int w = 0;
        #pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            if (s <= smax && (pass == 0 || cnt))
                enum_diag<C>(S, P, B, s, [&](int a, int e) {
                    int len = e - a + 1;
                    if ((double)len < P.minlen) return;
                    double sc = run_score<C>(S, P, B, s, a, len);
                    if (!(sc >= P.minbpscore)) return;
                    if (pass == 0) { cnt++; return; }
                    if (w < cap) { out[3 * (int64_t)w] = a; out[3 * (int64_t)w + 1] = s - a; out[3 * (int64_t)w + 2] = len; outsc[w] = sc; }
                    w++;
                });
            if (pass == 0) {
                int tot, off = team_exscan<TW>(S, cnt, tot);
                w = base + off;
                base += tot;
            }
        }
    }
    return base;
}

// round(x, 3) as CPython computes it when x * 1000 is far from a rounding tie
// (correctly rounded decimal -> nearest double = k / 1000 with IEEE division);
// ok = false: the caller must take the exact (host) path.  Same test as
// pyround3() in sqrn_params.h.
__device__ __forceinline__ double round3_fast(double x, bool &ok)
{
    if (!(fabs(x) < 1e9)) { ok = false; return x; }
    // y = x * 1000 carries a relative error <= 2^-53: k is the integer nearest to the exact product
    // whenever y is further than that from the half-way point
    double y = __dmul_rn(x, 1000.0), k = rint(y);
    if (!(fabs(__dsub_rn(y, k)) < __dsub_rn(0.4999999, __dmul_rn(fabs(y), 1e-15)))) { ok = false; return x; }
    return __ddiv_rn(k, 1000.0);
}

// the same as an integer number of thousandths (|k| < 2^31)
__device__ __forceinline__ double round3_milli(double x, bool &ok)
{
    if (!(fabs(x) < 2e6)) { ok = false; return 0.0; }
    double y = __dmul_rn(x, 1000.0), k = rint(y);
    if (!(fabs(__dsub_rn(y, k)) < __dsub_rn(0.4999999, __dmul_rn(fabs(y), 1e-15)))) { ok = false; return 0.0; }
    return k;
}

// ------------------------------------------------------------- finalisation
// ScoreStruct (seq.py:861-899) + dbn of the finished structure.
template <class C>
__device__ void team_finalize(State &S, const DevParams &P, const DevBatch &B, const DevWork &Wk, int item,
                              bool levels_valid = false)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank(), T = Team<TW>::T;
    const int N = S.N;
    if (!levels_valid) team_levels<C>(S);         // stlev[] of the final structure
    int64_t doff = Wk.dbn_off ? Wk.dbn_off[item] : 0;
    if (C::IO != 2 && (Wk.out_dbn_ascii || Wk.out_dbn_code)) {
        const int64_t so = B.off[Wk.item_seq ? Wk.item_seq[item] : item];
        #pragma unroll 1
        for (int p = r; p < N; p += T) {
            int pr = S.partner[p];
            int lv = pr < 0 ? 0 : S.stlev[S.owner[p]];
            if (Wk.out_dbn_code) Wk.out_dbn_code[doff + p] = (int8_t)(pr < 0 ? 0 : (p < pr ? (lv > 127 ? 127 : lv) : -(lv > 127 ? 127 : lv)));
            if (Wk.out_dbn_ascii) {
                uint8_t ch = '.';
                if (S.code[p] == CODE_SEP) ch = B.sym[so + p];
                else if (pr >= 0) {
                    const char *op = "([{<ABCDEFGHIJKLMNOPQRSTUVWXYZ", *cl = ")]}>abcdefghijklmnopqrstuvwxyz";
                    ch = lv <= 30 ? (uint8_t)(p < pr ? op[lv - 1] : cl[lv - 1]) : (uint8_t)'?';
                }
                Wk.out_dbn_ascii[doff + p] = ch;
            }
        }
    }
    const bool compact = C::IO == 2 || (C::IO == 0 && Wk.out_dbn_nib != nullptr);
    if (compact) {
        // two positions per byte; every item starts on a byte of its own ((doff >> 1) + item: at most one byte lost per item)
        uint8_t *nb = Wk.out_dbn_nib + (doff >> 1) + item;
        auto nib_of = [&](int p) -> uint32_t {
            if (p >= N) return 0u;
            const int pr = S.partner[p];
            if (pr < 0) return 0u;
            const int lv = S.stlev[S.owner[p]];
            return (uint32_t)((p < pr ? 0 : 8) | (lv & 7));
        };
#ifndef SQRN_HOST_EMU
        if (TW == 1) {
            #pragma unroll 1
            for (int p0 = 0; p0 < N; p0 += 32) {
                const int p = p0 + r;
                const uint32_t lo = nib_of(p), hi = __shfl_down_sync(0xffffffffu, lo, 1);
                if (!(r & 1) && p < N) nb[p >> 1] = (uint8_t)(lo | (hi << 4));
            }
        } else
#endif
        {
            #pragma unroll 1
            for (int k = r; 2 * k < N; k += T) nb[k] = (uint8_t)(nib_of(2 * k) | (nib_of(2 * k + 1) << 4));
        }
    }
    // bpsum of every stem in units of 0.5 (GU -0.5, AU 1.5, GC 4.0), one stem per thread
    #pragma unroll 1
    for (int t = r; t < S.nst; t += T) {
        int k2 = 0;
        #pragma unroll 1
        for (int q = 0; q < S.stl[t]; q++) {
            int a = S.code[S.sti[t] + q], b = S.code[S.stj[t] - q];
            int lo = a < b ? a : b, hi = a < b ? b : a;
            if (lo == CODE_G && hi == CODE_U) k2 -= 1;
            else if (lo == CODE_A && hi == CODE_U) k2 += 3;
            else if (lo == CODE_C && hi == CODE_G) k2 += 8;
        }
        S.cc[t] = k2;
    }
    Team<TW>::sync();
    if (r == 0) {
        uint8_t flags = 0;
        double thescore = 0.0; bool any = false; int maxlev = 0;
        #pragma unroll 1
        for (int t = 0; t < S.nst; t++) {                // summed in selection order (seq.py:884)
            int k2 = S.cc[t];
            if (k2 > 0) {
                double pw = (k2 < P.pw17_n) ? __ldg(&P.pw17_lut[k2]) : slow_pow(0.5 * k2, 1.7);
                thescore = __dadd_rn(thescore, pw); any = true;
            }
            if (S.stlev[t] > maxlev) maxlev = S.stlev[t];
        }
        double reactscore = 0.5;
        if (!C::PLAIN && S.has_react && !S.default_reacts) {
            int nsep = S.has_sep ? S.sepcnt[N] : 0;
            // builtin sum(): plain left-to-right for numpy floats, Neumaier for exact Python floats
            double sum = 0.0, comp = 0.0; bool first = true;
            #pragma unroll 1
            for (int p = 0; p < N; p++) {
                if (S.code[p] == CODE_SEP) continue;
                double rv = B.rvals[S.rcode[p]];
                double x = S.partner[p] >= 0 ? rv : __dsub_rn(1.0, rv);
                if (first || !B.react_comp) { sum = __dadd_rn(sum, x); first = false; continue; }
                double t = __dadd_rn(sum, x);
                if (fabs(sum) >= fabs(x)) comp = __dadd_rn(comp, __dadd_rn(__dsub_rn(sum, t), x));
                else comp = __dadd_rn(comp, __dadd_rn(__dsub_rn(x, t), sum));
                sum = t;
            }
            if (B.react_comp && comp != 0.0 && isfinite(comp)) sum = __dadd_rn(sum, comp);
            reactscore = __dsub_rn(1.0, __ddiv_rn(sum, (double)(N - nsep)));
        }
        if (!any) flags |= FLAG_INT0;
        if (maxlev > (compact ? 7 : 30)) flags |= FLAG_LEVELS;
        double v0 = __dmul_rn(thescore, reactscore), v1 = thescore, v2 = reactscore;   // seq.py:899, before round()
        if (compact) {
            // round(x, 3) as an integer number of thousandths; next to a rounding tie (or out of range) the host redoes it
            bool ok = true;
            const double k0 = round3_milli(v0, ok), k1 = round3_milli(v1, ok);
            if (ok) { Wk.out_milli[2 * (int64_t)item] = (int32_t)k0; Wk.out_milli[2 * (int64_t)item + 1] = (int32_t)k1; }
            else {
                flags |= FLAG_ROUND;
                const int slot = atomicAdd(Wk.rnd_count, 1);
                if (slot < Wk.rnd_cap) { double *q = Wk.rnd_list + 4 * (int64_t)slot; q[0] = (double)item; q[1] = v0; q[2] = v1; q[3] = v2; }
            }
            if (Wk.out_ns16) Wk.out_ns16[item] = (uint16_t)(S.nst > 65535 ? 65535 : S.nst);
        } else {
            if (Wk.round3) {
                bool ok = true;
                double q0 = round3_fast(v0, ok), q1 = round3_fast(v1, ok), q2 = round3_fast(v2, ok);
                if (ok) { v0 = q0; v1 = q1; v2 = q2; } else flags |= FLAG_ROUND;    // raw values stay: host rounds them
            }
            Wk.out_raw[3 * (int64_t)item] = v0;
            Wk.out_raw[3 * (int64_t)item + 1] = v1;
            Wk.out_raw[3 * (int64_t)item + 2] = v2;
            Wk.out_nstems[item] = S.nst;
        }
        if (Wk.out_flags) Wk.out_flags[item] = flags;
    }
    // stems in selection order
    if (!Wk.out_off) return;
    int64_t so = Wk.out_off[item], cap = Wk.out_off[item + 1] - so;
    #pragma unroll 1
    for (int t = r; t < S.nst && t < cap; t += T) {
        Wk.out_stems[3 * (so + t)] = S.sti[t];
        Wk.out_stems[3 * (so + t) + 1] = S.stj[t];
        Wk.out_stems[3 * (so + t) + 2] = S.stl[t];
    }
}

// ------------------------------------------------ cluster-wide arg-max (DSMEM)
// Every CTA of the cluster publishes the winner of its share of the anti-diagonals in rank 0's
// shared memory (a remote store through distributed shared memory), the cluster synchronises, and
// every CTA reads the CS entries back and picks the same overall winner with the same tie rule.
// Two parity buffers make one cluster barrier per greedy step enough.
struct XchgEntry { double fin; uint32_t key; int len; };

template <class C>
__device__ __forceinline__ Best cluster_best(State &S, Best b, int step)
{
#ifndef SQRN_HOST_EMU
    if (C::CLUSTER) {
        namespace cg = cooperative_groups;
        cg::cluster_group cl = cg::this_cluster();
        const unsigned cr = cl.block_rank(), cs = cl.num_blocks();
        XchgEntry *x0 = (XchgEntry *)cl.map_shared_rank(S.xchg, 0) + (step & 1) * 16;
        if (threadIdx.x == 0) { XchgEntry e; e.fin = b.fin; e.key = b.key; e.len = b.len; x0[cr] = e; }
        cl.sync();
        if (threadIdx.x == 0) {
            Best g; g.fin = -1e300; g.key = 0xffffffffu; g.len = 0;
            for (unsigned q = 0; q < cs; q++) {
                XchgEntry e = x0[q];
                if (better(e.fin, e.key, g.fin, g.key)) { g.fin = e.fin; g.key = e.key; g.len = e.len; }
            }
            S.red[0] = g.fin; ((uint32_t *)(S.red + 1))[0] = g.key; ((uint32_t *)(S.red + 1))[1] = (uint32_t)g.len;
        }
        __syncthreads();
        b.fin = S.red[0]; b.key = ((uint32_t *)(S.red + 1))[0]; b.len = (int)((uint32_t *)(S.red + 1))[1];
        __syncthreads();
    }
#endif
    return b;
}

// ------------------------------------------------------------ one work item
template <class C>
__device__ void team_run_item(State &S, const DevParams &P, const DevBatch &B, const DevWork &Wk,
                              const Layout &L, int item)
{
    constexpr int TW = C::TW;
    const int r = Team<TW>::rank();
    const int seq = Wk.item_seq ? Wk.item_seq[item] : item;
    const int mode = C::MODE >= 0 ? C::MODE : Wk.mode;
    S.region_mode = Wk.region_mode;
    GList g{};
    GState gs{};
    if (C::GLIST) {
#ifdef SQRN_HOST_EMU
        const long long slot = 0;
#else
        const long long slot = C::CLUSTER ? blockIdx.x / (unsigned)S.dstride : blockIdx.x;      // one list per cluster
#endif
        g.ent = (GEnt *)Wk.g_ent + 2 * slot * Wk.g_cap; g.bps = Wk.g_bps + 2 * slot * Wk.g_cap; g.qb = Wk.g_qb + 2 * slot * Wk.g_cap;
        g.cap = (int)Wk.g_cap; g.stat = Wk.g_stat; g.rebuild = Wk.g_rebuild;
        g.cnt = Wk.g_cnt ? Wk.g_cnt + GL_CNT_INTS * slot : nullptr;
    }
    if (C::PERSIST) {
        // hopeless for the list (too long for this parameter set): straight to the rescanning kernel
        const int N = (int)(B.off[seq + 1] - B.off[seq]);
        if (!(C::GLIST ? gl_wanted(N, P, Wk.g_cap) : persist_wanted(N, P, L))) {
            if (r == 0 && (!C::CLUSTER || S.doffset == 0)) { const int slot = atomicAdd(Wk.ovf_count, 1); Wk.ovf_list[slot] = item; }
            return;
        }
    }
    team_load<C>(S, B, P, seq);
    if (Wk.init_off) {
        const int64_t k0 = Wk.init_off[item], k1 = Wk.init_off[item + 1];
        #pragma unroll 1
        const bool light = (C::MODE >= 0 ? C::MODE : Wk.mode) == MODE_FINAL;
        for (int64_t k = k0; k < k1; k++)
            team_apply_stem<C>(S, Wk.init_stems[3 * k], Wk.init_stems[3 * k + 1], Wk.init_stems[3 * k + 2], false, light);
        if (light) Team<TW>::sync();
        if (k1 > k0) team_unpaired_prefix<C>(S);
    }
    unsigned long long calls = 0;
    // the sequence's base list, when the launch has one (pool rounds and their tails, CTA teams)
    BaseView bview{}; const BaseView *basep = nullptr;
    if (TW != 1 && Wk.base_n) {
        if (mode == MODE_BASE) {
            base_build<C>(S, P, B, Wk.base_ent + Wk.base_off[seq], Wk.base_off[seq + 1] - Wk.base_off[seq],
                          Wk.base_bend + (GL_NBIN + 1) * (int64_t)seq, Wk.base_n + seq);
            return;
        }
        const int nb = Wk.base_n[seq];
        if (nb >= 0 && (mode == MODE_STEP || mode == MODE_TAIL) && !C::PERSIST && !C::CLUSTER) {
            const int r_ = Team<TW>::rank();
            #pragma unroll 1
            for (int q = r_; q <= GL_NBIN; q += Team<TW>::T) S.gbend[q] = Wk.base_bend[(GL_NBIN + 1) * (int64_t)seq + q];
            Team<TW>::sync();
            bview.ent = Wk.base_ent + Wk.base_off[seq]; bview.n = nb; basep = &bview;
        }
    }
    if (mode == MODE_BASE) return;
    if (mode == MODE_TAIL || mode == MODE_FINAL) {
        // the pool loop of seq.py:1159-1199 once it can no longer branch
        // (cursize >= poollim => stopper = 1): take the top stem until none is left
        bool ok = true;                             // PERSIST: the run list has not overflowed
        int xc = 0;                                 // cluster flavour of the global list: exchanges done (buffer parity)
        int ui = 0, uj = 0, ul = 0;                 // the stem applied since the last pass over the list
        if (C::PERSIST && mode == MODE_TAIL && (double)S.nst != P.maxstemnum)
        {
            const long long t0 = (C::GLIST && g.stat) ? gl_clock() : 0;
            ok = C::GLIST ? gl_build<C>(S, P, B, g, gs) : persist_build<C>(S, P, B, L);
            if (C::GLIST && g.stat && r == 0) atomicAdd(&g.stat[9], (unsigned long long)(gl_clock() - t0));
        }
        bool lev_ok = false;                        // stlev[] matches the current stem set
        #pragma unroll 1
        while (ok && mode == MODE_TAIL && (double)S.nst != P.maxstemnum) {
            const long long t_it = (C::GLIST && g.stat) ? gl_clock() : 0;
            team_levels<C>(S, ul > 0 ? S.nst - 1 : -1);
            if (C::GLIST && g.stat && r == 0) atomicAdd(&g.stat[10], (unsigned long long)(gl_clock() - t_it));
            lev_ok = true;
            Best b;
            if (C::GLIST) {
                constexpr int T = Team<TW>::T;
                bool relevel = false;
                int ei = -1, ej = -1;
                if (ul > 0) {
                    // did the new stem (index nst - 1) move an older stem to another level?
                    bool ch = false;
                    #pragma unroll 1
                    for (int t = r; t < S.nst - 1; t += T) if (S.stlev[t] != S.stlev2[t]) ch = true;
                    relevel = Team<TW>::any(ch);
                    // its encloser: the older stem with the largest i whose arms lie outside both arms of it
                    if (r == 0) S.misc[10] = -1;
                    Team<TW>::sync();
                    #pragma unroll 1
                    for (int t = r; t < S.nst - 1; t += T) {
                        const int i = S.sti[t], j = S.stj[t], l = S.stl[t];
                        if (i + l - 1 < ui && j - l + 1 > uj) atomicMax(&S.misc[10], (i << 16) | j);      // largest i wins
                    }
                    Team<TW>::sync();
                    if (S.misc[10] >= 0) { ei = S.misc[10] >> 16; ej = S.misc[10] & 0xffff; }
                }
                #pragma unroll 1
                for (int t = r; t < S.nst; t += T) S.stlev2[t] = S.stlev[t];
                Team<TW>::sync();
                if (calls >= 4094) { ok = false; break; }        // the pass stamp of the records has 12 bits: let the rescanning kernel do it
                b = gl_step<C>(S, P, B, g, gs, ui, uj, ul, ei, ej, relevel, ok, xc, (int)calls);
                if (!ok) break;
            } else if (C::PERSIST) {
                b = persist_step<C>(S, P, B, L, ui, uj, ul, ok);
                if (!ok) break;
            } else b = team_scan<C>(S, P, B, L, -1.0, basep);
            if (!C::GLIST) b = cluster_best<C>(S, b, (int)calls);
            calls++;
            if (b.fin <= -1e300) break;
            int i = (int)(b.key & 0xffff);
            ui = i; uj = (int)(b.key >> 16) - i; ul = b.len;
            const long long t_ap = (C::GLIST && g.stat) ? gl_clock() : 0;
            team_apply_stem<C>(S, ui, uj, ul);
            if (C::GLIST && g.stat && r == 0) { atomicAdd(&g.stat[11], (unsigned long long)(gl_clock() - t_ap)); atomicAdd(&g.stat[8], (unsigned long long)(gl_clock() - t_it)); }
            lev_ok = false;
        }
        if (C::PERSIST && !ok) {
            if (r == 0 && (!C::CLUSTER || S.doffset == 0)) { const int slot = atomicAdd(Wk.ovf_count, 1); Wk.ovf_list[slot] = item; }
            Team<TW>::sync();
            return;
        }
        if (C::CLUSTER && S.doffset != 0) calls = 0;          // replicas: rank 0 reports
        else team_finalize<C>(S, P, B, Wk, item, lev_ok);
    } else if (mode == MODE_STEP) {
        int n = 0;
        int64_t so = Wk.out_off[item], cap = Wk.out_off[item + 1] - so;
        if ((double)S.nst != P.maxstemnum) {
            team_levels<C>(S);
            Best b = team_scan<C>(S, P, B, L, Wk.item_subopt[item], basep);
            calls++;
            n = team_choose<C>(S, L, b, Wk.item_subopt[item], Wk.out_stems + 3 * so,
                                Wk.out_stemfin ? Wk.out_stemfin + so : nullptr, (int)cap);
        }
        if (r == 0) Wk.out_nstems[item] = n;
    } else {
        int64_t so = Wk.out_off[item], cap = Wk.out_off[item + 1] - so;
        int n = team_yield<C>(S, P, B, Wk.out_stems + 3 * so, Wk.out_stemfin + so, cap);
        if (r == 0) Wk.out_nstems[item] = n;
    }
    if (r == 0 && Wk.n_calls && calls) atomicAdd(Wk.n_calls, calls);
    Team<TW>::sync();
}

#endif  // __CUDACC__
}  // namespace sqrn
