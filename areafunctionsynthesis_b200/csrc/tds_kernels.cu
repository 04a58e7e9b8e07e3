// tds_kernels.cu -- K1 `tds_batch_kernel`: batched time-domain vocal-tract
// simulation for sm_100a (B200), whole utterances in ONE persistent launch.
//
// Replaces, per audio sample and per tract, the loop body of
// Synthesizer::synthesizeSignalTds (reference Synthesizer.cpp:557-629):
// Tube::interpolate (Tube.cpp:438-505), TriangularGlottis::calcGeometry /
// incTime (TriangularGlottis.cpp:154-397), TdsModel::prepareTimeStep
// (TdsModel.cpp:718-1010), calcMatrix (:1785-2039), the linear solve (:2231-
// 2314), updateVariables (:2046-2098), the radiated-flow sum + skin filter
// (:687-705) and the driver's differentiator + Chebyshev-8 output filter
// (Synthesizer.cpp:614-627).
//
// B200 mapping ("tract per lane, one chain segment per warp"):
//   * one CTA = up to 32 independent tracts; LANE = TRACT everywhere, so every
//     warp instruction does useful work for 32 tracts and every branch on the
//     section type is warp-uniform;
//   * the 97-current network (a tree of chains: trachea-glottis-pharynx,
//     mouth, nose, piriform fossa; sinuses and radiation loads as leaves) is
//     cut into 14 contiguous SEGMENTS, one per warp: 8 "dynamic" warps own six
//     sections of 23..70 each (geometry changes every sample), 6 "static"
//     warps own the trachea, the rigid part of the nose (+ sinuses), the fossa
//     (+ both radiation loads).  A segment warp keeps its recurrent state (p,
//     p', current history c, last solution u) in REGISTERS and does the state
//     update, the network elements, its matrix rows AND the elimination of
//     its segment without leaving registers;
//   * only what crosses a segment boundary goes through SHARED MEMORY in
//     [item][lane] layout (256 B rows, conflict-free): the record of the last
//     section (for the neighbour's first row), the Schur-complement
//     contribution to the separator currents, the separator solutions; plus
//     the soft-wall state (touched once per step) and the control frames;
//   * the linear system (negated: SPD, graph = tree of 3-cliques) is solved by
//     parallel partition: every warp eliminates its interior currents towards
//     its last current (the coupling to the left separator travels along as a
//     "spike"), three warps then solve the reduced system over the 16
//     separators (three chains meeting in the velum clique {40,41,65}), and
//     every warp back-substitutes.  Segments whose sections are all static
//     have CONSTANT pivots: their factors are computed once in the prologue
//     and only the right-hand side is eliminated per step;
//   * two auxiliary warps run the glottis model + output filters and (noise
//     sources on) the noise-source logic concurrently with the solve.
// FP64 throughout.  Algebraically simplified w.r.t. the reference (wall
// admittance alpha = surface * const, one rcp + one sqrt + one rcp per dynamic
// section instead of ~15 div/sqrt, branch-free MUFU+Newton rcp/sqrt, zero-fill
// elimination order), FMA contraction on; parity gate 1e-6 relative max-abs.
#include <cmath>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <cstring>

#include "afs_internal.h"

#ifdef AFS_TDS_TIMING
#define SW_TICK(k) do { const long long now__ = clock64(); sw_t[k] += now__ - sw_last; sw_last = now__; } while (0)
#else
#define SW_TICK(k) do { } while (0)
#endif

namespace afs {

constexpr int LANES = 28;                     // tracts per CTA = row length of the [item][lane] arrays (224 B).  4096
                                              // utterances on 148 SMs are 27.7 per SM; hardware lanes 28..31 duplicate
                                              // lane 27 (same addresses, same values), which keeps 12.5 % of the shared
                                              // memory free for state
constexpr int NW = 16;                        // warps per CTA
constexpr int TDS_THREADS = NW * 32;
constexpr int KD = 6;                         // sections per dynamic-chain warp
constexpr int NSX = NS + 1;                   // + one all-zero dummy section (index 93)
constexpr double kTheta = 0.515;              // TdsModel.cpp:15
constexpr int NZ_RING = 44;                   // pre-generated noise sums per tract: a step draws at most 41 (40 pharynx /
                                              // mouth sources + the lips) and the refill leaves at most 2 entries empty
static_assert(NZ_RING >= 41 + 2 && NZ_RING <= 255, "NZ_RING");
constexpr double kTheta1 = 1.0 - kTheta;

// Warp roles.  The id also fixes the SM sub-partition (id % 4); the order balances the FP64 work of the four.
//   dynamic chains: GP 23..28 (glottis + pharynx), P1 29..34, P2 35..40, M0 41..46, M1 47..52, M2 53..58,
//                   M3 59..64, NV 65..70 (velum port 65..68 + two rigid nose sections)
//   static chains:  T0 0..7, T1 8..15, T2 16..22 (trachea), N1 71..75 (+ sinuses 89, 90), N2 76..83 (+ sinuses
//                   91, 92), F 84..88 (piriform fossa; + the radiation currents 93..96)
//   X1 glottis model + output filters (+ half of the noise-source stage), X2 radiation loads + noise sources
// (the two auxiliary warps come last: the arrays indexed by segment have NSEG = 14 rows)
enum { W_GP = 0, W_P1, W_P2, W_M0, W_M1, W_M2, W_M3, W_NV, W_T0, W_T1, W_N2, W_T2, W_F, W_N1, W_X1, W_X2 };
constexpr int NSEG = 14;
// separators of the partition (last current of every segment + the first currents of the branches)
enum { S7 = 0, S15, S22, S28, S29, S34, S40, S41, S46, S52, S58, S64, S65, S70, S75, S83, NSEP };
// uext[]: solution of the first current of every warp's segment (index = warp id) + radiated flows + zero
enum { UX_RADM = 16, UX_RADN = 17, UX_ZERO = 18, NUX = 19 };

struct TdsParams {
  int n_tracts, n_frames, lanes_per_cta, total_steps, sr_factor;
  int turbulence, soft_walls, skin, fossa, ilc;
  int dump_step;
  int noise;                        // generate_noise_sources (TdsModel.h:87)
  double nz_F, nm_a, nm_1a;         // 1 - exp(-2 pi 40 dt) (TdsModel.cpp:1637), exp(-2 pi 500 dt) and 1 - that (:2052)
  double nz_x2000;                  // exp(-2 pi 2000 dt): low-pass coefficient of a source at the upper cutoff limit
  double dt, idt, th, dtth, dtth1;  // 1/fs, 1/(dt*theta), theta1/theta, dt*theta, dt*theta1
  double ith, fs;                   // theta/theta1, fs
  double fsar, k_ent;
  int transvelar, var_loss;         // transvelarCoupling, glottisLossOption == VARIABLE_ENTRANCE_LOSS (TdsModel.h:75-87)
  int fsrc_sec;                     // TdsModel::setFlowSource (TdsModel.cpp:629-633): section of the flow source, -1 = none
  double fsrc_amp;                  // ... and its strength in cm^3/s
  int n_probes;                     // taps (getSectionPressure / getSectionFlow, TdsModel.cpp:1715-1734) ...
  unsigned long long probe_pack;    // ... and their tube sections, one per byte
  int reserved_[6];                 // (keeps the layout the plain kernels were tuned with)
  double tg_a[5], tg_b[5];          // 50 Hz Chebyshev low-pass of the transglottal pressure (TdsModel.cpp:474)
  double k_sondhi, k_bern;          // 8 rho/(3 pi^2) resp. rho/2, or 0 when the option is off
  double kalpha[3], c1[3], c2[3], c3[3];  // wall constants: 0 standard, 1 trachea, 2 sinus
  double kc[4][3];                        // kalpha*c1, kalpha*c2, kalpha*c3 per wall type; 3 = no walls (zeros)
  double sinus_surface[4];
  double cheb_a[9], cheb_b[9];
  double gs[AFS_NUM_GLOTTIS_STATIC];
  double g_sqrt_mk[2];  // sqrt(mass_i * spring_i): the tension Q cancels (TriangularGlottis.cpp:247-249)
};

// Static section record (same for all tracts).  R1 = resistance seen by the out-flowing current, R0 by the
// in-flowing one (they differ only for the decoupled piriform fossa entrance, TdsModel.cpp:883-889).
struct SRec { double E, L, R1, R0, invr, alpha; };

// Elimination constants of one static chain (computed once in the prologue).
struct STab {
  double T[8], W[8], H[8];     // per interior current: coupling/pivot, 1/pivot, spike/pivot
  double DB, FF, DDA;          // pivot of the last current, coupling left separator <-> last current, Schur term
  double le[2], lg[2], lw[2];  // sinus leaves: couplings e, g and 1/pivot
  double WB, CA;               // fossa: 1/DB and the constant part subtracted from the pivot of current 84
};

// ---- shared memory layout ---------------------------------------------------
struct TdsSmem {
  // per tract, [item][lane]
  double Ed[8 * KD][LANES], Al[8 * KD][LANES];   // E and alpha of the dynamic sections (index 6*warp + slot)
  double wall[NS][3][LANES];                     // wall current, its 1st and 2nd rate (TdsModel.h:184-186)
  double fA[2][AFS_NUM_PM_SECTIONS][LANES], fLen[2][AFS_NUM_PM_SECTIONS][LANES];  // control frames of the interval
  double velum[2][LANES];
  double gctl[2][AFS_NUM_GLOTTIS_CONTROL][LANES];
  double glLen[2][2][LANES], glArea[2][2][LANES];  // glottis geometry, double buffered by step parity
  // boundary record of every segment = its LAST section, for the first row of the next segment
  double bE[NSEG][LANES], bL[NSEG][LANES], bR[NSEG][LANES], bI[NSEG][LANES], bG[NSEG][LANES], bD[NSEG][LANES];
  double csib[4][LANES];                         // history term c of the branch heads 29, 84, 41, 65 (sibling rows)
  double sepD[NSEP][LANES], sepR[NSEP][LANES];
  double redJ[3][2][LANES];                      // pivot and right-hand side of the chain ends 40, 41, 65 (junction of the reduced system)
  double sideDD[NSEG][LANES], sideDR[NSEG][LANES], sideF[NSEG][LANES];
  double lf[5][LANES];                           // fossa current 84 folded into the clique {28, 29}
  double radc[2][2][LANES];                      // radiation pairs folded into currents 64 / 83
  double uext[NUX][LANES];
  double p4[4][LANES];            // pressures of sections 22..25 after the last update
  // non-default options (warp X2 owns the filters): pressures below / above the velum (sections 43, 67), the history
  // part of the two velum filters and of the transglottal-pressure filter (double buffered by step parity), the
  // transvelar flow of the current step
  double tvp[2][LANES], tvh[2][LANES], kenth[2][LANES], tvflow[LANES];
  // noise sources (TdsModel.cpp:1188-1708; warp X2).  Index j = section - 25 for the dipole sources of the
  // pharynx/mouth sections 25..64, 40 = lips; nmag: j = current - 25 for the currents 25..65, 41 = current 84.
  double nmag[42][LANES];                        // noiseMagnitude (TdsModel.h:143) of the currents a constriction can feed
  double namp[41][LANES], ncut[41][LANES], nout[41][LANES];  // currentAmp, cutoff, last filter output (TdsModel.h:101-112)
  // prepared by warp X1 for the vocal-fold constriction of the NEXT step, next to the glottis geometry: sqrt(A) / A^3
  // of the upper glottis area and the downstream factor (see noise_stage)
  double nzk0[2][LANES];
  unsigned long long nzfed[LANES];               // sources the four constriction slots feed (run records: kept by X1)
  unsigned long long nzdecay[LANES];             // decaying sources (nothing feeds them this step) that are still audible
  unsigned long long nztongue[3][LANES];         // sources fed by the tongue / lip constrictions (slots 1..3) that are audible
  double nzlip[LANES];                           // noiseMagnitude of the flow through the lips (currents 93 + 94)
  // glibc random() TYPE_3 (r[i] = r[i-3] + r[i-31]) in a ring of 36 = 3 x 12 words: the twelve draws of one noise
  // sample never wrap inside the ring and three samples bring it back to the same phase (static addresses)
  uint32_t rng[36][LANES];
  uint32_t nzring[LANES];                        // ring of pre-generated sums: head | valid entries << 8 | entries in use << 16 (written by X2)
  uint32_t nzsum[NZ_RING][LANES];                // the next sums of twelve rand() draws, generated ahead by X2
  // The constrictions of the current run of steps (they depend on the geometry only: found by a pre-pass kernel, kept
  // up to date by warp X1): packed {valid, kind, teeth, lateral, nar + 1, up} and the downstream factor as a function
  // of the control-frame ratio r, downF = (c0 + r c1) / (c2 + r c3)
  uint32_t nzc_i[4][LANES];
  double nzc_c[4][4][LANES];
  // same for all tracts
  SRec srec[NSX];
  double wallc[4][3];             // kalpha*c1, kalpha*c2, kalpha*c3 per wall type; 3 = zeros
  double rcn[8];                  // constant factors of the reduced system (reduced_constants)
  STab stab[6];                   // static chains only (stab_slot)
};

__device__ __forceinline__ int stab_slot(int w) { return w == W_F ? 4 : w == W_N1 ? 5 : w - W_T0; }  // T0, T1, N2, T2 = 8..11
__device__ __forceinline__ int wall_type(int s) { return s <= TRACHEA_LAST ? 1 : s >= SINUS0 ? 2 : 0; }
__device__ __forceinline__ bool is_dyn(int s) { return s >= GLOT_LO && s <= 68; }

// first current / section of a warp's segment, its length
__device__ __forceinline__ int seg_first(int w) {
  return w < 8 ? 23 + KD * w : w == W_T0 ? 0 : w == W_T1 ? 8 : w == W_T2 ? 16 : w == W_N1 ? 71 : w == W_N2 ? 76 : FOSSA0;
}
// warp that owns the source section of the segment's first current (-1: none)
__device__ __forceinline__ int seg_prev(int w) {
  switch (w) {
    case W_GP: return W_T2; case W_P1: return W_GP; case W_P2: return W_P1; case W_M0: return W_P2;
    case W_M1: return W_M0; case W_M2: return W_M1; case W_M3: return W_M2; case W_NV: return W_P2;
    case W_T1: return W_T0; case W_T2: return W_T1; case W_N1: return W_NV; case W_N2: return W_N1;
    case W_F: return W_GP; default: return -1;
  }
}
// where the currents flowing OUT of the segment's last section are published
__device__ __forceinline__ int seg_out0(int w) {
  switch (w) {
    case W_GP: return W_P1; case W_P1: return W_P2; case W_P2: return W_M0; case W_M0: return W_M1;
    case W_M1: return W_M2; case W_M2: return W_M3; case W_M3: return UX_RADM; case W_NV: return W_N1;
    case W_T0: return W_T1; case W_T1: return W_T2; case W_T2: return W_GP; case W_N1: return W_N2;
    case W_N2: return UX_RADN; default: return UX_ZERO;
  }
}
// slot in csib[] of a branch head's segment (P1: 29, F: 84, M0: 41, NV: 65), -1 otherwise
__device__ __forceinline__ int csib_slot(int w) { return w == W_P1 ? 0 : w == W_F ? 1 : w == W_M0 ? 2 : w == W_NV ? 3 : -1; }
__device__ __forceinline__ int seg_out1(int w) { return w == W_GP ? W_F : w == W_P2 ? W_NV : UX_ZERO; }
// separator ids: last current of the segment / separator on its left
__device__ __forceinline__ int seg_b(int w) {
  switch (w) {
    case W_GP: return S28; case W_P1: return S34; case W_P2: return S40; case W_M0: return S46;
    case W_M1: return S52; case W_M2: return S58; case W_M3: return S64; case W_NV: return S70;
    case W_T0: return S7; case W_T1: return S15; case W_T2: return S22; case W_N1: return S75;
    default: return S83;  // W_N2
  }
}
__device__ __forceinline__ int seg_a(int w) {
  switch (w) {
    case W_GP: return S22; case W_P1: return S29; case W_P2: return S34; case W_M0: return S41;
    case W_M1: return S46; case W_M2: return S52; case W_M3: return S58; case W_NV: return S65;
    case W_T1: return S7; case W_T2: return S15; case W_N1: return S70; default: return S75;  // W_N2
  }
}

// The per-warp topology constants above, packed into ONE opaque register once per launch (otherwise the compiler
// re-derives them with compare/branch chains in every step): out0:5 | out1:5 | b:4 | a:4 | prev:4 | csib+1:3.
__device__ __forceinline__ unsigned wcfg_make(int w) {
  unsigned v = (unsigned)seg_out0(w) | ((unsigned)seg_out1(w) << 5) | ((unsigned)seg_b(w) << 10) | ((unsigned)seg_a(w) << 14) |
               (((unsigned)seg_prev(w) & 15u) << 18) | ((unsigned)(csib_slot(w) + 1) << 22);
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ int wc_out0(unsigned c) { return (int)(c & 31u); }
__device__ __forceinline__ int wc_out1(unsigned c) { return (int)((c >> 5) & 31u); }
__device__ __forceinline__ int wc_b(unsigned c) { return (int)((c >> 10) & 15u); }
__device__ __forceinline__ int wc_a(unsigned c) { return (int)((c >> 14) & 15u); }
__device__ __forceinline__ int wc_prev(unsigned c) { return (int)((c >> 18) & 15u); }
__device__ __forceinline__ int wc_csib(unsigned c) { return (int)((c >> 22) & 7u) - 1; }

// ---- fast FP64 reciprocal / sqrt (branch-free) -----------------------------------
// IEEE division / sqrt compile to ~20-instruction sequences with a slow-path CALL that fences the
// scheduler; all operands here are positive, normal and well inside the range, so MUFU seed + two
// Newton steps (error <= ~1 ulp) is enough for the 1e-6 parity bar and lets independent sections
// interleave.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}
// One Newton step (relative error ~2e-14 after the 2^-23 seed): enough for the network ELEMENTS, whose rounding only
// perturbs coefficients; the pivots of the solve keep two steps.
__device__ __forceinline__ double fast_rcp1(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return fma(y, fma(-x, y, 1.0), y);
}
__device__ __forceinline__ double fast_rsqrt1(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return fma(y, fma(-0.5 * x * y, y, 0.5), y);
}
__device__ __forceinline__ double fast_div(double a, double b) {
  const double y = fast_rcp(b);
  const double q = a * y;
  return fma(fma(-b, q, a), y, q);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}
__device__ __forceinline__ double fast_sqrt(double x) {  // x > 0
  const double y = fast_rsqrt(x);
  const double sq = x * y;
  return fma(fma(-sq, sq, x), 0.5 * y, sq);
}

// Network elements of a non-sinus section (TdsModel.cpp:759-793).
__device__ __forceinline__ void section_elements(double area, double length, bool glottal, double &L, double &R,
                                                 double &invr, double &circ) {
  const double inva = fast_rcp1(area);
  const double q = area * (1.0 / kPi);
  const double r = q * fast_rsqrt1(q);
  invr = r * kPi * inva;
  circ = 2.0 * kPi * r;  // 2*sqrt(area*pi)
  L = (kRho * 0.5 * length) * inva;
  const double rmin = glottal ? 0.8 : 1.6;
  // elliptical duct with major half-axis rmin when r < rmin (TdsModel.cpp:772-788), else circular (a = b = r):
  // R = 2 mu l (a^2 + b^2) / (pi a^3 b^3) with a b = area / pi
  const double a = fmax(r, rmin), b = r < rmin ? area * (glottal ? 1.0 / (kPi * 0.8) : 1.0 / (kPi * 1.6)) : r;
  const double t = kPi * inva;
  R = (2.0 * kMu * length) * fma(a, a, b * b) * (t * t * inva);  // 1/(pi a^3 b^3) = pi^2 / area^3
}

// beta of TdsModel.cpp:826-831 with Lw = Mw/surface, Rw = Bw/surface folded in:
// beta = kalpha*(c1 wc + c2 wc' + c3 wc''), independent of the geometry.
__device__ __forceinline__ double wall_beta3(const double *w, double wc, double wcr, double wcr2) {
  return fma(wc, w[0], fma(wcr, w[1], wcr2 * w[2]));
}

// Filter for the sound radiated from the skin above the glottis (TdsModel.cpp:494-510, 544).
__device__ __constant__ double c_tone_a[5] = {5.027640021717718e-007, -7.995535578908732e-007,
                                              2.967895557191014e-007, 0.0, 0.0};
__device__ __constant__ double c_tone_b[5] = {0.0, 3.986308869708467, -5.959669638387298, 3.960408461107104,
                                              -0.987047716233603};

// Second velum filter H2 of the transvelar coupling (TdsModel.cpp:512-527); H1 has the coefficients of the skin filter.
__device__ __constant__ double c_tv2_a[5] = {6.589309727087047e-004, -0.001972281980771, 0.001968000742164,
                                             -6.546497341015677e-004, 0.0};

// ---- one matrix row (negated: SPD) of a dynamic segment ---------------------------
// Row of the current flowing from section A into section B (TdsModel.cpp:1916-2034):
//   diag = E_B + E_A + L_AB/(dt theta) + R_AB,   rhs = L_AB c + L_A c_sibling - D_B + D_A
// with c = u/(dt theta) + (theta1/theta) u' the history term of the current (H = -L c, TdsModel.cpp:2020).
// Junction inductance / Bernoulli loss (TdsModel.cpp:1745-1778, 850-877) and the glottal losses (:900-950) are
// added under warp-uniform flags; uB is the previous solution of this current.
enum { RF_JUNCTION = 1, RF_GLO = 2, RF_GUP = 4 };
__device__ __forceinline__ void row_dyn(const TdsParams &P, int fl, double E_A, double L_A, double R_A, double ia,
                                        double D_A, double E_B, double L_B, double R_B, double ib, double D_B, double c,
                                        double c_sib, double uB, double areaA, double areaB, double k_ent, double &gbf,
                                        double &diag, double &rhs) {
  // (branch-free: a flag only selects factors)
  const bool junction = (fl & RF_JUNCTION) != 0, glo = (fl & RF_GLO) != 0, gup = (fl & RF_GUP) != 0;
  const double qa = ia * ia * (1.0 / kPi), qb = ib * ib * (1.0 / kPi);  // 1/area_A, 1/area_B
  // Sondhi junction inductance (TdsModel.cpp:1745-1778); 0 when the option is off
  const double L_AB = fma(junction ? P.k_sondhi : 0.0, fabs(ia - ib), L_A + L_B);
  // Bernoulli loss wide -> narrow (TdsModel.cpp:850-877); area_B < area_A <=> ib > ia
  const bool on = junction && (((ib > ia) && (uB > 0)) || ((ib < ia) && (uB < 0)));
  const double kb = on ? uB * P.k_bern : 0.0;  // u rho/2, 0 when the option is off
  // glottal entrance loss (TdsModel.cpp:900-920) on R_B / lower -> upper glottis section (:930-950) on R_A
  const double target = (areaB < P.fsar * areaA) ? 1.0 : 0.0;  // compared on the areas themselves, like the reference:
  gbf = gup ? 0.8 * gbf + (1.0 - 0.8) * target : gbf;          // the two glottal areas can be equal to the last bit
  const double up = uB > 0 ? uB * (0.5 * kRho) : 0.0;
  const double dq = qb * qb - qa * qa;                          // 1/area_T^2 - 1/area_S^2
  R_A = fma(-kb, qa * qa, fma(gup ? gbf * up : 0.0, dq, R_A));
  R_B = fma(kb, qb * qb, fma(glo ? k_ent * up : 0.0, dq, R_B));
  diag = E_B + E_A + fma(L_AB, P.idt, R_A + R_B);
  rhs = fma(L_AB, c, fma(L_A, c_sib, D_A - D_B));
}
__device__ __forceinline__ int row_flags(int sec) {
  if (sec == GLOT_LO) return RF_GLO;
  if (sec == GLOT_UP) return RF_GUP;
  if (sec > PHAR0 && sec <= MOUTH_LAST && sec != 29 && sec != MOUTH0) return RF_JUNCTION;
  return 0;
}

// ==================================================================================
// Dynamic chain warp
// ==================================================================================
// Registers of a dynamic segment.  x: last solution u -> rhs -> G = rhs'/pivot -> new solution;
// y: diag -> T = coupling/pivot; h: spike/pivot.
struct DRegs {
  double p[KD], pr[KD], c[KD], x[KD], y[KD], h[KD];
  double L0, R0, I0, D0;  // record of the first section (its row is assembled after the barrier)
  double gbf;             // glottalBernoulliFactor (TdsModel.h:241), used by the owner of section 24
};

struct StepCtx {
  int n;            // loop iteration: update with the solution of step n-1, rows of step n
  bool upd, rows;
  double ratio, ratio1;
  bool dumping;
  // taps (OPT kernels only): where this lane's values of step n-1 go, null when the lane is idle or nothing is tapped
  double *tap_p, *tap_in, *tap_out;
  size_t tap_stride;  // distance between two probes of one tract (= total_steps)
};

// getSectionPressure / getSectionFlow of section `sec` after a step (TdsModel.cpp:1715-1734): pressure in the centre of
// the section, volume velocity into and out of it.
// (not inlined: taps are rare, and code that is not executed must stay out of the instruction stream of the time loop
//  -- the kernel lives off its instruction cache)
// Called from run_role around pass A, for the sections of the warp's segment that are tapped (taps_of): the flows are
// the solution of the last step, still in the registers when pass A begins; the pressure is there when it ends.  (The
// calls used to sit in the slot code of pass A: sixteen call sites in the hottest code of the kernel.)
// (everything by value -- the probe sections packed one per byte: a pointer into the kernel parameters or to the step
//  context handed to a function that is not inlined would move BOTH to local memory for the whole kernel; the
//  instantiations with taps ran 1.43 x slower than the plain ones for that reason)
__device__ __noinline__ void tap_section(const int n_probes, const unsigned long long probe_pack, double *tap_p, double *tap_in,
                                         double *tap_out, const size_t tap_stride, int sec, double p, double in, double out) {
  for (int q = 0; q < n_probes; q++) {
    if ((int)((probe_pack >> (8 * q)) & 255ull) != sec) continue;
    if (tap_p) tap_p[q * tap_stride] = p;
    if (tap_in) tap_in[q * tap_stride] = in;
    if (tap_out) tap_out[q * tap_stride] = out;
  }
}

__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }
// Named barriers of the solve phase.  Warp X1 (glottis model, output filters) only has to be back for the next pass A:
// it ARRIVES at barrier 1 once it has read what the others will overwrite and takes no part in barrier 2.
__device__ __forceinline__ void bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// noise sources (warps X2 and X1), see noise_scan / noise_stage below
struct NoiseCons {
  int valid, nar, kind, up, teeth, lateral;  // nar: narrowest section as pharynx/mouth index, -1 = upper glottis
  int branch;                                // how the obstacle moves with the control-frame ratio (noise_scan)
  double downF;                              // kind: 0 vocal folds, 1 tongue, 2 lower lip
};
struct NoiseRegs {
  int ph;                 // position in the random() ring: 0, 12 or 24
  int head, free, cap;    // ring of pre-generated sums: oldest entry; entries that wait for a new sum; entries in use
  double lips;            // sample of the lips dipole source (enters all four radiation rows)
  double nm93, nm94;      // noiseMagnitude of the mouth radiation currents
  double asp, gain;       // aspiration strength (dB) the cached vocal-fold source gain belongs to
};

// One run of steps with the same constrictions (tds_noise_runs_kernel): first step, first step of the tract's next run
// (0x7fffffff: none), the packed constriction slots and the coefficients of downF(r) = (c0 + r c1) / (c2 + r c3).
// The records of a tract form a linked list through one pool: the first record of piece p of tract t has the fixed
// place t * n_pieces + p, the others are taken from the rest of the pool as the bisection finds them.
struct NzRun {
  uint32_t start, next;  // first step of this run / of the tract's next run
  uint32_t slot[4];
  uint32_t next_idx;     // where the next run's record is
  uint32_t pad;
  double coef[4][4];
};
static_assert(sizeof(NzRun) == 160, "NzRun layout");

__device__ __noinline__ void dump_section(const double idt, const double ith, double *dump, int sec, double p, double pr, double wc,
                                          double u, double c) {
  dump[sec] = p; dump[93 + sec] = pr; dump[186 + sec] = wc;
  dump[279 + sec] = u; dump[376 + sec] = (c - u * idt) * ith;
}

// Pass A: state update with the solution of the previous step (TdsModel.cpp:2046-2098), then elements, E, D
// (:718-1010) and the rows of all currents but the first (:1785-2039) of the new step.
// HOT = true is the loop body of all steps but the first and the last: update and rows both on, and the slot body is
// BRANCH-FREE (the kind of a section only selects values), so that the update chain and the geometry chain of a slot
// -- and neighbouring slots -- overlap: the kernel is bound by the latency of dependent FP64 operations, not by the
// pipe.  HOT = false (first / last step, state dumps) takes the flags from X.
template <bool HOT, bool OPT>
__device__ __forceinline__ void dwarp_pass_a(const TdsParams &P, TdsSmem &S, const int lane, const int w, const unsigned wc_,
                                             DRegs &st, const StepCtx &X) {
  const bool upd = HOT ? true : X.upd, rows = HOT ? true : X.rows;
  const int n0 = 23 + KD * w;
  double outLast = 0.0;
  if (upd) outLast = S.uext[wc_out0(wc_)][lane] + S.uext[wc_out1(wc_)][lane];
  double pE = 0.0, pL = 0.0, pR = 0.0, pI = 0.0, pD = 0.0, pA = 0.0;  // record of the previous slot
  const int par = X.n & 1;
  double vel = 0.0;
  if (rows) vel = X.ratio1 * S.velum[0][lane] + X.ratio * S.velum[1][lane];
#pragma unroll
  for (int k = 0; k < KD; k++) {
    double p_ = st.p[k], pr_ = st.pr[k], c_ = st.c[k];
    const double u_ = st.x[k], un_ = (k < KD - 1) ? st.x[k + 1] : outLast;
    double x_ = u_, y_ = 0.0;
    const int sec = n0 + k, rec = KD * w + k;
    // kind of the section (warp-uniform): glottis only in slots 0, 1 (warp GP), velum port only in slots 0..3 (warp NV)
    const bool glottal = k < 2 && sec <= GLOT_UP;
    const bool velum = k < 4 && sec >= NOSE0;
    const bool rigid = k >= 4 && sec > MOUTH_LAST;
    // soft walls: none in the glottis (zero constants keep its wall state at 0; also when the option is off)
    const double ka = glottal ? 0.0 : P.kalpha[0];
    const double kc0 = glottal ? 0.0 : P.kc[0][0], kc1 = glottal ? 0.0 : P.kc[0][1], kc2 = glottal ? 0.0 : P.kc[0][2];
    double wc = S.wall[sec][0][lane], wcr = S.wall[sec][1][lane], wcr2 = S.wall[sec][2][lane];
    if (upd) {
      const double u = u_, out = un_;
      const double E_old = S.Ed[rec][lane], al_old = S.Al[rec][lane];
      const double beta_old = fma(wc, kc0, fma(wcr, kc1, wcr2 * kc2));
      double D_old = fma(-E_old, beta_old, fma(P.dtth1, pr_, p_));
      if (OPT && k == 2 && P.transvelar && w == W_NV) D_old = fma(E_old, S.tvflow[lane], D_old);  // source term of section 67
      if (OPT && P.fsrc_sec == sec) D_old = fma(E_old, P.fsrc_amp, D_old);                         // TdsModel.cpp:999-1002
      // current: u' = (u_new - u_old)/(dt theta) - (theta1/theta) u'_old = u_new/(dt theta) - c
      const double ur = fma(u, P.idt, -c_);
      c_ = fma(ur, P.th, u * P.idt);
      const double p_new = fma(E_old, u - out, D_old);
      const double pr_new = (p_new - p_) * P.idt - pr_ * P.th;
      p_ = p_new; pr_ = pr_new;
      if (OPT && k == 2 && P.transvelar && (w == W_M0 || w == W_NV)) S.tvp[w == W_NV ? 1 : 0][lane] = p_new;  // sections 43, 67
      const double wc_new = fma(pr_new, al_old, beta_old);
      const double wcr_new = (wc_new - wc) * P.idt - wcr * P.th;
      const double wcr2_new = (wcr_new - wcr) * P.idt - wcr2 * P.th;
      wc = wc_new; wcr = wcr_new; wcr2 = wcr2_new;
      S.wall[sec][0][lane] = wc; S.wall[sec][1][lane] = wcr; S.wall[sec][2][lane] = wcr2;
      if (k < 3 && sec <= PHAR0) S.p4[sec - TRACHEA_LAST][lane] = p_new;  // sections 23, 24, 25
    }
    if (rows) {
      // Tube::interpolate + setPharynxMouthGeometry (Tube.cpp:449-466, 323-350); the index is clamped for the other kinds
      const int j = min(max(sec - PHAR0, 0), AFS_NUM_PM_SECTIONS - 1);
      double area = X.ratio1 * S.fA[0][j][lane] + X.ratio * S.fA[1][j][lane];
      double length = X.ratio1 * S.fLen[0][j][lane] + X.ratio * S.fLen[1][j][lane];
      if (k < 2) {  // glottis sections from the glottis model (Synthesizer.cpp:576-579)
        const double ga = S.glArea[par][k][lane], gl = S.glLen[par][k][lane];
        area = glottal ? ga : area; length = glottal ? gl : length;
      }
      if (k < 4) {  // velum port: Tube::setVelumOpening (Tube.cpp:402-416)
        const double va = vel + ((double)(k * k) * (4.24 - vel)) / 16.0;
        area = velum ? va : area; length = velum ? (11.4 / 11.4) * 0.6 : length;
      } else {      // rigid nose sections 69, 70 (Tube.cpp:155-157)
        area = rigid ? (k == 4 ? 4.24 : 3.26) : area; length = rigid ? (11.4 / 11.4) * 0.6 : length;
      }
      area = fmax(area, kMinArea);
      double L, R, invr, circ;
      section_elements(area, length, glottal, L, R, invr, circ);
      const double C = (area * length) * (1.0 / (kRho * kC * kC));
      const double alpha = fmax(circ * length, kMinArea) * ka;  // TdsModel.cpp:805-832
      const double E = P.dtth * fast_rcp1(C + alpha);
      const double beta = fma(wc, kc0, fma(wcr, kc1, wcr2 * kc2));
      double D = fma(-E, beta, fma(P.dtth1, pr_, p_));  // TdsModel.cpp:988-1008
      if (OPT && P.fsrc_sec == sec) D = fma(E, P.fsrc_amp, D);
      S.Ed[rec][lane] = E; S.Al[rec][lane] = alpha;
      if (k == 0) {
        st.L0 = L; st.R0 = R; st.I0 = invr; st.D0 = D;
        if (wc_csib(wc_) >= 0) S.csib[wc_csib(wc_)][lane] = c_;
      } else {
        // (the lower -> upper glottis row can only be slot 1, the entrance row only slot 0: keeps their code and the
        //  registers they need out of the other slots)
        const int fl = row_flags(sec) & (k == 1 ? (RF_JUNCTION | RF_GUP) : RF_JUNCTION);
        row_dyn(P, fl, pE, pL, pR, pI, pD, E, L, R, invr, D, c_, 0.0, u_, k == 1 ? pA : 0.0, area, 0.0, st.gbf, y_, x_);
      }
      if (k == KD - 1) {
        S.bE[w][lane] = E; S.bL[w][lane] = L; S.bR[w][lane] = R; S.bI[w][lane] = invr; S.bD[w][lane] = D;
        S.bG[w][lane] = fma(L, P.idt, R);
      }
      pE = E; pL = L; pR = R; pI = invr; pD = D;
      if (k == 0) pA = area;
    }
    st.p[k] = p_; st.pr[k] = pr_; st.c[k] = c_; st.x[k] = x_; st.y[k] = y_;
  }
}

// TdsModel::getGlottalEntranceLossCoeffFlucher2011 (TdsModel.cpp:1019-1092): entrance loss coefficient from the
// low-pass filtered transglottal pressure and the width of the lower glottis section.
__device__ __noinline__ double fulcher_k_ent(double pressure_dPa, double area_lower) {
  double p = pressure_dPa / 979.7, D = area_lower / 1.25;
  if (p < 0.001) p = 0.001;
  if (D < 0.001) D = 0.001;
  const double logD = log10(D);
  const double a = pow(10.0, 0.7953 * logD * logD + 1.4741 * logD + 0.6529);
  const double b = -0.7427 * logD * logD + -1.6209 * logD + -0.875;
  double k = a * pow(p, b);
  if (k < 0.6) k = 0.6;
  if (k > 18.0) k = 18.0;
  return k;
}

__device__ __forceinline__ void dwarp_noise(const TdsParams &P, TdsSmem &S, const int lane, const int w, DRegs &st);

// After the barrier: the first row (its source section belongs to another warp), then elimination of the
// interior currents towards the last one.  `aown`: the first current is itself a separator (branch heads 29, 41, 65).
template <bool NOISE, bool OPT>
__device__ __forceinline__ void dwarp_eliminate(const TdsParams &P, TdsSmem &S, const int lane, const int w, const unsigned wc_,
                                                DRegs &st, const int par) {
  const int pw = wc_prev(wc_), n0 = 23 + KD * w;
  const bool aown = (w == W_P1) || (w == W_M0) || (w == W_NV);
  const double E_A = S.bE[pw][lane], E_0 = S.Ed[KD * w][lane];
  {
    const double c_sib = wc_csib(wc_) >= 0 ? S.csib[wc_csib(wc_) ^ 1][lane] : 0.0;  // 29 <-> 84, 41 <-> 65
    double k_ent = P.k_ent;
    if (OPT && P.var_loss && w == W_GP) {  // (the filter's history part comes from warp X2, its newest input is p22 - p25)
      const double filt = fma(P.tg_a[0], S.p4[0][lane] - S.p4[3][lane], S.kenth[par][lane]);
      k_ent = fulcher_k_ent(filt, fmax(S.glArea[par][0][lane], kMinArea));
    }
    double diag, rhs;
    row_dyn(P, row_flags(n0), E_A, S.bL[pw][lane], S.bR[pw][lane], S.bI[pw][lane], S.bD[pw][lane], E_0, st.L0, st.R0,
            st.I0, st.D0, st.c[0], c_sib, st.x[0], 0.0, 0.0, k_ent, st.gbf, diag, rhs);
    st.y[0] = diag; st.x[0] = rhs;
  }
  if (OPT && P.transvelar && w == W_NV) {
    // extra flow into the nasal cavity through the velum (TdsModel.cpp:966-980, 996-999): a source in section 67
    // (slot 2), i.e. D_67 += E_67 * flow, which enters the rows of the currents 67 and 68
    const double tv = fma(c_tone_a[0], S.tvp[0][lane], fma(c_tv2_a[0], st.p[2], S.tvh[par][lane]));
    const double dD = S.Ed[KD * w + 2][lane] * tv;
    st.x[2] -= dD; st.x[3] += dD;
    S.tvflow[lane] = tv;
  }
  if (NOISE) dwarp_noise(P, S, lane, w, st);  // dipole sources of the pharynx/mouth sections
  double f = aown ? -E_0 : -E_A;
  double dc = aown ? st.y[1] : st.y[0], rc = aown ? st.x[1] : st.x[0];
  double dda = 0.0, dra = 0.0;
#pragma unroll
  for (int k = 0; k < KD - 1; k++) {
    if (k == 0 && aown) continue;
    const double o = -S.Ed[KD * w + k][lane];
    const double dn = st.y[k + 1], rn = st.x[k + 1];
    const double wv = fast_rcp(dc);
    const double t = o * wv, h = f * wv;
    st.y[k] = t; st.x[k] = rc * wv; st.h[k] = h;
    dda = fma(h, f, dda); dra = fma(h, rc, dra);
    dc = fma(-t, o, dn);
    rc = fma(-t, rc, rn);
    f = -t * f;
  }
  const int b = wc_b(wc_), a = wc_a(wc_);
  S.sepD[b][lane] = dc; S.sepR[b][lane] = rc;
  S.sideF[w][lane] = f;
  if (aown) { S.sepD[a][lane] = st.y[0] - dda; S.sepR[a][lane] = st.x[0] - dra; }
  else { S.sideDD[w][lane] = dda; S.sideDR[w][lane] = dra; }
}

template <bool NOISE>
__device__ __forceinline__ void dwarp_back(const TdsParams &P, TdsSmem &S, const int lane, const int w, const unsigned wc_,
                                           DRegs &st) {
  const bool aown = (w == W_P1) || (w == W_M0) || (w == W_NV);
  const double ub = S.sepR[wc_b(wc_)][lane], ua = S.sepR[wc_a(wc_)][lane];
  st.x[KD - 1] = ub;
  double un = ub;
#pragma unroll
  for (int k = KD - 2; k >= 0; k--) {
    if (k == 0 && aown) { st.x[0] = ua; continue; }
    un = fma(-st.y[k], un, fma(-st.h[k], ua, st.x[k]));
    st.x[k] = un;
  }
  S.uext[w][lane] = st.x[0];
  if (NOISE) {  // noiseMagnitude of the currents 25..65 (TdsModel.cpp:2052-2060)
#pragma unroll
    for (int k = 0; k < KD; k++) {
      const int j = 23 + KD * w + k - PHAR0;
      if (j >= 0 && j <= 40) S.nmag[j][lane] = fma(P.nm_a, S.nmag[j][lane], P.nm_1a * st.x[k]);
    }
  }
}

// ==================================================================================
// Static chain warp: ONE instantiation shared by T0, T1, T2, N1, N2 and F (the hot loop has to stay small: the
// instruction cache is the scarce resource once 16 warps run different code).  Up to KS chain sections, right-
// aligned in the slot frame (valid slots ks..KS-1, slot k = section base + k), optionally two sinus leaves hanging
// on the chain slots j0 and j0+1; `fossa`: the piriform fossa 84..88, whose first current 84 is a branch head with a
// dynamic row and is folded into the clique {28, 29} locally.
// ==================================================================================
constexpr int KS = 8;
struct SCfg {
  int base, ks, leaf0;  // section of slot 0 (may be virtual), first valid slot (0, 1 or 3), first leaf section
  bool leaves, n2;      // sinus leaves present; they hang on slots 0, 1 (N2) instead of 5, 6 (N1)
  bool fossa, head;     // head: the segment starts at the lungs (no source section)
};
__device__ __forceinline__ SCfg scfg_of(int w) {
  SCfg c;
  c.ks = w == W_T2 ? 1 : (w == W_N1 || w == W_F) ? 3 : 0;
  c.base = (w == W_T0 ? 0 : w == W_T1 ? 8 : w == W_T2 ? 16 : w == W_N1 ? 71 : w == W_N2 ? 76 : FOSSA0) - c.ks;
  c.leaves = w == W_N1 || w == W_N2;
  c.n2 = w != W_N1;  // (warps without leaves: zero leaf constants, any position)
  c.leaf0 = w == W_N1 ? 89 : 91;
  c.fossa = w == W_F;
  c.head = w == W_T0;
  // (opaque to the optimiser: computed once per launch instead of once per step)
  asm volatile("" : "+r"(c.base), "+r"(c.ks), "+r"(c.leaf0));
  int fl = (c.leaves ? 1 : 0) | (c.n2 ? 2 : 0) | (c.fossa ? 4 : 0) | (c.head ? 8 : 0);
  asm volatile("" : "+r"(fl));
  c.leaves = fl & 1; c.n2 = fl & 2; c.fossa = fl & 4; c.head = fl & 8;
  return c;
}
struct SRegs {
  double p[KS], pr[KS], c[KS], x[KS];
  double lp[2], lpr[2], lc[2], lx[2];  // sinus leaves
  double D0;                           // D of the first section (its row is assembled after the barrier)
  double w84, e28, g28;                // fossa: factors kept between elimination and back substitution
};
// value of the first valid slot (ks is 0, 1 or 3)
__device__ __forceinline__ double first_of(const SCfg &c, const double (&a)[KS]) { return c.ks == 0 ? a[0] : c.ks == 1 ? a[1] : a[3]; }
__device__ __forceinline__ void set_first(const SCfg &c, double (&a)[KS], double v) {
  a[0] = c.ks == 0 ? v : a[0]; a[1] = c.ks == 1 ? v : a[1]; a[3] = c.ks == 3 ? v : a[3];
}

__device__ __forceinline__ double srec_diag(const SRec &A, const SRec &B, double idt) {
  return B.E + A.E + (A.L + B.L) * idt + A.R1 + B.R0;
}

// Prologue: pivots of the static rows (lane-uniform arithmetic, every lane computes the same values).  Slots that
// are not part of the chain keep T = W = H = 0, which makes the branch-free recurrences below skip them.
__device__ __noinline__ void swarp_build(const TdsParams &P, TdsSmem &S, const int lane, const int w) {
  const SCfg c = scfg_of(w);
  const int n0 = c.base + c.ks, j0 = c.n2 ? 0 : 5;
  const SRec zero{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double d[KS], o[KS];
  STab T;
  memset(&T, 0, sizeof T);
  for (int k = c.ks; k < KS; k++) {
    const int sec = c.base + k;
    const SRec B = S.srec[sec];
    const SRec A = sec == 0 ? zero : S.srec[sec - 1];
    d[k] = srec_diag(A, B, P.idt);  // (fossa: d[ks] is not used, the row of current 84 is dynamic)
    o[k] = -B.E;
  }
  if (c.leaves) {
    for (int j = 0; j < 2; j++) {
      const SRec A = S.srec[c.base + j0 + j], B = S.srec[c.leaf0 + j];
      const double dl = srec_diag(A, B, P.idt);
      const double e = A.E, g = e + A.L * P.idt + A.R1, wl = 1.0 / dl;
      d[j0 + j] -= e * e * wl;
      d[j0 + j + 1] -= g * g * wl;
      o[j0 + j] = -e + e * g * wl;
      T.le[j] = e; T.lg[j] = g; T.lw[j] = wl;
    }
  }
  const int k0 = c.fossa ? c.ks + 1 : c.ks;
  double f = c.head ? 0.0 : c.fossa ? o[c.ks] : -S.srec[n0 - 1].E;
  double dc = d[k0], dda = 0.0;
  for (int k = k0; k < KS - 1; k++) {
    const double wv = 1.0 / dc;
    const double t = o[k] * wv, h = f * wv;
    T.T[k] = t; T.W[k] = wv; T.H[k] = h;
    dda += h * f;
    dc = d[k + 1] - t * o[k];
    f = -t * f;
  }
  T.DB = dc; T.FF = f; T.DDA = dda;
  if (c.fossa) { T.WB = 1.0 / dc; T.CA = dda + f * f * T.WB; }
  if (lane == 0) S.stab[stab_slot(w)] = T;
  // boundary record of the last section and the constant separator data
  const SRec last = S.srec[c.base + KS - 1];
  S.bE[w][lane] = last.E; S.bL[w][lane] = last.L; S.bR[w][lane] = last.R1; S.bI[w][lane] = last.invr;
  S.bG[w][lane] = last.L * P.idt + last.R1; S.bD[w][lane] = 0.0;
  if (!c.fossa) {
    S.sepD[seg_b(w)][lane] = dc;
    S.sideDD[w][lane] = dda; S.sideF[w][lane] = f; S.sideDR[w][lane] = 0.0;
  }
}

// One static section: update with the last solution (TdsModel.cpp:2046-2098), then D of the new step (:988-1008).
template <bool OPT>
__device__ __forceinline__ void static_section(const TdsParams &P, TdsSmem &S, const int lane, const StepCtx &X, int sec,
                                               double &p, double &pr, double &c, double u, double out, double &Dnew) {
  const SRec &r = S.srec[sec];
  const double *wk = S.wallc[wall_type(sec)];
  double wc = S.wall[sec][0][lane], wcr = S.wall[sec][1][lane], wcr2 = S.wall[sec][2][lane];
  const double E = r.E;
  if (X.upd) {
    const double beta_old = wall_beta3(wk, wc, wcr, wcr2);
    double D_old = fma(-E, beta_old, fma(P.dtth1, pr, p));
    if (OPT && P.fsrc_sec == sec) D_old = fma(E, P.fsrc_amp, D_old);  // TdsModel.cpp:999-1002
    const double ur = fma(u, P.idt, -c);
    c = fma(ur, P.th, u * P.idt);
    const double p_new = fma(E, u - out, D_old);
    const double pr_new = (p_new - p) * P.idt - pr * P.th;
    const double wc_new = fma(pr_new, r.alpha, beta_old);
    const double wcr_new = (wc_new - wc) * P.idt - wcr * P.th;
    const double wcr2_new = (wcr_new - wcr) * P.idt - wcr2 * P.th;
    p = p_new; pr = pr_new; wc = wc_new; wcr = wcr_new; wcr2 = wcr2_new;
    S.wall[sec][0][lane] = wc; S.wall[sec][1][lane] = wcr; S.wall[sec][2][lane] = wcr2;
    if (sec == TRACHEA_LAST) S.p4[0][lane] = p_new;
  }
  Dnew = fma(-E, wall_beta3(wk, wc, wcr, wcr2), fma(P.dtth1, pr, p));
  if (OPT && P.fsrc_sec == sec) Dnew = fma(E, P.fsrc_amp, Dnew);
}

template <bool OPT>
__device__ __forceinline__ void swarp_pass_a(const TdsParams &P, TdsSmem &S, const int lane, const int w, const unsigned wc_,
                                             const SCfg &c, SRegs &st, const StepCtx &X) {
  double outLast = 0.0;
  if (X.upd) outLast = S.uext[wc_out0(wc_)][lane];
  const int j0 = c.n2 ? 0 : 5;
  double dn[KS], dl[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < KS; k++) {
    dn[k] = 0.0;
    if (k < c.ks) continue;
    double out = (k < KS - 1) ? st.x[k + 1] : outLast;
    if (k == j0) out += st.lx[0];
    if (k == j0 + 1) out += st.lx[1];
    static_section<OPT>(P, S, lane, X, c.base + k, st.p[k], st.pr[k], st.c[k], st.x[k], out, dn[k]);
  }
  if (c.leaves) {
#pragma unroll
    for (int j = 0; j < 2; j++) static_section<OPT>(P, S, lane, X, c.leaf0 + j, st.lp[j], st.lpr[j], st.lc[j], st.lx[j], 0.0, dl[j]);
  }
  if (X.rows) {
    // rows of the static currents: constant pivots, only the right-hand side is assembled (the rows of the slots
    // up to ks are placeholders: slot ks is rewritten after the barrier, the others are not part of the chain)
    st.D0 = first_of(c, dn);
#pragma unroll
    for (int k = KS - 1; k >= 1; k--) {
      const double LA = S.srec[c.base + k - 1].L, LAB = LA + S.srec[c.base + k].L;
      double rhs = fma(LAB, st.c[k], dn[k - 1] - dn[k]);
      if (k == j0 + 1) rhs = fma(LA, st.lc[0], rhs);
      if (k == j0 + 2) rhs = fma(LA, st.lc[1], rhs);
      st.x[k] = k > c.ks ? rhs : 0.0;
    }
    if (c.leaves) {
      const double cj0 = c.n2 ? st.c[1] : st.c[6], cj1 = c.n2 ? st.c[2] : st.c[7];
      const double dj0 = c.n2 ? dn[0] : dn[5], dj1 = c.n2 ? dn[1] : dn[6];
      const double LA0 = S.srec[c.base + j0].L, LA1 = S.srec[c.base + j0 + 1].L;
      st.lx[0] = fma(LA0 + S.srec[c.leaf0].L, st.lc[0], fma(LA0, cj0, dj0 - dl[0]));
      st.lx[1] = fma(LA1 + S.srec[c.leaf0 + 1].L, st.lc[1], fma(LA1, cj1, dj1 - dl[1]));
    }
    S.bD[w][lane] = dn[KS - 1];
    if (c.fossa) S.csib[1][lane] = st.c[3];
  }
}

__device__ __forceinline__ void swarp_eliminate(const TdsParams &P, TdsSmem &S, const int lane, const int w, const unsigned wc_,
                                                const SCfg &c, SRegs &st, double psrc
#ifdef AFS_TDS_TIMING
                                                , long long *sw_t
#endif
                                                ) {
#ifdef AFS_TDS_TIMING
  long long sw_last = clock64();
#endif
  const int n0 = c.base + c.ks, pw = wc_prev(wc_);
  const STab &C = S.stab[stab_slot(w)];
  double d84 = 0.0, r0;
  {  // first row: the source section belongs to another warp (or is the lung pressure source)
    const double c0 = first_of(c, st.c);
    if (c.head) {
      r0 = fma(S.srec[0].L, c0, psrc - st.D0);
    } else {
      const double L_A = S.bL[pw][lane], L_AB = L_A + S.srec[n0].L;
      const double c_sib = c.fossa ? S.csib[0][lane] : 0.0;
      r0 = fma(L_AB, c0, fma(L_A, c_sib, S.bD[pw][lane] - st.D0));
      if (c.fossa) d84 = S.srec[n0].E + S.bE[pw][lane] + fma(L_AB, P.idt, S.bR[pw][lane] + S.srec[n0].R0);
    }
    set_first(c, st.x, r0);
  }
  SW_TICK(0);
  {  // sinus leaves into their clique partners (zero constants in the warps without leaves)
    const double a0 = C.le[0] * C.lw[0] * st.lx[0], b0 = C.lg[0] * C.lw[0] * st.lx[0];
    const double a1 = C.le[1] * C.lw[1] * st.lx[1], b1 = C.lg[1] * C.lw[1] * st.lx[1];
    st.x[0] += c.n2 ? a0 : 0.0; st.x[1] += c.n2 ? a1 - b0 : 0.0; st.x[2] -= c.n2 ? b1 : 0.0;
    st.x[5] += c.n2 ? 0.0 : a0; st.x[6] += c.n2 ? 0.0 : a1 - b0; st.x[7] -= c.n2 ? 0.0 : b1;
  }
  SW_TICK(1);
  double rc = st.x[0], dra = 0.0;
#pragma unroll
  for (int k = 0; k < KS - 1; k++) {  // (T = W = H = 0 on the slots that are not interior currents)
    const double g = rc * C.W[k];
    dra = fma(C.H[k], rc, dra);
    rc = fma(-C.T[k], rc, st.x[k + 1]);
    st.x[k] = g;
  }
  st.x[KS - 1] = rc;
  SW_TICK(2);
  if (!c.fossa) {
    S.sepR[wc_b(wc_)][lane] = rc;
    S.sideDR[w][lane] = dra;
  } else {
    // current 88 into 84, then 84 out of the clique {28, 29, 84} (couplings through section 28)
    const double r84 = r0 - dra - C.FF * C.WB * rc;
    const double d84b = d84 - C.CA;
    const double e = S.bE[pw][lane], g = e + S.bG[pw][lane];
    const double wv = fast_rcp(d84b);
    S.lf[0][lane] = e * e * wv;        // pivot of 28
    S.lf[1][lane] = -e * wv * r84;     // rhs of 28
    S.lf[2][lane] = g * g * wv;        // pivot of 29
    S.lf[3][lane] = g * wv * r84;      // rhs of 29
    S.lf[4][lane] = -e + e * g * wv;   // coupling 28 <-> 29
    st.x[3] = r84; st.w84 = wv; st.e28 = e; st.g28 = g;
  }
  SW_TICK(3);
}

template <bool NOISE>
__device__ __forceinline__ void swarp_back(const TdsParams &P, TdsSmem &S, const int lane, const int w, const unsigned wc_,
                                           const SCfg &c, SRegs &st) {
  const STab &C = S.stab[stab_slot(w)];
  double ua, ub;
  if (c.fossa) {
    const double u28 = S.sepR[S28][lane], u29 = S.sepR[S29][lane];
    ua = (st.x[3] + st.e28 * u28 - st.g28 * u29) * st.w84;
    ub = (st.x[KS - 1] - C.FF * ua) * C.WB;
  } else {
    ub = S.sepR[wc_b(wc_)][lane];
    ua = c.head ? 0.0 : S.sepR[wc_a(wc_)][lane];
  }
  st.x[KS - 1] = ub;
  double un = ub;
#pragma unroll
  for (int k = KS - 2; k >= 0; k--) {
    un = fma(-C.T[k], un, fma(-C.H[k], ua, st.x[k]));
    st.x[k] = un;
  }
  if (c.fossa) {
    st.x[3] = ua;
    if (NOISE) S.nmag[41][lane] = fma(P.nm_a, S.nmag[41][lane], P.nm_1a * ua);
  }
  {
    const double xa0 = c.n2 ? st.x[0] : st.x[5], xm = c.n2 ? st.x[1] : st.x[6], xb1 = c.n2 ? st.x[2] : st.x[7];
    st.lx[0] = (st.lx[0] + C.le[0] * xa0 - C.lg[0] * xm) * C.lw[0];
    st.lx[1] = (st.lx[1] + C.le[1] * xm - C.lg[1] * xb1) * C.lw[1];
  }
  S.uext[w][lane] = first_of(c, st.x);
}

// State dump (diagnostics): the caller runs pass A in two halves around it so that x still holds the solution.
__device__ __forceinline__ void swarp_dump(const TdsParams &P, TdsSmem &S, const int lane, const SCfg &c, const SRegs &st,
                                           double *dump) {
  if (lane != 0) return;
#pragma unroll
  for (int k = 0; k < KS; k++)
    if (k >= c.ks) dump_section(P.idt, P.ith, dump, c.base + k, st.p[k], st.pr[k], S.wall[c.base + k][0][0], st.x[k], st.c[k]);
  if (c.leaves)
    for (int j = 0; j < 2; j++) dump_section(P.idt, P.ith, dump, c.leaf0 + j, st.lp[j], st.lpr[j], S.wall[c.leaf0 + j][0][0], st.lx[j], st.lc[j]);
}
__device__ __forceinline__ void dwarp_dump(const TdsParams &P, TdsSmem &S, const int lane, const int w, const DRegs &st,
                                           double *dump) {
  if (lane != 0) return;
#pragma unroll
  for (int k = 0; k < KD; k++) {
    const int sec = 23 + KD * w + k;
    dump_section(P.idt, P.ith, dump, sec, st.p[k], st.pr[k], S.wall[sec][0][0], st.x[k], st.c[k]);
  }
}

// ==================================================================================
// Radiation loads (warp X2): the two parallel radiation currents out of the mouth (93, 94) and the nostrils (95, 96)
// ==================================================================================
// Filters of the non-default options (warp X2): x / y histories (index 0 = previous sample) of the two velum filters
// (order 4, TdsModel.cpp:476-533) and of the 50 Hz Chebyshev filter of the transglottal pressure (:474).  Every
// filter is split into a0 * x(n) -- evaluated where the value is needed, right after the barrier -- and the history
// part, which X2 prepares one step ahead.
struct OptRegs {
  double x1[4], y1[4], x2[4], y2[4], xg[4], yg[4];
};
__device__ __forceinline__ double iir_history(const double *a, const double *b, const double (&x)[4], const double (&y)[4]) {
  double h = 0.0;
#pragma unroll
  for (int i = 1; i <= 4; i++) { h = fma(a[i], x[i - 1], h); h = fma(b[i], y[i - 1], h); }
  return h;
}
__device__ __forceinline__ void iir_push(double (&x)[4], double (&y)[4], double in, double out) {
#pragma unroll
  for (int i = 3; i > 0; i--) { x[i] = x[i - 1]; y[i] = y[i - 1]; }
  x[0] = in; y[0] = out;
}
struct RadRegs {
  double c[4], u[4];  // history terms and last solutions of the currents 93..96
  double keep[2][7];  // factors kept between elimination and back substitution
};
// Rows + elimination of one radiation pair (TdsModel.cpp:1841-1911) hanging on the last section of warp pw:
// iR through the radiation resistance, iL through the radiation inductance.  Contribution to the separator -> radc.
__device__ __forceinline__ void radiation_eliminate(const TdsParams &P, TdsSmem &S, const int lane, int pw, int q,
                                                    double cR, double cL, double lips, double (&keep)[7]) {
  const double E_A = S.bE[pw][lane], L_A = S.bL[pw][lane], R_A = S.bR[pw][lane], invr = S.bI[pw][lane];
  const double D_A = S.bD[pw][lane], G_A = S.bG[pw][lane];
  const double R_rad = (128.0 * kRho * kC / (9.0 * kPi * kPi)) * (invr * invr * (1.0 / kPi));
  const double L_rad = (8.0 * kRho / (3.0 * kPi * kPi)) * invr;
  const double L_AB = L_A + L_rad;
  const double dR = E_A + G_A + R_rad, rR = fma(L_A, cR + cL, D_A) + lips;  // S = -lipsDipole (TdsModel.cpp:1853)
  const double dL = E_A + fma(L_AB, P.idt, R_A), rL = fma(L_A, cR, L_AB * cL) + D_A + lips;
  const double e = E_A, g = e + G_A;
  const double wL = fast_rcp(dL);
  double dds = e * e * wL, drs = -e * wL * rL;
  const double dR2 = dR - g * g * wL, rR2 = rR - g * wL * rL;
  const double o = -e + e * g * wL;
  const double wR = fast_rcp(dR2);
  dds += o * o * wR; drs += o * wR * rR2;
  S.radc[q][0][lane] = dds; S.radc[q][1][lane] = drs;
  keep[0] = wL; keep[1] = rL; keep[2] = wR; keep[3] = rR2; keep[4] = o; keep[5] = e; keep[6] = g;
}
__device__ __forceinline__ void radiation_update(const TdsParams &P, const int lane, RadRegs &rr, NoiseRegs &nr,
                                                 const StepCtx &X, double *dump) {  // TdsModel.cpp:2063-2072
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double u = rr.u[j];
    const double ur = fma(u, P.idt, -rr.c[j]);
    rr.c[j] = fma(ur, P.th, u * P.idt);
  }
  if (X.dumping && lane == 0) {
    for (int j = 0; j < 4; j++) { dump[279 + 93 + j] = rr.u[j]; dump[376 + 93 + j] = (rr.c[j] - rr.u[j] * P.idt) * P.ith; }
  }
}
template <bool NOISE>
__device__ __forceinline__ void radiation_back(const TdsParams &P, TdsSmem &S, const int lane, RadRegs &rr, NoiseRegs &nr) {
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const double *kp = rr.keep[q];
    const double us = S.sepR[q == 0 ? S64 : S83][lane];
    const double uR = (kp[3] - kp[4] * us) * kp[2];
    const double uL = (kp[1] + kp[5] * us - kp[6] * uR) * kp[0];
    rr.u[2 * q] = uR; rr.u[2 * q + 1] = uL;
    S.uext[UX_RADM + q][lane] = uR + uL;
  }
  if (NOISE) {  // noiseMagnitude of the mouth radiation currents (TdsModel.cpp:2052), for the lip constriction of the next step
    nr.nm93 = fma(P.nm_a, nr.nm93, P.nm_1a * rr.u[0]);
    nr.nm94 = fma(P.nm_a, nr.nm94, P.nm_1a * rr.u[1]);
    S.nzlip[lane] = nr.nm93 + nr.nm94;
  }
}

// ==================================================================================
// Noise sources (pre-pass kernels; warps X1, X2 and the dynamic chain warps): TdsModel::calcNoiseSources /
// calcNoiseSample (TdsModel.cpp:1188-1708)
// ==================================================================================
// Everything that DECIDES something (which sections form a constriction, where its obstacle is, how many random
// numbers are drawn) repeats the reference's arithmetic operation by operation (explicit round-to-nearest
// intrinsics: no FMA contraction, exact divisions), so that the discrete choices and the libc rand() stream are
// the reference's; amplitudes are smooth functions of the state and follow it to rounding level.
//
// Division of labour:
//   * pre-pass kernels (state-independent): the constrictions of every step.  They depend on the geometry only, and
//     the geometry of an interval between two control frames is a LINEAR function of the step: every comparison the
//     scan makes changes its outcome at most once per interval, so the constrictions are piecewise constant in runs
//     of steps.  tds_noise_runs_kernel finds the run boundaries by bisection (a few hundred scans per utterance
//     instead of one per step) and writes one record per run; tds_noise_scan_kernel scans every step (control
//     frames only a few steps apart, Synthesizer::playTargetSequence style);
//   * warp X1 keeps the records of the current run in shared memory (nzc_i / nzc_c);
//   * warp X2, pass A (noise_stage): everything that depends on the state -- flow through the constrictions -> target
//     amplitudes -> amplitude filters, decay of the sources nothing feeds, one random sample per audible source in
//     the reference's order; the flow-independent factors are prepared one step ahead beside the solve;
//   * the chain warps add sample = out * amp of their own sections to the right-hand sides.
__device__ __forceinline__ uint32_t nzc_pack(const NoiseCons &c) {
  return (uint32_t)(c.valid & 1) | ((uint32_t)(c.kind & 3) << 1) | ((uint32_t)(c.teeth & 1) << 3) | ((uint32_t)(c.lateral & 1) << 4) |
         ((uint32_t)(c.nar + 1) << 8) | ((uint32_t)c.up << 16);
}
// the four slot words of a step, with the "a later slot feeds this source as well" bits (5: upstream, 6: downstream)
__device__ __forceinline__ void nzc_pack4(const NoiseCons (&C)[4], uint32_t (&w)[4]) {
  constexpr int NPM = AFS_NUM_PM_SECTIONS;
  unsigned long long later = 0;
#pragma unroll
  for (int k = 3; k >= 0; k--) {
    const int up = C[k].up, down = up < NPM - 1 ? up + 1 : NPM;
    w[k] = nzc_pack(C[k]) | (uint32_t)((later >> up) & 1ull) << 5 | (uint32_t)((later >> down) & 1ull) << 6;
    if (C[k].valid) later |= (1ull << up) | (1ull << down);
  }
}

// glibc srand() (stdlib/random_r.c, TYPE_3): the reference draws from libc rand() (TdsModel.cpp:1690-1696); every
// tract owns a private generator.  Word i of the sequence sits at ring position (i + 28) % 36, so that the first
// output (i = 344) starts a block of twelve at position 12.
__device__ __noinline__ void rng_seed(TdsSmem &S, const int lane, uint32_t seed, int &ph) {
  if (seed == 0) seed = 1;
  long long r = (int32_t)seed;
  S.rng[28][lane] = (uint32_t)(int32_t)r;
  for (int i = 1; i < 31; i++) {
    const long long hi = r / 127773, lo = r % 127773;
    long long word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    r = (int32_t)word;
    S.rng[(i + 28) % 36][lane] = (uint32_t)(int32_t)r;
  }
  for (int i = 31; i < 34; i++) S.rng[(i + 28) % 36][lane] = S.rng[(i - 31 + 28) % 36][lane];
  for (int i = 34; i < 344; i++) S.rng[(i + 28) % 36][lane] = S.rng[(i - 31 + 28) % 36][lane] + S.rng[(i - 3 + 28) % 36][lane];
  ph = 12;
}
// Twelve successive rand() calls summed with the reference's int wrap-around (TdsModel.cpp:1690-1696), the new words
// going to ring positions ph .. ph + 11 (ph = 0, 12 or 24).  The additive lagged-Fibonacci recurrence lets all operands
// be loaded up front (the fifteen old words the twelve new ones depend on).  ONE copy of the code for the three
// phases (the time loop lives off its instruction cache): the three runs of old words start at ring positions that
// depend on the phase, inside a run every address is base + constant.
__device__ __forceinline__ uint32_t rng_block(TdsSmem &S, const int lane, const int ph) {
  const uint32_t *a0 = &S.rng[ph + 5][lane];                        // words i - 31 of the first seven: ph + 5 .. ph + 11
  const uint32_t *a1 = &S.rng[ph == 24 ? 0 : ph + 12][lane];        // ... of the last five: (ph + 12 .. ph + 16) mod 36
  const uint32_t *m0 = &S.rng[ph == 0 ? 33 : ph - 3][lane];         // words i - 3 of the first three: (ph - 3 .. ph - 1) mod 36
  uint32_t a[12], m[3], o[12];
#pragma unroll
  for (int q = 0; q < 7; q++) a[q] = a0[q * LANES];
#pragma unroll
  for (int q = 7; q < 12; q++) a[q] = a1[(q - 7) * LANES];
#pragma unroll
  for (int q = 0; q < 3; q++) m[q] = m0[q * LANES];
  uint32_t isum = 0;
#pragma unroll
  for (int q = 0; q < 12; q++) {
    o[q] = a[q] + (q < 3 ? m[q] : o[q < 3 ? 0 : q - 3]);
    isum += o[q] >> 1;
  }
  uint32_t *w = &S.rng[ph][lane];
#pragma unroll
  for (int q = 0; q < 12; q++) w[q * LANES] = o[q];
  return isum;
}

// exp(t) for -0.75 <= t <= 0 (cutoff <= 2000 Hz at fs >= 16.8 kHz): Taylor polynomial of degree 17 (full double
// accuracy on that range), evaluated by Estrin's scheme -- five dependent levels instead of seventeen: warp X2 runs
// alone, so latency is what counts; libm exp else
__device__ __forceinline__ double exp_small(double t) {
  if (t < -0.75) return exp(t);
  const double t2 = t * t, t4 = t2 * t2, t8 = t4 * t4, t16 = t8 * t8;
  const double p0 = fma(t, 1.0, 1.0), p1 = fma(t, 1.0 / 6.0, 0.5), p2 = fma(t, 1.0 / 120.0, 1.0 / 24.0),
               p3 = fma(t, 1.0 / 5040.0, 1.0 / 720.0), p4 = fma(t, 1.0 / 362880.0, 1.0 / 40320.0),
               p5 = fma(t, 1.0 / 39916800.0, 1.0 / 3628800.0), p6 = fma(t, 1.0 / 6227020800.0, 1.0 / 479001600.0),
               p7 = fma(t, 1.0 / 1307674368000.0, 1.0 / 87178291200.0),
               p8 = fma(t, 1.0 / 355687428096000.0, 1.0 / 20922789888000.0);
  const double q0 = fma(t2, p1, p0), q1 = fma(t2, p3, p2), q2 = fma(t2, p5, p4), q3 = fma(t2, p7, p6);
  const double o0 = fma(t4, q1, q0), o1 = fma(t4, q3, q2);
  return fma(t16, p8, fma(t8, o1, o0));
}
// Tube::interpolate (Tube.cpp:449-466) exactly as the reference rounds it
__device__ __forceinline__ double nz_lerp(double r1, double a, double r, double b) { return __dadd_rn(__dmul_rn(r1, a), __dmul_rn(r, b)); }
// (works on the control frames staged in shared memory: TdsSmem in the time loop, ScanFrames in the pre-pass kernels)
template <class FR>
__device__ __forceinline__ double nz_area(const FR &S, const int lane, double r, double r1, int j) {
  const double a = nz_lerp(r1, S.fA[0][j][lane], r, S.fA[1][j][lane]);
  return a < kMinArea ? kMinArea : a;
}
template <class FR>
__device__ __forceinline__ double nz_len(const FR &S, const int lane, double r, double r1, int j) {
  return nz_lerp(r1, S.fLen[0][j][lane], r, S.fLen[1][j][lane]);
}

// Control frames of one interval as the constriction scan needs them, [item][lane] (NL lanes per row).
template <int NL>
struct ScanFrames {
  double fA[2][AFS_NUM_PM_SECTIONS][NL], fLen[2][AFS_NUM_PM_SECTIONS][NL], fLat[2][AFS_NUM_PM_SECTIONS][NL];
  double inc[2][NL];                              // incisor position of the two control frames
  unsigned char art[AFS_NUM_PM_SECTIONS][NL];     // articulator of the LEFT control frame (Tube.cpp:452)
};

// widen a constriction around its narrowest section (TdsModel.cpp:1260-1275 and the copies below it)
template <class FR>
__device__ __forceinline__ void nz_expand(const FR &S, const int lane, double r, double r1, int idx, int artX,
                                          double minA, int &first, int &last) {
  const double maxA = __dadd_rn(minA, 0.2);
  first = last = idx;
  while ((nz_area(S, lane, r, r1, first) < maxA) && (S.art[first][lane] == artX) && (first > 0)) first--;
  while ((nz_area(S, lane, r, r1, last) < maxA) && (S.art[last][lane] == artX) && (last < AFS_NUM_PM_SECTIONS - 1)) last++;
  first++; last--;
}

// What identifies the OUTCOME of a scan: every index and every yes / no the scan arrived at.  All of them come from
// comparisons of quantities that are linear in the control-frame ratio (areas, lengths, positions, laterality of ONE
// section), so a comparison changes its outcome at most once per interval: two steps of an interval with equal
// signatures enclose steps with that same signature, and the constriction records of those steps differ only in the
// (smooth) downstream factor.
struct ScanSig {
  unsigned long long a, b;
  __device__ bool operator==(const ScanSig &o) const { return a == o.a && b == o.b; }
};

// State-independent part: the constrictions of this sample and where their noise sources sit.
// C[k].branch tells how the obstacle position moves with r: 0 fixed at 1.5 cm, 1 the incisors, 2 the middle of the
// section behind the constriction (downF = 1/2), 3 the end of the constriction's last section (downF = 1).
template <class FR>
__device__ __noinline__ void noise_scan(const FR &S, const int lane, const double r, const double r1, NoiseCons (&C)[4],
                                        ScanSig *sig) {
  constexpr int NPM = AFS_NUM_PM_SECTIONS;
#pragma unroll
  for (int k = 0; k < 4; k++) { C[k].valid = 0; C[k].nar = -1; C[k].kind = 0; C[k].up = 0; C[k].teeth = 0; C[k].lateral = 0; C[k].branch = 0; C[k].downF = 0.0; }
  // narrowest tongue and lip sections
  double minT = 1000000.0, minL = 1000000.0;
  int iT = -1, iL = -1;
#pragma unroll 4
  for (int j = 0; j < NPM; j++) {
    const double a = nz_area(S, lane, r, r1, j);
    const int ar = S.art[j][lane];
    if ((ar == AFS_TONGUE) && (a < minT)) { minT = a; iT = j; }
    if ((ar == AFS_LOWER_LIP) && (a < minL)) { minL = a; iL = j; }
  }
  int f1 = 0, l1 = -1, f2 = 0, l2 = -1, fL = 0, lL = -1, n2 = -1;
  bool t1 = false, t2 = false, cand2 = false;
  double a2 = 1000000.0;
  if (minT < 1.0) { t1 = true; nz_expand(S, lane, r, r1, iT, AFS_TONGUE, minT, f1, l1); }
  if (t1) {  // a second tongue constriction outside the first (TdsModel.cpp:1310-1372)
#pragma unroll 4
    for (int j = 0; j < NPM; j++) {
      const double a = nz_area(S, lane, r, r1, j);
      if ((S.art[j][lane] == AFS_TONGUE) && (a < a2) && ((j < f1) || (j > l1))) { a2 = a; n2 = j; }
    }
    if (a2 < 1.0) {
      cand2 = true;
      nz_expand(S, lane, r, r1, n2, AFS_TONGUE, a2, f2, l2);
      if ((f2 > l1 + 1) || (l2 < f1 - 1)) t2 = true;
    }
  }
  const bool lipCand = minL < 1.0;
  if (lipCand) nz_expand(S, lane, r, r1, iL, AFS_LOWER_LIP, minL, fL, lL);
  // laterality of the tongue constrictions: only `max > 0.1` matters (TdsModel.cpp:1545); the section that holds the
  // maximum goes into the signature (a maximum over sections is not linear in r, one section's value is)
  int am1 = 0, am2 = 0;
  auto range_lat = [&](int f, int l, int &amax) {
    double m = 0.0;
    amax = 0;
    for (int i = f; i <= l; i++) {
      const double v = nz_lerp(r1, S.fLat[0][i][lane], r, S.fLat[1][i][lane]);
      if (v > m) { m = v; amax = i + 1; }
    }
    return m;
  };
  double lat1 = 0.0, lat2 = 0.0;
  if (t1) lat1 = range_lat(f1, l1, am1);
  if (t2) lat2 = range_lat(f2, l2, am2);
  // positions (Tube.cpp:329-346): one sequential pass that picks up what the constrictions need
  const double teeth = nz_lerp(r1, S.inc[0][lane], r, S.inc[1][lane]);
  double pos = 0.0;
  bool have15 = false, haveT = false;
  int s15 = -1, sT = -1;
  double p15 = 0.0, e15 = 1.0, pT = 0.0, eT = 1.0;
  double pL1 = 0.0, eL1 = 0.0, pN1 = 0.0, eN1 = 1.0, pL2 = 0.0, eL2 = 0.0, pN2 = 0.0, eN2 = 1.0, pLL = 0.0, eLL = 1.0, pNL = 0.0;
  for (int j = 0; j < NPM; j++) {
    const double len = nz_len(S, lane, r, r1, j);
    const double end = __dadd_rn(pos, len);
    if (!have15 && (pos <= 1.5) && (end >= 1.5)) { have15 = true; s15 = j; p15 = pos; e15 = len; }
    if (!haveT && (pos <= teeth) && (end >= teeth)) { haveT = true; sT = j; pT = pos; eT = len; }
    if (j == l1) { pL1 = pos; eL1 = len; }
    if (j == l1 + 1) { pN1 = pos; eN1 = len; }
    if (j == l2) { pL2 = pos; eL2 = len; }
    if (j == l2 + 1) { pN2 = pos; eN2 = len; }
    if (j == lL) { pLL = pos; eLL = len; }
    if (j == lL + 1) pNL = pos;
    pos = end;
    if (have15 && (haveT || !t1) && (j > l1) && (j > l2) && (j > lL)) break;  // (teeth position only matters with a tongue constriction)
  }
  // constriction 0: the glottis, obstacle 1.5 cm downstream (TdsModel.cpp:1222-1229)
  if (have15) {
    C[0].valid = 1; C[0].nar = -1; C[0].kind = 0; C[0].up = s15; C[0].branch = 0;
    C[0].downF = fast_div(__dsub_rn(1.5, p15), e15);
  }
  double minTongueAreaForTeeth = 1000000.0;
  unsigned side = 0;  // per tongue constriction: 0 teeth branch, else on which side of the incisors the obstacle sits
  // tongue constrictions: the jet hits the teeth when they are closer than 2 cm, else the next section
  auto tongue = [&](NoiseCons &c, int nar, double aNar, int last, double pL, double eL, double pN, double eN, double lat, int shift) {
    const double jetPos = __dadd_rn(pL, eL);
    double x, p, e;
    int sec;
    bool ok;
    if (__dsub_rn(teeth, jetPos) < 2.0) {
      x = teeth; sec = sT; p = pT; e = eT; ok = haveT;
      minTongueAreaForTeeth = aNar;
      c.branch = 1;
    } else {
      x = __dadd_rn(pN, __dmul_rn(0.5, eN)); sec = last + 1; p = pN; e = eN; ok = true;  // (inside section last+1)
      c.branch = 2;
      const double d = __dsub_rn(x, teeth);
      side |= (d <= -0.0001 ? 1u : d < 0.0001 ? 2u : 3u) << shift;
    }
    c.valid = ok ? 1 : 0; c.nar = nar; c.kind = 1; c.up = ok ? sec : 0;
    c.downF = fast_div(__dsub_rn(x, p), e);
    c.teeth = fabs(__dsub_rn(x, teeth)) < 0.0001 ? 1 : 0;
    c.lateral = lat > 0.1 ? 1 : 0;
  };
  if (t1) tongue(C[1], iT, minT, l1, pL1, eL1, pN1, eN1, lat1, 0);
  if (t2) tongue(C[2], n2, a2, l2, pL2, eL2, pN2, eN2, lat2, 2);
  // lip constriction (TdsModel.cpp:1380-1418): the obstacle is the downstream end of its last section
  if (lipCand && (minL < minTongueAreaForTeeth)) {
    C[3].valid = 1; C[3].nar = iL; C[3].kind = 2; C[3].up = lL; C[3].branch = 3;
    C[3].downF = fast_div(__dsub_rn(pNL, pLL), eLL);
  }
  if (sig) {
    auto six = [](int v) { return (unsigned long long)((v + 1) & 63); };  // -1 .. 62
    sig->a = six(iT) | six(iL) << 6 | six(t1 ? f1 : -1) << 12 | six(t1 ? l1 : -1) << 18 | six(cand2 ? n2 : -1) << 24 |
             six(cand2 ? f2 : -1) << 30 | six(cand2 ? l2 : -1) << 36 | six(lipCand ? fL : -1) << 42 | six(lipCand ? lL : -1) << 48 |
             six(have15 ? s15 : -1) << 54;
    sig->b = six((t1 && haveT) ? sT : -1) | (unsigned long long)am1 << 6 | (unsigned long long)am2 << 12 |
             (unsigned long long)(lat1 > 0.1) << 18 | (unsigned long long)(lat2 > 0.1) << 19 | (unsigned long long)t1 << 20 |
             (unsigned long long)t2 << 21 | (unsigned long long)cand2 << 22 | (unsigned long long)lipCand << 23 |
             (unsigned long long)C[3].valid << 24 | (unsigned long long)C[1].branch << 25 | (unsigned long long)C[2].branch << 27 |
             (unsigned long long)side << 29 | (unsigned long long)C[1].valid << 33 | (unsigned long long)C[2].valid << 34 |
             (unsigned long long)C[1].teeth << 35 | (unsigned long long)C[2].teeth << 36;
  }
}

// One sample of a source from the sum of twelve rand() draws (TdsModel.cpp:1680-1708).  S.ncut[j] holds the low-pass
// coefficient exp(-2 pi f_c dt) of the source's cutoff frequency (IirFilter::createSinglePoleLowpass,
// IirFilter.cpp:235-243), written by warp X2 whenever a constriction feeds the source: a source that only decays
// keeps its cutoff, so nobody evaluates an exponential for it.
__device__ __forceinline__ double nz_sample(const double x, const uint32_t isum, const double last) {
  double in = (double)(int32_t)isum;  // 12 x rand() summed in an int: wraps around like the reference's
  in = (in * (1.0 / 2147483647.0) - 6.0) * 0.28867513459481288225;  // / RAND_MAX, - 6, / sqrt(12)
  return __dadd_rn(__dmul_rn(__dsub_rn(1.0, x), in), __dmul_rn(x, last));
}
__device__ __forceinline__ void nz_lowpass(TdsSmem &S, const int lane, const int j, const uint32_t isum) {
  S.nout[j][lane] = nz_sample(S.ncut[j][lane], isum, S.nout[j][lane]);
}
// amplitude filter of calcNoiseSample (TdsModel.cpp:1637-1641), the reference's roundings
__device__ __forceinline__ double nz_amp(double old, double target, double F) { return __dadd_rn(old, __dmul_rn(F, __dsub_rn(target, old))); }
// Warp X2, behind its sample loop: new sums for the entries the step has consumed, three at a time (the generator is
// back at phase 12 after three sums, so all lanes run the same code); up to two entries wait for the next step.
// Publishes head and fill level.
__device__ __forceinline__ void nz_refill(TdsSmem &S, const int lane, NoiseRegs &N, int triples) {
  int t = N.head + (N.cap - N.free);
  if (t >= N.cap) t -= N.cap;
  for (; N.free >= 3 && triples > 0; triples--) {
#pragma unroll 1
    for (int j = 0; j < 3; j++) {  // phases 12, 24, 0
      S.nzsum[t][lane] = rng_block(S, lane, j == 2 ? 0 : 12 + 12 * j);
      t = t + 1 >= N.cap ? 0 : t + 1;
    }
    N.free -= 3;
  }
  S.nzring[lane] = (uint32_t)N.head | ((uint32_t)(N.cap - N.free) << 8) | ((uint32_t)N.cap << 16);
}

// gain * sqrt(A) / A^3 and 0.15 / (A d), d = sqrt(4 A / pi), for the constriction area A (TdsModel.cpp:1560-1600):
// fullAmp = amp * flow^3, cutoff = cut * flow.
__device__ __forceinline__ void nz_factors(double A, const double gain, double &amp, double &cut) {
  if (A < kMinArea) A = kMinArea;
  if (A < 0.1) A = 0.1;
  const double rs = fast_rsqrt(A), rs2 = rs * rs;  // v = flow / A = flow rs^2
  amp = gain * (A * rs) * (rs2 * rs2 * rs2);
  cut = 0.15 * rs2 * (rs * 0.88622692545275801365);
}
__device__ __forceinline__ double nz_down(const TdsSmem &S, const int lane, const int k, const double r) {
  return fast_div(fma(r, S.nzc_c[k][1][lane], S.nzc_c[k][0][lane]), fma(r, S.nzc_c[k][3][lane], S.nzc_c[k][2][lane]));
}
// Slot word of a constriction record: bit 0 valid, 1-2 kind, 3 teeth, 4 lateral, 5 / 6 the upstream / downstream source
// is fed by a LATER slot as well (the last one wins, TdsModel.cpp:1430-1612 writes the targets slot after slot),
// 8-15 nar + 1, 16-31 up.  nz_fed: the sources the four slots of a tract feed (bit j = source j).
__device__ __forceinline__ unsigned long long nz_fed(const TdsSmem &S, const int lane) {
  constexpr int NPM = AFS_NUM_PM_SECTIONS;
  unsigned long long fed = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t v = S.nzc_i[k][lane];
    const int up = (int)(v >> 16), down = up < NPM - 1 ? up + 1 : NPM;
    if (v & 1u) fed |= (1ull << up) | (1ull << down);
  }
  return fed;
}

// One constriction slot of calcNoiseSources, state-dependent part (TdsModel.cpp:1430-1612): flow through the
// constriction -> target amplitude -> amplitude filter (calcNoiseSample, :1637-1660) of the two sources it feeds.
// amp / cut / downF: the flow-independent factors (nz_factors, nz_down).  Returns the fed sources that are audible.
constexpr int NZ_BAR4 = 160;  // named barrier 4: F, X1, N1 and T2 arrive, X2 waits
struct NzConst { double F, x2000, dt; };  // P.nz_F, P.nz_x2000, P.dt by value (the callees are not inlined)
__device__ __forceinline__ unsigned long long nz_slot(const NzConst &P, TdsSmem &S, const int lane, const int k, const uint32_t v,
                                                      const double amp, const double cut, const double downF) {
  constexpr int NPM = AFS_NUM_PM_SECTIONS;
  const bool valid = v & 1u;
  const int nar = min(max((int)((v >> 8) & 255u) - 1, 0), NPM - 1), up = (int)(v >> 16), down = up < NPM - 1 ? up + 1 : NPM;
  double flow;
  if (k == 0) flow = S.nmag[0][lane];  // the glottis: current 25
  else {                               // pharynx/mouth section nar
    const double f = S.nmag[nar + 1][lane];
    const double extra = S.nmag[nar == 3 ? 41 : 40][lane];  // section 28: currents 29 and 84; section 40: currents 41 and 65
    flow = nar == NPM - 1 ? S.nzlip[lane] : (nar == 3 || nar == 15) ? f + extra : f;
  }
  if (flow < 0.0) flow = 0.0;
  const double fullAmp = amp * (flow * flow * flow);  // gain |v| v^2 sqrt(A), v = flow / A >= 0
  // cutoff: clamped to 50..2000 Hz; the vocal-fold and lip sources ask for 6000 Hz, i.e. sit at the upper limit
  double x = P.x2000;
  if (k == 1 || k == 2) {
    double cutoff = cut * flow;  // 0.15 v / d
    if (cutoff < 50.0) cutoff = 50.0;
    if (cutoff < 2000.0) x = exp_small(-2.0 * kPi * (cutoff * P.dt));
  }
  unsigned long long draw = 0;
#pragma unroll 1
  for (int side = 0; side < 2; side++) {
    const int j = side ? down : up;
    if (valid && !((v >> (5 + side)) & 1u)) {  // (not a source a later slot feeds as well)
      const double target = side ? downF * fullAmp : (1.0 - downF) * fullAmp;
      const double old = S.namp[j][lane];
      double a = nz_amp(old, target, P.F);
      S.ncut[j][lane] = x;
      if ((old >= 0.001) && (a < 0.001)) S.nout[j][lane] = 0.0;
      if (a >= 0.001) draw |= 1ull << j;
      if ((a < 1.0e-13) && (target == 0.0)) a = 0.0;  // (the reference lets it decay for ever; a source is silent below 1e-3)
      S.namp[j][lane] = a;
    }
  }
  return draw;
}

// The state-dependent part of the noise sources is spread over three warps that have time in pass A, each with a
// short piece, so that none of them becomes the longest path of the step (a lone warp pays an instruction fetch for
// almost every instruction it executes once per step: what counts is the number of instructions on its path):
//   warp F  (noise_decay, before its own pass A: it has the most slack of the static chain warps): the sources nothing
//           feeds this step decay (calcNoiseSample, TdsModel.cpp:1637-1660); the still audible ones -> nzdecay;
//   warps X1, N1, T2 (noise_slot; slots 1, 2, 3: X1 is idle in pass A otherwise, N1 and T2 are the static chain warps
//           with the most slack after F): the two tongue constrictions and the lip constriction, complete: factors,
//           flow, target amplitudes, amplitude filters; their audible sources -> nztongue;
//   warp X2 (noise_slot, noise_stage): the vocal-fold constriction (slot 0; its flow-independent factors were computed one step
//           ahead by X1 next to the glottis geometry: nzk0), then -- behind named barrier 4, where the others only arrive
//           -- one sample for every source that is not silent, in ascending source order: the reference draws from ONE
//           rand() stream in that order (TdsModel.cpp:1615-1620), here the k-th drawing source takes the k-th entry of
//           the ring of pre-generated sums.
// Every source is written by exactly one of them: a fed source by the warp of the LAST slot that feeds it (slot words,
// bits 5 / 6), the others by F.  Everything is finished before the barrier that ends pass A; the chain warps only add
// sample = out * amp to their right-hand sides.
__device__ __noinline__ void noise_decay(const double F, TdsSmem &S, const int lane, unsigned long long &live, const bool have_fed) {
  const unsigned long long fed = have_fed ? S.nzfed[lane] : nz_fed(S, lane);
  unsigned long long todo = live & ~fed, alive = fed, draw = 0;
  while (todo) {  // (one source per pass: the loop body stays in the instruction buffer of the sub-partition)
    const int j = __ffsll((long long)todo) - 1;
    todo &= todo - 1;
    const double old = S.namp[j][lane];
    double a = nz_amp(old, 0.0, F);
    if ((old >= 0.001) && (a < 0.001)) S.nout[j][lane] = 0.0;
    if (a >= 0.001) draw |= 1ull << j;
    if (a < 1.0e-13) a = 0.0;  // (the reference lets it decay for ever; a source is silent below 1e-3)
    else alive |= 1ull << j;
    S.namp[j][lane] = a;
  }
  live = alive;
  S.nzdecay[lane] = draw;
  bar_arrive(4, NZ_BAR4);
}

__device__ __noinline__ unsigned long long noise_slot(const NzConst P, TdsSmem &S, const int lane, const double r, const double r1,
                                                      const int k, const double gain0) {
  constexpr int NPM = AFS_NUM_PM_SECTIONS;
  unsigned long long draw = 0;
  const uint32_t v = S.nzc_i[k][lane];
  if (__any_sync(0xffffffffu, v & 1u)) {  // (the second tongue constriction and the lips are rare)
    double amp, cut, down;
    if (k == 0) {  // vocal folds: the flow-independent factors were computed one step ahead by X1 (nzk0)
      amp = gain0 * S.nzk0[0][lane]; cut = 0.0; down = S.nzk0[1][lane];
    } else {
      const bool teeth = (v >> 3) & 1u, lateral = (v >> 4) & 1u;
      const int nar = min(max((int)((v >> 8) & 255u) - 1, 0), NPM - 1);
      nz_factors(nz_area(S, lane, r, r1, nar), lateral ? 0.0 : k == 3 ? 2.0e-7 : teeth ? 10.0e-7 : 5.0e-7, amp, cut);
      down = nz_down(S, lane, k, r);
    }
    draw = nz_slot(P, S, lane, k, v, amp, cut, down);
  }
  return draw;
}

// Warp X2 behind barrier 4: one sample for every source that is not silent, ascending (returns the number of ring
// entries consumed); draw: the sources of its own slot 0.
__device__ __noinline__ int noise_stage(TdsSmem &S, const int lane, unsigned long long draw
#ifdef AFS_TDS_TIMING
                                        , long long *nz_t
#endif
                                        ) {
#ifdef AFS_TDS_TIMING
  long long sw_last = clock64();
  long long *sw_t = nz_t;
#endif
  SW_TICK(0);
  bar_sync(4, NZ_BAR4);
  SW_TICK(1);
  // samples, ascending: the k-th drawing source takes the k-th sum
  draw |= S.nzdecay[lane] | S.nztongue[0][lane] | S.nztongue[1][lane] | S.nztongue[2][lane];
  const uint32_t ring = S.nzring[lane];
  const int cap = (int)(ring >> 16);
  int i = (int)(ring & 255u), n_draw = 0;
  while (draw) {  // (one source per pass, see noise_decay)
    const int j = __ffsll((long long)draw) - 1;
    draw &= draw - 1;
    S.nout[j][lane] = nz_sample(S.ncut[j][lane], S.nzsum[i][lane], S.nout[j][lane]);
    i = i + 1 >= cap ? 0 : i + 1;
    n_draw++;
  }
  SW_TICK(2);
  return n_draw;
}

// The samples of the dipole sources of a chain warp's own sections into the right-hand sides: S = -sample
// (TdsModel.cpp:1936), sample = out * amp (0 while the source is silent: out = 0 then).
__device__ __forceinline__ void dwarp_noise(const TdsParams &P, TdsSmem &S, const int lane, const int w, DRegs &st) {
#pragma unroll
  for (int k = 0; k < KD; k++) {
    const int j = KD * w - 2 + k;  // section 23 + 6 w + k - 25
    if (j >= 0 && j < AFS_NUM_PM_SECTIONS) st.x[k] = fma(S.nout[j][lane], S.namp[j][lane], st.x[k]);
  }
}

// ==================================================================================
// Reduced system over the separators (three chains + the velum clique)
// ==================================================================================
__device__ __forceinline__ void red_elim(double &dx, double rx, double c, double &dy, double &ry) {  // x into y
  const double w = fast_rcp(dx);
  dx = w;
  const double t = c * w;
  dy = fma(-t, c, dy);
  ry = fma(-t, rx, ry);
}
// Constants of the trachea end of chain T (separators 7 and 15 have constant pivots): computed once per launch.
//   rcn[0] = 1/d7', [1] = c(7,15)/d7', [2] = 1/d15', [3] = c(15,22)/d15', [4] = c(15,22)^2/d15' (comes off d22)
__device__ __noinline__ void reduced_constants(TdsSmem &S, const int lane) {
  const double c0 = S.sideF[W_T1][lane], c1 = S.sideF[W_T2][lane];
  const double w7 = 1.0 / (S.sepD[S7][lane] - S.sideDD[W_T1][lane]);
  const double t7 = c0 * w7;
  const double w15 = 1.0 / (S.sepD[S15][lane] - S.sideDD[W_T2][lane] - t7 * c0);
  const double t15 = c1 * w15;
  if (lane == 0) { S.rcn[0] = w7; S.rcn[1] = t7; S.rcn[2] = w15; S.rcn[3] = t15; S.rcn[4] = t15 * c1; }
}
// One call per step and chain warp: forward elimination of the own chain -- chain T (7,15,22,28,29,34 -> 40) on T0,
// chain M (64,58,52,46 -> 41) on T1, chain N (83,75,70 -> 65) on T2 --, ONE barrier, the junction at the velum (clique
// {40, 41, 65} through section 40, the last section of warp P2) solved by all three warps for themselves, back
// substitution of the own chain.  The reduced system is pure latency of a warp that runs alone: the reciprocals and
// right-hand sides of the own chain stay in registers between the two sweeps, and in the two long chains two
// independent eliminations run interleaved (22 -> 28 beside 34 -> {29, 40}; 64 -> 58 beside 46 -> {41, 52}).
struct RedKeep { double w[4], r[6], c[7]; };
__device__ __forceinline__ void red_fwd_T(TdsSmem &S, const int lane, RedKeep &K, double &dj, double &rj) {
  const double c2 = S.sideF[W_GP][lane], c3 = S.lf[4][lane], c4 = S.sideF[W_P1][lane], c5 = S.sideF[W_P2][lane];
  const double r7 = S.sepR[S7][lane] - S.sideDR[W_T1][lane];
  const double r15 = fma(-S.rcn[1], r7, S.sepR[S15][lane] - S.sideDR[W_T2][lane]);
  double d22 = S.sepD[S22][lane] - S.sideDD[W_GP][lane] - S.rcn[4];
  const double r22 = fma(-S.rcn[3], r15, S.sepR[S22][lane] - S.sideDR[W_GP][lane]);
  double d28 = S.sepD[S28][lane] - S.lf[0][lane], r28 = S.sepR[S28][lane] - S.lf[1][lane];
  double d29 = S.sepD[S29][lane] - S.lf[2][lane], r29 = S.sepR[S29][lane] - S.lf[3][lane];
  double d34 = S.sepD[S34][lane] - S.sideDD[W_P2][lane];
  const double r34 = S.sepR[S34][lane] - S.sideDR[W_P2][lane];
  double d40 = S.sepD[S40][lane], r40 = S.sepR[S40][lane];
  // 22 -> 28  ||  34 -> {29, 40} (fill: coupling 29 <-> 40)
  const double w22 = fast_rcp(d22), w34 = fast_rcp(d34);
  const double t22 = c2 * w22, ta = c4 * w34, tb = c5 * w34;
  d28 = fma(-t22, c2, d28); r28 = fma(-t22, r22, r28);
  d29 = fma(-ta, c4, d29); r29 = fma(-ta, r34, r29);
  d40 = fma(-tb, c5, d40); r40 = fma(-tb, r34, r40);
  const double c6 = -ta * c5;
  // 28 -> 29, 29 -> 40
  const double w28 = fast_rcp(d28);
  const double t28 = c3 * w28;
  d29 = fma(-t28, c3, d29); r29 = fma(-t28, r28, r29);
  const double w29 = fast_rcp(d29);
  const double t29 = c6 * w29;
  d40 = fma(-t29, c6, d40); r40 = fma(-t29, r29, r40);
  K.w[0] = w22; K.w[1] = w28; K.w[2] = w29; K.w[3] = w34;
  K.r[0] = r7; K.r[1] = r15; K.r[2] = r22; K.r[3] = r28; K.r[4] = r29; K.r[5] = r34;
  K.c[2] = c2; K.c[3] = c3; K.c[4] = c4; K.c[5] = c5; K.c[6] = c6;
  dj = d40; rj = r40;
}
__device__ __forceinline__ void red_back_T(TdsSmem &S, const int lane, const RedKeep &K, const double u40) {
  const double c0 = S.sideF[W_T1][lane], c1 = S.sideF[W_T2][lane];
  const double u29 = (K.r[4] - K.c[6] * u40) * K.w[2];
  const double u34 = (K.r[5] - K.c[4] * u29 - K.c[5] * u40) * K.w[3];
  const double u28 = (K.r[3] - K.c[3] * u29) * K.w[1];
  const double u22 = (K.r[2] - K.c[2] * u28) * K.w[0];
  const double u15 = (K.r[1] - c1 * u22) * S.rcn[2];
  const double u7 = (K.r[0] - c0 * u15) * S.rcn[0];
  S.sepR[S40][lane] = u40;
  S.sepR[S34][lane] = u34; S.sepR[S29][lane] = u29; S.sepR[S28][lane] = u28; S.sepR[S22][lane] = u22;
  S.sepR[S15][lane] = u15; S.sepR[S7][lane] = u7;
}
__device__ __forceinline__ void red_fwd_M(TdsSmem &S, const int lane, RedKeep &K, double &dj, double &rj) {
  const double c0 = S.sideF[W_M3][lane], c1 = S.sideF[W_M2][lane], c2 = S.sideF[W_M1][lane], c3 = S.sideF[W_M0][lane];
  double d41 = S.sepD[S41][lane], r41 = S.sepR[S41][lane];
  double d46 = S.sepD[S46][lane] - S.sideDD[W_M1][lane];
  const double r46 = S.sepR[S46][lane] - S.sideDR[W_M1][lane];
  double d52 = S.sepD[S52][lane] - S.sideDD[W_M2][lane], r52 = S.sepR[S52][lane] - S.sideDR[W_M2][lane];
  double d58 = S.sepD[S58][lane] - S.sideDD[W_M3][lane], r58 = S.sepR[S58][lane] - S.sideDR[W_M3][lane];
  double d64 = S.sepD[S64][lane] - S.radc[0][0][lane];
  const double r64 = S.sepR[S64][lane] - S.radc[0][1][lane];
  // 64 -> 58  ||  46 -> {41, 52} (fill: coupling 41 <-> 52)
  const double w64 = fast_rcp(d64), w46 = fast_rcp(d46);
  const double t64 = c0 * w64, ta = c3 * w46, tb = c2 * w46;
  d58 = fma(-t64, c0, d58); r58 = fma(-t64, r64, r58);
  d41 = fma(-ta, c3, d41); r41 = fma(-ta, r46, r41);
  d52 = fma(-tb, c2, d52); r52 = fma(-tb, r46, r52);
  const double c6 = -ta * c2;
  // 58 -> 52, 52 -> 41
  const double w58 = fast_rcp(d58);
  const double t58 = c1 * w58;
  d52 = fma(-t58, c1, d52); r52 = fma(-t58, r58, r52);
  const double w52 = fast_rcp(d52);
  const double t52 = c6 * w52;
  d41 = fma(-t52, c6, d41); r41 = fma(-t52, r52, r41);
  K.w[0] = w64; K.w[1] = w58; K.w[2] = w52; K.w[3] = w46;
  K.r[0] = r64; K.r[1] = r58; K.r[2] = r52; K.r[3] = r46;
  K.c[0] = c0; K.c[1] = c1; K.c[2] = c2; K.c[3] = c3; K.c[6] = c6;
  dj = d41; rj = r41;
}
__device__ __forceinline__ void red_back_M(TdsSmem &S, const int lane, const RedKeep &K, const double u41) {
  const double u52 = (K.r[2] - K.c[6] * u41) * K.w[2];
  const double u46 = (K.r[3] - K.c[3] * u41 - K.c[2] * u52) * K.w[3];
  const double u58 = (K.r[1] - K.c[1] * u52) * K.w[1];
  const double u64 = (K.r[0] - K.c[0] * u58) * K.w[0];
  S.sepR[S41][lane] = u41;
  S.sepR[S46][lane] = u46; S.sepR[S52][lane] = u52; S.sepR[S58][lane] = u58; S.sepR[S64][lane] = u64;
}
__device__ __forceinline__ void red_fwd_N(TdsSmem &S, const int lane, RedKeep &K, double &dj, double &rj) {
  const double c0 = S.sideF[W_N2][lane], c1 = S.sideF[W_N1][lane], c2 = S.sideF[W_NV][lane];
  double d65 = S.sepD[S65][lane], r65 = S.sepR[S65][lane];
  double d70 = S.sepD[S70][lane] - S.sideDD[W_N1][lane], r70 = S.sepR[S70][lane] - S.sideDR[W_N1][lane];
  double d75 = S.sepD[S75][lane] - S.sideDD[W_N2][lane], r75 = S.sepR[S75][lane] - S.sideDR[W_N2][lane];
  double d83 = S.sepD[S83][lane] - S.radc[1][0][lane], r83 = S.sepR[S83][lane] - S.radc[1][1][lane];
  red_elim(d83, r83, c0, d75, r75);   // (leaves the reciprocal of the pivot in d)
  red_elim(d75, r75, c1, d70, r70);
  red_elim(d70, r70, c2, d65, r65);
  K.w[0] = d83; K.w[1] = d75; K.w[2] = d70;
  K.r[0] = r83; K.r[1] = r75; K.r[2] = r70;
  K.c[0] = c0; K.c[1] = c1; K.c[2] = c2;
  dj = d65; rj = r65;
}
__device__ __forceinline__ void red_back_N(TdsSmem &S, const int lane, const RedKeep &K, const double u65) {
  const double u70 = (K.r[2] - K.c[2] * u65) * K.w[2];
  const double u75 = (K.r[1] - K.c[1] * u70) * K.w[1];
  const double u83 = (K.r[0] - K.c[0] * u75) * K.w[0];
  S.sepR[S65][lane] = u65;
  S.sepR[S70][lane] = u70; S.sepR[S75][lane] = u75; S.sepR[S83][lane] = u83;
}
__device__ __noinline__ void reduced_solve(TdsSmem &S, const int lane, const int chain) {
  RedKeep K;
  double dj, rj;
  if (chain == 0) red_fwd_T(S, lane, K, dj, rj);
  else if (chain == 1) red_fwd_M(S, lane, K, dj, rj);
  else red_fwd_N(S, lane, K, dj, rj);
  S.redJ[chain][0][lane] = dj; S.redJ[chain][1][lane] = rj;
  bar_sync(3, 96);
  // junction (the same few instructions on the three warps, which share a sub-partition: no second barrier)
  const double e = S.bE[W_P2][lane], g = e + S.bG[W_P2][lane];
  const double r65 = S.redJ[2][1][lane], r41i = S.redJ[1][1][lane];
  const double w65 = fast_rcp(S.redJ[2][0][lane]);
  double d40 = S.redJ[0][0][lane] - e * e * w65;
  const double d41 = S.redJ[1][0][lane] - g * g * w65;
  const double o = -e + e * g * w65;
  double r40 = S.redJ[0][1][lane] + e * w65 * r65;
  const double r41 = r41i - g * w65 * r65;
  const double w41 = fast_rcp(d41);
  d40 -= o * o * w41;
  r40 -= o * w41 * r41;
  const double u40 = fast_div(r40, d40);
  if (chain == 0) { red_back_T(S, lane, K, u40); return; }
  const double u41 = (r41 - o * u40) * w41;
  if (chain == 1) { red_back_M(S, lane, K, u41); return; }
  const double u65 = (r65 + e * u40 - g * u41) * w65;
  red_back_N(S, lane, K, u65);
}

// ---- glottis (TriangularGlottis.cpp:154-576) -----------------------------------
struct GlottisState {
  double x0, x1;    // current relative displacements (lower, upper)
  double xp0, xp1;  // previous
};
struct GlottisDims {
  double cordLength, th0, th1, ol0, ol1, cl0, cl1, mw0, mw1, mz0, mz1, Q;
};
enum { G_F0, G_PRESSURE, G_REST1, G_REST2, G_ARY, G_ASP };
enum { S_REST_LENGTH, S_REST_THICK1, S_REST_THICK2, S_MASS1, S_MASS2, S_DAMP1, S_DAMP2, S_K1, S_K2, S_CK1, S_CK2,
       S_KC, S_INLET, S_OUTLET, S_NATURAL_F0, S_F0_DIV_Q };

__device__ __forceinline__ void glottis_open_close(double cordLength, double rest, double rel, double &ol, double &cl,
                                                   double &mw, double &mz) {
  const double back = rest + rel;
  const double front = (rest < 0.0) ? back : rel;
  ol = 0.0; mw = 0.0; cl = cordLength; mz = 0.5 * cordLength;
  if ((back > 0.0) && (front > 0.0)) {
    ol = cordLength; mw = back + front; cl = 0.0; mz = 0.0;
  } else if ((back <= 0.0) && (front <= 0.0)) {
  } else {
    if (fabs(rest) < 0.000000001) rest = 0.000000001;
    const double zApex = cordLength * (1.0 + fast_div(rel, rest));
    if ((zApex >= 0.0) && (zApex <= cordLength)) {
      if (back > 0.0) { ol = zApex; mw = back; cl = cordLength - zApex; mz = 0.5 * (zApex + cordLength); }
      else { ol = cordLength - zApex; mw = front; cl = zApex; mz = 0.5 * zApex; }
    }
  }
}
// tension Q, cord length, thicknesses (getTensionParameter / getLengthAndThickness)
__device__ __forceinline__ void glottis_tension(const TdsParams &P, double f0, GlottisDims &D) {
  double f0DivQ = P.gs[S_F0_DIV_Q];
  if (f0DivQ < 0.000001) f0DivQ = 0.000001;
  double Q = 1.0 + fast_div(f0 - P.gs[S_NATURAL_F0], f0DivQ);
  if (Q < 0.05) Q = 0.05;
  const double rs = fast_rsqrt(Q);  // 1/sqrt(Q)
  D.Q = Q;
  D.cordLength = P.gs[S_REST_LENGTH] * (Q * rs);
  D.th0 = P.gs[S_REST_THICK1] * rs;
  D.th1 = P.gs[S_REST_THICK2] * rs;
}
__device__ __forceinline__ void glottis_dims(const TdsParams &P, const double ctl[AFS_NUM_GLOTTIS_CONTROL],
                                             const GlottisState &g, GlottisDims &D) {
  glottis_tension(P, ctl[G_F0], D);
  glottis_open_close(D.cordLength, ctl[G_REST1], g.x0, D.ol0, D.cl0, D.mw0, D.mz0);
  glottis_open_close(D.cordLength, ctl[G_REST2], g.x1, D.ol1, D.cl1, D.mw1, D.mz1);
}
// TriangularGlottis::incTime
// `D` are the dimensions for the CURRENT displacement and control values: exactly what the geometry call of the
// previous step computed (same inputs), so they are cached instead of recomputed.
__device__ __forceinline__ void glottis_inc_time(const TdsParams &P, const double ctl[AFS_NUM_GLOTTIS_CONTROL],
                                                 GlottisState &g, const double p[4], const GlottisDims &D) {
  const double T = P.dt, Q = D.Q;
  const double iQ = fast_rcp(Q), iLen = fast_rcp(D.cordLength);
  const double m0 = P.gs[S_MASS1] * iQ, m1 = P.gs[S_MASS2] * iQ;
  const double al0 = D.cl0 * iLen, al1 = D.cl1 * iLen;
  const double k0 = P.gs[S_K1] * Q, k1 = P.gs[S_K2] * Q;
  const double ck0 = P.gs[S_CK1] * Q, ck1 = P.gs[S_CK2] * Q;
  const double kc = P.gs[S_KC] * Q * Q;
  const double dr0 = P.gs[S_DAMP1] + al0 * 1.0, dr1 = P.gs[S_DAMP2] + al1 * 1.0;
  // mass*spring = MASS*K (Q cancels): sqrt of a launch constant
  const double r0 = 2.0 * dr0 * P.g_sqrt_mk[0], r1 = 2.0 * dr1 * P.g_sqrt_mk[1];
  double f0 = p[1] * D.ol0 * D.th0, f1 = p[2] * D.ol1 * D.th1;
  f0 += 0.5 * 0.5 * (p[0] + p[1]) * P.gs[S_INLET] * D.cordLength;
  f1 += 0.5 * 0.5 * (p[3] + p[2]) * P.gs[S_OUTLET] * D.cordLength;
  const double rest0 = ctl[G_REST1], rest1 = ctl[G_REST2];
  const double rs0 = (rest0 >= 0.0) ? rest0 * (1.0 - D.mz0 * iLen) : rest0;
  const double rs1 = (rest1 >= 0.0) ? rest1 * (1.0 - D.mz1 * iLen) : rest1;
  const double A = m0 + r0 * T + T * T * (k0 + ck0 * al0) + kc * T * T;
  const double B = -kc * T * T;
  const double C = -kc * T * T;
  const double Dd = m1 + r1 * T + T * T * (k1 + ck1 * al1) + kc * T * T;
  const double E = f0 * T * T + 2.0 * m0 * g.x0 - m0 * g.xp0 + r0 * T * g.x0 - T * T * ck0 * al0 * rs0;
  const double F = f1 * T * T + 2.0 * m1 * g.x1 - m1 * g.xp1 + r1 * T * g.x1 - T * T * ck1 * al1 * rs1;
  double det = A * Dd - B * C;
  if (fabs(det) < 0.000000001) det = 0.000000001;
  const double idet = fast_rcp(det);
  const double n0 = (E * Dd - B * F) * idet;
  const double n1 = (A * F - E * C) * idet;
  g.xp0 = g.x0; g.xp1 = g.x1;
  g.x0 = n0; g.x1 = n1;
}
// TriangularGlottis::calcGeometry + getTubeData -> the two glottal tube sections
__device__ __forceinline__ void glottis_geometry(const TdsParams &P, const double ctl[AFS_NUM_GLOTTIS_CONTROL],
                                                 const GlottisState &g, double len[2], double area[2], GlottisDims &D) {
  glottis_dims(P, ctl, g, D);
  double chink = ctl[G_ARY];
  if (chink < 0.0) chink = 0.0;
  area[0] = D.ol0 * D.mw0 + chink;
  area[1] = D.ol1 * D.mw1 + chink;
  len[0] = D.th0;
  len[1] = D.th1;
}
// Direct-form IIR of IirFilter::getOutputSample (IirFilter.cpp:47-65); history x[i] = input i+1 samples
// ago, kept in the registers of warp X1.
template <int ORDER>
__device__ __forceinline__ double iir_step(const double *a, const double *b, double (&x)[ORDER], double (&y)[ORDER], double in) {
  double sum = a[0] * in;
#pragma unroll
  for (int i = 1; i <= ORDER; i++) { sum += a[i] * x[i - 1]; sum += b[i] * y[i - 1]; }
#pragma unroll
  for (int i = ORDER - 1; i > 0; i--) { x[i] = x[i - 1]; y[i] = y[i - 1]; }
  x[0] = in; y[0] = sum;
  return sum;
}

// Registers of warp X1: glottis state + cached dimensions, filter histories.
struct X1Regs {
  GlottisState G;
  GlottisDims GD;
  // The two output filters (IirFilter::getOutputSample, IirFilter.cpp:47-65) in TRANSPOSED direct form II:
  //   y = a0 x + s1;   s_i <- a_i x + b_i y + s_{i+1};   s_N <- a_N x + b_N y
  // -- the same recursion as the reference's direct form (identical in exact arithmetic, zero initial state alike),
  // with N states instead of 2 N histories, no shifting, and an output that needs ONE multiply-add once the input is
  // there: the state updates wait for the next pass A, where this warp has time (its path behind the barrier is the
  // longest of the step).
  double toneS[4];            // skin-radiation filter (TdsModel.cpp:494-510)
  double chebS[8];            // output low-pass (Synthesizer.cpp:52)
  double pend[4];             // input / output of the last sample of the two filters, not yet in the states
  double prevFlow;
};
template <int ORDER>
__device__ __forceinline__ void df2t_update(const double *a, const double *b, double (&s)[ORDER], double x, double y) {
#pragma unroll
  for (int i = 1; i < ORDER; i++) s[i - 1] = fma(a[i], x, fma(b[i], y, s[i]));
  s[ORDER - 1] = fma(a[ORDER], x, b[ORDER] * y);
}
// pass A of warp X1: the filter states take in the sample emitted behind the last barrier
__device__ __forceinline__ void emit_update(const TdsParams &P, X1Regs &xr) {
  if (P.skin) df2t_update<4>(c_tone_a, c_tone_b, xr.toneS, xr.pend[0], xr.pend[1]);
  df2t_update<8>(P.cheb_a, P.cheb_b, xr.chebS, xr.pend[2], xr.pend[3]);
}
// Output of the step whose radiated flows sit in uext[]: radiated flow (TdsModel.cpp:687-705), pressure
// differentiator + Chebyshev low-pass + decimation (Synthesizer.cpp:614-627).  `n_done` = steps finished.
__device__ __forceinline__ void emit_output(const TdsParams &P, X1Regs &xr, double u_rad, double p25, int n_done, bool active,
                                            int tract, int n_out, double *audio, double *flow, uint32_t &bad) {
  double u = u_rad;
  const double tone = fma(c_tone_a[0], p25, xr.toneS[0]);
  if (P.skin) u += tone;
  if (!isfinite(u)) bad |= 1u;
  const double outP = (u - xr.prevFlow) * P.fs;
  xr.prevFlow = u;
  const double filt = fma(P.cheb_a[0], outP, xr.chebS[0]);
  xr.pend[0] = p25; xr.pend[1] = tone; xr.pend[2] = outP; xr.pend[3] = filt;
  if (active) {
    if (flow) flow[(size_t)tract * P.total_steps + (n_done - 1)] = u;
    if ((n_done % P.sr_factor) == 0) audio[(size_t)tract * n_out + (n_done / P.sr_factor - 1)] = filt * (0.004 / 32767);
  }
}

// ---- the kernel -----------------------------------------------------------------
struct KArgs {
  const int32_t *bounds;
  const afs_tube_frame *tubes;
  const double *glottis;
  double *audio, *flow;
  uint32_t *status;
  double *dump;
  const uint32_t *seeds;
  // noise sources, the constriction records of the pre-pass: EITHER one record per run of steps and tract
  // (tds_noise_runs_kernel: runs[run_off[tract] ..], sorted by their first step) ...
  const NzRun *runs;
  int run_items;            // pieces per tract: tract t's list starts at runs[t * run_items]
  const uint32_t *run_ctl;  // [0] next free record of the pool, [1] != 0: the pool was too small (status bit 1)
  // ... OR one per step (tds_noise_scan_kernel: [cta][step][4][32])
  uint32_t *nzc_i;
  double *nzc_f;
  // taps: [tract][probe][step], each nullable
  double *tap_p, *tap_in, *tap_out;
  // progress[cta] = simulation steps whose outputs are complete and visible system-wide, published every
  // `progress_every` steps (nullable): lets the host copy finished time slices while the kernel runs
  unsigned int *progress;
  int progress_every;
  // chunked synthesis (STATE kernels only): recurrent state of every tract, ST_WORDS doubles each (see "state in / state
  // out" below); state_in null = a fresh model, state_out null = the state is dropped at the end
  const double *state_in;
  double *state_out;
};


// Warp X1 keeps the constriction records of the step the others work on in shared memory (nzc_i / nzc_c).
struct NzFeed {
  uint32_t idx;      // runs mode: the tract's next record ...
  uint32_t next;     // ... and its first step
};
__device__ __forceinline__ void nz_put_run(TdsSmem &S, const int lane, const NzRun *r) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    S.nzc_i[k][lane] = __ldg(&r->slot[k]);
#pragma unroll
    for (int c = 0; c < 4; c++) S.nzc_c[k][c][lane] = __ldg(&r->coef[k][c]);
  }
}
// per-step records ([cta][step][4][32], tds_noise_scan_kernel): the downstream factor is a constant
struct NzcRec { uint32_t i[4]; double f[4]; };
__device__ __forceinline__ void nzc_fetch_issue(const uint32_t *nzc_i, const double *nzc_f, int total_steps, int n, NzcRec &R) {
  const int hw_lane = threadIdx.x & 31;
  const size_t base = ((size_t)blockIdx.x * (size_t)total_steps + (size_t)n) * 4;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    R.i[k] = __ldg(nzc_i + (base + k) * 32 + hw_lane);
    R.f[k] = __ldg(nzc_f + (base + k) * 32 + hw_lane);
  }
}
__device__ __forceinline__ void nzc_fetch_store(TdsSmem &S, const int lane, const NzcRec &R) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    S.nzc_i[k][lane] = R.i[k];
    S.nzc_c[k][0][lane] = R.f[k]; S.nzc_c[k][1][lane] = 0.0; S.nzc_c[k][2][lane] = 1.0; S.nzc_c[k][3][lane] = 0.0;
  }
}

// ---- state in / state out ---------------------------------------------------------
// Everything that survives a step (pressures, wall currents, branch currents and their rates, noise
// sources and random() stream, glottis displacements, the filters of the driver) lives in the registers of the warp
// that owns it and in the per-tract rows of shared memory.  A chunked utterance (afs_tds_stream_*) leaves the time loop
// after the update of its last step, writes both to global memory -- one opaque block of ST_WORDS doubles per tract,
// independent of how the tracts are packed into CTAs -- and the next chunk starts from it: its iteration 0 assembles
// rows from that state exactly as iteration T of an unchunked run would.
constexpr size_t ST_U32_AT = offsetof(TdsSmem, rng), ST_D1_AT = offsetof(TdsSmem, nzc_c), ST_END_AT = offsetof(TdsSmem, srec);
static_assert(ST_U32_AT % (8 * LANES) == 0 && (ST_D1_AT - ST_U32_AT) % (4 * LANES) == 0 && (ST_END_AT - ST_D1_AT) % (8 * LANES) == 0 &&
              ST_D1_AT % 8 == 0, "per-tract rows of TdsSmem");
constexpr int ST_ROWS_A = (int)(ST_U32_AT / (8 * LANES));               // [row][lane] doubles in front of the random() ring
constexpr int ST_ROWS_U = (int)((ST_D1_AT - ST_U32_AT) / (4 * LANES));  // 32-bit rows: random() ring, ring of sums, constriction slots
constexpr int ST_ROWS_B = (int)((ST_END_AT - ST_D1_AT) / (8 * LANES));  // doubles behind them
constexpr int ST_REG_WORDS = 56;                                        // registers of one warp role (the largest: X2)
constexpr int ST_PARITY_AT = 33;                                       // where X1 keeps the parity of the chunk's step count
constexpr int ST_SMEM_AT = NW * ST_REG_WORDS;
constexpr int ST_WORDS = ST_SMEM_AT + ST_ROWS_A + ST_ROWS_B + (ST_ROWS_U + 1) / 2;

template <bool LOAD>
__device__ __forceinline__ void st_word(double *w, int &k, double &v) {
  if (LOAD) v = w[k]; else w[k] = v;
  k++;
}
template <bool LOAD>
__device__ __forceinline__ void st_int(double *w, int &k, int &v) {
  if (LOAD) v = (int)w[k]; else w[k] = (double)v;
  k++;
}
template <bool LOAD, int N>
__device__ __forceinline__ void st_arr(double *w, int &k, double (&a)[N]) {
#pragma unroll
  for (int i = 0; i < N; i++) st_word<LOAD>(w, k, a[i]);
}
template <bool LOAD>
__device__ __forceinline__ void st_dregs(double *w, DRegs &r) {
  int k = 0;
  st_arr<LOAD>(w, k, r.p); st_arr<LOAD>(w, k, r.pr); st_arr<LOAD>(w, k, r.c); st_arr<LOAD>(w, k, r.x); st_arr<LOAD>(w, k, r.y); st_arr<LOAD>(w, k, r.h);
  st_word<LOAD>(w, k, r.L0); st_word<LOAD>(w, k, r.R0); st_word<LOAD>(w, k, r.I0); st_word<LOAD>(w, k, r.D0); st_word<LOAD>(w, k, r.gbf);
}
template <bool LOAD>
__device__ __forceinline__ void st_sregs(double *w, SRegs &r, unsigned long long &live) {
  int k = 0;
  st_arr<LOAD>(w, k, r.p); st_arr<LOAD>(w, k, r.pr); st_arr<LOAD>(w, k, r.c); st_arr<LOAD>(w, k, r.x);
  st_arr<LOAD>(w, k, r.lp); st_arr<LOAD>(w, k, r.lpr); st_arr<LOAD>(w, k, r.lc); st_arr<LOAD>(w, k, r.lx);
  st_word<LOAD>(w, k, r.D0); st_word<LOAD>(w, k, r.w84); st_word<LOAD>(w, k, r.e28); st_word<LOAD>(w, k, r.g28);
  double lv = __longlong_as_double((long long)live);  // (warp F: the noise sources whose amplitude is not zero)
  st_word<LOAD>(w, k, lv);
  if (LOAD) live = (unsigned long long)__double_as_longlong(lv);
}
template <bool LOAD>
__device__ __forceinline__ void st_x1regs(double *w, X1Regs &r, int &parity) {
  int k = 0;
  st_word<LOAD>(w, k, r.G.x0); st_word<LOAD>(w, k, r.G.x1); st_word<LOAD>(w, k, r.G.xp0); st_word<LOAD>(w, k, r.G.xp1);
  st_word<LOAD>(w, k, r.GD.cordLength); st_word<LOAD>(w, k, r.GD.th0); st_word<LOAD>(w, k, r.GD.th1); st_word<LOAD>(w, k, r.GD.ol0);
  st_word<LOAD>(w, k, r.GD.ol1); st_word<LOAD>(w, k, r.GD.cl0); st_word<LOAD>(w, k, r.GD.cl1); st_word<LOAD>(w, k, r.GD.mw0);
  st_word<LOAD>(w, k, r.GD.mw1); st_word<LOAD>(w, k, r.GD.mz0); st_word<LOAD>(w, k, r.GD.mz1); st_word<LOAD>(w, k, r.GD.Q);
  st_arr<LOAD>(w, k, r.toneS); st_arr<LOAD>(w, k, r.chebS); st_arr<LOAD>(w, k, r.pend);
  st_word<LOAD>(w, k, r.prevFlow);
  st_int<LOAD>(w, k, parity);  // parity of the chunk's step count: which half of the step-parity buffers is current
}
template <bool LOAD>
__device__ __forceinline__ void st_x2regs(double *w, RadRegs &rr, OptRegs &o, NoiseRegs &nr) {
  int k = 0;
  st_arr<LOAD>(w, k, rr.c); st_arr<LOAD>(w, k, rr.u); st_arr<LOAD>(w, k, rr.keep[0]); st_arr<LOAD>(w, k, rr.keep[1]);
  st_arr<LOAD>(w, k, o.x1); st_arr<LOAD>(w, k, o.y1); st_arr<LOAD>(w, k, o.x2); st_arr<LOAD>(w, k, o.y2); st_arr<LOAD>(w, k, o.xg); st_arr<LOAD>(w, k, o.yg);
  st_int<LOAD>(w, k, nr.ph); st_int<LOAD>(w, k, nr.head); st_int<LOAD>(w, k, nr.free); st_int<LOAD>(w, k, nr.cap);
  st_word<LOAD>(w, k, nr.lips); st_word<LOAD>(w, k, nr.nm93); st_word<LOAD>(w, k, nr.nm94); st_word<LOAD>(w, k, nr.asp); st_word<LOAD>(w, k, nr.gain);
}
static_assert(6 * KD + 5 <= ST_REG_WORDS && 4 * KS + 8 + 4 + 1 <= ST_REG_WORDS && 4 + 12 + 16 + 1 + 1 <= ST_REG_WORDS &&
              8 + 14 + 24 + 4 + 5 <= ST_REG_WORDS, "ST_REG_WORDS");
// The per-tract rows of shared memory (all warps of the CTA; `blk`: the tract's block).  Saving: active lanes only.
template <bool LOAD>
__device__ __noinline__ void st_smem(TdsSmem &S, double *blk, const int lane, const int warp, const bool on) {
  if (!on) return;
  double *d0 = reinterpret_cast<double *>(&S), *d1 = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(&S) + ST_D1_AT);
  uint32_t *u = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(&S) + ST_U32_AT);
  double *ga = blk + ST_SMEM_AT, *gb = ga + ST_ROWS_A;
  uint32_t *gu = reinterpret_cast<uint32_t *>(gb + ST_ROWS_B);
  for (int r = warp; r < ST_ROWS_A; r += NW) { if (LOAD) d0[r * LANES + lane] = ga[r]; else ga[r] = d0[r * LANES + lane]; }
  for (int r = warp; r < ST_ROWS_B; r += NW) { if (LOAD) d1[r * LANES + lane] = gb[r]; else gb[r] = d1[r * LANES + lane]; }
  for (int r = warp; r < ST_ROWS_U; r += NW) { if (LOAD) u[r * LANES + lane] = gu[r]; else gu[r] = u[r * LANES + lane]; }
}


#ifdef AFS_TDS_TIMING
#define AFS_TICK(k) do { const long long now__ = clock64(); tacc[k] += now__ - tlast; tlast = now__; } while (0)
#else
#define AFS_TICK(k) do { } while (0)
#endif

// Stages the control frames of interval `interval` (all warps; the caller puts barriers around it).
__device__ __noinline__ void stage_interval(TdsSmem &S, const int lane, const int warp, const afs_tube_frame *my_frames,
                                               const double *my_glottis, int interval) {
  for (int item = warp; item < 2 * AFS_NUM_PM_SECTIONS + 2; item += NW) {
    if (item < 2 * AFS_NUM_PM_SECTIONS) {
      const int f = item / AFS_NUM_PM_SECTIONS, j = item % AFS_NUM_PM_SECTIONS;
      const afs_tube_frame *fr = my_frames + interval + f;
      double a = fr->area_cm2[j];
      if (a < kMinArea) a = kMinArea;  // Tube::setPharynxMouthGeometry clamps on the way in
      S.fA[f][j][lane] = a;
      S.fLen[f][j][lane] = fr->length_cm[j];
    } else {
      const int f = item - 2 * AFS_NUM_PM_SECTIONS;
      double v = my_frames[interval + f].velum_area_cm2;
      if (v < kMinArea) v = kMinArea;  // Tube::getVelumOpening_cm2 returns the clamped nose[0] area
      S.velum[f][lane] = v;
      for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++)
        S.gctl[f][c][lane] = my_glottis[(size_t)(interval + f) * AFS_NUM_GLOTTIS_CONTROL + c];
    }
  }
}

// Taps (OPT kernels): which slots of this warp's segment hold a tapped section (bit k: slot k; bits 8, 9: sinus leaves).
__device__ __forceinline__ unsigned taps_of(const TdsParams &P, const int role, const int warp, const SCfg &c) {
  unsigned m = 0;
  for (int q = 0; q < P.n_probes; q++) {
    const int sec = (int)((P.probe_pack >> (8 * q)) & 255ull);
    if (role == 0) { const int k = sec - (23 + KD * warp); if (k >= 0 && k < KD) m |= 1u << k; }
    if (role == 1) {
      const int k = sec - c.base;
      if (k >= c.ks && k < KS) m |= 1u << k;
      if (c.leaves && (sec == c.leaf0 || sec == c.leaf0 + 1)) m |= 1u << (8 + sec - c.leaf0);
    }
  }
  return m;
}
// before pass A: flows of step n-1 (x still holds its solution); after it: the pressures the update produced
template <bool AFTER>
__device__ __forceinline__ void dwarp_taps(const TdsParams &P, TdsSmem &S, const int lane, const int w, const unsigned wc_,
                                           const DRegs &st, const StepCtx &X, const unsigned mask) {
  const double outLast = AFTER ? 0.0 : S.uext[wc_out0(wc_)][lane] + S.uext[wc_out1(wc_)][lane];
#pragma unroll
  for (int k = 0; k < KD; k++) {
    if (!((mask >> k) & 1u)) continue;
    const int sec = 23 + KD * w + k;
    if (AFTER) tap_section(P.n_probes, P.probe_pack, X.tap_p, nullptr, nullptr, X.tap_stride, sec, st.p[k], 0.0, 0.0);
    else tap_section(P.n_probes, P.probe_pack, nullptr, X.tap_in, X.tap_out, X.tap_stride, sec, 0.0, st.x[k], k < KD - 1 ? st.x[k + 1] : outLast);
  }
}
template <bool AFTER>
__device__ __forceinline__ void swarp_taps(const TdsParams &P, TdsSmem &S, const int lane, const unsigned wc_, const SCfg &c,
                                           const SRegs &st, const StepCtx &X, const unsigned mask) {
  const double outLast = AFTER ? 0.0 : S.uext[wc_out0(wc_)][lane];
  const int j0 = c.n2 ? 0 : 5;
#pragma unroll
  for (int k = 0; k < KS; k++) {
    if (!((mask >> k) & 1u)) continue;
    if (AFTER) { tap_section(P.n_probes, P.probe_pack, X.tap_p, nullptr, nullptr, X.tap_stride, c.base + k, st.p[k], 0.0, 0.0); continue; }
    double out = (k < KS - 1) ? st.x[k + 1] : outLast;
    if (k == j0) out += st.lx[0];
    if (k == j0 + 1) out += st.lx[1];
    tap_section(P.n_probes, P.probe_pack, nullptr, X.tap_in, X.tap_out, X.tap_stride, c.base + k, 0.0, st.x[k], out);
  }
#pragma unroll
  for (int j = 0; j < 2; j++) {
    if (!((mask >> (8 + j)) & 1u)) continue;
    if (AFTER) tap_section(P.n_probes, P.probe_pack, X.tap_p, nullptr, nullptr, X.tap_stride, c.leaf0 + j, st.lp[j], 0.0, 0.0);
    else tap_section(P.n_probes, P.probe_pack, nullptr, X.tap_in, X.tap_out, X.tap_stride, c.leaf0 + j, 0.0, st.lx[j], 0.0);
  }
}

// The time loop, one instantiation per warp role.  All roles execute the same sequence of CTA barriers.
//   ROLE 0: dynamic chain, 1: static chain, 2: X1 (glottis + output), 3: X2 (radiation loads; noise sources)
template <int ROLE, bool NOISE, bool OPT, bool STATE>
__device__ __forceinline__ void run_role(const TdsParams &P, TdsSmem &S, const KArgs &A, const int lane, const int warp,
                                         const bool hw_ok) {
  const int tract_raw = blockIdx.x * P.lanes_per_cta + lane;
  const bool active = hw_ok && lane < P.lanes_per_cta && tract_raw < P.n_tracts;
  const int tract = tract_raw < P.n_tracts ? tract_raw : P.n_tracts - 1;
  const afs_tube_frame *my_frames = A.tubes + (size_t)tract * P.n_frames;
  const double *my_glottis = A.glottis + (size_t)tract * P.n_frames * AFS_NUM_GLOTTIS_CONTROL;
  const int n_out = P.total_steps / P.sr_factor;
  const int T = P.total_steps;
  uint32_t bad = 0;

  DRegs dst;
  SRegs sst;
  RadRegs rr;
  OptRegs orr;
  X1Regs xr;
  NoiseRegs nr;
  NzFeed feed{0u, 0x7fffffffu};
  int next_mark = (ROLE == 3 && A.progress && A.progress_every > 0) ? (A.progress_every < T ? A.progress_every : T) : -1;  // (warp X2)
  unsigned long long nz_live = 0;  // (warp F) sources whose amplitude is not zero
  nr.ph = 12; nr.head = 0; nr.free = 0; nr.cap = NZ_RING; nr.lips = 0.0; nr.nm93 = nr.nm94 = 0.0; nr.asp = nan(""); nr.gain = 0.0;
  const SCfg scfg = scfg_of(warp);
  const unsigned wcfg = wcfg_make(warp);
  const unsigned tapmask = (OPT && P.n_probes > 0 && ROLE <= 1) ? taps_of(P, ROLE, warp, scfg) : 0u;
  if (ROLE == 0) {
#pragma unroll
    for (int k = 0; k < KD; k++) { dst.p[k] = dst.pr[k] = dst.c[k] = dst.x[k] = dst.y[k] = dst.h[k] = 0.0; }
    dst.L0 = dst.R0 = dst.I0 = dst.D0 = 0.0; dst.gbf = 0.0;
  }
  if (ROLE == 1) {
#pragma unroll
    for (int k = 0; k < KS; k++) { sst.p[k] = sst.pr[k] = sst.c[k] = sst.x[k] = 0.0; }
#pragma unroll
    for (int j = 0; j < 2; j++) { sst.lp[j] = sst.lpr[j] = sst.lc[j] = sst.lx[j] = 0.0; }
    sst.D0 = 0.0; sst.w84 = sst.e28 = sst.g28 = 0.0;
    swarp_build(P, S, lane, warp);
  }
  // chunked synthesis: this chunk continues from the state the last one left (STATE kernels; see "state in / state out")
  const bool resume = STATE && A.state_in != nullptr;
  double *st_in = resume ? const_cast<double *>(A.state_in) + (size_t)tract * ST_WORDS + warp * ST_REG_WORDS : nullptr;
  int st_parity = 0;
  if (ROLE == 3) {
#pragma unroll
    for (int j = 0; j < 4; j++) { rr.c[j] = rr.u[j] = 0.0; orr.x1[j] = orr.y1[j] = orr.x2[j] = orr.y2[j] = orr.xg[j] = orr.yg[j] = 0.0; }
    if (STATE) {
#pragma unroll
      for (int j = 0; j < 7; j++) rr.keep[0][j] = rr.keep[1][j] = 0.0;
    }
    if (NOISE && !resume) {
      rng_seed(S, lane, A.seeds ? A.seeds[tract] : 1u, nr.ph);
      nr.cap = nr.free = NZ_RING;
      nz_refill(S, lane, nr, NZ_RING);  // (42 sums; the ring is never emptier than that when a step begins)
    }
  }
  if (ROLE == 2) {
    xr.G = GlottisState{0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 4; i++) { xr.toneS[i] = 0.0; xr.pend[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 8; i++) xr.chebS[i] = 0.0;
    xr.prevFlow = 0.0;
  }
  if (STATE && resume) {
    if (ROLE == 0) st_dregs<true>(st_in, dst);
    if (ROLE == 1) st_sregs<true>(st_in, sst, nz_live);
    if (ROLE == 2) {
      st_x1regs<true>(st_in, xr, st_parity);
      emit_update(P, xr);  // (the last sample of the previous chunk: its filter states were left for the next pass A)
    }
    if (ROLE == 3) {
      st_x2regs<true>(st_in, rr, orr, nr);
      // the history parts X2 prepared for the step that is now step 0 sit in the half of the previous chunk's parity
      const int par = (int)A.state_in[(size_t)tract * ST_WORDS + W_X1 * ST_REG_WORDS + ST_PARITY_AT];
      if (par) { S.tvh[0][lane] = S.tvh[1][lane]; S.kenth[0][lane] = S.kenth[1][lane]; }
    }
  }
  cta_sync();  // static boundary records / pivots visible
  if (warp == W_T0) reduced_constants(S, lane);

  int interval = 0, i0 = 0, i1 = 0;
  double inv_len = 1.0;
  if (T > 0) {
    i0 = A.bounds[0]; i1 = A.bounds[1];
    inv_len = 1.0 / (double)(i1 - i0);
    stage_interval(S, lane, warp, my_frames, my_glottis, 0);
    cta_sync();
    if (ROLE == 2 && NOISE) {  // constrictions of step 0
      if (A.runs) {
        const NzRun *rec = A.runs + (size_t)tract * (size_t)A.run_items;
        nz_put_run(S, lane, rec);
        S.nzfed[lane] = nz_fed(S, lane);
        feed.next = rec->next;
        feed.idx = rec->next_idx;
      } else {
        NzcRec rec;
        nzc_fetch_issue(A.nzc_i, A.nzc_f, T, 0, rec);
        nzc_fetch_store(S, lane, rec);
      }
    }
    if (ROLE == 2) {  // glottis geometry of step 0
      double ctl[AFS_NUM_GLOTTIS_CONTROL], len[2], ar[2];
      if (STATE && resume) {  // ... is what the last step of the previous chunk prepared (same control frame, ratio 1)
        len[0] = S.glLen[st_parity][0][lane]; len[1] = S.glLen[st_parity][1][lane];
        ar[0] = S.glArea[st_parity][0][lane]; ar[1] = S.glArea[st_parity][1][lane];
      } else {
        for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++) ctl[c] = S.gctl[0][c][lane];
        glottis_geometry(P, ctl, xr.G, len, ar, xr.GD);
      }
      if (NOISE) {
        double fa, fc;
        nz_factors(ar[1], 1.0, fa, fc);
        S.nzk0[0][lane] = fa;
        S.nzk0[1][lane] = nz_down(S, lane, 0, 0.0);
      }
      S.glLen[0][0][lane] = len[0]; S.glLen[0][1][lane] = len[1];
      S.glArea[0][0][lane] = ar[0]; S.glArea[0][1][lane] = ar[1];
    }
    cta_sync();
  }
#ifdef AFS_TDS_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
  long long sw_acc[4] = {0, 0, 0, 0};
#endif

  for (int n = 0; n <= T && T > 0; n++) {
    StepCtx X;
    X.n = n; X.upd = n > 0; X.rows = n < T;
    if (X.rows && n == i1) {  // control frames of the next interval (every warp is past its last read of the old ones)
      interval++;
      i0 = A.bounds[interval]; i1 = A.bounds[interval + 1];
      inv_len = 1.0 / (double)(i1 - i0);
      stage_interval(S, lane, warp, my_frames, my_glottis, interval);
      cta_sync();
    }
    X.ratio = (double)(n - i0) * inv_len;  // (1 ulp from the reference's i / N; irrelevant at the 1e-6 bar)
    X.ratio1 = 1.0 - X.ratio;
    X.dumping = A.dump != nullptr && P.dump_step == n && blockIdx.x == 0;
    X.tap_p = X.tap_in = X.tap_out = nullptr; X.tap_stride = (size_t)T;
    if (OPT && P.n_probes > 0 && active && n > 0) {  // values of step n-1 (0-based), written by the update of pass A
      const size_t at = (size_t)tract * P.n_probes * (size_t)T + (size_t)(n - 1);
      if (A.tap_p) X.tap_p = A.tap_p + at;
      if (A.tap_in) X.tap_in = A.tap_in + at;
      if (A.tap_out) X.tap_out = A.tap_out + at;
    }
    AFS_TICK(0);
    // ============ pass A: update (step n-1), elements + rows (step n) ============
    if (ROLE == 1 && NOISE && X.rows) {  // (see noise_decay)
      if (warp == W_F) noise_decay(P.nz_F, S, lane, nz_live, A.runs != nullptr);
      else if (warp == W_N1 || warp == W_T2) {  // the two static chain warps with the most slack after F
        const int k = warp == W_N1 ? 2 : 3;
        S.nztongue[k - 1][lane] = noise_slot(NzConst{P.nz_F, P.nz_x2000, P.dt}, S, lane, X.ratio, X.ratio1, k, 0.0);
        bar_arrive(4, NZ_BAR4);
      }
    }
    if (OPT && tapmask && X.upd) {  // taps: flows of step n-1
      if (ROLE == 0) dwarp_taps<false>(P, S, lane, warp, wcfg, dst, X, tapmask);
      if (ROLE == 1) swarp_taps<false>(P, S, lane, wcfg, scfg, sst, X, tapmask);
    }
    // (state dump requested for this step: update and rows in two separate passes, the dump in between)
    if (ROLE == 0 && X.upd && X.rows && !X.dumping) dwarp_pass_a<true, OPT>(P, S, lane, warp, wcfg, dst, X);
    else for (int half = 0; half < (X.dumping ? 2 : 1); half++) {
      StepCtx Y = X;
      if (X.dumping) { Y.upd = X.upd && half == 0; Y.rows = X.rows && half == 1; }
      if (ROLE == 0) dwarp_pass_a<false, OPT>(P, S, lane, warp, wcfg, dst, Y);
      if (ROLE == 1) swarp_pass_a<OPT>(P, S, lane, warp, wcfg, scfg, sst, Y);
      if (ROLE == 3 && Y.upd) radiation_update(P, lane, rr, nr, Y, A.dump);
      if (X.dumping && half == 0) {
        if (ROLE == 0) dwarp_dump(P, S, lane, warp, dst, A.dump);
        if (ROLE == 1) swarp_dump(P, S, lane, scfg, sst, A.dump);
      }
    }
    if (OPT && tapmask && X.upd) {  // taps: pressures of step n-1
      if (ROLE == 0) dwarp_taps<true>(P, S, lane, warp, wcfg, dst, X, tapmask);
      if (ROLE == 1) swarp_taps<true>(P, S, lane, wcfg, scfg, sst, X, tapmask);
    }
    if (ROLE == 2 && n > 1) emit_update(P, xr);  // (the sample emitted in iteration n-1)
    if (ROLE == 2 && NOISE && X.rows) {
      S.nztongue[0][lane] = noise_slot(NzConst{P.nz_F, P.nz_x2000, P.dt}, S, lane, X.ratio, X.ratio1, 1, 0.0);
      bar_arrive(4, NZ_BAR4);
    }
    if (ROLE == 3 && NOISE && X.rows) {  // the constrictions' share of the noise sources of step n
      // (the control-frame ratio only enters smooth quantities here; the decisions that need the reference's own
      //  rounding of it were made by the pre-pass)
      const double asp = X.ratio1 * S.gctl[0][G_ASP][lane] + X.ratio * S.gctl[1][G_ASP][lane];
      if (asp != nr.asp) {  // vocal-fold source gain 10^(dB/20) (TdsModel.cpp:1449): constant while the aspiration strength is
        nr.asp = asp;
        nr.gain = 0.5e-7 * exp(asp * (2.302585092994045684 / 20.0));
      }
      const unsigned long long draw0 = noise_slot(NzConst{P.nz_F, P.nz_x2000, P.dt}, S, lane, X.ratio, X.ratio1, 0, nr.gain);
      const int used = noise_stage(S, lane, draw0
#ifdef AFS_TDS_TIMING
                                   , sw_acc
#endif
                                   );
      nr.head += used; nr.free += used;
      if (nr.head >= nr.cap) nr.head -= nr.cap;
      // new sums for the entries just consumed: X2 waits for the dynamic chain warps at the end of pass A anyway (in
      // the reduced-system phase, where this used to be, it made X2 the warp the others wait for)
      nz_refill(S, lane, nr, NZ_RING);
    }
    AFS_TICK(1);
    cta_sync();
    AFS_TICK(2);
    double x1p[4] = {0.0, 0.0, 0.0, 0.0};
    if (ROLE == 2) {  // everything X1 reads from the other warps, then it is on its own until the end of the step
      const double u_rad = S.uext[UX_RADM][lane] + S.uext[UX_RADN][lane];
      for (int c = 0; c < 4; c++) x1p[c] = S.p4[c][lane];
      if (NOISE && n + 1 < T) {  // constriction records of step n+1 (X2 is past its reads of step n's; it looks at
        if (!A.runs) {  // them again behind barrier 1, hence before the arrive)
          NzcRec nzrec;
          nzc_fetch_issue(A.nzc_i, A.nzc_f, T, n + 1, nzrec);
          nzc_fetch_store(S, lane, nzrec);
        }
        else if ((uint32_t)(n + 1) == feed.next) {
          const NzRun *rec = A.runs + feed.idx;
          nz_put_run(S, lane, rec);
          S.nzfed[lane] = nz_fed(S, lane);
          feed.next = __ldg(&rec->next);
          feed.idx = __ldg(&rec->next_idx);
        }
      }
      if (X.rows) bar_arrive(1, TDS_THREADS);
      if (n > 0) emit_output(P, xr, u_rad, x1p[3], n, active, tract, n_out, A.audio, A.flow, bad);
    }
    if (!X.rows) break;
    // ============ first rows + elimination of the segments || glottis ============
    if (ROLE == 0) dwarp_eliminate<NOISE, OPT>(P, S, lane, warp, wcfg, dst, n & 1);
    if (ROLE == 1) {
      const double psrc = X.ratio1 * S.gctl[0][G_PRESSURE][lane] + X.ratio * S.gctl[1][G_PRESSURE][lane];
#ifdef AFS_TDS_TIMING
      swarp_eliminate(P, S, lane, warp, wcfg, scfg, sst, psrc, sw_acc);
#else
      swarp_eliminate(P, S, lane, warp, wcfg, scfg, sst, psrc);
#endif
    }
    if (ROLE == 3 && OPT && (P.transvelar || P.var_loss)) {
      // this step's filter outputs (same expression as their consumers use), then the history parts of the next step
      const int par = n & 1;
      if (P.transvelar) {
        const double in1 = S.tvp[0][lane], in2 = S.tvp[1][lane];
        const double h = S.tvh[par][lane];
        // (split the sum of the two filters: y1 = a0 x1 + h1, y2 = a0' x2 + h2 with h = h1 + h2 kept separately)
        const double h1 = iir_history(c_tone_a, c_tone_b, orr.x1, orr.y1), h2 = iir_history(c_tv2_a, c_tone_b, orr.x2, orr.y2);
        (void)h;
        iir_push(orr.x1, orr.y1, in1, fma(c_tone_a[0], in1, h1));
        iir_push(orr.x2, orr.y2, in2, fma(c_tv2_a[0], in2, h2));
        S.tvh[par ^ 1][lane] = iir_history(c_tone_a, c_tone_b, orr.x1, orr.y1) + iir_history(c_tv2_a, c_tone_b, orr.x2, orr.y2);
      }
      if (P.var_loss) {
        const double in = S.p4[0][lane] - S.p4[3][lane];
        const double hg = iir_history(P.tg_a, P.tg_b, orr.xg, orr.yg);
        iir_push(orr.xg, orr.yg, in, fma(P.tg_a[0], in, hg));
        S.kenth[par ^ 1][lane] = iir_history(P.tg_a, P.tg_b, orr.xg, orr.yg);
      }
    }
    if (ROLE == 3) {
      // sample of the lips dipole source (amplitude 0 while it is silent; it may be either warp's this step)
      if (NOISE) nr.lips = S.nout[AFS_NUM_PM_SECTIONS][lane] * S.namp[AFS_NUM_PM_SECTIONS][lane];
      radiation_eliminate(P, S, lane, W_M3, 0, rr.c[0], rr.c[1], nr.lips, rr.keep[0]);
      radiation_eliminate(P, S, lane, W_N2, 1, rr.c[2], rr.c[3], nr.lips, rr.keep[1]);
    }
    if (ROLE == 2) {
      // incTime for this step (pressures of the previous step), then the geometry step n+1 will use
      double ctl[AFS_NUM_GLOTTIS_CONTROL], len[2], ar[2];
      const double (&p)[4] = x1p;
      for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++) ctl[c] = X.ratio1 * S.gctl[0][c][lane] + X.ratio * S.gctl[1][c][lane];
      // (xr.GD: the dimensions for the current displacement and control values are what the geometry call of the last
      //  step computed -- same inputs, also across a change of control frames, where the old interval ends at ratio 1
      //  on the frame the new one starts from)
      glottis_inc_time(P, ctl, xr.G, p, xr.GD);
      const double rn = (double)(n + 1 - i0) * inv_len, rn1 = 1.0 - rn;
      for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++) ctl[c] = rn1 * S.gctl[0][c][lane] + rn * S.gctl[1][c][lane];
      glottis_geometry(P, ctl, xr.G, len, ar, xr.GD);
      if (NOISE) {  // vocal-fold noise source of step n+1: what does not depend on the flow (noise_stage)
        double fa, fc;
        nz_factors(ar[1], 1.0, fa, fc);
        S.nzk0[0][lane] = fa;
        S.nzk0[1][lane] = nz_down(S, lane, 0, n + 1 == i1 ? 0.0 : rn);  // (a new interval starts with new records)
      }
      const int par = (n + 1) & 1;
      S.glLen[par][0][lane] = len[0]; S.glLen[par][1][lane] = len[1];
      S.glArea[par][0][lane] = ar[0]; S.glArea[par][1][lane] = ar[1];
    }
    AFS_TICK(3);
    if (ROLE != 2) {
      bar_sync(1, TDS_THREADS);
      // ============ reduced system over the separators ============
      // (the three warps of the reduced system meet on their own barrier; the twelve others only wait for the result)
      if (warp == W_T0 || warp == W_T1 || warp == W_T2) {
        const int chain = warp == W_T0 ? 0 : warp == W_T1 ? 1 : 2;
        reduced_solve(S, lane, chain);
      }
      AFS_TICK(4);
      bar_sync(2, TDS_THREADS - 32);
    }
    AFS_TICK(5);
    // ============ back substitution inside the segments ============
    if (ROLE == 0) dwarp_back<NOISE>(P, S, lane, warp, wcfg, dst);
    if (ROLE == 1) swarp_back<NOISE>(P, S, lane, warp, wcfg, scfg, sst);
    if (ROLE == 3) radiation_back<NOISE>(P, S, lane, rr, nr);
    AFS_TICK(6);
    cta_sync();
    AFS_TICK(7);
    // Progress mark: the samples of the steps 1..n are written (warp X1, before the barrier above).  Published by warp
    // X2 -- X1's path is the longest of the step, X2 has time here; the barrier orders X1's stores before this fence.
    if (ROLE == 3 && n == next_mark) {
      next_mark = next_mark + A.progress_every < T ? next_mark + A.progress_every : T;
      if ((threadIdx.x & 31) == 0) {
        __threadfence_system();
        *(volatile unsigned int *)(A.progress + blockIdx.x) = (unsigned int)n;
      }
    }
  }
  if (A.progress && T > 0) {  // the last step's samples (the loop is left before its final barrier)
    cta_sync();
    if (ROLE == 3 && (threadIdx.x & 31) == 0) {
      __threadfence_system();
      *(volatile unsigned int *)(A.progress + blockIdx.x) = (unsigned int)T;
    }
  }
#ifdef AFS_TDS_TIMING
  if (A.dump && blockIdx.x == 0 && lane == 0) for (int k = 0; k < 8; k++) A.dump[warp * 8 + k] = (double)tacc[k];
  if ((ROLE == 1 || ROLE == 3) && A.dump && blockIdx.x == 0 && lane == 0) for (int k = 0; k < 4; k++) A.dump[136 + warp * 4 + k] = (double)sw_acc[k];
  // (pass-A cycles of the two auxiliary warps of the first 136 CTAs: how evenly the noise sources load the CTAs)
  if (ROLE == 2 && A.dump && blockIdx.x < 136 && lane == 0) A.dump[200 + blockIdx.x] = (double)tacc[1];
  if (ROLE == 3 && A.dump && blockIdx.x < 136 && lane == 0) {  // (whole time loop of the CTA)
    long long all = 0;
    for (int k = 0; k < 8; k++) all += tacc[k];
    A.dump[336 + blockIdx.x] = (double)all;
  }
#endif
  if (STATE && A.state_out && T > 0) {
    // (shared memory: nothing is written behind the barrier the loop was left at)
    double *blk = A.state_out + (size_t)tract * ST_WORDS, *st_out = blk + warp * ST_REG_WORDS;
    if (active) {
      if (ROLE == 0) st_dregs<false>(st_out, dst);
      if (ROLE == 1) st_sregs<false>(st_out, sst, nz_live);
      if (ROLE == 2) { st_parity = T & 1; st_x1regs<false>(st_out, xr, st_parity); }
      if (ROLE == 3) st_x2regs<false>(st_out, rr, orr, nr);
    }
    st_smem<false>(S, blk, lane, warp, active);
  }
  if (ROLE == 2 && NOISE && A.run_ctl && A.run_ctl[1]) bad |= 2u;  // the noise pre-pass ran out of record space
  if (ROLE == 2 && active && A.status) A.status[tract] = bad;
}

// OPT = transvelar coupling or variable entrance loss switched on; NOISE = generate_noise_sources: a separate instantiation, so that the (large) noise-source code of warp X2 does not
// sit in the instruction stream of the noise-free kernel.
template <bool NOISE, bool OPT, bool STATE>
__global__ void __launch_bounds__(TDS_THREADS, 1)
tds_batch_kernel(const TdsParams P, const KArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TdsSmem &S = *reinterpret_cast<TdsSmem *>(smem_raw);
  const int hw_lane = threadIdx.x & 31;
  const int lane = hw_lane < LANES ? hw_lane : LANES - 1;
  const int pw_ = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  // role id: the four warps of one SM sub-partition (physical id % 4) run the SAME code (instruction fetch locality)
  const int warp = (pw_ & 3) * 4 + (pw_ >> 2);

  // ---- prologue: static records, zero state ------------------------------------
  for (int s = threadIdx.x; s < NSX; s += TDS_THREADS) {
    SRec r{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (s < NS && !is_dyn(s)) {
      const StaticSection q = static_section(s);
      double L, R, invr, circ, surface;
      if (s >= SINUS0) {  // Helmholtz resonators (TdsModel.cpp:751-757, 808-811)
        L = kRho * (q.length / q.area);
        R = (8.0 * kMu * kPi * q.length) / (q.area * q.area);
        invr = 1.0 / sqrt(q.area / kPi);
        surface = P.sinus_surface[s - SINUS0];
      } else {
        section_elements(q.area, q.length, false, L, R, invr, circ);
        surface = circ * q.length;
      }
      if (surface < kMinArea) surface = kMinArea;
      const double C = q.volume / (kRho * kC * kC);
      const double alpha = P.soft_walls ? surface * P.kalpha[wall_type(s)] : 0.0;
      r.L = L; r.R1 = R; r.R0 = R; r.invr = invr; r.alpha = alpha;
      r.E = P.dtth / (C + alpha);
      if (s == FOSSA0 && !P.fossa) r.R0 = 8.0 * kMu * q.length * kPi / (kMinArea * kMinArea);  // TdsModel.cpp:883-889
    }
    S.srec[s] = r;
  }
  if (threadIdx.x < 12) {
    const int t = threadIdx.x / 3, j = threadIdx.x % 3;
    double v = 0.0;
    if (t < 3) v = P.kalpha[t] * (j == 0 ? P.c1[t] : j == 1 ? P.c2[t] : P.c3[t]);
    S.wallc[t][j] = v;
  }
  for (int i = warp; i < NS * 3; i += NW) (&S.wall[0][0])[i][lane] = 0.0;
  for (int i = warp; i < 8 * KD; i += NW) { S.Ed[i][lane] = 0.0; S.Al[i][lane] = 0.0; }
  for (int i = warp; i < 42; i += NW) {
    S.nmag[i][lane] = 0.0;
    if (i < 41) { S.namp[i][lane] = 0.0; S.ncut[i][lane] = 3000.0; S.nout[i][lane] = 0.0; }
  }
  {
    const int w = warp;
    if (w < NSEG) {
      S.bE[w][lane] = 0.0; S.bL[w][lane] = 0.0; S.bR[w][lane] = 0.0; S.bI[w][lane] = 0.0; S.bG[w][lane] = 0.0;
      S.bD[w][lane] = 0.0; S.sideDD[w][lane] = 0.0; S.sideDR[w][lane] = 0.0; S.sideF[w][lane] = 0.0;
    }
    S.csib[w & 3][lane] = 0.0;
    S.sepD[w][lane] = 1.0; S.sepR[w][lane] = 0.0;  // NSEP == NW
  }
  if (warp == 0) {
    S.nzk0[0][lane] = 0.0; S.nzk0[1][lane] = 0.0; S.nzdecay[lane] = 0; S.nzfed[lane] = 0; S.nztongue[0][lane] = 0; S.nztongue[1][lane] = 0; S.nztongue[2][lane] = 0;
    S.nzlip[lane] = 0.0;
    for (int i = 0; i < NUX; i++) S.uext[i][lane] = 0.0;
    for (int i = 0; i < 5; i++) S.lf[i][lane] = 0.0;
    for (int i = 0; i < 4; i++) { S.p4[i][lane] = 0.0; (&S.radc[0][0])[i][lane] = 0.0; }
    for (int i = 0; i < 2; i++) { S.tvp[i][lane] = 0.0; S.tvh[i][lane] = 0.0; S.kenth[i][lane] = 0.0; }
    S.tvflow[lane] = 0.0;
  }
  __syncthreads();
  if (STATE && A.state_in) {  // chunked synthesis: the per-tract rows as the previous chunk left them
    const int tr = blockIdx.x * P.lanes_per_cta + lane;
    st_smem<true>(S, const_cast<double *>(A.state_in) + (size_t)(tr < P.n_tracts ? tr : P.n_tracts - 1) * ST_WORDS, lane, warp, true);
    __syncthreads();
  }

  switch (warp) {
    case W_T0: case W_T1: case W_T2: case W_N1: case W_N2: case W_F: run_role<1, NOISE, OPT, STATE>(P, S, A, lane, warp, hw_lane < LANES); break;
    case W_X1: run_role<2, NOISE, OPT, STATE>(P, S, A, lane, warp, hw_lane < LANES); break;
    case W_X2: run_role<3, NOISE, OPT, STATE>(P, S, A, lane, warp, hw_lane < LANES); break;
    default: run_role<0, NOISE, OPT, STATE>(P, S, A, lane, warp, hw_lane < LANES); break;
  }
}
static_assert(NSEP == NW, "prologue zeroing assumes NSEP == NW");

// ==================================================================================
// Noise-source pre-pass: TdsModel::calcNoiseSources, state-independent part (TdsModel.cpp:1188-1430)
// ==================================================================================
// Which sections form a constriction and where its obstacle sits depends on the GEOMETRY of a step only, not on the
// acoustic state (inside the time loop the scan once cost one warp 27 k cycles per step, with the fifteen others
// waiting for it).  Two kernels, chosen by the spacing of the control frames:
//
// tds_noise_runs_kernel -- control frames many steps apart (an utterance = a few frames).  Between two frames every
// quantity the scan compares is linear in the step, so its outcome is piecewise constant; the kernel finds the run
// boundaries of every tract by BISECTION over the steps of an interval (two steps with the same ScanSig enclose steps
// with that signature) and writes one NzRun record per run: a few hundred scans per utterance instead of 44 100, a
// few KB of records instead of 2 MB.  One thread per (tract, piece of an interval): the bisection is sequential, long
// intervals are cut into pieces.  The records go into one pool as linked lists (NzRun), in a single pass.
//
// tds_noise_scan_kernel -- control frames a few steps apart (one per output sample, Synthesizer::playTargetSequence
// style): every step is scanned, (tract, step) pairs are independent, same CTA <-> tract mapping as tds_batch_kernel,
// the warps of a CTA take the steps of an interval round robin; 48 B per tract-step in [cta][step][4][32] layout.
constexpr int RUN_NL = 33;  // 32 tracts per CTA; rows padded: the staging below transposes
using RunFrames = ScanFrames<RUN_NL>;

// downF(r) = (c0 + r c1) / (c2 + r c3) of a constriction slot for the interval staged in S (see NoiseCons::branch)
template <class FR>
__device__ __forceinline__ void nz_run_coef(const FR &S, const int lane, const NoiseCons &c, double (&coef)[4]) {
  coef[0] = 0.0; coef[1] = 0.0; coef[2] = 1.0; coef[3] = 0.0;
  if (!c.valid) return;
  if (c.branch == 2) { coef[0] = 0.5; return; }
  if (c.branch == 3) { coef[0] = 1.0; return; }
  double p0 = 0.0, p1 = 0.0;  // position of the obstacle's section in the two control frames
  for (int i = 0; i < c.up; i++) { p0 += S.fLen[0][i][lane]; p1 += S.fLen[1][i][lane]; }
  const double x0 = c.branch == 1 ? S.inc[0][lane] : 1.5, x1 = c.branch == 1 ? S.inc[1][lane] : 1.5;
  coef[0] = x0 - p0; coef[1] = (x1 - p1) - coef[0];
  coef[2] = S.fLen[0][c.up][lane]; coef[3] = S.fLen[1][c.up][lane] - coef[2];
}

// One work item of the runs pre-pass: the steps [lo, hi] of control-frame interval iv (long intervals are cut into
// pieces so that the bisections of one utterance spread over many warps; a piece starts with a record of its own, which
// costs nothing but its 160 bytes); next_lo = first step of the next item (0x7fffffff: none).
struct NzItem { int32_t iv, lo, hi, next_lo; };

constexpr int RUN_PIECES = 16;  // pieces per interval = warps per CTA of the runs kernel
__global__ void __launch_bounds__(32 * RUN_PIECES)
tds_noise_runs_kernel(const int n_tracts, const int n_frames, const int n_items, const int32_t *__restrict__ bounds,
                      const afs_tube_frame *__restrict__ tubes, const NzItem *__restrict__ items,
                      const int32_t *__restrict__ iv_first /* per interval: its first item, then the total */,
                      NzRun *__restrict__ runs, const uint32_t pool_records, uint32_t *__restrict__ ctl) {
  // CTA = 32 tracts x one control-frame interval (staged once); warp = one piece of the interval
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RunFrames &S = *reinterpret_cast<RunFrames *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tract0 = blockIdx.y * 32, tract = tract0 + lane;
  const bool active = tract < n_tracts;
  const int iv = blockIdx.x, item_idx = iv_first[iv] + warp;
  const bool have_item = item_idx < iv_first[iv + 1];
  const int i0 = bounds[iv], i1 = bounds[iv + 1];
  // the two control frames of the interval, all 32 tracts: coalesced reads, transposed into [item][tract]
  for (int t = warp; t < 32; t += RUN_PIECES) {
    const int tr = tract0 + t < n_tracts ? tract0 + t : n_tracts - 1;
    for (int f = 0; f < 2; f++) {
      const afs_tube_frame *fr = tubes + (size_t)tr * n_frames + iv + f;
      for (int j = lane; j < AFS_NUM_PM_SECTIONS; j += 32) {
        const double a = fr->area_cm2[j];
        S.fA[f][j][t] = a < kMinArea ? kMinArea : a;  // Tube::setPharynxMouthGeometry clamps on the way in
        S.fLen[f][j][t] = fr->length_cm[j];
        S.fLat[f][j][t] = fr->laterality[j];
        if (f == 0) S.art[j][t] = (unsigned char)fr->articulator[j];
      }
      if (lane == 0) S.inc[f][t] = fr->incisor_pos_cm;
    }
  }
  __syncthreads();
  if (!have_item) return;
  if (!active) return;
  const NzItem item = items[item_idx];
  const uint32_t first = (uint32_t)((size_t)tract * (size_t)n_items + (size_t)item_idx);  // fixed place of the piece's first record
  const uint32_t after = item_idx + 1 < n_items ? first + 1 : 0u;                        // ... and of the next piece's
  NzRun *prev = nullptr;
  bool full = false;
  NoiseCons C[4];
  auto scan = [&](int n, ScanSig &sg) {
    const double re = __ddiv_rn((double)(n - i0), (double)(i1 - i0)), re1 = __dsub_rn(1.0, re);  // the reference's ratio
    noise_scan(S, lane, re, re1, C, &sg);
  };
  auto emit = [&](int n) {  // C holds the scan of step n
    if (full) return;
    uint32_t idx = first;
    if (prev) {
      idx = atomicAdd(ctl, 1u);
      if (idx >= pool_records) { full = true; atomicExch(ctl + 1, 1u); return; }  // (reported as status bit 1 of every tract)
      prev->next = (uint32_t)n; prev->next_idx = idx;
    }
    NzRun &R = runs[idx];
    R.start = (uint32_t)n; R.next = (uint32_t)item.next_lo; R.next_idx = after; R.pad = 0;
    uint32_t w4[4];
    nzc_pack4(C, w4);
    for (int k = 0; k < 4; k++) {
      R.slot[k] = w4[k];
      double cf[4];
      nz_run_coef(S, lane, C[k], cf);
      for (int c = 0; c < 4; c++) R.coef[k][c] = cf[c];
    }
    prev = &R;
  };
  int lo = item.lo;
  const int last = item.hi;
  ScanSig sg_lo, sg;
  scan(lo, sg_lo);
  emit(lo);  // (always: the coefficients belong to the interval, and the piece starts here)
  if (last > lo) {
    int hi_stack[34];
    ScanSig sg_stack[34];
    int sp = 0;
    scan(last, sg);
    hi_stack[0] = last; sg_stack[0] = sg; sp = 1;
    while (sp > 0) {
      const int hi = hi_stack[sp - 1];
      const ScanSig sg_hi = sg_stack[sp - 1];
      if (sg_hi == sg_lo) { lo = hi; sp--; continue; }        // nothing changes up to hi
      if (hi == lo + 1) {                                     // the change is at hi
        scan(hi, sg);
        emit(hi);
        lo = hi; sg_lo = sg_hi; sp--;
        continue;
      }
      const int mid = lo + (hi - lo) / 2;
      scan(mid, sg);
      hi_stack[sp] = mid; sg_stack[sp] = sg; sp++;
    }
  }
}

#ifndef AFS_SCAN_THREADS
#define AFS_SCAN_THREADS 640  // 20 warps at <= 102 registers: the scan is latency-bound, every warp more is throughput
#endif
constexpr int SCAN_THREADS = AFS_SCAN_THREADS, SCAN_WARPS = SCAN_THREADS / 32;
using StepFrames = ScanFrames<LANES>;
__global__ void __launch_bounds__(SCAN_THREADS, 1)
tds_noise_scan_kernel(const TdsParams P, const KArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepFrames &S = *reinterpret_cast<StepFrames *>(smem_raw);
  const int hw_lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lane = hw_lane < LANES ? hw_lane : LANES - 1;
  const int tract_raw = blockIdx.x * P.lanes_per_cta + min(lane, P.lanes_per_cta - 1);
  const int tract = tract_raw < P.n_tracts ? tract_raw : P.n_tracts - 1;
  const afs_tube_frame *my_frames = A.tubes + (size_t)tract * P.n_frames;
  const int T = P.total_steps;
  for (int iv = 0; iv + 1 < P.n_frames; iv++) {
    const int i0 = A.bounds[iv], i1 = A.bounds[iv + 1];
    __syncthreads();  // every warp is past its last read of the previous interval's frames
    for (int item = warp; item < 2 * AFS_NUM_PM_SECTIONS + 2; item += SCAN_WARPS) {
      if (item < 2 * AFS_NUM_PM_SECTIONS) {
        const int f = item / AFS_NUM_PM_SECTIONS, j = item % AFS_NUM_PM_SECTIONS;
        const afs_tube_frame *fr = my_frames + iv + f;
        const double a = fr->area_cm2[j];
        S.fA[f][j][lane] = a < kMinArea ? kMinArea : a;  // Tube::setPharynxMouthGeometry clamps on the way in
        S.fLen[f][j][lane] = fr->length_cm[j];
        S.fLat[f][j][lane] = fr->laterality[j];
        if (f == 0) S.art[j][lane] = (unsigned char)fr->articulator[j];
      } else {
        const int f = item - 2 * AFS_NUM_PM_SECTIONS;
        S.inc[f][lane] = my_frames[iv + f].incisor_pos_cm;
      }
    }
    __syncthreads();
    for (int n = i0 + warp; n < i1 && n < T; n += SCAN_WARPS) {
      const double re = __ddiv_rn((double)(n - i0), (double)(i1 - i0)), re1 = __dsub_rn(1.0, re);  // the reference's ratio
      NoiseCons cons[4];
      noise_scan(S, lane, re, re1, cons, nullptr);
      const size_t base = ((size_t)blockIdx.x * (size_t)T + (size_t)n) * 4;
      uint32_t w4[4];
      nzc_pack4(cons, w4);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        A.nzc_i[(base + k) * 32 + hw_lane] = w4[k];
        A.nzc_f[(base + k) * 32 + hw_lane] = cons[k].downF;
      }
    }
  }
}

// ==================================================================================
// K1s `tds_small_kernel`: the same simulation for SMALL batches -- one tract per CTA, one tube section per lane
// ==================================================================================
// tds_batch_kernel maps a tract to a lane: a CTA advances 28 tracts in ~5.5 us per step whether it holds 28 tracts or
// one, so a batch of a few hundred utterances (a shard of BASELINE configs[3] on one of eight GPUs) leaves the SM
// idle.  Here the roles are turned round: a CTA (4 section warps + 1 auxiliary warp) is ONE tract, a lane is one of
// the 93 tube sections (and the current flowing into it), several CTAs share an SM.  The step of a tract is then
// bound by the depth of its dependency chain, not by the serial work of a warp:
//   * update, elements, matrix rows: all sections at once (one lane each);
//   * the linear system (negated: SPD, a tree of chains) in O(log n) depth: the rigid chains (trachea, nose behind
//     the velum port, piriform fossa, sinuses, nostril radiation) have CONSTANT pivots -- factored once in the
//     prologue, only their right-hand sides are eliminated per step, a first-order linear recurrence evaluated as a
//     prefix scan with warp shuffles; the three chains with changing geometry (glottis + pharynx 23..40, mouth
//     41..64, velum port 65..69) are solved by parallel cyclic reduction with warp shuffles, each for two right-hand
//     sides (its own, and a unit load at its junction end), and joined by the 3 x 3 system of the velum clique
//     {40, 41, 65};
//   * the glottis model and the output filters run on the auxiliary warp beside the solve.
// Same arithmetic building blocks as tds_batch_kernel (section_elements, row_dyn, the glottis functions, fast_rcp);
// parity bar 1e-6 like there.  Default-path options only (turbulence, soft walls, skin radiation, fossa, inner length
// corrections, entrance loss constants); noise sources, transvelar coupling, the variable entrance loss, taps and the
// flow source stay with tds_batch_kernel (tds_small_supported).
namespace k1s {

constexpr int NSW = 4;                 // section warps
constexpr int THREADS = 32 * (NSW + 1);
constexpr int ZC = NC;                 // index of an always-zero current in U[]
// chain T = currents 23..40 on warp 0, chain M = 41..64 on warp 1, chain N = 65..69 on warp 2 (lanes 0..4)

struct Smem {
  // section records (index = tube section; NS = all-zero dummy), written in phase 1
  double E[NSX], D[NSX], L[NSX], R0[NSX], R1[NSX], I[NSX], AR[NSX], C[NSX];
  double U[NC + 1];                    // last solution of every current; U[ZC] = 0
  double fA[2][AFS_NUM_PM_SECTIONS], fLen[2][AFS_NUM_PM_SECTIONS], velum[2], gctl[2][AFS_NUM_GLOTTIS_CONTROL];
  double glLen[2][2], glArea[2][2];    // glottis geometry, double buffered by step parity
  double p4[4];                        // pressures of sections 22..25 after the last update
  double sch[8];                       // Schur contributions across warps: [0] rhs of 23 (trachea), [1..5] fossa -> {28, 29}
  double jn[6];                        // y, z of the junction ends of the chains T, M, N
  double urad[2][2];                   // radiated flow of the mouth / the nostrils, double buffered by step parity
  double toneX[4], toneY[4], chebX[8], chebY[8];  // output filters (auxiliary warp, lane 0)
};

// static description of a lane
struct Lane {
  int sec;        // tube section = current flowing into it; -1: idle lane
  int prev;       // section the current comes from (-1: lungs)
  int sib;        // the other current flowing out of `prev` (ZC: none)
  int out1, out2; // currents flowing out of the section (ZC: none)
  int wall;       // 0 standard, 1 trachea, 2 sinus, 3 none (glottis)
  bool dyn, glottal, velum;
};
__device__ __forceinline__ Lane lane_of(int warp, int lane) {
  Lane q;
  int s = -1;
  if (warp == 0) s = lane < 18 ? 23 + lane : -1;
  else if (warp == 1) s = lane < 24 ? 41 + lane : -1;
  else if (warp == 2) s = lane < 19 ? 65 + lane : lane < 23 ? SINUS0 + (lane - 19) : -1;
  else if (warp == 3) s = lane < 23 ? lane : lane < 28 ? FOSSA0 + (lane - 23) : -1;
  q.sec = s;
  q.prev = s <= 0 ? -1 : s == NOSE0 ? PHAR_LAST : s == FOSSA0 ? PHAR0 + FOSSA_COUPLING : s == 89 ? 73 : s == 90 ? 74 : s == 91 ? 76 : s == 92 ? 77 : s - 1;
  q.sib = s == 29 ? 84 : s == 84 ? 29 : s == 41 ? 65 : s == 65 ? 41 : s == 74 ? 89 : s == 89 ? 74 : s == 75 ? 90 : s == 90 ? 75 :
          s == 77 ? 91 : s == 91 ? 77 : s == 78 ? 92 : s == 92 ? 78 : ZC;
  q.out1 = s < 0 ? ZC : s == MOUTH_LAST ? 93 : s == NOSE_LAST ? 95 : (s == FOSSA_LAST || s >= SINUS0) ? ZC : s + 1;
  q.out2 = s == 28 ? 84 : s == 40 ? 65 : s == MOUTH_LAST ? 94 : s == NOSE_LAST ? 96 : s == 73 ? 89 : s == 74 ? 90 : s == 76 ? 91 : s == 77 ? 92 : ZC;
  q.glottal = s == GLOT_LO || s == GLOT_UP;
  q.velum = s >= NOSE0 && s <= NOSE0 + 3;
  q.dyn = s >= GLOT_LO && s <= NOSE0 + 3;
  q.wall = s < 0 ? 3 : q.glottal ? 3 : s <= TRACHEA_LAST ? 1 : s >= SINUS0 ? 2 : 0;
  return q;
}

__device__ __forceinline__ void bar_sec() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NSW) : "memory"); }
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Parallel cyclic reduction of a symmetric tridiagonal system on the lanes 0..n-1 of a warp, two right-hand sides.
// lo = coupling to lane - 1 (0 on lane 0), d = pivot, r1 / r2 = right-hand sides; lanes >= n pass d = 1, lo = r = 0.
// Returns the two solutions.
__device__ __forceinline__ void pcr2(const int lane, const int n, double lo, double d, double r1, double r2, double &y, double &z) {
  double up = shfl(lo, lane + 1);  // coupling to lane + 1 (symmetry)
  if (lane + 1 >= n) up = 0.0;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    if (s >= n) break;  // (warp-uniform)
    const double w = fast_rcp(d);
    const int im = lane - s, ip = lane + s;
    const double wm = shfl(w, im), wp = shfl(w, ip);
    const double lm = shfl(lo, im), um = shfl(up, im), lp = shfl(lo, ip), upp = shfl(up, ip);
    const double r1m = shfl(r1, im), r1p = shfl(r1, ip), r2m = shfl(r2, im), r2p = shfl(r2, ip);
    const double k1 = im >= 0 ? lo * wm : 0.0, k2 = ip < n ? up * wp : 0.0;
    d = fma(-k1, um, fma(-k2, lp, d));
    r1 = fma(-k1, r1m, fma(-k2, r1p, r1));
    r2 = fma(-k1, r2m, fma(-k2, r2p, r2));
    lo = -k1 * lm;
    up = -k2 * upp;
  }
  const double w = fast_rcp(d);
  y = r1 * w; z = r2 * w;
}

// Inclusive scan of the affine recurrence x_k = b_k + a_k x_{k-1} over the lanes lo..hi of a warp (upwards: dir = +1,
// x_{lo-1} = 0) or x_k = b_k + a_k x_{k+1} (downwards: dir = -1, x_{hi+1} = 0).  The a_k are constants of the launch:
// aw[j] = product of the a's of the window of 2^j lanes ending at this lane (precomputed), so a step is one shuffle
// and one FMA.
__device__ __forceinline__ double scan_affine(const int lane, const int lo, const int hi, const int dir, double b, const double (&aw)[5]) {
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const int src = lane - dir * (1 << j);
    const double bp = shfl(b, src);
    if (lane >= lo && lane <= hi && src >= lo && src <= hi) b = fma(aw[j], bp, b);
  }
  return b;
}
// window products for scan_affine (prologue): a = this lane's own factor
__device__ __forceinline__ void scan_windows(const int lane, const int lo, const int hi, const int dir, double a, double (&aw)[5]) {
  if (lane < lo || lane > hi) a = 0.0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    aw[j] = a;
    const int src = lane - dir * (1 << j);
    const double ap = shfl(a, src);
    a = (src >= lo && src <= hi) ? a * ap : 0.0;   // (a window that would leave the chain is never used)
  }
}

struct Args {
  const int32_t *bounds;
  const afs_tube_frame *tubes;
  const double *glottis;
  double *audio, *flow;
  uint32_t *status;
};

__global__ void __launch_bounds__(THREADS, 3)
tds_small_kernel(const TdsParams P, const Args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tract = blockIdx.x;
  const Lane Q = lane_of(warp, lane);
  const int sec = Q.sec;
  const bool aux = warp == NSW;
  const bool on = !aux && sec >= 0;
  const afs_tube_frame *my_frames = A.tubes + (size_t)tract * P.n_frames;
  const double *my_glottis = A.glottis + (size_t)tract * P.n_frames * AFS_NUM_GLOTTIS_CONTROL;
  const int T = P.total_steps, n_out = T / P.sr_factor;

  // ---- prologue: zero state, static records --------------------------------------------
  for (int i = threadIdx.x; i < NSX; i += THREADS) { S.E[i] = S.D[i] = S.L[i] = S.R0[i] = S.R1[i] = S.I[i] = S.AR[i] = S.C[i] = 0.0; }
  for (int i = threadIdx.x; i <= NC; i += THREADS) S.U[i] = 0.0;
  if (threadIdx.x < 8) S.sch[threadIdx.x] = 0.0;
  if (threadIdx.x < 6) S.jn[threadIdx.x] = 0.0;
  if (threadIdx.x < 4) { S.p4[threadIdx.x] = 0.0; (&S.urad[0][0])[threadIdx.x] = 0.0; S.toneX[threadIdx.x] = 0.0; S.toneY[threadIdx.x] = 0.0; }
  if (threadIdx.x < 8) { S.chebX[threadIdx.x] = 0.0; S.chebY[threadIdx.x] = 0.0; }
  // static record of this lane's section (as in tds_batch_kernel's prologue)
  SRec rec{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double kc0 = 0.0, kc1 = 0.0, kc2 = 0.0, ka = 0.0;
  if (on) {
    const int wt = Q.wall;
    if (wt < 3) { ka = P.kalpha[wt]; kc0 = P.kc[wt][0]; kc1 = P.kc[wt][1]; kc2 = P.kc[wt][2]; }
    if (!Q.dyn) {
      const StaticSection q = static_section(sec);
      double Ls, Rs, invr, circ, surface;
      if (sec >= SINUS0) {
        Ls = kRho * (q.length / q.area);
        Rs = (8.0 * kMu * kPi * q.length) / (q.area * q.area);
        invr = 1.0 / sqrt(q.area / kPi);
        surface = P.sinus_surface[sec - SINUS0];
      } else {
        section_elements(q.area, q.length, false, Ls, Rs, invr, circ);
        surface = circ * q.length;
      }
      if (surface < kMinArea) surface = kMinArea;
      const double Cc = q.volume / (kRho * kC * kC);
      const double alpha = P.soft_walls ? surface * P.kalpha[wt] : 0.0;
      rec.L = Ls; rec.R1 = Rs; rec.R0 = Rs; rec.invr = invr; rec.alpha = alpha;
      rec.E = P.dtth / (Cc + alpha);
      if (sec == FOSSA0 && !P.fossa) rec.R0 = 8.0 * kMu * q.length * kPi / (kMinArea * kMinArea);
      S.E[sec] = rec.E; S.L[sec] = rec.L; S.R0[sec] = rec.R0; S.R1[sec] = rec.R1; S.I[sec] = rec.invr; S.AR[sec] = q.area;
    }
  }
  __syncthreads();

  // ---- prologue: constant pivots of the rigid chains --------------------------------------
  // Rows whose two sections are rigid have constant coefficients: currents 0..22 (trachea), 70..83 (nose), 85..88
  // (fossa), the sinus leaves 89..92 and the nostril radiation pair 95, 96.  Elimination order: leaves first, then
  // every chain towards the current that links it to the moving part (23, 69, 84).
  //   sw   = 1 / pivot of this lane's current after the eliminations before it
  //   st   = coupling to the next current in elimination order * sw  (the factor of the right-hand-side recurrence)
  //   sdd  = what the chain takes off the pivot of the linking current (lane of the chain's last eliminated current)
  double sw = 0.0, st = 0.0;
  double fw[5] = {0, 0, 0, 0, 0}, bw[5] = {0, 0, 0, 0, 0};  // window products of the forward / backward scans
  // sinus leaf constants (leaf lanes 19..22 of warp 2) and what they do to their two hosts (host lanes)
  double leaf_w = 0.0, leaf_e = 0.0, leaf_g = 0.0;
  // nostril radiation pair (lane of section 83): factors of radiation_eliminate with constant geometry
  double nr_wL = 0.0, nr_wR = 0.0, nr_o = 0.0, nr_e = 0.0, nr_g = 0.0, nr_LA = 0.0, nr_LAB = 0.0;
  double sdd_link = 0.0;  // (lanes of the linking currents 23 / 69 / 84: constant Schur term of their pivot)
  {
    // plain diagonal of a static row and its coupling to the current before it
    auto sdiag = [&](int s) {
      const int a = s == 0 ? NS : (s == 89 ? 73 : s == 90 ? 74 : s == 91 ? 76 : s == 92 ? 77 : s - 1);
      return S.E[s] + S.E[a] + (S.L[a] + S.L[s]) * P.idt + S.R1[a] + S.R0[s];
    };
    // every lane evaluates the (short) serial recurrences it needs itself: launch-constant, a few dozen flops
    auto nose_pivot = [&](int k) {  // pivot of current k in 70..83 after leaves, radiation and the currents behind it
      double d_next = 0.0, o_next = 0.0;  // of current k+1 (eliminated before k)
      double dk = 0.0;
      for (int c = NOSE_LAST; c >= k; c--) {
        dk = sdiag(c);
        if (c == NOSE_LAST) {  // nostril radiation pair folded into 83 (radiation_eliminate)
          const double E_A = S.E[c], L_A = S.L[c], R_A = S.R1[c], invr = S.I[c], G_A = L_A * P.idt + R_A;
          const double R_rad = (128.0 * kRho * kC / (9.0 * kPi * kPi)) * (invr * invr * (1.0 / kPi));
          const double L_rad = (8.0 * kRho / (3.0 * kPi * kPi)) * invr;
          const double dR = E_A + G_A + R_rad, dL = E_A + (L_A + L_rad) * P.idt + R_A;
          const double e = E_A, g = e + G_A, wL = 1.0 / dL;
          const double dR2 = dR - g * g * wL, o = -e + e * g * wL, wR = 1.0 / dR2;
          dk -= e * e * wL + o * o * wR;
        }
        // sinus leaves: host section h has currents h (in), h+1 (out) and the leaf
        for (int q = 0; q < 4; q++) {
          const int h = q == 0 ? 73 : q == 1 ? 74 : q == 2 ? 76 : 77, l = SINUS0 + q;
          const double e = S.E[h], g = e + S.L[h] * P.idt + S.R1[h], wl = 1.0 / sdiag(l);
          if (c == h) dk -= e * e * wl;
          if (c == h + 1) dk -= g * g * wl;
        }
        if (c < NOSE_LAST) dk -= o_next * o_next / d_next;
        // coupling of current c to current c-1 (through section c-1), with the leaf on section c-1 folded in
        double oc = -S.E[c - 1];
        for (int q = 0; q < 4; q++) {
          const int h = q == 0 ? 73 : q == 1 ? 74 : q == 2 ? 76 : 77, l = SINUS0 + q;
          if (c - 1 == h) { const double e = S.E[h], g = e + S.L[h] * P.idt + S.R1[h]; oc += e * g / sdiag(l); }
        }
        d_next = dk; o_next = oc;
      }
      return dk;
    };
    auto nose_coupling = [&](int c) {  // coupling of current c (70..83) to current c-1
      double oc = -S.E[c - 1];
      for (int q = 0; q < 4; q++) {
        const int h = q == 0 ? 73 : q == 1 ? 74 : q == 2 ? 76 : 77, l = SINUS0 + q;
        if (c - 1 == h) { const double e = S.E[h], g = e + S.L[h] * P.idt + S.R1[h]; oc += e * g / sdiag(l); }
      }
      return oc;
    };
    auto up_pivot = [&](int first, int k) {  // chain eliminated upwards from `first` (trachea): pivot of current k
      double dk = sdiag(first);
      for (int c = first + 1; c <= k; c++) { const double o = -S.E[c - 1]; dk = sdiag(c) - o * o / dk; }
      return dk;
    };
    auto down_pivot = [&](int last, int k) {  // chain eliminated downwards from `last` (fossa): pivot of current k
      double dk = sdiag(last);
      for (int c = last - 1; c >= k; c--) { const double o = -S.E[c]; dk = sdiag(c) - o * o / dk; }
      return dk;
    };
    if (warp == 3 && lane < 23) {                 // trachea: eliminated upwards 0 -> 22, links to 23 through section 22
      sw = 1.0 / up_pivot(0, sec);
      st = -S.E[sec] * sw;                        // coupling to current sec+1 is -E_sec
    } else if (warp == 3 && lane >= 24 && lane < 28) {  // fossa 85..88: eliminated downwards 88 -> 85, links to 84 through section 84
      sw = 1.0 / down_pivot(FOSSA_LAST, sec);
      st = -S.E[sec - 1] * sw;                    // coupling to current sec-1 is -E_{sec-1}
    } else if (warp == 2 && lane >= 5 && lane < 19) {   // nose 70..83: eliminated downwards 83 -> 70, links to 69 through section 69
      sw = 1.0 / nose_pivot(sec);
      st = nose_coupling(sec) * sw;
    }
    // Schur terms on the linking currents
    if (warp == 0 && lane == 0) { const double e = S.E[TRACHEA_LAST]; sdd_link = e * e / up_pivot(0, TRACHEA_LAST); }
    if (warp == 2 && lane == 4) { const double o = nose_coupling(70); sdd_link = o * o / nose_pivot(70); }
    if (warp == 3 && lane == 23) { const double e = S.E[FOSSA0]; sdd_link = e * e / down_pivot(FOSSA_LAST, FOSSA0 + 1); }
    if (warp == 2 && lane >= 19 && lane < 23) {   // sinus leaves
      const int h = Q.prev;
      leaf_e = S.E[h]; leaf_g = leaf_e + S.L[h] * P.idt + S.R1[h]; leaf_w = 1.0 / sdiag(sec);
    }
    if (warp == 2 && lane == 18) {                // nostril radiation pair on section 83
      const double E_A = S.E[sec], L_A = S.L[sec], R_A = S.R1[sec], invr = S.I[sec], G_A = L_A * P.idt + R_A;
      const double R_rad = (128.0 * kRho * kC / (9.0 * kPi * kPi)) * (invr * invr * (1.0 / kPi));
      const double L_rad = (8.0 * kRho / (3.0 * kPi * kPi)) * invr;
      nr_LA = L_A; nr_LAB = L_A + L_rad;
      const double dR = E_A + G_A + R_rad, dL = E_A + nr_LAB * P.idt + R_A;
      nr_e = E_A; nr_g = nr_e + G_A; nr_wL = 1.0 / dL;
      nr_o = -nr_e + nr_e * nr_g * nr_wL;
      nr_wR = 1.0 / (dR - nr_g * nr_g * nr_wL);
    }
  }
  // right-hand-side recurrences: forward (elimination)  r'_k = r_k - st_prev * r'_prev,  backward (solution)
  // u_k = sw_k r'_k - st_k u_link...: both first-order affine recurrences over the lanes of a chain
  {
    // forward: factor applied to the value coming from the previously eliminated current = -st of THAT current
    const int fdir = (warp == 3 && lane < 23) ? +1 : -1;
    const int flo = (warp == 3 && lane < 23) ? 0 : warp == 3 ? 24 : 5, fhi = (warp == 3 && lane < 23) ? 22 : warp == 3 ? 27 : 18;
    const double st_src = shfl(st, lane - fdir);
    scan_windows(lane, flo, fhi, fdir, -st_src, fw);
    // backward: u_k = sw r'_k - st_k u_{toward link}: factor -st of THIS current, direction opposite
    scan_windows(lane, flo, fhi, -fdir, -st, bw);
  }
  const bool chain_tr = warp == 3 && lane < 23, chain_fo = warp == 3 && lane >= 24 && lane < 28, chain_no = warp == 2 && lane >= 5 && lane < 19;
  const int clo = chain_tr ? 0 : chain_fo ? 24 : 5, chi = chain_tr ? 22 : chain_fo ? 27 : 18, cdir = chain_tr ? +1 : -1;

  // ---- per-lane state ----------------------------------------------------------------------
  double p = 0.0, pr = 0.0, wc = 0.0, wcr = 0.0, wcr2 = 0.0;   // section
  double u = 0.0, c = 0.0;                                     // its in-current: last solution, history term
  double E_old = rec.E, al_old = rec.alpha;
  double gbf = 0.0;
  double ru[2] = {0.0, 0.0}, rc[2] = {0.0, 0.0};               // radiation pair of sections 64 / 83 (R, L)
  double rk[7] = {0, 0, 0, 0, 0, 0, 0};                        // (section 64: factors kept for the back substitution)
  // auxiliary warp, lane 0: glottis + output filters
  GlottisState G{0.0, 0.0, 0.0, 0.0};
  GlottisDims GD;
  double prevFlow = 0.0;
  uint32_t bad = 0;

  auto stage = [&](int interval) {  // control frames `interval`, `interval + 1` (all threads)
    for (int i = threadIdx.x; i < 2 * AFS_NUM_PM_SECTIONS; i += THREADS) {
      const int f = i / AFS_NUM_PM_SECTIONS, j = i % AFS_NUM_PM_SECTIONS;
      const afs_tube_frame *fr = my_frames + interval + f;
      const double a = fr->area_cm2[j];
      S.fA[f][j] = a < kMinArea ? kMinArea : a;
      S.fLen[f][j] = fr->length_cm[j];
    }
    if (threadIdx.x < 2) {
      const double v = my_frames[interval + threadIdx.x].velum_area_cm2;
      S.velum[threadIdx.x] = v < kMinArea ? kMinArea : v;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 32 + 2 * AFS_NUM_GLOTTIS_CONTROL) {
      const int i = threadIdx.x - 32, f = i / AFS_NUM_GLOTTIS_CONTROL, cc = i % AFS_NUM_GLOTTIS_CONTROL;
      S.gctl[f][cc] = my_glottis[(size_t)(interval + f) * AFS_NUM_GLOTTIS_CONTROL + cc];
    }
  };

  int interval = 0, i0 = 0, i1 = 0;
  double inv_len = 1.0;
  if (T > 0) {
    i0 = A.bounds[0]; i1 = A.bounds[1];
    inv_len = 1.0 / (double)(i1 - i0);
    stage(0);
    __syncthreads();
    if (aux && lane == 0) {  // glottis geometry of step 0
      double ctl[AFS_NUM_GLOTTIS_CONTROL], len[2], ar[2];
      for (int cc = 0; cc < AFS_NUM_GLOTTIS_CONTROL; cc++) ctl[cc] = S.gctl[0][cc];
      glottis_geometry(P, ctl, G, len, ar, GD);
      S.glLen[0][0] = len[0]; S.glLen[0][1] = len[1]; S.glArea[0][0] = ar[0]; S.glArea[0][1] = ar[1];
    }
    __syncthreads();
  }

#ifdef AFS_TDS_TIMING
  long long kt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, klast = clock64();
#define K_TICK(i) do { const long long now__ = clock64(); kt[i] += now__ - klast; klast = now__; } while (0)
#else
#define K_TICK(i) do { } while (0)
#endif
  for (int n = 0; n <= T && T > 0; n++) {
    const bool upd = n > 0, rows = n < T;
    if (rows && n == i1) {
      interval++;
      i0 = A.bounds[interval]; i1 = A.bounds[interval + 1];
      inv_len = 1.0 / (double)(i1 - i0);
      __syncthreads();
      stage(interval);
      __syncthreads();
    }
    const double ratio = (double)(n - i0) * inv_len, ratio1 = 1.0 - ratio;
    const int par = n & 1;
    double area = 0.0;
    // ============ phase 1: update with the solution of step n-1, elements and D of step n ============
    if (on) {
      double Ls = rec.L, R0s = rec.R0, R1s = rec.R1, invr = rec.invr, Es = rec.E, Dn;
      if (upd) {
        const double out = S.U[Q.out1] + S.U[Q.out2];
        const double beta_old = fma(wc, kc0, fma(wcr, kc1, wcr2 * kc2));
        const double D_old = fma(-E_old, beta_old, fma(P.dtth1, pr, p));
        const double ur = fma(u, P.idt, -c);
        c = fma(ur, P.th, u * P.idt);
        const double p_new = fma(E_old, u - out, D_old);
        const double pr_new = (p_new - p) * P.idt - pr * P.th;
        const double wc_new = fma(pr_new, al_old, beta_old);
        const double wcr_new = (wc_new - wc) * P.idt - wcr * P.th;
        const double wcr2_new = (wcr_new - wcr) * P.idt - wcr2 * P.th;
        p = p_new; pr = pr_new; wc = wc_new; wcr = wcr_new; wcr2 = wcr2_new;
        if (sec >= TRACHEA_LAST && sec <= PHAR0) S.p4[sec - TRACHEA_LAST] = p_new;
        if (sec == MOUTH_LAST || sec == NOSE_LAST) {  // radiation currents (TdsModel.cpp:2063-2072)
#pragma unroll
          for (int j = 0; j < 2; j++) { const double urr = fma(ru[j], P.idt, -rc[j]); rc[j] = fma(urr, P.th, ru[j] * P.idt); }
        }
      }
      if (rows) {
        if (Q.dyn) {
          const int j = min(max(sec - PHAR0, 0), AFS_NUM_PM_SECTIONS - 1);
          area = ratio1 * S.fA[0][j] + ratio * S.fA[1][j];
          double length = ratio1 * S.fLen[0][j] + ratio * S.fLen[1][j];
          if (Q.glottal) { area = S.glArea[par][sec - GLOT_LO]; length = S.glLen[par][sec - GLOT_LO]; }
          if (Q.velum) {
            const double vel = ratio1 * S.velum[0] + ratio * S.velum[1];
            const int k = sec - NOSE0;
            area = vel + ((double)(k * k) * (4.24 - vel)) / 16.0; length = (11.4 / 11.4) * 0.6;
          }
          area = fmax(area, kMinArea);
          double circ;
          section_elements(area, length, Q.glottal, Ls, R0s, invr, circ);
          R1s = R0s;
          const double Cc = (area * length) * (1.0 / (kRho * kC * kC));
          const double alpha = fmax(circ * length, kMinArea) * ka;
          Es = P.dtth * fast_rcp1(Cc + alpha);
          E_old = Es; al_old = alpha;
        } else area = S.AR[sec];
        const double beta = fma(wc, kc0, fma(wcr, kc1, wcr2 * kc2));
        Dn = fma(-Es, beta, fma(P.dtth1, pr, p));
        if (Q.dyn) { S.E[sec] = Es; S.L[sec] = Ls; S.R0[sec] = R0s; S.R1[sec] = R1s; S.I[sec] = invr; S.AR[sec] = area; }
        S.D[sec] = Dn; S.C[sec] = c;
      }
    }
    K_TICK(0);
    bar_all();   // records and p4 visible; the auxiliary warp starts the glottis model
    K_TICK(1);
    if (aux) {
      // ---- auxiliary warp: output of step n-1, glottis model of step n, geometry of step n+1 ----
      if (lane == 0) {
        if (n > 0) {
          double uo = S.urad[par ^ 1][0] + S.urad[par ^ 1][1];  // (step n-1: not touched again before this warp is done)
          if (P.skin) uo += iir_step<4>(c_tone_a, c_tone_b, S.toneX, S.toneY, S.p4[3]);
          if (!isfinite(uo)) bad |= 1u;
          const double outP = (uo - prevFlow) * P.fs;
          prevFlow = uo;
          const double filt = iir_step<8>(P.cheb_a, P.cheb_b, S.chebX, S.chebY, outP);
          if (A.flow) A.flow[(size_t)tract * T + (n - 1)] = uo;
          if ((n % P.sr_factor) == 0) A.audio[(size_t)tract * n_out + (n / P.sr_factor - 1)] = filt * (0.004 / 32767);
        }
        if (rows) {
          double ctl[AFS_NUM_GLOTTIS_CONTROL], len[2], ar[2], pp[4];
          for (int cc = 0; cc < 4; cc++) pp[cc] = S.p4[cc];
          for (int cc = 0; cc < AFS_NUM_GLOTTIS_CONTROL; cc++) ctl[cc] = ratio1 * S.gctl[0][cc] + ratio * S.gctl[1][cc];
          glottis_dims(P, ctl, G, GD);
          glottis_inc_time(P, ctl, G, pp, GD);
          const double rn = (double)(n + 1 - i0) * inv_len, rn1 = 1.0 - rn;
          for (int cc = 0; cc < AFS_NUM_GLOTTIS_CONTROL; cc++) ctl[cc] = rn1 * S.gctl[0][cc] + rn * S.gctl[1][cc];
          glottis_geometry(P, ctl, G, len, ar, GD);
          S.glLen[par ^ 1][0] = len[0]; S.glLen[par ^ 1][1] = len[1]; S.glArea[par ^ 1][0] = ar[0]; S.glArea[par ^ 1][1] = ar[1];
        }
      }
      if (!rows) break;
      K_TICK(2);
      bar_all();  // end of the step
      K_TICK(3);
      continue;
    }
    if (!rows) break;
    // ============ phase 2: matrix rows; right-hand sides of the rigid chains eliminated ============
    double diag = 1.0, rhs = 0.0, lo = 0.0;   // row of this lane's current: pivot, right-hand side, coupling to the current before it
    if (on) {
      const int a = Q.prev < 0 ? NS : Q.prev;
      const double E_A = S.E[a], L_A = S.L[a], R_A = S.R1[a], I_A = S.I[a], D_A = S.D[a];
      const double c_sib = Q.sib == ZC ? 0.0 : S.C[Q.sib];
      const double k_ent = P.k_ent;
      row_dyn(P, row_flags(sec), E_A, L_A, R_A, I_A, D_A, S.E[sec], S.L[sec], S.R0[sec], S.I[sec], S.D[sec], c, c_sib, u,
              sec == GLOT_UP ? S.AR[GLOT_LO] : 0.0, area, k_ent, gbf, diag, rhs);
      if (sec == 0) rhs += ratio1 * S.gctl[0][G_PRESSURE] + ratio * S.gctl[1][G_PRESSURE];  // pressure source at the lungs
      lo = -E_A;
    }
    // sinus leaves into their hosts (warp 2: leaf lanes 19..22, hosts 73, 74, 76, 77 and the current after each)
    if (warp == 2) {
      const double lr = rhs * leaf_w;   // (leaf lanes; 0 elsewhere)
      const double a0 = shfl(leaf_e * lr, 19), b0 = shfl(leaf_g * lr, 19), a1 = shfl(leaf_e * lr, 20), b1 = shfl(leaf_g * lr, 20);
      const double a2 = shfl(leaf_e * lr, 21), b2 = shfl(leaf_g * lr, 21), a3 = shfl(leaf_e * lr, 22), b3 = shfl(leaf_g * lr, 22);
      // leaf l on host section h: r_h += e w r_l, r_{h+1} -= g w r_l
      if (sec == 73) rhs += a0;
      if (sec == 74) rhs += a1 - b0;
      if (sec == 75) rhs -= b1;
      if (sec == 76) rhs += a2;
      if (sec == 77) rhs += a3 - b2;
      if (sec == 78) rhs -= b3;
      if (sec == NOSE_LAST) {  // nostril radiation pair into 83 (constant pivots)
        const double D_A = S.D[sec];
        const double rR = fma(nr_LA, rc[0] + rc[1], D_A), rL = fma(nr_LA, rc[0], nr_LAB * rc[1]) + D_A;
        const double rR2 = rR - nr_g * nr_wL * rL;
        rhs -= -nr_e * nr_wL * rL + nr_o * nr_wR * rR2;
        rk[1] = rL; rk[3] = rR2;
      }
    }
    // rigid chains: r'_k = r_k - st_prev r'_prev (prefix scan), then the Schur contribution to the linking current
    double rel = 0.0;  // eliminated right-hand side of a rigid-chain current
    if (warp >= 2) {
      rel = scan_affine(lane, clo, chi, cdir, (chain_tr || chain_fo || chain_no) ? rhs : 0.0, fw);
      if (warp == 2) {  // nose -> current 69 (lane 4), same warp
        const double t70 = shfl(st * rel, 5);  // coupling(69,70) * r'_70 / pivot_70
        if (lane == 4) { rhs -= t70; diag -= sdd_link; }
      } else {
        if (lane == 22) S.sch[0] = st * rel;           // trachea -> current 23: r_23 -= (-E_22 / pivot_22) r'_22
        const double t85 = shfl(st * rel, 24);          // fossa -> current 84 (lane 23)
        if (lane == 23) {
          rhs -= t85; diag -= sdd_link;
          // current 84 out of the clique {28, 29, 84} (couplings through section 28)
          const double e = S.E[28], g = e + fma(S.L[28], P.idt, S.R1[28]);
          const double wv = fast_rcp(diag);
          S.sch[1] = e * e * wv; S.sch[2] = -e * wv * rhs; S.sch[3] = g * g * wv; S.sch[4] = g * wv * rhs; S.sch[5] = e * g * wv;
          rk[0] = wv; rk[1] = e; rk[2] = g;
        }
      }
    }
    if (warp == 1 && sec == MOUTH_LAST) {  // mouth radiation pair into 64 (radiation_eliminate, geometry of this step)
      const double E_A = S.E[sec], L_A = S.L[sec], R_A = S.R1[sec], invr = S.I[sec], D_A = S.D[sec], G_A = fma(L_A, P.idt, R_A);
      const double R_rad = (128.0 * kRho * kC / (9.0 * kPi * kPi)) * (invr * invr * (1.0 / kPi));
      const double L_rad = (8.0 * kRho / (3.0 * kPi * kPi)) * invr;
      const double L_AB = L_A + L_rad;
      const double dR = E_A + G_A + R_rad, rR = fma(L_A, rc[0] + rc[1], D_A);
      const double dL = E_A + fma(L_AB, P.idt, R_A), rL = fma(L_A, rc[0], L_AB * rc[1]) + D_A;
      const double e = E_A, g = e + G_A;
      const double wL = fast_rcp(dL);
      double dds = e * e * wL, drs = -e * wL * rL;
      const double dR2 = dR - g * g * wL, rR2 = rR - g * wL * rL;
      const double o = -e + e * g * wL;
      const double wR = fast_rcp(dR2);
      dds += o * o * wR; drs += o * wR * rR2;
      diag -= dds; rhs -= drs;
      rk[0] = wL; rk[1] = rL; rk[2] = wR; rk[3] = rR2; rk[4] = o; rk[5] = e; rk[6] = g;
    }
    K_TICK(2);
    bar_sec();
    K_TICK(3);
    // ============ phase 3: the three moving chains by parallel cyclic reduction ============
    double y = 0.0, z = 0.0;
    if (warp <= 2) {
      int nch = warp == 0 ? 18 : warp == 1 ? 24 : 5;
      if (warp == 0) {
        if (lane == 0) { rhs -= S.sch[0]; diag -= sdd_link; lo = 0.0; }          // trachea folded into current 23
        if (lane == 5) { diag -= S.sch[1]; rhs -= S.sch[2]; }                     // fossa folded into 28 ...
        if (lane == 6) { diag -= S.sch[3]; rhs -= S.sch[4]; lo += S.sch[5]; }     // ... and 29 (coupling 28 <-> 29)
      }
      const bool in = lane < nch;
      // second right-hand side: unit load at the junction end (40: last of T; 41, 65: first of M, N)
      const double r2 = (warp == 0 ? lane == nch - 1 : lane == 0) ? 1.0 : 0.0;
      const double lo_c = (in && lane > 0) ? lo : 0.0;   // (the first currents of M and N couple to 40 through the junction)
      pcr2(lane, nch, lo_c, in ? diag : 1.0, in ? rhs : 0.0, in ? r2 : 0.0, y, z);
      if (warp == 0 && lane == nch - 1) { S.jn[0] = y; S.jn[1] = z; }
      if (warp == 1 && lane == 0) { S.jn[2] = y; S.jn[3] = z; }
      if (warp == 2 && lane == 0) { S.jn[4] = y; S.jn[5] = z; }
    }
    K_TICK(4);
    bar_sec();
    // ============ phase 4: the velum clique {40, 41, 65}; solution of the moving currents ============
    if (warp <= 2) {
      // a = u40, b = u41, cc = u65;  e = E_40, g = E_40 + G_40:
      //   a = yT + zT e (b + cc),   b = yM + zM (e a - g cc),   cc = yN + zN (e a - g b)
      const double e = S.E[40], g = e + fma(S.L[40], P.idt, S.R1[40]);
      const double yT = S.jn[0], zT = S.jn[1], yM = S.jn[2], zM = S.jn[3], yN = S.jn[4], zN = S.jn[5];
      // eliminate a:  b (1 - zM e zT e) + cc (zM g - zM e zT e) = yM + zM e yT   (and the same with N)
      const double q = e * zT * e;
      const double m11 = 1.0 - zM * q, m12 = zM * (g - q), m21 = zN * (g - q), m22 = 1.0 - zN * q;
      const double f1 = fma(zM * e, yT, yM), f2 = fma(zN * e, yT, yN);
      const double idet = fast_rcp(m11 * m22 - m12 * m21);
      const double b = (f1 * m22 - m12 * f2) * idet, cc = (m11 * f2 - m21 * f1) * idet;
      const double a = fma(zT * e, b + cc, yT);
      const double wj = warp == 0 ? e * (b + cc) : warp == 1 ? fma(e, a, -g * cc) : fma(e, a, -g * b);
      const int nch = warp == 0 ? 18 : warp == 1 ? 24 : 5;
      if (lane < nch) { u = fma(z, wj, y); S.U[sec] = u; }
    }
    K_TICK(5);
    bar_sec();
    // ============ phase 5: back substitution of the rigid chains, leaves, radiation ============
    if (warp == 1 && sec == MOUTH_LAST) {
      const double us = u;
      const double uR = (rk[3] - rk[4] * us) * rk[2];
      const double uL = (rk[1] + rk[5] * us - rk[6] * uR) * rk[0];
      ru[0] = uR; ru[1] = uL;
      S.U[93] = uR; S.U[94] = uL; S.urad[par][0] = uR + uL;
    }
    if (warp >= 2) {
      // value of the linking current, then u_k = sw r'_k - st u_{towards the link}
      double ulink = 0.0;
      if (warp == 3) {
        if (lane == 23) {  // current 84 from 28 and 29
          u = (rhs + rk[1] * S.U[28] - rk[2] * S.U[29]) * rk[0];
          S.U[sec] = u;
        }
        const double u84 = shfl(u, 23);
        ulink = chain_tr ? S.U[23] : u84;
      } else ulink = shfl(u, 4);   // current 69
      const bool first = chain_tr ? lane == 22 : chain_fo ? lane == 24 : chain_no ? lane == 5 : false;  // next to the link
      double b = sw * rel;
      if (first) b = fma(-st, ulink, b);
      const double uk = scan_affine(lane, clo, chi, -cdir, (chain_tr || chain_fo || chain_no) ? b : 0.0, bw);
      if (chain_tr || chain_fo || chain_no) { u = uk; S.U[sec] = u; }
      if (warp == 2) {
        // sinus leaves: u_l = w_l (r_l + e u_h - g u_{h+1})
        const double u73 = shfl(u, 8), u74 = shfl(u, 9), u75 = shfl(u, 10), u76 = shfl(u, 11), u77 = shfl(u, 12), u78 = shfl(u, 13);
        if (lane >= 19 && lane < 23) {
          const double uh = lane == 19 ? u73 : lane == 20 ? u74 : lane == 21 ? u76 : u77;
          const double un = lane == 19 ? u74 : lane == 20 ? u75 : lane == 21 ? u77 : u78;
          u = (rhs + leaf_e * uh - leaf_g * un) * leaf_w;
          S.U[sec] = u;
        }
        if (sec == NOSE_LAST) {
          const double uR = (rk[3] - nr_o * u) * nr_wR;
          const double uL = (rk[1] + nr_e * u - nr_g * uR) * nr_wL;
          ru[0] = uR; ru[1] = uL;
          S.U[95] = uR; S.U[96] = uL; S.urad[par][1] = uR + uL;
        }
      }
    }
    K_TICK(6);
    bar_all();  // end of the step: solution visible to the update and to the auxiliary warp
    K_TICK(7);
  }
#ifdef AFS_TDS_TIMING
  if (blockIdx.x == 0 && lane == 0)
    printf("k1s warp %d cycles/step: ph1 %lld | wait %lld | ph2 (aux: work) %lld | wait %lld | pcr %lld | junction %lld | back %lld | wait %lld\n", warp,
           kt[0] / T, kt[1] / T, kt[2] / T, kt[3] / T, kt[4] / T, kt[5] / T, kt[6] / T, kt[7] / T);
#endif
  if (aux && lane == 0 && A.status) A.status[tract] = bad;
}

}  // namespace k1s

bool tds_small_supported(const TdsLaunch &L) {
  TdsParams P;
  memcpy(&P, L.params.data(), sizeof P);
  return !P.noise && !P.transvelar && !P.var_loss && P.fsrc_sec < 0 && P.total_steps > 0;
}

cudaError_t launch_tds_small(const TdsLaunch &L, const int32_t *d_bounds, const afs_tube_frame *d_tubes, const double *d_glottis,
                             double *d_audio, double *d_flow, uint32_t *d_status, cudaStream_t stream) {
  TdsParams P;
  memcpy(&P, L.params.data(), sizeof P);
  k1s::Args A{d_bounds, d_tubes, d_glottis, d_audio, d_flow, d_status};
  k1s::tds_small_kernel<<<P.n_tracts, k1s::THREADS, sizeof(k1s::Smem), stream>>>(P, A);
  return cudaGetLastError();
}

// ---- host side ------------------------------------------------------------------
int tds_prepare_launch(const afs_tds_options *o, int n_tracts, int n_frames, const int32_t *frame_len,
                       double fs_tds_hz, int sr_factor, int spread_ctas, TdsLaunch *L) {
  TdsParams P;
  memset(&P, 0, sizeof P);
  long total = 0;
  L->step_bounds.assign((size_t)n_frames, 0);
  for (int k = 0; k + 1 < n_frames; k++) {
    L->step_bounds[k] = (int32_t)total;
    total += (long)frame_len[k] * sr_factor;
    if (total > 0x7fffffffL) return AFS_ERR_BAD_ARGUMENT;
  }
  L->step_bounds[n_frames - 1] = (int32_t)total;
  P.n_tracts = n_tracts; P.n_frames = n_frames; P.total_steps = (int)total; P.sr_factor = sr_factor;
  // tracts per CTA: as many as a CTA holds, i.e. as few CTAs as possible.  A step takes the same time whether a CTA
  // steps one tract or 28, and the SMs of a GPC share their instruction supply: 512 tracts on 19 CTAs are 8 % faster
  // than on 128 (noise sources on).  spread_ctas > 0 (debug key "tds.spread_sms") spreads over that many.
  int lpc = LANES;
  if (spread_ctas > 0) lpc = (n_tracts + spread_ctas - 1) / spread_ctas;
  if (lpc > LANES) lpc = LANES;
  if (lpc < 1) lpc = 1;
  P.lanes_per_cta = lpc;
  P.turbulence = o->turbulence_losses; P.soft_walls = o->soft_walls; P.skin = o->radiation_from_skin;
  P.fossa = o->piriform_fossa; P.ilc = o->inner_length_corrections;
  P.fsar = o->flow_separation_area_ratio;
  P.k_ent = o->glottis_loss_option == AFS_ENTRANCE_LOSS_VAN_DEN_BERG ? 1.375 : 1.0;
  P.transvelar = o->transvelar_coupling ? 1 : 0;
  P.var_loss = o->glottis_loss_option == AFS_VARIABLE_ENTRANCE_LOSS ? 1 : 0;
  if (o->flow_source_section < -1 || o->flow_source_section >= NS || o->reserved != 0) return AFS_ERR_BAD_ARGUMENT;
  P.fsrc_sec = o->flow_source_section;
  P.fsrc_amp = o->flow_source_section >= 0 ? o->flow_source_cm3_s : 0.0;
  P.n_probes = 0;
  chebyshev_lowpass(50.0 / fs_tds_hz, 4, P.tg_a, P.tg_b);
  P.k_sondhi = o->inner_length_corrections ? 8.0 * kRho / (3.0 * kPi * kPi) : 0.0;
  P.k_bern = o->turbulence_losses ? 0.5 * kRho : 0.0;
  const double dt = 1.0 / fs_tds_hz;
  P.dt = dt; P.dtth = dt * kTheta; P.dtth1 = dt * kTheta1; P.idt = 1.0 / (dt * kTheta); P.th = kTheta1 / kTheta;
  P.ith = kTheta / kTheta1; P.fs = fs_tds_hz;
  // wall types: standard (Tube.cpp:8-10), trachea (:103-105), sinus (:208-210)
  const double Mw[3] = {2.1, 0.25, 0.0}, Bw[3] = {800.0, 1000.0, 6500.0}, Kw[3] = {84500.0, 84500.0, 84500.0};
  for (int t = 0; t < 3; t++) {
    const double lq = Mw[t] / (dt * dt * kTheta * kTheta), rq = Bw[t] / (dt * kTheta);
    P.kalpha[t] = o->soft_walls ? 1.0 / (lq + rq + Kw[t]) : 0.0;  // rigid walls: alpha = beta = 0 without branches
    P.c1[t] = lq + rq;
    P.c2[t] = Mw[t] * (kTheta1 / kTheta + 1.0) / (dt * kTheta) + Bw[t] * (kTheta1 / kTheta);
    P.c3[t] = Mw[t] * (kTheta1 / kTheta);
  }
  static const double sinus_vol[4] = {11.3, 6.8, 33.0, 6.2};
  for (int i = 0; i < 4; i++) P.sinus_surface[i] = 4.0 * M_PI * pow((3.0 * sinus_vol[i]) / (4.0 * M_PI), 2.0 / 3.0);
  for (int t = 0; t < 4; t++) {
    P.kc[t][0] = t < 3 ? P.kalpha[t] * P.c1[t] : 0.0;
    P.kc[t][1] = t < 3 ? P.kalpha[t] * P.c2[t] : 0.0;
    P.kc[t][2] = t < 3 ? P.kalpha[t] * P.c3[t] : 0.0;
  }
  chebyshev_lowpass(7000.0 / fs_tds_hz, 8, P.cheb_a, P.cheb_b);  // Synthesizer.cpp:52
  for (int i = 0; i < AFS_NUM_GLOTTIS_STATIC; i++) P.gs[i] = o->glottis_static[i];
  P.g_sqrt_mk[0] = sqrt(P.gs[S_MASS1] * P.gs[S_K1]);
  P.g_sqrt_mk[1] = sqrt(P.gs[S_MASS2] * P.gs[S_K2]);
  P.noise = o->generate_noise_sources ? 1 : 0;
  P.nz_F = 1.0 - exp(-2.0 * M_PI * 40.0 * dt);
  P.nz_x2000 = exp(-2.0 * M_PI * (2000.0 * dt));
  P.nm_a = exp(-2.0 * M_PI * 500.0 * dt);
  P.nm_1a = 1.0 - P.nm_a;
  P.dump_step = 0;
  L->params.resize(sizeof P);
  memcpy(L->params.data(), &P, sizeof P);
  L->n_tracts = n_tracts; L->n_frames = n_frames;
  return AFS_OK;
}

// Noise pre-pass: runs mode (tds_noise_runs_kernel) when the control frames are at least 32 steps apart on average,
// else one record per step (tds_noise_scan_kernel).
bool tds_noise_runs_mode(const TdsLaunch &L) {
  TdsParams P;
  memcpy(&P, L.params.data(), sizeof P);
  return P.noise && P.n_frames >= 2 && (long)P.total_steps >= 32L * (long)(P.n_frames - 1);
}
size_t tds_noise_scan_bytes_per_cta(const TdsLaunch &L) {  // per-step records; 0 without noise sources or in runs mode
  TdsParams P;
  memcpy(&P, L.params.data(), sizeof P);
  return (P.noise && !tds_noise_runs_mode(L)) ? (size_t)P.total_steps * 4 * 32 * (sizeof(uint32_t) + sizeof(double)) : 0;
}
size_t tds_noise_run_bytes() { return sizeof(NzRun); }
int tds_lanes_per_cta(const TdsLaunch &L) {
  TdsParams P;
  memcpy(&P, L.params.data(), sizeof P);
  return P.lanes_per_cta;
}

// Runs mode: the work items of a launch (same for every tract: the control-frame intervals belong to the batch).
// Intervals are cut into pieces of >= 512 steps, at most 16 per interval.
void tds_noise_run_items(const TdsLaunch &L, std::vector<int32_t> *items4, std::vector<int32_t> *iv_first) {
  items4->clear();
  iv_first->clear();
  const int T = L.step_bounds.empty() ? 0 : L.step_bounds.back();
  for (int iv = 0; iv + 1 < L.n_frames; iv++) {
    const int i0 = L.step_bounds[(size_t)iv], i1 = L.step_bounds[(size_t)iv + 1] < T ? L.step_bounds[(size_t)iv + 1] : T;
    const int len = i1 - i0;
    iv_first->push_back((int32_t)(items4->size() / 4));
    if (len <= 0) continue;
    int pieces = len / 512;
    pieces = pieces < 1 ? 1 : pieces > RUN_PIECES ? RUN_PIECES : pieces;
    for (int q = 0; q < pieces; q++) {
      const int lo = i0 + (int)((long)len * q / pieces), hi = i0 + (int)((long)len * (q + 1) / pieces) - 1;
      items4->insert(items4->end(), {iv, lo, hi, 0x7fffffff});
    }
  }
  iv_first->push_back((int32_t)(items4->size() / 4));
  for (size_t k = 0; k + 1 < items4->size() / 4; k++) (*items4)[4 * k + 3] = (*items4)[4 * (k + 1) + 1];
}
// d_runs: pool of `pool_records` records, the first n_tracts * n_items of them the fixed places of the pieces' first
// records; d_ctl[0] (next free record) must hold n_tracts * n_items, d_ctl[1] (overflow flag) 0.
cudaError_t launch_noise_runs(const TdsLaunch &L, int n_items, const void *d_items, const int32_t *d_iv_first,
                              const int32_t *d_bounds, const afs_tube_frame *d_tubes, void *d_runs, uint32_t pool_records,
                              uint32_t *d_ctl, cudaStream_t stream) {
  TdsParams P;
  memcpy(&P, L.params.data(), sizeof P);
  const dim3 grid((unsigned)(P.n_frames - 1), (unsigned)((P.n_tracts + 31) / 32));
  cudaError_t e = cudaFuncSetAttribute(tds_noise_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RunFrames));
  if (e != cudaSuccess) return e;
  tds_noise_runs_kernel<<<grid, 32 * RUN_PIECES, sizeof(RunFrames), stream>>>(
      P.n_tracts, P.n_frames, n_items, d_bounds, d_tubes, (const NzItem *)d_items, d_iv_first, (NzRun *)d_runs, pool_records, d_ctl);
  return cudaGetLastError();
}

// Tracts [tract0, tract0 + n_wave) of the batch (one launch; with per-step noise records the caller splits a batch
// into waves when the records of all its CTAs would not fit `d_nzc`).  Noise sources: either the run records of the
// whole batch (d_runs, d_run_off: indexed by the tract's position in the batch) or the per-step buffer d_nzc.
cudaError_t launch_tds(const TdsLaunch &L, int tract0, int n_wave, const int32_t *d_bounds, const afs_tube_frame *d_tubes,
                       const double *d_glottis, const uint32_t *d_seeds, double *d_audio, double *d_flow,
                       uint32_t *d_status, double *d_dump, int dump_step, void *d_nzc, const void *d_runs,
                       const uint32_t *d_run_ctl, int run_items, const afs_tds_probes *probes, unsigned int *d_progress,
                       int progress_every, const double *d_state_in, double *d_state_out, cudaStream_t stream) {
  TdsParams P;
  memcpy(&P, L.params.data(), sizeof P);
  P.dump_step = dump_step;
  P.n_tracts = n_wave;
  if (probes && probes->n_probes > 0) {
    if (probes->n_probes > AFS_MAX_TDS_PROBES) return cudaErrorInvalidValue;
    P.n_probes = probes->n_probes;
    for (int q = 0; q < P.n_probes; q++) {
      if (probes->section[q] < 0 || probes->section[q] >= NS) return cudaErrorInvalidValue;
      P.probe_pack |= (unsigned long long)probes->section[q] << (8 * q);
    }
  }
  static_assert(sizeof(TdsSmem) <= 227 * 1024, "TdsSmem exceeds the 227 KB a CTA can have on sm_100a");
  // non-default options, a flow source, taps: separate instantiations, like the noise sources
  const bool opt = P.transvelar || P.var_loss || P.fsrc_sec >= 0 || P.n_probes > 0;
  // chunked synthesis (state in / state out): its own instantiations, the one-call kernels stay as they are
  const bool state = d_state_in || d_state_out;
  auto kernel = state ? (P.noise ? (opt ? tds_batch_kernel<true, true, true> : tds_batch_kernel<true, false, true>)
                                 : (opt ? tds_batch_kernel<false, true, true> : tds_batch_kernel<false, false, true>))
                      : (P.noise ? (opt ? tds_batch_kernel<true, true, false> : tds_batch_kernel<true, false, false>)
                                 : (opt ? tds_batch_kernel<false, true, false> : tds_batch_kernel<false, false, false>));
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(TdsSmem));  // per device; cheap
  if (e != cudaSuccess) return e;
  const int ctas = (P.n_tracts + P.lanes_per_cta - 1) / P.lanes_per_cta;
  const size_t n_out = (size_t)(P.total_steps / P.sr_factor);
  KArgs A;
  memset(&A, 0, sizeof A);
  A.bounds = d_bounds;
  A.tubes = d_tubes + (size_t)tract0 * P.n_frames;
  A.glottis = d_glottis + (size_t)tract0 * P.n_frames * AFS_NUM_GLOTTIS_CONTROL;
  A.audio = d_audio + (size_t)tract0 * n_out;
  A.flow = d_flow ? d_flow + (size_t)tract0 * (size_t)P.total_steps : nullptr;
  A.status = d_status ? d_status + tract0 : nullptr;
  A.dump = d_dump;
  A.seeds = d_seeds ? d_seeds + tract0 : nullptr;
  A.progress = progress_every > 0 ? d_progress : nullptr;
  A.progress_every = progress_every;
  A.state_in = d_state_in ? d_state_in + (size_t)tract0 * ST_WORDS : nullptr;
  A.state_out = d_state_out ? d_state_out + (size_t)tract0 * ST_WORDS : nullptr;
  if (P.n_probes > 0) {
    const size_t off = (size_t)tract0 * P.n_probes * (size_t)P.total_steps;
    A.tap_p = probes->pressure_dPa ? probes->pressure_dPa + off : nullptr;
    A.tap_in = probes->inflow_cm3_s ? probes->inflow_cm3_s + off : nullptr;
    A.tap_out = probes->outflow_cm3_s ? probes->outflow_cm3_s + off : nullptr;
  }
  if (P.noise) {
    if (d_runs && d_run_ctl) {
      A.runs = (const NzRun *)d_runs + (size_t)tract0 * (size_t)run_items;  // (runs mode is a single launch: tract0 = 0)
      A.run_items = run_items;
      A.run_ctl = d_run_ctl;
    } else {
      if (!d_nzc) return cudaErrorInvalidValue;
      A.nzc_i = (uint32_t *)d_nzc;
      A.nzc_f = (double *)((unsigned char *)d_nzc + (size_t)ctas * (size_t)P.total_steps * 4 * 32 * sizeof(uint32_t));
      e = cudaFuncSetAttribute(tds_noise_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepFrames));
      if (e != cudaSuccess) return e;
      tds_noise_scan_kernel<<<ctas, SCAN_THREADS, sizeof(StepFrames), stream>>>(P, A);
      e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
  }
  kernel<<<ctas, TDS_THREADS, sizeof(TdsSmem), stream>>>(P, A);
  return cudaGetLastError();
}

size_t tds_smem_bytes() { return sizeof(TdsSmem); }
size_t tds_state_words() { return (size_t)ST_WORDS; }
int tds_total_steps(const TdsLaunch &L) {
  TdsParams P;
  memcpy(&P, L.params.data(), sizeof P);
  return P.total_steps;
}

}  // namespace afs
