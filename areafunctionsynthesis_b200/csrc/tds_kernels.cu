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
// B200 mapping ("tract per lane, sections across warps"):
//   * one CTA = up to 32 independent tracts; LANE = TRACT everywhere, so every
//     warp instruction does useful work for 32 tracts and all branches on the
//     section type are warp-uniform (no divergence);
//   * 16 warps each own 3 "dynamic" slots (sections 23..68 whose geometry
//     changes every sample, blocked 3 per warp, + sections 69 and 84 whose
//     upstream neighbour is dynamic) and 3 static slots, together with the
//     branch current flowing INTO each owned section; the recurrent state (p,
//     p', three wall-current terms, the current's history term c) never leaves
//     the owner's REGISTERS for the whole utterance;
//   * everything that crosses sections (E, D, L, R, 1/r, alpha, matrix rows,
//     the solution) is exchanged through SHARED MEMORY in [item][lane] layout
//     (256 B per item: conflict-free); all per-slot addressing comes from a
//     small uint16 offset table built once, so the phases are straight-line
//     code with uniform-offset LDS/STS and no per-access index arithmetic;
//   * per step: phase 2 rows -> phase 3 solve (parallel partition over 12
//     warps + 3-warp reduced system) || glottis || output filters -> fused
//     phase 4 (state update) + phase 1 of the next step (elements, E, D);
//   * the two control frames of the current interval are staged in shared
//     memory; audio and flow are the only per-step HBM traffic.
// FP64 throughout.  Algebraically simplified w.r.t. the reference (wall
// admittance alpha = surface * const, one rcp + one sqrt + one rcp per dynamic
// section instead of ~15 div/sqrt, branch-free MUFU+Newton rcp/sqrt), FMA
// contraction on; parity gate 1e-6 relative max-abs on audio.
#include <cmath>
#include <cstddef>
#include <cstring>

#include "afs_internal.h"

namespace afs {

constexpr int LANES = 32;
constexpr int NWS = 16;                       // warps per CTA
constexpr int TDS_THREADS = NWS * 32;
constexpr int W_GLOTTIS = 3, W_OUTPUT = 4;    // warps that run the glottis / output filters during phase 3
constexpr int ND = 46;                        // truly dynamic sections 23..68
constexpr int DYN0 = 23, DYN_LAST = 68;
constexpr int NDX = 49;                       // per-lane exchange records: 0..45 sections 23..68, 46 = 69, 47 = 84, 48 = copy of 22
constexpr int DX_69 = 46, DX_84 = 47, DX_22 = 48;
constexpr int KD = 3, KS = 3;                 // dynamic / static slots per warp
constexpr int NSTAT = 45;                     // sections owned by static slots (all but 23..69 and 84)
constexpr int NCX = NC + 1;                   // + one always-zero dummy current (index 97)
constexpr int NSX = NS + 1;                   // + one all-zero dummy section (index 93): "no source section"
constexpr double kTheta = 0.515;              // TdsModel.cpp:15
constexpr double kTheta1 = 1.0 - kTheta;

struct TdsParams {
  int n_tracts, n_frames, lanes_per_cta, total_steps, sr_factor;
  int turbulence, soft_walls, skin, fossa, ilc;
  int dump_step;
  double dt, idt, th, dtth, dtth1;  // 1/fs, 1/(dt*theta), theta1/theta, dt*theta, dt*theta1
  double fsar, k_ent;
  double kalpha[3], c1[3], c2[3], c3[3];  // wall constants: 0 standard, 1 trachea, 2 sinus
  double kc[4][3];                        // kalpha*c1, kalpha*c2, kalpha*c3 per wall type; 3 = no walls (zeros)
  double sinus_surface[4];
  double cheb_a[9], cheb_b[9];
  double gs[AFS_NUM_GLOTTIS_STATIC];
  double g_sqrt_mk[2];  // sqrt(mass_i * spring_i): the tension Q cancels (TriangularGlottis.cpp:247-249)
};

// Static section record (same for all tracts).  R1 = resistance seen by the out-flowing current, R0 by the
// in-flowing one (they differ only for the decoupled piriform fossa entrance, TdsModel.cpp:883-889).
struct SRec { double E, L, R1, invr, alpha, R0; };

// Per (warp, slot) byte offsets, built once in the prologue.  All offsets are relative to array bases and fit 16 bits.
struct SlotTab {
  uint16_t flags;
  uint8_t sec;       // section (= in-flowing current) index: D[], diag[], sol[]
  uint8_t own;       // dynamic slots: record index into E/L/R/invr/alpha; static slots: section (srec[])
  uint8_t aRec;      // source section A: dynamic slots -> record index; static slots -> section (srec[])
  uint8_t aSec;      // A as an index into D[] (dummy section 93 = zero when there is no source section)
  uint8_t out0, out1;      // out-flowing currents (dummy current 97 = zero)
  uint8_t urbOwn, urbSib;  // slots in urb[]: where to publish the own history term / read the sibling's
  uint8_t p4;        // slot in p4[]
  uint8_t wall;      // wall-constant set in wallc[] (3 = no walls: glottis)
};
enum { FL_JUNCTION = 1, FL_GLO = 2, FL_GUP = 4, FL_ROW0 = 8, FL_VALID = 16, FL_G3 = 32, FL_OUT1 = 64, FL_SIB = 128,
       FL_URB = 256, FL_P4 = 512, FL_WALL = 1024 };

// ---- shared memory layout ---------------------------------------------------
struct TdsSmem {
  // per-tract exchange, [item][lane]
  double D[NSX][LANES];
  double E[NDX][LANES];
  double L[NDX][LANES], R[NDX][LANES], invr[NDX][LANES];  // contiguous: reused as solver scratch in phase 3
  double alpha[NDX][LANES];
  double diag[NC][LANES];
  double sol[2][NCX][LANES];
  double urb[18][LANES];          // history terms c of the 16 sibling (branch / radiation) currents, 16 = sink, 17 = zero
  double fA[2][AFS_NUM_PM_SECTIONS][LANES], fLen[2][AFS_NUM_PM_SECTIONS][LANES];  // control frames of the interval
  double velum[2][LANES];
  double gctl[2][AFS_NUM_GLOTTIS_CONTROL][LANES];
  double glLen[2][2][LANES], glArea[2][2][LANES];  // glottis geometry, double buffered by step parity
  double p4[5][LANES];            // pressures of sections 22..25 after the last update, 4 = sink
  double g3[3][LANES];            // G = L/(dt theta)+R of the dynamic branching sections 28, 40, 64
  double gx[4][LANES];            // glottis: x0, x1 (current), xp0, xp1 (previous) relative displacements
  double gdim[10][LANES];         // glottis: cached dimensions for the current displacement
  double toneX[4][LANES], toneY[4][LANES];   // skin-radiation filter history
  double chebX[8][LANES], chebY[8][LANES];   // output low-pass history
  double prevFlow[LANES];
  // same for all tracts
  SRec srec[NSX];
  double wallc[4][3];             // kalpha*c1, kalpha*c2, kalpha*c3 per wall type; 3 = zeros
  SlotTab slot[NWS][KD + KS];
};

__device__ __forceinline__ int wall_type(int s) { return s <= TRACHEA_LAST ? 1 : s >= SINUS0 ? 2 : 0; }
__device__ __forceinline__ bool is_dyn(int s) { return s >= DYN0 && s <= DYN_LAST; }
// section owned by dynamic slot index dx = 3*warp + k / static slot index t = 3*warp + k (t < 45)
__device__ __forceinline__ int dyn_slot_section(int dx) { return dx < ND ? DYN0 + dx : dx == DX_69 ? 69 : FOSSA0; }
__device__ __forceinline__ int stat_slot_section(int t) {
  return t < 23 ? t : t < 37 ? 70 + (t - 23) : t < 41 ? 85 + (t - 37) : SINUS0 + (t - 41);
}
// record index of a section in the per-lane exchange arrays (-1: none)
__device__ __forceinline__ int dyn_record_of(int s) {
  return is_dyn(s) ? s - DYN0 : s == 69 ? DX_69 : s == FOSSA0 ? DX_84 : s == TRACHEA_LAST ? DX_22 : -1;
}
// source section of current i (TdsModel.cpp:93-131); 93 (dummy) for the current out of the lungs
__device__ __forceinline__ int src_of(int i) {
  if (i == 0) return NS;
  if (i == NOSE0) return PHAR_LAST;
  if (i == FOSSA0) return PHAR0 + FOSSA_COUPLING;
  if (i >= SINUS0 && i <= SINUS_LAST) return i == 89 ? 73 : i == 90 ? 74 : i == 91 ? 76 : 77;
  return i - 1;
}
// sibling current at a branching source section, -1 if none
__device__ __forceinline__ int sibling_of(int i) {
  switch (i) {
    case 29: return 84; case 84: return 29;
    case 41: return 65; case 65: return 41;
    case 74: return 89; case 89: return 74;
    case 75: return 90; case 90: return 75;
    case 77: return 91; case 91: return 77;
    case 78: return 92; case 92: return 78;
    default: return -1;
  }
}
// slot in the small sibling table urb[]
__device__ __forceinline__ int urb_slot(int i) {
  switch (i) {
    case 29: return 0; case 84: return 1; case 41: return 2; case 65: return 3;
    case 74: return 4; case 89: return 5; case 75: return 6; case 90: return 7;
    case 77: return 8; case 91: return 9; case 78: return 10; case 92: return 11;
    case 93: return 12; case 94: return 13; case 95: return 14; case 96: return 15;
    default: return -1;
  }
}
// out-flowing currents of a section (dummy 97 where there is none)
__device__ __forceinline__ void outs_of(int s, int &o0, int &o1) {
  o0 = s + 1; o1 = NC;
  if (s == 28) o1 = 84;
  else if (s == 40) o1 = 65;
  else if (s == 64) { o0 = 93; o1 = 94; }
  else if (s == 83) { o0 = 95; o1 = 96; }
  else if (s == 73) o1 = 89;
  else if (s == 74) o1 = 90;
  else if (s == 76) o1 = 91;
  else if (s == 77) o1 = 92;
  else if (s == FOSSA_LAST || s >= SINUS0) o0 = NC;
}

// Recurrent state of one owned section + its in-flowing current (registers).
// c = u/(dt theta) + (theta1/theta) u' is all that the next step needs of the current's history:
// H = -L c (TdsModel.cpp:2020) and u'_new = u_new/(dt theta) - c (TdsModel.cpp:2070).
struct SecState {
  double p, pr, wc, wcr, wcr2, c;
};

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
  const double inva = fast_rcp(area);
  const double r = fast_sqrt(area * (1.0 / kPi));
  invr = r * kPi * inva;
  circ = 2.0 * kPi * r;  // 2*sqrt(area*pi)
  L = (kRho * 0.5 * length) * inva;
  const double rmin = glottal ? 0.8 : 1.6;
  if (r < rmin) {  // elliptical duct, major half-axis rmin (TdsModel.cpp:772-788)
    const double a = rmin, b = area * (1.0 / (kPi * rmin));
    const double t = kPi * inva;
    R = (2.0 * kMu * length) * (a * a + b * b) * (t * t * inva);  // 1/(pi a^3 b^3) = pi^2 / area^3
  } else {
    R = (4.0 * kMu * kPi * length) * (inva * inva);
  }
}


constexpr int ITEM = LANES * 8;  // bytes per [item] row

// beta of TdsModel.cpp:826-831 with Lw = Mw/surface, Rw = Bw/surface folded in:
// beta = kalpha*(c1 wc + c2 wc' + c3 wc''), independent of the geometry.
__device__ __forceinline__ double wall_beta3(const double *w, const SecState &st) {
  return fma(st.wc, w[0], fma(st.wcr, w[1], st.wcr2 * w[2]));
}

// ---- phase 1: per-section elements, E and D ----------------------------------
// Dynamic slot dx (uniform).
__device__ __forceinline__ void phase1_dyn(const TdsParams &P, TdsSmem &S, int lane, int dx, const SlotTab &T,
                                           const SecState &st, double ratio, double ratio1, int parity) {
  double E;
  const int own = T.own;
  const double beta = (T.flags & FL_WALL) ? wall_beta3(P.kc[0], st) : 0.0;  // dynamic slots: standard walls or none
  if (dx < ND) {
    double area, length;
    if (dx >= 2 && dx < 42) {  // pharynx / mouth: Tube::interpolate + setPharynxMouthGeometry (Tube.cpp:449-466, 323-350)
      const int j = dx - 2;
      area = ratio1 * S.fA[0][j][lane] + ratio * S.fA[1][j][lane];
      length = ratio1 * S.fLen[0][j][lane] + ratio * S.fLen[1][j][lane];
    } else if (dx < 2) {  // glottis sections from the glottis model (Synthesizer.cpp:576-579)
      area = S.glArea[parity][dx][lane];
      length = S.glLen[parity][dx][lane];
    } else {  // velum port: Tube::setVelumOpening (Tube.cpp:402-416)
      const int j = dx - 42;
      const double v = ratio1 * S.velum[0][lane] + ratio * S.velum[1][lane];
      area = v + ((double)(j * j) * (4.24 - v)) / 16.0;
      length = (11.4 / 11.4) * 0.6;
    }
    if (area < kMinArea) area = kMinArea;
    const double volume = area * length;
    double L, R, invr, circ;
    const bool glottal = dx < 2;
    section_elements(area, length, glottal, L, R, invr, circ);
    const double C = volume * (1.0 / (kRho * kC * kC));
    double alpha = 0.0;
    if (P.soft_walls && !glottal) {  // TdsModel.cpp:805-832
      double surface = circ * length;
      if (surface < kMinArea) surface = kMinArea;
      alpha = surface * P.kalpha[0];
    }
    E = P.dtth * fast_rcp(C + alpha);
    S.E[own][lane] = E; S.L[own][lane] = L; S.R[own][lane] = R; S.invr[own][lane] = invr; S.alpha[own][lane] = alpha;
    if (T.flags & FL_G3) S.g3[dx == 5 ? 0 : dx == 17 ? 1 : 2][lane] = L * P.idt + R;
  } else {  // sections 69 and 84: static geometry; E/alpha copies are permanent, but L, R, 1/r live in the arrays the
            // solver reuses as scratch, so they are refreshed every step
    E = S.E[own][lane];
    const SRec &r = S.srec[dx == DX_69 ? 69 : FOSSA0];
    S.L[own][lane] = r.L; S.R[own][lane] = dx == DX_84 ? r.R0 : r.R1; S.invr[own][lane] = r.invr;
  }
  if (dx == 0) {  // per-lane copy of static section 22, the source section of row 23 (same reason)
    const SRec &r = S.srec[TRACHEA_LAST];
    S.L[DX_22][lane] = r.L; S.R[DX_22][lane] = r.R1; S.invr[DX_22][lane] = r.invr;
  }
  S.D[T.sec][lane] = st.p + P.dtth1 * st.pr - E * beta;  // TdsModel.cpp:988-1008, no flow sources
}

__device__ __forceinline__ void phase1_stat(const TdsParams &P, TdsSmem &S, int lane, const SlotTab &T, const SecState &st) {
  const double beta = wall_beta3(S.wallc[T.wall], st);
  S.D[T.sec][lane] = st.p + P.dtth1 * st.pr - S.srec[T.own].E * beta;
}

// ---- phase 2: one matrix row (negated: SPD) per owned current -----------------
// Row of the current flowing from section A into the owned section B (TdsModel.cpp:1916-2034):
//   diag = E_B + E_A + L_AB/(dt theta) + R_AB,   rhs = L_AB c + L_A c_sibling + [p_lung] - D_B + D_A
// with c the history term of SecState.  The same formula covers plain, bifurcation and "no source section" rows
// (dummy zero records); junction inductance / Bernoulli and the glottal losses are added under uniform flags.
template <bool DYN>
__device__ __forceinline__ void phase2_row(const TdsParams &P, TdsSmem &S, int lane, const SlotTab &T, const SecState &st,
                                           int prev, int cur, double psrc, double &gbf) {
  double E_B, L_B, R_B, E_A, L_A, R_A, ia = 0.0, ib = 0.0;
  const int fl = T.flags;
  if (DYN) {
    const int own = T.own, a = T.aRec;
    E_B = S.E[own][lane]; L_B = S.L[own][lane]; R_B = S.R[own][lane]; ib = S.invr[own][lane];
    E_A = S.E[a][lane]; L_A = S.L[a][lane]; R_A = S.R[a][lane]; ia = S.invr[a][lane];
  } else {
    const SRec &b = S.srec[T.own], &a = S.srec[T.aRec];
    E_B = b.E; L_B = b.L; R_B = b.R0;
    E_A = a.E; L_A = a.L; R_A = a.R1;
  }
  const double D_B = S.D[T.sec][lane], D_A = S.D[T.aSec][lane];
  const double c_sib = (fl & FL_SIB) ? S.urb[T.urbSib][lane] : 0.0;
  double L_AB = L_A + L_B;
  if (DYN) {
    if (fl & (FL_JUNCTION | FL_GLO | FL_GUP)) {
      const double uB = S.sol[prev][T.sec][lane];
      if (fl & FL_JUNCTION) {
        if (P.ilc) L_AB += (8.0 * kRho / (3.0 * kPi * kPi)) * fabs(ia - ib);  // Sondhi (TdsModel.cpp:1745-1778)
        if (P.turbulence) {  // Bernoulli loss wide -> narrow (TdsModel.cpp:850-877); area_B < area_A <=> ib > ia
          if (((ib > ia) && (uB > 0)) || ((ib < ia) && (uB < 0))) {
            const double qa = ia * ia * (1.0 / kPi), qb = ib * ib * (1.0 / kPi);  // 1/area
            R_A -= uB * (0.5 * kRho) * (qa * qa);
            R_B += uB * (0.5 * kRho) * (qb * qb);
          }
        }
      } else if (fl & FL_GLO) {  // glottal entrance loss (TdsModel.cpp:900-920)
        if (uB > 0) {
          const double qT = ib * ib * (1.0 / kPi), qS = ia * ia * (1.0 / kPi);
          R_B += P.k_ent * 0.5 * kRho * fabs(uB) * (qT * qT - qS * qS);
        }
      } else {  // lower -> upper glottis section (TdsModel.cpp:930-950)
        const double qS = ia * ia * (1.0 / kPi), qT = ib * ib * (1.0 / kPi);
        const double target = (qT * P.fsar > qS) ? 1.0 : 0.0;  // area_T < fsar * area_S
        gbf = 0.8 * gbf + (1.0 - 0.8) * target;
        if (uB > 0) R_A += gbf * fabs(uB) * 0.5 * kRho * (qT * qT - qS * qS);
      }
    }
  }
  const double diag = E_B + E_A + fma(L_AB, P.idt, R_A + R_B);
  double rhs = fma(L_AB, st.c, fma(L_A, c_sib, D_A - D_B));
  if (!DYN) { if (fl & FL_ROW0) rhs += psrc; }
  S.diag[T.sec][lane] = diag;
  S.sol[cur][T.sec][lane] = rhs;
}

// Rows of the two parallel radiation currents out of section A (TdsModel.cpp:1841-1911):
// iR through the radiation resistance, iL = iR + 1 through the radiation inductance.
__device__ __forceinline__ void phase2_radiation(const TdsParams &P, TdsSmem &S, int lane, int A, int iR, double cR,
                                                 double cL, int cur) {
  const int iL = iR + 1;
  double E_A, L_A, R_A, invr;
  if (A == MOUTH_LAST) { const int d = MOUTH_LAST - DYN0; E_A = S.E[d][lane]; L_A = S.L[d][lane]; R_A = S.R[d][lane]; invr = S.invr[d][lane]; }
  else { E_A = S.srec[A].E; L_A = S.srec[A].L; R_A = S.srec[A].R1; invr = S.srec[A].invr; }
  const double D_A = S.D[A][lane];
  const double R_rad = (128.0 * kRho * kC / (9.0 * kPi * kPi)) * (invr * invr * (1.0 / kPi));
  const double L_rad = (8.0 * kRho / (3.0 * kPi * kPi)) * invr;
  const double L_AB = L_A + L_rad;
  S.diag[iR][lane] = E_A + L_A * P.idt + R_A + R_rad;
  S.sol[cur][iR][lane] = L_A * (cR + cL) + D_A;
  S.diag[iL][lane] = E_A + L_AB * P.idt + R_A;
  S.sol[cur][iL][lane] = fma(L_A, cR, L_AB * cL) + D_A;
}

// ---- phase 4: update of the recurrent state (TdsModel.cpp:2046-2098) -------------
template <bool DYN>
__device__ __forceinline__ void phase4_update(const TdsParams &P, TdsSmem &S, int lane, const SlotTab &T, SecState &st,
                                              int cur) {
  const int fl = T.flags;
  const double u_new = S.sol[cur][T.sec][lane];
  double out = S.sol[cur][T.out0][lane];  // (dummy zero current for dead ends)
  if (fl & FL_OUT1) out += S.sol[cur][T.out1][lane];
  double E, alpha, beta;
  if (DYN) {
    E = S.E[T.own][lane]; alpha = S.alpha[T.own][lane];
    beta = (fl & FL_WALL) ? wall_beta3(P.kc[0], st) : 0.0;
  } else {
    const SRec &b = S.srec[T.own];
    E = b.E; alpha = b.alpha;
    beta = wall_beta3(S.wallc[T.wall], st);
  }
  const double D = S.D[T.sec][lane];
  // current: u' = (u_new - u_old)/(dt theta) - (theta1/theta) u'_old = u_new/(dt theta) - c
  const double ur_new = fma(u_new, P.idt, -st.c);
  st.c = fma(ur_new, P.th, u_new * P.idt);
  if (fl & FL_URB) S.urb[T.urbOwn][lane] = st.c;
  // section
  const double p_new = fma(E, u_new - out, D);
  const double pr_new = (p_new - st.p) * P.idt - st.pr * P.th;
  const double wc_new = fma(pr_new, alpha, beta);
  const double wcr_new = (wc_new - st.wc) * P.idt - st.wcr * P.th;
  const double wcr2_new = (wcr_new - st.wcr) * P.idt - st.wcr2 * P.th;
  st.p = p_new; st.pr = pr_new; st.wc = wc_new; st.wcr = wcr_new; st.wcr2 = wcr2_new;
  if (fl & FL_P4) S.p4[T.p4][lane] = p_new;
}

// ---- phase 3: the linear solve (parallel partition) ------------------------------
// Negated system (SPD).  Couplings: two currents meeting in section s are coupled by -E_s (in/out) or
// by +E_s + G_s (two outs of a branching section), G_s = L_s/(dt theta) + R_s.  The graph is a tree of
// 3-cliques, so elimination from the leaves has no fill.  A single warp can only issue one FP64
// instruction every ~3 cycles (measured), so the 97 unknowns are cut into SEGMENTS that different
// warps eliminate concurrently (one tract per lane in every warp):
//   3a  each segment warp eliminates its interior nodes towards its two separator nodes (the coupling
//       to the left separator travels along as a "spike" f); leaf cliques (sinuses, radiation pairs,
//       piriform fossa) are folded in by the warp that owns their neighbours or by a leaf warp;
//   3b  the reduced system over the 11 separators 9,19,28,29,40 | 41,53,64 | 65,75,83 (three short
//       chains meeting in the velum clique {40,41,65}) is solved by three warps + one junction step;
//   3c  every segment warp back-substitutes its interior.
// Per interior node i the forward pass stores T = o_i/d'_i (in diag[]), G = r'_i/d'_i (in sol[]) and
// H = f_i/d'_i (scratch aliased on the L/R/invr exchange arrays, dead after phase 2).

struct Solver {
  TdsSmem &S;
  int lane, cur;
  double idt;
  __device__ __forceinline__ double &d(int i) { return S.diag[i][lane]; }
  __device__ __forceinline__ double &r(int i) { return S.sol[cur][i][lane]; }
  // scratch aliased on the L/R/invr exchange arrays (3*NDX = 147 contiguous items, dead between phase 2 and
  // the next phase 1): items 0..96 = H per node, 100..120 = side slots, 124..132 = leaf slots
  __device__ __forceinline__ double &H(int i) { return (&S.L[0])[i][lane]; }
  // side slot of a two-sided segment warp: warps 1,2 -> 0,1; 5..9 -> 2..6
  __device__ __forceinline__ double &side(int w, int k) { return (&S.L[0])[100 + 3 * (w < 3 ? w - 1 : w - 3) + k][lane]; }
  __device__ __forceinline__ double &leaf(int k) { return (&S.L[0])[124 + k][lane]; }
  __device__ __forceinline__ double E(int s) const { return is_dyn(s) ? S.E[s - DYN0][lane] : S.srec[s].E; }
  // G_s of the branching sections: static ones from the table, dynamic ones (28, 40, 64) published in phase 1
  __device__ __forceinline__ double G(int s) const {
    return s == 28 ? S.g3[0][lane] : s == 40 ? S.g3[1][lane] : s == 64 ? S.g3[2][lane] : S.srec[s].L * idt + S.srec[s].R1;
  }

  // Eliminate leaf current c (second out of branching section s) into its clique partners a (in-current of s)
  // and b (other out-current), both updated IN PLACE.  Returns the modified coupling o(a,b).
  __device__ __forceinline__ double leaf_in_place(int c, int s, int a, int b) {
    const double e = E(s), g = e + G(s);
    const double w = fast_rcp(d(c));
    d(c) = w;
    const double rc = r(c);
    d(a) -= e * e * w;
    d(b) -= g * g * w;
    r(a) += e * w * rc;
    r(b) -= g * w * rc;
    return -e + e * g * w;
  }
  __device__ __forceinline__ void leaf_back(int c, int s, int a, int b) {
    const double e = E(s), g = e + G(s);
    r(c) = (r(c) + e * r(a) - g * r(b)) * d(c);
  }

  // 3a: eliminate `count` interior nodes first, first+dir, ... towards separator b = first + dir*count; a (or -1)
  // is the separator on the other side.  Link sections ov0s / ov1s carry modified couplings ov0 / ov1.
  __device__ __forceinline__ void seg_forward(int w_id, int a, int first, int count, int dir, int ov0s, double ov0,
                                              int ov1s, double ov1) {
    double f = 0.0;
    if (a >= 0) { const int ls = dir > 0 ? a : a - 1; f = ls == ov0s ? ov0 : ls == ov1s ? ov1 : -E(ls); }
    double dc = d(first), rc = r(first), dda = 0.0, dra = 0.0;
    int i = first;
#pragma unroll 2
    for (int j = 0; j < count; j++, i += dir) {
      const int ls = dir > 0 ? i : i - 1;
      const double o = ls == ov0s ? ov0 : ls == ov1s ? ov1 : -E(ls);
      const double dn = d(i + dir), rn = r(i + dir);
      const double w = fast_rcp(dc);
      const double t = o * w, h = f * w;
      d(i) = t; r(i) = rc * w; H(i) = h;
      dda = fma(h, f, dda); dra = fma(h, rc, dra);
      dc = fma(-t, o, dn);
      rc = fma(-t, rc, rn);
      f = -t * f;
    }
    d(i) = dc; r(i) = rc;  // separator b, in place (only this warp touches it in 3a)
    if (a >= 0) { side(w_id, 0) = dda; side(w_id, 1) = dra; side(w_id, 2) = f; }
  }
  // 3c: u_i = G_i - H_i u_a - T_i u_next, walking back from b
  __device__ __forceinline__ void seg_back(int a, int first, int count, int dir) {
    const double ua = a >= 0 ? r(a) : 0.0;
    int i = first + dir * (count - 1);
    double un = r(i + dir);
#pragma unroll 2
    for (int j = 0; j < count; j++, i -= dir) {
      un = fma(-d(i), un, fma(-H(i), ua, r(i)));
      r(i) = un;
    }
  }
  // 3b helpers on separators (in place): eliminate x into y over coupling c / recover x from y
  __device__ __forceinline__ void elim_into(int x, int y, double c) {
    const double w = fast_rcp(d(x));
    d(x) = w;
    const double t = c * w;
    d(y) -= t * c;
    r(y) -= t * r(x);
  }
  __device__ __forceinline__ void back_from(int x, int y, double c) { r(x) = (r(x) - c * r(y)) * d(x); }
};

// warp roles in phase 3 (warps W_GLOTTIS = 3 and W_OUTPUT = 4 run the glottis and the output filters)
enum { R_T0 = 0, R_T1 = 1, R_T2 = 2, R_T4 = 5, R_M1 = 6, R_M2 = 7, R_N1 = 8, R_N2 = 9, R_FOSSA = 10, R_RADM = 11,
       R_RADN = 12, R_RED_T = 13, R_RED_M = 14, R_RED_N = 15 };

__device__ __forceinline__ void solve_3a(Solver &Z, int warp, double &k0, double &k1) {
  switch (warp) {
    case R_T0: Z.seg_forward(warp, -1, 0, 9, +1, -1, 0.0, -1, 0.0); break;    // 0..8   -> 9
    case R_T1: Z.seg_forward(warp, 9, 10, 9, +1, -1, 0.0, -1, 0.0); break;    // 10..18 -> 19
    case R_T2: Z.seg_forward(warp, 19, 20, 8, +1, -1, 0.0, -1, 0.0); break;   // 20..27 -> 28
    case R_T4: Z.seg_forward(warp, 29, 30, 10, +1, -1, 0.0, -1, 0.0); break;  // 30..39 -> 40
    case R_M1: Z.seg_forward(warp, 41, 42, 11, +1, -1, 0.0, -1, 0.0); break;  // 42..52 -> 53
    case R_M2: Z.seg_forward(warp, 53, 54, 10, +1, -1, 0.0, -1, 0.0); break;  // 54..63 -> 64
    case R_N1:  // sinuses 89 (on 73: {73,74}) and 90 (on 74: {74,75}), then 66..74 -> 75
      k0 = Z.leaf_in_place(89, 73, 73, 74);
      k1 = Z.leaf_in_place(90, 74, 74, 75);
      Z.seg_forward(warp, 65, 66, 9, +1, 73, k0, 74, k1);
      break;
    case R_N2:  // sinuses 91 (on 76: {76,77}) and 92 (on 77: {77,78}), then 76..82 -> 83
      k0 = Z.leaf_in_place(91, 76, 76, 77);
      k1 = Z.leaf_in_place(92, 77, 77, 78);
      Z.seg_forward(warp, 75, 76, 7, +1, 76, k0, 77, k1);
      break;
    case R_FOSSA: {  // fossa chain 88..85 -> 84, then 84 out of the clique {28, 29, 84}
      Z.seg_forward(warp, -1, 88, 4, -1, -1, 0.0, -1, 0.0);
      const double e = Z.E(28), g = e + Z.G(28);
      const double w = fast_rcp(Z.d(84));
      Z.d(84) = w;
      const double rc = Z.r(84);
      Z.leaf(0) = e * e * w;       // dd28
      Z.leaf(1) = -e * w * rc;     // dr28
      Z.leaf(2) = g * g * w;       // dd29
      Z.leaf(3) = g * w * rc;      // dr29
      Z.leaf(4) = -e + e * g * w;  // o(28,29)
      break;
    }
    case R_RADM:
    case R_RADN: {  // radiation pair: iL = iR+1 out of {s_in, iR, iL}, then iR; contributions to the separator
      const int sec = warp == R_RADM ? MOUTH_LAST : NOSE_LAST, iR = warp == R_RADM ? 93 : 95, iL = iR + 1;
      const double e = Z.E(sec), g = e + Z.G(sec);
      const double wL = fast_rcp(Z.d(iL));
      Z.d(iL) = wL;
      const double rL = Z.r(iL);
      double dds = e * e * wL, drs = -e * wL * rL;
      const double dR = Z.d(iR) - g * g * wL, rR = Z.r(iR) - g * wL * rL;
      const double o = -e + e * g * wL;
      const double wR = fast_rcp(dR);
      Z.d(iR) = wR; Z.r(iR) = rR;
      dds += o * o * wR; drs += o * wR * rR;
      Z.leaf(warp == R_RADM ? 5 : 7) = dds; Z.leaf(warp == R_RADM ? 6 : 8) = drs;
      k0 = o;
      break;
    }
    default: break;
  }
}

// 3b: reduced system.  Forward: three separator chains (one warp each), all operands loaded up front and the
// eliminations done in registers; junction; back.
__device__ __forceinline__ void red_elim(double &dx, double rx, double c, double &dy, double &ry) {  // x into y
  const double w = fast_rcp(dx);
  dx = w;
  const double t = c * w;
  dy = fma(-t, c, dy);
  ry = fma(-t, rx, ry);
}
__device__ __forceinline__ void solve_3b_forward(Solver &Z, int warp) {
  if (warp == R_RED_T) {
    const double c0 = Z.side(R_T1, 2), c1 = Z.side(R_T2, 2), c2 = Z.leaf(4), c3 = Z.side(R_T4, 2);
    double d9 = Z.d(9) - Z.side(R_T1, 0), r9 = Z.r(9) - Z.side(R_T1, 1);
    double d19 = Z.d(19) - Z.side(R_T2, 0), r19 = Z.r(19) - Z.side(R_T2, 1);
    double d28 = Z.d(28) - Z.leaf(0), r28 = Z.r(28) - Z.leaf(1);
    double d29 = Z.d(29) - (Z.leaf(2) + Z.side(R_T4, 0)), r29 = Z.r(29) - (Z.leaf(3) + Z.side(R_T4, 1));
    double d40 = Z.d(40), r40 = Z.r(40);
    red_elim(d9, r9, c0, d19, r19);
    red_elim(d19, r19, c1, d28, r28);
    red_elim(d28, r28, c2, d29, r29);
    red_elim(d29, r29, c3, d40, r40);
    Z.d(9) = d9; Z.r(9) = r9; Z.d(19) = d19; Z.r(19) = r19; Z.d(28) = d28; Z.r(28) = r28; Z.d(29) = d29; Z.r(29) = r29;
    Z.d(40) = d40; Z.r(40) = r40;
  } else if (warp == R_RED_M) {
    const double c0 = Z.side(R_M2, 2), c1 = Z.side(R_M1, 2);
    double d41 = Z.d(41) - Z.side(R_M1, 0), r41 = Z.r(41) - Z.side(R_M1, 1);
    double d53 = Z.d(53) - Z.side(R_M2, 0), r53 = Z.r(53) - Z.side(R_M2, 1);
    double d64 = Z.d(64) - Z.leaf(5), r64 = Z.r(64) - Z.leaf(6);
    red_elim(d64, r64, c0, d53, r53);
    red_elim(d53, r53, c1, d41, r41);
    Z.d(64) = d64; Z.r(64) = r64; Z.d(53) = d53; Z.r(53) = r53; Z.d(41) = d41; Z.r(41) = r41;
  } else if (warp == R_RED_N) {
    const double c0 = Z.side(R_N2, 2), c1 = Z.side(R_N1, 2);
    double d65 = Z.d(65) - Z.side(R_N1, 0), r65 = Z.r(65) - Z.side(R_N1, 1);
    double d75 = Z.d(75) - Z.side(R_N2, 0), r75 = Z.r(75) - Z.side(R_N2, 1);
    double d83 = Z.d(83) - Z.leaf(7), r83 = Z.r(83) - Z.leaf(8);
    red_elim(d83, r83, c0, d75, r75);
    red_elim(d75, r75, c1, d65, r65);
    Z.d(83) = d83; Z.r(83) = r83; Z.d(75) = d75; Z.r(75) = r75; Z.d(65) = d65; Z.r(65) = r65;
  }
}
// Junction at the velum: clique {40, 41, 65} through section 40.
__device__ __forceinline__ void solve_velum_junction(Solver &Z) {
  const double e = Z.E(40), g = e + Z.G(40);
  const double w65 = fast_rcp(Z.d(65));
  double d40 = Z.d(40) - e * e * w65;
  const double d41 = Z.d(41) - g * g * w65;
  const double o = -e + e * g * w65;
  double r40 = Z.r(40) + e * w65 * Z.r(65);
  const double r41 = Z.r(41) - g * w65 * Z.r(65);
  const double w41 = fast_rcp(d41);
  d40 -= o * o * w41;
  r40 -= o * w41 * r41;
  const double u40 = fast_div(r40, d40);
  const double u41 = (r41 - o * u40) * w41;
  const double u65 = (Z.r(65) + e * u40 - g * u41) * w65;
  Z.r(40) = u40; Z.r(41) = u41; Z.r(65) = u65;
}
__device__ __forceinline__ void solve_3b_back(Solver &Z, int warp) {
  if (warp == R_RED_T) {
    const double c0 = Z.side(R_T1, 2), c1 = Z.side(R_T2, 2), c2 = Z.leaf(4), c3 = Z.side(R_T4, 2);
    const double w9 = Z.d(9), w19 = Z.d(19), w28 = Z.d(28), w29 = Z.d(29);
    const double r9 = Z.r(9), r19 = Z.r(19), r28 = Z.r(28), r29 = Z.r(29), u40 = Z.r(40);
    const double u29 = (r29 - c3 * u40) * w29;
    const double u28 = (r28 - c2 * u29) * w28;
    const double u19 = (r19 - c1 * u28) * w19;
    const double u9 = (r9 - c0 * u19) * w9;
    Z.r(29) = u29; Z.r(28) = u28; Z.r(19) = u19; Z.r(9) = u9;
  } else if (warp == R_RED_M) {
    const double c0 = Z.side(R_M2, 2), c1 = Z.side(R_M1, 2);
    const double w53 = Z.d(53), w64 = Z.d(64), r53 = Z.r(53), r64 = Z.r(64), u41 = Z.r(41);
    const double u53 = (r53 - c1 * u41) * w53;
    const double u64 = (r64 - c0 * u53) * w64;
    Z.r(53) = u53; Z.r(64) = u64;
  } else if (warp == R_RED_N) {
    const double c0 = Z.side(R_N2, 2), c1 = Z.side(R_N1, 2);
    const double w75 = Z.d(75), w83 = Z.d(83), r75 = Z.r(75), r83 = Z.r(83), u65 = Z.r(65);
    const double u75 = (r75 - c1 * u65) * w75;
    const double u83 = (r83 - c0 * u75) * w83;
    Z.r(75) = u75; Z.r(83) = u83;
  }
}
__device__ __forceinline__ void solve_3c(Solver &Z, int warp, double k0) {
  switch (warp) {
    case R_T0: Z.seg_back(-1, 0, 9, +1); break;
    case R_T1: Z.seg_back(9, 10, 9, +1); break;
    case R_T2: Z.seg_back(19, 20, 8, +1); break;
    case R_T4: Z.seg_back(29, 30, 10, +1); break;
    case R_M1: Z.seg_back(41, 42, 11, +1); break;
    case R_M2: Z.seg_back(53, 54, 10, +1); break;
    case R_N1: Z.seg_back(65, 66, 9, +1); Z.leaf_back(89, 73, 73, 74); Z.leaf_back(90, 74, 74, 75); break;
    case R_N2: Z.seg_back(75, 76, 7, +1); Z.leaf_back(91, 76, 76, 77); Z.leaf_back(92, 77, 77, 78); break;
    case R_FOSSA: Z.leaf_back(84, 28, 28, 29); Z.seg_back(-1, 88, 4, -1); break;
    case R_RADM:
    case R_RADN: {
      const int sec = warp == R_RADM ? MOUTH_LAST : NOSE_LAST, iR = warp == R_RADM ? 93 : 95, iL = iR + 1;
      const double e = Z.E(sec), g = e + Z.G(sec);
      const double us = Z.r(sec);  // current into the radiating section carries the section's index
      const double uR = (Z.r(iR) - k0 * us) * Z.d(iR);
      Z.r(iR) = uR;
      Z.r(iL) = (Z.r(iL) + e * us - g * uR) * Z.d(iL);
      break;
    }
    default: break;
  }
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
// ago, kept in shared memory [i][lane].
template <int ORDER>
__device__ __forceinline__ double iir_step(const double *a, const double *b, double (*x)[LANES], double (*y)[LANES],
                                           int lane, double in) {
  double sum = a[0] * in;
#pragma unroll
  for (int i = 1; i <= ORDER; i++) { sum += a[i] * x[i - 1][lane]; sum += b[i] * y[i - 1][lane]; }
#pragma unroll
  for (int i = ORDER - 1; i > 0; i--) { x[i][lane] = x[i - 1][lane]; y[i][lane] = y[i - 1][lane]; }
  x[0][lane] = in; y[0][lane] = sum;
  return sum;
}

// Filter for the sound radiated from the skin above the glottis (TdsModel.cpp:494-510, 544).
__device__ __constant__ double c_tone_a[5] = {5.027640021717718e-007, -7.995535578908732e-007,
                                              2.967895557191014e-007, 0.0, 0.0};
__device__ __constant__ double c_tone_b[5] = {0.0, 3.986308869708467, -5.959669638387298, 3.960408461107104,
                                              -0.987047716233603};

// ---- the kernel -----------------------------------------------------------------
// Output of the step whose solution sits in sol[buf]: radiated flow (TdsModel.cpp:687-705), pressure
// differentiator + Chebyshev low-pass + decimation (Synthesizer.cpp:614-627).  `n_done` = steps finished.
__device__ __forceinline__ void emit_output(const TdsParams &P, TdsSmem &S, int lane, int buf, int n_done, bool active,
                                            int tract, int n_out, double *audio, double *flow, uint32_t &bad) {
  double u = 0.0;
  u += S.sol[buf][93][lane]; u += S.sol[buf][94][lane];
  u += S.sol[buf][95][lane]; u += S.sol[buf][96][lane];
  if (P.skin) u += iir_step<4>(c_tone_a, c_tone_b, S.toneX, S.toneY, lane, S.p4[3][lane]);
  if (!isfinite(u)) bad |= 1u;
  const double outP = (u - S.prevFlow[lane]) / P.dt;
  S.prevFlow[lane] = u;
  const double filt = iir_step<8>(P.cheb_a, P.cheb_b, S.chebX, S.chebY, lane, outP);
  if (active) {
    if (flow) flow[(size_t)tract * P.total_steps + (n_done - 1)] = u;
    if ((n_done % P.sr_factor) == 0) audio[(size_t)tract * n_out + (n_done / P.sr_factor - 1)] = (filt * 0.004) / 32767;
  }
}


// Slot table entry for dynamic (isDyn) slot index q = 3*warp + k, or static slot index q.
__device__ void build_slot(const TdsParams &P, bool isDyn, int q, SlotTab &T) {
  const bool valid = isDyn ? true : q < NSTAT;
  const int s = isDyn ? dyn_slot_section(q) : (valid ? stat_slot_section(q) : NS);
  const int A = valid ? src_of(s) : NS;
  int o0, o1;
  outs_of(valid ? s : FOSSA_LAST, o0, o1);
  const int sib = valid ? sibling_of(s) : -1;
  SlotTab t{};
  t.sec = (uint8_t)s; t.aSec = (uint8_t)A;
  t.out0 = (uint8_t)o0; t.out1 = (uint8_t)o1;
  t.urbOwn = (uint8_t)(valid && urb_slot(s) >= 0 ? urb_slot(s) : 16);
  t.urbSib = (uint8_t)(sib >= 0 ? urb_slot(sib) : 17);
  t.p4 = (uint8_t)(valid && s >= TRACHEA_LAST && s <= PHAR0 ? s - TRACHEA_LAST : 4);
  int fl = valid ? FL_VALID : 0;
  if (valid && o1 != NC) fl |= FL_OUT1;
  if (sib >= 0) fl |= FL_SIB;
  if (valid && urb_slot(s) >= 0) fl |= FL_URB;
  if (valid && s >= TRACHEA_LAST && s <= PHAR0) fl |= FL_P4;
  if (isDyn) {
    t.own = (uint8_t)q;
    t.aRec = (uint8_t)dyn_record_of(A);
    t.wall = 0;
    if (P.soft_walls && s != GLOT_LO && s != GLOT_UP) fl |= FL_WALL;
    if (s > PHAR0 && s <= MOUTH_LAST && sib < 0) fl |= FL_JUNCTION;
    if (s == GLOT_LO) fl |= FL_GLO;
    if (s == GLOT_UP) fl |= FL_GUP;
    if (s == 28 || s == 40 || s == 64) fl |= FL_G3;
  } else {
    t.own = (uint8_t)s;
    t.aRec = (uint8_t)A;
    t.wall = (uint8_t)(P.soft_walls && valid ? wall_type(s) : 3);
    if (s == 0) fl |= FL_ROW0;
  }
  t.flags = (uint16_t)fl;
  T = t;
}

__global__ void __launch_bounds__(TDS_THREADS, 1)
tds_batch_kernel(const TdsParams P, const int32_t *__restrict__ bounds, const afs_tube_frame *__restrict__ tubes,
                 const double *__restrict__ glottis, double *__restrict__ audio, double *__restrict__ flow,
                 uint32_t *__restrict__ status, double *__restrict__ dump) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TdsSmem &S = *reinterpret_cast<TdsSmem *>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int tract_raw = blockIdx.x * P.lanes_per_cta + lane;
  const bool active = lane < P.lanes_per_cta && tract_raw < P.n_tracts;
  const int tract = tract_raw < P.n_tracts ? tract_raw : P.n_tracts - 1;
  const afs_tube_frame *my_frames = tubes + (size_t)tract * P.n_frames;
  const double *my_glottis = glottis + (size_t)tract * P.n_frames * AFS_NUM_GLOTTIS_CONTROL;
  const int n_out = P.total_steps / P.sr_factor;

  // ---- prologue: static records, slot table, zero state ------------------------
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
  if (lane < KD + KS) {
    const int k = lane;
    build_slot(P, k < KD, 3 * warp + (k < KD ? k : k - KD), S.slot[warp][k]);
  }
  for (int i = warp; i < NCX; i += NWS) { S.sol[0][i][lane] = 0.0; S.sol[1][i][lane] = 0.0; if (i < NC) S.diag[i][lane] = 1.0; }
  for (int i = warp; i < NSX; i += NWS) S.D[i][lane] = 0.0;
  if (warp == 0) {
    for (int i = 0; i < 18; i++) S.urb[i][lane] = 0.0;
    for (int i = 0; i < 5; i++) S.p4[i][lane] = 0.0;
    for (int i = 0; i < 4; i++) { S.gx[i][lane] = 0.0; S.toneX[i][lane] = 0.0; S.toneY[i][lane] = 0.0; }
    for (int i = 0; i < 8; i++) { S.chebX[i][lane] = 0.0; S.chebY[i][lane] = 0.0; }
    S.prevFlow[lane] = 0.0;
  }
  __syncthreads();
  // per-lane copies of the static records a dynamic slot reads: 69, 84 (own) and 22 (source of 23)
  if (warp < 3) {
    const int s = warp == 0 ? 69 : warp == 1 ? FOSSA0 : TRACHEA_LAST, x = warp == 0 ? DX_69 : warp == 1 ? DX_84 : DX_22;
    const SRec r = S.srec[s];
    S.E[x][lane] = r.E; S.L[x][lane] = r.L; S.R[x][lane] = s == FOSSA0 ? r.R0 : r.R1; S.invr[x][lane] = r.invr;
    S.alpha[x][lane] = r.alpha;
  }

  // ---- per-thread recurrent state (registers) ----------------------------------
  SecState dyn[KD], sta[KS];
#pragma unroll
  for (int k = 0; k < KD; k++) dyn[k] = SecState{0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < KS; k++) sta[k] = SecState{0, 0, 0, 0, 0, 0};
  double cRad[4] = {0.0, 0.0, 0.0, 0.0};  // history terms of the radiation currents 93..96 (warp NWS-1)
  double gbf = 0.0;                        // glottalBernoulliFactor (owner of section 24)
  uint32_t bad = 0;
  const bool has_static = warp < NWS - 1;  // warp 15 owns 68, 69, 84 and the four radiation currents instead

  int interval = -1, i0 = 0, i1 = 0;  // current control interval [i0, i1)

  // Stages the control frames of the next interval (all warps; surrounded by barriers by the caller).
  auto stage_interval = [&]() {
    interval++;
    i0 = bounds[interval]; i1 = bounds[interval + 1];
    for (int item = warp; item < 2 * AFS_NUM_PM_SECTIONS + 2; item += NWS) {
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
  };
  // Phase 1 of step `step` for all owned sections.
  auto phase1_all = [&](int step) {
    const double ratio = (double)(step - i0) / (double)(i1 - i0);
    const double ratio1 = 1.0 - ratio;
#pragma unroll
    for (int k = 0; k < KD; k++) phase1_dyn(P, S, lane, 3 * warp + k, S.slot[warp][k], dyn[k], ratio, ratio1, step & 1);
    if (has_static) {
#pragma unroll
      for (int k = 0; k < KS; k++) phase1_stat(P, S, lane, S.slot[warp][KD + k], sta[k]);
    }
  };

  if (P.total_steps > 0) {
    __syncthreads();
    stage_interval();
    __syncthreads();
    if (warp == W_GLOTTIS) {  // glottis geometry of step 0
      double ctl[AFS_NUM_GLOTTIS_CONTROL], len[2], ar[2];
      for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++) ctl[c] = S.gctl[0][c][lane];
      const GlottisState G{0.0, 0.0, 0.0, 0.0};
      GlottisDims GD;
      glottis_geometry(P, ctl, G, len, ar, GD);
      S.gdim[0][lane] = GD.Q; S.gdim[1][lane] = GD.cordLength; S.gdim[2][lane] = GD.th0; S.gdim[3][lane] = GD.th1;
      S.gdim[4][lane] = GD.ol0; S.gdim[5][lane] = GD.ol1; S.gdim[6][lane] = GD.cl0; S.gdim[7][lane] = GD.cl1;
      S.gdim[8][lane] = GD.mz0; S.gdim[9][lane] = GD.mz1;
      S.glLen[0][0][lane] = len[0]; S.glLen[0][1][lane] = len[1];
      S.glArea[0][0][lane] = ar[0]; S.glArea[0][1][lane] = ar[1];
    }
    __syncthreads();
    phase1_all(0);
    __syncthreads();
  }
#ifdef AFS_TDS_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#define AFS_TICK(k) do { const long long now__ = clock64(); tacc[k] += now__ - tlast; tlast = now__; } while (0)
#else
#define AFS_TICK(k) do { } while (0)
#endif

  for (int n = 0; n < P.total_steps; n++) {
    const int cur = n & 1, prev = cur ^ 1;
    AFS_TICK(0);
    // ======================= phase 2: matrix rows =============================
    {
      const double ratio = (double)(n - i0) / (double)(i1 - i0);
      const double psrc = (1.0 - ratio) * S.gctl[0][G_PRESSURE][lane] + ratio * S.gctl[1][G_PRESSURE][lane];
#pragma unroll
      for (int k = 0; k < KD; k++) phase2_row<true>(P, S, lane, S.slot[warp][k], dyn[k], prev, cur, psrc, gbf);
      if (has_static) {
#pragma unroll
        for (int k = 0; k < KS; k++) phase2_row<false>(P, S, lane, S.slot[warp][KD + k], sta[k], prev, cur, psrc, gbf);
      } else {
        phase2_radiation(P, S, lane, MOUTH_LAST, 93, cRad[0], cRad[1], cur);
        phase2_radiation(P, S, lane, NOSE_LAST, 95, cRad[2], cRad[3], cur);
      }
    }
    AFS_TICK(1);
    __syncthreads();
    AFS_TICK(2);

    // ======================= phase 3: solve (+ glottis, + output) =============
    {
      Solver Z{S, lane, cur, P.idt};
      double k0 = 0.0, k1 = 0.0;
      if (warp == W_GLOTTIS) {
        // incTime for this step (pressures of the previous step), then the geometry step n+1 will use
        const double ratio = (double)(n - i0) / (double)(i1 - i0), ratio1 = 1.0 - ratio;
        double ctl[AFS_NUM_GLOTTIS_CONTROL], p[4], len[2], ar[2];
        GlottisState G{S.gx[0][lane], S.gx[1][lane], S.gx[2][lane], S.gx[3][lane]};
        for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++) ctl[c] = ratio1 * S.gctl[0][c][lane] + ratio * S.gctl[1][c][lane];
        for (int c = 0; c < 4; c++) p[c] = S.p4[c][lane];
        GlottisDims GD;
        GD.Q = S.gdim[0][lane]; GD.cordLength = S.gdim[1][lane]; GD.th0 = S.gdim[2][lane]; GD.th1 = S.gdim[3][lane];
        GD.ol0 = S.gdim[4][lane]; GD.ol1 = S.gdim[5][lane]; GD.cl0 = S.gdim[6][lane]; GD.cl1 = S.gdim[7][lane];
        GD.mz0 = S.gdim[8][lane]; GD.mz1 = S.gdim[9][lane]; GD.mw0 = 0.0; GD.mw1 = 0.0;
        glottis_inc_time(P, ctl, G, p, GD);
        const double rn = (double)(n + 1 - i0) / (double)(i1 - i0), rn1 = 1.0 - rn;
        for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++) ctl[c] = rn1 * S.gctl[0][c][lane] + rn * S.gctl[1][c][lane];
        glottis_geometry(P, ctl, G, len, ar, GD);
        S.gdim[0][lane] = GD.Q; S.gdim[1][lane] = GD.cordLength; S.gdim[2][lane] = GD.th0; S.gdim[3][lane] = GD.th1;
        S.gdim[4][lane] = GD.ol0; S.gdim[5][lane] = GD.ol1; S.gdim[6][lane] = GD.cl0; S.gdim[7][lane] = GD.cl1;
        S.gdim[8][lane] = GD.mz0; S.gdim[9][lane] = GD.mz1;
        S.gx[0][lane] = G.x0; S.gx[1][lane] = G.x1; S.gx[2][lane] = G.xp0; S.gx[3][lane] = G.xp1;
        S.glLen[prev][0][lane] = len[0]; S.glLen[prev][1][lane] = len[1];
        S.glArea[prev][0][lane] = ar[0]; S.glArea[prev][1][lane] = ar[1];
      } else if (warp == W_OUTPUT) {
        if (n > 0) emit_output(P, S, lane, prev, n, active, tract, n_out, audio, flow, bad);
      } else {
        solve_3a(Z, warp, k0, k1);
      }
      AFS_TICK(3);
      __syncthreads();
      solve_3b_forward(Z, warp);
      __syncthreads();
      if (warp == R_RED_T) solve_velum_junction(Z);
      __syncthreads();
      solve_3b_back(Z, warp);
      __syncthreads();
      solve_3c(Z, warp, k0);
    }
    AFS_TICK(4);
    __syncthreads();
    AFS_TICK(5);

    // ======================= control frames of the next interval ==============
    if (n + 1 == i1 && n + 1 < P.total_steps) {
      stage_interval();
      __syncthreads();
    }

    // ======================= phase 4 (update) fused with phase 1 of step n+1 ===
#pragma unroll
    for (int k = 0; k < KD; k++) phase4_update<true>(P, S, lane, S.slot[warp][k], dyn[k], cur);
    if (has_static) {
#pragma unroll
      for (int k = 0; k < KS; k++) phase4_update<false>(P, S, lane, S.slot[warp][KD + k], sta[k], cur);
    } else {  // radiation currents (TdsModel.cpp:2063-2072)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double u_new = S.sol[cur][93 + j][lane];
        const double ur_new = fma(u_new, P.idt, -cRad[j]);
        cRad[j] = fma(ur_new, P.th, u_new * P.idt);
      }
    }
    if (dump && P.dump_step == n + 1) {  // diagnostics: state of tract 0 after step n+1
      if (blockIdx.x == 0 && lane == 0) {
#define AFS_DUMP(T_, st)                                                                        \
  do {                                                                                          \
    if ((T_).flags & FL_VALID) {                                                                \
      const int sec = (T_).sec;                                                       \
      const double u = S.sol[cur][sec][lane];                                                   \
      dump[sec] = (st).p; dump[93 + sec] = (st).pr; dump[186 + sec] = (st).wc;                   \
      dump[279 + sec] = u; dump[376 + sec] = (st).c / P.th - u * P.idt / P.th;                   \
    }                                                                                           \
  } while (0)
#pragma unroll
        for (int k = 0; k < KD; k++) AFS_DUMP(S.slot[warp][k], dyn[k]);
#pragma unroll
        for (int k = 0; k < KS; k++) if (has_static) AFS_DUMP(S.slot[warp][KD + k], sta[k]);
#undef AFS_DUMP
        if (!has_static) {
          for (int j = 0; j < 4; j++) {
            const double u = S.sol[cur][93 + j][lane];
            dump[279 + 93 + j] = u; dump[376 + 93 + j] = cRad[j] / P.th - u * P.idt / P.th;
          }
        }
      }
    }
    AFS_TICK(6);
    if (n + 1 < P.total_steps) phase1_all(n + 1);
    AFS_TICK(7);
    __syncthreads();
  }
#ifdef AFS_TDS_TIMING
  if (dump && blockIdx.x == 0 && lane == 0) for (int k = 0; k < 8; k++) dump[warp * 8 + k] = (double)tacc[k];
#endif
  // epilogue: output of the last step
  if (warp == W_OUTPUT) {
    if (P.total_steps > 0)
      emit_output(P, S, lane, (P.total_steps - 1) & 1, P.total_steps, active, tract, n_out, audio, flow, bad);
    if (active && status) status[tract] = bad;
  }
}

// ---- host side ------------------------------------------------------------------
int tds_prepare_launch(const afs_tds_options *o, int n_tracts, int n_frames, const int32_t *frame_len,
                       double fs_tds_hz, int sr_factor, int sm_count, TdsLaunch *L) {
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
  // tracts per CTA: spread the batch over all SMs, at most 32 (one per lane)
  int lpc = (n_tracts + sm_count - 1) / (sm_count > 0 ? sm_count : 1);
  if (lpc > LANES) lpc = LANES;
  if (lpc < 1) lpc = 1;
  P.lanes_per_cta = lpc;
  P.turbulence = o->turbulence_losses; P.soft_walls = o->soft_walls; P.skin = o->radiation_from_skin;
  P.fossa = o->piriform_fossa; P.ilc = o->inner_length_corrections;
  P.fsar = o->flow_separation_area_ratio;
  P.k_ent = o->glottis_loss_option == AFS_ENTRANCE_LOSS_VAN_DEN_BERG ? 1.375 : 1.0;
  const double dt = 1.0 / fs_tds_hz;
  P.dt = dt; P.dtth = dt * kTheta; P.dtth1 = dt * kTheta1; P.idt = 1.0 / (dt * kTheta); P.th = kTheta1 / kTheta;
  // wall types: standard (Tube.cpp:8-10), trachea (:103-105), sinus (:208-210)
  const double Mw[3] = {2.1, 0.25, 0.0}, Bw[3] = {800.0, 1000.0, 6500.0}, Kw[3] = {84500.0, 84500.0, 84500.0};
  for (int t = 0; t < 3; t++) {
    const double lq = Mw[t] / (dt * dt * kTheta * kTheta), rq = Bw[t] / (dt * kTheta);
    P.kalpha[t] = 1.0 / (lq + rq + Kw[t]);
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
  P.dump_step = 0;
  L->params.resize(sizeof P);
  memcpy(L->params.data(), &P, sizeof P);
  L->n_tracts = n_tracts; L->n_frames = n_frames;
  return AFS_OK;
}

cudaError_t launch_tds(const TdsLaunch &L, const int32_t *d_bounds, const afs_tube_frame *d_tubes,
                       const double *d_glottis, double *d_audio, double *d_flow, uint32_t *d_status,
                       double *d_dump, int dump_step, cudaStream_t stream) {
  TdsParams P;
  memcpy(&P, L.params.data(), sizeof P);
  P.dump_step = dump_step;
  static_assert(sizeof(TdsSmem) <= 227 * 1024, "TdsSmem exceeds the 227 KB a CTA can have on sm_100a");
  cudaError_t e = cudaFuncSetAttribute(tds_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(TdsSmem));  // per device; cheap
  if (e != cudaSuccess) return e;
  const int ctas = (P.n_tracts + P.lanes_per_cta - 1) / P.lanes_per_cta;
  tds_batch_kernel<<<ctas, TDS_THREADS, sizeof(TdsSmem), stream>>>(P, d_bounds, d_tubes, d_glottis, d_audio, d_flow,
                                                                  d_status, d_dump);
  return cudaGetLastError();
}

size_t tds_smem_bytes() { return sizeof(TdsSmem); }

}  // namespace afs
