// tl_fast_kernels.cu -- K2f `tl_sweep_fast_direct_kernel` / `tl_sweep_fast_kernel` for sm_100a: the transmission-line sweep of
// TlModel::getSpectrum / prepareCalculations (reference TlModel.cpp:123-197, 484-676, 783-932) in the
// ALGEBRAICALLY REDUCED form, selected by afs_tl_options.arithmetic = AFS_TL_ARITHMETIC_FAST.
//
// tl_kernels.cu repeats the reference's operations one by one (19 IEEE divisions per section, no FMA) and is
// bit-identical to it; that costs 17 k FP64 instructions per (area function, bin).  The north-star bar for TL is
// 1e-9 relative, so this unit computes the same chain products with the arithmetic that the mathematics needs:
//   * a lumped T-section [[1+Za/Zb, Za^2/Zb+2Za],[1/Zb, 1+Za/Zb]] (TlModel.cpp:880-893) is the ladder
//     series(Za) * shunt(Yb) * series(Za) with Yb = 1/Zb, so the running product K is updated by
//     K.B += K.A*Z, K.D += K.C*Z (series) and K.A += K.B*Y, K.C += K.D*Y (shunt): 16 FMAs per section, and the
//     two half-sections that meet at a junction (plus the Sondhi junction inductance and the static
//     pressure-drop resistances, TlModel.cpp:573-611) merge into ONE series element;
//   * Zb is never formed: the wall admittance 1/((Bw + j(w Mw - Kw/w))/(l*circ)) (TlModel.cpp:813-822) is
//     l*circ times a factor that depends on the bin and on the wall class only (two classes: trachea and the
//     rest, Tube.cpp:8-10, 94-119) -- two reciprocals per BIN instead of six divisions per SECTION;
//   * all bin-independent section terms are staged once per CTA in shared memory, options folded in as zeros.
//   * for the flow-source transfer function of a trachea..pharynx section (the spectrum getFormants and
//     getImpulseResponse ask for) the product "rest of the chain after the source section" is accumulated
//     directly instead of as inverse(P[section]) * P[40] (TlModel.cpp:1336-1338): no cancellation, three live
//     matrices instead of five, the chains walked in straight-line code (tl_sweep_fast_direct_kernel); every
//     other spectrum type / source section goes through a small interpreter of a host-built chain program
//     (tl_sweep_fast_kernel).
// About 2.6 k FP64 instructions per bin (6.5x fewer), 2.25 M spectra/s at 4096 x 2048 (K2: 0.27 M).  Against the
// bit-identical kernel: 99.9 % of the points within 2.7e-10, maximum 2.85e-9 -- at those points the REFERENCE
// arithmetic is that far from an extended-precision evaluation of its own formulas and this kernel 1e-12
// (tools/tl_truth.py, tests/test_tl_gpu.py::test_fast_arithmetic_*; bench.py reports the error of every run).
// Wave-propagation mode and the RADIATION spectrum are served by tl_kernels.cu (afs_api.cu routes them).
#include "afs_internal.h"

namespace afs {
namespace {

#include "tl_cx.cuh"

// z = n / d as n * conj(d) / |d|^2 (operands of this model are far from the overflow limits)
__device__ __forceinline__ cx cx_div(cx n, cx d) {
  const double r = 1.0 / (d.re * d.re + d.im * d.im);
  return cx_make((n.re * d.re + n.im * d.im) * r, (n.im * d.re - n.re * d.im) * r);
}
__device__ __forceinline__ cx cx_rdiv(double x, cx d) {
  const double r = x / (d.re * d.re + d.im * d.im);
  return cx_make(d.re * r, -d.im * r);
}

#include "tl_algebra.cuh"

constexpr int TLF_THREADS = 128;
#ifndef AFS_TLF_MINB
#define AFS_TLF_MINB 4  // 128 registers, 16 warps/SM (96 registers / 20 warps spills: 1.96 vs 1.70 ms; unrolling the section loops by 2: 1.73 ms)
#endif

struct __align__(16) TlElem {
  double rsq;  // boundary layer: Re(Za) = rsq*sqrt(w)     (0 when the option is off)
  double rf;   // Hagen-Poiseuille resistance of the half section (0 when off)
  double ls;   // Im(Za) = w*ls
  double cp;   // Im(Yb) = w*cp (+ wall)
  double gp;   // heat conduction: Re(Yb) = gp*sqrt(w) (0 when off); sinus: the constant conductance
  double lc;   // length*circumference: wall admittance scale (0 without soft walls)
};

struct TlFastShared {
  TlElem e[NS];
  double2 pre[NS];      // series element in FRONT of section k: (static pressure-drop resistance (TlModel.cpp:592-611),
                        // Sondhi junction inductance (:943-977)); all zero unless `has_pre`
  double raw_area[NS];
  double rad[2][2];     // radiation constants of the mouth / nose opening
  int has_pre;
  unsigned short prog[128];  // chain program (OP_*), copied from the launch parameters
};
struct TlFastProgram {
  unsigned short op[128];
};

// Chain program, built once per CTA: a plain entry (< 0x80) is a section "series, shunt"; everything that makes
// a step more than that is its own entry, so that the hot loop carries nothing but the running product.
constexpr int OP_K = 0x7f;
constexpr int OP_BRANCH = 0x100;  // | first section of a chain (TlModel.cpp:565,616,648; 84: the fossa chain)
constexpr int OP_WALL = 0x200;    // | wall class of the following sections (1: trachea)
constexpr int OP_SINUS = 0x300;   // | sinus section: Helmholtz shunt (TlModel.cpp:659-670)
constexpr int OP_STORE = 0x400;   // | k: matrixProduct[k] is asked for by the transfer-function algebra
constexpr int OP_FOSSA = 0x500;   // fossa shunt follows (TlModel.cpp:586-590)
constexpr int OP_END = 0x600;

__device__ __forceinline__ void series(m22 &K, cx Z) {  // K := K * [[1, Z], [0, 1]]
  K.B.re = fma(K.A.re, Z.re, fma(-K.A.im, Z.im, K.B.re));
  K.B.im = fma(K.A.re, Z.im, fma(K.A.im, Z.re, K.B.im));
  K.D.re = fma(K.C.re, Z.re, fma(-K.C.im, Z.im, K.D.re));
  K.D.im = fma(K.C.re, Z.im, fma(K.C.im, Z.re, K.D.im));
}
__device__ __forceinline__ void shunt(m22 &K, cx Y) {  // K := K * [[1, 0], [Y, 1]]
  K.A.re = fma(K.B.re, Y.re, fma(-K.B.im, Y.im, K.A.re));
  K.A.im = fma(K.B.re, Y.im, fma(K.B.im, Y.re, K.A.im));
  K.C.re = fma(K.D.re, Y.re, fma(-K.D.im, Y.im, K.C.re));
  K.C.im = fma(K.D.re, Y.im, fma(K.D.im, Y.re, K.C.im));
}

// TlModel::getRadiationImpedance (TlModel.cpp:707-771) with the area-dependent factors c[0..1] staged per CTA
__device__ __forceinline__ cx radiation_fast(int type, double omega, const double *c) {
  if (omega < 0.0001) omega = 0.0001;
  if (type == AFS_PISTONINWALL_RADIATION) return cx_make(c[0] * omega * omega, c[1] * omega);
  if (type == AFS_PISTONINSPHERE_RADIATION) {
    const double freq = omega / (2.0 * kPi);
    const double K = freq < 1600 ? (0.6 * freq) / 1600.0 + 1.0 : 1.6;
    return cx_make((kRho * omega * omega * K) / (4.0 * kPi * kC), omega * c[1]);
  }
  if (type == AFS_PARALLEL_RADIATION) {  // R || jX = R X (X + jR) / (R^2 + X^2)
    const double R = c[0], X = omega * c[1];
    const double s = (R * X) / (R * R + X * X);
    return cx_make(s * X, s * R);
  }
  return cx_make(0, 0);
}
__device__ void radiation_constants(int type, double area, double *c) {
  if (area < kMinArea) area = kMinArea;
  c[0] = c[1] = 0.0;
  if (type == AFS_PISTONINWALL_RADIATION) {
    const double a = sqrt(area / kPi), analog = (kRho * kC) / area;
    c[0] = analog * 0.5 * (a / kC) * (a / kC);
    c[1] = analog * 8.0 * (a / kC) / (3.0 * kPi);
  } else if (type == AFS_PISTONINSPHERE_RADIATION) {
    c[1] = (8.0 * kRho) / (3.0 * kPi * sqrt(kPi * area));
  } else if (type == AFS_PARALLEL_RADIATION) {
    c[0] = (128.0 * kRho * kC) / (9.0 * kPi * kPi * area);
    c[1] = (8.0 * kRho) / (3.0 * kPi * sqrt(kPi * area));
  }
}

// Expand one frame to the 93 sections (Tube.cpp:323-416) and their bin-independent element factors
// (TlModel.cpp:796-811, 849-866), thread k < NS.
__device__ void stage_fast(const TlParams &p, const afs_tube_frame *f, TlFastShared &S, int k) {
  double area, length, volume;
  if (k >= PHAR0 && k <= MOUTH_LAST) {
    length = f->length_cm[k - PHAR0];
    area = f->area_cm2[k - PHAR0];
    if (area < kMinArea) area = kMinArea;
    volume = area * length;
  } else if (k == GLOT_LO || k == GLOT_UP) {
    length = f->glottis_length_cm[k - GLOT_LO];
    area = f->glottis_area_cm2[k - GLOT_LO];
    if (area < kMinArea) area = kMinArea;
    volume = area * length;
  } else {
    const StaticSection s = static_section(k);
    area = s.area; length = s.length; volume = s.volume;
    if (k >= NOSE0 && k < NOSE0 + 4) { area = velum_port_area(f->velum_area_cm2, k - NOSE0); volume = area * length; }
  }
  S.raw_area[k] = area;
  double a = area;
  if (a < kMinArea) a = kMinArea;
  const double circ = 2.0 * sqrt(a * kPi);
  TlElem e;
  if (k >= SINUS0) {
    e.rsq = p.o.boundary_layer ? (p.rs_factor * length * circ) / (a * a) : 0.0;
    e.rf = 0.0;
    e.ls = kRho * length / a;
    e.cp = p.cp_factor * volume;
    e.gp = p.sinus_gp[k - SINUS0];
    e.lc = 0.0;
  } else {
    e.rsq = p.o.boundary_layer ? (p.rs_factor * 0.5 * length * circ) / (a * a) : 0.0;
    e.rf = p.o.hagen_resistance ? (4.0 * kMu * length * kPi) / (a * a) : 0.0;
    e.ls = kRho * (0.5 * length) / a;
    e.cp = p.cp_factor * length * a;
    e.gp = p.o.heat_conduction ? p.gp_factor * length * circ : 0.0;
    e.lc = p.o.soft_walls ? length * circ : 0.0;
  }
  S.e[k] = e;
}


// The chain program depends on the options and on the source section only: built on the host per launch.
// [fossa 84..88 when the option is on,] 0..83 with the sinus sections 89..92 inserted after their coupling
// sections (Tube.cpp:35-38).
void build_program(const TlParams &p, TlFastProgram &prog) {
  int n = 0;
  unsigned short *o = prog.op;
  o[n++] = OP_WALL | 0;
  if (p.o.piriform_fossa) {
    o[n++] = OP_BRANCH | FOSSA0;
    for (int k = FOSSA0; k <= FOSSA_LAST; k++) {
      o[n++] = (unsigned short)k;
      if (k == p.section - 1 || k == p.section) o[n++] = (unsigned short)(OP_STORE | k);
    }
  }
  for (int k = 0; k <= NOSE_LAST; k++) {
    if (k == TRACHEA0 || k == MOUTH0 || k == NOSE0) o[n++] = (unsigned short)(OP_BRANCH | k);
    if (k == TRACHEA0) o[n++] = OP_WALL | 1;
    if (k == GLOT_LO) o[n++] = OP_WALL | 0;
    o[n++] = (unsigned short)k;
    if (k == p.section - 1 || k == p.section || k == PHAR_LAST || k == MOUTH_LAST || k == NOSE_LAST)
      o[n++] = (unsigned short)(OP_STORE | k);
    if (p.o.piriform_fossa && k == PHAR0 + FOSSA_COUPLING) o[n++] = OP_FOSSA;
    if (p.o.paranasal_sinuses) {
      const int kk = k - NOSE0;
      if (kk == 8 || kk == 9 || kk == 11 || kk == 12)
        o[n++] = (unsigned short)(OP_SINUS | (SINUS0 + (kk == 8 ? 0 : kk == 9 ? 1 : kk == 11 ? 2 : 3)));
    }
  }
  while (n < 128) o[n++] = OP_END;
}


// Per-CTA staging: sections, junction terms, static pressure drops, radiation constants, chain program.
__device__ __forceinline__ void stage_all(const TlParams &p, const TlFastProgram *prog, const afs_tube_frame *f,
                                          TlFastShared &S) {
  for (int k = threadIdx.x; k < NS; k += TLF_THREADS) stage_fast(p, f, S, k);
  if (prog) S.prog[threadIdx.x] = prog->op[threadIdx.x];
  __syncthreads();
  if (threadIdx.x < NS) {  // junction inductances need both neighbours
    const int k = threadIdx.x;
    double jl = 0.0;
    if (p.o.inner_length_corrections && k > PHAR0 && k <= MOUTH_LAST) {
      double A1 = S.raw_area[k - 1], A2 = S.raw_area[k];
      if (A1 < kMinArea) A1 = kMinArea;
      if (A2 < kMinArea) A2 = kMinArea;
      const double big = A1 > A2 ? A1 : A2, small = A1 > A2 ? A2 : A1;
      const double a = sqrt(big / kPi), b = sqrt(small / kPi);
      jl = 8.0 * kRho * (1 - b / a) / (3.0 * kPi * kPi * b);
    }
    S.pre[k] = make_double2(0.0, jl);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // mean flow through the tract (TlModel.cpp:526-541), most constricted section (:204-217)
    int ms = PHAR0;
    for (int i = PHAR0 + 1; i <= MOUTH_LAST; i++) if (S.raw_area[i] <= S.raw_area[ms]) ms = i;
    double A_g = S.raw_area[GLOT_LO], A_c = S.raw_area[ms];
    if (A_g < kMinArea) A_g = kMinArea;
    if (A_c < kMinArea) A_c = kMinArea;
    double mf = sqrt(2.0 * p.o.lung_pressure_dPa / (kRho * (1.0 / (A_g * A_g) + 1.0 / (A_c * A_c))));
    if (S.raw_area[NOSE0] > kMinArea) mf = 0.0;
    // the two resistors sit behind sections 22 and ms-1, i.e. in front of 23 and ms; the one behind the last
    // pharynx section is lost in the reference too (the mouth chain restarts from the unit matrix)
    S.has_pre = p.o.inner_length_corrections != 0;
    if (p.o.static_pressure_drops && mf != 0.0) {
      S.has_pre = 1;
      S.pre[GLOT_LO].x += kRho * mf / (A_g * A_g);
      if (ms != MOUTH0) S.pre[ms].x += kRho * mf / (A_c * A_c);
    }
    radiation_constants(p.o.radiation, S.raw_area[MOUTH_LAST], S.rad[0]);
    radiation_constants(p.o.radiation, S.raw_area[NOSE_LAST], S.rad[1]);
  }
  __syncthreads();
}

// Bin-dependent, section-independent factors.
struct BinCtx {
  double omega_raw, omega, sw;
  cx g0, g1;  // wall admittance per unit (length*circumference): 1 / (Bw + j (w Mw - Kw / w)), two wall classes
};
__device__ __forceinline__ BinCtx bin_ctx(const TlParams &p, int bin) {
  BinCtx c;
  const int fi = bin == 0 ? 1 : bin;  // every accessor aliases bin 0 to bin 1 (TlModel.cpp:992,1065,1127,1293)
  c.omega_raw = 2.0 * kPi * p.f0 * (double)fi;
  c.omega = c.omega_raw < 0.0001 ? 0.0001 : c.omega_raw;
  c.sw = sqrt(c.omega);
  const double iw = 1.0 / c.omega;
  const double x0 = c.omega * 2.1 - 84500.0 * iw, x1 = c.omega * 0.25 - 84500.0 * iw;
  const double r0 = 1.0 / (800.0 * 800.0 + x0 * x0), r1 = 1.0 / (1000.0 * 1000.0 + x1 * x1);
  c.g0 = cx_make(800.0 * r0, -x0 * r0);   // Tube.cpp:8-10
  c.g1 = cx_make(1000.0 * r1, -x1 * r1);  // trachea, Tube.cpp:94-119
  return c;
}

// Running state of a chain: product K, series impedance Zs not yet multiplied into K, elements of the last section.
struct Chain {
  m22 K;
  cx Zs, Za, Yb;
};
__device__ __forceinline__ void chain_reset(Chain &c) { c.K = m22_unit(); c.Zs = cx_make(0, 0); }
// one plain section: "series, shunt" (the second half-section stays pending in Zs)
__device__ __forceinline__ void chain_section(const TlFastShared &S, const BinCtx &b, bool pre, cx g, int k, Chain &c) {
  if (pre) {
    const double2 z = S.pre[k];
    c.Zs.re += z.x;
    c.Zs.im = fma(b.omega_raw, z.y, c.Zs.im);
  }
  const TlElem e = S.e[k];
  c.Za = cx_make(fma(e.rsq, b.sw, e.rf), b.omega * e.ls);
  c.Yb = cx_make(fma(e.gp, b.sw, e.lc * g.re), fma(b.omega, e.cp, e.lc * g.im));
  c.Zs = cx_add(c.Zs, c.Za);
  series(c.K, c.Zs);
  shunt(c.K, c.Yb);
  c.Zs = c.Za;
}
__device__ __forceinline__ void chain_shunt(Chain &c, cx Y) {  // a shunt element closes the pending series element
  series(c.K, c.Zs);
  c.Zs = cx_make(0, 0);
  shunt(c.K, Y);
}
// Helmholtz resonator shunt of a paranasal sinus (TlModel.cpp:659-670, 856-879)
__device__ __forceinline__ void chain_sinus(const TlFastShared &S, const BinCtx &b, int k, Chain &c) {
  const TlElem e = S.e[k];
  const cx za = cx_make(e.rsq * b.sw, b.omega * e.ls), yb = cx_make(e.gp, b.omega * e.cp);
  cx q = cx_mul(za, yb);
  q.re += 1.0;
  chain_shunt(c, cx_div(yb, q));
}
__device__ __forceinline__ m22 chain_product(const Chain &c) {  // matrixProduct[k] of the reference: K * series(Za)
  m22 F = c.K;
  series(F, c.Zs);
  return F;
}

}  // namespace

// Generic instantiation: every spectrum type / source section; the chains are walked by a small interpreter of
// the host-built chain program, the five products the algebra may ask for are kept.
__global__ void __launch_bounds__(TLF_THREADS, 3)
tl_sweep_fast_kernel(const TlParams p, const TlFastProgram prog, const afs_tube_frame *__restrict__ tubes,
                     double *__restrict__ out_re, double *__restrict__ out_im) {
  __shared__ TlFastShared S;
  const int tube = blockIdx.y;
  stage_all(p, &prog, tubes + tube, S);
  const bool pre = S.has_pre != 0;
  // a CTA serves every gridDim.x-th group of TLF_THREADS bins of its area function (staging amortised)
  for (int bin = blockIdx.x * TLF_THREADS + threadIdx.x; bin < p.num_freq; bin += gridDim.x * TLF_THREADS) {
    const BinCtx b = bin_ctx(p, bin);
    Prods P;
    P.section = p.section;
    P.sm1 = m22_unit(); P.s = m22_unit();
    P.p40 = m22_unit(); P.p64 = m22_unit(); P.p83 = m22_unit();
    Chain c;
    chain_reset(c);
    c.Za = cx_make(0, 0); c.Yb = cx_make(1, 0);
    cx fossaY = cx_make(0, 0), srcZa = cx_make(0, 0), srcYb = cx_make(1, 0);
    cx g = b.g0;
    const unsigned short *pc = S.prog;
    int op = *pc;
    for (;;) {
      while (op < 0x80) {  // plain sections
        chain_section(S, b, pre, g, op, c);
        op = *++pc;
      }
      if (op == OP_END) break;
      const int k = op & OP_K, what = op & ~OP_K;
      if (what == OP_WALL) {
        g = k ? b.g1 : b.g0;
      } else if (what == OP_BRANCH) {
        // end of the fossa chain (:554-561): input impedance A/C of the closed side branch (the pending series
        // element does not touch A and C)
        if (k == TRACHEA0 && p.o.piriform_fossa) fossaY = cx_div(c.K.C, c.K.A);
        chain_reset(c);
      } else if (what == OP_SINUS) {
        chain_sinus(S, b, k, c);
      } else if (what == OP_STORE) {  // stored BEFORE the series/shunt extras (TlModel.cpp:583)
        const m22 F = chain_product(c);
        if (k == p.section - 1) P.sm1 = F;
        if (k == p.section) { P.s = F; srcZa = c.Za; srcYb = c.Yb; }
        if (k == PHAR_LAST) P.p40 = F;
        if (k == MOUTH_LAST) P.p64 = F;
        if (k == NOSE_LAST) P.p83 = F;
      } else {  // OP_FOSSA
        chain_shunt(c, fossaY);
      }
      op = *++pc;
    }
    P.mouthRad = radiation_fast(p.o.radiation, b.omega_raw, S.rad[0]);
    P.noseRad = radiation_fast(p.o.radiation, b.omega_raw, S.rad[1]);
    cx v = cx_make(0, 0);
    if (p.type == AFS_INPUT_IMPEDANCE) v = input_impedance(P, p.section);
    else if (p.type == AFS_OUTPUT_IMPEDANCE) v = output_impedance(P, p.section);
    else if (p.type == AFS_PRESSURE_SOURCE_TF) v = pressure_source_tf(P, p.section);
    else if (p.type == AFS_FLOW_SOURCE_TF) v = flow_source_tf(P, p.section, srcZa, cx_rdiv(1.0, srcYb));
    const size_t o = (size_t)tube * p.num_freq + bin;
    out_re[o] = v.re;
    out_im[o] = v.im;
  }
}

// Direct instantiation: FLOW_SOURCE_TF with the source in the trachea..pharynx chain (sections 0..40) -- the
// spectrum getFormants and getImpulseResponse ask for.  Straight-line walk of the chains, three live matrices.
__global__ void __launch_bounds__(TLF_THREADS, AFS_TLF_MINB)
tl_sweep_fast_direct_kernel(const TlParams p, const afs_tube_frame *__restrict__ tubes, double *__restrict__ out_re,
                            double *__restrict__ out_im) {
  __shared__ TlFastShared S;
  const int tube = blockIdx.y;
  stage_all(p, nullptr, tubes + tube, S);
  const bool pre = S.has_pre != 0;
  const bool fossa = p.o.piriform_fossa != 0;
  const int s = p.section;
  for (int bin = blockIdx.x * TLF_THREADS + threadIdx.x; bin < p.num_freq; bin += gridDim.x * TLF_THREADS) {
    const BinCtx b = bin_ctx(p, bin);
    Chain c;
    chain_reset(c);
    c.Za = cx_make(0, 0); c.Yb = cx_make(1, 0);
    cx fossaY = cx_make(0, 0);
    // sections lo..hi of the main chain; splits where the wall class changes (22|23) and, with the fossa, behind its
    // coupling section 28, whose shunt follows unless the range ends there and `tail` is false
    auto run = [&](int lo, const int hi_all, const bool tail) {
      while (lo <= hi_all) {
        int hi = hi_all;
        if (lo <= TRACHEA_LAST && hi > TRACHEA_LAST) hi = TRACHEA_LAST;
        if (fossa && lo <= PHAR0 + FOSSA_COUPLING && hi > PHAR0 + FOSSA_COUPLING) hi = PHAR0 + FOSSA_COUPLING;
        const cx g = lo <= TRACHEA_LAST ? b.g1 : b.g0;
        for (int k = lo; k <= hi; k++) chain_section(S, b, pre, g, k, c);
        if (fossa && hi == PHAR0 + FOSSA_COUPLING && (tail || hi < hi_all)) chain_shunt(c, fossaY);
        lo = hi + 1;
      }
    };
    // The two radiating branches first: of the mouth and nose chains only (A*Zrad + B) and (C*Zrad + D) are ever
    // needed (their ratio is the branch's input impedance, and they close the bottom row of the coupled products),
    // so two complex numbers per branch stay live instead of a matrix.
    const cx mouthRad = radiation_fast(p.o.radiation, b.omega_raw, S.rad[0]);
    const cx noseRad = radiation_fast(p.o.radiation, b.omega_raw, S.rad[1]);
    for (int k = MOUTH0; k <= MOUTH_LAST; k++) chain_section(S, b, pre, b.g0, k, c);
    cx mA, mC, nA, nC;
    {
      const m22 F = chain_product(c);
      mA = cx_add(cx_mul(F.A, mouthRad), F.B); mC = cx_add(cx_mul(F.C, mouthRad), F.D);
    }
    // nose with the sinus shunts behind sections 73, 74, 76, 77 (Tube.cpp:35-38)
    chain_reset(c);
    if (p.o.paranasal_sinuses) {
      int k = NOSE0;
      for (; k <= NOSE0 + 8; k++) chain_section(S, b, pre, b.g0, k, c);
      chain_sinus(S, b, SINUS0, c);
      chain_section(S, b, pre, b.g0, k++, c);
      chain_sinus(S, b, SINUS0 + 1, c);
      for (; k <= NOSE0 + 11; k++) chain_section(S, b, pre, b.g0, k, c);
      chain_sinus(S, b, SINUS0 + 2, c);
      chain_section(S, b, pre, b.g0, k++, c);
      chain_sinus(S, b, SINUS0 + 3, c);
      for (; k <= NOSE_LAST; k++) chain_section(S, b, pre, b.g0, k, c);
    } else {
      for (int k = NOSE0; k <= NOSE_LAST; k++) chain_section(S, b, pre, b.g0, k, c);
    }
    {
      const m22 F = chain_product(c);
      nA = cx_add(cx_mul(F.A, noseRad), F.B); nC = cx_add(cx_mul(F.C, noseRad), F.D);
    }
    chain_reset(c);
    if (fossa) {  // side-branch chain (TlModel.cpp:554-561): its input impedance is A/C (closed end)
      for (int k = FOSSA0; k <= FOSSA_LAST; k++) chain_section(S, b, pre, b.g0, k, c);
      fossaY = cx_div(c.K.C, c.K.A);
      chain_reset(c);
    }
    // in front of the source: output impedance of matrixProduct[s-1] (stored before the extras of s-1), lung side
    // shorted (TlModel.cpp:1060-1105, :518)
    cx zout = cx_make(0, 0);
    if (s >= 1) {
      run(TRACHEA0, s - 1, false);
      const m22 F = chain_product(c);
      zout = cx_div(F.B, F.A);
      if (fossa && s - 1 == PHAR0 + FOSSA_COUPLING) chain_shunt(c, fossaY);
    }
    // the source section, then the rest of the chain as its own product R = inverse(P[s]) * P[40]
    run(s, s, false);
    const cx srcZa = c.Za, srcYb = c.Yb;
    chain_reset(c);
    if (fossa && s == PHAR0 + FOSSA_COUPLING) chain_shunt(c, fossaY);
    run(s + 1, PHAR_LAST, true);
    const m22 R = chain_product(c);

    // TlModel::getFlowSourceTF, trachea..pharynx branch (TlModel.cpp:1312-1368); only the bottom rows of the two
    // coupled products are needed: [R * shunt(1/zn) * P64] bottom row times (Zrad, 1) = (R.C + R.D/zn) mA + R.D mC
    const cx zn = cx_div(nA, nC), zm = cx_div(mA, mC);
    const cx zin = cx_add(load_through(R, par(zn, zm)), srcZa);
    zout = cx_add(zout, srcZa);
    cx factor = par(cx_rdiv(1.0, srcYb), zout);
    factor = cx_div(factor, cx_add(zin, factor));
    cx q = cx_add(R.C, cx_div(R.D, zn));
    cx v = cx_div(factor, cx_add(cx_mul(q, mA), cx_mul(R.D, mC)));
    q = cx_add(R.C, cx_div(R.D, zm));
    v = cx_add(v, cx_div(factor, cx_add(cx_mul(q, nA), cx_mul(R.D, nC))));
    const size_t o = (size_t)tube * p.num_freq + bin;
    out_re[o] = v.re;
    out_im[o] = v.im;
  }
}

cudaError_t launch_tl_sweep_fast(const TlParams &p, const afs_tube_frame *d_tubes, double *d_re, double *d_im,
                                 cudaStream_t stream) {
  constexpr int BINS_PER_THREAD = 4;
  dim3 grid((p.num_freq + TLF_THREADS * BINS_PER_THREAD - 1) / (TLF_THREADS * BINS_PER_THREAD), p.n_tubes);
  if (p.type == AFS_FLOW_SOURCE_TF && p.section >= TRACHEA0 && p.section <= PHAR_LAST) {
    tl_sweep_fast_direct_kernel<<<grid, TLF_THREADS, 0, stream>>>(p, d_tubes, d_re, d_im);
  } else {
    TlFastProgram prog;
    build_program(p, prog);
    tl_sweep_fast_kernel<<<grid, TLF_THREADS, 0, stream>>>(p, prog, d_tubes, d_re, d_im);
  }
  return cudaGetLastError();
}

}  // namespace afs
