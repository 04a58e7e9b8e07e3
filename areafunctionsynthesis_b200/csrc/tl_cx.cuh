// tl_cx.cuh -- complex scalars and 2x2 chain matrices shared by the two TL translation units
// (tl_kernels.cu: reference operation order, -fmad=false, bit-identical; tl_fast_kernels.cu: ladder
// recurrences, FMA contraction on, 1e-9).  Included INSIDE an anonymous namespace of each unit: the two
// units are compiled with different floating-point flags, the helpers must not be merged by the linker.
#pragma once

struct cx { double re, im; };
__device__ __forceinline__ cx cx_make(double re, double im) { cx z; z.re = re; z.im = im; return z; }
__device__ __forceinline__ cx cx_add(cx a, cx b) { return cx_make(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cx cx_neg(cx a) { return cx_make(-a.re, -a.im); }
// g++ inline complex product: (ac - bd, ad + bc)
__device__ __forceinline__ cx cx_mul(cx a, cx b) {
  return cx_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}

// Matrix2x2 (reference Matrix2x2.cpp:32-52, 74-86)
struct m22 { cx A, B, C, D; };
__device__ __forceinline__ m22 m22_unit() {
  m22 m; m.A = cx_make(1, 0); m.B = cx_make(0, 0); m.C = cx_make(0, 0); m.D = cx_make(1, 0); return m;
}
__device__ __forceinline__ m22 m22_zero() {
  m22 m; m.A = m.B = m.C = m.D = cx_make(0, 0); return m;
}
__device__ __forceinline__ m22 m22_mul(const m22 &k, const m22 &x) {
  m22 r;
  r.A = cx_add(cx_mul(k.A, x.A), cx_mul(k.B, x.C));
  r.B = cx_add(cx_mul(k.A, x.B), cx_mul(k.B, x.D));
  r.C = cx_add(cx_mul(k.C, x.A), cx_mul(k.D, x.C));
  r.D = cx_add(cx_mul(k.C, x.B), cx_mul(k.D, x.D));
  return r;
}
// K * [[1, Z], [0, 1]] and K * [[1, 0], [Y, 1]] (series impedance / shunt admittance as the reference builds them:
// Matrix2x2::unit() with one element replaced, then the full product).  In the full product the unit entries
// contribute x * 1 - y * 0 and x * 0 + y * 0, i.e. x and (signed) zero exactly, so leaving them out gives the same
// bits up to the sign of an exact zero.
__device__ __forceinline__ void m22_series(m22 &k, cx Z) {
  k.B = cx_add(cx_mul(k.A, Z), k.B);
  k.D = cx_add(cx_mul(k.C, Z), k.D);
}
__device__ __forceinline__ void m22_shunt(m22 &k, cx Y) {
  k.A = cx_add(k.A, cx_mul(k.B, Y));
  k.C = cx_add(k.C, cx_mul(k.D, Y));
}
__device__ __forceinline__ m22 m22_inv(const m22 &m) {  // det == 1 assumed, as in the reference
  m22 r; r.A = m.D; r.D = m.A; r.B = cx_neg(m.B); r.C = cx_neg(m.C); return r;
}
