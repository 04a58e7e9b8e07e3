// tl_algebra.cuh -- transfer-function algebra on the stored chain products: TlModel::getInputImpedance,
// getOutputImpedance, getPressureSourceTF, getFlowSourceTF (reference TlModel.cpp:987-1469).  Shared by the two
// TL translation units; the including unit defines cx_div / cx_rdiv BEFORE the include (libgcc Smith division in
// the bit-identical unit, conj/|z|^2 in the fast one).  Included inside an anonymous namespace.
#pragma once

// The products the transfer-function algebra may ask for.
struct Prods {
  m22 sm1, s, p40, p64, p83;
  int section;
  cx mouthRad, noseRad;
};
__device__ __forceinline__ m22 getP(const Prods &P, int k) {
  if (k == PHAR_LAST) return P.p40;
  if (k == MOUTH_LAST) return P.p64;
  if (k == NOSE_LAST) return P.p83;
  if (k == P.section) return P.s;
  if (k == P.section - 1) return P.sm1;
  return m22_unit();  // matrixProduct[][] is initialised to unit matrices (TlModel.cpp:52-58)
}
__device__ __forceinline__ cx load_through(const m22 &M, cx load) {
  return cx_div(cx_add(cx_mul(M.A, load), M.B), cx_add(cx_mul(M.C, load), M.D));
}

// TlModel::getInputImpedance (TlModel.cpp:987-1050)
__device__ cx input_impedance(const Prods &P, int section) {
  m22 M = m22_unit();
  cx load = cx_make(0, 0);
  if (section >= TRACHEA0 && section <= PHAR_LAST) {
    const cx zn = load_through(P.p83, P.noseRad);   // getInputImpedance(FIRST_NOSE_SECTION)
    const cx zm = load_through(P.p64, P.mouthRad);  // getInputImpedance(FIRST_MOUTH_SECTION)
    load = cx_div(cx_mul(zn, zm), cx_add(zn, zm));
    if (section > TRACHEA0) M = m22_inv(getP(P, section - 1));
    M = m22_mul(M, P.p40);
  } else if (section >= MOUTH0 && section <= MOUTH_LAST) {
    load = P.mouthRad;
    if (section > MOUTH0) M = m22_inv(getP(P, section - 1));
    M = m22_mul(M, P.p64);
  } else if (section >= NOSE0 && section <= NOSE_LAST) {
    load = P.noseRad;
    if (section > NOSE0) M = m22_inv(getP(P, section - 1));
    M = m22_mul(M, P.p83);
  } else {
    M = m22_zero();
  }
  return load_through(M, load);
}
// TlModel::getOutputImpedance (TlModel.cpp:1060-1105), lungTerminationImpedance == 0 (:518)
__device__ cx output_impedance(const Prods &P, int section) {
  m22 M = m22_zero(), K;
  const cx lung = cx_make(0, 0);
  if (section >= TRACHEA0 && section <= PHAR_LAST) {
    M = getP(P, section);
  } else if (section >= MOUTH0 && section <= MOUTH_LAST) {
    M = P.p40;
    K = m22_unit(); K.C = cx_rdiv(1.0, input_impedance(P, NOSE0));
    M = m22_mul(M, K); M = m22_mul(M, getP(P, section));
  } else if (section >= NOSE0 && section <= NOSE_LAST) {
    M = P.p40;
    K = m22_unit(); K.C = cx_rdiv(1.0, input_impedance(P, MOUTH0));
    M = m22_mul(M, K); M = m22_mul(M, getP(P, section));
  }
  return cx_div(cx_add(cx_mul(lung, M.D), M.B), cx_add(M.A, cx_mul(lung, M.C)));
}

__device__ __forceinline__ cx par(cx a, cx b) { return cx_div(cx_mul(a, b), cx_add(a, b)); }
__device__ __forceinline__ cx rad_term(cx factor, const m22 &K, cx zrad) {
  return cx_div(factor, cx_add(cx_mul(K.C, zrad), K.D));
}

// TlModel::getPressureSourceTF (TlModel.cpp:1118-1272)
__device__ cx pressure_source_tf(const Prods &P, int section) {
  m22 M, K;
  cx result = cx_make(0, 0), factor, zin, zout;
  if (section >= TRACHEA0 && section <= PHAR_LAST) {
    zin = input_impedance(P, section);
    if (section == TRACHEA0) zout = cx_make(0, 0); else zout = output_impedance(P, section - 1);
    factor = cx_rdiv(1.0, cx_add(zout, zin));
    M = m22_unit();
    if (section > TRACHEA0) M = getP(P, section - 1);
    M = m22_inv(M);
    M = m22_mul(M, P.p40);
    const m22 pharynx = M;
    K = pharynx; M = m22_unit(); M.C = cx_rdiv(1.0, input_impedance(P, NOSE0));
    K = m22_mul(K, M); K = m22_mul(K, P.p64);
    result = cx_add(result, rad_term(factor, K, P.mouthRad));
    K = pharynx; M = m22_unit(); M.C = cx_rdiv(1.0, input_impedance(P, MOUTH0));
    K = m22_mul(K, M); K = m22_mul(K, P.p83);
    result = cx_add(result, rad_term(factor, K, P.noseRad));
  } else if ((section >= MOUTH0 && section <= MOUTH_LAST) || section == -1) {
    if (section == -1) zin = P.mouthRad; else zin = input_impedance(P, section);
    if (section == -1) zout = output_impedance(P, MOUTH_LAST);
    else if (section == MOUTH0) zout = par(output_impedance(P, PHAR_LAST), input_impedance(P, NOSE0));
    else zout = output_impedance(P, section - 1);
    factor = cx_rdiv(1.0, cx_add(zout, zin));
    if (section == -1) M = m22_unit();
    else if (section > MOUTH0) { M = m22_inv(getP(P, section - 1)); M = m22_mul(M, P.p64); }
    else M = P.p64;
    result = cx_add(result, rad_term(factor, M, P.mouthRad));
    M = m22_unit();
    if (section == -1) M = P.p64;
    else if (section > MOUTH0) M = getP(P, section - 1);
    M = m22_inv(M);
    K = m22_unit(); K.C = cx_rdiv(1.0, output_impedance(P, PHAR_LAST));
    M = m22_mul(M, K); M = m22_mul(M, P.p83);
    result = cx_add(result, rad_term(factor, M, P.noseRad));
  }
  return result;
}

// TlModel::getFlowSourceTF (TlModel.cpp:1283-1469)
// Za, Zb: lumped impedances of `section` (getLumpedSectionImpedances, :1330/:1406), supplied by the caller when
// the section lies in the trachea..mouth range.
__device__ cx flow_source_tf(const Prods &P, int section, cx Za, cx Zb) {
  m22 M, K;
  cx result = cx_make(0, 0), factor, zin, zout;
  if (section == NOSE0) {
    K = P.p83;
    return cx_rdiv(1.0, cx_add(cx_mul(K.C, P.noseRad), K.D));
  }
  if (section >= TRACHEA0 && section <= PHAR_LAST) {
    if (section == PHAR_LAST) zin = par(input_impedance(P, MOUTH0), input_impedance(P, NOSE0));
    else zin = input_impedance(P, section + 1);
    if (section == TRACHEA0) zout = cx_make(0, 0); else zout = output_impedance(P, section - 1);
    zin = cx_add(zin, Za); zout = cx_add(zout, Za);
    factor = cx_div(cx_mul(Zb, zout), cx_add(Zb, zout));
    factor = cx_div(factor, cx_add(zin, factor));
    M = m22_inv(P.s);
    M = m22_mul(M, P.p40);
    m22 pharynx = m22_unit(); pharynx.B = Za; pharynx = m22_mul(pharynx, M);
    K = pharynx; M = m22_unit(); M.C = cx_rdiv(1.0, input_impedance(P, NOSE0));
    K = m22_mul(K, M); K = m22_mul(K, P.p64);
    result = cx_add(result, rad_term(factor, K, P.mouthRad));
    K = pharynx; M = m22_unit(); M.C = cx_rdiv(1.0, input_impedance(P, MOUTH0));
    K = m22_mul(K, M); K = m22_mul(K, P.p83);
    result = cx_add(result, rad_term(factor, K, P.noseRad));
  } else if (section >= MOUTH0 && section <= MOUTH_LAST) {
    if (section == MOUTH_LAST) zin = P.mouthRad; else zin = input_impedance(P, section + 1);
    if (section == MOUTH0) zout = par(output_impedance(P, PHAR_LAST), input_impedance(P, NOSE0));
    else zout = output_impedance(P, section - 1);
    zin = cx_add(zin, Za); zout = cx_add(zout, Za);
    factor = cx_div(cx_mul(Zb, zout), cx_add(Zb, zout));
    factor = cx_div(factor, cx_add(zin, factor));
    M = m22_inv(P.s);
    M = m22_mul(M, P.p64);
    K = m22_unit(); K.B = Za; K = m22_mul(K, M);
    result = cx_add(result, rad_term(factor, K, P.mouthRad));
    M = m22_inv(P.s);
    K = m22_unit(); K.C = cx_rdiv(1.0, output_impedance(P, PHAR_LAST));
    M = m22_mul(M, K); M = m22_mul(M, P.p83);
    K = m22_unit(); K.B = Za; K = m22_mul(K, M);
    result = cx_add(result, rad_term(factor, K, P.noseRad));
  }
  return result;
}

