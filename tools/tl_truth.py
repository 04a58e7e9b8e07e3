"""Extended-precision (x87 80-bit, 64-bit mantissa) evaluation of the reference's TL flow-source
transfer function (TlModel.cpp:484-676, 783-932, 1283-1400) for the DEFAULT options, used to attribute
the difference between the bit-identical kernel and the fast-arithmetic kernel: which of the two is
nearer to the value the formulas define?  Test-side analysis tool (uses the oracle to expand frames).

    python tools/tl_truth.py gpurun_out/tl_fast_worst.npz
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

LD, CLD = np.longdouble, np.clongdouble
RHO, CS, MU = LD("1.14e-3"), LD("3.5e4"), LD("1.86e-4")
PI = LD("3.14159265358979323846264338327950288")
MIN_AREA = LD("0.1e-2")
SINUS_AFTER = {65 + 8: 89, 65 + 9: 90, 65 + 11: 91, 65 + 12: 92}
SINUS_VOL = [LD("11.3"), LD("6.8"), LD("33.0"), LD("6.2")]


def truth_flow_tf(sections, f0_hz, bins, section=25, fossa=False):
    """sections: oracle expand_frame output (doubles, exact inputs); bins: integer array."""
    v = sections["v"].astype(LD)
    area = np.maximum(v[:, 1], MIN_AREA); length = v[:, 2]; volume = v[:, 3]
    Mw, Bw, Kw = v[:, 4], v[:, 5], v[:, 6]
    fi = np.where(bins == 0, 1, bins).astype(LD)
    w = 2 * PI * LD(f0_hz) * fi
    sw = np.sqrt(w)
    rs = np.sqrt(RHO * MU / 2); cpf = 1 / (RHO * CS * CS)
    one, zero = np.ones_like(w, dtype=CLD), np.zeros_like(w, dtype=CLD)

    def mul(K, M):
        return (K[0] * M[0] + K[1] * M[2], K[0] * M[1] + K[1] * M[3], K[2] * M[0] + K[3] * M[2], K[2] * M[1] + K[3] * M[3])

    def za_zb(k):
        circ = 2 * np.sqrt(area[k] * PI)
        Za = (rs * (length[k] / 2) * circ * sw / (area[k] * area[k])) + 1j * (w * RHO * (length[k] / 2) / area[k])
        Yb = 1j * (w * cpf * length[k] * area[k]) + 1 / ((Bw[k] + 1j * (w * Mw[k] - Kw[k] / w)) / (length[k] * circ))
        return Za.astype(CLD), (1 / Yb).astype(CLD)

    def smat(k):
        if k >= 89:
            circ = 2 * np.sqrt(area[k] * PI)
            Za = (rs * length[k] * circ * sw / (area[k] * area[k])) + 1j * (w * RHO * length[k] / area[k])
            inner = 4 * PI * ((3 * SINUS_VOL[k - 89]) / (4 * PI)) ** (LD(2) / 3)
            Yb = inner / 65000 + 1j * (w * cpf * volume[k])
            Zb = 1 / Yb
            return (1 + Za / Zb, -Zb, 1 / Zb, -one)
        Za, Zb = za_zb(k)
        q = 1 + Za / Zb
        return (q, Za * Za / Zb + 2 * Za, 1 / Zb, q)

    def rad(a):
        a = max(a, MIN_AREA)
        Z1 = (128 * RHO * CS) / (9 * PI * PI * a)
        Z2 = 1j * (w * 8 * RHO / (3 * PI * np.sqrt(PI * a)))
        return (Z1 * Z2) / (Z1 + Z2)

    P = {}
    unit = (one, zero, zero, one)
    fossa_zin = None
    if fossa:  # TlModel.cpp:554-561
        K = unit
        for k in range(84, 89):
            K = mul(K, smat(k))
        fossa_zin = K[0] / K[2]
    for first, last in ((0, 40), (41, 64), (65, 83)):
        K = unit
        for k in range(first, last + 1):
            K = mul(K, smat(k))
            P[k] = K
            if fossa and k == 25 + 3:  # Tube::FOSSA_COUPLING_SECTION (TlModel.cpp:586-590)
                K = mul(K, (one, zero, 1 / fossa_zin, one))
            if k in SINUS_AFTER:
                M = smat(SINUS_AFTER[k])
                K = mul(K, (one, zero, 1 / (M[0] / M[2]), one))
        # mean flow is 0 with lung pressure 0: the static pressure-drop resistors are 0

    def inv(M):
        return (M[3], -M[1], -M[2], M[0])

    def through(M, load):
        return (M[0] * load + M[1]) / (M[2] * load + M[3])

    mouthRad, noseRad = rad(v[64, 1]), rad(v[83, 1])
    zn, zm = through(P[83], noseRad), through(P[64], mouthRad)
    par = lambda a, b: a * b / (a + b)
    assert 0 < section < 40
    zin = through(mul(inv(P[section]), P[40]), par(zn, zm))
    zout = P[section - 1][1] / P[section - 1][0]
    Za, Zb = za_zb(section)
    zin = zin + Za; zout = zout + Za
    factor = par(Zb, zout)
    factor = factor / (zin + factor)
    M = mul(inv(P[section]), P[40])
    ph = mul((one, Za, zero, one), M)
    K = mul(mul(ph, (one, zero, 1 / zn, one)), P[64])
    res = factor / (K[2] * mouthRad + K[3])
    K = mul(mul(ph, (one, zero, 1 / zm, one)), P[83])
    res = res + factor / (K[2] * noseRad + K[3])
    return res


if __name__ == "__main__":
    from oracle.pyoracle import FRAME_DTYPE, Oracle
    d = np.load(sys.argv[1])
    orc = Oracle()
    frames = d["frames"].view(FRAME_DTYPE).reshape(-1)
    f0 = float(d["f0"])
    e_ref, e_fast = [], []
    section = int(d["section"]) if "section" in d else 25
    fossa = bool(d["fossa"]) if "fossa" in d else False
    from oracle.pyoracle import tl_options_default
    oo = tl_options_default(); oo.piriform_fossa = int(fossa)
    for i, (fr, b) in enumerate(zip(frames, d["bins"])):
        bins = np.array([b])
        t = truth_flow_tf(orc.expand_frame(fr.reshape(1)), f0, bins, section=section, fossa=fossa)[0]
        ro, io = orc.tl_spectrum(fr.reshape(1), f0, int(b) + 1, opts=oo, section=section)
        ref = CLD(ro[0, b]) + 1j * CLD(io[0, b])
        fast = CLD(d["fast_re"][i]) + 1j * CLD(d["fast_im"][i])
        gref = CLD(d["ref_re"][i]) + 1j * CLD(d["ref_im"][i])
        assert gref == ref, "device reference arithmetic is not the oracle's"
        e_ref.append(float(abs(ref - t) / abs(t))); e_fast.append(float(abs(fast - t) / abs(t)))
        print(f"tube {int(d['tubes'][i]):5d} bin {int(b):4d}: |ref-truth|/|truth| {e_ref[-1]:.3e}   |fast-truth|/|truth| {e_fast[-1]:.3e}"
              f"   |fast-ref|/|ref| {float(abs(fast - ref) / abs(ref)):.3e}")
    print(f"max over the {len(e_ref)} worst points: reference arithmetic {max(e_ref):.3e}, fast arithmetic {max(e_fast):.3e}")
