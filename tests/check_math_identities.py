"""Array-level numpy emulation of the CUDA formulation (time-domain analysis rows, complex-signal
FFT, two-term contraction, packed inverse FFT) against the oracle.  Design check, not a test."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sh_chain as oc

def emulate(g, symmetric=True):
    B = int(g["B"]); N2 = 2 * B; order = int(g["order"]); K = (order + 1) ** 2
    Yw = g["Yw"].astype(np.complex128); H = g["Hnm"][0].astype(np.complex128)
    sh_m = g["sh_m"]; rev = list(g["rev"])
    sig = []  # (row_re, row_im, k1, m1, k2, m2, sign2)
    rows = []
    if symmetric:
        for n in range(order + 1):
            for m in range(n + 1):
                kpos, kneg = n * n + n + m, n * n + n - m
                rows.append(Yw[kpos].real); rr = len(rows) - 1; ri = -1
                if m > 0:
                    rows.append(Yw[kpos].imag); ri = len(rows) - 1
                sig.append((rr, ri, kneg, -m, kpos if m > 0 else -1, m, -1.0 if m & 1 else 1.0))
    else:
        for k in range(K):
            rows.append(Yw[rev[k]].real); rows.append(Yw[rev[k]].imag)
            sig.append((2 * k, 2 * k + 1, k, int(sh_m[k]), -1, 0, 1.0))
    Rm = np.array(rows)
    T = g["x"].shape[0]
    zprev = np.zeros((Rm.shape[0], B)); rad_last = None
    w_in, w_out = oc.crossfade_windows(B)
    ys = []
    src16 = np.array([(float(g["source_azim"]), 0)], dtype=np.float16)
    for t in range(T):
        zcur = Rm @ g["x"][t].astype(np.float64)
        az = oc.individual_azimuth_deg(g["yaw"][t], src16, bool(g["is_hrir"]))
        rad = float(np.deg2rad(az)[0])
        acc = np.zeros((2, 2, B + 1), dtype=np.complex128)  # [cur/last][ear][f]
        f = np.arange(B + 1)
        for (rr, ri, k1, m1, k2, m2, sg) in sig:
            z = np.concatenate([zprev[rr], zcur[rr]]).astype(np.complex128)
            if ri >= 0:
                z = z + 1j * np.concatenate([zprev[ri], zcur[ri]])
            Z = np.fft.fft(z)
            Zf = Z[f]; Zc = np.conj(Z[(N2 - f) % N2])
            for which, r in ((0, rad), (1, rad_last)):
                if r is None:
                    continue
                acc[which] += H[k1] * np.exp(-1j * m1 * r) * Zf
                if k2 >= 0:
                    acc[which] += sg * H[k2] * np.exp(-1j * m2 * r) * Zc
        out = []
        for which in (0, 1):
            eL, eR = acc[which]
            eL = eL.copy(); eR = eR.copy()
            eL[0] = eL[0].real; eR[0] = eR[0].real; eL[B] = eL[B].real; eR[B] = eR[B].real
            Zo = np.zeros(N2, dtype=np.complex128)
            Zo[: B + 1] = eL + 1j * eR
            Zo[B + 1:] = (np.conj(eL) + 1j * np.conj(eR))[1:B][::-1]
            y = np.fft.ifft(Zo)
            out.append(np.stack([y.real[B:], y.imag[B:]]))
        ys.append(out[0] * w_in + out[1] * w_out)
        zprev = zcur; rad_last = rad
    return np.stack(ys)

for name, sym in (("sh_em32_n4_b256", True), ("sh_le110_n8_b128", True), ("sh_rnd50_n3_b32", True),
                  ("sh_em32_n2_b64_general", False), ("sh_rnd434_n12_b64", True)):
    g = np.load(f"tests/golden/{name}.npz")
    y = emulate(g, sym)
    print(name, "max err vs reference", np.abs(y - g["y"]).max(), "peak", np.abs(g["y"]).max())
