"""CPU checks of the 512-point real FFT the way gsn_fft.cu computes it (no GPU needed): the twiddle table compiled into
the library, and a numpy restatement of the kernels' decomposition -- 256 complex points (even / odd samples) as 16 x 16
with 4 x 4 inside, the register positions `pos16`, the transpose, the split / merge step of the real transform -- against
numpy's FFT.  The GPU tests (tests/test_gpu_parity.py) compare the kernels themselves with torch.stft / irfft."""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TABLE = os.path.join(os.path.dirname(HERE), "spiking_fullsubnet_b200", "csrc", "gsn_fft_tables.cuh")


def _table():
    txt = open(TABLE).read()
    vals = re.findall(r"\{(-?[0-9.e+-]+)f, (-?[0-9.e+-]+)f\}", txt)
    return np.array([complex(float(a), float(b)) for a, b in vals])


def test_twiddle_table_is_exp_minus_2pi_i_m_over_512():
    tw = _table()
    assert tw.shape == (512,)
    want = np.exp(-2j * np.pi * np.arange(512) / 512)
    assert np.abs(tw - want).max() < 6e-8  # float64 values rounded to float32
    assert tw[0] == 1 and abs(tw[128] - (-1j)) < 1e-12 and abs(tw[256] + 1) < 1e-12


def pos16(k):
    return 4 * (k & 3) + (k >> 2)


def dft4(a, inv):
    t0, t1, t2, d = a[0] + a[2], a[0] - a[2], a[1] + a[3], a[1] - a[3]
    t3 = d * (1j if inv else -1j)
    return [t0 + t2, t1 + t3, t0 - t2, t1 - t3]


def dft16(v, inv):
    """x[n] at v[n] -> X[k] at v[pos16(k)] (gsn_fft.cu: dft16)."""
    v = list(v)
    for b in range(4):
        v[b], v[4 + b], v[8 + b], v[12 + b] = dft4([v[b], v[4 + b], v[8 + b], v[12 + b]], inv)
    for c in range(1, 4):
        for b in range(1, 4):
            w = np.exp(-2j * np.pi * (b * c) / 16)
            v[4 * c + b] *= np.conj(w) if inv else w
    for c in range(4):
        v[4 * c:4 * c + 4] = dft4(v[4 * c:4 * c + 4], inv)
    return v


def fft256(x, inv):
    """Thread j holds x[16 n1 + j] at v[n1]; returns X with X[j + 16 k2] taken from thread j's v[pos16(k2)]."""
    tw = _table()
    s = np.zeros((16, 17), dtype=complex)  # the padded transpose buffer
    for j in range(16):
        v = dft16([x[16 * n1 + j] for n1 in range(16)], inv)
        for k1 in range(16):
            w = tw[(2 * j * k1) & 511]  # exp(-2 pi i j k1 / 256)
            s[k1, j] = v[pos16(k1)] * ((np.conj(w) if inv else w) if k1 else 1.0)
    out = np.zeros(256, dtype=complex)
    for j in range(16):
        v = dft16([s[j, n2] for n2 in range(16)], inv)
        for k2 in range(16):
            out[j + 16 * k2] = v[pos16(k2)]
    return out


def test_dft16_positions():
    rs = np.random.RandomState(0)
    x = rs.standard_normal(16) + 1j * rs.standard_normal(16)
    v = dft16(x, False)
    assert np.allclose([v[pos16(k)] for k in range(16)], np.fft.fft(x))
    v = dft16(x, True)
    assert np.allclose([v[pos16(k)] for k in range(16)], np.fft.ifft(x) * 16)


def test_forward_real_fft_as_the_kernel_computes_it():
    rs = np.random.RandomState(1)
    x = rs.standard_normal(512)
    tw = _table()
    z = fft256(x[0::2] + 1j * x[1::2], False)  # even samples real, odd imaginary
    X = np.zeros(257, dtype=complex)
    for k in range(256):
        zk, zm = z[k], z[(256 - k) & 255]
        e = 0.5 * (zk + np.conj(zm))
        o = -0.5j * (zk - np.conj(zm))
        X[k] = e + tw[k] * o
        if k == 0:
            X[256] = (e - o).real
    assert np.abs(X - np.fft.rfft(x)).max() < 1e-5 * np.abs(X).max()  # (float32 table values)


def test_inverse_real_fft_as_the_kernel_computes_it():
    rs = np.random.RandomState(2)
    X = rs.standard_normal(257) + 1j * rs.standard_normal(257)  # imaginary parts of DC / Nyquist must be ignored
    tw = _table()
    Xc = X.copy()
    Xc[0] = Xc[0].real
    Xc[256] = Xc[256].real
    Z = np.zeros(256, dtype=complex)
    for k in range(256):
        xk, xm = Xc[k], Xc[256 - k]
        e = 0.5 * (xk + np.conj(xm))
        o = 0.5 * (xk - np.conj(xm)) * np.conj(tw[k])
        Z[k] = e + 1j * o
    z = fft256(Z, True) / 256
    x = np.empty(512)
    x[0::2], x[1::2] = z.real, z.imag
    want = np.fft.irfft(X, n=512)
    assert np.abs(x - want).max() < 1e-5 * np.abs(want).max()
