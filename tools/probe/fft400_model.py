"""Index-math model of the warp-private FFT-400 (nww_fe5.cuh): 20 x 20 decomposition, radix-20 as a 4 x 5
prime-factor DFT, (k, 400 - k) pairing across lanes.  Checks against numpy.fft on random frames."""
import numpy as np

def dft20_pfa(v):
    # input map n = (5 n1 + 4 n2) % 20 ; output map k = (5 k1 + 16 k2) % 20
    t = np.zeros((4, 5), complex)
    for n2 in range(5):
        x = [v[(5 * n1 + 4 * n2) % 20] for n1 in range(4)]
        for k1 in range(4):
            t[k1, n2] = sum(x[n1] * np.exp(-2j * np.pi * n1 * k1 / 4) for n1 in range(4))
    out = np.zeros(20, complex)
    for k1 in range(4):
        for k2 in range(5):
            out[(5 * k1 + 16 * k2) % 20] = sum(t[k1, n2] * np.exp(-2j * np.pi * n2 * k2 / 5) for n2 in range(5))
    return out

rng = np.random.default_rng(0)
a, b = rng.standard_normal(400), rng.standard_normal(400)
z = a + 1j * b
assert np.allclose(dft20_pfa(z[:20]), np.fft.fft(z[:20]))
# pass 1: lane j
y = np.zeros((20, 20), complex)          # y[q][j]
for j in range(20):
    V = dft20_pfa(z[j::20])
    for q in range(20):
        y[q, j] = V[q] * np.exp(-2j * np.pi * j * q / 400)
# pass 2: lane q holds Z[q + 20 p] in register p
R = np.zeros((20, 20), complex)
for q in range(20):
    R[q] = dft20_pfa(y[q])
Z = np.fft.fft(z)
for q in range(20):
    for p in range(20):
        assert np.allclose(R[q, p], Z[q + 20 * p])
# pairing: lane q, p = 0..9 (lane 0 also bin 200)
PA, PB = np.zeros(201), np.zeros(201)
for q in range(20):
    for p in range(10):
        U = R[q, p]
        W = R[(20 - q) % 20, 19 - p] if q else R[0, (20 - p) % 20]
        ar, ai = U.real + W.real, U.imag - W.imag
        br, bi = U.imag + W.imag, U.real - W.real
        PA[q + 20 * p] = 0.25 * (ar * ar + ai * ai)
        PB[q + 20 * p] = 0.25 * (br * br + bi * bi)
U = R[0, 10]
PA[200], PB[200] = U.real ** 2, U.imag ** 2
assert np.allclose(PA, np.abs(np.fft.fft(a))[:201] ** 2) and np.allclose(PB, np.abs(np.fft.fft(b))[:201] ** 2)
print("fft400 model ok")
