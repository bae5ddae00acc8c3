"""
numpy prototype of the device algorithm (Stockham passes + packed real FFT +
pointwise filter + packed inverse).  Development aid: pins the index maps and
scalings that csrc/of_kernel.cuh transcribes.  Not imported by the product.
"""
import numpy as np

def stockham(z, radices, sign):
    """Govindaraju-style Stockham autosort; natural in, natural out. sign=-1 fwd."""
    M = z.shape[-1]
    a = z.astype(np.complex128).copy()
    Ns = 1
    for R in radices:
        b = np.empty_like(a)
        nb = M // R
        j = np.arange(nb)
        q = j % Ns
        v = np.stack([a[j + r * nb] * np.exp(sign * 2j * np.pi * q * r / (Ns * R)) for r in range(R)])  # [R, nb]
        # radix-R DFT
        W = np.exp(sign * 2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)
        o = W @ v
        idx = (j // Ns) * Ns * R + q
        for r in range(R):
            b[idx + r * Ns] = o[r]
        a = b
        Ns *= R
    return a

def fwd_real_packed(x, radices):
    """returns X[0..M] (M=N/2) = rfft(x) via M-point complex FFT."""
    N = x.shape[-1]; M = N // 2
    z = x[0::2] + 1j * x[1::2]
    Z = stockham(z, radices, -1)
    k = np.arange(M + 1)
    Zk = Z[k % M]; Zc = np.conj(Z[(M - k) % M])
    E = 0.5 * (Zk + Zc)
    O = -0.5j * (Zk - Zc)
    return E + np.exp(-2j * np.pi * k / N) * O

def inv_real_packed(F, radices):
    """F[0..M] hermitian half spectrum -> y[n] = sum_k F_k e^{+2 pi i k n/N} over all N bins (unnormalised irfft*N)."""
    M = F.shape[-1] - 1; N = 2 * M
    k = np.arange(M)
    Fk = F[k]; Fc = np.conj(F[M - k])
    E = (Fk + Fc)            # sum over even-sample spectrum (unnormalised)
    O = (Fk - Fc) * np.exp(2j * np.pi * k / N)
    Zp = E + 1j * O
    z = stockham(Zp, radices, +1)
    y = np.empty(N)
    y[0::2] = z.real; y[1::2] = z.imag
    return y

if __name__ == '__main__':
    rng = np.random.default_rng(0)
    for N, rad in [(64, [4, 8]), (1024, [16, 32]), (32768, [32, 32, 16]), (16384, [16, 16, 32])]:
        x = rng.standard_normal(N)
        X = fwd_real_packed(x, rad)
        ref = np.fft.rfft(x)
        e1 = np.abs(X - ref).max() / np.abs(ref).max()
        F = ref * rng.standard_normal(ref.shape)  # arbitrary hermitian-compatible (DC/Nyq must be real)
        F[0] = F[0].real; F[-1] = F[-1].real
        y = inv_real_packed(F, rad)
        yref = np.fft.irfft(F, n=N) * N
        e2 = np.abs(y - yref).max() / np.abs(yref).max()
        print(N, rad, e1, e2)
