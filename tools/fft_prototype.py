"""numpy prototype of the index algebra used by csrc/poisson_fft.cu (development aid;
run it to re-check the Stockham stage maps and the real<->complex packing)."""
import numpy as np


def plan(n, R=16):
    rad = []
    m = n
    while m > 1:
        r = min(R, m)
        # keep radices in {2,4,8,16}
        while m % r:
            r //= 2
        rad.append(r)
        m //= r
    return rad


def stockham(x, radices, sign=-1):
    """out-of-place model of the per-thread register algorithm: every 'thread' t holds the
    R values at positions t + p*T; each stage does R/r radix-r butterflies."""
    n = len(x)
    X = np.array(x, dtype=np.complex128)
    Ns = 1
    for r in radices:
        Y = np.empty_like(X)
        nb = n // r
        for j in range(nb):
            k = j % Ns
            v = np.array([X[j + q * nb] for q in range(r)])
            ang = sign * 2 * np.pi * k / (Ns * r)
            v = v * np.exp(1j * ang * np.arange(r))
            # radix-r DFT
            out = np.array([np.sum(v * np.exp(sign * 2j * np.pi * np.arange(r) * q / r)) for q in range(r)])
            j0 = (j // Ns) * Ns * r + k
            for q in range(r):
                Y[j0 + q * Ns] = out[q]
        X = Y
        Ns *= r
    return X


def r2c_via_half(x):
    """real length 2N -> N+1 bins using one complex FFT of length N."""
    N = len(x) // 2
    z = x[0::2] + 1j * x[1::2]
    Z = np.fft.fft(z)
    k = np.arange(N + 1)
    Zk = Z[k % N]
    Zc = np.conj(Z[(N - k) % N])
    w = np.exp(-2j * np.pi * k / (2 * N))
    return 0.5 * (Zk + Zc) - 0.5j * w * (Zk - Zc)


def c2r_via_half(X):
    """N+1 Hermitian bins -> real length 2N (unnormalised inverse: equals 2N * irfft)."""
    N = len(X) - 1
    k = np.arange(N)
    Xk = X[k]
    Xc = np.conj(X[N - k])
    w = np.exp(2j * np.pi * k / (2 * N))
    Z = (Xk + Xc) + 1j * w * (Xk - Xc)
    z = np.fft.ifft(Z) * N  # unnormalised inverse complex FFT of length N
    x = np.empty(2 * N)
    x[0::2] = z.real
    x[1::2] = z.imag
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (16, 32, 64, 128, 256, 512, 1024, 2048):
        x = rng.normal(size=n) + 1j * rng.normal(size=n)
        p = plan(n)
        assert np.prod(p) == n
        err = np.abs(stockham(x, p) - np.fft.fft(x)).max()
        erri = np.abs(stockham(x, p, +1) - np.fft.ifft(x) * n).max()
        print(n, p, err, erri)
    x = rng.normal(size=64)
    print("r2c", np.abs(r2c_via_half(x) - np.fft.rfft(x)).max())
    X = np.fft.rfft(x)
    print("c2r", np.abs(c2r_via_half(X) - x * 64).max())
