"""Design-time numpy model of the block FFT used by the CUDA kernels: multi-pass Stockham
(read t + b*T + q*N/R, twiddle W_{NS*R}^{q*k}, DFT_R, write expand(j)+q*NS) and the r2c
post-pass.  Run to validate the index algebra before touching CUDA."""
import numpy as np

def expand(j, ns, r):
    return (j // ns) * ns * r + (j % ns)

def block_fft(x, radices):
    N = len(x); cur = np.array(x, dtype=np.complex128); ns = 1
    for R in radices:
        out = np.zeros(N, dtype=np.complex128)
        for j in range(N // R):
            k = j % ns
            v = np.array([cur[j + q * (N // R)] for q in range(R)])
            v = v * np.exp(-2j * np.pi * np.arange(R) * k / (ns * R))
            V = np.fft.fft(v)
            o = expand(j, ns, R)
            for q in range(R):
                out[o + q * ns] = V[q]
        cur = out; ns *= R
    return cur

def r2c(x, radices):
    """real x of length 2N -> full 2N-point spectrum via N-point complex FFT; returns 2*X (0.5 folded out)."""
    N = len(x) // 2
    z = x[0::2] + 1j * x[1::2]
    Z = block_fft(z, radices)
    X = np.zeros(2 * N, dtype=np.complex128)
    for k in range(0, N // 2 + 1):
        A = Z[k]; B = np.conj(Z[(N - k) % N])
        E2 = A + B; O2 = A - B
        th = np.pi * k / N
        Wm = complex(-np.sin(th), -np.cos(th))   # -i * exp(-i th)
        T = Wm * O2
        X[k] = E2 + T
        X[N + k] = E2 - T
        if k > 0:
            X[N - k] = np.conj(E2 - T)
            X[2 * N - k] = np.conj(E2 + T)
    return X

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N, rad in [(512, [8, 8, 8]), (1024, [16, 8, 8]), (2048, [16, 16, 8]), (64, [4, 4, 4]), (32, [8, 4]), (128, [2, 8, 8])]:
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        err = np.abs(block_fft(x, rad) - np.fft.fft(x)).max()
        xr = rng.standard_normal(2 * N)
        err2 = np.abs(r2c(xr, rad) - 2 * np.fft.fft(xr)).max()
        print(N, rad, "c2c err", err, "r2c err", err2)
