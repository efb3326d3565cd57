"""Design-time tool: search padded shared-memory layouts pad(i) = i + (i >> A) * C for the
Stockham exchange after each FFT pass so that the strided writes are bank-conflict-free for
64-bit accesses (16 lanes x 8 B per wavefront).  Reads are t + const (always conflict-free
when A >= 4).  Prints, per exchange, the best (A, C) and the write wavefronts per request."""
import itertools, sys

def expand(j, ns, r):
    return (j // ns) * ns * r + (j % ns)

def cost(N, T, P, R, NS, A, C):
    B = P // R
    tot = 0; cnt = 0
    for b in range(B):
        for q in range(R):
            for h in range(max(T // 16, 1)):
                banks = {}
                for lane in range(min(16, T)):
                    t = 16 * h + lane
                    i = expand(t + b * T, NS, R) + q * NS
                    a = i + (i >> A) * C if A is not None else i
                    banks.setdefault(a % 16, set()).add(a)
                tot += max(len(s) for s in banks.values()); cnt += 1
    return tot / cnt

def search(N, T, radices):
    P = N // T
    ns = 1
    for p, R in enumerate(radices[:-1]):
        best = None
        for A in [None, 3, 4, 5, 6, 7, 8, 9]:
            for C in ([0] if A is None else [1, 2, 4, 8]):
                c = cost(N, T, P, R, ns, A, C)
                size = N if A is None else N + ((N - 1) >> A) * C + 1
                key = (round(c, 3), size)
                if best is None or key < best[0]:
                    best = (key, A, C)
        print(f"N={N} T={T} pass{p} R={R} NS={ns}: best wavefronts/req={best[0][0]} size={best[0][1]} A={best[1]} C={best[2]}  (nopad={cost(N,T,P,R,ns,None,0):.2f})")
        ns *= R

if __name__ == "__main__":
    search(512, 64, [8, 8, 8])
    search(1024, 64, [16, 8, 8])
    search(1024, 128, [8, 8, 16])
    search(1024, 64, [4, 16, 16])
    search(2048, 128, [16, 16, 8])
    search(2048, 128, [8, 16, 16])
    search(4096, 256, [16, 16, 16])
    search(8192, 512, [16, 16, 16, 2])
    search(8192, 512, [8, 8, 8, 16])
    search(256, 32, [8, 8, 4])
    search(256, 16, [16, 16])
    search(128, 16, [8, 8, 2])
    search(64, 16, [4, 4, 4])
