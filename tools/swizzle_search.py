"""Brute-force an XOR swizzle for the warp-private shared-memory FFT exchange buffers.

Model: float2 elements; a 64-bit warp access is served per half-warp (16 lanes); it costs as many
wavefronts as the maximum number of distinct addresses that fall into one of the 16 bank pairs.
"""
import itertools

def wavefronts(addrs):  # addrs: 32 float2 indices, one per lane
    total = 0
    for h in range(2):
        lanes = addrs[16 * h: 16 * h + 16]
        banks = {}
        for a in lanes:
            banks.setdefault(a % 16, set()).add(a)
        total += max(len(v) for v in banks.values())
    return total  # ideal = 2

def accesses(N2):
    G = min(32, N2 // 8)
    EPL = N2 // G
    groups = 32 // G
    radices = {64: (8, 8), 128: (8, 8, 2), 256: (8, 8, 4), 512: (8, 8, 8)}[N2]
    acc = []  # list of lane->idx lists (per group-local index)
    # reads: idx = l + G*u
    for u in range(EPL):
        acc.append(("read", [[l + G * u for l in range(G)]] ))
    Ns = 1
    for R in radices[:-1]:
        nb = EPL // R
        for b in range(nb):
            for t in range(R):
                idxs = []
                for l in range(G):
                    j = l + G * b
                    k = j % Ns
                    j0 = (j // Ns) * Ns * R + k
                    idxs.append(j0 + t * Ns)
                acc.append((f"write Ns={Ns}", [idxs]))
        Ns *= R
    return G, groups, acc

def evaluate(N2, swz, region_stride):
    G, groups, acc = accesses(N2)
    worst = 0; tot = 0; n = 0
    for name, (idxs,) in acc:
        addrs = []
        for g in range(groups):
            addrs += [g * region_stride + swz(i) for i in idxs]
        w = wavefronts(addrs)
        worst = max(worst, w); tot += w; n += 1
    return worst, tot / n

def make_swz(terms):
    def f(i):
        x = 0
        for sh, mask, lsh in terms:
            x ^= ((i >> sh) & mask) << lsh
        return i ^ x
    return f

for N2 in (64, 128, 256, 512):
    best = None
    cands = [(sh, mask, lsh) for sh in range(2, 9) for mask in (1, 3, 7, 15) for lsh in range(0, 4)
             if (mask << lsh) < 16 and sh >= (mask.bit_length() + lsh)]
    for n in (0, 1, 2):
        for terms in itertools.combinations(cands, n):
            for pad in (0, 4, 8):
                rs = N2 + pad
                worst, avg = evaluate(N2, make_swz(terms), rs)
                key = (worst, avg, n, pad)
                if best is None or key < best[0]:
                    best = (key, terms, pad)
        if best[0][0] == 2:
            break
    print(N2, best)
