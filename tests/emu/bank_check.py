#!/usr/bin/env python
"""Shared-memory bank-conflict model for the compile-time plans of csrc/hs_smem_fft.cuh.

A 16-byte LDS/STS of a warp is served per quarter-warp (8 lanes x 16 B = 128 B = all 32 banks);
lanes of a quarter-warp that hit different addresses in the same 16-byte bank group serialise.
For every stage this prints wavefronts / ideal wavefronts (1.00 = conflict free) for the reads and
the writes, using exactly the index arithmetic of hs_stage().  Used to pick LOGPAD per plan.
"""
import sys


def pad(p, logpad):
    return p + (p >> logpad) if logpad else p


def wavefronts(addrs):
    """addrs: list of 32 (or fewer) element addresses (None = inactive lane)."""
    total = 0
    ideal = 0
    for q in range(0, len(addrs), 8):
        lanes = [a for a in addrs[q:q + 8] if a is not None]
        if not lanes:
            continue
        groups = {}
        for a in set(lanes):
            groups.setdefault(a % 8, set()).add(a)
        total += max(len(v) for v in groups.values())
        ideal += 1
    return total, ideal


def check(N, T, logpad, radices, verbose=True):
    m = 1
    nst = len(radices)
    worst = 1.0
    for s, R in enumerate(radices):
        QP = N // (m * R)
        NB = N // R
        B = -(-NB // T)
        rd = [0, 0]
        wr = [0, 0]
        for w0 in range(0, T, 32):
            lanes = list(range(w0, min(w0 + 32, T)))
            for u in range(B):
                if s > 0:
                    for i in range(R):
                        ad = []
                        for t in lanes:
                            b = t + T * u
                            if b < NB:
                                q, k = b % QP, b // QP
                                ad.append(pad(q + QP * i + QP * R * k, logpad))
                            else:
                                ad.append(None)
                        a, b_ = wavefronts(ad)
                        rd[0] += a
                        rd[1] += b_
                if s < nst - 1:
                    for j in range(R):
                        ad = []
                        for t in lanes:
                            b = t + T * u
                            ad.append(pad(b + NB * j, logpad) if b < NB else None)
                        a, b_ = wavefronts(ad)
                        wr[0] += a
                        wr[1] += b_
        r = rd[0] / rd[1] if rd[1] else 1.0
        w = wr[0] / wr[1] if wr[1] else 1.0
        worst = max(worst, r, w)
        if verbose:
            print("  stage %d radix %2d m=%5d: read x%.2f  write x%.2f" % (s, R, m, r, w))
        m *= R
    return worst


def check_inplace(N, T, logpad, radices, verbose=True):
    """In-place DIF/DIT stages of csrc/hs_conv_fft.cuh: butterfly beta of stage s touches base + m*i,
    base = (beta // m) * L + beta % m; reads and writes hit the same addresses."""
    L = N
    worst = 1.0
    for s, R in enumerate(radices):
        m = L // R
        NB = N // R
        B = -(-NB // T)
        tot = [0, 0]
        for w0 in range(0, T, 32):
            lanes = list(range(w0, min(w0 + 32, T)))
            for u in range(B):
                for i in range(R):
                    ad = []
                    for t in lanes:
                        beta = t + T * u
                        ad.append(pad((beta // m) * L + beta % m + m * i, logpad) if beta < NB else None)
                    a, b_ = wavefronts(ad)
                    tot[0] += a
                    tot[1] += b_
        r = tot[0] / tot[1] if tot[1] else 1.0
        worst = max(worst, r)
        if verbose:
            print("  stage %d radix %2d L=%5d m=%5d: x%.2f" % (s, R, L, m, r))
        L = m
    return worst


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "inplace":
        N, T = int(sys.argv[2]), int(sys.argv[3])
        rad = [int(x) for x in sys.argv[4:]]
        for lp in (0, 2, 3, 4, 5, 6):
            print("N=%d T=%d radices=%s LOGPAD=%d: worst x%.2f" % (N, T, rad, lp, check_inplace(N, T, lp, rad, verbose=False)))
        sys.exit(0)
    plans = [
        (4096, 256, [16, 16, 16]), (4096, 512, [8, 8, 8, 8]), (3000, 200, [3, 5, 5, 5, 8]),
        (1024, 64, [16, 8, 8]), (1024, 64, [8, 8, 16]), (1024, 128, [8, 16, 8]), (2048, 128, [8, 16, 16]), (2048, 128, [16, 16, 8]),
        (512, 64, [8, 8, 8]), (256, 16, [16, 16]), (8192, 512, [8, 8, 8, 16]), (1000, 125, [5, 5, 5, 8]),
    ]
    if len(sys.argv) > 1:
        N, T = int(sys.argv[1]), int(sys.argv[2])
        plans = [(N, T, [int(x) for x in sys.argv[3:]])]
    for N, T, rad in plans:
        best = None
        for lp in (0, 2, 3, 4, 5, 6):
            w = check(N, T, lp, rad, verbose=False)
            if best is None or w < best[1] - 1e-9:
                best = (lp, w)
        print("N=%d T=%d radices=%s -> best LOGPAD=%d (worst x%.2f)" % (N, T, rad, best[0], best[1]))
        check(N, T, best[0], rad)
