#!/usr/bin/env python
"""Bank model behind WsLayout (sfem_b200/csrc/sshex8_laplacian_ws.cuh): shared-memory wavefronts of the 27 LDS.64 row
loads and the 9 STS.64 row stores of one z-step of a compute warp (lane = slot * 9 + y row, lanes 27..31 shadow lane 26),
fp64, level 8, for a slot-plane buffer [interior run at k | row y=0 | row y=8 | column x=0 | column x=8] with the
skeleton block of slot s at SK[s]. A 64-bit access is served per half-warp; lanes of a half-warp that hit the same
16-byte bank pair with different addresses serialise. The model reproduced the 2.8 wavefronts per LDS.64 ncu measured
for the fixed offsets (56, 58, 56), and predicted the 60 / 20 of the k-dependent offsets that ncu then confirmed
(load conflicts 19.5 M -> 8.1 M, store conflicts 9.9 M -> 4.0 M at C3).

Usage: python tools/dev/ws_bank_model.py [--search]
"""
import sys

L = 8


def layout(k, SK):
    d = {}
    for y in range(1, 8):
        for x in range(1, 8):
            d[(x, y)] = k + (y - 1) * 7 + (x - 1)
    for x in range(9):
        d[(x, 0)] = SK + x
        d[(x, 8)] = SK + 9 + x
    for y in range(1, 8):
        d[(0, y)] = SK + 18 + (y - 1)
        d[(8, y)] = SK + 25 + (y - 1)
    return d


def wavefronts(addrs):
    tot = 0
    for g in (addrs[:16], addrs[16:]):
        banks = {}
        for a in set(g):
            banks[a % 16] = banks.get(a % 16, 0) + 1
        tot += max(banks.values())
    return tot


def count(PS, ks, SKs):
    lay = [layout(ks[s], SKs[s]) for s in range(3)]
    loads = stores = 0
    for x in range(9):
        for iy in range(3):
            addrs = []
            for lane in range(32):
                cl = min(lane, 26)
                s, y = divmod(cl, 9)
                yy = min(max(y + iy - 1, 0), 8)
                addrs.append(s * PS + lay[s][(x, yy)])
            loads += wavefronts(addrs)
        addrs = []
        for lane in range(32):
            cl = min(lane, 26)
            s, y = divmod(cl, 9)
            addrs.append(s * PS + lay[s][(x, y)])
        stores += wavefronts(addrs)
    return loads, stores


if __name__ == "__main__":
    print("ideal (loads, stores):", (54, 18))
    for ks in ((0, 1, 0), (1, 0, 1)):
        print("k =", ks, "fixed SK (56, 58, 56):", count(96, ks, (56, 58, 56)),
              " SK(s) + DK(s) k with SK = (56, 57, 53), DK = (0, +1, -1):",
              count(96, ks, (56, 57 + ks[1], 53 - ks[2])))
    if "--search" in sys.argv:
        for ks in ((0, 1, 0), (1, 0, 1)):
            best = None
            for a in range(50, 65):
                for b in range(50, 65):
                    for c in range(50, 65):
                        l, s = count(96, ks, (a, b, c))
                        if best is None or l + s < best[0]:
                            best = (l + s, l, s, a, b, c)
            print("k =", ks, "best (total, loads, stores, SK0, SK1, SK2):", best)
