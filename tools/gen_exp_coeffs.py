#!/usr/bin/env python
"""Derive the constants of the kernels' fp64 exp: exp(a) = 2^q * T[j] * (1 + r*Q(r)).

a = (32 q + j) * ln2/32 + r,  |r| <= ln2/64;  T[j] = 2^(j/32);  1 + r*Q(r) = degree-5 relative-error minimax of e^r with the constant term then pinned to 1.
Prints C initialisers for pymbar_b200/csrc/exp_tables.h.
"""
import mpmath as mp

mp.mp.dps = 60
NT = 32
h = mp.log(2) / (2 * NT)
deg = 5  # of P(r) ~ e^r ; c0 is forced to 1 afterwards (T + T*p form)


def solve(nodes):
    # unknowns c0..c_deg, E:  e^{-x_i} P(x_i) - 1 = (-1)^i E   (relative error levelled)
    n = deg + 2
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(deg + 1):
            A[i, j] = mp.exp(-x) * x ** j
        A[i, deg + 1] = -((-1) ** i)
        b[i] = 1
    sol = mp.lu_solve(A, b)
    return [sol[j] for j in range(deg + 1)], sol[deg + 1]


def err(c, r):
    return mp.exp(-r) * sum(cj * r ** j for j, cj in enumerate(c)) - 1


nodes = [h * mp.cos(mp.pi * (deg + 1 - i) / (deg + 1)) for i in range(deg + 2)]
M = 6000
grid = [-h + 2 * h * i / M for i in range(M + 1)]
for it in range(40):
    c, E = solve(nodes)
    vals = [err(c, x) for x in grid]
    ext = [grid[0]]
    for i in range(1, M):
        if (vals[i] - vals[i - 1]) * (vals[i + 1] - vals[i]) <= 0:
            ext.append(grid[i])
    ext.append(grid[-1])
    best = []
    for x in ext:
        v = err(c, x)
        if best and mp.sign(err(c, best[-1])) == mp.sign(v):
            if abs(v) > abs(err(c, best[-1])):
                best[-1] = x
        else:
            best.append(x)
    while len(best) > deg + 2:
        if abs(err(c, best[0])) < abs(err(c, best[-1])):
            best.pop(0)
        else:
            best.pop()
    if len(best) < deg + 2 or max(abs(a - b) for a, b in zip(best, nodes)) < h * 1e-9:
        break
    nodes = best
print(f"// free-c0 minimax: c0 - 1 = {mp.nstr(c[0] - 1, 5)}")
c[0] = mp.mpf(1)
c = c[1:]
vals = [mp.exp(-x) * (1 + sum(cj * x ** (j + 1) for j, cj in enumerate(c))) - 1 for x in grid]
E = max(abs(v) for v in vals)
maxerr = max(abs(v) for v in vals)
print(f"// weighted-minimax relative error of exp reconstruction: {mp.nstr(maxerr, 5)} (levelled {mp.nstr(abs(E), 5)})")
for j, cj in enumerate(c):
    print(f"#define MBAR_EXP_C{j + 1} {mp.nstr(cj, 25)}   // {float(cj).hex()}")
print(f"#define MBAR_EXP_NT {NT}")
print(f"#define MBAR_EXP_SCALE {mp.nstr(NT / mp.log(2), 25)}   // {NT}/ln2")
ln2n = mp.log(2) / NT
print(f"#define MBAR_EXP_LN2N {mp.nstr(ln2n, 25)}   // ln2/{NT}")
hi = mp.mpf(float(ln2n))
print(f"#define MBAR_EXP_LN2N_LO {mp.nstr(ln2n - hi, 25)}   // ln2/{NT} - double(ln2/{NT})")
print("// 2^(j/%d), j = 0..%d" % (NT, NT - 1))
print("#define MBAR_EXP_TABLE_VALUES \\")
for j in range(NT):
    print(f"    {mp.nstr(mp.mpf(2) ** (mp.mpf(j) / NT), 25)}" + (", \\" if j < NT - 1 else ""))
