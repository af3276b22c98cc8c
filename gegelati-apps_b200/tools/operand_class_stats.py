#!/usr/bin/env python3
"""Operand-class statistics per (heavy line, 128-sample warp tile) on the C2 workload (numpy emulation of the generic code)."""
import sys, numpy as np, collections
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'python'))
import tpgx
np.seterr(all='ignore')
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 400
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 6
g = tpgx.synthetic_graph("mnist", roots=2000, edges=5, lines=20, depth=1, mix="uniform", seed=42)
fl = tpgx.flatten_programs(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5)
eff = fl['effective_lines']; code = fl['code']
off = np.concatenate([[0], np.cumsum(eff)])
img, lab = tpgx.synthetic_images(n=128 * NT)
X = img.astype(np.float64).reshape(-1, 28, 28)
S = X.shape[0]
consts = np.asarray(g.constants).reshape(-1, 5) if np.asarray(g.constants).ndim == 1 else np.asarray(g.constants)
print('constants', consts.shape, consts.dtype, consts[:2])

def win(loc):
    r, c = divmod(loc, 28)
    a = X[:, r:r + 3, c:c + 3]
    gx = -a[:, 0, 0] + a[:, 0, 2] - 2.0 * a[:, 1, 0] + 2.0 * a[:, 1, 2] - a[:, 2, 0] + a[:, 2, 2]
    gy = -a[:, 0, 0] - 2.0 * a[:, 0, 1] - a[:, 0, 2] + a[:, 2, 0] + 2.0 * a[:, 2, 1] + a[:, 2, 2]
    return gx, gy

def grp(mask):  # per 128-sample group: all
    return mask.reshape(NT, 128)

C = collections.Counter()
def classify(v):
    a = np.abs(v)
    zero = v == 0
    nan = np.isnan(v)
    inf = np.isinf(v)
    sub = (a > 0) & (a < 2.2250738585072014e-308)
    return zero, nan, inf, sub

for p in range(NP):
    words = code[off[p]:off[p + 1]]
    reg = np.zeros((9, S))
    for w in words:
        w = int(w)
        op, dst = w & 31, (w >> 5) & 7
        fa, fb = (w >> 8) & 0xfff, w >> 20
        def fetch(f):
            k, loc = f >> 10, f & 1023
            if k == 0: return reg[min(loc, 8)], 'R'
            if k == 2: return X.reshape(S, -1)[:, loc], 'P'
            raise ValueError(k)
        if op in (20, 21):
            gx, gy = win(fa & 1023)
            res = np.sqrt(gx * gx + gy * gy) if op == 20 else np.arctan(gy / gx)
        elif op == 11:
            a, _ = fetch(fa)
            c = float(consts[p, fb & 1023])
            prod = a * c
            res = prod / 10.0
            z, n, i, s = classify(prod)
            ap = np.abs(prod)
            comfy = (ap >= 2.0 ** -1000) & (ap <= 2.0 ** 1000)
            C['mulc'] += NT
            C['mulc all comfy'] += grp(comfy).all(1).sum()
            C['mulc all comfy|zero'] += grp(comfy | z).all(1).sum()
            C['mulc all comfy|zero|inf|nan'] += grp(comfy | z | n | i).all(1).sum()
        elif op == 5:
            a, _ = fetch(fa)
            res = np.exp(a)
            main = np.abs(a) < 700
            C['exp'] += NT
            C['exp all main'] += grp(main).all(1).sum()
            big = (a >= 710) | (a <= -746) | np.isnan(a)
            C['exp all main|far'] += grp(main | big).all(1).sum()
            C['exp none main'] += (~grp(main)).all(1).sum()
        elif op == 6:
            a, _ = fetch(fa)
            res = np.log(a)
            main = (a >= 2.2250738585072014e-308) & np.isfinite(a)
            C['ln'] += NT
            C['ln all main'] += grp(main).all(1).sum()
            C['ln none main'] += (~grp(main)).all(1).sum()
            C['ln all main|zero'] += grp(main | (a == 0)).all(1).sum()
        else:
            a, ka = fetch(fa); b, kb = fetch(fb)
            if op == 0: res = a - b
            elif op == 1: res = a + b
            elif op == 2: res = a * b
            elif op == 4: res = np.where(a < b, b, a)
            elif op == 3:
                res = a / b
                za, na, ia, sa = classify(a); zb, nb, ib, sb = classify(b)
                aa, ab = np.abs(a), np.abs(b)
                comfa = (aa >= 2.0 ** -500) & (aa <= 2.0 ** 500)
                comfb = (ab >= 2.0 ** -500) & (ab <= 2.0 ** 500)
                clean = comfa & comfb
                C['div'] += NT
                C['div all clean'] += grp(clean).all(1).sum()
                C['div all clean | a==0&comfb'] += grp(clean | (za & comfb)).all(1).sum()
                spec = za | na | ia | zb | nb | ib
                C['div all special'] += grp(spec).all(1).sum()
                C['div all clean|special'] += grp(clean | spec).all(1).sum()
                C['div kinds ' + ka + kb] += NT
            else: raise ValueError(op)
        reg[dst] = res
tot = C['exp']
for k in sorted(C): print('%-36s %8d  %.3f' % (k, C[k], C[k] / C[k.split()[0]]))
