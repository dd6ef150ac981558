#!/usr/bin/env python3
"""Generate straight-line register DFT codelets for the specialised STFT kernel (csrc/stft_fast.cuh).

A codelet `dftN(float2 (&v)[N])` transforms v in place (natural order in, natural order out,
forward transform exp(-2 pi i jk / N)).  Sizes are built by Cooley-Tukey recursion down to hand-written
radix-2/3/4/5 butterflies; every twiddle is a literal constant, quarter turns cost nothing.  The op list is
evaluated with numpy against numpy.fft before anything is written.  Coprime splits (10 = 2 x 5, 20 = 4 x 5) use the
Good-Thomas prime-factor map, which needs no twiddles between the stages (DFT-10: 140 -> 92 operations, DFT-20:
344 -> 224).

Usage:  python tools/gen_fft_codelets.py [--out PATH] [--sizes 10,20,16,32]
"""
from __future__ import annotations

import argparse
import math
import sys

import numpy as np


class Prog:
    def __init__(self, n_in):
        self.lines = []          # (name, expr) with expr over names / const names
        self.consts = {}         # name -> float value
        self.n = 0
        self.n_in = n_in

    def const(self, val):
        for k, v in self.consts.items():
            if v == val:
                return k
        name = f"c{len(self.consts)}"
        self.consts[name] = val
        return name

    def tmp(self, expr):
        name = f"t{self.n}"
        self.n += 1
        self.lines.append((name, expr))
        return name

    # real ops (ZERO is a symbolic exact zero: real-input codelets feed it as every imaginary part and the
    # simplifications below remove the arithmetic it touches)
    def add(self, a, b):
        if a == ZERO: return b
        if b == ZERO: return a
        return self.tmp(f"{a} + {b}")

    def sub(self, a, b):
        if b == ZERO: return a
        if a == ZERO: return self.neg(b)
        return self.tmp(f"{a} - {b}")

    def neg(self, a):
        return ZERO if a == ZERO else self.tmp(f"-{a}")

    def mulc(self, a, c):
        return ZERO if a == ZERO else self.tmp(f"{a} * {self.const(c)}")

    def mad(self, a, c, b):                                # a*c + b
        if a == ZERO: return b
        if b == ZERO: return self.mulc(a, c)
        return self.tmp(f"{a} * {self.const(c)} + {b}")

    def msub(self, a, c, b):                               # b - a*c
        if a == ZERO: return b
        if b == ZERO: return self.mulc(a, -c)
        return self.tmp(f"{b} - {a} * {self.const(c)}")


ZERO = "ZERO"


def cadd(P, a, b): return (P.add(a[0], b[0]), P.add(a[1], b[1]))
def csub(P, a, b): return (P.sub(a[0], b[0]), P.sub(a[1], b[1]))
def mul_mi(P, a): return (a[1], P.neg(a[0]))          # * (-i)
def mul_pi(P, a): return (P.neg(a[1]), a[0])          # * (+i)
def cneg(P, a): return (P.neg(a[0]), P.neg(a[1]))


def ctwiddle(P, a, m, n):
    """a * exp(-2 pi i m / n)"""
    m %= n
    if m == 0:
        return a
    if 4 * m == n:
        return mul_mi(P, a)
    if 2 * m == n:
        return cneg(P, a)
    if 4 * m == 3 * n:
        return mul_pi(P, a)
    ang = -2.0 * math.pi * m / n
    c, s = math.cos(ang), math.sin(ang)
    re = P.msub(a[1], s, P.mulc(a[0], c))             # a.re*c - a.im*s
    im = P.mad(a[1], c, P.mulc(a[0], s))              # a.re*s + a.im*c
    return (re, im)


def dft2(P, v):
    return [cadd(P, v[0], v[1]), csub(P, v[0], v[1])]


def dft3(P, v):
    s = math.sqrt(3.0) / 2
    t1 = cadd(P, v[1], v[2])
    d = csub(P, v[1], v[2])
    t2 = (P.msub(t1[0], 0.5, v[0][0]), P.msub(t1[1], 0.5, v[0][1]))
    o0 = cadd(P, v[0], t1)
    o1 = (P.mad(d[1], s, t2[0]), P.msub(d[0], s, t2[1]))
    o2 = (P.msub(d[1], s, t2[0]), P.mad(d[0], s, t2[1]))
    return [o0, o1, o2]


def dft4(P, v):
    t0 = cadd(P, v[0], v[2])
    t1 = csub(P, v[0], v[2])
    t2 = cadd(P, v[1], v[3])
    t3 = mul_mi(P, csub(P, v[1], v[3]))
    return [cadd(P, t0, t2), cadd(P, t1, t3), csub(P, t0, t2), csub(P, t1, t3)]


def dft5(P, v):
    c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
    s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
    a1, a2 = cadd(P, v[1], v[4]), cadd(P, v[2], v[3])
    b1, b2 = csub(P, v[1], v[4]), csub(P, v[2], v[3])
    o0 = tuple(P.add(P.add(v[0][i], a1[i]), a2[i]) for i in range(2))
    p1 = tuple(P.mad(a2[i], c2, P.mad(a1[i], c1, v[0][i])) for i in range(2))
    p2 = tuple(P.mad(a2[i], c1, P.mad(a1[i], c2, v[0][i])) for i in range(2))
    q1 = tuple(P.mad(b2[i], s2, P.mulc(b1[i], s1)) for i in range(2))
    q2 = tuple(P.msub(b2[i], s1, P.mulc(b1[i], s2)) for i in range(2))
    # X1 = p1 - i q1, X4 = p1 + i q1, X2 = p2 - i q2, X3 = p2 + i q2
    o1 = (P.add(p1[0], q1[1]), P.sub(p1[1], q1[0]))
    o4 = (P.sub(p1[0], q1[1]), P.add(p1[1], q1[0]))
    o2 = (P.add(p2[0], q2[1]), P.sub(p2[1], q2[0]))
    o3 = (P.sub(p2[0], q2[1]), P.add(p2[1], q2[0]))
    return [o0, o1, o2, o3, o4]


BASE = {2: dft2, 3: dft3, 4: dft4, 5: dft5}


def split(n):
    for r in (4, 5, 2, 3):
        if n % r == 0 and n != r:
            return r, n // r
    raise ValueError(f"cannot factor {n} into 2/3/4/5")


def dft_pfa(P, v, n1, n2):
    """Good-Thomas prime-factor step for coprime n1, n2: no twiddles between the two stages.
    Input map n = (n2*j1 + n1*j2) mod N, output map (CRT) k = (k1*n2*(n2^-1 mod n1) + k2*n1*(n1^-1 mod n2)) mod N."""
    n = n1 * n2
    A = [dft(P, [v[(n2 * j1 + n1 * j2) % n] for j1 in range(n1)]) for j2 in range(n2)]      # A[j2][k1]
    e1, e2 = n2 * pow(n2, -1, n1), n1 * pow(n1, -1, n2)
    out = [None] * n
    for k1 in range(n1):
        res = dft(P, [A[j2][k1] for j2 in range(n2)])
        for k2 in range(n2):
            out[(k1 * e1 + k2 * e2) % n] = res[k2]
    return out


def dft(P, v):
    n = len(v)
    if n == 1:
        return list(v)
    if n in BASE:
        return BASE[n](P, v)
    n1, n2 = split(n)            # x[n2*j1 + j2]; X[k1 + n1*k2]
    if math.gcd(n1, n2) == 1:
        return dft_pfa(P, v, n1, n2)
    A = [dft(P, [v[n2 * j1 + j2] for j1 in range(n1)]) for j2 in range(n2)]
    out = [None] * n
    for k1 in range(n1):
        col = [ctwiddle(P, A[j2][k1], j2 * k1, n) for j2 in range(n2)]
        res = dft(P, col)
        for k2 in range(n2):
            out[k1 + n1 * k2] = res[k2]
    return out


def build(n):
    P = Prog(n)
    v = [(f"x{i}r", f"x{i}i") for i in range(n)]
    outs = dft(P, v)
    return P, outs


def build_real(n, n_out):
    """DFT of n REAL inputs, outputs 0 .. n_out-1 only (the rest follows from Hermitian symmetry)."""
    P = Prog(n)
    v = [(f"x{i}", ZERO) for i in range(n)]
    outs = dft(P, v)[:n_out]
    return P, outs


def live_lines(P, outs):
    """dead-code elimination: the lines the requested outputs depend on, in order"""
    import re as _re
    need = set()
    for o in outs:
        need.update(o)
    keep = []
    for name, expr in reversed(P.lines):
        if name in need:
            keep.append((name, expr))
            need.update(_re.findall(r"[tx]\d+[ri]?", expr))
    return list(reversed(keep))


def verify_real(P, outs, n, trials=64):
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal((n, trials))
    env = {f"x{i}": x[i].copy() for i in range(n)}
    env.update(P.consts)
    env[ZERO] = np.zeros(trials)
    for name, expr in live_lines(P, outs):
        env[name] = eval(expr, {}, env)
    got = np.stack([env[o[0]] + 1j * env[o[1]] for o in outs])
    want = np.fft.fft(x, axis=0)[:len(outs)]
    return float(np.abs(got - want).max() / np.abs(want).max())


def emit_real(P, outs, n):
    lines_ = live_lines(P, outs)
    ops = sum(1 for _, e in lines_ if not e.startswith("-"))
    lines = [f"// real-input DFT-{n}, outputs 0..{len(outs) - 1}: {ops} float operations (before FMA contraction)",
             f"__device__ __forceinline__ void rdft{n}(const float (&x)[{n}], float2 (&y)[{len(outs)}]) {{"]
    used = " ".join(e for _, e in lines_)
    for k, val in P.consts.items():
        if k in used.split() or any(k == tok for tok in used.replace("-", " ").split()):
            lines.append(f"  constexpr float {k} = {val:.17g}f;")
    for i in range(n):
        lines.append(f"  const float x{i} = x[{i}];")
    for name, expr in lines_:
        lines.append(f"  const float {name} = {expr};")
    for k, o in enumerate(outs):
        re_, im_ = (("0.f" if c == ZERO else c) for c in o)
        lines.append(f"  y[{k}] = make_float2({re_}, {im_});")
    lines.append("}")
    return "\n".join(lines)


def verify(P, outs, n, trials=64):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, trials)) + 1j * rng.standard_normal((n, trials))
    env = {}
    for i in range(n):
        env[f"x{i}r"] = x[i].real.copy()
        env[f"x{i}i"] = x[i].imag.copy()
    env.update(P.consts)
    for name, expr in P.lines:
        env[name] = eval(expr, {}, env)
    got = np.stack([env[o[0]] + 1j * env[o[1]] for o in outs])
    want = np.fft.fft(x, axis=0)
    return float(np.abs(got - want).max() / np.abs(want).max())


def emit(P, outs, n):
    ops = sum(1 for _, e in P.lines if not e.startswith("-"))
    lines = [f"// DFT-{n}: {ops} float operations (before FMA contraction)",
             f"__device__ __forceinline__ void dft{n}(float2 (&v)[{n}]) {{"]
    for k, val in P.consts.items():
        lines.append(f"  constexpr float {k} = {val:.17g}f;")
    for i in range(n):
        lines.append(f"  const float x{i}r = v[{i}].x, x{i}i = v[{i}].y;")
    for name, expr in P.lines:
        lines.append(f"  const float {name} = {expr};")
    for k, o in enumerate(outs):
        lines.append(f"  v[{k}] = make_float2({o[0]}, {o[1]});")
    lines.append("}")
    return "\n".join(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="sm_hpss_mtl_b200/csrc/fft_codelets_gen.cuh")
    ap.add_argument("--sizes", default="8,10,16,20,32")
    args = ap.parse_args()
    chunks = ["// GENERATED by tools/gen_fft_codelets.py -- do not edit.",
              "// In-place register DFT codelets (forward, natural order).", "#pragma once", ""]
    for n in [int(s) for s in args.sizes.split(",")]:
        P, outs = build(n)
        err = verify(P, outs, n)
        print(f"DFT-{n}: {len(P.lines)} ops, max rel err {err:.2e}", file=sys.stderr)
        if err > 1e-12:
            raise SystemExit(f"verification failed for DFT-{n}")
        chunks.append(emit(P, outs, n))
        chunks.append("")
    for n, n_out in ((20, 11),):
        P, outs = build_real(n, n_out)
        err = verify_real(P, outs, n)
        print(f"real DFT-{n} ({n_out} outputs): {len(live_lines(P, outs))} ops, max rel err {err:.2e}", file=sys.stderr)
        if err > 1e-12:
            raise SystemExit(f"verification failed for real DFT-{n}")
        chunks.append(emit_real(P, outs, n))
        chunks.append("")
    chunks.append("template <int N> struct Dft;")
    for n in [int(s) for s in args.sizes.split(",")]:
        chunks.append(f"template <> struct Dft<{n}> {{ static __device__ __forceinline__ void run(float2 (&v)[{n}]) {{ dft{n}(v); }} }};")
    chunks.append("")
    with open(args.out, "w") as f:
        f.write("\n".join(chunks))
    print(f"wrote {args.out}", file=sys.stderr)


if __name__ == "__main__":
    main()
