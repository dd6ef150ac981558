#!/usr/bin/env python3
"""Generate straight-line min/max selection code for the sliding-median kernels.

One thread produces G consecutive outputs of a width-K sliding window from the
K+G-1 inputs x[0..K+G-2] it holds in registers (output j = element of rank K//2
of x[j..j+K-1]).  Instead of G independent median networks the generator uses
the overlap between neighbouring windows:

  * core  = x[G-1..K-1] is common to all G windows.  A pruned Batcher odd-even
    merge sort yields only its G middle order statistics m[0..G-1]
    (ranks h-G+1..h, h = K//2); every element below/above them can never be the
    answer of any of the G windows.
  * window j = core + E_j where E_j = X[j..j+G-2] slides over the 2G-2 "extras"
    X = x[0..G-2] ++ x[K..K+G-2].  The answer is the median of m (sorted, G) and
    E_j (G-1), i.e. rank G-1 of 2G-1 values.
  * outputs are paired: E_j and E_{j+1} share G-2 values s; with u,w = ranks
    G-2, G-1 of (m u s) (two min-of-max merge selections) the two answers are
    clamp(X[j],u,w) and clamp(X[j+G-1],u,w).

Everything is emitted as SSA min/max ops, dead-code eliminated, verified against
numpy on random data with ties, and written as a CUDA header.

Usage:  python tools/gen_median_networks.py [--out PATH] [--ks 3,5,...]
"""
from __future__ import annotations

import argparse
import sys
from dataclasses import dataclass, field

import numpy as np


# ------------------------------------------------------------------ IR
@dataclass
class Prog:
    n_in: int
    ops: list = field(default_factory=list)      # (dst, 'min'|'max', a, b)
    outs: list = field(default_factory=list)
    nreg: int = 0

    def __post_init__(self):
        self.nreg = self.n_in

    def op(self, kind, a, b):
        d = self.nreg
        self.nreg += 1
        self.ops.append((d, kind, a, b))
        return d

    def mn(self, a, b):
        return self.op('min', a, b)

    def mx(self, a, b):
        return self.op('max', a, b)

    def dce(self):
        live = set(self.outs)
        keep = []
        for (d, kind, a, b) in reversed(self.ops):
            if d in live:
                keep.append((d, kind, a, b))
                live.add(a)
                live.add(b)
        self.ops = keep[::-1]
        return self

    def run(self, x):
        """x: (n_in, N) array -> (len(outs), N)."""
        regs = {i: x[i] for i in range(self.n_in)}
        for (d, kind, a, b) in self.ops:
            regs[d] = np.minimum(regs[a], regs[b]) if kind == 'min' else np.maximum(regs[a], regs[b])
        return np.stack([regs[o] for o in self.outs])


# ------------------------------------------------------------------ networks
def batcher_pairs(n):
    """Comparator list (i<j) of Batcher's odd-even merge sort on n wires."""
    if n <= 1:
        return []
    p2 = 1
    while p2 < n:
        p2 *= 2
    pairs = []

    def merge(lo, hi, r):
        step = r * 2
        if step < hi - lo:
            merge(lo, hi, step)
            merge(lo + r, hi, step)
            for i in range(lo + r, hi - r, step):
                pairs.append((i, i + r))
        else:
            pairs.append((lo, lo + r))

    def sort(lo, hi):
        if hi - lo >= 1:
            mid = lo + (hi - lo) // 2
            sort(lo, mid)
            sort(mid + 1, hi)
            merge(lo, hi, 1)

    sort(0, p2 - 1)
    return [(i, j) for (i, j) in pairs if j < n]


def sort_wires(P: Prog, wires):
    """Apply a sorting network to the SSA values in ``wires``; returns sorted wires."""
    w = list(wires)
    for (i, j) in batcher_pairs(len(w)):
        lo = P.mn(w[i], w[j])
        hi = P.mx(w[i], w[j])
        w[i], w[j] = lo, hi
    return w


def merge_select(P: Prog, A, B, r):
    """Rank-r (0-based) element of the union of sorted wire lists A and B:
    min over a+b=r+1 of max(A[a-1], B[b-1])  (missing operand = -inf)."""
    terms = []
    for a in range(0, len(A) + 1):
        b = r + 1 - a
        if b < 0 or b > len(B):
            continue
        if a == 0:
            terms.append(B[b - 1])
        elif b == 0:
            terms.append(A[a - 1])
        else:
            terms.append(P.mx(A[a - 1], B[b - 1]))
    assert terms
    v = terms[0]
    for t in terms[1:]:
        v = P.mn(v, t)
    return v


def gen_group(K, G, pair=True):
    """Program with K+G-1 inputs and G outputs (see module docstring)."""
    h = K // 2
    assert 1 <= G <= h + 1 and h <= K - G
    P = Prog(K + G - 1)
    x = list(range(K + G - 1))
    core = x[G - 1:K]
    X = x[0:G - 1] + x[K:K + G - 1]
    cs = sort_wires(P, core)
    m = cs[h - G + 1:h + 1]
    outs = [None] * G
    j = 0
    while j < G:
        if pair and j + 1 < G and G >= 2:
            s = sort_wires(P, X[j + 1:j + G - 1])                    # G-2 shared extras
            u = merge_select(P, m, s, G - 2)
            w = merge_select(P, m, s, G - 1)
            outs[j] = P.mx(u, P.mn(X[j], w))
            outs[j + 1] = P.mx(u, P.mn(X[j + G - 1], w))
            j += 2
        else:
            e = sort_wires(P, X[j:j + G - 1])
            outs[j] = merge_select(P, m, e, G - 1)
            j += 1
    P.outs = outs
    return P.dce()


def verify(P: Prog, K, G, trials=4000, seed=0):
    rng = np.random.default_rng(seed + 131 * K + G)
    n = K + G - 1
    xs = [rng.standard_normal((n, trials)).astype(np.float32),
          rng.integers(0, 4, size=(n, trials)).astype(np.float32),          # heavy ties
          np.sort(rng.standard_normal((n, trials)).astype(np.float32), axis=0),
          -np.sort(rng.standard_normal((n, trials)).astype(np.float32), axis=0)]
    for x in xs:
        got = P.run(x)
        for j in range(G):
            want = np.sort(x[j:j + K], axis=0)[K // 2]
            if not np.array_equal(got[j], want):
                return False
    return True


def best_group(K, gmax=12):
    best = None
    for G in range(1, min(gmax, K // 2 + 1, K - K // 2) + 1):
        for pair in (True, False):
            P = gen_group(K, G, pair)
            cost = len(P.ops) / G
            if best is None or cost < best[0]:
                best = (cost, G, pair, P)
    return best


# ------------------------------------------------------------------ emit
def emit_cuda(P: Prog, K, G, name):
    """Straight-line device function; SSA values become local floats (ptxas allocates)."""
    lines = [f"// K={K} G={G}: {len(P.ops)} min/max ops = {len(P.ops) / G:.1f} per output",
             f"__device__ __forceinline__ void {name}(const float (&x)[{K + G - 1}], float (&o)[{G}]) {{"]

    def ref(r):
        return f"x[{r}]" if r < P.n_in else f"t{r}"

    for (d, kind, a, b) in P.ops:
        f = 'fminf' if kind == 'min' else 'fmaxf'
        lines.append(f"  const float t{d} = {f}({ref(a)}, {ref(b)});")
    for j, o in enumerate(P.outs):
        lines.append(f"  o[{j}] = {ref(o)};")
    lines.append("}")
    return "\n".join(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='sm_hpss_mtl_b200/csrc/median_networks_gen.cuh')
    ap.add_argument('--ks', default=','.join(str(k) for k in range(3, 64, 2)))
    ap.add_argument('--gmax', type=int, default=12)
    ap.add_argument('--report', action='store_true')
    args = ap.parse_args()
    ks = [int(s) for s in args.ks.split(',')]
    chunks = ["// GENERATED by tools/gen_median_networks.py -- do not edit.",
              "// Sliding-median selection networks: G outputs of a width-K window from K+G-1 registers.",
              "#pragma once", ""]
    table = []
    for K in ks:
        cost, G, pair, P = best_group(K, args.gmax)
        ok = verify(P, K, G)
        print(f"K={K:3d}  G={G:2d} pair={int(pair)}  ops={len(P.ops):5d}  per-output={cost:6.1f}  verified={ok}",
              file=sys.stderr)
        if not ok:
            raise SystemExit(f"verification failed for K={K}")
        chunks.append(emit_cuda(P, K, G, f"median_group_k{K}"))
        chunks.append("")
        table.append((K, G))
    chunks.append("template <int K> struct MedianGroup;   // G = outputs per group, run() = generated network")
    for K, G in table:
        chunks.append(f"template <> struct MedianGroup<{K}> {{ static constexpr int G = {G}; "
                      f"static __device__ __forceinline__ void run(const float (&x)[{K + G - 1}], float (&o)[{G}]) "
                      f"{{ median_group_k{K}(x, o); }} }};")
    chunks.append("")
    chunks.append("#define HPSS_MEDIAN_FAST_KS(X) " + " ".join(f"X({K})" for K, _ in table))
    chunks.append("")
    if not args.report:
        with open(args.out, 'w') as f:
            f.write("\n".join(chunks))
        print(f"wrote {args.out}", file=sys.stderr)


if __name__ == '__main__':
    main()
