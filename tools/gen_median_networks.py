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
import os
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

    def cemax(self, a, b, lo):
        """max(a, b) given lo = min(a, b): a + b - lo on the 32-bit patterns (exact, wraps mod 2^32).
        Integer adds can issue on the FMA pipe (IMAD.IADD) while FMNMX occupies the half-rate ALU pipe."""
        d = self.nreg
        self.nreg += 1
        self.ops.append((d, 'cemax', a, b, lo))
        return d

    def dce(self):
        # a cemax whose min has no other user becomes a plain max (and the min dies)
        users = {}
        for op in self.ops:
            for src in op[2:]:
                users[src] = users.get(src, 0) + 1
        for o in self.outs:
            users[o] = users.get(o, 0) + 1
        ops = []
        for op in self.ops:
            if op[1] == 'cemax' and users.get(op[4], 0) == 1:
                ops.append((op[0], 'max', op[2], op[3]))
            else:
                ops.append(op)
        live = set(self.outs)
        keep = []
        for op in reversed(ops):
            if op[0] in live:
                keep.append(op)
                live.update(op[2:])
        self.ops = keep[::-1]
        return self

    def run(self, x):
        """x: (n_in, N) array -> (len(outs), N)."""
        regs = {i: x[i] for i in range(self.n_in)}
        for op in self.ops:
            d, kind, a, b = op[:4]
            if kind == 'min':
                regs[d] = np.minimum(regs[a], regs[b])
            elif kind == 'max':
                regs[d] = np.maximum(regs[a], regs[b])
            else:   # cemax: integer identity on the bit patterns
                ia, ib, il = (regs[r].view(np.uint32) for r in (a, b, op[4]))
                regs[d] = (ia + ib - il).view(np.float32)
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


INT_EVERY = 0      # every INT_EVERY-th compare-exchange gets its max from integer adds (0 = never)
_ce_count = [0]


def sort_wires(P: Prog, wires):
    """Apply a sorting network to the SSA values in ``wires``; returns sorted wires."""
    w = list(wires)
    for (i, j) in batcher_pairs(len(w)):
        lo = P.mn(w[i], w[j])
        _ce_count[0] += 1
        if INT_EVERY and _ce_count[0] % INT_EVERY == 0:
            hi = P.cemax(w[i], w[j], lo)
        else:
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


def oe_merge_pairs(n):
    """Comparators of Batcher's odd-even merge of two sorted halves (n a power of two, wires 0..n-1)."""
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

    merge(0, n - 1, 1)
    return pairs


def merge_lists(P: Prog, A, B, cache):
    """Full merge of two sorted wire lists of any lengths (virtual +inf padding); memoised."""
    key = (tuple(A), tuple(B))
    if key in cache:
        return cache[key]
    if not A or not B:
        return list(A) + list(B)
    n = 1
    while n < max(len(A), len(B)):
        n *= 2
    INF = None
    w = list(A) + [INF] * (n - len(A)) + list(B) + [INF] * (n - len(B))
    for (i, j) in oe_merge_pairs(2 * n):
        a, b = w[i], w[j]
        if b is INF:
            continue
        if a is INF:
            w[i], w[j] = b, INF
            continue
        w[i], w[j] = P.mn(a, b), P.mx(a, b)
    out = [v for v in w if v is not INF]
    cache[key] = out
    return out


def shared_extras(P: Prog, X, G, cache):
    """sorted X[j+1 .. j+G-2] for the output pairs j = 0, 2, .. (and, for odd G, sorted X[G-1 .. 2G-3] for the
    last, unpaired output under key G-1), built from nested suffix sorts of the left extras X[:G-1] and nested
    prefix sorts of the right extras X[G-1:] (values are added two at a time, each pair is sorted once)."""
    L, R = X[:G - 1], X[G - 1:]
    suf_memo, pre_memo = {0: []}, {0: []}

    def suf(k):                                   # sorted last k values of L
        if k not in suf_memo:
            step = 2 if k >= 2 else 1
            head = sort_wires(P, L[G - 1 - k:G - 1 - k + step])
            suf_memo[k] = merge_lists(P, head, suf(k - step), cache)
        return suf_memo[k]

    def pre(k):                                   # sorted first k values of R
        if k not in pre_memo:
            step = 2 if k >= 2 else 1
            tail = sort_wires(P, R[k - step:k])
            pre_memo[k] = merge_lists(P, pre(k - step), tail, cache)
        return pre_memo[k]

    out = {}
    for j in range(0, G - 1, 2):
        out[j] = merge_lists(P, suf(G - 2 - j), pre(j), cache)
    if G % 2 == 1:
        out[G - 1] = pre(G - 1)
    return out


def step_layout(K, G):
    """Positions (relative to the first input of the double step) used by gen_step.
    K = 4G - 1 + 2E: the core of a group is the block-aligned middle 3G of the K - G + 1 values common to its G
    windows, E more values on each side are treated as extras."""
    assert (K - (4 * G - 1)) % 2 == 0 and K >= 4 * G - 1
    E = (K - (4 * G - 1)) // 2
    c0 = G - 1 + E                                  # first position of block C0
    blocks = [list(range(c0 + b * G, c0 + (b + 1) * G)) for b in range(4)]
    extras = []
    for base in (0, G):                             # the two groups of the step
        left = list(range(base, base + G - 1 + E))                       # x[base .. base+G-2+E]
        right = list(range(base + 4 * G - 1 + E, base + K + G - 1))      # x[base+4G-1+E .. base+K+G-2]
        extras.append((left, right))
    raw = sorted(set(extras[0][0] + extras[0][1] + extras[1][0] + extras[1][1] + blocks[2] + blocks[3]))
    return E, c0, blocks, extras, raw


def step_raw_index(K, G):
    """window-relative input positions a stateful step reads raw (see gen_step)"""
    return step_layout(K, G)[4]


def gen_step(K, G):
    """Stateful double step, K = 4G - 1 + 2E: 2G consecutive outputs per call, walking along a line.

    With x[i] the inputs of the 2G windows (x[j .. j+K-1] for output j), the blocks C0..C3 (G values each,
    starting at x[G-1+E]) tile the two cores (outputs 0..G-1: C0 u C1 u C2, outputs G..2G-1: C1 u C2 u C3).
    C0 and C1 arrive SORTED from the previous step (its C2, C3); C2, C3 are sorted here and handed on; C1 u C2
    is merged once and serves both cores.  The G-1+2E values of a window outside its core are its extras:
    neighbouring windows share all but one of them, and nested suffix / prefix sorts are shared by the pairs.
    Program inputs: ca[G], cb[G] (sorted), then the raw values x[i], i in step_raw_index(K, G).
    Program outputs: o[2G], then sorted C2[G], C3[G]."""
    E, c0, blocks, extras, raw_idx = step_layout(K, G)
    h = K // 2
    P = Prog(2 * G + len(raw_idx))
    ca, cb = list(range(0, G)), list(range(G, 2 * G))
    xw = {i: 2 * G + n for n, i in enumerate(raw_idx)}
    cache = {}
    cc = sort_wires(P, [xw[i] for i in blocks[2]])
    cd = sort_wires(P, [xw[i] for i in blocks[3]])
    pair = merge_lists(P, cb, cc, cache)
    outs = []
    for gi, core in enumerate((merge_lists(P, ca, pair, cache), merge_lists(P, pair, cd, cache))):
        left, right = extras[gi]
        nx = G - 1 + 2 * E                          # extras per window
        # answer = rank h of core (3G sorted) u extras (nx): candidates from the core are its ranks h-nx .. h
        m = core[h - nx:h + 1]
        Lw = [xw[i] for i in left]
        Rw = [xw[i] for i in right]
        # window j (j = 0..G-1) owns left[j:] and right[:E+j]
        suf_memo, pre_memo = {0: []}, {0: []}

        def suf(k):                                 # sorted last k values of the left extras
            if k not in suf_memo:
                step = 2 if k >= 2 else 1
                head = sort_wires(P, Lw[len(Lw) - k:len(Lw) - k + step])
                suf_memo[k] = merge_lists(P, head, suf(k - step), cache)
            return suf_memo[k]

        def pre(k):                                 # sorted first k values of the right extras
            if k not in pre_memo:
                step = 2 if k >= 2 else 1
                tail = sort_wires(P, Rw[k - step:k])
                pre_memo[k] = merge_lists(P, pre(k - step), tail, cache)
            return pre_memo[k]

        o = [None] * G
        for j in range(0, G - 1, 2):
            # windows j and j+1 share left[j+1:] and right[:E+j]; private: left[j] (window j), right[E+j] (window j+1)
            sh = merge_lists(P, suf(len(Lw) - j - 1), pre(E + j), cache)      # nx - 1 values
            u = merge_select(P, m, sh, nx - 1)
            w = merge_select(P, m, sh, nx)
            o[j] = P.mx(u, P.mn(Lw[j], w))
            o[j + 1] = P.mx(u, P.mn(Rw[E + j], w))
        if G % 2 == 1:                              # last output of an odd group
            j = G - 1
            e = merge_lists(P, suf(len(Lw) - j), pre(E + j), cache)           # all nx extras
            o[j] = merge_select(P, m, e, nx)
        outs += o
    P.outs = outs + cc + cd
    return P.dce()


def verify_step(P: Prog, K, G, trials=600, steps=4, seed=0):
    rng = np.random.default_rng(seed + 977 * K)
    E, c0, blocks, extras, raw_idx = step_layout(K, G)
    L = K - 1 + 2 * G * steps
    for data in (rng.standard_normal((L, trials)).astype(np.float32),
                 rng.integers(0, 3, size=(L, trials)).astype(np.float32),
                 np.sort(rng.standard_normal((L, trials)).astype(np.float32), axis=0)):
        ca = np.sort(data[blocks[0]], axis=0)
        cb = np.sort(data[blocks[1]], axis=0)
        for st in range(steps):
            base = 2 * G * st
            x = data[base:base + K + 2 * G - 1]
            res = P.run(np.concatenate([ca, cb, x[raw_idx]], axis=0))
            for j in range(2 * G):
                if not np.array_equal(res[j], np.sort(data[base + j:base + j + K], axis=0)[K // 2]):
                    return False
            ca, cb = res[2 * G:3 * G], res[3 * G:4 * G]
    return True


def gen_sort(n):
    P = Prog(n)
    P.outs = sort_wires(P, list(range(n)))
    return P.dce()


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

    lines += emit_ops(P, ref)
    for j, o in enumerate(P.outs):
        lines.append(f"  o[{j}] = {ref(o)};")
    lines.append("}")
    return "\n".join(lines)


def emit_ops(P: Prog, ref):
    lines = []
    for op in P.ops:
        d, kind, a, b = op[:4]
        if kind == 'cemax':
            lines.append(f"  const float t{d} = __int_as_float(__float_as_int({ref(a)}) + __float_as_int({ref(b)}) - "
                         f"__float_as_int({ref(op[4])}));")
        else:
            f = 'fminf' if kind == 'min' else 'fmaxf'
            lines.append(f"  const float t{d} = {f}({ref(a)}, {ref(b)});")
    return lines


def emit_step_cuda(P: Prog, K, G):
    """median_step_kK(ca, cb, xr, o, na, nb): see gen_step."""
    nr = len(step_raw_index(K, G))
    lines = [f"// K={K} stateful double step: {len(P.ops)} min/max ops = {len(P.ops) / (2 * G):.1f} per output",
             f"__device__ __forceinline__ void median_step_k{K}(const float (&ca)[{G}], const float (&cb)[{G}], "
             f"const float (&xr)[{nr}], float (&o)[{2 * G}], float (&na)[{G}], float (&nb)[{G}]) {{"]

    def ref(r):
        if r < G:
            return f"ca[{r}]"
        if r < 2 * G:
            return f"cb[{r - G}]"
        if r < P.n_in:
            return f"xr[{r - 2 * G}]"
        return f"t{r}"

    lines += emit_ops(P, ref)
    for j in range(2 * G):
        lines.append(f"  o[{j}] = {ref(P.outs[j])};")
    for j in range(G):
        lines.append(f"  na[{j}] = {ref(P.outs[2 * G + j])};")
    for j in range(G):
        lines.append(f"  nb[{j}] = {ref(P.outs[3 * G + j])};")
    lines.append("}")
    return "\n".join(lines)


def emit_sort_cuda(P: Prog, n):
    lines = [f"// sort of {n} values: {len(P.ops)} min/max ops",
             f"__device__ __forceinline__ void median_sort{n}(const float (&x)[{n}], float (&y)[{n}]) {{"]

    def ref(r):
        return f"x[{r}]" if r < P.n_in else f"t{r}"

    lines += emit_ops(P, ref)
    for j in range(n):
        lines.append(f"  y[{j}] = {ref(P.outs[j])};")
    lines.append("}")
    return "\n".join(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='sm_hpss_mtl_b200/csrc/median_networks_gen.cuh')
    ap.add_argument('--ks', default=','.join(str(k) for k in range(3, 64, 2)))
    ap.add_argument('--gmax', type=int, default=12)
    ap.add_argument('--report', action='store_true')
    ap.add_argument('--int-every', type=int, default=0,
                    help='every n-th compare-exchange takes its max from integer adds on the bit patterns (0 = off)')
    args = ap.parse_args()
    global INT_EVERY
    INT_EVERY = args.int_every
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
    # stateful double steps (K = 4G - 1) and the block sorts that start a line
    step_maxk = int(os.environ.get('HPSS_STEP_MAXK', '55'))     # measured: beyond 55 the stateless group is faster (registers)
    step_ks = [(K, (K + 1) // 4) for K, G in table if (K + 1) % 4 == 0 and 7 <= K <= step_maxk]
    sorts_done = set()
    for K, G in step_ks:
        P = gen_step(K, G)
        ok = verify_step(P, K, G)
        print(f"K={K:3d}  stateful step G={G:2d}  ops={len(P.ops):5d}  per-output={len(P.ops) / (2 * G):6.1f}  verified={ok}",
              file=sys.stderr)
        if not ok:
            raise SystemExit(f"step verification failed for K={K}")
        if G not in sorts_done:
            chunks.append(emit_sort_cuda(gen_sort(G), G))
            chunks.append("")
            sorts_done.add(G)
        chunks.append(emit_step_cuda(P, K, G))
        chunks.append("")
    chunks.append("// MedianStep<K>: stateful walk along a line (2G outputs per step) where K = 4G - 1 has one")
    chunks.append("template <int K> struct MedianStep { static constexpr bool available = false; static constexpr int G = 1; "
                  "static constexpr int NRAW = 1; };")
    for K, G in step_ks:
        raw = step_raw_index(K, G)
        chunks.append(f"template <> struct MedianStep<{K}> {{\n"
                      f"  static constexpr bool available = true;\n  static constexpr int G = {G};\n"
                      f"  static constexpr int NRAW = {len(raw)};\n"
                      f"  static __device__ __forceinline__ int raw_pos(int n) {{ return n < {G - 1} ? n : (n < {2 * G - 2} ? n + 1 : n + {G + 1}); }}\n"
                      f"  static __device__ __forceinline__ void sort(const float (&x)[{G}], float (&y)[{G}]) {{ median_sort{G}(x, y); }}\n"
                      f"  static __device__ __forceinline__ void run(const float (&ca)[{G}], const float (&cb)[{G}], const float (&xr)[{len(raw)}], "
                      f"float (&o)[{2 * G}], float (&na)[{G}], float (&nb)[{G}]) {{ median_step_k{K}(ca, cb, xr, o, na, nb); }}\n}};")
    chunks.append("")
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
