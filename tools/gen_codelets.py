#!/usr/bin/env python3
"""Straight-line DFT codelet generator for the sm_100a log-mel kernels.

The STFT stage is a two-stage Cooley-Tukey split N = N1 x N2 (400 = 20 x 20 for the Whisper
frontend, 1024 = 32 x 32 for the torchaudio one).  Each stage is a small DFT held entirely in
registers by one thread; this script traces those small DFTs symbolically (prime-factor /
Cooley-Tukey with radix 2, 4, 5 leaves), prunes everything a real-valued input or an unneeded
output makes dead, fuses multiply-adds, and prints C++ templates over a value type ``T`` that
is ``float`` on the host / scalar device path and the packed ``f32x2`` type on the fast path.

Nothing here is copied from FFTW's genfft; it is the same idea at toy size.

    python tools/gen_codelets.py > mlx8_ws_audio_transformer_b200/csrc/codelets_gen.cuh
    python tools/gen_codelets.py --selftest
"""
from __future__ import annotations

import cmath
import math
import sys

# ----------------------------------------------------------------------------------------
# symbolic values with hash-consing, sign tracking and eager FMA fusion
# ----------------------------------------------------------------------------------------


class Node:
    __slots__ = ("op", "args", "id", "name")

    def __init__(self, op, args, nid):
        self.op, self.args, self.id, self.name = op, args, nid, None


class Graph:
    def __init__(self):
        self.nodes = {}
        self.order = []

    def node(self, op, *args):
        key = (op,) + tuple(a.id if isinstance(a, Node) else ("c", repr(a)) for a in args)
        n = self.nodes.get(key)
        if n is None:
            n = Node(op, args, len(self.order))
            self.nodes[key] = n
            self.order.append(n)
        return n


class V:
    """sign * node, or exact zero (node is None)."""
    __slots__ = ("g", "n", "s")

    def __init__(self, g, n, s=1):
        self.g, self.n, self.s = g, n, s

    @property
    def zero(self):
        return self.n is None

    def __neg__(self):
        return V(self.g, self.n, -self.s)

    def __add__(self, o):
        g = self.g
        if self.zero:
            return o
        if o.zero:
            return self
        a, b = self, o
        # fuse a multiply into the add: x + a*k -> fma(a, k, x)
        for x, y in ((a, b), (b, a)):
            if y.n.op == "mulc":
                ya, yk = y.n.args
                # x + (ys * ya * yk): normalise so that x carries sign +1
                k = y.s * yk * x.s
                return V(g, g.node("fmac", ya, k, x.n), x.s)
            if y.n.op == "mulv":
                ya, yb = y.n.args
                if y.s * x.s > 0:
                    return V(g, g.node("fmav", ya, yb, x.n), x.s)
                return V(g, g.node("fnmav", ya, yb, x.n), x.s)       # x - ya*yb
        if a.s > 0 and b.s > 0:
            return V(g, g.node("add", *sorted((a.n, b.n), key=lambda n: n.id)), 1)
        if a.s < 0 and b.s < 0:
            return V(g, g.node("add", *sorted((a.n, b.n), key=lambda n: n.id)), -1)
        if a.s > 0:
            return V(g, g.node("sub", a.n, b.n), 1)
        return V(g, g.node("sub", b.n, a.n), 1)

    def __sub__(self, o):
        return self + (-o)

    def mulc(self, k: float):
        """multiply by a compile-time constant."""
        if self.zero or k == 0.0:
            return V(self.g, None)
        if k == 1.0:
            return self
        if k == -1.0:
            return -self
        n, s = self.n, self.s
        if n.op == "mulc":                      # (a*k1)*k2 -> a*(k1*k2)
            k = k * n.args[1]
            n = n.args[0]
        if k < 0:
            k, s = -k, -s
        return V(self.g, self.g.node("mulc", n, k), s)

    def mulv(self, o):
        """multiply by another runtime value (window sample, twiddle)."""
        if self.zero or o.zero:
            return V(self.g, None)
        a, b = sorted((self.n, o.n), key=lambda n: n.id)
        return V(self.g, self.g.node("mulv", a, b), self.s * o.s)


class C:
    """complex value made of two V's."""
    __slots__ = ("re", "im")

    def __init__(self, re, im):
        self.re, self.im = re, im

    def __add__(self, o):
        return C(self.re + o.re, self.im + o.im)

    def __sub__(self, o):
        return C(self.re - o.re, self.im - o.im)

    def __neg__(self):
        return C(-self.re, -self.im)

    def mul_i(self):      # * (+i)
        return C(-self.im, self.re)

    def mul_mi(self):     # * (-i)
        return C(self.im, -self.re)

    def scale(self, k):
        return C(self.re.mulc(k), self.im.mulc(k))

    def mulc(self, w: complex):
        """multiply by a compile-time complex constant, special-casing the cheap ones."""
        wr, wi = _snap(w.real), _snap(w.imag)
        if wi == 0.0:
            return self.scale(wr)
        if wr == 0.0:
            return C((-self.im).mulc(wi), self.re.mulc(wi))
        if abs(abs(wr) - abs(wi)) < 1e-15:
            # (a+ib)(c+id) with |c|=|d|: c*((a -/+ b) + i(b +/- a))
            sr = 1.0 if wr > 0 else -1.0
            si = 1.0 if wi > 0 else -1.0
            m = abs(wr)
            re = self.re.mulc(sr) - self.im.mulc(si)
            im = self.re.mulc(si) + self.im.mulc(sr)
            return C(re.mulc(m), im.mulc(m))
        re = self.re.mulc(wr) - self.im.mulc(wi)
        im = self.re.mulc(wi) + self.im.mulc(wr)
        return C(re, im)

    def mulv(self, wr: V, wi: V):
        """multiply by a runtime complex value (wr + i wi)."""
        re = self.re.mulv(wr) - self.im.mulv(wi)
        im = self.re.mulv(wi) + self.im.mulv(wr)
        return C(re, im)


def _snap(x: float) -> float:
    for t in (0.0, 1.0, -1.0, 0.5, -0.5):
        if abs(x - t) < 1e-15:
            return t
    return x


# ----------------------------------------------------------------------------------------
# small DFTs
# ----------------------------------------------------------------------------------------


def dft2(x):
    return [x[0] + x[1], x[0] - x[1]]


def dft4(x):
    a, b = x[0] + x[2], x[0] - x[2]
    c, d = x[1] + x[3], x[1] - x[3]
    return [a + c, b + d.mul_mi(), a - c, b + d.mul_i()]


def dft5(x):
    c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
    s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
    t1, t2 = x[1] + x[4], x[2] + x[3]
    t3, t4 = x[1] - x[4], x[2] - x[3]
    X0 = x[0] + t1 + t2
    a1 = x[0] + t1.scale(c1) + t2.scale(c2)
    a2 = x[0] + t1.scale(c2) + t2.scale(c1)
    b1 = t3.scale(s1) + t4.scale(s2)
    b2 = t3.scale(s2) - t4.scale(s1)
    return [X0, a1 + b1.mul_mi(), a2 + b2.mul_mi(), a2 + b2.mul_i(), a1 + b1.mul_i()]


def dft3(x):
    c, s = -0.5, math.sin(2 * math.pi / 3)
    t1, t2 = x[1] + x[2], x[1] - x[2]
    a = x[0] + t1.scale(c)
    b = t2.scale(s)
    return [x[0] + t1, a + b.mul_mi(), a + b.mul_i()]


FACTOR = {8: (2, 4), 10: (2, 5), 16: (4, 4), 20: (4, 5), 25: (5, 5), 32: (4, 8), 64: (8, 8),
          40: (8, 5), 50: (2, 25), 100: (4, 25)}


def dft(x):
    """DFT of a list of C; returns outputs in natural order."""
    n = len(x)
    if n == 1:
        return list(x)
    if n == 2:
        return dft2(x)
    if n == 3:
        return dft3(x)
    if n == 4:
        return dft4(x)
    if n == 5:
        return dft5(x)
    n1, n2 = FACTOR[n]
    out = [None] * n
    if math.gcd(n1, n2) == 1:
        # Good-Thomas: n = (n2*a + n1*b) mod n, k = CRT(k1 mod n1, k2 mod n2); no twiddles
        inner = []
        for b in range(n2):
            inner.append(dft([x[(n2 * a + n1 * b) % n] for a in range(n1)]))
        for k1 in range(n1):
            col = dft([inner[b][k1] for b in range(n2)])
            for k2 in range(n2):
                k = next(k for k in range(n) if k % n1 == k1 and k % n2 == k2)
                out[k] = col[k2]
        return out
    # Cooley-Tukey: n = n2*a + b, k = k1 + n1*k2
    inner = []
    for b in range(n2):
        y = dft([x[n2 * a + b] for a in range(n1)])
        inner.append([y[k1].mulc(cmath.exp(-2j * math.pi * b * k1 / n)) for k1 in range(n1)])
    for k1 in range(n1):
        col = dft([inner[b][k1] for b in range(n2)])
        for k2 in range(n2):
            out[k1 + n1 * k2] = col[k2]
    return out


# ----------------------------------------------------------------------------------------
# codelets
# ----------------------------------------------------------------------------------------


class Codelet:
    def __init__(self, name, doc):
        self.g = Graph()
        self.name, self.doc = name, doc
        self.params = []          # (kind, cname, count)  kind in {"in", "sin", "out"}
        self.scalars = set()      # ids of scalar-typed input nodes
        self.outputs = []         # (cexpr, V)
        self.neg_twin = {}        # input array -> array holding its negation (packed constants)

    def inputs(self, cname, count, scalar=False):
        """declare an input array; ``scalar`` inputs are plain floats shared by both packed
        lanes (window samples, twiddles), the others have the value type T."""
        self.params.append(("sin" if scalar else "in", cname, count))
        vs = []
        for i in range(count):
            n = self.g.node("in", f"{cname}[{i}]")
            if scalar:
                self.scalars.add(n.id)
            vs.append(V(self.g, n))
        return vs

    def out_array(self, cname, count):
        self.params.append(("out", cname, count))

    def emit_out(self, cexpr, v):
        self.outputs.append((cexpr, v))

    # ---- evaluation (self test) ---------------------------------------------------------
    def evaluate(self, env):
        val = {}
        for n in self.g.order:
            a = n.args
            if n.op == "in":
                val[n.id] = env[a[0]]
            elif n.op == "add":
                val[n.id] = val[a[0].id] + val[a[1].id]
            elif n.op == "sub":
                val[n.id] = val[a[0].id] - val[a[1].id]
            elif n.op == "mulc":
                val[n.id] = val[a[0].id] * a[1]
            elif n.op == "mulv":
                val[n.id] = val[a[0].id] * val[a[1].id]
            elif n.op == "fmac":
                val[n.id] = val[a[0].id] * a[1] + val[a[2].id]
            elif n.op == "fmav":
                val[n.id] = val[a[0].id] * val[a[1].id] + val[a[2].id]
            elif n.op == "fnmav":
                val[n.id] = val[a[2].id] - val[a[0].id] * val[a[1].id]
            else:
                raise ValueError(n.op)
        res = {}
        for cexpr, v in self.outputs:
            res[cexpr] = 0.0 if v.zero else v.s * val[v.n.id]
        return res

    # ---- liveness + emission -----------------------------------------------------------
    def live_nodes(self):
        live = set()
        stack = [v.n for _, v in self.outputs if not v.zero]
        while stack:
            n = stack.pop()
            if n.id in live:
                continue
            live.add(n.id)
            for a in n.args:
                if isinstance(a, Node):
                    stack.append(a)
        nodes = [n for n in self.g.order if n.id in live]
        if ORDER == "level":
            # experiment: emit by dependency level (all operations whose operands are ready first), i.e. with the
            # largest distance between dependent instructions; same operations, same results, another schedule
            level = {}
            for n in nodes:
                level[n.id] = 0 if n.op == "in" else 1 + max(level[a.id] for a in n.args if isinstance(a, Node))
            nodes.sort(key=lambda n: (level[n.id], n.id))
        return nodes

    def op_counts(self):
        cnt = {}
        for n in self.live_nodes():
            if n.op != "in":
                cnt[n.op] = cnt.get(n.op, 0) + 1
        cnt["neg_out"] = sum(1 for _, v in self.outputs if not v.zero and v.s < 0)
        cnt["total"] = sum(v for k, v in cnt.items())
        return cnt

    def emit(self):
        lines = []
        cnt = self.op_counts()
        lines.append(f"// {self.doc}")
        lines.append("// ops: " + ", ".join(f"{k}={v}" for k, v in sorted(cnt.items())))
        sig = []
        for kind, cname, count in self.params:
            if kind == "in":
                sig.append(f"const T (&{cname})[{count}]")
            elif kind == "sin":
                sig.append(f"const float (&{cname})[{count}]")
            else:
                sig.append(f"T (&{cname})[{count}]")
        lines.append("template <typename T>")
        lines.append(f"LM_HD void {self.name}({', '.join(sig)}) {{")
        for n in self.live_nodes():
            a = n.args
            if n.op == "in":
                n.name = a[0]
                continue
            n.name = f"t{n.id}"
            if n.op == "add":
                e = f"vadd({a[0].name}, {a[1].name})"
            elif n.op == "sub":
                e = f"vsub({a[0].name}, {a[1].name})"
            elif n.op == "mulc":
                e = f"vmulc({a[0].name}, {_lit(a[1])})"
            elif n.op in ("mulv", "fmav", "fnmav"):
                p, q = a[0], a[1]
                sc = (p.id in self.scalars) + (q.id in self.scalars)
                assert sc < 2, "scalar x scalar product is not expected"
                if p.id in self.scalars:
                    p, q = q, p
                fn = {"mulv": "vmul", "fmav": "vfma", "fnmav": "vfnma"}[n.op] + ("s" if sc else "")
                rest = "" if n.op == "mulv" else f", {a[2].name}"
                pn, qn = p.name, q.name
                if n.op == "fnmav" and not sc:
                    # c - p*q with q (or p) a packed constant that has a negated twin: one FFMA2
                    for cand in (q, p):
                        arr = cand.name.split("[")[0] if cand.op == "in" else None
                        if arr in self.neg_twin:
                            neg = cand.name.replace(arr + "[", self.neg_twin[arr] + "[", 1)
                            other = p if cand is q else q
                            pn, qn, fn = other.name, neg, "vfma"
                            break
                e = f"{fn}({pn}, {qn}{rest})"
            elif n.op == "fmac":
                e = f"vfmac({a[0].name}, {_lit(a[1])}, {a[2].name})"
            lines.append(f"  const T {n.name} = {e};")
        for cexpr, v in self.outputs:
            if v.zero:
                lines.append(f"  {cexpr} = vzero<T>();")
            elif v.s > 0:
                lines.append(f"  {cexpr} = {v.n.name};")
            else:
                lines.append(f"  {cexpr} = vneg({v.n.name});")
        lines.append("}")
        return "\n".join(lines)


ORDER = "trace"        # "trace": creation order of the symbolic trace (depth first); "level": see live_nodes


def _lit(k: float) -> str:
    return f"{k:.9e}f"


def make_stage1(n1: int, n2_total: int, packed_consts: bool = False):
    """windowed real DFT over the coarse index + inter-stage twiddle.

    packed_consts: the thread-per-frame kernel packs two COLUMNS (b, b+1) in one f32x2, so the
    window samples and twiddles differ between the two halves and are of type T themselves; the
    negated twiddle sine comes in as its own input (FFMA2 has no operand negation).

    in : x[n1]  samples x[N2*a + b] of one frame (b = this warp's column)
         w[n1]  window values w[N2*a + b]
         twr/twi[n1/2+1]  twiddles W_N^{b*k1} (index 0 unused)
    out: yr/yi[n1/2+1]    Y[b][k1] * W_N^{b*k1}, k1 = 0..n1/2
    """
    half = n1 // 2
    c = Codelet(f"stage1_r{n1}" + ("p" if packed_consts else ""),
                f"stage 1 of N={n1 * n2_total}: window, real DFT-{n1}, twiddle; outputs k1=0..{half}"
                + ("; packed (per-half) constants" if packed_consts else ""))
    x = c.inputs("x", n1)
    w = c.inputs("w", n1, scalar=not packed_consts)
    twr = c.inputs("twr", half + 1, scalar=not packed_consts)
    twi = c.inputs("twi", half + 1, scalar=not packed_consts)
    ntwi = c.inputs("ntwi", half + 1, scalar=False) if packed_consts else None
    if packed_consts:
        c.inputs("nw", n1, scalar=False)          # -w: turns c - x*w into one FFMA2 (see Codelet.emit)
        c.inputs("ntwr", half + 1, scalar=False)
        c.neg_twin = {"w": "nw", "twi": "ntwi", "ntwi": "twi", "twr": "ntwr"}
    c.out_array("yr", half + 1)
    c.out_array("yi", half + 1)
    zero = V(c.g, None)
    xs = [C(x[a].mulv(w[a]), zero) for a in range(n1)]
    y = dft(xs)
    for k1 in range(half + 1):
        if k1 == 0:
            v = y[k1]
        elif packed_consts:
            z = y[k1]
            v = C(z.re.mulv(twr[k1]) + z.im.mulv(ntwi[k1]), z.re.mulv(twi[k1]) + z.im.mulv(twr[k1]))
        else:
            v = y[k1].mulv(twr[k1], twi[k1])
        c.emit_out(f"yr[{k1}]", v.re)
        c.emit_out(f"yi[{k1}]", v.im)
    return c


def make_stage2(n2: int, name: str, real_input: bool, outs):
    """complex DFT over the fine index, followed by |X|^2.

    in : yr/yi[n2]; out: p[len(outs)] = |X[k2]|^2 for k2 in outs.
    """
    c = Codelet(name, f"stage 2: DFT-{n2} ({'real' if real_input else 'complex'} input) and power, "
                      f"outputs k2 in {outs[0]}..{outs[-1]}")
    yr = c.inputs("yr", n2)
    zero = V(c.g, None)
    if real_input:
        ys = [C(yr[b], zero) for b in range(n2)]
    else:
        yi = c.inputs("yi", n2)
        ys = [C(yr[b], yi[b]) for b in range(n2)]
    c.out_array("p", len(outs))
    X = dft(ys)
    for j, k2 in enumerate(outs):
        z = X[k2]
        p = z.re.mulv(z.re) + z.im.mulv(z.im)
        c.emit_out(f"p[{j}]", p)
    return c


def build_all():
    cl = []
    for n1, n2 in ((20, 20), (32, 32)):
        cl.append(make_stage1(n1, n2))
        if n1 == 20:
            cl.append(make_stage1(n1, n2, packed_consts=True))
        cl.append(make_stage2(n2, f"stage2_c{n2}", False, list(range(n2))))
        cl.append(make_stage2(n2, f"stage2_c{n2}_half", False, list(range(n2 // 2))))
        cl.append(make_stage2(n2, f"stage2_r{n2}_half", True, list(range(n2 // 2 + 1))))
    return cl


HEADER = '''// GENERATED by tools/gen_codelets.py -- do not edit by hand.
//
// Register-resident DFT codelets for the two-stage STFT (see DESIGN.md, "kernels").
// T is float (host emulation, scalar device path) or lm::f32x2 (packed sm_100a path);
// the v* helpers come from vec_ops.cuh.
#pragma once
#include "vec_ops.cuh"

namespace lm {
'''


def selftest():
    import random
    random.seed(1)
    ok = True
    for c in build_all():
        env = {}
        arrays = {}
        for kind, cname, count in c.params:
            if kind in ("in", "sin"):
                arrays[cname] = [random.uniform(-1, 1) for _ in range(count)]
                for i in range(count):
                    env[f"{cname}[{i}]"] = arrays[cname][i]
        res = c.evaluate(env)
        if c.name.startswith("stage1"):
            n1 = len(arrays["x"])
            half = n1 // 2
            err = 0.0
            for k1 in range(half + 1):
                ref = sum(arrays["x"][a] * arrays["w"][a] * cmath.exp(-2j * math.pi * a * k1 / n1)
                          for a in range(n1))
                if k1 and "ntwi" in arrays:
                    tw = complex(arrays["twr"][k1], arrays["twi"][k1])
                    # the codelet is told -twi separately; a consistent table has ntwi == -twi
                    ref = complex(ref.real * tw.real + ref.imag * arrays["ntwi"][k1],
                                  ref.real * tw.imag + ref.imag * tw.real)
                elif k1:
                    ref *= complex(arrays["twr"][k1], arrays["twi"][k1])
                err = max(err, abs(ref - complex(res[f"yr[{k1}]"], res[f"yi[{k1}]"])))
        else:
            n2 = len(arrays["yr"])
            yi = arrays.get("yi", [0.0] * n2)
            outs = [int(k) for k in range(len([o for o in c.outputs]))]
            err = 0.0
            for j, (cexpr, _) in enumerate(c.outputs):
                k2 = j
                ref = sum(complex(arrays["yr"][b], yi[b]) * cmath.exp(-2j * math.pi * b * k2 / n2)
                          for b in range(n2))
                err = max(err, abs(abs(ref) ** 2 - res[cexpr]))
        cnt = c.op_counts()
        print(f"{c.name:22s} err={err:.2e} ops={cnt['total']:4d} {cnt}", file=sys.stderr)
        ok = ok and err < 1e-11
    return ok


def main():
    global ORDER
    if "--level-order" in sys.argv:
        ORDER = "level"
    if "--selftest" in sys.argv:
        sys.exit(0 if selftest() else 1)
    print(HEADER)
    for c in build_all():
        print(c.emit())
        print()
    print("}  // namespace lm")


if __name__ == "__main__":
    main()
