"""Rigorous range analysis of the reference's 16-bit IDCT lane (idct_sse2.hpp:23-65).

Every intermediate of idct_1d_sse2 is a linear form of the 8 lane inputs plus a bounded rounding error
(each _mm_mulhi_epi16 floors: error in [0,1)).  Propagating (coefficients, error bound) through the
flow graph gives, for every operation, max |value| = sum_k |C_k| * X_k + E over all inputs with
|x_k| <= X_k.  Where that stays below 32767 the saturating add / wrapping shift can never trigger and
the kernel may use plain arithmetic -- bit-exactly.

Outputs: (1) which pass-1 operations can saturate given dequantised inputs (|F| <= 2048, first
coefficient <= 3036); (2) per-input gains G_k (max over all intermediates) and output gains O_jk used
by the kernel's per-block bound for a saturation-free pass 2.
"""
import numpy as np


class V:
    def __init__(self, c, e):
        self.c, self.e = np.asarray(c, float), float(e)


def lane(x, log):
    def note(name, v):
        log.append((name, v))
        return v

    def mulhi(a, c, name):
        return note(name, V(a.c * c / 65536.0, a.e * abs(c) / 65536.0 + 1.0))

    def shl(a, n, name):
        return note(name, V(a.c * (1 << n), a.e * (1 << n)))

    def add(a, b, name):
        return note(name, V(a.c + b.c, a.e + b.e))

    def sub(a, b, name):
        return note(name, V(a.c - b.c, a.e + b.e))

    v15 = add(shl(mulhi(x[0], 27145, "m0"), 1, "m0s"), shl(x[0], 1, "x0s"), "v15")
    v26 = add(mulhi(x[1], -5037, "m1"), shl(x[1], 2, "x1s"), "v26")
    v21 = add(mulhi(x[2], -19954, "m2"), shl(x[2], 2, "x2s"), "v21")
    v28 = add(shl(mulhi(x[3], -22089, "m3"), 1, "m3s"), shl(x[3], 2, "x3s"), "v28")
    v16 = add(shl(mulhi(x[4], 27145, "m4"), 1, "m4s"), shl(x[4], 1, "x4s"), "v16")
    v25 = add(mulhi(x[5], 14567, "m5"), shl(x[5], 1, "x5s"), "v25")
    v22 = add(shl(mulhi(x[6], 17391, "m6"), 1, "m6s"), x[6], "v22")
    v27 = shl(mulhi(x[7], 25570, "m7"), 1, "v27")
    v19 = sub(v25, v28, "v19"); v20 = sub(v26, v27, "v20"); v23 = add(v26, v27, "v23"); v24 = add(v25, v28, "v24")
    v7 = add(v23, v24, "v7"); v11 = add(v21, v22, "v11"); v13 = sub(v23, v24, "v13"); v17 = sub(v21, v22, "v17")
    v8 = add(v15, v16, "v8"); v9 = sub(v15, v16, "v9")
    v18 = mulhi(sub(v19, v20, "v19-v20"), 25079, "v18")
    v12 = sub(v18, add(v19, mulhi(v19, 20090, "op3m"), "op3"), "v12")
    v14 = sub(sub(v20, mulhi(v20, 30068, "op1m"), "op1"), v18, "v14")
    v6 = sub(shl(v14, 1, "v14s"), v7, "v6")
    v5 = sub(add(v13, mulhi(v13, 27145, "op0am"), "op0a"), v6, "v5")
    v4 = add(v5, shl(v12, 1, "v12s"), "v4")
    v10 = sub(add(v17, mulhi(v17, 27145, "op0bm"), "op0b"), v11, "v10")
    v0 = add(v8, v11, "v0"); v1 = add(v9, v10, "v1"); v2 = sub(v9, v10, "v2"); v3 = sub(v8, v11, "v3")
    outs = [add(v0, v7, "o0"), add(v1, v6, "o1"), add(v2, v5, "o2"), sub(v3, v4, "o3"),
            add(v3, v4, "o4"), sub(v2, v5, "o5"), sub(v1, v6, "o6"), sub(v0, v7, "o7")]
    return outs


def analyse():
    log = []
    x = [V(np.eye(8)[k], 0.0) for k in range(8)]
    outs = lane(x, log)
    return log, outs


if __name__ == "__main__":
    log, outs = analyse()
    X = np.array([3036.0] + [2048.0] * 7)     # pass-1 input magnitudes after dequantisation
    print("pass 1 with |F0| <= 3036, |F| <= 2048: worst-case magnitude of every intermediate")
    worst = 0
    for name, v in log:
        m = float(np.abs(v.c) @ X + v.e)
        worst = max(worst, m)
        flag = "  <-- can exceed int16" if m > 32767 else ""
        print("  %-8s %9.1f%s" % (name, m, flag))
    print("max", worst)
    G = np.max(np.array([np.abs(v.c) for _, v in log]), axis=0)
    E = max(v.e for _, v in log)
    O = np.array([np.abs(o.c) for o in outs])      # |O_jk|
    print("G_k  (max |coef| of input k over all intermediates):", np.round(G, 4))
    print("max rounding error bound over intermediates:", E)
    print("Omax_k (max_j |O_jk|):", np.round(O.max(axis=0), 4), " max output error", max(o.e for o in outs))
