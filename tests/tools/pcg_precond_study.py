"""Development tool (CPU, uses the oracle's linearisation): PCG iteration counts of the inexact LM on the C5 graph for
candidate preconditioners -- what a kernel change would buy before it is written.

    python tests/tools/pcg_precond_study.py [seg16 seg32 add16 mult16 half whole ...]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from omniswarm_b200 import synth
from oracle import solver_ref as R


def build(g):
    n = g["init"].shape[0]; m = len(g["ftype"])
    free = np.nonzero(g["fixed"] == 0)[0]
    col_of = -np.ones(n, np.int64); col_of[free] = np.arange(len(free))
    ia, ib = g["ia"], g["ib"]
    rows = (np.arange(m)[:, None, None] * 4 + np.arange(4)[None, :, None]) + np.zeros((1, 1, 4), np.int64)
    ca = (col_of[ia][:, None, None] * 4 + np.arange(4)[None, None, :]) + np.zeros((1, 4, 1), np.int64)
    cb = (col_of[ib][:, None, None] * 4 + np.arange(4)[None, None, :]) + np.zeros((1, 4, 1), np.int64)
    ma = np.broadcast_to((col_of[ia] >= 0)[:, None, None], (m, 4, 4))
    mb = np.broadcast_to((col_of[ib] >= 0)[:, None, None], (m, 4, 4))
    Rr = np.concatenate([rows[ma], rows[mb]]); Cc = np.concatenate([ca[ma], cb[mb]])
    return free, col_of, ma, mb, Rr, Cc


def chain_layout(g, n_drones, free, col_of):
    """position of every free node on its drone's chain: (chain id, index along the chain)"""
    n = g["init"].shape[0]
    chain = np.arange(n) % n_drones; pos = np.arange(n) // n_drones
    return chain[free], pos[free]


def pattern(chain, pos, seg, offset, nv):
    """block mask: same node, or chain neighbours inside the same segment (seg = 0: whole chain)"""
    k = len(chain)
    key = {(c, p): i for i, (c, p) in enumerate(zip(chain, pos))}
    I, Jc = list(range(k)), list(range(k))
    for i, (c, p) in enumerate(zip(chain, pos)):
        j = key.get((c, p + 1))
        if j is None: continue
        if seg and ((p + offset) // seg != (p + 1 + offset) // seg): continue
        I += [i, j]; Jc += [j, i]
    B = sp.csr_matrix((np.ones(len(I)), (I, Jc)), shape=(k, k))
    return sp.kron(B, np.ones((4, 4))).tocsr()


def run(g, n_drones, precond, tol=1e-2, verbose=False):
    free, col_of, ma, mb, Rr, Cc = build(g)
    chain, pos = chain_layout(g, n_drones, free, col_of)
    nv = 4 * len(free)
    pats = {}
    def pat(seg, off):
        if (seg, off) not in pats: pats[(seg, off)] = pattern(chain, pos, seg, off, nv)
        return pats[(seg, off)]
    x = g["init"].astype(np.float64).copy()
    m = len(g["ftype"])
    radius, decrease = 1e4, 2.0
    cost, r, Ja, Jb = R.evaluate_vec(g, x)
    total = 0; lm = 0
    for it in range(1, 60):
        J = sp.csr_matrix((np.concatenate([Ja[ma], Jb[mb]]), (Rr, Cc)), shape=(4 * m, nv))
        grad = J.T @ r.reshape(-1)
        if np.max(np.abs(grad)) < 1e-10: break
        H = (J.T @ J).tocsr()
        diag = np.clip(H.diagonal(), 1e-6, 1e32)
        ok = False
        while radius > 1e-32:
            A = (H + sp.diags(diag / radius)).tocsr()
            def fac(seg, off):
                return spla.splu(A.multiply(pat(seg, off)).tocsc())
            if precond == "jacobi":
                F = fac(1, 0); Minv = F.solve
            elif precond.startswith("seg"):
                F = fac(int(precond[3:]), 0); Minv = F.solve
            elif precond == "whole":
                F = fac(0, 0); Minv = F.solve
            elif precond.startswith("add"):
                L = int(precond[3:]); F1, F2 = fac(L, 0), fac(L, L // 2)
                Minv = lambda v: 0.5 * (F1.solve(v) + F2.solve(v))
            elif precond.startswith("mult"):          # symmetric multiplicative on the whole-chain tridiagonal T: A, B, A
                L = int(precond[4:]); F1, F2 = fac(L, 0), fac(L, L // 2)
                T = A.multiply(pat(0, 0)).tocsr()
                def Minv(v):
                    z = F1.solve(v); z = z + F2.solve(v - T @ z); return z + F1.solve(v - T @ z)
            elif precond.startswith("gs"):            # two-stage: A then B (non-symmetric; CG may still work in practice)
                L = int(precond[2:]); F1, F2 = fac(L, 0), fac(L, L // 2)
                T = A.multiply(pat(0, 0)).tocsr()
                def Minv(v):
                    z = F1.solve(v); return z + F2.solve(v - T @ z)
            else:
                raise SystemExit("unknown " + precond)
            # PCG
            b = -grad; xk = np.zeros(nv); rk = b.copy(); z = Minv(rk); p = z.copy(); rz = rk @ z; rr0 = rk @ rk
            k = 0
            while k < 2000:
                Ap = A @ p; alpha = rz / (p @ Ap); xk += alpha * p; rk -= alpha * Ap; k += 1
                if rk @ rk <= tol * tol * rr0: break
                z = Minv(rk); rz2 = rk @ z; p = z + (rz2 / rz) * p; rz = rz2
            total += k
            delta = xk
            xn = x.copy(); xn[free] += delta.reshape(-1, 4)
            new_cost = R.evaluate_vec(g, xn, want_jac=False)[0]
            model = -(grad @ delta) - 0.5 * (delta @ (H @ delta))
            rho = (cost - new_cost) / model if model > 0 else -1
            if rho > 1e-3:
                radius = min(radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 1e16); decrease = 2.0
                x = xn; ok = True
                break
            radius /= decrease; decrease *= 2.0
        lm += 1
        if not ok: break
        old = cost
        cost, r, Ja, Jb = R.evaluate_vec(g, x)
        if verbose: print(it, cost, total)
        if abs(old - cost) <= 1e-6 * old: break
    return lm, total, cost


if __name__ == "__main__":
    g = synth.pose_graph_c5(0)
    for name in (sys.argv[1:] or ["jacobi", "seg16", "seg32", "add16", "mult16", "gs16", "seg128", "whole"]):
        lm, total, cost = run(g, 5, name)
        print(f"{name:8s} LM {lm:2d}  PCG {total:5d}  cost {cost:.6f}", flush=True)
