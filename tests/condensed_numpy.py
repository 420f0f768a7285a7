"""NumPy mirror of the CUDA setup pipeline (condensed row-space formulation).

Test helper only: it restates, stage by stage, what
``direct_data_driven_mpc_b200/csrc`` computes per controller so that device
intermediates (Gram matrix, reduced Hessian, gains, ADMM operators) can be
compared stage-wise, and so the condensation itself is verified on the CPU
against the literal-KKT oracle.  DESIGN.md "Condensed formulation" has the math.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

NOMINAL, ROBUST = 0, 1
SLACK_NONE, SLACK_CONVEX = 0, 1


def hankel(X, L):
    N, nch = X.shape
    flat = np.ascontiguousarray(X).reshape(-1)
    idx = np.arange(L * nch)[:, None] + nch * np.arange(N - L + 1)[None, :]
    return flat[idx]


class Plan:
    """All per-controller constants the device keeps."""


def theta_layout(n, m, p):
    """theta = [u_past (n*m); y_past (n*p); u_s (m); y_s (p)]"""
    o_up, o_yp, o_us, o_ys = 0, n * m, n * m + n * p, n * m + n * p + m
    return o_up, o_yp, o_us, o_ys, n * (m + p) + m + p


def build_plan(n, m, p, u_d, y_d, L, Q, R, eps_max, lamb_alpha, lamb_sigma, c, slack_type,
               ctrl_type, use_terminal, rank_tol=1e-11, input_bounds=None, output_bounds=None):
    Lp = L + n
    nu, ny = Lp * m, Lp * p
    H = np.vstack([hankel(u_d, Lp), hankel(y_d, Lp)])
    r = H.shape[0]
    o_up, o_yp, o_us, o_ys, nth = theta_layout(n, m, p)
    pl = Plan()
    pl.n, pl.m, pl.p, pl.L, pl.nth, pl.H = n, m, p, L, nth, H
    W = H @ H.T
    pl.W = W
    Q = 0.5 * (Q + Q.T)
    R = 0.5 * (R + R.T)
    robust = ctrl_type == ROBUST
    pl.robust = robust
    pl.convex = robust and slack_type == SLACK_CONVEX
    # ---- x = [ubar (nu); ybar (ny); sigma (ny, robust)] : fixed/free split ----
    nx = nu + ny + (ny if robust else 0)
    fixed = np.zeros(nx, bool)
    fixed[0:n * m] = True
    fixed[nu:nu + n * p] = True
    if use_terminal:
        fixed[L * m:Lp * m] = True
        fixed[nu + L * p:nu + Lp * p] = True
    free = np.where(~fixed)[0]
    fix = np.where(fixed)[0]
    pl.free, pl.fix, pl.nx = free, fix, nx
    # x_c = Cth theta (selection / tiling)
    Cth = np.zeros((nx, nth))
    for i in range(n * m):
        Cth[i, o_up + i] = 1.0
    for i in range(n * p):
        Cth[nu + i, o_yp + i] = 1.0
    if use_terminal:
        for k in range(n):
            for j in range(m):
                Cth[(L + k) * m + j, o_us + j] = 1.0
            for j in range(p):
                Cth[nu + (L + k) * p + j, o_ys + j] = 1.0
    # D x  (cost weights) and linear term q/2 = -D Tile theta
    D = np.zeros((nx, nx))
    D[n * m:nu, n * m:nu] = R
    D[nu + n * p:nu + ny, nu + n * p:nu + ny] = Q
    Tile = np.zeros((nx, nth))
    for k in range(L):
        for j in range(m):
            Tile[(n + k) * m + j, o_us + j] = 1.0
        for j in range(p):
            Tile[nu + (n + k) * p + j, o_ys + j] = 1.0
    if robust:
        D[nu + ny:, nu + ny:] = lamb_sigma * np.eye(ny)
        # t = T x
        T = np.zeros((r, nx))
        T[:nu, :nu] = np.eye(nu)
        T[nu:, nu:nu + ny] = np.eye(ny)
        T[nu:, nu + ny:] = np.eye(ny)
        # Short data (fewer Hankel columns than rows: N - L - n + 1 < (L + n)(m + p), allowed by the reference down to
        # N_min): W is singular, t = T x must stay in range(H) and ||alpha||^2 = t^T W^+ t there.  Null-space rows of W
        # become the equality constraint Aeq x = 0, resolved by a Schur complement on top of the same reduced Hessian.
        lam_w, V_w = np.linalg.eigh(W)
        null = lam_w <= 1e-12 * lam_w.max()
        Aeq = None
        if null.any():
            Om = (V_w[:, ~null] / lam_w[~null]) @ V_w[:, ~null].T
            Aeq = V_w[:, null].T @ T
        else:
            Lw = np.linalg.cholesky(W)
            Om = sla.cho_solve((Lw, True), np.eye(r))
        Om = 0.5 * (Om + Om.T)
        pl.Om = Om
        P = D + (lamb_alpha * eps_max) * (T.T @ Om @ T)
        pl.P = P
        A = P[np.ix_(free, free)]
        Nmat = P[np.ix_(free, fix)] @ Cth[fix, :] - (D @ Tile)[free, :]
        La = np.linalg.cholesky(A)
        X0f = -sla.cho_solve((La, True), Nmat)          # x_f = X0f theta
        if Aeq is not None:
            Af, Ac = Aeq[:, free], Aeq[:, fix]
            Yq = sla.cho_solve((La, True), Af.T)
            Sq = Af @ Yq
            Sq = 0.5 * (Sq + Sq.T)
            X0f = X0f - Yq @ np.linalg.solve(Sq, Af @ X0f + Ac @ Cth[fix, :])
        X0 = np.zeros((nx, nth))
        X0[free] = X0f
        X0[fix] = Cth[fix]
        pl.A, pl.Nmat, pl.X0 = A, Nmat, X0
        pl.Ku = X0[n * m:nu, :]                         # optimal_u = Ku theta   (L*m, nth)
        # cost J0(theta) = x^T P x - 2 (D Tile theta)^T x + theta^T Tile^T D Tile theta
        pl.Z = X0.T @ P @ X0 - X0.T @ (D @ Tile) - (D @ Tile).T @ X0 + Tile.T @ D @ Tile
        pl.Z = 0.5 * (pl.Z + pl.Z.T)
        pl.feasF = None
        pl.nb = 0
        if pl.convex or input_bounds is not None or output_bounds is not None:
            # box rows: sigma_pred (CONVEX) then the free predicted inputs (input box), as in setup.cu
            pos = {g: i for i, g in enumerate(free)}
            rows, lo, hi = [], [], []
            if pl.convex:
                rows += [nu + ny + n * p + j for j in range(L * p)]
                lo += [-c * eps_max] * (L * p)
                hi += [c * eps_max] * (L * p)
            nbs = len(rows)
            if input_bounds is not None:
                bl = np.broadcast_to(np.asarray(-np.inf if input_bounds[0] is None else input_bounds[0], float).reshape(-1), (m,))
                bh = np.broadcast_to(np.asarray(np.inf if input_bounds[1] is None else input_bounds[1], float).reshape(-1), (m,))
                for j in range((L - n) * m if use_terminal else L * m):
                    rows.append(n * m + j); lo.append(bl[j % m]); hi.append(bh[j % m])
            nbu_end = len(rows)
            if output_bounds is not None:
                bl = np.broadcast_to(np.asarray(-np.inf if output_bounds[0] is None else output_bounds[0], float).reshape(-1), (p,))
                bh = np.broadcast_to(np.asarray(np.inf if output_bounds[1] is None else output_bounds[1], float).reshape(-1), (p,))
                for j in range((L - n) * p if use_terminal else L * p):
                    rows.append(nu + n * p + j); lo.append(bl[j % p]); hi.append(bh[j % p])
            nb = len(rows)
            bidx = np.array([pos[g] for g in rows])
            Bsel = np.zeros((nb, len(free)))
            Bsel[np.arange(nb), bidx] = 1.0
            Y = sla.cho_solve((La, True), Bsel.T)       # A^-1 B^T
            if Aeq is not None:
                Y = Y - Yq @ np.linalg.solve(Sq, Af @ Y)
            Lam = Bsel @ Y
            dg = np.diag(Lam)
            rs = np.ones(nb)                            # group equilibration as in k_lam_rho (first non-empty group = 1)
            groups = [g for g in (slice(0, nbs), slice(nbs, nbu_end), slice(nbu_end, nb)) if g.stop > g.start]
            for g in groups[1:]:
                rs[g] = np.sqrt(dg[groups[0]].mean() / dg[g].mean())
            Y = Y * rs[None, :]
            Lam = rs[:, None] * (0.5 * (Lam + Lam.T)) * rs[None, :]
            rho2 = nb / np.trace(Lam)                   # rho/2
            Phi = np.linalg.inv(np.eye(nb) + rho2 * Lam)
            Yfull = np.zeros((nx, nb))
            Yfull[free] = Y
            pl.Ks = rs[:, None] * X0[rows, :]           # s_unc = Ks theta (scaled rows)
            pl.Phi, pl.Lam, pl.rho2 = 0.5 * (Phi + Phi.T), Lam, rho2
            pl.Psi = rho2 * Yfull[n * m:nu, :]          # u = u0 + Psi (v - s)
            pl.bound = c * eps_max if pl.convex else 0.0
            pl.nb, pl.nbs, pl.rs = nb, nbs, rs
            pl.lo, pl.hi = rs * np.array(lo), rs * np.array(hi)
            fin = np.concatenate([np.abs(pl.lo[np.isfinite(pl.lo)]), np.abs(pl.hi[np.isfinite(pl.hi)]), [0.0]])
            pl.bmax = float(fin.max())
    else:
        # nominal: t = [ubar; ybar] in range(H)
        lam, V = np.linalg.eigh(W)
        keep = lam > rank_tol * lam.max()
        U = V[:, keep]
        pl.rank = int(keep.sum())
        Pi = U @ U.T
        E = np.zeros((len(fix), nx)); E[np.arange(len(fix)), fix] = 1.0
        Mc = E @ Pi @ E.T
        lc, Vc = np.linalg.eigh(Mc)
        kc = lc > 1e-9 * lc.max()
        Mc_p = (Vc[:, kc] / lc[kc]) @ Vc[:, kc].T
        tp_map = Pi @ E.T @ Mc_p @ Cth[fix, :]           # t_p = tp_map theta
        Pi0 = Pi - Pi @ E.T @ Mc_p @ E @ Pi
        G0 = Pi0 @ D @ Pi0
        G0 = 0.5 * (G0 + G0.T)
        l0, V0 = np.linalg.eigh(G0)
        k0 = l0 > 1e-11 * max(l0.max(), 1e-300)
        G0p = (V0[:, k0] / l0[k0]) @ V0[:, k0].T
        X0 = tp_map + Pi0 @ G0p @ Pi0 @ (D @ Tile - D @ tp_map)
        pl.X0 = X0
        pl.Ku = X0[n * m:nu, :]
        DT = D @ Tile
        pl.Z = X0.T @ D @ X0 - X0.T @ DT - DT.T @ X0 + Tile.T @ DT
        pl.Z = 0.5 * (pl.Z + pl.Z.T)
        pl.feasF = (Mc @ Mc_p - np.eye(len(fix))) @ Cth[fix, :]
    return pl


def make_theta(n, m, p, u_past, y_past, u_s, y_s):
    return np.concatenate([np.reshape(u_past, -1), np.reshape(y_past, -1), np.reshape(u_s, -1), np.reshape(y_s, -1)])


def solve(pl: Plan, theta, tol=1e-9, max_iter=500, relax=1.8):
    """Returns (optimal_u, cost, status, iters)."""
    u = pl.Ku @ theta
    cost = float(theta @ pl.Z @ theta)
    if pl.feasF is not None:
        fe = np.abs(pl.feasF @ theta).max()
        if fe > 1e-6 * (1.0 + np.abs(theta).max()):
            return u, cost, "infeasible", 1
    if not getattr(pl, "nb", 0):
        return u, cost, "optimal", 1
    s_unc = pl.Ks @ theta
    lo, hi = pl.lo, pl.hi
    if not np.any((s_unc < lo) | (s_unc > hi)):
        return u, cost, "optimal", 1
    z = np.clip(s_unc, lo, hi)
    w = np.zeros_like(z)
    status = "optimal_inaccurate"
    it = 0
    thr = tol * max(pl.bmax, np.abs(s_unc).max())
    for it in range(1, max_iter + 1):
        v = z - w
        s = v + pl.Phi @ (s_unc - v)
        sr = relax * s + (1 - relax) * z
        z_new = np.clip(sr + w, lo, hi)
        w = w + sr - z_new
        r_pri = np.abs(s - z_new).max()
        r_dua = np.abs(z_new - z).max()
        z = z_new
        if max(r_pri, r_dua) <= thr:
            status = "optimal"
            break
    v = z - w
    s = v + pl.Phi @ (s_unc - v)
    dv = v - s
    u = u + pl.Psi @ dv
    cost = cost + pl.rho2 ** 2 * float(dv @ pl.Lam @ dv)
    return u, cost, status, it
