"""Independent numpy/scipy twin of the numeric core of Optimizer::OptimizeEssentialGraph, used only to pin oracle_optimize_pose_graph.

Same Levenberg control flow as g2o (lambda from setUserLambdaInit, gain ratio with the +1e-3, nu doubling, the extra stop rule) but different numerics on
purpose: vertices and measurements are 4x4 similarity matrices [sR t; 0 1], the edge error is read off scipy.linalg.logm of Sji * Siw * Sjw^-1 (g2o: the closed
form of Sim3::log), the update is scipy.linalg.expm of the sim(3) generator, inverses are matrix inverses, the Jacobians are central differences with a 1e-6
step (g2o: 1e-9) and the normal equations are solved densely with numpy."""
import numpy as np
import scipy.linalg


def to_matrix(s):
    from scipy.spatial.transform import Rotation
    T = np.eye(4); T[:3, :3] = s[7] * Rotation.from_quat(s[:4] / np.linalg.norm(s[:4])).as_matrix(); T[:3, 3] = s[4:7]
    return T


def _gen(u):
    w, v, s = u[:3], u[3:6], u[6]
    A = np.zeros((4, 4))
    A[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) + s * np.eye(3)
    A[:3, 3] = v
    return A


def _log(T):
    L = np.real(scipy.linalg.logm(T))
    sigma = (L[0, 0] + L[1, 1] + L[2, 2]) / 3.0
    return np.array([L[2, 1], L[0, 2], L[1, 0], L[0, 3], L[1, 3], L[2, 3], sigma])


class Twin:
    def __init__(self, sim3, fixed, e_i, e_j, e_meas, fix_scale):
        self.V = [to_matrix(np.asarray(s, float)) for s in sim3]
        self.fixed = np.asarray(fixed) > 0
        self.ei, self.ej = np.asarray(e_i), np.asarray(e_j)
        self.M = [to_matrix(np.asarray(m, float)) for m in e_meas]
        self.fix = bool(fix_scale)
        self.hidx = np.full(len(self.V), -1); self.hidx[~self.fixed] = np.arange((~self.fixed).sum())
        self.iterations = self.trials = 0

    def err(self, e, Vi=None, Vj=None):
        Vi = self.V[self.ei[e]] if Vi is None else Vi
        Vj = self.V[self.ej[e]] if Vj is None else Vj
        return _log(self.M[e] @ Vi @ np.linalg.inv(Vj))

    def oplus(self, T, u):
        u = np.array(u, float)
        if self.fix:
            u[6] = 0
        return scipy.linalg.expm(_gen(u)) @ T

    def chi2(self):
        return sum(float(self.err(e) @ self.err(e)) for e in range(len(self.ei)) if self.hidx[self.ei[e]] >= 0 or self.hidx[self.ej[e]] >= 0)

    def optimize(self, iterations=20, lambda_init=1e-16):
        n = 7 * int((~self.fixed).sum())
        lam, ni, nbad = lambda_init, 2.0, 0
        h = 1e-6
        for it in range(iterations):
            cur = ini = self.chi2()
            H = np.zeros((n, n)); b = np.zeros(n)
            for e in range(len(self.ei)):
                ids = (self.ei[e], self.ej[e])
                hs = (self.hidx[ids[0]], self.hidx[ids[1]])
                if hs[0] < 0 and hs[1] < 0:
                    continue
                r = self.err(e)
                J = [None, None]
                for s in range(2):
                    if hs[s] < 0:
                        continue
                    J[s] = np.zeros((7, 7))
                    for d in range(7):
                        u = np.zeros(7); u[d] = h
                        Vp, Vm = self.oplus(self.V[ids[s]], u), self.oplus(self.V[ids[s]], -u)
                        ep = self.err(e, Vp, None) if s == 0 else self.err(e, None, Vp)
                        em = self.err(e, Vm, None) if s == 0 else self.err(e, None, Vm)
                        J[s][:, d] = (ep - em) / (2 * h)
                for s in range(2):
                    if hs[s] < 0:
                        continue
                    b[7 * hs[s]:7 * hs[s] + 7] -= J[s].T @ r
                    for t in range(2):
                        if hs[t] >= 0:
                            H[7 * hs[s]:7 * hs[s] + 7, 7 * hs[t]:7 * hs[t] + 7] += J[s].T @ J[t]
            if it == 0:
                lam, ni, nbad = lambda_init, 2.0, 0
            q, rho = 0, 0.0
            while True:
                backup = [v.copy() for v in self.V]
                x = np.linalg.solve(H + lam * np.eye(n), b)
                for k in range(len(self.V)):
                    if self.hidx[k] >= 0:
                        self.V[k] = self.oplus(self.V[k], x[7 * self.hidx[k]:7 * self.hidx[k] + 7])
                tmp = self.chi2()
                rho = (cur - tmp) / (x @ (lam * x + b) + 1e-3)
                if rho > 0 and np.isfinite(tmp):
                    lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3)); ni = 2.0; cur = tmp
                else:
                    lam *= ni; ni *= 2; self.V = backup
                q += 1; self.trials += 1
                if not (rho < 0 and q < 10):
                    break
            self.iterations += 1
            if q == 10 or rho == 0:
                break
            nbad = nbad + 1 if (ini - cur) * 1e3 < ini else 0
            if nbad >= 3:
                break
        return np.stack(self.V)
