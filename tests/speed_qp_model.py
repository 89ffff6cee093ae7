"""numpy model of the speed-QP kernel's linear algebra (csrc/st_qp.cu): the KKT systems of the coneqp iteration solved
through the structure the kernel uses - S = P + G' diag(w) G as three tridiagonal chains (one per component
s / s_dot / s_dot2) with bidiagonal Cholesky factors, Y = L^-1 A', dense K = Y'Y and its Cholesky factor - instead
of the dense chol2 of oracle/coneqp.py.  Plugged into the oracle's own iteration (coneqp.qp(..., kkt=...)) it must
reproduce the oracle's iterates: that is what tests/test_host_logic.py checks on the golden frames."""
import numpy as np


class ChainKkt:
    """Drop-in for oracle.coneqp._Chol2, valid for the speed QP's matrices (oracle/st_back.speed_qp_cone_form)."""

    def __init__(self, P, G, A):
        self.P, self.G, self.A = P, G, A
        self.nv = P.shape[0]
        self.q = self.nv // 3
        # structure checks: every row of G touches one variable or (s_i, s_{i+1}); P couples only s_dot2 neighbours
        for row in G:
            nz = np.nonzero(row)[0]
            assert len(nz) == 1 or (len(nz) == 2 and nz[0] % 3 == 0 and nz[1] == nz[0] + 3)
        off = P - np.diag(np.diag(P))
        for i, j in zip(*np.nonzero(off)):
            assert i % 3 == 2 and j % 3 == 2 and abs(i - j) == 3

    def factor(self, d):
        self.d = d
        w = 1.0 / (d * d)
        S = self.P + self.G.T @ (w[:, None] * self.G)
        q = self.q
        self.ldi = np.zeros((3, q))                       # inverse diagonal of the chain factors
        self.le = np.zeros((3, q))                        # sub-diagonal of the chain factors
        for a in range(3):
            prev = 0.0
            for k in range(q):
                v = 3 * k + a
                dd = S[v, v] - prev * prev
                if not dd > 0.0:
                    raise np.linalg.LinAlgError("chain pivot")
                inv = 1.0 / np.sqrt(dd)
                e = S[v + 3, v] * inv if k < q - 1 else 0.0
                self.ldi[a, k], self.le[a, k] = inv, e
                prev = e
        # the chains carry all of S: nothing couples different components
        mask = np.zeros_like(S, dtype=bool)
        for v in range(self.nv):
            mask[v, v] = True
            if v + 3 < self.nv:
                mask[v, v + 3] = mask[v + 3, v] = True
        assert np.all(S[~mask] == 0.0)
        self.Y = np.stack([self._fwd(col) for col in self.A], axis=1)       # L^-1 A'
        K = self.Y.T @ self.Y
        self.LK = np.linalg.cholesky(K)
        return self

    def _fwd(self, b):
        y = np.zeros(self.nv)
        for a in range(3):
            prev = 0.0
            for k in range(self.q):
                v = 3 * k + a
                prev = (b[v] - (self.le[a, k - 1] if k > 0 else 0.0) * prev) * self.ldi[a, k]
                y[v] = prev
        return y

    def _bwd(self, b):
        x = np.zeros(self.nv)
        for a in range(3):
            nxt = 0.0
            for k in range(self.q - 1, -1, -1):
                v = 3 * k + a
                nxt = (b[v] - self.le[a, k] * nxt) * self.ldi[a, k]
                x[v] = nxt
        return x

    def solve(self, bx, by, bz):
        z = bz / self.d
        x = self._fwd(bx + self.G.T @ (z / self.d))
        y = self.Y.T @ x - by
        y = np.linalg.solve(self.LK.T, np.linalg.solve(self.LK, y))
        x = self._bwd(x - self.Y @ y)
        return x, y, (self.G @ x) / self.d - z
