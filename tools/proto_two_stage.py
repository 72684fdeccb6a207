"""numpy prototype of the two-stage tridiagonalisation behind `mb_set_param("sytrd_mode", 3)` (csrc/sbr.cu).

Stage 1  dense -> band (b sub-diagonals): per panel a Householder QR of the block below the band (compact WY: V, T),
         Y = A22 V T, S = T' V' Y, W = Y - 1/2 V S, A22 -= V W' + W V', z <- Q' z.
Stage 2  band -> tridiagonal by bulge chasing (one reflector per sweep and step, Lang's scheme): sweep s may run step k
         once sweep s-1 has finished step k+1.

The functions mirror the kernels index for index (same panel / block boundaries, same reduction contents) so that the CUDA
code can be checked against them; `python tools/proto_two_stage.py` runs the self-checks (eigenvalues against
numpy.linalg.eigvalsh, RSS(lambda) through the transformed right-hand side, a randomised schedule of stage 2).
Test infrastructure only - nothing under machisplin_b200/ imports this file."""
from __future__ import annotations

import numpy as np


def house(alpha: float, xnorm2: float):
    """dlarfg for x = [alpha; x1], |x1|^2 = xnorm2  ->  beta, tau, scale  (v = [1; x1 * scale])."""
    if xnorm2 == 0.0:
        return alpha, 0.0, 0.0
    beta = -np.copysign(np.sqrt(alpha * alpha + xnorm2), alpha)
    return beta, (beta - alpha) / beta, 1.0 / (alpha - beta)


def panel_qr(P: np.ndarray):
    """In-place Householder QR of the r x b block P, one fused reduction per column (what k_sbr_qr does).
    Returns V (r x b, explicit unit diagonal, zeros above), T (b x b upper triangular), and leaves R in P's upper triangle."""
    r, b = P.shape
    nref = min(b, r - 1)
    V = np.zeros((r, b))
    T = np.zeros((b, b))
    X = P  # rows are "threads"
    for j in range(nref):
        # reduction over rows i > j of x_i[j] * x_i[c], all c
        s = X[j + 1:, j] @ X[j + 1:, :]
        alpha = X[j, j]
        beta, tau, scale = house(alpha, s[j])
        piv = X[j, :].copy()
        # T column j: u_q = V[j][q] + scale * s_q (q < j);  T[:j, j] = -tau T[:j,:j] u
        if j > 0:
            u = piv[:j] + scale * s[:j]
            T[:j, j] = -tau * (T[:j, :j] @ u)
        T[j, j] = tau
        w = piv + scale * s          # w_c = v' P[:, c], valid for c > j
        # update
        vcol = X[j + 1:, j] * scale
        X[j, j + 1:] -= tau * w[j + 1:]
        X[j + 1:, j + 1:] -= tau * np.outer(vcol, w[j + 1:])
        X[j + 1:, j] = vcol
        X[j, j] = beta
    for j in range(nref):
        V[j, j] = 1.0
        V[j + 1:, j] = X[j + 1:, j]
    return V, T, nref


def stage1(A: np.ndarray, z: np.ndarray, b: int):
    """Full symmetric A (m x m) -> band with b sub-diagonals, in place (both triangles maintained); z (m x L) <- Q'z."""
    m = A.shape[0]
    j0 = 0
    while m - j0 - b >= 2:
        r = m - j0 - b
        P = A[j0 + b:, j0:j0 + b]
        V, T, nref = panel_qr(P)
        # what the band keeps of the panel: R (rows i <= c); below is zero
        for c in range(b):
            P[c + 1:, c] = 0.0
        A[j0:j0 + b, j0 + b:] = P.T
        A22 = A[j0 + b:, j0 + b:]
        Z0 = A22 @ V                                   # k_sbr_av
        G0 = V.T @ Z0                                  # k_sbr_vtz (partials per CTA, fixed order)
        gz = V.T @ z[j0 + b:, :]
        S = T.T @ G0 @ T                               # k_sbr_w prologue
        W = Z0 @ T - 0.5 * V @ S
        z[j0 + b:, :] -= V @ (T.T @ gz)
        A22 -= V @ W.T + W @ V.T                       # k_sbr_r2k
        j0 += b
    return A


def to_band(A: np.ndarray, b: int):
    """B[off, j] = A[j + off, j], off = 0 .. 2b-1 (room for the bulge)."""
    m = A.shape[0]
    B = np.zeros((2 * b, m + 2 * b))
    for off in range(min(b, m - 1) + 1):
        B[off, :m - off] = np.diagonal(A, -off)
    return B


class BandView:
    """element access to the lower band storage as a symmetric matrix"""

    def __init__(self, B, m):
        self.B, self.m = B, m

    def blk(self, r0, nr, c0, nc):
        out = np.zeros((nr, nc))
        for p in range(nr):
            for q in range(nc):
                i, j = r0 + p, c0 + q
                if i >= j:
                    out[p, q] = self.B[i - j, j]
                else:
                    out[p, q] = self.B[j - i, i]
        return out

    def put(self, r0, c0, M, lower_only=False):
        nr, nc = M.shape
        for p in range(nr):
            for q in range(nc):
                i, j = r0 + p, c0 + q
                if i >= j:
                    self.B[i - j, j] = M[p, q]
                elif not lower_only:
                    raise AssertionError("upper write")


def chase_steps(m: int, b: int, s: int) -> int:
    """number of steps of sweep s (step 0 = type 1, step k >= 1 = type 2 + 3 on row block k)."""
    if min(b, m - 1 - s) < 2:
        return 0
    n = 1
    st = s + 1
    while True:
        lp = min(b, m - st)
        j1 = st + lp
        if j1 >= m:
            break
        n += 1
        st = j1
    return n


def chase_step(Bv: BandView, z: np.ndarray, b: int, s: int, k: int, state: dict):
    """One step of sweep s.  state carries the reflector (v, tau) from the previous step of the same sweep."""
    m = Bv.m
    if k == 0:
        ln = min(b, m - 1 - s)
        x = Bv.blk(s + 1, ln, s, 1)[:, 0]
        beta, tau, scale = house(x[0], float(x[1:] @ x[1:]))
        v = np.concatenate([[1.0], x[1:] * scale])
        if tau == 0.0:
            v[1:] = 0.0
        x[:] = 0.0
        x[0] = beta
        Bv.put(s + 1, s, x[:, None])
        r0 = s + 1
    else:
        st = s + 1 + (k - 1) * b
        lp = min(b, m - st)
        j1 = st + lp
        ln = min(b, m - j1)
        vp, taup = state["v"], state["tau"]
        Bk = Bv.blk(j1, ln, st, lp)
        u = Bk @ vp
        Bk -= taup * np.outer(u, vp)
        x = Bk[:, 0].copy()
        beta, tau, scale = house(x[0], float(x[1:] @ x[1:]))
        v = np.concatenate([[1.0], x[1:] * scale])
        if tau == 0.0:
            v[1:] = 0.0
        yv = v @ Bk[:, 1:]
        Bk[:, 1:] -= tau * np.outer(v, yv)
        Bk[:, 0] = 0.0
        Bk[0, 0] = beta
        Bv.put(j1, st, Bk)
        r0 = j1
    # two-sided on the diagonal block
    D = Bv.blk(r0, ln, r0, ln)
    p = tau * (D @ v)
    alpha = -0.5 * tau * float(p @ v)
    w = p + alpha * v
    D -= np.outer(v, w) + np.outer(w, v)
    _put_lower(Bv, r0, D)
    zr = z[r0:r0 + ln, :]
    z[r0:r0 + ln, :] = zr - tau * np.outer(v, v @ zr)
    state["v"], state["tau"] = v, tau


def _put_lower(Bv, r0, D):
    n = D.shape[0]
    for p in range(n):
        for q in range(p + 1):
            Bv.B[p - q, r0 + q] = D[p, q]


def stage2(B: np.ndarray, z: np.ndarray, m: int, b: int, rng=None):
    """Bulge chasing on the band storage.  rng = None: sweeps in order; otherwise a random schedule that only honours
    'sweep s step k after sweep s-1 step k+1' - the rule the persistent kernel spins on."""
    Bv = BandView(B, m)
    nsweep = max(m - 2, 0)
    total = [chase_steps(m, b, s) for s in range(nsweep)]
    if rng is None:
        for s in range(nsweep):
            st = {}
            for k in range(total[s]):
                chase_step(Bv, z, b, s, k, st)
    else:
        prog = [0] * nsweep
        states = [dict() for _ in range(nsweep)]
        live = [s for s in range(nsweep) if total[s] > 0]
        while live:
            ready = []
            for s in live:
                k = prog[s]
                if s == 0 or prog[s - 1] >= min(k + 2, total[s - 1]):
                    ready.append(s)
            s = ready[rng.integers(len(ready))]
            chase_step(Bv, z, b, s, prog[s], states[s])
            prog[s] += 1
            if prog[s] == total[s]:
                live.remove(s)
    d = B[0, :m].copy()
    e = B[1, :m - 1].copy()
    return d, e


def tri_rss(d, e, zh, lam):
    from scipy.linalg import solveh_banded
    ab = np.zeros((2, len(d)))
    ab[0] = d + lam
    ab[1, :-1] = e
    x = solveh_banded(ab, zh, lower=True)
    return lam * lam * float(x @ x)


def check(m, b, L=2, seed=0, random_schedule=False):
    rng = np.random.default_rng(seed)
    # PSD with a fast-decaying spectrum, like Q2'KQ2
    Qr, _ = np.linalg.qr(rng.standard_normal((m, m)))
    ev = 6.0 * np.exp(-np.linspace(0, 20, m))
    M = (Qr * ev) @ Qr.T
    M = 0.5 * (M + M.T)
    z = rng.standard_normal((m, L))
    A = M.copy()
    zh = z.copy()
    stage1(A, zh, b)
    # band check
    off = np.abs(np.tril(A, -(b + 1))).max() if m > b + 1 else 0.0
    ev1 = np.linalg.eigvalsh(np.tril(A) + np.tril(A, -1).T)
    B = to_band(A, b)
    d, e = stage2(B, zh, m, b, rng if random_schedule else None)
    from scipy.linalg import eigvalsh_tridiagonal
    ev2 = eigvalsh_tridiagonal(d, e)
    ref = np.linalg.eigvalsh(M)
    lam = 3e-3
    rss_ref = lam * lam * float(np.sum(np.linalg.solve(M + lam * np.eye(m), z[:, 0]) ** 2))
    rss = tri_rss(d, e, zh[:, 0], lam)
    return dict(m=m, b=b, below_band=off, ev_band=np.abs(ev1 - ref).max(), ev_tri=np.abs(ev2 - ref).max(),
                rss_rel=abs(rss - rss_ref) / rss_ref, znorm=abs(np.linalg.norm(zh) - np.linalg.norm(z)))


# ---------------------------------------------------------------------------------------------------------------------------
# Coefficients at the selected lambda from the BAND form instead of a dense Cholesky of M + lambda I (DESIGN.md section 9,
# not implemented on the device yet):  M = Q1 B Q1' after stage 1, so (M + lambda I)^-1 z = Q1 (B + lambda I)^-1 Q1'z.
# Q1'z is what stage 1 leaves in z; B + lambda I is banded (b sub-diagonals): block Cholesky with 32 x 32 blocks, O(m b^2);
# Q1 = H_1 ... H_K is applied panel by panel, last panel first, from the stored compact-WY factors (V_k, T_k).
# ---------------------------------------------------------------------------------------------------------------------------
def stage1_keep(A: np.ndarray, z: np.ndarray, b: int):
    """stage1() that also returns the compact-WY factors of every panel: list of (row offset, V, T)."""
    m = A.shape[0]
    j0 = 0
    panels = []
    while m - j0 - b >= 2:
        P = A[j0 + b:, j0:j0 + b]
        V, T, nref = panel_qr(P)
        for c in range(b):
            P[c + 1:, c] = 0.0
        A[j0:j0 + b, j0 + b:] = P.T
        A22 = A[j0 + b:, j0 + b:]
        Z0 = A22 @ V
        S = T.T @ (V.T @ Z0) @ T
        W = Z0 @ T - 0.5 * V @ S
        z[j0 + b:, :] -= V @ (T.T @ (V.T @ z[j0 + b:, :]))
        A22 -= V @ W.T + W @ V.T
        panels.append((j0 + b, V.copy(), T.copy()))
        j0 += b
    return panels


def band_cholesky_solve(B: np.ndarray, m: int, b: int, lam: float, rhs: np.ndarray) -> np.ndarray:
    """Solves (Bmat + lam I) x = rhs for the symmetric band matrix in lower band storage B[off, j] (off <= b), by a block
    Cholesky with b x b blocks: diagonal block potrf, one sub-diagonal block trsm, one trailing syrk per block column -
    what one persistent CTA would do, 156 steps at m = 5000."""
    nb = (m + b - 1) // b

    def blk(I, J):            # dense b x b block (I >= J) of the band matrix, zero outside the band / the matrix
        out = np.zeros((b, b))
        for p_ in range(b):
            for q_ in range(b):
                i, j = I * b + p_, J * b + q_
                if i < m and j < m and 0 <= i - j <= b:
                    out[p_, q_] = B[i - j, j]
                elif i < m and j < m and 0 < j - i <= b and I == J:
                    out[p_, q_] = B[j - i, i]
        if I == J:
            for p_ in range(b):
                i = I * b + p_
                out[p_, p_] = out[p_, p_] + lam if i < m else 1.0      # pad the last block with the identity
        return out

    Ld, Ls = [], []            # diagonal factors L_jj, sub-diagonal blocks L_{j+1,j}
    D = blk(0, 0)
    for j in range(nb):
        Ljj = np.linalg.cholesky(D)
        Ld.append(Ljj)
        if j + 1 < nb:
            S_ = blk(j + 1, j)
            Lsj = np.linalg.solve(Ljj, S_.T).T          # L_{j+1,j} = A_{j+1,j} L_jj^-T
            Ls.append(Lsj)
            D = blk(j + 1, j + 1) - Lsj @ Lsj.T
    x = np.zeros((nb * b, rhs.shape[1]))
    x[:m] = rhs
    for j in range(nb):                                  # forward
        x[j * b:(j + 1) * b] = np.linalg.solve(Ld[j], x[j * b:(j + 1) * b])
        if j + 1 < nb:
            x[(j + 1) * b:(j + 2) * b] -= Ls[j] @ x[j * b:(j + 1) * b]
    for j in range(nb - 1, -1, -1):                      # backward
        if j + 1 < nb:
            x[j * b:(j + 1) * b] -= Ls[j].T @ x[(j + 1) * b:(j + 2) * b]
        x[j * b:(j + 1) * b] = np.linalg.solve(Ld[j].T, x[j * b:(j + 1) * b])
    return x[:m]


def coefficients_from_band(M: np.ndarray, z: np.ndarray, lam: float, b: int) -> np.ndarray:
    A = M.copy()
    z1 = z.copy()
    panels = stage1_keep(A, z1, b)
    B = to_band(A, b)
    y = band_cholesky_solve(B, M.shape[0], b, lam, z1)
    for (r0, V, T) in reversed(panels):                  # beta = Q1 y,  H_k = I - V T V'
        y[r0:] -= V @ (T @ (V.T @ y[r0:]))
    return y


# ---------------------------------------------------------------------------------------------------------------------------
# Stage 2 with g consecutive sweeps per CTA and the band rows they work on kept in a sliding SHARED-MEMORY WINDOW (DESIGN.md
# section 9, option (c); not implemented on the device yet).  What this prototype pins down is the hand-over protocol:
#   * the window is organised by matrix ROWS (row i holds B[i - j, j] for its 64 band columns; a task of sweep s, step k only
#     touches rows of its row block R_k(s) = [s + 1 + k b, s + 1 + (k + 1) b));
#   * group-step K runs the tasks (s0 + i, K - 2 i), i < g - they are two steps apart, hence independent;
#   * before group-step K the group LOADS the rows below  s0 + 1 + (K + 1) b  that are not resident yet, which needs the previous
#     group to have RETIRED them;  after it the group RETIRES (writes back, publishes) every row none of its sweeps will touch again:
#     rows below  min_i (s0 + i + 1 + (K - 2 i + 1) b)  over its unfinished sweeps.
# Finality by rows needs no special cases (the by-column view has one element per step that changes hands early).
# ---------------------------------------------------------------------------------------------------------------------------
class _RowWindow:
    """BandView on top of a cache of rows; every access asserts that the row is resident."""

    def __init__(self, m, width):
        self.m, self.width, self.rows = m, width, {}

    def _get(self, i, j):
        if i < j:
            i, j = j, i
        assert i in self.rows, ("row not resident", i)
        return self.rows[i][i - j]

    def _set(self, i, j, v):
        assert i >= j and i in self.rows, ("row not resident", i)
        self.rows[i][i - j] = v

    def blk(self, r0, nr, c0, nc):
        return np.array([[self._get(r0 + p_, c0 + q_) for q_ in range(nc)] for p_ in range(nr)])

    def put(self, r0, c0, M, lower_only=False):
        for p_ in range(M.shape[0]):
            for q_ in range(M.shape[1]):
                if r0 + p_ >= c0 + q_:
                    self._set(r0 + p_, c0 + q_, M[p_, q_])

    @property
    def B(self):
        return self

    def __setitem__(self, key, v):          # _put_lower writes Bv.B[off, col]
        off, col = key
        self._set(col + off, col, v)


def stage2_grouped(B: np.ndarray, z: np.ndarray, m: int, b: int, g: int, rng):
    """Bulge chasing with groups of g sweeps, each group working on its own row window; groups are interleaved at random
    subject to the load rule above.  Returns (d, e) like stage2()."""
    width = 2 * b
    nsweep = max(m - 2, 0)
    total = [chase_steps(m, b, s) for s in range(nsweep)]
    ngroup = (nsweep + g - 1) // g

    class Group:
        pass

    groups = []
    for ga in range(ngroup):
        G = Group()
        G.s0 = ga * g
        G.sweeps = list(range(G.s0, min(G.s0 + g, nsweep)))
        G.K = 0
        G.win = _RowWindow(m, width)
        G.zrows = {}
        G.loaded = G.s0 + 1                 # rows [s0 + 1, loaded) are resident or already retired
        G.retired = G.s0 + 1                # rows below are final for this group (and written back)
        G.states = [dict() for _ in G.sweeps]
        G.nsteps = max([total[s] + 2 * i for i, s in enumerate(G.sweeps)] + [0])
        G.done = G.nsteps == 0
        if G.done:
            G.retired = m
        groups.append(G)

    def row_from_global(i):
        return np.array([B[off, i - off] if i - off >= 0 else 0.0 for off in range(width)])

    def row_to_global(i, row):
        for off in range(width):
            if i - off >= 0:
                B[off, i - off] = row[off]

    def can_run(ga):
        G = groups[ga]
        if G.done:
            return False
        need = min(m, G.s0 + 1 + (G.K + 1) * b)
        return ga == 0 or groups[ga - 1].retired >= need

    def run(ga):
        G = groups[ga]
        need = min(m, G.s0 + 1 + (G.K + 1) * b)
        for i in range(G.loaded, need):                       # load
            G.win.rows[i] = row_from_global(i)
            G.zrows[i] = z[i].copy()
        G.loaded = max(G.loaded, need)
        zview = _ZRows(G.zrows, z.shape[1])
        for idx, s in enumerate(G.sweeps):                    # the g independent tasks of this group-step
            k = G.K - 2 * idx
            if 0 <= k < total[s]:
                chase_step(G.win, zview, b, s, k, G.states[idx])
        G.K += 1
        bound = m
        for idx, s in enumerate(G.sweeps):                    # retire
            k_next = G.K - 2 * idx
            if k_next >= total[s]:
                continue                                      # this sweep has finished
            bound = min(bound, s + 1 + max(k_next, 0) * b)
        if G.K >= G.nsteps:
            bound = m
            G.done = True
        for i in range(G.retired, min(bound, G.loaded)):
            row_to_global(i, G.win.rows.pop(i))
            z[i] = G.zrows.pop(i)
        G.retired = max(G.retired, min(bound, G.loaded)) if not G.done else m
        G.max_resident = max(getattr(G, "max_resident", 0), len(G.win.rows))

    while True:
        ready = [ga for ga in range(ngroup) if can_run(ga)]
        if not ready:
            break
        run(ready[rng.integers(len(ready))])
    assert all(G.done for G in groups), "schedule stalled"
    stage2_grouped.max_resident = max(getattr(G, "max_resident", 0) for G in groups) if groups else 0
    return B[0, :m].copy(), B[1, :m - 1].copy()


class _ZRows:
    """z[r0:r0+ln, :] views on the cached right-hand-side rows"""

    def __init__(self, rows, L):
        self.rows, self.L = rows, L

    def __getitem__(self, key):
        sl, _ = key
        return np.array([self.rows[i] for i in range(sl.start, sl.stop)])

    def __setitem__(self, key, v):
        sl, _ = key
        for n_, i in enumerate(range(sl.start, sl.stop)):
            self.rows[i] = np.array(v[n_])


if __name__ == "__main__":
    for (m, b) in [(37, 8), (64, 8), (65, 8), (100, 32), (131, 32), (200, 32), (3, 8), (10, 8), (34, 32), (35, 32)]:
        print(check(m, b))
    print("random schedule", check(90, 8, random_schedule=True, seed=3), check(150, 32, random_schedule=True, seed=4))
