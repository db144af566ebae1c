"""CPU prototype (numpy/scipy) of the shifted-Laplacian multigrid preconditioner used to pick
the algorithm that the CUDA kernels implement.  Not product code, not imported anywhere."""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, "/root/repo")
from oracle import scattering_oracle as so, yee_oracle


def op_matrix(pol, ca, cb, cd, sx, sy):
    Ny, Nx = ca.shape
    DEX, DEY, DHX, DHY = yee_oracle.yeeder2d([Nx, Ny], [1.0, 1.0], [0, 0])
    a = sp.diags(ca.ravel()); b = sp.diags(cb.ravel()); d = sp.diags(cd.ravel())
    if pol == "TE":
        return (sx * (DHX @ a @ DEX) + sy * (DHY @ b @ DEY) + d).tocsr()
    return (sx * (DEX @ a @ DHX) + sy * (DEY @ b @ DHY) + d).tocsr()


def coarsen_coefs(pol, ca, cb, cd, dwall=None):
    """dwall: distance (in units of the *fine* spacing) of the wall-adjacent fine cell centre from the
    Dirichlet wall (1 on the finest level).  The coarse wall-face coefficient is scaled by H/d_coarse."""
    Ny, Nx = ca.shape
    assert Nx % 2 == 0 and Ny % 2 == 0
    # TE: ca_i lives on face i+1/2 ; coarse face I+1/2 = fine face (2I+1)+1/2 -> fine ca index 2I+1
    # TM: ca_i lives on face i-1/2 ; coarse face I-1/2 = fine face 2I-1/2     -> fine ca index 2I
    ix = 1 if pol == "TE" else 0
    cac = 0.5 * (ca[0::2, ix::2] + ca[1::2, ix::2])
    cbc = 0.5 * (cb[ix::2, 0::2] + cb[ix::2, 1::2])
    cdc = 0.25 * (cd[0::2, 0::2] + cd[1::2, 0::2] + cd[0::2, 1::2] + cd[1::2, 1::2])
    if dwall is not None:
        m = dwall / ((dwall + 0.5) / 2.0)     # delta_l / delta_{l+1}
        if pol == "TE":
            cac[:, -1] *= m; cbc[-1, :] *= m
        else:
            cac[:, 0] *= m; cbc[0, :] *= m
    return cac, cbc, cdc


def restrict(r):
    return 0.25 * (r[0::2, 0::2] + r[1::2, 0::2] + r[0::2, 1::2] + r[1::2, 1::2])


def prolong_const(e):
    return np.repeat(np.repeat(e, 2, axis=0), 2, axis=1)


def prolong_bilinear_wall(e, pol, dwall):
    """bilinear cell-centred interpolation; Neumann side clamps, Dirichlet side interpolates towards 0 at the wall.
    dwall = delta of the FINE level (relative distance of wall cell centre to wall)."""
    dc = (dwall + 0.5) / 2.0            # coarse delta (units of coarse spacing H=2h) -> distance 2*dc (in h)
    wq = dwall / (2.0 * dc)             # fine outer cell value = wq * E_last  (linear to 0 at the wall)
    def up1(a, axis):
        a = np.moveaxis(a, axis, 0)
        lo = np.concatenate([a[:1], a[:-1]], axis=0)
        hi = np.concatenate([a[1:], a[-1:]], axis=0)
        out = np.empty((2 * a.shape[0],) + a.shape[1:], dtype=a.dtype)
        out[0::2] = 0.75 * a + 0.25 * lo
        out[1::2] = 0.75 * a + 0.25 * hi
        if pol == "TE":
            out[-1] = wq * a[-1]
        else:
            out[0] = wq * a[0]
        return np.moveaxis(out, 0, axis)
    return up1(up1(e, 0), 1)


def prolong_bilinear(e):
    # cell-centred bilinear: fine cell (2I) = 3/4 e_I + 1/4 e_{I-1}, (2I+1) = 3/4 e_I + 1/4 e_{I+1}; edges: clamp
    def up1(a, axis):
        a = np.moveaxis(a, axis, 0)
        lo = np.concatenate([a[:1], a[:-1]], axis=0)
        hi = np.concatenate([a[1:], a[-1:]], axis=0)
        out = np.empty((2 * a.shape[0],) + a.shape[1:], dtype=a.dtype)
        out[0::2] = 0.75 * a + 0.25 * lo
        out[1::2] = 0.75 * a + 0.25 * hi
        return np.moveaxis(out, 0, axis)
    return up1(up1(e, 0), 1)


class MG:
    def __init__(self, pol, ca, cb, cd, sx, sy, shift=(1.0, 0.5), nlev=None, omega=0.8, pre=1, post=1,
                 prolong="bilinear", coarse="direct", cycle="V", coarse_its=30, min_n=32, wall=True, growth=1, wfrom=0):
        self.growth = growth; self.wfrom = wfrom
        self.pol = pol; self.omega = omega; self.pre = pre; self.post = post
        self.P = prolong_bilinear if prolong == "bilinear" else prolong_const
        self.cycle = cycle; self.coarse = coarse; self.coarse_its = coarse_its
        self.levels = []
        sh = shift[0] + 1j * shift[1]
        cdm = sh * cd
        self.prolong = prolong; self.wall = wall
        dwall = 1.0
        while True:
            A = op_matrix(pol, ca, cb, cdm, sx, sy)
            self.levels.append(dict(A=A, dinv=1.0 / A.diagonal(), shape=ca.shape, dwall=dwall))
            Ny, Nx = ca.shape
            if (nlev and len(self.levels) >= nlev) or Nx % 2 or Ny % 2 or min(Nx, Ny) // 2 < min_n:
                break
            ca, cb, cdm = coarsen_coefs(pol, ca, cb, cdm, dwall if wall else None)
            dwall = (dwall + 0.5) / 2.0
            sx, sy = sx / 4, sy / 4
        if coarse == "direct":
            self.lu = spla.splu(self.levels[-1]["A"].tocsc())
        self.work = 0.0

    def smooth(self, l, x, b, n):
        L = self.levels[l]
        for _ in range(n):
            x = x + self.omega * L["dinv"] * (b - L["A"] @ x)
            self.work += 0.25 ** l
        return x

    def vcycle(self, l, b):
        L = self.levels[l]
        if l == len(self.levels) - 1:
            if self.coarse == "direct":
                return self.lu.solve(b)
            return self.smooth(l, np.zeros_like(b), b, self.coarse_its)
        g = int(round(self.growth ** l))
        x = self.smooth(l, np.zeros_like(b), b, self.pre * g)
        r = b - L["A"] @ x
        self.work += 0.25 ** l
        rc = restrict(r.reshape(L["shape"])).ravel()
        ec = self.vcycle(l + 1, rc)
        if self.cycle == "W" and l + 1 >= self.wfrom:
            Lc = self.levels[l + 1]
            if l + 1 < len(self.levels) - 1:
                rc2 = rc - Lc["A"] @ ec
                ec = ec + self.vcycle(l + 1, rc2)
        ecs = ec.reshape(self.levels[l + 1]["shape"])
        if self.prolong == "bilinear" and self.wall:
            x = x + prolong_bilinear_wall(ecs, self.pol, L["dwall"]).ravel()
        else:
            x = x + self.P(ecs).ravel()
        x = self.smooth(l, x, b, self.post * g)
        return x

    def __call__(self, b):
        return self.vcycle(0, b)


def fgmres(A, b, M, rtol=1e-8, restart=30, maxit=500, verbose=False):
    n = b.size
    x = np.zeros(n, complex)
    bn = np.linalg.norm(b)
    its = 0
    hist = []
    while its < maxit:
        r = b - A @ x
        beta = np.linalg.norm(r)
        hist.append(beta / bn)
        if beta / bn < rtol:
            break
        V = np.zeros((restart + 1, n), complex); Z = np.zeros((restart, n), complex)
        H = np.zeros((restart + 1, restart), complex)
        V[0] = r / beta
        g = np.zeros(restart + 1, complex); g[0] = beta
        cs = np.zeros(restart, complex); sn = np.zeros(restart, complex)
        k = 0
        for j in range(restart):
            Z[j] = M(V[j])
            w = A @ Z[j]
            for _ in range(2):
                h = V[:j + 1].conj() @ w
                w = w - V[:j + 1].T @ h
                H[:j + 1, j] += h
            H[j + 1, j] = np.linalg.norm(w)
            V[j + 1] = w / H[j + 1, j]
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -np.conj(sn[i]) * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            a_, b_ = H[j, j], H[j + 1, j]
            d = np.sqrt(abs(a_) ** 2 + abs(b_) ** 2)
            cs[j] = abs(a_) / d if a_ != 0 else 0
            sn[j] = (a_ / abs(a_)) * np.conj(b_) / d if a_ != 0 else 1
            H[j, j] = cs[j] * a_ + sn[j] * b_
            H[j + 1, j] = 0
            g[j + 1] = -np.conj(sn[j]) * g[j]
            g[j] = cs[j] * g[j]
            its += 1; k = j + 1
            hist.append(abs(g[j + 1]) / bn)
            if verbose and its % 10 == 0:
                print("   it", its, hist[-1])
            if abs(g[j + 1]) / bn < rtol or its >= maxit:
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
        x = x + Z[:k].T @ y
    r = b - A @ x
    return x, its, np.linalg.norm(r) / bn, hist


def bicgstab(A, b, M, rtol=1e-8, maxit=2000):
    n = b.size; x = np.zeros(n, complex); r = b.copy(); rh = r.copy()
    rho = alpha = om = 1.0; v = np.zeros(n, complex); p = np.zeros(n, complex)
    bn = np.linalg.norm(b); hist = []
    for it in range(1, maxit + 1):
        rho_n = np.vdot(rh, r); beta = (rho_n / rho) * (alpha / om); rho = rho_n
        p = r + beta * (p - om * v)
        ph = M(p); v = A @ ph
        alpha = rho / np.vdot(rh, v)
        s = r - alpha * v
        sh = M(s); t = A @ sh
        om = np.vdot(t, s) / np.vdot(t, t)
        x += alpha * ph + om * sh
        r = s - om * t
        hist.append(np.linalg.norm(r) / bn)
        if hist[-1] < rtol:
            break
    return x, it, np.linalg.norm(b - A @ x) / bn, hist


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    pol = sys.argv[2] if len(sys.argv) > 2 else "TM"
    eps = float(sys.argv[3]) if len(sys.argv) > 3 else -1e6
    p = so.cylinder_problem(N, eps_r=eps)
    ca, cb, cd, sx, sy = so.stencil_coefficients(p, pol)
    A = op_matrix(pol, ca, cb, cd, sx, sy)
    Aref = so.build_A(p, pol)
    print("operator check", abs(A - Aref).max())
    b = so.rhs(Aref, p)
    for kw in (dict(shift=(1, 0.5)), dict(shift=(1, 0.5), prolong="const"), dict(shift=(1, 0.3)), dict(shift=(1, 0.5), pre=2, post=2),
               dict(shift=(1, 0.5), coarse="jacobi", coarse_its=40), dict(shift=(1,0.5), omega=0.6)):
        t = time.time()
        M = MG(pol, ca, cb, cd, sx, sy, **kw)
        x, its, rr, hist = fgmres(A, b, M, restart=40, maxit=400)
        print(N, pol, kw, "levels", len(M.levels), "its", its, "relres %.2e" % rr, "work(fine-applies) %.0f" % M.work, "t %.1f" % (time.time() - t))
