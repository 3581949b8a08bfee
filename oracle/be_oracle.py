"""CPU oracle for the Blurry-Edges render -> fold -> depth hot path (TEST INFRASTRUCTURE ONLY).

This module is a from-scratch torch restatement of the arithmetic the reference performs in
its eager PyTorch code.  It exists to CHECK the CUDA path; nothing in the product package
(`blurry_edges_b200/`) may import it.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` use it.

Parity pin: `tests/golden/make_golden.py` runs the UNMODIFIED reference classes (imported
from /root/reference in the build container) on seeded inputs and stores their outputs in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` holds this restatement to those vectors
(fp64: <=1e-10; fp32 reference within its own documented noise floor).

Layout convention used here (differs from the reference on purpose):
patch-major tensors `[N, ..., R, R]` with N = B*Hp*Wp and patch index l = py*Wp + px.
Every function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

DELTA = 0.07  # utils/postprocessing_loss.py:97 (normalized_gaussian default)


# --------------------------------------------------------------------------------------
# constants
# --------------------------------------------------------------------------------------
@dataclass
class Geometry:
    """Patch geometry, utils/postprocessing_loss.py:8-20,131-143 and utils/args.py:9-15."""
    R: int = 21
    stride: int = 2
    H: int = 147
    W: int = 147
    w: float = 1.0
    alpha_lambda: float = 5e-3

    @property
    def Hp(self) -> int:
        return (self.H - self.R) // self.stride + 1

    @property
    def Wp(self) -> int:
        return (self.W - self.R) // self.stride + 1

    @property
    def L(self) -> int:
        return self.Hp * self.Wp

    @property
    def lam(self) -> float:
        # :14; `ridge = lambda_ridge * torch.eye(3)` (:122,:133) is an fp32 tensor, so the value the
        # reference adds to the diagonal is the fp32 rounding of lambda even in its fp64 mode.
        return float(torch.tensor((self.alpha_lambda * self.R ** 2) ** 2, dtype=torch.float32))


@dataclass
class Camera:
    """utils/depth_etas.py:4-21.  intercept/theta_* are fp32-rounded there (torch.tensor of a
    python float) and stay fp32-rounded when the harness casts them to fp64."""
    s: float = 0.1104
    rho_1: float = 10.0
    rho_2: float = 10.2
    sigma_cam: float = 0.003
    pixel_pitch: float = 5.86e-6
    mag: float = 4.0
    R: int = 21
    numerator: float = field(init=False)
    k_const: float = field(init=False)
    k_root: float = field(init=False)
    k_fac: float = field(init=False)
    intercept: float = field(init=False)
    theta_mid: float = field(init=False)
    theta_wng: float = field(init=False)

    def __post_init__(self):
        nf = self.R // 2
        self.numerator = 2 * self.s ** 2 * (self.rho_2 - self.rho_1)
        self.k_const = -self.s * (self.rho_1 - self.rho_2) * (self.rho_1 * self.s + self.rho_2 * self.s - 2)
        self.k_root = nf * self.pixel_pitch * self.mag / self.sigma_cam
        self.k_fac = self.k_root ** 2
        f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))
        # torch.abs(torch.tensor(s*(rho2-rho1))) * sigma_cam / pitch / mag / nf, evaluated in fp32
        t = torch.abs(torch.tensor(self.s * (self.rho_2 - self.rho_1), dtype=torch.float32))
        self.intercept = float(t * self.sigma_cam / self.pixel_pitch / self.mag / nf)
        self.theta_mid = f32(3 / 4 * math.pi)
        self.theta_wng = f32(1 / 4 * math.pi)


def pixel_axis(R: int, dtype) -> torch.Tensor:
    """fp32 linspace(-1,1,R) cast to the working dtype (postprocessing_loss.py:15-17)."""
    return torch.linspace(-1.0, 1.0, R, dtype=torch.float32).to(dtype)


# --------------------------------------------------------------------------------------
# per-patch geometry: params -> signed wedge distances          (postprocessing_loss.py:26-86)
# --------------------------------------------------------------------------------------
def _edge(px, py, ang, X, Y, w):
    """One half-line edge: perpendicular distance d, axial coordinate a, and the
    'rounded cap' distance D used behind the vertex (:26-30, :57-78)."""
    sn, cs = torch.sin(ang), torch.cos(ang)
    dx, dy = X - px, Y - py
    d = -sn * dx + cs * dy
    a = cs * dx + sn * dy
    sg = torch.where(d < 0, -torch.ones_like(d), torch.ones_like(d))
    cap = torch.sqrt(d ** 2 + (a * w) ** 2) * sg
    return torch.where(a < 0, cap, d)


def wedge_distances(geo: torch.Tensor, R: int, w: float = 1.0) -> torch.Tensor:
    """geo [N,8] = (x0,y0,x1,y1,theta1,phi1,theta2,phi2) -> dist [N,2,R,R] (:43-86)."""
    dt = geo.dtype
    ax = pixel_axis(R, dt).to(geo.device)
    Y = ax.view(1, R, 1)
    X = ax.view(1, 1, R)
    g = lambda k: geo[:, k].view(-1, 1, 1)
    out = []
    for k in range(2):
        vx, vy, th, ph = g(2 * k), g(2 * k + 1), g(4 + 2 * k), g(5 + 2 * k)
        flip = torch.where(torch.remainder(ph, 2 * math.pi) < math.pi, torch.ones_like(ph), -torch.ones_like(ph))
        DA = _edge(vx, vy, th, X, Y, w)
        DB = _edge(vx, vy, th + ph, X, Y, w)
        if k == 0:
            inside = (flip * DA > 0) & (flip * DB < 0)        # :80 strict
        else:
            inside = (flip * DA >= 0) & (flip * DB <= 0)      # :81 non-strict
        sign = flip * torch.where(inside, torch.ones_like(DA), -torch.ones_like(DA))
        out.append(torch.min(DA.abs(), DB.abs()) * sign)
    return torch.stack(out, dim=1)


def eta_from_coef(e: torch.Tensor) -> torch.Tensor:
    """:88-89"""
    return 10 ** (torch.erf(e) * 2 - 2)


def soft_indicators(dist: torch.Tensor, eta: torch.Tensor) -> torch.Tensor:
    """dist [N,2,R,R], eta [N,2] -> u [N,3,R,R] (:91-95).  sqrt(2) is fp32-rounded there."""
    root2 = torch.sqrt(torch.tensor(2)).to(dist.dtype)  # int tensor -> fp32 sqrt, as in the reference
    h = 0.5 * (1.0 + torch.erf(dist / (root2 * eta.view(-1, 2, 1, 1))))
    h1, h2 = h[:, 0], h[:, 1]
    return torch.stack([(1 - h1) * (1 - h2), h1 * (1 - h2), h2], dim=1)


def bump(x: torch.Tensor, delta: float = DELTA) -> torch.Tensor:
    """:97-98"""
    return torch.exp(-x ** 2 / delta ** 2)


def boundary_map(dist: torch.Tensor) -> torch.Tensor:
    """blurry_edges_test.py:59-61 / global_training.py:80-82 / local_training.py:42-44."""
    d1, d2 = dist[:, 0], dist[:, 1]
    dB = torch.where(d2 >= 0, d2, torch.where(d1.abs() < d2.abs(), d1.abs(), d2.abs()))
    return bump(dB)


def depth_mask(dist: torch.Tensor, densify: str | None) -> torch.Tensor:
    """int32 mask in {0,1,2}: blurry_edges_test.py:47-54, global_training.py:84-86."""
    d1, d2 = dist[:, 0], dist[:, 1]
    if densify == 'w':
        m = (d1 > 0).to(torch.int32)
        t = (d2 > 0).to(torch.int32) * 2
        return torch.where(t == 2, t, m)
    m = (bump(d1) > 0.5).to(torch.int32)
    t = (bump(d2) > 0.5).to(torch.int32) * 2
    return torch.where((t == 2) | (d2 >= 0), t, m)


# --------------------------------------------------------------------------------------
# ridge regression colours
# --------------------------------------------------------------------------------------
def inv3_sym(M: torch.Tensor) -> torch.Tensor:
    """Inverse of [...,3,3] by cofactors.  The reference uses the Cayley-Hamilton trace
    formula (postprocessing_loss.py:104-112); both are exact algebra, they differ only in
    rounding (the trace form loses ~3e-3 in fp32)."""
    a, b, c = M[..., 0, 0], M[..., 0, 1], M[..., 0, 2]
    d, e, f = M[..., 1, 0], M[..., 1, 1], M[..., 1, 2]
    g, h, i = M[..., 2, 0], M[..., 2, 1], M[..., 2, 2]
    A = e * i - f * h
    B = -(d * i - f * g)
    C = d * h - e * g
    det = a * A + b * B + c * C
    adj = torch.stack([
        torch.stack([A, -(b * i - c * h), b * f - c * e], -1),
        torch.stack([B, a * i - c * g, -(a * f - c * d)], -1),
        torch.stack([C, -(a * h - b * g), a * e - b * d], -1)], -2)
    return adj / det[..., None, None]


def inv3_trace(M: torch.Tensor) -> torch.Tensor:
    """The reference's own formula (postprocessing_loss.py:104-112,127-128) - used only to
    reproduce the reference's fp32 noise floor in tests."""
    tr = torch.diagonal(M, dim1=-2, dim2=-1).sum(-1)
    M2 = M @ M
    tr2 = torch.diagonal(M2, dim1=-2, dim2=-1).sum(-1)
    tr3 = torch.diagonal(M2 @ M, dim1=-2, dim2=-1).sum(-1)
    det = (tr ** 3 - 3 * tr * tr2 + 2 * tr3) / 6
    eye = torch.eye(3, dtype=M.dtype, device=M.device)
    adj = M2 - tr[..., None, None] * M + ((tr ** 2 - tr2) / 2)[..., None, None] * eye
    return adj / det[..., None, None]


def ridge_colors(U: torch.Tensor, Yv: torch.Tensor, lam: float, trace_form: bool = False) -> torch.Tensor:
    """U [N,K,3] wedge rows, Yv [N,K,3] pixel rows -> C [N,3(wedge),3(channel)].
    (A^T A + lam I)^-1 A^T y: blurry_edges_test.py:19-28, global_training.py:62-67,
    local_training.py:37-40, global_data_pre_cal.py:43-46."""
    Ut = U.transpose(1, 2)
    M = Ut @ U + lam * torch.eye(3, dtype=U.dtype, device=U.device)
    inv = inv3_trace(M) if trace_form else inv3_sym(M)
    return inv @ (Ut @ Yv)


# --------------------------------------------------------------------------------------
# depth from the two defocus etas                                  (utils/depth_etas.py:23-37)
# --------------------------------------------------------------------------------------
def depth_from_etas(cam: Camera, e1: torch.Tensor, e2: torch.Tensor) -> torch.Tensor:
    dt = e1.dtype
    c = torch.tensor(cam.intercept, dtype=dt)
    tw = torch.tensor(cam.theta_wng, dtype=dt)
    tm = torch.tensor(cam.theta_mid, dtype=dt)
    c1 = -torch.sin(tw) * e1 + torch.cos(tw) * (e2 - c)
    c2 = -torch.sin(tm) * (e1 - c) + torch.cos(tm) * e2
    c3 = -torch.sin(tw) * (e1 - c) + torch.cos(tw) * e2
    half = (e1 + e2 - c) / 2
    a = torch.where(c1 > 0, half, torch.where(c2 > 0, c + (e1 - e2 - c) / 2, torch.where(c3 < 0, c + half, e1)))
    b = torch.where(c1 > 0, c + half, torch.where(c2 > 0, (e2 - e1 + c) / 2, torch.where(c3 < 0, half, e2)))
    return cam.numerator / (cam.k_fac * (a ** 2 - b ** 2) + cam.k_const)


def refocus_sigma(cam: Camera, z: torch.Tensor, rho_prime: float) -> torch.Tensor:
    return torch.abs((1 / z - rho_prime) * cam.s + 1) / cam.k_root


# --------------------------------------------------------------------------------------
# patch extraction / fold
# --------------------------------------------------------------------------------------
def extract(img: torch.Tensor, R: int, stride: int) -> torch.Tensor:
    """img [M,C,H,W] -> patches [M*Hp*Wp, C, R, R] (nn.Unfold, patch index = py*Wp+px)."""
    M, C, H, W = img.shape
    cols = F.unfold(img, R, stride=stride)                       # [M, C*R*R, L]
    return cols.transpose(1, 2).reshape(-1, C, R, R)


def fold_sum(p: torch.Tensor, M: int, g: Geometry) -> torch.Tensor:
    """p [M*L, C, R, R] -> overlap SUM [M, C, H, W] (torch.nn.Fold, postprocessing_loss.py:151-173)."""
    C = p.shape[1]
    cols = p.reshape(M, g.L, C * g.R * g.R).transpose(1, 2)
    return F.fold(cols, [g.H, g.W], g.R, stride=g.stride)


def cover_count(g: Geometry, dtype=torch.float32) -> torch.Tensor:
    """num_patches (postprocessing_loss.py:139-143) in closed form: #patches covering each pixel."""
    def axis(n, npatch):
        y = torch.arange(n)
        lo = torch.clamp((y - g.R + g.stride) // g.stride, min=0)   # ceil((y-R+1)/s)
        hi = torch.clamp(y // g.stride, max=npatch - 1)
        return (hi - lo + 1).clamp(min=0)
    return (axis(g.H, g.Hp)[:, None] * axis(g.W, g.Wp)[None, :]).to(dtype)


def sobel_mag(img: torch.Tensor) -> torch.Tensor:
    """[...,H,W] -> [...,H-2,W-2], sqrt(Sx^2+Sy^2+1e-8), valid padding (postprocessing_loss.py:114-117)."""
    sh = img.shape
    x = img.reshape(-1, 1, sh[-2], sh[-1])
    kx = torch.tensor([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=img.dtype, device=img.device).view(1, 1, 3, 3)
    ky = torch.tensor([[1, 2, 1], [0, 0, 0], [-1, -2, -1]], dtype=img.dtype, device=img.device).view(1, 1, 3, 3)
    out = torch.sqrt(F.conv2d(x, kx) ** 2 + F.conv2d(x, ky) ** 2 + 1e-8)
    return out.reshape(*sh[:-2], sh[-2] - 2, sh[-1] - 2)


# --------------------------------------------------------------------------------------
# compositions
# --------------------------------------------------------------------------------------
def colors_only(est: torch.Tensor, img: torch.Tensor, g: Geometry, trace_form: bool = False) -> torch.Tensor:
    """Pass A.  est [M,L,10] (angles already wrapped, eta coefficient raw), img [M,3,H,W]
    -> colours [M,3(channel),3(wedge),Hp,Wp]   (blurry_edges_test.py:81-92 with colors_only=True)."""
    M = est.shape[0]
    p = est.reshape(-1, 10)
    dist = wedge_distances(p[:, :8], g.R, g.w)
    u = soft_indicators(dist, eta_from_coef(p[:, 8:10]))
    pix = extract(img, g.R, g.stride)                              # [N,3,R,R]
    C = ridge_colors(u.flatten(2).transpose(1, 2), pix.flatten(2).transpose(1, 2), g.lam, trace_form)
    return C.reshape(M, g.Hp, g.Wp, 3, 3).permute(0, 4, 3, 1, 2)   # [M, c, w, Hp, Wp]


def colors_local(est: torch.Tensor, pat: torch.Tensor, g: Geometry) -> torch.Tensor:
    """global_data_pre_cal.py:39-47.  est [N,10], pat [N,R,R,3] -> colours [N,3(channel),3(wedge)]."""
    dist = wedge_distances(est[:, :8], g.R, g.w)
    u = soft_indicators(dist, eta_from_coef(est[:, 8:10]))
    C = ridge_colors(u.flatten(2).transpose(1, 2), pat.flatten(1, 2), g.lam)
    return C.transpose(1, 2)


def render_pair(est: torch.Tensor, img: torch.Tensor, g: Geometry, cam: Camera, trace_form: bool = False):
    """Shared front end of pass B and of the global loss.
    est [B,L,12] RESTORED params (xy, wrapped angles, eta COEFFICIENTS), img [B,2,3,H,W].
    Returns dict of patch-major tensors."""
    B = est.shape[0]
    p = est.reshape(-1, 12)
    dist = wedge_distances(p[:, :8], g.R, g.w)                    # [N,2,R,R]
    eta = eta_from_coef(p[:, 8:12])                               # [N,4] (w1i1, w2i1, w1i2, w2i2)
    u1 = soft_indicators(dist, eta[:, 0:2])
    u2 = soft_indicators(dist, eta[:, 2:4])
    pix = extract(img.reshape(B * 2, 3, g.H, g.W), g.R, g.stride).reshape(B, 2, g.L, 3, g.R, g.R)
    pix = pix.permute(0, 2, 1, 3, 4, 5).reshape(-1, 2, 3, g.R, g.R)  # [N,2,3,R,R]
    U = torch.cat([u1.flatten(2), u2.flatten(2)], dim=2).transpose(1, 2)       # [N,2RR,3]
    Yv = torch.cat([pix[:, 0].flatten(2), pix[:, 1].flatten(2)], dim=2).transpose(1, 2)
    C = ridge_colors(U, Yv, g.lam, trace_form)                    # [N,3w,3c]
    paint = lambda u: torch.einsum('nwij,nwc->ncij', u, C)
    z1 = depth_from_etas(cam, eta[:, 0], eta[:, 2])
    z2 = depth_from_etas(cam, eta[:, 1], eta[:, 3])
    return dict(dist=dist, eta=eta, u1=u1, u2=u2, C=C, P1=paint(u1), P2=paint(u2), z1=z1, z2=z2,
                lb=boundary_map(dist), paint=paint, pix=pix)


def inference(est: torch.Tensor, img: torch.Tensor, g: Geometry, cam: Camera, rho_prime: float = 10.39,
              densify: str | None = None, trace_form: bool = False, return_patches: bool = False):
    """Pass B (blurry_edges_test.py:30-100).  est [B,L,12] restored, img [B,2,3,H,W].
    Returns (image [B,2,3,H,W], sharpened [B,3,H,W], refocused [B,3,H,W], boundary [B,1,H,W],
    depth [B,H,W], confidence [B,H,W])."""
    B = est.shape[0]
    r = render_pair(est, img, g, cam, trace_form)
    dist, dt = r['dist'], est.dtype
    m = depth_mask(dist, densify)
    zero = torch.zeros((), dtype=dt)
    dmap = torch.where(m == 1, r['z1'].view(-1, 1, 1), torch.where(m == 2, r['z2'].view(-1, 1, 1), zero))
    sharp = r['paint'](soft_indicators(dist, torch.full_like(r['eta'][:, :2], 1e-4)))
    s1 = torch.where((m == 1).flatten(1).sum(1) > 0, refocus_sigma(cam, r['z1'], rho_prime), torch.full_like(r['z1'], 1e-4))
    s2 = torch.where((m == 2).flatten(1).sum(1) > 0, refocus_sigma(cam, r['z2'], rho_prime), torch.full_like(r['z2'], 1e-4))
    refoc = r['paint'](soft_indicators(dist, torch.stack([s1, s2], 1)))
    if return_patches:  # blurry_edges_test_big.py:73-87
        return dict(P1=r['P1'], P2=r['P2'], sharp=sharp, refoc=refoc, lb=r['lb'], dmap=dmap, mask=m)
    n = cover_count(g, dt)
    pl = lambda t: t.reshape(B * g.L, -1, g.R, g.R)
    image = torch.stack([fold_sum(pl(r['P1']), B, g), fold_sum(pl(r['P2']), B, g)], 1) / n
    cnt = fold_sum(pl((m > 0).to(dt)), B, g)[:, 0]
    depth = fold_sum(pl(dmap), B, g)[:, 0] / torch.where(cnt > 0, cnt, torch.ones_like(cnt))
    return (image, fold_sum(pl(sharp), B, g) / n, fold_sum(pl(refoc), B, g) / n,
            fold_sum(pl(r['lb']), B, g) / n, depth, cnt / n)


def restore_global(raw: torch.Tensor) -> torch.Tensor:
    """global_training.py:141-145 / blurry_edges_test.py:135-138: raw network output [.,12]
    -> (xy*3, wrapped angles, eta coefficient raw+0.5)."""
    return torch.cat([raw[..., :4] * 3, torch.remainder((raw[..., 4:8] + 1) * math.pi, 2 * math.pi),
                      raw[..., 8:] + 0.5], dim=-1)


def global_loss(raw: torch.Tensor, img_ny: torch.Tensor, img_gt: torch.Tensor, bndry_dist: torch.Tensor,
                deri: torch.Tensor, bndry_depth: torch.Tensor, gammas, g: Geometry, cam: Camera,
                trace_form: bool = False, return_terms: bool = False, global_batch: int | None = None, mask_sum=None):
    """global_training.py:93-157.  raw [B,L,12] network output, img_* [B,2,H,W,3],
    bndry_dist/bndry_depth [B,H,W], deri [B,2,H-2,W-2,3]; gammas = (color, color_cons,
    bndry_cons, smthns, smthns_cons, bndry_loc, depth)."""
    B, dt, R, s = raw.shape[0], raw.dtype, g.R, g.stride
    est = restore_global(raw)
    planar = lambda t: t.reshape(B * 2, g.H, g.W, 3).permute(0, 3, 1, 2)
    r = render_pair(est, planar(img_ny).reshape(B, 2, 3, g.H, g.W), g, cam, trace_form)
    P = torch.stack([r['P1'], r['P2']], 1)                                   # [N,2,3,R,R]
    lb = r['lb']                                                             # [N,R,R]
    n = cover_count(g, dt)
    pm = lambda t: t.reshape(B, g.L, 2, -1, R, R).permute(0, 2, 1, 3, 4, 5).reshape(B * 2 * g.L, -1, R, R)
    # patch-major [N,2,C,R,R] <-> image-major [(B*2)*L,C,R,R]
    back = lambda t, C, r_: t.reshape(B, 2, g.L, C, r_, r_).permute(0, 2, 1, 3, 4, 5).reshape(-1, 2, C, r_, r_)
    gimg = (fold_sum(pm(P), B * 2, g) / n).detach()                          # [2B,3,H,W]
    gbnd = (fold_sum(lb.reshape(B * g.L, 1, R, R), B, g) / n).detach()       # [B,1,H,W]
    gt_p = back(extract(planar(img_gt), R, s), 3, R)
    gi_p = back(extract(gimg, R, s), 3, R)
    gb_p = extract(gbnd, R, s)[:, 0]
    Pd = sobel_mag(P)                                                        # [N,2,3,R-2,R-2]
    dgt_p = back(extract(deri.permute(0, 1, 4, 2, 3).reshape(B * 2, 3, g.H - 2, g.W - 2), R - 2, s), 3, R - 2)
    dgi_p = back(extract(sobel_mag(gimg), R - 2, s), 3, R - 2)
    bd_p = extract(torch.log2(bndry_dist + 1).unsqueeze(1), R, s)[:, 0]
    zg_p = extract(bndry_depth.unsqueeze(1), R, s)[:, 0]
    m = depth_mask(r['dist'], None)
    zero = torch.zeros((), dtype=dt)
    dmap = torch.where(m == 1, r['z1'].view(-1, 1, 1), torch.where(m == 2, r['z2'].view(-1, 1, 1), zero))
    msk = ((zg_p != 0) & (m != 0)).to(dt)
    # `global_batch` / `mask_sum`: this call holds one rank's slice of a larger batch; the means then run over the
    # GLOBAL batch (data-parallel recipe of DESIGN.md section 6) so that the ranks' terms simply add up.
    scale = 1.0 if global_batch is None else B / global_batch
    msum = msk.sum() if mask_sum is None else torch.as_tensor(mask_sum, dtype=dt)
    terms = torch.stack([
        ((gt_p - P) ** 2).sum(2).mean() * scale,
        ((P - gi_p) ** 2).sum(2).mean() * scale,
        ((lb - gb_p) ** 2).mean() * scale,
        ((Pd - dgt_p) ** 2).sum(2).mean() * scale,
        ((Pd - dgi_p) ** 2).sum(2).mean() * scale,
        ((bd_p * lb) ** 2).mean() * scale,
        (((dmap - zg_p) * msk) ** 2).sum() / msum])
    loss = (torch.as_tensor(gammas, dtype=dt) * terms).sum()
    if return_terms:
        return loss, terms, dict(gimg=gimg.reshape(B, 2, 3, g.H, g.W), gbnd=gbnd, msum=msk.sum())
    return loss


def local_loss(est: torch.Tensor, img_ny: torch.Tensor, gt_img: torch.Tensor, bndry_dist: torch.Tensor,
               deri: torch.Tensor, betas, g: Geometry, return_terms: bool = False):
    """local_training.py:32-52.  est [B,10] raw network output (angles wrapped here, eta
    coefficient used as is), img_ny/gt_img [B,R,R,3], bndry_dist [B,R,R], deri [B,R-2,R-2,3];
    betas = (bndry_loc, smthns)."""
    ang = torch.remainder(est[:, 4:8], 2 * math.pi)
    dist = wedge_distances(torch.cat([est[:, :4], ang], 1), g.R, g.w)
    u = soft_indicators(dist, eta_from_coef(est[:, 8:10]))
    C = ridge_colors(u.flatten(2).transpose(1, 2), img_ny.flatten(1, 2), g.lam)
    P = torch.einsum('nwij,nwc->ncij', u, C)
    lb = boundary_map(dist)
    terms = torch.stack([((gt_img.permute(0, 3, 1, 2) - P) ** 2).sum(1).mean(),
                         ((bndry_dist * lb) ** 2).mean(),
                         ((deri.permute(0, 3, 1, 2) - sobel_mag(P)) ** 2).sum(1).mean()])
    loss = terms[0] + betas[0] * terms[1] + betas[1] * terms[2]
    if return_terms:
        return loss, terms, dict(P=P, lb=lb, C=C)
    return loss


def gamma_schedule(epoch_idx: int, ranges, dynamic_epoch=(30, 100, 200)):
    """global_training.py:25-51: the 7 gamma weights at gamma_idx = epoch_idx."""
    e0, e1, e2 = dynamic_epoch
    if epoch_idx < e0:
        rate, k = epoch_idx / (e0 - 1), 0
    elif epoch_idx < e1:
        rate, k = 1.0, 0
    elif epoch_idx < e2:
        rate, k = (epoch_idx - e1) / (e2 - e1 - 1), 1
    else:
        rate, k = 1.0, 1
    return [r[k] + rate * (r[k + 1] - r[k]) for r in ranges]


# --------------------------------------------------------------------------------------
# big image: block geometry and stitching                     (blurry_edges_test_big.py:116-190)
# --------------------------------------------------------------------------------------
def big_blocks(big_H: int, big_W: int, g: Geometry, n_margin: int = 10):
    """Returns (block_stride_px, (nby, nbx), list of (iv, ih, y0, x0, (Vs,Ve,Hs,He), (Vsl,Vel,Hsl,Hel)))."""
    bs = g.H - g.R + g.stride - g.stride * n_margin * 2, g.W - g.R + g.stride - g.stride * n_margin * 2
    nb = (math.ceil((big_H - g.R - g.stride * n_margin * 2 + g.stride) / bs[0]),
          math.ceil((big_W - g.R - g.stride * n_margin * 2 + g.stride) / bs[1]))
    out = []
    for iv in range(nb[0]):
        for ih in range(nb[1]):
            vs, ve = int(iv == 0), int(iv == nb[0] - 1)
            hs, he = int(ih == 0), int(ih == nb[1] - 1)
            Vs = iv * (g.Hp - 2 * n_margin) + (1 - vs) * n_margin
            Ve = (iv + 1) * (g.Hp - 2 * n_margin) + (1 + ve) * n_margin
            Hs = ih * (g.Wp - 2 * n_margin) + (1 - hs) * n_margin
            He = (ih + 1) * (g.Wp - 2 * n_margin) + (1 + he) * n_margin
            loc = ((1 - vs) * n_margin, (ve - 1) * n_margin + g.Hp, (1 - hs) * n_margin, (he - 1) * n_margin + g.Wp)
            out.append((iv, ih, iv * bs[0], ih * bs[1], (Vs, Ve, Hs, He), loc))
    return bs, nb, out


def inference_big(est_blocks: torch.Tensor, img: torch.Tensor, g: Geometry, cam: Camera, big_H: int, big_W: int,
                  rho_prime: float = 10.39, n_margin: int = 10, thres: float = 0.05):
    """Pass B on every 147x147 block + stitch + fold at big size (blurry_edges_test_big.py:135-190).
    est_blocks [nblk,L,12] restored params per block (row-major block order), img [2,3,bigH,bigW].
    Returns the six maps at big size plus the thresholded depth."""
    dt = img.dtype
    gb = Geometry(R=g.R, stride=g.stride, H=big_H, W=big_W, w=g.w, alpha_lambda=g.alpha_lambda)
    _, _, blocks = big_blocks(big_H, big_W, g, n_margin)
    names = ('P1', 'P2', 'sharp', 'refoc', 'lb', 'dmap', 'mask')
    full = {k: None for k in names}
    for b, (iv, ih, y0, x0, (Vs, Ve, Hs, He), (Vsl, Vel, Hsl, Hel)) in enumerate(blocks):
        blk = img[:, :, y0:y0 + g.H, x0:x0 + g.W].unsqueeze(0)
        r = inference(est_blocks[b:b + 1], blk, g, cam, rho_prime, None, return_patches=True)
        for k in names:
            t = r[k].to(dt)
            t = t.reshape(g.Hp, g.Wp, -1, g.R, g.R)
            if full[k] is None:
                full[k] = torch.zeros(gb.Hp, gb.Wp, t.shape[2], g.R, g.R, dtype=dt)
            full[k][Vs:Ve, Hs:He] = t[Vsl:Vel, Hsl:Hel]
    n = cover_count(gb, dt)
    fs = lambda t: fold_sum(t.reshape(gb.L, -1, g.R, g.R), 1, gb)
    image = torch.stack([fs(full['P1']), fs(full['P2'])], 1) / n
    cnt = fs((full['mask'] > 0).to(dt))[:, 0]
    depth = fs(full['dmap'])[:, 0] / torch.where(cnt > 0, cnt, torch.ones_like(cnt))
    conf = cnt / n
    return (image, fs(full['sharp']) / n, fs(full['refoc']) / n, fs(full['lb']) / n, depth, conf,
            torch.where(conf > thres, depth, torch.zeros_like(depth)))
