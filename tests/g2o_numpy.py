"""Independent numpy transcription of the reference's LocalBundleAdjustment / PoseOptimization arithmetic.

TEST INFRASTRUCTURE.  This file pins oracle/liblld_oracle.so: it is a second, separately written restatement of the
same reference code, organised the way g2o is organised (vertices with oplus, edge classes with computeError /
linearizeOplus / constructQuadraticForm, a block solver with the explicit Schur loop, the Levenberg driver), not the way
the oracle is organised (flat per-landmark loops, envelope LDL^T).  Numerics go through numpy (batched einsum,
np.linalg.inv / solve), so the two share no code and no summation order; tests/test_oracle_pin.py demands that they agree
on per-iteration chi2 to 1e-9, on LM trial counts, on every outlier flag and on the final state.

Every function cites the reference file:line it transcribes (paths relative to the reference root).
Input / output are the lld_ba_problem / lld_pose_problem field dictionaries of lld_slam_b200.synth (the flattening of
src/Optimizer.cc:938-1218 is the shim's job and is tested separately).
"""
from __future__ import annotations

import numpy as np

DBL_MAX = np.finfo(np.float64).max


# ------------------------------------------------------------------------------------------------------------------
# Eigen conventions (Eigen/src/Geometry/Quaternion.h).  Quaternions are (x, y, z, w) like Eigen's coeffs().
# ------------------------------------------------------------------------------------------------------------------
def quat_from_R(R):
    """Eigen::Quaterniond(Matrix3d) — used by SE3Quat(R, t) (se3quat.h:58) and AddLineMinimal (LineOptimizer.cc:48)."""
    R = np.asarray(R, np.float64)
    q = np.zeros(4)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0.0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t
        q[1] = (R[0, 2] - R[2, 0]) * t
        q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    return q


def quat_to_R(q):
    """Eigen toRotationMatrix(); q (...,4) -> (...,3,3)"""
    q = np.asarray(q, np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1.0 - (tyy + tzz); R[..., 0, 1] = txy - twz; R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz; R[..., 1, 1] = 1.0 - (txx + tzz); R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy; R[..., 2, 1] = tyz + twx; R[..., 2, 2] = 1.0 - (txx + tyy)
    return R


def quat_mul(a, b):
    """Hamilton product a*b"""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_rot(q, v):
    """Eigen q*v (_transformVector): v + w*uv + qv x uv, uv = 2 qv x v; batched over leading dims"""
    qv = q[..., :3]
    uv = 2.0 * np.cross(qv, v)
    return v + q[..., 3:4] * uv + np.cross(qv, uv)


def skew(t):
    """cpmat / skew (types_six_dof_expmap.cpp:36-43, se3_ops.hpp); batched"""
    t = np.asarray(t, np.float64)
    M = np.zeros(t.shape[:-1] + (3, 3))
    M[..., 0, 1] = -t[..., 2]; M[..., 0, 2] = t[..., 1]
    M[..., 1, 0] = t[..., 2]; M[..., 1, 2] = -t[..., 0]
    M[..., 2, 0] = -t[..., 1]; M[..., 2, 1] = t[..., 0]
    return M


class SE3Quat:
    """Thirdparty/g2o/g2o/types/se3quat.h"""

    def __init__(self, q, t):
        self.q = np.array(q, np.float64)
        self.t = np.array(t, np.float64)
        self.normalize_rotation()

    @classmethod
    def from_Rt(cls, R, t):  # :58-60
        return cls(quat_from_R(R), t)

    def normalize_rotation(self):  # :280-285
        if self.q[3] < 0:
            self.q = -self.q
        self.q = self.q / np.sqrt(np.dot(self.q, self.q))

    def mul(self, o):  # operator* :104-110
        r = SE3Quat.__new__(SE3Quat)
        r.t = self.t + quat_rot(self.q, o.t)
        r.q = quat_mul(self.q, o.q)
        r.normalize_rotation()
        return r

    def copy(self):
        r = SE3Quat.__new__(SE3Quat)
        r.q = self.q.copy(); r.t = self.t.copy()
        return r

    @staticmethod
    def exp(update):  # :223-257
        omega = np.array(update[:3], np.float64)
        upsilon = np.array(update[3:6], np.float64)
        theta = np.sqrt(np.dot(omega, omega))
        Om = skew(omega)
        if theta < 0.00001:
            R = np.eye(3) + Om + Om @ Om      # (sic: no 1/2)
            V = R
        else:
            Om2 = Om @ Om
            R = np.eye(3) + np.sin(theta) / theta * Om + (1 - np.cos(theta)) / (theta * theta) * Om2
            V = np.eye(3) + (1 - np.cos(theta)) / (theta * theta) * Om + (theta - np.sin(theta)) / (theta ** 3) * Om2
        return SE3Quat(quat_from_R(R), V @ upsilon)

    def to_Rt12(self):  # Converter::toCvMat(SE3Quat) = to_homogeneous_matrix :267-275
        return np.concatenate([quat_to_R(self.q).reshape(9), self.t])


# ------------------------------------------------------------------------------------------------------------------
# robust kernel (core/robust_kernel_impl.cpp:65-91) — batched
# ------------------------------------------------------------------------------------------------------------------
def huber(e, delta):
    """returns rho0, rho1 for squared errors e (N,) and per-edge delta (N,)"""
    dsqr = delta * delta
    inl = e <= dsqr
    sq = np.sqrt(np.where(inl, 1.0, e))
    rho0 = np.where(inl, e, 2 * sq * delta - dsqr)
    rho1 = np.where(inl, 1.0, delta / sq)
    return rho0, rho1


# ------------------------------------------------------------------------------------------------------------------
# point edges (types/types_six_dof_expmap.{h,cpp}); batched over edges
# ------------------------------------------------------------------------------------------------------------------
def point_error(q, t, X, intr, obs, stereo, bf_as_float):
    """EdgeSE3ProjectXYZ::computeError (.h:94-99 + cam_project .cpp:149-155) and EdgeStereoSE3ProjectXYZ::computeError
    (.h:125-130 + cam_project .cpp:158-165: float invz; the binary edge receives bf as `const float&`).
    Returns err (N,3) with err[:,2] = 0 for monocular edges and the camera-frame point."""
    Xc = quat_rot(q, X) + t
    fx, fy, cx, cy, bf = (intr[:, k] for k in range(5))
    err = np.zeros((len(X), 3))
    # mono: project2d then * f + c
    pu = Xc[:, 0] / Xc[:, 2] * fx + cx
    pv = Xc[:, 1] / Xc[:, 2] * fy + cy
    # stereo
    invz = (1.0 / Xc[:, 2]).astype(np.float32).astype(np.float64)
    su = Xc[:, 0] * invz * fx + cx
    sv = Xc[:, 1] * invz * fy + cy
    if bf_as_float:   # `const float& bf` times `const float invz`: a float * float product, rounded to float (.cpp:158-165)
        sr = su - (bf.astype(np.float32) * invz.astype(np.float32)).astype(np.float64)
    else:             # OnlyPose: double member bf times float invz (.cpp:307-314)
        sr = su - bf * invz
    err[:, 0] = np.where(stereo, obs[:, 0] - su, obs[:, 0] - pu)
    err[:, 1] = np.where(stereo, obs[:, 1] - sv, obs[:, 1] - pv)
    err[:, 2] = np.where(stereo, obs[:, 2] - sr, 0.0)
    return err, Xc


def point_jacobians(q, Xc, intr, stereo):
    """linearizeOplus of EdgeSE3ProjectXYZ (.cpp:111-147) and EdgeStereoSE3ProjectXYZ (.cpp:196-242).
    Returns J_point (N,3,3), J_pose (N,3,6); third rows are zero for monocular edges."""
    R = quat_to_R(q)
    fx, fy, bf = intr[:, 0], intr[:, 1], intr[:, 4]
    x, y, z = Xc[:, 0], Xc[:, 1], Xc[:, 2]
    z2 = z * z
    n = len(x)
    Ji = np.zeros((n, 3, 3))
    # stereo rows (explicit formula :210-220)
    for c in range(3):
        Ji[:, 0, c] = -fx * R[:, 0, c] / z + fx * x * R[:, 2, c] / z2
        Ji[:, 1, c] = -fy * R[:, 1, c] / z + fy * y * R[:, 2, c] / z2
        Ji[:, 2, c] = Ji[:, 0, c] - bf * R[:, 2, c] / z2
    # mono: -1/z * tmp * R  (:123-132)
    tmp = np.zeros((n, 2, 3))
    tmp[:, 0, 0] = fx; tmp[:, 0, 2] = -x / z * fx
    tmp[:, 1, 1] = fy; tmp[:, 1, 2] = -y / z * fy
    Jm = (-1.0 / z)[:, None, None] * np.einsum("nij,njk->nik", tmp, R)
    mono = ~stereo
    Ji[mono, :2, :] = Jm[mono]
    Ji[mono, 2, :] = 0.0
    Jj = np.zeros((n, 3, 6))
    Jj[:, 0, 0] = x * y / z2 * fx
    Jj[:, 0, 1] = -(1 + (x * x / z2)) * fx
    Jj[:, 0, 2] = y / z * fx
    Jj[:, 0, 3] = -1.0 / z * fx
    Jj[:, 0, 5] = x / z2 * fx
    Jj[:, 1, 0] = (1 + y * y / z2) * fy
    Jj[:, 1, 1] = -x * y / z2 * fy
    Jj[:, 1, 2] = -x / z * fy
    Jj[:, 1, 4] = -1.0 / z * fy
    Jj[:, 1, 5] = y / z2 * fy
    Jj[:, 2, 0] = Jj[:, 0, 0] - bf * y / z2
    Jj[:, 2, 1] = Jj[:, 0, 1] + bf * x / z2
    Jj[:, 2, 2] = Jj[:, 0, 2]
    Jj[:, 2, 3] = Jj[:, 0, 3]
    Jj[:, 2, 5] = Jj[:, 0, 5] - bf / z2
    Jj[mono, 2, :] = 0.0
    return Ji, Jj


def pose_only_point_jacobian(Xc, intr, stereo):
    """EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose::linearizeOplus (.cpp:274-296, 343-372): invz form"""
    fx, fy, bf = intr[:, 0], intr[:, 1], intr[:, 4]
    x, y = Xc[:, 0], Xc[:, 1]
    invz = 1.0 / Xc[:, 2]
    invz2 = invz * invz
    J = np.zeros((len(x), 3, 6))
    J[:, 0, 0] = x * y * invz2 * fx
    J[:, 0, 1] = -(1 + (x * x * invz2)) * fx
    J[:, 0, 2] = y * invz * fx
    J[:, 0, 3] = -invz * fx
    J[:, 0, 5] = x * invz2 * fx
    J[:, 1, 0] = (1 + y * y * invz2) * fy
    J[:, 1, 1] = -x * y * invz2 * fy
    J[:, 1, 2] = -x * invz * fy
    J[:, 1, 4] = -invz * fy
    J[:, 1, 5] = y * invz2 * fy
    J[:, 2, 0] = J[:, 0, 0] - bf * y * invz2
    J[:, 2, 1] = J[:, 0, 1] + bf * x * invz2
    J[:, 2, 2] = J[:, 0, 2]
    J[:, 2, 3] = J[:, 0, 3]
    J[:, 2, 5] = J[:, 0, 5] - bf * invz2
    J[~stereo, 2, :] = 0.0
    return J


# ------------------------------------------------------------------------------------------------------------------
# line edges
# ------------------------------------------------------------------------------------------------------------------
def line_K(f, cx, cy):
    K = np.zeros((len(f), 3, 3))
    K[:, 0, 0] = f; K[:, 1, 1] = f; K[:, 0, 2] = cx; K[:, 1, 2] = cy; K[:, 2, 2] = 1.0
    return K


def line_error_from_X(qk, tk, X1, X2, K, b, x1, x2):
    """EdgeSE3ProjectLine::computeError (.h:344-375) / OnlyPose (.h:402-418) given the two 3-D points"""
    X1m = np.einsum("nij,nj->ni", K, quat_rot(qk, X1) + tk + b)
    X2m = np.einsum("nij,nj->ni", K, quat_rot(qk, X2) + tk + b)
    lt = np.cross(X1m, X2m)
    l = lt / np.sqrt(lt[:, 0] ** 2 + lt[:, 1] ** 2)[:, None]
    return np.stack([np.einsum("ni,ni->n", x1, l), np.einsum("ni,ni->n", x2, l)], axis=1)


def form_jacobian_line_wrt_cam(X1m, X2m, b, K):
    """FormJacobianLineWRTCam (.cpp:472-499) -> J_l (N,3,6), D_l_ltilde (N,3,3)"""
    n_ = len(X1m)
    KX1 = np.einsum("nij,nj->ni", K, X1m + b)
    KX2 = np.einsum("nij,nj->ni", K, X2m + b)
    lt = np.cross(KX1, KX2)
    n = np.sqrt(lt[:, 0] ** 2 + lt[:, 1] ** 2)
    dn = np.zeros((n_, 3))
    dn[:, 0] = -lt[:, 0] / (n * n * n)
    dn[:, 1] = -lt[:, 1] / (n * n * n)
    D = np.einsum("ni,nj->nij", lt, dn) + (1.0 / n)[:, None, None] * np.eye(3)
    J2 = np.zeros((n_, 3, 6)); J1 = np.zeros((n_, 3, 6))
    J2[:, :, :3] = -np.einsum("nij,njk->nik", K, skew(X2m)); J2[:, :, 3:] = K
    J1[:, :, :3] = -np.einsum("nij,njk->nik", K, skew(X1m)); J1[:, :, 3:] = K
    Jl = np.einsum("nij,njk->nik", skew(KX1), J2) - np.einsum("nij,njk->nik", skew(KX2), J1)
    Jl = np.einsum("nij,njk->nik", D, Jl)
    return Jl, D


def line_R(ql):
    """LineParams::GetR (types_sba.cpp:77-92): q.normalized().toRotationMatrix()"""
    qn = ql / np.sqrt(np.einsum("ni,ni->n", ql, ql))[:, None]
    return quat_to_R(qn)


def line_linearize(qk, tk, ql, alpha, K, b, x1, x2):
    """EdgeSE3ProjectLine::linearize (.cpp:507-553) -> J_line (N,2,4), J_pose (N,2,6)"""
    R = line_R(ql)
    X1 = R[:, :, 1] * alpha[:, None]
    X2 = X1 + R[:, :, 0]
    X1m = quat_rot(qk, X1) + tk
    X2m = quat_rot(qk, X2) + tk
    Jl, D = form_jacobian_line_wrt_cam(X1m, X2m, b, K)
    Jj = np.stack([np.einsum("ni,nij->nj", x1, Jl), np.einsum("ni,nij->nj", x2, Jl)], axis=1)
    dX1 = np.zeros((len(ql), 3, 4))
    dX1[:, :, :3] = 2 * (-skew(alpha[:, None] * R[:, :, 1]))
    dX1[:, :, 3] = R[:, :, 1]
    dX2 = dX1.copy()
    dX2[:, :, :3] = dX2[:, :, :3] - 2 * skew(R[:, :, 0])
    Rc = quat_to_R(qk)   # to_homogeneous_matrix().block<3,3>(0,0)
    KRc = np.einsum("nij,njk->nik", K, Rc)
    KX1 = np.einsum("nij,nj->ni", K, X1m + b)
    KX2 = np.einsum("nij,nj->ni", K, X2m + b)
    dlt = np.einsum("nij,njk,nkl->nil", skew(KX1), KRc, dX2) - np.einsum("nij,njk,nkl->nil", skew(KX2), KRc, dX1)
    Dl = np.einsum("nij,njk->nik", D, dlt)
    Ji = np.stack([np.einsum("ni,nij->nj", x1, Dl), np.einsum("ni,nij->nj", x2, Dl)], axis=1)
    return Ji, Jj


def reproject_line_point_depth(X0, ld, pp, K):
    """vgl::ReprojectLinePointTo3D (src/vgl.cc:336-346): least squares of the 3x2 system (colPivHouseholderQr)"""
    n = len(X0)
    M = np.zeros((n, 3, 2))
    M[:, 0, 0] = pp[:, 0]; M[:, 1, 0] = pp[:, 1]; M[:, 2, 0] = 1.0
    M[:, :, 1] = -np.einsum("nij,nj->ni", K, ld)
    rhs = np.einsum("nij,nj->ni", K, X0)
    sol = np.einsum("nij,nj->ni", np.linalg.pinv(M), rhs)
    return sol[:, 0]


def line_depth_positive(qk, tk, ql, alpha, K, b, x1, x2):
    """EdgeSE3ProjectLine::IsDepthPositive (.h:312-342)"""
    R = line_R(ql)
    X0 = R[:, :, 1] * alpha[:, None]
    ld = R[:, :, 0]
    X0l = quat_rot(qk, X0) + tk + b
    ldl = quat_rot(qk, X0 + ld) + tk + b - X0l
    d1 = reproject_line_point_depth(X0l, ldl, x1[:, :2], K)
    d2 = reproject_line_point_depth(X0l, ldl, x2[:, :2], K)
    return ~((d1 < 0) | (d2 < 0))


# ------------------------------------------------------------------------------------------------------------------
# the graph of one LocalBundleAdjustment window
# ------------------------------------------------------------------------------------------------------------------
class LocalBAGraph:
    """Vertices: keyframes (VertexSE3Expmap), points (VertexSBAPointXYZ, marginalised), lines (VertexSBALine, marginalised).
    Edges in insertion order: all point edges (src/Optimizer.cc:1093-1178), then all line edges
    (src/Optimizer.cc:1182-1218 -> LineOptimizer::AddLineMinimal src/LineOptimizer.cc:39-127)."""

    def __init__(self, p, w):
        k0, k1 = int(p["kf_off"][w]), int(p["kf_off"][w + 1])
        p0, p1 = int(p["pt_off"][w]), int(p["pt_off"][w + 1])
        l0, l1 = int(p["ln_off"][w]), int(p["ln_off"][w + 1])
        self.k0, self.p0, self.l0 = k0, p0, l0
        self.nk, self.np_, self.nl = k1 - k0, p1 - p0, l1 - l0
        self.kf = [SE3Quat.from_Rt(p["kf_Tcw"][k][:9].reshape(3, 3), p["kf_Tcw"][k][9:]) for k in range(k0, k1)]  # Converter::toSE3Quat
        self.kf_fixed = np.asarray(p["kf_fixed"][k0:k1]).astype(bool)
        self.intr = np.asarray(p["kf_intr"][k0:k1], np.float64)
        self.lcam = np.asarray(p["kf_line_cam"][k0:k1], np.float64)
        self.pts = np.array(p["pt_xyz"][p0:p1], np.float64)
        # points edges
        eo = np.asarray(p["pt_obs_off"], np.int64)
        e0, e1 = int(eo[p0]), int(eo[p1])
        self.pe0 = e0
        self.pe_pt = np.repeat(np.arange(self.np_), np.diff(eo[p0:p1 + 1]))
        self.pe_kf = np.asarray(p["pt_obs_kf"][e0:e1], np.int64)
        self.pe_obs = np.asarray(p["pt_obs_uvr"][e0:e1], np.float64)
        self.pe_stereo = ~(np.asarray(p["pt_obs_uvr"][e0:e1, 2]) < 0)
        self.pe_info = np.asarray(p["pt_obs_info"][e0:e1], np.float64)
        self.pe_delta = np.where(self.pe_stereo, float(p["delta_pt_stereo"]), float(p["delta_pt_mono"]))
        self.pe_robust = bool(p["robust_points"])
        self.pe_level = np.zeros(len(self.pe_pt), np.int64)
        self.pe_err = np.zeros((len(self.pe_pt), 3))
        # lines: LineParams from (X0, dir)  LineOptimizer.cc:44-50
        self.ln_q = np.zeros((self.nl, 4)); self.ln_alpha = np.zeros(self.nl)
        for i in range(self.nl):
            xd = np.asarray(p["ln_x0_dir"][l0 + i], np.float64)
            X0, d = xd[:3], xd[3:]
            nX = np.sqrt(np.dot(X0, X0))
            Rl = np.stack([d, X0 / nX, np.cross(d, X0) / nX], axis=1)
            self.ln_q[i] = quat_from_R(Rl)
            self.ln_alpha[i] = nX
        self.ln_in = np.array(p["ln_x0_dir"][l0:l1], np.float64)
        self.ln_removed = np.zeros(self.nl, bool)
        co = np.asarray(p["ln_obs_off"], np.int64)
        c0, c1 = int(co[l0]), int(co[l1])
        self.lc0 = c0
        ln, kf, side, cell, x1, x2, info, delta, cam, bx = [], [], [], [], [], [], [], [], [], []
        for i in range(self.nl):
            for c in range(int(co[l0 + i]), int(co[l0 + i + 1])):
                k = int(p["ln_obs_kf"][c])
                for si in range(2):
                    seg = np.asarray((p["ln_obs_left"] if si == 0 else p["ln_obs_right"])[c], np.float64)
                    if si == 1 and seg[0] < 0:
                        continue
                    if int(p["ln_endpoints_normalized"]):   # src/Optimizer.cc:234-235
                        fx, fy, cx, cy = self.intr[k][:4]
                        a = np.array([(seg[0] - cx) / fx, (seg[1] - cy) / fy, 1.0]); b_ = np.array([(seg[2] - cx) / fx, (seg[3] - cy) / fy, 1.0])
                    else:                                    # Ki.setIdentity()  LineOptimizer.cc:106-113
                        a = np.array([seg[0], seg[1], 1.0]); b_ = np.array([seg[2], seg[3], 1.0])
                    ln.append(i); kf.append(k); side.append(si); cell.append(c - c0); x1.append(a); x2.append(b_)
                    info.append(float(p["ln_obs_info"][c][si]))
                    delta.append(float(p["delta_ln_stereo"]) if p["ln_obs_stereo"][c] else float(p["delta_ln_mono"]))
                    cam.append(self.lcam[k][:3]); bx.append(-self.lcam[k][3] if si == 1 else 0.0)
        n = len(ln)
        self.le_ln = np.array(ln, np.int64); self.le_kf = np.array(kf, np.int64); self.le_side = np.array(side, np.int64)
        self.le_cell = np.array(cell, np.int64)
        self.le_x1 = np.array(x1, np.float64).reshape(n, 3); self.le_x2 = np.array(x2, np.float64).reshape(n, 3)
        self.le_info = np.array(info, np.float64); self.le_delta = np.array(delta, np.float64)
        cam = np.array(cam, np.float64).reshape(n, 3)
        self.le_K = line_K(cam[:, 0], cam[:, 1], cam[:, 2])
        self.le_b = np.zeros((n, 3)); self.le_b[:, 0] = np.array(bx, np.float64)
        self.le_robust = True
        self.le_level = np.zeros(n, np.int64)
        self.le_err = np.zeros((n, 2))
        self.n_lcell = c1 - c0
        if n:   # e->computeError() at creation  LineOptimizer.cc:114
            self.le_err = self._line_err(np.arange(n))
        self.ln_filter = int(p["ln_filter"])
        self.chi2_pt = (float(p["chi2_pt_mono"]), float(p["chi2_pt_stereo"]))
        self.chi2_log, self.lambda_log, self.trials_log = [], [], []

    # ---- estimates as arrays ----
    def _kq(self):
        return np.array([s.q for s in self.kf]), np.array([s.t for s in self.kf])

    def _pt_err(self, idx):
        q, t = self._kq()
        k = self.pe_kf[idx]
        err, _ = point_error(q[k], t[k], self.pts[self.pe_pt[idx]], self.intr[k], self.pe_obs[idx], self.pe_stereo[idx], True)
        return err

    def _line_err(self, idx):
        q, t = self._kq()
        k = self.le_kf[idx]; l = self.le_ln[idx]
        R = line_R(self.ln_q[l])
        X1 = R[:, :, 1] * self.ln_alpha[l][:, None]
        X2 = X1 + R[:, :, 0]
        return line_error_from_X(q[k], t[k], X1, X2, self.le_K[idx], self.le_b[idx], self.le_x1[idx], self.le_x2[idx])

    # ---- SparseOptimizer ----
    def initialize_optimization(self, level=0):
        """sparse_optimizer.cpp:199-267: active edges = edges at `level` whose vertices all exist (removed lines do not)."""
        self.pe_act = np.nonzero(self.pe_level == level)[0]
        self.le_act = np.nonzero((self.le_level == level) & ~self.ln_removed[self.le_ln])[0] if len(self.le_ln) else np.zeros(0, np.int64)
        kf_used = np.zeros(self.nk, bool)
        kf_used[self.pe_kf[self.pe_act]] = True
        if len(self.le_act):
            kf_used[self.le_kf[self.le_act]] = True
        self.act_kf = np.nonzero(kf_used & ~self.kf_fixed)[0]              # buildIndexMapping :166-190: non-marginalised first
        self.act_pt = np.unique(self.pe_pt[self.pe_act])
        self.act_ln = np.unique(self.le_ln[self.le_act]) if len(self.le_act) else np.zeros(0, np.int64)
        self.kf_col = -np.ones(self.nk, np.int64); self.kf_col[self.act_kf] = np.arange(len(self.act_kf))
        self.pt_col = -np.ones(self.np_, np.int64); self.pt_col[self.act_pt] = np.arange(len(self.act_pt))
        self.ln_col = -np.ones(self.nl, np.int64); self.ln_col[self.act_ln] = np.arange(len(self.act_ln))
        return len(self.act_kf) + len(self.act_pt) + len(self.act_ln) > 0

    def compute_active_errors(self):  # :61-85
        if len(self.pe_act):
            self.pe_err[self.pe_act] = self._pt_err(self.pe_act)
        if len(self.le_act):
            self.le_err[self.le_act] = self._line_err(self.le_act)

    def _chi2_pt(self, idx):
        e = self.pe_err[idx]
        return np.einsum("ni,ni->n", e, self.pe_info[idx][:, None] * e)   # _error.dot(information()*_error)  base_edge.h:58-61

    def _chi2_ln(self, idx):
        e = self.le_err[idx]
        return np.einsum("ni,ni->n", e, self.le_info[idx][:, None] * e)

    def active_robust_chi2(self):  # :100-114, edge insertion order
        chi = 0.0
        if len(self.pe_act):
            c = self._chi2_pt(self.pe_act)
            if self.pe_robust:
                c = huber(c, self.pe_delta[self.pe_act])[0]
            for x in c:
                chi += x
        if len(self.le_act):
            c = self._chi2_ln(self.le_act)
            if self.le_robust:
                c = huber(c, self.le_delta[self.le_act])[0]
            for x in c:
                chi += x
        return chi

    # ---- BlockSolver::buildSystem (block_solver.hpp:502-560) with constructQuadraticForm (base_binary_edge.hpp:55-120) ----
    def build_system(self):
        nK, nP, nL = len(self.act_kf), len(self.act_pt), len(self.act_ln)
        self.Hpp = np.zeros((nK, nK, 6, 6)); self.bp = np.zeros((nK, 6))
        self.Hll_p = np.zeros((nP, 3, 3)); self.bl_p = np.zeros((nP, 3))
        self.Hll_l = np.zeros((nL, 4, 4)); self.bl_l = np.zeros((nL, 4))
        self.Hpl_p = {}   # (kf col, pt col) -> 6x3
        self.Hpl_l = {}   # (kf col, ln col) -> 6x4
        q, t = self._kq()
        if len(self.pe_act):
            idx = self.pe_act
            k = self.pe_kf[idx]; pt = self.pe_pt[idx]
            err, Xc = point_error(q[k], t[k], self.pts[pt], self.intr[k], self.pe_obs[idx], self.pe_stereo[idx], True)
            err = self.pe_err[idx]       # _error as left by computeActiveErrors
            A, B = point_jacobians(q[k], Xc, self.intr[k], self.pe_stereo[idx])
            info = self.pe_info[idx]
            rho1 = huber(self._chi2_pt(idx), self.pe_delta[idx])[1] if self.pe_robust else np.ones(len(idx))
            wO = rho1 * info
            omega_r = -(info[:, None] * err) * rho1[:, None]
            AtO = np.einsum("nda,n->nad", A, wO)
            np.add.at(self.bl_p, self.pt_col[pt], np.einsum("nda,nd->na", A, omega_r))
            np.add.at(self.Hll_p, self.pt_col[pt], np.einsum("nad,ndb->nab", AtO, A))
            free = self.kf_col[k] >= 0
            BtO = np.einsum("nda,n->nad", B, wO)
            kc = self.kf_col[k]
            np.add.at(self.bp, kc[free], np.einsum("nda,nd->na", B, omega_r)[free])
            HB = np.einsum("nad,ndb->nab", BtO, B)
            np.add.at(self.Hpp, (kc[free], kc[free]), HB[free])
            Wm = np.einsum("nad,ndb->nab", BtO, A)      # pose rows x landmark cols (Hpl block, _hessianTransposed layout)
            for e in np.nonzero(free)[0]:
                self.Hpl_p[(int(kc[e]), int(self.pt_col[pt[e]]))] = self.Hpl_p.get((int(kc[e]), int(self.pt_col[pt[e]])), 0) + Wm[e]
        if len(self.le_act):
            idx = self.le_act
            k = self.le_kf[idx]; l = self.le_ln[idx]
            A, B = line_linearize(q[k], t[k], self.ln_q[l], self.ln_alpha[l], self.le_K[idx], self.le_b[idx], self.le_x1[idx], self.le_x2[idx])
            err = self.le_err[idx]
            info = self.le_info[idx]
            rho1 = huber(self._chi2_ln(idx), self.le_delta[idx])[1] if self.le_robust else np.ones(len(idx))
            wO = rho1 * info
            omega_r = -(info[:, None] * err) * rho1[:, None]
            AtO = np.einsum("nda,n->nad", A, wO)
            np.add.at(self.bl_l, self.ln_col[l], np.einsum("nda,nd->na", A, omega_r))
            np.add.at(self.Hll_l, self.ln_col[l], np.einsum("nad,ndb->nab", AtO, A))
            kc = self.kf_col[k]
            free = kc >= 0
            BtO = np.einsum("nda,n->nad", B, wO)
            np.add.at(self.bp, kc[free], np.einsum("nda,nd->na", B, omega_r)[free])
            np.add.at(self.Hpp, (kc[free], kc[free]), np.einsum("nad,ndb->nab", BtO, B)[free])
            Wm = np.einsum("nad,ndb->nab", BtO, A)
            for e in np.nonzero(free)[0]:
                key = (int(kc[e]), int(self.ln_col[l[e]]))
                self.Hpl_l[key] = self.Hpl_l.get(key, 0) + Wm[e]

    def compute_lambda_init(self):  # optimization_algorithm_levenberg.cpp:166-180
        m = 0.0
        for a in range(len(self.act_kf)):
            m = max(m, np.abs(np.diag(self.Hpp[a, a])).max())
        if len(self.act_pt):
            m = max(m, np.abs(np.einsum("nii->ni", self.Hll_p)).max())
        if len(self.act_ln):
            m = max(m, np.abs(np.einsum("nii->ni", self.Hll_l)).max())
        return 1e-5 * m

    # ---- BlockSolver::solve with lambda on every diagonal (block_solver.hpp:354-486, setLambda :564-590) ----
    def solve(self, lam):
        nK = len(self.act_kf)
        Hs = np.zeros((6 * nK, 6 * nK))
        for a in range(nK):
            Hs[6 * a:6 * a + 6, 6 * a:6 * a + 6] = self.Hpp[a, a] + lam * np.eye(6)
        coeff = np.zeros(6 * nK)
        by_lm_p = {}
        for (a, j), W in self.Hpl_p.items():
            by_lm_p.setdefault(j, []).append((a, W))
        by_lm_l = {}
        for (a, j), W in self.Hpl_l.items():
            by_lm_l.setdefault(j, []).append((a, W))
        Dinv_p = np.linalg.inv(self.Hll_p + lam * np.eye(3)) if len(self.act_pt) else np.zeros((0, 3, 3))   # D->inverse() :389
        Dinv_l = np.linalg.inv(self.Hll_l + lam * np.eye(4)) if len(self.act_ln) else np.zeros((0, 4, 4))
        for Dinv, bl, by in ((Dinv_p, self.bl_p, by_lm_p), (Dinv_l, self.bl_l, by_lm_l)):
            for j in range(len(Dinv)):
                db = Dinv[j] @ bl[j]
                col = sorted(by.get(j, []), key=lambda x: x[0])
                for ii, (i1, Bi) in enumerate(col):
                    BDinv = Bi @ Dinv[j]
                    coeff[6 * i1:6 * i1 + 6] += Bi @ db
                    for (i2, Bj) in col[ii:]:
                        Hs[6 * i1:6 * i1 + 6, 6 * i2:6 * i2 + 6] -= BDinv @ Bj.T     # upper triangular blocks only :424-431
        bs = self.bp.reshape(-1) - coeff
        # LinearSolverEigen (solvers/linear_solver_eigen.h:94-124): LDL^T of the upper-triangular view
        Hfull = np.triu(Hs) + np.triu(Hs, 1).T
        ok = True
        try:
            xp = np.linalg.solve(Hfull, bs) if nK else np.zeros(0)
            if not np.all(np.isfinite(xp)):
                ok = False
        except np.linalg.LinAlgError:
            ok = False
            xp = np.zeros(6 * nK)
        if not ok:
            return False, None
        # xl = Dinv (bl - Hpl^T xp) :463-481
        cl_p = self.bl_p.copy(); cl_l = self.bl_l.copy()
        for (a, j), W in self.Hpl_p.items():
            cl_p[j] -= W.T @ xp[6 * a:6 * a + 6]
        for (a, j), W in self.Hpl_l.items():
            cl_l[j] -= W.T @ xp[6 * a:6 * a + 6]
        xl_p = np.einsum("nij,nj->ni", Dinv_p, cl_p) if len(cl_p) else cl_p
        xl_l = np.einsum("nij,nj->ni", Dinv_l, cl_l) if len(cl_l) else cl_l
        return True, (xp.reshape(nK, 6), xl_p, xl_l)

    def b_vector(self):
        return np.concatenate([self.bp.reshape(-1), self.bl_p.reshape(-1), self.bl_l.reshape(-1)])

    # ---- vertex updates ----
    def push(self):
        return ([s.copy() for s in self.kf], self.pts.copy(), self.ln_q.copy(), self.ln_alpha.copy())

    def pop(self, st):
        self.kf, self.pts, self.ln_q, self.ln_alpha = st

    def update(self, x):  # sparse_optimizer.cpp:422-435 -> oplusImpl of each vertex type
        xp, xlp, xll = x
        for a, k in enumerate(self.act_kf):   # VertexSE3Expmap::oplusImpl  types_six_dof_expmap.h:76-79
            self.kf[k] = SE3Quat.exp(xp[a]).mul(self.kf[k])
        if len(self.act_pt):                  # VertexSBAPointXYZ::oplusImpl  types_sba.h:55-59
            self.pts[self.act_pt] += xlp
        for j, l in enumerate(self.act_ln):   # VertexSBALine::oplusImpl  types_sba.h:97-108
            r = xll[j, :3]
            with np.errstate(invalid="ignore"):
                qr = np.array([r[0], r[1], r[2], np.sqrt(1.0 - np.dot(r, r))])
            qn = self.ln_q[l] / np.sqrt(np.dot(self.ln_q[l], self.ln_q[l]))    # GetQ() normalises
            self.ln_q[l] = quat_mul(qr, qn)
            self.ln_alpha[l] = self.ln_alpha[l] + xll[j, 3]

    # ---- OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:61-164) ----
    def lm_solve(self, iteration):
        self.compute_active_errors()
        current = self.active_robust_chi2()
        temp = current
        ini = current
        self.build_system()
        if iteration == 0:
            self.lam = self.compute_lambda_init()
            self.ni = 2.0
            self.nbad = 0
            self.chi2_log.append(current)
        rho = 0.0
        qmax = 0
        while True:
            backup = self.push()
            ok2, x = self.solve(self.lam)
            if ok2:
                self.update(x)
            self.compute_active_errors()
            temp = self.active_robust_chi2()
            if not ok2:
                temp = DBL_MAX
            rho = current - temp
            scale = 0.0
            if ok2:   # computeScale :182-189 over poses then landmarks (x of a failed solve is whatever was there; never accepted)
                xv = np.concatenate([x[0].reshape(-1), x[1].reshape(-1), x[2].reshape(-1)])
                bv = self.b_vector()
                for j in range(len(xv)):
                    scale += xv[j] * (self.lam * xv[j] + bv[j])
            scale += 1e-3
            rho /= scale
            if rho > 0 and np.isfinite(temp):
                alpha = 1.0 - (2 * rho - 1) ** 3
                alpha = min(alpha, 2.0 / 3.0)
                self.lam *= max(1.0 / 3.0, alpha)
                self.ni = 2.0
                current = temp
            else:
                self.lam *= self.ni
                self.ni *= 2
                self.pop(backup)
            qmax += 1
            if not (rho < 0 and qmax < 10):
                break
        self.chi2_log.append(current); self.lambda_log.append(self.lam); self.trials_log.append(qmax)
        if qmax == 10 or rho == 0:
            return False
        if (ini - current) * 1e3 < ini:
            self.nbad += 1
        else:
            self.nbad = 0
        return self.nbad < 3

    def optimize(self, iterations):  # sparse_optimizer.cpp:354-419
        done = 0
        ok = True
        i = 0
        while i < iterations and ok:
            ok = self.lm_solve(i)
            done += 1
            i += 1
        return done


def local_bundle_adjustment(p, its1=5, its2=15):
    """Optimizer::LocalBundleAdjustment src/Optimizer.cc:1220-1329 on every window of the batch; returns the lld_ba_result fields."""
    nw = int(p["n_win"])
    st = its1 + its2 + 2
    out = dict(kf_Tcw=np.zeros((int(p["kf_off"][-1]), 12)), pt_xyz=np.zeros((int(p["pt_off"][-1]), 3)),
               ln_x0_dir=np.zeros((int(p["ln_off"][-1]), 6)), pt_obs_bad=np.zeros(int(p["pt_obs_off"][-1]), np.uint8),
               ln_obs_bad=np.zeros((int(p["ln_obs_off"][-1]), 2), np.uint8), ln_removed=np.zeros(int(p["ln_off"][-1]), np.uint8),
               chi2_log=np.zeros((nw, st)), lambda_log=np.zeros((nw, st)), trials_log=np.zeros((nw, st), np.int32),
               n_iter_done=np.zeros((nw, 2), np.int32))
    for w in range(nw):
        G = LocalBAGraph(p, w)
        it1 = it2 = 0
        if G.initialize_optimization(0):
            it1 = G.optimize(its1)
        # :1234-1270 point gates on the stale _error of the last LM trial, fresh isDepthPositive
        q, t = G._kq()
        if len(G.pe_pt):
            chi = G._chi2_pt(np.arange(len(G.pe_pt)))
            zc = (quat_rot(q[G.pe_kf], G.pts[G.pe_pt]) + t[G.pe_kf])[:, 2]
            th = np.where(G.pe_stereo, G.chi2_pt[1], G.chi2_pt[0])
            G.pe_level[(chi > th) | ~(zc > 0)] = 1
        G.pe_robust = False
        # LineOptimizer::DisableOutliers src/LineOptimizer.cc:129-170
        if len(G.le_ln):
            alle = np.arange(len(G.le_ln))
            chi = G._chi2_ln(alle)
            dp = line_depth_positive(q[G.le_kf], t[G.le_kf], G.ln_q[G.le_ln], G.ln_alpha[G.le_ln], G.le_K, G.le_b, G.le_x1, G.le_x2)
            bad = (chi > G.le_delta * G.le_delta) | ~dp
            G.le_level[bad] = 1
            cnt = np.zeros(G.nl, np.int64)
            np.add.at(cnt, G.le_ln[~bad], 2)
            has_edge = np.zeros(G.nl, bool); has_edge[G.le_ln] = True
            G.ln_removed = has_edge & (cnt <= G.ln_filter)
        G.le_robust = False
        if G.initialize_optimization(0):
            it2 = G.optimize(its2)
        # :1276-1311 final point flags (stale errors again), LineOptimizer::GetLineData :172-201 (fresh errors)
        q, t = G._kq()
        if len(G.pe_pt):
            chi = G._chi2_pt(np.arange(len(G.pe_pt)))
            zc = (quat_rot(q[G.pe_kf], G.pts[G.pe_pt]) + t[G.pe_kf])[:, 2]
            th = np.where(G.pe_stereo, G.chi2_pt[1], G.chi2_pt[0])
            out["pt_obs_bad"][G.pe0:G.pe0 + len(G.pe_pt)] = ((chi > th) | ~(zc > 0)).astype(np.uint8)
        if len(G.le_ln):
            alle = np.arange(len(G.le_ln))
            dp = line_depth_positive(q[G.le_kf], t[G.le_kf], G.ln_q[G.le_ln], G.ln_alpha[G.le_ln], G.le_K, G.le_b, G.le_x1, G.le_x2)
            G.le_err = G._line_err(alle)
            chi = G._chi2_ln(alle)
            bad = ((chi > G.le_delta * G.le_delta) | ~dp) & ~G.ln_removed[G.le_ln]
            out["ln_obs_bad"][G.lc0 + G.le_cell[bad], G.le_side[bad]] = 1
        out["ln_removed"][G.l0:G.l0 + G.nl] = G.ln_removed
        for k in range(G.nk):
            out["kf_Tcw"][G.k0 + k] = G.kf[k].to_Rt12()
        out["pt_xyz"][G.p0:G.p0 + G.np_] = G.pts
        if G.nl:
            R = line_R(G.ln_q)
            xd = np.concatenate([G.ln_alpha[:, None] * R[:, :, 1], R[:, :, 0]], axis=1)     # LineOptimizer.cc:180-182
            xd[G.ln_removed] = G.ln_in[G.ln_removed]                                          # GetLineData false: caller keeps its value
            out["ln_x0_dir"][G.l0:G.l0 + G.nl] = xd
        n = len(G.chi2_log)
        out["chi2_log"][w, :min(n, st)] = G.chi2_log[:st]
        out["lambda_log"][w, :len(G.lambda_log)] = G.lambda_log
        out["trials_log"][w, :len(G.trials_log)] = G.trials_log
        out["n_iter_done"][w] = (it1, it2)
    return out


def bundle_adjustment(p, n_iter=10):
    """Optimizer::BundleAdjustment src/Optimizer.cc:321-559 (GlobalBundleAdjustemnt :312-319): one optimize(nIterations) over every
    keyframe, point and line; point edges robust only with bRobust (:411-416), line edges always (AddLineMinimalGlobal :217-219),
    line endpoints K^-1-normalised (:234-235), no outlier gates, nothing removed."""
    nw = int(p["n_win"])
    st = n_iter + 2
    out = dict(kf_Tcw=np.zeros((int(p["kf_off"][-1]), 12)), pt_xyz=np.zeros((int(p["pt_off"][-1]), 3)),
               ln_x0_dir=np.zeros((int(p["ln_off"][-1]), 6)), pt_obs_bad=np.zeros(int(p["pt_obs_off"][-1]), np.uint8),
               ln_obs_bad=np.zeros((int(p["ln_obs_off"][-1]), 2), np.uint8), ln_removed=np.zeros(int(p["ln_off"][-1]), np.uint8),
               chi2_log=np.zeros((nw, st)), lambda_log=np.zeros((nw, st)), trials_log=np.zeros((nw, st), np.int32),
               n_iter_done=np.zeros((nw, 2), np.int32))
    for w in range(nw):
        G = LocalBAGraph(p, w)
        it = G.optimize(n_iter) if G.initialize_optimization(0) else 0
        for k in range(G.nk):
            out["kf_Tcw"][G.k0 + k] = G.kf[k].to_Rt12()
        out["pt_xyz"][G.p0:G.p0 + G.np_] = G.pts
        if G.nl:
            R = line_R(G.ln_q)
            out["ln_x0_dir"][G.l0:G.l0 + G.nl] = np.concatenate([G.ln_alpha[:, None] * R[:, :, 1], R[:, :, 0]], axis=1)
        out["chi2_log"][w, :min(len(G.chi2_log), st)] = G.chi2_log[:st]
        out["lambda_log"][w, :len(G.lambda_log)] = G.lambda_log
        out["trials_log"][w, :len(G.trials_log)] = G.trials_log
        out["n_iter_done"][w] = (it, 0)
    return out


# ------------------------------------------------------------------------------------------------------------------
# Optimizer::PoseOptimization  src/Optimizer.cc:653-932 (+ AddLineMinOnlyPose :562-650)
# ------------------------------------------------------------------------------------------------------------------
def pose_optimization(p):
    F = int(p["n_frames"])
    out = dict(Tcw=np.zeros((F, 12)), pt_outlier=np.zeros(int(p["pt_off"][-1]), np.uint8),
               ln_outlier=np.zeros(int(p["ln_off"][-1]), np.uint8), n_inliers=np.zeros(F, np.int32))
    for f in range(F):
        T0 = SE3Quat.from_Rt(p["Tcw"][f][:9].reshape(3, 3), p["Tcw"][f][9:])
        a, b = int(p["pt_off"][f]), int(p["pt_off"][f + 1])
        la, lb = int(p["ln_off"][f]), int(p["ln_off"][f + 1])
        n = b - a
        Xw = np.asarray(p["pt_xw"][a:b], np.float64)
        obs = np.asarray(p["pt_uvr"][a:b], np.float64)
        stereo = ~(np.asarray(p["pt_uvr"][a:b, 2]) < 0)
        info = np.asarray(p["pt_info"][a:b], np.float64)
        delta = np.where(stereo, float(p["delta_stereo"]), float(p["delta_mono"]))
        intr = np.repeat(np.asarray(p["intr"][f], np.float64)[None], n, axis=0)
        # line edges in insertion order
        lcam = np.asarray(p["line_cam"][f], np.float64)
        l_ln, l_side, X1, X2, x1, x2, linfo, ldelta, lb_, lgate = [], [], [], [], [], [], [], [], [], []
        for i in range(la, lb):
            xd = np.asarray(p["ln_x0_dir"][i], np.float64)
            for si in range(2):
                seg = np.asarray((p["ln_left"] if si == 0 else p["ln_right"])[i], np.float64)
                if si == 1 and seg[0] < 0:
                    continue
                l_ln.append(i - la); l_side.append(si); X1.append(xd[:3]); X2.append(xd[:3] + xd[3:])
                x1.append([seg[0], seg[1], 1.0]); x2.append([seg[2], seg[3], 1.0])
                linfo.append(float(p["ln_info"][i][si]))
                ldelta.append(float(p["delta_ln_stereo"]) if p["ln_stereo"][i] else float(p["delta_ln_mono"]))
                lb_.append(-lcam[3] if si == 1 else 0.0)
                lgate.append(float(p["gate_ln_stereo"]) if p["ln_gate_stereo"][i][si] else float(p["gate_ln_mono"]))
        m = len(l_ln)
        l_ln = np.array(l_ln, np.int64); X1 = np.array(X1).reshape(m, 3); X2 = np.array(X2).reshape(m, 3)
        x1 = np.array(x1).reshape(m, 3); x2 = np.array(x2).reshape(m, 3); linfo = np.array(linfo); ldelta = np.array(ldelta)
        lbv = np.zeros((m, 3)); lbv[:, 0] = np.array(lb_); lgate = np.array(lgate)
        LK = line_K(np.full(m, lcam[0]), np.full(m, lcam[1]), np.full(m, lcam[2]))
        if n < 3:   # nInitialCorrespondences<3 :809-810
            out["Tcw"][f] = p["Tcw"][f]
            continue
        pt_level = np.zeros(n, np.int64); ln_level = np.zeros(m, np.int64)
        pt_err = np.zeros((n, 3)); ln_err = np.zeros((m, 2))
        pt_rob = True; ln_rob = True
        outl = np.zeros(n, bool); outl_ln = np.zeros(lb - la, bool)
        T = T0.copy()
        nbad_total = 0

        def errs(T, pidx, lidx):
            if len(pidx):
                qq = np.repeat(T.q[None], len(pidx), 0); tt = np.repeat(T.t[None], len(pidx), 0)
                pt_err[pidx] = point_error(qq, tt, Xw[pidx], intr[pidx], obs[pidx], stereo[pidx], False)[0]
            if len(lidx):
                qq = np.repeat(T.q[None], len(lidx), 0); tt = np.repeat(T.t[None], len(lidx), 0)
                ln_err[lidx] = line_error_from_X(qq, tt, X1[lidx], X2[lidx], LK[lidx], lbv[lidx], x1[lidx], x2[lidx])

        def chi2p(idx):
            return np.einsum("ni,ni->n", pt_err[idx], info[idx][:, None] * pt_err[idx])

        def chi2l(idx):
            return np.einsum("ni,ni->n", ln_err[idx], linfo[idx][:, None] * ln_err[idx])

        for it in range(int(p["n_rounds"])):
            T = T0.copy()                                    # vSE3->setEstimate(Converter::toSE3Quat(pFrame->mTcw)) :823
            pa = np.nonzero(pt_level == 0)[0]; lact = np.nonzero(ln_level == 0)[0]
            if len(pa) + len(lact) > 0:
                lam = 0.0; ni = 2.0; nbad = 0
                for iteration in range(int(p["its"])):
                    errs(T, pa, lact)

                    def robust_chi2():
                        chi = 0.0
                        c = chi2p(pa)
                        if pt_rob:
                            c = huber(c, delta[pa])[0]
                        for x in c:
                            chi += x
                        c = chi2l(lact)
                        if ln_rob:
                            c = huber(c, ldelta[lact])[0]
                        for x in c:
                            chi += x
                        return chi
                    current = robust_chi2(); ini = current
                    # buildSystem: unary quadratic forms (base_unary_edge.hpp:43-72)
                    H = np.zeros((6, 6)); bvec = np.zeros(6)
                    if len(pa):
                        qq = np.repeat(T.q[None], len(pa), 0); tt = np.repeat(T.t[None], len(pa), 0)
                        Xc = quat_rot(qq, Xw[pa]) + tt
                        A = pose_only_point_jacobian(Xc, intr[pa], stereo[pa])
                        r1 = huber(chi2p(pa), delta[pa])[1] if pt_rob else np.ones(len(pa))
                        bvec -= np.einsum("n,nda,n,nd->a", r1, A, info[pa], pt_err[pa])
                        H += np.einsum("nda,n,ndb->ab", A, r1 * info[pa], A)
                    if len(lact):
                        qq = np.repeat(T.q[None], len(lact), 0); tt = np.repeat(T.t[None], len(lact), 0)
                        Jl, _ = form_jacobian_line_wrt_cam(quat_rot(qq, X1[lact]) + tt, quat_rot(qq, X2[lact]) + tt, lbv[lact], LK[lact])
                        A = np.stack([np.einsum("ni,nij->nj", x1[lact], Jl), np.einsum("ni,nij->nj", x2[lact], Jl)], axis=1)
                        r1 = huber(chi2l(lact), ldelta[lact])[1] if ln_rob else np.ones(len(lact))
                        bvec -= np.einsum("n,nda,n,nd->a", r1, A, linfo[lact], ln_err[lact])
                        H += np.einsum("nda,n,ndb->ab", A, r1 * linfo[lact], A)
                    if iteration == 0:
                        lam = 1e-5 * np.abs(np.diag(H)).max(); ni = 2.0; nbad = 0
                    rho = 0.0; qmax = 0
                    while True:
                        Tb = T.copy()
                        Hl = H + lam * np.eye(6)
                        ok2 = True
                        try:  # LinearSolverDense: Eigen LDLT + isPositive (solvers/linear_solver_dense.h:65-113)
                            np.linalg.cholesky(Hl)
                            x = np.linalg.solve(Hl, bvec)
                        except np.linalg.LinAlgError:
                            ok2 = False
                        if ok2:
                            T = SE3Quat.exp(x).mul(T)
                        errs(T, pa, lact)
                        temp = robust_chi2()
                        if not ok2:
                            temp = DBL_MAX
                        rho = current - temp
                        scale = 1e-3
                        if ok2:
                            for j in range(6):
                                scale += x[j] * (lam * x[j] + bvec[j])
                        rho /= scale
                        if rho > 0 and np.isfinite(temp):
                            alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                            lam *= max(1.0 / 3.0, alpha); ni = 2.0; current = temp
                        else:
                            lam *= ni; ni *= 2; T = Tb
                        qmax += 1
                        if not (rho < 0 and qmax < 10):
                            break
                    if qmax == 10 or rho == 0:
                        break
                    if (ini - current) * 1e3 < ini:
                        nbad += 1
                    else:
                        nbad = 0
                    if nbad >= 3:
                        break
            # classification :825-884 : outliers are re-evaluated, inliers keep the error of the last LM trial; float compares
            nbad_total = 0
            re = np.nonzero(outl)[0]
            errs(T, re, np.zeros(0, np.int64))
            c = chi2p(np.arange(n)).astype(np.float32)
            th = np.where(stereo, np.float32(p["chi2_stereo"]), np.float32(p["chi2_mono"])).astype(np.float32)
            outl = c > th
            pt_level = outl.astype(np.int64)
            nbad_total = int(outl.sum())
            if it == 2:
                pt_rob = False
            if n + m < 10:          # optimizer.edges().size()<10 :886
                break
            errs(T, np.zeros(0, np.int64), np.arange(m))      # e->computeError() for every line edge :894
            cl = chi2l(np.arange(m)).astype(np.float32)
            lbad = cl.astype(np.float64) > lgate
            ln_level = lbad.astype(np.int64)
            for e in range(m):                                 # right edge has the last word on mvbOutlierLines
                outl_ln[l_ln[e]] = lbad[e]
            if it == 2:
                ln_rob = False
        out["Tcw"][f] = T.to_Rt12()
        out["pt_outlier"][a:b] = outl
        out["ln_outlier"][la:lb] = outl_ln
        out["n_inliers"][f] = n - nbad_total
    return out
