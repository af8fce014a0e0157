"""TEST INFRASTRUCTURE: an independent fp64 restatement of the kinematics the Cartesian / CoM producers need
(homogeneous 4x4 transforms, recursion over the tree), with Jacobians both analytic and by central finite
differences. The reference takes these quantities from XBot::ModelInterface, which cannot be built here:
this file is the oracle of tests/test_gpu_kinematics.py, and the row is "parity unpinned" (DESIGN.md)."""
import numpy as np


def _rot(axis, th):
    a = axis / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def fk(tree, q):
    """world transform (4x4) of every joint frame"""
    T = [None] * tree.n_joints
    for j in range(tree.n_joints):
        O = np.eye(4)
        O[:3, :3] = tree.origin[j, :9].reshape(3, 3)
        O[:3, 3] = tree.origin[j, 9:]
        M = np.eye(4)
        if tree.type[j] == 1:
            M[:3, :3] = _rot(tree.axis[j], q[tree.dof[j]])
        elif tree.type[j] == 2:
            M[:3, 3] = tree.axis[j] / np.linalg.norm(tree.axis[j]) * q[tree.dof[j]]
        P = np.eye(4) if tree.parent[j] < 0 else T[tree.parent[j]]
        T[j] = P @ O @ M
    return T


def pose(tree, q, distal, base=-1):
    T = fk(tree, q)
    return T[distal] if base < 0 else np.linalg.inv(T[base]) @ T[distal]


def com(tree, q):
    T = fk(tree, q)
    c = sum(tree.mass[j] * (T[j] @ np.append(tree.com[j], 1.0))[:3] for j in range(tree.n_joints))
    return c / tree.mass.sum()


def _vee(S):
    return np.array([S[2, 1] - S[1, 2], S[0, 2] - S[2, 0], S[1, 0] - S[0, 1]]) * 0.5


def jacobian_fd(tree, q, distal, base=-1, h=1e-6):
    """6 x n by central differences of the (relative) pose: rows linear then angular, in the base frame"""
    J = np.zeros((6, tree.n_dof))
    R0 = pose(tree, q, distal, base)[:3, :3]
    for i in range(tree.n_dof):
        dq = np.zeros(tree.n_dof); dq[i] = h
        Tp, Tm = pose(tree, q + dq, distal, base), pose(tree, q - dq, distal, base)
        J[:3, i] = (Tp[:3, 3] - Tm[:3, 3]) / (2 * h)
        J[3:, i] = _vee((Tp[:3, :3] - Tm[:3, :3]) / (2 * h) @ R0.T)
    return J


def com_jacobian_fd(tree, q, h=1e-6):
    J = np.zeros((3, tree.n_dof))
    for i in range(tree.n_dof):
        dq = np.zeros(tree.n_dof); dq[i] = h
        J[:, i] = (com(tree, q + dq) - com(tree, q - dq)) / (2 * h)
    return J


def _ancestors(tree, j):
    out = set()
    while j >= 0:
        out.add(j); j = tree.parent[j]
    return out


def jacobian(tree, q, distal, base=-1):
    """analytic: geometric Jacobian of the distal link, minus the motion of the base link, in the base frame"""
    T = fk(tree, q)
    pe = T[distal][:3, 3]
    ae, ab = _ancestors(tree, distal), (_ancestors(tree, base) if base >= 0 else set())
    J = np.zeros((6, tree.n_dof))
    for j in range(tree.n_joints):
        if tree.dof[j] < 0 or ((j in ae) == (j in ab)):
            continue
        a = T[j][:3, :3] @ (tree.axis[j] / np.linalg.norm(tree.axis[j]))
        col = np.concatenate([np.cross(a, pe - T[j][:3, 3]), a]) if tree.type[j] == 1 else np.concatenate([a, np.zeros(3)])
        J[:, tree.dof[j]] = -col if j in ab else col
    if base >= 0:
        Rb = T[base][:3, :3]
        J = np.vstack([Rb.T @ J[:3], Rb.T @ J[3:]])
    return J


def com_jacobian(tree, q):
    T = fk(tree, q)
    M = tree.mass.sum()
    J = np.zeros((3, tree.n_dof))
    cw = [(T[j] @ np.append(tree.com[j], 1.0))[:3] for j in range(tree.n_joints)]
    for j in range(tree.n_joints):
        if tree.dof[j] < 0:
            continue
        a = T[j][:3, :3] @ (tree.axis[j] / np.linalg.norm(tree.axis[j]))
        for k in range(tree.n_joints):
            if j in _ancestors(tree, k):
                J[:, tree.dof[j]] += tree.mass[k] / M * (np.cross(a, cw[k] - T[j][:3, 3]) if tree.type[j] == 1 else a)
    return J


def _quat(R):
    """(x, y, z, w), any sign"""
    w = np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    if w > 1e-3:
        return np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    vals, vecs = np.linalg.eigh((R + R.T) / 2)       # rotation by ~pi: the axis is the eigenvector of eigenvalue 1
    ax = vecs[:, np.argmax(vals)]
    ang = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    s = np.sign(_vee(R - R.T) @ ax) or 1.0
    return np.append(np.sin(ang / 2) * ax * s, np.cos(ang / 2))


def cartesian_b(T, Td, lam=1.0, gain=1.0, twist=None):
    """b of tasks::velocity::Cartesian (Cartesian.cpp:279-285) with cartesian_utils::computeCartesianError"""
    q, qd = _quat(T[:3, :3]), _quat(Td[:3, :3])
    if q @ qd < 0:
        q = -q
    eo = qd[3] * q[:3] - q[3] * qd[:3] + np.cross(qd[:3], q[:3])
    e = np.concatenate([Td[:3, 3] - T[:3, 3], -gain * eo])
    return (np.zeros(6) if twist is None else twist) + lam * e
