/*
 * sot_kinematics.cuh -- on-device producers of the Cartesian and CoM velocity tasks:
 * from a batch of joint positions q [B, n] and a kinematic tree (URDF-derived table, constant over the
 * batch) straight into the instance blobs,
 *     Cartesian:  A = J(distal) or the relative Jacobian of distal w.r.t. base (6 x n, linear rows first),
 *                 b = v_des + lambda [ p_des - p ; -gain * e_o(q, q_des) ]   src/tasks/velocity/Cartesian.cpp:68-105, 279-285
 *                 e_o = qd.w eps - q.w epsd + epsd x eps, short path          include/OpenSoT/utils/cartesian_utils.h:144-160,
 *                                                                             src/utils/cartesian_utils.cpp:79-95
 *     CoM:        A = J_com (3 x n), b = v_des + lambda (p_des - com)         src/tasks/velocity/CoM.cpp:59-74, 145-149
 * so that only joint states cross PCIe. The reference obtains J, poses and the CoM from XBot::ModelInterface
 * (absent here: parity of this file is pinned by an independent fp64 forward-kinematics restatement and by
 * finite differences, tests/test_gpu_kinematics.py, not by the reference itself).
 *
 * One warp per instance. World transforms of every joint frame live in shared memory. First every joint's transform
 * relative to its parent (lanes over ALL joints: the only place the tree tables are read from global memory on the
 * dependent path), then forward kinematics depth by depth from shared memory alone (lanes over the joints of one
 * depth), the subtree masses / first moments for the CoM Jacobian bottom-up, then lanes over the degrees of freedom
 * write the Jacobian columns.
 * Fixed-base trees of revolute / continuous / prismatic / fixed joints (Euclidean joint spaces, like
 * sot_posture_kernel); a floating base needs the quaternion joint of XBot and is rejected by the host side.
 */
#ifndef OSB_SOT_KINEMATICS_CUH_
#define OSB_SOT_KINEMATICS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include "sot_params.h"

#define KIN_WARPS 4     /* instances per CTA */

namespace {

struct kin3 { double x, y, z; };
__device__ __forceinline__ kin3 kcross(const kin3& a, const kin3& b)
{
    return kin3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
/* y = R x (R row-major 3x3 at T[0..8]) */
__device__ __forceinline__ kin3 krot(const double* T, const kin3& v)
{
    return kin3{T[0] * v.x + T[1] * v.y + T[2] * v.z, T[3] * v.x + T[4] * v.y + T[5] * v.z, T[6] * v.x + T[7] * v.y + T[8] * v.z};
}
__device__ __forceinline__ kin3 krotT(const double* T, const kin3& v)
{
    return kin3{T[0] * v.x + T[3] * v.y + T[6] * v.z, T[1] * v.x + T[4] * v.y + T[7] * v.z, T[2] * v.x + T[5] * v.y + T[8] * v.z};
}
/* unit quaternion (x, y, z, w) of a rotation matrix: Eigen::Quaterniond(Matrix3d), the construction the reference uses */
__device__ __forceinline__ void kquat(const double* R, double q[4])
{
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0.0) {
        double t = sqrt(tr + 1.0);
        q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        /* largest diagonal entry picks the branch; written out per branch: with a run-time index the matrix and the
         * quaternion went to local memory (192 bytes of stack) */
        int i = (R[4] > R[0]) ? 1 : 0;
        if (R[8] > (i ? R[4] : R[0])) i = 2;
        if (i == 0) {
            double t = sqrt(R[0] - R[4] - R[8] + 1.0);
            q[0] = 0.5 * t; t = 0.5 / t;
            q[3] = (R[7] - R[5]) * t; q[1] = (R[3] + R[1]) * t; q[2] = (R[6] + R[2]) * t;
        } else if (i == 1) {
            double t = sqrt(R[4] - R[8] - R[0] + 1.0);
            q[1] = 0.5 * t; t = 0.5 / t;
            q[3] = (R[2] - R[6]) * t; q[2] = (R[7] + R[5]) * t; q[0] = (R[1] + R[3]) * t;
        } else {
            double t = sqrt(R[8] - R[0] - R[4] + 1.0);
            q[2] = 0.5 * t; t = 0.5 / t;
            q[3] = (R[3] - R[1]) * t; q[0] = (R[2] + R[6]) * t; q[1] = (R[5] + R[7]) * t;
        }
    }
}

}  // namespace

extern "C" __global__ void __launch_bounds__(32 * KIN_WARPS) sot_kinematics_kernel(const __grid_constant__ kin_params g)
{
    extern __shared__ __align__(16) double kin_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int inst = blockIdx.x * KIN_WARPS + wid;
    if (inst >= g.batch) return;
    const int J = g.n_joints, n = g.n_dof;
    /* per warp: 15 J doubles + J bytes (64 joints: 7.7 KB, so that seven CTAs = 28 warps fit one SM and 4096 states are ONE
     * wave of the 148 SMs; with the subtree masses -- constants of the tree, k_msub -- and the joint positions in shared
     * memory too it was 8.7 KB: 24 warps, 1.15 waves) */
    double* T = kin_smem + (size_t)wid * g.smem_per_warp;   /* [J][12]: rotation row-major, translation */
    double* S = T + (size_t)J * 12;                         /* [J][3]: first moment of the subtree      */
    uint8_t* pars = reinterpret_cast<uint8_t*>(S + (size_t)J * 3);     /* [J]: parent of every joint, 255 = none */
    const double* q = g.q + (size_t)inst * g.q_stride;
    /* joint positions in registers: lane i holds q[i] and q[i + 32] (n <= 64), read by shuffles */
    const double q_lo = (lane < n) ? q[lane] : 0.0, q_hi = (lane + 32 < n) ? q[lane + 32] : 0.0;
    /* lane d keeps the first joint of depth d (and d + 32): the depth loop below reads no table */
    const int ds0 = (lane <= g.n_depths) ? g.depth_start[lane] : J;
    const int ds1 = (lane + 32 <= g.n_depths) ? g.depth_start[lane + 32] : J;
    __syncwarp();
    /* ---- every joint's transform relative to its parent, all joints at once (lanes over joints): the tree tables are
     * read from global memory here and nowhere else on the dependent path (one latency for the whole tree; read inside
     * the depth loop they cost ~700 cycles per depth and the kernel took 74 us per 4096 states, profiles/r02_ncu_kinematics) ---- */
    for (int j0 = 0; j0 < J; j0 += 32) {
        /* every small loop below is fully unrolled: the 3x3 blocks must live in registers (with runtime loop
         * bounds they went to local memory and the kernel took 81 us per 4096 states) */
        const bool valid = j0 + lane < J;
        const int j = valid ? j0 + lane : J - 1;             /* all lanes stay for the shuffles; extra lanes shadow the last joint */
        const double2* O2 = reinterpret_cast<const double2*>(g.origin + (size_t)j * 12);
        double Rl[9], pl[3];
        {
            const double2 o0 = O2[0], o1 = O2[1], o2 = O2[2], o3 = O2[3], o4 = O2[4], o5 = O2[5];
            Rl[0] = o0.x; Rl[1] = o0.y; Rl[2] = o1.x; Rl[3] = o1.y; Rl[4] = o2.x; Rl[5] = o2.y; Rl[6] = o3.x; Rl[7] = o3.y;
            Rl[8] = o4.x; pl[0] = o4.y; pl[1] = o5.x; pl[2] = o5.y;
        }
        const int type = g.type[j], dofj = g.dof[j];
        const double ax = g.axis[3 * j], ay = g.axis[3 * j + 1], az = g.axis[3 * j + 2];
        if (valid) pars[j] = (uint8_t)g.parent[j];            /* (-1 -> 255) */
        const int dsrc = dofj >= 0 ? dofj : 0;
        const double qa = __shfl_sync(0xffffffffu, q_lo, dsrc & 31), qb = __shfl_sync(0xffffffffu, q_hi, dsrc & 31);
        const double qj = dsrc < 32 ? qa : qb;
        if (type == 1) {           /* revolute: origin * Rot(axis, q) (Rodrigues) */
            double s, c;
            sincos(qj, &s, &c);
            const double v = 1.0 - c;
            const double M[9] = {c + ax * ax * v, ax * ay * v - az * s, ax * az * v + ay * s,
                                 ay * ax * v + az * s, c + ay * ay * v, ay * az * v - ax * s,
                                 az * ax * v - ay * s, az * ay * v + ax * s, c + az * az * v};
            double R2[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) R2[3 * r + cc] = Rl[3 * r] * M[cc] + Rl[3 * r + 1] * M[3 + cc] + Rl[3 * r + 2] * M[6 + cc];
#pragma unroll
            for (int i = 0; i < 9; ++i) Rl[i] = R2[i];
        } else if (type == 2) {    /* prismatic: translation along the axis in the joint frame */
            const double th = qj;
            const kin3 a = krot(Rl, kin3{ax, ay, az});
            pl[0] += th * a.x; pl[1] += th * a.y; pl[2] += th * a.z;
        }
        if (valid) {
            double2* Tj2 = reinterpret_cast<double2*>(T + (size_t)j * 12);
            Tj2[0] = make_double2(Rl[0], Rl[1]); Tj2[1] = make_double2(Rl[2], Rl[3]); Tj2[2] = make_double2(Rl[4], Rl[5]);
            Tj2[3] = make_double2(Rl[6], Rl[7]); Tj2[4] = make_double2(Rl[8], pl[0]); Tj2[5] = make_double2(pl[1], pl[2]);
        }
    }
    __syncwarp();
    /* ---- forward kinematics, depth by depth from shared memory only: T[j] <- T[parent] o T[j] (depth 0 is already in
     * the world frame) ---- */
    for (int d = 1; d < g.n_depths; ++d) {
        const int jb = (d < 64) ? __shfl_sync(0xffffffffu, d < 32 ? ds0 : ds1, d & 31) : g.depth_start[d];
        const int je = (d + 1 < 64) ? __shfl_sync(0xffffffffu, d + 1 < 32 ? ds0 : ds1, (d + 1) & 31) : g.depth_start[d + 1];
        for (int j = jb + lane; j < je; j += 32) {
            double2* Tj2 = reinterpret_cast<double2*>(T + (size_t)j * 12);
            const double2* Tp2 = reinterpret_cast<const double2*>(T + (size_t)pars[j] * 12);   /* (depth >= 1: a parent exists) */
            double Rl[9], pl[3], Tp[12];
            {
                const double2 a0 = Tj2[0], a1 = Tj2[1], a2 = Tj2[2], a3 = Tj2[3], a4 = Tj2[4], a5 = Tj2[5];
                Rl[0] = a0.x; Rl[1] = a0.y; Rl[2] = a1.x; Rl[3] = a1.y; Rl[4] = a2.x; Rl[5] = a2.y; Rl[6] = a3.x; Rl[7] = a3.y;
                Rl[8] = a4.x; pl[0] = a4.y; pl[1] = a5.x; pl[2] = a5.y;
                const double2 b0 = Tp2[0], b1 = Tp2[1], b2 = Tp2[2], b3 = Tp2[3], b4 = Tp2[4], b5 = Tp2[5];
                Tp[0] = b0.x; Tp[1] = b0.y; Tp[2] = b1.x; Tp[3] = b1.y; Tp[4] = b2.x; Tp[5] = b2.y; Tp[6] = b3.x; Tp[7] = b3.y;
                Tp[8] = b4.x; Tp[9] = b4.y; Tp[10] = b5.x; Tp[11] = b5.y;
            }
            double Rw[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) Rw[3 * r + cc] = Tp[3 * r] * Rl[cc] + Tp[3 * r + 1] * Rl[3 + cc] + Tp[3 * r + 2] * Rl[6 + cc];
            const kin3 pw = krot(Tp, kin3{pl[0], pl[1], pl[2]});
            Tj2[0] = make_double2(Rw[0], Rw[1]); Tj2[1] = make_double2(Rw[2], Rw[3]); Tj2[2] = make_double2(Rw[4], Rw[5]);
            Tj2[3] = make_double2(Rw[6], Rw[7]); Tj2[4] = make_double2(Rw[8], Tp[9] + pw.x);
            Tj2[5] = make_double2(Tp[10] + pw.y, Tp[11] + pw.z);
        }
        __syncwarp();
    }
    const int n_roots = __shfl_sync(0xffffffffu, ds0, 1);     /* joints of depth 0 come first */
    /* ---- first moment of every joint's subtree (for the CoM Jacobian), bottom-up; lanes 0..2 own the three
     * components, joints in reverse topological order (parents from shared memory); the subtree MASSES do not depend on
     * the configuration: g.msub, summed once on the host in the same order ---- */
    if (g.need_com) {
        for (int j = lane; j < J; j += 32) {
            const double* Tj = T + (size_t)j * 12;
            const kin3 cw = krot(Tj, kin3{g.com[3 * j], g.com[3 * j + 1], g.com[3 * j + 2]});
            const double m = g.mass[j];
            S[3 * j] = m * (Tj[9] + cw.x); S[3 * j + 1] = m * (Tj[10] + cw.y); S[3 * j + 2] = m * (Tj[11] + cw.z);
        }
        __syncwarp();
        if (lane < 3)
            for (int j = J - 1; j >= n_roots; --j) S[3 * pars[j] + lane] += S[3 * j + lane];
        __syncwarp();
    }
    /* ---- tasks ---- */
    double* blob = g.blobs + (size_t)inst * g.stride;
    for (int t = 0; t < g.n_tasks; ++t) {
        const kin_task& K = g.task[t];
        double* A = blob + K.off_A;
        double* b = blob + K.off_b;
        if (K.kind == OSB_KIN_COM) {
            double mtot = 0.0, mx = 0.0, my = 0.0, mz = 0.0;
            for (int j = 0; j < n_roots; ++j) { mtot += g.msub[j]; mx += S[3 * j]; my += S[3 * j + 1]; mz += S[3 * j + 2]; }
            const double inv = 1.0 / mtot;
            for (int j = lane; j < J; j += 32) {
                const int dj = g.dof[j];
                if (dj < 0) continue;
                const double* Tj = T + (size_t)j * 12;
                const kin3 a = krot(Tj, kin3{g.axis[3 * j], g.axis[3 * j + 1], g.axis[3 * j + 2]});
                const double msj = g.msub[j];
                const double ms = msj * inv;                        /* subtree mass fraction */
                kin3 col;
                if (g.type[j] == 1) {
                    const double is = msj > 0.0 ? 1.0 / msj : 0.0;
                    const kin3 r{S[3 * j] * is - Tj[9], S[3 * j + 1] * is - Tj[10], S[3 * j + 2] * is - Tj[11]};
                    col = kcross(a, r);
                    col.x *= ms; col.y *= ms; col.z *= ms;
                } else col = kin3{a.x * ms, a.y * ms, a.z * ms};
                A[dj] = col.x; A[n + dj] = col.y; A[2 * n + dj] = col.z;
            }
            if (lane < 3) {
                const double com = (lane == 0 ? mx : (lane == 1 ? my : mz)) * inv;
                const double des = K.desired ? K.desired[(size_t)inst * K.desired_stride + lane] : com;
                const double vel = K.desired_vel ? K.desired_vel[(size_t)inst * K.desired_vel_stride + lane] : 0.0;
                b[lane] = vel + K.lambda * (des - com);
            }
        } else {
            const double* Te = T + (size_t)K.distal * 12;
            const double* Tb = K.base >= 0 ? T + (size_t)K.base * 12 : nullptr;
            const kin3 pe{Te[9], Te[10], Te[11]};
            /* columns: joint j moves the distal link iff it is one of its ancestors (bit j of anc_distal), and the
             * base link likewise; the relative Jacobian is R_b' [ (v_e - v_b - w_b x (p_e - p_b)) ; (w_e - w_b) ] */
            for (int j = lane; j < J; j += 32) {
                const int dj = g.dof[j];
                if (dj < 0) continue;
                const bool on_e = (K.anc_distal[j >> 6] >> (j & 63)) & 1ull;
                const bool on_b = Tb && ((K.anc_base[j >> 6] >> (j & 63)) & 1ull);
                kin3 v{0, 0, 0}, w{0, 0, 0};
                if (on_e != on_b) {
                    const double* Tj = T + (size_t)j * 12;
                    const kin3 a = krot(Tj, kin3{g.axis[3 * j], g.axis[3 * j + 1], g.axis[3 * j + 2]});
                    if (g.type[j] == 1) { v = kcross(a, kin3{pe.x - Tj[9], pe.y - Tj[10], pe.z - Tj[11]}); w = a; }
                    else v = a;
                    if (on_b) { v = kin3{-v.x, -v.y, -v.z}; w = kin3{-w.x, -w.y, -w.z}; }
                }
                if (Tb) { v = krotT(Tb, v); w = krotT(Tb, w); }
                A[dj] = v.x; A[n + dj] = v.y; A[2 * n + dj] = v.z;
                A[3 * n + dj] = w.x; A[4 * n + dj] = w.y; A[5 * n + dj] = w.z;
            }
            if (lane == 0) {
                /* actual pose of the distal link (in the base frame when a base link is given) */
                double R[9], p[3];
                if (Tb) {
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) R[3 * r + cc] = Tb[r] * Te[cc] + Tb[3 + r] * Te[3 + cc] + Tb[6 + r] * Te[6 + cc];
                    const kin3 dp = krotT(Tb, kin3{pe.x - Tb[9], pe.y - Tb[10], pe.z - Tb[11]});
                    p[0] = dp.x; p[1] = dp.y; p[2] = dp.z;
                } else {
#pragma unroll
                    for (int i = 0; i < 9; ++i) R[i] = Te[i];
                    p[0] = pe.x; p[1] = pe.y; p[2] = pe.z;
                }
                double e[6] = {0, 0, 0, 0, 0, 0};
                if (K.desired) {
                    const double* D = K.desired + (size_t)inst * K.desired_stride;
                    double qa[4], qd[4];
                    kquat(R, qa); kquat(D, qd);
                    if (qa[0] * qd[0] + qa[1] * qd[1] + qa[2] * qd[2] + qa[3] * qd[3] < 0.0) { qa[0] = -qa[0]; qa[1] = -qa[1]; qa[2] = -qa[2]; qa[3] = -qa[3]; }
                    const kin3 eps{qa[0], qa[1], qa[2]}, epsd{qd[0], qd[1], qd[2]};
                    const kin3 cr = kcross(epsd, eps);
                    const double eo[3] = {qd[3] * eps.x - qa[3] * epsd.x + cr.x, qd[3] * eps.y - qa[3] * epsd.y + cr.y, qd[3] * eps.z - qa[3] * epsd.z + cr.z};
#pragma unroll
                    for (int i = 0; i < 3; ++i) { e[i] = D[9 + i] - p[i]; e[3 + i] = -K.orientation_gain * eo[i]; }
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const double vel = K.desired_vel ? K.desired_vel[(size_t)inst * K.desired_vel_stride + i] : 0.0;
                    b[i] = vel + K.lambda * e[i];
                }
            }
        }
    }
}

#endif
