"""CPU oracle for SURVEY.md section 8f row 4 (data synthesis + metrics) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain NumPy restatements of what the reference computes (all citations relative to /root/reference):

    synth_imu            dataset/tools/dataset_generators/photorealistic_generator.py:145-216  (interp > 1: key poses are
                         interpolated, linear position + Slerp, before they are differentiated; noise = white + walk INCREMENT)
                         dataset/tools/dataset_generators/synthetic_generator.py:145-186       (interp == 1: poses at IMU rate,
                         accumulated random-walk bias)
    project_landmarks    synthetic_generator.py:189-223  (Camera.W2Ci / project_point, src/msckf/Camera.py:13-52)
    metrics_stream       main.py:312-391, 436-437 and the summary 440-445

Parity status: PINNED against tests/golden/synth_generator.npz, synth_photoreal.npz and metrics_main.npz, which
tools/make_golden_synth.py produced by executing the reference's own code in the build container
(tests/test_synth_oracle_vs_golden.py).  Quaternions follow scipy.spatial.transform.Rotation (x, y, z, w; from_matrix picks
the largest of the trace / diagonal entries), because the reference's R2quaternion / Slerp / R2euler are thin wrappers of it
(src/utils/geometry.py:132-194).  Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module."""
import numpy as np


# ---------------------------------------------------------------------------------------------------------------------
# rotation helpers (scipy.spatial.transform conventions)
# ---------------------------------------------------------------------------------------------------------------------
def quat_from_matrix(R):
    """Rotation.from_matrix(R).as_quat() for an orthonormal R: (x, y, z, w), built from the largest of (R00, R11, R22, trace)."""
    R = np.asarray(R, float)
    dec = np.array([R[0, 0], R[1, 1], R[2, 2], R[0, 0] + R[1, 1] + R[2, 2]])
    c = int(np.argmax(dec))
    q = np.zeros(4)
    if c != 3:
        i, j, k = c, (c + 1) % 3, (c + 2) % 3
        q[i] = 1 - dec[3] + 2 * R[i, i]
        q[j] = R[j, i] + R[i, j]
        q[k] = R[k, i] + R[i, k]
        q[3] = R[k, j] - R[j, k]
    else:
        q[0] = R[2, 1] - R[1, 2]
        q[1] = R[0, 2] - R[2, 0]
        q[2] = R[1, 0] - R[0, 1]
        q[3] = 1 + dec[3]
    return q / np.linalg.norm(q)


def quat_to_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def slerp_matrix(R0, R1, t):
    """Slerp([0, 1], [R0, R1])(t).as_matrix(): R0 exp(t log(R0^T R1))."""
    q0, q1 = quat_from_matrix(R0), quat_from_matrix(R1)
    d = quat_mul(np.array([-q0[0], -q0[1], -q0[2], q0[3]]), q1)
    if d[3] < 0:
        d = -d                                   # as_rotvec works on the w >= 0 representative
    nv = np.linalg.norm(d[:3])
    ang = 2 * np.arctan2(nv, d[3])
    rotvec = d[:3] * (ang / nv) if nv > 1e-300 else np.zeros(3)
    a = t * ang
    ax = rotvec / ang if ang > 0 else np.zeros(3)
    dq = np.concatenate([np.sin(a / 2) * ax, [np.cos(a / 2)]])
    return quat_to_matrix(quat_mul(q0, dq))


def euler_XYZ(R):
    """R2euler(R) = Rotation.from_matrix(R).as_euler('XYZ') (src/utils/geometry.py:132-143): intrinsic X-Y-Z angles."""
    R = np.asarray(R, float)
    return np.array([np.arctan2(-R[1, 2], R[2, 2]), np.arcsin(np.clip(R[0, 2], -1.0, 1.0)), np.arctan2(-R[0, 1], R[0, 0])])


# ---------------------------------------------------------------------------------------------------------------------
# IMU synthesis from ground-truth poses
# ---------------------------------------------------------------------------------------------------------------------
def _omega(R_prev, R_cur, dt, sign_fix):
    """(2 / dt) vec(q_prev^* q_cur) with scipy's quaternions (photorealistic_generator.py:178-187, synthetic_generator.py:158-164)."""
    qp, qc = quat_from_matrix(R_prev), quat_from_matrix(R_cur)
    if sign_fix and np.dot(qp, qc) < 0:
        qc = -qc
    q1 = np.array([qp[3], qp[0], qp[1], qp[2]])
    q2 = np.array([qc[3], qc[0], qc[1], qc[2]])
    return (2 / dt) * np.array([q1[0] * q2[1] - q1[1] * q2[0] - q1[2] * q2[3] + q1[3] * q2[2],
                                q1[0] * q2[2] + q1[1] * q2[3] - q1[2] * q2[0] - q1[3] * q2[1],
                                q1[0] * q2[3] - q1[1] * q2[2] + q1[2] * q2[1] - q1[3] * q2[0]])


def synth_imu(key_poses, key_times, interp, gravity, noise, accumulate_bias, sign_fix):
    """key_poses [K, 12] (R row-major, t), key_times [K].  interp = points per key interval (the reference's
    np.linspace(0, 1, 10) -> 10, i.e. 9 samples per interval; 1 = the poses are already at IMU rate).
    noise [T, 4, 3] already scaled: acc white, acc walk increment, gyro white, gyro walk increment (or None).
    Returns poses [T + 1, 12], times [T + 1], imu_gt [T + 1, 6] (w, a; row 0 zero), imu_noisy [T + 1, 6]."""
    key_poses = np.asarray(key_poses, float)
    K = len(key_poses)
    g = np.asarray(gravity, float)
    sub = max(interp - 1, 1)
    poses, times = [key_poses[0].copy()], [float(key_times[0])]
    for i in range(1, K):
        R0, p0 = key_poses[i - 1][:9].reshape(3, 3), key_poses[i - 1][9:]
        R1, p1 = key_poses[i][:9].reshape(3, 3), key_poses[i][9:]
        if interp > 1:
            ts = np.linspace(0, 1, interp)
            # the reference restarts every interval from slerp(0), a quaternion round trip of the key pose (:165-166)
            poses[-1] = np.concatenate([slerp_matrix(R0, R1, 0.0).flatten(), p0])
            for j in range(1, interp):
                t = ts[j]
                poses.append(np.concatenate([slerp_matrix(R0, R1, t).flatten(), (1 - t) * p0 + t * p1]))
                times.append((1 - t) * key_times[i - 1] + t * key_times[i])
        else:
            poses.append(key_poses[i].copy())
            times.append(float(key_times[i]))
    # (with interpolation the pose a sample differentiates FROM is the restarted one; rebuild the exact sequence per interval)
    T = sub * (K - 1)
    imu_gt = np.zeros((T + 1, 6))
    imu_noisy = np.zeros((T + 1, 6))
    out_poses = np.zeros((T + 1, 12))
    out_poses[0] = key_poses[0]
    v_prev = np.zeros(3)
    bias_a, bias_g = np.zeros(3), np.zeros(3)
    k = 0
    for i in range(1, K):
        R0, p0 = key_poses[i - 1][:9].reshape(3, 3), key_poses[i - 1][9:]
        R1, p1 = key_poses[i][:9].reshape(3, 3), key_poses[i][9:]
        if interp > 1:
            ts = np.linspace(0, 1, interp)
            Rs = [slerp_matrix(R0, R1, t) for t in ts]
            ps = [(1 - t) * p0 + t * p1 for t in ts]
            tt = [(1 - t) * key_times[i - 1] + t * key_times[i] for t in ts]
        else:
            Rs, ps, tt = [R0, R1], [p0, p1], [key_times[i - 1], key_times[i]]
        for j in range(1, len(Rs)):
            k += 1
            dt = tt[j] - tt[j - 1]
            v = (ps[j] - ps[j - 1]) / dt
            acc = Rs[j - 1].T @ ((v - v_prev) / dt + g)
            om = _omega(Rs[j - 1], Rs[j], dt, sign_fix)
            imu_gt[k] = np.concatenate([om, acc])
            if noise is not None:
                na, nba, ng, nbg = noise[k - 1]
                if accumulate_bias:                      # synthetic_generator.py:168-181
                    bias_g = bias_g + nbg
                    bias_a = bias_a + nba
                    imu_noisy[k] = np.concatenate([om + ng + bias_g, acc + na + bias_a])
                else:                                    # photorealistic_generator.py:189-198: the increment, not the sum
                    imu_noisy[k] = np.concatenate([om + nbg + ng, acc + nba + na])
            else:
                imu_noisy[k] = imu_gt[k]
            out_poses[k] = np.concatenate([Rs[j].flatten(), ps[j]])
            v_prev = v
    return out_poses, np.array(times), imu_gt, imu_noisy


# ---------------------------------------------------------------------------------------------------------------------
# landmark projection
# ---------------------------------------------------------------------------------------------------------------------
def project_landmarks(body_pose, R_I_C, t_I_C, K, width, height, landmarks, noise_uv):
    """One pose: body_pose [12]; returns (visible [L] bool, kp_gt [L, 2], kp_noisy [L, 2]) (synthetic_generator.py:189-223)."""
    R, t = body_pose[:9].reshape(3, 3), body_pose[9:]
    Rc, tc = R @ R_I_C, R @ t_I_C + t
    L = len(landmarks)
    vis = np.zeros(L, bool)
    kp = np.zeros((L, 2))
    kpn = np.zeros((L, 2))
    for j in range(L):
        c = Rc.T @ (landmarks[j] - tc)
        if c[2] <= 0:
            continue
        im = K @ c
        im = im[:2] / im[2]
        if im[0] < 0 or im[0] >= width or im[1] < 0 or im[1] >= height:
            continue
        vis[j] = True
        kp[j] = im
        kpn[j] = im + (noise_uv[j] if noise_uv is not None else 0.0)
    return vis, kp, kpn


# ---------------------------------------------------------------------------------------------------------------------
# metrics
# ---------------------------------------------------------------------------------------------------------------------
def metrics_stream(gt_R, gt_p, est_R, est_p, P):
    """main.py:312-391 + 436-445 over a whole stream.  Returns dict(rte, roe (smoothed, per step), ate [n, 3], aoe [n, 3],
    bounds [n, 6], rel_sq_pos, rel_sq_ori, summary [6] = rmse_position, rmse_orientation, mean/std rte, mean/std roe)."""
    n = len(gt_p)

    # Isometry3D.inv / __mul__ go through 4x4 homogeneous matrices and np.linalg.inv (src/utils/geometry.py:31-38); the
    # orientation errors below are arccos of a value within 1e-8 of 1, where the last bits of those products matter, so the
    # same arithmetic is used here
    def T4(R, p):
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = p
        return T

    def inv_mul(A, B):
        Ai = np.linalg.inv(A)
        Ai = T4(Ai[:3, :3], Ai[:3, 3])
        return Ai @ B

    T0g, T0e = np.eye(4), np.eye(4)
    dq_p, dq_o = [], []
    out = dict(rte=np.zeros(n), roe=np.zeros(n), ate=np.zeros((n, 3)), aoe=np.zeros((n, 3)), bounds=np.zeros((n, 6)),
               rel_sq_pos=np.zeros(n), rel_sq_ori=np.zeros(n))
    ang = lambda R: np.arccos(np.clip((np.trace(R) - 1) / 2.0, -1.0, 1.0))  # noqa: E731
    from scipy.spatial.transform import Rotation
    for i in range(n):
        T1g, T1e = T4(gt_R[i], gt_p[i]), T4(est_R[i], est_p[i])
        Trg = inv_mul(T0g, T1g)                                       # T_I0_I1_gt
        Tre = inv_mul(T0e, T1e)                                       # T_I0_I1_est
        Trg, Tre = T4(Trg[:3, :3], Trg[:3, 3]), T4(Tre[:3, :3], Tre[:3, 3])
        rel_t_gt, rel_o_gt = np.linalg.norm(Trg[:3, 3]), ang(Trg[:3, :3])
        Ta = inv_mul(T1g, T1e)                                        # abs_T_error
        out["ate"][i] = Ta[:3, 3]
        out["aoe"][i] = Rotation.from_matrix(Ta[:3, :3]).as_euler("XYZ")
        out["bounds"][i] = np.concatenate([3 * np.sqrt(np.diag(P[i][12:15, 12:15])), 3 * np.sqrt(np.diag(P[i][:3, :3]))])
        Tr = inv_mul(Trg, Tre)                                        # rel_T_error
        rte = np.linalg.norm(Tr[:3, 3]) / rel_t_gt if rel_t_gt != 0 else 0
        roe = ang(Tr[:3, :3]) / rel_o_gt if rel_o_gt != 0 else 0
        rte_s = 0.001 * rte + 0.999 * np.mean(dq_p) if len(dq_p) > 0 and rte > 10 * np.mean(dq_p) else rte
        roe_s = 0.001 * roe + 0.999 * np.mean(dq_o) if len(dq_o) > 0 and roe > 10 * np.mean(dq_o) else roe
        dq_p = (dq_p + [rte_s])[-10:]
        dq_o = (dq_o + [roe_s])[-10:]
        out["rel_sq_pos"][i], out["rel_sq_ori"][i] = rte ** 2, roe ** 2
        out["rte"][i], out["roe"][i] = rte_s, roe_s
        T0g, T0e = T1g, T1e
    out["summary"] = np.array([np.sqrt(np.mean(out["rel_sq_pos"])), np.sqrt(np.mean(out["rel_sq_ori"])), np.mean(out["rte"]),
                               np.std(out["rte"]), np.mean(out["roe"]), np.std(out["roe"])])
    return out
