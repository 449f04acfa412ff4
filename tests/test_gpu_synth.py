"""GPU: data synthesis + metrics kernels against fixtures produced by executing the reference's
own generator / driver code (tools/make_golden_synth.py) and against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import synth_oracle as so

pytestmark = pytest.mark.gpu

from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, MetricsAccumulator, UpdateParams, project_landmarks, synth_imu  # noqa: E402


def _dev(x, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda", dtype).contiguous()


def test_synth_imu_photorealistic_generator(golden_dir):
    z = np.load(os.path.join(golden_dir, "synth_photoreal.npz"))
    sig = z["sigmas"]
    noise = (z["normals"] * np.array([sig[0], sig[2], sig[1], sig[3]])[None, :, None]).reshape(-1, 12)
    B = 3                                                   # the same stream three times: batch indexing
    out = synth_imu(_dev(np.repeat(z["key_poses"][None], B, 0)), _dev(np.repeat(z["key_times"][None], B, 0)), 10, z["gravity"],
                    noise=_dev(np.repeat(noise[None], B, 0)), sign_fix=True)
    torch.cuda.synchronize()
    for b in range(B):
        gt, ny = out["imu_gt"][b].cpu().numpy(), out["imu_noisy"][b].cpu().numpy()
        assert gt.shape == z["imu_gt"].shape
        assert rel_err(gt[:, 0], z["imu_gt"][:, 0]) < 1e-14
        assert rel_err(gt[:, 1:4], z["imu_gt"][:, 1:4]) < 1e-9 and rel_err(gt[:, 4:], z["imu_gt"][:, 4:]) < 1e-9
        assert rel_err(ny[:, 1:4], z["imu_noisy"][:, 1:4]) < 1e-9 and rel_err(ny[:, 4:], z["imu_noisy"][:, 4:]) < 1e-9
        tr = z["transforms"][:, 1:].reshape(-1, 3, 4)
        po = out["poses"][b].cpu().numpy()
        assert rel_err(po[:, :9], tr[:, :, :3].reshape(-1, 9)) < 1e-12 and rel_err(po[:, 9:], tr[:, :, 3]) < 1e-12
        assert rel_err(out["times"][b].cpu().numpy(), z["transforms"][:, 0]) < 1e-14


def test_synth_imu_synthetic_generator(golden_dir):
    z = np.load(os.path.join(golden_dir, "synth_generator.npz"))
    P = np.vstack([z["poses"], z["last_pose"][None]])
    T = len(z["poses"])
    d = z["imu_draws"]                                      # gyro white, acc white, gyro walk, acc walk
    noise = np.concatenate([d[:, 1], d[:, 3], d[:, 0], d[:, 2]], 1)
    out = synth_imu(_dev(P[None]), _dev((np.arange(T + 1) * float(z["delta_t"]))[None]), 1, z["gravity"], noise=_dev(noise[None]),
                    accumulate_bias=True, sign_fix=False)
    ny = out["imu_noisy"][0].cpu().numpy()
    assert rel_err(ny[1:, 1:], z["imu_noisy"]) < 1e-9       # (the reference's gt stream is its noisy stream, see the CPU test)
    ref = so.synth_imu(P, np.arange(T + 1) * float(z["delta_t"]), 1, z["gravity"], None, True, False)[2]
    assert rel_err(out["imu_gt"][0].cpu().numpy()[:, 1:], ref) < 1e-12


def test_synth_imu_device_noise_statistics():
    """Draws made on the device: white noise + walk increments with the requested standard deviations, reproducible per seed."""
    B, K = 4, 400
    t = np.arange(K) * 0.05
    poses = np.zeros((B, K, 12))
    poses[:, :, [0, 4, 8]] = 1.0
    poses[:, :, 9] = np.sin(t)[None]
    sig = np.tile(np.array([[1e-2, 1e-3, 2e-3, 2e-4]]), (B, 1))
    a = synth_imu(_dev(poses), _dev(np.tile(t, (B, 1))), 10, (0, 0, -9.81), sigmas=_dev(sig), seed=99)
    b = synth_imu(_dev(poses), _dev(np.tile(t, (B, 1))), 10, (0, 0, -9.81), sigmas=_dev(sig), seed=99)
    c = synth_imu(_dev(poses), _dev(np.tile(t, (B, 1))), 10, (0, 0, -9.81), sigmas=_dev(sig), seed=100)
    assert torch.equal(a["imu_noisy"], b["imu_noisy"]) and not torch.equal(a["imu_noisy"], c["imu_noisy"])
    dn = (a["imu_noisy"] - a["imu_gt"])[:, 1:, 1:].cpu().numpy()      # [B, T, 6]: gyro, acc
    n = dn.shape[1]
    assert abs(dn[..., :3].std() / np.hypot(1e-3, 2e-4) - 1) < 0.02 and abs(dn[..., 3:].std() / np.hypot(1e-2, 2e-3) - 1) < 0.02
    assert abs(dn.mean()) < 5 * 1e-2 / np.sqrt(n * B * 6)
    assert not np.allclose(dn[0], dn[1])                               # filters draw from different streams


def test_project_landmarks(golden_dir):
    z = np.load(os.path.join(golden_dir, "synth_generator.npz"))
    poses, L = z["poses"], len(z["landmarks"])
    kp, ids, cnt, sc = project_landmarks(_dev(poses[None]), z["R_W_C"], z["t_W_C"], z["K"], int(z["width"]), int(z["height"]),
                                         _dev(z["landmarks"][None]), noise_uv=_dev(z["noise_uv"][None]), sigma_px=float(z["sigma_px"]),
                                         want_scores=True)
    kp, ids, cnt, sc = kp[0].cpu().numpy(), ids[0].cpu().numpy(), cnt[0].cpu().numpy(), sc[0].cpu().numpy()
    n = 0
    for i in range(len(poses)):
        vis = z["vis"][i]
        want = np.nonzero(vis)[0]
        assert cnt[i] == len(want) and np.array_equal(ids[i, :cnt[i]], want)          # id order, as the reference lists them
        if len(want):
            assert rel_err(kp[i, :cnt[i]], z["kp_noisy"][i][want]) < 1e-12
            assert np.allclose(sc[i, :cnt[i]], z["score"][i][want], rtol=0, atol=1e-15)
        n += len(want)
    assert n == int(z["vis"].sum()) > 3000
    # no noise supplied, sigma 0: the clean projections
    kp0, ids0, cnt0 = project_landmarks(_dev(poses[None]), z["R_W_C"], z["t_W_C"], z["K"], int(z["width"]), int(z["height"]),
                                        _dev(z["landmarks"][None]))
    i = 40
    want = np.nonzero(z["vis"][i])[0]
    assert rel_err(kp0[0, i, :len(want)].cpu().numpy(), z["kp_gt"][i][want]) < 1e-12


def test_metrics_stream(golden_dir):
    """main.py:312-391 step by step.  Translation metrics to 1e-9; the orientation-error angles are arccos of a value within
    ~1e-9 of 1 (condition number 1 / angle^2), so a different but equally valid evaluation of the same rotation products moves
    them by ~1e-16 / angle: they are compared in absolute terms (1e-9 rad before the division by the ground-truth angle)."""
    z = np.load(os.path.join(golden_dir, "metrics_main.npz"))
    n = len(z["gt_p"])
    B = 2
    flt = BatchedMSCKF(B, 4, 2, 1, UpdateParams())
    acc = MetricsAccumulator(B)
    gt = np.concatenate([z["gt_R"].reshape(n, 9), z["gt_p"]], 1)
    recs = []
    for i in range(n):
        st = np.concatenate([z["est_R"][i].flatten(), z["est_p"][i], np.zeros(9)])
        flt.imu_state.copy_(_dev(np.tile(st, (B, 1))))
        flt.P[:, :21, :21] = _dev(z["P"][i])
        recs.append(acc.step(_dev(np.tile(gt[i], (B, 1))), flt, record=True)[1].cpu().numpy())
    rec = np.array(recs)
    summ = acc.summary()[1].cpu().numpy()
    assert rel_err(rec[:, 0], z["rte"]) < 1e-9
    assert rel_err(rec[:, 2:5], z["ate"]) < 1e-9 and rel_err(rec[:, 5:8], z["aoe"]) < 1e-9
    assert rel_err(rec[:, 8:14], z["bounds"]) < 1e-12
    # orientation: compare the angles (roe x ground-truth relative angle) absolutely
    rel_o_gt = np.array([np.arccos(np.clip((np.trace((z["gt_R"][i - 1] if i else np.eye(3)).T @ z["gt_R"][i]) - 1) / 2, -1, 1))
                         for i in range(n)])
    assert np.max(np.abs(rec[:, 1] - z["roe"]) * np.maximum(rel_o_gt, 1e-12)) < 1e-9
    assert abs(summ[0] - np.sqrt(np.mean(z["rel_sq_pos"]))) < 1e-9 * summ[0]
    assert abs(summ[2] - np.mean(z["rte"])) < 1e-9 * summ[2] and abs(summ[3] - np.std(z["rte"])) < 1e-9 * summ[3]
    assert abs(summ[1] - np.sqrt(np.mean(z["rel_sq_ori"]))) < 1e-5 * summ[1]
    assert abs(summ[4] - np.mean(z["roe"])) < 1e-5 * summ[4] and abs(summ[5] - np.std(z["roe"])) < 1e-5 * summ[5]
    assert abs(summ[6] - np.sqrt(np.mean(np.sum(z["ate"] ** 2, 1)))) < 1e-9 * summ[6] and summ[7] == n
    assert np.max(np.abs(np.round(summ[[0, 2, 3]], 8) - z["summary"][[0, 2, 3]])) <= 1e-8        # the reference's rounded summary
