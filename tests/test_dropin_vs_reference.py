"""CPU: the HOST logic of the drop-in binding INTEGRATION.md shows -- `class MSCKF_B200(B200UpdateMixin, MSCKF)` over the
UNMODIFIED reference class -- driven through 60 camera frames of the reference's own feature_callback / imu_callback.

The device context is replaced by tests/fake_context.FakeContext (the five device calls answered by the CPU oracle), so this
checks what the GPU tests cannot see from golden cases: camera keys advance with imu.id (reference MSCKF.py:164, 255), the
window grows and is pruned like the reference's, tracks and last_camera_measurement stay what the reference's matcher
expects (MSCKF.py:772-778), the debug lists are filled (MSCKF.py:489-490), the covariance is uploaded once and then stays
resident.  Needs /root/reference (present in the build container, absent on the GPU box -> skipped there)."""
import os
import sys

import numpy as np
import pytest

from helpers import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference/src/msckf"), reason="the reference is not on this machine")


def _drive(f, ns, seq, per_frame):
    image = np.zeros((seq.height, seq.width, 3), np.uint8)
    frame_at = {fr.imu_index: fr for fr in seq.frames}
    for i in range(len(seq.imu)):
        row = seq.imu[i]
        f.imu_callback(ns.IMUMeasurement(row[0], row[1:4].copy(), row[4:7].copy()))
        fr = frame_at.get(i)
        if fr is not None:
            f.feature_callback(image, ns.CameraMeasurement(keypoints=fr.keypoints.copy(), descriptors=fr.descriptors.copy(),
                                                           scores=fr.scores.copy()))
            per_frame(f)


def _snapshot(f):
    imu = f.state.imu
    lcm = f.last_camera_measurement
    return dict(
        cam_keys=list(f.state.cameras.keys()), feat_keys=list(f.features.keys()), imu_id=imu.id,
        tracks={k: (list(ft.camera_indices), len(ft.keypoints), len(ft.descriptors), len(ft.lines), ft.tracked_for_n_frames,
                    ft.lost_for_n_frames) for k, ft in f.features.items()},
        idp={k: (np.array(ft.inverse_depth_point.m, float), float(ft.inverse_depth_point.rho)) for k, ft in f.features.items()},
        R=np.array(imu.T_W_Ii.R, float), t=np.array(imu.T_W_Ii.t, float), v=np.array(imu.v_W_Ii, float),
        bg=np.array(imu.gyroscope_bias, float), ba=np.array(imu.accelerometer_bias, float),
        P=np.array(f.state.covariance, float),
        cam_R=np.array([c.T_W_Ci.R for c in f.state.cameras.values()], float),
        cam_t=np.array([c.T_W_Ci.t for c in f.state.cameras.values()], float),
        lcm_idx=None if lcm is None else np.array(lcm.features_indices).copy(),
        lcm_desc=None if lcm is None else np.array(lcm.descriptors, float).copy(),
        n_world=len(f.estimated_world_points), n_cur=len(f.currently_processed_world_points),
        world=np.array(f.estimated_world_points[-5:], float),
        n_gate=f.number_of_residuals_discarded_for_gasting_test,
        aliased=all(ln.base is f.state.cameras[c].T_W_Ci.t for ft in f.features.values() for ln, c in zip(ft.lines, ft.camera_indices)),
        null_aliased=imu.T_W_Ii_null is imu.T_W_Ii and imu.v_W_Ii_null is imu.v_W_Ii)


@needs_reference
def test_binding_over_the_unmodified_reference_class():
    import ref_shim
    from fake_context import FakeContext
    from monocular_visual_inertial_msckf_b200 import B200UpdateMixin, synth
    import make_golden

    ns = ref_shim.load_reference()

    class MSCKF_B200(B200UpdateMixin, ns.MSCKF):          # INTEGRATION.md section 1, verbatim
        def _b200_make_context(self, F, M, N, params):    # ... except that the device is played by the CPU oracle
            return FakeContext(F, M, N, params)

    kw = dict(seed=42, n_frames=60, n_landmarks=40, noise_level="mid", p_drop=0.06, p_outlier=0.03, shuffle=True)
    fp = dict(min_parallax=2.0, max_number_of_camera_states=8, epipolar_rejection_threshold=1e9,
              homography_rejection_threshold=1e9, min_number_of_frames_to_be_tracked=3)
    seq = synth.make_sequence(**kw)
    params, _ = make_golden.make_params(ns, seq, **fp)

    ref_frames, got_frames = [], []
    _drive(ns.MSCKF(parameters=params, rr=None), ns, seq, lambda f: ref_frames.append(_snapshot(f)))
    FakeContext.totals.clear()
    dut = MSCKF_B200(parameters=params, rr=None)
    _drive(dut, ns, seq, lambda f: got_frames.append(_snapshot(f)))

    assert len(ref_frames) == len(got_frames) == 60
    worst = 0.0
    for k, (a, b) in enumerate(zip(ref_frames, got_frames)):
        where = f"frame {k}"
        assert b["cam_keys"] == a["cam_keys"], where              # keys are imu ids: they advance with every process_imu
        assert b["feat_keys"] == a["feat_keys"] and b["imu_id"] == a["imu_id"], where
        assert b["tracks"] == a["tracks"], where
        assert b["n_gate"] == a["n_gate"] and b["n_world"] == a["n_world"] and b["n_cur"] == a["n_cur"], where
        assert b["aliased"] and a["aliased"] and b["null_aliased"] == a["null_aliased"], where
        assert (b["lcm_idx"] is None) == (a["lcm_idx"] is None), where
        if a["lcm_idx"] is not None:
            assert np.array_equal(b["lcm_idx"], a["lcm_idx"]), where
            assert rel_err(b["lcm_desc"], a["lcm_desc"]) < 1e-12, where
        for key in ("R", "t", "v", "bg", "ba", "P", "cam_R", "cam_t", "world"):
            if a[key].size:
                e = rel_err(b[key], a[key])
                worst = max(worst, e)
                assert e < 1e-9, (where, key, e)
        for fid, (m, rho) in a["idp"].items():
            assert rel_err(b["idp"][fid][0], m) < 1e-9 and abs(b["idp"][fid][1] - rho) <= 1e-9 * abs(rho), (where, fid)
    # the window grew past one camera, reached its cap and was pruned; updates actually ran
    assert max(len(a["cam_keys"]) for a in ref_frames) >= 8 and len(set(tuple(a["cam_keys"]) for a in ref_frames)) > 30
    ctx = dut._b200_res.ctx
    # (the samples before the first camera frame are buffered and replayed by the reference's imu_callback)
    assert ctx.calls["update"] > 20 and ctx.calls["remove"] > 20 and ctx.calls["propagate"] == len(seq.imu)
    assert ctx.calls["augment"] == 59
    # residency: one upload of the covariance for the whole run, although it was read back after every frame
    assert dut._b200_res.p_uploads == 1
    print(f"drop-in binding vs reference over 60 frames: worst relative error {worst:.2e}, "
          f"{ctx.calls['update']} updates, {ctx.calls['remove']} camera removals")


def test_covariance_property_roundtrip():
    """Assigning state.covariance makes the host matrix authoritative (uploaded by the next call); reading it after a
    device call downloads the device copy; in-place edits of a downloaded matrix are refused."""
    from fake_context import FakeContext
    from monocular_visual_inertial_msckf_b200.msckf import (B200UpdateMixin, IMUMeasurement, MSCKF, MSCKFParameters)

    class F(MSCKF):
        def _b200_make_context(self, F, M, N, params):
            return FakeContext(F, M, N, params)

    f = F(MSCKFParameters(), device="cpu")
    P0 = np.diag(np.linspace(1e-4, 2e-4, 15))
    f.state.covariance = P0
    f.process_imu(IMUMeasurement(0.01, np.array([0.01, 0.02, -0.01]), np.array([0.1, 0.0, 9.81])))
    assert f.state.imu.id == 1
    P1 = f.state.covariance
    assert P1.shape == (15, 15) and not np.array_equal(P1, P0) and not P1.flags.writeable
    with pytest.raises(ValueError):
        P1[0, 0] = 1.0
    f.process_imu(IMUMeasurement(0.02, np.array([0.01, 0.02, -0.01]), np.array([0.1, 0.0, 9.81])))
    assert f._b200_res.p_uploads == 1 and f.state.imu.id == 2
    f.state.covariance = 2.0 * np.array(f.state.covariance)          # a new matrix: uploaded by the next call
    f.state_augmentation()
    assert f._b200_res.p_uploads == 2 and list(f.state.cameras.keys()) == [2] and f.state.covariance.shape == (21, 21)
    assert isinstance(f, B200UpdateMixin)


def test_reference_run_replayed_through_the_drop_in_class_host_logic(golden_dir):
    """The replay the GPU suite runs on the device (tests/test_gpu_dropin.py), here with the CPU stand-in: 763 calls through
    process_imu / state_augmentation / get_valid_features / update / remove_cameras, camera keys from imu.id, one upload
    of the covariance."""
    from dropin_replay import replay_trace_through_drop_in
    from fake_context import FakeContext
    from monocular_visual_inertial_msckf_b200.msckf import MSCKF

    class F(MSCKF):
        def _b200_make_context(self, F, M, N, params):
            return FakeContext(F, M, N, params)

    f, worst = replay_trace_through_drop_in(lambda p: F(p, device="cpu"), golden_dir)
    assert worst < 1e-9
