"""TEST INFRASTRUCTURE: a CPU stand-in for BatchedMSCKF(1, F, M, N) whose five device calls (front, update, propagate,
augment, remove_cameras) are answered by the CPU oracle.  It exists so that the HOST logic of the drop-in mixin
(monocular_visual_inertial_msckf_b200.msckf.B200UpdateMixin: key handling, object aliasing, track surgery, residency of the
covariance, last_camera_measurement upkeep) can be run against the UNMODIFIED reference class without a GPU.  It proves
nothing about the kernels -- the -m gpu tests do that -- and nothing in the product imports it."""
import numpy as np
import torch

from oracle import msckf_oracle as o
from monocular_visual_inertial_msckf_b200 import _lib


class FakeContext:
    totals = {}          # calls over every context created since the last reset (contexts are re-created when they grow)

    def __init__(self, F, M, N, params):
        self.B, self.F, self.M, self.N = 1, int(F), int(M), int(N)
        self.ldp = 15 + 6 * self.N
        self.params = params
        self.device = torch.device("cpu")
        f64, i32 = torch.float64, torch.int32
        z = lambda *s, dtype=f64: torch.zeros(*s, dtype=dtype)  # noqa: E731
        self.P = z(1, self.ldp, self.ldp)
        self.imu_state, self.imu_null, self.imu_time, self.Qc_diag = z(1, 21), z(1, 15), z(1), z(1, 4)
        self.cam_R, self.cam_t, self.samples_stage = z(1, N, 9), z(1, N, 3), z(1, 16, 7)
        self.n_cams = z(1, dtype=i32)
        self.feat_len = z(1, F, dtype=i32)
        self.obs_uv, self.obs_dir, self.obs_conf = z(1, F, M, 2), z(1, F, M, 3), z(1, F, M)
        self.obs_slot = z(1, F, M, dtype=i32)
        self.idp_base, self.idp_m, self.idp_rho = z(1, F, 3), z(1, F, 3), z(1, F)
        self.idp_anchor_slot = torch.full((1, F), -1, dtype=i32)
        self.lost_cnt, self.tracked_cnt = z(1, F, dtype=i32), z(1, F, dtype=i32)
        self.feat_status = z(1, F, dtype=torch.uint8)
        self.gamma, self.dof, self.delta_x = z(1, F), z(1, F, dtype=i32), z(1, self.ldp)
        self.n_accepted, self.n_rejected = z(1, dtype=i32), z(1, dtype=i32)
        self.calls = FakeContext.totals
        for k in ("front", "update", "propagate", "augment", "remove"):
            self.calls.setdefault(k, 0)

    # the "device" arrays ARE host memory here: the mirror is the thing itself
    def small_host_views(self):
        return {k: getattr(self, k).numpy() for k in ("imu_state", "imu_null", "imu_time", "Qc_diag", "cam_R", "cam_t",
                                                      "samples_stage")}

    def push_small(self):
        pass

    def pull_small(self):
        pass

    # ------------------------------------------------------------------------------------------------------------
    def _case(self):
        p = self.params
        n = int(self.n_cams[0])
        d = 15 + 6 * n
        F = int((self.feat_len[0] > 0).sum())
        assert bool((self.feat_len[0, :F] > 0).all()), "the mixin packs its features densely"
        st = self.imu_state[0].numpy()
        K = np.asarray(p.K, float)
        return dict(K=K, Kinv=np.linalg.inv(K), width=p.width, height=p.height, sigma=p.sigma_image,
                    gravity=np.asarray(p.gravity, float), min_lost=max(int(p.min_lost), 1), min_tracked=max(int(p.min_tracked), 2),
                    use_parallax=int(p.use_parallax), min_parallax=float(p.min_parallax), n_cams=n,
                    cam_R=self.cam_R[0, :n].numpy().reshape(n, 3, 3).copy(), cam_t=self.cam_t[0, :n].numpy().copy(),
                    imu_R=st[0:9].reshape(3, 3).copy(), imu_t=st[9:12].copy(), imu_v=st[12:15].copy(), imu_bg=st[15:18].copy(),
                    imu_ba=st[18:21].copy(), P=self.P[0, :d, :d].numpy().copy(),
                    feat_len=self.feat_len[0, :F].numpy().copy(), obs_uv=self.obs_uv[0, :F].numpy().copy(),
                    obs_dir=self.obs_dir[0, :F].numpy().copy(), obs_conf=self.obs_conf[0, :F].numpy().copy(),
                    obs_slot=self.obs_slot[0, :F].numpy().copy(), idp_base=self.idp_base[0, :F].numpy().copy(),
                    idp_anchor_slot=self.idp_anchor_slot[0, :F].numpy().copy(), idp_m=self.idp_m[0, :F].numpy().copy(),
                    idp_rho=self.idp_rho[0, :F].numpy().copy(), lost_cnt=self.lost_cnt[0, :F].numpy().copy(),
                    tracked_cnt=self.tracked_cnt[0, :F].numpy().copy()), F, n, d

    def _selection(self, c, F):
        valid, lost, m, rho, tri = o.select_and_triangulate(c, return_triangulated=True)
        st = np.zeros(self.F, np.uint8)
        for j in range(F):
            if valid[j]:
                code = _lib.FEAT_ACCEPTED
            elif lost[j]:
                code = _lib.FEAT_LOST_DISCARD
            else:
                code = _lib.FEAT_SKIPPED
            st[j] = code | (_lib.FEAT_LOST_FLAG if lost[j] else 0) | (_lib.FEAT_TRIANGULATED if tri[j] else 0)
        self.idp_m[0, :F] = torch.from_numpy(m)
        self.idp_rho[0, :F] = torch.from_numpy(rho)
        return st, valid

    def front(self):
        self.calls["front"] += 1
        c, F, n, d = self._case()
        st, _ = self._selection(c, F)
        self.feat_status[0] = torch.from_numpy(st)

    def update(self):
        self.calls["update"] += 1
        c, F, n, d = self._case()
        st, valid = self._selection(c, F)
        out = o.update_case(c)
        for j in range(F):
            if valid[j] and not out["out_gate"][j]:
                st[j] = (st[j] & 0xF0) | _lib.FEAT_GATED_OUT
        self.feat_status[0] = torch.from_numpy(st)
        self.gamma[0, :F] = torch.from_numpy(out["out_gamma"])
        self.delta_x[0].zero_()
        self.delta_x[0, :d] = torch.from_numpy(out["out_delta_x"])
        self.n_rejected[0] = int(out["out_n_rejected"])
        self.n_accepted[0] = int(out["out_gate"].sum()) if out["out_updated"] else 0
        if out["out_updated"]:
            self.P[0, :d, :d] = torch.from_numpy(out["out_P"])
            s = self.imu_state[0].numpy()
            s[0:9] = out["out_imu_R"].reshape(9)
            s[9:12], s[12:15], s[15:18], s[18:21] = out["out_imu_t"], out["out_imu_v"], out["out_imu_bg"], out["out_imu_ba"]
            self.cam_R[0, :n] = torch.from_numpy(out["out_cam_R"].reshape(n, 9))
            self.cam_t[0, :n] = torch.from_numpy(out["out_cam_t"])

    def propagate(self, samples, use_null_state=False):
        assert use_null_state and samples.shape[:2] == (1, 1)
        self.calls["propagate"] += 1
        n = int(self.n_cams[0])
        d = 15 + 6 * n
        s, z = self.imu_state[0].numpy(), self.imu_null[0].numpy()
        c = dict(sample=samples[0, 0].numpy().copy(), timestamp=float(self.imu_time[0]), gravity=np.asarray(self.params.gravity, float),
                 Qc_diag=np.repeat(self.Qc_diag[0].numpy(), 3), imu_R=s[0:9].reshape(3, 3).copy(), imu_t=s[9:12].copy(),
                 imu_v=s[12:15].copy(), imu_bg=s[15:18].copy(), imu_ba=s[18:21].copy(), null_R=z[0:9].reshape(3, 3).copy(),
                 null_t=z[9:12].copy(), null_v=z[12:15].copy(), P=self.P[0, :d, :d].numpy().copy())
        out = o.propagate_case(c)
        s[0:9], s[9:12], s[12:15] = out["out_imu_R"].reshape(9), out["out_imu_t"], out["out_imu_v"]
        z[0:15] = s[0:15]
        self.imu_time[0] = out["out_timestamp"]
        self.P[0, :d, :d] = torch.from_numpy(out["out_P"])

    def augment(self, R_I_C, t_I_C, active=None):
        self.calls["augment"] += 1
        n = int(self.n_cams[0])
        assert n < self.N
        d = 15 + 6 * n
        s = self.imu_state[0].numpy()
        R, t = s[0:9].reshape(3, 3), s[9:12]
        R_I_C, t_I_C = np.asarray(R_I_C, float), np.asarray(t_I_C, float)
        self.cam_R[0, n] = torch.from_numpy((R @ R_I_C).reshape(9))
        self.cam_t[0, n] = torch.from_numpy(R @ t_I_C + t)
        self.P[0, :d + 6, :d + 6] = torch.from_numpy(o.augment_covariance(self.P[0, :d, :d].numpy().copy(), R_I_C, t_I_C, R.copy()))
        self.n_cams[0] = n + 1

    def remove_cameras(self, mask):
        self.calls["remove"] += 1
        n = int(self.n_cams[0])
        P = self.P[0, :15 + 6 * n, :15 + 6 * n].numpy().copy()
        keep = [s for s in range(n) if not int(mask[0, s])]
        for s in sorted((s for s in range(n) if int(mask[0, s])), reverse=True):
            P = o.remove_camera_covariance(P, s)
        R, t = self.cam_R[0].numpy().copy(), self.cam_t[0].numpy().copy()
        self.cam_R[0].zero_()
        self.cam_t[0].zero_()
        for k, s in enumerate(keep):
            self.cam_R[0, k] = torch.from_numpy(R[s])
            self.cam_t[0, k] = torch.from_numpy(t[s])
        self.P[0].zero_()
        self.P[0, :P.shape[0], :P.shape[0]] = torch.from_numpy(P)
        self.n_cams[0] = len(keep)
