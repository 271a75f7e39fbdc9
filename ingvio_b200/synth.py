"""Seeded synthetic IMU / feature / GNSS streams for the BASELINE.json configs.

Trajectory: helix, radius 20 m, speed 5 m/s, yaw-rate 0.25 rad/s, vertical sinusoid +-2 m; IMU 200 Hz,
camera 20 Hz (K_imu = 10); IMU noise from /root/reference/config/sportsfield/ingvio_mono.yaml:16-19.
Every feature of a frame is observed in all SW clones of the window (worst-case classic MSCKF).
Seeds: 20240925 + 1000*config_id + sequence_id. Streams are open-loop (they do not depend on the
filter's estimate), so the same packets drive the CUDA path, the oracle and the CPU port.
"""
from dataclasses import dataclass

import numpy as np

from .frames import FramePacket, GnssArrays

G_NORM = 9.8
IMU_RATE = 200.0
CAM_RATE = 20.0
K_IMU = int(IMU_RATE / CAM_RATE)

# camera axes in the IMU frame: cam z -> body x (forward), cam x -> body -y, cam y -> body -z
R_C2I = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
P_C2I = np.array([0.05, 0.02, -0.01])
# stereo extrinsics as in /root/reference/ingvio_estimator/test/TestTriangulator.cpp:37-38
R_CL2CR = np.eye(3)
P_CL2CR = np.array([0.001, -0.12, 0.003])


@dataclass(frozen=True)
class Workload:
    name: str
    config_id: int
    sw: int            # clones in the window during the visual update
    feats: int
    sats: int
    stereo: bool = False

    @property
    def rho(self):
        return 4 if self.stereo else 2

    @property
    def n_gnss(self):
        return 6 if self.sats > 0 else 0

    @property
    def dim(self):
        """N = 21 + g + 6*SW."""
        return 21 + self.n_gnss + 6 * self.sw


WORKLOADS = {
    "c1": Workload("c1", 1, 5, 30, 0),
    "c2": Workload("c2", 2, 11, 150, 12),
    "c3": Workload("c3", 3, 11, 110, 12, stereo=True),
    "c4": Workload("c4", 4, 11, 150, 12),      # batch=64 of the c2 shape, distinct seeds
    "c5": Workload("c5", 5, 30, 400, 20),
    "tiny": Workload("tiny", 9, 4, 12, 5),     # fast parity case
    "tiny_stereo": Workload("tiny_stereo", 10, 4, 10, 5, stereo=True),
}


def _rotz(a):
    c, s = np.cos(a), np.sin(a)
    z, o = np.zeros_like(a), np.ones_like(a)
    return np.stack([np.stack([c, -s, z], -1), np.stack([s, c, z], -1), np.stack([z, z, o], -1)], -2)


def _rotx(a):
    c, s = np.cos(a), np.sin(a)
    z, o = np.zeros_like(a), np.ones_like(a)
    return np.stack([np.stack([o, z, z], -1), np.stack([z, c, -s], -1), np.stack([z, s, c], -1)], -2)


class Trajectory:
    """Analytic ground truth, vectorised over a batch of phase offsets."""

    def __init__(self, phase):
        self.phase = np.asarray(phase, dtype=np.float64)  # (B,)
        self.r, self.w = 20.0, 0.25
        self.az, self.wz = 2.0, 0.5
        self.aroll, self.wroll = 0.05, 0.7

    def pos(self, t):
        th = self.w * t + self.phase
        return np.stack([self.r * np.cos(th), self.r * np.sin(th), self.az * np.sin(self.wz * t + self.phase)], -1)

    def vel(self, t):
        th = self.w * t + self.phase
        return np.stack([-self.r * self.w * np.sin(th), self.r * self.w * np.cos(th),
                         self.az * self.wz * np.cos(self.wz * t + self.phase)], -1)

    def acc(self, t):
        th = self.w * t + self.phase
        return np.stack([-self.r * self.w ** 2 * np.cos(th), -self.r * self.w ** 2 * np.sin(th),
                         -self.az * self.wz ** 2 * np.sin(self.wz * t + self.phase)], -1)

    def rot(self, t):
        yaw = self.w * t + self.phase + np.pi / 2
        roll = self.aroll * np.sin(self.wroll * t + self.phase)
        return _rotz(yaw) @ _rotx(roll)

    def omega_body(self, t):
        # R = Rz(yaw) Rx(roll): body rate = Rx^T [0,0,yaw'] + [roll',0,0]
        roll = self.aroll * np.sin(self.wroll * t + self.phase)
        droll = self.aroll * self.wroll * np.cos(self.wroll * t + self.phase)
        zdot = np.stack([np.zeros_like(roll), np.zeros_like(roll), np.full_like(roll, self.w)], -1)
        wb = np.einsum("bji,bj->bi", _rotx(roll), zdot)
        wb[..., 0] += droll
        return wb


class SyntheticStream:
    """Generates FramePackets for B sequences of one workload."""

    sigma_g, sigma_a = 0.004, 0.08
    obs_noise = 0.01
    pf_noise = 0.05

    def __init__(self, wl: Workload, batch: int, seq0: int = 0):
        self.wl = wl
        self.B = batch
        self.seeds = [20240925 + 1000 * wl.config_id + seq0 + b for b in range(batch)]
        self.rngs = [np.random.Generator(np.random.PCG64(s)) for s in self.seeds]
        self.traj = Trajectory(np.array([r.uniform(0, 2 * np.pi) for r in self.rngs]))
        self.bg = np.stack([r.normal(0, 2e-3, 3) for r in self.rngs])
        self.ba = np.stack([r.normal(0, 2e-2, 3) for r in self.rngs])
        self.frame_idx = 0   # index of the next camera frame; frame i is at t = i / CAM_RATE
        self.n_clones = 0    # clones in the filter window before the next frame (tracked open-loop)
        lat, lon = np.deg2rad(40.0), np.deg2rad(116.3)
        sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
        self.R_enu2ecef = np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0.0, cl, sl]])

    # -- initial condition --------------------------------------------------------------------
    def initial_state(self):
        t0 = 0.0
        z = np.zeros(self.B)
        return dict(t=t0, R=self.traj.rot(z + t0), p=self.traj.pos(z + t0), v=self.traj.vel(z + t0),
                    bg=np.zeros((self.B, 3)), ba=np.zeros((self.B, 3)))

    def cam_pose(self, t):
        tt = np.zeros(self.B) + t
        R = self.traj.rot(tt)
        p = self.traj.pos(tt)
        return R @ R_C2I, np.einsum("bij,j->bi", R, P_C2I) + p

    def _normal(self, shape_tail, scale):
        return np.stack([r.normal(0.0, scale, shape_tail) for r in self.rngs])

    def _uniform(self, lo, hi, shape_tail):
        return np.stack([r.uniform(lo, hi, shape_tail) for r in self.rngs])

    # -- one frame --------------------------------------------------------------------------------
    def next_frame(self, with_visual=True, with_gnss=True, marg_oldest=True):
        """Frame packet for camera frame `frame_idx+1` (IMU over (t_i, t_{i+1}]). While the window
        fills, fewer than wl.sw clones take part; from then on the oldest clone is marginalised."""
        wl, B = self.wl, self.B
        i0 = self.frame_idx
        t0, t1 = i0 / CAM_RATE, (i0 + 1) / CAM_RATE
        ts = t0 + (np.arange(K_IMU) + 1) / IMU_RATE          # sample stamps; dt = stamp - state time
        gyro = np.empty((B, K_IMU, 3))
        accel = np.empty((B, K_IMU, 3))
        for k in range(K_IMU):
            # zero-order hold of the sample at the *start* of each step (mid-point would also do)
            tk = np.zeros(B) + ts[k] - 0.5 / IMU_RATE
            Rk = self.traj.rot(tk)
            a_w = self.traj.acc(tk) + np.array([0.0, 0.0, G_NORM])
            accel[:, k] = np.einsum("bji,bj->bi", Rk, a_w)
            gyro[:, k] = self.traj.omega_body(tk)
        gyro += self.bg[:, None, :] + self._normal((K_IMU, 3), self.sigma_g * np.sqrt(IMU_RATE) * 0.05)
        accel += self.ba[:, None, :] + self._normal((K_IMU, 3), self.sigma_a * np.sqrt(IMU_RATE) * 0.05)
        dt = np.full((B, K_IMU), 1.0 / IMU_RATE)
        self.frame_idx += 1
        ncl = self.n_clones + 1
        assert ncl <= wl.sw
        F, rho = wl.feats, wl.rho
        pf_w = np.zeros((B, F, 3))
        obs = np.zeros((B, F, wl.sw, rho))
        mask = np.zeros((B, F, wl.sw), dtype=np.uint8)
        anchor = np.zeros((B, F), dtype=np.int32)
        mode = None
        if with_visual and ncl >= 3 and F > 0:
            mode = "all_obs"
            frames = [self.frame_idx - (ncl - 1) + s for s in range(ncl)]   # slot s <-> camera frame
            Rm, pm = self.cam_pose(frames[ncl // 2] / CAM_RATE)
            depth = self._uniform(3.0, 40.0, (F,))
            xy = self._uniform(-0.45, 0.45, (F, 2))
            pc = np.concatenate([xy * depth[..., None], depth[..., None]], -1)    # (B,F,3) in mid camera
            truth = np.einsum("bij,bfj->bfi", Rm, pc) + pm[:, None, :]
            pf_w = truth + self._normal((F, 3), self.pf_noise)
            for s, fr in enumerate(frames):
                Rc, pcw = self.cam_pose(fr / CAM_RATE)
                q = np.einsum("bji,bfj->bfi", Rc, truth - pcw[:, None, :])
                obs[:, :, s, 0:2] = q[..., 0:2] / q[..., 2:3]
                if wl.stereo:
                    qr = np.einsum("ij,bfj->bfi", R_CL2CR, q) + P_CL2CR
                    obs[:, :, s, 2:4] = qr[..., 0:2] / qr[..., 2:3]
                mask[:, :, s] = 1
            obs[:, :, :ncl, :] += self._normal((F, ncl, rho), self.obs_noise)
            # anchors: newest clone for most, a random slot for every 4th track
            anchor[:] = ncl - 1
            rnd = np.stack([r.integers(0, ncl, F) for r in self.rngs]).astype(np.int32)
            anchor[:, ::4] = rnd[:, ::4]
        gn = None
        if with_gnss and wl.sats > 0:
            S = wl.sats
            az = self._uniform(0.0, 2 * np.pi, (S,))
            el = self._uniform(np.deg2rad(20.0), np.deg2rad(85.0), (S,))
            enu = np.stack([np.cos(el) * np.sin(az), np.cos(el) * np.cos(az), np.sin(el)], -1)
            unit = np.einsum("ij,bsj->bsi", self.R_enu2ecef, enu)
            ura = np.full((B, S), 2.0)
            psr_std = np.full((B, S), 1.0)
            dstd = np.full((B, S), 0.2)
            sys = np.tile((np.arange(S) % 4).astype(np.int32), (B, 1))
            gn = GnssArrays(unit=unit, res_pos=np.zeros((B, S)), res_vel=np.zeros((B, S)), sys=sys, ura=ura,
                            psr_std=psr_std, dopp_std_mps=dstd, el=el,
                            R_enu2ecef=np.tile(self.R_enu2ecef, (B, 1, 1)))
            gn.res_pos = self._normal((S,), 1.0) * gn.sigma_psr()
            gn.res_vel = self._normal((S,), 1.0) * gn.sigma_dopp()
        marg = [0] if (marg_oldest and ncl >= wl.sw) else []
        self.n_clones = ncl - len(marg)
        return FramePacket(t=t1, gyro=gyro, accel=accel, dt=dt, pf_w=pf_w, anchor_slot=anchor, obs=obs,
                           obs_mask=mask, obs_total=mask.sum(-1).astype(np.int32), visual_mode=mode,
                           selected_slots=[], marg_slots=marg, max_valid=F, gnss=gn)


# ------------------------------------------------------------------------------------------------
# Raw GNSS epochs (the inputs of igv_gnss_residuals = the outputs of gnss_comm::sat_states + raw L1 observations)
# ------------------------------------------------------------------------------------------------
KLOBUCHAR = np.array([0.1118e-7, -0.7451e-8, -0.5961e-7, 0.1192e-6, 0.1167e6, -0.2294e6, -0.1311e6, 0.1049e7])
L1_FREQ = {0: 1575.42e6, 1: 1602.0e6, 2: 1575.42e6, 3: 1561.098e6}   # GPS, GLO (channel 0), GAL, BDS


def geo2ecef(lat_deg, lon_deg, h):
    a, e2 = 6378137.0, 6.69437999014e-3
    la, lo = np.deg2rad(lat_deg), np.deg2rad(lon_deg)
    N = a / np.sqrt(1 - e2 * np.sin(la) ** 2)
    return np.array([(N + h) * np.cos(la) * np.cos(lo), (N + h) * np.cos(la) * np.sin(lo), (N * (1 - e2) + h) * np.sin(la)])


def enu2ecef_rotation(lat_deg, lon_deg):
    la, lo = np.deg2rad(lat_deg), np.deg2rad(lon_deg)
    sl, cl, so, co = np.sin(la), np.cos(la), np.sin(lo), np.cos(lo)
    return np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0.0, cl, sl]])


def raw_gnss_epoch(rng, rcv_ecef, rcv_vel_ecef, clock_bias4, clock_drift, S, lat_deg, lon_deg, el_range=(10.0, 85.0),
                   no_l1=(), below_horizon=()):
    """One synthetic raw epoch for B receivers (rcv_ecef (B,3), rcv_vel_ecef (B,3), clock_bias4 (B,4), clock_drift (B,)):
    satellites on a 26 560 km shell at random azimuth / elevation, orbital speed 3.9 km/s, small clock terms, and
    pseudo-ranges / Dopplers consistent with the receiver up to ~10 m of unmodelled delay + noise. `no_l1` lists
    satellite indices without an L1 observation (freq = -1), `below_horizon` indices placed under the horizon."""
    c = 2.99792458e8
    B = rcv_ecef.shape[0]
    Re = enu2ecef_rotation(lat_deg, lon_deg)
    az = rng.uniform(0, 2 * np.pi, (B, S))
    el = np.deg2rad(rng.uniform(el_range[0], el_range[1], (B, S)))
    for i in below_horizon:
        el[:, i] = np.deg2rad(-5.0)
    enu = np.stack([np.cos(el) * np.sin(az), np.cos(el) * np.cos(az), np.sin(el)], -1)
    u = np.einsum("ij,bsj->bsi", Re, enu)
    r0 = np.linalg.norm(rcv_ecef, axis=-1)[:, None]
    ru = np.einsum("bi,bsi->bs", rcv_ecef, u)
    rho = -ru + np.sqrt(ru * ru + 26.56e6 ** 2 - r0 * r0)
    pos = rcv_ecef[:, None, :] + rho[..., None] * u
    t = rng.standard_normal((B, S, 3))
    t -= np.einsum("bsi,bsi->bs", t, pos)[..., None] * pos / np.einsum("bsi,bsi->bs", pos, pos)[..., None]
    vel = 3.9e3 * t / np.linalg.norm(t, axis=-1, keepdims=True)
    sys = np.tile((np.arange(S) % 4).astype(np.int32), (B, 1))
    freq = np.vectorize(L1_FREQ.get)(sys).astype(np.float64)
    sdt = rng.normal(0, 1e-4, (B, S))
    sddt = rng.normal(0, 1e-11, (B, S))
    tgd = rng.normal(0, 1e-8, (B, S))
    cb = np.take_along_axis(clock_bias4, sys.astype(np.int64), axis=1)
    psr = rho + cb - sdt * c + tgd * c + rng.uniform(4.0, 12.0, (B, S)) + rng.normal(0, 1.0, (B, S))
    rate = np.einsum("bsi,bsi->bs", vel - rcv_vel_ecef[:, None, :], u) + clock_drift[:, None] - sddt * c
    dopp = -(rate + rng.normal(0, 0.1, (B, S))) * freq / c
    for i in no_l1:
        freq[:, i] = -1.0
    obs = np.stack([psr, dopp, freq], -1)
    obs_std = np.stack([np.full((B, S), 2.0), np.full((B, S), 1.0), np.full((B, S), 1.0)], -1)
    ttx = np.stack([rng.uniform(1.0, 366.0, (B, S)), rng.uniform(0.0, 604800.0, (B, S))], -1)
    return dict(sat_pos=pos, sat_vel=vel, sat_clk=np.stack([sdt, sddt, tgd], -1), obs=obs, obs_std=obs_std, ttx=ttx,
                sys=sys, iono=np.tile(KLOBUCHAR, (B, 1)))


def random_ephemerides(rng, B, S, geo=()):
    """Plausible broadcast ephemerides in the IGV_EPH_* record layout (B,S,24) for constellations sys = i % 4
    (GPS, GLO, GAL, BDS); indices in `geo` become BDS GEO satellites (prn <= 5). GLONASS records carry a state on a
    25 500 km orbit. Returns (eph, sys, t_obs_rel, psr)."""
    eph = np.zeros((B, S, 24))
    sys = np.tile((np.arange(S) % 4).astype(np.int32), (B, 1))
    for i in geo:
        sys[:, i] = 3
    for b in range(B):
        for i in range(S):
            k = sys[b, i]
            if k == 1:
                r = 25.5e6
                u = rng.standard_normal(3); u /= np.linalg.norm(u)
                t = np.cross(u, rng.standard_normal(3)); t /= np.linalg.norm(t)
                eph[b, i, 0:3] = r * u
                eph[b, i, 3:6] = 3.95e3 * t - 7.292115e-5 * np.cross([0, 0, 1.0], r * u)   # ECEF (rotating frame) velocity
                eph[b, i, 6:9] = rng.normal(0, 1e-6, 3)
                eph[b, i, 9] = rng.normal(0, 1e-4)      # tau_n
                eph[b, i, 10] = rng.normal(0, 1e-11)    # gamma
            else:
                is_geo = i in geo
                A = 42.164e6 if is_geo else rng.uniform(26.4e6, 29.7e6)
                eph[b, i, :22] = [A, rng.uniform(1e-4, 0.02), 0.09 if is_geo else rng.uniform(0.93, 0.99),
                                  rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-3, 3), rng.normal(0, 4e-9),
                                  rng.normal(-8e-9, 1e-9), rng.normal(0, 3e-10), rng.normal(0, 2e-6), rng.normal(0, 5e-6),
                                  rng.normal(200, 80), rng.normal(0, 60), rng.normal(0, 1e-7), rng.normal(0, 1e-7),
                                  rng.normal(0, 2e-4), rng.normal(0, 1e-11), 0.0, rng.uniform(0, 604800.0),
                                  rng.normal(0, 8e-9), rng.choice([0.0, 16.0, -7200.0]), 3 if is_geo else rng.integers(6, 30)]
    t_obs = rng.uniform(-3600.0, 3600.0, (B, S))
    psr = rng.uniform(2.0e7, 2.6e7, (B, S))
    return eph, sys, t_obs, psr


# ------------------------------------------------------------------------------------------------
# Tracker messages (the inputs of igv_tracks_collect = feature_tracker/msg/{Mono,Stereo}Frame.msg)
# ------------------------------------------------------------------------------------------------
class TrackerStream:
    """Persistent landmarks seen from a SyntheticStream's true camera: per frame one message per sequence with
    (id: uint64, u0 v0 [u1 v1]: float64) in shuffled order, tracks being born in front of the camera and dying at random
    or when they leave the field of view. Open loop and seeded like the stream."""

    def __init__(self, stream: "SyntheticStream", n_tracks: int, meas_stride: int, death_prob=0.12, id_base=0):
        self.s = stream
        self.n, self.M = n_tracks, meas_stride
        self.death = death_prob
        self.id_base = id_base
        self.rngs = [np.random.Generator(np.random.PCG64(seed + 77_000)) for seed in stream.seeds]
        self.lm = [dict() for _ in range(stream.B)]
        self.next_id = [1] * stream.B

    def message(self, t):
        B, rho = self.s.B, self.s.wl.rho
        R, p = self.s.cam_pose(t)
        n_meas = np.zeros(B, np.int32)
        ids = np.zeros((B, self.M), np.uint64)
        uv = np.zeros((B, self.M, rho))
        for b in range(B):
            rng, lm = self.rngs[b], self.lm[b]
            vis = {}
            for k, x in list(lm.items()):
                q = R[b].T @ (x - p[b])
                if q[2] < 1.0 or abs(q[0] / q[2]) > 0.8 or abs(q[1] / q[2]) > 0.8 or rng.random() < self.death:
                    del lm[k]
                else:
                    vis[k] = q
            while len(lm) < self.n:
                d = rng.uniform(3.0, 40.0)
                q = np.array([rng.uniform(-0.45, 0.45) * d, rng.uniform(-0.45, 0.45) * d, d])
                k = self.next_id[b]
                self.next_id[b] += 1
                lm[k] = R[b] @ q + p[b]
                vis[k] = q
            keys = list(vis.keys())
            keys = [keys[i] for i in rng.permutation(len(keys))][:self.M]
            n_meas[b] = len(keys)
            for i, k in enumerate(keys):
                q = vis[k]
                z = [q[0] / q[2], q[1] / q[2]]
                if rho == 4:
                    qr = R_CL2CR @ q + P_CL2CR
                    z += [qr[0] / qr[2], qr[1] / qr[2]]
                ids[b, i] = np.uint64(self.id_base + k)
                uv[b, i] = np.array(z) + rng.normal(0.0, SyntheticStream.obs_noise, rho)
        return n_meas, ids, uv
