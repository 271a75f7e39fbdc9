"""IMU mean propagation + analytic transition matrices + the per-frame propagate/augment loop.

TEST INFRASTRUCTURE (oracle). CPU restatement of
/root/reference/ingvio_estimator/src/ImuPropagator.cpp:98-162 (analytic branch),
:163-229 (RK4/Taylor branch, used only by the convergence property test), :232-314.
"""
from dataclasses import dataclass

import numpy as np

from .lie import gamma_func, psi1_func, psi2_func, skew
from .state import BDS, FS, GPS, State
from .state_manager import StateManager


@dataclass
class ImuCtrl:
    timestamp: float
    gyro_raw: np.ndarray
    accel_raw: np.ndarray


def _rot_from_angle_axis(angle, axis):
    return np.eye(3) + np.sin(angle) * skew(axis) + (1 - np.cos(angle)) * skew(axis) @ skew(axis)


class ImuPropagator:
    def __init__(self, gravity_norm=9.8):
        self.gravity = np.array([0.0, 0.0, -gravity_norm])
        self.has_gravity_set = True
        self.imu_ctrl_buffer = []

    def store_imu(self, ctrl: ImuCtrl):
        self.imu_ctrl_buffer.append(ctrl)

    def state_and_cov_transition(self, state: State, ctrl: ImuCtrl, dt, is_analytic=True):
        """ImuPropagator.cpp:98-230. Returns (Phi 15x15, G 15x12) and moves the mean."""
        Phi = np.eye(15)
        G = np.zeros((15, 12))
        R_hat = state.extended_pose.value_linear().copy()
        p_hat = state.extended_pose.value_trans1().copy()
        v_hat = state.extended_pose.value_trans2().copy()
        G[0:3, 0:3] = R_hat
        G[3:6, 0:3] = skew(p_hat) @ R_hat
        G[6:9, 0:3] = skew(v_hat) @ R_hat
        G[6:9, 3:6] = R_hat
        G[9:12, 6:9] = np.eye(3)
        G[12:15, 9:12] = np.eye(3)
        gyro = ctrl.gyro_raw - state.bg.value()
        acc = ctrl.accel_raw - state.ba.value()
        g = self.gravity
        state.timestamp += dt
        if is_analytic:
            G0 = gamma_func(dt * gyro, 0)
            G1 = gamma_func(dt * gyro, 1)
            G2 = gamma_func(dt * gyro, 2)
            state.extended_pose.rot = R_hat @ G0
            v_new = v_hat + g * dt + R_hat @ G1 @ acc * dt
            state.extended_pose.vec2 = v_new
            p_new = p_hat + v_hat * dt + 0.5 * g * dt ** 2 + R_hat @ G2 @ acc * dt ** 2
            state.extended_pose.vec1 = p_new
            self._propagate_clock(state, dt)
            Phi[3:6, 0:3] = 0.5 * skew(g) * dt ** 2
            Phi[3:6, 6:9] = dt * np.eye(3)
            Phi[6:9, 0:3] = skew(g) * dt
            Phi[0:3, 9:12] = -R_hat @ G1 * dt
            Phi[6:9, 12:15] = -R_hat @ G1 * dt
            Phi[3:6, 12:15] = -R_hat @ G2 * dt ** 2
            Phi[6:9, 9:12] = -skew(v_new) @ R_hat @ G1 * dt + R_hat @ psi1_func(gyro, acc, dt)
            Phi[3:6, 9:12] = -skew(p_new) @ R_hat @ G1 * dt + R_hat @ psi2_func(gyro, acc, dt)
        else:
            dang = dt * gyro
            nrm = np.linalg.norm(dang)
            axis = dang / nrm if nrm > 0 else np.array([1.0, 0, 0])
            R_dt2 = R_hat @ _rot_from_angle_axis(nrm / 2, axis)
            R_dt = R_hat @ _rot_from_angle_axis(nrm, axis)
            k1_v = R_hat @ acc + g
            k1_p = v_hat
            k2_v = R_dt2 @ acc + g
            k2_p = v_hat + k1_v * dt / 2.0
            k3_v = R_dt2 @ acc + g
            k3_p = v_hat + k2_v * dt / 2.0
            k4_v = R_dt @ acc + g
            k4_p = v_hat + k3_v * dt
            state.extended_pose.vec2 = v_hat + dt / 6.0 * (k1_v + 2 * k2_v + 2 * k3_v + k4_v)
            state.extended_pose.vec1 = p_hat + dt / 6.0 * (k1_p + 2 * k2_p + 2 * k3_p + k4_p)
            state.extended_pose.rot = R_dt
            self._propagate_clock(state, dt)
            F = np.zeros((15, 15))
            F[3:6, 6:9] = np.eye(3)
            F[6:9, 0:3] = skew(g)
            F[0:3, 9:12] = -R_hat
            F[3:6, 9:12] = -skew(p_hat) @ R_hat
            F[6:9, 9:12] = -skew(v_hat) @ R_hat
            F[6:9, 12:15] = -R_hat
            F2 = F @ F / 2.0
            F3 = F2 @ F / 3.0
            Phi = np.eye(15) + F * dt + F2 * dt * dt + F3 * dt ** 3
        return Phi, G

    @staticmethod
    def _propagate_clock(state: State, dt):
        """ImuPropagator.cpp:139-148."""
        if state.state_params.enable_gnss and FS in state.gnss:
            for i in range(GPS, BDS + 1):
                if i in state.gnss:
                    state.gnss[i].set_value(state.gnss[i].value() + dt * state.gnss[FS].value())

    def propagate_until(self, state: State, t_end, is_analytic=True):
        """ImuPropagator.cpp:232-292."""
        if not self.has_gravity_set or t_end <= state.timestamp:
            return
        if len(self.imu_ctrl_buffer) == 0 or self.imu_ctrl_buffer[0].timestamp > t_end:
            return
        propa_cnt = 0
        last = self.imu_ctrl_buffer[-1]
        for ctrl in self.imu_ctrl_buffer:
            if ctrl.timestamp < state.timestamp:
                propa_cnt += 1
                continue
            if ctrl.timestamp > t_end:
                break
            propa_cnt += 1
            dt = ctrl.timestamp - state.timestamp
            if dt < 1e-6:
                continue
            last = ctrl
            Phi, G = self.state_and_cov_transition(state, ctrl, dt, is_analytic)
            StateManager.propagate_state_cov(state, Phi, G, dt)
        if state.timestamp < t_end:
            dt_last = t_end - state.timestamp
            if dt_last > 1e-6:
                Phi, G = self.state_and_cov_transition(state, last, dt_last, is_analytic)
                StateManager.propagate_state_cov(state, Phi, G, dt_last)
            else:
                state.timestamp = t_end
        del self.imu_ctrl_buffer[:propa_cnt]

    def propagate_augment_at_end(self, state: State, t_end, is_analytic=True):
        """ImuPropagator.cpp:294-314."""
        if not self.has_gravity_set:
            return
        self.propagate_until(state, t_end, is_analytic)
        if state.timestamp != t_end:
            return
        StateManager.augment_sliding_window_pose(state)

    # Convenience used by the synthetic-frame driver: explicit (gyro, accel, dt) steps, i.e. the body
    # of the loop at ImuPropagator.cpp:260-271 with the time bookkeeping already resolved.
    def propagate_steps(self, state: State, gyro, accel, dts):
        for w, a, dt in zip(gyro, accel, dts):
            if dt < 1e-6:
                continue
            Phi, G = self.state_and_cov_transition(state, ImuCtrl(0.0, np.asarray(w, float), np.asarray(a, float)),
                                                   float(dt), True)
            StateManager.propagate_state_cov(state, Phi, G, float(dt))
