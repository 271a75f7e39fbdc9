"""Per-frame visual pipeline driven by the device-resident track table.

`DeviceMapServer` chains, in DEVICE pointer mode and without any host round trip, the calls that reproduce the visual
part of IngvioFilter::callbackMonoFrame / callbackStereoFrame
(/root/reference/ingvio_estimator/src/IngvioFilter.cpp:143-205, :271-333):

    MapServerManager::collect*Meas            -> collect()
    RemoveLostUpdate::updateState*            -> remove_lost_update()   (mark, select, triangulate, update, erase)
    SwMargUpdate / KeyframeUpdate::updateState* -> selected_update()
    clean*ObsAtMargTime, changeMSCKFAnchor, margSwPose -> slide()
    MapServerManager::eraseInvalidFeatures    -> erase_invalid()

The scratch arrays between the calls (gathered observations, masks, triangulated points) are CUDA tensors owned by this
object; torch is used for device memory only.
"""
import numpy as np

from . import capi


class DeviceMapServer:
    def __init__(self, filt, max_tracks, obs_slots=None, tri_params=None, device="cuda"):
        """device: where the scratch tensors live. "cuda" (default) chains the calls in DEVICE pointer mode; host tensors make
        every call a HOST-pointer call (staged copies), which is how the CPU-side tests drive the same sequence."""
        import torch
        self.f = filt
        self.torch = torch
        filt.create_map_server(max_tracks)
        B, F = filt.B, filt.max_feats
        SW = int(obs_slots or filt.max_clones)
        dev = torch.device(device)
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)   # noqa: E731
        self.buf = dict(track_entry=z((B, F), torch.int32), n_sel=z((B,), torch.int32), track_id=z((B, F), torch.int32),
                        obs=z((B, F, SW, filt.rho), torch.float64), mask_all=z((B, F, SW), torch.uint8),
                        mask_upd=z((B, F, SW), torch.uint8), anchor_slot=z((B, F), torch.int32),
                        chi2_dof=z((B, F), torch.int32), feat_ok=z((B, F), torch.uint8))
        self.pf = z((B, F, 3), torch.float64)
        self.ok = z((B, F), torch.uint8)
        self.SW, self.F = SW, F
        self.tri_params = dict(tri_params or {})
        if dev.type == "cuda":
            torch.cuda.synchronize()   # the zero-fills ran on torch's stream; the handle enqueues on its own

    def collect(self, n_meas, ids, uv):
        """One tracker message per sequence (host arrays or CUDA tensors) at the newest clone."""
        self.f.collect_meas(n_meas, ids, uv)

    def _triangulate_commit(self):
        g = self.buf
        self.f.triangulate(g["obs"], g["mask_all"], g["anchor_slot"], pf_out=self.pf, ok_out=self.ok, **self.tri_params)
        self.f.commit_triangulation(g["track_entry"], self.pf, self.ok, g["feat_ok"])

    def remove_lost_update(self, noise, max_valid=20, **want):
        """RemoveLostUpdate::updateStateMono / Stereo (RemoveLostUpdate.cpp:40-167, :276-405)."""
        f, g = self.f, self.buf
        f.mark_marg_features()
        f.gather_tracks(capi.TRK_LOST, n_feats=self.F, obs_slots=self.SW, out=g)
        self._triangulate_commit()
        out = f.msckf_update(capi.VIS_ALL_OBS, self.pf, g["anchor_slot"], g["obs"], g["mask_upd"], g["chi2_dof"], noise,
                             max_valid=max_valid, feat_ok=g["feat_ok"], **want)
        f.erase_tracks(g["track_entry"])
        return out

    def selected_update(self, selected_slots, noise, dof_fixed=0, **want):
        """SwMargUpdate::updateState* (dof = #selected-1) / KeyframeUpdate::updateState* (dof_fixed = 2)."""
        f, g = self.f, self.buf
        f.gather_tracks(capi.TRK_SEEN_AT, selected_slots=selected_slots, dof_fixed=dof_fixed, n_feats=self.F,
                        obs_slots=self.SW, out=g)
        self._triangulate_commit()
        return f.msckf_update(capi.VIS_SELECTED, self.pf, g["anchor_slot"], g["obs"], g["mask_upd"], g["chi2_dof"], noise,
                              max_valid=0, feat_ok=g["feat_ok"], **want)

    def slide(self, marg_slots, min_depth):
        """clean*ObsAtMargTime -> changeMSCKFAnchor (min_depth: 0 SwMargUpdate, 0.3 KeyframeUpdate) -> margSwPose."""
        f = self.f
        f.clean_obs_at(marg_slots)
        f.change_msckf_anchor(marg_slots, min_depth)
        for s in sorted(marg_slots, reverse=True):
            f.marg_sliding_window_pose(s)

    def erase_invalid(self, min_depth=0.2):
        self.f.erase_invalid_features(min_depth)

    def selected_counts(self):
        """Tracks selected by the last gather, per sequence (synchronises the handle's stream)."""
        self.f.synchronize()
        return self.buf["n_sel"].cpu().numpy().astype(np.int32)
