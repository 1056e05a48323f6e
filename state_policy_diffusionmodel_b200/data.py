"""Device-resident dataset -> batch path.

Mirrors the reference's `utils/load_data.py::CarRacingDataset` / `CarRacingDatasetForInference` and the helpers of
`utils/data_utils.py` it is built from, with the arrays resident in HBM and a batch built by two CUDA launches
(`spdm_gather_windows`, csrc/data_kernels.cu) instead of B `__getitem__` calls + a collate on the host:

    ds = DeviceWindowDataset(images_u8_hwc, position, velocity, action, episode_ends, pred_horizon, obs_horizon,
                             stats=None, step_size=5)             # same arguments / meaning as CarRacingDataset
    batch, translation, start, end = ds.batch(idxs)               # == default_collate([ds_ref[i] for i in idxs])
    traj = ds.unnormalize_position(sampled_positions, translation)  # utils/data_utils.py:35-40, on the device

Window indices and the normalisation statistics are one-off host preprocessing in the reference as well
(`create_sample_indices_sparse`, `_compute_stats`); they are restated here in numpy with the reference's arithmetic
(float32).  Everything per batch runs on the GPU; there is no CPU fallback for it.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def create_sample_indices_sparse(ends, sequence_length, step_size):
    """utils/data_utils.py:46-56 — [start, end, 0, sequence_length] for every window that fits inside its episode."""
    indices = []
    prev_end = 0
    for end in ends:
        end = int(end)
        for start in range(prev_end, end - sequence_length + 1):
            if start + sequence_length * step_size <= end:
                indices.append([start, start + sequence_length * step_size, 0, sequence_length])
        prev_end = end
    return indices


def get_data_stats(data):
    """utils/data_utils.py:10-16"""
    data = np.asarray(data).reshape(-1, np.asarray(data).shape[-1])
    return {"min": np.min(data, axis=0), "max": np.max(data, axis=0)}


def compute_stats(position, velocity, action, indices, step_size):
    """utils/load_data.py:58-76 — position: scalar mean of the per-window minima / maxima; velocity, action: per-dimension."""
    pmin, pmax = [], []
    for start, end, _, _ in indices:
        s = get_data_stats(position[start:end:step_size])
        pmax.append(s["max"])
        pmin.append(s["min"])
    return {"position": {"max": np.average(pmax), "min": np.average(pmin)},
            "velocity": get_data_stats(velocity), "action": get_data_stats(action)}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceWindowDataset:
    """CarRacingDataset with its arrays in HBM.  `images`: uint8 (N, H, W, 3) [decoded / 255 on the fly] or float (N, 3, H, W)."""

    def __init__(self, images, position, velocity, action, episode_ends, pred_horizon, obs_horizon, stats=None, step_size=1,
                 device="cuda", image_frames=None):
        self.obs_horizon, self.pred_horizon = int(obs_horizon), int(pred_horizon)
        self.sequence_len = self.obs_horizon + self.pred_horizon
        self.step_size = int(step_size)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceWindowDataset needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        position = np.asarray(position, dtype=np.float32)
        velocity = np.asarray(velocity, dtype=np.float32)
        action = np.asarray(action, dtype=np.float32)
        self.indices = create_sample_indices_sparse(episode_ends, self.sequence_len, self.step_size)
        self.stats = stats if stats is not None else compute_stats(position, velocity, action, self.indices, self.step_size)
        st = _lib.SpdmDataStats()
        st.pos_min, st.pos_max = float(np.float32(self.stats["position"]["min"])), float(np.float32(self.stats["position"]["max"]))
        for d in range(2):
            st.vel_min[d], st.vel_max[d] = float(self.stats["velocity"]["min"][d]), float(self.stats["velocity"]["max"][d])
        for d in range(3):
            st.act_min[d], st.act_max[d] = float(self.stats["action"]["min"][d]), float(self.stats["action"]["max"][d])
        self._st = st
        images = torch.as_tensor(images)
        if images.dtype == torch.uint8:
            if images.dim() != 4 or images.shape[-1] != 3:
                raise ValueError("uint8 images must be (N, H, W, 3)")
            self.image_kind, self.H, self.W = _lib.IMAGES_U8_HWC, int(images.shape[1]), int(images.shape[2])
        else:
            if images.dim() != 4 or images.shape[1] != 3:
                raise ValueError("float images must be (N, 3, H, W)")
            images = images.float()
            self.image_kind, self.H, self.W = _lib.IMAGES_F32_CHW, int(images.shape[2]), int(images.shape[3])
        if self.W % 4:
            raise ValueError("image width must be a multiple of 4")
        self.n_frames = int(images.shape[0])
        if not (len(position) == len(velocity) == len(action) == self.n_frames):
            raise ValueError("images / position / velocity / action disagree on the number of frames")
        self.images = images.contiguous().to(self.device)
        self.position = torch.from_numpy(position).contiguous().to(self.device)
        self.velocity = torch.from_numpy(velocity).contiguous().to(self.device)
        self.action = torch.from_numpy(action).contiguous().to(self.device)
        # frames of a window that get images: the reference item carries all of them, the model reads obs_horizon
        self.image_frames = self.sequence_len if image_frames is None else int(image_frames)
        self._starts_all = torch.tensor([i[0] for i in self.indices], dtype=torch.int64, device=self.device)

    def __len__(self):
        return len(self.indices)

    def batch(self, idxs, image_frames=None):
        """The collated batch of windows `idxs` -> (dict(image, position, velocity, action), translation, start_idx, end_idx)."""
        idxs = torch.as_tensor(idxs, dtype=torch.int64, device=self.device)
        if idxs.numel() == 0:
            raise ValueError("empty batch")
        if int(idxs.min()) < 0 or int(idxs.max()) >= len(self.indices):
            raise IndexError("window index out of range")
        B, T = int(idxs.numel()), self.sequence_len
        T_img = self.image_frames if image_frames is None else int(image_frames)
        starts = self._starts_all[idxs].contiguous()
        out = {"position": torch.empty((B, T, 2), device=self.device), "velocity": torch.empty((B, T, 2), device=self.device),
               "action": torch.empty((B, T, 3), device=self.device)}
        tr = torch.empty((B, 2), device=self.device)
        img = torch.empty((B, T_img, 3, self.H, self.W), device=self.device) if T_img > 0 else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_gather_windows(
                _ptr(self.images), self.image_kind, self.H, self.W, _ptr(self.position), _ptr(self.velocity), _ptr(self.action),
                _ptr(starts), B, T, T_img, self.step_size, ctypes.byref(self._st), _ptr(img) if img is not None else None,
                _ptr(out["position"]), _ptr(out["velocity"]), _ptr(out["action"]), _ptr(tr), _stream()))
        if img is not None:
            out["image"] = img
        return out, tr, starts, starts + T * self.step_size

    def unnormalize_position(self, npos, translation):
        """utils/data_utils.py:35-40 on device tensors: npos (..., rows, 2) with translation (n, 2), n = prod(leading dims)."""
        npos = npos.to(self.device, torch.float32).contiguous()
        translation = translation.to(self.device, torch.float32).contiguous()
        rows = int(npos.shape[-2])
        n = npos.numel() // (rows * 2)
        if npos.shape[-1] != 2 or translation.numel() != n * 2:
            raise ValueError("npos must be (..., rows, 2) with one translation vector per sample")
        out = torch.empty_like(npos)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.spdm_unnormalize_position(_ptr(npos), _ptr(translation), self._st.pos_min, self._st.pos_max, n, rows,
                                                          _ptr(out), _stream()))
        return out
