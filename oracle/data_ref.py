"""oracle/data_ref.py — TEST INFRASTRUCTURE (see oracle/__init__.py): CPU restatement of the reference's dataset ->
batch path, numpy only.

  utils/data_utils.py:10-16    get_data_stats            per-dimension min / max over all rows
  utils/data_utils.py:18-21    normalize_data            (x - min) / (max - min) * 2 - 1
  utils/data_utils.py:23-26    unnormalize_data
  utils/data_utils.py:35-40    unnormalize_position      x * 2 + translation, then unnormalize_data
  utils/data_utils.py:46-56    create_sample_indices_sparse
  utils/data_utils.py:58-62    sample_sequence(_array)_sparse   data[start:end:step]
  utils/load_data.py:25-42     CarRacingDataset._create_dataset (velocity / action normalised once, position per window)
  utils/load_data.py:58-76     _compute_stats            position: scalar mean of the per-window minima / maxima
  utils/load_data.py:128-144   _normalize_position + __getitem__ (window, centred on its first point, halved)
  torch DataLoader default_collate: the batch is the stack of the items.

Pinned by tests/golden/dataset.npz, which oracle/make_golden.py::golden_dataset produces by running the reference's own
CarRacingDataset (its zarr reader replaced by the synthetic arrays below) — tests/test_oracle_golden.py checks it on CPU.
"""
import numpy as np


def make_synthetic_dataset(seed=0, n_frames=64, episode_ends=(25, 47, 64), image_hw=96):
    """Deterministic synthetic CarRacing-shaped dataset (legacy RandomState: bit-stable across numpy versions):
    uint8 HWC frames (the reference stores uint8 / 255.0, generateData/trajectory_control_utils.py:170), float32 state."""
    rs = np.random.RandomState(seed)
    img_u8 = rs.randint(0, 256, size=(n_frames, image_hw, image_hw, 3)).astype(np.uint8)
    position = np.cumsum(rs.normal(0.0, 0.7, size=(n_frames, 2)), axis=0).astype(np.float32) + np.float32(40.0)
    velocity = rs.uniform(-30.0, 60.0, size=(n_frames, 2)).astype(np.float32)
    action = rs.uniform(-1.0, 1.0, size=(n_frames, 3)).astype(np.float32)
    return {"img_u8": img_u8, "position": position, "velocity": velocity, "action": action,
            "episode_ends": np.asarray(episode_ends, dtype=np.int64)}


def create_sample_indices_sparse(ends, sequence_length, step_size):
    indices = []
    prev_end = 0
    for end in ends:
        for start in np.arange(prev_end, end - sequence_length + 1, 1):
            if start + sequence_length * step_size <= end:
                indices.append([int(start), int(start + sequence_length * step_size), 0, sequence_length])
        prev_end = end
    return indices


def get_data_stats(data):
    data = data.reshape(-1, data.shape[-1])
    return {"min": np.min(data, axis=0), "max": np.max(data, axis=0)}


def normalize_data(data, stats):
    return (data - stats["min"]) / (stats["max"] - stats["min"]) * 2 - 1


def unnormalize_data(ndata, stats):
    return (ndata + 1) / 2 * (stats["max"] - stats["min"]) + stats["min"]


def unnormalize_position(nsample, translation_vec, position_stats):
    return unnormalize_data(np.array(nsample) * 2.0 + translation_vec, position_stats)


def compute_stats(data, indices, step_size):
    pmin, pmax = [], []
    for start, end, _, _ in indices:
        s = get_data_stats(data["position"][start:end:step_size])
        pmax.append(s["max"])
        pmin.append(s["min"])
    return {"position": {"max": np.average(pmax), "min": np.average(pmin)},
            "velocity": get_data_stats(data["velocity"]), "action": get_data_stats(data["action"])}


class RefWindowDataset:
    """CarRacingDataset without the zarr reader: `data` holds image (N,3,H,W) float, position, velocity, action."""

    def __init__(self, data, episode_ends, pred_horizon, obs_horizon, stats=None, step_size=1):
        self.sequence_len = obs_horizon + pred_horizon
        self.step_size = step_size
        self.indices = create_sample_indices_sparse(episode_ends, self.sequence_len, step_size)
        self.stats = stats if stats is not None else compute_stats(data, self.indices, step_size)
        self.data = {"position": data["position"], "velocity": normalize_data(data["velocity"], self.stats["velocity"]),
                     "action": normalize_data(data["action"], self.stats["action"]), "image": data["image"]}

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, idx):
        start, end, _, _ = self.indices[idx]
        sample = {k: v[start:end:self.step_size] for k, v in self.data.items()}
        pos = normalize_data(sample["position"], self.stats["position"])
        translation = pos[0, :]
        sample["position"] = (pos - translation) / 2.0
        return sample, translation, start, end

    def collate(self, idxs):
        items = [self[i] for i in idxs]
        batch = {k: np.stack([it[0][k] for it in items]) for k in ("image", "position", "velocity", "action")}
        return batch, np.stack([it[1] for it in items]), np.asarray([it[2] for it in items]), np.asarray([it[3] for it in items])


def image_chw_float(img_u8):
    """uint8 HWC frames -> the float CHW array the reference holds in memory (load_data.py:47: np.moveaxis(img, -1, 1))."""
    return np.moveaxis(img_u8.astype(np.float32) / np.float32(255.0), -1, 1)
