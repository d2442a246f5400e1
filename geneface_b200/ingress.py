"""Frame ingress for sequence rendering: everything between the files on disk and the
`(poses, condition windows, background)` arrays sequence.SequenceRenderer takes.  Host code (numpy), no GPU.

Restates, vectorised over the whole sequence instead of per frame:
  data_gen/nerf/binarizer.py:24-59                 get_win_conds             -> window / windows
  inference/nerfs/lm3d_radnerf_infer.py:48-72      normalise, clamp per landmark group, exponential smoothing (lambda 0.2)
                                                                            -> regularize_lm3d
  inference/nerfs/lm3d_radnerf_infer.py:74-86      cond_win / cond_wins      -> cond_windows
  tasks/radnerfs/dataset_utils.py:16-36            smooth_camera_path        -> smooth_camera_path
  tasks/radnerfs/dataset_utils.py:39-105           the camera / background / statistics part of RADNeRFDataset.__init__
                                                                            -> SequenceInputs.load (`trainval_dataset.npy`)
"""
import os

import numpy as np

from .utils import nerf_matrix_to_ngp


# ------------------------------------------------------------------------------------------------------ windows
def _window_index(T, win, pad):
    """[T, win] source frame of every window slot, and a validity mask (False = zero padding)."""
    half = win // 2
    src = np.arange(T)[:, None] - half + np.arange(win)[None, :]
    ok = (src >= 0) & (src < T)
    if pad == 'edge':
        return np.clip(src, 0, T - 1), np.ones_like(ok)
    if pad == 'zero':
        return np.clip(src, 0, T - 1), ok
    raise NotImplementedError(pad)


def windows(conds, win, pad='zero'):
    """All windows at once: conds [T, ...] -> [T, win, ...]; slot k of frame t is frame t - win//2 + k (edge-replicated or zero)."""
    conds = np.asarray(conds)
    src, ok = _window_index(conds.shape[0], win, pad)
    out = conds[src]
    if not ok.all():
        out = out * ok.reshape(ok.shape + (1,) * (conds.ndim - 1)).astype(conds.dtype)
    return out


def window(conds, idx, win=8, pad='zero'):
    """One window (the reference's per-frame call); idx is clamped into the sequence."""
    conds = np.asarray(conds)
    idx = min(max(0, int(idx)), conds.shape[0] - 1)
    return windows(conds, win, pad)[idx]


# ------------------------------------------------------------------------------------------------------ landmarks
# 68-point layout: jaw/yaw 0-16, brows 17-26, nose 27-35, eyes 36-47, mouth 48-67
_HALF_XY = (slice(17, 27), slice(36, 48))          # brows and eyes: x,y clamped at clamp_std / 2, z at clamp_std


def regularize_lm3d(idexp_lm3d, mean, std, clamp_std, lam=0.2):
    """idexp_lm3d [T, 68*3 or 68, 3] raw -> normalised, outlier-clamped, exponentially smoothed [T, 204] float32.
    y_0 = x_0;  y_i = lam * y_{i-1} + (1 - lam) * x_i  (the same lambda for every landmark group in the reference)."""
    x = (np.asarray(idexp_lm3d, np.float32).reshape(-1, 68, 3) - np.asarray(mean, np.float32).reshape(-1, 68, 3)) / np.asarray(std, np.float32).reshape(-1, 68, 3)
    c = np.float32(clamp_std)
    x = np.clip(x, -c, c)
    for g in _HALF_XY:
        x[:, g, 0:2] = np.clip(x[:, g, 0:2], -c / 2, c / 2)
    y = x.copy()
    lam = np.float32(lam)
    for i in range(1, y.shape[0]):
        y[i] = lam * y[i - 1] + (np.float32(1) - lam) * x[i]
    y[0] = lam * x[0] + (np.float32(1) - lam) * x[0]            # the reference also passes frame 0 through the blend
    return y.reshape(-1, 204)


def cond_windows(cond, cond_win_size=1, smo_win_size=5):
    """cond [T, C] -> cond_win [T, cond_win, C] and cond_wins [T, smo_win, cond_win, C] (edge padding), the two tensors the
    inference samples carry (`cond_wins` is what RADNeRF.cal_cond_feat consumes with attention)."""
    cw = windows(np.asarray(cond, np.float32), cond_win_size, 'edge')
    return cw, windows(cw, smo_win_size, 'edge')


# ------------------------------------------------------------------------------------------------------ camera path
def _quat_from_matrix(R):
    """[N,3,3] rotation matrices -> unit quaternions [N,4] (x, y, z, w); branch on the largest diagonal term for stability."""
    R = np.asarray(R, np.float64)
    q = np.empty((R.shape[0], 4))
    t = np.trace(R, axis1=1, axis2=2)
    for n in range(R.shape[0]):
        m = R[n]
        if t[n] > 0:
            s = np.sqrt(t[n] + 1.0) * 2
            q[n] = ((m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s)
        else:
            i = int(np.argmax(np.diag(m)))
            j, k = (i + 1) % 3, (i + 2) % 3
            s = np.sqrt(1.0 + m[i, i] - m[j, j] - m[k, k]) * 2
            v = np.empty(4)
            v[i], v[j], v[k], v[3] = 0.25 * s, (m[j, i] + m[i, j]) / s, (m[k, i] + m[i, k]) / s, (m[k, j] - m[j, k]) / s
            q[n] = v
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def _matrix_from_quat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def mean_rotation(Rs):
    """Chordal L2 mean of rotations: the quaternion that maximises sum (q . q_i)^2 = top eigenvector of sum q_i q_i^T."""
    q = _quat_from_matrix(Rs)
    w, v = np.linalg.eigh(q.T @ q)
    return _matrix_from_quat(v[:, -1])


def smooth_camera_path(poses, kernel_size=7):
    """poses [N,4,4]: box-filter the translation and average the rotation over a centred window (clipped at the ends)."""
    poses = np.array(poses, copy=True)
    N, K = poses.shape[0], kernel_size // 2
    trans, rots = poses[:, :3, 3].copy(), poses[:, :3, :3].copy()
    for i in range(N):
        lo, hi = max(0, i - K), min(N, i + K + 1)
        poses[i, :3, 3] = trans[lo:hi].mean(0)
        poses[i, :3, :3] = mean_rotation(rots[lo:hi])
    return poses


# ------------------------------------------------------------------------------------------------------ dataset file
class SequenceInputs:
    """Camera, background and landmark statistics of a GeneFace `trainval_dataset.npy` (data_gen/nerf/binarizer.py:175-199):
    a pickled dict with H, W, focal, cx, cy, bg_img uint8 [H,W,3], idexp_lm3d_mean/std and `train_samples` / `val_samples`
    lists of per-frame dicts (`c2w` 4x4, `idexp_lm3d_normalized_win`, ...)."""

    def __init__(self, H, W, focal, cx, cy, poses, bg_img, lm3d_mean, lm3d_std, conds=None):
        self.H, self.W, self.focal, self.cx, self.cy = int(H), int(W), float(focal), float(cx), float(cy)
        self.intrinsics = (self.focal, self.focal, self.cx, self.cy)
        self.poses = np.asarray(poses, np.float32)                       # [F,4,4] ngp convention
        self.bg_img = np.asarray(bg_img, np.float32)                     # [H,W,3] in [0,1]
        self.lm3d_mean, self.lm3d_std = np.asarray(lm3d_mean, np.float32), np.asarray(lm3d_std, np.float32)
        self.conds = conds                                               # [F, cond_win, 204] of the stored frames, or None

    @classmethod
    def load(cls, path, prefix='val', camera_scale=4, camera_offset=(0, 0, 0), smooth_kernel=0, cond_win_size=1):
        if os.path.isdir(path):
            path = os.path.join(path, "trainval_dataset.npy")
        ds = np.load(path, allow_pickle=True).tolist()
        keys = {'train': ['train_samples'], 'val': ['val_samples'], 'trainval': ['train_samples', 'val_samples']}
        if prefix not in keys:
            raise ValueError("prefix should be train, val or trainval")
        samples = [s for k in keys[prefix] for s in ds[k]]
        poses = np.stack([nerf_matrix_to_ngp(np.asarray(s['c2w']), scale=camera_scale, offset=list(camera_offset)) for s in samples])
        if np.isnan(poses).any():
            raise ValueError("NaN in the camera poses: check the face tracker output")
        if smooth_kernel:
            poses = smooth_camera_path(poses, smooth_kernel)
        conds = None
        if samples and 'idexp_lm3d_normalized_win' in samples[0]:
            conds = np.stack([np.asarray(s['idexp_lm3d_normalized_win'], np.float32).reshape(cond_win_size, 204) for s in samples])
        return cls(ds['H'], ds['W'], ds['focal'], ds['cx'], ds['cy'], poses, np.asarray(ds['bg_img'], np.float32) / 255.0,
                   ds['idexp_lm3d_mean'], ds['idexp_lm3d_std'], conds)

    def sequence(self, idexp_lm3d, clamp_std, cond_win_size=1, smo_win_size=5):
        """Raw landmark sequence [T,68,3] -> condition windows [T, smo_win, cond_win, 204] float32 for SequenceRenderer; the stored
        poses are cycled when the driving sequence is longer than the stored camera path (the reference indexes idx % len)."""
        cond = regularize_lm3d(idexp_lm3d, self.lm3d_mean, self.lm3d_std, clamp_std)
        _, wins = cond_windows(cond, cond_win_size, smo_win_size)
        T = wins.shape[0]
        poses = self.poses[np.arange(T) % self.poses.shape[0]]
        return poses, wins
