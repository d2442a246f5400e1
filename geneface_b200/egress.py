"""Frame egress for sequence rendering.

The reference converts every finished frame with `.cpu().numpy()` and writes it with `cv2.imwrite` on the rendering thread
(inference/nerfs/base_nerf_infer.py:93-105,165-175), i.e. one device sync and ~10-20 ms of PNG encoding per frame in series with the
renderer.  At >100 frames/s per GPU that writer, not the renderer, would set the frame rate.  Here the renderer hands RGB8 frames
(packed on the GPU by k_finish, drained to pinned host memory by sequence.SequenceRenderer) to a pool of encoder threads:

  encode_png(rgb)             one frame -> PNG bytes (8-bit RGB, non-interlaced; Sub or Up prediction filter + zlib), stdlib only
  PngSequenceWriter(dir)      thread pool; submit(index, frame) returns at once; files are named like the reference's
                              (`00000.png`, `00001.png`, ...) so the reference's ffmpeg step (base_nerf_infer.py:307) runs unchanged

zlib releases the GIL while compressing, so the pool scales with host cores.  Host code only: nothing here touches the GPU.
"""
import os
import struct
import threading
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_PNG_SIG = b"\x89PNG\r\n\x1a\n"


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)


def encode_png(rgb, level=1, prediction="sub"):
    """rgb: uint8 [H, W, 3] (RGB order) -> PNG file content.  prediction: 'none' | 'sub' | 'up' (one filter type for all rows)."""
    a = np.ascontiguousarray(rgb)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("encode_png expects a uint8 [H, W, 3] array, got %s %s" % (a.dtype, a.shape))
    H, W, _ = a.shape
    if prediction == "none":
        ftype, body = 0, a.reshape(H, W * 3)
    elif prediction == "sub":                  # byte minus the same channel of the pixel to the left
        body = a.reshape(H, W * 3).copy()
        body[:, 3:] -= a.reshape(H, W * 3)[:, :-3]
        ftype = 1
    elif prediction == "up":                   # byte minus the byte above
        body = a.reshape(H, W * 3).copy()
        body[1:] -= a.reshape(H, W * 3)[:-1]
        ftype = 2
    else:
        raise ValueError("unknown prediction %r" % (prediction,))
    rows = np.empty((H, W * 3 + 1), np.uint8)
    rows[:, 0] = ftype
    rows[:, 1:] = body
    ihdr = struct.pack(">IIBBBBB", W, H, 8, 2, 0, 0, 0)          # 8-bit, colour type 2 (RGB), deflate, adaptive filtering, no interlace
    return _PNG_SIG + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(rows.tobytes(), level)) + _chunk(b"IEND", b"")


class PngSequenceWriter:
    """Asynchronous `%05d.png` writer.  submit() copies nothing: the caller must not overwrite `frame` until wait()/close() or until
    the returned future is done (SequenceRenderer's host ring holds every frame of the call, so this holds there)."""

    def __init__(self, out_dir, workers=None, level=1, prediction="sub", name_fmt="{:05d}.png", backend="auto"):
        """backend: 'zlib' = encode_png above (stdlib only); 'cv2' = OpenCV's encoder (the reference's own, ~1.5x faster per thread,
        also releases the GIL); 'auto' = cv2 when importable."""
        self.out_dir, self.level, self.prediction, self.name_fmt = out_dir, level, prediction, name_fmt
        self._cv2 = None
        if backend in ("auto", "cv2"):
            try:
                import cv2
                self._cv2 = cv2
            except ImportError:
                if backend == "cv2":
                    raise
        os.makedirs(out_dir, exist_ok=True)
        self.pool = ThreadPoolExecutor(max_workers=workers or min(16, os.cpu_count() or 1))
        self.futures = []
        self.bytes_written = 0
        self._lock = threading.Lock()

    def _write(self, index, frame):
        if self._cv2 is not None:
            ok, buf = self._cv2.imencode(".png", np.asarray(frame)[..., ::-1], [self._cv2.IMWRITE_PNG_COMPRESSION, self.level])
            if not ok:
                raise RuntimeError("cv2.imencode failed for frame %d" % index)
            data = buf.tobytes()
        else:
            data = encode_png(np.asarray(frame), self.level, self.prediction)
        path = os.path.join(self.out_dir, self.name_fmt.format(index))
        with open(path, "wb") as f:
            f.write(data)
        with self._lock:
            self.bytes_written += len(data)
        return path

    def submit(self, index, frame):
        fut = self.pool.submit(self._write, int(index), frame)
        self.futures.append(fut)
        return fut

    def wait(self):
        paths = [f.result() for f in self.futures]        # re-raises encoder/IO errors here
        self.futures = []
        return paths

    def close(self):
        try:
            return self.wait()
        finally:
            self.pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
