"""av.VideoFrame / av.AudioFrame when PyAV is installed (the deployment case: webrtc.py consumes
them); otherwise minimal stand-ins with the same constructor surface so the plumbing can be
exercised without PyAV (tests, bench)."""
import numpy as np

try:
    from av import AudioFrame, VideoFrame     # noqa: F401
    HAVE_AV = True
except ImportError:
    HAVE_AV = False

    class _Plane:
        def __init__(self):
            self.data = b""

        def update(self, b):
            self.data = bytes(b)

    class AudioFrame:
        def __init__(self, format="s16", layout="mono", samples=0):
            self.format, self.layout, self.samples = format, layout, samples
            self.planes = [_Plane()]
            self.sample_rate = None
            self.pts = None

        def to_ndarray(self):
            return np.frombuffer(self.planes[0].data, dtype=np.int16)

    class VideoFrame:
        def __init__(self, arr, format):
            self._arr, self.format = arr, format
            self.height, self.width = arr.shape[:2]
            self.pts = None

        @classmethod
        def from_ndarray(cls, arr, format="bgr24"):
            # av.VideoFrame.from_ndarray COPIES the pixels into the frame's planes; the plugin relies on that (it hands over
            # views into a pinned ring that is overwritten a few batches later), so the stand-in copies too
            return cls(np.array(arr, copy=True), format)

        def to_ndarray(self, format=None):
            return self._arr
