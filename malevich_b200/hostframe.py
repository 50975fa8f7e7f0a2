"""The host frame a job of several rank PROCESSES composes: one row-major colour image per slot in a shared mapping
(/dev/shm), every rank copying the rows it owns into it over its own PCIe link (Device.present_owned_rows_async), plus two
small counters that tell the consumer -- rank 0, the process that would blit the frame like the reference's present()
(main.c:1301-1306) -- when a frame is whole, and the producers when a slot may be overwritten.

    frame f lives in slot f % slots
    done[r]   = number of frames whose rows rank r has delivered (written by rank r after its present_wait)
    consumed  = number of frames rank 0 has taken AND let go of (written by rank 0: taking frame f lets go of every frame before f)

No CUDA here: the mapping is page-locked for a device with Device.register_host_memory by its user."""
import mmap
import os
import time

import numpy as np


class SharedHostFrames:
    def __init__(self, name: str, height: int, width: int, world: int, rank: int, slots: int = 2, create: bool = False):
        self.path = os.path.join("/dev/shm", name)
        self.world, self.rank, self.slots = world, rank, slots
        self.frame_bytes = height * width * 4
        header = 4096  # counters on their own page
        total = header + slots * self.frame_bytes
        if create:
            try:
                os.unlink(self.path)  # left behind by a run that died
            except FileNotFoundError:
                pass
            # tmpfs pages are allocated on first touch and a full /dev/shm answers with SIGBUS, not with an error: refuse up
            # front (a container's default /dev/shm is 64 MB, two 4K frames are 66 MB)
            vfs = os.statvfs("/dev/shm")
            if vfs.f_bavail * vfs.f_frsize < total + (8 << 20):
                raise OSError(f"/dev/shm has {vfs.f_bavail * vfs.f_frsize >> 20} MB free, the shared frames need {total >> 20} MB")
            with open(self.path, "wb") as f:
                f.truncate(total)
        self._fd = os.open(self.path, os.O_RDWR)
        self._map = mmap.mmap(self._fd, total, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        self._counters = np.frombuffer(self._map, dtype=np.int64, count=world + 1, offset=0)
        self.frames = [np.frombuffer(self._map, dtype=np.uint32, count=height * width, offset=header + s * self.frame_bytes).reshape(height, width) for s in range(slots)]
        self.delivered = 0  # frames this rank has issued

    # -- producer side (every rank) ------------------------------------------------------------------
    def slot_for_next(self, timeout_s: float = 10.0) -> np.ndarray:
        """The slot of the frame about to be delivered; waits until rank 0 has consumed the frame that used it before."""
        f = self.delivered
        need = f - self.slots + 1  # frames that must have been consumed before slot f % slots is written again
        if need > 0:
            self._spin(lambda: int(self._counters[self.world]) >= need, timeout_s, "the consumer to free a slot")
        self.delivered += 1
        return self.frames[f % self.slots]

    def publish(self, frames_done: int):
        """This rank's rows of the first `frames_done` frames are in host memory (call after present_wait)."""
        self._counters[self.rank] = frames_done

    # -- consumer side (rank 0) ----------------------------------------------------------------------
    def take(self, f: int, timeout_s: float = 10.0) -> np.ndarray:
        """Frame f, once every rank has delivered its rows of it. The view stays valid until the next take() (or release()):
        taking frame f lets go of the frames before it, whose slots the producers may then overwrite."""
        self._spin(lambda: int(self._counters[:self.world].min()) > f, timeout_s, f"frame {f}")
        self._counters[self.world] = max(int(self._counters[self.world]), f)
        return self.frames[f % self.slots]

    def release(self, f: int):
        """The consumer is done with frame f (and everything before it)."""
        self._counters[self.world] = max(int(self._counters[self.world]), f + 1)

    def reset(self):
        self._counters[self.rank] = 0
        if self.rank == 0:
            self._counters[self.world] = 0
        self.delivered = 0

    @staticmethod
    def _spin(cond, timeout_s, what):
        t0 = time.perf_counter()
        while not cond():
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError(f"shared host frame: waited {timeout_s} s for {what}")

    @property
    def address(self) -> int:
        return np.frombuffer(self._map, dtype=np.uint8).ctypes.data

    @property
    def nbytes(self) -> int:
        return len(self._map)

    def close(self, unlink: bool = False):
        self.frames, self._counters = [], None
        try:
            self._map.close()
        except BufferError:  # a view is still alive somewhere; the mapping goes with the process
            pass
        os.close(self._fd)
        if unlink:
            try:
                os.unlink(self.path)
            except FileNotFoundError:
                pass
