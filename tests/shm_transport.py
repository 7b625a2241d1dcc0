"""Test-only transport for zles.dist.ShardedCodec: the destination buffer lives in POSIX shared
memory, so world_size-2 gloo processes on the CPU (emulator library, where a "device pointer" is
a host pointer) exercise the same exchange / offset / peer-write logic as CUDA IPC does on GPUs."""
import ctypes
from multiprocessing import shared_memory


class ShmTransport:
    def __init__(self):
        self.shm = None
        self.base = 0
        self.owner = False

    def _map(self):
        self._keep = (ctypes.c_uint8 * self.shm.size).from_buffer(self.shm.buf)
        self.base = ctypes.addressof(self._keep)

    def create(self, nbytes: int) -> str:
        self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
        self.owner = True
        self._map()
        return self.shm.name

    def open(self, name: str):
        self.shm = shared_memory.SharedMemory(name=name)
        self._map()

    def close(self):
        if self.shm is None:
            return
        self._keep = None
        self.base = 0
        import gc
        gc.collect()
        try:
            self.shm.close()
        except BufferError:
            self.shm._mmap = None  # a ctypes view may still be referenced by a frame; the mapping dies with the process
        if self.owner:
            self.shm.unlink()
        self.shm = None
