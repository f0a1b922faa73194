"""`mpi4py.MPI` subset: COMM_WORLD / COMM_SELF with Get_rank, Get_size, py2f, Barrier, allgather (python ints),
Split (only splits that keep the world or isolate the caller are representable by the shim's two communicators)."""
import ctypes
import os

_here = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.path.join(_here, "..", "..", "lib", "libmpi.so"), mode=ctypes.RTLD_GLOBAL)
_MPI_INT, _MPI_LONG = 4, 5
_COMM_WORLD, _COMM_SELF = 1, 2

_flag = ctypes.c_int(0)
_lib.MPI_Initialized(ctypes.byref(_flag))
if not _flag.value:
    _prov = ctypes.c_int(0)
    _lib.MPI_Init_thread(None, None, 3, ctypes.byref(_prov))
    import atexit

    atexit.register(_lib.MPI_Finalize)


class Comm:
    def __init__(self, handle: int):
        self._h = handle

    def Get_rank(self) -> int:
        v = ctypes.c_int(0)
        _lib.MPI_Comm_rank(self._h, ctypes.byref(v))
        return v.value

    def Get_size(self) -> int:
        v = ctypes.c_int(0)
        _lib.MPI_Comm_size(self._h, ctypes.byref(v))
        return v.value

    rank = property(Get_rank)
    size = property(Get_size)

    def py2f(self) -> int:
        return int(_lib.MPI_Comm_c2f(self._h))

    def Barrier(self) -> None:
        _lib.MPI_Barrier(self._h)

    def allgather(self, value: int):
        n = self.Get_size()
        send = ctypes.c_long(int(value))
        recv = (ctypes.c_long * n)()
        rc = _lib.MPI_Allgather(ctypes.byref(send), 1, _MPI_LONG, recv, 1, _MPI_LONG, self._h)
        if rc != 0:
            raise RuntimeError(f"MPI_Allgather failed with error code {rc}")
        return [int(v) for v in recv]


COMM_WORLD = Comm(_COMM_WORLD)
COMM_SELF = Comm(_COMM_SELF)
SUM, MAX, MIN = 3, 1, 2
