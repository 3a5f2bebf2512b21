"""MPI task farm for per-particle likelihood calls: API of ``pocomc.parallel.MPIPool``
(pocomc/parallel.py:20-178).  This is the reference's CPU likelihood farm, not the GPU
communication layer (that is torch.distributed/NCCL in pocomc_b200.dist); it is kept so scripts
that pass ``pool=MPIPool()`` keep working when mpi4py is installed."""
import atexit
import sys

__all__ = ["MPIPool"]


def _import_mpi(use_dill=False):
    try:
        from mpi4py import MPI as _MPI
    except ImportError as e:
        raise ImportError("Please install mpi4py") from e
    if use_dill:
        import dill
        _MPI.pickle.__init__(dill.dumps, dill.loads, dill.HIGHEST_PROTOCOL)
    return _MPI


class MPIPool:
    """Master (rank 0) hands tasks to workers (other ranks) one at a time and collects results in
    task order.  Workers block in ``wait()`` until the master closes the pool."""

    def __init__(self, comm=None, use_dill=True):
        self.MPI = _import_mpi(use_dill=use_dill)
        self.comm = self.MPI.COMM_WORLD if comm is None else comm
        self.master = 0
        self.rank = self.comm.Get_rank()
        atexit.register(lambda: MPIPool.close(self))
        if not self.is_master():
            self.wait()
            sys.exit(0)
        self.workers = set(range(self.comm.size))
        self.workers.discard(self.master)
        self.size = self.comm.Get_size() - 1
        if self.size == 0:
            raise ValueError("Tried to create an MPI pool, but there was only one MPI process available. "
                             "Need at least two.")

    def is_master(self):
        return self.rank == self.master

    def is_worker(self):
        return self.rank != self.master

    def wait(self):
        """Worker loop: receive (func, arg), reply with func(arg) under the same tag; None ends."""
        if self.is_master():
            return
        status = self.MPI.Status()
        while True:
            task = self.comm.recv(source=self.master, tag=self.MPI.ANY_TAG, status=status)
            if task is None:
                break
            func, arg = task
            self.comm.ssend(func(arg), self.master, status.tag)

    def map(self, worker, tasks):
        if not self.is_master():
            self.wait()
            return
        idle = self.workers.copy()
        todo = [(tid, (worker, arg)) for tid, arg in enumerate(tasks)]
        out = [None] * len(todo)
        waiting = len(todo)
        while waiting:
            if idle and todo:
                dest = idle.pop()
                tid, task = todo.pop()
                self.comm.send(task, dest=dest, tag=tid)
            if todo:
                if not self.comm.Iprobe(source=self.MPI.ANY_SOURCE, tag=self.MPI.ANY_TAG):
                    continue
            else:
                self.comm.Probe(source=self.MPI.ANY_SOURCE, tag=self.MPI.ANY_TAG)
            status = self.MPI.Status()
            res = self.comm.recv(source=self.MPI.ANY_SOURCE, tag=self.MPI.ANY_TAG, status=status)
            out[status.tag] = res
            idle.add(status.source)
            waiting -= 1
        return out

    def close(self):
        if self.is_worker():
            return
        for w in self.workers:
            self.comm.send(None, w, 0)

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
