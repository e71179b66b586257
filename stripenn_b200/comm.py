"""Exchanges of the sharded `compute` pipeline.

The path needs no data-path collective: every (chromosome, window range) unit is independent.
What the ranks do exchange is small and typed — never pickled Python objects:

* per-chromosome availability counts (one int64 per background chromosome),
* the background columns: each GPU writes the columns of its chromosomes into its slot range of
  a device buffer and ONE all-gather (NCCL over NVLink, in place) completes the buffer on every
  GPU; the columns never visit the host,
* the raw candidates of a chromosome that was cut into window ranges (fixed-dtype records),
* the final result rows (fixed-dtype records).

A *virtual rank* is one host thread with its own library context; `tpp` of them share a GPU
(one process per GPU).  Four communicators implement the same interface:

    LocalComm    one virtual rank
    ThreadComm   `tpp` threads of one process (shared memory, a barrier)
    TorchComm    one virtual rank per process, torch.distributed (nccl or gloo)
    HybridComm   `tpp` threads in each of N processes: threads meet in shared memory, thread 0 of
                 every process talks to the other processes

Interface: rank, world, proc, nproc, tpp; all_gather_arrays(a) -> one array per virtual rank
(same dtype and trailing shape everywhere, any length); once(fn) -> fn() evaluated once per
process, every thread of the process gets the object; bg_buffer(dev, nslots) -> device pointer of
this process's background buffer; bg_exchange(per_proc) -> completes it on every process.
"""
from __future__ import annotations

import threading

import numpy as np

BG_SLOT = 1600          # doubles per background sample: 4 tables x 400 distances


class LocalComm:
    """One virtual rank: nothing to exchange."""
    rank = 0
    world = 1
    proc = 0
    nproc = 1
    tpp = 1

    def __init__(self):
        self._bg = None

    def all_gather_arrays(self, a):
        return [np.asarray(a)]

    all_gather_fixed = all_gather_arrays

    def once(self, fn):
        return fn()

    def bg_buffer(self, dev, nslots):
        nbytes = max(int(nslots), 1) * BG_SLOT * 8
        if self._bg is None or self._bg[1] < nbytes or self._bg[2] is not dev:
            self.bg_release()
            self._bg = (dev.device_alloc(nbytes), nbytes, dev)
        return self._bg[0]

    def bg_exchange(self, per_proc):
        pass

    def bg_release(self):
        if self._bg is not None:
            try:
                self._bg[2].device_free(self._bg[0])
            except Exception:
                pass
            self._bg = None


class ThreadComm:
    """`world` host threads of one process feeding ONE GPU through separate contexts ("virtual
    ranks"): while one thread post-processes its chromosomes on the host (dedup, table assembly),
    the kernels of the others keep the GPU busy.  Create with ThreadComm.group(k)."""
    proc = 0
    nproc = 1

    def __init__(self, rank, world, shared):
        self.rank, self.world, self._sh = rank, world, shared
        self.tpp = world

    @classmethod
    def group(cls, world):
        shared = {"slots": [None] * world, "barrier": threading.Barrier(world), "bg": None}
        return [cls(r, world, shared) for r in range(world)]

    def all_gather(self, obj):
        sh = self._sh
        sh["slots"][self.rank] = obj
        sh["barrier"].wait()
        out = list(sh["slots"])
        sh["barrier"].wait()
        return out

    def all_gather_arrays(self, a):
        return self.all_gather(np.asarray(a))

    all_gather_fixed = all_gather_arrays

    def once(self, fn):
        """fn() evaluated by thread 0 only; every thread gets the same object."""
        sh = self._sh
        if self.rank == 0:
            try:
                sh["once"] = (fn(), None)
            except BaseException as e:  # noqa: BLE001 - re-raised on every thread
                sh["once"] = (None, e)
        sh["barrier"].wait()
        val, err = sh["once"]
        sh["barrier"].wait()
        if err is not None:
            raise err
        return val

    def barrier(self):
        self._sh["barrier"].wait()

    def bg_buffer(self, dev, nslots):
        nbytes = max(int(nslots), 1) * BG_SLOT * 8

        def make():
            bg = self._sh["bg"]
            if bg is None or bg[1] < nbytes:
                if bg is not None:
                    bg[2].device_free(bg[0])
                self._sh["bg"] = (dev.device_alloc(nbytes), nbytes, dev)
            return self._sh["bg"][0]
        return self.once(make)

    def bg_exchange(self, per_proc):
        self.barrier()          # every thread's columns are in the buffer


class TorchComm:
    """One virtual rank per process over torch.distributed.  With the nccl backend the payloads
    travel as CUDA tensors (the background buffer is a torch allocation that NCCL gathers in
    place); with gloo (CPU tests) as CPU tensors."""
    tpp = 1

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = self.proc = dist.get_rank()
        self.world = self.nproc = dist.get_world_size()
        self.cuda = dist.get_backend() == "nccl"
        self.dev = torch.device("cuda", torch.cuda.current_device()) if self.cuda else torch.device("cpu")
        self._bg = None
        # The small host arrays (counts, candidate and result rows) travel over a gloo group even when
        # the default backend is nccl: an NCCL collective is a kernel that has to find room on GPUs the
        # searches keep full, on all ranks at once — measured at 8 GPUs as milliseconds per exchange
        # (profiles/bench_r02p_n8.json: the first exchange completed 23 ms into a 49 ms step).  Only
        # the background buffer, which lives on the device, goes through NCCL.
        self.host_group = dist.new_group(backend="gloo") if self.cuda and self.nproc > 1 else None

    def _gather_equal(self, a):
        """a: 1-D array of the same length and dtype on every process -> (nproc, len)."""
        torch = self.torch
        t = torch.from_numpy(np.ascontiguousarray(a).reshape(-1))
        if self.host_group is None and self.cuda:
            t = t.to(self.dev)
        out = torch.empty(self.nproc * t.numel(), dtype=t.dtype, device=t.device)   # flat: gloo insists
        self.dist.all_gather_into_tensor(out, t, group=self.host_group)
        return out.cpu().numpy().reshape(self.nproc, -1)

    def all_gather_fixed(self, a):
        """Arrays of the SAME shape and dtype on every process: one collective."""
        a = np.ascontiguousarray(a)
        got = self._gather_equal(a.reshape(-1).view(np.uint8))
        return [got[p].copy().view(a.dtype).reshape(a.shape) for p in range(self.nproc)]

    def all_gather_arrays(self, a):
        a = np.ascontiguousarray(a)
        sizes = self._gather_equal(np.array([a.shape[0]], dtype=np.int64))[:, 0]
        row = int(np.prod(a.shape[1:], dtype=np.int64)) * a.dtype.itemsize
        cap = int(sizes.max()) * row
        if cap == 0:
            return [np.zeros((0,) + a.shape[1:], dtype=a.dtype) for _ in range(self.nproc)]
        buf = np.zeros(cap, dtype=np.uint8)
        buf[:a.shape[0] * row] = a.reshape(-1).view(np.uint8)
        got = self._gather_equal(buf)
        return [got[p, :int(sizes[p]) * row].copy().view(a.dtype).reshape((int(sizes[p]),) + a.shape[1:])
                for p in range(self.nproc)]

    def once(self, fn):
        return fn()

    def bg_buffer(self, dev, nslots):
        n = max(int(nslots), 1) * BG_SLOT
        if self._bg is None or self._bg.numel() < n:
            self._bg = self.torch.empty(n, dtype=self.torch.float64, device=self.dev)
        return int(self._bg.data_ptr())

    def bg_exchange(self, per_proc):
        """In-place all-gather: process p's slot range [p * per_proc, (p + 1) * per_proc) goes to
        every process (ncclAllGather straight between the buffers k_background wrote and
        k_bg_sort reads)."""
        n = int(per_proc) * BG_SLOT
        if n == 0 or self.nproc == 1:
            return
        full = self._bg[:n * self.nproc]
        mine = full[self.proc * n:(self.proc + 1) * n]
        try:
            self.dist.all_gather_into_tensor(full, mine)
        except (RuntimeError, ValueError):      # a backend that refuses aliased buffers
            self.dist.all_gather_into_tensor(full, mine.clone())
        if self.cuda:
            self.torch.cuda.current_stream().synchronize()


class HybridComm:
    """`tpp` host threads in each of `nproc` processes (one process per GPU): virtual rank =
    proc * tpp + thread.  Threads meet through a ThreadComm; thread 0 of each process performs the
    inter-process exchange through a TorchComm.  Create one per thread with HybridComm.group."""

    def __init__(self, tc: ThreadComm, pc: TorchComm):
        self.tc, self.pc = tc, pc
        self.tpp = tc.world
        self.proc, self.nproc = pc.proc, pc.nproc
        self.rank = pc.proc * tc.world + tc.rank
        self.world = pc.nproc * tc.world

    @classmethod
    def group(cls, tpp, pc: TorchComm):
        return [cls(tc, pc) for tc in ThreadComm.group(tpp)]

    def all_gather_arrays(self, a):
        a = np.ascontiguousarray(a)
        local = self.tc.all_gather(a)

        def across():
            # two collectives: the per-thread lengths, then the rows padded to the longest process
            lens = np.array([x.shape[0] for x in local], dtype=np.int64)
            all_lens = self.pc._gather_equal(lens)                       # (nproc, tpp)
            row = int(np.prod(a.shape[1:], dtype=np.int64)) * a.dtype.itemsize
            cap = int(all_lens.sum(axis=1).max()) * row
            if cap == 0:
                return [np.zeros((0,) + a.shape[1:], dtype=a.dtype) for _ in range(self.world)]
            buf = np.zeros(cap, dtype=np.uint8)
            mine = np.concatenate(local).reshape(-1).view(np.uint8)
            buf[:mine.shape[0]] = mine
            got = self.pc._gather_equal(buf)
            out = []
            for p in range(self.nproc):
                off = 0
                for t in range(self.tpp):
                    n = int(all_lens[p, t]) * row
                    out.append(got[p, off:off + n].copy().view(a.dtype).reshape((-1,) + a.shape[1:]))
                    off += n
            return out
        return self.tc.once(across)

    def all_gather_fixed(self, a):
        """Same shape on every virtual rank: threads meet in shared memory, one collective."""
        a = np.ascontiguousarray(a)
        local = self.tc.all_gather(a)

        def across():
            parts = self.pc.all_gather_fixed(np.stack(local))            # (tpp, ...) per process
            return [parts[p][t] for p in range(self.nproc) for t in range(self.tpp)]
        return self.tc.once(across)

    def once(self, fn):
        return self.tc.once(fn)

    def bg_buffer(self, dev, nslots):
        return self.tc.once(lambda: self.pc.bg_buffer(dev, nslots))

    def bg_exchange(self, per_proc):
        self.tc.barrier()                                   # every thread's columns are in the buffer
        self.tc.once(lambda: self.pc.bg_exchange(per_proc))


def run_threads(fn, comms):
    """Run fn(comm) on one host thread per communicator (ThreadComm or HybridComm of one group);
    returns the per-thread results and re-raises the first exception."""
    if isinstance(comms, int):
        comms = ThreadComm.group(comms)
    world = len(comms)
    res, err = [None] * world, [None] * world

    def barrier_of(c):
        return (c.tc if isinstance(c, HybridComm) else c)._sh["barrier"]

    def work(i):
        try:
            res[i] = fn(comms[i])
        except BaseException as e:  # noqa: BLE001 - re-raised below
            err[i] = e
            barrier_of(comms[i]).abort()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return res


def background_layout(names2b, n_of, owner, tpp, nproc):
    """Slot layout of the background buffer: process p owns the slot range
    [p * per_proc, (p + 1) * per_proc); inside it the chromosomes whose primary virtual rank lives
    in p follow each other in `names2b` order.  Pure function of its arguments, so every rank
    derives the same layout.  Returns (slot0 per chromosome, per_proc, slots of all columns in
    names2b order)."""
    used = [0] * nproc
    slot_in_proc = {}
    for c in names2b:
        p = owner[c] // tpp
        slot_in_proc[c] = used[p]
        used[p] += int(n_of[c])
    per_proc = max(used) if used else 0
    slot0 = {c: (owner[c] // tpp) * per_proc + slot_in_proc[c] for c in names2b}
    slots = np.concatenate([np.arange(slot0[c], slot0[c] + int(n_of[c]), dtype=np.int32) for c in names2b]) \
        if names2b else np.zeros(0, np.int32)
    return slot0, per_proc, slots
