"""Multi-GPU use of the grid hot path: one process per GPU, ``torch.distributed`` for the plumbing.

The path shards two ways:

* **independent environments** (BASELINE config 3): :func:`shard_envs` deals environment indices to
  ranks; there is no data-path collective, only the final gather of whatever the caller collects;
* **row bands of one large world** (BASELINE config 4): every rank holds the replicated robot state
  and labels / reduces rows ``band_rows(N, world_size, rank)`` of the map only.  Per step the
  per-robot moment partials ``[E][R][4]`` float64 are combined with ONE ``all_reduce(SUM)`` (NCCL over
  NVLink on the GPU box, gloo in the CPU tests); IDF generation is per band with one scalar
  ``all_reduce(MAX)`` for the normalisation factor and one ``all_gather`` of the bands (once, at
  set-up) so that window crops near a band edge need no halo exchange.

The reference has no multi-process form of this path (its Voronoi, voronoi.cpp:55-327, and its
data generation, python/scripts/data_generation/data_generation.py:217-259, are single-process loops).
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["band_rows", "shard_envs", "float_to_ordered", "ordered_to_float", "allreduce_sum_", "allreduce_max_",
           "BandedWorldIDF", "BandedCoverageSystem", "PeerExchange"]


def band_rows(n_rows: int, world_size: int, rank: int, align: int = 64) -> tuple[int, int]:
    """Rows ``[row0, row1)`` of rank ``rank``: contiguous, disjoint, covering ``[0, n_rows)``; band edges are
    multiples of ``align`` (the Voronoi kernel's region height) whenever ``n_rows`` allows it."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    units = (n_rows + align - 1) // align
    if units < world_size:  # fewer aligned chunks than ranks: fall back to single rows
        align, units = 1, n_rows
    lo = (units * rank) // world_size
    hi = (units * (rank + 1)) // world_size
    return min(lo * align, n_rows), min(hi * align, n_rows)


def shard_envs(total_envs: int, world_size: int, rank: int) -> range:
    """Environment indices of rank ``rank`` (contiguous blocks, sizes differ by at most one)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    lo = (total_envs * rank) // world_size
    hi = (total_envs * (rank + 1)) // world_size
    return range(lo, hi)


def float_to_ordered(x):
    """Order-preserving int32 image of float32 (the encoding the IDF kernel's atomicMax uses)."""
    i = np.asarray(x, np.float32).view(np.int32)
    return np.where(i >= 0, i, i ^ np.int32(0x7FFFFFFF)).astype(np.int32)


def ordered_to_float(i):
    i = np.asarray(i, np.int32)
    return np.where(i >= 0, i, i ^ np.int32(0x7FFFFFFF)).astype(np.int32).view(np.float32)


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def allreduce_sum_(t):
    """In-place SUM over ranks (no-op in a single process).  ``t``: torch tensor on the backend's device."""
    dist = _dist()
    if dist is not None and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_max_(t):
    dist = _dist()
    if dist is not None and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def _rank_world():
    dist = _dist()
    return (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)


class BandedWorldIDF:
    """WorldIDF generation tiled in row bands over the ranks of the default process group.

    Each rank integrates the features over its rows only (generate_world_map.cu:76-164 per cell),
    the map maximum is all-reduced (one int32), every band is normalised with the global factor
    (generate_world_map.cu:252-259) and the bands are all-gathered so that every rank ends up with
    the full, bit-identical map."""

    def __init__(self, idf, mode: int = 0):
        import torch

        self.idf = idf
        rank, world = _rank_world()
        n = idf._params.pWorldMapSize
        self.rows = band_rows(n, world, rank)
        lib = _lib.load()
        r0, r1 = self.rows
        _lib.check(lib.ccb_idf_generate_band(idf._h, int(mode), r0, r1))
        dev = torch.device("cuda", idf._device)
        mx = _wrap_dev(lib.ccb_idf_band_max_dev(idf._h), (1,), torch.int32, dev)
        allreduce_max_(mx)
        _lib.check(lib.ccb_idf_normalize_band(idf._h, r0, r1))
        if world > 1:
            import torch.distributed as dist

            full = _wrap_dev(lib.ccb_idf_map_dev_mut(idf._h), (n, n), torch.float32, dev)
            bounds = [band_rows(n, world, g) for g in range(world)]
            if len({b[1] - b[0] for b in bounds}) == 1:
                dist.all_gather_into_tensor(full, full[r0:r1].clone())
            else:  # ragged bands: one broadcast per band
                for g, (a, b) in enumerate(bounds):
                    if b > a:
                        dist.broadcast(full[a:b], src=g)
        idf._mutable = None


def _wrap_dev(ptr: int, shape, dtype, device):
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4"}[dtype]
    h.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                  "strides": None}
    return torch.as_tensor(h, device=device)


class PeerExchange:
    """The step's all-reduce(SUM) as ONE kernel over peer memory (``csrc/peer.cuh``, ``ccb_peer_allreduce_sum``).

    The partials of a step are E x R x 4 doubles (8 KB at R = 256): latency, not bandwidth.  Every rank allocates the
    same exchange buffer from torch's symmetric-memory allocator, the rendezvous maps every peer's buffer into this
    process (NVLink peer access), and the kernel stores / flags / waits / adds in rank order -- no NCCL call on the
    data path, bit-identical sums on every rank and from run to run.  :meth:`create` returns ``None`` (on EVERY rank
    alike) when the buffers cannot be set up -- single process, CPU backend, no peer access -- and the caller keeps
    NCCL's ``all_reduce``."""

    def __init__(self, hdl, buf, state, count, device, rank, world):
        self._hdl, self._buf, self._state = hdl, buf, state  # keep the mappings alive
        self.count, self.device, self.rank, self.world = count, device, rank, world
        import ctypes as C

        self._ptrs = (C.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])

    @classmethod
    def create(cls, count: int, device: int):
        import torch

        dist = _dist()
        if dist is None or dist.get_world_size() < 2 or dist.get_backend() != "nccl":
            return None
        rank, world = dist.get_rank(), dist.get_world_size()
        lib = _lib.load()
        dev = torch.device("cuda", device)
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        hdl = buf = None
        try:
            import torch.distributed._symmetric_memory as symm_mem

            nbytes = int(lib.ccb_peer_buffer_bytes(world, count))
            buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            if len(hdl.buffer_ptrs) != world or not all(int(p) for p in hdl.buffer_ptrs):
                ok.zero_()
        except Exception:  # noqa: BLE001 - any failure here means "no peer path": all ranks must still agree below
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # also orders the zero_() of every rank before the first exchange
        torch.cuda.synchronize(dev)
        if int(ok.item()) == 0:
            return None
        state = torch.zeros(2, dtype=torch.int64, device=dev)
        dist.barrier()
        return cls(hdl, buf, state, count, device, rank, world)

    def allreduce_sum_(self, t, stream=None):
        """In-place rank-ordered SUM of the float64 tensor ``t`` (``count`` elements) on ``stream``."""
        import torch

        if t.numel() != self.count or t.dtype != torch.float64 or not t.is_contiguous():
            raise ValueError("PeerExchange.allreduce_sum_: tensor does not match the exchange buffer")
        st = stream if stream is not None else torch.cuda.current_stream(t.device)
        _lib.check(_lib.load().ccb_peer_allreduce_sum(self.device, st.cuda_stream, t.data_ptr(), t.data_ptr(), self.count,
                                                      self._ptrs, self.rank, self.world, self._state.data_ptr()))
        return t

    def check(self) -> None:
        """Raise if a peer failed to arrive in some exchange (the kernel gives up after ~3 s instead of hanging)."""
        import torch

        torch.cuda.synchronize(self.device)
        done, bad = (int(v) for v in self._state.tolist())
        if bad:
            raise RuntimeError(f"peer all-reduce: exchange {bad} of {done} timed out waiting for a peer")


class BandedCoverageSystem:
    """One large world tiled over the ranks (BASELINE config 4).

    Every rank constructs the same :class:`BatchedCoverageSystem` (replicated positions and maps; the controllers are
    deterministic, so the replicas stay bit-identical) and shares out the *work* of a step:

    * ClairvoyantCVT / CentralizedCVT / the objective: rank g labels and reduces map rows ``band_rows(N, G, g)`` only
      (``ccb_env_voronoi_partials_dev``), the per-robot moment partials ``[E][R][4]`` float64 are combined with ONE
      ``all_reduce(SUM)`` and every rank finalises the cells (``ccb_env_finalize_from_partials_dev``);
    * DecentralizedCVT: the robots' local windows are independent, so they are dealt out by ROBOT: rank g reduces the
      windows of robots ``shard_envs(R, G, g)`` (``ccb_env_window_partials_dev``), the same all-reduce, and every rank
      takes the result as its window moments (``ccb_env_set_window_sums_dev``).

    Memory is NOT sharded: every rank holds the whole world map and all robot bit planes (``per_rank_bytes``); a band
    owner would otherwise need a halo of L/2 rows for the windows of robots near its band edge.

    ``graph=True``: a step (partials, the NCCL all-reduce, finalisation, actions, kinematics, window marks) is captured
    into ONE CUDA graph on its second call and replayed afterwards -- the environment must then have been created on
    ``stream`` (a ``torch.cuda.Stream``), which the all-reduce is issued on as well.
    """

    def __init__(self, env, src: int = 0, stream=None, graph: bool = False, collective: str = "auto"):
        import torch

        if collective not in ("auto", "peer", "nccl"):
            raise ValueError("collective: 'auto', 'peer' or 'nccl'")

        self.env = env
        self.src = int(src)
        rank, world = _rank_world()
        self.rank, self.world = rank, world
        self.rows = band_rows(env._params.pWorldMapSize, world, rank)
        r = shard_envs(env._R, world, rank)
        self.robots = (r.start, r.stop)
        dev = torch.device("cuda", env._device)
        self._sums = torch.zeros((env._E, env._R, 4), dtype=torch.float64, device=dev)
        self._obj = torch.zeros((env._E,), dtype=torch.float64, device=dev)
        self.stream = stream
        if graph and stream is None:
            raise ValueError("graph=True needs the torch.cuda.Stream the environment was created on")
        if stream is not None and int(stream.cuda_stream) != int(env._stream):
            raise ValueError("the environment was not created on `stream`")
        self.graph = bool(graph)
        self._graphs = {}
        self._seen = {}
        _lib.check(_lib.load().ccb_env_set_band_mode(env._h, 1))
        # the all-reduce of a step: one kernel over peer memory where the ranks can map each other's buffers,
        # NCCL (gloo in the CPU tests) otherwise; `collective` says which one is in use
        self._peer = PeerExchange.create(self._sums.numel(), env._device) if collective != "nccl" else None
        if collective == "peer" and self._peer is None and world > 1:
            raise RuntimeError("collective='peer': the peer exchange buffers could not be set up")
        self.collective = "peer" if self._peer is not None else ("nccl" if world > 1 else "none")

    def _allreduce(self, t):
        if self._peer is not None:
            return self._peer.allreduce_sum_(t, self.stream)
        return allreduce_sum_(t)

    def close(self) -> None:
        """Release the captured graphs.  Call before ``torch.distributed.destroy_process_group()``: a graph that
        holds a captured NCCL all-reduce keeps the communicator busy and the teardown waits for it forever."""
        import torch

        if self.stream is not None:
            self.stream.synchronize()
        self._graphs.clear()
        self._seen.clear()
        torch.cuda.synchronize()
        if self._peer is not None:
            self._peer.check()

    @property
    def per_rank_bytes(self) -> int:
        """Device bytes of the replicated state: world map, robot bit planes (+ exploration planes), system planes."""
        n, r, e = self.env._params.pWorldMapSize, self.env._R, self.env._E
        planes = r * (2 if self.env._params.pUpdateExplorationMap else 1) + 2
        return e * (4 * n * n + planes * n * ((n + 31) // 32) * 4)

    def _ctx(self):
        import contextlib

        import torch

        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def compute_cells(self, src: int | None = None):
        """voronoi.cpp:55-327 over the whole world: band partials, one all-reduce, finalisation.
        ``src`` 0: world map (ClairvoyantCVT, objective), 1: explored-IDF map (CentralizedCVT)."""
        lib = _lib.load()
        r0, r1 = self.rows
        src = self.src if src is None else int(src)
        with self._ctx():
            _lib.check(lib.ccb_env_voronoi_partials_dev(self.env._h, src, r0, r1, self._sums.data_ptr()))
            self._allreduce(self._sums)
            _lib.check(lib.ccb_env_finalize_from_partials_dev(self.env._h, self._sums.data_ptr()))
        return self._sums

    def compute_window_moments(self):
        """decentralized_cvt.h:85-143 for every robot: this rank's robots, one all-reduce, handed to the controller."""
        lib = _lib.load()
        with self._ctx():
            _lib.check(lib.ccb_env_window_partials_dev(self.env._h, self.robots[0], self.robots[1], self._sums.data_ptr()))
            self._allreduce(self._sums)
            _lib.check(lib.ccb_env_set_window_sums_dev(self.env._h, self._sums.data_ptr()))
        return self._sums

    def _step_eager(self, ctrl, with_objective: bool):
        lib = _lib.load()
        with self._ctx():
            if ctrl._KIND == 2:
                self.compute_window_moments()
            else:
                self.compute_cells(1 if ctrl._KIND == 1 else 0)
            _lib.check(lib.ccb_ctrl_compute_actions(ctrl._h))
            _lib.check(lib.ccb_env_step_actions_dev(self.env._h, lib.ccb_ctrl_actions_dev(ctrl._h)))
            if with_objective:  # coverage_system.h:598-601: always on the world map
                self.compute_cells(0)
                _lib.check(lib.ccb_env_objective_dev(self.env._h, self._obj.data_ptr()))
        self.env._invalidate()

    def step(self, ctrl, with_objective: bool = False):
        """``ComputeActions(); StepActions(GetActions())`` of a CVT controller on the tiled world, device-resident.
        Returns the objective (E,) at the new positions if asked (a device tensor, valid once the stream has run)."""
        if not self.graph:
            self._step_eager(ctrl, with_objective)
            return self._obj if with_objective else None
        import torch

        key = (id(ctrl), bool(with_objective))
        g = self._graphs.get(key)
        if g is not None:
            with self._ctx():
                g.replay()  # (the library's version counters keep their relative state: nothing to update on the host)
            self.env._invalidate()
            return self._obj if with_objective else None
        n = self._seen.get(key, 0)
        self._seen[key] = n + 1
        if n < 1:  # first call: eager (allocates whatever the kernels need; NCCL warms up)
            self._step_eager(ctrl, with_objective)
            return self._obj if with_objective else None
        g = torch.cuda.CUDAGraph()
        self.stream.synchronize()
        with torch.cuda.graph(g, stream=self.stream):
            self._step_eager(ctrl, with_objective)
        self._graphs[key] = g
        with self._ctx():  # nothing ran during the capture: run the step this call asked for
            g.replay()
        self.env._invalidate()
        return self._obj if with_objective else None
