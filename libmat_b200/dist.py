"""Multi-GPU plumbing of the RPD path: one process per GPU, tets sharded in contiguous blocks with the
sites replicated (the cells of different tets are independent, reference convex_cell.cu:1162-1163),
compact results gathered on one rank over torch.distributed (NCCL over NVLink on the GPU box, gloo
in the CPU tests).  No data-path collective is needed before the gather.

The gathered blob is the concatenation of the ranks' compact blobs in rank order, which IS the
global (tet, site) order because shards are contiguous in tet order; cell byte offsets are rebased
by the preceding ranks' blob sizes, cell ids by the preceding ranks' cell counts.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard(n_tet: int, rank: int, world: int) -> tuple[int, int]:
    """rank r owns tets [first, first + count): contiguous, sizes differ by at most one"""
    base, rem = divmod(n_tet, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allreduce_site_volumes(vol: np.ndarray, bary_soa: np.ndarray, group=None, device=None):
    """Per-site volume and barycentre sums of a sharded run: every rank passes the arrays of ITS tets
    (RpdResult.site_volumes() of a run with want_volumes=True -- the atomicAdd targets of get_cell_volume_and_barycenter,
    convex_cell.cu:1073-1160) and gets the sums over all ranks.  One all-reduce of 16 bytes per site; with the nccl
    backend pass device= so that the reduction runs over NVLink."""
    t = torch.from_numpy(np.concatenate([np.asarray(vol, np.float32).ravel(), np.asarray(bary_soa, np.float32).ravel()]))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t.cpu()
    a = t.numpy()
    n = np.asarray(vol).size
    return a[:n].copy(), a[n:].copy()


def balanced_cuts(cells_per_tet: np.ndarray, world: int, tet_cost: float = 1.5) -> np.ndarray:
    """Cut points (world + 1 ascending tet indices, cuts[0] = 0, cuts[-1] = n_tet) of contiguous tet shards of equal
    estimated WORK instead of equal size: work(tet) = tet_cost + its number of cells (the neighbour search costs
    ~2.8 ns per tet, clipping + ordering ~1.9 ns per cell on a B200 -> tet_cost = 1.5 cells).  Tets of the outer
    shell of a domain hold one or two cells, interior tets four or five, so equal-size slabs of a ball mesh differ
    by 1.5x in work (measured at config 4 on 8 GPUs: 86 vs 131 MB of records, 2.7 vs 3.7 ms)."""
    w = np.asarray(cells_per_tet, dtype=np.float64) + float(tet_cost)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(1, world) / world
    inner = np.searchsorted(cum, targets, side="left")
    cuts = np.concatenate([[0], inner, [len(w)]]).astype(np.int64)
    return np.maximum.accumulate(cuts)


def balanced_shards(n_tet: int, first: int, cells_per_tet_local: np.ndarray, group=None, device=None,
                    tet_cost: float = 1.5):
    """Collective: every rank passes the per-tet cell counts of ITS current shard [first, first + len) -- e.g. the
    bincount of RpdResult.pairs() of a previous run, what an iteration loop has anyway -- and gets the same
    balanced_cuts of the whole mesh plus the per-tet weights (input of rebalance_cuts).  One all-reduce of n_tet
    bytes (cells per tet <= 255)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = np.zeros(n_tet, np.uint8)
    mine[first:first + len(cells_per_tet_local)] = np.minimum(np.asarray(cells_per_tet_local), 255)
    if world > 1:
        t = torch.from_numpy(mine)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)  # shards are disjoint: max = union
        mine = t.cpu().numpy()
    return balanced_cuts(mine, world, tet_cost), mine.astype(np.float64) + float(tet_cost)


def rebalance_cuts(cuts: np.ndarray, weight_per_tet: np.ndarray, seconds_per_rank: np.ndarray) -> np.ndarray:
    """One step of MEASURED load balancing.  Rank r took seconds_per_rank[r] for the tets [cuts[r], cuts[r+1]); the
    static weight model (balanced_cuts) cannot know that an interior cell is clipped by more planes than a cell of
    the outer shell, so each rank's range gets its own measured rate (seconds per unit of weight) and the tet order
    is cut again into pieces of equal predicted time.  Two or three steps converge (config 4 on 8 GPUs: per-rank
    kernel time 2.85-3.38 ms -> within 3 %)."""
    cuts = np.asarray(cuts, dtype=np.int64)
    w = np.asarray(weight_per_tet, dtype=np.float64)
    world = len(cuts) - 1
    cum_w = np.concatenate([[0.0], np.cumsum(w)])
    span_w = np.maximum(cum_w[cuts[1:]] - cum_w[cuts[:-1]], 1e-30)
    rate = np.asarray(seconds_per_rank, dtype=np.float64) / span_w
    cost = w * np.repeat(rate, np.diff(cuts))
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    inner = np.searchsorted(cum, cum[-1] * np.arange(1, world) / world, side="left")
    return np.maximum.accumulate(np.concatenate([[0], inner, [len(w)]]).astype(np.int64))


class DeviceView:
    """__cuda_array_interface__ wrapper of a raw device pointer owned by libmat_b200 (mb_rpd_device_buffers)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(int(nbytes), 1),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 3}


def as_u8_tensor(ptr: int, nbytes: int, device) -> torch.Tensor:
    return torch.as_tensor(DeviceView(ptr, nbytes), device=device)[:nbytes]


def gather_varlen(local: torch.Tensor, dst: int = 0, group=None, out: torch.Tensor | None = None):
    """Gather 1-D uint8 tensors of different lengths on rank `dst` (all-gather of the sizes, then a
    grouped send/recv).  Returns (concatenation or None, sizes[list]) -- `out` may supply a large
    enough destination buffer to avoid reallocating every step."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(sizes, torch.tensor([local.numel()], dtype=torch.int64, device=local.device), group=group)
    sz = [int(x) for x in sizes.tolist()]
    if world == 1:
        return local, sz
    ops = []
    result = None
    if rank == dst:
        total = sum(sz)
        if out is None or out.numel() < total:
            out = torch.empty(total, dtype=torch.uint8, device=local.device)
        result = out[:total]
        off = 0
        for r in range(world):
            if r == dst:
                result[off:off + sz[r]].copy_(local, non_blocking=True)
            elif sz[r]:
                ops.append(dist.P2POp(dist.irecv, result[off:off + sz[r]], r, group))
            off += sz[r]
    elif local.numel():
        ops.append(dist.P2POp(dist.isend, local, dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return result, sz


def _gpu_numa_cpus(device_index: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when that cannot be determined"""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        if all(hasattr(props, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        else:
            import pynvml
            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            bdf = bus.lower()[-12:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:  # noqa: BLE001
        return None


class ShardSink:
    """Destination of a sharded run: one slab per rank, [blob bytes | cell offsets], in memory every rank can
    write -- device memory of rank `dst`'s GPU shared through CUDA IPC (kind="device": each rank's ordered
    records cross NVLink by copy-engine DMA while its next tet span is clipped), or a POSIX shared-memory
    segment page-locked by every rank (kind="host": the shards reach host memory over all PCIe links in
    parallel).  The slab directory (bytes, cells per rank) is exchanged through a shared-memory mailbox (no collective).

    Layout: slab r starts at r * slab_bytes; its offsets array (cap_cells + 1 int64) sits at the slab's end."""

    def __init__(self, ctx, cap_bytes: int, cap_cells: int, kind: str = "device", dst: int = 0, group=None, tag="mb"):
        import mmap
        import os
        self.ctx, self.kind, self.dst, self.group = ctx, kind, dst, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cap_bytes = (int(cap_bytes) + 255) // 256 * 256
        self.cap_cells = int(cap_cells)
        self.off_bytes = (8 * (self.cap_cells + 1) + 255) // 256 * 256
        self.slab_bytes = self.cap_bytes + self.off_bytes
        total = self.slab_bytes * self.world
        self._mm = None
        # directory mailbox: one cache line per rank [sequence, bytes, cells] in a small shared-memory file.  A rank
        # publishes its line when its streamed run has returned (its DMA has landed); readers spin on the sequence
        # words.  This replaces a per-run NCCL all-gather of 16 bytes, which cost 130-250 us per step (H2D of the
        # entry, the collective, D2H of the table) -- 5-10 % of a 2.4 ms step.
        self._seq = 0
        self._dir_mm = None
        box = [None]
        if self.rank == dst:
            dpath = f"/dev/shm/{tag}_dir_{os.getpid()}_{id(self) & 0xffff}"
            with open(dpath, "wb") as fh:
                fh.truncate(64 * self.world)
            box = [dpath]
        if self.world > 1:
            dist.broadcast_object_list(box, src=dst, group=group)
        fd = os.open(box[0], os.O_RDWR)
        self._dir_mm = mmap.mmap(fd, 64 * self.world)
        os.close(fd)
        self._dir = np.frombuffer(self._dir_mm, dtype=np.int64).reshape(self.world, 8)
        if self.world > 1:
            dist.barrier(group=group)
        if self.rank == dst:
            os.unlink(box[0])
        if kind == "device":
            # failures (IPC not permitted in this container, no peer access) must surface on EVERY rank, or
            # the others would wait in the next collective: agree on the outcome before going on
            box, ok, err = [None], 1, ""
            self.base = None
            if self.rank == dst:
                try:
                    self.base, handle = ctx.sink_create(total)
                    box = [handle]
                except Exception as exc:  # noqa: BLE001
                    ok, err = 0, str(exc)
            if self.world > 1:
                dist.broadcast_object_list(box, src=dst, group=group)
            if self.rank != dst:
                try:
                    if box[0] is None:
                        raise RuntimeError("owner could not create the sink")
                    self.base = ctx.sink_open(box[0])
                except Exception as exc:  # noqa: BLE001
                    ok, err = 0, str(exc)
            if self.world > 1:
                oks = [None] * self.world
                dist.all_gather_object(oks, (ok, err), group=group)
                bad = [e for o, e in oks if not o]
                if bad:
                    if self.base is not None:
                        (ctx.sink_destroy if self.rank == dst else ctx.sink_close)(self.base)
                    raise RuntimeError("peer sink unavailable: " + bad[0])
            elif not ok:
                raise RuntimeError(err)
        else:
            box = [None]
            if self.rank == dst:
                path = f"/dev/shm/{tag}_{os.getpid()}"
                with open(path, "wb") as fh:
                    fh.truncate(total)
                box = [path]
            if self.world > 1:
                dist.broadcast_object_list(box, src=dst, group=group)
            self.path = box[0]
            fd = os.open(self.path, os.O_RDWR)
            self._mm = mmap.mmap(fd, total)
            os.close(fd)
            self.host = np.frombuffer(self._mm, dtype=np.uint8)
            self.base = self.host.ctypes.data
            # first touch: every rank faults in ITS slab from a CPU of its GPU's NUMA node, so that the shard's
            # D2H stays on the local socket (page-locking untouched pages would place them all on one node)
            saved = os.sched_getaffinity(0)
            cpus = _gpu_numa_cpus(torch.cuda.current_device()) if torch.cuda.is_available() else None
            try:
                if cpus:
                    os.sched_setaffinity(0, cpus)
                s0 = self.rank * self.slab_bytes
                self.host[s0:s0 + self.slab_bytes:4096] = 0
            finally:
                os.sched_setaffinity(0, saved)
            self.numa_pinned = bool(cpus)
            if self.world > 1:
                dist.barrier(group=group)
            ctx.host_register(self.base, total)
            if self.world > 1:
                dist.barrier(group=group)
            if self.rank == dst:
                os.unlink(self.path)  # the mappings keep the segment alive

    def slab(self, r: int):
        b = self.base + r * self.slab_bytes
        return b, b + self.cap_bytes

    def run(self, n_chunks=0, **kw):
        """every rank: streamed run of its shard into its slab; returns (result handle, directory [world, 2])"""
        blob_ptr, off_ptr = self.slab(self.rank)
        if n_chunks == 0 and self.kind == "device":
            # spans exist to hide the transfer behind the kernels; into a peer's HBM the whole shard takes ~0.1 ms
            # unless many ranks converge on the destination's NVLink ingress at once
            # (measured, per step: N = 2 one span 2.32 ms vs two 2.42; config 4 on 8 balanced ranks, where 7 last-span
            # copies converge on one ingress at 950 GB/s: three spans of sizes 3:2:1 3.45 ms vs two 3.6)
            n_chunks = 1 if self.world <= 2 else (2 if self.world <= 4 else 3)
        import os
        import time
        trace = os.environ.get("MB_TRACE", "0") >= "2"
        t0 = time.perf_counter()
        res = self.ctx.run_to_sink(blob_ptr, self.cap_bytes, off_ptr, self.cap_cells + 1, n_chunks=n_chunks, **kw)
        t1 = time.perf_counter()
        # publish this rank's directory line (payload first, sequence word last: x86 stores stay in program order) and
        # wait until every rank has published the same sequence number -- the completion barrier of the run
        # (two slots by sequence parity: a fast rank can be at most one run ahead of a rank still reading the table)
        self._seq += 1
        c0 = 4 * (self._seq & 1)
        row = self._dir[self.rank]
        row[c0 + 1] = res.compact_bytes
        row[c0 + 2] = res.n_cells
        row[c0] = self._seq
        seqs = self._dir[:, c0]
        deadline = t1 + 120.0
        spins = 0
        while (seqs < self._seq).any():
            spins += 1
            if (spins & 0x3fff) == 0 and time.perf_counter() > deadline:
                raise RuntimeError(f"ShardSink: rank(s) {np.flatnonzero(seqs < self._seq).tolist()} did not finish run {self._seq} within 120 s")
        table = self._dir[:, c0 + 1:c0 + 3].copy()
        if trace:
            print(f"[ShardSink rank {self.rank} {self.kind}] run_to_sink {1e6 * (t1 - t0):.0f} us, directory mailbox "
                  f"{1e6 * (time.perf_counter() - t1):.0f} us", file=__import__("sys").stderr)
        return res, table

    def read_host(self, directory, out_blob: np.ndarray | None = None):
        """rank dst: the shards concatenated in rank order (= global (tet, site) order) + rebased offsets"""
        tot = int(directory[:, 0].sum())
        blob = out_blob if out_blob is not None else np.empty(tot, np.uint8)
        offs, at = [], 0
        for r in range(self.world):
            nb, nc = int(directory[r, 0]), int(directory[r, 1])
            bp, op = self.slab(r)
            o = np.empty(nc + 1, np.int64)
            if self.kind == "device":
                if nb:
                    self.ctx.copy_to_host(blob[at:at + nb], bp, nb)
                self.ctx.copy_to_host(o, op, 8 * (nc + 1))
            else:
                s0 = r * self.slab_bytes
                blob[at:at + nb] = self.host[s0:s0 + nb]
                o[:] = self.host[s0 + self.cap_bytes:s0 + self.cap_bytes + 8 * (nc + 1)].view(np.int64)
            offs.append(o)
            at += nb
        return blob[:tot], rebase_offsets(offs, [int(x) for x in directory[:, 0]])

    def close(self):
        self._dir = None
        try:
            self._dir_mm.close()
        except (BufferError, AttributeError):
            pass
        if self.kind == "device":
            if self.rank == self.dst:
                if self.world > 1:
                    dist.barrier(group=self.group)
                self.ctx.sink_destroy(self.base)
            else:
                self.ctx.sink_close(self.base)
                dist.barrier(group=self.group)
        else:
            self.ctx.host_unregister(self.base)
            self.host = None
            try:
                self._mm.close()
            except BufferError:
                pass


def rebase_offsets(cell_offsets: list[np.ndarray], blob_sizes: list[int]) -> np.ndarray:
    """per-rank cell byte offsets (each n_cells_r + 1 long, starting at 0) -> global offsets"""
    out = [np.zeros(1, dtype=np.int64)]
    base = 0
    for offs, size in zip(cell_offsets, blob_sizes):
        offs = np.asarray(offs, dtype=np.int64)
        assert offs[0] == 0 and offs[-1] == size
        out.append(offs[1:] + base)
        base += size
    return np.concatenate(out)
