"""One process per GPU; instances shard contiguously by rank and never exchange data.
The only collective on the path is a SUM all-reduce of the int64 counters
(solved count, instance count): rewards are +-1, so the reward total is derived
exactly as 2*solved - count (an fp32 sum of +-1 over 64 Mi instances is not exact)."""
import os

import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous slice [lo, hi) of rank `rank`; the first n_total % world_size ranks get one extra."""
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns
    (rank, local_rank, world_size); world_size 1 needs no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def reduce_counters(counters):
    """SUM all-reduce of an int64 counter tensor across ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def reduce_counters_async(counters):
    """Start the SUM all-reduce and return a handle whose `.wait()` orders the current stream after
    it (None for a single process).  The collective runs on NCCL's own stream, so the next step's
    kernel -- which does not depend on it -- overlaps the reduction."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(counters, op=dist.ReduceOp.SUM, async_op=True)
    return None


class PeerCounters:
    """The counters' SUM all-reduce over NVLink peer memory (C ABI cube_peer_allreduce_i64, csrc/peer.cu) for the
    ranks of one box: every rank's exchange buffer is torch symmetric memory, mapped by all ranks once, here; an
    all-reduce is then ONE small kernel per rank on the current stream (no NCCL call, no host synchronisation).

        peer = PeerCounters(capacity=256)        # collective: every rank constructs it
        peer.allreduce_(counters)                # int64 CUDA tensor, <= capacity elements, in place

    Raises RuntimeError when symmetric memory cannot be set up (no process group, no peer access); callers fall
    back to reduce_counters (NCCL) -- both give the same totals."""

    def __init__(self, capacity=256, group=None):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerCounters needs an initialised process group")
        group = group if group is not None else dist.group.WORLD
        self._lib = _lib
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.capacity = int(capacity)
        nbytes = int(_lib.load().cube_peer_buffer_bytes(self.capacity))
        if nbytes <= 0:
            raise RuntimeError("cube_peer_buffer_bytes(%d) failed" % self.capacity)
        dev = torch.device("cuda", torch.cuda.current_device())
        # the allocation is local and may fail on one rank only; the rendezvous behind it is a collective, so the
        # ranks first agree that every one of them got its buffer (all of them raise, or none)
        err = None
        try:
            self.buffer = symm_mem.empty(nbytes // 8, dtype=torch.int64, device=dev)
            self.buffer.zero_()
        except Exception as e:                            # noqa: BLE001 - reported below, on every rank
            err = e
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok[0]) == 0:
            raise RuntimeError("symmetric memory allocation failed on %s: %s" % ("this rank" if err is not None else "another rank", err))
        self.handle = symm_mem.rendezvous(self.buffer, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world:
            raise RuntimeError("symmetric memory returned %d buffer pointers for %d ranks" % (len(ptrs), self.world))
        off = self.buffer.data_ptr() - ptrs[self.rank]    # the tensor's offset inside the (symmetric) allocation
        self._ptrs = (ctypes.c_uint64 * self.world)(*[p + off for p in ptrs])
        self.epoch = 0
        torch.cuda.synchronize(dev)                       # every buffer is zero ...
        dist.barrier(group)                               # ... before any rank's first flag can arrive

    def allreduce_(self, values):
        import ctypes

        if values.dtype != torch.int64 or not values.is_cuda or not values.is_contiguous():
            raise TypeError("values must be a contiguous int64 CUDA tensor")
        n = values.numel()
        if n > self.capacity:
            raise ValueError("%d values exceed the exchange buffer's capacity %d" % (n, self.capacity))
        self.epoch += 1
        lib = self._lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream(values.device).cuda_stream)
        self._lib.check(lib.cube_peer_allreduce_i64(self.world, self.rank, self._ptrs, ctypes.c_void_p(values.data_ptr()), n,
                                                    self.capacity, self.epoch, stream), "cube_peer_allreduce_i64")
        return values


def reward_total(counters):
    """Exact sum of the +-1 rewards behind the counters: 2*solved - produced."""
    return 2 * int(counters[0]) - int(counters[1])


def max_over_ranks(value, device):
    """MAX all-reduce of a Python float (timing: the slowest rank defines the step)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu_node(device_index):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the pinned
    host buffers it allocates afterwards are first-touched on that node: with one process per GPU the
    host-buffer pipeline's copies (75 GB/s per GPU in both directions together) then stay off the
    inter-socket link.  Returns a dict describing what was done; never raises."""
    info = {"numa_node": None, "cpus": None}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        info["pci"] = bdf
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read()) & set(os.sched_getaffinity(0))
        if not cpus:
            return info
        os.sched_setaffinity(0, cpus)
        info["numa_node"], info["cpus"] = node, len(cpus)
    except Exception as exc:  # noqa: BLE001 - best effort: containers may hide sysfs or forbid affinity changes
        info["error"] = "%s: %s" % (type(exc).__name__, exc)
    return info
