"""Multi-GPU plumbing for the predict path: one process per GPU, clips sharded
across ranks with no data-path collective, ONE exchange step - an all-gather of the per-rank
score rows (NCCL over NVLink through the engine's ``nisqa_gather_nccl``; ``gloo`` on CPU for
the host-logic tests).  The reference's own multi-GPU mechanism is ``nn.DataParallel`` over
the batch (reference nisqa/NISQA_model.py:56-57); clips are independent, so sharding the file
list is result-equivalent.
"""
import os

import numpy as np


def env_world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process if absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun env if WORLD_SIZE > 1.  Returns
    (rank, world, local_rank)."""
    rank, world, local = env_world()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def usable_cpus():
    """CPUs this process may really use: the affinity mask cut down to the cgroup CPU quota (cgroup v2 ``cpu.max``).  The
    GPU boxes show 128 logical CPUs but grant 16: decode threads beyond the grant only fight each other
    (profiles/r02ac_bench_files.txt: 8 threads per batch 15 400 clips/s, 16 threads 10 500)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()[:2]
        if quota != "max" and int(quota) > 0:
            n = max(1, min(n, -(-int(quota) // int(period))))
    except Exception:
        pass
    return n


def shutdown(engine=None):
    """Orderly end of a multi-rank run: barrier, close the engine (its NCCL communicator), destroy the process group.
    A no-op for the parts that do not exist (single process, no engine)."""
    rank, world, _ = env_world()
    initialised = False
    if world > 1:
        import torch.distributed as dist
        initialised = dist.is_initialized()
        if initialised:
            dist.barrier()
    if engine is not None:
        engine.close()
    if initialised:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def shard_rows(weights, world):
    """Deterministic balanced partition of row indices over ranks.

    Longest-processing-time-first on ``weights`` (file sizes ~ audio seconds): rows sorted by
    decreasing weight, each handed to the currently lightest rank.  Every rank computes the
    same table, so no communication is needed.  Returns a list of ``world`` int64 arrays, each
    in increasing row order."""
    w = np.asarray(weights, dtype=np.float64)
    order = np.argsort(-w, kind="stable")
    load = np.zeros(world, dtype=np.float64)
    count = np.zeros(world, dtype=np.int64)
    owner = np.empty(len(w), dtype=np.int64)
    for i in order:
        r = int(np.lexsort((np.arange(world), count, load))[0])
        owner[i] = r
        load[r] += w[i]
        count[r] += 1
    return [np.flatnonzero(owner == r).astype(np.int64) for r in range(world)]


def scatter_rows(gathered, shards, n_total):
    """gathered: [world, max_rows, n_out]; shards: per-rank row indices -> [n_total, n_out]."""
    gathered = np.asarray(gathered)
    out = np.full((n_total, gathered.shape[2]), np.nan, dtype=np.float32)
    for r, idx in enumerate(shards):
        out[idx] = gathered[r, :len(idx)]
    return out


def all_gather_scores(local_scores, shards, n_total, engine=None):
    """The single exchange step.  ``local_scores`` [n_local, n_out] float32 rows of this rank in
    the order of ``shards[rank]``.  Returns the full [n_total, n_out] matrix on every rank."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    n_out = local_scores.shape[1]
    max_rows = max(1, max(len(s) for s in shards))
    if engine is not None and dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        if getattr(engine, "_nccl_ready", False) is False:
            uid = [engine.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            engine.nccl_init(world, rank, uid[0])
            engine._nccl_ready = True
        loc = torch.full((max_rows, n_out), float("nan"), dtype=torch.float32, device=dev)
        loc[:local_scores.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_scores)).to(dev)
        glob = torch.empty((world, max_rows, n_out), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        engine.gather_nccl(loc.data_ptr(), max_rows, glob.data_ptr())
        gathered = glob.cpu().numpy()
    else:
        loc = torch.full((max_rows, n_out), float("nan"), dtype=torch.float32)
        loc[:local_scores.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_scores))
        parts = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(parts, loc)
        gathered = torch.stack(parts).numpy()
    return scatter_rows(gathered, shards, n_total)


def bind_to_gpu_numa(device_index=0):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers
    are allocated next to the GPU's PCIe root (a remote node roughly halves H2D bandwidth).
    Best effort: returns the node number, or None when the topology is not visible."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None
