"""Multi-GPU pool scoring: one process per GPU, torch.distributed (NCCL over NVLink on the GPU
box, gloo in CPU tests) for the two small exchanges the path has:

  1. the fitted surrogate (PoolModel: dictionaries, G, w_mu, scalars; a few MB) is broadcast
     once from the rank that fitted it;
  2. every rank scores its own contiguous shard of the candidate pool — no data-path
     collective — and the per-rank top-n lists (n x (value, global index)) are all-gathered and
     merged with the reference's distinct-by-value rule (bayesopt/acquisitions.py:101-105).

Because each candidate is scored by exactly one rank from identical replicated state, scores
are bitwise independent of the number of ranks.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .pool_model import PoolModel


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise from torchrun's environment; returns (rank, world_size, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (sysfs cpulist) -> [0, 1, 2, 3, 8, 10, 11]."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host pools it allocates
    afterwards (first touch) and the threads that fill them are local to that GPU's PCIe root: with 8 ranks on one
    host the pools otherwise land wherever the scheduler started the process.  Best effort: returns the node, or
    None (and changes nothing) when the topology cannot be read or the node's CPUs are not available to us."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s" % (dom[-4:].lower(), rest.lower())
        node = int(open(path + "/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set(parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read()))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous index range [lo, hi) of `rank` ([r*M/R, (r+1)*M/R))."""
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def deal_row_blocks(n_rows: int, block: int, world: int, symmetric: bool = True) -> List[List[int]]:
    """Row blocks of a (symmetric) N x N Gram for every rank: starts[r] = first rows of rank r's blocks, ascending.
    With the symmetric schedule a block starting at row r0 only computes the tiles right of its diagonal —
    rows x (N - r0) of work — so equal blocks are unequal work: blocks are dealt largest first to the least loaded
    rank (ties to the lowest rank: every rank computes the same answer without talking).  Otherwise round-robin."""
    starts = list(range(0, n_rows, block))
    if world <= 1:
        return [starts]
    if not symmetric:
        return [[r0 for i, r0 in enumerate(starts) if i % world == r] for r in range(world)]
    load = [0.0] * world
    mine: List[List[int]] = [[] for _ in range(world)]
    for r0 in sorted(starts, key=lambda r: (-(min(block, n_rows - r) * (n_rows - r)), r)):
        r = min(range(world), key=lambda i: (load[i], i))
        mine[r].append(r0)
        load[r] += min(block, n_rows - r0) * (n_rows - r0)
    return [sorted(m) for m in mine]


def merge_topk(vals: np.ndarray, idx: np.ndarray, k: int, distinct: bool = True):
    """Merge candidate (value, global index) pairs into the global top-k: k largest DISTINCT
    values, each with the lowest index at which it occurs (np.unique(return_index=True) +
    argpartition in the reference); distinct=False gives plain top-k with ties by lowest index.
    Entries with idx < 0 or NaN values are ignored.  Sorted by descending value."""
    vals = np.asarray(vals, dtype=np.float32).reshape(-1)
    idx = np.asarray(idx, dtype=np.int64).reshape(-1)
    ok = (idx >= 0) & ~np.isnan(vals)
    vals, idx = vals[ok], idx[ok]
    order = np.lexsort((idx, -vals))                 # value desc, index asc
    vals, idx = vals[order], idx[order]
    if distinct and len(vals):
        keep = np.concatenate([[True], vals[1:] != vals[:-1]])
        vals, idx = vals[keep], idx[keep]
    return vals[:k], idx[:k]


def allgather_topk(vals: np.ndarray, idx: np.ndarray, k: int, distinct: bool = True, device=None):
    """All-gather every rank's local top-k list and merge; every rank returns the same result."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return merge_topk(vals, idx, k, distinct)
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    v = torch.full((k,), float("nan"), dtype=torch.float32)
    i = torch.full((k,), -1, dtype=torch.int64)
    v[:len(vals)] = torch.from_numpy(np.asarray(vals, dtype=np.float32))
    i[:len(idx)] = torch.from_numpy(np.asarray(idx, dtype=np.int64))
    v, i = v.to(device), i.to(device)
    gv = [torch.empty_like(v) for _ in range(world)]
    gi = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(gv, v)
    dist.all_gather(gi, i)
    return merge_topk(torch.cat(gv).cpu().numpy(), torch.cat(gi).cpu().numpy(), k, distinct)


def allgather_topk_device(eng, local_cands: torch.Tensor, k: int, distinct: bool = True,
                          gathered: Optional[torch.Tensor] = None):
    """Device-resident variant used on the GPU path: one all-gather of k 16-byte entries per rank
    (NCCL over NVLink), merged by the same tournament kernel; a single host sync per pool."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return eng.topk_merge(local_cands, k, distinct)
    world = dist.get_world_size()
    if gathered is None:
        gathered = torch.empty((world * local_cands.numel(),), dtype=torch.uint8, device=local_cands.device)
    dist.all_gather_into_tensor(gathered, local_cands)
    return eng.topk_merge(gathered, k, distinct)


def select_topk(eng, scores: torch.Tensor, k: int, distinct: bool = True, index_base: int = 0,
                cand_buf: Optional[torch.Tensor] = None, gathered: Optional[torch.Tensor] = None):
    """Global top-k of the ranks' score shards: (values, global indices) on the host, identical on
    every rank.  One rank: one tournament + readback (nbw_topk).  Several: local tournament left on the
    device, one all-gather of k entries per rank, merge (nbw_topk_device / nbw_topk_merge)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return eng.topk(scores, k, distinct=distinct, index_base=index_base)
    local = eng.topk_device(scores, k, distinct=distinct, index_base=index_base, out=cand_buf)
    return allgather_topk_device(eng, local, k, distinct, gathered=gathered)


def broadcast_pool_model(pm: Optional[PoolModel], src: int = 0, device=None) -> PoolModel:
    """Broadcast the fitted surrogate from `src` to every rank (NCCL: over NVLink) as ONE flat byte
    buffer: a 256-byte head (json length, payload length), the json meta and every tensor at an aligned
    offset — two collectives (head, buffer) instead of a pickled object list plus one per tensor.
    Receivers get views into the received buffer."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return pm
    rank = dist.get_rank()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    head = torch.zeros((2,), dtype=torch.int64, device=device)
    buf = None
    if rank == src:
        buf = pm.pack_flat()
        if buf.device != device:
            buf = buf.to(device)
        head.copy_(buf[:16].view(torch.int64))
    dist.broadcast(head, src=src)
    if rank == src:
        dist.broadcast(buf, src=src)
        return pm
    hlen, plen = [int(v) for v in head.cpu()]
    buf = torch.empty((256 + ((hlen + 255) & ~255) + plen,), dtype=torch.uint8, device=device)
    dist.broadcast(buf, src=src)
    header = bytes(buf[256:256 + hlen].cpu().numpy())
    return PoolModel.unpack_flat(buf, header)


def score_and_select(eng, model, cells: torch.Tensor, n_cells: int, k: int, distinct: bool = True, index_base: int = 0,
                     scores: Optional[torch.Tensor] = None, cand_buf: Optional[torch.Tensor] = None,
                     gathered: Optional[torch.Tensor] = None):
    """One pool step of a rank: score its shard and agree on the global top-k.  The local top-k is fused
    into the scoring kernel when the packed path serves the model (nbw_score_pool_topk); the per-rank
    lists are all-gathered (k x 16 B per rank) and merged on every rank.  Returns (values, global indices)."""
    if eng.fused_topk_ok(model, k):
        local = eng.score_pool_topk(model, cells, n_cells, k, distinct, index_base, scores=scores, out=cand_buf)
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return eng.cands_read(local, k)
        return allgather_topk_device(eng, local, k, distinct, gathered=gathered)
    sc, _, _, _ = eng.score_pool(model, cells, n_cells, scores=scores)
    return select_topk(eng, sc, k, distinct, index_base, cand_buf, gathered)


CAND_DTYPE = np.dtype([("v", "<f4"), ("pad", "<i4"), ("idx", "<i8")])


def parse_cands(raw: np.ndarray):
    """16-byte device list entries {f32 value; i32 pad; i64 index} -> (values, indices), empties dropped."""
    c = np.frombuffer(raw.tobytes(), dtype=CAND_DTYPE)
    keep = c["idx"] >= 0
    return c["v"][keep].astype(np.float32), c["idx"][keep].astype(np.int64)


class PoolPipeline:
    """Pool steps in flight: while pool i+1 is being scored on the main stream, the exchange of pool i — the
    all-gather of the ranks' k-entry lists (NCCL over NVLink), the device merge and the 16 k-byte copy to pinned
    host memory — runs on a side stream.  The per-pool collective thus leaves the critical path, and
    a single GPU no longer stalls on a read-back between pools.

        t = pipe.submit(model, cells, n, index_base)    # enqueue; returns a ticket
        vals, idx = pipe.collect(t)                      # blocks until that pool's selection is on the host
    """

    def __init__(self, eng, k: int, distinct: bool = True, depth: int = 2):
        self.eng, self.k, self.distinct, self.depth = eng, int(k), bool(distinct), int(depth)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.side = torch.cuda.Stream(device=eng.device)
        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(cand=eng.empty((k * 16,), torch.uint8),
                                   gathered=eng.empty((self.world * k * 16,), torch.uint8),
                                   merged=eng.empty((k * 16,), torch.uint8),
                                   host=torch.empty((k * 16,), dtype=torch.uint8).pin_memory(),
                                   scored=torch.cuda.Event(), done=torch.cuda.Event(), busy=False))
        self.head = 0

    def submit(self, model, cells: torch.Tensor, n_cells: int, index_base: int = 0,
               scores: Optional[torch.Tensor] = None) -> int:
        eng, k = self.eng, self.k
        ticket = self.head
        slot = self.slots[ticket % self.depth]
        if slot["busy"]:
            raise RuntimeError("PoolPipeline: collect() ticket %d before submitting more pools" % (ticket - self.depth))
        self.head += 1
        if eng.fused_topk_ok(model, k):
            eng.score_pool_topk(model, cells, n_cells, k, self.distinct, index_base, scores=scores, out=slot["cand"])
        else:
            sc, _, _, _ = eng.score_pool(model, cells, n_cells, scores=scores)
            eng.topk_device(sc, k, distinct=self.distinct, index_base=index_base, out=slot["cand"])
        main = torch.cuda.current_stream(eng.device)
        slot["scored"].record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(slot["scored"])
            src = slot["cand"]
            if self.world > 1:
                dist.all_gather_into_tensor(slot["gathered"], slot["cand"])
                src = eng.topk_merge_device(slot["gathered"], k, self.distinct, out=slot["merged"])
            slot["host"].copy_(src, non_blocking=True)
            slot["done"].record(self.side)
        slot["busy"] = True
        return ticket

    def collect(self, ticket: int):
        slot = self.slots[ticket % self.depth]
        slot["done"].synchronize()
        slot["busy"] = False
        return parse_cands(slot["host"].numpy())


def propose_location_sharded(acq, pool_shard, index_base: int, top_n: int = 5, return_distinct: bool = True):
    """Score this rank's shard (device records) and agree on the global top-n.
    Returns (top values, global indices, local scores device tensor)."""
    acq.iters += 1
    cells, M = acq.gp.upload_pool(pool_shard)
    eng = acq.gp._dev['eng']
    scores = eng.empty((M,), torch.float32)
    vals, idx = score_and_select(eng, acq._model(), cells, M, top_n, bool(return_distinct), index_base, scores=scores)
    return vals, idx, scores
