"""Multi-GPU sharding of the collision path; host-side logic only.

Two modes, both one process per GPU with ``torch.distributed`` as plumbing:

* **Independent worlds** (config C3, or one scene per rank): ``shard_worlds`` gives rank r the
  worlds ``[r*W/R, (r+1)*W/R)``.  No data-path collective exists or is needed.
* **One huge scene** (config C4): 1-D slab decomposition along x.  ``plan_slabs`` picks equal-count
  splitters from the AABB centres; every rank then needs all bodies whose x-interval meets its slab
  (its own plus *ghosts*), which ``exchange_ghosts`` moves with point-to-point sends (NCCL over
  NVLink on GPUs, gloo in the CPU tests).  De-duplication rule: a pair (i, j) is reported by the
  rank whose slab contains ``x* = max(min_i.x, min_j.x)`` — the left end of the pair's x-overlap.
  Both x-intervals contain x*, so both bodies are present on that rank, and slabs partition the
  axis, so exactly one rank reports each pair.

Everything here works on numpy arrays; the per-rank compute is whatever collision backend the
caller plugs in (the CUDA ``CollisionWorld`` on GPUs, the CPU oracle in the gloo tests).
"""
import numpy as np

from . import SHAPE_CONVEX, SHAPE_DT, Scene


# ---------------------------------------------------------------------------------------------
# independent worlds
# ---------------------------------------------------------------------------------------------
def world_range(num_worlds, rank, size):
    return rank * num_worlds // size, (rank + 1) * num_worlds // size


def _subset(scene, idx, world_id=None, num_worlds=1):
    """Scene restricted to bodies idx (hull vertex ranges are re-packed)."""
    shapes = scene.shapes[idx].copy()
    hull_parts, used = [], 0
    is_hull = shapes["type"] == SHAPE_CONVEX
    if is_hull.any():
        first = shapes["p0"].view(np.uint32)
        count = shapes["p1"].view(np.uint32)
        new_first = first.copy()
        for k in np.nonzero(is_hull)[0]:
            hull_parts.append(scene.hull[first[k]:first[k] + count[k]])
            new_first[k] = used
            used += int(count[k])
        shapes["p0"] = new_first.view(np.float32)
    hull = np.concatenate(hull_parts) if hull_parts else np.zeros((0, 3), np.float32)
    return Scene(scene.xf[idx], shapes, hull, world_id, num_worlds, scene.name)


def shard_worlds(scene, rank, size):
    """Returns (local_scene, global_index_of_local_body)."""
    lo, hi = world_range(scene.num_worlds, rank, size)
    idx = np.nonzero((scene.world_id >= lo) & (scene.world_id < hi))[0]
    local = _subset(scene, idx, scene.world_id[idx] - lo, max(1, hi - lo))
    return local, idx.astype(np.uint32)


# ---------------------------------------------------------------------------------------------
# slab decomposition of one scene
# ---------------------------------------------------------------------------------------------
def plan_slabs(center_x, size):
    """Equal-count splitters on the centre x coordinate: slab r = [edges[r], edges[r+1])."""
    cx = np.sort(np.asarray(center_x, dtype=np.float64))
    cx = cx[np.isfinite(cx)]
    edges = np.empty(size + 1, np.float64)
    edges[0], edges[-1] = -np.inf, np.inf
    for r in range(1, size):
        edges[r] = cx[min(len(cx) - 1, r * len(cx) // size)] if len(cx) else 0.0
    return edges


def owner_of(center_x, edges):
    """Rank owning each body: the slab that contains its AABB centre (NaN -> rank 0)."""
    cx = np.nan_to_num(np.asarray(center_x, np.float64), nan=-np.inf)
    return np.clip(np.searchsorted(edges, cx, side="right") - 1, 0, len(edges) - 2).astype(np.int32)


def slab_mask(aabb, edges, r):
    """Bodies whose closed x-interval [min.x, max.x] meets the half-open slab r."""
    return (aabb[:, 0] < edges[r + 1]) & (aabb[:, 3] >= edges[r])


def _pack(scene, idx, gid):
    """Flat float32 payload for bodies idx: per body 10 xf + 4 shape words + 1 id, then hull xyz."""
    sub = _subset(scene, idx)
    body = np.zeros((len(idx), 15), np.float32)
    body[:, :10] = sub.xf
    body[:, 10:14] = sub.shapes.view(np.float32).reshape(-1, 4)
    body[:, 14] = gid[idx].astype(np.uint32).view(np.float32)
    header = np.array([len(idx), len(sub.hull)], np.uint32).view(np.float32)
    return np.concatenate([header, body.ravel(), sub.hull.ravel()]).astype(np.float32)


def _unpack(buf):
    nb, nh = buf[:2].view(np.uint32)
    body = buf[2:2 + 15 * nb].reshape(nb, 15)
    hull = buf[2 + 15 * nb:2 + 15 * nb + 3 * nh].reshape(nh, 3)
    shapes = np.ascontiguousarray(body[:, 10:14]).view(SHAPE_DT).reshape(-1)
    return body[:, :10].copy(), shapes.copy(), hull.copy(), body[:, 14].copy().view(np.uint32)


def exchange_ghosts(owned, owned_gid, owned_aabb, edges, rank, size, dist=None, device="cpu"):
    """Sends every owned body to each other rank whose slab its x-interval meets; returns the
    local scene (owned + ghosts, ordered by global id) and the global id of each local body.

    ``dist`` is ``torch.distributed`` (initialised) or None for a single process."""
    parts = [(owned.xf, owned.shapes, owned.hull, owned_gid)]
    if size > 1:
        import torch
        payloads = {}
        for r in range(size):
            if r == rank:
                continue
            idx = np.nonzero(slab_mask(owned_aabb, edges, r))[0]
            payloads[r] = _pack(owned, idx, owned_gid)
        # sizes first (one all_gather), then pairwise non-blocking sends / receives
        my_sizes = torch.zeros(size, dtype=torch.int64, device=device)
        for r, p in payloads.items():
            my_sizes[r] = len(p)
        all_sizes = [torch.zeros(size, dtype=torch.int64, device=device) for _ in range(size)]
        dist.all_gather(all_sizes, my_sizes)
        recv = {r: torch.empty(int(all_sizes[r][rank]), dtype=torch.float32, device=device)
                for r in range(size) if r != rank}
        send = {r: torch.from_numpy(p).to(device) for r, p in payloads.items()}
        # one grouped exchange (ncclGroupStart/End under NCCL: un-grouped send-then-recv on both sides
        # of a pair would wait on each other); empty messages are skipped on both ends
        ops = []
        for r in range(size):
            if r == rank:
                continue
            if send[r].numel():
                ops.append(dist.P2POp(dist.isend, send[r], r))
            if recv[r].numel():
                ops.append(dist.P2POp(dist.irecv, recv[r], r))
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        for r in sorted(recv):
            parts.append(_unpack(recv[r].cpu().numpy()))
    xf = np.concatenate([p[0] for p in parts])
    shapes = np.concatenate([p[1] for p in parts])
    gid = np.concatenate([p[3] for p in parts]).astype(np.uint32)
    # re-base hull ranges of the appended parts
    hull_parts, used, off = [], 0, 0
    for p in parts:
        nb = len(p[0])
        is_hull = shapes["type"][off:off + nb] == SHAPE_CONVEX
        if is_hull.any():
            first = shapes["p0"][off:off + nb].view(np.uint32)
            first[is_hull] += used
        hull_parts.append(p[2])
        used += len(p[2])
        off += nb
    hull = np.concatenate(hull_parts) if hull_parts else np.zeros((0, 3), np.float32)
    # local index order = global id order, so every pair is evaluated with the same body in the
    # "A" role as in a single-process run (contact floats then match bit for bit)
    order = np.argsort(gid, kind="stable")
    return Scene(xf[order], shapes[order], hull, name=owned.name), gid[order]


def filter_pairs_for_rank(pairs_local, local_aabb, local_gid, edges, rank):
    """Keeps the local pairs this rank must report (x* rule) and returns them as canonical
    (a<b, sorted) GLOBAL id pairs plus the boolean mask over pairs_local."""
    if len(pairs_local) == 0:
        return np.zeros((0, 2), np.uint32), np.zeros(0, bool)
    i, j = pairs_local[:, 0], pairs_local[:, 1]
    xs = np.maximum(local_aabb[i, 0], local_aabb[j, 0]).astype(np.float64)
    keep = (xs >= edges[rank]) & (xs < edges[rank + 1])
    gi, gj = local_gid[i[keep]], local_gid[j[keep]]
    a, b = np.minimum(gi, gj), np.maximum(gi, gj)
    order = np.lexsort((b, a))
    return np.stack([a[order], b[order]], axis=1).astype(np.uint32), keep


class CudaBackend:
    """Per-rank collision backend on this rank's GPU (the product path: CollisionWorld over the C ABI)."""

    def __init__(self, device=0, pairs_per_body=8):
        self.device = device
        self.pairs_per_body = pairs_per_body

    def refit(self, scene):
        from . import CollisionWorld
        w = CollisionWorld.for_scene(scene, pairs_per_body=self.pairs_per_body, device=self.device)
        w.refit()
        bb = w.aabbs()
        w.close()
        return bb

    def step(self, scene):
        from . import CollisionWorld
        w = CollisionWorld.for_scene(scene, pairs_per_body=self.pairs_per_body, device=self.device)
        w.step()
        out = (w.aabbs(), w.pairs().copy(), w.contacts().copy())
        w.close()
        return out


class _DevBuf:
    """Zero-copy view of a device buffer owned by the C library (__cuda_array_interface__)."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2}


class SlabRank:
    """One rank of the x-slab decomposition on the CUDA path (fast variant of slab_step).

    The context keeps the owned bodies resident; per step only the owned AABBs come back to the host
    (to pick the ghosts), the ghosts travel rank-to-rank, and the ownership rule and the pair
    orientation by global id are applied inside the traversal kernel (axcd_set_slab /
    axcd_set_body_keys), so there is no host-side filtering and no pair is narrow-phased twice.
    Hull shapes cannot be ghosts in this version; use slab_step for scenes with hulls."""

    def __init__(self, owned, owned_gid, edges, rank, size, device=0, ghost_frac=0.5, pairs_per_body=8):
        from . import CollisionWorld
        self.owned, self.gid = owned, np.ascontiguousarray(owned_gid, np.uint32)
        self.edges, self.rank, self.size = edges, rank, size
        cap = int(owned.n * (1.0 + ghost_frac)) + 4096
        # hull ghosts bring their vertices (axcd_slab_step): room for them behind the owned ones
        hull_cap = int(len(owned.hull) * (1.0 + ghost_frac)) + (8192 if len(owned.hull) else 0)
        self.w = CollisionWorld(cap, max_pairs=pairs_per_body * cap, max_hull_verts=hull_cap, device=device)
        self.w.set_shapes(owned.shapes, owned.hull)
        self.w.set_transforms(owned.xf)
        self.w.set_body_keys(self.gid, 0)
        self.w.set_slab(np.float32(edges[rank]), np.float32(edges[rank + 1]))
        self.local_gid = self.gid
        self._bb = np.zeros((cap, 6), np.float32)

    def owned_aabbs(self):
        import ctypes as C
        self.w.refit()
        self.w._check(self.w._lib.axcd_get_aabbs(self.w._ctx, self._bb.ctypes.data_as(C.c_void_p), len(self._bb)),
                      "axcd_get_aabbs")
        return self._bb[:self.owned.n]

    def ghost_payloads(self, bb):
        """Flat float32 payload for every other rank (see _pack)."""
        return {r: _pack(self.owned, np.nonzero(slab_mask(bb, self.edges, r))[0], self.gid)
                for r in range(self.size) if r != self.rank}

    def exchange(self, payloads, dist, device):
        import torch
        my_sizes = torch.zeros(self.size, dtype=torch.int64, device=device)
        for r, p in payloads.items():
            my_sizes[r] = len(p)
        all_sizes = [torch.zeros(self.size, dtype=torch.int64, device=device) for _ in range(self.size)]
        dist.all_gather(all_sizes, my_sizes)
        recv = {r: torch.empty(int(all_sizes[r][self.rank]), dtype=torch.float32, device=device)
                for r in range(self.size) if r != self.rank}
        send = {r: torch.from_numpy(p).to(device) for r, p in payloads.items()}
        ops = []
        for r in range(self.size):
            if r == self.rank:
                continue
            ops.append(dist.P2POp(dist.isend, send[r], r))
            ops.append(dist.P2POp(dist.irecv, recv[r], r))
        for q in dist.batch_isend_irecv(ops):
            q.wait()
        return [recv[r].cpu().numpy() for r in sorted(recv)]

    def step_with(self, received):
        """received: list of payloads from the other ranks.  Returns AxcdStats."""
        parts = [_unpack(b) for b in received]
        if parts:
            gxf = np.concatenate([p[0] for p in parts])
            gsh = np.concatenate([p[1] for p in parts])
            ggid = np.concatenate([p[3] for p in parts]).astype(np.uint32)
        else:
            gxf, gsh, ggid = np.zeros((0, 10), np.float32), np.zeros(0, SHAPE_DT), np.zeros(0, np.uint32)
        self.w.set_ghosts(self.owned.n, gxf, gsh, ggid)
        self.local_gid = np.concatenate([self.gid, ggid])
        return self.w.step()

    def step(self, dist=None, device="cpu"):
        bb = self.owned_aabbs()
        received = self.exchange(self.ghost_payloads(bb), dist, device) if self.size > 1 else []
        return self.step_with(received)

    # ---- device-resident variant: no AABB download, no host-side selection ---------------------------
    def pack_device(self):
        """Refit + device-side ghost selection.  Returns {rank: float32 CUDA tensor view (records*16)}."""
        import torch
        self.w.refit()
        ptrs, counts = self.w.pack_ghosts(np.asarray(self.edges, np.float32), self.size, self.rank)
        dev = f"cuda:{self.w.cfg.deviceOrdinal}"
        out = {}
        for r in range(self.size):
            if r == self.rank:
                continue
            n = int(counts[r]) * 16
            out[r] = (torch.as_tensor(_DevBuf(ptrs[r], n), device=dev) if n else
                      torch.empty(0, dtype=torch.float32, device=dev))
        return out

    def exchange_device(self, send, dist):
        import torch
        dev = f"cuda:{self.w.cfg.deviceOrdinal}"
        my_sizes = torch.zeros(self.size, dtype=torch.int64, device=dev)
        for r, t in send.items():
            my_sizes[r] = t.numel()
        all_sizes = [torch.zeros(self.size, dtype=torch.int64, device=dev) for _ in range(self.size)]
        dist.all_gather(all_sizes, my_sizes)
        recv = {r: torch.empty(int(all_sizes[r][self.rank]), dtype=torch.float32, device=dev)
                for r in range(self.size) if r != self.rank}
        ops = []
        for r in range(self.size):
            if r == self.rank:
                continue
            if send[r].numel():
                ops.append(dist.P2POp(dist.isend, send[r], r))
            if recv[r].numel():
                ops.append(dist.P2POp(dist.irecv, recv[r], r))
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        return [recv[r] for r in sorted(recv)]

    def step_with_device(self, received):
        """received: list of CUDA float32 tensors of ghost records from the other ranks."""
        import torch
        torch.cuda.current_stream().synchronize()   # the records must be complete before our stream reads them
        if received:
            allrec = torch.cat(received) if len(received) > 1 else received[0]
        else:
            allrec = None
        ng = (allrec.numel() // 16) if allrec is not None else 0
        self.w.set_ghosts_device(self.owned.n, ng, allrec.data_ptr() if ng else 0)
        self._keep = allrec   # keep the buffer alive until the step has consumed it
        self.local_gid = None
        return self.w.step()

    def step_device(self, dist=None):
        send = self.pack_device()
        received = self.exchange_device(send, dist) if self.size > 1 else []
        return self.step_with_device(received)

    # ---- the exchange inside the library (axcd_slab_init / axcd_slab_step: NCCL, no torch on the data path) ----
    def init_native(self, dist=None):
        """Creates the library's own NCCL communicator.  The 128-byte unique id comes from rank 0 and
        travels through torch.distributed (plumbing only; a C++ host would use its own launcher)."""
        from . import nccl_unique_id
        if self.size > 1 and dist is not None:
            import torch
            dev = f"cuda:{self.w.cfg.deviceOrdinal}" if dist.get_backend() == "nccl" else "cpu"
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
            if self.rank == 0:
                t.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(t, 0)
            uid = bytes(t.cpu().numpy().tobytes())
        else:
            uid = nccl_unique_id()
        self.w.slab_init(uid, self.rank, self.size, np.asarray(self.edges, np.float32))

    def step_native(self):
        """One collective slab step entirely behind the C ABI.  Returns AxcdStats (with exchangeMs, ghostBodies)."""
        st = self.w.slab_step()
        self.local_gid = None
        return st

    def pairs_global(self):
        if self.local_gid is None:
            self.local_gid = self.w.body_keys()
        p = self.local_gid[self.w.pairs()]
        return p[np.lexsort((p[:, 1], p[:, 0]))]

    def contacts_global(self):
        if self.local_gid is None:
            self.local_gid = self.w.body_keys()
        c = self.w.contacts().copy()
        c["a"], c["b"] = self.local_gid[c["a"]], self.local_gid[c["b"]]
        return c[np.lexsort((c["b"], c["a"]))]

    def close(self):
        self.w.close()


def slab_step(scene_owned, owned_gid, edges, rank, size, backend, dist=None, device="cpu"):
    """One sharded broadphase+narrowphase step.

    backend(scene) -> (aabb (n,6), pairs (m,2) local canonical, contacts structured array) is the
    per-rank collision implementation.  Returns (global pairs this rank reports, their contacts
    with global ids)."""
    owned_aabb = backend.refit(scene_owned)
    local, gid = exchange_ghosts(scene_owned, owned_gid, owned_aabb, edges, rank, size, dist, device)
    aabb, pairs, contacts = backend.step(local)
    gp, keep = filter_pairs_for_rank(pairs, aabb, gid, edges, rank)
    # contacts follow the same rule, keyed by their pair (both lists are in canonical local order)
    if len(contacts):
        pk = pairs[:, 0].astype(np.uint64) << np.uint64(32) | pairs[:, 1].astype(np.uint64)
        ck = contacts["a"].astype(np.uint64) << np.uint64(32) | contacts["b"].astype(np.uint64)
        cm = keep[np.searchsorted(pk, ck)]
        c = contacts[cm].copy()
        ga, gb = gid[c["a"]], gid[c["b"]]
        swap = ga > gb
        c["a"], c["b"] = np.minimum(ga, gb), np.maximum(ga, gb)
        for f in ("nx", "ny", "nz"):   # normal points from a to b: flip when the ids swap order
            c[f] = np.where(swap, -c[f], c[f])
        c = c[np.lexsort((c["b"], c["a"]))]
    else:
        c = contacts
    return gp, c
