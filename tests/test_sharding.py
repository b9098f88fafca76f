"""Multi-process sharding logic on CPU: world_size-2 `gloo`, the per-rank
collision backend is the CPU oracle.  The union of what the ranks report must equal the
single-process answer exactly, with no duplicates."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import axcd
from axcd import sharding
import oracle_lib as O


class OracleBackend:
    def refit(self, s):
        rc, bb = O.refit(s.xf, s.shapes, s.hull)
        assert rc == 0
        return bb

    def step(self, s):
        bb = self.refit(s)
        pairs = O.broadphase(bb, s.world_id)
        con, _, _ = O.narrowphase(s.xf, s.shapes, pairs, s.hull)
        return bb, pairs, con


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _slab_worker(rank, size, port, scene_name, scale, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        s = axcd.config_scene(scene_name, scale=scale)
        be = OracleBackend()
        bb = be.refit(s)
        cx = (bb[:, 0].astype(np.float64) + bb[:, 3]) * 0.5
        edges = sharding.plan_slabs(cx, size)
        mine = np.nonzero(sharding.owner_of(cx, edges) == rank)[0]
        owned = sharding._subset(s, mine)
        gp, gc = sharding.slab_step(owned, mine.astype(np.uint32), edges, rank, size, be, dist)
        gathered = [None] * size
        dist.all_gather_object(gathered, (gp, gc, len(mine)))
        if rank == 0:
            out.put(gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scene_name,scale", [("C1", 0.03), ("C2", 0.003)])
def test_slab_decomposition_two_ranks_matches_single_process(scene_name, scale):
    size = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, size, port, scene_name, scale, q)) for r in range(size)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = axcd.config_scene(scene_name, scale=scale)
    _, ref_pairs, ref_con = OracleBackend().step(s)
    got = np.concatenate([g[0] for g in gathered])
    key = got[:, 0].astype(np.uint64) << np.uint64(32) | got[:, 1]
    assert len(np.unique(key)) == len(key), "a pair was reported by two ranks"
    order = np.argsort(key)
    assert np.array_equal(got[order], ref_pairs)
    con = np.concatenate([g[1] for g in gathered])
    con = con[np.lexsort((con["b"], con["a"]))]
    assert np.array_equal(con["a"], ref_con["a"]) and np.array_equal(con["b"], ref_con["b"])
    # local index order follows global ids, so each contact is computed exactly as in one process
    for f in ("px", "py", "pz", "nx", "ny", "nz", "depth"):
        assert np.array_equal(con[f].view(np.uint32), ref_con[f].view(np.uint32)), f
    assert sum(g[2] for g in gathered) == s.n and min(g[2] for g in gathered) > 0.4 * s.n


def _worlds_worker(rank, size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        s = axcd.config_scene("C3", scale=16 / 4096)
        local, gidx = sharding.shard_worlds(s, rank, size)
        _, pairs, con = OracleBackend().step(local)
        gp = gidx[pairs]
        t = torch.tensor([len(pairs), len(con)], dtype=torch.int64)
        dist.all_reduce(t)                      # the only collective: counters for reporting
        gathered = [None] * size
        dist.all_gather_object(gathered, gp)
        if rank == 0:
            out.put((gathered, t.tolist()))
    finally:
        dist.destroy_process_group()


def test_world_sharding_two_ranks_no_communication_on_data_path():
    size = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worlds_worker, args=(r, size, port, q)) for r in range(size)]
    for p in procs:
        p.start()
    gathered, totals = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = axcd.config_scene("C3", scale=16 / 4096)
    _, ref_pairs, ref_con = OracleBackend().step(s)
    got = np.concatenate(gathered)
    assert np.array_equal(got, ref_pairs)       # rank order == world order == canonical order
    assert totals == [len(ref_pairs), len(ref_con)]


def test_world_range_covers_everything_once():
    for w in (1, 7, 4096):
        for size in (1, 2, 3, 8):
            spans = [sharding.world_range(w, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == w
            assert all(spans[k][1] == spans[k + 1][0] for k in range(size - 1))


def test_xstar_rule_partitions_pairs_for_any_rank_count():
    s = axcd.config_scene("C1", scale=0.02)
    be = OracleBackend()
    bb, ref_pairs, _ = be.step(s)
    cx = (bb[:, 0].astype(np.float64) + bb[:, 3]) * 0.5
    for size in (3, 8):
        edges = sharding.plan_slabs(cx, size)
        total = []
        for r in range(size):
            m = np.nonzero(sharding.slab_mask(bb, edges, r))[0]        # owned + ghosts of rank r
            sub = sharding._subset(s, m)
            lbb, lp, _ = be.step(sub)
            gp, _ = sharding.filter_pairs_for_rank(lp, lbb, m.astype(np.uint32), edges, r)
            total.append(gp)
        got = np.concatenate(total)
        key = got[:, 0].astype(np.uint64) << np.uint64(32) | got[:, 1]
        assert len(np.unique(key)) == len(key)
        assert np.array_equal(got[np.argsort(key)], ref_pairs)
