"""CPU: host logic of the N>1 path (slab partition + halo bookkeeping), with world_size-2 gloo.

Each rank builds its slab descriptor, the ranks cross-check them, and a numpy emulation of the exchange
pattern the CUDA path uses (narrow halo of the search direction before every operator application, sum
all-reduce of the dot products) reproduces a global CG solve on a lattice stencil matrix."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_slabs_tile_the_lattice():
    part = importlib.import_module("lpm-c_b200.partition")
    for nz, world in [(216, 8), (216, 4), (216, 2), (100, 8), (21, 2), (33, 4)]:
        slabs = [part.make_slab(nz, 7, r, world) for r in range(world)]
        assert slabs[0].z0 == 0 and slabs[-1].z1 == nz
        for a, b in zip(slabs[:-1], slabs[1:]):
            assert a.z1 == b.z0
            assert a.g_hi == b.send_wide_lo and b.g_lo == a.send_wide_hi          # wide exchange counts agree
            assert a.narrow_hi == b.send_narrow_lo and b.narrow_lo == a.send_narrow_hi
        assert slabs[0].g_lo == 0 and slabs[-1].g_hi == 0
        assert sum(s.own1 - s.own0 for s in slabs) == nz * 7
        assert max(s.z1 - s.z0 for s in slabs) - min(s.z1 - s.z0 for s in slabs) <= 1
    with pytest.raises(ValueError):
        part.make_slab(12, 7, 0, 8)     # fewer than 4 owned layers per rank


def _worker(rank, world, port, q, ragged=False):
    import torch.distributed as dist
    import scipy.sparse as sp
    sys.path.insert(0, str(ROOT))
    lpm_lat = importlib.import_module("lpm-c_b200.lattice")
    part = importlib.import_module("lpm-c_b200.partition")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    n = 10
    lat = lpm_lat.sc_block(n)
    N = n ** 3
    conn = lat["conn"]
    rows = np.repeat(np.arange(N), conn.shape[1])[conn.ravel() >= 0]
    cols = conn.ravel()[conn.ravel() >= 0]
    vals = np.where(rows == cols, 70.0, -1.0 - ((rows + cols) % 5) / 5.0)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(N, N))
    if ragged:
        # carve a notch out of layers 3..5 (order preserved, as the reference's carving helpers do): layers of 100 / 60 particles
        x = lat["xyz"]
        keep = ~((x[:, 2] > 1.4) & (x[:, 2] < 2.6) & (x[:, 0] > 2.9))
        A = A[keep][:, keep].tocsr()
        N = int(keep.sum())
        counts = part.layer_counts(x[keep, 2], lat["h"])
        weights = np.bincount(np.rint(x[keep, 2] / lat["h"]).astype(int), weights=np.diff(A.indptr)).tolist()   # sum of nb_conn per layer
        s = part.make_ragged_slab(counts, rank, world, weights)
        n_lo_r, n_hi_r, n_lo_s, n_hi_s = s.narrow_recv_lo, s.narrow_recv_hi, s.narrow_send_lo, s.narrow_send_hi
    else:
        s = part.make_slab(n, n * n, rank, world)
        L = s.layer_size
        n_lo_r, n_hi_r, n_lo_s, n_hi_s = s.narrow_lo * L, s.narrow_hi * L, s.send_narrow_lo * L, s.send_narrow_hi * L
    b = np.sin(1e-2 * np.arange(N))
    g0 = s.first_global
    loc = slice(g0, g0 + s.n_local)
    Al = A[loc, loc].tocsr()                       # local operator incl. ghost columns
    own = np.zeros(s.n_local, dtype=bool)
    own[s.own0:s.own1] = True

    def exchange(v):
        # particle counts, exactly what lpmb_dist_set_slab / lpmb_dist_exchange use (csrc/lpmb_dist.cu)
        reqs = []
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(v[s.own0:s.own0 + n_lo_s].copy()), rank - 1))
            rl = torch.empty(n_lo_r, dtype=torch.float64)
            reqs.append(dist.irecv(rl, rank - 1))
        if rank < world - 1:
            reqs.append(dist.isend(torch.from_numpy(v[s.own1 - n_hi_s:s.own1].copy()), rank + 1))
            rh = torch.empty(n_hi_r, dtype=torch.float64)
            reqs.append(dist.irecv(rh, rank + 1))
        for r in reqs:
            r.wait()
        if rank > 0:
            v[s.own0 - n_lo_r:s.own0] = rl.numpy()
        if rank < world - 1:
            v[s.own1:s.own1 + n_hi_r] = rh.numpy()

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    bl = b[loc] * own
    x = np.zeros(s.n_local)
    r = bl.copy()
    p = r.copy()
    rr = allsum(r @ r)
    thresh = 1e-8 * rr + 1e-12
    it = 0
    while rr > thresh and it < 500:
        exchange(p)
        ap = (Al @ p) * own
        alpha = rr / allsum(p @ ap)
        x += alpha * p
        r -= alpha * ap
        rr_new = allsum(r @ r)
        it += 1
        if rr_new <= thresh:
            break
        p = r + (rr_new / rr) * p
        rr = rr_new
    # the narrow ghosts of x followed the iteration exactly (x = sum alpha_k p_k elementwise)
    xg = np.zeros(N)
    xg[g0 + s.own0:g0 + s.own1] = x[s.own0:s.own1]
    t = torch.from_numpy(xg)
    dist.all_reduce(t)
    res = np.linalg.norm(b - A @ t.numpy()) / np.linalg.norm(b)
    narrow_ok = True
    if rank > 0:
        lo = slice(s.own0 - n_lo_r, s.own0)
        narrow_ok &= np.array_equal(x[lo], t.numpy()[g0 + lo.start:g0 + lo.stop])
    if rank == 0:
        q.put((it, res, narrow_ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True], ids=["full_layers", "carved"])
def test_distributed_cg_emulation_with_gloo(ragged):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 500) + (500 if ragged else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, ragged)) for r in range(2)]
    for p in procs:
        p.start()
    it, res, narrow_ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert 0 < it < 200 and res <= 1.01e-4 and narrow_ok


def test_ragged_slabs_for_carved_specimens():
    """irregular specimens are cut at layer boundaries, balanced on the layer weights (sum of nb_conn);
    the counts handed to lpmb_dist_set_slab must agree pairwise between neighbouring ranks"""
    part = importlib.import_module("lpm-c_b200.partition")
    # uniform layers that divide evenly: identical to make_slab
    for nz, world in [(216, 8), (32, 4), (20, 2), (9, 1)]:
        for r in range(world):
            a, b = part.make_slab(nz, 7, r, world), part.make_ragged_slab([7] * nz, r, world)
            assert a.set_slab_args() == b.set_slab_args() and a.n_local == b.n_local and a.first_global == b.first_global
    # a notched plate: thin layers in the middle, weights != counts
    rng = np.random.default_rng(20240607)
    counts = [int(c) for c in np.r_[np.full(20, 400), np.full(9, 150), np.full(25, 400)] + rng.integers(0, 30, 54)]
    weights = [c * (61 if 2 <= z < 52 else 40) for z, c in enumerate(counts)]
    for world in (2, 3, 4, 8):
        slabs = [part.make_ragged_slab(counts, r, world, weights) for r in range(world)]
        assert slabs[0].z0 == 0 and slabs[-1].z1 == len(counts) and slabs[0].first_global == 0
        assert sum(s.own1 - s.own0 for s in slabs) == sum(counts)
        for s in slabs:
            assert s.z1 - s.z0 >= 4
            assert s.n_local == sum(counts[s.z0 - s.g_lo:s.z1 + s.g_hi]) and s.first_global == sum(counts[:s.z0 - s.g_lo])
            assert s.own1 - s.own0 == sum(counts[s.z0:s.z1])
        for a, b in zip(slabs[:-1], slabs[1:]):
            assert a.z1 == b.z0
            assert a.narrow_recv_hi == b.narrow_send_lo and b.narrow_recv_lo == a.narrow_send_hi
            assert a.n_local - a.own1 == b.wide_send_lo and b.own0 == a.wide_send_hi
            # global index ranges line up: my upper ghosts are the first particles my upper neighbour owns
            assert a.first_global + a.own1 == b.first_global + b.own0
        # optimal bottleneck: no contiguous partition with >= 4 layers per rank does better (brute force for small worlds)
        if world <= 3:
            import itertools
            nz, best = len(counts), float("inf")
            for cuts in itertools.combinations(range(4, nz - 3), world - 1):
                c = (0,) + cuts + (nz,)
                if min(b - a for a, b in zip(c[:-1], c[1:])) >= 4:
                    best = min(best, max(sum(weights[a:b]) for a, b in zip(c[:-1], c[1:])))
            assert max(s.weight for s in slabs) == best
        # better than the equal-layer split whenever the layers differ
        equal = [part.owned_layers(len(counts), r, world) for r in range(world)]
        assert max(s.weight for s in slabs) <= max(sum(weights[a:b]) for a, b in equal)
    with pytest.raises(ValueError):
        part.make_ragged_slab([5] * 7, 0, 2)          # fewer than 4 layers per rank
    # layer detection from z-slowest coordinates
    z = np.repeat(0.5 * np.arange(6), [4, 4, 2, 2, 4, 4]) - 1.25
    assert part.layer_counts(z, 0.5) == [4, 4, 2, 2, 4, 4]
    with pytest.raises(ValueError):
        part.layer_counts(z[::-1], 0.5)               # not z-slowest
    with pytest.raises(ValueError):
        part.layer_counts(np.r_[z[:8], z[12:]], 0.5)  # a missing layer


def test_ragged_slab_invariants_on_random_specimens():
    """property test (hypothesis): for any layer counts / weights and any world that fits, the slabs tile the lattice, every
    rank owns >= 4 layers, neighbouring ranks agree on every exchanged count, and the bottleneck is never worse than the
    equal-layer split"""
    from hypothesis import given, settings, strategies as st
    part = importlib.import_module("lpm-c_b200.partition")

    @settings(max_examples=120, deadline=None)
    @given(st.integers(1, 6).flatmap(lambda world: st.tuples(
        st.just(world),
        st.lists(st.tuples(st.integers(1, 500), st.integers(1, 70)), min_size=4 * world if world > 1 else 1, max_size=4 * world + 30))))
    def check(arg):
        world, layers = arg
        counts = [c for c, _ in layers]
        weights = [c * w for c, w in layers]
        nz = len(counts)
        slabs = [part.make_ragged_slab(counts, r, world, weights) for r in range(world)]
        assert slabs[0].z0 == 0 and slabs[-1].z1 == nz and slabs[0].own0 == 0 and slabs[-1].n_local == slabs[-1].own1
        assert sum(s.own1 - s.own0 for s in slabs) == sum(counts)
        for s in slabs:
            assert world == 1 or s.z1 - s.z0 >= 4
            a = s.set_slab_args()
            assert 0 <= a[0] < a[1] <= s.n_local                      # what lpmb_dist_set_slab checks (csrc/lpmb_dist.cu)
            assert a[2] <= a[0] and a[3] <= s.n_local - a[1]
            assert max(a[4:]) <= a[1] - a[0]
        for lo, hi in zip(slabs[:-1], slabs[1:]):
            assert lo.z1 == hi.z0
            assert lo.narrow_recv_hi == hi.narrow_send_lo and hi.narrow_recv_lo == lo.narrow_send_hi
            assert lo.n_local - lo.own1 == hi.wide_send_lo and hi.own0 == lo.wide_send_hi
            assert lo.first_global + lo.own1 == hi.first_global + hi.own0
        if world > 1:
            equal = [part.owned_layers(nz, r, world) for r in range(world)]
            assert max(s.weight for s in slabs) <= max(sum(weights[a:b]) for a, b in equal) + 1e-9

    check()


@pytest.mark.parametrize("world,n,ghost", [(2, 24, 4), (8, 216, 4), (4, 216, 4), (3, 100, 4), (8, 64, 4), (5, 47, 4)])
def test_multigrid_levels_on_slabs(world, n, ghost):
    """lpmb_mg_slab_plan (the level geometry of the fast mode's DISTRIBUTED multigrid hierarchy, lpmb_mg.cu): for every rank of
    a z-slab decomposition -- coarse site I sits on fine site 2 I in global coordinates; the ranks' owned layers tile every
    level; a distributed level gives every rank >= 2 owned layers and 2 ghost layers either side (clipped at the faces), its
    block holds every fine layer the restriction of its owned coarse layers reads and every coarse layer the prolongation of
    its owned fine layers reads; all ranks agree on the level count, on which level is the first replicated one and on the
    owned ranges gathered there; the coarsest level is whole on every rank."""
    import importlib
    capi = importlib.import_module("lpm-c_b200.capi")
    partition = importlib.import_module("lpm-c_b200.partition")
    slabs = [partition.make_slab(n, n * n, r, world, ghost_layers=ghost) for r in range(world)]
    owned = [s.z1 - s.z0 for s in slabs]
    plans = []
    for r, s in enumerate(slabs):
        nz_local = s.n_local // (n * n)
        plans.append(capi.mg_slab_plan(world, r, owned, n, n, nz_local, s.own0 // (n * n)))
    nlev, lrep = len(plans[0]["levels"]), plans[0]["lrep"]
    assert all(len(p["levels"]) == nlev and p["lrep"] == lrep and p["gat_off"] == plans[0]["gat_off"] and p["gat_cnt"] == plans[0]["gat_cnt"]
               for p in plans)
    assert 1 <= lrep < nlev and not plans[0]["levels"][-1]["dist"]
    for l in range(nlev):
        L = [p["levels"][l] for p in plans]
        nzg = L[0]["nzg"]
        assert all(x["nzg"] == nzg and x["nx"] == L[0]["nx"] and x["dist"] == L[0]["dist"] for x in L)
        if l > 0:
            assert nzg == (plans[0]["levels"][l - 1]["nzg"] + 1) // 2 and L[0]["nx"] == (plans[0]["levels"][l - 1]["nx"] + 1) // 2
        assert (l < lrep) == bool(L[0]["dist"])
        if not L[0]["dist"]:
            assert all(x["gz0"] == 0 and x["nz"] == nzg for x in L)
        # the owned global ranges tile the level, in rank order
        og = [(x["gz0"] + x["oz0"], x["gz0"] + x["oz1"]) for x in L] if (L[0]["dist"] or l == lrep) else None
        if og:
            assert og[0][0] == 0 and og[-1][1] == nzg and all(og[r][1] == og[r + 1][0] for r in range(world - 1))
        if l == lrep:
            lay = L[0]["nx"] * L[0]["ny"]
            assert plans[0]["gat_off"] == [a * lay for a, _ in og] and plans[0]["gat_cnt"] == [(b - a) * lay for a, b in og]
        if L[0]["dist"]:
            for r, x in enumerate(L):
                a, b = og[r]
                assert b - a >= 2
                assert x["gz0"] <= max(0, a - 2) and x["gz0"] + x["nz"] >= min(nzg, b + 2)      # two ghost layers, clipped
                if l > 0:   # coarse layer Z is owned by the owner of fine layer 2 Z
                    fa, fb = plans[r]["levels"][l - 1]["gz0"] + plans[r]["levels"][l - 1]["oz0"], plans[r]["levels"][l - 1]["gz0"] + plans[r]["levels"][l - 1]["oz1"]
                    assert a == (fa + 1) // 2 and b == (fb + 1) // 2
        if l > 0:
            nzf = plans[0]["levels"][l - 1]["nzg"]
            for r in range(world):
                F, Cc = plans[r]["levels"][l - 1], plans[r]["levels"][l]
                if not F["dist"]:
                    continue
                fa, fb = F["gz0"] + F["oz0"], F["gz0"] + F["oz1"]
                ca, cb = (fa + 1) // 2, (fb + 1) // 2
                # restriction of the owned coarse layers reads fine layers 2 Z - 1 .. 2 Z + 1: inside the fine block
                assert F["gz0"] <= max(0, 2 * ca - 1) and F["gz0"] + F["nz"] >= min(nzf, 2 * (cb - 1) + 2)
                # prolongation to the owned fine layers reads coarse layers z >> 1 and (z >> 1) + 1: inside the coarse block
                lo, hi = fa >> 1, min(Cc["nzg"] - 1, ((fb - 1) >> 1) + ((fb - 1) & 1))
                assert Cc["gz0"] <= lo and Cc["gz0"] + Cc["nz"] > hi


def test_distributed_vcycle_schedule_reproduces_the_global_vcycle():
    """The communication schedule of the distributed multigrid V-cycle (lpmb_mg.cu::mg_vcycle / mg_smooth) on a 1-D column
    of the lattice, emulated rank by rank in numpy on the level geometry lpmb_mg_slab_plan returns: two ghost layers of the
    iterate refreshed before every stencil pass EXCEPT the first post-smoothing sweep, the residual's before the restriction,
    the coarse correction's before the prolongation, the restricted residual all-gathered at the first replicated level,
    boundary classes and interpolation weights in GLOBAL coordinates.  The owned parts of the result must equal the V-cycle
    of one rank on the whole column -- for several decompositions, with a Dirichlet end (mask) and free ends.  Ghost and
    not-owned entries are poisoned before every use that the schedule does not cover, so a missing exchange shows."""
    import importlib
    capi = importlib.import_module("lpm-c_b200.capi")
    partition = importlib.import_module("lpm-c_b200.partition")
    S = {-2: 0.3, -1: 1.0, 1: 1.0, 2: 0.3}          # 2-hop stencil like the lattice tangent's z-reach
    NU, OM = 2, (0.56, 1.39)

    def w1(f, X, nc):                                # mg_w1
        d = f - 2 * X
        return 1.0 if d == 0 else (0.5 if X + 1 < nc else 1.0) if d == 1 else 0.5 if d == -1 else 0.0

    def apply_level(u, f, mask, gz0, nzg, scale, mode, om):
        """mg_stencil_kernel on a local block: neighbours outside the BLOCK are skipped, the diagonal is the global class's"""
        n = len(f)
        out = np.empty(n)
        for i in range(n):
            g = i + gz0
            a = 0.0
            if mode != 2:
                for o, s in S.items():
                    if 0 <= i + o < n:
                        a += s * (u[i + o] - u[i])
            diag = -scale * sum(s for o, s in S.items() if 0 <= g + o < nzg)
            r = f[i] - scale * a
            if mode == 1:
                out[i] = mask[i] * r
            else:
                out[i] = (0.0 if mode == 2 else u[i]) + om * mask[i] * (mask[i] * r) / diag
        return out

    def vcycle(world, owned, nz_local0, ghost_lo0, f0_global, mask0_global):
        plans = [capi.mg_slab_plan(world, r, owned, 64, 64, nz_local0[r], ghost_lo0[r]) for r in range(world)]
        nlev, lrep = len(plans[0]["levels"]), plans[0]["lrep"]
        lev = lambda r, l: plans[r]["levels"][l]
        POISON = 1e30

        def own(r, l):
            L = lev(r, l)
            return L["gz0"] + L["oz0"], L["gz0"] + L["oz1"]

        def exchange(vs, l):                         # mg_exchange: owners' values into the neighbours' two ghost layers
            if not lev(0, l)["dist"]:
                return
            for r in range(world):
                a, b = own(r, l)
                for q in (r - 1, r + 1):
                    if 0 <= q < world:
                        for g in (range(a, a + 2) if q == r - 1 else range(b - 2, b)):
                            vs[q][g - lev(q, l)["gz0"]] = vs[r][g - lev(r, l)["gz0"]]

        def poison_unowned(vs, l, keep_ghosts=0, before_gather=False):
            for r in range(world):
                # on a replicated level every rank computes every site itself; only what is about to be all-gathered (the
                # restricted residual and the mask of the first replicated level) is valid on the owned layers alone
                if world == 1 or not (lev(r, l)["dist"] or (before_gather and l == lrep)):
                    continue
                a, b = own(r, l)
                for i in range(len(vs[r])):
                    g = i + lev(r, l)["gz0"]
                    if not (a - keep_ghosts <= g < b + keep_ghosts):
                        vs[r][i] = POISON

        def sync_coarse(vs, l):                      # mg_sync_coarse
            if world == 1:
                return
            if lev(0, l)["dist"]:
                exchange(vs, l)
            elif l == lrep:
                for r in range(world):
                    a, b = own(r, l)
                    for q in range(world):
                        vs[q][a:b] = vs[r][a:b]

        # masks: level 0 = the boundary conditions on every local site; coarser = min over the interpolation support
        masks = [[mask0_global[lev(r, 0)["gz0"]:lev(r, 0)["gz0"] + lev(r, 0)["nz"]].copy() for r in range(world)]]
        for l in range(1, nlev):
            ms = []
            for r in range(world):
                F, Cc = lev(r, l - 1), lev(r, l)
                m = np.ones(Cc["nz"])
                for Z in range(Cc["nz"]):
                    Zg = Z + Cc["gz0"]
                    for dz in (-1, 0, 1):
                        fz = 2 * Zg + dz - F["gz0"]
                        if 0 <= fz < F["nz"] and w1(fz + F["gz0"], Zg, Cc["nzg"]) != 0.0:
                            m[Z] = min(m[Z], masks[l - 1][r][fz])
                ms.append(m)
            poison_unowned(ms, l, before_gather=True)
            sync_coarse(ms, l)
            masks.append(ms)

        U = [[None] * world for _ in range(nlev)]
        Fv = [[None] * world for _ in range(nlev)]
        # level-0 right-hand side: the PCG residual, ZERO on the ghost rows (the solve's mask)
        for r in range(world):
            L = lev(r, 0)
            f = f0_global[L["gz0"]:L["gz0"] + L["nz"]].copy()
            a, b = own(r, 0)
            for i in range(L["nz"]):
                if world > 1 and not (a <= i + L["gz0"] < b):
                    f[i] = 0.0
            Fv[0][r] = f

        def smooth(l, first, ghosts_valid=False):
            for s in range(NU):
                if first and s == 0:
                    for r in range(world):
                        L = lev(r, l)
                        U[l][r] = apply_level(None, Fv[l][r], masks[l][r], L["gz0"], L["nzg"], 2.0 ** l, 2, OM[s & 1])
                else:
                    if not (ghosts_valid and s == 0):
                        poison_unowned(U[l], l)
                        exchange(U[l], l)
                    else:
                        poison_unowned(U[l], l, keep_ghosts=2)
                    for r in range(world):
                        L = lev(r, l)
                        U[l][r] = apply_level(U[l][r], Fv[l][r], masks[l][r], L["gz0"], L["nzg"], 2.0 ** l, 0, OM[s & 1])

        def cycle(l):
            if l == nlev - 1:
                for r in range(world):
                    L = lev(r, l)
                    u = None
                    for s in range(6):
                        u = apply_level(u, Fv[l][r], masks[l][r], 0, L["nzg"], 2.0 ** l, 2 if s == 0 else 0, OM[0])
                    U[l][r] = u
                return
            smooth(l, True)
            poison_unowned(U[l], l)
            exchange(U[l], l)
            res = [apply_level(U[l][r], Fv[l][r], masks[l][r], lev(r, l)["gz0"], lev(r, l)["nzg"], 2.0 ** l, 1, 0.0) for r in range(world)]
            poison_unowned(res, l)
            exchange(res, l)
            for r in range(world):
                F, Cc = lev(r, l), lev(r, l + 1)
                fc = np.zeros(Cc["nz"])
                for Z in range(Cc["nz"]):
                    Zg = Z + Cc["gz0"]
                    for dz in (-1, 0, 1):
                        fz = 2 * Zg + dz - F["gz0"]
                        if 0 <= fz < F["nz"]:
                            fc[Z] += w1(fz + F["gz0"], Zg, Cc["nzg"]) * res[r][fz]
                    fc[Z] *= masks[l + 1][r][Z]
                Fv[l + 1][r] = fc
            if world > 1 and not lev(0, l + 1)["dist"] and l + 1 == lrep:
                poison_unowned(Fv[l + 1], l + 1, before_gather=True)
                sync_coarse(Fv[l + 1], l + 1)
            cycle(l + 1)
            poison_unowned(U[l + 1], l + 1)
            exchange(U[l + 1], l + 1)
            for r in range(world):
                F, Cc = lev(r, l), lev(r, l + 1)
                for fz in range(F["nz"]):
                    fzg = fz + F["gz0"]
                    s = 0.0
                    for az in range((fzg & 1) + 1):
                        Zg = (fzg >> 1) + az
                        Z = Zg - Cc["gz0"]
                        if Zg < Cc["nzg"] and 0 <= Z < Cc["nz"]:
                            s += w1(fzg, Zg, Cc["nzg"]) * U[l + 1][r][Z]
                    U[l][r][fz] += masks[l][r][fz] * s
            smooth(l, False, ghosts_valid=world > 1 and bool(lev(0, l)["dist"]))

        with np.errstate(all="ignore"):
            cycle(0)
        out = np.full(len(f0_global), np.nan)
        for r in range(world):
            a, b = own(r, 0)
            out[a:b] = U[0][r][a - lev(r, 0)["gz0"]:b - lev(r, 0)["gz0"]]
        return out, nlev, lrep

    rng = np.random.default_rng(11)
    for world, nz in ((2, 64), (4, 64), (8, 64), (3, 50), (5, 47)):
        f = rng.standard_normal(nz)
        for dirichlet in (False, True):
            mask = np.ones(nz)
            if dirichlet:
                mask[0] = 0.0
                f[0] = 0.0
            ref, nlev1, _ = vcycle(1, [nz], [nz], [0], f, mask)
            slabs = [partition.make_slab(nz, 64 * 64, r, world) for r in range(world)]
            owned = [s.z1 - s.z0 for s in slabs]
            got, nlev, lrep = vcycle(world, owned, [s.n_local // 4096 for s in slabs], [s.own0 // 4096 for s in slabs], f, mask)
            assert nlev == nlev1 and 1 <= lrep < nlev
            assert np.isfinite(got).all() and np.abs(got).max() < 1e6, "a poisoned (never exchanged) value reached the result"
            assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max(), (world, nz, dirichlet, np.abs(got - ref).max())


def _worker_round2(rank, world, port, q):
    """the N > 1 host logic added in round 2, over gloo: (a) the brittle law's global selection -- every rank lists the
    candidates of its own particle range, the lists are all-gathered, every rank runs lpmb_brittle_select on the same global
    list; (b) the preconditioned CG of the fast mode on slabs -- halo exchange of p before every product, THREE all-reduced dot
    products per iteration (p.Ap, r.r for the stop rule on the true residual, r.z), every rank taking the same decisions."""
    import torch
    import torch.distributed as dist
    import scipy.sparse as sp
    sys.path.insert(0, str(ROOT))
    capi = importlib.import_module("lpm-c_b200.capi")
    lpm_lat = importlib.import_module("lpm-c_b200.lattice")
    part = importlib.import_module("lpm-c_b200.partition")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 10
    lat = lpm_lat.sc_block(n)
    N, nn = n ** 3, 18
    s = part.make_slab(n, n * n, rank, world)
    g0 = s.first_global
    # ---- (a) brittle selection: strains = a hash of the global bond key with many ties
    key = np.arange(N * nn, dtype=np.int64)
    strain_all = ((key * 2654435761) % 97).astype(np.float64) / 8.0
    crit, nbreak = 11.5, 9
    mine_lo, mine_hi = (g0 + s.own0) * nn, (g0 + s.own1) * nn
    mk = key[mine_lo:mine_hi][strain_all[mine_lo:mine_hi] >= crit]
    lists = [None] * world
    dist.all_gather_object(lists, (mk.tolist(), strain_all[mk].tolist()))
    gk = np.array(sum((l[0] for l in lists), []), dtype=np.int64)        # rank order = ascending keys
    gs = np.array(sum((l[1] for l in lists), []))
    picked, _ = capi.brittle_select(gk, gs, nbreak)
    ref_keys = key[strain_all >= crit]
    ref_picked, _ = capi.brittle_select(ref_keys, strain_all[ref_keys], nbreak)
    sel_ok = np.array_equal(gk, ref_keys) and np.array_equal(picked, ref_picked) and len(picked) == nbreak
    # ---- (b) preconditioned CG on slabs (block-diagonal preconditioner = distributed trivially; same stop rule as pcg_run)
    conn = lat["conn"]
    rows = np.repeat(np.arange(N), conn.shape[1])[conn.ravel() >= 0]
    cols = conn.ravel()[conn.ravel() >= 0]
    vals = np.where(rows == cols, 70.0 + (rows % 7), -1.0 - ((rows + cols) % 5) / 5.0)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(N, N))
    dinv = 1.0 / A.diagonal()
    b = np.cos(3e-2 * np.arange(N))
    loc = slice(g0, g0 + s.n_local)
    Al = A[loc, loc].tocsr()
    own = np.zeros(s.n_local, dtype=bool)
    own[s.own0:s.own1] = True
    L = s.layer_size
    n_lo, n_hi = s.narrow_lo * L, s.narrow_hi * L

    def exchange(v):
        reqs = []
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(v[s.own0:s.own0 + s.send_narrow_lo * L].copy()), rank - 1))
            rl = torch.empty(n_lo, dtype=torch.float64)
            reqs.append(dist.irecv(rl, rank - 1))
        if rank < world - 1:
            reqs.append(dist.isend(torch.from_numpy(v[s.own1 - s.send_narrow_hi * L:s.own1].copy()), rank + 1))
            rh = torch.empty(n_hi, dtype=torch.float64)
            reqs.append(dist.irecv(rh, rank + 1))
        for r in reqs:
            r.wait()
        if rank > 0:
            v[s.own0 - n_lo:s.own0] = rl.numpy()
        if rank < world - 1:
            v[s.own1:s.own1 + n_hi] = rh.numpy()

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def pcg(matvec, rhs, mask, dot):
        x = np.zeros_like(rhs)
        r = rhs * mask
        rr0 = dot(r, r)
        thresh = 1e-8 * rr0 + 1e-12
        z = dinv_l * r if matvec is not global_mv else dinv * r
        rho = dot(r, z)
        p = z.copy()
        it = 0
        while it < 500:
            ap = matvec(p) * mask
            alpha = rho / dot(p, ap)
            x += alpha * p
            r -= alpha * ap
            it += 1
            if dot(r, r) <= thresh:
                break
            z = (dinv_l if matvec is not global_mv else dinv) * r
            rho_new = dot(r, z)
            p = z + (rho_new / rho) * p
            rho = rho_new
        return x, it

    dinv_l = dinv[loc]

    def local_mv(p):
        exchange(p)
        return Al @ p

    def global_mv(p):
        return A @ p

    x, it = pcg(local_mv, b[loc].copy(), own.astype(float), lambda u, v: allsum(u @ v))
    xr, itr = pcg(global_mv, b.copy(), np.ones(N), lambda u, v: float(u @ v))
    err = np.abs(x[s.own0:s.own1] - xr[g0 + s.own0:g0 + s.own1]).max() / np.abs(xr).max()
    errs = [None] * world
    dist.all_gather_object(errs, (it, float(err), bool(sel_ok)))
    if rank == 0:
        q.put((itr, errs))
    dist.destroy_process_group()


def test_round2_slab_logic_with_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30400 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker_round2, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    itr, errs = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert all(sel for _, _, sel in errs), "ranks disagree with the single-list brittle selection"
    assert all(it == itr for it, _, _ in errs) and 0 < itr < 100, (itr, errs)
    assert all(e <= 1e-12 for _, e, _ in errs), errs
