"""Particle-slab partition of a z-slowest ordered lattice across the GPUs of one box (host logic only).

The reference numbers particles z-slowest (src/initialization.c:266-284), so a contiguous global index
range is a z-slab.  Rank r holds the layers [z0-g_lo, z1+g_hi): its owned layers [z0, z1) plus
`ghost_layers` (4) ghost layers towards each existing neighbour rank; the two ghost layers next to the
owned range are computed redundantly (complete 2-hop stars), the outer two supply positions only
(csrc/lpmb_dist.cu).  `narrow_layers` (2 = reach of conn) is what the CG exchanges per iteration.

Everything here is integer bookkeeping; tests/test_partition.py checks it with world_size-2 gloo.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class Slab:
    rank: int
    world: int
    layer_size: int      # particles per z-layer
    nz: int              # global number of layers
    z0: int              # first owned layer (global)
    z1: int              # one past the last owned layer
    g_lo: int            # ghost layers below / above
    g_hi: int
    narrow_lo: int       # layers exchanged per CG iteration with rank-1 / rank+1
    narrow_hi: int
    send_narrow_lo: int  # layers this rank sends down / up per CG iteration
    send_narrow_hi: int
    send_wide_lo: int    # layers this rank sends down / up in a full ghost refresh
    send_wide_hi: int

    # ---- particle counts (what lpmb_dist_set_slab takes) ----
    @property
    def n_local(self) -> int:
        return (self.z1 - self.z0 + self.g_lo + self.g_hi) * self.layer_size

    @property
    def own0(self) -> int:
        return self.g_lo * self.layer_size

    @property
    def own1(self) -> int:
        return (self.g_lo + self.z1 - self.z0) * self.layer_size

    @property
    def first_global(self) -> int:
        """global index of local particle 0"""
        return (self.z0 - self.g_lo) * self.layer_size

    def set_slab_args(self):
        L = self.layer_size
        return (self.own0, self.own1, self.narrow_lo * L, self.narrow_hi * L, self.send_narrow_lo * L, self.send_narrow_hi * L,
                self.send_wide_lo * L, self.send_wide_hi * L)

    def as_dict(self):
        return asdict(self)


def owned_layers(nz: int, rank: int, world: int):
    base, rem = divmod(nz, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def make_slab(nz: int, layer_size: int, rank: int, world: int, ghost_layers: int = 4, narrow_layers: int = 2) -> Slab:
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    z0, z1 = owned_layers(nz, rank, world)
    if world > 1 and (z1 - z0) < ghost_layers:
        raise ValueError(f"each rank must own at least {ghost_layers} layers (nz={nz}, world={world})")

    def ghosts(r):
        a, b = owned_layers(nz, r, world)
        return (min(ghost_layers, a) if r > 0 else 0), (min(ghost_layers, nz - b) if r < world - 1 else 0)

    g_lo, g_hi = ghosts(rank)
    n_lo, n_hi = min(narrow_layers, g_lo), min(narrow_layers, g_hi)
    # what my neighbours expect from me = their ghost / narrow depth on the side facing me
    lo_wants_wide = ghosts(rank - 1)[1] if rank > 0 else 0
    hi_wants_wide = ghosts(rank + 1)[0] if rank < world - 1 else 0
    return Slab(rank, world, layer_size, nz, z0, z1, g_lo, g_hi, n_lo, n_hi, min(narrow_layers, lo_wants_wide),
                min(narrow_layers, hi_wants_wide), lo_wants_wide, hi_wants_wide)


# ---- ragged layers: carved specimens (compact tension, notches, holes) -------------------------------------------
# "irregular specimens split at layer boundaries balanced on sum(nb_conn)".  The helpers of the
# reference that carve a specimen keep the z-slowest order (initialization.c:299-313, 869-905, 1058-1094), so a slab is
# still a contiguous index range -- only the number of particles per layer varies.  lpmb_dist_set_slab takes particle
# counts, so nothing changes below the ABI.

@dataclass(frozen=True)
class RaggedSlab:
    rank: int
    world: int
    nz: int
    z0: int                   # owned layers [z0, z1)
    z1: int
    g_lo: int                 # ghost layers below / above
    g_hi: int
    first_global: int         # global index of local particle 0
    n_local: int
    own0: int                 # owned local index range
    own1: int
    narrow_recv_lo: int       # particle counts, the arguments of lpmb_dist_set_slab in its order
    narrow_recv_hi: int
    narrow_send_lo: int
    narrow_send_hi: int
    wide_send_lo: int
    wide_send_hi: int
    weight: float             # sum of the layer weights this rank owns

    def set_slab_args(self):
        return (self.own0, self.own1, self.narrow_recv_lo, self.narrow_recv_hi, self.narrow_send_lo, self.narrow_send_hi,
                self.wide_send_lo, self.wide_send_hi)

    def as_dict(self):
        return asdict(self)


def layer_counts(z, spacing: float, tol: float = 1e-6):
    """particles per z-layer of a z-slowest ordered lattice (z = third coordinate of xyz_initial in index order);
    raises if the order is not layer by layer or a layer index is skipped (then index ranges are not slabs)"""
    import numpy as np
    z = np.asarray(z, dtype=np.float64)
    if z.size == 0:
        raise ValueError("empty lattice")
    f = (z - z.min()) / spacing
    k = np.rint(f).astype(np.int64)
    if np.abs(f - k).max() > tol:
        raise ValueError("particles are not on z-layers of the given spacing")
    if (np.diff(k) < 0).any():
        raise ValueError("particles are not numbered z-slowest: a contiguous index range is not a slab")
    counts = np.bincount(k)
    if (counts == 0).any():
        raise ValueError("an empty z-layer separates the specimen: slabs on either side do not couple")
    return counts.tolist()


def balanced_cuts(weights, world: int, min_layers: int):
    """cut points c[0]=0 < c[1] < ... < c[world]=nz of the contiguous partition of the layers that minimises the largest
    per-rank weight, every rank owning at least `min_layers` layers (exact dynamic programme: nz and world are tiny)"""
    nz = len(weights)
    if world < 1 or nz < world * max(1, min_layers):
        raise ValueError(f"{nz} layers cannot give {world} ranks at least {max(1, min_layers)} layers each")
    m = max(1, min_layers)
    pre = [0.0]
    for w in weights:
        pre.append(pre[-1] + float(w))
    INF = float("inf")
    # best[k][j] = minimal bottleneck of splitting layers [0, j) into k slabs
    best = [[INF] * (nz + 1) for _ in range(world + 1)]
    arg = [[-1] * (nz + 1) for _ in range(world + 1)]
    best[0][0] = 0.0
    for k in range(1, world + 1):
        for j in range(k * m, nz - (world - k) * m + 1):
            for i in range((k - 1) * m, j - m + 1):
                if best[k - 1][i] == INF:
                    continue
                v = max(best[k - 1][i], pre[j] - pre[i])
                if v < best[k][j]:       # strict: ties keep the lowest cut (deterministic on every rank)
                    best[k][j], arg[k][j] = v, i
    cuts = [nz]
    for k in range(world, 0, -1):
        cuts.append(arg[k][cuts[-1]])
    return cuts[::-1]


def make_ragged_slab(counts, rank: int, world: int, weights=None, ghost_layers: int = 4, narrow_layers: int = 2) -> RaggedSlab:
    """slab of `rank` for layers with counts[z] particles; `weights[z]` = the work of layer z (sum of nb_conn over its
    particles = its share of the SpMV bytes; default: the particle count).  Every rank computes the same cuts."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    counts = [int(c) for c in counts]
    nz = len(counts)
    weights = counts if weights is None else list(weights)
    if len(weights) != nz:
        raise ValueError("weights and counts differ in length")
    cuts = balanced_cuts(weights, world, ghost_layers if world > 1 else 1)
    off = [0]
    for c in counts:
        off.append(off[-1] + c)

    def ghosts(r):
        a, b = cuts[r], cuts[r + 1]
        return (min(ghost_layers, a) if r > 0 else 0), (min(ghost_layers, nz - b) if r < world - 1 else 0)

    def npart(za, zb):
        return off[zb] - off[za]

    z0, z1 = cuts[rank], cuts[rank + 1]
    g_lo, g_hi = ghosts(rank)
    n_lo, n_hi = min(narrow_layers, g_lo), min(narrow_layers, g_hi)
    # what the neighbours expect from me: their ghost / narrow depth on the side facing me, counted in MY top / bottom layers
    lo_wide = ghosts(rank - 1)[1] if rank > 0 else 0
    hi_wide = ghosts(rank + 1)[0] if rank < world - 1 else 0
    lo_narrow, hi_narrow = min(narrow_layers, lo_wide), min(narrow_layers, hi_wide)
    return RaggedSlab(
        rank=rank, world=world, nz=nz, z0=z0, z1=z1, g_lo=g_lo, g_hi=g_hi,
        first_global=off[z0 - g_lo], n_local=npart(z0 - g_lo, z1 + g_hi),
        own0=npart(z0 - g_lo, z0), own1=npart(z0 - g_lo, z1),
        narrow_recv_lo=npart(z0 - n_lo, z0), narrow_recv_hi=npart(z1, z1 + n_hi),
        narrow_send_lo=npart(z0, z0 + lo_narrow), narrow_send_hi=npart(z1 - hi_narrow, z1),
        wide_send_lo=npart(z0, z0 + lo_wide), wide_send_hi=npart(z1 - hi_wide, z1),
        weight=float(sum(weights[z0:z1])))
