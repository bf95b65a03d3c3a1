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
