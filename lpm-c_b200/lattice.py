"""Synthetic lattice builders (host, numpy) for benchmarks and tests.

These produce exactly the arrays the reference's set-up code produces for an axis-aligned
simple-cubic block -- `neighbors`/`nsign` in ascending-j order with the two shells interleaved
(reference src/neighbor.c:16-41) and `conn` = sorted unique(self + direct neighbours + 1st-of-1st
+ 2nd-of-2nd) (src/neighbor.c:56-112) -- but in O(N) from the lattice stencil instead of the
reference's O(N^2) all-pairs search.  Particle order is x-fastest, z-slowest like
src/initialization.c:266-284.  tests/test_lattice.py checks them bit for bit against the
reference's own output on the default 21^3 case.

They are plain index arithmetic (no physics) and are not on the timed path; the device-side
builder for arbitrary point sets is lpmb_build_topology.
"""
from __future__ import annotations

import itertools

import numpy as np

SC_NN = 18          # nneighbors  (initialization.c:251)
SC_NCONN = 61       # nneighbors_AFEM + 1 (initialization.c:256)


def sc_offsets():
    """(first-shell offsets, second-shell offsets, conn offsets) of the simple-cubic lattice"""
    first = [o for o in itertools.product((-1, 0, 1), repeat=3) if sum(abs(v) for v in o) == 1]
    second = [o for o in itertools.product((-1, 0, 1), repeat=3) if sum(abs(v) for v in o) == 2]
    conn = {(0, 0, 0)}
    conn.update(first)
    conn.update(second)
    for a in first:
        for b in first:
            conn.add((a[0] + b[0], a[1] + b[1], a[2] + b[2]))
    for a in second:
        for b in second:
            conn.add((a[0] + b[0], a[1] + b[1], a[2] + b[2]))
    return first, second, sorted(conn)


def _stencil_table(nx, ny, nz, offsets, fill=-1):
    """[N][len(offsets)] neighbour index per offset (fill where outside the block)"""
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    # particle index = x + nx*(y + ny*z): x fastest
    ix = ix.transpose(2, 1, 0).ravel()
    iy = iy.transpose(2, 1, 0).ravel()
    iz = iz.transpose(2, 1, 0).ravel()
    N = nx * ny * nz
    out = np.full((N, len(offsets)), fill, dtype=np.int64)
    for k, (ox, oy, oz) in enumerate(offsets):
        jx, jy, jz = ix + ox, iy + oy, iz + oz
        ok = (jx >= 0) & (jx < nx) & (jy >= 0) & (jy < ny) & (jz >= 0) & (jz < nz)
        out[ok, k] = (jx + nx * (jy + ny * jz))[ok]
    return out


def sc_block(nx: int, ny: int | None = None, nz: int | None = None, h: float = 0.5, origin=(0.0, 0.0, 0.0)):
    """Axis-aligned simple-cubic block.  Returns dict(xyz, neighbors, nsign, conn, nb, nb_conn)."""
    ny = ny or nx
    nz = nz or nx
    N = nx * ny * nz
    first, second, conn_off = sc_offsets()
    # direct neighbours, ascending j, -1 padded at the end
    offs = first + second
    shell = np.array([0] * len(first) + [1] * len(second), dtype=np.int32)
    tab = _stencil_table(nx, ny, nz, offs)
    big = np.iinfo(np.int64).max
    key = np.where(tab >= 0, tab, big)
    order = np.argsort(key, axis=1, kind="stable")
    nbr = np.take_along_axis(tab, order, axis=1).astype(np.int32)
    nsign = np.where(nbr >= 0, shell[order], -1).astype(np.int32)
    nb = (nbr >= 0).sum(axis=1).astype(np.int32)
    # conn
    ctab = _stencil_table(nx, ny, nz, conn_off)
    ckey = np.where(ctab >= 0, ctab, big)
    ckey.sort(axis=1)
    conn = np.where(ckey == big, -1, ckey).astype(np.int32)
    nb_conn = (conn >= 0).sum(axis=1).astype(np.int32)
    i = np.arange(N)
    xyz = np.empty((N, 3), dtype=np.float64)
    xyz[:, 0] = origin[0] + h * (i % nx)
    xyz[:, 1] = origin[1] + h * ((i // nx) % ny)
    xyz[:, 2] = origin[2] + h * (i // (nx * ny))
    return {"xyz": xyz, "neighbors": nbr, "nsign": nsign, "conn": conn, "nb": nb, "nb_conn": nb_conn,
            "shape": (nx, ny, nz), "h": h}


def k_pointer(conn: np.ndarray, dim: int) -> np.ndarray:
    """K_pointer[N+1][2] as src/neighbor.c:114-130 computes it (64-bit)."""
    N = conn.shape[0]
    ge = ((conn >= np.arange(N)[:, None]) & (conn >= 0)).sum(axis=1)
    inc = dim * dim * ge - (3 if dim == 3 else 1)
    kp = np.zeros((N + 1, 2), dtype=np.int64)
    kp[:N, 0] = ge
    kp[1:, 1] = np.cumsum(inc)
    return kp
