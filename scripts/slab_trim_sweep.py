"""Per-rank SpMV of a slab run, emulated on one GPU: the brick kernel on the slab shapes of the 216^3 case with and without
streaming only the needed rows of every class tile (param brick_trim; test hook brick_own_z0 / brick_own_z1).
usage: python scripts/slab_trim_sweep.py [shape index ...]       (run on the GPU box)"""
import importlib
import sys

sys.path.insert(0, ".")
lpm = importlib.import_module("lpm-c_b200")

# (nz, own_z0, own_z1): interior slab of 8 ranks, edge slabs of 8 ranks, interior slab of 4 ranks, slab of 2 ranks
SHAPES = (("8 ranks, interior", 31, (2, 29)), ("8 ranks, rank 0", 29, (0, 27)), ("8 ranks, rank 7", 29, (2, 29)),
          ("4 ranks, interior", 58, (2, 56)), ("2 ranks, rank 0", 110, (0, 108)))
PICK = [int(a) for a in sys.argv[1:]] or range(len(SHAPES))
for name, nz, own in [SHAPES[k] for k in PICK]:
    lat = lpm.lattice.sc_block(216, 216, nz)
    N = lat["xyz"].shape[0]
    c = lpm.Context(N, 3, 2, 18, 61)
    c.set_params(radius=0.25, brick_own_z0=own[0], brick_own_z1=own[1])
    c.set_field("xyz_initial", lat["xyz"])
    c.set_connectivity(lat["conn"])
    del lat
    c.fill_test_pattern()
    out = []
    for trim in (0.0, 1.0):
        c.set_params(brick_trim=trim)
        c.enable_bricks(True)
        c.spmv_bench(5, 2)
        t = c.spmv_bench(40, 2)
        out.append((t, c.spmv_bytes_bricks()))
        c.enable_bricks(False)
    (t0, b0), (t1, b1) = out
    print(f"{name}: {nz} layers, owned {own}: all rows {t0:.4f} ms ({b0 / 1e9:.3f} GB)  needed rows {t1:.4f} ms ({b1 / 1e9:.3f} GB, "
          f"{b1 / t1 / 1e6:.0f} GB/s)  speed-up {t0 / t1:.3f}", flush=True)
    c.close()
