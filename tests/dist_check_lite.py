"""Multi-GPU correctness without torch: the slab-decomposed Newton iteration (halo exchange + all-reduces inside
liblpmb200.so, brick SpMV streaming only the needed rows of every tile) reproduces the single-GPU full-format one.

    python tests/dist_check_lite.py [n=24] [world=2]

Spawns `world` processes itself (one per GPU); the NCCL unique id travels through a file.  Same checks as
tests/dist_check.py (which rendezvouses through torch.distributed), a fraction of its start-up time."""
import importlib
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
KEYS = ("xyz", "F", "stress_tensor", "dLp0", "damage_nonlocal0", "damage_w", "damage_broken", "Pin")


def run_steps(c):
    its, nrs = [], []
    for _ in range(2):          # two consecutive Newton iterations (no snapshot restore): state really evolves
        it, nr = c.newton_iteration(0, 1)
        its.append(it)
        nrs.append(nr)
    broken, _ = c.update_damage(0)
    c.update_crack()
    c.switch_state(1)
    return its, nrs, broken


def brittle_phase(c, first_global=0, n_total=None):
    """Brittle law on slabs (updateBrittleDamage, constitutive.c:1437-1526: the nbreak largest bond strains are selected
    GLOBALLY).  Bond stretches are replaced by a hash of the GLOBAL bond id -- well separated values, so that the selection
    cannot hinge on the 1e-15 differences between a slab run and a single-GPU run -- with ~60 candidates above the critical
    strain and nbreak = 7 of them to break; then the elastic law on the damaged lattice (the ghosts' broken flags enter the
    owned rows through the neighbours' shell sums, so F / Pin of the owned rows check them too)."""
    nn, n_local = c.nn, c.N
    n_total = n_total or n_local
    L0 = c.get_field("distance_initial")
    key = (np.uint64(first_global) + np.arange(n_local, dtype=np.uint64))[:, None] * np.uint64(nn) + np.arange(nn, dtype=np.uint64)[None, :]
    h = ((key * np.uint64(2654435761)) % np.uint64(1000003)).astype(np.float64) / 1000003.0
    c.set_field("dL", h * 2e-3 * L0)
    c.set_params(critical_bstrain=2e-3 * (1.0 - 60.0 / (n_total * nn)), nbreak=7.0)
    cand, pairs = c.update_damage(6)
    c.update_crack()
    c.bond_force(6)
    return [cand] + (sorted(int(p) for p in np.asarray(pairs)[:, 0] if p >= 0) if len(pairs) else [])


def next_step_of(c):
    """the first Newton iteration of the next load step (assembly, predictor, residual, solve, constitutive update)"""
    c.copy_field("xyz_temp", "xyz")
    c.copy_field("F_temp", "F")
    c.fd_stiffness(False)
    c.bond_force(4)
    c.update_rr()
    return c.newton_iteration(0, 1)


def child(rank, world, n, d):
    import bench
    lpm = importlib.import_module("lpm-c_b200")
    partition = importlib.import_module("lpm-c_b200.partition")
    uid_file = Path(d) / "uid.bin"
    if rank == 0:
        uid = lpm.Context.dist_unique_id()
        (Path(d) / "uid.tmp").write_bytes(uid)
        os.replace(Path(d) / "uid.tmp", uid_file)
    else:
        t0 = time.time()
        while not uid_file.exists():
            if time.time() - t0 > 30:
                raise SystemExit("no unique id from rank 0")
            time.sleep(0.05)
        uid = uid_file.read_bytes()
    slab = partition.make_slab(n, n * n, rank, world)
    c, info = bench.build_workload(lpm, n, rank, slab=slab, unique_id=uid)
    its, nrs, broken = run_steps(c)
    own = slice(slab.own0, slab.own1)
    # per-rank snapshot / resume of a slab run (lpmb_io.cu): save, run on, wipe, load, run on again -> identical
    snap = Path(d) / f"rank{rank}.snap"
    c.snapshot_save(snap)

    def next_step():
        return next_step_of(c)

    a = next_step()
    xa = c.get_field("xyz")
    c.set_field("xyz", np.zeros_like(xa))
    c.snapshot_load(snap)
    b = next_step()
    xb = c.get_field("xyz")
    c.snapshot_load(snap)
    b2 = next_step()
    xb2 = c.get_field("xyz")
    print(f"rank {rank}: snapshot resume: live {a} | loaded {b} | loaded again {b2}; xyz live-vs-loaded max {np.abs(xa - xb).max():.3e}, "
          f"loaded-vs-loaded {np.abs(xb - xb2).max():.3e}", flush=True)
    assert b == b2 and np.array_equal(xb, xb2), ("two resumes from the same slab snapshot differ", b, b2)
    assert a[0] == b[0] and abs(a[1] - b[1]) <= 1e-12 * abs(a[1]) and np.abs(xa - xb).max() <= 1e-13, ("slab snapshot resume differs", a, b)
    # opt-in fast mode on slabs (CG preconditioned with one multigrid V-cycle per slab = block-Jacobi over the ranks,
    # lpmb_mg.cu / pcg_run): the same step from the same snapshot -- far fewer iterations, the same stop rule on the true
    # residual, and a Newton residual after the iteration within the accepted solve error of the parity-mode one
    c.snapshot_load(snap)
    c.set_params(cg_precond=1.0)
    f = next_step()
    c.set_params(cg_precond=0.0)
    print(f"rank {rank}: fast mode on slabs: {f[0]} PCG iterations (parity mode {b[0]} CG iterations), Newton residual {f[1]:.6e} vs {b[1]:.6e}", flush=True)
    assert 0 < f[0] <= b[0] // 2 and abs(f[1] - b[1]) <= 1e-3 * abs(b[1]), ("fast mode on slabs", f, b)
    # ... and the same with one hierarchy per slab and no communication inside the V-cycle (block-Jacobi over the slabs)
    c.snapshot_load(snap)
    c.set_params(cg_precond=1.0, mg_dist=0.0)
    fj = next_step()
    c.set_params(cg_precond=0.0, mg_dist=1.0)
    print(f"rank {rank}: fast mode, block-Jacobi over the slabs: {fj[0]} PCG iterations", flush=True)
    assert 0 < fj[0] < b[0] and abs(fj[1] - b[1]) <= 1e-3 * abs(b[1]), ("block-Jacobi fast mode on slabs", fj, b)
    # back to the state after the load step, then the brittle selection across the slabs
    c.snapshot_load(snap)
    brittle = brittle_phase(c, slab.first_global, n ** 3)
    out = {k: np.ascontiguousarray(c.get_field(k).reshape(slab.n_local, -1)[own]) for k in KEYS}
    np.savez(Path(d) / f"rank{rank}.npz", its=np.array(its), nrs=np.array(nrs), broken=np.array([broken] + brittle), mode=np.array([c.dist_mode()]), fast=np.array([f[0], b[0], fj[0]]), fast_nr=np.array([f[1]]),
             norm0=np.array([info["norm_residual0"]]), spmv_bytes=np.array([c.spmv_bytes_bricks()]), **out)
    c.close()


def main():
    if "--rank" in sys.argv:
        a = sys.argv
        child(int(a[a.index("--rank") + 1]), int(a[a.index("--world") + 1]), int(a[a.index("--n") + 1]), a[a.index("--dir") + 1])
        return
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    with tempfile.TemporaryDirectory() as d:
        procs = [subprocess.Popen([sys.executable, __file__, "--rank", str(r), "--world", str(world), "--n", str(n), "--dir", d])
                 for r in range(world)]
        rcs = [p.wait(timeout=240) for p in procs]
        assert rcs == [0] * world, rcs
        parts = [np.load(Path(d) / f"rank{r}.npz") for r in range(world)]
        import bench
        lpm = importlib.import_module("lpm-c_b200")
        c1, info1 = bench.build_workload(lpm, n, 0, bricks=False)   # full-format SELL kernel as the cross-check
        its1, nrs1, broken1 = run_steps(c1)
        # the distributed multigrid hierarchy is the single-GPU one: same PCG iteration count, same Newton residual
        snap1 = Path(d) / "single.snap"
        c1.snapshot_save(snap1)
        c1.set_params(cg_precond=1.0)
        f1 = next_step_of(c1)
        c1.set_params(cg_precond=0.0)
        c1.snapshot_load(snap1)
        fd_its, fd_nr = int(parts[0]["fast"][0]), float(parts[0]["fast_nr"][0])
        print(f"fast mode: slabs {fd_its} PCG iterations, Newton residual {fd_nr:.9e}; one GPU {f1[0]} iterations, {f1[1]:.9e}; "
              f"block-Jacobi over the slabs {int(parts[0]['fast'][2])} iterations; parity mode {int(parts[0]['fast'][1])} CG iterations")
        fast_ok = abs(fd_its - f1[0]) <= 1 and abs(fd_nr - f1[1]) <= 1e-5 * abs(f1[1])
        broken1 = [broken1] + brittle_phase(c1)
        its, nrs, broken = list(parts[0]["its"]), list(parts[0]["nrs"]), [int(x) for x in parts[0]["broken"]]
        assert all([int(x) for x in p["broken"]] == broken for p in parts), "ranks disagree on the broken bonds"
        print(f"world={world} n={n}: CG iterations dist {its} single {its1}; residual norms dist {nrs} single {nrs1}; "
              f"broken bonds [nonlocal law, brittle candidates, particles of the 7 brittle breaks] dist {broken} single {broken1}; comm mode {[int(p['mode'][0]) for p in parts]}; fast mode on slabs {int(parts[0]['fast'][0])} PCG iterations against {int(parts[0]['fast'][1])} CG iterations; "
              f"brick SpMV bytes per rank {[int(p['spmv_bytes'][0]) for p in parts]}")
        ok = its == its1 and broken == broken1 and fast_ok
        ok &= all(abs(a - b) <= 1e-9 * abs(b) for a, b in zip(nrs, nrs1))
        ok &= abs(float(parts[0]["norm0"][0]) - info1["norm_residual0"]) <= 1e-12 * info1["norm_residual0"]
        x0 = c1.get_field("xyz_initial")
        for k in KEYS:
            a = np.concatenate([p[k] for p in parts])
            b = c1.get_field(k).reshape(n ** 3, -1)
            if k == "xyz":
                a, b = a - x0, b - x0
            err = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
            print(f"  {k}: rel.err {err:.2e}")
            ok &= err <= 1e-9
        c1.close()
        print("DIST_CHECK", "OK" if ok else "FAILED")
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
