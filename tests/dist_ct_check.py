"""Multi-GPU run of BASELINE config 5 AS SHIPPED: the carved, pre-cracked compact-tension specimen of
examples/CT_sc_ductile_nonlocal.c (75 030 particles = 15 z-layers of 5 002; two pin holes and a notch cut out of the plate,
an initial crack defined on the bonds) on `world` z-slabs over the C ABI, against the SERIAL all-CPU run of the unchanged
example (tests/golden/c5src_log.txt): Newton iteration counts, the CG iteration count of EVERY solve (400 288 | 400 366 |
400 351 ...) and the printed residual / reaction norms of every load step.

    python tests/dist_ct_check.py [world=2] [steps=3]

Geometry, particle types and the initial crack come from the reference's own host code (oracle/_ref/liblpmc_ref.so:
createCuboid / createCrack / removeCircle / setTypeRect / defineCrack, examples/CT_sc_ductile_nonlocal.c:60-237) in the
parent process -- test infrastructure; the slabs never see it: they get xyz / type / initial bond damage slices and build
their topology on the device.  Material, damage law and the nine displacement BCs per step are the example's (:171-300)."""
import importlib
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"
RADIUS = 0.33
STEP = -6.4e-3
DBP = [(3, "x", 0.0), (3, "y", -STEP), (3, "z", 0.0), (4, "x", 0.0), (4, "y", STEP), (4, "z", 0.0), (5, "y", 0.0), (6, "y", 0.0), (6, "z", 0.0)]


def reference_setup(path):
    """geometry / types / initial crack by the reference's host code -> npz"""
    import ctypes as C
    from oracle.ref import RefLPM
    r = RefLPM.instance()
    r.threads(len(os.sched_getaffinity(0)))
    r.setup_ct_geometry()
    r.threads(1)
    L, N = r.lib, r.N
    pc1, pc2 = (9.9, 13.13), (9.9, 34.91)
    ca1, ca2, w, ch = -10.0, 18.0, 1.0, 24.021
    ntype = 0
    r.set_ptr("type", L.allocInt1D(N, ntype))
    ntype += 1
    rects = [(pc2[0] - 5.2, pc2[0] + 5.2, pc2[1], 100.0, -100.0, 100.0), (pc1[0] - 5.2, pc1[0] + 5.2, -100.0, pc1[1], -100.0, 100.0),
             (pc2[0] - RADIUS, pc2[0] + RADIUS, 40.2 - 2 * RADIUS, 40.2 + RADIUS, -100.0, 100.0),
             (pc1[0] - RADIUS, pc1[0] + RADIUS, 7.9 - RADIUS, 7.9 + 2 * RADIUS, -100.0, 100.0),
             (50.0 - 2 * RADIUS, 100.0, 24.0 - 2 * RADIUS, 24.0 + 2 * RADIUS, -100.0, 100.0),
             (50.0 - 2 * RADIUS, 100.0, 24.0 - 2 * RADIUS, 24.0 + 2 * RADIUS, 5.1 - RADIUS, 5.1 + RADIUS),
             (-100.0, 1.5 * RADIUS, 22.0, 22.5, -100.0, 100.0), (-100.0, 1.5 * RADIUS, 25.5, 26.0, -100.0, 100.0)]
    for rc in rects:
        L.setTypeRect(*rc, ntype)
        ntype += 1
    r.si("ntype", ntype)
    L.defineCrack.argtypes = [C.c_double] * 3
    L.defineCrack.restype = None
    L.defineCrack(ca1, ca2 + w, ch)
    xyz = r.get("xyz")
    z = np.rint((xyz[:, 2] - xyz[:, 2].min()) / (2 * RADIUS)).astype(int)
    counts = np.bincount(z)
    assert np.all(np.diff(z) >= 0) and len(set(counts.tolist())) == 1, "specimen is not a stack of equal z-layers numbered z-slowest"
    np.savez(path, xyz=xyz, type=r.get("type"), damage_broken=r.get("damage_broken"), damage_w=r.get("damage_w"), damage_D=r.get("damage_D"),
             nb=r.get("nb"), neighbors=r.get("neighbors"), nz=len(counts), layer=int(counts[0]), ntype=ntype,
             particle_volume=r.gd("particle_volume"), cut1=r.gd("neighbor1_cutoff"), cut2=r.gd("neighbor2_cutoff"))


def child(rank, world, steps, d):
    lpm = importlib.import_module("lpm-c_b200")
    partition = importlib.import_module("lpm-c_b200.partition")
    g = np.load(Path(d) / "setup.npz")
    uid_file = Path(d) / "uid.bin"
    if rank == 0:
        (Path(d) / "uid.tmp").write_bytes(lpm.Context.dist_unique_id())
        os.replace(Path(d) / "uid.tmp", uid_file)
    else:
        t0 = time.time()
        while not uid_file.exists():
            if time.time() - t0 > 60:
                raise SystemExit("no unique id from rank 0")
            time.sleep(0.05)
    uid = uid_file.read_bytes()
    nz, layer = int(g["nz"]), int(g["layer"])
    slab = partition.make_slab(nz, layer, rank, world)
    sl = slice(slab.first_global, slab.first_global + slab.n_local)
    N = slab.n_local
    c = lpm.Context(N, 3, 2, 18, 61, device=rank)
    if world > 1:
        c.dist_init(uid, rank, world)
        c.dist_set_slab(*slab.set_slab_args())
    E0, mu0 = 115e3, 0.28
    C11, C12, C44 = E0 * (1 - mu0) / (1 + mu0) / (1 - 2 * mu0), E0 * mu0 / (1 + mu0) / (1 - 2 * mu0), E0 / 2 / (1 + mu0)
    c.set_params(radius=RADIUS, particle_volume=float(g["particle_volume"]), J2_H=2401.8, J2_xi=0.0, damage_L=0.6, damage_threshold=0.85,
                 damagec_A=400.0, nbreak=20, critical_bstrain=1.0e-2)
    xyz, typ = g["xyz"][sl], g["type"][sl]
    c.set_field("xyz", xyz)
    c.set_field("xyz_initial", xyz)
    c.build_topology(float(g["cut1"]), float(g["cut2"]))
    # device topology == the reference's lists (re-numbered to the slab)
    nb_ref = g["neighbors"][sl]
    own = slice(slab.own0, slab.own1)
    mine = c.get_field("neighbors")[own]
    want = np.where(nb_ref[own] >= 0, nb_ref[own] - slab.first_global, -1)
    assert np.array_equal(mine, want), "device neighbour lists of the owned particles differ from the reference's"
    c.set_field("type", typ)
    sig = np.full(N, 955.0)
    sig[(typ >= 1) & (typ <= 6)] = 1e6
    c.set_field("sigmay", sig)
    c.calc_kntv(np.tile([C11, C12, C44], (int(g["ntype"]), 1)))
    c.compute_dl()
    c.set_field("damage_broken", g["damage_broken"][sl])
    c.set_field("damage_w", g["damage_w"][sl])
    c.set_field("damage_D0", np.ascontiguousarray(g["damage_D"][sl][..., 0]))
    c.set_field("damage_D1", np.ascontiguousarray(g["damage_D"][sl][..., 1]))
    c.set_field("nb", g["nb"][sl])
    out = {"newton": [], "cg": [], "nr0": [], "nf0": [], "t": []}
    for step in range(1, steps + 1):
        t0 = time.time()
        c.copy_field("xyz_temp", "xyz")
        c.copy_field("F_temp", "F")
        c.copy_field("Pex_temp", "Pex")
        c.fd_stiffness(True)
        for t, ax, s in DBP:
            c.apply_disp_bc(t, ax, s)
        c.bond_force(4)
        nr, nf = c.update_rr()
        out["nr0"].append(nr)
        out["nf0"].append(nf)
        tol, ni, cg = max(nr, nf), 0, []
        while nr > 1e-4 * tol and ni < 100:
            it, nr = c.newton_iteration(0, 1)
            cg.append(it)
            ni += 1
        broken, _ = c.update_damage(0)
        c.update_crack()
        c.switch_state(1)
        assert broken == 0, "a bond broke within the first load steps (the golden run has none)"
        c.synchronize()
        out["newton"].append(ni)
        out["cg"].append(cg)
        out["t"].append(time.time() - t0)
    np.savez(Path(d) / f"rank{rank}.npz", newton=np.array(out["newton"]), cg=np.array(sum(out["cg"], [])), nr0=np.array(out["nr0"]),
             nf0=np.array(out["nf0"]), t=np.array(out["t"]), mode=np.array([c.dist_mode()]))
    c.close()


def main():
    a = sys.argv
    if "--rank" in a:
        child(int(a[a.index("--rank") + 1]), int(a[a.index("--world") + 1]), int(a[a.index("--steps") + 1]), a[a.index("--dir") + 1])
        return
    if "--setup" in a:
        reference_setup(a[a.index("--setup") + 1])
        return
    world = int(a[1]) if len(a) > 1 else 2
    steps = int(a[2]) if len(a) > 2 else 3
    with tempfile.TemporaryDirectory() as d:
        t0 = time.time()
        subprocess.run([sys.executable, __file__, "--setup", str(Path(d) / "setup.npz")], check=True, timeout=600, stdout=subprocess.DEVNULL)
        t_setup = time.time() - t0
        procs = [subprocess.Popen([sys.executable, __file__, "--rank", str(r), "--world", str(world), "--steps", str(steps), "--dir", d])
                 for r in range(world)]
        rcs = [p.wait(timeout=600) for p in procs]
        assert rcs == [0] * world, rcs
        res = [np.load(Path(d) / f"rank{r}.npz") for r in range(world)]
    log = (GOLD / "c5src_log.txt").read_text()
    g_cg = [int(m) for m in re.findall(r"The system has been solved after (\d+) iterations", log)]
    g_newton = [int(m) for m in re.findall(r"Loading step \d+ has finished in (\d+) iterations", log)]
    g_norms = [(float(x), float(y)) for x, y in re.findall(r"Norm of residual is (\S+), norm of reaction is (\S+),", log)]
    r0 = res[0]
    newton, cg = r0["newton"].tolist(), r0["cg"].tolist()
    print(f"world={world}: {steps} load steps of the CT specimen; Newton {newton} (golden {g_newton[:steps]}); CG {cg} (golden {g_cg[:len(cg)]}); "
          f"norms {[(float(a), float(b)) for a, b in zip(r0['nr0'], r0['nf0'])]} (golden {g_norms[:steps]}); "
          f"{r0['t'].round(2).tolist()} s per load step; comm mode {[int(q['mode'][0]) for q in res]}; reference set-up {t_setup:.1f} s")
    ok = newton == g_newton[:steps] and cg == g_cg[:len(cg)]
    for k in range(steps):
        ok &= abs(float(r0["nr0"][k]) / g_norms[k][0] - 1) <= 2e-5 and abs(float(r0["nf0"][k]) / g_norms[k][1] - 1) <= 2e-5
    for q in res[1:]:
        ok &= q["cg"].tolist() == cg and q["newton"].tolist() == newton
    print("DIST_CT_CHECK", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
