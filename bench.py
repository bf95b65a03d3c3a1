#!/usr/bin/env python
"""bench.py -- Newton iterations / second of the LPM hot path on B200, with the SpMV HBM roofline.

Metric (BASELINE.json): "Newton iters/sec + PCG SpMV HBM GB/s at 1/2/4/8 B200 vs MKL host CPU".
  value     = Newton iterations / s, whole job, state resident in HBM
  roofline  = the CG SpMV (dominant kernel) against the measured HBM peak
  e2e       = the same Newton iteration through the C ABI with HOST buffers (H2D + D2H every step)
  step      = one pass of the reference's Newton loop body (src/lpmc_project.c:426-464):
              switchStateV(0); displacement-BC treatment of K/residual; solverCG (+ xyz += disp);
              computeBondForceGeneral(plmode 0: J2 elastoplastic); computeStress; updateRR; ||residual||.
              Every step replays Newton iteration 0 of load step 1 from the same snapshot, so all steps
              do identical work (the snapshot restore is inside the timed region; it is two device copies).
  workload  = BASELINE.json configs[4] ("C5"): examples/CT_sc_ductile_nonlocal.c physics (E=115e3, nu=0.28,
              sigma_y=955, H=2401.8, damagec_A=400, damage_L=0.6, threshold 0.85) on a synthetic simple-cubic
              block of n^3 particles (n=216 -> 10 077 696 particles, 30.2M DoF), bottom z-layer fixed in z,
              top z-layer displaced in z (displacement control like the CT example).

Launch:  python bench.py --gpus N --steps K --warmup W      (N>1: under torchrun, one rank per GPU)
         python bench.py --impl reference ...               (the reference's own CPU code, see below)
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "newton_iterations_per_second"
UNIT = "Newton it/s"

# C5 family physics (examples/CT_sc_ductile_nonlocal.c:171,195-202,230-235) on radius 0.25
PHYS = dict(radius=0.25, E0=115e3, mu0=0.28, sigmay=955.0, J2_xi=0.0, J2_H=2401.8, damagec_A=400.0, damage_L=0.6,
            damage_threshold=0.85, damageb_A=10.0, nbreak=20, critical_bstrain=1.0e-2)
STRAIN_STEP = 1.0e-2   # top layer displaced by 1 % of the block height (plastic from the first iteration)


def peaks():
    try:
        p = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- GPU arm
def build_workload(lpm, n: int, device: int, slab=None, unique_id: bytes | None = None, bricks: bool = True, brick_trim: bool = True):
    """set up the n^3 block (or this rank's slab of it: owned layers + ghosts, see lpm-c_b200/partition.py) on the
    device: lattice -> O(N) topology -> material -> first FD tangent -> BCs -> predictor -> residual; snapshot"""
    t0 = time.time()
    h = 2 * PHYS["radius"]
    if slab is None:
        first, N = 0, n ** 3
    else:
        first, N = slab.first_global, slab.n_local
    i = np.arange(first, first + N, dtype=np.int64)   # global particle indices, x fastest / z slowest
    xyz = np.empty((N, 3))
    xyz[:, 0] = h * (i % n)
    xyz[:, 1] = h * ((i // n) % n)
    xyz[:, 2] = h * (i // (n * n))
    typ = np.zeros(N, dtype=np.int32)
    typ[i // (n * n) == n - 1] = 1      # top layer   (type 1: displaced)
    typ[i // (n * n) == 0] = 2          # bottom layer (type 2: fixed in z)
    c = lpm.Context(N, 3, 2, 18, 61, device=device)
    if slab is not None:
        c.dist_init(unique_id, slab.rank, slab.world)
        c.dist_set_slab(*slab.set_slab_args())
    E0, mu0 = PHYS["E0"], PHYS["mu0"]
    C11 = E0 * (1.0 - mu0) / (1.0 + mu0) / (1.0 - 2.0 * mu0)
    C12 = E0 * mu0 / (1.0 + mu0) / (1.0 - 2.0 * mu0)
    C44 = E0 / 2.0 / (1.0 + mu0)
    c.set_params(radius=PHYS["radius"], particle_volume=h ** 3, J2_H=PHYS["J2_H"], J2_xi=PHYS["J2_xi"],
                 damage_L=PHYS["damage_L"], damage_threshold=PHYS["damage_threshold"], damagec_A=PHYS["damagec_A"],
                 nbreak=PHYS["nbreak"], critical_bstrain=PHYS["critical_bstrain"])
    c.set_field("xyz", xyz)
    c.set_field("xyz_initial", xyz)
    del xyz
    c.build_topology(h, np.sqrt(2.0) * h)
    c.set_field("type", typ)
    c.set_field("sigmay", np.full(N, PHYS["sigmay"]))
    c.calc_kntv(np.tile([C11, C12, C44], (3, 1)))
    c.compute_dl()
    t1 = time.time()
    # load step 1 up to the first Newton iteration (lpmc_project.c:387-414)
    c.copy_field("xyz_temp", "xyz")
    c.copy_field("F_temp", "F")
    c.copy_field("Pex_temp", "Pex")
    c.fd_stiffness(False)
    c.synchronize()
    t2 = time.time()
    c.apply_disp_bc(2, "z", 0.0)
    c.apply_disp_bc(1, "z", STRAIN_STEP * h * (n - 1))
    c.bond_force(4)
    nr, nf = c.update_rr()
    c.copy_field("xyz_save", "xyz")
    c.copy_field("residual_save", "residual")
    if bricks:
        if not brick_trim:      # A/B switch: stream every row of every class tile (slab runs; a single GPU has no partial tiles at 216)
            c.set_params(brick_trim=0.0)
        if os.environ.get("LPMB_BRICK_LAZY_WAIT", "0") not in ("", "0"):
            # experimental (slab runs with the peer-memory halo push): wait for the neighbours' pushes brick by brick, interior
            # brick layers first (brick_spmv_kernel<true>); not measured yet, off by default
            c.set_params(brick_lazy_wait=1.0)
        c.enable_bricks(True)   # brick-blocked symmetric SpMV for the CG (lpmb_brick.cu); SC lattice: eligible
    c.synchronize()
    info = {"N": N, "n": n, "setup_s": round(t1 - t0, 2), "fd_assembly_s": round(t2 - t1, 3), "norm_residual0": nr,
            "norm_reaction0": nf, "bricks": bool(bricks)}
    return c, info


def one_step(c):
    c.copy_field("xyz", "xyz_save")
    c.copy_field("residual", "residual_save")
    return c.newton_iteration(0, 1)


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lpm = importlib.import_module("lpm-c_b200")
    if lpm.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    if world > 1:
        from importlib import import_module
        return import_module("lpm-c_b200.dist_bench").run(args, lpm, dist, rank, world, local, sys.modules[__name__])

    n = args.n
    c, info = build_workload(lpm, n, local, bricks=args.spmv == "bricks", brick_trim=not args.no_brick_trim)
    N = info["N"]
    hbm_peak, peak_src = peaks()

    def barrier():
        c.synchronize()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        it, nr = one_step(c)
    barrier()
    c.set_profiling(True)
    launches0 = c.launches
    sampler = ClockSampler(local)
    sampler.start()
    stream = torch.cuda.ExternalStream(c.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    iters = []
    for _ in range(args.steps):
        it, nr = one_step(c)
        iters.append(it)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = c.launches - launches0
    spmv_ms, spmv_calls = c.get_profile()
    c.set_profiling(False)
    ms_per_step = ms_total / args.steps
    value = 1000.0 / ms_per_step
    spmv_avg_ms = spmv_ms / max(1, spmv_calls)
    # achieved GB/s is computed from the bytes of the representation that is streamed -- the brick
    # format stores each symmetric block once -- and the full-format (both triangles) figure is reported beside it
    full_bytes = c.spmv_bytes()
    alg_bytes = c.spmv_bytes_bricks() if info["bricks"] else full_bytes
    achieved = alg_bytes / (spmv_avg_ms * 1e-3) / 1e9
    kernel = ("brick_spmv_kernel + brick_gather_kernel<true> (symmetric CG SpMV + fused mask and p.Ap)" if info["bricks"]
              else "spmv_sell_kernel<3,true> (CG SpMV + fused p.Ap)")
    traffic = None
    try:
        prof = json.load(open(ROOT / "profiles" / "spmv_dram_traffic.json"))
        key = "bricks" if info["bricks"] else "full"
        if int(prof[key].get("n", -1)) == n:
            traffic = prof[key]["dram_bytes_per_launch"]
    except Exception:
        pass

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    nd = 3 * N
    h_xyz = torch.empty((N, 3), dtype=torch.float64, pin_memory=True).numpy()
    h_res = torch.empty(nd, dtype=torch.float64, pin_memory=True).numpy()
    h_bc = torch.empty(nd, dtype=torch.int32, pin_memory=True).numpy()
    h_fix = torch.empty(nd, dtype=torch.int32, pin_memory=True).numpy()
    o_xyz = torch.empty((N, 3), dtype=torch.float64, pin_memory=True).numpy()
    o_disp = torch.empty(nd, dtype=torch.float64, pin_memory=True).numpy()
    o_pin = torch.empty(nd, dtype=torch.float64, pin_memory=True).numpy()
    o_res = torch.empty(nd, dtype=torch.float64, pin_memory=True).numpy()
    h_xyz[:] = c.get_field("xyz_save")
    h_res[:] = c.get_field("residual_save")
    h_bc[:] = c.get_field("dispBC_index")
    h_fix[:] = c.get_field("fix_index")
    capi = importlib.import_module("lpm-c_b200.capi")

    def e2e_step():
        for nm, a in (("xyz", h_xyz), ("residual", h_res), ("dispBC_index", h_bc), ("fix_index", h_fix)):
            capi._check(capi.lib.lpmb_field_set(c._h, nm.encode(), a.ctypes.data, a.size))
        it, nr = c.newton_iteration(0, 1)
        for nm, a in (("xyz", o_xyz), ("disp", o_disp), ("Pin", o_pin), ("residual", o_res)):
            capi._check(capi.lib.lpmb_field_get(c._h, nm.encode(), a.ctypes.data, a.size))
        return it, nr

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        it_e, nr_e = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d = h_xyz.nbytes + h_res.nbytes + h_bc.nbytes + h_fix.nbytes
    d2h = o_xyz.nbytes + o_disp.nbytes + o_pin.nbytes + o_res.nbytes
    assert np.isfinite(o_xyz).all() and np.isfinite(nr_e)

    # ---- opt-in fast mode: CG preconditioned with the matrix-free multigrid V-cycle (lpmb_mg.cu).  Reported BESIDE the
    # parity-mode headline above, never instead of it: another Krylov sequence, same stop rule on the true residual.
    fast = None
    if not args.no_fast_mode:
        try:
            c.set_params(cg_precond=1.0)
            for _ in range(2):
                one_step(c)
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            f_iters, f_nr = [], []
            for _ in range(args.steps):
                itf, nrf = one_step(c)
                f_iters.append(itf)
                f_nr.append(nrf)
            f1.record(stream)
            barrier()
            f_ms = f0.elapsed_time(f1) / args.steps
            fast = {"newton_it_per_s": 1000.0 / f_ms, "ms_per_step": f_ms, "pcg_iterations_per_step": f_iters,
                    "speedup_vs_parity_mode": ms_per_step / f_ms, "norm_residual_after_the_iteration": f_nr[-1],
                    "norm_residual_after_the_iteration_parity_mode": nr,
                    "preconditioner": "matrix-free geometric multigrid V-cycle (2+2 damped block-Jacobi sweeps, trilinear transfer, "
                                      "one 61-point stencil of the assembled tangent for all levels), param cg_precond = 1",
                    "note": "not the parity path: the reference's CG is unpreconditioned (solver.c:219-220); same stop rule on the true residual"}
        except Exception as e:   # the fast mode is optional; the parity-mode line stands on its own
            fast = {"error": str(e)[:300]}
        finally:
            c.set_params(cg_precond=0.0)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"C5 physics (J2 plasticity + nonlocal damage) on a synthetic SC {n}^3 lattice, "
                               f"{N} particles, {nd} DoF; Newton iteration 0 of load step 1 replayed from a snapshot",
                   "lattice_n": n, "particles": N, "dof": nd, "cg_iterations_per_step": iters,
                   "cg_iterations_per_s": float(sum(iters)) / (ms_total * 1e-3),   # secondary metric of SURVEY 8(d)
                   "cg": "unpreconditioned, rel 1e-8 / abs 1e-12 on squared norms (solver.c:217-222)",
                   "l2": "inputs larger than L2 (matrix %.1f GB)" % (c.spmv_bytes_stored() / 1e9),
                   "setup_s": info["setup_s"], "fd_assembly_s": info["fd_assembly_s"], "parallelism": "1 GPU",
                   "spmv_kernel": "brick-blocked symmetric (each block of the symmetric tangent streamed once)" if info["bricks"]
                                  else "full-format SELL-32 (both triangles)"},
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved,
                     "full_format_equivalent_gbs": full_bytes / (spmv_avg_ms * 1e-3) / 1e9,
                     "full_format_bytes_per_launch": full_bytes,
                     "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                     "stored_bytes_per_launch": alg_bytes if info["bricks"] else c.spmv_bytes_stored(), "avg_launch_ms": spmv_avg_ms,
                     "launches_timed": spmv_calls, "share_of_step": spmv_ms / ms_total},
        "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "fast_mode": fast,
    }
    c.close()
    if not args.no_cpu_baseline:
        out.update(cpu_reference_sample(args, n_full=n, with_gpu_dropin=True, device=local))
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------- CPU arm
class _QuietStdout:
    """The reference's solverCG() printf()s a line per solve (solver.c:257): keep the C-level stdout away from this
    process's stdout, which carries exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self._saved, 1)
        os.close(self._saved)
        os.close(self._null)
        return False


# CG iterations of Newton iteration 0 of load step 1 of this workload, by lattice size n (n^3 particles).  MEASURED: 116 at
# 48^3 and 226 at 100^3 by the reference's own solverCG() and by the GPU arm alike (parity_at_sample / reference arm print
# them), 458 at 216^3 by the GPU arm (every BENCH / SCALE line, config.cg_iterations_per_step) -- the reference cannot
# set 216^3 up at all (32-bit K_pointer, neighbor.c:114-134).  Used to carry a measured CPU time to the 216^3 config.
CG_ITERS_MEASURED = {48: 116, 100: 226, 216: 458}


def usable_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))   # the cores this process may run on (os.cpu_count() ignores an affinity mask)
    except AttributeError:
        return os.cpu_count() or 1


def reference_newton_sample(m: int, steps: int, warmup: int, cores: int, keep_outputs: bool = False) -> dict:
    """One load-step start + `warmup + steps` replays of Newton iteration 0 on an SC m^3 block of the bench physics, through
    the reference-named entry points of whatever oracle.ref.REF_SO is loaded in THIS process: oracle/_ref/liblpmc_ref.so (the
    unmodified reference, CPU) or oracle/_ref/liblpmc_b200host.so (the reference's host code + the GPU drop-in library).
    Topology by oracle.ref.inject_sc_topology (O(N), bit-identical to the reference's O(N^2) search).  The reference's FD
    assembly races under OpenMP when threads work on nearby layers (measured here: 8 threads on 32 layers
    -> K wrong by 2e-4, 494 instead of 79 CG iterations), so it runs on m // 12 threads (>= 12 layers per static chunk;
    48^3: 1 vs 4 threads agree to 2e-16 in K, 5e-16 in disp)."""
    from oracle import ref as oref
    r = oref.RefLPM.instance()
    r.threads(cores)
    hi = 0.5 * (m - 1)
    t_setup = time.perf_counter()
    for extra in (0.0, 0.25, -0.2):
        box = (-0.2, hi + extra, -0.2, hi + extra, -0.2, hi + extra)
        r.setup_sc(box=box, radius=PHYS["radius"], E0=PHYS["E0"], mu0=PHYS["mu0"],
                   plmode=0, sigmay=PHYS["sigmay"], J2_xi=PHYS["J2_xi"], J2_H=PHYS["J2_H"], nbreak=PHYS["nbreak"],
                   critical_bstrain=PHYS["critical_bstrain"], damageb_A=PHYS["damageb_A"], damagec_A=PHYS["damagec_A"],
                   damage_threshold=PHYS["damage_threshold"], damage_L=PHYS["damage_L"], top_z="auto", neighbor_search="lattice")
        if r.N == m ** 3:
            break
    assert r.N == m ** 3, (r.N, m)
    t_setup = time.perf_counter() - t_setup
    N, L = r.N, r.lib
    z = r.get("xyz")[:, 2]
    height = float(z.max() - z.min())
    dropin = hasattr(L, "lpmc_dropin_last_cg_iterations")
    fd_threads = max(1, min(cores, m // 12))
    L.omp_set_num_threads(fd_threads)
    t0 = time.perf_counter()
    # reference types from setup_sc: 1 = top layer, 2 = bottom layer; same BCs as the GPU arm
    nr0, nf0 = r.begin_step([(2, "z", 0.0), (1, "z", STRAIN_STEP * height)], [])
    t_fd = time.perf_counter() - t0
    L.omp_set_num_threads(cores)
    xyz_s, res_s = r.get("xyz"), r.get("residual")
    times, t_solve, its = [], [], []

    def timed_solve():
        a = time.perf_counter()
        L.solverCG()
        t_solve.append(time.perf_counter() - a)
        its.append(int(L.lpmc_dropin_last_cg_iterations() if dropin else L.lpmb_shim_last_itercount()))

    for k in range(warmup + steps):
        r.put("xyz", xyz_s)
        r.put("residual", res_s)
        a = time.perf_counter()
        nr = r.newton_iteration(hooks={"solve": timed_solve})
        if k >= warmup:
            times.append(time.perf_counter() - a)
    out = {"m": m, "particles": N, "cores": cores, "fd_threads": fd_threads, "setup_s": t_setup, "begin_step_s": t_fd,
           "t_step": float(np.mean(times)), "t_cg": float(np.mean(t_solve[warmup:])), "cg_iterations": int(its[-1]), "steps": steps,
           "norm_residual0": float(nr0), "norm_residual1": float(nr), "impl": "dropin" if dropin else "cpu"}
    if keep_outputs:
        # one more, UNTIMED replay whose outputs are compared (parity_at_sample): the reference's threaded force loop writes
        # its neighbours' dL / dL_total rows from several threads -- usually benign, but one bench run
        # on the 16-core box came back with F off by 4.6e-4 while disp agreed to 8e-15; single-threaded it is deterministic
        def law_on_one_thread():
            L.omp_set_num_threads(1)
            L.computeBondForceGeneral(r.gi("plmode"), 1)
            L.omp_set_num_threads(cores)
        r.put("xyz", xyz_s)
        r.put("residual", res_s)
        r.newton_iteration(hooks={"solve": timed_solve, "bondforce": law_on_one_thread})
        out["arrays"] = {"disp": r.get("disp"), "F": r.get("F"), "Pin": r.get("Pin"), "xyz0": xyz_s, "xyz": r.get("xyz")}
    return out


def extrapolate_to(n_full: int, smp: dict) -> float:
    """seconds per Newton iteration at n_full^3 from a measured m^3 sample: per-particle cost, CG time additionally scaled
    with the MEASURED iteration counts (CG_ITERS_MEASURED), the rest of the iteration linearly"""
    m = smp["m"]
    scale_n = (n_full ** 3) / smp["particles"]
    it_ratio = CG_ITERS_MEASURED[n_full] / smp["cg_iterations"] if n_full in CG_ITERS_MEASURED else n_full / m
    return smp["t_cg"] * scale_n * it_ratio + (smp["t_step"] - smp["t_cg"]) * scale_n


def dropin_sample_child(args):
    """child process of the GPU arm (LPMB_REF_SO = oracle/_ref/liblpmc_b200host.so): the same sample through the reference's
    host code + reference-named GPU drop-in entry points; results to args.sample_child (npz)"""
    with _QuietStdout():
        smp = reference_newton_sample(args.cpu_sample_n, args.cpu_steps, 1, args.cpu_threads or usable_cores(), keep_outputs=True)
    arrays = smp.pop("arrays")
    np.savez(args.sample_child, meta=np.array([json.dumps(smp)]), **arrays)


def run_dropin_sample(args, m: int, steps: int, device: int = 0, device_bc: bool = False):
    """GPU drop-in run of the sample in a child process (its own CUDA context; the reference keeps its state in process
    globals, so the CPU and the drop-in build cannot share a process)"""
    import tempfile
    host_so = ROOT / "oracle" / "_ref" / "liblpmc_b200host.so"
    if not host_so.exists():
        return None, None
    with tempfile.TemporaryDirectory() as d:
        out = Path(d) / "dropin.npz"
        env = dict(os.environ, LPMB_REF_SO=str(host_so), LPMB_DEVICE=str(device))
        if device_bc:
            env["LPMB_DROPIN_DEVICE_BC"] = "1"
        cmd = [sys.executable, str(ROOT / "bench.py"), "--sample-child", str(out), "--cpu-sample-n", str(m), "--cpu-steps", str(steps),
               "--cpu-threads", str(args.cpu_threads or 0)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        if r.returncode != 0 or not out.exists():
            return {"error": (r.stderr or r.stdout)[-500:]}, None
        z = np.load(out)
        return json.loads(str(z["meta"][0])), {k: z[k] for k in z.files if k != "meta"}


def _rel(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cpu_reference_sample(args, n_full: int, with_gpu_dropin: bool = False, device: int = 0) -> dict:
    """cpu_baseline leg of the GPU arm: the reference's OWN functions (oracle/_ref = its unmodified sources + open MKL
    stand-in, NOT Intel MKL) on all host cores on the m^3 sample (default 48^3), carried to n_full^3 with the measured
    CG iteration counts.  with_gpu_dropin: the same sample once more through liblpmc_dropin.so (child process) ->
    parity_at_sample (CG iterations equal, disp / F / Pin to 1e-9) and same_config_sample (one measured GPU/CPU pair on
    identical inputs through the reference's own API)."""
    from oracle import ref as oref
    if not oref.available():
        return {"cpu_baseline": {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}}
    m, cores = args.cpu_sample_n, args.cpu_threads or usable_cores()
    with _QuietStdout():
        smp = reference_newton_sample(m, args.cpu_steps, 1, cores, keep_outputs=with_gpu_dropin)
    cpu_arrays = smp.pop("arrays", None)
    t_full = extrapolate_to(n_full, smp)
    res = {"cpu_baseline": {
        "value": 1.0 / t_full, "unit": UNIT, "cores": cores, "kind": "reference",
        "sample": (f"reference's own switchStateV / setDispBC_stiffnessUpdate3D / solverCG / computeBondForceGeneral(0) / updateRR "
                   f"(unmodified sources + open MKL stand-in with a threaded symmetric SpMV, not Intel MKL) on an SC {m}^3 block "
                   f"({smp['particles']} particles): {smp['t_step']:.3f} s per Newton iteration ({smp['cg_iterations']} CG its, solverCG "
                   f"{smp['t_cg']:.3f} s), {smp['steps']} timed; carried to {n_full}^3 as t_cg*(N/Ns)*({CG_ITERS_MEASURED.get(n_full)}/"
                   f"{smp['cg_iterations']} measured CG iterations) + t_rest*(N/Ns); the --impl reference arm measures 100^3"),
        "sample_newton_it_per_s": 1.0 / smp["t_step"], "sample_particles": smp["particles"], "sample_cg_iterations": smp["cg_iterations"],
        "sample_fd_assembly_s": smp["begin_step_s"], "sample_fd_threads": smp["fd_threads"]}}
    if not with_gpu_dropin:
        return res
    g, g_arrays = run_dropin_sample(args, m, args.cpu_steps, device)
    if g is None or "error" in g:
        res["parity_at_sample"] = {"ok": False, "error": "drop-in sample did not run: " + str(g)}
        return res
    x0 = cpu_arrays["xyz0"]
    par = {"particles": smp["particles"], "through": "reference host code + liblpmc_dropin.so (strict mode: K_global re-imported per solve)",
           "cg_iterations_reference": smp["cg_iterations"], "cg_iterations_gpu": g["cg_iterations"],
           "cg_iterations_equal": smp["cg_iterations"] == g["cg_iterations"],
           "rel_err_disp": _rel(g_arrays["disp"], cpu_arrays["disp"]), "rel_err_F": _rel(g_arrays["F"], cpu_arrays["F"]),
           "rel_err_Pin": _rel(g_arrays["Pin"], cpu_arrays["Pin"]),
           "rel_err_xyz_moved": _rel(g_arrays["xyz"] - x0, cpu_arrays["xyz"] - x0), "tolerance": 1e-9}
    par["ok"] = bool(par["cg_iterations_equal"] and max(par["rel_err_disp"], par["rel_err_F"], par["rel_err_Pin"]) <= 1e-9)
    res["parity_at_sample"] = par
    same = {"particles": smp["particles"], "cpu_it_per_s": 1.0 / smp["t_step"], "cpu_cores": cores,
            "gpu_it_per_s": 1.0 / g["t_step"], "gpu_through": "liblpmc_dropin.so, host arrays in / out on every call",
            "gpu_begin_step_s": g["begin_step_s"], "cpu_begin_step_s": smp["begin_step_s"], "cpu_fd_threads": smp["fd_threads"]}
    g2, _ = run_dropin_sample(args, m, args.cpu_steps, device, device_bc=True)
    if g2 and "error" not in g2:
        same["gpu_it_per_s_device_bc"] = 1.0 / g2["t_step"]   # LPMB_DROPIN_DEVICE_BC=1: tangent stays on the device (DoF mask)
    if args.dropin_s1 > 0:
        g3, _ = run_dropin_sample(args, args.dropin_s1, 2, device)
        if g3 and "error" not in g3:
            same["s1"] = {"particles": g3["particles"], "gpu_it_per_s": 1.0 / g3["t_step"], "cg_iterations_gpu": g3["cg_iterations"],
                          "gpu_begin_step_s": g3["begin_step_s"], "cpu_it_per_s": "see the --impl reference line (s1_measured)"}
    res["same_config_sample"] = same
    return res


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref = unmodified sources + open MKL
    stand-in) on all host threads.  The 216^3 config cannot be set up by the reference at all (32-bit CSR offsets), so:
      * the K timed `steps` (after W warm-ups) are Newton iteration 0 on the 48^3 sample -- `ms_per_step` is their MEASURED
        wall time, the same sample the GPU arm's same_config_sample / parity_at_sample use;
      * S1 = 100^3 is set up with the injected O(N) topology and timed for --ref-s1-steps Newton
        iterations; `value` carries THAT measurement to 216^3 with the measured CG iteration counts (226 -> 458)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import ref as oref
    if not oref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built"}), flush=True)
        return
    cores = args.cpu_threads or usable_cores()
    n = args.n
    with _QuietStdout():
        small = reference_newton_sample(args.cpu_sample_n, args.steps, max(1, args.warmup), cores)
        s1 = reference_newton_sample(args.ref_s1_n, args.ref_s1_steps, 1, cores) if args.ref_s1_n > 0 else None
    anchor = s1 or small
    t_full = extrapolate_to(n, anchor)
    value = 1.0 / t_full
    N = n ** 3
    describe = lambda s: {"particles": s["particles"], "newton_it_per_s": 1.0 / s["t_step"], "s_per_newton_iteration": s["t_step"],
                          "solverCG_s": s["t_cg"], "cg_iterations": s["cg_iterations"], "steps_timed": s["steps"],
                          "fd_assembly_s": s["begin_step_s"], "fd_threads": s["fd_threads"], "setup_s": s["setup_s"],
                          "extrapolated_to_config_it_per_s": 1.0 / extrapolate_to(n, s)}
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": (f"unmodified reference sources + open MKL stand-in (not Intel MKL), {cores} OpenMP threads; measured: SC "
                       f"{anchor['m']}^3 ({anchor['particles']} particles) {anchor['t_step']:.3f} s per Newton iteration "
                       f"({anchor['cg_iterations']} CG its, solverCG {anchor['t_cg']:.3f} s); carried to {n}^3 as t_cg*(N/Ns)*"
                       f"({CG_ITERS_MEASURED.get(n)}/{anchor['cg_iterations']} measured CG iterations) + t_rest*(N/Ns)")}
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * small["t_step"],
           "ms_per_step_is": f"MEASURED wall time of one timed step = Newton iteration 0 on the {args.cpu_sample_n}^3 sample; "
                             f"`value` is the {n}^3 config (ms_per_step_config below), carried over from the measured {anchor['m']}^3 run",
           "ms_per_step_config": 1000.0 * t_full,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"C5 physics (J2 plasticity + nonlocal damage) on a synthetic SC {n}^3 lattice, {N} particles; "
                                  "Newton iteration 0 of load step 1", "lattice_n": n, "particles": N,
                      "parallelism": f"{cores} host threads (OpenMP)"},
           "cpu_baseline": base,
           "same_config_sample": {"particles": small["particles"], "cpu_it_per_s": 1.0 / small["t_step"], "cpu_cores": cores,
                                  "gpu_it_per_s": "see the b200 arm's same_config_sample"},
           "sample_measured": describe(small), "s1_measured": describe(s1) if s1 else None,
           "cg_iterations_measured": CG_ITERS_MEASURED,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lattice-n", dest="n", type=int, default=int(os.environ.get("LPMB_BENCH_N", 216)), help="lattice points per side")
    ap.add_argument("--cpu-sample-n", type=int, default=48)   # 110 592 particles: ~1 s per reference Newton iteration
    ap.add_argument("--cpu-steps", type=int, default=4)       # FD assembly ~20 s + 5 Newton iterations: ~25-30 s of CPU work
    ap.add_argument("--ref-s1-n", type=int, default=100, help="--impl reference: lattice size of the measured S1 anchor (0 = skip)")
    ap.add_argument("--ref-s1-steps", type=int, default=1)
    ap.add_argument("--dropin-s1", type=int, default=100, help="b200 arm: also run the drop-in sample at this lattice size (0 = skip)")
    ap.add_argument("--dist-parity-n", type=int, default=40, help="N>1: lattice size of the slabs-vs-single-GPU check run before the timed region (0 = skip)")
    ap.add_argument("--sample-child", default="", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spmv", default="bricks", choices=["bricks", "full"],
                    help="CG SpMV kernel: brick-blocked symmetric (default) or the full-format SELL kernel")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the opt-in preconditioned (multigrid) fast-mode line")
    ap.add_argument("--no-brick-trim", action="store_true", help="A/B: stream whole class tiles instead of only the needed z-layers")
    args = ap.parse_args()
    if args.sample_child:
        dropin_sample_child(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
