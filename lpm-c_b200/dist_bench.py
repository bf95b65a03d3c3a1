"""Multi-GPU arm of bench.py: strong scaling of the same n^3 workload over particle slabs.

One process per GPU (torchrun).  torch.distributed is plumbing only: it carries the 128-byte NCCL unique id
to the ranks, provides the barrier and the max-over-ranks of the device-measured time.  The data path --
SpMV halo exchange and CG dot-product all-reduce -- is NCCL inside liblpmb200.so (csrc/lpmb_dist.cu).
"""
from __future__ import annotations

import json

import numpy as np


PARITY_KEYS = ("xyz", "F", "Pin", "stress_tensor", "dLp0", "damage_nonlocal0", "damage_w", "damage_broken")


def _parity_steps(c):
    """two consecutive Newton iterations (state really evolves), then the nonlocal damage update, updateCrack and the commit"""
    its, nrs = [], []
    for _ in range(2):
        it, nr = c.newton_iteration(0, 1)
        its.append(int(it))
        nrs.append(float(nr))
    broken, _ = c.update_damage(0)
    c.update_crack()
    c.switch_state(1)
    return its, nrs, int(broken)


def dist_parity(args, lpm, dist, rank, world, local, bench, n):
    """Before the timed region: the slab-decomposed path (this world size, the comm mode and SpMV kernel of the timed run)
    against ONE GPU running the full-format kernel on the same n^3 block of the bench workload -- CG iteration counts and
    broken-bond counts equal, residual norms and xyz / F / Pin / stress / plastic stretch / damage fields of every owned
    particle within 1e-9 (tests/dist_check_lite.py, made driver-visible).  Returns the report (rank 0) and ok (all ranks)."""
    import numpy as np
    import torch
    from . import partition
    slab = partition.make_slab(n, n * n, rank, world)
    uid = [lpm.Context.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    c, info = bench.build_workload(lpm, n, local, slab=slab, unique_id=uid[0], bricks=args.spmv == "bricks",
                                    brick_trim=not getattr(args, "no_brick_trim", False))
    its, nrs, broken = _parity_steps(c)
    own = slice(slab.own0, slab.own1)
    mine = {k: np.ascontiguousarray(c.get_field(k).reshape(slab.n_local, -1)[own]) for k in PARITY_KEYS}
    mine["_meta"] = (its, nrs, broken, int(c.dist_mode()), float(info["norm_residual0"]))
    c.close()
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    report = None
    if rank == 0:
        c1, info1 = bench.build_workload(lpm, n, local, bricks=False)
        its1, nrs1, broken1 = _parity_steps(c1)
        x0 = c1.get_field("xyz_initial")
        errs = {}
        for k in PARITY_KEYS:
            a = np.concatenate([q[k] for q in parts])
            b = c1.get_field(k).reshape(n ** 3, -1)
            if k == "xyz":
                a, b = a - x0, b - x0
            errs[k] = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
        c1.close()
        same_meta = all(q["_meta"][0] == its and q["_meta"][2] == broken for q in parts)
        ok = (same_meta and its == its1 and broken == broken1 and max(errs.values()) <= 1e-9
              and all(abs(a - b) <= 1e-9 * abs(b) for a, b in zip(nrs, nrs1))
              and abs(info["norm_residual0"] - info1["norm_residual0"]) <= 1e-12 * info1["norm_residual0"])
        report = {"ok": bool(ok), "lattice_n": n, "particles": n ** 3, "world": world, "comm_mode": [q["_meta"][3] for q in parts],
                  "cg_iterations_slabs": its, "cg_iterations_single_gpu": its1, "cg_iterations_equal": its == its1,
                  "broken_bonds_slabs": broken, "broken_bonds_single_gpu": broken1,
                  "rel_err": errs, "rel_err_residual_norms": [abs(a - b) / abs(b) for a, b in zip(nrs, nrs1)], "tolerance": 1e-9,
                  "checked": "2 Newton iterations + update_damage(0) + update_crack + switch_state(1); slabs on the timed run's SpMV kernel "
                             "and comm mode vs one GPU on the full-format kernel"}
    flag = torch.tensor([1 if (report and report["ok"]) else 0], device=f"cuda:{local}")
    dist.broadcast(flag, src=0)
    return report, bool(flag.item())


def run(args, lpm, dist, rank, world, local, bench):
    import torch
    from . import partition

    n = args.n
    parity_report, parity_ok = (None, True)
    if getattr(args, "dist_parity_n", 0) > 0:
        parity_report, parity_ok = dist_parity(args, lpm, dist, rank, world, local, bench, args.dist_parity_n)
        if not parity_ok:
            if rank == 0:
                print(json.dumps({"metric": bench.METRIC, "value": None, "n_gpus": world, "dist_parity": parity_report,
                                  "error": "slab-decomposed path does not reproduce the single-GPU path; no timing reported"}), flush=True)
            dist.barrier()
            dist.destroy_process_group()
            raise SystemExit(3)
    slab = partition.make_slab(n, n * n, rank, world)
    # rank 0 creates the NCCL id; torch.distributed ships it
    uid = [lpm.Context.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    c, info = bench.build_workload(lpm, n, local, slab=slab, unique_id=uid[0], bricks=args.spmv == "bricks",
                                    brick_trim=not getattr(args, "no_brick_trim", False))
    hbm_peak, peak_src = bench.peaks()
    comm_mode = c.dist_mode()

    def barrier():
        c.synchronize()
        torch.cuda.synchronize()
        dist.barrier()

    for _ in range(args.warmup):
        it, nr = bench.one_step(c)
    barrier()
    c.set_profiling(True)
    launches0 = c.launches
    sampler = bench.ClockSampler(local)
    sampler.start()
    stream = torch.cuda.ExternalStream(c.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    iters = []
    for _ in range(args.steps):
        it, nr = bench.one_step(c)
        iters.append(it)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_local = e0.elapsed_time(e1)
    t = torch.tensor([ms_local], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    launches = c.launches - launches0
    spmv_ms, spmv_calls = c.get_profile()
    c.set_profiling(False)
    spmv_avg_ms = spmv_ms / max(1, spmv_calls)
    # per-rank SpMV bytes: only the slices holding owned rows are streamed
    own_frac = (slab.own1 - slab.own0) / slab.n_local
    # (brick kernel: the needed z-layers of every class tile of owned +- 2 layers are streamed; lpmb_spmv_bytes_bricks counts exactly that)
    alg_bytes_rank = c.spmv_bytes_bricks() if info["bricks"] else int(c.spmv_bytes() * own_frac)
    kernel = ("brick_spmv_kernel + brick_gather_kernel<true> (symmetric CG SpMV + fused mask and p.Ap)" if info["bricks"]
              else "spmv_sell_kernel<3,true> (CG SpMV + fused p.Ap)")
    achieved = alg_bytes_rank / (spmv_avg_ms * 1e-3) / 1e9
    stats = torch.tensor([spmv_avg_ms, achieved, float(launches)], dtype=torch.float64, device=f"cuda:{local}")
    gathered = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(gathered, stats)

    # end to end: every rank uploads its slab's inputs from pinned host memory and reads its results back
    N = slab.n_local
    nd = 3 * N
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
    h_xyz, h_res = pin((N, 3), torch.float64), pin(nd, torch.float64)
    h_bc, h_fix = pin(nd, torch.int32), pin(nd, torch.int32)
    o_xyz, o_disp, o_pin, o_res = pin((N, 3), torch.float64), pin(nd, torch.float64), pin(nd, torch.float64), pin(nd, torch.float64)
    h_xyz[:] = c.get_field("xyz_save")
    h_res[:] = c.get_field("residual_save")
    h_bc[:] = c.get_field("dispBC_index")
    h_fix[:] = c.get_field("fix_index")

    import importlib
    capi = importlib.import_module("lpm-c_b200.capi")

    def e2e_step():   # straight between the pinned buffers and the device, exactly like the single-GPU arm (bench.py)
        for nm, a in (("xyz", h_xyz), ("residual", h_res), ("dispBC_index", h_bc), ("fix_index", h_fix)):
            capi._check(capi.lib.lpmb_field_set(c._h, nm.encode(), a.ctypes.data, a.size))
        r = c.newton_iteration(0, 1)
        for nm, a in (("xyz", o_xyz), ("disp", o_disp), ("Pin", o_pin), ("residual", o_res)):
            capi._check(capi.lib.lpmb_field_get(c._h, nm.encode(), a.ctypes.data, a.size))
        return r

    import time
    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = h_xyz.nbytes + h_res.nbytes + h_bc.nbytes + h_fix.nbytes
    d2h = o_xyz.nbytes + o_disp.nbytes + o_pin.nbytes + o_res.nbytes
    bytes_t = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(bytes_t, op=dist.ReduceOp.SUM)

    # ---- opt-in fast mode on slabs (param cg_precond = 1): CG preconditioned with the multigrid V-cycle of the single-GPU fast
    # mode, its hierarchy distributed over the slabs (lpmb_mg.cu, pcg_run).  Reported BESIDE the parity-mode headline.
    fast = None
    if not getattr(args, "no_fast_mode", False):
        try:
            c.set_params(cg_precond=1.0)
            for _ in range(2):
                bench.one_step(c)
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            f_iters, f_nr = [], []
            for _ in range(args.steps):
                itf, nrf = bench.one_step(c)
                f_iters.append(itf)
                f_nr.append(nrf)
            f1.record(stream)
            barrier()
            tf = torch.tensor([f0.elapsed_time(f1) / args.steps], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            f_ms = float(tf.item())
            fast = {"newton_it_per_s": 1000.0 / f_ms, "ms_per_step": f_ms, "pcg_iterations_per_step": f_iters,
                    "speedup_vs_parity_mode": (ms_total / args.steps) / f_ms, "norm_residual_after_the_iteration": f_nr[-1],
                    "norm_residual_after_the_iteration_parity_mode": nr,
                    "preconditioner": "matrix-free geometric multigrid V-cycle, ONE hierarchy distributed over the slabs (global coarse "
                                      "grids, two ghost layers exchanged before every stencil pass, coarse levels replicated); "
                                      "param cg_precond = 1",
                    "note": "not the parity path: the reference's CG is unpreconditioned (solver.c:219-220); same stop rule on the true residual"}
        except Exception as e:   # optional; the parity-mode line stands on its own
            fast = {"error": str(e)[:300]}
        finally:
            c.set_params(cg_precond=0.0)

    if rank == 0:
        ms_per_step = ms_total / args.steps
        Ng = n ** 3
        per_rank = [{"rank": r, "spmv_ms": float(g[0]), "spmv_GBs": float(g[1]), "launches": int(g[2])} for r, g in enumerate(gathered)]
        worst = min(per_rank, key=lambda d: d["spmv_GBs"])
        out = {
            "metric": bench.METRIC, "value": 1000.0 / ms_per_step, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5 physics (J2 plasticity + nonlocal damage) on a synthetic SC {n}^3 lattice, {Ng} particles, "
                                   f"{3 * Ng} DoF; Newton iteration 0 of load step 1 replayed from a snapshot",
                       "lattice_n": n, "particles": Ng, "dof": 3 * Ng, "cg_iterations_per_step": iters,
                       "cg_iterations_per_s": float(sum(iters)) / (ms_total * 1e-3),
                       "spmv_kernel": "brick-blocked symmetric" if info["bricks"] else "full-format SELL-32",
                       "cg": "unpreconditioned, rel 1e-8 / abs 1e-12 on squared norms (solver.c:217-222)",
                       "l2": "inputs larger than L2 (per-rank matrix %.1f GB)" % (c.spmv_bytes_stored() * own_frac / 1e9),
                       "parallelism": f"{world} z-slabs (owned layers per rank {slab.z1 - slab.z0}, 4 ghost layers); per CG iteration a halo "
                                      "exchange of p (2 layers) + all-reduce of 2 scalars",
                       "comm": {1: "NCCL", 2: "scalars: NVLink peer memory (CUDA IPC); halo: NCCL",
                                3: "scalars and halo push: NVLink peer memory (CUDA IPC); NCCL only outside the CG loop"}[comm_mode]},
            "roofline": {"bound": "hbm", "kernel": kernel + ", slowest rank", "achieved": worst["spmv_GBs"],
                         "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": worst["spmv_GBs"] / hbm_peak, "traffic": None,
                         "algorithmic_bytes_per_launch": alg_bytes_rank, "avg_launch_ms": worst["spmv_ms"], "per_rank": per_rank,
                         "share_of_step": float(spmv_ms / ms_local)},
            "e2e": {"value": 1.0 / e2e_s, "unit": bench.UNIT, "h2d_bytes_per_step": int(bytes_t[0].item()),
                    "d2h_bytes_per_step": int(bytes_t[1].item()), "steps": e2e_steps},
            "gpu_launches": int(sum(d["launches"] for d in per_rank)),
            "clocks": clocks,
            "dist_parity": parity_report,
        }
        if fast is not None:
            out["fast_mode"] = fast
        print(json.dumps(out), flush=True)
    c.close()
    dist.barrier()
    dist.destroy_process_group()
