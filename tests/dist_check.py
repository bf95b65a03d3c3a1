"""Multi-GPU correctness: the slab-decomposed Newton iteration (NCCL halo exchange + all-reduce inside
liblpmb200.so) reproduces the single-GPU one.   torchrun --nproc-per-node 2 tests/dist_check.py [n]"""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lpm = importlib.import_module("lpm-c_b200")
    from importlib import import_module
    partition = import_module("lpm-c_b200.partition")
    slab = partition.make_slab(n, n * n, rank, world)
    uid = [lpm.Context.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    c, info = bench.build_workload(lpm, n, local, slab=slab, unique_id=uid[0])
    its, nrs = [], []
    for _ in range(2):          # two consecutive Newton iterations (no snapshot restore): state really evolves
        it, nr = c.newton_iteration(0, 1)
        its.append(it)
        nrs.append(nr)
    # end of the load step: nonlocal damage (Gaussian gather across the slab boundary), crack update, commit
    broken, _ = c.update_damage(0)
    c.update_crack()
    c.switch_state(1)
    own = slice(slab.own0, slab.own1)
    KEYS = ("xyz", "F", "stress_tensor", "dLp0", "damage_nonlocal0", "damage_w", "damage_broken", "Pin")
    mine = {k: torch.from_numpy(np.ascontiguousarray(c.get_field(k).reshape(slab.n_local, -1)[own])).cuda() for k in KEYS}
    gathered = {}
    for k, t in mine.items():
        sizes = [None] * world
        dist.all_gather_object(sizes, tuple(t.shape))
        bufs = [torch.empty(s, dtype=t.dtype, device=t.device) for s in sizes]
        dist.all_gather(bufs, t)
        gathered[k] = torch.cat(bufs).cpu().numpy()
    c.close()
    ok = True
    if rank == 0:
        c1, info1 = bench.build_workload(lpm, n, local, bricks=False)   # full-format SELL kernel as the cross-check
        assert abs(info1["norm_residual0"] - info["norm_residual0"]) <= 1e-12 * info1["norm_residual0"], (info1, info)
        its1, nrs1 = [], []
        for _ in range(2):
            it, nr = c1.newton_iteration(0, 1)
            its1.append(it)
            nrs1.append(nr)
        broken1, _ = c1.update_damage(0)
        c1.update_crack()
        c1.switch_state(1)
        print(f"world={world} n={n}: CG iterations dist {its} single {its1}; residual norms dist {nrs} single {nrs1}; "
              f"broken bonds dist {broken} single {broken1}")
        ok &= its == its1 and broken == broken1
        ok &= all(abs(a - b) <= 1e-9 * abs(b) for a, b in zip(nrs, nrs1))
        x0 = c1.get_field("xyz_initial")
        for k in gathered:
            ref = c1.get_field(k).reshape(n ** 3, -1)
            a, b = gathered[k], ref
            if k == "xyz":
                a, b = a - x0, b - x0
            err = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
            print(f"  {k}: rel.err {err:.2e}")
            ok &= err <= 1e-9
        c1.close()
        print("DIST_CHECK", "OK" if ok else "FAILED")
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
