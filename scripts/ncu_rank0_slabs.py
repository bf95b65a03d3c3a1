"""per-kernel launch list of ONE rank of a slab run (the other ranks run unprofiled): rank 0 is started under
`ncu --metrics gpu__time_duration.sum` (one pass, no kernel replay -- the peer-memory flags of lpmb_peer.cu must not be replayed),
profiling only between cudaProfilerStart/Stop around a few Newton iterations of the bench workload.
    python scripts/ncu_rank0_slabs.py [world=2] [n=216] [out=gpurun_out/slab_rank0_launches.csv]"""
import ctypes, importlib, os, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def child(rank, world, n, d):
    import bench
    lpm = importlib.import_module("lpm-c_b200")
    partition = importlib.import_module("lpm-c_b200.partition")
    uid_file = Path(d) / "uid.bin"
    if rank == 0:
        (Path(d) / "uid.tmp").write_bytes(lpm.Context.dist_unique_id())
        os.replace(Path(d) / "uid.tmp", uid_file)
    else:
        t0 = time.time()
        while not uid_file.exists():
            if time.time() - t0 > 120:
                raise SystemExit("no unique id from rank 0")
            time.sleep(0.05)
    uid = uid_file.read_bytes()
    slab = partition.make_slab(n, n * n, rank, world)
    c, info = bench.build_workload(lpm, n, rank, slab=slab, unique_id=uid)
    for _ in range(2):
        bench.one_step(c)
    c.synchronize()
    rt = None
    for name in ("libcudart.so", "libcudart.so.12"):
        try:
            rt = ctypes.CDLL(name); break
        except OSError:
            pass
    if rank == 0 and rt:
        rt.cudaProfilerStart()
    t0 = time.perf_counter()
    it, nr = bench.one_step(c)
    c.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0 and rt:
        rt.cudaProfilerStop()
    if rank == 0:
        print(f"world={world} n={n}: one Newton iteration {dt*1e3:.1f} ms (under ncu on rank 0), {it} CG iterations, comm mode {c.dist_mode()}", flush=True)
    c.close()


def main():
    a = sys.argv
    if "--rank" in a:
        child(int(a[a.index("--rank") + 1]), int(a[a.index("--world") + 1]), int(a[a.index("--n") + 1]), a[a.index("--dir") + 1])
        return
    world = int(a[1]) if len(a) > 1 else 2
    n = int(a[2]) if len(a) > 2 else 216
    out = a[3] if len(a) > 3 else "gpurun_out/slab_rank0_launches.csv"
    with tempfile.TemporaryDirectory() as d:
        procs = []
        for r in range(world):
            cmd = [sys.executable, __file__, "--rank", str(r), "--world", str(world), "--n", str(n), "--dir", d]
            if r == 0:
                cmd = ["ncu", "--profile-from-start", "off", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", out] + cmd
            procs.append(subprocess.Popen(cmd))
        rcs = [p.wait(timeout=600) for p in procs]
        print("exit codes", rcs)


if __name__ == "__main__":
    main()
