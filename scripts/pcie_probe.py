"""Host <-> device copy bandwidth per GPU, alone and with every GPU of the box copying at once (what limits the
end-to-end number of bench.py at N > 1: b2sim_task_step_host moves 8 B in and 41 B out per env-step).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 scripts/pcie_probe.py

Every rank allocates 256 MB pinned buffers after binding to its GPU's NUMA node (bench.bind_near_gpu), then times
cudaMemcpyAsync D2H, H2D and both directions together: first rank by rank while the others wait (solo), then all ranks
between the same barriers (concurrent). Rank 0 prints one JSON line.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    binding = bench.bind_near_gpu(local)
    nbytes = 256 << 20
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(kind, reps=8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return reps * nbytes * (2 if kind == "both" else 1) / dt / 1e9

    for kind in ("d2h", "h2d"):
        run(kind, 2)
    res = {}
    for kind in ("d2h", "h2d", "both"):
        solo = torch.zeros(world, dtype=torch.float64, device="cuda")
        for r in range(world):
            barrier()
            if r == rank:
                solo[r] = run(kind)
            barrier()
        barrier()
        conc = torch.zeros(world, dtype=torch.float64, device="cuda")
        conc[rank] = run(kind)
        barrier()
        if world > 1:
            dist.all_reduce(solo)
            dist.all_reduce(conc)
        res[kind] = {"solo_gbs_per_gpu": [round(v, 1) for v in solo.tolist()],
                     "concurrent_gbs_per_gpu": [round(v, 1) for v in conc.tolist()],
                     "concurrent_total_gbs": round(float(conc.sum().item()), 1)}
    if rank == 0:
        print(json.dumps({"probe": "pinned host <-> device copies, 256 MB, GB/s", "n_gpus": world, "binding_rank0": binding,
                          "cpus_visible": len(os.sched_getaffinity(0)), **res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
