"""What the host side of the box can move: pinned host memory <-> device copies on N GPUs at once, no kernels.

The end-to-end number of bench.py (`e2e`) is bounded by these copies (3.2 GB up and 2.4 GB down per step at config 5),
so this probe measures the ceiling it should be compared with: per-GPU and aggregate GB/s for host->device alone,
device->host alone and both directions at once, with N = 1, 2, 4, 8 processes (one per GPU, each bound to the CPU cores
NVML reports as local to its GPU before it allocates its pinned buffers, like bench.py does).

usage: python tools/host_link_probe.py [max_gpus]      (writes gpurun_out/host_link_probe.json)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BYTES = 1 << 30      # per direction and GPU per repetition
REPS = 6


def worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from __graft_entry__ import load_package
    scb = load_package()
    bound = scb.bind_host_to_device(rank)
    up_h = torch.empty(BYTES, dtype=torch.uint8).pin_memory()
    dn_h = torch.empty(BYTES, dtype=torch.uint8).pin_memory()
    up_h.fill_(1)
    up_d = torch.empty(BYTES, dtype=torch.uint8, device="cuda")
    dn_d = torch.ones(BYTES, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def run(do_up, do_dn):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(REPS):
            if do_up:
                with torch.cuda.stream(s_up):
                    up_d.copy_(up_h, non_blocking=True)
            if do_dn:
                with torch.cuda.stream(s_dn):
                    dn_h.copy_(dn_d, non_blocking=True)
        torch.cuda.synchronize()
        dist.barrier()
        return time.perf_counter() - t0

    res = {}
    for name, (u, d) in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True))):
        run(u, d)
        t = min(run(u, d) for _ in range(3))
        res[name] = {"seconds": t, "GBps_per_gpu": (int(u) + int(d)) * REPS * BYTES / t / 1e9}
    if rank == 0:
        res["cores_bound_rank0"] = len(bound) if bound else None
        with open(os.path.join(out_dir, "link_%d.json" % world), "w") as f:
            json.dump(res, f)
    dist.destroy_process_group()


def main():
    import torch
    import torch.multiprocessing as mp
    ngpu = torch.cuda.device_count()
    cap = int(sys.argv[1]) if len(sys.argv) > 1 else ngpu
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    report = {"bytes_per_copy": BYTES, "reps": REPS, "host_cpus": os.cpu_count(), "gpus_on_box": ngpu, "runs": {}}
    for world in (1, 2, 4, 8):
        if world > min(ngpu, cap):
            break
        mp.spawn(worker, args=(world, 29700 + world, out_dir), nprocs=world, join=True)
        with open(os.path.join(out_dir, "link_%d.json" % world)) as f:
            r = json.load(f)
        os.remove(os.path.join(out_dir, "link_%d.json" % world))
        for k in ("h2d", "d2h", "both"):
            r[k]["GBps_aggregate"] = r[k]["GBps_per_gpu"] * world
        # one config-5 step moves 3.2 GB up and 2.4 GB down in total (shared by the ranks): the floor the host side sets
        up = 3.2e9 / (r["h2d"]["GBps_aggregate"] * 1e9)
        dn = 2.4e9 / (r["d2h"]["GBps_aggregate"] * 1e9)
        both = 5.6e9 / (r["both"]["GBps_aggregate"] * 1e9)
        r["config5_step_floor_ms"] = {"blocking (upload, then download)": 1e3 * (up + dn),
                                      "pipelined (both directions busy)": 1e3 * max(both, up, dn)}
        report["runs"]["%d_gpus" % world] = r
    print(json.dumps(report, indent=1))
    with open(os.path.join(out_dir, "host_link_probe.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
