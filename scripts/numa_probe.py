"""Development tool (multi-GPU box): host side of the e2e path at N > 1.  Every rank copies the bench step's record buffers (41.5 MB H2D,
24.9 MB D2H, pinned) at the same time as all other ranks; variants: CPU affinity left alone / bound to the GPU-local CPUs (NVML) BEFORE the
pinned buffers are allocated (first touch decides the NUMA node), one direction at a time, both.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/numa_probe.py"""
import json, os, subprocess, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 2073600
if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["bash", "-c", "for d in /sys/bus/pci/devices/*; do if [ -e $d/numa_node ] && grep -qi 0x10de $d/vendor; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done"],
                ["bash", "-c", "cat /sys/devices/system/node/node*/meminfo | grep MemTotal; nproc"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            print("$", " ".join(cmd)); print("\n".join(l for l in out.splitlines() if not cmd[0] == "lscpu" or "NUMA" in l or "Model name" in l or "Socket" in l or "CPU(s):" in l))
        except Exception as e:
            print("failed", cmd, e)
    sys.stdout.flush()

def local_cpus():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    n = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
    cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
    return cpus

def run(tag, bind):
    before = sorted(os.sched_getaffinity(0))
    cpus = None
    if bind:
        try:
            cpus = local_cpus()
            if cpus: os.sched_setaffinity(0, cpus)
        except Exception as e:
            cpus = f"failed: {e}"
    pin_in = torch.empty((N, 5), dtype=torch.float32).pin_memory(); pin_in.fill_(1.0)
    pin_out = torch.empty((N, 3), dtype=torch.float32).pin_memory(); pin_out.fill_(0.0)
    d_in = torch.empty((N, 5), dtype=torch.float32, device="cuda"); d_out = torch.zeros((N, 3), dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for it in range(2):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(10):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1): d_in.copy_(pin_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2): pin_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 100
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = round(float(t.item()), 3)
    os.sched_setaffinity(0, before)
    allc = [None] * world
    dist.all_gather_object(allc, (local, (cpus[:4] + ["..."] + cpus[-2:]) if isinstance(cpus, list) and len(cpus) > 6 else cpus, len(before)))
    if rank == 0:
        print(json.dumps({"tag": tag, "world": world, "ms_per_step_max_over_ranks": res, "cpus": allc}))
        sys.stdout.flush()

run("default affinity", False)
run("GPU-local CPUs (NVML) before pinning", True)
run("default affinity again", False)
dist.destroy_process_group()
