"""Development tool (torchrun, 2 GPUs): calibrates the NVLink peer path -- cudaMemcpyPeer-style copies (torch) at the exchange's sizes."""
import os, time, json
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (14, 28, 256):
    n = mb * (1 << 20) // 2
    a = torch.ones(n, dtype=torch.float16, device="cuda")
    b = torch.empty_like(a)
    # all-to-all style exchange through NCCL send/recv (NVLink), bidirectional
    for it in range(3):
        ops = [dist.P2POp(dist.isend, a, (rank + 1) % world), dist.P2POp(dist.irecv, b, (rank - 1) % world)]
        for r in dist.batch_isend_irecv(ops): r.wait()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(20):
        ops = [dist.P2POp(dist.isend, a, (rank + 1) % world), dist.P2POp(dist.irecv, b, (rank - 1) % world)]
        for r in dist.batch_isend_irecv(ops): r.wait()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    if rank == 0: print(json.dumps({"nccl_sendrecv_MB": mb, "us": round(t * 1e3, 1), "GBps_per_direction": round(mb * 1.048576 / t, 1)}))
    # all-reduce of the same size for reference
    for it in range(3): dist.all_reduce(a)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for it in range(20): dist.all_reduce(a)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    if rank == 0: print(json.dumps({"nccl_allreduce_MB": mb, "us": round(t * 1e3, 1)}))
dist.destroy_process_group()
