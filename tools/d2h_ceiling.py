#!/usr/bin/env python
"""Host-ingest ceiling of the box: bare cudaMemcpyAsync device -> page-locked host copies of 33 MB buffers (one 4K 32-bit picture,
what bench.py's e2e leg reads back per frame) from N GPUs at once, nothing else running.

    python tools/d2h_ceiling.py                                   one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_ceiling.py [--no-numa]

Every rank keeps 4 copies in flight on 4 streams for --seconds, the aggregate is the sum over ranks of bytes / the slowest rank's
time.  Prints one JSON line (rank 0).  bench.py's e2e value x 33.2 MB should be compared with THIS number, not with PCIe's nominal
rate: beyond a few GPUs the host memory system, not the links, bounds the read-back."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--no-numa", action="store_true", help="do not pin the rank to its GPU's NUMA node before allocating")
    args = ap.parse_args()
    import torch

    import bench

    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    numa = {"pinned": False} if args.no_numa else bench.pin_to_gpu_numa_node(local_rank)
    n_buf, nbytes = 4, bench.WIDTH * bench.HEIGHT * 4
    dev = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(n_buf)]
    host = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(n_buf)]
    for h in host:
        h.zero_()  # first touch on this rank's node
    streams = [torch.cuda.Stream() for _ in range(n_buf)]

    def run(seconds):
        copies = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for i in range(n_buf):
                with torch.cuda.stream(streams[i]):
                    host[i].copy_(dev[i], non_blocking=True)
            copies += n_buf
            if copies % (8 * n_buf) == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        return copies, time.perf_counter() - t0

    run(0.3)
    if dist is not None:
        dist.barrier()
    copies, dt = run(args.seconds)
    gbs = copies * nbytes / dt / 1e9
    if dist is not None:
        t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        total = float(t.item())
        dist.barrier()
        dist.destroy_process_group()
    else:
        total = gbs
    if rank == 0:
        print(json.dumps({
            "what": "bare cudaMemcpyAsync D2H into page-locked host memory, 33.2 MB buffers, 4 in flight per GPU",
            "n_gpus": world, "aggregate_gb_per_s": total, "rank0_gb_per_s": gbs, "pictures_per_s": total * 1e9 / nbytes, "numa": numa,
            "host_cores": os.cpu_count(),
        }))


if __name__ == "__main__":
    main()
