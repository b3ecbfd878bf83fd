#!/usr/bin/env python
"""Host-link ceiling of the box: every rank runs concurrent page-locked H2D + D2H copies at the same time
(rr_host_link_probe) and rank 0 prints one JSON line with the per-rank and aggregate GB/s.

  python tools/pcie_ceiling.py                                     # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/pcie_ceiling.py [--no-bind] [--wc] [--direction 3]

This is the denominator of bench.py's end-to-end numbers (e2e.host_link_gbs): frames are staged through
page-locked host buffers, so N ranks cannot move more bytes than this.  --no-bind leaves NUMA placement to
the OS (RAIN_B200_NUMA_BIND=0), the default binds each rank to its GPU's node first (dist.bind_to_gpu_numa).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--no-bind", action="store_true")
    ap.add_argument("--wc", action="store_true", help="write-combined input buffer")
    ap.add_argument("--direction", type=int, default=3, help="1 h2d, 2 d2h, 3 both at once")
    args = ap.parse_args()
    from rain_rendering_b200 import api, dist as rdist
    rank, world, local = rdist.env_rank()
    bind = {"bound": False, "reason": "--no-bind"} if args.no_bind else rdist.bind_to_gpu_numa(local)
    import torch
    import torch.distributed as dist
    rdist.init_process_group("nccl")
    torch.cuda.set_device(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    h2d, d2h = api.host_link_probe(local, args.mb, args.seconds, args.wc, args.direction)
    t = torch.tensor([h2d, d2h], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        binds = [None] * world
        dist.all_gather_object(binds, bind)
    else:
        allv, binds = [t], [bind]
    if rank == 0:
        per = [[float(v[0]), float(v[1])] for v in allv]
        print(json.dumps({"tool": "pcie_ceiling", "n_gpus": world, "mb": args.mb, "direction": args.direction, "write_combined": args.wc,
                          "numa_bind": binds, "per_rank_h2d_d2h_gbs": per, "h2d_gbs": sum(p[0] for p in per), "d2h_gbs": sum(p[1] for p in per),
                          "total_gbs": sum(p[0] + p[1] for p in per)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
