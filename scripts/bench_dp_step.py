"""Data-parallel flow-training step (C4 and C5 flow shapes, global batch 16384) on the ranks of one box:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_dp_step.py
FLOWMC_DP_PEER=0 selects NCCL all-reduce + flowmc_clip_adamw instead of the fused peer-memory optimiser kernel;
FLOWMC_TC_SPLIT=0 switches the feature-split kernels off.  Rank 0 prints one JSON line per shape."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import flow_train_dp  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    for tag, d, L in (("C4", 32, 10), ("C5", 64, 8)):
        out = flow_train_dp(dev, rank, world, d=d, n_layers=L, tag=tag)
        out["env"] = {k: os.environ.get(k) for k in ("FLOWMC_DP_PEER", "FLOWMC_TC_SPLIT") if os.environ.get(k)}
        if rank == 0:
            print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
