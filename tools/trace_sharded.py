"""Phase table of a sharded control step (MPOPIS_TRACE=1 prints CUDA-event phase times per rank to stderr).
    torchrun --nproc-per-node N tools/trace_sharded.py <samples per GPU>"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist
from bench import make_engine
from mpopis_b200 import _lib, sharding

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
K = int(sys.argv[1]) * world
env, eng = make_engine(_lib.product(), K, rank, world, lr)
if world > 1:
    print("transport", sharding.connect(eng, dist, rank, world), file=sys.stderr)
U, st = np.zeros(eng.cs), env.state.copy()
for i in range(4):
    ctrl, U, its = eng.plan(st, i, U)
if rank == 0:
    print("ok", K, ctrl, its)
if world > 1:
    dist.destroy_process_group()
