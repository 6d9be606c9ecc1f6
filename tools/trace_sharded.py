"""Phase times of the NCCL variant of the sharded tier (development tool; torchrun, config-4 shape at reduced rows)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scone_b200 as sb  # noqa: E402
from scone_b200 import sharded  # noqa: E402
from scone_b200.utils import synthetic as S  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rows_per_gpu, D, V, B, L = 2_000_000, 4096, 128_000, 256, 2048
N = rows_per_gpu * world
toks, lens, longest = S.make_vocab_device(N, 5, V, seed=0, device=dev, return_longest=True)
index = sb.FGramIndex(toks, lens)
table = sb.CacheTable(sharded.shard_rows(N, rank, world), D, "fp16", device=dev)
base = S.make_base_device(V, D, torch.bfloat16, seed=3, device=dev)
q = S.make_stream_device(toks, lens, B, L, V, seed=100 + rank, p_plant=1.0, pick_ids=longest)
out = torch.empty((B, L, D), dtype=torch.bfloat16, device=dev)
for M in (1, 4):
    cache = sharded.ShardedEmbeddingCache(sharded.CudaOps(index, table, base), micro_batches=M)
    os.environ.pop("SCONE_SHARDED_TRACE", None)
    for _ in range(3):
        cache.lookup(q, out=out)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        cache.lookup(q, out=out)
    e1.record()
    torch.cuda.synchronize()
    os.environ["SCONE_SHARDED_TRACE"] = "1"
    cache.lookup(q, out=out)
    if rank == 0:
        print(f"micro {M}: {e0.elapsed_time(e1) / 5:.3f} ms per lookup untraced; serialised phases (ms):", flush=True)
        print("   " + ", ".join(f"{n} {t:.3f}" for n, t in cache.trace_log), flush=True)
        print(f"   sum {sum(t for _, t in cache.trace_log):.3f}", flush=True)
dist.destroy_process_group()
