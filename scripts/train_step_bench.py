"""cfg5: train-mode forward (search + straight-through outputs + EMA k-means update) step time.
Single process or under torchrun (NCCL all-reduce of the statistics).  SURVEY.md 8d cfg5."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from academicodec_b200 import synth
from academicodec_b200.quantization import ResidualVectorQuantizer

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

CONFIGS = [(128, 8, 16), (128, 8, 640), (512, 12, 16), (512, 12, 640)]
if len(sys.argv) > 1:            # e.g. `train_step_bench.py 3` = only the fourth config (for ncu launch lists)
    CONFIGS = [CONFIGS[int(sys.argv[1])]]
for (d, n_q, b) in CONFIGS:
    q = ResidualVectorQuantizer(dimension=d, n_q=n_q, bins=1024, kmeans_init=False)
    cb = torch.from_numpy(synth.rvq_codebooks(n_q, 1024, d, 4321, "decay"))
    for i, layer in enumerate(q.vq.layers):
        layer._codebook.embed.data.copy_(cb[i]); layer._codebook.embed_avg.data.copy_(cb[i])
    q = q.to(dev).train()
    x = torch.from_numpy(synth.latents(b, d, 100, 1234 + rank)).to(dev)
    for _ in range(3):
        q(x, 100)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        q(x, 100)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        payload = n_q * 1024 * (d + 1) * 4 / 1e6
        print(f"cfg5 world={world} D={d} n_q={n_q} frames/GPU={b*100}: {ms.item():.3f} ms/step "
              f"({world*b*100/ms.item()/1e3:.2f} M frames/s), all-reduce payload {payload:.1f} MB")
if world == 1 and len(sys.argv) == 1:
    # CPU arm of the same step (oracle port) for the smallest config
    from oracle import rvq_oracle
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    for (d, n_q, b) in [(128, 8, 16), (512, 12, 16)]:
        cb = torch.from_numpy(synth.rvq_codebooks(n_q, 1024, d, 4321, "decay"))
        states = rvq_oracle.make_states(cb)
        x = torch.from_numpy(synth.latents(b, d, 100, 1234))
        rvq_oracle.rvq_forward(x, states, None, training=True)
        t0 = time.perf_counter()
        for _ in range(3):
            rvq_oracle.rvq_forward(x, states, None, training=True)
        print(f"cfg5 CPU oracle port D={d} n_q={n_q} frames={b*100}: {(time.perf_counter()-t0)/3*1e3:.1f} ms/step "
              f"({torch.get_num_threads()} threads)")
if world > 1:
    dist.destroy_process_group()
