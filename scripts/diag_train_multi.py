"""Training-forward step time under torchrun, with and without a host sync per step, peer exchange vs NCCL."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import synth
from academicodec_b200.quantization import ResidualVectorQuantizer
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d, n_q, b = 512, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 16
q = ResidualVectorQuantizer(dimension=d, n_q=n_q, bins=1024, kmeans_init=False)
cb = torch.from_numpy(synth.rvq_codebooks(n_q, 1024, d, 4321, "decay"))
for i, layer in enumerate(q.vq.layers):
    layer._codebook.embed.data.copy_(cb[i]); layer._codebook.embed_avg.data.copy_(cb[i])
q = q.to(dev).train()
x = torch.from_numpy(synth.latents(b, d, 100, 1234 + rank)).to(dev)
for _ in range(5):
    q(x, 100)
torch.cuda.synchronize(); dist.barrier()
for mode in ("sync_each_step", "async", "async"):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        q(x, 100)
        if mode == "sync_each_step":
            torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"peer={os.environ.get('ACQ_PEER_REDUCE', '1')} b={b} {mode}: host {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/step, "
              f"events {e0.elapsed_time(e1) / 20:.3f} ms/step", flush=True)
    dist.barrier()
dist.destroy_process_group()
