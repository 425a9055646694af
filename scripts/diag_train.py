import os, sys, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import synth
from academicodec_b200.quantization import ResidualVectorQuantizer
dev = torch.device("cuda:0")
d, n_q, b = 512, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 640
q = ResidualVectorQuantizer(dimension=d, n_q=n_q, bins=1024, kmeans_init=False)
cb = torch.from_numpy(synth.rvq_codebooks(n_q, 1024, d, 4321, "decay"))
for i, layer in enumerate(q.vq.layers):
    layer._codebook.embed.data.copy_(cb[i]); layer._codebook.embed_avg.data.copy_(cb[i])
q = q.to(dev).train()
x = torch.from_numpy(synth.latents(b, d, 100, 1234)).to(dev)
for _ in range(3):
    q(x, 100)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    q(x, 100)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        q(x, 100)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
