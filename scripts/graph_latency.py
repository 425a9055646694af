"""Small-batch latency of encode / decode: eager module calls vs CUDA-graph replay (wall clock per call,
synchronised, i.e. what a streaming caller sees)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import synth
from academicodec_b200.graphs import GraphedCodec
from academicodec_b200.quantization import ResidualVectorQuantizer
dev = torch.device("cuda:0")
for (d, n_q, b, t) in [(128, 8, 16, 100), (128, 8, 1, 100), (512, 12, 16, 100), (512, 1, 1, 750)]:
    q = ResidualVectorQuantizer(dimension=d, n_q=n_q, bins=1024, kmeans_init=False)
    cb = torch.from_numpy(synth.rvq_codebooks(n_q, 1024, d, 4321, "decay"))
    for i, layer in enumerate(q.vq.layers):
        layer._codebook.embed.data.copy_(cb[i])
    q = q.to(dev).eval()
    x = torch.from_numpy(synth.latents(b, d, t, 1)).to(dev)
    g = GraphedCodec(q, x, 100)
    def wall(fn, n=200):
        for _ in range(10): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn(); torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e6
    codes = q.encode(x, 100)
    print(f"D={d} n_q={n_q} frames={b*t}: encode eager {wall(lambda: q.encode(x, 100)):7.1f} us  graph {wall(lambda: g.encode(x)):7.1f} us"
          f" | decode eager {wall(lambda: q.decode(codes)):7.1f} us  graph {wall(lambda: g.decode(codes)):7.1f} us")
