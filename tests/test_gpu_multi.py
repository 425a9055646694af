"""Multi-GPU checks (need >= 2 CUDA devices; run with `gpurun --gpus 2`):
the EMA exchange over NCCL equals the reference on the concatenated global batch and keeps the
replicas bit-identical; encode/decode shard clips with no collective."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ACQ_ROOT"])
from academicodec_b200.quantization import ResidualVectorQuantizer
from oracle import rvq_oracle
from tests import cases

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
case = cases.RVQ_CASES["cfg1_small"]
x, cb = cases.rvq_inputs(case)
q = ResidualVectorQuantizer(dimension=case["D"], n_q=case["n_q"], bins=case["bins"], kmeans_init=False)
for i, layer in enumerate(q.vq.layers):
    layer._codebook.embed.data.copy_(cb[i]); layer._codebook.embed_avg.data.copy_(cb[i])
q = q.to(dev).train()
shards = [x, x.flip(0) * 0.5][:world] if world == 2 else [x * (1.0 - 0.1 * r) for r in range(world)]
qz, codes, bw, pen = q(shards[rank].to(dev), case["frame_rate"])
torch.cuda.synchronize()
# oracle: ONE process, concatenated global batch
states = rvq_oracle.make_states(cb)
oq, ocodes, _, _ = rvq_oracle.quantizer_forward(torch.cat(shards, 0), states, case["bins"],
                                                 case["frame_rate"], None, training=True)
b = x.shape[0]
assert torch.equal(codes.cpu(), ocodes[:, rank * b:(rank + 1) * b]), "codes differ from the oracle shard"
for i, layer in enumerate(q.vq.layers):
    c = layer._codebook
    torch.testing.assert_close(c.cluster_size.cpu(), states[i]["cluster_size"], rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(c.embed.cpu(), states[i]["embed"], rtol=1e-5, atol=1e-6)
flat = torch.cat([l._codebook.embed.reshape(-1) for l in q.vq.layers])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert all(torch.equal(gathered[0], g) for g in gathered[1:]), "replicas diverged"
# encode / decode: no collective, each rank its own clips
q.eval()
c2 = q.encode(shards[rank].to(dev), case["frame_rate"])
assert torch.equal(q.decode(c2), q(shards[rank].to(dev), case["frame_rate"])[0])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                    reason="needs >= 2 CUDA devices")
def test_ema_allreduce_nccl(tmp_path):
    world = min(torch.cuda.device_count(), 2)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29700 + (os.getpid() % 2000)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ACQ_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
