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
# the exchange itself: this package's kernel over NVLink peer memory (or NCCL when switched off)
from academicodec_b200.quantization.distrib import PeerExchange
exch = PeerExchange.create(1024 * 129 * 3 + 4, dev)
want_peer = os.environ.get("ACQ_PEER_REDUCE", "1") != "0"
if want_peer and os.environ.get("ACQ_REQUIRE_PEER") == "1":
    assert exch is not None, "symmetric memory unavailable"
if exch is not None:
    n = exch.numel
    g = torch.Generator(device="cpu").manual_seed(5)
    parts = [torch.randn(n, generator=g) * (1.0 + r) for r in range(world)]
    for rep in range(3):
        exch.buf.copy_(parts[rank].to(dev))
        got = exch.all_reduce_().clone()
        torch.cuda.synchronize()
        acc = parts[0].clone()
        for r in range(1, world):
            acc = acc + parts[r]
        if exch.mode == "p2p":
            assert torch.equal(got.cpu(), acc), "peer all-reduce differs from the rank-ordered fp32 sum"
        else:
            torch.testing.assert_close(got.cpu(), acc, rtol=1e-5, atol=1e-5)
        gl = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gl, got)
        assert all(torch.equal(gl[0], t) for t in gl[1:]), "ranks hold different sums"
    print("rank", rank, "peer exchange", exch.mode, "ok")
else:
    print("rank", rank, "peer exchange off: nccl")
# encode / decode: no collective, each rank its own clips
q.eval()
c2 = q.encode(shards[rank].to(dev), case["frame_rate"])
assert torch.equal(q.decode(c2), q(shards[rank].to(dev), case["frame_rate"])[0])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                    reason="needs >= 2 CUDA devices")
@pytest.mark.parametrize("mode", ["peer", "peer_p2p", "nccl"])
def test_ema_allreduce_multi_gpu(tmp_path, mode):
    """EMA statistics exchange at every available world size (2 .. 8): the module forward on sharded batches equals
    the oracle on the concatenated batch, replicas stay bit-identical, and the exchange kernel itself
    (NVLS multimem / P2P / NCCL) returns the same sums on every rank."""
    world = min(torch.cuda.device_count(), 8)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29700 + (os.getpid() % 2000) + {"peer": 0, "peer_p2p": 1, "nccl": 2}[mode]
    extra = {"peer": {"ACQ_PEER_REDUCE": "1"}, "peer_p2p": {"ACQ_PEER_REDUCE": "1", "ACQ_PEER_MULTICAST": "0"},
             "nccl": {"ACQ_PEER_REDUCE": "0"}}[mode]
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ACQ_ROOT=ROOT, **extra)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    print(outs[0][-400:])
