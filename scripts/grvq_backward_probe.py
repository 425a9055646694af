"""GRVQ training step (forward + backward of the HiFi-Codec Quantizer) with the backward kernel vs the eager
expression of the same gradients.   python scripts/grvq_backward_probe.py [B] [T]"""
import os, sys, torch
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from academicodec_b200 import grvq
b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
h = SimpleNamespace(n_code_groups=2, n_codes=1024, codebook_loss_lambda=1.0, commitment_loss_lambda=0.25)
q = grvq.Quantizer(h).to(dev)
x = torch.randn(b, 512, t, device=dev)


def step():
    xg = x.clone().requires_grad_(True)
    for w in q._weights():
        w.grad = None
    qo, loss, _ = q(xg)
    ((qo * qo).mean() + loss).backward()


for eager in (True, False, True, False):
    grvq._EAGER_BACKWARD = eager
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"B={b} T={t} backward={'eager' if eager else 'kernel'}: {e0.elapsed_time(e1) / 10:.3f} ms per forward+backward", flush=True)
