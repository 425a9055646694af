"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the package
never routes through the oracle, host-side logic (bandwidth mapping, state-dict keys, loud
failure without CUDA), and the world_size-2 EMA exchange over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from academicodec_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from academicodec_b200 import build
        build.build()
    return _lib.LIB_PATH


def test_cabi_exports_every_declared_symbol(lib_path):
    header = open(os.path.join(ROOT, "include", "acq_b200.h")).read()
    declared = set(re.findall(r"\b(acq_[a-z0-9_]+)\s*\(", header))
    declared.discard("acq_pipeline")
    from academicodec_b200 import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    lib.acq_version.restype = ctypes.c_int
    assert lib.acq_version() == 100


def test_argument_validation_without_gpu(lib_path):
    """Pure host-side validation paths return ACQ_EINVAL before touching the device."""
    from academicodec_b200 import _lib
    lib = _lib.load()
    tab = (ctypes.c_void_p * 1)(None)
    rc = lib.acq_rvq_search(None, ctypes.cast(tab, ctypes.POINTER(ctypes.c_void_p)), None, None, None,
                            1, 1, 1024, 128, 1, 1, 0, 0, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.acq_last_error()
    rc = lib.acq_vq_decode(None, 1, 1, None, 1, 1, 1024, 128, 1, 1, None, None, None)
    assert rc == -1


def test_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "academicodec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
    # and importing the package must not pull the oracle in
    code = "import sys; import academicodec_b200.quantization, academicodec_b200.grvq; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], cwd=ROOT, check=True)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from academicodec_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_cpu_tensor_is_rejected(lib_path):
    from academicodec_b200.quantization import ResidualVectorQuantizer
    q = ResidualVectorQuantizer(dimension=16, n_q=2, bins=32, kmeans_init=False).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        q.encode(torch.randn(1, 16, 5), 100)
    with pytest.raises(RuntimeError, match="no CPU path"):
        q.decode(torch.zeros(2, 1, 5, dtype=torch.int64))


def test_rvq_surface_matches_reference_contract():
    from academicodec_b200.quantization import QuantizedResult, ResidualVectorQuantizer  # noqa: F401
    q = ResidualVectorQuantizer(dimension=128, n_q=8, bins=1024)
    keys = list(q.state_dict().keys())
    want = [f"vq.layers.{i}._codebook.{k}" for i in range(8)
            for k in ("inited", "cluster_size", "embed", "embed_avg")]
    assert keys == want
    assert sum(p.numel() for p in q.parameters()) == 0          # buffers only
    assert float(q.vq.layers[0]._codebook.inited) == 0.0        # kmeans_init=True default
    assert float(q.vq.layers[0]._codebook.embed.abs().sum()) == 0.0
    assert tuple(q.vq.layers[3].codebook.shape) == (1024, 128)
    # bandwidth -> n_q (vq.py:88-101): 1024 bins @ 100 fps = 1 kbps per quantizer
    assert q.get_bandwidth_per_quantizer(100) == 1.0
    assert q.get_num_quantizers_for_bandwidth(100, 6.0) == 6
    assert q.get_num_quantizers_for_bandwidth(100, 0.3) == 1
    assert q.get_num_quantizers_for_bandwidth(100, None) == 8
    assert q.get_num_quantizers_for_bandwidth(100, 0.0) == 8
    assert q.get_num_quantizers_for_bandwidth(750, 7.5) == 1    # cfg2: 750 fps, 7.5 kbps
    q2 = ResidualVectorQuantizer(dimension=8, n_q=2, bins=16, kmeans_init=False)
    assert float(q2.vq.layers[0]._codebook.inited) == 1.0


def test_grvq_surface_matches_reference_contract():
    import types
    from academicodec_b200.grvq import Quantizer
    h = types.SimpleNamespace(n_code_groups=2, n_codes=1024, codebook_loss_lambda=1.0,
                              commitment_loss_lambda=0.25)
    q = Quantizer(h)
    keys = list(q.state_dict().keys())
    assert keys == ["quantizer_modules.0.embedding.weight", "quantizer_modules.1.embedding.weight",
                    "quantizer_modules2.0.embedding.weight", "quantizer_modules2.1.embedding.weight"]
    w = q.quantizer_modules[0].embedding.weight
    assert tuple(w.shape) == (1024, 256) and float(w.abs().max()) <= 1.0 / 1024
    assert q.residul_layer == 2 and q.n_code_groups == 2
    with pytest.raises(AssertionError):
        Quantizer(types.SimpleNamespace(n_code_groups=3, n_codes=8, codebook_loss_lambda=1.0,
                                        commitment_loss_lambda=0.25))


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ACQ_ROOT"])
from academicodec_b200 import ops
from academicodec_b200.quantization import core_vq
from oracle import rvq_oracle
from tests import cases

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
case = cases.RVQ_CASES["odd_dims"]
x, cb = cases.rvq_inputs(case)
s, k, d = cb.shape

# CPU stand-ins for the two kernels (test-only): this test covers the host-side exchange --
# every rank contributes its shard's statistics, one all-reduce, identical apply everywhere.
def fake_stats(x_bdt, codes, embeds, flags=0):
    n = len(embeds)
    sums = torch.zeros(n, k, d); counts = torch.zeros(n, k)
    r = x_bdt.transpose(1, 2).reshape(-1, d).clone()
    for i in range(n):
        c = codes[i].reshape(-1)
        sums[i].index_add_(0, c, r); counts[i] += torch.bincount(c, minlength=k).float()
        q = embeds[i][c]; r = r - (r + (q - r))
    return torch.cat([sums.reshape(-1), counts.reshape(-1)])
def fake_apply(stats, embed, embed_avg, cluster_size, decay, eps):
    n = len(embed)
    sums = stats[: n * k * d].view(n, k, d); counts = stats[n * k * d:].view(n, k)
    for i in range(n):
        if float(counts[i].sum()) == 0.0:
            continue                      # stage no rank used (acq_ema_apply's rule)
        cluster_size[i].mul_(decay).add_(counts[i], alpha=1 - decay)
        embed_avg[i].mul_(decay).add_(sums[i], alpha=1 - decay)
        n = cluster_size[i].sum()
        sm = (cluster_size[i] + eps) / (n + k * eps) * n
        embed[i].copy_(embed_avg[i] / sm.unsqueeze(1))
ops.ema_stats, ops.ema_apply = fake_stats, fake_apply

books = [core_vq.EuclideanCodebook(d, k) for _ in range(s)]
for i, b_ in enumerate(books):
    b_.embed.copy_(cb[i]); b_.embed_avg.copy_(cb[i])
# global batch = concat over ranks; the oracle runs it in one process
xs = [x, x.flip(0) * 0.5]
states = rvq_oracle.make_states(cb)
_, codes_all, _ = rvq_oracle.rvq_forward(torch.cat(xs, 0), states, None, training=True)
bsz = x.shape[0]
my_codes = codes_all[:, rank * bsz:(rank + 1) * bsz].reshape(s, -1)
if os.environ.get("ACQ_RAGGED"):
    # ranks disagree on the number of stages (per-process bandwidth draw, net3.py:41): rank 1 stops one
    # stage early.  The collective must still match in size (whole stack) and every rank must apply the
    # update of every stage some rank used.
    used = s if rank == 0 else s - 1
    core_vq.ema_update_(books[:used], xs[rank], my_codes[:used], flags=ops.ACQ_STE, all_codebooks=books)
    solo = rvq_oracle.make_states(cb)          # last stage: only rank 0's batch contributed
    rvq_oracle.rvq_forward(xs[0], solo, None, training=True)
    states[s - 1] = solo[s - 1]
else:
    core_vq.ema_update_(books, xs[rank], my_codes, flags=ops.ACQ_STE)
for i in range(s):
    torch.testing.assert_close(books[i].cluster_size, states[i]["cluster_size"], rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(books[i].embed, states[i]["embed"], rtol=1e-5, atol=1e-6)
# replicas identical without any broadcast
flat = torch.cat([b_.embed.reshape(-1) for b_ in books])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert torch.equal(gathered[0], gathered[1])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("ragged", [False, True], ids=["same_n_q", "ranks_pick_different_n_q"])
def test_ema_exchange_world_size_2_gloo(tmp_path, ragged):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + (os.getpid() % 2000) + (7 if ragged else 0)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), ACQ_ROOT=ROOT, OMP_NUM_THREADS="1")
        if ragged:
            env["ACQ_RAGGED"] = "1"
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) works without a GPU and
    prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-clips", "2", "--workload", "cfg1_enc24k_240d_rvq"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "rvq_encode_decode_frames_per_sec"
    assert line["unit"] == "frames/s" and line["value"] > 0 and line["higher_is_better"] is True
    # "reference" when baseline/_ref (the unmodified reference package) is installed, the oracle port otherwise
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == "cfg1_enc24k_240d_rvq"
    # the oracle port stays selectable, and the training workload has its own metric name
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-clips", "2", "--ref-port", "--workload", "cfg5_rvq_ema_train"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["kind"] == "port" and line["metric"] == "rvq_ema_train_frames_per_sec"
