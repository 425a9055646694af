"""Pin the oracle: it must reproduce what the live reference returned (tests/golden/), exactly
for codes / decode / straight-through values (same ATen calls in the same order on the same
host class) and to 1e-6 for multi-threaded reductions.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import adjudicate, grvq_oracle, rvq_oracle
from tests import cases

ROW_STRIDE = 64


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)      # the fixtures were generated single-threaded
    yield
    torch.set_num_threads(n)


def _codes_equal_or_near(x, cb, ref, new, st=0, ste=False):
    """Exact equality is expected; if the host's sgemm differs from the generating host's, any
    disagreement must be an fp64 near-tie."""
    ref = np.asarray(ref).astype(np.int64)
    new = np.asarray(new).astype(np.int64)
    if np.array_equal(ref, new):
        return 0
    rep = adjudicate.compare_rvq_codes(x, cb, ref, new, st=st, straight_through=ste)
    assert rep["hard_mismatch"] == 0, rep
    return rep["near_tie"]


@pytest.mark.parametrize("name", list(cases.RVQ_CASES))
def test_rvq_encode_decode(golden, name):
    case = cases.RVQ_CASES[name]
    x, cb = cases.rvq_inputs(case)
    embeds = list(cb)
    codes = rvq_oracle.rvq_encode(x, embeds)
    near = _codes_equal_or_near(x, cb, golden[f"{name}/codes"], codes.numpy())
    if near == 0:
        dec = rvq_oracle.rvq_decode(codes, embeds)
        assert np.array_equal(dec.numpy(), golden[f"{name}/decode"])
    # decode of the REFERENCE codes must be bit-exact regardless
    ref_codes = torch.from_numpy(golden[f"{name}/codes"].astype(np.int64))
    assert np.array_equal(rvq_oracle.rvq_decode(ref_codes, embeds).numpy(), golden[f"{name}/decode"])
    # bandwidth -> n_q mapping and st slicing
    bw = float(golden[f"{name}/bw"])
    n_q = rvq_oracle.num_quantizers_for_bandwidth(case["n_q"], case["bins"], case["frame_rate"], bw)
    assert n_q == golden[f"{name}/codes_bw"].shape[0] == max(1, case["n_q"] // 2)
    _codes_equal_or_near(x, cb, golden[f"{name}/codes_bw"],
                         rvq_oracle.rvq_encode(x, embeds, n_q).numpy())
    if case["n_q"] >= 3:
        _codes_equal_or_near(x, cb, golden[f"{name}/codes_st2"],
                             rvq_oracle.rvq_encode(x, embeds, None, 2).numpy(), st=2)


@pytest.mark.parametrize("name", list(cases.RVQ_CASES))
def test_rvq_forward_eval(golden, name):
    case = cases.RVQ_CASES[name]
    x, cb = cases.rvq_inputs(case)
    states = rvq_oracle.make_states(cb)
    bw = float(golden[f"{name}/bw"])
    q, codes, bwt, pen = rvq_oracle.quantizer_forward(x, states, case["bins"], case["frame_rate"], bw)
    near = _codes_equal_or_near(x, cb, golden[f"{name}/fwd_eval_codes"], codes.numpy())
    if near == 0:
        assert np.array_equal(q.numpy(), golden[f"{name}/fwd_eval_quantized"])
    assert np.array_equal(bwt.numpy(), golden[f"{name}/fwd_eval_bw"])
    assert float(pen) == float(golden[f"{name}/fwd_eval_penalty"]) == 0.0


@pytest.mark.parametrize("name", list(cases.RVQ_CASES))
def test_rvq_forward_train_ema(golden, name):
    """Two train-mode steps; also pins 'dead-code expiry leaves no trace in the state': the
    fixtures ran with the reference default threshold 2, the oracle has no expiry at all."""
    case = cases.RVQ_CASES[name]
    x, cb = cases.rvq_inputs(case)
    states = rvq_oracle.make_states(cb)
    for step in range(2):
        xs = x if step == 0 else x.flip(0) * 0.5
        q, codes, bwt, pen = rvq_oracle.quantizer_forward(xs, states, case["bins"],
                                                          case["frame_rate"], None, training=True)
        assert np.array_equal(codes.numpy(), golden[f"{name}/train{step}_codes"].astype(np.int64)), \
            "train-mode codes differ (sgemm differs from the generating host?)"
        assert np.array_equal(q.detach().numpy(), golden[f"{name}/train{step}_quantized"])
        np.testing.assert_allclose(pen.detach().numpy(), golden[f"{name}/train{step}_penalty"],
                                   rtol=1e-6)
    for i, st in enumerate(states):
        np.testing.assert_array_equal(st["cluster_size"].numpy(),
                                      golden[f"{name}/train_cluster_size{i}"])
        for key in ("embed", "embed_avg"):
            a = st[key].numpy()
            np.testing.assert_allclose(a[::ROW_STRIDE], golden[f"{name}/train_{key}{i}_rows"],
                                       rtol=1e-6, atol=1e-7)
            sums = golden[f"{name}/train_{key}{i}_sums"]
            np.testing.assert_allclose(np.abs(a.astype(np.float64)).sum(), sums[1], rtol=1e-6)


def test_ties(golden):
    x, cb = cases.tie_inputs()
    codes = rvq_oracle.rvq_encode(x, list(cb))
    assert np.array_equal(codes.numpy(), golden["ties/codes"].astype(np.int64))
    assert (codes[0] == 0).all()            # zero codebook -> every distance ties -> index 0
    assert (codes[1] < 32).all()            # duplicated rows -> the first copy wins
    assert np.array_equal(rvq_oracle.rvq_decode(codes, list(cb)).numpy(), golden["ties/decode"])


@pytest.mark.parametrize("name", list(cases.GRVQ_CASES))
def test_grvq(golden, name):
    case = cases.GRVQ_CASES[name]
    x, w = cases.grvq_inputs(case)
    q, loss, ids = grvq_oracle.grvq_forward(x, w)
    codes = torch.stack(ids, -1).reshape(x.shape[0], x.shape[2], -1)
    assert np.array_equal(codes.numpy(), golden[f"{name}/codes"].astype(np.int64))
    assert np.array_equal(q.numpy(), golden[f"{name}/quantized"])
    np.testing.assert_allclose(loss.numpy(), golden[f"{name}/loss"], rtol=1e-6)
    emb = grvq_oracle.grvq_embed(codes, w)
    assert np.array_equal(emb.numpy(), golden[f"{name}/embed"])


def test_kmeans(golden):
    samples, k, iters = cases.kmeans_inputs()
    pick = torch.from_numpy(golden["kmeans/pick"])
    means, bins = rvq_oracle.kmeans_from_means(samples, samples[pick], iters)
    assert np.array_equal(bins.numpy(), golden["kmeans/bins"])
    np.testing.assert_allclose(means.numpy(), golden["kmeans/means"], rtol=1e-6, atol=1e-7)


def test_audit_accepts_reference_and_rejects_garbage(golden):
    case = cases.RVQ_CASES["cfg1_small"]
    x, cb = cases.rvq_inputs(case)
    ref = golden["cfg1_small/codes"].astype(np.int64)
    rep = adjudicate.audit_rvq_codes(x, cb, ref)
    assert sum(rep["wrong"]) == 0, rep
    bad = ref.copy()
    bad[3] = (bad[3] + 1) % case["bins"]
    rep = adjudicate.audit_rvq_codes(x, cb, bad)
    assert rep["wrong"][3] > 0.99 * bad[3].size
    cmp_ = adjudicate.compare_rvq_codes(x, cb, ref, bad)
    assert cmp_["hard_mismatch"] > 0.99 * bad[3].size and cmp_["near_tie"] <= 2


@pytest.mark.parametrize("bits,n", cases.BITPACK_CASES)
def test_bitpack_oracle(golden, bits, n):
    from oracle import bitpack_oracle
    vals = cases.bitpack_values(bits, n)
    want = golden[f"bitpack/{bits}_{n}"]
    got = np.frombuffer(bitpack_oracle.pack(vals, bits), dtype=np.uint8)
    assert np.array_equal(got, want)
    assert np.array_equal(bitpack_oracle.unpack(want.tobytes(), n, bits), vals)
