"""GPU parity of hot path (i): waveform -> HuBERT features -> k-means unit ids -> dedup, through the C ABI, against
oracle/hubert_oracle.py (pinned to the reference's extract() by tests/golden/hubert_tiny.npz).

Bars: unit ids / durations / counts are integers -> compared exactly (mismatches only tolerated where the fp64 top-2
margin of the oracle's own distances is below the fp32 noise of the feature, and then counted and bounded);
fp32 features within 2e-4 relative (split-bf16 products are ~2^-16 accurate, accumulation is fp32)."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mk(o, seed, max_batch, max_samples):
    from oracle import hubert_oracle as HO
    from slamkit_b200.feature_extractor import HubertB200Config, HubertB200FeatureExtractor
    p = HO.init_hubert_params(o, seed=seed)
    c = HubertB200Config(conv_dim=o.conv_dim, conv_kernel=o.conv_kernel, conv_stride=o.conv_stride, hidden=o.hidden,
                         n_heads=o.n_heads, ffn=o.ffn, layer=o.layer, pos_conv_kernel=o.pos_conv_kernel,
                         pos_conv_groups=o.pos_conv_groups, n_units=o.n_units, ln_eps=o.ln_eps, pad=o.pad)
    fe = HubertB200FeatureExtractor(c, p, device=DEV, max_batch=max_batch, max_samples=max_samples)
    return fe, p


def _tiny():
    from oracle import hubert_oracle as HO
    return HO.OracleHubertConfig(conv_dim=64, hidden=128, n_heads=2, ffn=256, n_layers=3, pos_conv_kernel=16,
                                 pos_conv_groups=4, n_units=50, layer=3)


def test_units_match_reference_golden(golden_dir):
    """The fixture holds what the reference's own HubertFeatureExtractor.extract returned for these weights/clips."""
    z = np.load(os.path.join(golden_dir, "hubert_tiny.npz"))
    fe, p = _mk(_tiny(), 11, 2, 16000)
    wav, lens = torch.from_numpy(z["wav"]), torch.from_numpy(z["lens"])
    feat = fe.features(wav).cpu()
    assert rel_err(feat, torch.from_numpy(z["feat"])) < 2e-4
    toks = fe.extract(wav, lens)
    assert [len(t) for t in toks] == [len(z["tok0"]), len(z["tok1"])]
    assert np.array_equal(toks[0], z["tok0"]) and np.array_equal(toks[1], z["tok1"])


def test_every_stage_against_oracle():
    """Stage-by-stage taps (conv layers, projection, positional conv, each encoder layer) on a ragged 2-clip batch."""
    from oracle import hubert_oracle as HO
    o = HO.OracleHubertConfig(conv_dim=128, hidden=192, n_heads=3, ffn=384, n_layers=2, pos_conv_kernel=32,
                              pos_conv_groups=4, n_units=100, layer=2)
    fe, p = _mk(o, 4, 2, 24000)
    g = torch.Generator().manual_seed(1)
    wav = (0.1 * torch.randn(2, 24000, generator=g)).clamp(-1, 1)
    wav[1, 17000:] = 0
    T = HO.frame_counts(o, 24000)
    taps = {}
    HO.features(p, o, wav, taps=taps)
    # conv stack: compare layer by layer by re-running the oracle's conv encoder incrementally
    h = torch.nn.functional.pad(wav, (o.pad, o.pad))[:, None]
    h = torch.nn.functional.gelu(torch.nn.functional.group_norm(
        torch.nn.functional.conv1d(h, p["conv0.weight"], stride=o.conv_stride[0]), o.conv_dim, p["gn.weight"], p["gn.bias"], 1e-5))
    got = fe.debug_stage(wav, 100, 2 * T[0], o.conv_dim).cpu().view(2, T[0], -1).transpose(1, 2)
    assert rel_err(got, h) < 1e-4, ("conv0", rel_err(got, h))
    for i in range(1, 8):
        h = torch.nn.functional.gelu(torch.nn.functional.conv1d(h, p[f"conv{i}.weight"], stride=o.conv_stride[i]))
        got = fe.debug_stage(wav, 100 + i, 2 * T[i], o.conv_dim).cpu().view(2, T[i], -1).transpose(1, 2)
        assert rel_err(got, h) < 1e-4, (f"conv{i}", rel_err(got, h))
    Tf = T[-1]
    x = torch.nn.functional.linear(torch.nn.functional.layer_norm(h.transpose(1, 2), (o.conv_dim,), p["fp.ln.weight"], p["fp.ln.bias"], o.ln_eps),
                                   p["fp.proj.weight"], p["fp.proj.bias"])
    got = fe.debug_stage(wav, 200, 2 * Tf, o.hidden).cpu().view(2, Tf, -1)
    assert rel_err(got, x) < 1e-4, ("proj", rel_err(got, x))
    pc = torch.nn.functional.conv1d(x.transpose(1, 2), HO.pos_conv_weight(p), p["pos.bias"], padding=o.pos_conv_kernel // 2,
                                    groups=o.pos_conv_groups)[:, :, :-1]
    got = fe.debug_stage(wav, 201, 2 * Tf, o.hidden).cpu().view(2, Tf, -1)
    assert rel_err(got, torch.nn.functional.gelu(pc).transpose(1, 2)) < 1e-4, "posconv"
    got = fe.debug_stage(wav, 0, 2 * Tf, o.hidden).cpu().view(2, Tf, -1)
    assert rel_err(got, taps["embed"]) < 1e-4, "embed"
    for l in range(o.layer):
        got = fe.debug_stage(wav, l + 1, 2 * Tf, o.hidden).cpu().view(2, Tf, -1)
        assert rel_err(got, taps[f"layer{l}"]) < 2e-4, (f"layer{l}", rel_err(got, taps[f"layer{l}"]))


@pytest.mark.parametrize("B,S", [(3, 48000), (2, 80000)])
def test_units_vs_oracle_midsize(B, S):
    """mHuBERT-25Hz conv geometry with a narrower body; ragged lengths sorted descending like cli/extract_features.py."""
    from oracle import hubert_oracle as HO
    o = HO.OracleHubertConfig(conv_dim=256, hidden=256, n_heads=4, ffn=512, n_layers=4, pos_conv_kernel=128,
                              pos_conv_groups=16, n_units=500, layer=4)
    fe, p = _mk(o, 21, B, S)
    g = torch.Generator().manual_seed(S)
    wav = (0.1 * torch.randn(B, S, generator=g)).clamp(-1, 1)
    lens = torch.tensor(sorted([S] + [int(S * f) for f in (0.71, 0.33)][:B - 1], reverse=True))
    for b in range(B):
        wav[b, lens[b]:] = 0
    want = HO.extract(p, o, wav, lens)
    got = fe.extract(wav, lens)
    assert [len(x) for x in got] == [len(x) for x in want]
    feat = HO.features(p, o, wav).numpy()
    _, margin = HO.kmeans_margins(feat.reshape(-1, o.hidden), p["kmeans.centers"].numpy())
    margin = margin.reshape(B, -1)
    n_bad = 0
    for b in range(B):
        diff = np.nonzero(got[b] != want[b])[0]
        # any disagreement must sit on a near-tie of the fp64 distances (margin below fp32 feature noise)
        assert all(margin[b, t] < 5e-3 for t in diff), (b, diff[:5], margin[b, diff[:5]])
        n_bad += len(diff)
    total = sum(len(x) for x in want)
    assert n_bad <= max(1, total // 500), (n_bad, total)
    assert rel_err(fe.features(wav).cpu(), torch.from_numpy(feat)) < 2e-4


def test_kmeans_argmin_first_min_tie_break():
    import ctypes as C
    from oracle import hubert_oracle as HO
    from slamkit_b200 import _lib as L
    lib = L.require_cuda()
    rng = np.random.default_rng(0)
    U, D, M = 500, 64, 3000
    c = rng.standard_normal((U, D)).astype(np.float32)
    c[17], c[400] = c[3], c[399]
    x = rng.standard_normal((M, D)).astype(np.float32)
    x[:64] = c[17]
    x[64:90] = c[400]
    dot = torch.from_numpy(x @ c.T).to(DEV).contiguous()
    cd = torch.from_numpy(c).to(DEV)
    csq = torch.empty(U, device=DEV)
    labels = torch.empty(M, dtype=torch.int32, device=DEV)
    L.check(lib.sk_row_sqnorm(L.ptr(cd), L.ptr(csq), U, D, L.stream_ptr()))
    L.check(lib.sk_kmeans_argmin(L.ptr(dot), L.ptr(csq), L.ptr(labels), M, U, U, L.stream_ptr()))
    got = labels.cpu().numpy()
    # same fp32 dot products -> the argmin / tie-break rule itself must agree exactly with the sklearn restatement
    d = (csq.cpu().numpy()[None, :] + np.float32(-2.0) * dot.cpu().numpy())
    assert np.array_equal(got, np.argmin(d, axis=1))
    assert not np.any(got == 17) and not np.any(got == 400)
    assert (got == HO.kmeans_predict(x, c)).mean() > 0.999


def test_rle_matches_groupby_including_empty_and_ragged():
    from oracle import hubert_oracle as HO
    from slamkit_b200 import _lib as L
    lib = L.require_cuda()
    rng = np.random.default_rng(1)
    B, T = 7, 300
    ids = rng.integers(0, 5, size=(B, T)).astype(np.int32)
    ids[2] = 3                      # one long run
    ids[3, :] = np.arange(T) % 7    # no repeats
    nf = np.array([300, 1, 300, 299, 0, 33, 64], dtype=np.int32)
    d_ids, d_nf = torch.from_numpy(ids).to(DEV), torch.from_numpy(nf).to(DEV)
    units, dur = torch.empty_like(d_ids), torch.empty_like(d_ids)
    cnt = torch.empty(B, dtype=torch.int32, device=DEV)
    L.check(lib.sk_rle(L.ptr(d_ids), L.ptr(d_nf), L.ptr(units), L.ptr(dur), L.ptr(cnt), B, T, L.stream_ptr()))
    u, d, c = units.cpu().numpy(), dur.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        wu, wd = HO.dedup(ids[b, :nf[b]].tolist())
        assert c[b] == len(wu)
        assert u[b, :c[b]].tolist() == wu and d[b, :c[b]].tolist() == wd
        assert int(d[b, :c[b]].sum()) == nf[b]


def test_tokeniser_end_to_end_on_gpu():
    """audio_represent -> strings -> ids through the mirrors equals the oracle's extract + dedup + stringify + ids."""
    from oracle import hubert_oracle as HO
    from slamkit_b200.tokeniser import B200UnitTokeniser
    fe, p = _mk(_tiny(), 11, 2, 16000)
    tok = B200UnitTokeniser(fe, num_units=50)
    g = torch.Generator().manual_seed(9)
    wav = (0.1 * torch.randn(2, 16000, generator=g)).clamp(-1, 1)
    lens = torch.tensor([16000, 9000])
    wav[1, 9000:] = 0
    reps = tok.audio_represent(wav, lens)
    want = HO.extract(p, _tiny(), wav, lens)
    for r, w in zip(reps, want):
        wu, wd = HO.dedup(w.tolist())
        assert list(r["units"]) == wu and list(r["duration"]) == wd
        assert tok.stringify_representation([r])[0] == HO.stringify(wu)
        assert tok(r)["input_ids"] == HO.token_ids(wu)


def test_full_size_properties():
    """cfg-3 geometry (mHuBERT-25Hz, 30 s clips) at a small batch: frame count 750, valid id range, batch-order
    equivariance and determinism (size-independent properties in place of a full-size CPU oracle run)."""
    from oracle import hubert_oracle as HO
    o = HO.OracleHubertConfig()
    fe, p = _mk(o, 2, 4, 480000)
    g = torch.Generator().manual_seed(3)
    wav = (0.1 * torch.randn(4, 480000, generator=g)).clamp(-1, 1)
    ids, nf = fe.units_device(wav, None)
    assert ids.shape == (4, 750) and nf.tolist() == [750] * 4
    a = ids.cpu().numpy()
    assert a.min() >= 0 and a.max() < 500
    ids2, _ = fe.units_device(wav, None)
    assert torch.equal(ids, ids2)
    perm = torch.tensor([2, 0, 3, 1])
    ids3, _ = fe.units_device(wav[perm], None)
    assert (ids3.cpu().numpy() != a[perm.numpy()]).mean() < 2e-3   # clips are independent when all are full length
    u, d, c = fe.dedup_device(ids, nf)
    assert int(d[0, :int(c[0])].sum()) == 750
