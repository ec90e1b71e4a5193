"""Pins oracle/hubert_oracle.py: against tests/golden/hubert_tiny.npz (the reference's own HubertFeatureExtractor.extract
run on seeded weights by oracle/make_goldens.py), against sklearn's KMeans.predict directly, and against the known
answers carried by the reference's example_data golden files (frame counts, rel_l rule, dedup, strings, token ids)."""
import os

import numpy as np
import torch

from oracle import hubert_oracle as HO

TINY = HO.OracleHubertConfig(conv_dim=64, hidden=128, n_heads=2, ffn=256, n_layers=3, pos_conv_kernel=16,
                             pos_conv_groups=4, n_units=50, layer=3)


def test_features_and_units_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "hubert_tiny.npz"))
    p = HO.init_hubert_params(TINY, seed=11)
    wav, lens = torch.from_numpy(z["wav"]), torch.from_numpy(z["lens"])
    feat = HO.features(p, TINY, wav)
    assert feat.shape == z["feat"].shape
    assert float((feat - torch.from_numpy(z["feat"])).abs().max()) < 2e-5   # same fp32 ops, same order
    toks = HO.extract(p, TINY, wav, lens)
    assert np.array_equal(toks[0], z["tok0"]) and np.array_equal(toks[1], z["tok1"])   # integer ids: exact


def test_kmeans_predict_matches_sklearn_including_ties():
    from sklearn.cluster import KMeans
    rng = np.random.default_rng(0)
    c = rng.standard_normal((500, 96)).astype(np.float32)
    c[17] = c[3]              # duplicated centre: the lower index must win
    c[400] = c[399]
    x = rng.standard_normal((5000, 96)).astype(np.float32)
    x[:50] = c[17] + 1e-3 * rng.standard_normal((50, 96)).astype(np.float32)
    x[50:60] = c[400]
    km = KMeans(n_clusters=500, n_init=1)
    km.cluster_centers_, km._n_threads, km.n_features_in_ = c, 1, 96
    want = km.predict(x)
    got = HO.kmeans_predict(x, c)
    assert np.array_equal(got, want)
    assert not np.any(got == 17) and not np.any(got == 400)


def test_frame_counts_known_answers():
    cfg = HO.OracleHubertConfig()
    # example_data: audio2 = 255120 samples -> 398 frames; audio1 = 225360 samples padded to 255120 -> ceil rule -> 352
    assert HO.frame_counts(cfg, 255120)[-1] == 398
    assert int(HO.rel_lengths(torch.tensor([225360]), 255120, 398)[0]) == 352
    assert HO.frame_counts(cfg, 480000) == [96015, 48007, 24003, 12001, 6000, 3000, 1500, 750]


def test_dedup_strings_and_ids_match_reference_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "tokeniser.npz"))
    for i, frames in ((0, 398), (1, 352)):
        units, dur = z[f"units{i}"], z[f"dur{i}"]
        assert int(dur.sum()) == frames
        expanded = np.repeat(units, dur)
        u2, d2 = HO.dedup(expanded.tolist())
        assert u2 == units.tolist() and d2 == dur.tolist()
        assert HO.token_ids(units.tolist()) == z[f"ids{i}"].tolist()
    assert HO.stringify([3, 49, 7]) == "<Un3><Un49><Un7>"
    assert HO.dedup([]) == ([], [])
    assert HO.dedup([5, 5, 5]) == ([5], [3])
