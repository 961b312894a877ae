"""Dataset __getitem__ restatements against fixtures produced by executing the reference's AudioDataset classes
(tests/golden/make_golden_datasets.py), and the native draw planner against per-draw Python."""
import json
import os
import random

import numpy as np
import torch

from cases import DATASET_SPECS, hash_spec, sha
from oracle import frontend as F

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "ref_datasets.json")))["cases"]
SPECS64 = [hash_spec(T, 64, seed=300 + T) for T in DATASET_SPECS["rows64"]]
SPECS128 = [hash_spec(T, 128, seed=500 + T) for T in DATASET_SPECS["rows128"]]
ORDER = DATASET_SPECS["order"]


def rng_fingerprint():
    return sha(np.array(random.getstate()[1], dtype=np.uint64))


def test_oracle_cola_items_reproduce_the_executed_reference():
    for windowing in (False, True):
        for augment in (False, True):
            g = G[f"mae_training/cola/w{int(windowing)}a{int(augment)}"]
            random.seed(4242)
            items = [F.dataset_cola_item(SPECS64[i], 251, augment, windowing) for i in ORDER]
            assert [sha(a) for a, _ in items] == g["x1"] and [sha(b) for _, b in items] == g["x2"]
            assert rng_fingerprint() == g["rng_after"]
    g = G["cola_training/cola/a1"]
    random.seed(4243)
    items = [F.dataset_cola_item(SPECS64[i], 251, True, False) for i in ORDER]
    assert [sha(a) for a, _ in items] == g["x1"] and rng_fingerprint() == g["rng_after"]


def test_oracle_mae_items_reproduce_the_executed_reference():
    random.seed(77)
    assert [sha(F.dataset_mae_item(SPECS64[i], 256)) for i in ORDER] == G["mae_training/mae/256"]["x"]
    assert rng_fingerprint() == G["mae_training/mae/256"]["rng_after"]
    random.seed(78)
    assert [sha(F.dataset_mae_item(s, 1024)) for s in SPECS128] == G["mae_training/audiomae/1024"]["x"]
    assert rng_fingerprint() == G["mae_training/audiomae/1024"]["rng_after"]


def test_oracle_finetune_items_reproduce_the_executed_reference():
    for name in ("first", "random_aug", "specaug", "specaug_only"):
        g = G[f"finetuning/{name}"]
        random.seed(990)
        torch.manual_seed(991)
        xs = [F.dataset_finetune_item(SPECS64[i], **g["kw"]) for i in ORDER]
        assert [list(x.shape) for x in xs] == g["shape"]
        for x, s, zr, zc in zip(xs, g["sum"], g["zero_rows"], g["zero_cols"]):
            assert float(np.asarray(x, np.float64).sum()) == s
            assert [int(r) for r in np.flatnonzero((x == 0).all(axis=1))] == zr
            assert [int(c) for c in np.flatnonzero((x == 0).all(axis=0))] == zc
        assert rng_fingerprint() == g["rng_after"]
        assert sha(torch.get_rng_state().numpy()) == g["torch_rng_after"]


def test_native_draw_planner_equals_per_draw_python():
    """hmfe_cola_draws (host-only entry of libhmfe.so) consumes Python's Mersenne Twister exactly as the per-item
    Python code does: masks, window / crop starts, gains and the generator state afterwards are identical."""
    from heart_murmur_detection_b200 import datasets as D
    from heart_murmur_detection_b200.util import draw_mask_rows

    rows = [300, 251, 1000, 777, 2500, 260, 754, 753]
    for windowing in (False, True):
        for augment in (True, False):
            random.seed(99)
            dr = D.cola_draws(rows, 251, augment, windowing)
            after = random.getstate()
            random.seed(99)
            for i, T in enumerate(rows):
                w = 0
                if windowing and T > 753:
                    w = int(random.random() * (T - 753))
                    T = 753
                if augment:
                    m = draw_mask_rows(T)
                    assert (m == dr["mask"][dr["mask_off"][i] : dr["mask_off"][i] + T]).all()
                s1, s2 = int(random.random() * (T - 251)), int(random.random() * (T - 251))
                assert (w, s1, s2) == (dr["win_start"][i], dr["start1"][i], dr["start2"][i])
                if augment:
                    assert np.float32(0.9 + random.random() / 5.0) == dr["gain1"][i]
                    assert np.float32(0.9 + random.random() / 5.0) == dr["gain2"][i]
            assert random.getstate() == after
