#!/usr/bin/env python
"""Golden fixtures of the Dataset ``__getitem__`` methods, produced by EXECUTING the reference's classes.

The three ``AudioDataset`` classes (src/pretrain/cola_training.py, src/pretrain/mae_training.py,
src/benchmark/other_eval/finetuning.py) live in modules whose other imports (lightning, hydra, wandb, the model zoo)
are not installable here, so the class definitions are taken from the reference files AT GENERATION TIME (``ast``; no
reference text enters this repository) and executed with the reference's own ``random_crop / random_mask /
random_multiply / crop_first`` (the real ``src/util.py``, imported as in make_golden.py).  ``SpecAugmentation`` is the
oracle's restatement of torchlibrosa (absent here): the class's control flow and draw order are the reference's own.

Runs only in the build container (needs /root/reference); writes ``ref_datasets.json``.
Usage:  python tests/golden/make_golden_datasets.py
"""
from __future__ import annotations

import ast
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from cases import DATASET_SPECS, hash_spec, sha  # noqa: E402
from make_golden import REFERENCE, install_shims, rng_fingerprint  # noqa: E402

from oracle import frontend as F  # noqa: E402


def reference_class(rel_path, ns):
    src = open(os.path.join(REFERENCE, rel_path)).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "AudioDataset")
    code = compile(ast.Module(body=[node], type_ignores=[]), rel_path, "exec")
    exec(code, ns)
    return ns["AudioDataset"]


def torch_fingerprint():
    return sha(torch.get_rng_state().numpy())


def main():
    install_shims({})
    import src.util as ref  # the real reference

    ns = {"torch": torch, "np": np, "random_crop": ref.random_crop, "random_mask": ref.random_mask,
          "random_multiply": ref.random_multiply, "crop_first": ref.crop_first, "SpecAugmentation": F.SpecAugmentation}
    specs64 = [hash_spec(T, 64, seed=300 + T) for T in DATASET_SPECS["rows64"]]
    specs128 = [hash_spec(T, 128, seed=500 + T) for T in DATASET_SPECS["rows128"]]
    order = DATASET_SPECS["order"]
    cases = {}

    Mae = reference_class("src/pretrain/mae_training.py", dict(ns))
    for windowing in (False, True):
        for augment in (False, True):
            ds = Mae(specs64, max_len=251, augment=augment, method="cola", windowing=windowing)
            random.seed(4242)
            items = [ds[i] for i in order]
            cases[f"mae_training/cola/w{int(windowing)}a{int(augment)}"] = {
                "x1": [sha(a.numpy()) for a, _ in items], "x2": [sha(b.numpy()) for _, b in items],
                "x1_sum": [float(a.double().sum()) for a, _ in items], "rng_after": rng_fingerprint()}
    ds = Mae(specs64, max_len=256, method="mae")
    random.seed(77)
    items = [ds[i] for i in order]
    cases["mae_training/mae/256"] = {"x": [sha(np.ascontiguousarray(a)) for a in items], "rng_after": rng_fingerprint()}
    ds = Mae(specs128, max_len=1024, method="audiomae")
    random.seed(78)
    items = [ds[i] for i in range(len(specs128))]
    cases["mae_training/audiomae/1024"] = {"x": [sha(np.ascontiguousarray(a)) for a in items], "rng_after": rng_fingerprint()}

    Cola = reference_class("src/pretrain/cola_training.py", dict(ns))
    ds = Cola(specs64, max_len=251, augment=True)
    random.seed(4243)
    items = [ds[i] for i in order]
    cases["cola_training/cola/a1"] = {"x1": [sha(a.numpy()) for a, _ in items], "x2": [sha(b.numpy()) for _, b in items],
                                      "x1_sum": [float(a.double().sum()) for a, _ in items], "rng_after": rng_fingerprint()}

    Ft = reference_class("src/benchmark/other_eval/finetuning.py", dict(ns))
    labels = list(range(len(specs64)))
    for name, kw in (("first", dict(max_len=251, augment=False, crop_mode="first")),
                     ("random_aug", dict(max_len=251, augment=True, crop_mode="random")),
                     ("specaug", dict(max_len=251, augment=True, crop_mode="random", spec_augment=True, time_drop_width=64,
                                      time_stripes_num=2, freq_drop_width=8, freq_stripes_num=2)),
                     ("specaug_only", dict(max_len=251, augment=False, crop_mode="first", spec_augment=True))):
        ds = Ft((specs64, labels), **kw)
        random.seed(990)
        torch.manual_seed(991)
        items = [ds[i] for i in order]
        xs = [x.numpy() for x, _ in items]
        cases[f"finetuning/{name}"] = {
            "kw": kw, "sum": [float(np.asarray(x, np.float64).sum()) for x in xs],
            "zero_rows": [[int(r) for r in np.flatnonzero((x == 0).all(axis=1))] for x in xs],
            "zero_cols": [[int(c) for c in np.flatnonzero((x == 0).all(axis=0))] for x in xs],
            "probe": [[float(v) for v in x.reshape(-1)[:: max(1, x.size // 16)][:16]] for x in xs],
            "shape": [list(x.shape) for x in xs], "rng_after": rng_fingerprint(), "torch_rng_after": torch_fingerprint()}
    out = os.path.join(HERE, "ref_datasets.json")
    json.dump({"generator": "tests/golden/make_golden_datasets.py", "reference_files": [
        "src/pretrain/mae_training.py", "src/pretrain/cola_training.py", "src/benchmark/other_eval/finetuning.py",
        "src/util.py"], "cases": cases}, open(out, "w"), indent=0)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
