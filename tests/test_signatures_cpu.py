"""The drop-in mirrors keep the reference's call signatures verbatim: argument names, positional order and
defaults, compared with ``tests/golden/ref_signatures.json`` (read from the reference's source by
``tests/golden/make_signatures.py``)."""
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "ref_signatures.json")))


@pytest.mark.parametrize("key", sorted(REF))
def test_mirror_signature_is_the_reference_signature(key):
    ref = REF[key]
    fn = getattr(importlib.import_module(ref["mirror"]), key.split("::")[1])
    params = list(inspect.signature(fn).parameters.values())
    assert [p.name for p in params] == ref["args"], f"{key} (reference line {ref['lineno']})"
    got_defaults = {p.name: p.default for p in params if p.default is not inspect.Parameter.empty}
    assert got_defaults == ref["defaults"], f"{key} (reference line {ref['lineno']})"
    assert all(p.kind == inspect.Parameter.POSITIONAL_OR_KEYWORD for p in params)
