"""Drop-in check of the mirror package's public signatures against the REAL reference's (tests/golden/ref_signatures.json,
captured by tests/golden/make_signatures_golden.py): same parameter names, order and defaults; extra trailing parameters
(e.g. `precision`, `weights_seed` of the model builders) are allowed."""
import importlib
import inspect
import json
import os

import pytest

SIGS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_signatures.json')))


@pytest.mark.parametrize('key', sorted(SIGS))
def test_signature_is_a_superset_of_the_reference(key):
    mod, path = key.split(':')
    o = importlib.import_module('ssd_keras_b200.' + mod)
    for p in path.split('.'):
        o = getattr(o, p)
    ours = [[n, None if p.default is inspect.Parameter.empty else repr(p.default)]
            for n, p in inspect.signature(o).parameters.items() if p.kind != inspect.Parameter.VAR_KEYWORD]
    ref = SIGS[key]
    assert ours[:len(ref)] == ref, (ours, ref)
