"""The C-ABI library loads and exports every symbol include/babyai_b200.h declares
(no compute calls here: this runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

from babyai_b200 import lib as bl
from babyai_b200.levels import LEVELS, VOCAB, LevelSpec, level_spec

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'babyai_b200.h')).read()
    return sorted(set(re.findall(r'\b(bb_[a-z_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(bl.LIB_PATH):
        import babyai_b200.build as b
        b.build()
    L = bl.load()
    declared = _header_symbols()
    assert declared, 'no symbols parsed from the header'
    for s in declared:
        assert hasattr(L, s), s
    assert sorted(bl.SYMBOLS) == declared


def test_vocab_matches_library():
    L = bl.load()
    assert L.bb_vocab_size() == len(VOCAB) - 1
    assert [L.bb_vocab_word(i).decode() for i in range(len(VOCAB))] == VOCAB


def test_level_spec_layout_matches_header():
    assert C.sizeof(LevelSpec) == 120         # 8 int32, double, 3 int32, 1+4 int32, 1+3 int32, 2 int32, 2 int32 (verifier modes), 3 int32 (bonus) + pad
    assert LevelSpec.locked_room_prob.offset == 32
    for name in LEVELS:
        s = level_spec(name)
        assert 3 <= s.room_size <= 20


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    L = bl.load()
    h = C.c_void_p()
    spec = level_spec('GoToLocal')
    rc = L.bb_pool_create(C.byref(spec), 4, 0, C.byref(h))
    assert rc != 0 and b'CUDA' in L.bb_last_error()
    with pytest.raises(RuntimeError):
        from babyai_b200 import BabyAIVecEnv
        BabyAIVecEnv('GoToLocal', 4)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under babyai_b200/ may reference it."""
    pk = os.path.join(ROOT, 'babyai_b200')
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'babyai_oracle' not in txt and 'hostemu' not in txt.replace('tests/hostemu', ''), f
