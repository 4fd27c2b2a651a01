"""The built library is what DESIGN.md says it is: every kernel is there (sm_100a SASS), the observation tiles of the persistent
kernels leave through cp.async.bulk (UBLKCP), k_rollout_cta's roles meet at producer / consumer named barriers (BAR.ARV),
k_step8 stages its records with cp.async (LDGSTS), and the cooperative generator has not grown back to its inlined size.
Needs cuobjdump (CUDA toolkit), no GPU."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
LIB = os.path.join(ROOT, 'babyai_b200', 'libbabyai_b200.so')


@pytest.fixture(scope='module')
def sass():
    exe = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(exe):
        pytest.skip('cuobjdump not available')
    import sys
    sys.path.insert(0, ROOT)
    from babyai_b200 import build as b
    b.build()
    out = subprocess.run([exe, '-sass', LIB], capture_output=True, text=True, timeout=300).stdout
    assert 'sm_100a' in out
    fns, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = fns.setdefault(m.group(1), [])
        elif cur is not None and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
            cur.append(re.sub(r'^\s+/\*[0-9a-f]+\*/\s+', '', line))
    return fns


def _named(fns, prefix):
    return {k: v for k, v in fns.items() if re.match(r'_Z\d+' + prefix + r'(I|P|N|8|\d)', k)}


def test_every_kernel_is_built(sass):
    for name in ('k_seed', 'k_gen_scan', 'k_gen_small', 'k_gen', 'k_rollout', 'k_rollout_cta', 'k_step8', 'k_render_rgb'):
        assert _named(sass, name), name
    assert len(_named(sass, 'k_gen')) == 2 and len(_named(sass, 'k_rollout_cta')) == 2 and len(_named(sass, 'k_step8')) == 4
    assert len(_named(sass, 'k_rollout')) == 5          # generic <1>, <8>, untracked <1, true>, GoTo-only and Pickup-only single-room


def test_tiles_leave_through_the_bulk_copy_engine(sass):
    for name in ('k_rollout', 'k_rollout_cta'):
        for k, code in _named(sass, name).items():
            assert sum('UBLKCP' in i for i in code) >= 1, k
            assert sum('UTMACMDFLUSH' in i for i in code) >= 1, k      # cp.async.bulk.commit_group


def test_cta_roles_use_arrive_and_sync_barriers(sass):
    for k, code in _named(sass, 'k_rollout_cta').items():
        assert sum('BAR.ARV' in i for i in code) >= 4 and sum('BAR.SYNC' in i for i in code) >= 6, k
    for k, code in _named(sass, 'k_step8').items():
        assert sum('LDGSTS' in i for i in code) >= 8, k                # cp.async staging of the env record


def test_generator_code_size_and_shared_accesses(sass):
    gen = _named(sass, 'k_gen')
    small = min(len(v) for v in gen.values())
    assert small < 12000, small                                    # k_gen<false>: ~8 100 instructions (41 700 when everything was inlined)
    for k, code in gen.items():
        lds = sum(bool(re.match(r'(@!?U?P\d+\s+)?(LDS|STS)', i)) for i in code)
        generic = sum(bool(re.match(r'(@!?U?P\d+\s+)?(LD\.E|ST\.E)', i)) for i in code)
        assert lds > 5 * generic, (k, lds, generic)                # the working memory is accessed as shared memory
