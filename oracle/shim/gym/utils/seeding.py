"""gym 0.1x seeding, restated: np_random(seed) -> (RandomState, seed), the
RandomState being seeded with init_by_array(uint32 words of the first 8 bytes of
sha512(str(seed))).  [recalled; SURVEY.md App. A.3]

Two back-ends, chosen process-wide with set_backend():
  'mt'     -- numpy legacy RandomState (fidelity to historical gym seeds)
  'philox' -- oracle/philox.py, the stream the CUDA generator uses
"""
import hashlib
import os
import struct
import sys

import numpy as np

from .. import error

_BACKEND = os.environ.get('BABYAI_ORACLE_RNG', 'mt')


def set_backend(name):
    global _BACKEND
    assert name in ('mt', 'philox')
    _BACKEND = name


def get_backend():
    return _BACKEND


def _bigint_from_bytes(b):
    sizeof_int = 4
    padding = sizeof_int - len(b) % sizeof_int
    b += b'\0' * padding
    int_count = int(len(b) / sizeof_int)
    unpacked = struct.unpack("{}I".format(int_count), b)
    accum = 0
    for i, val in enumerate(unpacked):
        accum += 2 ** (sizeof_int * 8 * i) * val
    return accum


def _int_list_from_bigint(bigint):
    if bigint < 0:
        raise error.Error('Seed must be non-negative, not {}'.format(bigint))
    elif bigint == 0:
        return [0]
    ints = []
    while bigint > 0:
        bigint, mod = divmod(bigint, 2 ** 32)
        ints.append(mod)
    return ints


def create_seed(a=None, max_bytes=8):
    if a is None:
        a = _bigint_from_bytes(os.urandom(max_bytes))
    elif isinstance(a, str):
        a = a.encode('utf8')
        a += hashlib.sha512(a).digest()
        a = _bigint_from_bytes(a[:max_bytes])
    elif isinstance(a, int):
        a = a % 2 ** (8 * max_bytes)
    else:
        raise error.Error('Invalid type for seed: {} ({})'.format(type(a), a))
    return a


def hash_seed(seed=None, max_bytes=8):
    if seed is None:
        seed = create_seed(max_bytes=max_bytes)
    h = hashlib.sha512(str(seed).encode('utf8')).digest()
    return _bigint_from_bytes(h[:max_bytes])


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, (int, np.integer)) and 0 <= seed):
        raise error.Error('Seed must be a non-negative integer or omitted, not {}'.format(seed))
    seed = create_seed(None if seed is None else int(seed))
    if _BACKEND == 'philox':
        here = os.path.dirname(os.path.abspath(__file__))
        root = os.path.normpath(os.path.join(here, '..', '..', '..'))
        if root not in sys.path:
            sys.path.insert(0, root)
        from philox import PhiloxRandom
        return PhiloxRandom(seed), seed
    rng = np.random.RandomState()
    rng.seed(_int_list_from_bigint(hash_seed(seed)))
    return rng, seed
