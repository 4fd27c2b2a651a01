"""Name-only stand-in for the `blosc` package (reference setup.py:13): the
reference imports it in babyai/utils/demos.py for demo (de)serialisation, which
is outside the hot path.  pack_array/unpack_array are provided with pickle so
that importing babyai works.  TEST INFRASTRUCTURE ONLY."""
import pickle


def pack_array(arr):
    return pickle.dumps(arr)


def unpack_array(buf):
    return pickle.loads(buf)
