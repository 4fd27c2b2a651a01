"""Builds oracle/_ref/libbabyai_oracle.so from oracle/babyai_oracle.c.
TEST INFRASTRUCTURE ONLY (the checker / CPU baseline, never the product).
-ffp-contract=off keeps `1 - 0.9*(step_count/max_steps)` as a separately
rounded multiply and subtract, like the reference's Python float arithmetic."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'babyai_oracle.c')
OUT_DIR = os.path.join(HERE, '_ref')
OUT = os.path.join(OUT_DIR, 'libbabyai_oracle.so')


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ['gcc', '-O2', '-std=gnu11', '-Wall', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC',
           '-pthread', SRC, '-o', OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
