"""Host-side pieces of bench.py that need no GPU: the one-JSON-line contract on stdout and the reference arm's
rank gating (the driver parses stdout; under torchrun NCCL's version banner must not join the line)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_emit_line_goes_to_the_saved_descriptor():
    """After the descriptor swap of the multi-rank set-up, anything written to fd 1 lands on stderr and only
    emit_line reaches the real stdout."""
    code = (
        "import os, sys; sys.path.insert(0, %r); import bench\n"
        "sys.stdout.flush(); bench._REAL_STDOUT = os.dup(1); os.dup2(2, 1)\n"
        "os.write(1, b'NCCL version banner\\n'); print('library chatter')\n"
        "bench.emit_line({'metric': 'ray samples/s', 'value': 1.0})\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "ray samples/s", "value": 1.0}
    assert "NCCL version banner" in r.stderr and "library chatter" in r.stderr


def test_emit_line_plain_stdout():
    code = "import sys; sys.path.insert(0, %r); import bench; bench.emit_line({'a': 1})" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'


def test_bench_tf_visible_range():
    """The bench transfer function is transparent below value 65: bricks whose maximum stays below it are the ones
    the e2e leg never uploads."""
    sys.path.insert(0, ROOT)
    import bench
    rgba, lo, hi = bench.bench_tf()
    assert rgba.shape == (256, 4) and (rgba[:int(lo), 3] == 0).all() and rgba[int(lo), 3] >= 0 and hi == 255.0
    assert lo == 65.0
