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
    """The bench transfer functions have a transparent low end: bricks whose maximum stays below the first non-zero
    alpha entry are the ones visibility flags empty and the e2e leg never pages in."""
    sys.path.insert(0, ROOT)
    import bench
    for n in (256, 4096):
        rgba, lo, hi = bench.tf1d_std(n)
        assert rgba.shape == (n, 4) and (rgba[:int(lo), 3] == 0).all() and rgba[int(lo), 3] > 0 and hi == n - 1
        assert 0.25 * n <= lo <= 0.4 * n
    t2, nz = bench.tf2d(4096)
    assert t2.shape == (256, 4096, 4) and (t2[:, :int(nz[0]), 3] == 0).all() and t2[:, int(nz[0]), 3].max() > 0
    assert nz[2:] == (0.0, 255.0)


def test_bench_workloads_name_the_baseline_configs():
    """N=1 is configs[2] (2048^3 u16, 1080p, 2D-TF lit), N>1 is configs[3] (4096^3 u16, 4K): the line says which."""
    sys.path.insert(0, ROOT)
    import bench
    c = bench.workload_config(1, "2d_lit")
    assert "configs[2]" in c["workload"] and c["volume"] == [2048] * 3 and c["resolution"] == [1920, 1080]
    c = bench.workload_config(8, "1d_lit")
    assert "configs[3]" in c["workload"] and c["volume"] == [4096] * 3 and c["resolution"] == [3840, 2160]
