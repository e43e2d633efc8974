"""Time-dependent UVF files: a two-step file assembled by the reference's own block classes and read back
through the reference's UVFDataset (tests/golden/make_golden_timesteps.py) against the product's reader and renderer."""
import gzip
import os

import numpy as np
import pytest

from imagevis3d_b200.uvf import UvfFile

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _file(tmp_path):
    z = np.load(os.path.join(GOLD, "timesteps_case.npz"))
    f = tmp_path / "ts.uvf"
    f.write_bytes(gzip.decompress(z["uvf_gz"].tobytes()))
    return z, str(f)


def test_reader_serves_every_time_step(tmp_path):
    z, f = _file(tmp_path)
    for t in (0, 1):
        with UvfFile(f, timestep=t) as u:
            assert u.info.timesteps == 2
            assert (u.toc()[3] == z["t%d/sizes" % t]).all()
            assert (u.read_bricks() == z["t%d/payload" % t]).all()
            assert (u.minmax()[:, :2] == z["t%d/minmax" % t]).all()
    assert not (z["t0/payload"] == z["t1/payload"]).all()
    with pytest.raises(Exception):
        UvfFile(f, timestep=2)


@pytest.mark.gpu
def test_renderer_set_timestep(tmp_path):
    """renderer.setTimestep: the other step's bricks (written back out and compared with the reference's), the transfer
    function kept, a different picture; a step the file does not have is refused"""
    from imagevis3d_b200.renderer import Renderer
    z, f = _file(tmp_path)
    r = Renderer(0)
    r.loadDataset(f)
    assert r.timesteps() == (2, 0)
    r.resize((128, 96))
    tf = np.zeros((256, 4), np.uint8)
    tf[60:, :] = (200, 120, 40, 30)
    r.set1DTrans(tf)
    r.paintUntilConverged(100)
    a = r.frameFloat()
    r.setTimestep(1)
    assert r.timesteps() == (2, 1)
    r.paintUntilConverged(100)
    b = r.frameFloat()
    assert np.abs(a - b).max() > 0.05
    out = tmp_path / "t1.uvf"
    r.saveUVF(str(out))
    with UvfFile(str(out)) as u:
        assert (u.read_bricks() == z["t1/payload"]).all()
        assert (u.minmax()[:, :2] == z["t1/minmax"]).all()
    r.setTimestep(0)
    r.paintUntilConverged(100)
    assert np.abs(r.frameFloat() - a).max() <= 1e-6            # same bricks, same table: the first picture again
    with pytest.raises(Exception):
        r.setTimestep(2)
    r.cleanup()
