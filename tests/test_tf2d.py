"""T2 TransferFunction2D: the ".2dt" file (Load / Save pinned to the compiled reference class) and the swatch rasteriser
(the reference paints with Qt's QPainter, which is not available: the rules are restated and checked on their own)."""
import os
import subprocess

import numpy as np
import pytest

from imagevis3d_b200.renderer import Renderer
from oracle import oracle as O

C60 = "/root/reference/DemoData/c60.2dt"


def _write_2dt(path, w, h, swatches, alpha1d=0.0):
    with open(path, "w") as f:
        f.write("%d %d\n%d\n" % (w, h, w))
        for i in range(w):
            f.write("%g %g %g %g \n" % (i / (w - 1.0), 0.25, 0.5, alpha1d))
        f.write("%d\n" % len(swatches))
        for radial, pts, grad, stops in swatches:
            f.write("%d\n%d\n" % (radial, len(pts)))
            for p in pts:
                f.write("%g %g \n" % p)
            f.write("%g %g %g %g\n%d\n" % (grad[0][0], grad[0][1], grad[1][0], grad[1][1], len(stops)))
            for pos, c in stops:
                f.write("%g  %g %g %g %g \n" % ((pos,) + tuple(c)))


RECT = [(0.25, 0.25), (0.25, 0.75), (0.5, 0.75), (0.5, 0.25)]


@pytest.mark.skipif(not (os.path.isfile(O.ref_driver_path() or "") and os.path.isfile(C60)),
                    reason="oracle/_ref/ref_driver or the reference's DemoData is not available (/root/reference absent)")
def test_2dt_load_save_round_trip_is_the_reference_s(tmp_path):
    mine = tmp_path / "synthetic.2dt"
    _write_2dt(str(mine), 64, 32, [(0, RECT, ((0.25, 0.5), (0.5, 0.5)), [(0, (1, 0, 0, 0.125)), (0.333333, (0.1, 0.7, 0.3, 1)), (1, (0, 0, 1, 0.5))]),
                                   (1, [(0.6, 0.1), (0.9, 0.2), (0.75, 0.9)], ((0.75, 0.4), (0.9, 0.4)), [(0.5, (0.479873, 0.585009, 0.80874, 0.14))])])
    for src in (C60, str(mine)):
        a, b = str(tmp_path / "ours.2dt"), str(tmp_path / "ref.2dt")
        Renderer.renderTF2D(src, 64, 64, save_as=a)
        subprocess.check_call([O.ref_driver_path(), "tf2d", src, b])
        assert open(a, "rb").read() == open(b, "rb").read()


def test_2dt_rasteriser_rules(tmp_path):
    f = str(tmp_path / "t.2dt")
    # constant-colour rectangle over a transparent 1-D function
    _write_2dt(f, 64, 64, [(0, RECT, ((0.25, 0.5), (0.5, 0.5)), [(0, (1, 0.5, 0, 0.5)), (1, (1, 0.5, 0, 0.5))])])
    t, nz = Renderer.renderTF2D(f, 64, 64)
    inside = np.zeros((64, 64), bool)
    inside[16:48, 16:32] = True                                # pixel centres inside [0.25,0.5) x [0.25,0.75): columns 16..31, rows 16..47
    assert (t[inside] == (255, 127, 0, 127)).all()             # QColor(int(c*255)): 0.5 -> 127
    assert (t[~inside][:, 3] == 0).all()
    assert nz == [16.0, 31.0, 16.0, 47.0]                      # ComputeNonZeroLimits: min x, max x, min y, max y
    # linear gradient along x: alpha rises monotonically across the swatch, pad spread outside the gradient line
    _write_2dt(f, 64, 64, [(0, RECT, ((0.3, 0.5), (0.45, 0.5)), [(0, (1, 1, 1, 0.1)), (1, (1, 1, 1, 0.9))])])
    t, _ = Renderer.renderTF2D(f, 64, 64)
    a = t[30, 16:32, 3].astype(int)
    assert (np.diff(a) >= 0).all() and a[0] == 25 and a[-1] == 229 and (t[20, 16:32, 3] == t[40, 16:32, 3]).all()
    # radial gradient: opaque centre, transparent rim; drawn OVER the 1-D function (source-over), which fills every row
    _write_2dt(f, 64, 64, [(1, [(0.1, 0.1), (0.1, 0.9), (0.9, 0.9), (0.9, 0.1)], ((0.5, 0.5), (0.75, 0.5)), [(0, (0, 1, 0, 1)), (1, (0, 1, 0, 0))])], alpha1d=0.25)
    t, nz = Renderer.renderTF2D(f, 64, 64)
    assert tuple(t[32, 32]) == (0, 255, 0, 255) or t[32, 32, 3] >= 250
    assert t[32, 50, 3] == 63 and t[2, 2, 3] == 63             # beyond the radius / outside the polygon: the 1-D function's alpha (int(0.25*255))
    assert (t[0, :, 0] == np.arange(64) * 255 // 63).all() or abs(int(t[0, 63, 0]) - 255) <= 1   # 1-D ramp stretched over every row
    assert nz == [0.0, 63.0, 0.0, 63.0]
    # target size differs from the file's: the 1-D part is resampled, swatches are resolution independent
    t2, nz2 = Renderer.renderTF2D(f, 128, 256)
    assert t2.shape == (256, 128, 4) and t2[128, 64, 3] >= 250 and nz2 == [0.0, 127.0, 0.0, 255.0]


@pytest.mark.gpu
def test_renderer_load_2d_trans(tmp_path):
    """renderer.load2DTrans(file) == rasterise + set2DTrans with the same table"""
    from oracle import host as H
    f = str(tmp_path / "t.2dt")
    _write_2dt(f, 256, 256, [(0, [(0.1, 0.0), (0.1, 1.0), (0.9, 1.0), (0.9, 0.0)], ((0.1, 0.5), (0.9, 0.5)), [(0, (1, 0.2, 0.1, 0.0)), (1, (0.2, 0.4, 1, 0.6))])])
    raw = H.synth_sphere((96, 96, 96), np.uint8)
    frames = []
    for how in ("file", "table"):
        r = Renderer(0)
        r.loadRawVolume(raw, hist2d=True)
        r.resize((160, 120))
        if how == "file":
            r.load2DTrans(f)
        else:
            t, nz = Renderer.renderTF2D(f, 256, 256)
            r.set2DTrans(t, nz)
        r.setRenderMode(1)
        r.paintUntilConverged(100)
        frames.append(r.frameFloat())
        r.cleanup()
    assert (frames[0][..., 3] > 0).sum() > 1000
    assert np.abs(frames[0] - frames[1]).max() == 0.0
