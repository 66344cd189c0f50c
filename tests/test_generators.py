"""the dataset generators.  The unmodified reference functions (nn/datasets/generators.py, from
oracle/_ref or /root/reference) are run with this package's rasteriser standing in for the two scikit-image calls they
make (skimage.draw.circle no longer exists; matplotlib is stubbed): trajectories, rejection sampling and numpy RNG
consumption are then the reference's own, so the .npz files must be byte-identical to ours."""
import importlib.machinery
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fetch_reference  # noqa: E402
from paig_reproduction_b200 import generators as mine  # noqa: E402

REF = fetch_reference.reference_path()
needs_ref = pytest.mark.skipif(REF is None, reason="reference copy not present (oracle/fetch_reference.py)")


def _reference_generators():
    def stub(name, **attrs):
        m = sys.modules.get(name)
        if m is None or not hasattr(m, "__paig_stub__"):
            try:
                __import__(name)
                return sys.modules[name]
            except ImportError:
                m = types.ModuleType(name)
                m.__spec__ = importlib.machinery.ModuleSpec(name, None)
                m.__paig_stub__ = True
                sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    class _Ax:
        def __getattr__(self, _):
            return lambda *a, **k: _Ax()

    plt = stub("matplotlib.pyplot")
    if hasattr(plt, "__paig_stub__"):
        plt.switch_backend = lambda *a, **k: None
        plt.Normalize = lambda *a, **k: None
        plt.subplots = lambda *a, **k: (_Ax(), _Ax())
        plt.close = lambda *a, **k: None
        cm = stub("matplotlib.cm", Greys_r=None)
        mpl = stub("matplotlib")
        mpl.pyplot, mpl.cm = plt, cm
    sk = stub("skimage")
    sk.draw = stub("skimage.draw", circle=lambda r, c, radius, shape=None: mine.disc_indices(r, c, radius, shape))
    sk.transform = stub("skimage.transform", resize=lambda img, size, anti_aliasing=True: mine.downscale(img, size))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from nn.datasets import generators as ref
    return ref


def _same_npz(a, b):
    da, db = np.load(a), np.load(b)
    assert sorted(da.files) == sorted(db.files) == ["test_x", "train_x", "valid_x"]
    for k in da.files:
        assert da[k].dtype == db[k].dtype and da[k].shape == db[k].shape and np.array_equal(da[k], db[k]), k
    return da


@needs_ref
def test_spring_balls_matches_reference(tmp_path):
    ref = _reference_generators()
    kw = dict(img_size=[32, 32], radius=2, dt=0.3, k=4, equil=6, vx0_max=8, vy0_max=8, color=True)   # spring_color
    np.random.seed(3)
    ref.generate_spring_balls_dataset(str(tmp_path / "ref.npz"), 5, 2, 2, 12, **kw)
    np.random.seed(3)
    mine.generate_spring_balls_dataset(str(tmp_path / "mine.npz"), 5, 2, 2, 12, **kw)
    d = _same_npz(tmp_path / "ref.npz", tmp_path / "mine.npz")
    assert d["train_x"].shape == (5, 12, 32, 32, 3) and d["train_x"].dtype == np.uint8
    assert d["train_x"][..., 2].max() == 255 and d["train_x"][..., 1].max() == 255 and d["train_x"][..., 0].max() == 0
    try:
        import PIL  # noqa: F401
        assert os.path.exists(tmp_path / "mine_samples.jpg")            # the reference's '<dest>_samples.jpg'
    except ImportError:
        pass


@needs_ref
def test_three_body_matches_reference(tmp_path):
    ref = _reference_generators()
    kw = dict(img_size=[36, 36], radius=2, dt=0.5, g=60, m=1.0, vx0_max=2, vy0_max=2, color=True)    # 3bp_color
    np.random.seed(11)
    ref.generate_3_body_problem_dataset(str(tmp_path / "ref.npz"), 4, 1, 1, 20, **kw)
    np.random.seed(11)
    mine.generate_3_body_problem_dataset(str(tmp_path / "mine.npz"), 4, 1, 1, 20, **kw)
    d = _same_npz(tmp_path / "ref.npz", tmp_path / "mine.npz")
    assert d["train_x"].shape == (4, 20, 36, 36, 3) and all(d["train_x"][..., c].max() == 255 for c in range(3))


@needs_ref
def test_falling_and_bouncing_sets_match_reference(tmp_path):
    ref = _reference_generators()
    np.random.seed(5)
    ref.generate_falling_bouncing_ball_dataset(str(tmp_path / "ref.npz"), 3, 1, 1, 10, img_size=[32, 32], radius=3, g=9.8,
                                               vx0_max=4.0, vy0_max=4.0)
    np.random.seed(5)
    mine.generate_falling_bouncing_ball_dataset(str(tmp_path / "mine.npz"), 3, 1, 1, 10, img_size=[32, 32], radius=3, g=9.8,
                                                vx0_max=4.0, vy0_max=4.0)
    _same_npz(tmp_path / "ref.npz", tmp_path / "mine.npz")
    ref.generate_bouncing_ball_dataset(str(tmp_path / "refb.npz"), 6, 2, 2, 15, 8.0)
    mine.generate_bouncing_ball_dataset(str(tmp_path / "mineb.npz"), 6, 2, 2, 15, 8.0)
    d = _same_npz(tmp_path / "refb.npz", tmp_path / "mineb.npz")
    assert d["train_x"].shape == (6, 15, 2)


def test_rasteriser_and_loader_contract(tmp_path):
    rr, cc = mine.disc_indices(10, 12, 4, (32, 32))
    assert len(rr) == 45 and ((rr - 10) ** 2 + (cc - 12) ** 2 < 16).all()             # lattice points strictly inside r = 4
    rr, cc = mine.disc_indices(1, 1, 5, (8, 8))
    assert rr.min() == 0 and cc.min() == 0                                            # clipped at the border
    img = np.zeros((40, 40, 3), np.float32)
    img[:20] = 1.0
    small = mine.downscale(img, (4, 4))
    assert small.shape == (4, 4, 3) and np.allclose(small[:2], 1.0) and np.allclose(small[2:], 0.0)
    # mnist-style set from caller-supplied glyphs, then through the iterators' reshape
    np.random.seed(0)
    glyph = np.zeros((22, 22), np.float32)
    glyph[4:18, 9:13] = 1.0
    mine.generate_spring_mnist_dataset(str(tmp_path / "m.npz"), 2, 1, 1, 6, digits=[glyph, glyph.T], img_size=[64, 64], k=2,
                                       equil=12, vx0_max=8, vy0_max=8)
    d = np.load(tmp_path / "m.npz")
    assert d["train_x"].shape == (2, 6, 64, 64, 3) and d["train_x"].dtype == np.uint8 and d["train_x"].max() == 255
    x = d["train_x"]
    as_loaded = x.astype(np.float32).reshape(x.shape[:2] + (3, 64, 64)) / 255          # iterators.py:57-64
    assert as_loaded.shape == (2, 6, 3, 64, 64) and 0.0 <= as_loaded.min() and as_loaded.max() == 1.0
