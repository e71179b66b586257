"""`stripenn seeimage` (reference src/stripenn/seeimage.py:32-97).
CPU: the oracle restatement against the image arrays the UNMODIFIED reference hands to imshow
(tests/golden/seeimage.npz, made by make_golden_seeimage.py), the PNG writer, the CLI surface.
GPU: snn_heatmap through the C ABI against the same golden arrays, bit for bit."""
import os

import numpy as np
import pytest

from stripenn_b200 import seeimage as see


@pytest.fixture(scope="module")
def see_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "seeimage.npz"))


def _cases(g):
    for k in range(int(g["ncases"])):
        yield (k, str(g["k%d_map" % k]), str(g["k%d_position" % k]), str(g["k%d_maxpixel" % k]),
               str(g["k%d_norm" % k]), bool(g["k%d_slow" % k]))


def _cm(name, tiny_cm, tiny2_cm):
    return {"tiny": tiny_cm, "tiny2": tiny2_cm}[name]


# ------------------------------------------------------------------------------------------- CPU
def test_oracle_seeimage_matches_reference(tiny_cm, tiny2_cm, see_golden):
    from snn_oracle import driver
    g = see_golden
    for k, mp_, pos, mpx, norm, slow in _cases(g):
        imgs = driver.seeimage(_cm(mp_, tiny_cm, tiny2_cm), pos, mpx, norm, slow)
        assert len(imgs) == len(mpx.split(","))
        for i, (suffix, chan) in enumerate(imgs):
            assert "heat" + suffix == str(g["k%d_name%d" % (k, i)])
            ref = g["k%d_g%d" % (k, i)]
            assert chan.shape == ref.shape
            assert np.array_equal(chan, ref, equal_nan=True), (k, i)


def test_golden_covers_the_edge_cases(see_golden):
    g = see_golden
    assert np.isnan(g["k0_g0"]).any() and np.isnan(g["k2_g0"]).any()     # NaN-weight bins
    assert g["k1_g0"].shape[0] > 448                                       # wider than the band
    assert not np.array_equal(g["k0_g0"], g["k0_g1"], equal_nan=True)     # maxpixel matters


def test_write_png_round_trip(tmp_path):
    import cv2
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    p = str(tmp_path / "a.png")
    see.write_png(p, a)
    assert np.array_equal(cv2.imread(p)[:, :, ::-1], a)
    with pytest.raises(ValueError):
        see.write_png(p, a.astype(np.float64))


def test_parse_position():
    assert see.parse_position("chr1:135,010,000-136000000", 2 * 10**8) == ("chr1", 135010000, 136000000)
    assert see.parse_position("chrB", 7499923) == ("chrB", 0, 7499923)


def test_cli_has_the_reference_commands_and_flags():
    import typer.main
    from stripenn_b200 import cli
    cmd = typer.main.get_command(cli.app)
    assert {"compute", "seeimage", "score"} <= set(cmd.commands)
    opts = {o for p in cmd.commands["seeimage"].params for o in p.opts}
    assert {"--cool", "--position", "-p", "--maxpixel", "-m", "--out", "-o", "--norm", "-s", "--seed"} <= opts


def test_invalid_chromosome_exits(tiny_cm):
    with pytest.raises(SystemExit):
        see.window_images(tiny_cm, "nope:1-1000", "0.95", "None", device=object())


# ------------------------------------------------------------------------------------------- GPU
@pytest.fixture
def dev(device):
    for cid in list(device.loaded):
        device.unload_chrom(cid)
    yield device
    for cid in list(device.loaded):
        device.unload_chrom(cid)


@pytest.mark.gpu
def test_gpu_seeimage_matches_reference(dev, tiny_cm, tiny2_cm, see_golden):
    g = see_golden
    for k, mp_, pos, mpx, norm, slow in _cases(g):
        imgs = see.window_images(_cm(mp_, tiny_cm, tiny2_cm), pos, mpx, norm, slow, device=dev, verbose=False)
        for i, (suffix, chan, rgb) in enumerate(imgs):
            ref = g["k%d_g%d" % (k, i)]
            assert "heat" + suffix == str(g["k%d_name%d" % (k, i)])
            assert np.array_equal(chan, ref, equal_nan=True), (k, i)
            # imshow's quantisation of the float image: (255*v) truncated, NaN -> 0, red = 255
            want = np.where(np.isnan(ref), 0.0, 255.0 * ref).astype(np.uint8)
            assert np.all(rgb[:, :, 0] == 255)
            assert np.array_equal(rgb[:, :, 1], want) and np.array_equal(rgb[:, :, 2], want)
        for cid in list(dev.loaded):
            dev.unload_chrom(cid)


@pytest.mark.gpu
def test_gpu_seeimage_writes_the_reference_file_names(tmp_path, tiny2_cm, see_golden):
    import cv2
    p = str(tmp_path / "tiny2.npz")
    tiny2_cm.save_npz(p)
    pos = str(see_golden["k0_position"])
    assert see.seeimage(p, pos, "0.95,0.98", "KR", str(tmp_path / "heat"), False, 123456789) == 0
    for i in range(2):
        f = str(tmp_path / str(see_golden["k0_name%d" % i]))
        assert os.path.exists(f)
        img = cv2.imread(f)[:, :, ::-1]
        ref = see_golden["k0_g%d" % i]
        assert img.shape == ref.shape + (3,)
        assert np.array_equal(img[:, :, 1], np.where(np.isnan(ref), 0.0, 255.0 * ref).astype(np.uint8))


@pytest.mark.gpu
def test_gpu_heatmap_rejects_bad_windows(dev, tiny_cm):
    from stripenn_b200 import getstripe as gs
    from stripenn_b200._lib import StripennB200Error as SnnError
    lib = gs.DeviceMatrix(tiny_cm, False, dev)
    cid = lib.ensure("T1")
    with pytest.raises(SnnError):
        dev.heatmap(cid, 10, 5, 7.0)
    with pytest.raises(SnnError):
        dev.heatmap(cid, 0, tiny_cm.nbins("T1") + 1, 7.0)
    chan, rgb = dev.heatmap(cid, 5, 5, 7.0)
    assert chan.shape == (0, 0) and rgb.shape == (0, 0, 3)
