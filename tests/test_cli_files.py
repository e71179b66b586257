"""Drop-in surface: the file-level drivers `compute()` / `getScore()` and the typer CLI
(reference src/stripenn/cli.py:8-58, stripenn.py:64-163, score.py:7-61): same flags, same
files written (stripenn.log, result_unfiltered.tsv, result_filtered.tsv, the score table),
same columns.  CPU part: flag names/defaults and the log/coordinate-file logic.  GPU part: the
commands end to end on .npz contact maps against the reference's golden tables."""
import os

import numpy as np
import pandas as pd
import pytest
from typer.testing import CliRunner

from stripenn_b200 import cli
from stripenn_b200 import compute as cp
from stripenn_b200 import score as sc
from stripenn_b200 import synth

RTOL = 1e-5
COLS = ['chr', 'pos1', 'pos2', 'chr2', 'pos3', 'pos4', 'length', 'width', 'Mean', 'maxpixel', 'pvalue', 'Stripiness']


def test_cli_flags_are_the_references():
    """cli.py:10-23 / :47-53 — option names and defaults verbatim."""
    r = CliRunner().invoke(cli.app, ["compute", "--help"])
    assert r.exit_code == 0
    for flag in ("--cool", "--out", "-o", "--norm", "--chrom", "-k", "--canny", "-c", "--minL", "-l", "--maxW", "-w",
                 "--maxpixel", "-m", "--numcores", "-n", "--pvalue", "-p", "--mask", "-s", "--bfilter", "-b", "--seed"):
        assert flag in r.output, flag
    r = CliRunner().invoke(cli.app, ["score", "--help"])
    assert r.exit_code == 0
    for flag in ("--cool", "--coord", "-c", "--norm", "--numcores", "-n", "--out", "-o", "--mask", "--seed"):
        assert flag in r.output, flag
    import inspect
    d = {k: v.default.default for k, v in inspect.signature(cli.execute).parameters.items()}
    assert (d["norm"], d["chrom"], d["canny"], d["minL"], d["maxW"], d["maxpixel"], d["pvalue"], d["mask"], d["slow"],
            d["bfilter"], d["seed"]) == ('KR', 'all', 2.0, 10, 8, '0.95,0.96,0.97,0.98,0.99', 0.1, '0', False, 3,
                                         123456789)


def test_log_file_and_outdir(tmp_path, monkeypatch):
    """stripenn.py:12-61: the log keys in order; an existing directory is emptied on 'Y'."""
    out = str(tmp_path / "res")
    cp.makeOutDir(out)
    assert os.path.isdir(out)
    open(os.path.join(out, "old.txt"), "w").write("x")
    monkeypatch.setattr("builtins.input", lambda: "Y")
    cp.makeOutDir(out)
    assert os.listdir(out) == []
    monkeypatch.setattr("builtins.input", lambda: "n")
    with pytest.raises(SystemExit):
        cp.makeOutDir(out)
    cp.addlog("a.mcool", out, "KR", "all", 2.0, 10, 8, "0.95", 8, 0.1, "0", 3)
    keys = [l.split(":")[0] for l in open(os.path.join(out, "stripenn.log"))]
    assert keys == ["cool", "out", "norm", "chrom", "canny", "minL", "maxW", "maxpixel", "num_cores", "pvalue", "mask",
                    "blur filter"]


def test_coordinate_file_header_sniffing(tmp_path):
    """score.py:11-22."""
    rows = "chrA\t5001\t30000\tchrA\t5001\t400000\nchrA\t905001\t930000\tchrA\t605001\t930000\n"
    a = tmp_path / "a.tsv"
    a.write_text(rows)
    b = tmp_path / "b.tsv"
    b.write_text("c\tx1\tx2\tc2\ty1\ty2\textra\n" + rows.replace("\n", "\tfoo\n"))
    ta, tb = sc.read_coordinates(str(a)), sc.read_coordinates(str(b))
    assert list(ta.columns) == ['chr', 'pos1', 'pos2', 'chr2', 'pos3', 'pos4']
    assert list(tb.columns) == ['chr', 'pos1', 'pos2', 'chr2', 'pos3', 'pos4', 'extra']
    assert ta['pos4'].tolist() == tb['pos4'].tolist() == [400000, 930000]


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_compute_command_writes_the_references_tables(tmp_path, tiny2_cm, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny2_compute.npz"))
    npz = str(tmp_path / "tiny2.npz")
    tiny2_cm.save_npz(npz)
    out = str(tmp_path / "out")
    r = CliRunner().invoke(cli.app, ["compute", "--cool", npz, "--out", out, "--norm", "KR", "-m", "0.95,0.98", "-n", "2"])
    assert r.exit_code == 0, r.output
    assert sorted(os.listdir(out)) == ["result_filtered.tsv", "result_unfiltered.tsv", "stripenn.log"]
    for name, pre in (("result_unfiltered.tsv", "unf_"), ("result_filtered.tsv", "fil_")):
        df = pd.read_csv(os.path.join(out, name), sep="\t", dtype={"chr": str, "chr2": str})
        assert list(df.columns) == COLS
        for k in ("pos1", "pos2", "pos3", "pos4", "length", "width"):
            assert df[k].tolist() == g[pre + k].tolist(), (name, k)
        assert df["chr"].tolist() == g[pre + "chr"].tolist() and df["maxpixel"].tolist() == g[pre + "maxpixel"].tolist()
        for k in ("Mean", "pvalue", "Stripiness"):
            np.testing.assert_allclose(df[k].to_numpy(float), g[pre + k], rtol=RTOL, equal_nan=True)
    # -k with one chromosome, --yes to reuse the directory
    r = CliRunner().invoke(cli.app, ["compute", "--cool", npz, "--out", out, "--norm", "KR", "-m", "0.95,0.98", "-k", "chrB",
                                     "--yes"])
    assert r.exit_code == 0, r.output
    df = pd.read_csv(os.path.join(out, "result_unfiltered.tsv"), sep="\t", dtype={"chr": str})
    assert set(df["chr"]) == {"chrB"}


@pytest.mark.gpu
def test_score_command_matches_reference(tmp_path, tiny2_cm, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny2_score.npz"))
    npz = str(tmp_path / "tiny2.npz")
    tiny2_cm.save_npz(npz)
    tab = synth.make_stripe_table(tiny2_cm, 160)
    coord = str(tmp_path / "coords.tsv")
    with open(coord, "w") as f:
        for i in range(len(tab["pos1"])):
            f.write("%s\t%d\t%d\t%s\t%d\t%d\n" % (tab["chr"][i], tab["pos1"][i], tab["pos2"][i], tab["chr"][i],
                                                   tab["pos3"][i], tab["pos4"][i]))
    out = str(tmp_path / "scored.tsv")
    r = CliRunner().invoke(cli.app, ["score", "--cool", npz, "--coord", coord, "--norm", "KR", "--out", out])
    assert r.exit_code == 0, r.output
    df = pd.read_csv(out, sep="\t", dtype={"chr": str, "chr2": str})
    assert list(df.columns) == ['chr', 'pos1', 'pos2', 'chr2', 'pos3', 'pos4', 'pvalue_added', 'Stripiness_added',
                                'O_Mean_added', 'O_Sum_added', 'O/E_Mean_added', 'O/E_Total_added']
    for k in ("pvalue_added", "Stripiness_added", "O_Mean_added", "O_Sum_added", "O/E_Mean_added", "O/E_Total_added"):
        np.testing.assert_allclose(df[k].to_numpy(float), g[k.replace("/", "_")].astype(float), rtol=RTOL, equal_nan=True)
