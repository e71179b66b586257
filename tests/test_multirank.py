"""CPU, torch.distributed over gloo: the sharded pipeline (plan_units, typed exchanges of the
availability counts / background columns / cut-chromosome candidates / final rows, re-ordering)
returns on every rank exactly the reference's tables for 'tiny2':
  * world 2, one chromosome per rank;
  * world 4, both chromosomes cut into window ranges (third exchange);
  * 2 processes x 2 host threads (HybridComm: the bench's strong-scaling layout)."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(tmp_path, nproc, tpp=1, chrom="all"):
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc,
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_sharded_worker.py"), out, str(tpp), chrom]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    subprocess.run(cmd, check=True, timeout=900, env=env, cwd=ROOT)
    return [pickle.load(open("%s.rank%d" % (out, r), "rb")) for r in range(nproc)]


def _check(results, g, chrom=None):
    for r, (unf, fil, _) in enumerate(results):
        for df, pre in ((unf, "unf_"), (fil, "fil_")):
            sel = np.ones(len(g[pre + "pos1"]), bool) if chrom is None else g[pre + "chr"] == chrom
            assert len(df) == int(sel.sum())
            for k in ("pos1", "pos2", "pos3", "pos4"):
                assert df[k].astype(np.int64).tolist() == g[pre + k][sel].tolist(), (r, pre, k)
            assert df["chr"].astype(str).tolist() == g[pre + "chr"][sel].tolist()
            assert df["maxpixel"].tolist() == g[pre + "maxpixel"][sel].tolist()
            for k in ("Mean", "pvalue", "Stripiness"):
                np.testing.assert_allclose(df[k].to_numpy(dtype=float), g[pre + k][sel], rtol=1e-9, equal_nan=True)


def test_two_ranks_match_reference(tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny2_compute.npz"))
    res = _run(tmp_path, 2)
    _check(res, g)
    # each rank touched only its own chromosome
    loaded = [r[2] for r in res]
    assert loaded[0] != loaded[1] and len(loaded[0]) == len(loaded[1]) == 1


def test_four_ranks_cut_chromosomes_match_reference(tmp_path, golden_dir):
    """World 4 on two chromosomes: plan_units cuts both into two window ranges, the pieces meet at
    the primaries (exchange 3) and the tables are still the reference's."""
    from stripenn_b200 import compute as cp
    from stripenn_b200 import synth
    cm = synth.make_config("tiny2")
    units = cp.plan_units(cm.chromnames, cm.chromsizes, cm.binsize, 4)
    assert sorted(u["npieces"] for u in units) == [2, 2, 2, 2] and sorted(u["rank"] for u in units) == [0, 1, 2, 3]
    g = np.load(os.path.join(golden_dir, "tiny2_compute.npz"))
    _check(_run(tmp_path, 4), g)


def test_two_processes_two_threads_match_reference(tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny2_compute.npz"))
    _check(_run(tmp_path, 2, tpp=2), g)


def test_selected_chromosome_uses_every_rank(tmp_path):
    """-k chrB on a two-chromosome file at world 2: the search units are planned over the SELECTED
    chromosome (cut in two), the unselected one still gets an owner for the background stage."""
    from stripenn_b200 import compute as cp
    from stripenn_b200 import synth
    cm = synth.make_config("tiny2")
    units = cp.plan_units(cm.chromnames, cm.chromsizes, cm.binsize, 2, selected=["chrB"])
    srch = [u for u in units if u["search"]]
    assert [u["chrom"] for u in srch] == ["chrB", "chrB"] and sorted(u["rank"] for u in srch) == [0, 1]
    assert [u["chrom"] for u in units if not u["search"]] == ["chrA"]
    res = _run(tmp_path, 2, chrom="chrB")
    assert res[0][0].equals(res[1][0]) and len(res[0][0]) > 0
    assert set(res[0][0]["chr"].astype(str)) == {"chrB"}
    # same rows as the single-rank run of the same selection
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_device import FakeDevice
    unf, fil = cp.run_stages(cm, norm="KR", chrom="chrB", maxpixel="0.95,0.98", device=FakeDevice(cm, "KR"))
    for k in ("pos1", "pos2", "pos3", "pos4", "maxpixel"):
        assert unf[k].tolist() == res[0][0][k].tolist()
    np.testing.assert_allclose(unf["pvalue"].to_numpy(float), res[0][0]["pvalue"].to_numpy(float), rtol=1e-12)
