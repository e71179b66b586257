"""CPU, world_size 2 over gloo: the sharded pipeline (LPT partition by chromosome, two
host-side exchanges for the background, final gather + reordering) returns on every rank
exactly the reference's tables for 'tiny2'."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_match_reference(tmp_path, golden_dir):
    out = str(tmp_path / "res")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_sharded_worker.py"), out]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    subprocess.run(cmd, check=True, timeout=900, env=env, cwd=ROOT)
    g = np.load(os.path.join(golden_dir, "tiny2_compute.npz"))
    loaded = []
    for r in (0, 1):
        unf, fil, mine = pickle.load(open("%s.rank%d" % (out, r), "rb"))
        loaded.append(mine)
        for df, pre in ((unf, "unf_"), (fil, "fil_")):
            assert len(df) == len(g[pre + "pos1"])
            for k in ("pos1", "pos2", "pos3", "pos4"):
                assert df[k].astype(np.int64).tolist() == g[pre + k].tolist(), (r, pre, k)
            assert df["chr"].astype(str).tolist() == g[pre + "chr"].tolist()
            assert df["maxpixel"].tolist() == g[pre + "maxpixel"].tolist()
            for k in ("Mean", "pvalue", "Stripiness"):
                np.testing.assert_allclose(df[k].to_numpy(dtype=float), g[pre + k], rtol=1e-9, equal_nan=True)
    # each rank touched only its own chromosome
    assert loaded[0] != loaded[1] and len(loaded[0]) == len(loaded[1]) == 1
