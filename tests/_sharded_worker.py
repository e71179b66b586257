"""Worker of test_multirank.py: one process of a gloo group runs the product's sharded pipeline
(stripenn_b200.compute.run_pipeline) with the oracle-backed FakeDevice — one virtual rank per
process (TorchComm) or `tpp` host threads per process (HybridComm)."""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch.distributed as dist  # noqa: E402

from fake_device import FakeDevice  # noqa: E402
from stripenn_b200 import compute as cp  # noqa: E402
from stripenn_b200 import synth  # noqa: E402


def main():
    out = sys.argv[1]
    tpp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    chrom = sys.argv[3] if len(sys.argv) > 3 else "all"
    dist.init_process_group("gloo")
    cm = synth.make_config("tiny2")
    pc = cp.TorchComm()
    if tpp == 1:
        dev = FakeDevice(cm, "KR")
        unf, fil = cp.run_stages(cm, norm="KR", chrom=chrom, maxpixel="0.95,0.98", device=dev, comm=pc)
        loaded = sorted(dev.loaded.keys())
    else:
        devs = [FakeDevice(cm, "KR") for _ in range(tpp)]
        res = cp.run_threads(lambda hc: cp.run_stages(cm, norm="KR", chrom=chrom, maxpixel="0.95,0.98",
                                                      device=devs[hc.tc.rank], comm=hc),
                             cp.HybridComm.group(tpp, pc))
        unf, fil = res[0]
        for other in res[1:]:       # every thread of the process holds the same tables
            assert other[0].equals(unf) and other[1].equals(fil)
        loaded = sorted(set(k for d in devs for k in d.loaded))
    with open("%s.rank%d" % (out, dist.get_rank()), "wb") as f:
        pickle.dump((unf, fil, loaded), f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
