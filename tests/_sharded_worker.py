"""Worker of test_multirank.py: rank r of a 2-process gloo group runs the product's sharded
pipeline (stripenn_b200.compute.run_pipeline) with the oracle-backed FakeDevice."""
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
    dist.init_process_group("gloo")
    cm = synth.make_config("tiny2")
    dev = FakeDevice(cm, "KR")
    unf, fil = cp.run_stages(cm, norm="KR", maxpixel="0.95,0.98", device=dev, comm=cp.TorchComm())
    with open("%s.rank%d" % (out, dist.get_rank()), "wb") as f:
        pickle.dump((unf, fil, sorted(dev.loaded.keys())), f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
