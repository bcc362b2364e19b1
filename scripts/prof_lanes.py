"""Where does the per-tensor headline spend its time?  Run under torchrun:
times GradientSync for subsets of the ResNet-50 tensors, lane counts and forced
algorithms.  (Experiment driver, prints one line per configuration.)"""
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
from faabric_b200.models import GradientSync, resnet50_grad_sizes  # noqa: E402


def main():
    dist = bench.Dist(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    comm_blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    comm, group = dist.make_comm(heapBytes=(512 << 20), stageBytes=(16 << 20), channels=8, maxBlocks=comm_blocks)

    class A:
        tuning = ""
        gpus = dist.world

    bench.load_tuning(comm, A, dist)
    sizes = resnet50_grad_sizes()
    small = [s for s in sizes if s * 4 <= 32 * 1024]
    mid = [s for s in sizes if 32 * 1024 < s * 4 < 128 * 1024]
    big = [s for s in sizes if s * 4 >= 128 * 1024]
    subsets = {"all": sizes, "small<=32K": small, "mid": mid, "big>=128K": big}
    for name, subset in subsets.items():
        for channels in (1, 4, 8):
            for algo in ("auto", "twoshot", "nvls"):
                if name == "small<=32K" and algo != "auto":
                    continue
                try:
                    sync = GradientSync(comm, subset, dtype=torch.int32, algo=algo, channels=channels)
                    sync.send.fill_(1)
                    ms = bench.timed(dist, sync.step, 20, 5)
                    err = comm.check_error()
                    sync.close()
                except Exception as e:  # unsupported combination
                    ms, err = float("nan"), str(e)[:60]
                if dist.rank == 0:
                    print(json.dumps({"subset": name, "tensors": len(subset), "MiB": round(sum(subset) * 4 / 2**20, 1),
                                      "channels": channels, "algo": algo, "blocks": comm_blocks,
                                      "ms": round(ms, 4), "us_per_tensor": round(ms * 1000 / max(1, len(subset)), 2),
                                      "err": err}), flush=True)
    dist.barrier()
    dist.close()


if __name__ == "__main__":
    main()
