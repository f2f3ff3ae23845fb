import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from tests.test_gpu_dist import _run_rank
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for case, variant in (("sim5", "default"), ("sim5", "union_skip")):
    out = _run_rank(rank, world, case, variant, dev)
    if rank == 0:
        print(case, variant, "keys:", sorted(k for k in out if k.startswith("grad/")), flush=True)
        g = np.load(os.path.join(ROOT, "tests", "golden", f"{case}.npz"))
        for k in out:
            if k.startswith("grad/"):
                ref = g[f"model/{variant}/{k}"]
                print("   ", k, float(np.abs(out[k] - ref).max() / max(np.abs(ref).max(), 1e-30)), flush=True)
dist.barrier(); dist.destroy_process_group()
