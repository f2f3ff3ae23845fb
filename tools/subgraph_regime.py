#!/usr/bin/env python3
"""The reference's actual training regime: per-group sub-graphs, -b 32.  Times the dataset build and the
per-batch step of the re-hosted loop on config 2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pangnn_b200 import setup, train
from pangnn_b200.data import DataLoader
bs = sys.argv[1] if len(sys.argv) > 1 else "32"
args = setup.parse(["--simulate_dataset", "10000", "2", "0.5", "10", "3", "--train", "-e", "1", "-b", bs, "-o", "/tmp/runs", "-m", "/tmp/none.pkl"])
t0 = time.time()
res = train.run(args, device="cuda:0")
print("total wall", round(time.time() - t0, 2), "s;", len(res["dataset"].train), "train graphs")
ds, model = res["dataset"], res["model"]
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
pw = float(ds.class_balance)
from pangnn_b200.data import DeviceLoader
for name, loader in (("host collate + H2D (DataLoader)", DataLoader(ds.train, batch_size=args.batch_size, shuffle=True, device="cuda:0", seed=0)),
                     ("device collate (DeviceLoader, pangnn_collate)", DeviceLoader(ds.train, batch_size=args.batch_size, shuffle=True, device="cuda:0", seed=0))):
    for rep in range(2):                                    # second epoch is the measurement
        torch.cuda.synchronize(); t0 = time.time(); nb = 0; ne = 0
        for batch in loader:
            opt.zero_grad(); loss, logits = model.forward_loss(batch, pw); loss.backward(); opt.step(); loss.item()
            nb += 1; ne += batch.y.numel()
        torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{name}: epoch of {nb} batches (-b {bs}): {dt:.2f} s = {dt / nb * 1e3:.2f} ms / batch, {ne / dt:.3e} scored edges/s, {ne / nb:.0f} edges / batch")
    torch.cuda.synchronize(); t0 = time.time()
    for batch in loader:
        pass
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"    loader alone: {dt / nb * 1e3:.3f} ms / batch")

# ---- the same regime replayed as one CUDA graph per size bucket (pangnn_b200.graphs.GraphedBatchStep)
from pangnn_b200.graphs import GraphedBatchStep
import gc
del loss, logits, batch, opt                                # no autograd graph of the eager loops may stay alive
gc.collect()
gopt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
stepper = GraphedBatchStep(model, gopt, pw)
loader = DeviceLoader(ds.train, batch_size=args.batch_size, shuffle=True, device="cuda:0", seed=0)
for rep in range(3):                                        # the first epochs capture the buckets
    torch.cuda.synchronize(); t0 = time.time(); nb = 0; ne = 0
    for packed, ids, ids_dev in loader.iter_ids():
        loss, logits = stepper.step_ids(packed, ids, ids_dev); loss.item()
        nb += 1; ne += logits.numel()
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"graphed (epoch {rep}): {nb} batches (-b {bs}): {dt / nb * 1e3:.2f} ms / batch, {ne / dt:.3e} scored edges/s; {stepper.stats}, {len(stepper.buckets)} buckets")
