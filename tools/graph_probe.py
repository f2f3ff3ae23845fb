import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.conftest import load_golden
from tests.helpers import golden_graph, VARIANT_FLAGS
from oracle.params import make_state_dict
from pangnn_b200 import ops, setup
from pangnn_b200.gnn import AlternateGCN
from pangnn_b200.graphs import GraphedStep
DEV = "cuda:0"
g = load_golden("c2"); pw = float(g["model/default/pos_weight"]); graph = golden_graph(g, "default", device=DEV)
def build():
    setup.reset(); ops.clear_cache()
    m = AlternateGCN(DEV, None, False); m.load_state_dict(make_state_dict(64, 128, False, seed=1234)); return m.to(DEV)
for cap in (False, True):
    model = build(); opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=cap); out = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True); loss, _ = model.forward_loss(graph, pw); loss.backward(); opt.step(); out.append(loss.item())
    print("eager capturable=", cap, out)
model = build(); opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
step = GraphedStep(model, graph, opt, pw, warmup=2)
print("graph (after 2 eager):", [step().item() for _ in range(4)])
