"""CUDA-graph replay of a static-shape training step.

A whole-graph step over a small pan-genome (BASELINE.json configs[0..1]: the graph lives in L2) is
~90 kernel launches of a few microseconds each — launch-latency-bound, not bandwidth-bound.  When the
same graph is stepped repeatedly (whole-graph epochs, ``pangnn.py:167-238`` with one batch), the step is
captured once and replayed: one launch per step, no Python between kernels.
"""
import torch


class GraphedStep:
    """``loss = GraphedStep(model, graph, optimizer, pos_weight)()`` replays one fused
    forward + BCE loss + backward + optimizer step.  The graph tensors must not be reallocated; parameter
    and optimizer state are updated in place exactly as by the eager step.  ``optimizer`` must be built
    with ``capturable=True`` (Adam)."""

    def __init__(self, model, graph, optimizer, pos_weight, warmup=3):
        self.model, self.graph, self.opt, self.pw = model, graph, optimizer, float(pos_weight)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the capturing stream: lazy inits, CSR cache
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.cuda_graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.cuda_graph):
            self.loss = self._eager()

    def _eager(self):
        self.opt.zero_grad(set_to_none=True)
        loss, self.logits = self.model.forward_loss(self.graph, self.pw)
        loss.backward()
        self.opt.step()
        return loss

    def __call__(self):
        self.cuda_graph.replay()
        return self.loss
