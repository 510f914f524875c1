"""CUDA-graph capture of a whole hot-path step (forward + backward + flat gradient all-reduce).

A training step at per-GPU batch 8 is ~670 kernel launches of 5-250 us each; launched one by one from Python
the host side (ctypes call, TMA descriptor encoding, torch.empty) costs about as much as the GPU work of the
small kernels.  All otter_b200 entry points are capture-safe (no allocation, no synchronisation, TMA
descriptors are passed by value as __grid_constant__ kernel parameters), so the step is captured once and
replayed: `GraphedStep(fn, *static_inputs)`; refresh the static input tensors in place, then `.replay()`.
"""
import torch


class GraphedStep:
    def __init__(self, fn, *static_inputs, warmup=2):
        self.fn, self.inputs = fn, static_inputs
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up off the capture stream (lazy inits, func attributes)
            for _ in range(warmup):
                fn(*static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):   # NCCL watchdog threads may poll
            self.outputs = fn(*static_inputs)

    def replay(self):
        self.graph.replay()
        return self.outputs
