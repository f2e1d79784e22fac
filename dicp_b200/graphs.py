"""CUDA-graph replay of a whole ICP step (not in the reference).

A `sync_free` constant-iteration call of `ICP.icp` never waits for the GPU, and neither does its reverse
pass, so everything a training or tracking step does with it -- forward loop, loss, backward -- can be recorded once in
a CUDA graph and replayed on new data.  Below a few thousand pairs x points a step is bound by its ~130 kernel
launches; the replay removes them from the host (profiles/r02b_small_batches.md: 0.65 -> 0.34 ms for one 4096-point
pair, 10 iterations, forward + backward).

    icp = ICP(icp_type='pt2pl', max_iterations=10); icp.const_iter = icp.sync_free = True; icp.nn_search = 'grid'
    def step(source, target, T_init):
        res = icp.icp(source, target, T_init, trim_dist=2.0, loss_fn={"name": "cauchy", "metric": 0.5})
        g_src, = torch.autograd.grad(res["T"].sum(), [source])
        return res["T"], g_src
    graphed = GraphedStep(step, (source, target, T_init))       # captures with these tensors' shapes / dtypes / requires_grad
    T, g_src = graphed(new_source, new_target, new_T_init)      # copies the inputs in, replays, returns the (static) outputs
"""
import torch


class GraphedStep:
    """Records `fn(*inputs)` in a CUDA graph and replays it.  `fn` must be capturable: CUDA tensors in, CUDA tensors
    (or nested tuples / lists / dicts of them) out, no host synchronisation inside -- for `ICP.icp` that means
    `const_iter` and `sync_free` set and tensor (not list) clouds.  Gradients belong inside `fn` (`torch.autograd.grad`).
    The outputs of a call are views of buffers owned by the graph: they are overwritten by the next call."""

    def __init__(self, fn, example_inputs, warmup=2):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device")
        self._inputs = [t.detach().clone().requires_grad_(t.requires_grad) if isinstance(t, torch.Tensor) else t for t in example_inputs]
        dev = next(t.device for t in self._inputs if isinstance(t, torch.Tensor))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                  # warm-up outside the capture: allocator pools, lazy initialisation
            for _ in range(max(1, warmup)):
                fn(*self._inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._outputs = fn(*self._inputs)

    def __call__(self, *inputs):
        if len(inputs) != len(self._inputs):
            raise ValueError("GraphedStep was captured with %d inputs, got %d" % (len(self._inputs), len(inputs)))
        with torch.no_grad():
            for dst, src in zip(self._inputs, inputs):
                if isinstance(dst, torch.Tensor):
                    if src.shape != dst.shape or src.dtype != dst.dtype:
                        raise ValueError("GraphedStep inputs must keep the captured shapes and dtypes")
                    if src.data_ptr() != dst.data_ptr():
                        dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._outputs

    @property
    def static_inputs(self):
        """The graph's own input buffers (write into them directly to skip the copy of `__call__`)."""
        return self._inputs

    def replay(self):
        """Replay on whatever the static input buffers hold."""
        self._graph.replay()
        return self._outputs
