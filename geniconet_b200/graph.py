"""Whole-step CUDA graph: forward + loss + backward + gradient exchange + optimizer update as ONE graph launch.

At I5/B=36 a training step is ~300 kernel launches of a few to a few hundred microseconds each; issued one by one from
Python (autograd + ctypes) the host cannot keep the GPU busy.  The C-ABI library allocates nothing and keeps no device
state, so every call is capturable as is; torch's caching allocator gives the capture a private pool.

    step = GraphedStep(fn, (x_static, t_static))      # fn(x, t) -> loss tensor (reads the static buffers)
    x_static.copy_(batch_x); t_static.copy_(batch_t); loss = step()      # loss is a static device tensor

Rules for `fn`: no host synchronisation (.item(), .cpu()), optimizers built with capturable=True, shapes fixed.
"""
import torch


class GraphedStep:
    def __init__(self, fn, static_inputs, warmup=3):
        self.fn, self.inputs = fn, tuple(static_inputs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # allocator + lazy-init warm-up outside the capture
            for _ in range(max(int(warmup), 1)):
                out = fn(*self.inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        del out
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.output = fn(*self.inputs)

    def __call__(self):
        self.graph.replay()
        return self.output
