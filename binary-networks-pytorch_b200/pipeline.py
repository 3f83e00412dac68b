"""Host-to-logits serving pipeline: overlap the PCIe upload of batch i+1 with the forward of batch i.

``HostPipeline(model_or_engine, example_batch)`` owns two device input buffers and (optionally) one CUDA
graph per buffer.  ``submit(host_batch)`` enqueues  H2D copy (copy stream) -> forward (compute stream) ->
D2H of the logits into pinned memory, and returns immediately; ``results()`` yields the logits in order.
The upload of the next batch runs while the previous forward is still computing, so the steady-state cost
per batch is max(forward, upload) instead of their sum (a 154 MB fp32 batch of 256 images takes ~2.9 ms
over PCIe Gen5, a third of the forward).
"""
from collections import deque
from typing import Deque, Optional, Tuple

import torch


class HostPipeline:
    def __init__(self, engine: torch.nn.Module, example: torch.Tensor, device: Optional[torch.device] = None,
                 use_graphs: bool = True, post=None) -> None:
        self.engine = engine
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.post = post                                   # e.g. the logits all-gather of sharded inference
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)       # logits download off the compute stream
        self.bufs = [torch.empty(example.shape, dtype=example.dtype, device=self.device) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.read = [torch.cuda.Event() for _ in range(2)]     # slot k's logits have left the device
        self.graphs = [None, None]
        self.outs = [None, None]
        self.host_out = [None, None, None]      # three pinned result buffers: a retired view survives one more submit
        self.k = 0
        self.step = 0
        self.pending: Deque[Tuple[int, torch.cuda.Event]] = deque()
        with torch.no_grad():
            for k in range(2):
                self.bufs[k].copy_(example.to(self.device, non_blocking=True))
                with torch.cuda.stream(self.compute_stream):
                    out = self.engine(self.bufs[k])         # warm-up: packs weights, folds BN
                    self.compute_stream.synchronize()
                    if use_graphs:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=self.compute_stream):   # forward only, collectives outside
                            out = self.engine(self.bufs[k])
                        self.graphs[k] = g
                    self.outs[k] = out
                    final = self.post(out) if self.post is not None else out
                    self.compute_stream.synchronize()
                self.host_out[k] = torch.empty(final.shape, dtype=final.dtype).pin_memory()
                if k == 1:
                    self.host_out[2] = torch.empty(final.shape, dtype=final.dtype).pin_memory()
                self.done[k].record(self.compute_stream)
                self.read[k].record(self.compute_stream)
        torch.cuda.synchronize(self.device)

    @property
    def h2d_bytes(self) -> int:
        return self.bufs[0].numel() * self.bufs[0].element_size()

    @property
    def d2h_bytes(self) -> int:
        return self.host_out[0].numel() * self.host_out[0].element_size()

    def submit(self, host_batch: torch.Tensor):
        """Enqueue one batch (pinned host memory for a truly asynchronous copy).  When both slots are in flight
        the oldest batch is retired first and its logits are returned: a view of that slot's pinned buffer, valid
        until the next ``submit`` (copy it if it has to live longer; note that ``clone()`` of a pinned tensor
        allocates pinned memory, which is slow -- use ``torch.empty_like(t, pin_memory=False).copy_(t)``)."""
        k = self.k
        self.k ^= 1
        hk = self.step % 3
        self.step += 1
        retired = None
        if len(self.pending) == 2:                          # both slots in flight: retire the oldest first
            ko, ev = self.pending.popleft()
            ev.synchronize()
            retired = self.host_out[ko]
        self.copy_stream.wait_event(self.done[k])           # slot k's previous forward has consumed its input
        with torch.cuda.stream(self.copy_stream):
            self.bufs[k].copy_(host_batch, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.compute_stream.wait_event(self.ready[k])
        self.compute_stream.wait_event(self.read[k])        # the previous logits of this slot have been downloaded
        with torch.cuda.stream(self.compute_stream), torch.no_grad():
            if self.graphs[k] is not None:
                self.graphs[k].replay()
            else:
                self.outs[k] = self.engine(self.bufs[k])
            final = self.post(self.outs[k]) if self.post is not None else self.outs[k]
            self.done[k].record(self.compute_stream)
        # the download runs on its own stream: the next forward does not queue behind a PCIe round trip
        self.d2h_stream.wait_event(self.done[k])
        fin = torch.cuda.Event()
        with torch.cuda.stream(self.d2h_stream):
            final.record_stream(self.d2h_stream)
            self.host_out[hk].copy_(final, non_blocking=True)
            self.read[k].record(self.d2h_stream)
            fin.record(self.d2h_stream)
        self.pending.append((hk, fin))
        return retired

    def results(self):
        """Retire everything still in flight, oldest first; yields the logits of each batch (pinned views)."""
        while self.pending:
            k, ev = self.pending.popleft()
            ev.synchronize()
            yield self.host_out[k]

    def drain(self):
        """Wait for everything submitted; returns the logits of the last batch (pinned host tensor)."""
        last = None
        while self.pending:
            k, ev = self.pending.popleft()
            ev.synchronize()
            last = self.host_out[k]
        return last
