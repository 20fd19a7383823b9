"""Data-parallel plumbing for ZiRa fine-tuning: ONE flat all-reduce of the trainable branch gradients.

The reference wraps the whole model in DDP *before* freezing (train_multidatasets.py:406 vs :239-246), with
``find_unused_parameters=True``, so every step walks the autograd graph and reduces buckets covering all
~172 M parameters although only the ZiRa branches (4.6 M parameters, 18.5 MB) train.  Here the gradients of
exactly the trainable parameters live as views into one contiguous buffer, reduced by a single NCCL
all-reduce (NVLS/tree over NVSwitch: the payload is latency-bound), issued after backward.  The path itself
does not shard below an image: replicas plus this one exchange is the whole multi-GPU story.
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    def __init__(self, params, world_size=None, dtype=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket needs at least one trainable parameter")
        dev = self.params[0].device
        dtype = dtype or self.params[0].dtype
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.dtype != dtype:
                raise ValueError("all bucketed parameters must share one dtype")
            p.grad = self.flat[off:off + n].view_as(p)   # autograd accumulates in place into the view
            off += n

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def all_reduce(self, async_op=False):
        """Average the bucket over ranks (sum-reduce then scale, as DDP does). No-op for world size 1."""
        if self.world <= 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            return work
        self.flat.div_(self.world)
        return None

    def finish(self, work):
        if work is not None:
            work.wait()
            self.flat.div_(self.world)

    def zero(self):
        self.flat.zero_()
