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
        for p in self.params:
            if p.dtype != dtype:
                raise ValueError("all bucketed parameters must share one dtype")
        self.attach()

    def attach(self):
        """(Re-)point every parameter's ``.grad`` at its slice of the flat buffer (autograd then accumulates in place)."""
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def check_attached(self):
        """Raise if a gradient no longer lives in the bucket -- ``optimizer.zero_grad()`` with its default
        ``set_to_none=True`` or a re-created Parameter (``RepZero*.__rep__`` replaces ``scaling``) detaches it, and the
        all-reduce would then average stale zeros while the ranks silently diverge.  Host-side pointer checks only."""
        off, esz = 0, self.flat.element_size()
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None or g.data_ptr() != self.flat.data_ptr() + off * esz or not g.is_contiguous():
                raise RuntimeError("FlatGradBucket: the gradient of parameter %d (shape %s) is not a view of the bucket any more; "
                                   "use bucket.zero_grad() instead of optimizer.zero_grad(set_to_none=True), and rebuild the "
                                   "bucket after re-creating parameters" % (i, tuple(p.shape)))
            off += p.numel()

    def zero_grad(self):
        """The bucket's replacement for ``optimizer.zero_grad()``: zero in place, keep the views attached."""
        self.flat.zero_()

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def all_reduce(self, async_op=False):
        """Average the bucket over ranks (sum-reduce then scale, as DDP does). No-op for world size 1."""
        self.check_attached()
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
