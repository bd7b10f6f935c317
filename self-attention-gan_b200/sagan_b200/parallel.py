"""Batch data parallelism of the SAGAN training step: one process per GPU, replica gradients SUMMED.

Replaces what `tf.distribute.MirroredStrategy` does implicitly in the reference
(/root/reference/sagan/main.py:91-98 strategy + dataset sharding, :190,:205 gradient all-reduce inside
`apply_gradients`, :216-229 loss reduction for reporting):

  * the global batch `batch_size * n_replicas` (main.py:358) is split evenly, remainder dropped
    (sagan/dataset.py:39);
  * every replica differentiates `mean(loss_elems) / global_batch` (main.py:184,201), so the replica gradients are
    SUMMED, not averaged;
  * weights, Adam state and spectral-norm `u` are replicated; G's BatchNorm statistics stay per replica (plain
    `BatchNormalization`, generator.py:10): no activation exchange.

The exchange itself is one all-reduce per network over its single flat fp32 gradient bucket (D 0.70 MB, G 4.91 MB at
church64).  The tensors may live on any device: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """[start, stop) of this replica's samples inside a global batch (even split, remainder dropped)."""
    per = global_batch // world
    return rank * per, (rank + 1) * per


class ReplicaGradientSum:
    def __init__(self, process_group=None):
        self.pg = process_group
        self.active = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if self.active else 1
        self.rank = dist.get_rank(process_group) if self.active else 0

    def global_batch(self, per_replica_batch):
        """main.py:358: global_batch_size = batch_size * len(gpu)."""
        return per_replica_batch * self.world

    def broadcast_(self, *flat_tensors, src=0):
        """Identical initial state on every replica (MirroredStrategy creates mirrored variables)."""
        if self.world > 1:
            for t in flat_tensors:
                dist.broadcast(t, src, group=self.pg)

    def sum_(self, flat_grads):
        """main.py:190,205: SUM over replicas, in place, on the tensor's current stream."""
        if self.world > 1:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.pg)
        return flat_grads

    def sum_losses_(self, loss_sums):
        """main.py:216-220: per-replica loss sums -> global sums (reporting only)."""
        if self.world > 1:
            dist.all_reduce(loss_sums, op=dist.ReduceOp.SUM, group=self.pg)
        return loss_sums
