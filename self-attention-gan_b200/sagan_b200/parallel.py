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

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.pg)

    def sum_losses_(self, loss_sums):
        """main.py:216-220: per-replica loss sums -> global sums (reporting only)."""
        if self.world > 1:
            dist.all_reduce(loss_sums, op=dist.ReduceOp.SUM, group=self.pg)
        return loss_sums


class PeerAdam:
    """Replica gradient sum FUSED with Keras Adam over NVLink peer memory (csrc/dp.cu, sagan_dp_sum_adam): replica r
    reduces slice r of every replica's gradient bucket, updates it with its shard of the second-moment state and
    writes the new weights into every replica's parameter buffer.  One kernel per network per update, no NCCL on the
    data path.  The network's flat buffers must be symmetric-memory allocations (see `symmetric_allocator`)."""

    def __init__(self, net, opt, dp, loss_sums=None):
        """loss_sums: optional symmetric-memory [2] tensor {sum L_D, sum L_G}; when given, every `step()` also sums it
        over the replicas into `self.loss_global` (sagan/main.py:216-220) inside the same kernel."""
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.net, self.opt, self.dp = net, opt, dp
        lib = _lib.load()
        if dp.world > lib.sagan_dp_max_world():
            raise _lib.SaganError(f"PeerAdam supports at most {lib.sagan_dp_max_world()} replicas, got {dp.world}")
        dev = net.flat_params.device
        n = net.flat_params.numel()
        if n % (4 * dp.world):
            raise _lib.SaganError(f"flat bucket length {n} is not a multiple of 4 * world")
        group = dp.pg if dp.pg is not None else dist.group.WORLD
        self.flags = symm.empty(lib.sagan_dp_flag_bytes() // 4, dtype=torch.int32, device=dev)
        self.flags.zero_()
        hp = symm.rendezvous(net.flat_params, group)
        hg = symm.rendezvous(net.flat_grads, group)
        hf = symm.rendezvous(self.flags, group)
        if int(hp.buffer_ptrs[dp.rank]) != net.flat_params.data_ptr() or int(hg.buffer_ptrs[dp.rank]) != net.flat_grads.data_ptr():
            raise _lib.SaganError("symmetric-memory buffer pointer does not match the tensor (storage offset?)")
        self.peers = _lib.DpPeers()
        for q in range(dp.world):
            self.peers.grads[q] = int(hg.buffer_ptrs[q])
            self.peers.params[q] = int(hp.buffer_ptrs[q])
            self.peers.flags[q] = int(hf.buffer_ptrs[q])
        self._handles = (hp, hg, hf)
        self.loss_global = None
        if loss_sums is not None:
            hl = symm.rendezvous(loss_sums, group)
            self._loss_peers = (C.c_void_p * dp.world)(*[int(hl.buffer_ptrs[q]) for q in range(dp.world)])
            self.loss_global = torch.zeros(2, device=dev)
            self._handles += (hl,)
        self.n = n
        self.v_shard = torch.zeros(n // dp.world, device=dev)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._C = C
        torch.cuda.synchronize()
        dist.barrier(group=dp.pg)          # every replica's flag pad is zeroed before the first kernel signals

    def step(self):
        from . import _lib
        if self.loss_global is not None:
            _lib.check(_lib.load().sagan_dp_sum_adam_losses(
                self._C.byref(self.peers), self.dp.rank, self.dp.world, self.n, self.v_shard.data_ptr(),
                self.opt.hyper.data_ptr(), self.epoch.data_ptr(), self.status.data_ptr(), self._loss_peers,
                self.loss_global.data_ptr(), torch.cuda.current_stream().cuda_stream), "sagan_dp_sum_adam_losses")
            return
        _lib.check(_lib.load().sagan_dp_sum_adam(self._C.byref(self.peers), self.dp.rank, self.dp.world, self.n,
                                                 self.v_shard.data_ptr(), self.opt.hyper.data_ptr(),
                                                 self.epoch.data_ptr(), self.status.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream), "sagan_dp_sum_adam")

    def check(self):
        """Raises if a peer never arrived at a barrier (synchronises)."""
        if int(self.status.item()) != 0:
            from . import _lib
            raise _lib.SaganError("sagan_dp_sum_adam: a replica did not reach the exchange barrier (timed out)")


def symmetric_allocator():
    """Allocator for nets.set_flat_allocator: zeroed symmetric-memory buffers (peer-addressable over NVLink)."""
    import torch.distributed._symmetric_memory as symm

    def alloc(n, device):
        t = symm.empty(n, dtype=torch.float32, device=device)
        t.zero_()
        return t
    return alloc
