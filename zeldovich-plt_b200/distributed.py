"""Slab-decomposed runs: one process per GPU, ``torch.distributed`` for the plumbing.

Stage 1 (generation, x and z transforms) is sharded over y-row pairs, stage 2 (y transform and
emission) over z planes; between them every rank sends block r of its stage-1 buffer to rank r —
one all-to-all, the y<->z transpose the reference does through ``BlockArray``
(reference src/block_array.cpp:387-414, 466-504).  The layout math lives in ``csrc/zplt_slab.h``
(device + host); this module only moves the blocks.
"""
import ctypes as C

import torch
import torch.distributed as dist


def exchange_tensors(send: torch.Tensor, recv: torch.Tensor, group=None) -> None:
    """all-to-all of equal contiguous blocks: block r of ``send`` goes to rank r, ``recv`` is filled in
    rank order.  Works on any backend (nccl for device tensors, gloo on CPU)."""
    world = dist.get_world_size(group)
    assert send.numel() == recv.numel() and send.numel() % world == 0
    dist.all_to_all_single(recv, send, group=group)


class SlabWorkspace:
    """Torch-owned workspace of a slab context: [send | recv] halves, viewed as float64."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        nbytes = ctx.workspace_bytes()
        self.buf = torch.empty(nbytes // 8, dtype=torch.float64, device=device)
        ctx.set_workspace(self.buf.data_ptr(), nbytes)
        send_ptr, recv_ptr, self.bytes_per_peer = ctx.exchange_info()
        world = dist.get_world_size()
        half = self.bytes_per_peer * world // 8  # float64 elements of one slab buffer (the workspace may hold more behind them)
        assert send_ptr == self.buf.data_ptr() and recv_ptr == self.buf.data_ptr() + half * 8
        self.send = self.buf[:half]
        self.recv = self.buf[half:2 * half]

    def begin(self):
        """Nothing to wait for: the all-to-all below is ordered on the stream after this rank's emission."""

    def exchange(self, group=None):
        """Run the all-to-all on the current torch stream and mark the context as exchanged."""
        exchange_tensors(self.send, self.recv, group)
        self.ctx.exchange_done()


def plane_range(ppd: int, rank: int, world: int):
    """Global z planes [z0, z1) rank owns in stage 2."""
    n = ppd // world
    return rank * n, (rank + 1) * n


class PeerExchange:
    """Fused exchange: the z-pass kernel stores into the peers' stage-2 buffers over NVLink
    (CUDA IPC mapped peer memory).  ``torch.distributed`` only carries the 64-byte handles and
    the barrier that separates the writers from the readers."""

    def __init__(self, ctx, group=None):
        self.ctx, self.group = ctx, group
        world = dist.get_world_size(group)
        mine = ctx.ipc_export()
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        ctx.ipc_import(handles)

    def close(self):
        """Unmap the peers' buffers on every rank; only then may any rank free (or destroy the context that owns) its own."""
        self.ctx.synchronize()
        dist.barrier(group=self.group)
        self.ctx.ipc_close()
        dist.barrier(group=self.group)

    def begin(self):
        """Call before ctx.generate() of every step but the first: the peers' z pass of the coming step stores into the
        receive buffer this rank may still be emitting planes from (write-after-read), so everybody first finishes
        emitting, then everybody may generate."""
        self.ctx.synchronize()
        dist.barrier(group=self.group)

    def generate(self):
        """One whole stage 1 on this rank: begin(), the potential pass with its two barriers when ZD_f_NL != 0,
        ctx.generate(), exchange()."""
        self.begin()
        if self.ctx.cfg.f_NL != 0.0:
            self.ctx.potential_begin()
            self.ctx.synchronize()
            dist.barrier(group=self.group)
            self.ctx.potential_exchange()
            self.ctx.synchronize()
            dist.barrier(group=self.group)
        self.ctx.generate()
        self.exchange()

    def exchange(self):
        """Call after ctx.generate(): wait for this rank's peer stores, then for everybody else's."""
        self.ctx.synchronize()
        dist.barrier(group=self.group)
        self.ctx.exchange_done()
