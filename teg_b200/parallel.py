"""Multi-GPU plumbing for the one exchange step of the path: summing reverse-mode parameter gradients.

Instances (pixels / parameter sets) are independent, so they are sharded across ranks in contiguous
blocks with no data-path collective; the value image stays sharded.  Only the
per-parameter gradient vector [P] (P = 6..2304 numbers) is combined, with ONE all-reduce per step.
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) carries it; the C ABI offers the
same operation without torch (tegb_nccl_comm_create / tegb_allreduce_sum).

On one node the collective library can be left out of the step altogether: ``PeerGroup`` maps a small mailbox of
every rank into every peer (CUDA IPC over NVLink / NVSwitch) and the runtime's last column-sum kernel writes the local
sums straight into all mailboxes and adds the ranks' contributions in rank order (csrc/tegb200.cu,
``colsum_stage2_peer``): no extra launch, a few microseconds instead of a collective's latency, and the same bits on
every rank.  ``torch.distributed`` is used once, to gather the 64-byte IPC handles.
"""
from __future__ import annotations

from typing import Tuple


def shard_rows(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of rows owned by ``rank``: sizes differ by at most one, earlier ranks get the extra."""
    assert world_size >= 1 and 0 <= rank < world_size and n_rows >= 0
    base, extra = divmod(n_rows, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_instances(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    return shard_rows(n, world_size, rank)


def allreduce_gradients(grad_sum, group=None):
    """In-place sum of the local per-parameter gradient sums over all ranks (one collective)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad_sum, op=dist.ReduceOp.SUM, group=group)
    return grad_sum


class PeerGroup:
    """Mailboxes for the fused column-sum + all-reduce over NVLink peer memory (one node, <= 16 ranks).

    ``attach(program)``: ``reduce`` launches of that compiled program return the sum over all ranks.
    ``allreduce_(tensor)``: stand-alone in-place all-reduce of a small CUDA vector.
    Every rank must make the same sequence of reducing calls.  ``close()`` synchronises the ranks first."""

    def __init__(self, device: int, max_values: int = 64, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import runtime
        assert dist.is_initialized(), 'PeerGroup needs an initialised torch.distributed process group'
        self._L = runtime.lib()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        self.handle = ctypes.c_void_p()
        mine = (ctypes.c_ubyte * 64)()
        # Every step that can fail locally (allocation, IPC export, IPC import) is followed by an agreement between the
        # ranks, so that a failure anywhere raises on EVERY rank instead of leaving the others inside a collective.
        rc = self._L.tegb_peer_create(self.rank, self.world, int(max_values), int(device),
                                      ctypes.byref(self.handle), ctypes.cast(mine, ctypes.c_void_p))
        why = '' if rc == 0 else runtime.last_error()
        on_gpu = dist.get_backend(group) == 'nccl'
        dev = torch.device('cuda', device) if on_gpu else torch.device('cpu')
        t = torch.tensor(list(bytes(mine)) + [1 if rc == 0 else 0], dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(gathered, t, group=group)
        rows = [bytes(g.cpu().tolist()) for g in gathered]
        ok = all(r[64] == 1 for r in rows)
        if ok:
            blob = b''.join(r[:64] for r in rows)
            self._blob = ctypes.create_string_buffer(blob, len(blob))
            rc = self._L.tegb_peer_connect(self.handle, ctypes.cast(self._blob, ctypes.c_void_p))
            why = '' if rc == 0 else runtime.last_error()
        flag = torch.tensor([1 if (ok and rc == 0) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)     # also: every mailbox is mapped before anyone writes
        if int(flag.item()) == 0:
            if self.handle:
                self._L.tegb_peer_destroy(self.handle)
                self.handle = None
            raise RuntimeError('peer mailboxes unavailable on some rank (CUDA IPC between the GPUs of this node is '
                               f'required){": " + why if why else ""}')

    def attach(self, program) -> None:
        """program: a ``runtime.Program`` (``CUDA_EvalMode.compiled(...)[1]``, ``PreparedRender.prog``)"""
        from . import runtime
        runtime.check(self._L.tegb_program_set_peer(program.handle, self.handle))

    def attach_groups(self, program, n_groups_total: int, local_of=None) -> None:
        """Grouped reduce programs (``CUDA_EvalMode.grouped_program(..., reduce=True)[1]``): their launches then
        deliver the all-reduced ``[n_groups_total, k]`` sums -- every row, whatever group range a rank launched (the
        groups it did not cover contribute zero).  ``max_values`` of this PeerGroup must be >= n_groups_total * k."""
        import ctypes
        from . import runtime
        if local_of is not None:
            # int32 CUDA tensor [n_groups_total]: this rank's group index of every global group, -1 where it holds none
            assert local_of.is_cuda and local_of.is_contiguous() and local_of.numel() == n_groups_total
            self._maps = getattr(self, '_maps', []) + [local_of]          # keep the device array alive
        ptr = ctypes.c_void_p(local_of.data_ptr()) if local_of is not None else None
        runtime.check(self._L.tegb_program_set_peer_groups(program.handle, self.handle, int(n_groups_total), ptr))

    def detach(self, program) -> None:
        from . import runtime
        runtime.check(self._L.tegb_program_set_peer(program.handle, None))

    def allreduce_(self, tensor, stream: int = 0):
        import ctypes
        from . import runtime
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.element_size() in (4, 8)
        runtime.check(self._L.tegb_peer_allreduce_sum(self.handle, ctypes.c_void_p(tensor.data_ptr()),
                                                      int(tensor.numel()), int(tensor.element_size()),
                                                      ctypes.c_void_p(stream)))
        return tensor

    def check(self) -> None:
        """Raises if an exchange timed out waiting for a peer (synchronises the device).  A time-out
        (``TEGB200_PEER_TIMEOUT_S``, default 120 s) also poisons the sums of that call with NaN and makes every later
        launch on this group raise ``TegbError``, so a missed ``check()`` cannot hide it; call it on every rank
        before trusting results that were produced without a later launch."""
        if self._L.tegb_peer_error(self.handle):
            raise RuntimeError('peer all-reduce: a rank did not arrive within the time-out')

    def close(self) -> None:
        import torch.distributed as dist
        if self.handle:
            self.check()
            dist.barrier(group=self.group)
            self._L.tegb_peer_destroy(self.handle)
            self.handle = None
