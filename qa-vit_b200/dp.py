"""Data-parallel gradient all-reduce (new functionality: the reference is single-GPU).

One process per GPU; gradients live in FusedAdamW's flat fp32 buffer, cut into a few contiguous buckets in
*reverse registration order* (head / stage4 first ... patch_embed / cnn_stem last, the order autograd produces them).
Each bucket is all-reduced (mean) on a side stream as soon as autograd has accumulated its last gradient
(post-accumulate-grad hooks), overlapping the remaining backward.  The only collective on the path.
The GlobalTokenBank state (global_k / global_v are parameters mutated inside forward) is averaged with the last
bucket so replicas do not drift (SURVEY.md hard part 2, option b)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

import ctypes as C

from ._lib import check, lib


def native_comm(group=None):
    """ncclComm_t of the process group's NCCL backend on the current device (ProcessGroupNCCL._comm_ptr), or None when the group is
    not NCCL / the communicator is not created yet / this torch build does not expose it.  The C ABI's qavit_dp_* take it."""
    try:
        if not (dist.is_initialized() and torch.cuda.is_available() and lib.qavit_dp_available()):
            return None
        pg = group if group is not None else dist.distributed_c10d._get_default_group()
        be = pg._get_backend(torch.device("cuda", torch.cuda.current_device()))
        ptr = be._comm_ptr()
        return int(ptr) if ptr else None
    except Exception:
        return None


def native_allreduce_sum(comm: int, tensors) -> None:
    """qavit_dp_allreduce_sum on the current stream: one NCCL group over the given contiguous fp32 CUDA tensors, in place."""
    n = len(tensors)
    bufs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    counts = (C.c_size_t * n)(*[t.numel() for t in tensors])
    check(lib.qavit_dp_allreduce_sum(comm, bufs, counts, n, torch.cuda.current_stream().cuda_stream))


class GradAllReducer:
    def __init__(self, opt, n_buckets: int = 4, bank_params: Optional[List[torch.nn.Parameter]] = None, group=None,
                 overlap: bool = True):
        """overlap=False: no autograd hooks; the caller invokes reduce_flat() once after backward (the form used when
        forward + backward are replayed as a CUDA graph: one 26 MB all-reduce is ~0.2 % of the step, so overlapping it
        buys nothing there and the collective stays outside the captured graphs)."""
        self.opt, self.group = opt, group
        self.native = True          # reduce_flat() goes through the C ABI (qavit_dp_allreduce_sum) when the group's ncclComm_t is available
        self._comm = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        params = opt.param_groups[0]["params"]
        offs = opt.seg_off.tolist()
        total = offs[-1]
        # bucket boundaries on parameter boundaries, roughly equal bytes, indexed from the END of the buffer
        target = total / n_buckets
        bounds, acc = [len(params)], 0
        for i in range(len(params) - 1, -1, -1):
            acc += offs[i + 1] - offs[i]
            if acc >= target and i > 0 and len(bounds) < n_buckets:
                bounds.append(i)
                acc = 0
        bounds.append(0)
        self.buckets = []                      # (first_param, last_param_exclusive, flat slice)
        for hi, lo in zip(bounds[:-1], bounds[1:]):
            if hi > lo:
                self.buckets.append((lo, hi, opt.flat_g[offs[lo]:offs[hi]]))
        self.bank_params = bank_params or []
        nb = sum(p.numel() for p in self.bank_params)
        if nb > getattr(opt, "tail", 0):
            raise ValueError(f"GradAllReducer: the bank state ({nb} floats) rides in the tail of the optimizer's gradient buffer; "
                             f"build FusedAdamW(..., tail_elems>={nb})")
        self._tail = opt._flat_g_full[total:total + nb]
        self._full = opt._flat_g_full[:total + nb]
        self._producers = []
        opt.grad_prescale = 1.0 / self.world        # buffers hold SUMS over ranks; the mean is folded into opt.clip()
        self._pending = [0] * len(self.buckets)
        self._param_bucket = {}
        for bi, (lo, hi, _) in enumerate(self.buckets):
            for i in range(lo, hi):
                self._param_bucket[i] = bi
        self._stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._handles = []
        self._hooks = []
        self._pindex = {id(p): i for i, p in enumerate(params)}
        if self.world > 1 and overlap:
            for i, p in enumerate(params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
            try:        # parameters whose gradients the native kernels accumulate in place announce themselves here
                from . import functional as QF
                QF.register_grad_ready_callback(self._on_native_ready)
            except Exception:      # pragma: no cover  (CPU-only unit tests of the bucket logic)
                pass
        self.reset()

    def reset(self):
        """Call before every backward: every bucket waits for all of its parameters that will get a gradient."""
        if not self.opt._have_flags and hasattr(self.opt, "_sync_flags"):
            self.opt._sync_flags()               # name-derived has-gradient mask: without it a bucket would wait for parameters
        flags = self.opt._flags_host if self.opt._have_flags else None      # that never report and only leave in finish()
        for bi, (lo, hi, _) in enumerate(self.buckets):
            self._pending[bi] = sum(1 for i in range(lo, hi) if flags is None or int(flags[i]) & 1)
        self._handles = []
        self._done = set()                       # parameter indices already counted this backward
        self._launched = set()
        self._multi = {id(p) for p in self.bank_params}
        self._flags = flags

    def _mark_ready(self, i):
        """Parameter i's gradient kernels are enqueued.  Two sources report -- autograd's post-accumulate hook and the native
        kernels' callback (gradients accumulated in place into the flat buffer) -- and BOTH may fire for the same parameter:
        count each parameter once (counting twice launched every bucket at its half-way point: 10 % wrong sums, r2 debug)."""
        if i in self._done:
            return
        self._done.add(i)
        if self._flags is not None and not (int(self._flags[i]) & 1):
            return                               # not part of the bucket's count
        bi = self._param_bucket[i]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and bi not in self._launched:
            self._launch(bi)

    def _make_hook(self, i):
        def hook(p):
            if id(p) in self._multi:             # the bank is written by every block: it leaves in finish()
                return
            self._mark_ready(i)
        return hook

    def _on_native_ready(self, params):
        for p in params:
            i = self._pindex.get(id(p))
            if i is None or id(p) in self._multi:
                continue
            self._mark_ready(i)

    def add_producer_stream(self, stream):
        """A stream besides the current one on which gradient kernels run (HQAViT's lateral path runs its backward on a
        side stream): every bucket launch waits for it too, so a bucket mixing main-stream and side-stream parameters
        is never reduced before both producers are done."""
        if stream is not None and stream not in self._producers:
            self._producers.append(stream)

    def _wait_producers(self):
        self._stream.wait_stream(torch.cuda.current_stream())
        for ps in self._producers:
            self._stream.wait_stream(ps)

    def _launch(self, bi):
        """SUM all-reduce of one bucket; the 1 / world of the mean is folded into the optimizer's clip pass."""
        buf = self.buckets[bi][2]
        self._launched.add(bi)
        if self._stream is not None:
            self._wait_producers()
            with torch.cuda.stream(self._stream):
                self._handles.append(dist.all_reduce(buf, group=self.group, async_op=True))
        else:
            self._handles.append(dist.all_reduce(buf, group=self.group, async_op=True))

    def _bank_to_tail(self):
        o = 0
        for p in self.bank_params:
            n = p.numel()
            if p.is_cuda:
                check(lib.qavit_scaled_copy(p.data.data_ptr(), 1.0, n, self._tail[o:o + n].data_ptr(), torch.cuda.current_stream().cuda_stream))
            else:
                self._tail[o:o + n].copy_(p.data.reshape(-1))
            o += n

    def _tail_to_bank(self):
        o = 0
        for p in self.bank_params:
            n = p.numel()
            if p.is_cuda:
                check(lib.qavit_scaled_copy(self._tail[o:o + n].data_ptr(), 1.0 / self.world, n, p.data.data_ptr(), torch.cuda.current_stream().cuda_stream))
            else:
                p.data.copy_((self._tail[o:o + n] / self.world).view(p.shape))
            o += n

    @torch.no_grad()
    def reduce_flat(self):
        """ONE all-reduce (sum) of the flat gradient buffer with the GlobalTokenBank state riding in its tail, on the current
        stream; 1 / world is applied to the gradients inside the clip pass and to the bank on its way back."""
        if self.world == 1:
            return
        if self.bank_params:
            self._bank_to_tail()
        if self.native and self._comm is None and self._full.is_cuda:
            self._comm = native_comm(self.group)           # exists once the group has run its first collective
        if self.native and self._comm is not None and self._full.is_cuda:
            native_allreduce_sum(self._comm, [self._full])  # the C ABI's qavit_dp_allreduce_sum on the current stream
        else:
            dist.all_reduce(self._full, group=self.group)
        if self.bank_params:
            self._tail_to_bank()

    def finish(self):
        """After backward: flush buckets whose hooks did not all fire, average the bank state, join the side stream."""
        if self.world == 1:
            return
        for bi in range(len(self.buckets)):
            if bi not in self._launched:
                self._pending[bi] = 0
                self._launch(bi)
        if self.bank_params:
            self._bank_to_tail()
            if self._stream is not None:
                self._wait_producers()
                with torch.cuda.stream(self._stream):
                    self._handles.append(dist.all_reduce(self._tail, group=self.group, async_op=True))
            else:
                self._handles.append(dist.all_reduce(self._tail, group=self.group, async_op=True))
        for h in self._handles:
            h.wait()
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        if self.bank_params:
            self._tail_to_bank()
        self._handles = []
