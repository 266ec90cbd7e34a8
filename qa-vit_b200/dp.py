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

from ._lib import check, lib


class GradAllReducer:
    def __init__(self, opt, n_buckets: int = 4, bank_params: Optional[List[torch.nn.Parameter]] = None, group=None,
                 overlap: bool = True):
        """overlap=False: no autograd hooks; the caller invokes reduce_flat() once after backward (the form used when
        forward + backward are replayed as a CUDA graph: one 26 MB all-reduce is ~0.2 % of the step, so overlapping it
        buys nothing there and the collective stays outside the captured graphs)."""
        self.opt, self.group = opt, group
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
        flags = self.opt._flags_host if self.opt._have_flags else None
        for bi, (lo, hi, _) in enumerate(self.buckets):
            self._pending[bi] = sum(1 for i in range(lo, hi) if flags is None or int(flags[i]) & 1)
        self._handles = []
        self._seen = set()
        self._multi = {id(p) for p in self.bank_params}

    def _make_hook(self, i):
        def hook(_p):
            bi = self._param_bucket[i]
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch(bi)
        return hook

    def _on_native_ready(self, params):
        for p in params:
            i = self._pindex.get(id(p))
            if i is None:
                continue
            bi = self._param_bucket[i]
            if p in self._seen:            # shared parameters (the bank) are reported by every block: count once,
                continue                   # when the LAST user has run -- handled by finish() for those
            self._seen.add(p)
            if id(p) in self._multi:
                continue
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch(bi)

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
        dist.all_reduce(self._full, group=self.group)
        if self.bank_params:
            self._tail_to_bank()

    def finish(self):
        """After backward: flush buckets whose hooks did not all fire, average the bank state, join the side stream."""
        if self.world == 1:
            return
        for bi in range(len(self.buckets)):
            if self._pending[bi] > 0:
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
