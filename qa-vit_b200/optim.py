"""Gradient clipping + AdamW on flat fp32 buffers (reference: HQAViT_CIFAR100.py:1413-1439, 1566-1586).

``FusedAdamW`` re-homes every parameter (and its gradient) as a view into one flat fp32 buffer -- each tensor
starting on a 4-float boundary -- so the per-parameter / global clip is 3 launches and the AdamW step is 1,
instead of torch's ~10 multi-tensor launches plus ~60 single-tensor norms with host syncs.  lr and beta1 are
per-step device scalars (OneCycleLR rewrites both every iteration): ``param_groups[0]['lr']`` /
``['betas']`` are read at each step, so torch LR schedulers drive it unchanged."""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

from ._lib import check, lib

# parameters the reference never trains: GlobalTokenBank.write runs under no_grad (H:296-321) and the branch `.norm` only
# feeds it, so autograd leaves their .grad at None and torch.optim.AdamW skips them, weight decay included
NO_GRAD_NAME_PARTS = ("swa.norm.", "msda.norm.", "cga.norm.", "write_norm.", "write_compression.", "write_gate.")
_HYPER_SLOTS = 4


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=1e-2, max_grad_norm: Optional[float] = None, per_param_clip: float = 0.1,
                 per_param_clip_names: Sequence[str] = ("cnn_stem", "dwconv"), tail_elems: int = 0):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        if not named:
            raise ValueError("no parameters")
        self.names = [n for n, _ in named]
        params = [p for _, p in named]
        dev = params[0].device
        # CPU parameters are accepted for the host-side logic only (state_dict round trips, bucket layout tests);
        # clip() / step() raise there: the arithmetic has no CPU path
        self._cuda = dev.type == "cuda"
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm, self.per_param_clip = max_grad_norm, per_param_clip
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        offs.append(off)
        self.total = off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        # the gradient buffer carries a TAIL after the last parameter: the data-parallel reducer parks non-gradient state there
        # (the GlobalTokenBank's global_k / global_v) so that one all-reduce moves gradients and bank together (dp.py)
        self.tail = int(tail_elems)
        self._flat_g_full = torch.zeros(off + self.tail, dtype=torch.float32, device=dev)
        self.flat_g = self._flat_g_full[:off]
        self.grad_prescale = 1.0          # set by the data-parallel reducer: the buffer then holds the SUM over ranks
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self._pviews: List[torch.Tensor] = []
        self._gviews: List[torch.Tensor] = []
        with torch.no_grad():
            for p, o in zip(params, offs):
                v = self.flat_p[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v                                  # parameter now lives in the flat buffer
                g = self.flat_g[o:o + p.numel()].view(p.shape)
                self._pviews.append(v)
                self._gviews.append(g)
        self.seg_off = torch.tensor(offs, dtype=torch.int64, device=dev)
        clip_bit = [2 if any(s in n for s in per_param_clip_names) else 0 for n in self.names]
        self._clip_bits = clip_bit
        self.seg_flags = torch.zeros(len(params), dtype=torch.int32, device=dev)
        pin = (lambda t: t.pin_memory()) if self._cuda else (lambda t: t)
        self._flags_host = pin(torch.zeros(len(params), dtype=torch.int32))
        self.norms = torch.zeros(len(params) + 2, dtype=torch.float32, device=dev)
        # per-step scalars travel through a RING of pinned slots, each guarded by an event recorded after its H2D copy: the
        # host may run up to _HYPER_SLOTS steps ahead of the GPU without rewriting a slot a pending copy still has to read
        self._hyper_ring = [pin(torch.zeros(12, dtype=torch.float32)) for _ in range(_HYPER_SLOTS)]
        self._hyper_evt = [None] * _HYPER_SLOTS
        self._hyper_i = 0
        self._hyper_host = self._hyper_ring[0]            # the slot written last (kept for introspection / tests)
        self.hyper = torch.zeros(12, dtype=torch.float32, device=dev)
        self.step_count = 0
        self._have_flags = False
        self._user_mask = False
        self._default_mask = [not any(s in n for s in NO_GRAD_NAME_PARTS) for n in self.names]
        self.flat_ema: Optional[torch.Tensor] = None      # enable_ema(): EMA of flat_p, updated inside the step kernel
        self.ema_decay = 0.0
        self._all_flags = torch.ones(len(params), dtype=torch.int32, device=dev)
        self._mon = torch.zeros(2 * len(params), dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ EMA folded into the step
    def enable_ema(self, decay: float = 0.9999) -> torch.Tensor:
        """Start tracking ``ema = decay * ema + (1 - decay) * p`` for every parameter inside the AdamW kernel
        (``ModelEMA.update``, HQAViT_CIFAR100.py:139-149).  Returns the flat EMA buffer (same layout as ``flat_p``)."""
        if self.flat_ema is None:
            self.flat_ema = self.flat_p.clone()
        self.ema_decay = float(decay)
        return self.flat_ema

    def ema_views(self):
        """{parameter name: view into the EMA buffer}."""
        if self.flat_ema is None:
            raise RuntimeError("enable_ema() first")
        offs = self.seg_off.tolist()
        return {n: self.flat_ema[o:o + p.numel()].view(p.shape) for n, p, o in zip(self.names, self.param_groups[0]["params"], offs)}

    # ------------------------------------------------------------------ gradient monitoring without host syncs
    @torch.no_grad()
    def monitor_norms(self):
        """Per-tensor gradient and parameter L2 norms (``GradientMonitor.log_gradients``, H:198-242) as two device
        vectors in ``self.names`` order, from two launches; call before ``clip()`` for the unclipped gradients."""
        n = len(self.names)
        if not self._have_flags:
            self._sync_flags()
        check(lib.qavit_segment_norms(self.flat_g.data_ptr(), self.seg_off.data_ptr(), self.seg_flags.data_ptr(), n,
                                      self._mon.data_ptr(), _stream()))
        check(lib.qavit_segment_norms(self.flat_p.data_ptr(), self.seg_off.data_ptr(), self._all_flags.data_ptr(), n,
                                      self._mon[n:].data_ptr(), _stream()))
        return self._mon[:n], self._mon[n:]

    def attach_grads(self):
        """Point every .grad at its slice of the flat gradient buffer (zeroed): autograd then accumulates in place
        and the DP all-reduce / clip / AdamW kernels see one contiguous tensor."""
        if self._cuda:
            check(lib.qavit_memset_zero(self.flat_g.data_ptr(), self.flat_g.numel() * 4, _stream()))
        else:
            self.flat_g.zero_()
        for p, g in zip(self.param_groups[0]["params"], self._gviews):
            p.grad = g
        self._scaled = False

    def zero_grad(self, set_to_none: bool = True):
        self.attach_grads()

    def _sync_flags(self, grads_present: Optional[Sequence[bool]] = None):
        """seg_flags bit 0 = "this parameter has a gradient".  Without an explicit mask: .grad is not None -- except that an
        ATTACHED gradient view (attach_grads() makes every .grad non-None) only counts for parameters the reference trains
        (NO_GRAD_NAME_PARTS: the bank's write_* and the branch .norm stay grad=None there and are skipped by AdamW)."""
        params = self.param_groups[0]["params"]
        for i, (p, gv) in enumerate(zip(params, self._gviews)):
            if grads_present is not None:
                has = grads_present[i]
            elif p.grad is None:
                has = False
            elif p.grad.data_ptr() == gv.data_ptr():
                has = self._default_mask[i]
            else:
                has = True
            self._flags_host[i] = (1 if has else 0) | self._clip_bits[i]
        self.seg_flags.copy_(self._flags_host, non_blocking=True)
        self._have_flags = True

    def set_grad_mask(self, has_grad: Sequence[bool]):
        """Explicit has-gradient mask (overrides the name-derived default used with attached gradients)."""
        self._user_mask = True
        self._sync_flags(list(has_grad))

    @torch.no_grad()
    def clip(self) -> torch.Tensor:
        """Per-parameter clip (names containing cnn_stem / dwconv, to 0.1) then global clip to max_grad_norm.
        Returns the device scalar of the global norm (what clip_grad_norm_ returns); no host sync."""
        if not self._cuda:
            raise RuntimeError("FusedAdamW.clip: CPU parameters -- the optimizer arithmetic has no CPU path")
        self._gather_foreign_grads()
        if not self._have_flags:
            self._sync_flags()
        check(lib.qavit_clip_grads_scaled(self.flat_g.data_ptr(), self.seg_off.data_ptr(), self.seg_flags.data_ptr(), len(self.names),
                                          float(self.per_param_clip), float(self.max_grad_norm if self.max_grad_norm else 3.0e38),
                                          float(self.grad_prescale), self.norms.data_ptr(), self.total, _stream()))
        self._scaled = True
        return self.norms[len(self.names)]

    def _gather_foreign_grads(self):
        """Gradients produced as fresh tensors (attach_grads() not used) are copied into the flat buffer."""
        for p, g in zip(self.param_groups[0]["params"], self._gviews):
            if p.grad is not None and p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g

    def push_hyper(self, lam: Optional[float] = None):
        """Advance the step counter and send this step's scalars (lr, betas, eps, wd, bias corrections) to the device buffer
        the step kernel reads: written into the next pinned ring slot (after waiting for the copy that last read that slot),
        copied H2D on the current stream, and the slot's event re-recorded.  Stream order protects the device buffer (the
        previous step's kernel has run before this copy lands); the ring protects the host buffer.  This is the only
        per-step host work when the step is graph-replayed (GraphedTrainStep calls it before every replay)."""
        grp = self.param_groups[0]
        self.step_count += 1
        b1, b2 = grp["betas"]
        t = self.step_count
        k = self._hyper_i
        self._hyper_i = (k + 1) % _HYPER_SLOTS
        if self._hyper_evt[k] is not None:
            self._hyper_evt[k].synchronize()
        h = self._hyper_ring[k]
        h[0], h[1], h[2], h[3], h[4] = grp["lr"], b1, b2, grp["eps"], grp["weight_decay"]
        h[5], h[6] = 1.0 - b1 ** t, 1.0 - b2 ** t
        h[7], h[8] = self.ema_decay, 1.0 - self.ema_decay
        h[9] = 1.0 if lam is None else float(lam)     # CutMix / MixUp weight of this step (GraphedTrainStep(mix=True) reads hyper[9])
        self._hyper_host = h
        self.hyper.copy_(h, non_blocking=True)
        if self._cuda:
            if self._hyper_evt[k] is None:
                self._hyper_evt[k] = torch.cuda.Event()
            self._hyper_evt[k].record()

    write_hyper = push_hyper      # former name

    @torch.no_grad()
    def step(self, closure=None):
        if not self._cuda:
            raise RuntimeError("FusedAdamW.step: CPU parameters -- the optimizer arithmetic has no CPU path")
        self._gather_foreign_grads()
        if not self._have_flags:
            self._sync_flags()
        if self.grad_prescale != 1.0 and not getattr(self, "_scaled", False):
            self.clip()               # data-parallel mean (1 / world) rides in the clip pass: it must run once per step
        if not torch.cuda.is_current_stream_capturing():
            self.push_hyper()         # a captured step reads self.hyper, which the replaying caller refreshes eagerly
        if self.flat_ema is not None:
            check(lib.qavit_adamw_ema_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                           self.exp_avg_sq.data_ptr(), self.flat_ema.data_ptr(), self.seg_off.data_ptr(),
                                           self.seg_flags.data_ptr(), len(self.names), self.hyper.data_ptr(), self.total, _stream()))
        else:
            check(lib.qavit_adamw_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                       self.seg_off.data_ptr(), self.seg_flags.data_ptr(), len(self.names), self.hyper.data_ptr(),
                                       self.total, _stream()))
        return None


    # ------------------------------------------------------------------ checkpoint compatibility with torch.optim.AdamW
    def state_dict(self):
        """torch.optim.AdamW's layout (the reference checkpoints ``optimizer.state_dict()``, H:1692 / 1728): per parameter
        index ``{'step', 'exp_avg', 'exp_avg_sq'}`` (copies out of the flat buffers) for every parameter that has been
        stepped, plus the param group.  A checkpoint written here loads into torch.optim.AdamW and vice versa."""
        offs = self.seg_off.tolist()
        params = self.param_groups[0]["params"]
        state = {}
        if self.step_count > 0:
            flags = self._flags_host.tolist()
            for i, p in enumerate(params):
                if not (int(flags[i]) & 1):
                    continue                      # never received a gradient: torch holds no state for it either
                o, n = offs[i], p.numel()
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(params)))
        out = {"state": state, "param_groups": [group]}
        if self.flat_ema is not None:
            out["qavit_ema"] = {"decay": self.ema_decay, "flat": self.flat_ema.clone()}
        return out

    def load_state_dict(self, sd):
        """Accepts the dict above or one written by torch.optim.AdamW over the same parameter list (same order)."""
        groups = sd["param_groups"]
        params = self.param_groups[0]["params"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(params):
            raise ValueError(f"loaded state dict holds {len(order)} parameters, this optimizer {len(params)}")
        g0 = groups[0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in g0:
                self.param_groups[0][k] = tuple(g0[k]) if k == "betas" else g0[k]
        offs = self.seg_off.tolist()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        pos = {pid: i for i, pid in enumerate(order)}
        for pid, st in sd["state"].items():
            i = pos[pid]
            o, n = offs[i], params[i].numel()
            if tuple(st["exp_avg"].shape) != tuple(params[i].shape):
                raise ValueError(f"state of parameter {self.names[i]}: shape {tuple(st['exp_avg'].shape)} != {tuple(params[i].shape)}")
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): one shared counter is kept here")
        self.step_count = steps.pop() if steps else 0
        if "qavit_ema" in sd:
            self.enable_ema(sd["qavit_ema"]["decay"]).copy_(sd["qavit_ema"]["flat"])
        self.state.clear()


def clip_grad_norms_(opt: FusedAdamW) -> torch.Tensor:
    return opt.clip()


class ModelEMA:
    """Drop-in for the reference's ``ModelEMA`` (HQAViT_CIFAR100.py:128-184): ``.ema`` is an eval-mode copy of the model,
    ``update(model)`` / ``set_decay`` / ``compute_distance`` keep their signatures.  With ``optimizer=FusedAdamW`` the
    parameter average is computed INSIDE the AdamW kernel (one extra fp32 stream instead of 815 ``mul_/add_`` launches):
    the copy's parameters are re-homed as views of the optimizer's flat EMA buffer and ``update()`` only mirrors the
    buffers (BatchNorm statistics, ``update_count``), as H:151-156 does."""

    def __init__(self, model: torch.nn.Module, decay: float = 0.9999, device=None, optimizer: Optional[FusedAdamW] = None):
        from copy import deepcopy
        if optimizer is None:
            raise RuntimeError("qavit_b200.ModelEMA needs optimizer=FusedAdamW (the average is fused into its step kernel)")
        self.ema = deepcopy(model).eval()
        self.decay = decay
        self.device = device
        self.opt = optimizer
        optimizer.enable_ema(decay)
        views = optimizer.ema_views()
        with torch.no_grad():
            for n, p in self.ema.named_parameters():
                p.requires_grad_(False)
                if n in views:
                    p.data = views[n]

    @torch.no_grad()
    def update(self, model: torch.nn.Module):
        ema_buffers = dict(self.ema.named_buffers())
        for name, buf in model.named_buffers():
            if name in ema_buffers:
                ema_buffers[name].copy_(buf)

    def set_decay(self, decay: float):
        self.decay = decay
        self.opt.ema_decay = float(decay)

    @torch.no_grad()
    def compute_distance(self, model: torch.nn.Module):
        param_dist = (self.opt.flat_ema - self.opt.flat_p).norm().item()
        ema_buffers = dict(self.ema.named_buffers())
        sq = 0.0
        for name, buf in model.named_buffers():
            if name in ema_buffers and buf.dtype == torch.float32:
                sq += (ema_buffers[name] - buf).norm().item() ** 2
        return param_dist, sq ** 0.5


class GradientMonitor:
    """Drop-in for the reference's ``GradientMonitor`` (HQAViT_CIFAR100.py:189-250): same attributes, same return values of
    ``log_gradients(model, detailed)`` and ``check_explosion(threshold)``.  With ``optimizer=FusedAdamW`` the ~1.6 k
    ``.norm().item()`` host syncs per call become two launches over the flat gradient / parameter buffers and ONE
    device-to-host copy; the per-tensor detail statistics are only computed (with torch ops) for the rare tensors the
    reference singles out (gradient norm > 10 or non-finite)."""

    def __init__(self, optimizer: Optional[FusedAdamW] = None):
        self.grad_norms = []
        self.param_norms = []
        self.layer_grad_history = {}
        self.explosion_count = 0
        self.opt = optimizer

    @torch.no_grad()
    def log_gradients(self, model, detailed=False):
        if self.opt is None:
            raise RuntimeError("qavit_b200.GradientMonitor needs optimizer=FusedAdamW (norms come from its flat buffers)")
        opt = self.opt
        gn_d, pn_d = opt.monitor_norms()
        both = torch.cat([gn_d, pn_d]).tolist()                # the only host sync of the call
        n = len(opt.names)
        gn, pn = both[:n], both[n:]
        has_grad = [(int(f) & 1) != 0 for f in opt._flags_host.tolist()]
        params = opt.param_groups[0]["params"]
        total_sq, param_sq = 0.0, 0.0
        grad_stats, layer_stats = {}, {}
        for i, name in enumerate(opt.names):
            if not has_grad[i]:
                continue                                       # the reference skips parameters whose .grad is None
            g_norm, p_norm = gn[i], pn[i]
            total_sq += g_norm ** 2
            param_sq += p_norm ** 2
            layer_name = ".".join(name.split(".")[:2])
            st = layer_stats.setdefault(layer_name, {"grad_norm": 0, "param_norm": 0, "count": 0})
            st["grad_norm"] += g_norm
            st["param_norm"] += p_norm
            st["count"] += 1
            if g_norm > 10.0 or not math.isfinite(g_norm):
                g = params[i].grad
                grad_stats[name] = {
                    "grad_norm": g_norm, "grad_mean": g.mean().item(), "grad_max": g.abs().max().item(),
                    "grad_min": g.abs().min().item(), "has_nan": torch.isnan(g).any().item(),
                    "has_inf": torch.isinf(g).any().item(), "param_norm": p_norm,
                }
        total_norm, param_norm = total_sq ** 0.5, param_sq ** 0.5
        self.grad_norms.append(total_norm)
        self.param_norms.append(param_norm)
        if detailed:
            for layer, st in layer_stats.items():
                self.layer_grad_history.setdefault(layer, []).append(st["grad_norm"] / max(st["count"], 1))
        return total_norm, param_norm, grad_stats, layer_stats

    def check_explosion(self, threshold=50.0):
        if len(self.grad_norms) > 0:
            is_exploding = self.grad_norms[-1] > threshold
            if is_exploding:
                self.explosion_count += 1
            return is_exploding
        return False
