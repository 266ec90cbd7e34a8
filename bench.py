#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: HQAViT CIFAR-100-shape training (bf16) images/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = forward + backward + per-parameter/global clip + AdamW (+ the bucketed gradient all-reduce when N > 1)
over one synthetic batch of `--batch` images per GPU (weak scaling).  Prints ONE JSON line (rank 0).

  value      whole-job images/sec, batch already resident in HBM, CUDA-event timed, max over ranks
  e2e        same step through the public API with the batch coming from pinned host memory every step
             (H2D inside the timed region) and the loss read back to the host every step
  roofline   the dominant kernel (tcgen05 projection GEMM, qkv shape) timed alone with CUDA events
  cpu_baseline  the reference training step on the host cores, on a bounded sample: the UNMODIFIED reference modules
             from baseline/_ref (kind "reference"; `python baseline/install_ref.py` copies them in the build container,
             the directory is git-ignored but travels to the GPU box), else the oracle port (kind "port")
  gpu_eager_baseline  the same live reference's eager CUDA step under bf16 autocast on this GPU (like-for-like baseline)
  --impl reference   times the CPU path alone
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train images/sec (HQAViT CIFAR-100 shape, fwd+bwd+clip+AdamW)"
WORKLOAD = "HQAViT CIFAR-100 32x32 training step, bf16, synthetic batch"

DEFAULT_DROPOUT, DEFAULT_DROP_PATH = 0.1, 0.1     # the reference defaults (H:56-57); --dropout 0 --drop-path 0 = the parity configuration

# BASELINE.json configs: [1] is the default bench line; [2] / [3] / [4] are selectable for the record (profiles/).
# fwd MFLOP / image = F_min of SURVEY.md 8d (blocks) + rest of the model.
WORKLOADS = {
    "hqavit_c100": dict(family="hqavit", img=32, classes=100, kw={}, ctor={}, fwd_mflop=185.4 + 1.22 + 199.7,
                        label="HQAViT CIFAR-100 32x32"),
    # HQAViTv2_CIFAR100.py: same blocks around the ConvNeXt-patchify stem (7 LayerScale blocks at 8 x 8: ~205 MFLOP instead of 98.9)
    "hqavitv2_c100": dict(family="hqavit", img=32, classes=100, kw={}, ctor=dict(variant="v2"), fwd_mflop=185.4 + 1.22 + 305.8,
                          label="HQAViTv2 CIFAR-100 32x32"),
    "qavitv2_c100": dict(family="qavit", img=32, classes=100, kw={}, ctor=dict(variant="v2"), fwd_mflop=713.4 + 1.22,
                         label="QAViTv2 CIFAR-100 32x32 (= QAViTV2_EXTREME model)"),
    "qavit_224": dict(family="qavit", img=224, classes=100, kw=dict(patch_size=16, window_size=7, dilation_factors=(1, 2, 3), linformer_k=64),
                      ctor=dict(variant="v2b"), fwd_mflop=2421.0 + 57.8, label="QAViTv2.py defaults 224x224 / patch 16 (196 tokens)"),
    # BASELINE config 4b: the CIFAR-100 model fed 96 x 96 images after adjust_positional_embedding + a 10-class head (STL-10 recipe,
    # HQAViT_Tiny_stl10.py:250-282, 405-412).  F_min: lateral path at 24 x 24 maps (9 x stem / LMFA, RRCV + SplitFusion unchanged) ~ 1271,
    # patch embed 10.6, blocks 185.4 + block 0's 576-token TokenLearner ~ 6
    "hqavit_stl96": dict(family="hqavit", img=96, built_img=32, classes=10, kw={}, ctor={}, fwd_mflop=1271.0 + 10.6 + 191.4,
                         label="HQAViT CIFAR-100 model at STL-10 96x96 (resized pos_embed)"),
    "hqavit_tinyin": dict(family="hqavit", img=64, classes=200, kw=dict(depth=12, num_learned_tokens=64),
                          ctor=dict(stage_depths=(2, 2, 6, 2), square_tokens=True), fwd_mflop=1296.6 + 803.0,
                          label="HQAViT TinyImageNet 64x64"),
}


def build_workload(Q, name, dropout=0.0, drop_path=0.0):
    w = WORKLOADS[name]
    common = dict(img_size=w.get("built_img", w["img"]), num_classes=w["classes"], dropout=dropout, drop_path=drop_path, **w["kw"])
    if w["family"] == "hqavit":
        model = Q.HQAViT(Q.HQAViTConfig(**common), **w["ctor"])
        if w.get("built_img", w["img"]) != w["img"]:
            Q.adjust_positional_embedding(model, w["img"])
        if dropout == 0.0:    # the parity configuration also silences SplitFusion's hard-coded Dropout(0.1) (H:930)
            for n in ("fuse2", "fuse3", "fuse4"):
                getattr(model, n).cat_mlp[3].p = 0.0
    else:
        model = Q.QAViT(Q.QAViTConfig(**common), **w["ctor"])
    return model, w


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def live_reference_available():
    try:
        from baseline import live_reference as LR
        return LR.available()
    except Exception:
        return False


def reference_step_rate(device, batch, steps, warmup, dropout, drop_path, threads=None, flash=None):
    """The UNMODIFIED reference (baseline/_ref, `python baseline/install_ref.py`) driven the way its own train loop drives
    it (HQAViT_CIFAR100.py:1401-1439): model.train(), CrossEntropyLoss(label_smoothing), per-parameter 0.1 clip of
    cnn_stem / dwconv, global 0.5 clip, torch.optim.AdamW, zero_grad.  device 'cpu': fp32 on the host cores (BASELINE.md
    section 4); a CUDA device: the reference's eager GPU path under torch.autocast(bfloat16) with its GradScaler, stock
    attention selection (flash_attn when importable and head_dim % 8 == 0, else SDPA; H:359-392).  Returns
    (images/sec, sec/step, info dict)."""
    from baseline import live_reference as LR
    if threads:
        torch.set_num_threads(threads)
    mod = LR.import_reference("HQAViT_CIFAR100", flash=flash)
    torch.manual_seed(42)
    model = mod.HQAViT(mod.HQAViTConfig(dropout=dropout, drop_path=drop_path))
    if dropout == 0.0:
        for n in ("fuse2", "fuse3", "fuse4"):
            getattr(model, n).cat_mlp[3].p = 0.0
    dev = torch.device(device)
    cuda = dev.type == "cuda"
    model = model.to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=6e-4, betas=(0.95, 0.999), weight_decay=0.06)
    crit = torch.nn.CrossEntropyLoss(label_smoothing=0.12)
    scaler = torch.amp.GradScaler("cuda", enabled=cuda)          # H:1590 GradScaler(enabled=use_amp), also with bf16
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, 3, 32, 32, generator=g).to(dev)
    y = torch.randint(0, 100, (batch,), generator=g).to(dev)
    clipped = [p for n, p in model.named_parameters() if "cnn_stem" in n or "dwconv" in n]

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=cuda):
            loss = crit(model(x), y)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        for p in clipped:
            if p.grad is not None:
                torch.nn.utils.clip_grad_norm_([p], max_norm=0.1)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        return loss

    ts = []
    for it in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = step()
        if cuda:
            torch.cuda.synchronize()
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    info = {"has_flash_attn": bool(getattr(mod, "HAS_FLASH_ATTN", False)), "loss": float(loss.item()),
            "threads": torch.get_num_threads()}
    del model, opt
    if cuda:
        torch.cuda.empty_cache()
    return batch / dt, dt, info


def cpu_step_rate(sample_b, steps, warmup, threads=None, dropout=0.0, drop_path=0.0):
    """The reference training step on the host cores, fp32, on `sample_b` images; (images/sec, sec/step, threads, kind).
    kind "reference": the live reference modules from baseline/_ref; kind "port": the oracle restatement (only when the
    reference copy is not installed).  dropout / drop_path > 0: masks drawn per step for every site."""
    if live_reference_available():
        rate, dt, info = reference_step_rate("cpu", sample_b, steps, warmup, dropout, drop_path, threads=threads)
        return rate, dt, info["threads"], "reference"
    from oracle import qavit_oracle as O
    if threads:
        torch.set_num_threads(threads)
    ocfg = O.OracleConfig(family="hqavit")
    sd = O.synthetic_state(ocfg)
    keys = O.trainable_keys(ocfg)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(sample_b, 3, 32, 32, generator=g)
    y = torch.randint(0, 100, (sample_b,), generator=g)
    state = {}
    ts = []
    mask_fn = None
    if dropout > 0 or drop_path > 0:
        nblk = sum(ocfg.stage_depths)

        def mask_fn(i, B, nt):
            if i == "pos":
                return (torch.rand(B, nt, ocfg.embed_dim) >= dropout).float() / (1.0 - dropout) if dropout > 0 else None
            return O.random_masks(ocfg, B, nt, dropout, drop_path * i / max(1, nblk - 1))     # H:1187 linspace(0, drop_path, depth)
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, grads, new_state = O.loss_and_grads(sd, ocfg, x, y, label_smoothing=0.12, mask_fn=mask_fn)
        sd.update(new_state)
        O.clip_grads_(grads)
        O.adamw_step_({k: sd[k] for k in keys}, grads, state, lr=6e-4, beta1=0.95, wd=0.06)
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    return sample_b / dt, dt, torch.get_num_threads(), "port"


def _cpu_kind_note(kind):
    return ("the UNMODIFIED reference modules (baseline/_ref) driven like HQAViT_CIFAR100.py:1401-1439" if kind == "reference"
            else "oracle port of the reference step (baseline/_ref not installed)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 256
    steps, warm = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all the host threads it can
    rate, dt, cores, kind = cpu_step_rate(sample, steps, warm, threads=host_threads(), dropout=args.dropout, drop_path=args.drop_path)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": args.batch, "dropout": args.dropout, "drop_path": args.drop_path,
                   "note": "CPU path of the reference step: " + _cpu_kind_note(kind) + "; each step = a bounded sample of the workload.  "
                           "This arm is ONE host process per node whatever --gpus says: a ratio against it is only meaningful at N = 1"},
        "cpu_baseline": {"value": rate, "unit": "images/sec", "cores": cores, "kind": kind,
                         "sample": f"{sample} images/step (of per-GPU batch {args.batch}), fp32, fwd+bwd+clip+AdamW, dropout {args.dropout} / "
                                   f"drop_path {args.drop_path} masks drawn per step"},
        "e2e": {"value": rate, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def gemm_roofline(device, batch):
    """Dominant kernel (tc_gemm_nt_kernel, ~21 % of the step over ~280 launches) at its largest block shape, the SWA qkv
    projection [B*16, 192] x [576, 192]^T.  The step is a CUDA graph, so per-kernel events cannot be placed inside it; the
    kernel is timed live on its own launching stream with CUDA events in two ways:
      train  : 48 launches back to back between ONE pair of events, rotating over 6 operand sets (6 x 116 MB >> 126 MB L2, so no
               launch finds its operands in cache) -> average launch duration in the middle of a stream of kernels, which is how it
               runs inside the step (roofline.achieved uses this one);
      single : one launch between a pair of events, L2 flushed before -- includes ~7 us of event / launch latency that the kernel
               never pays inside the graph (in-kernel clock64 trace, tools/gemm_trace.py: 25 us from first CTA start to last CTA exit).
    Returns (t_train, t_single, algorithmic bytes, flops, shape)."""
    import math
    from qavit_b200 import _lib as L
    M, N, K = batch * 16, 576, 192
    W = torch.randn(N, K, device=device) / math.sqrt(K)
    Wb = W.bfloat16()
    bias = torch.zeros(N, device=device)
    sets = [(torch.randn(M, K, device=device).bfloat16(), torch.empty(M, N, device=device, dtype=torch.bfloat16)) for _ in range(6)]
    flush = torch.empty(1 << 30, dtype=torch.uint8, device=device)   # > L2; also keeps the GPU busy while the launch is enqueued
    s = torch.cuda.current_stream().cuda_stream

    def launch(A, C):
        L.check(L.lib.qavit_test_gemm_nt(1, A.data_ptr(), K, M, N, K, W.data_ptr(), Wb.data_ptr(), bias.data_ptr(), C.data_ptr(), 0, s))

    singles = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(*sets[0])
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            singles.append(e0.elapsed_time(e1) * 1e-3)
    trains, R = [], 48
    for i in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(R):
            launch(*sets[r % len(sets)])
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            trains.append(e0.elapsed_time(e1) * 1e-3 / R)
    bytes_alg = M * K * 2 + N * K * 2 + M * N * 2
    flops = 2.0 * M * N * K
    return sum(trains) / len(trains), sum(singles) / len(singles), bytes_alg, flops, (M, N, K)


def measured_traffic(shape):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch at `shape` from the committed ncu --set full capture."""
    for name in ("r2_roofline_traffic.json", "r1_roofline_traffic.json"):      # the newest capture of the kernel as it is now
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p)).get("x".join(str(v) for v in shape))
    return None


def run_ours(args):
    import torch.distributed as dist

    import qavit_b200 as Q
    from qavit_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(42)
    Q.functional.manual_seed(42 + rank)     # dropout masks: independent per replica
    model, wl = build_workload(Q, args.workload, args.dropout, args.drop_path)
    infer = args.mode == "infer"
    model = model.to(dev).set_precision("bf16")
    model = model.eval() if infer else model.train()
    if world > 1:       # identical replicas
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    names = [n for n, _ in model.named_parameters()]
    bank = [model.global_bank.global_k, model.global_bank.global_v]
    # (the parameters the reference never trains -- bank write_*, branch .norm -- are masked off by FusedAdamW's default)
    opt = Q.FusedAdamW(model.named_parameters(), lr=6e-4, betas=(0.95, 0.999), weight_decay=0.06, max_grad_norm=0.5,
                       tail_elems=sum(p.numel() for p in bank))
    # graph mode: forward + backward and clip + AdamW are two CUDA graphs with ONE eager all-reduce of the flat gradient
    # buffer between them; eager mode (--no-graph): bucketed all-reduces overlapped with backward from autograd hooks
    reducer = Q.GradAllReducer(opt, n_buckets=4, bank_params=bank, overlap=not args.graph) if world > 1 else None

    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, 3, wl["img"], wl["img"], generator=g).pin_memory()
    y_host = torch.randint(0, wl["classes"], (B,), generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step(x, y):
        if infer:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return model(x).float().logsumexp(1).mean()      # a scalar that depends on every logit
        opt.zero_grad()
        if reducer and not args.graph:
            reducer.reset()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(x)
        loss = Q.cross_entropy(logits, y, label_smoothing=0.12)
        loss.backward()
        if reducer:
            reducer.finish() if not args.graph else reducer.reduce_flat()
        opt.clip()
        opt.step()
        return loss

    # our kernels per step: counted by the library on one eager step (graph replays re-issue exactly these launches)
    step(x_dev, y_dev)
    torch.cuda.synchronize()
    if reducer is not None:
        reducer.add_producer_stream(getattr(model, "_lat_stream", None))   # HQAViT's lateral path runs on a side stream
    l0 = L.lib.qavit_launch_count()
    step(x_dev, y_dev)
    torch.cuda.synchronize()
    launches_per_step = L.lib.qavit_launch_count() - l0
    graphed = None
    if args.graph and not infer:
        try:
            graphed = Q.GraphedTrainStep(model, opt, x_dev, y_dev, label_smoothing=0.12, autocast_bf16=True, warmup=3,
                                         capture_error_mode="thread_local" if world > 1 else "global",
                                         between=reducer.reduce_flat if reducer else None)
        except Exception as e:      # e.g. a process-group build that cannot be captured: run the step eagerly
            print(f"[bench] rank {rank}: CUDA-graph capture of the step failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            graphed = None
            torch.cuda.synchronize()

    if args.graph and infer:
        class _GraphedInfer:
            """Eval-mode forward captured once and replayed (the eager forward is launch-bound from Python)."""

            def __init__(self):
                self.x = x_dev.clone()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        step(self.x, None)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                self.g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g):
                    self.out = step(self.x, None)

            def __call__(self, x=None, y=None):
                if x is not None:
                    self.x.copy_(x, non_blocking=True)
                self.g.replay()
                return self.out

        graphed = _GraphedInfer()

    def timed(nsteps, from_host):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        pipelined = from_host and graphed is not None and hasattr(graphed, "prefetch")
        if pipelined:
            graphed.prefetch(x_host, y_host)          # step 0's batch: its copy is inside the timed region like every other step's
        for i in range(nsteps):
            if pipelined:
                # the public training-loop form: launch step i on its prefetched batch, start the H2D copy of batch i + 1 under it,
                # then read the loss back (one H2D of the whole batch and one D2H per step, all inside the timed region)
                loss_t = graphed()
                if i + 1 < nsteps:
                    graphed.prefetch(x_host, y_host)
                last = loss_t.item()
            elif graphed is not None:
                last = graphed(x_host, y_host).item() if from_host else graphed()
            elif from_host:
                xb = x_host.to(dev, non_blocking=True)
                yb = y_host.to(dev, non_blocking=True)
                last = step(xb, yb).item()          # D2H read of the step's result
            else:
                last = step(x_dev, y_dev)
                if world > 1:
                    last = last.item()              # eager DP: keep the caching allocator's cross-stream reuse bounded
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, last

    for _ in range(max(3, args.warmup)):
        graphed() if graphed is not None else step(x_dev, y_dev)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, loss = timed(args.steps, from_host=False)
    launches = launches_per_step
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, loss_e2e = timed(args.steps, from_host=True)
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        hbm, tf_burst, tf_sus, src = peaks()
        t_g, t_single, bytes_g, flops_g, shape_g = gemm_roofline(dev, B)
        roof = {"kernel": f"tc_gemm_nt_kernel (SWA qkv projection, M={shape_g[0]} N={shape_g[1]} K={shape_g[2]})",
                "timing": "CUDA events around 48 back-to-back launches on the launching stream, 6 rotating operand sets (700 MB >> L2): "
                          "average launch duration; single_launch_us = one launch per event pair, L2 flushed (adds ~7 us of event / "
                          "launch latency)",
                "bound": "hbm", "achieved": bytes_g / t_g / 1e9, "peak": hbm, "unit": "GB/s", "frac": bytes_g / t_g / 1e9 / hbm,
                "launch_us": t_g * 1e6, "single_launch_us": t_single * 1e6, "frac_single_launch": bytes_g / t_single / 1e9 / hbm,
                "traffic": measured_traffic(shape_g), "algorithmic_bytes": bytes_g, "peak_source": src + " (burst: kernel timed alone)",
                "tensor_tflops": flops_g / t_g / 1e12,
                "tensor_frac_of_burst": flops_g / t_g / 1e12 / tf_burst}
        # whole-step tensor utilisation against the sustained peak, F_min_train = 3 x (185.4 + 1.2 + 199.7) MFLOP / image
        f_train = (1 if infer else 3) * wl["fwd_mflop"] * 1e6
        roof["step_tflops_fmin"] = value / world * f_train / 1e12
        roof["step_frac_of_sustained_peak"] = roof["step_tflops_fmin"] / tf_sus
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.workload == "hqavit_c100" and not infer:
            rate, dt, cores, kind = cpu_step_rate(256, 8, 1, threads=host_threads(), dropout=args.dropout, drop_path=args.drop_path)
            cpu = {"value": rate, "unit": "images/sec", "cores": cores, "kind": kind,
                   "sample": f"256 images/step x 8 steps ({8 * dt:.1f} s of CPU work), fp32, {_cpu_kind_note(kind)}, fwd+bwd+clip+AdamW, "
                             f"dropout {args.dropout} / drop_path {args.drop_path} masks drawn per step"}
        eager = None
        if world == 1 and not args.no_gpu_eager_baseline and args.workload == "hqavit_c100" and not infer and live_reference_available():
            # the like-for-like GPU baseline (SURVEY 8d / BASELINE.md 4.5): the live reference's eager CUDA path on this B200
            eager = {"unit": "images/sec", "what": "UNMODIFIED reference (baseline/_ref) eager CUDA step, torch.autocast(bfloat16) + GradScaler, "
                     "fwd+bwd+clips+torch.optim.AdamW (HQAViT_CIFAR100.py:1401-1439), same dropout / drop_path, batch resident"}
            for bb in sorted({256, B}):
                try:
                    r, dt_e, info = reference_step_rate(dev, bb, 5, 3, args.dropout, args.drop_path)
                    eager[f"b{bb}"] = {"value": r, "ms_per_step": dt_e * 1e3}
                    eager["attention_path"] = ("flash_attn_func where head_dim % 8 == 0, fp32 SDPA for channel-group attention (hd 4)"
                                               if info["has_flash_attn"] else "F.scaled_dot_product_attention (flash_attn not importable)")
                except Exception as e:          # e.g. out of memory at the large batch: report, keep the line
                    eager[f"b{bb}"] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
                    torch.cuda.empty_cache()
        default = args.workload == "hqavit_c100" and not infer
        metric = METRIC if default else f"{'inference' if infer else 'train'} images/sec ({wl['label']}" + (")" if infer else ", fwd+bwd+clip+AdamW)")
        workload = WORKLOAD if default else f"{wl['label']} {'eval-mode forward' if infer else 'training step'}, bf16, synthetic batch"
        line = {
            "metric": metric, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload, "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "dropout": args.dropout, "drop_path": args.drop_path, "label_smoothing": 0.12, "optimizer": "AdamW+clip(0.1 per-param, 0.5 global)",
                       "l2": "working set (saved activations ~2.6 MB/image incl. the lateral path, x batch) >> 126 MB L2; no explicit flush",
                       "lateral_cnn_path": "native (qavit_lateral_* / qavit_splitfusion_*)",
                       "cuda_graph": graphed is not None},
            "e2e": {"value": e2e, "unit": "images/sec", "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 8),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "input_pipeline": ("GraphedTrainStep.prefetch(): the pinned-host batch of step i + 1 is copied on a copy stream while "
                                       "step i runs; every step still does one whole-batch H2D and one loss D2H inside the timed region")
                    if (graphed is not None and hasattr(graphed, "prefetch")) else "copy, then step, then loss read-back"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager,
            "final_loss": float(loss.item() if hasattr(loss, "item") else loss),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4736, help="images per GPU per step.  The reference's 256 was sized for a 6 GB "
                    "laptop GPU (SURVEY.md 8d lists 256 / 1024 / 4096 / 16384 per GPU); 4736 = 148 SMs x 32 makes every "
                    "row count of the step (B*16 and B*64 token rows) a whole number of 128-row GEMM tiles per SM")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hqavit_c100", choices=sorted(WORKLOADS), help="default: BASELINE.json configs[1]")
    ap.add_argument("--mode", default="train", choices=["train", "infer"], help="infer: eval-mode forward throughput")
    ap.add_argument("--dropout", type=float, default=DEFAULT_DROPOUT, help="config.dropout (reference default 0.1, H:56)")
    ap.add_argument("--drop-path", type=float, default=DEFAULT_DROP_PATH, help="config.drop_path (reference default 0.1, H:57)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true", help="skip timing the live reference's eager CUDA step")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="run the step eagerly instead of as one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
        run_ours(args)


if __name__ == "__main__":
    main()
