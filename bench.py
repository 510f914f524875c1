#!/usr/bin/env python
"""bench.py — samples/sec of the vision-fusion hot path (harness mode M1, BASELINE.md §2 / SURVEY.md §8d).

One step = CLIP ViT-L/14 forward (frozen, no grad) -> PerceiverResampler (6 blocks, 64 latents) ->
8 x GatedCrossAttentionBlock(dim 4096) chained on a synthetic hidden state, loss = mean(out^2), backward
through the gated blocks and the perceiver (dgrad + wgrad, fp32 gradients into one flat buffer), and for
N > 1 the single NCCL all-reduce of that buffer.  Workload = BASELINE.json configs[1]
(OTTER-Image-MPT7B shape, per-GPU batch 8, 1 image/sample, L = 256, bf16), weak scaling over ranks.

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference [...]                      the reference's CPU implementation of the same path

The CPU arm (`--impl reference`, and the `cpu_baseline` leg of the GPU arm) times oracle/restatement.py — the
CPU restatement of the reference's modules pinned against the real reference by tests/golden — because
/root/reference does not exist on the GPU box (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

SEED = 0
CFG = dict(L=256, D=4096, n_gated=8, T=1, F=1, img=224, clip_layers=24, vis_dim=1024, latents=64)

# ---- algorithmic FLOPs (2*M*N*K per GEMM) ---------------------------------------
def flops_per_sample(L=256, D=4096, T=1, Fr=1, n_gated=8):
    clip = 24 * (257 * (4 * 1024 ** 2 * 2 + 2 * 1024 * 4096 * 2) + 4 * 257 ** 2 * 1024) + 256 * 588 * 1024 * 2
    Nx = Fr * 256
    perc = 6 * (2 * (Nx + 64) * 1024 * 1024 + 2 * 64 * 1024 * 512 + 4 * 8 * 64 * (Nx + 64) * 64
                + 2 * 64 * 512 * 1024 + 4 * 64 * 1024 * 4096)
    gated = 2 * L * D * 512 + 2 * (T * 64) * 1024 * 1024 + 4 * 8 * L * (T * 64) * 64 + 2 * L * 512 * D + 4 * L * D * 4 * D
    return clip * T * Fr + 3 * perc * T + 3 * n_gated * gated


def host_cores():
    """CPU cores this process may actually use: min(affinity mask, cgroup quota, os.cpu_count())."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


# ---- clocks sampler (nvidia-smi during the timed region) -----------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- the CUDA arm -------------------------------------------------------------------------------
def build_modules(device):
    from transformers import CLIPVisionConfig
    from otter_b200.modeling_clip import CLIPVisionModel
    from otter_b200.modeling_otter import OtterGatedCrossAttentionBlock, OtterPerceiverResampler
    torch.manual_seed(SEED)       # identical weights on every rank (DDP semantics, no broadcast needed)
    vc = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=CFG["clip_layers"],
                          num_attention_heads=16, image_size=CFG["img"], patch_size=14, hidden_act="quick_gelu")
    clip = CLIPVisionModel(vc).to(device).requires_grad_(False)
    perceiver = OtterPerceiverResampler(dim=CFG["vis_dim"]).to(device)
    gated = torch.nn.ModuleList([OtterGatedCrossAttentionBlock(dim=CFG["D"], dim_visual=CFG["vis_dim"])
                                 for _ in range(CFG["n_gated"])]).to(device)
    with torch.no_grad():         # gates at 0 would make the path an identity
        for g in gated:
            g.attn_gate.fill_(0.5)
            g.ff_gate.fill_(0.5)
    return clip, perceiver, gated


def host_batch(batch, rank):
    g = torch.Generator().manual_seed(SEED + rank)     # rank r uses seed + r (train_utils.py:33-36)
    vision_x = torch.randn(batch, CFG["T"], CFG["F"], 3, CFG["img"], CFG["img"], generator=g).to(torch.bfloat16)
    hidden = torch.randn(batch, CFG["L"], CFG["D"], generator=g).to(torch.bfloat16)
    loc = torch.zeros(batch, CFG["L"], dtype=torch.bool)
    loc[:, 0] = True                                    # one <image> at position 0
    return vision_x.pin_memory(), hidden.pin_memory(), loc.pin_memory()


def dominant_gemm_times(dev, M, D):
    from otter_b200 import functional as F
    g = torch.Generator(device=dev).manual_seed(1)

    def rn(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)

    x, h, dy, dz = rn(M, D), rn(M, 4 * D), rn(M, D), rn(M, 4 * D)
    w1, w2 = rn(4 * D, D, scale=D ** -0.5), rn(D, 4 * D, scale=(4 * D) ** -0.5)
    z, a2 = torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16), torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    gw1 = torch.zeros(4 * D, D, device=dev)
    gw2 = torch.zeros(D, 4 * D, device=dev)
    gate = torch.full((1,), 0.5, device=dev)
    calls = {
        "fwd_up_gelu": lambda: F.linear_fwd(x, w1, act=1, aux_out=z),
        "fwd_down_gate_res": lambda: F.linear_fwd(h, w2, aux_out=a2, scale_ptr=gate, scale_tanh=True, residual=x),
        "dgrad_down_dgelu": lambda: F.linear_dgrad(dy, w2, aux_in=z, scale_ptr=gate, scale_tanh=True),
        "dgrad_up": lambda: F.linear_dgrad(dz, w1),
        "wgrad_down": lambda: F.linear_wgrad(dy, h, out=gw2, accumulate=False, scale_ptr=gate, scale_tanh=True),
        "wgrad_up": lambda: F.linear_wgrad(dz, x, out=gw1, accumulate=False),
    }
    F.linear_fwd(x, w1, act=1, aux_out=z)        # realistic pre-activations for the dGELU epilogue
    per, tot_ms = {}, 0.0
    flops = 2.0 * M * D * 4 * D
    for name, fn in calls.items():
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        per[name] = round(ms * 1e3, 1)
        tot_ms += ms
    return {"tflops": 6 * flops / (tot_ms * 1e-3) / 1e12, "per_class_us": per, "sum_ms": tot_ms, "flops_avg": flops}


def run_cuda(args):
    import torch.distributed as dist
    from otter_b200 import functional as F
    from otter_b200 import params as P
    from otter_b200.dp import FlatGradBuffer
    from otter_b200.blocks import MediaFromClipFn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    batch = args.per_gpu_batch
    clip, perceiver, gated = build_modules(dev)
    trainable = list(perceiver.parameters()) + list(gated.parameters())
    comm_dtype = torch.bfloat16 if args.grad_comm_dtype == "bf16" else None
    flat = FlatGradBuffer(trainable, device=dev, comm_dtype=comm_dtype, nccl_registered=args.nccl_registered and world > 1)
    h_vis, h_hid, h_loc = host_batch(batch, rank)
    d_vis, d_hid, d_loc = h_vis.to(dev), h_hid.to(dev), h_loc.to(dev)
    B, L, D = batch, CFG["L"], CFG["D"]
    loss_host = torch.zeros(1).pin_memory()

    def clip_step(vis):
        """Frozen CLIP tower on this step's images (no grad, independent of the trainable weights)."""
        b_, T_, F_ = vis.shape[:3]
        return clip.last_hidden_bf16(vis.reshape(b_ * T_ * F_, *vis.shape[3:]))      # bf16 [bTF, 257, 1024]

    def train_step(hidden, hid, loc):
        """perceiver + 8 gated blocks forward/backward on precomputed CLIP features."""
        # weights "just updated by the optimizer": re-derive the bf16 compute copies (autocast-equivalent work)
        if os.environ.get("OTB_MULTI_CAST") == "1":       # candidate: one multi-tensor launch instead of 67 casts
            P.refresh(trainable)
        else:
            P.invalidate(trainable)
        flat.begin_step()
        media = MediaFromClipFn.apply(hidden, perceiver.frame_embs, CFG["F"])            # drop CLS (+frame_embs)
        media = perceiver.resample_media(media, B * CFG["T"])                              # [B*T*64, 1024] bf16
        tt = F.text_time(loc, True)
        x = hid.view(B * L, D).detach().requires_grad_(True)
        for g in gated:
            x = g.forward_2d(x, media, tt, B, L, CFG["T"], CFG["latents"])
        loss, dx = F.sqmean_loss(x)
        x.backward(dx)
        flat.finish_step()
        return loss

    def step(vis, hid, loc):
        return train_step(clip_step(vis), hid, loc)

    def log(msg):
        if rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    graphed = None          # (graph of clip_step, graph of train_step) when CUDA graphs are on
    launches_per_step = None
    state = {"hidden": None}

    def run_clip():
        state["hidden"] = graphed[0].replay() if graphed is not None else clip_step(d_vis)

    def run_step():
        """One step = train_step on the CLIP features of this batch, then the ONE gradient all-reduce, overlapped
        with the frozen CLIP forward of the next batch (data prefetch: it does not depend on the weight update)."""
        loss = graphed[1].replay() if graphed is not None else train_step(state["hidden"], d_hid, d_loc)
        work = flat.all_reduce(async_op=True)
        run_clip()
        if work is not None:
            work.wait()
            if dist.get_backend() == "gloo" and not getattr(work, "averaged", False):
                flat.flat.div_(world)
        return loss

    prefetch = None
    if args.e2e_prefetch:
        # opt-in loader pipeline for the e2e leg: a copy stream fills device staging buffers with the next step's
        # inputs while the current step computes; the step then starts with three device-to-device copies.
        cs = torch.cuda.Stream(device=dev)
        prefetch = {"vis": torch.empty_like(d_vis), "hid": torch.empty_like(d_hid), "loc": torch.empty_like(d_loc),
                    "landed": torch.cuda.Event(), "free": torch.cuda.Event()}

        def _issue():
            cs.wait_event(prefetch["free"])
            with torch.cuda.stream(cs):
                prefetch["vis"].copy_(h_vis, non_blocking=True)
                prefetch["hid"].copy_(h_hid, non_blocking=True)
                prefetch["loc"].copy_(h_loc, non_blocking=True)
                prefetch["landed"].record(cs)

        prefetch["issue"] = _issue
        prefetch["free"].record(torch.cuda.current_stream())
        _issue()                                         # prologue: inputs of the first timed step

    def timed(n, e2e):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start.record()
        for _ in range(n):
            if e2e and prefetch is not None:             # same bytes per step, H2D of the NEXT step's inputs overlapped
                main = torch.cuda.current_stream()
                main.wait_event(prefetch["landed"])      # this step's inputs sit in the staging buffers
                d_vis.copy_(prefetch["vis"]); d_hid.copy_(prefetch["hid"]); d_loc.copy_(prefetch["loc"])   # D2D, us
                prefetch["free"].record(main)
                prefetch["issue"]()                      # pinned host -> staging on the copy stream, behind `free`
                loss = run_step()
                loss_host.copy_(loss, non_blocking=True)
                main.synchronize()                                          # the user reads the loss every step
            elif e2e:                                    # pinned host -> static device inputs, step, loss -> host
                d_vis.copy_(h_vis, non_blocking=True)
                d_hid.copy_(h_hid, non_blocking=True)
                d_loc.copy_(h_loc, non_blocking=True)
                loss = run_step()
                loss_host.copy_(loss, non_blocking=True)
                torch.cuda.current_stream().synchronize()                   # the user reads the loss every step
            else:
                run_step()
        end.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = start.elapsed_time(end)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    log(f"modules built (world {world}, per-GPU batch {batch})")
    step(d_vis, d_hid, d_loc)                 # first step also builds the frozen CLIP weight shadows
    n0 = F.launch_count()
    step(d_vis, d_hid, d_loc)
    launches_per_step = F.launch_count() - n0
    if not args.no_graph:
        from otter_b200.graph import GraphedStep
        ga = GraphedStep(clip_step, d_vis)
        gb = GraphedStep(train_step, ga.outputs, d_hid, d_loc)
        graphed = (ga, gb)
        log("step captured as two CUDA graphs (CLIP forward | perceiver + gated fwd/bwd)")
    run_clip()                                  # pipeline prologue: features of the first batch
    for _ in range(max(args.warmup, 3)):
        run_step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(args.steps, e2e=False)
    log(f"timed region done: {ms / args.steps:.3f} ms/step")
    launches = launches_per_step
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(args.steps, e2e=True)

    # ---- roofline of the dominant kernel: the six 4096<->16384 FFN GEMM launch classes of the gated blocks
    # (8 launches each per step, 86 % of the step's algorithmic FLOPs), each timed live with CUDA events over
    # 10 back-to-back launches on fresh N(0,1) operands (operands + outputs > L2) ----
    dom = dominant_gemm_times(dev, B * L, D) if rank == 0 else None
    # eager per-launch event timing of EVERY GEMM of one step (includes host launch gaps; kept as a cross-check)
    prof = []
    F.set_gemm_profiler(prof)
    step(d_vis, d_hid, d_loc)
    flat.all_reduce()
    F.set_gemm_profiler(None)
    torch.cuda.synchronize()
    g_flops = sum(r[0] for r in prof)
    g_ms = sum(r[1].elapsed_time(r[2]) for r in prof)
    if rank == 0 and os.environ.get("OTB_GEMM_BREAKDOWN"):
        agg = {}
        for f, a_, b_, shape in prof:
            t = agg.setdefault(shape, [0, 0.0, 0.0])
            t[0] += 1
            t[1] += a_.elapsed_time(b_)
            t[2] += f
        with open(os.environ["OTB_GEMM_BREAKDOWN"], "w") as fh:
            fh.write("M,N,K,a_mn,b_mn,launches,total_ms,avg_us,TFLOPs\n")
            for shape, (n, ms_, fl_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                fh.write(",".join(map(str, shape)) + f",{n},{ms_:.3f},{ms_ / n * 1e3:.1f},{fl_ / ms_ / 1e9:.1f}\n")
    if world > 1:
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)                # kernel timed alone -> burst figure
    peak_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else \
        "fallback 1590 TFLOP/s (B200_PROFILING.md)"
    step_ms = ms / args.steps
    value = batch * world * args.steps / (ms / 1e3)
    e2e_value = batch * world * args.steps / (ms_e2e / 1e3)
    fl = flops_per_sample(CFG["L"], CFG["D"], CFG["T"], CFG["F"], CFG["n_gated"])
    ach = dom["tflops"]
    out = {
        "metric": "samples/sec perceiver+gated-xattn fwd+bwd", "value": round(value, 2), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(step_ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "OTTER-Image-MPT7B shape (BASELINE configs[1]), M1 hot path: CLIP ViT-L/14 fwd -> "
                               "perceiver(6x64 latents) -> 8 gated x-attn blocks D=4096, fwd+bwd",
                   "per_gpu_batch": batch, "global_batch": batch * world, "L": CFG["L"], "images_per_sample": 1,
                   "parallelism": f"dp{world}", "random_init": True, "gates": 0.5,
                   "launch": "eager" if args.no_graph else "two CUDA graphs per step (frozen CLIP forward | perceiver + gated "
                             "fwd/bwd), replayed" + ("; the single NCCL gradient all-reduce runs between them, overlapped "
                             "with the CLIP forward of the next batch" if world > 1 else ""),
                   "weights": "fp32 master, bf16 compute copies re-cast every step; fp32 grads in one flat buffer",
                   "cache": "per-step working set (2.4 GB bf16 weights + activations) >> 126 MB L2; no explicit flush",
                   "grad_allreduce_bytes": flat.comm_nbytes() if world > 1 else 0,
                   "grad_allreduce_dtype": args.grad_comm_dtype, "nccl_registered_buffer": bool(args.nccl_registered)},
        "e2e": {"value": round(e2e_value, 2), "unit": "samples/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": (h_vis.numel() * 2 + h_hid.numel() * 2 + h_loc.numel()) * 1,
                "d2h_bytes_per_step": 4, "h2d_overlapped_with_previous_step": bool(args.e2e_prefetch)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor",
                     "kernel": "otb::gemm2_bf16_kernel (tcgen05 cta_group::2 / TMA) — the six 4096<->16384 FFN GEMM classes "
                               "(fwd, dgrad, wgrad) of the gated blocks, 48 launches/step",
                     "achieved": round(ach, 1), "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": round(ach / peak_tf, 4) if peak_tf else None,
                     "traffic": 269.4e6, "traffic_note": "dram read+write of one 2048x16384x4096 GELU+aux launch of "
                                                         "gemm2_bf16_kernel, ncu --set full, profiles/r01_ncu_full_v2.md "
                                                         "(algorithmic 285 MB)",
                     "peak_source": peak_src, "frac_of_sustained_peak": round(ach / peak_sus, 4),
                     "per_class_us": dom["per_class_us"], "flops_per_launch_avg": dom["flops_avg"],
                     "share_of_step": round(8 * dom["sum_ms"] / step_ms, 3),
                     "all_gemms_eager_event_ms": round(g_ms, 3), "all_gemm_launches": len(prof),
                     "step_algorithmic_tflops": round(fl * batch / (step_ms * 1e-3) / 1e12, 1),
                     "step_frac_of_sustained_peak": round(fl * batch / (step_ms * 1e-3) / 1e12 / peak_sus, 4)},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_port_sample(steps=1, warmup=0, batch=1)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---- the CPU arm: reference modules' math on the host cores (oracle port) -----------------------
def cpu_port_sample(steps, warmup, batch):
    """Time `steps` M1 steps of the reference math on CPU (fp32, all host threads) at a bounded batch."""
    from oracle import restatement as R
    cores = host_cores()               # affinity / cgroup quota aware (the GPU box: 128 visible, quota 16)
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(SEED)
    D, Dv = CFG["D"], CFG["vis_dim"]

    def lin(o, i):
        return torch.randn(o, i, generator=g) / i ** 0.5

    def ln(d):
        return torch.ones(d), torch.zeros(d)

    clip_p = {}
    pre = "vision_model."
    clip_p[pre + "embeddings.patch_embedding.weight"] = torch.randn(Dv, 3, 14, 14, generator=g) * 0.02
    clip_p[pre + "embeddings.class_embedding"] = torch.randn(Dv, generator=g) * 0.02
    clip_p[pre + "embeddings.position_embedding.weight"] = torch.randn(257, Dv, generator=g) * 0.02
    clip_p[pre + "pre_layrnorm.weight"], clip_p[pre + "pre_layrnorm.bias"] = ln(Dv)
    for i in range(CFG["clip_layers"]):
        lp = f"{pre}encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            clip_p[lp + f"self_attn.{nm}.weight"], clip_p[lp + f"self_attn.{nm}.bias"] = lin(Dv, Dv), torch.zeros(Dv)
        clip_p[lp + "layer_norm1.weight"], clip_p[lp + "layer_norm1.bias"] = ln(Dv)
        clip_p[lp + "layer_norm2.weight"], clip_p[lp + "layer_norm2.bias"] = ln(Dv)
        clip_p[lp + "mlp.fc1.weight"], clip_p[lp + "mlp.fc1.bias"] = lin(4 * Dv, Dv), torch.zeros(4 * Dv)
        clip_p[lp + "mlp.fc2.weight"], clip_p[lp + "mlp.fc2.bias"] = lin(Dv, 4 * Dv), torch.zeros(Dv)
    perc_p = {"latents": torch.randn(64, Dv, generator=g)}
    perc_p["norm.weight"], perc_p["norm.bias"] = ln(Dv)
    for i in range(6):
        lp = f"layers.{i}."
        for nm in ("norm_media", "norm_latents", "feed_forward.0"):
            perc_p[lp + nm + ".weight"], perc_p[lp + nm + ".bias"] = ln(Dv)
        perc_p[lp + "to_q.weight"], perc_p[lp + "to_kv.weight"], perc_p[lp + "to_out.weight"] = lin(512, Dv), lin(1024, Dv), lin(Dv, 512)
        perc_p[lp + "feed_forward.1.weight"], perc_p[lp + "feed_forward.3.weight"] = lin(4 * Dv, Dv), lin(Dv, 4 * Dv)
    gated_ps = []
    for _ in range(CFG["n_gated"]):
        gp = {"attn_gate": torch.tensor([0.5]), "ff_gate": torch.tensor([0.5])}
        gp["attn.norm.weight"], gp["attn.norm.bias"] = ln(D)
        gp["feed_forward.0.weight"], gp["feed_forward.0.bias"] = ln(D)
        gp["attn.to_q.weight"], gp["attn.to_kv.weight"], gp["attn.to_out.weight"] = lin(512, D), lin(1024, Dv), lin(D, 512)
        gp["feed_forward.1.weight"], gp["feed_forward.3.weight"] = lin(4 * D, D), lin(D, 4 * D)
        gated_ps.append(gp)
    train = [t for t in perc_p.values()] + [t for gp in gated_ps for t in gp.values()]
    for t in train:
        t.requires_grad_(True)
    vision_x = torch.randn(batch, CFG["T"], CFG["F"], 3, CFG["img"], CFG["img"], generator=g)
    hidden = torch.randn(batch, CFG["L"], D, generator=g).requires_grad_(True)
    loc = torch.zeros(batch, CFG["L"], dtype=torch.bool)
    loc[:, 0] = True

    def one():
        for t in train:
            t.grad = None
        out, _ = R.m1_forward(vision_x, hidden, loc, clip_p, perc_p, gated_ps)
        loss = out.float().pow(2).mean()
        loss.backward()
        return loss.item()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return {"value": round(batch * steps / dt, 4), "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) of the same M1 workload at batch {batch} (fp32, torch CPU, {cores} threads), "
                      f"{dt / steps:.2f} s/step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_port_sample(steps=args.steps, warmup=min(args.warmup, 1), batch=1)
    dt_ms = 1e3 / res["value"]
    out = {"impl": "reference", "metric": "samples/sec perceiver+gated-xattn fwd+bwd", "value": res["value"],
           "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
           "ms_per_step": round(dt_ms, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "OTTER-Image-MPT7B shape (BASELINE configs[1]), M1 hot path, reference math on host "
                                  "cores (oracle/restatement.py port; /root/reference is not on the GPU box)",
                      "per_step_batch": 1, "L": CFG["L"]},
           "cpu_baseline": res,
           "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="otter_b200", choices=["otter_b200", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--grad-comm-dtype", default="fp32", choices=["fp32", "bf16"],
                    help="wire format of the single gradient all-reduce (gradients stay fp32 on both sides); bf16 "
                         "halves the payload to SURVEY.md §8e's 2.36 GB at the cost of two cast passes per step")
    ap.add_argument("--e2e-prefetch", action="store_true",
                    help="e2e leg: copy the next step's inputs host->device on a copy stream while this step computes")
    ap.add_argument("--nccl-registered", action="store_true",
                    help="allocate the flat gradient buffer from NCCL's allocator and register it (zero-copy / NVLS)")
    ap.add_argument("--seq-len", type=int, default=256, help="text length L (SURVEY.md §8d sweeps 128/256/512/1024)")
    args = ap.parse_args()
    CFG["L"] = args.seq_len
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
