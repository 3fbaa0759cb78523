#!/usr/bin/env python
"""Benchmark of the WorldForge guided-sampling hot path (BASELINE.json metric).

metric   denoising-steps/sec for Wan2.1-I2V-14B 480p, 81 frames, IRR + FLF + DSG on (BASELINE.json configs[1]),
         synthetic warped inputs, random-init weights.
step     one outer denoising step of the guided loop (reference pipeline_wan_i2v_clean.py:563-728).  The full
         run is 50 steps of which the first 15 are guided (4 DiT forwards + 2 VAE round trips + FLF + DSG) and 35
         are plain (2 DiT forwards): a K-step measurement keeps that 3:7 mix - of the K timed steps the first
         round(0.3*K) are guided.  The W warm-up steps precede them (all guided), so with the default
         W=6, K=10 the timed guided steps have step index >= 6 and run the FLF optical-flow selection like
         steps 6..14 of the real run do.
value    K / device time of the K steps with every input already resident in HBM.
e2e      the same K steps driven through the public pipeline object from pinned HOST buffers: every step uploads
         its inputs (latents, condition, embeddings, and for guided steps the warped clip and mask) and reads the
         new latents back; those copies are inside the timed region.
roofline the dominant kernel is the self-attention (52% of the forward's FLOPs at 480p): algorithmic FLOPs per
         launch 4*L^2*dim over the mean launch time measured with CUDA events inside the timed region, against
         the measured sustained cuBLAS bf16 peak of MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
T_START = time.time()
# the two comparator legs (PyTorch + flash-attn on the GPU, the CPU port) are reported extras: a run that is already this
# many seconds old when it gets to one of them skips it and says so, instead of running into the caller's time limit
EXTRAS_DEADLINE_S = float(os.environ.get("WF_BENCH_EXTRAS_DEADLINE_S", "900"))
# stdout carries exactly one JSON line.  Libraries write there too (the box exports NCCL_DEBUG=VERSION and NCCL prints its
# "NCCL version ..." banner to stdout), so file descriptor 1 is pointed at stderr for the whole run and the result line is
# written to the saved descriptor at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "denoising_steps_per_sec_wan2.1_i2v_14b_480p_81f_irr_flf_dsg"     # the default configuration (BASELINE.json configs[1])


def metric_name(args):
    res = {(480, 832): "480p", (720, 1280): "720p"}.get((args.height, args.width), f"{args.height}x{args.width}")
    return f"denoising_steps_per_sec_wan2.1_i2v_14b_{res}_{args.frames}f_irr_flf_dsg"
UNIT = "steps/s"


def guided_of(K: int) -> int:
    """How many of K timed steps are guided: the 15:35 mix of the 50-step run."""
    return round(0.3 * K)


def wan_config(args, world: int) -> dict:
    """The `config` object of the line - the workload, identical for both arms (`--impl reference` measures the same
    configuration on the host cores and says how in `cpu_baseline.sample`)."""
    f, h, w = (args.frames - 1) // 4 + 1, args.height // 8, args.width // 8
    K, kg = args.steps, guided_of(args.steps)
    cfg_layout = world > 1 and world % 2 == 0 and os.environ.get("WF_LAYOUT", "cfg") != "ulysses"
    sp_world = world // 2 if cfg_layout else world
    par = "single GPU" if world == 1 else ("cfg2 x " if cfg_layout else "") + \
        f"ulysses{sp_world} (DiT tokens, {os.environ.get('WF_ULYSSES', 'peer')} exchange) + vae-rows{world} + flf-channels{world}"
    return {"workload": f"Wan2.1-I2V-14B {args.height}x{args.width} {args.frames}f guided sampling (IRR+FLF+DSG), "
                        f"{kg} guided + {K - kg} plain timed steps (the 15:35 mix of the 50-step run)",
            "tokens": f * (h // 2) * (w // 2), "dit_layers": args.layers, "dit_forwards_timed": 4 * kg + 2 * (K - kg),
            "vae": "fp32 storage, tf32 tensor-core convs", "parallelism": par,
            "l2_policy": "inputs larger than L2 (33 GB of weights, 0.67 GB activations streamed per GEMM)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--height", type=int, default=None, help="default 480 (704 for longcat-refine)")
    ap.add_argument("--width", type=int, default=None, help="default 832 (1280 for longcat-refine)")
    ap.add_argument("--frames", type=int, default=None, help="default 81 (wan) / 93 (longcat)")
    ap.add_argument("--model", default="wan", choices=["wan", "longcat", "longcat-refine"],
                    help="wan: Wan2.1-I2V-14B guided sampling (BASELINE configs[1], the headline); longcat: LongCat-Video distilled "
                         "16-step guided i2v (configs[3]; 93 frames unless --frames is given); longcat-refine: the 480p->720p refine "
                         "pass with block-sparse attention (configs[4]; 16 latent frames of 704x1280, context parallel under torchrun)")
    ap.add_argument("--layers", type=int, default=None, help="DiT depth (default: 40 for Wan2.1-14B, 48 for LongCat; smaller only for dry runs)")
    ap.add_argument("--bsa-chunk", default="4x4x8", help="longcat-refine: BSA chunk (t x h x w latent tokens). 4x4x8 = the code default "
                    "(bsa_interface.py:619-622); 4x4x4 with --height 768 is what the reference's cp = 8 resolution bucket needs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the PyTorch + flash-attn comparator leg (N=1 only)")
    a = ap.parse_args()
    a.height = a.height if a.height is not None else (704 if a.model == "longcat-refine" else 480)
    a.width = a.width if a.width is not None else (1280 if a.model == "longcat-refine" else 832)
    a.frames = a.frames if a.frames is not None else (81 if a.model == "wan" else 93)
    a.layers = a.layers if a.layers is not None else (40 if a.model == "wan" else 48)
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d.get("bf16_tflops_sustained", 1400.0), hbm=d.get("hbm_gbs", 6650.0), src="measured")
    return dict(bf16=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (port of the reference's algorithm) on the host cores, bounded sample, extrapolated
# ----------------------------------------------------------------------------------------------------------------

def _median_spread(ts):
    med = statistics.median(ts)
    return med, (max(ts) - min(ts)) / med


def cpu_baseline(args, reps: int = int(os.environ.get("WF_CPU_REPS", "2"))):
    """The oracle (CPU port of the reference's algorithm) on all host cores, on a bounded sample of the benchmark
    workload, extrapolated to steps/s.  Sample (fixed seeds): ONE full-width Wan-14B DiT block at L = 4680 tokens - the
    token count of BASELINE config 1 (480p, 9 frames), i.e. the benchmark's 60 x 104 latent frames with 3 instead of 21
    of them - its self-attention alone at the same L, and one VAE decode + encode of a 9 x 128 x 192 clip; each timed
    ``reps`` times after one warm-up, MEDIAN taken, spread (max-min)/median reported.  Extrapolation: attention with
    (L/4680)^2 = 49, the rest of the block with L/4680 = 7, x 40 blocks; the VAE with pixels x frames.  Baseline only."""
    import torch
    from oracle import wan_dit, wan_vae
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = wan_dit.DitConfig(num_layers=1)
    P = wan_dit.init_params(cfg, 3)
    f, h, w = (args.frames - 1) // 4 + 1, args.height // 8, args.width // 8
    L = f * (h // 2) * (w // 2)
    grid_s = tuple(int(v) for v in os.environ.get("WF_CPU_GRID", "3x30x52").split("x"))   # tests shrink the sample
    Ls = grid_s[0] * grid_s[1] * grid_s[2]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(Ls, cfg.dim, generator=g)
    e0 = torch.randn(1, 6, cfg.dim, generator=g) * 0.1
    ctx = torch.randn(cfg.img_len + cfg.text_len, cfg.dim, generator=g)
    q = torch.randn(Ls, cfg.num_heads, 128, generator=g)

    def timed(fn, n=reps):
        fn()                                       # warm-up (thread pool, allocator)
        ts = []
        for _ in range(max(n, 1)):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return ts

    with torch.no_grad():
        ts_blk = timed(lambda: wan_dit.block_forward(P, cfg, 0, x, e0, grid_s, ctx, amp=True))
        ts_att = timed(lambda: wan_dit.attention(q, q, q, amp=True), 2 * reps + 1)    # x49 in the extrapolation: more samples
    (t_blk, sp_blk), (t_att, sp_att) = _median_spread(ts_blk), _median_spread(ts_att)
    t_lin = max(t_blk - t_att, 1e-6)
    t_fwd = 40 * (t_lin * L / Ls + t_att * (L / Ls) ** 2)
    vcfg = wan_vae.VaeConfig()
    PV = wan_vae.init_params(vcfg, 4)
    Fs, Hs, Ws = (9, 128, 192) if "WF_CPU_GRID" not in os.environ else (5, 32, 48)
    with torch.no_grad():
        z = torch.randn(16, (Fs - 1) // 4 + 1, Hs // 8, Ws // 8, generator=g)
        ts_vae = timed(lambda: wan_vae.encode_mode(PV, vcfg, wan_vae.decode(PV, vcfg, z)))
    t_vs, sp_vae = _median_spread(ts_vae)
    t_vae = t_vs * (args.frames * args.height * args.width) / (Fs * Hs * Ws)
    # the timed steps of this configuration: kg guided steps (4 forwards + 2 VAE round trips) and K - kg plain steps (2 forwards)
    K, kg = max(args.steps, 1), guided_of(args.steps)
    t_step = (kg * (4 * t_fwd + 2 * t_vae) + (K - kg) * 2 * t_fwd) / K
    return {"value": 1.0 / t_step, "unit": UNIT, "cores": cores, "kind": "port",
            "reps": reps, "spread": {"block": round(sp_blk, 4), "attention": round(sp_att, 4), "vae": round(sp_vae, 4)},
            "block_s": t_blk, "attention_s": t_att, "vae_round_trip_s": t_vs,
            "sample": f"oracle (CPU port of the reference) on {cores} threads, fixed seeds, median of {reps} after a warm-up: one "
                      f"Wan-14B-width DiT block at L={Ls} tokens ({t_blk:.2f}s, spread {sp_blk:.1%}; its self-attention "
                      f"{t_att:.2f}s, spread {sp_att:.1%}) and one VAE decode+encode of {Fs}x{Hs}x{Ws} ({t_vs:.2f}s, spread "
                      f"{sp_vae:.1%}), extrapolated to L={L} tokens x 40 blocks / {args.frames}x{args.height}x{args.width} and "
                      f"the {kg}:{K - kg} guided:plain mix of the {K} timed steps"}


def cpu_baseline_longcat(args, refine: bool, reps: int = int(os.environ.get("WF_CPU_REPS", "2"))):
    """The LongCat lines' CPU baseline, same recipe as ``cpu_baseline``: ONE full-width LongCat-13.6B block of the oracle
    (oracle/longcat_dit.py, bf16 autocast arithmetic) at a bounded token count, its self-attention alone, and (guided i2v
    only) one VAE decode + encode of a 9 x 128 x 192 clip; medians after a warm-up; attention scaled with L^2, the rest
    with L, x 48 blocks.  i2v (distilled, configs[3]): a guided step is 2 forwards + 1 VAE round trip, a plain step 1
    forward, mixed like the timed steps.  Refine pass (configs[4]): 1 forward per step with block-sparse attention - the
    attention term is scaled by the selected fraction (1 - sparsity) of the dense cost; chunk gating is not timed."""
    import torch
    from oracle import longcat_dit, wan_vae
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = longcat_dit.LongCatConfig(depth=1)
    P = longcat_dit.init_params(cfg, 3)
    T = 16 if refine else (args.frames - 1) // 4 + 1
    h, w = args.height // 8, args.width // 8
    L = T * (h // 2) * (w // 2)
    grid_s = tuple(int(v) for v in os.environ.get("WF_CPU_GRID", "3x30x52").split("x"))
    Ls = grid_s[0] * grid_s[1] * grid_s[2]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(Ls, cfg.hidden_size, generator=g).to(torch.bfloat16)
    y = torch.randn(64, cfg.hidden_size, generator=g).to(torch.bfloat16)
    t = torch.randn(grid_s[0], cfg.adaln_tembed_dim, generator=g)
    q = torch.randn(Ls, cfg.num_heads, cfg.head_dim, generator=g).to(torch.bfloat16)

    def timed(fn, n=reps):
        fn()
        ts = []
        for _ in range(max(n, 1)):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return ts

    with torch.no_grad():
        ts_blk = timed(lambda: longcat_dit.block_forward(P, cfg, 0, x, y, t, grid_s, 1, True))
        ts_att = timed(lambda: longcat_dit.attention(q, q, q, True), 2 * reps + 1)
    (t_blk, sp_blk), (t_att, sp_att) = _median_spread(ts_blk), _median_spread(ts_att)
    t_lin = max(t_blk - t_att, 1e-6)
    keep = (1.0 - 0.9375) if refine else 1.0
    t_fwd = 48 * (t_lin * L / Ls + keep * t_att * (L / Ls) ** 2)
    K = max(args.steps, 1)
    if refine:
        t_step, t_vs, sp_vae, mix = t_fwd, None, None, "1 forward per step (block-sparse attention: 1/16 of the dense cost)"
    else:
        vcfg = wan_vae.VaeConfig()
        PV = wan_vae.init_params(vcfg, 4)
        Fs, Hs, Ws = (9, 128, 192) if "WF_CPU_GRID" not in os.environ else (5, 32, 48)
        with torch.no_grad():
            z = torch.randn(16, (Fs - 1) // 4 + 1, Hs // 8, Ws // 8, generator=g)
            ts_vae = timed(lambda: wan_vae.encode_mode(PV, vcfg, wan_vae.decode(PV, vcfg, z)))
        t_vs, sp_vae = _median_spread(ts_vae)
        t_vae = t_vs * (args.frames * args.height * args.width) / (Fs * Hs * Ws)
        kg = round(0.625 * args.steps)
        t_step = (kg * (2 * t_fwd + t_vae) + (K - kg) * t_fwd) / K
        mix = f"{kg}:{K - kg} guided:plain mix of the {K} timed steps (guided: 2 forwards + 1 VAE round trip of {Fs}x{Hs}x{Ws} scaled by pixels x frames)"
    return {"value": 1.0 / t_step, "unit": UNIT, "cores": cores, "kind": "port", "reps": reps,
            "spread": {"block": round(sp_blk, 4), "attention": round(sp_att, 4), "vae": None if sp_vae is None else round(sp_vae, 4)},
            "block_s": t_blk, "attention_s": t_att, "vae_round_trip_s": t_vs,
            "sample": f"oracle (CPU port of the reference) on {cores} threads, fixed seeds, median of {reps} after a warm-up: one "
                      f"LongCat-13.6B-width block at L={Ls} tokens ({t_blk:.2f}s, spread {sp_blk:.1%}; its self-attention {t_att:.2f}s, "
                      f"spread {sp_att:.1%}), extrapolated to L={L} tokens x 48 blocks; {mix}"}


def guarded_baseline(fn, args, *a):
    """A reported baseline never fails (or delays past the deadline) the bench line it rides on."""
    if args.no_cpu_baseline:
        return None
    if time.time() - T_START > EXTRAS_DEADLINE_S:
        return {"skipped": f"the run was older than {EXTRAS_DEADLINE_S:.0f} s (WF_BENCH_EXTRAS_DEADLINE_S) when this leg was due"}
    try:
        return fn(args, *a)
    except Exception as ex:  # noqa: BLE001
        return {"error": f"{type(ex).__name__}: {ex}"}


def run_reference(args):
    """--impl reference: the reference's CPU path = the oracle port on all host cores (the Python reference itself cannot
    travel to the GPU box and has no compiled component).  One bounded sample, every piece timed >= 5 times, medians."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.model != "wan":      # the CPU arm times the headline configuration only (BASELINE.json configs[1] / [2] shapes)
        emit({"impl": "reference", "unavailable": f"the CPU arm covers --model wan (the metric's configuration), not {args.model}"})
        return
    base = cpu_baseline(args, reps=int(os.environ.get("WF_CPU_REPS_REF", max(5, min(args.steps, 7)))))
    v = base["value"]
    emit({
        "impl": "reference", "metric": metric_name(args), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": wan_config(args, int(os.environ.get("WORLD_SIZE", "1"))),     # the same configuration object as our arm's
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def gpu_reference(args, tr, dev, devt, n_guided: int = 2, n_plain: int = 2):
    """The comparator BASELINE.json's metric names - "the reference's PyTorch+flash-attn path on the same box" (BASELINE.md
    §3, kind port-gpu): oracle/gpu_path.py = the vendored WanModel op sequence on cuBLAS bf16 GEMMs + eager fp32 torch ops +
    flash_attn_varlen_func (flash-attn 2.8.3), the chunked / cached WanVAE_ schedule on cuDNN (tf32), the reference loop
    and scheduler, OpenCV FLF scoring on the host - on the SAME weights, inputs, GPU.  One guided warm-up step, then
    ``n_guided`` guided + ``n_plain`` plain steps timed with CUDA events; steps/s for the guided:plain mix of the timed steps of
    our arm (``guided_of``: 15:35 of the 50-step run)."""
    import torch
    from oracle import gpu_path, pipeline as opipe, unipc, wan_vae
    torch.cuda.empty_cache()
    ref_tr = gpu_path.RefGpuTransformer(tr)
    vcfg = wan_vae.VaeConfig()
    ref_vae = gpu_path.RefGpuVAE(wan_vae.init_params(vcfg, 4321), vcfg, dev)
    warm = 1
    guide, total = warm + n_guided, warm + n_guided + n_plain
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(total + 1)]
    ev[0].record()
    opipe.denoise_loop(ref_tr, ref_vae, unipc.OracleUniPC(flow_shift=3.0), devt["latents"].clone(), devt["condition"],
                       devt["prompt_embeds"], devt["negative_prompt_embeds"], devt["image_embeds"], 50, 4.0,
                       video_ref=devt["video_ref"], mask=devt["mask"], guided=True, resample_steps=2, guide_steps=guide,
                       omega=4.0, omega_resample=4.0, resample_round=guide, use_pca_channel_selection=True, static=True,
                       generator=torch.Generator().manual_seed(42), on_step=lambda i, lat: ev[i + 1].record(), max_steps=total)
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(total)]
    g_ms, p_ms = statistics.mean(ms[warm:guide]), statistics.mean(ms[guide:])
    K, kg = max(args.steps, 1), guided_of(args.steps)
    t_step = (kg * g_ms + (K - kg) * p_ms) / K           # the guided:plain mix of our arm's timed steps (15:35 for K = 20)
    del ref_tr, ref_vae
    torch.cuda.empty_cache()
    import flash_attn
    return {"value": 1000.0 / t_step, "unit": UNIT, "kind": "port-gpu", "guided_step_ms": g_ms, "plain_step_ms": p_ms,
            "dit_forward_ms": p_ms / 2, "step_ms": ms,
            "sample": f"oracle/gpu_path.py on the same GPU, weights and inputs: {warm} guided warm-up step, then {n_guided} guided + "
                      f"{n_plain} plain steps (CUDA events), weighted {kg}:{K - kg} like the {K} timed steps of this run; cuBLAS bf16 Linears, eager fp32 norms / complex128 "
                      f"RoPE, flash-attn {flash_attn.__version__} varlen, chunked cuDNN tf32 VAE, OpenCV FLF scoring on the host "
                      "(the timed guided steps have step index < 6, where the reference's selector returns without scoring)"}


# ----------------------------------------------------------------------------------------------------------------
# ours
# ----------------------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from worldforge_b200 import lib, pipeline as wpipe, scheduler as wsched, synth, transformer as wtr, vae as wvae

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()

    cfg = wtr.WanDitConfig(num_layers=args.layers)
    tr = wtr.WfWanTransformer.random_init(cfg, dev, seed=1234)
    cfgp, sp_world = None, world
    if world > 1:
        from worldforge_b200 import ulysses
        # WF_ULYSSES=nccl: the all-to-all form (A/B measurements); default: q|k|v and attention output stored straight into the
        # peers' memory over NVLink by the producing kernels.  WF_LAYOUT=ulysses: every rank in one Ulysses group (A/B);
        # default on an even rank count: the two CFG forwards in the two halves of the ranks, Ulysses inside a half
        peer = os.environ.get("WF_ULYSSES", "peer") != "nccl"
        if world % 2 == 0 and os.environ.get("WF_LAYOUT", "cfg") != "ulysses":
            sp_group, cfgp = ulysses.cfg_layout(world, rank)
            sp_world = world // 2
            if sp_group is not None:
                ulysses.enable(tr, sp_group, peer=peer)
        else:
            ulysses.enable(tr, dist.group.WORLD, peer=peer)
    vae = wvae.WfWanVAE.random_init(dev, seed=4321)
    if world > 1:
        vae.enable_row_sharding(dist.group.WORLD)      # encode / decode split by image rows; FLF scoring by channels
    inp = synth.make_inputs(args.frames, args.height, args.width, seed=42)
    f, h, w = (args.frames - 1) // 4 + 1, args.height // 8, args.width // 8
    L = f * (h // 2) * (w // 2)

    K, W = args.steps, args.warmup
    k_guided = guided_of(K)
    total = W + K
    guide = W + k_guided
    knobs = dict(guidance_scale=4.0, guided=True, resample_steps=2, guide_steps=guide, omega=4.0, omega_resample=4.0,
                 resample_round=guide, use_pca_channel_selection=True, static=True)

    host = {k: getattr(inp, k).contiguous().pin_memory() for k in
            ("latents", "condition", "prompt_embeds", "negative_prompt_embeds", "image_embeds", "video_ref", "mask")}
    devt = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    lat_host_out = torch.empty(inp.latents.shape, dtype=torch.bfloat16).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(from_host: bool):
        sched = wsched.WfUniPCScheduler(flow_shift=3.0)
        sampler = wpipe.GuidedSampler(tr, vae, sched, generator=torch.Generator().manual_seed(42), cfg_parallel=cfgp, **knobs)
        sampler.begin(total, dev)
        latents = devt["latents"].clone()
        lat_cpu = host["latents"]
        h2d = d2h = 0
        t_ev = None
        for i in range(total):
            if i == W:
                barrier()
                lib.launches = 0
                lib.timed_attention = [] if not from_host else None
                lib.trace = {} if (os.environ.get("WF_TRACE") and not from_host) else None
                clocks.start() if not from_host else None
                t_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                t_ev[0].record()
            if from_host:
                guided_step = i < guide
                ins = {k: host[k].to(dev, non_blocking=True) for k in ("condition", "prompt_embeds", "negative_prompt_embeds", "image_embeds")}
                lat_in = lat_cpu.to(dev, non_blocking=True)
                vr = host["video_ref"].to(dev, non_blocking=True) if guided_step else None
                mk = host["mask"].to(dev, non_blocking=True) if guided_step else None
                if i >= W:
                    h2d += sum(v.numel() * v.element_size() for v in ins.values()) + lat_in.numel() * lat_in.element_size()
                    h2d += (vr.numel() * 4 + mk.numel() * 4) if guided_step else 0
                latents = sampler.step(i, lat_in, ins["condition"], ins["prompt_embeds"],
                                       ins["negative_prompt_embeds"], ins["image_embeds"], vr, mk)
                lat_host_out.copy_(latents.to(torch.bfloat16), non_blocking=True)
                torch.cuda.current_stream().synchronize()
                lat_cpu = lat_host_out
                if i >= W:
                    d2h += lat_host_out.numel() * 2
            else:
                latents = sampler.step(i, latents, devt["condition"], devt["prompt_embeds"], devt["negative_prompt_embeds"],
                                       devt["image_embeds"], devt["video_ref"], devt["mask"])
        t_ev[1].record()
        barrier()
        ms = t_ev[0].elapsed_time(t_ev[1])
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, h2d // max(K, 1), d2h // max(K, 1), sampler.forwards

    clocks = ClockSampler(local)
    ms, _, _, fwd = run(from_host=False)
    clk = clocks.stop()
    if lib.trace is not None:                 # WF_TRACE=1: where the timed steps went (development; rank 0, stderr)
        summ = lib.trace_summary()
        lib.trace = None
        if rank == 0:
            for k, (n, t_ms) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
                print(f"[trace] {k:28s} {n:4d} calls {t_ms:10.1f} ms  ({t_ms / ms * 100:5.1f} % of the timed region)", file=sys.stderr, flush=True)
    launches = lib.launches
    attn = list(lib.timed_attention or [])
    lib.timed_attention = None
    attn_ms = [a.elapsed_time(b) for a, b, _, _ in attn]
    e2e = None
    if not args.no_e2e:
        ms_e, h2d, d2h, _ = run(from_host=True)
        e2e = {"value": K / (ms_e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
    if rank != 0:
        return
    pk = peaks()
    roof = None
    if attn_ms:
        mean_ms = statistics.mean(attn_ms)
        flops = 4.0 * L * L * cfg.dim / sp_world    # Ulysses: each rank runs heads/P of the full-sequence attention
        ach = flops / (mean_ms / 1000.0) / 1e12
        traffic = None
        pj = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(pj) and (args.height, args.width, args.frames, world) == (480, 832, 81, 1):   # the captured shape
            traffic = json.load(open(pj)).get("attention_dram_bytes_per_launch")
        roof = {"kernel": "attention_tcgen05 (self-attention)", "bound": "tensor", "achieved": ach, "peak": pk["bf16"],
                "unit": "TFLOP/s", "frac": ach / pk["bf16"], "traffic": traffic, "peak_source": pk["src"] + " sustained cuBLAS bf16",
                "launches_timed": len(attn_ms), "mean_launch_ms": mean_ms,
                "share_of_step": sum(attn_ms) / ms}
    value = K / (ms / 1000.0)
    line = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": dict(wan_config(args, world), dit_forwards_timed=fwd - 4 * W),
        "clocks": clk, "gpu_launches": launches, "e2e": e2e, "roofline": roof,
        "dit_forwards_per_sec": (fwd - 4 * W) / (ms / 1000.0),
    }
    late = lambda: time.time() - T_START > EXTRAS_DEADLINE_S
    skipped = {"skipped": f"the run was older than {EXTRAS_DEADLINE_S:.0f} s (WF_BENCH_EXTRAS_DEADLINE_S) when this leg was due"}
    if world == 1 and not args.no_gpu_reference and late():
        line["gpu_reference"] = dict(skipped)
    elif world == 1 and not args.no_gpu_reference:
        try:
            ref = gpu_reference(args, tr, dev, devt)
            ref["ours_over_reference"] = {"device_resident": value / ref["value"], "e2e": (e2e["value"] / ref["value"]) if e2e else None}
            line["gpu_reference"] = ref
        except Exception as ex:  # comparator only - never fail the bench for it
            line["gpu_reference"] = {"error": f"{type(ex).__name__}: {ex}"}
    if world == 1 and not args.no_cpu_baseline and late():
        line["cpu_baseline"] = dict(skipped)
    elif world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args)
        except Exception as ex:  # baseline only - never fail the bench for it
            line["cpu_baseline"] = {"error": str(ex)}
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
# LongCat-Video (BASELINE configs[3]): distilled 16-step guided i2v, 480p, one GPU
# ----------------------------------------------------------------------------------------------------------------

def run_longcat(args):
    """K outer steps of generate_i2v's loop (pipeline_longcat_video.py:828-994) on the engine: LongCat-Video 13.6 B DiT
    (48 blocks, 4096 wide, 32 heads; no CFG in distilled mode -> one forward per IRR round), Euler scheduler, IRR
    (resample_steps 2), FLF (VAE decode -> blend -> encode, per-channel Farneback selection on the host) and DSG.  The
    16-step distilled run guides its first 10 steps (run_test_case.sh: 8-11): a K-step measurement keeps that 5:3 mix.
    e2e: after every step the latents are read back to pinned host memory and every input of the next step (latents, text
    embeddings, and for guided steps the warped clip and mask) is uploaded again, inside the timed region."""
    import torch
    from worldforge_b200 import lib, longcat, longcat_pipeline as wlp, synth, vae as wvae
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != 1:
        if int(os.environ.get("RANK", "0")) == 0:
            emit({"metric": "denoising_steps_per_sec_longcat_video", "unavailable": "the guided LongCat i2v loop is benchmarked on one GPU (run with --gpus 1); its context parallel is measured with --model longcat-refine"})
        return
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib.load()
    cfg = longcat.LongCatConfig(depth=args.layers)
    dit = longcat.WfLongCatTransformer.random_init(cfg, dev, seed=1234)
    vae = wvae.WfWanVAE.random_init(dev, seed=4321)
    inp = synth.make_inputs(args.frames, args.height, args.width, seed=42)
    T, h, w = (args.frames - 1) // 4 + 1, args.height // 8, args.width // 8
    N = T * (h // 2) * (w // 2)
    K, W = args.steps, args.warmup
    k_guided = round(0.625 * K)
    total, guide = W + K, W + k_guided
    assert total <= 50, "the distilled schedule has 50 anchor steps"
    pe = inp.prompt_embeds.unsqueeze(0)                                   # [1, 1, 512, 4096] bf16; distilled mode: no CFG batch
    pm = torch.zeros(1, pe.shape[2], dtype=torch.int64); pm[:, :64] = 1
    host = {k: v.contiguous().pin_memory() for k, v in dict(latents=inp.latents, pe=pe, pm=pm, video_ref=inp.video_ref, mask=inp.mask).items()}
    devt = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    lat_host = torch.empty_like(host["latents"]).pin_memory()
    knobs = dict(guidance_scale=1.0, do_cfg=False, use_distill=True, guided=True, resample_steps=2, guide_steps=guide,
                 resample_round=guide, omega=4.0, omega_resample=4.0, use_pca_channel_selection=True, max_replace_threshold=3)
    clocks = ClockSampler(0)
    state = {}

    def run(from_host: bool):
        sched = wlp.WfFlowMatchEulerScheduler(shift=1.0)
        ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
        cnt = dict(h2d=0, d2h=0)
        latents = devt["latents"].clone()

        def on_step(i, lat):
            if i == W - 1:                     # the last warm-up step is done: the timed region starts here
                torch.cuda.synchronize()
                lib.launches = 0
                if not from_host:
                    lib.timed_attention = []
                    clocks.start()
                ev[0].record()
            if from_host:                      # result to the host, then the next step's inputs from the host
                lat_host.copy_(lat, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                if i >= W:
                    cnt["d2h"] += lat_host.numel() * 4
                if i < total - 1:
                    nxt_guided = (i + 1) < guide
                    lat.copy_(lat_host, non_blocking=True)
                    devt["pe"].copy_(host["pe"], non_blocking=True); devt["pm"].copy_(host["pm"], non_blocking=True)
                    if nxt_guided:
                        devt["video_ref"].copy_(host["video_ref"], non_blocking=True); devt["mask"].copy_(host["mask"], non_blocking=True)
                    if i >= W - 1:
                        cnt["h2d"] += lat_host.numel() * 4 + host["pe"].numel() * 2 + host["pm"].numel() * 8
                        cnt["h2d"] += (host["video_ref"].numel() * 4 + host["mask"].numel() * 4) if nxt_guided else 0
            if i == total - 1:
                ev[1].record()
        wlp.denoise_loop(dit, vae, sched, latents, devt["pe"], devt["pm"], total, video_ref=devt["video_ref"], mask=devt["mask"],
                         generator=torch.Generator().manual_seed(42), on_step=on_step, **knobs)
        torch.cuda.synchronize()
        state["fuse_calls"] = sched.fuse_calls
        return ev[0].elapsed_time(ev[1]), cnt["h2d"] // max(K, 1), cnt["d2h"] // max(K, 1)

    calls0 = dit.calls
    ms, _, _ = run(False)
    clk = clocks.stop()
    if lib.trace is not None:                 # WF_TRACE=1: where the timed steps went (development; rank 0, stderr)
        summ = lib.trace_summary()
        lib.trace = None
        if rank == 0:
            for k, (n, t_ms) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
                print(f"[trace] {k:28s} {n:4d} calls {t_ms:10.1f} ms  ({t_ms / ms * 100:5.1f} % of the timed region)", file=sys.stderr, flush=True)
    rank = 0
    launches = lib.launches
    fwd = dit.calls - calls0 - 2 * W
    attn = list(lib.timed_attention or [])
    lib.timed_attention = None
    e2e = None
    if not args.no_e2e:
        ms_e, h2d, d2h = run(True)
        e2e = {"value": K / (ms_e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
    pk = peaks()
    roof = None
    if attn:
        tot_ms = sum(a.elapsed_time(b) for a, b, _, _ in attn)
        flops = sum(4.0 * lq * lk * cfg.hidden_size for _, _, lq, lk in attn)
        ach = flops / (tot_ms / 1000.0) / 1e12
        roof = {"kernel": "attention_tcgen05 (self-attention: noise tokens x all tokens, condition frame x itself)", "bound": "tensor",
                "achieved": ach, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": ach / pk["bf16"], "traffic": None,
                "peak_source": pk["src"] + " sustained cuBLAS bf16", "launches_timed": len(attn), "share_of_step": tot_ms / ms}
    value = K / (ms / 1000.0)
    emit({
        "metric": f"denoising_steps_per_sec_longcat_video_distill_480p_{args.frames}f_irr_flf_dsg", "value": value, "unit": UNIT,
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"LongCat-Video 13.6B {args.height}x{args.width} {args.frames}f distilled guided i2v (IRR+FLF+DSG, no CFG), "
                               f"{k_guided} guided + {K - k_guided} plain timed steps (the 10:6 mix of the 16-step run)",
                   "tokens": N, "dit_layers": cfg.depth, "dit_forwards_timed": fwd, "vae": "fp32 storage, tf32 tensor-core convs",
                   "parallelism": "single GPU", "l2_policy": "inputs larger than L2 (27 GB of weights streamed per forward)"},
        "clocks": clk, "gpu_launches": launches, "e2e": e2e, "roofline": roof, "dit_forwards_per_sec": fwd / (ms / 1000.0),
        "flops_per_forward": dit.flops_per_forward(N, (h // 2) * (w // 2), 64),
        "cpu_baseline": guarded_baseline(cpu_baseline_longcat, args, False)})


# ----------------------------------------------------------------------------------------------------------------
# LongCat-Video refine pass (BASELINE configs[4]): 480p -> 720p, block-sparse attention, context parallel over the ranks
# ----------------------------------------------------------------------------------------------------------------

def run_longcat_refine(args):
    """K steps of generate_refine's loop (pipeline_longcat_video.py:1467-1498) on the engine: one DiT forward with block-sparse
    self-attention (chunks 4x4x8, sparsity 0.9375) and an Euler step per timestep; no CFG / IRR / FLF / DSG in this pass.
    16 latent frames (4 clean condition + 12 noise, the BSA padding of :1406-1421) of 704x1280 -> 56 320 tokens (SURVEY.md
    §8a row a12).  Under torchrun the DiT runs LongCat's 2-D context parallel (`cp_split_hw`, every rank a block of every
    frame).  e2e: the latents go to pinned host memory and back every step."""
    import torch
    import torch.distributed as dist
    from worldforge_b200 import lib, longcat, longcat_pipeline as wlp
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    cfg = longcat.LongCatConfig(depth=args.layers)
    dit = longcat.WfLongCatTransformer.random_init(cfg, dev, seed=1234)
    ck = [int(v) for v in args.bsa_chunk.split("x")]
    assert len(ck) == 3 and ck[0] * ck[1] * ck[2] in (64, 128), "--bsa-chunk: 64- or 128-token chunks (t x h x w)"
    ct = ck[0] * ck[1] * ck[2]
    dit.bsa_params = dict(sparsity=0.9375, cdf_threshold=None, chunk_3d_shape_q=list(ck), chunk_3d_shape_k=list(ck))
    dit.enable_bsa()
    T, h, w, ncl = 16, args.height // 8, args.width // 8, 4
    split = [1, 1]
    if world > 1:
        split = min(([i, world // i] for i in range(1, int(world ** 0.5) + 1) if world % i == 0), key=lambda f: abs(f[0] - f[1]))
        if (h // 2) % (ck[1] * split[0]) or (w // 2) % (ck[2] * split[1]):
            if rank == 0:
                emit({"metric": "refine_steps_per_sec_longcat_video_720p_bsa", "unavailable": f"{h // 2}x{w // 2} patch grid does not split into "
                      f"{split[0]}x{split[1]} blocks of whole {ck[1]}x{ck[2]} BSA chunks - the reference's own assert (bsa_interface.py:638-639). It "
                      "switches resolution buckets with cp (pipeline_longcat_video.py:1334-1337 -> bukcet_config.py:82-109: 768x1280 at cp = 8, "
                      "whose 24x20 blocks need the checkpoint's 4x4x4 chunks): --height 768 --width 1280 --bsa-chunk 4x4x4"})
            return
        dit.enable_context_parallel(dist.group.WORLD, split)
    N = T * (h // 2) * (w // 2)
    K, W = args.steps, args.warmup
    total = W + K
    g = torch.Generator().manual_seed(42)
    lat0 = torch.randn(1, 16, T, h, w, generator=g)
    pe = torch.randn(1, 1, 512, 4096, generator=g).to(torch.bfloat16)
    pm = torch.zeros(1, 512, dtype=torch.int64); pm[:, :64] = 1
    host = {k: v.contiguous().pin_memory() for k, v in dict(latents=lat0, pe=pe, pm=pm).items()}
    devt = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    lat_host = torch.empty_like(host["latents"]).pin_memory()
    clocks = ClockSampler(local)

    def run(from_host: bool):
        sched = wlp.WfFlowMatchEulerScheduler(shift=1.0)
        ts = wlp.refine_schedule(sched, 50, 0.5, device=dev)
        assert total <= len(ts), f"the refine schedule has {len(ts)} steps"
        ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
        cnt = dict(h2d=0, d2h=0)
        latents = devt["latents"].clone()

        def on_step(i, lat):
            if i == W - 1:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                lib.launches = 0
                if not from_host:
                    clocks.start()
                ev[0].record()
            if from_host:
                lat_host.copy_(lat, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                if i >= W:
                    cnt["d2h"] += lat_host.numel() * 4
                if i < total - 1:
                    lat.copy_(lat_host, non_blocking=True)
                    devt["pe"].copy_(host["pe"], non_blocking=True); devt["pm"].copy_(host["pm"], non_blocking=True)
                    if i >= W - 1:
                        cnt["h2d"] += lat_host.numel() * 4 + host["pe"].numel() * 2 + host["pm"].numel() * 8
            if i == total - 1:
                ev[1].record()
        wlp.refine_loop(dit, sched, latents, devt["pe"], devt["pm"], ncl, ts[:total], on_step=on_step)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1])
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, cnt["h2d"] // max(K, 1), cnt["d2h"] // max(K, 1)

    ms, _, _ = run(False)
    clk = clocks.stop()
    launches = lib.launches
    e2e = None
    if not args.no_e2e:
        ms_e, h2d, d2h = run(True)
        e2e = {"value": K / (ms_e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
    if rank != 0:
        return
    pk = peaks()
    C, Fd, per = cfg.hidden_size, cfg.ffn_dim, (h // 2) * (w // 2)
    Nn, ctx = N - ncl * per, 64
    n_sel = int((1 - 0.9375) * (N // ct))
    # algorithmic FLOPs of one forward: token-side GEMMs, cross-attention, and the SELECTED blocks of the sparse self-attention
    gemm = cfg.depth * (2 * N * (4 * C * C + 3 * C * Fd) + 2 * Nn * 2 * C * C + 2 * ctx * 2 * C * C + 4 * Nn * ctx * C)
    sparse = cfg.depth * 4.0 * C * ct * ct * n_sel * (N // ct)
    step_ms = ms / K
    ach = (gemm + sparse) / world / (step_ms / 1000.0) / 1e12
    emit({
        "metric": "refine_steps_per_sec_longcat_video_720p_16lf_bsa", "value": K / (ms / 1000.0), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"LongCat-Video 13.6B refine pass {args.height}x{args.width}, 16 latent frames (4 condition + 12 noise), block-sparse "
                               f"self-attention ({args.bsa_chunk} chunks, sparsity 0.9375: {n_sel} of {N // ct} key chunks per query chunk), Euler steps from t = 0.5",
                   "tokens": N, "dit_layers": cfg.depth, "parallelism": "single GPU" if world == 1 else f"context parallel cp_split_hw {split[0]}x{split[1]} (NCCL all-to-all)",
                   "l2_policy": "inputs larger than L2 (27 GB of weights streamed per forward)"},
        "clocks": clk, "gpu_launches": launches, "e2e": e2e,
        "roofline": {"kernel": "DiT forward (tcgen05 GEMMs + block-sparse attention), whole step", "bound": "tensor", "achieved": ach, "peak": pk["bf16"],
                     "unit": "TFLOP/s", "frac": ach / pk["bf16"], "traffic": None, "peak_source": pk["src"] + " sustained cuBLAS bf16",
                     "flops_per_forward": gemm + sparse},
        "cpu_baseline": guarded_baseline(cpu_baseline_longcat, args, True) if world == 1 else None})


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.model == "longcat-refine":
        run_longcat_refine(args)
    elif args.model == "longcat":
        run_longcat(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
