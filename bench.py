#!/usr/bin/env python
"""Fusion-head fwd+bwd throughput on B200 (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config rgbnt201|rgbnt100|msvr310] [--dim 768|512]
                    [--batch B] [--impl reference] [--exchange nvlink|nccl] [--check]

One "step" = Select_Interactive_Module fwd + AlignmentM fwd (GAM + LAM) + backward to the three [B,129,d] token maps and
all head parameters, on one synthetic batch (B per GPU, bf16 tokens, fp32 master parameters).  --config picks the
reference configuration whose shape is run (configs/<NAME>/Signal.yml: grid, TOPK, loss weights); the default is the
one BASELINE.json's metric is quoted on (RGBNT201, B=128, d=768).  N>1 (torchrun, one rank per GPU): the batch is sharded
(weak scaling, B per rank; --strong: global batch 1024 split over the ranks) and the head gradients are averaged over
the ranks inside the backward -- by the library's own NVLink kernel (sig_xchg_allreduce_f32) or by NCCL.

value   : samples/s, tokens already resident in HBM, CUDA-graph replay of the FusionHead step, CUDA events, max over ranks
variants: the same step eager through FusionHead, and eager through the reference's call sequence SIM(...); AlignM(...)
e2e     : the step from pinned HOST token maps (H2D of the tokens, D2H of out + losses inside the timed region)
roofline: dominant kernel (CUDA events around each launch in a profiled pass) against MEASURED_PEAKS.json
cpu_baseline / --impl reference: the CPU oracle port (oracle/signal_oracle.py, torch CPU ops) on the host cores.
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

METRIC = "fusion_head_fwd_bwd_samples_per_s"
UNIT = "samples/s"
L = 128
NSETS = 4            # token sets rotated through the timed region (4 x 76 MB > 126 MB L2)

# configs/<NAME>/Signal.yml of the reference: INPUT.SIZE_TRAIN / 16 -> grid, MODEL.TOPK, Gram_Loss_weight, PAT_Loss_weight,
# SOLVER.IMS_PER_BATCH; "batch" is what BASELINE.json's configs quote (B=128 on one B200; MSVR310: the YAML's 64)
CONFIGS = {
    "rgbnt201": dict(name="RGBNT201", grid=(16, 8), topk=80, w_gam=0.2, w_lam=0.2, batch=128, ref="configs/RGBNT201/Signal.yml:9-12,17"),
    "rgbnt100": dict(name="RGBNT100", grid=(8, 16), topk=112, w_gam=0.1, w_lam=0.1, batch=128, ref="configs/RGBNT100/Signal.yml:9-12,17"),
    "msvr310": dict(name="MSVR310", grid=(8, 16), topk=64, w_gam=0.2, w_lam=0.01, batch=64, ref="configs/MSVR310/Signal.yml:9-12,17,38"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm": j["hbm_gbs"], "tensor_burst": j["bf16_tflops"], "tensor": j["bf16_tflops_sustained"], "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def workload_name(cfg, d, B):
    return (f"{cfg['name']} fusion head (SIM+GAM+LAM) fwd+bwd, B={B}/GPU, 3x129 tokens, d={d}, grid {cfg['grid'][0]}x{cfg['grid'][1]}, "
            f"TOPK={cfg['topk']}, bf16")


# ---------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ---------------------------------------------------------------------------------------------
def cpu_step_fn(cfg, d, B, seed=1234, fwd_only=False):
    import torch
    from oracle import signal_oracle as so
    from signal_b200 import synthetic as syn
    sim_p = {k: v.requires_grad_(not fwd_only) for k, v in syn.make_params(syn.sim_param_shapes(d), seed).items()}
    al_p = {k: v.requires_grad_(not fwd_only) for k, v in syn.make_params(syn.align_param_shapes(d), seed + 1).items()}
    toks = [t.to(torch.bfloat16).float().requires_grad_(not fwd_only) for t in syn.make_tokens(B, d, seed=seed + 2)]
    cot = syn.make_cotangent(B, d, seed=seed + 3)
    leaves = toks + [p for p in sim_p.values()] + [p for p in al_p.values()]

    def step():
        if fwd_only:
            with torch.no_grad():
                so.head_forward(sim_p, al_p, toks, cfg["topk"], cfg["grid"][0], cfg["grid"][1])
            return
        out, gam, lam, _ = so.head_forward(sim_p, al_p, toks, cfg["topk"], cfg["grid"][0], cfg["grid"][1])
        loss = (out * cot).sum() + cfg["w_gam"] * gam + cfg["w_lam"] * lam
        torch.autograd.grad(loss, leaves, allow_unused=True)
    return step


def time_cpu(cfg, d, B, steps, warmup, threads=None, fwd_only=False):
    import torch
    cores = threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    step = cpu_step_fn(cfg, d, B, fwd_only=fwd_only)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args, cfg):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the reference is pure Python and
    /root/reference does not travel).  Honours --steps / --warmup; the batch per step is the arm's B when the whole run
    fits the time budget, else the largest power-of-two slice that does (stated in `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    d = args.dim
    B = args.batch or cfg["batch"]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget = float(os.environ.get("SIG_REF_BUDGET_S", "150"))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # probe: one step at B=8 -> seconds per sample (the oracle is linear in B apart from the B x B GAM grid)
    probe = cpu_step_fn(cfg, d, 8)
    probe()
    t0 = time.perf_counter()
    probe()
    per_sample = (time.perf_counter() - t0) / 8
    Bs = B
    while Bs > 8 and per_sample * Bs * (steps + warmup) > budget:
        Bs //= 2
    val, ms, cores = time_cpu(cfg, d, Bs, steps, warmup)
    one = time_cpu(cfg, d, min(Bs, 16), 2, 1, threads=1)
    c1 = time_cpu(CONFIGS["rgbnt201"], 768, 8, 5, 2, fwd_only=True)
    sample = (f"B={Bs} per step ({'the full batch' if Bs == B else 'a slice of the B=%d batch: the full one exceeds the %.0f s budget' % (B, budget)}), "
              f"{warmup} warm-up + {steps} timed fwd+bwd steps, fp32, torch CPU ops, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, d, B), "global_batch": B, "parallelism": "cpu",
                   "note": "CPU oracle port of the reference modules, one host process whatever --gpus says"},
        "cpu_baseline": {"value": round(val, 2), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "one_thread": {"value": round(one[0], 2), "unit": UNIT, "cores": 1, "sample": f"B={min(Bs, 16)}, 1 warm-up + 2 timed fwd+bwd steps"},
                         "config1_cpu_fwd_b8": {"value": round(c1[0], 2), "unit": UNIT, "ms_per_step": round(c1[1], 2), "cores": c1[2],
                                                "sample": "BASELINE.json configs[0]: forward only, fp32, B=8, d=768, grid 16x8, TOPK=80, torch.no_grad(), 2 warm-up + 5 timed"}},
        "e2e": {"value": round(val, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: MSVR310 full train step (random-init ViT-B/16 backbone in stock PyTorch + the head)
# ---------------------------------------------------------------------------------------------
def _torch_triplet(feat, labels):
    """soft-margin batch-hard triplet loss (layers/triplet_loss.py:16-31,51-104,121-135) in stock torch ops"""
    import torch
    import torch.nn.functional as F
    d2 = (feat ** 2).sum(1, keepdim=True)
    dist = (d2 + d2.t() - 2 * feat @ feat.t()).clamp(min=1e-12).sqrt()
    same = labels[:, None] == labels[None, :]
    ap = torch.where(same, dist, dist.new_full((), -1e30)).max(1)[0]
    an = torch.where(same, dist.new_full((), 1e30), dist).min(1)[0]
    return F.soft_margin_loss(an - ap, torch.ones_like(ap))


def _torch_head(step, patches, cls):
    """Comparison arm: the reference's head arithmetic in stock torch ops on the device (the oracle port, fp32 like the
    reference's autocast islands) + torch's own cross entropy / the triplet restatement above."""
    import torch.nn.functional as F
    from oracle import signal_oracle as so
    sim_p = dict(step.SIM.state_dict(keep_vars=True))
    al_p = dict(step.AlignM.state_dict(keep_vars=True))
    import torch
    pf, cf = [p.float() for p in patches], [c.float() for c in cls]
    # fp32 arithmetic (autocast off): under autocast torch would run these matmuls -- including the 1536-wide distance
    # matrix of the triplet loss -- in half precision, which moves the loss by several percent; the B200 kernels
    # accumulate in fp32, so the comparison arm does too
    with torch.autocast("cuda", enabled=False):
        out, _ = so.sim_forward(sim_p, pf, cf, step.cfg["topk"])
        gam = so.gam_loss(pf, al_p["contra_temp"])
        lam = so.lam_loss(al_p, pf, step.h, step.w)

    def head_loss(score, feat, target, c):
        with torch.autocast("cuda", enabled=False):
            return c["w_id"] * F.cross_entropy(score.float(), target, label_smoothing=0.1) + c["w_tri"] * _torch_triplet(feat.float(), target)
    return out.to(patches[0].dtype), gam, lam, head_loss


def run_full_step(args, cfg):
    """One complete training iteration of the reference's MSVR310 configuration (make_model.py:148-255, processor.py:165-261):
    backbone (stock PyTorch, bf16 autocast) + fusion head + four BNNeck heads + ID/triplet/GAM/LAM losses + Adam; timed
    with the B200 head and with a stock-PyTorch head (same backbone, same optimizer)."""
    import torch
    import __graft_entry__ as entry
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --full-step needs a CUDA device")
    entry.build()
    from signal_b200 import fullstep as fs, lib
    lib.load()
    dev = torch.device("cuda", 0)
    B = args.batch or cfg["batch"]
    H, W = cfg["grid"][0] * 16, cfg["grid"][1] * 16
    g = torch.Generator().manual_seed(1234)
    imgs = [torch.randn(B, 3, H, W, generator=g).to(dev) for _ in range(3)]
    ids = 16                                                    # B / NUM_INSTANCE (4) identities per batch
    target = (torch.arange(B) // (B // ids)).to(dev)
    cam = torch.randint(0, 8, (B,), generator=g).to(dev)
    out = {}
    for arm, th in (("b200_head", None), ("torch_head", _torch_head)):
        torch.manual_seed(1234)
        step = fs.SignalTrainStep(cfg["grid"], cfg["topk"], 155, 8, 512, 0.25, 1.0, cfg["w_gam"], cfg["w_lam"], torch_head=th).to(dev)
        opt = torch.optim.Adam([p for p in step.parameters() if p.requires_grad], lr=5e-6)

        def it():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = step(imgs, target, cam)
            loss.backward()
            opt.step()
            return loss
        for _ in range(max(3, args.warmup)):
            l0 = it()
        torch.cuda.synchronize()
        l_before = lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = max(3, min(args.steps, 30))
        e0.record()
        for _ in range(n):
            loss = it()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[arm] = {"ms_per_step": round(ms, 3), "samples_per_s": round(B / (ms * 1e-3), 1), "loss": round(float(loss), 5),
                    "library_launches_per_step": (lib.launch_count() - l_before) // n, "steps": n}
        del step, opt
        torch.cuda.empty_cache()
    line = {"metric": "full_train_step_samples_per_s", "value": out["b200_head"]["samples_per_s"], "unit": UNIT, "n_gpus": 1,
            "ms_per_step": out["b200_head"]["ms_per_step"], "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{cfg['name']} full Signal train step: random-init ViT-B/16 backbone (stock PyTorch, bf16 autocast, one pass over 3x{B} "
                                   f"{H}x{W} images) + fusion head (TOPK={cfg['topk']}, grid {cfg['grid'][0]}x{cfg['grid'][1]}) + 4 BNNeck heads + "
                                   f"0.25*ID + 1.0*triplet per head + {cfg['w_gam']}*GAM + {cfg['w_lam']}*LAM + Adam, B={B}",
                       "reference_config": cfg["ref"], "launch": "eager"},
            "arms": out,
            "speedup_of_the_step": round(out["torch_head"]["ms_per_step"] / out["b200_head"]["ms_per_step"], 3),
            "note": "same backbone, same optimizer in both arms; torch_head = the reference's head algorithm in stock torch ops (oracle port) + "
                    "torch cross entropy / triplet; the losses of the two arms agree to bf16 rounding (both printed)"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler: NVML in-process (brackets warm-up + timed region), nvidia-smi as a fall-back
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.sm, self.reasons, self.mx, self.power = [], set(), None, []
        self.stop_flag = False
        self.src = "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._loop_nvml, daemon=True)
        except Exception:
            self.nv = None
            self.src = "nvidia-smi"
            self.index = index
            self.th = threading.Thread(target=self._loop_smi, daemon=True)
        self.mark = None
        self.th.start()

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                return index
        return index

    def _loop_nvml(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                c = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                p = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.sm.append((time.perf_counter(), c))
                self.power.append(p)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.0005)

    def _loop_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"
        try:
            proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in proc.stdout:
            if self.stop_flag:
                break
            f = [x.strip() for x in line.split(",")]
            try:
                self.sm.append((time.perf_counter(), float(f[0])))
                self.mx = float(f[1])
                self.power.append(float(f[6]))
            except Exception:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)
        proc.terminate()

    def mark_timed(self):
        self.mark = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        self.th.join(timeout=2)
        allc = sorted(c for _, c in self.sm)
        timed = sorted(c for t, c in self.sm if self.mark is not None and t >= self.mark)
        med = lambda v: v[len(v) // 2] if v else None
        return {"sm_mhz": med(timed) if timed else med(allc), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(allc), "samples_in_timed_region": len(timed), "sm_mhz_warmup_and_timed": med(allc),
                "power_w_max": round(max(self.power), 1) if self.power else None, "source": self.src}


# ---------------------------------------------------------------------------------------------
# algorithmic work per phase (per step of B samples); DESIGN.md states the derivation
# ---------------------------------------------------------------------------------------------
def phase_work(B, d, s=2):
    """Algorithmic work per launch of the bf16 path, as implemented: {phase: (bound, work, kernel)}.
    T*s = bytes of one sample's patch tokens; the folded offset net is ONE d x d GEMM per modality.  Every phase
    listed here brackets exactly one kernel launch."""
    T = 3 * L * d            # token elements per sample
    gemm = 3 * 2 * L * d * d * B                                   # folded 1x1 conv, three modalities
    return {
        "lam_offsetnet_fwd": ("tensor", gemm, "pair_gemm_kernel (cta_group::2, 256x256 units, K-major x K-major)  H = X W'^T + b'"),
        "lam_offsetnet_bwd_dx": ("tensor", gemm, "pair_gemm_kernel (K-major x MN-major)  dX = dH W' (+GAM rows, + SIM's k-blocks under FusionHead)"),
        "lam_offsetnet_bwd_dw": ("tensor", gemm, "pair_gemm_kernel (MN-major x MN-major)  dW' = dH^T X (split-K)"),
        "lam_dwconv_fwd": ("hbm", B * T * s, "lam_dw_fwd_ring_kernel  (read H)"),
        "lam_dwconv_bwd": ("hbm", 2 * B * T * s, "lam_dw_bwd_ring_kernel  (read H, write dH)"),
        "sim_scores": ("hbm", B * T * s, "sim_scores_ring_kernel  (one pass over the tokens)"),
        "gam_pool": ("hbm", B * T * s, "pool_ring_kernel  (one pass over the tokens)"),
        "sim_attn_logits_fwd": ("hbm", B * T * s, "pipeline_kernel<32,RowsProblem>  (one pass over the tokens)"),
        "sim_attn_pool_fwd": ("hbm", B * T * s, "pipeline_kernel<32,ColsProblem>  (one pass over the tokens)"),
        "sim_attn_dlogits_bwd": ("hbm", B * T * s, "pipeline_kernel<32,RowsProblem>  (one pass over the tokens)"),
        "sim_attn_dx_bwd": ("hbm", B * T * s, "pipeline_kernel<256,DxProblem>  (write d(patches); fused into the dX GEMM under FusionHead)"),
        "sim_attn_dq_bwd": ("hbm", B * T * s, "pipeline_kernel<32,ColsProblem>  (one pass over the tokens)"),
    }


def ncu_traffic():
    """{phase: dram bytes read + written per launch} from the committed ncu --set full summary (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("per_launch_dram_bytes", {})
        except Exception:
            return {}
    return {}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args, cfg):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        opts = None
        if os.environ.get("SIG_NCCL_PRIO", "1") != "0":
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from signal_b200 import lib, modules as M, parallel, synthetic as syn
    lib.load()

    if args.check:
        if world < 2:
            raise SystemExit("bench.py --check needs --gpus >= 2 under torchrun")
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import dp_worker
        dp_worker.check_exchange(dev, rank, world)
        dp_worker.check_head(dev, rank, world)
        if rank == 0:
            print(json.dumps({"check": "ok", "n_gpus": world, "what": "NVLink exchange kernel == ncclAllReduce; exchanged head gradients == mean of the "
                              "shard-local gradients (tests/dp_worker.py)"}), flush=True)
        torch.cuda.synchronize()
        os._exit(0)

    d = args.dim
    B = args.batch or cfg["batch"]
    if args.strong and world > 1:
        B = max(1, 1024 // world)
    GRID, TOPK, W_GAM, W_LAM = cfg["grid"], cfg["topk"], cfg["w_gam"], cfg["w_lam"]
    sampler = ClockSampler(local) if rank == 0 else None      # brackets warm-up + timed region

    sim = M.Select_Interactive_Module(d, k=TOPK)
    al = M.AlignmentM(d, GRID[0], GRID[1])
    sim.load_state_dict(syn.make_params(syn.sim_param_shapes(d), 1234))
    al.load_state_dict(syn.make_params(syn.align_param_shapes(d), 1235))
    sim, al = sim.to(dev), al.to(dev)
    params = [p for p in list(sim.parameters()) + list(al.parameters())]
    ncoll = [0]     # exchange calls issued by the last step

    def allreduce_grads():
        if args.allreduce and not overlap:
            ncoll[0] = parallel.allreduce_param_grads(params, world)

    host_sets = []
    for k in range(NSETS):
        toks = syn.make_tokens(B, d, seed=1234 + 17 * k + 1000 * rank, dtype=torch.bfloat16)
        host_sets.append([t.pin_memory() for t in toks])
    dev_sets = [[t.to(dev).requires_grad_(True) for t in hs] for hs in host_sets]
    cot = syn.make_cotangent(B, d).to(dev, torch.bfloat16)
    wg = torch.tensor(W_GAM, device=dev)
    wl = torch.tensor(W_LAM, device=dev)

    head = M.FusionHead(sim, al) if args.fused else None
    overlap = head is not None and world > 1 and args.allreduce and args.overlap
    exchange = None
    nsync = [0]
    auto_exchange = args.exchange == "auto"
    if auto_exchange:
        args.exchange = "nvlink"
    if overlap:   # pieces of the gradient arena are averaged over the ranks inside the backward as soon as they are final
        if args.exchange == "nvlink":
            # symmetric memory is torch plumbing that a box may not offer (no NVLink peer access, an older driver): with
            # --exchange auto every rank then agrees to fall back to ncclAllReduce instead of failing the run
            try:
                exchange = parallel.GradExchange(head.grad_numel(), dev)
                ok = torch.ones(1, device=dev)
            except Exception as e:   # noqa: BLE001
                if not auto_exchange:
                    raise
                print(f"[bench] rank {rank}: NVLink exchange unavailable ({type(e).__name__}: {e}); falling back to NCCL", file=sys.stderr)
                exchange, ok = None, torch.zeros(1, device=dev)
            if auto_exchange:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if float(ok) < 1.0:
                    exchange, args.exchange = None, "nccl"
        if exchange is not None:
            head.grad_arena = exchange.arena

            def _sync(flat):
                nsync[0] += 1
                exchange.allreduce(flat)
        else:
            def _sync(flat):
                nsync[0] += 1
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        head.grad_sync = _sync

    def fwd_bwd(toks, use_head=True):
        patches = [t[:, 1:] for t in toks]
        cls = [t[:, 0] for t in toks]
        if args.only == "sim":   # (diagnostic) one module only
            out = sim(*patches, *cls)
            torch.autograd.backward([out], [cot])
            return out, wg, wl
        if args.only == "align":
            gam, lam = al(*patches, stage="together_CLS_Patch")
            torch.autograd.backward([gam, lam], [wg, wl])
            return cot, gam, lam
        if head is not None and use_head:     # one call: AlignM on a side stream next to SIM, one token-gradient writer
            out, gam, lam = head(*patches, *cls, stage="together_CLS_Patch")
        else:                                 # the reference's two consecutive module calls (make_model.py:191,205)
            out = sim(*patches, *cls)
            gam, lam = al(*patches, stage="together_CLS_Patch")
        if overlap:
            nsync[0] = 0
        torch.autograd.backward([out, gam, lam], [cot, wg, wl])
        if overlap:
            ncoll[0] = nsync[0]
        return out, gam, lam

    def clear_grads(toks):
        for t in toks:
            t.grad = None
        for p in params:
            p.grad = None

    def eager_step(i, use_head=True):
        toks = dev_sets[i % NSETS]
        clear_grads(toks)
        res = fwd_bwd(toks, use_head)
        if world > 1:
            allreduce_grads()
        return res

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def capture_graphs():
        graphs, nl = [], None
        for k in range(NSETS):
            clear_grads(dev_sets[k])
            g = torch.cuda.CUDAGraph()
            lb = lib.launch_count()
            with torch.cuda.graph(g):
                fwd_bwd(dev_sets[k])
                if world > 1 and args.allreduce_in_graph:
                    allreduce_grads()
            nl = lib.launch_count() - lb
            graphs.append((g, parallel.grad_arenas(params)))
        return graphs, nl

    def timed(step_fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for i in range(n):
            step_fn(i)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W = max(args.warmup, 3)
    for i in range(W):
        eager_step(i)
    sync_all()
    graph_launches = None
    step = eager_step
    if args.graph:
        # one CUDA graph per token set (static inputs); replay = the same module calls without per-launch host work.
        # The library is capturable: no syncs, no allocations; the exchange kernel / NCCL calls are captured too.
        graphs, graph_launches = capture_graphs()

        def step(i):
            g, arenas = graphs[i % NSETS]
            g.replay()
            if world > 1 and not args.allreduce_in_graph and args.allreduce and not overlap:
                ncoll[0] = parallel.allreduce_arenas(arenas, world)

        for i in range(W):
            step(i)
        sync_all()
    l0 = lib.launch_count()
    if sampler:
        sampler.mark_timed()
    ms_total = timed(step, args.steps)
    launches = lib.launch_count() - l0 if graph_launches is None else graph_launches * args.steps
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)

    # ---- N > 1: the same step without any gradient exchange -> what the exchange costs on top of the compute
    exposed_us = None
    if world > 1 and overlap and args.graph:
        hook, arena = head.grad_sync, head.grad_arena
        head.grad_sync, head.grad_arena = None, None
        overlap_saved, overlap = overlap, False
        args_allreduce, args.allreduce = args.allreduce, False
        g2, _ = capture_graphs()

        def step_nox(i):
            g2[i % NSETS][0].replay()
        for i in range(W):
            step_nox(i)
        ms_nox = timed(step_nox, args.steps) / args.steps
        exposed_us = round((ms_step - ms_nox) * 1e3, 1)
        head.grad_sync, head.grad_arena, overlap, args.allreduce = hook, arena, overlap_saved, args_allreduce
        del g2

    # ---- variants (N = 1): the same step without graph replay -- through FusionHead, and through the reference's own
    # call sequence SIM(...); AlignM(...) (make_model.py:191,205)
    variants = None
    if world == 1 and not args.no_variants and not args.only:
        nv = max(3, min(args.steps, 50))
        variants = {}
        for name, use_head in (("eager_fusion_head", True), ("eager_two_module_calls", False)):
            if use_head and head is None:
                continue
            fn = lambda i, u=use_head: eager_step(i, u)
            for i in range(3):
                fn(i)
            ms = timed(fn, nv) / nv
            variants[name] = {"value": round(B / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms, 4), "steps": nv,
                              "api": "FusionHead(SIM, AlignM)(...)" if use_head else "SIM(...); AlignM(...)  (reference call sequence)",
                              "launch": "eager (host-bound: ctypes calls + torch allocations per step)"}

        if head is not None:
            # the same step through FusionHead.make_graphed: forward and backward graphs behind the ordinary autograd API
            # (inputs are copied into the graphs' static buffers every call -- part of the measured time)
            graphed = head.make_graphed(*dev_sets[0])

            def gstep(i):
                toks = dev_sets[i % NSETS]
                clear_grads(toks)
                out, gam, lam = graphed(*toks)
                torch.autograd.backward([out, gam, lam], [cot, wg, wl])
            for i in range(3):
                gstep(i)
            ms = timed(gstep, nv) / nv
            variants["graphed_callable"] = {"value": round(B / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms, 4), "steps": nv,
                                            "api": "FusionHead.make_graphed(...)(rgb, ni, ti); autograd.backward(...)",
                                            "launch": "two CUDA graphs (forward, backward) replayed by autograd; inputs copied into static buffers"}

    # ---- e2e: pinned HOST token maps -> device -> the same module calls -> results back on the host, every step.
    # Double buffered: the H2D copy of step i+1 runs on a copy stream while step i computes; each step ends
    # with the D2H read of its fused feature + the two losses and a host-side wait for them.
    e2e_steps = max(3, min(args.steps, 50))
    h2d = 3 * B * (L + 1) * d * 2
    d2h = B * 3 * d * 2 + 8
    copy_stream = torch.cuda.Stream(dev)
    main_stream = torch.cuda.current_stream(dev)
    stage = [[torch.empty_like(t, device=dev).requires_grad_(True) for t in host_sets[0]] for _ in range(2)]
    out_host = [torch.empty(B, 3 * d, dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    loss_host = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    for ev in done:
        ev.record(main_stream)

    def run_stage(k):
        clear_grads(stage[k])
        res = fwd_bwd(stage[k])
        if world > 1:
            allreduce_grads()
        return res

    for k in range(2):          # warm up on the staging buffers, then (graph mode) capture one graph per buffer
        with torch.no_grad():
            for dst, src in zip(stage[k], host_sets[k]):
                dst.copy_(src)
        run_stage(k)
    sync_all()
    runners = []
    for k in range(2):
        if args.graph:
            clear_grads(stage[k])
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                res = fwd_bwd(stage[k])
                if world > 1 and (args.allreduce_in_graph or overlap):
                    allreduce_grads()
                losses2 = torch.stack([res[1], res[2]])
            runners.append((g.replay, res[0], losses2, parallel.grad_arenas(params)))
        else:
            runners.append((None, None, None, None))
    sync_all()

    def issue_copy(i):
        k = i % 2
        copy_stream.wait_event(done[k])                 # the step that last read this buffer has finished
        with torch.cuda.stream(copy_stream), torch.no_grad():
            for dst, src in zip(stage[k], host_sets[i % NSETS]):
                dst.copy_(src, non_blocking=True)
            ready[k].record(copy_stream)

    def e2e_run(n):
        issue_copy(0)
        for i in range(n):
            k = i % 2
            if i + 1 < n:
                issue_copy(i + 1)                       # overlaps this step's compute
            main_stream.wait_event(ready[k])
            replay, out_k, loss_k, arenas = runners[k]
            if replay is not None:
                replay()
                if world > 1 and args.allreduce and not overlap and not args.allreduce_in_graph:
                    parallel.allreduce_arenas(arenas, world)
            else:
                out_k, gam_k, lam_k = run_stage(k)
                loss_k = torch.stack([gam_k, lam_k])
            out_host[k].copy_(out_k.detach(), non_blocking=True)
            loss_host[k].copy_(loss_k.detach(), non_blocking=True)
            done[k].record(main_stream)
            done[k].synchronize()                       # this step's results are on the host

    e2e_run(4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    e2e_run(e2e_steps)
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = B * world * e2e_steps / (float(t.item()) * 1e-3)

    # ---- profiled pass (rank 0): CUDA events recorded by the library around each phase, on the stream that runs
    # it, with the two modules called one after the other (no cross-module overlap) -> per-kernel rooflines.
    roof, phases, kernels = None, None, None
    if rank == 0:
        prof_steps = 10
        if head is not None:
            head.grad_sync = None            # rank-local pass: no collectives
            head.grad_arena = None

        def seq_step(toks):
            patches = [t[:, 1:] for t in toks]
            cls = [t[:, 0] for t in toks]
            out = sim(*patches, *cls)
            gam, lam = al(*patches, stage="together_CLS_Patch")
            torch.autograd.backward([out, gam, lam], [cot, wg, wl])

        for i in range(2):
            seq_step(dev_sets[i % NSETS])
        torch.cuda.synchronize()
        lib.profile_enable(True)
        for i in range(prof_steps):          # local work only: the other ranks are not in this pass
            toks = dev_sets[i % NSETS]
            clear_grads(toks)
            seq_step(toks)
        torch.cuda.synchronize()
        lib.profile_enable(False)
        prof = lib.profile_collect()
        pk = peaks()
        work = phase_work(B, d)
        phases = {k: round(v[0] / prof_steps * 1e3, 1) for k, v in prof.items()}   # us per step
        traffic = ncu_traffic()
        kernels = {}
        for k, (bound, w, kname) in work.items():
            if k not in phases or phases[k] <= 0:
                continue
            sec = phases[k] * 1e-6
            # single launches timed on their own (eager pass, GPU otherwise idle, clocks at their maximum): the BURST
            # tensor peak is the denominator; the fraction of the sustained peak is given beside it
            if bound == "hbm":
                ach, peak, unit = w / sec / 1e9, pk["hbm"], "GB/s"
                extra = {}
            else:
                ach, peak, unit = w / sec / 1e12, pk["tensor_burst"], "TFLOP/s"
                extra = {"frac_of_sustained_peak": round(ach / pk["tensor"], 4)}
            kernels[k] = dict({"kernel": kname, "bound": bound, "us": phases[k], "achieved": round(ach, 1), "peak": peak, "unit": unit,
                               "frac": round(ach / peak, 4), "traffic": traffic.get(k)}, **extra)
        # dominant kernel of the step: the tcgen05 GEMM of the LAM offset net (three launches per step: H = X W'^T,
        # dX = dH W', dW' = dH^T X -- one template, one algorithmic work figure)
        gem = [kernels[k] for k in ("lam_offsetnet_fwd", "lam_offsetnet_bwd_dx", "lam_offsetnet_bwd_dw") if k in kernels]
        if gem:
            us = sum(g["us"] for g in gem) / len(gem)
            w = work["lam_offsetnet_fwd"][1]
            ach = w / (us * 1e-6) / 1e12
            tr = [g["traffic"] for g in gem if g["traffic"]]
            roof = {"kernel": "pair_gemm_kernel (tcgen05 cta_group::2, 256x256 units; LAM offset-net GEMMs: H = X W'^T, dX = dH W', dW' = dH^T X)",
                    "bound": "tensor", "achieved": round(ach, 1), "peak": pk["tensor_burst"], "unit": "TFLOP/s", "frac": round(ach / pk["tensor_burst"], 4),
                    "frac_of_sustained_peak": round(ach / pk["tensor"], 4),
                    "traffic": round(sum(tr) / len(tr)) if tr else None,
                    "peak_source": pk["src"] + ": burst cuBLAS bf16 %.0f TFLOP/s (kernel timed alone at full clocks); sustained %.0f" % (pk["tensor_burst"], pk["tensor"]),
                    "launches_per_step": len(gem), "avg_launch_us": round(us, 1),
                    "algorithmic_flops_per_launch": w, "share_of_step": round(sum(g["us"] for g in gem) / (ms_step * 1e3), 3),
                    "how": "CUDA events around each launch on its stream, modules run back to back (eager), mean of %d steps; "
                           "traffic = dram bytes read+written per launch from profiles/ (ncu --set full)" % prof_steps}
    if world > 1:
        dist.barrier()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            Bs = min(B, 128)
            v, ms, cores = time_cpu(cfg, d, Bs, 3, 1)
            one = time_cpu(cfg, d, 16, 2, 1, threads=1)
            c1 = time_cpu(CONFIGS["rgbnt201"], 768, 8, 5, 2, fwd_only=True)
            cpu = {"value": round(v, 2), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"B={Bs} per step, 1 warm-up + 3 timed fwd+bwd steps of the CPU oracle port (torch CPU ops, fp32, all host threads)",
                   "one_thread": {"value": round(one[0], 2), "unit": UNIT, "cores": 1, "sample": "B=16, 1 warm-up + 2 timed fwd+bwd steps"},
                   "config1_cpu_fwd_b8": {"value": round(c1[0], 2), "unit": UNIT, "ms_per_step": round(c1[1], 2), "cores": c1[2],
                                          "sample": "BASELINE.json configs[0]: forward only, fp32, B=8, d=768, grid 16x8, TOPK=80, torch.no_grad(), 2 warm-up + 5 timed"}}
        if world > 1:
            if not args.allreduce:
                xchg = "none (--no-allreduce)"
            elif overlap:
                how = ("sig_xchg_allreduce_f32: the library's own two-shot all-reduce kernel over NVLink symmetric memory"
                       + (" with NVLS multimem.ld_reduce / multimem.st" if exchange is not None and exchange.multicast else " (peer loads/stores)")) \
                    if exchange is not None else "ncclAllReduce(avg)"
                xchg = (f"{how}, {ncoll[0]} calls per step on a communication stream inside the backward, each as soon as its piece of the flat "
                        f"gradient arena is final (FusionHead.grad_sync)" + (", captured in the step graph" if args.graph else ""))
            else:
                xchg = f"ncclAllReduce(avg) of the flat gradient arenas after the backward, {ncoll[0]} per step"
        else:
            xchg = "n/a"
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(cfg, d, B), "global_batch": B * world, "parallelism": f"dp{world}", "reference_config": cfg["ref"],
                       "loss_weights": {"gam": W_GAM, "lam": W_LAM},
                       "launch": "cuda_graph_replay" if args.graph else "eager", "api": "FusionHead(SIM, AlignM)" if args.fused else "SIM(...); AlignM(...)",
                       "l2": f"inputs rotate over {NSETS} token sets ({NSETS * 3 * B * (L + 1) * d * 2 / 1e6:.0f} MB > 126 MB L2)",
                       "grad_exchange": xchg},
            "e2e": {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "how": "pinned host tokens, H2D of step i+1 on a copy stream under step i's compute (2 staging buffers), "
                           + ("graph replay of the module calls" if args.graph else "eager module calls")
                           + ", D2H of out + losses and a host wait every step"
                           + ("; every rank feeds its own tokens over the host's shared PCIe / memory paths, which bounds this number at N > 1" if world > 1 else "")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "variants": variants,
            "exposed_exchange_us_per_step": exposed_us,
            "kernels": kernels,
            "phases_us": phases,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # the captured graphs hold NCCL work: drop them before the communicator, and do not let a slow
        # communicator teardown keep the launcher waiting after the result line is out
        sys.stdout.flush()
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="rgbnt201", choices=sorted(CONFIGS), help="reference configuration whose head shape is run")
    ap.add_argument("--dim", type=int, default=768, choices=[512, 768])
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (default: the configuration's, 128 / 64)")
    ap.add_argument("--strong", action="store_true", help="N>1: global batch 1024 split over the ranks (BASELINE.json configs[4])")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--exchange", default=os.environ.get("SIG_EXCHANGE", "auto"), choices=["auto", "nvlink", "nccl"],
                    help="N>1: gradient exchange inside the backward: the library's NVLink/NVLS kernel (auto) or ncclAllReduce.  "
                         "Measured (profiles/r2_exchange_*.json, r2_eager_dp.json): with AlignM's eager backward chain the kernel "
                         "wins at every N (0.637 vs 0.664 ms at N=2, 0.634 vs 0.718 at N=4, 0.683 vs 0.734 at N=8 without the chain)")
    ap.add_argument("--check", action="store_true", help="N>1: numerical check of the data-parallel path (no timing)")
    ap.add_argument("--full-step", action="store_true",
                    help="BASELINE.json configs[3]: time one complete training iteration (backbone + head + losses + Adam) with the B200 "
                         "head and with a stock-PyTorch head (use with --config msvr310 --dim 512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the eager variants")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="time eager module calls instead of CUDA-graph replays")
    ap.add_argument("--allreduce-eager", dest="allreduce_in_graph", action="store_false",
                    help="(diagnostic) issue the NCCL all-reduce after each graph replay instead of capturing it in the step graph")
    ap.add_argument("--only", default="", choices=["", "sim", "align"], help="(diagnostic) time one module's fwd+bwd only")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="(diagnostic) all-reduce after the backward instead of inside it (FusionHead.grad_sync)")
    ap.add_argument("--no-allreduce", dest="allreduce", action="store_false", help="(diagnostic) skip the gradient exchange at N>1")
    ap.add_argument("--no-fused", dest="fused", action="store_false",
                    help="call SIM and AlignM one after the other instead of through signal_b200.FusionHead")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.full_step:
        run_full_step(args, cfg)
    else:
        run_gpu(args, cfg)


if __name__ == "__main__":
    main()
