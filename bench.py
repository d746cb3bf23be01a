#!/usr/bin/env python
"""Fusion-head fwd+bwd throughput on B200 (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dim 768] [--impl reference]

One "step" = Select_Interactive_Module fwd + AlignmentM fwd (GAM + LAM) + backward to the three
[B,129,d] token maps and all head parameters, on one synthetic RGBNT201 batch (B=128 per GPU,
bf16 tokens, fp32 master parameters).  N>1 (launched by torchrun, one rank per GPU): the batch
is sharded (weak scaling, B=128 per rank) and the head gradients are all-reduced with NCCL.

value   : samples/s, tokens already resident in HBM, CUDA-event timed, max over ranks
e2e     : same step through the nn.Module API starting from pinned HOST token maps (H2D copy of
          the tokens and D2H read of out + losses inside the timed region)
roofline: dominant phase of the step (CUDA events recorded by the library around each phase in a
          profiled pass right after the timed region) against MEASURED_PEAKS.json
cpu_baseline / --impl reference: the CPU oracle port (oracle/signal_oracle.py, torch CPU ops, all
          host threads) on a bounded sample of the same workload.
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
B_PER_GPU = 128
L = 128
GRID = (16, 8)
TOPK = 80
NSETS = 4            # token sets rotated through the timed region (4 x 76 MB > 126 MB L2)
W_GAM, W_LAM = 0.2, 0.2   # configs/RGBNT201/Signal.yml:9-10


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm": j["hbm_gbs"], "tensor_burst": j["bf16_tflops"], "tensor": j["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "src": "fallback"}


def workload_name(d):
    return f"RGBNT201 fusion head (SIM+GAM+LAM) fwd+bwd, B={B_PER_GPU}/GPU, 3x129 tokens, d={d}, grid 16x8, TOPK={TOPK}, bf16"


# ---------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ---------------------------------------------------------------------------------------------
def cpu_step_fn(d, B, seed=1234):
    import torch
    from oracle import signal_oracle as so
    from signal_b200 import synthetic as syn
    sim_p = {k: v.requires_grad_(True) for k, v in syn.make_params(syn.sim_param_shapes(d), seed).items()}
    al_p = {k: v.requires_grad_(True) for k, v in syn.make_params(syn.align_param_shapes(d), seed + 1).items()}
    toks = [t.to(torch.bfloat16).float().requires_grad_(True) for t in syn.make_tokens(B, d, seed=seed + 2)]
    cot = syn.make_cotangent(B, d, seed=seed + 3)
    leaves = toks + [p for p in sim_p.values()] + [p for p in al_p.values()]

    def step():
        out, gam, lam, _ = so.head_forward(sim_p, al_p, toks, TOPK, GRID[0], GRID[1])
        loss = (out * cot).sum() + W_GAM * gam + W_LAM * lam
        torch.autograd.grad(loss, leaves, allow_unused=True)
    return step


def time_cpu(d, B, steps, warmup):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(d, B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Bs = 32
    steps = max(1, min(args.steps, 6))
    warmup = max(1, min(args.warmup, 2))
    val, ms, cores = time_cpu(args.dim, Bs, steps, warmup)
    sample = f"B={Bs} slice of the workload per step, {warmup} warm-up + {steps} timed fwd+bwd steps, fp32, torch CPU ops"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.dim), "note": "CPU oracle port of the reference modules (the reference is "
                   "pure Python and /root/reference does not travel to the GPU box)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
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
        # phase: (bound, work [bytes for hbm, flops for tensor], kernel)
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
def run_gpu(args):
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
            # the gradient exchange runs next to GPU-filling kernels: NCCL's CTAs must be dispatched as soon as an
            # SM has room, not after the compute kernels launched before them have drained
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from signal_b200 import lib, modules as M, synthetic as syn
    lib.load()

    d, B = args.dim, B_PER_GPU
    sim = M.Select_Interactive_Module(d, k=TOPK)
    al = M.AlignmentM(d, GRID[0], GRID[1])
    sim.load_state_dict(syn.make_params(syn.sim_param_shapes(d), 1234))
    al.load_state_dict(syn.make_params(syn.align_param_shapes(d), 1235))
    sim, al = sim.to(dev), al.to(dev)
    params = [p for p in list(sim.parameters()) + list(al.parameters())]
    # each backward returns its parameter gradients as views of a flat fp32 arena (FusionHead: one arena for both
    # modules, exchanged in two pieces inside the backward; separate module calls: one arena per module)

    from signal_b200 import parallel
    ncoll = [0]     # collectives issued by the last step

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
    if overlap:   # pieces of the gradient arenas are all-reduced inside the backward as soon as they are final
        nsync = [0]

        skip = int(os.environ.get("SIG_SYNC_SKIP", "0"))   # (diagnostic bit mask: leave out piece 0/1/2 of the exchange)

        def _sync(flat):
            nsync[0] += 1
            if not (skip >> (nsync[0] - 1)) & 1:
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        head.grad_sync = _sync

    def fwd_bwd(toks):
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
        if head is not None:     # one call: AlignM on a side stream next to SIM, one token-gradient writer
            out, gam, lam = head(*patches, *cls, stage="together_CLS_Patch")
        else:                    # the reference's two consecutive module calls (make_model.py:191,205)
            out = sim(*patches, *cls)
            gam, lam = al(*patches, stage="together_CLS_Patch")
        if overlap:
            nsync[0] = 0
        torch.autograd.backward([out, gam, lam], [cot, wg, wl])
        if overlap:
            ncoll[0] = nsync[0]
        return out, gam, lam

    def step(i):
        toks = dev_sets[i % NSETS]
        for t in toks:
            t.grad = None
        for p in params:
            p.grad = None
        res = fwd_bwd(toks)
        if world > 1:
            allreduce_grads()
        return res

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    sync_all()
    graph_launches = None
    if args.graph:
        # capture one CUDA graph per token set (static inputs); replay = the same module calls,
        # without per-launch host work.  The library is capturable: no syncs, no allocations.
        graphs = []
        for k in range(NSETS):
            for t in dev_sets[k]:
                t.grad = None
            for p in params:
                p.grad = None
            g = torch.cuda.CUDAGraph()
            lb = lib.launch_count()
            with torch.cuda.graph(g):
                fwd_bwd(dev_sets[k])
                if world > 1 and args.allreduce_in_graph:
                    allreduce_grads()
            graph_launches = lib.launch_count() - lb
            graphs.append((g, parallel.grad_arenas(params)))

        def step(i):
            g, arenas = graphs[i % NSETS]
            g.replay()
            if world > 1 and not args.allreduce_in_graph and args.allreduce and not overlap:
                ncoll[0] = parallel.allreduce_arenas(arenas, world)

        for i in range(max(args.warmup, 3)):
            step(i)
        sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    launches = lib.launch_count() - l0 if graph_launches is None else graph_launches * args.steps
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)

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
        for t in stage[k]:
            t.grad = None
        for p_ in params:
            p_.grad = None
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
            for t in stage[k]:
                t.grad = None
            for p_ in params:
                p_.grad = None
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
            for t in toks:
                t.grad = None
            for p_ in params:
                p_.grad = None
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
            if bound == "hbm":
                ach, peak, unit = w / sec / 1e9, pk["hbm"], "GB/s"
            else:
                ach, peak, unit = w / sec / 1e12, pk["tensor"], "TFLOP/s"
            kernels[k] = {"kernel": kname, "bound": bound, "us": phases[k], "achieved": round(ach, 1), "peak": peak, "unit": unit,
                          "frac": round(ach / peak, 4), "traffic": traffic.get(k)}
        # dominant kernel of the step: the tcgen05 GEMM of the LAM offset net (three launches per step: H = X W'^T,
        # dX = dH W', dW' = dH^T X -- one template, one algorithmic work figure)
        gem = [kernels[k] for k in ("lam_offsetnet_fwd", "lam_offsetnet_bwd_dx", "lam_offsetnet_bwd_dw") if k in kernels]
        if gem:
            us = sum(g["us"] for g in gem) / len(gem)
            w = work["lam_offsetnet_fwd"][1]
            ach = w / (us * 1e-6) / 1e12
            tr = [g["traffic"] for g in gem if g["traffic"]]
            roof = {"kernel": "pair_gemm_kernel (tcgen05 cta_group::2, 256x256 units; LAM offset-net GEMMs: H = X W'^T, dX = dH W', dW' = dH^T X)",
                    "bound": "tensor", "achieved": round(ach, 1), "peak": pk["tensor"], "unit": "TFLOP/s", "frac": round(ach / pk["tensor"], 4),
                    "traffic": round(sum(tr) / len(tr)) if tr else None, "peak_source": pk["src"] + " (sustained cuBLAS bf16; burst %.0f)" % pk["tensor_burst"],
                    "launches_per_step": len(gem), "avg_launch_us": round(us, 1),
                    "algorithmic_flops_per_launch": w, "share_of_step": round(sum(g["us"] for g in gem) / (ms_step * 1e3), 3),
                    "how": "CUDA events around each launch on its stream, modules run back to back (eager), mean of %d steps; "
                           "traffic = dram bytes read+written per launch from profiles/ (ncu --set full)" % prof_steps}
    if world > 1:
        dist.barrier()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            Bs = 32
            v, ms, cores = time_cpu(d, Bs, 3, 1)
            cpu = {"value": round(v, 2), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"B={Bs} slice of the workload, 1 warm-up + 3 timed fwd+bwd steps of the CPU oracle port, fp32"}
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(d), "global_batch": B * world, "parallelism": f"dp{world}",
                       "launch": "cuda_graph_replay" if args.graph else "eager", "api": "FusionHead(SIM, AlignM)" if args.fused else "SIM(...); AlignM(...)", "l2": f"inputs rotate over {NSETS} token sets ({NSETS * 3 * B * (L + 1) * d * 2 / 1e6:.0f} MB > 126 MB L2)",
                       "grad_allreduce": (f"NCCL all-reduce (avg) of the flat head-gradient arenas, {ncoll[0]} collectives per step, "
                                          + ("issued inside the backward on a communication stream as each piece becomes final (FusionHead.grad_sync), " if overlap else "after the backward, ")
                                          + ("captured in the step graph" if args.graph and (args.allreduce_in_graph or overlap) else "eager")) if world > 1 else "n/a"},
            "e2e": {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "how": "pinned host tokens, H2D of step i+1 on a copy stream under step i's compute (2 staging buffers), "
                           + ("graph replay of the module calls" if args.graph else "eager module calls")
                           + ", D2H of out + losses and a host wait every step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
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
    ap.add_argument("--dim", type=int, default=768, choices=[512, 768])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="time eager module calls instead of CUDA-graph replays")
    ap.add_argument("--allreduce-eager", dest="allreduce_in_graph", action="store_false",
                    help="(diagnostic) issue the NCCL all-reduce after each graph replay instead of capturing it in the step graph")
    ap.add_argument("--only", default="", choices=["", "sim", "align"], help="(diagnostic) time one module's fwd+bwd only")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="(diagnostic) all-reduce after the backward instead of inside it (FusionHead.grad_sync)")
    ap.add_argument("--no-allreduce", dest="allreduce", action="store_false", help="(diagnostic) skip the gradient all-reduce at N>1")
    ap.add_argument("--no-fused", dest="fused", action="store_false",
                    help="call SIM and AlignM one after the other instead of through signal_b200.FusionHead")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
