#!/usr/bin/env python
"""Benchmark of the Tiny-NewsRec hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W [--workload kd4|kd2|table|table_long|eval] [--no-graph] [--impl reference]

Default workload ``kd4`` is BASELINE.json configs[1]: the 4-layer student, finetuning-stage
multi-teacher KD step (forward 4 layers + backward layers {2,3} + fused Adam(amsgrad)), M=4
teachers, per-GPU batch 32, history 50, npratio 4, title 30 tokens, bf16 compute, synthetic
MIND-shaped data, random-init weights.  A "step" is one such train step on one batch.

One JSON line is printed by rank 0.  ``value`` is device-timed whole-job impressions/s with
inputs resident in HBM; ``e2e`` is the same metric through the public API with pinned HOST inputs
copied to the device every step and the loss read back.  The train step runs the way the public
API runs it at speed: ``tinyrec.run.GraphedTrainStep`` (zero_grad + Model.forward + backward +
bucketed NCCL all-reduce + fused Adam captured once into a CUDA graph, replayed per batch;
``--no-graph`` launches every step eagerly; ``config.launch`` says which ran).  ``roofline`` comes
from CUDA events around every GEMM launch over eagerly launched steps right after the timed region
(events cannot be recorded inside a replay).  The secondary workloads add one more leg through the
public drivers on the whole BASELINE-sized input: ``full_table`` (``tinyrec.run.build_news_table`` over
the 161k-row table) and ``full_dev_set`` (``tinyrec.run.evaluate`` over 376 471 impressions).
``--impl reference`` times the CPU oracle port of the reference (all host threads).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "kd4": dict(layers=4, trainable=[2, 3], name="Tiny-NewsRec 4-layer student finetuning-stage KD, M=4 teachers"),
    "kd2": dict(layers=2, trainable=[0, 1], name="Tiny-NewsRec 2-layer student finetuning-stage KD, M=4 teachers"),
}
B, H, K, L, M, D, N_NEWS = 32, 50, 5, 30, 4, 256, 161013
N_BATCHES = 8
# secondary workloads (BASELINE.json configs[3], configs[4]); one "step" = one batch of rows / impressions
TABLE_WORKLOADS = {
    "table": dict(layers=12, L=30, rows=4096, mean_len=14.0, std_len=4.0, min_len=4,
                  name="get_teacher_emb: 12-layer teacher news-table build, title rows (30 tokens)"),
    "table_long": dict(layers=12, L=180, rows=1024, mean_len=140.0, std_len=40.0, min_len=10,
                       name="get_teacher_emb: 12-layer teacher news-table build, title+abstract+body rows (180 tokens)"),
}
EVAL_WORKLOAD = dict(imps=4096, name="Impression scoring eval from a precomputed news table: gather + user attention "
                                     "+ dot scoring + AUC/MRR/nDCG@5/10")


def flops_per_step(layers, n_train, tokens):
    """per token per layer fwd 14 247 936 FLOP; bwd of a trainable layer 2x,
    lowest trainable layer omits the QKV dgrad (3 538 944); heads 307 200 /token + 393 216 /news."""
    fwd = 14247936 * layers
    bwd = 28495872 * n_train - 3538944
    return tokens * (fwd + bwd) + tokens * 307200 * 3 + (tokens // L) * 393216 * 3


def ncu_traffic():
    """DRAM bytes per GEMM launch (dram__bytes_read.sum + dram__bytes_write.sum averaged over the GEMM launches
    of one kd4 step) from the newest committed `ncu --set full` capture -> (bytes, file) or (None, None)."""
    for name in ("r02_gemm_ncu_full.json", "r01_s5_gemm_ncu_full.json", "r01_gemm_ncu_full_v7.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)["traffic_bytes_per_launch"]), "profiles/" + name
        except Exception:  # noqa: BLE001
            continue
    return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML from a thread every 20 ms;
    `nvidia-smi -lms` takes longer to start than a short timed region lasts)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread, self.err = index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, rs))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.02)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nv is None or self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable: " + str(self.err)]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        sm = [m for m, _ in self.samples]
        reasons = set()
        for _, rs in self.samples:
            for bit, nm in self.REASONS.items():
                if rs & bit:
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- problem setup
def make_inputs(rank, device):
    """The device-resident tables (news tokens int32 [N+1, 2L], M teacher tables fp32 [N+1, D]) with their batcher,
    N_BATCHES distinct INDEX batches (hist_idx, hist_mask, cand_idx, label) on the device and the same in pinned host
    memory -- what the training loader ships per step (dataloader.py:118-172 after id -> row mapping)."""
    import tinyrec.dataloader as dl
    import tinyrec.synth as synth
    news = synth.news_table(N_NEWS, L=L, seed=1234)
    g = torch.Generator(device=device).manual_seed(4321)
    teachers = [(torch.randn(N_NEWS + 1, D, generator=g, device=device) * 0.1) for _ in range(M)]
    tables = dl.DeviceTables(news, [], device)
    tables.teachers = teachers
    hist_idx, hmask, cand_idx, label = synth.train_impressions(B * N_BATCHES, N_NEWS, H, K, seed=1234 + rank)
    batcher = dl.TrainBatcher(tables, B, H, K)
    dev_batches, host_batches = [], []
    for i in range(N_BATCHES):
        sl = slice(i * B, (i + 1) * B)
        hb = tuple(torch.from_numpy(np.ascontiguousarray(a[sl])).pin_memory() for a in (hist_idx, hmask, cand_idx, label))
        host_batches.append(hb)
        dev_batches.append(tuple(t.to(device) for t in hb))
    return batcher, dev_batches, host_batches


def assembled(batcher, idx_batch, clone=False):
    """index batch -> the six tensors Model.forward takes (row gathers on the device)."""
    hist_idx, hmask, cand_idx, label = idx_batch
    history, candidate, th, tc = batcher.assemble(hist_idx, cand_idx)
    if clone:
        return (history.clone(), hmask, candidate.clone(), label, [t.clone() for t in th], [t.clone() for t in tc])
    return (history, hmask, candidate, label, th, tc)


def make_model(layers, trainable, device):
    import tinyrec.model_bert as mb
    import tinyrec.synth as synth
    args = synth.demo_args(num_student_layers=layers, num_teachers=M, bert_trainable_layer=trainable)
    torch.manual_seed(0)
    model = mb.Model(args)
    model.load_state_dict(synth.kd_model_state(layers, M, 0), strict=True)
    model.to(device)
    for p in model.teachers.parameters():                       # run.py:101-112
        p.requires_grad = False
    bm = model.student.news_encoder.bert_model
    for p in bm.parameters():
        p.requires_grad = False
    for i, layer in enumerate(bm.bert.encoder.layer):
        if i in trainable:
            for p in layer.parameters():
                p.requires_grad = True
    return model, args


def bytes_of(batch):
    n = 0
    for t in batch:
        for u in (t if isinstance(t, (list, tuple)) else [t]):
            n += u.numel() * u.element_size()
    return n


# ----------------------------------------------------------------------------- CPU legs (oracle)
def oracle_train_step_fn(layers, trainable, b_sample, seed=0):
    """Reference algorithm on CPU (oracle port): forward + backward + Adam(amsgrad) for b_sample
    impressions at the workload's shape.  Returns a callable running one step."""
    import tinyrec.synth as synth
    from oracle import model as om, optim as oopt
    sd = synth.kd_model_state(layers, M, 0)
    keys = om.trainable_keys(sd, trainable)
    for k in keys:
        sd[k].requires_grad_(True)
    news = synth.news_table(5000, L=L, seed=1234)
    tabs = synth.teacher_tables(5000, M, D, seed=1234)
    hist_idx, hmask, cand_idx, label = synth.train_impressions(b_sample, 5000, H, K, seed=seed)
    history = torch.from_numpy(news[hist_idx].astype(np.int64))
    candidate = torch.from_numpy(news[cand_idx].astype(np.int64))
    th = [torch.from_numpy(t[hist_idx]) for t in tabs]
    tc = [torch.from_numpy(t[cand_idx]) for t in tabs]
    state = {k: [torch.zeros_like(sd[k]) for _ in range(3)] for k in keys}
    step = [0]

    def run():
        for k in keys:
            sd[k].grad = None
        total = om.kd_model_forward(sd, history, torch.from_numpy(hmask), candidate, torch.from_numpy(label), th, tc,
                                    layers, False, 1.0, 0.2)[0]
        total.backward()
        step[0] += 1
        with torch.no_grad():
            for k in keys:
                m_, v_, vm_ = state[k]
                oopt.adam_amsgrad_step(sd[k], sd[k].grad, m_, v_, vm_, step[0], lr=1e-4)
        return float(total.detach())
    return run


def cpu_baseline(layers, trainable, budget_s=20.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b_sample = 4
    fn = oracle_train_step_fn(layers, trainable, b_sample)
    t0 = time.perf_counter()
    fn()
    t1 = time.perf_counter() - t0
    reps = max(1, min(3, int(budget_s / max(t1, 1e-3)) - 1))
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts)) if ts else t1
    return {"value": b_sample / t, "unit": "impressions/s", "cores": cores, "kind": "port",
            "sample": f"{1 + reps} train steps (fwd+bwd+Adam) of {b_sample} impressions at the workload shape, fp32 torch CPU oracle",
            "seconds_per_step": t}


def run_reference(a):
    """Reference arm: the CPU oracle port of the reference's train step (fwd + bwd + Adam(amsgrad), fp32 torch, all host
    threads) at the workload's shape and -- host time permitting -- its full per-GPU batch.  The port omits the
    reference's dead BertPooler / classifier compute and its per-forward one-hot rel-pos bias, so it is FASTER than the
    reference would be: the GPU / reference ratio the driver computes from it is a lower bound."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[a.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # the full batch of 32 impressions per step unless (steps + warmup) such steps would not finish in ~6 minutes
    probe = oracle_train_step_fn(wl["layers"], wl["trainable"], 2)
    t0 = time.perf_counter(); probe(); per_imp = (time.perf_counter() - t0) / 2.0
    budget = 360.0 / max(1, a.steps + a.warmup)
    b_sample = int(max(1, min(B, budget / max(per_imp, 1e-4))))
    fn = oracle_train_step_fn(wl["layers"], wl["trainable"], b_sample)
    for _ in range(a.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    dt = time.perf_counter() - t0
    val = b_sample * a.steps / dt
    line = {"impl": "reference", "metric": "kd_train_impressions_per_sec", "value": val, "unit": "impressions/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(a.workload, 1, b_sample),
            "cpu_baseline": {"value": val, "unit": "impressions/s", "cores": cores, "kind": "port",
                             "sample": f"{b_sample} impressions per step (of {B}) at the workload shape, eval-mode fwd + bwd + "
                                       "Adam; oracle port of the reference modules (the Python reference cannot travel to "
                                       "the GPU box); omits the reference's dead pooler/classifier and one-hot rel-pos "
                                       "work, i.e. a lower bound on the speed-up"},
            "e2e": {"value": val, "unit": "impressions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_dict(workload, n, batch=B):
    wl = WORKLOADS[workload]
    return {"workload": wl["name"], "student_layers": wl["layers"], "trainable_layers": wl["trainable"],
            "per_gpu_batch": batch, "global_batch": batch * n, "history": H, "npratio": K - 1, "title_tokens": L,
            "teachers": M, "news_dim": D, "vocab": 30522, "parallelism": f"dp{n}",
            "l2": "per-step activation working set (~5 GB) >> 126 MB L2; 8 rotating input batches",
            "dropout": "0.1 active (training mode, as the reference train(); counter-based masks regenerated in backward)"}


# ----------------------------------------------------------------------------- GPU arm
def run_tinyrec(a):
    import torch.distributed as dist
    import tinyrec.ops as ops
    import tinyrec.optim as topt
    import tinyrec.run as trun
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 and world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import tinyrec.parallel as par
        par.nccl_defaults()                                       # NCCL_MAX_CTAS: see tinyrec.optim.DistributedOptimizer
        dist.init_process_group("nccl", device_id=device)
    wl = WORKLOADS[a.workload]
    model, margs = make_model(wl["layers"], wl["trainable"], device)
    opt = topt.Adam(model, lr=1e-4)
    if world > 1:
        topt.broadcast_parameters(model, 0)
        opt = topt.DistributedOptimizer(opt)
    batcher, dev_batches, host_batches = make_inputs(rank, device)

    def step(idx_batch):
        """one eager step from an index batch: device row gathers (loader) + forward + backward + optimizer"""
        batch = assembled(batcher, idx_batch)
        opt.zero_grad()
        out = model(*batch)
        out[0].backward()
        opt.step()
        return out[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The train step as the public API (tinyrec.run.train) runs it: tinyrec.run.GraphedTrainStep captures the loader's
    # row gathers + zero_grad + forward + backward (+ bucketed NCCL all-reduce) + Adam once and replays it per batch;
    # only the index arrays are inputs.  --no-graph keeps the eager loop.
    gstep, graph_note = None, "eager (--no-graph)"
    if not a.no_graph:
        try:
            gstep = trun.GraphedTrainStep(model, opt, dev_batches[0], warmup=max(a.warmup, 3), batcher=batcher)
            graph_note = "CUDA graph replay of the whole step incl. the loader's row gathers (tinyrec.run.GraphedTrainStep)"
        except Exception as exc:                                  # noqa: BLE001  (reported, not hidden)
            gstep, graph_note = None, f"eager: graph capture failed ({type(exc).__name__}: {exc})"
            torch.cuda.synchronize()
    if world > 1:                                                 # all ranks must take the same path
        flag = torch.tensor([1 if gstep is not None else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and gstep is not None:
            gstep, graph_note = None, "eager: graph capture failed on another rank"

    def run_step(batch):
        if gstep is not None:
            return gstep(*batch)[0]
        return step(batch)

    def timed(n_steps):
        """n_steps device-timed steps over the rotating device-resident batches -> ms (max over ranks)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(n_steps):
            run_step(dev_batches[i % N_BATCHES])
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(a.warmup, 3)):
        run_step(dev_batches[i % N_BATCHES])
    barrier()
    # ---- device-timed region (inputs resident in HBM)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.stats.launches
    ms = timed(a.steps)
    launches = (gstep.launches_per_step * a.steps) if gstep is not None else (ops.stats.launches - launches0)
    clocks = sampler.stop() if rank == 0 else None
    value = B * world * a.steps / (ms / 1e3)
    # ---- the same loop for >= 2.5 s: the step runs at the board power cap, so a 0.15 s region (20 steps) is measured
    # before the clocks settle; this is the number to quote for long runs
    n_sus = max(a.steps, int(2500.0 / max(ms / a.steps, 1e-3)) + 1)
    sampler2 = ClockSampler(local)
    if rank == 0:
        sampler2.start()
    ms_sus = timed(n_sus)
    clocks_sus = sampler2.stop() if rank == 0 else None
    sustained = {"value": B * world * n_sus / (ms_sus / 1e3), "unit": "impressions/s", "steps": n_sus,
                 "seconds": ms_sus / 1e3, "ms_per_step": ms_sus / n_sus, "clocks": clocks_sus}

    # ---- e2e: pinned host INDEX arrays -> H2D every step -> public API (row gathers + step) -> loss read back
    def e2e_step(hb):
        if gstep is not None:                                     # H2D straight into the graph's static inputs
            return float(gstep(*hb)[0].item())
        return float(step(tuple(t.to(device, non_blocking=True) for t in hb)).item())

    def wall(fn, batches):
        for i in range(3):
            fn(batches[i % N_BATCHES])
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            fn(batches[i % N_BATCHES])
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return B * world * a.steps / float(t.item())
    e2e_val = wall(e2e_step, host_batches)

    # ---- the round-1 form of the same leg for comparison: the six ASSEMBLED model inputs (8 MB) shipped from pinned
    # host memory every step, no loader work on the device (single GPU only: it needs a second captured graph)
    e2e_pre = None
    if world == 1 and gstep is not None and not a.no_preassembled:
        pre_dev = assembled(batcher, dev_batches[0], clone=True)
        pre_host = []
        for i in range(N_BATCHES):
            bt = assembled(batcher, dev_batches[i], clone=True)
            pin = lambda t: t.cpu().pin_memory()  # noqa: E731
            pre_host.append((pin(bt[0]), pin(bt[1]), pin(bt[2]), pin(bt[3]), [pin(t) for t in bt[4]], [pin(t) for t in bt[5]]))
        g2 = trun.GraphedTrainStep(model, opt, pre_dev, warmup=3)
        e2e_pre = {"value": wall(lambda hb: float(g2(*hb)[0].item()), pre_host), "unit": "impressions/s",
                   "h2d_bytes_per_step": bytes_of(pre_host[0]), "d2h_bytes_per_step": 4,
                   "what": "six pre-assembled Model.forward inputs per step from pinned host memory (round-1 e2e)"}
        g2.close()

    # ---- roofline leg: the same steps launched eagerly with CUDA events around every tnr_gemm_bf16 launch (events
    # cannot be recorded inside a graph replay, so the per-kernel durations come from this pass, run right after the
    # timed region in the same process and clock state; without --no-graph it is a separate pass, with it the same)
    n_roof = min(a.steps, 20)
    barrier()
    ops.stats.gemm_events = []
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for i in range(n_roof):
        step(dev_batches[i % N_BATCHES])
    r1.record()
    barrier()
    eager_ms = r0.elapsed_time(r1) / n_roof
    gemm_events, ops.stats.gemm_events = ops.stats.gemm_events, None

    # ---- N > 1: CUDA-event timeline of the gradient exchange over eagerly launched steps (where each bucket's
    # all-reduce starts and ends relative to the end of the backward; "exposed" = what the optimizer waits for)
    comm = None
    if world > 1 and hasattr(opt, "trace"):
        opt.trace = []
        for i in range(12):
            step(dev_batches[i % N_BATCHES])
        barrier()
        tr, opt.trace = opt.trace[2:], None
        med = lambda xs: float(np.median(xs)) if len(xs) else None  # noqa: E731
        nb = min(len(t.get("buckets", [])) for t in tr)
        buckets = []
        for j in range(nb):
            b0 = tr[0]["buckets"][j]
            buckets.append({"mbytes": (b0["hi"] - b0["lo"]) * 4 / 1e6,
                            "ready_ms_before_backward_end": med([t["buckets"][j]["ready"].elapsed_time(t["backward_end"]) for t in tr]),
                            "start_ms_before_backward_end": med([t["buckets"][j]["start"].elapsed_time(t["backward_end"]) for t in tr]),
                            "duration_ms": med([t["buckets"][j]["start"].elapsed_time(t["buckets"][j]["done"]) for t in tr]),
                            "done_ms_after_backward_end": med([t["backward_end"].elapsed_time(t["buckets"][j]["done"]) for t in tr])})
        comm = {"steps": len(tr), "launch": "eager", "step_ms": med([t["t0"].elapsed_time(t["step_end"]) for t in tr]),
                "backward_end_to_step_end_ms": med([t["backward_end"].elapsed_time(t["step_end"]) for t in tr]),
                "exposed_ms": med([max(0.0, max(t["backward_end"].elapsed_time(b["done"]) for b in t["buckets"])) for t in tr]),
                "sm_reserve": getattr(opt, "sm_reserve", 0), "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS"),
                "collective": (("tnr_allreduce_p2p, NVLS multimem.ld_reduce / multimem.st (one kernel per bucket)"
                                if getattr(model.train_state().flat.symm, "mc_ptr", 0) else
                                "tnr_allreduce_p2p (one kernel per bucket over NVLink peer memory)")
                               if model.train_state().flat.symm is not None and getattr(opt, "use_p2p", False)
                               else f"nccl all_reduce ({model.train_state().flat.symm_error or 'TNR_P2P_ALLREDUCE=0'})"),
                "buckets": buckets}

    if rank == 0:
        peak_tf, peak_bw, how = peaks()
        gflop = sum(f for f, _, _ in gemm_events)
        gms = sum(s.elapsed_time(e) for _, s, e in gemm_events)
        ach = gflop / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
        tokens = B * (H + K) * L
        algo = flops_per_step(wl["layers"], len(wl["trainable"]), tokens)
        traffic, traffic_src = ncu_traffic() if a.workload == "kd4" else (None, None)
        line = {"metric": "kd_train_impressions_per_sec", "value": value, "unit": "impressions/s", "n_gpus": world,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_dict(a.workload, world), "launch": graph_note, "clocks": clocks,
                "sustained": sustained,
                "e2e": {"value": e2e_val, "unit": "impressions/s", "h2d_bytes_per_step": bytes_of(host_batches[0]),
                        "d2h_bytes_per_step": 4,
                        "what": "pinned host index arrays (hist_idx, hist_mask, cand_idx, label) in, loader row gathers + "
                                "train step on the device, loss read back"},
                "e2e_preassembled": e2e_pre,
                "comm_timeline": comm,
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "kernel": "tnr::gemm::gemm_kernel (tcgen05, all GEMM launches of the step)",
                             "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                             "peak_source": f"{how} bf16_tflops_sustained",
                             "traffic": traffic,
                             "traffic_note": f"DRAM bytes per GEMM launch, ncu --set full over one step ({traffic_src}); "
                                             "algorithmic operand+output bytes average 345 MB per launch (operands are "
                                             "read from HBM once: no re-reads)",
                             "measured": f"CUDA events around every GEMM launch over {n_roof} eagerly launched steps right "
                                         "after the timed region (per-kernel events cannot be recorded inside a graph replay)",
                             "gemm_launches_per_step": len(gemm_events) / n_roof,
                             "gemm_ms_per_step": gms / n_roof,
                             "gemm_share_of_step": (gms / n_roof) / (ms / a.steps) if ms > 0 else None,
                             "eager_ms_per_step_with_events": eager_ms,
                             "algorithmic_tflop_per_step": algo / 1e12,
                             "step_frac_of_tensor_roofline": (algo / 1e12) / (ms / a.steps * 1e-3) / peak_tf,
                             "sustained_step_frac_of_tensor_roofline": (algo / 1e12) / (ms_sus / n_sus * 1e-3) / peak_tf}}
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl["layers"], wl["trainable"])
        print(json.dumps(line), flush=True)
    if world > 1:
        # A captured graph holds NCCL kernels: tearing the communicator down while it is alive hung the ranks after
        # the JSON line was out (observed at N=2).  Drop the graph, drain, and leave without the collective teardown.
        gstep = None
        import gc
        gc.collect()
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ----------------------------------------------------------------------------- shared timing helpers
def _dist_setup(a):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 and world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import tinyrec.parallel as par
        par.nccl_defaults()
        dist.init_process_group("nccl", device_id=device)
    return rank, world, local, device


def _timed_steps(step, a, rank, world, local, device):
    """W warm-up steps, then K steps bracketed by barrier + synchronize, CUDA events, max over ranks.
    Returns (ms total, launches, gemm_events, op_events, clocks)."""
    import torch.distributed as dist
    import tinyrec.ops as ops

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for i in range(max(a.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.stats.gemm_events = []
    launches0 = ops.stats.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.stats.launches - launches0
    gemm_events, ops.stats.gemm_events = ops.stats.gemm_events, None
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), launches, gemm_events, clocks, barrier


def _wall_steps(step, a, world, device, barrier):
    import torch.distributed as dist
    for i in range(3):
        step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        step(i)
    barrier()
    t = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------- table build (configs[3])
def table_flops_per_news(layers, Lw):
    """per token per layer 14 155 776 (GEMMs) + 4 L E (attention); heads 307 200 / token
    + 393 216 / news."""
    return layers * (14155776 + 4 * Lw * 768) * Lw + 307200 * Lw + 393216


def run_table(a):
    import torch.distributed as dist
    import tinyrec.model_bert_2 as mb2
    import tinyrec.parallel as par
    import tinyrec.synth as synth
    wl = TABLE_WORKLOADS[a.workload]
    rank, world, local, device = _dist_setup(a)
    Lw, rows, layers = wl["L"], wl["rows"], wl["layers"]
    margs = synth.demo_args(num_hidden_layers=layers)
    model = mb2.ModelBert(margs)
    model.load_state_dict(synth.model_bert_state("", layers, 0), strict=True)
    model.to(device).eval()
    enc = model.news_encoder
    news = synth.news_table(N_NEWS, L=Lw, seed=1234, mean_len=wl["mean_len"], std_len=wl["std_len"], min_len=wl["min_len"])
    lo, hi = par.shard_rows(news.shape[0], rank, world)                 # contiguous row ranges per rank (run.py:438-447)
    nb = min(N_BATCHES, (hi - lo) // rows)
    shard = torch.from_numpy(news[lo:lo + nb * rows])
    dev_tab = shard.to(device)                                          # int32 rows resident in HBM
    host_tab = shard.pin_memory()
    out = torch.empty(rows, D, device=device, dtype=torch.float32)
    host_out = torch.empty(rows, D, dtype=torch.float32).pin_memory()

    @torch.no_grad()
    def step(i):
        b = i % nb
        x = dev_tab[b * rows:(b + 1) * rows].to(torch.int64)            # torch.LongTensor(arr), run.py:276-277
        out.copy_(enc(x))

    @torch.no_grad()
    def e2e_step(i):
        b = i % nb
        x = host_tab[b * rows:(b + 1) * rows].to(device, non_blocking=True).to(torch.int64)
        host_out.copy_(enc(x), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms, launches, gemm_events, clocks, barrier = _timed_steps(step, a, rank, world, local, device)
    value = rows * world * a.steps / (ms / 1e3)
    e2e_val = rows * world * a.steps / _wall_steps(e2e_step, a, world, device, barrier)
    # ---- the whole 161k-news table (BASELINE.json configs[3]) through the public driver tinyrec.run.build_news_table:
    # host int32 rows in, rows sharded over the ranks, all-gathered fp32 [N+1, D] table out; wall clock, max over ranks
    import tinyrec.run as trun
    barrier()
    t0 = time.perf_counter()
    full = trun.build_news_table(enc, news, batch_size=rows)
    barrier()
    tt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    full_s, full_rows = float(tt.item()), int(full.shape[0])
    del full
    if rank == 0:
        peak_tf, peak_bw, how = peaks()
        gflop = sum(f for f, _, _ in gemm_events)
        gms = sum(s_.elapsed_time(e_) for _, s_, e_ in gemm_events)
        ach = gflop / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
        algo = table_flops_per_news(layers, Lw) * rows
        line = {"metric": "news_embeddings_per_sec", "value": value, "unit": "news/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": wl["name"], "teacher_layers": layers, "tokens_per_news": Lw, "rows_per_step": rows,
                           "table_rows": N_NEWS + 1, "news_dim": D, "vocab": 30522, "parallelism": f"rows sharded x{world}",
                           "l2": f"{nb} rotating row batches; per-step activations >> 126 MB L2",
                           "note": "eval-mode forward (no dropout), random-init weights"},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "news/s", "h2d_bytes_per_step": rows * 2 * Lw * 4,
                        "d2h_bytes_per_step": rows * D * 4},
                "gpu_launches": launches,
                "full_table": {"api": "tinyrec.run.build_news_table (host int32 rows in, all-gathered fp32 table out)",
                               "rows": full_rows, "seconds": full_s, "news_per_sec": full_rows / full_s},
                "roofline": {"bound": "tensor", "kernel": "tnr::gemm::gemm_kernel (tcgen05, all GEMM launches of the step)",
                             "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                             "peak_source": f"{how} bf16_tflops_sustained", "traffic": None,
                             "gemm_ms_per_step": gms / a.steps, "gemm_share_of_step": gms / ms if ms > 0 else None,
                             "algorithmic_tflop_per_step": algo / 1e12,
                             "step_frac_of_tensor_roofline": (algo / 1e12) / (ms / a.steps * 1e-3) / peak_tf}}
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = table_cpu_baseline(layers, Lw, news)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def table_cpu_baseline(layers, Lw, news, budget_s=20.0):
    import tinyrec.synth as synth
    from oracle import model as om
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.model_bert_state("", layers, 0)
    n = 16 if Lw <= 32 else 4
    x = torch.from_numpy(news[1:1 + n].astype(np.int64))
    with torch.no_grad():
        t0 = time.perf_counter()
        om.news_encoder(sd, "news_encoder.", x, layers)
        t1 = time.perf_counter() - t0
        reps = max(1, min(4, int(budget_s / max(t1, 1e-3)) - 1))
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            om.news_encoder(sd, "news_encoder.", x, layers)
            ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    return {"value": n / t, "unit": "news/s", "cores": cores, "kind": "port",
            "sample": f"{1 + reps} forward passes of {n} news rows ({Lw} tokens, {layers} layers), fp32 torch CPU oracle"}


# ----------------------------------------------------------------------------- eval scoring (configs[4])
def run_eval(a):
    import torch.distributed as dist
    import tinyrec.dataloader as dl
    import tinyrec.model_bert as mb
    import tinyrec.ops as ops
    import tinyrec.synth as synth
    rank, world, local, device = _dist_setup(a)
    n_imp = EVAL_WORKLOAD["imps"]
    margs = synth.demo_args()
    ue = mb.UserEncoder(margs)
    ue.load_state_dict(synth.user_encoder_state("", D, margs.user_query_vector_dim, 0), strict=True)
    ue.to(device).eval()
    g = torch.Generator(device=device).manual_seed(4321)
    table = torch.randn(N_NEWS + 1, D, generator=g, device=device) * 0.1           # precomputed news vectors (165 MB)
    hist, hmask, ptr, cand, lab = synth.eval_impressions(n_imp * N_BATCHES, N_NEWS, H, seed=1234 + rank)
    batches, host_batches = [], []
    for b in range(N_BATCHES):
        s0, s1 = b * n_imp, (b + 1) * n_imp
        p0, p1 = int(ptr[s0]), int(ptr[s1])
        hb = (torch.from_numpy(hist[s0:s1]), torch.from_numpy(hmask[s0:s1]), torch.from_numpy(ptr[s0:s1 + 1] - p0),
              torch.from_numpy(cand[p0:p1]), torch.from_numpy(lab[p0:p1]))
        host_batches.append(tuple(t.pin_memory() for t in hb) + (int(np.diff(ptr[s0:s1 + 1]).max()),))
        batches.append(tuple(t.to(device) for t in hb) + (int(np.diff(ptr[s0:s1 + 1]).max()),))
    per = torch.zeros(n_imp, 5, device=device, dtype=torch.float64)
    sums = torch.zeros(5, device=device, dtype=torch.float64)

    @torch.no_grad()
    def score(bt):
        hi_t, hm_t, ptr_t, cand_t, lab_t, max_c = bt
        user = ue.forward_gather(table, hi_t, hm_t)      # news_scoring[log_ids] fused into the user-encoder kernel
        ops.eval_metrics(table, user, ptr_t, cand_t, lab_t, max_c, per, sums)

    def step(i):
        score(batches[i % N_BATCHES])

    def e2e_step(i):
        hb = host_batches[i % N_BATCHES]
        score(tuple(t.to(device, non_blocking=True) for t in hb[:5]) + (hb[5],))
        return sums.cpu()                                                       # the 5 running sums read back

    ms, launches, _, clocks, barrier = _timed_steps(step, a, rank, world, local, device)
    value = n_imp * world * a.steps / (ms / 1e3)
    e2e_val = n_imp * world * a.steps / _wall_steps(e2e_step, a, world, device, barrier)
    # per-kernel device times for the roofline of the dominant kernel
    ops.stats.op_events = []
    for i in range(N_BATCHES):
        step(i)
    torch.cuda.synchronize()
    ev, ops.stats.op_events = ops.stats.op_events, None
    per_op = {}
    for name, _, s_, e_ in ev:
        per_op[name] = per_op.get(name, 0.0) + s_.elapsed_time(e_) / N_BATCHES
    # ---- the whole dev-shaped set (BASELINE.json configs[4]: ~376k impressions) through the public driver
    # tinyrec.run.evaluate: host index arrays in, pipelined pinned H2D, metric means out; wall clock, max over ranks
    import tinyrec.run as trun
    n_full = 376471
    fh, fm, fptr, fcand, flab = synth.eval_impressions(n_full, N_NEWS, H, seed=99)
    trun.evaluate(ue, table, fh, fm, fptr, fcand, flab)    # warm-up at the timed sizes: staging slots, workspaces, allocator
    barrier()
    t0 = time.perf_counter()
    full_mean, full_total = trun.evaluate(ue, table, fh, fm, fptr, fcand, flab)
    barrier()
    tt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    full_s = float(tt.item())
    nnz = float(np.mean([bt[3].numel() for bt in batches]))
    algo_bytes = {"user_encoder_score": n_imp * H * (D + 1 + 1) * 4 + n_imp * D * 4,
                  "eval_metrics": nnz * (D * 4 + 5) + n_imp * D * 4}
    if rank == 0:
        _, peak_bw, how = peaks()
        top = max(per_op, key=per_op.get)
        ach = algo_bytes.get(top, 0.0) / (per_op[top] * 1e-3) / 1e9
        h2d = sum(t.numel() * t.element_size() for t in host_batches[0][:5])
        line = {"metric": "eval_impressions_per_sec", "value": value, "unit": "impressions/s", "n_gpus": world,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (scores), f64 (metrics)", "data": "synthetic",
                "config": {"workload": EVAL_WORKLOAD["name"], "impressions_per_step": n_imp, "history": H,
                           "mean_candidates": nnz / n_imp, "table_rows": N_NEWS + 1, "news_dim": D,
                           "parallelism": f"impressions sharded x{world}",
                           "l2": f"{N_BATCHES} rotating batches; 165 MB table, 210 MB of history rows gathered from it per step > 126 MB L2"},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "impressions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 40},
                "gpu_launches": launches,
                "full_dev_set": {"api": "tinyrec.run.evaluate (host index arrays in, metric means out)",
                                 "impressions": int(full_total), "seconds": full_s,
                                 "impressions_per_sec": full_total / full_s,
                                 "auc_mrr_ndcg5_ndcg10": [float(x) for x in full_mean]},
                "roofline": {"bound": "hbm", "kernel": "tnr_" + top, "achieved": ach, "peak": peak_bw, "unit": "GB/s",
                             "frac": ach / peak_bw, "peak_source": f"{how} hbm_gbs", "traffic": None,
                             "kernel_ms_per_step": per_op, "algorithmic_bytes_per_step": algo_bytes}}
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = eval_cpu_baseline(hist, hmask, ptr, cand, lab)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def eval_cpu_baseline(hist, hmask, ptr, cand, lab, n=1500):
    """The reference's per-impression Python loop (run.py:335-361) restated with the oracle: user vector,
    dot scores, AUC / MRR / nDCG@5/10."""
    import tinyrec.synth as synth
    from oracle import metrics as omet, model as om
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    table = (rng.standard_normal((N_NEWS + 1, D)) * 0.1).astype(np.float32)
    sd = synth.user_encoder_state("user_encoder.", D, 200, 0)
    t0 = time.perf_counter()
    with torch.no_grad():
        for s0 in range(0, n, 128):                                            # batches of 128 like the eval loader
            s1 = min(n, s0 + 128)
            vecs = torch.from_numpy(table[hist[s0:s1]])
            user = om.user_encoder(sd, "user_encoder.", vecs, torch.from_numpy(hmask[s0:s1]), False).numpy()
            for i in range(s0, s1):
                c = table[cand[ptr[i]:ptr[i + 1]]]
                omet.impression_metrics(lab[ptr[i]:ptr[i + 1]], c @ user[i - s0])
    t = time.perf_counter() - t0
    return {"value": n / t, "unit": "impressions/s", "cores": cores, "kind": "port",
            "sample": f"{n} impressions, per-impression Python loop as the reference's test()"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="tinyrec", choices=["tinyrec", "reference"])
    ap.add_argument("--workload", default="kd4", choices=sorted(WORKLOADS) + sorted(TABLE_WORKLOADS) + ["eval"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-preassembled", dest="no_preassembled", action="store_true",
                    help="kd workloads: skip the extra e2e leg that ships pre-assembled model inputs (round-1 form)")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true",
                    help="kd workloads: launch every step eagerly instead of replaying the captured CUDA graph")
    a = ap.parse_args()
    if a.impl == "reference":
        if a.workload not in WORKLOADS:
            raise SystemExit("--impl reference is defined for the train workloads (kd4 / kd2)")
        run_reference(a)
    elif a.workload in TABLE_WORKLOADS:
        run_table(a)
    elif a.workload == "eval":
        run_eval(a)
    else:
        run_tinyrec(a)


if __name__ == "__main__":
    main()
