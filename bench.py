#!/usr/bin/env python
"""Benchmark of the guided-DDIM + exemplar-retrieval hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Metric: guided DDIM clip-steps/s.  One "step" of this script = ONE guided batch of B=64 clips per GPU
(BASELINE.json configs[1]; per-GPU work fixed -> weak scaling, 8 GPUs = the 512 clips of configs[2]):
discourse retrieval against a 4096-entry synthetic annotated DB, batched 50-step DDIM inversion of the
E retrieved exemplars, 50 insertion-guided sampling steps  =  50 * (B + E) clip-steps.
  value : the device-resident loops (K6 state + inversion + guided sampling), CUDA-event timed;
  e2e   : the same batches from pinned HOST buffers through the public API -- H2D, retrieval, codec, loops,
          decode, D2H of the latents every step -- via GuidedPipeline (stage 1 of the next batch overlaps the
          loops of the current one); e2e.sync = one synchronous MotionDiffusion.forward(**host_batch) per
          step, the call tools/visualize.py:200 makes.
Also reported under "knn": the kNN sweep of configs[3] (queries/s at Q = 1, 8, 64, 4096; HBM roofline of the
exact scan, tensor roofline of the large-batch similarity kernel).
--impl reference times the reference's algorithm (oracle/ port: as-written op sequence, exemplars
inverted one by one at B=1) on the host cores.
"""
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

import torch  # noqa: E402

METRIC, UNIT = "guided_ddim_clip_steps_per_sec", "clip-steps/s"
B_PER_GPU, N_DB, STEPS = 64, 4096, 50
GFLOP_PER_CLIP_STEP = 3.348          # algorithmic, invariants hoisted
GEMM_GFLOP_PER_CLIP_STEP = 3.291     # dense-GEMM share of the above
GUIDANCE = [0] * 25 + list(range(25))   # decreasing_till_25 (tools/visualize.py:90-91)
WORKLOAD = ("configs[1]: 64 clips/GPU guided DDIM (discourse retrieval, inversion + insertion guidance "
            "decreasing_till_25, len150@15fps)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(int(r[0]) for r in rows), "sm_max_mhz": int(rows[0][1]),
                "reasons": reasons, "samples": len(rows)}


def dist_env():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    return rank, world, int(os.environ.get("LOCAL_RANK", 0))


# ------------------------------------------------------------------------------------------------
def make_batch(first_clip, n_clips):
    """Synthetic query clips (pinned host memory) with discourse annotations, collated like
    beatx_collate_fn; clip i depends only on its global index, so shards agree with the whole."""
    from rag_gesture_b200 import synthetic as S
    qs = S.SyntheticGestureDataset(first_clip + n_clips, seed=8)
    batch = S.collate([qs[first_clip + i] for i in range(n_clips)])
    for k, v in batch.items():
        if torch.is_tensor(v):
            batch[k] = v.pin_memory()
        elif isinstance(v, list) and v and torch.is_tensor(v[0]):
            batch[k] = [t.pin_memory() for t in v]
    batch["retrieval_method"] = "discourse"
    return batch


def infer_kwargs():
    return dict(use_inversion=True, outpaint=False, inversion_start_time=-1, insertion_guidance=True,
                guidance_iters=list(GUIDANCE), guidance_lr=0.1)


def h2d_bytes(batch):
    n = 0
    for v in batch.values():
        if torch.is_tensor(v):
            n += v.numel() * v.element_size()
        elif isinstance(v, list) and v and torch.is_tensor(v[0]):
            n += sum(t.numel() * t.element_size() for t in v)
    return n


def run_b200(args, emit):
    import torch.distributed as dist
    import rag_gesture_b200 as R
    from rag_gesture_b200 import config as C
    from rag_gesture_b200 import synthetic as S
    from rag_gesture_b200.engine import launch_count
    rank, world, local = dist_env()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)"
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, tc_burst, tc_sus, peak_src = peaks()

    from rag_gesture_b200 import _lib
    prec = {"bf16": _lib.PREC_BF16, "bf16x3": _lib.PREC_BF16X3, "fp32": _lib.PREC_FP32}[args.precision]
    if args.gemm_kernel or args.gemm2_min_rows or args.gemm2_persist_tiles or args.pair128_min_rows:
        _lib.check(_lib.load().rg_set_gemm_kernel(args.gemm_kernel, args.gemm2_min_rows, args.gemm2_persist_tiles, args.pair128_min_rows))
    cfg = C.model_cfg()
    cfg["use_retrieval_for_test"] = True
    cfg["model"]["precision"] = prec
    arch = R.build_architecture(cfg, database=S.SyntheticGestureDataset(N_DB, seed=7))
    arch.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
    arch = arch.to(dev).eval()
    B = B_PER_GPU
    batch = make_batch(rank * B, B)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    if args.ncu_steps:
        return ncu_pass(args, arch, batch, dev, emit)
    # ---- value: device-resident loops -------------------------------------------------------------
    # K prepared batches (inputs, exemplars and retrieval results resident in HBM) through
    # MotionDiffusion.run_many: pass k = guided loop of batch k-1 fused level by level with the inversion loop
    # of batch k's exemplars (one kernel chain per level for both); the un-fused head (inversion of batch 0)
    # and tail (guided loop of batch K-1) are inside the timed region.  "sequential" = run_prepared per batch.
    n_gb = max(2, args.steps)
    gbs = [arch.prepare(**dict(batch, inference_kwargs=infer_kwargs())) for _ in range(n_gb)]
    gb = gbs[0]
    E = len(gb.jobs)
    clip_steps = gb.clip_steps(STEPS)
    sampler = ClockSampler(local)
    out_holder = {}

    def hot():
        out_holder["x"] = arch.run_prepared(gb)
    t_seq = timed(hot, args.steps, args.warmup)
    for _ in range(max(1, args.warmup // 2)):
        arch.run_many(gbs[:2])
    barrier()
    flush.zero_()
    n0 = launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out_holder["x"] = arch.run_many(gbs[:args.steps])
    b.record()
    barrier()
    launches = (launch_count() - n0) // args.steps
    t = torch.tensor([a.elapsed_time(b) / 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_hot = float(t)
    total_steps = torch.tensor([clip_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_steps)
    value = float(total_steps) * args.steps / t_hot
    value_seq = float(total_steps) * args.steps / t_seq

    # ---- e2e: the public API with host buffers ---------------------------------------------------------
    # Every step: H2D of that step's pinned host batch, retrieval from scratch (caches cleared), codec,
    # loops, decode, D2H of the latents into pinned memory.  Headline = GuidedPipeline (the drop-in for
    # `for data in loader: model(**data)`, tools/visualize.py:189-200: stage 1 of batch i+1 overlaps the
    # loops of batch i; the first batch's stage 1 is NOT overlapped and is inside the timed region);
    # "sync" = one synchronous MotionDiffusion.forward(**batch) per step.
    from rag_gesture_b200.architecture import GuidedPipeline
    host_out = torch.empty(B, C.N_TOKENS, C.LATENT_DIM).pin_memory()
    db = arch.model.database

    def fresh_batch():
        for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):      # no retrieval-cache hits
            d.clear()
        return dict(batch, inference_kwargs=infer_kwargs())

    def e2e_sync():
        res = arch(**fresh_batch())
        host_out.copy_(res["prev_latentout"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t_sync = timed(e2e_sync, args.steps, max(1, args.warmup // 2))

    pipe = GuidedPipeline(arch)

    host_bufs = [host_out, torch.empty_like(host_out).pin_memory()]

    def e2e_pipelined(n):
        # the consumer copies batch k's latents to pinned host memory asynchronously and waits for that copy
        # one batch later, so the device never drains between passes
        pending = None
        for k, res in enumerate(pipe.run(fresh_batch() for _ in range(n))):
            host_bufs[k % 2].copy_(res["prev_latentout"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending.synchronize()
            pending = ev
        if pending is not None:
            pending.synchronize()
    e2e_pipelined(max(2, args.warmup // 2))
    barrier()
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e2e_pipelined(args.steps)
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b) / 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t)
    clocks = sampler.stop()
    e2e_value = float(total_steps) * args.steps / t_e2e
    e2e_sync_value = float(total_steps) * args.steps / t_sync

    # ---- e2e with the TransformerVAE codec (4 body-part VAEs, synthetic YAML + checkpoints) in place of the stand-in ----
    e2e_vae = None
    if rank == 0 and world == 1:
        import tempfile
        with tempfile.TemporaryDirectory() as root:
            vae_kw = dict(num_heads=4, ff_size=1024, num_layers=4)
            cfg_v = C.model_cfg()
            cfg_v["use_retrieval_for_test"] = True
            cfg_v["model"]["precision"] = prec
            cfg_v["model"]["vae_cfg"] = S.write_vae_files(root, latent_dim=C.LATENT_DIM, **vae_kw)
            arch_v = R.build_architecture(cfg_v, database=arch.model.database.dataset)
        arch_v.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
        arch_v = arch_v.to(dev).eval()
        pipe_v, dbv = GuidedPipeline(arch_v), arch_v.model.database

        def run_v(n):
            def gen():
                for _ in range(n):
                    for d in (dbv.test_indexes, dbv.test_dbounds, dbv.test_qbounds):
                        d.clear()
                    yield dict(batch, inference_kwargs=infer_kwargs())
            for k, res in enumerate(pipe_v.run(gen())):
                host_bufs[k % 2].copy_(res["prev_latentout"], non_blocking=True)
            torch.cuda.synchronize()
        run_v(2)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run_v(args.steps)
        b.record()
        torch.cuda.synchronize()
        tv = a.elapsed_time(b) / 1e3
        e2e_vae = {"value": round(clip_steps * args.steps / tv, 2), "unit": UNIT, "ms_per_step": round(1e3 * tv / args.steps, 3),
                   "codec": "vae.GestureRepEncoder: 4 TransformerVAEs (latent 512, 4 heads, ff 1024, 4 layers -- sizes guessed, the "
                            "shipped YAMLs / checkpoints are not in the reference repo), encode of B clips + E exemplars and "
                            "decode inside the timed region; projections on the tcgen05 GEMM (bf16x3), attention rg_op_mha, "
                            "the rest PyTorch kernels",
                   "api": "GuidedPipeline(model).run(batches)"}
        del arch_v, pipe_v
        torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel family: the dense GEMMs of one denoiser evaluation --------------
    roof = gemm_roofline(arch, [B, E], dev, flush, tc_sus, peak_src, args.precision)
    if args.precision != "fp32":
        roof = gemm_chain_roofline(arch, [B, E], dev, flush, tc_sus, roof, args.steps)
    elementwise = elementwise_roofline(arch, dev, hbm_peak, peak_src)
    knn = knn_bench(args, dev, rank, world, hbm_peak, tc_sus, peak_src, flush, tc_burst)
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        knn["cpu_baseline"] = knn_cpu_baselines(n_full=args.knn_n, seconds=min(10.0, args.cpu_seconds))

    # ---- configs[4]: long-form synthesis, a 10-minute stream (67 windows of 150 frames, 15 overlapping) -------
    longform = longform_bench(args, arch, dev, rank, world, barrier) if args.longform_seconds > 0 else None

    # ---- configs[0]: one clip, plain 50-step DDIM, no retrieval (latency; CUDA-graph replay of the chain) ----
    def plain_b1():
        kw1 = arch.model.get_precompute_condition(device=dev, text=batch["word"][:1].to(dev), audio=batch["audio"][:1].to(dev),
                                                  speaker_ids=batch["speaker_ids"][:1].to(dev), re_dict=1)
        qm1 = torch.ones(1, C.N_TOKENS, device=dev)
        qm1[:, C.QUERY_MASK_ZERO_ROWS] = 0
        kw1 = dict(xf_out=kw1["xf_out"], re_dict=None, sample_idx=None, query_mask={c: qm1 for c in C.CONDS},
                   motion_mask=S.motion_mask(1).to(dev))
        arch.model._state_cache = (None, None)
        return arch.diffusion_test.ddim_sample_loop(arch.model, (1, C.N_TOKENS, C.LATENT_DIM), clip_denoised=False,
                                                    model_kwargs=kw1, eta=0)
    t_b1 = timed(plain_b1, 5, 3) / 5
    configs0 = {"workload": "configs[0]: 1 clip, plain 50-step DDIM, no retrieval (condition encode + K6 state + loop)",
                "seconds_per_clip": round(t_b1, 5), "clip_steps_per_sec": round(STEPS / t_b1, 1),
                "us_per_evaluation": round(1e6 * t_b1 / STEPS, 1)}

    # ---- the other precision tiers on the same device-resident batches (value only, 2 batches) ----------------
    tiers = {}
    if rank == 0 and world == 1:
        for tier in ("bf16x3", "fp32"):
            if tier == args.precision:
                continue
            cfg_t = C.model_cfg()
            cfg_t["use_retrieval_for_test"] = True
            cfg_t["model"]["precision"] = {"bf16": _lib.PREC_BF16, "bf16x3": _lib.PREC_BF16X3, "fp32": _lib.PREC_FP32}[tier]
            arch_t = R.build_architecture(cfg_t, database=arch.model.database.dataset)
            arch_t.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
            arch_t = arch_t.to(dev).eval()
            gbs_t = [arch_t.prepare(**dict(batch, inference_kwargs=infer_kwargs())) for _ in range(2)]
            arch_t.run_many(gbs_t)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            arch_t.run_many(gbs_t)
            b.record()
            torch.cuda.synchronize()
            tt = a.elapsed_time(b) / 1e3
            tiers[tier] = {"value": round(2 * gbs_t[0].clip_steps(STEPS) / tt, 1), "unit": UNIT, "ms_per_step": round(1e3 * tt / 2, 2),
                           "parity": "rel-L2 <= 1e-3 vs reference (tests)", "batches": 2}
            del arch_t, gbs_t
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * t_hot / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "bf16x3", "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "clips_per_gpu": B, "exemplars_rank0": E, "clip_steps_per_step_rank0": clip_steps,
                       "db_entries": N_DB, "ddim_steps": STEPS,
                       "precision": {"bf16": "tcgen05 bf16 operands, fp32 TMEM accumulate (parity tier rel-L2 <= 2e-2)",
                                     "bf16x3": "tcgen05 hi+lo bf16 split, 3 products (parity tier rel-L2 <= 1e-3)",
                                     "fp32": "fp32 FMA GEMMs (exact tier)"}[args.precision],
                       "l2": "256 MiB flush before the timed region; each pass streams bf16 weights (77 MB) + K6 state "
                             "(250 MB) + activations, i.e. more than the 126 MB L2",
                       "schedule": "run_many: guided loop of batch k-1 fused with the inversion loop of batch k "
                                   "(rg_denoise_groups); head and tail passes un-fused, inside the timed region",
                       "sequential": {"value": round(value_seq, 2), "ms_per_step": round(1e3 * t_seq / args.steps, 3),
                                      "api": "run_prepared per batch: inversion loop, then guided loop"},
                       "parallelism": f"clips sharded over {world} GPU(s), no collective in the loop"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(batch),
                    "d2h_bytes_per_step": host_out.numel() * 4, "ms_per_step": round(1e3 * t_e2e / args.steps, 3),
                    "api": "GuidedPipeline(model).run(batches): K batches back to back, pipeline fill included",
                    "sync": {"value": round(e2e_sync_value, 2), "ms_per_step": round(1e3 * t_sync / args.steps, 3),
                             "api": "MotionDiffusion.forward(**host_batch), one synchronous call per step"},
                    "codec": "SyntheticGestureCodec stand-in (on both arms); `vae_codec` = the same run with the TransformerVAE codec",
                    "vae_codec": e2e_vae},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            # the reference's own CPU code next to the GPU number: at N = 1 only (the other ranks would sit in a collective
            # while rank 0 computes for tens of seconds; the driver's reference arm covers every N)
            "cpu_baseline": cpu_baseline(sample_seconds=args.cpu_seconds) if world == 1 and args.cpu_seconds > 0 else None,
            "knn": knn,
            "configs0": configs0, "tiers": tiers, "longform": longform, "elementwise": elementwise,
            "gflop_per_clip_step": GFLOP_PER_CLIP_STEP,
            "achieved_tflops_loop": round(value * GFLOP_PER_CLIP_STEP / 1e3 / world, 2),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def elementwise_roofline(arch, dev, hbm_peak, peak_src):
    """K8 / K9: the DDIM update (3 fp32 tensors: x, x0 in, x' out) and the insertion blend (per row: in_seq in for the
    row mask, then noise OR x in, x out = 3 tensors' worth of bytes) timed alone on inputs larger than L2 (2048
    clips = 180 MB per tensor), CUDA events, algorithmic bytes / time against the measured HBM copy bandwidth."""
    from rag_gesture_b200 import config as C
    eng = arch.model.rg_engine(arch.diffusion_test)
    n_clips = 2048
    shape = (n_clips, C.N_TOKENS, C.LATENT_DIM)
    x, x0, noise = (torch.randn(shape, device=dev) for _ in range(3))
    in_seq = torch.zeros(shape, device=dev)
    in_seq[:, 2:9] = 1.0
    out = torch.empty(shape, device=dev)
    bytes_t = x.numel() * 4

    def ev(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 1e3)
        return statistics.median(ts)
    t_up = ev(lambda: eng.ddim_update(x, x0, 10, -1, out=out))
    t_bl = ev(lambda: eng.blend_in_seq(x, in_seq, noise, 10, out=out))
    res = {}
    for name, t, n_t in (("ddim_update_kernel", t_up, 3), ("blend_kernel", t_bl, 3)):
        gbs = n_t * bytes_t / t / 1e9
        res[name] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 4),
                     "launch_ms": round(t * 1e3, 3), "bytes_per_launch": n_t * bytes_t, "tensors": n_t, "peak_source": peak_src}
    return res


def longform_bench(args, arch, dev, rank, world, barrier):
    """configs[4]: one synthetic audio stream of --longform-seconds (default 600 s = 9000 frames = 67 windows) through
    LongformSynthesizer.  (a) ONE stream on all GPUs: retrieval + DDIM inversions of the windows sharded over the
    ranks, one all-gather of the window payloads, the B=1 sampling chain (prev-latent dependency) on rank 0
    (SURVEY 8e row 4); (b) replicas: every rank its own stream (N streams at once).  Device-synchronised wall
    clock, max over ranks."""
    import torch.distributed as dist
    from rag_gesture_b200 import longform as LF
    from rag_gesture_b200 import synthetic as S
    n_frames = int(args.longform_seconds * 15)
    qs = S.SyntheticGestureDataset(96, seed=9)
    db = arch.model.database

    def window_fn(c, f0, f1):
        b = S.collate([qs[c % len(qs)]])
        b["retrieval_method"] = "discourse"
        return b

    def clear():
        for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
            d.clear()
    lf = LF.LongformSynthesizer(arch)
    n_win = len(LF.chunk_starts(n_frames))

    def timed_run(fn):
        clear()
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    warm = 150 + 135 * max(2 * world - 1, 3)                    # a short stream: graph capture, corpus upload
    lf.run_sharded(warm, window_fn, infer_kwargs())
    phases = {}
    t_stream = timed_run(lambda: lf.run_sharded(n_frames, window_fn, infer_kwargs(), timings=phases))
    out = {"workload": f"configs[4]: {args.longform_seconds:.0f} s stream @ 15 fps = {n_frames} frames, {n_win} windows of 150 "
                       "frames (15 overlap), discourse retrieval + inversion + insertion guidance per window, prev-latent chain",
           "windows": n_win, "one_stream": {"seconds": round(t_stream, 3), "windows_per_sec": round(n_win / t_stream, 2),
                                            "x_realtime": round(args.longform_seconds / t_stream, 1),
                                            "phases_rank0_s": {k: round(v, 3) for k, v in phases.items()},
                                            "partition": f"windows' retrieval + inversions sharded over {world} GPU(s), chain on rank 0"}}
    # several streams per GPU: their chains advance together, one batch of S clips per window index
    n_streams = 8

    def stream_fn(si):
        def fn(c, f0, f1):
            b = S.collate([qs[(c + 11 * si) % len(qs)]])
            b["retrieval_method"] = "discourse"
            return b
        return fn
    fns = [stream_fn(si) for si in range(n_streams)]
    lf.run_streams(150 + 135 * 2, fns, infer_kwargs())
    t_multi = timed_run(lambda: lf.run_streams(n_frames, fns, infer_kwargs()))
    out["streams_per_gpu"] = {"streams": n_streams * world, "per_gpu": n_streams, "seconds": round(t_multi, 3),
                              "windows_per_sec": round(world * n_streams * n_win / t_multi, 2),
                              "x_realtime": round(world * n_streams * args.longform_seconds / t_multi, 1),
                              "how": "LongformSynthesizer.run_streams: the prev-latent chains of 8 independent streams per GPU "
                                     "as one batch of 8 clips per window (an evaluation at B = 1 is latency-bound)"}
    if world > 1:
        t_rep = timed_run(lambda: lf.run(n_frames, window_fn, infer_kwargs(), batch_inversions=True))
        out["replicas"] = {"streams": world, "seconds": round(t_rep, 3), "windows_per_sec": round(world * n_win / t_rep, 2),
                           "x_realtime": round(world * args.longform_seconds / t_rep, 1)}
    return out


def ncu_pass(args, arch, batch, dev, emit):
    """`bench.py --ncu-steps S`: the SAME guided batch with both loops cut to their first S DDIM levels, one
    exact kNN pass and one tensor-core kNN call, nothing timed -- the command profiled under ncu (which
    serialises every launch it intercepts: the full 10k-launch step takes tens of minutes there)."""
    from rag_gesture_b200.parallel import KnnIndex, knn_topk
    S = args.ncu_steps
    diff = arch.diffusion_test
    ik = dict(infer_kwargs(), guidance_iters=list(GUIDANCE)[:S])
    gbs = [arch.prepare(**dict(batch, inference_kwargs=dict(ik, guidance_iters=list(ik["guidance_iters"]))))
           for _ in range(2)]
    gb = gbs[0]
    arch.model.rg_engine(diff)                   # 50-level schedule / timestep table (K7) uploaded first
    diff.num_timesteps = S                       # then the loop range only is cut
    arch.run_many(gbs)                           # inversion pass, fused guided+inversion pass, guided pass
    g = torch.Generator(device=dev).manual_seed(42)
    db = torch.nn.functional.normalize(torch.randn(args.knn_n, 768, device=dev, generator=g), dim=1)
    index = KnnIndex(db)
    for Q in (8, 4096):
        q = torch.nn.functional.normalize(torch.randn(Q, 768, device=dev, generator=g), dim=1)
        knn_topk(db, q, 8, index=index)
    torch.cuda.synchronize()
    emit({"ncu_pass": True, "ddim_levels": S, "clips": B_PER_GPU, "exemplars": len(gb.jobs), "knn_n": args.knn_n})


def gemm_roofline(arch, n_clips, dev, flush, tc_peak, peak_src, precision):
    """Every dense contraction of one denoiser evaluation, timed per shape (same kernel, same shapes
    as inside rg_denoise) with CUDA events on the launching stream, L2 flushed between launches."""
    from rag_gesture_b200 import _lib, ops
    if precision != "fp32":
        return gemm_roofline_tc(n_clips, dev, flush, tc_peak, peak_src, precision == "bf16x3")
    M = sum(n_clips) * 43
    shapes = [("qkv", 1536, 512, 8), ("sa_proj", 512, 512, 8), ("ca_q", 1536, 512, 8), ("ca_proj", 512, 512, 24),
              ("ca_mix", 512, 1536, 8), ("ffn1", 1024, 512, 8), ("ffn2", 512, 1024, 8), ("ffn_proj", 512, 512, 8),
              ("embed/out", 512, 512, 2)]
    tot_t, tot_f, per = 0.0, 0.0, {}
    for name, N, K, count in shapes:
        x, w, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.randn(N, device=dev)
        for _ in range(3):
            ops.linear(x, w, b)
        ts = []
        for _ in range(5):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.linear(x, w, b)
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) / 1e3)
        t = statistics.median(ts)
        per[name] = round(2.0 * M * N * K / t / 1e12, 2)
        tot_t += t * count
        tot_f += 2.0 * M * N * K * count
    ach = tot_f / tot_t / 1e12
    return {"bound": "tensor", "kernel": "gemm_tn_f32_kernel (fp32 FMA pipe; all GEMMs of one denoiser "
            "evaluation, time-weighted)", "achieved": round(ach, 2), "peak": tc_peak, "unit": "TFLOP/s",
            "frac": round(ach / tc_peak, 4), "traffic": None, "peak_source": f"{peak_src} bf16 sustained",
            "rows": M, "per_shape_tflops": per,
            "note": "exact fp32 tier runs on CUDA cores; the tcgen05 bf16 path is the next kernel"}


def gemm_chain_roofline(arch, n_clips, dev, flush, tc_peak, isolated, k_steps=2):
    """The dense contractions as they run INSIDE a step: rg_probe_gemm_only makes rg_denoise launch only its
    GEMMs -- the model's own weights (a different matrix per layer), the step's shapes, epilogues and
    programmatic-dependent-launch chain -- and 10 such evaluations are timed back to back with CUDA events on
    the launching stream (one L2 flush before them, as in the loop where weights stay L2-resident between
    levels).  achieved = algorithmic GEMM flops (3.291 GFLOP per clip-step) / that time, over the row counts the
    timed region of `value` launches, weighted by how often: per level one evaluation of E clips (head pass:
    inversion only), K-1 evaluations of B+E clips (fused guided + inversion passes: the persistent 2-CTA kernel)
    and one of B clips (tail pass).  The isolated, L2-flushed per-launch figures stay in `isolated`;
    `gemm_share` is GEMM-only time / full evaluation time."""
    from rag_gesture_b200 import _lib, synthetic as S
    from rag_gesture_b200 import config as C
    lib = _lib.load()
    eng = arch.model.rg_engine(arch.diffusion_test)

    def evals(fn, n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n / 1e3

    tot_t, tot_f, per, share = 0.0, 0.0, {}, {}
    B_, E_ = n_clips
    mix = [(E_, 1), (B_, 1), (B_ + E_, max(1, k_steps - 1))]          # the fused level last: reused below
    for clips, weight in mix:
        if clips <= 0:
            continue
        cond = S.synthetic_conditions(clips, seed=5)
        xf = eng.encode_conditions(cond["word"].to(dev), cond["audio"].to(dev), cond["speaker_ids"].to(dev))
        state = eng.precompute_state(xf)
        x, sm = S.synthetic_latents(clips, seed=6).to(dev), S.motion_mask(clips).to(dev)
        qm = torch.stack([S.query_masks(clips)[c] for c in C.CONDS], 0).to(dev).contiguous()
        out = torch.empty_like(x)
        step = lambda: eng.denoise(x, sm, qm, state, step_idx=10, out=out)
        for _ in range(3):
            step()
        t_full = evals(step, 10)
        _lib.check(lib.rg_probe_gemm_only(eng._h, 1))
        try:
            for _ in range(2):
                step()
            t_gemm = evals(step, 10)
        finally:
            _lib.check(lib.rg_probe_gemm_only(eng._h, 0))
        flops = GEMM_GFLOP_PER_CLIP_STEP * 1e9 * clips
        per[f"M{clips * 43}"] = round(flops / t_gemm / 1e12, 1)
        share[f"M{clips * 43}"] = round(t_gemm / t_full, 3)
        tot_t += t_gemm * weight
        tot_f += flops * weight
    ach = tot_f / tot_t / 1e12
    # Second roofline of the same chain: bytes that cross L2 <-> SM.  With 128x128 tiles every tile pulls its A and W
    # panels (bf16) out of L2 = 64 flop per operand byte, and writes fp32 / bf16 outputs (+ reads the fp32 residual).
    # Denominator: an L2-resident copy (24 MB buffers, read + write bytes) measured here, like the HBM figure.
    def chain_l2_bytes(M):
        tiles_m = (M + 127) // 128
        layer = [(1536, 512, 4, 0), (512, 512, 6, 4), (1536, 512, 4, 0), (512, 2048, 6, 0), (1024, 512, 2, 0),
                 (512, 1024, 4, 0), (512, 512, 6, 4)]             # (N, K, output bytes / element, residual bytes / element)
        tot = 0
        for N, K, ob, rb in layer * C.NUM_LAYERS + [(512, 512, 4, 0), (512, 512, 4, 0)]:
            tot += tiles_m * (N // 128) * (128 * K * 2 + 128 * K * 2) + M * N * (ob + rb)
        return tot
    import ctypes
    l2buf = torch.empty(48 << 20, dtype=torch.uint8, device=dev)
    gbs = ctypes.c_float()
    _lib.check(lib.rg_probe_l2_read(_lib.ptr(l2buf), l2buf.numel(), 40, ctypes.byref(gbs), _lib.stream_ptr()))
    l2_copy = float(gbs.value)
    l2_ach = chain_l2_bytes(clips * 43) / t_gemm / 1e9          # the fused level (measured last)
    l2 = {"bound": "l2", "achieved": round(l2_ach, 1), "peak": round(l2_copy, 1), "unit": "GB/s", "frac": round(l2_ach / l2_copy, 3),
          "rows": clips * 43, "peak_source": "measured here: rg_probe_l2_read, all SMs streaming a 48 MB L2-resident buffer with 128-bit loads",
          "bytes_per_evaluation": chain_l2_bytes(clips * 43),
          "note": "tile-level bytes (operand panels re-read per 128x128 tile = 64 flop/B, + outputs + residual) over the chain's "
                  "time, against the measured L2->SM read rate.  High, but kernels with 25 % / 50 % fewer operand bytes "
                  "(gemm2_tc_kernel_same_probe) are not faster: ~6 us of per-launch latency x 58 launches is the other "
                  "~40 % of the chain (DESIGN 6)"}
    # the same probe with the 2-CTA kernel forced (cta_group::2, 256x256 pair tiles, TMA-store epilogue), fused shape
    two_cta = {}
    try:
        for name, mode, pt in (("pair128_shared_weight_tile", 3, 0), ("pair256_persistent_2acc", 2, 1), ("pair256_one_tile_per_pair", 2, 10 ** 6)):
            _lib.check(lib.rg_set_gemm_kernel(mode, 0, pt, 0))
            for _ in range(3):
                step()
            _lib.check(lib.rg_probe_gemm_only(eng._h, 1))
            try:
                for _ in range(2):
                    step()
                t2 = evals(step, 10)
            finally:
                _lib.check(lib.rg_probe_gemm_only(eng._h, 0))
            two_cta[name] = {"rows": clips * 43, "tflops": round(GEMM_GFLOP_PER_CLIP_STEP * 1e9 * clips / t2 / 1e12, 1)}
    finally:
        _lib.check(lib.rg_set_gemm_kernel(0, 0, 296, 0))
    return {"bound": "tensor", "kernel": "gemm_tc_kernel<128,*> (tcgen05.mma cta_group::1 kind::f16, 128x128 tiles, TMA-fed, TMEM "
            "accumulator, two CTAs per SM): the 58 GEMM launches of one denoiser evaluation as their own PDL chain "
            "(rg_probe_gemm_only), replayed from the CUDA graph, 10 evaluations back to back",
            "mix": {f"M{c * 43}": w for c, w in mix}, "gemm2_tc_kernel_same_probe": two_cta, "l2_roofline": l2,
            "achieved": round(ach, 1), "peak": tc_peak, "unit": "TFLOP/s", "frac": round(ach / tc_peak, 4),
            "traffic": isolated.get("traffic"), "traffic_source": isolated.get("traffic_source"),
            "peak_source": isolated.get("peak_source"), "rows": isolated.get("rows"), "in_chain_tflops": per,
            "gemm_share": share, "executed_flop_multiplier": isolated.get("executed_flop_multiplier"),
            "isolated": {"achieved": isolated["achieved"], "frac": isolated["frac"],
                         "how": "every shape launched alone, L2 flushed before each launch, time-weighted",
                         "per_shape_tflops": isolated["per_shape_tflops"]}}


def gemm_roofline_tc(n_clips, dev, flush, tc_peak, peak_src, split):
    """tcgen05 GEMM launches alone, at the row counts the two loops of a bench step actually launch
    (M = 43 * clips for the guided loop, 43 * exemplars for the inversion loop; 50 evaluations each),
    operands pre-converted to bf16 planes as inside rg_denoise.  achieved = ALGORITHMIC flops
    (2*M*N*K per launch; bf16x3 executes 3x that on the pipe) / CUDA-event time, time-weighted."""
    import ctypes
    from rag_gesture_b200 import _lib
    lib = _lib.load()
    shapes = [("qkv", 1536, 512, 8), ("sa_proj", 512, 512, 8), ("ca_q", 1536, 512, 8), ("ca_proj", 512, 512, 24),
              ("ca_mix", 512, 1536, 8), ("ffn1", 1024, 512, 8), ("ffn2", 512, 1024, 8), ("ffn_proj", 512, 512, 8),
              ("embed/out", 512, 512, 2)]
    tot_t, tot_f, per = 0.0, 0.0, {}
    for clips in n_clips:
        M = clips * 43
        if M <= 0:
            continue
        for name, N, K, count in shapes:
            x, w, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.randn(N, device=dev)
            out = torch.empty(M, N, device=dev)
            ts = ctypes.c_float()
            _lib.check(lib.rg_probe_gemm_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), M, N, K, int(split),
                                            5, _lib.ptr(flush), flush.numel(), ctypes.byref(ts), _lib.stream_ptr()))
            t = ts.value / 1e3
            per[f"{name}@M{M}"] = round(2.0 * M * N * K / t / 1e12, 1)
            tot_t += t * count
            tot_f += 2.0 * M * N * K * count
    ach = tot_f / tot_t / 1e12
    return {"bound": "tensor", "kernel": "gemm_tc_kernel<128,*> (tcgen05.mma kind::f16, TMA-fed, TMEM accumulator; all "
            "GEMM launches of one inversion + one sampling evaluation, time-weighted, L2 flushed before each)",
            "achieved": round(ach, 1), "peak": tc_peak, "unit": "TFLOP/s", "frac": round(ach / tc_peak, 4),
            # dram__bytes_read+write of ONE launch of the costliest shape (qkv / ca_q, N=1536, M=4128) from the
            # committed ncu --set full capture (profiles/ncu_full_gemm_tc_r02.txt); its algorithmic DRAM bytes are
            # A 4.2 MB + W 1.6 MB (the fp32 output stays in L2 for the next kernel): no re-reads
            "traffic": 5832704, "traffic_source": "profiles/ncu_full_gemm_tc_r02.txt, gemm_tc_kernel<128,3,0> grid (12,33)",
            "peak_source": f"{peak_src} bf16 sustained", "rows": [c * 43 for c in n_clips],
            "per_shape_tflops": per, "executed_flop_multiplier": 3 if split else 1}


def knn_bench(args, dev, rank, world, hbm_peak, tc_peak, peak_src, flush, tc_burst=None):
    """configs[3]: 1M x 768 fp32 embeddings row-sharded over the ranks, top-8, sweep Q in {1, 8, 64, 4096}
    through sharded_knn (local top-k, all-gather, merge).  Q <= 8: the exact scan, one pass over the shard
    (HBM-bound); Q > 8: the tensor-core path (bf16 similarity GEMM, certified over-selection, exact fp32
    re-score; bit-identical results).  Two rooflines: the scan kernel alone against HBM, knn_tc_kernel
    alone against the bf16 tensor peak."""
    import ctypes
    import torch.distributed as dist
    from rag_gesture_b200 import _lib
    from rag_gesture_b200.parallel import KnnIndex, shard_range, sharded_knn
    n_total, dim, k = args.knn_n, 768, 8
    lo, hi = shard_range(n_total, rank, world)
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    db = torch.nn.functional.normalize(torch.randn(hi - lo, dim, device=dev, generator=g), dim=1)
    index = KnnIndex(db)
    out = {}
    uncert = 0
    for Q in (8, 1, 64, 4096):
        q = torch.nn.functional.normalize(torch.randn(Q, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(43)), dim=1)
        for _ in range(2):
            sharded_knn(db, q, k, n_total, index=index)
        ts = []
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            sharded_knn(db, q, k, n_total, index=index)
            e.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(e) / 1e3], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t))
        t = statistics.median(ts)
        row = {"queries_per_sec": round(Q / t, 1), "ms": round(t * 1e3, 3)}
        if Q <= KnnIndex.min_queries:
            gbs = (hi - lo) * dim * 4 / t / 1e9               # the fp32 shard is streamed once
            row.update(path="exact scan", hbm_gbs=round(gbs, 1), frac_hbm=round(gbs / hbm_peak, 4))
        else:
            tf = 2.0 * Q * (hi - lo) * dim / t / 1e12
            row.update(path="tcgen05 + certified re-score", tflops=round(tf, 1), frac_tensor=round(tf / tc_peak, 4),
                       uncertified=index.last_uncertified)
            uncert += index.last_uncertified
        out[f"q{Q}"] = row
    # the sharded result checked on the hardware it ran on: every rank all-gathers the shards (3 GB at 1M rows),
    # runs the UNSHARDED exact scan for a query sample and compares indices and scores bit for bit
    same = None
    if world > 1 and n_total % world == 0:
        from rag_gesture_b200.parallel import knn_topk
        full = torch.empty(n_total, dim, device=dev)
        dist.all_gather_into_tensor(full, db)
        ok = True
        for Q in (8, 64):
            qv = torch.nn.functional.normalize(torch.randn(Q, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(44)), dim=1)
            si, ss = sharded_knn(db, qv, k, n_total, index=index)
            ui, us = knn_topk(full, qv, k)
            ok = ok and torch.equal(si, ui) and torch.equal(ss, us)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
        del full
        torch.cuda.empty_cache()
    # the kernels alone (what the rooflines are about): CUDA events around the launch only
    lib = _lib.load()
    q8 = torch.nn.functional.normalize(torch.randn(8, dim, device=dev), dim=1)
    ms = ctypes.c_float()
    _lib.check(lib.rg_probe_knn_scan(_lib.ptr(db), hi - lo, dim, _lib.ptr(q8), 8, k, 5, _lib.ptr(flush),
                                     flush.numel(), ctypes.byref(ms), _lib.stream_ptr()))
    scan_gbs = (hi - lo) * dim * 4 / (ms.value / 1e3) / 1e9
    q4k = torch.nn.functional.normalize(torch.randn(4096, dim, device=dev), dim=1)
    ms_tc = ctypes.c_float()
    _lib.check(lib.rg_probe_knn_tc(index.handle, _lib.ptr(q4k), 4096, 5, _lib.ptr(flush), flush.numel(),
                                   ctypes.byref(ms_tc), _lib.stream_ptr()))
    tc_tf = 2.0 * 4096 * (hi - lo) * dim / (ms_tc.value / 1e3) / 1e12
    head = out["q4096"]
    index.close()
    return {"queries_per_sec": head["queries_per_sec"], "n": n_total, "dim": dim, "k": k, "q": 4096, "ms": head["ms"],
            "uncertified_queries": uncert, "sharded_equals_unsharded": same, "sweep": out,
            "roofline": {"bound": "hbm", "kernel": "knn_scan768_kernel<8> alone (exact fp32, 8 queries, 1 pass over the shard)",
                         "launch_ms": round(ms.value, 3), "achieved": round(scan_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(scan_gbs / hbm_peak, 4), "peak_source": peak_src,
                         "bytes_per_launch": (hi - lo) * dim * 4},
            "roofline_tc": {"bound": "tensor", "kernel": "knn_tc_kernel alone (tcgen05 bf16 similarity of 4096 queries x shard, "
                            "top-16 per (query, chunk) selected in the epilogue; scores never leave the SM)",
                            # a kernel timed alone: the BURST cuBLAS figure is the denominator
                            "launch_ms": round(ms_tc.value, 3), "achieved": round(tc_tf, 1), "peak": tc_burst or tc_peak,
                            "unit": "TFLOP/s", "frac": round(tc_tf / (tc_burst or tc_peak), 4),
                            "peak_source": f"{peak_src} bf16 burst (kernel timed in isolation)",
                            "frac_of_sustained": round(tc_tf / tc_peak, 4),
                            # ncu --set full at N = 1M, Q = 4096 (profiles/ncu_full_knn_tc_r01.txt): the bf16 shard
                            # (1.536 GB) is read from DRAM once, candidates written once
                            "traffic": 1544255000 + 7865600 if (hi - lo) == 1_000_000 else None,
                            "flop_per_launch": 2.0 * 4096 * (hi - lo) * dim}}


# ---- the reference's algorithm on the host cores ------------------------------------------------------
def cpu_reference_sample(n_clips, n_exemplars, n_steps, threads):
    """Bounded sample of the guided batch, executed like the reference: exemplars inverted one by one
    at B=1 (diffusion_architecture.py:323-354), then guided sampling of the clips, `n_steps` levels
    each.  Returns (clip_steps, seconds)."""
    from oracle import denoiser as OD
    from oracle import diffusion as ODF
    from rag_gesture_b200 import config as C
    from rag_gesture_b200 import synthetic as S
    torch.set_num_threads(threads)
    sd = cpu_reference_sample.sd = getattr(cpu_reference_sample, "sd", None) or S.synthetic_state_dict(0)
    model, diff = OD.OracleDenoiser(sd), ODF.OracleDiffusion()
    T, D = C.N_TOKENS, C.LATENT_DIM
    cond = S.synthetic_conditions(n_clips + n_exemplars, seed=5)
    t0 = time.perf_counter()
    with torch.no_grad():
        inv_rows = []
        for e in range(n_exemplars):
            c1 = {k: v[n_clips + e:n_clips + e + 1] for k, v in cond.items()}
            kw = dict(xf_out=OD.encode_conditions(sd, c1["word"], c1["audio"], c1["speaker_ids"]),
                      query_mask=S.query_masks(1), motion_mask=S.motion_mask(1))
            img = S.synthetic_latents(1, seed=6 + e, scale=0.5)
            for i in range(n_steps):
                img = diff.ddim_reverse_sample(model, img, torch.tensor([i]), kw)
            inv_rows.append(img)
        cb = {k: v[:n_clips] for k, v in cond.items()}
        kw = dict(xf_out=OD.encode_conditions(sd, cb["word"], cb["audio"], cb["speaker_ids"]),
                  query_mask=S.query_masks(n_clips), motion_mask=S.motion_mask(n_clips))
        in_seq = torch.zeros(n_clips, T, D)
        for e in range(min(n_exemplars, n_clips)):
            in_seq[e, 2:5] = inv_rows[e][0, 2:5]
        img = torch.randn(n_clips, T, D)
        for i in reversed(range(STEPS - n_steps, STEPS)):
            img, _ = diff.ddim_sample(model, img, torch.tensor([i] * n_clips), kw, in_seq)
    return (n_clips + n_exemplars) * n_steps, time.perf_counter() - t0


class _ShapeOnlyCodec(torch.nn.Module):
    """Stand-in for the reference's GestureRepEncoder (needs VAE yaml + checkpoints that are not in the repo);
    the denoiser and the sampling loops never call it (tests/golden/make_golden.py uses the same one)."""

    def __init__(self, vae_cfg, body_part_cat_axis="time"):
        super().__init__()
        self.vae_latent_dim = vae_cfg["latent_dim"]
        self.body_part_cat_axis = body_part_cat_axis


_REF = {}


def reference_modules():
    """The UNMODIFIED reference (oracle/refshim.py: /root/reference here, oracle/_ref on the GPU box), its
    ReGestureTransformer with the synthetic weights and its SpacedDiffusion; None when it is not available."""
    if "ns" in _REF:
        return _REF["ns"]
    _REF["ns"] = None
    try:
        from oracle import refshim
        from rag_gesture_b200 import config as C
        from rag_gesture_b200 import synthetic as S
        if not refshim.available():
            return None
        ns = refshim.load()
        ns.dt.GestureRepEncoder = _ShapeOnlyCodec
        cfg = C.denoiser_cfg()
        cfg.pop("type")
        model = ns.rg.ReGestureTransformer(**cfg, database=None, use_retrieval_for_test=False)
        missing, unexpected = model.load_state_dict(S.synthetic_state_dict(0), strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        ns.model, ns.diffusion = model.eval(), ns.arch.build_diffusion(C.diffusion_test_cfg())
        _REF["ns"] = ns
    except Exception as e:                                  # noqa: BLE001  (report, then fall back to the port)
        sys.stderr.write(f"bench.py: reference not importable ({type(e).__name__}: {e}); using the oracle port\n")
    return _REF["ns"]


def reference_sample(n_clips, n_exemplars, threads, plain=False):
    """Bounded sample of configs[1] executed by the reference's OWN code (gaussian_diffusion.py:1137
    ddim_reverse_sample_loop once per exemplar at B=1, as diffusion_architecture.py:323-354 does, then :1233
    ddim_guided_sample_loop with decreasing_till_25 insertion guidance -- autograd gradient steps included), all
    50 levels.  plain=True: configs[0], ddim_sample_loop at B=n_clips.  Returns (clip_steps, seconds)."""
    from rag_gesture_b200 import config as C
    from rag_gesture_b200 import synthetic as S
    ns = reference_modules()
    model, diff = ns.model, ns.diffusion
    torch.set_num_threads(threads)
    T, D, n = C.N_TOKENS, C.LATENT_DIM, C.N_CHUNKS

    def kwargs(cond, B):
        pc = model.get_precompute_condition(device="cpu", text=cond["word"], audio=cond["audio"],
                                            speaker_ids=cond["speaker_ids"], re_dict=1)
        return dict(xf_out=pc["xf_out"], re_dict=None, query_mask=S.query_masks(B), motion_mask=S.motion_mask(B),
                    sample_idx=None)
    cond = S.synthetic_conditions(n_clips + n_exemplars, seed=5)
    t0 = time.perf_counter()
    cb = {k: v[:n_clips] for k, v in cond.items()}
    if plain:
        with torch.no_grad():
            diff.ddim_sample_loop(model, (n_clips, T, D), clip_denoised=False, model_kwargs=kwargs(cb, n_clips), eta=0)
        return n_clips * STEPS, time.perf_counter() - t0
    inv_list = torch.zeros(STEPS, n_clips, T, D)
    with torch.no_grad():
        for e in range(n_exemplars):
            c1 = {k: v[n_clips + e:n_clips + e + 1] for k, v in cond.items()}
            inv = diff.ddim_reverse_sample_loop(model, start_img=S.synthetic_latents(1, seed=6 + e, scale=0.5),
                                                clip_denoised=False, model_kwargs=kwargs(c1, 1), eta=0,
                                                return_all_timesteps=True)
            inv = torch.cat(inv, 0)
            b = e % n_clips
            inv_list[:, b, 2:5] = inv[:, 4:7]
            inv_list[:, b, n + 3:n + 6] = inv[:, n + 5:n + 8]
        kw = kwargs(cb, n_clips)
    start = torch.randn(n_clips, T, D)
    nz = inv_list[-1] != 0
    start[nz] = inv_list[-1][nz]
    with torch.inference_mode(False):
        diff.ddim_guided_sample_loop(model, (n_clips, T, D), noise=start, clip_denoised=False, model_kwargs=kw, eta=0,
                                     in_seq=None, guidance_iters=list(GUIDANCE), inverted_latent_list=inv_list,
                                     guidance_lr=0.1)
    return (n_clips + n_exemplars) * STEPS, time.perf_counter() - t0


def cpu_sample(n_clips, n_exemplars, threads):
    """(clip_steps, seconds, kind): the unmodified reference when it is available, else the oracle port."""
    if reference_modules() is not None:
        return reference_sample(n_clips, n_exemplars, threads) + ("reference",)
    return cpu_reference_sample(n_clips, n_exemplars, STEPS, threads) + ("port",)


def _cpu_sample_text(kind, n_clips, n_exemplars, cs, dt, threads):
    what = ("the UNMODIFIED reference (mogen.models: ReGestureTransformer + SpacedDiffusion.ddim_reverse_sample_loop per "
            "exemplar at B=1 + ddim_guided_sample_loop with autograd insertion guidance)" if kind == "reference"
            else "oracle port of the reference algorithm")
    return (f"{what}: {n_exemplars} exemplars inverted at B=1 + {n_clips} clips guided, all 50 DDIM levels = {cs} "
            f"clip-steps in {dt:.1f} s, torch CPU fp32, {threads} threads")


def cpu_sizes(sample_seconds, threads):
    """Clips / exemplars of the bounded sample: the B : E ratio of configs[1] (64 : 96) where the budget allows,
    scaled so that all 50 levels of both loops take about `sample_seconds`; never below 1 clip + 1 exemplar."""
    cs, dt, _ = cpu_sample(1, 1, threads)                       # calibration = warm-up: 100 clip-steps
    rate = cs / dt
    budget = rate * sample_seconds / STEPS                      # clips + exemplars that fit
    if budget < 4:
        return 1, 1
    units = max(1, min(8, int(budget / 5)))                     # 1 unit = 2 clips + 3 exemplars
    return 2 * units, 3 * units


def cpu_baseline(sample_seconds=15.0):
    threads = os.cpu_count() or 1
    nc, ne = cpu_sizes(sample_seconds, threads)
    cs, dt, kind = cpu_sample(nc, ne, threads)
    out = {"value": round(cs / dt, 2), "unit": UNIT, "cores": threads, "kind": kind,
           "sample": _cpu_sample_text(kind, nc, ne, cs, dt, threads)}
    if kind == "reference":                                     # configs[0] on the host cores, next to the GPU latency
        cs0, dt0 = reference_sample(1, 0, threads, plain=True)
        out["configs0_plain_ddim_b1"] = {"seconds": round(dt0, 3), "clip_steps_per_sec": round(cs0 / dt0, 2)}
    return out


def knn_cpu_baselines(n_full=1_000_000, dim=768, k=8, seconds=10.0):
    """(i) the reference's own ranking, sort_sidx_by_textsimilarity (rag/utils.py:86-132), on a
    10k-entry slice (entries/s); (ii) the strong baseline, fp32 torch.mm + torch.topk(k=8) over the full 1M x 768
    database on all host cores (queries/s), on a bounded number of queries."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    out = {"cores": threads}
    g = torch.Generator().manual_seed(42)
    ns = reference_modules()
    n_slice, Td = 10_000, 32
    feats = torch.randn(n_slice, Td, dim, generator=g)
    q = torch.randn(Td, dim, generator=g)
    if ns is not None:
        cache = {f"e{i}": [feats[i], 0] for i in range(n_slice)}
        names = list(cache.keys())
        t0 = time.perf_counter()
        order = ns.rag_utils.sort_sidx_by_textsimilarity(names, "", q, cache)
        dt = time.perf_counter() - t0
        out["text_similarity_reference"] = {"entries_per_sec": round(n_slice / dt, 1), "entries": n_slice, "tokens": Td,
                                            "seconds": round(dt, 2), "kind": "reference",
                                            "fn": "sort_sidx_by_textsimilarity (rag/utils.py:86)", "top": order[0]}
    db = torch.nn.functional.normalize(torch.randn(n_full, dim, generator=g), dim=1)
    qs = torch.nn.functional.normalize(torch.randn(64, dim, generator=g), dim=1)
    torch.topk(qs[:8] @ db.T, k, dim=1)                         # warm-up
    t0, nq = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds and nq < 4096:
        torch.topk(qs @ db.T, k, dim=1)
        nq += qs.shape[0]
    dt = time.perf_counter() - t0
    out["mm_topk_fp32"] = {"queries_per_sec": round(nq / dt, 1), "n": n_full, "dim": dim, "k": k, "queries": nq,
                           "seconds": round(dt, 2), "kind": "torch.mm + torch.topk, fp32, all host cores"}
    return out


def run_reference(args, emit):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores, same metric and
    config; each step is a bounded sample of the configs[1] workload (all 50 levels of both loops)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)              # torchrun exports OMP_NUM_THREADS=1: undo it before the first CPU op
    # whole run (K steps) bounded to about 2.5 minutes of host time; each step is one sample of both loops
    nc, ne = cpu_sizes(max(args.cpu_seconds, 150.0) / max(1, args.steps), threads)
    for _ in range(min(args.warmup, 1)):                        # the calibration run above already warmed up
        cpu_sample(1, 1, threads)
    cs_tot, t_tot, kind = 0, 0.0, "port"
    for _ in range(args.steps):
        cs, dt, kind = cpu_sample(nc, ne, threads)
        cs_tot, t_tot = cs_tot + cs, t_tot + dt
    v = cs_tot / t_tot
    sample = "per step: " + _cpu_sample_text(kind, nc, ne, cs_tot // args.steps, t_tot / args.steps, threads)
    emit({
        "impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_tot / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_gpu": B_PER_GPU, "db_entries": N_DB, "ddim_steps": STEPS,
                   "sample": f"bounded sample of that workload on the host cores: {nc} clips + {ne} exemplars per step, all 50 "
                             "levels of the inversion and the guided loop (reference's own code, no retrieval kernels)"},
        "cpu_baseline": {"value": round(v, 2), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--knn-n", type=int, default=1_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ncu-steps", type=int, default=0, help="profiling pass: loops cut to S levels, nothing timed")
    ap.add_argument("--longform-seconds", type=float, default=600.0, help="configs[4] stream length (0: skip)")
    ap.add_argument("--pair128-min-rows", type=int, default=0, help="row threshold of the pair128 kernel (0: library default)")
    ap.add_argument("--gemm-kernel", type=int, default=0, choices=[0, 1, 2, 3],
                    help="rg_set_gemm_kernel: 0 automatic, 1 always the 128x128 kernel, 2 the 2-CTA kernel when eligible")
    ap.add_argument("--gemm2-min-rows", type=int, default=0, help="row threshold of the automatic choice (0: library default)")
    ap.add_argument("--gemm2-persist-tiles", type=int, default=0, help="pair tiles from which the 2-CTA kernel is persistent (0: default)")
    a = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line on the first
    # communicator), so fd 1 points at stderr while the run lasts and the JSON lines go to the saved descriptor.
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    _emit = lambda obj: print(json.dumps(obj), file=_real_stdout, flush=True)
    if a.impl == "reference":
        run_reference(a, _emit)
    else:
        run_b200(a, _emit)
