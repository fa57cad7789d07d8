"""bench.py --config 3|4|5: the other BASELINE.json configs under the same JSON contract (one line on stdout).

  3  CelebHQ LDM ControlNet on VAE latents (config/celebhq.yaml), canny hint 3x1024x1024, 1000 steps + VAE decode,
     global batch 256 sharded over the ranks               metric: images/s of the whole job
  4  consistency-distilled ControlNet, single step, batch 4096: MNIST shapes and the CelebHQ-latent variant
     (dict(ldm_params, im_channels=4, im_size=32))   metric: samples/s (MNIST), second shape in config
  5  distribution-matching distilled ControlNet, single step, CIFAR 32x32, batch sweep 1..8192 (graph replay)
     metric: samples/s at the best batch; the sweep is in config.sweep

Not collected by pytest (no test_ prefix); imported by bench.py only.  Inputs are created on the host and copied to the
device inside the `e2e` region; `value` is device-timed with inputs resident.
"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mods():
    import torch
    pk = "controlnet-pytorch_b200."
    return (torch, importlib.import_module(pk + "runtime"), importlib.import_module(pk + "sampler"),
            importlib.import_module(pk + "utils.synthetic"))


def _fill(m, syn, seed=0):
    m.load_state_dict(syn.det_state_dict(m.state_dict(), seed))
    return m.cuda().eval()


def _timeit(torch, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _emit(line):
    print(json.dumps(line), flush=True)


def run(args):
    if args.impl == "reference":
        _emit({"impl": "reference", "unavailable": "the CPU arm is implemented for the headline config (2) only"})
        return
    {3: config3, 4: config4, 5: config5}[args.config](args)


def config3(args):
    torch, rt, S, syn = _mods()
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench as BN
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    world, rank, local, dev = BN._setup_dist()
    rt.lib()
    rt.set_mode(args.mode)
    G = args.batch or 256
    lo, hi = S.shard_bounds(G, world, rank)
    B = hi - lo
    CN = importlib.import_module("controlnet-pytorch_b200.models.controlnet_ldm").ControlNet
    VAE = importlib.import_module("controlnet-pytorch_b200.models.vae").VAE
    sch = importlib.import_module("controlnet-pytorch_b200.scheduler.linear_noise_scheduler")
    m = _fill(CN(4, syn.CELEBHQ_LDM_PARAMS, down_sample_factor=32), syn)
    vae = _fill(VAE(3, syn.CELEBHQ_VAE_PARAMS), syn, 1)
    sched = sch.LinearNoiseScheduler(ldm_scheduler=True, **syn.CELEBHQ_DIFFUSION)
    smp = S.LDMSampler(m, sched, vae, seed=3, use_graph=True)
    # hints as uint8 on the host (0.8 GB for 256 samples), expanded to the reference's fp32 {0,1} x 3 channels on device
    g = torch.Generator().manual_seed(100 + rank)
    hint_u8 = (torch.rand(B, 1, 1024, 1024, generator=g) < 0.05).to(torch.uint8).pin_memory()
    per = 4 * 32 * 32

    def hint_dev():
        return hint_u8.to(dev, non_blocking=True).float().expand(B, 3, 1024, 1024).contiguous()
    hint = hint_dev()
    x_T = smp.draw_xT((B, 4, 32, 32), dev, elem_offset=lo * per)
    with torch.no_grad():
        smp._capture(x_T, hint, 1000, lo * per)
    clk = BN.ClockSampler(local)
    if rank == 0:
        clk.start()
        time.sleep(0.3)
    ms_step, w0, w1 = BN._timed_replay(smp, args.steps, args.warmup, world, dev)
    if rank == 0:
        clk.stop()
    with torch.no_grad():
        ms_dec = _timeit(torch, lambda: smp.decode(smp.xt), 2, warm=1)
        m._hint_cache.clear()
        ms_hint = _timeit(torch, lambda: (m._hint_cache.clear(), m._hint_feat(hint, rt.get_mode())), 1, warm=1)
    t = torch.tensor([ms_dec, ms_hint], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dec, ms_hint = float(t[0]), float(t[1])
    job_s = (1000 * ms_step + ms_dec + ms_hint) * 1e-3
    value = G / job_s
    # end to end, bounded: K timesteps + decode + gather with host buffers, scaled to 1000 steps
    k = args.steps
    xh = torch.randn(B, 4, 32, 32).pin_memory()
    out_h = torch.empty(G, 3, 128, 128).pin_memory() if rank == 0 else None
    smp2 = S.LDMSampler(m, sched, vae, seed=4, use_graph=True)

    def job():
        h = hint_dev()
        x = xh.to(dev, non_blocking=True)
        xt, _ = smp2.sample(x, h, steps=k, elem_offset=lo * per)
        ims = smp2.decode(xt)
        full = S.gather_shards(ims, G) if world > 1 else ims
        if rank == 0:
            out_h.copy_(full, non_blocking=True)
    with torch.no_grad():
        job()                        # captures the step graph
        job()                        # first replayed call: its eager parts (hint refresh, decode) size the allocator's pools
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        job()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
    loop_s = dt - (ms_dec + ms_hint) * 1e-3
    e2e_s = dt + max(loop_s, 0.0) * (1000 - k) / k
    line = {"metric": "controlnet_ldm_images_per_sec", "value": round(value, 3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else "f16", "data": "synthetic",
            "config": {"workload": "celebhq_ldm_controlnet_1000step_plus_vae_decode", "global_batch": G, "batch_per_gpu": B,
                       "hint": "3x1024x1024 Bernoulli(0.05)", "latent": "4x32x32", "decode_ms": round(ms_dec, 1),
                       "hint_pyramid_ms_once": round(ms_hint, 1), "job_seconds_1000_steps": round(job_s, 2),
                       "model_tflops": round(G * 64.64e9 / (ms_step * 1e-3) / 1e12, 1),
                       "l2": "activations per step far exceed the 126 MB L2"},
            "e2e": {"value": round(G / e2e_s, 3), "unit": "images/s", "h2d_bytes_per_step": int((hint_u8.numel() + xh.numel() * 4) / 1000),
                    "d2h_bytes_per_step": int(G * 3 * 128 * 128 * 4 / 1000), "timesteps_run": k,
                    "note": "uint8 hint + x_T H2D, %d timesteps, VAE decode, all-gather, D2H of the images; the loop part is "
                            "scaled to 1000 timesteps" % k},
            "gpu_launches": int(smp.launches_per_step * args.steps), "launches_per_step": int(smp.launches_per_step),
            "clocks": clk.summary(w0, w1) if rank == 0 else None,
            "mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        _emit(line)


def _student_line(metric, workload, value, ms, B, extra, args, e2e):
    return {"metric": metric, "value": round(value, 1), "unit": "samples/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else "f16", "data": "synthetic",
            "config": dict({"workload": workload, "batch": B}, **extra), "e2e": e2e}


def _hints(torch, B, s):
    return (torch.rand(B, 1, s, s) < 0.1).float().expand(B, 3, s, s).contiguous()


def _e2e_student(torch, model, xh, ch, hh, reps=3):
    """public forward with pinned-host inputs copied in and the result copied out inside the timed region"""
    out_h = torch.empty_like(xh).pin_memory()

    def call():
        y = model(xh.cuda(non_blocking=True), ch.cuda(non_blocking=True), hh.cuda(non_blocking=True))
        out_h.copy_(y, non_blocking=True)
        torch.cuda.synchronize()
    call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    return (time.perf_counter() - t0) / reps


def config4(args):
    torch, rt, S, syn = _mods()
    rt.lib()
    rt.set_mode(args.mode)
    torch.cuda.set_device(0)
    CS = importlib.import_module("controlnet-pytorch_b200.models.consistency_controlnet_distilled").ConsistencyControlNet
    B = args.batch or 4096
    with torch.no_grad():
        m = _fill(CS(syn.MNIST_PARAMS), syn)
        x, h, sg = torch.randn(B, 1, 28, 28).pin_memory(), _hints(torch, B, 28).pin_memory(), torch.full((B,), 80.0).pin_memory()
        xd, hd, sd = x.cuda(), h.cuda(), sg.cuda()
        c0 = rt.launch_count()
        m(xd, sd, hd)
        launches = rt.launch_count() - c0
        ms = _timeit(torch, lambda: m(xd, sd, hd), args.steps, warm=max(args.warmup, 3))
        e2e_s = _e2e_student(torch, m, x, sg, h)
        del m
        torch.cuda.empty_cache()
        lat = dict(syn.CELEBHQ_LDM_PARAMS, im_channels=4, im_size=32)
        m2 = _fill(CS(lat), syn)
        x2, h2 = torch.randn(B, 4, 32, 32, device="cuda"), _hints(torch, B, 32).cuda()
        ms2 = _timeit(torch, lambda: m2(x2, sd, h2), max(2, args.steps // 5), warm=2)
    line = _student_line("consistency_controlnet_samples_per_sec", "consistency_single_step_mnist", B / ms * 1e3, ms, B,
                         {"sigma": 80.0, "model_tflops": round(2.119e9 * B / (ms * 1e-3) / 1e12, 1),
                          "celebhq_latent_shape": {"batch": B, "ms": round(ms2, 2), "samples_per_s": round(B / ms2 * 1e3, 1),
                                                   "model_tflops": round(33.85e9 * B / (ms2 * 1e-3) / 1e12, 1)},
                          "l2": "activations per call exceed L2"}, args,
                         {"value": round(B / e2e_s, 1), "unit": "samples/s", "h2d_bytes_per_step": int((x.numel() + h.numel() + B) * 4),
                          "d2h_bytes_per_step": int(x.numel() * 4)})
    line["gpu_launches"] = int(launches * args.steps)
    _emit(line)


def config5(args):
    torch, rt, S, syn = _mods()
    rt.lib()
    rt.set_mode(args.mode)
    torch.cuda.set_device(0)
    DM = importlib.import_module("controlnet-pytorch_b200.models.distribution_matching_controlnet").DistributionMatchingControlNet
    sweep = []
    with torch.no_grad():
        m = _fill(DM(syn.CIFAR_PARAMS), syn)
        g = S.GraphedStudent(m)
        best = None
        launches = 0
        for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
            x, h, t = torch.randn(B, 3, 32, 32, device="cuda"), _hints(torch, B, 32).cuda(), torch.full((B,), 999, device="cuda")
            fn = (lambda: g(x, t, h)) if B <= 512 else (lambda: m(x, t, h))
            ms = _timeit(torch, fn, 10 if B <= 1024 else 3, warm=2)
            rec = {"batch": B, "ms": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1), "path": "graph" if B <= 512 else "eager",
                   "model_tflops": round(9.394e9 * B / (ms * 1e-3) / 1e12, 1)}
            sweep.append(rec)
            if best is None or rec["samples_per_s"] > best["samples_per_s"]:
                best = rec
        Bb = best["batch"]
        x, h, t = torch.randn(Bb, 3, 32, 32).pin_memory(), _hints(torch, Bb, 32).pin_memory(), torch.full((Bb,), 999).pin_memory()
        c0 = rt.launch_count()
        m(x.cuda(), t.cuda(), h.cuda())
        launches = rt.launch_count() - c0
        e2e_s = _e2e_student(torch, m, x, t, h)
    line = _student_line("dm_controlnet_samples_per_sec", "dm_single_step_cifar_batch_sweep", best["samples_per_s"], best["ms"], Bb,
                         {"sweep": sweep, "latency_ms_batch1": sweep[0]["ms"]}, args,
                         {"value": round(Bb / e2e_s, 1), "unit": "samples/s", "h2d_bytes_per_step": int((x.numel() + h.numel()) * 4 + Bb * 8),
                          "d2h_bytes_per_step": int(x.numel() * 4)})
    line["gpu_launches"] = int(launches)
    _emit(line)
