#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 denoising engine (see DESIGN.md "Measurement").

Metric (BASELINE.json): images/sec (+ UNet it/s) for SD1.5 512x512 txt2img, 20 Euler steps, cfg 7,
random-init weights generated locally (no network), synthetic prompt, seeds 42+i.

One "step" = one call of the public API (mlis_generate) producing a batch of B images per GPU:
CLIP text encode (cond + uncond), 20 sampler steps = 20 batched UNet evaluations of 2B latents
(cond | uncond halves), VAE decode, RGB8 pack. Ranks are independent (images of a batch are
data-parallel, weights replicated); NCCL only gathers the final RGB8 images on rank 0.

  e2e   : THE HEADLINE. images/s through the public API with host buffers: prompt string in, RGB8 images out
          (tokenise + CLIP + noise upload + denoise + decode + D2H + NCCL gather of the device-resident
          RGB8 images), wall clock, max over ranks
  value : images/s over all ranks, GPU time of the generation (CUDA events on the engine stream),
          conditioning already resident in HBM (MLIS_TUF_CONDITIONING), max over ranks
  roofline     : the tcgen05 GEMM / implicit-conv kernel family (dominant kernel), algorithmic
                 FLOPs / CUDA-event time of those launches in one profiled UNet evaluation
  roofline_hbm : the HBM-bound kernel families of the same evaluation (GB/s, fraction of the measured copy bandwidth)
  cpu_baseline : the reference host code on the CPU oracle (oracle/_ref), bounded sample

Workloads (`--workload`): c1 (default, BASELINE configs[0] shape, weak scaling, + the SDXL side measurement),
c2 (SD2.1 768x768 v-prediction, 25 DPM++2M steps, global batch 8 sharded over the GPUs: strong scaling),
c3 (SDXL 1024x1024, 30 Euler steps, cfg 7, global batch 16: strong scaling),
c5 (SDXL 2048x2048 tiled VAE decode, vae-tile 512, the 16 tiles spread across the GPUs + TAE decode: strong scaling).

`--impl reference` times the reference's own CPU path (reference host objects + oracle ggml ops) on the same
configuration: cfg 7, Euler, 512x512, every sampler step = 2 UNet evaluations.
"""
import argparse, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

PROMPT = "a photograph of an astronaut riding a horse"
WORKLOAD = "SD1.5 txt2img 512x512, 20 Euler steps, cfg 7, uniform schedule, random-init weights (BASELINE configs[0] shape)"
N_STEPS_SAMPLER = 20
FLOP_PER_NFE_SD15_512 = 0.803e12      # BASELINE.md section 2 (dense 2*MAC, conv + linear + attention)
FLOP_VAE_512 = 2.515e12
FLOP_PER_NFE_SD21_768 = 2.149e12
FLOP_PER_NFE_SDXL_1024 = 6.761e12
FLOP_VAE_1024 = 10.470e12
FLOP_VAE_TILE_80 = 3.959e12           # one 80x80-latent tile of the SDXL decoder
FLOP_TAE_2048 = 2.262e12
SDXL_STEPS = 30

GEN_WORKLOADS = {
    "c1": dict(kind="sd1", dim=(512, 512), steps=20, method="euler", cfg=7, nfe_per_step=2, flop_nfe=FLOP_PER_NFE_SD15_512,
               name=WORKLOAD, scaling="weak"),
    "c2": dict(kind="sd2", dim=(768, 768), steps=25, method="dpmpp2m", cfg=7, nfe_per_step=2, flop_nfe=FLOP_PER_NFE_SD21_768, global_batch=8,
               name="SD2.1 txt2img 768x768 v-prediction, 25 DPM++2M steps, cfg 7, global batch 8 sharded over the GPUs (BASELINE configs[1])", scaling="strong"),
    "c3": dict(kind="sdxl", dim=(1024, 1024), steps=30, method="euler", cfg=7, nfe_per_step=2, flop_nfe=FLOP_PER_NFE_SDXL_1024, global_batch=16,
               name="SDXL base txt2img 1024x1024, 30 Euler steps, cfg 7, full VAE decode, global batch 16 sharded over the GPUs (BASELINE configs[2])", scaling="strong"),
}


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        p = {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "src": "measured"}
    except Exception:
        pass
    return p


def weights_path(kind="sd1"):
    """Random-init checkpoint, generated once per box (untimed)."""
    d = os.environ.get("MLIS_BENCH_WEIGHTS", "/tmp/mlis_b200_weights")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, kind + ".safetensors")
    if not os.path.exists(p):
        import gen_weights
        tmp = p + ".tmp%d" % os.getpid()
        gen_weights.write_safetensors(tmp, gen_weights.build_spec(kind), 1234, "f16")
        os.replace(tmp, p)
    return p


class ClockSampler(threading.Thread):
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append([x.strip() for x in o])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        reasons = set()
        for s in self.samples:
            if len(s) >= 7:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_run(model, n_sampler_steps, threads=None, cfg=7):
    """The reference's host code + the oracle's CPU ops on the bench configuration (512x512, Euler, cfg 7 => every
    sampler step is TWO UNet evaluations, mlimgsynth.c:1572-1587). Returns the per-NFE, CLIP and VAE-decode seconds the
    reference itself logs."""
    exe = os.path.join(ROOT, "oracle", "_ref", "mlimgsynth_cpu")
    if not os.path.exists(exe):
        return None
    out = "/tmp/mlis_ref_out_%d" % os.getpid()
    cmd = [exe, "generate", "-m", model, "-p", PROMPT, "-d", "512,512", "-S", "42", "-s", str(n_sampler_steps), "--method", "euler",
           "--cfg-scale", str(cfg), "-o", out + ".pnm", "-v"]
    if threads:
        cmd += ["-t", str(threads)]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    wall = time.time() - t0
    nfe, vae, clip = [], None, []
    for line in r.stderr.splitlines():
        if "NFE" in line and "done {" in line:
            nfe.append(float(line.split("{")[1].split("s}")[0]))
        if "VAE decode done {" in line:
            vae = float(line.split("{")[1].split("s}")[0])
        if "CLIP text encode done {" in line:
            clip.append(float(line.split("{")[1].split("s}")[0]))
    if r.returncode != 0 or not nfe:
        return None
    return {"nfe_s": nfe, "vae_s": vae, "clip_s": clip, "wall_s": wall}


def reference_images_per_s(res, warm_nfe):
    """images/s of ONE image of the workload = 1 / (2 CLIP encodes + 40 UNet evaluations + 1 VAE decode), from the times of a
    run at the workload's own settings; when that run had fewer than 20 sampler steps, the mean UNet evaluation stands
    for the missing ones (said in the sample text)."""
    nfe = res["nfe_s"][warm_nfe:] or res["nfe_s"]
    t_nfe = sum(nfe) / len(nfe)
    full = len(res["nfe_s"]) >= 2 * N_STEPS_SAMPLER
    t_unet = sum(res["nfe_s"][:2 * N_STEPS_SAMPLER]) if full else 2 * N_STEPS_SAMPLER * t_nfe
    total = t_unet + (res["vae_s"] or 0.0) + sum(res["clip_s"])
    return 1.0 / total, t_nfe, full


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, SAME configuration as the
    B200 arm (512x512, Euler, cfg 7 -- two UNet evaluations per sampler step, prompt + empty negative prompt through CLIP,
    VAE decode). A step = one sampler step; min(steps + warmup, 20) sampler steps are run, so the driver's default
    (--steps 20) is the complete 20-step generation, measured, not extrapolated."""
    if rank != 0:
        return
    cores = os.cpu_count()
    model = weights_path("sd1")
    n = max(1, min(args.steps + args.warmup, N_STEPS_SAMPLER))
    res = cpu_reference_run(model, n, cores)
    if not res:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/mlimgsynth_cpu missing or failed"}))
        return
    warm = 2 * min(args.warmup, n - 1)
    img_s, t_nfe, full = reference_images_per_s(res, warm)
    sample = ("%s20-step generation: %d sampler steps = %d UNet evaluations at cfg 7 + 2 CLIP encodes + 1 VAE decode at 512x512 by the reference's host code "
              "(oracle/_ref, built from /root/reference) on the restated ggml CPU ops; ggml itself is not on this machine; s/NFE %.2f, CLIP %.2f s, VAE %.2f s%s"
              % ("complete " if full else "sample of the ", n, len(res["nfe_s"]), t_nfe, sum(res["clip_s"]), res["vae_s"] or 0.0,
                 "" if full else "; the remaining sampler steps are counted at the mean UNet-evaluation time"))
    line = {
        "impl": "reference", "metric": "images_per_sec", "value": img_s, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 2 * t_nfe * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate (ggml CPU rounding points)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": 1, "global_batch": 1, "cfg_scale": 7, "sampler_steps_run": n, "complete_generation": full,
                   "step": "one sampler step = 2 UNet evaluations (cond, uncond); the reference generates one image per call (mlimgsynth.c:1640), so its images/s does not depend on the batch",
                   "unet_it_per_s": 1.0 / (2 * t_nfe), "nfe_per_s": 1.0 / t_nfe, "vae_decode_s": res["vae_s"], "clip_encode_s": sum(res["clip_s"])},
        "cpu_baseline": {"value": img_s, "unit": "images/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": img_s, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ B200 arm
class Engine:
    """The engine library instance the host library uses (stats, device timer, per-kernel profile)."""

    def __init__(self):
        import ctypes as C
        self.C = C
        self.lib = C.CDLL(os.path.join(ROOT, "mlimgsynth_b200", "lib", "libggml_b200.so"), mode=C.RTLD_LOCAL)
        self.lib.ggml_b200_timer_stop.restype = C.c_double

        class St(C.Structure):
            _fields_ = [(k, C.c_uint64) for k in ("kernel_launches", "graph_launches", "plans_built", "h2d_bytes", "d2h_bytes")]
        self.St = St
        self.lib.ggml_b200_get_stats.restype = C.POINTER(St)

    def stats(self):
        s = self.lib.ggml_b200_get_stats().contents
        return {k: getattr(s, k) for k, _ in self.St._fields_}

    def timer_start(self):
        self.lib.ggml_b200_timer_start()

    def timer_stop_ms(self):
        return self.lib.ggml_b200_timer_stop()

    def profile(self, fn):
        C = self.C
        self.lib.ggml_b200_profile_enable(1)
        fn()
        names = {0: "copy", 7: "gemm_simt", 8: "groupnorm", 9: "layernorm", 10: "geglu", 11: "im2col", 13: "attention", 14: "gemm_tc", 15: "conv3x3_tc", 3: "upscale"}
        prof = {}
        for k, nm in names.items():
            ms, fl, by, ln = C.c_double(), C.c_double(), C.c_double(), C.c_uint64()
            self.lib.ggml_b200_profile_get(k, C.byref(ms), C.byref(fl), C.byref(by), C.byref(ln))
            if ln.value:
                prof[nm] = {"ms": ms.value, "tflop": fl.value / 1e12, "gb": by.value / 1e9, "launches": ln.value}
        self.lib.ggml_b200_profile_enable(0)
        return prof


class Dist:
    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x):
        from mlimgsynth_b200 import dist as D
        return D.all_max(x, device="cuda") if self.world > 1 else x

    def prepare_weights(self, *kinds):
        if self.rank == 0:
            for k in kinds:
                weights_path(k)
        self.barrier()

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def roofline_blocks(eng, ctx, B, pk, nfe_ms_fn):
    """Per-kernel-family numbers of ONE UNet evaluation of 2B latents (profiled eager pass, outside the timed regions)."""
    import numpy as np
    x = np.random.default_rng(0).standard_normal((2 * B, 4, 64, 64)).astype(np.float32)
    cond = (np.random.default_rng(1).standard_normal((2 * B, 77, 768)) * 0.5).astype(np.float32)
    ctx.unet_eval(x, cond, None, 5.0)      # builds/captures the batch-2B graph used for NFE timing
    ctx.unet_eval(x, cond, None, 5.0)
    eng.timer_start()
    for _ in range(5):
        ctx.unet_eval(x, cond, None, 5.0)
    nfe_ms = eng.timer_stop_ms() / 5
    prof = eng.profile(lambda: ctx.unet_eval(x, cond, None, 5.0))
    g = [prof.get("gemm_tc", {}), prof.get("conv3x3_tc", {})]
    g_ms = sum(p.get("ms", 0) for p in g); g_fl = sum(p.get("tflop", 0) for p in g); g_n = sum(p.get("launches", 0) for p in g); g_gb = sum(p.get("gb", 0) for p in g)
    achieved = g_fl / (g_ms / 1e3) if g_ms else 0.0
    traffic, traffic_launches, traffic_src = None, 0, None
    for f in ("r2_gemm_traffic.json", "r1_gemm_traffic.json"):      # DRAM bytes per launch of the same family from the committed ncu capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", f)))
            traffic = tj["dram_bytes_per_launch"]; traffic_launches = tj["launches"]; traffic_src = f
            break
        except Exception:
            pass
    unet_tf = FLOP_PER_NFE_SD15_512 * 2 * B / (nfe_ms / 1e3) / 1e12
    roof = {"bound": "tensor", "kernel": "gemm_tc_persistent_kernel (tcgen05 cta_group::2 GEMM + implicit 3x3 conv; all linear / conv launches of one UNet evaluation)", "achieved": achieved, "peak": pk["bf16_tflops"],
            "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"], "traffic": traffic,
            "traffic_note": "dram__bytes_read+write per launch, ncu capture of all %d tcgen05 GEMM/conv launches of one evaluation (profiles/%s); algorithmic bytes per launch %.1f MB" % (traffic_launches, traffic_src, g_gb * 1e3 / max(g_n, 1)),
            "algorithmic_tflop_per_launch": g_fl / max(g_n, 1), "peak_source": pk["src"],
            "launches_per_unet_eval": g_n, "avg_launch_us": g_ms * 1e3 / g_n if g_n else None,
            "unet_eval_ms_batch%d" % (2 * B): nfe_ms, "unet_tflops": unet_tf, "unet_frac_of_peak": unet_tf / pk["bf16_tflops"]}
    att = prof.get("attention")
    if att:
        roof["attention_tflops"] = att["tflop"] / (att["ms"] / 1e3)
        roof["attention_frac_of_peak"] = roof["attention_tflops"] / pk["bf16_tflops"]
    hbm = {"peak": pk["hbm_gbs"], "unit": "GB/s", "peak_source": pk["src"], "families": {}}
    tot_ms = 0.0
    for nm in ("groupnorm", "layernorm", "geglu", "copy", "upscale", "im2col"):
        p = prof.get(nm)
        if p and p["ms"] > 0:
            gbs = p["gb"] / (p["ms"] / 1e3)
            hbm["families"][nm] = {"achieved": gbs, "frac": gbs / pk["hbm_gbs"], "ms": p["ms"], "launches": p["launches"], "algorithmic_gb": p["gb"]}
            tot_ms += p["ms"]
    hbm["share_of_unet_eval"] = tot_ms / nfe_ms if nfe_ms else None
    return roof, hbm, prof


def gen_call(ctx, api, D, seed, n_global, rank, world, cached_cond, nb):
    # image i of the global batch gets seed + i whatever the world size; this rank owns a contiguous block
    ctx.set("seed", D.image_seeds(seed * 1000, n_global, rank, world)[0])
    ctx.set("prompt", PROMPT)
    if cached_cond:
        ctx.set("tensor_use_flags", api.TUF_CONDITIONING)
    ctx.generate()


def collect_images(ctx, D, dd, nb):
    """The e2e leg's output side: every rank's RGB8 images end up in host memory of rank 0. With several ranks the RGB8
    images are gathered straight from the device buffers the pack kernel wrote (one NCCL gather, then one D2H on rank 0)."""
    if dd.world > 1:
        return D.gather_images_device(ctx, dst=0)
    return [ctx.image(i) for i in range(nb)]


def run_generate(args, dd, wl_name):
    import numpy as np
    from mlimgsynth_b200 import api, dist as D
    wl = GEN_WORKLOADS[wl_name]
    rank, world, local = dd.rank, dd.world, dd.local
    if wl["scaling"] == "weak":
        B = args.batch; n_global = B * world
    else:
        n_global = wl["global_batch"]
        if n_global % world:
            raise SystemExit("global batch %d does not divide over %d GPUs" % (n_global, world))
        B = n_global // world
    dd.prepare_weights(wl["kind"])
    os.environ.setdefault("GGML_B200_QUIET", "1")
    ctx = api.Ctx(backend="B200:%d" % local, model=weights_path(wl["kind"]), image_dim=wl["dim"], steps=wl["steps"], method=wl["method"],
                  cfg_scale=wl["cfg"], batch_size=B)
    eng = Engine()
    gen = lambda seed, cached: gen_call(ctx, api, D, seed, n_global, rank, world, cached, B)

    # ---- warm-up (builds graphs, uploads weights, captures CUDA graphs)
    W = max(args.warmup, 3)
    for i in range(W):
        gen(i, i > 0)
        collect_images(ctx, D, dd, B)     # also warms up the NCCL channels

    # ---- `value`: device time of K generations, conditioning resident
    clocks = ClockSampler(local); clocks.start()
    dd.barrier()
    s0 = eng.stats()
    eng.timer_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        gen(100 + i, True)
    dev_ms = eng.timer_stop_ms()
    dd.barrier()
    wall_cached = time.perf_counter() - t0
    s1 = eng.stats()
    dev_s = dd.allmax(dev_ms / 1e3)

    # ---- `e2e`: public API, host buffers in and out, + gather of the RGB8 images to rank 0
    dd.barrier()
    e0 = eng.stats()
    t0 = time.perf_counter()
    for i in range(args.steps):
        gen(200 + i, False)
        collect_images(ctx, D, dd, B)
    dd.barrier()
    e2e_s = dd.allmax(time.perf_counter() - t0)
    e1 = eng.stats()
    clk = clocks.summary()

    total_images = n_global * args.steps
    value = total_images / dev_s
    e2e = total_images / e2e_s
    unet_it_s = (wl["steps"] * total_images) / dev_s              # sampler steps x images per second
    pk = peaks()

    roof = hbm = None; prof = {}; vae = None; sdxl = None; cpu = None
    if wl_name == "c1":
        if rank == 0:
            roof, hbm, prof = roofline_blocks(eng, ctx, B, pk, None)
            # VAE decode of the batch alone (device time incl. the latent upload and the RGB8 download of the call)
            lat = (np.random.default_rng(3).standard_normal((B, 4, 64, 64)) * 0.18).astype(np.float32)
            ctx.decode(lat); ctx.decode(lat)
            eng.timer_start()
            for _ in range(3):
                ctx.decode(lat)
            v_ms = eng.timer_stop_ms() / 3
            vprof = eng.profile(lambda: ctx.decode(lat))
            k_ms = sum(p["ms"] for p in vprof.values())
            vae = {"vae_decode_call_ms_batch%d" % B: v_ms, "vae_kernels_ms_batch%d" % B: k_ms,
                   "note": "call = mlis_image_decode on host tensors (latent H2D, decoder graph, RGB8 pack, D2H of the f32 image and the RGB8 bytes); kernels = sum of the decoder's kernel times (profiled pass)",
                   "vae_tflops": FLOP_VAE_512 * B / (k_ms / 1e3) / 1e12, "vae_frac_of_peak": FLOP_VAE_512 * B / (k_ms / 1e3) / 1e12 / pk["bf16_tflops"],
                   "kernel_profile": vprof}
        if not args.no_sdxl:
            ctx.close(); ctx = None
            sdxl = sdxl_side(args, dd, eng, pk)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            res = cpu_reference_run(weights_path("sd1"), 1, os.cpu_count())
            if res:
                img_s, t_nfe, _ = reference_images_per_s(res, 0)
                cpu = {"value": img_s, "unit": "images/s", "cores": os.cpu_count(), "kind": "reference",
                       "sample": "1 sampler step at cfg 7 (2 UNet evaluations) + 2 CLIP encodes + 1 VAE decode at 512x512 by the reference host code on the restated ggml CPU ops "
                                 "(oracle/_ref); 40 UNet evaluations counted at the mean; s/NFE %.2f, CLIP %.2f s, VAE %.2f s" % (t_nfe, sum(res["clip_s"]), res["vae_s"] or 0.0)}
    else:
        if rank == 0:
            shape = (2 * B, 4, wl["dim"][1] // 8, wl["dim"][0] // 8)
            x = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
            nctx = {"sd2": 1024, "sdxl": 2048}[wl["kind"]]
            cond = (np.random.default_rng(1).standard_normal((2 * B, 77, nctx)) * 0.5).astype(np.float32)
            lab = (np.random.default_rng(2).standard_normal((2 * B, 2816)) * 0.5).astype(np.float32) if wl["kind"] == "sdxl" else None
            ctx.unet_eval(x, cond, lab, 5.0); ctx.unet_eval(x, cond, lab, 5.0)
            eng.timer_start()
            for _ in range(3):
                ctx.unet_eval(x, cond, lab, 5.0)
            nfe_ms = eng.timer_stop_ms() / 3
            tf = wl["flop_nfe"] * 2 * B / (nfe_ms / 1e3) / 1e12
            roof = {"bound": "tensor", "kernel": "whole UNet evaluation of %d latents" % (2 * B), "achieved": tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": tf / pk["bf16_tflops"], "traffic": None, "unet_eval_ms_batch%d" % (2 * B): nfe_ms, "peak_source": pk["src"]}

    if rank == 0:
        line = {
            "metric": "images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": dev_s * 1e3 / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": B, "global_batch": n_global, "parallelism": "dp%d (images), cfg halves batched" % world,
                       "unet_it_per_s": unet_it_s, "nfe_per_s": wl["nfe_per_step"] * unet_it_s,
                       "l2": "working set (weights 1.7-5.1 GB + activations) exceeds the 126 MB L2: every UNet evaluation streams all weights from HBM",
                       "headline": "e2e (public API, host buffers); value is the device-timed generation with conditioning resident",
                       "sdxl_1024": sdxl},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": (e1["h2d_bytes"] - e0["h2d_bytes"]) // args.steps,
                    "d2h_bytes_per_step": (e1["d2h_bytes"] - e0["d2h_bytes"]) // args.steps, "ms_per_step": e2e_s * 1e3 / args.steps},
            "gpu_launches": s1["kernel_launches"] - s0["kernel_launches"], "graph_replays": s1["graph_launches"] - s0["graph_launches"],
            "wall_ms_per_step_cached_cond": wall_cached * 1e3 / args.steps,
            "clocks": clk, "roofline": roof, "roofline_hbm": hbm, "vae": vae, "cpu_baseline": cpu, "kernel_profile": prof,
        }
        print(json.dumps(line))
    if ctx is not None:
        ctx.close()


def sdxl_side(args, dd, eng, pk):
    """Second half of the BASELINE metric: SDXL base 1024x1024, 30 Euler steps, cfg 7, full VAE decode (BASELINE configs[2]
    shape at --sdxl-batch images per GPU), --sdxl-reps timed generations. Reported under config.sdxl_1024; the headline stays SD1.5."""
    import numpy as np
    from mlimgsynth_b200 import api, dist as D
    rank, world, local = dd.rank, dd.world, dd.local
    dd.prepare_weights("sdxl")
    xb = args.sdxl_batch
    cx = api.Ctx(backend="B200:%d" % local, model=weights_path("sdxl"), image_dim=(1024, 1024), steps=SDXL_STEPS, method="euler", cfg_scale=7, batch_size=xb)
    genx = lambda seed, cached: gen_call(cx, api, D, seed, xb * world, rank, world, cached, xb)
    genx(0, False); genx(1, True)
    clocks = ClockSampler(local); clocks.start()
    reps = max(1, args.sdxl_reps)
    dev, e2e = [], []
    for r in range(reps):
        dd.barrier()
        eng.timer_start()
        genx(10 + r, True)
        dev.append(dd.allmax(eng.timer_stop_ms() / 1e3))
    for r in range(reps):
        dd.barrier()
        t0 = time.perf_counter()
        genx(20 + r, False)
        collect_images(cx, D, dd, xb)
        dd.barrier()
        e2e.append(dd.allmax(time.perf_counter() - t0))
    clk = clocks.summary()
    out = None
    if rank == 0:
        lat = np.random.default_rng(0).standard_normal((2 * xb, 4, 128, 128)).astype(np.float32)
        cond = (np.random.default_rng(1).standard_normal((2 * xb, 77, 2048)) * 0.5).astype(np.float32)
        lab = (np.random.default_rng(2).standard_normal((2 * xb, 2816)) * 0.5).astype(np.float32)
        cx.unet_eval(lat, cond, lab, 5.0); cx.unet_eval(lat, cond, lab, 5.0)
        nfe = []
        for _ in range(3):
            eng.timer_start()
            cx.unet_eval(lat, cond, lab, 5.0)
            nfe.append(eng.timer_stop_ms())
        x_nfe = sorted(nfe)[1]
        tf = FLOP_PER_NFE_SDXL_1024 * 2 * xb / (x_nfe / 1e3) / 1e12
        md, me = sorted(dev)[len(dev) // 2], sorted(e2e)[len(e2e) // 2]
        out = {"workload": "SDXL base txt2img 1024x1024, %d Euler steps, cfg 7, full VAE decode, random-init weights" % SDXL_STEPS,
               "batch_per_gpu": xb, "repeats": reps, "images_per_sec": xb * world / md, "e2e_images_per_sec": xb * world / me,
               "images_per_sec_all": [xb * world / d for d in dev],
               "unet_it_per_s": SDXL_STEPS * xb * world / md, "ms_per_generation": md * 1e3,
               "unet_eval_ms_batch%d" % (2 * xb): x_nfe, "unet_eval_ms_all": nfe, "unet_tflops": tf, "unet_frac_of_peak": tf / pk["bf16_tflops"], "clocks": clk}
    cx.close()
    return out


def run_c5(args, dd):
    """BASELINE configs[4]: SDXL 2048x2048 tiled VAE decode (vae-tile 512 = 16 tiles of 80x80 latent pixels, vae.c:331-391) with the
    tiles spread round-robin over the GPUs, one NCCL gather of the decoded tiles (device to device) and the ordered merge on
    rank 0; plus the TAE decode of the same latent (one GPU, full frame). Strong scaling: the 16 tiles are the fixed work."""
    import numpy as np
    from mlimgsynth_b200 import api, dist as D
    rank, world, local = dd.rank, dd.world, dd.local
    dd.prepare_weights("sdxl", "sd1", "tae")
    os.environ.setdefault("GGML_B200_QUIET", "1")
    ctx = api.Ctx(backend="B200:%d" % local, model=weights_path("sdxl"), vae_tile=512)
    eng = Engine()
    lat = (np.random.default_rng(42).standard_normal((1, 4, 256, 256)) * 0.18).astype(np.float32)
    n_tiles, tw, th = ctx.vae_tile_plan(256, 256)
    W = max(args.warmup, 3)
    for _ in range(W):
        D.vae_decode_tiled(ctx, lat)
    clocks = ClockSampler(local); clocks.start()
    dd.barrier()
    s0 = eng.stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        img = D.vae_decode_tiled(ctx, lat)
    dd.barrier()
    e2e_s = dd.allmax(time.perf_counter() - t0)
    s1 = eng.stats()
    # device-resident variant: the latent upload / image download of the call are there too (the API takes host tensors); the
    # engine-stream time of this rank's tile decodes alone is what `value` reports
    slots = (n_tiles + world - 1) // world
    import torch
    buf = torch.empty((slots, 3 * th * tw), dtype=torch.float32, device="cuda")
    dd.barrier()
    eng.timer_start()
    for _ in range(args.steps):
        ctx.vae_tiles_decode(lat, rank, world, buf.data_ptr())
    dev_s = dd.allmax(eng.timer_stop_ms() / 1e3)
    clk = clocks.summary()
    tae = None
    if rank == 0:
        ct = api.Ctx(backend="B200:%d" % local, model=weights_path("sd1"), tae=weights_path("tae"))
        ct.decode(lat); ct.decode(lat)
        t0 = time.perf_counter()
        for _ in range(3):
            ct.decode(lat)
        t_tae = (time.perf_counter() - t0) / 3
        tae = {"workload": "TAE decode of the same 256x256 latent -> 2048x2048, full frame, one GPU (tae.c:117)", "ms": t_tae * 1e3,
               "tflops": FLOP_TAE_2048 / t_tae / 1e12, "images_per_sec": 1.0 / t_tae}
        ct.close()
    if rank == 0:
        pk = peaks()
        tf = FLOP_VAE_TILE_80 * n_tiles * args.steps / dev_s / 1e12
        line = {
            "metric": "images_per_sec", "value": args.steps / dev_s, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": dev_s * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
            "config": {"workload": "SDXL VAE decode 2048x2048, vae-tile 512 (%d tiles of 80x80 latent pixels) spread over the GPUs + TAE decode (BASELINE configs[4])" % n_tiles,
                       "tiles": n_tiles, "tiles_per_gpu": slots, "parallelism": "tiles round-robin over %d GPUs, one NCCL gather (%.1f MB per tile, device to device), ordered merge on rank 0" % (world, 3 * th * tw * 4 / 1e6),
                       "value_is": "max over ranks of the engine-stream time of a rank's tile decodes", "tae": tae,
                       "l2": "each tile's activations (up to 640x640x128 f16 = 105 MB per tensor) exceed L2 at the last two levels"},
            "e2e": {"value": args.steps / e2e_s, "unit": "images/s", "ms_per_step": e2e_s * 1e3 / args.steps,
                    "h2d_bytes_per_step": lat.nbytes, "d2h_bytes_per_step": 2048 * 2048 * 3,
                    "note": "host latent in, RGB8 image out on rank 0: tile decodes + NCCL gather + merge + RGB8 pack + D2H"},
            "gpu_launches": s1["kernel_launches"] - s0["kernel_launches"], "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "SDXL VAE decoder, one 80x80 tile (conv3x3 family dominates)", "achieved": tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": tf / pk["bf16_tflops"] / world, "traffic": None, "note": "frac is per GPU: achieved / (peak x n_gpus)", "peak_source": pk["src"]},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c1", choices=["c1", "c2", "c3", "c5"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step (c1: weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sdxl", action="store_true", help="skip the SDXL 1024x1024 side measurement (second half of the BASELINE metric)")
    ap.add_argument("--sdxl-batch", type=int, default=2, help="SDXL images per GPU per generation")
    ap.add_argument("--sdxl-reps", type=int, default=3, help="timed SDXL generations (median reported)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    dd = Dist()
    if args.workload == "c5":
        run_c5(args, dd)
    else:
        run_generate(args, dd, args.workload)
    dd.close()


if __name__ == "__main__":
    main()
