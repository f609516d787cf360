#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 denoising engine (see DESIGN.md "Measurement").

Metric (BASELINE.json): images/sec (+ UNet it/s) for SD1.5 512x512 txt2img, 20 Euler steps, cfg 7,
random-init weights generated locally (no network), synthetic prompt, seeds 42+i.

One "step" = one call of the public API (mlis_generate) producing a batch of B images per GPU:
CLIP text encode (cond + uncond), 20 sampler steps = 20 batched UNet evaluations of 2B latents
(cond | uncond halves), VAE decode, RGB8 pack. Ranks are independent (images of a batch are
data-parallel, weights replicated); NCCL only gathers the final RGB8 images on rank 0.

  value : images/s over all ranks, GPU time of the generation (CUDA events on the engine stream),
          conditioning already resident in HBM (MLIS_TUF_CONDITIONING), max over ranks
  e2e   : images/s through the public API with host buffers: prompt string in, RGB8 images out
          (tokenise + CLIP + noise upload + denoise + decode + D2H + NCCL gather), wall clock,
          max over ranks
  roofline     : the tcgen05 GEMM / implicit-conv kernel family (dominant kernel), algorithmic
                 FLOPs / CUDA-event time of those launches in one profiled UNet evaluation
  cpu_baseline : the reference host code on the CPU oracle (oracle/_ref), bounded sample

`--impl reference` times the reference's own CPU path (reference host objects + oracle ggml ops).
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
FLOP_PER_NFE_SDXL_1024 = 6.761e12
SDXL_STEPS = 30


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


def cpu_reference_run(model, n_nfe, threads=None):
    """Reference host code + oracle CPU ops: per-NFE seconds and one VAE decode (512x512)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "mlimgsynth_cpu")
    if not os.path.exists(exe):
        return None
    out = "/tmp/mlis_ref_out_%d" % os.getpid()
    cmd = [exe, "generate", "-m", model, "-p", PROMPT, "-d", "512,512", "-S", "42", "-s", str(n_nfe), "--method", "euler",
           "--cfg-scale", "1", "-o", out + ".pnm", "-v"]
    if threads:
        cmd += ["-t", str(threads)]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    wall = time.time() - t0
    nfe, vae = [], None
    for line in r.stderr.splitlines():
        if "NFE" in line and "done {" in line:
            nfe.append(float(line.split("{")[1].split("s}")[0]))
        if "VAE decode done {" in line:
            vae = float(line.split("{")[1].split("s}")[0])
    if r.returncode != 0 or not nfe:
        return None
    return {"nfe_s": nfe, "vae_s": vae, "wall_s": wall}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    cores = os.cpu_count()
    model = weights_path("sd1")
    n = args.steps + args.warmup
    res = cpu_reference_run(model, n, cores)
    if not res:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/mlimgsynth_cpu missing or failed"}))
        return
    nfe = res["nfe_s"][args.warmup:] or res["nfe_s"]
    t_nfe = sum(nfe) / len(nfe)
    vae = res["vae_s"] or 0.0
    img_s = 1.0 / (2 * N_STEPS_SAMPLER * t_nfe + vae)
    line = {
        "impl": "reference", "metric": "images_per_sec", "value": img_s, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_nfe * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate (ggml CPU rounding points)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": "one UNet evaluation (N=1) of the workload; images/s = 1/(40 NFE + 1 VAE decode)",
                   "unet_it_per_s": 1.0 / (2 * t_nfe), "nfe_per_s": 1.0 / t_nfe, "vae_decode_s": vae},
        "cpu_baseline": {"value": img_s, "unit": "images/s", "cores": cores, "kind": "reference",
                         "sample": "%d UNet evaluations + 1 VAE decode at 512x512 by the reference's host code (oracle/_ref, built from /root/reference) on the restated ggml CPU ops; ggml itself is not on this machine" % len(nfe)},
        "e2e": {"value": img_s, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sdxl", action="store_true", help="skip the SDXL 1024x1024 side measurement (second half of the BASELINE metric)")
    ap.add_argument("--sdxl-batch", type=int, default=2, help="SDXL images per GPU per generation")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        weights_path("sd1")
    if world > 1:
        dist.barrier()
    model = weights_path("sd1")

    from mlimgsynth_b200 import api, dist as D
    import ctypes as C
    B = args.batch
    os.environ.setdefault("GGML_B200_QUIET", "1")
    ctx = api.Ctx(backend="B200:%d" % local, model=model, image_dim=(512, 512), steps=N_STEPS_SAMPLER, method="euler",
                  cfg_scale=7, batch_size=B)
    eng = C.CDLL(os.path.join(ROOT, "mlimgsynth_b200", "lib", "libggml_b200.so"), mode=C.RTLD_LOCAL)   # same library instance as the host lib uses
    eng.ggml_b200_timer_stop.restype = C.c_double

    class St(C.Structure):
        _fields_ = [(k, C.c_uint64) for k in ("kernel_launches", "graph_launches", "plans_built", "h2d_bytes", "d2h_bytes")]
    eng.ggml_b200_get_stats.restype = C.POINTER(St)

    def stats():
        s = eng.ggml_b200_get_stats().contents
        return {k: getattr(s, k) for k, _ in St._fields_}

    def gen(seed, cached_cond):
        # image i of the global batch gets seed + i whatever the world size; this rank owns a contiguous block
        ctx.set("seed", D.image_seeds(seed * 1000, B * world, rank, world)[0])
        ctx.set("prompt", PROMPT)
        if cached_cond:
            ctx.set("tensor_use_flags", api.TUF_CONDITIONING)
        ctx.generate()
        return [ctx.image(i) for i in range(B)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        return D.all_max(x, device="cuda") if world > 1 else x

    # ---- warm-up (builds graphs, uploads weights, captures CUDA graphs)
    for i in range(max(args.warmup, 3)):
        imgs = gen(i, cached_cond=(i > 0))
        if world > 1:
            D.gather_arrays(np.stack(imgs), dst=0, device="cuda")     # also warms up the NCCL point-to-point channels

    # ---- `value`: device time of K generations, conditioning resident
    clocks = ClockSampler(local); clocks.start()
    barrier()
    s0 = stats()
    eng.ggml_b200_timer_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        gen(100 + i, cached_cond=True)
    dev_ms = eng.ggml_b200_timer_stop()
    barrier()
    wall_cached = time.perf_counter() - t0
    s1 = stats()
    dev_s = allmax(dev_ms / 1e3)

    # ---- `e2e`: public API, host buffers in and out, + gather of the RGB8 images to rank 0
    barrier()
    e0 = stats()
    t0 = time.perf_counter()
    for i in range(args.steps):
        imgs = gen(200 + i, cached_cond=False)
        if world > 1:
            D.gather_arrays(np.stack(imgs), dst=0, device="cuda")     # NCCL: only the final RGB8 images travel
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    e1 = stats()
    clk = clocks.summary()

    total_images = B * world * args.steps
    value = total_images / dev_s
    e2e = total_images / e2e_s
    nfe_per_gen = 2 * N_STEPS_SAMPLER * B          # UNet evaluations (latents) per generation per GPU
    unet_it_s = (N_STEPS_SAMPLER * B * world * args.steps) / dev_s   # sampler steps x images per second

    # ---- roofline of the dominant kernel family + per-NFE time (profiled eager pass, outside timed regions)
    roof, prof = None, {}
    if rank == 0:
        pk = peaks()
        x = np.random.default_rng(0).standard_normal((2 * B, 4, 64, 64)).astype(np.float32)
        cond = (np.random.default_rng(1).standard_normal((2 * B, 77, 768)) * 0.5).astype(np.float32)
        ctx.unet_eval(x, cond, None, 5.0)      # builds/captures the batch-2B graph used for NFE timing
        ctx.unet_eval(x, cond, None, 5.0)
        eng.ggml_b200_timer_start()
        for _ in range(5):
            ctx.unet_eval(x, cond, None, 5.0)
        nfe_ms = eng.ggml_b200_timer_stop() / 5
        eng.ggml_b200_profile_enable(1)
        ctx.unet_eval(x, cond, None, 5.0)
        names = {0: "copy", 7: "gemm_simt", 8: "groupnorm", 9: "layernorm", 10: "geglu", 11: "im2col", 13: "attention", 14: "gemm_tc", 15: "conv3x3_tc", 3: "upscale"}
        for k, nm in names.items():
            ms, fl, by, ln = C.c_double(), C.c_double(), C.c_double(), C.c_uint64()
            eng.ggml_b200_profile_get(k, C.byref(ms), C.byref(fl), C.byref(by), C.byref(ln))
            if ln.value:
                prof[nm] = {"ms": ms.value, "tflop": fl.value / 1e12, "gb": by.value / 1e9, "launches": ln.value}
        eng.ggml_b200_profile_enable(0)
        g_ms = prof.get("gemm_tc", {}).get("ms", 0) + prof.get("conv3x3_tc", {}).get("ms", 0)
        g_fl = prof.get("gemm_tc", {}).get("tflop", 0) + prof.get("conv3x3_tc", {}).get("tflop", 0)
        g_n = prof.get("gemm_tc", {}).get("launches", 0) + prof.get("conv3x3_tc", {}).get("launches", 0)
        achieved = g_fl / (g_ms / 1e3) if g_ms else 0.0
        traffic, traffic_launches = None, 0      # DRAM bytes per launch of the same kernel family from the committed ncu capture (profiles/)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")))
            traffic = tj["dram_bytes_per_launch"]; traffic_launches = tj["launches"]
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "gemm_tc_persistent_kernel (tcgen05 cta_group::2 GEMM + implicit 3x3 conv; all linear / conv launches of one UNet evaluation)", "achieved": achieved, "peak": pk["bf16_tflops"],
                "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"], "traffic": traffic,
                "traffic_note": "dram__bytes_read+write per launch, ncu capture of all %d tcgen05 GEMM/conv launches of one evaluation (profiles/r1_ncu_gemm_traffic.md); algorithmic bytes per launch %.1f MB" % (traffic_launches, (prof.get("gemm_tc", {}).get("gb", 0) + prof.get("conv3x3_tc", {}).get("gb", 0)) * 1e3 / max(g_n, 1)),
                "algorithmic_tflop_per_launch": g_fl / max(g_n, 1), "peak_source": pk["src"],
                "launches_per_unet_eval": g_n, "avg_launch_us": g_ms * 1e3 / g_n if g_n else None,
                "unet_eval_ms_batch%d" % (2 * B): nfe_ms,
                "unet_tflops": FLOP_PER_NFE_SD15_512 * 2 * B / (nfe_ms / 1e3) / 1e12,
                "unet_frac_of_peak": FLOP_PER_NFE_SD15_512 * 2 * B / (nfe_ms / 1e3) / 1e12 / pk["bf16_tflops"]}

    # ---- second half of the BASELINE metric: SDXL base 1024x1024, 30 Euler steps, cfg 7, full VAE decode
    # (BASELINE configs[2] shape at --sdxl-batch images per GPU). Reported under config.sdxl_1024; the headline stays SD1.5.
    sdxl = None
    if not args.no_sdxl:
        ctx.close(); ctx = None
        if rank == 0:
            weights_path("sdxl")
        barrier()
        xb = args.sdxl_batch
        cx = api.Ctx(backend="B200:%d" % local, model=weights_path("sdxl"), image_dim=(1024, 1024), steps=SDXL_STEPS, method="euler", cfg_scale=7, batch_size=xb)
        def genx(seed, cached):
            cx.set("seed", D.image_seeds(seed * 1000, xb * world, rank, world)[0]); cx.set("prompt", PROMPT)
            if cached:
                cx.set("tensor_use_flags", api.TUF_CONDITIONING)
            cx.generate()
            return [cx.image(i) for i in range(xb)]
        genx(0, False); genx(1, True)
        barrier()
        eng.ggml_b200_timer_start()
        genx(2, True)
        x_dev = allmax(eng.ggml_b200_timer_stop() / 1e3)
        barrier()
        t0 = time.perf_counter()
        imgs = genx(3, False)
        if world > 1:
            D.gather_arrays(np.stack(imgs), dst=0, device="cuda")
        barrier()
        x_e2e = allmax(time.perf_counter() - t0)
        if rank == 0:
            lat = np.random.default_rng(0).standard_normal((2 * xb, 4, 128, 128)).astype(np.float32)
            cond = (np.random.default_rng(1).standard_normal((2 * xb, 77, 2048)) * 0.5).astype(np.float32)
            lab = (np.random.default_rng(2).standard_normal((2 * xb, 2816)) * 0.5).astype(np.float32)
            cx.unet_eval(lat, cond, lab, 5.0); cx.unet_eval(lat, cond, lab, 5.0)
            eng.ggml_b200_timer_start()
            for _ in range(3):
                cx.unet_eval(lat, cond, lab, 5.0)
            x_nfe = eng.ggml_b200_timer_stop() / 3
            tf = FLOP_PER_NFE_SDXL_1024 * 2 * xb / (x_nfe / 1e3) / 1e12
            sdxl = {"workload": "SDXL base txt2img 1024x1024, %d Euler steps, cfg 7, full VAE decode, random-init weights" % SDXL_STEPS,
                    "batch_per_gpu": xb, "images_per_sec": xb * world / x_dev, "e2e_images_per_sec": xb * world / x_e2e,
                    "unet_it_per_s": SDXL_STEPS * xb * world / x_dev, "ms_per_generation": x_dev * 1e3,
                    "unet_eval_ms_batch%d" % (2 * xb): x_nfe, "unet_tflops": tf, "unet_frac_of_peak": tf / peaks()["bf16_tflops"]}
        cx.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_reference_run(model, 2, os.cpu_count())
        if res:
            t_nfe = sum(res["nfe_s"][1:] or res["nfe_s"]) / len(res["nfe_s"][1:] or res["nfe_s"])
            vae = res["vae_s"] or 0.0
            cpu = {"value": 1.0 / (2 * N_STEPS_SAMPLER * t_nfe + vae), "unit": "images/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": "2 UNet evaluations + 1 VAE decode at 512x512 by the reference host code on the restated ggml CPU ops (oracle/_ref); "
                             "extrapolated to 40 NFE + decode; s/NFE %.2f, VAE %.2f s" % (t_nfe, vae)}

    if rank == 0:
        line = {
            "metric": "images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_s * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d (images), cfg halves batched" % world,
                       "unet_it_per_s": unet_it_s, "nfe_per_s": 2 * unet_it_s,
                       "l2": "working set (weights 1.7 GB + activations) exceeds the 126 MB L2: every UNet evaluation streams all weights from HBM",
                       "sdxl_1024": sdxl},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": (e1["h2d_bytes"] - e0["h2d_bytes"]) // args.steps,
                    "d2h_bytes_per_step": (e1["d2h_bytes"] - e0["d2h_bytes"]) // args.steps, "ms_per_step": e2e_s * 1e3 / args.steps},
            "gpu_launches": s1["kernel_launches"] - s0["kernel_launches"], "graph_replays": s1["graph_launches"] - s0["graph_launches"],
            "wall_ms_per_step_cached_cond": wall_cached * 1e3 / args.steps,
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "kernel_profile": prof,
        }
        print(json.dumps(line))
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
