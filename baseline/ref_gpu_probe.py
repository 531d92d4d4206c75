#!/usr/bin/env python
"""Baseline measurement only (no product code): time the *unmodified* reference
nn.Modules under PyTorch eager on whatever GPU this runs on, and on the host CPU.

Usage (on the GPU box):  python baseline/ref_gpu_probe.py
Writes gpurun_out/ref_gpu_probe.json and prints a summary.

The reference modules are imported from /root/reference if present, else from
baseline/_ref (a verbatim, git-ignored copy of the reference's models/ tree).
"""
import importlib, json, os, sys, time, copy, math, platform

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference" if os.path.isdir("/root/reference/models") else os.path.join(HERE, "_ref")
sys.path.insert(0, os.path.dirname(HERE))      # repository root: baseline.refload

import torch

OUT = {"ref_path": REF, "torch": torch.__version__, "cuda": torch.version.cuda,
       "host_cpus": os.cpu_count(), "torch_threads": torch.get_num_threads(),
       "platform": platform.platform(), "cases": [], "notes": []}


def model(name, seed=0):
    torch.manual_seed(seed)
    from baseline.refload import reference_model_class
    return reference_model_class(os.path.join(REF, "models"), name)().eval()


def psnr(a, b):
    mse = ((a.float() - b.float()) ** 2).mean().item()
    return 10 * math.log10(1.0 / mse) if mse > 0 else float("inf")


def time_cuda(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / 1e3)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def run_case(tag, name, shape, kw, modes, flop_per_frame=None):
    B = shape[0]
    base = model(name)
    torch.manual_seed(123)
    x = torch.rand(*shape)
    rec = {"tag": tag, "model": name, "shape": list(shape), "kw": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
           "gflop_per_frame": flop_per_frame, "modes": {}}
    y_ref = None
    for mode in modes:
        try:
            torch.backends.cudnn.allow_tf32 = mode != "fp32_notf32"
            torch.backends.cuda.matmul.allow_tf32 = mode == "fp32_tf32all"
            torch.backends.cudnn.benchmark = True
            m = copy.deepcopy(base).cuda()
            xin = x.cuda()
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode.startswith("autocast_bf16") else torch.autocast("cuda", enabled=False)
            if mode.startswith("pure_bf16"):
                m = m.bfloat16(); xin = xin.bfloat16()
            if mode.endswith("_cl"):
                m = m.to(memory_format=torch.channels_last); xin = xin.contiguous(memory_format=torch.channels_last)
            def fn():
                with torch.no_grad(), ctx:
                    return m(xin, **kw)
            torch.cuda.reset_peak_memory_stats()
            y = fn().float().cpu()
            best, med = time_cuda(fn)
            r = {"best_s": best, "median_s": med, "fps": B / best, "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30,
                 "out_shape": list(y.shape)}
            if flop_per_frame:
                r["model_TFLOPs"] = flop_per_frame * B / best / 1e3
            if mode == "fp32_notf32":
                y_ref = y
            elif y_ref is not None:
                r["maxabs_vs_fp32_notf32"] = (y - y_ref).abs().max().item()
                r["psnr_vs_fp32_notf32"] = psnr(y, y_ref)
            rec["modes"][mode] = r
            print(f"[{tag}] {mode:18s} best {best*1e3:9.2f} ms  {B/best:9.2f} fps  mem {r['peak_mem_GB']:.2f} GB "
                  + (f" maxabs {r.get('maxabs_vs_fp32_notf32', 0):.2e}" if 'maxabs_vs_fp32_notf32' in r else ""), flush=True)
            del m, xin, y
            torch.cuda.empty_cache()
        except Exception as e:  # OOM etc.
            rec["modes"][mode] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
            print(f"[{tag}] {mode}: ERROR {type(e).__name__}: {str(e)[:160]}", flush=True)
            torch.cuda.empty_cache()
    rec["_y_ref"] = y_ref
    OUT["cases"].append(rec)
    return rec


def main():
    has_gpu = torch.cuda.is_available()
    OUT["gpu"] = torch.cuda.get_device_name(0) if has_gpu else None
    OUT["n_gpu"] = torch.cuda.device_count() if has_gpu else 0
    if has_gpu:
        OUT["cudnn"] = torch.backends.cudnn.version()
    print(json.dumps({k: v for k, v in OUT.items() if k not in ("cases",)}), flush=True)

    # ---- CPU baseline, config 1 (FastTransformer x2 of one 360x640 frame, fp32) + Window 720p->1080p
    cpu = {}
    for tag, name, shape, kw in [("cfg1_fast_360x640_x2", "FastTransformer", (1, 3, 360, 640), {"upscale_factor": 2}),
                                 ("window_720p_1080p_b1", "WindowTransformer", (1, 3, 720, 1280), {})]:
        m = model(name); torch.manual_seed(123); x = torch.rand(*shape)
        with torch.no_grad():
            y = m(x, **kw)
            ts = []
            for _ in range(3):
                t = time.perf_counter(); m(x, **kw); ts.append(time.perf_counter() - t)
        cpu[tag] = {"best_s": min(ts), "fps": shape[0] / min(ts), "threads": torch.get_num_threads(), "y": y}
        print(f"[cpu] {tag}: best {min(ts):.3f} s  ({shape[0]/min(ts):.3f} fps) on {torch.get_num_threads()} threads", flush=True)
    OUT["cpu"] = {k: {kk: vv for kk, vv in v.items() if kk != "y"} for k, v in cpu.items()}

    if not has_gpu:
        finish(); return

    modes = ["fp32_notf32", "fp32_default", "autocast_bf16", "autocast_bf16_cl", "pure_bf16", "pure_bf16_cl"]
    # config 1 on GPU too (oracle noise CPU-fp32 vs CUDA-fp32)
    r = run_case("cfg1_fast_360x640_x2", "FastTransformer", (1, 3, 360, 640), {"upscale_factor": 2}, modes, 139.84)
    if r.get("_y_ref") is not None:
        OUT["cpu_vs_cuda_fp32_maxabs_cfg1"] = (r["_y_ref"] - cpu["cfg1_fast_360x640_x2"]["y"]).abs().max().item()
    r = run_case("window_720p_1080p_b1", "WindowTransformer", (1, 3, 720, 1280), {}, modes, 126.94)
    if r.get("_y_ref") is not None:
        OUT["cpu_vs_cuda_fp32_maxabs_window_b1"] = (r["_y_ref"] - cpu["window_720p_1080p_b1"]["y"]).abs().max().item()
    print("cpu-vs-cuda fp32 maxabs:", OUT.get("cpu_vs_cuda_fp32_maxabs_cfg1"), OUT.get("cpu_vs_cuda_fp32_maxabs_window_b1"), flush=True)
    # config 2
    run_case("cfg2_window_720p_1080p_b8", "WindowTransformer", (8, 3, 720, 1280), {}, modes, 126.94)
    # config 4 sweep (batch 1)
    for s, gf in [(2, 559.36), (3, 916.51), (4, 1688.92), (6, 2845.16)]:
        run_case(f"cfg4_fast_720p_x{s}_b1", "FastTransformer", (1, 3, 720, 1280), {"upscale_factor": s},
                 ["fp32_notf32", "autocast_bf16", "pure_bf16_cl"], gf)
    # config 5 as the reference can actually run it (720p input only)
    run_case("cfg5_residual_720p_4k_b2", "ResidualTransformer", (2, 3, 720, 1280), {"res_out": (2160, 3840)},
             ["fp32_notf32", "autocast_bf16", "pure_bf16"], 179.45)
    run_case("cfg5alt_fast_1080p_x2_b1", "FastTransformer", (1, 3, 1080, 1920), {"upscale_factor": 2},
             ["autocast_bf16", "pure_bf16_cl"], 1247.79)

    # ---- where does the time go on the GPU today? (config 2, autocast bf16)
    try:
        from torch.profiler import profile, ProfilerActivity
        m = model("WindowTransformer").cuda(); x = torch.rand(8, 3, 720, 1280, device="cuda")
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            m(x); torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                m(x); torch.cuda.synchronize()
        tab = prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=70)
        print(tab, flush=True)
        OUT["profile_cfg2_autocast_bf16"] = tab
        m = model("FastTransformer").cuda(); x = torch.rand(1, 3, 720, 1280, device="cuda")
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            m(x, upscale_factor=2); torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                m(x, upscale_factor=2); torch.cuda.synchronize()
        tab = prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=70)
        print(tab, flush=True)
        OUT["profile_fast_720p_x2_autocast_bf16"] = tab
    except Exception as e:
        OUT["notes"].append(f"profile failed: {type(e).__name__}: {e}")
    finish()


def finish():
    for c in OUT["cases"]:
        c.pop("_y_ref", None)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ref_gpu_probe.json", "w") as f:
        json.dump(OUT, f, indent=1)
    print("wrote gpurun_out/ref_gpu_probe.json")


if __name__ == "__main__":
    main()
