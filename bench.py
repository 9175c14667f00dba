#!/usr/bin/env python
"""Benchmark of the per-ray rendering hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # the reference's own PyTorch path on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # one rank per GPU

A "step" = one pass of CrossAttentionRenderer.forward(input, z=z) over one batch of synthetic
scenes: 256x256 target rays, 2 source views, 64 epipolar samples, 12 scenes per GPU (BASELINE
config 2; at N>1 every rank renders its own 12 scenes = config 3's layout, weak scaling, tiles
all-gathered at the end of the step).  Prints ONE JSON line.  After the headline region the same
line gets `extra_configs` (config 3's bf16 arithmetic at this N; config 4 at N=1), each measured the
same way with its own roofline and parity; `general_branches` (N=1: the reference's other forward branches,
n_view 1 / 3 and no_latent_concat, one scene each, with a parity check); at N>1 a strong-scaling block; and
`reference_gpu`: the UNMODIFIED reference (oracle/_ref) timed on the same B200 (N=1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cross_attention_renderer_b200 import synthetic  # noqa: E402

METRIC = "rendered rays/sec at 256x256, 64 epipolar samples, 2 source views"
_CPU_THREADS = None
FLOP_PER_SAMPLE_VIEW_ENC1 = 2 * 579 * 576


def tap_bytes_per_ray(P, elt):
    """n * P * 2 gathers * 576 channels * 4 taps * sizeof(element) per ray."""
    return 2 * P * 2 * 576 * 4 * elt


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("CAR_CLOCK_SAMPLE_MS", "200"), "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw)}


# ------------------------------------------------------------------------------------------------
# the reference's own implementation on the host cores (oracle/_ref), else the oracle port
# ------------------------------------------------------------------------------------------------
def _pick_threads(run):
    """"All the host threads it can use": torch's intra-op pool stops scaling (and can collapse) well before
    128 threads on these elementwise-heavy ops, so pick the thread count that is fastest on a small probe
    and report it as `cores`."""
    global _CPU_THREADS
    if _CPU_THREADS is None:
        avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        best = None
        for nt in sorted({n for n in (8, 16, 32, 64, avail) if n <= avail}):
            torch.set_num_threads(nt)
            run(64, 200)
            dt = run(128, 201)
            if best is None or dt < best[0]:
                best = (dt, nt)
        _CPU_THREADS = best[1]
    torch.set_num_threads(_CPU_THREADS)
    return _CPU_THREADS


class CpuArm:
    """Times forward(input, z=z) of the reference on the host for `rays` rays of one HxH / P-sample scene.
    kind "reference": the unmodified reference files in oracle/_ref (Tensor.cuda made a no-op);
    kind "port": oracle/car_oracle.py, only when oracle/_ref was not built."""

    def __init__(self, H, P):
        from oracle import ref_loader
        self.H, self.P = H, P
        self.z = synthetic.make_features(1, H, seed=0)
        self.sd = synthetic.make_state_dict(seed=0)
        self.kind = "reference" if ref_loader.available() else "port"
        self.rl = ref_loader
        self.model = None
        if self.kind == "reference":
            with ref_loader.host_mode():
                self.model = ref_loader.build_model(self.sd, H, P, device="cpu")
        _pick_threads(self.run)

    def run(self, rays, seed):
        inp = synthetic.make_inputs(1, self.H, self.H, seed=seed, rays=rays)
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.kind == "reference":
                with self.rl.host_mode():
                    out = self.rl.render(self.model, inp, self.z, chunk_rays=8192)   # render_realestate10k_traj.py:96
            else:
                from oracle import car_oracle as orc
                out = orc.render(self.sd, inp, self.z, self.H, self.H, self.P)
        float(out["rgb"].sum())
        return time.perf_counter() - t0


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    H, P = args.size, args.samples
    rays = args.cpu_rays
    arm = CpuArm(H, P)
    for _ in range(args.warmup):
        arm.run(max(64, rays // 8), 300)
    t_tot, n_tot = 0.0, 0
    for i in range(args.steps):
        t_tot += arm.run(rays, i)
        n_tot += rays
    value = n_tot / t_tot
    cores = torch.get_num_threads()
    sample = f"{rays} of the {H * H} target rays of one scene per step, {H}x{H} maps, {P} samples, 2 views"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{H}x{H} target, 2 views, {P} samples (BASELINE config 2), bounded sample on host cores",
                   "sample": sample,
                   "what": "unmodified reference files (oracle/_ref) through CrossAttentionRenderer.forward(input, z=z)"
                           if arm.kind == "reference" else "oracle port of the reference (oracle/_ref not built)"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# ncu profile lookup keyed on the build being timed
# ------------------------------------------------------------------------------------------------
def _profile_numbers(kind, precision, build):
    """(dram bytes per launch, lts bytes per launch, file) from profiles/r02_ncu_<kind>_<precision>.json if that
    capture is of THIS build (`build_id` field), else (None, None, reason)."""
    path = os.path.join(ROOT, "profiles", f"r02_ncu_{kind}_{precision}.json")
    if not os.path.exists(path):
        return None, None, "no ncu capture of this build under profiles/"
    try:
        pj = json.load(open(path))
        recs = pj["launches"] if isinstance(pj, dict) and "launches" in pj else (pj if isinstance(pj, list) else [pj])
        bid = pj.get("build_id") if isinstance(pj, dict) else None
        if bid != build:
            return None, None, f"profiles/{os.path.basename(path)} is of build {bid}, timed build is {build}"
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}

        def val(rec, key):
            v, u = rec[key].split()[:2]
            return float(v) * mult[u]
        dram = sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in recs) / len(recs)
        cyc = [float(r["sm__cycles_elapsed.max"].split()[0]) for r in recs if "sm__cycles_elapsed.max" in r]
        lts = None
        if all("lts__t_bytes.sum" in r for r in recs):
            lts = sum(val(r, "lts__t_bytes.sum") for r in recs) / len(recs)
        elif all("lts__t_sectors.sum" in r for r in recs):                       # 32-byte sectors
            lts = sum(float(r["lts__t_sectors.sum"].split()[0]) * 32.0 for r in recs) / len(recs)
        extra = {"file": os.path.basename(path), "launches_averaged": len(recs)}
        if lts is not None and cyc:
            extra["lts_bytes_per_clk"] = lts / (sum(cyc) / len(cyc))
            # /opt/skills/guides/B300_MICROARCH.md, "L2 cache": LTS throughput cap ~6300 B/cyc full chip (LDG and TMA alike)
            extra["lts_cap_bytes_per_clk_guide"] = 6300.0
        tp = [r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") for r in recs]
        if all(tp):
            extra["tensor_pipe_active_pct"] = sum(float(t.split()[0]) for t in tp) / len(tp)
        return dram, lts, extra
    except Exception as exc:  # pragma: no cover
        return None, None, f"unreadable profile: {exc!r}"[:120]


# ------------------------------------------------------------------------------------------------
# one measured configuration
# ------------------------------------------------------------------------------------------------
def measure(ctx, precision, b, H, P, steps, warmup, e2e_steps, with_clocks, seed_base=100):
    """Device-timed throughput of `steps` steps + e2e with host buffers + per-stage roofline + parity of what
    was timed.  Returns a dict; every rank calls this (the step contains collectives at N>1)."""
    import ctypes as C
    import torch.distributed as dist
    from cross_attention_renderer_b200 import _lib
    from cross_attention_renderer_b200.models import CrossAttentionRenderer
    world, rank, local, dev, lib = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"], ctx["lib"]
    R = H * H
    inp_h = synthetic.make_inputs(b, H, H, seed=seed_base + rank)
    z_h = synthetic.make_features(b, H, seed=seed_base + rank)
    sd = synthetic.make_state_dict(seed=0)
    pin = lambda t: t.pin_memory()
    inp_h = {k: {kk: pin(vv) for kk, vv in v.items()} for k, v in inp_h.items()}
    z_h = [pin(t) for t in z_h]
    model = CrossAttentionRenderer(n_view=2, npoints=P, precision=precision).to(dev).eval()
    model.load_state_dict(sd, strict=False)
    model.H = model.W = H
    if ctx["chunk_rays"]:
        model.chunk_rays = ctx["chunk_rays"]
    model.pixel_val_to_cpu = False          # metric excludes the optional pixel_val D2H
    inp_d = synthetic.to_device(inp_h, dev)
    z_d = [t.to(dev) for t in z_h]
    total_rays = b * R * world

    def step_resident():
        model.release_features()            # re-pack NCHW->NHWC every step (no cached work)
        with torch.no_grad():
            out = model(inp_d, z=z_d)
        if world > 1:                       # final gather of rendered tiles (north_star)
            for k in ("rgb", "valid_mask", "depth_ray"):
                t = out[k].reshape(b * R, -1)
                recv = torch.empty(world * t.shape[0], t.shape[1], device=dev, dtype=t.dtype)
                dist.all_gather_into_tensor(recv, t.contiguous())
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_w = time.perf_counter()
    for _ in range(warmup):
        step_resident()
    sync_all()
    # a fresh box's first seconds run slow (clocks, page-in): warm up for at least ~3 s in total.
    # The number of extra steps is agreed across ranks (the step contains a collective).
    t_el = time.perf_counter() - t_w
    extra_warm = 0 if t_el >= 3.0 else min(6, int((3.0 - t_el) / max(1e-3, t_el / warmup)) + 1)
    if world > 1:
        ew = torch.tensor([extra_warm], device=dev)
        dist.all_reduce(ew, op=dist.ReduceOp.MAX)
        extra_warm = int(ew.item())
    for _ in range(extra_warm):
        step_resident()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0 and with_clocks and not os.environ.get("CAR_NO_CLOCK_SAMPLER"):
        sampler.start()
    nst = len(_lib.STAGES)
    ms_arr, ln_arr = (C.c_float * nst)(), (C.c_int * nst)()
    if ctx["stage_profile"]:
        lib.car_profile_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    step_evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev0.record()
    for i_ in range(steps):
        step_resident()
        step_evs[i_].record()
    ev1.record()
    sync_all()
    ms_total = ev0.elapsed_time(ev1)
    step_ms = [round(([ev0] + step_evs)[i_].elapsed_time(step_evs[i_]), 2) for i_ in range(steps)]
    if ctx["stage_profile"]:
        lib.car_profile_end(ms_arr, ln_arr, nst)
    clocks = sampler.stop() if (rank == 0 and with_clocks) else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    launches_per_step = model.last_launch_count + 3
    value = total_rays * steps / (ms_total * 1e-3)

    # ---- end to end through the public API with HOST buffers ------------------------------
    e2e = None
    if e2e_steps > 0:
        h2d = sum(t_.numel() * t_.element_size() for t_ in z_h) + \
            sum(v.numel() * v.element_size() for d_ in inp_h.values() for v in d_.values())
        d2h = (b * R * 3 + b * R + b * R) * 4

        def step_e2e():
            inp_g = synthetic.to_device(inp_h, dev)      # H2D from pinned memory
            z_g = [t_.to(dev, non_blocking=True) for t_ in z_h]
            with torch.no_grad():
                o = model(inp_g, z=z_g)                   # fresh device tensors: features are re-packed
            return o["rgb"].cpu(), o["valid_mask"].cpu(), o["depth_ray"].cpu()   # D2H of the result
        # (uploading step k+1 on a copy stream while step k renders was measured and is slower: 694 vs 680 ms per step -
        # the 931 MB of DMA writes compete with the L2-bound kernels)
        step_e2e()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        sync_all()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": total_rays * e2e_steps / float(tt.item()), "unit": "rays/s",
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": e2e_steps}
    res = {"value": value, "ms_per_step": ms_total / steps, "step_ms": step_ms, "clocks": clocks, "e2e": e2e,
           "gpu_launches": launches_per_step * steps, "extra_warm": extra_warm, "total_rays": total_rays,
           "feature_mb": sum(t_.numel() * 4 for t_ in z_h) / 1e6}
    if rank != 0:
        return res

    # ---- roofline of the dominant kernel + the other stages -------------------------------
    peaks = measured_peaks()
    build = _lib.build_id()
    stage_ms = {n: float(ms_arr[i]) for i, n in enumerate(_lib.STAGES)}
    stage_ln = {n: int(ln_arr[i]) for i, n in enumerate(_lib.STAGES)}
    rays_prof = b * R * steps                            # rank-0 rays covered by the profile
    feat_elt = 2 if precision == "bf16" else 4
    tap_bytes = tap_bytes_per_ray(P, feat_elt)
    roof = {}
    if stage_ms["gather"] > 0:
        ach = rays_prof * tap_bytes / (stage_ms["gather"] * 1e-3) / 1e9
        roof["gather"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": ach / peaks["hbm_gbs"], "traffic": None, "kernel": "k_gather",
                          "avg_launch_ms": stage_ms["gather"] / max(1, stage_ln["gather"]),
                          "algorithmic_bytes_per_ray": tap_bytes, "peak_source": peaks["source"]}
    if stage_ms["gemm_enc1"] > 0:
        fl = rays_prof * 2 * P * 2 * FLOP_PER_SAMPLE_VIEW_ENC1
        ach = fl / (stage_ms["gemm_enc1"] * 1e-3) / 1e12
        roof["gemm_enc1"] = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"],
                             "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None,
                             "kernel": "k_gemm_simt (fp32 FFMA)" if precision == "fp32_simt" else "k_gemm_umma",
                             "avg_launch_ms": stage_ms["gemm_enc1"] / max(1, stage_ln["gemm_enc1"]),
                             "peak_source": peaks["source"] + " bf16 dense, sustained"}
    if stage_ms.get("fused", 0) > 0:
        # k_fused_encode: gather + encoder GEMMs of one ray chunk per launch.  Algorithmic bytes =
        # bilinear tap bytes (n*P*2 gathers*576 ch*4 taps*elt per ray); algorithmic
        # flops = 2*128*(592*576 + 576*416)*2 per ray (the 3xbf16 split executes 3x that).
        t_ = stage_ms["fused"] * 1e-3
        ach = rays_prof * tap_bytes / t_ / 1e9
        flop_ray = 2 * 128 * (592 * 576 + 576 * 416) * 2 * (P / 64.0)
        mma_mult = 3 if precision == "fp32" else 1
        dram, lts, src = _profile_numbers("fused", precision, build)
        rays_launch = rays_prof / max(1, stage_ln["fused"])
        roof["fused"] = {
            "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
            "traffic": dram, "lts_bytes": lts, "traffic_source": src, "kernel": "k_fused_encode",
            "avg_launch_ms": stage_ms["fused"] / max(1, stage_ln["fused"]), "rays_per_launch": rays_launch,
            "algorithmic_bytes_per_ray": tap_bytes, "algorithmic_bytes_per_launch": tap_bytes * rays_launch,
            "peak_source": peaks["source"],
            "note": "taps are served by L1/L2 (feature maps of a scene fit the 126 MB L2), hence DRAM traffic << tap bytes",
            "tensor": {"bound": "tensor", "algorithmic_tflops": rays_prof * flop_ray / t_ / 1e12,
                       "mma_tflops_executed": rays_prof * flop_ray * mma_mult / t_ / 1e12,
                       "peak_bf16_tflops_sustained": peaks["bf16_tflops_sustained"],
                       "frac_executed": rays_prof * flop_ray * mma_mult / t_ / 1e12 / peaks["bf16_tflops_sustained"],
                       "frac_algorithmic": rays_prof * flop_ray / t_ / 1e12 / peaks["bf16_tflops_sustained"]}}
    if stage_ms.get("attention", 0) > 0 and P in (64, 128) and precision != "fp32_simt":
        # per-ray attention tail: algorithmic HBM bytes per ray from ctx["tail_bytes"] (follows the kernel structure)
        rows = 2 * P
        kh = rows * 128 * 2 * (2 if precision == "fp32" else 1)
        v = rows * 288 * 4
        bytes_ray = ctx["tail_bytes"](rows, kh, v)
        t_ = stage_ms["attention"] * 1e-3
        ach = rays_prof * bytes_ray / t_ / 1e9
        dram, lts, src = _profile_numbers("tail", precision, build)
        roof["tail"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / peaks["hbm_gbs"], "traffic": dram, "lts_bytes": lts, "traffic_source": src,
                        "kernel": "k_tail", "avg_launch_ms": stage_ms["attention"] / max(1, stage_ln["attention"]),
                        "algorithmic_bytes_per_ray": bytes_ray, "peak_source": peaks["source"]}
    dom = max(stage_ms, key=stage_ms.get)
    res["roofline"] = roof.get(dom) or roof.get("fused") or roof.get("gemm_enc1") or roof.get("gather")
    res["roofline_all"] = roof
    res["stage_share"] = {k: round(v / max(1e-9, sum(stage_ms.values())), 4) for k, v in stage_ms.items() if v > 0}
    res["stage_ms_per_step"] = {k: round(v / steps, 3) for k, v in stage_ms.items() if v > 0}
    res["build_id"] = build

    # ---- parity of what was just timed: 512 rays of scene 0 against the oracle (untimed) ----
    try:
        from oracle import car_oracle as orc
        idx = torch.randperm(R, generator=torch.Generator().manual_seed(0))[:512].sort().values
        one = lambda t: t[:1].cpu()
        inp1 = {"context": {k: one(v) for k, v in inp_h["context"].items()},
                "query": {k: one(v) for k, v in inp_h["query"].items()}}
        inp1["query"]["uv"] = inp1["query"]["uv"][:, :, idx]
        z1 = [t[:2].cpu() for t in z_h]
        cams = orc.prepare_cameras(inp1)
        iv = torch.linspace(0, 1, P)
        torch.set_num_threads(min(32, os.cpu_count()))
        with torch.no_grad():
            ref = orc.render({k: v.cpu() for k, v in sd.items()}, inp1, z1, H, H, P, interval=iv, cams=cams)
            model.release_features()
            got = model.render_prepared({k: v.to(dev).contiguous() for k, v in cams.items()},
                                        inp1["query"]["uv"][:, 0].contiguous().to(dev), iv.to(dev),
                                        [t.to(dev) for t in z1], 1, idx.numel())
        rgb = got["rgb"].cpu()
        g = torch.Generator().manual_seed(5)
        target = ref["rgb"] + 0.18 * torch.randn(ref["rgb"].shape, generator=g)      # ~21 dB from the render
        res["parity"] = {"rays": int(idx.numel()),
                         "rgb_rel_err_vs_oracle": float((rgb - ref["rgb"]).abs().max() / ref["rgb"].abs().max()),
                         "psnr_db_vs_oracle": orc.psnr(rgb, ref["rgb"]),
                         "delta_psnr_db_vs_noisy_target": orc.psnr(rgb, target) - orc.psnr(ref["rgb"], target),
                         "valid_mask_equal": bool(torch.equal(got["valid_mask"].cpu(), ref["valid_mask"])),
                         "pixel_val_bit_equal": bool(torch.equal(got["pixel_val"].cpu(), ref["pixel_val"]))}
    except Exception as exc:  # pragma: no cover - parity is reported, never fatal for the bench line
        res["parity"] = {"error": repr(exc)[:200]}
    return res


def strong_scaling(ctx, scenes_total, H, P, reps):
    """`scenes_total` scenes, ONE at a time: features on rank 0 only, NCCL broadcast (one scene ahead, overlapped
    with the previous scene's rendering), each rank renders 1/N of the scene's rays, tiles all-gathered.  Device
    time, max over ranks; the same with the broadcast disabled (every rank already holds the maps) isolates its cost."""
    import torch.distributed as dist
    from cross_attention_renderer_b200 import sharding
    from cross_attention_renderer_b200.models import CrossAttentionRenderer
    world, rank, dev = ctx["world"], ctx["rank"], ctx["dev"]
    model = CrossAttentionRenderer(n_view=2, npoints=P, precision="fp32").to(dev).eval()
    model.load_state_dict(synthetic.make_state_dict(seed=0), strict=False)
    model.H = model.W = H
    model.pixel_val_to_cpu = False
    scenes, local = [], []
    for k in range(scenes_total):
        inp = synthetic.to_device(synthetic.make_inputs(1, H, H, seed=500 + k), dev)
        z = [t.to(dev) for t in synthetic.make_features(1, H, seed=500 + k)]
        scenes.append((inp, z if rank == 0 else None))
        local.append((inp, z))

    def timed(fn):
        fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    nb = [0]

    def with_bcast():
        with torch.no_grad():
            _, nb[0] = sharding.render_scenes_pipelined(model, scenes, src=0, device=dev)

    def without_bcast():
        with torch.no_grad():
            for inp, z in local:
                sharding.render_sharded(model, inp, z, rank=rank, world=world, gather=True)
    ms_b, ms_n = timed(with_bcast), timed(without_bcast)
    rays = scenes_total * H * H
    return {"workload": f"{scenes_total} scenes of {H}x{H} rays, {P} samples, rendered one scene at a time with the rays of "
                        f"each scene split across {world} GPUs; fp32 maps broadcast from rank 0 (NCCL), one scene ahead",
            "value": rays / (ms_b * 1e-3), "unit": "rays/s", "ms": ms_b, "scaling": "strong",
            "ms_maps_already_resident": ms_n, "value_maps_already_resident": rays / (ms_n * 1e-3),
            "broadcast_bytes_per_rank": nb[0], "broadcast_exposed_ms": ms_b - ms_n,
            "broadcast_gbs_if_serial": nb[0] / max(1e-9, (ms_b - ms_n) * 1e-3) / 1e9 if ms_b > ms_n else None}


def general_branches(ctx, H, P, steps, warmup, precision="fp32"):
    """The other forward branches (n_view 1 / 3, no_latent_concat) through the public
    forward: one scene per step, inputs resident, device-timed; parity of a 256-ray slice against the oracle."""
    from cross_attention_renderer_b200.models import CrossAttentionRenderer
    dev = ctx["dev"]
    out = {}
    for name, nv, noconcat in (("n_view_1", 1, False), ("n_view_3", 3, False), ("no_latent_concat", 2, True)):
        inp = synthetic.make_inputs(1, H, H, seed=500 + nv, n_ctx=nv)
        z = synthetic.make_features(1, H, seed=500 + nv, n_view=nv)
        sd = synthetic.make_state_dict(seed=0, n_view=nv, no_latent_concat=noconcat)
        m = CrossAttentionRenderer(n_view=nv, npoints=P, no_latent_concat=noconcat, precision=precision).to(dev).eval()
        m.load_state_dict(sd, strict=False)
        m.H = m.W = H
        m.pixel_val_to_cpu = False
        inp_d, z_d = synthetic.to_device(inp, dev), [t.to(dev) for t in z]

        def step():
            m.release_features()
            with torch.no_grad():
                return m(inp_d, z=z_d)
        for _ in range(warmup):
            res = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            res = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = m.last_launch_count
        # parity: the same inputs cut to 256 rays through the oracle restatement of this branch
        from oracle import car_oracle as orc
        sel = slice(0, 256)
        sub = {"query": dict(inp["query"]), "context": inp["context"]}
        sub["query"]["uv"] = inp["query"]["uv"][:, :, sel].contiguous()
        fn = {1: orc.render_single_view, 2: orc.render, 3: orc.render_three_views}[nv]
        with torch.no_grad():
            ref = fn(sd, sub, z, H, H, P, **({"no_latent_concat": True} if noconcat else {}))
        got = res["rgb"][:, :, sel].cpu()
        rel = float((got - ref["rgb"]).abs().max() / ref["rgb"].abs().max().clamp_min(1e-12))
        out[name] = {"workload": f"{H}x{H} target, n_view={nv}{', no_latent_concat' if noconcat else ''}, {P} samples, 1 scene, "
                                 "car_render_forward_general (unfused; per-sample GEMMs: "
                                 + ("tcgen05 hi+lo" if precision == "fp32" else "exact fp32 SIMT") + ")",
                     "value": round(H * H / ms * 1e3, 1), "unit": "rays/s", "ms_per_step": round(ms, 3),
                     "gpu_launches_per_step": launches,
                     "parity": {"rays": 256, "rgb_rel_err_vs_oracle": rel, "ok": rel <= 1e-4}}
    return out


def reference_on_gpu(dev, H, P, rays=8192):
    """the unmodified reference (oracle/_ref) executed on the same B200, forward(input, z=z) on
    8192-ray chunks like render_realestate10k_traj.py:96; TF32 off (the parity-grade setting) and torch's
    default flags (cuDNN convolutions may use TF32).  A reported row, not the driver's ratio."""
    from oracle import ref_loader
    if not ref_loader.available():
        return {"unavailable": "oracle/_ref not built"}
    out = {"what": "unmodified reference files (oracle/_ref), CrossAttentionRenderer.forward(input, z=z) on this GPU",
           "rays_per_call": rays, "workload": f"{H}x{H} maps, {P} samples, 2 views, 1 scene"}
    try:
        sd = synthetic.make_state_dict(seed=0)
        inp = synthetic.to_device(synthetic.make_inputs(1, H, H, seed=100, rays=rays), dev)
        z = [t.to(dev) for t in synthetic.make_features(1, H, seed=100)]
        m = ref_loader.build_model(sd, H, P, device=dev)

        def timed(n):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ev0.record()
            for _ in range(n):
                ref_loader.render(m, inp, z)
            ev1.record()
            torch.cuda.synchronize()
            return n * rays / (ev0.elapsed_time(ev1) * 1e-3), n * rays / (time.perf_counter() - t0)
        with ref_loader.strict_fp32():
            timed(2)
            dev_rps, wall_rps = timed(5)
            out["tf32_off"] = {"value": dev_rps, "unit": "rays/s", "wall_value": wall_rps}
        timed(2)
        dev_rps, wall_rps = timed(5)
        out["torch_defaults"] = {"value": dev_rps, "unit": "rays/s", "wall_value": wall_rps,
                                 "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32),
                                 "matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32)}
        out["peak_mem_gb"] = torch.cuda.max_memory_allocated(dev) / 1e9
        del m, z, inp
        torch.cuda.empty_cache()
    except Exception as exc:  # pragma: no cover
        out["error"] = repr(exc)[:300]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CAR_PRECISION", "fp32"),
                    choices=["fp32_simt", "fp32", "bf16"])
    ap.add_argument("--scenes", type=int, default=12, help="scenes per GPU")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--cpu-rays", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_configs and reference_gpu")
    ap.add_argument("--no-stage-profile", action="store_true", help="diagnostic: no per-kernel events in the timed region")
    ap.add_argument("--chunk-rays", type=int, default=0, help="diagnostic: rays per workspace chunk (0 = library default)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import torch.distributed as dist
    from cross_attention_renderer_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    b, H, P = args.scenes, args.size, args.samples
    lib = _lib.load()

    def tail_bytes(rows, kh, v):
        # k_tail<A> + k_tail<B>, per ray.  A: reads relu(key_map) + V + the geometry record, writes Q1 (fp32),
        # at_wt, zsum; B: reads V + Q1 + local_coords + row bias + zsum, writes z.
        q1 = rows * 128 * 4
        return (kh + v + rows * 128 + q1 + rows * 4 + 288 * 4) + (v + q1 + rows * 64 + 128 * 4 + 288 * 4 * 2)
    ctx = {"world": world, "rank": rank, "local": local, "dev": dev, "lib": lib, "chunk_rays": args.chunk_rays,
           "stage_profile": not args.no_stage_profile, "tail_bytes": tail_bytes}
    e2e_steps = 0 if args.no_e2e else max(10, args.steps)
    main_res = measure(ctx, args.precision, b, H, P, args.steps, args.warmup, e2e_steps, with_clocks=True)

    # ---- extra configurations measured after the headline region (same method, fewer e2e steps) ----
    extra = {}
    if not args.no_extra and (H, P, args.precision) == (256, 64, "fp32"):
        r = measure(ctx, "bf16", b, 256, 64, args.steps, args.warmup, 0 if args.no_e2e else 5, with_clocks=True, seed_base=100)
        if rank == 0:
            extra["c3_bf16"] = {"workload": f"256x256 target, 2 views, 64 samples, bf16 maps + single bf16 MMA, {b} scenes per GPU "
                                            f"(BASELINE config 3 layout at {world} GPU(s))", **_public(r)}
        if world == 1:
            r = measure(ctx, "fp32", 1, 512, 128, args.steps, args.warmup, 0 if args.no_e2e else 5, with_clocks=True, seed_base=300)
            extra["c4_512_p128"] = {"workload": "512x512 target, 2 views, 128 samples, fp32 maps, 1 scene (BASELINE config 4)", **_public(r)}
            extra["general_branches"] = general_branches(ctx, 256, 64, max(2, min(3, args.steps)), 2)
    # ---- strong scaling over a few scenes (N > 1): the scenes' rays are split across all ranks, the feature maps
    # exist on rank 0 only and are broadcast one scene ahead (sharding.render_scenes_pipelined) -------------------
    if world > 1 and not args.no_extra and (H, P, args.precision) == (256, 64, "fp32"):
        st = strong_scaling(ctx, scenes_total=4, H=256, P=64, reps=max(3, args.steps))
        if rank == 0:
            extra["strong_4_scenes"] = st
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_base = None
    ref_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(H, P)
        arm.run(128, 100)
        dt = arm.run(args.cpu_rays, 0)
        cpu_base = {"value": args.cpu_rays / dt, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": arm.kind,
                    "sample": f"{args.cpu_rays} rays of one {H}x{H}/{P}-sample scene, {dt:.1f} s"}
    if world == 1 and not args.no_extra:
        ref_gpu = reference_on_gpu(dev, H, P)
    maps = "bf16 maps" if args.precision == "bf16" else "fp32 maps"
    if (H, P) == (256, 64):
        tag = "BASELINE config 3 layout: 12 scenes per GPU" if args.precision == "bf16" else "BASELINE config 2"
    elif (H, P) == (512, 128):
        tag = "BASELINE config 4"
    else:
        tag = "non-BASELINE size"
    workload = f"{H}x{H} target, 2 views, {P} samples, {maps}, {b} scenes per GPU ({tag})"
    m = main_res
    line = {
        "metric": METRIC, "value": m["value"], "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32_simt": "f32", "fp32": "f32 (3xbf16 tcgen05 split, fp32 accumulate)", "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": {"workload": workload,
                   "scenes_per_gpu": b, "rays_per_step": m["total_rays"], "precision": args.precision,
                   "parallelism": f"ray/scene sharding x{world}, all_gather of tiles",
                   "l2": "inputs (feature maps %.0f MB per GPU) exceed the 126 MB L2" % m["feature_mb"],
                   "repacked_every_step": True, "extra_warmup_steps": m["extra_warm"]},
        "step_ms": m["step_ms"], "clocks": m["clocks"], "e2e": m["e2e"], "gpu_launches": m["gpu_launches"],
        "roofline": m.get("roofline"), "roofline_all": m.get("roofline_all"), "stage_share": m.get("stage_share"),
        "stage_ms_per_step": m.get("stage_ms_per_step"), "build_id": m.get("build_id"),
        "cpu_baseline": cpu_base, "parity": m.get("parity"),
        "extra_configs": extra or None, "reference_gpu": ref_gpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _public(r):
    keys = ("value", "ms_per_step", "step_ms", "clocks", "e2e", "gpu_launches", "roofline", "stage_ms_per_step", "parity")
    out = {k: r.get(k) for k in keys}
    out["unit"] = "rays/s"
    return out


if __name__ == "__main__":
    main()
