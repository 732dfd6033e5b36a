#!/usr/bin/env python
"""Benchmark of the MFT diffraction hot path (BASELINE.json metric):

    polychromatic PSFs/s, forward + gradient, 1024 -> 512 px, 64 wavelengths (config 3)

One step = one polychromatic PSF of one point source (64 wavelengths, hex-NRM pupil with a
21-mode OPD basis) plus the gradient of sum(G * psf) w.r.t. the 21 basis coefficients:
basis eval -> pupil phasor -> 2 phasor-GEMM stages per wavelength -> |E|^2 spectral sum ->
cotangent -> 2 adjoint stages per wavelength -> OPD-bar -> basis reduce.
With N GPUs every rank owns one source (PointSources with N stars, weak scaling) and the
summed image and the coefficient gradient are all-reduced over NCCL.

  python bench.py [--gpus N --steps K --warmup W]            our CUDA arm
  python bench.py --impl reference [...]                     the reference's algorithm on the
                                                             host CPU cores (oracle/, NumPy+torch)
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "polychromatic PSFs/s fwd+grad at 1024->512 px (64 wavelengths); MFT TFLOP/s vs tensor peak"
UNIT = "PSF+grad/s"
WORKLOAD = ("c3: 1024 px hex-NRM pupil (7 holes, 21-mode OPD basis), 64 wavelengths 4.1-4.5 um, "
            "oversampled MFT to 512x512, forward + adjoint (grad of sum(G*psf) w.r.t. 21 coefficients)")


def mft_flops(n, m):
    return 8.0 * m * n * (n + m)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return pk, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def measure_tensor_peak(dev, seconds=3.0):
    """Live MMA-only probe (dlux_tc_peak_probe): burst and sustained TFLOP/s of kind::tf32 and of the
    GEMM's own 4 tf32 + 4 bf16 instruction mix, on this board, in this process."""
    import ctypes as C
    import torch
    from dlux_b200 import _lib
    st = torch.cuda.current_stream(dev).cuda_stream
    sink = torch.zeros(1, device=dev)
    sp = C.c_void_p(sink.data_ptr())
    out = {}
    for name, kind in (("tf32", 0), ("mix", 2)):
        nb = 4000
        for _ in range(2):                                # calibrate to ~15 ms per launch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fl = _lib.tc_peak_probe(kind, nb, st, sp)
            e1.record()
            torch.cuda.synchronize()
            nb = max(200, int(nb * 15.0 / max(e0.elapsed_time(e1), 1e-3)))
        burst = 0.0
        for _ in range(3):
            time.sleep(0.3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fl = _lib.tc_peak_probe(kind, nb, st, sp)
            e1.record()
            torch.cuda.synchronize()
            burst = max(burst, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out[name + "_burst"] = burst
        if seconds > 0:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0, tot = time.time(), 0.0
            e0.record()
            while time.time() - t0 < seconds:
                for _ in range(10):
                    tot += _lib.tc_peak_probe(kind, nb, st, sp)
                torch.cuda.current_stream().synchronize()
            e1.record()
            torch.cuda.synchronize()
            out[name + "_sustained"] = tot / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return out


def load_traffic():
    """dram__bytes per GEMM launch from the committed ncu capture of this round (profiles/)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            t["source"] = "profiles/" + name
            return t
        except Exception:
            continue
    return None


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock, power and throttle reasons of one board, sampled every ~10 ms by a thread of this process through
    NVML (pynvml); `nvidia-smi -lms 20` through a line-buffered pipe when NVML cannot be loaded.  Every sample carries
    the wall-clock time it was taken at, so `stop(t0, t1)` reports exactly the samples of the timed window."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self._stop = index, [], None, None, False
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["stdbuf", "-oL", "nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [(n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else 0x8),
                (n.nvmlClocksEventReasonHwThermalSlowdown if hasattr(n, "nvmlClocksEventReasonHwThermalSlowdown") else 0x40),
                (n.nvmlClocksEventReasonSwThermalSlowdown if hasattr(n, "nvmlClocksEventReasonSwThermalSlowdown") else 0x20),
                (n.nvmlClocksEventReasonSwPowerCap if hasattr(n, "nvmlClocksEventReasonSwPowerCap") else 0x4)]
        while not self._stop:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append((time.time(), sm, self.mx, pw, [nm for nm, b in zip(self.NAMES, bits) if rs & b]))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        import datetime
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                pw = float(c[3]) if c[3].replace(".", "", 1).isdigit() else None
                self.rows.append((ts, float(c[1]), float(c[2]), pw,
                                  [nm for nm, v in zip(self.NAMES, c[4:8]) if v.lower().startswith("active")]))
            except Exception:
                pass

    def alive(self):
        return bool(self.rows)

    def stop(self, t0=None, t1=None):
        """Statistics of the samples taken in [t0, t1] (time.time(); default: all of them)."""
        self._stop = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if not self.source:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"], "samples": 0}
        rows = [r for r in list(self.rows) if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        sm = [r[1] for r in rows]
        pw = [r[3] for r in rows if r[3] is not None]
        reasons = sorted({x for r in rows for x in r[4]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": reasons, "samples": len(sm),
                "source": self.source}


# ---------------------------------------------------------------------------- CPU arm
def cpu_reference_sample(cfg, n_sample, threads):
    """Oracle forward + torch-autograd gradient for `n_sample` of the wavelengths."""
    import torch
    from oracle import mft_oracle as O
    from oracle import torch_twin
    torch.set_num_threads(threads)
    idx = np.linspace(0, len(cfg["wavelengths"]) - 1, n_sample).round().astype(int)
    wls, w = cfg["wavelengths"][idx], cfg["weights"][idx]
    M = cfg["psf_npixels"] * cfg["oversample"]
    ps = O.arcsec2rad(np.float32(cfg["psf_pixel_scale"]) / np.float32(cfg["oversample"]))
    c = torch.tensor(cfg["coefficients"], requires_grad=True)
    t0 = time.perf_counter()
    psf = torch_twin.poly_psf(cfg["transmission"], None, wls, w, diameter=cfg["diameter"], psf_npixels=M,
                              pixel_scale_rad=ps, offset=cfg["positions"][0], basis=cfg["basis"],
                              coefficients=c, dtype=np.float32)
    (psf * torch.as_tensor(cfg["G"])).sum().backward()
    dt = time.perf_counter() - t0
    return dt, psf.detach().numpy(), c.grad.numpy(), idx


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dlux_b200 import workloads
    cfg = workloads.config("c3")
    cores = os.cpu_count() or 1
    L = len(cfg["wavelengths"])
    n_sample = L                                          # every step is one FULL 64-wavelength PSF + gradient
    times = []
    for i in range(args.warmup + args.steps):
        dt, _, _, _ = cpu_reference_sample(cfg, n_sample, cores)
        if i >= args.warmup:
            times.append(dt)
    t_unit = float(np.mean(times))                        # seconds per full 64-wavelength PSF+grad
    value = 1.0 / t_unit
    sample = (f"all {L} wavelengths of the c3 PSF+grad per step (NumPy complex64 oracle transfer matrices "
              f"+ torch-CPU complex64 matmul/autograd on {cores} host threads), nothing extrapolated")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_unit,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex64",
           "data": "synthetic",
           "config": {"workload": WORKLOAD, "n_pupil": cfg["wf_npixels"],
                      "n_psf": cfg["psf_npixels"] * cfg["oversample"], "n_wavelengths": L,
                      "n_basis": len(cfg["coefficients"]), "sources_per_gpu": 1},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "restatement of the reference (oracle/), not the reference itself: JAX is not installable here"}
    print(json.dumps(out))



def run_c4_strong(dev, world, rank, steps=3, n_stars=64):
    """BASELINE config 4 at fixed TOTAL size (strong scaling): 2048 px binary-phase-mask pupil, `n_stars` stars x
    64 wavelengths -> 256x256, gradients w.r.t. star positions, fluxes and the phase mask.  Stars are sharded
    over the ranks; one NCCL all-reduce of the image forward, one flat all-reduce of the gradients backward."""
    import torch
    import torch.distributed as dist
    import dlux_b200 as dl
    from dlux_b200 import distributed as D
    from dlux_b200 import workloads
    cfg = workloads.config("c4")
    N, M, L = cfg["wf_npixels"], cfg["psf_npixels"], len(cfg["wavelengths"])
    up = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    phase = up(cfg["phase"]).requires_grad_(True)
    pos = up(cfg["positions"][:n_stars]).requires_grad_(True)
    flux = up(cfg["fluxes"][:n_stars]).requires_grad_(True)
    G = up(cfg["G"])
    layer = dl.Optic(up(cfg["transmission"]), None, phase, normalise=True, device=dev)
    optics = dl.AngularOpticalSystem(N, cfg["diameter"], [("mask", layer)], M, cfg["psf_pixel_scale"], device=dev)

    def step():
        for t in (phase, pos, flux):
            t.grad = None
        psf = D.sharded_point_sources_model(optics, cfg["wavelengths"], pos, flux, cfg["weights"])
        (psf * G).sum().backward()
        D.all_reduce_grads([phase, pos, flux])
        return psf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    flops = 2 * n_stars * L * mft_flops(N, M)
    return {"workload": f"c4-lite: 2048 px binary phase mask, {n_stars} stars x {L} wavelengths -> {M}x{M}, "
                        "PSF + gradients w.r.t. positions, fluxes and the phase mask; TOTAL size fixed (strong scaling)",
            "ms_per_step": ms, "star_wavelength_psf_grad_per_s": n_stars * L * 1e3 / ms,
            "psf_grad_per_s": 1e3 / ms, "mft_tflops_algorithmic_all_gpus": flops / (ms * 1e-3) / 1e12,
            "stars_per_gpu": n_stars / world,
            "nccl_bytes_per_step": 0 if world == 1 else int(4 * M * M + 4 * (N * N + 3 * n_stars)),
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}


# ---------------------------------------------------------------------------- CUDA arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dlux_b200 as dl
    from dlux_b200 import _lib, ops, workloads
    from dlux_b200.utils import propagation as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    cfg = workloads.config("c3")
    N, M = cfg["wf_npixels"], cfg["psf_npixels"] * cfg["oversample"]
    L, nz = len(cfg["wavelengths"]), len(cfg["coefficients"])
    # one source per rank (PointSources with `world` stars); rank r owns star r
    all_positions = np.stack([(np.random.default_rng(100 + r).uniform(-1, 1, 2) * 2e-7).astype(np.float32)
                              for r in range(world)]) if world > 1 else cfg["positions"][:1]
    all_fluxes = np.ones(world, np.float32)
    position = all_positions[rank]
    flux = np.float32(1.0)

    up = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    T_d, basis_d, G_d = up(cfg["transmission"]), up(cfg["basis"]), up(cfg["G"])
    coeffs_d = up(cfg["coefficients"])
    ps_in = np.float32(np.float32(cfg["diameter"]) / np.float32(N))
    ps_out = P.arcsec2rad(np.float32(cfg["psf_pixel_scale"]) / np.float32(cfg["oversample"]))
    s_h, nrm_h = P.mft_geometry(cfg["wavelengths"], N, ps_in, M, ps_out)
    k_d = up((np.float32(2 * np.pi) / cfg["wavelengths"]).astype(np.float32))
    s_d, nrm_d = up(s_h.astype(np.float32)), up(nrm_h.astype(np.float32))
    w_d = up((cfg["weights"] * flux).astype(np.float32)).reshape(1, L)
    delta_d = up((position[None, None, :] * np.float32(cfg["diameter"]) /
                  cfg["wavelengths"][None, :, None]).astype(np.float32))

    def step_device(sparse=False):
        """Hot path with every input already resident in HBM."""
        opd = ops.basis_eval(basis_d, coeffs_d)
        psf, field = ops.polypsf_fwd(T_d, opd, None, k_d, s_d, nrm_d, w_d, delta_d, N, M, True, None, True,
                                     sparse=sparse)
        if world > 1:
            dist.all_reduce(psf)
        opd_bar = ops.polypsf_bwd(T_d, opd, None, k_d, s_d, nrm_d, w_d, delta_d, field, G_d, N, M,
                                        True, None, True, False, False, sparse=sparse)[0]
        cbar = ops.basis_reduce(basis_d, opd_bar, coeffs_d.shape)
        if world > 1:
            dist.all_reduce(cbar)
        return psf, cbar

    # ---- end-to-end arm: the public API with HOST buffers (pinned), H2D + D2H inside the step
    coeffs_h = torch.as_tensor(cfg["coefficients"]).pin_memory()
    G_h = torch.as_tensor(cfg["G"]).pin_memory()
    psf_h = torch.empty((M, M), dtype=torch.float32).pin_memory()
    grad_h = torch.empty(nz, dtype=torch.float32).pin_memory()
    from dlux_b200 import distributed as D

    # the model is built once (as in a fitting loop); every step gets new coefficients and a new
    # image-plane cotangent from the host
    e2e_layer = dl.BasisOptic(basis_d, T_d, coeffs_d, normalise=True, effect="opd", device=dev)
    optics = dl.AngularOpticalSystem(N, cfg["diameter"], [("pupil", e2e_layer)], cfg["psf_npixels"],
                                     cfg["psf_pixel_scale"], cfg["oversample"], device=dev)

    def step_e2e():
        c = coeffs_h.to(dev, non_blocking=True).requires_grad_(True)
        G = G_h.to(dev, non_blocking=True)
        e2e_layer.coefficients = c
        # PointSources(world stars).model(optics), sources sharded one per rank + NCCL all-reduce
        psf = D.sharded_point_sources_model(optics, cfg["wavelengths"], all_positions, all_fluxes,
                                            cfg["weights"])
        loss = (psf * G).sum()
        loss.backward()
        D.all_reduce_grads([c])
        psf_h.copy_(psf.detach(), non_blocking=True)
        grad_h.copy_(c.grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(psf_h[0, 0])

    # the same step captured once into a CUDA graph through the public API (dl.GraphedValueAndGrad, the role
    # jax.jit plays for the reference's value_and_grad loop); single GPU only (no NCCL inside the capture)
    gstep = None
    if world == 1:
        stars = dl.PointSources(cfg["wavelengths"], all_positions, all_fluxes, weights=cfg["weights"])

        def loss_fn(c, G):
            e2e_layer.coefficients = c
            psf = stars.model(optics)
            return (psf * G).sum(), psf
        try:
            gstep = dl.GraphedValueAndGrad(loss_fn, [coeffs_d, G_d], has_aux=True, argnums=[0])
        except Exception as e:                          # pragma: no cover
            print("bench: CUDA-graph capture of the e2e step failed:", repr(e), file=sys.stderr)
            gstep = None

    # ... and with the host traffic inside the graph, the cotangent's upload under the forward pass and the
    # image's download under the backward pass (dl.GraphedFitStep)
    fstep = None
    if world == 1:
        def model_fn(c):
            e2e_layer.coefficients = c
            return stars.model(optics)
        try:
            fstep = dl.GraphedFitStep(model_fn, lambda psf, G: (psf * G).sum(), [coeffs_d], [G_d],
                                      host_params=[coeffs_h], host_data=[G_h], host_image=psf_h, host_grads=[grad_h])
        except Exception as e:                          # pragma: no cover
            print("bench: GraphedFitStep capture failed:", repr(e), file=sys.stderr)
            fstep = None

    # N > 1: the local part of the step (this rank's shard, forward and backward) replayed from two CUDA graphs
    # (torch.cuda.make_graphed_callables), the two NCCL all-reduces issued between / after them as usual
    # OPT-IN (DLUX_BENCH_SHARD_GRAPHS=1): measured at N = 2 (+11 % over eager launches), but the capture hung
    # the N = 4 run, so the default e2e path at N > 1 stays the eager step
    gshard = None
    if world > 1 and os.environ.get("DLUX_BENCH_SHARD_GRAPHS", "0") == "1":
        def local_model(c):
            e2e_layer.coefficients = c
            return D.sharded_point_sources_model(optics, cfg["wavelengths"], all_positions, all_fluxes,
                                                 cfg["weights"], reduce=False)
        try:
            gshard = torch.cuda.make_graphed_callables(local_model, (coeffs_d.detach().clone().requires_grad_(True),))
        except Exception as e:                          # pragma: no cover
            print("bench: graph capture of the sharded step failed:", repr(e), file=sys.stderr)
            gshard = None
        ok = torch.tensor([1 if gshard is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)       # every rank takes the same path
        if int(ok.item()) == 0:
            gshard = None

    def step_e2e_sharded_graphs():
        c = coeffs_h.to(dev, non_blocking=True).requires_grad_(True)
        G = G_h.to(dev, non_blocking=True)
        psf_local = gshard(c)
        psf = psf_local.detach().clone()
        dist.all_reduce(psf)
        psf_h.copy_(psf, non_blocking=True)
        psf_local.backward(G)                           # d sum(G * psf_total) / d psf_local = G on every rank
        D.all_reduce_grads([c])
        grad_h.copy_(c.grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(psf_h[0, 0])

    def step_e2e_fit():
        fstep.step()
        return float(psf_h[0, 0])

    def step_e2e_graph():
        gstep.static[0].detach().copy_(coeffs_h, non_blocking=True)
        gstep.static[1].copy_(G_h, non_blocking=True)
        gstep.graph.replay()
        psf_h.copy_(gstep.aux.detach(), non_blocking=True)
        grad_h.copy_(gstep.grads[0], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(psf_h[0, 0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)                      # every rank samples its own board
    sampler.start()                                    # (early: nvidia-smi's first line takes ~0.2 s)
    t_w = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    while time.perf_counter() - t_w < 0.4 and not sampler.alive():   # untimed: wait for the sampler to be alive
        step_device()
        torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    t_c0 = time.time()
    ms_total = timed(step_device, args.steps)
    launches = _lib.launch_count() - launches0
    ms_step = ms_total / args.steps
    value = world * 1e3 / ms_step                      # PSF+grad per second, all ranks

    # ---- roofline pass: per-kernel time of the phasor GEMMs (events on the launching stream)
    _lib.profile_enable(True)
    timed(step_device, args.steps)
    _lib.profile_enable(False)
    t_c1 = time.time()
    gemm_ms, gemm_launches, gemm_flops = _lib.profile_read()
    # clocks over the timed region + the identical roofline pass right after it (>= 0.1 s even at 20 steps)
    clocks = sampler.stop(t_c0, t_c1)
    clocks["window_s"] = t_c1 - t_c0

    # ---- sustained regime: the same step back to back for >= 3 s (power-capped clocks)
    sus = None
    if args.sustained > 0:
        n_sus = max(args.steps, int(args.sustained * 1e3 / ms_step))
        s2 = ClockSampler(local)
        s2.start()
        ms_sus = timed(step_device, n_sus) / n_sus
        clk_sus = s2.stop()
        _lib.profile_enable(True)
        timed(step_device, min(n_sus, 200))
        _lib.profile_enable(False)
        g_ms, g_n, g_fl = _lib.profile_read()
        sus = {"value": world * 1e3 / ms_sus, "unit": UNIT, "ms_per_step": ms_sus, "steps": n_sus,
               "gemm_ms_per_step": g_ms / min(n_sus, 200), "gemm_tflops_algorithmic": g_fl / (g_ms * 1e-3) / 1e12,
               "clocks": clk_sus}

    # ---- opt-in exact zero-block skipping (reported separately; the headline stays dense)
    sparse_rec = None
    if args.sparse:
        d_psf, d_grad = step_device()
        s_psf, s_grad = step_device(True)
        for _ in range(3):
            step_device(True)
        ms_sp = timed(lambda: step_device(True), args.steps) / args.steps
        blk = lambda bh, bw: float((T_d.reshape(N // bh, bh, N // bw, bw).abs().amax((1, 3)) == 0).float().mean()) \
            if N % bh == 0 and N % bw == 0 else None
        sparse_rec = {"value": world * 1e3 / ms_sp, "unit": UNIT, "ms_per_step": ms_sp,
                      "skip_fraction_stage1_chunks": blk(256, 16), "skip_fraction_adjoint_output_blocks": blk(128, 256),
                      "bit_identical_psf": bool(torch.equal(d_psf, s_psf)),
                      "grad_rel_diff_vs_dense": float((d_grad - s_grad).norm() / d_grad.norm()),
                      "note": "blocks of the pupil with transmission == 0 are neither contracted (forward stage 1) nor "
                              "produced (last adjoint stage); exact, partial sums keep the dense boundaries"}

    # ---- e2e arm
    for _ in range(3):
        step_e2e()
    psf_ref_h, grad_ref_h = psf_h.clone(), grad_h.clone()      # what the eager step delivered to the host

    def same_as_eager(step):                                   # the captured variants must deliver the same
        psf_h.zero_(); grad_h.zero_()
        step()
        r = lambda a, b: float((a - b).norm() / b.norm())
        return max(r(psf_h, psf_ref_h), r(grad_h, grad_ref_h))
    e2e_dev = {}
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_eager = world * 1e3 / ms_e2e
    e2e_api = "eager public API (BasisOptic + AngularOpticalSystem + PointSources.model + backward)"
    if gstep is not None:
        e2e_dev["graph"] = same_as_eager(step_e2e_graph)
        for _ in range(3):
            step_e2e_graph()
        ms_g = timed(step_e2e_graph, args.steps) / args.steps
        if ms_g < ms_e2e:
            ms_e2e = ms_g
            e2e_api = ("the same public-API step captured once by dl.GraphedValueAndGrad and replayed (host buffers copied "
                       "in and out every step)")
    if gshard is not None:
        e2e_dev["sharded_graphs"] = same_as_eager(step_e2e_sharded_graphs)
        for _ in range(3):
            step_e2e_sharded_graphs()
        ms_s = timed(step_e2e_sharded_graphs, args.steps) / args.steps
        if ms_s < ms_e2e:
            ms_e2e = ms_s
            e2e_api = ("the public-API step with this rank's shard (forward and backward) replayed from two CUDA graphs "
                       "(torch.cuda.make_graphed_callables) and the two NCCL all-reduces issued eagerly; host buffers "
                       "copied in and out every step")
    e2e_serial = e2e_fit = None
    if fstep is not None:
        e2e_dev["graph_with_host_io"] = same_as_eager(step_e2e_fit)
        for _ in range(3):
            step_e2e_fit()
        ms_f = timed(step_e2e_fit, args.steps) / args.steps
        e2e_fit = world * 1e3 / ms_f
        if ms_f < ms_e2e:
            e2e_serial = world * 1e3 / ms_e2e
            ms_e2e = ms_f
            e2e_api = ("the public-API step (BasisOptic + AngularOpticalSystem + PointSources.model + autograd) captured "
                       "once by dl.GraphedFitStep: pinned host buffers copied in and out EVERY step by memcpy nodes of "
                       "the graph, the cotangent's upload under the forward pass, the image's download under the "
                       "backward pass")
    e2e_value = world * 1e3 / ms_e2e

    # ---- the north-star multi-GPU workload at fixed total size (every rank takes part)
    c4 = None
    if args.c4_stars > 0:
        try:
            c4 = run_c4_strong(dev, world, rank, n_stars=args.c4_stars)
        except Exception as e:                           # pragma: no cover
            c4 = {"error": repr(e)}

    # clocks of the other ranks (the driver only sees rank 0's line)
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        clocks = dict(clocks, per_rank_sm_mhz=[c.get("sm_mhz") for c in allc],
                      per_rank_reasons=[c.get("reasons") for c in allc])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = load_peaks()
    probe = None
    try:
        probe = measure_tensor_peak(dev, seconds=3.0 if args.sustained > 0 else 0.0)
    except Exception as e:                               # pragma: no cover
        probe = None
        probe_err = repr(e)
    if probe:
        tf32_burst, tf32_sus = probe["tf32_burst"], probe.get("tf32_sustained")
        peak_note = ("dense kind::tf32 tcgen05 rate MEASURED live by the library's MMA-only probe (cta_group::2, "
                     "M256xN128, operands resident): burst = best ~15 ms launch, sustained = >= 3 s back to back")
    else:
        tf32_burst, tf32_sus = peaks["bf16_tflops"] / 2.0, peaks["bf16_tflops_sustained"] / 2.0
        peak_note = f"probe unavailable ({probe_err}); FALLBACK dense TF32 = bf16/2 from MEASURED_PEAKS.json ({peak_src})"
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12        # algorithmic TFLOP/s inside the GEMM kernels
    flops_step = 2 * L * mft_flops(N, M)                   # forward + adjoint, per source
    cores = os.cpu_count() or 1
    cpu_reference_sample(cfg, 4, cores)                       # warm the thread pools
    cpu_dt, cpu_psf, cpu_grad, _ = cpu_reference_sample(cfg, L, cores)   # one full 64-wavelength PSF + gradient
    cpu_value = 1.0 / cpu_dt
    parity = None
    if world == 1:   # same star, same inputs: the full-size step against the oracle (checker, not product)
        psf_d, cbar_d = step_device()
        rel = lambda a, b: float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
        parity = {"psf_rel_l2_vs_oracle": rel(psf_d.cpu().numpy(), cpu_psf),
                  "grad_rel_l2_vs_oracle": rel(cbar_d.cpu().numpy(), cpu_grad), "tolerance": 1e-5}

    traffic = load_traffic()
    # SURVEY 8(d) algorithmic bytes of one C3 PSF + gradient: T + OPD 8 MiB, basis 2 x 4 nz N^2, PSF 4 M^2,
    # OPD-bar 4 N^2 -- spread over the GEMM launches of a step for the per-launch figure
    alg_bytes_step = 8.0 * N * N + 2 * 4.0 * nz * N * N + 4.0 * M * M + 4.0 * N * N
    n_gemm = max(1.0, gemm_launches / args.steps)
    roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 TS-form split-precision phasor GEMM)",
            "achieved": achieved, "peak": tf32_burst, "unit": "TFLOP/s", "frac": achieved / tf32_burst,
            "frac_executed": 2 * achieved / tf32_burst,
            "peak_note": peak_note + "; achieved = algorithmic FLOPs (8 per complex MAC) / CUDA-event time of the GEMM "
                         "launches in the timed (burst) region; the split-precision scheme executes 2x the algorithmic "
                         "FLOPs in tf32-equivalent tensor time (1x tf32 + 2x bf16 at twice the rate), so frac <= 0.5 by "
                         "construction and frac_executed is the tensor-pipe utilisation",
            "traffic": (traffic or {}).get("dram_bytes_per_gemm_launch"),
            "traffic_source": (traffic or {}).get("source"),
            "traffic_algorithmic": alg_bytes_step / n_gemm,
            "traffic_algorithmic_note": "SURVEY 8(d): %.3f GB per PSF+grad step / %.0f GEMM launches" % (alg_bytes_step / 1e9, n_gemm),
            "gemm_ms_per_step": gemm_ms / args.steps,
            "gemm_share_of_step": (gemm_ms / args.steps) / ms_step,
            "gemm_launches_per_step": gemm_launches / args.steps}
    if probe:
        roof["probe"] = probe
        roof["frac_of_mix_ceiling"] = achieved / (probe["mix_burst"] / 3.0)
    if sus is not None:
        speak = tf32_sus if tf32_sus else tf32_burst
        sus["peak"] = speak
        sus["frac"] = sus["gemm_tflops_algorithmic"] / speak
        sus["frac_executed"] = 2 * sus["frac"]
        if probe and probe.get("mix_sustained"):
            sus["frac_of_mix_ceiling"] = sus["gemm_tflops_algorithmic"] / (probe["mix_sustained"] / 3.0)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "tf32 + 2x bf16 split, fp32 accumulate (complex64 parity)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_pupil": N, "n_psf": M, "n_wavelengths": L, "n_basis": nz,
                   "sources_per_gpu": 1, "parallelism": f"sources sharded over {world} GPU(s), NCCL all-reduce of PSF + coefficient gradient",
                   "l2": "no explicit flush: each step streams > 1 GB of operand planes (> 126 MB L2)"},
        "mft_tflops": {"algorithmic": world * flops_step / (ms_step * 1e-3) / 1e12,
                       "executed_tensor_tf32_equivalent": 2 * world * flops_step / (ms_step * 1e-3) / 1e12,
                       "flops_per_step_per_gpu": flops_step},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "api": e2e_api, "eager_value": e2e_eager,
                "serial_copies_value": e2e_serial, "graph_with_host_io_value": e2e_fit,
                "rel_diff_vs_eager_step": e2e_dev,
                # (the eager step also re-uploads the wavelength / weight vectors; the captured steps keep them resident)
                "h2d_bytes_per_step": int(coeffs_h.numel() * 4 + G_h.numel() * 4 +
                                          (4 * 3 * L + 8 * L if e2e_api.startswith("eager") else 0)),
                "d2h_bytes_per_step": int(psf_h.numel() * 4 + grad_h.numel() * 4)},
        "gpu_launches": int(launches),
        "parity": parity,
        "roofline": roof,
        "sustained": sus,
        "c4_strong": c4,
        "sparse": sparse_rec,
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "one full c3 PSF+grad, all 64 wavelengths (NumPy complex64 oracle transfer "
                                   "matrices + torch-CPU complex64 matmul/autograd on all host cores)"},
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sparse", type=int, default=1, help="also time the opt-in zero-block-skipping path (0 = skip)")
    ap.add_argument("--c4-stars", type=int, default=64,
                    help="stars of the fixed-size C4-lite strong-scaling record (0 = skip)")
    ap.add_argument("--sustained", type=float, default=3.0,
                    help="seconds of back-to-back steps for the `sustained` sub-record (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
