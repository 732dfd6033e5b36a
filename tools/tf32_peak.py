"""Measure the tcgen05 tensor-pipe rate the roofline of gemm_tc_kernel is quoted against: the
MMA-only probe of libdlux_b200.so (dlux_tc_peak_probe: cta_group::2, M256 x N128, operands resident)
for kind::tf32, kind::f16 (bf16) and the GEMM's own 4 tf32 + 4 bf16 mix; burst (one ~20 ms launch,
best of 5) and sustained (back-to-back launches for >= 4 s), with SM clock / power sampled during
the sustained run.  Writes profiles/tf32_peak.json (or the path given).

    python tools/tf32_peak.py [out.json]
"""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dlux_b200 import _lib  # noqa: E402


def sample_smi(stop, rows, idx=0):
    q = "clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown"
    p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(idx)],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    def rd():
        for line in p.stdout:
            rows.append([c.strip() for c in line.split(",")])
    t = threading.Thread(target=rd, daemon=True)
    t.start()
    stop.wait()
    p.terminate()


def measure(kind, seconds=4.0):
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream(dev).cuda_stream
    sink = torch.zeros(1, device=dev)
    import ctypes as C
    sp = C.c_void_p(sink.data_ptr())
    # calibrate: batches for ~20 ms
    nb = 2000
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fl = _lib.tc_peak_probe(kind, nb, st, sp)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        nb = max(200, int(nb * 20.0 / ms))
    burst = 0.0
    for _ in range(5):
        torch.cuda.synchronize()
        time.sleep(0.5)                     # let the board cool to its idle clocks between bursts
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fl = _lib.tc_peak_probe(kind, nb, st, sp)
        e1.record()
        torch.cuda.synchronize()
        burst = max(burst, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    rows, stop = [], threading.Event()
    th = threading.Thread(target=sample_smi, args=(stop, rows), daemon=True)
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    tot = 0.0
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(10):
            tot += _lib.tc_peak_probe(kind, nb, st, sp)
        torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    sus = tot / (e0.elapsed_time(e1) * 1e-3) / 1e12
    stop.set()
    th.join(timeout=3)
    clk = sorted(float(r[0]) for r in rows[len(rows) // 4:] if r and r[0].replace(".", "").isdigit())
    pw = sorted(float(r[1]) for r in rows[len(rows) // 4:] if len(r) > 1 and r[1].replace(".", "").isdigit())
    cap = any(len(r) > 2 and r[2].lower().startswith("active") for r in rows)
    return {"burst_tflops": burst, "sustained_tflops": sus, "sustained_seconds": seconds,
            "sm_mhz_median_sustained": clk[len(clk) // 2] if clk else None,
            "power_w_median_sustained": pw[len(pw) // 2] if pw else None, "sw_power_cap": cap,
            "batches_per_launch": nb, "value_check": float(sink.item())}


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "tf32_peak.json")
    _lib.load()
    res = {"what": "tcgen05.mma.cta_group::2 M256xN128, A in TMEM, B in smem, operands resident, 74 CTA pairs; "
                   "real FLOPs = 2*M*N*K per MMA; CUDA events around the launch(es)",
           "gpu": torch.cuda.get_device_name(0), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "kind_tf32_k8": measure(0), "kind_f16_bf16_k16": measure(1), "gemm_mix_4tf32_4bf16": measure(2)}
    m = res["gemm_mix_4tf32_4bf16"]
    # the GEMM's algorithmic FLOPs per mix FLOP: 8 MMAs (4 x K8 + 4 x K16 = 96 k-units of real work) realise
    # 16 k of a complex product's 4 real MACs... quoted simply: algorithmic = mix_flops / 3
    res["algorithmic_ceiling_tflops"] = {"burst": m["burst_tflops"] / 3.0, "sustained": m["sustained_tflops"] / 3.0,
                                         "note": "per 16-k chunk and tile the GEMM issues 4 tf32 MMAs (K=8) + 4 bf16 MMAs "
                                                 "(K=16): 3x the real FLOPs of the algorithmic complex product "
                                                 "(2*256*128*16*2 per chunk = one tf32-precision pass)"}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
