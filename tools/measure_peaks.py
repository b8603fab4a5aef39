"""Roofline denominators measured on the box the kernels run on -> gpurun_out/peaks_r2.json.

* cuBLAS GEMM 8192^3 through torch.matmul in TF32 (allow_tf32) and BF16: best of 10 (burst) and a 4 s back-to-back
  loop (sustained, the chip settles near 1.3 GHz under the 1 kW cap);
* the tcgen05 issue-rate peak of tools/micro/tc_peak.cu (M = 128 MMAs back to back on every SM for 5 ms and for 2 s);
* SM clock / throttle reasons sampled by nvidia-smi while each leg runs.
The 3xTF32 roofline of the Bartlett kernels is TF32 / 3 (three MMAs per product).  Copy the file to profiles/.
"""
from __future__ import annotations

import json
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


class Clocks:
    def __init__(self):
        self.samples, self.reasons, self._stop = [], set(), False
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
        while not self._stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append((float(out[0]), float(out[1]), float(out[2])))
                self.reasons.add(out[3].strip())
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop = True
        self._t.join()

    def summary(self):
        if not self.samples:
            return {}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz_median": sm[len(sm) // 2], "sm_mhz_min": sm[0], "sm_mhz_max": sm[-1], "sm_max_mhz": self.samples[0][1],
                "power_w_max": max(s[2] for s in self.samples), "reasons": sorted(self.reasons), "samples": len(sm)}


def gemm(dtype, tf32: bool, n=8192):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flop = 2.0 * n ** 3
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) * 1e-12)
        time.sleep(0.05)
    with Clocks() as ck:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0, k = time.time(), 0
        e0.record()
        while time.time() - t0 < 4.0:
            for _ in range(20):
                torch.matmul(a, b, out=c)
            k += 20
            torch.cuda.synchronize()
        e1.record(); e1.synchronize()
        sustained = flop * k / (e0.elapsed_time(e1) * 1e-3) * 1e-12
    torch.backends.cuda.matmul.allow_tf32 = False
    return {"burst_tflops": round(best, 1), "sustained_tflops": round(sustained, 1), "clocks_sustained": ck.summary()}


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    out["cublas_tf32_8192"] = gemm(torch.float32, True)
    time.sleep(2.0)
    out["cublas_bf16_8192"] = gemm(torch.bfloat16, False)
    time.sleep(2.0)
    exe = ROOT / "tools" / "micro" / "tc_peak"
    if exe.exists():
        with Clocks() as ck:
            r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
        rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
        out["tcgen05_micro"] = {"rows": rows, "clocks": ck.summary(), "note": "last row = 2 s sustained; others 5 ms bursts"}
        tf = [x for x in rows if x.get("kind") == "tf32" and x.get("a_operand") == "tmem"]
        if tf:
            out["tf32_micro_burst_tflops"] = max(x["tflops"] for x in tf[:-1])
            out["tf32_micro_sustained_tflops"] = tf[-1]["tflops"]
    else:
        out["tcgen05_micro"] = {"error": "tools/micro/tc_peak not built"}
    tfb = out["cublas_tf32_8192"]["burst_tflops"]
    out["roofline_3xtf32"] = {"cublas_burst_div3": round(tfb / 3, 1),
                              "cublas_sustained_div3": round(out["cublas_tf32_8192"]["sustained_tflops"] / 3, 1),
                              "micro_burst_div3": round(out.get("tf32_micro_burst_tflops", 0.0) / 3, 1),
                              "nominal_4096flop_clk_sm_div3": round(4096 * 148 * 1.965e9 / 3 * 1e-12, 1)}
    dst = ROOT / "gpurun_out" / "peaks_r2.json"
    dst.parent.mkdir(exist_ok=True)
    dst.write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
