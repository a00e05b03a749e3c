"""Per-tile timeline of the slab convolution kernel (csrc/gemm.cu) on one B200: builds a PRIVATE copy of the library with
-DB2RL_TRACE (clock64 stamps per CTA / role / tile), runs the three forward convolutions of NatureConvBody at B = 512 and
prints, for a few CTAs, what the TMA producer, the MMA issuer and one epilogue warp were waiting for.  Diagnostic only."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from deeprl_b200 import _lib  # noqa: E402

TRACE_LIB = os.path.join(ROOT, "scripts", "microbench", "libb2rl_trace.so")


def build():
    cmd = ["nvcc"] + _lib.NVCC_FLAGS + ["-DB2RL_TRACE", "-o", TRACE_LIB] + [os.path.join(_lib.CSRC, s) for s in _lib.SOURCES]
    subprocess.run(cmd, check=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
        sys.exit(0)
    if not os.path.exists(TRACE_LIB):
        build()
    _lib.LIB_PATH = TRACE_LIB
    import deeprl_b200 as rl
    from deeprl_b200.network import nature_tc
    rl.select_device(0)
    dev = torch.device("cuda", 0)
    B = 512
    bf = torch.bfloat16
    trace = torch.zeros(148 * 4 * 16 * 4, dtype=torch.int64, device=dev)
    cases = {
        "conv1 fwd (BN=32, 1764 tiles)": lambda: nature_tc.conv_gemm(
            0, x0, w1, 32, 4, 2, 21, 1, x1, bias=b1, relu=True, out_map=1, G=21, V=20, block_n=32),
        "conv2 fwd (BN=64, 400 tiles)": lambda: nature_tc.conv_gemm(0, x1, w2, 64, 4, 2, 10, 1, y2, bias=b2, relu=True, block_n=64),
        "conv3 fwd (BN=64, 400 tiles)": lambda: nature_tc.conv_gemm(
            0, y2, w3, 64, 9, 3, 10, 1, y3, bias=b2, relu=True, out_map=2, G=10, V=7, block_n=64),
    }
    g1 = torch.randn(B * 441, 32, device=dev).to(bf)
    g2 = torch.randn(B * 100, 64, device=dev).to(bf)
    cases["conv1 wgrad (147 CTAs x ~24 k-tiles)"] = lambda: nature_tc.wgrad_partials(x0, g1, 32, 4, 2, 21)
    cases["conv2 wgrad"] = lambda: nature_tc.wgrad_partials(x1, g2, 64, 4, 2, 10)
    x0 = torch.randint(0, 255, (B * 441, 64), device=dev).to(bf)
    w1 = (torch.randn(32, 256, device=dev) * 0.01).to(bf)
    b1 = torch.zeros(32, device=dev)
    x1 = torch.empty(B * 100, 128, device=dev, dtype=bf)
    w2 = (torch.randn(64, 512, device=dev) * 0.01).to(bf)
    b2 = torch.zeros(64, device=dev)
    y2 = torch.empty(B * 100, 64, device=dev, dtype=bf)
    w3 = (torch.randn(64, 576, device=dev) * 0.01).to(bf)
    y3 = torch.empty(B * 49, 64, device=dev, dtype=bf)
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        # the library instance that torch-side wrappers call is the same file: set the trace buffer through it
        L = _lib.lib()
        L.b2rl_debug_set_trace.argtypes = [ctypes.c_void_p]
        L.b2rl_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
        trace.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = trace.cpu().numpy().reshape(148, 4, 16, 4).astype(np.int64)
        print("=" * 100)
        print("%s: %.1f us (events, incl. launch)" % (name, e0.elapsed_time(e1) * 1e3))
        for cta in (0, 73, 147):
            c = t[cta]
            t0 = c[3, 0, 0]
            if t0 == 0:
                continue
            cyc = lambda v: (v - t0) if v else -1
            print("  CTA %3d: kernel body start 0, producer start %d, weights resident %d, end %d cycles" % (
                cta, cyc(c[3, 0, 1]), cyc(c[3, 0, 2]), cyc(c[3, 0, 3])))
            print("    tile | producer: wait-empty issue | MMA: start tmem-free slab-full committed | epilogue: start acc-full tmem-read stored")
            for it in range(16):
                if c[1, it, 0] == 0:
                    break
                print("    %4d | %8d %8d | %8d %8d %8d %8d | %8d %8d %8d %8d" % (
                    it, cyc(c[0, it, 0]), cyc(c[0, it, 1]), cyc(c[1, it, 0]), cyc(c[1, it, 1]), cyc(c[1, it, 2]), cyc(c[1, it, 3]),
                    cyc(c[2, it, 0]), cyc(c[2, it, 1]), cyc(c[2, it, 2]), cyc(c[2, it, 3])))
    print("done")
