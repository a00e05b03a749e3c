"""Graph-timed microbenchmarks of single launches of the step (one B200): each candidate is captured REPS times in a
CUDA graph (so host launch overhead is out) and the graph is timed with CUDA events.  Numbers are warm-L2 per-launch
times -- use them to choose between variants, not as bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deeprl_b200 as rl  # noqa: E402
from deeprl_b200 import _lib, ops  # noqa: E402
from deeprl_b200.network import fused  # noqa: E402

REPS = 20
dev = torch.device("cuda", 0)
rl.select_device(0)


def timed(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (10 * REPS)
    print("%-58s %8.2f us" % (name, us), flush=True)
    return us


B = 512
bf = torch.bfloat16
y3 = torch.randn(B, 3136, device=dev).to(bf)
w4 = (torch.randn(512, 3136, device=dev) * 0.02).to(bf)
b4 = torch.randn(512, device=dev)
y4 = torch.empty(B, 512, device=dev, dtype=bf)
acc = torch.zeros(B, 512, device=dev)


def fc4_split(splits):
    def f():
        ops.gemm_bf16(y3, w4, out_dtype=torch.float32, out=acc, splits=splits, block_n=64)
        _lib.call("b2rl_bias_act_f32_to_bf16", _lib.ptr(acc), _lib.ptr(b4), _lib.ptr(y4), B, 512, 1, _lib.stream())
    return f


for s in (2, 4, 8):
    timed("fc4 fwd: split-K %d + zero + bias/ReLU pass" % s, fc4_split(s))
for bn in (32, 64, 128):
    timed("fc4 fwd: one kernel, fused bias+ReLU, BN=%d" % bn,
          lambda bn=bn: ops.gemm_bf16(y3, w4, bias=b4, relu=True, out=y4, block_n=bn))

# fc4 backward GEMMs
g4 = torch.randn(B, 512, device=dev).to(bf)
gw = torch.empty(512, 3136, device=dev)
gy3 = torch.empty(B, 3136, device=dev, dtype=bf)
for bn in (64, 128):
    timed("fc4 wgrad [512x3136] K=512, BN=%d" % bn,
          lambda bn=bn: ops.gemm_bf16(g4, y3, a_major="mn", b_major="mn", out_dtype=torch.float32, out=gw, block_n=bn))
    timed("fc4 dgrad [512x3136] K=512, BN=%d" % bn,
          lambda bn=bn: ops.gemm_bf16(g4, w4, a_major="k", b_major="mn", out=gy3, block_n=bn))

# narrow heads
phi = torch.randn(B, 512, device=dev).to(bf)
for A, duel in ((4, False), (6, True), (18, False)):
    wa, ba = torch.randn(A, 512, device=dev), torch.randn(A, device=dev)
    wv, bv = (torch.randn(1, 512, device=dev), torch.randn(1, device=dev)) if duel else (None, None)
    q = torch.empty(B, A, device=dev)
    gq = torch.randn(B, A, device=dev)
    gphi = torch.empty_like(phi)
    gwa, gba = torch.zeros_like(wa), torch.zeros_like(ba)
    gwv, gbv = (torch.zeros_like(wv), torch.zeros_like(bv)) if duel else (None, None)
    timed("head_fwd A=%d dueling=%s" % (A, duel),
          lambda: _lib.call("b2rl_head_fwd", _lib.ptr(phi), _lib.ptr(wa), _lib.ptr(ba), _lib.ptr(wv), _lib.ptr(bv), B, 512, A,
                            _lib.ptr(q), _lib.stream()))
    timed("head_bwd A=%d dueling=%s" % (A, duel),
          lambda: _lib.call("b2rl_head_bwd", _lib.ptr(gq), _lib.ptr(phi), _lib.ptr(wa), _lib.ptr(wv), B, 512, A, _lib.ptr(gphi),
                            _lib.ptr(gwa), _lib.ptr(gba), _lib.ptr(gwv), _lib.ptr(gbv), _lib.stream()))

# optimizer tail on the DQN parameter count
n = 1_687_000
opt = ops.FlatOptimizer([torch.nn.Parameter(torch.randn(n, device=dev))], kind="rmsprop", lr=2.5e-4, alpha=0.95, eps=0.01,
                        centered=True)
opt.grad.normal_()
timed("clip + RMSprop(centered) over %d params (2 launches)" % n, lambda: opt.step(max_norm=10.0))
print("done")

# ---------------------------------------------------------------------------------------------- convolution GEMMs
import ctypes  # noqa: E402
from deeprl_b200.network import nature_tc  # noqa: E402

x0 = torch.randint(0, 255, (B * 441, 64), device=dev).to(bf)
w1 = (torch.randn(32, 256, device=dev) * 0.01).to(bf)
b1 = torch.randn(32, device=dev)
x1 = torch.empty(B * 100, 128, device=dev, dtype=bf)
w2 = (torch.randn(64, 512, device=dev) * 0.01).to(bf)
b2 = torch.randn(64, device=dev)
y2 = torch.empty(B * 100, 64, device=dev, dtype=bf)
w3 = (torch.randn(64, 576, device=dev) * 0.01).to(bf)
y3b = torch.empty(B * 49, 64, device=dev, dtype=bf)
g1 = torch.randn(B * 441, 32, device=dev).to(bf)
g2 = torch.randn(B * 100, 64, device=dev).to(bf)
w2d = (torch.randn(128, 256, device=dev) * 0.01).to(bf)
gy1 = torch.empty(B * 100, 128, device=dev, dtype=bf)
timed("conv1 fwd", lambda: nature_tc.conv_gemm(0, x0, w1, 32, 4, 2, 21, 1, x1, bias=b1, relu=True, out_map=1, G=21, V=20, block_n=32))
timed("conv2 fwd", lambda: nature_tc.conv_gemm(0, x1, w2, 64, 4, 2, 10, 1, y2, bias=b2, relu=True, block_n=64))
timed("conv3 fwd", lambda: nature_tc.conv_gemm(0, y2, w3, 64, 9, 3, 10, 1, y3b, bias=b2, relu=True, out_map=2, G=10, V=7, block_n=64))
timed("conv3 dgrad", lambda: nature_tc.conv_gemm(0, g2, w3, 64, 9, 3, 10, -1, y2, block_n=64))
timed("conv2 dgrad", lambda: nature_tc.conv_gemm(0, g2, w2d, 128, 4, 2, 10, -1, gy1, block_n=128))
timed("conv3 wgrad (partials)", lambda: nature_tc.wgrad_partials(y2, g2, 64, 9, 3, 10))
timed("conv2 wgrad (partials)", lambda: nature_tc.wgrad_partials(x1, g2, 64, 4, 2, 10))
timed("conv1 wgrad (partials)", lambda: nature_tc.wgrad_partials(x0, g1, 32, 4, 2, 21))

# dual launches: two operand sets in one grid, checked against two single launches
x0b = torch.randint(0, 255, (B * 441, 64), device=dev).to(bf)
w1b = (torch.randn(32, 256, device=dev) * 0.01).to(bf)
b1b = torch.randn(32, device=dev)
x1a, x1b, x1r = torch.zeros_like(x1), torch.zeros_like(x1), torch.zeros_like(x1)


def dual_conv1():
    _lib.call("b2rl_conv_gemm_dual_bf16", _lib.ptr(x0), _lib.ptr(x0b), B * 441, 64, _lib.ptr(w1), _lib.ptr(w1b), 32, 4, 2, 21, 1,
              _lib.ptr(x1a), _lib.ptr(x1b), 128, _lib.ptr(b1), _lib.ptr(b1b), 1, 0, 1, 21, 20, 32, _lib.stream())


dual_conv1()
nature_tc.conv_gemm(0, x0b, w1b, 32, 4, 2, 21, 1, x1r, bias=b1b, relu=True, out_map=1, G=21, V=20, block_n=32)
nature_tc.conv_gemm(0, x0, w1, 32, 4, 2, 21, 1, x1, bias=b1, relu=True, out_map=1, G=21, V=20, block_n=32)
torch.cuda.synchronize()
print("dual conv1 == two single launches:", bool(torch.equal(x1a, x1)), bool(torch.equal(x1b, x1r)))
timed("conv1 fwd DUAL (both nets)", dual_conv1)
y4a, y4b = torch.empty(B, 512, device=dev, dtype=bf), torch.empty(B, 512, device=dev, dtype=bf)
y3x = torch.randn(B, 3136, device=dev).to(bf)
w4x = (torch.randn(512, 3136, device=dev) * 0.02).to(bf)


def dual_fc4(bn):
    _lib.call("b2rl_gemm_dual_bf16", _lib.ptr(y3), _lib.ptr(y3x), 3136, _lib.ptr(w4), _lib.ptr(w4x), 3136, _lib.ptr(y4a),
              _lib.ptr(y4b), 512, B, 512, 3136, _lib.ptr(b4), _lib.ptr(b4), 1, 0, bn, _lib.stream())


dual_fc4(64)
ref_a = ops.gemm_bf16(y3, w4, bias=b4, relu=True, block_n=64)
ref_b = ops.gemm_bf16(y3x, w4x, bias=b4, relu=True, block_n=64)
torch.cuda.synchronize()
print("dual fc4 == two single launches:", bool(torch.equal(y4a, ref_a)), bool(torch.equal(y4b, ref_b)))
for bn in (32, 64, 128):
    timed("fc4 fwd DUAL, fused bias+ReLU, BN=%d" % bn, lambda bn=bn: dual_fc4(bn))
print("done")
