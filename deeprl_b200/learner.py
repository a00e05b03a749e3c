"""CUDA-graph learner for the DQN family: one gradient update of ``DQNAgent.step`` (DQN_agent.py:101-138) --
``sgd_update_frequency`` feeds, sample, target / online forward, fused loss, backward, [gradient all-reduce],
fused clip + optimizer, PER priority update -- replayed as (at most two) captured graphs, with every scalar
that changes between updates (ring cursor, Philox counter, PER beta, Adam step, max priority) living in device
memory.  This is the throughput path of bench.py; ``DQNAgent`` without it runs the same kernels eagerly.

Multi-GPU: one learner per rank, rank-local replay shard, parameters identical on every rank;
gradients are summed with ONE NCCL all-reduce of the flat gradient arena between the two graphs and scaled by
1 / world_size inside the clip kernel.
"""
import os
import sys

import numpy as np
import torch

from . import _lib, ops, parallel
from .component.replay import PrioritizedReplay
from .network import nature_tc
from .network.fused import frame_scale
from .utils.config import Config


class StepTrace:
    """Timing events recorded INSIDE the captured update (``torch.cuda.Event(external=True)`` -> event-record nodes of the
    graph): the real timeline of one graph replay, with the overlap between the branches, which neither the serialised ncu
    launch list nor eager launches show.  ``scripts/trace_step.py`` prints it."""

    def __init__(self):
        self.marks = []

    def mark(self, name, stream=None):
        e = torch.cuda.Event(enable_timing=True, external=True)
        e.record(stream if stream is not None else torch.cuda.current_stream())
        self.marks.append((name, e))

    def timeline(self):
        t0 = self.marks[0][1]
        return [(n, t0.elapsed_time(e) * 1e3) for n, e in self.marks]


class GraphedDQNLearner:
    def __init__(self, network, target_network, optimizer, replay, kind="dqn", discount=0.99, n_step=1, double_q=False,
                 gradient_clip=5.0, feeds_per_update=4, compute_dtype=torch.bfloat16, state_scale=1.0 / 255,
                 replay_eps=0.01, replay_alpha=0.5, categorical=(-10.0, 10.0), world_size=1, target_sync_every=10000,
                 prefetch=False, dual=False):
        self.net, self.tgt, self.opt, self.replay = network, target_network, optimizer, replay
        self.kind, self.gamma_n, self.double_q = kind, discount ** n_step, double_q
        self.clip, self.feeds = gradient_clip, feeds_per_update
        self.dtype, self.scale = compute_dtype, state_scale
        self.eps, self.alpha, self.cat = replay_eps, replay_alpha, categorical
        self.world = world_size
        self.per = isinstance(replay, PrioritizedReplay)
        self.sync_every = target_sync_every
        dev = replay.device
        self.dev = dev
        B = replay.batch_size
        self.B = B
        n = max(self.feeds, 1)
        rb = replay.row_bytes
        # ONE packed pinned staging buffer for the env transitions of an update (+ PER beta) and ONE device mirror:
        # [frames n*rb | reward n*8 | action n*4 | mask n*4 | beta 4] -> a single host->device copy node in the graph
        o_r = (n * rb + 7) // 8 * 8
        o_a, o_m, o_b = o_r + 8 * n, o_r + 12 * n, o_r + 16 * n
        total = (o_b + 4 + 15) // 16 * 16
        self.h_pack = torch.zeros(total, dtype=torch.uint8, pin_memory=True)
        self.d_pack = torch.zeros(total, dtype=torch.uint8, device=dev)

        def views(buf):
            return (buf[:n * rb].view(n, rb), buf[o_a:o_a + 4 * n].view(torch.int32), buf[o_r:o_r + 8 * n].view(torch.float64),
                    buf[o_m:o_m + 4 * n].view(torch.int32), buf[o_b:o_b + 4].view(torch.float32))

        self.h_frames, self.h_action, self.h_reward, self.h_mask, self.h_beta = views(self.h_pack)
        self.d_frames, self.d_action, self.d_reward, self.d_mask, self.d_beta = views(self.d_pack)
        self.h_mask.fill_(1), self.h_beta.fill_(0.4)
        self.d_pack.copy_(self.h_pack)
        self.h_loss = torch.zeros(1, dtype=torch.float32, pin_memory=True)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.g_main = self.g_opt = None
        self._side = None
        # prefetch = the graph form of ReplayWrapper(async_=True) (replay.py:214-262: the worker hands out the batch it
        # sampled right after the previous request and immediately samples the next one into the other cache): update k
        # trains on the batch sampled during update k-1 while a third branch of the graph feeds + samples batch k+1
        # into the other buffer set.  Two graphs (one per buffer parity) are captured and replayed alternately.
        self.prefetch = bool(prefetch)
        # dual: one launch per body layer for online(s) + target(s') (nature_tc.forward_dual).  Measured on B200 at B = 512:
        # 27 instead of 33 launches per update, +1 % updates/s with synchronous replay, -2 % with the prefetch branch (the
        # two-stream fork already hides the per-launch fixed cost), hence off by default.
        self.dual = bool(dual)
        self._batch = [None, None]
        self._parity = 0
        self.updates = 0
        self.with_h2d = False
        self._tail = None                 # network/tail.py NatureTail (built on first use), False = not applicable
        self._overlap = False             # multi-GPU: NCCL captured inside the update graph, fc4's all-reduce beside the backward
        self._early_work = None
        body = getattr(network, "body", None)
        self.ring = (not self.prefetch and not self.dual and compute_dtype == torch.bfloat16 and Config.DENSE_BACKEND == "tcgen05"
                     and hasattr(body, "repack") and not getattr(body, "noisy_linear", False) and replay.history_length == 4
                     and tuple(getattr(replay, "item_shape", ())) == (84, 84) and os.environ.get("B2RL_K1", "1") != "0")
        self.opt.zero_grad()              # the fused tail writes / re-zeroes the gradient arena itself: start from zeros

    # ------------------------------------------------------------------ fused tail / fused head (csrc/tail.cu, csrc/head.cu)
    def tail(self):
        """The two-launch update tail (gradient reduce + clip / optimizer / operand pack) when the online network has a
        tcgen05 NatureConvBody and the backward epilogues are fused; None otherwise (generic unpack + FlatOptimizer.step)."""
        if self._tail is None:
            from .network.tail import NatureTail
            body = getattr(self.net, "body", None)
            ok = (self.dtype == torch.bfloat16 and Config.DENSE_BACKEND == "tcgen05" and nature_tc.FUSED_BWD and _lib.CONV_SLAB
                  and hasattr(body, "repack") and not getattr(body, "noisy_linear", False)
                  and os.environ.get("B2RL_TAIL", "1") != "0" and self.opt.kind in ("rmsprop", "adam"))
            if ok:
                self._repack(self.net, self.scale)
                self._tail = NatureTail(self.opt, body, self.scale)
                self._tail.max_norm, self._tail.grad_scale = self.clip, 1.0 / self.world
                self._refresh_head_operands(True)
                if self.world > 1:
                    # fc4's gradient is reduced into the arena and all-reduced right after its GEMM (beside the convolution
                    # backward); the small remainder follows the last weight-gradient GEMM
                    self._tail.split = True
                    self._tail.early = self._allreduce_early
            else:
                self._tail = False
        return self._tail or None

    def _heads(self):
        """(online head modules, target head modules) when the DQN head can run in the fused head + loss + backward kernel."""
        # (off by default: measured on B200 at batch 512 the separate head_fwd | dqn_loss | head_bwd kernels -- the target head on
        # the side branch -- give 224.7 us / update, the two-launch fused form 243.2 us, the one-launch form 252.9 us)
        if self.kind != "dqn" or self.tail() is None or os.environ.get("B2RL_FUSED_HEAD", "0") == "0":
            return None
        out = []
        for n in (self.net, self.tgt):
            fa = getattr(n, "fc_head", None) or getattr(n, "fc_advantage", None)
            fv = getattr(n, "fc_value", None) if hasattr(n, "fc_advantage") else None
            if not isinstance(fa, torch.nn.Linear) or fa.out_features >= 32 or (fv is not None and not isinstance(fv, torch.nn.Linear)):
                return None
            if any(m is not None and m.weight.data_ptr() % 16 for m in (fa, fv)):
                return None
            out.append((fa, fv))
        return out

    # ------------------------------------------------------------------ the update, as eager code
    def _h2d(self):
        self.d_pack.copy_(self.h_pack, non_blocking=True)

    def _sample(self, tag=0, phase=None):
        """feeds of this update + one sampled batch into buffer set ``tag`` (on the current stream).  ``phase`` "select" =
        feed + index draw only, "gather" = the batch from those indices (uniform replay), None = everything."""
        rp = self.replay
        # DQN_agent.py:104-112 calls feed() once per env transition; `feeds` single-item calls are exactly one multi-item
        # call with each item in its own slot (reference_feed_quirk off) plus `feeds` tree.add(max_priority) -- one launch
        if self.feeds and phase != "gather":
            quirk, rp.quirk = rp.quirk, False
            rp._device_cursor = True                     # (a captured feed advances the device cursor only)
            if self.per:
                rp.feed_device(self.d_frames, self.d_action, self.d_reward, self.d_mask, self.feeds, add_leaf=False)
                rp.tree.add_n(self.feeds, rp.max_priority_dev)
            else:
                rp.feed_device(self.d_frames, self.d_action, self.d_reward, self.d_mask, self.feeds)
            rp.quirk = quirk
        if self.dtype == torch.bfloat16:
            # exact integer frames, space-to-depth layout; ImageNormalizer's scale is folded into conv1's weights.
            # K1 (self.ring): no batch at all -- conv1 reads the sampled stacks from the uint8 ring (synchronous replay only:
            # a prefetched index could be overwritten by the next update's feeds before its frames are read)
            kw = dict(phase=phase) if phase else {}
            return rp.sample_normalized(out_dtype=self.dtype, scale=None, layout="ring" if self.ring else "s2d", tag=tag, **kw)
        kw = dict(phase=phase) if phase else {}
        return rp.sample_normalized(out_dtype=self.dtype, scale=self.scale, layout="nchw", tag=tag, **kw)

    def _main(self, parity=None):
        rp = self.replay
        cur = torch.cuda.current_stream()
        nature_tc.mark("start")
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
            self._pre = torch.cuda.Stream(device=self.dev)
            self._packed_ev, self._sampled_ev = torch.cuda.Event(), torch.cuda.Event()
        side, pre = self._side, self._pre
        if os.environ.get("B2RL_SINGLE_STREAM", "0") == "1":       # experiment: no parallel graph branches except the prefetch
            side = cur
        fs = self.scale if self.dtype == torch.bfloat16 else 1.0
        tail = self.tail()
        if tail is None:
            # online weights changed in the previous optimizer step: re-pack them on the side branch, next to feed + sample
            # (with the fused tail the optimizer kernel itself writes the packed bf16 operands)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._repack(self.net, fs)
                self._packed_ev.record(side)
        # async replay: the batch of the NEXT update is fed + sampled on a parallel branch.  Uniform replay: that branch starts
        # after the backward pass, beside the (small-footprint, L2-bound) update tail -- started beside the forward pass its 512
        # gather CTAs held the shared memory the convolution kernels need and delayed them by ~20 us (in-graph timeline,
        # profiles/r02_timeline.txt).  Prioritized replay keeps the early start: the reference's replay worker draws the next
        # batch BEFORE this update's priorities arrive (replay.py:219-261), and the graph keeps that order.
        late = (self.prefetch and not self.per and getattr(self, "_one_graph", True)
                and os.environ.get("B2RL_PREFETCH_LATE", "1") == "1")
        if self.prefetch:
            eager = parity is None
            if eager:
                parity = self._parity
            if self._batch[parity] is None:                              # very first update: nothing prefetched yet
                self._batch[parity] = self._sample(parity)
            t = self._batch[parity]
            if late:
                self._prefetch_branch(parity, "select")  # feed + index draw now (two one-CTA kernels), the gather later
            else:
                self._prefetch_branch(parity)
            if eager:
                self._parity = 1 - parity
        else:
            t = self._sample(0)
        per = dict(is_prob=t.sampling_prob, eps=self.eps, alpha=self.alpha, beta_dev=self.d_beta) if self.per else {}
        heads = self._heads()
        if heads is not None:
            self._main_fused_head(t, per, heads, tail, fs)
            if late:
                self._prefetch_branch(parity, "gather")
            if self.prefetch:
                cur.wait_stream(pre)
            return
        body_a, body_b = getattr(self.net, "body", None), getattr(self.tgt, "body", None)
        if (self.dual and self.dtype == torch.bfloat16 and Config.DENSE_BACKEND == "tcgen05"
                and hasattr(body_a, "repack") and hasattr(body_b, "repack")):
            # online(s) and target(s') share every launch of the convolutional body (nature_tc.forward_dual): the grid of
            # each kernel is split between the two networks, so the per-launch fixed cost is paid once
            if tail is None:
                cur.wait_event(self._packed_ev)
            with nature_tc.dual_forward(body_a, body_b, t.next_state), frame_scale(fs):
                out = self.net(t.state)
                with torch.no_grad():
                    nxt_t = self.tgt(t.next_state)
                    nxt_o = self.net(t.next_state) if self.double_q else None
        else:
            # the target forward on s' and the online forward on s are independent: fork them onto two streams (two
            # parallel branches of the captured graph) so that the prologue / tail of one chain overlaps the other
            side.wait_stream(cur)
            with torch.cuda.stream(side), frame_scale(fs), torch.no_grad():
                nxt_t = self.tgt(t.next_state)
            if tail is None:
                cur.wait_event(self._packed_ev)
            with frame_scale(fs):
                with torch.no_grad():
                    nxt_o = self.net(t.next_state) if self.double_q else None
                out = self.net(t.state)
            cur.wait_stream(side)
        nature_tc.mark("fwd_joined")
        if self.kind == "dqn":
            head = out["q"]
            r = ops.dqn_loss_fused(head.detach(), nxt_t["q"], nxt_o["q"] if nxt_o else None, t.action, t.reward, t.mask,
                                   self.gamma_n, **per)
            grad = r["dq"]
        elif self.kind == "c51":
            head = out["log_prob"]
            r = ops.c51_loss_fused(head.detach(), nxt_t["prob"], nxt_o["prob"] if nxt_o else None, t.action, t.reward,
                                   t.mask, self.gamma_n, self.cat[0], self.cat[1], **per)
            grad = r["dlogp"]
        else:
            head = out["quantile"]
            r = ops.qr_loss_fused(head.detach(), nxt_t["quantile"], t.action, t.reward, t.mask, self.gamma_n)
            grad = r["dquant"]
        if self.per:
            if self.prefetch:                            # the sum tree is read by the prefetch branch: update after it
                cur.wait_event(self._sampled_ev)
            rp.update_priorities((t.idx, r["priority"]))
        nature_tc.mark("loss")
        if tail is None:
            self.opt.zero_grad()
        fired = []
        if late and os.environ.get("B2RL_PREFETCH_AT", "end") == "dgrad":
            # option (B2RL_PREFETCH_AT=dgrad): fork the gather right after the last dgrad GEMM, beside the conv2 / conv1
            # weight-gradient GEMMs.  Measured 227.6 us / update against 226.1 us for the fork after the backward pass (default):
            # the gather slows the conv1 weight gradient and kernel A by as much as it gains
            nature_tc.AFTER_DGRAD = lambda: (self._prefetch_branch(parity, "gather"), fired.append(1))
        try:
            with nature_tc.wgrad_stream(None if side is cur else side), nature_tc.grad_sink(tail):   # weight-gradient GEMMs on the side branch
                head.backward(grad)
        finally:
            nature_tc.AFTER_DGRAD = None
        nature_tc.mark("bwd_done")
        self.loss.copy_(r["loss"])
        if late:
            if not fired:
                self._prefetch_branch(parity, "gather")
            self._late_join = True                       # joined after the optimizer kernels (_opt)
        elif self.prefetch:
            cur.wait_stream(pre)

    def _prefetch_branch(self, parity, phase=None):
        cur, pre = torch.cuda.current_stream(), self._pre
        pre.wait_stream(cur)                                             # after the host->device copy of this update's feeds
        with torch.cuda.stream(pre):
            b = self._sample(1 - parity, phase)
            if phase != "select":
                self._batch[1 - parity] = b
                self._sampled_ev.record(pre)
                nature_tc.mark("sampled")

    def _main_fused_head(self, t, per, heads, tail, fs):
        """DQN with a VanillaNet / DuelingNet head: bodies on the tcgen05 kernels, then ONE launch for the online / target
        [/ double-Q] head forwards + target / loss / PER block + head backward (csrc/head.cu dqn_head_fused_kernel)."""
        cur, side = torch.cuda.current_stream(), self._side
        side.wait_stream(cur)
        with torch.cuda.stream(side), frame_scale(fs), torch.no_grad():
            phi_t = self.tgt.body(t.next_state)
        with frame_scale(fs):
            with torch.no_grad():
                phi_o = self.net.body(t.next_state) if self.double_q else None
            phi = self.net.body(t.state)
        cur.wait_stream(side)
        nature_tc.mark("fwd_joined")
        r = ops.dqn_head_fused(phi.detach(), phi_t, phi_o, heads[0], heads[1], t.action, t.reward, t.mask, self.gamma_n,
                               tail.db4, two=os.environ.get("B2RL_FUSED_HEAD", "0") != "one", **per)
        if self.per:
            if self.prefetch:                            # the sum tree is read by the prefetch branch: update after it
                cur.wait_event(self._sampled_ev)
            self.replay.update_priorities((t.idx, r["priority"]))
        gphi = r["gphi"]
        nature_tc.mark("head")
        nature_tc.PREMASKED[gphi.data_ptr()] = tail.db4  # already masked by relu(fc4); its column sums are in the tail's db4
        with nature_tc.wgrad_stream(side), nature_tc.grad_sink(tail):
            phi.backward(gphi)
        self.loss.copy_(r["loss"])

    def _repack(self, net, fs):
        """tcgen05 backend: the learner owns the packed bf16 operands of both networks -- the online body is re-packed
        once per update (one launch), the target body only when it is synchronised."""
        body = getattr(net, "body", None)
        if body is not None and hasattr(body, "repack") and self.dtype == torch.bfloat16:
            body.auto_repack = False
            body.repack(fs)

    def sync_target(self):
        self.tgt.load_state_dict(self.net.state_dict())        # DQN_agent.py:136-138
        self._repack(self.tgt, self.scale if self.dtype == torch.bfloat16 else 1.0)
        self._refresh_head_operands(False)

    def _opt(self):
        self._opt_kernels()
        if getattr(self, "_late_join", False):           # the late prefetch branch ran beside the update tail
            torch.cuda.current_stream().wait_stream(self._pre)
            self._late_join = False

    def _opt_kernels(self):
        tail = self.tail()
        if tail is not None:
            tail.step(max_norm=self.clip, grad_scale=1.0 / self.world, reduced_elsewhere=self.world > 1)
            nature_tc.mark("opt")
        else:
            self.opt.step(max_norm=self.clip, grad_scale=1.0 / self.world)
            nature_tc.mark("opt")

    def refresh_packed(self):
        """Re-derive the packed bf16 operands of both networks from the fp32 parameters (after the parameters were changed
        from outside: load_state_dict, a broadcast, a copied arena)."""
        fs = self.scale if self.dtype == torch.bfloat16 else 1.0
        self._repack(self.net, fs)
        self._repack(self.tgt, fs)
        self._refresh_head_operands(True)

    def _dist_fc(self, net):
        return getattr(net, "fc_categorical", None) or getattr(net, "fc_quantiles", None)

    def _refresh_head_operands(self, online):
        """Distributional heads (C51 / QR-DQN) on the tcgen05 GEMM: the online head reads its bf16 weight from the optimizer's
        arena-wide bf16 shadow (written by the fused optimizer kernel), the target head from a copy refreshed at target sync."""
        if self.kind not in ("c51", "qr") or self.tail() is None or os.environ.get("B2RL_DIST_HEAD", "1") == "0":
            return
        fa, ft = self._dist_fc(self.net), self._dist_fc(self.tgt)
        if not isinstance(fa, torch.nn.Linear) or not isinstance(ft, torch.nn.Linear) or fa.weight.data_ptr() % 16:
            return
        o = self.opt
        if o.shadow is None:
            o.shadow = torch.zeros(o.n, dtype=torch.bfloat16, device=o.flat.device)
        if online:
            o.shadow.copy_(o.flat)
            off = (fa.weight.data_ptr() - o.flat.data_ptr()) // 4
            fa._w16 = o.shadow[off:off + fa.weight.numel()].view_as(fa.weight)
        if getattr(ft, "_w16", None) is None:
            ft._w16 = torch.empty_like(ft.weight, dtype=torch.bfloat16)
        ft._w16.copy_(ft.weight.detach())

    # ------------------------------------------------------------------ capture / replay
    def capture(self, warmup=3, with_h2d=False):
        """Warm up eagerly on a side stream (cuDNN autotune, lazy allocations), then capture."""
        self.with_h2d = with_h2d
        self.refresh_packed()
        # multi GPU, default: the collectives are captured INSIDE the update graph -- fc4's slice of the gradient arena (95 % of
        # the bytes) is all-reduced asynchronously right after its weight-gradient GEMM, beside the convolution backward, the
        # small remainder after the last GEMM.  B2RL_NCCL_IN_GRAPH=0: [sample..backward] graph | eager all-reduce of the whole
        # arena | [clip + optimizer] graph (the round-1 form: +54 us per update at 2 and 4 GPUs, +85 us at 8).
        one_graph = self.world == 1 or os.environ.get("B2RL_NCCL_IN_GRAPH", "1") == "1"
        self._one_graph = one_graph      # (the late prefetch branch is joined after the optimizer kernels: one graph only)
        self._overlap = self.world > 1 and one_graph and self.tail() is not None
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                if with_h2d:
                    self._h2d()
                self._main()
                self._allreduce()
                self._opt()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.g_main, self.g_opt = [], None
        try:
            self._capture_main(with_h2d, one_graph)
        except Exception as e:                            # noqa: BLE001 -- any capture error: use the split form
            if self.world == 1 or not one_graph:
                raise
            print("b2rl: NCCL capture failed (%s); using the split-graph form" % str(e).splitlines()[0], file=sys.stderr)
            torch.cuda.synchronize()
            one_graph = False
            self._one_graph = False
            self._overlap = False
            self.g_main = []
            self._capture_main(with_h2d, False)
        if not one_graph:                                # [sample..backward] | NCCL all-reduce | [clip+opt]
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt):
                self._opt()
        # host mirrors of the ring cursor advanced during warm-up / capture calls; re-derive from the device
        st = self.replay.ring_state.cpu()
        self.replay.pos, self.replay._size = int(st[0]), int(st[1])
        self.launches_per_update = None
        return self

    def _capture_main(self, with_h2d, one_graph):
        for parity in ((0, 1) if self.prefetch else (None,)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.g_main[0].pool() if self.g_main else None):
                if with_h2d:
                    self._h2d()
                self._main(parity)
                if one_graph:
                    self._allreduce()
                    self._opt()
            self.g_main.append(g)

    def _allreduce_early(self):
        """fc4's slice of the gradient arena, asynchronously (called on the weight-gradient branch right after kernel A4): the
        convolution backward proceeds while NCCL runs; ``_allreduce`` waits for it."""
        self._early_work = None
        if self._overlap:
            lo, hi = self._tail.early_slice
            self._early_work = parallel.allreduce_gradients(self.opt.grad[lo:hi], async_op=True)

    def _allreduce(self):
        if self.world <= 1:
            return
        tail = self.tail()
        if tail is not None and self._overlap:
            for lo, hi in tail.late_slices:
                parallel.allreduce_gradients(self.opt.grad[lo:hi])
            if getattr(self, "_early_work", None) is not None:
                self._early_work.wait()
                self._early_work = None
        else:
            parallel.allreduce_gradients(self.opt.grad)

    def update(self):
        """One gradient update (graph replay).  Returns the device loss tensor (no sync)."""
        if self.prefetch:
            self.g_main[self._parity].replay()
            self._parity = 1 - self._parity
        else:
            self.g_main[0].replay()
        if self.g_opt is not None:
            self._allreduce()
            self.g_opt.replay()
        self.updates += 1
        if self.sync_every and self.updates % self.sync_every == 0:
            self.sync_target()
        return self.loss

    def update_from_host(self, frames, action, reward, mask, beta=None):
        """End-to-end update through HOST buffers: the env transitions of this update are written to the pinned
        staging area, copied host->device inside the captured graph, and the loss is read back."""
        assert self.with_h2d, "capture(with_h2d=True) first"
        self.h_frames.numpy()[...] = frames
        self.h_action.numpy()[...] = action
        self.h_reward.numpy()[...] = reward
        self.h_mask.numpy()[...] = mask
        if beta is not None:
            self.h_beta[0] = beta
        self.update()
        self.h_loss.copy_(self.loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.h_loss[0])

    @property
    def h2d_bytes(self):
        return self.h_pack.numel()


class GraphedPPOLearner:
    """The PPO minibatch update (PPO_agent.py:73-99, non-shared representation) as ONE captured graph replayed once per
    minibatch: row gather by a device-resident index matrix, network forward, ``b2rl_ppo_loss`` (clipped surrogate, value
    loss, approx-KL and their gradients in one launch), backward, KL-GATED actor Adam step (``b2rl_clip_adam_gated``: the
    reference's ``if approx_kl <= 1.5 * target_kl`` decided on the device) and the critic Adam step.  An iteration of
    examples.py:496-522 is 5 120 such updates of an 11 k-parameter MLP: eager, each costs ~1.8 ms of Python and launch
    latency; as a graph replay the host cost is one ``cudaGraphLaunch``.

    The rollout rows live in persistent device buffers (``load``); the minibatch index rows of ALL epochs are uploaded
    once per iteration (``set_batches``) and consumed through a device cursor."""

    KEYS = ("state", "action", "log_pi_a", "ret", "advantage")

    def __init__(self, network, actor_opt, critic_opt, rows, state_dim, action_dim, mini_batch_size, ppo_ratio_clip,
                 entropy_weight, target_kl, max_batches):
        self.net, self.actor_opt, self.critic_opt = network, actor_opt, critic_opt
        self.mb, self.clip, self.ent_w, self.target_kl = int(mini_batch_size), ppo_ratio_clip, entropy_weight, target_kl
        dev = actor_opt.flat.device
        f = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        self.buf = dict(state=f(rows, state_dim), action=f(rows, action_dim), log_pi_a=f(rows, 1), ret=f(rows, 1),
                        advantage=f(rows, 1))
        self.perm = torch.zeros((max_batches, self.mb), dtype=torch.int64, device=dev)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=dev)
        self.stats = torch.zeros(4, dtype=torch.float32, device=dev)      # policy loss, value loss, approx KL of the last update
        self.graph = None
        self.dev = dev

    def load(self, entries):
        for k in self.KEYS:
            self.buf[k].copy_(getattr(entries, k).reshape(self.buf[k].shape))

    def set_batches(self, index_rows):
        rows = np.stack([np.asarray(r, dtype=np.int64) for r in index_rows])
        assert rows.shape[1] == self.mb and rows.shape[0] <= self.perm.shape[0]
        self.perm[:rows.shape[0]].copy_(torch.from_numpy(rows), non_blocking=False)
        self.cursor.zero_()
        return rows.shape[0]

    def _step(self):
        idx = self.perm.index_select(0, self.cursor).view(-1)
        e = {k: v.index_select(0, idx) for k, v in self.buf.items()}
        pred = self.net(e["state"], e["action"])
        r = ops.ppo_loss_fused(pred["log_pi_a"].detach(), pred["entropy"].detach(), pred["v"].detach(), e["log_pi_a"],
                               e["advantage"], e["ret"], self.clip, self.ent_w)
        shape = pred["v"].shape
        self.actor_opt.zero_grad()
        torch.autograd.backward([pred["log_pi_a"], pred["entropy"]], [r["dlogp"].view(shape), r["dent"].view(shape)])
        self.actor_opt.step(gate=r["out"][2:3], gate_max=1.5 * self.target_kl)        # PPO_agent.py:94-97
        self.critic_opt.zero_grad()
        pred["v"].backward(r["dv"].view(shape))
        self.critic_opt.step()                                                        # PPO_agent.py:98-99
        self.stats.copy_(r["out"])
        self.cursor.add_(1)

    def _state(self):
        return [t for o in (self.actor_opt, self.critic_opt) for t in (o.flat, o.s1, o.s2, o.step_dev)]

    def capture(self, warmup=3):
        """Warm-up + capture run on whatever is in the buffers; parameters and optimizer state are restored afterwards."""
        saved = [t.clone() for t in self._state()]
        validate = torch.distributions.Distribution._validate_args
        torch.distributions.Distribution.set_default_validate_args(False)     # argument checks synchronise: not capturable
        try:
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    self.cursor.zero_()
                    _lib.reset_launch_count()
                    self._step()
                    self.launches_per_update = _lib.launch_count()     # of OUR kernels (the MLP itself is torch / cuBLAS)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.cursor.zero_()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._step()
        finally:
            torch.distributions.Distribution.set_default_validate_args(validate)
        for t, v in zip(self._state(), saved):
            t.copy_(v)
        self.cursor.zero_()
        torch.cuda.synchronize()
        return self

    def run(self, n_batches):
        for _ in range(n_batches):
            self.graph.replay()


class PersistentPPOLearner:
    """The whole minibatch loop of a PPO iteration (PPO_agent.py:68-99, non-shared representation) as ONE launch of one
    persistent thread block (``b2rl_ppo_minibatch_updates``, csrc/ppo_persistent.cu): weights in shared memory, Adam moments in
    L2, the rows of the next minibatch fetched during the current update, the KL gate decided on the device.  Same interface as
    ``GraphedPPOLearner`` (which replays ~45 small kernels per update and remains the path for every network this kernel does
    not cover)."""

    KEYS = GraphedPPOLearner.KEYS
    A_NAMES = ("actor_body.layers.0.weight", "actor_body.layers.0.bias", "actor_body.layers.1.weight",
               "actor_body.layers.1.bias", "fc_action.weight", "fc_action.bias", "std")
    C_NAMES = ("critic_body.layers.0.weight", "critic_body.layers.0.bias", "critic_body.layers.1.weight",
               "critic_body.layers.1.bias", "fc_critic.weight", "fc_critic.bias")

    @staticmethod
    def supported(network, mini_batch_size):
        """GaussianActorCriticNet with DummyBody phi and two-layer tanh FCBody actor / critic bodies of equal widths."""
        from .network.network_bodies import DummyBody, FCBody
        from .network.network_heads import GaussianActorCriticNet
        if not isinstance(network, GaussianActorCriticNet) or not isinstance(network.phi_body, DummyBody):
            return False
        ab, cb = network.actor_body, network.critic_body
        if not all(isinstance(b, FCBody) and len(b.layers) == 2 and b.gate is torch.tanh and not b.noisy_linear for b in (ab, cb)):
            return False
        dims = lambda b: (b.layers[0].in_features, b.layers[0].out_features, b.layers[1].out_features)
        D, H1, H2 = dims(ab)
        A = network.fc_action.out_features
        if not (dims(cb) == (D, H1, H2) and D <= 256 and H1 <= 128 and H2 <= 128 and A <= 32 and 4 <= mini_batch_size <= 128
                and mini_batch_size % 4 == 0 and network.fc_action.weight.is_cuda):
            return False
        # weights + double-buffered rows + both networks' activations must fit the shared memory of one SM
        # (examples.py:496-522: D = 17, hidden 64, A = 6, mini batch 64 -> 205 KB)
        return int(_lib.lib().b2rl_ppo_minibatch_smem_bytes(D, A, H1, H2, int(mini_batch_size))) <= 227 * 1024

    def __init__(self, network, actor_opt, critic_opt, rows, state_dim, action_dim, mini_batch_size, ppo_ratio_clip,
                 entropy_weight, target_kl, max_batches):
        self.net, self.actor_opt, self.critic_opt = network, actor_opt, critic_opt
        self.mb, self.clip, self.ent_w, self.target_kl = int(mini_batch_size), ppo_ratio_clip, entropy_weight, target_kl
        if actor_opt.kind != "adam" or critic_opt.kind != "adam":
            raise NotImplementedError("the persistent PPO kernel implements Adam (examples.py:508-509)")
        dev = actor_opt.flat.device
        self.dev = dev
        f = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        self.buf = dict(state=f(rows, state_dim), action=f(rows, action_dim), log_pi_a=f(rows, 1), ret=f(rows, 1),
                        advantage=f(rows, 1))
        self.perm = torch.zeros((max_batches, self.mb), dtype=torch.int64, device=dev)
        self.stats = torch.zeros(4, dtype=torch.float32, device=dev)
        named = dict(network.named_parameters())
        self.a_off = self._offsets(actor_opt, [named[n] for n in self.A_NAMES])
        self.c_off = self._offsets(critic_opt, [named[n] for n in self.C_NAMES])
        ab = network.actor_body
        self.D, self.H1, self.H2 = ab.layers[0].in_features, ab.layers[0].out_features, ab.layers[1].out_features
        self.A = network.fc_action.out_features
        self.launches_per_update = 0.0
        self.n = 0

    @staticmethod
    def _offsets(opt, params):
        """Element offsets of ``params`` inside the optimizer's arena (host int32 array for the kernel's argument block)."""
        base = opt.flat.data_ptr()
        offs = []
        for p in params:
            off = (p.data_ptr() - base) // 4
            if not (0 <= off and off + p.numel() <= opt.n) or getattr(p, "_b2rl_flat_owner", None) != id(opt):
                raise _lib.B2RLError("PersistentPPOLearner: a parameter does not live in the optimizer's arena")
            offs.append(off)
        return torch.tensor(offs, dtype=torch.int32)

    def load(self, entries):
        for k in self.KEYS:
            self.buf[k].copy_(getattr(entries, k).reshape(self.buf[k].shape))

    def set_batches(self, index_rows):
        rows = np.stack([np.asarray(r, dtype=np.int64) for r in index_rows])
        assert rows.shape[1] == self.mb and rows.shape[0] <= self.perm.shape[0]
        self.perm[:rows.shape[0]].copy_(torch.from_numpy(rows), non_blocking=False)
        return rows.shape[0]

    def capture(self, warmup=0):
        return self

    def run(self, n_batches):
        a, c, b = self.actor_opt, self.critic_opt, self.buf
        _lib.call("b2rl_ppo_minibatch_updates", _lib.ptr(b["state"]), _lib.ptr(b["action"]), _lib.ptr(b["log_pi_a"]),
                  _lib.ptr(b["ret"]), _lib.ptr(b["advantage"]), self.D, self.A, self.H1, self.H2, self.mb, _lib.ptr(self.perm),
                  int(n_batches), _lib.ptr(a.flat), _lib.ptr(a.s1), _lib.ptr(a.s2), _lib.ptr(a.step_dev), _lib.ptr(self.a_off),
                  _lib.ptr(c.flat), _lib.ptr(c.s1), _lib.ptr(c.s2), _lib.ptr(c.step_dev), _lib.ptr(self.c_off),
                  float(a.lr), float(a.betas[0]), float(a.betas[1]), float(a.eps), float(c.lr), float(c.betas[0]),
                  float(c.betas[1]), float(c.eps), float(self.clip), float(self.ent_w), float(1.5 * self.target_kl),
                  _lib.ptr(self.stats), _lib.stream())
        self.n = int(n_batches)         # (the one launch is counted by _lib.call itself; launches_per_update stays 0)

