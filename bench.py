#!/usr/bin/env python
"""Benchmark of the differentiable ball rollout (BASELINE.json metric: ball-steps/s, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A "step" is one pass of the hot path over one batch: forward rollout + L1 loss + backward
rollout of the whole workload (default: BASELINE configs[1] = C2, 4,096 balls x 201 segments x
500 simulation steps).  Prints ONE JSON line (rank 0).  See DESIGN.md §7 for every field.

  value     ball-steps/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public API (doper_b200.sim.ball_sim.vmapped_grad_and_value)
            with HOST numpy inputs: pinned H2D + 3 launches + pinned D2H inside the timed region
  roofline  dominant kernel (rollout_fwd_kernel): algorithmic FP32 flops / its event-timed duration
            against the non-FMA FP32 issue rate measured in this run (doper_fp32_probe)
  cpu_baseline / --impl reference   the oracle's C restatement of the reference (OpenMP, all host
            cores) — the reference itself is JAX, which is not installable in this image
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "ball_steps_per_sec_fwd_bwd"
UNIT = "ball-steps/s"


def workload_label(name, B, S, n_steps, per_env):
    scene = f"{S} segments per env ({B} private maps)" if per_env else f"one shared {S}-segment map"
    return f"{name}: {B} balls x {n_steps} steps, {scene}, rollout fwd + L1 loss + bwd"


def config_block(name, B, S, n_steps, per_env, args):
    """`config` of the JSON line: the workload only, identical for both arms of one command line."""
    return {"workload": workload_label(name, B, S, n_steps, per_env) +
            (" [reference regime: v0 = U(-1,2)^2, no impulse stand-in]" if args.reference_regime else ""),
            "balls_per_gpu": int(B), "sim_steps": int(n_steps), "segments": int(S), "per_env_maps": bool(per_env),
            "l2": "256 MiB buffer rewritten between timed iterations (flush), outside the event pairs"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's C restatement on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(w, steps, warmup, min_seconds=0.0):
    from oracle import ball_sim_np as onp
    from oracle.cref import CRef

    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        cr, flavour = CRef("native"), "native(-O3 -march=native)"
    except Exception:  # no compiler on the box: the prebuilt portable build travels with the repo
        cr, flavour = CRef("portable"), "portable(-O2)"
    c = onp.make_constants(w["constants"])
    n = w["n_steps"]
    dt = np.float32(w["sim_time"] / n)
    B = w["x0"].shape[0]

    def one():
        return cr.value_and_grad(n, dt, c, w["attractor"], w["target"], w["segments"], w["x0"], w["v0"])

    for _ in range(warmup):
        one()
    times = []
    t_all = time.perf_counter()
    k = 0
    while k < steps or (time.perf_counter() - t_all) < min_seconds:
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        k += 1
    ms = 1e3 * float(np.mean(times))
    return {"ms_per_step": ms, "value": B * n / (ms * 1e-3), "cores": cores, "steps": k,
            "flavour": flavour, "balls": B}


def bounded_sample(w, max_ball_steps_x_segs=4096 * 500 * 201):
    """Cap the CPU work of one step at about one C2 (~0.4 s on 8 cores): first balls only."""
    B, n, S = w["x0"].shape[0], w["n_steps"], w["segments"].shape[-3]
    cap = max(64, int(max_ball_steps_x_segs // (n * S)))
    if B <= cap:
        return w, f"the full workload ({B} balls x {n} steps)"
    s = dict(w)
    s["x0"], s["v0"] = w["x0"][:cap], w["v0"][:cap]
    if w["per_env"]:
        s["segments"] = w["segments"][:cap]
    return s, f"first {cap} of {B} balls x all {n} steps (per-ball work is independent)"


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML, polls during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.stop_flag, self.max_mhz, self.err = index, [], 0, False, None, None
        self.ready, self.window, self.handle, self.nv = threading.Event(), None, None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.handle, self.nv = h, nv
            while not self.stop_flag:
                self.sample()
                self.ready.set()
                time.sleep(0.002)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
            self.ready.set()

    def sample(self):
        """One (time, SM MHz, throttle reasons) sample; also called from the main thread while the timed steps are
        still running on the GPU, so that a timed region of a few milliseconds is never left without one."""
        nv, h = self.nv, self.handle
        if nv is None:
            return
        mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(
            nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        self.samples.append((time.perf_counter(), mhz, mask))

    def summary(self):
        """Median SM clock and the union of throttle reasons over the samples taken inside the timed region
        (`window` = its perf_counter bounds; samples before it are idle clocks and do not count)."""
        inside = [x for x in self.samples if self.window is None or self.window[0] <= x[0] <= self.window[1]]
        mask = 0
        for x in inside:
            mask |= x[2]
        reasons = [n for b, n in self.REASONS.items() if mask & b and n != "gpu_idle"]
        return {"sm_mhz": float(np.median([x[1] for x in inside])) if inside else None,
                "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(inside),
                **({"error": self.err} if self.err else {})}


def launches_per_step(lib, dref, separate):
    """Kernels of this repository per timed step: doper_value_and_grad launches the forward (+ the hand-over drain
    kernel of the baked plans), then either nothing more (fused), or loss + Jacobian pass as ONE kernel and the fold
    (default backward), or loss and the sequential backward; --separate-launches: forward, loss, backward kernels."""
    fwd = 1 + (1 if lib.doper_rollout_plan_name(dref, 3).decode() else 0)
    if separate:
        return fwd + 1 + len(lib.doper_rollout_plan_name(dref, 1).decode().split("+"))
    name = lib.doper_rollout_plan_name(dref, 2).decode()
    if "fused" in name:
        return 1
    if "loss_jacobian_kernel" in name:
        return fwd + len(name.split("+"))
    return fwd + 1 + len(name.split("+"))


def physical_gpu_index(local_index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


# ------------------------------------------------------------------------------------------------
# parity of the timed plan, measured on the timed workload (untimed, product code only)
# ------------------------------------------------------------------------------------------------
def parity_block(ball_sim, lib, check, ptr, st, desc, sws, bufs, w, scene, opts, gv_default, B, n_steps):
    """(a) the benchmarked backward (binary64 chunk-parallel) against the sequential reference-order fp32
    backward (flags bit 1: bit-identical to the oracle's adjoint, tests/test_parity_gpu.py) on the SAME
    forward workspace, over ALL balls; (b) the benchmarked forward against the exact full sweep (every
    (ball, segment) pair through the reference arithmetic, exact index log): hit flags at every step,
    segment index at every collision, final states bit for bit."""
    import copy
    import ctypes

    import torch

    d2 = copy.copy(desc)
    d2.flags = int(desc.flags) | 2
    gx2 = torch.empty(2 * B, dtype=torch.float32, device=gv_default.device)
    gv2 = torch.empty(2 * B, dtype=torch.float32, device=gv_default.device)
    # the seeds sign(x_T - target) of the backward (the fused launch computes them in place and leaves none behind)
    xT_t = bufs.out[4:4 + 2 * B]
    seeds = torch.empty(2 * B, dtype=torch.float32, device=gv_default.device)
    tmp = torch.empty(1, dtype=torch.float32, device=gv_default.device)
    check(lib.doper_l1_loss(st, B, ptr(xT_t), ball_sim._target2(w["target"]), 0, ptr(tmp), ptr(seeds)), "loss (seeds)")
    check(lib.doper_rollout_bwd(st, ctypes.byref(d2), ptr(sws), ptr(bufs.ws), 0, ptr(seeds), 0, ptr(gx2), ptr(gv2)),
          "bwd (sequential)")
    torch.cuda.synchronize()
    a = gv_default.double().view(B, 2).cpu().numpy()
    b = gv2.double().view(B, 2).cpu().numpy()
    fin = np.isfinite(a).all(1) & np.isfinite(b).all(1)
    nrm = np.maximum(np.linalg.norm(b[fin], axis=1), 1e-30)
    e = np.abs(a[fin] - b[fin]).max(axis=1) / nrm
    back = {"compared_to": "sequential backward (fp32, the reference's operation order; bit-exact vs the oracle adjoint in tests)",
            "balls": int(B), "balls_finite_in_both": int(fin.sum()),
            "balls_nonfinite_in_exactly_one": int((np.isfinite(a).all(1) != np.isfinite(b).all(1)).sum()),
            "rel_err_vs_ball_gradient_norm": {"median": float(np.median(e)), "q99": float(np.quantile(e, 0.99)),
                                              "q999": float(np.quantile(e, 0.999)), "max": float(e.max())},
            "fraction_within_1e-4": float((e <= 1e-4).mean()),
            "note": "the default backward is the float64 adjoint of the fp32 trajectory rounded once to fp32 "
                    "(within 1e-6 of oracle_grad_truth64 in tests); this gap is the fp32 reference-order adjoint's own rounding error"}
    cap = 8192 if B * n_steps > 64_000_000 else B
    sl = slice(0, cap)
    seg = w["segments"][sl] if w["per_env"] else w["segments"]
    args_h = (w["sim_time"], n_steps, seg, w["x0"][sl], w["v0"][sl], w["attractor"], w["constants"])
    h_def = ball_sim.rollout_history(*args_h, opts)
    h_ex = ball_sim.rollout_history(*args_h, ball_sim.RolloutOptions(force_exact_sweep=True, broad_phase=False,
                                                                    exact_index_log=True))
    hit = h_ex["hit"]
    fwd = {"compared_to": "exact full sweep: every (ball, segment) pair through the reference arithmetic, index logged at every step",
           "balls_compared": int(cap), "ball_steps_compared": int(cap) * int(n_steps), "collisions": int(hit.sum()),
           "hit_flag_mismatches": int((h_def["hit"] != hit).sum()),
           "collision_index_mismatches": int((h_def["idx"][hit] != h_ex["idx"][hit]).sum()),
           "final_state_bit_mismatches": int((h_def["x_T"].view(np.uint32) != h_ex["x_T"].view(np.uint32)).sum() +
                                             (h_def["v_T"].view(np.uint32) != h_ex["v_T"].view(np.uint32)).sum())}
    if cap != B:
        fwd["note"] = f"first {cap} balls of the batch (per-ball results do not depend on the batch: tests)"
    return {"backward": back, "forward": fwd}


# ------------------------------------------------------------------------------------------------
# the sharded BASELINE workloads (configs[2] and configs[4]) next to the headline line
# ------------------------------------------------------------------------------------------------
def measure_sharded(name, world, rank, device, exchange, steps, strong):
    """Device-timed fwd + loss + bwd (+ the per-step exchange when world > 1) of one rank's contiguous slice of a
    GLOBALLY seeded batch (results independent of the GPU count): c5 = 1,048,576 balls / 8 = 131,072 per GPU
    (weak: every GPU runs a full 131,072-ball share), c3 = 65,536 environments with their own maps (strong: the
    fixed total is split over the ranks).  Returns the per-rank timing; the caller takes the max over ranks."""
    import ctypes

    import torch

    from doper_b200 import synthetic as syn
    from doper_b200._abi import check, lib
    from doper_b200._device import ptr
    from doper_b200.distributed import ball_shard
    from doper_b200.sim import ball_sim
    from doper_b200.sim.jax_geometry import JaxScene

    full = syn.WORKLOADS[name][0]
    if name == "c5":
        per_gpu = full // 8
        total = per_gpu * world
    else:
        per_gpu = full // world if strong else full
        total = per_gpu * world
    lo, hi = ball_shard(total, rank, world)
    w = syn.make_workload(name, n_balls=total, lo=lo, hi=hi)
    B, n_steps = hi - lo, w["n_steps"]
    scene = JaxScene(w["segments"])
    opts = ball_sim.RolloutOptions()
    desc = ball_sim._make_desc(B, scene, w["sim_time"], n_steps, w["attractor"], w["constants"], opts)
    _, sws = scene.prepared(device, float(desc.consts[0]))
    x0, v0 = torch.tensor(w["x0"], device=device), torch.tensor(w["v0"], device=device)
    bufs = ball_sim._Buffers.get(device, desc)
    o = bufs.out
    xT, vT = o[4:4 + 2 * B], o[4 + 2 * B:4 + 4 * B]
    gv, gx, lpb = o[4 + 4 * B:4 + 6 * B], o[4 + 6 * B:4 + 8 * B], o[4 + 8 * B:4 + 9 * B]
    loss_sum = torch.zeros(1, dtype=torch.float32, device=device)
    tgt = ball_sim._target2(w["target"])
    dref = ctypes.byref(desc)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def step():
        st = torch.cuda.current_stream().cuda_stream
        check(lib.doper_rollout_fwd(st, dref, ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT), 0, 0, 0, ptr(bufs.ws)), "fwd")
        check(lib.doper_l1_loss(st, B, ptr(xT), tgt, ptr(lpb), ptr(loss_sum), ptr(bufs.scratch)), "loss")
        check(lib.doper_rollout_bwd(st, dref, ptr(sws), ptr(bufs.ws), ptr(bufs.bwd_scratch), ptr(bufs.scratch), 0, ptr(gx), ptr(gv)), "bwd")
        if world > 1:
            exchange()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    for a, b in ev:
        flush.fill_(3)
        a.record(); step(); b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    return {"ms": ms, "B": B, "n_steps": n_steps, "S": int(w["segments"].shape[-3]), "per_env": bool(w["per_env"]),
            "fwd_kernel": lib.doper_rollout_plan_name(dref, 0).decode(), "bwd_kernel": lib.doper_rollout_plan_name(dref, 1).decode(),
            "finite": bool(torch.isfinite(gv).all().item())}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def ours(args):
    import ctypes

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 when the first
        # communicator is created, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from doper_b200 import synthetic as syn
    from doper_b200._abi import check, lib
    from doper_b200._device import ptr
    from doper_b200.sim import ball_sim
    from doper_b200.sim.jax_geometry import JaxScene

    # ---- inputs: global arrays generated with the global seed, sliced by rank (weak scaling) ----
    name = args.workload
    per_gpu = args.balls or syn.WORKLOADS[name][0]
    if name == "c5" and not args.balls:
        per_gpu = syn.WORKLOADS[name][0] // 8
    if name == "c3" and not args.balls:
        per_gpu = syn.WORKLOADS[name][0] // max(world, 1) if args.strong else syn.WORKLOADS[name][0]
    total = per_gpu * world
    from doper_b200.distributed import ball_shard
    # Shards.  The named single-GPU workloads (c1, c2, c4) are defined by their seeds for ONE GPU; under
    # weak scaling every rank runs that same seeded batch (identical work per GPU, which is what the
    # scaling figure is about: launch + all-reduce overheads).  --distinct-shards slices a globally
    # seeded batch of world x per_gpu balls instead: the synchronous step then waits for the rank that
    # drew the ball with the longest collision chain (DESIGN.md "stragglers"), a property of the inputs.
    # c3 / c5 are DEFINED as global batches and are always sliced.
    sigma = 0.0 if args.reference_regime else 10.0
    replicated = world > 1 and name in ("c1", "c2", "c4") and not args.distinct_shards
    if replicated:
        lo, hi = 0, per_gpu
        w = syn.make_workload(name, n_balls=per_gpu, impulse_sigma=sigma)
    else:
        lo, hi = ball_shard(total, rank, world)  # contiguous slice of the globally seeded batch
        w = syn.make_workload(name, n_balls=total, lo=lo, hi=hi, impulse_sigma=sigma)
    B, n_steps = per_gpu, (args.sim_steps or w["n_steps"])
    w["n_steps"], w["sim_time"] = n_steps, n_steps / 1000.0
    S = int(w["segments"].shape[-3])
    opts = ball_sim.RolloutOptions(ckpt_every=args.ckpt_every, lanes_per_ball=args.lanes,
                                   force_exact_sweep=args.exact, sequential_backward=args.seq_bwd,
                                   smem_forward=args.no_reg,
                                   broad_phase=(False if args.full_sweep else None))
    scene = JaxScene(w["segments"])
    desc = ball_sim._make_desc(B, scene, w["sim_time"], n_steps, w["attractor"], w["constants"], opts)
    _, sws = scene.prepared(device, float(desc.consts[0]))
    x0 = torch.tensor(w["x0"], device=device)
    v0 = torch.tensor(w["v0"], device=device)
    bufs = ball_sim._Buffers.get(device, desc)
    o = bufs.out
    # the one exchange step of a data-parallel trainer: controller grads (6,602 fp32,
    # reference controllers.py:12-18) + the scalar loss, summed over ranks.  The loss kernel writes its sum
    # straight into slot 6,602 of the exchange buffer (no copy launch between backward and exchange).
    ar_buf = torch.zeros(6602 + 2, dtype=torch.float32, device=device)
    loss_sum, xT, vT = ar_buf[6602:6603], o[4:4 + 2 * B], o[4 + 2 * B:4 + 4 * B]
    gv, gx, lpb = o[4 + 4 * B:4 + 6 * B], o[4 + 6 * B:4 + 8 * B], o[4 + 8 * B:4 + 9 * B]
    tgt = ball_sim._target2(w["target"])
    cur = lambda: torch.cuda.current_stream().cuda_stream  # launches follow torch's current stream (graph capture)
    st = cur()
    dref = ctypes.byref(desc)
    ar_out = torch.zeros(6602 + 2, dtype=torch.float32, device=device)
    peer = None
    if world > 1 and not args.nccl:
        from doper_b200.distributed import PeerAllReducer
        try:
            peer = PeerAllReducer(6602 + 2, device)  # one-shot all-reduce over NVLink peer memory (csrc/peer.cuh)
        except RuntimeError as e:  # no CUDA IPC between these GPUs (raised on every rank together): NCCL
            if rank == 0:
                print(f"bench: peer-memory all-reduce unavailable, using NCCL ({e})", file=sys.stderr)
            peer = None

    def exchange():
        # the optimiser step needs the reduced gradient: the exchange is on the step's critical path
        if peer is not None:
            peer.all_reduce(ar_buf, out=ar_out)
        else:
            ar_out.copy_(ar_buf, non_blocking=True)
            dist.all_reduce(ar_out)

    def fwd():
        check(lib.doper_rollout_fwd(cur(), dref, ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT), 0, 0, 0,
                                    ptr(bufs.ws)), "fwd")

    def loss_bwd():
        check(lib.doper_l1_loss(cur(), B, ptr(xT), tgt, ptr(lpb), ptr(loss_sum), ptr(bufs.scratch)), "loss")
        check(lib.doper_rollout_bwd(cur(), dref, ptr(sws), ptr(bufs.ws), ptr(bufs.bwd_scratch), ptr(bufs.scratch), 0, ptr(gx), ptr(gv)),
              "bwd")

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def value_and_grad():
        # the C-ABI entry of the path (= the reference's vmapped_grad_and_value): ONE launch when the pipelined plan
        # applies (forward + loss + backward fused per block), else forward, loss and backward launches
        check(lib.doper_value_and_grad(cur(), dref, ptr(sws), ptr(x0), ptr(v0), tgt, ptr(bufs.ws), ptr(bufs.bwd_scratch),
                                       ptr(bufs.scratch), ptr(loss_sum), ptr(lpb), ptr(xT), ptr(vT), ptr(gv), ptr(gx)), "value_and_grad")

    def one_step():
        if args.separate_launches:
            fwd(); loss_bwd()
        else:
            value_and_grad()
        if world > 1:
            exchange()

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    # The timed loop replays ONE CUDA graph per step (memset + forward + loss + backward (2 kernels) + exchange):
    # the same kernels on the same buffers, without the host's per-launch gaps.  --no-graph launches eagerly.
    step_graph = None
    if not args.no_graph:
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            one_step()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize()
        step_graph, fwd_graph = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(step_graph, stream=side):
            one_step()
        with torch.cuda.graph(fwd_graph, stream=side):
            fwd()
        torch.cuda.synchronize()
        for _ in range(3):
            step_graph.replay()
        barrier()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    sampler.ready.wait(5.0)  # NVML is initialised and sampling before the timed region starts
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    pending = None
    barrier()
    t_wall = time.perf_counter()
    for k in range(K):
        flush.fill_(k & 0xFF)  # L2 flush between timed iterations (untimed: outside the event pairs)
        ev[k][0].record()
        if step_graph is not None:
            step_graph.replay()
        else:
            one_step()
        ev[k][2].record()
    try:
        sampler.sample()  # the steps just enqueued are still running: a sample under load whatever the region's length
    except Exception:  # pragma: no cover
        pass
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    sampler.window = (t_wall, time.perf_counter())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    step_ms = np.array([e[0].elapsed_time(e[2]) for e in ev])
    if step_graph is not None:
        # the forward kernel alone (the roofline's dominant kernel), same buffers, after an L2 flush each
        fev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(min(K, 20))]
        for a, b in fev:
            flush.fill_(1)
            a.record(); fwd_graph.replay(); b.record()
        torch.cuda.synchronize()
        fwd_ms = np.array([a.elapsed_time(b) for a, b in fev])
    else:
        fev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(min(K, 20))]
        for a, b in fev:
            flush.fill_(1)
            a.record(); fwd(); b.record()
        torch.cuda.synchronize()
        fwd_ms = np.array([a.elapsed_time(b) for a, b in fev])
    tot = torch.tensor([step_ms.sum()], dtype=torch.float64, device=device)
    tot_min = tot.clone()
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_min, op=dist.ReduceOp.MIN)
    ms_per_step = float(tot.item()) / K
    ms_per_step_fastest_rank = float(tot_min.item()) / K
    value = world * B * n_steps / (ms_per_step * 1e-3)

    # ---- e2e: public API with host inputs (H2D + launches + D2H inside the timed region) ----
    xh, vh = w["x0"], w["v0"]
    api_args = (w["sim_time"], n_steps, scene, xh, vh, w["target"], w["attractor"], w["constants"], opts)
    for _ in range(3):
        ball_sim.vmapped_grad_and_value(*api_args)
    barrier()
    e2e_steps = max(5, min(K, 50))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        (l_e2e, (xT_e2e, _)), gv_e2e = ball_sim.vmapped_grad_and_value(*api_args)
        if world > 1:
            exchange()  # the same per-step exchange as in the device-timed loop
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([1e3 * (time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * n_steps / (float(e2e_ms.item()) * 1e-3)

    # ---- parity of the timed plan on the timed workload (untimed) -----------------------------
    parity = None
    if rank == 0 and not args.no_parity:
        # eager, the timed entry point, WITHOUT the exchange (a collective: rank 0 alone must not enter it)
        if args.separate_launches:
            fwd(); loss_bwd()
        else:
            value_and_grad()
        torch.cuda.synchronize()
        parity = parity_block(ball_sim, lib, check, ptr, cur(), desc, sws, bufs, w, scene, opts, gv, B, n_steps)
    # ---- N > 1: the peer-memory exchange against NCCL on the same vectors (every run) ---------
    exch_check = None
    if world > 1 and peer is not None:
        g = torch.Generator(device=device); g.manual_seed(1234 + rank)
        worst, bits_equal = 0.0, True
        for _ in range(20):
            v = torch.randn(6602 + 2, device=device, generator=g) * 100.0
            ref_v = v.clone(); dist.all_reduce(ref_v)
            mag = v.abs(); dist.all_reduce(mag)  # sum over ranks of |v_i|: the scale of the rounding error of any summation order
            got_v = peer.all_reduce(v).clone()
            worst = max(worst, float(((got_v - ref_v).abs() / mag.clamp_min(1e-30)).max()))
            gathered = [torch.empty_like(got_v) for _ in range(world)]
            dist.all_gather(gathered, got_v)
            bits_equal = bits_equal and all(torch.equal(gathered[0].view(torch.int32), t.view(torch.int32)) for t in gathered)
        peer.check()
        exch_check = {"vectors": 20, "max_diff_vs_nccl_relative_to_sum_of_magnitudes": worst, "bit_identical_across_ranks": bool(bits_equal),
                      "what": "doper_allreduce_oneshot vs torch.distributed.all_reduce (NCCL) on the same random vectors"}
    # ---- the sharded BASELINE workloads on the same ranks (configs[4] weak, configs[2] strong) -------
    sharded = None
    if args.sharded or (world > 1 and not args.no_sharded and name == "c2" and not args.balls and not args.sim_steps):
        sharded = {}
        for wl, strong in (("c5", False), ("c3", True)):
            barrier()
            r = measure_sharded(wl, world, rank, device, exchange, steps=5, strong=strong)
            t = torch.tensor([r["ms"]], dtype=torch.float64, device=device)
            ok = torch.tensor([1.0 if r["finite"] else 0.0], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            tot_b = r["B"] * world
            sharded[wl] = {
                "workload": workload_label(wl, r["B"], r["S"], r["n_steps"], r["per_env"]) + f" per GPU, {tot_b} in total",
                "scaling": "strong (65,536 environments split over the ranks)" if strong else "weak (131,072 balls per GPU = 1,048,576 / 8)",
                "n_gpus": world, "steps": 5, "warmup": 3, "ms_per_step": round(float(t.item()), 4),
                "value": round(tot_b * r["n_steps"] / (float(t.item()) * 1e-3), 1), "unit": UNIT,
                "forward_kernel": r["fwd_kernel"], "backward_kernel": r["bwd_kernel"], "gradients_finite": bool(ok.item() > 0.5),
                "timing": "CUDA events around fwd + loss + bwd" + (" + exchange" if world > 1 else "") + ", L2 flushed between steps, max over ranks",
            }
    barrier()
    # ---- roofline of the dominant kernel -----------------------------------------------------
    roof = None
    cpu = None
    if rank == 0:
        probe_out = torch.empty(148 * 8 * 256, dtype=torch.float32, device=device)
        rates = {}
        for use_fma in (0, 1):
            iters = 8192
            for _ in range(2):
                check(lib.doper_fp32_probe(st, 148 * 8, 256, iters, use_fma, ptr(probe_out)), "probe")
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            check(lib.doper_fp32_probe(st, 148 * 8, 256, iters, use_fma, ptr(probe_out)), "probe")
            b.record()
            torch.cuda.synchronize()
            rates[use_fma] = 148 * 8 * 256 * iters * 16 / (a.elapsed_time(b) * 1e-3)
        fwd_flops = B * n_steps * (24.0 * S + 34.0)  # algorithmic flops of the fwd kernel
        achieved = fwd_flops / (float(fwd_ms.mean()) * 1e-3) / 1e12
        peak = rates[0] / 1e12
        K_ck = int(lib.doper_rollout_ckpt_every(dref))
        ws_bytes = (-(-n_steps // K_ck) * 16 + n_steps * 4) * B  # ckpt + hit log written
        traffic = None
        tp = ROOT / "profiles" / "roofline_traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("traffic", {}).get(
                    name + ("_full_sweep" if args.full_sweep else "")) if (not args.balls and not args.sim_steps) else None
            except Exception:
                traffic = None
        peaks = {}
        mp = ROOT / "MEASURED_PEAKS.json"
        if mp.exists():
            peaks = json.loads(mp.read_text())
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        fwd_kernel = lib.doper_rollout_plan_name(dref, 0).decode()
        bwd_kernel = lib.doper_rollout_plan_name(dref, 1 if args.separate_launches else 2).decode()
        prof = {}
        if tp.exists():
            try:
                prof = json.loads(tp.read_text()).get("kernels", {}).get(fwd_kernel, {})
            except Exception:
                prof = {}
        pruned = "cull" in fwd_kernel or "tpc" in fwd_kernel or "baked" in fwd_kernel
        # executed work (what the silicon did, from the committed ncu capture of this kernel on this workload)
        executed = None
        wi = prof.get("warp_instructions_per_launch") if (not args.balls and not args.sim_steps and prof.get("workload") == name) else None
        sm_mhz = (sampler.summary().get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        chain_cycles = 88  # dependent chain of one free-flight step: 22 FP32 operations x 4 cycles (physics.cuh:free_step)
        if wi:
            kernel_cycles = float(fwd_ms.mean()) * 1e-3 * sm_mhz * 1e6
            executed = {
                "warp_instructions_per_ball_step": round(wi / (B * n_steps), 2),
                "active_lanes_per_warp_instruction": prof.get("active_lanes_per_instruction"),
                "fp32_lane_ops_per_ball_step": prof.get("fp32_lane_ops_per_ball_step"),
                "issue_slots_used": round(wi / (kernel_cycles * 148 * 4), 4),
                "lane_issue_slots_used": round(wi * (prof.get("active_lanes_per_instruction") or 32) / (kernel_cycles * 148 * 4 * 32), 4),
                "source": prof.get("report"),
            }
        roof = {
            "kernel": fwd_kernel, "backward_kernel": bwd_kernel, "bound": "fp32", "achieved": round(achieved, 3),
            "peak": round(peak, 3), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
            "achieved_definition": "ALGORITHMIC flops of the reference's sweep, (24 S + 34) per ball-step (SURVEY 8d), "
                                   "divided by the forward kernel's event-timed duration"
                                   + ("; this kernel's broad phase proves most (ball, segment) pairs irrelevant and "
                                      "never evaluates them (results bit-identical), so frac measures speed against "
                                      "the full-sweep roofline and can exceed 1; --full-sweep times the kernel that "
                                      "evaluates every pair" if pruned else ""),
            "fma_pipe_pct_of_peak_ncu": prof.get("fma_pipe_pct"), "issue_active_pct_ncu": prof.get("issue_active_pct"),
            "peak_source": "non-FMA FADD/FMUL lane-op rate measured in this run (doper_fp32_probe); "
                           "a bit-exact sweep cannot contract the reference's mul+add pairs",
            "frac_of_fma_peak": round(fwd_flops / (float(fwd_ms.mean()) * 1e-3) / (2 * rates[1]), 4),
            "fma_peak_tflops": round(2 * rates[1] / 1e12, 3),
            "kernel_ms": round(float(fwd_ms.mean()), 4),
            "kernel_share_of_step": round(float(fwd_ms.mean() / step_ms.mean()), 4),
            "algorithmic_flops_per_launch": fwd_flops,
            "traffic": traffic,
            "executed": executed,
            "serial_floor_us": round(n_steps * chain_cycles / sm_mhz, 2),
            "serial_floor_definition": "n_steps x the dependent FP32 chain of one free-flight step (22 operations x 4 cycles): "
                                       "a ball's steps cannot overlap, so no forward kernel can be faster than this",
            "hbm": {"achieved": round(ws_bytes / (float(fwd_ms.mean()) * 1e-3) / 1e9, 2), "peak": hbm_peak,
                    "unit": "GB/s", "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "algorithmic_bytes_per_launch": ws_bytes},
        }
        if world == 1 and not args.no_cpu:
            sample, what = bounded_sample(w)
            r = cpu_reference_run(sample, steps=3, warmup=1, min_seconds=10.0)
            cpu = {"value": round(r["value"], 1), "unit": UNIT, "cores": r["cores"], "kind": "port",
                   "sample": f"{what}, {r['steps']} repetitions, oracle/ball_sim_ref.c {r['flavour']} + OpenMP "
                             f"(restated reference; the reference's JAX is not installable here)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(name, B, S, n_steps, bool(w["per_env"]), args),
            "plan": {"forward_kernel": lib.doper_rollout_plan_name(dref, 0).decode(),
                     "backward_kernel": lib.doper_rollout_plan_name(dref, 1 if args.separate_launches else 2).decode(),
                     "ckpt_every": int(lib.doper_rollout_ckpt_every(dref)),
                     "sweep": "exact" if args.exact else ("filtered+exact, every pair" if args.full_sweep else
                                                        "broad phase (baked per-cell lists / per-ball lists) + exact"),
                     "launch": "eager" if args.no_graph else "one CUDA graph per step",
                     "shards": ("n/a" if world == 1 else ("every rank runs the workload's own seeded batch" if replicated
                                                          else f"a globally seeded batch of {total} balls sliced by rank")),
                     "exchange": ("none" if world == 1 else
                                  ("one-shot NVLink peer-memory all-reduce kernel (csrc/peer.cuh)" if peer is not None else "NCCL all-reduce")
                                  + " of 6,604 fp32 (controller gradients + loss) per step, inside the timed region"),
                     "ms_per_step_fastest_rank": round(ms_per_step_fastest_rank, 5)},
            "clocks": sampler.summary(),
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": 4 * B * 4,
                    "d2h_bytes_per_step": (4 + 6 * B) * 4, "ms_per_step": round(float(e2e_ms.item()), 5),
                    "api": "doper_b200.sim.ball_sim.vmapped_grad_and_value(host numpy arrays)"},
            # kernels of this repository per step: forward, loss, backward (1 or 2), exchange (N > 1, peer kernel)
            "gpu_launches": (launches_per_step(lib, dref, args.separate_launches) + (1 if peer is not None else 0)) * K,
            "wall_ms_timed_region": round(wall_ms, 3),
            "roofline": roof,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if parity:
            line["parity"] = parity
        if exch_check:
            line["exchange_self_check"] = exch_check
        if sharded:
            line["sharded_workloads"] = sharded
        print(json.dumps(line), flush=True)
    if world > 1:
        if peer is not None:
            peer.close()
        dist.destroy_process_group()


def reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from doper_b200 import synthetic as syn

    name = args.workload
    per_gpu = args.balls or syn.WORKLOADS[name][0]
    w = syn.make_workload(name, n_balls=per_gpu, impulse_sigma=0.0 if args.reference_regime else 10.0)
    if args.sim_steps:
        w["n_steps"], w["sim_time"] = args.sim_steps, args.sim_steps / 1000.0
    sample, what = bounded_sample(w)
    S = int(w["segments"].shape[-3])
    r = cpu_reference_run(sample, steps=args.steps, warmup=args.warmup)  # exactly --warmup untimed, --steps timed
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 1), "unit": UNIT,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
        "ms_per_step": round(r["ms_per_step"], 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same `config` object as the product arm prints for this command line (kernel-plan details of
        # the product arm live under its own `plan` key)
        "config": config_block(name, per_gpu, S, w["n_steps"], bool(w["per_env"]), args),
        "cpu_baseline": {"value": round(r["value"], 1), "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": f"each step = {what}; oracle/ball_sim_ref.c {r['flavour']} + OpenMP on all host "
                                   f"cores (restated reference: the reference is JAX/XLA:CPU, not installable here)"},
        "e2e": {"value": round(r["value"], 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["c1", "c2", "c3", "c4", "c5"], default="c2")
    ap.add_argument("--balls", type=int, default=0, help="balls per GPU (default: the workload's)")
    ap.add_argument("--sim-steps", type=int, default=0, help="rollout length (default: the workload's)")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--ckpt-every", type=int, default=0)
    ap.add_argument("--exact", action="store_true", help="force the exact (unfiltered) sweep")
    ap.add_argument("--seq-bwd", action="store_true", help="force the sequential (oracle-order) backward")
    ap.add_argument("--no-reg", action="store_true", help="do not use the register-resident forward kernel")
    ap.add_argument("--full-sweep", action="store_true",
                    help="forbid the broad-phase kernels: every (ball, segment) pair is evaluated every step")
    ap.add_argument("--strong", action="store_true")
    ap.add_argument("--reference-regime", action="store_true",
                    help="v0 = U(-1,2)^2 only (SURVEY 8d: the trainer's draw without the impulse stand-in): slow balls, rare collisions")
    ap.add_argument("--nccl", action="store_true", help="N > 1: exchange through NCCL instead of the peer-memory kernel")
    ap.add_argument("--distinct-shards", action="store_true",
                    help="N > 1: slice one globally seeded batch instead of running the workload's batch on every rank")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch the timed steps eagerly instead of replaying one CUDA graph")
    ap.add_argument("--no-parity", action="store_true", help="skip the (untimed) parity block")
    ap.add_argument("--separate-launches", action="store_true",
                    help="time doper_rollout_fwd + doper_l1_loss + doper_rollout_bwd instead of doper_value_and_grad")
    ap.add_argument("--sharded", action="store_true",
                    help="also time one rank's slice of the sharded BASELINE workloads (c5 weak, c3 strong); default when N > 1")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the sharded workloads block")
    args = ap.parse_args()
    if args.impl == "reference":
        reference(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
