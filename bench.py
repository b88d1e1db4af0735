#!/usr/bin/env python
"""bench.py -- DAE train-step throughput (playlists/s) on B200, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "cfg2"): untied DAE (--dae), B=256 playlists per GPU, 250 000 tracks
+ 40 000 artists, latent 256, bf16 tensor-core operands / fp32 accumulate+master, synthetic MPD-shaped
batches (tools/synth_mpd.py).  One step = one `sess.run([optimizer, cost])` of the reference:
COO->CSR, encode, decode + weighted BCE + backward, dense TF1 Adam on every variable.

Printed JSON (one line, rank 0):
  value  : playlists/s, device-timed (CUDA events on the launching stream), inputs already staged in HBM
  e2e    : the same metric through the public API (models.DAEs.DAE.train_step) with HOST buffers: pinned
           H2D of the batch and D2H of the cost inside the timed region
  roofline: dominant kernel (decoder Adam, HBM-bound): algorithmic bytes / measured launch time vs the
           measured copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline: the TF1 graph restated op-for-op on torch-CPU (oracle/tf1_graph_cpu.py), bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (n_tracks, n_artists, hidden, batch, tied)
    "cfg2": (250000, 40000, 256, 256, False),
    "cfg1": (5000, 1000, 64, 128, True),
    # secondary lines (not the headline; measured through the host API, H2D inside the timed region):
    "cfg3": (250000, 40000, 256, 256, False),     # DAE + char-CNN title head train step (--title)
    "cfg5": (2000000, 0, 256, 4096, False),       # challenge inference: top-500 over a 2M-item decoder, batch 4096
}
KP, KP_IN = 0.8, 0.75          # [DAE] keep_prob / input_kp of the shipped configs (0to1_inorder/config.ini:18-19)
LR = 0.005


def make_batches(wl, n, seed, rank=0):
    from tools.synth_mpd import SynthMPD
    T, A, H, B, tied = WORKLOADS[wl]
    g = SynthMPD(T, A, n_clusters=64, seed=180610 + rank)
    rng = np.random.default_rng(seed + 1000 * rank)
    out = []
    for i in range(n):
        trk, art, y, titles, tv, av = g.coo_batch(B, rng)
        # hide-and-seek: tracks-only or artists-only input, tracks+artists target (main_train.py:202-213)
        x, xv = (trk, tv) if i % 2 == 0 else (art, av)
        out.append((np.ascontiguousarray(x), xv.astype(np.float32), np.ascontiguousarray(y),
                    np.ones(len(y), np.float32)))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference(wl, steps, warmup):
    """The reference's CPU path (TF1 graph restated on torch-CPU, all host threads), bounded sample."""
    import torch
    from oracle.tf1_graph_cpu import TF1GraphCPU
    T, A, H, B, tied = WORKLOADS[wl]
    N = T + A
    m = TF1GraphCPU(N, H, LR, tied=tied, seed=0)
    batches = make_batches(wl, 2, seed=7)
    for i in range(warmup):
        m.train_step(*batches[i % 2], B, KP, KP_IN)
    t0 = time.perf_counter()
    for i in range(steps):
        m.train_step(*batches[i % 2], B, KP, KP_IN)
    dt = time.perf_counter() - t0
    return {"value": B * steps / dt, "unit": "playlists/s", "cores": torch.get_num_threads(), "kind": "port",
            "host_cpus": os.cpu_count(), "ms_per_step": 1e3 * dt / steps,
            "sample": "%d steps (after %d warm-up) of the %s workload: dense fp32 TF1 graph of models/DAEs.py "
                      "restated on torch-CPU (TF1 not installable: py3.12, no network)" % (steps, warmup, wl)}


def aux_workload(args, wl, rank=0, world=1):
    """cfg3 (title-mode train step, one GPU) / cfg5 (challenge inference, item-sharded over `world` GPUs) through the
    public host API."""
    import torch
    from spotify_recsys_challenge_2018_b200.models.DAEs import DAE, DAE_title
    from spotify_recsys_challenge_2018_b200.models.title_get import get_model
    from tools.synth_mpd import SynthMPD
    T, A, H, B, tied = WORKLOADS[wl]
    N = T + A

    class Conf:
        pass
    conf = Conf()
    conf.save = "/tmp/bench_w"; conf.n_input = N; conf.n_tracks = T; conf.n_output = N; conf.hidden = H
    conf.lr = LR; conf.reg_lambda = 0.0; conf.initval = "NULL"; conf.DAEval = "NULL"; conf.seed = 0
    conf.device = int(os.environ.get("LOCAL_RANK", "0"))
    conf.charsize = 41; conf.strmaxlen = 25; conf.char_emb = 50; conf.char_model = "Char_CNN"
    conf.filter_num = 100; conf.filter_size = [3, 5, 7, 9]                    # */config.ini [TITLE]
    g = SynthMPD(T, max(A, 1), n_clusters=64, seed=180610)
    rng = np.random.default_rng(7)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if wl == "cfg3":
        conf.batch = B
        tm = get_model(conf)
        m = DAE_title(conf, tm).fit()
        tm.fit(m)
        batches = []
        for i in range(4):
            trk, art, y, titles, tv, av = g.coo_batch(B, rng)
            batches.append((np.ascontiguousarray(y), np.ones(len(y), np.float32), np.asarray(titles, np.int64)))
        run = lambda i: tm.train_step(m, batches[i % 4][0], batches[i % 4][1], batches[i % 4][2], KP, 0.7, 0.01)
        h2d = int(np.mean([2 * (y.nbytes + v.nbytes) + t.nbytes for y, v, t in batches]))
        d2h, units, metric = 4, B, "dae_title_train_playlists_per_sec"
        desc = "cfg3: DAE + char-CNN title head train step (--title), B=%d, %d tracks + %d artists, latent %d, 4x100 filters" % (B, T, A, H)
        launches = lambda: tm.launch_count() + m.launch_count()
    else:
        # challenge inference: ONE call ranks the whole 4096-playlist batch (fused decode + top-K, 16 batch tiles of 256 rows
        # inside the kernel grid); with --gpus N every rank ranks its slice of the item axis and the lists are merged
        conf.batch = B
        m = DAE(conf)
        m.trainable = False
        m.fit()
        trk, art, y, titles, tv, av = g.coo_batch(B, rng)
        trk = np.ascontiguousarray(trk); tv = tv.astype(np.float32)
        order = np.argsort(trk[:, 0], kind="stable")
        bounds = np.searchsorted(trk[order, 0], np.arange(B + 1))
        seeds = (bounds.astype(np.int32), trk[order, 1].astype(np.int32))     # CSR of the seed tracks (= the input tracks)
        rec = m
        if world > 1:
            from spotify_recsys_challenge_2018_b200.dp import ShardedRecommender
            rec = ShardedRecommender(m)
        run = lambda i: rec.recommend(trk, tv, seeds, k=500)
        if world > 1:
            # the merged per-shard lists must be the unsharded list (checked once, outside the timed region)
            got = rec.recommend(trk, tv, seeds, k=500, return_scores=True)
            want = m.recommend(trk, tv, seeds, k=500, return_scores=True)
            if not (np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])):
                bad = np.nonzero((got[0] != want[0]).any(1))[0]
                raise SystemExit("bench.py: item-sharded top-k differs from the unsharded list (rank %d, range %s): %d rows differ, "
                                 "idx mismatch %.4f, first bad row %d: got %s / %s want %s / %s"
                                 % (rank, rec.range, len(bad), (got[0] != want[0]).mean(), bad[0] if len(bad) else -1,
                                    got[0][bad[0]][:6] if len(bad) else "", got[1][bad[0]][:6] if len(bad) else "",
                                    want[0][bad[0]][:6] if len(bad) else "", want[1][bad[0]][:6] if len(bad) else ""))
        h2d = int(trk.nbytes + tv.nbytes + 4 * (B + 1) + 4 * len(trk))
        d2h, units, metric = B * 500 * 4, B, "dae_challenge_topk_playlists_per_sec"
        desc = ("cfg5: challenge inference, top-500 over a %d-item decoder, batch %d in one call (fused decode + top-K, "
                "16 batch tiles of 256 rows), latent %d, item axis sharded over %d GPU(s)" % (T, B, H, world))
        launches = m.launch_count
    for i in range(max(args.warmup, 3) if wl == "cfg3" else 2):
        run(i)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    steps = args.steps if wl == "cfg3" else max(1, min(args.steps, 10))
    l0 = launches()
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank != 0:
        m.close()
        return 0
    line = {"metric": metric, "value": units * steps / dt, "unit": "playlists/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "note": "secondary line: timed through the host API (H2D + D2H inside)"},
            "e2e": {"value": units * steps / dt, "unit": "playlists/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches() - l0)}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-flags", type=int, default=0, help="extra dae_model_set_debug bits (include/dae_b200.h)")
    ap.add_argument("--no-overlap", action="store_true", help="decoder update on the main stream (no overlap with the encoder tail)")
    ap.add_argument("--two-kernel", action="store_true", help="decoder dW and Adam as two kernels (gradient through HBM)")
    args = ap.parse_args()
    wl = args.workload
    T, A, H, B, tied = WORKLOADS[wl]
    N = T + A
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "%s: %s DAE train step, B=%d playlists/GPU, %d tracks + %d artists, latent %d"
                          % (wl, "tied" if tied else "untied", B, T, A, H),
              "global_batch": B * world, "parallelism": "dp%d" % world,
              "l2": ("working set per step (parameters + Adam state, %.1f GB) >> 126 MB L2; no explicit flush"
                     % ((4 if not tied else 2) * 3 * N * H * 4 / 1e9)) if N * H * 4 > 126e6 else
                    "working set fits L2 (the reference's CPU-runnable parity case, not a bench line)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, args.cpu_steps if wl == "cfg2" else args.steps))
        warm = max(1, min(args.warmup, 2))
        base = cpu_reference(wl, steps, warm)
        line = {"impl": "reference", "metric": "dae_train_playlists_per_sec", "value": base["value"],
                "unit": "playlists/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "playlists/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if wl == "cfg3":
        return aux_workload(args, wl) if rank == 0 else 0
    if wl == "cfg5":
        rc = aux_workload(args, wl, rank, world)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return rc
    from spotify_recsys_challenge_2018_b200.dp import DataParallelDAE
    from spotify_recsys_challenge_2018_b200.models.DAEs import DAE, DAE_tied

    class Conf:
        pass
    conf = Conf()
    conf.save = "/tmp/bench_w"; conf.batch = B; conf.n_input = N; conf.n_tracks = T; conf.hidden = H
    conf.lr = LR; conf.reg_lambda = 0.0; conf.initval = "NULL"; conf.seed = 0; conf.device = local_rank
    conf.world = world; conf.rank = rank
    stream = torch.cuda.Stream()
    conf.stream = stream.cuda_stream
    with torch.cuda.stream(stream):
        model = (DAE_tied if tied else DAE)(conf).fit()
        if args.two_kernel or args.no_overlap or args.debug_flags:
            model.set_debug((4 if args.two_kernel else 0) | (8 if args.no_overlap else 0) | args.debug_flags)
        trainer = DataParallelDAE(model) if world > 1 else None
        batches = make_batches(wl, 8, seed=7, rank=rank)
        model.stage_batch(0, *batches[0])
        model.stage_batch(1, *batches[1])

        def step(i):
            # inputs stay resident in HBM; the slot's COO->CSR / bitmask preparation is re-run every step
            # on the library's side stream, one step ahead of the main stream (dae_model_restage)
            if trainer is not None:
                trainer.train_step_staged(i & 1, KP, KP_IN)
            else:
                model.train_step_staged(i & 1, KP, KP_IN)
            model.restage(i & 1)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- device-timed throughput, inputs resident in HBM ------------------------------
        # clocks / throttle reasons are sampled under load: from the warm-up through the timed region and the
        # end-to-end loop (the timed region alone lasts ~50 ms, too short for nvidia-smi's sampling period)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for i in range(args.warmup):
            step(i)
        model.sync_cost()
        barrier()
        l0 = model.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            step(i)
        e1.record(stream)
        barrier()
        launches = model.launch_count() - l0
        ms = e0.elapsed_time(e1)
        cost = model.sync_cost()
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        value = B * world * args.steps / (ms / 1e3)

        # ---- end to end through the public API: host COO in, cost out ------------------------
        e2e = None
        if world == 1:
            for i in range(3):
                model.train_step_async(*batches[i % len(batches)], KP, KP_IN)
            model.flush()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(args.steps):
                model.train_step_async(*batches[i % len(batches)], KP, KP_IN)   # returns the previous step's cost
            model.flush()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            h2d = int(np.mean([x.nbytes + xv.nbytes + y.nbytes + yv.nbytes for x, xv, y, yv in batches]))
            e2e = {"value": B * args.steps / dt, "unit": "playlists/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 8, "ms_per_step": 1e3 * dt / args.steps,
                   "api": "models.DAEs.DAE.train_step_async (dae_model_train_step_async), the call main_runner/main_train.py "
                          "makes: host int64 COO + fp32 values in (pinned H2D every step), every step's cost read back "
                          "(D2H, one step late), final flush inside the timed region"}
        else:
            # DP: every rank feeds its own B rows of the global batch from host memory: H2D of the rank's batch,
            # the step (NVLink exchange inside the kernels) and D2H of the cost inside the timed region
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                model.train_step_async(*batches[i % len(batches)], KP, KP_IN)
            model.flush()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda")
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            h2d = int(np.mean([x.nbytes + xv.nbytes + y.nbytes + yv.nbytes for x, xv, y, yv in batches]))
            e2e = {"value": B * world * args.steps / float(dt.item()), "unit": "playlists/s",
                   "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 8 * world,
                   "ms_per_step": 1e3 * float(dt.item()) / args.steps,
                   "api": "models.DAEs.DAE.train_step_async on every rank (world = N model attached with dp.DataParallelDAE): "
                          "per-rank host COO in, per-step cost back"}

        clocks = sampler.stop() if rank == 0 else None
        # ---- per-phase device times (profiled steps; not part of `value`) -----------------------
        model.set_profiling(True)
        for i in range(min(args.steps, 20)):
            step(i)
        model.sync_cost()
        phases = {k: (ms_ / max(n, 1)) for k, (ms_, n) in model.phase_times().items() if n}
        model.set_profiling(False)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        # dominant kernel.  Default step: k_dw_adam_fused = dW_dec tile in tensor memory + dense TF1 Adam on the decoder
        # rows this GPU owns (N/world): algorithmic bytes per launch = dz read 2 B x K/256 + w,m,v read 12 B
        # + w,m,v write 12 B + bf16 operand copy 2 B = 28 B / parameter at K = 256 (the gradient never exists in HBM).
        # --two-kernel (debug bit 2): k_adam_rows_vec4 = 30 B / parameter (g read 4 + 24 + 2).
        n_own = N / world
        fused = "adam_dec" not in phases
        if fused:
            roof_kernel, roof_ms = "k_dw_adam_fused", phases.get("dw_dec")
            roof_bytes = (26.0 + 2.0 * world) * n_own * H
            roof_desc = "k_dw_adam_fused (dW_dec tile in TMEM + dense TF1 Adam on the decoder rows + bf16 operand refresh)"
        else:
            roof_kernel, roof_ms = "k_adam_rows_vec4", phases.get("adam_dec")
            roof_bytes = 30.0 * n_own * H
            roof_desc = "k_adam_rows_vec4 (decoder rows: dense TF1 Adam + bf16 operand refresh)"
        roofline = None
        # DRAM bytes of the same launch from the committed ncu --set full capture (profiles/traffic.json, written by
        # tools/summarize_profile.py): only quoted for the single-GPU shape it was captured on
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                caps = json.load(f)["kernels"].get(roof_kernel, [])
            if caps and world == 1 and wl == "cfg2":
                traffic = caps[0]["dram_bytes"]
        except Exception:
            pass
        if roof_ms:
            ach = roof_bytes / (roof_ms / 1e3) / 1e9
            roofline = {"kernel": roof_desc, "bound": "hbm",
                        "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": roof_bytes,
                        "launch_ms": roof_ms}
        # whole step (per GPU): decoder dW+Adam 26 (fused; 34 as two kernels) + encoder Adam 24 (untied) on the owned rows,
        # + W operand read by decode and dh (2 + 2) + dz write (2) and re-read by dh and dW (2 + 2) over all N rows
        step_bytes = ((26.0 if fused else 34.0) + (24.0 if not tied else 0.0)) * n_own * H + 10.0 * N * H
        line = {"metric": "dae_train_playlists_per_sec", "value": value, "unit": "playlists/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "phase_ms": phases, "last_cost": cost,
                "step_hbm_gbs_algorithmic": step_bytes / (ms / args.steps / 1e3) / 1e9}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(wl, args.cpu_steps if wl == "cfg2" else 50, 2)
        print(json.dumps(line))
    barrier()          # no rank unmaps its arena while a peer may still read it
    model.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
