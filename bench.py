#!/usr/bin/env python
"""Headline benchmark: batched hyperplane-truncated MVN sampling, k = 1000, k2 = 1 (the reference's
README example shape, BASELINE.json `metric`), samples/s on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref)

A "step" is one pass of the hot path over one batch: ONE call of the batched entry point that
factorises cov once and draws `--batch` samples (per GPU) from per-sample seeded streams, inputs
already resident in HBM (`value`), and the same call through the C ABI with HOST buffers, host<->device
copies inside the timed region (`e2e`).  Prints one JSON line (see README / DESIGN.md section 6).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K, K2 = 1000, 1
METRIC = "samples/sec (hyperplane-truncated MVN k=1000, k2=1, batched, shared covariance)"


def make_inputs(np):
    """README example shape (README.md:120-125: G = 1^T, r = 0, sum-to-zero) on a well-conditioned SPD
    covariance (T T^T / d + I), seeded."""
    import cases
    rng = np.random.default_rng(1)
    cov = cases.spd(rng, K)
    g = np.ones((K2, K))
    r = np.zeros(K2)
    mean = rng.random(K)
    return mean, cov, g, r


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML (nvidia_ml_py) polled every 2 ms from a thread -
    the timed region of the default run is a few tens of milliseconds, shorter than one period of `nvidia-smi -lms` -
    with `nvidia-smi -lms 50` (B200_PROFILING.md recipe) as the fallback.  Samples are time-stamped so that those
    inside [t0, t1] can be picked out."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc, self.nvml, self.stop_flag, self.src = gpu, [], None, None, False, None
        self.sm_max = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except (ValueError, IndexError):
                pass
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml, self.src = pynvml, "nvml, 2 ms period"
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:       # noqa: BLE001  (no NVML: fall back to the CLI)
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self._physical_index())], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.src = "nvidia-smi -lms 50"
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:   # noqa: BLE001  (older binding name)
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.perf_counter(), ("nvml", sm, mask)))
            except Exception:       # noqa: BLE001
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    @property
    def active(self):
        return self.nvml is not None or self.proc is not None

    def count_since(self, t0):
        return sum(1 for t, _ in self.rows if t >= t0)

    def stop(self, t0, t1, note=None):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        reasons = set()
        if self.nvml is not None:
            sm = sorted(r[1] for r in rows)
            mx = [self.sm_max] if self.sm_max else []
            for r in rows:
                for name, bit in self.REASONS:
                    if r[2] & bit:
                        reasons.add(name)
        else:
            rows = [r for r in rows if len(r) > 8]

            def num(x):
                try:
                    return float(x)
                except ValueError:
                    return None
            sm = sorted(v for v in (num(r[1]) for r in rows) if v is not None)
            mx = [v for v in (num(r[2]) for r in rows) if v is not None]
            for r in rows:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
               "reasons": sorted(reasons), "samples": len(sm), "source": self.src}
        if note:
            out["note"] = note
        return out


def cpu_reference_run(nsamples, nthreads, seed0=12345):
    """Time the UNMODIFIED reference (oracle/_ref) on `nsamples` samples of the bench workload using
    `nthreads` host threads (independent samples in parallel, OpenBLAS single-threaded inside each call -
    the best-throughput way to use the cores for one-sample-per-call code).  Runs in a subprocess so the
    OpenBLAS thread setting does not leak.  Falls back to the C oracle port if oracle/_ref is absent."""
    code = f"""
import os, sys, json, time
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, 'tests'))
import numpy as np
sys.argv = ['x']
import bench
mean, cov, g, r = bench.make_inputs(np)
from oracle import Oracle, Reference, have_reference
if have_reference():
    ref = Reference(); kind = 'reference'
    ref.hyperplane('pcg64', {nthreads}, mean, cov, g, r, seed0=1, nthreads={nthreads})   # warm-up
    x, info, rc = ref.hyperplane('pcg64', {nsamples}, mean, cov, g, r, seed0={seed0}, nthreads={nthreads})
    sec = ref.last_seconds
else:
    orc = Oracle(); kind = 'port'
    t0 = time.perf_counter()
    for i in range({nsamples}):     # the port refactorises per call too: one problem per sample
        orc.hyperplane('pcg64', 1, mean, cov, g, r, seed0={seed0} + i)
    sec = time.perf_counter() - t0
print(json.dumps(dict(kind=kind, seconds=sec, nsamples={nsamples})))
"""
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


def cpu_baseline(target_seconds=12.0):
    cores = os.cpu_count() or 1
    probe = cpu_reference_run(cores, cores)
    per_round = max(probe["seconds"], 1e-3)
    rounds = max(1, min(5000, int(target_seconds / per_round)))
    res = cpu_reference_run(cores * rounds, cores)
    n = res["nsamples"]
    if res["kind"] == "port":
        cores_used = 1
    else:
        cores_used = cores
    return {"value": n / res["seconds"], "unit": "samples/s", "cores": cores_used, "kind": res["kind"],
            "sample": f"{n} samples of the bench workload (k={K}, k2={K2}), one reference call per sample, "
                      f"{cores_used} host thread(s), OpenBLAS single-threaded per call, {res['seconds']:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation, all host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = cores * 16
    times = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        res = cpu_reference_run(per_step, cores, seed0=12345 + i * per_step)
        kind = res["kind"]
        if i >= args.warmup:
            times.append(res["seconds"])
    total = sum(times)
    value = per_step * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"hyperplane-truncated MVN k={K}, k2={K2} (G=1^T, r=0), PCG64 per-sample seeds",
                       "samples_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores if kind == "reference" else 1,
                             "kind": kind,
                             "sample": f"{per_step} samples per step, one reference call per sample (the "
                                       "reference refactorises cov on every call)"},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def gather_leg(torch, dist, h, world, rank, dev, rows, k, sample_into, redraw, steps, first_buffer=None, warm=2):
    """`steps` sampler steps, each followed by the all-gather of the rows x k output shards, double-buffered so that
    the gather of step i runs beside step i + 1.  Transport: copy-engine pushes into peer-mapped buffers
    (htnorm_b200.dist.PeerGather - no kernel, no SM taken from the sampling GEMM); NCCL all_gather_into_tensor when
    the peer mapping cannot be set up.  Device-timed, max over ranks.  Checks: every received shard has the checksum
    its owner computed, and a slice of the LAST rank's shard equals the same rows drawn locally (bit for bit)."""
    res = {"value": None}
    ok_local = 1
    outs = pg = full = None
    transport = ("copy engines: one cudaMemcpyAsync per peer shard, ring order on one copy stream, into peer-mapped, double-buffered "
                 "receive buffers (torch symmetric memory); the sampler writes its shard in place; no kernel, no SM "
                 "taken from the sampling GEMM")
    try:
        if os.environ.get("HTN_BENCH_GATHER", "peer") != "peer":
            raise RuntimeError("NCCL requested")
        pg = h.PeerGather(rows, k, dev, nbuf=2)
        ok_local = 1
    except Exception as ex:     # noqa: BLE001
        ok_local = 0
        res["peer_gather_error"] = repr(ex)[:200]
    flag = torch.tensor([ok_local], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if bool(flag.item()):
        # the sampler writes its shard straight into its slice of the (double-buffered) receive buffer: no local copy
        outs = [pg.own(0), pg.own(1)]
        fulls = [pg.full[0], pg.full[1]]
    else:
        pg = None
        transport = "ncclAllGather (all_gather_into_tensor) on NCCL's stream"
        try:
            outs = [first_buffer if first_buffer is not None else torch.empty((rows, k), dtype=torch.float64, device=dev),
                    torch.empty((rows, k), dtype=torch.float64, device=dev)]
            full = torch.empty((world * rows, k), dtype=torch.float64, device=dev)
            ok_local = 1
        except RuntimeError:
            ok_local = 0
        flag = torch.tensor([ok_local], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not bool(flag.item()):
            return {"value": None, "skipped": "shard / receive buffers did not fit"}
        fulls = [full, full]
    pend = [None, None]

    def step(i):
        o = outs[i & 1]
        if pend[i & 1] is not None:             # this shard buffer has left for the peers
            if pg is not None:
                pg.wait(pend[i & 1])
            else:
                pend[i & 1].wait()
        sample_into(i, o)
        if pg is not None:
            pend[i & 1] = pg.push(i & 1)
        else:
            pend[i & 1] = dist.all_gather_into_tensor(full.view(-1), o.view(-1), async_op=True)

    def drain():                                # the current stream has seen every copy / collective of this rank
        for w in pend:
            if w is not None:
                if pg is not None:
                    pg.wait(w)
                else:
                    w.wait()

    def everywhere():
        if pg is not None:
            pg.finish()
        else:
            torch.cuda.synchronize()
            dist.barrier()
    for i in range(warm):
        step(i)
    drain()
    everywhere()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(steps):
        step(warm + i)
    drain()
    g1.record()
    everywhere()
    t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    last_i = warm + steps - 1
    last = outs[last_i & 1]
    full = fulls[last_i & 1]
    # checks on the last step's batch
    own = bool(torch.equal(full[rank * rows:(rank + 1) * rows], last))
    sums = torch.zeros(world, dtype=torch.float64, device=dev)
    sums[rank] = last.sum()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    got = torch.stack([full[q * rows:(q + 1) * rows].sum() for q in range(world)])
    sums_ok = bool(torch.equal(got, sums))
    q, n = world - 1, min(128, rows)
    lo = rows - n
    sub_ok = bool(torch.equal(redraw(last_i, q, lo, n), full[q * rows + lo:q * rows + lo + n]))
    okt = torch.tensor([int(own), int(sums_ok), int(sub_ok)], dtype=torch.int32, device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    res.update({"seconds": float(t.item()) * 1e-3, "ms_per_step": float(t.item()) / steps, "transport": transport,
                "collective": "all-gather of the output shards over NVLink, overlapped with the next step "
                              "(double-buffered shards); every rank ends with all %d samples" % (rows * world),
                "own_shard_intact": bool(okt[0].item()), "peer_shards_match_owner_checksums": bool(okt[1].item()),
                "slice_of_last_rank_equals_local_redraw": bool(okt[2].item())})
    if pg is not None:
        pg.close()
    return res


def cfg5_over_gpus(torch, dist, h, world, rank, dev, peak_tf, steps, no_gather):
    """BASELINE configs[4] as north_star states it: ONE 16384 x 16384 covariance, k2 = 64, 65 536 samples sharded over
    the N GPUs (8192 per rank at N = 8), the set-up (factorisation) replicated on every rank, no data-path collective;
    then the same with the output gather.  Inputs are broadcast from rank 0 (untimed) so that every rank factorises
    the same bits."""
    import bench_configs as bc
    total, k = 65536, 16384
    rows = total // world
    mu, cov, g, r = bc.cfg5_inputs(torch, dev)
    for t in (mu, cov, g, r):
        dist.broadcast(t, src=0)
    out = torch.empty((rows, k), dtype=torch.float64, device=dev)
    info = torch.empty(rows, dtype=torch.int32, device=dev)
    s0 = 12345 + rank * rows

    def sample(i, o):
        h.hyperplane_truncated_mvnorm_batch(rows, mu, cov, g, r, gen="pcg64", seed0=s0 + i * total, out=o, info=info)

    def redraw(i, q, lo, n):
        o = torch.empty((n, k), dtype=torch.float64, device=dev)
        h.hyperplane_truncated_mvnorm_batch(n, mu, cov, g, r, gen="pcg64", seed0=12345 + q * rows + i * total + lo, out=o)
        return o
    sample(0, out)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        sample(1 + i, out)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    bad = torch.tensor([int((info != 0).sum().item())], dtype=torch.int32, device=dev)
    dist.all_reduce(bad, op=dist.ReduceOp.SUM)
    resid = ((out[:256] @ g.T - r).norm(dim=1) / r.norm()).max().reshape(1)
    dist.all_reduce(resid, op=dist.ReduceOp.MAX)
    flops_rank = bc.cfg5_flops(rows)
    tf = flops_rank / (ms * 1e-3) / 1e12
    res = {"workload": "hyperplane-truncated MVN k=16384, k2=64: one covariance, 65 536 samples over %d GPUs (%d per "
                       "rank), factorisation replicated per rank, no data-path collective" % (world, rows),
           "ms_per_step": ms, "samples_per_s": total / (ms * 1e-3), "steps": steps,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s per GPU",
                        "frac": tf / peak_tf if peak_tf and peak_tf > 0 else None, "algorithmic_flops_per_rank": flops_rank,
                        "algorithmic_flops_unreplicated": bc.cfg5_flops(total)},
           "failed_samples": int(bad.item()), "rel_constraint_residual_f64_first_256_per_rank": float(resid.item())}
    if not no_gather:
        gl = gather_leg(torch, dist, h, world, rank, dev, rows, k, sample, redraw, steps, first_buffer=out, warm=1)
        if gl.get("seconds"):
            gl["samples_per_s"] = total * steps / gl.pop("seconds")
            gl["bytes_per_rank_per_step"] = 8 * rows * k
        gl.pop("value", None)
        res["gathered"] = gl
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the figures that include the output gather")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the `configs` object (cfg2 ... cfg5, mid-size batch, generator alone; N = 1) and cfg5 over N GPUs (N > 1)")
    ap.add_argument("--configs", default="cfg2,cfg3,cfg4,cfg5_shard,mid256,normal_fill", help="which entries of `configs` to run")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import htnorm_b200 as h

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner (and INFO lines) to stdout at any
        # NCCL_DEBUG level, so its log goes to a file instead (NCCL_DEBUG=INFO from outside still works, there)
        if "HTN_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["HTN_NCCL_DEBUG"]
        elif os.environ.get("NCCL_DEBUG", "").upper() in ("", "WARN", "VERSION"):
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/htn_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # nvidia-smi needs ~1 s to deliver its first sample

    B = args.batch                      # per GPU: weak scaling, the batch shards naturally
    mean, cov, g, r = make_inputs(np)
    d_in = [torch.from_numpy(x).to(dev) for x in (mean, cov, g, r)]
    out = torch.empty((B, K), dtype=torch.float64, device=dev)      # 524 MB at B = 65536: larger than L2
    info = torch.empty(B, dtype=torch.int32, device=dev)
    seed0 = 12345 + rank * B            # rank q draws samples [q B, (q+1) B) of the global batch

    def step(i):
        h.hyperplane_truncated_mvnorm_batch(B, *d_in, gen="pcg64", seed0=seed0 + i * B * world, out=out, info=info)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    h.phase_timing(True)                    # on during the warm-up too: the phase-mark events are created there
    for i in range(args.warmup):
        step(i)
    barrier()
    # ---- timed: device-resident inputs ----
    h.phase_timing(True)                    # resets the marks
    launches0 = h.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gemm_ms = []
    phases = []
    barrier()
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)               # phase marks are CUDA events: recorded per step, read after the loop
    e1.record()
    barrier()
    launches = h.launch_count() - launches0
    t_wall1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    phases = [h.phase_timing()]             # per-phase averages over the timed steps
    h.phase_timing(False)
    clk = None
    if rank == 0:
        note = None
        if clocks.count_since(t_wall0) < 3 and clocks.active:
            # the timed region is shorter than the sampler's period: keep the SAME steps running (untimed)
            # until a few samples have arrived, and report those
            note = ("timed region (%.0f ms) shorter than the nvidia-smi sampling period; clocks sampled over an "
                    "identical untimed continuation of the same steps" % ms_total)
            t_probe = time.perf_counter()
            i = 0
            while clocks.count_since(t_probe) < 6 and time.perf_counter() - t_probe < 4.0:
                step(args.warmup + args.steps + i)
                i += 1
                if i % 8 == 0:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            t_wall0, t_wall1 = t_probe, time.perf_counter()
        clk = clocks.stop(t_wall0, t_wall1, note)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = B * world * args.steps / (ms_total * 1e-3)
    gemm_ms = [p["gemm"] for p in phases if p["gemm"] > 0]

    # parity spot check inside the bench: constraint residual of the last batch (sum-to-zero)
    resid = float((out @ d_in[2].T - d_in[3]).abs().max().item())

    # ---- timed (N > 1 only): the same steps followed by the one collective of the path, the all-gather of the
    # output shards.  Shards are double-buffered, so the gather of step i runs beside step i + 1. ----
    gathered = None
    if world > 1 and not args.no_gather:
        def sample_headline(i, o):
            h.hyperplane_truncated_mvnorm_batch(B, *d_in, gen="pcg64", seed0=seed0 + i * B * world, out=o, info=info)

        def redraw_headline(i, q, lo, n):      # rows [lo, lo + n) of rank q's shard of step i, drawn on THIS rank
            o = torch.empty((n, K), dtype=torch.float64, device=dev)
            h.hyperplane_truncated_mvnorm_batch(n, *d_in, gen="pcg64", seed0=12345 + q * B + i * B * world + lo, out=o)
            return o
        gathered = gather_leg(torch, dist, h, world, rank, dev, B, K, sample_headline, redraw_headline, args.steps,
                              first_buffer=out)
        if gathered.get("seconds"):
            gathered["value"] = B * world * args.steps / gathered.pop("seconds")
            gathered["unit"] = "samples/s"
            gathered["bytes_per_rank_per_step"] = 8 * B * K

    # ---- timed: end to end through the C ABI with HOST buffers (pinned), copies inside ----
    h_in = [torch.from_numpy(x).pin_memory() for x in (mean, cov, g, r)]
    h_out = torch.empty((B, K), dtype=torch.float64).pin_memory()
    h_info = torch.empty(B, dtype=torch.int32).pin_memory()
    np_in = [x.numpy() for x in h_in]

    def step_host(i):
        h.hyperplane_truncated_mvnorm_batch(B, *np_in, gen="pcg64", seed0=seed0 + i * B * world, out=h_out.numpy(),
                                            info=h_info.numpy())
    e2e_steps = max(2, min(args.steps, 5))
    step_host(0)
    step_host(1)                        # two untimed calls: pool growth, first touch of the pinned buffers
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_host(2 + i)                # synchronous: returns when the samples are in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / float(t.item())
    h2d = 8 * (K * K + K2 * K + K + K2)
    d2h = 8 * B * K + 4 * B
    # the ceiling of that figure on THIS box: every rank copies one shard-sized buffer device -> pinned host at the same
    # time, nothing else running (N = 1: the bare PCIe rate; N > 1: what the host side absorbs from N GPUs at once)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        h_out.copy_(out, non_blocking=True)
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    d2h_ceiling_gbs = world * 8.0 * B * K / float(t.item()) / 1e9

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    measured = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    dmma_peak = h.fp64_tensor_peak_tflops(local, 100.0)         # in-run FP64 tensor (DMMA) peak
    dfma_peak = h.fp64_fma_peak_tflops(local, 50.0) if rank == 0 else None
    # ---- the other BASELINE configurations (not the headline): N = 1 -> cfg2, cfg3, cfg4, cfg5's shard, the generator
    # alone; N > 1 -> cfg5 as north_star states it, sharded over the N GPUs, without and with the output gather ----
    configs = None
    if not args.no_configs:
        import bench_configs
        del out, info
        torch.cuda.empty_cache()
        if world == 1:
            hbm = float(measured.get("hbm_gbs", bench_configs.HBM_FALLBACK_GBS))
            configs = bench_configs.run_all(h, torch, dev, dmma_peak, hbm, tuple(args.configs.split(",")))
            configs["peaks"] = {"fp64_tensor_tflops": dmma_peak, "hbm_gbs": hbm,
                                "hbm_source": "MEASURED_PEAKS.json (of measured)" if "hbm_gbs" in measured
                                else "B200_PROFILING.md fallback (of fallback)"}
        else:
            try:
                configs = {"cfg5": cfg5_over_gpus(torch, dist, h, world, rank, dev, dmma_peak, 2, args.no_gather)}
            except Exception as ex:     # noqa: BLE001  (reported in place; the headline stands)
                configs = {"cfg5": {"error": repr(ex)[:300]}}
    if rank == 0:
        alg_flops = B * (K * K + 2.0 * K * K2)                  # per launch of the sampling GEMM
        gemm_avg = sum(gemm_ms) / len(gemm_ms) if gemm_ms else None
        achieved = alg_flops / (gemm_avg * 1e-3) / 1e12 if gemm_avg else None
        roof = {"bound": "tensor", "kernel": "gemm_tma_kernel (sampling GEMM out = Z~ F^T + mean; TMA + mbarrier ring, DMMA m8n8k4 f64)",
                "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
                "frac": achieved / dmma_peak if achieved and dmma_peak > 0 else None,
                "traffic": None,
                "peak_source": "in-run DMMA (mma.sync m8n8k4 f64) microbenchmark, htn_fp64_tensor_peak_tflops; "
                               "MEASURED_PEAKS.json has no FP64 entry (hbm_gbs=%s)" % measured.get("hbm_gbs"),
                "fp64_fma_peak_tflops": dfma_peak,
                "kernel_ms": gemm_avg,
                "step_phases_ms": {k_: v for k_, v in phases[0].items() if k_ != "calls"},
                "kernel_launches_timed": phases[0]["calls"]}
        traffic_path = os.path.join(ROOT, "profiles", "gemm_traffic_bytes.json")
        if os.path.exists(traffic_path):
            tj = json.load(open(traffic_path))
            roof["traffic"] = tj.get("dram_bytes_per_launch")
            roof["traffic_source"] = "one `ncu --set full` capture of this kernel, not this run: " + str(tj.get("source"))
            roof["algorithmic_bytes"] = tj.get("algorithmic_bytes")
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"hyperplane-truncated MVN k={K}, k2={K2} (G=1^T, r=0; README example "
                                       "shape), well-conditioned SPD cov, PCG64 per-sample seeds, one "
                                       "factorisation + batch per step",
                           "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2": "output (524 MB per step at the default batch) and Z~ exceed the 126 MB L2"},
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "d2h_gbs": e2e_value * 8.0 * K / 1e9, "d2h_ceiling_gbs": d2h_ceiling_gbs,
                        "frac_of_d2h_ceiling": e2e_value * 8.0 * K / 1e9 / d2h_ceiling_gbs,
                        "d2h_ceiling_how": "all %d ranks copy one 524 MB shard device -> pinned host concurrently, "
                                           "3 repetitions, max over ranks" % world},
                "gpu_launches": launches, "roofline": roof, "clocks": clk,
                "max_abs_constraint_residual": resid}
        if gathered is not None:
            line["gathered"] = gathered
        if configs is not None:
            line["configs"] = configs
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline()
            except Exception as ex:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": 0, "kind": "unavailable",
                                        "sample": repr(ex)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
