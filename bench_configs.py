"""The other BASELINE.json configurations, measured by bench.py next to the headline (the `configs` object of its
JSON line): cfg2 (2^20 independent small problems), cfg3 / cfg4 (structured precision), cfg5's one-GPU shard, 4096
independent mid-size problems (k = 256) in one call, and the standalone ziggurat generator with BOTH of its rooflines.  Device-resident synthetic inputs (SURVEY 8d shapes),
CUDA events on the launching stream, a few steps each.  `frac` = SURVEY 8(d)'s algorithmic flops / bytes of the
configuration divided by the measured time, over the measured peak (FP64 tensor: in-run DMMA microbenchmark; HBM:
MEASURED_PEAKS.json).  Not the headline: `value` of the bench line stays the k = 1000 workload."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

HBM_FALLBACK_GBS = 6650.0     # B200_PROFILING.md fallback, used only when MEASURED_PEAKS.json is absent


def _time(torch, fn, warmup=2, steps=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _tensor_entry(name, ms, nsamples, flops, peak_tf, extra=None):
    tf = flops / (ms * 1e-3) / 1e12
    d = {"workload": name, "ms": ms, "samples_per_s": nsamples / (ms * 1e-3),
         "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s",
                      "frac": tf / peak_tf if peak_tf > 0 else None, "algorithmic_flops": flops}}
    if extra:
        d.update(extra)
    return d


def spd_lowrank_device(torch, d, dev, seed, rank=64):
    """diag(U(1,2)) + V V^T / d on the device (input synthesis only: torch.matmul is not on the measured path)"""
    g = torch.Generator(device=dev).manual_seed(seed)
    v = torch.randn(d, rank, dtype=torch.float64, device=dev, generator=g)
    s = (v @ v.T) / d
    s.diagonal().add_(1.0 + torch.rand(d, dtype=torch.float64, device=dev, generator=g))
    return s


def cfg2(h, torch, dev, hbm_gbs, nproblems=1 << 20):
    """configs[1]: k = 64, k2 = 4, 2^20 independent problems (own mean / cov / G / r each): HBM-bound on paper"""
    import cases
    k, k2 = 64, 4
    rng = np.random.default_rng(2)
    base = torch.from_numpy(np.stack([cases.spd(rng, k) for _ in range(64)])).to(dev)
    g_ = torch.Generator(device=dev).manual_seed(2)
    cov = base.repeat(nproblems // 64, 1, 1)
    cov.mul_(1.0 + torch.rand(nproblems, 1, 1, dtype=torch.float64, device=dev, generator=g_))   # distinct matrices
    g = torch.randn(nproblems, k2, k, dtype=torch.float64, device=dev, generator=g_)
    r = torch.randn(nproblems, k2, dtype=torch.float64, device=dev, generator=g_)
    mu = torch.randn(nproblems, k, dtype=torch.float64, device=dev, generator=g_)
    out = torch.empty(nproblems, k, dtype=torch.float64, device=dev)
    info = torch.empty(nproblems, dtype=torch.int32, device=dev)
    ms = _time(torch, lambda: h.hyperplane_truncated_mvnorm_batch(nproblems, mu, cov, g, r, independent=True,
                                                                   seed0=12345, out=out, info=info))
    bad = int((info != 0).sum().item())
    resid = float((torch.einsum("bck,bk->bc", g[:4096], out[:4096]) - r[:4096]).abs().max().item())
    nbytes = 8.0 * (k * k + k2 * k + 2 * k + k2) * nproblems          # 35 872 B per problem
    flops = 127371.0 * nproblems
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"workload": "hyperplane-truncated MVN k=64, k2=4, %d independent problems" % nproblems, "ms": ms,
            "samples_per_s": nproblems / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs,
                         "algorithmic_bytes": nbytes, "algorithmic_flops": flops,
                         "fp64_tflops": flops / (ms * 1e-3) / 1e12},
            "failed_problems": bad, "max_abs_constraint_residual_first_4096": resid}


def mid256(h, torch, dev, peak_tf, nproblems=4096, k=256, k2=4):
    """independent MID-SIZE problems (own mean / cov / G / r each, one sample per problem) in one call:
    the batched left-looking factorisation + projection of csrc/dev/mid_batched.cu.  Flops per problem as SURVEY 8(d) counts
    them for one hyperplane call with its own covariance: k^3/3 + k^2 + 2 k^2 k2 + 2 k k2^2 + k2^3/3 + 2 k2^2 + 4 k k2."""
    import cases
    rng = np.random.default_rng(k)
    base = torch.from_numpy(np.stack([cases.spd(rng, k) for _ in range(8)])).to(dev)
    g_ = torch.Generator(device=dev).manual_seed(9)
    cov = base.repeat(nproblems // 8, 1, 1)
    cov.mul_(1.0 + torch.rand(nproblems, 1, 1, dtype=torch.float64, device=dev, generator=g_))
    g = torch.randn(nproblems, k2, k, dtype=torch.float64, device=dev, generator=g_)
    r = torch.randn(nproblems, k2, dtype=torch.float64, device=dev, generator=g_)
    mu = torch.randn(nproblems, k, dtype=torch.float64, device=dev, generator=g_)
    out = torch.empty(nproblems, k, dtype=torch.float64, device=dev)
    info = torch.empty(nproblems, dtype=torch.int32, device=dev)
    ms = _time(torch, lambda: h.hyperplane_truncated_mvnorm_batch(nproblems, mu, cov, g, r, independent=True, seed0=12345,
                                                                   out=out, info=info))
    flops = nproblems * (k ** 3 / 3 + k * k + 2 * k * k * k2 + 2 * k * k2 * k2 + k2 ** 3 / 3 + 2 * k2 * k2 + 4 * k * k2)
    resid = float((torch.einsum("bck,bk->bc", g[:256], out[:256]) - r[:256]).abs().max().item())
    return _tensor_entry("hyperplane-truncated MVN k=%d, k2=%d, %d independent problems (one batched call)"
                         % (k, k2, nproblems), ms, nproblems, flops, peak_tf,
                         {"failed_problems": int((info != 0).sum().item()), "max_abs_constraint_residual_first_256": resid})


def cfg3(h, torch, dev, peak_tf):
    """configs[2]: p = 2000, n = 500, diagonal Omega, 4096 samples sharing A and Phi"""
    import cases
    rng = np.random.default_rng(3)
    p, n, B = 2000, 500, 4096
    a = torch.from_numpy(cases.spd(rng, p)).to(dev)
    phi = torch.from_numpy(rng.standard_normal((n, p)) / np.sqrt(p)).to(dev)
    om = torch.from_numpy(np.diag(rng.uniform(0.5, 1.5, n))).to(dev)
    mu = torch.from_numpy(rng.standard_normal(p)).to(dev)
    out = torch.empty(B, p, dtype=torch.float64, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    ms = _time(torch, lambda: h.structured_precision_mvnorm_batch(B, mu, a, phi, om, False, 0, 1, seed0=12345, out=out,
                                                                   info=info), steps=5)
    return _tensor_entry("structured precision p=2000, n=500, diagonal Omega, 4096 samples", ms, B, 4.253e10, peak_tf,
                         {"finite": bool(torch.isfinite(out).all().item())})


def cfg4(h, torch, dev, peak_tf):
    """configs[3]: structured precision + structured mean, p = 8192, n = 1024, dense A / Omega, batch 256"""
    p, n, B = 8192, 1024, 256
    a = spd_lowrank_device(torch, p, dev, 41)
    om = spd_lowrank_device(torch, n, dev, 42)
    g_ = torch.Generator(device=dev).manual_seed(43)
    phi = torch.randn(n, p, dtype=torch.float64, device=dev, generator=g_) / np.sqrt(p)
    t = torch.randn(n, dtype=torch.float64, device=dev, generator=g_)
    out = torch.empty(B, p, dtype=torch.float64, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    ms = _time(torch, lambda: h.structured_precision_mvnorm_batch(B, t, a, phi, om, True, 0, 0, seed0=12345, out=out,
                                                                   info=info))
    return _tensor_entry("structured precision + structured mean p=8192, n=1024, dense A/Omega, 256 samples", ms, B,
                         3.659e11, peak_tf, {"finite": bool(torch.isfinite(out).all().item())})


def cfg5_inputs(torch, dev):
    k, k2 = 16384, 64
    cov = spd_lowrank_device(torch, k, dev, 51)
    g_ = torch.Generator(device=dev).manual_seed(52)
    g = torch.randn(k2, k, dtype=torch.float64, device=dev, generator=g_) / np.sqrt(k)
    r = torch.randn(k2, dtype=torch.float64, device=dev, generator=g_)
    mu = torch.randn(k, dtype=torch.float64, device=dev, generator=g_)
    return mu, cov, g, r


def cfg5_flops(B):
    k, k2 = 16384.0, 64.0
    setup = k ** 3 / 3 + 2 * k * k * k2 + 2 * k * k2 * k2 + k2 ** 3 / 3
    return setup + B * (k * k + 4 * k * k2 + 2 * k2 * k2)


def cfg5_shard(h, torch, dev, peak_tf, B=8192):
    """configs[4], one GPU's share: k = 16384, k2 = 64, one factorisation + 8192 of the 65 536 samples"""
    mu, cov, g, r = cfg5_inputs(torch, dev)
    out = torch.empty(B, 16384, dtype=torch.float64, device=dev)
    info = torch.empty(B, dtype=torch.int32, device=dev)
    ms = _time(torch, lambda: h.hyperplane_truncated_mvnorm_batch(B, mu, cov, g, r, seed0=12345, out=out, info=info),
               warmup=1, steps=2)
    resid = float(((out[:512] @ g.T - r).norm(dim=1) / r.norm()).max().item())
    return _tensor_entry("hyperplane-truncated MVN k=16384, k2=64: one factorisation + %d samples (one GPU's shard of "
                         "65 536)" % B, ms, B, cfg5_flops(B), peak_tf,
                         {"rel_constraint_residual_first_512_f64": resid, "failed": int((info != 0).sum().item())})


def normal_fill(h, torch, dev, hbm_gbs, issue_peak_gops, nstreams=65536, n=1000):
    """the generator alone (the headline's Z producer): 65 536 PCG64 streams x 1000 ziggurat normals.  Two rooflines:
    HBM (8 B written per normal) and integer ISSUE (the kernel's executed warp-instructions per 32 normals, from the
    committed ncu capture, over the issue rate measured in this run with the generator's instruction mix)."""
    import ctypes as C
    import json
    out = torch.empty(nstreams, n, dtype=torch.float64, device=dev)
    b = h._lib.Batch()
    h.lib.htn_batch_init(C.byref(b), C.c_size_t(nstreams))
    b.gen, b.seed0, b.mem, b.device = 0, 12345, h._lib.HTN_MEM_DEVICE, torch.cuda.current_device()
    b.stream = torch.cuda.current_stream().cuda_stream

    def fn():
        rc = h.lib.htn_std_normal_fill(C.byref(b), C.c_size_t(n), C.c_void_p(out.data_ptr()), None)
        assert rc == 0
    ms = _time(torch, fn, steps=5)
    normals = float(nstreams) * n
    gbs = 8.0 * normals / (ms * 1e-3) / 1e9
    entry = {"workload": "ziggurat normals alone: %d PCG64 streams x %d" % (nstreams, n), "ms": ms,
             "normals_per_s": normals / (ms * 1e-3),
             "roofline_hbm": {"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs}}
    path = os.path.join(ROOT, "profiles", "normal_fill_instr.json")
    if os.path.exists(path) and issue_peak_gops and issue_peak_gops > 0:
        w = json.load(open(path))
        per32 = float(w["warp_instructions_per_32_normals"])
        ach = normals / 32.0 * per32 / (ms * 1e-3) / 1e9
        entry["roofline_issue"] = {"bound": "issue", "achieved": ach, "peak": issue_peak_gops,
                                   "unit": "G warp-instructions/s", "frac": ach / issue_peak_gops,
                                   "warp_instructions_per_32_normals": per32, "source": w.get("source")}
    return entry


def run_all(h, torch, dev, peak_tf, hbm_gbs, which=("cfg2", "cfg3", "cfg4", "cfg5_shard", "mid256", "normal_fill")):
    """Every entry is computed independently: a failure is reported in place and does not take the bench line down."""
    out = {}
    issue = None
    if "normal_fill" in which:
        try:
            issue = h.int_issue_peak_gops(torch.cuda.current_device(), 50.0)
        except Exception:   # noqa: BLE001
            issue = None
    table = {"cfg2": lambda: cfg2(h, torch, dev, hbm_gbs), "cfg3": lambda: cfg3(h, torch, dev, peak_tf),
             "cfg4": lambda: cfg4(h, torch, dev, peak_tf), "cfg5_shard": lambda: cfg5_shard(h, torch, dev, peak_tf),
             "mid256": lambda: mid256(h, torch, dev, peak_tf),
             "normal_fill": lambda: normal_fill(h, torch, dev, hbm_gbs, issue)}
    for name in which:
        try:
            out[name] = table[name]()
        except Exception as ex:     # noqa: BLE001
            out[name] = {"error": repr(ex)[:300]}
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    return out
