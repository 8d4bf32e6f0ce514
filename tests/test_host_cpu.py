"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the public headers
declare, struct layouts match the reference's, the generator constructors reproduce the reference
streams, argument errors surface as the reference's exceptions, a missing GPU fails loudly (no CPU
fallback), and the multi-GPU sharding logic gathers identical results under gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def h():
    import htnorm_b200
    return htnorm_b200


def declared_functions():
    names = set()
    for hdr in ("htnorm.h", "htnorm_rng.h", "htnorm_batch.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b([a-z_0-9]+)\s*\([^;{]*\)\s*;", src):
            if m.group(1).startswith(("htn_", "rng_", "init_")):
                names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol(h):
    names = declared_functions()
    assert {"htn_hyperplane_truncated_mvn", "htn_structured_precision_mvn", "init_ht_config",
            "init_sp_config", "rng_free", "rng_pcg64_new", "rng_pcg64_new_seeded", "rng_xrs128p_new",
            "rng_xrs128p_new_seeded", "htn_hyperplane_truncated_mvn_batch",
            "htn_structured_precision_mvn_batch"} <= names
    for n in names:
        assert hasattr(h.lib, n), f"{n} declared in include/ but not exported by libhtnorm_b200.so"
    assert set(h.PUBLIC_SYMBOLS) <= names


def test_struct_layouts_match_reference(h):
    # sizes probed on the compiled reference (x86-64)
    assert C.sizeof(h._lib.RngT) == 24
    assert C.sizeof(h._lib.HtConfig) == 56
    assert C.sizeof(h._lib.SpConfig) == 64
    assert h._lib.SpConfig.a_id.offset == 0 and h._lib.SpConfig.o_id.offset == 4   # enums first
    assert h._lib.HtConfig.diag.offset == 48


@pytest.mark.parametrize("gen", ["pcg64", "xrs128p"])
def test_host_generators_match_reference_streams(h, oracle, gen):
    """rng_*_new_seeded hand out objects whose function pointers work on the host exactly like the
    reference's (a foreign caller may call them directly)."""
    for seed in (0, 10, 12345, 2**64 - 1):
        g = h.HTNGenerator(seed, gen)
        rng = g._rng.contents
        f = C.CFUNCTYPE(C.c_uint64, C.c_void_p)(rng.next_uint64)
        d = C.CFUNCTYPE(C.c_double, C.c_void_p)(rng.next_double)
        got = [f(rng.base) for _ in range(50)]
        want = oracle.raw(gen, 1, 51, seed0=seed)[0]
        assert got == [int(v) for v in want[:50]]
        assert d(rng.base) == float(int(want[50]) >> 11) * (1.0 / 9007199254740992.0)


def test_argument_errors_before_any_device_work(h):
    k = 8
    cov = np.eye(k)
    with pytest.raises(ValueError, match="C-contiguous"):       # reference tests/test_pyhtnorm.py:37-38
        h.hyperplane_truncated_mvnorm(np.zeros(k), np.asfortranarray(np.random.rand(k, k)), np.ones((1, k)),
                                      np.zeros(1))
    with pytest.raises(ValueError, match="incompatible shapes"):  # :41-43
        h.hyperplane_truncated_mvnorm(np.zeros(k), cov, np.ones((1, k + 1)), np.zeros(1))
    with pytest.raises(ValueError):
        h.structured_precision_mvnorm(np.zeros(k), cov, np.ones((2, k)), np.eye(2), a_type=-1000)   # :86-89
    with pytest.raises(ValueError):
        h.HTNGenerator(-1)                                       # :121-124
    with pytest.raises(ValueError):
        h.HTNGenerator(1.5)
    with pytest.raises(ValueError, match="incompatible shapes"):
        h.hyperplane_truncated_mvnorm_batch(4, np.zeros(k), cov, np.ones((1, k + 1)), np.zeros(1))


def test_no_gpu_fails_loudly(h):
    """the product has no CPU path: without a device every sampling call raises"""
    if h.lib.htn_device_count() > 0:
        pytest.skip("a GPU is present")
    k = 8
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        h.hyperplane_truncated_mvnorm_batch(4, np.zeros(k), np.eye(k), np.ones((1, k)), np.zeros(1))
    with pytest.raises(RuntimeError):
        h.raw_streams(2, 4)
    with pytest.raises(RuntimeError):
        h.hyperplane_truncated_mvnorm(np.zeros(k), np.eye(k), np.ones((1, k)), np.zeros(1), random_state=1)


def test_product_does_not_reference_the_oracle():
    """the oracle is test infrastructure: nothing under htnorm_b200/ may import, link or call it"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "htnorm_b200")):
        if "_obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h", ".cuh", ".map")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\boracle\b|liboracle|_ref/", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_shard_ranges(h):
    for n in (0, 1, 7, 64, 65536, 100003):
        for w in (1, 2, 3, 4, 8):
            r = [h.shard_range(n, q, w) for q in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[q][1] == r[q + 1][0] for q in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == h.shard_sizes(n, w)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch, torch.distributed as dist
    import cases
    from functools import partial
    from oracle import Oracle
    from htnorm_b200.dist import sample_sharded, gather_rows
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    c = cases.ht_case("k4_dense_64")
    orc = Oracle()
    def sampler(n, seeds=None, seed0=0):
        # stand-in for the CUDA batch call with the same signature and the same per-sample seeding rule
        x, info, _ = orc.hyperplane("pcg64", n, c["mean"], c["cov"], c["g"], c["r"], seeds=seeds, seed0=seed0)
        return torch.from_numpy(x), torch.from_numpy(info)
    B = 37   # ragged: shards of 19 and 18
    seeds = np.arange(1000, 1000 + B, dtype=np.uint64) * 7919
    for kw in (dict(seed0=12345), dict(seeds=seeds)):
        local, info = sample_sharded(sampler, B, rank, world, **kw)
        full = gather_rows(local, B)
        want = orc.hyperplane("pcg64", B, c["mean"], c["cov"], c["g"], c["r"], **kw)[0]
        assert full.shape == (B, 64) and np.array_equal(full.numpy(), want), "sharded != unsharded"
        top = gather_rows(local, B, dst=0)
        assert (top is None) == (rank != 0)
        if rank == 0:
            assert np.array_equal(top.numpy(), want)
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_sharded_sampling_is_independent_of_world_size_gloo(tmp_path):
    """world_size 2 over gloo on CPU: rank q draws its contiguous slice with offset seeds; the gathered
    result equals the single-process result bit for bit (the only collective is the output gather)."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


def test_batch_result_buffers_are_validated(h):
    """`out` / `info` reach the C ABI as raw pointers, so shape, dtype, contiguity and memory space are
    checked in the wrapper - before any device work (this runs without a GPU)."""
    k = 8
    args = (np.zeros(k), np.eye(k), np.ones((1, k)), np.zeros(1))
    for bad in (np.empty((4, k), dtype=np.float32), np.empty((3, k)), np.empty((4, k + 1)),
                np.empty((k, 4)).T, np.empty((4, 2 * k))[:, ::2]):
        with pytest.raises(ValueError, match="`out`"):
            h.hyperplane_truncated_mvnorm_batch(4, *args, out=bad)
    with pytest.raises(ValueError, match="`info`"):
        h.hyperplane_truncated_mvnorm_batch(4, *args, info=np.empty(4, dtype=np.int64))
    with pytest.raises(ValueError, match="`info`"):
        h.structured_precision_mvnorm_batch(4, np.zeros(k), np.eye(k), np.ones((2, k)), np.eye(2),
                                            info=np.empty(3, dtype=np.int32))
    with pytest.raises(TypeError):
        h.hyperplane_truncated_mvnorm_batch(4, *args, out=[[0.0] * k] * 4)
    with pytest.raises(ValueError, match="seeds"):
        h.hyperplane_truncated_mvnorm_batch(4, *args, seeds=np.arange(3, dtype=np.uint64))


def test_product_library_exports_only_the_declared_boundary(h):
    """W7 / ADVICE r1: the release library exports the symbols include/*.h declares and nothing else; the kernel-level
    test hooks (htn_dbg_*) and the host layer's internals live in libhtnorm_b200_dev.so only."""
    out = subprocess.run(["nm", "-D", "--defined-only", h.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.split()[1] in "TDBR"}
    assert exported == declared_functions(), exported ^ declared_functions()
    dev = h._lib.load_dev()
    assert hasattr(dev, "htn_dbg_gemm_tn") and not hasattr(h.lib, "htn_dbg_gemm_tn")
