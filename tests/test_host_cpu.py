"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/aquagym.h
declares, argument validation and the no-fallback behaviour, host-side table/slab logic, and the
world_size-2 statistics all-reduce over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests.conftest import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from aquaticgymenv_b200.build import build
    build()
    from aquaticgymenv_b200 import _native
    return _native


def _header_functions():
    src = open(os.path.join(ROOT, "include", "aquagym.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aqua_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(native):
    lib = native.load()
    declared = _header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libaquagym.so does not export %s" % name
    assert sorted(native.SIGNATURES) == declared, "binding table and header disagree"
    assert lib.aqua_abi_version() == native.ABI_VERSION
    # only the C ABI is exported with C linkage: nm shows the symbols unmangled
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(r"\bT %s\b" % name, out), name


def test_config_struct_matches_header(native):
    # sizeof(aqua_config) as the C compiler sees it == ctypes mirror
    prog = '#include <stdio.h>\n#include "aquagym.h"\nint main(){printf("%zu %zu", sizeof(aqua_config), sizeof(aqua_state_view));return 0;}'
    exe = os.path.join(ROOT, ".pytest_cache", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog,
                   text=True, check=True)
    a, b = map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split())
    assert a == C.sizeof(native.Config) and b == C.sizeof(native.StateView)


def test_no_cpu_fallback_and_validation(native):
    import torch
    lib = native.load()
    cfg = native.Config()
    h = C.c_void_p()
    cfg.struct_size = 12  # wrong ABI size
    assert lib.aqua_create(C.byref(h), C.byref(cfg)) == native.ERR_INVALID
    assert b"struct_size" in lib.aqua_last_error()
    cfg.struct_size = C.sizeof(native.Config)
    cfg.n_envs = 0
    assert lib.aqua_create(C.byref(h), C.byref(cfg)) == native.ERR_INVALID
    cfg.n_envs = 16
    cfg.n_obstacles = 65
    assert lib.aqua_create(C.byref(h), C.byref(cfg)) == native.ERR_INVALID
    cfg.n_obstacles = 0
    cfg.precision = 7
    assert lib.aqua_create(C.byref(h), C.byref(cfg)) == native.ERR_INVALID
    cfg.precision = native.F32
    cfg.env_id_base = (1 << 32) - 4
    assert lib.aqua_create(C.byref(h), C.byref(cfg)) == native.ERR_INVALID
    cfg.env_id_base = 0
    bad = (C.c_double * 5)(1.0, 2.0, 3.0, 1.0, 1.0)  # type 3 is neither circle nor rectangle
    cfg.n_obstacles, cfg.obstacles = 1, bad
    assert lib.aqua_create(C.byref(h), C.byref(cfg)) == native.ERR_INVALID
    cfg.n_obstacles, cfg.obstacles = 0, None
    if not torch.cuda.is_available():
        rc = lib.aqua_create(C.byref(h), C.byref(cfg))
        assert rc == native.ERR_NO_DEVICE and not h
        assert b"no CPU fallback" in lib.aqua_last_error()
        from aquaticgymenv_b200 import AquaVecEnv, gym_aqua
        with pytest.raises(native.AquaError):
            AquaVecEnv(8)
        with pytest.raises(native.AquaError):
            gym_aqua.make("AquaEnv-v0")
        with pytest.raises(ValueError):
            AquaVecEnv(8, device="cpu")
    # NULL handle is an error, not a crash
    assert lib.aqua_reset(None, None, None) == native.ERR_INVALID
    assert lib.aqua_destroy(None) == native.OK


def test_build_script_loads_without_the_package():
    """__graft_entry__.build() must be able to build a fresh checkout: the package refuses to
    import before libaquagym.so exists, so the build script is loaded by path and must not pull
    the package in."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import importlib.util, os, sys\n"
            "spec = importlib.util.spec_from_file_location('_b', os.path.join(%r, 'aquaticgymenv_b200', 'build.py'))\n"
            "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
            "assert callable(m.build) and 'aquaticgymenv_b200' not in sys.modules\n" % root)
    subprocess.run([sys.executable, "-c", code], check=True)
    src = open(os.path.join(root, "__graft_entry__.py")).read()
    assert "spec_from_file_location" in src and "from aquaticgymenv_b200.build import" not in src


def test_product_package_never_touches_the_oracle():
    """The product path must not import or load anything under oracle/."""
    pkg = os.path.join(ROOT, "aquaticgymenv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "aqua_device.cuh", (dirpath, f)
    code = ("import sys; sys.path.insert(0, %r); import aquaticgymenv_b200; "
            "assert not [m for m in sys.modules if m.startswith('oracle')]" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_obstacle_table_conversions(native):
    from aquaticgymenv_b200 import DEFAULT_OBSTACLES, DIFFICULT_OBSTACLES, obstacle_table
    from aquaticgymenv_b200.vec_env import table_to_reference_list
    assert np.array_equal(DEFAULT_OBSTACLES, np.load(os.path.join(GOLDEN, "v1_table.npy")))
    assert np.array_equal(DIFFICULT_OBSTACLES, np.load(os.path.join(GOLDEN, "v2_table.npy")))
    assert obstacle_table(False).shape == (0, 5)
    assert np.array_equal(obstacle_table(True), DEFAULT_OBSTACLES)
    lst = table_to_reference_list(DIFFICULT_OBSTACLES)
    assert lst[1][1] == "r" and lst[1][2] == (10.0, 10.0) and lst[0][1] == "c"
    assert np.array_equal(obstacle_table(lst), DIFFICULT_OBSTACLES)
    weird = [(np.array([1, 2]), "x", (3, 4))]  # non-'c' -> rectangle (aqua.py:434-435)
    assert obstacle_table(weird).tolist() == [[1.0, 2.0, 1.0, 3.0, 4.0]]
    with pytest.raises(ValueError):
        obstacle_table(np.zeros((65, 5)))


def test_registry_matches_reference_ids(native):
    from aquaticgymenv_b200 import gym_aqua
    assert sorted(gym_aqua.REGISTRY) == sorted(
        ["AquaEnv-v0", "AquaEnv-v1", "AquaEnv-v2", "AquaContinuousEnv-v0", "AquaContinuousEnv-v1",
         "AquaContinuousEnv-v2"])
    assert gym_aqua.REGISTRY["AquaEnv-v1"][2] == {"obstacles": True}
    assert len(gym_aqua.REGISTRY["AquaContinuousEnv-v2"][2]["obstacles"]) == 6
    with pytest.raises(KeyError):
        gym_aqua.make("AquaEnv-v9")


def test_spaces_surface():
    from aquaticgymenv_b200 import spaces
    box = spaces.Box(np.array([0, 0, -np.pi, 0, 0]), np.array([100, 100, np.pi, 100, 100]), dtype=np.float64)
    assert box.shape == (5,) and box.high[2] == np.pi and box.low[2] == -np.pi
    assert box.contains(box.sample()) and not box.contains(np.array([0, 0, 4.0, 0, 0]))
    act = spaces.Box(0.2, 0.5, shape=[2], dtype=np.float64)
    assert act.contains(np.array([0.2, 0.5])) and not act.contains(np.array([0.1, 0.5]))
    d = spaces.Discrete(3)
    assert d.n == 3 and all(0 <= d.sample() < 3 for _ in range(20))


def test_slab_partition():
    from aquaticgymenv_b200 import sharding
    for total, world in [(1 << 20, 8), (1 << 22, 8), (10, 3), (7, 8), (65536, 1)]:
        got = [sharding.slab(total, r, world) for r in range(world)]
        assert got[0][0] == 0 and sum(n for _, n in got) == total
        for (b0, n0), (b1, _) in zip(got, got[1:]):
            assert b0 + n0 == b1
        assert max(n for _, n in got) - min(n for _, n in got) <= 1
    with pytest.raises(ValueError):
        sharding.slab(8, 8, 8)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from aquaticgymenv_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
base, n = sharding.slab(1001, rank, 2)
stats = torch.tensor([n, rank, 0, 1, 10 * n, -5 * (rank + 1), 7 * n, 0, rank], dtype=torch.int64)
sharding.all_reduce_stats(stats)
assert stats.tolist() == [1001, 1, 0, 2, 10010, -15, 7007, 0, 1], stats
w = sharding.all_reduce_stats(torch.ones(9, dtype=torch.int64), async_op=True)
w.wait()
try:
    sharding.all_reduce_stats(torch.ones(9))
    raise SystemExit("float stats accepted")
except ValueError:
    pass
dist.destroy_process_group()
print("rank", rank, "ok", base, n)
"""


def test_stats_allreduce_gloo_world2(native, tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "rank 0 ok 0 501" in outs[0] and "rank 1 ok 501 500" in outs[1]
