"""The C-ABI shared library loads and exports exactly what include/gymchess_b200.h declares (no GPU needed)."""
import os
import re

import pytest

from gym_chess_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gymchess_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcb_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.SO_PATH), "run __graft_entry__.build() first"
    L = _lib.lib()
    assert L.gcb_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "declared in the header but not exported: " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert sorted(_lib.SIGNATURES) == names


def test_no_cpu_fallback_without_gpu():
    L = _lib.lib()
    if L.gcb_device_count() > 0:
        pytest.skip("a GPU is present")
    from gym_chess_b200 import BatchedChessEngine, GcbError
    import numpy as np

    with pytest.raises(GcbError):
        BatchedChessEngine().update_state(np.zeros((1, 64), np.int8), np.ones((1, 4), np.uint8))
    with pytest.raises(GcbError):
        from gym_chess_b200 import BatchedChessEnv
        BatchedChessEnv(4, opponent="none")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gym_chess_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle/", "ORACLE_DIR/") or f.endswith((".cuh", ".cu")), f
                assert "import oracle" not in src and "from oracle" not in src and "libgc_oracle" not in src, f
                assert "host_emul" not in src or f.endswith(".cuh"), f


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """include/gymchess_b200.h is the whole contract: a C99 program that includes it and links libgymchess_b200.so calls
    the library (no torch / C++ / CUDA types in the signatures); without a GPU the compute entry points refuse (GCB_E_NOGPU)."""
    import subprocess

    src = tmp_path / "consumer.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "gymchess_b200.h"
int main(void) {
    gcb_env_config cfg;
    gcb_env *env = NULL;
    memset(&cfg, 0, sizeof cfg);
    cfg.num_envs = 4, cfg.moves_max = -1;
    printf("version %d devices %d\n", gcb_version(), gcb_device_count());
    int rc = gcb_env_create(&cfg, &env);
    printf("create rc %d (%s)\n", rc, rc ? gcb_last_error() : "ok");
    if (rc == GCB_OK) {
        uint64_t st[16];
        rc = gcb_env_step_sampled(env, 3, NULL, NULL, NULL, NULL, NULL, NULL);
        if (rc == GCB_OK) rc = gcb_env_stats(env, st, NULL);
        printf("steps %llu rc %d\n", (unsigned long long)st[0], rc);
        gcb_env_destroy(env);
    }
    return 0;
}
''')
    exe = tmp_path / "consumer"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           _lib.SO_PATH, "-Wl,-rpath," + os.path.dirname(_lib.SO_PATH)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "version 200" in out.stdout
    if _lib.lib().gcb_device_count() > 0:
        assert "steps 12 rc 0" in out.stdout, out.stdout
    else:
        assert "create rc -3" in out.stdout and "no CPU fallback" in out.stdout, out.stdout
