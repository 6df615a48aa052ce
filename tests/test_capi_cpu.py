"""No-GPU checks of the boundary: the C-ABI library loads, exports every symbol that
include/dppo_b200.h declares, and its host-only entry points (layout / workspace sizing /
argument validation) agree with the oracle's layout.  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import dppo_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dppo_b200 import build
    build.build()
    from dppo_b200 import _lib
    return _lib.load_library()


def _cfg(obs, hidden, act, max_rows=1024):
    from dppo_b200._lib import DppoConfig
    c = DppoConfig()
    c.obs_dim, c.act_dim, c.n_hidden, c.max_rows = obs, act, len(hidden), max_rows
    for i, h in enumerate(hidden):
        c.hidden[i] = h
    c.entcoeff, c.vcoeff, c.clip_param = 0.01, 0.5, 0.2
    return c


def test_header_symbols_all_exported(lib):
    from dppo_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dppo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dppo_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 18
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dppo_[a-z_0-9]+)", nm))
    assert declared <= exported, f"declared but not exported: {declared - exported}"
    assert declared == set(_lib.EXPORTED), "ctypes prototypes out of sync with the header"
    assert lib.dppo_version() >= 100


@pytest.mark.parametrize("obs,hidden,act,n", [(3, (16,), 1, 115), (17, (256, 256), 6, 73741),
                                               (376, (256, 256), 17, 171299), (376, (1024,) * 4, 17, 3570723),
                                               (5, (), 2, 30)])
def test_param_layout_matches_oracle(lib, obs, hidden, act, n):
    from dppo_b200 import ParamLayout
    c = _cfg(obs, hidden, act)
    assert lib.dppo_param_count(C.byref(c)) == n
    assert lib.dppo_grad_buffer_len(C.byref(c)) == n + 8
    lay = O.make_layout(obs, hidden, act)
    offs = (C.c_int64 * 32)()
    k = lib.dppo_param_offsets(C.byref(c), offs, 32)
    assert k == 2 * (len(hidden) + 2)
    assert list(offs[:k]) == [e[2] for e in lay.entries]
    pl = ParamLayout(obs, hidden, act)
    assert pl.n_params == n and pl.offsets() == list(offs[:k])
    assert [e.shape for e in pl.entries] == [e[1] for e in lay.entries]


def test_workspace_and_validation(lib):
    from dppo_b200._lib import last_error
    c = _cfg(376, (256, 256), 17, max_rows=131072)
    nbytes = lib.dppo_workspace_bytes(C.byref(c))
    # at least: 2 pi activations + 2 scratch [B,256] + 3 head tiles [B,35]
    assert nbytes > 131072 * 4 * (4 * 256 + 3 * 35)
    assert nbytes < 4 << 30
    bad = _cfg(0, (16,), 1)
    assert lib.dppo_param_count(C.byref(bad)) == -1 and "obs_dim" in last_error(lib)
    bad = _cfg(3, (16,) * 8, 1)
    bad.n_hidden = 9
    assert lib.dppo_param_count(C.byref(bad)) == -1 and "n_hidden" in last_error(lib)
    # no device here: create must fail loudly, not fall back
    ctx = C.c_void_p()
    import torch
    if not torch.cuda.is_available():
        rc = lib.dppo_create(C.byref(c), None, 0, C.byref(ctx))
        assert rc != 0 and ctx.value is None and last_error(lib)


def test_engine_refuses_to_run_without_cuda():
    import torch
    from dppo_b200 import DppoEngine
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        DppoEngine(3, 1, (16,), 800)


def test_missing_library_is_loud(tmp_path):
    from dppo_b200 import _lib
    with pytest.raises(_lib.LibraryMissing):
        _lib.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under tensorflow-dppo_b200/ may reference it."""
    pkg = os.path.join(ROOT, "tensorflow-dppo_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.replace("# oracle-free", ""), f"{fn} mentions the oracle"


def test_minibatch_perm_header_matches_oracle(tmp_path):
    """csrc/perm.cuh is plain C++ on the host side: compile it with g++ and compare the keyed
    permutation with the oracle's numpy restatement, bit for bit (the device kernels inline the same header)."""
    src = tmp_path / "perm_main.cpp"
    src.write_text('''
#include <cstdio>
#include <cstdlib>
#include "%s/tensorflow-dppo_b200/csrc/perm.cuh"
int main(int argc, char** argv) {
    const unsigned n = strtoul(argv[1], 0, 10);
    const unsigned long long seed = strtoull(argv[2], 0, 10), epoch = strtoull(argv[3], 0, 10);
    const dppo::PermKey k = dppo::make_perm_key(n, seed, epoch);
    for (unsigned i = 0; i < n; ++i) printf("%%u\\n", dppo::perm_index(k, i));
    return 0;
}
''' % ROOT)
    exe = tmp_path / "perm_main"
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), str(src)], check=True)
    for n, seed, epoch in ((1, 0, 0), (2, 5, 1), (3, 5, 1), (7, 1234, 0), (1000, 1234, 3), (4096, 2**63 + 11, 2),
                           (100003, 99, 7)):
        out = subprocess.run([str(exe), str(n), str(seed), str(epoch)], capture_output=True, text=True, check=True).stdout
        got = np.array(out.split(), dtype=np.int64)
        ref = O.minibatch_perm(n, seed, epoch)
        assert np.array_equal(got, ref), (n, seed, epoch)
        assert np.array_equal(np.sort(ref), np.arange(n))


@pytest.mark.parametrize("obs,hidden,ncat", [(4, (16,), 2), (128, (256, 256), 18), (5, (), 3)])
def test_param_layout_categorical_matches_oracle(lib, obs, hidden, ncat):
    """pd_type = DPPO_PD_CATEGORICAL: the policy head is ncat wide (distributions.py:53-54, Model.py:14)."""
    c = _cfg(obs, hidden, ncat)
    c.pd_type = 1
    lay = O.make_layout(obs, hidden, ncat, pd="cat")
    assert lib.dppo_param_count(C.byref(c)) == lay.n_params
    offs = (C.c_int64 * 32)()
    k = lib.dppo_param_offsets(C.byref(c), offs, 32)
    assert list(offs[:k]) == [e[2] for e in lay.entries]
    c.pd_type = 5
    assert lib.dppo_param_count(C.byref(c)) < 0 and b"pd_type" in lib.dppo_last_error()


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/dppo_b200.h must compile as C99 (no C++-isms, no torch types)."""
    hdr = os.path.join(ROOT, "include", "dppo_b200.h")
    subprocess.run(["gcc", "-x", "c", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", hdr], check=True)
    txt = open(hdr).read()
    assert "torch" not in txt.lower() and "at::" not in txt and "std::" not in txt


def test_c_program_links_and_calls_the_abi(lib, tmp_path):
    """A plain C consumer (what a cgo/JNI/FFI binding compiles down to): link libdppo_b200.so, call the host-only entry
    points, check the layout numbers and the error path -- no Python, no torch in the process."""
    from dppo_b200 import _lib
    src = tmp_path / "consumer.c"
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "dppo_b200.h"
int main(void) {
    dppo_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.obs_dim = 376; cfg.act_dim = 17; cfg.n_hidden = 2; cfg.hidden[0] = 256; cfg.hidden[1] = 256;
    cfg.max_rows = 131072; cfg.gemm_mode = DPPO_GEMM_TF32X3; cfg.pd_type = DPPO_PD_DIAG_GAUSSIAN;
    cfg.entcoeff = 0.01f; cfg.vcoeff = 0.5f; cfg.clip_param = 0.2f;
    printf("%d %lld %lld %zu\\n", dppo_version(), (long long)dppo_param_count(&cfg), (long long)dppo_grad_buffer_len(&cfg),
           dppo_workspace_bytes(&cfg));
    cfg.pd_type = 9;
    printf("%lld|%s\\n", (long long)dppo_param_count(&cfg), dppo_last_error());
    return 0;
}
''')
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-ldppo_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    ver, n, glen, ws = out[0].split()
    assert int(ver) >= 100 and int(n) == 171299 and int(glen) == 171307 and int(ws) > 500 * (1 << 20)
    assert out[1].startswith("-1|") and "pd_type" in out[1]
