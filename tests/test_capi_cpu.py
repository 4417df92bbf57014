"""The C-ABI library without a GPU: it loads, exports every symbol include/bsreg_cuda.h
declares, the ctypes mirror has the layout a C compiler gives the header's structs, and
the product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "bsreg_cuda.h"


def declared_functions():
    src = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|int64_t|char)\s*\*?\s*(bsreg_\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("bsreg_create", "bsreg_run", "bsreg_destroy", "bsreg_get_state", "bsreg_set_state",
                 "bsreg_eval_spmm", "bsreg_eval_gram", "bsreg_eval_logdet", "bsreg_eval_conditionals",
                 "bsreg_last_error", "bsreg_abi_version"):
        assert must in names
    # every entry point cites the reference interface it replaces
    text = HEADER.read_text()
    for cite in ("R/50_execute.R", "R/40_setup.R", "R/10_lm.R", "R/15_sar.R", "R/15_sem.R", "R/20_mh.R",
                 "R/41_set_options.R", "R/42_set_mh.R", "R/60_interface.R"):
        assert cite in text


def test_library_exports_every_declared_symbol(capi):
    lib = capi.load()
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in bsreg_cuda.h but not exported"
        assert name in capi.SYMBOLS, f"{name} has no ctypes prototype"
    assert sorted(capi.SYMBOLS) == names
    assert lib.bsreg_abi_version() == 2


def test_ctypes_structs_match_the_c_layout(capi, tmp_path):
    """Compile a probe with gcc against the real header and compare sizeof/offsetof."""
    structs = {"bsreg_problem": capi.Problem, "bsreg_options": capi.Options, "bsreg_mh": capi.MH}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == C.sizeof(ct)
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback(capi):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((50, 2))
    y = X @ [1.0, 2.0] + rng.standard_normal(50)
    with pytest.raises(capi.BsregError) as e:
        capi.Handle(y, X, model=capi.MODEL_LM)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)          # BSREG_ENOGPU
    with pytest.raises(capi.BsregError):
        capi.eval_rng(0, 1, 0, 0, 1, 4)


def test_argument_validation_happens_before_any_device_work(capi):
    lib = capi.load()
    out = C.c_void_p()
    assert lib.bsreg_create(None, None, C.byref(out)) == -1                  # BSREG_EINVAL
    assert b"null" in lib.bsreg_last_error()
    prob, opt = capi.Problem(), capi.Options()
    assert lib.bsreg_create(C.byref(prob), C.byref(opt), C.byref(out)) == -1
    assert lib.bsreg_run(None, 1, 1, None, None, None, None) == -1
    assert lib.bsreg_dims(None, None, None, None, None, None) == -1
    y = np.zeros(4); X = np.ones((4, 1), order="F")
    prob.n, prob.k = 4, 1
    prob.y, prob.X = y.ctypes.data_as(capi.c_dp), X.ctypes.data_as(capi.c_dp)
    opt.model, opt.n_chains = capi.MODEL_SAR, 1                              # SAR without W
    assert lib.bsreg_create(C.byref(prob), C.byref(opt), C.byref(out)) == -1
    assert b"connectivity" in lib.bsreg_last_error()                         # message of R/15_sar.R:52
    lib.bsreg_destroy(None)                                                  # no-op, must not crash


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under bsreg_b200/ may import or execute it."""
    for path in (ROOT / "bsreg_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
    code = "import sys; import bsreg_b200, bsreg_b200.capi; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def _order(capi, W, force=0):
    n = W.shape[0]
    perm = np.full(n, -1, dtype=np.int32)
    rp, ci = W.indptr.astype(np.int32), W.indices.astype(np.int32)
    rc = capi.load().bsreg_debug_internal_order(n, rp.ctypes.data_as(capi.c_ip), ci.ctypes.data_as(capi.c_ip), force,
                                                perm.ctypes.data_as(capi.c_ip))
    return rc, perm


def test_internal_ordering_of_the_observations(capi):
    """Host logic of bsreg_create (no device): W without locality is relabelled breadth-first, local W is kept; the
    relabelling is a permutation that brings the columns of a 64-row tile together."""
    from bsreg_b200.synthetic import knn_weights
    n = 20000
    W_rand, _ = knn_weights(n, 10, 1001, False, False)         # points in random order
    W_sorted, _ = knn_weights(n, 10, 1001, False, True)        # Morton order: local already
    rc, perm = _order(capi, W_rand)
    assert rc == 1
    assert np.array_equal(np.sort(perm), np.arange(n))
    rc_s, _ = _order(capi, W_sorted)
    assert rc_s == 0
    rc_f, perm_f = _order(capi, W_sorted, force=1)
    assert rc_f == 1 and np.array_equal(np.sort(perm_f), np.arange(n))

    def distinct_per_tile(W):
        rp, ci = W.indptr, W.indices
        return np.mean([len(np.unique(ci[rp[t]:rp[min(t + 64, n)]])) for t in range(0, n, 64)])
    before = distinct_per_tile(W_rand)
    Wp = W_rand[perm][:, perm].tocsr()
    after = distinct_per_tile(Wp)
    assert before > 600 and after < 0.45 * before               # 640 references per tile, about 220 distinct after
    # breadth-first queue order: the first observation stays first, its neighbours follow
    assert perm[0] == 0
    first = set(W_rand.indices[W_rand.indptr[0]:W_rand.indptr[1]].tolist()) - {0}
    assert first == set(perm[1:1 + len(first)].tolist())


def test_internal_ordering_handles_disconnected_and_empty_rows(capi):
    import scipy.sparse as sp
    n = 6000
    rng = np.random.default_rng(0)
    rows = np.repeat(np.arange(n), 3)
    cols = rng.integers(0, n, size=3 * n)
    A = sp.csr_matrix((np.ones(3 * n), (rows, cols)), shape=(n, n)).tolil()
    for i in (0, 17, n - 1):
        A.rows[i] = []; A.data[i] = []
    A = sp.csr_matrix(A)
    rc, perm = _order(capi, A, force=1)
    assert rc == 1 and np.array_equal(np.sort(perm), np.arange(n))
