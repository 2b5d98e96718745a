"""CPU tests (no GPU): the C-ABI shared library builds/loads and exports every symbol that
include/grapes_b200.h declares; no compute entry point is called here."""
import ctypes
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so_path():
    from grapes_b200.build import build_library
    return build_library()


def test_header_declares_the_expected_entry_points():
    from grapes_b200._lib import parse_header
    protos = parse_header()
    for name in ("grapes_ctx_create", "grapes_ctx_destroy", "grapes_last_error", "grapes_expand_rows",
                 "grapes_rank_nodes", "grapes_slice_block", "grapes_relabel", "grapes_build_csr", "grapes_aggregate",
                 "grapes_gemm", "grapes_gemm_tn", "grapes_sampler_l1_fwd", "grapes_sampler_l1_bwd",
                 "grapes_select_topk", "grapes_classifier_loss", "grapes_adam_step"):
        assert name in protos, name
    assert len(protos) >= 30


def test_library_exports_every_declared_symbol(so_path):
    from grapes_b200._lib import parse_header
    cdll = ctypes.CDLL(so_path)
    missing = [n for n in parse_header() if not hasattr(cdll, n)]
    assert not missing, missing
    cdll.grapes_abi_version.restype = ctypes.c_int
    assert cdll.grapes_abi_version() == 1


def test_ctx_create_fails_loudly_without_gpu(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from grapes_b200._lib import lib
    L = lib()
    ctx = ctypes.c_void_p()
    rc = L.cdll.grapes_ctx_create(0, 100, 1000, 1 << 20, ctypes.byref(ctx))
    assert rc != 0 and L.last_error()


def test_product_path_has_no_cpu_fallback():
    """The product package must not import the oracle (tests-only) and must refuse CPU tensors."""
    import torch
    import grapes_b200.engine, grapes_b200.gcn, grapes_b200.utils, grapes_b200.graph   # noqa: F401
    import sys
    pkg_dir = os.path.join(ROOT, "grapes_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
    from grapes_b200._lib import GrapesError
    from grapes_b200.utils import sample_neighborhoods_from_probs
    with pytest.raises(GrapesError):
        sample_neighborhoods_from_probs(torch.zeros(4, 1), torch.arange(4), 2)
    if not torch.cuda.is_available():
        from grapes_b200.graph import DeviceGraph
        with pytest.raises(GrapesError):
            DeviceGraph.from_edge_index(torch.zeros(2, 3, dtype=torch.long), 4, device="cpu")


def test_built_library_carries_the_id_of_this_source_tree():
    """The loader refuses (or rebuilds) a binary whose build id is not the sha1 of the sources + header it parses its
    prototypes from: a stale library would mean silently corrupted arguments."""
    from grapes_b200 import build
    from grapes_b200._lib import lib
    L = lib()
    assert not build.is_stale()
    L.cdll.grapes_build_id.restype = __import__("ctypes").c_char_p
    assert L.cdll.grapes_build_id().decode() == build.tree_id() == build.built_id()


def test_flag_word_raises_with_the_reason():
    """GRAPES_OVF_* bits of a step (scal[15] / the overflow word) -> GrapesError naming the cause; 0 is silent."""
    import pytest
    from grapes_b200._lib import GrapesError
    from grapes_b200.engine import GrapesEngine
    GrapesEngine.raise_on_flags(0)
    with pytest.raises(GrapesError, match="nodes>cap_n"):
        GrapesEngine.raise_on_flags(4)
    with pytest.raises(GrapesError, match="exchange failed"):
        GrapesEngine.raise_on_flags(32 | 2)


def test_hot_kernels_are_blackwell_native_in_sass(so_path):
    """B200_PROFILING.md, "What proves a Blackwell-native kernel": the default sampler-head kernels issue tcgen05.mma with an
    operand in tensor memory (UTC*MMA + STTM), read their accumulators with tcgen05.ld (LDTM) and stage tiles by TMA (UTMALDG);
    the aggregation stages rows with cp.async.bulk (UBLKCP); nothing in the library uses the legacy mma.sync path (HMMA)."""
    import shutil
    import sys
    if not (shutil.which("cuobjdump") or os.path.isfile("/usr/local/cuda/bin/cuobjdump")) or not shutil.which("c++filt"):
        pytest.skip("cuobjdump / c++filt not available")
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from sass_evidence import sass_counts
    fams, archs = sass_counts(so_path)
    assert "sm_100a" in archs and archs <= {"sm_100a", "sm_52"}          # sm_52: nvcc's empty device-link stub
    for k in ("k_l1_fwd_ts", "k_l1_bwd_ts"):
        r = fams[k]
        assert r["UTC*MMA"] > 0 and r["LDTM"] > 0 and r["STTM"] > 0 and r["UTMALDG"] > 0 and r["SYNCS"] > 0, (k, dict(r))
    assert fams["k_agg_tma"]["UBLKCP"] > 0 and fams["k_agg_tma"]["SYNCS"] > 0
    assert fams["k_step_tail"]["MEMBAR.*SYS"] > 0                         # peer-memory exchange: system-scope fences
    assert sum(r["HMMA"] for r in fams.values()) == 0
