"""CPU: the C-ABI library builds/loads and exports every symbol include/simseg_b200.h declares (no compute calls —
there is no GPU here), the host mirror keeps the reference's names / state-dict keys / error behaviour, and the
product path fails LOUDLY without a CUDA device (there is no CPU fallback to route through).
"""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "simseg_b200.h")


@pytest.fixture(scope="module")
def lib():
    from simseg_b200 import _lib, build
    build.build()                      # no-op when libsimseg_b200.so is newer than its sources
    return _lib.load()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(simseg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    names = _declared_functions()
    assert len(names) >= 25
    raw = ctypes.CDLL(os.path.join(ROOT, "simseg_b200", "lib", "libsimseg_b200.so"))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/simseg_b200.h but not exported"


def test_ctypes_prototypes_cover_header(lib):
    from simseg_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared_functions()


def test_gemm_args_struct_matches_header():
    """Field order of the ctypes struct == field order of simseg_gemm_args in the header."""
    from simseg_b200._lib import GemmArgs
    src = open(HEADER).read()
    body = re.search(r"typedef struct simseg_gemm_args \{(.*?)\} simseg_gemm_args;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if " " in decl else decl
        names = re.sub(r"^(const\s+)?(void|float|int64_t|int32_t)\s*\*?", "", decl).strip()
        fields += [n.strip().lstrip("*").strip() for n in names.split(",")]
    assert fields == [f[0] for f in GemmArgs._fields_]


def test_no_gpu_calls_fail_loudly(lib):
    from simseg_b200 import _lib, ops
    assert lib.simseg_version() >= 100
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = ctypes.c_void_p()
    rc = lib.simseg_ctx_create(0, ctypes.byref(p))
    assert rc < 0 and lib.simseg_last_error()          # an error code + message, not a crash and not a CPU path
    with pytest.raises(_lib.SimsegError):
        ops.ctx()
    with pytest.raises(_lib.SimsegError):
        ops.patch_text_sim(torch.zeros(4, 512), torch.zeros(3, 512))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "simseg_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "/root/reference" not in txt, f


# ------------------------------------------------------------------------------------------ host mirror
def _cfg(extra=()):
    from simseg_b200.config import load_cfg
    return load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                                          "transforms.input_size=224"] + list(extra))


def test_config_surface():
    from simseg_b200.config import load_cfg
    cfg = _cfg()
    assert cfg.model.pool.name == "loda" and cfg.model.pool.loda.image_k == 5 and cfg.model.pool.loda.text_k == 1
    assert cfg.model.projection.dim == 512 and cfg.loss.name == "NCE" and cfg.loss.global_reduce is True
    assert cfg.loss.temperature.name == "parameter" and abs(cfg.loss.temperature.value - 0.02) < 1e-12
    assert isinstance(cfg.optim.param.betas, tuple)
    b = load_cfg("simseg.vit-b.yaml")
    assert b.model.image_encoder.embedding_dim == 768
    with pytest.raises(KeyError):                       # core/config.py:194-195: unknown keys are rejected
        _cfg(["model.no_such_key=1"])


def test_state_dict_keys_match_reference_naming_cpu():
    from oracle import simseg_oracle as O
    from simseg_b200.pipeline import PIPELINE
    model = PIPELINE["clip"](_cfg())
    sd = O.make_state_dict(384, 6)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    # attribute paths other reference code reaches into
    vit = model.image_encoder.model.model
    assert vit.patch_embed.num_patches == 196 and tuple(vit.pos_embed.shape) == (1, 197, 384)
    assert model.loss.temperature.shape == () and model.cfg is not None
    for m in ("forward_image_feature", "forward_image_project", "image_projection", "forward_text_feature",
              "forward_text_project", "forward_loss"):
        assert hasattr(model, m)


def test_unknown_options_raise_like_the_reference():
    from simseg_b200.pipeline import PIPELINE
    for ov in ("model.projection.name=complex", "model.pool.name=avg", "loss.temperature.name=cosine"):
        with pytest.raises(NotImplementedError):
            PIPELINE["clip"](_cfg([ov]))


def test_model_forward_without_gpu_raises():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oracle import simseg_oracle as O
    from simseg_b200._lib import SimsegError
    from simseg_b200.pipeline import PIPELINE
    model = PIPELINE["clip"](_cfg())
    with pytest.raises((SimsegError, AssertionError, RuntimeError)):
        model(O.make_batch(2, 25))


def test_library_is_sm100a_tcgen05_and_tma_code(lib):
    """The shipped library holds sm_100a cubins only, and the hot kernels really are tcgen05 / TMEM / TMA code
    (SASS mnemonics per B200_PROFILING.md): a rebuild that silently lost them would still pass a parity test."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    so = os.path.join(ROOT, "simseg_b200", "lib", "libsimseg_b200.so")
    elfs = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True, timeout=120).stdout
    archs = set(re.findall(r"\.(sm_[0-9a-z]+)\.cubin", elfs))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, timeout=300).stdout
    for mnemonic, least in (("UTCHMMA", 100), ("UTMALDG", 100), ("UTMASTG", 20), ("LDTM", 50), ("UTCBAR", 20), ("UBLKCP", 1)):
        assert sass.count(mnemonic) >= least, (mnemonic, sass.count(mnemonic))
    assert "HMMA.16816" in sass            # the mma.sync attention kernels kept as the tested alternate (short sequences)
