"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/ishara_b200.h
declares, and its host-side logic (parameter table, argument checks, error reporting) behaves without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ishara_b200 as ib
from ishara_b200 import _lib
from oracle import ishara_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ishara_b200.h")).read()
    return sorted(set(re.findall(r"ISHARA_API[^;(]*?\b(ishara_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ishara_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in ishara_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_no_torch_types_in_header():
    src = open(os.path.join(ROOT, "include", "ishara_b200.h")).read()
    assert "#include <torch" not in src and "at::" not in src and "torch::" not in src and "#include <cuda" not in src


def test_version_and_device_count():
    lib = _lib.load()
    assert b"sm_100a" in lib.ishara_version()
    assert lib.ishara_device_count() >= 0


@pytest.mark.parametrize("kw", [
    {},
    dict(dim=384, num_conv_squeeze_blocks=4, num_conv_conform_blocks=4, input_shape=(1024, 276)),
    dict(dim=128, num_conv_squeeze_blocks=1, num_conv_conform_blocks=0, kernel_sizes=(5,), num_conv_per_block=2,
         num_heads=4, expansion_factor=4, transformer_kernel_size=7, input_shape=(96, 64), num_classes=28),
])
def test_param_table_matches_oracle(kw):
    m = ib.get_model(**kw)
    okw = dict(kw)
    if "input_shape" in okw:
        okw["frames"], okw["features"] = okw.pop("input_shape")
    if "kernel_sizes" in okw:
        okw["kernel_sizes"] = tuple(okw["kernel_sizes"])
    cfg = O.Config(**okw)
    assert m.param_specs == O.param_specs(cfg)
    assert m.count_params() == O.count_params(cfg)[0]


def test_get_model_signature_matches_reference():
    import inspect

    sig = inspect.signature(ib.get_model)
    names = list(sig.parameters)
    # c7:1-11, in order, with the reference's defaults
    assert names[:9] == ["dim", "num_conv_squeeze_blocks", "num_conv_conform_blocks", "kernel_sizes",
                         "num_conv_per_block", "dropout_rate", "num_heads", "expansion_factor",
                         "transformer_kernel_size"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["dim"], d["num_conv_squeeze_blocks"], d["num_conv_conform_blocks"], tuple(d["kernel_sizes"]),
            d["num_conv_per_block"], d["dropout_rate"], d["num_heads"], d["expansion_factor"],
            d["transformer_kernel_size"]) == (256, 2, 2, (11, 5, 3), 3, 0.2, 8, 2, 15)


def test_weights_round_trip_and_keras_default_init(tmp_path):
    m = ib.get_model(dim=64, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1, num_heads=4, input_shape=(32, 20),
                     num_classes=12)
    w = m.get_weights()
    assert np.all(w["stem_bn.gamma"] == 1) and np.all(w["stem_bn.moving_variance"] == 1)
    assert np.all(w["top_conv.bias"] == 0) and np.all(w["stem_bn.moving_mean"] == 0)
    lim = np.sqrt(6.0 / (20 + 64))
    assert np.abs(w["stem_conv.kernel"]).max() <= lim and w["stem_conv.kernel"].std() > 0.3 * lim
    p = O.init_params(O.Config(dim=64, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1, num_heads=4, frames=32,
                               features=20, num_classes=12))
    m.load_weights(p)
    path = tmp_path / "w.npz"
    m.save_weights(path)
    m2 = ib.get_model(dim=64, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1, num_heads=4, input_shape=(32, 20),
                      num_classes=12).load_weights(str(path))
    for k, v in m2.get_weights().items():
        assert np.array_equal(v, p[k]), k
    # same through the safetensors container (readable by the `safetensors` package: u64 header length + JSON + raw fp32)
    sp = tmp_path / "w.safetensors"
    m.save_weights(sp)
    raw = open(sp, "rb").read()
    n = int.from_bytes(raw[:8], "little")
    import json
    hdr = json.loads(raw[8:8 + n])
    assert hdr["stem_conv.kernel"]["dtype"] == "F32" and hdr["stem_conv.kernel"]["shape"] == [20, 64] and n % 8 == 0
    m3 = ib.get_model(dim=64, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1, num_heads=4, input_shape=(32, 20),
                      num_classes=12).load_weights(sp)
    for k, v in m3.get_weights().items():
        assert np.array_equal(v, p[k]), k


def test_error_reporting_without_gpu():
    lib = _lib.load()
    m = ib.get_model(dim=64, num_conv_squeeze_blocks=0, num_conv_conform_blocks=1, num_heads=4, input_shape=(32, 20),
                     num_classes=12)
    with pytest.raises(KeyError):
        m.load_weights({"nope.kernel": np.zeros(3, np.float32)})
    with pytest.raises(ValueError):
        m.load_weights({"stem_conv.kernel": np.zeros((3, 3), np.float32)})
    a = np.zeros(5, np.float32)
    st = lib.ishara_model_set_param(m._h, b"nope", a.ctypes.data_as(C.c_void_p), 5)
    assert st == _lib.ERR_INVALID and b"unknown parameter" in lib.ishara_last_error()
    st = lib.ishara_model_set_param(m._h, b"stem_conv.kernel", a.ctypes.data_as(C.c_void_p), 5)
    assert st == _lib.ERR_SHAPE
    assert lib.ishara_model_forward(None, None, 1, None, None) == _lib.ERR_INVALID
    # forward before finalize (finalize itself needs the GPU)
    st = lib.ishara_model_forward(m._h, C.c_void_p(16), 1, C.c_void_p(16), None)
    assert st == _lib.ERR_STATE
    with pytest.raises(ValueError):
        m(np.zeros((1, 31, 20), np.float32))
    with pytest.raises(ValueError):
        ib.get_model(mask_mode="something-else")
    mp = ib.get_model(dim=64, num_conv_squeeze_blocks=0, num_conv_conform_blocks=1, num_heads=4, input_shape=(32, 20),
                      num_classes=12, mask_mode="propagated")       # accepted; the mode is a host-side switch until forward
    assert lib.ishara_model_set_mask_mode(mp._h, 7) == _lib.ERR_INVALID
    assert lib.ishara_model_comm_init(mp._h, None, 0, 2) == _lib.ERR_INVALID   # null unique id
    assert lib.ishara_model_train_counters(mp._h, None, None, None) == _lib.ERR_STATE  # no training state yet
    mp.close()
    bad = _lib.Config()
    h = C.c_void_p()
    assert lib.ishara_model_create(C.byref(bad), 0, C.byref(h)) == _lib.ERR_SHAPE


def test_no_cpu_fallback_compute_fails_loudly_without_gpu():
    lib = _lib.load()
    if lib.ishara_device_count() > 0:
        pytest.skip("GPU present")
    m = ib.get_model(dim=64, num_conv_squeeze_blocks=0, num_conv_conform_blocks=0, num_heads=4, input_shape=(32, 20),
                     num_classes=12)
    with pytest.raises(ib.IsharaError) as e:
        m(np.zeros((1, 32, 20), np.float32))
    assert e.value.status == _lib.ERR_CUDA


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ishara_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_host_char_map_and_postprocess_match_oracle():
    assert ib.char_to_num == O.CHAR_TO_NUM and ib.num_to_char == O.NUM_TO_CHAR
    assert "".join(ib.num_to_char_fn(ib.FALLBACK_IDS)) == "2 a-e -aroe"
    for ids in ([1, 2], list(range(13)), []):
        assert np.array_equal(ib.tflite_postprocess(ids), O.tflite_postprocess(np.asarray(ids)))


def test_dlpack_view_of_numpy_is_rejected_as_device_operand():
    from ishara_b200 import _dlpack

    v = _dlpack.view(np.zeros((2, 3), np.float32))
    assert v.shape == (2, 3) and v.dtype == "float32" and not v.on_cuda
    with pytest.raises(ValueError):
        _dlpack.view(np.zeros((4, 4), np.float32)[:, ::2])


def test_edit_distances_host_function():
    """ishara_edit_distances is a host function: known answers + a brute-force DP (c18:1-15 scorer)."""
    import random

    from ishara_b200 import edit_distances, levenshtein_scores

    def dp(a, b):
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
            prev = cur
        return prev[-1]

    assert edit_distances(["kitten", "", "abc", "flaw"], ["sitting", "abc", "", "lawn"]).tolist() == [3, 3, 3, 2]
    rnd = random.Random(0)
    pairs = [("".join(rnd.choice("abc -") for _ in range(rnd.randint(0, 20))), "".join(rnd.choice("abc -") for _ in range(rnd.randint(1, 20))))
             for _ in range(200)]
    got = edit_distances([p for p, _ in pairs], [t for _, t in pairs])
    assert got.tolist() == [dp(p, t) for p, t in pairs]
    s = levenshtein_scores(["3 creekhouse", "x"], ["3 creekhouse", "ab"])
    assert s[0] == 1.0 and s[1] == 0.0


def test_new_entry_points_report_errors_without_gpu():
    """Training, preprocessing, pipelined inference and scorer entry points: argument and state errors are status codes
    with a message; no compute is attempted without a device and nothing falls back to the CPU."""
    lib = _lib.load()
    m = ib.get_model(dim=128, num_conv_squeeze_blocks=0, num_conv_conform_blocks=1, num_heads=4, input_shape=(32, 20),
                     num_classes=12)
    # state errors: nothing submitted / no training state yet
    assert lib.ishara_model_infer_collect(m._h) == _lib.ERR_STATE and b"nothing in flight" in lib.ishara_last_error()
    assert lib.ishara_model_train_apply(m._h, None, 1.0, None) == _lib.ERR_STATE
    ptr, n = C.c_void_p(), C.c_int64()
    assert lib.ishara_model_train_grad_buffer(m._h, C.byref(ptr), C.byref(n)) == _lib.ERR_STATE
    a = np.zeros(4, np.float32)
    assert lib.ishara_model_train_param_grad(m._h, b"stem_conv.kernel", a.ctypes.data_as(C.c_void_p), 4) == _lib.ERR_STATE
    assert lib.ishara_model_train_sync(m._h) == _lib.OK                      # nothing to sync is not an error
    # argument errors
    assert lib.ishara_model_train_configure(None, 0.0, 0, 0) == _lib.ERR_INVALID
    assert lib.ishara_model_train_configure(m._h, 1.5, 0, 0) == _lib.ERR_INVALID and b"dropout" in lib.ishara_last_error()
    assert lib.ishara_model_infer_submit(m._h, None, 1, None, 0, None, None, None, None) == _lib.ERR_INVALID
    assert lib.ishara_model_train_step_host(m._h, None, None, 1, 1, None, None) == _lib.ERR_INVALID
    assert lib.ishara_preprocess(None, None, 1, 0, None, None, 32, 1, None, None) == _lib.ERR_INVALID
    assert lib.ishara_edit_distances(None, None, None, None, 3, None) == _lib.ERR_INVALID
    assert lib.ishara_edit_distances(None, None, None, None, 0, None) == _lib.OK
    if lib.ishara_device_count() == 0:
        # compute needs the device: loud CUDA error, no fallback
        with pytest.raises(ib.IsharaError) as e:
            m.train_step(np.zeros((1, 32, 20), np.float32), np.zeros((1, 4), np.int32))
        assert e.value.status == _lib.ERR_CUDA
        with pytest.raises(ib.IsharaError):
            ib.LandmarkPreprocessor({g: (0.0, 1.0) for g in ("lip", "rhand", "lhand", "rpose", "lpose")}, frame_len=32)
    m.close()


def test_c_host_example_builds_and_uses_the_abi(tmp_path):
    """examples/infer.c: a plain-C host (no Python, no torch) drives the ABI; 235 tensors / 7,591,096 parameters are the
    reference's model.summary() count for the BASELINE configuration. Without a GPU the first compute call fails loudly."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = tmp_path / "infer"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "infer.c"),
                    "-L", libdir, "-lishara_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert "235 tensors, 7591096 parameters" in r.stdout
    if _lib.load().ishara_device_count() == 0:
        assert r.returncode == 3 and "status 3" in r.stderr
    else:
        assert r.returncode == 0 and "sequence 1:" in r.stdout
