"""ctypes binding of libcar_b200.so (include/car_b200.h).

The library is the product's only arithmetic path; there is no CPU or eager
fallback.  ``load()`` raises ``RuntimeError`` if the shared object is missing
(run ``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C
cross_attention_renderer_b200/csrc``).
"""
import ctypes as C
import os

ABI_VERSION = 7
PREC_FP32_SIMT, PREC_FP32_3XBF16, PREC_BF16 = 0, 1, 2
PRECISIONS = {"fp32_simt": PREC_FP32_SIMT, "fp32": PREC_FP32_3XBF16, "bf16": PREC_BF16}
K_ENC = 592
GEOM_STRIDE = 32

STAGES = ("raysetup", "sample_geom", "gather", "gemm_enc1", "gemm_enc2", "gemm_kv", "gemm_small",
          "attention", "phi", "pack", "fused", "backward",
          "bwd_dgrad", "bwd_wgrad", "bwd_operand_copies", "bwd_scatter")   # backward = attention / colour-MLP / per-ray part

_LIB_PATH = os.environ.get("CAR_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcar_b200.so")

c_fp = C.c_void_p


class car_mat(C.Structure):
    _fields_ = [("f32", c_fp), ("hi", c_fp), ("lo", c_fp), ("bias", c_fp),
                ("N", C.c_int32), ("K", C.c_int32)]


class car_weights(C.Structure):
    _fields_ = [("enc1", car_mat), ("enc2", car_mat), ("value", car_mat), ("key1", car_mat),
                ("key2", car_mat), ("qry1", car_mat), ("qry2", car_mat), ("rep1_loc", car_mat),
                ("rep1_g", car_mat), ("rep2", car_mat), ("enc_lat", car_mat), ("phi_in", car_mat),
                ("phi_z", car_mat * 3), ("phi_fc0", car_mat * 3), ("phi_fc1", car_mat * 3),
                ("phi_out", car_mat), ("kv_fold", car_mat), ("kv_fold64", car_mat), ("rowb_fold", car_mat), ("phi_pack", car_mat)]


class car_cameras(C.Structure):
    _fields_ = [("Q", c_fp), ("Cself", c_fp), ("Rel", c_fp), ("qinv", c_fp), ("K", c_fp),
                ("Kq", c_fp)]


class car_debug(C.Structure):
    _fields_ = [("geom", c_fp), ("x", c_fp), ("interp", c_fp), ("value", c_fp), ("key", c_fp),
                ("q1", c_fp), ("q2", c_fp), ("zfinal", c_fp)]


class car_render_args(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("precision", C.c_int32),
                ("b", C.c_int32), ("R", C.c_int32), ("P", C.c_int32), ("H", C.c_int32),
                ("W", C.c_int32), ("ray_begin", C.c_int32), ("ray_end", C.c_int32),
                ("feat_bf16", C.c_int32), ("feat", c_fp * 3),
                ("weights", car_weights), ("cams", car_cameras),
                ("uv", c_fp), ("interval", c_fp),
                ("rgb", c_fp), ("valid_mask", c_fp), ("depth_ray", c_fp), ("at_wt", c_fp),
                ("at_wt_max", c_fp), ("pixel_val", c_fp), ("coords", c_fp),
                ("workspace", c_fp), ("workspace_bytes", C.c_size_t),
                ("debug", car_debug), ("stream", c_fp), ("use_fused", C.c_int32),
                ("train", C.c_int32), ("chunk_rays", C.c_int32)]


class car_general_weights(C.Structure):
    _fields_ = [("enc1", car_mat), ("enc2", car_mat), ("merge", car_mat), ("value", car_mat), ("key1", car_mat),
                ("key2", car_mat), ("qry1", car_mat), ("qry2", car_mat), ("rep1_loc", car_mat), ("rep1_g", car_mat),
                ("rep2", car_mat), ("enc_lat", car_mat), ("phi_in", car_mat), ("phi_z", car_mat * 3),
                ("phi_fc0", car_mat * 3), ("phi_fc1", car_mat * 3), ("phi_out", car_mat)]


class car_general_args(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_view", C.c_int32), ("flags", C.c_int32),
                ("b", C.c_int32), ("R", C.c_int32), ("P", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("ray_begin", C.c_int32), ("ray_end", C.c_int32), ("feat", c_fp * 3),
                ("weights", car_general_weights), ("cams", car_cameras), ("uv", c_fp), ("interval", c_fp),
                ("rgb", c_fp), ("valid_mask", c_fp), ("depth_ray", c_fp), ("at_wt", c_fp), ("at_wt_max", c_fp),
                ("pixel_val", c_fp), ("coords", c_fp), ("workspace", c_fp), ("workspace_bytes", C.c_size_t),
                ("stream", c_fp), ("chunk_rays", C.c_int32), ("debug_interp", c_fp), ("debug_zfinal", c_fp),
                ("precision", C.c_int32)]


FLAG_NO_SAMPLE, FLAG_NO_LATENT_CONCAT = 1, 2


class car_mat_grad(C.Structure):
    _fields_ = [("w", c_fp), ("bias", c_fp)]


GRAD_MATS = ("enc1", "enc2", "value", "key1", "key2", "qry1", "qry2", "rep1_loc", "rep1_g", "rep2",
             "enc_lat", "phi_in")


class car_weight_grads(C.Structure):
    _fields_ = [(n, car_mat_grad) for n in GRAD_MATS] + [
        ("phi_z", car_mat_grad * 3), ("phi_fc0", car_mat_grad * 3), ("phi_fc1", car_mat_grad * 3),
        ("phi_out", car_mat_grad)]


class car_backward_args(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("fwd", C.POINTER(car_render_args)),
                ("d_rgb", c_fp), ("d_depth_ray", c_fp), ("grads", car_weight_grads),
                ("d_feat", c_fp * 3), ("workspace", c_fp), ("workspace_bytes", C.c_size_t),
                ("stream", c_fp), ("precision", C.c_int32)]


# every symbol include/car_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "car_version": (C.c_int, []),
    "car_last_error": (C.c_char_p, []),
    "car_features_bytes": (C.c_size_t, [C.c_int] * 5),
    "car_pack_features": (C.c_int, [c_fp, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp]),
    "car_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "car_default_chunk_rays": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "car_render_forward": (C.c_int, [C.POINTER(car_render_args)]),
    "car_train_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "car_backward_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "car_render_backward": (C.c_int, [C.POINTER(car_backward_args)]),
    "car_unpack_features": (C.c_int, [c_fp, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, c_fp]),
    "car_general_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "car_general_default_chunk_rays": (C.c_int, [C.c_int] * 3),
    "car_render_forward_general": (C.c_int, [C.POINTER(car_general_args)]),
    "car_adam_step": (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_int, c_fp, c_fp]),
    "car_last_launch_count": (C.c_int, []),
    "car_profile_begin": (C.c_int, []),
    "car_profile_end": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
    "car_debug_set_fused_stats": (C.c_int, [c_fp]),
    "car_gemm_umma_test": (C.c_int, [c_fp] * 6 + [C.c_int] * 5 + [c_fp]),
}
# include/car_b200_test.h (libcar_b200_test.so: micro-benchmarks / building-block kernels, not the product)
TEST_SYMBOLS = {
    "car_mma_rate_test": (C.c_int, [C.c_int] * 7 + [c_fp, c_fp]),
    "car_gemm_pair_test": (C.c_int, [c_fp] * 7 + [C.c_int] * 8 + [c_fp]),
    "car_tap_fetch_ab": (C.c_int, [c_fp, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_int] + [C.c_int] * 5 + [c_fp, c_fp, C.c_int]),
}

_lib = None
_test_lib = None


def lib_path():
    return _LIB_PATH


def build_id():
    """Short hash of the kernel sources the library is built from: ncu summaries under profiles/ carry it,
    and bench.py only quotes a profile's DRAM / L2 traffic for the build it is timing."""
    import hashlib
    csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
    h = hashlib.sha256()
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh")) or name == "Makefile":
            h.update(name.encode())
            h.update(open(os.path.join(csrc, name), "rb").read())
    return h.hexdigest()[:12]


def load():
    """dlopen the library (once) and declare prototypes.  Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} not found: the CUDA extension is not built and there is no CPU "
            "fallback. Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)           # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.car_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libcar_b200.so ABI {got} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def load_test():
    """dlopen libcar_b200_test.so (test / micro-benchmark kernels); the product library is loaded first."""
    global _test_lib
    if _test_lib is not None:
        return _test_lib
    load()
    path = os.path.join(os.path.dirname(_LIB_PATH), "libcar_b200_test.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build with `make -C cross_attention_renderer_b200/csrc`")
    lib = C.CDLL(path)
    for name, (res, args) in TEST_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.car_last_error = _lib.car_last_error
    _test_lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().car_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")
