"""ctypes binding of libnerfw_sm100.so (include/nerfw.h).  No fallback: a missing library is an ImportError."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# NERFW_PROFILE_LIB=1 (scripts/ only): the separate profiling build (`make -C csrc PROFILE=1`), whose kernels honour the
# timeline / skip switches.  The product library ignores the environment.
LIB_PATH = os.path.join(_HERE, "libnerfw_sm100_profile.so" if os.environ.get("NERFW_PROFILE_LIB") == "1" else "libnerfw_sm100.so")
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")

N_LAYERS = 8
MLP_FP32, MLP_BF16X3, MLP_BF16, MLP_FP16 = 0, 1, 2, 3
MODE_NAMES = {"fp32": MLP_FP32, "bf16x3": MLP_BF16X3, "bf16": MLP_BF16, "fp16": MLP_FP16}

c_float_p = C.c_void_p  # device pointers are passed as integers


class NerfwWeights(C.Structure):
    _fields_ = [
        ("pts_w", C.c_void_p * N_LAYERS),
        ("pts_b", C.c_void_p * N_LAYERS),
        ("density_w", C.c_void_p),
        ("density_b", C.c_void_p),
        ("dir_w", C.c_void_p),
        ("dir_b", C.c_void_p),
        ("app_w", C.c_void_p),
        ("app_b", C.c_void_p),
        ("rgb_w", C.c_void_p),
        ("rgb_b", C.c_void_p),
    ]


NerfwGrads = NerfwWeights  # same layout (include/nerfw.h)


class NerfwRenderOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("rgb", "depth", "acc", "weights", "z_vals", "rgb_coarse", "depth_coarse",
                                          "acc_coarse", "weights_coarse", "z_coarse")]


# name -> (restype, argtypes); mirrors include/nerfw.h one to one
SIGNATURES = {
    "nerfw_last_error": (C.c_char_p, []),
    "nerfw_abi_version": (C.c_int, []),
    "nerfw_check_device": (C.c_int, []),
    "nerfw_launch_count": (C.c_uint64, []),
    "nerfw_raygen": (C.c_int, [C.c_int, C.c_int, C.c_float, C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p]),
    "nerfw_normalize_dirs": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nerfw_stratified": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "nerfw_ray_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "nerfw_sample_pdf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nerfw_sample_pdf_general": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nerfw_posenc": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nerfw_packed_bytes": (C.c_size_t, []),
    "nerfw_pack_weights": (C.c_int, [C.POINTER(NerfwWeights), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nerfw_mlp_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "nerfw_mlp_mask_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "nerfw_mlp_fwd": (C.c_int, [C.POINTER(NerfwWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                 C.c_void_p]),
    "nerfw_mlp_bwd_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int64]),
    "nerfw_mlp_bwd": (C.c_int, [C.POINTER(NerfwWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_int64, C.c_int, C.c_void_p, C.POINTER(NerfwGrads), C.c_void_p, C.c_void_p,
                                 C.c_size_t, C.c_void_p]),
    "nerfw_mlp_bwd_tc_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "nerfw_mlp_bwd_tc": (C.c_int, [C.POINTER(NerfwWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(NerfwGrads),
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nerfw_composite_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "nerfw_composite_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "nerfw_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                              C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "nerfw_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nerfw_quantize_u8": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nerfw_merge_raw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nerfw_unmerge_raw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "nerfw_volume_render_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int, C.c_int64]),
    "nerfw_volume_render": (C.c_int, [C.POINTER(NerfwWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                       C.c_int, C.POINTER(NerfwRenderOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nerfw_max_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nerfw_fog": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                            C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    "nerfw_depth_edges": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "nerfw_toon": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                             C.c_void_p]),
    "nerfw_hologram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p]),
    "nerfw_selftest_umma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nerfw_selftest_umma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nerfw_selftest_umma_mn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}


def build(verbose: bool = False, profile: bool = False) -> str:
    """Compile the CUDA sources for sm_100a (nvcc cross-compiles without a GPU) and return the .so path.
    profile=True builds the separate profiling library (`make PROFILE=1`) that scripts/ load with NERFW_PROFILE_LIB=1."""
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))] + (["PROFILE=1"] if profile else [])
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("building libnerfw_sm100%s.so failed (see output above)" % ("_profile" if profile else ""))
    return os.path.join(_HERE, "libnerfw_sm100_profile.so") if profile else os.path.join(_HERE, "libnerfw_sm100.so")


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension is not built.  Run `python __graft_entry__.py` "
                "(or `make -C depth-aware-shader-effects-for-nerf_b200/csrc`).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.nerfw_abi_version() != 1:
            raise ImportError("libnerfw_sm100.so ABI version mismatch; rebuild it")
        _lib = handle
    return _lib


class NerfwError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().nerfw_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise NerfwError(f"[{rc}] {msg}")
