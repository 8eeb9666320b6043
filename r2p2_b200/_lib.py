"""ctypes loader for the CUDA library (r2p2_b200/csrc/libr2p2_b200.so, C ABI in include/r2p2_b200.h).

There is deliberately no fallback: if the shared library is missing or no B200 is present, every
entry point raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``r2p2_b200.build()``.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libr2p2_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]


class R2P2Error(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    hdr = os.path.join(_HERE, "..", "include", "r2p2_b200.h")
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs + [hdr]):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "r2p2_capi.cu")]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None

_D = C.POINTER(C.c_double)
_SIGS = {
    "r2p2_version": (C.c_int, []),
    "r2p2_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "r2p2_destroy": (C.c_int, [C.c_void_p]),
    "r2p2_last_error": (C.c_char_p, [C.c_void_p]),
    "r2p2_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "r2p2_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "r2p2_set_stage": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "r2p2_set_stage_i32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "r2p2_set_robot": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int32]),
    "r2p2_set_controller": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]),
    "r2p2_set_pid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double]),
    "r2p2_read_pid": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6),
    "r2p2_set_noise": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p, C.c_int32, C.c_int32]),
    "r2p2_read_noise_draws": (C.c_int, [C.c_void_p, C.c_void_p]),
    "r2p2_set_robot_index_offset": (C.c_int, [C.c_void_p, C.c_uint32]),
    "r2p2_population_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "r2p2_write_pose": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 5),
    "r2p2_ga_state_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "r2p2_ga_save": (C.c_int, [C.c_void_p, C.c_void_p]),
    "r2p2_ga_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "r2p2_set_mapping": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "r2p2_upload_genomes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "r2p2_set_genomes_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "r2p2_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double]),
    "r2p2_run": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "r2p2_evaluate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                                C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_void_p]),
    "r2p2_sync": (C.c_int, [C.c_void_p]),
    "r2p2_last_run_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "r2p2_read_fitness": (C.c_int, [C.c_void_p, C.c_void_p]),
    "r2p2_read_state": (C.c_int, [C.c_void_p] + [C.c_void_p] * 8),
    "r2p2_read_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "r2p2_fitness_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "r2p2_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "r2p2_set_trace": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "r2p2_read_trace": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 6),
    "r2p2_read_trace_collisions": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "r2p2_state_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "r2p2_save_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "r2p2_load_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "r2p2_cost_grid": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "r2p2_set_goal": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double]),
    "r2p2_read_goal": (C.c_int, [C.c_void_p, C.c_void_p]),
    "r2p2_sense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "r2p2_move": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double]),
    "r2p2_ga_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_double, C.c_double]),
    "r2p2_ga_set_operators": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32]),
    "r2p2_ga_population_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "r2p2_ga_fitness_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "r2p2_ga_write_fitness": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "r2p2_ga_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "r2p2_ga_evaluate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "r2p2_ga_breed": (C.c_int, [C.c_void_p]),
    "r2p2_ga_best": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p]),
    "r2p2_ga_generation": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "r2p2_cast_rays": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 9),
    "r2p2_cast_rays_plain": (C.c_int, [C.c_void_p, C.c_int32] + [C.c_void_p] * 9),
}
EXPORTS = tuple(_SIGS)


def lib():
    """The loaded library; raises R2P2Error when it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise R2P2Error("CUDA library %s is missing: run r2p2_b200.build() (needs nvcc). "
                            "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(handle, rc):
    if rc != 0:
        msg = lib().r2p2_last_error(handle)
        raise R2P2Error("r2p2_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
