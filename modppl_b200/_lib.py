"""ctypes binding of libmodppl_b200.so (include/modppl_b200.h).

The library is the product: there is no Python or CPU fallback.  If the shared object is missing the import fails
loudly; if no CUDA device is usable every compute call raises `MplError`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmodppl_b200.so")


class MplError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"modppl_b200 error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    # a fresh checkout: compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)
    import subprocess
    try:
        subprocess.check_call(["make", "-C", os.path.dirname(_HERE), "modppl_b200/lib/libmodppl_b200.so"], stdout=subprocess.DEVNULL)
    except Exception as e:  # noqa: BLE001
        raise ImportError(f"could not build {LIB_PATH}: {e}.  modppl_b200 has no CPU fallback.") from e
if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "modppl_b200 has no CPU fallback."
    )
lib = C.CDLL(LIB_PATH)

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)
c_u64_p = C.POINTER(C.c_uint64)


class Move(C.Structure):
    _fields_ = [("kind", C.c_int32), ("proposal", C.c_int32), ("arg", C.c_double), ("mask", C.c_uint32), ("repeat", C.c_uint32)]


class PfConfig(C.Structure):
    _fields_ = [("dtype", C.c_int), ("device", C.c_int), ("seed", C.c_uint64), ("gid_offset", C.c_uint64), ("n_global", C.c_uint64)]


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


# every symbol include/modppl_b200.h declares
SYMBOLS = {
    "mpl_last_error": (C.c_char_p,),
    "mpl_version": (C.c_char_p,),
    "mpl_device_count": (C.c_int, C.POINTER(C.c_int)),
    "mpl_model_create": (C.c_void_p, C.c_char_p, c_double_p, C.c_size_t),
    "mpl_model_compile": (C.c_void_p, C.c_char_p),
    "mpl_model_jit_compile": (C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_size_t),
    "mpl_model_jit_source": (C.c_char_p, C.c_void_p, C.c_int),
    "mpl_model_destroy": (None, C.c_void_p),
    "mpl_model_state_dim": (C.c_int, C.c_void_p),
    "mpl_model_obs_dim": (C.c_int, C.c_void_p),
    "mpl_model_num_latents": (C.c_int, C.c_void_p),
    "mpl_particle_system_new": (C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(PfConfig)),
    "mpl_ps_destroy": (None, C.c_void_p),
    "mpl_ps_init_step": (C.c_int, C.c_void_p, c_double_p, C.c_size_t),
    "mpl_ps_step": (C.c_int, C.c_void_p, c_double_p, C.c_size_t),
    "mpl_ps_effective_sample_size": (C.c_int, C.c_void_p, C.c_int, c_double_p),
    "mpl_ps_resample": (C.c_int, C.c_void_p, C.c_int, c_double_p),
    "mpl_ps_checkpoint_size": (C.c_int, C.c_void_p, C.POINTER(C.c_uint64)),
    "mpl_ps_checkpoint": (C.c_int, C.c_void_p, C.c_void_p, C.c_uint64),
    "mpl_ps_restore": (C.c_int, C.c_void_p, C.c_void_p, C.c_uint64),
    "mpl_ps_step_resample": (C.c_int, C.c_void_p, c_double_p, C.c_size_t, C.c_int, c_double_p),
    "mpl_ps_log_marginal_likelihood_estimate": (C.c_int, C.c_void_p, c_double_p),
    "mpl_ps_read": (C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t),
    "mpl_ps_write": (C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t),
    "mpl_ps_num_particles": (C.c_int, C.c_void_p, c_u64_p),
    "mpl_ps_history_enable": (C.c_int, C.c_void_p, C.c_uint64),
    "mpl_ps_trajectories": (C.c_int, C.c_void_p, c_i64_p, C.c_uint64, c_double_p, C.c_size_t, c_u64_p),
    "mpl_ps_sync": (C.c_int, C.c_void_p),
    "mpl_ps_upload_observations": (C.c_int, C.c_void_p, c_double_p, C.c_size_t, C.c_size_t),
    "mpl_ps_run": (C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_double, c_float_p),
    "mpl_ps_num_resamples": (C.c_int, C.c_void_p, c_u64_p),
    "mpl_ps_profile_enable": (C.c_int, C.c_void_p, C.c_int),
    "mpl_ps_profile_get": (C.c_int, C.c_void_p, C.c_char_p, c_double_p, c_u64_p),
    "mpl_ps_launch_count": (C.c_int, C.c_void_p, c_u64_p),
    "mpl_importance_sampling": (C.c_int, C.c_void_p, c_double_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_uint64, c_double_p, c_double_p, c_double_p),
    "mpl_importance_resampling": (C.c_int, C.c_void_p, c_double_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, c_double_p, c_i64_p, c_double_p),
    "mpl_chains_new": (C.c_void_p, C.c_void_p, c_double_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int),
    "mpl_chains_destroy": (None, C.c_void_p),
    "mpl_model_num_proposals": (C.c_int, C.c_void_p),
    "mpl_model_proposal_name": (C.c_char_p, C.c_void_p, C.c_int),
    "mpl_model_proposal_index": (C.c_int, C.c_void_p, C.c_char_p),
    "mpl_mh": (C.c_int, C.c_void_p, C.c_char_p, C.c_double, C.c_uint32, c_u64_p),
    "mpl_regen_mh": (C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, c_u64_p),
    "mpl_mh_schedule": (C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, c_u64_p, c_float_p),
    "mpl_chains_num_slots": (C.c_int, C.c_void_p),
    "mpl_chains_read": (C.c_int, C.c_void_p, c_double_p, C.c_size_t),
    "mpl_chains_write": (C.c_int, C.c_void_p, c_double_p, C.c_size_t),
    "mpl_resample_indices": (C.c_int, c_double_p, c_double_p, C.c_uint64, C.c_uint64, C.c_int, c_i64_p),
    "mpl_cumsum_sequential": (C.c_int, c_double_p, C.c_uint64, c_double_p),
    "mpl_logsumexp_stats": (C.c_int, C.c_void_p, C.c_uint64, C.c_int, c_double_p, c_double_p, c_double_p),
    "mpl_fixed_resample": (C.c_int, c_float_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint32, c_i32_p, c_double_p, c_u64_p),
    "mpl_logpdf": (C.c_int, C.c_char_p, c_double_p, c_double_p, C.c_size_t, c_double_p),
    "mpl_ps_peer_export": (C.c_int, C.c_void_p, C.c_void_p),
    "mpl_ps_peer_attach": (C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p),
    "mpl_ps_peer_detach": (C.c_int, C.c_void_p),
    "mpl_ps_peer_barrier": (C.c_int, C.c_void_p),
    "mpl_ps_peer_error": (C.c_int, C.c_void_p, C.POINTER(C.c_int)),
    "mpl_ps_trace": (C.c_int, C.c_void_p, C.POINTER(C.c_longlong)),
    "mpl_ps_nvlink_bytes": (C.c_int, C.c_void_p, c_u64_p),
    "mpl_test_set_inline_level1": (C.c_int, C.c_int),
    "mpl_ps_island_export": (C.c_int, C.c_void_p, C.c_void_p),
    "mpl_ps_island_attach": (C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p),
    "mpl_ps_live_buffer": (C.c_int, C.c_void_p, C.POINTER(C.c_int)),
    "mpl_ps_island_copy_from": (C.c_int, C.c_void_p, C.c_int, C.c_int),
    "mpl_ps_copy_state": (C.c_int, C.c_void_p, C.c_void_p),
    "mpl_test_virtual_shards": (C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, c_double_p, C.c_size_t, C.c_size_t, c_double_p, c_double_p, c_double_p, c_double_p),
    "mpl_test_virtual_shards_scheme": (C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_int, c_double_p, C.c_size_t, C.c_size_t, c_double_p, c_double_p, c_double_p, c_double_p),
}
for _name, _s in SYMBOLS.items():
    _sig(_name, _s[0], *_s[1:])


def last_error():
    return lib.mpl_last_error().decode()


def check(rc):
    if rc != 0:
        raise MplError(rc, last_error())


def check_handle(h):
    if not h:
        raise MplError(-1, last_error())
    return h


def device_count():
    n = C.c_int(0)
    rc = lib.mpl_device_count(C.byref(n))
    return n.value if rc == 0 else 0
