"""ctypes loader for libgams_cuda.so.  Fails loudly; never falls back to a CPU path."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_ROOT = _PKG.parent
_LIB = None

ERR = {0: "OK", -1: "EINVAL", -2: "ECUDA", -3: "ENOMEM", -4: "ELAG", -5: "ENODEV", -6: "EPARSE", -7: "ECAP"}


class GamsCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gams_cuda error {code} ({ERR.get(code, '?')}): {msg}")
        self.code = code


def lib_path() -> Path:
    """The in-tree library; GAMS_CUDA_LIB selects another build of it (kernel-variant experiments)."""
    import os
    return Path(os.environ.get("GAMS_CUDA_LIB") or _PKG / "libgams_cuda.so")


def header_symbols() -> list[str]:
    """Every function include/gams_cuda.h declares (used by the symbol-coverage test)."""
    text = (_ROOT / "include" / "gams_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gams_cuda_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.exists():
        raise ImportError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C gams_b200/csrc`). gams_b200 has no CPU fallback.")
    L = C.CDLL(str(p))
    vp, i32, i64, u64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
    sig = {
        "gams_cuda_version": (i32, []),
        "gams_cuda_last_error": (C.c_char_p, []),
        "gams_cuda_device_count": (i32, []),
        "gams_cuda_create": (i32, [vp, i32, C.POINTER(vp)]),
        "gams_cuda_destroy": (None, [vp]),
        "gams_cuda_ctx_num_devices": (i32, [vp]),
        "gams_cuda_set_zscore_mode": (i32, [vp, i32]),
        "gams_cuda_zscore_stats": (i32, [vp, vp, i32]),
        "gams_cuda_set_gunzip_options": (i32, [vp, i32, i32]),
        "gams_cuda_host_alloc": (vp, [sz]),
        "gams_cuda_host_free": (None, [vp]),
        "gams_cuda_memcpy_probe": (i32, [vp, vp, sz, i32]),
        "gams_cuda_wave_nwin": (i64, [i64, i32, i32]),
        "gams_cuda_wave_win_offsets": (i32, [i64, vp, i32, i32, vp]),
        "gams_cuda_partition": (i32, [i64, vp, i32, vp]),
        "gams_cuda_wave_batch": (i32, [vp, i64, vp, vp, i32, i32, u64, f32, f32, vp, vp, vp, vp]),
        "gams_cuda_wave_gz_batch": (i32, [vp, i64, vp, vp, vp, i32, i32, u64, f32, f32, vp, vp, vp, vp, vp]),
        "gams_cuda_thresholding_batch": (i32, [vp, i64, vp, vp, u64, f32, f32, vp]),
        "gams_cuda_wave_peaks_batch": (i32, [vp, i64, vp, vp, i32, i32, u64, f32, f32, f32, vp, u64, vp, vp, vp, vp, vp, vp]),
        "gams_cuda_wave_peaks_gz_batch": (i32, [vp, i64, vp, vp, vp, i32, i32, u64, f32, f32, f32, vp, u64, vp, vp, vp, vp, vp, vp, vp]),
        "gams_cuda_wave_peaks_dev": (i32, [vp, vp, vp, vp, f32, u64, vp, vp, vp, vp, vp, vp]),
        "gams_cuda_peak_link_batch": (i32, [vp, i64, vp, vp, vp, i64] + [vp] * 11),
        "gams_cuda_wave_peaks_text": (i32, [C.c_char_p, i32, i32, i32, i64, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(sz)]),
        "gams_cuda_wave_plan_create": (i32, [vp, i32, i64, vp, i32, i32, u64, C.POINTER(vp)]),
        "gams_cuda_wave_plan_destroy": (None, [vp]),
        "gams_cuda_wave_plan_total_windows": (u64, [vp]),
        "gams_cuda_wave_plan_total_bases": (u64, [vp]),
        "gams_cuda_wave_plan_seq_capacity": (u64, [vp]),
        "gams_cuda_wave_gc_dev": (i32, [vp, vp, vp, vp, vp]),
        "gams_cuda_wave_signal_dev": (i32, [vp, vp, vp, f32, f32, vp, vp]),
        "gams_cuda_wave_dev": (i32, [vp, vp, vp, f32, f32, vp, vp, vp, vp]),
        "gams_cuda_wave_plan_fused": (i32, [vp, f32, f32]),
        "gams_cuda_wave_plan_launches": (u64, [vp]),
        "gams_cuda_anno_coverage": (i32, [vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]),
        "gams_cuda_gen_ctgs": (i32, [vp, i64, vp, vp, i32, i32, i32, u64, vp, vp, vp, vp]),
        "gams_cuda_gc_ranges": (i32, [vp, i64, vp, vp, i64, vp, vp, vp, vp]),
        "gams_cuda_sw_batch": (i32, [vp, i64, vp, vp, vp, i64, vp, vp, vp, i32, i32, i32] + [vp] * 9),
        "gams_cuda_sw_count_batch": (i32, [vp, i64, vp, vp, vp, i64, vp, vp, vp, i32, i32, i32, vp, vp, vp] + [vp] * 10),
        "gams_cuda_count_overlaps": (i32, [vp, i64, vp, vp, vp, i64, vp, vp, vp, vp]),
        "gams_cuda_count_plan_create": (i32, [vp, i32, i64, vp, C.POINTER(vp)]),
        "gams_cuda_count_plan_destroy": (None, [vp]),
        "gams_cuda_count_overlaps_dev": (i32, [vp, vp, vp, vp, i64, vp, vp, vp, i32, vp]),
        "gams_cuda_count_plan_launches": (u64, [vp]),
        "gams_cuda_count_plan_invalidate": (None, [vp]),
        "gams_cuda_count_windows": (i32, [vp, i64, vp, vp, i32, i32, vp, vp, vp, vp]),
        "gams_cuda_cwin_plan_create": (i32, [vp, i32, i64, vp, vp, i32, i32, vp, C.POINTER(vp)]),
        "gams_cuda_cwin_plan_destroy": (None, [vp]),
        "gams_cuda_cwin_plan_total_windows": (u64, [vp]),
        "gams_cuda_cwin_plan_launches": (u64, [vp]),
        "gams_cuda_cwin_plan_invalidate": (None, [vp]),
        "gams_cuda_count_windows_dev": (i32, [vp, vp, vp, vp, vp]),
        "gams_cuda_gz_sizes": (i32, [i64, vp, vp, vp]),
        "gams_cuda_gunzip_batch": (i32, [vp, i64, vp, vp, vp, vp, vp]),
        "gams_cuda_gunzip_dev": (i32, [vp, i32, vp, i64, vp, vp, vp, vp, vp]),
        "gams_cuda_parse_ranges": (i32, [vp, vp, u64, i64, vp, vp, u64, C.POINTER(u64), vp, vp, vp, vp]),
        "gams_cuda_locate_count_text": (i32, [vp, vp, u64, i64, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(sz),
                                              C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "gams_cuda_locate_count_lines": (i32, [vp, vp, u64, i64, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(sz),
                                               C.POINTER(vp), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "gams_cuda_build_idx_rg": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "gams_cuda_locate_ctg": (i32, [vp, i64, vp, vp, vp, i64, vp, vp, vp, vp]),
        "gams_cuda_lapper_blob_view": (i32, [vp, sz, C.POINTER(u64), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint32)]),
        "gams_cuda_wave_text": (i32, [C.c_char_p, i32, i32, i32, i64, vp, vp, f32, i32, C.POINTER(vp), C.POINTER(sz)]),
        "gams_cuda_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here == the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    L._gams_signatures = sig
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise GamsCudaError(rc, load().gams_cuda_last_error().decode(errors="replace"))
