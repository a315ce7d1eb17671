"""ctypes mirror of include/pcramp_gpu.h.

One method per C entry point, same names and argument meaning.  numpy arrays in, numpy arrays out;
words are (n, 2) uint64 arrays of {buffer[0], buffer[1]}.  There is no CPU fallback: if the CUDA
library is missing or no GPU is present this module raises.
"""
import ctypes
import os

import numpy as np

TARGET, BACKGROUND, MULTIPLEX = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCRAMP_GPU_LIB") or os.path.join(_HERE, "csrc", "libpcramp_gpu.so")

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f32p = ctypes.POINTER(ctypes.c_float)


class GpuError(RuntimeError):
    pass


class Stats(ctypes.Structure):
    _fields_ = [
        ("n_patterns", ctypes.c_uint64),
        ("n_positions", ctypes.c_uint64),
        ("n_hits", ctypes.c_uint64),
        ("n_entries", ctypes.c_uint64),
        ("n_keys", ctypes.c_uint64),
        ("kernel_launches", ctypes.c_uint64),
        ("n_seeded", ctypes.c_uint64),
        ("n_seed_entries", ctypes.c_uint64),
        ("n_indexed", ctypes.c_uint64),
        ("n_index_queries", ctypes.c_uint64),
        ("n_index_entries", ctypes.c_uint64),
        ("ms_seed", ctypes.c_float),
        ("ms_scan", ctypes.c_float),
        ("ms_edge", ctypes.c_float),
        ("ms_db", ctypes.c_float),
        ("ms_score", ctypes.c_float),
        ("ms_index_kernel", ctypes.c_float),
        ("ms_index_build", ctypes.c_float),
        ("index_bytes", ctypes.c_uint64),
        ("n_index_builds", ctypes.c_uint64),
        ("n_index_stale", ctypes.c_uint64),
        ("n_fast", ctypes.c_uint64),
        ("n_fast_redo", ctypes.c_uint64),
        ("n_edge_words", ctypes.c_uint64),
        ("edge_table_bytes", ctypes.c_uint64),
        ("ms_edge_table_build", ctypes.c_float),
        ("edge_table_used", ctypes.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ThermoStats(ctypes.Structure):
    _fields_ = [
        ("n_problems", ctypes.c_uint64),
        ("dp_cells", ctypes.c_uint64),
        ("kernel_launches", ctypes.c_uint64),
        ("ms_kernel", ctypes.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class OptimizeOptions(ctypes.Structure):
    """pcramp_gpu_optimize_options; defaults = the reference's (pcramp.h:14-51)"""
    _fields_ = [
        ("target_threshold", ctypes.c_float), ("target_search_multiplier", ctypes.c_float),
        ("target_amplicon_min", ctypes.c_int), ("target_amplicon_max", ctypes.c_int),
        ("background_threshold", ctypes.c_float), ("background_search_multiplier", ctypes.c_float),
        ("background_amplicon_min", ctypes.c_int), ("background_amplicon_max", ctypes.c_int),
        ("use_taq_mama", ctypes.c_int), ("use_multiplex", ctypes.c_int), ("degen", ctypes.c_uint32),
        ("primer_min", ctypes.c_int), ("primer_max", ctypes.c_int),
        ("salt", ctypes.c_float), ("primer_strand", ctypes.c_float),
        ("primer_tm_min", ctypes.c_float), ("primer_tm_max", ctypes.c_float), ("max_hairpin", ctypes.c_float),
    ]

    def __init__(self, **kw):
        d = dict(target_threshold=1.0, target_search_multiplier=0.9, target_amplicon_min=80, target_amplicon_max=200,
                 background_threshold=0.8, background_search_multiplier=0.9, background_amplicon_min=0, background_amplicon_max=2000,
                 use_taq_mama=0, use_multiplex=1, degen=1, primer_min=18, primer_max=25, salt=0.05, primer_strand=900.0e-9,
                 primer_tm_min=50.0, primer_tm_max=75.0, max_hairpin=40.0)
        d.update(kw)
        super().__init__(**d)


class RandomAssayOptions(ctypes.Structure):
    """pcramp_gpu_random_assay_options (include/pcramp_gpu.h); defaults = the reference's (pcramp.h:14-31)"""
    _fields_ = [("primer_min", ctypes.c_int), ("primer_max", ctypes.c_int), ("amplicon_min", ctypes.c_int), ("amplicon_max", ctypes.c_int),
                ("degen", ctypes.c_uint32), ("salt", ctypes.c_float), ("primer_strand", ctypes.c_float), ("primer_tm_min", ctypes.c_float),
                ("primer_tm_max", ctypes.c_float), ("max_hairpin", ctypes.c_float), ("max_dimer", ctypes.c_float)]

    def __init__(self, primer_range=(18, 25), amplicon_range=(80, 200), degen=1, salt=0.05, primer_strand=900.0e-9, primer_tm_range=(50.0, 75.0),
                 max_hairpin=40.0, max_dimer=40.0):
        super().__init__(primer_range[0], primer_range[1], amplicon_range[0], amplicon_range[1], degen, salt, primer_strand, primer_tm_range[0],
                         primer_tm_range[1], max_hairpin, max_dimer)



class DesignOptions(ctypes.Structure):
    """pcramp_gpu_design_options (include/pcramp_gpu.h); filled by pcramp_gpu_design_default_options"""
    _fields_ = [("num_trial", ctypes.c_uint32), ("n_streams", ctypes.c_uint32), ("degen", ctypes.c_uint32),
                ("optimize_5", ctypes.c_int), ("optimize_3", ctypes.c_int), ("primer_min", ctypes.c_int), ("primer_max", ctypes.c_int),
                ("primer_tm_min", ctypes.c_float), ("primer_tm_max", ctypes.c_float), ("primer_strand", ctypes.c_float), ("salt", ctypes.c_float),
                ("max_hairpin", ctypes.c_float), ("max_dimer", ctypes.c_float),
                ("target_amplicon_min", ctypes.c_int), ("target_amplicon_max", ctypes.c_int),
                ("background_amplicon_min", ctypes.c_int), ("background_amplicon_max", ctypes.c_int),
                ("target_threshold", ctypes.c_float), ("target_search_multiplier", ctypes.c_float),
                ("background_threshold", ctypes.c_float), ("background_search_multiplier", ctypes.c_float),
                ("min_target_cover", ctypes.c_float), ("max_background_cover", ctypes.c_float),
                ("pack_max_degen", ctypes.c_uint32), ("pack_min_gc", ctypes.c_float), ("pack_max_gc", ctypes.c_float),
                ("use_taq_mama", ctypes.c_int), ("use_multiplex", ctypes.c_int)]


class DesignResult(ctypes.Structure):
    """pcramp_gpu_design_result (include/pcramp_gpu.h)"""
    _fields_ = [("found", ctypes.c_int), ("iteration", ctypes.c_uint32), ("major_id", ctypes.c_uint32), ("minor_id", ctypes.c_uint32),
                ("targets_remaining", ctypes.c_uint32), ("num_active_target", ctypes.c_uint32), ("num_active_background", ctypes.c_uint32),
                ("trial", ctypes.c_uint32), ("active_target_norm", ctypes.c_float), ("active_background_norm", ctypes.c_float),
                ("f", ctypes.c_uint64 * 2), ("r", ctypes.c_uint64 * 2), ("degeneracy_f", ctypes.c_double), ("degeneracy_r", ctypes.c_double),
                ("reused_f", ctypes.c_int), ("reused_r", ctypes.c_int),
                ("target_coverage", ctypes.c_float), ("background_coverage", ctypes.c_float), ("oligo_overlap", ctypes.c_float),
                ("n_target_entries", ctypes.c_uint64), ("n_background_entries", ctypes.c_uint64), ("n_amplicons_added", ctypes.c_uint64),
                ("n_splits", ctypes.c_uint64), ("n_multiplex_keys", ctypes.c_uint64),
                ("ms_total", ctypes.c_float), ("ms_candidates", ctypes.c_float), ("ms_select_background", ctypes.c_float),
                ("ms_select_target", ctypes.c_float), ("ms_optimize", ctypes.c_float), ("ms_screen", ctypes.c_float), ("ms_accept", ctypes.c_float)]

    def as_dict(self):
        out = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            out[k] = [int(v[0]), int(v[1])] if k in ("f", "r") else v
        return out


MOVES = {"IncreaseDegeneracy": 0, "DecreaseDegeneracy": 1, "Trim5": 2, "Trim3": 3, "Grow5": 4, "Grow3": 5}

# op codes of pcramp_gpu_thermo_batch (include/pcramp_gpu.h)
TM_PM_DUPLEX, TM_HAIRPIN, TM_HOMODIMER, TM_HETERODIMER, TM_HETERODIMER_DIAG, TM_HOMODIMER_DIAG = range(6)
THERMO_STRIDE = 33  # bytes per string slot used by this wrapper (32 bases + NUL)

_lib = None

# name -> (restype, argtypes); also the list of symbols tests check against include/pcramp_gpu.h
SIGNATURES = {
    "pcramp_gpu_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "pcramp_gpu_destroy": (None, [ctypes.c_void_p]),
    "pcramp_gpu_create_worker": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "pcramp_gpu_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "pcramp_gpu_stream": (ctypes.c_void_p, [ctypes.c_void_p]),
    "pcramp_gpu_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "pcramp_gpu_upload_sequences": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, _u8p, _u64p, _u32p, _f32p]),
    "pcramp_gpu_set_active": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u8p]),
    "pcramp_gpu_set_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _f32p]),
    "pcramp_gpu_split_sequence": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32]),
    "pcramp_gpu_split_sequences": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, _u32p, _u32p]),
    "pcramp_gpu_unique_amplicons": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p, _u64p, ctypes.c_uint32, ctypes.c_float, ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int, _u64p, _u64p, _u64p]),
    "pcramp_gpu_unique_amplicons_copy": (ctypes.c_int, [ctypes.c_void_p, _u32p, _u64p, ctypes.c_char_p, _u32p, _u32p]),
    "pcramp_gpu_pool_amplicon_coverage": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p, _u64p, ctypes.c_uint32, ctypes.c_float, ctypes.c_int,
                                                         ctypes.c_int, ctypes.c_float, ctypes.c_int, _f32p]),
    "pcramp_gpu_accept_assay": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _u64p, _u64p]),
    "pcramp_gpu_random_assays": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, _u32p, _u32p, ctypes.POINTER(RandomAssayOptions),
                                                _u64p, _u64p, _u32p]),
    "pcramp_gpu_upload_fasta_groups": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p), _u64p, _u32p,
                                                      ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p),
                                                      _u32p]),
    "pcramp_gpu_best_assay": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, _f32p, _f32p, _f32p, _u64p, _u64p, ctypes.c_float,
                                             ctypes.POINTER(ctypes.c_int64), _f32p, _f32p, ctypes.POINTER(ctypes.c_double)]),
    "pcramp_gpu_sw_timing": (ctypes.c_int, [ctypes.c_void_p, _f32p]),
    "pcramp_gpu_pack": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_float,
                                       ctypes.c_uint32, ctypes.c_uint64, _u64p, _i32p, _u32p, _u64p]),
    "pcramp_gpu_select_words": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p, _u64p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_float, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_uint32, _u64p, _u64p]),
    "pcramp_gpu_db_size": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p, _u64p]),
    "pcramp_gpu_db_copy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p, _u32p, _i32p, _u32p, _u32p]),
    "pcramp_gpu_keys_copy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p]),
    "pcramp_gpu_score_pairs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p, _u64p, ctypes.c_uint32, ctypes.c_float, ctypes.c_float,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _u32p]),
    "pcramp_gpu_score_variants": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p, _u64p, _u64p, _u64p, ctypes.c_uint32, ctypes.c_float,
                                                 ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _u32p]),
    "pcramp_gpu_optimize": (ctypes.c_int, [ctypes.c_void_p, _u64p, _u64p, ctypes.c_uint32, _i32p, ctypes.c_uint32, ctypes.c_void_p, _f32p, _f32p,
                                           _f32p, _u32p]),
    "pcramp_gpu_upload_fasta": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p), _u64p, ctypes.c_uint64,
                                               ctypes.c_uint64, ctypes.c_uint32, ctypes.POINTER(ctypes.c_char_p), _u32p]),
    "pcramp_gpu_fasta_records": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u32p, _u64p, _u32p, _u32p, _f32p]),
    "pcramp_gpu_fasta_free": (None, [ctypes.c_void_p]),
    "pcramp_gpu_fasta_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _f32p, _f32p, _u64p, _u64p]),
    "pcramp_fasta_scan": (ctypes.c_uint32, [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint32, _u64p, _u32p, _u64p, _u64p, _f32p]),
    "pcramp_gpu_sequences_copy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u32p, _u64p, _u64p, _u32p, _u8p]),
    "pcramp_gpu_multiplex_keys": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, _u64p]),
    "pcramp_gpu_multiplex_keys_copy": (ctypes.c_int, [ctypes.c_void_p, _u64p]),
    "pcramp_gpu_set_pool": (ctypes.c_int, [ctypes.c_void_p, _u64p, _u64p, ctypes.c_uint32]),
    "pcramp_gpu_multiplex_coverage": (ctypes.c_int, [ctypes.c_void_p, _u64p, _u64p, _u64p, _u64p, ctypes.c_uint32, ctypes.c_float, ctypes.c_int,
                                                     _f32p]),
    "pcramp_gpu_oligo_overlap": (ctypes.c_int, [ctypes.c_void_p, _u64p, _u64p, ctypes.c_uint32, _f32p]),
    "pcramp_gpu_stage_pairs": (ctypes.c_int, [ctypes.c_void_p, _u64p, _u64p, ctypes.c_uint32]),
    "pcramp_gpu_set_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32]),
    "pcramp_gpu_select_words_staged": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_uint32,
                                                      ctypes.c_float, ctypes.c_float, ctypes.c_uint32, _u64p, _u64p]),
    "pcramp_gpu_score_pairs_staged": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_int]),
    "pcramp_gpu_device_coverage": (ctypes.c_void_p, [ctypes.c_void_p]),
    "pcramp_gpu_device_bitsets": (ctypes.c_void_p, [ctypes.c_void_p]),
    "pcramp_gpu_device_bitsets_pass1": (ctypes.c_void_p, [ctypes.c_void_p]),
    "pcramp_gpu_merge_shards": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, _u32p, _f32p, ctypes.c_uint32,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    "pcramp_gpu_exchange_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, _u32p, ctypes.c_uint32, _f32p]),
    "pcramp_gpu_exchange_buffer": (ctypes.c_void_p, [ctypes.c_void_p]),
    "pcramp_gpu_exchange_ipc_handle": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "pcramp_gpu_exchange_connect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "pcramp_gpu_exchange_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "pcramp_gpu_exchange_coverage": (ctypes.c_void_p, [ctypes.c_void_p]),
    "pcramp_gpu_exchange_bitsets": (ctypes.c_void_p, [ctypes.c_void_p]),
    "pcramp_gpu_exchange_words": (ctypes.c_uint32, [ctypes.c_void_p]),
    "pcramp_gpu_exchange_fetch": (ctypes.c_int, [ctypes.c_void_p, _f32p, _u32p]),
    "pcramp_gpu_exchange_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pcramp_gpu_design_default_options": (None, [ctypes.POINTER(DesignOptions)]),
    "pcramp_gpu_design_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(DesignOptions), ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p)]),
    "pcramp_gpu_design_destroy": (None, [ctypes.c_void_p]),
    "pcramp_gpu_design_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "pcramp_gpu_design_iteration": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(DesignResult)]),
    "pcramp_gpu_design_matches": (ctypes.c_int, [ctypes.c_void_p, _u32p, _u32p]),
    "pcramp_gpu_design_active": (ctypes.c_int, [ctypes.c_void_p, _u8p, _u32p]),
    "pcramp_gpu_exchange_pairs": (ctypes.c_uint32, [ctypes.c_void_p]),
    "pcramp_gpu_exchange_set_timeout_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32]),
    "pcramp_gpu_exchange_status": (ctypes.c_int, [ctypes.c_void_p, _u32p]),
    "pcramp_gpu_reduce_best": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_double, ctypes.c_int64,
                                               _u32p, _f32p, _f32p, _f32p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "pcramp_gpu_measure_int_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]),
    "pcramp_gpu_measure_int32_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]),
    "pcramp_gpu_bitset_words": (ctypes.c_uint32, [ctypes.c_void_p, ctypes.c_int]),
    "pcramp_gpu_fetch_results": (ctypes.c_int, [ctypes.c_void_p, _f32p, _u32p]),
    "pcramp_gpu_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]),
    "pcramp_gpu_get_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Stats)]),
    "pcramp_gpu_thermo_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint32,
                                               ctypes.c_float, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p]),
    "pcramp_gpu_thermo_words": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, _u64p, _u64p, ctypes.c_float, _f32p, _f32p, _f32p, _f32p,
                                               _f32p, _f32p]),
    "pcramp_gpu_thermo_stage": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint32,
                                               ctypes.c_float, _f32p, _f32p]),
    "pcramp_gpu_thermo_run_staged": (ctypes.c_int, [ctypes.c_void_p]),
    "pcramp_gpu_thermo_fetch": (ctypes.c_int, [ctypes.c_void_p, _f32p, _f32p, _f32p, _f32p]),
    "pcramp_gpu_is_valid": (ctypes.c_int, [ctypes.c_void_p, _u64p, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                           ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, _u8p]),
    "pcramp_gpu_max_dimer_tm": (ctypes.c_int, [ctypes.c_void_p, _u64p, _u64p, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                               _f32p]),
    "pcramp_gpu_multiplex_compatible": (ctypes.c_int, [ctypes.c_void_p, _u64p, _u64p, ctypes.c_uint32, _u64p, _u64p, ctypes.c_uint32,
                                                       ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, _u8p]),
    "pcramp_gpu_sw_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, _u64p, _u64p, _i32p, _i32p, _i32p, _i32p, _i32p, _u8p]),
    "pcramp_gpu_background_match": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p, _u64p, ctypes.c_uint32, ctypes.c_float, ctypes.c_float,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, _u32p, _u64p]),
    "pcramp_gpu_multiplex_background_match": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, _u64p, _u64p, ctypes.c_uint32, ctypes.c_float,
                                                             ctypes.c_int, _u32p]),
    "pcramp_gpu_get_thermo_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ThermoStats)]),
    "pcramp_word_from_string": (None, [ctypes.c_char_p, ctypes.c_int, _u64p]),
    "pcramp_word_to_string": (ctypes.c_int, [_u64p, ctypes.c_char_p]),
    "pcramp_word_and": (ctypes.c_uint32, [_u64p, _u64p]),
    "pcramp_word_size": (ctypes.c_uint32, [_u64p]),
    "pcramp_word_start": (ctypes.c_int, [_u64p]),
    "pcramp_word_stop": (ctypes.c_int, [_u64p]),
    "pcramp_word_complement": (None, [_u64p, _u64p]),
    "pcramp_word_center": (None, [_u64p, _u64p]),
    "pcramp_word_max_overlap": (ctypes.c_float, [_u64p, _u64p]),
}


def load_library():
    """dlopen the CUDA library; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpuError("%s is missing: run `python -m pcramp_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _words(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2 or a.shape[1] != 2:
        raise ValueError("words must have shape (n, 2)")
    return a


class PcrampGpu:
    """One context = one GPU (include/pcramp_gpu.h)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.pcramp_gpu_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise GpuError("pcramp_gpu_create failed (rc=%d): no usable CUDA device %d; there is no CPU fallback" % (rc, device))
        self.h = h
        self.n_seq = {}
        self.n_pairs = 0
        self._db = {}

    def worker(self):
        """pcramp_gpu_create_worker: a context with its own stream / scratch / database that reads this context's collections and text
        index in place.  Drive it from its own host thread (ctypes releases the GIL during a call); close it before this context."""
        h = ctypes.c_void_p()
        self._ck(self.lib.pcramp_gpu_create_worker(self.h, ctypes.byref(h)))
        w = PcrampGpu.__new__(PcrampGpu)
        w.lib, w.h, w.n_seq, w.n_pairs, w._db = self.lib, h, dict(self.n_seq), 0, {}
        self._workers = getattr(self, "_workers", [])
        self._workers.append(w)
        return w

    def close(self):
        for w in getattr(self, "_workers", []):
            w.close()
        self._workers = []
        if getattr(self, "h", None):
            self.lib.pcramp_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise GpuError(self.lib.pcramp_gpu_last_error(self.h).decode())

    @property
    def stream(self):
        return self.lib.pcramp_gpu_stream(self.h)

    def synchronize(self):
        self._ck(self.lib.pcramp_gpu_synchronize(self.h))

    # ---- sequences ----------------------------------------------------------------------------
    def upload_sequences(self, kind, nibbles, byte_off, length, weight=None):
        nibbles = np.ascontiguousarray(nibbles, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.uint64)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        if weight is not None:
            weight = np.ascontiguousarray(weight, dtype=np.float32)
        n = len(length)
        self._ck(self.lib.pcramp_gpu_upload_sequences(self.h, kind, n, _ptr(nibbles, _u8p), _ptr(byte_off, _u64p), _ptr(length, _u32p),
                                                      _ptr(weight, _f32p)))
        self.n_seq[kind] = n

    def set_active(self, kind, active):
        active = np.ascontiguousarray(active, dtype=np.uint8)
        self._ck(self.lib.pcramp_gpu_set_active(self.h, kind, _ptr(active, _u8p)))

    def split_sequence(self, kind, seq, pos):
        self._ck(self.lib.pcramp_gpu_split_sequence(self.h, kind, int(seq), int(pos)))

    def split_sequences(self, kind, seq, pos):
        seq, pos = np.ascontiguousarray(seq, dtype=np.uint32), np.ascontiguousarray(pos, dtype=np.uint32)
        self._ck(self.lib.pcramp_gpu_split_sequences(self.h, kind, len(seq), _ptr(seq, _u32p), _ptr(pos, _u32p)))

    # ---- multiplex bookkeeping (amplicon.cuh) --------------------------------------------------
    def unique_amplicons(self, kind, f, r, threshold, amplicon_min=80, amplicon_max=200, want_bounds=True, copy=True):
        """PCR::collect_unique_amplicons for every pair -> list per pair of (amplicon strings in the returned deque's order,
        bounds (n, 3) uint32 {index, begin, end} in push order); copy=False leaves everything on the device and returns the counts"""
        f, r = _words(f), _words(r)
        cnt = np.zeros(3, np.uint64)
        self._ck(self.lib.pcramp_gpu_unique_amplicons(self.h, kind, _ptr(f, _u64p), _ptr(r, _u64p), len(f), float(threshold), int(amplicon_min),
                                                      int(amplicon_max), int(bool(want_bounds)), _ptr(cnt[0:1], _u64p), _ptr(cnt[1:2], _u64p),
                                                      _ptr(cnt[2:3], _u64p)))
        n_amp, n_bases, n_bounds = (int(x) for x in cnt)
        if not copy:
            return n_amp, n_bases, n_bounds
        pair_off = np.zeros(len(f) + 1, np.uint32)
        text_off = np.zeros(n_amp + 1, np.uint64)
        text = ctypes.create_string_buffer(max(1, n_bases))
        b_pair = np.zeros(max(1, n_bounds), np.uint32)
        bounds = np.zeros((max(1, n_bounds), 3), np.uint32)
        self._ck(self.lib.pcramp_gpu_unique_amplicons_copy(self.h, _ptr(pair_off, _u32p), _ptr(text_off, _u64p), text, _ptr(b_pair, _u32p),
                                                           _ptr(bounds, _u32p) if want_bounds else None))
        raw = text.raw[:n_bases]
        b_pair, bounds = b_pair[:n_bounds], bounds[:n_bounds]
        out = []
        for p in range(len(f)):
            amps = [raw[int(text_off[u]):int(text_off[u + 1])].decode("ascii") for u in range(int(pair_off[p]), int(pair_off[p + 1]))]
            out.append((amps, bounds[b_pair == p].copy() if want_bounds else None))
        return out

    def pool_amplicon_coverage(self, kind, f, r, target_threshold, amplicon_min, amplicon_max, background_threshold, use_taq_mama=False):
        f, r = _words(f), _words(r)
        cov = np.zeros(len(f), np.float32)
        self._ck(self.lib.pcramp_gpu_pool_amplicon_coverage(self.h, kind, _ptr(f, _u64p), _ptr(r, _u64p), len(f), float(target_threshold),
                                                            int(amplicon_min), int(amplicon_max), float(background_threshold), int(use_taq_mama),
                                                            _ptr(cov, _f32p)))
        return cov

    def accept_assay(self, pair=0, pack_max_degen=256, min_oligo_length=18):
        """main.cpp:989-1017,1123 for one pair of the last unique_amplicons(want_bounds=True) call -> (amplicons appended, multiplex keys)"""
        out = np.zeros(2, np.uint64)
        self._ck(self.lib.pcramp_gpu_accept_assay(self.h, int(pair), int(pack_max_degen), int(min_oligo_length), _ptr(out[0:1], _u64p),
                                                  _ptr(out[1:2], _u64p)))
        self.n_seq[MULTIPLEX] = self.n_seq.get(MULTIPLEX, 0) + int(out[0])
        return int(out[0]), int(out[1])

    def best_assay(self, target_cov, background_cov, overlap, f, r, max_background_cover=0.0):
        """main.cpp:829-858 over the trials in order -> (index or -1, accuracy, overlap, total degeneracy)"""
        t = np.ascontiguousarray(target_cov, dtype=np.float32)
        b = np.ascontiguousarray(background_cov, dtype=np.float32)
        o = None if overlap is None else np.ascontiguousarray(overlap, dtype=np.float32)
        f, r = _words(f), _words(r)
        idx, acc, ov, dg = ctypes.c_int64(-1), np.zeros(1, np.float32), np.zeros(1, np.float32), ctypes.c_double(0.0)
        self._ck(self.lib.pcramp_gpu_best_assay(self.h, len(t), _ptr(t, _f32p), _ptr(b, _f32p), _ptr(o, _f32p), _ptr(f, _u64p), _ptr(r, _u64p),
                                                float(max_background_cover), ctypes.byref(idx), _ptr(acc, _f32p), _ptr(ov, _f32p), ctypes.byref(dg)))
        return int(idx.value), float(acc[0]), float(ov[0]), float(dg.value)

    def pack(self, kind, seq, pack_max_degen=256, pack_min_gc=0.0, pack_max_gc=1.0, min_oligo_length=18):
        """Sequence::pack of one sequence -> (words, loc, strand), unordered."""
        n = ctypes.c_uint64()
        args = (self.h, kind, int(seq), int(pack_max_degen), float(pack_min_gc), float(pack_max_gc), int(min_oligo_length))
        self._ck(self.lib.pcramp_gpu_pack(*args, 0, None, None, None, ctypes.byref(n)))
        words = np.zeros((n.value, 2), np.uint64)
        loc = np.zeros(n.value, np.int32)
        strand = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(self.lib.pcramp_gpu_pack(*args, n.value, _ptr(words, _u64p), _ptr(loc, _i32p), _ptr(strand, _u32p), ctypes.byref(n)))
        return words, loc, strand

    # ---- seed scan ----------------------------------------------------------------------------
    def select_words(self, kind, f, r, threshold, optimize_5=False, optimize_3=False, pack_max_degen=256, pack_min_gc=0.0,
                     pack_max_gc=1.0, min_oligo_length=18, want_keys=True, want_entries=True):
        """-> (entries, keys); want_keys=False skips the canonical order / keys() numbering (keys = None); want_entries=False does not
        ask for the entry count either (entries = None), which lets a batch in the fast form return without a host round trip"""
        f, r = _words(f), _words(r)
        ne, nk = ctypes.c_uint64(), ctypes.c_uint64()
        self._ck(self.lib.pcramp_gpu_select_words(self.h, kind, _ptr(f, _u64p), _ptr(r, _u64p), len(f), int(optimize_5), int(optimize_3),
                                                  float(threshold), int(pack_max_degen), float(pack_min_gc), float(pack_max_gc),
                                                  int(min_oligo_length), ctypes.byref(ne) if want_entries else None,
                                                  ctypes.byref(nk) if want_keys else None))
        self.n_pairs = len(f)
        return (ne.value if want_entries else None), (nk.value if want_keys else None)

    def db_size(self, kind):
        ne, nk = ctypes.c_uint64(), ctypes.c_uint64()
        self._ck(self.lib.pcramp_gpu_db_size(self.h, kind, ctypes.byref(ne), ctypes.byref(nk)))
        return ne.value, nk.value

    def db_copy(self, kind):
        ne, nk = self.db_size(kind)
        words = np.zeros((ne, 2), np.uint64)
        index = np.zeros(ne, np.uint32)
        loc = np.zeros(ne, np.int32)
        strand = np.zeros(ne, np.uint32)
        key = np.zeros(ne, np.uint32)
        self._ck(self.lib.pcramp_gpu_db_copy(self.h, kind, _ptr(words, _u64p), _ptr(index, _u32p), _ptr(loc, _i32p), _ptr(strand, _u32p),
                                             _ptr(key, _u32p)))
        return words, index, loc, strand, key

    def keys_copy(self, kind):
        ne, nk = self.db_size(kind)
        keys = np.zeros((nk, 2), np.uint64)
        self._ck(self.lib.pcramp_gpu_keys_copy(self.h, kind, _ptr(keys, _u64p)))
        return keys

    # ---- pair scoring -------------------------------------------------------------------------
    def score_pairs(self, kind, f, r, search_threshold, detect_threshold, amplicon_min=80, amplicon_max=200, use_taq_mama=False, out=None):
        """-> (coverage float32[n], bitsets uint32[n, words]); out = (coverage, bitsets) arrays to fill instead (e.g. views of pinned
        host memory: the device-to-host copy then runs at full PCIe rate instead of through the driver's staging buffer)"""
        f, r = _words(f), _words(r)
        n = len(f)
        nw = (self.n_seq[kind] + 31) // 32
        if out is not None:
            cov, bits = out
            assert cov.dtype == np.float32 and bits.dtype == np.uint32 and cov.shape == (n,) and bits.shape == (n, nw)
            assert cov.flags.c_contiguous and bits.flags.c_contiguous
        else:
            cov = np.zeros(n, np.float32)
            bits = np.zeros((n, nw), np.uint32)
        self._ck(self.lib.pcramp_gpu_score_pairs(self.h, kind, _ptr(f, _u64p), _ptr(r, _u64p), n, float(search_threshold),
                                                 float(detect_threshold), int(amplicon_min), int(amplicon_max), int(use_taq_mama),
                                                 _ptr(cov, _f32p), _ptr(bits, _u32p)))
        return cov, bits

    def score_variants(self, kind, base_f, base_r, var_f, var_r, search_threshold, detect_threshold, amplicon_min=80, amplicon_max=200,
                       use_taq_mama=False):
        """move scoring: trial oligos (var_*) against the candidate amplicons of the unmoved assays (base_*)"""
        bf, br, vf, vr = _words(base_f), _words(base_r), _words(var_f), _words(var_r)
        n = len(bf)
        nw = (self.n_seq[kind] + 31) // 32
        cov = np.zeros(n, np.float32)
        bits = np.zeros((n, nw), np.uint32)
        self._ck(self.lib.pcramp_gpu_score_variants(self.h, kind, _ptr(bf, _u64p), _ptr(br, _u64p), _ptr(vf, _u64p), _ptr(vr, _u64p), n,
                                                    float(search_threshold), float(detect_threshold), int(amplicon_min), int(amplicon_max),
                                                    int(use_taq_mama), _ptr(cov, _f32p), _ptr(bits, _u32p)))
        return cov, bits

    def optimize(self, f, r, moves, options):
        """optimize() for a batch of trials -> (f, r, target_coverage, background_coverage, oligo_overlap, iterations)"""
        f, r = _words(f).copy(), _words(r).copy()
        n = len(f)
        mv = np.ascontiguousarray([MOVES[m] if isinstance(m, str) else int(m) for m in moves], dtype=np.int32)
        tc, bc, ov = (np.zeros(n, np.float32) for _ in range(3))
        it = np.zeros(n, np.uint32)
        self._ck(self.lib.pcramp_gpu_optimize(self.h, _ptr(f, _u64p), _ptr(r, _u64p), n, _ptr(mv, _i32p), len(mv), ctypes.byref(options),
                                              _ptr(tc, _f32p), _ptr(bc, _f32p), _ptr(ov, _f32p), _ptr(it, _u32p)))
        return f, r, tc, bc, ov, it

    # ---- FASTA ingest on the device (fasta.cuh) ------------------------------------------------
    def upload_fasta(self, kind, texts, min_length=0, max_length=1 << 40, ignore=()):
        """texts: list of bytes (inflated FASTA files, in order) -> [(file, defline bytes, length, weight)] of the kept records"""
        blobs = [bytes(t) for t in texts]
        arr = (ctypes.c_char_p * max(1, len(blobs)))(*blobs)
        nb = np.array([len(b) for b in blobs], np.uint64)
        ig = [x.encode() for x in ignore]
        iga = (ctypes.c_char_p * max(1, len(ig)))(*ig)
        n = np.zeros(1, np.uint32)
        self._ck(self.lib.pcramp_gpu_upload_fasta(self.h, kind, len(blobs), arr, _ptr(nb, _u64p), int(min_length), int(max_length), len(ig), iga,
                                                  _ptr(n, _u32p)))
        n = int(n[0])
        self.n_seq[kind] = n
        f, off, dl, ln, w = np.zeros(n, np.uint32), np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.float32)
        self._ck(self.lib.pcramp_gpu_fasta_records(self.h, kind, _ptr(f, _u32p), _ptr(off, _u64p), _ptr(dl, _u32p), _ptr(ln, _u32p), _ptr(w, _f32p)))
        return [(int(f[i]), blobs[int(f[i])][int(off[i]):int(off[i]) + int(dl[i])], int(ln[i]), float(w[i])) for i in range(n)]

    def upload_fasta_groups(self, kind, texts, file_group, min_length=0, max_length=1 << 40, num_pad=1, ignore=()):
        """append_fasta_group: the files of a group form one sequence -> [(first file, first kept defline, length, weight)] per sequence"""
        blobs = [bytes(t) for t in texts]
        arr = (ctypes.c_char_p * max(1, len(blobs)))(*blobs)
        nb = np.array([len(b) for b in blobs], np.uint64)
        fg = np.ascontiguousarray(file_group, dtype=np.uint32)
        assert len(fg) == len(blobs)
        ig = [x.encode() for x in ignore]
        iga = (ctypes.c_char_p * max(1, len(ig)))(*ig)
        n = np.zeros(1, np.uint32)
        self._ck(self.lib.pcramp_gpu_upload_fasta_groups(self.h, kind, len(blobs), arr, _ptr(nb, _u64p), _ptr(fg, _u32p), int(min_length),
                                                         int(max_length), int(num_pad), len(ig), iga, _ptr(n, _u32p)))
        n = int(n[0])
        self.n_seq[kind] = n
        f, off, dl, ln, w = np.zeros(n, np.uint32), np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.float32)
        self._ck(self.lib.pcramp_gpu_fasta_records(self.h, kind, _ptr(f, _u32p), _ptr(off, _u64p), _ptr(dl, _u32p), _ptr(ln, _u32p), _ptr(w, _f32p)))
        return [(int(f[i]), blobs[int(f[i])][int(off[i]):int(off[i]) + int(dl[i])], int(ln[i]), float(w[i])) for i in range(n)]

    def fasta_timing(self, kind):
        a, b, t, n = np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.uint64), np.zeros(1, np.uint64)
        self._ck(self.lib.pcramp_gpu_fasta_timing(self.h, kind, _ptr(a, _f32p), _ptr(b, _f32p), _ptr(t, _u64p), _ptr(n, _u64p)))
        return {"ms_count": float(a[0]), "ms_pack": float(b[0]), "text_bytes": int(t[0]), "n_bases": int(n[0])}

    def set_weights(self, kind, weight):
        w = np.ascontiguousarray(weight, dtype=np.float32)
        self._ck(self.lib.pcramp_gpu_set_weights(self.h, kind, _ptr(w, _f32p)))

    def sequences_copy(self, kind):
        """the collection as the reference stores it -> (byte_off uint64[n], length uint32[n], nibbles uint8[total])"""
        n, tot = np.zeros(1, np.uint32), np.zeros(1, np.uint64)
        self._ck(self.lib.pcramp_gpu_sequences_copy(self.h, kind, _ptr(n, _u32p), _ptr(tot, _u64p), None, None, None))
        off, ln, nib = np.zeros(int(n[0]), np.uint64), np.zeros(int(n[0]), np.uint32), np.zeros(max(1, int(tot[0])), np.uint8)
        self._ck(self.lib.pcramp_gpu_sequences_copy(self.h, kind, None, None, _ptr(off, _u64p), _ptr(ln, _u32p), _ptr(nib, _u8p)))
        return off, ln, nib[:int(tot[0])]

    # ---- the multiplex terms of optimize() ---------------------------------------------------
    def multiplex_keys(self, pack_max_degen=256, min_oligo_length=18):
        """keys() of the whole-sequence pack of the MULTIPLEX collection (main.cpp:989-1003) -> (n_keys, 2) uint64"""
        n = np.zeros(1, np.uint64)
        self._ck(self.lib.pcramp_gpu_multiplex_keys(self.h, int(pack_max_degen), int(min_oligo_length), _ptr(n, _u64p)))
        words = np.zeros((int(n[0]), 2), np.uint64)
        self._ck(self.lib.pcramp_gpu_multiplex_keys_copy(self.h, _ptr(words, _u64p)))
        return words

    def set_pool(self, pool_f, pool_r):
        pf, pr = _words(pool_f), _words(pool_r)
        self._ck(self.lib.pcramp_gpu_set_pool(self.h, _ptr(pf, _u64p), _ptr(pr, _u64p), len(pf)))

    def multiplex_coverage(self, base_f, base_r, var_f, var_r, threshold, use_taq_mama=False):
        bf, br, vf, vr = _words(base_f), _words(base_r), _words(var_f), _words(var_r)
        cov = np.zeros(len(bf), np.float32)
        self._ck(self.lib.pcramp_gpu_multiplex_coverage(self.h, _ptr(bf, _u64p), _ptr(br, _u64p), _ptr(vf, _u64p), _ptr(vr, _u64p), len(bf),
                                                        float(threshold), int(use_taq_mama), _ptr(cov, _f32p)))
        return cov

    def oligo_overlap(self, f, r):
        f, r = _words(f), _words(r)
        ov = np.zeros(len(f), np.float32)
        self._ck(self.lib.pcramp_gpu_oligo_overlap(self.h, _ptr(f, _u64p), _ptr(r, _u64p), len(f), _ptr(ov, _f32p)))
        return ov

    # ---- resident variants --------------------------------------------------------------------
    def stage_pairs(self, f, r):
        f, r = _words(f), _words(r)
        self._ck(self.lib.pcramp_gpu_stage_pairs(self.h, _ptr(f, _u64p), _ptr(r, _u64p), len(f)))
        self.n_pairs = len(f)

    def set_batch(self, first, count):
        self._ck(self.lib.pcramp_gpu_set_batch(self.h, int(first), int(count)))
        self.n_pairs = int(count)

    def select_words_staged(self, kind, threshold, optimize_5=False, optimize_3=False, pack_max_degen=256, pack_min_gc=0.0, pack_max_gc=1.0,
                            min_oligo_length=18, want_keys=True, want_entries=True):
        ne, nk = ctypes.c_uint64(), ctypes.c_uint64()
        self._ck(self.lib.pcramp_gpu_select_words_staged(self.h, kind, int(optimize_5), int(optimize_3), float(threshold), int(pack_max_degen),
                                                         float(pack_min_gc), float(pack_max_gc), int(min_oligo_length),
                                                         ctypes.byref(ne) if want_entries else None, ctypes.byref(nk) if want_keys else None))
        return (ne.value if want_entries else None), (nk.value if want_keys else None)

    def score_pairs_staged(self, kind, search_threshold, detect_threshold, amplicon_min=80, amplicon_max=200, use_taq_mama=False):
        self._ck(self.lib.pcramp_gpu_score_pairs_staged(self.h, kind, float(search_threshold), float(detect_threshold), int(amplicon_min),
                                                        int(amplicon_max), int(use_taq_mama)))

    def fetch_results(self, kind):
        nw = (self.n_seq[kind] + 31) // 32
        cov = np.zeros(self.n_pairs, np.float32)
        bits = np.zeros((self.n_pairs, nw), np.uint32)
        self._ck(self.lib.pcramp_gpu_fetch_results(self.h, _ptr(cov, _f32p), _ptr(bits, _u32p)))
        return cov, bits

    def device_pointers(self):
        """(coverage, bitsets, pass-1 bitsets) device addresses of the last staged scoring call"""
        return (self.lib.pcramp_gpu_device_coverage(self.h), self.lib.pcramp_gpu_device_bitsets(self.h),
                self.lib.pcramp_gpu_device_bitsets_pass1(self.h))

    def merge_shards(self, d_any, d_pass1, shard_nseq, n_pairs, d_out_bits, d_out_cov, weight_all=None):
        shard_nseq = np.ascontiguousarray(shard_nseq, dtype=np.uint32)
        if weight_all is not None:
            weight_all = np.ascontiguousarray(weight_all, dtype=np.float32)
        self._ck(self.lib.pcramp_gpu_merge_shards(self.h, d_any, d_pass1, len(shard_nseq), _ptr(shard_nseq, _u32p), _ptr(weight_all, _f32p),
                                                  int(n_pairs), d_out_bits, d_out_cov))

    # ---- multi-GPU exchange over peer memory (xchg.cuh) ----------------------------------------
    def exchange_create(self, rank, world, shard_nseq, max_pairs, weight_all=None):
        sn = np.ascontiguousarray(shard_nseq, dtype=np.uint32)
        w = None if weight_all is None else np.ascontiguousarray(weight_all, dtype=np.float32)
        self._ck(self.lib.pcramp_gpu_exchange_create(self.h, int(rank), int(world), _ptr(sn, _u32p), int(max_pairs), _ptr(w, _f32p)))
        self._xchg_world = int(world)

    def exchange_buffer(self):
        return self.lib.pcramp_gpu_exchange_buffer(self.h)

    def exchange_ipc_handle(self):
        buf = ctypes.create_string_buffer(64)
        self._ck(self.lib.pcramp_gpu_exchange_ipc_handle(self.h, buf))
        return buf.raw

    def exchange_connect_pointers(self, pointers):
        arr = (ctypes.c_void_p * len(pointers))(*[int(p) for p in pointers])
        self._ck(self.lib.pcramp_gpu_exchange_connect(self.h, arr, 0))

    def exchange_connect_ipc(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self._xchg_world
        self._ck(self.lib.pcramp_gpu_exchange_connect(self.h, ctypes.create_string_buffer(blob, len(blob)), 1))

    def exchange_step(self, kind):
        self._ck(self.lib.pcramp_gpu_exchange_step(self.h, kind))

    def exchange_pointers(self):
        """(coverage, bitsets, words per row) of the merged result of the last step"""
        return (self.lib.pcramp_gpu_exchange_coverage(self.h), self.lib.pcramp_gpu_exchange_bitsets(self.h),
                int(self.lib.pcramp_gpu_exchange_words(self.h)))

    def exchange_fetch(self, n_pairs=None):
        """host copies of the merged result of the last step; the row count is the library's (n_pairs, when given, must agree)"""
        words = int(self.lib.pcramp_gpu_exchange_words(self.h))
        rows = int(self.lib.pcramp_gpu_exchange_pairs(self.h))
        if n_pairs is not None and int(n_pairs) != rows:
            raise GpuError("exchange_fetch: the last step held %d pairs, not %d" % (rows, int(n_pairs)))
        cov = np.zeros(rows, np.float32)
        bits = np.zeros((rows, max(1, words)), np.uint32)
        self._ck(self.lib.pcramp_gpu_exchange_fetch(self.h, _ptr(cov, _f32p), _ptr(bits, _u32p)))
        return cov, bits

    def exchange_set_timeout_ms(self, ms):
        self._ck(self.lib.pcramp_gpu_exchange_set_timeout_ms(self.h, int(ms)))

    def exchange_status(self):
        """-> bit mask of the ranks that never arrived (0 = healthy); synchronises"""
        m = ctypes.c_uint32()
        self._ck(self.lib.pcramp_gpu_exchange_status(self.h, ctypes.byref(m)))
        return int(m.value)

    def exchange_destroy(self):
        self._ck(self.lib.pcramp_gpu_exchange_destroy(self.h))

    def reduce_best(self, target_coverage, background_coverage, oligo_overlap, degeneracy, global_trial):
        """pcramp_gpu_reduce_best -> (owner rank, target, background, overlap, degeneracy, global trial or -1)"""
        o, t, b, v = ctypes.c_uint32(), ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        d, k = ctypes.c_double(), ctypes.c_int64()
        self._ck(self.lib.pcramp_gpu_reduce_best(self.h, float(target_coverage), float(background_coverage), float(oligo_overlap), float(degeneracy),
                                                 int(global_trial), ctypes.byref(o), ctypes.byref(t), ctypes.byref(b), ctypes.byref(v),
                                                 ctypes.byref(d), ctypes.byref(k)))
        return int(o.value), np.float32(t.value), np.float32(b.value), np.float32(v.value), float(d.value), int(k.value)

    def _thermo_args(self, op, seq_a, seq_b, strand_a, strand_b):
        a = seq_a if isinstance(seq_a, np.ndarray) else pack_strings(seq_a)
        n = len(a)
        b = None
        if seq_b is not None:
            b = seq_b if isinstance(seq_b, np.ndarray) else pack_strings(seq_b, a.shape[1])
        sa = None if strand_a is None else np.ascontiguousarray(np.broadcast_to(np.asarray(strand_a, dtype=np.float32), (n,)))
        sb = None if strand_b is None else np.ascontiguousarray(np.broadcast_to(np.asarray(strand_b, dtype=np.float32), (n,)))
        return n, a, b, sa, sb

    def thermo_batch(self, op, seq_a, seq_b=None, salt=0.05, strand_a=9e-7, strand_b=None, out=None):
        """-> (tm, dH, dS, dG_dp) float32 arrays; seq_* are lists of str or (n, stride) uint8 arrays; out = four float32 arrays of n
        to fill (page-locked arrays, like page-locked seq_* arrays, move at link speed)"""
        n, a, b, sa, sb = self._thermo_args(op, seq_a, seq_b, strand_a, strand_b)
        if out is None:
            out = [np.zeros(n, np.float32) for _ in range(4)]
        else:
            assert len(out) == 4 and all(o.dtype == np.float32 and o.size == n and o.flags.c_contiguous for o in out)
        self._ck(self.lib.pcramp_gpu_thermo_batch(self.h, int(op), n, a.ctypes.data_as(ctypes.c_char_p),
                                                  None if b is None else b.ctypes.data_as(ctypes.c_char_p), a.shape[1], float(salt),
                                                  _ptr(sa, _f32p), _ptr(sb, _f32p), *[_ptr(o, _f32p) for o in out]))
        return tuple(out)

    def thermo_words(self, op, words_a, words_b=None, salt=0.05, strand_a=9e-7, strand_b=None, out=None):
        """pcramp_gpu_thermo_words: the batch of thermo_batch with the oligos as (n, 2) uint64 words"""
        a = _words(words_a)
        n = len(a)
        b = None if words_b is None else _words(words_b)
        sa = None if strand_a is None else np.ascontiguousarray(np.broadcast_to(np.asarray(strand_a, dtype=np.float32), (n,)))
        sb = None if strand_b is None else np.ascontiguousarray(np.broadcast_to(np.asarray(strand_b, dtype=np.float32), (n,)))
        if out is None:
            out = [np.zeros(n, np.float32) for _ in range(4)]
        self._ck(self.lib.pcramp_gpu_thermo_words(self.h, int(op), n, _ptr(a, _u64p), _ptr(b, _u64p), float(salt), _ptr(sa, _f32p), _ptr(sb, _f32p),
                                                  *[_ptr(o, _f32p) for o in out]))
        return tuple(out)

    def thermo_stage(self, op, seq_a, seq_b=None, salt=0.05, strand_a=9e-7, strand_b=None):
        n, a, b, sa, sb = self._thermo_args(op, seq_a, seq_b, strand_a, strand_b)
        self._ck(self.lib.pcramp_gpu_thermo_stage(self.h, int(op), n, a.ctypes.data_as(ctypes.c_char_p),
                                                  None if b is None else b.ctypes.data_as(ctypes.c_char_p), a.shape[1], float(salt),
                                                  _ptr(sa, _f32p), _ptr(sb, _f32p)))
        self._thermo_n = n

    def thermo_run_staged(self):
        self._ck(self.lib.pcramp_gpu_thermo_run_staged(self.h))

    def thermo_fetch(self):
        out = [np.zeros(self._thermo_n, np.float32) for _ in range(4)]
        self._ck(self.lib.pcramp_gpu_thermo_fetch(self.h, *[_ptr(o, _f32p) for o in out]))
        return tuple(out)

    def is_valid(self, words, salt=0.05, primer_strand=9e-7, tm_range=(50.0, 75.0), max_hairpin=40.0, max_dimer=40.0,
                 check_homo_dimer=True, fast_alignment=False):
        words = _words(words)
        valid = np.zeros(len(words), np.uint8)
        self._ck(self.lib.pcramp_gpu_is_valid(self.h, _ptr(words, _u64p), len(words), float(salt), float(primer_strand), float(tm_range[0]),
                                              float(tm_range[1]), float(max_hairpin), float(max_dimer), int(check_homo_dimer),
                                              int(fast_alignment), _ptr(valid, _u8p)))
        return valid

    def max_dimer_tm(self, f, r, salt=0.05, primer_strand=9e-7, fast_alignment=False):
        f, r = _words(f), _words(r)
        tm = np.zeros(len(f), np.float32)
        self._ck(self.lib.pcramp_gpu_max_dimer_tm(self.h, _ptr(f, _u64p), _ptr(r, _u64p), len(f), float(salt), float(primer_strand),
                                                  int(fast_alignment), _ptr(tm, _f32p)))
        return tm

    def multiplex_compatible(self, f, r, pool_f, pool_r, salt=0.05, primer_strand=9e-7, max_dimer=40.0, fast_alignment=False):
        f, r, pool_f, pool_r = _words(f), _words(r), _words(pool_f), _words(pool_r)
        ok = np.zeros(len(f), np.uint8)
        self._ck(self.lib.pcramp_gpu_multiplex_compatible(self.h, _ptr(f, _u64p), _ptr(r, _u64p), len(f), _ptr(pool_f, _u64p),
                                                          _ptr(pool_r, _u64p), len(pool_f), float(salt), float(primer_strand),
                                                          float(max_dimer), int(fast_alignment), _ptr(ok, _u8p)))
        return ok

    # ---- K4: Smith-Waterman and the background tests ------------------------------------------------
    def sw_batch(self, query, target, out=None):
        """-> (n, 6) int32 {score, q_start, q_stop, t_start, t_stop, (last_two.first << 4) | last_two.second};
        out = (five int32 arrays of n, one uint8 array of (n, 2)): the library copies straight into them (page-locked arrays
        move at link speed) and they are returned as they are"""
        q, t = _words(query), _words(target)
        n = len(q)
        if out is not None:
            cols, l2 = list(out[0]), out[1]
            assert all(c.dtype == np.int32 and c.size == n and c.flags.c_contiguous for c in cols) and l2.dtype == np.uint8 and l2.size == 2 * n
        else:
            cols = [np.zeros(n, np.int32) for _ in range(5)]
            l2 = np.zeros((n, 2), np.uint8)
        self._ck(self.lib.pcramp_gpu_sw_batch(self.h, n, _ptr(q, _u64p), _ptr(t, _u64p), *[_ptr(c, _i32p) for c in cols], _ptr(l2, _u8p)))
        if out is not None:
            return out
        return np.stack(cols + [(l2[:, 0].astype(np.int32) << 4) | l2[:, 1]], 1)

    def sw_timing(self):
        ms = np.zeros(1, np.float32)
        self._ck(self.lib.pcramp_gpu_sw_timing(self.h, _ptr(ms, _f32p)))
        return float(ms[0])

    def background_match(self, kind, f, r, search_threshold, detect_threshold, amplicon_min=0, amplicon_max=2000, use_taq_mama=False):
        """-> (bitsets (n_pairs, words) uint32, number of candidate amplicons)"""
        f, r = _words(f), _words(r)
        nw = (self.n_seq[kind] + 31) // 32
        bits = np.zeros((len(f), nw), np.uint32)
        n_amp = ctypes.c_uint64()
        self._ck(self.lib.pcramp_gpu_background_match(self.h, kind, _ptr(f, _u64p), _ptr(r, _u64p), len(f), float(search_threshold),
                                                      float(detect_threshold), int(amplicon_min), int(amplicon_max), int(use_taq_mama),
                                                      _ptr(bits, _u32p), ctypes.byref(n_amp)))
        return bits, n_amp.value

    def multiplex_background_match(self, kind, f, r, threshold, use_taq_mama=False):
        f, r = _words(f), _words(r)
        nw = (self.n_seq[kind] + 31) // 32
        bits = np.zeros((len(f), nw), np.uint32)
        self._ck(self.lib.pcramp_gpu_multiplex_background_match(self.h, kind, _ptr(f, _u64p), _ptr(r, _u64p), len(f), float(threshold),
                                                                int(use_taq_mama), _ptr(bits, _u32p)))
        return bits

    def random_assays(self, kind, seeds, trials_per_stream, options=None):
        """PCR::random_assay, one GPU thread per seed stream -> (f, r (n_trials, 2) uint64, seeds after the run, attempts per trial)"""
        seeds = np.array(seeds, dtype=np.uint32).reshape(-1)
        per = np.ascontiguousarray(trials_per_stream, dtype=np.uint32).reshape(-1)
        assert len(seeds) == len(per)
        n = int(per.sum())
        f, r = np.zeros((n, 2), np.uint64), np.zeros((n, 2), np.uint64)
        att = np.zeros(max(1, n), np.uint32)
        opt = options if options is not None else RandomAssayOptions()
        self._ck(self.lib.pcramp_gpu_random_assays(self.h, kind, len(seeds), _ptr(seeds, _u32p), _ptr(per, _u32p), ctypes.byref(opt),
                                                   _ptr(f, _u64p), _ptr(r, _u64p), _ptr(att, _u32p)))
        return f, r, seeds, att[:n]

    def thermo_stats(self):
        s = ThermoStats()
        self._ck(self.lib.pcramp_gpu_get_thermo_stats(self.h, ctypes.byref(s)))
        return s.as_dict()

    def measure_int_peak(self):
        v = ctypes.c_double()
        self._ck(self.lib.pcramp_gpu_measure_int_peak(self.h, ctypes.byref(v)))
        return v.value

    def measure_int32_peak(self):
        """INT32 operations/s of the DP kernels' instruction mix, issue-bound (the K3 / K4 roofline denominator)"""
        v = ctypes.c_double()
        self._ck(self.lib.pcramp_gpu_measure_int32_peak(self.h, ctypes.byref(v)))
        return v.value

    def set_option(self, name, value):
        self._ck(self.lib.pcramp_gpu_set_option(self.h, name.encode(), int(value)))

    def stats(self):
        s = Stats()
        self._ck(self.lib.pcramp_gpu_get_stats(self.h, ctypes.byref(s)))
        return s.as_dict()


class DesignLoop:
    """pcramp_gpu_design_*: the body of pcramp's `while(true)` design loop (main.cpp:471-1130), one iteration per call"""

    def __init__(self, gpu, seed, **options):
        self.g = gpu
        self.opt = DesignOptions()
        gpu.lib.pcramp_gpu_design_default_options(ctypes.byref(self.opt))
        for k, v in options.items():
            if not hasattr(self.opt, k):
                raise KeyError(k)
            setattr(self.opt, k, v)
        self.h = ctypes.c_void_p()
        gpu._ck(gpu.lib.pcramp_gpu_design_create(gpu.h, ctypes.byref(self.opt), int(seed), ctypes.byref(self.h)))

    def iteration(self):
        res = DesignResult()
        if self.g.lib.pcramp_gpu_design_iteration(self.h, ctypes.byref(res)):
            raise GpuError(self.g.lib.pcramp_gpu_design_last_error(self.h).decode())
        return res

    def matches(self, n_target, n_background=0):
        t = np.zeros(max(1, (n_target + 31) // 32), np.uint32)
        b = np.zeros(max(1, (n_background + 31) // 32), np.uint32)
        self.g.lib.pcramp_gpu_design_matches(self.h, _ptr(t, _u32p), _ptr(b, _u32p))
        return unpack_bits(t[None, :], n_target)[0], unpack_bits(b[None, :], n_background)[0]

    def close(self):
        if self.h:
            self.g.lib.pcramp_gpu_design_destroy(self.h)
            self.h = None


def pack_strings(strs, stride=THERMO_STRIDE):
    """list of str -> (n, stride) uint8, NUL padded (the fixed-stride string layout of pcramp_gpu_thermo_batch)"""
    buf = np.zeros((len(strs), stride), dtype=np.uint8)
    for i, s in enumerate(strs):
        b = s.encode() if isinstance(s, str) else bytes(s)
        if len(b) >= stride:
            raise ValueError("sequence longer than %d bases" % (stride - 1))
        buf[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    return buf


def unpack_bits(bits, n_seq):
    """(n_pairs, words) uint32, LSB-first -> (n_pairs, n_seq) uint8."""
    b = np.unpackbits(np.ascontiguousarray(bits).view(np.uint8), axis=1, bitorder="little")
    return b[:, :n_seq]
