"""ctypes binding of libnrrt.so (include/nrrt.h).  There is no CPU fallback: a missing library is a hard error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnrrt.so")

NRRT_SOLVED, NRRT_MAX_ITER_BEST_EFFORT, NRRT_NO_NODE_CONNECTED, NRRT_SAMPLER_STUCK = 0, 1, 2, 3


class NrrtError(RuntimeError):
    pass


class NrrtPlanParams(C.Structure):
    _fields_ = [("dim", C.c_int32), ("shape", C.c_int32 * 3), ("max_iterations", C.c_int32), ("node_stride", C.c_int32),
                ("max_reject", C.c_int32), ("path_cap", C.c_int32), ("goal_threshold", C.c_double),
                ("neural_bias", C.c_double), ("seed", C.c_uint64), ("draw_start", C.c_uint64), ("sampler_mode", C.c_int32)]


class NrrtPlanOut(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("cost", C.c_void_p), ("parent", C.c_void_p), ("n_nodes", C.c_void_p),
                ("iterations", C.c_void_p), ("status", C.c_void_p), ("goal_node", C.c_void_p), ("best_node", C.c_void_p),
                ("best_dist", C.c_void_p), ("path_n", C.c_void_p), ("path_idx", C.c_void_p), ("path_len", C.c_void_p),
                ("trace", C.c_void_p), ("trace_n", C.c_void_p), ("trace_cap", C.c_int32)]


class NrrtConvLayer(C.Structure):
    _fields_ = [("dims", C.c_int32), ("kind", C.c_int32), ("n", C.c_int32), ("in_d", C.c_int32), ("in_h", C.c_int32),
                ("in_w", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("in_", C.c_void_p), ("in_cstride", C.c_int64),
                ("weight", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("out", C.c_void_p),
                ("out_cstride", C.c_int64), ("out_coffset", C.c_int64), ("res", C.c_void_p), ("res_cstride", C.c_int64),
                ("res_coffset", C.c_int64), ("relu", C.c_int32), ("post_scale", C.c_void_p), ("post_shift", C.c_void_p),
                ("poly", C.c_void_p), ("poly_cases", C.c_int32), ("head_weight", C.c_void_p), ("head_bias", C.c_float),
                ("logits", C.c_void_p), ("force_per_tap", C.c_int32), ("in2", C.c_void_p), ("in2_cstride", C.c_int64),
                ("cin2", C.c_int32), ("in_n", C.c_int32), ("in2_n", C.c_int32), ("in_index", C.c_void_p), ("in2_index", C.c_void_p),
                ("res_index", C.c_void_p), ("res_n", C.c_int32),
                ("no_poly_mma", C.c_int32), ("no_tma_store", C.c_int32),
                ("poly3", C.c_void_p), ("res_mma", C.c_int32), ("res_group", C.c_int32), ("tag", C.c_int32)]


class NrrtConvW(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("post_scale", C.c_void_p),
                ("post_shift", C.c_void_p), ("kind", C.c_int32), ("cin", C.c_int32), ("cin2", C.c_int32), ("cout", C.c_int32),
                ("relu", C.c_int32), ("tag", C.c_int32)]


class NrrtCnnWeights(C.Structure):
    _fields_ = [("dims", C.c_int32), ("stem_cin", C.c_int32), ("stem_cout", C.c_int32), ("stem_w", C.c_void_p),
                ("stem_b", C.c_void_p), ("stem2", NrrtConvW), ("enc", ((NrrtConvW * 3) * 2) * 4), ("coords_w", C.c_void_p * 4),
                ("coords_b", C.c_void_p * 4), ("up6", NrrtConvW), ("up7", NrrtConvW), ("up8", NrrtConvW), ("c6full", NrrtConvW),
                ("c6e", NrrtConvW), ("c7e", NrrtConvW), ("c8e", NrrtConvW), ("c60", NrrtConvW), ("c61", NrrtConvW),
                ("c70", NrrtConvW), ("c71", NrrtConvW), ("c80", NrrtConvW), ("c81", NrrtConvW), ("c7f", NrrtConvW),
                ("c8f", NrrtConvW), ("bc7", C.c_void_p), ("bc8", C.c_void_p), ("mt_up6", C.c_void_p), ("mt_c6", C.c_void_p),
                ("mt_c7", C.c_void_p), ("mt_c8", C.c_void_p), ("w6taps", C.c_void_p), ("b6", C.c_void_p), ("mt3_c6", C.c_void_p),
                ("mt3_c7", C.c_void_p), ("mt3_c8", C.c_void_p), ("head_w", C.c_void_p), ("head_bias", C.c_float)]


# NrrtLayerTag (include/nrrt.h)
TAG_STEM2, TAG_ENC, TAG_UP6, TAG_UP7, TAG_UP8 = 1, 10, 40, 41, 42
TAG_C6FULL, TAG_C6E, TAG_C7E, TAG_C8E = 45, 46, 47, 48
TAG_C60, TAG_C61, TAG_C70, TAG_C71, TAG_C80, TAG_C81, TAG_C7F, TAG_C8F = 50, 51, 52, 53, 54, 55, 60, 61


def tag_name(tag: int) -> str:
    names = {TAG_STEM2: "stem2", TAG_UP6: "up6", TAG_UP7: "up7", TAG_UP8: "up8", TAG_C6FULL: "c6full", TAG_C6E: "c6e",
             TAG_C7E: "c7e", TAG_C8E: "c8e", TAG_C60: "c6.0", TAG_C61: "c6.1", TAG_C70: "c7.0", TAG_C71: "c7.1", TAG_C80: "c8.0",
             TAG_C81: "c8.1", TAG_C7F: "c7f", TAG_C8F: "c8f"}
    if TAG_ENC <= tag < TAG_ENC + 24:
        k = tag - TAG_ENC
        return f"enc{k // 6 + 2}.{(k // 3) % 2}.{('c1', 'c2', 'ds')[k % 3]}"
    return names.get(tag, f"tag{tag}")


_SIGNATURES = {
    "nrrt_version": (C.c_int, []),
    "nrrt_last_error": (C.c_char_p, []),
    "nrrt_words_per_map": (C.c_size_t, [C.c_int64]),
    "nrrt_pack_occupancy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "nrrt_cdf_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int64]),
    "nrrt_cdf_build": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrrt_draw_samples": (C.c_int, [C.POINTER(NrrtPlanParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_plan_batch": (C.c_int, [C.POINTER(NrrtPlanParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.POINTER(NrrtPlanOut), C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_plan_smem_bytes": (C.c_int, [C.POINTER(NrrtPlanParams), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "nrrt_collision_free": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "nrrt_tree_search": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "nrrt_conv_layer": (C.c_int, [C.POINTER(NrrtConvLayer), C.c_void_p]),
    "nrrt_conv_profile_start": (C.c_int, [C.c_int]),
    "nrrt_conv_profile_stop": (C.c_int, []),
    "nrrt_conv_profile_resume": (C.c_int, []),
    "nrrt_conv_profile_read": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "nrrt_conv_forward_workspace_bytes": (C.c_size_t, [C.POINTER(NrrtCnnWeights), C.c_int, C.c_int, C.c_int, C.c_int]),
    "nrrt_conv_forward_2d": (C.c_int, [C.POINTER(NrrtCnnWeights), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrrt_conv_forward_3d": (C.c_int, [C.POINTER(NrrtCnnWeights), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nrrt_stem_conv": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_stem_conv_maps": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_coords_branch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_coords_fill": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "nrrt_coords_poly": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                   C.c_void_p, C.c_void_p]),
    "nrrt_coords_poly_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_coords_pieces_3d": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_poly_relu_3d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nrrt_poly16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_poly_relu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nrrt_poly_fold": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_gather_channels": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64,
                                       C.c_void_p, C.c_int, C.c_void_p]),
    "nrrt_grid_paths": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_grid_paths_ex": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "nrrt_generate_maps": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_generate_task_candidates": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_select_tasks": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_path_masks_2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_mask_tables_host": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "nrrt_paths_to_volume": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_enlarge_paths_3d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_to_channel_major": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_conv_wgrad": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "nrrt_conv_wgrad_nhwc": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p]),
    "nrrt_relu_bwd_sum": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "nrrt_bn_train_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_bn_eval_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nrrt_bn_train_backward": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_remap_channels_last": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nrrt_bce_with_logits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_mask_targets": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "nrrt_head_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "nrrt_grad_check": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nrrt_loss_scale_update": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "nrrt_to_half": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "nrrt_pack_dgrad_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_pack_dgrad_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nrrt_stem_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_coords_fill_backward": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nrrt_coords_layer_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nrrt_head_1x1": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None
launches = 0  # kernels of libnrrt.so launched through the Python wrappers (bench.py reports it as gpu_launches)


def count_launch(n: int = 1):
    global launches
    launches += n


def load():
    """Load libnrrt.so; raise loudly when the CUDA extension has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NrrtError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a). mgr_sampling_cnn_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise NrrtError(f"libnrrt error {rc}: {load().nrrt_last_error().decode()}")
