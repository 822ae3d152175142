"""ctypes binding of librelpose_b200.so (the C ABI declared in include/relpose_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or cannot be loaded every
compute entry point raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C rel_pose_b200/csrc`.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librelpose_b200.so")

_c_int = ctypes.c_int
_c_i64 = ctypes.c_int64
_c_f32 = ctypes.c_float
_c_size = ctypes.c_size_t
_ptr = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/relpose_b200.h one to one
_SIGNATURES = {
    "rp_last_error": (ctypes.c_char_p, []),
    "rp_version": (_c_int, []),
    "rp_device_arch": (_c_int, [_c_int]),
    "rp_copy_rows_h2d": (_c_int, [_ptr, _ptr, _c_i64, _c_int, _c_i64, _c_int, _c_int, _ptr]),
    "rp_preprocess_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_preprocess_u8": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_conv2d_workspace_bytes": (_c_size, [_c_int] * 9),
    "rp_conv2d_nhwc_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr] + [_c_int] * 10 + [_ptr, _c_size, _c_int, _ptr]),
    "rp_preprocess_nhwc4_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_preprocess_nhwc4_u8": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_preprocess_stem_windows_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_preprocess_stem_windows_u8": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_stem_weight_windows_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_preprocess_stem_compact_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_preprocess_stem_compact_u8": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_stem_compact_supported": (_c_int, [_c_int]),
    "rp_stem_pool_tc": (_c_int, [_ptr, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_maxpool3x3s2_nhwc_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_permute_conv_weight_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_bn_fold_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _c_f32, _ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_intrinsics_prepare_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_tokens_posembed_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_layernorm_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_f32, _c_int, _ptr]),
    "rp_linear_workspace_bytes": (_c_size, [_c_int, _c_int, _c_int]),
    "rp_linear_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_split_planes_bf16": (_c_int, [_ptr, _ptr, _c_i64, _c_int, _c_int, _ptr]),
    "rp_transpose_split_planes_bf16": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_layernorm_planes_bf16": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_f32, _c_int, _c_int, _ptr]),
    "rp_linear_tc": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_regressor_tail_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_linear_tc_splitk_workspace_bytes": (_c_size, [_c_int, _c_int, _c_int, _ptr]),
    "rp_linear_tc_splitk": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_conv3x3_halo_supported": (_c_int, [_c_int] * 8),
    "rp_conv3x3_halo_tc": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr] + [_c_int] * 9 + [_ptr]),
    "rp_ln_linear_tc": (_c_int, [_ptr, _ptr, _ptr, _c_f32, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_mlp_tc": (_c_int, [_ptr, _ptr, _ptr, _c_f32, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_ln_linear_tc_ex": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_f32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_f32,
                                    _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_mlp_tc_ex": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_f32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_f32,
                              _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_conv2d_tc": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr] + [_c_int] * 13 + [_ptr]),
    "rp_maxpool3x3s2_planes": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_self_attention_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_self_attention_tc": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_cross_attention_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_cross_attention_tc": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_posenc_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_essential_workspace_bytes": (_c_size, [_c_int]),
    "rp_posenc_ex_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _ptr]),
    "rp_essential_ex_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_essential_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_essential_tc_workspace_bytes": (_c_size, [_c_int, _c_int]),
    "rp_essential_tc": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_essential_ex_tc": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_em_project_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_normalize_pose_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_se3_mul_fwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_se3_mul_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_se3_inv_fwd_f32": (_c_int, [_ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_se3_inv_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_se3_log_fwd_f32": (_c_int, [_ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_se3_log_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_se3_exp_fwd_f32": (_c_int, [_ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_se3_exp_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_gemm_f32": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_f32, _ptr, _c_int, _ptr, _c_int, _c_f32, _ptr, _c_int, _c_int, _c_int, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_i64, _c_int, _ptr]),
    "rp_gelu_fwd_f32": (_c_int, [_ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_gelu_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_relu_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_mul_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_axpby_f32": (_c_int, [_c_f32, _ptr, _c_f32, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_add_bcast_rows_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _c_int, _c_int, _ptr]),
    "rp_sum_over_period_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_colsum_workspace_bytes": (_c_size, [_c_i64, _c_int]),
    "rp_colsum_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_layernorm_train_fwd_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_f32, _c_int, _ptr]),
    "rp_layernorm_bwd_workspace_bytes": (_c_size, [_c_int, _c_int]),
    "rp_layernorm_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_softmax_rows_fwd_f32": (_c_int, [_ptr, _ptr, _c_i64, _c_int, _c_f32, _c_int, _ptr]),
    "rp_softmax_rows_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_i64, _c_int, _c_f32, _c_int, _c_int, _ptr]),
    "rp_softmax_cols_fwd_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_f32, _c_int, _ptr]),
    "rp_softmax_cols_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_f32, _c_int, _c_int, _ptr]),
    "rp_bn_workspace_bytes": (_c_size, [_c_i64, _c_int]),
    "rp_bn_train_stats_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _c_f32, _c_i64, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_bn_apply_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _c_f32, _c_int, _c_int, _ptr]),
    "rp_bn_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c_f32, _ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_im2col_nhwc_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_im2col_t_planes_bf16": (_c_int, [_ptr, _ptr] + [_c_int] * 10 + [_ptr]),
    "rp_col2im_nhwc_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr]),
    "rp_maxpool3x3s2_bwd_workspace_bytes": (_c_size, [_c_int] * 4),
    "rp_maxpool3x3s2_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_normalize_pose_bwd_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_concat_vpos_f32": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_scatter_dv_f32": (_c_int, [_ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_self_attention_tc_lse": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _ptr]),
    "rp_attention_bwd_prep": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_attention_bwd_tc": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _ptr]),
    "rp_conv_dw_tc_supported": (_c_int, [_c_int] * 5),
    "rp_conv_dw_tc_workspace_bytes": (_c_size, [_c_int] * 10),
    "rp_conv_dw_tc": (_c_int, [_ptr, _ptr, _ptr] + [_c_int] * 9 + [_ptr, _c_size, _c_int, _ptr]),
    "rp_linear_dw_tc_supported": (_c_int, [_c_int] * 3),
    "rp_linear_dw_tc_workspace_bytes": (_c_size, [_c_int] * 4),
    "rp_linear_dw_tc": (_c_int, [_ptr, _ptr, _ptr] + [_c_int] * 3 + [_ptr, _c_size, _c_int, _ptr]),
    "rp_em_bwd_tc_workspace_bytes": (_c_size, [_c_int]),
    "rp_em_bwd_tc": (_c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr, _c_size, _c_int, _ptr]),
    "rp_grad_norm_multi": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr, _ptr, _c_int, _ptr]),
    "rp_adam_clip_step_multi": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr] + [ctypes.c_double] * 6 + [_c_int, _c_int, _ptr]),
    "rp_adam_clip_step_multi_dev": (_c_int, [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr, ctypes.c_double, _ptr] + [ctypes.c_double] * 4 +
                                    [_c_int, _ptr]),
    "rp_svd3_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
    "rp_essential_to_rt_f32": (_c_int, [_ptr, _ptr, _ptr, _ptr, _c_i64, _c_int, _ptr]),
}

_lock = threading.Lock()
_lib = None


class RelposeLibraryError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load (once) and return the ctypes handle.  Raises RelposeLibraryError if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RelposeLibraryError(
                f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
                "(run `make -C rel_pose_b200/csrc`); rel_pose_b200 has no CPU / PyTorch fallback.")
        try:
            handle = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise RelposeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as e:
                raise RelposeLibraryError(f"{LIB_PATH} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().rp_last_error()
        msg = msg.decode("utf-8", "replace") if msg else ""
        raise RelposeLibraryError(f"{what or 'relpose_b200'} failed (code {rc}): {msg}")
