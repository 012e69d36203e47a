"""ctypes binding of libtts_b200.so (the C ABI declared in include/tts_b200.h).

There is deliberately no fallback: if the shared library is missing or no CUDA device is
present, every product call raises.  PyTorch is used only to own device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libtts_b200.so")
TTS_MAX_LAYERS = 16
ABI_VERSION = 10

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u8p = C.POINTER(C.c_uint8)


class GemmEpilogue(C.Structure):
    _fields_ = [("alpha", C.c_float), ("scale", C.c_void_p), ("shift", C.c_void_p), ("bias", C.c_void_p),
                ("act", C.c_int32), ("residual", C.c_void_p), ("ldr", C.c_int32), ("row_len", C.c_void_p),
                ("rows_per_batch", C.c_int32), ("valid_rows", C.c_int32), ("out_rows_per_batch", C.c_int32),
                ("out_row_offset", C.c_int32), ("head_dim", C.c_int32), ("n_heads", C.c_int32),
                ("head_rows", C.c_int32), ("out_v", C.c_void_p)]


class GemmBf16(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32), ("a_rows", C.c_int64),
                ("B", C.c_void_p), ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
                ("C", C.c_void_p), ("ldc", C.c_int64), ("out_bf16", C.c_int32),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("taps", C.c_int32), ("split_k", C.c_int32),
                ("bias", C.c_void_p), ("act", C.c_int32), ("alpha", C.c_float),
                ("residual", C.c_void_p), ("ldr", C.c_int64),
                ("drop_p", C.c_float), ("seed", C.c_uint64), ("rng_stream", C.c_uint32),
                ("gate", C.c_void_p), ("ldg", C.c_int64), ("gate_scale", C.c_float),
                ("row_len", C.c_void_p), ("rows_per_batch", C.c_int32), ("valid_rows", C.c_int32),
                ("out_rows_per_batch", C.c_int32), ("out_row_offset", C.c_int32)]


class AttnTrain(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("ldq", C.c_int64), ("ldk", C.c_int64),
                ("ldv", C.c_int64), ("out", C.c_void_p), ("ldo", C.c_int64), ("lse", C.c_void_p),
                ("batch", C.c_int32), ("n_heads", C.c_int32), ("tq", C.c_int32), ("tk", C.c_int32),
                ("head_dim", C.c_int32), ("causal", C.c_int32), ("key_len", C.c_void_p),
                ("drop_p", C.c_float), ("seed", C.c_uint64), ("rng_stream", C.c_uint32),
                ("d_out", C.c_void_p), ("lddo", C.c_int64), ("delta", C.c_void_p),
                ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p), ("lddq", C.c_int64), ("lddk", C.c_int64),
                ("lddv", C.c_int64), ("dq_acc", C.c_void_p), ("keep_mask", C.c_void_p)]


class GriffinLim(C.Structure):
    _fields_ = [("mel", C.c_void_p), ("lengths", C.c_void_p), ("inv_basis_t", C.c_void_p), ("window", C.c_void_p),
                ("twiddle", C.c_void_p), ("batch", C.c_int32), ("frames_max", C.c_int32), ("min_frames", C.c_int32),
                ("n_mels", C.c_int32), ("n_fft", C.c_int32), ("hop_length", C.c_int32), ("win_length", C.c_int32),
                ("n_iter", C.c_int32), ("max_abs", C.c_float), ("max_db", C.c_float), ("ref_db", C.c_float),
                ("power", C.c_float), ("preemphasis", C.c_float), ("mag", C.c_void_p), ("frames", C.c_void_p),
                ("y", C.c_void_p), ("ldy", C.c_int64), ("wav", C.c_void_p), ("ldw", C.c_int64)]


class DecLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln_self_g", "ln_self_b", "w_qkv", "w_self_out", "ln_cross_g", "ln_cross_b", "w_cross_q", "w_cross_kv",
        "w_cross_out", "ln_ffn_g", "ln_ffn_b", "w_ffn_in", "w_ffn_out", "w_qkv_ln", "c_qkv_ln", "w_cross_q_ln",
        "c_cross_q_ln", "w_ffn_in_ln", "c_ffn_in_ln", "pk_qkv", "pk_self_out", "pk_cross_q", "pk_cross_out", "pk_ffn_in", "pk_ffn_out")]


class DecoderWeights(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("n_layers", "d_model", "n_heads", "d_ffn", "n_mels", "prenet_hidden")] +
                [(n, C.c_void_p) for n in ("prenet_w0", "prenet_b0", "prenet_w1", "prenet_b1", "prenet_w2", "pe_scale",
                                           "pe_table", "ln_out_g", "ln_out_b", "w_mel", "w_stop", "b_stop", "w_mel_ln",
                                           "w_stop_ln", "c_out_ln", "pk_pre0", "pk_pre1", "pk_pre2",
                                           "pk_final")] +
                [("pk_ksplit", C.c_int32), ("pk_reserved", C.c_int32)] +
                [("layer", DecLayerWeights * TTS_MAX_LAYERS)])


class DecodeState(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("batch", "mem_len", "t_max")] +
                [(n, C.c_void_p) for n in ("memory", "input_lengths", "self_k", "self_v", "cross_k", "cross_v",
                                           "lengths", "finished", "frames", "stop_logits", "align_self", "align_cross",
                                           "step_counter", "n_unfinished", "scratch")] +
                [("drop_p_prenet", C.c_float), ("drop_p_transformer", C.c_float), ("drop_seed", C.c_uint64)])


_EXPORTS = {
    "tts_abi_version": (C.c_int, []),
    "tts_last_error": (C.c_char_p, []),
    "tts_launch_count": (C.c_int64, []),
    "tts_launch_count_reset": (None, []),
    "tts_gemm_nt": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                              C.c_int32, C.c_int32, C.POINTER(GemmEpilogue), C.c_void_p]),
    "tts_gemm_use_tensor_cores": (C.c_int, [C.c_int]),
    "tts_gemm_bf16": (C.c_int, [C.POINTER(GemmBf16), C.c_void_p]),
    "tts_gemm_bf16_status": (C.c_int, []),
    "tts_griffin_lim": (C.c_int, [C.POINTER(GriffinLim), C.c_void_p]),
    "tts_attn_train_fwd": (C.c_int, [C.POINTER(AttnTrain), C.c_void_p]),
    "tts_attn_train_bwd": (C.c_int, [C.POINTER(AttnTrain), C.c_void_p]),
    "tts_attn_keep_words": (C.c_int32, [C.c_int32]),
    "tts_attn_tc_status": (C.c_int, []),
    "tts_attn_tc_trace": (C.c_int, [C.c_void_p]),
    "tts_ln_fwd_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_int32, C.c_void_p]),
    "tts_ln_bwd_train": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p]),
    "tts_ln_bwd_scratch_floats": (C.c_size_t, [C.c_int32]),
    "tts_dropout_cast": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int32, C.c_void_p]),
    "tts_multi_cast_bf16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "tts_multi_chunk_elems": (C.c_int32, []),
    "tts_embed_train_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p]),
    "tts_embed_train_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p]),
    "tts_shift_pe_train_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p]),
    "tts_shift_pe_train_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p]),
    "tts_colsum_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "tts_sum_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "tts_rowdot_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "tts_bn_scratch_floats": (C.c_size_t, [C.c_int32]),
    "tts_bn_train_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tts_bn_train_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tts_pad_cast_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tts_loss_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tts_sumsq_multi": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "tts_adam_multi": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, C.c_float, C.c_float, C.c_void_p]),
    "tts_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float,
                                C.c_void_p, C.c_int32, C.c_void_p]),
    "tts_embed_pe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                               C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tts_shift_pe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                               C.c_int32, C.c_void_p]),
    "tts_pad_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_void_p]),
    "tts_cond_embed": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "tts_attention": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                C.c_int32, C.c_void_p, C.c_void_p]),
    "tts_decode_scratch_bytes": (C.c_size_t, [C.POINTER(DecoderWeights), C.c_int32, C.c_int32, C.c_int32]),
    "tts_decode_begin": (C.c_int, [C.POINTER(DecoderWeights), C.POINTER(DecodeState), C.c_void_p]),
    "tts_decode_steps": (C.c_int, [C.POINTER(DecoderWeights), C.POINTER(DecodeState), C.c_int32, C.c_void_p,
                                   C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "tts_decode_profile": (C.c_int, [C.POINTER(DecoderWeights), C.POINTER(DecodeState), C.c_void_p, C.c_int32]),
}

_lib = None


def load(check_device: bool = True):
    """Load the shared library (once) and bind every symbol include/tts_b200.h declares."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("tts_b200: %s is missing - run `python __graft_entry__.py` (build()) first; "
                               "there is no CPU or eager fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _EXPORTS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if lib.tts_abi_version() != ABI_VERSION:
            raise RuntimeError("tts_b200: ABI mismatch (library %d, binding %d)" % (lib.tts_abi_version(), ABI_VERSION))
        _lib = lib
    if check_device and not torch.cuda.is_available():
        raise RuntimeError("tts_b200: no CUDA device; the mel path has no CPU fallback")
    return _lib


def exported_symbols():
    return sorted(_EXPORTS)


def check(rc: int, what: str):
    if rc != 0:
        msg = _lib.tts_last_error().decode("utf-8", "replace") if _lib is not None else "?"
        raise RuntimeError("tts_b200.%s failed (%d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous, CUDA - what every kernel expects."""
    if not t.is_cuda:
        raise RuntimeError("tts_b200: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
