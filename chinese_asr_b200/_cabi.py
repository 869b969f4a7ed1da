"""ctypes binding of libasr_b200.so - the C ABI declared in include/asr_b200.h.

The library is built in-tree by build.sh / __graft_entry__.build().  There is NO fallback: if the
shared object is missing or a call fails, an exception is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libasr_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with ./build.sh (nvcc, sm_100a). "
        "chinese_asr_b200 has no CPU / PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)

PCM_F32, PCM_S16 = 0, 1          # enum asr_pcm_format


class AsrWeights(C.Structure):
    _fields_ = [("enc_w_ih", c_float_p * 8), ("enc_w_hh", c_float_p * 8),
                ("enc_b_ih", c_float_p * 8), ("enc_b_hh", c_float_p * 8),
                ("embedding", c_float_p), ("dec_w_ih", c_float_p), ("dec_w_hh", c_float_p),
                ("dec_b_ih", c_float_p), ("dec_b_hh", c_float_p), ("proj_w", c_float_p),
                ("proj_b", c_float_p), ("att_w_enc", c_float_p), ("att_b", c_float_p),
                ("att_w_hidden", c_float_p), ("att_v", c_float_p)]


class AsrFeatureConsts(C.Structure):
    _fields_ = [("mel_fb", c_float_p), ("window", c_float_p), ("taps", c_float_p),
                ("preemphasis", C.c_float)]


class AsrLmTables(C.Structure):
    _fields_ = [("uni_logp", c_float_p), ("uni_bo", c_float_p), ("bi_keys", c_int64_p),
                ("bi_vals", c_float_p), ("bi_cap", C.c_int64), ("tri_keys", c_int64_p),
                ("tri_vals", c_float_p), ("tri_cap", C.c_int64), ("vocab", C.c_int32),
                ("skip_id", C.c_int32), ("id_map", c_int32_p)]


# every symbol include/asr_b200.h declares, with its signature
SIGNATURES = {
    "asr_last_error": (C.c_char_p, []),
    "asr_version": (C.c_int, []),
    "asr_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(AsrWeights), C.POINTER(AsrFeatureConsts)]),
    "asr_destroy": (C.c_int, [C.c_void_p]),
    "asr_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int]),
    "asr_features": (C.c_int, [C.c_void_p, C.c_void_p, c_int64_p, C.c_int, C.c_void_p, c_int32_p,
                               C.c_int, C.c_void_p]),
    "asr_num_frames": (C.c_int, [C.c_int64]),
    "asr_features_pcm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_int64_p, C.c_int, C.c_void_p, c_int32_p,
                                   C.c_int, C.c_float, C.c_void_p]),
    "asr_cmvn": (C.c_int, [C.c_void_p, C.c_void_p, c_int32_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "asr_set_vocab": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, C.c_int]),
    "asr_wer": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, C.c_int, c_int32_p, c_int64_p, C.c_int, c_int32_p,
                          c_int32_p, C.c_void_p]),
    "asr_transcribe_pcm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, c_int64_p, C.c_int, C.c_int,
                                     C.c_int, C.c_float, C.c_int, C.c_double, C.c_double, c_int32_p, c_int32_p,
                                     c_float_p, C.c_void_p]),
    "asr_transcribe_device_pcm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, c_int64_p, C.c_int,
                                            C.c_int, C.c_int, C.c_float, C.c_int, C.c_double, C.c_double,
                                            c_int32_p, c_int32_p, c_float_p, C.c_void_p]),
    "asr_prefetch_pcm_fmt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_int64_p, C.c_int]),
    "asr_encode": (C.c_int, [C.c_void_p, C.c_void_p, c_int32_p, C.c_int, C.c_void_p]),
    "asr_export_encoder": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "asr_encode_layers": (C.c_int, [C.c_void_p, C.c_void_p, c_int32_p, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p]),
    "asr_decode_greedy": (C.c_int, [C.c_void_p, C.c_int, c_int32_p, c_int32_p, c_float_p, c_int32_p,
                                    c_int32_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "asr_set_lm": (C.c_int, [C.c_void_p, C.POINTER(AsrLmTables)]),
    "asr_lm_score": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, C.c_int, C.c_int, c_float_p,
                               C.c_void_p]),
    "asr_decode_beam": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double,
                                  C.c_double, c_int32_p, c_int32_p, c_float_p, c_int32_p, C.c_void_p]),
    "asr_beam_trace": (C.c_int, [C.c_void_p, c_float_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                 c_float_p]),
    "asr_check_guards": (C.c_int, [C.c_void_p]),
    "asr_decode_info": (C.c_int, [C.c_void_p, c_int32_p]),
    "asr_beam_nbest": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int32_p, c_int32_p, c_int32_p, c_float_p]),
    "asr_transcribe": (C.c_int, [C.c_void_p, C.c_void_p, c_int64_p, C.c_int, C.c_int, C.c_int,
                                 C.c_float, C.c_int, C.c_double, C.c_double, c_int32_p, c_int32_p,
                                 c_float_p, C.c_void_p]),
    "asr_transcribe_device": (C.c_int, [C.c_void_p, C.c_void_p, c_int64_p, C.c_int, C.c_int, C.c_int,
                                        C.c_float, C.c_int, C.c_double, C.c_double, c_int32_p,
                                        c_int32_p, c_float_p, C.c_void_p]),
    "asr_set_recurrence_chunks": (C.c_int, [C.c_void_p, C.c_int]),
    "asr_test_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_int, C.c_int, C.c_void_p]),
    "asr_prefetch_pcm": (C.c_int, [C.c_void_p, C.c_void_p, c_int64_p, C.c_int]),
    "asr_bench_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, C.c_void_p]),
    "asr_launch_count": (C.c_int64, [C.c_void_p, C.c_int]),
    "asr_convert_audio_length": (C.c_int64, [C.c_int64, C.c_int]),
    "asr_convert_audio": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_float,
                                    C.POINTER(C.c_int16), C.c_int64, c_int64_p, C.c_void_p]),
    "asr_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "asr_stage_times": (C.c_int, [C.c_void_p, c_float_p, C.c_int]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError if the library does not export it
    _fn.restype = _res
    _fn.argtypes = _args


class AsrError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib.asr_last_error()
        raise AsrError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def fptr(arr):
    """float32 numpy array -> float* (the array must stay alive during the call)."""
    return arr.ctypes.data_as(c_float_p)
