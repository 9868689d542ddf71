"""ishara_b200 — B200-native (sm_100a) implementation of the Ishara landmark-encoder hot path.

Drop-in for the reference's ``get_model(...)`` model call, ``CTCLoss`` and greedy CTC decode
(nb:conv-hybrid-model c5-c8). All compute is hand-written CUDA behind the C ABI in
``include/ishara_b200.h``; importing this package loads ``ishara_b200/lib/libishara_b200.so`` and fails
loudly when it has not been built. There is no CPU or PyTorch fallback.
"""
from . import _lib
from ._dlpack import DeviceTensor, from_host
from ._lib import IsharaError
from .model import (CTCLoss, FALLBACK_IDS, IsharaModel, char_to_num, decode_batch_predictions, decode_ids,
                    decode_phrase, get_model, lr_schedule, lrfn, num_to_char, num_to_char_fn, pad_token, pad_token_idx,
                    tflite_postprocess)
from .preprocess import LandmarkPreprocessor, sel_cols
from .deploy import TFLiteModel, edit_distances, levenshtein_scores

_lib.load()  # fail at import time if the CUDA library is missing

__all__ = [
    "get_model", "IsharaModel", "CTCLoss", "decode_phrase", "decode_batch_predictions", "decode_ids",
    "num_to_char_fn", "tflite_postprocess", "char_to_num", "num_to_char", "pad_token", "pad_token_idx",
    "FALLBACK_IDS", "DeviceTensor", "from_host", "IsharaError", "LandmarkPreprocessor", "sel_cols",
    "TFLiteModel", "edit_distances", "levenshtein_scores", "lrfn", "lr_schedule",
]
