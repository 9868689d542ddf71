"""The deployed wrapper and the scorer of the reference.

``TFLiteModel`` mirrors ``TFLiteModel.__call__`` (nb:conv-hybrid-model c13:1-25): raw frames ``[N, len(SEL_COLS)]`` ->
pre_process00 + pre_process1 (GPU) -> model -> decode_phrase (GPU) -> short-prediction fallback -> ``one_hot(x, 59)``.
``levenshtein_scores`` is the evaluation loop's metric (c18:1-15): ``(len(target) - distance(pred, target)) / len(target)``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import numpy as np

from . import _lib
from .model import decode_phrase, num_to_char_fn, tflite_postprocess
from .preprocess import LandmarkPreprocessor


class TFLiteModel:
    def __init__(self, model, preprocessor: LandmarkPreprocessor):
        if preprocessor.frame_len != model.frames:
            raise ValueError(f"preprocessor frame_len {preprocessor.frame_len} != model frames {model.frames}")
        self.model, self.pre = model, preprocessor

    def __call__(self, inputs: np.ndarray, training: bool = False) -> Dict[str, np.ndarray]:
        x = self.pre([np.asarray(inputs, np.float32)])          # c13:9-15
        logits = self.model(x)                                  # c13:17
        ids = decode_phrase(logits[0])                          # c13:19
        return {"outputs": tflite_postprocess(ids)}             # c13:20-25 (fallback + one_hot)

    def predict_str(self, inputs: np.ndarray) -> str:
        """"".join(rev_character_map[s] for s in argmax(outputs)) — what the scorer consumes (c18:7)."""
        return "".join(num_to_char_fn(np.argmax(self(inputs)["outputs"], axis=1)))


def edit_distances(predictions: Sequence[str], targets: Sequence[str]) -> np.ndarray:
    """Levenshtein distance of every (prediction, target) pair (the `Levenshtein.distance` of c18:1), on UTF-8 bytes."""
    if len(predictions) != len(targets):
        raise ValueError("predictions and targets differ in length")
    n = len(targets)
    pb = [p.encode("utf-8") for p in predictions]
    tb = [t.encode("utf-8") for t in targets]
    P = (C.c_char_p * n)(*pb)
    T = (C.c_char_p * n)(*tb)
    pl = np.array([len(b) for b in pb], np.int32)
    tl = np.array([len(b) for b in tb], np.int32)
    out = np.zeros(n, np.int32)
    _lib.check(_lib.load().ishara_edit_distances(C.cast(P, C.c_void_p), pl.ctypes.data_as(C.c_void_p), C.cast(T, C.c_void_p),
                                                 tl.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)))
    return out


def levenshtein_scores(predictions: Sequence[str], targets: Sequence[str]) -> np.ndarray:
    """score_i = (len(target_i) - distance_i) / len(target_i)   (c18:9); the reference reports the mean (c18:15)."""
    d = edit_distances(predictions, targets).astype(np.float64)
    lens = np.array([len(t) for t in targets], np.float64)
    return (lens - d) / lens
