"""Explicit map between the Keras model that ``get_model`` builds (nb:conv-hybrid-model c5, c7) and the canonical
parameter names of this library -- for loading ``model.save_weights("model.h5")`` files (c9:10) and
for ``tools/dump_tf_reference.py``.

Why a table and not name matching: the variables inside the custom layers carry Keras AUTO names
(``squeezeformer_0/sequential/dense/kernel:0``, ``.../dense_1/...`` with process-global counters), which no suffix rule
recovers. What IS deterministic is the order: a Keras layer lists ``trainable_weights`` (then ``non_trainable_weights``)
own-variables first, then sub-layers in the order their attributes were assigned in ``__init__`` -- and that order can be
read off the notebook:

  SqueezeformerBlock c5:155-183   norm1, ffn1 (Sequential: Dense, Dropout, Dense), norm2, mha (qkv, proj), conv
                                  (ConvModule c5:135-143: norm, conv1, conv2 = CausalDWConv1D, conv3, se = fc1, fc2), norm3, ffn2
  ConformerBlock c5:311-319       ffn1 (FeedForwardModule.sequential c5:237-244), mha, conv (ConvolutionModule c5:249-284:
                                  pointwise_conv1, depthwise_conv, pointwise_conv2, batch_norm, layer_norm), ffn2,
                                  layer_norm1, layer_norm2
  Conv1DBlock c5:41-89            a FUNCTION: its layers are top-level and explicitly named ``<name>_expand_conv``,
                                  ``<name>_dwconv``, ``<name>_bn``, ``<name>_project_conv``; only ECA (c5:75) is unnamed
                                  (``eca``, ``eca_1``, ... in creation order = Conv1DBlock order)
  Dense / Conv1D: kernel, bias; LayerNormalization / BatchNormalization: gamma, beta (+ moving_mean, moving_variance,
  non-trainable)

Every assignment is shape-checked, so a Keras version that orders differently fails loudly instead of loading garbage.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

LayerSpec = Tuple[str, List[str], List[str]]  # (Keras layer name or 'ECA#<n>', trainable names, non-trainable names)


def _dense(base: str, bias: bool = True) -> List[str]:
    return [base + ".kernel"] + ([base + ".bias"] if bias else [])


def _ln(base: str) -> List[str]:
    return [base + ".gamma", base + ".beta"]


def _ffn(base: str) -> List[str]:
    return _dense(base + ".0") + _dense(base + ".2")


def keras_layer_table(num_conv_squeeze_blocks: int = 2, num_conv_conform_blocks: int = 2, num_conv_per_block: int = 3) -> List[LayerSpec]:
    """Weight-carrying layers of ``get_model`` in ``model.layers`` order with the canonical names of their
    ``trainable_weights`` and ``non_trainable_weights`` (in Keras' order)."""
    out: List[LayerSpec] = [("stem_conv", ["stem_conv.kernel"], []),
                            ("stem_bn", _ln("stem_bn"), ["stem_bn.moving_mean", "stem_bn.moving_variance"])]
    eca = 0

    def conv_blocks(tag: str, i: int):
        nonlocal eca
        for j in range(1, num_conv_per_block + 1):
            n = f"conv{tag}_{i}_{j}"
            out.append((n + "_expand_conv", _dense(n + "_expand_conv"), []))
            out.append((n + "_dwconv", [n + "_dwconv.depthwise_kernel"], []))
            out.append((n + "_bn", _ln(n + "_bn"), [n + "_bn.moving_mean", n + "_bn.moving_variance"]))
            out.append((f"ECA#{eca}", [n + "_eca.kernel"], []))
            eca += 1
            out.append((n + "_project_conv", _dense(n + "_project_conv"), []))

    for i in range(num_conv_squeeze_blocks):
        conv_blocks("squeeze", i)
        s = f"squeezeformer_{i}"
        tw = (_ln(s + ".norm1") + _ffn(s + ".ffn1") + _ln(s + ".norm2") + [s + ".mha.qkv.kernel", s + ".mha.proj.kernel"]
              + _ln(s + ".conv.norm") + _dense(s + ".conv.conv1") + [s + ".conv.conv2.depthwise_kernel"] + _dense(s + ".conv.conv3")
              + _dense(s + ".conv.se.fc1") + _dense(s + ".conv.se.fc2") + _ln(s + ".norm3") + _ffn(s + ".ffn2"))
        out.append((s, tw, []))
    for i in range(num_conv_conform_blocks):
        conv_blocks("conform", i)
        c = f"conformer_{i}"
        tw = (_ffn(c + ".ffn1") + [c + ".mha.qkv.kernel", c + ".mha.proj.kernel"]
              + _dense(c + ".conv.pointwise_conv1") + _dense(c + ".conv.depthwise_conv") + _dense(c + ".conv.pointwise_conv2")
              + _ln(c + ".conv.batch_norm") + _ln(c + ".conv.layer_norm") + _ffn(c + ".ffn2") + _ln(c + ".layer_norm1") + _ln(c + ".layer_norm2"))
        out.append((c, tw, [c + ".conv.batch_norm.moving_mean", c + ".conv.batch_norm.moving_variance"]))
    out.append(("top_conv", _dense("top_conv"), []))
    out.append(("classifier", _dense("classifier"), []))
    return out


def map_keras_layers(layers: Sequence, table: List[LayerSpec]) -> Dict[str, object]:
    """``layers`` = ``model.layers`` of the real Keras model. Returns canonical name -> Keras variable, checking counts.
    ECA layers are matched by class name in creation order; everything else by layer name."""
    by_name = {l.name: l for l in layers}
    ecas = [l for l in layers if type(l).__name__ == "ECA"]
    out: Dict[str, object] = {}
    for lname, tw, ntw in table:
        layer = ecas[int(lname[4:])] if lname.startswith("ECA#") else by_name[lname]
        got_t, got_n = list(layer.trainable_weights), list(layer.non_trainable_weights)
        if len(got_t) != len(tw) or len(got_n) != len(ntw):
            raise ValueError(f"layer {lname}: Keras has {len(got_t)}+{len(got_n)} variables, the table expects {len(tw)}+{len(ntw)}")
        for name, var in zip(tw + ntw, got_t + got_n):
            out[name] = var
    return out


def assign_to_keras(model, params: Dict[str, np.ndarray], table: List[LayerSpec]) -> int:
    """Copy canonical-named arrays into the Keras model (shape-checked). Returns the number of variables assigned."""
    m = map_keras_layers(model.layers, table)
    for name, var in m.items():
        a = np.asarray(params[name], np.float32)
        if tuple(var.shape) != a.shape:
            if int(np.prod(var.shape)) != a.size:
                raise ValueError(f"{name}: Keras shape {tuple(var.shape)} vs {a.shape}")
            a = a.reshape(var.shape)
        var.assign(a)
    missing = set(params) - set(m)
    if missing:
        raise ValueError(f"parameters without a Keras variable: {sorted(missing)[:5]} ...")
    return len(m)


def load_keras_h5(path: str, table: List[LayerSpec]) -> Dict[str, np.ndarray]:
    """Read a ``model.save_weights("model.h5")`` file (c9:10; Keras legacy HDF5 layout: one group per layer, attribute
    ``weight_names`` = ``layer.weights`` order = trainable then non-trainable) into canonical names. Needs ``h5py``, which is
    not part of the build image: the import error says so instead of pretending."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - h5py is absent here
        raise ImportError("load_keras_h5 needs h5py (not installed in this image); convert the file to .npz with "
                          "tools/dump_tf_reference.py on a machine that has TensorFlow") from e
    out: Dict[str, np.ndarray] = {}
    with h5py.File(path, "r") as f:
        root = f["model_weights"] if "model_weights" in f else f
        eca_groups = sorted((k for k in root.keys() if k == "eca" or k.startswith("eca_")), key=lambda k: int(k[4:] or 0) if k != "eca" else 0)
        for lname, tw, ntw in table:
            g = root[eca_groups[int(lname[4:])]] if lname.startswith("ECA#") else root[lname]
            names = [n.decode() if isinstance(n, bytes) else n for n in g.attrs["weight_names"]]
            if len(names) != len(tw) + len(ntw):
                raise ValueError(f"layer {lname}: file has {len(names)} arrays, the table expects {len(tw) + len(ntw)}")
            for cname, wname in zip(tw + ntw, names):
                out[cname] = np.asarray(g[wname], np.float32)
    return out
