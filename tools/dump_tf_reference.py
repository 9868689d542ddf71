"""Emit true golden vectors from the REAL reference (needs TensorFlow 2.12-era + tensorflow_addons and a checkout
of tanmayrainanda/ishara; neither exists in the build image, so this script is the documented upgrade path from
"parity unpinned" to "pinned" — DESIGN.md §3).

It does not contain reference code: it loads cells c5-c8 of `Test Notebooks/conv-hybrid-model.ipynb` from the
checkout you point it at, executes them, builds get_model() with the BASELINE kwargs, loads the seeded weights
that oracle.init_params generates (so both sides share weights), runs seeded inputs and writes

    tests/golden/tf_reference.npz        weights (Keras names), x, labels, logits, per-sequence CTC loss, decoded ids
    tests/golden/tf_reference_train.npz  one training-mode step with every Dropout layer switched off: mean CTC loss,
                                         gradients of all trainable variables, updated BatchNorm moving statistics

usage: python tools/dump_tf_reference.py /path/to/ishara [--frames 384] [--batch 4]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference_checkout")
    ap.add_argument("--frames", type=int, default=384)
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args()

    import tensorflow as tf
    import tensorflow_addons as tfa  # noqa: F401  (get_model's optimizer)

    from oracle import ishara_oracle as O

    nb = json.load(open(os.path.join(args.reference_checkout, "Test Notebooks", "conv-hybrid-model.ipynb")))
    code = [c for c in nb["cells"] if c["cell_type"] == "code"]
    ns = {"tf": tf, "tfa": tfa, "np": np, "INPUT_SHAPE": [args.frames, 276], "pad_token_idx": O.PAD_TOKEN_IDX,
          "char_to_num": dict(O.CHAR_TO_NUM), "num_to_char": dict(O.NUM_TO_CHAR)}
    for idx in (5, 6):
        exec("".join(code[idx]["source"]), ns)
    c7 = "".join(code[7]["source"]).split("tf.keras.backend.clear_session()")[0]   # the def only, not the demo call
    exec(c7, ns)
    exec("".join(code[8]["source"]), ns)

    cfg = O.Config(frames=args.frames)
    params = O.init_params(cfg, seed=42)
    model = ns["get_model"](num_conv_squeeze_blocks=cfg.num_conv_squeeze_blocks, num_conv_conform_blocks=cfg.num_conv_conform_blocks,
                            kernel_sizes=list(cfg.kernel_sizes), num_conv_per_block=cfg.num_conv_per_block, dropout_rate=cfg.dropout_rate)
    x = O.make_inputs(cfg, args.batch, seed=1234)
    y = O.make_labels(cfg, args.batch)
    model(x)  # build

    # explicit Keras-layer -> canonical-name table (ishara_b200/keras_names.py: order of trainable / non-trainable weights
    # per layer read off c5 / c7, every assignment shape-checked) instead of guessing from auto-generated variable names
    from ishara_b200.keras_names import assign_to_keras, keras_layer_table, map_keras_layers

    table = keras_layer_table(cfg.num_conv_squeeze_blocks, cfg.num_conv_conform_blocks, cfg.num_conv_per_block)
    assigned = assign_to_keras(model, params, table)
    assert assigned == len(params), (assigned, len(params))
    logits = model(x, training=False).numpy()
    nll = tf.nn.ctc_loss(labels=y, logits=logits, label_length=(y != 59).sum(-1).astype(np.int32),
                         logit_length=np.full(args.batch, args.frames, np.int32), blank_index=59,
                         logits_time_major=False).numpy()
    ids = [ns["decode_phrase"](l).numpy() for l in logits]
    out = os.path.join(ROOT, "tests", "golden", "tf_reference.npz")
    np.savez(out, x=x, labels=y, logits=logits, nll=nll, ids=np.array(ids, dtype=object), **{"w:" + k: v for k, v in params.items()})
    print("wrote", out)

    # ---- training-mode step. The head's Dropout(0.4) and ConformerBlock's attention dropout cannot
    # be disabled through get_model's kwargs, so every Dropout layer is made the identity for this dump: what is pinned is
    # BatchNormalization on batch statistics + its moving-average update, the loss and the gradients.
    tf.keras.layers.Dropout.call = lambda self, inputs, training=None: inputs
    name_of = {var.ref(): name for name, var in map_keras_layers(model.layers, table).items()}
    with tf.GradientTape() as tape:
        lg = model(x, training=True)
        loss = ns["CTCLoss"](tf.constant(y), lg)
    grads = tape.gradient(loss, model.trainable_variables)
    gout = {"g:" + name_of[v.ref()]: (np.zeros(v.shape, np.float32) if g is None else tf.convert_to_tensor(g).numpy().reshape(v.shape))
            for v, g in zip(model.trainable_variables, grads)}
    stats = {"s:" + name_of[v.ref()]: v.numpy() for v in model.variables
             if name_of[v.ref()].endswith(("moving_mean", "moving_variance"))}
    out_t = os.path.join(ROOT, "tests", "golden", "tf_reference_train.npz")
    np.savez(out_t, x=x, labels=y, loss=np.float32(loss.numpy()), logits_train=lg.numpy(), **gout, **stats,
             **{"w:" + k: v for k, v in params.items()})
    print("wrote", out_t)


if __name__ == "__main__":
    main()
