"""Offline converter: the reference's TensorFlow SavedModel directory -> this package's npz checkpoint.

    python -m cubecobrarecommender_b200.scripts.convert_savedmodel ml_files/recommender ml_files/recommender_npz

The reference saves with ``autoencoder.save(dest, save_format='tf')`` (reference ``src/ml/train.py:112-115``): a
``saved_model.pb`` plus a TF2 object-graph checkpoint under ``variables/``.  This script needs TensorFlow ONLY to read
that checkpoint (``tf.train.load_checkpoint``) -- it runs wherever the model was trained, not on the serving box; this
package itself never imports TensorFlow.  Checkpoint keys of the subclassed model follow its attribute names
(reference ``src/ml/model.py:27-33, 58-64, 92-98``):

    encoder/encoded_1|encoded_2|encoded_3|bottleneck/kernel|bias
    decoder|decoder_for_reg/decoded_1|decoded_2|decoded_3|reconstruct/kernel|bias   (+ /.ATTRIBUTES/VARIABLE_VALUE)

and map onto the Keras layer names this package uses (``encoder_e1/kernel`` ... ``reg_reconstruction/bias``); Adam's
``m`` / ``v`` slots (``.OPTIMIZER_SLOT/optimizer/m|v``) and ``optimizer/iter`` are carried over when present, so a
converted checkpoint can also be resumed.  Kernels are stored ``(in, out)`` on both sides: no transposes.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ATTR_TO_LAYER = {
    ("encoder", "encoded_1"): "encoder_e1", ("encoder", "encoded_2"): "encoder_e2", ("encoder", "encoded_3"): "encoder_e3",
    ("encoder", "bottleneck"): "encoder_bottleneck",
    ("decoder", "decoded_1"): "main_d1", ("decoder", "decoded_2"): "main_d2", ("decoder", "decoded_3"): "main_d3",
    ("decoder", "reconstruct"): "main_reconstruction",
    ("decoder_for_reg", "decoded_1"): "reg_d1", ("decoder_for_reg", "decoded_2"): "reg_d2",
    ("decoder_for_reg", "decoded_3"): "reg_d3", ("decoder_for_reg", "reconstruct"): "reg_reconstruction",
}
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def map_checkpoint_keys(keys):
    """{checkpoint key: npz key} for the variables this package knows (pure string logic: unit-tested without TF)."""
    out = {}
    for key in keys:
        if not key.endswith(SUFFIX):
            continue
        parts = key[:-len(SUFFIX)].split("/")
        if len(parts) >= 3 and (parts[0], parts[1]) in ATTR_TO_LAYER and parts[2] in ("kernel", "bias"):
            name = ATTR_TO_LAYER[(parts[0], parts[1])] + "/" + parts[2]
            if len(parts) == 3:
                out[key] = name
            elif len(parts) == 6 and parts[3] == ".OPTIMIZER_SLOT" and parts[4] == "optimizer" and parts[5] in ("m", "v"):
                out[key] = f"adam_{parts[5]}/" + name
        elif parts == ["optimizer", "iter"]:
            out[key] = "step"
    return out


def convert(src: str, dst: str) -> dict:
    try:
        import tensorflow as tf
    except ImportError as e:  # pragma: no cover - needs the training machine
        raise SystemExit("convert_savedmodel needs TensorFlow to read the reference checkpoint; run it where the model "
                         "was trained (this package's runtime does not depend on TensorFlow)") from e
    reader = tf.train.load_checkpoint(os.path.join(src, "variables", "variables"))
    mapping = map_checkpoint_keys(reader.get_variable_to_shape_map().keys())
    blob = {}
    for key, name in mapping.items():
        v = reader.get_tensor(key)
        blob[name] = np.asarray(v, dtype=np.int64).reshape(1) if name == "step" else np.asarray(v, dtype=np.float32)
    missing = [f"{layer}/{p}" for layer in ATTR_TO_LAYER.values() for p in ("kernel", "bias") if f"{layer}/{p}" not in blob]
    if missing:
        raise SystemExit(f"checkpoint under {src} lacks {missing[:4]}... ({len(missing)} tensors)")
    blob["num_cards"] = np.int64(blob["encoder_e1/kernel"].shape[0])
    blob.setdefault("step", np.zeros(1, dtype=np.int64))
    blob["completed_epochs"] = np.int64(0)
    os.makedirs(dst, exist_ok=True)
    np.savez(os.path.join(dst, "cc_recommender.npz"), **blob)
    return blob


def main(argv=None):
    args = sys.argv[1:] if argv is None else argv
    if len(args) != 2:
        raise SystemExit(__doc__)
    blob = convert(args[0], args[1])
    print(f"wrote {args[1]}/cc_recommender.npz: {int(blob['num_cards'])} cards, "
          f"{sum(v.size for k, v in blob.items() if k.endswith(('kernel', 'bias')) and not k.startswith('adam'))} parameters")


if __name__ == "__main__":
    main()
