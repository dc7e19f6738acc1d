"""TensorBundle (SavedModel variables) reader / writer.  TensorFlow is absent, so reader and
writer are checked against each other, against the format constants, and through the model's save / load."""
import os
import struct

import numpy as np
import pytest

from kcvae_testlib import make, pkg, small_config
import importlib

tfb = importlib.import_module("trustedai-cl-vae-ad_b200.tf_bundle")


def test_crc32c_known_answers():
    assert tfb.crc32c(b"123456789") == 0xE3069283                 # the CRC-32C check value
    assert tfb.crc32c(bytes(32)) == 0x8A9136AA                    # RFC 3720 B.4: 32 zero bytes
    assert tfb.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43           # RFC 3720 B.4: 32 bytes of 0xff
    assert tfb.mask_crc(0) == 0xA282EAD8


def test_bundle_roundtrip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    t = {f"layer_with_weights-{i}/{n}/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal(s).astype(np.float32)
         for i, (n, s) in enumerate([("kernel", (3, 3, 3, 8)), ("bias", (8,)), ("kernel", (40, 5)), ("bias", (5,))] * 6)}
    t["save_counter/.ATTRIBUTES/VARIABLE_VALUE"] = np.array(7, np.int64)
    prefix = str(tmp_path / "variables" / "variables")
    tfb.write_bundle(prefix, t)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack_from("<Q", raw, len(raw) - 8)[0] == 0xDB4775248B80FB57
    back = tfb.read_bundle(prefix)
    assert sorted(back) == sorted(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and np.array_equal(back[k], t[k]), k
    # a flipped byte in the data file is caught by the per-tensor checksum
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[10] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        tfb.read_bundle(prefix)


def test_model_loads_a_savedmodel_style_directory(tmp_path):
    """<log_dir>/{config.yml, encoder/variables/*, decoder/variables/*} as train.py:127-128 leaves it."""
    cfg = small_config()
    m, ws = make(cfg, "emu")
    n_enc = len(m.encoder.variables)
    pkg.save_config(cfg, str(tmp_path / "config.yml"))
    for name, part in (("encoder", ws[:n_enc]), ("decoder", ws[n_enc:])):
        os.makedirs(tmp_path / name)
        (tmp_path / name / "saved_model.pb").write_bytes(b"")     # present in a real SavedModel; never parsed here
        tfb.keras_weights_to_bundle(str(tmp_path / name / "variables" / "variables"), part)
    m2, _ = make(cfg, "emu", seed=99)
    m2.load_model(str(tmp_path))
    for a, b in zip(m2.get_weights(), ws):
        assert np.array_equal(a, b)
    # and the directories this runtime writes carry the same bundle next to weights.npz
    m2.encoder.save(str(tmp_path / "out_enc"))
    got = tfb.keras_weights_from_bundle(str(tmp_path / "out_enc" / "variables" / "variables"))
    for a, b in zip(got, ws[:n_enc]):
        assert np.array_equal(a, b)
    # shape mismatch (a checkpoint of another topology) is reported, not silently loaded
    m3, _ = make(small_config(layers=(4, 5)), "emu")
    os.remove(tmp_path / "out_enc" / "weights.npz")
    with pytest.raises(ValueError, match="do not match"):
        m3.encoder.load(str(tmp_path / "out_enc"))
